# round 2, session 2: C3 (one grid, 4.4 waves of 256-thread blocks) with the tile-striding variant, 6 / 5 / 4 blocks per SM
python tools/r2_perf.py c3 > gpurun_out/r2l_c3_base.log 2>&1
GFB_PERSIST_MAX_WAVES=6 python tools/r2_perf.py c3 > gpurun_out/r2l_c3_persist6.log 2>&1
GFB_PERSIST_MAX_WAVES=8 GFB_LIB_PATH=ab/libgf_ng1_5.so python tools/r2_perf.py c3 > gpurun_out/r2l_c3_persist5.log 2>&1
GFB_PERSIST_MAX_WAVES=8 GFB_LIB_PATH=ab/libgf_ng1_4.so python tools/r2_perf.py c3 > gpurun_out/r2l_c3_persist4.log 2>&1
grep "pdl=1" gpurun_out/r2l_c3_*.log
