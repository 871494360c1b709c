# round 2, session 2: C3 (one replica of 1M atoms, one grid) with the tile-striding variant: 6 / 5 / 4 blocks of 256 per SM
python -m pytest tests/test_gpu_lines.py tests/test_gpu_modes.py tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -q -x 2>&1 | tail -3 > gpurun_out/r2r_tests.log
GFB_DEFER=0 python tools/r2_perf.py c3 > gpurun_out/r2r_c3_nodefer.log 2>&1
python tools/r2_perf.py c3 > gpurun_out/r2r_c3_ng1_6.log 2>&1
for b in 5 4; do GFB_LIB_PATH=ab/libgf_ng1_$b.so python tools/r2_perf.py c3 > gpurun_out/r2r_c3_ng1_$b.log 2>&1; done
tail -2 gpurun_out/r2r_tests.log
grep "pdl=1" gpurun_out/r2r_c3_*.log
