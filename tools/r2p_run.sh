# round 2, session 2: plain-specialised tile-striding variant at 32 / 36 / 40 warps per SM
python -m pytest tests/test_gpu_lines.py tests/test_gpu_modes.py -m gpu -q -x 2>&1 | tail -3 > gpurun_out/r2p_tests.log
python tools/r2_perf.py strong sweep c4 > gpurun_out/r2p_perf_default.log 2>&1
for v in t1152 t1280; do GFB_LIB_PATH=ab/libgf_$v.so python tools/r2_perf.py strong sweep c4 > gpurun_out/r2p_perf_$v.log 2>&1; done
for v in t1152 t1280; do GFB_PERSIST_MAX_WAVES=6 GFB_LIB_PATH=ab/libgf_$v.so python tools/r2_perf.py strong sweep > gpurun_out/r2p_perf_${v}_w6.log 2>&1; done
tail -2 gpurun_out/r2p_tests.log
grep "shard 1/[48].*pdl=1 graph=1\|C4 pdl=1 graph=1 fixed\|replicas per launch" gpurun_out/r2p_perf_*.log
