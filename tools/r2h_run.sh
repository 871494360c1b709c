# round 2, session 2: tile-striding (persist) variant of the lines kernel under launch overlap — A/B against the one-tile-per-block launch
python -m pytest tests/test_gpu_lines.py tests/test_gpu_fullsize.py -m gpu -q -x 2>&1 | tail -15 > gpurun_out/r2h_tests.log
python tools/r2_perf.py strong c4 > gpurun_out/r2h_perf_persist.log 2>&1
GFB_DEFER=0 python tools/r2_perf.py strong c4 > gpurun_out/r2h_perf_nodefer.log 2>&1
GFB_PERSIST_MAX_WAVES=20 python tools/r2_perf.py modes strong > gpurun_out/r2h_perf_persist20.log 2>&1
GFB_LIB_PATH=ab/libgf_p1024.so python tools/r2_perf.py strong c4 > gpurun_out/r2h_perf_p1024.log 2>&1
tail -5 gpurun_out/r2h_tests.log
grep "shard 1/8" gpurun_out/r2h_perf_*.log
