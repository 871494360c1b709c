python -m pytest tests -m gpu -q 2>&1 | tail -40 > gpurun_out/r2c_tests.log
python tools/r2_perf.py modes strong c4 c3 > gpurun_out/r2c_perf_default.log 2>&1
GFB_POS_PREFETCH=0 python tools/r2_perf.py modes strong c4 c3 > gpurun_out/r2c_perf_noprefetch.log 2>&1
tail -5 gpurun_out/r2c_tests.log
