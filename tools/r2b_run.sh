python -m pytest tests -m gpu -q 2>&1 | tail -40 > gpurun_out/r2b_tests.log
python tools/r2_perf.py modes strong c4 c3 > gpurun_out/r2b_perf_default.log 2>&1
GFB_LIB_PATH=$PWD/ab/libgf_b64.so python tools/r2_perf.py modes strong c4 > gpurun_out/r2b_perf_b64.log 2>&1
for v in 1 3; do GFB_LIB_PATH=$PWD/ab/libgf_ld$v.so python tools/r2_perf.py c3 > gpurun_out/r2b_perf_ld$v.log 2>&1; done
tail -5 gpurun_out/r2b_tests.log
