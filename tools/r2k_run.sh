# round 2, session 2: persist variant with the occupancy-derived grid, register budgets 64 (default) / 56 / 48
python -m pytest tests/test_gpu_lines.py tests/test_gpu_modes.py -m gpu -q -x 2>&1 | tail -4 > gpurun_out/r2k_tests.log
python tools/r2_perf.py strong c4 > gpurun_out/r2k_perf_default.log 2>&1
for v in t1152 t1280; do GFB_LIB_PATH=ab/libgf_$v.so python tools/r2_perf.py strong c4 > gpurun_out/r2k_perf_$v.log 2>&1; done
tail -2 gpurun_out/r2k_tests.log
grep "shard 1/[48].*pdl=1 graph=1\|C4 pdl=1 graph=1 fixed\|C4 pdl=1 graph=1 energy" gpurun_out/r2k_perf_*.log
