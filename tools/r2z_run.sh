# round 2, session 2: scaled corners summed over the grids before ONE interpolation (lines kernel, 2-4 grids)
python -m pytest tests -m gpu -q -x 2>&1 | tail -4 > gpurun_out/r2z_tests.log; tail -3 gpurun_out/r2z_tests.log
python tools/r2_perf.py modes strong sweep c4 > gpurun_out/r2z_perf_combined.log 2>&1
grep "C5 mixed\|shard 1/[248].*pdl=1 graph=1\|C4 pdl=1 graph=1 fixed\|C4 pdl=1 graph=1 energy\|replicas per launch" gpurun_out/r2z_perf_combined.log
