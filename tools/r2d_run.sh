python -m pytest tests -m gpu -q -x 2>&1 | tail -30 > gpurun_out/r2d_tests.log
python bench.py > gpurun_out/r2d_bench.json 2> gpurun_out/r2d_bench.err
python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/r2d_bench_ref.json 2> gpurun_out/r2d_bench_ref.err
tail -3 gpurun_out/r2d_tests.log; tail -c 600 gpurun_out/r2d_bench.err; nproc
