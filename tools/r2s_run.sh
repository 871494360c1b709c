# round 2, session 2: the bench as the driver runs it at N=1 (both arms), plus the default invocation
python bench.py --impl reference --gpus 1 --steps 20 --warmup 3 > gpurun_out/r2s_ref.json 2> gpurun_out/r2s_ref.err
python bench.py --gpus 1 --steps 20 --warmup 3 > gpurun_out/r2s_bench_s20.json 2> gpurun_out/r2s_bench_s20.err
python bench.py > gpurun_out/r2s_bench_default.json 2> gpurun_out/r2s_bench_default.err
tail -c 300 gpurun_out/r2s_bench_s20.err; tail -c 600 gpurun_out/r2s_bench_default.json
