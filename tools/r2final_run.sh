# round 2: final measurement batch on one GPU (after the last kernel change)
bash tools/r2b_profile.sh > gpurun_out/r2b_profile.log 2>&1
python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2s_ref.json 2> gpurun_out/r2s_ref.err
python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2s_bench_s20.json 2> gpurun_out/r2s_bench_s20.err
python bench.py > gpurun_out/r2s_bench_default.json 2> gpurun_out/r2s_bench_default.err
python tools/r2_perf.py modes strong sweep c4 c3 double > gpurun_out/r2f_perf_default.log 2>&1
GFB_DEFER=0 python tools/r2_perf.py strong sweep c4 > gpurun_out/r2f_perf_nodefer.log 2>&1
GFB_PERSIST_MAX_WAVES=24 python tools/r2_perf.py modes strong sweep > gpurun_out/r2f_perf_w24.log 2>&1
grep "us/launch" gpurun_out/r2b_profile.log; tail -c 300 gpurun_out/r2s_bench_s20.err
