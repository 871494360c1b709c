N=${1:-2}
python -m pytest tests/test_gpu_multi.py -m gpu -q -x 2>&1 | tail -30 > gpurun_out/r2f_tests_multi_n$N.log
for g in fused push nccl; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 3 --energy-gather $g > gpurun_out/r2f_bench_n${N}_$g.json 2> gpurun_out/r2f_bench_n${N}_$g.err
done
tail -3 gpurun_out/r2f_tests_multi_n$N.log; tail -c 800 gpurun_out/r2f_bench_n${N}_fused.err
