N=${1:-8}
nvidia-smi topo -m > gpurun_out/r2g_topo_n$N.txt 2>&1; nproc >> gpurun_out/r2g_topo_n$N.txt
for g in push nccl; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 3 --energy-gather $g > gpurun_out/r2g_bench_n${N}_$g.json 2> gpurun_out/r2g_bench_n${N}_$g.err
done
tail -c 600 gpurun_out/r2g_bench_n${N}_push.err
