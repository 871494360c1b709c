nvidia-smi topo -m > gpurun_out/r2e_topo.txt 2>&1
python -m pytest tests/test_gpu_multi.py -m gpu -q -x 2>&1 | tail -30 > gpurun_out/r2e_tests_multi.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r2e_bench_n2.json 2> gpurun_out/r2e_bench_n2.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 3 --energy-gather nccl > gpurun_out/r2e_bench_n2_nccl.json 2> gpurun_out/r2e_bench_n2_nccl.err
python bench.py --steps 20 --warmup 3 --no-extras > gpurun_out/r2e_bench_n1.json 2> gpurun_out/r2e_bench_n1.err
tail -3 gpurun_out/r2e_tests_multi.log; tail -c 1500 gpurun_out/r2e_bench_n2.err
