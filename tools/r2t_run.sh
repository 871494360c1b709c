# round 2, session 2 (2 GPUs): flag-in-data gather (gfb_comm_gather) against push+wait
python -m pytest tests/test_gpu_multi.py -m gpu -q -x 2>&1 | tail -5 > gpurun_out/r2t_tests_multi_n2.log
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 "$@"; }
run > gpurun_out/r2t_bench_n2_ll.json 2> gpurun_out/r2t_bench_n2_ll.err
run --energy-gather push > gpurun_out/r2t_bench_n2_push.json 2> gpurun_out/r2t_bench_n2_push.err
tail -3 gpurun_out/r2t_tests_multi_n2.log; tail -c 400 gpurun_out/r2t_bench_n2_ll.err
python - <<'PY'
import json
for n in ("ll","push"):
    for line in open(f"gpurun_out/r2t_bench_n2_{n}.json"):
        if line.startswith("{"):
            d=json.loads(line); print(n, "us/step %.2f"%(d["ms_per_step"]*1e3), "own launch %.2f"%d["roofline"]["launch_us"], {k:round(v["ms_per_step"]*1e3,2) for k,v in d["variants"].items()}, "e2e %.2f G"%(d["e2e"]["value"]/1e9), d["run"]["gather_check"])
PY
