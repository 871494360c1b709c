# round 2, session 2 (N GPUs): the bench as the driver runs it
N=${1:-8}
nvidia-smi topo -m > gpurun_out/r2o_topo_n$N.txt 2>&1; nproc >> gpurun_out/r2o_topo_n$N.txt
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/r2o_bench_n${N}.json 2> gpurun_out/r2o_bench_n${N}.err
tail -c 600 gpurun_out/r2o_bench_n${N}.err
python - <<PY
import json
for line in open("gpurun_out/r2o_bench_n${N}.json"):
    if line.startswith("{"):
        d=json.loads(line); print("N", d["n_gpus"], "value %.1f G"%(d["value"]/1e9), "us/step %.2f"%(d["ms_per_step"]*1e3), "own launch %.2f"%d["roofline"]["launch_us"], {k:round(v["ms_per_step"]*1e3,2) for k,v in d["variants"].items()}, "e2e %.2f G"%(d["e2e"]["value"]/1e9), "weak", d.get("weak_scaling",{}).get("value"))
PY
