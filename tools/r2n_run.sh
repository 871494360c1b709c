# round 2, session 2: depth of the tile-striding launch, rule D = floor(parked tiles x resident / tiles)
python tools/r2_perf.py strong sweep c4 > gpurun_out/r2n_perf_default.log 2>&1
GFB_DEFER=0 python tools/r2_perf.py sweep > gpurun_out/r2n_perf_nodefer.log 2>&1
for d in 2 3 4; do GFB_PERSIST_MAX_DEPTH=$d GFB_LIB_PATH=ab/libgf_d6.so python tools/r2_perf.py strong sweep c4 > gpurun_out/r2n_perf_d6_depth$d.log 2>&1; done
grep "shard 1/[48].*pdl=1 graph=1\|C4 pdl=1 graph=1 fixed\|replicas per launch" gpurun_out/r2n_perf_*.log
