# round 2, session 2: how far up in launch size the tile-striding variant pays (64-register default vs 56-register build)
GFB_PERSIST_MAX_WAVES=24 python tools/r2_perf.py modes strong sweep > gpurun_out/r2q_perf_default_w24.log 2>&1
GFB_PERSIST_MAX_WAVES=24 GFB_LIB_PATH=ab/libgf_t1152.so python tools/r2_perf.py modes strong sweep > gpurun_out/r2q_perf_t1152_w24.log 2>&1
python tools/r2_perf.py modes > gpurun_out/r2q_perf_default_modes.log 2>&1
grep "shard 1/[248].*pdl=1 graph=1\|C5 mixed path=1 pdl=1\|replicas per launch" gpurun_out/r2q_perf_*.log
