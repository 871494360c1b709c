# round 2, session 2: GPU suite with the tile-striding variant as the default + A/B of its register budget / parked tiles
python -m pytest tests -m gpu -q -x 2>&1 | tail -8 > gpurun_out/r2i_tests.log
python tools/r2_perf.py strong c4 > gpurun_out/r2i_perf_default.log 2>&1
for v in d3 d3_896 d3_768; do GFB_LIB_PATH=ab/libgf_$v.so python tools/r2_perf.py strong c4 > gpurun_out/r2i_perf_$v.log 2>&1; done
tail -4 gpurun_out/r2i_tests.log
grep "shard 1/8.*pdl=1\|C4 pdl=1 graph=1 fixed\|C4 pdl=1 graph=1 energy" gpurun_out/r2i_perf_*.log
