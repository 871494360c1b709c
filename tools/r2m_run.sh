# round 2, session 2: pipeline depth of the tile-striding launch (grid = resident / D)
python -m pytest tests/test_gpu_lines.py tests/test_gpu_modes.py tests/test_gpu_fullsize.py -m gpu -q -x 2>&1 | tail -4 > gpurun_out/r2m_tests.log
python tools/r2_perf.py strong c4 > gpurun_out/r2m_perf_default.log 2>&1
GFB_PERSIST_MAX_DEPTH=1 python tools/r2_perf.py strong c4 > gpurun_out/r2m_perf_depth1.log 2>&1
GFB_PERSIST_MIN_TILES=4 python tools/r2_perf.py strong c4 > gpurun_out/r2m_perf_min4.log 2>&1
GFB_PERSIST_MIN_TILES=5.2 python tools/r2_perf.py strong c4 > gpurun_out/r2m_perf_min5.log 2>&1
GFB_PERSIST_MIN_TILES=7.7 GFB_PERSIST_MAX_WAVES=6 python tools/r2_perf.py strong c4 > gpurun_out/r2m_perf_min8.log 2>&1
tail -2 gpurun_out/r2m_tests.log
grep "shard 1/[48].*pdl=1 graph=1\|C4 pdl=1 graph=1 fixed\|C4 pdl=1 graph=1 energy" gpurun_out/r2m_perf_*.log
