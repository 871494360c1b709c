# round 2, session 2: DOUBLE B-spline record kernel (gf_eval_bspline_f64_kernel) — parity, then timing against the general kernel
python -m pytest tests/test_gpu_bspline.py tests/test_gpu_modes.py tests/test_plugin.py -m gpu -q -x 2>&1 | tail -6 > gpurun_out/r2v_tests.log
tail -3 gpurun_out/r2v_tests.log
python tools/bspline_perf.py 65536 > gpurun_out/r2v_bspline_perf.log 2>&1
GFB_BSPLINE_F64=0 python tools/bspline_perf.py 65536 > gpurun_out/r2v_bspline_perf_general.log 2>&1
grep "bspline double" gpurun_out/r2v_bspline_perf.log gpurun_out/r2v_bspline_perf_general.log; grep "bspline mixed" gpurun_out/r2v_bspline_perf.log
