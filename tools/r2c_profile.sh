# Round-2 ncu captures of the record kernels of interpolation methods 1 and 2 on the C5 shape (run under gpurun, one GPU).
# Every ncu command runs after the same command has exited 0 without ncu.
set -x
run_cfg() {  # name layout
  python tools/profile_run.py c5full 0 2 8 0 $2 || return 1
  ncu --set full --clock-control none --cache-control none --import-source on -k regex:gf_eval_bspline_kernel -s 5 -c 1 -f -o gpurun_out/r2c_$1 \
      python tools/profile_run.py c5full 0 2 8 0 $2 > gpurun_out/r2c_ncu_$1.log 2>&1
  ncu -i gpurun_out/r2c_$1.ncu-rep --page raw --csv > gpurun_out/r2c_$1_raw.csv
  rm -f gpurun_out/r2c_$1.ncu-rep
}
run_cfg c5full_tricubic hermite
run_cfg c5full_bspline bspline
ls -la gpurun_out | tail -8
