# round 2: DOUBLE record kernel with the scaled corners summed over the grids (one interpolation, three divisions per atom)
python -m pytest tests -m gpu -q 2>&1 | tail -3
python tools/r2_perf.py double > gpurun_out/r2f64_perf.log 2>&1; grep "C5 double" gpurun_out/r2f64_perf.log
python tools/profile_run.py c5full 1 2 8 1
ncu --set full --clock-control none --cache-control none --import-source on -k regex:gf_eval_lines_f64_kernel -s 5 -c 1 -f -o gpurun_out/r2b_c5full_double python tools/profile_run.py c5full 1 2 8 1 > gpurun_out/r2b_ncu_c5full_double.log 2>&1
ncu -i gpurun_out/r2b_c5full_double.ncu-rep --page raw --csv > gpurun_out/r2b_c5full_double_raw.csv
rm -f gpurun_out/r2b_c5full_double.ncu-rep
