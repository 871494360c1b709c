# Round-2 final ncu recipe (run under gpurun, one GPU): launch list of the bench command + one full capture per workload.
# Every ncu command runs after the same command has exited 0 without ncu. pdl = 1 mirrors the bench's launches (small
# launches then run the tile-striding instantiation <..., PERSIST = true>).
set -x
python bench.py --steps 20 --warmup 3 --no-extras > gpurun_out/r2b_bench_noextras.json 2> gpurun_out/r2b_bench_noextras.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r2b_launches_bench.csv \
    python bench.py --steps 20 --warmup 3 --no-extras > gpurun_out/r2b_ncu_launch.log 2>&1
run_cfg() {  # name cfg prec fmode kernel-regex
  python tools/profile_run.py $2 $3 $4 8 1 || return 1
  ncu --set full --clock-control none --cache-control none --import-source on -k regex:$5 -s 5 -c 1 -f -o gpurun_out/r2b_$1 \
      python tools/profile_run.py $2 $3 $4 8 1 > gpurun_out/r2b_ncu_$1.log 2>&1
  ncu -i gpurun_out/r2b_$1.ncu-rep --page raw --csv > gpurun_out/r2b_$1_raw.csv
}
run_cfg c5full c5full 0 2 gf_eval_lines_kernel
run_cfg c5shard8 c5shard8 0 2 gf_eval_lines_kernel
run_cfg c5shard4 c5shard4 0 2 gf_eval_lines_kernel
run_cfg c3 c3 0 2 gf_eval_lines_kernel
run_cfg c4 c4 0 2 gf_eval_lines_kernel
run_cfg c5full_energy_only c5full 0 -1 gf_eval_lines_kernel
run_cfg c5full_double c5full 1 2 gf_eval_lines_f64_kernel
ncu -i gpurun_out/r2b_c5shard8.ncu-rep --page source --csv > gpurun_out/r2b_c5shard8_source.csv 2>/dev/null
ncu -i gpurun_out/r2b_c5shard8.ncu-rep --page details --csv > gpurun_out/r2b_c5shard8_details.csv 2>/dev/null
rm -f gpurun_out/r2b_c3.ncu-rep gpurun_out/r2b_c4.ncu-rep gpurun_out/r2b_c5full_energy_only.ncu-rep gpurun_out/r2b_c5full_double.ncu-rep gpurun_out/r2b_c5shard4.ncu-rep gpurun_out/r2b_c5full.ncu-rep
ls -la gpurun_out | tail -30
