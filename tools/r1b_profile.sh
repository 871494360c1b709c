set -x
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r1b_ref.json 2> gpurun_out/r1b_ref.err; echo "ref rc=$?"
python bench.py --steps 30 --warmup 3 --no-extras > gpurun_out/r1b_bench_noextras.json 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r1b_launches.csv python bench.py --steps 30 --warmup 3 --no-extras > gpurun_out/r1b_ncu_launch.log 2>&1
for cfg in c5full c3 c4; do
  python tools/profile_run.py $cfg 0 2 6 || exit 1
  ncu --set full --clock-control none --cache-control none --import-source on -k regex:gf_eval_lines -s 3 -c 1 -f -o gpurun_out/r1b_${cfg}_lines_warm python tools/profile_run.py $cfg 0 2 6 > gpurun_out/r1b_ncu_${cfg}.log 2>&1
  ncu -i gpurun_out/r1b_${cfg}_lines_warm.ncu-rep --page raw --csv > gpurun_out/r1b_${cfg}_lines_warm_raw.csv
done
ncu -i gpurun_out/r1b_c5full_lines_warm.ncu-rep --page source --csv > gpurun_out/r1b_c5full_lines_warm_source.csv 2>/dev/null
rm -f gpurun_out/r1b_c3_lines_warm.ncu-rep gpurun_out/r1b_c4_lines_warm.ncu-rep
ls -la gpurun_out
