set -e
for c in n3_fixed n5_v2 n10_grp n3_safe; do python tools/profile_cases.py $c > gpurun_out/plain_$c.log 2>&1; cat gpurun_out/plain_$c.log; done
for c in n3_fixed n5_v2 n10_grp n3_safe; do
  ncu --set full --clock-control none --import-source on -k regex:rollout_kernel -s 2 -c 1 -f -o gpurun_out/r01c_$c python tools/profile_cases.py $c 1 > gpurun_out/ncu_$c.log 2>&1 || echo "ncu failed $c"
done
python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/plain_bench.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01c_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/ncu_bench.log 2>&1
ls -la gpurun_out
