#!/bin/bash
# round 2, call 18: bench line + ncu of the final factorised-path kernels (call 17's outputs exceeded the copy-back limit)
mkdir -p gpurun_out
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/c18_bench.json 2> gpurun_out/c18_bench.err; echo "bench rc=$?"
PROF="python tools/profile_step.py"
$PROF > gpurun_out/c18_plain_prof.log 2>&1; echo "profile_step plain rc=$?"
$PROF > /dev/null 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_prof_r02c.csv $PROF > gpurun_out/c18_ncu_list.log 2>&1; echo "list rc=$?"
for k in k_window_gram k_quad k_basis_f64 k_p3d_hankel_f64s; do
  $PROF > /dev/null 2>&1 && ncu --set full --clock-control none -k "regex:$k" -s 1 -c 1 -f -o gpurun_out/prof_r02c_$k $PROF > gpurun_out/c18_ncu_$k.log 2>&1
  echo "ncu $k rc=$?"
done
ls -la gpurun_out/*.ncu-rep
