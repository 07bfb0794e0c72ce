#!/bin/bash
# round 2, call 7: ncu of the factorised-path kernels after the rewrite
mkdir -p gpurun_out
PROF="python tools/profile_step.py"
$PROF > gpurun_out/c7_plain_prof.log 2>&1
echo "profile_step plain rc=$?"
$PROF > /dev/null 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_prof_r02b.csv $PROF > gpurun_out/c7_ncu_list.log 2>&1
echo "list rc=$?"
for k in k_gram k_basis_f64 "k_window"; do
  $PROF > /dev/null 2>&1 && ncu --set full --clock-control none --import-source on -k "regex:^$k" -s 1 -c 1 -f -o gpurun_out/prof_r02b_$k $PROF > gpurun_out/c7_ncu_$k.log 2>&1
  echo "ncu $k rc=$?"
done
ls -la gpurun_out/prof_r02b*
