#!/bin/bash
# round 2, call 19: bench line + ncu of the final factorised-path kernels, summarised on the box (reports are deleted)
mkdir -p gpurun_out
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/c19_bench.json 2> gpurun_out/c19_bench.err; echo "bench rc=$?"
PROF="python tools/profile_step.py"
$PROF > gpurun_out/c19_plain_prof.log 2>&1; echo "profile_step plain rc=$?"
$PROF > /dev/null 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_prof_r02c.csv $PROF > gpurun_out/c19_ncu_list.log 2>&1; echo "list rc=$?"
for k in k_window_gram k_quad k_basis_f64; do
  $PROF > /dev/null 2>&1 && ncu --set full --clock-control none --import-source on -k "regex:$k" -s 1 -c 1 -f -o /tmp/prof_$k $PROF > gpurun_out/c19_ncu_$k.log 2>&1
  echo "ncu $k rc=$?"
  python tools/summarize_ncu.py /tmp/prof_$k.ncu-rep gpurun_out/${k}_ncu_r02c.txt > /dev/null 2>&1; echo "summary $k rc=$?"
  rm -f /tmp/prof_$k.ncu-rep
done
du -sh gpurun_out
