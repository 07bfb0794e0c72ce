#!/bin/bash
# round 2, call 32: final-code ncu evidence -- launch list of the bench command, full captures of the two hot kernels (52 nodes)
mkdir -p gpurun_out
BENCH="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --alt-n-q 0"
$BENCH > gpurun_out/c32_plain_bench.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_r02_final.csv $BENCH > gpurun_out/c32_ncu_bench.log 2>&1
echo "launch list rc=$?"
python tools/summarize_launches.py gpurun_out/launches_r02_final.csv gpurun_out/launches_r02_final_summary.txt "$BENCH" > /dev/null; echo "summary rc=$?"
PROF="python tools/profile_step.py"
$PROF > gpurun_out/c32_plain_prof.log 2>&1; echo "profile_step plain rc=$?"
for k in k_p3d_hankel k_window_tc; do
  ncu --set full --clock-control none --import-source on -k "regex:$k" -s 2 -c 1 -f -o /tmp/prof_$k $PROF > gpurun_out/c32_ncu_$k.log 2>&1
  echo "ncu $k rc=$?"
  python tools/summarize_ncu.py /tmp/prof_$k.ncu-rep gpurun_out/${k}_ncu_r02_final.txt gpurun_out/${k}_traffic_final.json > /dev/null 2>&1; echo "summary $k rc=$?"
  rm -f /tmp/prof_$k.ncu-rep
done
du -sh gpurun_out
