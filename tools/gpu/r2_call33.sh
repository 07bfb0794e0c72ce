#!/bin/bash
# round 2, call 33: source-level ncu view of k_window_tc (stall and bank-conflict attribution)
mkdir -p gpurun_out
PROF="python tools/profile_step.py"
$PROF > gpurun_out/c33_plain_prof.log 2>&1; echo "profile_step plain rc=$?"
k=k_window_tc
ncu --set full --clock-control none --import-source on -k "regex:$k" -s 2 -c 1 -f -o /tmp/prof_$k $PROF > gpurun_out/c33_ncu_$k.log 2>&1
echo "ncu $k rc=$?"
ncu -i /tmp/prof_$k.ncu-rep --page source --csv --print-source sass > /tmp/src.csv 2> gpurun_out/c33_src.err; echo "source rc=$?"
gzip -c /tmp/src.csv > gpurun_out/${k}_source_r02.csv.gz
ncu -i /tmp/prof_$k.ncu-rep --page source --csv --print-source cuda,sass > /tmp/src2.csv 2>> gpurun_out/c33_src.err; echo "source2 rc=$?"
gzip -c /tmp/src2.csv > gpurun_out/${k}_source_cuda_r02.csv.gz
rm -f /tmp/prof_$k.ncu-rep
du -sh gpurun_out
