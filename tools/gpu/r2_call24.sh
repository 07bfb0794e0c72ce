#!/bin/bash
# round 2, call 24: ncu of k_p3d_hankel_imma (full set + source page), summarised on the box
mkdir -p gpurun_out
PROF="python tools/profile_step.py"
$PROF > gpurun_out/c24_plain_prof.log 2>&1; echo "profile_step plain rc=$?"
k=k_p3d_hankel_imma
ncu --set full --clock-control none --import-source on -k "regex:$k" -s 2 -c 1 -f -o /tmp/prof_$k $PROF > gpurun_out/c24_ncu_$k.log 2>&1
echo "ncu $k rc=$?"
python tools/summarize_ncu.py /tmp/prof_$k.ncu-rep gpurun_out/${k}_ncu_r02.txt > /dev/null 2>&1; echo "summary rc=$?"
ncu -i /tmp/prof_$k.ncu-rep --page source --csv --print-source sass > /tmp/src.csv 2> gpurun_out/c24_src.err; echo "source rc=$?"
gzip -c /tmp/src.csv > gpurun_out/${k}_source_r02.csv.gz
rm -f /tmp/prof_$k.ncu-rep
du -sh gpurun_out
