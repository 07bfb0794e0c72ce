#!/bin/bash
# round 2, call 43: ncu capture of k_p3d_hankel on the final code (16-byte RowPar)
mkdir -p gpurun_out
PROF="python tools/profile_step.py"
k=k_p3d_hankel
ncu --set full --clock-control none --import-source on -k "regex:$k" -s 2 -c 1 -f -o /tmp/prof_$k $PROF > gpurun_out/c43_ncu_$k.log 2>&1
echo "ncu $k rc=$?"
python tools/summarize_ncu.py /tmp/prof_$k.ncu-rep gpurun_out/${k}_ncu_r02_final.txt gpurun_out/${k}_traffic_final.json > /dev/null 2>&1; echo "summary $k rc=$?"
rm -f /tmp/prof_$k.ncu-rep
