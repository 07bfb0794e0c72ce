#!/bin/bash
# round 2, call 5: rest of the GPU suite, then ncu: launch list of the bench command, full captures of the hot kernels
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q --durations=5 ) > gpurun_out/c5_suite.log 2>&1
echo "gpu suite rc=$?"
tail -n 15 gpurun_out/c5_suite.log
BENCH="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --alt-n-q 0"
$BENCH > gpurun_out/c5_plain_bench.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_r02.csv $BENCH > gpurun_out/c5_ncu_bench.log 2>&1
echo "launch list rc=$?"
PROF="python tools/profile_step.py"
$PROF > gpurun_out/c5_plain_prof.log 2>&1
echo "profile_step plain rc=$?"
tail -n 2 gpurun_out/c5_plain_prof.log
for k in k_p3d_hankel k_window_tc k_basis_f64 k_quad k_gram "k_window<"; do
  name=$(echo $k | tr -d '<')
  skip=1; [ "$k" = "k_window<" ] && skip=0
  $PROF > /dev/null 2>&1 && ncu --set full --clock-control none --import-source on -k "regex:$k" -s $skip -c 1 -f -o gpurun_out/prof_r02_$name $PROF > gpurun_out/c5_ncu_$name.log 2>&1
  echo "ncu $k rc=$?"
done
ls -la gpurun_out/*.ncu-rep
