#!/bin/bash
# round 2, call 38: final code (wave-aware chunk size): suite, smoke, bench, then the ncu evidence of the same code
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q ) > gpurun_out/c38_suite.log 2>&1
echo "suite rc=$?"; tail -n 6 gpurun_out/c38_suite.log
( time python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/c38_smoke.log 2>&1; echo "smoke rc=$?"; tail -n 4 gpurun_out/c38_smoke.log
( time timeout 1200 python bench.py ) > gpurun_out/c38_bench.json 2> gpurun_out/c38_bench.err; echo "bench rc=$?"; tail -n 3 gpurun_out/c38_bench.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/c38_bench.json").read().strip().splitlines()[-1])
print("value %.4g e2e %.4g ms/step %.2f frac %.3f grid %.4g precise %.4g alt_n_q %.4g alt_bao %.4g launches %s" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["roofline"]["frac"], d["grid"]["value"], d["precise"]["value"], d["alt_n_q"]["value"], d["alt_bao"]["value"], d["gpu_launches"]))
print("stage", d["roofline"]["stage_ms"], "parity", d["parity"]["max_abs_dchi2"], d["parity"]["truth"]["abs_dchi2"])
PY
BENCH="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --alt-n-q 0"
$BENCH > gpurun_out/c38_plain_bench.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_r02_final.csv $BENCH > gpurun_out/c38_ncu_bench.log 2>&1
echo "launch list rc=$?"
python tools/summarize_launches.py gpurun_out/launches_r02_final.csv gpurun_out/launches_r02_final_summary.txt "$BENCH" > /dev/null; echo "summary rc=$?"
PROF="python tools/profile_step.py"
$PROF > gpurun_out/c38_plain_prof.log 2>&1; echo "profile_step plain rc=$?"
for k in k_p3d_hankel k_window_tc; do
  ncu --set full --clock-control none --import-source on -k "regex:$k" -s 2 -c 1 -f -o /tmp/prof_$k $PROF > gpurun_out/c38_ncu_$k.log 2>&1
  echo "ncu $k rc=$?"
  python tools/summarize_ncu.py /tmp/prof_$k.ncu-rep gpurun_out/${k}_ncu_r02_final.txt gpurun_out/${k}_traffic_final.json > /dev/null 2>&1; echo "summary $k rc=$?"
  rm -f /tmp/prof_$k.ncu-rep
done
du -sh gpurun_out
