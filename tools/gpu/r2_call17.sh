#!/bin/bash
# round 2, call 17: suite, smoke, parity report, bench; then ncu of the final factorised-path kernels
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q ) > gpurun_out/c17_suite.log 2>&1
echo "suite rc=$?"
tail -n 6 gpurun_out/c17_suite.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/c17_smoke.log 2>&1; echo "smoke rc=$?"; tail -n 5 gpurun_out/c17_smoke.log
timeout 600 python tools/parity_report.py > gpurun_out/parity_report_r02.txt 2> gpurun_out/c17_parity.err; echo "parity report rc=$?"; cat gpurun_out/parity_report_r02.txt
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/c17_bench.json 2> gpurun_out/c17_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
j=json.load(open("gpurun_out/c17_bench.json"))
print("value %.4g e2e %.4g ms/step %.2f" % (j["value"], j["e2e"]["value"], j["ms_per_step"]), j["roofline"]["stage_ms"], "frac %.3f" % j["roofline"]["frac"])
g=j["grid"]; print("grid value %.4g e2e %.4g ms %.2f" % (g["value"], g["e2e"]["value"], g["ms_per_scan"]), g["stage_ms"], g["parity"]["max_abs_dchi2"])
print("precise %.4g" % j["precise"]["value"], j["precise"]["parity"]["max_abs_dchi2"], "alt_bao %.4g" % j["alt_bao"]["value"])
PY
PROF="python tools/profile_step.py"
$PROF > gpurun_out/c17_plain_prof.log 2>&1; echo "profile_step plain rc=$?"
$PROF > /dev/null 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_prof_r02c.csv $PROF > gpurun_out/c17_ncu_list.log 2>&1; echo "list rc=$?"
for k in k_window_gram k_quad k_basis_f64; do
  $PROF > /dev/null 2>&1 && ncu --set full --clock-control none --import-source on -k "regex:$k" -s 1 -c 1 -f -o gpurun_out/prof_r02c_$k $PROF > gpurun_out/c17_ncu_$k.log 2>&1
  echo "ncu $k rc=$?"
done
