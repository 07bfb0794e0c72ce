#!/bin/bash
# round 2, call 27: whole GPU suite, smoke, bench (default run) on the final code
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q ) > gpurun_out/c27_suite.log 2>&1
echo "suite rc=$?"; tail -n 12 gpurun_out/c27_suite.log
( time python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/c27_smoke.log 2>&1; echo "smoke rc=$?"; tail -n 6 gpurun_out/c27_smoke.log
( time timeout 1200 python bench.py ) > gpurun_out/c27_bench.json 2> gpurun_out/c27_bench.err; echo "bench rc=$?"; tail -n 3 gpurun_out/c27_bench.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/c27_bench.json").read().strip().splitlines()[-1])
print("value %.4g e2e %.4g ms/step %.2f frac %.3f grid %.4g precise %.4g alt_n_q %.4g alt_bao %.4g setup_s %s" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["roofline"]["frac"], d["grid"]["value"], d["precise"]["value"], d["alt_n_q"]["value"], d["alt_bao"]["value"], d.get("setup_s")))
print("parity", d["parity"]["max_abs_dchi2"], d["parity"]["truth"], "cpu", d["cpu_baseline"]["value"], "clocks", d["clocks"])
PY
