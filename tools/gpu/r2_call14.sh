#!/bin/bash
# round 2, call 14: k_quad rewrite: factor tests, whole suite, bench
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q ) > gpurun_out/c14_suite.log 2>&1
echo "suite rc=$?"
tail -n 8 gpurun_out/c14_suite.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/c14_bench.json 2> gpurun_out/c14_bench.err
echo "bench rc=$?"
tail -n 3 gpurun_out/c14_bench.err
python - <<'PY'
import json
j=json.load(open("gpurun_out/c14_bench.json"))
print("value %.4g e2e %.4g ms/step %.2f" % (j["value"], j["e2e"]["value"], j["ms_per_step"]), j["roofline"]["stage_ms"], "frac %.3f" % j["roofline"]["frac"])
g=j["grid"]; print("grid value %.4g e2e %.4g ms %.2f" % (g["value"], g["e2e"]["value"], g["ms_per_scan"]), g["stage_ms"], g["parity"])
print("alt", j["alt_n_q"]); print("alt_bao", j["alt_bao"]); print("precise %.4g" % j["precise"]["value"], "setup", j["setup_s"], "cpu", j["cpu_baseline"]["value"])
print("parity", j["parity"])
PY
