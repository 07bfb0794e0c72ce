#!/bin/bash
# round 2, call 9: suite, bench, smoke after the Gram cache / k_gram v3
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q ) > gpurun_out/c9_suite.log 2>&1
echo "gpu suite rc=$?"
tail -n 8 gpurun_out/c9_suite.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/c9_bench.json 2> gpurun_out/c9_bench.err
echo "bench rc=$?"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/c9_smoke.log 2>&1
echo "smoke rc=$?"
tail -n 4 gpurun_out/c9_smoke.log
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/c9_bench_ref.json 2> gpurun_out/c9_bench_ref.err
echo "bench ref rc=$?"
python - <<'PY'
import json
j=json.load(open("gpurun_out/c9_bench.json"))
print("value %.4g e2e %.4g ms/step %.2f" % (j["value"], j["e2e"]["value"], j["ms_per_step"]), j["roofline"]["stage_ms"], "frac %.3f" % j["roofline"]["frac"])
g=j["grid"]; print("grid value %.4g e2e %.4g ms %.2f" % (g["value"], g["e2e"]["value"], g["ms_per_scan"]), g["stage_ms"])
print("precise %.4g" % j["precise"]["value"], "alt", j["alt_n_q"]["value"], "cpu", j["cpu_baseline"]["value"], "setup", j["setup_s"])
r=json.load(open("gpurun_out/c9_bench_ref.json")); print("ref", r["value"], r["cpu_baseline"]["cores"])
PY
