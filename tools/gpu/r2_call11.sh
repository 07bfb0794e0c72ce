#!/bin/bash
# round 2, call 11: whole suite (ns + SiIII, stream test, Gram cache), bench
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q ) > gpurun_out/c11_suite.log 2>&1
echo "gpu suite rc=$?"
tail -n 25 gpurun_out/c11_suite.log
timeout 900 python bench.py --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/c11_bench.json 2> gpurun_out/c11_bench.err
echo "bench rc=$?"
python - <<'PY'
import json
j=json.load(open("gpurun_out/c11_bench.json"))
print("value %.4g e2e %.4g ms/step %.2f" % (j["value"], j["e2e"]["value"], j["ms_per_step"]), j["roofline"]["stage_ms"])
g=j["grid"]; print("grid value %.4g e2e %.4g ms %.2f" % (g["value"], g["e2e"]["value"], g["ms_per_scan"]), g["stage_ms"])
PY
