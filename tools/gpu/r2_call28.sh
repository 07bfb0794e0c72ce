#!/bin/bash
# round 2, call 28: cheap birthday pre-check of the explicit-point grouping -- factorised-path tests, bench (e2e should equal value again)
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_factor.py tests/test_gpu_scripts.py -q ) > gpurun_out/c28_tests.log 2>&1
echo "tests rc=$?"; tail -n 6 gpurun_out/c28_tests.log
( time timeout 1200 python bench.py ) > gpurun_out/c28_bench.json 2> gpurun_out/c28_bench.err; echo "bench rc=$?"; tail -n 3 gpurun_out/c28_bench.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/c28_bench.json").read().strip().splitlines()[-1])
print("value %.4g e2e %.4g ms/step %.2f frac %.3f grid %.4g (e2e %.4g) precise %.4g" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["roofline"]["frac"], d["grid"]["value"], d["grid"]["e2e"]["value"], d["precise"]["value"]))
PY
