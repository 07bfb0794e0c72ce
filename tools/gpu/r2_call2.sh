#!/bin/bash
# round 2, call 2: factorised path -- its tests, the whole GPU suite, a bench line with the grid scan
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests/test_gpu_factor.py -x -q ) > gpurun_out/c2_factor.log 2>&1
echo "factor tests rc=$?"
( time timeout 900 python -m pytest tests -m gpu -q ) > gpurun_out/c2_suite.log 2>&1
echo "gpu suite rc=$?"
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/c2_bench.json 2> gpurun_out/c2_bench.err
echo "bench rc=$?"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/c2_smoke.log 2>&1
echo "smoke rc=$?"
tail -n 25 gpurun_out/c2_factor.log
tail -n 12 gpurun_out/c2_suite.log
tail -n 5 gpurun_out/c2_bench.err
python - <<'PY'
import json
try:
    j=json.load(open("gpurun_out/c2_bench.json"))
    print("value %.4g e2e %.4g ms/step %.2f" % (j["value"], j["e2e"]["value"], j["ms_per_step"]))
    print("parity", json.dumps(j["parity"]))
    print("grid", json.dumps(j["grid"]))
    print("setup_s", j["setup_s"])
except Exception as e:
    print("bench parse failed", e)
PY
