#!/bin/bash
# round 2, call 13: guard-band test of the factorised path; quick knobs of the general pipeline at n_q = 56
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_factor.py -m gpu -q ) > gpurun_out/c13_tests.log 2>&1
echo "tests rc=$?"
tail -n 12 gpurun_out/c13_tests.log
B="python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-e2e --alt-n-q 0 --no-precise --no-grid"
for v in "CPX_PP=1" "CPX_PP=2" "CPX_WTC_AGROUPS=2" "CPX_WTC_AGROUPS=5" "CPX_WTC_AGROUPS=10" "CPX_WTC_AGROUPS=20"; do
  env $v timeout 300 $B > gpurun_out/c13_$v.json 2> gpurun_out/c13_$v.err
  python - "$v" <<'PY'
import json, sys
v=sys.argv[1]
try:
    j=json.load(open("gpurun_out/c13_%s.json"%v)); print(v, "value %.4g ms/step %.2f" % (j["value"], j["ms_per_step"]), j["roofline"]["stage_ms"])
except Exception as e: print(v, "failed", e)
PY
done
