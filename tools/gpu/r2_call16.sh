#!/bin/bash
# round 2, call 16: staged float64-cell kernel: suite, bench (precise record), A/B against the unstaged kernel
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q ) > gpurun_out/c16_suite.log 2>&1
echo "suite rc=$?"
tail -n 8 gpurun_out/c16_suite.log
B="python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-e2e --alt-n-q 0 --no-grid"
timeout 600 $B > gpurun_out/c16_bench.json 2> gpurun_out/c16_bench.err; echo "bench rc=$?"
CPX_F64_STAGE=0 timeout 600 $B > gpurun_out/c16_bench_nostage.json 2> gpurun_out/c16_bench_nostage.err; echo "bench nostage rc=$?"
python - <<'PY'
import json
for n in ("c16_bench","c16_bench_nostage"):
    j=json.load(open("gpurun_out/%s.json"%n)); print(n, "value %.4g" % j["value"], "precise", j["precise"]["value"], j["precise"]["ms_per_step"])
PY
