#!/bin/bash
# round 2, call 12: fused window + Gram kernel: factor tests, bench (grid leg), A/B against the two-kernel form
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_factor.py tests/test_gpu_parity.py tests/test_gpu_configs.py tests/test_posterior_golden.py -m gpu -q ) > gpurun_out/c12_tests.log 2>&1
echo "tests rc=$?"
tail -n 15 gpurun_out/c12_tests.log
timeout 900 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-e2e --alt-n-q 0 --no-precise > gpurun_out/c12_bench.json 2> gpurun_out/c12_bench.err
echo "bench rc=$?"
CPX_FUSED_GRAM=0 timeout 900 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-e2e --alt-n-q 0 --no-precise > gpurun_out/c12_bench_unfused.json 2> gpurun_out/c12_bench_unfused.err
echo "bench unfused rc=$?"
python - <<'PY'
import json
for n in ("c12_bench","c12_bench_unfused"):
    j=json.load(open("gpurun_out/%s.json"%n))
    g=j["grid"]; print(n, "grid value %.4g e2e %.4g ms %.2f" % (g["value"], g["e2e"]["value"], g["ms_per_scan"]), g["stage_ms"])
PY
