#!/bin/bash
# round 2, call 8: factor-path tests after the second kernel pass, bench, basis-kernel warps A/B
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_factor.py tests/test_gpu_configs.py tests/test_gpu_parity.py tests/test_gpu_scripts.py tests/test_posterior_golden.py -m gpu -q ) > gpurun_out/c8_suite.log 2>&1
echo "tests rc=$?"
tail -n 6 gpurun_out/c8_suite.log
timeout 900 python bench.py --steps 8 --warmup 3 > gpurun_out/c8_bench.json 2> gpurun_out/c8_bench.err
echo "bench rc=$?"
CPX_BASIS_WARPS=8 timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --alt-n-q 0 --no-precise > gpurun_out/c8_bench_w8.json 2> gpurun_out/c8_bench_w8.err
echo "bench w8 rc=$?"
tail -n 5 gpurun_out/c8_bench.err
python - <<'PY'
import json
for n in ("c8_bench","c8_bench_w8"):
    try:
        j=json.load(open("gpurun_out/%s.json"%n))
        print(n, "value %.4g e2e %s ms/step %.2f" % (j["value"], j["e2e"] and "%.4g" % j["e2e"]["value"], j["ms_per_step"]), j["roofline"]["stage_ms"])
        g=j["grid"]; print("  grid value %.4g e2e %.4g ms %.2f" % (g["value"], g["e2e"]["value"], g["ms_per_scan"]), g["stage_ms"], g.get("parity"))
        print("  precise", j.get("precise"))
    except Exception as e:
        print(n, "parse failed", e)
PY
