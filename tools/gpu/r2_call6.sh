#!/bin/bash
# round 2, call 6: suite after the factorised-path kernel rewrites, bench, quick A/B of the basis staging
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q --durations=5 ) > gpurun_out/c6_suite.log 2>&1
echo "gpu suite rc=$?"
tail -n 12 gpurun_out/c6_suite.log
timeout 900 python bench.py --steps 8 --warmup 3 > gpurun_out/c6_bench.json 2> gpurun_out/c6_bench.err
echo "bench rc=$?"
CPX_BASIS_STAGE=0 timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --alt-n-q 0 > gpurun_out/c6_bench_nostage.json 2> gpurun_out/c6_bench_nostage.err
echo "bench nostage rc=$?"
tail -n 5 gpurun_out/c6_bench.err
python - <<'PY'
import json
for n in ("c6_bench","c6_bench_nostage"):
    try:
        j=json.load(open("gpurun_out/%s.json"%n))
        print(n, "value %.4g ms/step %.2f" % (j["value"], j["ms_per_step"]), j["roofline"]["stage_ms"])
        g=j["grid"]; print("  grid value %.4g e2e %.4g ms %.2f" % (g["value"], g["e2e"]["value"], g["ms_per_scan"]), g["stage_ms"], g.get("parity"))
        print("  setup_s", j["setup_s"], "parity", j.get("parity"))
    except Exception as e:
        print(n, "parse failed", e)
PY
