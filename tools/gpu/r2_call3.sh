#!/bin/bash
# round 2, call 3: suite with kpar0 + node-set selection, bench at the tolerance-driven node count (grid leg, parity on the
# filled measurement, n_q = 88 alongside), PP choice for a 128-node set
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -q -x ) > gpurun_out/c3_suite.log 2>&1
echo "gpu suite rc=$?"
timeout 900 python bench.py --steps 6 --warmup 3 > gpurun_out/c3_bench.json 2> gpurun_out/c3_bench.err
echo "bench rc=$?"
CPX_PP=1 timeout 600 python bench.py --steps 4 --warmup 3 --n-q 128 --alt-n-q 0 --no-grid --no-cpu-baseline --no-e2e > gpurun_out/c3_bench_nq128_pp1.json 2> gpurun_out/c3_bench_nq128_pp1.err
echo "bench nq128 pp1 rc=$?"
timeout 600 python bench.py --steps 4 --warmup 3 --n-q 128 --alt-n-q 0 --no-grid --no-cpu-baseline --no-e2e > gpurun_out/c3_bench_nq128_pp2.json 2> gpurun_out/c3_bench_nq128_pp2.err
echo "bench nq128 pp2 rc=$?"
tail -n 15 gpurun_out/c3_suite.log
tail -n 5 gpurun_out/c3_bench.err
python - <<'PY'
import json
try:
    j=json.load(open("gpurun_out/c3_bench.json"))
    print("n_q", j["config"]["n_q"], "value %.4g e2e %.4g ms/step %.2f" % (j["value"], j["e2e"]["value"], j["ms_per_step"]), j["roofline"]["stage_ms"], "frac %.3f" % j["roofline"]["frac"])
    print("parity", json.dumps(j["parity"]))
    g=j["grid"]; print("grid value %.4g e2e %.4g ms %.2f" % (g["value"], g["e2e"]["value"], g["ms_per_scan"]), g["stage_ms"], g.get("parity"))
    print("alt", j["alt_n_q"]); print("setup_s", j["setup_s"]); print("cpu", j["cpu_baseline"]["value"])
except Exception as e:
    print("bench parse failed", e)
for n in ("pp1","pp2"):
    try:
        j=json.load(open("gpurun_out/c3_bench_nq128_%s.json"%n))
        print("nq128", n, "value %.4g ms/step %.2f" % (j["value"], j["ms_per_step"]), j["roofline"]["stage_ms"])
    except Exception as e:
        print(n, "failed", e)
PY
