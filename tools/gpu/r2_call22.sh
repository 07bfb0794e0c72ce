#!/bin/bash
# round 2, call 22: k_p3d_hankel_imma (INT8 mma.sync Hankel kernel) -- parity tests, then the whole suite, then bench A/B
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests/test_gpu_tc.py -x -q ) > gpurun_out/c22_tc.log 2>&1
echo "tc tests rc=$?"; tail -n 15 gpurun_out/c22_tc.log
( time timeout 1500 python -m pytest tests -m gpu -q ) > gpurun_out/c22_suite.log 2>&1
echo "suite rc=$?"; tail -n 25 gpurun_out/c22_suite.log
CPX_IMMA=0 timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/c22_bench_dmma.json 2> gpurun_out/c22_bench_dmma.err; echo "bench dmma rc=$?"
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/c22_bench_imma.json 2> gpurun_out/c22_bench_imma.err; echo "bench imma rc=$?"
python - <<'PY'
import json
for n in ("dmma", "imma"):
    try:
        d = json.loads(open("gpurun_out/c22_bench_%s.json" % n).read().strip().splitlines()[-1])
        print(n, "value %.4g e2e %.4g ms/step %.2f" % (d["value"], d["e2e"]["value"], d["ms_per_step"]), "stage_ms", d["roofline"].get("stage_ms"), "parity", d["parity"]["max_abs_dchi2"], d["parity"]["truth"]["abs_dchi2"], "alt_n_q", d.get("alt_n_q", {}).get("value"), "grid", d["grid"]["value"])
    except Exception as e:
        print(n, "failed", e)
PY
