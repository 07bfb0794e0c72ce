#!/bin/bash
# round 2, call 1: new oracle-vs-GPU tests on the quoted configs; float32 Px_Zam scratch (CPX_PX32) parity and speed
mkdir -p gpurun_out
( time python -m pytest tests/test_gpu_oracle_configs.py -x -q -s ) > gpurun_out/c1_oracle_configs.log 2>&1
echo "oracle configs rc=$?"
( time CPX_PX32=1 python -m pytest tests -m gpu -x -q ) > gpurun_out/c1_px32_suite.log 2>&1
echo "px32 suite rc=$?"
( CPX_PX32=1 python -m pytest tests/test_gpu_oracle_configs.py -x -q -s ) > gpurun_out/c1_px32_oracle_configs.log 2>&1
echo "px32 oracle configs rc=$?"
python bench.py --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/c1_bench_px64.json 2> gpurun_out/c1_bench_px64.err
echo "bench px64 rc=$?"
CPX_PX32=1 python bench.py --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/c1_bench_px32.json 2> gpurun_out/c1_bench_px32.err
echo "bench px32 rc=$?"
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/c1_smoke.log 2>&1
echo "smoke rc=$?"
tail -3 gpurun_out/c1_oracle_configs.log gpurun_out/c1_px32_suite.log gpurun_out/c1_px32_oracle_configs.log gpurun_out/c1_smoke.log
python - <<'PY'
import json
for n in ("px64","px32"):
    try:
        j=json.load(open("gpurun_out/c1_bench_%s.json"%n))
        print(n, "value %.4g e2e %.4g ms/step %.2f stage %s" % (j["value"], j["e2e"]["value"], j["ms_per_step"], j["roofline"]["stage_ms"]))
    except Exception as e:
        print(n, "failed", e)
PY
