#!/bin/bash
# round 2, call 4: whole GPU suite (set-up on the device, mocks, posterior, sampler, IGM mode, components, scripts), bench
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q -x --durations=12 ) > gpurun_out/c4_suite.log 2>&1
echo "gpu suite rc=$?"
timeout 900 python bench.py --steps 8 --warmup 3 > gpurun_out/c4_bench.json 2> gpurun_out/c4_bench.err
echo "bench rc=$?"
( for s in "validate_pipeline.py" "chi2_scan_mock.py arinyo tight 0 bias beta --grid 128" "chi2_scan_forecast.py bias beta q1 kp_Mpc --grid 64" "fit_many_mocks.py bias beta --nmocks 1000"; do
    echo "== python scripts/$s"; ( time timeout 600 python scripts/$s ) 2>&1 | grep -v "^$" ; done ) > gpurun_out/c4_scripts.log 2>&1
echo "scripts rc=$?"
tail -n 40 gpurun_out/c4_suite.log
tail -n 5 gpurun_out/c4_bench.err
python - <<'PY'
import json
try:
    j=json.load(open("gpurun_out/c4_bench.json"))
    print("n_q", j["config"]["n_q"], "value %.4g e2e %.4g ms/step %.2f" % (j["value"], j["e2e"]["value"], j["ms_per_step"]), j["roofline"]["stage_ms"], "frac %.3f" % j["roofline"]["frac"])
    g=j["grid"]; print("grid value %.4g e2e %.4g ms %.2f" % (g["value"], g["e2e"]["value"], g["ms_per_scan"]), g["stage_ms"])
    print("setup_s", j["setup_s"])
except Exception as e:
    print("bench parse failed", e)
PY
cat gpurun_out/c4_scripts.log | tail -40
