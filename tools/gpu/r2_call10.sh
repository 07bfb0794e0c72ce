#!/bin/bash
# round 2, call 10 (2 GPUs): the NCCL scan test, the factor test (Gram cache), the bench under torchrun at N = 2
mkdir -p gpurun_out
nvidia-smi -L
( time timeout 900 python -m pytest tests/test_gpu_scripts.py tests/test_gpu_factor.py -m gpu -q ) > gpurun_out/c10_tests.log 2>&1
echo "tests rc=$?"
tail -n 6 gpurun_out/c10_tests.log
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 6 --warmup 3 > gpurun_out/c10_bench_2gpu.json 2> gpurun_out/c10_bench_2gpu.err
echo "bench 2gpu rc=$?"
tail -n 5 gpurun_out/c10_bench_2gpu.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/c10_bench_ref_2gpu.json 2> gpurun_out/c10_bench_ref_2gpu.err
echo "bench ref 2gpu rc=$?"
( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29535 scripts/chi2_scan_forecast.py bias beta q1 kp_Mpc --grid 64 ) > gpurun_out/c10_scan_2gpu.log 2>&1
echo "scan 2gpu rc=$?"
tail -n 3 gpurun_out/c10_scan_2gpu.log
python - <<'PY'
import json
j=json.load(open("gpurun_out/c10_bench_2gpu.json"))
print("N=2 value %.4g e2e %.4g ms/step %.2f" % (j["value"], j["e2e"]["value"], j["ms_per_step"]), j["e2e"]["api"][:60])
g=j["grid"]; print("grid value %.4g e2e %.4g ms %.2f tuples/gpu %d" % (g["value"], g["e2e"]["value"], g["ms_per_scan"], g["tuples_per_scan_per_gpu"]), g["stage_ms"])
print(open("gpurun_out/c10_bench_ref_2gpu.json").read()[:300])
PY
