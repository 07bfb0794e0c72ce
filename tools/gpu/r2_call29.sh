#!/bin/bash
# round 2, call 29 (2 GPUs): the two-rank NCCL scan test and the bench under torchrun at N = 2 on the final code
mkdir -p gpurun_out
nvidia-smi -L | wc -l
( time timeout 600 python -m pytest tests/test_gpu_scripts.py -q ) > gpurun_out/c29_scripts.log 2>&1; echo "scripts tests rc=$?"; tail -n 4 gpurun_out/c29_scripts.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29602 bench.py --gpus 2 --steps 8 --warmup 3 > gpurun_out/c29_bench_2gpu.json 2> gpurun_out/c29_bench_2gpu.err
echo "bench N=2 rc=$?"; tail -n 2 gpurun_out/c29_bench_2gpu.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29603 bench.py --impl reference --gpus 2 --steps 1 --warmup 0 > gpurun_out/c29_ref_2gpu.json 2> gpurun_out/c29_ref_2gpu.err
echo "reference arm N=2 rc=$?"; tail -c 600 gpurun_out/c29_ref_2gpu.json
python - <<'PY'
import json
j=json.loads(open("gpurun_out/c29_bench_2gpu.json").read().strip().splitlines()[-1])
g=j["grid"]
print("N=2 value %.4g e2e %.4g ms/step %.2f | grid %.4g e2e %.4g" % (j["value"], j["e2e"]["value"], j["ms_per_step"], g["value"], g["e2e"]["value"]))
PY
