#!/bin/bash
# round 2, call 31 (8 GPUs): bench under torchrun at N = 8 and N = 4
mkdir -p gpurun_out
nvidia-smi -L | wc -l
for N in 8 4; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2960$N bench.py --gpus $N --steps 8 --warmup 3 > gpurun_out/c31_bench_${N}gpu.json 2> gpurun_out/c31_bench_${N}gpu.err
  echo "bench N=$N rc=$?"
  tail -n 2 gpurun_out/c31_bench_${N}gpu.err
done
python - <<'PY'
import json
for N in (8,4):
    try:
        j=json.load(open("gpurun_out/c31_bench_%dgpu.json"%N))
        g=j["grid"]
        print("N=%d value %.4g e2e %.4g ms/step %.2f | grid %.4g e2e %.4g ms %.2f tuples/gpu %d %s" % (N, j["value"], j["e2e"]["value"], j["ms_per_step"], g["value"], g["e2e"]["value"], g["ms_per_scan"], g["tuples_per_scan_per_gpu"], g["stage_ms"]))
    except Exception as e:
        print(N, "failed", e)
PY
