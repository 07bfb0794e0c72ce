#!/bin/bash
# round 2, call 25: k_p3d_hankel_imma iterations -- A/B timing + parity tests
mkdir -p gpurun_out
timeout 600 python tools/hankel_ab.py 0 88 2>&1 | tee gpurun_out/c25_ab.txt
( time timeout 600 python -m pytest tests/test_gpu_tc.py -x -q ) > gpurun_out/c25_tests.log 2>&1
echo "tests rc=$?"; tail -n 30 gpurun_out/c25_tests.log
