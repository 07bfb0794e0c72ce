#!/bin/bash
# round 2, call 41: k_rowpars reading exp(-k_par L_H) / tanh((k_par/kC)^pC) from per-z tables when those parameters are at their defaults
mkdir -p gpurun_out
timeout 600 python tools/stage_ab.py before:CUPIX_B200_LIB=cupix_b200/lib/variants/lib_before.so tables: before2:CUPIX_B200_LIB=cupix_b200/lib/variants/lib_before.so tables2: 2>&1 | tee gpurun_out/c41_rowtables.txt
( timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_oracle_configs.py tests/test_gpu_factor.py -x -q ) > gpurun_out/c41_tests.log 2>&1; echo "tests rc=$?"; tail -n 3 gpurun_out/c41_tests.log
