#!/bin/bash
# round 2, call 40: workers of k_window_tc walking the chunks of a tile in reverse (LRU-friendlier second read)
mkdir -p gpurun_out
timeout 600 python tools/stage_ab.py base: rev:CUPIX_B200_LIB=cupix_b200/lib/variants/lib_rev.so base2: rev2:CUPIX_B200_LIB=cupix_b200/lib/variants/lib_rev.so 2>&1 | tee gpurun_out/c40_rev.txt
( CUPIX_B200_LIB=cupix_b200/lib/variants/lib_rev.so timeout 600 python -m pytest tests/test_gpu_tc.py tests/test_gpu_oracle_configs.py -x -q ) > gpurun_out/c40_tests.log 2>&1; echo "tests (reverse build) rc=$?"; tail -n 3 gpurun_out/c40_tests.log
