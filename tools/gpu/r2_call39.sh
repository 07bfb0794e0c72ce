#!/bin/bash
# round 2, call 39: L2 cache-policy hints on the window kernel's TMA loads (scouts evict_last, workers evict_first)
mkdir -p gpurun_out
timeout 900 python tools/stage_ab.py base: l2hint:CUPIX_B200_LIB=cupix_b200/lib/variants/lib_l2hint.so base2: l2hint2:CUPIX_B200_LIB=cupix_b200/lib/variants/lib_l2hint.so 2>&1 | tee gpurun_out/c39_l2hint.txt
