#!/bin/bash
# round 2, call 26: k_window_tc ring depths (C: 2 operand stages + 3 worker slots + 2 scout slots)
mkdir -p gpurun_out
timeout 900 python tools/stage_ab.py base: C:CUPIX_B200_LIB=cupix_b200/lib/variants/lib_C.so 2>&1 | tee -a gpurun_out/c26_window_ab.txt
