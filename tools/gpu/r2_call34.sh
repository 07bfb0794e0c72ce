#!/bin/bash
# round 2, call 34: k_window_tc with the scouts' 128-bit loads kept whole (no 4-way bank conflicts) against the shipped build
mkdir -p gpurun_out
timeout 900 python tools/stage_ab.py base: lds128:CUPIX_B200_LIB=cupix_b200/lib/variants/lib_lds128.so base2: lds128b:CUPIX_B200_LIB=cupix_b200/lib/variants/lib_lds128.so 2>&1 | tee gpurun_out/c34_window_lds128.txt
