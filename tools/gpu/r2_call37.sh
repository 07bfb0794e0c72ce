#!/bin/bash
# round 2, call 37: chunk size against the wave count of k_window_tc (58 / 48 / 74 point tiles x 12 z on 148 SMs = 4.70 / 3.89 / 6.00 waves)
mkdir -p gpurun_out
timeout 900 python tools/stage_ab.py c7424: c6144:CPX_SCRATCH_MB=3330,CPX_WTC_AGROUPS=1 c9472:CPX_SCRATCH_MB=5130 c4608:CPX_SCRATCH_MB=2500,CPX_WTC_AGROUPS=1 c7424b: 2>&1 | tee gpurun_out/c37_chunk.txt
