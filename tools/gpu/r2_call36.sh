#!/bin/bash
# round 2, call 36: points per warp of k_p3d_hankel at the 52-node default, window theta-group count
mkdir -p gpurun_out
timeout 900 python tools/stage_ab.py ag4: ag1:CPX_WTC_AGROUPS=1 ag2:CPX_WTC_AGROUPS=2 ag4b: ag2b:CPX_WTC_AGROUPS=2 2>&1 | tee -a gpurun_out/c36_knobs.txt
