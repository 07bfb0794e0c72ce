#!/bin/bash
# round 2, call 35: k_p3d_hankel_imma2 (MMAs of the previous unit issued between the cell pairs of the current one)
mkdir -p gpurun_out
timeout 900 python tools/stage_ab.py dmma: imma:CPX_IMMA=1 imma2:CPX_IMMA=2 2>&1 | tee gpurun_out/c35_imma2.txt
timeout 600 python tools/stage_ab.py --n-q 56 dmma: imma:CPX_IMMA=1 imma2:CPX_IMMA=2 2>&1 | tee -a gpurun_out/c35_imma2.txt
