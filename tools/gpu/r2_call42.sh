#!/bin/bash
# round 2, call 42: 16-byte RowPar (amp * cont folded, continuum factor in its own array only where the additive terms need it)
mkdir -p gpurun_out
timeout 600 python tools/stage_ab.py before:CUPIX_B200_LIB=cupix_b200/lib/variants/lib_before.so rowpar16: before2:CUPIX_B200_LIB=cupix_b200/lib/variants/lib_before.so rowpar16b: 2>&1 | tee gpurun_out/c42_rowpar16.txt
( time timeout 1200 python -m pytest tests -m gpu -q ) > gpurun_out/c42_suite.log 2>&1; echo "suite rc=$?"; tail -n 5 gpurun_out/c42_suite.log
