#!/bin/bash
# round 2, call 30: opt-in Hankel kernels after trimming the instantiations (tests), whole suite once more
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q ) > gpurun_out/c30_suite.log 2>&1
echo "suite rc=$?"; tail -n 8 gpurun_out/c30_suite.log
