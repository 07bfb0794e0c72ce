#!/bin/bash
# round 2, call 21: generic P3D callable path (cpx_hankel_apply), Mpc accessors, new Sampler interface; whole suite
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q ) > gpurun_out/c21_suite.log 2>&1
echo "suite rc=$?"
tail -n 30 gpurun_out/c21_suite.log
