#!/bin/bash
# round 2, call 20: grouped explicit points; whole suite
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q ) > gpurun_out/c20_suite.log 2>&1
echo "suite rc=$?"
tail -n 25 gpurun_out/c20_suite.log
python - <<'PY'
# explicit-meshgrid timing on the S5 bench workload: the 10.5 M-point grid as a shuffled point list through get_chi2_batch
import time, numpy as np, bench
filled, data, likes, joint = bench.build_workload("S5", 100, 0)
import cupix_b200.plan as _plan
d0 = _plan.default_vector(likes[0].theory, likes[0].theory.fid_cosmo)
lo, hi = bench.grid_bounds(d0)
n = [64, 64, 16, 10]
axes = [np.linspace(lo[j], hi[j], n[j]) for j in range(4)]
pts = np.stack(np.meshgrid(*axes, indexing="ij"), -1).reshape(-1, 4)
pts = pts[np.random.default_rng(0).permutation(len(pts))]
joint.get_chi2_batch(pts[:1000], bench.NAMES)
for rep in range(2):
    t = time.time(); c = joint.get_chi2_batch(pts, bench.NAMES); dt = time.time() - t
    print("explicit shuffled meshgrid, %d points, %d tuples: %.3f s = %.3g evals/s (factored calls %d)" % (len(pts), n[2]*n[3], dt, len(pts)/dt, joint.engine().info("factored_calls")))
t = time.time(); g = joint.get_chi2_grid(bench.NAMES, lo, hi, n); dt = time.time() - t
print("same grid on-device: %.3f s = %.3g evals/s" % (dt, len(pts)/dt))
PY
