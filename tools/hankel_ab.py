"""Per-stage device times of the general-points pipeline on the S5 bench workload for the Hankel kernels on offer
(k_p3d_hankel_imma = default, CPX_IMMA=0 = the FP64 DMMA kernel), each in its own process, plus the chi2 difference
between them on the same points.     python tools/hankel_ab.py [n_q ...]"""
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def child(n_q):
    import torch
    import bench
    filled, data, likes, joint = bench.build_workload("S5", 100, n_q)
    eng = joint.engine()
    import cupix_b200.plan as _plan
    from cupix_b200.engine import slots_from_names
    d0 = _plan.default_vector(likes[0].theory, likes[0].theory.fid_cosmo)
    lo, hi = bench.grid_bounds(d0)
    slots, zsel = slots_from_names(bench.NAMES)
    B = 131072
    dev = torch.device("cuda:0")
    chi2 = torch.empty(B, dtype=torch.float64, device=dev)
    stream = torch.cuda.Stream()
    kw = dict(grid=(lo, hi, bench.GRID_N), zsel=zsel, chi2_ptr=chi2.data_ptr(), stream=stream.cuda_stream, factor=False)
    for i in range(3):
        eng.eval_device(B, len(bench.NAMES), slots, grid_first=i * B, **kw)
    stage = np.zeros(4)
    for i in range(3):
        stage += np.asarray(eng.eval_device(B, len(bench.NAMES), slots, grid_first=(3 + i) * B, stage_ms=True, **kw))
    stage /= 3
    eng.eval_device(B, len(bench.NAMES), slots, grid_first=12345, **kw)
    torch.cuda.synchronize()
    print(json.dumps({"kernel": eng.info("hankel_kernel"), "n_q": likes[0].theory.rule.n_q, "stage_ms": stage.tolist(),
                      "chi2": chi2[:4096].cpu().numpy().tolist()}))


if __name__ == "__main__":
    if len(sys.argv) > 2 and sys.argv[1] == "--child":
        child(int(sys.argv[2]))
        sys.exit(0)
    for n_q in [int(x) for x in sys.argv[1:]] or [0]:
        res = {}
        for im in ("0", "1"):
            env = dict(os.environ, CPX_IMMA=im)
            out = subprocess.run([sys.executable, __file__, "--child", str(n_q)], env=env, capture_output=True, text=True)
            if out.returncode != 0:
                print("CPX_IMMA=%s failed:\n%s" % (im, out.stderr[-2000:]))
                continue
            res[im] = json.loads(out.stdout.strip().splitlines()[-1])
            r = res[im]
            print("n_q %3d  CPX_IMMA=%s  kernel %d  params %.2f  p3d_hankel %.2f  window_chi2 %.2f  reduce %.2f  total %.2f ms per 131072 points x 12 z"
                  % (r["n_q"], im, r["kernel"], *r["stage_ms"], sum(r["stage_ms"])))
        if len(res) == 2:
            a, b = np.array(res["0"]["chi2"]), np.array(res["1"]["chi2"])
            print("         chi2 of 4096 grid points: max |imma - dmma| = %.3e at chi2 in [%.4g, %.4g]; max relative %.2e"
                  % (np.abs(a - b).max(), a.min(), a.max(), (np.abs(a - b) / a).max()))
