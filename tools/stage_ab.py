"""Per-stage device times of the general-points pipeline on the S5 bench workload under different settings, each in its own
process:   python tools/stage_ab.py label1:ENV=v,ENV2=v label2:... [--n-q N]
e.g.       python tools/stage_ab.py base: A:CUPIX_B200_LIB=cupix_b200/lib/variants/lib_A.so"""
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
    chi2 = torch.empty(B, dtype=torch.float64, device="cuda:0")
    stream = torch.cuda.Stream()
    kw = dict(grid=(lo, hi, bench.GRID_N), zsel=zsel, chi2_ptr=chi2.data_ptr(), stream=stream.cuda_stream, factor=False)
    for i in range(3):
        eng.eval_device(B, len(bench.NAMES), slots, grid_first=i * B, **kw)
    stage = np.zeros(4)
    for i in range(5):
        stage += np.asarray(eng.eval_device(B, len(bench.NAMES), slots, grid_first=(3 + i) * B, stage_ms=True, **kw))
    stage /= 5
    eng.eval_device(B, len(bench.NAMES), slots, grid_first=12345, **kw)
    torch.cuda.synchronize()
    print(json.dumps({"n_q": likes[0].theory.rule.n_q, "stage_ms": stage.tolist(), "chi2": chi2[:4096].cpu().numpy().tolist()}))


if __name__ == "__main__":
    args = sys.argv[1:]
    n_q = 0
    if "--n-q" in args:
        n_q = int(args[args.index("--n-q") + 1])
        del args[args.index("--n-q"):args.index("--n-q") + 2]
    if args and args[0] == "--child":
        child(int(args[1]))
        sys.exit(0)
    ref = None
    for spec in args:
        label, _, envs = spec.partition(":")
        env = dict(os.environ)
        for kv in filter(None, envs.split(",")):
            k, _, v = kv.partition("=")
            env[k] = v
        out = subprocess.run([sys.executable, __file__, "--child", str(n_q)], env=env, capture_output=True, text=True)
        if out.returncode != 0:
            print("%s failed:\n%s" % (label, out.stderr[-1500:]))
            continue
        r = json.loads(out.stdout.strip().splitlines()[-1])
        c = np.array(r["chi2"])
        if ref is None:
            ref = c
        print("%-10s n_q %3d  params %.2f  p3d_hankel %.2f  window_chi2 %.2f  reduce %.2f  total %.2f ms per 131072 points x 12 z;  max |dchi2| vs first %.2e"
              % (label, r["n_q"], *r["stage_ms"], sum(r["stage_ms"]), np.abs(c - ref).max()))
