#!/usr/bin/env python
"""Headline benchmark: Px likelihood evaluations per second on the S5 stress shape
(12 z x 20 theta x 256 k_par, dense window matrices; BASELINE.json configs[4]).

    python bench.py --gpus N --steps K --warmup W            (N>1: launched under torchrun)
    python bench.py --impl reference ...                     (CPU arm: the reference algorithm)

A step = one pass of the hot path over one batch of `--points` parameter points per GPU, taken as
consecutive slices of a 4-parameter (bias, beta, q1, kp_Mpc) scan grid; every point is scored
against the whole data vector (all 12 z bins).  See DESIGN.md section 6 for what each JSON key means.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

NAMES = ["bias", "beta", "q1", "kp_Mpc"]
GRID_N = [64, 64, 64, 40]              # 10 485 760 points ("1e7 grid points")


def grid_bounds(defaults):
    from cupix_b200.plan import SLOT_INDEX
    lo, hi = [], []
    for nm, half in zip(NAMES, (0.02, 0.2, 0.15, 4.0)):
        c = defaults[SLOT_INDEX[nm]]
        lo.append(c - half)
        hi.append(c + half)
    return lo, hi


def build_workload(shape, n_theta_average, n_q, add_noise=True, need_engine=True, precise=False, filled_path=None,
                   timings=None):
    """Synthetic S5 measurement: windows/weights seeded, data = model(truth) + noise, covariance from the model
    amplitude (px_data/synthetic.py).  The truth model comes from the GPU; CPU workers read the filled data vectors and
    covariances from ``filled_path`` (written by the B200 arm) so that both arms score the same measurement.  Without
    either (the --impl reference arm, which is timed only) the placeholder data (zeros, unit covariance) stay."""
    from cupix_b200.cosmology import Cosmology
    from cupix_b200.engine import PxEngine
    from cupix_b200.likelihood.likelihood import JointLikelihood, Likelihood
    from cupix_b200.likelihood.theory import Theory
    from cupix_b200.px_data import synthetic
    from cupix_b200.px_data.data_DESI_DR2 import DESI_DR2
    d = synthetic.make_measurement_dict(shape)
    if filled_path is not None:
        with np.load(filled_path) as f:
            d["Px_ZAM"], d["cov_ZAM"] = f["Px_ZAM"], f["cov_ZAM"]
    cosmo = Cosmology()
    device = int(os.environ.get("LOCAL_RANK", "0"))
    cfg = {"default_lya_model": "best_fit_arinyo_from_gadget", "include_hcd": True, "device": device}
    if n_q:
        cfg["n_q"] = n_q
    theories = [Theory(float(z), fid_cosmo=cosmo, config=cfg) for z in d["z_centers"]]
    data = DESI_DR2(d)
    likes = [Likelihood(data, th, iz, config={"N_theta_average": n_theta_average}) for iz, th in enumerate(theories)]
    if not need_engine:
        return d, data, likes, None
    t0 = time.time()
    plans = [lk.build_plan() for lk in likes]            # Hankel weights, linear power on the cells: the host side of set-up
    t1 = time.time()
    eng0 = PxEngine(plans, device=device, precise=precise)      # cpx_create: Cholesky, operators, digit tables, upload
    t2 = time.time()
    if timings is not None:
        timings.update(plans_s=t1 - t0, cpx_create_s=t2 - t1)
    slots = np.zeros(0, dtype=np.int32)
    model = eng0.eval(points=np.zeros((1, 0)), slots=slots, want=("model",))["model"][0]    # truth = per-z defaults
    eng0.close()
    filled = synthetic.fill_from_model(d, model, add_noise=add_noise)
    data = DESI_DR2(filled)
    likes = [Likelihood(data, th, iz, config={"N_theta_average": n_theta_average}) for iz, th in enumerate(theories)]
    joint = JointLikelihood(likes, device=device, precise_cells=precise)       # second create: now with the real covariance
    joint.engine()
    return filled, data, likes, joint


class ClockSampler(object):
    """nvidia-smi clocks / throttle reasons during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, smax, reasons, power = [], None, set(), []
        for t, line in self.rows:
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9 or not (t0 - 0.05 <= t <= t1 + 0.15):
                continue
            try:
                sm.append(float(f[1]))
                smax = float(f[2])
                power.append(float(f[3]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": smax,
                "reasons": sorted(reasons), "power_w_max": max(power) if power else None, "samples": len(sm)}


def algorithmic_work(likes, n_q):
    """Per evaluation (one point, all z): FP64 Hankel FMAs, P3D cells, window FMAs."""
    hankel = cells = window = chi2 = 0
    for lk in likes:
        n_a, n_m = len(lk.data.theta_min_a_arcmin), len(lk.data.k_m[lk.iz])
        n_M, n_A = len(lk.data.k_M_edges[lk.iz]) - 1, len(lk.data.theta_min_A_arcmin)
        cells += n_m * n_q
        hankel += n_a * n_m * n_q
        window += int(np.count_nonzero(lk.data.B_A_a)) * n_M * n_m
        chi2 += n_A * n_M
    return {"p3d_cells": cells, "hankel_dfma": hankel, "window_dfma": window, "chi2_dfma": chi2}


def window_roofline(work, likes, B, k2_s, dmma_peak_tflops, share):
    """Roofline entry of the chi2 stage.  Default: k_window_tc (window contraction as an exact INT8 split on
    tcgen05) -- bound by the one read of Px_Zam from HBM (n_a * n_m_pad float64 per point and z);
    CPX_WTC=0: k_window, the FP64 DMMA GEMM, bound by the FP64 tensor pipe."""
    if os.environ.get("CPX_WTC", "1") != "0":
        px_bytes = 0
        for lk in likes:
            n_a, n_m = len(lk.data.theta_min_a_arcmin), len(lk.data.k_m[lk.iz])
            px_bytes += 8 * n_a * (-(-n_m // 32) * 32)
        peak = None
        try:
            peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))).get("hbm_gbs"))
        except Exception:
            pass
        achieved = px_bytes * B / k2_s / 1e9
        return {"kernel": "k_window_tc", "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": (achieved / peak) if peak else None, "algorithmic_bytes_per_point_z": px_bytes / len(likes),
                "int8_mac_per_s": 15.0 * work["window_dfma"] * B / k2_s, "share_of_step": share}
    achieved = 2.0 * work["window_dfma"] * B / k2_s / 1e12
    return {"kernel": "k_window", "bound": "tensor", "achieved": achieved, "peak": dmma_peak_tflops, "unit": "TFLOP/s",
            "frac": achieved / dmma_peak_tflops, "share_of_step": share}


def cpu_worker(args):
    """One process of the CPU arm: evaluates its block of points serially (the reference pattern,
    scripts/chi2_scan_mock.py:140), BLAS pinned to one thread (scripts/profile.py:2-7)."""
    flavour, shape, n_t, n_q, pts = args[:5]
    want_truth = len(args) > 5 and args[5]
    filled_path = args[6] if len(args) > 6 else None
    try:
        from threadpoolctl import threadpool_limits
        threadpool_limits(1)
    except Exception:
        pass
    from oracle.hankel import FFTLogHankel
    from tests import helpers as H
    d, data, likes, _ = build_workload(shape, n_t, n_q, need_engine=False, filled_path=filled_path)
    orcs = []
    for lk in likes:
        o = H.oracle_likelihood(lk)
        if flavour == "fftlog":
            o.theory.hankel = FFTLogHankel()
        else:
            # fixed-node rule: per-theta weights are parameter independent -> computed once, outside the timing
            r = lk.theory.rule
            cache = {}

            def wfn(rr, r=r, cache=cache):
                key = rr.tobytes()
                if key not in cache:
                    cache[key] = r.weights(rr, nthreads=1)
                return cache[key]
            o.theory.hankel.weight_fn = wfn
            o.get_chi2({})
        orcs.append(o)
    t0 = time.time()
    out = [sum(o.get_chi2(dict(zip(NAMES, p))) for o in orcs) for p in pts]
    dt = time.time() - t0
    # outside the timing: chi2 at the per-z truth (all parameters at their defaults), the best-fit region of the data
    truth = sum(o.get_chi2({}) for o in orcs) if want_truth else None
    return dt, out, truth


def sample_grid_points(defaults, n, seed=7):
    """``n`` random points OF THE SCAN GRID (multi-indices [n, 4] in NAMES order, and their coordinates)."""
    lo, hi = grid_bounds(defaults)
    rng = np.random.default_rng(seed)
    idx = np.stack([rng.integers(0, g, n) for g in GRID_N], axis=1)
    axes = [np.linspace(lo[j], hi[j], GRID_N[j]) for j in range(len(NAMES))]
    return idx, np.stack([axes[j][idx[:, j]] for j in range(len(NAMES))], axis=1)


def run_cpu(flavour, shape, n_t, n_q, per_core, defaults, filled_path=None):
    import multiprocessing as mp
    cores = len(os.sched_getaffinity(0))
    idx, pts = sample_grid_points(defaults, cores * per_core)
    ctx = mp.get_context("fork")
    t0 = time.time()
    with ctx.Pool(cores) as pool:
        res = pool.map(cpu_worker, [(flavour, shape, n_t, n_q, pts[i::cores], i == 0, filled_path) for i in range(cores)])
    wall = time.time() - t0
    busy = max(r[0] for r in res)
    chi2 = np.empty(len(pts))
    for i in range(cores):
        chi2[i::cores] = res[i][1]
    return {"value": len(pts) / busy, "unit": "evals/s", "cores": cores, "evals": len(pts),
            "busy_s": busy, "wall_s": wall, "points": pts, "grid_index": idx, "chi2": chi2, "chi2_truth": res[0][2]}


def emit(line):
    """The one JSON line goes to the real stdout; everything else a library prints (NCCL banners...)
    was diverted to stderr by main()."""
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())


_REAL_STDOUT = 1


def main():
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)                      # stdout of this process and its children -> stderr
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--points", type=int, default=1 << 17, help="parameter points per GPU per step")
    ap.add_argument("--shape", default="S5")
    ap.add_argument("--n-theta-average", type=int, default=100)
    ap.add_argument("--n-q", type=int, default=0,
                    help="k_perp nodes; 0 = what Theory selects for px_rtol = 1e-5 and this linear power (52 log-uniform nodes)")
    ap.add_argument("--alt-n-q", type=int, default=88, help="a second, short device-timed measurement at this node count (0 = skip)")
    ap.add_argument("--cpu-evals-per-core", type=int, default=2)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-grid", action="store_true", help="skip the factorised whole-grid scan (`grid` key)")
    ap.add_argument("--no-precise", action="store_true", help="skip the short float64-cell measurement (`precise` key)")
    ap.add_argument("--precise", action="store_true", help="float64 P3D cells (CPX_PRECISE_CELLS, reference-precision mode)")
    a = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    workload = "S5 stress: 12 z x 20 theta x 256 k_par, dense windows, %d-point 4-parameter grid" % int(np.prod(GRID_N)) \
        if a.shape == "S5" else a.shape
    config = {"workload": workload, "shape": a.shape, "points_per_gpu_per_step": a.points, "params": NAMES,
              "grid_n": GRID_N, "N_theta_average": a.n_theta_average, "n_q": a.n_q, "include_hcd": True, "precise_cells": bool(a.precise),
              "parallelism": "points sharded over %d GPU(s), NCCL all_gather of chi2" % a.gpus,
              "l2": "per-chunk Px_Zam scratch (GiBs) >> 126 MB L2, no explicit flush"}

    # ------------------------------------------------------------------ CPU reference arm
    if a.impl == "reference":
        if rank != 0:
            return
        import cupix_b200.plan as _plan
        d, data, likes, _ = build_workload(a.shape, a.n_theta_average, a.n_q, need_engine=False)
        config.update(n_q=likes[0].theory.rule.n_q, kperp_family=likes[0].theory.kperp_family, px_rtol=likes[0].theory.px_rtol)
        defaults = _plan.default_vector(likes[0].theory, likes[0].theory.fid_cosmo)
        # W untimed and K timed steps like the B200 arm; a step is a bounded sample of the workload (one
        # evaluation per host core, a few seconds), and the counts are capped so that the run ends within minutes
        n_warm, n_steps = min(max(a.warmup, 0), 2), min(max(a.steps, 1), 20)
        for _ in range(n_warm):
            run_cpu("fftlog", a.shape, a.n_theta_average, a.n_q, 1, defaults)
        vals = [run_cpu("fftlog", a.shape, a.n_theta_average, a.n_q, 1, defaults) for _ in range(n_steps)]
        busy = sum(r["busy_s"] for r in vals)
        best = dict(vals[0], value=sum(r["evals"] for r in vals) / busy, busy_s=busy / n_steps)
        line = {"impl": "reference", "metric": "Px likelihood evals/sec", "value": best["value"], "unit": "evals/s",
                "n_gpus": a.gpus, "steps": n_steps, "warmup": n_warm, "ms_per_step": 1e3 * best["busy_s"],
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
                "data": "synthetic", "config": config,
                "cpu_baseline": {"value": best["value"], "unit": "evals/s", "cores": best["cores"], "kind": "port",
                                 "sample": "%d evaluations per step (1 per core) of the S5 workload with the reference's "
                                           "FFTLog Hankel algorithm (oracle A), N_theta_average=%d, one process per core"
                                           % (best["evals"], a.n_theta_average)},
                "e2e": {"value": best["value"], "unit": "evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        emit(line)
        return

    # ------------------------------------------------------------------ B200 arm
    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the B200 arm has no CPU fallback (use --impl reference)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)

    setup = {}
    filled, data, likes, joint = build_workload(a.shape, a.n_theta_average, a.n_q, precise=a.precise, timings=setup)
    eng = joint.engine()
    n_q = likes[0].theory.rule.n_q          # the node set Theory selected (or --n-q)
    config.update(n_q=n_q, kperp_family=likes[0].theory.kperp_family, px_rtol=likes[0].theory.px_rtol)
    # what a user pays once per measurement before the first chi2: host plans (cold Hankel-weight cache) + cpx_create
    setup_s = dict(setup, total_s=setup["plans_s"] + setup["cpx_create_s"])
    import cupix_b200.plan as _plan
    from cupix_b200.engine import slots_from_names, probe_peaks
    defaults = _plan.default_vector(likes[0].theory, likes[0].theory.fid_cosmo)
    lo, hi = grid_bounds(defaults)
    slots, zsel = slots_from_names(NAMES)
    total = int(np.prod(GRID_N))
    B = a.points
    stream = torch.cuda.Stream(device=dev)
    chi2_dev = torch.empty(B, dtype=torch.float64, device=dev)
    gathered = torch.empty(B * world, dtype=torch.float64, device=dev) if world > 1 else None

    def step_device(i):
        first = ((i * world + rank) * B) % max(total - B, 1)
        # `value` is the general-points pipeline: every point through P3D + Hankel + window (factor=False); the
        # factorised scan of the same grid is reported separately as `grid`
        eng.eval_device(B, len(NAMES), slots, grid=(lo, hi, GRID_N), grid_first=first, zsel=zsel,
                        chi2_ptr=chi2_dev.data_ptr(), stream=stream.cuda_stream, factor=False)
        if world > 1:
            with torch.cuda.stream(stream):
                dist.all_gather_into_tensor(gathered, chi2_dev)

    def sync_all():
        stream.synchronize()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    for i in range(a.warmup):
        step_device(i)
    sync_all()
    launches0 = eng.info("kernel_launches")
    sampler = ClockSampler(local_rank) if rank == 0 else None
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_wall0 = time.time()
    e0.record(stream)
    for i in range(a.steps):
        step_device(a.warmup + i)
    e1.record(stream)
    sync_all()
    t_wall1 = time.time()
    ms = e0.elapsed_time(e1)
    launches = eng.info("kernel_launches") - launches0
    clocks = sampler.stop(t_wall0, t_wall1) if sampler else None
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    value = a.steps * B * world / (ms * 1e-3)

    # ---- per-kernel device times (CUDA events on the launching stream, inside the library) ----
    stage = np.zeros(4)
    for i in range(a.steps):
        first = ((i * world + rank) * B) % max(total - B, 1)
        stage += eng.eval_device(B, len(NAMES), slots, grid=(lo, hi, GRID_N), grid_first=first, zsel=zsel,
                                 chi2_ptr=chi2_dev.data_ptr(), stream=stream.cuda_stream, stage_ms=True, factor=False)
    stage /= a.steps
    chunk = eng.info("chunk")
    n_chunks = (B + chunk - 1) // chunk

    # ---- end to end through the public API with pinned host buffers ----
    e2e = None
    if not a.no_e2e:
        rng = np.random.default_rng(1234 + rank)
        n_sets = a.warmup + a.steps
        # every step has its own input points, prepared in pinned host memory before the clock starts
        pts_host = torch.empty((n_sets, B, len(NAMES)), dtype=torch.float64).pin_memory()
        out_host = torch.empty((n_sets, B), dtype=torch.float64).pin_memory()
        pts_np, out_np = pts_host.numpy(), out_host.numpy()
        pts_np[:] = rng.uniform(lo, hi, size=pts_np.shape)

        gath_pin = torch.empty(B * world, dtype=torch.float64).pin_memory() if world > 1 else None

        def step_host(i):
            if world == 1:
                # the public batched call: H2D of this step's points + all kernels + D2H of chi2 inside cpx_eval
                # (host-pointer mode), result written into the caller's (pinned) array
                joint.get_chi2_batch(pts_np[i], NAMES, out=out_np[i])
            else:
                # N ranks: the step's points go up from pinned memory, cpx_eval runs on device pointers, the chi2 of all
                # ranks are gathered on the device (NCCL) and the gathered array comes down to pinned host memory
                with torch.cuda.stream(stream):
                    p_dev = pts_host[i].to(dev, non_blocking=True)
                    eng.eval_device(B, len(NAMES), slots, points_ptr=p_dev.data_ptr(), zsel=zsel, chi2_ptr=chi2_dev.data_ptr(),
                                    stream=stream.cuda_stream)
                    dist.all_gather_into_tensor(gathered, chi2_dev)
                    gath_pin.copy_(gathered, non_blocking=True)
                    p_dev.record_stream(stream)
                stream.synchronize()
        for i in range(a.warmup):
            step_host(i)
        sync_all()
        t0 = time.time()
        for i in range(a.steps):
            step_host(a.warmup + i)
        sync_all()
        dt = time.time() - t0
        if world > 1:
            t = torch.tensor([dt], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        e2e = {"value": a.steps * B * world / dt, "unit": "evals/s", "h2d_bytes_per_step": B * len(NAMES) * 8,
               "d2h_bytes_per_step": B * 8 * world,
               "api": "JointLikelihood.get_chi2_batch(points, names, out=pinned array) -> C ABI cpx_eval with host pointers" if world == 1
               else "pinned points -> device, cpx_eval with device pointers, NCCL all_gather of chi2, gathered chi2 -> pinned host"}

    # ---- the same device-timed step with other k_perp node sets (the accuracy / cost knob of the Hankel rule) ----
    # alt_n_q: the round-1 node count on the same smooth power; alt_bao: a linear power with a 5 % BAO-like wiggle, i.e. what a
    # CAMB / LaCE cosmology looks like -- Theory then selects the refined 'bao' node set by itself (128 nodes at px_rtol = 1e-5)
    def timed_alt(cfg2, cosmo2):
        from cupix_b200.likelihood.likelihood import JointLikelihood, Likelihood
        from cupix_b200.likelihood.theory import Theory
        likes2 = [Likelihood(data, Theory(lk.theory.z, fid_cosmo=cosmo2 or lk.theory.fid_cosmo, config=cfg2), lk.iz,
                             config={"N_theta_average": a.n_theta_average}) for lk in likes]
        eng2 = JointLikelihood(likes2, device=local_rank).engine()
        n_alt = max(2, min(a.steps, 5))

        def step_alt(i):
            first = ((i * world + rank) * B) % max(total - B, 1)
            eng2.eval_device(B, len(NAMES), slots, grid=(lo, hi, GRID_N), grid_first=first, zsel=zsel,
                             chi2_ptr=chi2_dev.data_ptr(), stream=stream.cuda_stream, factor=False)
        for i in range(2):
            step_alt(i)
        sync_all()
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record(stream)
        for i in range(n_alt):
            step_alt(2 + i)
        a1.record(stream)
        sync_all()
        ams = a0.elapsed_time(a1)
        if world > 1:
            t = torch.tensor([ams], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ams = float(t.item())
        th2 = likes2[0].theory
        eng2.close()
        return {"n_q": th2.rule.n_q, "kperp_family": th2.kperp_family, "value": n_alt * B * world / (ams * 1e-3), "unit": "evals/s",
                "ms_per_step": ams / n_alt, "steps": n_alt, "rule_error_vs_reference_algorithm": th2.rule_error}

    alt = alt_bao = None
    base_cfg = {"default_lya_model": "best_fit_arinyo_from_gadget", "include_hcd": True, "device": local_rank}
    if a.alt_n_q and a.alt_n_q != n_q and not a.precise:
        alt = timed_alt(dict(base_cfg, n_q=a.alt_n_q), None)
        alt["note"] = "same step, device-timed, with the round-1 node count"
        from cupix_b200.cosmology import Cosmology
        alt_bao = timed_alt(dict(base_cfg), Cosmology({"bao_amp": 0.05}))
        alt_bao["note"] = ("same step on a linear power with a 5 % BAO-like wiggle (what CAMB delivers): the node family and count "
                           "were selected by Theory from that power (kperp_rule = 'auto', px_rtol = 1e-5)")

    # ---- the reference-precision mode of the general-points pipeline (float64 P3D cells, CPX_PRECISE_CELLS) ----
    precise_rec = None
    if not a.precise and not a.no_precise:
        Bp = min(B, 1 << 14)
        eng.precise = True
        try:
            for i in range(2):
                eng.eval_device(Bp, len(NAMES), slots, grid=(lo, hi, GRID_N), grid_first=i * Bp, zsel=zsel,
                                chi2_ptr=chi2_dev.data_ptr(), stream=stream.cuda_stream, factor=False)
            sync_all()
            p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            n_p = 3
            p0.record(stream)
            for i in range(n_p):
                eng.eval_device(Bp, len(NAMES), slots, grid=(lo, hi, GRID_N), grid_first=(2 + i) * Bp, zsel=zsel,
                                chi2_ptr=chi2_dev.data_ptr(), stream=stream.cuda_stream, factor=False)
            p1.record(stream)
            sync_all()
            pms = p0.elapsed_time(p1)
        finally:
            eng.precise = False
        if world > 1:
            t = torch.tensor([pms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            pms = float(t.item())
        precise_rec = {"value": n_p * Bp * world / (pms * 1e-3), "unit": "evals/s", "points_per_gpu_per_step": Bp, "steps": n_p,
                       "ms_per_step": pms / n_p,
                       "what": "general-points pipeline with float64 P3D cells (Theory config precise_cells / CPX_PRECISE_CELLS): "
                               "chi2 within 1e-3 absolute of the oracle at any chi2"}

    # ---- the whole 4-parameter grid as ONE scan through the factorised path (cpx_factor.cuh) ----
    # Same 10 485 760 points, flattened with the tuple axes (q1, kp_Mpc) slowest so that a rank's contiguous slice holds
    # whole tuples; each rank scans total / N points (strong scaling: the job is the grid), chi2 all-gathered.
    grid = None
    grid_host = None
    if not a.no_grid and not a.precise:
        from cupix_b200.parallel import block_partition
        perm = [2, 3, 0, 1]
        gnames = [NAMES[j] for j in perm]
        glo, ghi, gn = [lo[j] for j in perm], [hi[j] for j in perm], [GRID_N[j] for j in perm]
        gslots, gzsel = slots_from_names(gnames)
        counts, starts = block_partition(total, world)
        cnt, first = counts[rank], starts[rank]
        g_dev = torch.empty(max(counts), dtype=torch.float64, device=dev)
        g_all = torch.empty(max(counts) * world, dtype=torch.float64, device=dev) if world > 1 else None

        def scan_device():
            eng.eval_device(cnt, len(gnames), gslots, grid=(glo, ghi, gn), grid_first=first, zsel=gzsel,
                            chi2_ptr=g_dev.data_ptr(), stream=stream.cuda_stream)
            if world > 1:
                with torch.cuda.stream(stream):
                    dist.all_gather_into_tensor(g_all, g_dev)
        n_scan = max(2, min(a.steps, 5))
        f0 = eng.info("factored_calls")
        t0_ = eng.info("factored_tuples")
        for _ in range(2):
            scan_device()
        sync_all()
        tuples_per_scan = (eng.info("factored_tuples") - t0_) // 2
        assert eng.info("factored_calls") - f0 == 2, "the grid scan did not take the factorised path"
        l0 = eng.info("kernel_launches")
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record(stream)
        for _ in range(n_scan):
            scan_device()
        g1.record(stream)
        sync_all()
        gms = g0.elapsed_time(g1)
        g_launches = eng.info("kernel_launches") - l0
        gstage = eng.eval_device(cnt, len(gnames), gslots, grid=(glo, ghi, gn), grid_first=first, zsel=gzsel,
                                 chi2_ptr=g_dev.data_ptr(), stream=stream.cuda_stream, stage_ms=True)
        # end to end: host-pointer C-ABI call, chi2 of the rank's slice copied into pinned host memory
        g_pin = torch.empty(cnt, dtype=torch.float64).pin_memory()
        g_np = g_pin.numpy()
        for _ in range(2):
            eng.eval(slots=gslots, zsel=gzsel, grid=(glo, ghi, gn), grid_first=first, B=cnt, chi2_out=g_np)
        sync_all()
        tg = time.time()
        for _ in range(n_scan):
            eng.eval(slots=gslots, zsel=gzsel, grid=(glo, ghi, gn), grid_first=first, B=cnt, chi2_out=g_np)
        sync_all()
        gdt = time.time() - tg
        if world > 1:
            t = torch.tensor([gms, gdt], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            gms, gdt = float(t[0].item()), float(t[1].item())
        grid_host = (g_np, first, cnt, gn, perm)
        grid = {"value": n_scan * total / (gms * 1e-3), "unit": "evals/s", "scaling": "strong",
                "what": "the whole %d-point grid scanned through the factorised path: P3D (float64 cells) + 3-moment Hankel + "
                        "window once per distinct (q1, kp_Mpc) tuple and z, chi2 per point as a quadratic form" % total,
                "points_per_scan": total, "scans": n_scan, "ms_per_scan": gms / n_scan, "axis_order": gnames, "grid_n": gn,
                "tuples_per_scan_per_gpu": int(tuples_per_scan), "gpu_launches_per_scan": int(g_launches // n_scan),
                "stage_ms": {"tuple_params": float(gstage[0]), "basis_f64_hankel": float(gstage[1]),
                             "window_gram": float(gstage[2]), "quadratic_form": float(gstage[3])},
                "e2e": {"value": n_scan * total / gdt, "unit": "evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": cnt * 8,
                        "api": "PxEngine.eval(grid=...) = C ABI cpx_eval in grid mode with a pinned host output buffer"},
                "algorithmic": None}
        w_ = algorithmic_work(likes, n_q)         # per evaluation = summed over z
        nb_ = 9                                   # basis tables with HCD: F^i H_j, i, j = 0..2 (cpx_factor.cuh)
        grid["algorithmic"] = {"per": "tuple (all z) / point", "tuples": int(tuples_per_scan), "basis_tables": nb_,
                               "p3d_cells_f64_per_tuple": w_["p3d_cells"], "hankel_dfma_per_tuple": 3 * w_["hankel_dfma"],
                               "window_dfma_per_tuple": nb_ * w_["window_dfma"],
                               "gram_dfma_per_tuple": (nb_ * nb_ + nb_) * w_["chi2_dfma"],
                               "quadratic_form_dfma_per_point": len(likes) * (nb_ * nb_ + nb_),
                               "bytes_out_per_point": 8}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel ----
    work = algorithmic_work(likes, n_q)
    peaks = probe_peaks(local_rank)
    k1_s = stage[1] * 1e-3
    k2_s = stage[2] * 1e-3
    # dominant kernel: k_p3d_hankel, bound by the FP64 tensor pipe (DMMA m8n8k4).  `achieved` counts the
    # ALGORITHMIC flops (2 n_a n_m n_q per point and z); the kernel issues 24/20 of that because the
    # 20 theta bins occupy three 8-wide MMA column tiles.
    useful_flops = 2.0 * work["hankel_dfma"] * B
    achieved = useful_flops / k1_s / 1e12
    peak = 2.0 * peaks["dmma_fma_per_s"] / 1e12
    n_a = len(likes[0].data.theta_min_a_arcmin)
    pad = (8.0 * ((n_a + 7) // 8) if n_a <= 24 else float(n_a)) / n_a
    traffic = None
    tfile = os.path.join(ROOT, "profiles", "k_p3d_hankel_traffic.json")
    if os.path.exists(tfile):
        try:
            traffic = json.load(open(tfile)).get("dram_bytes_per_launch")
        except Exception:
            traffic = None
    roofline = {"kernel": "k_p3d_hankel", "bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                "frac": achieved / peak, "traffic": traffic,
                "peak_source": "FP64 tensor pipe (DMMA m8n8k4) rate of this GPU measured in this process by "
                               "cpx_probe_peaks; MEASURED_PEAKS.json holds only HBM and bf16 figures, which do not "
                               "bound this FP64 path",
                "pipe_frac_incl_padding": achieved * pad / peak,
                "launches_per_step": int(n_chunks), "ms_per_launch": float(stage[1] / n_chunks),
                "share_of_step": float(stage[1] / stage.sum()),
                "algorithmic": dict(work, per="evaluation (one point, all z)"),
                "fp32_fma_pipe_frac": float(31.0 * work["p3d_cells"] * B / k1_s / peaks["ffma_per_s"]),
                "mufu_pipe_frac": float(2.0 * work["p3d_cells"] * B / k1_s / peaks["mufu_per_s"]),
                "window_kernel": window_roofline(work, likes, B, k2_s, peak, float(stage[2] / stage.sum())),
                "stage_ms": {"params": float(stage[0]), "p3d_hankel": float(stage[1]), "window_chi2": float(stage[2]),
                             "reduce": float(stage[3])},
                "measured_peaks": peaks}
    hbm = None
    try:
        hbm = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        roofline["hbm_gbs_measured"] = hbm.get("hbm_gbs")
    except Exception:
        pass

    cpu = None
    parity = None
    if not a.no_cpu_baseline and a.gpus == 1:
        import tempfile
        with tempfile.TemporaryDirectory() as td:
            fpath = os.path.join(td, "filled.npz")
            np.savez(fpath, Px_ZAM=filled["Px_ZAM"], cov_ZAM=filled["cov_ZAM"])
            r = run_cpu("node", a.shape, a.n_theta_average, n_q, a.cpu_evals_per_core, defaults, filled_path=fpath)
        # the CPU leg's sample points scored by the GPU arm as well, inside a full production batch (same kernels,
        # same host-pointer C-ABI call as the e2e leg): oracle chi2 against device chi2 on the workload the number is quoted on
        pb = np.random.default_rng(99).uniform(lo, hi, size=(B, len(NAMES)))
        pb[:len(r["points"])] = r["points"]
        got = eng.eval_host_into(np.ascontiguousarray(pb), slots, zsel, np.empty(B))[:len(r["points"])]
        dchi = np.abs(got - r["chi2"])
        n_data = joint.get_ndata()
        near = r["chi2"] <= 2.0 * n_data
        parity = {"n": int(len(dchi)), "against": "float64 NumPy oracle (same node rule), the cpu_baseline sample points",
                  "max_abs_dchi2": float(dchi.max()), "max_rel_dchi2": float((dchi / np.abs(r["chi2"])).max()),
                  "frac_within_1e-3": float(np.mean(dchi <= 1e-3)),
                  "n_within_2ndata": int(near.sum()),
                  "max_abs_dchi2_within_2ndata": float(dchi[near].max()) if near.any() else None,
                  "chi2_min": float(r["chi2"].min()), "chi2_max": float(r["chi2"].max()), "n_data": int(n_data),
                  "rule": "|dchi2| <= 1e-3 where chi2 <= 2 n_data, 1e-6 relative beyond (DESIGN.md section 5)"}
        if grid_host is not None:
            # the same sample points read out of the factorised whole-grid scan (its flattening puts q1, kp_Mpc first)
            g_np, g_first, g_cnt, gn, perm = grid_host
            flat = np.ravel_multi_index(tuple(r["grid_index"][:, j] for j in perm), gn)
            dg = np.abs(g_np[flat - g_first] - r["chi2"])
            grid["parity"] = {"n": int(len(dg)), "against": parity["against"], "max_abs_dchi2": float(dg.max()),
                              "max_rel_dchi2": float((dg / np.abs(r["chi2"])).max()), "frac_within_1e-3": float(np.mean(dg <= 1e-3))}
        if precise_rec is not None:
            eng.precise = True
            try:
                gotp = eng.eval_host_into(np.ascontiguousarray(r["points"]), slots, zsel, np.empty(len(r["points"])), factor=False)
            finally:
                eng.precise = False
            dp = np.abs(gotp - r["chi2"])
            precise_rec["parity"] = {"n": int(len(dp)), "against": parity["against"], "max_abs_dchi2": float(dp.max()),
                                     "max_rel_dchi2": float((dp / np.abs(r["chi2"])).max()), "frac_within_1e-3": float(np.mean(dp <= 1e-3))}
        if r.get("chi2_truth") is not None:
            # the grid applies one (bias, beta, q1, kp) to all 12 z, so all of it is far from the best fit of this
            # data; the per-z truth (no columns: every z at its own defaults) is the point inside chi2 <= 2 n_data
            t_gpu = float(joint.get_chi2({}))
            parity["truth"] = {"chi2_oracle": float(r["chi2_truth"]), "chi2_gpu": t_gpu,
                               "abs_dchi2": abs(t_gpu - float(r["chi2_truth"]))}
        cpu = {"value": r["value"], "unit": "evals/s", "cores": r["cores"], "kind": "port",
               "sample": "%d evaluations (%d per core) of the same S5 workload with the float64 NumPy oracle using the "
                         "same fixed-node Hankel rule as the GPU (oracle B); the reference's own FFTLog algorithm is "
                         "timed by --impl reference" % (r["evals"], a.cpu_evals_per_core)}

    line = {"metric": "Px likelihood evals/sec", "value": value, "unit": "evals/s", "n_gpus": a.gpus, "steps": a.steps,
            "warmup": a.warmup, "ms_per_step": ms / a.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64" if a.precise else "f64 accumulate / f32 P3D cell", "data": "synthetic", "config": config,
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu, "parity": parity, "grid": grid, "alt_n_q": alt, "alt_bao": alt_bao, "precise": precise_rec, "setup_s": setup_s,
            "per_point_z_rate": value * len(likes)}
    emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
