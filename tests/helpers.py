"""Shared test helpers: golden fixtures, oracle construction, and a float64 NumPy emulation of the
device algorithm (same plan arrays, same operation order at the formula level) used to validate the
host-side plan on machines without a GPU."""
import os

import numpy as np

from cupix_b200 import plan as _plan
from cupix_b200.cosmology import Cosmology
from cupix_b200.likelihood.theory import Theory
from cupix_b200.quadrature import KperpRule
from oracle.hankel import NodeHankel
from oracle.px_oracle import OracleLikelihood, OracleTheory

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
LYA_KEYS = ("bias", "beta", "q1", "q2", "kv_Mpc", "av", "bv", "kp_Mpc")
CONT_KEYS = ("b_H", "beta_H", "L_H_Mpc", "b_X", "beta_X", "b_noise_Mpc", "kC_Mpc", "pC")

_RULES = {}


def rule(n_q=96):
    if n_q not in _RULES:
        _RULES[n_q] = KperpRule(n_q)
    return _RULES[n_q]


def golden(name):
    with np.load(os.path.join(GOLDEN, name), allow_pickle=False) as f:
        return {k: f[k] for k in f.files}


def make_theory(z, model, flags, cosmo=None, n_q=96, extra=None):
    cfg = {"default_lya_model": model, "verbose": False, "n_q": n_q}
    cfg.update(flags)
    cfg.update(extra or {})
    return Theory(z, fid_cosmo=cosmo or Cosmology(), config=cfg)


def oracle_for(theory, hankel=None):
    """OracleTheory with the same defaults / flags / cosmology as a product Theory."""
    r = theory.rule
    hankel = hankel or NodeHankel(r.k, lambda rr: r.weights(rr, nthreads=1),
                                  kpar0=_plan.KPAR0_REFERENCE_MPC if getattr(theory, "kpar0", "exact") == "reference" else None)
    d = _plan.default_vector(theory, theory.fid_cosmo)
    lya = {k: d[_plan.SLOT_INDEX[k]] for k in LYA_KEYS}
    cont = {k: d[_plan.SLOT_INDEX[k]] for k in CONT_KEYS}
    c = theory.cont_model
    return OracleTheory(theory.z, theory.fid_cosmo, lya, cont, (c._xi_theta, c._xi_val), hankel,
                        include_hcd=theory.include_hcd, include_metal=theory.include_metal,
                        include_sky=theory.include_sky, include_continuum=theory.include_continuum)


def emulate_plan_px(p, params):
    """float64 NumPy evaluation of Px_Zam[a, m] from a ZbinPlan exactly as k_p3d_hankel composes it."""
    v = np.array(p.defaults, dtype=np.float64)
    for k, x in params.items():
        if k in _plan.SLOT_INDEX:
            v[_plan.SLOT_INDEX[k]] = x
    bias, beta, q1, q2, kv, av, bv, kp = v[:8]
    b_H, beta_H, L_H, b_X, beta_X, b_noise, kC, pC = v[8:16]
    kpar, kperp = p.kpar_Mpc, p.kperp_Mpc          # kpar: what the P3D sees (k_par = 0 replaced in kpar0 = 'reference' mode)
    kpar_row = getattr(p, "kpar_row_Mpc", None)
    kpar_row = kpar if kpar_row is None else kpar_row
    k2 = kpar[:, None] ** 2 + kperp[None, :] ** 2
    k = np.sqrt(k2)
    pos = k > 0
    ks = np.where(pos, k, 1.0)
    mu = np.where(pos, kpar[:, None] / ks, 0.0)
    PL = p.linP_cell
    # As / ns columns: linear power of the plan's cosmology rescaled per point (k_params, CPX_AS / CPX_NS)
    amp = 1.0
    if getattr(p, "k_pivot_Mpc", 0.0) > 0 and (v[16] != p.defaults[16] or v[17] != p.defaults[17]):
        amp = v[16] / p.defaults[16]
        PL = PL * amp * (ks / p.k_pivot_Mpc) ** (v[17] - p.defaults[17])
    D = ks ** 3 * PL / (2 * np.pi ** 2)
    with np.errstate(all="ignore"):
        vel = (ks / kv) ** av * mu ** bv
        dnl = np.exp(D * (q1 + q2 * D) * (1 - vel) - k2 / kp ** 2)
    bT, betaT = bias * np.ones_like(kpar), beta * np.ones_like(kpar)
    if p.flags & _plan.INCLUDE_HCD:
        F = np.exp(-kpar * L_H)
        bT = bias + b_H * F
        betaT = (bias * beta + b_H * beta_H * F) / bT
    cell = np.where(pos, PL * (1 + betaT[:, None] * mu ** 2) ** 2 * dnl, 0.0)        # [n_m, n_q]
    px = (p.hankel_w @ cell.T) * (bT ** 2 * p.dAA_dMpc)[None, :]                        # [n_a, n_m]
    if p.flags & _plan.INCLUDE_METAL:
        ma, mc = p.metal_auto, p.metal_cross
        dn = v[17] - p.defaults[17] if getattr(p, "k_pivot_Mpc", 0.0) > 0 else 0.0
        if dn != 0.0:                 # tilt: Taylor tables of the metal moments in dn (k_p3d_hankel epilogue, metal_moment)
            ma = sum(p.metal_auto_dn[n] * dn ** n for n in range(p.metal_dn_order + 1))
            mc = sum(p.metal_cross_dn[n] * dn ** n for n in range(p.metal_dn_order + 1))
        px = px + amp * b_X ** 2 * (ma[0] + 2 * beta_X * ma[1] + beta_X ** 2 * ma[2])
        px = px + amp * bias * b_X * (mc[0] + (beta + beta_X) * mc[1] + beta * beta_X * mc[2])
    if p.flags & _plan.INCLUDE_SKY:
        px = px + b_noise * np.outer(p.sky_xi_a, p.sky_smooth_m)
    if p.flags & _plan.INCLUDE_CONTINUUM:
        px = px * np.tanh((kpar_row / kC) ** pC)[None, :]
    return px


def emulate_plan_chi2(p, params, data_index=0):
    """float64 NumPy evaluation of (model[A, M], chi2_A[A]) the way k_window composes them
    (Cholesky whitening instead of the reference's explicit inverse)."""
    px = emulate_plan_px(p, params)
    model, chi2 = [], []
    for A in range(p.n_A):
        idx = np.where(p.B_A_a[A] != 0)[0]
        vs = p.V_aM[idx].sum(axis=0)
        mod = sum((p.V_aM[a] / vs) * (p.U_aMn[a] @ px[a]) for a in idx)
        L = np.linalg.cholesky(p.cov_AMM[A])
        y = np.linalg.solve(L, p.Px_AM[data_index, A] - mod)
        model.append(mod)
        chi2.append(y @ y)
    return np.array(model), np.array(chi2)


def oracle_likelihood(like):
    """OracleLikelihood twin of a product Likelihood."""
    return OracleLikelihood.from_data(like.data, oracle_for(like.theory), like.iz, like.N_theta_average)


class MockEmulator(object):
    """Stand-in for ``forestflow.P3D_cINN.P3DEmulator`` (absent here): a smooth, deterministic map from the six emulator
    inputs (Delta2_p, n_p, mF, gamma, sigT_Mpc, kF_Mpc) to the eight Arinyo parameters in ForestFlow's naming, written
    in torch so that the batch runs on the GPU like the real network (model_lya.py:159-171, 200-235)."""

    def predict_Arinyos_tensor(self, x):
        import torch
        d2p, n_p, mF, gamma, sigT, kF = [x[:, i] for i in range(6)]
        bias = -0.12 - 0.25 * (0.75 - mF)
        beta = 1.5 + 0.4 * (gamma - 1.5)
        q1 = 0.3 + 0.4 * (d2p - 0.35)
        kvav = 0.55 + 0.1 * (n_p + 2.3)
        av = 0.35 + 0.5 * (sigT - 0.13)
        bv = 1.6 + 0.0 * mF
        kp = 12.0 + 0.8 * (kF - 10.0)
        q2 = 0.25 + 0.1 * (mF - 0.75)
        return torch.stack([bias, beta, q1, kvav, av, bv, kp, q2], dim=1)

    def predict_Arinyos(self, emu_params):
        import torch
        from cupix_b200.likelihood.model_lya import EMU_INPUT_NAMES, FF_OUTPUT_NAMES
        x = torch.tensor([[float(emu_params[k]) for k in EMU_INPUT_NAMES]], dtype=torch.float64)
        return dict(zip(FF_OUTPUT_NAMES, [float(v) for v in self.predict_Arinyos_tensor(x)[0]]))
