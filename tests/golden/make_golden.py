"""Generate the golden fixtures in this directory by running the UNMODIFIED reference modules.

Run in the build container only (needs /root/reference):   python tests/golden/make_golden.py

What runs from /root/reference, untouched:
    cupix.likelihood.theory.Theory            (orchestration, P3D models, contaminants, units)
    cupix.likelihood.model_lya.LyaModel       (defaults from the CoLoRe FITS headers / Gadget CSV)
    cupix.likelihood.model_contaminants.ContaminantsModel
    cupix.likelihood.likelihood.Likelihood    (theta average, window, rebin, chi2)
    cupix.likelihood.window_and_rebin
    cupix.px_data.base_px_data.BaseDataPx     (container + range cuts)

What is absent from the container and therefore replaced by stand-ins injected into sys.modules:
    lace.cosmo.cosmology / rescale_cosmology  -> cupix_b200.cosmology (analytic LCDM + EH98)
    forestflow.pcross.Px_Mpc_detailed         -> the fixed-node Hankel rule (oracle.hankel.NodeHankel
                                                  with cupix_b200.quadrature.KperpRule weights)
    forestflow.model_p3d_arinyo.coordinates   -> identity decorator
    forestflow.priors / P3D_cINN              -> empty shells (not exercised)
    astropy.io.fits / astropy.table, h5py, matplotlib -> minimal readers / empty shells

So the fixtures pin everything the reference itself computes on this path; the third-party
Hankel transform is the one piece whose parity stays unpinned (see oracle/px_oracle.py header).
"""
import os
import re
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, REPO)
sys.path.insert(0, REF)

from cupix_b200 import cosmology as b200_cosmology          # noqa: E402
from cupix_b200.quadrature import KperpRule                   # noqa: E402
from cupix_b200.px_data import synthetic                      # noqa: E402
from oracle.hankel import NodeHankel                          # noqa: E402

N_Q = 96
RULE = KperpRule(N_Q)
HANKEL = NodeHankel(RULE.k, lambda r: RULE.weights(r, nthreads=1))


def _module(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def install_stand_ins():
    # --- lace ---
    lace = _module("lace", __path__=[os.path.join(REPO, "_no_lace")])
    cosmo_pkg = _module("lace.cosmo", __path__=[])
    cmod = _module("lace.cosmo.cosmology", Cosmology=b200_cosmology.Cosmology)
    rmod = _module("lace.cosmo.rescale_cosmology",
                   RescaledCosmology=lambda fid_cosmo, new_params_dict: b200_cosmology.RescaledCosmology(fid_cosmo, new_params_dict))
    tmod = _module("lace.cosmo.thermal_broadening", thermal_broadening_kms=lambda T0: 9.1 * np.sqrt(T0 / 1e4))
    lace.cosmo = cosmo_pkg
    cosmo_pkg.cosmology, cosmo_pkg.rescale_cosmology, cosmo_pkg.thermal_broadening = cmod, rmod, tmod

    # --- forestflow ---
    def coordinates(_kind):
        return lambda fn: fn

    def Px_Mpc_detailed(z, kpar_iMpc, rperp_Mpc, p3d_fun_Mpc, p3d_params=None, max_k_for_p3d=200.0, **kw):
        return HANKEL(rperp_Mpc, kpar_iMpc, lambda k, mu: p3d_fun_Mpc(z, k, mu, ari_pp=p3d_params))

    ff = _module("forestflow", __path__=[os.path.join(REPO, "_no_forestflow", "forestflow")])
    ff.pcross = _module("forestflow.pcross", Px_Mpc_detailed=Px_Mpc_detailed)
    ff.model_p3d_arinyo = _module("forestflow.model_p3d_arinyo", coordinates=coordinates)
    ff.priors = _module("forestflow.priors")
    ff.P3D_cINN = _module("forestflow.P3D_cINN", P3DEmulator=object)

    # --- astropy: the two readers the reference uses ---
    class _HDU(object):
        def __init__(self, header):
            self.header = header

    class _FitsFile(list):
        def __enter__(self):
            return self

        def __exit__(self, *a):
            return False

    def fits_open(path):
        raw = open(path, "rb").read()
        hdr = {}
        for i in range(0, min(len(raw), 2880 * 64), 80):
            card = raw[i:i + 80].decode("ascii", "replace")
            m = re.match(r"^(HIERARCH\s+)?([A-Za-z0-9_ ]+?)\s*=\s*([^/]+)", card)
            if m:
                key, val = m.group(2).strip(), m.group(3).strip()
                try:
                    hdr.setdefault(key, float(val))
                except ValueError:
                    pass
        return _FitsFile([_HDU({}), _HDU(hdr)])

    class Table(dict):
        @classmethod
        def read(cls, fname):
            tab = np.loadtxt(fname, delimiter=",", skiprows=1)
            cols = open(fname).readline().strip().split(",")
            return cls({c: tab[:, i] for i, c in enumerate(cols)})

    ap = _module("astropy", __path__=[])
    ap.io = _module("astropy.io", __path__=[])
    ap.io.fits = _module("astropy.io.fits", open=fits_open)
    ap.table = _module("astropy.table", Table=Table)

    # --- things the reference imports at module level but does not use on this path ---
    _module("h5py")
    mpl = _module("matplotlib", __path__=[])
    mpl.pyplot = _module("matplotlib.pyplot")
    mpl.colors = _module("matplotlib.colors", ListedColormap=object)


def main():
    install_stand_ins()
    from cupix.likelihood.theory import Theory as RefTheory
    from cupix.likelihood.likelihood import Likelihood as RefLikelihood
    from cupix.px_data.base_px_data import BaseDataPx as RefBaseDataPx

    rng = np.random.default_rng(synthetic.SEED)
    cosmo = b200_cosmology.Cosmology()

    # ---------------- fixture 1: Theory.get_px_obs for every contaminant combination -----------
    theta = np.array([0.6, 2.5, 9.0, 31.0, 57.0])
    k_AA = np.array([0.0, 0.00767, 0.0614, 0.3, 0.9, 1.19])
    flag_sets = {
        "lya": {},
        "hcd": {"include_hcd": True},
        "metal": {"include_metal": True},
        "sky": {"include_sky": True},
        "continuum": {"include_continuum": True},
        "all": {"include_hcd": True, "include_metal": True, "include_sky": True, "include_continuum": True},
    }
    param_sets = {
        "defaults": {},
        "bias_beta": {"bias": -0.13, "beta": 1.3},
        "arinyo": {"bias": -0.15, "beta": 1.7, "q1": 0.6, "q2": 0.35, "kv_Mpc": 0.8, "av": 0.6, "bv": 1.7, "kp_Mpc": 15.0},
        "linear": {"q1": 0.0, "q2": 0.0, "kp_Mpc": 150.0},
        "bv0": {"bv": 0.0},
        "contaminants": {"b_H": -0.03, "beta_H": 0.7, "L_H_Mpc": 5.0, "b_X": -0.008, "beta_X": 0.4,
                         "b_noise_Mpc": 0.002, "kC_Mpc": 0.03, "pC": 0.8, "not_a_parameter": 3.0},
    }
    out = {"theta_arc": theta, "k_AA": k_AA, "n_q": N_Q}
    for model, z in (("best_fit_arinyo_from_colore", 2.2), ("best_fit_arinyo_from_gadget", 2.4),
                     ("best_fit_arinyo_from_colore_pressure_only", 2.8)):
        for fname, flags in flag_sets.items():
            cfg = {"default_lya_model": model, "verbose": False}
            cfg.update(flags)
            th = RefTheory(z, fid_cosmo=cosmo, config=cfg)
            for pname, params in param_sets.items():
                px = th.get_px_obs(theta_arc=theta, k_AA=k_AA, params=params)
                out[f"px|{model}|{z}|{fname}|{pname}"] = px
            out[f"lyadef|{model}|{z}"] = np.array([th.lya_model.default_lya_params[k] for k in
                                                   ("bias", "beta", "q1", "q2", "kv_Mpc", "av", "bv", "kp_Mpc")])
    np.savez_compressed(os.path.join(HERE, "ref_theory_px.npz"), **out)
    with open(os.path.join(HERE, "ref_theory_param_sets.py"), "w") as f:
        f.write("# written by make_golden.py -- the parameter/flag sets behind ref_theory_px.npz\n")
        f.write("FLAG_SETS = %r\nPARAM_SETS = %r\n" % (flag_sets, param_sets))

    # ---------------- fixture 2: Likelihood on a synthetic measurement ---------------------------
    for shape, N_t, cfg_extra in (("tiny", 1, {}), ("tiny_rebin", 5, {"include_hcd": True, "include_metal": True,
                                                                     "include_sky": True, "include_continuum": True})):
        d = synthetic.make_measurement_dict(shape, z0=2.2, dz=0.2)
        Nz = len(d["z_centers"])
        truth = {"bias": -0.115, "beta": 1.5}
        trial = {"bias": -0.12, "beta": 1.62, "q1": 0.5}
        res = {k: v for k, v in d.items()}
        res["N_theta_average"] = N_t
        likes = []
        for iz in range(Nz):
            cfg = {"default_lya_model": "best_fit_arinyo_from_colore", "verbose": False}
            cfg.update(cfg_extra)
            th = RefTheory(float(d["z_centers"][iz]), fid_cosmo=cosmo, config=cfg)
            data = RefBaseDataPx(d["Px_ZAM"], d["cov_ZAM"], d["z_centers"], d["k_M_edges"], d["theta_min_A"],
                                 d["theta_max_A"], d["N_fft"], d["L_fft"], k_m_AA=d["k_m"],
                                 theta_min_a_arcmin=d["theta_min_a"], theta_max_a_arcmin=d["theta_max_a"],
                                 B_A_a=d["B_A_a"], U_ZaMn=d["U_ZaMn"], V_ZaM=d["V_ZaM"])
            likes.append(RefLikelihood(data, th, iz, config={"N_theta_average": N_t}))
        # forecast recipe: data = convolved truth + noise (likelihood.py:37-50 adds L n)
        model_truth = np.array([lk.get_convolved_px(params=truth) for lk in likes])
        filled = synthetic.fill_from_model(d, model_truth, rng=rng, add_noise=True)
        res["Px_ZAM"], res["cov_ZAM"] = filled["Px_ZAM"], filled["cov_ZAM"]
        for iz, lk in enumerate(likes):
            lk.data.Px_ZAM, lk.data.cov_ZAM = filled["Px_ZAM"], filled["cov_ZAM"]
            res[f"Px_Zam_trial_{iz}"] = lk._compute_Px_Zam(params=trial)
            res[f"model_trial_{iz}"] = lk.get_convolved_px(params=trial)
            chi2, info = lk.get_chi2(params=trial, return_info=True)
            res[f"chi2_trial_{iz}"] = chi2
            res[f"log_like_all_trial_{iz}"] = info["log_like_all"]
            res[f"chi2_truth_{iz}"] = lk.get_chi2(params=truth)
            res[f"model_truth_{iz}"] = model_truth[iz]
        res["truth_keys"], res["truth_vals"] = np.array(list(truth)), np.array(list(truth.values()))
        res["trial_keys"], res["trial_vals"] = np.array(list(trial)), np.array(list(trial.values()))
        res["flags"] = np.array(sorted(cfg_extra))
        np.savez_compressed(os.path.join(HERE, f"ref_likelihood_{shape}.npz"), **res)

    # ---------------- fixture 3: BaseDataPx range cuts ------------------------------------------
    d = synthetic.make_measurement_dict("tiny_rebin")
    data = RefBaseDataPx(d["Px_ZAM"], d["cov_ZAM"], d["z_centers"], d["k_M_edges"], d["theta_min_A"],
                         d["theta_max_A"], d["N_fft"], d["L_fft"], k_m_AA=d["k_m"],
                         theta_min_a_arcmin=d["theta_min_a"], theta_max_a_arcmin=d["theta_max_a"],
                         B_A_a=d["B_A_a"], U_ZaMn=d["U_ZaMn"], V_ZaM=d["V_ZaM"],
                         theta_min_cut_arcmin=1.0, kM_max_cut_AA=0.2, km_max_cut_AA=0.22)
    np.savez_compressed(os.path.join(HERE, "ref_data_cuts.npz"), k_m=np.asarray(data.k_m),
                        k_M_edges=np.asarray(data.k_M_edges), U_shape=np.array(data.U_ZaMn.shape),
                        U_sum=data.U_ZaMn.sum(), V_sum=data.V_ZaM.sum(), B_A_a=data.B_A_a,
                        theta_min_A=data.theta_min_A_arcmin, theta_min_a=data.theta_min_a_arcmin,
                        Px_shape=np.array(data.Px_ZAM.shape), cov_shape=np.array(data.cov_ZAM.shape))
    print("golden fixtures written to", HERE)
    for f in sorted(os.listdir(HERE)):
        print("  ", f, os.path.getsize(os.path.join(HERE, f)))


if __name__ == "__main__":
    main()
