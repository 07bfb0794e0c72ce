"""Golden vectors for the P3D -> Px Hankel step, produced by RUNNING the reference's own in-tree
implementation of that step:  cupix/likelihood/lyaP3D.py, functions ``Px_Mpc_withSiIII`` (:127-362)
and ``P1D_Mpc`` (:66-106) -- the code its header calls "copied from ForestFlow", i.e. the algorithm of
``forestflow.pcross.Px_Mpc_detailed`` (FFTLog on 2^11 log-spaced k_perp in [1e-7, 1e3], P1D splice
below r ~ 0.02-0.2 Mpc, cubic spline to the requested separations).

Run in the build container only (needs /root/reference):   python tests/golden/make_golden_hankel.py

The module is imported untouched.  Its two absent imports are satisfied by stand-ins:
    forestflow.pcross   empty module (imported at lyaP3D.py:2, not used by the functions run here)
    hankl.FFTLog        scipy.fft.fht, SciPy's implementation of the same published algorithm
                        (Hamilton 2000): hankl.FFTLog(x, f, q=0, mu=0) returns y = 1 / x[::-1] and
                        G(y) = int F(x) J_0(xy) y dx, which is scipy.fft.fht(f, dln, mu=0, offset=0, bias=0).
The SiIII terms of ``Px_Mpc_withSiIII`` are switched off with bias_SiIII = 0, which leaves exactly
the Lya transform of the P3D callable that is passed in.

The P3D callable is the Arinyo model of cupix/likelihood/theory.py:285-308 evaluated with the
linear power of cupix_b200.cosmology (the stand-in for LaCE/CAMB used by every other fixture).
"""
import os
import sys
import types

import numpy as np
import scipy.fft

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, REPO)

from cupix_b200.cosmology import Cosmology   # noqa: E402

PARAM_SETS = {
    "gadget_z2.4": dict(z=2.4, bias=-0.1483, beta=1.49, q1=0.297, q2=0.302, kv=0.527 ** (1 / 0.347), av=0.347, bv=1.68, kp=13.4),
    "colore_like": dict(z=2.2, bias=-0.115, beta=1.55, q1=0.1, q2=0.0, kv=1.0, av=0.3, bv=1e-13, kp=0.6),
    "strong_nl": dict(z=3.0, bias=-0.23, beta=1.2, q1=0.9, q2=0.4, kv=0.8, av=0.6, bv=1.9, kp=20.0),
    "linear": dict(z=2.8, bias=-0.2, beta=1.0, q1=0.0, q2=0.0, kv=1.0, av=0.5, bv=1.5, kp=1e4),
    # the same Arinyo parameters on a linear power with a 5 % BAO-like wiggle (Cosmology(bao_amp=0.05)): what the
    # refined 'bao' node sets exist for
    "gadget_z2.4_bao": dict(z=2.4, bias=-0.1483, beta=1.49, q1=0.297, q2=0.302, kv=0.527 ** (1 / 0.347), av=0.347, bv=1.68, kp=13.4,
                            bao_amp=0.05),
    "strong_nl_bao": dict(z=3.0, bias=-0.23, beta=1.2, q1=0.9, q2=0.4, kv=0.8, av=0.6, bv=1.9, kp=20.0, bao_amp=0.05),
}
RPERP = np.concatenate([[0.012, 0.05, 0.15, 0.3], np.geomspace(0.5, 80.0, 14)])
# 1e-4 is what the reference's model_Px substitutes for k_par = 0 (lyaP3D.py:35-39) before calling Px_Mpc_withSiIII,
# which itself refuses a zero (:208-209): the first column IS the reference's k_par = 0 row
KPAR = np.array([1e-4, 1e-3, 0.01, 0.05, 0.2, 0.5, 1.0, 2.0, 4.0])


def load_reference_lyaP3D():
    def FFTLog(x, f_x, q=0, mu=0, **kw):
        assert q == 0
        dln = np.log(x[-1] / x[0]) / (len(x) - 1)
        return 1.0 / x[::-1], scipy.fft.fht(f_x, dln, mu=mu, offset=0.0, bias=0.0)

    ff = types.ModuleType("forestflow")
    ff.__path__ = []
    ff.pcross = types.ModuleType("forestflow.pcross")
    sys.modules.update({"forestflow": ff, "forestflow.pcross": ff.pcross, "hankl": types.SimpleNamespace(FFTLog=FFTLog)})
    mpl = types.ModuleType("matplotlib")
    mpl.__path__ = []
    mpl.pyplot = types.ModuleType("matplotlib.pyplot")
    sys.modules.setdefault("matplotlib", mpl)
    sys.modules.setdefault("matplotlib.pyplot", mpl.pyplot)
    import importlib.util
    spec = importlib.util.spec_from_file_location("ref_lyaP3D", os.path.join(REF, "cupix", "likelihood", "lyaP3D.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def arinyo_p3d(cosmo):
    """P3D(z, k, mu, params) of theory.py:285-308 (k in 1/Mpc)."""
    def p3d(z, k, mu, pp):
        linP = cosmo.get_linP_Mpc(z, k)
        delta2 = k ** 3 * linP / (2 * np.pi ** 2)
        nonlin = delta2 * (pp["q1"] + pp["q2"] * delta2)
        vel = (k / pp["kv"]) ** pp["av"] * np.abs(mu) ** pp["bv"]
        dnl = np.exp(nonlin * (1 - vel) - (k / pp["kp"]) ** 2)
        return pp["bias"] ** 2 * (1 + pp["beta"] * mu ** 2) ** 2 * linP * dnl
    return p3d


def main():
    ref = load_reference_lyaP3D()
    # the substitution itself, executed by the reference: model_Px -> (patched) Px_Mpc_withSiIII sees 1e-4 where k_par was 0
    seen = {}
    orig = ref.Px_Mpc_withSiIII
    ref.Px_Mpc_withSiIII = lambda z, kpar, *a, **k: seen.setdefault("kpar", np.array(kpar)) * 0.0
    obj = ref.LyaP3D(np.array([2.4]), None, None, {"bias": np.array([-0.1])}, Si_contam=True, contam_coeffs={"x": 1})
    obj.model_Px(np.array([[0.0, 0.5]]), np.array([[1.0, 2.0]]))
    ref.Px_Mpc_withSiIII = orig
    assert seen["kpar"][0, 0] == 1e-4 and seen["kpar"][0, 1] == 0.5, seen

    out = {"rperp_Mpc": RPERP, "kpar_Mpc": KPAR, "names": np.array(sorted(PARAM_SETS)),
           "kpar0_substitute": np.array(seen["kpar"][0, 0])}
    for name in sorted(PARAM_SETS):
        ps = dict(PARAM_SETS[name])
        z = ps.pop("z")
        bao_amp = ps.pop("bao_amp", 0.0)
        cosmo = Cosmology({"bao_amp": bao_amp}) if bao_amp else Cosmology()
        p3d = arinyo_p3d(cosmo)

        class Model(object):
            @staticmethod
            def linP_Mpc(z, k, cosmo=cosmo):
                return cosmo.get_linP_Mpc(z, k)

        pp = {k: np.array([v]) for k, v in ps.items()}        # the reference indexes its coefficients by z bin
        zero = np.array([0.0])
        px = ref.Px_Mpc_withSiIII(np.array([z]), KPAR, RPERP, Model, lambda zz, k, mu, p: p3d(zz, k, mu, {kk: float(np.ravel(vv)[0]) for kk, vv in p.items()}),
                                  P3D_mode="pol", Si_coeffs={"bias_SiIII": zero, "beta_SiIII": zero + 1.0, "k_p_SiIII": zero + 10.0},
                                  P3D_params=pp)
        px = np.asarray(px)[0]                                   # [Nr, Nk]
        out["px|" + name] = px
        out["params|" + name] = np.array([z] + [PARAM_SETS[name][k] for k in ("bias", "beta", "q1", "q2", "kv", "av", "bv", "kp")] + [bao_amp])
        print(name, px.shape, "Px range %.3e .. %.3e" % (px.min(), px.max()))
    np.savez_compressed(os.path.join(HERE, "ref_hankel_lyaP3D.npz"), **out)
    print("wrote ref_hankel_lyaP3D.npz")


if __name__ == "__main__":
    main()
