"""Self-contained stand-in for ``lace.cosmo.cosmology.Cosmology``.

The reference takes its linear power and unit conversions from LaCE/CAMB
(``cupix/likelihood/theory.py:5,132-136,300``), which are not importable here.  ``Theory`` only
ever calls five methods of that object, so any object with the same five methods drops in
(a real ``lace`` Cosmology included):

    get_linP_Mpc(z, k_Mpc)            linear matter power [Mpc^3]
    get_darc_dMpc(z)                  arcmin per transverse comoving Mpc
    get_dAA_dMpc(z, lambda_rest_AA)   observed Angstrom per line-of-sight comoving Mpc
    same_background(cosmo_params)     True if only the primordial power changed
    get_linP_Mpc_params(z, kp_Mpc)    {'Delta2_p', 'n_p'} at the pivot

This implementation is a flat LCDM background with an Eisenstein & Hu (1998) no-wiggle transfer
function and the exact LCDM growth integral.  It is *synthetic-data grade*: smooth, analytic and
deterministic, so that the CPU oracle and the device plan sample exactly the same function.
"""
import numpy as np

C_KMS = 299792.458


class Cosmology(object):
    def __init__(self, cosmo_params_dict=None):
        # bao_amp > 0 superimposes a BAO-like oscillation (period 2 pi / 147 Mpc, Silk-damped) of that relative amplitude:
        # what a CAMB / LaCE linear power looks like to the k_perp quadrature (0.05 is realistic); 0 = smooth
        p = dict(H0=67.66, omch2=0.119, ombh2=0.0224, As=2.105e-9, ns=0.9665, Tcmb=2.7255, bao_amp=0.0)
        if cosmo_params_dict:
            for k, v in cosmo_params_dict.items():
                if k in p:
                    p[k] = float(v)
        self.params = p
        self.h = p["H0"] / 100.0
        self.Om = (p["omch2"] + p["ombh2"]) / self.h ** 2
        self.Ob = p["ombh2"] / self.h ** 2
        self.As = p["As"]
        self.ns = p["ns"]
        self.k_pivot = 0.05
        self._growth_cache = {}
        self._dist_cache = {}

    # ---- background -------------------------------------------------------------------
    def _E(self, z):
        return np.sqrt(self.Om * (1.0 + z) ** 3 + (1.0 - self.Om))

    def get_Hz(self, z):
        return self.params["H0"] * self._E(z)

    def _comoving_distance_Mpc(self, z):
        key = round(float(z), 12)
        if key not in self._dist_cache:
            x, w = np.polynomial.legendre.leggauss(64)
            zz = 0.5 * z * (x + 1.0)
            self._dist_cache[key] = C_KMS / self.params["H0"] * 0.5 * z * np.sum(w / self._E(zz))
        return self._dist_cache[key]

    def get_darc_dMpc(self, z):
        """arcmin subtended by one transverse comoving Mpc at z"""
        return (180.0 * 60.0 / np.pi) / self._comoving_distance_Mpc(z)

    def get_dAA_dMpc(self, z, lambda_rest_AA=1215.67):
        """observed Angstroms per comoving Mpc along the line of sight"""
        return lambda_rest_AA * self.get_Hz(z) / C_KMS

    def get_dkms_dMpc(self, z):
        return self.get_Hz(z) / (1.0 + z)

    # ---- linear power -----------------------------------------------------------------
    def _growth(self, z):
        key = round(float(z), 12)
        if key not in self._growth_cache:
            a = 1.0 / (1.0 + z)
            x, w = np.polynomial.legendre.leggauss(96)
            # D(a) = 5/2 Om E(a) int_0^a da' / (a' E(a'))^3, substituting a' = a s^2 for the a'^(3/2) cusp
            s = 0.5 * (x + 1.0)
            ap = a * s * s
            Ea = np.sqrt(self.Om / ap ** 3 + (1.0 - self.Om))
            integ = np.sum(0.5 * w * 2.0 * a * s / (ap * Ea) ** 3)
            self._growth_cache[key] = 2.5 * self.Om * self._E(z) * integ
        return self._growth_cache[key]

    def _transfer_nowiggle(self, k_Mpc):
        h, Om, Ob = self.h, self.Om, self.Ob
        omh2, obh2, fb = Om * h * h, Ob * h * h, Ob / Om
        theta = self.params["Tcmb"] / 2.7
        s = 44.5 * np.log(9.83 / omh2) / np.sqrt(1.0 + 10.0 * obh2 ** 0.75)
        alpha = 1.0 - 0.328 * np.log(431.0 * omh2) * fb + 0.38 * np.log(22.3 * omh2) * fb ** 2
        gamma = Om * h * (alpha + (1.0 - alpha) / (1.0 + (0.43 * k_Mpc * s) ** 4))
        q = k_Mpc * theta ** 2 / (gamma * h)
        L0 = np.log(2.0 * np.e + 1.8 * q)
        C0 = 14.2 + 731.0 / (1.0 + 62.5 * q)
        return L0 / (L0 + C0 * q * q)

    def get_linP_Mpc(self, z, k_Mpc):
        k = np.asarray(k_Mpc, dtype=np.float64)
        T = self._transfer_nowiggle(k)
        D = self._growth(z)
        c_H0 = C_KMS / self.params["H0"]
        with np.errstate(divide="ignore", invalid="ignore"):
            d2 = (4.0 / 25.0) * self.As * (k / self.k_pivot) ** (self.ns - 1.0) \
                * (k * c_H0) ** 4 * T * T * D * D / self.Om ** 2
            P = np.where(k > 0, 2.0 * np.pi ** 2 * d2 / np.where(k > 0, k, 1.0) ** 3, 0.0)
        if self.params["bao_amp"]:
            P = P * (1.0 + self.params["bao_amp"] * np.sin(k * 147.0) * np.exp(-(k / 0.25) ** 1.4))
        return P

    def get_linP_Mpc_params(self, z, kp_Mpc=0.7):
        eps = 1e-3
        lk = np.log(kp_Mpc)
        lp = np.log(self.get_linP_Mpc(z, np.exp([lk - eps, lk, lk + eps])))
        return {"Delta2_p": kp_Mpc ** 3 * np.exp(lp[1]) / (2 * np.pi ** 2),
                "n_p": (lp[2] - lp[0]) / (2 * eps),
                "alpha_p": (lp[2] - 2 * lp[1] + lp[0]) / eps ** 2}

    def same_background(self, cosmo_params=None):
        """True when none of the background parameters differ from this cosmology."""
        if not cosmo_params:
            return True
        for key in ("H0", "omch2", "ombh2", "mnu", "omk", "w"):
            if key in cosmo_params and key in self.params:
                if not np.isclose(cosmo_params[key], self.params[key], rtol=1e-12, atol=0.0):
                    return False
            elif key in cosmo_params and cosmo_params[key] not in (0, 0.0, -1, -1.0):
                return False
        return True


class RescaledCosmology(Cosmology):
    """Same background, primordial amplitude/tilt replaced (stand-in for
    ``lace.cosmo.rescale_cosmology.RescaledCosmology``, used at theory.py:54-56)."""

    def __init__(self, fid_cosmo, new_params_dict=None):
        p = dict(fid_cosmo.params)
        for key in ("As", "ns"):
            if new_params_dict and key in new_params_dict:
                p[key] = float(new_params_dict[key])
        super().__init__(p)

    def changes_linP(self, fid_cosmo):
        return (self.As != fid_cosmo.As) or (self.ns != fid_cosmo.ns)


def cache_key(cosmo, fid_cosmo):
    """Hashable identity of a cosmology for the engine caches of Theory / Likelihood.

    The fiducial object maps to 0.  Cosmologies of this module are keyed by value, so the
    RescaledCosmology built afresh by every scalar call with the same ``As, ns`` reuses one device
    plan.  Foreign objects (a real ``lace`` Cosmology) are keyed by ``id``; the caches keep a reference
    to the object next to its engine so that the id cannot be recycled for another cosmology while
    the entry lives."""
    if cosmo is None or cosmo is fid_cosmo:
        return 0
    if isinstance(cosmo, Cosmology):
        return ("params",) + tuple(sorted(cosmo.params.items()))
    return ("id", id(cosmo))
