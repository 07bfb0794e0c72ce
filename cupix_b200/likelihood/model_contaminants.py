"""Default values and per-call merging of the contaminant parameters.

Host-side mirror of ``cupix/likelihood/model_contaminants.py`` (ContaminantsModel): HCD
(b_H, beta_H, L_H_Mpc; :34-45), SiIII metal (b_X, beta_X; :48-57), sky residuals (b_noise_Mpc;
:60-79), continuum distortion (kC_Mpc, pC; :82-100) and the xi_noise(theta) table (:147-165).
These are slots 8-15 of the device parameter vector.
"""
import numpy as np

from cupix_b200.utils.utils import data_path

CONT_PARAM_NAMES = ("b_H", "beta_H", "L_H_Mpc", "b_X", "beta_X", "b_noise_Mpc", "kC_Mpc", "pC")


class ContaminantsModel(object):

    def __init__(self, z, config={'verbose': False}):
        self.z_lya = z
        self.lr_metal = 1206.52   # SiIII
        self.setup_from_config(config)

    def setup_from_config(self, config):
        """Defaults of all contaminant parameters and the xi_noise table (model_contaminants.py:17-31)."""
        self.verbose = config.get('verbose', False)
        self.default_hcd_params = self.get_default_hcd_params(config)
        self.default_metal_params = self.get_default_metal_params(config)
        self.default_sky_params = self.get_default_sky_params(config)
        self.default_continuum_params = self.get_default_continuum_params(config)
        self._xi_theta, self._xi_val = self.read_xi_noise(config)

    def get_default_hcd_params(self, config):
        """L_H in Mpc, not Mpc/h (model_contaminants.py:34-45)."""
        return self._with_config({'b_H': -0.02, 'beta_H': 0.5, 'L_H_Mpc': 5.0 / 0.67}, config)

    def get_default_metal_params(self, config):
        """model_contaminants.py:48-57"""
        return self._with_config({'b_X': -0.005, 'beta_X': 0.5}, config)

    def get_default_sky_params(self, config):
        """Preliminary DESI DR2 fits, by redshift (model_contaminants.py:60-79)."""
        return self._with_config({'b_noise_Mpc': 0.0035 if self.z_lya == 2.2 else 0.00125}, config)

    def get_default_continuum_params(self, config):
        """Preliminary fits on mocks, by redshift (model_contaminants.py:82-100)."""
        cont = {'kC_Mpc': 0.019, 'pC': 0.63} if self.z_lya == 2.2 else {'kC_Mpc': 0.012, 'pC': 0.45}
        return self._with_config(cont, config)

    @staticmethod
    def _with_config(defaults, config):
        for par in defaults:
            if par in config:
                defaults[par] = config[par]
        return defaults

    @staticmethod
    def _merge(defaults, params):
        out = dict(defaults)
        for key in out:
            if key in params:
                out[key] = params[key]
        return out

    def get_hcd_params(self, params):
        return self._merge(self.default_hcd_params, params)

    def get_metal_params(self, params):
        return self._merge(self.default_metal_params, params)

    def get_sky_params(self, params):
        return self._merge(self.default_sky_params, params)

    def get_continuum_params(self, params):
        return self._merge(self.default_continuum_params, params)

    def read_xi_noise(self, config):
        tab = np.loadtxt(data_path("desi-instrument-syst-for-forest-auto-correlation_arcmin.csv"),
                         delimiter=",", skiprows=1)
        return tab[:, 0].copy(), tab[:, 1].copy()

    def get_xi_noise(self, theta_arc):
        """Linear interpolation of the tabulated xi_noise (scipy interp1d(kind='linear') in the
        reference, which raises outside the table; so do we)."""
        theta_arc = np.asarray(theta_arc, dtype=np.float64)
        if np.any(theta_arc < self._xi_theta[0]) or np.any(theta_arc > self._xi_theta[-1]):
            raise ValueError("A value in theta_arc is outside the xi_noise interpolation range.")
        return np.interp(theta_arc, self._xi_theta, self._xi_val)
