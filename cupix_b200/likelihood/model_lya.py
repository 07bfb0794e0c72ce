"""Default values and per-call merging of the clean Lya P3D parameters.

Host-side mirror of ``cupix/likelihood/model_lya.py`` (LyaModel): same constructor, same config
keys (``default_lya_model``, ``emulator_label``, ``Nrealizations`` plus any Arinyo name as an
override, model_lya.py:42-56,150-152), same ``get_lya_params(cosmo, params)`` contract
(defaults, overridden by matching keys of ``params``; unknown keys ignored, model_lya.py:190-196).
The eight values it returns are slots 0-7 of the device parameter vector.
"""
import json

import numpy as np

from cupix_b200.utils.utils import data_path

LYA_PARAM_NAMES = ("bias", "beta", "q1", "q2", "kv_Mpc", "av", "bv", "kp_Mpc")
IGM_PARAM_NAMES = ("mF", "gamma", "sigT_Mpc", "kF_Mpc")                                 # model_lya.py:205-208
EMU_INPUT_NAMES = ("Delta2_p", "n_p") + IGM_PARAM_NAMES                                  # model_lya.py:203-222
FF_OUTPUT_NAMES = ("bias", "beta", "q1", "kvav", "av", "bv", "kp", "q2")                 # ForestFlow's Arinyo naming


def lya_params_from_forestflow_params(ff_params):
    """ForestFlow names -> cupix names: kp -> kp_Mpc, kvav -> kv_Mpc = exp(ln(kvav)/av)
    (model_lya.py:13-18)."""
    lya_params = dict(ff_params)
    lya_params["kp_Mpc"] = lya_params.pop("kp")
    kvav = lya_params.pop("kvav")
    lya_params["kv_Mpc"] = np.exp(np.log(kvav) / lya_params["av"])
    return lya_params


def _read_training_info():
    import csv
    with open(data_path("ff_training_info.csv")) as f:
        rows = list(csv.DictReader(f))
    return {k: np.array([float(r[k]) for r in rows]) for k in rows[0]}


class LyaModel(object):
    """Arinyo-parameter bookkeeping for one redshift bin."""

    def __init__(self, z, config={'verbose': False}):
        self.z = z
        self.lr_lya = 1215.67
        self._setup_from_config(config)

    def _setup_from_config(self, config):
        """Default parameter values, or IGM defaults plus an emulator (model_lya.py:35-57)."""
        self.verbose = config.get('verbose', False)
        self.default_lya_model = config.get('default_lya_model', 'best_fit_arinyo_from_p1d')
        if 'igm' in self.default_lya_model:
            self.default_igm_params = self.get_default_igm_params(config)
            self.default_lya_params = None
            # config['emulator']: any object with ForestFlow's ``predict_Arinyos(emu_params=dict) -> dict`` (and, for
            # batches, ``predict_Arinyos_tensor(x[B, 6]) -> [B, 8]`` in torch, columns EMU_INPUT_NAMES -> FF_OUTPUT_NAMES);
            # default: ForestFlow's cINN, as in the reference
            self.emulator = config['emulator'] if config.get('emulator') is not None else \
                self.get_emulator(config.get('emulator_label', 'forest_mpg'), config.get('Nrealizations', 3000))
        else:
            self.default_lya_params = self.get_default_lya_params(config)
            self.default_igm_params = None
            self.emulator = None

    # -- defaults -------------------------------------------------------------------------
    def _closest_training_row(self):
        info = _read_training_info()
        dz = np.abs(info['z'] - self.z)
        iz = int(np.argmin(dz))
        assert dz[iz] < 0.2, "could not find training info for redshift {}, cannot use Gadget default theory".format(self.z)
        return info, iz

    def get_default_igm_params(self, config):
        name = self.default_lya_model.lower()
        if 'gadget' in name:
            info, iz = self._closest_training_row()
            igm_params = {p: info[p + "_central"][iz]
                          for p in ('Delta2_p', 'n_p', 'mF', 'gamma', 'sigT_Mpc', 'kF_Mpc')}
        elif 'p1d' in name:
            from forestflow import priors          # optional third-party dependency
            igm_params = priors.get_IGM_priors(z=self.z, tag='DESI_DR1_P1D')['mean']
        else:
            raise ValueError("unknown default_lya_model", self.default_lya_model)
        for par in igm_params:
            if par in config:
                igm_params[par] = config[par]
        return igm_params

    def get_default_lya_params(self, config):
        name = self.default_lya_model.lower()
        if 'colore' in name:
            assert self.z in [2.2, 2.4, 2.6, 2.8], "We only have CoLoRe fits for redshifts in [2.2, 2.4, 2.6, 2.8]"
            with open(data_path("colore_arinyo_fits.json")) as f:
                hdr = json.load(f)["{:.1f}".format(self.z)]
            ff = {'bias': hdr['bias_LYA'], 'beta': hdr['beta_LYA'], 'q1': hdr['dnl_arinyo_q1'],
                  'kvav': hdr['dnl_arinyo_kv'] ** hdr['dnl_arinyo_av'], 'av': hdr['dnl_arinyo_av'],
                  'bv': hdr['dnl_arinyo_bv'], 'kp': hdr['dnl_arinyo_kp'],
                  'q2': hdr.get('dnl_arinyo_q2', 0)}
            if 'pressure_only' in name:
                ff['q1'] = 0.0
                ff['q2'] = 0.0
                ff['kp'] = {2.2: 0.325, 2.4: 0.315}.get(self.z, 0.300)
        elif 'p1d' in name:
            from forestflow import priors          # optional third-party dependency
            ff = priors.get_arinyo_priors(z=self.z, tag='DESI_DR1_P1D')['mean']
        elif 'gadget' in name:
            info, iz = self._closest_training_row()
            ff = {p: info[p + "_central"][iz]
                  for p in ('bias', 'beta', 'q1', 'kvav', 'av', 'bv', 'kp', 'q2')}
        else:
            raise ValueError("unknown default_lya_model", self.default_lya_model)
        lya_params = lya_params_from_forestflow_params(ff)
        for par in lya_params:
            if par in config:
                lya_params[par] = config[par]
        return lya_params

    def get_emulator(self, emulator_label, Nrealizations):
        """ForestFlow cINN (model_lya.py:159-171).  Runs in PyTorch; its [B, 8] Arinyo output is
        handed to the device path as a tensor (``Likelihood.get_chi2_from_arinyo_tensor``)."""
        if emulator_label != "forest_mpg":
            raise ValueError("implement emulator_label", emulator_label)
        import forestflow                          # optional third-party dependency
        from forestflow.P3D_cINN import P3DEmulator
        path_program = forestflow.__path__[0][:-10]
        return P3DEmulator(model_path=path_program + "/data/emulator_models/forest_mpg",
                           Nrealizations=Nrealizations)

    # -- per-call values --------------------------------------------------------------------
    def get_lya_params(self, cosmo, params):
        if self.default_igm_params is not None:
            assert self.emulator is not None, "need emulator for IGM params"
            igm_params = dict(self.default_igm_params)
            for key in igm_params:
                if key in params:
                    igm_params[key] = params[key]
            return self.emulate_lya_params(cosmo, igm_params)
        lya_params = dict(self.default_lya_params)
        for key in lya_params:
            if key in params:
                lya_params[key] = params[key]
        return lya_params

    def emulate_lya_params_tensor(self, cosmo, igm_names, igm_values):
        """Batched ``emulate_lya_params``: ``igm_values`` [B, len(igm_names)] (torch, any device) overrides the default IGM
        parameters column by column; returns the emulator's Arinyo parameters [B, 8] in ForestFlow's naming
        (FF_OUTPUT_NAMES) on the same device.  Uses ``emulator.predict_Arinyos_tensor`` when the emulator has it (one
        network call for the batch); otherwise falls back to one ``predict_Arinyos`` call per row."""
        import torch
        assert self.default_igm_params is not None and self.emulator is not None, "need emulator for IGM params"
        B = igm_values.shape[0]
        lin = cosmo.get_linP_Mpc_params(z=self.z, kp_Mpc=0.7)
        base = dict(self.default_igm_params, Delta2_p=lin['Delta2_p'], n_p=lin['n_p'])
        x = torch.tensor([float(base[k]) for k in EMU_INPUT_NAMES], dtype=igm_values.dtype, device=igm_values.device).repeat(B, 1)
        for j, name in enumerate(igm_names):
            assert name in IGM_PARAM_NAMES, "not an IGM parameter: %r" % name
            x[:, EMU_INPUT_NAMES.index(name)] = igm_values[:, j]
        if hasattr(self.emulator, "predict_Arinyos_tensor"):
            return self.emulator.predict_Arinyos_tensor(x)
        rows = []
        for row in x.cpu().numpy():
            ff = self.emulator.predict_Arinyos(emu_params=dict(zip(EMU_INPUT_NAMES, row)))
            rows.append([float(ff[k]) for k in FF_OUTPUT_NAMES])
        return torch.tensor(rows, dtype=igm_values.dtype, device=igm_values.device)

    def emulate_lya_params(self, cosmo, igm_params):
        emu_params = {k: igm_params[k] for k in ('mF', 'gamma', 'sigT_Mpc', 'kF_Mpc')}
        linP_params = cosmo.get_linP_Mpc_params(z=self.z, kp_Mpc=0.7)
        emu_params['Delta2_p'] = linP_params['Delta2_p']
        emu_params['n_p'] = linP_params['n_p']
        ff_params = self.emulator.predict_Arinyos(emu_params=emu_params)
        return lya_params_from_forestflow_params(ff_params)
