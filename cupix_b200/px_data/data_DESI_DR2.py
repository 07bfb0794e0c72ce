"""Reader/writer for DESI-DR2-style Px files (real data, CoLoRe mocks, forecasts).

Mirror of ``cupix/px_data/data_DESI_DR2.py``.  HDF5 layout (:66-99): ``metadata.attrs{k_m,
k_M_edges, theta_min_a, theta_max_a, theta_min_A, theta_max_A, z_centers, N_fft, L_fft}``,
``metadata/B_A_a``, ``P_Z_AM/z_i/theta_rebin_j``, ``C_Z_AMN/z_i/theta_rebin_j``,
``U_Z_aMn/z_i/theta_j``, ``V_Z_aM/z_i/theta_j``.  The file is opened once (the reference re-opens
it per dataset).  h5py is optional here, so the same content can also live in an ``.npz`` twin
with flat keys (``k_m, k_M_edges, theta_min_a, ..., B_A_a, Px_ZAM, cov_ZAM, U_ZaMn, V_ZaM``).
"""
import numpy as np

from cupix_b200.px_data.base_px_data import BaseDataPx

_META = ("k_m", "k_M_edges", "theta_min_a", "theta_max_a", "theta_min_A", "theta_max_A",
         "z_centers", "N_fft", "L_fft")


def _read_hdf5(filepath):
    import h5py
    d = {}
    with h5py.File(filepath, "r") as f:
        for key in _META:
            d[key] = f["metadata"].attrs[key]
        d["B_A_a"] = f["metadata/B_A_a"][:]
        Nz, N_A, N_a = len(d["z_centers"]), len(d["theta_min_A"]), len(d["theta_min_a"])
        d["Px_ZAM"] = np.array([[f[f"P_Z_AM/z_{iz}/theta_rebin_{A}/"][:] for A in range(N_A)] for iz in range(Nz)])
        d["cov_ZAM"] = np.array([[f[f"C_Z_AMN/z_{iz}/theta_rebin_{A}/"][:] for A in range(N_A)] for iz in range(Nz)])
        d["U_ZaMn"] = np.array([[f[f"U_Z_aMn/z_{iz}/theta_{a}/"][:] for a in range(N_a)] for iz in range(Nz)])
        d["V_ZaM"] = np.array([[f[f"V_Z_aM/z_{iz}/theta_{a}/"][:] for a in range(N_a)] for iz in range(Nz)])
    return d


def write_px_file(filepath, d, extra_attrs=None):
    """Write the dict produced by ``_read_hdf5``/``synthetic`` as .hdf5 (needs h5py) or .npz."""
    if str(filepath).endswith(".npz"):
        np.savez(filepath, **d, **({} if not extra_attrs else {"attr_" + k: v for k, v in extra_attrs.items()}))
        return
    import h5py
    with h5py.File(filepath, "w") as f:
        meta = f.create_group("metadata")
        for key in _META:
            meta.attrs[key] = d[key]
        for k, v in (extra_attrs or {}).items():
            meta.attrs[k] = v
        meta.create_dataset("B_A_a", data=d["B_A_a"])
        for iz in range(len(d["z_centers"])):
            for A in range(len(d["theta_min_A"])):
                f.create_dataset(f"P_Z_AM/z_{iz}/theta_rebin_{A}", data=d["Px_ZAM"][iz][A])
                f.create_dataset(f"C_Z_AMN/z_{iz}/theta_rebin_{A}", data=d["cov_ZAM"][iz][A])
            for a in range(len(d["theta_min_a"])):
                f.create_dataset(f"U_Z_aMn/z_{iz}/theta_{a}", data=d["U_ZaMn"][iz][a])
                f.create_dataset(f"V_Z_aM/z_{iz}/theta_{a}", data=d["V_ZaM"][iz][a])


def data_to_dict(data):
    """The dict layout of ``_read_hdf5`` / ``synthetic`` from a loaded (possibly cut) data object."""
    return {"k_m": np.asarray(data.k_m), "k_M_edges": np.asarray(data.k_M_edges),
            "theta_min_a": np.asarray(data.theta_min_a_arcmin), "theta_max_a": np.asarray(data.theta_max_a_arcmin),
            "theta_min_A": np.asarray(data.theta_min_A_arcmin), "theta_max_A": np.asarray(data.theta_max_A_arcmin),
            "z_centers": np.asarray(data.z), "N_fft": data.N_fft, "L_fft": data.L_fft,
            "B_A_a": np.asarray(data.B_A_a), "Px_ZAM": np.asarray(data.Px_ZAM), "cov_ZAM": np.asarray(data.cov_ZAM),
            "U_ZaMn": np.asarray(data.U_ZaMn), "V_ZaM": np.asarray(data.V_ZaM)}


def write_forecast(likes, savepath, params={}, add_noise=False, seed=None):
    """Forecast file: the measurement of ``likes[i].data`` with P_Z_AM replaced by the theory of each z bin
    convolved with the data's own window (``Likelihood.generate_px_forecast``), noise from the data covariance
    optional -- what notebooks/examples/generate_forecast_data.py:132-173 writes.  Windows, covariances and
    binning are copied; per z bin the default Lya parameters and the model label used are stored next to the
    data vector (HDF5: attributes of ``P_Z_AM/z_i`` and its ``lya_params`` group; .npz: ``z{i}_lya_<name>``,
    ``z{i}_default_lya_model``), plus ``z_centers_orig``.  Returns the forecast Px_ZAM."""
    data = likes[0].data
    d = data_to_dict(data)
    rng = np.random.default_rng(seed)
    px = []
    for like in likes:
        m = like.get_convolved_px(params=params)
        if add_noise:                                  # likelihood.py:42-49, with a seeded generator
            L = np.linalg.cholesky(data.cov_ZAM[like.iz])
            m = m + np.einsum("aij,aj->ai", L, rng.normal(size=m.shape))
        px.append(m)
    d["Px_ZAM"] = np.array(px)
    extra = {"z_centers_orig": np.asarray(data.z), "forecast_add_noise": bool(add_noise)}
    per_z = []
    for like in likes:
        lm = like.theory.lya_model
        info = {"default_lya_model": str(lm.default_lya_model)}
        defaults = getattr(lm, "default_lya_params", None) or {}
        info.update({"lya_" + k: float(v) for k, v in defaults.items()})
        info.update({"forecast_" + k: float(v) for k, v in params.items()})
        per_z.append(info)
    if str(savepath).endswith(".npz"):
        for iz, info in enumerate(per_z):
            extra.update({"z%d_%s" % (iz, k): v for k, v in info.items()})
        write_px_file(savepath, d, extra_attrs=extra)
    else:
        import h5py
        write_px_file(savepath, d, extra_attrs=extra)
        with h5py.File(savepath, "a") as f:
            for iz, info in enumerate(per_z):
                g = f["P_Z_AM/z_%d" % iz]
                g.attrs["default_lya_model"] = info["default_lya_model"]
                lp = g.create_group("lya_params")
                for k, v in info.items():
                    if k.startswith("lya_"):
                        lp.attrs[k[4:]] = v
    return d["Px_ZAM"]


class DESI_DR2(BaseDataPx):
    """Px measurement in the DESI DR2 analysis format.  ``filepath`` may be an .hdf5/.h5 file, an
    .npz twin, or an already-loaded dict with the same keys."""

    def __init__(self, filepath, theta_min_cut_arcmin=None, theta_max_cut_arcmin=None,
                 kM_min_cut_AA=None, kM_max_cut_AA=None, km_min_cut_AA=None, km_max_cut_AA=None):
        if isinstance(filepath, dict):
            d, path = filepath, None
        elif str(filepath).endswith(".npz"):
            with np.load(filepath) as f:
                d = {k: f[k] for k in f.files}
            path = filepath
        else:
            d, path = _read_hdf5(filepath), filepath
        self._file = d            # the file's content before any cut: what the reference's per-dataset getters return
        super().__init__(d["Px_ZAM"], d["cov_ZAM"], d["z_centers"], d["k_M_edges"],
                         d["theta_min_A"], d["theta_max_A"], d["N_fft"], d["L_fft"],
                         has_theta_rebinning=True, has_k_rebinning=True, k_m_AA=d["k_m"],
                         theta_min_a_arcmin=d["theta_min_a"], theta_max_a_arcmin=d["theta_max_a"],
                         B_A_a=d["B_A_a"], U_ZaMn=d["U_ZaMn"], V_ZaM=d["V_ZaM"], filepath=path,
                         theta_min_cut_arcmin=theta_min_cut_arcmin, theta_max_cut_arcmin=theta_max_cut_arcmin,
                         kM_min_cut_AA=kM_min_cut_AA, kM_max_cut_AA=kM_max_cut_AA,
                         km_min_cut_AA=km_min_cut_AA, km_max_cut_AA=km_max_cut_AA)

    # -- the reference's per-dataset getters (data_DESI_DR2.py:66-99; there each call re-opens the file) -------------
    def read_from_file(self):
        d = self._file
        return (d["k_m"], d["k_M_edges"], d["theta_min_a"], d["theta_max_a"], d["theta_min_A"], d["theta_max_A"],
                d["z_centers"], d["N_fft"], d["L_fft"], d["B_A_a"])

    def get_Px_z_T(self, zbin_ind, theta_bin_ind):
        return np.array(self._file["Px_ZAM"][zbin_ind][theta_bin_ind])

    def get_cov_matrix_z_T(self, zbin_ind, theta_bin_ind):
        return np.array(self._file["cov_ZAM"][zbin_ind][theta_bin_ind])

    def get_window_matrix_z_t(self, zbin_ind, theta_bin_ind):
        return np.array(self._file["U_ZaMn"][zbin_ind][theta_bin_ind])

    def get_V_Z_aM(self, zbin_ind, theta_bin_ind):
        return np.array(self._file["V_ZaM"][zbin_ind][theta_bin_ind])
