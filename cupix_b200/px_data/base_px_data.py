"""Container for a binned Px measurement: the arrays that get uploaded once to HBM.

Host-side mirror of ``cupix/px_data/base_px_data.py`` (BaseDataPx): same constructor arguments,
same public attributes (``z, k_m, k_M_edges, theta_{min,max}_{a,A}_arcmin, B_A_a, Px_ZAM, cov_ZAM,
U_ZaMn, V_ZaM``), same range cuts (``limit_theta_range`` :98-140, ``limit_k_M_range`` :142-184,
``limit_k_m_range`` :186-205).  Index letters: Z redshift bin, a/A fine/coarse theta bin,
m/M fine/coarse k_par bin.
"""
from warnings import warn

import numpy as np


def _as_array(x):
    return None if x is None else np.asarray(x)


class BaseDataPx(object):

    def __init__(self, Px_ZAM, cov_ZAM, z, k_M_edges_AA, theta_min_A_arcmin, theta_max_A_arcmin,
                 N_fft, L_fft, has_theta_rebinning=True, has_k_rebinning=True, k_m_AA=None,
                 theta_min_a_arcmin=None, theta_max_a_arcmin=None, B_A_a=None, U_ZaMn=None,
                 V_ZaM=None, filepath=None, theta_min_cut_arcmin=None, theta_max_cut_arcmin=None,
                 kM_min_cut_AA=None, kM_max_cut_AA=None, km_min_cut_AA=None, km_max_cut_AA=None):
        self.z = np.array(z)
        self.Nz = len(self.z)
        self.N_fft, self.L_fft = N_fft, L_fft
        self.filepath = filepath
        self.theta_min_A_arcmin = np.asarray(theta_min_A_arcmin, dtype=np.float64)
        self.theta_max_A_arcmin = np.asarray(theta_max_A_arcmin, dtype=np.float64)
        self.theta_min_a_arcmin = _as_array(theta_min_a_arcmin)
        self.theta_max_a_arcmin = _as_array(theta_max_a_arcmin)
        self.has_k_rebinning = bool(has_k_rebinning)
        self.has_theta_rebinning = bool(has_theta_rebinning)

        if has_k_rebinning:
            assert k_m_AA is not None, "k_m_AA must be provided if has_k_rebinning is True"
            k_m_AA = np.asarray(k_m_AA, dtype=np.float64)
            k_M_edges_AA = np.asarray(k_M_edges_AA, dtype=np.float64)
            if k_m_AA.ndim == 1:
                # one wavenumber grid shared by all redshift bins -> one row per z
                self.k_m = np.tile(k_m_AA, (self.Nz, 1))
                self.k_M_edges = np.tile(k_M_edges_AA, (self.Nz, 1))
            else:
                self.k_m, self.k_M_edges = k_m_AA, k_M_edges_AA
            self.k_M_centers_AA = 0.5 * (self.k_M_edges[:, :-1] + self.k_M_edges[:, 1:])
        if has_theta_rebinning:
            assert theta_min_a_arcmin is not None, "theta_min_a_arcmin must be provided if has_theta_rebinning is True"
            assert theta_max_a_arcmin is not None, "theta_max_a_arcmin must be provided if has_theta_rebinning is True"
            assert B_A_a is not None, "B_A_a must be provided if has_theta_rebinning is True"
            self.B_A_a = np.asarray(B_A_a)
        self.Px_ZAM = np.asarray(Px_ZAM, dtype=np.float64)
        self.cov_ZAM = np.asarray(cov_ZAM, dtype=np.float64)
        self.U_ZaMn = _as_array(U_ZaMn)
        self.V_ZaM = _as_array(V_ZaM)

        if theta_min_cut_arcmin is not None or theta_max_cut_arcmin is not None:
            self.limit_theta_range(theta_min_cut_arcmin, theta_max_cut_arcmin)
        if kM_min_cut_AA is not None or kM_max_cut_AA is not None:
            self.limit_k_M_range(kM_min_cut_AA, kM_max_cut_AA)
        if km_min_cut_AA is not None or km_max_cut_AA is not None:
            self.limit_k_m_range(km_min_cut_AA, km_max_cut_AA)
        self.theta_centers_arcmin = 0.5 * (self.theta_min_A_arcmin + self.theta_max_A_arcmin)

    # ---- cuts ----------------------------------------------------------------------------
    def limit_theta_range(self, theta_min_arcmin=None, theta_max_arcmin=None):
        """Keep the coarse (and fine) theta bins fully inside [theta_min, theta_max]."""
        if theta_min_arcmin is None and theta_max_arcmin is None:
            warn("No theta limits provided. No changes made.")
            return
        lo = self.theta_min_A_arcmin[0] if theta_min_arcmin is None else theta_min_arcmin
        hi = self.theta_max_A_arcmin[-1] if theta_max_arcmin is None else theta_max_arcmin
        keep_A = np.where((self.theta_min_A_arcmin >= lo) & (self.theta_max_A_arcmin <= hi))[0]
        if keep_A.size == 0:
            raise ValueError("No theta bins found within the specified range.")
        if self.has_theta_rebinning:
            keep_a = np.where((self.theta_min_a_arcmin >= lo) & (self.theta_max_a_arcmin <= hi))[0]
            if keep_a.size == 0:
                raise ValueError("No theta bins found within the specified range.")
            self.theta_min_a_arcmin = self.theta_min_a_arcmin[keep_a]
            self.theta_max_a_arcmin = self.theta_max_a_arcmin[keep_a]
            self.U_ZaMn = self.U_ZaMn[:, keep_a]
            self.V_ZaM = self.V_ZaM[:, keep_a]
            self.B_A_a = self.B_A_a[np.ix_(keep_A, keep_a)]
        self.theta_min_A_arcmin = self.theta_min_A_arcmin[keep_A]
        self.theta_max_A_arcmin = self.theta_max_A_arcmin[keep_A]
        self.Px_ZAM = self.Px_ZAM[:, keep_A]
        self.cov_ZAM = self.cov_ZAM[:, keep_A]

    def limit_k_M_range(self, k_min_AA=None, k_max_AA=None):
        """Keep the coarse k bins inside [k_min, k_max] (same edge conventions as the reference:
        only-max compares the lower edge, both compares lower and upper edges)."""
        if k_min_AA is None and k_max_AA is None:
            warn("No k limits provided. No changes made.")
            return
        edges, px, cov, U, V = [], [], [], [], []
        for iz in range(self.Nz):
            lo_e, hi_e = self.k_M_edges[iz][:-1], self.k_M_edges[iz][1:]
            if k_min_AA is None:
                keep = np.where(lo_e <= k_max_AA)[0]
            elif k_max_AA is None:
                keep = np.where(lo_e >= k_min_AA)[0]
            else:
                keep = np.where((lo_e >= k_min_AA) & (hi_e <= k_max_AA))[0]
            if keep.size == 0:
                raise ValueError("No k bins found within the specified range.")
            edges.append(self.k_M_edges[iz][np.concatenate(([keep[0]], keep + 1))])
            px.append(self.Px_ZAM[iz][:, keep])
            cov.append(self.cov_ZAM[iz][:, keep][:, :, keep])
            if self.has_theta_rebinning:
                U.append(self.U_ZaMn[iz][:, keep])
                V.append(self.V_ZaM[iz][:, keep])
        self.k_M_edges = np.array(edges)
        self.k_M_centers_AA = 0.5 * (self.k_M_edges[:, :-1] + self.k_M_edges[:, 1:])
        self.Px_ZAM, self.cov_ZAM = np.array(px), np.array(cov)
        if self.has_theta_rebinning:
            self.U_ZaMn, self.V_ZaM = np.array(U), np.array(V)

    def limit_k_m_range(self, k_min_AA=None, k_max_AA=None):
        """Keep the fine k_m inside [k_min, k_max]; drops the matching window-matrix columns.
        (As in the reference the last k_m is never kept: the test looks at k_m[:-1], k_m[1:].)"""
        if k_min_AA is None and k_max_AA is None:
            warn("No k limits provided. No changes made.")
            return
        k_m, U = [], []
        for iz in range(self.Nz):
            lo_k, hi_k = self.k_m[iz][:-1], self.k_m[iz][1:]
            if k_min_AA is None:
                keep = np.where(lo_k <= k_max_AA)[0]
            elif k_max_AA is None:
                keep = np.where(lo_k >= k_min_AA)[0]
            else:
                keep = np.where((lo_k >= k_min_AA) & (hi_k <= k_max_AA))[0]
            if keep.size == 0:
                raise ValueError("No k bins found within the specified range.")
            k_m.append(self.k_m[iz][keep])
            U.append(self.U_ZaMn[iz][:, :, keep])
        self.k_m, self.U_ZaMn = np.array(k_m), np.array(U)
