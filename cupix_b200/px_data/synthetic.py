"""Seeded synthetic DESI-shaped Px measurements (bins, windows, weights, covariance).

The reference ships no usable measurement file (its one in-tree forecast is a missing blob), so
tests and benchmarks build their own in the layout of ``data_DESI_DR2.py:12-62``.  Shapes S1-S5 of
SURVEY.md section 8(d).  The Px values are filled in afterwards from a model (the forecast recipe
of ``cupix/likelihood/likelihood.py:37-50``: data = convolved theory (+ L n)).
"""
import numpy as np

SEED = 20251019


def make_bins(Nz=1, z0=2.2, dz=0.2, N_A=20, fine_per_coarse=1, N_m=156, rebin_k=3,
              theta_lo=0.5, theta_hi=60.0, N_fft=1024, pix_AA=0.8, skip_k0=False):
    """Binning arrays.  k_m = 2 pi j / L with L = N_fft * pix; coarse k bins are groups of
    ``rebin_k`` fine modes; theta edges are log-spaced; fine theta bins split each coarse bin."""
    z = np.round(z0 + dz * np.arange(Nz), 10)
    L = N_fft * pix_AA
    j0 = 1 if skip_k0 else 0
    k_m = 2 * np.pi * (np.arange(N_m) + j0) / L
    N_M = N_m // rebin_k
    dk = 2 * np.pi / L
    k_M_edges = k_m[0] - 0.5 * dk + rebin_k * dk * np.arange(N_M + 1)
    eA = np.geomspace(theta_lo, theta_hi, N_A + 1)
    ea = np.concatenate([np.linspace(eA[i], eA[i + 1], fine_per_coarse + 1)[:-1] for i in range(N_A)] + [eA[-1:]])
    N_a = N_A * fine_per_coarse
    B = np.zeros((N_A, N_a))
    for A in range(N_A):
        B[A, A * fine_per_coarse:(A + 1) * fine_per_coarse] = 1.0
    return dict(z_centers=z, k_m=k_m, k_M_edges=k_M_edges, theta_min_a=ea[:-1], theta_max_a=ea[1:],
                theta_min_A=eA[:-1], theta_max_A=eA[1:], N_fft=N_fft, L_fft=L, B_A_a=B)


def make_windows(bins, rng, half_width=8, rebin_k=None):
    """Dense window matrices U[z,a,M,m]: row-normalised |sinc-like band| + 1e-3 U(0,1), and
    weights V[z,a,M] ~ U(0.5, 1.5)."""
    Nz, N_a = len(bins["z_centers"]), len(bins["theta_min_a"])
    N_m, N_M = len(bins["k_m"]), len(bins["k_M_edges"]) - 1
    rebin_k = rebin_k or max(1, N_m // N_M)
    m = np.arange(N_m)[None, :]
    centre = (rebin_k * np.arange(N_M) + 0.5 * (rebin_k - 1))[:, None]
    band = np.abs(np.sinc((m - centre) / half_width)) * (np.abs(m - centre) <= 4 * half_width)
    U = np.empty((Nz, N_a, N_M, N_m))
    for iz in range(Nz):
        for a in range(N_a):
            u = band * (1.0 + 0.1 * rng.uniform(-1, 1, size=band.shape)) + 1e-3 * rng.uniform(size=band.shape)
            U[iz, a] = u / u.sum(axis=1, keepdims=True)
    V = rng.uniform(0.5, 1.5, size=(Nz, N_a, N_M))
    return U, V


def make_covariance(Px_ZAM, rho=0.3, frac=0.05, floor_frac=1e-3):
    """cov = D^1/2 R D^1/2 with R_ij = rho^|i-j| and sqrt(D) = frac |Px| + floor (SPD)."""
    Nz, N_A, N_M = Px_ZAM.shape
    i = np.arange(N_M)
    R = rho ** np.abs(i[:, None] - i[None, :])
    floor = floor_frac * np.max(np.abs(Px_ZAM))
    sig = frac * np.abs(Px_ZAM) + floor
    return sig[..., :, None] * R[None, None] * sig[..., None, :]


def make_measurement_dict(shape="S1", seed=SEED, **overrides):
    """Bins + windows for a named shape, with zero Px and identity covariance as placeholders.
    Use ``fill_from_model`` once a model prediction is available."""
    shapes = {
        "tiny": dict(Nz=1, N_A=4, fine_per_coarse=1, N_m=24, rebin_k=3),
        "tiny_rebin": dict(Nz=2, N_A=3, fine_per_coarse=2, N_m=32, rebin_k=4),
        "S1": dict(Nz=1, N_A=20, fine_per_coarse=1, N_m=156, rebin_k=3),
        "S1_rebin": dict(Nz=1, N_A=20, fine_per_coarse=3, N_m=156, rebin_k=3),
        "S2": dict(Nz=4, N_A=20, fine_per_coarse=1, N_m=156, rebin_k=3),
        "S5": dict(Nz=12, N_A=20, fine_per_coarse=1, N_m=256, rebin_k=4, z0=2.0),
    }
    cfg = dict(shapes[shape])
    cfg.update(overrides)
    rng = np.random.default_rng(seed)
    bins = make_bins(**cfg)
    U, V = make_windows(bins, rng, rebin_k=cfg.get("rebin_k"))
    Nz, N_A, N_M = len(bins["z_centers"]), len(bins["theta_min_A"]), len(bins["k_M_edges"]) - 1
    d = dict(bins)
    d.update(U_ZaMn=U, V_ZaM=V, Px_ZAM=np.zeros((Nz, N_A, N_M)),
             cov_ZAM=np.tile(np.eye(N_M), (Nz, N_A, 1, 1)))
    return d


def fill_from_model(d, model_ZAM, rng=None, add_noise=False):
    """data = model (+ L n), covariance from the model amplitude."""
    d = dict(d)
    cov = make_covariance(model_ZAM)
    px = np.array(model_ZAM, dtype=np.float64, copy=True)
    if add_noise:
        rng = rng or np.random.default_rng(SEED + 1)
        L = np.linalg.cholesky(cov)
        px = px + np.einsum("zaij,zaj->zai", L, rng.normal(size=px.shape))
    d["Px_ZAM"], d["cov_ZAM"] = px, cov
    return d
