"""ctypes binding of the C ABI in include/cupix_b200.h (libcupix_b200.so, built in-tree by
``__graft_entry__.build()``).  There is no CPU fallback: if the library or a CUDA device is
missing, construction raises."""
import ctypes as C
import os

import numpy as np

from cupix_b200 import plan as _plan

_LIB = None
LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "lib", "libcupix_b200.so")

NSLOTS = 18
DEVICE_IO = 1
PRECISE_CELLS = 2
NO_FACTOR = 4
MODEL_WHITENED = 8

c_dp = C.POINTER(C.c_double)
c_ip = C.POINTER(C.c_int32)


class ZbinDesc(C.Structure):
    _fields_ = [
        ("n_a", C.c_int32), ("n_m", C.c_int32), ("n_q", C.c_int32), ("n_A", C.c_int32),
        ("n_M", C.c_int32), ("n_data", C.c_int32), ("flags", C.c_int32), ("reserved", C.c_int32),
        ("dAA_dMpc", C.c_double),
        ("kpar_Mpc", c_dp), ("kperp_Mpc", c_dp), ("hankel_w", c_dp), ("linP_cell", c_dp),
        ("metal_auto", c_dp), ("metal_cross", c_dp), ("sky_xi_a", c_dp), ("sky_smooth_m", c_dp),
        ("defaults", C.c_double * NSLOTS),
        ("B_A_a", c_dp), ("U_aMn", c_dp), ("V_aM", c_dp), ("Px_AM", c_dp), ("cov_AMM", c_dp),
        ("k_pivot_Mpc", C.c_double),
        ("kpar_row_Mpc", c_dp),
        ("metal_auto_dn", c_dp), ("metal_cross_dn", c_dp), ("metal_dn_order", C.c_int32), ("reserved2", C.c_int32),
    ]


class EvalArgs(C.Structure):
    _fields_ = [
        ("B", C.c_int64), ("P", C.c_int32), ("flags", C.c_int32),
        ("points", C.c_void_p), ("slot", c_ip), ("zsel", c_ip),
        ("grid_lo", c_dp), ("grid_hi", c_dp), ("grid_n", c_ip), ("grid_first", C.c_int64),
        ("data_idx", C.c_void_p), ("z_list", c_ip), ("n_z_list", C.c_int32), ("reserved", C.c_int32),
        ("chi2", C.c_void_p), ("chi2_zA", C.c_void_p), ("model_zAM", C.c_void_p), ("px_zam", C.c_void_p),
        ("stream", C.c_void_p), ("stage_ms", C.POINTER(C.c_float)),
    ]


class CpxError(RuntimeError):
    pass


def load_library(path=None):
    """dlopen the in-tree shared library; raises with build instructions if it is absent."""
    global _LIB
    if _LIB is not None and path is None:
        return _LIB
    path = path or os.environ.get("CUPIX_B200_LIB") or LIB_PATH      # CUPIX_B200_LIB: another build of the same library (A/B runs)
    if not os.path.exists(path):
        raise CpxError("%s not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                       "(cupix_b200 has no CPU fallback)" % path)
    lib = C.CDLL(path)
    lib.cpx_create.argtypes = [C.POINTER(ZbinDesc), C.c_int32, C.c_int32, C.POINTER(C.c_void_p)]
    lib.cpx_create.restype = C.c_int
    lib.cpx_eval.argtypes = [C.c_void_p, C.POINTER(EvalArgs)]
    lib.cpx_eval.restype = C.c_int
    lib.cpx_set_data.argtypes = [C.c_void_p, C.c_int32, c_dp, C.c_int32]
    lib.cpx_set_data.restype = C.c_int
    lib.cpx_destroy.argtypes = [C.c_void_p]
    lib.cpx_destroy.restype = None
    lib.cpx_last_error.argtypes = []
    lib.cpx_last_error.restype = C.c_char_p
    lib.cpx_version.argtypes = []
    lib.cpx_version.restype = C.c_int
    lib.cpx_get_info.argtypes = [C.c_void_p, C.c_int32, C.POINTER(C.c_int64)]
    lib.cpx_get_info.restype = C.c_int
    lib.cpx_probe_peaks.argtypes = [C.c_int32, c_dp, c_dp, c_dp, c_dp]
    lib.cpx_probe_peaks.restype = C.c_int
    lib.cpx_generate_mocks.argtypes = [C.c_void_p, C.c_int32, c_ip, c_ip, c_dp, C.c_int32, C.c_uint64, c_dp]
    lib.cpx_generate_mocks.restype = C.c_int
    lib.cpx_device_count.argtypes = []
    lib.cpx_device_count.restype = C.c_int
    lib.cpx_hankel_jbar.argtypes = [C.c_int32, C.c_int32, c_ip, c_dp, c_dp, C.c_int64, c_dp, c_dp, c_dp]
    lib.cpx_hankel_jbar.restype = C.c_int
    lib.cpx_hankel_apply.argtypes = [C.c_int32, C.c_int32, C.c_int32, C.c_int32, c_dp, c_dp, c_dp]
    lib.cpx_hankel_apply.restype = C.c_int
    lib.cpx_probe_imma.argtypes = [C.c_int32, C.POINTER(C.c_double)]
    lib.cpx_probe_imma.restype = C.c_int
    _LIB = lib
    return lib


def _dp(a):
    return None if a is None else a.ctypes.data_as(c_dp)


def _f64(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float64)


def probe_imma(device=0):
    """INT8 mma.sync rate of this GPU in integer operations per second (cpx_probe_imma)."""
    lib = load_library()
    v = C.c_double(0.0)
    rc = lib.cpx_probe_imma(int(device), C.byref(v))
    if rc != 0:
        raise CpxError("cpx_probe_imma failed (%d): %s" % (rc, lib.cpx_last_error().decode()))
    return v.value


def probe_peaks(device=0):
    """Measured FP64-FMA, FP32-FMA, MUFU.EX2 and FP64-tensor (DMMA, in FMAs) rates of the GPU [ops/s]."""
    lib = load_library()
    d, f, m, t = C.c_double(), C.c_double(), C.c_double(), C.c_double()
    if lib.cpx_probe_peaks(device, C.byref(d), C.byref(f), C.byref(m), C.byref(t)) != 0:
        raise CpxError(lib.cpx_last_error().decode())
    return {"dfma_per_s": d.value, "ffma_per_s": f.value, "mufu_per_s": m.value, "dmma_fma_per_s": t.value}


def device_count():
    """CUDA devices visible to the library (0 when there is none; raises CpxError if the library is not built)."""
    return int(load_library().cpx_device_count())


def hankel_jbar(device, group_start, r_Mpc, weight, kf, meas):
    """jbar[g, f] = meas[f] sum_{i in g} w_i J0(kf[f] r_i) / sum_{i in g} w_i on the GPU (cpx_hankel_jbar): the J0 panel
    sums of KperpRule.weights, the only expensive part of building a plan."""
    lib = load_library()
    gs = np.ascontiguousarray(group_start, dtype=np.int32)
    r, w, kf, meas = (_f64(x) for x in (r_Mpc, weight, kf, meas))
    out = np.empty((gs.size - 1, kf.size))
    rc = lib.cpx_hankel_jbar(int(device), gs.size - 1, gs.ctypes.data_as(c_ip), _dp(r), _dp(w), kf.size, _dp(kf), _dp(meas), _dp(out))
    if rc != 0:
        raise CpxError("cpx_hankel_jbar failed (%d): %s" % (rc, lib.cpx_last_error().decode()))
    return out


def hankel_apply(device, W, cells):
    """px[r, k] = sum_q W[r, q] cells[k, q] on the GPU (cpx_hankel_apply): the Hankel contraction for P3D cells evaluated by
    the caller (Theory._compute_px_from_p3d)."""
    lib = load_library()
    W, cells = _f64(W), _f64(cells)
    assert W.ndim == 2 and cells.ndim == 2 and W.shape[1] == cells.shape[1], "W [n_r, n_q] and cells [n_k, n_q] expected"
    out = np.empty((W.shape[0], cells.shape[0]))
    rc = lib.cpx_hankel_apply(int(device), W.shape[0], W.shape[1], cells.shape[0], _dp(W), _dp(cells), _dp(out))
    if rc != 0:
        raise CpxError("cpx_hankel_apply failed (%d): %s" % (rc, lib.cpx_last_error().decode()))
    return out


class PxEngine(object):
    """Device-resident Px likelihood for a list of redshift-bin plans (cupix_b200.plan.ZbinPlan)."""

    INFO = {"n_zbins": 0, "max_na": 1, "max_nm": 2, "max_nA": 3, "max_nM": 4, "chunk": 5,
            "device_bytes": 6, "kernel_launches": 7, "sm_count": 8, "guard_buffers": 9,
            "guard_violations": 10, "factored_calls": 11, "factored_tuples": 12, "chunk_allocated": 13,
            "factored_cache_hits": 14, "hankel_kernel": 15}

    def __init__(self, plans, device=0, precise=False):
        """``precise=True`` evaluates the P3D cells in float64 as well (CPX_PRECISE_CELLS): the
        reference-precision mode, ~4x slower, chi2 within ~1e-9 of the float64 oracle everywhere."""
        self.precise = bool(precise)
        self.lib = load_library()
        self.plans = list(plans)
        self.device = int(device)
        descs = (ZbinDesc * len(self.plans))()
        self._keep = []
        for d, p in zip(descs, self.plans):
            arrs = {k: _f64(getattr(p, k)) for k in
                    ("kpar_Mpc", "kperp_Mpc", "hankel_w", "linP_cell", "metal_auto", "metal_cross",
                     "sky_xi_a", "sky_smooth_m", "B_A_a", "U_aMn", "V_aM", "Px_AM", "cov_AMM", "kpar_row_Mpc",
                     "metal_auto_dn", "metal_cross_dn")}
            self._keep.append(arrs)
            d.n_a, d.n_m, d.n_q = p.n_a, p.n_m, p.n_q
            d.n_A, d.n_M, d.n_data = p.n_A, p.n_M, p.n_data
            d.flags = p.flags
            d.dAA_dMpc = p.dAA_dMpc
            d.k_pivot_Mpc = float(getattr(p, "k_pivot_Mpc", 0.0))
            d.metal_dn_order = int(getattr(p, "metal_dn_order", 0))
            for k, v in arrs.items():
                setattr(d, k, _dp(v))
            for i in range(NSLOTS):
                d.defaults[i] = float(p.defaults[i])
        h = C.c_void_p()
        rc = self.lib.cpx_create(descs, len(self.plans), self.device, C.byref(h))
        if rc != 0:
            raise CpxError("cpx_create failed (%d): %s" % (rc, self.lib.cpx_last_error().decode()))
        self._h = h
        self._keep = None           # host arrays were copied to the device
        self.n_z = len(self.plans)

    def close(self):
        if getattr(self, "_h", None):
            self.lib.cpx_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def info(self, what):
        v = C.c_int64()
        if self.lib.cpx_get_info(self._h, self.INFO[what], C.byref(v)) != 0:
            raise CpxError(self.lib.cpx_last_error().decode())
        return v.value

    def set_data(self, iz, Px_AM):
        Px_AM = _f64(Px_AM)
        if Px_AM.ndim == 2:
            Px_AM = Px_AM[None]
        rc = self.lib.cpx_set_data(self._h, iz, _dp(Px_AM), Px_AM.shape[0])
        if rc != 0:
            raise CpxError("cpx_set_data failed (%d): %s" % (rc, self.lib.cpx_last_error().decode()))
        self.plans[iz].n_data = Px_AM.shape[0]

    def generate_mocks(self, n_mocks, seed, point=(), slots=(), zsel=None, return_mocks=True):
        """``n_mocks`` realisations d = model(point) + L n of the data vector of every z bin, drawn on the device
        (cpx_generate_mocks) and installed as data vectors 0 .. n_mocks-1.  Returns them in data space
        [n_z, n_mocks, n_A_max, n_M_max] unless ``return_mocks`` is False."""
        slots = np.ascontiguousarray(slots, dtype=np.int32)
        point = np.ascontiguousarray(point, dtype=np.float64).reshape(-1)
        assert point.size == slots.size
        zs = None if zsel is None else np.ascontiguousarray(zsel, dtype=np.int32)
        out = None
        if return_mocks:
            out = np.zeros((self.n_z, int(n_mocks), max(p.n_A for p in self.plans), max(p.n_M for p in self.plans)))
        rc = self.lib.cpx_generate_mocks(self._h, slots.size, slots.ctypes.data_as(c_ip) if slots.size else None,
                                         None if zs is None else zs.ctypes.data_as(c_ip), _dp(point) if point.size else None,
                                         int(n_mocks), int(seed) & 0xFFFFFFFFFFFFFFFF, _dp(out))
        if rc != 0:
            raise CpxError("cpx_generate_mocks failed (%d): %s" % (rc, self.lib.cpx_last_error().decode()))
        for p in self.plans:
            p.n_data = int(n_mocks)
        return out

    # ----------------------------------------------------------------------------------------
    def _dims(self, z_list):
        zs = list(range(self.n_z)) if z_list is None else list(z_list)
        pl = [self.plans[i] for i in zs]
        return zs, max(p.n_a for p in pl), max(p.n_m for p in pl), max(p.n_A for p in pl), max(p.n_M for p in pl)

    def _flags(self, factor, extra=0):
        """cpx_eval flags.  ``factor``: None = the library decides (chi2-only calls whose points share tuples take the
        factorised path, include/cupix_b200.h), False = every point through the full pipeline (CPX_NO_FACTOR)."""
        return extra | (PRECISE_CELLS if self.precise else 0) | (NO_FACTOR if factor is False else 0)

    def eval(self, points=None, slots=(), zsel=None, grid=None, grid_first=0, B=None, data_idx=None,
             z_list=None, want=("chi2",), stage_ms=False, factor=None, chi2_out=None):
        """Host-buffer evaluation.  ``points`` [B, P] float64 (or ``grid=(lo, hi, n)`` for on-device
        meshgrid generation), ``slots`` [P] parameter slot per column.  Returns a dict of numpy
        arrays for the requested outputs in ``want`` (chi2, chi2_zA, model, px_zam).  ``chi2_out``: caller's
        float64 host array [B] (e.g. pinned memory) to receive chi2 instead of a fresh array."""
        slots = np.ascontiguousarray(slots, dtype=np.int32)
        P = int(slots.size)
        args = EvalArgs()
        keep = [slots]
        if grid is not None:
            lo, hi, n = (np.ascontiguousarray(grid[0], dtype=np.float64), np.ascontiguousarray(grid[1], dtype=np.float64),
                         np.ascontiguousarray(grid[2], dtype=np.int32))
            assert lo.size == hi.size == n.size == P
            total = int(np.prod(n.astype(np.int64)))
            B = total - grid_first if B is None else B
            args.grid_lo, args.grid_hi, args.grid_n = _dp(lo), _dp(hi), n.ctypes.data_as(c_ip)
            args.grid_first = int(grid_first)
            keep += [lo, hi, n]
        else:
            if points is None:
                points = np.zeros((1 if B is None else B, 0))
            points = np.ascontiguousarray(points, dtype=np.float64)
            if points.ndim == 1:
                points = points.reshape(1, -1)
            assert points.shape[1] == P, "points must be [B, len(slots)]"
            B = points.shape[0]
            args.points = points.ctypes.data_as(C.c_void_p) if P > 0 else None
            keep.append(points)
        args.B, args.P, args.flags = int(B), P, self._flags(factor)
        args.slot = slots.ctypes.data_as(c_ip) if P > 0 else None
        if zsel is not None:
            zsel = np.ascontiguousarray(zsel, dtype=np.int32)
            args.zsel = zsel.ctypes.data_as(c_ip)
            keep.append(zsel)
        if data_idx is not None:
            data_idx = np.ascontiguousarray(data_idx, dtype=np.int32)
            assert data_idx.size == B
            args.data_idx = data_idx.ctypes.data_as(C.c_void_p)
            keep.append(data_idx)
        zs, na, nm, nA, nM = self._dims(z_list)
        if z_list is not None:
            zl = np.ascontiguousarray(zs, dtype=np.int32)
            args.z_list, args.n_z_list = zl.ctypes.data_as(c_ip), len(zs)
            keep.append(zl)
        out = {}
        if "chi2" in want:
            if chi2_out is not None:
                assert chi2_out.dtype == np.float64 and chi2_out.flags.c_contiguous and chi2_out.size >= B
                out["chi2"] = chi2_out.reshape(-1)[:B]
            else:
                out["chi2"] = np.empty(B)
            args.chi2 = out["chi2"].ctypes.data_as(C.c_void_p)
        if "chi2_zA" in want:
            out["chi2_zA"] = np.zeros((B, len(zs), nA))
            args.chi2_zA = out["chi2_zA"].ctypes.data_as(C.c_void_p)
        if "model" in want:
            out["model"] = np.zeros((B, len(zs), nA, nM))
            args.model_zAM = out["model"].ctypes.data_as(C.c_void_p)
        if "px_zam" in want:
            out["px_zam"] = np.zeros((B, len(zs), na, nm))
            args.px_zam = out["px_zam"].ctypes.data_as(C.c_void_p)
        ms = (C.c_float * 4)()
        if stage_ms:
            args.stage_ms = ms
        rc = self.lib.cpx_eval(self._h, C.byref(args))
        if rc != 0:
            raise CpxError("cpx_eval failed (%d): %s" % (rc, self.lib.cpx_last_error().decode()))
        if stage_ms:
            out["stage_ms"] = np.array(list(ms))
        return out

    def eval_host_into(self, points, slots, zsel, chi2_out, data_idx=None, factor=None):
        """chi2 for host ``points`` [B, P] written into the caller's host array ``chi2_out`` [B]
        (e.g. pinned memory); H2D, kernels and D2H all happen inside this one cpx_eval call."""
        slots = np.ascontiguousarray(slots, dtype=np.int32)
        assert points.flags.c_contiguous and points.dtype == np.float64 and chi2_out.flags.c_contiguous
        args = EvalArgs()
        args.B, args.P, args.flags = points.shape[0], points.shape[1], self._flags(factor)
        args.points = points.ctypes.data_as(C.c_void_p)
        args.slot = slots.ctypes.data_as(c_ip)
        if zsel is not None:
            zsel = np.ascontiguousarray(zsel, dtype=np.int32)
            args.zsel = zsel.ctypes.data_as(c_ip)
        if data_idx is not None:
            data_idx = np.ascontiguousarray(data_idx, dtype=np.int32)
            args.data_idx = data_idx.ctypes.data_as(C.c_void_p)
        args.chi2 = chi2_out.ctypes.data_as(C.c_void_p)
        rc = self.lib.cpx_eval(self._h, C.byref(args))
        if rc != 0:
            raise CpxError("cpx_eval failed (%d): %s" % (rc, self.lib.cpx_last_error().decode()))
        return chi2_out

    def eval_device(self, B, P, slots, points_ptr=None, grid=None, grid_first=0, zsel=None,
                    data_idx_ptr=None, z_list=None, chi2_ptr=None, chi2_zA_ptr=None, model_ptr=None,
                    px_ptr=None, stream=None, stage_ms=False, factor=None):
        """Device-pointer evaluation (asynchronous on ``stream``): pointers are raw device addresses
        (e.g. ``torch.Tensor.data_ptr()``).  ``stream`` is the raw ``cudaStream_t`` the buffers are ordered on; None or 0
        is the CUDA default (legacy) stream -- what ``torch.cuda.current_stream().cuda_stream`` is for torch's default
        stream.  Returns stage times if requested (this synchronises)."""
        slots = np.ascontiguousarray(slots, dtype=np.int32)
        args = EvalArgs()
        args.B, args.P, args.flags = int(B), int(P), self._flags(factor, DEVICE_IO)
        args.slot = slots.ctypes.data_as(c_ip) if P > 0 else None
        args.points = points_ptr
        keep = [slots]
        if grid is not None:
            lo, hi, n = (np.ascontiguousarray(grid[0], dtype=np.float64), np.ascontiguousarray(grid[1], dtype=np.float64),
                         np.ascontiguousarray(grid[2], dtype=np.int32))
            args.grid_lo, args.grid_hi, args.grid_n = _dp(lo), _dp(hi), n.ctypes.data_as(c_ip)
            args.grid_first = int(grid_first)
            keep += [lo, hi, n]
        if zsel is not None:
            zsel = np.ascontiguousarray(zsel, dtype=np.int32)
            args.zsel = zsel.ctypes.data_as(c_ip)
            keep.append(zsel)
        if z_list is not None:
            zl = np.ascontiguousarray(list(z_list), dtype=np.int32)
            args.z_list, args.n_z_list = zl.ctypes.data_as(c_ip), zl.size
            keep.append(zl)
        args.data_idx = data_idx_ptr
        args.chi2, args.chi2_zA, args.model_zAM, args.px_zam = chi2_ptr, chi2_zA_ptr, model_ptr, px_ptr
        args.stream = stream
        ms = (C.c_float * 4)()
        if stage_ms:
            args.stage_ms = ms
        rc = self.lib.cpx_eval(self._h, C.byref(args))
        if rc != 0:
            raise CpxError("cpx_eval failed (%d): %s" % (rc, self.lib.cpx_last_error().decode()))
        return np.array(list(ms)) if stage_ms else None


def slots_from_names(names, n_z=None):
    """Parameter names -> (slots, zsel).  A trailing ``_<iz>`` restricts a parameter to one z bin
    (the convention of the reference scan scripts, scripts/chi2_scan_mock.py:138).  Unknown names
    raise here; the dict-based scalar API drops them silently like the reference does."""
    slots, zsel = [], []
    for name in names:
        base, iz = name, -1
        if name not in _plan.SLOT_INDEX:
            head, _, tail = name.rpartition("_")
            if head in _plan.SLOT_INDEX and tail.isdigit():
                base, iz = head, int(tail)
            else:
                raise KeyError("unknown parameter name %r" % name)
        slots.append(_plan.SLOT_INDEX[base])
        zsel.append(iz)
    return np.array(slots, dtype=np.int32), np.array(zsel, dtype=np.int32)
