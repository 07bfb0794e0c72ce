"""CPU: the C-ABI library loads and exports every symbol include/cupix_b200.h declares."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_declared_symbols():
    import __graft_entry__ as g
    lib_path = g.build()
    lib = ctypes.CDLL(lib_path)
    header = open(os.path.join(ROOT, "include", "cupix_b200.h")).read()
    declared = set(re.findall(r"^\s*(?:int|void|const char\*)\s+(cpx_\w+)\s*\(", header, flags=re.M))
    assert {"cpx_create", "cpx_eval", "cpx_set_data", "cpx_destroy", "cpx_last_error", "cpx_version",
            "cpx_get_info", "cpx_probe_peaks"} <= declared
    for name in declared:
        assert hasattr(lib, name), "symbol %s missing from libcupix_b200.so" % name
    assert lib.cpx_version() >= 100


def test_struct_layout_matches_header():
    """ctypes mirrors of cpx_zbin_desc / cpx_eval_args have the sizes the C compiler gives them."""
    import subprocess
    import tempfile
    from cupix_b200 import engine
    src = '#include "cupix_b200.h"\n#include <stdio.h>\nint main(){printf("%zu %zu\\n", sizeof(cpx_zbin_desc), sizeof(cpx_eval_args));return 0;}\n'
    with tempfile.TemporaryDirectory() as td:
        open(os.path.join(td, "s.c"), "w").write(src)
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), "-o", os.path.join(td, "s"), os.path.join(td, "s.c")])
        a, b = map(int, subprocess.check_output([os.path.join(td, "s")]).split())
    assert a == ctypes.sizeof(engine.ZbinDesc)
    assert b == ctypes.sizeof(engine.EvalArgs)


def test_no_cpu_fallback():
    """Without a CUDA device the product path fails loudly instead of computing on the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA device present")
    from cupix_b200.engine import CpxError
    from tests import helpers as H
    import numpy as np
    th = H.make_theory(2.2, "best_fit_arinyo_from_colore", {})
    with pytest.raises(CpxError, match="no CUDA device|no CPU fallback"):
        th.get_px_obs(np.array([1.0, 2.0]), np.array([0.1, 0.2]))


def test_device_helpers_fail_loudly_without_a_gpu():
    """cpx_hankel_apply / cpx_hankel_jbar / cpx_probe_imma and Theory._compute_px_from_p3d have no CPU path either."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA device present")
    import numpy as np
    from cupix_b200 import engine
    from tests import helpers as H
    with pytest.raises(engine.CpxError, match="cpx_hankel_apply failed"):
        engine.hankel_apply(0, np.ones((3, 8)), np.ones((5, 8)))
    with pytest.raises(engine.CpxError, match="cpx_probe_imma failed"):
        engine.probe_imma(0)
    with pytest.raises(engine.CpxError, match="cpx_hankel_jbar failed"):
        engine.hankel_jbar(0, [0, 1], [1.0], [1.0], [0.1, 0.2], [1.0, 1.0])
    th = H.make_theory(2.2, "best_fit_arinyo_from_colore", {})
    with pytest.raises(engine.CpxError, match="needs a CUDA device"):
        th._compute_px_from_p3d(np.array([1.0]), np.array([0.1]), lambda k, mu: np.exp(-k * k), 200.0)
    with pytest.raises(AssertionError):
        engine.hankel_apply(0, np.ones((3, 8)), np.ones((5, 7)))


def test_product_does_not_import_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "cupix_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M), f


def test_slot_enum_info_enum_and_field_offsets_match_header():
    """The Python side's slot order (plan.SLOT_NAMES), info items (PxEngine.INFO) and ctypes field offsets are the
    header's: compiled from include/cupix_b200.h with gcc and compared value by value."""
    import subprocess
    import tempfile
    from cupix_b200 import engine, plan
    enum_of = {"bias": "CPX_BIAS", "beta": "CPX_BETA", "q1": "CPX_Q1", "q2": "CPX_Q2", "kv_Mpc": "CPX_KV_MPC",
               "av": "CPX_AV", "bv": "CPX_BV", "kp_Mpc": "CPX_KP_MPC", "b_H": "CPX_B_H", "beta_H": "CPX_BETA_H",
               "L_H_Mpc": "CPX_L_H_MPC", "b_X": "CPX_B_X", "beta_X": "CPX_BETA_X", "b_noise_Mpc": "CPX_B_NOISE_MPC",
               "kC_Mpc": "CPX_KC_MPC", "pC": "CPX_PC", "As": "CPX_AS", "ns": "CPX_NS"}
    assert set(enum_of) == set(plan.SLOT_NAMES)
    info_of = {"n_zbins": "CPX_INFO_N_ZBINS", "max_na": "CPX_INFO_MAX_NA", "max_nm": "CPX_INFO_MAX_NM",
               "max_nA": "CPX_INFO_MAX_N_A", "max_nM": "CPX_INFO_MAX_N_M", "chunk": "CPX_INFO_CHUNK",
               "device_bytes": "CPX_INFO_DEVICE_BYTES", "kernel_launches": "CPX_INFO_KERNEL_LAUNCHES",
               "sm_count": "CPX_INFO_SM_COUNT", "guard_buffers": "CPX_INFO_GUARD_BUFFERS",
               "guard_violations": "CPX_INFO_GUARD_VIOLATIONS", "factored_calls": "CPX_INFO_FACTORED_CALLS",
               "factored_tuples": "CPX_INFO_FACTORED_TUPLES", "chunk_allocated": "CPX_INFO_CHUNK_ALLOCATED",
               "factored_cache_hits": "CPX_INFO_FACTORED_CACHE_HITS", "hankel_kernel": "CPX_INFO_HANKEL_KERNEL"}
    assert set(info_of) == set(engine.PxEngine.INFO)
    zfields = [f[0] for f in engine.ZbinDesc._fields_]
    efields = [f[0] for f in engine.EvalArgs._fields_]
    lines = ['#include "cupix_b200.h"', "#include <stdio.h>", "#include <stddef.h>", "int main(){"]
    lines += ['printf("slot %s %%d\\n", (int)%s);' % (n, e) for n, e in enum_of.items()]
    lines += ['printf("info %s %%d\\n", (int)%s);' % (n, e) for n, e in info_of.items()]
    lines += ['printf("z %s %%zu\\n", offsetof(cpx_zbin_desc, %s));' % (f, f) for f in zfields]
    lines += ['printf("e %s %%zu\\n", offsetof(cpx_eval_args, %s));' % (f, f) for f in efields]
    lines += ['printf("nslots x %d\\n", (int)CPX_NSLOTS);', 'printf("flag DEVICE_IO %d\\n", (int)CPX_DEVICE_IO);',
              'printf("flag PRECISE_CELLS %d\\n", (int)CPX_PRECISE_CELLS);',
              'printf("flag NO_FACTOR %d\\n", (int)CPX_NO_FACTOR);', "return 0;}"]
    with tempfile.TemporaryDirectory() as td:
        open(os.path.join(td, "s.c"), "w").write("\n".join(lines) + "\n")
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), "-o", os.path.join(td, "s"), os.path.join(td, "s.c")])
        out = [ln.split() for ln in subprocess.check_output([os.path.join(td, "s")]).decode().splitlines()]
    for kind, name, val in out:
        val = int(val)
        if kind == "slot":
            assert plan.SLOT_INDEX[name] == val, name
        elif kind == "info":
            assert engine.PxEngine.INFO[name] == val, name
        elif kind == "z":
            assert getattr(engine.ZbinDesc, name).offset == val, name
        elif kind == "e":
            assert getattr(engine.EvalArgs, name).offset == val, name
        elif kind == "nslots":
            assert engine.NSLOTS == val == len(plan.SLOT_NAMES)
        elif kind == "flag":
            assert getattr(engine, name) == val


def test_integration_stub_is_current():
    """The ctypes stub printed in INTEGRATION.md (what a cupix maintainer would paste) still matches the library:
    it is executed against the in-tree .so and its structs / slot table are compared with the engine's."""
    from cupix_b200 import engine, plan
    text = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    start = text.index("# cupix/likelihood/_b200.py")
    stub = text[start:text.index("```", start)]
    stub = stub.replace('"libcupix_b200.so"', repr(engine.LIB_PATH))
    ns = {}
    exec(compile(stub, "INTEGRATION.md", "exec"), ns)
    assert ctypes.sizeof(ns["ZbinDesc"]) == ctypes.sizeof(engine.ZbinDesc)
    assert ctypes.sizeof(ns["EvalArgs"]) == ctypes.sizeof(engine.EvalArgs)
    for name, _ in engine.ZbinDesc._fields_:
        assert getattr(ns["ZbinDesc"], name).offset == getattr(engine.ZbinDesc, name).offset, name
    for name, _ in engine.EvalArgs._fields_:
        assert getattr(ns["EvalArgs"], name).offset == getattr(engine.EvalArgs, name).offset, name
    assert ns["SLOT"] == plan.SLOT_INDEX
    assert callable(ns["chi2_batch"])
