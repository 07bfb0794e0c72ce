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


def test_product_does_not_import_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "cupix_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M), f
