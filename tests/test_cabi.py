"""The C-ABI library loads on a CPU-only box and exports every symbol the header declares
(no compute calls here: those need a GPU and live in the -m gpu tests)."""
import os
import re
import subprocess

from conftest import ROOT


def _header_symbols():
    text = open(os.path.join(ROOT, "include", "bskit_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(bsk_[a-z0-9_]+)\s*\(", text)))


def test_library_builds_and_exports_header_symbols():
    import __graft_entry__ as g
    g.build()
    from bskit_b200 import _native
    lib = _native.lib()
    declared = _header_symbols()
    assert declared, "no declarations parsed from the header"
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/bskit_b200.h but not exported"
    assert sorted(_native.EXPORTS) == declared
    assert lib.bsk_version() >= 100


def test_library_is_built_for_sm100a_with_tma_bulk_copies():
    from bskit_b200 import _native
    out = subprocess.run(["cuobjdump", "-lelf", _native.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out
    sass = subprocess.run(["cuobjdump", "-sass", _native.LIB_PATH], capture_output=True, text=True).stdout
    assert "tile_contract_kernel" in sass
    assert "UBLKCP" in sass               # cp.async.bulk (TMA 1-D bulk copy) feeds the tile ring
    assert "SYNCS.ARRIVE.TRANS64" in sass  # mbarrier expect_tx
    # the tensor-core contraction path: tcgen05.mma with the A operand in TMEM, tcgen05.st/ld
    assert "tc_contract_kernel" in sass
    assert "UTCHMMA" in sass and "STTM" in sass and "LDTM" in sass


def test_argument_errors_are_reported_not_crashes():
    import ctypes as C
    from bskit_b200 import _native
    lib = _native.lib()
    rc = lib.bsk_plan_create(None, None, None, None, None, None)
    assert rc == -1 and b"null" in lib.bsk_last_error()
    handle = C.c_void_p()
    geom = _native.Geometry(7, 7, 1, 0, 1, 0, 1, 0)      # odd mesh
    import numpy as np
    t = np.zeros(8)
    rc = lib.bsk_plan_create(C.byref(handle), C.byref(geom), _native.dptr(t), _native.dptr(t),
                             _native.dptr(t), None)
    assert rc == -1 and b"nmesh" in lib.bsk_last_error()
    # entry points added for the tensor-core path and for particle painting validate their arguments
    # before any CUDA call
    assert lib.bsk_cplan_set_path(None, 1) == -1 and b"bsk_cplan_set_path" in lib.bsk_last_error()
    assert lib.bsk_cplan_path(None, None) == -1
    box = np.array([100.0, 100.0, 100.0])
    assert lib.bsk_paint_cic(None, 0, 10, 16, _native.dptr(box), None, None) == -1
    assert b"bsk_paint_cic" in lib.bsk_last_error()


def test_no_cpu_fallback():
    import torch
    import pytest
    from bskit_b200 import engine as eng, _native
    g = eng.choose_grid(16, 100.0, 0.2, "full")
    with pytest.raises(_native.NativeError):
        eng.Engine(g, 100.0, _native.F32, device=torch.device("cpu"))
