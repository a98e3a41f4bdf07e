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


def test_tensor_core_schedule_is_injective_and_pruned():
    """Host logic of the tcgen05 contraction (no device): every triangle of a dense list reads
    its own accumulator slot (pair row x column), the S=40 all-triangle list needs 7 units whose
    column counts sum to 184 (vs 7 x 40 dense), and lists that do not fit are reported ineligible."""
    import ctypes as C
    import numpy as np
    from bskit_b200 import _native
    from bskit_b200.bins import generate_triangle_bin_list
    lib = _native.lib()

    def info(rows, nrows):
        rows = np.ascontiguousarray(rows, dtype=np.int32)
        out = (C.c_int64 * 6)()
        assert lib.bsk_tc_schedule_info(len(rows), rows.ctypes.data_as(C.POINTER(C.c_int32)), nrows, out) == 0
        return list(out)

    kf = 2 * np.pi / 1000.0
    for nb, units, cols in ((40, 7, 184), (24, 5, None), (12, 3, None)):
        idx = generate_triangle_bin_list(kmin=0.5 * kf, kmax=(nb + 1.0) * kf, dk=kf, return_indices=True)
        assert idx.max() == nb - 1
        got = info(idx, (nb + 3) // 4 * 4)
        assert got[0] == units, got
        assert got[1] == len(idx) and got[4] == 1          # injective, in range
        if cols:
            assert got[2] == cols
        assert got[3] == nb * (nb + 1) // 2 and got[3] <= 128 * units     # one pair row per row pair
    # random dense list over 40 rows in arbitrary row order: still injective
    rng = np.random.default_rng(0)
    tri = np.array([[a, b, c] for a in range(40) for b in range(a, 40) for c in range(b, 40) if (a + b + c) % 3 == 0])
    tri = np.array([rng.permutation(t) for t in tri])
    got = info(tri, 40)
    assert got[0] > 0 and got[1] == len(tri) and got[4] == 1
    # not eligible: too many rows, too few triangles
    assert info(np.array([[a, a, a] for a in range(80)] * 4), 80)[0] == 0
    assert info(np.array([[0, 1, 2]]), 4)[0] == 0
