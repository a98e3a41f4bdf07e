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
    its own accumulator slot (pair row x column); the class cover needs fewer 128-row units than
    one pair row per row pair would (S=40: 5 units instead of 7); lists that do not fit one launch
    (S=80, cross lists) are cut into passes; short lists are reported ineligible."""
    import ctypes as C
    import numpy as np
    from bskit_b200 import _native
    from bskit_b200.bins import generate_triangle_bin_list
    lib = _native.lib()

    def info(rows, nrows):
        rows = np.ascontiguousarray(rows, dtype=np.int32)
        out = (C.c_int64 * 6)()
        assert lib.bsk_tc_schedule_info(len(rows), rows.ctypes.data_as(C.POINTER(C.c_int32)), nrows, out) == 0
        return dict(units=out[0], distinct=out[1], cols=out[2], passes=out[3], in_range=out[4], cost=out[5])

    kf = 2 * np.pi / 1000.0

    def all_list(nb, nf=1):
        idx = generate_triangle_bin_list(kmin=0.5 * kf, kmax=(nb + 1.0) * kf, dk=kf, num_fields=nf,
                                         return_indices=True)
        seg = (nb + 3) // 4 * 4
        return np.asarray(idx) + np.array([0, 0, seg if nf == 2 else 0]), seg * nf

    # S = 40: the two-halves schedule -- 2 + 2 units (420 pair rows on 512 lanes), 152 accumulator columns, one pass
    for nb, max_units, max_passes in ((40, 4, 1), (24, 3, 1), (12, 2, 1), (80, 32, 6)):
        idx, nrows = all_list(nb)
        got = info(idx, nrows)
        assert 0 < got["units"] <= max_units and got["passes"] <= max_passes, got
        assert got["distinct"] == len(idx) and got["in_range"] == 1, got      # injective, in range
        assert got["cols"] <= 96 * 2 * got["passes"], got                     # TMEM accumulator budget
        if nb == 40:
            assert got["cols"] <= 160 and got["cost"] <= 80, got
    # two-field <AAB> list: 80 rows, the third row lives in the second segment
    idx, nrows = all_list(40, 2)
    got = info(idx, nrows)
    assert got["units"] > 0 and got["distinct"] == len(idx) and got["in_range"] == 1, got
    # random dense list over 40 rows in arbitrary row order: still injective
    rng = np.random.default_rng(0)
    tri = np.array([[a, b, c] for a in range(40) for b in range(a, 40) for c in range(b, 40) if (a + b + c) % 3 == 0])
    tri = np.array([rng.permutation(t) for t in tri])
    got = info(tri, 40)
    assert got["units"] > 0 and got["distinct"] == len(tri) and got["in_range"] == 1
    # not eligible: too few triangles
    assert info(np.array([[a, a, a] for a in range(80)]), 80)["units"] == 0
    assert info(np.array([[0, 1, 2]]), 4)["units"] == 0


def test_tensor_core_schedule_semantics_on_the_cpu():
    """bsk_tc_schedule_eval follows the schedule tables the way tc_contract_kernel routes data (lane -> pair
    rows, window column -> field row, (team, column, lane) -> slot, triangle -> slot) in float64 on the host:
    every builder (two halves, class cover, multi-pass, cross lists, arbitrary row order) must deliver
    sum_x f[r1] f[r2] f[r3] for every triangle.  Shapes the GPU tests do not run are covered here."""
    import ctypes as C
    import numpy as np
    from bskit_b200 import _native
    from bskit_b200.bins import generate_triangle_bin_list
    lib = _native.lib()
    kf = 2 * np.pi / 1000.0
    rng = np.random.default_rng(5)

    def check(tri, nrows, ncells=3):
        tri = np.ascontiguousarray(tri, dtype=np.int32)
        f = rng.standard_normal((nrows, ncells))
        out = np.zeros(len(tri))
        rc = lib.bsk_tc_schedule_eval(len(tri), tri.ctypes.data_as(C.POINTER(C.c_int32)), nrows, ncells,
                                      f.ctypes.data_as(C.POINTER(C.c_double)), out.ctypes.data_as(C.POINTER(C.c_double)))
        assert rc == 0, _native.lib().bsk_last_error()
        want = (f[tri[:, 0]] * f[tri[:, 1]] * f[tri[:, 2]]).sum(axis=1)
        np.testing.assert_allclose(out, want, rtol=1e-12, atol=1e-12)

    for nb in (12, 16, 20, 24, 28, 32, 36, 40, 48, 80):          # 1+1, 1+2, 2+1, 2+2 units, several passes
        idx = np.asarray(generate_triangle_bin_list(kmin=0.5 * kf, kmax=(nb + 1.0) * kf, dk=kf, return_indices=True))
        check(idx, (nb + 3) // 4 * 4)
    # <AAB>: third row in a second 40-row segment
    idx = np.asarray(generate_triangle_bin_list(kmin=0.5 * kf, kmax=41.0 * kf, dk=kf, num_fields=2, return_indices=True))
    check(idx + np.array([0, 0, 40]), 80)
    # the reference's production binning (80 bins, 24138 triangles)
    kf75 = 2 * np.pi / 75.0
    idx = np.asarray(generate_triangle_bin_list(0.5 * kf75, 23.5, kf75, num_lowk_bins=40, dk_high=6 * kf75, return_indices=True))
    check(idx, 80, ncells=2)
    # arbitrary row order and a non-closure selection rule
    tri = np.array([[a, b, c] for a in range(40) for b in range(a, 40) for c in range(b, 40) if (a + b + c) % 3 == 0])
    check(np.array([rng.permutation(t) for t in tri]), 40)
