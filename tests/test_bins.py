"""Host-side bin / triangle generators: bit-identical to the oracle's restatement of
main.py:1008-1389 and to the reference golden rows; error behaviour of SURVEY.md 8b."""
import os

import numpy as np
import pytest

import bskit_b200 as bk
from oracle import bskit_oracle as orc
from conftest import REF_OUT, fmt_e

KF = 2 * np.pi / 1000.0
SCHEMES = [
    dict(kmin=0.00314, kmax=0.1, dk=0.00628, num_lowk_bins=3, dk_high=0.01884),
    dict(kmin=0.5 * KF, kmax=0.5 * KF + 12.5 * KF, dk=KF),
    dict(kmin=0.5 * KF, kmax=0.5 * KF + 20.5 * KF, dk=KF, num_lowk_bins=5, dk_high=3 * KF),
    dict(kmin=0.02, kmax=0.31, dk=0.03),
]


@pytest.mark.parametrize("scheme", SCHEMES)
def test_generators_bit_identical_to_oracle(scheme):
    e = bk.generate_bin_edge_list(**scheme)
    eo = orc.bin_edges(**scheme)
    assert np.array_equal(e, eo)
    for nf in (1, 2, 3):
        E = bk.generate_triangle_bin_list(num_fields=nf, **scheme)
        I = bk.generate_triangle_bin_list(num_fields=nf, return_indices=True, **scheme)
        Eo, Io = orc.triangles_all(eo, nf)
        assert np.array_equal(E, Eo) and np.array_equal(I, Io)
    assert np.array_equal(bk.generate_equilateral_triangle_bin_list(**scheme), orc.triangles_equilateral(eo)[0])
    for q in (0, 2):
        Eo, Io = orc.triangles_squeezed(eo, q)
        assert np.array_equal(bk.generate_squeezed_triangle_bin_list(squeezed_bin_index=q, **scheme), Eo)
        assert np.array_equal(bk.generate_squeezed_triangle_bin_list(squeezed_bin_index=q, return_indices=True, **scheme), Io)
    for m in (1.5, 2, 3):
        Eo, Io = orc.triangles_isosceles(eo, m)
        assert np.array_equal(bk.generate_isosceles_triangle_bin_list(isos_mult=m, **scheme), Eo)
        assert np.array_equal(bk.generate_isosceles_triangle_bin_list(isos_mult=m, return_indices=True, **scheme), Io)


def test_golden_rows_edges_and_order():
    g = np.loadtxt(os.path.join(REF_OUT, "Lbox1000_512_kf_3kf_3lowkbins.dat"))
    E = bk.generate_triangle_bin_list(**SCHEMES[0])
    assert len(E) == 59
    for t in range(59):
        assert [fmt_e(v) for v in E[t]] == [fmt_e(v) for v in g[t, 4:10]]


def test_survey_sizes():
    # S=40 -> 6730 (1 field), 18911 (2 fields); S=80 -> 48260
    kmin, kmax, dk = 0.5 * KF, 0.5 * KF + 40.5 * KF, KF
    assert len(bk.generate_bin_edge_list(kmin, kmax, dk)) == 40
    assert len(bk.generate_triangle_bin_list(kmin, kmax, dk)) == 6730
    assert len(bk.generate_triangle_bin_list(kmin, kmax, dk, num_fields=2)) == 18911
    assert len(bk.generate_triangle_bin_list(kmin, 0.5 * KF + 80.5 * KF, dk, return_indices=True)) == 48260
    iso = bk.generate_isosceles_triangle_bin_list(kmin, kmax, dk, isos_mult=2, return_indices=True)
    assert len(iso) == 34                                 # T = 34 for S = 40, m = 2


def test_error_behaviour():
    with pytest.raises(ValueError):
        bk.generate_bin_edge_list(-1, 1, 0.1)
    with pytest.raises(ValueError):
        bk.generate_bin_edge_list(0.1, -1, 0.1)
    with pytest.raises(ValueError):
        bk.generate_bin_edge_list(0.1, 1, 0)
    with pytest.raises(ValueError):
        bk.generate_bin_edge_list(0.1, 1, 0.1, num_lowk_bins=2)          # dk_high missing
    with pytest.raises(ValueError):
        bk.generate_isosceles_triangle_bin_list(0.1, 1, 0.1, isos_mult=0.5)
    with pytest.raises(ValueError):
        bk.generate_triangle_bin_list(0.1, 1, 0.1, num_fields=4)
    with pytest.raises(NotImplementedError):
        bk.generate_triangle_bin_list(0.1, 1, 0.1, dmu=0.1)


def test_fftbispectrum_ctor_errors_without_gpu():
    mesh = np.zeros((8, 8, 8), dtype=np.float32)
    with pytest.raises(ValueError):
        bk.FFTBispectrum(mesh, BoxSize=100.0, for_grid_info_only=True)            # no kmin/kmax/k_edges
    with pytest.raises(ValueError):
        bk.FFTBispectrum(mesh, BoxSize=100.0, kmin=0.1, kmax=1, dk=0.1, third=mesh, for_grid_info_only=True)
    with pytest.raises(ValueError):
        bk.FFTBispectrum(mesh, BoxSize=100.0, kmin=0.1, kmax=1, dk=0.1, second=np.zeros((4, 4, 4)),
                         for_grid_info_only=True)                                  # Nmesh mismatch
    with pytest.raises(ValueError):
        bk.FFTBispectrum(mesh, kmin=0.1, kmax=1, dk=0.1, for_grid_info_only=True)  # bare array, no BoxSize
    fb = bk.FFTBispectrum(mesh, BoxSize=100.0, kmin=0.1, kmax=1, for_grid_info_only=True)
    assert fb.attrs["dk"] == 2 * np.pi / 100.0 and fb.attrs["painted"] is False
    assert fb.k_edges.shape[1] == 6 and fb.k_indices.shape[1] == 3 and fb.num_fields == 1
    with pytest.raises(NotImplementedError):
        fb.measure_bispectrum(0, 1, kmeas_min=0.1, kmeas_max=0.2)


def test_combine_matches_reference_golden():
    gi = np.loadtxt(os.path.join(REF_OUT, "Lbox1000_512_kf_3kf_3lowkbins.dat"))
    un = np.loadtxt(os.path.join(REF_OUT, "test_grid_512_1_unnormbs_kf_3kf_3lowbins.dat"))
    want = np.loadtxt(os.path.join(REF_OUT, "test_grid_512_1_bs_comb_kf_3kf_3lowbins.dat"))
    got = bk.combine_gridinfo_and_unnormalized(gi, un, k_max=1.0)
    assert got.shape == want.shape
    for r in range(len(want)):
        assert ["%.6e" % v for v in got[r, 1:]] == ["%.6e" % v for v in want[r, 1:]]


def test_reference_production_binning_80_bins_24138_triangles():
    """The reference's production job (examples/batch/sub_measure_bs_faster_ill.sh:33-37): 40 bins of width k_f
    from k_f/2, then bins of width 6 k_f up to 280.51 k_f -> 80 bins, 24 138 closed triangles
    (`bench.py --scheme paper80`); generator == oracle, and the tensor-core schedule covers the list in
    passes over its 80 rows with every triangle in its own accumulator slot."""
    import ctypes as C
    from bskit_b200 import _native as nat
    kf = 2 * np.pi / 75.0                       # LBOX = 75 in the batch script: DK = 0.0837758 = k_f
    kmin, kmax, dk, dk_high = 0.5 * kf, 23.5, kf, 6.0 * kf
    edges = bk.generate_bin_edge_list(kmin, kmax, dk, 40, dk_high)
    idx = bk.generate_triangle_bin_list(kmin, kmax, dk, num_lowk_bins=40, dk_high=dk_high, return_indices=True)
    assert len(edges) == 80 and len(idx) == 24138
    oe = orc.bin_edges(kmin, kmax, dk, 40, dk_high)
    _, oi = orc.triangles_all(oe, 1)
    assert np.array_equal(edges, oe) and np.array_equal(np.asarray(idx), oi)
    rows = np.ascontiguousarray(idx, dtype=np.int32)
    out = (C.c_int64 * 6)()
    assert nat.lib().bsk_tc_schedule_info(len(rows), rows.ctypes.data_as(C.POINTER(C.c_int32)), 80, out) == 0
    assert out[0] > 0 and out[1] == len(rows) and out[4] == 1, list(out)      # eligible, injective, in range
