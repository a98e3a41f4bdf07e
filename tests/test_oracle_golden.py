"""Pin the CPU oracle against the reference's own golden outputs
(/root/reference/examples/tests/output_ref, copied to tests/golden/reference_output_ref)."""
import os

import numpy as np
import pytest

from oracle import bskit_oracle as orc
from conftest import REF_OUT, fmt_e

# examples/tests/scripts/test_compute_bs_gridinfo.sh:17-26
BINS = dict(kmin=0.00314, kmax=0.1, dk=0.00628, num_lowk_bins=3, dk_high=0.01884)


def _load(name):
    return np.loadtxt(os.path.join(REF_OUT, name))


def test_bin_edges_golden_scheme():
    e = orc.bin_edges(**BINS)
    assert e.shape == (7, 2)                       # 7 bins, last upper edge .09734
    assert fmt_e(e[-1, 1]) == "9.734000e-02"
    assert fmt_e(e[0, 0]) == "3.140000e-03"


def test_triangle_list_matches_golden_rows():
    g = _load("Lbox1000_512_kf_3kf_3lowkbins.dat")
    edges6, idx = orc.triangles_all(orc.bin_edges(**BINS), 1)
    assert len(edges6) == len(g) == 59
    for t in range(59):
        assert int(g[t, 0]) == t
        for c in range(6):
            assert fmt_e(edges6[t, c]) == fmt_e(g[t, 4 + c])
    e = orc.bin_edges(**BINS)
    assert np.array_equal(e[idx[:, 0], 0], edges6[:, 0])
    assert np.all(idx[:, 0] >= idx[:, 1]) and np.all(idx[:, 1] >= idx[:, 2])


@pytest.mark.parametrize("nmesh", [64])
def test_gridinfo_matches_golden_all_digits(nmesh):
    """N_tri and k_mean are grid-size independent while 3*n_max < N,
    so a 64^3 oracle run must reproduce the 512^3 golden to every printed digit."""
    g = _load("Lbox1000_512_kf_3kf_3lowkbins.dat")
    e = orc.bin_edges(**BINS)
    _, idx = orc.triangles_all(e, 1)
    ntri, kmean = orc.measure_gridinfo(nmesh, 1000.0, e, idx, workers=4)
    bad = []
    for t in range(59):
        got = [fmt_e(kmean[t, 0]), fmt_e(kmean[t, 1]), fmt_e(kmean[t, 2]), fmt_e(ntri[t])]
        want = [fmt_e(g[t, 1]), fmt_e(g[t, 2]), fmt_e(g[t, 3]), fmt_e(g[t, 10])]
        if got != want:
            bad.append((t, got, want))
    assert not bad, bad


def test_slow_golden_identical_to_fast_golden():
    a = open(os.path.join(REF_OUT, "Lbox1000_512_kf_3kf_3lowkbins.dat")).read()
    b = open(os.path.join(REF_OUT, "Lbox1000_512_kf_3kf_3lowkbins_slow.dat")).read()
    assert a == b


def test_gridinfo_subbox_golden_nonempty_rows():
    """First 59 rows of the sub-box golden: L=500, N=256.  Rows whose
    bin is empty (every triangle touching bin 0, which lies below k_f = 2pi/500) hold
    nan / fp noise in the golden itself; there the oracle must give N_tri ~ 0."""
    g = _load("Lbox1000_512_kf_3kf_3lowkbins_subbox0.dat")[:59]
    e = orc.bin_edges(**BINS)
    _, idx = orc.triangles_all(e, 1)
    ntri, kmean = orc.measure_gridinfo(64, 500.0, e, idx, workers=4)
    checked = 0
    for t in range(59):
        if not np.isfinite(g[t, 1:4]).all() or abs(g[t, 10]) < 0.5:
            assert abs(ntri[t]) < 1e-6
            continue
        checked += 1
        assert fmt_e(ntri[t]) == fmt_e(g[t, 10])
        for c in range(3):
            assert fmt_e(kmean[t, c]) == fmt_e(g[t, 1 + c])
    assert checked == int((np.abs(g[:, 10]) > 0.5).sum()) and checked >= 40


def test_equilateral_list_matches_cross_golden():
    g = _load("test_grid_512_1_cross_test_grid_512_2_bs_eq_kf_3kf_3lowbins_slow.dat")
    e = orc.bin_edges(**BINS)
    edges6, idx = orc.triangles_equilateral(e)
    assert len(g) == len(edges6) == 7
    for t in range(7):
        for c in range(6):
            assert fmt_e(edges6[t, c]) == fmt_e(g[t, 4 + c])
    # N_tri column of the equilateral golden equals the all-triangle golden rows (i,i,i)
    ntri, kmean = orc.measure_gridinfo(64, 1000.0, e, idx, workers=4)
    for t in range(7):
        assert fmt_e(ntri[t]) == fmt_e(g[t, 11])
        assert fmt_e(kmean[t, 0]) == fmt_e(g[t, 1])


def test_unnorm_golden_layout_and_combine_arithmetic():
    """B values depend on pmesh's RNG (parity unpinned); layout, ordering and the
    combine step B = B_unnorm / N_tri (process_fast_bs_measurement.py:73-78) are pinned."""
    u = _load("test_grid_512_1_unnormbs_kf_3kf_3lowbins.dat")
    gi = _load("Lbox1000_512_kf_3kf_3lowkbins.dat")
    comb = _load("test_grid_512_1_bs_comb_kf_3kf_3lowbins.dat")
    assert u.shape == (59, 8) and gi.shape == (59, 11) and comb.shape == (59, 12)
    assert np.array_equal(u[:, 0], gi[:, 0])
    np.testing.assert_allclose(u[:, 1:7], gi[:, 4:10], rtol=0, atol=0)
    np.testing.assert_allclose(comb[:, 10], u[:, 7] / gi[:, 10], rtol=2e-6)


def test_mode_counts_first_bin():
    # shell n^2 in {1,2} holds 6 + 12 = 18 modes
    kf = 2 * np.pi / 1000.0
    assert orc.modes_per_bin(32, 1000.0, [[0.5 * kf, 1.5 * kf]])[0] == 18
