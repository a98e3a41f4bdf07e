"""Host logic of the bskit.main mirror on a CPU-only box: result dict, index slicing, units,
file formats, accumulation across calls, error paths.  The numeric stages are the numpy test
double (tests/fake_backend.py), patched over engine.NativeBackend by the fixture below; the
product itself has no such switch and never runs without the CUDA library."""
import os
import pickle

import numpy as np
import pytest
import torch

import bskit_b200 as bk
from bskit_b200 import engine as bkengine
from bskit_b200 import main as bkmain
from fake_backend import FakeBackend
from oracle import bskit_oracle as orc
from conftest import REF_OUT, fmt_e

BINS = dict(kmin=0.00314, kmax=0.1, dk=0.00628, num_lowk_bins=3, dk_high=0.01884)


@pytest.fixture(autouse=True)
def cpu_backend(monkeypatch):
    monkeypatch.setattr(bkengine, "NativeBackend", FakeBackend)
    yield
    bk.clear_cache()


def _mesh(n, seed=3):
    g = np.random.default_rng(seed).standard_normal((n, n, n))
    return g + 0.3 * g ** 2


def test_gridinfo_golden_file_through_the_api(tmp_path):
    """Same file as examples/tests/output_ref/Lbox1000_512_kf_3kf_3lowkbins.dat, via the public
    API (fast and slow entry points), N=48 (results are N independent while 3 n_max < N)."""
    want = open(os.path.join(REF_OUT, "Lbox1000_512_kf_3kf_3lowkbins.dat")).read().strip().splitlines()
    mesh = np.zeros((48, 48, 48))
    fb = bk.FFTBispectrum(mesh, BoxSize=np.ones(3) * 1000.0, for_grid_info_only=True,
                          device=torch.device("cpu"), **BINS)
    out = tmp_path / "gi.dat"
    r = fb.measure_gridinfo_faster(imin=0, imax=200, out_file=str(out))
    assert out.read_text().strip().splitlines() == want
    assert r["N_tri"].shape == (59,) and r["k_mean"].shape == (59, 3) and r["k_edge"].shape == (59, 6)
    out2 = tmp_path / "gi_slow.dat"
    fb.measure_bispectrum(0, 200, out_file=str(out2), meas_type="grid_info")
    assert out2.read_text().strip().splitlines() == want
    assert len(fb.b["index"]) == 118                       # results accumulate across calls


def test_fast_path_values_slices_units_and_files(tmp_path):
    n, box, u = 16, 200.0, 2.0
    kf = 2 * np.pi / box
    mesh = _mesh(n)
    fb = bk.FFTBispectrum(mesh, BoxSize=box, kmin=0.5 * kf, kmax=5.0 * kf, dk=kf,
                          pos_units_mpcoverh=u, device=torch.device("cpu"))
    assert fb.attrs["painted"] is True
    edges = orc.bin_edges(0.5 * kf, 5.0 * kf, kf)
    e6, idx = orc.triangles_all(edges, 1)
    out = tmp_path / "b.dat"
    r = fb.measure_bispectrum_faster(imin=2, imax=7, out_file=str(out))
    want = orc.measure_unnormalized([mesh], box, edges, idx[2:7], pos_units=u)
    np.testing.assert_allclose(r["B"], want, rtol=1e-10)
    assert np.array_equal(r["index"], np.arange(2, 7)) and np.array_equal(r["k_edge"], e6[2:7])
    rows = out.read_text().strip().splitlines()
    assert len(rows) == 5
    assert rows[0] == "%d %e %e %e %e %e %e %e" % ((2,) + tuple(e6[2]) + (r["B"][0],))
    # imax beyond the list end is clipped, None means "to the end" (superset of the reference)
    r2 = fb.measure_bispectrum_faster(imin=len(idx) - 2, imax=10 ** 6)
    assert len(r2["B"]) == 2
    r3 = fb.measure_bispectrum_faster()
    assert len(r3["B"]) == len(idx)
    # slow path 'full' = unnormalised / N_tri with edges scaled by the unit factor (main.py:1708)
    full = fb.measure_bispectrum(0, 3, out_file=str(tmp_path / "full.dat"), meas_type="full")
    ed, tr = np.unique((e6[0:3] * u).reshape(-1, 2), axis=0, return_inverse=True)
    tr = np.asarray(tr).reshape(-1, 3)
    wn, wk = orc.measure_gridinfo(n, box, ed, tr, pos_units=u)
    wb = orc.measure_unnormalized([mesh], box, ed, tr, pos_units=u)
    np.testing.assert_allclose(full["B"], wb / wn, rtol=1e-9)
    np.testing.assert_allclose(full["k_mean"], wk, rtol=1e-10)
    cols = np.loadtxt(tmp_path / "full.dat")
    assert cols.shape == (3, 12) and fmt_e(cols[1, 11]) == fmt_e(full["N_tri"][1])
    un = fb.measure_bispectrum(0, 2, out_file=str(tmp_path / "un.dat"), meas_type="unnorm_b_value")
    assert (tmp_path / "un.dat").read_text().splitlines()[0] == "%d %e" % (0, un["B"][0])
    # save_bispectrum and pickling of the measured state (main.py:1598-1606, 2135-2161)
    fb.save_bispectrum(str(tmp_path / "all.dat"))
    assert np.loadtxt(tmp_path / "all.dat").shape == (len(fb.b["B"]), 11)
    state = pickle.loads(pickle.dumps(fb.__getstate__()))
    assert set(state) == {"b", "k_edges", "attrs"}


def test_explicit_k_edges_and_cross_routing():
    n, box = 16, 200.0
    kf = 2 * np.pi / box
    a, b = _mesh(n, 1), _mesh(n, 2)
    k_edges = np.array([[1.5 * kf, 2.5 * kf, 0.5 * kf, 1.5 * kf, 0.5 * kf, 1.5 * kf],
                        [2.5 * kf, 3.5 * kf, 1.5 * kf, 2.5 * kf, 0.5 * kf, 1.5 * kf]])
    fb = bk.FFTBispectrum(a, BoxSize=box, k_edges=k_edges, second=b, device=torch.device("cpu"))
    assert fb.num_fields == 2 and fb.k_indices is None
    r = fb.measure_bispectrum(0, 2, meas_type="unnorm_b_value")
    ed, tr = np.unique(k_edges.reshape(-1, 2), axis=0, return_inverse=True)
    want = orc.measure_unnormalized([a, b], box, ed, np.asarray(tr).reshape(-1, 3))
    np.testing.assert_allclose(r["B"], want, rtol=1e-10)
    # the fast path also runs with explicit k_edges (the reference cannot)
    r2 = fb.measure_bispectrum_faster(0, 2)
    np.testing.assert_allclose(r2["B"], want, rtol=1e-10)
    with pytest.raises(ValueError):
        fb.measure_bispectrum(0, 1, meas_type="nonsense")


def test_empty_bins_give_zero_count_and_nan_kmean():
    # bins below the fundamental mode are empty: N_tri = 0 exactly, k_mean = NaN
    n, box = 16, 100.0
    kf = 2 * np.pi / box
    fb = bk.FFTBispectrum(np.zeros((n, n, n)), BoxSize=box, kmin=0.1 * kf, kmax=2.7 * kf, dk=0.5 * kf,
                          triangle_type="equilateral", for_grid_info_only=True,
                          device=torch.device("cpu"))
    r = fb.measure_gridinfo_faster()
    assert r["N_tri"][0] == 0 and np.isnan(r["k_mean"][0]).all()
    assert r["N_tri"][-1] > 0 and np.isfinite(r["k_mean"][-1]).all()


def test_subbox_helpers_and_subbox_gridinfo_golden(tmp_path):
    """Sub-box path: index helpers (main.py:30-81), exact sub-cube extraction, and
    the sub-box golden: bins stay in full-box units, so in the L=500 sub-box every triangle
    touching bin 0 is empty (N_tri = 0, k_mean = nan) and the others match the golden digits."""
    assert bk.subbox_multiindex_to_index((1, 0, 1), 2) == 5
    assert np.array_equal(bk.subbox_index_to_multiindex(5, 2), [1, 0, 1])
    for i in range(8):
        assert bk.subbox_multiindex_to_index(bk.subbox_index_to_multiindex(i, 2), 2) == i
    full = np.arange(16 ** 3, dtype=np.float64).reshape(16, 16, 16)
    sub = bk.field_subbox_pm(np.array([1, 0, 1]), 2, bk.ArrayMesh(full, 1000.0))
    assert np.array_equal(sub.array, full[8:16, 0:8, 8:16])
    assert np.array_equal(sub.attrs["BoxSize"], [500.0] * 3) and np.array_equal(sub.attrs["Nmesh"], [8] * 3)
    with pytest.raises(ValueError):
        bk.field_subbox_pm((0, 0, 0), 3, bk.ArrayMesh(full, 1000.0))
    # sub-box of a 96^3 mesh -> 48^3, L = 500: first block of the reference's sub-box golden
    g = np.loadtxt(os.path.join(REF_OUT, "Lbox1000_512_kf_3kf_3lowkbins_subbox0.dat"))[:59]
    src = bk.field_subbox_pm((0, 0, 0), 2, bk.ArrayMesh(np.zeros((96, 96, 96)), 1000.0))
    fb = bk.FFTBispectrum(src, for_grid_info_only=True, device=torch.device("cpu"), **BINS)
    r = fb.measure_gridinfo_faster(0, 200)
    for t in range(59):
        if abs(g[t, 10]) < 0.5 or not np.isfinite(g[t, 1:4]).all():
            assert r["N_tri"][t] == 0 and np.isnan(r["k_mean"][t]).all()
        else:
            assert fmt_e(r["N_tri"][t]) == fmt_e(g[t, 10])
            assert [fmt_e(v) for v in r["k_mean"][t]] == [fmt_e(v) for v in g[t, 1:4]]


def test_non_cubic_box_and_dtype_override():
    """BoxSize with three different side lengths: per-axis wavenumber tables and the crop radius
    (from the longest side) must reproduce the oracle; compute_dtype overrides the mesh dtype."""
    n = 16
    box = np.array([200.0, 160.0, 240.0])
    kf = 2 * np.pi / box.max()
    mesh = _mesh(n, 9)
    fb = bk.FFTBispectrum(mesh, BoxSize=box, kmin=0.6 * kf, kmax=5.1 * kf, dk=1.1 * kf,
                          device=torch.device("cpu"))
    edges = orc.bin_edges(0.6 * kf, 5.1 * kf, 1.1 * kf)
    _, idx = orc.triangles_all(edges, 1)
    r = fb.measure_bispectrum_faster()
    g = fb.measure_gridinfo_faster()
    want = orc.measure_unnormalized([mesh], box, edges, idx)
    wn, wk = orc.measure_gridinfo(n, box, edges, idx)
    np.testing.assert_allclose(r["B"], want, rtol=1e-10, atol=1e-12 * np.abs(want).max())
    assert np.array_equal(g["N_tri"], np.rint(wn))
    ok = wn > 0.5
    np.testing.assert_allclose(g["k_mean"][ok], wk[ok], rtol=1e-10)
    assert fb._meas().precision == bkmain.F64
    fb32 = bk.FFTBispectrum(mesh, BoxSize=box, kmin=0.6 * kf, kmax=5.1 * kf, dk=1.1 * kf,
                            compute_dtype=np.float32, device=torch.device("cpu"))
    assert fb32._meas().precision == bkmain.F32


def test_measure_subboxes_driver(tmp_path):
    """The per-sub-box loop of scripts/measure/measure_subbox_bs_fast.py:246-295: every sub-box
    is measured with Nmesh/nsub, BoxSize/nsub and unchanged k bins, one output file per sub-box,
    and equals a direct measurement of the sliced sub-cube."""
    n, box = 16, 400.0
    kf_sub = 2 * np.pi / (box / 2)
    mesh = _mesh(n, seed=9)
    bins = dict(kmin=0.5 * kf_sub, kmax=3.6 * kf_sub, dk=kf_sub)
    prefix = str(tmp_path / "sb")
    res = bk.measure_subboxes(bk.ArrayMesh(mesh, box), 2, 3, 5, out_file_prefix=prefix,
                              device=torch.device("cpu"), **bins)
    assert sorted(res) == [3, 4, 5]
    for ind in (3, 4, 5):
        a, b, c = (int(v) for v in bk.subbox_index_to_multiindex(ind, 2))
        cube = mesh[8 * a:8 * a + 8, 8 * b:8 * b + 8, 8 * c:8 * c + 8]
        fb = bk.FFTBispectrum(np.ascontiguousarray(cube), BoxSize=box / 2, device=torch.device("cpu"), **bins)
        want = fb.measure_bispectrum_faster(0, 10 ** 6)
        np.testing.assert_allclose(res[ind]["B"], want["B"], rtol=1e-12, atol=0)
        rows = np.loadtxt(prefix + "_subbox%d.dat" % ind)
        assert rows.shape == (len(want["B"]), 8)
        np.testing.assert_allclose(rows[:, 7], want["B"], rtol=2e-6)
    gi = bk.measure_subboxes(bk.ArrayMesh(mesh, box), 2, 0, 0, meas_type="grid_info",
                             device=torch.device("cpu"), **bins)
    assert np.all(gi[0]["N_tri"] >= 0) and gi[0]["k_mean"].shape[1] == 3
    with pytest.raises(ValueError):
        bk.measure_subboxes(bk.ArrayMesh(mesh, box), 2, 0, 8, device=torch.device("cpu"), **bins)
    with pytest.raises(ValueError):
        bk.measure_subboxes(bk.ArrayMesh(mesh, box), 2, meas_type="full", device=torch.device("cpu"), **bins)
