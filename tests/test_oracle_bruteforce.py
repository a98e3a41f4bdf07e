"""Pin the oracle's FFT estimator against an explicit sum over closed triangles
(the idea of the reference's own bk_binned cross-check, main.py:834-1004)."""
import numpy as np
import pytest

from oracle import bskit_oracle as orc


def _field(n, seed):
    rng = np.random.default_rng(seed)
    g = rng.standard_normal((n, n, n))
    return g + 0.4 * g ** 2 - 0.4          # non-Gaussian, so B != 0


@pytest.mark.parametrize("n", [10, 12])
def test_fft_estimator_equals_bruteforce_auto(n):
    box = 100.0
    kf = 2 * np.pi / box
    edges = np.array([[0.5 * kf, 1.5 * kf], [1.5 * kf, 2.5 * kf], [2.5 * kf, 3.5 * kf]])
    triples = np.array([[0, 0, 0], [1, 0, 0], [1, 1, 0], [1, 1, 1], [2, 1, 1], [2, 2, 1]])
    m = _field(n, 3)
    b = orc.measure_unnormalized([m], box, edges, triples)
    ntri, kmean = orc.measure_gridinfo(n, box, edges, triples)
    for t, tr in enumerate(triples):
        bb, cnt, km = orc.brute_force_triangle([m], box, [edges[i] for i in tr])
        assert abs(ntri[t] - cnt) < 1e-6 * max(cnt, 1)
        assert abs(b[t] - bb) <= 1e-11 * max(abs(bb), 1e-30)
        np.testing.assert_allclose(kmean[t], km, rtol=1e-11)


def test_known_triangle_counts():
    # counts 120, 174, 456 for golden rows 0,1,2 (bins [.5,1.5],[1.5,2.5] k_f)
    kf = 2 * np.pi / 1000.0
    edges = np.array([[0.5 * kf, 1.5 * kf], [1.5 * kf, 2.5 * kf]])
    ntri, _ = orc.measure_gridinfo(16, 1000.0, edges, [[0, 0, 0], [1, 0, 0], [1, 1, 0]])
    assert [int(round(v)) for v in ntri] == [120, 174, 456]


def test_cross_routing_aab_and_abc():
    n, box = 10, 50.0
    kf = 2 * np.pi / box
    edges = np.array([[0.5 * kf, 1.5 * kf], [1.5 * kf, 2.5 * kf]])
    a, b_, c = _field(n, 1), _field(n, 2), _field(n, 5)
    for meshes in ([a, b_], [a, b_, c]):
        triples = np.array([[1, 0, 0], [1, 1, 0], [0, 0, 1]])
        got = orc.measure_unnormalized(meshes, box, edges, triples)
        for t, tr in enumerate(triples):
            bb, _, _ = orc.brute_force_triangle(meshes, box, [edges[i] for i in tr])
            assert abs(got[t] - bb) <= 1e-11 * abs(bb)
    # <AAB> != <BBA>  (the two cross goldens differ)
    x = orc.measure_unnormalized([a, b_], box, edges, [[1, 1, 0]])
    y = orc.measure_unnormalized([b_, a], box, edges, [[1, 1, 0]])
    assert abs(x[0] - y[0]) > 1e-6 * abs(x[0])


def test_pk_fft_matches_direct_mode_average():
    n, box = 12, 80.0
    kf = 2 * np.pi / box
    m = _field(n, 7)
    lo, hi = 1.5 * kf, 2.5 * kf
    p, nb, km = orc.pk_fft(m, box, lo, hi)
    cube = orc.full_spectrum(m)
    f = 2 * np.pi * np.fft.fftfreq(n, 1.0 / n) / box
    kk = np.sqrt(f[:, None, None] ** 2 + f[None, :, None] ** 2 + f[None, None, :] ** 2)
    sel = (kk >= lo) & (kk <= hi)
    assert abs(nb - sel.sum()) < 1e-8
    np.testing.assert_allclose(p, (np.abs(cube[sel]) ** 2).mean() * box ** 3, rtol=1e-11)
    np.testing.assert_allclose(km, kk[sel].mean(), rtol=1e-11)


def test_oracle_cic_painting_properties():
    """CIC oracle: a particle on a mesh point puts all its weight there, one at a
    cell centre 1/8 on each corner, periodic wrap, and the total weight equals the particle count."""
    from oracle import bskit_oracle as orc
    n, box = 8, 16.0
    m = orc.paint_cic(np.array([[4.0, 6.0, 2.0]]), n, box)
    assert m[2, 3, 1] == 1.0 and m.sum() == 1.0
    m = orc.paint_cic(np.array([[15.0, 15.0, 15.0]]), n, box)           # centre of the last cell: wraps
    assert np.allclose(m[[7, 0]][:, [7, 0]][:, :, [7, 0]], 0.125) and np.isclose(m.sum(), 1.0)
    m = orc.paint_cic(np.array([[-1.0, 17.0, 33.0]]), n, box)           # outside the box: periodic
    assert np.isclose(m.sum(), 1.0) and np.isclose(m[7, 0, 0], 0.125)
    rng = np.random.default_rng(0)
    pos = rng.uniform(0, box, size=(1000, 3))
    assert np.isclose(orc.paint_cic(pos, n, box).sum(), 1000.0)


def test_tiled_mesh_identity_behind_the_c5_check():
    """scripts/make_golden_c5.py: a mesh that repeats a small field t times per axis has
    B_big = t^6 B_small for the same physical bins (here t = 2, 16^3 -> 32^3, aliased bins included)."""
    from bskit_b200 import synthetic as syn
    t, ns, nb = 2, 16, 6
    box = 200.0
    small = syn.gaussian_mesh(ns, seed=3, box=box).astype(np.float64)
    big = np.tile(small, (t, t, t))
    kmin, kmax, dk = syn.bench_bins(nb, box=box)
    edges = orc.bin_edges(kmin, kmax, dk)
    _, idx = orc.triangles_all(edges, 1)
    b_small = orc.measure_unnormalized([small], box, edges, idx)
    b_big = orc.measure_unnormalized([big], box * t, edges, idx)
    assert np.allclose(b_big, b_small * float(t) ** 6, rtol=1e-9, atol=1e-9 * np.abs(b_small).max() * t ** 6)
