"""float64 numpy/scipy restatement of bskit's FFT bispectrum estimator.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).

Why a restatement: the reference (``/root/reference/bskit/main.py``) delegates
all arithmetic to nbodykit / pmesh / pfft / mpi4py, none of which is vendored,
pinned or installable here (the reference also uses
``np.float``/``np.int`` which numpy >= 1.24 removed, so it does not even
import).  Each function below names the reference lines it follows.

Parity pinning
--------------
* PINNED: bin edges, triangle lists, N_tri and k_mean against the reference's
  golden ``examples/tests/output_ref/Lbox1000_512_kf_3kf_3lowkbins.dat``
  (59/59 rows to every printed digit) and the first block of
  ``..._subbox0.dat`` (57/57 non-empty rows); see ``tests/test_oracle_golden.py``.
* PINNED: the B estimator and its normalisation against a brute-force sum over
  closed triangles on tiny grids (the idea of the reference's own
  ``bk_binned``, main.py:834-1004); see ``tests/test_oracle_bruteforce.py``.
* PARITY UNPINNED: B *values* of the shipped goldens
  (``test_grid_512_1_unnormbs_*.dat``).  They depend on the pmesh random-number
  stream of ``nbk.LinearMesh(seed)`` which cannot be regenerated without pmesh.
  Only their column layout / ordering is checked.

Conventions (pmesh/nbodykit, recalled and then verified by the pins above):
``k_axis = 2*pi*fftfreq(N, 1/N)/L``; half spectrum on the last axis; forward
transform divided by N^3, inverse un-normalised.
"""
from __future__ import annotations

import itertools

import numpy as np
import scipy.fft as sfft


# --------------------------------------------------------------------------- #
# bin specifications (host side of the path)
# --------------------------------------------------------------------------- #
def bin_edges(kmin, kmax, dk, num_lowk_bins=0, dk_high=-1.0):
    """(S,2) lower/upper k-bin edges.  Follows main.py:1008-1071."""
    if kmin <= 0.0:
        raise ValueError("kmin must be > 0!")
    if kmax <= 0.0:
        raise ValueError("kmax must be > 0!")
    if dk <= 0.0:
        raise ValueError("dk must be > 0!")
    if num_lowk_bins > 0 and dk_high < 0.0:
        raise ValueError("Must specify dk_high if num_lowk_bins > 0!")
    lo = np.arange(kmin, kmax - dk, dk)                       # main.py:1052
    width = np.ones_like(lo) * dk
    hi = lo + width                                           # main.py:1054
    if 0 < num_lowk_bins < len(lo):                           # main.py:1059
        lo2 = np.arange(hi[num_lowk_bins - 1], kmax - dk_high, dk_high)
        hi2 = lo2 + np.ones_like(lo2) * dk_high
        lo = np.hstack((lo[:num_lowk_bins], lo2))
        hi = np.hstack((hi[:num_lowk_bins], hi2))
    return np.vstack((lo, hi)).T


def geometric_k_mean(kmin, kmax):
    """main.py:408-428."""
    return 0.75 * (kmax ** 4 - kmin ** 4) / (kmax ** 3 - kmin ** 3)


def _closed(t):
    # main.py:1334-1340: closure test on the *upper* edges, all permutations
    return (t[0][1] + t[1][1] >= t[2][0]
            and t[0][1] + t[2][1] >= t[1][0]
            and t[1][1] + t[2][1] >= t[0][0])


def triangles_all(edges, num_fields=1):
    """All closed bin triples -> ((T,6) edges, (T,3) k-bin indices).

    Follows main.py:1272-1389 (itertools enumeration, closure filter, column
    swap, structured sort on the three lower edges, index lookup by exact
    float equality on the lower edges).
    """
    if num_fields not in (1, 2, 3):
        raise ValueError("num_fields must be 1, 2, or 3!")
    edges = np.asarray(edges, dtype=np.float64)
    if num_fields == 1:
        raw = [t for t in itertools.combinations_with_replacement(edges, 3) if _closed(t)]
        arr = np.array(raw)
        arr = arr[:, ::-1, :].copy()                          # (c,b,a): k1>=k2>=k3
    elif num_fields == 2:
        pairs = list(itertools.combinations_with_replacement(edges, 2))
        raw = [(a[0], a[1], b) for a in pairs for b in edges]
        arr = np.array([t for t in raw if _closed(t)])
        arr = arr[:, [1, 0, 2], :].copy()                     # main.py:1363-1365
    else:
        raw = [t for t in itertools.product(edges, repeat=3) if _closed(t)]
        arr = np.array(raw)
    flat = arr.reshape(arr.shape[0], 6)
    rec = flat.copy().view("f8,f8,f8,f8,f8,f8")
    flat = np.sort(rec, order=["f0", "f2", "f4"], axis=0).view(np.float64)
    lows = edges[:, 0]
    idx = np.array([[np.where(lows == row[c])[0][0] for c in (0, 2, 4)] for row in flat])
    return flat, idx


def triangles_equilateral(edges):
    """main.py:1074-1118."""
    edges = np.asarray(edges, dtype=np.float64)
    i = np.arange(len(edges))
    return np.hstack((edges, edges, edges)), np.vstack((i, i, i)).T


def triangles_squeezed(edges, squeezed_bin_index=0):
    """main.py:1121-1178: (q, s, s) for every s > q."""
    edges = np.asarray(edges, dtype=np.float64)
    q = squeezed_bin_index
    s = np.arange(q + 1, len(edges))
    sq = np.array([edges[q] for _ in s]).reshape(len(s), 2)
    return (np.hstack((sq, edges[q + 1:], edges[q + 1:])),
            np.vstack((np.full(len(s), q), s, s)).T)


def triangles_isosceles(edges, isos_mult, isos_tol=0.1):
    """main.py:1181-1269: (L(s), s, s), L = bin whose geometric mean is
    nearest mean_s/isos_mult, kept when the fractional miss < isos_tol."""
    if isos_mult < 1.0:
        raise ValueError("isos_mult must be greater than 1! Come on man...")
    edges = np.asarray(edges, dtype=np.float64)
    kmean = np.array([geometric_k_mean(e[0], e[1]) for e in edges])
    target = kmean / isos_mult
    near, dev = [], []
    for v in target:
        d = np.abs(kmean - v)
        near.append(int(d.argmin()))
        dev.append(d.min() / v)
    near = np.array(near)
    keep = np.array(dev) < isos_tol
    s = np.arange(len(edges))[keep]
    big = near[keep]
    return np.hstack((edges[big], edges[s], edges[s])), np.vstack((big, s, s)).T


# --------------------------------------------------------------------------- #
# grid conventions
# --------------------------------------------------------------------------- #
def _box3(box):
    b = np.atleast_1d(np.asarray(box, dtype=np.float64))
    return np.ones(3) * b if b.size == 1 else b


def k_axes(nmesh, box):
    """Per-axis wavenumber tables (x, y full; z half).  pmesh convention."""
    L = _box3(box)
    n = int(nmesh)
    kx = 2.0 * np.pi * np.fft.fftfreq(n, 1.0 / n) / L[0]
    ky = 2.0 * np.pi * np.fft.fftfreq(n, 1.0 / n) / L[1]
    kz = 2.0 * np.pi * np.fft.rfftfreq(n, 1.0 / n) / L[2]
    return kx, ky, kz


def k_norm(nmesh, box):
    """|k| on the half-spectrum grid: ``sum(ki**2. for ki in k)**0.5``
    (main.py:266, 316, 384, 618, 1850)."""
    kx, ky, kz = k_axes(nmesh, box)
    parts = (kx[:, None, None], ky[None, :, None], kz[None, None, :])
    return sum(ki ** 2.0 for ki in parts) ** 0.5


def shell_mask(kk, lo, hi):
    """Both ends inclusive (main.py:268, 320, 620, 1852)."""
    return (kk <= hi) & (kk >= lo)


def hermitian_weights(nmesh):
    """Multiplicity of each stored half-spectrum mode in the full cube."""
    n = int(nmesh)
    w = np.full(n // 2 + 1, 2, dtype=np.int64)
    w[0] = 1
    if n % 2 == 0:
        w[-1] = 1
    return w


def modes_per_bin(nmesh, box, edges):
    """Exact integer number of full-cube modes in each k-bin."""
    kk = k_norm(nmesh, box)
    w = hermitian_weights(nmesh)[None, None, :]
    return np.array([int((shell_mask(kk, lo, hi) * w).sum()) for lo, hi in np.asarray(edges)],
                    dtype=np.int64)


def forward(mesh, workers=None):
    """delta_k = rfftn(delta)/N^3  (pmesh r2c; main.py:1612)."""
    mesh = np.asarray(mesh)
    return sfft.rfftn(mesh, workers=workers) / mesh.size


def compensate_cic(delta_k, nmesh, nmesh_cic):
    """CIC window compensation at the original resolution
    (scripts/measure/measure_bs_fast.py:45-57; kind='circular')."""
    n = int(nmesh)
    w_full = 2.0 * np.pi * np.fft.fftfreq(n)
    w_half = 2.0 * np.pi * np.fft.rfftfreq(n)
    v = delta_k
    for ax, w in enumerate((w_full, w_full, w_half)):
        shape = [1, 1, 1]
        shape[ax] = -1
        v = v / (1.0 - 2.0 / 3.0 * np.sin(0.5 * w.reshape(shape) * nmesh / nmesh_cic) ** 2) ** 0.5
    return v


def to_real(spec, nmesh, workers=None):
    """Un-normalised inverse (pmesh c2r)."""
    n = int(nmesh)
    return sfft.irfftn(spec, s=(n, n, n), workers=workers) * float(n) ** 3


def data_shell(delta_k, kk, lo, hi, workers=None):
    """I_i(x) = c2r(delta_k * mask)  (main.py:1846-1861, 614-660)."""
    return to_real(delta_k * shell_mask(kk, lo, hi), delta_k.shape[0], workers)


def number_field(nmesh, box, lo, hi, kk=None, workers=None):
    """n_i(x) = c2r(mask)  (main.py:280-329)."""
    kk = k_norm(nmesh, box) if kk is None else kk
    return to_real(shell_mask(kk, lo, hi).astype(np.complex128), nmesh, workers)


def k_field(nmesh, box, lo, hi, p, kk=None, workers=None):
    """kappa_i(x) = c2r(|k|^p * mask)  (main.py:227-277)."""
    kk = k_norm(nmesh, box) if kk is None else kk
    return to_real((kk ** p * shell_mask(kk, lo, hi)).astype(np.complex128), nmesh, workers)


# --------------------------------------------------------------------------- #
# measurements
# --------------------------------------------------------------------------- #
def _route(num_fields):
    """Which input mesh feeds k1,k2,k3 (main.py:627-640): <AAA>, <AAB>, <ABC>."""
    return {1: (0, 0, 0), 2: (0, 0, 1), 3: (0, 1, 2)}[num_fields]


def measure_unnormalized(meshes, box, edges, triples, pos_units=1.0,
                         nmesh_cic=None, workers=None, delta_k=None):
    """Unnormalised B for every index triple.

    ``B = sum_x I_a I_b I_c * V^2 / N^3 * u^6`` (main.py:1871-1882; cross
    routing main.py:627-640).  ``meshes`` is a list of 1-3 real (N,N,N)
    arrays, promoted to float64.
    """
    if delta_k is None:
        fields = [np.asarray(m, dtype=np.float64) for m in meshes]
        n = fields[0].shape[0]
        delta_k = [forward(f, workers) for f in fields]
        if nmesh_cic:
            delta_k = [compensate_cic(d, n, nmesh_cic) for d in delta_k]
    n = delta_k[0].shape[0]
    L = _box3(box)
    kk = k_norm(n, L)
    edges = np.asarray(edges, dtype=np.float64)
    triples = np.asarray(triples, dtype=np.int64).reshape(-1, 3)
    route = _route(len(delta_k))
    cache = {}

    def shell(slot, ibin):
        key = (route[slot], int(ibin))
        if key not in cache:
            cache[key] = data_shell(delta_k[key[0]], kk, edges[ibin, 0], edges[ibin, 1], workers)
        return cache[key]

    norm = L.prod() ** 2 / float(n) ** 3
    out = np.empty(len(triples))
    for t, (a, b, c) in enumerate(triples):
        out[t] = np.sum(shell(0, a) * shell(1, b) * shell(2, c)) * norm * pos_units ** 6.0
    return out


def measure_unnormalized_dense(meshes, box, edges, triples, pos_units=1.0, workers=None,
                               progress=False):
    """Same numbers as :func:`measure_unnormalized` (main.py:1871-1882) for long triangle lists
    on large grids: every shell is built once (as the reference's fast path does,
    main.py:1846-1861), the product I_a*I_b is formed once per (a,b) pair and the full-grid sum
    ``np.sum((I_a*I_b)*I_c)`` is evaluated as a float64 dot product, x-slabs on a thread pool.
    Used to generate the committed 512^3 fixture (scripts/make_golden_metric512.py)."""
    from concurrent.futures import ThreadPoolExecutor
    import time
    fields = [np.asarray(m, dtype=np.float64) for m in meshes]
    n = fields[0].shape[0]
    delta_k = [forward(f, workers) for f in fields]
    del fields
    L = _box3(box)
    kk = k_norm(n, L)
    edges = np.asarray(edges, dtype=np.float64)
    triples = np.asarray(triples, dtype=np.int64).reshape(-1, 3)
    route = _route(len(delta_k))
    shells = {}
    for slot in range(3):
        for ibin in np.unique(triples[:, slot]):
            key = (route[slot], int(ibin))
            if key not in shells:
                shells[key] = data_shell(delta_k[key[0]], kk, edges[ibin, 0], edges[ibin, 1], workers)
    del kk, delta_k
    nw = max(1, int(workers or 1))
    slabs = [(int(s[0]), int(s[-1]) + 1) for s in np.array_split(np.arange(n), min(nw, n)) if len(s)]
    pool = ThreadPoolExecutor(nw)
    norm = L.prod() ** 2 / float(n) ** 3
    out = np.empty(len(triples))
    order = np.lexsort((triples[:, 2], triples[:, 1], triples[:, 0]))
    prod = np.empty((n, n, n))
    last = None
    t0 = time.time()
    for cnt, t in enumerate(order):
        a, b, c = (int(v) for v in triples[t])
        fa, fb, fc = shells[(route[0], a)], shells[(route[1], b)], shells[(route[2], c)]
        if last != (a, b):
            list(pool.map(lambda s: np.multiply(fa[s[0]:s[1]], fb[s[0]:s[1]], out=prod[s[0]:s[1]]), slabs))
            last = (a, b)
        parts = pool.map(lambda s: float(np.dot(prod[s[0]:s[1]].reshape(-1), fc[s[0]:s[1]].reshape(-1))), slabs)
        out[t] = sum(parts) * norm * pos_units ** 6.0
        if progress and cnt % 500 == 0:
            print(f"  triangle {cnt}/{len(order)}  {time.time() - t0:.0f}s", flush=True)
    pool.shutdown()
    return out


def measure_gridinfo(nmesh, box, edges, triples, pos_units=1.0, workers=None):
    """(N_tri, k_mean[T,3]) per triple (main.py:2006-2064).

    ``N_tri = sum n_a n_b n_c / N^3``; ``k1 = sum kappa_a n_b n_c / N^3 / N_tri``
    etc.; NaN where the division fails (main.py:2037-2057).
    """
    n = int(nmesh)
    L = _box3(box)
    kk = k_norm(n, L)
    edges = np.asarray(edges, dtype=np.float64)
    triples = np.asarray(triples, dtype=np.int64).reshape(-1, 3)
    used = sorted(set(triples.ravel().tolist()))
    nf = {i: number_field(n, L, edges[i, 0], edges[i, 1], kk, workers) for i in used}
    kf = {i: k_field(n, L, edges[i, 0], edges[i, 1], 1.0, kk, workers) for i in used}
    ntri = np.empty(len(triples))
    kmean = np.empty((len(triples), 3))
    n3 = float(n) ** 3
    for t, (a, b, c) in enumerate(triples):
        nb = np.sum(nf[a] * nf[b] * nf[c]) / n3
        ntri[t] = nb
        with np.errstate(divide="ignore", invalid="ignore"):
            kmean[t, 0] = np.sum(kf[a] * nf[b] * nf[c]) / n3 / nb
            kmean[t, 1] = np.sum(nf[a] * kf[b] * nf[c]) / n3 / nb
            kmean[t, 2] = np.sum(nf[a] * nf[b] * kf[c]) / n3 / nb
    return ntri, kmean / pos_units


def pk_fft(mesh, box, lo, hi, workers=None):
    """FFT-estimator P(k) in one bin -> (P, N_modes, k_mean)  (main.py:333-405)."""
    mesh = np.asarray(mesh, dtype=np.float64)
    n = mesh.shape[0]
    L = _box3(box)
    kk = k_norm(n, L)
    n3 = float(n) ** 3
    nbin = np.sum(number_field(n, L, lo, hi, kk, workers) ** 2.0) / n3
    kmean = np.sum(k_field(n, L, lo, hi, 0.5, kk, workers) ** 2.0) / n3 / nbin
    shell = data_shell(forward(mesh, workers), kk, lo, hi, workers)
    return np.sum(shell ** 2.0) * L.prod() / n3 / nbin, nbin, kmean


# --------------------------------------------------------------------------- #
# brute force over explicit closed triangles (tiny grids only)
# --------------------------------------------------------------------------- #
def full_spectrum(mesh):
    """Full complex cube delta_k = fftn(delta)/N^3 and integer mode numbers."""
    mesh = np.asarray(mesh, dtype=np.float64)
    return np.fft.fftn(mesh) / mesh.size


def brute_force_triangle(meshes, box, bins3):
    """Sum delta(k1) delta(k2) delta(k3) over all k1+k2+k3 = 0 (mod N) with
    |k_i| in bins3[i] -> (B_unnormalised, N_tri, k_mean[3]).

    Independent of any FFT-estimator algebra; same idea as the reference's
    own cross-check ``bk_binned`` (main.py:834-1004) but closing modulo N as
    the FFT estimator does.  O(M1*M2) with M = modes per bin.
    """
    cubes = [full_spectrum(m) for m in meshes]
    n = cubes[0].shape[0]
    L = _box3(box)
    route = _route(len(cubes))
    f = np.fft.fftfreq(n, 1.0 / n).astype(np.int64)
    ix, iy, iz = np.meshgrid(f, f, f, indexing="ij")
    kx = 2.0 * np.pi * np.fft.fftfreq(n, 1.0 / n) / L[0]
    ky = 2.0 * np.pi * np.fft.fftfreq(n, 1.0 / n) / L[1]
    kz = 2.0 * np.pi * np.fft.fftfreq(n, 1.0 / n) / L[2]
    kk = sum(ki ** 2.0 for ki in (kx[:, None, None], ky[None, :, None], kz[None, None, :])) ** 0.5
    sel = [np.argwhere(shell_mask(kk, b[0], b[1])) for b in bins3]   # index triples in-cube
    in3 = shell_mask(kk, bins3[2][0], bins3[2][1])
    c1, c2, c3 = (cubes[route[0]], cubes[route[1]], cubes[route[2]])
    total = 0.0 + 0.0j
    count = 0
    ksum = np.zeros(3)
    for p1 in sel[0]:
        n1 = f[p1]
        v1 = c1[tuple(p1)]
        k1 = kk[tuple(p1)]
        for p2 in sel[1]:
            n3 = (-(n1 + f[p2])) % n                           # cube index of k3
            if in3[n3[0], n3[1], n3[2]]:
                total += v1 * c2[tuple(p2)] * c3[n3[0], n3[1], n3[2]]
                count += 1
                ksum += (k1, kk[tuple(p2)], kk[n3[0], n3[1], n3[2]])
    b = total.real * L.prod() ** 2
    with np.errstate(divide="ignore", invalid="ignore"):
        return b, count, ksum / count


def paint_cic(pos, nmesh, boxsize):
    """Cloud-in-cell mass assignment, float64 (test oracle for bsk_paint_cic; the reference gets
    this from nbodykit's ``catalog.to_mesh(window='cic')``, scripts/measure/measure_bs_fast.py:
    209-217).  Mesh points at integer multiples of BoxSize/Nmesh, weights prod (1 - |g - i|) on
    the 8 surrounding points, periodic.  Returns the weight sums (N,N,N)."""
    pos = np.asarray(pos, dtype=np.float64)
    box = np.ones(3) * np.asarray(boxsize, dtype=np.float64)
    n = int(nmesh)
    g = pos * (n / box)[None, :]
    f = np.floor(g)
    d = g - f
    i0 = np.mod(f.astype(np.int64), n)
    i1 = np.mod(i0 + 1, n)
    mesh = np.zeros((n, n, n))
    for ax in (0, 1):
        for ay in (0, 1):
            for az in (0, 1):
                w = ((d[:, 0] if ax else 1 - d[:, 0]) * (d[:, 1] if ay else 1 - d[:, 1]) *
                     (d[:, 2] if az else 1 - d[:, 2]))
                np.add.at(mesh, ((i1 if ax else i0)[:, 0], (i1 if ay else i0)[:, 1], (i1 if az else i0)[:, 2]), w)
    return mesh
