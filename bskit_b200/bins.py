"""k-bin and triangle-bin specifications (host side of the bispectrum path).

Same names, argument order, return shapes and error behaviour as the
generators in the reference's ``bskit/main.py:1008-1389``; the enumeration is
vectorised index arithmetic instead of itertools over float pairs, so the
80-bin scheme (48 260 triangles) builds in milliseconds and the k-bin index
triples come for free instead of a float-equality search per triangle.

Bit-exactness contract (BASELINE.json: "k-bin and triangle-bin assignment
bit-exact"): edges are produced by the very same float64 numpy expressions as
the reference (``np.arange(kmin, kmax-dk, dk)``, ``lower + dk``), because the
*length* of that arange is fp-fragile and any other
formula would change which bins exist.
"""
from __future__ import annotations

import numpy as np

__all__ = [
    "generate_bin_edge_list",
    "geometric_k_mean",
    "generate_equilateral_triangle_bin_list",
    "generate_squeezed_triangle_bin_list",
    "generate_isosceles_triangle_bin_list",
    "generate_triangle_bin_list",
]


def generate_bin_edge_list(kmin=-1.0, kmax=-1.0, dk=-1.0, num_lowk_bins=0, dk_high=-1.0):
    """(Nbins, 2) array of lower/upper k-bin edges (ref. main.py:1008-1071).

    Lower edges run from ``kmin`` in steps of ``dk`` and stop before
    ``kmax - dk``; with ``num_lowk_bins > 0`` the bins after the first
    ``num_lowk_bins`` have width ``dk_high`` instead.
    """
    if kmin <= 0.0:
        raise ValueError("kmin must be > 0!")
    if kmax <= 0.0:
        raise ValueError("kmax must be > 0!")
    if dk <= 0.0:
        raise ValueError("dk must be > 0!")
    if num_lowk_bins > 0 and dk_high < 0.0:
        raise ValueError("Must specify dk_high if num_lowk_bins > 0!")

    lower = np.arange(kmin, kmax - dk, dk)
    upper = lower + np.ones_like(lower) * dk
    if 0 < num_lowk_bins < len(lower):
        n_low = int(num_lowk_bins)
        lower_hi = np.arange(upper[n_low - 1], kmax - dk_high, dk_high)
        upper_hi = lower_hi + np.ones_like(lower_hi) * dk_high
        lower = np.concatenate((lower[:n_low], lower_hi))
        upper = np.concatenate((upper[:n_low], upper_hi))
    return np.stack((lower, upper), axis=1)


def geometric_k_mean(kmin, kmax):
    """Volume-weighted mean |k| of a spherical shell (ref. main.py:408-428)."""
    return 0.75 * (kmax ** 4 - kmin ** 4) / (kmax ** 3 - kmin ** 3)


def _edges6(edges, idx):
    return np.hstack((edges[idx[:, 0]], edges[idx[:, 1]], edges[idx[:, 2]]))


def generate_equilateral_triangle_bin_list(kmin=-1.0, kmax=-1.0, dk=-1.0, num_lowk_bins=0,
                                           dk_high=-1.0, return_indices=False):
    """(i, i, i) for every k bin (ref. main.py:1074-1118)."""
    edges = generate_bin_edge_list(kmin, kmax, dk, num_lowk_bins, dk_high)
    i = np.arange(len(edges))
    idx = np.stack((i, i, i), axis=1)
    return idx if return_indices else _edges6(edges, idx)


def generate_squeezed_triangle_bin_list(kmin=-1.0, kmax=-1.0, dk=-1.0, squeezed_bin_index=0,
                                        num_lowk_bins=0, dk_high=-1.0, return_indices=False):
    """(q, s, s) for every s > q = ``squeezed_bin_index`` (ref. main.py:1121-1178)."""
    edges = generate_bin_edge_list(kmin, kmax, dk, num_lowk_bins, dk_high)
    q = int(squeezed_bin_index)
    s = np.arange(q + 1, len(edges))
    idx = np.stack((np.full_like(s, q), s, s), axis=1)
    return idx if return_indices else _edges6(edges, idx).reshape(len(s), 6)


def generate_isosceles_triangle_bin_list(kmin=-1.0, kmax=-1.0, dk=-1.0, isos_mult=0, isos_tol=0.1,
                                         num_lowk_bins=0, dk_high=-1.0, return_indices=False):
    """(L(s), s, s): for each k_S bin the k_L bin whose geometric mean is nearest
    ``mean(k_S)/isos_mult``, kept when the fractional miss is below ``isos_tol``
    (ref. main.py:1181-1269)."""
    if isos_mult < 1.0:
        raise ValueError("isos_mult must be greater than 1! Come on man...")
    edges = generate_bin_edge_list(kmin, kmax, dk, num_lowk_bins, dk_high)
    means = geometric_k_mean(edges[:, 0], edges[:, 1])
    target = means / isos_mult
    miss = np.abs(means[None, :] - target[:, None])        # [short bin, candidate long bin]
    nearest = miss.argmin(axis=1)
    frac = miss.min(axis=1) / target
    keep = frac < isos_tol
    short = np.arange(len(edges))[keep]
    idx = np.stack((nearest[keep], short, short), axis=1)
    return idx if return_indices else _edges6(edges, idx).reshape(len(short), 6)


def _closed(edges, a, b, c):
    """Closure on the *upper* edges for all three permutations (ref. main.py:1334-1340)."""
    lo, hi = edges[:, 0], edges[:, 1]
    return ((hi[a] + hi[b] >= lo[c]) & (hi[a] + hi[c] >= lo[b]) & (hi[b] + hi[c] >= lo[a]))


def _triangle_indices(edges, num_fields):
    n = len(edges)
    r = np.arange(n)
    if num_fields == 1:
        a, b, c = np.meshgrid(r, r, r, indexing="ij")
        keep = (a <= b) & (b <= c)
        a, b, c = a[keep], b[keep], c[keep]
        ok = _closed(edges, a, b, c)
        idx = np.stack((c[ok], b[ok], a[ok]), axis=1)         # k1 >= k2 >= k3
    elif num_fields == 2:
        a, b, c = np.meshgrid(r, r, r, indexing="ij")
        keep = a <= b
        a, b, c = a[keep], b[keep], c[keep]
        ok = _closed(edges, a, b, c)
        idx = np.stack((b[ok], a[ok], c[ok]), axis=1)         # k1 >= k2, k3 free
    else:
        a, b, c = (x.ravel() for x in np.meshgrid(r, r, r, indexing="ij"))
        ok = _closed(edges, a, b, c)
        idx = np.stack((a[ok], b[ok], c[ok]), axis=1)
    return idx


def generate_triangle_bin_list(kmin=-1.0, kmax=-1.0, dk=-1.0, mu_min=None, mu_max=None, dmu=None,
                               num_fields=1, num_lowk_bins=0, dk_high=-1.0, return_indices=False):
    """All closed (k1,k2,k3) bin triples (ref. main.py:1272-1389).

    One field: k1 >= k2 >= k3.  Two fields (<AAB>): k1 >= k2, any k3.  Three
    fields: every ordered triple.  Closure: each pair of upper edges must reach
    the third lower edge.  Rows are sorted by (k1_low, k2_low, k3_low).
    Returns (N_tri, 6) edges, or (N_tri, 3) k-bin indices with
    ``return_indices=True``.
    """
    if num_fields not in [1, 2, 3]:
        raise ValueError("num_fields must be 1, 2, or 3!")
    if mu_min is not None or mu_max is not None or dmu is not None:
        raise NotImplementedError("Mu binning not implemented yet!")
    edges = generate_bin_edge_list(kmin, kmax, dk, num_lowk_bins, dk_high)
    idx = _triangle_indices(edges, num_fields)
    e6 = _edges6(edges, idx).reshape(len(idx), 6)
    # the reference sorts the edge records on the three lower edges; the remaining
    # columns only break ties between bins sharing a lower edge
    order = np.lexsort((e6[:, 5], e6[:, 3], e6[:, 1], e6[:, 4], e6[:, 2], e6[:, 0]))
    if not return_indices:
        return e6[order]
    # the reference recovers indices as the FIRST bin whose lower edge equals the
    # triangle's (main.py:1388); identical to idx unless two bins share a lower edge
    first = np.array([np.flatnonzero(edges[:, 0] == lo)[0] for lo in edges[:, 0]], dtype=np.int64)
    return first[idx[order]]
