#!/usr/bin/env python
"""Oracle fixture for the decimated check of config C5 (2048^3, equilateral + squeezed bins).

A mesh that repeats a small seeded field t times along every axis has Fourier modes only on the
sub-lattice k = t * m * k_f, with the small field's amplitudes; its shell fields are the small
field's shell fields, repeated.  With the same physical bin edges

    B_big(a,b,c) = t^6 * B_small(a,b,c)       (V^2/N^3 grows by t^3, the cell sum by t^3)

so a float64 oracle run on the small field (here 256^3, BoxSize 1000/8) checks a 2048^3 GPU
measurement (BoxSize 1000) of the tiled mesh -- forward transform, shells and sums at full size --
without a 2048^3 CPU run.  tests/test_oracle_bruteforce.py checks the identity itself on the CPU.

    python scripts/make_golden_c5.py        # writes tests/golden/c5_tiled_oracle.npz (~1 min)
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bskit_b200 import synthetic as syn      # noqa: E402
from oracle import bskit_oracle as orc       # noqa: E402

N_SMALL, TILE, NBINS, SEED = 256, 8, 100, 5


def main():
    box = syn.BOX / TILE
    kmin, kmax, dk = syn.bench_bins(NBINS, box=box)
    mesh = syn.gaussian_mesh(N_SMALL, seed=SEED, box=box)
    edges = orc.bin_edges(kmin, kmax, dk)
    assert len(edges) == NBINS
    t0 = time.time()
    _, eq = orc.triangles_equilateral(edges)
    _, sq = orc.triangles_squeezed(edges, 0)
    b_eq = orc.measure_unnormalized([mesh], box, edges, eq, workers=8)
    b_sq = orc.measure_unnormalized([mesh], box, edges, sq, workers=8)
    out = os.path.join(ROOT, "tests", "golden", "c5_tiled_oracle.npz")
    np.savez(out, n_small=N_SMALL, tile=TILE, nbins=NBINS, seed=SEED, box_small=box, kmin=kmin, kmax=kmax, dk=dk,
             eq_idx=eq, sq_idx=sq, b_eq_small=b_eq, b_sq_small=b_sq)
    print(f"wrote {out}: {len(b_eq)} equilateral + {len(b_sq)} squeezed triangles in {time.time() - t0:.0f} s")


if __name__ == "__main__":
    main()
