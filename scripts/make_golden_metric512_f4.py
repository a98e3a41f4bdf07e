#!/usr/bin/env python
"""What the reference's own float32 arithmetic gives on the headline configuration.

The reference keeps f4 meshes in f4: delta_k is complex64, every shell is a
complex64 -> float32 c2r, and the triangle sum is numpy's float32 pairwise ``np.sum(a*b*c)``
(bskit/main.py:1846-1879).  This script runs exactly that arithmetic (scipy pocketfft in single
precision standing in for PFFT/FFTW) for the 300 triangles with the smallest |B| and 300 others
spread over the list, and stores the values next to the float64 oracle fixture:

    tests/golden/metric512_ref_f4.npz   {index, B_f4}

The parity test uses it to show how far the reference's f4 path itself is from the float64
oracle on cancellation-dominated triangles (where no float32 pipeline holds 1e-5 relative).
"""
import os
import sys
import time

import numpy as np
import scipy.fft as sfft

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import bskit_oracle as orc          # noqa: E402
from bskit_b200 import synthetic as syn          # noqa: E402


def main():
    fx = np.load(os.path.join(ROOT, "tests", "golden", "metric512_oracle.npz"))
    n, want, triples, edges = int(fx["nmesh"]), fx["B"], fx["triples"].astype(int), fx["edges"]
    workers = len(os.sched_getaffinity(0))
    order = np.argsort(np.abs(want))
    pick = np.unique(np.concatenate([order[:300], order[300::max(1, (len(order) - 300) // 300)]]))
    mesh = syn.lognormal_mesh(n, seed=int(fx["seed"]), workers=workers)          # float32
    dk32 = (sfft.rfftn(mesh, workers=workers) / np.float32(mesh.size)).astype(np.complex64)
    kk = orc.k_norm(n, syn.BOX)
    shells = {}
    t0 = time.time()
    for b in np.unique(triples[pick]):
        m = (dk32 * orc.shell_mask(kk, edges[b, 0], edges[b, 1])).astype(np.complex64)
        shells[int(b)] = (sfft.irfftn(m, s=(n, n, n), workers=workers) * np.float32(n) ** 3).astype(np.float32)
    print(f"{len(shells)} float32 shells ({time.time() - t0:.0f}s)", flush=True)
    norm = syn.BOX ** 6 / float(n) ** 3
    out = np.empty(len(pick))
    for i, t in enumerate(pick):
        a, b, c = triples[t]
        out[i] = float(np.sum(shells[a] * shells[b] * shells[c])) * norm      # float32 product and sum
        if i % 100 == 0:
            print(f"  {i}/{len(pick)} {time.time() - t0:.0f}s", flush=True)
    rel = np.abs(out - want[pick]) / np.abs(want[pick])
    print("f4 reference arithmetic vs f64 oracle: median %.2e, 99%% %.2e, max %.2e, n > 1e-5: %d of %d"
          % (np.median(rel), np.quantile(rel, 0.99), rel.max(), int((rel > 1e-5).sum()), len(pick)))
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "metric512_ref_f4.npz"), index=pick.astype(np.int32),
                        B_f4=out)


if __name__ == "__main__":
    main()
