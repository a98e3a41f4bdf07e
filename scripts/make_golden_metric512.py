#!/usr/bin/env python
"""Oracle values of the headline configuration, committed as a fixture.

Runs the CPU oracle (oracle/bskit_oracle.py: the reference's algorithm, bskit/main.py:1846-1882,
float64, full 512^3 grid, one masked inverse FFT per k-bin and one full-grid sum of I_a*I_b*I_c
per triangle) on the benchmark mesh (bskit_b200.synthetic.lognormal_mesh(512, seed=1), float32,
BoxSize 1000, S = 40 bins of width k_f from k_f/2) for ALL 6730 triangles and writes

    tests/golden/metric512_oracle.npz   {B, triples, edges, nmesh, seed, box}

The GPU parity test (tests/test_gpu_parity.py::test_metric_config_512_all_triangles_vs_oracle) and
bench.py's `checks.vs_oracle` compare the CUDA path with these values; the test also recomputes a
few entries with the oracle on the spot.  Needs ~50 GB of host memory and ~20 min on 8 cores.

    python scripts/make_golden_metric512.py [--nmesh 512] [--nbins 40] [--out tests/golden/...]
"""
import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import bskit_oracle as orc          # noqa: E402
from bskit_b200 import synthetic as syn          # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--nmesh", type=int, default=512)
    ap.add_argument("--nbins", type=int, default=40)
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--workers", type=int, default=len(os.sched_getaffinity(0)))
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    out = a.out or os.path.join(ROOT, "tests", "golden", f"metric{a.nmesh}_oracle.npz")
    t0 = time.time()
    mesh = syn.lognormal_mesh(a.nmesh, seed=a.seed, workers=a.workers)
    kmin, kmax, dk = syn.bench_bins(a.nbins)
    edges = orc.bin_edges(kmin, kmax, dk)
    _, triples = orc.triangles_all(edges, 1)
    print(f"mesh {mesh.shape} {mesh.dtype}, {len(edges)} bins, {len(triples)} triangles "
          f"({time.time() - t0:.0f}s)", flush=True)
    B = orc.measure_unnormalized_dense([mesh], syn.BOX, edges, triples, workers=a.workers, progress=True)
    np.savez_compressed(out, B=B, triples=triples.astype(np.int16), edges=edges, nmesh=a.nmesh,
                        seed=a.seed, box=syn.BOX)
    print(f"wrote {out} ({time.time() - t0:.0f}s); rms(B) = {np.sqrt(np.mean(B ** 2)):.6e}", flush=True)


if __name__ == "__main__":
    main()
