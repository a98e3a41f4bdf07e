#!/usr/bin/env python
"""Per-triangle error of the CUDA paths against the committed 512^3 oracle fixture
(tests/golden/metric512_oracle.npz): distribution of |B_gpu - B_oracle| / |B_oracle| with no floor."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import bskit_b200 as bk                       # noqa: E402
from bskit_b200 import synthetic as syn       # noqa: E402


def main():
    fx = np.load(os.path.join(ROOT, "tests", "golden", "metric512_oracle.npz"))
    want = fx["B"]
    n, nb = int(fx["nmesh"]), len(fx["edges"])
    mesh = syn.lognormal_mesh(n, seed=int(fx["seed"]), workers=os.cpu_count())
    kmin, kmax, dk = syn.bench_bins(nb)
    rms = np.sqrt(np.mean(want ** 2))
    out = {}
    for name, kw in (("tensor_full", dict(grid="full", contraction="tensor")),
                     ("fp32_full", dict(grid="full", contraction="fp32")),
                     ("f64acc_full", dict(grid="full", accum_dtype=np.float64)),
                     ("tensor_auto", dict(grid="auto", contraction="tensor")),
                     ("fp32_auto", dict(grid="auto", contraction="fp32"))):
        fb = bk.FFTBispectrum(mesh, BoxSize=syn.BOX, kmin=kmin, kmax=kmax, dk=dk, **kw)
        got = fb.measure_bispectrum_faster()["B"]
        path = fb.attrs.get("contraction_path")
        fb.close()
        rel = np.abs(got - want) / np.abs(want)
        absr = np.abs(got - want) / rms
        q = np.quantile(rel, [0.5, 0.9, 0.99, 0.999, 1.0])
        small = np.argsort(np.abs(want))[: len(want) // 10]
        out[name] = dict(path=path, rel_q50=q[0], rel_q90=q[1], rel_q99=q[2], rel_q999=q[3], rel_max=q[4],
                         frac_gt_1e5=float(np.mean(rel > 1e-5)), n_gt_1e5=int(np.sum(rel > 1e-5)),
                         abs_over_rms_max=float(absr.max()),
                         smallest_decile_rel_max=float(rel[small].max()),
                         smallest_decile_frac_gt_1e5=float(np.mean(rel[small] > 1e-5)),
                         smallest_abs_over_rms=float(np.abs(want[small]).max() / rms),
                         signed_mean_rel=float(np.mean((got - want) / want)))
        print(name, json.dumps(out[name]), flush=True)
    bk.clear_cache()
    print(json.dumps(out))


if __name__ == "__main__":
    main()
