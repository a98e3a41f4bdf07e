#!/usr/bin/env python
"""CPU model of the arithmetic of the tensor-core contraction (contract_tc.cuh), run on the headline
configuration (512^3 seeded lognormal mesh, S = 40) for a uniform sample of the 6730 triangles.

Why: the per-triangle error distribution of the tcgen05 path was measured on the GPU for two schedules
("two smallest rows" and "class cover", profiles/r2_parity512.txt, r2d_bench_n1.json); the final
"two halves" schedule generates other pairs ((b,c) with column a for triangles whose two largest rows lie
in the upper half) and was timed / profiled / checked on synthetic fields only before the round's GPU
budget ended.  This model predicts its distribution and is validated against the two measured ones.

What is modelled (profiles/r1_tensor_core_rounding_probe.txt):
  * shell fields: float64 inverse FFT, stored as float32 (the pipeline's storage);
  * pair product P = fl32(F_x F_y) (round to nearest), P_hi = tf32(P) rounded to nearest, P_lo = fl32(P - P_hi)
    truncated to tf32 by the tensor core; C_hi / C_lo likewise;
  * per 8 cells three MMAs (P_lo C_hi, P_hi C_lo, P_hi C_hi), each accumulating two groups of 4 exact
    products into an fp32 accumulator that is TRUNCATED after every group (the probe: "chunk 4, RZ");
  * one accumulator lives for a 128-cell window, is then added (fp32, round to nearest) to a register
    accumulator, which is flushed to float64 every 128 windows.
Not modelled: the order in which CTAs visit tiles (float64 stage, irrelevant at 1e-16).

Output: one JSON line per pair rule with the same statistics bench.py prints in checks.vs_oracle.
Usage: python scripts/dev/tc_precision_model.py [n_sample=240] [n_procs=6] [variants]   (~35 GB of host memory)
"variants": instead of the three pair rules, design variants of the accumulation (truncation once per MMA, 64-cell
windows, the two small split terms in a second accumulator) with the two-halves pairs.
"""
import json
import multiprocessing as mp
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import bskit_oracle as orc          # noqa: E402
from bskit_b200 import synthetic as syn          # noqa: E402

WINDOW, GROUP, FLUSH = 128, 4, 128
SLAB = 1 << 23                                   # cells per slab of the emulation (65536 windows)
_SHELLS = {}
VARIANT_RUN = len(sys.argv) > 3 and sys.argv[3] == "variants"
HALF = 20                                        # the cut of build_schedule_halves for 40 rows (set in main)


def tf32_rn(x):
    """float32 -> nearest value with 10 explicit mantissa bits (ties to even), as float32."""
    u = x.view(np.uint32)
    r = (u + np.uint32(0xFFF) + ((u >> np.uint32(13)) & np.uint32(1))) & np.uint32(0xFFFFE000)
    return r.view(np.float32)


def tf32_rz(x):
    return (x.view(np.uint32) & np.uint32(0xFFFFE000)).view(np.float32)


def rz32(x):
    """float64 -> float32 rounded toward zero."""
    f = x.astype(np.float32)
    over = np.abs(f.astype(np.float64)) > np.abs(x)
    if over.any():
        f[over] = np.nextafter(f[over], np.float32(0))
    return f


def emulate(fx, fy, fc, group=GROUP, window=WINDOW, split_small=False):
    """Tensor-core sum of fx*fy*fc (float32 arrays of equal length, a multiple of window*FLUSH).
    group: products summed exactly between truncations of the accumulator (4: the probe's reading; 8: once per
    MMA); window: cells an accumulator lives for; split_small: the two small terms of the split go to a second
    accumulator (design variant)."""
    total = 0.0
    for s0 in range(0, len(fx), SLAB):
        x, y, c = fx[s0:s0 + SLAB], fy[s0:s0 + SLAB], fc[s0:s0 + SLAB]
        p = x * y                                  # float32, round to nearest
        p_hi = tf32_rn(p)
        p_lo = tf32_rz(p - p_hi)
        c_hi = tf32_rn(c)
        c_lo = tf32_rz(c - c_hi)
        ph, pl = p_hi.astype(np.float64), p_lo.astype(np.float64)
        ch, cl = c_hi.astype(np.float64), c_lo.astype(np.float64)
        nwin = len(x) // window
        # exact sums of `group` products: [term][window][k-step][part]
        terms = [(a * b).reshape(nwin, window // 8, 8 // group, group).sum(axis=3)
                 for a, b in ((pl, ch), (ph, cl), (ph, ch))]
        acc = np.zeros(nwin, dtype=np.float32)
        small = np.zeros(nwin, dtype=np.float32)
        for ks in range(window // 8):
            for it, t in enumerate(terms):
                for part in range(8 // group):
                    if split_small and it < 2:
                        small = rz32(small.astype(np.float64) + t[:, ks, part])
                    else:
                        acc = rz32(acc.astype(np.float64) + t[:, ks, part])
        if split_small:
            acc = acc + small                      # drained separately, added in fp32 (round to nearest)
        # register accumulators (fp32 round-to-nearest adds), flushed to float64 every FLUSH windows
        w = acc.reshape(-1, FLUSH)
        reg = np.zeros(w.shape[0], dtype=np.float32)
        for j in range(FLUSH):
            reg = reg + w[:, j]
        total += float(reg.astype(np.float64).sum())
    return total


VARIANTS = {"group8": dict(group=8), "window64": dict(window=64), "split_small": dict(split_small=True),
            "group8_split_small": dict(group=8, split_small=True)}


def work(args):
    t, rows = args
    r = sorted(int(v) for v in rows)
    f = [_SHELLS[v] for v in r]
    exact = 0.0
    for s0 in range(0, len(f[0]), SLAB):
        exact += float((f[0][s0:s0 + SLAB].astype(np.float64) * f[1][s0:s0 + SLAB] * f[2][s0:s0 + SLAB]).sum())
    if VARIANT_RUN:
        # design variants, with the pair the two-halves schedule generates for this triangle
        a, b, c = (f[0], f[1], f[2]) if r[1] < HALF else (f[1], f[2], f[0])
        return t, r, exact, {k: emulate(a, b, c, **kw) for k, kw in VARIANTS.items()}, None
    low = emulate(f[0], f[1], f[2])                # pair = two smallest rows, column = largest
    high = emulate(f[1], f[2], f[0])               # pair = two largest rows, column = smallest
    return t, r, exact, low, high


def stats(got, want):
    rel = np.abs(got - want) / np.abs(want)
    return {"n": int(len(rel)), "median": float(np.median(rel)), "q90": float(np.quantile(rel, 0.9)),
            "q99": float(np.quantile(rel, 0.99)), "max": float(rel.max()),
            "n_above_1e-5": int((rel > 1e-5).sum()), "frac_above_1e-5": float((rel > 1e-5).mean()),
            "mean_signed": float(np.mean((got - want) / want))}


def main():
    n_sample = int(sys.argv[1]) if len(sys.argv) > 1 else 240
    n_procs = int(sys.argv[2]) if len(sys.argv) > 2 else 6
    fx = np.load(os.path.join(ROOT, "tests", "golden", "metric512_oracle.npz"))
    n, want, triples, edges = int(fx["nmesh"]), fx["B"], fx["triples"].astype(int), fx["edges"]
    workers = len(os.sched_getaffinity(0))
    pick = np.unique(np.linspace(0, len(want) - 1, n_sample).astype(int))
    t0 = time.time()
    mesh = syn.lognormal_mesh(n, seed=int(fx["seed"]), workers=workers)
    dk = orc.forward(mesh.astype(np.float64), workers=workers)
    kk = orc.k_norm(n, syn.BOX)
    for b in np.unique(triples[pick]):
        _SHELLS[int(b)] = orc.data_shell(dk, kk, edges[b, 0], edges[b, 1], workers=workers).astype(np.float32).ravel()
    del dk, kk, mesh
    print(f"{len(_SHELLS)} shells, {time.time() - t0:.0f}s", flush=True)
    norm = syn.BOX ** 6 / float(n) ** 3
    rms = float(np.sqrt(np.mean(want ** 2)))
    global HALF
    h = HALF = (((len(edges) + 7) // 8 * 8) // 2) & ~1  # the cut of build_schedule_halves (tc_schedule.h)
    res = {}
    with mp.get_context("fork").Pool(n_procs) as pool:
        for i, (t, r, exact, low, high) in enumerate(pool.imap_unordered(work, [(int(t), triples[t]) for t in pick])):
            if VARIANT_RUN:
                res[t] = (r, exact * norm, {k: v * norm for k, v in low.items()}, None)
            else:
                res[t] = (r, exact * norm, low * norm, high * norm)
            if i % 20 == 0:
                print(f"  {i}/{len(pick)} {time.time() - t0:.0f}s", flush=True)
    ts = np.array(sorted(res))
    exact = np.array([res[t][1] for t in ts])
    if VARIANT_RUN:
        out = {"sample": f"{len(ts)} triangles, uniformly spaced over the list of {len(want)}; pairs as the two-halves "
                         "schedule generates them"}
        for k in VARIANTS:
            got = np.array([res[t][2][k] for t in ts]) * 1.0
            out[k] = dict(stats(got, want[ts]), max_abs_over_rms=float(np.abs(got - want[ts]).max() / rms),
                          noise_abs_over_rms_q50=float(np.median(np.abs(got - exact)) / rms),
                          noise_abs_over_rms_max=float(np.abs(got - exact).max() / rms))
        out["seconds"] = time.time() - t0
        print(json.dumps(out, indent=1))
        return
    low = np.array([res[t][2] for t in ts])
    high = np.array([res[t][3] for t in ts])
    halves = np.array([res[t][2] if res[t][0][1] < h else res[t][3] for t in ts])
    out = {"sample": f"{len(ts)} triangles, uniformly spaced over the list of {len(want)}; rms(B) = {rms:.6e}",
           "float32_storage_only_vs_oracle": stats(exact, want[ts]),
           "rule_two_smallest_rows_vs_oracle": dict(stats(low, want[ts]),
                                                    max_abs_over_rms=float(np.abs(low - want[ts]).max() / rms)),
           "rule_two_largest_rows_vs_oracle": dict(stats(high, want[ts]),
                                                   max_abs_over_rms=float(np.abs(high - want[ts]).max() / rms)),
           "rule_two_halves_vs_oracle": dict(stats(halves, want[ts]),
                                             max_abs_over_rms=float(np.abs(halves - want[ts]).max() / rms),
                                             team1_fraction=float(np.mean([res[t][0][1] >= h for t in ts]))),
           "noise_abs_over_rms_two_halves": {
               "q50": float(np.median(np.abs(halves - exact)) / rms), "q99": float(np.quantile(np.abs(halves - exact), 0.99) / rms),
               "max": float(np.abs(halves - exact).max() / rms)},
           "seconds": time.time() - t0}
    print(json.dumps(out, indent=1))
    np.savez_compressed("/tmp/tc_precision_model.npz", index=ts, exact=exact, low=low, high=high, halves=halves)


if __name__ == "__main__":
    main()
