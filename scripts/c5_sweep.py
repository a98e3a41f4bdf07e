#!/usr/bin/env python
"""Config C5 (BASELINE.json configs[4]): 2048^3 mesh, equilateral + squeezed-isosceles bins, x-slab
sharded over 1/2/4/8 GPUs; one JSON line per run (committed as profiles/r2_c5_n<N>.json).

    python scripts/c5_sweep.py                                            # 1 GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port 29511 scripts/c5_sweep.py

Every rank builds its own slab on the device: the 256^3 seeded Gaussian field of
scripts/make_golden_c5.py repeated 8 times along every axis (no 34 GB host mesh).  The result is
checked against the committed float64 oracle of the small field: B_big = 8^6 B_small for the same
physical bins (the "decimated oracle check"; identity tested on the CPU in
tests/test_oracle_bruteforce.py).  Times: CUDA events on the launching stream, max over ranks,
second of two runs per list (the first builds cuFFT plans and scratch).
"""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bskit_b200 as bk                         # noqa: E402
from bskit_b200 import synthetic as syn         # noqa: E402


def main():
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    lr = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(lr)
    dev = torch.device("cuda", lr)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    fx = np.load(os.path.join(ROOT, "tests", "golden", "c5_tiled_oracle.npz"))
    ns, tile, nb = int(fx["n_small"]), int(fx["tile"]), int(fx["nbins"])
    n = ns * tile
    if len(sys.argv) > 1:                       # smaller tiling factor for quick checks
        tile = int(sys.argv[1])
        n = ns * tile
    # bins used: the first `nb` of the fixture's 100 (default 36: k up to 292 k_f of the 2048^3 grid, the
    # "S up to 300" of SURVEY 8d; the cropped spectrum and one shell's scratch then fit beside three
    # 32 GiB fields on one GPU)
    nb = int(sys.argv[2]) if len(sys.argv) > 2 else 36
    small = syn.gaussian_mesh(ns, seed=int(fx["seed"]), box=float(fx["box_small"]))
    nxl = n // world
    x0 = rank * nxl
    sm = torch.from_numpy(small).to(dev)
    rows = (torch.arange(x0, x0 + nxl, device=dev) % ns)
    slab = sm[rows].repeat(1, tile, tile).contiguous()          # [nxl][n][n] float32
    del sm
    box = float(fx["box_small"]) * tile
    kmin, dk = float(fx["kmin"]), float(fx["dk"])
    kmax = kmin + (nb + 0.5) * dk

    def timed(fn):
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = fn()
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return out, float(t.item())

    res = {"config": "C5", "workload": f"{n}^3 float32 Gaussian mesh (256^3 seed {int(fx['seed'])} field tiled {tile}x per axis), "
                                       f"BoxSize {box:g}, S={nb} k-bins of width {tile} k_f, equilateral + squeezed(0) lists, auto-bispectrum",
           "nmesh": n, "nbins": nb, "n_gpus": world, "parallelism": f"x-slab x{world}", "grid": "full", "lists": {}}
    for tt, kw, key in (("equilateral", {}, "b_eq_small"), ("squeezed", dict(squeezed_bin_index=0), "b_sq_small")):
        for rep in range(2):
            fb = bk.FFTBispectrum(slab, Nmesh=n, BoxSize=box, kmin=kmin, kmax=kmax, dk=dk, triangle_type=tt,
                                  grid="full", device=dev, **kw)
            meas = fb._meas()
            edges, _ = fb._fast_bins()          # the engine the measurement will use: forward timed on its own
            eng = meas.session.engine(np.asarray(edges)[:, 1].max(), meas.precision)
            _, t_fwd = timed(lambda: meas.cubes(eng))
            got, t_meas = timed(lambda: fb.measure_bispectrum_faster(0, 10 ** 6))
            e = list(meas.session._engines.values())[0]
            info = {"forward_ms": t_fwd, "measure_ms": t_meas, "triangles": int(len(got["B"])), "batches": int(e.last_batches),
                    "schedule": e.last_schedule, "resident_fields": int(e.row_capacity())}
            b = np.asarray(got["B"])
            fb.close()
        want = np.asarray(fx[key])[:len(b)] * float(tile) ** 6
        assert np.array_equal(np.asarray(fx["eq_idx" if tt == "equilateral" else "sq_idx"])[:len(b)], np.asarray(fb.k_indices))
        rms = float(np.sqrt(np.mean(want ** 2)))
        info["vs_decimated_oracle"] = {"n": int(len(want)), "max_abs_err_over_rms": float(np.abs(b - want).max() / rms),
                                       "median_rel_err": float(np.median(np.abs(b - want) / np.abs(want))),
                                       "max_rel_err": float((np.abs(b - want) / np.abs(want)).max())}
        res["lists"][tt] = info
    res["total_ms"] = sum(v["forward_ms"] + v["measure_ms"] for v in res["lists"].values())
    if rank == 0:
        print(json.dumps(res))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
