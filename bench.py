#!/usr/bin/env python
"""Benchmark of the FFT-bispectrum hot path (BASELINE.json: "s per 512^3 all-triangle
bispectrum at 1/2/4/8 B200").

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N ...             # the reference algorithm on host cores

Workload ("metric" row): 512^3 float32 lognormal mesh, BoxSize 1000,
S = 40 k-bins of width k_f from k_f/2, all 6730 closed triangles, auto-bispectrum plus
normalisation (N_tri and the three k-means).  One step = forward transform of the resident
mesh, synthesis of the 40 shell fields, the triangle contraction, and the normalisation.
Strong scaling: the same mesh is x-slab sharded over N ranks.

The headline `value` runs every stage on the mesh's own 512^3 grid (grid='full'), i.e. the
transforms and cell sums the reference performs.  The exact band-limited evaluation
(grid='auto', the library default) is reported beside it under "auto_grid".

`checks` (outside the timed region): every result of the timed steps is compared with the
committed float64 oracle fixture of this exact configuration (tests/golden/metric512_oracle.npz,
all 6730 triangles, produced by oracle/ on the same seeded mesh), the normalisation with an oracle
run on the spot, and `checksum` makes results comparable across GPU counts.
"""
from __future__ import annotations

import argparse
import hashlib
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "s per 512^3 all-triangle bispectrum"
GOLDEN_GRIDINFO = os.path.join(ROOT, "tests", "golden", "reference_output_ref", "Lbox1000_512_kf_3kf_3lowkbins.dat")
FIXTURE = os.path.join(ROOT, "tests", "golden", "metric512_oracle.npz")
NCU_SUMMARY = os.path.join(ROOT, "profiles", "r2_tc_contract_ncu.json")     # committed per round (scripts/dev/ncu_summary.py)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--nmesh", type=int, default=512)
    ap.add_argument("--nbins", type=int, default=40)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--accum", default="f32", choices=["f32", "f64"])
    ap.add_argument("--contraction", default=None, choices=["tensor", "fp32"],
                    help="contraction kernel of the headline step (default: the library default)")
    ap.add_argument("--mesh", default="host", choices=["host", "device"],
                    help="device: every rank generates its own slab on the GPU (no host mesh; no e2e leg)")
    ap.add_argument("--oracle-shells", default="",
                    help="comma separated k-bin indices: rank 0 gathers the mesh to the host, builds these shells with "
                         "the float64 oracle on the full grid and checks every closed triangle among them (large grids)")
    ap.add_argument("--scheme", default="kf", choices=["kf", "paper80"],
                    help="paper80: the reference's production binning (examples/batch/sub_measure_bs_faster_ill.sh:33-37): "
                         "40 bins of width k_f from k_f/2, then bins of width 6 k_f up to 280.51 k_f (80 bins, 24138 triangles)")
    ap.add_argument("--profile", action="store_true",
                    help="only the full-grid device-resident steps (for ncu); prints stage times")
    return ap.parse_args()


# --------------------------------------------------------------------------- #
# CPU leg: the reference algorithm (oracle port) on the host cores
# --------------------------------------------------------------------------- #
def oracle_golden_gate():
    """BASELINE.md section 3 validity gate: before it is timed, the oracle port must reproduce the
    reference's golden grid-info file (59 rows, every printed digit; 64^3 suffices)."""
    from oracle import bskit_oracle as orc
    g = np.loadtxt(GOLDEN_GRIDINFO)
    e = orc.bin_edges(kmin=0.00314, kmax=0.1, dk=0.00628, num_lowk_bins=3, dk_high=0.01884)
    _, idx = orc.triangles_all(e, 1)
    ntri, kmean = orc.measure_gridinfo(64, 1000.0, e, idx, workers=host_cores())
    ok = len(g) == len(idx) == 59
    for t in range(59):
        got = ["%e" % kmean[t, 0], "%e" % kmean[t, 1], "%e" % kmean[t, 2], "%e" % ntri[t]]
        want = ["%e" % g[t, 1], "%e" % g[t, 2], "%e" % g[t, 3], "%e" % g[t, 10]]
        ok = ok and got == want
    if not ok:
        raise RuntimeError("the oracle port does not reproduce the reference's golden grid-info file")
    return "59/59 golden rows reproduced"


def cpu_reference_sample(mesh32, nbins, ntri, cores, shell_ids, n_tri_sample=64, n_field64=2, n_tri64=4):
    """Time a bounded sample of the reference's algorithm (bskit/main.py:1846-1879 and
    2006-2061) and extrapolate linearly to the whole job: S masked inverse FFTs (f4, the
    mesh dtype) + T full-grid triple-product sums, then 2S f8 fields + 4T f8 sums.
    `shell_ids`: which of the S f4 shells are really built (all of them
    for the reported baseline), n_tri_sample >= 64 triangle sums."""
    import scipy.fft as sfft
    from concurrent.futures import ThreadPoolExecutor
    from oracle import bskit_oracle as orc
    from bskit_b200 import synthetic as syn

    n = mesh32.shape[0]
    kmin, kmax, dk = syn.bench_bins(nbins)
    edges = orc.bin_edges(kmin, kmax, dk)
    pool = ThreadPoolExecutor(cores)

    def tri_sum(a, b, c):
        sl = np.array_split(np.arange(n), cores)
        parts = pool.map(lambda s: np.sum(a[s[0]:s[-1] + 1] * b[s[0]:s[-1] + 1] * c[s[0]:s[-1] + 1]), sl)
        return sum(parts)

    t0 = time.perf_counter()
    dk32 = (sfft.rfftn(mesh32, workers=cores) / mesh32.size).astype(np.complex64)
    t_fwd = time.perf_counter() - t0
    t0 = time.perf_counter()
    keep = []
    for i in shell_ids:
        kk = orc.k_norm(n, syn.BOX)                    # the reference recomputes |k| per bin
        m = dk32 * orc.shell_mask(kk, edges[i, 0], edges[i, 1])
        sh = (sfft.irfftn(m, s=(n, n, n), workers=cores) * np.float32(n) ** 3).astype(np.float32)
        if len(keep) < 3:
            keep.append(sh)                            # three resident shells feed the timed sums
        del kk, m, sh
    t_shell = (time.perf_counter() - t0) / len(shell_ids)
    t0 = time.perf_counter()
    for j in range(n_tri_sample):
        tri_sum(keep[j % len(keep)], keep[(j + 1) % len(keep)], keep[(2 * j) % len(keep)])
    t_tri = (time.perf_counter() - t0) / n_tri_sample
    del keep
    # normalisation leg: float64 number / k fields (always f8 in the reference)
    t0 = time.perf_counter()
    kk = orc.k_norm(n, syn.BOX)
    i0 = shell_ids[0]
    f64 = []
    for j in range(n_field64):
        f64.append(orc.number_field(n, syn.BOX, edges[i0, 0], edges[i0, 1], kk, cores) if j % 2 == 0
                   else orc.k_field(n, syn.BOX, edges[i0, 0], edges[i0, 1], 1.0, kk, cores))
    t_nfield = (time.perf_counter() - t0) / n_field64
    t0 = time.perf_counter()
    for j in range(n_tri64):
        tri_sum(f64[0], f64[-1], f64[0])
    t_tri64 = (time.perf_counter() - t0) / n_tri64
    pool.shutdown()
    total = t_fwd + nbins * t_shell + ntri * t_tri + 2 * nbins * t_nfield + 4 * ntri * t_tri64
    sample = (f"oracle port (numpy/scipy restatement, not nbodykit): timed 1 forward rfftn, "
              f"{len(shell_ids)} of {nbins} f4 shell builds, {n_tri_sample} of {ntri} f4 triangle sums, "
              f"{n_field64} of {2 * nbins} f8 number/k fields, {n_tri64} of {4 * ntri} f8 sums on {n}^3; "
              f"extrapolated linearly (t_fwd={t_fwd:.2f}s t_shell={t_shell:.2f}s t_tri={t_tri:.3f}s "
              f"t_field64={t_nfield:.2f}s t_tri64={t_tri64:.3f}s)")
    return total, sample


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


# --------------------------------------------------------------------------- #
# clocks sampler
# --------------------------------------------------------------------------- #
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self._stop = index, [], threading.Event()
        self.thread = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                      "--format=csv,noheader,nounits"], capture_output=True,
                                     text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            self._stop.wait(0.2)

    def __enter__(self):
        self.thread.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self.thread.join(timeout=6)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        sm = sorted(float(r[0]) for r in self.rows if r[0].replace(".", "").isdigit())
        reasons = []
        for col, name in ((2, "hw_slowdown"), (3, "hw_thermal_slowdown"), (4, "sw_thermal_slowdown"),
                          (5, "sw_power_cap")):
            if any(len(r) > col and r[col].lower().startswith("active") for r in self.rows):
                reasons.append(name)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None,
                "sm_max_mhz": float(self.rows[0][1]) if self.rows[0][1].replace(".", "").isdigit() else None,
                "reasons": reasons, "samples": len(self.rows)}


# --------------------------------------------------------------------------- #
# checks
# --------------------------------------------------------------------------- #
def rel_error_stats(got, want):
    """Per-triangle relative error distribution, no floor."""
    got, want = np.asarray(got, dtype=np.float64), np.asarray(want, dtype=np.float64)
    rel = np.abs(got - want) / np.abs(want)
    rms = float(np.sqrt(np.mean(want ** 2)))
    bad = rel > 1e-5
    return {"n": int(len(rel)), "median": float(np.median(rel)), "q99": float(np.quantile(rel, 0.99)),
            "max": float(rel.max()), "n_above_1e-5": int(bad.sum()), "frac_above_1e-5": float(bad.mean()),
            "max_abs_over_rms": float(np.abs(got - want).max() / rms),
            "largest_abs_B_over_rms_among_those_above_1e-5": float(np.abs(want[bad]).max() / rms) if bad.any() else 0.0,
            "mean_signed": float(np.mean((got - want) / want))}


def checksum(b, ntri, kmean):
    """Numbers that must agree between runs on different GPU counts (the float64 partial sums are
    all-reduced in a different order and the float32 tile sums cut the cells differently, so B agrees to
    ~1e-9 relative; N_tri exactly)."""
    b, ntri, kmean = np.asarray(b), np.asarray(ntri), np.asarray(kmean)
    ok = np.isfinite(kmean)
    txt = ",".join("%d" % v for v in ntri)          # integers: must be identical on every GPU count
    return {"B_sum": float(b.sum()), "B_abs_sum": float(np.abs(b).sum()), "N_tri_sum": int(ntri.sum()),
            "k_mean_sum": float(kmean[ok].sum()), "digest_Ntri": hashlib.sha1(txt.encode()).hexdigest()[:16]}


def measure_tf32_peak(dev):
    """cuBLAS TF32 GEMM 8192^3, best of 10 (the way MEASURED_PEAKS.json measures bf16), outside timing."""
    import torch
    old = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = True
    try:
        a = torch.randn(8192, 8192, device=dev)
        b = torch.randn(8192, 8192, device=dev)
        for _ in range(3):
            a @ b
        best = 1e9
        for _ in range(10):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            a @ b
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        del a, b
        return 2.0 * 8192 ** 3 / (best * 1e-3) / 1e12
    finally:
        torch.backends.cuda.matmul.allow_tf32 = old
        torch.cuda.empty_cache()


# --------------------------------------------------------------------------- #
def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    nmesh, nbins = args.nmesh, args.nbins

    from bskit_b200 import synthetic as syn
    import bskit_b200 as bk
    if args.scheme == "paper80":
        kf = 2.0 * np.pi / syn.BOX
        kmin, kmax, dk = 0.5 * kf, 280.51 * kf, kf
        edges = bk.generate_bin_edge_list(kmin, kmax, dk, 40, 6.0 * kf)
        triples = bk.generate_triangle_bin_list(kmin, kmax, dk, num_lowk_bins=40, dk_high=6.0 * kf, return_indices=True)
        binning = "k-bins: 40 of width k_f from k_f/2, then width 6 k_f up to 280.5 k_f (the reference's production scheme)"
    else:
        kmin, kmax, dk = syn.bench_bins(nbins)
        edges = bk.generate_bin_edge_list(kmin, kmax, dk)
        triples = bk.generate_triangle_bin_list(kmin, kmax, dk, return_indices=True)
        binning = "k-bins of width k_f from k_f/2"
    ntri = len(triples)
    cores = host_cores()
    config = {"workload": f"{nmesh}^3 float32 lognormal mesh (seed 1, BoxSize 1000), S={len(edges)} "
                          f"{binning}, all {ntri} triangles, auto + normalisation",
              "nmesh": nmesh, "nbins": int(len(edges)), "ntriangles": int(ntri),
              "parallelism": f"x-slab x{max(world, 1)}", "l2": "inputs larger than L2 (no flush needed)"}

    # ---------------------------------------------------------------- reference arm
    if args.impl == "reference":
        if rank != 0:
            return
        gate = oracle_golden_gate()
        mesh = syn.lognormal_mesh(nmesh, seed=1, workers=cores)
        vals = []
        samples = []
        S = len(edges)
        # the K timed steps together build every one of the S shells once; each step times 64 sums
        per_step = max(1, -(-S // max(args.steps, 1)))
        for i in range(args.warmup + args.steps):
            if i < args.warmup:
                ids, nt = [S // 2], 8
            else:
                k = i - args.warmup
                ids = [(k * per_step + j) % S for j in range(per_step)]
                nt = 64
            v, sample = cpu_reference_sample(mesh, S, ntri, cores, ids, n_tri_sample=nt, n_field64=2, n_tri64=4)
            if i >= args.warmup:
                vals.append(v)
                samples.append(sample)
        v = float(np.median(vals))
        sample = (f"validity gate: {gate}; {args.steps} timed steps, each a bounded sample extrapolated linearly "
                  f"(step k builds shells k*{per_step}..k*{per_step}+{per_step - 1} mod {S}: all {S} shells over the run; "
                  f"64 triangle sums per step); median reported; last step: " + samples[-1])
        print(json.dumps({
            "impl": "reference", "metric": METRIC, "value": v, "unit": "s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": v * 1e3,
            "higher_is_better": False, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": config,
            "cpu_baseline": {"value": v, "unit": "s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": v, "unit": "s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}))
        return

    # ---------------------------------------------------------------- our arm
    import torch
    import torch.distributed as dist
    from bskit_b200 import engine as eng, _native as nat

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    assert world == args.gpus or world == 1, "launch with torchrun --nproc-per-node N for --gpus N"

    host = None
    if args.mesh == "host":
        # synthetic mesh: rank 0 generates, everyone receives (host copy kept pinned for the e2e leg)
        host = torch.empty((nmesh, nmesh, nmesh), dtype=torch.float32).pin_memory()
        if rank == 0:
            host.copy_(torch.from_numpy(syn.lognormal_mesh(nmesh, seed=1, workers=cores)))
        if world > 1:
            tmp = host.to(dev)
            dist.broadcast(tmp, 0)
            host.copy_(tmp.cpu())
            del tmp
    accum = nat.F32 if args.accum == "f32" else nat.F64

    def device_slab(e_data):
        """This rank's x-planes of a lognormal-like field generated on the GPU (large grids: no host mesh)."""
        f = e_data.info
        # seeded per block of nmesh/8 planes, so the mesh does not depend on the number of ranks
        blk = max(1, nmesh // 8)
        parts = []
        for x0 in range(int(f.nx0), int(f.nx0 + f.nxl), blk):
            gen = torch.Generator(device=dev)
            gen.manual_seed(1234 + x0)
            parts.append(torch.randn((min(blk, int(f.nx0 + f.nxl) - x0), nmesh, nmesh), generator=gen, device=dev,
                                     dtype=torch.float32))
        g = torch.cat(parts) if len(parts) > 1 else parts[0]
        del parts
        # smooth along z and y (cheap separable box filter) so that low-k shells carry signal, then lognormal
        for dim in (1, 2):
            g = (g + torch.roll(g, 1, dim) + torch.roll(g, -1, dim)) / 3.0
        g = g * (0.8 / 0.58)
        return (torch.exp(g - 0.32) - 1.0).contiguous()

    def make(policy, contraction=None):
        g = eng.choose_grid(nmesh, syn.BOX, edges[:, 1].max(), policy, world)
        gn = eng.choose_grid(nmesh, syn.BOX, edges[:, 1].max(), "auto", world)
        e_data = eng.Engine(g, syn.BOX, nat.F32, device=dev, accum_precision=accum, contraction=contraction)
        e_norm = eng.Engine(gn, syn.BOX, nat.F64, device=dev)
        return e_data, e_norm

    vol2 = syn.BOX ** 6

    def step(e_data, e_norm, slab, marks=None):
        eng._mark(marks, "start", e_data)
        cube = e_data.forward(slab)
        eng._mark(marks, "forward_done", e_data)
        # both measurements are enqueued before the first result is fetched: the host work of the
        # (mesh independent) normalisation hides behind the contraction kernel
        fetch_b = eng.measure_triangle_sums(e_data, [cube], edges, triples, marks=marks, defer=True)
        eng._mark(marks, "data_done", e_data)
        fetch_n = eng.measure_grid_sums(e_norm, edges, triples, defer=True)
        eng._mark(marks, "norm_done", e_data)
        b = fetch_b() * vol2
        ntri_v, kmean = fetch_n()
        return b, ntri_v, kmean

    kept = {}

    def timed(policy, steps, warmup, with_clocks=False, contraction=None, keep_slab=False):
        e_data, e_norm = make(policy, contraction)
        slab = e_data.local_slab(host) if host is not None else device_slab(e_data)
        if keep_slab:
            kept["slab"] = slab.cpu()
        for _ in range(warmup):
            out = step(e_data, e_norm, slab)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        launches0 = nat.launch_count()
        all_marks = []
        sampler = ClockSampler(local_rank) if with_clocks else None
        if sampler:
            sampler.__enter__()
        t0 = torch.cuda.Event(enable_timing=True)
        t1 = torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        t0.record()
        for _ in range(steps):
            marks = []
            out = step(e_data, e_norm, slab, marks)
            all_marks.append(marks)
        t1.record()
        torch.cuda.synchronize()
        if sampler:
            sampler.__exit__()
        if world > 1:
            dist.barrier()
        ms = t0.elapsed_time(t1)
        launches = nat.launch_count() - launches0
        if world > 1:
            tms = torch.tensor([ms], dtype=torch.float64, device=dev)
            dist.all_reduce(tms, op=dist.ReduceOp.MAX)
            ms = float(tms.item())
        stages = {}
        for marks in all_marks:
            d = dict(marks)
            for a, b_, name in (("start", "forward_done", "forward"), ("forward_done", "shells_done", "shells"),
                                ("shells_done", "contract_done", "contract"),
                                ("data_done", "norm_done", "normalisation")):
                stages.setdefault(name, []).append(d[a].elapsed_time(d[b_]))
        stages = {k: float(np.mean(v)) for k, v in stages.items()}
        info = e_data.backend.cplan_info()
        res = dict(ms_per_step=ms / steps, stages_ms=stages, launches=launches, out=out,
                   grid=e_data.grid, ncells=e_data.ncells, cplan=info, schedule=e_data.last_schedule,
                   ntri_residual=getattr(e_norm, "last_ntri_residual", None),
                   clocks=sampler.summary() if sampler else None)
        e_data.close()
        e_norm.close()
        del slab
        torch.cuda.empty_cache()
        return res

    full = timed("full", args.steps, args.warmup, with_clocks=not args.profile, contraction=args.contraction,
                 keep_slab=bool(args.oracle_shells) and host is None)
    if args.profile:
        if rank == 0:
            print(json.dumps({"profile_only": True, "ms_per_step": full["ms_per_step"],
                              "stages_ms": full["stages_ms"], "gpu_launches": full["launches"]}))
        return
    auto = timed("auto", args.steps, args.warmup, contraction=args.contraction)
    # the same step with the other contraction kernel (tensor cores <-> FP32 pipe), reported beside the default
    default_tensor = full["schedule"] == "tensor"
    alt = timed("full", 2, 1, contraction="fp32" if default_tensor else "tensor") if accum == nat.F32 else None

    # ---- e2e: the user-facing API from (pinned) host memory, result back on the host.  The
    # process-wide grid-info cache is switched off: every step recomputes the normalisation.
    e2e_s = e2e_cold_s = None
    b_api = g_api = None
    n_e2e = 0
    if host is not None:
        old_cache = bk.set_gridinfo_cache(False)

        def e2e_once():
            fb = bk.FFTBispectrum(host, BoxSize=syn.BOX, kmin=kmin, kmax=kmax, dk=dk, grid="full",
                                  accum_dtype=np.float32 if accum == nat.F32 else np.float64,
                                  contraction=args.contraction, device=dev)
            g = fb.measure_gridinfo_faster(0, ntri)      # step 1 of the reference workflow; overlaps the upload
            b = fb.measure_bispectrum_faster(0, ntri)
            fb.close()
            return b, g

        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        c0 = time.perf_counter()
        e2e_once()                                    # cold: cuFFT plans, contraction schedules, scratch
        torch.cuda.synchronize()
        e2e_cold_s = time.perf_counter() - c0
        for _ in range(max(0, min(args.warmup, 2) - 1)):
            e2e_once()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        wall0 = time.perf_counter()
        n_e2e = max(1, min(args.steps, 3))
        for _ in range(n_e2e):
            b_api, g_api = e2e_once()
        torch.cuda.synchronize()
        e2e_s = (time.perf_counter() - wall0) / n_e2e       # host-side work (plan lookups, D2H) counts too
        if world > 1:
            t = torch.tensor([e2e_s, e2e_cold_s], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            e2e_s, e2e_cold_s = float(t[0].item()), float(t[1].item())
        bk.set_gridinfo_cache(old_cache)

    oracle_sample = None
    if args.oracle_shells:
        bins = sorted(int(x) for x in args.oracle_shells.split(","))
        if host is not None:
            mesh_np = host.numpy() if rank == 0 else None
        else:
            # every rank drops its slab into shared memory on the host, rank 0 assembles the mesh
            shm = "/dev/shm" if os.path.isdir("/dev/shm") else "/tmp"
            np.save(os.path.join(shm, f"bsk_slab_{rank}.npy"), kept["slab"].numpy())
            if world > 1:
                dist.barrier()
            mesh_np = None
            if rank == 0:
                mesh_np = np.concatenate([np.load(os.path.join(shm, f"bsk_slab_{r}.npy")) for r in range(world)], axis=0)
            if world > 1:
                dist.barrier()
            os.remove(os.path.join(shm, f"bsk_slab_{rank}.npy"))
        if rank == 0:
            from oracle import bskit_oracle as orc
            tarr = np.asarray(triples)
            sel = np.flatnonzero(np.isin(tarr, bins).all(axis=1))
            t0 = time.perf_counter()
            want = orc.measure_unnormalized([mesh_np], syn.BOX, edges, tarr[sel], workers=cores)
            got = full["out"][0][sel]
            rel = np.abs(got - want) / np.abs(want)
            oracle_sample = {"shells": bins, "triangles": int(len(sel)), "oracle": "oracle/bskit_oracle.py, float64, full grid, "
                             "mesh gathered from the ranks", "seconds": time.perf_counter() - t0,
                             "rel_err": [float(x) for x in rel], "max_rel_err": float(rel.max()),
                             "B_over_rms": [float(abs(w) / np.sqrt(np.mean(full["out"][0] ** 2))) for w in want]}
            del mesh_np
    tf32_peak = measure_tf32_peak(dev) if rank == 0 else None
    if world > 1:
        dist.barrier()
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- checks (outside timing)
    bf, ba = full["out"][0], auto["out"][0]
    rms = float(np.sqrt(np.mean(bf ** 2)))
    checks = {"auto_vs_full_max_abs_over_rms": float(np.max(np.abs(bf - ba)) / rms),
              "N_tri_max_abs_residual_before_rounding": full["ntri_residual"],
              "checksum": checksum(bf, full["out"][1], full["out"][2])}
    if oracle_sample is not None:
        checks["vs_oracle_sample"] = oracle_sample
    if b_api is not None:
        checks["api_vs_engine_max_abs_over_rms"] = float(np.max(np.abs(b_api["B"] - bf)) / rms)
    if host is not None and nmesh == 512 and len(edges) == 40 and os.path.exists(FIXTURE):
        fx = np.load(FIXTURE)
        if np.array_equal(fx["triples"].astype(np.int64), np.asarray(triples)) and np.array_equal(fx["edges"], edges):
            checks["vs_oracle"] = {
                "oracle": "tests/golden/metric512_oracle.npz: oracle/bskit_oracle.py (float64, full 512^3 grid, "
                          "one np.sum(I_a*I_b*I_c) per triangle) on the same seeded mesh, ALL triangles; "
                          "per-triangle |B - B_oracle| / |B_oracle|, no floor",
                "default_full_grid": dict(rel_error_stats(bf, fx["B"]), schedule=full["schedule"]),
                "auto_grid": dict(rel_error_stats(ba, fx["B"]), schedule=auto["schedule"])}
            if alt is not None:
                checks["vs_oracle"]["other_contraction_path"] = dict(rel_error_stats(alt["out"][0], fx["B"]),
                                                                     schedule=alt["schedule"])
    # normalisation against the oracle on the spot (64 triangles spread over the list; the oracle needs only
    # a grid with 3 n_max < N)
    try:
        from oracle import bskit_oracle as orc
        pick = np.unique(np.linspace(0, ntri - 1, 64).astype(int))
        n_or = 128 if 3 * (len(edges) + 1) < 128 else nmesh
        wn, wk = orc.measure_gridinfo(n_or, syn.BOX, edges, np.asarray(triples)[pick], workers=cores)
        gn, gk = full["out"][1][pick], full["out"][2][pick]
        checks["normalisation_vs_oracle_sample"] = {
            "triangles": int(len(pick)), "N_tri_equal": bool(np.array_equal(gn, np.rint(wn))),
            "k_mean_max_rel": float(np.max(np.abs(gk - wk) / np.abs(wk)))}
    except Exception as exc:   # pragma: no cover
        checks["normalisation_vs_oracle_sample"] = {"error": repr(exc)}

    # ---- roofline of the dominant kernel
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
    ncu = {}
    try:
        ncu = json.load(open(NCU_SUMMARY))
    except Exception:
        pass
    t_contract = full["stages_ms"]["contract"] * 1e-3
    cells = float(full["ncells"])
    alg_bytes = float(len(edges)) * 4.0 * cells                     # every shell value read once
    npairs = len({(min(a, b), max(a, b)) for a, b, _ in np.sort(np.asarray(triples), axis=1)[:, [0, 1, 2]]})
    useful_flops = 2.0 * cells * ntri + cells * npairs             # 2XT + XP
    clk = (full["clocks"] or {}).get("sm_mhz") or 1900.0
    if default_tensor:
        import ctypes as C
        rows32 = np.ascontiguousarray(triples, dtype=np.int32)
        out6 = (C.c_int64 * 6)()
        nat.lib().bsk_tc_schedule_info(len(rows32), rows32.ctypes.data_as(C.POINTER(C.c_int32)),
                                       (len(edges) + 3) // 4 * 4, out6)
        sched = {"units": int(out6[0]), "accumulator_columns": int(out6[2]), "passes": int(out6[3])}
        issued = 6.0 * 128.0 * sched["accumulator_columns"] * cells   # 3 MMAs of 2*128*N flops per cell and unit
        roofline = {
            "kernel": "tc_contract_kernel (tcgen05.mma kind::tf32, 3xTF32, pair products in TMEM), "
                      f"{sched['passes']} launch(es) per step",
            "bound": "tensor", "unit": "TFLOP/s", "peak": tf32_peak,
            "peak_source": "cuBLAS TF32 8192^3 measured in this run (best of 10, outside timing); "
                           "MEASURED_PEAKS.json holds bf16 only",
            "achieved": useful_flops / t_contract / 1e12, "frac": useful_flops / t_contract / 1e12 / tf32_peak,
            "frac_basis": "useful flops 2*X*T + X*P (SURVEY 8d)",
            "achieved_issued": issued / t_contract / 1e12, "frac_issued": issued / t_contract / 1e12 / tf32_peak,
            "frac_issued_of_half_measured_bf16_peak": (issued / t_contract / 1e12 / (0.5 * float(peaks["bf16_tflops"])))
            if peaks.get("bf16_tflops") else None,
            "useful_flops_per_launch_set": useful_flops, "issued_flops_per_launch_set": issued, "schedule": sched,
            "traffic": ncu.get("dram_bytes_per_step") if (nmesh == 512 and len(edges) == 40 and world == 1) else None,
            "traffic_source": ncu.get("source"),
            "algorithmic_bytes_per_pass": alg_bytes, "hbm_GBps_algorithmic": alg_bytes * sched["passes"] / t_contract / 1e9,
            "kernel_ms": t_contract * 1e3,
            "ncu": {k: ncu.get(k) for k in ("tensor_pipe_active_pct", "issue_active_pct", "smem_wavefront_pct",
                                            "local_load_requests", "kernel_ms")} if ncu else None,
            "note": "bounded by the generation of the A operand (pair products on the CUDA cores, TMEM stores at "
                    "~64 B/cycle per lane quarter, hand-over latency), not by the tensor pipe: DESIGN.md section 3, "
                    "profiles/r2_tc_contract_history.md"}
    else:
        issued = full["cplan"]["nblocks"] * 80.0 * 2.0 * cells if full["cplan"] else None
        fp32_peak = 72.5e12 * clk / 1965.0        # scripts/dev/ffma_probe.cu at 1965 MHz (round 1)
        roofline = {"kernel": "tile_contract_kernel<float,%s>" % ("float" if accum == nat.F32 else "double"),
                    "bound": "hbm", "achieved": alg_bytes / t_contract / 1e9, "peak": hbm_peak, "unit": "GB/s",
                    "frac": alg_bytes / t_contract / 1e9 / hbm_peak, "traffic": None, "peak_source": peak_src,
                    "algorithmic_bytes_per_launch": alg_bytes, "kernel_ms": t_contract * 1e3,
                    "note": "the kernel is FP32-FMA-pipe bound, not HBM bound: see fp32_pipe",
                    "fp32_pipe": {"issued_flops_per_launch": issued, "useful_flops_per_launch": useful_flops,
                                  "issued_tflops": (issued / t_contract / 1e12) if issued else None,
                                  "frac_issued": (issued / t_contract / fp32_peak) if issued else None}}
    shells_bytes = float(len(edges)) * cells * (8 * (nmesh // 2 + 1) / nmesh + 4)
    stage_roofs = {"shells": {"algorithmic_bytes": shells_bytes,
                              "achieved_GBps": shells_bytes / (full["stages_ms"]["shells"] * 1e-3) / 1e9,
                              "frac_of_hbm_peak": shells_bytes / (full["stages_ms"]["shells"] * 1e-3) / 1e9 / hbm_peak,
                              "output_bytes": float(len(edges)) * cells * 4,
                              "frac_of_hbm_peak_on_output_bytes": float(len(edges)) * cells * 4
                              / (full["stages_ms"]["shells"] * 1e-3) / 1e9 / hbm_peak}}

    cpu = None
    if world == 1 and host is not None and not args.no_cpu_baseline:
        gate = oracle_golden_gate()
        v, sample = cpu_reference_sample(host.numpy(), len(edges), ntri, cores, list(range(len(edges))),
                                         n_tri_sample=64, n_field64=2, n_tri64=4)
        cpu = {"value": v, "unit": "s", "cores": cores, "kind": "port",
               "sample": f"validity gate: {gate}; " + sample}

    value = full["ms_per_step"] * 1e-3
    line = {
        "metric": METRIC, "value": value, "unit": "s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": full["ms_per_step"], "higher_is_better": False,
        "scaling": "strong", "vs_baseline": None,
        "dtype": "f32 shell fields and products; f64 forward/inverse FFT, f64 tile reduction, f64 normalisation"
                 if accum == nat.F32 else "f32 shell fields; f64 FFT, products and accumulation",
        "data": "synthetic" if host is not None else "synthetic (generated per rank on the GPU)",
        "config": config,
        "grid": {"eval_grid": int(full["grid"].neval), "ncrop": int(full["grid"].ncrop)},
        "contraction": full["schedule"],
        "stages_ms": full["stages_ms"], "roofline": roofline, "stage_rooflines": stage_roofs,
        "cpu_baseline": cpu,
        "e2e": ({"value": e2e_s, "unit": "s", "h2d_bytes_per_step": int(nmesh ** 3 * 4),
                 "d2h_bytes_per_step": int(ntri * 8 * 5), "api": "bskit_b200.FFTBispectrum(host mesh)"
                 ".measure_bispectrum_faster + measure_gridinfo_faster", "steps": n_e2e,
                 "cold_first_call_s": e2e_cold_s,
                 "note": "warm: cuFFT plans, contraction schedules and scratch are cached per process; "
                         "cold_first_call_s is the first call of the process; the (N_tri, k_mean) cache is off"}
                if e2e_s is not None else None),
        "gpu_launches": int(full["launches"]),
        "clocks": full["clocks"],
        "auto_grid": {"value": auto["ms_per_step"] * 1e-3, "unit": "s", "eval_grid": int(auto["grid"].neval),
                      "stages_ms": auto["stages_ms"], "schedule": auto["schedule"],
                      "note": "exact band-limited evaluation (library default); same outputs"},
        "checks": checks,
    }
    if alt is not None:
        t_alt = alt["stages_ms"]["contract"] * 1e-3
        line["other_contraction_path"] = {
            "schedule": alt["schedule"], "ms_per_step": alt["ms_per_step"], "contract_ms": t_alt * 1e3,
            "default_contract_ms": t_contract * 1e3,
            "kernel": "tile_contract_kernel<float,float> (FP32 pipe, packed FFMA2; contraction='fp32')" if default_tensor
            else "tc_contract_kernel (tcgen05, 3xTF32; contraction='tensor')"}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
