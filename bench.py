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
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

TC_DRAM_BYTES = 24.6e9      # ncu dram read + write of tc_contract_kernel at 512^3, S=40 (profiles/)
METRIC = "s per 512^3 all-triangle bispectrum"


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
    ap.add_argument("--profile", action="store_true",
                    help="only the full-grid device-resident steps (for ncu); prints stage times")
    return ap.parse_args()


# --------------------------------------------------------------------------- #
# CPU leg: the reference algorithm (oracle port) on the host cores
# --------------------------------------------------------------------------- #
def cpu_reference_sample(mesh32, nbins, ntri, cores, n_shell_sample=2, n_tri_sample=6):
    """Time a bounded sample of the reference's algorithm (bskit/main.py:1846-1879 and
    2006-2061) and extrapolate linearly to the whole job: S masked inverse FFTs (f4, the
    mesh dtype) + T full-grid triple-product sums, then 2S f8 fields + 4T f8 sums."""
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
    pick = np.linspace(0, nbins - 1, n_shell_sample).astype(int)
    t0 = time.perf_counter()
    shells = []
    for i in pick:
        kk = orc.k_norm(n, syn.BOX)                    # the reference recomputes |k| per bin
        m = dk32 * orc.shell_mask(kk, edges[i, 0], edges[i, 1])
        shells.append((sfft.irfftn(m, s=(n, n, n), workers=cores) * np.float32(n) ** 3).astype(np.float32))
        del kk, m
    t_shell = (time.perf_counter() - t0) / len(pick)
    t0 = time.perf_counter()
    for j in range(n_tri_sample):
        tri_sum(shells[0], shells[-1], shells[j % len(shells)])
    t_tri = (time.perf_counter() - t0) / n_tri_sample
    # normalisation leg: float64 number / k fields (always f8 in the reference)
    t0 = time.perf_counter()
    kk = orc.k_norm(n, syn.BOX)
    nf = orc.number_field(n, syn.BOX, edges[pick[0], 0], edges[pick[0], 1], kk, cores)
    kf = orc.k_field(n, syn.BOX, edges[pick[0], 0], edges[pick[0], 1], 1.0, kk, cores)
    t_nfield = (time.perf_counter() - t0) / 2 + t_shell * 0.0
    t0 = time.perf_counter()
    for j in range(2):
        tri_sum(nf, kf, nf)
    t_tri64 = (time.perf_counter() - t0) / 2
    pool.shutdown()
    total = t_fwd + nbins * t_shell + ntri * t_tri + 2 * nbins * t_nfield + 4 * ntri * t_tri64
    sample = (f"oracle port (numpy/scipy restatement, not nbodykit): timed 1 forward rfftn, "
              f"{len(pick)} of {nbins} f4 shell builds, {n_tri_sample} of {ntri} f4 triangle sums, "
              f"2 of {2 * nbins} f8 number/k fields, 2 of {4 * ntri} f8 sums on {n}^3; "
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
def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    nmesh, nbins = args.nmesh, args.nbins

    from bskit_b200 import synthetic as syn
    import bskit_b200 as bk
    kmin, kmax, dk = syn.bench_bins(nbins)
    edges = bk.generate_bin_edge_list(kmin, kmax, dk)
    triples = bk.generate_triangle_bin_list(kmin, kmax, dk, return_indices=True)
    ntri = len(triples)
    cores = host_cores()
    config = {"workload": f"{nmesh}^3 float32 lognormal mesh (seed 1, BoxSize 1000), S={len(edges)} "
                          f"k-bins of width k_f from k_f/2, all {ntri} triangles, auto + normalisation",
              "nmesh": nmesh, "nbins": int(len(edges)), "ntriangles": int(ntri),
              "parallelism": f"x-slab x{max(world, 1)}", "l2": "inputs larger than L2 (no flush needed)"}

    # ---------------------------------------------------------------- reference arm
    if args.impl == "reference":
        if rank != 0:
            return
        mesh = syn.lognormal_mesh(nmesh, seed=1, workers=cores)
        vals = []
        sample = ""
        for i in range(args.warmup + args.steps):
            v, sample = cpu_reference_sample(mesh, len(edges), ntri, cores,
                                             n_shell_sample=1 if i < args.warmup else 2,
                                             n_tri_sample=2 if i < args.warmup else 4)
            if i >= args.warmup:
                vals.append(v)
        v = float(np.median(vals))
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

    def make(policy, tensor=None):
        g = eng.choose_grid(nmesh, syn.BOX, edges[:, 1].max(), policy, world)
        gn = eng.choose_grid(nmesh, syn.BOX, edges[:, 1].max(), "auto", world)
        e_data = eng.Engine(g, syn.BOX, nat.F32, device=dev, accum_precision=accum)
        if tensor is not None:                      # None: the library default
            e_data.backend.contraction_path = 1 if tensor else 0   # include/bskit_b200.h, bsk_cplan_set_path
        e_norm = eng.Engine(gn, syn.BOX, nat.F64, device=dev)
        return e_data, e_norm

    vol2 = syn.BOX ** 6

    def step(e_data, e_norm, slab, marks=None):
        eng._mark(marks, "start", e_data)
        cube = e_data.forward(slab)
        eng._mark(marks, "forward_done", e_data)
        b = eng.measure_triangle_sums(e_data, [cube], edges, triples, marks=marks) * vol2
        eng._mark(marks, "data_done", e_data)
        ntri_v, kmean = eng.measure_grid_sums(e_norm, edges, triples)
        eng._mark(marks, "norm_done", e_data)
        return b, ntri_v, kmean

    def timed(policy, steps, warmup, with_clocks=False, tensor=None):
        e_data, e_norm = make(policy, tensor)
        slab = e_data.local_slab(host)
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
                   clocks=sampler.summary() if sampler else None)
        e_data.close()
        e_norm.close()
        del slab
        torch.cuda.empty_cache()
        return res

    full = timed("full", args.steps, args.warmup, with_clocks=not args.profile)
    if args.profile:
        if rank == 0:
            print(json.dumps({"profile_only": True, "ms_per_step": full["ms_per_step"],
                              "stages_ms": full["stages_ms"], "gpu_launches": full["launches"]}))
        return
    auto = timed("auto", args.steps, args.warmup)
    # the same step with the other contraction kernel (tensor cores <-> FP32 pipe), reported beside the default
    default_tensor = full["schedule"] == "tensor"
    alt = timed("full", 2, 1, tensor=not default_tensor) if accum == nat.F32 else None

    # ---- e2e: the user-facing API from (pinned) host memory, result back on the host
    def e2e_once():
        fb = bk.FFTBispectrum(host, BoxSize=syn.BOX, kmin=kmin, kmax=kmax, dk=dk, grid="full",
                              accum_dtype=np.float32 if accum == nat.F32 else np.float64, device=dev)
        g = fb.measure_gridinfo_faster(0, ntri)      # step 1 of the reference workflow; overlaps the upload
        b = fb.measure_bispectrum_faster(0, ntri)
        fb.close()
        return b, g

    for _ in range(max(1, min(args.warmup, 2))):
        e2e_once()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    wall0 = time.perf_counter()
    e0.record()
    n_e2e = max(1, min(args.steps, 3))
    for _ in range(n_e2e):
        b_api, g_api = e2e_once()
    e1.record()
    torch.cuda.synchronize()
    e2e_s = (time.perf_counter() - wall0) / n_e2e       # host-side work (plan build, D2H) counts too
    if world > 1:
        t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # consistency of the two evaluations and of the API leg (cheap sanity, outside timing)
    bf, ba = full["out"][0], auto["out"][0]
    rms = float(np.sqrt(np.mean(bf ** 2)))
    agree = float(np.max(np.abs(bf - ba)) / rms)
    api_agree = float(np.max(np.abs(b_api["B"] - bf)) / rms)

    # ---- roofline of the dominant kernel (tile_contract_kernel, float32 fields)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
    t_contract = full["stages_ms"]["contract"] * 1e-3
    alg_bytes = float(len(edges)) * 4.0 * full["ncells"]           # every shell value read once
    achieved = alg_bytes / t_contract / 1e9
    cells = float(full["ncells"])
    useful_flops = 2.0 * cells * ntri + cells * 820.0 * (len(edges) == 40)
    issued = full["cplan"]["nblocks"] * 80.0 * 2.0 * cells if full["cplan"] else None
    clk = (full["clocks"] or {}).get("sm_mhz") or 1900.0
    fp32_peak = 72.5e12 * clk / 1965.0        # measured by scripts/dev/ffma_probe.cu at 1965 MHz
    # issued tensor-core work: 3 MMAs (P_lo*C_hi, P_hi*C_lo, P_hi*C_hi) of 2*128*N flops per cell and unit;
    # the N of the units of the S=40 all-triangle schedule sum to 184
    issued_tc = 6.0 * 128.0 * 184.0 * cells if (len(edges) == 40 and ntri == 6730) else None
    tf32_peak = 741.0            # TFLOP/s, cuBLAS TF32 8192^3 measured on this pool's B200 (profiles/r1_extra_peaks.txt)
    fp32_info = {"issued_flops_per_launch": issued, "useful_flops_per_launch": useful_flops,
                 "peak_tflops_measured_ffma_probe": fp32_peak / 1e12}
    if default_tensor:
        roofline = {"kernel": "tc_contract_kernel (tcgen05.mma kind::tf32, 3xTF32, pair products in TMEM)",
                    "bound": "tensor", "achieved": (issued_tc / t_contract / 1e12) if issued_tc else None,
                    "peak": tf32_peak, "unit": "TFLOP/s",
                    "frac": (issued_tc / t_contract / 1e12 / tf32_peak) if issued_tc else None,
                    # dram__bytes_read.sum + dram__bytes_write.sum of one `ncu --set full` capture of this kernel
                    # at this configuration (profiles/r1_tc_contract_ncu_summary.txt)
                    "traffic": TC_DRAM_BYTES if (nmesh == 512 and len(edges) == 40 and world == 1) else None,
                    "peak_source": "measured TF32 cuBLAS 8192^3 (profiles/r1_extra_peaks.txt); MEASURED_PEAKS.json holds bf16 only",
                    "frac_of_half_measured_bf16_peak": (issued_tc / t_contract / 1e12 / (0.5 * float(peaks["bf16_tflops"])))
                    if (issued_tc and peaks.get("bf16_tflops")) else None,
                    "issued_flops_per_launch": issued_tc, "useful_flops_per_launch": useful_flops,
                    "algorithmic_bytes_per_launch": alg_bytes, "hbm_GBps": achieved, "kernel_ms": t_contract * 1e3,
                    "note": "bounded by operand generation (shared-memory reads of the two rows of every pair) "
                            "rather than the tensor pipe: see DESIGN.md"}
    else:
        roofline = {"kernel": "tile_contract_kernel<float,%s>" % ("float" if accum == nat.F32 else "double"),
                    "bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
                    "frac": achieved / hbm_peak,
                    # dram__bytes_read.sum + dram__bytes_write.sum of one `ncu --set full` capture of this
                    # kernel at this exact configuration (profiles/r1_tile_contract_packed_ncu_raw.csv)
                    "traffic": 21.513e9 if (nmesh == 512 and len(edges) == 40 and world == 1) else None,
                    "peak_source": peak_src,
                    "algorithmic_bytes_per_launch": alg_bytes, "kernel_ms": t_contract * 1e3,
                    "note": "the kernel is FP32-FMA-pipe bound, not HBM bound: see fp32_pipe",
                    "fp32_pipe": dict(fp32_info, issued_tflops=(issued / t_contract / 1e12) if issued else None,
                                      frac_issued=(issued / t_contract / fp32_peak) if issued else None)}
    shells_bytes = float(len(edges)) * full["ncells"] * (8 * (nmesh // 2 + 1) / nmesh + 4)
    stage_roofs = {"shells": {"algorithmic_bytes": shells_bytes,
                              "achieved_GBps": shells_bytes / (full["stages_ms"]["shells"] * 1e-3) / 1e9,
                              "frac_of_hbm_peak": shells_bytes / (full["stages_ms"]["shells"] * 1e-3) / 1e9 / hbm_peak}}

    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        v, sample = cpu_reference_sample(host.numpy(), len(edges), ntri, cores)
        cpu = {"value": v, "unit": "s", "cores": cores, "kind": "port", "sample": sample}

    value = full["ms_per_step"] * 1e-3
    line = {
        "metric": METRIC, "value": value, "unit": "s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": full["ms_per_step"], "higher_is_better": False,
        "scaling": "strong", "vs_baseline": None,
        "dtype": "f32 shell fields and products; f64 forward/inverse FFT, f64 tile reduction, f64 normalisation"
                 if accum == nat.F32 else "f32 shell fields; f64 FFT, products and accumulation",
        "data": "synthetic",
        "config": dict(config, eval_grid=int(full["grid"].neval), ncrop=int(full["grid"].ncrop)),
        "stages_ms": full["stages_ms"], "roofline": roofline, "stage_rooflines": stage_roofs,
        "cpu_baseline": cpu,
        "e2e": {"value": e2e_s, "unit": "s", "h2d_bytes_per_step": int(nmesh ** 3 * 4),
                "d2h_bytes_per_step": int(ntri * 8 * 5), "api": "bskit_b200.FFTBispectrum(host mesh)"
                ".measure_bispectrum_faster + measure_gridinfo_faster", "steps": n_e2e},
        "gpu_launches": int(full["launches"]),
        "clocks": full["clocks"],
        "auto_grid": {"value": auto["ms_per_step"] * 1e-3, "unit": "s", "eval_grid": int(auto["grid"].neval),
                      "stages_ms": auto["stages_ms"],
                      "max_abs_diff_vs_full_over_rms": agree,
                      "note": "exact band-limited evaluation (library default); same outputs"},
        "checks": {"api_vs_engine_max_abs_over_rms": api_agree},
    }
    if alt is not None:
        t_alt = alt["stages_ms"]["contract"] * 1e-3
        info = {"schedule": alt["schedule"], "ms_per_step": alt["ms_per_step"], "contract_ms": t_alt * 1e3,
                "default_contract_ms": t_contract * 1e3,
                "max_abs_diff_vs_default_over_rms": float(np.max(np.abs(alt["out"][0] - bf)) / rms)}
        if default_tensor:
            info.update(kernel="tile_contract_kernel<float,float> (FP32 pipe, packed FFMA2; BSKIT_B200_CONTRACTION=fp32)",
                        fp32_pipe=dict(fp32_info, issued_tflops=(issued / t_alt / 1e12) if issued else None,
                                       frac_issued=(issued / t_alt / fp32_peak) if issued else None),
                        note="exact round-to-nearest products: error floor ~5e-8 of max|B| instead of ~1e-6")
        else:
            info.update(kernel="tc_contract_kernel (tcgen05.mma kind::tf32, 3xTF32, pair products in TMEM; "
                               "BSKIT_B200_CONTRACTION=tensor)",
                        roofline={"bound": "tensor", "achieved": (issued_tc / t_alt / 1e12) if issued_tc else None,
                                  "peak": tf32_peak, "unit": "TFLOP/s",
                                  "frac": (issued_tc / t_alt / 1e12 / tf32_peak) if issued_tc else None},
                        note="3xTF32 with truncating tensor-core accumulators: ~1e-6 relative instead of ~5e-8")
        line["other_contraction_path"] = info
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
