"""Dev: capacity / overflow check at 1024^3 (all triangles, forces batches on one GPU) and
2048^3 (equilateral + squeezed).  White-noise device-generated meshes; prints stage times."""
import os, sys, time
# (expandable_segments makes 100+ GiB allocations take tens of seconds: not used)
import numpy as np, torch
sys.path.insert(0, ".")
import bskit_b200 as bk
from bskit_b200 import engine as eng, _native as nat, synthetic as syn
dev = torch.device("cuda", 0)
which = sys.argv[1]
def t():
    torch.cuda.synchronize(); return time.perf_counter()
if which == "1024":
    n, nb = 1024, 40
    mesh = torch.randn((n, n, n), dtype=torch.float32, device=dev)
    kmin, kmax, dk = syn.bench_bins(nb)
    t0 = t()
    fb = bk.FFTBispectrum(mesh, BoxSize=syn.BOX, kmin=kmin, kmax=kmax, dk=dk, grid="full", device=dev)
    t1 = t()
    b = fb.measure_bispectrum_faster(0, 10**9)
    t2 = t()
    e = list(fb._meas().session._engines.values())[0]
    print("1024 full: ctor %.2fs measure %.2fs batches %d chunk %d rowcap %d  B[0:3]=%s" % (t1 - t0, t2 - t1, e.last_batches, e.chunk, e.row_capacity(), b["B"][:3]))
    fb.close(); del fb; torch.cuda.empty_cache()
    fa = bk.FFTBispectrum(mesh, BoxSize=syn.BOX, kmin=kmin, kmax=kmax, dk=dk, grid="auto", device=dev)
    t3 = t(); ba = fa.measure_bispectrum_faster(0, 10**9); t4 = t()
    rms = np.sqrt(np.mean(b["B"] ** 2))
    print("1024 auto: measure %.3fs  max|full-auto|/rms = %.2e" % (t4 - t3, np.abs(ba["B"] - b["B"]).max() / rms))
else:
    n, nb = 2048, 300
    mesh = torch.randn((n, n, n), dtype=torch.float32, device=dev)
    kmin, kmax, dk = syn.bench_bins(nb)
    for tt, kw in (("equilateral", {}), ("squeezed", dict(squeezed_bin_index=0))):
        t0 = t()
        fb = bk.FFTBispectrum(mesh, BoxSize=syn.BOX, kmin=kmin, kmax=kmax, dk=dk, triangle_type=tt, grid="full", device=dev, **kw)
        t1 = t()
        b = fb.measure_bispectrum_faster(0, 40)           # first 40 triangles of the list
        t2 = t()
        e = list(fb._meas().session._engines.values())[0]
        print("2048 %s: ctor %.2fs measure(40 tri) %.2fs batches %d chunk %d rowcap %d grid %s mem %.1f GB" % (
            tt, t1 - t0, t2 - t1, e.last_batches, e.chunk, e.row_capacity(), e.grid, torch.cuda.max_memory_allocated() / 2**30))
        fb.close(); torch.cuda.empty_cache()
        fa = bk.FFTBispectrum(mesh, BoxSize=syn.BOX, kmin=kmin, kmax=kmax, dk=dk, triangle_type=tt, grid="auto", device=dev, **kw)
        t3 = t(); ba = fa.measure_bispectrum_faster(0, 40); t4 = t()
        rms = np.sqrt(np.mean(b["B"] ** 2))
        print("   auto grid %.2fs %s: max|full-auto|/rms = %.2e" % (t4 - t3, list(fa._meas().session._engines.values())[0].grid, np.abs(ba["B"] - b["B"]).max() / rms))
        fa.close(); del fb, fa; torch.cuda.empty_cache()
