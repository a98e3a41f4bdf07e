"""Dev: config C5 in reduced form — 2048^3 mesh, equilateral + squeezed bins, x-slab sharded.
Each rank generates its own white-noise slab on the device.  Prints per-stage times (max over
ranks) for the forward transform, the equilateral and the squeezed measurement."""
import os, sys, time
# (expandable_segments makes 100+ GiB allocations take tens of seconds: not used)
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, ".")
import bskit_b200 as bk
from bskit_b200 import synthetic as syn
rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1)); lr = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(lr); dev = torch.device("cuda", lr)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
nb = int(sys.argv[2]) if len(sys.argv) > 2 else 100
ntri = int(sys.argv[3]) if len(sys.argv) > 3 else 32
g = torch.Generator(device=dev); g.manual_seed(1234 + rank)
slab = torch.randn((n // world, n, n), dtype=torch.float32, device=dev, generator=g)
kmin, kmax, dk = syn.bench_bins(nb)
def sync_t():
    torch.cuda.synchronize()
    if world > 1: dist.barrier()
    return time.perf_counter()
out = {}
for tt, kw in (("equilateral", {}), ("squeezed", dict(squeezed_bin_index=0))):
    for rep in range(2):
        t0 = sync_t()
        fb = bk.FFTBispectrum(slab, Nmesh=n, BoxSize=syn.BOX, kmin=kmin, kmax=kmax, dk=dk, triangle_type=tt, grid="full", device=dev, **kw)
        fb._meas().cubes(list(fb._meas().session._engines.values())[0])
        t1 = sync_t()
        b = fb.measure_bispectrum_faster(0, ntri)
        t2 = sync_t()
        e = list(fb._meas().session._engines.values())[0]
        out[tt] = (t1 - t0, t2 - t1, e.last_batches, e.last_schedule, e.row_capacity())
        fb.close()
if rank == 0:
    for tt, v in out.items():
        print(f"C5 {n}^3 S={nb} {tt} first {ntri} triangles, {world} GPU(s): forward {v[0]*1e3:.0f} ms, measure {v[1]*1e3:.0f} ms, batches {v[2]}, schedule {v[3]}, rowcap {v[4]}")
if world > 1:
    dist.destroy_process_group()
