"""Dev: per-stage times of one shell + a 3-row contraction at 2048^3 (S=300 binning)."""
import os, sys, time
# (expandable_segments makes 100+ GiB allocations take tens of seconds: not used)
import numpy as np, torch
sys.path.insert(0, ".")
from bskit_b200 import engine as eng, _native as nat, synthetic as syn
dev = torch.device("cuda", 0)
n = int(sys.argv[1]); nb = int(sys.argv[2])
kmin, kmax, dk = syn.bench_bins(nb)
import bskit_b200 as bk
edges = bk.generate_bin_edge_list(kmin, kmax, dk)
g = eng.choose_grid(n, syn.BOX, edges[:, 1].max(), "full")
e = eng.Engine(g, syn.BOX, nat.F32, device=dev)
print("grid", g, "chunk", e.chunk, "info ky,kz", e.info.ky, e.info.kz)
mesh = torch.randn((n, n, n), dtype=torch.float32, device=dev)
def ev():
    x = torch.cuda.Event(enable_timing=True); x.record(); return x
t0 = ev(); cube = e.forward(mesh); t1 = ev()
table = torch.empty((3, e.ncells), dtype=torch.float32, device=dev)
for rep in range(2):
    a = ev()
    e.synthesize(cube, nat.KIND_DATA, 0.0, edges[10:11, 0], edges[10:11, 1], table[0:1])
    b = ev()
    e.synthesize(cube, nat.KIND_DATA, 0.0, edges[200:202, 0], edges[200:202, 1], table[1:3])
    c = ev()
    rows = np.array([[0, 0, 0], [1, 1, 1], [2, 2, 2], [0, 1, 1]])
    fields = [table[0], table[1], table[2], table[0]]
    s = e.contract(fields, rows)
    d = ev(); torch.cuda.synchronize()
    print("rep", rep, "forward %.1f ms; 1 shell %.1f ms; 2 shells %.1f ms; contract(4 tri, 3 rows) %.1f ms" % (
        t0.elapsed_time(t1), a.elapsed_time(b), b.elapsed_time(c), c.elapsed_time(d)))
