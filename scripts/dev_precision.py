"""Dev experiment: where does the float32 error of B come from (FFT vs accumulation)?"""
import numpy as np, torch, sys
sys.path.insert(0, ".")
from bskit_b200 import engine as eng, _native as nat, synthetic as syn
from oracle import bskit_oracle as orc

n, nb = 64, 12
kmin, kmax, dk = syn.bench_bins(nb)
mesh = syn.lognormal_mesh(n, seed=1)
edges = orc.bin_edges(kmin, kmax, dk)
_, idx = orc.triangles_all(edges, 1)
want = orc.measure_unnormalized([mesh], syn.BOX, edges, idx, workers=4) / syn.BOX ** 6
dev = torch.device("cuda", 0)
for pol in ("full",):
    g = eng.choose_grid(n, syn.BOX, edges[:, 1].max(), pol)
    e = eng.Engine(g, syn.BOX, nat.F32, device=dev)
    cube = e.forward(mesh)
    sp = 12
    table = torch.empty((sp, e.ncells), dtype=torch.float32, device=dev)
    e.synthesize(cube, nat.KIND_DATA, 0.0, edges[:, 0], edges[:, 1], table)
    mine = e.contract(table, idx)[0] / n ** 3
    t64 = table.double()
    ref64 = np.array([(t64[a] * t64[b] * t64[c]).sum().item() for a, b, c in idx]) / n ** 3
    p32 = np.array([((table[a] * table[b]) * table[c]).double().sum().item() for a, b, c in idx]) / n ** 3
    # oracle shells in fp64 -> compare shell fields
    dk64 = orc.forward(mesh.astype(np.float64))
    kk = orc.k_norm(n, syn.BOX)
    sh = np.stack([orc.data_shell(dk64, kk, lo, hi) for lo, hi in edges]).reshape(nb, -1)
    d = (table.cpu().numpy().astype(np.float64) - sh)
    print("shell field rel err (rms err / rms):", np.sqrt((d ** 2).mean(1)) / np.sqrt((sh ** 2).mean(1)))
    scale = np.abs(want).max()
    big = np.abs(want) > 1e-3 * scale
    for name, v in (("kernel", mine), ("fp32 shells, fp64 products+sum", ref64), ("fp32 products, fp64 sum", p32)):
        rel = np.abs(v - want)[big] / np.abs(want)[big]
        print(f"{pol:5s} {name:34s} max rel {rel.max():.2e}  median {np.median(rel):.2e}  max abs/scale {np.abs(v-want).max()/scale:.2e}")

print("---- precision modes (storage f32) ----")
import bskit_b200 as bk
for fft, acc in ((np.float32, np.float32), (np.float64, np.float32), (np.float64, np.float64)):
    for meshname, m in (("lognormal", mesh), ("gaussian", syn.gaussian_mesh(n, seed=1))):
        w = orc.measure_unnormalized([m], syn.BOX, edges, idx, workers=4)
        fb = bk.FFTBispectrum(m, BoxSize=syn.BOX, kmin=kmin, kmax=kmax, dk=dk, grid="full", fft_dtype=fft, accum_dtype=acc)
        got = fb.measure_bispectrum_faster(0, 10**9)["B"]
        fb.close()
        sc = np.abs(w).max(); rms = np.sqrt((w**2).mean())
        for thr in (1e-3, 1e-2, 5e-2):
            bigm = np.abs(w) > thr * sc
            print(f"fft={np.dtype(fft).name} acc={np.dtype(acc).name} {meshname:9s} thr={thr:g}: max rel {np.max(np.abs(got-w)[bigm]/np.abs(w)[bigm]):.2e}", end=" | ")
        print(f"max abs/max {np.abs(got-w).max()/sc:.2e} abs/rms {np.abs(got-w).max()/rms:.2e}")
