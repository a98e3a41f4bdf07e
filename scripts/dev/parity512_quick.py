import os, sys, json, numpy as np, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
if len(sys.argv) > 1:
    from bskit_b200 import _native
    _native.LIB_PATH = os.path.abspath(sys.argv[1])
import torch
import bskit_b200 as bk
from bskit_b200 import synthetic as syn
fx = np.load(os.path.join(ROOT, "tests", "golden", "metric512_oracle.npz"))
want = fx["B"]; n = int(fx["nmesh"]); nb = len(fx["edges"])
mesh = syn.lognormal_mesh(n, seed=1, workers=os.cpu_count())
kmin, kmax, dk = syn.bench_bins(nb)
rms = np.sqrt(np.mean(want ** 2))
fb = bk.FFTBispectrum(mesh, BoxSize=syn.BOX, kmin=kmin, kmax=kmax, dk=dk, grid="full", contraction="tensor")
got = fb.measure_bispectrum_faster()["B"]
torch.cuda.synchronize(); t0 = time.time()
for _ in range(3): fb.b = None; got = fb.measure_bispectrum_faster()["B"]
torch.cuda.synchronize(); dt = (time.time() - t0) / 3
rel = np.abs(got - want) / np.abs(want)
print(sys.argv[1:] , "ms/measure %.1f" % (dt * 1e3), "q50 %.2e q99 %.2e max %.2e n>1e-5 %d abs/rms max %.2e signed mean %.2e" % (
    np.median(rel), np.quantile(rel, 0.99), rel.max(), int((rel > 1e-5).sum()), np.abs(got - want).max() / rms, np.mean((got - want) / want)))
