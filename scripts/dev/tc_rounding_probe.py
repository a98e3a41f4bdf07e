"""Dev probe: how do B200 tensor cores round when accumulating TF32/BF16 products in FP32?
(Matters for a tcgen05 version of the triangle contraction: sums over ~1e6 cells per CTA.)
C = A @ B with K long; row 0 of A / col 0 of B engineered so that the exact result needs
rounding at every accumulation step.  Compare with simulated round-to-nearest / toward-zero."""
import numpy as np, torch
torch.backends.cuda.matmul.allow_tf32 = True
dev = "cuda"
def simulate(vals, chunk, mode):
    acc = np.float32(0.0)
    for i in range(0, len(vals), chunk):
        s = float(np.sum(vals[i:i + chunk].astype(np.float64)))      # chunk summed exactly
        exact = float(acc) + s
        if mode == "rn":
            acc = np.float32(exact)
        else:                                                          # toward zero
            r = np.float32(exact)
            if abs(float(r)) > abs(exact):
                r = np.nextafter(r, np.float32(0.0))
            acc = r
    return float(acc)
for dtype, name in ((torch.float32, "tf32"), (torch.bfloat16, "bf16")):
    K = 8192
    a = torch.zeros(128, K, dtype=torch.float32); b = torch.zeros(K, 128, dtype=torch.float32)
    a[:, 0] = 4096.0; b[0, :] = 4096.0          # 2^24: fp32 ulp = 2 afterwards
    a[:, 1:] = 1.1875; b[1:, :] = 1.0           # each product 1.1875 (exact in tf32 and bf16)
    c = (a.to(dev).to(dtype) @ b.to(dev).to(dtype)).float().cpu().numpy()[0, 0]
    vals = np.full(K, 1.1875); vals[0] = 2.0 ** 24
    exact = vals.sum()
    print(f"{name}: gpu={c:.1f} exact={exact:.1f} gpu-exact={c - exact:+.1f}")
    for chunk in (4, 8, 16, 32, 64):
        print(f"   chunk {chunk:3d}: RN -> {simulate(vals, chunk, 'rn') - exact:+9.1f}   RZ -> {simulate(vals, chunk, 'rz') - exact:+9.1f}")
    # random-sign cancelling sums: relative error of a long fp32-accumulated dot product
    rng = np.random.default_rng(0)
    x = rng.standard_normal((128, 1 << 16)).astype(np.float32); y = rng.standard_normal((1 << 16, 128)).astype(np.float32)
    xt = torch.from_numpy(x).to(dev); yt = torch.from_numpy(y).to(dev)
    if dtype == torch.float32:
        got = (xt @ yt).double().cpu().numpy()
        xq = (xt.view(torch.int32) & ~0x1fff).view(torch.float32); yq = (yt.view(torch.int32) & ~0x1fff).view(torch.float32)
        ref_q = (xq.double() @ yq.double()).cpu().numpy()            # exact result for truncated-to-tf32 inputs
        ref_rn = None
    else:
        got = (xt.to(dtype) @ yt.to(dtype)).double().cpu().numpy()
        ref_q = (xt.to(dtype).double() @ yt.to(dtype).double()).cpu().numpy()
    scale = np.sqrt((ref_q ** 2).mean())
    d = got - ref_q
    print(f"   random dot K=65536: mean err/rms = {d.mean()/scale:+.2e}, rms err/rms = {np.sqrt((d**2).mean())/scale:.2e}")
