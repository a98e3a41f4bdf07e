"""world_size-2 (and 4) runs of the host-side multi-rank logic under gloo on CPU:
x-slab partition of the mesh, all-gather of the cropped (y,z)-transformed planes,
sharded shell synthesis (each rank its own x-planes) and the final all-reduce.
The numeric stages are a numpy test double (tests/fake_backend.py); the result must
equal the single-process float64 oracle."""
import os
import sys
import tempfile

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT


def _worker(rank, world, init_file, grid_policy, outdir, nb=4):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    dist.init_process_group("gloo", init_method="file://" + init_file, rank=rank, world_size=world)
    from bskit_b200 import engine as eng, synthetic as syn
    from fake_backend import FakeBackend
    n = 16
    kmin, kmax, dk = syn.bench_bins(nb)
    from bskit_b200 import generate_bin_edge_list, generate_triangle_bin_list
    edges = generate_bin_edge_list(kmin, kmax, dk)
    triples = generate_triangle_bin_list(kmin, kmax, dk, return_indices=True)
    mesh = syn.lognormal_mesh(n, seed=1, dtype=np.float64)
    g = eng.choose_grid(n, syn.BOX, edges[:, 1].max(), grid_policy, world)
    e = eng.Engine(g, syn.BOX, 1, device=torch.device("cpu"), backend_cls=FakeBackend)
    assert e.world == world and e.info.nxl == n // world
    assert e.transposed == bool(g.full and world > 1)
    if e.transposed:      # exact mode counts are a collective over the ky blocks
        from oracle import bskit_oracle as orc
        assert np.array_equal(e.backend.modes_per_bin(edges[:, 0], edges[:, 1]), orc.modes_per_bin(n, syn.BOX, edges))
    cube = e.forward(mesh)                                   # full mesh given: each rank slices
    cube2 = e.forward(mesh[e.info.nx0:e.info.nx0 + e.info.nxl])   # or the local slab directly
    assert torch.equal(cube, cube2)
    b = eng.measure_triangle_sums(e, [cube], edges, triples) * syn.BOX ** 6
    ntri, kmean = eng.measure_grid_sums(e, edges, triples)
    np.save(os.path.join(outdir, f"b_{rank}.npy"), b)
    np.save(os.path.join(outdir, f"n_{rank}.npy"), ntri)
    np.save(os.path.join(outdir, f"k_{rank}.npy"), kmean)
    np.save(os.path.join(outdir, f"grid_{rank}.npy"), np.array([g.neval, g.ncrop]))
    dist.destroy_process_group()


# nb = 7: bins up to 7.5 k_f on a 16^3 mesh -> nothing can be cropped -> the spectrum is distributed in
# ky blocks and both exchanges are all-to-all transposes (engine.Engine.transposed)
@pytest.mark.parametrize("world,policy,nb", [(2, "full", 4), (2, "auto", 4), (4, "full", 4), (2, "full", 7), (4, "auto", 7)])
def test_sharded_pipeline_matches_oracle(world, policy, nb):
    from oracle import bskit_oracle as orc
    from bskit_b200 import synthetic as syn
    with tempfile.TemporaryDirectory() as d:
        init = os.path.join(d, "rendezvous")
        mp.spawn(_worker, args=(world, init, policy, d, nb), nprocs=world, join=True)
        n = 16
        kmin, kmax, dk = syn.bench_bins(nb)
        edges = orc.bin_edges(kmin, kmax, dk)
        _, idx = orc.triangles_all(edges, 1)
        mesh = syn.lognormal_mesh(n, seed=1, dtype=np.float64)
        want = orc.measure_unnormalized([mesh], syn.BOX, edges, idx)
        wn, wk = orc.measure_gridinfo(n, syn.BOX, edges, idx)
        for r in range(world):
            b = np.load(os.path.join(d, f"b_{r}.npy"))
            np.testing.assert_allclose(b, want, rtol=1e-10, atol=1e-12 * np.abs(want).max())
            assert np.array_equal(np.load(os.path.join(d, f"n_{r}.npy")), np.rint(wn))
            np.testing.assert_allclose(np.load(os.path.join(d, f"k_{r}.npy")), wk, rtol=1e-10)
            m = np.load(os.path.join(d, f"grid_{r}.npy"))[0]
            assert m % world == 0
