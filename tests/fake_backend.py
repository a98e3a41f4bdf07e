"""A numpy stand-in for ``engine.NativeBackend`` used ONLY by the CPU tests of the
multi-rank host logic (slab partition, all-gather of cropped planes, all-reduce).
It follows the same stage contract as the C ABI (include/bskit_b200.h) on CPU tensors.
The product never selects it: ``Engine`` defaults to the CUDA backend and has no fallback."""
import types

import numpy as np
import torch

from bskit_b200.engine import axis_tables, mode_of


class FakeBackend:
    name = "numpy-test-double"

    def __init__(self, grid, boxsize, precision, world, rank, device, max_shells,
                 fft_precision=None, accum_precision=None, no_prune=False, contraction=None, transposed=False):
        self.grid, self.world, self.rank = grid, world, rank
        self.transposed, self.group = bool(transposed), None
        n, m = grid.nmesh, grid.neval
        kxy = n if grid.full else 2 * grid.ncrop + 1
        kzn = n // 2 + 1 if grid.full else grid.ncrop + 1
        self.kx, self.ky, self.kz, self.modes = axis_tables(grid, boxsize)
        nxl, mxl = n // world, m // world
        kyl = kxy // world if self.transposed else kxy
        self.ky_loc = self.ky[rank * kyl:(rank + 1) * kyl] if self.transposed else self.ky
        self.info = types.SimpleNamespace(
            kyl=kyl, ky0=rank * kyl if self.transposed else 0, pruned=0,
            xplanes_complex_per_shell=mxl * kxy * kzn if self.transposed else 0,
            kx=kxy, ky=kxy, kz=kzn, nx0=nxl * rank, nxl=nxl, mx0=mxl * rank, mxl=mxl, fwd_batch=nxl,
            fwd_work_complex=0, planes_local_complex=nxl * kxy * kzn, planes_all_complex=n * kxy * kzn,
            cube_complex=kxy * kyl * kzn, xcols_complex_per_shell=m * kyl * kzn if self.transposed else 1,
            planes2d_complex_per_shell=mxl * m * (m // 2 + 1) if self.transposed else 1,
            field_real_per_shell=mxl * m * m, fft_work_bytes=0)
        self.comp = None
        self.rdtype = torch.float64

    def close(self):
        pass

    def set_compensation(self, tables):
        self.comp = tables

    def prepare_shells(self, nsh):
        pass

    def forward_local(self, slab):
        n = self.grid.nmesh
        spec = np.fft.rfft2(slab.numpy().astype(np.float64), axes=(1, 2)) / float(n) ** 3
        iy = self.modes % n
        out = spec[:, iy, :][:, :, :self.info.kz]
        if self.comp is not None:
            out = out * self.comp[1][None, :, None] * self.comp[2][None, None, :]
        return torch.from_numpy(np.ascontiguousarray(out))

    def forward_finish(self, planes_all):
        n = self.grid.nmesh
        spec = np.fft.fft(planes_all.numpy(), axis=0)[self.modes % n]
        if self.comp is not None:
            spec = spec * self.comp[0][:, None, None]
        return torch.from_numpy(np.ascontiguousarray(spec))

    def shells(self, cube, kind, kpow, lo, hi, xcols, planes2d, fields_out):
        m = self.grid.neval
        kk = (self.kx[:, None, None] ** 2.0 + self.ky[None, :, None] ** 2.0
              + self.kz[None, None, :] ** 2.0) ** 0.5
        for s in range(len(lo)):
            mask = (kk <= hi[s]) & (kk >= lo[s])
            src = cube.numpy() if kind == 0 else (np.ones_like(kk) if kind == 1 else kk ** kpow)
            big = np.zeros((m, m, m // 2 + 1), dtype=np.complex128)
            ix = self.modes % m
            big[np.ix_(ix, ix, np.arange(self.info.kz))] = src * mask
            cols = np.fft.ifft(big, axis=0) * m                       # un-normalised inverse
            loc = cols[self.info.mx0:self.info.mx0 + self.info.mxl]
            real = np.fft.irfft2(loc, s=(m, m), axes=(1, 2)) * float(m) ** 2
            fields_out[s].copy_(torch.from_numpy(real.reshape(-1)))

    # transposed plans (uncropped spectrum, ky blocks): the two halves of `shells`
    def shells_x(self, cube, kind, kpow, lo, hi, xcols):
        m, f = self.grid.neval, self.info
        kk = (self.kx[:, None, None] ** 2.0 + self.ky_loc[None, :, None] ** 2.0
              + self.kz[None, None, :] ** 2.0) ** 0.5
        out = xcols[:len(lo) * f.xcols_complex_per_shell].view(m, len(lo), f.kyl, f.kz)
        for s in range(len(lo)):
            mask = (kk <= hi[s]) & (kk >= lo[s])
            src = cube.numpy() if kind == 0 else (np.ones_like(kk) if kind == 1 else kk ** kpow)
            out[:, s] = torch.from_numpy(np.fft.ifft(src * mask, axis=0) * m)

    def shells_yz(self, nsh, xplanes, planes2d, fields_out):
        m, f = self.grid.neval, self.info
        loc = xplanes[:nsh * f.xplanes_complex_per_shell].view(f.mxl, nsh, f.ky, f.kz).numpy()
        for s in range(nsh):
            real = np.fft.irfft2(loc[:, s], s=(m, m), axes=(1, 2)) * float(m) ** 2
            fields_out[s].copy_(torch.from_numpy(real.reshape(-1)))

    def modes_per_bin(self, lo, hi):
        import torch.distributed as dist
        kk = (self.kx[:, None, None] ** 2.0 + self.ky_loc[None, :, None] ** 2.0 + self.kz[None, None, :] ** 2.0) ** 0.5
        n = self.grid.nmesh
        w = np.where((np.arange(self.info.kz) == 0) | (np.arange(self.info.kz) == n // 2), 1, 2)[None, None, :]
        out = np.array([int((w * ((kk <= h) & (kk >= l))).sum()) for l, h in zip(lo, hi)], dtype=np.int64)
        if self.transposed and self.world > 1:
            t = torch.from_numpy(out)
            dist.all_reduce(t, group=self.group)
            out = t.numpy()
        return out

    def contract(self, fields, rows, ncells, job_off):
        t = [f.numpy() for f in fields]
        out = np.zeros((len(job_off), len(rows)))
        for j, off in enumerate(np.asarray(job_off).reshape(-1, 3)):
            for i, (a, b, c) in enumerate(np.asarray(rows)):
                out[j, i] = np.sum(t[a + off[0]] * t[b + off[1]] * t[c + off[2]])
        return torch.from_numpy(out)
