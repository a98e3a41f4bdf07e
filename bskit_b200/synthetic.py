"""Synthetic density meshes for tests and benchmarks.

All meshes: BoxSize 1000, white noise from ``numpy.random.default_rng(seed)``
coloured with P(k) = 2e4 (k/0.02) / (1 + (k/0.02)^2)^1.7.  No network, no
datasets: these are the "synthetic Gaussian/lognormal density grids of the named
shapes" BASELINE.json asks for.
"""
from __future__ import annotations

import numpy as np
import scipy.fft as sfft

BOX = 1000.0
KF = 2.0 * np.pi / BOX


def bench_bins(nbins, box=BOX):
    """(kmin, kmax, dk) giving exactly `nbins` bins of width k_f starting at k_f/2.
    kmax = kmin + (S+0.5) dk keeps np.arange's length robust."""
    kf = 2.0 * np.pi / box
    return 0.5 * kf, 0.5 * kf + (nbins + 0.5) * kf, kf


def _power(k):
    x = k / 0.02
    return 2.0e4 * x / (1.0 + x * x) ** 1.7


def _colour(white, box, workers=None, extra=None):
    n = white.shape[0]
    wk = sfft.rfftn(white, workers=workers)
    kx = 2 * np.pi * np.fft.fftfreq(n, 1.0 / n) / box
    kz = 2 * np.pi * np.fft.rfftfreq(n, 1.0 / n) / box
    kk = np.sqrt(kx[:, None, None] ** 2 + kx[None, :, None] ** 2 + kz[None, None, :] ** 2)
    amp = np.sqrt(_power(kk) * n ** 3 / box ** 3)
    amp[0, 0, 0] = 0.0
    if extra is not None:
        amp = amp * extra(kk)
    wk *= amp
    return sfft.irfftn(wk, s=white.shape, workers=workers)


def gaussian_mesh(n, seed=1, box=BOX, dtype=np.float32, workers=None):
    rng = np.random.default_rng(seed)
    white = rng.standard_normal((n, n, n))
    return _colour(white, box, workers).astype(dtype)


def lognormal_mesh(n, seed=1, box=BOX, dtype=np.float32, workers=None):
    g = _colour(np.random.default_rng(seed).standard_normal((n, n, n)), box, workers)
    g *= 0.8 / g.std()
    sigma2 = g.var()
    return (np.exp(g - 0.5 * sigma2) - 1.0).astype(dtype)


def baryon_like_mesh(matter, seed=2, box=BOX, dtype=np.float32, workers=None):
    """delta_b(k) = delta_m(k) exp(-(k/(40 k_f))^2 / 2) + 5% independent Gaussian."""
    n = matter.shape[0]
    kf = 2 * np.pi / box
    mk = sfft.rfftn(np.asarray(matter, dtype=np.float64), workers=workers)
    kx = 2 * np.pi * np.fft.fftfreq(n, 1.0 / n) / box
    kz = 2 * np.pi * np.fft.rfftfreq(n, 1.0 / n) / box
    kk = np.sqrt(kx[:, None, None] ** 2 + kx[None, :, None] ** 2 + kz[None, None, :] ** 2)
    mk *= np.exp(-0.5 * (kk / (40.0 * kf)) ** 2)
    smooth = sfft.irfftn(mk, s=matter.shape, workers=workers)
    noise = _colour(np.random.default_rng(seed).standard_normal((n, n, n)), box, workers)
    return (smooth + 0.05 * noise * (smooth.std() / noise.std())).astype(dtype)
