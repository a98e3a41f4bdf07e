"""Mesh sources accepted by :class:`bskit_b200.main.FFTBispectrum`.

The reference takes nbodykit ``MeshSource`` objects (``BigFileMesh``,
``ArrayMesh``, ``FieldMesh`` ...; ``bskit/main.py:1509-1524``).  nbodykit is not
a dependency here: a source is anything that yields a real (N,N,N) density
array plus ``attrs['BoxSize']`` / ``attrs['Nmesh']``:

* a numpy array or torch tensor (``BoxSize`` must then be passed explicitly);
* an :class:`ArrayMesh` (array + BoxSize, optional CIC compensation);
* any object with ``.attrs`` and one of ``.compute(mode='real')``,
  ``.paint(mode='real')`` or ``.preview()`` returning array-likes — which is
  what nbodykit meshes expose, so they can be passed straight through when
  nbodykit is installed.
"""
from __future__ import annotations

import numpy as np

try:
    import torch
except Exception:  # pragma: no cover
    torch = None


class CompensateCIC:
    """CIC window compensation at the original painting resolution
    (ref. ``CompensateCICShotnoiseNgrid``, scripts/measure/measure_bs_fast.py:45-57):
    ``v / prod_i (1 - 2/3 sin^2(w_i Nmesh / (2 NmeshCIC)))^(1/2)``.  Applied on the
    GPU as three per-axis tables inside the forward transform."""

    def __init__(self, nmesh_cic):
        self.nmesh_cic = int(nmesh_cic)

    def __repr__(self):
        return f"CompensateCIC(nmesh_cic={self.nmesh_cic})"


class ArrayMesh:
    """A real density mesh held in memory (cf. ``nbodykit.lab.ArrayMesh``)."""

    def __init__(self, array, BoxSize, compensation=None):
        shape = tuple(array.shape)
        # (N,N,N), or this rank's x-slab (N/world, N, N) of a mesh sharded over ranks
        if len(shape) != 3 or shape[1] != shape[2] or shape[0] > shape[1] or shape[1] % shape[0]:
            raise ValueError(f"mesh must be a cubic 3-D array (or an x-slab of one), got shape {shape}")
        self.array = array
        box = np.atleast_1d(np.asarray(BoxSize, dtype=np.float64)).ravel()
        self.attrs = {"BoxSize": np.ones(3) * box if box.size == 1 else box.copy(),
                      "Nmesh": np.array([shape[1]] * 3, dtype=np.int64)}
        self.compensation = compensation

    @classmethod
    def from_npy(cls, path, BoxSize, mmap=True):
        """A mesh stored as a .npy array (cf. scripts/grids/convert_npy_grid_to_bigfile.py)."""
        return cls(np.load(path, mmap_mode="r" if mmap else None), BoxSize)

    def apply(self, func, kind="circular", mode="complex"):
        """Queue a k-space action like nbodykit's ``mesh.apply``.  Only the CIC
        compensation object is supported (the one action the reference's scripts use)."""
        if not isinstance(func, CompensateCIC):
            raise NotImplementedError("only CompensateCIC actions are supported by bskit_b200")
        return ArrayMesh(self.array, self.attrs["BoxSize"], compensation=func)

    def view(self):
        return ArrayMesh(self.array, self.attrs["BoxSize"], compensation=self.compensation)

    def compute(self, mode="real"):
        if mode != "real":
            raise NotImplementedError("ArrayMesh only stores the real field")
        return self.array


def cast_source(source, Nmesh=None, BoxSize=None):
    """Normalise a user source to an :class:`ArrayMesh` (cf. nbodykit ``_cast_source``)."""
    if isinstance(source, ArrayMesh):
        mesh = source
    elif isinstance(source, np.ndarray) or (torch is not None and isinstance(source, torch.Tensor)):
        if BoxSize is None:
            raise ValueError("BoxSize is required when the source is a bare array")
        mesh = ArrayMesh(source, BoxSize)
    elif hasattr(source, "attrs"):
        arr = None
        for name, kw in (("compute", {"mode": "real"}), ("paint", {"mode": "real"}), ("preview", {})):
            fn = getattr(source, name, None)
            if fn is not None:
                arr = np.asarray(fn(**kw))
                break
        if arr is None:
            raise TypeError("mesh source exposes none of compute/paint/preview")
        mesh = ArrayMesh(arr, source.attrs["BoxSize"] if BoxSize is None else BoxSize)
    else:
        raise TypeError(f"cannot interpret {type(source).__name__} as a mesh source")
    n = int(mesh.attrs["Nmesh"][0])
    if Nmesh is not None:
        want = np.atleast_1d(np.asarray(Nmesh)).ravel()
        if not np.all(want == n):
            raise ValueError(f"Nmesh={Nmesh} does not match the mesh resolution {n} "
                             "(resampling is out of scope for the bispectrum path)")
    if BoxSize is not None:
        box = np.atleast_1d(np.asarray(BoxSize, dtype=np.float64)).ravel()
        box = np.ones(3) * box if box.size == 1 else box
        if not np.array_equal(box, mesh.attrs["BoxSize"]):
            raise ValueError("BoxSize does not match the mesh source")
    return mesh
