"""Reader (and writer) for density grids stored in the *bigfile* format.

The reference reads its production meshes with ``nbodykit.lab.BigFileMesh(path, 'Field')``
(``scripts/measure/measure_bs_slow.py:226``, ``scripts/grids/downsample_bigfile_grid.py:42``,
``scripts/measure/measure_ps_from_bigfile_subboxes.py``) and writes them with
``mesh.save(path, dataset='Field', mode='real')``.  Neither nbodykit nor the ``bigfile`` library
is a dependency here; this module restates the on-disk layout of bigfile (rainwoodman/bigfile,
``src/bigfile.c``: ``big_block_open`` / ``big_block_flush``), which is plain files:

    <path>/<dataset>/header     text:  ``DTYPE: <f4`` / ``NMEMB: 1`` / ``NFILE: 2`` and one line per
                                physical file ``000000: <items> : <checksum> : <unused>``
    <path>/<dataset>/attr-v2    text, one attribute per line:
                                ``name dtype nmemb <raw bytes in hex> #HUMANE [ v1 v2 ... ]``
    <path>/<dataset>/000000 ..  raw little/big-endian items, ``NMEMB`` members each, C order

A mesh written by nbodykit carries ``ndarray.shape`` (= Nmesh), ``Nmesh`` and ``BoxSize``
attributes; the real field is the flattened (Nx, Ny, Nz) array.  The checksum is bigfile's
``sysvsum`` (sum of all bytes, folded to 16 bits twice); it is verified when ``verify=True``.

Parity note: no bigfile produced by the reference ships in its tree and the library cannot be
installed here, so the layout is pinned to the library's documented format, to a fixture built
byte by byte in ``tests/test_bigfile.py`` and to a write/read round trip, not to a
reference-produced file.

Every rank of a multi-GPU job maps only the physical files that hold its x-slab
(``BigFileMesh.slab``): nothing is read that the rank does not own.
"""
from __future__ import annotations

import os

import numpy as np

from .mesh import ArrayMesh

_HEX = "0123456789abcdef"


def _sysvsum(buf):
    """bigfile's checksum of a byte buffer (``sysvsum`` in bigfile.c): the plain byte sum, to be
    folded by :func:`_fold` after all buffers of a physical file have been added up."""
    return int(np.frombuffer(buf, dtype=np.uint8).sum(dtype=np.uint64))


def _fold(s):
    r = (s & 0xFFFF) + ((s & 0xFFFFFFFF) >> 16)
    return (r & 0xFFFF) + (r >> 16)


def read_header(block_dir):
    """(dtype, nmemb, [items per physical file], [checksums]) of one block."""
    dtype = nmemb = nfile = None
    sizes, sums = {}, {}
    with open(os.path.join(block_dir, "header")) as f:
        for line in f:
            line = line.strip()
            if not line:
                continue
            key, _, rest = line.partition(":")
            key, rest = key.strip(), rest.strip()
            if key == "DTYPE":
                dtype = np.dtype(rest)
            elif key == "NMEMB":
                nmemb = int(rest)
            elif key == "NFILE":
                nfile = int(rest)
            else:
                parts = [x.strip() for x in rest.split(":")]
                fid = int(key, 16)
                sizes[fid] = int(parts[0])
                sums[fid] = int(parts[1]) if len(parts) > 1 and parts[1] else None
    if dtype is None or nmemb is None or nfile is None:
        raise ValueError(f"{block_dir}/header: DTYPE, NMEMB and NFILE are required")
    if sorted(sizes) != list(range(nfile)):
        raise ValueError(f"{block_dir}/header: NFILE={nfile} but file entries {sorted(sizes)}")
    return dtype, nmemb, [sizes[i] for i in range(nfile)], [sums[i] for i in range(nfile)]


def read_attrs(block_dir):
    """Attributes of a block (``attr-v2``) as a dict of numpy arrays / str."""
    out = {}
    path = os.path.join(block_dir, "attr-v2")
    if not os.path.exists(path):
        return out
    with open(path) as f:
        for line in f:
            line = line.split("#HUMANE")[0].strip()
            if not line:
                continue
            name, dt, nmemb, *rest = line.split()
            raw = bytes.fromhex(rest[0]) if rest else b""
            dt = np.dtype(dt)
            if dt.kind == "S":
                out[name] = raw[: int(nmemb) * dt.itemsize].decode("utf-8", "replace").rstrip("\0")
            else:
                out[name] = np.frombuffer(raw, dtype=dt, count=int(nmemb)).copy()
    return out


def _file_name(i):
    return "%06X" % i


class BigFileMesh(ArrayMesh):
    """A real density mesh stored in bigfile format (cf. ``nbodykit.lab.BigFileMesh``).

    ``BigFileMesh(path, dataset)`` exposes ``attrs['BoxSize']``, ``attrs['Nmesh']`` and the other
    stored attributes, and the field as a read-only memory map (one physical file) or a lazily
    assembled array (several); :meth:`slab` returns the x-planes ``[x0, x1)`` reading only the
    physical files that hold them.
    """

    def __init__(self, path, dataset="Field", verify=False, compensation=None):
        self.path, self.dataset = path, dataset
        self.block_dir = os.path.join(path, dataset)
        if not os.path.isdir(self.block_dir):
            raise FileNotFoundError(f"no dataset {dataset!r} in bigfile {path!r}")
        self.dtype, self.nmemb, self.sizes, self.checksums = read_header(self.block_dir)
        if self.nmemb != 1 or self.dtype.kind != "f":
            raise ValueError(f"{self.block_dir}: a real mesh needs NMEMB=1 and a float dtype, found "
                             f"{self.dtype.str} x {self.nmemb} (complex-mode meshes are not supported)")
        stored = read_attrs(self.block_dir)
        shape = stored.get("ndarray.shape", stored.get("Nmesh"))
        if shape is None or "BoxSize" not in stored:
            raise ValueError(f"{self.block_dir}/attr-v2: 'ndarray.shape' (or 'Nmesh') and 'BoxSize' are required")
        shape = tuple(int(v) for v in np.atleast_1d(shape))
        if len(shape) == 1:
            shape = shape * 3
        if int(np.prod(shape)) != sum(self.sizes):
            raise ValueError(f"{self.block_dir}: shape {shape} does not match {sum(self.sizes)} stored items")
        self.shape = shape
        if verify:
            self.verify()
        if len(self.sizes) == 1:
            array = np.memmap(os.path.join(self.block_dir, _file_name(0)), dtype=self.dtype, mode="r", shape=shape)
        else:
            array = _LazyField(self)
        box = np.atleast_1d(np.asarray(stored["BoxSize"], dtype=np.float64)).ravel()
        self.array = array
        self.attrs = {k: v for k, v in stored.items()}
        self.attrs["BoxSize"] = np.ones(3) * box if box.size == 1 else box.copy()
        self.attrs["Nmesh"] = np.array(shape, dtype=np.int64)
        self.compensation = compensation

    def verify(self):
        """Check every physical file against the header's checksum."""
        for i, (n, want) in enumerate(zip(self.sizes, self.checksums)):
            if want is None:
                continue
            raw = np.fromfile(os.path.join(self.block_dir, _file_name(i)), dtype=np.uint8)
            if raw.size != n * self.dtype.itemsize or _fold(_sysvsum(raw)) != want:
                raise ValueError(f"{self.block_dir}/{_file_name(i)}: size or checksum mismatch")

    def slab(self, x0, x1):
        """x-planes [x0, x1) as a native-endian array; reads only the files that overlap them."""
        plane = self.shape[1] * self.shape[2]
        lo, hi = x0 * plane, x1 * plane
        out = np.empty(hi - lo, dtype=self.dtype.newbyteorder("="))
        start = 0
        for i, n in enumerate(self.sizes):
            a, b = max(lo, start), min(hi, start + n)
            if a < b:
                mm = np.memmap(os.path.join(self.block_dir, _file_name(i)), dtype=self.dtype, mode="r", shape=(n,))
                out[a - lo:b - lo] = mm[a - start:b - start]
            start += n
        return out.reshape((x1 - x0,) + tuple(self.shape[1:]))

    def compute(self, mode="real"):
        if mode != "real":
            raise NotImplementedError("BigFileMesh only holds the real field")
        return self.array if isinstance(self.array, np.memmap) else self.slab(0, self.shape[0])

    def view(self):
        return self

    def apply(self, func, kind="circular", mode="complex"):
        from .mesh import CompensateCIC
        if not isinstance(func, CompensateCIC):
            raise NotImplementedError("only CompensateCIC actions are supported by bskit_b200")
        return BigFileMesh(self.path, self.dataset, compensation=func)


class _LazyField:
    """Array-like over several physical files: supports ``shape``, ``dtype`` and x-slicing."""

    def __init__(self, mesh):
        self._m = mesh
        self.shape = mesh.shape
        self.dtype = mesh.dtype.newbyteorder("=")
        self.ndim = 3

    def __getitem__(self, idx):
        if isinstance(idx, slice):
            x0, x1, step = idx.indices(self.shape[0])
            if step == 1:
                return self._m.slab(x0, x1)
        return self._m.slab(0, self.shape[0])[idx]

    def __array__(self, dtype=None, copy=None):
        a = self._m.slab(0, self.shape[0])
        return a.astype(dtype) if dtype is not None else a


def save_mesh(path, array, BoxSize, dataset="Field", nfile=1, attrs=None):
    """Write a real (N,N,N) mesh as ``<path>/<dataset>`` in bigfile format (what
    ``mesh.save(path, dataset='Field', mode='real')`` produces, cf.
    ``scripts/grids/convert_npygrid_to_bigfile.py:55``)."""
    array = np.ascontiguousarray(array)
    if array.ndim != 3 or array.dtype.kind != "f":
        raise ValueError("save_mesh expects a real 3-D floating point array")
    block_dir = os.path.join(path, dataset)
    os.makedirs(block_dir, exist_ok=True)
    flat = array.reshape(-1)
    nfile = int(max(1, min(nfile, flat.size)))
    bounds = [flat.size * i // nfile for i in range(nfile + 1)]
    lines = ["DTYPE: %s" % array.dtype.str, "NMEMB: 1", "NFILE: %d" % nfile]
    for i in range(nfile):
        part = flat[bounds[i]:bounds[i + 1]]
        part.tofile(os.path.join(block_dir, _file_name(i)))
        lines.append("%s: %d : %d : 0" % (_file_name(i), part.size, _fold(_sysvsum(part.view(np.uint8)))))
    with open(os.path.join(block_dir, "header"), "w") as f:
        f.write("\n".join(lines) + "\n")
    box = np.atleast_1d(np.asarray(BoxSize, dtype=np.float64)).ravel()
    box = np.ones(3) * box if box.size == 1 else box
    all_attrs = {"ndarray.shape": np.array(array.shape, dtype="<i8"), "BoxSize": box.astype("<f8"),
                 "Nmesh": np.array(array.shape, dtype="<i8")}
    for k, v in (attrs or {}).items():
        all_attrs[k] = v
    with open(os.path.join(block_dir, "attr-v2"), "w") as f:
        for name, v in all_attrs.items():
            if isinstance(v, str):
                raw, dt, n = v.encode(), "|S1", len(v.encode())
                human = v
            else:
                v = np.atleast_1d(np.asarray(v))
                raw, dt, n = v.tobytes(), v.dtype.str, v.size
                human = " ".join(repr(x) for x in v.tolist())
            f.write("%s %s %d %s #HUMANE [ %s ]\n" % (name, dt, n, raw.hex(), human))
    return path
