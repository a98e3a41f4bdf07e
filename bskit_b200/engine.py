"""Host-side driver of the GPU bispectrum pipeline.

PyTorch is plumbing here (device memory, streams, ``torch.distributed``); every
numeric stage is a call into the C-ABI CUDA library (``_native``):

    forward   : x-slab 2-D r2c -> crop/compensate -> all-gather (NCCL) -> x c2c -> cube
    shells    : k-shell filter -> inverse x c2c -> scatter -> batched 2-D c2r
    contract  : all triangles of a list in one pass over the shell fields
    all-reduce: one NCCL all-reduce of the float64 triangle sums

It replaces the two hot loops of the reference
(``bskit/main.py:1846-1879`` and ``2006-2061``).  One process drives one GPU;
under ``torchrun`` every rank owns ``N/world`` x-planes of the mesh and of every
shell field, and all calls are collective.
"""
from __future__ import annotations

import ctypes as C
import os
import math
from dataclasses import dataclass

import numpy as np
import torch

from . import _native as nat

try:  # torch.distributed is optional plumbing: a single process needs none of it
    import torch.distributed as dist
except Exception:  # pragma: no cover
    dist = None

TWO_PI = 2.0 * np.pi


# --------------------------------------------------------------------------- #
# small helpers
# --------------------------------------------------------------------------- #
def dist_info(group=None):
    if dist is not None and dist.is_available() and dist.is_initialized():
        return dist.get_world_size(group), dist.get_rank(group)
    return 1, 0


def box3(box):
    b = np.atleast_1d(np.asarray(box, dtype=np.float64)).ravel()
    if b.size == 1:
        b = np.ones(3) * b[0]
    if b.size != 3:
        raise ValueError("BoxSize must be a scalar or have 3 elements")
    return b


def _good_size(n, multiple):
    """Smallest even m >= n, divisible by `multiple`, with only factors 2,3,5,7 (cuFFT-friendly)."""
    m = max(int(n), 4)
    step = int(np.lcm(2, multiple))
    m = ((m + step - 1) // step) * step
    while True:
        r = m
        for f in (2, 3, 5, 7):
            while r % f == 0:
                r //= f
        if r == 1:
            break
        m += step
    # a power of two within 15% is preferred: the pruned z-pass kernel handles those
    p2 = 1 << (int(n) - 1).bit_length()
    if p2 % step == 0 and p2 <= 1.15 * m:
        return p2
    return m


def mode_of(j, k, n):
    """Signed mode number of cropped-cube index j (mirrors csrc/common.cuh)."""
    if k == n:
        return j if j <= n // 2 - (1 if n % 2 == 0 else 0) else j - n
    nc = (k - 1) // 2
    return j if j <= nc else j - k


@dataclass(frozen=True)
class GridChoice:
    nmesh: int
    neval: int      # M: grid the shells are synthesised on
    ncrop: int      # modes with |n_axis| <= ncrop are kept
    full: bool      # the whole half-spectrum is kept (no crop)


def choose_grid(nmesh, boxsize, kmax, policy="auto", world=1):
    """Pick the crop radius and the evaluation grid for shells up to |k| <= kmax.

    ``policy='full'``: shells live on the mesh's own N^3 grid, exactly the
    transforms the reference performs.  ``policy='auto'``: when every shell is
    band-limited to |n_axis| <= n_c with 3*n_c < N, the triple-product sums are
    evaluated on the smallest cuFFT-friendly grid M > 3*n_c instead.  That is
    algebraically exact: sum_x I_a I_b I_c / N^3 only couples modes with
    k1+k2+k3 = 0 (mod N), and with |k1+k2+k3| <= 3*n_c < M <= N "mod M" and
    "mod N" select the same triangles (results are grid-size
    independent while 3*n_max < N).  An integer policy forces that M.
    """
    n = int(nmesh)
    L = box3(boxsize)
    # |n_i| <= k L_i / 2pi for every mode inside a shell; +1 guards fp rounding
    ncrop = int(math.floor(float(kmax) * float(L.max()) / TWO_PI)) + 1
    if 2 * ncrop + 1 >= n:
        return GridChoice(n, n, n // 2, True)
    if policy == "full":
        return GridChoice(n, n, ncrop, False)
    if policy == "auto":
        m = _good_size(3 * ncrop + 1, world)
        return GridChoice(n, m if m < n else n, ncrop, False)
    m = int(policy)
    if m > n or m % 2 or m % world or m < 3 * ncrop + 1:
        raise ValueError(f"evaluation grid {m} is invalid for nmesh={n}, ncrop={ncrop}, world={world}")
    return GridChoice(n, m, ncrop, False)


def axis_tables(grid: GridChoice, boxsize):
    """float64 wavenumber per cropped-cube index, computed with the same numpy
    expressions as the reference grid (2*pi*fftfreq(N,1/N)/L) so the in-kernel
    |k| is bit-identical to numpy's."""
    n = grid.nmesh
    L = box3(boxsize)
    kfull = [TWO_PI * np.fft.fftfreq(n, 1.0 / n) / L[a] for a in range(3)]
    khalf = TWO_PI * np.fft.rfftfreq(n, 1.0 / n) / L[2]
    kxy = n if grid.full else 2 * grid.ncrop + 1
    kzn = n // 2 + 1 if grid.full else grid.ncrop + 1
    modes = np.array([mode_of(j, kxy, n) for j in range(kxy)])
    kx = np.ascontiguousarray(kfull[0][modes % n])
    ky = np.ascontiguousarray(kfull[1][modes % n])
    kz = np.ascontiguousarray(khalf[:kzn])
    return kx, ky, kz, modes


def cic_tables(grid: GridChoice, nmesh_cic):
    """Per-axis CIC compensation factors (ref. scripts/measure/measure_bs_fast.py:45-57):
    v / (1 - 2/3 sin^2(w N / (2 N_cic)))^(1/2), w the circular frequency."""
    n = grid.nmesh
    kxy = n if grid.full else 2 * grid.ncrop + 1
    kzn = n // 2 + 1 if grid.full else grid.ncrop + 1
    w_full = TWO_PI * np.fft.fftfreq(n)
    w_half = TWO_PI * np.fft.rfftfreq(n)
    modes = np.array([mode_of(j, kxy, n) for j in range(kxy)])

    def comp(w):
        return 1.0 / (1.0 - 2.0 / 3.0 * np.sin(0.5 * w * n / nmesh_cic) ** 2) ** 0.5

    cxy = np.ascontiguousarray(comp(w_full[modes % n]))
    return cxy, cxy.copy(), np.ascontiguousarray(comp(w_half[:kzn]))


# --------------------------------------------------------------------------- #
# native backend: thin, typed wrappers over the C ABI
# --------------------------------------------------------------------------- #
class NativeBackend:
    """Owns one ``bsk_plan`` and launches its stages on the current CUDA stream."""

    name = "cuda-sm100a"

    def __init__(self, grid: GridChoice, boxsize, precision, world, rank, device, max_shells,
                 fft_precision=None, accum_precision=None, no_prune=False, contraction=None, transposed=False):
        if device.type != "cuda":
            raise nat.NativeError("bskit_b200 runs on CUDA devices only (no CPU fallback)")
        self.transposed = bool(transposed)
        self.world, self.group = world, None        # the engine sets `group` (collective mode counts)
        self.lib = nat.lib()
        self.grid, self.precision, self.device = grid, precision, device
        self.fft_precision = nat.F64 if fft_precision is None else max(fft_precision, precision)
        self.accum_precision = precision if accum_precision is None else max(accum_precision, precision)
        self.rdtype = torch.float32 if precision == nat.F32 else torch.float64
        self.cdtype = torch.complex64 if self.fft_precision == nat.F32 else torch.complex128
        kx, ky, kz, _ = axis_tables(grid, boxsize)
        self._tables = (kx, ky, kz)
        geom = nat.Geometry(grid.nmesh, grid.neval, grid.ncrop, precision, world, rank, max_shells,
                            self.fft_precision, int(bool(no_prune)), int(self.transposed))
        self.stream = torch.cuda.current_stream(device).cuda_stream
        handle = C.c_void_p()
        with torch.cuda.device(device):
            nat.check(self.lib.bsk_plan_create(C.byref(handle), C.byref(geom), nat.dptr(kx),
                                               nat.dptr(ky), nat.dptr(kz), C.c_void_p(self.stream)),
                      "bsk_plan_create")
        self.handle = handle
        self.info = nat.Info()
        nat.check(self.lib.bsk_plan_info(handle, C.byref(self.info)), "bsk_plan_info")
        self._cplans = {}           # insertion-ordered: least recently used first
        #: 1 = tcgen05 tensor cores for eligible dense lists (3xTF32), 0 = FP32-pipe tile kernel
        #: everywhere (exact round-to-nearest products).  Chosen by the `contraction` argument
        #: ('tensor' | 'fp32'), else BSKIT_B200_CONTRACTION, else tensor.
        #: include/bskit_b200.h, bsk_cplan_set_path
        if contraction is None:
            contraction = os.environ.get("BSKIT_B200_CONTRACTION", "") or "tensor"
        if contraction not in ("tensor", "fp32"):
            raise ValueError("contraction must be 'tensor' or 'fp32'")
        self.contraction_path = 0 if contraction == "fp32" else 1
        self.last_path = 0

    MAX_CPLANS = 8

    def _use_current_stream(self):
        """Every stage runs on the stream that is current when it is called (a plan built under
        one stream may be used under another: sessions are cached across objects)."""
        st = torch.cuda.current_stream(self.device).cuda_stream
        if st != self.stream:
            nat.check(self.lib.bsk_plan_set_stream(self.handle, C.c_void_p(st)), "bsk_plan_set_stream")
            self.stream = st

    def close(self):
        if getattr(self, "handle", None):
            for cp, _ in self._cplans.values():
                self.lib.bsk_cplan_destroy(cp)
            self._cplans = {}
            self.lib.bsk_plan_destroy(self.handle)
            self.handle = None

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass

    def set_compensation(self, tables):
        if tables is None and not getattr(self, "_has_comp", False):
            return                                   # already all ones: skip the upload + sync
        self._use_current_stream()
        self._has_comp = tables is not None
        cx, cy, cz = tables if tables is not None else (None, None, None)
        nat.check(self.lib.bsk_set_compensation(self.handle, nat.dptr(cx), nat.dptr(cy), nat.dptr(cz)),
                  "bsk_set_compensation")

    def forward_local(self, slab):
        """slab: CUDA tensor [nxl][N][N] f32/f64 -> planes_local [nxl][Ky][Kz] complex128."""
        self._use_current_stream()
        f, n = self.info, self.grid.nmesh
        work = torch.empty(f.fwd_work_complex, dtype=torch.complex128, device=self.device)
        conv = None
        mesh_dtype = nat.F32 if slab.dtype == torch.float32 else nat.F64
        if mesh_dtype == nat.F32:
            conv = torch.empty(f.fwd_batch * n * n, dtype=torch.float64, device=self.device)
        planes = torch.empty((f.nxl, f.ky, f.kz), dtype=torch.complex128, device=self.device)
        nat.check(self.lib.bsk_forward_local(self.handle, slab.data_ptr(), mesh_dtype, work.data_ptr(),
                                             conv.data_ptr() if conv is not None else None,
                                             planes.data_ptr()), "bsk_forward_local")
        return planes

    def forward_finish(self, planes_all):
        self._use_current_stream()
        f = self.info
        cube = torch.empty((f.kx, f.kyl, f.kz), dtype=torch.complex128, device=self.device)
        nat.check(self.lib.bsk_forward_finish(self.handle, planes_all.data_ptr(), cube.data_ptr()),
                  "bsk_forward_finish")
        return cube

    def modes_per_bin(self, lo, hi):
        self._use_current_stream()
        lo = np.ascontiguousarray(lo, dtype=np.float64)
        hi = np.ascontiguousarray(hi, dtype=np.float64)
        out = np.zeros(len(lo), dtype=np.int64)
        nat.check(self.lib.bsk_modes_per_bin(self.handle, len(lo), nat.dptr(lo), nat.dptr(hi),
                                             out.ctypes.data_as(C.POINTER(C.c_int64))),
                  "bsk_modes_per_bin")
        if self.transposed and self.world > 1:       # every rank counted its ky block of the cube
            t = torch.from_numpy(out).to(self.device)
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
            out = t.cpu().numpy()
        return out

    def _retry_after_cache_release(self, call, what):
        """cuFFT allocates its work areas with cudaMalloc, outside torch's caching allocator: when a plan
        cannot be created because freed field tables are still cached, release them and try once more."""
        rc = call()
        if rc != 0 and self.device.type == "cuda":
            torch.cuda.synchronize(self.device)
            torch.cuda.empty_cache()
            rc = call()
        nat.check(rc, what)

    def prepare_shells(self, nsh):
        self._use_current_stream()
        with torch.cuda.device(self.device):
            self._retry_after_cache_release(lambda: self.lib.bsk_shells_prepare(self.handle, int(nsh)),
                                            "bsk_shells_prepare")

    def shells(self, cube, kind, kpow, lo, hi, xcols, planes2d, fields_out):
        self._use_current_stream()
        lo = np.ascontiguousarray(lo, dtype=np.float64)
        hi = np.ascontiguousarray(hi, dtype=np.float64)
        self._retry_after_cache_release(
            lambda: self.lib.bsk_shells(self.handle, cube.data_ptr() if cube is not None else None,
                                        kind, float(kpow), len(lo), nat.dptr(lo), nat.dptr(hi),
                                        xcols.data_ptr(), planes2d.data_ptr(), fields_out.data_ptr()),
            "bsk_shells")

    def shells_x(self, cube, kind, kpow, lo, hi, xcols):
        """First half of `shells` (transposed plans): filter of this rank's ky block + inverse x transform."""
        self._use_current_stream()
        lo = np.ascontiguousarray(lo, dtype=np.float64)
        hi = np.ascontiguousarray(hi, dtype=np.float64)
        nat.check(self.lib.bsk_shells_x(self.handle, cube.data_ptr() if cube is not None else None,
                                        kind, float(kpow), len(lo), nat.dptr(lo), nat.dptr(hi),
                                        xcols.data_ptr()), "bsk_shells_x")

    def shells_yz(self, nsh, xplanes, planes2d, fields_out):
        """Second half: (y,z) transforms of this rank's x-planes (all ky) into the fields."""
        self._use_current_stream()
        nat.check(self.lib.bsk_shells_yz(self.handle, int(nsh), xplanes.data_ptr(), planes2d.data_ptr(),
                                         fields_out.data_ptr()), "bsk_shells_yz")

    def contract(self, fields, rows, ncells, job_off):
        """fields: list of 1-D field tensors (len % 4 == 0; entries may alias); rows: (T,3)
        int32 triples of indices into it.  Returns this rank's float64 sums [njobs][T] (CUDA)."""
        nrows = len(fields)
        row_ptrs = [f.data_ptr() for f in fields]
        rows = np.ascontiguousarray(rows, dtype=np.int32)
        job_off = np.ascontiguousarray(job_off, dtype=np.int32).reshape(-1, 3)
        njobs = len(job_off)
        key = (rows.tobytes(), nrows)
        ent = self._cplans.pop(key, None)           # re-inserted below: most recently used last
        if ent is None or ent[1] < njobs:
            if ent is not None:
                self.lib.bsk_cplan_destroy(ent[0])
            while len(self._cplans) >= self.MAX_CPLANS:      # schedules hold device tables and partials
                old = self._cplans.pop(next(iter(self._cplans)))
                self.lib.bsk_cplan_destroy(old[0])
            cp = C.c_void_p()
            with torch.cuda.device(self.device):
                nat.check(self.lib.bsk_cplan_create(C.byref(cp), len(rows),
                                                    rows.ctypes.data_as(C.POINTER(C.c_int32)),
                                                    nrows, max(njobs, 4)), "bsk_cplan_create")
            ent = (cp, max(njobs, 4))
        self._cplans[key] = ent
        cp = ent[0]
        self._last_cplan = cp
        nat.check(self.lib.bsk_cplan_set_path(cp, int(self.contraction_path)), "bsk_cplan_set_path")
        sums = torch.empty((njobs, len(rows)), dtype=torch.float64, device=self.device)
        ptrs = (C.c_void_p * nrows)(*row_ptrs)
        with torch.cuda.device(self.device):
            nat.check(self.lib.bsk_contract(cp, ptrs, self.precision, self.accum_precision, ncells, njobs,
                                            job_off.ctypes.data_as(C.POINTER(C.c_int32)),
                                            C.cast(sums.data_ptr(), C.POINTER(C.c_double)),
                                            C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)),
                      "bsk_contract")
        out = (C.c_int64 * 3)()
        nat.check(self.lib.bsk_cplan_path(cp, out), "bsk_cplan_path")
        self.last_path = int(out[2])
        return sums

    def fold_even(self, table, neval, mxl, fold_x):
        """Octant (or y-z quadrant) of even fields with the cube-root multiplicity weights
        (bsk_fold_even).  table: [nrows][mxl*M*M] contiguous; returns [nrows][ncell_out]."""
        h = neval // 2 + 1
        nx = h if fold_x else mxl
        ncell_out = (nx * h * h + 3) // 4 * 4
        out = torch.empty((table.shape[0], ncell_out), dtype=table.dtype, device=table.device)
        with torch.cuda.device(self.device):
            nat.check(self.lib.bsk_fold_even(C.c_void_p(table.data_ptr()),
                                             nat.F32 if table.dtype == torch.float32 else nat.F64,
                                             int(table.shape[0]), int(neval), int(mxl), int(bool(fold_x)),
                                             C.c_void_p(out.data_ptr()), ncell_out,
                                             C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)),
                      "bsk_fold_even")
        return out

    def reduce_list(self, fields, rows, ncells):
        """Streaming per-triangle reduction for sparse lists; rows: (T,3) indices into fields.
        Returns this rank's float64 sums [T] (CUDA tensor)."""
        rows = np.ascontiguousarray(rows, dtype=np.int32)
        nrows = len(fields)
        ptrs = (C.c_void_p * nrows)(*[f.data_ptr() for f in fields])
        sums = torch.empty(len(rows), dtype=torch.float64, device=self.device)
        with torch.cuda.device(self.device):
            nat.check(self.lib.bsk_reduce_list(ptrs, nrows, self.precision, ncells, len(rows),
                                               rows.ctypes.data_as(C.POINTER(C.c_int32)),
                                               C.cast(sums.data_ptr(), C.POINTER(C.c_double)),
                                               C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)),
                      "bsk_reduce_list")
        return sums

    def cplan_info(self):
        """Schedule of the most recent contraction: blocks, split, rounds, threads."""
        cp = getattr(self, "_last_cplan", None)
        if cp is None:
            return None
        out = (C.c_int64 * 4)()
        nat.check(self.lib.bsk_cplan_info(cp, out), "bsk_cplan_info")
        return dict(nblocks=int(out[0]), split=int(out[1]), rounds=int(out[2]), threads=int(out[3]))


# --------------------------------------------------------------------------- #
# engine: one grid geometry, collectives, buffers
# --------------------------------------------------------------------------- #
class Engine:
    """Shell synthesis + triangle contraction on one (N, M, crop, precision) geometry."""

    def __init__(self, grid: GridChoice, boxsize, precision=nat.F32, device=None, group=None,
                 backend_cls=None, scratch_bytes=None, fft_precision=None,
                 accum_precision=None, no_prune=False, max_rows=None, contraction=None):
        self.grid = grid
        self.max_rows = max_rows
        self.last_batches = 0
        self.last_schedule = None
        self.boxsize = box3(boxsize)
        self.precision = precision
        self.group = group
        self.world, self.rank = dist_info(group)
        if device is None:
            device = torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() \
                else torch.device("cpu")
        self.device = torch.device(device)
        self.itemsize = 4 if precision == nat.F32 else 8
        fft_precision = nat.F64 if fft_precision is None else max(fft_precision, precision)
        fft_itemsize = 4 if fft_precision == nat.F32 else 8
        m = grid.neval
        mxl = m // self.world
        kxy = grid.nmesh if grid.full else 2 * grid.ncrop + 1
        kzn = grid.nmesh // 2 + 1 if grid.full else grid.ncrop + 1
        pruned = (fft_precision == nat.F64 and not no_prune and m in (64, 128, 256, 512, 1024, 2048))
        per_shell = (m * kxy * kzn + mxl * m * (kzn if pruned else m // 2 + 1)) * 2 * fft_itemsize
        if scratch_bytes is None:
            scratch_bytes = 6 << 30
        # keep each cuFFT batch below 2^31 elements and the scratch within budget
        by_elems = (2 ** 31 - 1) // max(mxl * m * m, 1)
        self.chunk = int(max(1, min(nat.MAX_CHUNK, scratch_bytes // max(per_shell, 1), by_elems)))
        # Exchange of the distributed transforms.  A spectrum that can be cropped is all-gathered (a few
        # MB) and every rank keeps the whole cropped cube.  When the bins reach Nyquist nothing can be
        # cropped: the cube is then distributed in ky blocks and the exchanges are the all-to-all
        # transposes of a slab-decomposed FFT (what pfft does under the reference, main.py:1612).
        # BSKIT_B200_EXCHANGE=allgather|alltoall overrides (A/B timing, single-GPU tests of the split path).
        mode = os.environ.get("BSKIT_B200_EXCHANGE", "")
        if mode not in ("", "allgather", "alltoall"):
            raise ValueError("BSKIT_B200_EXCHANGE must be 'allgather' or 'alltoall'")
        self.transposed = bool(grid.full and (mode == "alltoall" or (self.world > 1 and mode != "allgather")))
        if self.transposed:
            kzn_t, kyl_t = grid.nmesh // 2 + 1, grid.nmesh // self.world
            per_shell = (m * kyl_t * kzn_t + mxl * m * (kzn_t if pruned else m // 2 + 1)) * 2 * fft_itemsize
            self.chunk = int(max(1, min(nat.MAX_CHUNK, scratch_bytes // max(per_shell, 1), by_elems)))
        extra = {} if contraction is None else {"contraction": contraction}
        if self.transposed:
            extra["transposed"] = True
        # the CUDA backend unless the caller brings its own implementation of the stage interface
        make_backend = backend_cls or NativeBackend
        self.backend = make_backend(grid, self.boxsize, precision, self.world, self.rank,
                                    self.device, self.chunk, fft_precision=fft_precision,
                                    accum_precision=accum_precision, no_prune=no_prune, **extra)
        self.backend.group = self.group
        self.info = self.backend.info
        self.ncells = int(self.info.field_real_per_shell)
        self.rdtype = torch.float32 if precision == nat.F32 else torch.float64
        self.cdtype = torch.complex64 if fft_precision == nat.F32 else torch.complex128
        self._scratch = None

    # -- forward ------------------------------------------------------------ #
    def local_slab(self, mesh):
        """This rank's x-planes of `mesh` as a CUDA tensor.  `mesh` is the full
        (N,N,N) array (numpy / torch, any device) or already the local slab."""
        n, f = self.grid.nmesh, self.info
        if not isinstance(mesh, np.ndarray) and not torch.is_tensor(mesh) and hasattr(mesh, "shape"):
            # file-backed sources (bigfile.BigFileMesh over several physical files): x-slicing reads
            # only the planes this rank owns
            if tuple(mesh.shape) == (n, n, n):
                mesh = mesh[f.nx0:f.nx0 + f.nxl]
            mesh = np.asarray(mesh)
        if isinstance(mesh, np.ndarray):
            if tuple(mesh.shape) == (n, n, n):
                mesh = mesh[f.nx0:f.nx0 + f.nxl]          # slice first: memory-mapped files stay lazy
            if mesh.dtype not in (np.float32, np.float64) or not mesh.dtype.isnative:
                mesh = mesh.astype(np.float32 if mesh.dtype.itemsize == 4 and mesh.dtype.kind == "f" else np.float64)
            if not mesh.flags.writeable:                  # e.g. np.load(mmap_mode='r')
                mesh = np.array(mesh)
            t = torch.from_numpy(mesh)
        else:
            t = mesh
            if t.dtype not in (torch.float32, torch.float64):
                t = t.to(torch.float64)
        if tuple(t.shape) == (n, n, n):
            t = t[f.nx0:f.nx0 + f.nxl]
        elif tuple(t.shape) != (f.nxl, n, n):
            raise ValueError(f"mesh has shape {tuple(t.shape)}, expected {(n, n, n)} or the local "
                             f"slab {(f.nxl, n, n)}")
        if t.device != self.device:
            if t.device.type == "cpu" and self.device.type == "cuda":
                t = t.contiguous()
                t = (t if t.is_pinned() else t.pin_memory()).to(self.device, non_blocking=True)
            else:
                t = t.to(self.device)
        return t.contiguous()

    def upload_async(self, mesh):
        """Start the host-to-device copy of this rank's slab on a side stream and return
        (device slab, event).  The copy engine then runs beside whatever the compute stream is
        doing (e.g. the mesh-independent normalisation); `forward` waits for the event."""
        if self.device.type != "cuda":
            return self.local_slab(mesh), None
        if not hasattr(self, "_copy_stream"):
            self._copy_stream = torch.cuda.Stream(self.device)
        # device inputs that need a cast / contiguous copy are read by a kernel on the copy stream:
        # it must see what the caller's stream has written
        self._copy_stream.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(self._copy_stream):
            slab = self.local_slab(mesh)
            ev = torch.cuda.Event()
            ev.record(self._copy_stream)
        return slab, ev

    def forward(self, mesh, compensation=None, ready=None):
        """delta_k cube of `mesh` (1/N^3-normalised, compensated, cropped).  `ready`: event of
        an `upload_async` whose slab is passed as `mesh`."""
        self.backend.set_compensation(compensation)
        if ready is not None:
            torch.cuda.current_stream(self.device).wait_event(ready)
        slab = self.local_slab(mesh)
        if ready is not None:
            slab.record_stream(torch.cuda.current_stream(self.device))
        planes = self.backend.forward_local(slab)
        if self.transposed:
            # x-slabs -> y-slabs: ky block r of every plane goes to rank r; the received blocks, in
            # rank order, are the planes [N][kyl][kz] of this rank's ky block
            f = self.info
            if self.world > 1:
                send = planes.view(f.nxl, self.world, f.kyl, f.kz).permute(1, 0, 2, 3).contiguous()
                allp = torch.empty_like(send)
                all_to_all_blocks(torch.view_as_real(allp), torch.view_as_real(send), self.group)
                allp = allp.view(self.grid.nmesh, f.kyl, f.kz)
            else:
                allp = planes
        elif self.world > 1:
            n = self.grid.nmesh
            allp = torch.empty((n,) + tuple(planes.shape[1:]), dtype=planes.dtype, device=planes.device)
            all_gather_concat(torch.view_as_real(allp), torch.view_as_real(planes), self.group)
        else:
            allp = planes
        return self.backend.forward_finish(allp)

    # -- shells --------------------------------------------------------------- #
    def _get_scratch(self, nsh):
        f = self.info
        need = (nsh * f.xcols_complex_per_shell, nsh * f.planes2d_complex_per_shell)
        if self._scratch is None or self._scratch[0].numel() < need[0] or self._scratch[1].numel() < need[1]:
            self._scratch = None
            self._scratch = (torch.empty(need[0], dtype=self.cdtype, device=self.device),
                             torch.empty(need[1], dtype=self.cdtype, device=self.device))
        return self._scratch

    def release_scratch(self):
        self._scratch = None

    def set_chunk(self, nsh):
        """Change how many shells are synthesised per call (the scratch scales with it)."""
        self.chunk = int(max(1, min(nsh, self.chunk)))
        self._scratch = None
        self._row_cap = None
        if self.device.type == "cuda":
            torch.cuda.empty_cache()

    def row_capacity(self):
        """How many shell fields ([ncells] of the storage dtype) fit in device memory next to
        the synthesis scratch.  `max_rows` (tests) overrides the measurement."""
        if self.max_rows is not None:
            return int(self.max_rows)
        if self.device.type != "cuda":
            return 1 << 30
        if getattr(self, "_row_cap", None) is not None:
            return self._row_cap                     # measured once per engine (cudaMemGetInfo syncs)
        # allocate the synthesis scratch and the cuFFT work areas first, then see what is left
        self._get_scratch(self.chunk)
        self.backend.prepare_shells(self.chunk)
        torch.cuda.synchronize(self.device)
        free, _ = torch.cuda.mem_get_info(self.device)
        cached = torch.cuda.memory_reserved(self.device) - torch.cuda.memory_allocated(self.device)
        avail = 0.94 * (free + cached) - (1 << 30)
        cap = max(0, int(avail // (self.ncells * self.itemsize)))
        if self.world > 1:      # every rank must cut the triangle list into the same batches
            t = torch.tensor([cap], dtype=torch.int64, device=self.device)
            dist.all_reduce(t, op=dist.ReduceOp.MIN, group=self.group)
            cap = int(t.item())
        self._row_cap = cap
        return self._row_cap

    def synthesize(self, cube, kind, kpow, lo, hi, out):
        """Fill out[i] (i over bins) with the shell field of bin (lo[i], hi[i]).
        `out` is a [nbins][ncells] CUDA tensor (rows may be a view of a larger table)."""
        nb = len(lo)
        assert out.shape[0] == nb and out.shape[1] == self.ncells and out.is_contiguous()
        for s0 in range(0, nb, self.chunk):
            s1 = min(nb, s0 + self.chunk)
            xcols, planes2d = self._get_scratch(s1 - s0)
            if not self.transposed:
                self.backend.shells(cube, kind, kpow, lo[s0:s1], hi[s0:s1], xcols, planes2d, out[s0:s1])
                continue
            nsh, f = s1 - s0, self.info
            self.backend.shells_x(cube, kind, kpow, lo[s0:s1], hi[s0:s1], xcols)
            if self.world > 1:
                # y-slabs -> x-slabs: x chunk r of the transformed columns goes to rank r (staged in the
                # planes2d scratch, which the (y,z) passes only write later), then the ky blocks of
                # the senders are interleaved back into the full ky axis, in place of xcols
                nx = nsh * f.xcols_complex_per_shell
                send, stage = xcols[:nx], planes2d[:nx]
                all_to_all_blocks(torch.view_as_real(stage).view(self.world, -1),
                                  torch.view_as_real(send).view(self.world, -1), self.group)
                if f.pruned:      # inner layout [kz][ky]
                    src = stage.view(self.world, f.mxl, nsh, f.kz, f.kyl).permute(1, 2, 3, 0, 4)
                    send.view(f.mxl, nsh, f.kz, self.world, f.kyl).copy_(src)
                else:             # inner layout [ky][kz]
                    src = stage.view(self.world, f.mxl, nsh, f.kyl, f.kz).permute(1, 2, 0, 3, 4)
                    send.view(f.mxl, nsh, self.world, f.kyl, f.kz).copy_(src)
            self.backend.shells_yz(nsh, xcols, planes2d, out[s0:s1])

    # -- contraction ---------------------------------------------------------- #
    def contract(self, fields, rows, job_off=((0, 0, 0),), marks=None, on_device=False):
        """Triangle sums over this rank's cells, all-reduced over ranks.

        fields: [nrows][ncells] tensor or a list of 1-D field tensors (nrows % 4 == 0; list
        entries may alias, which is how padding rows cost no memory);
        rows: (T,3) row triples into `fields`.  Returns float64 numpy [njobs][T], or with
        `on_device` the CUDA tensor (no host synchronisation: the caller fetches it later).
        """
        if torch.is_tensor(fields):
            fields = [fields[r] for r in range(fields.shape[0])]
        rows = np.asarray(rows, dtype=np.int64).reshape(-1, 3)
        job_off = np.asarray(job_off, dtype=np.int64).reshape(-1, 3)
        ncells = int(fields[0].numel())
        # Sparse lists (each field feeds few triangles) stream straight from HBM, one pass per
        # triangle; dense lists go through the shared-memory tile kernel.  Cost model: bytes.
        self.last_schedule = "tile"
        ntot = len(rows) * len(job_off)
        if hasattr(self.backend, "reduce_list") and ntot <= min(65535, 2 * len(fields)):
            allrows = (rows[None, :, :] + job_off[:, None, :]).reshape(-1, 3)
            srt = np.sort(allrows, axis=1)
            per_tri = 1 + (srt[:, 1] != srt[:, 0]) + (srt[:, 2] != srt[:, 1])
            if per_tri.sum() <= 2 * len(np.unique(allrows)):
                self.last_schedule = "stream"
        if self.last_schedule == "stream":
            sums = self.backend.reduce_list(fields, allrows, ncells).reshape(len(job_off), len(rows))
        else:
            sums = self.backend.contract(fields, rows, ncells, job_off)
            if getattr(self.backend, "last_path", 0) == 1:
                self.last_schedule = "tensor"
        _mark(marks, "contract_done", self)
        if self.world > 1:
            dist.all_reduce(sums, op=dist.ReduceOp.SUM, group=self.group)
        return sums if on_device else sums.cpu().numpy()

    def close(self):
        self._scratch = None
        self.backend.close()
        if self.device.type == "cuda":
            torch.cuda.empty_cache()      # hand the (often multi-GB) scratch back to the driver


def all_to_all_blocks(out, inp, group=None):
    """out[r] = block `rank` of rank r's inp (dim 0 split in `world` equal blocks): NCCL all-to-all;
    falls back to a list all-to-all where the backend has no single-tensor form."""
    try:
        dist.all_to_all_single(out, inp.contiguous(), group=group)
    except (RuntimeError, NotImplementedError):
        world = dist.get_world_size(group)
        outs = [t.contiguous() for t in out.chunk(world, dim=0)]
        dist.all_to_all(outs, [t.contiguous() for t in inp.chunk(world, dim=0)], group=group)
        for dst, t in zip(out.chunk(world, dim=0), outs):
            dst.copy_(t)


def all_gather_concat(out, inp, group=None):
    """out = concat over ranks of inp along dim 0 (NCCL all-gather; gloo-safe)."""
    try:
        dist.all_gather_into_tensor(out, inp.contiguous(), group=group)
    except (RuntimeError, NotImplementedError):
        world = dist.get_world_size(group)
        parts = list(out.chunk(world, dim=0))
        tmp = [torch.empty_like(inp) for _ in range(world)]
        dist.all_gather(tmp, inp.contiguous(), group=group)
        for p, t in zip(parts, tmp):
            p.copy_(t)


# --------------------------------------------------------------------------- #
# measurements on top of an Engine
# --------------------------------------------------------------------------- #
def _pad4(n):
    return (int(n) + 3) // 4 * 4


FIELD_ROUTE = {1: (0, 0, 0), 2: (0, 0, 1), 3: (0, 1, 2)}  # which mesh feeds k1,k2,k3 (main.py:627-640)


_UNIQ_CACHE = {}


def _unique_triples(triples):
    """Distinct index triples and the inverse map (cached: the same list is measured again for
    every mesh / every step, and np.unique over rows costs milliseconds of idle GPU)."""
    triples = np.ascontiguousarray(np.asarray(triples, dtype=np.int64).reshape(-1, 3))
    key = (triples.shape[0], hash(triples.tobytes()))
    hit = _UNIQ_CACHE.get(key)
    if hit is not None and np.array_equal(hit[0], triples):
        return hit[1], hit[2]
    uniq, inverse = np.unique(triples, axis=0, return_inverse=True)
    inverse = np.asarray(inverse).reshape(-1)
    if len(_UNIQ_CACHE) > 16:
        _UNIQ_CACHE.clear()
    _UNIQ_CACHE[key] = (triples.copy(), uniq, inverse)
    return uniq, inverse


def _mark(marks, name, engine):
    """Optional stage timing hook: record a CUDA event on the launching stream."""
    if marks is not None and engine.device.type == "cuda":
        ev = torch.cuda.Event(enable_timing=True)
        ev.record(torch.cuda.current_stream(engine.device))
        marks.append((name, ev))


def _alloc_table(shape, engine):
    """Field table allocation; under memory pressure cached blocks of other sizes are released
    first (the caching allocator cannot split a cached 96 GiB block into a 100 GiB request)."""
    import os
    if os.environ.get("BSK_DEBUG"):
        free, tot = torch.cuda.mem_get_info(engine.device)
        print(f"[bsk] table {shape} {engine.rdtype} = {shape[0]*shape[1]*engine.itemsize/2**30:.1f} GiB; free {free/2**30:.1f} "
              f"reserved {torch.cuda.memory_reserved(engine.device)/2**30:.1f} allocated "
              f"{torch.cuda.memory_allocated(engine.device)/2**30:.1f} rowcap {engine.row_capacity()}", flush=True)
    table, oom = None, 0
    try:
        table = torch.empty(shape, dtype=engine.rdtype, device=engine.device)
    except torch.OutOfMemoryError:
        oom = 1
    if engine.world > 1:
        # every rank must take the same path (the retry invalidates the cached row capacity, whose
        # re-measurement is collective): agree on "somebody ran out of memory"
        flag = torch.tensor([oom], dtype=torch.int32, device=engine.device)
        dist.all_reduce(flag, op=dist.ReduceOp.MAX, group=engine.group)
        any_oom = int(flag.item())
    else:
        any_oom = oom
    if not any_oom:
        return table
    del table
    engine._row_cap = None
    torch.cuda.empty_cache()
    return torch.empty(shape, dtype=engine.rdtype, device=engine.device)


def _batched_contract(engine: Engine, nseg, synth, job_seg_off, uniq, marks=None, symmetric=False, defer=False):
    """Evaluate every unique bin triple of `uniq` for each job.

    The field table has `nseg` segments (one per source: mesh A/B/C, or unit / |k| shells) with
    the same bin -> row layout; job j reads slot s from segment job_seg_off[j][s].  When all
    needed shells fit in device memory this is one synthesis + one contraction pass; otherwise
    the triangle list is cut (in order) into batches whose distinct bins fit, and the shells of
    each batch are synthesised afresh — the memory-for-recompute trade of the reference's
    "slow" path (bskit/main.py:1656-1661), but per batch instead of per triangle.
    """
    njobs = len(job_seg_off)
    out = np.empty((njobs, len(uniq)))
    pending = []                 # (batch, device tensor): fetched at the end, or by the caller when deferred
    nbins_all = len(np.unique(uniq))
    if engine.row_capacity() // nseg < 1.3 * nbins_all and engine.chunk > 1 and engine.max_rows is None:
        engine.set_chunk(1)      # trade synthesis batching for field memory before cutting the list
    cap = engine.row_capacity()
    if symmetric:        # the folded copy of the table lives next to it: 1/8 of it on one GPU, 1/4 on a slab
        cap = int(cap / (1.16 if engine.world == 1 else 1.30))
    seg_cap = max(1, cap // nseg)                          # distinct bins resident per segment
    if int(uniq.max()) + 1 <= seg_cap or len(np.unique(uniq)) <= seg_cap:
        batches = [np.arange(len(uniq))]              # the common case: everything is resident
    else:
        batches, cur, cur_bins = [], [], set()
        for t, tri in enumerate(uniq.tolist()):
            new_bins = cur_bins.union(tri)
            if cur and len(new_bins) > seg_cap:
                batches.append(np.asarray(cur))
                cur, new_bins = [], set(tri)
            cur.append(t)
            cur_bins = new_bins
        if cur:
            batches.append(np.asarray(cur))
    # one field table for all batches, sized for the largest: allocating (and releasing) tens of GiB per
    # batch costs more than the synthesis of a shell, and a cached block of another size cannot be reused
    max_nb = max(len(np.unique(uniq[b])) for b in batches)
    big_table = _alloc_table((nseg * max_nb, engine.ncells), engine) if max_nb <= seg_cap else None
    for batch in batches:
        tri = uniq[batch]
        bins = np.unique(tri)
        pos = np.full(int(bins.max()) + 1, -1, dtype=np.int64)
        pos[bins] = np.arange(len(bins))
        nb = len(bins)
        seg = _pad4(nb)
        if nb > seg_cap:
            raise MemoryError(f"one triangle needs {nseg * nb} resident shell fields of "
                              f"{engine.ncells * engine.itemsize / 2**30:.1f} GiB each; only "
                              f"{engine.row_capacity()} fit on this device (use grid='auto' or more GPUs)")
        table = big_table[:nseg * nb]
        fields = []
        for sidx in range(nseg):
            for run in _runs(bins.tolist()):
                r0 = sidx * nb + int(pos[run[0]])
                synth(sidx, run, table[r0: r0 + len(run)])
            # the 4x4x4 blocks read rows in aligned groups of four: pad each segment's row list by
            # repeating its first field (the products of padding rows are discarded)
            fields += [table[sidx * nb + r] for r in range(nb)] + [table[sidx * nb]] * (seg - nb)
        _mark(marks, "shells_done", engine)
        rows = pos[tri]
        job_off = np.asarray(job_seg_off, dtype=np.int64) * seg
        m = engine.grid.neval
        mxl = table.shape[1] // (m * m)
        if (symmetric and m % 2 == 0 and table.shape[1] == mxl * m * m and (engine.world > 1 or mxl == m)
                and hasattr(engine.backend, "fold_even")):
            # Unit-amplitude and |k|-weighted shells depend on |k| only, so they are even in every
            # axis: f(x,y,z) = f(-x,y,z) = ...  The sum over the grid is the sum over [0, M/2] per
            # mirrored axis with multiplicity w = prod w_axis (1 on the planes 0 and M/2, else 2).
            # Scaling every field by w^(1/3) puts the weight into the triple product, so one
            # contraction over the reduced cells replaces the full-grid one: all three axes on one
            # GPU (8x fewer cells), y and z on this rank's x-slab otherwise (4x fewer).
            octant = engine.backend.fold_even(table, m, mxl, fold_x=engine.world == 1)
            ofields = []
            for sidx in range(nseg):
                ofields += [octant[sidx * nb + r] for r in range(nb)] + [octant[sidx * nb]] * (seg - nb)
            pending.append((batch, engine.contract(ofields, rows, job_off, on_device=True)))
            _mark(marks, "contract_done", engine)
            del octant, ofields
        else:
            pending.append((batch, engine.contract(fields, rows, job_off, marks=marks, on_device=True)))
        del table, fields
    del big_table
    engine.last_batches = len(batches)

    def fetch():
        for batch, sums in pending:
            out[:, batch] = sums.cpu().numpy() if torch.is_tensor(sums) else np.asarray(sums)
        return out

    return fetch if defer else fetch()


def measure_triangle_sums(engine: Engine, cubes, edges, triples, marks=None, defer=False):
    """sum_x I_a I_b I_c / M^3 for every (a,b,c) in `triples` (indices into `edges`).

    cubes: 1-3 spectrum cubes (auto, <AAB>, <ABC> routing as the reference's slow
    path, main.py:627-640).  Multiply by V^2 for the unnormalised bispectrum
    (main.py:1875-1877: sum * V^2 / N^3).
    """
    edges = np.asarray(edges, dtype=np.float64).reshape(-1, 2)
    uniq, inverse = _unique_triples(triples)
    route = FIELD_ROUTE[len(cubes)]

    def synth(sidx, run, out):
        engine.synthesize(cubes[sidx], nat.KIND_DATA, 0.0, edges[run, 0], edges[run, 1], out)

    if defer:      # everything is enqueued; the returned callable synchronises and fetches the result
        fetch = _batched_contract(engine, len(cubes), synth, [route], uniq, marks, defer=True)
        return lambda: fetch()[0][inverse] / float(engine.grid.neval) ** 3
    sums = _batched_contract(engine, len(cubes), synth, [route], uniq, marks)[0]
    return sums[inverse] / float(engine.grid.neval) ** 3


def measure_grid_sums(engine: Engine, edges, triples, marks=None, defer=False):
    """(N_tri, k_mean[T,3]) from unit-amplitude and |k|-weighted shells
    (main.py:2006-2061): N_tri = sum n_a n_b n_c / M^3, k_1 = sum kappa_a n_b n_c / M^3 / N_tri ..."""
    edges = np.asarray(edges, dtype=np.float64).reshape(-1, 2)
    uniq, inverse = _unique_triples(triples)

    def synth(sidx, run, out):
        if sidx == 0:
            engine.synthesize(None, nat.KIND_UNIT, 0.0, edges[run, 0], edges[run, 1], out)
        else:
            engine.synthesize(None, nat.KIND_KPOW, 1.0, edges[run, 0], edges[run, 1], out)

    jobs = [(0, 0, 0), (1, 0, 0), (0, 1, 0), (0, 0, 1)]
    if defer:
        fetch = _batched_contract(engine, 2, synth, jobs, uniq, marks, symmetric=True, defer=True)
        return lambda: _finish_grid_sums(engine, fetch() / float(engine.grid.neval) ** 3, inverse)
    sums = _batched_contract(engine, 2, synth, jobs, uniq, marks, symmetric=True) / float(engine.grid.neval) ** 3
    return _finish_grid_sums(engine, sums, inverse)


def _finish_grid_sums(engine, sums, inverse):
    ntri = np.rint(sums[0])                      # an exact triangle count (integer valued)
    resid = float(np.max(np.abs(sums[0] - ntri))) if len(ntri) else 0.0
    engine.last_ntri_residual = resid
    if not resid < 0.05:
        raise nat.NativeError(f"N_tri is not integer valued (max |N_tri - round| = {resid:.3g}): the "
                              "float64 normalisation contraction has lost precision")
    with np.errstate(divide="ignore", invalid="ignore"):
        kmean = np.where(ntri[None, :] > 0, sums[1:4] / ntri[None, :], np.nan).T
    return ntri[inverse], np.ascontiguousarray(kmean[inverse])


def _runs(sorted_ids):
    """Split a sorted list of bin ids into maximal runs of consecutive ids."""
    runs, cur = [], []
    for b in sorted_ids:
        if cur and b != cur[-1] + 1:
            runs.append(cur)
            cur = []
        cur.append(b)
    if cur:
        runs.append(cur)
    return runs
