/* bskit_b200 C ABI — the drop-in boundary for the FFT-bispectrum hot path.
 *
 * The reference (sjforeman/bskit) has no FFI layer of its own: its hot path is
 * Python (bskit/main.py) calling nbodykit / pmesh / pfft / mpi4py.  Each entry
 * point below names the reference lines whose arithmetic it replaces, so a
 * maintainer can bind it from bskit/main.py with ctypes.
 *
 * Conventions
 *  - plain pointers and sizes only; no torch / numpy types cross this boundary;
 *  - every function returns 0 (BSK_OK) or a negative code; bsk_last_error()
 *    returns a thread-local message for the last failure;
 *  - "device" pointers are CUDA device memory owned by the CALLER (the Python
 *    host allocates them as torch tensors); the library only owns cuFFT plans,
 *    small lookup tables and the contraction schedule;
 *  - all work is enqueued on the cudaStream_t given at plan creation and is
 *    asynchronous with respect to the host unless stated otherwise;
 *  - a plan is not thread-safe; in a multi-GPU run every rank (one process per
 *    GPU) makes the same calls in the same order and the host performs the
 *    collectives (all-gather of cropped planes, all-reduce of triangle sums)
 *    between them with torch.distributed/NCCL.
 *
 * Grid conventions (pmesh): cubic mesh N^3, x slowest, z contiguous; half
 * spectrum on z; forward transform divided by N^3, inverse un-normalised;
 * k_axis = 2*pi*fftfreq(N,1/N)/L.
 */
#ifndef BSKIT_B200_H
#define BSKIT_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BSK_OK 0
#define BSK_ERR_ARG (-1)
#define BSK_ERR_CUDA (-2)
#define BSK_ERR_CUFFT (-3)
#define BSK_ERR_STATE (-4)

#define BSK_F32 0
#define BSK_F64 1

/* shell kinds, bsk_shells() */
#define BSK_KIND_DATA 0 /* delta_k * mask      main.py:1846-1861, 614-660 */
#define BSK_KIND_UNIT 1 /* mask                number_field, main.py:280-329 */
#define BSK_KIND_KPOW 2 /* |k|^p * mask        k_field,      main.py:227-277 */

typedef struct bsk_plan bsk_plan;   /* grid geometry + cuFFT plans + k tables */
typedef struct bsk_cplan bsk_cplan; /* triangle-list contraction schedule */

typedef struct bsk_geometry {
  int32_t nmesh;     /* N: input mesh is N^3 */
  int32_t neval;     /* M: grid the shell fields are synthesised on. M == N
                        reproduces the reference's transforms exactly; M < N is
                        the band-limited evaluation (exact when M > 3*ncrop) */
  int32_t ncrop;     /* keep modes with |n_axis| <= ncrop; >= N/2 keeps all */
  int32_t precision; /* BSK_F32 or BSK_F64: storage dtype of the shell fields */
  int32_t world;     /* number of x-slab shards (ranks) */
  int32_t rank;      /* this shard */
  int32_t max_shells;/* largest nsh a bsk_shells() call will pass */
  int32_t fft_precision; /* precision of the inverse transforms (>= precision).  F64
                        with F32 storage keeps the only float32 rounding per-cell and
                        uncorrelated; a float32 FFT's correlated error does not average
                        out of the heavily cancelling triangle sums */
  int32_t no_prune;  /* 1 = always use the generic cuFFT 2-D c2r path (debug / A-B timing);
                        0 = use the pruned y pass + fused z-pass kernel when neval is a
                        power of two in [64, 2048] and fft_precision is F64 */
  int32_t transposed;/* 1 = the spectrum cube is distributed in ky blocks (y-slabs) and the exchanges
                        around the x transforms are all-to-all transposes, the pfft scheme the
                        reference relies on (main.py:1612, every mesh.paint / c2r); for spectra
                        that cannot be cropped (2*ncrop+1 >= nmesh).  0 = every rank holds the whole
                        cropped cube (all-gather of the cropped planes) */
} bsk_geometry;

/* Derived sizes, bsk_plan_info(): counts in ELEMENTS (complex counts are numbers of
 * complex values).  The forward transform and the spectrum cube are ALWAYS float64 /
 * complex128 (a float32 FFT's rounding error scales with the rms of the whole spectrum
 * and would swamp the weak high-k modes); xcols / planes2d are in fft_precision, fields
 * in precision. */
typedef struct bsk_info {
  int64_t kx, ky, kz;          /* cropped spectrum cube dims */
  int64_t nx0, nxl;            /* local x-planes of the N grid  [nx0, nx0+nxl) */
  int64_t mx0, mxl;            /* local x-planes of the M grid */
  int64_t fwd_batch;           /* planes per forward 2-D transform chunk (divides nxl) */
  int64_t fwd_work_complex;    /* fwd_batch * N * (N/2+1)  (complex128) */
  int64_t planes_local_complex;/* nxl * ky * kz */
  int64_t planes_all_complex;  /* N * ky * kz */
  int64_t cube_complex;        /* kx * ky * kz */
  int64_t xcols_complex_per_shell;  /* M * ky * kz       (W buffer) */
  int64_t planes2d_complex_per_shell;/* mxl * M * (M/2+1), or mxl * M * kz on the pruned path */
  int64_t field_real_per_shell;     /* mxl * M * M = local cells */
  int64_t fft_work_bytes;      /* cuFFT work areas owned by the plan */
  int64_t kyl, ky0;            /* ky block of this rank's part of the cube: [ky0, ky0+kyl); kyl == ky
                                  unless the plan is transposed.  Then planes_all = N * kyl * kz,
                                  cube = kx * kyl * kz and xcols per shell = M * kyl * kz */
  int64_t xplanes_complex_per_shell; /* transposed plans: mxl * ky * kz, the buffer bsk_shells_yz reads */
  int64_t pruned;              /* 1: pruned y pass + fused z pass (inner layout of xcols / xplanes is [kz][ky]) */
} bsk_info;

int bsk_version(void);
const char* bsk_last_error(void);

/* Plan.  kx_tab/ky_tab/kz_tab: HOST float64 wavenumber per cropped-cube index
 * along each axis, computed by the host exactly as the reference's grid does
 * (2*pi*fftfreq(N,1/N)/L, negative frequencies in the upper half), so that the
 * in-kernel |k| = sqrt(kx^2+ky^2+kz^2) is bit-identical to numpy's
 * (main.py:1850).  comp_x/y/z: optional HOST float64 multiplicative
 * compensation per cropped-cube index (CIC window, measure_bs_fast.py:45-57);
 * NULL = none. */
int bsk_plan_create(bsk_plan** out, const bsk_geometry* geom, const double* kx_tab,
                    const double* ky_tab, const double* kz_tab, void* cuda_stream);
int bsk_plan_destroy(bsk_plan* plan);
int bsk_plan_info(const bsk_plan* plan, bsk_info* out);
int bsk_set_compensation(bsk_plan* plan, const double* comp_x, const double* comp_y,
                         const double* comp_z);
/* Enqueue every later call of this plan (kernels and cuFFT plans) on `cuda_stream`.  The host calls
 * it with its current stream before each stage, so a plan built under one stream can be used under
 * another (the reference has no streams: every pmesh call is synchronous). */
int bsk_plan_set_stream(bsk_plan* plan, void* cuda_stream);

/* Forward transform of this rank's x-slab (replaces mesh.paint(mode='complex'),
 * main.py:1608-1621, plus the queued compensation action).  Runs in float64, in chunks
 * of fwd_batch planes.
 *   mesh_slab    device, [nxl][N][N] real, dtype mesh_dtype (BSK_F32/BSK_F64)
 *   work         device scratch, fwd_work_complex complex128 values
 *   convert_work device scratch, fwd_batch*N*N float64 (float32 meshes only, else NULL)
 *   planes_local device out, [nxl][ky][kz] complex128: 2-D r2c over (y,z), scaled by
 *                1/N^3, y/z compensation applied, cropped to the kept modes */
int bsk_forward_local(bsk_plan* plan, const void* mesh_slab, int mesh_dtype, void* work,
                      void* convert_work, void* planes_local);
/* After the host all-gathered planes_local into planes_all [N][ky][kz] (complex128):
 * in-place x transform and crop -> cube [kx][ky][kz] complex128 (x compensation applied). */
int bsk_forward_finish(bsk_plan* plan, void* planes_all, void* cube);

/* Exact integer number of full-cube modes in each k-bin (inclusive both ends,
 * main.py:1852), Hermitian multiplicity included.  lo/hi: HOST float64[nbins];
 * counts: HOST out int64[nbins].  Synchronous. */
int bsk_modes_per_bin(bsk_plan* plan, int nbins, const double* lo, const double* hi,
                      int64_t* counts);

/* Shell synthesis for nsh k-bins (replaces the per-bin mask + c2r of
 * main.py:1846-1861 / number_field / k_field): k-shell filter of the cube,
 * inverse x transform, scatter into zero-padded (y,z) half-spectra of the local
 * planes, batched 2-D c2r.
 *   cube     device [kx][ky][kz] complex128 (ignored for UNIT / KPOW kinds)
 *   lo, hi   HOST float64[nsh] bin edges, both inclusive
 *   xcols    device scratch, nsh * xcols_complex_per_shell complex
 *   planes2d device scratch, nsh * planes2d_complex_per_shell complex
 *   fields   device out, [nsh][mxl*M*M] real (contiguous) */
/* Create the cuFFT plans (and their work areas) a bsk_shells() call with nsh bins will use,
 * so that the caller can size its field table from the memory that is really left. */
int bsk_shells_prepare(bsk_plan* plan, int nsh);
int bsk_shells(bsk_plan* plan, const void* cube, int kind, double kpow, int nsh,
               const double* lo, const double* hi, void* xcols, void* planes2d, void* fields);
/* The two halves of bsk_shells() for transposed plans (the distributed c2r of main.py:1859-1861 as
 * pfft does it): bsk_shells_x filters this rank's ky block of the cube and runs the inverse x
 * transform -> xcols [M][nsh][kyl][kz] ([M][nsh][kz][kyl] on the pruned path, fft_precision complex);
 * the host then transposes with one all-to-all (chunk r of the x axis goes to rank r, the ky blocks
 * are concatenated in rank order) into xplanes [mxl][nsh][ky][kz] ([mxl][nsh][kz][ky] pruned), and
 * bsk_shells_yz finishes the (y,z) transforms of this rank's planes into fields [nsh][mxl*M*M].
 * Likewise bsk_forward_finish() of a transposed plan takes planes_all = [N][kyl][kz]: the planes of
 * bsk_forward_local() after the all-to-all that sends ky block r to rank r. */
int bsk_shells_x(bsk_plan* plan, const void* cube, int kind, double kpow, int nsh, const double* lo,
                 const double* hi, void* xcols);
int bsk_shells_yz(bsk_plan* plan, int nsh, const void* xplanes, void* planes2d, void* fields);

/* Triangle contraction schedule for a list of triangles given as TILE-ROW
 * triples: rows[t] = (r1, r2, r3) indexes the array of field pointers handed to
 * bsk_contract().  Triangles are grouped into 4x4x4 register blocks. */
int bsk_cplan_create(bsk_cplan** out, int ntri, const int32_t* rows, int nrows, int max_jobs);
int bsk_cplan_destroy(bsk_cplan* cp);
int bsk_cplan_info(const bsk_cplan* cp, int64_t out[4]); /* nblocks, split, rounds, threads */

/* Contraction path of a schedule.  path 0 (default): FP32-pipe tile kernel (packed FFMA2, exact
 * round-to-nearest products).  path 1: tcgen05 tensor cores, 3xTF32 operands with the pair
 * products written to TMEM, accumulators drained every 128 cells (relative error ~8e-7 of the
 * largest sums instead of ~5e-8; the Python host selects it by default).  Path 1 is used only when
 * the call is eligible: float32 fields and products, one job with zero offsets, at least 256
 * triangles, ncells a multiple of 128; lists that do not fit one launch (more than 40 column rows,
 * 96 accumulator columns per team or 120 raw rows: S = 80 bins, two- and three-field lists) run as
 * several passes.  Otherwise bsk_contract runs path 0.  bsk_cplan_path: out = {tensor-core units
 * of the schedule (0: not eligible), requested path, path the last bsk_contract call ran}. */
int bsk_cplan_set_path(bsk_cplan* cp, int path);
/* Host only (no device needed): the tensor-core schedule bsk_cplan_create builds for a list.
 * out = {128-row units over all passes (0: list not eligible), distinct accumulator slots read by
 * the triangles (== number of distinct sorted triangles; negative if two different triangles
 * share a slot), sum of the units' accumulator columns, passes (kernel launches), all slots in
 * range, sum over units of max(11, columns/2) (MMA cycles per 8 cells and operand term)}. */
int bsk_tc_schedule_info(int ntri, const int32_t* rows, int nrows, int64_t out[6]);
/* Host only: evaluate the tensor-core schedule of a list on HOST float64 fields [nrows][ncells] by following
 * its tables exactly as tc_contract_kernel does (lane -> pair rows, unit window column -> raw slot -> field
 * row, (team, accumulator column, lane) -> partial slot, triangle -> slot), in plain float64 loops: sums[t] must
 * equal sum_x f[r1] f[r2] f[r3] of triangle t.  Pins the schedule builders for any list shape without a GPU
 * (tests/test_cabi.py); O(units * 128 * columns * ncells), meant for a few hundred cells. */
int bsk_tc_schedule_eval(int ntri, const int32_t* rows, int nrows, int64_t ncells, const double* fields,
                         double* sums);
int bsk_cplan_path(const bsk_cplan* cp, int64_t out[3]);

/* sums[j][t] = sum over local cells x of
 *     F[r1 + off[j][0]](x) * F[r2 + off[j][1]](x) * F[r3 + off[j][2]](x)
 * (replaces the per-triangle np.sum of main.py:1875, 2027-2055; the caller
 * applies V^2/N^3 etc. and all-reduces across ranks).
 *   row_ptrs  HOST array of nrows DEVICE pointers, one real field of ncells each
 *   job_off   HOST int32[njobs][3] row offsets per job (0,0,0 for plain B)
 *   precision        storage dtype of the fields (BSK_F32 / BSK_F64)
 *   accum_precision  dtype of the products and of the per-tile partial sums: equal to
 *                    precision, or BSK_F64 for float32 fields (every product and add in
 *                    float64; ~2x the time).  Tile partials are always folded in float64.
 *   sums      device out float64 [njobs][ntri] */
int bsk_contract(bsk_cplan* cp, const void* const* row_ptrs, int precision, int accum_precision,
                 int64_t ncells, int njobs, const int32_t* job_off, double* sums,
                 void* cuda_stream);

/* Fold of even fields for the normalisation (main.py:2006-2061).  n_i and kappa_i depend on |k|
 * only, so they are even in every axis and sum_x over the grid equals the sum over [0, M/2] per
 * mirrored axis with weight prod_axis (1 on the planes 0 and M/2, else 2).  Writes, for each of
 * the `nrows` contiguous fields [mxl][M][M], the cells [0,nx) x [0,M/2] x [0,M/2] scaled by the
 * cube root of that weight (so the weight lands in the triple product), rows zero-padded to
 * ncell_out.  fold_x = 1 (whole grid local, nx = M/2+1) folds all three axes, fold_x = 0
 * (x-slab of a multi-GPU run, nx = mxl) folds y and z only. */
int bsk_fold_even(const void* fields, int precision, int nrows, int neval, int mxl, int fold_x, void* out,
                  int64_t ncell_out, void* cuda_stream);

/* Sparse triangle lists (equilateral / squeezed / isosceles, T ~ S: every triangle touches
 * <= 3 fields): one coalesced streaming pass per triangle straight from HBM,
 *     sums[t] = sum_x F[rows[t][0]](x) * F[rows[t][1]](x) * F[rows[t][2]](x)
 * (the same np.sum of main.py:1875 / 2027-2055, for lists where the 4x4x4-block schedule of
 * bsk_contract would waste work).  row_ptrs, rows: HOST arrays; sums: device float64[ntri].
 * Returns after the work is enqueued and the host tables have been consumed. */
int bsk_reduce_list(const void* const* row_ptrs, int nrows, int precision, int64_t ncells, int ntri,
                    const int32_t* rows, double* sums, void* cuda_stream);

/* Cloud-in-cell mass assignment (the step before the path for particle inputs; replaces
 * catalog.to_mesh(window='cic') + paint of scripts/measure/measure_bs_fast.py:209-217).
 *   pos     device [npart][3] positions (float32 or float64 per `precision`), any real values:
 *           they are wrapped periodically into the box
 *   mesh    device out float32 [nmesh][nmesh][nmesh], zeroed by the call; holds the weight sums
 *           (divide by npart / nmesh^3 for 1 + delta) */
int bsk_paint_cic(const void* pos, int precision, int64_t npart, int nmesh, const double boxsize[3],
                  float* mesh, void* cuda_stream);

/* number of kernels this library has launched since load (bench bookkeeping) */
int64_t bsk_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* BSKIT_B200_H */
