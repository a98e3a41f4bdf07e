// Forward transform, k-shell filter and shell synthesis around cuFFT (sm_100a).
//
// Replaces, on the GPU, the reference's per-bin "copy delta_k, mask by |k|, c2r"
// loop (bskit/main.py:1846-1861), number_field / k_field (main.py:227-329) and
// the one-off forward paint (main.py:1608-1621) with its CIC compensation action
// (scripts/measure/measure_bs_fast.py:45-57).
//
// Data flow per rank (x-slab decomposition, N = mesh, M = evaluation grid):
//   mesh slab [nxl][N][N] --2-D R2C--> [nxl][N][N/2+1] --crop_yz--> [nxl][Ky][Kz]
//   (host all-gather over ranks) [N][Ky][Kz] --1-D C2C along x--> --crop_x--> cube
//   cube [Kx][Ky][Kz] --shell_filter--> xcols [M][nsh][Ky][Kz] --1-D inverse x-->
//   --scatter_planes--> [nsh][mxl][M][M/2+1] --2-D C2R--> fields [nsh][mxl][M][M]
#include "common.cuh"

#include <cmath>

namespace bsk {

thread_local std::string g_err;
std::atomic<int64_t> g_launches{0};

void set_error(const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  g_err = buf;
}

template <typename T> struct Cx;
template <> struct Cx<float> { using type = float2; };
template <> struct Cx<double> { using type = double2; };

// ---------------------------------------------------------------------------
// dtype conversion of the input slab (f64 mesh on an f32 plan and vice versa)
// ---------------------------------------------------------------------------
template <typename Tin, typename Tout>
__global__ void convert_kernel(const Tin* __restrict__ in, Tout* __restrict__ out, int64_t n) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) out[i] = (Tout)in[i];
}

// ---------------------------------------------------------------------------
// crop after the (y,z) transform: scale by 1/N^3, y/z compensation, keep modes
// ---------------------------------------------------------------------------
template <typename T>
__global__ void crop_yz_kernel(const typename Cx<T>::type* __restrict__ spec,  // [nxl][N][N/2+1]
                               typename Cx<T>::type* __restrict__ planes,      // [nxl][Ky][Kz]
                               int nxl, int N, int Ky, int Kz, double scale,
                               const double* __restrict__ cy, const double* __restrict__ cz) {
  const int64_t total = (int64_t)nxl * Ky * Kz;
  const int nzh = N / 2 + 1;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    int jz = (int)(i % Kz);
    int64_t r = i / Kz;
    int jy = (int)(r % Ky);
    int xl = (int)(r / Ky);
    int ny = mode_of(jy, Ky, N);
    int iy = ny < 0 ? ny + N : ny;
    typename Cx<T>::type v = spec[((int64_t)xl * N + iy) * nzh + jz];
    T f = (T)(scale * cy[jy] * cz[jz]);
    v.x *= f;
    v.y *= f;
    planes[i] = v;
  }
}

template <typename T>
__global__ void crop_x_kernel(const typename Cx<T>::type* __restrict__ planes_all,  // [N][Ky][Kz]
                              typename Cx<T>::type* __restrict__ cube,              // [Kx][Ky][Kz]
                              int N, int Kx, int64_t plane, const double* __restrict__ cx) {
  const int64_t total = (int64_t)Kx * plane;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    int jx = (int)(i / plane);
    int64_t rem = i - (int64_t)jx * plane;
    int nx = mode_of(jx, Kx, N);
    int ix = nx < 0 ? nx + N : nx;
    typename Cx<T>::type v = planes_all[(int64_t)ix * plane + rem];
    T f = (T)cx[jx];
    v.x *= f;
    v.y *= f;
    cube[i] = v;
  }
}

// ---------------------------------------------------------------------------
// k-shell filter.  One thread per cube mode: |k| in float64 with the exact
// operation order of numpy's  sum(ki**2. for ki in k)**0.5  (main.py:1850) — no
// FMA contraction — then an inclusive test against each bin of the chunk
// (main.py:1852).  Output is laid out [x][shell][ky][kz] so that ONE strided
// cuFFT call does the inverse x transform of every shell of the chunk.
// ---------------------------------------------------------------------------
constexpr int kMaxChunk = 64;
struct BinEdges {
  double lo[kMaxChunk];
  double hi[kMaxChunk];
};

__device__ __forceinline__ double knorm_exact(double kx, double ky, double kz) {
  double s = __dadd_rn(__dadd_rn(__dmul_rn(kx, kx), __dmul_rn(ky, ky)), __dmul_rn(kz, kz));
  return __dsqrt_rn(s);
}

template <typename T>
__global__ void shell_filter_kernel(const double2* __restrict__ cube,  // [Kx][Ky][Kz] float64
                                    typename Cx<T>::type* __restrict__ xcols,  // [M][nsh][Ky][Kz]
                                    int Kx, int Ky, int Kz, int N, int M, int nsh, int kind,
                                    double kpow, int ky_inner, BinEdges bins,
                                    const double* __restrict__ kxt,
                                    const double* __restrict__ kyt,
                                    const double* __restrict__ kzt) {
  const int64_t plane = (int64_t)Ky * Kz;
  const int64_t total = (int64_t)Kx * plane;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    int jx = (int)(i / plane);
    int64_t rem = i - (int64_t)jx * plane;
    int jy = (int)(rem / Kz);
    int jz = (int)(rem - (int64_t)jy * Kz);
    double kk = knorm_exact(kxt[jx], kyt[jy], kzt[jz]);
    typename Cx<T>::type v;
    if (kind == BSK_KIND_DATA) {
      const double2 c = cube[i];
      v.x = (T)c.x;
      v.y = (T)c.y;
    } else if (kind == BSK_KIND_UNIT) {
      v.x = (T)1;
      v.y = (T)0;
    } else {
      double w = (kpow == 1.0) ? kk : (kpow == 0.5 ? sqrt(kk) : pow(kk, kpow));
      v.x = (T)w;
      v.y = (T)0;
    }
    int nx = mode_of(jx, Kx, N);
    int mx = nx < 0 ? nx + M : nx;
    typename Cx<T>::type zero;
    zero.x = (T)0;
    zero.y = (T)0;
    // inner layout [ky][kz] for the generic path, [kz][ky] (ky contiguous) for the pruned path,
    // whose y-scatter then reads and writes contiguous runs
    typename Cx<T>::type* dst =
        xcols + (int64_t)mx * nsh * plane + (ky_inner ? (int64_t)jz * Ky + jy : rem);
#pragma unroll 4
    for (int s = 0; s < nsh; ++s) {
      bool in = (kk <= bins.hi[s]) & (kk >= bins.lo[s]);
      dst[(int64_t)s * plane] = in ? v : zero;
    }
  }
}

// exact mode counts per bin (Hermitian multiplicity of the half spectrum)
__global__ void mode_count_kernel(int Kx, int Ky, int Kz, int N, int nbins,
                                  const double* __restrict__ lo, const double* __restrict__ hi,
                                  const double* __restrict__ kxt, const double* __restrict__ kyt,
                                  const double* __restrict__ kzt,
                                  unsigned long long* __restrict__ counts) {
  const int64_t plane = (int64_t)Ky * Kz;
  const int64_t total = (int64_t)Kx * plane;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    int jx = (int)(i / plane);
    int64_t rem = i - (int64_t)jx * plane;
    int jy = (int)(rem / Kz);
    int jz = (int)(rem - (int64_t)jy * Kz);
    double kk = knorm_exact(kxt[jx], kyt[jy], kzt[jz]);
    unsigned long long w = (jz == 0 || (N % 2 == 0 && jz == N / 2)) ? 1ull : 2ull;
    for (int b = 0; b < nbins; ++b)
      if ((kk <= hi[b]) & (kk >= lo[b])) atomicAdd(&counts[b], w);
  }
}

// ---------------------------------------------------------------------------
// scatter the x-transformed columns of the local planes into zero-padded
// (y,z) half-spectra: the pre-processing of the batched 2-D C2R.  Write-bound:
// every output element is written exactly once, coalesced along z.
// ---------------------------------------------------------------------------
template <typename T>
__global__ void scatter_planes_kernel(const typename Cx<T>::type* __restrict__ xcols,  // [M][nsh][Ky][Kz]
                                      typename Cx<T>::type* __restrict__ planes2d,  // [nsh][mxl][M][M/2+1]
                                      int M, int Ky, int Kz, int nsh, int mx0, int mxl) {
  const int mzh = M / 2 + 1;
  const int64_t rows = (int64_t)nsh * mxl * M;  // (s, xl, iy)
  const int nc = (Ky - 1) / 2;
  typename Cx<T>::type zero;
  zero.x = (T)0;
  zero.y = (T)0;
  for (int64_t row = blockIdx.x; row < rows; row += gridDim.x) {
    int iy = (int)(row % M);
    int64_t r2 = row / M;
    int xl = (int)(r2 % mxl);
    int s = (int)(r2 / mxl);
    int jy;
    if (Ky == M) jy = iy;
    else if (iy <= nc) jy = iy;
    else if (iy >= M - nc) jy = iy - M + Ky;
    else jy = -1;
    typename Cx<T>::type* dst = planes2d + row * mzh;
    if (jy < 0) {
      for (int iz = threadIdx.x; iz < mzh; iz += blockDim.x) dst[iz] = zero;
    } else {
      const typename Cx<T>::type* src =
          xcols + (((int64_t)(mx0 + xl) * nsh + s) * Ky + jy) * Kz;
      for (int iz = threadIdx.x; iz < mzh; iz += blockDim.x) dst[iz] = iz < Kz ? src[iz] : zero;
    }
  }
}

// ---------------------------------------------------------------------------
// Even-field fold.  Unit-amplitude and |k|-weighted shells depend on |k| only, so they are even in
// every axis: sum_x f g h over the grid = sum over [0, M/2] per mirrored axis of w f g h with
// w = prod_axis (1 on the planes 0 and M/2, 2 elsewhere).  Scaling each field by w^(1/3) puts the
// weight into the triple product, so the normalisation (main.py:2024-2061) contracts nx*h*h cells
// instead of mxl*M*M.  out rows are zero-padded to `ncell_out` (a multiple of 4).
// ---------------------------------------------------------------------------
template <typename T>
__global__ void fold_even_kernel(const T* __restrict__ fields, T* __restrict__ out, int nrows, int M, int mxl,
                                 int nx, int fold_x, int64_t ncell_out) {
  const int h = M / 2 + 1;
  const int64_t nred = (int64_t)nx * h * h;
  const double c1 = 1.0, c2 = 1.2599210498948731648;   // 2^(1/3)
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < (int64_t)nrows * ncell_out;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int r = (int)(i / ncell_out);
    const int64_t c = i - (int64_t)r * ncell_out;
    T v = (T)0;
    if (c < nred) {
      const int z = (int)(c % h);
      const int64_t t = c / h;
      const int y = (int)(t % h), x = (int)(t / h);
      double w = ((y == 0 || y == h - 1) ? c1 : c2) * ((z == 0 || z == h - 1) ? c1 : c2);
      if (fold_x) w *= (x == 0 || x == h - 1) ? c1 : c2;
      v = (T)((double)fields[((int64_t)r * mxl + x) * M * M + (int64_t)y * M + z] * w);
    }
    out[i] = v;
  }
}

static inline int grid_for(int64_t n, int block, int cap = 148 * 16) {
  int64_t g = (n + block - 1) / block;
  return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

// ---------------------------------------------------------------------------
// cuFFT plan helpers
// ---------------------------------------------------------------------------
static int make_plan_many(cufftHandle* h, int rank, long long* n, long long* inembed,
                          long long istride, long long idist, long long* onembed,
                          long long ostride, long long odist, cufftType type, long long batch,
                          cudaStream_t st, size_t* work) {
  BSK_FFT(cufftCreate(h));
  size_t ws = 0;
  BSK_FFT(cufftMakePlanMany64(*h, rank, n, inembed, istride, idist, onembed, ostride, odist, type,
                              batch, &ws));
  BSK_FFT(cufftSetStream(*h, st));
  *work += ws;
  return BSK_OK;
}

static int get_invx(bsk_plan* p, int nsh, cufftHandle* out) {
  auto it = p->invx.find(nsh);
  if (it != p->invx.end()) {
    *out = it->second;
    return BSK_OK;
  }
  int M = p->g.neval;
  long long cols = (long long)nsh * p->info.kyl * p->info.kz;
  long long n[1] = {M};
  long long emb[1] = {M};
  cufftHandle h;
  int rc = make_plan_many(&h, 1, n, emb, cols, 1, emb, cols, 1,
                          p->g.fft_precision == BSK_F32 ? CUFFT_C2C : CUFFT_Z2Z, cols, p->stream,
                          &p->fft_work_bytes);
  if (rc) return rc;
  p->invx[nsh] = h;
  *out = h;
  return BSK_OK;
}

static int get_inv2d(bsk_plan* p, int nsh, cufftHandle* out) {
  auto it = p->inv2d.find(nsh);
  if (it != p->inv2d.end()) {
    *out = it->second;
    return BSK_OK;
  }
  int M = p->g.neval;
  long long n[2] = {M, M};
  long long inembed[2] = {M, M / 2 + 1};
  // same precision for transform and storage: out of place into the compact field array;
  // float64 transform with float32 storage: in place (rows padded to M+2 reals), narrowed after
  const bool inplace = p->g.fft_precision != p->g.precision;
  long long onembed[2] = {M, inplace ? M + 2 : M};
  cufftHandle h;
  int rc = make_plan_many(&h, 2, n, inembed, 1, (long long)M * (M / 2 + 1), onembed, 1,
                          inplace ? (long long)M * (M + 2) : (long long)M * M,
                          p->g.fft_precision == BSK_F32 ? CUFFT_C2R : CUFFT_Z2D,
                          (long long)nsh * p->info.mxl, p->stream, &p->fft_work_bytes);
  if (rc) return rc;
  p->inv2d[nsh] = h;
  *out = h;
  return BSK_OK;
}

static int get_invy(bsk_plan* p, int nsh, cufftHandle* out) {
  auto it = p->invy.find(nsh);
  if (it != p->invy.end()) {
    *out = it->second;
    return BSK_OK;
  }
  long long M = p->g.neval;
  long long n[1] = {M};
  long long emb[1] = {M};
  cufftHandle h;
  int rc = make_plan_many(&h, 1, n, emb, 1, M, emb, 1, M, CUFFT_Z2Z,
                          (long long)nsh * p->info.mxl * p->info.kz, p->stream, &p->fft_work_bytes);
  if (rc) return rc;
  p->invy[nsh] = h;
  *out = h;
  return BSK_OK;
}

}  // namespace bsk

using namespace bsk;

// float64 rows padded to M+2 (in-place c2r output) -> compact float32 field rows
__global__ void narrow_rows_kernel(const double* __restrict__ src, float* __restrict__ dst,
                                   int64_t rows, int M) {
  const int64_t total = rows * (M / 2);
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / (M / 2);
    const int c = (int)(i - r * (M / 2));
    const double2 v = reinterpret_cast<const double2*>(src + r * (M + 2))[c];
    reinterpret_cast<float2*>(dst + r * M)[c] = make_float2((float)v.x, (float)v.y);
  }
}

// Shell synthesis, first half: k-shell filter of the (local ky block of the) cube and the inverse x
// transform of every kept column.  xcols: [M][nsh][kyl][kz], or [M][nsh][kz][kyl] on the pruned path.
template <typename TF>
static int shells_x_impl(bsk_plan* p, const void* cube, int kind, double kpow, int nsh,
                         const BinEdges& be, void* xcols) {
  using C = typename Cx<TF>::type;
  const bsk_info& f = p->info;
  const int N = p->g.nmesh, M = p->g.neval;
  const int64_t xc = (int64_t)nsh * f.xcols_complex_per_shell;
  if (f.kx != M)  // rows of the padded x axis that no kept mode maps to must be zero
    BSK_CUDA(cudaMemsetAsync(xcols, 0, sizeof(C) * (size_t)xc, p->stream));
  shell_filter_kernel<TF><<<grid_for(f.cube_complex, 256), 256, 0, p->stream>>>(
      (const double2*)cube, (C*)xcols, (int)f.kx, (int)f.kyl, (int)f.kz, N, M, nsh, kind, kpow,
      p->use_zpass ? 1 : 0, be, p->d_kx, p->d_ky + f.ky0, p->d_kz);
  count_launch();
  BSK_CUDA(cudaGetLastError());
  cufftHandle hx;
  int rc;
  if ((rc = get_invx(p, nsh, &hx))) return rc;
  if (sizeof(TF) == 4)
    BSK_FFT(cufftExecC2C(hx, (cufftComplex*)xcols, (cufftComplex*)xcols, CUFFT_INVERSE));
  else
    BSK_FFT(cufftExecZ2Z(hx, (cufftDoubleComplex*)xcols, (cufftDoubleComplex*)xcols, CUFFT_INVERSE));
  return BSK_OK;
}

// Second half: the (y,z) transforms of this rank's x-planes.  xplanes: the x-transformed columns of the
// planes [mx0, mx0+mxl) with the FULL ky range, [mxl][nsh][ky][kz] ([mxl][nsh][kz][ky] pruned).
template <typename TF, typename TS>  // transform precision, storage precision
static int shells_yz_impl(bsk_plan* p, int nsh, const void* xplanes, void* planes2d, void* fields) {
  using C = typename Cx<TF>::type;
  const bsk_info& f = p->info;
  const int M = p->g.neval;
  cufftHandle h2;
  int rc;
  if (p->use_zpass) {  // pruned y pass (cuFFT on the kept kz columns) + fused z pass
    cufftHandle hy;
    if ((rc = get_invy(p, nsh, &hy))) return rc;
    return zpass_run(M, sizeof(TS) == 4, xplanes, planes2d, fields, (int)f.ky, (int)f.kz, nsh,
                     0, (int)f.mxl, p->d_wtab, hy, p->stream);
  }
  if ((rc = get_inv2d(p, nsh, &h2))) return rc;
  const int64_t rows = (int64_t)nsh * f.mxl * M;
  int grid = (int)(rows < 148 * 32 ? rows : 148 * 32);
  scatter_planes_kernel<TF><<<grid, 128, 0, p->stream>>>((const C*)xplanes, (C*)planes2d, M, (int)f.ky,
                                                         (int)f.kz, nsh, 0, (int)f.mxl);
  count_launch();
  BSK_CUDA(cudaGetLastError());
  if (sizeof(TF) == sizeof(TS)) {
    if (sizeof(TF) == 4)
      BSK_FFT(cufftExecC2R(h2, (cufftComplex*)planes2d, (cufftReal*)fields));
    else
      BSK_FFT(cufftExecZ2D(h2, (cufftDoubleComplex*)planes2d, (cufftDoubleReal*)fields));
  } else {
    BSK_FFT(cufftExecZ2D(h2, (cufftDoubleComplex*)planes2d, (cufftDoubleReal*)planes2d));
    narrow_rows_kernel<<<grid_for(rows * (M / 2), 256), 256, 0, p->stream>>>(
        (const double*)planes2d, (float*)fields, rows, M);
    count_launch();
    BSK_CUDA(cudaGetLastError());
  }
  return BSK_OK;
}

static int shells_x_dispatch(bsk_plan* p, const void* cube, int kind, double kpow, int nsh, const double* lo,
                             const double* hi, void* xcols) {
  BinEdges be;
  for (int s = 0; s < nsh; ++s) {
    be.lo[s] = lo[s];
    be.hi[s] = hi[s];
  }
  for (int s = nsh; s < kMaxChunk; ++s) be.lo[s] = be.hi[s] = 0.0;
  return p->g.fft_precision == BSK_F64 ? shells_x_impl<double>(p, cube, kind, kpow, nsh, be, xcols)
                                       : shells_x_impl<float>(p, cube, kind, kpow, nsh, be, xcols);
}

static int shells_yz_dispatch(bsk_plan* p, int nsh, const void* xplanes, void* planes2d, void* fields) {
  if (p->g.precision == BSK_F64) return shells_yz_impl<double, double>(p, nsh, xplanes, planes2d, fields);
  return p->g.fft_precision == BSK_F64 ? shells_yz_impl<double, float>(p, nsh, xplanes, planes2d, fields)
                                       : shells_yz_impl<float, float>(p, nsh, xplanes, planes2d, fields);
}

// ---------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------
extern "C" {

int bsk_version(void) { return 100; }
const char* bsk_last_error(void) { return g_err.c_str(); }
int64_t bsk_launch_count(void) { return g_launches.load(); }

static int upload(double** dst, const double* src, int64_t n, cudaStream_t st) {
  BSK_CUDA(cudaMalloc((void**)dst, sizeof(double) * (size_t)n));
  BSK_CUDA(cudaMemcpyAsync(*dst, src, sizeof(double) * (size_t)n, cudaMemcpyHostToDevice, st));
  BSK_CUDA(cudaStreamSynchronize(st));
  return BSK_OK;
}

int bsk_plan_create(bsk_plan** out, const bsk_geometry* geom, const double* kx_tab,
                    const double* ky_tab, const double* kz_tab, void* cuda_stream) {
  BSK_REQUIRE(out && geom && kx_tab && ky_tab && kz_tab, "bsk_plan_create: null argument");
  const bsk_geometry& g = *geom;
  const int N = g.nmesh, M = g.neval;
  BSK_REQUIRE(N >= 4 && N % 2 == 0, "nmesh must be even and >= 4 (got %d)", N);
  BSK_REQUIRE(M >= 4 && M % 2 == 0 && M <= N, "neval must be even, >= 4 and <= nmesh (got %d)", M);
  BSK_REQUIRE(g.precision == BSK_F32 || g.precision == BSK_F64, "bad precision %d", g.precision);
  BSK_REQUIRE(g.fft_precision == BSK_F32 || g.fft_precision == BSK_F64, "bad fft_precision %d",
              g.fft_precision);
  BSK_REQUIRE(g.fft_precision >= g.precision, "fft_precision may not be lower than precision");
  BSK_REQUIRE(g.world >= 1 && g.rank >= 0 && g.rank < g.world, "bad world/rank %d/%d", g.world,
              g.rank);
  BSK_REQUIRE(N % g.world == 0 && M % g.world == 0,
              "nmesh (%d) and neval (%d) must be divisible by world (%d)", N, M, g.world);
  BSK_REQUIRE(g.max_shells >= 1 && g.max_shells <= kMaxChunk, "max_shells must be in [1,%d]",
              kMaxChunk);
  const bool full = (2 * g.ncrop + 1 >= N);
  BSK_REQUIRE(full ? (M == N) : (M >= 2 * g.ncrop + 2),
              "neval=%d incompatible with ncrop=%d at nmesh=%d", M, g.ncrop, N);
  BSK_REQUIRE(g.transposed == 0 || g.transposed == 1, "transposed must be 0 or 1");
  BSK_REQUIRE(!g.transposed || full,
              "a transposed (y-slab) spectrum is for uncropped spectra only (2*ncrop+1 >= nmesh)");

  bsk_plan* p = new bsk_plan();
  p->g = g;
  p->stream = (cudaStream_t)cuda_stream;
  bsk_info& f = p->info;
  f.kx = f.ky = full ? N : 2 * g.ncrop + 1;
  f.kz = full ? N / 2 + 1 : g.ncrop + 1;
  f.nxl = N / g.world;
  f.nx0 = f.nxl * g.rank;
  f.mxl = M / g.world;
  f.mx0 = f.mxl * g.rank;
  // the forward transform runs in float64 whatever the shell precision (a float32 FFT's error
  // is relative to the whole spectrum's rms and would swamp the weak high-k modes); it is
  // chunked over planes so that its work buffers stay below ~1 GiB
  f.fwd_batch = 1;
  for (int64_t b = 1; b <= f.nxl; ++b)
    if (f.nxl % b == 0 && b * (int64_t)N * (N / 2 + 1) * 16 <= (1ll << 30)) f.fwd_batch = b;
  f.fwd_work_complex = f.fwd_batch * (int64_t)N * (N / 2 + 1);
  // transposed: the spectrum cube (and the x-transformed columns of the shells) hold this rank's
  // ky block only; the exchanges around them are all-to-all transposes (x-slabs <-> y-slabs)
  f.kyl = g.transposed ? f.ky / g.world : f.ky;
  f.ky0 = g.transposed ? f.kyl * g.rank : 0;
  f.planes_local_complex = f.nxl * f.ky * f.kz;
  f.planes_all_complex = (int64_t)N * f.kyl * f.kz;
  f.cube_complex = f.kx * f.kyl * f.kz;
  f.xcols_complex_per_shell = (int64_t)M * f.kyl * f.kz;
  f.xplanes_complex_per_shell = g.transposed ? f.mxl * f.ky * f.kz : 0;
  p->use_zpass = (g.fft_precision == BSK_F64) && zpass_supported(M) && !g.no_prune;
  f.planes2d_complex_per_shell =
      p->use_zpass ? f.mxl * (int64_t)M * f.kz : f.mxl * (int64_t)M * (M / 2 + 1);
  f.field_real_per_shell = f.mxl * (int64_t)M * M;
  f.pruned = p->use_zpass ? 1 : 0;

  int rc;
  if ((rc = upload(&p->d_kx, kx_tab, f.kx, p->stream))) return rc;
  if ((rc = upload(&p->d_ky, ky_tab, f.ky, p->stream))) return rc;
  if ((rc = upload(&p->d_kz, kz_tab, f.kz, p->stream))) return rc;
  std::vector<double> ones((size_t)(f.kx > f.kz ? f.kx : f.kz), 1.0);
  if ((rc = upload(&p->d_cx, ones.data(), f.kx, p->stream))) return rc;
  if ((rc = upload(&p->d_cy, ones.data(), f.ky, p->stream))) return rc;
  if ((rc = upload(&p->d_cz, ones.data(), f.kz, p->stream))) return rc;

  {  // forward: batched 2-D real-to-complex over chunks of the local planes (float64)
    long long n[2] = {N, N};
    long long inembed[2] = {N, N};
    long long onembed[2] = {N, N / 2 + 1};
    rc = make_plan_many(&p->fwd2d, 2, n, inembed, 1, (long long)N * N, onembed, 1,
                        (long long)N * (N / 2 + 1), CUFFT_D2Z, f.fwd_batch, p->stream,
                        &p->fft_work_bytes);
    if (rc) return rc;
  }
  {  // forward: strided 1-D complex transform along x on the kept (y,z) columns
    long long cols = f.kyl * f.kz;
    long long n[1] = {N};
    long long emb[1] = {N};
    rc = make_plan_many(&p->fwdx, 1, n, emb, cols, 1, emb, cols, 1, CUFFT_Z2Z, cols, p->stream,
                        &p->fft_work_bytes);
    if (rc) return rc;
  }
  if (p->use_zpass) {
    std::vector<double2> w((size_t)M);
    for (int j = 0; j < M; ++j) {
      const double a = 2.0 * 3.14159265358979323846264338327950288 * (double)j / (double)M;
      w[j] = make_double2(cos(a), sin(a));
    }
    // exact values on the axes and diagonals keep the table symmetric
    for (int j = 0; j < M; ++j) {
      if (j % (M / 4) == 0) {
        const int q = j / (M / 4);
        w[j] = make_double2(q == 0 ? 1.0 : q == 2 ? -1.0 : 0.0, q == 1 ? 1.0 : q == 3 ? -1.0 : 0.0);
      }
    }
    BSK_CUDA(cudaMalloc((void**)&p->d_wtab, sizeof(double2) * (size_t)M));
    // pageable source: the copy returns once staged and is ordered in the legacy stream only; the plan's
    // stream may be a non-blocking one
    BSK_CUDA(cudaMemcpyAsync(p->d_wtab, w.data(), sizeof(double2) * (size_t)M, cudaMemcpyHostToDevice, p->stream));
    BSK_CUDA(cudaStreamSynchronize(p->stream));
  }
  f.fft_work_bytes = (int64_t)p->fft_work_bytes;
  *out = p;
  return BSK_OK;
}

int bsk_plan_destroy(bsk_plan* p) {
  if (!p) return BSK_OK;
  if (p->fwd2d) cufftDestroy(p->fwd2d);
  if (p->fwdx) cufftDestroy(p->fwdx);
  for (auto& kv : p->invx) cufftDestroy(kv.second);
  for (auto& kv : p->inv2d) cufftDestroy(kv.second);
  for (auto& kv : p->invy) cufftDestroy(kv.second);
  cudaFree(p->d_wtab);
  cudaFree(p->d_kx);
  cudaFree(p->d_ky);
  cudaFree(p->d_kz);
  cudaFree(p->d_cx);
  cudaFree(p->d_cy);
  cudaFree(p->d_cz);
  delete p;
  return BSK_OK;
}

int bsk_plan_info(const bsk_plan* p, bsk_info* out) {
  BSK_REQUIRE(p && out, "bsk_plan_info: null argument");
  *out = p->info;
  out->fft_work_bytes = (int64_t)p->fft_work_bytes;
  return BSK_OK;
}

int bsk_set_compensation(bsk_plan* p, const double* cx, const double* cy, const double* cz) {
  BSK_REQUIRE(p, "bsk_set_compensation: null plan");
  std::vector<double> ones((size_t)(p->info.kx > p->info.kz ? p->info.kx : p->info.kz), 1.0);
  const double* sx = cx ? cx : ones.data();
  const double* sy = cy ? cy : ones.data();
  const double* sz = cz ? cz : ones.data();
  BSK_CUDA(cudaMemcpyAsync(p->d_cx, sx, sizeof(double) * p->info.kx, cudaMemcpyHostToDevice, p->stream));
  BSK_CUDA(cudaMemcpyAsync(p->d_cy, sy, sizeof(double) * p->info.ky, cudaMemcpyHostToDevice, p->stream));
  BSK_CUDA(cudaMemcpyAsync(p->d_cz, sz, sizeof(double) * p->info.kz, cudaMemcpyHostToDevice, p->stream));
  BSK_CUDA(cudaStreamSynchronize(p->stream));
  p->has_comp = cx || cy || cz;
  return BSK_OK;
}

int bsk_plan_set_stream(bsk_plan* p, void* cuda_stream) {
  BSK_REQUIRE(p, "bsk_plan_set_stream: null plan");
  cudaStream_t st = (cudaStream_t)cuda_stream;
  if (st == p->stream) return BSK_OK;
  p->stream = st;
  if (p->fwd2d) BSK_FFT(cufftSetStream(p->fwd2d, st));
  if (p->fwdx) BSK_FFT(cufftSetStream(p->fwdx, st));
  for (auto& kv : p->invx) BSK_FFT(cufftSetStream(kv.second, st));
  for (auto& kv : p->inv2d) BSK_FFT(cufftSetStream(kv.second, st));
  for (auto& kv : p->invy) BSK_FFT(cufftSetStream(kv.second, st));
  return BSK_OK;
}

int bsk_fold_even(const void* fields, int precision, int nrows, int neval, int mxl, int fold_x, void* out,
                  int64_t ncell_out, void* cuda_stream) {
  BSK_REQUIRE(fields && out && nrows > 0 && neval >= 4 && neval % 2 == 0 && mxl >= 1,
              "bsk_fold_even: bad argument");
  BSK_REQUIRE(precision == BSK_F32 || precision == BSK_F64, "bsk_fold_even: bad precision");
  const int h = neval / 2 + 1;
  const int nx = fold_x ? h : mxl;
  BSK_REQUIRE(!fold_x || mxl == neval, "bsk_fold_even: the x axis can only be folded when the whole grid is local");
  BSK_REQUIRE(ncell_out >= (int64_t)nx * h * h && ncell_out % 4 == 0,
              "bsk_fold_even: ncell_out must be a multiple of 4 and hold nx*(M/2+1)^2 cells");
  cudaStream_t st = (cudaStream_t)cuda_stream;
  const int64_t total = (int64_t)nrows * ncell_out;
  if (precision == BSK_F64)
    fold_even_kernel<double><<<grid_for(total, 256), 256, 0, st>>>((const double*)fields, (double*)out, nrows, neval,
                                                                    mxl, nx, fold_x, ncell_out);
  else
    fold_even_kernel<float><<<grid_for(total, 256), 256, 0, st>>>((const float*)fields, (float*)out, nrows, neval, mxl,
                                                                   nx, fold_x, ncell_out);
  count_launch();
  BSK_CUDA(cudaGetLastError());
  return BSK_OK;
}

int bsk_forward_local(bsk_plan* p, const void* mesh_slab, int mesh_dtype, void* work,
                      void* convert_work, void* planes_local) {
  BSK_REQUIRE(p && mesh_slab && work && planes_local, "bsk_forward_local: null argument");
  BSK_REQUIRE(mesh_dtype == BSK_F32 || mesh_dtype == BSK_F64, "bad mesh_dtype %d", mesh_dtype);
  BSK_REQUIRE(mesh_dtype == BSK_F64 || convert_work,
              "float32 mesh: convert_work (fwd_batch*N*N float64) is required");
  const bsk_info& f = p->info;
  const int N = p->g.nmesh;
  const int64_t chunk_real = f.fwd_batch * (int64_t)N * N;
  const int64_t chunk_out = f.fwd_batch * f.ky * f.kz;
  const double scale = 1.0 / ((double)N * (double)N * (double)N);
  for (int64_t c = 0; c < f.nxl / f.fwd_batch; ++c) {
    const double* src;
    if (mesh_dtype == BSK_F32) {
      convert_kernel<float, double><<<grid_for(chunk_real, 256), 256, 0, p->stream>>>(
          (const float*)mesh_slab + c * chunk_real, (double*)convert_work, chunk_real);
      count_launch();
      BSK_CUDA(cudaGetLastError());
      src = (const double*)convert_work;
    } else {
      src = (const double*)mesh_slab + c * chunk_real;
    }
    BSK_FFT(cufftExecD2Z(p->fwd2d, (cufftDoubleReal*)src, (cufftDoubleComplex*)work));
    crop_yz_kernel<double><<<grid_for(chunk_out, 256), 256, 0, p->stream>>>(
        (const double2*)work, (double2*)planes_local + c * chunk_out, (int)f.fwd_batch, N,
        (int)f.ky, (int)f.kz, scale, p->d_cy, p->d_cz);
    count_launch();
    BSK_CUDA(cudaGetLastError());
  }
  return BSK_OK;
}

int bsk_forward_finish(bsk_plan* p, void* planes_all, void* cube) {
  BSK_REQUIRE(p && planes_all && cube, "bsk_forward_finish: null argument");
  const bsk_info& f = p->info;
  const int N = p->g.nmesh;
  const int64_t plane = f.kyl * f.kz;
  BSK_FFT(cufftExecZ2Z(p->fwdx, (cufftDoubleComplex*)planes_all, (cufftDoubleComplex*)planes_all,
                       CUFFT_FORWARD));
  crop_x_kernel<double><<<grid_for(f.cube_complex, 256), 256, 0, p->stream>>>(
      (const double2*)planes_all, (double2*)cube, N, (int)f.kx, plane, p->d_cx);
  count_launch();
  BSK_CUDA(cudaGetLastError());
  return BSK_OK;
}

int bsk_modes_per_bin(bsk_plan* p, int nbins, const double* lo, const double* hi, int64_t* counts) {
  BSK_REQUIRE(p && lo && hi && counts && nbins > 0, "bsk_modes_per_bin: bad argument");
  const bsk_info& f = p->info;
  double* d_lo = nullptr;
  double* d_hi = nullptr;
  unsigned long long* d_cnt = nullptr;
  BSK_CUDA(cudaMalloc((void**)&d_lo, sizeof(double) * nbins));
  BSK_CUDA(cudaMalloc((void**)&d_hi, sizeof(double) * nbins));
  BSK_CUDA(cudaMalloc((void**)&d_cnt, sizeof(unsigned long long) * nbins));
  BSK_CUDA(cudaMemcpyAsync(d_lo, lo, sizeof(double) * nbins, cudaMemcpyHostToDevice, p->stream));
  BSK_CUDA(cudaMemcpyAsync(d_hi, hi, sizeof(double) * nbins, cudaMemcpyHostToDevice, p->stream));
  BSK_CUDA(cudaMemsetAsync(d_cnt, 0, sizeof(unsigned long long) * nbins, p->stream));
  mode_count_kernel<<<grid_for(f.cube_complex, 256), 256, 0, p->stream>>>(
      (int)f.kx, (int)f.kyl, (int)f.kz, p->g.nmesh, nbins, d_lo, d_hi, p->d_kx, p->d_ky + f.ky0, p->d_kz,
      d_cnt);
  count_launch();
  BSK_CUDA(cudaGetLastError());
  std::vector<unsigned long long> h((size_t)nbins);
  BSK_CUDA(cudaMemcpyAsync(h.data(), d_cnt, sizeof(unsigned long long) * nbins,
                           cudaMemcpyDeviceToHost, p->stream));
  BSK_CUDA(cudaStreamSynchronize(p->stream));
  for (int b = 0; b < nbins; ++b) counts[b] = (int64_t)h[b];
  cudaFree(d_lo);
  cudaFree(d_hi);
  cudaFree(d_cnt);
  return BSK_OK;
}

int bsk_shells_prepare(bsk_plan* p, int nsh) {
  BSK_REQUIRE(p && nsh >= 1 && nsh <= p->g.max_shells, "bsk_shells_prepare: bad argument");
  cufftHandle h;
  int rc;
  if ((rc = get_invx(p, nsh, &h))) return rc;
  if (p->use_zpass) return get_invy(p, nsh, &h);
  return get_inv2d(p, nsh, &h);
}

static int shells_check(bsk_plan* p, const void* cube, int kind, int nsh, const char* who) {
  BSK_REQUIRE(nsh >= 1 && nsh <= p->g.max_shells, "%s: nsh=%d outside [1,%d]", who, nsh, p->g.max_shells);
  BSK_REQUIRE(kind == BSK_KIND_DATA || kind == BSK_KIND_UNIT || kind == BSK_KIND_KPOW, "%s: bad kind %d", who, kind);
  BSK_REQUIRE(kind != BSK_KIND_DATA || cube, "%s: data kind needs the spectrum cube", who);
  return BSK_OK;
}

int bsk_shells(bsk_plan* p, const void* cube, int kind, double kpow, int nsh, const double* lo,
               const double* hi, void* xcols, void* planes2d, void* fields) {
  BSK_REQUIRE(p && lo && hi && xcols && planes2d && fields, "bsk_shells: null argument");
  BSK_REQUIRE(!p->g.transposed, "bsk_shells: a transposed plan needs bsk_shells_x, the host's all-to-all, bsk_shells_yz");
  int rc;
  if ((rc = shells_check(p, cube, kind, nsh, "bsk_shells"))) return rc;
  if ((rc = shells_x_dispatch(p, cube, kind, kpow, nsh, lo, hi, xcols))) return rc;
  // this rank's planes of the x-transformed columns
  const size_t csize = p->g.fft_precision == BSK_F64 ? 16 : 8;
  const char* xplanes = (const char*)xcols + csize * (size_t)p->info.mx0 * nsh * p->info.ky * p->info.kz;
  return shells_yz_dispatch(p, nsh, xplanes, planes2d, fields);
}

int bsk_shells_x(bsk_plan* p, const void* cube, int kind, double kpow, int nsh, const double* lo,
                 const double* hi, void* xcols) {
  BSK_REQUIRE(p && lo && hi && xcols, "bsk_shells_x: null argument");
  int rc;
  if ((rc = shells_check(p, cube, kind, nsh, "bsk_shells_x"))) return rc;
  return shells_x_dispatch(p, cube, kind, kpow, nsh, lo, hi, xcols);
}

int bsk_shells_yz(bsk_plan* p, int nsh, const void* xplanes, void* planes2d, void* fields) {
  BSK_REQUIRE(p && xplanes && planes2d && fields, "bsk_shells_yz: null argument");
  BSK_REQUIRE(nsh >= 1 && nsh <= p->g.max_shells, "bsk_shells_yz: nsh=%d outside [1,%d]", nsh, p->g.max_shells);
  return shells_yz_dispatch(p, nsh, xplanes, planes2d, fields);
}

}  // extern "C"
