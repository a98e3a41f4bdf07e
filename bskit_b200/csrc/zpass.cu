// Pruned shell synthesis for power-of-two evaluation grids (sm_100a).
//
// The reference materialises, per k-bin, a full zero-padded half spectrum and runs a full 3-D
// c2r on it (bskit/main.py:1859-1861).  A shell only has modes with |n_axis| <= n_c, so here
//   x pass : cuFFT on the Ky*Kz kept columns                      (spectral.cu)
//   y pass : scatter_y_kernel -> [plane][kz][y] zero-padded in y only, cuFFT Z2Z along y
//            on the Kz kept columns (a factor (M/2+1)/Kz less data than the 2-D transform)
//   z pass : zpass_c2r_kernel — reads the Kz kept coefficients of a row, does the length-M
//            real inverse FFT in shared memory / registers in float64 (zfft_core.h) and writes
//            the row once, already narrowed to the storage dtype.
// HBM traffic per shell drops from ~5 full-grid passes to one N^3 write plus O(N^2 Kz) reads.
#include "common.cuh"
#include "zfft_core.h"

namespace bsk {

using zfft::cplx;

// [x][nsh][Kz][Ky] (after the x transform; ky contiguous) -> [nsh][mxl][Kz][M]  (y contiguous,
// zero padded): contiguous reads and writes.  One CTA per (shell, plane): no 64-bit divisions in
// the element loop (M is a power of two on this path: log2m).
__global__ void __launch_bounds__(256)
scatter_y_kernel(const double2* __restrict__ xcols, double2* __restrict__ ycols,
                 int M, int log2m, int Ky, int Kz, int nsh, int mx0, int mxl) {
  const int nc = (Ky - 1) / 2;
  const int xl = blockIdx.x % mxl, s = blockIdx.x / mxl;
  const double2* __restrict__ src = xcols + ((int64_t)(mx0 + xl) * nsh + s) * Kz * Ky;
  double2* __restrict__ dst = ycols + ((int64_t)s * mxl + xl) * Kz * M;
  const int total = Kz * M;
  for (int i = threadIdx.x; i < total; i += blockDim.x) {
    const int iy = i & (M - 1), kz = i >> log2m;
    int jy;
    if (Ky == M) jy = iy;
    else if (iy <= nc) jy = iy;
    else if (iy >= M - nc) jy = iy - M + Ky;
    else jy = -1;
    double2 v = make_double2(0.0, 0.0);
    if (jy >= 0) v = src[kz * Ky + jy];
    dst[i] = v;
  }
}

constexpr int kZThreads = 128;  // small CTAs: 4-5 resident per SM, cheap barriers

__device__ __forceinline__ int padidx(int i) { return i + (i >> 4); }  // 1 pad per 16 complex

// One Stockham stage of radix R on the rows of this CTA.  TPR threads cooperate on a row.
// In the first stage (p == 1) only packed entries with index < Kz or > H-Kz can be non-zero
// (the rest were never written): they are taken as zero without touching shared memory.
// The LAST stage writes its outputs straight to global memory from registers (x[2n] = Re z[n],
// x[2n+1] = Im z[n]; lanes hold consecutive n, so each warp store covers whole 128-byte lines)
// instead of going back through shared memory.
template <int R, int H, int TPR, bool LAST, typename TS>
__device__ __forceinline__ void stockham_stage(cplx* __restrict__ row, int lt, int p, bool active,
                                               int Kz,
                                               const cplx* __restrict__ tw /* smem [R][p]: e^{2 pi i k m/(pR)} */,
                                               TS* __restrict__ out_row) {
  constexpr int T = H / R;          // butterflies per row in this stage
  constexpr int PER = T / TPR;      // butterflies per thread (TPR = H/16, so PER = 16/R)
  cplx v[PER][R];
  if (active) {
#pragma unroll
  for (int b = 0; b < PER; ++b) {
    const int i = lt + b * TPR;
    const int k = i & (p - 1);
#pragma unroll
    for (int m = 0; m < R; ++m) {
      const int idx = i + m * T;
      cplx u{0.0, 0.0};
      if (p > 1 || idx < Kz || idx > H - Kz) u = row[padidx(idx)];
      if (m > 0 && p > 1) {
        u = zfft::cmul(u, tw[m * p + k]);   // lanes with consecutive k: conflict-free
      }
      v[b][m] = u;
    }
  }
  }
  __syncthreads();  // every input of the stage has been read: the buffer can be overwritten
  if (active) {
#pragma unroll
  for (int b = 0; b < PER; ++b) {
    zfft::dft_inverse_bitrev<R>(v[b]);
    const int i = lt + b * TPR;
    const int k = i & (p - 1);
    const int j = (i - k) * R + k;
    if constexpr (LAST) {
#pragma unroll
      for (int m = 0; m < R; ++m) {
        const cplx z = v[b][zfft::bitrev<R>(m)];
        TS* dst = out_row + 2 * (j + m * p);
        if (sizeof(TS) == 4)
          *reinterpret_cast<float2*>(dst) = make_float2((float)z.x, (float)z.y);
        else
          *reinterpret_cast<double2*>(dst) = make_double2(z.x, z.y);
      }
    } else {
#pragma unroll
      for (int m = 0; m < R; ++m) row[padidx(j + m * p)] = v[b][zfft::bitrev<R>(m)];
    }
  }
  }
  if constexpr (!LAST) __syncthreads();
}

template <int H, int TPR, int P0, int REM>
struct Stages {
  static constexpr int R = (REM % 16 == 0) ? 16 : (REM % 8 == 0) ? 8 : (REM % 4 == 0) ? 4 : 2;
  // per-stage twiddle tables are stored back to back: [R][P0] entries each
  template <typename TS>
  static __device__ __forceinline__ void run(cplx* row, int lt, bool active, int Kz,
                                             const cplx* tw, TS* out_row) {
    stockham_stage<R, H, TPR, (REM / R == 1), TS>(row, lt, P0, active, Kz, tw, out_row);
    if constexpr (REM / R > 1)
      Stages<H, TPR, P0 * R, REM / R>::run(row, lt, active, Kz, tw + R * P0, out_row);
  }
  static __device__ __forceinline__ void fill(cplx* tw, const double2* __restrict__ wtab, int tid,
                                              int nthreads) {
    for (int idx = tid; idx < R * P0; idx += nthreads) {
      const int m = idx / P0, k = idx % P0;
      const double2 w = wtab[(2 * (H / (P0 * R)) * k * m) & (2 * H - 1)];
      tw[idx] = cplx{w.x, w.y};
    }
    if constexpr (REM / R > 1) Stages<H, TPR, P0 * R, REM / R>::fill(tw + R * P0, wtab, tid, nthreads);
  }
};

// M = 2H.  Rows are (plane, y); a CTA handles RPC consecutive y of one plane.
template <int H, typename TS>
__global__ void __launch_bounds__(kZThreads, 5)
zpass_c2r_kernel(const double2* __restrict__ ycols,  // [planes][Kz][M]
                 TS* __restrict__ fields,            // [planes][M][M]
                 int Kz, int64_t planes, const double2* __restrict__ wtab) {
  constexpr int M = 2 * H;
  constexpr int TPR = (H / 16) > 0 ? (H / 16) : 1;   // threads per row
  constexpr int RPC = kZThreads / TPR > 32 ? 32 : kZThreads / TPR;  // rows per CTA
  constexpr int NT = RPC * TPR;                      // active threads
  constexpr int ROWLEN = (H + 1 + ((H + 1) >> 4) + 1) | 1;  // padded complex per row, odd
  extern __shared__ __align__(16) unsigned char zsm[];
  cplx* wsm = reinterpret_cast<cplx*>(zsm);          // pack twiddles e^{2 pi i k/M}, k <= H/2
  cplx* tws = wsm + (H / 2 + 1);                     // per-stage twiddle tables (< 2H entries)
  cplx* sm = tws + 2 * H;                            // RPC padded rows

  const int tid = threadIdx.x;
  for (int j = tid; j <= H / 2; j += blockDim.x) {
    const double2 w = wtab[j];
    wsm[j] = cplx{w.x, w.y};
  }
  Stages<H, TPR, 1, H>::fill(tws, wtab, tid, blockDim.x);
  __syncthreads();
  const int64_t groups_per_plane = M / RPC;
  const int64_t ngroups = planes * groups_per_plane;
  // The kept coefficients of the NEXT row group are fetched into registers while the current
  // group is transformed (the loads are the only long-latency operation of the loop); groups
  // with more than NPRE elements per thread (bins up to Nyquist) load in place instead.
  constexpr int NPRE = 4;
  const bool prefetch = RPC * Kz <= NPRE * kZThreads;
  double2 pre[NPRE];
  auto fetch = [&](int64_t g) {
    const int64_t pl = g / groups_per_plane;
    const int yy = (int)(g - pl * groups_per_plane) * RPC;
#pragma unroll
    for (int q = 0; q < NPRE; ++q) {
      const int e = tid + q * kZThreads;
      if (e < RPC * Kz) pre[q] = ycols[((int64_t)pl * Kz + e / RPC) * M + yy + e % RPC];
    }
  };
  if (prefetch && (int64_t)blockIdx.x < ngroups) fetch(blockIdx.x);
  for (int64_t grp = blockIdx.x; grp < ngroups; grp += gridDim.x) {
    const int64_t plane = grp / groups_per_plane;
    const int y0 = (int)(grp - plane * groups_per_plane) * RPC;
    // ---- the kept coefficients X[kz], kz < Kz (entries >= Kz are implicitly zero)
    if (prefetch) {
#pragma unroll
      for (int q = 0; q < NPRE; ++q) {
        const int e = tid + q * kZThreads;
        if (e < RPC * Kz) sm[(e % RPC) * ROWLEN + padidx(e / RPC)] = cplx{pre[q].x, pre[q].y};
      }
      if (grp + gridDim.x < ngroups) fetch(grp + gridDim.x);
    } else {
      for (int e = tid; e < RPC * Kz; e += blockDim.x) {
        const int r = e % RPC, kz = e / RPC;
        const double2 g = ycols[((int64_t)plane * Kz + kz) * M + y0 + r];
        sm[r * ROWLEN + padidx(kz)] = cplx{g.x, g.y};
      }
    }
    __syncthreads();
    if (tid < NT) {
      const int r = tid / TPR, lt = tid % TPR;
      cplx* row = sm + r * ROWLEN;
      // ---- pack the Hermitian half spectrum into a complex sequence of length H (in place)
      for (int k = lt; k <= H / 2; k += TPR) {
        if (k >= Kz && H - k >= Kz) continue;         // both inputs zero -> both outputs zero
        const cplx zero{0.0, 0.0};
        const cplx xk = k < Kz ? row[padidx(k)] : zero;
        const cplx xhk = H - k < Kz ? row[padidx(H - k)] : zero;
        cplx zk, zhk;
        zfft::pack_pair(xk, xhk, wsm[k], zk, zhk);
        row[padidx(k)] = zk;
        if (k != 0 && k != H - k) row[padidx(H - k)] = zhk;
      }
    }
    __syncthreads();
    // ---- Stockham stages: every thread runs the barriers, threads without a row do no work;
    //      the last stage stores the finished rows to global memory
    {
      const bool active = tid < NT;
      const int r = active ? tid / TPR : 0;
      TS* out_row = fields + ((int64_t)plane * M + y0 + r) * M;
      Stages<H, TPR, 1, H>::run(sm + r * ROWLEN, tid % TPR, active, Kz, tws, out_row);
    }
    // the last stage synchronised after reading shared memory, so the next group may load
  }
}

template <int H, typename TS>
static int launch_zpass(const void* ycols, void* fields, int Kz, int64_t planes, const double2* wtab,
                        cudaStream_t st) {
  constexpr int TPR = (H / 16) > 0 ? (H / 16) : 1;
  constexpr int RPC = kZThreads / TPR > 32 ? 32 : kZThreads / TPR;
  constexpr int ROWLEN = (H + 1 + ((H + 1) >> 4) + 1) | 1;
  const size_t smem = ((size_t)RPC * ROWLEN + 2 * H + H / 2 + 1) * sizeof(cplx);
  BSK_CUDA(cudaFuncSetAttribute(zpass_c2r_kernel<H, TS>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                (int)smem));
  const int64_t ngroups = planes * (2 * H / RPC);
  const int grid = (int)(ngroups < 148 * 16 ? ngroups : 148 * 16);
  zpass_c2r_kernel<H, TS><<<grid, kZThreads, smem, st>>>((const double2*)ycols, (TS*)fields, Kz, planes, wtab);
  count_launch();
  BSK_CUDA(cudaGetLastError());
  return BSK_OK;
}

bool zpass_supported(int M) {
  return M == 64 || M == 128 || M == 256 || M == 512 || M == 1024 || M == 2048;
}

// ycols scratch must hold nsh*mxl*Kz*M complex128; wtab = e^{2 pi i j/M}, j < M
int zpass_run(int M, bool store_f32, const void* xcols, void* ycols, void* fields, int Ky, int Kz,
              int nsh, int mx0, int mxl, const double2* wtab, cufftHandle yplan, cudaStream_t st) {
  int log2m = 0;
  while ((1 << log2m) < M) ++log2m;
  scatter_y_kernel<<<nsh * mxl, 256, 0, st>>>((const double2*)xcols, (double2*)ycols, M, log2m, Ky, Kz, nsh, mx0, mxl);
  count_launch();
  BSK_CUDA(cudaGetLastError());
  BSK_FFT(cufftExecZ2Z(yplan, (cufftDoubleComplex*)ycols, (cufftDoubleComplex*)ycols, CUFFT_INVERSE));
  const int64_t planes = (int64_t)nsh * mxl;
#define BSK_ZP(HH)                                                                          \
  case 2 * HH:                                                                              \
    return store_f32 ? launch_zpass<HH, float>(ycols, fields, Kz, planes, wtab, st)         \
                     : launch_zpass<HH, double>(ycols, fields, Kz, planes, wtab, st);
  switch (M) {
    BSK_ZP(32)
    BSK_ZP(64)
    BSK_ZP(128)
    BSK_ZP(256)
    BSK_ZP(512)
    BSK_ZP(1024)
    default:
      set_error("zpass_run: unsupported M=%d", M);
      return BSK_ERR_ARG;
  }
#undef BSK_ZP
}

}  // namespace bsk
