// Shared helpers for the bskit_b200 CUDA library (sm_100a).
#pragma once

#include <cuda_runtime.h>
#include <cufft.h>

#include <atomic>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <map>
#include <string>
#include <vector>

#include "../../include/bskit_b200.h"

namespace bsk {

void set_error(const char* fmt, ...);
extern std::atomic<int64_t> g_launches;
inline void count_launch(int n = 1) { g_launches.fetch_add(n, std::memory_order_relaxed); }

#define BSK_CUDA(expr)                                                               \
  do {                                                                               \
    cudaError_t e__ = (expr);                                                        \
    if (e__ != cudaSuccess) {                                                        \
      bsk::set_error("%s:%d %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(e__)); \
      return BSK_ERR_CUDA;                                                           \
    }                                                                                \
  } while (0)

#define BSK_FFT(expr)                                                        \
  do {                                                                       \
    cufftResult r__ = (expr);                                                \
    if (r__ != CUFFT_SUCCESS) {                                              \
      bsk::set_error("%s:%d %s -> cufft %d", __FILE__, __LINE__, #expr, (int)r__); \
      return BSK_ERR_CUFFT;                                                  \
    }                                                                        \
  } while (0)

#define BSK_REQUIRE(cond, ...)      \
  do {                              \
    if (!(cond)) {                  \
      bsk::set_error(__VA_ARGS__);  \
      return BSK_ERR_ARG;           \
    }                               \
  } while (0)

// signed mode number of cropped-cube index j on an axis that keeps `k` of `n` modes
__host__ __device__ inline int mode_of(int j, int k, int n) {
  if (k == n) return (j <= n / 2 - (n % 2 == 0 ? 1 : 0)) ? j : j - n;  // fftfreq order
  int nc = (k - 1) / 2;
  return j <= nc ? j : j - k;
}

bool zpass_supported(int M);
int zpass_run(int M, bool store_f32, const void* xcols, void* ycols, void* fields, int Ky, int Kz,
              int nsh, int mx0, int mxl, const double2* wtab, cufftHandle yplan, cudaStream_t st);

}  // namespace bsk

struct bsk_plan {
  bsk_geometry g{};
  bsk_info info{};
  cudaStream_t stream = nullptr;
  // device tables (float64)
  double* d_kx = nullptr;
  double* d_ky = nullptr;
  double* d_kz = nullptr;
  double* d_cx = nullptr;  // compensation (all ones when unset)
  double* d_cy = nullptr;
  double* d_cz = nullptr;
  bool has_comp = false;
  // cuFFT plans
  cufftHandle fwd2d = 0;   // batched 2-D R2C/D2Z over local planes
  cufftHandle fwdx = 0;    // strided 1-D C2C/Z2Z along x, length N
  std::map<int, cufftHandle> invx;   // by nsh: strided 1-D inverse along x, length M
  std::map<int, cufftHandle> inv2d;  // by nsh: batched 2-D C2R/Z2D
  std::map<int, cufftHandle> invy;   // by nsh: pruned path, 1-D inverse along y on kept kz columns
  bool use_zpass = false;            // pruned y/z passes (float64 transforms, power-of-two M)
  double2* d_wtab = nullptr;         // e^{2 pi i j/M}, j < M
  size_t fft_work_bytes = 0;
};
