// Probe: FP32 FMA issue throughput on B200 — scalar FFMA vs packed FFMA2 (fma.rn.f32x2).
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void __launch_bounds__(256) probe(float* out, int iters, float s) {
  float a[16];
  float2 b[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) { a[i] = threadIdx.x * 1e-3f + i; b[i] = make_float2(a[i], a[i] + 1.f); }
  float m = s, c = s * 0.5f;
  float2 m2 = make_float2(s, s * 1.01f), c2 = make_float2(c, c);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      if (MODE == 0) {
#pragma unroll
        for (int i = 0; i < 16; ++i) a[i] = fmaf(a[i], m, c);
      } else {
#pragma unroll
        for (int i = 0; i < 16; ++i) b[i] = __ffma2_rn(b[i], m2, c2);
      }
    }
  }
  float acc = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) acc += (MODE == 0) ? a[i] : (b[i].x + b[i].y);
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}
int main() {
  float* d; cudaMalloc(&d, 148 * 8 * 256 * 4);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  int iters = 20000;
  for (int mode = 0; mode < 2; ++mode)
    for (int ctas_per_sm : {1, 2, 4}) {
      int grid = 148 * ctas_per_sm;
      for (int rep = 0; rep < 2; ++rep) {
        cudaEventRecord(e0);
        if (mode == 0) probe<0><<<grid, 256>>>(d, iters, 0.999f); else probe<1><<<grid, 256>>>(d, iters, 0.999f);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
      }
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      double fmas = (double)grid * 256 * iters * 8 * 16 * (mode == 0 ? 1 : 2);
      printf("mode=%s ctas/sm=%d  %.3f ms  %.2f TFMA/s (%.1f TFLOP/s)\n", mode ? "FFMA2" : "FFMA ", ctas_per_sm, ms,
             fmas / ms / 1e9, 2 * fmas / ms / 1e9);
    }
  return 0;
}
