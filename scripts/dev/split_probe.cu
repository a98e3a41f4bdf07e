// Cost of forming pair products and their tf32 hi/lo split on the CUDA cores (sm_100a):
// which instruction mix is cheapest for the tensor-core contraction's generator warps?
// build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o build/split_probe scripts/dev/split_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ float2 fmul2(float2 a, float2 b) { return __fmul2_rn(a, b); }
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) { return __ffma2_rn(a, b, c); }

template <int V>
__device__ __forceinline__ void body(const float2 (&a)[4], const float2 (&b)[4], uint32_t (&hi)[8], uint32_t (&lo)[8]) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 pr = fmul2(a[i], b[i]);
    float2 ph;
    if (V == 0 || (V == 2 && i < 2)) {           // integer round-to-nearest of the top 11 bits
      ph = make_float2(__uint_as_float((__float_as_uint(pr.x) + 0x1000u) & 0xFFFFE000u),
                       __uint_as_float((__float_as_uint(pr.y) + 0x1000u) & 0xFFFFE000u));
    } else if (V == 1 || V == 2) {               // Veltkamp split on the FMA pipe (packed)
      const float2 t = fmul2(pr, make_float2(8193.f, 8193.f));
      const float2 u = ffma2(pr, make_float2(-1.f, -1.f), t);     // t - pr
      ph = ffma2(u, make_float2(-1.f, -1.f), t);                  // t - u
    } else if (V == 3) {                         // integer add through IMAD (FMA pipe), AND on the ALU
      uint32_t x, y;
      asm("mad.lo.u32 %0, %1, 1, 0x1000;" : "=r"(x) : "r"(__float_as_uint(pr.x)));
      asm("mad.lo.u32 %0, %1, 1, 0x1000;" : "=r"(y) : "r"(__float_as_uint(pr.y)));
      ph = make_float2(__uint_as_float(x & 0xFFFFE000u), __uint_as_float(y & 0xFFFFE000u));
    } else if (V == 5) {                         // baseline: no split at all (loop + consumer overhead)
      ph = pr;
    } else {                                     // V == 4: truncation split (1 AND per element)
      ph = make_float2(__uint_as_float(__float_as_uint(pr.x) & 0xFFFFE000u),
                       __uint_as_float(__float_as_uint(pr.y) & 0xFFFFE000u));
    }
    const float2 pl = V == 5 ? pr : ffma2(ph, make_float2(-1.f, -1.f), pr);
    hi[2 * i] = __float_as_uint(ph.x); hi[2 * i + 1] = __float_as_uint(ph.y);
    lo[2 * i] = __float_as_uint(pl.x); lo[2 * i + 1] = __float_as_uint(pl.y);
  }
}

template <int V>
__global__ void probe(long long* out, uint32_t* sink, int iters) {
  float2 a[4], b[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    a[i] = make_float2(1.0f + threadIdx.x * 1e-3f + i, 0.5f + i * 0.25f);
    b[i] = make_float2(0.75f + threadIdx.x * 2e-3f, 1.25f - i * 0.125f);
  }
  uint32_t acc = 0;
  __syncthreads();
  const long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < iters; ++it) {
    uint32_t hi[8], lo[8];
    body<V>(a, b, hi, lo);
#pragma unroll
    for (int i = 0; i < 8; ++i) acc ^= hi[i] ^ lo[i];        // one LOP3 per element keeps the results live
#pragma unroll
    for (int i = 0; i < 4; ++i) asm volatile("" : "+f"(a[i].x), "+f"(a[i].y), "+f"(b[i].x), "+f"(b[i].y));  // opaque inputs
  }
  const long long t1 = clock64();
  __syncthreads();
  if (threadIdx.x == 0) out[0] = t1 - t0;
  if (acc == 0x12345u) sink[0] = acc;
}

template <int V>
void run(const char* name, long long* d_out, uint32_t* d_sink) {
  const int iters = 4096;
  for (int nw : {4, 8, 16}) {
    probe<V><<<1, nw * 32, 0>>>(d_out, d_sink, iters);
    long long h = 0;
    cudaMemcpy(&h, d_out, 8, cudaMemcpyDeviceToHost);
    printf("%-44s warps %2d: %.1f cycles per 8-product body per SMSP-warp slot (%.1f per body, SM-wide)\n", name, nw,
           (double)h / iters / (nw / 4.0), (double)h / iters / nw);
  }
}

int main() {
  long long* d_out; uint32_t* d_sink;
  cudaMalloc(&d_out, 64); cudaMalloc(&d_sink, 4);
  run<0>("V0 int add+and (today)", d_out, d_sink);
  run<1>("V1 Veltkamp packed (FMA pipe only)", d_out, d_sink);
  run<2>("V2 half int, half Veltkamp", d_out, d_sink);
  run<3>("V3 IMAD add + AND", d_out, d_sink);
  run<4>("V4 truncation split (AND only)", d_out, d_sink);
  run<5>("V5 products only (overhead baseline)", d_out, d_sink);
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
