// Shared-memory wavefront probe: how many cycles does a warp-wide LDS.128 cost when several lanes
// read the same 16-byte chunk?  (decides the pair-row layout of the tensor-core generator)
// build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o build/lds_probe scripts/dev/lds_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

constexpr int STRIDE = 528;   // bytes between rows (as RAW_STRIDE in contract_tc.cuh)
constexpr int ITERS = 2048;

__device__ __forceinline__ float4 lds128(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}

__global__ void probe(int pattern, long long* out, float* sink) {
  extern __shared__ __align__(128) unsigned char smem[];
  for (int i = threadIdx.x; i < 48 * STRIDE / 4; i += blockDim.x) reinterpret_cast<float*>(smem)[i] = (float)i;
  __syncthreads();
  const int l = threadIdx.x & 31;
  uint32_t off;
  switch (pattern) {
    case 0: off = l * 16; break;                       // 32 distinct chunks, contiguous
    case 1: off = 0; break;                            // one chunk for all lanes
    case 2: off = (l & 7) * STRIDE; break;             // 8 rows; every quarter-warp sees all 8
    case 3: off = (l >> 3) * STRIDE; break;            // 4 rows; a quarter-warp reads one row
    case 4: off = (l >> 2) * STRIDE; break;            // 8 rows; a quarter-warp reads two rows
    case 5: off = l * STRIDE; break;                   // 32 distinct rows (today's generator)
    case 6: off = (l & 15) * STRIDE; break;            // 16 rows
    case 7: off = (l >> 1) * STRIDE; break;            // 16 rows, pairs of lanes
    case 8: off = (l & 3) * STRIDE; break;             // 4 rows, interleaved
    default: off = 0;
  }
  const uint32_t base = (uint32_t)__cvta_generic_to_shared(smem) + off;
  float4 acc = make_float4(0, 0, 0, 0);
  __syncthreads();
  const long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const float4 v = lds128(base + u * 16);
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
  }
  const long long t1 = clock64();
  __syncthreads();
  if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
  if (acc.x + acc.y + acc.z + acc.w == 12345.f) sink[0] = acc.x;
}

int main() {
  long long* d_out; float* d_sink;
  cudaMalloc(&d_out, 8 * 256); cudaMalloc(&d_sink, 4);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 48 * STRIDE);
  const char* names[] = {"32 distinct contiguous", "1 chunk (broadcast)", "8 rows, l&7", "4 rows, l>>3", "8 rows, l>>2",
                         "32 distinct rows", "16 rows, l&15", "16 rows, l>>1", "4 rows, l&3"};
  for (int nw : {8, 16}) {
    for (int p = 0; p < 9; ++p) {
      probe<<<1, nw * 32, 48 * STRIDE>>>(p, d_out, d_sink);
      long long h = 0;
      cudaMemcpy(&h, d_out, 8, cudaMemcpyDeviceToHost);
      printf("warps %2d  pattern %d (%-24s): %.2f cycles per warp-LDS.128 (SM-wide)\n", nw, p, names[p],
             (double)h / ((double)ITERS * 8 * nw));
    }
  }
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
