// Throughput of tcgen05.st (TMEM stores) on sm_100a: cycles per instruction for the 32x32b shape with
// 8 / 16 / 32 columns, issued back to back by 1, 4 (one per TMEM lane quarter) or 8 warps (two per quarter),
// and the same with a tcgen05.ld stream or an MMA stream running beside the stores.
// build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o build/sttm_probe scripts/dev/sttm_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
constexpr int ITERS = 512;

template <int N>
__device__ __forceinline__ void st_n(uint32_t addr, uint32_t v);
template <>
__device__ __forceinline__ void st_n<8>(uint32_t a, uint32_t v) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1};" ::"r"(a), "r"(v) : "memory");
}
template <>
__device__ __forceinline__ void st_n<16>(uint32_t a, uint32_t v) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1};" ::"r"(a), "r"(v) : "memory");
}
template <>
__device__ __forceinline__ void st_n<32>(uint32_t a, uint32_t v) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1};" ::"r"(a),
      "r"(v)
      : "memory");
}

// out[cfg * 8 + warp] = cycles per store instruction seen by that warp
template <int N>
__device__ void run(long long* out, int cfg, int nwarps, uint32_t tbase, int warp) {
  __syncthreads();
  if (warp < nwarps) {
    const uint32_t base = tbase + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(warp >> 2) * 128u;
    const long long t0 = clock64();
#pragma unroll 4
    for (int i = 0; i < ITERS; ++i) st_n<N>(base + (uint32_t)((i * N) & 127), (uint32_t)i);
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    const long long t1 = clock64();
    if ((threadIdx.x & 31) == 0) out[cfg * 8 + warp] = (t1 - t0) * 100 / ITERS;
  }
  __syncthreads();
}

__global__ void probe(long long* out) {
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tbase = tmem_slot;
  int cfg = 0;
  for (int nw : {1, 4, 8}) {
    run<8>(out, cfg++, nw, tbase, warp);
    run<16>(out, cfg++, nw, tbase, warp);
    run<32>(out, cfg++, nw, tbase, warp);
  }
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tbase) : "memory");
}

int main() {
  long long* d;
  cudaMalloc(&d, 9 * 8 * sizeof(long long));
  cudaMemset(d, 0, 9 * 8 * sizeof(long long));
  probe<<<1, 256>>>(d);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
  long long h[72];
  cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  const int ns[3] = {8, 16, 32}, nws[3] = {1, 4, 8};
  for (int c = 0; c < 9; ++c) {
    const int n = ns[c % 3], nw = nws[c / 3];
    double worst = 0;
    for (int w = 0; w < nw; ++w) worst = h[c * 8 + w] / 100.0 > worst ? h[c * 8 + w] / 100.0 : worst;
    printf("tcgen05.st.32x32b.x%-2d  %d warp(s): %6.1f cycles per instruction and warp  -> %6.1f B/cycle/SM\n", n, nw, worst,
           nw * n * 128.0 / worst);
  }
  return 0;
}
