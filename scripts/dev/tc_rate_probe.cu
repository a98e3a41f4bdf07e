// tcgen05.mma issue-rate probe (sm_100a): cycles per MMA for M=128, A in TMEM or shared memory,
// measured with a clean warp-uniform issue loop (one elected lane, 16 MMAs per loop trip).
//   nvcc -O2 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tc_rate_probe tc_rate_probe.cu
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); fflush(stdout); exit(2); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t elect_one() {
  uint32_t pred;
  asm volatile("{\n.reg .b32 rx;\n.reg .pred px;\nelect.sync rx|px, 0xffffffff;\nselp.u32 %0, 1, 0, px;\n}\n" : "=r"(pred));
  return pred;
}
__host__ __device__ inline uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) | ((uint64_t)1 << 46);
}
__host__ __device__ constexpr uint32_t make_idesc(int fmt, int M, int N) {
  return (1u << 4) | ((uint32_t)fmt << 7) | ((uint32_t)fmt << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
template <int KIND>  // 0 f16, 2 tf32
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc) {
  if constexpr (KIND == 2)
    asm volatile("tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, 1;" ::"r"(d), "r"(a), "l"(b), "r"(idesc) : "memory");
  else
    asm volatile("tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, 1;" ::"r"(d), "r"(a), "l"(b), "r"(idesc) : "memory");
}
__device__ __forceinline__ void mma_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc) {
  asm volatile("tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, 1;" ::"r"(d), "l"(a), "l"(b), "r"(idesc) : "memory");
}

// MODE 0: TS one accumulator; 1: TS two accumulators alternating; 2: SS one accumulator;
// MODE 3: TS, three operand pairs rotating (the 3xTF32 pattern: different A columns / B images)
template <int KIND, int N, int MODE>
__global__ void __launch_bounds__(128, 1) rate_kernel(long long* out, int slot, int trips) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tslot;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 24576 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3f800000u;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tslot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tb = tslot;
  if (warp == 0) {
    const uint32_t leader = elect_one();
    const uint64_t b0 = make_desc(smem_u32(smem), 128, 256), b1 = make_desc(smem_u32(smem + 8192), 128, 256);
    const uint64_t a0 = make_desc(smem_u32(smem + 16384), 128, 256);
    constexpr uint32_t id = make_idesc(KIND, 128, N);
    const uint32_t A0 = tb + 256, A1 = tb + 264, D0 = tb, D1 = tb + 128;
    const long long t0 = clock64();
    for (int t = 0; t < trips; ++t) {
      if (leader) {
#pragma unroll
        for (int u = 0; u < 16; ++u) {
          if constexpr (MODE == 0) mma_ts<KIND>(D0, A0, b0, id);
          else if constexpr (MODE == 1) mma_ts<KIND>((u & 1) ? D1 : D0, A0, b0, id);
          else if constexpr (MODE == 2) mma_ss(D0, a0, b0, id);
          else mma_ts<KIND>(D0, (u % 3 == 0) ? A1 : A0, (u % 3 == 1) ? b1 : b0, id);
        }
      }
      __syncwarp();
    }
    if (leader) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    __syncwarp();
    uint32_t ok = 0;
    for (int it = 0; it < (1 << 26) && !ok; ++it)
      asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(ok) : "r"(smem_u32(&bar)) : "memory");
    const long long t1 = clock64();
    if (leader) out[slot] = ok ? (t1 - t0) : -1;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tb) : "memory");
}

template <int KIND, int N, int MODE>
static void run(const char* name, long long* d, int& slot) {
  const int trips = 512;
  CK(cudaFuncSetAttribute(rate_kernel<KIND, N, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 24576));
  rate_kernel<KIND, N, MODE><<<1, 128, 24576>>>(d, slot, trips);
  CK(cudaDeviceSynchronize());
  long long c;
  CK(cudaMemcpy(&c, d + slot, 8, cudaMemcpyDeviceToHost));
  printf("%-34s N=%3d : %7.1f cycles/MMA\n", name, N, (double)c / (trips * 16.0));
  ++slot;
}

int main() {
  long long* d;
  CK(cudaMalloc(&d, 8 * 64));
  int slot = 0;
  run<2, 16, 0>("tf32 TS", d, slot);  run<2, 32, 0>("tf32 TS", d, slot);  run<2, 40, 0>("tf32 TS", d, slot);
  run<2, 48, 0>("tf32 TS", d, slot);  run<2, 64, 0>("tf32 TS", d, slot);  run<2, 96, 0>("tf32 TS", d, slot);
  run<2, 128, 0>("tf32 TS", d, slot); run<2, 256, 0>("tf32 TS", d, slot);
  run<2, 48, 1>("tf32 TS 2 accumulators", d, slot);  run<2, 128, 1>("tf32 TS 2 accumulators", d, slot);
  run<2, 48, 2>("tf32 SS", d, slot);  run<2, 128, 2>("tf32 SS", d, slot);  run<2, 256, 2>("tf32 SS", d, slot);
  run<2, 48, 3>("tf32 TS rotating operands", d, slot);
  run<0, 16, 0>("f16 TS", d, slot);   run<0, 48, 0>("f16 TS", d, slot);   run<0, 128, 0>("f16 TS", d, slot);  run<0, 256, 0>("f16 TS", d, slot);
  fflush(stdout);
  return 0;
}
