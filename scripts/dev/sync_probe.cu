// Latencies of the synchronisation primitives the tensor-core contraction uses (sm_100a):
// what does one producer -> consumer hand-over cost?
// build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o build/sync_probe scripts/dev/sync_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
#define T0() const long long t0 = clock64()
#define T1(slot) do { const long long t1 = clock64(); if (threadIdx.x == 0) out[slot] = (t1 - t0) / ITERS; } while (0)
constexpr int ITERS = 256;

__global__ void probe(long long* out) {
  __shared__ __align__(8) uint64_t bar[4];
  __shared__ uint32_t tmem_slot;
  __shared__ volatile int flag;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int i = 0; i < 4; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar[i])), "r"(1));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    flag = 0;
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 64;" ::"r"(smem_u32(&tmem_slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tbase = tmem_slot;
  const uint32_t b0 = smem_u32(&bar[0]), b1 = smem_u32(&bar[1]);
  if (warp == 0) {
    // complete phase 0 of bar[0] once so that parity-0 waits succeed immediately
    if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(b0) : "memory");
    __syncwarp();
    uint32_t ok = 0, sink = 0;
    { T0();   // (0) mbarrier.test_wait on a completed phase, result consumed at once
      for (int i = 0; i < ITERS; ++i) {
        asm volatile("{\n.reg .pred p;\nmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(ok) : "r"(b0), "r"(0) : "memory");
        sink += ok;
        asm volatile("" : "+r"(sink));
      }
      T1(0); }
    { T0();   // (1) mbarrier.try_wait on a completed phase
      for (int i = 0; i < ITERS; ++i) {
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(ok) : "r"(b0), "r"(0) : "memory");
        sink += ok;
        asm volatile("" : "+r"(sink));
      }
      T1(1); }
    { T0();   // (2) tcgen05.wait::st with nothing pending
      for (int i = 0; i < ITERS; ++i) asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      T1(2); }
    { T0();   // (3) two tcgen05.st.x8 + wait::st
      for (int i = 0; i < ITERS; ++i) {
        asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1};" ::"r"(tbase), "r"(sink) : "memory");
        asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1};" ::"r"(tbase + 8), "r"(sink) : "memory");
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      }
      T1(3); }
    { T0();   // (4) fence::before + fence::after
      for (int i = 0; i < ITERS; ++i) {
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      }
      T1(4); }
    { T0();   // (5) tcgen05.ld.x8 + wait::ld
      uint32_t r[8];
      for (int i = 0; i < ITERS; ++i) {
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(tbase) : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        sink += r[0];
      }
      T1(5); }
    { T0();   // (6) tcgen05.commit (nothing pending) -> own try_wait until the phase completes
      for (int i = 0; i < ITERS; ++i) {
        if (lane == 0) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(b1) : "memory");
        uint32_t done = 0;
        while (!done)
          asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(done) : "r"(b1), "r"(i & 1) : "memory");
      }
      T1(6); }
    { T0();   // (7) mbarrier.arrive (lane 0) -> own try_wait
      const uint32_t b2 = smem_u32(&bar[2]);
      for (int i = 0; i < ITERS; ++i) {
        if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(b2) : "memory");
        uint32_t done = 0;
        while (!done)
          asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(done) : "r"(b2), "r"(i & 1) : "memory");
      }
      T1(7); }
    if (sink == 0x7fffffff) out[31] = sink;
  }
  __syncthreads();
  // (8) ping-pong between warp 0 and warp 1 through two mbarriers (round trip / 2 = one hop)
  {
    __shared__ __align__(8) uint64_t pp[2];
    if (threadIdx.x == 0) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&pp[0])), "r"(1));
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&pp[1])), "r"(1));
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const uint32_t p0 = smem_u32(&pp[0]), p1 = smem_u32(&pp[1]);
    if (warp < 2) {
      T0();
      for (int i = 0; i < ITERS; ++i) {
        const uint32_t mine = warp == 0 ? p0 : p1, other = warp == 0 ? p1 : p0;
        if (warp == 0) {
          if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(other) : "memory");
          uint32_t done = 0;
          while (!done) asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(done) : "r"(mine), "r"(i & 1) : "memory");
        } else {
          uint32_t done = 0;
          while (!done) asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(done) : "r"(mine), "r"(i & 1) : "memory");
          if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(other) : "memory");
        }
      }
      if (warp == 0) T1(8);
    }
  }
  __syncthreads();
  // (9) ping-pong through two named barriers (bar.sync 1 / 2 with 64 threads)
  if (warp < 2) {
    T0();
    for (int i = 0; i < ITERS; ++i) {
      asm volatile("bar.sync 1, 64;" ::: "memory");
      asm volatile("bar.sync 2, 64;" ::: "memory");
    }
    if (warp == 0) T1(9);
  }
  __syncthreads();
  // (10) ping-pong through volatile shared-memory flags
  if (warp < 2) {
    T0();
    for (int i = 0; i < ITERS; ++i) {
      if (warp == 0) { if (lane == 0) flag = 2 * i + 1; while (flag != 2 * i + 2) {} }
      else { while (flag != 2 * i + 1) {} if (lane == 0) flag = 2 * i + 2; }
      __syncwarp();
    }
    if (warp == 0) T1(10);
  }
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 64;" ::"r"(tbase) : "memory");
}

int main() {
  long long* d; cudaMalloc(&d, 32 * 8); cudaMemset(d, 0, 32 * 8);
  probe<<<1, 128>>>(d);
  long long h[32];
  cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  const char* n[] = {"mbarrier.test_wait (completed phase), result used", "mbarrier.try_wait (completed phase), result used",
                     "tcgen05.wait::st, nothing pending", "2 x tcgen05.st.x8 + wait::st", "tcgen05.fence before + after",
                     "tcgen05.ld.x8 + wait::ld", "tcgen05.commit (idle pipe) -> try_wait success (same warp)",
                     "mbarrier.arrive -> try_wait success (same warp)", "mbarrier ping-pong between two warps (round trip)",
                     "named-barrier ping-pong (2 x bar.sync, 64 threads)", "shared-memory flag ping-pong (round trip)"};
  for (int i = 0; i <= 10; ++i) printf("%-62s %5lld cycles\n", n[i], h[i]);
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
