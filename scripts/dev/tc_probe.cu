// tcgen05 probe for the tensor-core triangle contraction (sm_100a, standalone: no torch).
//   nvcc -O2 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tc_probe tc_probe.cu && ./tc_probe
// Answers, on the real B200:
//   1. TMEM st/ld round trip with the 32x32b shape (lane = thread, column = register index)
//   2. kind::tf32 MMA with both operands in shared memory (K-major, no swizzle descriptors)
//   3. kind::tf32 MMA with A in TMEM (lane = M row, column = k), B in shared memory
//   4. how the tensor core narrows fp32 containers to tf32 (truncate vs round)
//   5. kind::f16 MMA with A in TMEM (two k per column)
//   6. cycles per MMA for M=128, several N, A in TMEM vs shared memory, tf32 vs f16
//   7. whether N=40 is accepted with M=128 (last: may fault)
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>
#include <cuda_fp16.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); fflush(stdout); exit(2); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ bool mbar_wait_bounded(uint64_t* bar, uint32_t parity) {
  for (int it = 0; it < (1 << 24); ++it) {
    uint32_t ok;
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    if (ok) return true;
  }
  return false;
}
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// K-major, no swizzle: 8-row x 16-byte core matrices; LBO = byte step between the two 16-byte
// K chunks of one MMA, SBO = byte step between 8-row groups.
__host__ __device__ inline uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;  // descriptor version (Blackwell)
  return d;
}
// kind: 0 = f16 inputs, 2 = tf32 inputs; fp32 accumulate, both operands K-major
__host__ __device__ inline uint32_t make_idesc(int fmt, int M, int N) {
  return (1u << 4) | ((uint32_t)fmt << 7) | ((uint32_t)fmt << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

__device__ __forceinline__ void mma_tf32_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
               "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}\n" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_tf32_ts(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
               "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n}\n" ::"r"(d), "r"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_f16_ts(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
               "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n}\n" ::"r"(d), "r"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}

__device__ __forceinline__ void tmem_st8(uint32_t addr, const uint32_t (&v)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(addr), "r"(v[0]), "r"(v[1]), "r"(v[2]),
               "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]) : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t addr, uint32_t (&v)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]) : "r"(addr) : "memory");
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

constexpr int MAXN = 256;
// shared-memory operand images (byte offsets inside the dynamic buffer)
constexpr int OFF_A = 0;                 // 128 rows: 16 groups x 256 B = 4 KB
constexpr int OFF_B = 4096;              // 256 rows: 32 groups x 256 B = 8 KB
constexpr int OFF_BAR = 4096 + 8192;     // mbarrier + tmem base
constexpr int SMEM_BYTES = OFF_BAR + 64;

// element (row, k) of a K-major 32-bit operand in the no-swizzle canonical layout
__host__ __device__ inline int canon32(int row, int k) { return (row >> 3) * 256 + (k >> 2) * 128 + (row & 7) * 16 + (k & 3) * 4; }
// same for 16-bit elements (K = 16 per MMA: two chunks of 8 elements)
__host__ __device__ inline int canon16(int row, int k) { return (row >> 3) * 256 + (k >> 3) * 128 + (row & 7) * 16 + (k & 7) * 2; }

struct Out {
  uint32_t roundtrip_bad;
  uint32_t timeout;
  float d_ss[128 * 48];
  float d_ts[128 * 48];
  float d_f16[128 * 48];
  float d_acc[128 * 48];   // accumulate=1 on top of d_ts
  long long cyc[40];
};

__global__ void __launch_bounds__(128, 1)
probe_kernel(const float* __restrict__ A, const float* __restrict__ B, Out* out, int try_n40) {
  extern __shared__ __align__(1024) unsigned char smem[];
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + OFF_BAR + 16);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    out->roundtrip_bad = 0;
    out->timeout = 0;
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(tmem_slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = *tmem_slot;
  const uint32_t lane_base = tbase + ((uint32_t)(warp * 32) << 16);
  uint32_t phase = 0;

  // ---- 1. round trip
  {
    uint32_t v[8], r[8];
    for (int c = 0; c < 8; ++c) v[c] = 0xA0000000u | (tid << 8) | c;
    tmem_st8(lane_base + 0, v);
    tmem_wait_st();
    tmem_ld8(lane_base + 0, r);
    tmem_wait_ld();
    int bad = 0;
    for (int c = 0; c < 8; ++c) bad += (r[c] != v[c]);
    if (bad) atomicAdd(&out->roundtrip_bad, bad);
  }

  // ---- operand images in shared memory (tf32 containers = fp32 bits)
  for (int i = tid; i < 128 * 8; i += 128) {
    const int m = i / 8, k = i % 8;
    *reinterpret_cast<float*>(smem + OFF_A + canon32(m, k)) = A[m * 8 + k];
  }
  for (int i = tid; i < MAXN * 8; i += 128) {
    const int n = i / 8, k = i % 8;
    *reinterpret_cast<float*>(smem + OFF_B + canon32(n, k)) = B[n * 8 + k];
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();

  const uint64_t adesc = make_desc(smem_u32(smem + OFF_A), 128, 256);
  const uint64_t bdesc = make_desc(smem_u32(smem + OFF_B), 128, 256);
  const uint32_t D0 = tbase + 64, D1 = tbase + 128, D2 = tbase + 192, ACOL = tbase + 256, A16COL = tbase + 272;

  auto read_d = [&](uint32_t dcol, float* dst) {
    for (int c0 = 0; c0 < 48; c0 += 8) {
      uint32_t r[8];
      tmem_ld8(dcol + ((uint32_t)(warp * 32) << 16) + c0, r);
      tmem_wait_ld();
      for (int c = 0; c < 8; ++c) dst[tid * 48 + c0 + c] = __uint_as_float(r[c]);
    }
  };
  auto wait_mma = [&]() {
    if (!mbar_wait_bounded(bar, phase)) { if (lane == 0) atomicAdd(&out->timeout, 1); }
    phase ^= 1;
    tc_fence_after();
  };

  // ---- 2. SS
  if (tid == 0) {
    mma_tf32_ss(D0, adesc, bdesc, make_idesc(2, 128, 48), 0);
    tc_commit(bar);
  }
  wait_mma();
  read_d(D0, out->d_ss);

  // ---- 3. TS: A row m (= this thread's TMEM lane) columns k = 0..7
  {
    uint32_t v[8];
    for (int k = 0; k < 8; ++k) v[k] = __float_as_uint(A[tid * 8 + k]);
    tmem_st8(ACOL + ((uint32_t)(warp * 32) << 16), v);
    tmem_wait_st();
  }
  tc_fence_before();
  __syncthreads();
  if (tid == 0) {
    tc_fence_after();
    mma_tf32_ts(D1, ACOL, bdesc, make_idesc(2, 128, 48), 0);
    tc_commit(bar);
  }
  wait_mma();
  read_d(D1, out->d_ts);
  // accumulate a second time on top (enable_input_d = 1): expect 2x
  tc_fence_before();
  __syncthreads();
  if (tid == 0) {
    tc_fence_after();
    mma_tf32_ts(D1, ACOL, bdesc, make_idesc(2, 128, 48), 1);
    tc_commit(bar);
  }
  wait_mma();
  read_d(D1, out->d_acc);

  // ---- 5. f16 TS: K = 16; A column c holds (k = 2c in the low half, k = 2c+1 in the high half)
  __syncthreads();
  for (int i = tid; i < MAXN * 16; i += 128) {
    const int n = i / 16, k = i % 16;
    *reinterpret_cast<__half*>(smem + OFF_B + canon16(n, k)) = __float2half_rn(B[n * 8 + (k & 7)] * (k < 8 ? 1.f : 0.5f));
  }
  {
    uint32_t v[8];
    for (int c = 0; c < 8; ++c) {
      const __half lo = __float2half_rn(A[tid * 8 + ((2 * c) & 7)] * ((2 * c) < 8 ? 1.f : 0.25f));
      const __half hi = __float2half_rn(A[tid * 8 + ((2 * c + 1) & 7)] * ((2 * c + 1) < 8 ? 1.f : 0.25f));
      v[c] = (uint32_t)__half_as_ushort(lo) | ((uint32_t)__half_as_ushort(hi) << 16);
    }
    tmem_st8(A16COL + ((uint32_t)(warp * 32) << 16), v);
    tmem_wait_st();
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  if (tid == 0) {
    tc_fence_after();
    mma_f16_ts(D2, A16COL, bdesc, make_idesc(0, 128, 48), 0);
    tc_commit(bar);
  }
  wait_mma();
  read_d(D2, out->d_f16);

  // ---- 6. timing: NI back-to-back MMAs, one accumulator; (variant, N) table
  const int NI = 4096;
  const int Ns[6] = {16, 32, 48, 64, 128, 256};
  int slot = 0;
  for (int variant = 0; variant < 4; ++variant) {     // 0 tf32 TS, 1 tf32 SS, 2 f16 TS, 3 tf32 TS alternating 2 accumulators
    for (int ni = 0; ni < 6; ++ni) {
      const int N = Ns[ni];
      tc_fence_before();
      __syncthreads();
      long long t0 = 0;
      if (tid == 0) {
        tc_fence_after();
        const uint32_t idt = make_idesc(2, 128, N), idh = make_idesc(0, 128, N);
        // accumulators from column 0 (N up to 256); A operands live at columns >= 256
        t0 = clock64();
        for (int i = 0; i < NI; ++i) {
          if (variant == 0) mma_tf32_ts(tbase, ACOL, bdesc, idt, 1);
          else if (variant == 1) mma_tf32_ss(tbase, adesc, bdesc, idt, 1);
          else if (variant == 2) mma_f16_ts(tbase, A16COL, bdesc, idh, 1);
          else mma_tf32_ts((i & 1) && N <= 128 ? tbase + 128 : tbase, ACOL, bdesc, idt, 1);
        }
        tc_commit(bar);
      }
      wait_mma();
      if (tid == 0) out->cyc[slot] = clock64() - t0;
      ++slot;
    }
  }

  // ---- 7. N = 40 with M = 128 (may be rejected by the hardware)
  if (try_n40) {
    __syncthreads();
    for (int i = tid; i < MAXN * 8; i += 128) {
      const int n = i / 8, k = i % 8;
      *reinterpret_cast<float*>(smem + OFF_B + canon32(n, k)) = B[n * 8 + k];
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tc_fence_before();
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
      mma_tf32_ts(D1, ACOL, bdesc, make_idesc(2, 128, 40), 0);
      tc_commit(bar);
    }
    wait_mma();
    read_d(D1, out->d_acc);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tbase) : "memory");
}

static float tf32_trunc(float x) { uint32_t u; memcpy(&u, &x, 4); u &= 0xFFFFE000u; memcpy(&x, &u, 4); return x; }
static float tf32_rn(float x) { uint32_t u; memcpy(&u, &x, 4); u += 0x00000FFFu + ((u >> 13) & 1u); u &= 0xFFFFE000u; memcpy(&x, &u, 4); return x; }

int main(int argc, char** argv) {
  const int try_n40 = argc > 1 ? atoi(argv[1]) : 0;
  std::vector<float> A(128 * 8), B(MAXN * 8);
  srand(12345);
  for (auto& v : A) v = (float)rand() / RAND_MAX * 2.f - 1.f;
  for (auto& v : B) v = (float)rand() / RAND_MAX * 2.f - 1.f;
  float *dA, *dB; Out* dO;
  CK(cudaMalloc(&dA, A.size() * 4)); CK(cudaMalloc(&dB, B.size() * 4)); CK(cudaMalloc(&dO, sizeof(Out)));
  CK(cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemset(dO, 0, sizeof(Out)));
  CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
  probe_kernel<<<1, 128, SMEM_BYTES>>>(dA, dB, dO, try_n40);
  CK(cudaGetLastError());
  CK(cudaDeviceSynchronize());
  Out* o = new Out;
  CK(cudaMemcpy(o, dO, sizeof(Out), cudaMemcpyDeviceToHost));
  printf("roundtrip mismatches: %u   mbarrier timeouts: %u\n", o->roundtrip_bad, o->timeout);
  auto cmp = [&](const char* name, const float* got, int nact, float (*narrow)(float), double scale) {
    double maxerr = 0, maxref = 0;
    for (int m = 0; m < 128; ++m)
      for (int n = 0; n < nact; ++n) {
        double ref = 0;
        for (int k = 0; k < 8; ++k) ref += (double)narrow(A[m * 8 + k]) * (double)narrow(B[n * 8 + k]);
        ref *= scale;
        maxerr = fmax(maxerr, fabs(ref - got[m * 48 + n]));
        maxref = fmax(maxref, fabs(ref));
      }
    printf("%-28s max|err| = %.3e (max|ref| %.3f)\n", name, maxerr, maxref);
  };
  cmp("SS tf32 vs truncated inputs", o->d_ss, 48, tf32_trunc, 1.0);
  cmp("SS tf32 vs RN-rounded inputs", o->d_ss, 48, tf32_rn, 1.0);
  cmp("TS tf32 vs truncated inputs", o->d_ts, 48, tf32_trunc, 1.0);
  cmp("TS tf32 vs RN-rounded inputs", o->d_ts, 48, tf32_rn, 1.0);
  if (!try_n40) cmp("TS accumulate x2 (trunc)", o->d_acc, 48, tf32_trunc, 2.0);
  else cmp("TS N=40 (trunc)", o->d_acc, 40, tf32_trunc, 1.0);
  {  // f16: k<8 plain, k>=8 scaled copies: sum_k a_k b_k (1 + 0.25*0.5)
    double maxerr = 0;
    for (int m = 0; m < 128; ++m)
      for (int n = 0; n < 48; ++n) {
        double ref = 0;
        for (int k = 0; k < 16; ++k) {
          const float a = __half2float(__float2half_rn(A[m * 8 + (k & 7)] * (k < 8 ? 1.f : 0.25f)));
          const float b = __half2float(__float2half_rn(B[n * 8 + (k & 7)] * (k < 8 ? 1.f : 0.5f)));
          ref += (double)a * b;
        }
        maxerr = fmax(maxerr, fabs(ref - o->d_f16[m * 48 + n]));
      }
    printf("%-28s max|err| = %.3e\n", "TS f16 (K=16)", maxerr);
  }
  const char* vn[4] = {"tf32 TS", "tf32 SS", "f16 TS", "tf32 TS 2 accumulators"};
  const int Ns[6] = {16, 32, 48, 64, 128, 256};
  for (int v = 0; v < 4; ++v) {
    printf("%-24s cycles/MMA (M=128):", vn[v]);
    for (int i = 0; i < 6; ++i) printf("  N=%d: %.1f", Ns[i], (double)o->cyc[v * 6 + i] / 4096.0);
    printf("\n");
  }
  fflush(stdout);
  return 0;
}
