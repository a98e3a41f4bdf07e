// Tensor-core triangle contraction (sm_100a, tcgen05 + TMEM), included by contract.cu.
//
//   D[(a,b), c] = sum_x (F_a(x) F_b(x)) * F_c(x)        for every row pair (a,b) and every row c
//
// i.e. the reference's T separate np.sum(f[a]*f[b]*f[c]) loops (bskit/main.py:1871-1879) recast
// as one skinny GEMM over the grid axis: M = row pairs (128 per "unit"), N = rows (<= 40),
// K = cells.  Measured on B200 (profiles/r1_tcgen05_probe.txt): with M = 128 a tcgen05.mma costs
// 128*N/256 cycles only when the A operand comes from TMEM (44 instead of 24 cycles at N = 48
// from shared memory), so the pair products -- which have to be generated on the CUDA cores
// anyway -- are written straight into TMEM and never touch shared memory.
//
// Numerics.  3xTF32: P = F_a*F_b (fp32, RN) is split P = P_hi + P_lo, F_c = C_hi + C_lo with
// 11-bit pieces, and D += P_lo*C_hi + P_hi*C_lo + P_hi*C_hi.  The tensor core truncates its fp32
// accumulator (round toward zero, profiles/r1_tensor_core_rounding_probe.txt), which would bias a
// long heavily-cancelling sum, so an accumulator lives in TMEM for ONE 32-cell chunk only
// (12 MMAs); it is then drained into fp32 registers with round-to-nearest adds, and those are
// flushed into float64 partials every few hundred chunks.
//
// Roles (384 threads, one CTA per SM, persistent over tiles of 128 cells):
//   warps 0-3 / 4-7  two generator+drain teams; lane v of every warp keeps row v of the current
//                    chunk in registers, so a unit costs one 128-byte shared-memory read per lane
//   warp 8           TMA producer: raw [row][cell] tiles, cp.async.bulk + mbarrier
//   warp 9           MMA issuer (one elected lane)
//   warps 10-11      convert the raw tile to the B operand images (hi / lo, K-major core
//                    matrices: 8 rows x 16 bytes)
#pragma once

namespace bsk {
namespace tc {

constexpr int CH = 32;            // cells per chunk = K extent of one unit
constexpr int TL = 128;           // cells per tile
constexpr int NCH = TL / CH;
constexpr int MAXR = 40;          // rows (N <= 40)
constexpr int UPT = 4;            // units per team (2 teams -> at most 8 units = 1024 pair rows)
// D columns a unit may need, by position j in its team (units are dealt to the teams in order of
// decreasing width): a pair (a <= b) only meets rows c >= b, so most units need few columns
__host__ __device__ constexpr int cap(int j) { return j == 0 ? 40 : j == 1 ? 32 : j == 2 ? 24 : 8; }
__host__ __device__ constexpr int capoff(int j) { return j == 0 ? 0 : j == 1 ? 40 : j == 2 ? 72 : 96; }
constexpr int CAPSUM = 104;
constexpr int NTHREADS = 384;
constexpr bool USE_OWN = false;   // keep lane v's row v in registers (saves shared-memory reads, costs 32 registers)
constexpr int RAW_STRIDE = TL * 4 + 16;            // bytes; +16 keeps lanes on distinct banks
constexpr int RAW_BYTES = ((MAXR + 1) * RAW_STRIDE + 127) / 128 * 128;  // + one all-zero row for idle lanes
constexpr int BIMG_BYTES = (TL / 4) * (MAXR / 8) * 128;   // one hi or lo image of a tile
constexpr int OFF_BAR = 0;
constexpr int OFF_RAW = 256;
constexpr int OFF_BIMG = OFF_RAW + 2 * RAW_BYTES;  // [buf][hi|lo]
constexpr int SMEM_BYTES = OFF_BIMG + 4 * BIMG_BYTES;
static_assert(OFF_BIMG % 128 == 0, "operand images must be 128-byte aligned");

// barrier indices
enum { RAW_FULL = 0, RAW_EMPTY = 2, B_FULL = 4, B_EMPTY = 6, A_FULL = 8, A_EMPTY = 12, D_FULL = 16, D_EMPTY = 20, NBAR = 24 };

// bounded wait: a protocol bug traps instead of hanging the GPU
__device__ __forceinline__ void tc_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok = 0;
  for (uint32_t it = 0; it < (1u << 27); ++it) {
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    if (ok) return;
  }
  __trap();
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ uint32_t elect_one() {
  uint32_t pred;
  asm volatile("{\n.reg .b32 rx;\n.reg .pred px;\nelect.sync rx|px, 0xffffffff;\nselp.u32 %0, 1, 0, px;\n}\n" : "=r"(pred));
  return pred;
}
// K-major, no swizzle: LBO = byte step between the two 16-byte K chunks of one MMA, SBO = byte
// step between 8-row groups (include/..., cute mma_sm100_desc.hpp SmemDescriptor bit layout)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) |
         ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) | ((uint64_t)1 << 46);
}
// tf32 x tf32 -> fp32, both operands K-major, M = 128
__host__ __device__ constexpr uint32_t make_idesc_tf32(int n) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}
__device__ __forceinline__ void mma_tf32_ts(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
               "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n}\n" ::"r"(d), "r"(a), "l"(b), "r"(idesc), "r"(acc)
               : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t addr, const uint32_t (&v)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(addr), "r"(v[0]), "r"(v[1]),
               "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
               : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t addr, float (&v)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7])
               : "r"(addr)
               : "memory");
}
__device__ __forceinline__ float4 lds128(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ void sts128(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

struct Params {
  const float* const* rowptr;
  int nrows;        // R <= 40
  int ncols;        // N = R rounded up to 8
  int64_t ntiles;   // ncells / TL
  int nu0, nu1;     // units of team 0 / team 1
  int ucol0[2 * UPT];        // [team * UPT + j]: first D column the unit needs (multiple of 8)
  int uncol[2 * UPT];        //                   number of columns (multiple of 8, <= CAP[j])
  const uint32_t* slot_tab;  // [team * UPT + j][4][32]: ra | rb << 8 | resident << 16  (row R = zero row)
  double* partial;           // [cta][CAPSUM][256]
  int64_t partial_stride;
  int flush_chunks;
};

__device__ __forceinline__ void tmem_st16(uint32_t addr, const uint32_t (&v)[16]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(addr),
               "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
               "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
               : "memory");
}
// issue a TMEM load of 8 columns as four float2 (no wait)
__device__ __forceinline__ void tmem_ld8v(uint32_t addr, float2 (&v)[4]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=f"(v[0].x), "=f"(v[0].y), "=f"(v[1].x), "=f"(v[1].y), "=f"(v[2].x), "=f"(v[2].y), "=f"(v[3].x), "=f"(v[3].y)
               : "r"(addr)
               : "memory");
}
// the loaded values may be used only after tcgen05.wait::ld: route them through an (empty)
// volatile asm placed after the wait so the compiler cannot hoist their uses above it
__device__ __forceinline__ void pin8(float2 (&v)[4]) {
  asm volatile("" : "+f"(v[0].x), "+f"(v[0].y), "+f"(v[1].x), "+f"(v[1].y), "+f"(v[2].x), "+f"(v[2].y), "+f"(v[3].x), "+f"(v[3].y));
}

// Generator + drain team member.  team is 0 or 1, q = warp % 4 is the TMEM lane quarter.
__device__ __forceinline__ void team_loop(const Params& p, const unsigned char* smem, uint64_t* bars, uint32_t tbase,
                                          int team, int q, int lane) {
  const int my_nu = team ? p.nu1 : p.nu0;
  const int R = p.nrows;
  uint32_t rb_off[UPT], ra_off[UPT];
  bool resident[UPT];
#pragma unroll
  for (int j = 0; j < UPT; ++j) {
    uint32_t e = (uint32_t)R | ((uint32_t)R << 8);
    if (j < my_nu) e = p.slot_tab[((team * UPT + j) * 4 + q) * 32 + lane];
    ra_off[j] = (e & 0xFFu) * RAW_STRIDE;
    rb_off[j] = ((e >> 8) & 0xFFu) * RAW_STRIDE;
    resident[j] = USE_OWN && ((e >> 16) & 1u);   // warp-uniform by construction
  }
  const uint32_t own_off = (uint32_t)(lane < R ? lane : R) * RAW_STRIDE;
  const uint32_t lane_sel = (uint32_t)(q * 32) << 16;
  float2 acc0[cap(0) / 2], acc1[cap(1) / 2], acc2_[cap(2) / 2], acc3[cap(3) / 2];
#pragma unroll
  for (int c = 0; c < cap(0) / 2; ++c) acc0[c] = make_float2(0.f, 0.f);
#pragma unroll
  for (int c = 0; c < cap(1) / 2; ++c) acc1[c] = make_float2(0.f, 0.f);
#pragma unroll
  for (int c = 0; c < cap(2) / 2; ++c) acc2_[c] = make_float2(0.f, 0.f);
#pragma unroll
  for (int c = 0; c < cap(3) / 2; ++c) acc3[c] = make_float2(0.f, 0.f);
  int ncol[UPT];
#pragma unroll
  for (int j = 0; j < UPT; ++j) ncol[j] = p.uncol[team * UPT + j];

  uint32_t n_gen = 0;     // units generated by this team so far
  uint32_t n_drain = 0;   // units drained so far
  int since_flush = 0;
  double* my_partial = p.partial + (int64_t)blockIdx.x * p.partial_stride + team * 128 + q * 32 + lane;

  auto drain_into = [&](auto jj, auto& acc) {
    constexpr int J = decltype(jj)::value;
    const uint32_t dbuf = (uint32_t)team * 2u + (n_drain & 1u);
    const uint32_t d_tmem = tbase + lane_sel + dbuf * 64u;
    tc_wait(&bars[D_FULL + dbuf], (n_drain >> 1) & 1u);
    tc_fence_after();
    float2 v[cap(J) / 8][4];
#pragma unroll
    for (int g = 0; g < cap(J) / 8; ++g)
      if (g * 8 < ncol[J]) tmem_ld8v(d_tmem + g * 8, v[g]);
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int g = 0; g < cap(J) / 8; ++g)
      if (g * 8 < ncol[J]) pin8(v[g]);
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(&bars[D_EMPTY + dbuf]);
#pragma unroll
    for (int g = 0; g < cap(J) / 8; ++g)
      if (g * 8 < ncol[J]) {
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[g * 4 + c] = __fadd2_rn(acc[g * 4 + c], v[g][c]);
      }
    ++n_drain;
  };
  auto drain_dyn = [&](int j) {   // statically indexed accumulators behind a warp-uniform switch
    if (j == 0) drain_into(std::integral_constant<int, 0>{}, acc0);
    else if (j == 1) drain_into(std::integral_constant<int, 1>{}, acc1);
    else if (j == 2) drain_into(std::integral_constant<int, 2>{}, acc2_);
    else drain_into(std::integral_constant<int, 3>{}, acc3);
  };
  auto flush = [&]() {
#pragma unroll
    for (int c = 0; c < cap(0) / 2; ++c)
      if (2 * c < ncol[0]) {
        atomicAdd(my_partial + (int64_t)(capoff(0) + 2 * c) * 256, (double)acc0[c].x);
        atomicAdd(my_partial + (int64_t)(capoff(0) + 2 * c + 1) * 256, (double)acc0[c].y);
        acc0[c] = make_float2(0.f, 0.f);
      }
#pragma unroll
    for (int c = 0; c < cap(1) / 2; ++c)
      if (2 * c < ncol[1]) {
        atomicAdd(my_partial + (int64_t)(capoff(1) + 2 * c) * 256, (double)acc1[c].x);
        atomicAdd(my_partial + (int64_t)(capoff(1) + 2 * c + 1) * 256, (double)acc1[c].y);
        acc1[c] = make_float2(0.f, 0.f);
      }
#pragma unroll
    for (int c = 0; c < cap(2) / 2; ++c)
      if (2 * c < ncol[2]) {
        atomicAdd(my_partial + (int64_t)(capoff(2) + 2 * c) * 256, (double)acc2_[c].x);
        atomicAdd(my_partial + (int64_t)(capoff(2) + 2 * c + 1) * 256, (double)acc2_[c].y);
        acc2_[c] = make_float2(0.f, 0.f);
      }
#pragma unroll
    for (int c = 0; c < cap(3) / 2; ++c)
      if (2 * c < ncol[3]) {
        atomicAdd(my_partial + (int64_t)(capoff(3) + 2 * c) * 256, (double)acc3[c].x);
        atomicAdd(my_partial + (int64_t)(capoff(3) + 2 * c + 1) * 256, (double)acc3[c].y);
        acc3[c] = make_float2(0.f, 0.f);
      }
  };

  int it = 0;
  for (int64_t tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x, ++it) {
    const int buf = it & 1;
    tc_wait(&bars[RAW_FULL + buf], (uint32_t)(it >> 1) & 1u);
    const unsigned char* raw = smem + OFF_RAW + buf * RAW_BYTES;
#pragma unroll 1
    for (int c = 0; c < NCH; ++c) {
      const unsigned char* cbase = raw + c * (CH * 4);
      float2 own[CH / 2];
      if constexpr (USE_OWN) {
#pragma unroll
        for (int v4 = 0; v4 < CH / 4; ++v4) {
          const float4 t = *reinterpret_cast<const float4*>(cbase + own_off + v4 * 16);
          own[v4 * 2] = make_float2(t.x, t.y);
          own[v4 * 2 + 1] = make_float2(t.z, t.w);
        }
      }
#pragma unroll
      for (int j = 0; j < UPT; ++j) {
        if (j < my_nu) {
          const uint32_t abuf = (uint32_t)team * 2u + (n_gen & 1u);
          const uint32_t a_tmem = tbase + lane_sel + 256u + abuf * 64u;
          bool waited = false;
#pragma unroll
          for (int h = 0; h < 4; ++h) {
            // 8 cells: partner row from shared memory, own row from registers (resident slots)
            float2 pb[4], pa[4];
#pragma unroll
            for (int v4 = 0; v4 < 2; ++v4) {
              const float4 t = *reinterpret_cast<const float4*>(cbase + rb_off[j] + h * 32 + v4 * 16);
              pb[v4 * 2] = make_float2(t.x, t.y);
              pb[v4 * 2 + 1] = make_float2(t.z, t.w);
            }
            if (USE_OWN && resident[j]) {
#pragma unroll
              for (int i = 0; i < 4; ++i) pa[i] = own[h * 4 + i];
            } else {
#pragma unroll
              for (int v4 = 0; v4 < 2; ++v4) {
                const float4 t = *reinterpret_cast<const float4*>(cbase + ra_off[j] + h * 32 + v4 * 16);
                pa[v4 * 2] = make_float2(t.x, t.y);
                pa[v4 * 2 + 1] = make_float2(t.z, t.w);
              }
            }
            uint32_t hi[8], lo[8];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const float2 pr = __fmul2_rn(pa[i], pb[i]);
              const float2 ph = make_float2(__uint_as_float(__float_as_uint(pr.x) & 0xFFFFE000u),
                                            __uint_as_float(__float_as_uint(pr.y) & 0xFFFFE000u));
              const float2 pl = __ffma2_rn(ph, make_float2(-1.f, -1.f), pr);   // exact
              hi[2 * i] = __float_as_uint(ph.x); hi[2 * i + 1] = __float_as_uint(ph.y);
              lo[2 * i] = __float_as_uint(pl.x); lo[2 * i + 1] = __float_as_uint(pl.y);
            }
            if (!waited) {
              tc_wait(&bars[A_EMPTY + abuf], ((n_gen >> 1) & 1u) ^ 1u);
              tc_fence_after();
              waited = true;
            }
            tmem_st8(a_tmem + h * 8, hi);
            tmem_st8(a_tmem + 32 + h * 8, lo);
          }
          asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&bars[A_FULL + abuf]);
          ++n_gen;
          // drain the unit generated before this one (its MMAs overlap this generation)
          if (n_gen > 1) drain_dyn(j > 0 ? j - 1 : my_nu - 1);
        }
      }
      if (++since_flush == p.flush_chunks) {
        flush();
        since_flush = 0;
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&bars[RAW_EMPTY + buf]);
  }
  if (n_gen > 0) drain_dyn(my_nu - 1);
  flush();
}

__global__ void __launch_bounds__(NTHREADS, 1) tc_contract_kernel(const __grid_constant__ Params p) {
  extern __shared__ __align__(1024) unsigned char smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + OFF_BAR + NBAR * 8);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int R = p.nrows, N = p.ncols;

  if (tid == 0) {
    for (int b = 0; b < 2; ++b) {
      mbar_init(&bars[RAW_FULL + b], 1);
      mbar_init(&bars[RAW_EMPTY + b], 10);
      mbar_init(&bars[B_FULL + b], 2);
      mbar_init(&bars[B_EMPTY + b], 1);
    }
    for (int b = 0; b < 4; ++b) {
      mbar_init(&bars[D_FULL + b], 1);
      mbar_init(&bars[D_EMPTY + b], 4);
      mbar_init(&bars[A_FULL + b], 4);
      mbar_init(&bars[A_EMPTY + b], 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  // zero row of both raw buffers and the whole operand-image area (rows R..N-1 stay zero)
  for (int i = tid; i < RAW_STRIDE / 4; i += NTHREADS) {
    reinterpret_cast<uint32_t*>(smem + OFF_RAW + R * RAW_STRIDE)[i] = 0u;
    reinterpret_cast<uint32_t*>(smem + OFF_RAW + RAW_BYTES + R * RAW_STRIDE)[i] = 0u;
  }
  for (int i = tid; i < 4 * BIMG_BYTES / 4; i += NTHREADS) reinterpret_cast<uint32_t*>(smem + OFF_BIMG)[i] = 0u;
  if (warp == 9) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(tmem_slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = *tmem_slot;

  if (warp < 8) {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 232;");
    team_loop(p, smem, bars, tbase, warp >> 2, warp & 3, lane);
  } else {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
    if (warp == 8) {
      // ---- TMA producer
      int it = 0;
      for (int64_t tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x, ++it) {
        const int buf = it & 1;
        tc_wait(&bars[RAW_EMPTY + buf], ((uint32_t)(it >> 1) & 1u) ^ 1u);
        if (lane == 0) mbar_expect_tx(&bars[RAW_FULL + buf], (uint32_t)(R * TL * 4));
        __syncwarp();
        unsigned char* dst = smem + OFF_RAW + buf * RAW_BYTES;
        for (int r = lane; r < R; r += 32)
          bulk_g2s(dst + r * RAW_STRIDE, p.rowptr[r] + tile * TL, TL * 4, &bars[RAW_FULL + buf]);
      }
    } else if (warp == 9) {
      // ---- MMA issuer
      const uint32_t leader = elect_one();
      const uint32_t lbo = (uint32_t)(N / 8) * 128u;     // next 4-cell group of the image
      uint32_t n_unit[2] = {0u, 0u};
      int it = 0;
      for (int64_t tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x, ++it) {
        const int buf = it & 1;
        tc_wait(&bars[B_FULL + buf], (uint32_t)(it >> 1) & 1u);
        const uint32_t img_hi = smem_u32(smem + OFF_BIMG + (buf * 2 + 0) * BIMG_BYTES);
        const uint32_t img_lo = smem_u32(smem + OFF_BIMG + (buf * 2 + 1) * BIMG_BYTES);
        for (int c = 0; c < NCH; ++c) {
          for (int j = 0; j < UPT; ++j) {
            for (int team = 0; team < 2; ++team) {
              if (j >= (team ? p.nu1 : p.nu0)) continue;
              const uint32_t n = n_unit[team]++;
              const uint32_t abuf = (uint32_t)team * 2u + (n & 1u);
              tc_wait(&bars[A_FULL + abuf], (n >> 1) & 1u);
              tc_wait(&bars[D_EMPTY + abuf], ((n >> 1) & 1u) ^ 1u);
              tc_fence_after();
              if (leader) {
                const uint32_t idesc = make_idesc_tf32(p.uncol[team * UPT + j]);
                const uint32_t coff = (uint32_t)(p.ucol0[team * UPT + j] / 8) * 128u;   // first 8-row group
                const uint32_t d = tbase + abuf * 64u;          // D and A buffers share the index
                const uint32_t a = tbase + 256u + abuf * 64u;
#pragma unroll
                for (int ks = 0; ks < CH / 8; ++ks) {
                  const uint32_t goff = (uint32_t)(c * (CH / 4) + ks * 2) * lbo + coff;
                  const uint64_t bh = make_desc(img_hi + goff, lbo, 128u);
                  const uint64_t bl = make_desc(img_lo + goff, lbo, 128u);
                  mma_tf32_ts(d, a + 32 + ks * 8, bh, idesc, ks > 0);   // P_lo * C_hi
                  mma_tf32_ts(d, a + ks * 8, bl, idesc, 1);             // P_hi * C_lo
                  mma_tf32_ts(d, a + ks * 8, bh, idesc, 1);             // P_hi * C_hi
                }
                tc_commit(&bars[A_EMPTY + abuf]);
                tc_commit(&bars[D_FULL + abuf]);
              }
              __syncwarp();
            }
          }
        }
        if (leader) tc_commit(&bars[B_EMPTY + buf]);
        __syncwarp();
      }
    } else {
      // ---- operand-image converters (64 threads): raw fp32 -> tf32 hi (round to nearest) + lo
      const int ct = tid - 320;
      const uint32_t ngrp = (uint32_t)(N / 8) * 128u;
      int it = 0;
      for (int64_t tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x, ++it) {
        const int buf = it & 1;
        tc_wait(&bars[RAW_FULL + buf], (uint32_t)(it >> 1) & 1u);
        tc_wait(&bars[B_EMPTY + buf], ((uint32_t)(it >> 1) & 1u) ^ 1u);
        const uint32_t raw = smem_u32(smem + OFF_RAW + buf * RAW_BYTES);
        const uint32_t img_hi = smem_u32(smem + OFF_BIMG + (buf * 2 + 0) * BIMG_BYTES);
        const uint32_t img_lo = smem_u32(smem + OFF_BIMG + (buf * 2 + 1) * BIMG_BYTES);
        for (int i = ct; i < R * (TL / 4); i += 64) {
          const int g = i / R, l = i - g * R;          // consecutive threads: consecutive rows
          const float4 v = lds128(raw + (uint32_t)l * RAW_STRIDE + (uint32_t)g * 16u);
          const float x[4] = {v.x, v.y, v.z, v.w};
          uint32_t h[4], lo[4];
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            h[k] = (__float_as_uint(x[k]) + 0x1000u) & 0xFFFFE000u;
            lo[k] = __float_as_uint(x[k] - __uint_as_float(h[k]));
          }
          const uint32_t o = (uint32_t)g * ngrp + (uint32_t)(l >> 3) * 128u + (uint32_t)(l & 7) * 16u;
          sts128(img_hi + o, h[0], h[1], h[2], h[3]);
          sts128(img_lo + o, lo[0], lo[1], lo[2], lo[3]);
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(&bars[B_FULL + buf]);
          mbar_arrive(&bars[RAW_EMPTY + buf]);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 9) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tbase) : "memory");
}

}  // namespace tc
}  // namespace bsk
