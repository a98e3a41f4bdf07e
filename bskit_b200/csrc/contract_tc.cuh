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
// 11-bit pieces (hi parts rounded to nearest), and D += P_lo*C_hi + P_hi*C_lo + P_hi*C_hi.  The
// tensor core truncates its fp32 accumulator (round toward zero,
// profiles/r1_tensor_core_rounding_probe.txt), which biases a long heavily-cancelling sum
// (~3.5e-7 per 32 cells accumulated), so an accumulator lives in TMEM for one window of WIN
// chunks only; it is then drained into fp32 registers with round-to-nearest adds, and those are
// flushed into float64 partials every few hundred chunks.  TMEM reads run at 64 B/cycle/SM, so
// the window is a speed/bias trade (WIN = 1 / 2 / 4: 70.7 / 52.3 / 47.6 ms, 3.7e-7 / 6.0e-7 /
// 9.2e-7 of max|sum| at 512^3, S = 40; FP32-pipe kernel: 71 ms, 5e-8).
//
// Roles (640 threads, one CTA per SM, persistent over tiles of 128 cells; setmaxnreg 72/144/40):
//   2 x 4 warps      generator teams, pure producers: warp q of a team owns TMEM lanes
//                    [32q, 32q+32) and writes the hi / lo pair products of one unit (128 pair rows x
//                    32 cells) into one of the team's two A buffers
//   2 x 4 warps      drain warpgroups, one per team: own the second-level accumulators
//   1 warp           TMA producer: raw [row][cell] tiles, cp.async.bulk + mbarrier
//   1 warp           MMA issuer (one elected lane): 12 MMAs per unit into the unit's own accumulator
//   2 warps          convert the raw tile to the B operand images (hi / lo, K-major core
//                    matrices: 8 rows x 16 bytes)
#pragma once

namespace bsk {
namespace tc {

constexpr int CH = 32;            // cells per chunk = K extent of one unit
constexpr int TL = 128;           // cells per tile
constexpr int NCH = TL / CH;
constexpr int WIN = 4;            // chunks accumulated in TMEM before a drain (divides NCH): the
                                  // accumulator truncates, so the window bounds the bias (~3.5e-7 per chunk)
static_assert(NCH % WIN == 0, "window");
constexpr int MAXR = 40;          // rows (N <= 40)
constexpr int NTEAMS = 2;
constexpr int UPT = 4;            // units per team (NTEAMS * UPT * 128 pair rows at most)
constexpr int NTEAMTHREADS = NTEAMS * 128;
constexpr int NTHREADS = 2 * NTEAMTHREADS + 128;    // generator teams + one drain warpgroup per team + 4 auxiliary warps
#ifndef BSK_TC_PROF
#define BSK_TC_PROF 0
#endif
constexpr bool PROF = BSK_TC_PROF;   // cycle counters per phase (block 0), see Params::prof
// 640 threads are launched with 96 registers each; setmaxnreg then moves registers from the
// generator and auxiliary warpgroups to the drain warpgroups, which hold the second-level
// accumulators: 256*72 + 256*144 + 128*40 <= 640*96
constexpr int REGS_TEAM = 72, REGS_DRAIN = 144, REGS_AUX = 40;
constexpr bool USE_OWN = false;   // lane v keeps row v of the chunk in registers across its team's units
static_assert(NTEAMTHREADS * (REGS_TEAM + REGS_DRAIN) + 128 * REGS_AUX <= NTHREADS * (65536 / NTHREADS / 8 * 8), "register split");
static_assert(NTEAMTHREADS * (96 - REGS_TEAM) + 128 * (96 - REGS_AUX) >= NTEAMTHREADS * (REGS_DRAIN - 96), "setmaxnreg pool");
// D columns a unit may need, by its position j in the team (units are sorted by decreasing width
// and dealt round-robin to the teams): a pair (a <= b) only meets rows c >= b, so most units need
// few columns
__host__ __device__ constexpr int cap(int j) { return j == 0 ? 40 : j == 1 ? 32 : j == 2 ? 24 : 8; }
__host__ __device__ constexpr int capoff(int j) { return j == 0 ? 0 : j == 1 ? 40 : j == 2 ? 72 : 96; }
constexpr int CAPSUM = 104;       // per team
// TMEM columns: one accumulator per unit (cap(j) columns) and two A buffers (hi 32 + lo 32
// columns) per team
constexpr int TM_D = 0, TM_A = 256;
static_assert(NTEAMS * CAPSUM <= TM_A && TM_A + NTEAMS * 2 * 64 <= 512, "TMEM budget");

constexpr int RAW_STRIDE = TL * 4 + 16;            // bytes; +16 keeps lanes on distinct banks
constexpr int RAW_BYTES = ((MAXR + 1) * RAW_STRIDE + 127) / 128 * 128;  // + one all-zero row for idle lanes
constexpr int BIMG_BYTES = (TL / 4) * (MAXR / 8) * 128;   // one hi or lo image of a tile
constexpr int OFF_BAR = 0;
constexpr int OFF_RAW = 512;
constexpr int OFF_BIMG = OFF_RAW + 2 * RAW_BYTES;  // [buf][hi|lo]
constexpr int SMEM_BYTES = OFF_BIMG + 4 * BIMG_BYTES;
static_assert(OFF_BIMG % 128 == 0, "operand images must be 128-byte aligned");

// barrier indices
enum {
  RAW_FULL = 0, RAW_EMPTY = 2, B_FULL = 4, B_EMPTY = 6,
  A_FULL = 8, A_EMPTY = A_FULL + 2 * NTEAMS, D_FULL = A_EMPTY + 2 * NTEAMS, D_EMPTY = D_FULL + NTEAMS * UPT,
  NBAR = D_EMPTY + NTEAMS * UPT
};
static_assert(NBAR * 8 + 8 <= OFF_RAW, "barrier area");

// bounded wait: a protocol bug traps instead of hanging the GPU
__device__ __forceinline__ void tc_wait(uint32_t bar_addr, uint32_t parity) {
  uint32_t ok = 0;
#pragma unroll 1
  for (uint32_t it = 0; it < (1u << 26); ++it) {
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                 : "=r"(ok) : "r"(bar_addr), "r"(parity) : "memory");
    if (ok) return;
  }
  __trap();
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar_addr) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar_addr) : "memory");
}
__device__ __forceinline__ void tc_commit(uint32_t bar_addr) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar_addr) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ uint32_t elect_one() {
  uint32_t pred;
  asm volatile("{\n.reg .b32 rx;\n.reg .pred px;\nelect.sync rx|px, 0xffffffff;\nselp.u32 %0, 1, 0, px;\n}\n" : "=r"(pred));
  return pred;
}
// K-major, no swizzle: LBO = byte step between the two 16-byte K chunks of one MMA, SBO = byte
// step between 8-row groups (bit layout as in CUTLASS cute/arch/mma_sm100_desc.hpp)
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
template <int ACC>
__device__ __forceinline__ void mma_tf32_ts2(uint32_t d, uint32_t a, uint32_t desc_lo, uint32_t desc_hi, uint32_t idesc) {
  asm volatile("{\n.reg .b64 bd;\nmov.b64 bd, {%2, %3};\n"
               "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], bd, %4, %5;\n}\n" ::"r"(d), "r"(a), "r"(desc_lo), "r"(desc_hi),
               "r"(idesc), "n"(ACC)
               : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t addr, const uint32_t (&v)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(addr), "r"(v[0]), "r"(v[1]),
               "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
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
__device__ __forceinline__ void sts128(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
// not volatile: the compiler may hoist and batch these loads
__device__ __forceinline__ float4 lds128(uint32_t addr) {
  float4 v;
  asm("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr) : "memory");
  return v;
}

struct Params {
  const float* const* rowptr;
  int nrows;        // R <= 40
  int ncols;        // N = R rounded up to 8
  int64_t ntiles;   // ncells / TL
  int nu[NTEAMS];             // units per team
  int ucol0[NTEAMS * UPT];    // [team * UPT + j]: first D column the unit needs (multiple of 8)
  int uncol[NTEAMS * UPT];    //                   number of columns (multiple of 8, <= cap(j))
  const uint32_t* slot_tab;   // [team * UPT + j][4][32]: ra | rb << 8  (row index R = zero row)
  double* partial;            // [cta][team][CAPSUM][128]
  int64_t partial_stride;
  int flush_chunks;
  long long* prof;            // PROF only: [team warp q=0: 8 counters per team][mma: 8 counters]
};

// Generator team member; q = warp % 4 is the TMEM lane quarter.  Pure producer: waits for a
// free A buffer, writes the 128 x 32 pair products (hi, lo) of one unit, signals the MMA issuer.
__device__ __forceinline__ void team_loop(const Params& p, const unsigned char* smem, uint32_t bars, uint32_t tbase,
                                          int team, int q, int lane) {
  const int my_nu = p.nu[team];
  const int R = p.nrows;
  uint32_t rb_off[UPT], ra_off[UPT];
  bool resident[UPT];   // warp-uniform: every lane's first row is row `lane` (or the lane is idle)
#pragma unroll
  for (int j = 0; j < UPT; ++j) {
    uint32_t e = (uint32_t)R | ((uint32_t)R << 8) | (1u << 16);
    if (j < my_nu) e = p.slot_tab[((team * UPT + j) * 4 + q) * 32 + lane];
    ra_off[j] = (e & 0xFFu) * RAW_STRIDE;
    rb_off[j] = ((e >> 8) & 0xFFu) * RAW_STRIDE;
    resident[j] = USE_OWN && __all_sync(0xffffffffu, (e >> 16) & 1u);
  }
  const uint32_t own_off = (uint32_t)(lane < R ? lane : R) * RAW_STRIDE;
  const uint32_t lane_sel = (uint32_t)(q * 32) << 16;
  const uint32_t a_tmem0 = tbase + lane_sel + TM_A + (uint32_t)team * 128u;
  const uint32_t bar_afull = bars + (A_FULL + team * 2) * 8, bar_aempty = bars + (A_EMPTY + team * 2) * 8;
  uint32_t n_gen = 0;     // units generated by this team so far
  long long t_raw = 0, t_gen = 0, t_aempty = 0, t_st = 0, t_mark = 0;
  auto tick = [&](long long& acc_t) {
    if constexpr (PROF) { const long long now = clock64(); acc_t += now - t_mark; t_mark = now; }
  };
  if constexpr (PROF) t_mark = clock64();

  int it = 0;
  for (int64_t tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x, ++it) {
    const int buf = it & 1;
    tc_wait(bars + (RAW_FULL + buf) * 8, (uint32_t)(it >> 1) & 1u);
    tick(t_raw);
    const uint32_t raw = smem_u32(smem + OFF_RAW + buf * RAW_BYTES);
#pragma unroll 1
    for (int c = 0; c < NCH; ++c) {
      const uint32_t cbase = raw + c * (CH * 4);
      float4 own[CH / 4];
      if constexpr (USE_OWN) {
#pragma unroll
        for (int v4 = 0; v4 < CH / 4; ++v4) own[v4] = lds128(cbase + own_off + v4 * 16);
      }
#pragma unroll
      for (int j = 0; j < UPT; ++j) {
        if (j < my_nu) {
          const uint32_t ab = n_gen & 1u;
          const uint32_t a_tmem = a_tmem0 + ab * 64u;
          bool waited = false;
#pragma unroll
          for (int h = 0; h < 4; ++h) {   // 8 cells at a time
            const float4 b0 = lds128(cbase + rb_off[j] + h * 32), b1 = lds128(cbase + rb_off[j] + h * 32 + 16);
            float4 a0, a1;
            if (USE_OWN && resident[j]) {
              a0 = own[2 * h];
              a1 = own[2 * h + 1];
            } else {
              a0 = lds128(cbase + ra_off[j] + h * 32);
              a1 = lds128(cbase + ra_off[j] + h * 32 + 16);
            }
            const float2 pa[4] = {make_float2(a0.x, a0.y), make_float2(a0.z, a0.w), make_float2(a1.x, a1.y), make_float2(a1.z, a1.w)};
            const float2 pb[4] = {make_float2(b0.x, b0.y), make_float2(b0.z, b0.w), make_float2(b1.x, b1.y), make_float2(b1.z, b1.w)};
            uint32_t hi[8], lo[8];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const float2 pr = __fmul2_rn(pa[i], pb[i]);
              // hi = product rounded to nearest tf32 (11 bits): the low part is then sign-symmetric and
              // at most 2^-12 |P|, so the tensor core's truncation of it costs 2^-23 instead of 2^-22
              const float2 ph = make_float2(__uint_as_float((__float_as_uint(pr.x) + 0x1000u) & 0xFFFFE000u),
                                            __uint_as_float((__float_as_uint(pr.y) + 0x1000u) & 0xFFFFE000u));
              const float2 pl = __ffma2_rn(ph, make_float2(-1.f, -1.f), pr);   // exact
              hi[2 * i] = __float_as_uint(ph.x); hi[2 * i + 1] = __float_as_uint(ph.y);
              lo[2 * i] = __float_as_uint(pl.x); lo[2 * i + 1] = __float_as_uint(pl.y);
            }
            if (!waited) {
              tick(t_gen);
              tc_wait(bar_aempty + ab * 8, ((n_gen >> 1) & 1u) ^ 1u);
              tc_fence_after();
              waited = true;
              tick(t_aempty);
            }
            tmem_st8(a_tmem + h * 8, hi);
            tmem_st8(a_tmem + 32 + h * 8, lo);
          }
          tick(t_gen);
          asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(bar_afull + ab * 8);
          ++n_gen;
          tick(t_st);
        }
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(bars + (RAW_EMPTY + buf) * 8);
  }
  if constexpr (PROF) {
    if (blockIdx.x == 0 && q == 0 && lane == 0 && p.prof) {
      long long* o = p.prof + team * 8;
      o[0] = t_raw; o[1] = t_gen; o[2] = t_aempty; o[3] = t_st; o[6] = n_gen;
    }
  }
}

// Drain warp (team, q): after each of its team's units has had its 12 MMAs, adds the accumulator of
// TMEM lanes [32q, 32q+32) into fp32 registers (round to nearest) and releases the accumulator;
// flushes to the float64 partials every flush_chunks chunks.
template <int J>
__device__ __forceinline__ void drain_unit(float2 (&acc)[cap(J) / 2], int ncol, uint32_t d_tmem, uint32_t bar_full,
                                           uint32_t bar_empty, uint32_t parity, int lane, long long& prof_wait,
                                           long long& prof_work) {
  long long t0 = 0;
  if constexpr (PROF) t0 = clock64();
  tc_wait(bar_full, parity);
  tc_fence_after();
  if constexpr (PROF) { const long long now = clock64(); prof_wait += now - t0; t0 = now; }
#pragma unroll
  for (int g0 = 0; g0 < cap(J) / 8; g0 += 3) {   // up to three 8-column groups in flight
    if (g0 * 8 < ncol) {
      float2 v[3][4];
#pragma unroll
      for (int g = 0; g < 3; ++g)
        if (g0 + g < cap(J) / 8 && (g0 + g) * 8 < ncol) tmem_ld8v(d_tmem + (g0 + g) * 8, v[g]);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
      for (int g = 0; g < 3; ++g)
        if (g0 + g < cap(J) / 8 && (g0 + g) * 8 < ncol) {
          pin8(v[g]);
#pragma unroll
          for (int c = 0; c < 4; ++c) acc[(g0 + g) * 4 + c] = __fadd2_rn(acc[(g0 + g) * 4 + c], v[g][c]);
        }
    }
  }
  tc_fence_before();
  __syncwarp();
  if (lane == 0) mbar_arrive(bar_empty);
  if constexpr (PROF) prof_work += clock64() - t0;
}
template <int J>
__device__ __forceinline__ void flush_unit(float2 (&acc)[cap(J) / 2], int ncol, double* my_partial) {
#pragma unroll
  for (int c = 0; c < cap(J) / 2; ++c)
    if (2 * c < ncol) {
      atomicAdd(my_partial + (int64_t)(capoff(J) + 2 * c) * 128, (double)acc[c].x);
      atomicAdd(my_partial + (int64_t)(capoff(J) + 2 * c + 1) * 128, (double)acc[c].y);
      acc[c] = make_float2(0.f, 0.f);
    }
}
template <int N>
__device__ __forceinline__ void zero_acc(float2 (&acc)[N]) {
#pragma unroll
  for (int c = 0; c < N; ++c) acc[c] = make_float2(0.f, 0.f);
}

__device__ __forceinline__ void drain_loop(const Params& p, uint32_t bars, uint32_t tbase, int team, int q, int lane) {
  float2 a0[cap(0) / 2], a1[cap(1) / 2], a2[cap(2) / 2], a3[cap(3) / 2];
  zero_acc(a0); zero_acc(a1); zero_acc(a2); zero_acc(a3);
  const int my_nu = p.nu[team];
  auto ncol = [&](int j) { return j < my_nu ? p.uncol[team * UPT + j] : 0; };
  const uint32_t d_base = tbase + ((uint32_t)(q * 32) << 16) + TM_D + (uint32_t)team * CAPSUM;
  const uint32_t bar_full = bars + (D_FULL + team * UPT) * 8, bar_empty = bars + (D_EMPTY + team * UPT) * 8;
  double* my_partial = p.partial + ((int64_t)blockIdx.x * NTEAMS + team) * p.partial_stride + q * 32 + lane;
  int since_flush = 0;
  long long d_wait = 0, d_work = 0;
  uint32_t n = 0;     // windows drained so far
  auto flush = [&]() {
    flush_unit<0>(a0, ncol(0), my_partial); flush_unit<1>(a1, ncol(1), my_partial);
    flush_unit<2>(a2, ncol(2), my_partial); flush_unit<3>(a3, ncol(3), my_partial);
  };
#define BSK_TC_DRAIN(J, ACC)                                                                              \
  if (J < my_nu)                                                                                          \
    drain_unit<J>(ACC, ncol(J), d_base + capoff(J), bar_full + J * 8, bar_empty + J * 8, n & 1u, lane,    \
                  d_wait, d_work);
  for (int64_t tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x) {
#pragma unroll 1
    for (int c = 0; c < NCH / WIN; ++c) {
      BSK_TC_DRAIN(0, a0) BSK_TC_DRAIN(1, a1) BSK_TC_DRAIN(2, a2) BSK_TC_DRAIN(3, a3)
      ++n;
      if ((since_flush += WIN) >= p.flush_chunks) {
        flush();
        since_flush = 0;
      }
    }
  }
#undef BSK_TC_DRAIN
  flush();
  if constexpr (PROF) {
    if (blockIdx.x == 0 && q == 0 && lane == 0 && p.prof) {
      p.prof[(NTEAMS + 1 + team) * 8 + 0] = d_wait;
      p.prof[(NTEAMS + 1 + team) * 8 + 1] = d_work;
      p.prof[(NTEAMS + 1 + team) * 8 + 2] = n;
    }
  }
}

__global__ void __launch_bounds__(NTHREADS, 1) tc_contract_kernel(const __grid_constant__ Params p) {
  extern __shared__ __align__(1024) unsigned char smem[];
  uint64_t* bar_ptr = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + OFF_BAR + NBAR * 8);
  const uint32_t bars = smem_u32(bar_ptr);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int R = p.nrows, N = p.ncols;
  constexpr int W_DRAIN = NTEAMS * 4, W_TMA = 2 * W_DRAIN, W_MMA = W_TMA + 1;

  if (tid == 0) {
    for (int b = 0; b < 2; ++b) {
      mbar_init(&bar_ptr[RAW_FULL + b], 1);
      mbar_init(&bar_ptr[RAW_EMPTY + b], NTEAMS * 4 + 2);
      mbar_init(&bar_ptr[B_FULL + b], 2);
      mbar_init(&bar_ptr[B_EMPTY + b], 1);
    }
    for (int b = 0; b < 2 * NTEAMS; ++b) {
      mbar_init(&bar_ptr[A_FULL + b], 4);
      mbar_init(&bar_ptr[A_EMPTY + b], 1);
    }
    for (int b = 0; b < NTEAMS * UPT; ++b) {
      mbar_init(&bar_ptr[D_FULL + b], 1);
      mbar_init(&bar_ptr[D_EMPTY + b], 4);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  // zero row of both raw buffers and the whole operand-image area (rows R..N-1 stay zero)
  for (int i = tid; i < RAW_STRIDE / 4; i += NTHREADS) {
    reinterpret_cast<uint32_t*>(smem + OFF_RAW + R * RAW_STRIDE)[i] = 0u;
    reinterpret_cast<uint32_t*>(smem + OFF_RAW + RAW_BYTES + R * RAW_STRIDE)[i] = 0u;
  }
  for (int i = tid; i < 4 * BIMG_BYTES / 4; i += NTHREADS) reinterpret_cast<uint32_t*>(smem + OFF_BIMG)[i] = 0u;
  if (warp == W_MMA) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(tmem_slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = *tmem_slot;

  if (warp < W_DRAIN) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(REGS_TEAM));
    team_loop(p, smem, bars, tbase, warp >> 2, warp & 3, lane);
  } else if (warp < W_TMA) {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(REGS_DRAIN));
    drain_loop(p, bars, tbase, (warp - W_DRAIN) >> 2, warp & 3, lane);
  } else {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(REGS_AUX));
    if (warp == W_TMA) {
      // ---- TMA producer
      int it = 0;
      for (int64_t tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x, ++it) {
        const int buf = it & 1;
        tc_wait(bars + (RAW_EMPTY + buf) * 8, ((uint32_t)(it >> 1) & 1u) ^ 1u);
        if (lane == 0) mbar_expect_tx(&bar_ptr[RAW_FULL + buf], (uint32_t)(R * TL * 4));
        __syncwarp();
        unsigned char* dst = smem + OFF_RAW + buf * RAW_BYTES;
        for (int r = lane; r < R; r += 32)
          bulk_g2s(dst + r * RAW_STRIDE, p.rowptr[r] + tile * TL, TL * 4, &bar_ptr[RAW_FULL + buf]);
      }
    } else if (warp == W_MMA) {
      // ---- MMA issuer
      const uint32_t leader = elect_one();
      const uint32_t lbo = (uint32_t)(N / 8) * 128u;     // next 4-cell group of the image
      uint32_t n_unit[NTEAMS];
      uint32_t idesc[NTEAMS * UPT], coff[NTEAMS * UPT];
#pragma unroll
      for (int t = 0; t < NTEAMS; ++t) n_unit[t] = 0u;
#pragma unroll
      for (int i = 0; i < NTEAMS * UPT; ++i) {
        idesc[i] = make_idesc_tf32(p.uncol[i]);
        coff[i] = (uint32_t)p.ucol0[i];                  // first column, = 16-byte units into an image group
      }
      long long m_afull = 0, m_dempty = 0, m_issue = 0, m_bfull = 0, m_mark = 0, m_start = 0;
      if constexpr (PROF) m_start = clock64();
      int it = 0;
      for (int64_t tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x, ++it) {
        const int buf = it & 1;
        if constexpr (PROF) m_mark = clock64();
        tc_wait(bars + (B_FULL + buf) * 8, (uint32_t)(it >> 1) & 1u);
        if constexpr (PROF) m_bfull += clock64() - m_mark;
        const uint32_t img_hi = smem_u32(smem + OFF_BIMG + (buf * 2 + 0) * BIMG_BYTES);
        const uint32_t img_lo = smem_u32(smem + OFF_BIMG + (buf * 2 + 1) * BIMG_BYTES);
        // descriptor of (image, 4-cell group g, first column col0): base + g * N + col0 in 16-byte units
        const uint32_t dh_lo = (uint32_t)make_desc(img_hi, lbo, 128u), dl_lo = (uint32_t)make_desc(img_lo, lbo, 128u);
        const uint32_t d_hi32 = (uint32_t)(make_desc(img_hi, lbo, 128u) >> 32);
#pragma unroll
        for (int c = 0; c < NCH; ++c) {
          const uint32_t n_win = (uint32_t)it * (NCH / WIN) + (uint32_t)(c / WIN);
#pragma unroll
          for (int j = 0; j < UPT; ++j) {
#pragma unroll
            for (int team = 0; team < NTEAMS; ++team) {
              if (j >= p.nu[team]) continue;
              const uint32_t n = n_unit[team]++;
              const uint32_t ab = n & 1u;
              if constexpr (PROF) m_mark = clock64();
              tc_wait(bars + (A_FULL + team * 2 + ab) * 8, (n >> 1) & 1u);
              if constexpr (PROF) { const long long now = clock64(); m_afull += now - m_mark; m_mark = now; }
              if (c % WIN == 0)    // first chunk of a window: the unit's accumulator must have been drained
                tc_wait(bars + (D_EMPTY + team * UPT + j) * 8, (n_win & 1u) ^ 1u);
              tc_fence_after();
              if constexpr (PROF) { const long long now = clock64(); m_dempty += now - m_mark; m_mark = now; }
              const uint32_t d = tbase + TM_D + (uint32_t)team * CAPSUM + capoff(j);
              const uint32_t a = tbase + TM_A + (uint32_t)team * 128u + ab * 64u;
              const uint32_t id = idesc[team * UPT + j];
              const uint32_t o0 = coff[team * UPT + j] + (uint32_t)(c * (CH / 4)) * (uint32_t)N;
              if (leader) {
#pragma unroll
                for (int ks = 0; ks < CH / 8; ++ks) {
                  const uint32_t o = o0 + (uint32_t)(ks * 2) * (uint32_t)N;
                  if (ks == 0 && c % WIN == 0) mma_tf32_ts2<0>(d, a + 32 + ks * 8, dh_lo + o, d_hi32, id);   // P_lo * C_hi
                  else mma_tf32_ts2<1>(d, a + 32 + ks * 8, dh_lo + o, d_hi32, id);
                  mma_tf32_ts2<1>(d, a + ks * 8, dl_lo + o, d_hi32, id);                       // P_hi * C_lo
                  mma_tf32_ts2<1>(d, a + ks * 8, dh_lo + o, d_hi32, id);                       // P_hi * C_hi
                }
                tc_commit(bars + (A_EMPTY + team * 2 + ab) * 8);
                if (c % WIN == WIN - 1) tc_commit(bars + (D_FULL + team * UPT + j) * 8);
              }
              __syncwarp();
              if constexpr (PROF) { const long long now = clock64(); m_issue += now - m_mark; m_mark = now; }
            }
          }
        }
        if (leader) tc_commit(bars + (B_EMPTY + buf) * 8);
        __syncwarp();
      }
      if constexpr (PROF) {
        if (blockIdx.x == 0 && lane == 0 && p.prof) {
          long long* o = p.prof + NTEAMS * 8;
          o[0] = m_afull; o[1] = m_dempty; o[2] = m_issue; o[3] = m_bfull; o[4] = clock64() - m_start;
        }
      }
    } else {
      // ---- operand-image converters (64 threads): raw fp32 -> tf32 hi (round to nearest) + lo
      const int ct = tid - (W_MMA + 1) * 32;
      const uint32_t ngrp = (uint32_t)(N / 8) * 128u;
      int it = 0;
      for (int64_t tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x, ++it) {
        const int buf = it & 1;
        tc_wait(bars + (RAW_FULL + buf) * 8, (uint32_t)(it >> 1) & 1u);
        tc_wait(bars + (B_EMPTY + buf) * 8, ((uint32_t)(it >> 1) & 1u) ^ 1u);
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
          mbar_arrive(bars + (B_FULL + buf) * 8);
          mbar_arrive(bars + (RAW_EMPTY + buf) * 8);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == W_MMA) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tbase) : "memory");
}

}  // namespace tc
}  // namespace bsk
