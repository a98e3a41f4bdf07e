// Tensor-core triangle contraction (sm_100a, tcgen05 + TMEM), included by contract.cu.
//
//   D[(a,b), c] = sum_x (F_a(x) F_b(x)) * F_c(x)        for row pairs (a,b) and rows c
//
// i.e. the reference's T separate np.sum(f[a]*f[b]*f[c]) loops (bskit/main.py:1871-1879) recast
// as a skinny GEMM over the grid axis: M = row pairs (128 per "unit"), N = the rows c the unit's
// triangles need, K = cells.  Measured on B200 (profiles/r1_tcgen05_probe.txt): with M = 128 a
// tcgen05.mma costs max(~11, 128*N/256) cycles only when the A operand comes from TMEM, so the
// pair products -- generated on the CUDA cores anyway -- are written straight to TMEM with
// tcgen05.st and never touch shared memory; B = the fields, converted once per tile into K-major
// core matrices (hi and lo images).
//
// What bounds the kernel is shared-memory bandwidth: every pair row needs the cells of its two
// fields in the registers of the lane that owns the row.  Round 1 read both rows per lane and
// unit (8 wavefronts per 4 cells and warp).  Measured (scripts/dev/lds_probe.cu,
// profiles/r2_lds_probe.txt): a warp-wide LDS.128 costs 4 cycles when the lanes read 32 different
// 16-byte chunks, 2 cycles when every even/odd lane pair reads the same chunk.  The layout here:
//   * rows are handled in groups of 8; a generator warp "hosts" one group for the whole chunk:
//     lane l keeps row 8*host + (l & 7) of the chunk in registers (32 cells) across the units;
//   * in unit j the warp meets a partner group: lane pair k = l >> 1 reads row
//     8*part + ((k & 7) + s) & 7, s = 2*piece + (k >> 3) -- lane pairs share the chunk they read
//     (2-cycle LDS), the 8 pairs of a half warp read 8 consecutive rows (distinct banks), and
//     two such "pieces" cover all 64 pairs of two groups.
// Shared-memory reads per 4 cells, warp and unit drop from 8 wavefronts to 2 (+ 4/U for the
// resident row, U = units that share the host group).
//
// Numerics.  3xTF32: P = F_a*F_b (fp32, RN) is split P = P_hi + P_lo, F_c = C_hi + C_lo with
// 11-bit pieces (hi parts rounded to nearest), D += P_lo*C_hi + P_hi*C_lo + P_hi*C_hi.  The
// tensor core truncates its fp32 accumulator (profiles/r1_tensor_core_rounding_probe.txt), so an
// accumulator lives in TMEM for one window of WIN chunks only; it is then drained into fp32
// registers (round-to-nearest adds), flushed into float64 partials every flush_chunks chunks.
// Window ends are staggered over the units so the TMEM reads (64 B/cycle/SM) spread evenly.
//
// Roles (640 threads, one CTA per SM, persistent over tiles of 128 cells; setmaxnreg 96/120/40):
//   2 x 4 warps  generator teams: warp q of a team owns TMEM lanes [32q, 32q+32)
//   2 x 4 warps  drain warpgroups, one per team: second-level accumulators of the team's units
//                (96 columns per team, handed out to the units in blocks of 8)
//   1 warp       TMA producer (raw [row][cell] tiles, cp.async.bulk + mbarrier), up to 3 tiles ahead
//   2 warps      MMA issuers, one per team (one elected lane each)
//   (the generators also convert the raw tile into the operand images at the start of a tile)
//
// A launch processes one "pass": <= 8 units with <= 96 accumulator columns per team, over a
// window of <= 40 column rows and <= MAXRAW raw rows.  Lists that need more (S = 80 bins, 2-3
// field cross lists) run as several passes (host schedule: contract.cu, build_tc_schedule_host).
#pragma once

namespace bsk {
namespace tc {

constexpr int CH = 32;            // cells per chunk = K extent of one unit
constexpr int TL = 128;           // cells per tile
constexpr int NCH = TL / CH;
#ifndef BSK_TC_WIN
#define BSK_TC_WIN 4
#endif
constexpr int WIN = BSK_TC_WIN;   // chunks accumulated in TMEM before a drain
constexpr int MAXCOL = 40;        // column rows per pass (N of an MMA <= 40)
constexpr int MAXRAW = 120;       // raw rows per pass (8-row groups: 15)
constexpr int NTEAMS = 2;
constexpr int UPT = 4;            // units per team
constexpr int NUNITS = NTEAMS * UPT;
constexpr int NGEN = NTEAMS * 128;
constexpr int NTHREADS = 2 * NGEN + 128;    // generator teams + one drain warpgroup per team + 4 auxiliary warps
#ifndef BSK_TC_PROF
#define BSK_TC_PROF 0
#endif
constexpr bool PROF = BSK_TC_PROF;   // cycle counters per phase (block 0), see Params::prof
// 640 threads are launched with 96 registers each; setmaxnreg then moves registers from the
// auxiliary warps to the drain warpgroups, which hold the second-level accumulators
constexpr int REGS_GEN = 96, REGS_DRAIN = 120, REGS_AUX = 48;
static_assert(NGEN * (REGS_GEN + REGS_DRAIN) + 128 * REGS_AUX <= NTHREADS * 96, "register split");
// accumulator columns per team: 12 blocks of 8 columns, handed out to the team's units by the
// host (Params::ublk0); a unit of width ncol uses ncol/8 consecutive blocks
constexpr int TEAMCOLS = 96, NBLK = TEAMCOLS / 8;
constexpr int CAPTOT = NTEAMS * TEAMCOLS;   // 192
// TMEM columns: the teams' accumulators and two A buffers (hi 32 + lo 32 columns) per team
constexpr int TM_D = 0, TM_A = 256;
static_assert(CAPTOT <= TM_A && TM_A + NTEAMS * 2 * 64 <= 512, "TMEM budget");

constexpr int RAW_STRIDE = TL * 4 + 16;            // bytes; 8 consecutive rows tile the 32 banks
constexpr int MAXRAWBUF = 4;                       // raw tiles in flight (as many as fit next to the images)
constexpr int BIMG_BYTES = (TL / 4) * (MAXCOL / 8) * 128;   // one hi or lo image of a tile
constexpr int OFF_BAR = 0;
constexpr int OFF_BIMG = 512;                      // [buf][hi|lo]
constexpr int OFF_RAW = OFF_BIMG + 4 * BIMG_BYTES;
constexpr int SMEM_MAX = 227 * 1024;
static_assert(OFF_BIMG % 128 == 0 && OFF_RAW % 128 == 0, "operand images must be 128-byte aligned");
// bytes of one raw buffer: nraw rows + one all-zero row
__host__ __device__ constexpr int raw_bytes(int nraw) { return ((nraw + 1) * RAW_STRIDE + 127) / 128 * 128; }
__host__ __device__ constexpr int raw_bufs(int nraw) {
  return (SMEM_MAX - OFF_RAW) / raw_bytes(nraw) < MAXRAWBUF ? (SMEM_MAX - OFF_RAW) / raw_bytes(nraw) : MAXRAWBUF;
}
static_assert(raw_bufs(MAXRAW) >= 2, "shared memory");

// barrier indices
enum {
  RAW_FULL = 0, RAW_EMPTY = RAW_FULL + MAXRAWBUF, B_FULL = RAW_EMPTY + MAXRAWBUF, B_EMPTY = B_FULL + 2,
  A_FULL = B_EMPTY + 2, A_EMPTY = A_FULL + 2 * NTEAMS, D_FULL = A_EMPTY + 2 * NTEAMS, D_EMPTY = D_FULL + NUNITS,
  NBAR = D_EMPTY + NUNITS
};
static_assert(NBAR * 8 + 8 <= OFF_BIMG, "barrier area");

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
// named (hardware) barriers for the generator -> MMA issuer hand-over: a hop costs ~16 cycles instead
// of ~110 through an mbarrier (profiles/r2_sync_probe.txt).  id in [1, 15]; count = the team's 4 generator
// warps + the issuer warp = 160 threads
__device__ __forceinline__ void nbar_arrive(int id, int count) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(count) : "memory"); }
__device__ __forceinline__ void nbar_sync(int id, int count) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory"); }
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
__device__ __forceinline__ void mma_tf32_ts(uint32_t d, uint32_t a, uint32_t desc_lo, uint32_t desc_hi, uint32_t idesc,
                                            uint32_t acc) {
  asm volatile("{\n.reg .pred p;\n.reg .b64 bd;\nsetp.ne.b32 p, %5, 0;\nmov.b64 bd, {%2, %3};\n"
               "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], bd, %4, p;\n}\n" ::"r"(d), "r"(a), "r"(desc_lo), "r"(desc_hi),
               "r"(idesc), "r"(acc)
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
// Shared-memory load that the compiler may hoist, batch and reorder freely (no volatile, no memory
// clobber).  The tile must not change while such loads are in flight, and the address must depend
// on order_token() of the wait that made the tile visible.
__device__ __forceinline__ float4 lds128(uint32_t addr) {
  float4 v;
  asm("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}
// an opaque zero produced after everything before it (volatile + memory clobber): adding it to an
// address keeps the loads through that address behind the wait
__device__ __forceinline__ uint32_t order_token() {
  uint32_t t;
  asm volatile("mov.u32 %0, 0;" : "=r"(t)::"memory");
  return t;
}
// L2 policies: the fields stream through once per pass (evict first), the float64 partials are
// re-read by every flush (evict last) -- keeps the flushes out of DRAM
__device__ __forceinline__ uint64_t policy_evict_first() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ uint64_t policy_evict_last() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ void bulk_g2s_hint(void* dst_smem, const void* src, uint32_t bytes, uint64_t* bar, uint64_t pol) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::
          "r"(smem_u32(dst_smem)),
      "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(pol)
      : "memory");
}
// float -> double conversion inside the asm statement: left to the compiler, the conversions of
// all accumulators are hoisted in front of the reductions and double the register demand
__device__ __forceinline__ void red_add_f64_hint(double* addr, float v, uint64_t pol) {
  asm volatile("{\n.reg .f64 d;\ncvt.f64.f32 d, %1;\nred.global.add.L2::cache_hint.f64 [%0], d, %2;\n}\n" ::"l"(addr), "f"(v), "l"(pol)
               : "memory");
}

struct Params {
  const float* const* rowptr;   // field-table rows (device pointers)
  const int* rawrow;            // [nraw]: field-table row loaded into raw slot s, -1: the slot stays zero
  int nraw;                     // raw slots (multiple of 8, <= MAXRAW); slot nraw is the zero row
  int nbuf, rawb;               // raw tiles in flight (2..4) and bytes per raw buffer
  int ncols;                    // column rows of the window (multiple of 8, <= MAXCOL)
  int colslot[MAXCOL / 8];      // raw slot of the first row of each 8-column block
  int64_t ntiles;               // ncells / TL
  int nu[NTEAMS];               // units per team
  int ucol0[NUNITS];            // [team * UPT + j]: first D column the unit needs (multiple of 8)
  int uncol[NUNITS];            //                   number of columns (multiple of 8)
  int ublk0[NUNITS];            //                   first accumulator block of the unit within its team
  const uint32_t* lane_tab;     // [team * UPT + j][128]: a_slot | b_slot << 8
  double* partial;              // [cta][team][TEAMCOLS][128]
  int flush_chunks;
  long long* prof;              // PROF only
};

// window bookkeeping shared by the MMA issuers and the drain warps: unit u closes its window
// after chunk g (running count over this CTA's chunks) when (g + u) % WIN == WIN - 1 or g is
// the last chunk
__device__ __forceinline__ bool win_closes(uint32_t g, int u, uint32_t gtot) {
  return ((g + (uint32_t)u) % WIN) == WIN - 1 || g + 1 == gtot;
}
__device__ __forceinline__ bool win_opens(uint32_t g, int u) { return g == 0 || ((g + (uint32_t)u) % WIN) == 0; }

// operand-image conversion, done by the 256 generator threads at the start of a tile (a separate
// converter warp pair could not keep up once the generators got faster): raw fp32 -> tf32 hi
// (round to nearest) + lo.  Thread gt owns column gt & 63 (if it exists) and the 4-cell groups
// g = (gt >> 6) + 4k: 8 independent items per thread, 5 % of a tile's generator instructions.
template <int UNR>
__device__ __forceinline__ void convert_share(const Params& p, uint32_t raw, uint32_t img_hi, uint32_t img_lo, int gt,
                                              uint32_t col_slot) {
  const int l = gt & 63;
  if (l < p.ncols) {
    const uint32_t ngrp = (uint32_t)(p.ncols / 8) * 128u;
    const uint32_t row = raw + col_slot * RAW_STRIDE;
    const uint32_t o = (uint32_t)(l >> 3) * 128u + (uint32_t)(l & 7) * 16u;
#pragma unroll UNR
    for (int k = 0; k < TL / 16; ++k) {
      const uint32_t g = (uint32_t)(gt >> 6) + 4u * k;
      const float4 v = lds128(row + g * 16u);
      const float x[4] = {v.x, v.y, v.z, v.w};
      uint32_t h[4], lo[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        h[e] = (__float_as_uint(x[e]) + 0x1000u) & 0xFFFFE000u;
        lo[e] = __float_as_uint(x[e] - __uint_as_float(h[e]));
      }
      sts128(img_hi + g * ngrp + o, h[0], h[1], h[2], h[3]);
      sts128(img_lo + g * ngrp + o, lo[0], lo[1], lo[2], lo[3]);
    }
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// Generator team member; q = warp % 4 is the TMEM lane quarter.  Pure producer: writes the
// 128 x 32 pair products (hi, lo) of one unit into one of the team's two A buffers.  The
// hand-over of a unit (tcgen05.wait::st + arrive) is deferred until the first products of the next
// unit are in registers, and the A buffer is probed (test_wait) before the loads are issued, so
// neither latency is exposed; a warp without a piece in a unit only keeps the protocol going.
__device__ __forceinline__ uint32_t mbar_test(uint32_t bar_addr, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n.reg .pred p;\nmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
               : "=r"(ok) : "r"(bar_addr), "r"(parity) : "memory");
  return ok;
}
__device__ __forceinline__ void team_loop(const Params& p, const unsigned char* smem, uint32_t bars, uint32_t tbase,
                                          int team, int q, int lane) {
  const int my_nu = p.nu[team];
  uint32_t rb_off[UPT], ra_off[UPT];
  bool reload[UPT];     // warp-uniform: the resident row changes with this unit
  bool idle[UPT];       // warp-uniform: this warp has no piece in the unit
#pragma unroll
  for (int j = 0; j < UPT; ++j) {
    const uint32_t none = (uint32_t)p.nraw | ((uint32_t)p.nraw << 8);
    uint32_t e = none;
    if (j < my_nu) e = p.lane_tab[((team * UPT + j) * 4 + q) * 32 + lane];
    ra_off[j] = (e & 0xFFu) * RAW_STRIDE;
    rb_off[j] = ((e >> 8) & 0xFFu) * RAW_STRIDE;
    idle[j] = __all_sync(0xffffffffu, e == none);
    reload[j] = j == 0 || __any_sync(0xffffffffu, ra_off[j] != ra_off[j - 1]);
  }
  // the resident row must be (re)loaded in the first non-idle unit of a chunk
  {
    bool have = false;
#pragma unroll
    for (int j = 0; j < UPT; ++j) {
      if (idle[j]) continue;
      if (!have) reload[j] = true;
      have = true;
    }
  }
  // A team with exactly two units alternates their order from chunk to chunk (even chunks: unit 1
  // first): a unit whose window closes is then the FIRST of its chunk and the second of the next, so
  // two unit slots instead of one separate the last MMA of a window from the first of the next and the
  // drain (TMEM reads: 64 B/cycle/SM) no longer stalls the in-order issuer.  Issuer and generator
  // compute the same order; the drain only depends on which unit closes after which chunk.
  const bool alt = my_nu == 2;
  const bool same01 = __all_sync(0xffffffffu, ra_off[0] == ra_off[1]);   // alt: resident row shared by both units
  const uint32_t lane_sel = (uint32_t)(q * 32) << 16;
  const uint32_t a_tmem0 = tbase + lane_sel + TM_A + (uint32_t)team * 128u;
  const uint32_t bar_afull = bars + (A_FULL + team * 2) * 8, bar_aempty = bars + (A_EMPTY + team * 2) * 8;
  uint32_t n_gen = 0;     // units generated by this team so far
  bool pending = false;   // the previous unit's stores have not been handed over yet
  uint32_t pending_ab = 0;
  long long t_raw = 0, t_gen = 0, t_aempty = 0, t_st = 0, t_mark = 0;
  auto tick = [&](long long& acc_t) {
    if constexpr (PROF) { const long long now = clock64(); acc_t += now - t_mark; t_mark = now; }
  };
  auto hand_over = [&]() {
    if (pending) {
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      tc_fence_before();
      nbar_arrive(1 + team * 2 + (int)pending_ab, 160);
      pending = false;
    }
  };
  if constexpr (PROF) t_mark = clock64();

  int it = 0;
  for (int64_t tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x, ++it) {
    const int buf = it % p.nbuf;
    tc_wait(bars + (RAW_FULL + buf) * 8, (uint32_t)(it / p.nbuf) & 1u);
    tick(t_raw);
    const uint32_t raw = smem_u32(smem + OFF_RAW + buf * p.rawb) + order_token();
#pragma unroll 1
    for (int c = 0; c < NCH; ++c) {
      const uint32_t cbase = raw + c * (CH * 4);
      float4 own[CH / 4];
      bool have_own = false;      // alt teams: the resident row of this chunk is in registers
#pragma unroll
      for (int j = 0; j < UPT; ++j) {
        if (j < my_nu) {
          // unit executed in slot j of this chunk, and its (warp-uniform) flags and (per-lane) row offsets
          const bool swp = alt && !(c & 1);
          const uint32_t ra_j = j < 2 && swp ? ra_off[j ^ 1] : ra_off[j], rb_j = j < 2 && swp ? rb_off[j ^ 1] : rb_off[j];
          const bool idle_j = j < 2 && swp ? idle[j ^ 1] : idle[j];
          const bool reload_j = alt ? (!have_own || !same01) : reload[j];
          const uint32_t ab = n_gen & 1u;
          const uint32_t a_tmem = a_tmem0 + ab * 64u;
          const uint32_t ok = mbar_test(bar_aempty + ab * 8, ((n_gen >> 1) & 1u) ^ 1u);
          if (idle_j) {
            hand_over();
            if (!ok) tc_wait(bar_aempty + ab * 8, ((n_gen >> 1) & 1u) ^ 1u);
          } else {
            if (reload_j) {
#pragma unroll
              for (int v4 = 0; v4 < CH / 4; ++v4) own[v4] = lds128(cbase + ra_j + v4 * 16);
              have_own = true;
            }
#pragma unroll
            for (int h = 0; h < 4; ++h) {   // 8 cells at a time
              const float4 b0 = lds128(cbase + rb_j + h * 32), b1 = lds128(cbase + rb_j + h * 32 + 16);
              const float4 a0 = own[2 * h], a1 = own[2 * h + 1];
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
              if (h == 0) {
                tick(t_gen);
                hand_over();          // the previous unit: its stores have long landed
                tick(t_st);
                if (!ok) tc_wait(bar_aempty + ab * 8, ((n_gen >> 1) & 1u) ^ 1u);
                tc_fence_after();
                tick(t_aempty);
              }
              tmem_st8(a_tmem + h * 8, hi);
              tmem_st8(a_tmem + 32 + h * 8, lo);
            }
            tick(t_gen);
          }
          pending = true;
          pending_ab = ab;
          ++n_gen;
        }
      }
    }
    hand_over();     // before the raw buffer is released and a (possibly long) wait for the next tile
    tick(t_st);
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

// Drain warp (team, q): after a unit's window has had its MMAs, adds the accumulator of TMEM
// lanes [32q, 32q+32) into fp32 registers (round to nearest) and releases the accumulator.  The
// registers are 12 blocks of 8 columns; the block a column group goes to is warp-uniform.
// (separate members and an if-chain: an array indexed through a switch is turned into a dynamically
// indexed local-memory array by the compiler)
struct DrainAcc {
  float2 b0[4], b1[4], b2[4], b3[4], b4[4], b5[4], b6[4], b7[4], b8[4], b9[4], b10[4], b11[4];
};
#define BSK_TC_FOR_BLOCKS(X) X(0, b0) X(1, b1) X(2, b2) X(3, b3) X(4, b4) X(5, b5) X(6, b6) X(7, b7) X(8, b8) X(9, b9) X(10, b10) X(11, b11)
__device__ __forceinline__ void add4(float2 (&a)[4], const float2 (&v)[4]) {
#pragma unroll
  for (int c = 0; c < 4; ++c) a[c] = __fadd2_rn(a[c], v[c]);
}
__device__ __forceinline__ void add_block(DrainAcc& acc, int blk, const float2 (&v)[4]) {
#define BSK_TC_ADD(K, M) if (blk == K) { add4(acc.M, v); return; }
  BSK_TC_FOR_BLOCKS(BSK_TC_ADD)
#undef BSK_TC_ADD
}
__device__ __forceinline__ void drain_unit(DrainAcc& acc, int blk0, int nblk, uint32_t d_tmem, uint32_t bar_full,
                                           uint32_t bar_empty, uint32_t parity, int lane, long long& prof_wait,
                                           long long& prof_work) {
  long long t0 = 0;
  if constexpr (PROF) t0 = clock64();
  tc_wait(bar_full, parity);
  tc_fence_after();
  if constexpr (PROF) { const long long now = clock64(); prof_wait += now - t0; t0 = now; }
#pragma unroll 1
  for (int g0 = 0; g0 < nblk; ++g0) {   // one 8-column group at a time (the drain is not throughput critical)
    float2 v0[4];
    tmem_ld8v(d_tmem + g0 * 8, v0);
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    pin8(v0);
    add_block(acc, blk0 + g0, v0);
  }
  tc_fence_before();
  nbar_arrive((int)bar_empty, 160);      // named barrier 5 + unit: the issuer may start the unit's next window
  if constexpr (PROF) prof_work += clock64() - t0;
}

__device__ __forceinline__ void drain_loop(const Params& p, const unsigned char* smem, uint32_t bars, uint32_t tbase, int team,
                                           int q, int lane) {
  DrainAcc acc;
#define BSK_TC_ZERO(K, M) _Pragma("unroll") for (int c = 0; c < 4; ++c) acc.M[c] = make_float2(0.f, 0.f);
  BSK_TC_FOR_BLOCKS(BSK_TC_ZERO)
#undef BSK_TC_ZERO
  // few live scalars next to the 96 accumulators (120 registers): the units' first block and block
  // count are packed into one word (8 bits per unit), the window parities into another
  const int my_nu = p.nu[team];
  uint32_t cfg = 0, used = 0;
#pragma unroll
  for (int j = 0; j < UPT; ++j) {
    const uint32_t b0 = (uint32_t)p.ublk0[team * UPT + j], nb = j < my_nu ? (uint32_t)p.uncol[team * UPT + j] / 8u : 0u;
    cfg |= (b0 | (nb << 4)) << (8 * j);
    used = max(used, b0 + nb);
  }
  const uint32_t d_base = tbase + ((uint32_t)(q * 32) << 16) + TM_D + (uint32_t)team * TEAMCOLS;
  const uint32_t bar_full = bars + (D_FULL + team * UPT) * 8;     // D_EMPTY follows NUNITS barriers later
  uint32_t parity = 0;     // bit j: parity of unit j's next window
  int since_flush = 0;
  long long d_wait = 0, d_work = 0;
  uint32_t n0 = 0;
  auto flush = [&]() {
    double* my_partial = p.partial + ((int64_t)blockIdx.x * NTEAMS + team) * ((int64_t)TEAMCOLS * 128) + q * 32 + lane;
    const uint64_t pol = policy_evict_last();
#define BSK_TC_FLUSH(K, M)                                                                         \
  if (K < used) {                                                                                  \
    _Pragma("unroll") for (int c = 0; c < 4; ++c) {                                                \
      red_add_f64_hint(my_partial + (int64_t)(K * 8 + 2 * c) * 128, acc.M[c].x, pol);      \
      red_add_f64_hint(my_partial + (int64_t)(K * 8 + 2 * c + 1) * 128, acc.M[c].y, pol);  \
      acc.M[c] = make_float2(0.f, 0.f);                                                            \
    }                                                                                              \
  }
    BSK_TC_FOR_BLOCKS(BSK_TC_FLUSH)
#undef BSK_TC_FLUSH
  };
  const uint32_t gtot = (uint32_t)((p.ntiles > (int64_t)blockIdx.x ? (p.ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0) * NCH);
  // The drain warps, which wait most of the time, also turn raw tile `it` into the operand images
  // (hi = tf32 rounded to nearest, lo = rest) one tile ahead of the MMAs.  On the generator warps the
  // same conversion stalled the A pipeline at every tile boundary (7 ms of 50,
  // profiles/r2_tc_contract_history.md).
  const int gt = team * 128 + q * 32 + lane;
  const uint32_t col_slot = (gt & 63) < p.ncols ? (uint32_t)(p.colslot[(gt & 63) >> 3] + (gt & 7)) : 0u;
  auto convert_tile = [&](uint32_t it) {
    const uint32_t buf = it % (uint32_t)p.nbuf, bbuf = it & 1u;
    tc_wait(bars + (RAW_FULL + buf) * 8, (it / (uint32_t)p.nbuf) & 1u);
    tc_wait(bars + (B_EMPTY + bbuf) * 8, ((it >> 1) & 1u) ^ 1u);     // the MMAs of tile it - 2 have released the images
    const uint32_t raw = smem_u32(smem + OFF_RAW + buf * p.rawb) + order_token();
    convert_share<1>(p, raw, smem_u32(smem + OFF_BIMG + (bbuf * 2 + 0) * BIMG_BYTES),
                     smem_u32(smem + OFF_BIMG + (bbuf * 2 + 1) * BIMG_BYTES), gt, col_slot);
    __syncwarp();
    if (lane == 0) {
      mbar_arrive(bars + (B_FULL + bbuf) * 8);
      mbar_arrive(bars + (RAW_EMPTY + buf) * 8);
    }
  };
  if (gtot > 0) convert_tile(0);
  auto drain_j = [&](uint32_t j) {
    const uint32_t nb = (cfg >> (8 * j + 4)) & 15u, b0 = (cfg >> (8 * j)) & 15u;
    if (nb > 0) {
      drain_unit(acc, (int)b0, (int)nb, d_base + b0 * 8, bar_full + j * 8, (uint32_t)(5 + team * UPT + (int)j), (parity >> j) & 1u,
                 lane, d_wait, d_work);
      parity ^= 1u << j;
      if (j == 0) ++n0;
    }
  };
#pragma unroll 1
  for (uint32_t g = 0; g + 1 < gtot; ++g) {
    if (g % NCH == 0 && g + NCH < gtot) convert_tile(g / NCH + 1);
    // the units that close their window after chunk g: (g + team * UPT + j) % WIN == WIN - 1
    for (uint32_t j = (2 * WIN - 1 - (g + (uint32_t)team * UPT) % WIN) % WIN; j < UPT; j += WIN) drain_j(j);
    if (++since_flush >= p.flush_chunks) {
      flush();
      since_flush = 0;
    }
  }
  if (gtot > 0)      // after the last chunk every unit closes
    for (uint32_t j = 0; j < UPT; ++j) drain_j(j);
  flush();
  if constexpr (PROF) {
    if (blockIdx.x == 0 && q == 0 && lane == 0 && p.prof) {
      p.prof[(2 * NTEAMS + team) * 8 + 0] = d_wait;
      p.prof[(2 * NTEAMS + team) * 8 + 1] = d_work;
      p.prof[(2 * NTEAMS + team) * 8 + 2] = n0;
    }
  }
}

// MMA issuer of one team: 12 MMAs per unit and chunk into the unit's own accumulator
__device__ __forceinline__ void mma_loop(const Params& p, const unsigned char* smem, uint32_t bars, uint32_t tbase, int team) {
  const uint32_t leader = elect_one();
  const int N = p.ncols;
  const uint32_t lbo = (uint32_t)(N / 8) * 128u;     // next 4-cell group of the image
  const int my_nu = p.nu[team];
  // (per-unit constants are read from the parameter bank where they are used: the issuer runs on 48 registers)
  const int64_t my_tiles = p.ntiles > (int64_t)blockIdx.x ? (p.ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
  const uint32_t gtot = (uint32_t)(my_tiles * NCH);
  uint32_t n_unit = 0, g = 0;
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
    // descriptor of (image, 4-cell group gq, first column col0): base + gq * N + col0 in 16-byte units
    const uint32_t dh_lo = (uint32_t)make_desc(img_hi, lbo, 128u), dl_lo = (uint32_t)make_desc(img_lo, lbo, 128u);
    const uint32_t d_hi32 = (uint32_t)(make_desc(img_hi, lbo, 128u) >> 32);
#pragma unroll 1
    for (int c = 0; c < NCH; ++c, ++g) {
#pragma unroll
      for (int j = 0; j < UPT; ++j) {
        if (j >= my_nu) continue;
        const int jj = (my_nu == 2 && !(c & 1)) ? (j ^ 1) : j;      // the generators' unit order (team_loop)
        const int u = team * UPT + jj;
        const uint32_t n = n_unit++;
        const uint32_t ab = n & 1u;
        if constexpr (PROF) m_mark = clock64();
        nbar_sync(1 + team * 2 + (int)ab, 160);
        if constexpr (PROF) { const long long now = clock64(); m_afull += now - m_mark; m_mark = now; }
        const bool opens = win_opens(g, u);
        if (opens && g > 0)    // first chunk of a later window: the unit's accumulator must have been drained
          nbar_sync(5 + u, 160);
        tc_fence_after();
        if constexpr (PROF) { const long long now = clock64(); m_dempty += now - m_mark; m_mark = now; }
        const uint32_t d = tbase + TM_D + (uint32_t)team * TEAMCOLS + (uint32_t)p.ublk0[u] * 8u;
        const uint32_t a = tbase + TM_A + (uint32_t)team * 128u + ab * 64u;
        const uint32_t id = make_idesc_tf32(p.uncol[u]);
        // first column of the unit (= 16-byte units into an image group) + the chunk's 4-cell groups
        const uint32_t o0 = (uint32_t)p.ucol0[u] + (uint32_t)(c * (CH / 4)) * (uint32_t)N;
        const bool closes = win_closes(g, u, gtot);
        if (leader) {
#pragma unroll
          for (int ks = 0; ks < CH / 8; ++ks) {
            const uint32_t o = o0 + (uint32_t)(ks * 2) * (uint32_t)N;
            mma_tf32_ts(d, a + 32 + ks * 8, dh_lo + o, d_hi32, id, (ks == 0 && opens) ? 0u : 1u);   // P_lo * C_hi
            mma_tf32_ts(d, a + ks * 8, dl_lo + o, d_hi32, id, 1u);                                  // P_hi * C_lo
            mma_tf32_ts(d, a + ks * 8, dh_lo + o, d_hi32, id, 1u);                                  // P_hi * C_hi
          }
          tc_commit(bars + (A_EMPTY + team * 2 + ab) * 8);
          if (closes) tc_commit(bars + (D_FULL + u) * 8);
        }
        __syncwarp();
        if constexpr (PROF) { const long long now = clock64(); m_issue += now - m_mark; m_mark = now; }
      }
    }
    if (leader) tc_commit(bars + (B_EMPTY + buf) * 8);
    __syncwarp();
  }
  if constexpr (PROF) {
    if (blockIdx.x == 0 && (threadIdx.x & 31) == 0 && p.prof) {
      long long* o = p.prof + (NTEAMS + team) * 8;
      o[0] = m_afull; o[1] = m_dempty; o[2] = m_issue; o[3] = m_bfull; o[4] = clock64() - m_start;
    }
  }
}

__global__ void __launch_bounds__(NTHREADS, 1) tc_contract_kernel(const __grid_constant__ Params p) {
  extern __shared__ __align__(1024) unsigned char smem[];
  uint64_t* bar_ptr = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + OFF_BAR + NBAR * 8);
  const uint32_t bars = smem_u32(bar_ptr);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  constexpr int W_DRAIN = NTEAMS * 4, W_TMA = 2 * W_DRAIN, W_MMA0 = W_TMA + 1, W_CONV = W_MMA0 + NTEAMS;

  if (tid == 0) {
    for (int b = 0; b < MAXRAWBUF; ++b) {
      mbar_init(&bar_ptr[RAW_FULL + b], 1);
      mbar_init(&bar_ptr[RAW_EMPTY + b], 2 * NTEAMS * 4);     // generator warps (pair products) + drain warps (images)
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&bar_ptr[B_FULL + b], NTEAMS * 4);
      mbar_init(&bar_ptr[B_EMPTY + b], NTEAMS);
    }
    for (int b = 0; b < 2 * NTEAMS; ++b) {
      mbar_init(&bar_ptr[A_FULL + b], 4);
      mbar_init(&bar_ptr[A_EMPTY + b], 1);
    }
    for (int b = 0; b < NUNITS; ++b) {
      mbar_init(&bar_ptr[D_FULL + b], 1);
      mbar_init(&bar_ptr[D_EMPTY + b], 4);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  // the operand images and every raw buffer (rows that are never loaded must read as zero)
  for (int i = tid; i < (4 * BIMG_BYTES + p.nbuf * p.rawb) / 16; i += NTHREADS)
    reinterpret_cast<uint4*>(smem + OFF_BIMG)[i] = make_uint4(0u, 0u, 0u, 0u);
  if (warp == W_MMA0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(tmem_slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = *tmem_slot;

  if (warp < W_DRAIN) {
    team_loop(p, smem, bars, tbase, warp >> 2, warp & 3, lane);     // keeps its 96 registers
  } else if (warp < W_TMA) {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(REGS_DRAIN));
    drain_loop(p, smem, bars, tbase, (warp - W_DRAIN) >> 2, warp & 3, lane);
  } else {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(REGS_AUX));
    if (warp == W_TMA) {
      // ---- TMA producer: raw [row][cell] tiles, nbuf - 1 tiles ahead of the generators
      const uint64_t pol = policy_evict_first();
      // this lane's raw rows (slots lane, lane + 32, ...): pointers of the rows that are really loaded
      const float* src[(MAXRAW + 31) / 32];
      int nload = 0;
#pragma unroll
      for (int k = 0; k < (MAXRAW + 31) / 32; ++k) {
        const int sl = lane + 32 * k;
        const int r = sl < p.nraw ? p.rawrow[sl] : -1;
        src[k] = r >= 0 ? p.rowptr[r] : nullptr;
        nload += r >= 0;
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) nload += __shfl_xor_sync(0xffffffffu, nload, o);
      int it = 0;
      for (int64_t tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x, ++it) {
        const int buf = it % p.nbuf;
        tc_wait(bars + (RAW_EMPTY + buf) * 8, ((uint32_t)(it / p.nbuf) & 1u) ^ 1u);
        if (lane == 0) mbar_expect_tx(&bar_ptr[RAW_FULL + buf], (uint32_t)(nload * TL * 4));
        __syncwarp();
        unsigned char* dst = const_cast<unsigned char*>(smem) + OFF_RAW + buf * p.rawb;
#pragma unroll
        for (int k = 0; k < (MAXRAW + 31) / 32; ++k)
          if (src[k]) bulk_g2s_hint(dst + (lane + 32 * k) * RAW_STRIDE, src[k] + tile * TL, TL * 4, &bar_ptr[RAW_FULL + buf], pol);
      }
    } else if (warp == W_CONV) {
      // spare warp
    } else {
      mma_loop(p, smem, bars, tbase, warp - W_MMA0);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == W_MMA0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tbase) : "memory");
}

}  // namespace tc
}  // namespace bsk
