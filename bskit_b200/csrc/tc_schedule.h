// Host-side schedule of the tensor-core contraction (plain C++, no CUDA): which pair rows each
// generator lane forms in each unit, which accumulator column every triangle reads, and how a
// list that does not fit one launch is cut into passes.  See contract_tc.cuh for the kernel.
//
// Vocabulary.  Rows of the field table are handled in aligned groups of 8.  A triangle
// (r1,r2,r3) -- the reference's np.sum(f[r1]*f[r2]*f[r3]), bskit/main.py:1875 -- is sorted
// a <= b <= c: the pair (a,b) is generated on the CUDA cores, c is the accumulator column.
// A "class" is a pair of groups (ga <= gb); one of the two groups "hosts" it: a generator warp
// that hosts group h keeps row 8h + (lane & 7) in registers and meets the partner group through
// shared memory.  A class is two "pieces" (warp x unit slots, 32 pair rows each, together all
// 64 ordered pairs of the two groups).  A pass has 8 warp slots (2 teams x 4 warps) x 4 units.
#pragma once

#include <algorithm>
#include <cstdint>
#include <map>
#include <set>
#include <vector>
#include <cstdio>
#include <cstdlib>

// BSK_TC_DEBUG=1 prints why a list was found not eligible
#define BSK_TCS_FAIL(msg) do { if (getenv("BSK_TC_DEBUG")) fprintf(stderr, "tc schedule: not eligible: %s (line %d)\n", msg, __LINE__); return false; } while (0)

namespace bsk {
namespace tcs {

constexpr int kTeams = 2, kUpt = 4, kUnits = 8, kWarpSlots = 8;
constexpr int kMaxColGroups = 5;     // 40 column rows per pass
constexpr int kMaxRawGroups = 15;    // 120 raw rows per pass
constexpr int kTeamCols = 96;        // accumulator columns per team (== tc::TEAMCOLS)
constexpr int kMaxUnitCols = 40;     // N of one MMA
constexpr int kCapTot = kTeams * kTeamCols;

// row of the partner group met by lane l of a piece (see contract_tc.cuh)
inline int partner_index(int lane, int piece) {
  const int k = lane >> 1;
  return ((k & 7) + 2 * piece + (k >> 3)) & 7;
}

struct Pass {
  int nraw = 0;                      // raw slots (multiple of 8)
  std::vector<int> rawrow;           // [nraw] field-table row or -1
  int ncols = 0;                     // window columns (multiple of 8)
  int colslot[kMaxColGroups] = {0, 0, 0, 0, 0};
  int nu[kTeams] = {0, 0};
  int col0[kUnits] = {0}, ncol[kUnits] = {0};
  int blk0[kUnits] = {0};            // first accumulator block (8 columns) of the unit within its team
  std::vector<uint32_t> lane_tab;    // [kUnits][128]: a_slot | b_slot << 8
  int64_t mma_cost = 0;              // sum over units of max(11, ncol/2): cycles per 8 cells and MMA term
  int pieces = 0;
};

struct Schedule {
  std::vector<Pass> passes;
  std::vector<int64_t> tri_slot;     // per triangle: pass * pass_stride + column * 128 + lane-in-unit
  int64_t pass_stride = 0;           // slots per pass and CTA (kCapTot * 128)
  double est_cycles = 0.0;           // estimated cycles per 32-cell chunk, summed over the passes
  bool cover = false;                // pairs chosen by the class cover (else: the two smallest rows)
};

namespace detail {

struct Lcg {
  uint64_t s;
  explicit Lcg(uint64_t seed) : s(seed * 2862933555777941757ull + 3037000493ull) {}
  uint32_t next() {
    s = s * 6364136223846793005ull + 1442695040888963407ull;
    return (uint32_t)(s >> 33);
  }
  int below(int n) { return (int)(next() % (uint32_t)n); }
};

struct Piece {
  int host = -1, part = -1, piece = 0;   // groups, piece index
  int lo = 1 << 30, hi = -1;             // window column range its needed pairs read (hi < lo: nothing)
};

// Column range (window positions) of the needed pairs one piece covers.  prange[(a,b)] is the
// range of pair a <= b; a pair of a diagonal class may be covered by two lanes -- both count here,
// the final assignment picks one.
inline void piece_range(Piece& pc, const std::map<std::pair<int, int>, std::pair<int, int>>& prange) {
  pc.lo = 1 << 30;
  pc.hi = -1;
  for (int l = 0; l < 32; ++l) {
    const int a = 8 * pc.host + (l & 7), b = 8 * pc.part + partner_index(l, pc.piece);
    auto it = prange.find({std::min(a, b), std::max(a, b)});
    if (it == prange.end()) continue;
    pc.lo = std::min(pc.lo, it->second.first);
    pc.hi = std::max(pc.hi, it->second.second);
  }
}

struct Placement {
  int slot_host[kWarpSlots];
  int pos[kWarpSlots][kUpt];     // piece index or -1
};

inline void unit_spans(const Placement& pl, const std::vector<Piece>& pcs, int lo[kUnits], int n[kUnits]) {
  for (int t = 0; t < kTeams; ++t)
    for (int j = 0; j < kUpt; ++j) {
      int l = 1 << 30, h = -1;
      for (int q = 0; q < 4; ++q) {
        const int pi = pl.pos[t * 4 + q][j];
        if (pi < 0 || pcs[pi].hi < 0) continue;
        l = std::min(l, pcs[pi].lo);
        h = std::max(h, pcs[pi].hi);
      }
      const int u = t * kUpt + j;
      if (h < 0) {
        bool any = false;
        for (int q = 0; q < 4; ++q) any |= pl.pos[t * 4 + q][j] >= 0;
        lo[u] = 0;
        n[u] = any ? 8 : 0;     // pieces without needed pairs still occupy a unit
      } else {
        lo[u] = l / 8 * 8;
        n[u] = h / 8 * 8 + 8 - lo[u];
      }
    }
}

// Estimated cycles per 32-cell chunk of a placement: the generator teams work through their units
// one after the other (kGenCycles each, whatever the number of pieces in the unit), the MMA
// issuers need 12 MMAs of max(11, N/2) cycles per unit; + a penalty when a unit is wider than one
// MMA or the units of a team need more than its accumulator columns.
constexpr double kGenCycles = 450.0, kReloadCycles = 50.0;
inline double placement_cost(const Placement& pl, const std::vector<Piece>& pcs, bool* feasible = nullptr) {
  int lo[kUnits], n[kUnits];
  unit_spans(pl, pcs, lo, n);
  double mma = 0.0, penalty = 0.0;
  int team_units[kTeams] = {0, 0};
  bool ok = true;
  for (int t = 0; t < kTeams; ++t) {
    int sum = 0;
    for (int j = 0; j < kUpt; ++j) {
      const int w = n[t * kUpt + j];
      if (w > 0) { mma += std::max(11.0, w / 2.0); ++team_units[t]; }
      sum += w;
      if (w > kMaxUnitCols) { ok = false; penalty += 1000.0 + 10.0 * (w - kMaxUnitCols); }
    }
    if (sum > kTeamCols) { ok = false; penalty += 1000.0 + 10.0 * (sum - kTeamCols); }
  }
  if (feasible) *feasible = ok;
  // a warp slot whose host group changes from one unit to the next reloads its resident row there
  // (8 more 128-bit loads for that warp): the unit takes about kReloadCycles longer
  double reload[kTeams] = {0.0, 0.0};
  for (int t = 0; t < kTeams; ++t)
    for (int j = 1; j < kUpt; ++j) {
      bool any = false;
      for (int q = 0; q < 4; ++q) {
        const int pi = pl.pos[t * 4 + q][j];
        if (pi < 0) continue;
        int prev = -1;
        for (int jj = j - 1; jj >= 0 && prev < 0; --jj)
          if (pl.pos[t * 4 + q][jj] >= 0) prev = pcs[pl.pos[t * 4 + q][jj]].host;
        any |= prev >= 0 && prev != pcs[pi].host;
      }
      if (any) reload[t] += kReloadCycles;
    }
  const double gen = std::max(kGenCycles * team_units[0] + reload[0], kGenCycles * team_units[1] + reload[1]);
  return std::max(gen, 12.0 * mma) + 0.05 * (gen + 12.0 * mma) + penalty;
}

// Place the classes of one pass.  Returns false when no feasible placement was found.
inline bool place_pass(const std::vector<std::pair<int, int>>& cls,
                       const std::map<std::pair<int, int>, std::pair<int, int>>& prange, uint64_t seed,
                       Placement& best, std::vector<Piece>& best_pcs, double* best_cost_out = nullptr) {
  if ((int)cls.size() * 2 > kWarpSlots * kUpt) return false;
  double best_cost = 1e300;
  bool found = false;
  Lcg rnd(seed);
  auto slot_host = [](const Placement& pl, const std::vector<Piece>& pcs, int w) {
    for (int j = 0; j < kUpt; ++j)
      if (pl.pos[w][j] >= 0) return pcs[pl.pos[w][j]].host;
    return -1;
  };
  for (int attempt = 0; attempt < 48 || (!found && attempt < 400); ++attempt) {
    // orientation: which group hosts each class (diagonal classes host themselves)
    std::vector<Piece> pcs;
    for (size_t i = 0; i < cls.size(); ++i) {
      const bool first = cls[i].first == cls[i].second || (rnd.next() & 1);
      for (int piece = 0; piece < 2; ++piece) {
        Piece pc;
        pc.host = first ? cls[i].first : cls[i].second;
        pc.part = first ? cls[i].second : cls[i].first;
        pc.piece = piece;
        piece_range(pc, prange);
        pcs.push_back(pc);
      }
    }
    // initial placement: pieces in order of their first column, each into the first free position of a
    // slot that already hosts its group, else of an empty slot (low units first)
    std::vector<int> order(pcs.size());
    for (size_t i = 0; i < order.size(); ++i) order[i] = (int)i;
    std::stable_sort(order.begin(), order.end(), [&](int x, int y) { return pcs[x].lo > pcs[y].lo; });
    Placement pl;
    for (int w = 0; w < kWarpSlots; ++w)
      for (int j = 0; j < kUpt; ++j) pl.pos[w][j] = -1;
    bool ok = true;
    for (int pi : order) {
      int bw = -1, bj = -1, bkey = 1 << 30;
      for (int w = 0; w < kWarpSlots; ++w) {
        const int h = slot_host(pl, pcs, w);
        for (int j = 0; j < kUpt; ++j)
          if (pl.pos[w][j] < 0) {
            const int key = j * 64 + (h == -1 ? 32 : (h != pcs[pi].host ? 48 : 0)) + (int)(rnd.next() & 15);
            if (key < bkey) { bkey = key; bw = w; bj = j; }
            break;
          }
      }
      if (bw < 0) { ok = false; break; }
      pl.pos[bw][bj] = pi;
    }
    if (!ok) continue;
    double cur = placement_cost(pl, pcs);
    for (int it = 0; it < 3000; ++it) {
      const int m = rnd.below(10);
      const int w1 = rnd.below(kWarpSlots), w2 = rnd.below(kWarpSlots), j1 = rnd.below(kUpt), j2 = rnd.below(kUpt);
      if (m < 3) {                       // permute two units of one warp slot
        if (j1 == j2) continue;
        std::swap(pl.pos[w1][j1], pl.pos[w1][j2]);
        const double c = placement_cost(pl, pcs);
        if (c <= cur) cur = c; else std::swap(pl.pos[w1][j1], pl.pos[w1][j2]);
      } else if (m < 8) {                // exchange two positions of different slots when the hosts allow it
        if (w1 == w2) continue;
        const int p1 = pl.pos[w1][j1], p2 = pl.pos[w2][j2];
        if (p1 < 0 && p2 < 0) continue;
        std::swap(pl.pos[w1][j1], pl.pos[w2][j2]);
        const double c = placement_cost(pl, pcs);
        if (c <= cur) cur = c; else std::swap(pl.pos[w1][j1], pl.pos[w2][j2]);
      } else {                           // exchange two whole warp slots (changes the teams)
        if (w1 == w2) continue;
        for (int j = 0; j < kUpt; ++j) std::swap(pl.pos[w1][j], pl.pos[w2][j]);
        const double c = placement_cost(pl, pcs);
        if (c <= cur) cur = c;
        else
          for (int j = 0; j < kUpt; ++j) std::swap(pl.pos[w1][j], pl.pos[w2][j]);
      }
    }
    bool feasible = false;
    cur = placement_cost(pl, pcs, &feasible);
    if (feasible && cur < best_cost) {
      best_cost = cur;
      best = pl;
      best_pcs = pcs;
      found = true;
    }
  }
  if (!found) return false;
  for (int w = 0; w < kWarpSlots; ++w) best.slot_host[w] = slot_host(best, best_pcs, w);
  // the units of a team in order of decreasing width, empty unit slots last
  int lo[kUnits], n[kUnits];
  unit_spans(best, best_pcs, lo, n);
  for (int t = 0; t < kTeams; ++t) {
    int order[kUpt] = {0, 1, 2, 3};
    std::stable_sort(order, order + kUpt, [&](int x, int y) { return n[t * kUpt + x] > n[t * kUpt + y]; });
    Placement tmp = best;
    for (int q = 0; q < 4; ++q)
      for (int j = 0; j < kUpt; ++j) tmp.pos[t * 4 + q][j] = best.pos[t * 4 + q][order[j]];
    best = tmp;
  }
  if (best_cost_out) *best_cost_out = best_cost;
  return true;
}

}  // namespace detail

// "Two halves" schedule for lists over <= 40 rows (one column window): the rows are cut into a low and
// a high half; every sorted triangle a <= b <= c has its two smallest rows in the low half (pair (a,b),
// column c, team 0) or its two largest in the high half (pair (b,c), column a, team 1) -- pigeonhole.
// Only pairs WITHIN a half are generated: 2 * 210 instead of 820 for 40 rows, which is fewer than any
// cover by 8-row group classes (436 needed rows on 640 lanes) and fits 2 + 2 units, so both teams work
// through two units per chunk instead of three and two.
//   A lane pair (2m, 2m+1) is a "slot": its lanes keep the two rows of a ROW PAIR (2t, 2t+1) in registers
// for the whole chunk and meet ONE partner row per unit, which both lanes read (2-cycle shared-memory
// loads).  A slot use (t, y) covers the pairs (2t, y) and (2t+1, y).  Row pairs t != u are related through
// up to two uses on either side; the side with the lighter load takes them.
inline bool build_schedule_halves(int ntri, const int32_t* rows, int nrows, Schedule& out) {
  using namespace detail;
  if (ntri < 256 || nrows < 4 || nrows > 8 * kMaxColGroups) return false;
  const int nraw = (nrows + 7) / 8 * 8;
  int h = (nraw / 2) & ~1;                    // rows [0, h) | [h, nraw)
  out.passes.clear();
  out.est_cycles = 0.0;
  out.cover = true;
  out.tri_slot.assign((size_t)ntri, -1);
  out.pass_stride = (int64_t)kCapTot * 128;
  std::vector<int> ta((size_t)ntri), tb((size_t)ntri), tcol((size_t)ntri), tteam((size_t)ntri);
  std::set<std::pair<int, int>> need[kTeams];
  for (int t = 0; t < ntri; ++t) {
    int r[3] = {rows[3 * t], rows[3 * t + 1], rows[3 * t + 2]};
    std::sort(r, r + 3);
    if (r[0] < 0 || r[2] >= nrows) return false;
    if (r[1] < h) { ta[t] = r[0]; tb[t] = r[1]; tcol[t] = r[2]; tteam[t] = 0; }
    else { ta[t] = r[1]; tb[t] = r[2]; tcol[t] = r[0]; tteam[t] = 1; }
    need[tteam[t]].insert({ta[t], tb[t]});
  }
  Pass ps;
  ps.nraw = nraw;
  ps.rawrow.assign((size_t)nraw, -1);
  for (int i = 0; i < nrows; ++i) ps.rawrow[i] = i;
  ps.ncols = nraw;
  for (int g = 0; g < nraw / 8; ++g) ps.colslot[g] = 8 * g;
  const uint32_t zero_row = (uint32_t)nraw;
  ps.lane_tab.assign((size_t)kUnits * 128, zero_row | (zero_row << 8));
  std::map<std::pair<int, int>, std::pair<int, int>> where[kTeams];     // pair -> (unit, lane-in-unit)
  for (int team = 0; team < kTeams; ++team) {
    if (need[team].empty()) continue;
    // uses[t] = partner rows the slots of row pair t have to meet
    std::map<int, std::vector<int>> uses;
    std::map<std::pair<int, int>, std::set<int>> rel_rows_lo, rel_rows_hi;   // relation (t<u): rows of t / of u involved
    std::map<int, std::set<int>> within;
    for (auto& pr : need[team]) {
      const int t = pr.first / 2, u = pr.second / 2;
      if (t == u) { within[t].insert(pr.first); within[t].insert(pr.second); if (pr.first != pr.second) within[t].insert(-1); }
      else { rel_rows_lo[{t, u}].insert(pr.first); rel_rows_hi[{t, u}].insert(pr.second); }
    }
    std::map<int, int> load;
    for (auto& kv : within) {
      // pairs (2t,2t), (2t,2t+1), (2t+1,2t+1): partner 2t covers the first two, partner 2t+1 the last two
      const int t = kv.first;
      bool need00 = false, need01 = kv.second.count(-1) > 0, need11 = false;
      for (auto& pr : need[team]) {
        if (pr.first == 2 * t && pr.second == 2 * t) need00 = true;
        if (pr.first == 2 * t + 1 && pr.second == 2 * t + 1) need11 = true;
      }
      if (need00 || (need01 && !need11)) uses[t].push_back(2 * t);
      if (need11) uses[t].push_back(2 * t + 1);
      load[t] = (int)uses[t].size();
    }
    // relations between different row pairs, heaviest first, to the lighter side
    std::vector<std::pair<int, int>> rels;
    for (auto& kv : rel_rows_lo) rels.push_back(kv.first);
    std::stable_sort(rels.begin(), rels.end(), [&](const std::pair<int, int>& x, const std::pair<int, int>& y) {
      return rel_rows_lo[x].size() + rel_rows_hi[x].size() > rel_rows_lo[y].size() + rel_rows_hi[y].size();
    });
    for (auto& r : rels) {
      // own = t: one use per involved row of u (and vice versa)
      const int cost_t = (int)rel_rows_hi[r].size(), cost_u = (int)rel_rows_lo[r].size();
      const bool to_t = load[r.first] + cost_t < load[r.second] + cost_u ||
                        (load[r.first] + cost_t == load[r.second] + cost_u && cost_t <= cost_u);
      if (to_t) { for (int y : rel_rows_hi[r]) uses[r.first].push_back(y); load[r.first] += cost_t; }
      else { for (int x : rel_rows_lo[r]) uses[r.second].push_back(x); load[r.second] += cost_u; }
    }
    // smallest number of units whose 64 slots hold every row pair's uses
    int U = 0;
    for (int cand = 1; cand <= kUpt && U == 0; ++cand) {
      int slots = 0;
      for (auto& kv : uses) slots += ((int)kv.second.size() + cand - 1) / cand;
      if (slots <= 64 && cand * kMaxUnitCols <= kTeamCols + (cand == 2 ? 16 : 0)) U = cand;
    }
    if (U == 0 || U * kMaxUnitCols > kTeamCols + 16) BSK_TCS_FAIL("two halves: the pairs of a half do not fit a team");
    ps.nu[team] = U;
    // slots: (row pair, partner row per unit); unit j takes the j-th block of the row pair's sorted
    // partner list, so neighbouring columns share a unit
    struct Slot { int t; int partner[kUpt]; };
    std::vector<Slot> slots;
    for (auto& kv : uses) {
      std::vector<int> list = kv.second;
      std::sort(list.begin(), list.end());
      const int nslots = ((int)list.size() + U - 1) / U;
      for (int k = 0; k < nslots; ++k) {
        Slot sl;
        sl.t = kv.first;
        for (int j = 0; j < kUpt; ++j) sl.partner[j] = (j < U && j * nslots + k < (int)list.size()) ? list[j * nslots + k] : -1;
        slots.push_back(sl);
      }
    }
    // (slots stay in row-pair order: the lanes of a warp then keep few distinct resident rows and read
    // neighbouring partner rows; a search that spread the partner rows of a warp over the bank residues
    // was slower, 32.5 against 31.0 ms)
    while ((int)slots.size() < 64) { Slot sl; sl.t = -1; for (int j = 0; j < kUpt; ++j) sl.partner[j] = -1; slots.push_back(sl); }
    int used = 0;
    for (int slot = 0; slot < 64; ++slot) {
      const Slot& sl = slots[slot];
      if (sl.t < 0) continue;
      ++used;
      for (int j = 0; j < U; ++j) {
        const int u = team * kUpt + j;
        for (int e = 0; e < 2; ++e) {
          const int lane = slot * 2 + e, own = 2 * sl.t + e;
          const uint32_t a_slot = own < nrows ? (uint32_t)own : zero_row;
          const uint32_t b_slot = sl.partner[j] >= 0 ? (uint32_t)sl.partner[j] : zero_row;
          ps.lane_tab[(size_t)u * 128 + lane] = a_slot | (b_slot << 8);    // idle uses keep the resident row
          if (sl.partner[j] >= 0 && own < nrows) {
            const auto key = std::make_pair(std::min(own, sl.partner[j]), std::max(own, sl.partner[j]));
            if (!where[team].count(key)) where[team][key] = {u, lane};
          }
        }
      }
    }
    ps.pieces += U * ((used + 15) / 16);
  }
  // unit column ranges from the triangles that really read them
  int ulo[kUnits], uhi[kUnits];
  for (int u = 0; u < kUnits; ++u) { ulo[u] = 1 << 30; uhi[u] = -1; }
  for (int t = 0; t < ntri; ++t) {
    auto wh = where[tteam[t]].find({ta[t], tb[t]});
    if (wh == where[tteam[t]].end()) BSK_TCS_FAIL("two halves: pair without a lane");
    ulo[wh->second.first] = std::min(ulo[wh->second.first], tcol[t]);
    uhi[wh->second.first] = std::max(uhi[wh->second.first], tcol[t]);
  }
  for (int u = 0; u < kUnits; ++u) {
    const bool live = (u % kUpt) < ps.nu[u / kUpt];
    if (uhi[u] < 0) { ps.col0[u] = 0; ps.ncol[u] = live ? 8 : 0; }
    else { ps.col0[u] = ulo[u] / 8 * 8; ps.ncol[u] = uhi[u] / 8 * 8 + 8 - ps.col0[u]; }
    if (ps.ncol[u] > kMaxUnitCols) BSK_TCS_FAIL("two halves: unit wider than one MMA");
    if (ps.ncol[u] > 0) ps.mma_cost += std::max(11, ps.ncol[u] / 2);
  }
  int tu[kTeams] = {0, 0};
  for (int t = 0; t < kTeams; ++t) {
    int b = 0;
    for (int j = 0; j < kUpt; ++j) { ps.blk0[t * kUpt + j] = b; b += ps.ncol[t * kUpt + j] / 8; tu[t] += ps.ncol[t * kUpt + j] > 0; }
    if (b * 8 > kTeamCols) BSK_TCS_FAIL("two halves: team needs more accumulator columns than it has");
  }
  for (int t = 0; t < ntri; ++t) {
    const auto wh = where[tteam[t]][{ta[t], tb[t]}];
    const int u = wh.first;
    out.tri_slot[t] = (int64_t)((u / kUpt) * kTeamCols + ps.blk0[u] * 8 + tcol[t] - ps.col0[u]) * 128 + wh.second;
  }
  out.est_cycles = std::max(kGenCycles * std::max(tu[0], tu[1]), 12.0 * (double)ps.mma_cost) + 200.0;
  out.passes.push_back(std::move(ps));
  return true;
}

// Build the schedule; returns false when the list is not eligible (the caller then keeps the
// FP32-pipe kernel).
inline bool build_schedule_mode(int ntri, const int32_t* rows, int nrows, bool cover, Schedule& out) {
  using namespace detail;
  if (ntri < 256 || nrows < 1 || nrows > 8 * 64) return false;
  out.passes.clear();
  out.est_cycles = 0.0;
  out.cover = cover;
  out.tri_slot.assign((size_t)ntri, -1);
  out.pass_stride = (int64_t)kCapTot * 128;
  // Which two rows of a triangle form the generated pair is free (the product commutes): pick a
  // small set of classes that covers every triangle (greedy set cover -- the all-triangle list of
  // S = 40 bins needs 9 of its 15 classes, i.e. 40 % fewer pair rows to generate), then give each
  // triangle the covering class whose column range grows least.
  std::vector<int> ta((size_t)ntri), tb((size_t)ntri), tc_((size_t)ntri);
  std::vector<char> cgroup_used((size_t)(nrows + 7) / 8, 0);
  {
    struct Opt { int a, b, c; };
    std::vector<Opt> opts((size_t)ntri * 3);
    for (int t = 0; t < ntri; ++t) {
      int r[3] = {rows[3 * t], rows[3 * t + 1], rows[3 * t + 2]};
      std::sort(r, r + 3);
      if (r[0] < 0 || r[2] >= nrows) return false;
      opts[3 * t + 0] = {r[0], r[1], r[2]};
      opts[3 * t + 1] = {r[0], r[2], r[1]};
      opts[3 * t + 2] = {r[1], r[2], r[0]};
    }
    auto cls_of = [](const Opt& o) { return std::make_pair(o.a / 8, o.b / 8); };
    std::set<std::pair<int, int>> chosen;
    std::vector<char> covered((size_t)ntri, 0);
    int left = cover ? ntri : 0;
    while (left > 0) {
      std::map<std::pair<int, int>, int> gain;
      for (int t = 0; t < ntri; ++t) {
        if (covered[t]) continue;
        std::pair<int, int> seen[3];
        int ns = 0;
        for (int k = 0; k < 3; ++k) {
          const auto c = cls_of(opts[3 * t + k]);
          bool dup = false;
          for (int i = 0; i < ns; ++i) dup |= seen[i] == c;
          if (!dup) { seen[ns++] = c; ++gain[c]; }
        }
      }
      std::pair<int, int> best{-1, -1};
      int best_gain = -1;
      for (auto& kv : gain)
        if (kv.second > best_gain) { best_gain = kv.second; best = kv.first; }
      chosen.insert(best);
      for (int t = 0; t < ntri; ++t) {
        if (covered[t]) continue;
        for (int k = 0; k < 3; ++k)
          if (cls_of(opts[3 * t + k]) == best) { covered[t] = 1; --left; break; }
      }
    }
    std::map<std::pair<int, int>, std::pair<int, int>> span;   // class -> [lo, hi] of its columns (rows)
    for (int t = 0; t < ntri; ++t) {
      int pick = cover ? -1 : 0, pick_grow = 1 << 30;
      for (int k = 0; k < 3 && cover; ++k) {
        const Opt& o = opts[3 * t + k];
        if (!chosen.count(cls_of(o))) continue;
        auto it = span.find(cls_of(o));
        const int grow = it == span.end() ? 8 : std::max(0, it->second.first - o.c) + std::max(0, o.c - it->second.second);
        if (grow < pick_grow) { pick_grow = grow; pick = k; }
      }
      const Opt& o = opts[3 * t + pick];
      ta[t] = o.a; tb[t] = o.b; tc_[t] = o.c;
      auto it = span.find(cls_of(o));
      if (it == span.end()) span[cls_of(o)] = {o.c, o.c};
      else { it->second.first = std::min(it->second.first, o.c); it->second.second = std::max(it->second.second, o.c); }
      cgroup_used[o.c / 8] = 1;
    }
  }
  // column windows: consecutive runs of <= 5 used column groups
  std::vector<std::vector<int>> windows;
  for (int g = 0; g < (int)cgroup_used.size(); ++g) {
    if (!cgroup_used[g]) continue;
    if (windows.empty() || (int)windows.back().size() == kMaxColGroups) windows.push_back({});
    windows.back().push_back(g);
  }
  for (const auto& win : windows) {
    std::map<int, int> colpos_of_group;
    for (size_t i = 0; i < win.size(); ++i) colpos_of_group[win[i]] = (int)i * 8;
    // pairs (and their classes) with a triangle in this window
    std::map<std::pair<int, int>, std::pair<int, int>> prange;     // pair -> window column range
    std::map<std::pair<int, int>, int> class_lo;                   // class -> lowest column
    std::vector<int> tris;
    for (int t = 0; t < ntri; ++t) {
      auto it = colpos_of_group.find(tc_[t] / 8);
      if (it == colpos_of_group.end()) continue;
      tris.push_back(t);
      const int col = it->second + tc_[t] % 8;
      auto key = std::make_pair(ta[t], tb[t]);
      auto pr = prange.find(key);
      if (pr == prange.end()) prange[key] = {col, col};
      else { pr->second.first = std::min(pr->second.first, col); pr->second.second = std::max(pr->second.second, col); }
      auto ck = std::make_pair(ta[t] / 8, tb[t] / 8);
      auto cl = class_lo.find(ck);
      if (cl == class_lo.end()) class_lo[ck] = col; else cl->second = std::min(cl->second, col);
    }
    // classes ordered by their lowest column (descending): neighbours share column ranges
    std::vector<std::pair<int, int>> classes;
    for (auto& kv : class_lo) classes.push_back(kv.first);
    std::stable_sort(classes.begin(), classes.end(), [&](const std::pair<int, int>& x, const std::pair<int, int>& y) {
      return class_lo[x] > class_lo[y];
    });
    size_t next = 0;
    while (next < classes.size()) {
      size_t take = std::min<size_t>(16, classes.size() - next);
      Placement pl;
      std::vector<Piece> pcs;
      std::vector<std::pair<int, int>> cls;
      for (;; --take) {
        if (take == 0) BSK_TCS_FAIL("no feasible placement for a pass");
        cls.assign(classes.begin() + next, classes.begin() + next + take);
        // raw groups this pass touches
        std::vector<int> groups(win.begin(), win.end());
        for (auto& c : cls) { groups.push_back(c.first); groups.push_back(c.second); }
        std::sort(groups.begin(), groups.end());
        groups.erase(std::unique(groups.begin(), groups.end()), groups.end());
        if ((int)groups.size() > kMaxRawGroups) continue;
        if (place_pass(cls, prange, 12345u + out.passes.size() * 977u + take, pl, pcs)) break;
      }
      // ---- materialise the pass
      Pass ps;
      std::vector<int> groups(win.begin(), win.end());
      for (auto& c : cls) { groups.push_back(c.first); groups.push_back(c.second); }
      std::sort(groups.begin(), groups.end());
      groups.erase(std::unique(groups.begin(), groups.end()), groups.end());
      std::map<int, int> slot_of_group;
      for (size_t i = 0; i < groups.size(); ++i) slot_of_group[groups[i]] = (int)i * 8;
      ps.nraw = (int)groups.size() * 8;
      ps.rawrow.assign((size_t)ps.nraw, -1);
      for (size_t i = 0; i < groups.size(); ++i)
        for (int k = 0; k < 8; ++k)
          if (groups[i] * 8 + k < nrows) ps.rawrow[i * 8 + k] = groups[i] * 8 + k;
      ps.ncols = (int)win.size() * 8;
      for (size_t i = 0; i < win.size(); ++i) ps.colslot[i] = slot_of_group[win[i]];
      const uint32_t idle = (uint32_t)ps.nraw | ((uint32_t)ps.nraw << 8);
      ps.lane_tab.assign((size_t)kUnits * 128, idle);
      // where every ordered pair of this pass lives: (unit, lane-in-unit)
      std::map<std::pair<int, int>, std::pair<int, int>> where;
      std::set<std::pair<int, int>> cls_set(cls.begin(), cls.end());
      for (int w = 0; w < kWarpSlots; ++w)
        for (int j = 0; j < kUpt; ++j) {
          const int pi = pl.pos[w][j];
          if (pi < 0) continue;
          const Piece& pc = pcs[pi];
          const int u = (w / 4) * kUpt + j, q = w % 4;
          ps.nu[w / 4] = std::max(ps.nu[w / 4], j + 1);
          ++ps.pieces;
          for (int l = 0; l < 32; ++l) {
            const int a = 8 * pc.host + (l & 7), b = 8 * pc.part + partner_index(l, pc.piece);
            ps.lane_tab[(size_t)u * 128 + q * 32 + l] =
                (uint32_t)(slot_of_group[pc.host] + (l & 7)) | ((uint32_t)(slot_of_group[pc.part] + b % 8) << 8);
            auto key = std::make_pair(std::min(a, b), std::max(a, b));
            if (!where.count(key)) where[key] = {u, q * 32 + l};
          }
        }
      // unit column ranges from the triangles that really read them
      int ulo[kUnits], uhi[kUnits];
      for (int u = 0; u < kUnits; ++u) { ulo[u] = 1 << 30; uhi[u] = -1; }
      std::vector<int> mine;
      for (int t : tris) {
        if (!cls_set.count({ta[t] / 8, tb[t] / 8})) continue;
        auto wh = where.find({ta[t], tb[t]});
        if (wh == where.end()) BSK_TCS_FAIL("pair without a lane");
        const int col = colpos_of_group[tc_[t] / 8] + tc_[t] % 8;
        ulo[wh->second.first] = std::min(ulo[wh->second.first], col);
        uhi[wh->second.first] = std::max(uhi[wh->second.first], col);
        mine.push_back(t);
      }
      for (int u = 0; u < kUnits; ++u) {
        const bool live = (u % kUpt) < ps.nu[u / kUpt];
        if (uhi[u] < 0) { ps.col0[u] = 0; ps.ncol[u] = live ? 8 : 0; }
        else { ps.col0[u] = ulo[u] / 8 * 8; ps.ncol[u] = uhi[u] / 8 * 8 + 8 - ps.col0[u]; }
        if (ps.ncol[u] > kMaxUnitCols) BSK_TCS_FAIL("unit wider than one MMA");
        if (ps.ncol[u] > 0) ps.mma_cost += std::max(11, ps.ncol[u] / 2);
      }
      for (int t = 0; t < kTeams; ++t) {
        int b = 0;
        for (int j = 0; j < kUpt; ++j) { ps.blk0[t * kUpt + j] = b; b += ps.ncol[t * kUpt + j] / 8; }
        if (b * 8 > kTeamCols) {
          if (getenv("BSK_TC_DEBUG")) {
            for (int u = 0; u < kUnits; ++u) fprintf(stderr, " unit %d: col0 %d ncol %d\n", u, ps.col0[u], ps.ncol[u]);
            int lo[kUnits], n[kUnits];
            unit_spans(pl, pcs, lo, n);
            for (int u = 0; u < kUnits; ++u) fprintf(stderr, " placement unit %d: lo %d n %d\n", u, lo[u], n[u]);
          }
          BSK_TCS_FAIL("team needs more accumulator columns than it has");
        }
      }
      const int64_t base = (int64_t)out.passes.size() * out.pass_stride;
      for (int t : mine) {
        const auto wh = where[{ta[t], tb[t]}];
        const int col = colpos_of_group[tc_[t] / 8] + tc_[t] % 8;
        if (out.tri_slot[t] >= 0) BSK_TCS_FAIL("triangle in two passes");
        const int u = wh.first;
        out.tri_slot[t] = base + (int64_t)((u / kUpt) * kTeamCols + ps.blk0[u] * 8 + col - ps.col0[u]) * 128 + wh.second;
      }
      {
        int tu[kTeams] = {0, 0};
        for (int u = 0; u < kUnits; ++u) tu[u / kUpt] += ps.ncol[u] > 0;
        out.est_cycles += std::max(kGenCycles * std::max(tu[0], tu[1]), 12.0 * (double)ps.mma_cost) + 200.0;
      }
      out.passes.push_back(std::move(ps));
      next += take;
    }
  }
  for (int t = 0; t < ntri; ++t)
    if (out.tri_slot[t] < 0) BSK_TCS_FAIL("triangle without a slot");
  return !out.passes.empty();
}

// Build the schedule; returns false when the list is not eligible (the caller then keeps the
// FP32-pipe kernel).  Two ways of choosing the generated pair are tried -- the two smallest rows
// of every triangle, and the greedy class cover -- and the one with the lower estimated time wins.
inline bool build_schedule(int ntri, const int32_t* rows, int nrows, Schedule& out) {
  Schedule a, b;
  const bool oka = build_schedule_mode(ntri, rows, nrows, false, a);
  const bool okb = build_schedule_mode(ntri, rows, nrows, true, b);
  if (!oka && !okb) return false;
  bool pick_b = okb && (!oka || b.est_cycles < a.est_cycles);
  if (const char* force = getenv("BSK_TC_COVER")) {       // A/B knob: 0 = two smallest rows, 1 = class cover
    if (force[0] == '0' && oka) pick_b = false;
    if (force[0] == '1' && okb) pick_b = true;
  }
  out = pick_b ? std::move(b) : std::move(a);
  // third candidate: the two-halves schedule (lists over <= 40 rows)
  Schedule c;
  const char* fh = getenv("BSK_TC_HALVES");              // A/B knob: 0 = never, 1 = whenever it is feasible
  if (!(fh && fh[0] == '0') && build_schedule_halves(ntri, rows, nrows, c) &&
      ((fh && fh[0] == '1') || c.est_cycles < out.est_cycles))
    out = std::move(c);
  return true;
}

}  // namespace tcs
}  // namespace bsk
