// rows5.cu -- forward / backward banded DP rows, scaled linear domain (dp3.cuh), ROTATING wavefront.
//
// Same rows as rows4.cu (Node::NextRow, node_next_row.h:6-61, driven by dtw.cpp:48-81,182-197), different schedule.
// rows4.cu cuts the rows of a pass into stripes of 31 (A-row, B-row) pairs; inside a stripe the band drifts by ~10
// columns per pair while the wavefront lags by one column per lane, so every lane idles for half of the stripe
// (utilisation W / (W + 11*31) = 47 % at the default band).  Here ONE warp per (read, direction) runs one continuous
// wavefront: pair g always works on column C0 +- (t - g) at step t, and lane g mod 32 takes pair g+32 as soon as
// pair g has left its band -- nominally a pair is busy W+1 = 302 of every 352 steps (86 %).  Neighbour cells still
// travel by warp shuffle (a rotation, lane 31 feeds lane 0), tagged with the pair index because a lane may already
// have moved on.
//
// When a pair is not finished by the time its lane must start the next one (bands are irregular: ~5 % of the pairs
// of a default read) the new pair starts in a SECOND slot of the lane, and the step body runs a second time, for the
// whole warp, only while some lane has a live second slot.  Reads that would need a third slot (very wide or very
// slowly drifting bands) are recognised by the band kernel (NVB_READ_NO_ROTATION) and go through rows4.cu.
// Pairs change slots only at multiples of TS steps, so the store tiles (see rows4.cu) always hold one pair per lane.
#include "dp3.cuh"
#include "kernels.h"

namespace {

#ifndef NVB_ROT_MIN_BLOCKS
#define NVB_ROT_MIN_BLOCKS 8  // 64-thread CTAs per SM: 128 registers per thread, 16 warps per SM
#endif

// Steps per store tile; pairs are (de)activated only at multiples of it.  The per-tile work (slot management, ring
// upkeep, the transposing flush: ~300 warp instructions) is amortised over the tile; a coarser tile keeps pairs in
// their slots longer, so the second slot -- which makes the whole warp run the step body twice -- is live a little
// more often (3.6 % of the 8-step tiles against 2.4 % of the 4-step tiles on the bench reads).  Measured at 1000 reads:
// 8 steps win for the plain sweep (18.4 -> 16.6 ms, profiles/r02e) and, since the round-2 emission code, for the wobble
// sweep too (25.3 against 26.9 ms, profiles/r02w; with the round-1 emission 4 steps were better, 25.6 against 26.7).
// 8 is the largest value the slot-reuse guarantee of band.cu (be[j] - bs[j+63] <= 48) allows.
#ifndef NVB_ROT_TS_PLAIN
#define NVB_ROT_TS_PLAIN 8
#endif
#ifndef NVB_ROT_TS_OTHER
#define NVB_ROT_TS_OTHER 8
#endif
__host__ __device__ constexpr int tile_steps(int mode) { return mode == NVB_MODE_PLAIN ? NVB_ROT_TS_PLAIN : NVB_ROT_TS_OTHER; }

// What the transposing flush needs to know about the row a lane parked in a store tile, reduced to the step index: the
// cell of step tt lies at plane index base + tt (forward) / base - tt (reverse) and is inside the row's band for
// t_lo <= tt <= t_hi (an empty range for lanes that store nothing).  16 bytes: one shared-memory load per flushed cell.
struct RowMeta {
  long long base;
  int t_lo, t_hi;
};
struct StoreTile {
  double *f;      // [32][TS + 1]
  int32_t *x;     // [32][TS + 1]
  RowMeta *meta;  // [32]
};
__host__ __device__ constexpr size_t tile_bytes(int ts) {
  return 32 * (ts + 1) * (sizeof(double) + sizeof(int32_t)) + 32 * sizeof(RowMeta);
}
// tiles per warp: the B rows of the two slots, plus their A rows in the transition sweep (the only one that stores them)
__host__ __device__ constexpr int tiles_per_warp(int mode) { return mode == NVB_MODE_TRANS ? 4 : 2; }
// Parameters of the NEXT row pair of a lane, staged in shared memory by cp.async while the lane's current pair is
// still running (~350 steps ahead): band rows i and i+1, their cell offsets and the emission row of position i.  A lane
// that takes a new pair reads them back with shared-memory latency; fetched at that moment from global memory they
// left the whole warp waiting (long-scoreboard stalls were 29 % of the stall cycles of the sweep, ncu r02z).
struct PairRec {
  int bs0, bs1, be0, be1;
  long long coff0, coff1;
  double mu, ac, mc, flags;  // BatchDev::row_emis[i]
};
static_assert(sizeof(PairRec) == 64, "PairRec layout");
static_assert((2 * tile_bytes(NVB_ROT_TS_PLAIN)) % 256 == 0 && (2 * tile_bytes(NVB_ROT_TS_OTHER)) % 256 == 0,
              "the signal ring behind the store tiles must start on a 256-byte boundary (TMA bulk copies)");
// per warp: store tiles | signal ring | 32 PairRec
__host__ __device__ constexpr size_t rec_offset(int mode) {
  return ((tiles_per_warp(mode) * tile_bytes(tile_steps(mode)) + ring_bytes(8) + 15) / 16) * 16;
}
__host__ __device__ constexpr size_t warp_bytes(int mode) {  // keeps every ring 256-byte aligned
  return ((rec_offset(mode) + NVB_WARP * sizeof(PairRec) + 255) / 256) * 256;
}

__device__ __forceinline__ void rec_prefetch(PairRec *rec, const ReadView &v, int i) {
  const unsigned a = smem_u32(rec);
  const int32_t *bs = v.bs + i, *be = v.be + i;
  const int64_t *co = v.coff + i;
  const double *em = v.emis + 4 * (size_t)i;
  asm volatile(
      "cp.async.ca.shared.global [%0], [%1], 4;\n"
      "cp.async.ca.shared.global [%0 + 4], [%1 + 4], 4;\n"
      "cp.async.ca.shared.global [%0 + 8], [%2], 4;\n"
      "cp.async.ca.shared.global [%0 + 12], [%2 + 4], 4;\n"
      "cp.async.ca.shared.global [%0 + 16], [%3], 8;\n"
      "cp.async.ca.shared.global [%0 + 24], [%3 + 8], 8;\n"
      "cp.async.ca.shared.global [%0 + 32], [%4], 16;\n"
      "cp.async.ca.shared.global [%0 + 48], [%4 + 16], 16;\n"
      "cp.async.commit_group;\n" ::"r"(a), "l"(bs), "l"(be), "l"(co), "l"(em)
      : "memory");
}

template <bool REV, int TS>
__device__ __forceinline__ void flush_tile(const StoreTile &tile, double *F, int32_t *X, int C0, int t0, int lane) {
  constexpr int TSTRIDE = TS + 1;     // padded tile row stride
  constexpr int RPI = NVB_WARP / TS;  // rows per iteration
  const int kk = lane & (TS - 1);
#pragma unroll
  for (int i = 0; i < NVB_WARP / RPI; i++) {
    const int r = RPI * i + lane / TS;
    const RowMeta m = tile.meta[r];
    const int tt = t0 + kk;
    if (tt >= m.t_lo && tt <= m.t_hi) {
      const long long idx = REV ? m.base - tt : m.base + tt;
      F[idx] = tile.f[r * TSTRIDE + kk];
      X[idx] = tile.x[r * TSTRIDE + kk];
    }
  }
}

// One (A-row, B-row) pair in flight in a lane.
template <int MEL>
struct Slot {
  LaneCfg L;
  LaneState<MEL> S;
  LaneOut out;      // outputs of the last step (consumed by pair + 1 at the next step)
  XD aout;
  int pair;         // pair index g in sweep order, -1 = free
  int t_end;        // last step with an in-band column
  long long boff, aoff;
  int ws, awe;      // A-row band for the stores of the transition sweep (empty when nothing is stored)
  double p_cur;     // emission of the NEXT step to run, evaluated one step ahead (lane_emit): p_cur * 2^k_cur
  int k_cur;
};

template <int MEL>
__device__ __forceinline__ void slot_clear(Slot<MEL> &Q) {
  lane_cfg_clear(Q.L);
  lane_reset(Q.S);
  Q.out.f = 0.0; Q.out.E = NVB_EZERO; Q.out.p = 1.0; Q.out.k = 0;
  Q.aout = xd_zero();
  Q.pair = -1; Q.t_end = -1; Q.boff = 0; Q.aoff = 0; Q.ws = 1; Q.awe = 0; Q.p_cur = 1.0; Q.k_cur = 0;
}

template <bool REV>
__device__ __forceinline__ int pair_base(int n, int g) { return REV ? n - 1 - g : g; }

// Steps at which pair g has its first / last in-band column (C0 = near end of the initial row's band).
template <bool REV>
__device__ __forceinline__ int pair_t_start(const ReadView &v, int C0, int g) {
  const int i = pair_base<REV>(v.n, g);
  return REV ? (C0 - v.be[i + 1]) + g : (v.bs[i] - C0) + g;
}
template <bool REV>
__device__ __forceinline__ int pair_t_end(const ReadView &v, int C0, int g) {
  const int i = pair_base<REV>(v.n, g);
  return REV ? (C0 - v.bs[i]) + g : (v.be[i + 1] - C0) + g;
}

template <bool REV>
__device__ __forceinline__ int sample_index(const ReadView &v, int C0, int g, int t) {
  const int c = REV ? C0 - (t - g) : C0 + (t - g);
  return min(max(REV ? c : c - 1, 0), v.N - 1);
}

template <int MEL, int MODE, bool REV>
__device__ __forceinline__ void slot_start(Slot<MEL> &Q, const ReadView &v, const SignalRing<8> &R, unsigned exp_tab,
                                           const PairRec *rec, long long coff1, int C0, int g, int t) {
  const int n = v.n;
  const int i = pair_base<REV>(n, g);
  const double C_E2 = 0.1353352832366127;  // exp(-2): the "/ 2" of kmer_model.cpp:60 is "- 2.0" in log space
  asm volatile("cp.async.wait_group 0;" ::: "memory");  // this lane's record of pair g (rec_prefetch) has landed
  const PairRec r = *rec;
  slot_clear(Q);
  Q.pair = g;
  Q.t_end = REV ? (C0 - r.bs0) + g : (r.be1 - C0) + g;  // pair_t_end
  const bool hasA = (MODE != NVB_MODE_PLAIN) && (REV ? (i <= n - 2) : (i >= 1));
  // A-row on band row (REV ? i + 1 : i), B-row on band row (REV ? i : i + 1)
  Q.L.role = NVB_ROLE_PAIR;
  Q.L.mu = r.mu; Q.L.ac = r.ac; Q.L.mc = r.mc;
  Q.L.ms = REV ? r.bs0 : r.bs1; Q.L.me = REV ? r.be0 : r.be1;
  if (MODE == NVB_MODE_TRANS) {  // trans_row_off(): rows 2i / 2i+1 live on band rows i / i+1
    const long long even = r.coff0 + r.coff1 - coff1, odd = 2 * r.coff1 - coff1;
    Q.aoff = REV ? odd : even;
    Q.boff = REV ? even : odd;
  } else {
    Q.boff = REV ? r.coff0 : r.coff1;
  }
  if (hasA) {
    Q.L.ws = REV ? r.bs1 : r.bs0; Q.L.we = REV ? r.be1 : r.be0;
    if (MODE == NVB_MODE_TRANS) {  // GetTransitionDistribution (kmer_model.cpp:64-94): constant 0.01, or 0
      Q.ws = Q.L.ws; Q.awe = Q.L.we;
      const bool dead = ((int)r.flags >> (REV ? 1 : 0)) & 1;  // same mean as the row on the other side of the transition
      Q.L.pc = dead ? 0.0 : 0.01 * 64.0;  // 0.01 as mantissa 0.64 and exponent -6 (an exact rescaling)
      Q.L.kc = dead ? NVB_EZERO : -6;
    } else {
      Q.L.cm = C_E2; Q.L.abias = 0;
    }
  }
  lane_emit(Q.L, exp_tab, ring_read(R, sample_index<REV>(v, C0, g, t)), Q.p_cur, Q.k_cur);  // emission of its first step
}

// One step of one slot.  `pf .. pk, ptag` = outputs of the producer lane's slot(s) at the previous step.
template <int MEL, int MODE, bool REV>
__device__ __forceinline__ void slot_step(Slot<MEL> &Q, const ReadView &v, const SignalRing<8> &R, unsigned exp_tab,
                                          int C0, int t,
                                          const LaneOut &in0, int tag0, const LaneOut &in1, int tag1, bool have1,
                                          int ones_s, int ones_e) {
  // Runs for EVERY lane, also those whose slot is free (pair -1: cleared configuration, empty bands, zero state, so
  // the step computes zeros that nobody reads): a per-lane `if (pair >= 0)` around ~150 instructions costs a
  // BSSY / BRA / BSYNC triple and a reconvergence stall at every step (ncu r02o: 13 % of the stall samples of the
  // sweep sat on control-flow instructions).  For the same reason the inflow is chosen by selects, not branches.
  const int g = Q.pair;
  const bool live = g >= 0;
  const int c = REV ? C0 - (t - g) : C0 + (t - g);
  // the emission of step t+1 depends on nothing but its sample (staged in shared memory): it is evaluated here,
  // independently of the state update of step t below, so that the two dependency chains overlap
  const double p = Q.p_cur;
  const int kk = Q.k_cur;
  // (a free slot would read a sample outside the staged window: give it a harmless one)
  const double x_next = ring_read(R, sample_index<REV>(v, C0, g, t + 1));
  lane_emit(Q.L, exp_tab, live ? x_next : 0.0, Q.p_cur, Q.k_cur);
  const int want = g - 1;
  const bool first = g == 0;                               // the all-ones initial row (dtw.cpp:50,66,182,190)
  const bool ones = first && c >= ones_s && c <= ones_e;
  const bool m0 = live && !first && tag0 == want;
  const bool m1 = live && !first && !m0 && have1 && tag1 == want;
  LaneOut in;  // neither: the producer has left its band, nothing flows any more
  in.f = ones ? 1.0 : (m0 ? in0.f : (m1 ? in1.f : 0.0));
  in.E = ones ? 0 : (m0 ? in0.E : (m1 ? in1.E : NVB_EZERO));
  in.p = m0 ? in0.p : (m1 ? in1.p : 1.0);
  in.k = m0 ? in0.k : (m1 ? in1.k : 0);
  lane_update<MEL, MODE, false, false>(Q.L, Q.S, c, p, kk, in, 1.0, 0, Q.out, Q.aout);
}

template <int MODE>
__device__ __forceinline__ LaneOut rotate_out(const LaneOut &o, int src) {
  LaneOut r;
  r.f = __shfl_sync(NVB_FULL, o.f, src);
  r.E = __shfl_sync(NVB_FULL, o.E, src);
  if (MODE == NVB_MODE_WOBBLE) {
    r.p = __shfl_sync(NVB_FULL, o.p, src);
    r.k = __shfl_sync(NVB_FULL, o.k, src);
  } else {
    r.p = 1.0; r.k = 0;
  }
  return r;
}

// Row [s, e] at plane offset `off`, swept by pair `pair`: column at step tt is C0 + (tt - pair) forward, C0 - (tt - pair)
// in reverse, plane index off + (column - s).
template <bool REV>
__device__ __forceinline__ RowMeta row_meta(bool live, long long off, int s, int e, int pair, int C0) {
  RowMeta m;
  if (REV) { m.base = off - s + C0 + pair; m.t_lo = C0 + pair - e; m.t_hi = C0 + pair - s; }
  else { m.base = off - s + C0 - pair; m.t_lo = s - C0 + pair; m.t_hi = e - C0 + pair; }
  if (!live) { m.t_lo = 1; m.t_hi = 0; }
  return m;
}

template <int MEL, bool REV>
__device__ __forceinline__ void write_meta(const StoreTile &tb, const StoreTile &ta, const Slot<MEL> &Q, int lane,
                                           bool trans, int C0) {
  const bool live = Q.pair >= 0;
  tb.meta[lane] = row_meta<REV>(live, Q.boff, Q.L.ms, Q.L.me, Q.pair, C0);
  if (trans) ta.meta[lane] = row_meta<REV>(live, Q.aoff, Q.ws, Q.awe, Q.pair, C0);
}

template <int MEL, int MODE, bool REV>
__device__ void sweep_rotate(const ModelDev &M, const ReadView &v, double *F, int32_t *X, int lane,
                             const StoreTile (&tiles)[4], const SignalRing<8> &R, unsigned exp_tab, PairRec *recs) {
  const int n = v.n;
  constexpr bool TRANS = (MODE == NVB_MODE_TRANS);
  constexpr int TS = tile_steps(MODE), TSTRIDE = TS + 1;
  // tiles: B rows of slot P and slot Q2, then (transition sweep only) their A rows
  const StoreTile &tPB = tiles[0], &tSB = tiles[1], &tPA = tiles[TRANS ? 2 : 0], &tSA = tiles[TRANS ? 3 : 1];

  // all-ones first row (dtw.cpp:50,66,182,190): written to HBM here, generated on the fly for pair 0
  const int j0 = REV ? n : 0;
  const int ones_s = v.bs[j0], ones_e = v.be[j0];
  {
    const int64_t off0 = REV ? (TRANS ? trans_row_off(v, 2 * n - 1) : v.coff[n]) : 0;
    for (int c = ones_s + lane; c <= ones_e; c += NVB_WARP) { F[off0 + c - ones_s] = 1.0; X[off0 + c - ones_s] = 0; }
  }
  const int C0 = REV ? ones_e : ones_s;
  const int T_total = ((pair_t_end<REV>(v, C0, n - 1) + 1 + TS - 1) / TS) * TS;

  Slot<MEL> P, Q2;
  slot_clear(P);
  slot_clear(Q2);
  int next_g = lane;                                   // next pair this lane will take
  int next_start = (next_g < n) ? pair_t_start<REV>(v, C0, next_g) : 0x7fffffff;
  PairRec *rec = recs + lane;                          // parameters of pair next_g, in flight or landed
  if (next_g < n) rec_prefetch(rec, v, pair_base<REV>(n, next_g));
  const long long coff1 = TRANS ? v.coff[1] : 0;
  bool any2_tile = false;                              // some lane had a live second slot during this tile

  // One iteration per store tile: slot management, then TS steps (two copies of the step loop, with and without the
  // second slot, so that the steps themselves carry no tile / flush / renormalisation / second-slot tests), then the
  // flush.
  static_assert(((NVB_RENORM_MASK + 1) % TS) == 0 || (TS % (NVB_RENORM_MASK + 1)) == 0, "tile vs renormalisation period");
  const int src = (lane + NVB_WARP - 1) & (NVB_WARP - 1);
  for (int t = 0; t < T_total; t += TS) {
    {
      // ---- slot management, only at tile boundaries -----------------------------------------------------------
      // a pair keeps its slot until the step AFTER its last in-band one: that is when its last cell is consumed
      if (P.pair >= 0 && t >= P.t_end + 2) {
        if (Q2.pair >= 0) { P = Q2; slot_clear(Q2); } else slot_clear(P);
      }
      const bool starting = next_start < t + TS;  // its first in-band column falls into this tile
      {
        // signal window of this tile: the samples the live pairs (and the one starting now) read at steps t .. t+TS
        // (the last step evaluates the emission of step t+TS), plus what to have in flight beyond it
        int glo = 0x7fffffff, ghi = -1;
        if (P.pair >= 0) { glo = min(glo, P.pair); ghi = max(ghi, P.pair); }
        if (Q2.pair >= 0) { glo = min(glo, Q2.pair); ghi = max(ghi, Q2.pair); }
        if (starting) { glo = min(glo, next_g); ghi = max(ghi, next_g); }
        glo = __reduce_min_sync(NVB_FULL, glo);
        ghi = __reduce_max_sync(NVB_FULL, ghi);
        if (ghi >= 0) {
          const int last = v.N - 1;
          const int lo = REV ? C0 - t - TS + glo : C0 + t - ghi - 1;
          const int hi = REV ? C0 - t + ghi : C0 + t + TS - glo - 1;
          const int need_lo = min(max(lo, 0), last), need_hi = min(max(hi, 0), last);
          const int ahead_lo = min(max(lo - (REV ? 48 : 32), 0), last), ahead_hi = min(max(hi + (REV ? 32 : 48), 0), last);
          ring_advance<REV, 8>(R, need_lo, need_hi, ahead_lo, ahead_hi, lane);
        }
      }
      if (starting) {
        if (P.pair < 0) slot_start<MEL, MODE, REV>(P, v, R, exp_tab, rec, coff1, C0, next_g, t);
        else slot_start<MEL, MODE, REV>(Q2, v, R, exp_tab, rec, coff1, C0, next_g, t);  // the band kernel guarantees Q2 is free
        next_g += NVB_WARP;
        next_start = (next_g < n) ? pair_t_start<REV>(v, C0, next_g) : 0x7fffffff;
        if (next_g < n) rec_prefetch(rec, v, pair_base<REV>(n, next_g));  // its record was read into the slot above
      }
      any2_tile = __any_sync(NVB_FULL, Q2.pair >= 0);
      write_meta<MEL, REV>(tPB, tPA, P, lane, TRANS, C0);
      if (any2_tile) write_meta<MEL, REV>(tSB, tSA, Q2, lane, TRANS, C0);
      __syncwarp();
    }
    if (!any2_tile) {
#pragma unroll 1
      for (int k = 0; k < TS; k++) {
        // neighbour outputs of the previous step: rotation by one lane, tagged with the pair index
        const LaneOut in0 = rotate_out<MODE>(P.out, src);
        const int tag0 = __shfl_sync(NVB_FULL, P.pair, src);
        LaneOut none;
        none.f = 0.0; none.E = NVB_EZERO; none.p = 1.0; none.k = 0;
        slot_step<MEL, MODE, REV>(P, v, R, exp_tab, C0, t + k, in0, tag0, none, -1, false, ones_s, ones_e);
        tPB.f[lane * TSTRIDE + k] = P.out.f;
        tPB.x[lane * TSTRIDE + k] = P.out.E;
        if (TRANS) { tPA.f[lane * TSTRIDE + k] = P.aout.f; tPA.x[lane * TSTRIDE + k] = P.aout.e; }
        if (((t + k) & NVB_RENORM_MASK) == NVB_RENORM_MASK && TS > NVB_RENORM_MASK + 1) lane_renorm(P.S);
      }
    } else {
#pragma unroll 1
      for (int k = 0; k < TS; k++) {
        const LaneOut in0 = rotate_out<MODE>(P.out, src);
        const int tag0 = __shfl_sync(NVB_FULL, P.pair, src);
        const LaneOut in1 = rotate_out<MODE>(Q2.out, src);
        const int tag1 = __shfl_sync(NVB_FULL, Q2.pair, src);
        slot_step<MEL, MODE, REV>(P, v, R, exp_tab, C0, t + k, in0, tag0, in1, tag1, true, ones_s, ones_e);
        tPB.f[lane * TSTRIDE + k] = P.out.f;
        tPB.x[lane * TSTRIDE + k] = P.out.E;
        if (TRANS) { tPA.f[lane * TSTRIDE + k] = P.aout.f; tPA.x[lane * TSTRIDE + k] = P.aout.e; }
        slot_step<MEL, MODE, REV>(Q2, v, R, exp_tab, C0, t + k, in0, tag0, in1, tag1, true, ones_s, ones_e);
        tSB.f[lane * TSTRIDE + k] = Q2.out.f;
        tSB.x[lane * TSTRIDE + k] = Q2.out.E;
        if (TRANS) { tSA.f[lane * TSTRIDE + k] = Q2.aout.f; tSA.x[lane * TSTRIDE + k] = Q2.aout.e; }
        if (((t + k) & NVB_RENORM_MASK) == NVB_RENORM_MASK && TS > NVB_RENORM_MASK + 1) {
          lane_renorm(P.S);
          lane_renorm(Q2.S);
        }
      }
    }
#ifndef NVB_EXPERIMENT_NO_STORE  // (timing experiment only, no results: what a step costs without the tile flush)
    __syncwarp();
    flush_tile<REV, TS>(tPB, F, X, C0, t, lane);
    if (TRANS) flush_tile<REV, TS>(tPA, F, X, C0, t, lane);
    if (any2_tile) {
      flush_tile<REV, TS>(tSB, F, X, C0, t, lane);
      if (TRANS) flush_tile<REV, TS>(tSA, F, X, C0, t, lane);
    }
    __syncwarp();
#endif
    // renormalisation every NVB_RENORM_MASK + 1 steps: at the end of the tile that completes a period (tiles longer than
    // a period renormalise inside their step loop above)
    if (TS <= NVB_RENORM_MASK + 1 && ((t + TS - 1) & NVB_RENORM_MASK) == NVB_RENORM_MASK) {
      lane_renorm(P.S);
      if (any2_tile) lane_renorm(Q2.S);
    }
  }
  ring_drain(R);
}

template <int MEL, int MODE>
__global__ void __launch_bounds__(64, NVB_ROT_MIN_BLOCKS) sweep5_kernel(ModelDev M, BatchDev B, int b0, int n_items,
                                                     const int64_t *mat_base, double *pF, int32_t *pX, double *sF,
                                                     int32_t *sX) {
  extern __shared__ unsigned long long smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int item = blockIdx.x * (blockDim.x >> 5) + warp;  // (read, direction)
  const unsigned exp_tab = exp_table_init();
  if (item >= n_items) return;
  const int b = b0 + (item >> 1);
  if (B.flags[b] != 0) return;  // bad band
  unsigned char *base = reinterpret_cast<unsigned char *>(smem_raw) + (size_t)warp * warp_bytes(MODE);
  StoreTile tiles[4];
#pragma unroll
  for (int i = 0; i < tiles_per_warp(MODE); i++) {
    unsigned char *p = base + (size_t)i * tile_bytes(tile_steps(MODE));
    tiles[i].f = reinterpret_cast<double *>(p);
    tiles[i].meta = reinterpret_cast<RowMeta *>(tiles[i].f + 32 * (tile_steps(MODE) + 1));
    tiles[i].x = reinterpret_cast<int32_t *>(tiles[i].meta + 32);
  }
  ReadView v = read_view(B, b);
  const int64_t mb = mat_base[b];
  // signal ring (dp3.cuh) behind the four store tiles: 8 chunks of 32 samples + 8 mbarriers, filled by TMA bulk copies
  SignalRing<8> R;
  ring_init(R, base + tiles_per_warp(MODE) * tile_bytes(tile_steps(MODE)), B.signal, B.sig_off[b], B.sig_off[B.n_reads], lane);
  PairRec *recs = reinterpret_cast<PairRec *>(base + rec_offset(MODE));
  if (item & 1) sweep_rotate<MEL, MODE, true>(M, v, sF + mb, sX + mb, lane, tiles, R, exp_tab, recs);
  else sweep_rotate<MEL, MODE, false>(M, v, pF + mb, pX + mb, lane, tiles, R, exp_tab, recs);
}

template <int MEL, int MODE>
void launch_mode(const ModelDev &M, const BatchDev &B, int b0, int n_items, const int64_t *mb, double *pF, int32_t *pX,
                 double *sF, int32_t *sX, cudaStream_t st) {
  const int warps = 2;  // 2 x (2 tiles x 3.9 KB + 2 KB ring + 2 KB pair records) = 24 KB per CTA (40 KB in the transition
                        // sweep): 8 CTAs per SM (5) beside the 2 KB exp table and the 1 KB the driver reserves per CTA
  const size_t smem = (size_t)warps * warp_bytes(MODE);
  sweep5_kernel<MEL, MODE><<<(n_items + warps - 1) / warps, warps * NVB_WARP, smem, st>>>(M, B, b0, n_items, mb, pF, pX,
                                                                                         sF, sX);
}

template <int MEL>
void launch_sweep5(const ModelDev &M, const BatchDev &B, int mode, int b0, int n_items, const int64_t *mb, double *pF,
                   int32_t *pX, double *sF, int32_t *sX, cudaStream_t st) {
  switch (mode) {
    case NVB_MODE_PLAIN: launch_mode<MEL, NVB_MODE_PLAIN>(M, B, b0, n_items, mb, pF, pX, sF, sX, st); break;
    case NVB_MODE_TRANS: launch_mode<MEL, NVB_MODE_TRANS>(M, B, b0, n_items, mb, pF, pX, sF, sX, st); break;
    default: launch_mode<MEL, NVB_MODE_WOBBLE>(M, B, b0, n_items, mb, pF, pX, sF, sX, st); break;
  }
}

}  // namespace

// Rotating-wavefront sweep of the reads [b0, b1) whose flag is NVB_READ_OK.  Returns -1 for an unsupported
// min_event_length.
int nvbk_sweep_rotate(const ModelDev &M, const BatchDev &B, int mode, int b0, int b1, const int64_t *d_mat_base,
                      double *pF, int32_t *pX, double *sF, int32_t *sX, cudaStream_t st) {
  const int n_items = 2 * (b1 - b0);
  if (n_items <= 0) return 0;
  switch (B.mel) {
    case 0: launch_sweep5<0>(M, B, mode, b0, n_items, d_mat_base, pF, pX, sF, sX, st); break;
    case 1: launch_sweep5<1>(M, B, mode, b0, n_items, d_mat_base, pF, pX, sF, sX, st); break;
    case 2: launch_sweep5<2>(M, B, mode, b0, n_items, d_mat_base, pF, pX, sF, sX, st); break;
    case 3: launch_sweep5<3>(M, B, mode, b0, n_items, d_mat_base, pF, pX, sF, sX, st); break;
    case 4: launch_sweep5<4>(M, B, mode, b0, n_items, d_mat_base, pF, pX, sF, sX, st); break;
    case 5: launch_sweep5<5>(M, B, mode, b0, n_items, d_mat_base, pF, pX, sF, sX, st); break;
    case 6: launch_sweep5<6>(M, B, mode, b0, n_items, d_mat_base, pF, pX, sF, sX, st); break;
    default: return -1;
  }
  return 0;
}
