// dp3.cuh -- the scaled linear-domain DP cell shared by the row sweeps (rows4.cu, rows5.cu) and the SNP kernel (snp3.cu).
//
// The reference carries every DP cell as a log-probability and pays one exp + one log per cell
// (probability.cpp:33-40).  Here a cell is a double mantissa f and an int32 binary exponent e,
// value = f * 2^e, so a cell update is a couple of FMAs; the only transcendental left is ONE exp per lane and step
// (the Gaussian emission), computed directly in (mantissa, exponent) form so that no likelihood, however small,
// ever underflows.  Mathematically this is the same sum-product recurrence as Node::NextRow
// (node_next_row.h:6-61):
//     A-row (wobble / transition, m = 0):  A[c] = P[c] + mixemis(c-1) * A[c-1]
//     B-row (model, m = min_event_length): B[c] = e(c-1) * B[c-1] + (prod_{j=c-m}^{c-1} e(j)) * A[c-m]
// with P the previous B-row.  One lane owns one (A-row, B-row) pair; lanes form a wavefront skewed by one step,
// neighbour values travel by warp shuffle.
//
// Exponent bookkeeping: every running value (A cell, B cell, each A value in flight to the B-row) carries its OWN
// exponent.  Multiplying by an emission p * 2^k multiplies the mantissa by p and adds k to the exponent (exact, any
// range); adding two terms aligns BOTH on the larger exponent (select-free: two power-of-two scale factors, one of
// them 1.0), so a term is dropped only when it is below 2^-1022 of the other term OF THE SAME CELL -- the reference
// itself drops it below e^-37 (log(1 + exp(b - a)) == 0).
//
// Zero is (0.0, NVB_EZERO) with NVB_EZERO = -2^29.  Exponents of zero mantissas are never special-cased: they drift
// by the emission exponents like any other (a few thousand per row at most) and stay hundreds of millions below
// every real exponent (>= -12M for 20k samples at the worst emission), so max() never picks them, and the sum of two
// of them cannot overflow int32.  Mantissas are renormalised every 8 steps (see NVB_RENORM_MASK).
//
// Instruction diet (ncu, profiles/r01b: the v2 kernels were instruction-issue bound at ~190 warp instructions per
// wavefront step): polynomial coefficients come from the constant bank as DFMA operands instead of 2 UMOV each,
// the JOIN role shares the B-row code (emission 1, inflow multiplied by the suffix cell), rows without an A-row are
// handled by data (zero mixture weight, open band) instead of a divergent branch, and the kernel is specialised on the
// row mode so the plain sweep carries no A-row code and the transition sweep no mixture.
#pragma once
#include "common.cuh"

// Mantissas are renormalised when (step & mask) == mask, i.e. every 8 steps.  The period bounds how stale a
// mantissa can get: alignment takes the larger EXPONENT, so a value whose mantissa has decayed hands its slack to
// whatever is added to it, and the slack compounds lane after lane (one hop per step).  With at most 2^-1 per step
// and hop (emission mantissas in [0.98, 1.98], transition weight 0.64 * 2^-6) 8 steps x 8 hops stay far inside the
// double range; 32 steps did not (measured: garbage in the far tails of transition rows).
#ifndef NVB_RENORM_MASK
#define NVB_RENORM_MASK 7
#endif

#define NVB_EZERO (-(1 << 29))  // exponent carried by values that are exactly zero

#define NVB_LN2_HI 6.93147180369123816490e-01
#define NVB_LN2_LO 1.90821492927058770002e-10
#define NVB_LOG2E 1.44269504088896338700e+00

// 2^e as a double for e <= 0; e <= -1023 gives +0.0
__device__ __forceinline__ double pow2neg(int e) {
  e = max(e, -1023);
  return __hiloint2double((e + 1023) << 20, 0);
}

// exp(l) = p * 2^k with p in [0.98, 1.98], any finite l (no underflow: k is returned, not applied).
// l = (32 k + j) ln2/32 + r, |r| <= ln2/64: Cody-Waite reduction against ln2/32 (hi/lo; the hi part has 32 significant
// bits, so n * hi is exact for |n| < 2^21), 2^(j/32) from a 32-entry table in shared memory, exp(r) - 1 by a degree-6
// Taylor polynomial (truncation 3.5e-18; worst relative error of p measured against 60-digit arithmetic: 2.0e-16).
// The table replaces 6 of the 13 dependent DFMAs of the single-interval polynomial this started as: the FP64 pipe
// issues one warp instruction every two cycles, and in the latency-bound sweeps the Horner chain is on the critical path.
#define NVB_32_LOG2E 46.16624130844683
#define NVB_LN2_32_HI 0.02166084938653512
#define NVB_LN2_32_LO 5.9631716539705866e-12

static __constant__ double c_exp_tab[32] = {
    1.0,                1.0218971486541166, 1.0442737824274138, 1.0671404006768237, 1.0905077326652577,
    1.1143867425958924, 1.1387886347566916, 1.1637248587775775, 1.189207115002721,  1.215247359980469,
    1.241857812073484,  1.2690509571917332, 1.2968395546510096, 1.3252366431597413, 1.3542555469368927,
    1.383909881963832,  1.4142135623730951, 1.4451808069770467, 1.4768261459394993, 1.5091644275934228,
    1.5422108254079407, 1.5759808451078865, 1.6104903319492543, 1.645755478153965,  1.681792830507429,
    1.718619298122478,  1.7562521603732995, 1.7947090750031072, 1.8340080864093424, 1.8741676341103,
    1.9152065613971474, 1.9571441241754002};
// 1/6! .. 1/3!; read as constant-bank operands of the DFMAs
static __constant__ double c_exp_poly[4] = {1.388888888888889e-03, 8.333333333333333e-03, 4.1666666666666664e-02,
                                            1.6666666666666666e-01};

static __shared__ double s_exp_tab[32];

// Every kernel that evaluates emissions calls this once, with ALL threads of the CTA, before anything else.
__device__ __forceinline__ void exp_table_init() {
  if (threadIdx.x < 32) s_exp_tab[threadIdx.x] = c_exp_tab[threadIdx.x];
  __syncthreads();
}

#ifndef NVB_EXP_TABLE
#define NVB_EXP_TABLE 1
#endif

#if !NVB_EXP_TABLE
// The single-interval form (round 1): Cody-Waite against ln2 and a degree-13 Taylor polynomial on |r| <= 0.347
// (error < 5e-18), p in [0.70, 1.42].  No table, 13 dependent DFMAs.
#define NVB_LOG2E 1.44269504088896338700e+00
static __constant__ double c_exp_poly13[11] = {
    1.6059043836821613e-10, 2.08767569878681e-09,   2.505210838544172e-08,  2.755731922398589e-07,
    2.7557319223985893e-06, 2.48015873015873e-05,   1.984126984126984e-04,  1.388888888888889e-03,
    8.333333333333333e-03,  4.1666666666666664e-02, 1.6666666666666666e-01};
__device__ __forceinline__ void exp_ext(double l, double &p, int &k) {
  const double magic = 6755399441055744.0;
  double t = fma(l, NVB_LOG2E, magic);
  k = __double2loint(t);
  double kd = t - magic;
  double r = fma(-kd, NVB_LN2_HI, l);
  r = fma(-kd, NVB_LN2_LO, r);
  double q = c_exp_poly13[0];
#pragma unroll
  for (int i = 1; i < 11; i++) q = fma(q, r, c_exp_poly13[i]);
  q = fma(q, r, 0.5);
  q = fma(q, r, 1.0);
  p = fma(q, r, 1.0);
}
#else
__device__ __forceinline__ void exp_ext(double l, double &p, int &k) {
  const double magic = 6755399441055744.0;  // 1.5 * 2^52: rint via add/sub, integer in the low word
  const double t = fma(l, NVB_32_LOG2E, magic);
  const int n = __double2loint(t);
  const double kd = t - magic;
  double r = fma(-kd, NVB_LN2_32_HI, l);
  r = fma(-kd, NVB_LN2_32_LO, r);
  k = n >> 5;
  const double T = s_exp_tab[n & 31];
  double q = fma(c_exp_poly[0], r, c_exp_poly[1]);
  q = fma(q, r, c_exp_poly[2]);
  q = fma(q, r, c_exp_poly[3]);
  q = fma(q, r, 0.5);
  q = fma(q, r, 1.0);
  q *= r;             // exp(r) - 1
  p = fma(T, q, T);
}
#endif

// natural log of f * 2^E (f > 0), E*ln2 added in two pieces
__device__ __forceinline__ double log_ext(double f, int E) {
  if (!(f > 0.0)) return nvb_neg_inf();
  return fma((double)E, NVB_LN2_HI, log(f)) + (double)E * NVB_LN2_LO;
}

// value = f * 2^e; zero is (0, NVB_EZERO)
struct XD {
  double f;
  int e;
};

__device__ __forceinline__ XD xd_make(double f, int e) {
  XD v;
  v.f = f; v.e = e;
  return v;
}
__device__ __forceinline__ XD xd_zero() { return xd_make(0.0, NVB_EZERO); }

// a + b, both aligned on the larger exponent (no selects; one of the two scale factors is 1.0)
__device__ __forceinline__ XD xd_add(XD a, XD b) {
  XD r;
  r.e = max(a.e, b.e);
  r.f = fma(b.f, pow2neg(b.e - r.e), a.f * pow2neg(a.e - r.e));
  return r;
}

// mantissa back into [1,2) (normal inputs only); zero gets the zero exponent back
__device__ __forceinline__ void xd_renorm(XD &v) {
  if (v.f > 0.0) {
    const int hi = __double2hiint(v.f);
    v.e += ((hi >> 20) & 0x7ff) - 1023;
    v.f = __hiloint2double((hi & 0x800fffff) | 0x3ff00000, __double2loint(v.f));
  } else {
    v.e = NVB_EZERO;
  }
}

// ---- signal staging: a per-warp ring in shared memory, filled by 1-D TMA bulk copies -------------------------------
// A sweep warp reads, at every step, one sample per lane out of a window of at most 64 consecutive samples that slides
// by (at most) one sample per step.  Fetching them with per-lane loads put the global-load latency on the critical
// path of every step (ncu, profiles/r01e: 27 % of the stall samples of the hot loop were the first use of that
// load).  Instead the warp keeps NS = 4 or 8 chunks of 32 samples in shared memory: chunk k holds the samples with
// ABSOLUTE index [32k, 32k+32) of the batch's signal array (absolute, so that every chunk starts on a 256-byte
// boundary of the allocation: cp.async.bulk needs 16-byte alignment and read slices start anywhere), one elected lane
// issues cp.async.bulk (global -> shared, completion on an mbarrier) a few dozen steps before the window reaches the
// chunk, and the warp waits on the mbarrier only when it first needs it.
// Bookkeeping of a ring; lives in shared memory next to the ring (it is touched only every few steps, by the whole
// warp with uniform values, and must not cost registers in the step loop).
struct RingState {
  const double *base;       // signal array, rounded down to a 256-sample boundary in front of the read's slice
  int limit;                // chunks at or beyond this id lie outside the allocation
  int issued_lo, issued_hi; // chunk ids (relative to `base`) issued so far: [issued_lo, issued_hi]; empty when lo > hi
  int ready_lo, ready_hi;   // ... and known to have landed
  unsigned phases;          // bit q: parity the NEXT copy into slot q completes
  int pad;
};
// NS slots of 32 samples, NS mbarriers, the state: NS = 8 for the rotating sweep (window of up to 64 samples), 4 for the
// striped sweep (window of 32)
__host__ __device__ constexpr int ring_bytes(int NS) { return NS * 256 + NS * 8 + (int)sizeof(RingState); }

template <int NS>
struct SignalRing {
  unsigned buf;  // shared-memory address of the NS * 32 samples; the NS mbarriers and the RingState follow
  int off;       // index of the read's sample 0 relative to RingState::base (0..255)
};

__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
template <int NS>
__device__ __forceinline__ RingState *ring_state(const SignalRing<NS> &R) {
  return reinterpret_cast<RingState *>(__cvta_shared_to_generic(R.buf + NS * 256 + NS * 8));
}

// `mem` = ring_bytes(NS) of 16-byte aligned shared memory owned by this warp, `signal` = the batch's signal array
// (256-byte aligned allocation of `total` samples), `first` = absolute index of the read's sample 0.
template <int NS>
__device__ __forceinline__ void ring_init(SignalRing<NS> &R, void *mem, const double *signal, long long first,
                                          long long total, int lane) {
  const long long aligned = first & ~255LL;
  R.buf = smem_u32(mem);
  R.off = (int)(first - aligned);
  if (lane < NS) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(R.buf + NS * 256 + lane * 8));
  if (lane == 0) {
    RingState *S = ring_state(R);
    const long long chunks = (total - aligned + 31) >> 5;
    S->base = signal + aligned;
    S->limit = (int)(chunks < 0x7fffffff ? chunks : 0x7fffffff);
    S->issued_lo = 1; S->issued_hi = 0; S->ready_lo = 1; S->ready_hi = 0;
    S->phases = 0;
  }
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncwarp();
}

// Lane 0 issues the bulk copy of chunk `k` into its ring slot (256 bytes, both addresses 256-byte aligned) and flips
// the slot's phase bit.
template <int NS>
__device__ __forceinline__ void ring_issue(const SignalRing<NS> &R, RingState *S, int k) {
  if (k < 0 || k >= S->limit) return;
  const unsigned slot = (unsigned)k & (unsigned)(NS - 1);
  S->phases ^= 1u << slot;
  const unsigned bar = R.buf + NS * 256 + slot * 8, dst = R.buf + slot * 256;
  // the slot was last read through the generic proxy: order those reads before the async-proxy write
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], 256;" ::"r"(bar) : "memory");
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], 256, [%2];"
               ::"r"(dst), "l"(S->base + (long long)k * 32), "r"(bar) : "memory");
}

// All lanes wait until the latest copy into chunk k's slot has landed (at most one copy per slot is in flight: a slot
// is reused NS chunks later, long after its previous chunk was waited for).
template <int NS>
__device__ __forceinline__ void ring_wait(const SignalRing<NS> &R, int k, int limit, unsigned phases) {
  if (k < 0 || k >= limit) return;
  const unsigned slot = (unsigned)k & (unsigned)(NS - 1);
  const unsigned bar = R.buf + NS * 256 + slot * 8, parity = ((phases >> slot) & 1u) ^ 1u;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "RING_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra RING_DONE;\n"
      "bra RING_WAIT;\n"
      "RING_DONE:\n"
      "}\n" ::"r"(bar), "r"(parity) : "memory");
}

// Keep the ring ahead of the window.  Called by the whole warp with warp-uniform arguments (sample indices of the read,
// clamped by the caller to [0, N-1]): the steps until the next call read samples [need_lo, need_hi]; samples
// [ahead_lo, ahead_hi] (a superset, fewer than (NS - 1) * 32 samples) are to be in flight.  A forward sweep (REV = false) only
// ever extends the ring upwards after its first call, a reverse sweep only downwards: callers keep one chunk of
// slack on the trailing side so that the small backward wobbles of the window never reach below what is resident.
template <int NS>
__device__ __forceinline__ void ring_reset(const SignalRing<NS> &R, int lane);

template <bool REV, int NS>
__device__ __forceinline__ void ring_advance(const SignalRing<NS> &R, int need_lo, int need_hi, int ahead_lo, int ahead_hi,
                                             int lane) {
  RingState *S = ring_state(R);
  const int nl = (R.off + need_lo) >> 5, nh = (R.off + need_hi) >> 5;
  const int al = (R.off + ahead_lo) >> 5, ah = (R.off + ahead_hi) >> 5;
  bool first = S->issued_lo > S->issued_hi;
  // a window that jumped by a whole ring (it never does in the sweeps as they are) starts over instead of queueing
  // several copies on one slot's mbarrier
  if (!first && (REV ? (S->issued_lo - al >= NS) : (ah - S->issued_hi >= NS))) {
    ring_reset(R, lane);
    first = true;
  }
  if (first || (!REV ? (ah > S->issued_hi) : (al < S->issued_lo))) {
    __syncwarp();  // every lane is done reading the slots that are about to be overwritten (and the state)
    if (lane == 0) {
      if (first) {  // the initial fill
        if (!REV) { S->issued_lo = al; S->issued_hi = al - 1; S->ready_lo = al; S->ready_hi = al - 1; }
        else { S->issued_hi = ah; S->issued_lo = ah + 1; S->ready_hi = ah; S->ready_lo = ah + 1; }
      }
      if (!REV) while (S->issued_hi < ah) ring_issue(R, S, ++S->issued_hi);
      else while (S->issued_lo > al) ring_issue(R, S, --S->issued_lo);
    }
    __syncwarp();
  }
  const int limit = S->limit;
  const unsigned phases = S->phases;
  if (!REV) {
    int r = S->ready_hi;
    if (r < nh) {
      while (r < nh) ring_wait(R, ++r, limit, phases);
      __syncwarp();
      if (lane == 0) S->ready_hi = r;
    }
  } else {
    int r = S->ready_lo;
    if (r > nl) {
      while (r > nl) ring_wait(R, --r, limit, phases);
      __syncwarp();
      if (lane == 0) S->ready_lo = r;
    }
  }
}

// Before the warp exits: no bulk copy may still be in flight towards its shared memory.  Waits for the latest copy
// into each of the 8 slots (a slot that was never used reports completion at once).
template <int NS>
__device__ __forceinline__ void ring_drain(const SignalRing<NS> &R) {
  __syncwarp();
  const unsigned phases = ring_state(R)->phases;
#pragma unroll 1
  for (int slot = 0; slot < NS; slot++) ring_wait(R, slot, NS, phases);
}

// Forget what is resident (after draining): the next ring_advance starts with an initial fill.  For sweeps whose
// window jumps (every stripe of the striped sweep starts at its own band start).
template <int NS>
__device__ __forceinline__ void ring_reset(const SignalRing<NS> &R, int lane) {
  ring_drain(R);
  if (lane == 0) {
    RingState *S = ring_state(R);
    S->issued_lo = 1; S->issued_hi = 0; S->ready_lo = 1; S->ready_hi = 0;
  }
  __syncwarp();
}

// Sample i of the read (its chunk must be resident).
template <int NS>
__device__ __forceinline__ double ring_read(const SignalRing<NS> &R, int i) {
  double x;
  asm volatile("ld.shared.f64 %0, [%1];" : "=d"(x) : "r"(R.buf + (((R.off + i) & (NS * 32 - 1)) << 3)));
  return x;
}

enum { NVB_ROLE_IDLE = 0, NVB_ROLE_LOADER = 1, NVB_ROLE_PAIR = 2, NVB_ROLE_JOIN = 3 };

// Per-lane constants of one stripe / task.
struct LaneCfg {
  int role;
  int ws, we;        // A-row band (INT_MIN..INT_MAX for lanes without an A-row: their A-row is P itself).  Both ends
                     // are needed: the far end of the band is `we` in a forward sweep and `ws` in a reverse sweep
  int ms, me;        // B-row band; LOADER: band of the row it loads; JOIN: band of the closing suffix row
  double mu, ac, mc; // own Gaussian emission (PAIR: model row; LOADER: the row before the first pair; JOIN: last+1)
  double cm;         // A-row mixture weight: exp(-2) (kmer_model.cpp:60), 0 for lanes without an A-row
  int abias;         // 0, or NVB_EZERO for lanes without an A-row (their mixture term must stay zero-class)
  double pc;         // transition rows: constant emission 0.01 = 0.64 * 2^-6, or 0 (kmer_model.cpp:64-94)
  int kc;            // ... and its exponent (-6 / NVB_EZERO)
};

__device__ __forceinline__ void lane_cfg_clear(LaneCfg &L) {
  L.role = NVB_ROLE_IDLE; L.ws = -0x7fffffff - 1; L.we = 0x7fffffff; L.ms = 0x7fffffff; L.me = 0;  // empty B band
  L.mu = 0; L.ac = 0; L.mc = 0; L.cm = 0; L.abias = NVB_EZERO; L.pc = 0; L.kc = NVB_EZERO;
}

template <int MEL>
struct LaneState {
  XD mod, w;                    // running B-row / A-row cells
  XD q[MEL > 0 ? MEL : 1];      // A-row outputs in flight to the B-row (delay = min_event_length)
};

template <int MEL>
__device__ __forceinline__ void lane_reset(LaneState<MEL> &S) {
  S.mod = xd_zero(); S.w = xd_zero();
#pragma unroll
  for (int i = 0; i < (MEL > 0 ? MEL : 1); i++) S.q[i] = xd_zero();
}

// Outputs of a lane at one step, consumed by lane+1 at the next step.
struct LaneOut {
  double f;  // B-row value at this step's column (masked to the row's band): f * 2^E
  int E;
  double p;  // own emission at this step's sample: p * 2^k
  int k;
};

// One wavefront step of one lane.  `c` is the column, `x` the sample this step's emissions are evaluated at,
// `in` the neighbour's output of the previous step.  JOIN lanes (WITH_JOIN kernels only) run the same code with
// emission 1 on their B-row and their A-row cell multiplied by the closing suffix cell (sF, sX), which makes the B
// cell the running sum of Node::TotalLikelihood (node.cpp:31-37), delayed by MEL steps; other lanes pass (1.0, 0).
// `aout` receives the A-row cell (for callers that store it).
//
// PH >= 0 = step index mod MEL: the A-row cells in flight to the B-row live in a ring (slot PH is the one pushed MEL
// steps ago) for callers that unroll their step loop MEL times.  PH < 0: a plain shifting delay line for rolled loops.
// Both kernels use PH < 0: unrolling saved ~6 register moves per step but cost registers -- the SNP kernel dropped
// from 7 to 6 resident CTAs per SM and got 12 % slower (measured), the sweeps would lose their 14 CTAs per SM.
// FWD_ONLY (forward-only callers with wobble rows): the A-row needs only its upper band limit (below the band its
// inflow is already zero) and the B-row output only its lower one (above the band the consumer's own A-row limit cuts
// it off); lanes without an output use ms = INT_MAX.  Measured neutral on the SNP kernel, so currently unused.
// The step comes in two halves so that the latency-bound sweeps can evaluate the emission of step t+1 (which depends
// on nothing but the sample) while the state update of step t waits for its neighbour: lane_emit + lane_update.
__device__ __forceinline__ void lane_emit(const LaneCfg &L, double x, double &p, int &kk) {
#ifdef NVB_EXPERIMENT_NO_EMIT  // timing experiment only (wrong results): what a step costs without its emission
  p = 1.0 + 1e-9 * x; kk = 0;
  return;
#endif
  // own emission, reference formula ac - d*d*mc (kmer_model.cpp:47-51)
  const double d = x - L.mu;
  const double l = L.ac - d * d * L.mc;
  exp_ext(l, p, kk);
}

template <int MEL, int MODE, bool WITH_JOIN, int PH, bool FWD_ONLY>
__device__ __forceinline__ void lane_update(const LaneCfg &L, LaneState<MEL> &S, int c, double p, int kk,
                                            const LaneOut &in, double sF, int sX, LaneOut &out, XD &aout);

template <int MEL, int MODE, bool WITH_JOIN, int PH, bool FWD_ONLY>
__device__ __forceinline__ void lane_step(const LaneCfg &L, LaneState<MEL> &S, int c, double x, const LaneOut &in,
                                          double sF, int sX, LaneOut &out, XD &aout) {
  double p;
  int kk;
  lane_emit(L, x, p, kk);
  lane_update<MEL, MODE, WITH_JOIN, PH, FWD_ONLY>(L, S, c, p, kk, in, sF, sX, out, aout);
}

template <int MEL, int MODE, bool WITH_JOIN, int PH, bool FWD_ONLY>
__device__ __forceinline__ void lane_update(const LaneCfg &L, LaneState<MEL> &S, int c, double p, int kk,
                                            const LaneOut &in, double sF, int sX, LaneOut &out, XD &aout) {
  out.p = p;
  out.k = kk;

  const XD pm = xd_make(in.f, in.E);
  XD wout;
  if (MODE == NVB_MODE_PLAIN) {
    wout = pm;
  } else {
    // A[c] = P[c] + mix * A[c-1]
    XD mix;
    if (MODE == NVB_MODE_TRANS) {
      mix = xd_make(L.pc, L.kc);
    } else {
      mix = xd_add(xd_make(p, kk), xd_make(in.p, in.k));  // own + neighbour emission, weight exp(-2) each
      mix.f *= L.cm;
      mix.e += L.abias;
    }
    S.w = xd_add(xd_make(mix.f * S.w.f, S.w.e + mix.e), pm);
    const bool ina = (FWD_ONLY && MODE == NVB_MODE_WOBBLE) ? (c <= L.we) : (c >= L.ws && c <= L.we);
    wout = ina ? S.w : xd_zero();
  }
  aout = wout;

  double pb = p;
  int kb = kk;
  XD push = wout;
  if (WITH_JOIN) {
    const bool join = L.role == NVB_ROLE_JOIN;
    pb = join ? 1.0 : p;
    kb = join ? 0 : kk;
    push = xd_make(wout.f * sF, wout.e + sX);
  }
  // B[c] = e(c-1) * B[c-1] + (product of the last m emissions) * A[c-m]
  XD popped;
  if (MEL == 0) {
    popped = push;
  } else {
#pragma unroll
    for (int i = 0; i < MEL; i++) { S.q[i].f *= pb; S.q[i].e += kb; }
    if (PH >= 0) {  // ring: slot PH was pushed MEL steps ago
      popped = S.q[PH >= 0 ? PH : 0];
      S.q[PH >= 0 ? PH : 0] = push;
    } else {        // PH < 0: plain delay line (callers that do not unroll their step loop)
      popped = S.q[MEL - 1];
#pragma unroll
      for (int i = MEL - 1; i > 0; i--) S.q[i] = S.q[i - 1];
      S.q[0] = push;
    }
  }
  S.mod = xd_add(xd_make(pb * S.mod.f, S.mod.e + kb), popped);
  const bool inb = (FWD_ONLY && MODE == NVB_MODE_WOBBLE) ? (c >= L.ms) : ((c >= L.ms) && (c <= L.me));
  out.f = inb ? S.mod.f : 0.0;
  out.E = inb ? S.mod.e : NVB_EZERO;
}

// Mantissa renormalisation (call on a warp-uniform schedule, every 32 steps).
template <int MEL>
__device__ __forceinline__ void lane_renorm(LaneState<MEL> &S) {
  xd_renorm(S.mod);
  xd_renorm(S.w);
#pragma unroll
  for (int i = 0; i < (MEL > 0 ? MEL : 1); i++) xd_renorm(S.q[i]);
}

template <int MODE>
__device__ __forceinline__ LaneOut shfl_up_out(const LaneOut &o) {
  LaneOut r;
  r.f = __shfl_up_sync(NVB_FULL, o.f, 1);
  r.E = __shfl_up_sync(NVB_FULL, o.E, 1);
  if (MODE == NVB_MODE_WOBBLE) {  // only the wobble mixture needs the neighbour's emission
    r.p = __shfl_up_sync(NVB_FULL, o.p, 1);
    r.k = __shfl_up_sync(NVB_FULL, o.k, 1);
  } else {
    r.p = 0.0; r.k = NVB_EZERO;
  }
  return r;
}
