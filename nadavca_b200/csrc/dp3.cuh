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
// row mode so the plain sweep carries no A-row code and the transition sweep no mixture.  Round 2 (ncu source page,
// DESIGN.md section 6): the emission exp takes a scaled argument (no hi/lo reduction, no re-materialised constants),
// the B-row recurrence is factored so that a step costs m multiplications and no delay-line moves, and the step loops
// carry no renormalisation / tile tests.
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

// exp(l) = p * 2^k with p in [0.998, 2.0], any finite l (no underflow: k is returned, not applied).
//
// The argument arrives SCALED: ls = l * 256 / ln 2 (the lanes scale their emission constants once, lane_set_emission),
// so n = rint(ls) comes from one add of the 1.5 * 2^52 constant, the remainder ls - n is exact (|ls - n| <= 1/2), and
// l = (256 k + j) ln2/256 + r with r = (ls - n) * ln2/256, |r| <= 1.36e-3:  exp(l) = 2^k * 2^(j/256) * exp(r).
// 2^(j/256) comes from a 256-entry table in shared memory, exp(r) - 1 from a degree-4 Taylor polynomial (truncation
// 3.8e-17; the r^4 coefficient is 1/24 cut to its upper 32 bits, which perturbs exp(r) by 3e-18 and lets ptxas encode
// it as an immediate).  Worst relative error of p against 60-digit arithmetic: 2.3e-16
// (tests/test_host_logic.py::test_device_exp_restated_in_numpy reads the table and the constants out of this file).
// History (profiles/r02q, r02w): the single-interval degree-13 polynomial of round 1 cost 13 dependent DFMAs; the
// 32-entry table with a Cody-Waite reduction against ln2/32 cost 15 FP64 instructions per emission plus 8 integer /
// uniform instructions that only re-materialised its constants (the SNP kernel is bound by instruction issue, and
// ptxas prefers re-materialising to spilling under the 64-register cap); this form needs 12 FP64 and no constant
// traffic.
#define NVB_EXP_SCALE 369.3299304675746        // 256 / ln 2
// ln 2 / 256, 1/24 (lower 32 bits cleared), 1/6 -- twice.  From constant memory (BANK = true) for the SNP kernel, which
// is bound by instruction issue: ptxas hoists constant-bank LOADS into uniform registers outside its step loop, whereas
// it re-materialises literals with two moves each inside it (-6 instructions per step).  As literals (BANK = false) for
// the sweeps, which are bound by the latency of one warp: in their much larger loops the loads are not hoisted and a
// constant-cache access in the dependent chain costs more than two moves (measured: wobble sweep 25.6 -> 27.8 ms).
static __constant__ double c_exp_k[3] = {0.0027076061740622863, 0.041666656732559204, 0.16666666666666666};
#define NVB_EXP_STEP 0.0027076061740622863
#define NVB_EXP_C4 0.041666656732559204
#define NVB_EXP_C3 0.16666666666666666

// 2^(j/256), correctly rounded
static __device__ const double g_exp_tab[256] = {
    1.0, 1.0027112750502025, 1.0054299011128027, 1.0081558981184175,
    1.0108892860517005, 1.0136300849514894, 1.016378314910953, 1.019133996077738,
    1.0218971486541166, 1.0246677928971357, 1.0274459491187637, 1.030231637686041,
    1.0330248790212284, 1.0358256936019572, 1.0386341019613787, 1.041450124688316,
    1.0442737824274138, 1.0471050958792898, 1.0499440858006872, 1.0527907730046264,
    1.0556451783605572, 1.0585073227945128, 1.061377227289262, 1.0642549128844645,
    1.0671404006768237, 1.0700337118202419, 1.0729348675259756, 1.075843889062791,
    1.0787607977571199, 1.0816856149932152, 1.0846183622133092, 1.0875590609177697,
    1.0905077326652577, 1.0934643990728858, 1.0964290818163769, 1.099401802630222,
    1.102382583307841, 1.1053714457017412, 1.1083684117236787, 1.1113735033448175,
    1.1143867425958924, 1.1174081515673693, 1.1204377524096067, 1.12347556733302,
    1.1265216186082418, 1.129575928566288, 1.1326385195987192, 1.1357094141578055,
    1.1387886347566916, 1.1418762039695616, 1.1449721444318042, 1.148076478840179,
    1.1511892299529827, 1.154310420590216, 1.1574400736337511, 1.1605782120274988,
    1.1637248587775775, 1.1668800369524817, 1.1700437696832502, 1.1732160801636373,
    1.1763969916502812, 1.1795865274628758, 1.182784710984341, 1.1859915656609938,
    1.189207115002721, 1.1924313825831512, 1.1956643920398273, 1.1989061670743806,
    1.202156731452703, 1.2054161090051239, 1.2086843236265816, 1.2119613992768012,
    1.215247359980469, 1.2185422298274085, 1.2218460329727576, 1.2251587936371455,
    1.22848053610687, 1.2318112847340759, 1.2351510639369334, 1.2384998981998165,
    1.241857812073484, 1.245224830175258, 1.2486009771892048, 1.2519862778663162,
    1.255380757024691, 1.2587844395497165, 1.2621973503942507, 1.2656195145788063,
    1.2690509571917332, 1.2724917033894028, 1.275941778396392, 1.2794012075056693,
    1.2828700160787783, 1.2863482295460256, 1.2898358734066657, 1.2933329732290895,
    1.2968395546510096, 1.3003556433796506, 1.3038812651919358, 1.3074164459346773,
    1.3109612115247644, 1.3145155879493546, 1.318079601266064, 1.3216532776031575,
    1.3252366431597413, 1.3288297242059544, 1.3324325470831615, 1.3360451382041458,
    1.339667524053303, 1.3432997311868353, 1.3469417862329458, 1.3505937158920345,
    1.3542555469368927, 1.3579273062129011, 1.3616090206382248, 1.365300717204012,
    1.3690024229745905, 1.3727141650876684, 1.3764359707545302, 1.380167867260238,
    1.383909881963832, 1.387662042298529, 1.3914243757719262, 1.3951969099662003,
    1.3989796725383112, 1.4027726912202048, 1.4065759938190154, 1.4103896082172707,
    1.4142135623730951, 1.4180478843204152, 1.4218926021691656, 1.4257477441054942,
    1.42961333839197, 1.433489413367789, 1.4373759974489824, 1.4412731191286257,
    1.4451808069770467, 1.449099089642035, 1.4530279958490526, 1.4569675544014438,
    1.460917794180647, 1.4648787441464057, 1.4688504333369818, 1.4728328908693675,
    1.4768261459394993, 1.4808302278224719, 1.4848451658727524, 1.488870989524397,
    1.4929077282912648, 1.4969554117672355, 1.5010140696264256, 1.5050837316234065,
    1.5091644275934228, 1.5132561874526098, 1.5173590411982147, 1.5214730189088146,
    1.5255981507445384, 1.529734466947287, 1.533881997840956, 1.5380407738316568,
    1.5422108254079407, 1.5463921831410214, 1.550584877685, 1.5547889397770887,
    1.559004400237837, 1.5632312899713576, 1.567469639965553, 1.5717194812923414,
    1.5759808451078865, 1.5802537626528246, 1.5845382652524937, 1.588834384317164,
    1.593142151342267, 1.597461597908627, 1.6017927556826934, 1.606135656416771,
    1.6104903319492543, 1.6148568142048607, 1.6192351351948637, 1.6236253270173289,
    1.6280274218573478, 1.632441451987275, 1.6368674497669644, 1.6413054476440063,
    1.645755478153965, 1.6502175739206177, 1.6546917676561943, 1.6591780921616162,
    1.6636765803267364, 1.6681872651305825, 1.6727101796415966, 1.6772453570178785,
    1.681792830507429, 1.6863526334483934, 1.6909247992693053, 1.6955093614893326,
    1.7001063537185235, 1.7047158096580513, 1.709337763100463, 1.713972247929926,
    1.718619298122478, 1.723278947746274, 1.7279512309618377, 1.732636182022311,
    1.7373338352737062, 1.7420442251551564, 1.746767386199169, 1.7515033530318782,
    1.7562521603732995, 1.761013843037584, 1.7657884359332727, 1.7705759740635547,
    1.7753764925265212, 1.7801900265154245, 1.785016611318935, 1.789856282321401,
    1.7947090750031072, 1.7995750249405351, 1.804454167806624, 1.809346539371032,
    1.8142521755003989, 1.8191711121586085, 1.8241033854070534, 1.8290490314048973,
    1.8340080864093424, 1.8389805867758937, 1.843966568958626, 1.8489660695104508,
    1.8539791250833855, 1.8590057724288205, 1.864046048397789, 1.8690999899412386,
    1.8741676341103, 1.8792490180565602, 1.8843441790323345, 1.8894531543909392,
    1.8945759815869656, 1.8997126981765553, 1.9048633418176741, 1.9100279502703899,
    1.9152065613971474, 1.9203992131630474, 1.925605943636125, 1.930826790987627,
    1.9360617934922943, 1.9413109895286405, 1.9465744175792332, 1.9518521162309783,
    1.9571441241754002, 1.9624504802089273, 1.9677712232331759, 1.9731063922552343,
    1.978456026387951, 1.9838201648502194, 1.9891988469672663, 1.9945921121709402};

static __shared__ double s_exp_tab[256];

__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }

// Every kernel that evaluates emissions calls this once, with ALL threads of the CTA, before anything else; returns
// the shared-memory address of the table (kept in a register: taking it inside the step loop costs three instructions).
__device__ __forceinline__ unsigned exp_table_init() {
  for (int i = threadIdx.x; i < 256; i += blockDim.x) s_exp_tab[i] = g_exp_tab[i];
  __syncthreads();
  unsigned tab = smem_u32(s_exp_tab);
  asm volatile("" : "+r"(tab));  // opaque: otherwise the address is re-derived (S2UR + UMOV + ULEA) at every use
  return tab;
}

template <bool BANK>
__device__ __forceinline__ void exp_ext_scaled(double ls, unsigned tab, double &p, int &k) {
  const double magic = 6755399441055744.0;  // 1.5 * 2^52: rint via add/sub, integer in the low word
  const double t = ls + magic;
  const int n = __double2loint(t);
  const double r = (ls - (t - magic)) * (BANK ? c_exp_k[0] : NVB_EXP_STEP);
  k = n >> 8;
  double T;
  // (bank conflicts of the 256-entry lookup do not show: a timing build that folded the index to 32 entries ran the
  // sweeps and the SNP kernel in the same time, profiles/r02w)
  asm("ld.shared.f64 %0, [%1];" : "=d"(T) : "r"(tab + ((unsigned)(n & 255) << 3)));
  double q = BANK ? fma(r, c_exp_k[1], c_exp_k[2]) : fma(r, NVB_EXP_C4, NVB_EXP_C3);
  q = fma(q, r, 0.5);
  q = fma(q, r, 1.0);
  q *= r;             // exp(r) - 1
  p = fma(T, q, T);
}

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
  double mu, ac, mc; // own Gaussian emission (PAIR: model row; LOADER: the row before the first pair; JOIN: last+1);
                     // ac and mc are SCALED by 256 / ln 2 (lane_set_emission), see exp_ext_scaled
  double cm;         // A-row mixture weight: exp(-2) (kmer_model.cpp:60), 0 for lanes without an A-row
  int abias;         // 0, or NVB_EZERO for lanes without an A-row (their mixture term must stay zero-class)
  double pc;         // transition rows: constant emission 0.01 = 0.64 * 2^-6, or 0 (kmer_model.cpp:64-94)
  int kc;            // ... and its exponent (-6 / NVB_EZERO)
};

__device__ __forceinline__ void lane_cfg_clear(LaneCfg &L) {
  L.role = NVB_ROLE_IDLE; L.ws = -0x7fffffff - 1; L.we = 0x7fffffff; L.ms = 0x7fffffff; L.me = 0;  // empty B band
  L.mu = 0; L.ac = 0; L.mc = 0; L.cm = 0; L.abias = NVB_EZERO; L.pc = 0; L.kc = NVB_EZERO;
}

// Gaussian emission of a lane from the model tables: log density ac - (x - mu)^2 * mc (kmer_model.cpp:47-51).
__device__ __forceinline__ void lane_set_emission(LaneCfg &L, double mu, double ac, double mc) {
  L.mu = mu; L.ac = ac * NVB_EXP_SCALE; L.mc = mc * NVB_EXP_SCALE;
}

// ... the same from a row of BatchDev::row_emis (already scaled)
__device__ __forceinline__ void lane_set_emission_row(LaneCfg &L, const double *row) {
  const double4 e = *reinterpret_cast<const double4 *>(row);
  L.mu = e.x; L.ac = e.y; L.mc = e.z;
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
// FWD_ONLY (forward-only callers with wobble rows): the A-row needs only its upper band limit (below the band its
// inflow is already zero) and the B-row output only its lower one (above the band the consumer's own A-row limit cuts
// it off); lanes without an output use ms = INT_MAX.  Measured neutral on the SNP kernel, so currently unused.
// The step comes in two halves so that the latency-bound sweeps can evaluate the emission of step t+1 (which depends
// on nothing but the sample) while the state update of step t waits for its neighbour: lane_emit + lane_update.
template <bool BANK = false>
__device__ __forceinline__ void lane_emit(const LaneCfg &L, unsigned tab, double x, double &p, int &kk) {
#ifdef NVB_EXPERIMENT_NO_EMIT  // timing experiment only (wrong results): what a step costs without its emission
  p = 1.0 + 1e-9 * x; kk = 0;
  return;
#endif
  // own emission, reference formula ac - d*d*mc (kmer_model.cpp:47-51), in units of ln2/256
  const double d = x - L.mu;
  exp_ext_scaled<BANK>(fma(-(d * d), L.mc, L.ac), tab, p, kk);
}

template <int MEL, int MODE, bool WITH_JOIN, bool FWD_ONLY>
__device__ __forceinline__ void lane_update(const LaneCfg &L, LaneState<MEL> &S, int c, double p, int kk,
                                            const LaneOut &in, double sF, int sX, LaneOut &out, XD &aout);

template <int MEL, int MODE, bool WITH_JOIN, bool FWD_ONLY>
__device__ __forceinline__ void lane_step(const LaneCfg &L, LaneState<MEL> &S, unsigned tab, int c, double x,
                                          const LaneOut &in, double sF, int sX, LaneOut &out, XD &aout) {
  double p;
  int kk;
  lane_emit<true>(L, tab, x, p, kk);  // lane_step is the SNP kernel's form
  lane_update<MEL, MODE, WITH_JOIN, FWD_ONLY>(L, S, c, p, kk, in, sF, sX, out, aout);
}

template <int MEL, int MODE, bool WITH_JOIN, bool FWD_ONLY>
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
  // B[c] = e(c-1) * B[c-1] + (product of the last m emissions) * A[c-m], factored as
  //   B[c] = e(c-1) * (B[c-1] + Q_{m-1}[c-1]),  Q_i[c] = e(c-1) * Q_{i-1}[c-1],  Q_0[c] = A[c]:
  // the oldest value in flight joins the cell BEFORE the emission is applied, so a step costs m multiplications
  // instead of m + 1, and every in-flight value is computed straight into its next slot (no shifting moves: the
  // delay line of the unfactored form cost 6 register moves per step at m = 2).  S.q[0] = A of the previous step,
  // S.q[i] = Q_i.
  if (MEL == 0) {
    S.mod = xd_add(xd_make(pb * S.mod.f, S.mod.e + kb), push);
  } else {
    const XD joined = xd_add(S.mod, S.q[MEL - 1]);
    S.mod = xd_make(pb * joined.f, joined.e + kb);
#pragma unroll
    for (int i = MEL - 1; i > 0; i--) S.q[i] = xd_make(pb * S.q[i - 1].f, S.q[i - 1].e + kb);
    S.q[0] = push;
  }
  const bool inb = (FWD_ONLY && MODE == NVB_MODE_WOBBLE) ? (c >= L.ms) : ((c >= L.ms) && (c <= L.me));
  out.f = inb ? S.mod.f : 0.0;
  out.E = inb ? S.mod.e : NVB_EZERO;
}

// Mantissa renormalisation (call on a warp-uniform schedule: every NVB_RENORM_MASK + 1 steps in the sweeps, every 16 in
// the SNP kernel).
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
