// dp2.cuh -- the scaled linear-domain DP cell shared by the row sweep (rows2.cu) and the SNP kernel (snp2.cu).
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
// range); adding the two terms of a recurrence aligns them on the larger exponent, so a term is dropped only when it
// is below 2^-1022 of the other term OF THE SAME CELL -- the reference itself drops it below e^-37
// (log(1 + exp(b - a)) == 0).  Zero is (0, NVB_EZERO).  Mantissas are renormalised every 32 steps.
#pragma once
#include "common.cuh"

#define NVB_EZERO (-(1 << 30))  // exponent carried by lanes / cells whose value is exactly zero

#define NVB_LN2_HI 6.93147180369123816490e-01
#define NVB_LN2_LO 1.90821492927058770002e-10
#define NVB_LOG2E 1.44269504088896338700e+00
#define NVB_LN2 0.693147180559945309417

// 2^e as a double; e <= -1023 gives +0.0, e must be <= 1023
__device__ __forceinline__ double pow2i(int e) {
  e = max(e, -1023);
  return __hiloint2double((e + 1023) << 20, 0);
}

// exp(l) = p * 2^k with p in [0.70, 1.42], any finite l (no underflow: k is returned, not applied).
// Cody-Waite reduction r = l - k*ln2 (hi/lo) and a degree-13 Taylor polynomial on |r| <= 0.347 (error < 5e-18).
__device__ __forceinline__ void exp_ext(double l, double &p, int &k) {
  const double magic = 6755399441055744.0;  // 1.5 * 2^52: rint via add/sub, integer in the low word
  double t = fma(l, NVB_LOG2E, magic);
  k = __double2loint(t);
  double kd = t - magic;
  double r = fma(-kd, NVB_LN2_HI, l);
  r = fma(-kd, NVB_LN2_LO, r);
  double q = 1.6059043836821613e-10;              // 1/13!
  q = fma(q, r, 2.08767569878681e-09);            // 1/12!
  q = fma(q, r, 2.505210838544172e-08);           // 1/11!
  q = fma(q, r, 2.755731922398589e-07);           // 1/10!
  q = fma(q, r, 2.7557319223985893e-06);          // 1/9!
  q = fma(q, r, 2.48015873015873e-05);            // 1/8!
  q = fma(q, r, 1.984126984126984e-04);           // 1/7!
  q = fma(q, r, 1.388888888888889e-03);           // 1/6!
  q = fma(q, r, 8.333333333333333e-03);           // 1/5!
  q = fma(q, r, 4.1666666666666664e-02);          // 1/4!
  q = fma(q, r, 1.6666666666666666e-01);          // 1/3!
  q = fma(q, r, 0.5);
  q = fma(q, r, 1.0);
  p = fma(q, r, 1.0);
}

// natural log of f * 2^E (f > 0), E*ln2 added in two pieces
__device__ __forceinline__ double log_ext(double f, int E) {
  if (!(f > 0.0)) return nvb_neg_inf();
  return fma((double)E, NVB_LN2_HI, log(f)) + (double)E * NVB_LN2_LO;
}

enum { NVB_ROLE_IDLE = 0, NVB_ROLE_LOADER = 1, NVB_ROLE_PAIR = 2, NVB_ROLE_JOIN = 3 };

// value = f * 2^e; zero is (0, NVB_EZERO)
struct XD {
  double f;
  int e;
};

__device__ __forceinline__ XD xd_zero() {
  XD z;
  z.f = 0.0; z.e = NVB_EZERO;
  return z;
}

// a + b, aligned on the larger exponent
__device__ __forceinline__ XD xd_add(XD a, XD b) {
  const int d = b.e - a.e;
  const bool bb = d > 0;
  const double s = pow2i(-abs(d));
  XD r;
  r.f = fma(bb ? a.f : b.f, s, bb ? b.f : a.f);
  r.e = max(a.e, b.e);
  return r;
}

// mantissa back into [1,2) (normal inputs only); zero gets the zero exponent
__device__ __forceinline__ void xd_renorm(XD &v) {
  if (v.f > 0.0) {
    const int hi = __double2hiint(v.f);
    v.e += ((hi >> 20) & 0x7ff) - 1023;
    v.f = __hiloint2double((hi & 0x800fffff) | 0x3ff00000, __double2loint(v.f));
  } else {
    v.e = NVB_EZERO;
  }
}

// Per-lane constants of one stripe / task.
struct LaneCfg {
  int role;
  int hasA;          // the lane has an A-row (wobble / transition row)
  int ws, we;        // A-row band (inclusive)
  int ms, me;        // B-row band; LOADER: band of the row it loads; JOIN: band of the closing suffix row
  double mu, ac, mc; // own Gaussian emission (PAIR: model row; LOADER: the row before the first pair; JOIN: last+1)
  double a1, a2;     // A-row emission = a1 * neighbour emission + a2 * own emission
  double pc;         // constant neighbour emission (transition rows), used when nb_const
  int kc;
  int nb_const;
};

template <int MEL>
struct LaneState {
  XD mod, w;                    // running B-row / A-row cells
  XD q[MEL > 0 ? MEL : 1];      // A-row outputs in flight to the B-row (delay = min_event_length)
  XD acc;                       // JOIN accumulator
};

template <int MEL>
__device__ __forceinline__ void lane_reset(LaneState<MEL> &S) {
  S.mod = xd_zero(); S.w = xd_zero(); S.acc = xd_zero();
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
// `in` the neighbour's output of the previous step, (sF, sX) the closing suffix cell (JOIN only).  `aout` receives
// the A-row cell (for callers that store it).
template <int MEL>
__device__ __forceinline__ void lane_step(const LaneCfg &L, LaneState<MEL> &S, int c, double x, LaneOut in,
                                          double sF, int sX, LaneOut &out, XD &aout) {
  // own emission, reference formula ac - d*d*mc (kmer_model.cpp:47-51)
  const double d = x - L.mu;
  const double l = L.ac - d * d * L.mc;
  double p;
  int kk;
  exp_ext(l, p, kk);
  out.p = p;
  out.k = kk;

  XD pm;
  pm.f = in.f; pm.e = in.E;

  XD wout;
  if (L.hasA) {
    // A[c] = P[c] + (a1 * e_nb + a2 * e_own) * A[c-1]; the two emissions are aligned on the larger exponent
    double np = in.p;
    int nk = in.k;
    if (L.nb_const) { np = L.pc; nk = L.kc; }
    const int dk = nk - kk;
    const int kref = max(kk, nk);
    const double et = p * pow2i(-max(dk, 0));
    const double en = np * pow2i(-max(-dk, 0));
    const double mix = fma(et, L.a2, en * L.a1);
    XD t;
    t.f = mix * S.w.f;
    t.e = (t.f == 0.0) ? NVB_EZERO : S.w.e + kref;  // a zero must not carry a live exponent into the alignment
    S.w = xd_add(t, pm);
    if (S.w.f == 0.0) S.w.e = NVB_EZERO;
    wout = (c >= L.ws && c <= L.we) ? S.w : xd_zero();
  } else {
    wout = pm;
  }
  aout = wout;

  if (L.role == NVB_ROLE_JOIN) {
    // Node::TotalLikelihood (node.cpp:31-37): acc += A[c] * suffix[c]
    XD t;
    t.f = wout.f * sF;
    t.e = wout.e + sX;
    if (t.f != 0.0) S.acc = xd_add(S.acc, t);
    out.f = 0.0;
    out.E = NVB_EZERO;
  } else {
    // B[c] = e(c-1) * B[c-1] + (product of the last m emissions) * A[c-m]
    XD popped;
    if (MEL == 0) {
      popped = wout;
    } else {
#pragma unroll
      for (int i = 0; i < MEL; i++) { S.q[i].f *= p; S.q[i].e += kk; }
      popped = S.q[MEL - 1];
#pragma unroll
      for (int i = MEL - 1; i > 0; i--) S.q[i] = S.q[i - 1];
      S.q[0] = wout;
    }
    XD t;
    t.f = p * S.mod.f;
    t.e = S.mod.e + kk;
    S.mod = xd_add(t, popped);
    // the B-row cell is what other lanes, later stripes and other kernels consume: keep it normalised at every step
    // (an un-normalised inflow would hand its exponent slack to everything aligned against it)
    xd_renorm(S.mod);
    const bool inb = (c >= L.ms && c <= L.me);
    out.f = inb ? S.mod.f : 0.0;
    out.E = inb ? S.mod.e : NVB_EZERO;
  }
}

// Mantissa renormalisation (call on a warp-uniform schedule, e.g. every 32 steps).
template <int MEL>
__device__ __forceinline__ void lane_renorm(LaneState<MEL> &S) {
  xd_renorm(S.w);
  xd_renorm(S.acc);
#pragma unroll
  for (int i = 0; i < (MEL > 0 ? MEL : 1); i++) xd_renorm(S.q[i]);
}

__device__ __forceinline__ LaneOut shfl_up_out(const LaneOut &o) {
  LaneOut r;
  r.f = __shfl_up_sync(NVB_FULL, o.f, 1);
  r.E = __shfl_up_sync(NVB_FULL, o.E, 1);
  r.p = __shfl_up_sync(NVB_FULL, o.p, 1);
  r.k = __shfl_up_sync(NVB_FULL, o.k, 1);
  return r;
}
