// rows3.cu -- forward / backward banded DP rows, scaled linear domain (see dp3.cuh), one warp per (read, direction).
//
// Replaces the driver loops of RefineAlignment (reference nadavca/dtw/dtw.cpp:182-197) and EstimateLogLikelihoods
// (dtw.cpp:48-81) with Node::NextRow (node_next_row.h:6-61).  The rows of a pass are grouped in pairs
// (A-row, B-row): A = the wobble row (dtw.cpp:53-58) or the transition row (dtw.cpp:170-172) in front of base i,
// B = the model row of base i.  A stripe is one LOADER lane (re-reads the last stored row of the previous stripe
// from HBM and supplies the emission of the base before the stripe) plus up to 31 pair lanes; lane l works one
// step behind lane l-1.  Stored rows are written as (mantissa double, exponent int32) planes.
#include <stdio.h>
#include "dp3.cuh"
#include "kernels.h"

namespace {

struct PairGeom {
  int hasA, aband, bband, storeA;
  int64_t aoff, boff;
  int base;  // reference base index of the model row
  int nb;    // base index of the neighbouring model row (the other component of the wobble mixture)
};

template <bool REV>
__device__ __forceinline__ PairGeom pair_geom(const ReadView &v, int mode, int g) {
  PairGeom p;
  const int n = v.n;
  const int i = REV ? n - 1 - g : g;
  p.base = i;
  p.storeA = (mode == NVB_MODE_TRANS);
  if (!REV) {
    p.hasA = (i >= 1) && mode != NVB_MODE_PLAIN;
    p.aband = i; p.bband = i + 1; p.nb = i - 1;
    if (mode == NVB_MODE_TRANS) { p.aoff = trans_row_off(v, 2 * i); p.boff = trans_row_off(v, 2 * i + 1); }
    else { p.aoff = 0; p.boff = v.coff[i + 1]; }
  } else {
    p.hasA = (i <= n - 2) && mode != NVB_MODE_PLAIN;
    p.aband = i + 1; p.bband = i; p.nb = i + 1;
    if (mode == NVB_MODE_TRANS) { p.aoff = trans_row_off(v, 2 * i + 1); p.boff = trans_row_off(v, 2 * i); }
    else { p.aoff = 0; p.boff = v.coff[i]; }
  }
  return p;
}

template <int MEL, int MODE, bool REV>
__device__ void sweep2(const ModelDev &M, const ReadView &v, double *F, int32_t *X, int lane) {
  constexpr int mode = MODE;
  const int n = v.n, N = v.N;
  const double C_E2 = 0.1353352832366127;  // exp(-2): the "/ 2" of kmer_model.cpp:60 is "- 2.0" in log space

  // all-ones first row (dtw.cpp:50,66,182,190)
  int ls, le;           // band of the row the loader re-reads
  int64_t loff;
  {
    const int j0 = REV ? n : 0;
    ls = v.bs[j0]; le = v.be[j0];
    loff = REV ? ((mode == NVB_MODE_TRANS) ? trans_row_off(v, 2 * n - 1) : v.coff[n]) : 0;
    for (int c = ls + lane; c <= le; c += NVB_WARP) { F[loff + c - ls] = 1.0; X[loff + c - ls] = 0; }
  }
  __syncwarp();

  for (int g0 = 0; g0 < n; g0 += NVB_WARP - 1) {
    const int npairs = min(NVB_WARP - 1, n - g0);
    LaneCfg L;
    lane_cfg_clear(L);
    int ws = 1, awe = 0;  // A-row band, for the stores of the transition rows
    int64_t aoff = 0, boff = 0;
    int storeA = 0, storeB = 0;
    if (lane == 0) {
      L.role = NVB_ROLE_LOADER;
      L.ms = ls; L.me = le;
      const PairGeom pg = pair_geom<REV>(v, mode, g0);
      const int nb = (g0 > 0) ? pg.nb : pg.base;  // emission of the model row before the stripe
      const int id = kmer_id(M, v, nb, INT32_MIN, 0);
      L.mu = M.mean[id]; L.ac = M.ac[id]; L.mc = M.mc[id];
    } else if (lane <= npairs) {
      const PairGeom pg = pair_geom<REV>(v, mode, g0 + lane - 1);
      const int id = kmer_id(M, v, pg.base, INT32_MIN, 0);
      L.role = NVB_ROLE_PAIR;
      L.mu = M.mean[id]; L.ac = M.ac[id]; L.mc = M.mc[id];
      ws = v.bs[pg.aband]; awe = v.be[pg.aband];
      L.ms = v.bs[pg.bband]; L.me = v.be[pg.bband];
      aoff = pg.aoff; boff = pg.boff; storeA = pg.storeA && pg.hasA; storeB = 1;
      if (pg.hasA) {
        L.ws = ws; L.we = awe;
        if (mode == NVB_MODE_TRANS) {  // GetTransitionDistribution (kmer_model.cpp:64-94): constant 0.01, or 0
          const double mo = M.mean[kmer_id(M, v, pg.nb, INT32_MIN, 0)];
          const bool dead = (mo == L.mu);
          L.pc = dead ? 0.0 : 0.01 * 64.0;  // 0.01 as mantissa 0.64 and exponent -6 (an exact rescaling)
          L.kc = dead ? NVB_EZERO : -6;
        } else {
          L.cm = C_E2; L.abias = 0;
        }
      }
    }
    const int C0 = REV ? le : ls;
    const int endcol = __shfl_sync(NVB_FULL, REV ? L.ms : L.me, npairs);
    const int T = (REV ? C0 - endcol : endcol - C0) + npairs + 1;
    const double *lF = F + loff;
    const int32_t *lX = X + loff;

    LaneState<MEL> S;
    lane_reset(S);
    LaneOut out;
    out.f = 0.0; out.E = NVB_EZERO; out.p = 1.0; out.k = 0;
    for (int t = 0; t < T; t++) {
      const int c = REV ? C0 - (t - lane) : C0 + (t - lane);
      const int xi = min(max(REV ? c : c - 1, 0), N - 1);
      const double x = __ldg(v.sig + xi);
      LaneOut in = shfl_up_out<MODE>(out);
      XD aout;
      lane_step<MEL, MODE, false>(L, S, c, x, in, 1.0, 0, out, aout);
      if (lane == 0) {
        const bool inb = (c >= L.ms && c <= L.me);
        out.f = inb ? __ldcg(lF + (c - L.ms)) : 0.0;
        out.E = inb ? __ldcg(lX + (c - L.ms)) : NVB_EZERO;
      } else {
        if (storeB && c >= L.ms && c <= L.me) { F[boff + c - L.ms] = out.f; X[boff + c - L.ms] = out.E; }
        if (storeA && c >= ws && c <= awe) { F[aoff + c - ws] = aout.f; X[aoff + c - ws] = aout.e; }
      }
      if ((t & NVB_RENORM_MASK) == NVB_RENORM_MASK) lane_renorm(S);
    }
    // the last pair's B-row feeds the next stripe
    ls = __shfl_sync(NVB_FULL, L.ms, npairs);
    le = __shfl_sync(NVB_FULL, L.me, npairs);
    loff = __shfl_sync(NVB_FULL, boff, npairs);
    __syncwarp();
  }
}

template <int MEL, int MODE>
__global__ void __launch_bounds__(128) sweep3_kernel(ModelDev M, BatchDev B, int b0, int n_items,
                                                     const int64_t *mat_base, double *pF, int32_t *pX, double *sF,
                                                     int32_t *sX) {
  const int wic = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int item = blockIdx.x * (blockDim.x >> 5) + wic;
  if (item >= n_items) return;
  const int b = b0 + (item >> 1);
  if (B.flags[b]) return;
  ReadView v = read_view(B, b);
  const int64_t base = mat_base[b];
  if (item & 1) sweep2<MEL, MODE, true>(M, v, sF + base, sX + base, lane);
  else sweep2<MEL, MODE, false>(M, v, pF + base, pX + base, lane);
}

// Node::TotalLikelihood(prefix[n], suffix[n]) (dtw.cpp:83-85); suffix[n] is all ones.  Two passes over the row:
// largest exponent, then the mantissa sum relative to it.
__global__ void __launch_bounds__(128) no_snp2_kernel(ModelDev M, BatchDev B, int b0, int n_items,
                                                      const int64_t *mat_base, const double *pF, const int32_t *pX,
                                                      const double *sF, const int32_t *sX, double *out_ll) {
  const int wic = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int item = blockIdx.x * (blockDim.x >> 5) + wic;
  if (item >= n_items) return;
  const int b = b0 + item;
  if (B.flags[b]) return;
  ReadView v = read_view(B, b);
  const int n = v.n;
  const int64_t off = mat_base[b] + v.coff[n];
  const int w = v.be[n] - v.bs[n] + 1;
  int emax = NVB_EZERO;
  for (int i = lane; i < w; i += NVB_WARP) {
    const double f = pF[off + i] * sF[off + i];
    if (f > 0.0) emax = max(emax, pX[off + i] + sX[off + i]);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) emax = max(emax, __shfl_xor_sync(NVB_FULL, emax, o));
  double acc = 0.0;
  for (int i = lane; i < w; i += NVB_WARP) {
    const double f = pF[off + i] * sF[off + i];
    if (f > 0.0) acc += f * pow2neg(pX[off + i] + sX[off + i] - emax);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(NVB_FULL, acc, o);
  const double total = log_ext(acc, emax);
  const int A = M.alphabet;
  double *out = out_ll + B.ref_off[b] * A;
  for (int i = lane; i < n; i += NVB_WARP) out[(int64_t)i * A + v.ref[i]] = total;
}

template <int MEL>
void launch_sweep2(const ModelDev &M, const BatchDev &B, int mode, int b0, int n_items, const int64_t *mb, double *pF,
                   int32_t *pX, double *sF, int32_t *sX, cudaStream_t st) {
  const unsigned grid = (n_items + 3) / 4;
  switch (mode) {
    case NVB_MODE_PLAIN: sweep3_kernel<MEL, NVB_MODE_PLAIN><<<grid, 128, 0, st>>>(M, B, b0, n_items, mb, pF, pX, sF, sX); break;
    case NVB_MODE_TRANS: sweep3_kernel<MEL, NVB_MODE_TRANS><<<grid, 128, 0, st>>>(M, B, b0, n_items, mb, pF, pX, sF, sX); break;
    default: sweep3_kernel<MEL, NVB_MODE_WOBBLE><<<grid, 128, 0, st>>>(M, B, b0, n_items, mb, pF, pX, sF, sX); break;
  }
}

}  // namespace

int nvbk_sweep2(const ModelDev &M, const BatchDev &B, int mode, int b0, int b1, const int64_t *d_mat_base,
                double *pF, int32_t *pX, double *sF, int32_t *sX, cudaStream_t st) {
  const int n_items = 2 * (b1 - b0);
  if (n_items <= 0) return 0;
  switch (B.mel) {
    case 0: launch_sweep2<0>(M, B, mode, b0, n_items, d_mat_base, pF, pX, sF, sX, st); break;
    case 1: launch_sweep2<1>(M, B, mode, b0, n_items, d_mat_base, pF, pX, sF, sX, st); break;
    case 2: launch_sweep2<2>(M, B, mode, b0, n_items, d_mat_base, pF, pX, sF, sX, st); break;
    case 3: launch_sweep2<3>(M, B, mode, b0, n_items, d_mat_base, pF, pX, sF, sX, st); break;
    case 4: launch_sweep2<4>(M, B, mode, b0, n_items, d_mat_base, pF, pX, sF, sX, st); break;
    case 5: launch_sweep2<5>(M, B, mode, b0, n_items, d_mat_base, pF, pX, sF, sX, st); break;
    case 6: launch_sweep2<6>(M, B, mode, b0, n_items, d_mat_base, pF, pX, sF, sX, st); break;
    default: return -1;
  }
  return 0;
}

void nvbk_no_snp2(const ModelDev &M, const BatchDev &B, int b0, int b1, const int64_t *d_mat_base, const double *pF,
                  const int32_t *pX, const double *sF, const int32_t *sX, double *d_out_ll, cudaStream_t st) {
  const int n_items = b1 - b0;
  if (n_items <= 0) return;
  no_snp2_kernel<<<(n_items + 3) / 4, 128, 0, st>>>(M, B, b0, n_items, d_mat_base, pF, pX, sF, sX, d_out_ll);
}
