// rows.cu -- forward / backward banded DP rows as a skewed wavefront (log-space fp64, sum semiring).
//
// Replaces the driver loops of RefineAlignment (reference nadavca/dtw/dtw.cpp:182-197) and EstimateLogLikelihoods
// (dtw.cpp:48-81) together with Node::NextRow (node_next_row.h:6-61).
//
// Mapping: one warp per (read, direction).  The rows of a pass form a "program" q = 0..Q-1 (q = 0 is the all-ones
// row).  A warp takes a stripe of up to 32 consecutive program rows, lane = row, and sweeps the columns: at step t
// lane l works on column C0 +/- (t - l).  Cell (row, c) needs the same row at the previous column (a register) and
// the previous row at column c -/+ m, which lane l-1 produced 1+m steps earlier; it is fetched from a small
// time-indexed ring in shared memory.  Lane 0 reads its predecessor (the last row of the previous stripe) from
// the matrix in HBM.  Each lane runs the plain recurrence
//     res[c] = (sum_{j<m} l(.) + pred[c -/+ m])  (+)  (l(.) + res[c -/+ 1])
// from the first column of the stripe; cells before the row's own band start are "virtual": they are carried but
// published as log(0), which reproduces the reference's first-cell sum (node_next_row.h:37-48 / :13-24).
#include "common.cuh"
#include "kernels.h"

namespace {

struct RowSpec {
  int s, e;       // inclusive band
  int m;          // minimum event length of this row's emission
  int stored;     // row is kept in the matrix
  int64_t off;    // packed offset of the row when stored
  Emis em;
};

template <bool REV>
__device__ __forceinline__ void program_row(int mode, int q, const ModelDev &M, const ReadView &v, int mel,
                                            RowSpec &r, bool want_emis) {
  const int n = v.n;
  int j;  // band row
  r.stored = 1;
  r.m = mel;
  r.em.kind = NVB_EM_CONST;
  r.em.ac1 = 0; r.em.mu1 = 0; r.em.mc1 = 0; r.em.mu2 = 0; r.em.ac2 = 0; r.em.mc2 = 0;
  if (!REV) {
    if (q == 0) {
      j = 0; r.off = 0; r.m = 0;
    } else if (mode == NVB_MODE_PLAIN) {  // dtw.cpp:176-179
      j = q; r.off = v.coff[q];
      if (want_emis) r.em = emis_gauss(M, v, q - 1, INT32_MIN, 0);
    } else if (mode == NVB_MODE_TRANS) {  // dtw.cpp:145-174
      int ei = q - 1;
      j = (q + 1) >> 1; r.off = trans_row_off(v, q);
      if (ei & 1) {
        r.m = 0;
        if (want_emis) r.em = emis_transition(M, v, ei >> 1, (ei >> 1) + 1);
      } else if (want_emis) {
        r.em = emis_gauss(M, v, ei >> 1, INT32_MIN, 0);
      }
    } else {  // wobble program, dtw.cpp:51-64
      if (q & 1) {
        int i = (q - 1) >> 1;
        j = i + 1; r.off = v.coff[i + 1];
        if (want_emis) r.em = emis_gauss(M, v, i, INT32_MIN, 0);
      } else {
        int i = q >> 1;
        j = i; r.off = 0; r.stored = 0; r.m = 0;
        if (want_emis) r.em = emis_mix(M, v, i - 1, i, INT32_MIN, 0);
      }
    }
  } else {
    if (q == 0) {
      j = n; r.m = 0;
      r.off = (mode == NVB_MODE_TRANS) ? trans_row_off(v, 2 * n - 1) : v.coff[n];
    } else if (mode == NVB_MODE_PLAIN) {
      int t = n - q;
      j = t; r.off = v.coff[t];
      if (want_emis) r.em = emis_gauss(M, v, t, INT32_MIN, 0);
    } else if (mode == NVB_MODE_TRANS) {
      int t = 2 * n - 1 - q;
      j = (t + 1) >> 1; r.off = trans_row_off(v, t);
      if (t & 1) {
        r.m = 0;
        if (want_emis) r.em = emis_transition(M, v, t >> 1, (t >> 1) + 1);
      } else if (want_emis) {
        r.em = emis_gauss(M, v, t >> 1, INT32_MIN, 0);
      }
    } else {  // dtw.cpp:68-81
      if (q & 1) {
        int i = n - ((q - 1) >> 1);
        j = i - 1; r.off = v.coff[i - 1];
        if (want_emis) r.em = emis_gauss(M, v, i - 1, INT32_MIN, 0);
      } else {
        int i = n - (q >> 1);
        j = i; r.off = 0; r.stored = 0; r.m = 0;
        if (want_emis) r.em = emis_mix(M, v, i, i - 1, INT32_MIN, 0);
      }
    }
  }
  r.s = v.bs[j];
  r.e = v.be[j];
}

template <bool REV>
__device__ void sweep(const ModelDev &M, const ReadView &v, int mode, int mel, double *mat, double *ring, int D,
                      int lane) {
  const int n = v.n, N = v.N;
  const int Q = (mode == NVB_MODE_PLAIN) ? n + 1 : 2 * n;
  const double NINF = nvb_neg_inf();
  const int dmask = D - 1;

  RowSpec pr;
  program_row<REV>(mode, 0, M, v, mel, pr, false);
  for (int c = pr.s + lane; c <= pr.e; c += NVB_WARP) mat[pr.off + c - pr.s] = 0.0;  // Node(start,end): all ones
  __syncwarp();

  int P = 0;
  while (P < Q - 1) {
    int L = min(NVB_WARP, Q - 1 - P);
    if (mode == NVB_MODE_WOBBLE && ((P + L) & 1) == 0) L -= 1;  // a stripe must end on a stored (odd) row
    const bool have = lane < L;
    RowSpec me;
    if (have) program_row<REV>(mode, P + 1 + lane, M, v, mel, me, true);
    else { me = pr; me.stored = 0; }
    const int C0 = REV ? pr.e : pr.s;
    const int endcol = __shfl_sync(NVB_FULL, REV ? me.s : me.e, L - 1);
    const int T = (REV ? C0 - endcol : endcol - C0) + L;
    const double *prow = mat + pr.off;

    for (int d = 0; d < D; d++) ring[d * NVB_WARP + lane] = NINF;
    __syncwarp();

    double cur = NINF;
    for (int t = 0; t < T; t++) {
      const int c = REV ? C0 - (t - lane) : C0 + (t - lane);
      const bool active = have && t >= lane && (REV ? c >= me.s : c <= me.e);
      double pub = NINF;
      if (active) {
        const int cp = REV ? c + me.m : c - me.m;
        double pv;
        if (lane == 0) {
          pv = (cp >= pr.s && cp <= pr.e) ? __ldcg(prow + (cp - pr.s)) : NINF;
        } else {
          const int tt = t - 1 - me.m;
          pv = tt >= 0 ? ring[(tt & dmask) * NVB_WARP + lane - 1] : NINF;
        }
        const int jb = REV ? c : c - 1;
        double p = 0.0;
        for (int k = 0; k < me.m; k++) {
          int j = REV ? jb + k : jb - k;
          j = min(max(j, 0), N - 1);
          p = p + emis_eval(me.em, __ldg(v.sig + j));
        }
        const double a = p + pv;
        const double b = emis_eval(me.em, __ldg(v.sig + min(max(jb, 0), N - 1))) + cur;
        cur = lp_add(a, b);
        const bool inband = REV ? (c <= me.e) : (c >= me.s);
        if (inband) {
          pub = cur;
          if (me.stored) mat[me.off + c - me.s] = cur;
        }
      }
      ring[(t & dmask) * NVB_WARP + lane] = pub;
      __syncwarp();
    }
    // the last row of this stripe is the predecessor of the next one
    pr.s = __shfl_sync(NVB_FULL, me.s, L - 1);
    pr.e = __shfl_sync(NVB_FULL, me.e, L - 1);
    pr.off = __shfl_sync(NVB_FULL, me.off, L - 1);
    P += L;
    __syncwarp();
  }
}

__global__ void __launch_bounds__(128) sweep_kernel(ModelDev M, BatchDev B, int mode, int b0, int n_items,
                                                    const int64_t *mat_base, double *prefix, double *suffix,
                                                    int D) {
  extern __shared__ double s_ring[];
  const int wic = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int item = blockIdx.x * (blockDim.x >> 5) + wic;
  if (item >= n_items) return;
  const int b = b0 + (item >> 1);
  const bool rev = item & 1;
  if (B.flags[b]) return;
  ReadView v = read_view(B, b);
  double *ring = s_ring + (size_t)wic * D * NVB_WARP;
  if (rev) sweep<true>(M, v, mode, B.mel, suffix + mat_base[b], ring, D, lane);
  else sweep<false>(M, v, mode, B.mel, prefix + mat_base[b], ring, D, lane);
}

// Node::TotalLikelihood(prefix[n], suffix[n]) (dtw.cpp:83-85, node.cpp:31-37); suffix[n] is all ones.
__global__ void __launch_bounds__(128) no_snp_kernel(ModelDev M, BatchDev B, int b0, int n_items,
                                                     const int64_t *mat_base, const double *prefix,
                                                     const double *suffix, double *out_ll) {
  const int wic = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int item = blockIdx.x * (blockDim.x >> 5) + wic;
  if (item >= n_items) return;
  const int b = b0 + item;
  if (B.flags[b]) return;
  ReadView v = read_view(B, b);
  const int n = v.n;
  const double *prow = prefix + mat_base[b] + v.coff[n];
  const double *srow = suffix + mat_base[b] + v.coff[n];
  const int w = v.be[n] - v.bs[n] + 1;
  const int seg = (w + NVB_WARP - 1) / NVB_WARP;
  const int lo = min(w, lane * seg), hi = min(w, lo + seg);
  double acc = nvb_neg_inf();
  for (int i = lo; i < hi; i++) acc = lp_add(acc, __ldcg(prow + i) + __ldcg(srow + i));
  double total = nvb_neg_inf();
  for (int l = 0; l < NVB_WARP; l++) total = lp_add(total, __shfl_sync(NVB_FULL, acc, l));
  const int A = M.alphabet;
  double *out = out_ll + B.ref_off[b] * A;
  for (int i = lane; i < n; i += NVB_WARP) out[(int64_t)i * A + v.ref[i]] = total;
}

}  // namespace

static int ring_depth(int mel) {
  int D = 4;
  while (D < mel + 2) D <<= 1;
  return D;
}

void nvbk_sweep(const ModelDev &M, const BatchDev &B, int mode, int b0, int b1, const int64_t *d_mat_base,
                double *d_prefix, double *d_suffix, cudaStream_t st) {
  const int n_items = 2 * (b1 - b0);
  if (n_items <= 0) return;
  const int D = ring_depth(B.mel);
  const int warps = 4;
  size_t smem = (size_t)warps * D * NVB_WARP * sizeof(double);
  if (smem > 48 * 1024) cudaFuncSetAttribute(sweep_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  sweep_kernel<<<(n_items + warps - 1) / warps, warps * NVB_WARP, smem, st>>>(M, B, mode, b0, n_items, d_mat_base,
                                                                            d_prefix, d_suffix, D);
}

void nvbk_no_snp(const ModelDev &M, const BatchDev &B, int b0, int b1, const int64_t *d_mat_base,
                 const double *d_prefix, const double *d_suffix, double *d_out_ll, cudaStream_t st) {
  const int n_items = b1 - b0;
  if (n_items <= 0) return;
  no_snp_kernel<<<(n_items + 3) / 4, 128, 0, st>>>(M, B, b0, n_items, d_mat_base, d_prefix, d_suffix, d_out_ll);
}
