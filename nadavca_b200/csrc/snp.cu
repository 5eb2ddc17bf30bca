// snp.cu -- the per-position SNP refinement DP (the hottest loop of the reference: dtw.cpp:93-129).
//
// Work item ("task") = (read, reference position i, alternative base): re-run the <= k model rows (and their
// wobble rows) whose k-mer contains position i, starting from the stored prefix row `first` and closing against
// the stored suffix row `last+1` with Node::TotalLikelihood (node.cpp:31-37).  A task is a short row program
//     [wobble(j-1,j) on band j]  model(j) on band j+1   for j = first..last
//     [trailing wobble(last,last+1) on band `last`]      (dtw.cpp:116-123 -- band row `last`, not last+1)
//     join with suffix[last+1]
// of at most 2k+2 rows; lane = row, the same skewed wavefront as rows.cu, several tasks packed per warp.  The
// join is itself a row: acc[c] = acc[c-1] (+) (cur[c] + suffix[c]), i.e. m = 0 and a unit emission.
#include "common.cuh"
#include "kernels.h"

namespace {

enum { ROW_NONE = 0, ROW_WOB = 1, ROW_MOD = 2, ROW_TRAIL = 3, ROW_JOIN = 4 };

__global__ void __launch_bounds__(128) snp_kernel(ModelDev M, BatchDev B, int wobbling, int b0, int b1, int64_t g0,
                                                  int64_t n_tasks, int LT, int TPW, const int64_t *mat_base,
                                                  const double *prefix, const double *suffix, double *out_ll,
                                                  int D) {
  extern __shared__ double s_ring[];
  const int wic = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t warp_id = blockIdx.x * (int64_t)(blockDim.x >> 5) + wic;
  double *ring = s_ring + (size_t)wic * D * NVB_WARP;
  const double NINF = nvb_neg_inf();
  const int dmask = D - 1;
  const int A = M.alphabet;

  const int slot = lane / LT, ri = lane - slot * LT;
  const int64_t task = warp_id * TPW + slot;
  bool have = slot < TPW && task < n_tasks;

  // ---- decode the task and this lane's row -------------------------------------------------------------
  int kind = ROW_NONE, s = 0, e = -1, m = 0, C0 = 0, nrows = 0, N = 1;
  Emis em;
  em.kind = NVB_EM_CONST; em.ac1 = 0; em.mu1 = 0; em.mc1 = 0; em.mu2 = 0; em.ac2 = 0; em.mc2 = 0;
  const double *prow = nullptr, *srow = nullptr, *sig = nullptr;
  int ps = 0, pe = -1, ss = 0, se = -1;
  double *out = nullptr;
  if (have) {
    const int64_t g = g0 + task / (A - 1);
    const int alt = (int)(task % (A - 1));
    int lo = b0, hi = b1;  // last read with ref_off <= g
    while (hi - lo > 1) {
      int mid = (lo + hi) >> 1;
      if (B.ref_off[mid] <= g) lo = mid; else hi = mid;
    }
    const int b = lo;
    if (B.flags[b]) {
      have = false;
    } else {
      ReadView v = read_view(B, b);
      const int n = v.n;
      const int i = (int)(g - B.ref_off[b]);
      const int refbase = v.ref[i];
      const int base = alt + (alt >= refbase ? 1 : 0);
      const int back = M.k - M.central - 1, fwd = M.central;  // dtw.cpp:88-89
      const int first = max(0, i - back), last = min(n - 1, i + fwd);
      N = v.N; sig = v.sig;
      int idx = 0, prev_band = first;
      for (int j = first; j <= last; j++) {
        if (wobbling && j > 0) {
          if (idx == ri) { kind = ROW_WOB; s = v.bs[j]; e = v.be[j]; m = 0; em = emis_mix(M, v, j - 1, j, i, base); }
          idx++;
        }
        if (idx == ri) { kind = ROW_MOD; s = v.bs[j + 1]; e = v.be[j + 1]; m = B.mel; em = emis_gauss(M, v, j, i, base); }
        idx++;
        prev_band = j + 1;
      }
      if (wobbling && last + 1 < n) {
        if (idx == ri) { kind = ROW_TRAIL; s = v.bs[last]; e = v.be[last]; m = 0; em = emis_mix(M, v, last, last + 1, i, base); }
        idx++;
        prev_band = last;
      }
      if (idx == ri) {
        kind = ROW_JOIN; s = v.bs[prev_band]; e = v.be[prev_band]; m = 0;
        srow = suffix + mat_base[b] + v.coff[last + 1];
        ss = v.bs[last + 1]; se = v.be[last + 1];
        out = out_ll + (B.ref_off[b] + i) * A + base;
      }
      idx++;
      nrows = idx;
      C0 = v.bs[first];
      if (ri == 0) { prow = prefix + mat_base[b] + v.coff[first]; ps = v.bs[first]; pe = v.be[first]; }
      if (kind == ROW_NONE) have = false;
    }
  }
  int T = have ? (e - C0 + ri + 1) : 0;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) T = max(T, __shfl_xor_sync(NVB_FULL, T, o));

  for (int d = 0; d < D; d++) ring[d * NVB_WARP + lane] = NINF;
  __syncwarp();

  double cur = NINF;
  for (int t = 0; t < T; t++) {
    const int c = C0 + (t - ri);
    const bool active = have && t >= ri && c <= e;
    double pub = NINF;
    if (active) {
      const int cp = c - m;
      double pv;
      if (ri == 0) {
        pv = (cp >= ps && cp <= pe) ? __ldg(prow + (cp - ps)) : NINF;
      } else {
        const int tt = t - 1 - m;
        pv = tt >= 0 ? ring[(tt & dmask) * NVB_WARP + lane - 1] : NINF;
      }
      double b;
      if (kind == ROW_JOIN) {
        const double sv = (c >= ss && c <= se) ? __ldg(srow + (c - ss)) : NINF;
        pv = pv + sv;  // prefix[i] * suffix[i]
        b = cur;
      } else {
        b = emis_eval(em, __ldg(sig + min(max(c - 1, 0), N - 1))) + cur;
      }
      double p = 0.0;
      for (int k = 0; k < m; k++) p = p + emis_eval(em, __ldg(sig + min(max(c - 1 - k, 0), N - 1)));
      cur = lp_add(p + pv, b);
      if (c >= s) pub = cur;
    }
    ring[(t & dmask) * NVB_WARP + lane] = pub;
    __syncwarp();
  }
  if (have && kind == ROW_JOIN) *out = cur;
  (void)nrows;
}

}  // namespace

static int ring_depth(int mel) {
  int D = 4;
  while (D < mel + 2) D <<= 1;
  return D;
}

int nvbk_snp(const ModelDev &M, const BatchDev &B, int wobbling, int b0, int b1, int64_t g0, int64_t g1,
             const int64_t *d_mat_base, const double *d_prefix, const double *d_suffix, double *d_out_ll,
             cudaStream_t st) {
  const int LT = wobbling ? 2 * M.k + 2 : M.k + 1;  // rows of the longest task
  if (LT > NVB_WARP) return -1;
  const int TPW = NVB_WARP / LT;
  const int64_t n_tasks = (g1 - g0) * (M.alphabet - 1);
  if (n_tasks <= 0) return 0;
  const int warps = 4;
  const int64_t n_warps = (n_tasks + TPW - 1) / TPW;
  const int D = ring_depth(B.mel);
  size_t smem = (size_t)warps * D * NVB_WARP * sizeof(double);
  if (smem > 48 * 1024) cudaFuncSetAttribute(snp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  snp_kernel<<<(unsigned)((n_warps + warps - 1) / warps), warps * NVB_WARP, smem, st>>>(
      M, B, wobbling, b0, b1, g0, n_tasks, LT, TPW, d_mat_base, d_prefix, d_suffix, d_out_ll, D);
  return 0;
}
