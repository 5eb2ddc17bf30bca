// snp2.cu -- the per-position SNP refinement DP (reference dtw.cpp:93-129), scaled linear domain (dp2.cuh).
//
// Task = (read, position i, alternative base).  Lanes of a task: LOADER (streams the stored prefix row `first` from
// HBM and evaluates the emission of base first-1), one PAIR lane per influenced base j = first..last (wobble row on
// band j + model row on band j+1 under the modified sequence, dtw.cpp:103-115), JOIN (the trailing wobble row on
// band `last` -- dtw.cpp:116-123 -- and Node::TotalLikelihood against the stored suffix row last+1).  That is k+2
// lanes; 32/(k+2) tasks share a warp (4 for the 6-mer model).  Every lane evaluates exactly one exp per step.
#include "dp2.cuh"
#include "kernels.h"

namespace {

template <int MEL>
__global__ void __launch_bounds__(128) snp2_kernel(ModelDev M, BatchDev B, int wobbling, int b0, int b1, int64_t g0,
                                                   int64_t n_tasks, int LT, int TPW, const int64_t *mat_base,
                                                   const double *pF, const int32_t *pX, const double *sF,
                                                   const int32_t *sX, double *out_ll) {
  const int wic = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t warp_id = blockIdx.x * (int64_t)(blockDim.x >> 5) + wic;
  const int A = M.alphabet;
  const double C_E2 = 0.1353352832366127;  // exp(-2)

  const int slot = lane / LT, ri = lane - slot * LT;
  const int64_t task = warp_id * TPW + slot;
  const bool have = slot < TPW && task < n_tasks;

  LaneCfg L;
  L.role = NVB_ROLE_IDLE; L.hasA = 0; L.ws = 1; L.we = 0; L.ms = 1; L.me = 0;
  L.mu = 0; L.ac = 0; L.mc = 0; L.a1 = C_E2; L.a2 = C_E2; L.pc = 0; L.kc = 0; L.nb_const = 0;
  const double *rowF = nullptr;   // LOADER: prefix row; JOIN: suffix row
  const int32_t *rowX = nullptr;
  const double *sig = B.signal;
  int N = 1, C0 = 0, T = 0;
  double *out = nullptr;

  if (have) {
    const int64_t g = g0 + task / (A - 1);
    const int alt = (int)(task % (A - 1));
    int lo = b0, hi = b1;  // last read with ref_off <= g
    while (hi - lo > 1) {
      const int mid = (lo + hi) >> 1;
      if (B.ref_off[mid] <= g) lo = mid; else hi = mid;
    }
    const int b = lo;
    if (!B.flags[b]) {
      ReadView v = read_view(B, b);
      const int n = v.n;
      const int i = (int)(g - B.ref_off[b]);
      const int refbase = v.ref[i];
      const int base = alt + (alt >= refbase ? 1 : 0);
      const int back = M.k - M.central - 1, fwd = M.central;  // dtw.cpp:88-89
      const int first = max(0, i - back), last = min(n - 1, i + fwd);
      const int npairs = last - first + 1;
      N = v.N; sig = v.sig;
      C0 = v.bs[first];
      const int64_t mb = mat_base[b];
      int kmer_at = -1;
      if (ri == 0) {
        L.role = NVB_ROLE_LOADER;
        L.ms = v.bs[first]; L.me = v.be[first];
        rowF = pF + mb + v.coff[first]; rowX = pX + mb + v.coff[first];
        kmer_at = first > 0 ? first - 1 : first;
      } else if (ri <= npairs) {
        const int j = first + ri - 1;
        L.role = NVB_ROLE_PAIR;
        L.hasA = wobbling && j > 0;
        L.ws = v.bs[j]; L.we = v.be[j];
        L.ms = v.bs[j + 1]; L.me = v.be[j + 1];
        kmer_at = j;
      } else if (ri == npairs + 1) {
        L.role = NVB_ROLE_JOIN;
        L.hasA = wobbling && last + 1 < n;
        L.ws = v.bs[last]; L.we = v.be[last];          // band row `last`, not last+1 (dtw.cpp:120-121)
        L.ms = v.bs[last + 1]; L.me = v.be[last + 1];  // band of the closing suffix row
        rowF = sF + mb + v.coff[last + 1]; rowX = sX + mb + v.coff[last + 1];
        kmer_at = last + 1 < n ? last + 1 : last;
        out = out_ll + (B.ref_off[b] + i) * A + base;
        const int endcol = L.hasA ? L.we : L.me;       // range of the row that is joined
        T = endcol - C0 + ri + 1;
      }
      if (kmer_at >= 0) {
        const int id = kmer_id(M, v, kmer_at, i, base);  // ModifiedSequence (sequence.cpp:30-38)
        L.mu = M.mean[id]; L.ac = M.ac[id]; L.mc = M.mc[id];
      }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) T = max(T, __shfl_xor_sync(NVB_FULL, T, o));

  LaneState<MEL> S;
  lane_reset(S);
  LaneOut res;
  res.f = 0.0; res.E = NVB_EZERO; res.p = 1.0; res.k = 0;
  const bool is_loader = L.role == NVB_ROLE_LOADER, is_join = L.role == NVB_ROLE_JOIN;
  for (int t = 0; t < T; t++) {
    const int c = C0 + (t - ri);
    const double x = __ldg(sig + min(max(c - 1, 0), N - 1));
    LaneOut in = shfl_up_out(res);
    if (is_loader || ri == 0) { in.f = 0.0; in.E = NVB_EZERO; }
    double rF = 0.0;
    int rX = NVB_EZERO;
    if ((is_loader || is_join) && c >= L.ms && c <= L.me) {
      rF = __ldg(rowF + (c - L.ms));
      rX = __ldg(rowX + (c - L.ms));
    }
    XD aout;
    lane_step<MEL>(L, S, c, x, in, rF, rX, res, aout);
    if (is_loader) { res.f = rF; res.E = rX; }
    if (L.role == NVB_ROLE_IDLE) { res.f = 0.0; res.E = NVB_EZERO; }
    if ((t & 31) == 31) lane_renorm(S);
  }
  if (is_join) {
    xd_renorm(S.acc);  // canonical mantissa: the result must not depend on how many extra steps the warp ran
    *out = log_ext(S.acc.f, S.acc.e);
  }
}

template <int MEL>
void launch_snp2(const ModelDev &M, const BatchDev &B, int wobbling, int b0, int b1, int64_t g0, int64_t n_tasks, int LT,
                 int TPW, const int64_t *mb, const double *pF, const int32_t *pX, const double *sF, const int32_t *sX,
                 double *out, cudaStream_t st) {
  const int warps = 4;
  const int64_t n_warps = (n_tasks + TPW - 1) / TPW;
  snp2_kernel<MEL><<<(unsigned)((n_warps + warps - 1) / warps), warps * NVB_WARP, 0, st>>>(
      M, B, wobbling, b0, b1, g0, n_tasks, LT, TPW, mb, pF, pX, sF, sX, out);
}

}  // namespace

int nvbk_snp2(const ModelDev &M, const BatchDev &B, int wobbling, int b0, int b1, int64_t g0, int64_t g1,
              const int64_t *d_mat_base, const double *pF, const int32_t *pX, const double *sF, const int32_t *sX,
              double *d_out_ll, cudaStream_t st) {
  const int LT = M.k + 2;  // LOADER + k pairs + JOIN
  if (LT > NVB_WARP) return -1;
  const int TPW = NVB_WARP / LT;
  const int64_t n_tasks = (g1 - g0) * (M.alphabet - 1);
  if (n_tasks <= 0) return 0;
#define NVB_SNP2(MEL) launch_snp2<MEL>(M, B, wobbling, b0, b1, g0, n_tasks, LT, TPW, d_mat_base, pF, pX, sF, sX, d_out_ll, st)
  switch (B.mel) {
    case 0: NVB_SNP2(0); break;
    case 1: NVB_SNP2(1); break;
    case 2: NVB_SNP2(2); break;
    case 3: NVB_SNP2(3); break;
    case 4: NVB_SNP2(4); break;
    case 5: NVB_SNP2(5); break;
    case 6: NVB_SNP2(6); break;
    default: return -1;
  }
#undef NVB_SNP2
  return 0;
}
