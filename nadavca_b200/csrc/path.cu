// path.cu -- posterior rows, max-product path search and traceback of RefineAlignment.
//
// Replaces reference nadavca/dtw/dtw.cpp:199-227 with Node::operator* (node.cpp:23-29),
// PathSearchingNode::NextRow / GetBestIndex / GetPrevious (node.cpp:39-91).
//
// One warp per read walks the rows in order.  A row's score is log(prefix * suffix), formed from the (mantissa,
// exponent) planes written by rows2.cu (one log per cell, read coalesced from HBM); the running "best predecessor
// over i' <= c - m" is an inclusive (max, first-argmax) scan along the row -- strict '>' so the lowest index wins
// ties, index -1 while everything is log(0).  The back-pointer of a cell overwrites that cell's slot in the prefix
// exponent plane (dead once the score is formed), so the traceback needs no extra HBM.  Lane 0 then walks the
// back-pointers from the last row.
#include "dp2.cuh"
#include "kernels.h"

namespace {

struct Best {
  double v;
  int i;
};

__device__ __forceinline__ Best best_shfl_up(Best x, int o) {
  Best r;
  r.v = __shfl_up_sync(NVB_FULL, x.v, o);
  r.i = __shfl_up_sync(NVB_FULL, x.i, o);
  return r;
}

// inclusive scan, op(left, right) = right strictly greater ? right : left
__device__ __forceinline__ Best best_scan(Best x, int lane) {
#pragma unroll
  for (int o = 1; o < NVB_WARP; o <<= 1) {
    Best l = best_shfl_up(x, o);
    if (lane >= o && !(x.v > l.v)) x = l;
  }
  return x;
}

__device__ __forceinline__ void row_geom(const ReadView &v, int mode, int r, int &s, int &e, int64_t &off) {
  int j;
  if (mode == NVB_MODE_TRANS) { j = (r + 1) >> 1; off = trans_row_off(v, r); }
  else { j = r; off = v.coff[r]; }
  s = v.bs[j];
  e = v.be[j];
}

__global__ void __launch_bounds__(128) path_kernel(BatchDev B, int mode, int b0, int n_items, const int64_t *mat_base,
                                                   const double *pF, int32_t *pX, const double *sF,
                                                   const int32_t *sX, double *dp, const int64_t *dp_base,
                                                   int32_t *events, int32_t *status) {
  const int wic = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int item = blockIdx.x * (blockDim.x >> 5) + wic;
  if (item >= n_items) return;
  const int b = b0 + item;
  ReadView v = read_view(B, b);
  const int n = v.n;
  int32_t *ev = events + 2 * B.ref_off[b];
  if (B.flags[b]) {
    for (int i = lane; i < 2 * n; i += NVB_WARP) ev[i] = -1;
    if (lane == 0) status[b] = B.flags[b];
    return;
  }
  const double NINF = nvb_neg_inf();
  const int R = (mode == NVB_MODE_TRANS) ? 2 * n : n + 1;
  const double *PF = pF + mat_base[b];
  const double *SF = sF + mat_base[b];
  const int32_t *SX = sX + mat_base[b];
  int32_t *bp = pX + mat_base[b];  // prefix exponents, overwritten cell by cell with back-pointers
  double *dp_prev = dp + dp_base[b];
  double *dp_cur = dp_prev + B.max_width[b];

  int s, e, ps, pe;
  int64_t off;
  row_geom(v, mode, 0, s, e, off);
  for (int c = s + lane; c <= e; c += NVB_WARP) {  // dp[0] = PathSearchingNode(all_paths_sum[0]) (dtw.cpp:205)
    int64_t x = off + c - s;
    double sc = log_ext(PF[x] * SF[x], bp[x] + SX[x]);
    __stcg(dp_prev + (c - s), sc);
    bp[x] = -1;
  }
  __syncwarp();

  for (int r = 1; r < R; r++) {
    ps = s; pe = e;
    row_geom(v, mode, r, s, e, off);
    const int m = (mode == NVB_MODE_TRANS && ((r - 1) & 1)) ? 0 : B.mel;  // dtw.cpp:165-179
    Best carry;
    carry.v = NINF; carry.i = -1;
    // predecessors strictly before the row's first admissible index (node.cpp:68-76)
    const int pre_hi = min(pe, s - m - 1);
    for (int base = ps; base <= pre_hi; base += NVB_WARP) {
      const int i = base + lane;
      Best x;
      x.v = (i <= pre_hi) ? __ldcg(dp_prev + (i - ps)) : NINF;
      x.i = (x.v > NINF) ? i : -1;
      x = best_scan(x, lane);
      if (!(x.v > carry.v)) x = carry;
      carry.v = __shfl_sync(NVB_FULL, x.v, NVB_WARP - 1);
      carry.i = __shfl_sync(NVB_FULL, x.i, NVB_WARP - 1);
    }
    // node.cpp:78-89
    for (int cb = s; cb <= e; cb += NVB_WARP) {
      const int c = cb + lane;
      const int from = c - m;
      const bool ok = c <= e && from >= ps && from <= pe;
      Best x;
      x.v = ok ? __ldcg(dp_prev + (from - ps)) : NINF;
      x.i = (x.v > NINF) ? from : -1;
      x = best_scan(x, lane);
      if (!(x.v > carry.v)) x = carry;
      if (c <= e) {
        const int64_t k = off + c - s;
        const double sc = log_ext(PF[k] * SF[k], bp[k] + SX[k]);
        __stcg(dp_cur + (c - s), x.v + sc);
        bp[k] = x.i;
      }
      carry.v = __shfl_sync(NVB_FULL, x.v, NVB_WARP - 1);
      carry.i = __shfl_sync(NVB_FULL, x.i, NVB_WARP - 1);
    }
    double *t = dp_prev; dp_prev = dp_cur; dp_cur = t;
    __syncwarp();
  }

  // GetBestIndex on the last row (node.cpp:48-58)
  Best best;
  best.v = NINF; best.i = -1;
  for (int cb = s; cb <= e; cb += NVB_WARP) {
    const int c = cb + lane;
    Best x;
    x.v = (c <= e) ? __ldcg(dp_prev + (c - s)) : NINF;
    x.i = (x.v > NINF) ? c : -1;
    x = best_scan(x, lane);
    if (!(x.v > best.v)) x = best;
    best.v = __shfl_sync(NVB_FULL, x.v, NVB_WARP - 1);
    best.i = __shfl_sync(NVB_FULL, x.i, NVB_WARP - 1);
  }
  if (best.i < 0) {  // no valid path in the band (dtw.cpp:211-213)
    for (int i = lane; i < 2 * n; i += NVB_WARP) ev[i] = -1;
    if (lane == 0) status[b] = 1;
    return;
  }
  __threadfence_block();
  if (lane == 0) {  // dtw.cpp:215-227
    int bi = best.i;
    for (int r = R - 1; r >= 0; r--) {
      row_geom(v, mode, r, s, e, off);
      if (mode == NVB_MODE_TRANS) {
        ev[r] = bi;  // events[r/2][r%2]
      } else {
        if (r > 0) ev[2 * (r - 1) + 1] = bi;
        if (r + 1 < R) ev[2 * r] = bi;
      }
      bi = __ldcg(bp + (off + bi - s));
    }
    status[b] = 0;
  }
}

}  // namespace

void nvbk_path(const BatchDev &B, int mode, int b0, int b1, const int64_t *d_mat_base, const double *pF,
               int32_t *pX, const double *sF, const int32_t *sX, double *d_dp, const int64_t *d_dp_base,
               int32_t *d_events, int32_t *d_status, cudaStream_t st) {
  const int n_items = b1 - b0;
  if (n_items <= 0) return;
  path_kernel<<<(n_items + 3) / 4, 128, 0, st>>>(B, mode, b0, n_items, d_mat_base, pF, pX, sF, sX, d_dp, d_dp_base,
                                                 d_events, d_status);
}
