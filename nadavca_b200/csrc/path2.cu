// path2.cu -- posterior rows, max-product path search and traceback of RefineAlignment, second generation.
//
// Replaces reference nadavca/dtw/dtw.cpp:199-227 with Node::operator* (node.cpp:23-29) and
// PathSearchingNode::NextRow / GetBestIndex / GetPrevious (node.cpp:39-91).
//
// Stage 1, score_kernel (fully parallel, HBM-bound): score = log(prefix * suffix) for every stored cell, formed
//   from the (mantissa, exponent) planes written by the sweeps (rows4.cu, rows5.cu) and written over the prefix mantissa plane.
// Stage 2, path2_kernel (one warp per read, rows in order): dp[r][c] = score[r][c] + max_{i' <= c - m} dp[r-1][i'].
//   The running "best predecessor" of the reference (node.cpp:68-89) is a prefix maximum of the previous dp row; it
//   is kept in shared memory as M[i] = max(dp[r-1][..i]) so a row is: K consecutive columns per lane, K lookups,
//   one lane-local scan, one 5-step warp max-scan.  The reference's back-pointer prev[c] is the FIRST index
//   attaining that maximum (strict '>' while scanning upwards, node.cpp:72-75,82-85), i.e. the last column i' <= c-m
//   at which dp[r-1] set a new strict record.  So one bit per cell ("this cell is a new record of its row") replaces
//   the 4-byte back-pointer, and the traceback (dtw.cpp:215-227) is a find-last-set over the record bits of the
//   row above, done by the whole warp with one ballot per row.
#include <stdio.h>
#include "dp3.cuh"
#include "kernels.h"

namespace {

__global__ void __launch_bounds__(256) score_kernel(int64_t total, double *pF, const int32_t *pX, const double *sF,
                                                    const int32_t *sX) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t x = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; x < total; x += stride)
    pF[x] = log_ext(pF[x] * __ldg(sF + x), __ldg(pX + x) + __ldg(sX + x));
}

__device__ __forceinline__ void row_geom2(const ReadView &v, int mode, int r, int &s, int &e, int64_t &off) {
  int j;
  if (mode == NVB_MODE_TRANS) { j = (r + 1) >> 1; off = trans_row_off(v, r); }
  else { j = r; off = v.coff[r]; }
  s = v.bs[j];
  e = v.be[j];
}

__device__ __forceinline__ double warp_excl_max(double x, int lane) {
#pragma unroll
  for (int o = 1; o < NVB_WARP; o <<= 1) {
    const double y = __shfl_up_sync(NVB_FULL, x, o);
    if (lane >= o) x = fmax(x, y);
  }
  const double y = __shfl_up_sync(NVB_FULL, x, 1);
  return lane ? y : nvb_neg_inf();
}

// Last record bit at relative index <= t of a row whose record words start at `fl` (chunks of 32 words, PK columns
// per word).  Returns -1 when there is none.  Called by one whole warp.
template <int PK>
__device__ __forceinline__ int find_last_record(const uint32_t *fl, bool in_smem, int t, int lane) {
  constexpr int PCH = PK * NVB_WARP;
  if (t < 0) return -1;
  int ch = t / PCH;
  int tl = t - ch * PCH;
  for (;;) {
    // staged rows are read from shared memory; unstaged ones (rows too wide to stage) straight from L2
    uint32_t w = in_smem ? fl[ch * NVB_WARP + lane] : __ldcg(fl + ch * NVB_WARP + lane);
    const int lane_t = tl / PK, bit_t = tl - lane_t * PK;
    if (lane > lane_t) w = 0;
    else if (lane == lane_t) w &= (2u << bit_t) - 1u;
    const unsigned any = __ballot_sync(NVB_FULL, w != 0);
    if (any) {
      const int hl = 31 - __clz(any);
      const uint32_t wv = __shfl_sync(NVB_FULL, w, hl);
      return ch * PCH + hl * PK + (31 - __clz(wv));
    }
    if (--ch < 0) return -1;
    tl = PCH - 1;
  }
}

// One CTA per read, NWP = blockDim.x / 32 warps.  A row is cut into chunks of 32 * PK consecutive columns (PK per
// lane); chunk ch of a row is worked by warp ch % NWP, all chunks of a round of NWP at the same time: each warp scans
// its chunk, publishes the chunk maximum, and after a CTA barrier takes the maximum of the chunks in front of it as
// its carry.  Rows are inherently sequential (dp[r] needs the prefix maxima of dp[r-1]); the kernel is bound by the
// dependent-issue latency of one row, so everything that is not the recurrence is kept off it:
//   * band geometry (start, end, offset of every row) is computed GB rows at a time into a shared-memory table
//     (ncu profiles/r02f: 472 warp instructions per row and warp, most of them geometry arithmetic and loads);
//   * the scores of row r+D are copied global -> shared with cp.async while row r is worked (no registers, no
//     stall on first use: 21 % of the stall samples with a one-row register prefetch);
//   * the traceback stages the record words AND the geometry of 64 rows at a time (all warps), then warp 0 walks them.
constexpr int GB = 64;  // rows per geometry block

struct RowGeo {
  int s, e;
  long long off;
};

__device__ __forceinline__ void cp_async8(void *dst_smem, const void *src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((unsigned)__cvta_generic_to_shared(dst_smem)), "l"(src)
               : "memory");
}

// 16 bytes global -> shared past L1 (the record words were written by this CTA a moment ago)
__device__ __forceinline__ void cp_async16_cg(void *dst_smem, const void *src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(dst_smem)), "l"(src)
               : "memory");
}

template <int PK, int D>
__global__ void __launch_bounds__(256) path2_kernel(BatchDev B, int mode, int b0, int n_items, const int64_t *mat_base,
                                                    const double *score, uint32_t *records, const int64_t *rec_base,
                                                    double *gscratch, const int64_t *dp_base, int smem_width,
                                                    int stage_words, int32_t *events, int32_t *status) {
  extern __shared__ double smem[];
  constexpr int PCH = PK * NVB_WARP;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, NWP = blockDim.x >> 5, NT = blockDim.x;
  const int item = blockIdx.x;
  if (item >= n_items) return;
  const int b = b0 + item;
  ReadView v = read_view(B, b);
  const int n = v.n;
  int32_t *ev = events + 2 * B.ref_off[b];
  if (B.flags[b]) {
    for (int i = threadIdx.x; i < 2 * n; i += blockDim.x) ev[i] = -1;
    if (threadIdx.x == 0) status[b] = B.flags[b];
    return;
  }
  const double NINF = nvb_neg_inf();
  const int R = (mode == NVB_MODE_TRANS) ? 2 * n : n + 1;
  const int maxw = B.max_width[b];
  const int nch = (maxw + PCH - 1) / PCH;
  const double *SC = score + mat_base[b];
  uint32_t *FL = records + rec_base[b];
  // shared memory: chunk maxima [2][8] | geometry table [GB + D + 1] | score ring [D][NT * PK] | two prefix-max rows
  double *tot = smem;
  RowGeo *geo = reinterpret_cast<RowGeo *>(smem + 16);
  double *ring = smem + 16 + 2 * (GB + D + 1);
  double *body = ring + (size_t)D * NT * PK;  // also the staging area of the traceback
  double *Mprev, *Mcur;
  if (smem_width > 0) { Mprev = body; Mcur = Mprev + smem_width; }
  else { Mprev = gscratch + dp_base[b]; Mcur = Mprev + maxw; }

  int gbase = 0;
  auto fill_geo = [&](int base, int count) {  // rows base .. base+count-1 into geo[0 .. count)
    for (int i = threadIdx.x; i < count; i += NT) {
      RowGeo g;
      g.s = 0; g.e = -1; g.off = 0;
      if (base + i >= 0 && base + i < R) {
        int64_t o;
        row_geom2(v, mode, base + i, g.s, g.e, o);
        g.off = o;
      }
      geo[i] = g;
    }
  };
  // this thread's PK scores of row r (first round of chunks) -> ring slot r % D
  auto request = [&](int r) {
    double *dst = ring + (size_t)(r % D) * NT * PK + threadIdx.x * PK;
    const RowGeo g = geo[r - gbase];
    const int c0 = g.s + warp * PCH + lane * PK;
#pragma unroll
    for (int j = 0; j < PK; j++) {
      // columns outside the band are never copied: their slots hold stale values, which the consumer masks (a plain
      // store here would have to wait for every cp.async in flight: 25 % of the stall samples, ncu profiles/r02m)
      if (r < R && c0 + j <= g.e) cp_async8(dst + j, SC + g.off + (c0 + j - g.s));
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  fill_geo(0, GB + D + 1);
  __syncthreads();
#pragma unroll
  for (int d = 0; d < D; d++) request(d);

  int ps = 0, pe = -1;
  int flip = 0;
  for (int r = 0; r < R; r++) {
    if (r - gbase == GB) {  // next geometry block (the end-of-row barrier of row r-1 separates it from the readers)
      gbase = r;
      fill_geo(gbase, GB + D + 1);
      __syncthreads();
    }
    const int m = (r == 0) ? 0 : ((mode == NVB_MODE_TRANS && ((r - 1) & 1)) ? 0 : B.mel);  // dtw.cpp:165-179
    const RowGeo g = geo[r - gbase];
    const int s = g.s, e = g.e;
    asm volatile("cp.async.wait_group %0;" ::"n"(D - 1) : "memory");  // this thread's scores of row r have landed
    double cur[PK];
    {
      const double *src = ring + (size_t)(r % D) * NT * PK + threadIdx.x * PK;
#pragma unroll
      for (int j = 0; j < PK; j++) cur[j] = src[j];
    }
    request(r + D);  // into the slot just read (own elements only)
    const int w = e - s + 1;
    double round_carry = NINF;  // maximum over all chunks of the rounds before this one
    for (int base = 0; base * PCH < w; base += NWP) {
      const int ch = base + warp;
      const int c0 = s + ch * PCH + lane * PK;
      double dp[PK];
#pragma unroll
      for (int j = 0; j < PK; j++) {
        const int c = c0 + j;
        double sc = (c <= e) ? cur[j] : NINF;  // (slots of columns outside the band hold stale values)
        if (base > 0) sc = (c <= e) ? __ldg(SC + g.off + (c - s)) : NINF;  // rows wider than NWP chunks: later rounds
        if (r > 0) {
          // best predecessor over i' <= c - m inside the previous row's band (node.cpp:68-89)
          const int q = c - m;
          const double bv = (c <= e && q >= ps) ? Mprev[min(q, pe) - ps] : NINF;
          sc = bv + sc;
        }
        dp[j] = (c <= e) ? sc : NINF;
      }
      // lane-local inclusive prefix maxima, the warp-wide exclusive carry, then the carry over the chunks in front
      double lm = dp[0];
#pragma unroll
      for (int j = 1; j < PK; j++) lm = fmax(lm, dp[j]);
      const double lane_excl = warp_excl_max(lm, lane);
      double *T = tot + 8 * flip;
      if (NWP > 1) {
        const double chunk_max = __shfl_sync(NVB_FULL, fmax(lane_excl, lm), NVB_WARP - 1);
        if (lane == 0) T[warp] = chunk_max;
        __syncthreads();
      }
      double carry = round_carry;
      if (NWP > 1) {
        for (int k = 0; k < warp; k++) carry = fmax(carry, T[k]);
      }
      uint32_t bits = 0;
      double before = fmax(carry, lane_excl);
#pragma unroll
      for (int j = 0; j < PK; j++) {
        if (dp[j] > before) bits |= 1u << j;  // strict '>': the lowest index wins ties (node.cpp:72-75)
        before = fmax(before, dp[j]);
        if (c0 + j <= e) Mcur[c0 + j - s] = before;
      }
      if (ch < nch) FL[((int64_t)r * nch + ch) * NVB_WARP + lane] = bits;
      if (NWP > 1) {
        if (w > NWP * PCH) {  // more rounds follow
          for (int k = 0; k < NWP; k++) round_carry = fmax(round_carry, T[k]);
        }
        flip ^= 1;
      } else {
        round_carry = __shfl_sync(NVB_FULL, before, NVB_WARP - 1);
      }
    }
    __syncthreads();  // Mcur is complete: the next row reads it as Mprev
    double *t = Mprev; Mprev = Mcur; Mcur = t;
    ps = s; pe = e;
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  __threadfence_block();
  __syncthreads();

  // Traceback (dtw.cpp:211-227).  The chain "record of row r -> column in row r-1" is sequential, but what it reads is
  // not: all warps copy the record words and the geometry of a group of rows into shared memory at once (one global
  // round trip per group instead of several per row), then warp 0 walks the group from there.  Step r (R down to 1)
  // turns the column of row r into the column of row r-1 with the records of row r-1; the virtual step R picks the end
  // point: GetBestIndex on the last row (node.cpp:48-58), the first index of the row maximum = its last record.
  uint32_t *stage = reinterpret_cast<uint32_t *>(body);
  const int words_per_row = nch * NVB_WARP;
  const bool staged = stage_words >= words_per_row;
  const int group = staged ? min(GB, stage_words / words_per_row) : GB;
  int rel = 0;
  bool no_path = false;
  for (int hi = R - 1; hi >= 0; hi -= group) {
    const int lo = max(0, hi - group + 1);
    __syncthreads();
    fill_geo(lo, hi - lo + 2);  // geo[i] = row lo + i, up to row hi + 1 (empty when hi + 1 == R)
    if (staged) {
      const uint32_t *src = FL + (int64_t)lo * words_per_row;
      // asynchronous 16-byte copies, all in flight at once: a load -> store loop left the CTA waiting on one L2
      // round trip per word and thread (24 % of the kernel's stall samples, ncu r02x).  Rows are whole chunks of 32
      // words and the record planes start on 128-byte boundaries, so every piece is aligned.
      const int words = (hi - lo + 1) * words_per_row;
      for (int i = 4 * threadIdx.x; i < words; i += 4 * blockDim.x) cp_async16_cg(stage + i, src + i);
      asm volatile("cp.async.commit_group;" ::: "memory");
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    __syncthreads();
    if (warp != 0 || no_path) continue;
    const uint32_t *rows0 = staged ? stage : FL + (int64_t)lo * words_per_row;
    for (int r = hi + 1; r > lo; r--) {
      int q;
      const RowGeo below = geo[r - 1 - lo];  // row r-1
      if (r == R) {
        q = below.e - below.s;
      } else {
        const int col = geo[r - lo].s + rel;
        if (lane == 0) {
          if (mode == NVB_MODE_TRANS) {
            ev[r] = col;  // events[r/2][r%2]
          } else {
            ev[2 * (r - 1) + 1] = col;
            if (r + 1 < R) ev[2 * r] = col;
          }
        }
        const int mm = (mode == NVB_MODE_TRANS && ((r - 1) & 1)) ? 0 : B.mel;
        q = min(col - mm, below.e) - below.s;
      }
      rel = find_last_record<PK>(rows0 + (int64_t)(r - 1 - lo) * words_per_row, staged, q, lane);
      if (r == R && rel < 0) { no_path = true; break; }  // no valid path in the band (dtw.cpp:211-213)
    }
  }
  if (warp != 0) return;
  if (no_path) {
    for (int i = lane; i < 2 * n; i += NVB_WARP) ev[i] = -1;
    if (lane == 0) status[b] = 1;
    return;
  }
  if (lane == 0) {  // row 0
    int rs, re;
    int64_t o;
    row_geom2(v, mode, 0, rs, re, o);
    const int col = rs + rel;
    if (mode == NVB_MODE_TRANS) ev[0] = col;
    else if (R > 1) ev[0] = col;
    status[b] = 0;
  }
}

template <int PK, int D>
int launch_path(const BatchDev &B, int mode, int b0, int n_items, const int64_t *d_mat_base, const double *score,
                uint32_t *d_records, const int64_t *d_rec_base, double *d_dp, const int64_t *d_dp_base, int wave_maxw,
                int32_t *d_events, int32_t *d_status, cudaStream_t st) {
  // warps per read: one per chunk of the widest row, at most 8 (wider rows take several rounds)
  int warps = (wave_maxw + PK * NVB_WARP - 1) / (PK * NVB_WARP);
  warps = warps < 1 ? 1 : (warps > 8 ? 8 : warps);
  const int nch = (wave_maxw + PK * NVB_WARP - 1) / (PK * NVB_WARP);
  // shared memory: chunk maxima, geometry table, score ring, then two prefix-maximum rows (global scratch rows for
  // very wide bands), reused by the traceback as a staging area for the record words of up to GB rows
  const size_t head = (16 + 2 * (GB + D + 1) + (size_t)D * warps * NVB_WARP * PK) * sizeof(double);
  const size_t want_stage = (size_t)32 * nch * NVB_WARP * sizeof(uint32_t);
  int smem_width = wave_maxw;
  size_t rows_bytes = (size_t)2 * wave_maxw * sizeof(double);
  if (head + rows_bytes > 200 * 1024) { smem_width = 0; rows_bytes = 0; }
  size_t body = rows_bytes > want_stage ? rows_bytes : want_stage;
  if (head + body > 200 * 1024) body = rows_bytes;
  const size_t smem = head + body;
  const int stage_words = (int)(body / sizeof(uint32_t));
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(path2_kernel<PK, D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return -1;
  }
  path2_kernel<PK, D><<<n_items, warps * NVB_WARP, smem, st>>>(B, mode, b0, n_items, d_mat_base, score, d_records,
                                                                d_rec_base, d_dp, d_dp_base, smem_width, stage_words,
                                                                d_events, d_status);
  return 0;
}

}  // namespace

// Columns per lane of the path search for a batch whose widest band row has `maxw` columns: 4 warps cover rows of
// up to 384 columns at 3 per lane, 8 warps 1536 columns at 6; wider rows use 11 per lane (8 warps: 2816 columns per
// round).
int nvbk_path2_columns_per_lane(int maxw) { return maxw <= 384 ? 3 : (maxw <= 1536 ? 6 : 11); }

void nvbk_score(int64_t total_cells, double *pF, const int32_t *pX, const double *sF, const int32_t *sX,
                cudaStream_t st) {
  if (total_cells <= 0) return;
  const int64_t want = (total_cells + 255) / 256;
  const unsigned blocks = (unsigned)(want < 148 * 32 ? want : 148 * 32);
  score_kernel<<<blocks, 256, 0, st>>>(total_cells, pF, pX, sF, sX);
}

// pk: columns per lane the record words of this batch were laid out for (nvbk_path2_columns_per_lane of the batch's
// widest row).  Returns -1 when the shared-memory reservation fails.
int nvbk_path2(const BatchDev &B, int mode, int b0, int b1, int pk, const int64_t *d_mat_base, const double *score,
               uint32_t *d_records, const int64_t *d_rec_base, double *d_dp, const int64_t *d_dp_base, int wave_maxw,
               int32_t *d_events, int32_t *d_status, cudaStream_t st) {
  const int n_items = b1 - b0;
  if (n_items <= 0) return 0;
  switch (pk) {
    case 3: return launch_path<3, 4>(B, mode, b0, n_items, d_mat_base, score, d_records, d_rec_base, d_dp, d_dp_base, wave_maxw, d_events, d_status, st);
    case 6: return launch_path<6, 3>(B, mode, b0, n_items, d_mat_base, score, d_records, d_rec_base, d_dp, d_dp_base, wave_maxw, d_events, d_status, st);
    default: return launch_path<11, 2>(B, mode, b0, n_items, d_mat_base, score, d_records, d_rec_base, d_dp, d_dp_base, wave_maxw, d_events, d_status, st);
  }
}
