// band.cu -- band geometry on the device.
// Replaces ComputeBandStarts / ComputeBandEnds (reference nadavca/dtw/dtw.cpp:7-35): anchors (signal idx, ref idx)
// give row [max(0,s-bw), min(N,s+bw)] (inclusive), rows without anchor default to [0,N]; then a running max of
// the starts (ascending) and a running min of the ends (descending).  Also produces the exclusive scan of the
// band widths (packed-row offsets of the DP matrices), the widest row and a bad-band flag per read.
#include "common.cuh"
#include "kernels.h"

namespace {

constexpr int kThreads = 256;

__global__ void __launch_bounds__(kThreads) band_kernel(BatchDev B, int64_t *summary /* [n_reads][4] */) {
  const int b = blockIdx.x;
  const int tid = threadIdx.x;
  const int64_t r0 = B.ref_off[b];
  const int n = (int)(B.ref_off[b + 1] - r0);
  const int N = (int)(B.sig_off[b + 1] - B.sig_off[b]);
  const int rows = n + 1;
  int32_t *bs = B.bs + r0 + b;
  int32_t *be = B.be + r0 + b;
  int64_t *coff = B.cell_off + r0 + 2 * (int64_t)b;  // n+2 entries; first used as anchor tags
  const int32_t *anc = B.anchors + 2 * B.anc_off[b];
  const int n_anc = (int)(B.anc_off[b + 1] - B.anc_off[b]);
  const int bw = B.bandwidth;

  __shared__ long long s_part[kThreads];
  __shared__ int s_bad, s_maxw, s_overlap;
  if (tid == 0) { s_bad = 0; s_maxw = 0; s_overlap = -0x7fffffff; }

  for (int j = tid; j < rows; j += kThreads) { bs[j] = 0; be[j] = N; coff[j] = -1; }
  __syncthreads();
  // the reference applies anchors in order, so the LAST anchor naming a row wins (dtw.cpp:11-15)
  for (int a = tid; a < n_anc; a += kThreads) {
    int r = anc[2 * a + 1];
    if (r >= 0 && r <= n) atomicMax((long long *)&coff[r], (long long)a);
  }
  __syncthreads();
  for (int j = tid; j < rows; j += kThreads) {
    long long a = coff[j];
    if (a >= 0) {
      int s = anc[2 * a];
      bs[j] = max(0, s - bw);
      be[j] = min(N, s + bw);
    }
  }
  __syncthreads();

  const int seg = (rows + kThreads - 1) / kThreads;
  const int lo = min(rows, tid * seg), hi = min(rows, lo + seg);

  // running max of starts, ascending (dtw.cpp:16-18)
  int m = 0;
  for (int j = lo; j < hi; j++) m = max(m, bs[j]);
  s_part[tid] = m;
  __syncthreads();
  int carry = 0;
  for (int t = 0; t < tid; t++) carry = max(carry, (int)s_part[t]);
  for (int j = lo; j < hi; j++) { carry = max(carry, bs[j]); bs[j] = carry; }
  __syncthreads();

  // running min of ends, descending (dtw.cpp:31-33)
  m = N;
  for (int j = lo; j < hi; j++) m = min(m, be[j]);
  s_part[tid] = m;
  __syncthreads();
  carry = N;
  for (int t = kThreads - 1; t > tid; t--) carry = min(carry, (int)s_part[t]);
  for (int j = hi - 1; j >= lo; j--) { carry = min(carry, be[j]); be[j] = carry; }
  __syncthreads();

  // widths -> exclusive scan (packed row offsets), widest row, bad-band flag
  long long sum = 0;
  int maxw = 0, bad = 0;
  for (int j = lo; j < hi; j++) {
    int w = be[j] - bs[j] + 1;
    if (w <= 0) { bad = 1; w = 0; }
    sum += w;
    maxw = max(maxw, w);
  }
  // rotating sweep (rows5.cu): a lane's slot must be free again two generations (64 pairs) later, i.e. band row j must
  // end at most 48 columns beyond the start of band row j + 63
  int overlap = -0x7fffffff;
  for (int j = lo; j < hi; j++)
    if (j + 63 < rows) overlap = max(overlap, be[j] - bs[j + 63]);
  atomicMax(&s_overlap, overlap);
  s_part[tid] = sum;
  if (bad) atomicOr(&s_bad, 1);
  atomicMax(&s_maxw, maxw);
  __syncthreads();
  long long base = 0;
  for (int t = 0; t < tid; t++) base += s_part[t];
  for (int j = lo; j < hi; j++) {
    coff[j] = base;
    int w = be[j] - bs[j] + 1;
    base += w > 0 ? w : 0;
  }
  __syncthreads();
  if (tid == 0) {
    long long total = 0;
    for (int t = 0; t < kThreads; t++) total += s_part[t];
    coff[rows] = total;
    int bad_all = s_bad || n <= 0 || N <= 0;
    B.flags[b] = bad_all ? 2 : 0;
    B.max_width[b] = s_maxw;
    summary[4 * b + 0] = total;
    summary[4 * b + 1] = rows > 0 ? max(0, be[0] - bs[0] + 1) : 0;
    summary[4 * b + 2] = rows > 0 ? max(0, be[n] - bs[n] + 1) : 0;
    const int no_rotation = s_overlap > 48;
    summary[4 * b + 3] = ((long long)(bad_all ? 1 : 0) << 32) | ((long long)no_rotation << 33) | (unsigned)s_maxw;
  }
}

// KmerModel::GetExpectedSignal (kmer_model.cpp:32-42), one thread per reference position of the batch.
__global__ void expected_signal_kernel(ModelDev M, BatchDev B, int64_t total, double *out) {
  int64_t g = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (g >= total) return;
  int lo = 0, hi = B.n_reads;  // last read with ref_off <= g
  while (hi - lo > 1) {
    int mid = (lo + hi) >> 1;
    if (B.ref_off[mid] <= g) lo = mid; else hi = mid;
  }
  ReadView v = read_view(B, lo);
  int i = (int)(g - B.ref_off[lo]);
  out[g] = M.mean[kmer_id(M, v, i, INT32_MIN, 0)];
}

// Emission parameters of every reference position of the batch (its unmodified k-mer), one thread per position:
// [mean, ac * S, mc * S, same-mean flags] with S = NVB_EXP_SCALE (dp3.cuh).  The sweeps read one 32-byte row when a lane takes a new
// row pair instead of walking reference bases -> k-mer id -> three model tables (three dependent global loads in
// front of a stalled warp).
__global__ void row_emission_kernel(ModelDev M, BatchDev B, int64_t total, double scale, double *out) {
  int64_t g = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (g >= total) return;
  int lo = 0, hi = B.n_reads;  // last read with ref_off <= g
  while (hi - lo > 1) {
    int mid = (lo + hi) >> 1;
    if (B.ref_off[mid] <= g) lo = mid; else hi = mid;
  }
  ReadView v = read_view(B, lo);
  const int i = (int)(g - B.ref_off[lo]);
  const int id = kmer_id(M, v, i, INT32_MIN, 0);
  const double mean = M.mean[id];
  // bit 0 / 1: the position before / after has the same mean (GetTransitionDistribution, kmer_model.cpp:64-94: the
  // transition between two such rows has probability 0)
  int same = 0;
  if (i >= 1 && M.mean[kmer_id(M, v, i - 1, INT32_MIN, 0)] == mean) same |= 1;
  if (i + 1 < v.n && M.mean[kmer_id(M, v, i + 1, INT32_MIN, 0)] == mean) same |= 2;
  double4 row;
  row.x = mean; row.y = M.ac[id] * scale; row.z = M.mc[id] * scale; row.w = (double)same;
  reinterpret_cast<double4 *>(out)[g] = row;
}

}  // namespace

void nvbk_row_emission(const ModelDev &M, const BatchDev &B, int64_t total, double scale, double *d_out, cudaStream_t st) {
  if (total > 0) row_emission_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(M, B, total, scale, d_out);
}

void nvbk_band(const BatchDev &B, int64_t *d_summary, cudaStream_t st) {
  if (B.n_reads > 0) band_kernel<<<B.n_reads, kThreads, 0, st>>>(B, d_summary);
}

void nvbk_expected_signal(const ModelDev &M, const BatchDev &B, int64_t total, double *d_out, cudaStream_t st) {
  if (total > 0) expected_signal_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(M, B, total, d_out);
}
