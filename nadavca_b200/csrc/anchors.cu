// anchors.cu -- batched anchor construction: the step in front of the DP (reference nadavca/alignment.py).
//
// For every read: walk the CIGAR of its BWA hit, keep the MATCHING bases as (read index, reference index) anchors and
// flip them to read orientation for the reverse strand (alignment.py:109-140); drop anchors whose base the basecaller
// did not place in the signal, convert to (sample index, reference index) relative to the extended signal range and
// the first anchored reference position, and derive the ranges (alignment.py:142-186).  Pure integer work, one warp
// per read: the CIGAR operations are walked in order (uniformly by the warp), the positions of an M block are
// compared 32 at a time and compacted with ballots.
#include "common.cuh"
#include "kernels.h"

namespace {

enum { OP_M = 0, OP_I = 1, OP_D = 2, OP_S = 3 };

struct Walk {
  int total, kept;          // matching bases; those present in the base -> sample map
  int first_read, last_read;  // read index (read orientation) of the first / last matching base in OUTPUT order
};

// One pass over the CIGAR.  WRITE = false counts, WRITE = true stores the anchors at their final position (ascending
// in read orientation: for the reverse strand the walk order is reversed).
template <bool WRITE>
__device__ Walk walk_cigar(const AnchorBatch &A, int b, int lane, int total_kept, int32_t *out) {
  const int64_t o0 = A.cigar_off[b], o1 = A.cigar_off[b + 1];
  const int64_t r0 = A.read_off[b];
  const int L = (int)(A.read_off[b + 1] - r0);
  const int8_t *seq = A.read_seq + r0;
  const int32_t *map = A.mapping + r0;
  const bool rev = A.reverse[b] != 0;
  const int64_t G = A.genome_len;
  int index_in_read = 0;
  int64_t index_in_ref = A.mapped_pos[b];
  Walk w;
  w.total = 0; w.kept = 0; w.first_read = -1; w.last_read = -1;
  for (int64_t o = o0; o < o1; o++) {
    const int op = A.cigar_op[o], num = A.cigar_len[o];
    if (op == OP_S || op == OP_I) { index_in_read += num; continue; }
    if (op == OP_D) { index_in_ref += num; continue; }
    for (int base = 0; base < num; base += NVB_WARP) {
      const int i = base + lane;
      bool match = false, placed = false;
      int read_idx = 0, sample = 0;
      int64_t ref_idx = 0;
      if (i < num) {
        const int j = index_in_read + i;          // index in the ORIENTED read
        const int64_t g = index_in_ref + i;
        if (j < L && g >= 0 && g < G) {
          const int orig = rev ? L - 1 - j : j;   // index in read.sequence
          const int base_code = rev ? 3 - seq[orig] : seq[orig];
          match = A.genome[g] == base_code;
          read_idx = orig;
          ref_idx = rev ? G - 1 - g : g;          // counted from the END of the contig on the reverse strand
          sample = map[orig];
          placed = match && sample >= 0;
        }
      }
      const unsigned mm = __ballot_sync(NVB_FULL, match), mp = __ballot_sync(NVB_FULL, placed);
      if (mm) {
        // first / last matching base in walk order (read_sequence_range uses the UNFILTERED anchors)
        const int lo_lane = __ffs(mm) - 1, hi_lane = 31 - __clz(mm);
        const int lo_read = __shfl_sync(NVB_FULL, read_idx, lo_lane), hi_read = __shfl_sync(NVB_FULL, read_idx, hi_lane);
        if (w.total == 0) w.first_read = lo_read;
        w.last_read = hi_read;
      }
      if (WRITE && placed) {
        const int k = w.kept + __popc(mp & ((1u << lane) - 1u));  // position in walk order
        const int dst = rev ? total_kept - 1 - k : k;
        out[2 * dst] = sample;                 // both made relative by the caller
        out[2 * dst + 1] = (int32_t)ref_idx;
      }
      w.total += __popc(mm);
      w.kept += __popc(mp);
    }
    index_in_read += num;
    index_in_ref += num;
  }
  return w;
}

// meta per read: [n_anchors, ref_start, ref_end, sig_ext_start, sig_ext_end, read_first, read_end]  (-1s when empty)
__global__ void __launch_bounds__(128) anchors_kernel(AnchorBatch A, int32_t *anchors, int64_t *meta) {
  const int wic = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.x * (blockDim.x >> 5) + wic;
  if (b >= A.n_reads) return;
  const int64_t r0 = A.read_off[b];
  int32_t *out = anchors + 2 * r0;  // capacity: one (sample, reference) pair per base of the read
  const Walk c = walk_cigar<false>(A, b, lane, 0, nullptr);
  int64_t *m = meta + 7 * (int64_t)b;
  if (c.kept == 0) {
    if (lane < 7) m[lane] = lane == 0 ? 0 : -1;
    return;
  }
  walk_cigar<true>(A, b, lane, c.kept, out);
  __syncwarp();
  const bool rev = A.reverse[b] != 0;
  const int n = c.kept;
  const int64_t sig_first = out[0], ref_first = out[1], sig_last = out[2 * (n - 1)], ref_last = out[2 * (n - 1) + 1];
  const int64_t ext_start = max((int64_t)0, sig_first - A.bandwidth);
  const int64_t ext_end = min((int64_t)A.n_signal[b], sig_last + 1 + A.bandwidth);
  __syncwarp();
  // (sample - extended start, reference index - first anchored index)   alignment.py:153-168
  for (int i = lane; i < n; i += NVB_WARP) {
    out[2 * i] -= (int32_t)ext_start;
    out[2 * i + 1] -= (int32_t)ref_first;
  }
  if (lane == 0) {
    int64_t ref_start = ref_first, ref_end = ref_last + 1;
    if (rev) { const int64_t s = A.genome_len - ref_end, e = A.genome_len - ref_start; ref_start = s; ref_end = e; }
    // the walk order of the reverse strand is descending in read orientation: first / last swap (alignment.py:137-138)
    const int read_first = rev ? c.last_read : c.first_read, read_last = rev ? c.first_read : c.last_read;
    m[0] = n; m[1] = ref_start; m[2] = ref_end; m[3] = ext_start; m[4] = ext_end; m[5] = read_first; m[6] = read_last + 1;
  }
}

}  // namespace

void nvbk_anchors(const AnchorBatch &A, int32_t *d_anchors, int64_t *d_meta, cudaStream_t st) {
  if (A.n_reads > 0) anchors_kernel<<<(A.n_reads + 3) / 4, 128, 0, st>>>(A, d_anchors, d_meta);
}
