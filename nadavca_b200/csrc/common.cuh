// common.cuh -- device-side data layout and helpers shared by all kernels of libnadavca_b200.so (sm_100a).
//
// Data layout in HBM (all per batch, CSR over reads; see DESIGN.md "Data layout"):
//   signal   double[sum N]          sig_off[b]   .. sig_off[b+1]
//   ref      int32 [sum n]          ref_off[b]   .. ref_off[b+1]       numeric bases
//   bs, be   int32 [sum (n+1)]      index ref_off[b] + b + j           inclusive band of band-row j = 0..n
//   cell_off int64 [sum (n+2)]      index ref_off[b] + 2b + j          exclusive scan of band widths (j = n+1: total)
//   prefix/suffix matrices: packed band rows, row j of read b at mat_base[b] + cell_off[j] + (c - bs[j])
//   (with transitions the 2n rows interleave band rows i and i+1, see trans_row_off()).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define NVB_WARP 32
#define NVB_FULL 0xffffffffu

enum { NVB_MODE_PLAIN = 0, NVB_MODE_TRANS = 1, NVB_MODE_WOBBLE = 2 };
enum { NVB_EM_GAUSS = 0, NVB_EM_MIX = 1, NVB_EM_CONST = 2 };

struct ModelDev {
  int k, central, alphabet;
  int64_t n_kmers;
  const double *mean, *ac, *mc;  // per k-mer tables (kmer_model.cpp:6-14)
  double log_p_in;               // log(0.01), kmer_model.cpp:79
};

struct BatchDev {
  int n_reads;
  const double *signal;
  const int64_t *sig_off;
  const int32_t *ref;
  const int64_t *ref_off;
  const int32_t *ctxb;
  const int64_t *ctxb_off;
  const int32_t *ctxa;
  const int64_t *ctxa_off;
  const int32_t *anchors;
  const int64_t *anc_off;
  int bandwidth, mel;
  int32_t *bs, *be;
  int64_t *cell_off;
  int32_t *flags;      // per read NVB_READ_* (bad band)
  int32_t *max_width;  // per read
  const double *row_emis;  // per reference position [mu, ac * S, mc * S, same-mean flags]: Gaussian emission of the unmodified k-mer
                           // there, ac and mc scaled for exp_ext_scaled (dp3.cuh); filled once per batch (band.cu)
};

// Inputs of the batched anchor construction (anchors.cu), CSR over reads.
struct AnchorBatch {
  int n_reads;
  const int32_t *cigar_len;   // operations of every read's CIGAR ...
  const int8_t *cigar_op;     // ... 0 = M, 1 = I, 2 = D, 3 = S
  const int64_t *cigar_off;
  const int64_t *mapped_pos;  // 0-based position of the hit in the contig
  const int32_t *reverse;
  const int8_t *read_seq;     // basecalled bases 0..3 ...
  const int32_t *mapping;     // ... and the sample index of each base (-1: not placed), Read.sequence_to_signal_mapping
  const int64_t *read_off;
  const int32_t *n_signal;    // len(read.normalized_signal)
  const int8_t *genome;       // contig bases 0..3 (4 = other)
  int64_t genome_len;
  int bandwidth;
};

struct ReadView {
  int n, N, nb, na;
  const double *sig;
  const int32_t *ref, *cb, *ca;
  const int32_t *bs, *be;
  const int64_t *coff;
  const double *emis;  // BatchDev::row_emis of this read
};

__device__ __forceinline__ double nvb_neg_inf() { return __longlong_as_double(0xfff0000000000000LL); }

__device__ __forceinline__ ReadView read_view(const BatchDev &B, int b) {
  ReadView v;
  int64_t r0 = B.ref_off[b];
  v.n = (int)(B.ref_off[b + 1] - r0);
  v.N = (int)(B.sig_off[b + 1] - B.sig_off[b]);
  v.nb = (int)(B.ctxb_off[b + 1] - B.ctxb_off[b]);
  v.na = (int)(B.ctxa_off[b + 1] - B.ctxa_off[b]);
  v.sig = B.signal + B.sig_off[b];
  v.ref = B.ref + r0;
  v.cb = B.ctxb + B.ctxb_off[b];
  v.ca = B.ctxa + B.ctxa_off[b];
  v.bs = B.bs + r0 + b;
  v.be = B.be + r0 + b;
  v.coff = B.cell_off + r0 + 2 * (int64_t)b;
  v.emis = B.row_emis + 4 * r0;
  return v;
}

// ExtendedSequence / ModifiedSequence (sequence.cpp:6-38): context ++ reference ++ context, base 0 outside,
// one optional overridden position.
__device__ __forceinline__ int base_at(const ReadView &v, int idx, int mod_pos, int mod_val) {
  if (idx == mod_pos) return mod_val;
  int j = idx + v.nb;
  if (j < 0 || j >= v.nb + v.n + v.na) return 0;
  if (j < v.nb) return v.cb[j];
  if (j < v.nb + v.n) return v.ref[j - v.nb];
  return v.ca[j - v.nb - v.n];
}

// KmerModel::GetKmerId (kmer_model.cpp:22-30)
__device__ __forceinline__ int kmer_id(const ModelDev &M, const ReadView &v, int i, int mod_pos, int mod_val) {
  int id = 0;
  for (int p = i - M.central; p < i - M.central + M.k; p++) id = id * M.alphabet + base_at(v, p, mod_pos, mod_val);
  return id;
}

struct Emis {
  int kind;
  double mu1, ac1, mc1, mu2, ac2, mc2;  // EM_CONST keeps its value in ac1
};

__device__ __forceinline__ Emis emis_gauss(const ModelDev &M, const ReadView &v, int i, int mp, int mv) {
  Emis e;
  int id = kmer_id(M, v, i, mp, mv);
  e.kind = NVB_EM_GAUSS;
  e.mu1 = M.mean[id]; e.ac1 = M.ac[id]; e.mc1 = M.mc[id];
  e.mu2 = 0; e.ac2 = 0; e.mc2 = 0;
  return e;
}
__device__ __forceinline__ Emis emis_mix(const ModelDev &M, const ReadView &v, int i1, int i2, int mp, int mv) {
  Emis e;
  int a = kmer_id(M, v, i1, mp, mv), b = kmer_id(M, v, i2, mp, mv);
  e.kind = NVB_EM_MIX;
  e.mu1 = M.mean[a]; e.ac1 = M.ac[a]; e.mc1 = M.mc[a];
  e.mu2 = M.mean[b]; e.ac2 = M.ac[b]; e.mc2 = M.mc[b];
  return e;
}
// GetTransitionDistribution (kmer_model.cpp:64-94): log(0) for equal means, else the constant log(0.01).
__device__ __forceinline__ Emis emis_transition(const ModelDev &M, const ReadView &v, int i1, int i2) {
  Emis e;
  double m1 = M.mean[kmer_id(M, v, i1, INT32_MIN, 0)], m2 = M.mean[kmer_id(M, v, i2, INT32_MIN, 0)];
  e.kind = NVB_EM_CONST;
  e.ac1 = (m1 == m2) ? nvb_neg_inf() : M.log_p_in;
  e.mu1 = 0; e.mc1 = 0; e.mu2 = 0; e.ac2 = 0; e.mc2 = 0;
  return e;
}

// Probability::operator+ (probability.cpp:33-40): a + log(1 + exp(b - a)) with the larger operand first.
__device__ __forceinline__ double lp_add(double a, double b) {
  double hi = a < b ? b : a;
  double lo = a < b ? a : b;
  if (lo == nvb_neg_inf()) return hi;
  return hi + log(1 + exp(lo - hi));
}

// Files are compiled with -fmad=false so that ac - d*d*mc rounds like the reference's scalar code.
__device__ __forceinline__ double emis_eval(const Emis &e, double x) {
  if (e.kind == NVB_EM_CONST) return e.ac1;
  double d = x - e.mu1;
  double l1 = e.ac1 - d * d * e.mc1;
  if (e.kind == NVB_EM_GAUSS) return l1;
  double d2 = x - e.mu2;
  double l2 = e.ac2 - d2 * d2 * e.mc2;
  return lp_add(l1, l2) - 2.0;  // kmer_model.cpp:60, "/ 2" is "- 2.0" in log space
}

// Row offset of row rho (0..2n-1) in the transition layout: rows 2i / 2i+1 live on band rows i / i+1.
__device__ __forceinline__ int64_t trans_row_off(const ReadView &v, int rho) {
  int i = rho >> 1;
  if (rho & 1) return 2 * v.coff[i + 1] - v.coff[1];
  return v.coff[i] + v.coff[i + 1] - v.coff[1];
}
