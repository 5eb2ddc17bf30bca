// select.cu -- pooled median / MAD normalisation of raw signals on the device (reference nadavca/read.py:67-81).
//
// Read.normalize_reads pools the raw samples of ALL reads, takes shift = median(values) and scale = median(|values -
// shift|) and clips (raw - shift) / scale to [-5, 5].  The reference does it with statistics.median over a Python list
// -- O(n log n) on the host, 4e8 values at the scale of BASELINE configs[2].  Here the two medians are EXACT order
// statistics found by a most-significant-digit radix select over the 64-bit order-preserving keys of the doubles:
// 8 passes of 8 bits, each one histogram kernel over the resident values (HBM bound, 8 B per value and pass).  The
// histogram of a pass is a 256-bin device array, so a job sharded over several GPUs all-reduces 2 KB per pass and
// normalises exactly like one host would (numpy.median = mean of the two middle elements for an even count).
#include "common.cuh"
#include "kernels.h"

namespace {

// float64 -> uint64 whose unsigned order is the numeric order (no NaNs expected)
__device__ __forceinline__ unsigned long long ordered_key(double v) {
  const unsigned long long bits = (unsigned long long)__double_as_longlong(v);
  return (bits >> 63) ? ~bits : (bits | 0x8000000000000000ull);
}

// values -> keys: mode 0 = the value itself, mode 1 = |value - shift| (the operand of the second median)
__device__ __forceinline__ unsigned long long key_of(const double *values, int64_t i, int mode, double shift) {
  const double v = values[i];
  return ordered_key(mode ? fabs(v - shift) : v);
}

// numpy.clip leaves NaN alone (a constant signal has MAD 0 and normalises to 0/0); fmin/fmax would turn it into a bound
__device__ __forceinline__ double clip_keep_nan(double x, double lo, double hi) { return x < lo ? lo : (x > hi ? hi : x); }

// Histogram of the 8 key bits below the `fixed` leading bits, over the values whose leading bits equal `prefix`.
__global__ void __launch_bounds__(256) radix_hist_kernel(const double *values, int64_t n, int mode, double shift,
                                                         unsigned long long prefix, int fixed,
                                                         unsigned long long *hist) {
  __shared__ unsigned int s_hist[256];
  s_hist[threadIdx.x] = 0;
  __syncthreads();
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int shift_bits = 56 - fixed;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += stride) {
    const unsigned long long key = key_of(values, i, mode, shift);
    if (fixed == 0 || (key >> (64 - fixed)) == prefix) atomicAdd(&s_hist[(key >> shift_bits) & 255], 1u);
  }
  __syncthreads();
  if (s_hist[threadIdx.x]) atomicAdd(&hist[threadIdx.x], (unsigned long long)s_hist[threadIdx.x]);
}

// out = clip((values - shift) / scale, lo, hi), read.py:80-81 (same IEEE operations as numpy's)
__global__ void __launch_bounds__(256) normalize_clip_kernel(const double *values, int64_t n, double shift, double scale,
                                                             double lo, double hi, double *out) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += stride) {
    const double x = (values[i] - shift) / scale;
    out[i] = clip_keep_nan(x, lo, hi);
  }
}

// Per-read median / MAD (align_signal.py:54 normalises every read on its own): one CTA per read runs the 8-pass radix
// select in shared memory, four times (two order statistics for each of the two medians when the count is even),
// then clips its read.  values / out are CSR over reads.
__device__ double block_kth(const double *values, int n, long long k, int mode, double shift, unsigned int *s_hist,
                            unsigned long long *s_prefix) {
  unsigned long long prefix = 0;
  for (int fixed = 0; fixed < 64; fixed += 8) {
    for (int i = threadIdx.x; i < 256; i += blockDim.x) s_hist[i] = 0;
    __syncthreads();
    const int shift_bits = 56 - fixed;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
      const unsigned long long key = key_of(values, i, mode, shift);
      if (fixed == 0 || (key >> (64 - fixed)) == prefix) atomicAdd(&s_hist[(key >> shift_bits) & 255], 1u);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      long long below = 0;
      int digit = 0;
      for (; digit < 255; digit++) {
        if (below + s_hist[digit] > k) break;
        below += s_hist[digit];
      }
      *s_prefix = (prefix << 8) | (unsigned long long)digit;
      s_hist[256] = (unsigned int)below;
    }
    __syncthreads();
    prefix = *s_prefix;
    k -= s_hist[256];
    __syncthreads();
  }
  const unsigned long long bits = (prefix >> 63) ? (prefix & 0x7fffffffffffffffull) : ~prefix;
  return __longlong_as_double((long long)bits);
}

__device__ double block_median(const double *values, int n, int mode, double shift, unsigned int *s_hist,
                               unsigned long long *s_prefix) {
  if (n & 1) return block_kth(values, n, n / 2, mode, shift, s_hist, s_prefix);
  const double a = block_kth(values, n, n / 2 - 1, mode, shift, s_hist, s_prefix);
  const double b = block_kth(values, n, n / 2, mode, shift, s_hist, s_prefix);
  return (a + b) / 2.0;  // numpy.median: mean of the two middle elements
}

__global__ void __launch_bounds__(256) normalize_each_kernel(const double *values, const int64_t *off, double lo,
                                                             double hi, double *out, double *shift_scale) {
  __shared__ unsigned int s_hist[257];
  __shared__ unsigned long long s_prefix;
  const int b = blockIdx.x;
  const double *v = values + off[b];
  const int n = (int)(off[b + 1] - off[b]);
  if (n <= 0) return;
  const double shift = block_median(v, n, 0, 0.0, s_hist, &s_prefix);
  const double scale = block_median(v, n, 1, shift, s_hist, &s_prefix);
  double *o = out + off[b];
  for (int i = threadIdx.x; i < n; i += blockDim.x) o[i] = clip_keep_nan((v[i] - shift) / scale, lo, hi);
  if (threadIdx.x == 0 && shift_scale) { shift_scale[2 * b] = shift; shift_scale[2 * b + 1] = scale; }
}

unsigned grid_for(int64_t n) {
  const int64_t want = (n + 255) / 256;
  return (unsigned)(want < 1 ? 1 : (want > 148 * 16 ? 148 * 16 : want));  // 16 CTAs of 256 threads per SM, grid-stride
}

}  // namespace

void nvbk_radix_hist(const double *d_values, int64_t n, int mode, double shift, unsigned long long prefix, int fixed,
                     unsigned long long *d_hist, cudaStream_t st) {
  radix_hist_kernel<<<grid_for(n), 256, 0, st>>>(d_values, n, mode, shift, prefix, fixed, d_hist);
}

void nvbk_normalize_each(const double *d_values, const int64_t *d_off, int n_reads, double lo, double hi, double *d_out,
                         double *d_shift_scale, cudaStream_t st) {
  if (n_reads > 0) normalize_each_kernel<<<n_reads, 256, 0, st>>>(d_values, d_off, lo, hi, d_out, d_shift_scale);
}

void nvbk_normalize_clip(const double *d_values, int64_t n, double shift, double scale, double lo, double hi,
                         double *d_out, cudaStream_t st) {
  if (n > 0) normalize_clip_kernel<<<grid_for(n), 256, 0, st>>>(d_values, n, shift, scale, lo, hi, d_out);
}
