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
    out[i] = fmin(fmax(x, lo), hi);
  }
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

void nvbk_normalize_clip(const double *d_values, int64_t n, double shift, double scale, double lo, double hi,
                         double *d_out, cudaStream_t st) {
  if (n > 0) normalize_clip_kernel<<<grid_for(n), 256, 0, st>>>(d_values, n, shift, scale, lo, hi, d_out);
}
