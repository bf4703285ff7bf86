// reverb_b200 — sample-rate conversion to 16 kHz on the GPU (front of the path).
//
// The reference resamples with torchaudio.transforms.Resample(orig, 16000) on the CPU (asr/wenet/cli/reverb.py:125-128):
// a polyphase windowed-sinc FIR (torchaudio/functional/functional.py `_get_sinc_resample_kernel` /
// `_apply_sinc_resample_kernel`): with o = orig/gcd, n = new/gcd, width = ceil(6 o / (0.99 min(o, n))),
//   y[b*n + p] = sum_{k < 2 width + o} kernel[p][k] * xpad[b*o + k],   xpad = x zero-padded by (width, width + o)
// truncated to ceil(n * len / o) samples.  The (n, 2 width + o) kernel table is built on the host exactly like
// torchaudio builds it (reverb_b200/resample.py) — this file is the convolution: one thread per output sample, the
// phase's filter row and the signal window both stream through L1/L2 (the whole table is <= a few hundred KB).
#include "kernels.h"

namespace rvb {

template <typename T>
__global__ void __launch_bounds__(256)
resample_kernel(const T* __restrict__ x, long long n_in, const float* __restrict__ kern, int orig, int new_, int width,
                int taps, float* __restrict__ y, long long n_out) {
  const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n_out) return;
  const long long blk = j / new_;
  const int phase = (int)(j - blk * new_);
  const float* kr = kern + (size_t)phase * taps;
  const long long base = blk * orig - width;  // index into x of tap 0
  float acc = 0.f;
  int k0 = 0, k1 = taps;
  if (base < 0) k0 = (int)(-base);
  if (base + taps > n_in) k1 = (int)(n_in - base);
  for (int k = k0; k < k1; ++k) acc = fmaf(__ldg(kr + k), (float)x[base + k], acc);
  y[j] = acc;
}

int launch_resample(const void* x, int is_i16, long long n_in, const float* kern, int orig, int new_, int width, float* y,
                    long long n_out, cudaStream_t stream) {
  RVB_REQUIRE(orig >= 1 && new_ >= 1 && width >= 1, "resample: bad factors %d -> %d (width %d)", orig, new_, width);
  if (n_out <= 0) return 0;
  const int taps = 2 * width + orig;
  const unsigned grid = (unsigned)((n_out + 255) / 256);
  if (is_i16)
    resample_kernel<short><<<grid, 256, 0, stream>>>(reinterpret_cast<const short*>(x), n_in, kern, orig, new_, width, taps,
                                                     y, n_out);
  else
    resample_kernel<float><<<grid, 256, 0, stream>>>(reinterpret_cast<const float*>(x), n_in, kern, orig, new_, width, taps,
                                                     y, n_out);
  RVB_COUNT_LAUNCH();
  RVB_CHECK_LAUNCH();
  return 0;
}

}  // namespace rvb
