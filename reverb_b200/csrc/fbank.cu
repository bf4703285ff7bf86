// reverb_b200 — Kaldi-compatible 80-bin log-mel filterbank, one fused kernel (sm_100a).
//
// Replaces `torchaudio.compliance.kaldi.fbank(waveform, num_mel_bins=80, frame_length=25, frame_shift=10,
// dither=0.0, energy_floor=0.0, sample_frequency=16000)` as called by the reference at
// asr/wenet/cli/reverb.py:130-138 (torchaudio kaldi.py:514-645): snip_edges framing (400/160), per-frame DC removal,
// pre-emphasis 0.97 (replicate-left), povey window, zero-pad to 512, |rFFT|^2, 80 triangular mel bins
// (20 Hz .. 8 kHz, mel domain), log(max(., eps)).  Input samples are int16-VALUED (not normalised).
//
// One warp per frame, 8 frames per CTA: the frame lives in shared memory from load to mel; a 512-point radix-2 FFT is
// done in place by the warp (9 stages x 8 butterflies per lane).  HBM traffic = 4 B/sample in (2 B for the int16
// entry point) + 320 B/frame out; everything else is on chip.
#include <math.h>

#include <mutex>
#include <vector>

#include "kernels.h"

namespace rvb {

constexpr int FB_WIN = 400, FB_SHIFT = 160, FB_NFFT = 512, FB_NBIN = 80, FB_MAXW = 64;
constexpr int FB_WARPS = 8;

struct FbankTables {
  float* window = nullptr;   // [400] povey
  float2* twiddle = nullptr; // [256] (cos, -sin)(2 pi k / 512)
  float* mel_w = nullptr;    // [80][FB_MAXW]
  int* mel_start = nullptr;  // [80]
  int* mel_len = nullptr;    // [80]
};
// one table set per device (a process may hold models on several GPUs); entries are written once under the mutex
constexpr int FB_MAXDEV = 64;
static FbankTables g_fb_dev[FB_MAXDEV];
static std::mutex g_fb_mutex;

static int init_fbank_tables(const FbankTables** out) {
  std::lock_guard<std::mutex> lock(g_fb_mutex);
  int dev = 0;
  RVB_CHECK_CUDA(cudaGetDevice(&dev));
  RVB_REQUIRE(dev >= 0 && dev < FB_MAXDEV, "fbank: device index %d out of range", dev);
  *out = &g_fb_dev[dev];
  if (g_fb_dev[dev].window != nullptr) return 0;
  FbankTables g_fb;  // published to g_fb_dev[dev] only when complete
  std::vector<float> win(FB_WIN);
  for (int i = 0; i < FB_WIN; ++i) {
    double h = 0.5 - 0.5 * cos(2.0 * M_PI * i / (FB_WIN - 1));
    win[i] = (float)pow(h, 0.85);
  }
  std::vector<float2> tw(FB_NFFT / 2);
  for (int k = 0; k < FB_NFFT / 2; ++k) {
    double a = 2.0 * M_PI * k / FB_NFFT;
    tw[k] = make_float2((float)cos(a), (float)(-sin(a)));
  }
  // mel banks, float32 arithmetic like torchaudio's get_mel_banks (kaldi.py:436-510)
  const float low = 20.0f, high = 8000.0f, bin_width = 16000.0f / FB_NFFT;
  const float mlow = 1127.0f * logf(1.0f + low / 700.0f), mhigh = 1127.0f * logf(1.0f + high / 700.0f);
  const float delta = (mhigh - mlow) / (FB_NBIN + 1);
  std::vector<float> w(FB_NBIN * FB_MAXW, 0.f);
  std::vector<int> st(FB_NBIN, 0), ln(FB_NBIN, 0);
  for (int b = 0; b < FB_NBIN; ++b) {
    const float left = mlow + b * delta, center = mlow + (b + 1.0f) * delta, right = mlow + (b + 2.0f) * delta;
    int first = -1, last = -1;
    std::vector<float> row(FB_NFFT / 2, 0.f);
    for (int k = 0; k < FB_NFFT / 2; ++k) {  // Nyquist bin (256) has weight 0 (kaldi.py:627)
      float mel = 1127.0f * logf(1.0f + (bin_width * k) / 700.0f);
      float up = (mel - left) / (center - left), down = (right - mel) / (right - center);
      float v = fmaxf(0.0f, fminf(up, down));
      row[k] = v;
      if (v > 0.f) {
        if (first < 0) first = k;
        last = k;
      }
    }
    if (first < 0) { first = 0; last = -1; }
    RVB_REQUIRE(last - first + 1 <= FB_MAXW, "fbank: mel bin %d too wide", b);
    st[b] = first;
    ln[b] = last - first + 1;
    for (int k = first; k <= last; ++k) w[b * FB_MAXW + (k - first)] = row[k];
  }
  RVB_CHECK_CUDA(cudaMalloc(&g_fb.window, sizeof(float) * FB_WIN));
  RVB_CHECK_CUDA(cudaMalloc(&g_fb.twiddle, sizeof(float2) * FB_NFFT / 2));
  RVB_CHECK_CUDA(cudaMalloc(&g_fb.mel_w, sizeof(float) * FB_NBIN * FB_MAXW));
  RVB_CHECK_CUDA(cudaMalloc(&g_fb.mel_start, sizeof(int) * FB_NBIN));
  RVB_CHECK_CUDA(cudaMalloc(&g_fb.mel_len, sizeof(int) * FB_NBIN));
  RVB_CHECK_CUDA(cudaMemcpy(g_fb.window, win.data(), sizeof(float) * FB_WIN, cudaMemcpyHostToDevice));
  RVB_CHECK_CUDA(cudaMemcpy(g_fb.twiddle, tw.data(), sizeof(float2) * FB_NFFT / 2, cudaMemcpyHostToDevice));
  RVB_CHECK_CUDA(cudaMemcpy(g_fb.mel_w, w.data(), sizeof(float) * FB_NBIN * FB_MAXW, cudaMemcpyHostToDevice));
  RVB_CHECK_CUDA(cudaMemcpy(g_fb.mel_start, st.data(), sizeof(int) * FB_NBIN, cudaMemcpyHostToDevice));
  RVB_CHECK_CUDA(cudaMemcpy(g_fb.mel_len, ln.data(), sizeof(int) * FB_NBIN, cudaMemcpyHostToDevice));
  g_fb_dev[dev] = g_fb;
  return 0;
}

__device__ __forceinline__ float ld_sample(const float* p, long long i) { return p[i]; }
__device__ __forceinline__ float ld_sample(const short* p, long long i) { return (float)p[i]; }

template <typename TIn>
__global__ void __launch_bounds__(FB_WARPS * 32)
fbank_kernel(const TIn* __restrict__ wave_all, long long wave_stride, long long n_frames, float* __restrict__ feats_all,
             const float* __restrict__ window, const float2* __restrict__ twiddle, const float* __restrict__ mel_w,
             const int* __restrict__ mel_start, const int* __restrict__ mel_len) {
  __shared__ float s_re[FB_WARPS][FB_NFFT];
  __shared__ float s_im[FB_WARPS][FB_NFFT];
  __shared__ float2 s_tw[FB_NFFT / 2];
  for (int i = threadIdx.x; i < FB_NFFT / 2; i += blockDim.x) s_tw[i] = twiddle[i];
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long frame = (long long)blockIdx.x * FB_WARPS + warp;
  if (frame >= n_frames) return;
  const TIn* wave = wave_all + (long long)blockIdx.y * wave_stride;      // blockIdx.y = recording in the batch
  float* feats = feats_all + (long long)blockIdx.y * n_frames * FB_NBIN;
  float* re = s_re[warp];
  float* im = s_im[warp];
  const TIn* src = wave + frame * FB_SHIFT;

  // 1. load + DC removal
  float x[13];
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < 13; ++i) {
    int j = lane + 32 * i;
    x[i] = (j < FB_WIN) ? ld_sample(src, j) : 0.f;
    sum += x[i];
  }
  const float mean = warp_sum(sum) / (float)FB_WIN;
#pragma unroll
  for (int i = 0; i < 13; ++i) {
    int j = lane + 32 * i;
    if (j < FB_WIN) im[j] = x[i] - mean;  // stash the centred frame (im is free until the FFT)
  }
  __syncwarp();
  // 2. pre-emphasis + window, written in bit-reversed order for the in-place DIT FFT
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    int j = lane + 32 * i;
    float v = 0.f;
    if (j < FB_WIN) {
      float cur = im[j];
      float prev = im[j > 0 ? j - 1 : 0];
      v = (cur - 0.97f * prev) * __ldg(window + j);
    }
    re[__brev((unsigned)j) >> 23] = v;  // 9-bit reversal
  }
  __syncwarp();
#pragma unroll
  for (int i = 0; i < 16; ++i) im[lane + 32 * i] = 0.f;
  __syncwarp();
  // 3. radix-2 DIT FFT, 9 stages
#pragma unroll 1
  for (int s = 1; s <= 9; ++s) {
    const int half = 1 << (s - 1);
    const int tstep = FB_NFFT >> s;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      int bf = lane + 32 * i;            // butterfly id 0..255
      int grp = bf >> (s - 1);
      int pos = bf & (half - 1);
      int i0 = (grp << s) + pos, i1 = i0 + half;
      float2 w = s_tw[pos * tstep];
      float xr = re[i1], xi = im[i1];
      float tr = xr * w.x - xi * w.y;
      float ti = xr * w.y + xi * w.x;
      float ur = re[i0], ui = im[i0];
      re[i0] = ur + tr;
      im[i0] = ui + ti;
      re[i1] = ur - tr;
      im[i1] = ui - ti;
    }
    __syncwarp();
  }
  // 4. power spectrum (bins 0..255; the Nyquist bin carries zero mel weight)
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    int k = lane + 32 * i;
    float a = re[k], b = im[k];
    re[k] = a * a + b * b;
  }
  __syncwarp();
  // 5. mel + log
  for (int b = lane; b < FB_NBIN; b += 32) {
    const int st = __ldg(mel_start + b), ln = __ldg(mel_len + b);
    float acc = 0.f;
    for (int k = 0; k < ln; ++k) acc = fmaf(re[st + k], __ldg(mel_w + b * FB_MAXW + k), acc);
    feats[frame * FB_NBIN + b] = logf(fmaxf(acc, 1.1920928955078125e-07f));
  }
}

template <typename TIn>
static int launch_fbank_t(const TIn* wave, long long n_samples, float* feats, long long n_frames, cudaStream_t stream,
                          int batch = 1, long long wave_stride = 0) {
  const FbankTables* fb = nullptr;
  if (init_fbank_tables(&fb)) return -1;
  const FbankTables& g_fb = *fb;
  long long expect = n_samples < FB_WIN ? 0 : 1 + (n_samples - FB_WIN) / FB_SHIFT;
  RVB_REQUIRE(n_frames <= expect, "fbank: %lld frames requested but only %lld fit %lld samples", n_frames, expect,
              n_samples);
  if (n_frames <= 0) return 0;
  const long long blocks = (n_frames + FB_WARPS - 1) / FB_WARPS;
  dim3 grid((unsigned)blocks, (unsigned)batch);
  fbank_kernel<TIn><<<grid, FB_WARPS * 32, 0, stream>>>(wave, wave_stride, n_frames, feats, g_fb.window, g_fb.twiddle,
                                                       g_fb.mel_w, g_fb.mel_start, g_fb.mel_len);
  RVB_COUNT_LAUNCH();
  RVB_CHECK_LAUNCH();
  return 0;
}

int launch_fbank(const float* wave, long long n_samples, float* feats, long long n_frames, cudaStream_t stream) {
  return launch_fbank_t<float>(wave, n_samples, feats, n_frames, stream);
}
int launch_fbank_i16(const short* wave, long long n_samples, float* feats, long long n_frames, cudaStream_t stream) {
  return launch_fbank_t<short>(wave, n_samples, feats, n_frames, stream);
}
int launch_fbank_batch(const void* wave, int is_i16, int batch, long long wave_stride, long long n_samples,
                       float* feats, long long n_frames, cudaStream_t stream) {
  RVB_REQUIRE(batch >= 1 && batch <= 65535 && wave_stride >= n_samples, "fbank_batch: bad batch/stride");
  if (is_i16)
    return launch_fbank_t<short>(reinterpret_cast<const short*>(wave), n_samples, feats, n_frames, stream, batch,
                                 wave_stride);
  return launch_fbank_t<float>(reinterpret_cast<const float*>(wave), n_samples, feats, n_frames, stream, batch,
                               wave_stride);
}

}  // namespace rvb
