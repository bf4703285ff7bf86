// reverb_b200 — HBM-bound kernels of the Conformer stack: LayerNorm (single / fused pair), CMVN+conv1+ReLU,
// GLU + depthwise conv + norm + SiLU, decoder embedding, small casts.  Warp-shuffle reductions, 128-bit loads.
#include <math.h>

#include "kernels.h"

namespace rvb {

// ---------------------------------------------------------------------------------------------------------------
// LayerNorm: one warp per row, row cached in registers (NV float4 per lane), two-pass variance like ATen.
// reference: nn.LayerNorm uses of transformer/encoder_layer.py:149-159, encoder.py:107, decoder_layer.py:53-55,241-243
template <int NV>
__global__ void __launch_bounds__(256) layernorm_kernel(const float* __restrict__ x, const float* __restrict__ gamma,
                                                        const float* __restrict__ beta, float eps, int M, int d,
                                                        bf16* __restrict__ out_bf16, float* __restrict__ out_f32,
                                                        const int* __restrict__ row_lens, int rows_per_batch,
                                                        int mask_rows) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= M) return;
  const int nvec = d >> 2;
  const float4* xr = reinterpret_cast<const float4*>(x + (long long)warp * d);
  bool masked = false;
  if (mask_rows && row_lens != nullptr) {
    int b = warp / rows_per_batch;
    masked = (warp - b * rows_per_batch) >= __ldg(row_lens + b);
  }
  float4 v[NV];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    int idx = lane + 32 * i;
    if (idx < nvec) {
      v[i] = xr[idx];
      s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    } else {
      v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
  const float mean = warp_sum(s) / (float)d;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    int idx = lane + 32 * i;
    if (idx < nvec) {
      float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, e = v[i].w - mean;
      q += (a * a + b * b) + (c * c + e * e);
    }
  }
  const float rstd = rsqrtf(warp_sum(q) / (float)d + eps);
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    int idx = lane + 32 * i;
    if (idx < nvec) {
      float4 g = __ldg(reinterpret_cast<const float4*>(gamma) + idx);
      float4 bb = __ldg(reinterpret_cast<const float4*>(beta) + idx);
      float4 o;
      o.x = (v[i].x - mean) * rstd * g.x + bb.x;
      o.y = (v[i].y - mean) * rstd * g.y + bb.y;
      o.z = (v[i].z - mean) * rstd * g.z + bb.z;
      o.w = (v[i].w - mean) * rstd * g.w + bb.w;
      if (masked) o = make_float4(0.f, 0.f, 0.f, 0.f);
      if (out_f32) reinterpret_cast<float4*>(out_f32 + (long long)warp * d)[idx] = o;
      if (out_bf16) {
        uint2 u;
        u.x = pack_bf16x2(o.x, o.y);
        u.y = pack_bf16x2(o.z, o.w);
        reinterpret_cast<uint2*>(out_bf16 + (long long)warp * d)[idx] = u;
      }
    }
  }
}

int launch_layernorm(const float* x, const float* gamma, const float* beta, float eps, int M, int d, bf16* out_bf16,
                     float* out_f32, const int* row_lens, int rows_per_batch, int mask_rows, cudaStream_t stream) {
  RVB_REQUIRE(d % 4 == 0 && d <= 4096, "layernorm: d=%d unsupported (need d %% 4 == 0, d <= 4096)", d);
  if (M <= 0) return 0;
  const int grid = (M + 7) / 8;
  const int nv = (d / 4 + 31) / 32;
  if (rows_per_batch <= 0) rows_per_batch = M;
#define RVB_LN(NV)                                                                                           \
  layernorm_kernel<NV><<<grid, 256, 0, stream>>>(x, gamma, beta, eps, M, d, out_bf16, out_f32, row_lens,     \
                                                 rows_per_batch, mask_rows)
  if (nv <= 1) RVB_LN(1);
  else if (nv <= 2) RVB_LN(2);
  else if (nv <= 4) RVB_LN(4);
  else if (nv <= 8) RVB_LN(8);
  else if (nv <= 16) RVB_LN(16);
  else RVB_LN(32);
#undef RVB_LN
  RVB_COUNT_LAUNCH();
  RVB_CHECK_LAUNCH();
  return 0;
}

// x2 = LN_a(x) (+ y_add) ; n = LN_b(x2).  Fuses `norm_final` of block i (and the LSL `x = x + y`,
// encoder_layer.py:397-400) with the first pre-norm of block i+1 (or encoder.after_norm): one read, two writes.
template <int NV>
__global__ void __launch_bounds__(256)
double_layernorm_kernel(const float* __restrict__ x, const float* __restrict__ ga, const float* __restrict__ ba,
                        const float* __restrict__ y_add, float* __restrict__ x2, const float* __restrict__ gb,
                        const float* __restrict__ bb, float eps, int M, int d, bf16* __restrict__ n_out,
                        float* __restrict__ n_out_f32) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= M) return;
  const int nvec = d >> 2;
  const float4* xr = reinterpret_cast<const float4*>(x + (long long)warp * d);
  float4 v[NV];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    int idx = lane + 32 * i;
    if (idx < nvec) {
      v[i] = xr[idx];
      s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    } else {
      v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
  float mean = warp_sum(s) / (float)d;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    int idx = lane + 32 * i;
    if (idx < nvec) {
      float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, e = v[i].w - mean;
      q += (a * a + b * b) + (c * c + e * e);
    }
  }
  float rstd = rsqrtf(warp_sum(q) / (float)d + eps);
  s = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    int idx = lane + 32 * i;
    if (idx < nvec) {
      float4 g = __ldg(reinterpret_cast<const float4*>(ga) + idx);
      float4 b4 = __ldg(reinterpret_cast<const float4*>(ba) + idx);
      float4 o;
      o.x = (v[i].x - mean) * rstd * g.x + b4.x;
      o.y = (v[i].y - mean) * rstd * g.y + b4.y;
      o.z = (v[i].z - mean) * rstd * g.z + b4.z;
      o.w = (v[i].w - mean) * rstd * g.w + b4.w;
      if (y_add) {
        float4 y = reinterpret_cast<const float4*>(y_add + (long long)warp * d)[idx];
        o.x += y.x;
        o.y += y.y;
        o.z += y.z;
        o.w += y.w;
      }
      reinterpret_cast<float4*>(x2 + (long long)warp * d)[idx] = o;
      v[i] = o;
      s += (o.x + o.y) + (o.z + o.w);
    }
  }
  if (n_out == nullptr && n_out_f32 == nullptr) return;
  mean = warp_sum(s) / (float)d;
  q = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    int idx = lane + 32 * i;
    if (idx < nvec) {
      float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, e = v[i].w - mean;
      q += (a * a + b * b) + (c * c + e * e);
    }
  }
  rstd = rsqrtf(warp_sum(q) / (float)d + eps);
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    int idx = lane + 32 * i;
    if (idx < nvec) {
      float4 g = __ldg(reinterpret_cast<const float4*>(gb) + idx);
      float4 b4 = __ldg(reinterpret_cast<const float4*>(bb) + idx);
      float4 o;
      o.x = (v[i].x - mean) * rstd * g.x + b4.x;
      o.y = (v[i].y - mean) * rstd * g.y + b4.y;
      o.z = (v[i].z - mean) * rstd * g.z + b4.z;
      o.w = (v[i].w - mean) * rstd * g.w + b4.w;
      if (n_out_f32) reinterpret_cast<float4*>(n_out_f32 + (long long)warp * d)[idx] = o;
      if (n_out) {
        uint2 u;
        u.x = pack_bf16x2(o.x, o.y);
        u.y = pack_bf16x2(o.z, o.w);
        reinterpret_cast<uint2*>(n_out + (long long)warp * d)[idx] = u;
      }
    }
  }
}

int launch_double_layernorm(const float* x, const float* ga, const float* ba, const float* y_add, float* x2,
                            const float* gb, const float* bb, float eps, int M, int d, bf16* n_out,
                            float* n_out_f32, cudaStream_t stream) {
  RVB_REQUIRE(d % 4 == 0 && d <= 4096, "double_layernorm: d=%d unsupported", d);
  if (M <= 0) return 0;
  const int grid = (M + 7) / 8;
  const int nv = (d / 4 + 31) / 32;
#define RVB_DLN(NV) \
  double_layernorm_kernel<NV><<<grid, 256, 0, stream>>>(x, ga, ba, y_add, x2, gb, bb, eps, M, d, n_out, n_out_f32)
  if (nv <= 1) RVB_DLN(1);
  else if (nv <= 2) RVB_DLN(2);
  else if (nv <= 4) RVB_DLN(4);
  else if (nv <= 8) RVB_DLN(8);
  else if (nv <= 16) RVB_DLN(16);
  else RVB_DLN(32);
#undef RVB_DLN
  RVB_COUNT_LAUNCH();
  RVB_CHECK_LAUNCH();
  return 0;
}

// ---------------------------------------------------------------------------------------------------------------
// GlobalCMVN + Conv2d(1, C, 3, stride 2) + ReLU  (transformer/cmvn.py:36-47, subsampling.py:186-187).
// Output is channels-last bf16 with the time axis split by parity, (B, 2, T1h, F1, C), so that the second conv's
// implicit-GEMM A tiles (gemm.cu conv_mode) are plain unit-stride 4-D TMA boxes.  One CTA per (b, t1).
__global__ void __launch_bounds__(256)
conv1_kernel(const float* __restrict__ feats, const float* __restrict__ mean, const float* __restrict__ istd,
             const float* __restrict__ w, const float* __restrict__ bias, bf16* __restrict__ out, int T, int F, int C,
             int T1, int T1h, int F1) {
  extern __shared__ float s_in[];  // 3 rows x F, CMVN applied
  const int t1 = blockIdx.x;       // 0 .. 2*T1h-1
  const int b = blockIdx.y;
  const int par = t1 & 1, th = t1 >> 1;
  bf16* orow = out + (((long long)(b * 2 + par) * T1h + th) * F1) * C;
  const int CG = C >> 3;
  if (t1 >= T1) {  // padding row (T1 odd): keep it finite
    for (int i = threadIdx.x; i < F1 * CG; i += blockDim.x)
      reinterpret_cast<uint4*>(orow)[i] = make_uint4(0u, 0u, 0u, 0u);
    return;
  }
  for (int i = threadIdx.x; i < 3 * F; i += blockDim.x) {
    int kh = i / F, f = i - kh * F;
    int t = 2 * t1 + kh;
    float x = (t < T) ? feats[((long long)b * T + t) * F + f] : 0.f;
    s_in[i] = (x - __ldg(mean + f)) * __ldg(istd + f);
  }
  __syncthreads();
  const int cg = threadIdx.x % CG;
  const int fstep = blockDim.x / CG;
  float wr[8][9], br[8];
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    br[c] = __ldg(bias + cg * 8 + c);
#pragma unroll
    for (int k = 0; k < 9; ++k) wr[c][k] = __ldg(w + (cg * 8 + c) * 9 + k);
  }
  for (int f = threadIdx.x / CG; f < F1; f += fstep) {
    float in[9];
#pragma unroll
    for (int kh = 0; kh < 3; ++kh)
#pragma unroll
      for (int kw = 0; kw < 3; ++kw) in[kh * 3 + kw] = s_in[kh * F + 2 * f + kw];
    float o[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      float a = br[c];
#pragma unroll
      for (int k = 0; k < 9; ++k) a = fmaf(wr[c][k], in[k], a);
      o[c] = fmaxf(a, 0.f);
    }
    uint4 u;
    u.x = pack_bf16x2(o[0], o[1]);
    u.y = pack_bf16x2(o[2], o[3]);
    u.z = pack_bf16x2(o[4], o[5]);
    u.w = pack_bf16x2(o[6], o[7]);
    reinterpret_cast<uint4*>(orow + (long long)f * C)[cg] = u;
  }
}

int launch_conv1(const float* feats, const float* mean, const float* istd, const float* w, const float* bias,
                 bf16* out, int B, int T, int F, int C, int T1, int T1h, int F1, cudaStream_t stream) {
  RVB_REQUIRE(C % 8 == 0 && C / 8 <= 1024, "conv1: C=%d unsupported", C);
  const int CG = C / 8;
  int threads = CG;
  while (threads < 256 && threads * 2 <= 1024) threads *= 2;
  if (threads > 1024) threads = CG;
  dim3 grid(2 * T1h, B);
  conv1_kernel<<<grid, threads, 3 * F * sizeof(float), stream>>>(feats, mean, istd, w, bias, out, T, F, C, T1, T1h,
                                                                 F1);
  RVB_COUNT_LAUNCH();
  RVB_CHECK_LAUNCH();
  return 0;
}

// ---------------------------------------------------------------------------------------------------------------
// Conformer convolution module, middle part (transformer/convolution.py:129-138):
//   GLU over channels -> depthwise conv (K taps, causal left pad K-1 or symmetric (K-1)/2) + bias
//   -> LayerNorm over channels (or BatchNorm1d eval) -> SiLU.
// Input (B, T, 2C) bf16 = pointwise_conv1 output incl. bias; output (B, T, C) bf16 feeds pointwise_conv2.
// One CTA per (b, 16 output frames): GLU'd halo rows staged in smem (bf16), conv results in smem (fp32),
// then one warp per frame does the channel reduction.  Layout stays (B, T, C): no transposes.
constexpr int CM_TT = 16;

// Fast path: K compile-time, each thread owns NIT channel pairs and keeps its CM_TT x 2 x NIT conv outputs in
// registers (fully unrolled sliding window), LayerNorm statistics by two block reductions; only the GLU'd halo rows
// live in shared memory (bf16), so 3 CTAs fit per SM at C = 1024.
template <int K, int NIT, int NT>
__global__ void __launch_bounds__(NT)
conv_mid_fast_kernel(const bf16* __restrict__ x, const float* __restrict__ pw1_bias, const float* __restrict__ dw_w,
                     const float* __restrict__ dw_b,
                     const float* __restrict__ norm_w, const float* __restrict__ norm_b,
                     const float* __restrict__ bn_mean, const float* __restrict__ bn_var, int use_ln, float eps,
                     bf16* __restrict__ out, int T, int C, int causal) {
  constexpr int ROWS = CM_TT + K - 1;
  extern __shared__ __align__(16) uint8_t smem_cm[];
  uint32_t* s_glu = reinterpret_cast<uint32_t*>(smem_cm);  // [ROWS][C/2] packed bf16 pairs
  __shared__ float s_part[NT / 32][CM_TT];
  __shared__ float s_stat[2][CM_TT];
  const int b = blockIdx.y;
  const int t0 = blockIdx.x * CM_TT;
  const int left = causal ? (K - 1) : (K - 1) / 2;
  const int C2 = C >> 1;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // staging: batches of 8 independent (row, channel-pair) items per thread so that 16 loads are in flight at once
  constexpr int U = 8;
  for (int base = 0; base < ROWS * C2; base += NT * U) {
    uint32_t ra[U], rg[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int i = base + u * NT + threadIdx.x;
      ra[u] = 0u;
      rg[u] = 0u;
      if (i < ROWS * C2) {
        const int r = i / C2, cp = i - r * C2;
        const int t = t0 - left + r;
        if (t >= 0 && t < T) {
          const uint32_t* xr = reinterpret_cast<const uint32_t*>(x + ((long long)b * T + t) * (2 * C));
          ra[u] = xr[cp];
          rg[u] = xr[C2 + cp];
        }
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int i = base + u * NT + threadIdx.x;
      if (i >= ROWS * C2) continue;
      const int r = i / C2, cp = i - r * C2;
      const int t = t0 - left + r;
      uint32_t o = 0u;
      if (t >= 0 && t < T) {
        float2 a = unpack_bf16x2(ra[u]);
        float2 g = unpack_bf16x2(rg[u]);
        o = pack_bf16x2(a.x * sigmoid_f(g.x), a.y * sigmoid_f(g.y));
      } else if (t < 0 && causal) {
        // the reference left-pads K-1 zero frames BEFORE pointwise_conv1 (convolution.py:113-114,129-130), so the
        // pad frames reach the depthwise conv as GLU(bias), not as zeros
        float2 a = unpack_bf16x2(pack_bf16x2(__ldg(pw1_bias + 2 * cp), __ldg(pw1_bias + 2 * cp + 1)));
        float2 g = unpack_bf16x2(pack_bf16x2(__ldg(pw1_bias + C + 2 * cp), __ldg(pw1_bias + C + 2 * cp + 1)));
        o = pack_bf16x2(a.x * sigmoid_f(g.x), a.y * sigmoid_f(g.y));
      }
      s_glu[i] = o;
    }
  }
  __syncthreads();
  float acc[NIT][CM_TT][2];
#pragma unroll
  for (int it = 0; it < NIT; ++it) {
    const int cp = threadIdx.x + it * NT;
    const bool ok = cp < C2;
    float w0[K], w1[K];
    float b0 = 0.f, b1 = 0.f;
    if (ok) {
      b0 = __ldg(dw_b + 2 * cp);
      b1 = __ldg(dw_b + 2 * cp + 1);
#pragma unroll
      for (int k = 0; k < K; ++k) {
        w0[k] = __ldg(dw_w + (2 * cp) * K + k);
        w1[k] = __ldg(dw_w + (2 * cp + 1) * K + k);
      }
    } else {
#pragma unroll
      for (int k = 0; k < K; ++k) w0[k] = w1[k] = 0.f;
    }
#pragma unroll
    for (int t = 0; t < CM_TT; ++t) {
      acc[it][t][0] = b0;
      acc[it][t][1] = b1;
    }
#pragma unroll
    for (int r = 0; r < ROWS; ++r) {
      float2 v = ok ? unpack_bf16x2(s_glu[r * C2 + cp]) : make_float2(0.f, 0.f);
#pragma unroll
      for (int k = 0; k < K; ++k) {
        constexpr int dummy = 0;
        (void)dummy;
        const int t = r - k;  // compile-time after unrolling
        if (t >= 0 && t < CM_TT) {
          acc[it][t][0] = fmaf(w0[k], v.x, acc[it][t][0]);
          acc[it][t][1] = fmaf(w1[k], v.y, acc[it][t][1]);
        }
      }
    }
  }
  float mean[CM_TT], rstd[CM_TT];
  if (use_ln) {
    // pass 1: mean over channels
#pragma unroll
    for (int t = 0; t < CM_TT; ++t) {
      float s = 0.f;
#pragma unroll
      for (int it = 0; it < NIT; ++it)
        if (threadIdx.x + it * NT < C2) s += acc[it][t][0] + acc[it][t][1];
      s = warp_sum(s);
      if (lane == 0) s_part[warp][t] = s;
    }
    __syncthreads();
    if (threadIdx.x < CM_TT) {
      float s = 0.f;
      for (int w = 0; w < NT / 32; ++w) s += s_part[w][threadIdx.x];
      s_stat[0][threadIdx.x] = s / (float)C;
    }
    __syncthreads();
#pragma unroll
    for (int t = 0; t < CM_TT; ++t) mean[t] = s_stat[0][t];
    // pass 2: variance (two-pass, like ATen)
#pragma unroll
    for (int t = 0; t < CM_TT; ++t) {
      float q = 0.f;
#pragma unroll
      for (int it = 0; it < NIT; ++it)
        if (threadIdx.x + it * NT < C2) {
          float d0 = acc[it][t][0] - mean[t], d1 = acc[it][t][1] - mean[t];
          q += d0 * d0 + d1 * d1;
        }
      q = warp_sum(q);
      if (lane == 0) s_part[warp][t] = q;
    }
    __syncthreads();
    if (threadIdx.x < CM_TT) {
      float q = 0.f;
      for (int w = 0; w < NT / 32; ++w) q += s_part[w][threadIdx.x];
      s_stat[1][threadIdx.x] = rsqrtf(q / (float)C + eps);
    }
    __syncthreads();
#pragma unroll
    for (int t = 0; t < CM_TT; ++t) rstd[t] = s_stat[1][t];
  }
#pragma unroll
  for (int it = 0; it < NIT; ++it) {
    const int cp = threadIdx.x + it * NT;
    if (cp >= C2) continue;
    const int c0 = 2 * cp, c1 = 2 * cp + 1;
    const float g0 = __ldg(norm_w + c0), g1 = __ldg(norm_w + c1), be0 = __ldg(norm_b + c0), be1 = __ldg(norm_b + c1);
    float m0 = 0.f, m1 = 0.f, r0 = 1.f, r1 = 1.f;
    if (!use_ln) {
      m0 = __ldg(bn_mean + c0);
      m1 = __ldg(bn_mean + c1);
      r0 = rsqrtf(__ldg(bn_var + c0) + eps);
      r1 = rsqrtf(__ldg(bn_var + c1) + eps);
    }
#pragma unroll
    for (int t = 0; t < CM_TT; ++t) {
      if (t0 + t >= T) continue;
      float y0, y1;
      if (use_ln) {
        y0 = (acc[it][t][0] - mean[t]) * rstd[t] * g0 + be0;
        y1 = (acc[it][t][1] - mean[t]) * rstd[t] * g1 + be1;
      } else {
        y0 = (acc[it][t][0] - m0) * r0 * g0 + be0;
        y1 = (acc[it][t][1] - m1) * r1 * g1 + be1;
      }
      reinterpret_cast<uint32_t*>(out + ((long long)b * T + t0 + t) * C)[cp] = pack_bf16x2(silu_f(y0), silu_f(y1));
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Split version (default for K in {7,15,31}): kernel A = GLU + depthwise conv for one (batch, 16-frame tile,
// 128-channel-pair slice) per 128-thread CTA — no block-wide phases, so the loads of one CTA overlap the FMAs of the
// others (4+ CTAs / SM) — writes the fp32 conv result and accumulates per-frame sum / sum-of-squares with one atomic
// pair per (frame, CTA); kernel B = LayerNorm(from the accumulated statistics) + SiLU -> bf16, one warp per frame.
// With BatchNorm (eval) kernel A applies norm + SiLU itself and writes bf16 directly.
template <int K>
__global__ void __launch_bounds__(128)
conv_dw_kernel(const bf16* __restrict__ x, const float* __restrict__ pw1_bias, const float* __restrict__ dw_w,
               const float* __restrict__ dw_b, const float* __restrict__ norm_w, const float* __restrict__ norm_b,
               const float* __restrict__ bn_mean, const float* __restrict__ bn_var, int use_ln, float eps,
               float* __restrict__ conv_out, float* __restrict__ stats, bf16* __restrict__ out, int T, int C,
               int causal) {
  constexpr int ROWS = CM_TT + K - 1;
  __shared__ float s_part[4][CM_TT][2];
  const int b = blockIdx.y;
  const int t0 = blockIdx.x * CM_TT;
  const int C2 = C >> 1;
  const int cp = blockIdx.z * 128 + threadIdx.x;
  const bool ok = cp < C2;
  const int left = causal ? (K - 1) : (K - 1) / 2;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint32_t ra[ROWS], rg[ROWS];
#pragma unroll
  for (int r = 0; r < ROWS; ++r) {
    const int t = t0 - left + r;
    ra[r] = 0u;
    rg[r] = 0u;
    if (ok && t >= 0 && t < T) {
      const uint32_t* xr = reinterpret_cast<const uint32_t*>(x + ((long long)b * T + t) * (2 * C));
      ra[r] = __ldg(xr + cp);
      rg[r] = __ldg(xr + C2 + cp);
    }
  }
  float w0[K], w1[K];
  float b0 = 0.f, b1 = 0.f;
  float2 padv = make_float2(0.f, 0.f);
  if (ok) {
    b0 = __ldg(dw_b + 2 * cp);
    b1 = __ldg(dw_b + 2 * cp + 1);
#pragma unroll
    for (int k = 0; k < K; ++k) {
      w0[k] = __ldg(dw_w + (2 * cp) * K + k);
      w1[k] = __ldg(dw_w + (2 * cp + 1) * K + k);
    }
    if (causal) {
      // the reference left-pads K-1 zero frames BEFORE pointwise_conv1 (convolution.py:113-114,129-130): the pad
      // frames reach the depthwise conv as GLU(bias), not as zeros
      float2 a = unpack_bf16x2(pack_bf16x2(__ldg(pw1_bias + 2 * cp), __ldg(pw1_bias + 2 * cp + 1)));
      float2 g = unpack_bf16x2(pack_bf16x2(__ldg(pw1_bias + C + 2 * cp), __ldg(pw1_bias + C + 2 * cp + 1)));
      padv = unpack_bf16x2(pack_bf16x2(a.x * sigmoid_f(g.x), a.y * sigmoid_f(g.y)));
    }
  } else {
#pragma unroll
    for (int k = 0; k < K; ++k) w0[k] = w1[k] = 0.f;
  }
  float acc[CM_TT][2];
#pragma unroll
  for (int t = 0; t < CM_TT; ++t) {
    acc[t][0] = b0;
    acc[t][1] = b1;
  }
#pragma unroll
  for (int r = 0; r < ROWS; ++r) {
    const int tt = t0 - left + r;
    float2 v;
    if (tt >= 0 && tt < T) {
      float2 a = unpack_bf16x2(ra[r]);
      float2 g = unpack_bf16x2(rg[r]);
      v = unpack_bf16x2(pack_bf16x2(a.x * sigmoid_f(g.x), a.y * sigmoid_f(g.y)));  // GLU output is stored as bf16
    } else {
      v = (tt < 0) ? padv : make_float2(0.f, 0.f);
    }
#pragma unroll
    for (int k = 0; k < K; ++k) {
      const int t = r - k;  // compile-time after unrolling
      if (t >= 0 && t < CM_TT) {
        acc[t][0] = fmaf(w0[k], v.x, acc[t][0]);
        acc[t][1] = fmaf(w1[k], v.y, acc[t][1]);
      }
    }
  }
  if (!use_ln) {
    if (!ok) return;
    const int c0 = 2 * cp, c1 = 2 * cp + 1;
    const float g0 = __ldg(norm_w + c0), g1 = __ldg(norm_w + c1), be0 = __ldg(norm_b + c0), be1 = __ldg(norm_b + c1);
    const float m0 = __ldg(bn_mean + c0), m1 = __ldg(bn_mean + c1);
    const float r0 = rsqrtf(__ldg(bn_var + c0) + eps), r1 = rsqrtf(__ldg(bn_var + c1) + eps);
#pragma unroll
    for (int t = 0; t < CM_TT; ++t) {
      if (t0 + t >= T) continue;
      const float y0 = (acc[t][0] - m0) * r0 * g0 + be0, y1 = (acc[t][1] - m1) * r1 * g1 + be1;
      reinterpret_cast<uint32_t*>(out + ((long long)b * T + t0 + t) * C)[cp] = pack_bf16x2(silu_f(y0), silu_f(y1));
    }
    return;
  }
#pragma unroll
  for (int t = 0; t < CM_TT; ++t) {
    float s = ok ? acc[t][0] + acc[t][1] : 0.f;
    float q = ok ? acc[t][0] * acc[t][0] + acc[t][1] * acc[t][1] : 0.f;
    s = warp_sum(s);
    q = warp_sum(q);
    if (lane == 0) {
      s_part[warp][t][0] = s;
      s_part[warp][t][1] = q;
    }
    if (ok && t0 + t < T)
      reinterpret_cast<float2*>(conv_out + ((long long)b * T + t0 + t) * C)[cp] = make_float2(acc[t][0], acc[t][1]);
  }
  __syncthreads();
  if (threadIdx.x < 2 * CM_TT) {
    const int t = threadIdx.x >> 1, which = threadIdx.x & 1;
    if (t0 + t < T) {
      const float v = s_part[0][t][which] + s_part[1][t][which] + s_part[2][t][which] + s_part[3][t][which];
      atomicAdd(stats + ((long long)b * T + t0 + t) * 2 + which, v);
    }
  }
}

// y = SiLU(LN(conv_out)) with mean / variance from the accumulated (sum, sum of squares): one warp per frame
template <int NV>
__global__ void __launch_bounds__(256)
conv_norm_silu_kernel(const float* __restrict__ conv_out, const float* __restrict__ stats,
                      const float* __restrict__ gamma, const float* __restrict__ beta, float eps, long long M, int C,
                      bf16* __restrict__ out) {
  const long long row = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= M) return;
  const float mean = stats[row * 2] / (float)C;
  const float var = fmaxf(stats[row * 2 + 1] / (float)C - mean * mean, 0.f);
  const float rstd = rsqrtf(var + eps);
  const int nvec = C >> 2;
  const float4* xr = reinterpret_cast<const float4*>(conv_out + row * C);
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int idx = lane + 32 * i;
    if (idx < nvec) {
      const float4 v = xr[idx];
      const float4 g = __ldg(reinterpret_cast<const float4*>(gamma) + idx);
      const float4 bb = __ldg(reinterpret_cast<const float4*>(beta) + idx);
      uint2 u;
      u.x = pack_bf16x2(silu_f((v.x - mean) * rstd * g.x + bb.x), silu_f((v.y - mean) * rstd * g.y + bb.y));
      u.y = pack_bf16x2(silu_f((v.z - mean) * rstd * g.z + bb.z), silu_f((v.w - mean) * rstd * g.w + bb.w));
      reinterpret_cast<uint2*>(out + row * C)[idx] = u;
    }
  }
}

template <int K>
static int launch_conv_split(const bf16* x, const float* pw1_bias, const float* dw_w, const float* dw_b,
                             const float* norm_w, const float* norm_b, const float* bn_mean, const float* bn_var,
                             int use_ln, float eps, bf16* out, float* conv_tmp, float* stats, int B, int T, int C,
                             int causal, cudaStream_t stream) {
  const int C2 = C / 2;
  if (use_ln) RVB_CHECK_CUDA(cudaMemsetAsync(stats, 0, (size_t)B * T * 2 * sizeof(float), stream));
  dim3 grid((T + CM_TT - 1) / CM_TT, B, (C2 + 127) / 128);
  conv_dw_kernel<K><<<grid, 128, 0, stream>>>(x, pw1_bias, dw_w, dw_b, norm_w, norm_b, bn_mean, bn_var, use_ln, eps,
                                              conv_tmp, stats, out, T, C, causal);
  RVB_COUNT_LAUNCH();
  RVB_CHECK_LAUNCH();
  if (use_ln) {
    const long long M = (long long)B * T;
    const int nv = (C / 4 + 31) / 32;
    const unsigned g2 = (unsigned)((M + 7) / 8);
    if (nv <= 1) conv_norm_silu_kernel<1><<<g2, 256, 0, stream>>>(conv_tmp, stats, norm_w, norm_b, eps, M, C, out);
    else if (nv <= 2) conv_norm_silu_kernel<2><<<g2, 256, 0, stream>>>(conv_tmp, stats, norm_w, norm_b, eps, M, C, out);
    else if (nv <= 4) conv_norm_silu_kernel<4><<<g2, 256, 0, stream>>>(conv_tmp, stats, norm_w, norm_b, eps, M, C, out);
    else if (nv <= 8) conv_norm_silu_kernel<8><<<g2, 256, 0, stream>>>(conv_tmp, stats, norm_w, norm_b, eps, M, C, out);
    else if (nv <= 16) conv_norm_silu_kernel<16><<<g2, 256, 0, stream>>>(conv_tmp, stats, norm_w, norm_b, eps, M, C, out);
    else conv_norm_silu_kernel<32><<<g2, 256, 0, stream>>>(conv_tmp, stats, norm_w, norm_b, eps, M, C, out);
    RVB_COUNT_LAUNCH();
    RVB_CHECK_LAUNCH();
  }
  return 0;
}

// Generic fallback (any K <= 64, any even C): conv results staged in shared memory.
__global__ void __launch_bounds__(256)
conv_mid_kernel(const bf16* __restrict__ x, const float* __restrict__ pw1_bias, const float* __restrict__ dw_w,
                const float* __restrict__ dw_b, const float* __restrict__ norm_w, const float* __restrict__ norm_b, const float* __restrict__ bn_mean,
                const float* __restrict__ bn_var, int use_ln, float eps, bf16* __restrict__ out, int T, int C, int K,
                int causal) {
  extern __shared__ __align__(16) uint8_t smem_cm[];
  const int rows = CM_TT + K - 1;
  bf16* s_glu = reinterpret_cast<bf16*>(smem_cm);                                   // [rows][C]
  float* s_conv = reinterpret_cast<float*>(smem_cm + (size_t)rows * C * sizeof(bf16));  // [CM_TT][C]
  const int b = blockIdx.y;
  const int t0 = blockIdx.x * CM_TT;
  const int left = causal ? (K - 1) : (K - 1) / 2;
  const int C2 = C >> 1;
  for (int i = threadIdx.x; i < rows * C2; i += blockDim.x) {
    int r = i / C2, cp = i - r * C2;
    int t = t0 - left + r;
    uint32_t o = 0u;
    if (t >= 0 && t < T) {
      const bf16* xr = x + ((long long)b * T + t) * (2 * C);
      float2 a = unpack_bf16x2(reinterpret_cast<const uint32_t*>(xr)[cp]);
      float2 g = unpack_bf16x2(reinterpret_cast<const uint32_t*>(xr + C)[cp]);
      o = pack_bf16x2(a.x * sigmoid_f(g.x), a.y * sigmoid_f(g.y));
    } else if (t < 0 && causal) {
      float2 a = unpack_bf16x2(pack_bf16x2(__ldg(pw1_bias + 2 * cp), __ldg(pw1_bias + 2 * cp + 1)));
      float2 g = unpack_bf16x2(pack_bf16x2(__ldg(pw1_bias + C + 2 * cp), __ldg(pw1_bias + C + 2 * cp + 1)));
      o = pack_bf16x2(a.x * sigmoid_f(g.x), a.y * sigmoid_f(g.y));
    }
    reinterpret_cast<uint32_t*>(s_glu)[i] = o;
  }
  __syncthreads();
  for (int cp = threadIdx.x; cp < C2; cp += blockDim.x) {
    float acc0[CM_TT], acc1[CM_TT];
    const float b0 = __ldg(dw_b + 2 * cp), b1 = __ldg(dw_b + 2 * cp + 1);
#pragma unroll
    for (int t = 0; t < CM_TT; ++t) {
      acc0[t] = b0;
      acc1[t] = b1;
    }
    for (int r = 0; r < rows; ++r) {
      float2 v = unpack_bf16x2(reinterpret_cast<const uint32_t*>(s_glu)[r * C2 + cp]);
#pragma unroll
      for (int t = 0; t < CM_TT; ++t) {
        int k = r - t;
        if (k >= 0 && k < K) {
          acc0[t] = fmaf(__ldg(dw_w + (2 * cp) * K + k), v.x, acc0[t]);
          acc1[t] = fmaf(__ldg(dw_w + (2 * cp + 1) * K + k), v.y, acc1[t]);
        }
      }
    }
#pragma unroll
    for (int t = 0; t < CM_TT; ++t)
      reinterpret_cast<float2*>(s_conv + (size_t)t * C)[cp] = make_float2(acc0[t], acc1[t]);
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  for (int t = warp; t < CM_TT; t += nwarps) {
    if (t0 + t >= T) continue;
    const float* row = s_conv + (size_t)t * C;
    bf16* orow = out + ((long long)b * T + t0 + t) * C;
    float mean = 0.f, rstd = 1.f;
    if (use_ln) {
      float s = 0.f;
      for (int c = lane; c < C; c += 32) s += row[c];
      mean = warp_sum(s) / (float)C;
      float q = 0.f;
      for (int c = lane; c < C; c += 32) {
        float dlt = row[c] - mean;
        q += dlt * dlt;
      }
      rstd = rsqrtf(warp_sum(q) / (float)C + eps);
    }
    for (int cp = lane; cp < C2; cp += 32) {
      float2 v = reinterpret_cast<const float2*>(row)[cp];
      const int c0 = 2 * cp, c1 = 2 * cp + 1;
      float y0, y1;
      if (use_ln) {
        y0 = (v.x - mean) * rstd * __ldg(norm_w + c0) + __ldg(norm_b + c0);
        y1 = (v.y - mean) * rstd * __ldg(norm_w + c1) + __ldg(norm_b + c1);
      } else {
        y0 = (v.x - __ldg(bn_mean + c0)) * rsqrtf(__ldg(bn_var + c0) + eps) * __ldg(norm_w + c0) + __ldg(norm_b + c0);
        y1 = (v.y - __ldg(bn_mean + c1)) * rsqrtf(__ldg(bn_var + c1) + eps) * __ldg(norm_w + c1) + __ldg(norm_b + c1);
      }
      reinterpret_cast<uint32_t*>(orow)[cp] = pack_bf16x2(silu_f(y0), silu_f(y1));
    }
  }
}

template <int K, int NIT, int NT>
static int launch_conv_mid_fast(const bf16* x, const float* pw1_bias, const float* dw_w, const float* dw_b, const float* norm_w,
                                const float* norm_b, const float* bn_mean, const float* bn_var, int use_ln, float eps,
                                bf16* out, int B, int T, int C, int causal, cudaStream_t stream) {
  const size_t smem = (size_t)(CM_TT + K - 1) * C * sizeof(bf16);
  static size_t configured = 0;
  if (smem > 48 * 1024 && smem > configured) {
    RVB_CHECK_CUDA(cudaFuncSetAttribute(conv_mid_fast_kernel<K, NIT, NT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        (int)smem));
    configured = smem;
  }
  dim3 grid((T + CM_TT - 1) / CM_TT, B);
  conv_mid_fast_kernel<K, NIT, NT><<<grid, NT, smem, stream>>>(x, pw1_bias, dw_w, dw_b, norm_w, norm_b, bn_mean, bn_var, use_ln,
                                                           eps, out, T, C, causal);
  RVB_COUNT_LAUNCH();
  RVB_CHECK_LAUNCH();
  return 0;
}

int launch_conv_mid(const bf16* x, const float* pw1_bias, const float* dw_w, const float* dw_b, const float* norm_w, const float* norm_b,
                    const float* bn_mean, const float* bn_var, int use_layer_norm, float eps, bf16* out, int B, int T,
                    int C, int K, int causal, cudaStream_t stream, float* conv_tmp, float* stats) {
  RVB_REQUIRE(C % 2 == 0 && K >= 1 && K <= 64, "conv_mid: unsupported C=%d K=%d", C, K);
  RVB_REQUIRE(!causal || pw1_bias != nullptr, "conv_mid: causal mode needs the pointwise_conv1 bias");
  if (conv_tmp != nullptr && stats != nullptr && C % 4 == 0 && C <= 4096) {
#define RVB_CS(KK)                                                                                                   \
  return launch_conv_split<KK>(x, pw1_bias, dw_w, dw_b, norm_w, norm_b, bn_mean, bn_var, use_layer_norm, eps, out,   \
                               conv_tmp, stats, B, T, C, causal, stream)
    if (K == 15) RVB_CS(15);
    if (K == 31) RVB_CS(31);
    if (K == 7) RVB_CS(7);
#undef RVB_CS
  }
  // one channel pair per thread when it fits a 512-thread CTA (C <= 1024), else two per thread (C <= 2048)
  const int C2 = C / 2;
#define RVB_CM(KK, NN, TT)                                                                                        \
  return launch_conv_mid_fast<KK, NN, TT>(x, pw1_bias, dw_w, dw_b, norm_w, norm_b, bn_mean, bn_var, use_layer_norm, eps, out, \
                                          B, T, C, causal, stream)
  if (K == 15 || K == 31 || K == 7) {
    if (C2 <= 128) { if (K == 15) RVB_CM(15, 1, 128); if (K == 31) RVB_CM(31, 1, 128); RVB_CM(7, 1, 128); }
    if (C2 <= 256) { if (K == 15) RVB_CM(15, 1, 256); if (K == 31) RVB_CM(31, 1, 256); RVB_CM(7, 1, 256); }
    if (C2 <= 512) { if (K == 15) RVB_CM(15, 1, 512); if (K == 31) RVB_CM(31, 1, 512); RVB_CM(7, 1, 512); }
    if (C2 <= 1024) { if (K == 15) RVB_CM(15, 2, 512); if (K == 31) RVB_CM(31, 2, 512); RVB_CM(7, 2, 512); }
  }
#undef RVB_CM
  const size_t smem = (size_t)(CM_TT + K - 1) * C * sizeof(bf16) + (size_t)CM_TT * C * sizeof(float);
  RVB_REQUIRE(smem <= 200 * 1024, "conv_mid: C=%d K=%d needs %zu B of shared memory", C, K, smem);
  static size_t configured = 0;
  if (smem > 48 * 1024 && smem > configured) {
    RVB_CHECK_CUDA(cudaFuncSetAttribute(conv_mid_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = smem;
  }
  dim3 grid((T + CM_TT - 1) / CM_TT, B);
  conv_mid_kernel<<<grid, 256, smem, stream>>>(x, pw1_bias, dw_w, dw_b, norm_w, norm_b, bn_mean, bn_var, use_layer_norm, eps,
                                               out, T, C, K, causal);
  RVB_COUNT_LAUNCH();
  RVB_CHECK_LAUNCH();
  return 0;
}

// ---------------------------------------------------------------------------------------------------------------
__global__ void scale_cast_kernel(const float* __restrict__ x, float scale, float* __restrict__ of,
                                  bf16* __restrict__ ob, long long n) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (; i < n; i += stride) {
    float v = x[i] * scale;
    if (of) of[i] = v;
    if (ob) ob[i] = __float2bfloat16(v);
  }
}

int launch_scale_cast(const float* x, float scale, float* out_f32, bf16* out_bf16, long long n, cudaStream_t stream) {
  if (n <= 0) return 0;
  long long blocks = (n + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  scale_cast_kernel<<<(int)blocks, 256, 0, stream>>>(x, scale, out_f32, out_bf16, n);
  RVB_COUNT_LAUNCH();
  RVB_CHECK_LAUNCH();
  return 0;
}

int launch_f32_to_bf16(const float* x, bf16* out, long long n, cudaStream_t stream) {
  return launch_scale_cast(x, 1.0f, nullptr, out, n, stream);
}

struct WSumPtrs {
  const float* p[8];
  float c[8];
};
__global__ void weighted_sum_kernel(WSumPtrs in, int n_in, long long n, bf16* __restrict__ ob, float* __restrict__ of) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (; i < n; i += stride) {
    // same association order as the reference: y = c0*L0(x); y = y + c1*L1(x); ...  (encoder_layer.py:378-390)
    float v = in.c[0] * in.p[0][i];
    for (int j = 1; j < n_in; ++j) v = v + in.c[j] * in.p[j][i];
    if (ob) ob[i] = __float2bfloat16(v);
    if (of) of[i] = v;
  }
}

int launch_weighted_sum_bf16(const float* const* ins, const float* coef, int n_in, long long n, bf16* out_bf16,
                             float* out_f32, cudaStream_t stream) {
  RVB_REQUIRE(n_in >= 1 && n_in <= 8, "weighted_sum: n_in=%d unsupported", n_in);
  WSumPtrs w;
  for (int i = 0; i < n_in; ++i) {
    w.p[i] = ins[i];
    w.c[i] = coef[i];
  }
  long long blocks = (n + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  weighted_sum_kernel<<<(int)blocks, 256, 0, stream>>>(w, n_in, n, out_bf16, out_f32);
  RVB_COUNT_LAUNCH();
  RVB_CHECK_LAUNCH();
  return 0;
}

// sinusoidal positional table (transformer/embedding.py:39-56): pe[pos, 2i] = sin(pos * w_i), pe[pos, 2i+1] = cos(..)
__global__ void sinusoid_kernel(int T, int d, float nlod, float* __restrict__ of, bf16* __restrict__ ob) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)T * d) return;
  int pos = (int)(i / d), c = (int)(i - (long long)pos * d);
  int i2 = c & ~1;
  // torch: div_term = exp(arange(0, d, 2) * -(ln(10000)/d)) in fp32, then sin/cos(position * div_term) in fp32
  float div = expf((float)i2 * nlod);
  float ang = (float)pos * div;
  float v = (c & 1) ? cosf(ang) : sinf(ang);
  if (of) of[i] = v;
  if (ob) ob[i] = __float2bfloat16(v);
}

int launch_sinusoid(int T, int d, float* out_f32, bf16* out_bf16, cudaStream_t stream) {
  long long n = (long long)T * d;
  if (n <= 0) return 0;
  sinusoid_kernel<<<(int)((n + 255) / 256), 256, 0, stream>>>(T, d, (float)(-(log(10000.0) / (double)d)), out_f32,
                                                              out_bf16);
  RVB_COUNT_LAUNCH();
  RVB_CHECK_LAUNCH();
  return 0;
}

// decoder embedding + absolute positional encoding (transformer/decoder.py:147 `self.embed`,
// embedding.py:58-76: x * sqrt(d) + pe[:L])
__global__ void embed_posenc_kernel(const int* __restrict__ tok, const float* __restrict__ emb, int L, int d,
                                    float nlod, float* __restrict__ out) {
  const int r = blockIdx.x;
  const int pos = r % L;
  const int id = tok[r];
  const float xs = sqrtf((float)d);
  for (int c = threadIdx.x; c < d; c += blockDim.x) {
    int i2 = c & ~1;
    float div = expf((float)i2 * nlod);
    float ang = (float)pos * div;
    float pe = (c & 1) ? cosf(ang) : sinf(ang);
    out[(long long)r * d + c] = emb[(long long)id * d + c] * xs + pe;
  }
}

int launch_embed_posenc(const int* tokens, const float* emb, int N, int L, int d, float* out, cudaStream_t stream) {
  if (N * L <= 0) return 0;
  embed_posenc_kernel<<<N * L, 128, 0, stream>>>(tokens, emb, L, d, (float)(-(log(10000.0) / (double)d)), out);
  RVB_COUNT_LAUNCH();
  RVB_CHECK_LAUNCH();
  return 0;
}

}  // namespace rvb
