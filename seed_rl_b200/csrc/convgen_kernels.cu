// Generic 'valid' strided convolution (any square kernel / stride; cin, cout multiples
// of 4) for the IMPALA-paper shallow net (conv 8x8/4 -> 16, conv 4x4/2 -> 32), which is
// not in the reference.  Simple direct kernels: these layers are ~6% of the
// deep net's FLOPs, so they are kept straightforward (coalesced float4 over channels,
// weights through the read-only path).
#include "kernels.h"

namespace seedrl {

// one thread = one output pixel x 4 output channels
template <bool U8>
__global__ void __launch_bounds__(256)
convgen_fwd_kernel(int N, int H, int W, int CIN, int COUT, int K, int S, int Ho, int Wo,
                   const void* __restrict__ in_, const float* __restrict__ w,
                   const float* __restrict__ bias, int relu, float* __restrict__ out) {
  const int O4 = COUT >> 2;
  const long long total = (long long)N * Ho * Wo * O4;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int c4 = (int)(i % O4);
  long long r = i / O4;
  const int wo = (int)(r % Wo); r /= Wo;
  const int ho = (int)(r % Ho);
  const long long n = r / Ho;
  float4 acc = __ldg(reinterpret_cast<const float4*>(bias) + c4);
  for (int kh = 0; kh < K; ++kh) {
    const int h = ho * S + kh;
    for (int kw = 0; kw < K; ++kw) {
      const int x = wo * S + kw;
      const size_t pix = ((size_t)n * H + h) * W + x;
      const float4* wp = reinterpret_cast<const float4*>(w + ((size_t)(kh * K + kw) * CIN) * COUT) + c4;
      for (int ci = 0; ci < CIN; ci += 4) {
        float xv[4];
        if (U8) {
          const uchar4 u = __ldg(reinterpret_cast<const uchar4*>(in_) + (pix * CIN + ci) / 4);
          const float k = 1.0f / 255.0f;
          xv[0] = u.x * k; xv[1] = u.y * k; xv[2] = u.z * k; xv[3] = u.w * k;
        } else {
          const float4 f = __ldg(reinterpret_cast<const float4*>(in_) + (pix * CIN + ci) / 4);
          xv[0] = f.x; xv[1] = f.y; xv[2] = f.z; xv[3] = f.w;
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float4 ww = __ldg(wp + (size_t)(ci + q) * O4);
          acc.x = fmaf(xv[q], ww.x, acc.x); acc.y = fmaf(xv[q], ww.y, acc.y);
          acc.z = fmaf(xv[q], ww.z, acc.z); acc.w = fmaf(xv[q], ww.w, acc.w);
        }
      }
    }
  }
  if (relu) {
    acc.x = fmaxf(acc.x, 0.f); acc.y = fmaxf(acc.y, 0.f);
    acc.z = fmaxf(acc.z, 0.f); acc.w = fmaxf(acc.w, 0.f);
  }
  reinterpret_cast<float4*>(out)[i] = acc;
}

// one thread = one input pixel x one input channel; dx masked by mask > 0 (relu of the
// producing layer).
__global__ void __launch_bounds__(256)
convgen_dgrad_kernel(int N, int H, int W, int CIN, int COUT, int K, int S, int Ho, int Wo,
                     const float* __restrict__ dy, const float* __restrict__ w,
                     const float* __restrict__ mask, float* __restrict__ dx) {
  const long long total = (long long)N * H * W * CIN;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int ci = (int)(i % CIN);
  long long r = i / CIN;
  const int x = (int)(r % W); r /= W;
  const int h = (int)(r % H);
  const long long n = r / H;
  float acc = 0.f;
  if (!mask || mask[i] > 0.f) {
    for (int kh = 0; kh < K; ++kh) {
      const int hh = h - kh;
      if (hh < 0 || hh % S) continue;
      const int ho = hh / S;
      if (ho >= Ho) continue;
      for (int kw = 0; kw < K; ++kw) {
        const int xx = x - kw;
        if (xx < 0 || xx % S) continue;
        const int wo = xx / S;
        if (wo >= Wo) continue;
        const float4* d4 = reinterpret_cast<const float4*>(dy + (((size_t)n * Ho + ho) * Wo + wo) * COUT);
        const float4* w4 = reinterpret_cast<const float4*>(w + ((size_t)(kh * K + kw) * CIN + ci) * COUT);
        for (int c = 0; c < COUT / 4; ++c) {
          const float4 d = __ldg(d4 + c), ww = __ldg(w4 + c);
          acc += d.x * ww.x + d.y * ww.y + d.z * ww.z + d.w * ww.w;
        }
      }
    }
  }
  dx[i] = acc;
}

// CTA b handles images n = b, b+grid, ...; thread owns weights e = tid + 256*j (j < EPT).
template <bool U8, int EPT>
__global__ void __launch_bounds__(256)
convgen_wgrad_kernel(int N, int H, int W, int CIN, int COUT, int K, int S, int Ho, int Wo,
                     const void* __restrict__ x_, const float* __restrict__ dy,
                     float* __restrict__ partial) {
  const int NWt = K * K * CIN * COUT;
  const int tid = threadIdx.x;
  float acc[EPT];
  int xoff[EPT], co[EPT];
#pragma unroll
  for (int j = 0; j < EPT; ++j) {
    acc[j] = 0.f;
    const int e = tid + 256 * j;
    const int c = e % COUT;
    const int ci = (e / COUT) % CIN;
    const int kw = (e / (COUT * CIN)) % K;
    const int kh = e / (COUT * CIN * K);
    co[j] = c;
    xoff[j] = (kh * W + kw) * CIN + ci;
  }
  float bacc = 0.f;
  for (int n = blockIdx.x; n < N; n += gridDim.x) {
    for (int ho = 0; ho < Ho; ++ho)
      for (int wo = 0; wo < Wo; ++wo) {
        const float* d = dy + (((size_t)n * Ho + ho) * Wo + wo) * COUT;
        const size_t xb = (((size_t)n * H + ho * S) * W + wo * S) * CIN;
#pragma unroll
        for (int j = 0; j < EPT; ++j) {
          if (tid + 256 * j < NWt) {
            float xv;
            if (U8) xv = (float)__ldg(reinterpret_cast<const uint8_t*>(x_) + xb + xoff[j]) * (1.0f / 255.0f);
            else xv = __ldg(reinterpret_cast<const float*>(x_) + xb + xoff[j]);
            acc[j] = fmaf(xv, __ldg(d + co[j]), acc[j]);
          }
        }
        if (tid < COUT) bacc += __ldg(d + tid);
      }
  }
  float* dst = partial + (size_t)blockIdx.x * (NWt + COUT);
#pragma unroll
  for (int j = 0; j < EPT; ++j)
    if (tid + 256 * j < NWt) dst[tid + 256 * j] = acc[j];
  if (tid < COUT) dst[NWt + tid] = bacc;
}

int convgen_forward(int N, int H, int W, int cin, int cout, int k, int stride, int in_u8,
                    const void* in, const float* w, const float* bias, int relu, float* out,
                    cudaStream_t st) {
  if (cin % 4 || cout % 4) return set_error(SEEDRL_ERR_INVALID_ARGUMENT, "convgen: channels % 4");
  const int Ho = (H - k) / stride + 1, Wo = (W - k) / stride + 1;
  const long long total = (long long)N * Ho * Wo * (cout / 4);
  const unsigned grid = (unsigned)((total + 255) / 256);
  if (in_u8) convgen_fwd_kernel<true><<<grid, 256, 0, st>>>(N, H, W, cin, cout, k, stride, Ho, Wo, in, w, bias, relu, out);
  else convgen_fwd_kernel<false><<<grid, 256, 0, st>>>(N, H, W, cin, cout, k, stride, Ho, Wo, in, w, bias, relu, out);
  count_launch(PC_CONV_FWD, st);
  SEEDRL_CHECK_LAUNCH();
  return SEEDRL_OK;
}

int convgen_dgrad(int N, int H, int W, int cin, int cout, int k, int stride, const float* dy,
                  const float* w, const float* mask, float* dx, cudaStream_t st) {
  const int Ho = (H - k) / stride + 1, Wo = (W - k) / stride + 1;
  const long long total = (long long)N * H * W * cin;
  convgen_dgrad_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(N, H, W, cin, cout, k, stride,
                                                                        Ho, Wo, dy, w, mask, dx);
  count_launch(PC_CONV_DGRAD, st);
  SEEDRL_CHECK_LAUNCH();
  return SEEDRL_OK;
}

int convgen_wgrad(int N, int H, int W, int cin, int cout, int k, int stride, int in_u8,
                  const void* x, const float* dy, float* dw, float* db, float* partial,
                  size_t partial_bytes, cudaStream_t st) {
  const int Ho = (H - k) / stride + 1, Wo = (W - k) / stride + 1;
  const int nw = k * k * cin * cout;
  if (nw > 256 * 32) return set_error(SEEDRL_ERR_INVALID_ARGUMENT, "convgen_wgrad: too many weights");
  int grid = kNumSMs * 2;
  if (grid > N) grid = N;
  if ((size_t)grid * (nw + cout) * sizeof(float) > partial_bytes)
    return set_error(SEEDRL_ERR_INVALID_ARGUMENT, "convgen_wgrad: partial buffer too small");
  const bool small = nw <= 256 * 16;
  if (in_u8) {
    if (small) convgen_wgrad_kernel<true, 16><<<grid, 256, 0, st>>>(N, H, W, cin, cout, k, stride, Ho, Wo, x, dy, partial);
    else convgen_wgrad_kernel<true, 32><<<grid, 256, 0, st>>>(N, H, W, cin, cout, k, stride, Ho, Wo, x, dy, partial);
  } else {
    if (small) convgen_wgrad_kernel<false, 16><<<grid, 256, 0, st>>>(N, H, W, cin, cout, k, stride, Ho, Wo, x, dy, partial);
    else convgen_wgrad_kernel<false, 32><<<grid, 256, 0, st>>>(N, H, W, cin, cout, k, stride, Ho, Wo, x, dy, partial);
  }
  count_launch(PC_CONV_WGRAD, st);
  SEEDRL_CHECK_LAUNCH();
  return wgrad_reduce(grid, nw, cout, partial, dw, db, st);
}

}  // namespace seedrl
