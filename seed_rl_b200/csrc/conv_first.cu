// First layer of ImpalaDeep (dmlab/networks.py:31-37: Conv2D(16, 3, 'same') on the uint8 frames,
// then MaxPool 3x3 / 2 'same') -- backward, fused.
//
// The gradient that reaches the convolution output is the max-pool's scatter of the pooled gradient
// g: at most one position per (pooled pixel, channel) is non-zero.  Materialising that full-
// resolution tensor (607 MB fp32 at 1 344 frames) and running a dense weight-gradient convolution over
// it was 0.69 ms of the 5.5 ms step (pool backward 0.30 + weight gradient 0.39).  Here each
// (pooled pixel q, channel co) adds  g[q][co] * x[argmax(q, co) + tap]  to dW[tap][:][co] directly:
//     dW[kh][kw][ci][co] = sum_{n,q} g[n,q,co] * x[n][p(q,co) + (kh-1, kw-1)][ci] / 255,
//     db[co]            = sum_{n,q} g[n,q,co],
// p(q, co) = the window position stored by the forward pool (idx).  4x fewer MACs than the dense
// form, no full-resolution gradient, fp32 accumulation (exact products: frames are integers).
// A CTA stages one frame at a time in shared memory as bf16 (exact for 0..255) with a zero border;
// thread = (channel co, pooled-pixel lane): 9 x 8-byte patch loads + 36 FMAs per (q, co) into 36
// register accumulators; per-CTA partials go through the deterministic deferred reduce.
#include "kernels.h"
#include "tc_common.cuh"

namespace seedrl {

constexpr int kFwThreads = 256;
constexpr int kFwCo = 16;
constexpr int kFwBands = 4;

struct FirstWgradArgs {
  int N, H, W, Ho, Wo, pt, pl;
  int Lpp, PWp, RHp;             // pooled plane-tensor geometry
  const uint8_t* frames;         // [N,H,W,4]
  const uint4* g;                // pooled gradient planes: 2 hi planes then 2 lo planes, [Lpp] x 16 B
  const uint8_t* idx;            // [N,Ho,Wo,16] window tap kh*3+kw of the forward arg-max
  float* partial;                // [grid][9*4*16 + 16]
};

__global__ void __launch_bounds__(kFwThreads, 3) first_wgrad_pooled_kernel(const FirstWgradArgs a) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  uint2* s_x = reinterpret_cast<uint2*>(smem_raw);               // [(H+2)][(W+2)] pixels x 4 bf16
  const int tid = threadIdx.x;
  const int co = tid & (kFwCo - 1), ql = tid >> 4;               // 16 pooled-pixel lanes
  const int SW = a.W + 2, SH = a.H + 2;
  const int nq = a.Ho * a.Wo;
  float acc[36];
#pragma unroll
  for (int i = 0; i < 36; ++i) acc[i] = 0.f;
  float accb = 0.f;
  // zero border once (the interior is rewritten per frame)
  for (int i = tid; i < SH * SW; i += kFwThreads) s_x[i] = make_uint2(0u, 0u);
  __syncthreads();
  const unsigned short* gh = reinterpret_cast<const unsigned short*>(a.g);
  const size_t lo_off = (size_t)2 * a.Lpp * 8;                   // in bf16 elements: 2 hi planes
  const size_t plane_off = (size_t)(co >> 3) * a.Lpp * 8 + (co & 7);
  // work unit = (frame, band of pooled rows): 4 bands per frame keep the static schedule balanced
  // (1 344 frames over 444 CTA slots would leave a 4-vs-3 frame imbalance)
  const int RB = (a.Ho + kFwBands - 1) / kFwBands;
  const int units = a.N * kFwBands;
  for (int u = blockIdx.x; u < units; u += gridDim.x) {
    const int n = u / kFwBands, band = u - n * kFwBands;
    const int r0 = band * RB, r1 = min(a.Ho, r0 + RB);
    if (r0 >= r1) continue;
    // ---- stage the frame rows this band's patches can touch: uint8 x4 -> 4 halves ----------------
    const int fs = max(0, 2 * r0 - a.pt - 1), fe = min(a.H, 2 * (r1 - 1) - a.pt + 4);
    const uchar4* src = reinterpret_cast<const uchar4*>(a.frames) + (size_t)n * a.H * a.W;
    for (int i = fs * a.W + tid; i < fe * a.W; i += kFwThreads) {
      const int h = i / a.W, w = i - h * a.W;
      const uchar4 u4 = __ldg(src + i);
      // bf16 pairs (exact for 0..255): the consumer turns them back into floats with one shift /
      // one mask each (fp16 needed F2F conversions: the conversion pipe, not the FMAs, bounded the loop)
      const uint32_t f0 = __float_as_uint((float)u4.x) >> 16, f1 = __float_as_uint((float)u4.y) & 0xFFFF0000u;
      const uint32_t f2 = __float_as_uint((float)u4.z) >> 16, f3 = __float_as_uint((float)u4.w) & 0xFFFF0000u;
      s_x[(h + 1) * SW + w + 1] = make_uint2(f0 | f1, f2 | f3);
    }
    __syncthreads();
    const uint8_t* idx_n = a.idx + (size_t)n * nq * kFwCo;
    // The arg-max tap and the gradient of pooled pixel q + 16 are fetched (L2 latency) while the 36
    // FMAs of pixel q run: without this the loop is a chain of dependent global loads.
    auto fetch = [&](int q, int* t, uint32_t* ghi, uint32_t* glo, int* qh_, int* qw_) {
      const int qh = q / a.Wo, qw = q - qh * a.Wo;
      const size_t sp = (size_t)(n * a.RHp + qh + 1) * a.PWp + qw + 1;
      *t = idx_n[(size_t)q * kFwCo + co];
      *ghi = gh[plane_off + sp * 8];
      *glo = gh[lo_off + plane_off + sp * 8];
      *qh_ = qh; *qw_ = qw;
    };
    const int qend = r1 * a.Wo;
    int q = r0 * a.Wo + ql;
    int t_n = 0, qh_n = 0, qw_n = 0;
    uint32_t ghi_n = 0, glo_n = 0;
    if (q < qend) fetch(q, &t_n, &ghi_n, &glo_n, &qh_n, &qw_n);
    while (q < qend) {
      const int t = t_n, qh = qh_n, qw = qw_n;
      const float gv = __uint_as_float(ghi_n << 16) + __uint_as_float(glo_n << 16);
      q += kFwThreads / kFwCo;
      if (q < qend) fetch(q, &t_n, &ghi_n, &glo_n, &qh_n, &qw_n);
      const int kh = t / 3, kw = t - kh * 3;
      // arg-max position in frame coordinates; its 3x3 patch starts at smem (ph, pw)
      const int ph = qh * 2 - a.pt + kh, pw = qw * 2 - a.pl + kw;
      accb += gv;
      const uint2* patch = s_x + ph * SW + pw;
      uint2 v[9];
#pragma unroll
      for (int dh = 0; dh < 3; ++dh)
#pragma unroll
        for (int dw = 0; dw < 3; ++dw) v[dh * 3 + dw] = patch[dh * SW + dw];
#pragma unroll
      for (int k = 0; k < 9; ++k) {
        float* c = acc + k * 4;
        c[0] = fmaf(gv, __uint_as_float(v[k].x << 16), c[0]);
        c[1] = fmaf(gv, __uint_as_float(v[k].x & 0xFFFF0000u), c[1]);
        c[2] = fmaf(gv, __uint_as_float(v[k].y << 16), c[2]);
        c[3] = fmaf(gv, __uint_as_float(v[k].y & 0xFFFF0000u), c[3]);
      }
    }
    __syncthreads();                       // the staged rows are rewritten by the next unit
  }
  // ---- reduce the 16 pooled-pixel lanes per channel (fixed order) -> this CTA's partial ----------
  float* s_red = reinterpret_cast<float*>(smem_raw);              // [16 ql][37][16 co] (frame buffer is free)
#pragma unroll
  for (int i = 0; i < 36; ++i) s_red[(ql * 37 + i) * kFwCo + co] = acc[i];
  s_red[(ql * 37 + 36) * kFwCo + co] = accb;
  __syncthreads();
  float* dst = a.partial + (size_t)blockIdx.x * (36 * kFwCo + kFwCo);
  for (int e = tid; e < 37 * kFwCo; e += kFwThreads) {
    const int i = e / kFwCo, c = e - i * kFwCo;
    float s = 0.f;
#pragma unroll
    for (int l = 0; l < kFwThreads / kFwCo; ++l) s += s_red[(l * 37 + i) * kFwCo + c];
    if (i < 36) dst[i * kFwCo + c] = s * (1.0f / 255.0f);         // (tap, ci) x co: HWIO order
    else dst[36 * kFwCo + c] = s;
  }
}

static void same_pad3s2_(int in, int* out, int* before) {
  *out = (in + 1) / 2;
  const int total = (*out - 1) * 2 + 3 - in;
  *before = total > 0 ? total / 2 : 0;
}

bool first_wgrad_pooled_supported(int cin, int cout, int H, int W) {
  return cin == 4 && cout == 16 && (size_t)(H + 2) * (W + 2) * 8 <= 72 * 1024;
}

// dW / db of the first convolution from the POOLED gradient planes + the pool's arg-max taps.
int first_wgrad_pooled(int N, int H, int W, const uint8_t* frames, const void* g_planes, const uint8_t* idx,
                       float* dw, float* db, WgradBatch* batch, cudaStream_t st) {
  int Ho, Wo, pt, pl;
  same_pad3s2_(H, &Ho, &pt);
  same_pad3s2_(W, &Wo, &pl);
  FirstWgradArgs a;
  a.N = N; a.H = H; a.W = W; a.Ho = Ho; a.Wo = Wo; a.pt = pt; a.pl = pl;
  a.Lpp = (int)planes_positions(N, Ho, Wo); a.PWp = Wo + 2; a.RHp = Ho + 1;
  a.frames = frames; a.g = reinterpret_cast<const uint4*>(g_planes); a.idx = idx;
  size_t smem = (size_t)(H + 2) * (W + 2) * 8;
  const size_t red = (size_t)(kFwThreads / kFwCo) * 37 * kFwCo * 4;
  if (red > smem) smem = red;
  const int NW = 36 * kFwCo + kFwCo;
  int grid = 3 * kNumSMs;
  if (grid > N * kFwBands) grid = N * kFwBands;
  if (!batch || batch->n >= kMaxReduceJobs || batch->used + (size_t)grid * NW > batch->cap_floats)
    return set_error(SEEDRL_ERR_INVALID_ARGUMENT, "first_wgrad_pooled: partial buffer too small");
  a.partial = batch->buf + batch->used;
  batch->used += (size_t)grid * NW;
  static bool attr = false;
  if (!attr) {
    SEEDRL_CUDA(cudaFuncSetAttribute(first_wgrad_pooled_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     72 * 1024));
    attr = true;
  }
  first_wgrad_pooled_kernel<<<grid, kFwThreads, smem, st>>>(a);
  count_launch(PC_CONV_WGRAD, st);
  SEEDRL_CHECK_LAUNCH();
  batch->jobs[batch->n++] = ReduceJob{a.partial, dw, db, grid, 36 * kFwCo, kFwCo};
  return SEEDRL_OK;
}

}  // namespace seedrl
