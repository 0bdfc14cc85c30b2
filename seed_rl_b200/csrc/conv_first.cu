// First layer of ImpalaDeep (dmlab/networks.py:31-37: Conv2D(16, 3, 'same') on the uint8 frames,
// then MaxPool 3x3 / 2 'same'), fused: backward here, forward (conv0pool_kernel) at the end of the file.
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
#include <cuda.h>

#include <cstdlib>

#include "kernels.h"
#include "tc_common.cuh"

namespace seedrl {

constexpr int kFwThreads = 256;
constexpr int kFwCo = 16;
constexpr int kFwRowsPerBand = 6;   // pooled rows per work unit

struct FirstWgradArgs {
  int N, H, W, Ho, Wo, pt, pl;
  int Lpp, PWp, RHp;             // pooled plane-tensor geometry
  int rb;                        // pooled rows per work unit
  const uint8_t* frames;         // [N,H,W,4]
  const uint4* g;                // pooled gradient planes: 2 hi planes then 2 lo planes, [Lpp] x 16 B
  const uint8_t* idx;            // [N,Ho,Wo,16] window tap kh*3+kw of the forward arg-max
  float* partial;                // [grid][9*4*16 + 16]
};

__global__ void __launch_bounds__(kFwThreads, 3) first_wgrad_pooled_kernel(const FirstWgradArgs a) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  const int tid = threadIdx.x;
  const int co = tid & (kFwCo - 1), ql = tid >> 4;               // 16 pooled-pixel lanes
  const int SW = a.W + 2;
  const int RB = a.rb;                                           // pooled rows per band
  const int SR = 2 * RB + 3;                                     // staged frame rows (incl. the two border rows)
  uint2* s_x = reinterpret_cast<uint2*>(smem_raw);               // [SR][SW] pixels x 4 bf16
  float* s_g = reinterpret_cast<float*>(s_x + SR * SW);          // [RB*Wo][16] pooled gradient (hi + lo)
  uint8_t* s_t = reinterpret_cast<uint8_t*>(s_g + RB * a.Wo * kFwCo);   // [RB*Wo][16] arg-max taps
  float acc[36];
#pragma unroll
  for (int i = 0; i < 36; ++i) acc[i] = 0.f;
  float accb = 0.f;
  const int bands = (a.Ho + RB - 1) / RB;
  const int units = a.N * bands;
  // work unit = (frame, band of pooled rows): keeps the static schedule balanced and the staged
  // working set small enough for 4+ CTAs per SM
  for (int u = blockIdx.x; u < units; u += gridDim.x) {
    const int n = u / bands, band = u - n * bands;
    const int r0 = band * RB, r1 = min(a.Ho, r0 + RB);
    const int nq = (r1 - r0) * a.Wo;
    // ---- stage: frame rows f0 .. f0+SR-1 (zero outside the frame / at the two border columns) ----
    const int f0 = 2 * r0 - a.pt - 1;
    const uint32_t* src = reinterpret_cast<const uint32_t*>(a.frames) + (size_t)n * a.H * a.W;
    for (int i = tid; i < SR * SW; i += kFwThreads) {
      const int lr = i / SW, bc = i - lr * SW, fr = f0 + lr;
      uint32_t w32 = 0u;
      if (fr >= 0 && fr < a.H && bc >= 1 && bc <= a.W) w32 = __ldg(src + (size_t)fr * a.W + bc - 1);
      // bf16 x 4 (exact for 0..255): the consumer turns them back into floats with one shift / mask each
      const uint32_t f0_ = __float_as_uint(__uint_as_float(__byte_perm(w32, 0x4B000000u, 0x7540)) - 8388608.0f);
      const uint32_t f1_ = __float_as_uint(__uint_as_float(__byte_perm(w32, 0x4B000000u, 0x7541)) - 8388608.0f);
      const uint32_t f2_ = __float_as_uint(__uint_as_float(__byte_perm(w32, 0x4B000000u, 0x7542)) - 8388608.0f);
      const uint32_t f3_ = __float_as_uint(__uint_as_float(__byte_perm(w32, 0x4B000000u, 0x7543)) - 8388608.0f);
      s_x[i] = make_uint2(__byte_perm(f0_, f1_, 0x7632), __byte_perm(f2_, f3_, 0x7632));
    }
    // ---- stage: the band's pooled gradient (hi + lo planes -> fp32) and arg-max taps: coalesced
    //      16-byte loads, all independent (the main loop then touches shared memory only) ----
    for (int i = tid; i < nq * 2; i += kFwThreads) {
      const int gq = i & 1, qq = i >> 1;
      const int qr = qq / a.Wo, qw = qq - qr * a.Wo;
      const size_t sp = (size_t)(n * a.RHp + r0 + qr + 1) * a.PWp + qw + 1;
      const uint4 hh = __ldg(a.g + (size_t)gq * a.Lpp + sp), ll = __ldg(a.g + (size_t)(2 + gq) * a.Lpp + sp);
      float4* dst = reinterpret_cast<float4*>(s_g + (size_t)qq * kFwCo + gq * 8);
      dst[0] = make_float4(__uint_as_float(hh.x << 16) + __uint_as_float(ll.x << 16),
                           __uint_as_float(hh.x & 0xFFFF0000u) + __uint_as_float(ll.x & 0xFFFF0000u),
                           __uint_as_float(hh.y << 16) + __uint_as_float(ll.y << 16),
                           __uint_as_float(hh.y & 0xFFFF0000u) + __uint_as_float(ll.y & 0xFFFF0000u));
      dst[1] = make_float4(__uint_as_float(hh.z << 16) + __uint_as_float(ll.z << 16),
                           __uint_as_float(hh.z & 0xFFFF0000u) + __uint_as_float(ll.z & 0xFFFF0000u),
                           __uint_as_float(hh.w << 16) + __uint_as_float(ll.w << 16),
                           __uint_as_float(hh.w & 0xFFFF0000u) + __uint_as_float(ll.w & 0xFFFF0000u));
    }
    {
      const uint4* isrc = reinterpret_cast<const uint4*>(a.idx + ((size_t)n * a.Ho + r0) * a.Wo * kFwCo);
      for (int i = tid; i < nq; i += kFwThreads) reinterpret_cast<uint4*>(s_t)[i] = __ldg(isrc + i);
    }
    __syncthreads();
    // ---- thread = (channel co, pooled-pixel lane): 9 x 8-byte patch loads + 36 FMAs per pixel ----
    int qh = ql / a.Wo, qw = ql - qh * a.Wo;                    // band-local pooled row / column
    for (int q = ql; q < nq; q += kFwThreads / kFwCo) {
      const int t = s_t[q * kFwCo + co];
      const float gv = s_g[q * kFwCo + co];
      const int kh = t / 3, kw = t - kh * 3;
      // arg-max position (frame row 2*(r0+qh) - pt + kh); its 3x3 patch starts one row / column
      // earlier = staged row 2*qh + kh, band column 2*qw - pl + kw
      const uint2* patch = s_x + (2 * qh + kh) * SW + (2 * qw - a.pl + kw);
      accb += gv;
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
      qw += kFwThreads / kFwCo;
      while (qw >= a.Wo) { qw -= a.Wo; ++qh; }
    }
    __syncthreads();                       // the staged band is rewritten by the next unit
  }
  // ---- reduce the 16 pooled-pixel lanes per channel (fixed order) -> this CTA's partial ----------
  float* s_red = reinterpret_cast<float*>(smem_raw);              // [16 ql][37][16 co] (staging buffers are free)
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
  const int wo = (W + 1) / 2;
  return cin == 4 && cout == 16 && H >= 3 && W >= 3 &&
         (size_t)(2 * kFwRowsPerBand + 3) * (W + 2) * 8 + (size_t)kFwRowsPerBand * wo * kFwCo * 5 <= 70 * 1024;
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
  a.rb = kFwRowsPerBand < Ho ? kFwRowsPerBand : Ho;
  size_t smem = (size_t)(2 * a.rb + 3) * (W + 2) * 8 + (size_t)a.rb * Wo * kFwCo * 5;
  const size_t red = (size_t)(kFwThreads / kFwCo) * 37 * kFwCo * 4;
  if (red > smem) smem = red;
  smem = (smem + 127) / 128 * 128;
  const int NW = 36 * kFwCo + kFwCo;
  const int units = N * ((Ho + a.rb - 1) / a.rb);
  int grid = 3 * kNumSMs;
  if (grid > units) grid = units;
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

// =================================================================================================
// First layer, forward, fused: Conv2D(16, 3, 'same') on the uint8 frames + bias + MaxPool 3x3/2
// 'same' (dmlab/networks.py:31-37,47-48) -> the pooled activation as plane tensors (raw and ReLU'd)
// + the arg-max taps.  The full-resolution conv output (607 MB fp32 at 1 344 frames, written once and
// read once by a separate pool kernel: 0.53 ms of the step) never leaves the SM.
//
// Work unit = (frame, band of kCpRows pooled rows).  Per unit:
//   1. the band's frame rows -> shared memory as a PAIR array: entry e = [pixel e, pixel e+1] of the
//      zero-bordered band (row pitch SW = W + 2), 4 channels each, bf16 (exact for 0..255) = 16 bytes.
//      That array IS a K-major UMMA operand whose row p reads, for kernel row kh, the 16 bytes at
//      e = p + kh*SW (taps kw = 0, 1) and at e + 2 (taps kw = 2 and a 4th, zero-weight, tap): the two
//      K-groups of one K = 16 instruction are the same array 32 bytes apart (LBO = 32 B).  A whole
//      kernel row per MMA: 3 MMAs of 128 x 32 x 16 per 128 positions, no im2col pass.
//   2. one elected thread issues them: D[128, 0:16] = A hi(W), D[128, 16:32] = A lo(W)
//      (frames are exact in bf16, so two products make the fp32-faithful result);
//   3. epilogue: TMEM -> (hi + lo) / 255 + bias -> fp32 tile in shared memory;
//   4. pooling from shared memory, TF-SAME windows, first maximum wins; hi/lo split; coalesced
//      16-byte plane stores; padding positions of the plane tensors written as zeros.
// 2 CTAs / SM (TMEM 2 x 256 columns): the phases of one CTA overlap the other's.
// kCpRows pooled rows per unit (template parameter: 3 -> up to 6 blocks of 128 positions, TMEM 256 columns,
// 2 CTAs / SM; 2 -> up to 4 blocks, TMEM 128 columns, 3 CTAs / SM).
constexpr int kCpThreadsF = 256;
constexpr int kCpOutStride = 20;    // floats per position in the fp32 tile (16 + 4: pool reads 2-way conflict)

struct Conv0PoolArgs {
  int N, H, W, Ho, Wo, pt, pl;
  int Lpp, PWp, RHp;                 // pooled plane-tensor geometry
  unsigned int sw_mul; int sw_sh;    // division by SW = W + 2
  const uint8_t* frames;             // [N,H,W,4]
  const float* w;                    // [3,3,4,16]
  const float* bias;                 // [16]
  uint4* praw; uint4* prelu;         // plane tensors, 16 channels (2 hi planes, 2 lo planes)
  uint8_t* idx;                      // [N,Ho,Wo,16]
  int* err;
};

__device__ __forceinline__ uint2 u8x4_to_bf16x4(uint32_t w32) {
  // byte -> float without I2F: 0x4B0000kk is 2^23 + kk; the high half of the float is its bf16
  const uint32_t f0 = __float_as_uint(__uint_as_float(__byte_perm(w32, 0x4B000000u, 0x7540)) - 8388608.0f);
  const uint32_t f1 = __float_as_uint(__uint_as_float(__byte_perm(w32, 0x4B000000u, 0x7541)) - 8388608.0f);
  const uint32_t f2 = __float_as_uint(__uint_as_float(__byte_perm(w32, 0x4B000000u, 0x7542)) - 8388608.0f);
  const uint32_t f3 = __float_as_uint(__uint_as_float(__byte_perm(w32, 0x4B000000u, 0x7543)) - 8388608.0f);
  return make_uint2(__byte_perm(f0, f1, 0x7632), __byte_perm(f2, f3, 0x7632));
}

// The observation tile of a unit -- frame rows cr0-1 .. cr0+7 of frame n, all W pixels x 4 channels --
// arrives by ONE TMA tensor copy (cp.async.bulk.tensor.3d over the [N][H][W] uint32 view of the
// frames; rows above / below the frame are zero-filled by the TMA unit: the 'same' padding costs
// nothing) into one of two raw stages; the copy of the NEXT unit's tile is in flight while this
// unit converts, multiplies and pools.
template <int kCpRows>
__global__ void __launch_bounds__(kCpThreadsF, kCpRows == 2 ? 3 : 2) conv0pool_kernel(const __grid_constant__ CUtensorMap tm_frames,
                                                                    const Conv0PoolArgs a) {
  constexpr int kCpMaxBlocks = kCpRows == 2 ? 4 : 6;
  constexpr uint32_t kTmemCols = kCpRows == 2 ? 128u : 256u;
  extern __shared__ __align__(128) uint8_t smem_raw[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int W = a.W, H = a.H, SW = W + 2;
  const int npair = kCpMaxBlocks * 128 + 2 * SW + 8;          // pair entries an MMA may touch
  uint4* s_p = reinterpret_cast<uint4*>(smem_raw);            // pair array, 16 B per entry
  float* s_out = reinterpret_cast<float*>(smem_raw + (size_t)npair * 16);           // [positions][20] fp32
  uint8_t* s_bq = reinterpret_cast<uint8_t*>(s_out + (size_t)kCpMaxBlocks * 128 * kCpOutStride);
  s_bq = reinterpret_cast<uint8_t*>(((uintptr_t)s_bq + 127) & ~(uintptr_t)127);     // B: 48 x 32 bf16 = 3 KB
  float* s_bias = reinterpret_cast<float*>(s_bq + 48 * 32 * 2);
  uint64_t* s_bar = reinterpret_cast<uint64_t*>(s_bias + 16);
  uint64_t* s_full = s_bar + 1;                               // [2] TMA stage filled
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(s_full + 2);
  constexpr int kRawRows = 2 * kCpRows + 3;                   // frame rows per tile (incl. the halo rows)
  uint32_t* s_raw = reinterpret_cast<uint32_t*>(((uintptr_t)(s_tmem + 4) + 127) & ~(uintptr_t)127);   // [2][kRawRows][W]
  const uint32_t raw_bytes = (uint32_t)(kRawRows * W) * 4u;
  const uint32_t raw_stride = (raw_bytes + 127u) & ~127u;     // stage pitch (TMA destinations are 128-byte aligned)

  // ---- one-time setup: B operand (K-major, [N = 32][K = 48]: hi(w) | lo(w); k = kh*16 + kw*4 + ci,
  //      the 4th tap of a row has zero weights), bias, barrier, TMEM -------------------------------
  for (int i = tid; i < 48 * 32; i += kCpThreadsF) {
    const int k = i / 32, nn = i - k * 32;
    const int kh = k >> 4, kw = (k >> 2) & 3, ci = k & 3;
    float v = 0.f;
    if (kw < 3) {
      const float wv = __ldg(a.w + ((kh * 3 + kw) * 4 + ci) * 16 + (nn & 15));
      v = nn < 16 ? wv : bf16_resid(wv);
    }
    const uint32_t off = (uint32_t)(k >> 3) * 512u + (uint32_t)(nn >> 3) * 128u + (uint32_t)(nn & 7) * 16u +
                         (uint32_t)(k & 7) * 2u;
    *reinterpret_cast<__nv_bfloat16*>(s_bq + off) = __float2bfloat16_rn(v);
  }
  if (tid < 16) s_bias[tid] = __ldg(a.bias + tid);
  for (int i = tid; i < npair; i += kCpThreadsF) s_p[i] = make_uint4(0u, 0u, 0u, 0u);
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(s_bar)));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(s_full)));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(s_full + 1)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tm_frames)) : "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(s_tmem)), "r"(kTmemCols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(s_tmem);
  constexpr uint32_t idesc = umma_idesc(128, 32);
  const uint32_t a_base = smem_u32(s_p), b_base = smem_u32(s_bq);

  const int bands = (a.Ho + kCpRows - 1) / kCpRows;
  const int units = a.N * bands;
  uint32_t phase = 0;
  bool timed_out = false;
  auto unit_geom = [&](int u, int* n, int* r0, int* r1, int* cr0, int* cr1) {
    *n = u / bands;
    const int band = u - *n * bands;
    *r0 = band * kCpRows; *r1 = min(a.Ho, *r0 + kCpRows);
    *cr0 = max(0, 2 * *r0 - a.pt); *cr1 = min(H - 1, 2 * (*r1 - 1) - a.pt + 2);
  };
  // one thread issues the tile copy of unit u into stage st: box = W pixels x kRawRows rows x 1 frame
  // starting one row above the band's first conv row (negative / >= H rows come back as zeros)
  auto issue_tile = [&](int u, int st) {
    int n, r0, r1, cr0, cr1;
    unit_geom(u, &n, &r0, &r1, &cr0, &cr1);
    const uint32_t dst = smem_u32(s_raw) + (uint32_t)st * raw_stride, bar = smem_u32(s_full + st);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(raw_bytes) : "memory");
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::
            "r"(dst), "l"(reinterpret_cast<uint64_t>(&tm_frames)), "r"(0), "r"(cr0 - 1), "r"(n), "r"(bar)
        : "memory");
  };
  if (tid == 0 && (int)blockIdx.x < units) issue_tile(blockIdx.x, 0);
  int it = 0;
  for (int u = blockIdx.x; u < units; u += gridDim.x, ++it) {
    int n, r0, r1, cr0, cr1;
    unit_geom(u, &n, &r0, &r1, &cr0, &cr1);
    const int band = u - n * bands;
    const int CR = cr1 - cr0 + 1, npos = CR * SW, nblk = (npos + 127) >> 7;
    const int st = it & 1;
    // the next unit's tile goes into the other stage (its previous contents were converted one unit ago)
    if (tid == 0 && u + (int)gridDim.x < units) issue_tile(u + gridDim.x, st ^ 1);
    if (!mbar_wait_bounded(s_full + st, (uint32_t)((it >> 1) & 1))) timed_out = true;
    // ---- 1. raw tile (row lr = frame row cr0-1+lr, W pixels) -> pair array: entry (lr, bc) =
    //      [pixel bc-1, pixel bc] of that row as bf16 x 4 each, zero at the two border columns ----
    const uint32_t* raw = s_raw + (size_t)st * (raw_stride / 4);
    for (int e = tid; e < (CR + 2) * SW; e += kCpThreadsF) {
      const int lr = (int)(__umulhi((unsigned)e, a.sw_mul) >> a.sw_sh), bc = e - lr * SW;
      const uint32_t w0 = (bc >= 1 && bc <= W) ? raw[lr * W + bc - 1] : 0u;
      const uint32_t w1 = (bc + 1 <= W) ? raw[lr * W + bc] : 0u;
      const uint2 p0 = u8x4_to_bf16x4(w0), p1 = u8x4_to_bf16x4(w1);
      s_p[e] = make_uint4(p0.x, p0.y, p1.x, p1.y);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    // ---- 2. MMAs: one per kernel row and 128-position block -------------------------------------
    if (warp == 0 && elect_one()) {
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      for (int m = 0; m < nblk; ++m) {
#pragma unroll
        for (int kh = 0; kh < 3; ++kh) {
          const uint64_t da = umma_desc(a_base + (uint32_t)(m * 128 + kh * SW) * 16u, 32u, 128u);
          const uint64_t db = umma_desc(b_base + (uint32_t)(2 * kh) * 512u, 512u, 128u);
          umma_f16(tmem_base + (uint32_t)(m * 32), da, db, idesc, kh > 0 ? 1u : 0u);
        }
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(s_bar))
                   : "memory");
    }
    if (!mbar_wait_bounded(s_bar, phase)) timed_out = true;
    phase ^= 1;
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    // ---- 3. epilogue: accumulators -> fp32 tile [position o = lr*SW + c][16 (+4 pad)] --------------
    for (int m = warp >> 2; m < nblk; m += 2) {
      const int q = warp & 3;
      const int o = m * 128 + q * 32 + lane;
      float v[32];
      tmem_ld<32>(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(m * 32), v);
      if (o < npos) {
        float4* dst = reinterpret_cast<float4*>(s_out + (size_t)o * kCpOutStride);
#pragma unroll
        for (int c4 = 0; c4 < 4; ++c4) {
          const float4 bq = reinterpret_cast<const float4*>(s_bias)[c4];
          dst[c4] = make_float4(fmaf(v[c4 * 4 + 0] + v[16 + c4 * 4 + 0], 1.0f / 255.0f, bq.x),
                                fmaf(v[c4 * 4 + 1] + v[16 + c4 * 4 + 1], 1.0f / 255.0f, bq.y),
                                fmaf(v[c4 * 4 + 2] + v[16 + c4 * 4 + 2], 1.0f / 255.0f, bq.z),
                                fmaf(v[c4 * 4 + 3] + v[16 + c4 * 4 + 3], 1.0f / 255.0f, bq.w));
        }
      }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    // ---- 4. max-pool 3x3 / 2 from the tile; thread = (pooled pixel, 8-channel group) ----------------
    const int nout = (r1 - r0) * a.Wo * 2;
    for (int i = tid; i < nout; i += kCpThreadsF) {
      const int gq = i & 1, qq = i >> 1;
      const int qr = qq / a.Wo, qw = qq - qr * a.Wo, qh = r0 + qr;
      float best[8];
      unsigned char arg[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) { best[e] = -INFINITY; arg[e] = 0; }
#pragma unroll
      for (int kh = 0; kh < 3; ++kh) {
        const int h = qh * 2 - a.pt + kh;
        if (h < 0 || h >= H) continue;
#pragma unroll
        for (int kw = 0; kw < 3; ++kw) {
          const int w = qw * 2 - a.pl + kw;
          if (w < 0 || w >= W) continue;
          const float4* s4 = reinterpret_cast<const float4*>(s_out + (size_t)((h - cr0) * SW + w) * kCpOutStride + gq * 8);
          const float4 x0 = s4[0], x1 = s4[1];
          const float vv[8] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
          const unsigned char t = (unsigned char)(kh * 3 + kw);
#pragma unroll
          for (int e = 0; e < 8; ++e)
            if (vv[e] > best[e]) { best[e] = vv[e]; arg[e] = t; }
        }
      }
      const size_t pix = ((size_t)n * a.Ho + qh) * a.Wo + qw;
      uint2 packed;
      packed.x = arg[0] | (arg[1] << 8) | (arg[2] << 16) | ((uint32_t)arg[3] << 24);
      packed.y = arg[4] | (arg[5] << 8) | (arg[6] << 16) | ((uint32_t)arg[7] << 24);
      *reinterpret_cast<uint2*>(a.idx + pix * 16 + gq * 8) = packed;
      const size_t sp = (size_t)(n * a.RHp + qh + 1) * a.PWp + qw + 1;
      const float4 va = make_float4(best[0], best[1], best[2], best[3]), vb = make_float4(best[4], best[5], best[6], best[7]);
      a.praw[(size_t)gq * a.Lpp + sp] = pack8_bf16(va, vb);
      a.praw[(size_t)(2 + gq) * a.Lpp + sp] = pack8_bf16(bf16_resid4(va), bf16_resid4(vb));
      const float4 ra = make_float4(fmaxf(va.x, 0.f), fmaxf(va.y, 0.f), fmaxf(va.z, 0.f), fmaxf(va.w, 0.f));
      const float4 rb = make_float4(fmaxf(vb.x, 0.f), fmaxf(vb.y, 0.f), fmaxf(vb.z, 0.f), fmaxf(vb.w, 0.f));
      a.prelu[(size_t)gq * a.Lpp + sp] = pack8_bf16(ra, rb);
      a.prelu[(size_t)(2 + gq) * a.Lpp + sp] = pack8_bf16(bf16_resid4(ra), bf16_resid4(rb));
    }
    // ---- padding positions of the plane tensors owned by this unit: zeros -------------------------
    {
      // per pooled row: columns 0 and Wo+1; band 0 also the separator row above the image; the last
      // unit also the tail [N*RHp*PWp, Lpp)
      const uint4 z = make_uint4(0u, 0u, 0u, 0u);
      const int rows = r1 - r0;
      for (int i = tid; i < rows * 2 * 4; i += kCpThreadsF) {
        const int pl_ = i & 3, side = (i >> 2) & 1, rr = i >> 3;
        const size_t sp = (size_t)(n * a.RHp + r0 + rr + 1) * a.PWp + (side ? a.Wo + 1 : 0);
        a.praw[(size_t)pl_ * a.Lpp + sp] = z;
        a.prelu[(size_t)pl_ * a.Lpp + sp] = z;
      }
      if (band == 0) {
        for (int i = tid; i < a.PWp * 4; i += kCpThreadsF) {
          const int pl_ = i & 3, cc = i >> 2;
          const size_t sp = (size_t)(n * a.RHp) * a.PWp + cc;
          a.praw[(size_t)pl_ * a.Lpp + sp] = z;
          a.prelu[(size_t)pl_ * a.Lpp + sp] = z;
        }
      }
      if (u == units - 1) {
        const int t0 = a.N * a.RHp * a.PWp;
        for (int i = tid; i < (a.Lpp - t0) * 4; i += kCpThreadsF) {
          const int pl_ = i & 3, cc = i >> 2;
          a.praw[(size_t)pl_ * a.Lpp + t0 + cc] = z;
          a.prelu[(size_t)pl_ * a.Lpp + t0 + cc] = z;
        }
      }
    }
    __syncthreads();      // s_out and the pair array are rewritten by the next unit
  }
  if (timed_out && a.err) atomicExch(a.err, 1);
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols));
  }
}

static int g_c0_rows = getenv("SEEDRL_C0_ROWS") ? (atoi(getenv("SEEDRL_C0_ROWS")) == 2 ? 2 : 3) : 3;
void conv0pool_set_rows(int rows) { g_c0_rows = rows == 2 ? 2 : 3; }
static int c0_rows(int W) {                      // pooled rows per unit that fit the block budget
  if (g_c0_rows == 2 && 5 * (W + 2) <= 4 * 128) return 2;
  return 3;
}
bool conv0pool_supported(int cin, int cout, int H, int W) {
  // (W % 4: the TMA row pitch W*4 bytes must be a multiple of 16; W <= 256: box width)
  return cin == 4 && cout == 16 && W % 4 == 0 && W <= 256 && 7 * (W + 2) <= 6 * 128 && H >= 3 && W >= 3;
}

template <int ROWS>
static int launch_conv0pool(Conv0PoolArgs a, const uint8_t* frames, cudaStream_t st) {
  constexpr int MAXB = ROWS == 2 ? 4 : 6;
  const int N = a.N, H = a.H, W = a.W;
  const size_t raw_stride = (((size_t)(2 * ROWS + 3) * W * 4) + 127) / 128 * 128;
  const size_t smem = (size_t)(MAXB * 128 + 2 * (W + 2) + 8) * 16 + (size_t)MAXB * 128 * kCpOutStride * 4 + 128 +
                      48 * 32 * 2 + 16 * 4 + 64 + 128 + 2 * raw_stride;
  static bool attr = false;
  if (!attr) {
    SEEDRL_CUDA(cudaFuncSetAttribute(conv0pool_kernel<ROWS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 112 * 1024));
    attr = true;
  }
  if (smem > 112 * 1024) return set_error(SEEDRL_ERR_INVALID_ARGUMENT, "conv0pool: image too wide");
  // [N][H][W] view of the frames with one uint32 (= 4 uint8 channels) per pixel
  typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  static EncodeTiledFn enc = nullptr;
  if (!enc) {
    void* q = nullptr;
    cudaDriverEntryPointQueryResult qr;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &q, cudaEnableDefault, &qr) != cudaSuccess ||
        qr != cudaDriverEntryPointSuccess)
      return set_error(SEEDRL_ERR_INTERNAL, "cuTensorMapEncodeTiled is not available");
    enc = reinterpret_cast<EncodeTiledFn>(q);
  }
  CUtensorMap tm;
  const cuuint64_t gdim[3] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
  const cuuint64_t gstr[2] = {(cuuint64_t)W * 4, (cuuint64_t)H * W * 4};
  const cuuint32_t box[3] = {(cuuint32_t)W, (cuuint32_t)(2 * ROWS + 3), 1};
  const cuuint32_t estr[3] = {1, 1, 1};
  if (enc(&tm, CU_TENSOR_MAP_DATA_TYPE_UINT32, 3, const_cast<uint8_t*>(frames), gdim, gstr, box, estr,
          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
    return set_error(SEEDRL_ERR_INTERNAL, "conv0pool: cuTensorMapEncodeTiled failed");
  const int units = N * ((a.Ho + ROWS - 1) / ROWS);
  const int per_sm = ROWS == 2 ? 3 : 2;
  const int grid = units < per_sm * kNumSMs ? units : per_sm * kNumSMs;
  conv0pool_kernel<ROWS><<<grid, kCpThreadsF, smem, st>>>(tm, a);
  count_launch(PC_CONV_FWD, st);
  SEEDRL_CHECK_LAUNCH();
  return SEEDRL_OK;
}

int conv0pool_forward(int N, int H, int W, const uint8_t* frames, const float* w, const float* bias, void* praw,
                      void* prelu, uint8_t* idx, int* err, cudaStream_t st) {
  Conv0PoolArgs a;
  a.N = N; a.H = H; a.W = W;
  same_pad3s2_(H, &a.Ho, &a.pt);
  same_pad3s2_(W, &a.Wo, &a.pl);
  a.Lpp = (int)planes_positions(N, a.Ho, a.Wo); a.PWp = a.Wo + 2; a.RHp = a.Ho + 1;
  fast_div_setup((unsigned int)(W + 2), &a.sw_mul, &a.sw_sh);
  a.frames = frames; a.w = w; a.bias = bias;
  a.praw = reinterpret_cast<uint4*>(praw); a.prelu = reinterpret_cast<uint4*>(prelu); a.idx = idx; a.err = err;
  return c0_rows(W) == 2 ? launch_conv0pool<2>(a, frames, st) : launch_conv0pool<3>(a, frames, st);
}

}  // namespace seedrl
