// "Planes" convolution path (conv_mode 3, 'tc3p'): activations and gradients of the 16/32-channel
// layers of dmlab/networks.py:26-60 live in HBM in the tensor core's own operand format, so that
// every 3x3 convolution (forward, data gradient, weight gradient) is
//      TMA tile (cp.async.bulk.tensor)  ->  tcgen05.mma  ->  epilogue
// with no thread ever touching an input element.
//
// HBM format of a [N, H, W, C] activation ("plane tensor"):
//   the N images form one zero-padded tall image (PW = W + 2 columns, RH = H + 1 rows per image,
//   one shared zero row between images); pixel (n, h, w) sits at storage position
//       s = (n * RH + h + 1) * PW + (w + 1),          0 <= s < Lp,
//   and the tensor is 2 * C/8 planes [Lp][8 ch] bf16 (16 bytes per position): first the C/8 "hi"
//   planes (bf16(v)), then the C/8 "lo" planes (bf16(v - hi)) -- v = hi + lo to ~2^-17 relative,
//   the bf16x3 operand split of the 'tc3' mode, done ONCE by the producing kernel's epilogue
//   instead of by every consumer.  Padding positions hold zeros.
//   One plane is exactly the canonical no-swizzle UMMA layout (core matrix = 8 positions x 16 B):
//   K-major A operand of the forward / data-gradient GEMM (M = positions, K = channels: LBO = plane
//   stride, SBO = 128 B) and MN-major operand of the weight-gradient GEMM (K = positions: LBO =
//   128 B, SBO = plane stride); a filter tap (kh, kw) is the descriptor start address moved by
//   (kh * PW + kw) * 16 bytes.
//
// Kernels (all persistent, 1 CTA / SM, warp-specialised, mbarrier pipelines, bounded waits):
//   convp_kernel<CIN, COUT, NSUB>   forward / data gradient.  warp 0 = TMA producer (one
//       cp.async.bulk.tensor.3d per tile: box = 16-position chunks x (2 * CIN/8 planes), two
//       smem stages), warp 1 = MMA issuer (NSUB*9*CIN/16*3 tcgen05.mma kind::f16 128 x COUT x 16
//       per tile into one of TWO TMEM accumulator sets, tcgen05.commit -> stage-empty and
//       accumulator-full barriers), warps 4..11 = epilogue (tcgen05.ld -> bias / ReLU-mask /
//       residual -> hi/lo split -> coalesced 16-byte plane stores; optionally a second, ReLU'd
//       copy for the next conv, or fp32 NHWC for the max-pool / Dense consumers).
//   wgradp_kernel<CP, COUT, KC>     weight + bias gradient: M = (kw, ci) rows from three
//       kw-shifted TMA copies of the x planes (+ a constant ones row whose accumulator is the bias
//       gradient), N = c_out from the dy planes, K = positions, one TMEM accumulator per kernel
//       row kh; per-CTA partials reduced in fixed order by the deferred reduce of conv_tc_kernels.cu.
//   poolp_fwd / poolp_bwd / to_planes / from_planes: elementwise format kernels.
#include <cuda.h>
#include <cuda_bf16.h>

#include <cstdio>
#include <cstdlib>

#include "kernels.h"
#include "tc_common.cuh"

namespace seedrl {

// ------------------------------------------------------------------------------------------------
// geometry
long long planes_positions(int N, int H, int W) {
  const long long Q = (long long)N * (H + 1) * (W + 2);
  // every storage position a consumer's TMA box can touch inside its declared extent is written
  // (zeros) by the producer: tiles read up to 2*PW+2 past Q, shifted maps drop up to 64 positions
  return ((Q + 2 * (W + 2) + 2 + 256 + 127) / 128) * 128;
}
size_t planes_bytes(int N, int H, int W, int C) {
  return (size_t)planes_positions(N, H, W) * 16 * 2 * (C / 8);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* q = nullptr;
    cudaDriverEntryPointQueryResult qr;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &q, cudaEnableDefault, &qr) == cudaSuccess &&
        qr == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(q);
    (void)cudaGetLastError();
  }
  return fn;
}

// Positions per TMA box row.  The TMA engine pays a fixed cost per box row, so rows are as long as
// the 256-element box limit allows: 128 positions x 16 B = 256 x uint64 (measured: 16-position rows
// made the conv kernels TMA-issue-bound at ~40 cycles per row).  Tiles start on multiples of it.
static int g_chunk = 0;
int planes_chunk() {
  if (g_chunk == 0) {
    const char* e = getenv("SEEDRL_PLANES_CHUNK");
    g_chunk = e ? atoi(e) : 128;
    if (g_chunk != 16 && g_chunk != 32 && g_chunk != 64 && g_chunk != 128) g_chunk = 128;
  }
  return g_chunk;
}

// 3-D view of a plane tensor starting `shift` positions in: {one chunk of CH positions as 2*CH
// uint64, chunks, planes}; box = {2*CH, box_chunks, box_planes}.  Out-of-extent chunks read as zeros.
static int make_plane_map(CUtensorMap* tm, const void* base, long long Lp, int planes, int shift,
                          int CH, int box_chunks, int box_planes) {
  EncodeTiledFn enc = encode_tiled_fn();
  if (!enc) return set_error(SEEDRL_ERR_INTERNAL, "cuTensorMapEncodeTiled is not available");
  if (box_chunks < 1 || box_chunks > 256 || box_planes < 1 || box_planes > planes)
    return set_error(SEEDRL_ERR_INVALID_ARGUMENT, "conv planes: TMA box out of range");
  const cuuint64_t gdim[3] = {(cuuint64_t)(2 * CH), (cuuint64_t)((Lp - shift) / CH), (cuuint64_t)planes};
  const cuuint64_t gstr[2] = {(cuuint64_t)CH * 16, (cuuint64_t)Lp * 16};
  const cuuint32_t box[3] = {(cuuint32_t)(2 * CH), (cuuint32_t)box_chunks, (cuuint32_t)box_planes};
  const cuuint32_t estr[3] = {1, 1, 1};
  void* addr = const_cast<char*>(reinterpret_cast<const char*>(base)) + (size_t)shift * 16;
  const CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_UINT64, 3, addr, gdim, gstr, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                         CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    char msg[128];
    snprintf(msg, sizeof msg, "cuTensorMapEncodeTiled failed (%d) Lp=%lld planes=%d shift=%d box=%dx%d", (int)r, Lp,
             planes, shift, box_chunks, box_planes);
    return set_error(SEEDRL_ERR_INTERNAL, msg);
  }
  return SEEDRL_OK;
}

__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}\n" ::"r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* tm, int c0, int c1, int c2,
                                            uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::
          "r"(smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(tm)), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ float bf16lo(uint32_t w) { return __uint_as_float(w << 16); }
__device__ __forceinline__ float bf16hi(uint32_t w) { return __uint_as_float(w & 0xFFFF0000u); }

// ------------------------------------------------------------------------------------------------
// forward / data gradient
struct ConvpArgs {
  ConvGeom g;
  int Lp;                      // storage positions per plane (input and output share the geometry)
  int nch;                     // chunks per staged tile (box_chunks of the map)
  int chunk;                   // positions per chunk
  int ntiles;
  const uint4* wq;             // packed weights [hi | lo], conv_tc_kernels.cu layout
  const float* bias;           // [COUT] or null
  const uint4* mask;           // hi planes of the ReLU'd forward activation (COUT/8 planes) or null
  const uint4* res;            // residual plane tensor (COUT channels) or null
  uint4* out_raw;              // plane tensor or null
  uint4* out_relu;             // plane tensor (ReLU applied) or null
  float* out_nhwc;             // fp32 [N,H,W,COUT] or null
  int* err;
};

constexpr int kCpThreads = 384;
constexpr int kCpM = 128;

template <int CIN, int COUT, int NSUB>
__global__ void __launch_bounds__(kCpThreads, 1)
convp_kernel(const __grid_constant__ CUtensorMap tm_in, const ConvpArgs a) {
  constexpr int G = CIN / 8, GO = COUT / 8, NS = CIN / 16;
  constexpr int MT = NSUB * kCpM;
  // one accumulator set: per M block 2*COUT columns [a*hi(w) + lo(a)*hi(w) | hi(a)*lo(w)]
  constexpr int ACC_COLS = NSUB * 2 * COUT;
  constexpr int TCOLS = 2 * ACC_COLS <= 32 ? 32 : (2 * ACC_COLS <= 64 ? 64 : (2 * ACC_COLS <= 128 ? 128
                        : (2 * ACC_COLS <= 256 ? 256 : 512)));
  static_assert(2 * ACC_COLS <= 512, "TMEM has 512 columns");
  constexpr int NEPI = NSUB >= 2 ? 8 : 4;                // epilogue warps that own an M block
  extern __shared__ __align__(128) uint8_t smem_raw[];
  const int PW = a.g.PW;
  const uint32_t P = (uint32_t)(a.nch * a.chunk) * 16u;  // plane stride in a stage (bytes)
  const uint32_t stage_bytes = 2u * G * P;
  uint8_t* s_stage = smem_raw;                           // [2][hi G planes | lo G planes]
  uint4* s_b = reinterpret_cast<uint4*>(smem_raw + 2 * (size_t)stage_bytes);   // 2 * 9*CIN*COUT bf16
  float* s_bias = reinterpret_cast<float*>(s_b + 2 * 9 * CIN * COUT / 8);
  uint64_t* s_full = reinterpret_cast<uint64_t*>(s_bias + COUT);
  uint64_t* s_empty = s_full + 2;
  uint64_t* s_tfull = s_empty + 2;
  uint64_t* s_tempty = s_tfull + 2;
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(s_tempty + 2);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  for (int i = tid; i < 2 * 9 * CIN * COUT / 8; i += kCpThreads) s_b[i] = __ldg(a.wq + i);
  if (tid < COUT) s_bias[tid] = a.bias ? __ldg(a.bias + tid) : 0.f;
  if (tid == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(s_full + i, 1);
      mbar_init(s_empty + i, 1);
      mbar_init(s_tfull + i, 1);
      mbar_init(s_tempty + i, NEPI);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tm_in)) : "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(s_tmem)),
                 "r"(TCOLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // weights: generic -> async proxy
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(s_tmem);

  const int my_tiles = ((int)blockIdx.x < a.ntiles) ? (a.ntiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
  bool timed_out = false;

  if (warp == 0) {
    // ================================ TMA producer ============================================
    if (elect_one()) {
      for (int it = 0; it < my_tiles; ++it) {
        const int s = it & 1;
        if (it >= 2 && !mbar_wait_bounded(s_empty + s, (uint32_t)(((it >> 1) - 1) & 1))) { timed_out = true; break; }
        const int tile = (int)blockIdx.x + it * (int)gridDim.x;
        mbar_expect_tx(s_full + s, stage_bytes);
        tma_load_3d(s_stage + (size_t)s * stage_bytes, &tm_in, 0, tile * MT / a.chunk, 0, s_full + s);
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ================================ MMA issuer ==============================================
    // The tensor core reads its A tile (128 positions x 16 channels = 4 KB) from shared memory
    // at ~128 B/clk for EVERY instruction -- that, not the FLOP rate, bounds small-N MMAs (measured:
    // 39 clk per 128x16x16).  So the bf16x3 product is issued as TWO reads of the activations per
    // (tap, slab): hi(a) x [hi(w) | lo(w)] as one N = 2*COUT instruction, lo(a) x hi(w) accumulated
    // onto its first half; the epilogue adds the halves.
    constexpr uint32_t idesc2 = umma_idesc(kCpM, 2 * COUT), idesc1 = umma_idesc(kCpM, COUT);
    const uint32_t b_base = smem_u32(s_b);
    for (int it = 0; it < my_tiles; ++it) {
      const int s = it & 1;
      if (!mbar_wait_bounded(s_full + s, (uint32_t)((it >> 1) & 1))) { timed_out = true; break; }
      if (it >= 2 && !mbar_wait_bounded(s_tempty + s, (uint32_t)(((it >> 1) - 1) & 1))) { timed_out = true; break; }
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (elect_one()) {
        const uint32_t a_base = smem_u32(s_stage + (size_t)s * stage_bytes);
        const uint32_t d_base = tmem_base + (uint32_t)(s * ACC_COLS);
        // descriptors are advanced by adding to their address field (16-byte units)
        const uint64_t da0 = umma_desc(a_base, P, 128u);
        const uint64_t db0 = umma_desc(b_base, (uint32_t)(2 * GO) * 128u, 128u);
        const uint64_t lo_off = (uint64_t)((G * P) >> 4), slab_off = (uint64_t)((2 * P) >> 4);
#pragma unroll 1
        for (int m = 0; m < NSUB; ++m) {
          uint32_t acc = 0;
          const uint32_t d = d_base + (uint32_t)(m * 2 * COUT);
#pragma unroll
          for (int tap = 0; tap < 9; ++tap) {
            const uint64_t off = (uint64_t)(m * kCpM + (tap / 3) * PW + (tap % 3));
#pragma unroll
            for (int sl = 0; sl < NS; ++sl) {
              const uint64_t da = da0 + off + (uint64_t)sl * slab_off;
              const uint64_t db = db0 + (uint64_t)((tap * NS + sl) * (COUT * 64 / 16));
              umma_f16(d, da, db, idesc2, acc);
              umma_f16(d, da + lo_off, db, idesc1, 1u);
              acc = 1;
            }
          }
        }
        umma_commit(s_empty + s);       // this stage's smem may be refilled once these MMAs retire
        umma_commit(s_tfull + s);       // ... and the accumulator set is complete
      }
      __syncwarp();
    }
  } else if (warp >= 4) {
    // ================================ epilogue ================================================
    const int q = warp & 3;                       // TMEM lane quadrant this warp may read
    const int half = (warp - 4) >> 2;             // 0 / 1: which M blocks of a tile
    const size_t plane_u = (size_t)a.Lp;          // plane stride in 16-byte units (global)
    if (blockIdx.x == 0) {                        // head margin s in [0, PW + 1): zeros
      const int et = tid - 128;
      const uint4 z = make_uint4(0u, 0u, 0u, 0u);
      for (int i = et; i < (PW + 1) * 2 * GO; i += kCpThreads - 128) {
        const int pl = i / (PW + 1), s = i - pl * (PW + 1);
        if (a.out_raw) a.out_raw[(size_t)pl * plane_u + s] = z;
        if (a.out_relu) a.out_relu[(size_t)pl * plane_u + s] = z;
      }
    }
    const bool active = NSUB >= 2 || half == 0;
    for (int it = 0; it < my_tiles && active; ++it) {
      const int as = it & 1;
      const int tile = (int)blockIdx.x + it * (int)gridDim.x;
      if (!mbar_wait_bounded(s_tfull + as, (uint32_t)((it >> 1) & 1))) { timed_out = true; break; }
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      constexpr int NB = NSUB >= 2 ? NSUB / 2 : 1;     // M blocks per warp
#pragma unroll 1
      for (int bi = 0; bi < NB; ++bi) {
        const int m = NSUB >= 2 ? half + 2 * bi : 0;
        const int p = tile * MT + m * kCpM + q * 32 + lane;
        const int s = p + PW + 1;
        const int pix = out_pixel(a.g, p);
        const bool in_store = s < a.Lp;
        // residual / mask operands of this position (issued before the TMEM read)
        uint4 rh[GO], rl[GO], mk[GO];
#pragma unroll
        for (int go = 0; go < GO; ++go) {
          rh[go] = make_uint4(0u, 0u, 0u, 0u); rl[go] = rh[go]; mk[go] = rh[go];
          if (pix >= 0) {
            if (a.res) {
              rh[go] = __ldg(a.res + (size_t)go * plane_u + s);
              rl[go] = __ldg(a.res + (size_t)(GO + go) * plane_u + s);
            }
            if (a.mask) mk[go] = __ldg(a.mask + (size_t)go * plane_u + s);
          }
        }
        float v[COUT];
#pragma unroll
        for (int hc = 0; hc < COUT / 16; ++hc) {
          float v2[16];
          const uint32_t col = (uint32_t)(as * ACC_COLS + m * 2 * COUT + hc * 16);
          tmem_ld<16>(tmem_base + ((uint32_t)(q * 32) << 16) + col, v + hc * 16);
          tmem_ld<16>(tmem_base + ((uint32_t)(q * 32) << 16) + col + COUT, v2);
#pragma unroll
          for (int e = 0; e < 16; ++e) v[hc * 16 + e] += v2[e];
        }
        if (bi == NB - 1) {          // accumulator set drained: hand it back to the MMA warp
          asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
          __syncwarp();
          if (lane == 0) mbar_arrive(s_tempty + as);
        }
#pragma unroll
        for (int go = 0; go < GO; ++go) {
          float x[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) x[e] = v[go * 8 + e] + s_bias[go * 8 + e];
          if (a.mask) {
            const uint32_t mw[4] = {mk[go].x, mk[go].y, mk[go].z, mk[go].w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              x[2 * e] = bf16lo(mw[e]) > 0.f ? x[2 * e] : 0.f;
              x[2 * e + 1] = bf16hi(mw[e]) > 0.f ? x[2 * e + 1] : 0.f;
            }
          }
          if (a.res) {
            const uint32_t hw[4] = {rh[go].x, rh[go].y, rh[go].z, rh[go].w};
            const uint32_t lw[4] = {rl[go].x, rl[go].y, rl[go].z, rl[go].w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              x[2 * e] += bf16lo(hw[e]) + bf16lo(lw[e]);
              x[2 * e + 1] += bf16hi(hw[e]) + bf16hi(lw[e]);
            }
          }
          if (pix < 0) {
#pragma unroll
            for (int e = 0; e < 8; ++e) x[e] = 0.f;
          }
          const float4 xa = make_float4(x[0], x[1], x[2], x[3]), xb = make_float4(x[4], x[5], x[6], x[7]);
          if (a.out_raw && in_store) {
            a.out_raw[(size_t)go * plane_u + s] = pack8_bf16(xa, xb);
            a.out_raw[(size_t)(GO + go) * plane_u + s] = pack8_bf16(bf16_resid4(xa), bf16_resid4(xb));
          }
          if (a.out_relu && in_store) {
            const float4 ra = make_float4(fmaxf(xa.x, 0.f), fmaxf(xa.y, 0.f), fmaxf(xa.z, 0.f), fmaxf(xa.w, 0.f));
            const float4 rb = make_float4(fmaxf(xb.x, 0.f), fmaxf(xb.y, 0.f), fmaxf(xb.z, 0.f), fmaxf(xb.w, 0.f));
            a.out_relu[(size_t)go * plane_u + s] = pack8_bf16(ra, rb);
            a.out_relu[(size_t)(GO + go) * plane_u + s] = pack8_bf16(bf16_resid4(ra), bf16_resid4(rb));
          }
          if (a.out_nhwc && pix >= 0) {
            float4* o = reinterpret_cast<float4*>(a.out_nhwc + (size_t)pix * COUT + go * 8);
            o[0] = xa; o[1] = xb;
          }
        }
      }
    }
  }
  if (timed_out && a.err) atomicExch(a.err, 1);
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TCOLS));
  }
}

template <int CIN, int COUT, int NSUB>
static int launch_convp(const PlaneConv& c, cudaStream_t st) {
  const ConvGeom g = make_geom(c.N, c.H, c.W);
  constexpr int MT = NSUB * kCpM;
  const long long Lp = planes_positions(c.N, c.H, c.W);
  if (Lp + MT + 4 * g.PW >= (1LL << 31))
    return set_error(SEEDRL_ERR_INVALID_ARGUMENT, "convp: batch too large for 32-bit positions");
  const int L = MT + 2 * g.PW + 2;
  const int CH = planes_chunk();
  const int nch = (L + CH - 1) / CH;
  const size_t stage = (size_t)2 * (CIN / 8) * nch * CH * 16;
  const size_t smem = 2 * stage + (size_t)2 * 9 * CIN * COUT * 2 + COUT * 4 + 8 * 8 + 16;
  if (smem > 227 * 1024) return kPlanesTryNext;
  CUtensorMap tm;
  SEEDRL_TRY_RC(make_plane_map(&tm, c.in, Lp, 2 * (CIN / 8), 0, CH, nch, 2 * (CIN / 8)));
  static bool attr = false;
  if (!attr) {
    SEEDRL_CUDA(cudaFuncSetAttribute(convp_kernel<CIN, COUT, NSUB>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     227 * 1024));
    attr = true;
  }
  ConvpArgs a;
  a.g = g; a.Lp = (int)Lp; a.nch = nch; a.chunk = CH;
  a.ntiles = (int)((Lp - g.PW - 1 + MT - 1) / MT);       // every storage position >= PW + 1 is written
  a.wq = reinterpret_cast<const uint4*>(c.wq); a.bias = c.bias;
  a.mask = reinterpret_cast<const uint4*>(c.mask); a.res = reinterpret_cast<const uint4*>(c.res);
  a.out_raw = reinterpret_cast<uint4*>(c.out_raw); a.out_relu = reinterpret_cast<uint4*>(c.out_relu);
  a.out_nhwc = c.out_nhwc; a.err = c.err;
  const int grid = a.ntiles < kNumSMs ? a.ntiles : kNumSMs;
  convp_kernel<CIN, COUT, NSUB><<<grid, kCpThreads, smem, st>>>(tm, a);
  count_launch(g_conv_cat, st);
  SEEDRL_CHECK_LAUNCH();
  return SEEDRL_OK;
}

bool convp_supported(int cin, int cout) {
  return (cin == 16 || cin == 32) && (cout == 16 || cout == 32);
}

int convp_forward(int cin, int cout, const PlaneConv& c, cudaStream_t st) {
  // big tiles amortise the halo; small problems (inference batches) take 128-position tiles so
  // that every SM still gets work
  const long long Lp = planes_positions(c.N, c.H, c.W);
  const bool small = Lp / 512 < 2 * kNumSMs;
#define SEEDRL_CP_CASE(CI, CO_)                                                   \
  if (cin == CI && cout == CO_) {                                                 \
    int rc = kPlanesTryNext;                                                      \
    if (!small) rc = launch_convp<CI, CO_, 4>(c, st);                             \
    if (rc == kPlanesTryNext) rc = launch_convp<CI, CO_, 2>(c, st);               \
    if (rc == kPlanesTryNext) rc = launch_convp<CI, CO_, 1>(c, st);               \
    if (rc == kPlanesTryNext) return set_error(SEEDRL_ERR_INVALID_ARGUMENT, "convp: image too wide"); \
    return rc;                                                                    \
  }
  SEEDRL_CP_CASE(16, 16)
  SEEDRL_CP_CASE(16, 32)
  SEEDRL_CP_CASE(32, 16)
  SEEDRL_CP_CASE(32, 32)
#undef SEEDRL_CP_CASE
  return set_error(SEEDRL_ERR_INVALID_ARGUMENT, "convp: unsupported (cin,cout)");
}

// ------------------------------------------------------------------------------------------------
// weight + bias gradient
//   dW[kh][kw][ci][co] = sum_p x~[p + kh*PW + kw][ci] * dy[p][co]
//                      = sum_j x~[j + PW + kw][ci] * dy[j + (1 - kh)*PW][co]        (j = p + (kh-1)*PW)
// GEMM with K = positions j (chunks of KC), M = (kw, ci) rows, N = (kh, co) columns:
//   A  = three kw-shifted TMA copies of the x planes (no halo) + one constant plane whose first
//        channel is 1 (its accumulator row is the bias gradient);  MN-major, 128 rows
//   B  = three kh-shifted TMA copies of the dy planes; MN-major, hi planes then lo planes
// so the activation tile -- the operand whose shared-memory read (4 KB per instruction at
// ~128 B/clk) bounds small-N MMAs -- is read TWICE per 16 positions:
//   hi(x) x [hi(dy) kh=0..2 | lo(dy) kh=0..2]   N = 6*COUT   -> D[:, 0 : 6*COUT]
//   lo(x) x  hi(dy) kh=0..2                     N = 3*COUT   -> accumulated onto D[:, 0 : 3*COUT]
// (the first version kept kh as a start-address offset of the x planes: three accumulators, nine
// instructions and nine reads of the x tile per 16 positions -- the tensor pipe was 92 % busy at 19 %
// of HBM bandwidth).  Terms with j < 0 pair the zero row above the first image with dy and vanish.
struct WgradpArgs {
  int PW, nchx, nchunks, nb;            // TMA chunks per box (x and dy alike); K chunks; stages
  int chunk;                            // positions per TMA chunk
  float* partial;                       // [grid][9*CIN*COUT + COUT]
  int* err;
};

constexpr int kWpThreads = 192;
constexpr int kWpMaxStages = 4;

struct WgradMaps { CUtensorMap x[3], dy[3]; };

template <int CP, int COUT, int KC>
__global__ void __launch_bounds__(kWpThreads, 1)
wgradp_kernel(const __grid_constant__ WgradMaps tm, const WgradpArgs a) {
  constexpr int G = CP / 8, GO = COUT / 8;
  constexpr int XG = 3 * G + 1;                          // M groups per half: (kw, g) planes + ones/zeros plane
  // 16-channel inputs fill only 49 of an M = 128 instruction's rows; M = 64 (8 row groups >= XG = 7)
  // reads half the A tile per instruction -- the small-N MMA is bound by that read
  constexpr int MROWS = XG <= 8 ? 64 : 128;
  constexpr int TCOLS = 6 * COUT <= 128 ? 128 : 256;
  constexpr int NW = 9 * CP * COUT + COUT;
  extern __shared__ __align__(128) uint8_t smem_raw[];
  const int nb = a.nb;
  const uint32_t Pk = (uint32_t)(a.nchx * a.chunk) * 16u;   // plane stride (bytes), x and dy alike
  const uint32_t xh_bytes = (uint32_t)XG * Pk;
  const uint32_t dh_bytes = (uint32_t)(3 * GO) * Pk;
  const uint32_t stage_bytes = 2u * xh_bytes + 2u * dh_bytes;   // [x hi | x lo | dy hi (kh,go) | dy lo (kh,go)]
  uint64_t* s_full = reinterpret_cast<uint64_t*>(smem_raw);
  uint64_t* s_empty = s_full + kWpMaxStages;
  uint64_t* s_done = s_empty + kWpMaxStages;
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(s_done + 1);
  uint8_t* s_stage = smem_raw + 128;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  // the constant planes of every stage: ones (x hi, channel 0 of each position) and zeros (x lo)
  for (int i = tid; i < nb * (int)(Pk / 16); i += kWpThreads) {
    const int sg = i / (int)(Pk / 16), k = i - sg * (int)(Pk / 16);
    uint4* xh = reinterpret_cast<uint4*>(s_stage + (size_t)sg * stage_bytes + (size_t)(XG - 1) * Pk);
    uint4* xl = reinterpret_cast<uint4*>(s_stage + (size_t)sg * stage_bytes + xh_bytes + (size_t)(XG - 1) * Pk);
    xh[k] = make_uint4(0x00003F80u, 0u, 0u, 0u);        // bf16 1.0 in channel 0
    xl[k] = make_uint4(0u, 0u, 0u, 0u);
  }
  if (tid == 0) {
    for (int i = 0; i < kWpMaxStages; ++i) { mbar_init(s_full + i, 1); mbar_init(s_empty + i, 1); }
    mbar_init(s_done, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(s_tmem)),
                 "r"(TCOLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(s_tmem);

  const int my_chunks = ((int)blockIdx.x < a.nchunks) ? (a.nchunks - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
  bool timed_out = false;

  if (warp == 0) {
    if (elect_one()) {
      const uint32_t tx = (uint32_t)(6 * G + 6 * GO) * Pk;
      for (int it = 0; it < my_chunks; ++it) {
        const int s = it % nb;
        if (it >= nb && !mbar_wait_bounded(s_empty + s, (uint32_t)(((it / nb) - 1) & 1))) { timed_out = true; break; }
        const int c0 = ((int)blockIdx.x + it * (int)gridDim.x) * KC / a.chunk;
        uint8_t* base = s_stage + (size_t)s * stage_bytes;
        mbar_expect_tx(s_full + s, tx);
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          tma_load_3d(base + (size_t)(k * G) * Pk, &tm.x[k], 0, c0, 0, s_full + s);                  // x hi, kw = k
          tma_load_3d(base + xh_bytes + (size_t)(k * G) * Pk, &tm.x[k], 0, c0, G, s_full + s);       // x lo
          tma_load_3d(base + 2 * (size_t)xh_bytes + (size_t)(k * GO) * Pk, &tm.dy[k], 0, c0, 0, s_full + s);   // dy hi, kh = k
          tma_load_3d(base + 2 * (size_t)xh_bytes + dh_bytes + (size_t)(k * GO) * Pk, &tm.dy[k], 0, c0, GO, s_full + s);
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // both operands MN-major
    constexpr uint32_t idesc6 = umma_idesc(MROWS, 6 * COUT) | (1u << 15) | (1u << 16);
    constexpr uint32_t idesc3 = umma_idesc(MROWS, 3 * COUT) | (1u << 15) | (1u << 16);
    for (int it = 0; it < my_chunks; ++it) {
      const int s = it % nb;
      if (!mbar_wait_bounded(s_full + s, (uint32_t)((it / nb) & 1))) { timed_out = true; break; }
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (elect_one()) {
        const uint32_t xb = smem_u32(s_stage + (size_t)s * stage_bytes);
        const uint32_t db = xb + 2u * xh_bytes;
        // descriptors with start address 0 (the address field counts 16-byte units)
        const uint64_t d0 = umma_desc(0u, 128u, Pk);
        const uint64_t xh = d0 + (xb >> 4), xl = xh + (xh_bytes >> 4);
        const uint64_t dh = d0 + (db >> 4);
        const uint32_t acc0 = it > 0 ? 1u : 0u;
#pragma unroll
        for (int ks = 0; ks < KC / 16; ++ks) {
          const uint32_t ko = (uint32_t)(ks * 16);
          umma_f16(tmem_base, xh + ko, dh + ko, idesc6, (ks > 0) ? 1u : acc0);
          umma_f16(tmem_base, xl + ko, dh + ko, idesc3, 1u);
        }
        umma_commit(s_empty + s);
      }
      __syncwarp();
    }
    if (elect_one()) umma_commit(s_done);
    __syncwarp();
  }
  // ---- drain, then rows (kw, ci) of the accumulator -> this CTA's partial ------------------------
  if (my_chunks > 0 && !timed_out) {
    if (!mbar_wait_bounded(s_done, 0u)) timed_out = true;
  }
  if (timed_out && a.err) atomicExch(a.err, 1);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  __syncthreads();
  float* dst = a.partial + (size_t)blockIdx.x * NW;
  if (warp >= 2) {
    // TMEM lane -> accumulator row: M = 128: lane = row; M = 64 (cute tmem_frg_1sm, M_MMA == 64): row m
    // lives in lane (m % 16) + 32 * (m / 16), i.e. 16 rows per 32-lane sub-partition
    const int row = MROWS == 128 ? (warp & 3) * 32 + lane : (lane < 16 ? (warp & 3) * 16 + lane : 1 << 20);
    const int kw = row / CP, ci = row - kw * CP;
    const uint32_t lane_addr = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
#pragma unroll 1
    for (int kh = 0; kh < 3; ++kh) {
#pragma unroll
      for (int hc = 0; hc < COUT / 16; ++hc) {
        float v[16], v2[16];
        tmem_ld<16>(lane_addr + (uint32_t)(kh * COUT + hc * 16), v);                 // x * hi(dy)
        tmem_ld<16>(lane_addr + (uint32_t)(3 * COUT + kh * COUT + hc * 16), v2);     // hi(x) * lo(dy)
        if (row < 3 * CP) {
#pragma unroll
          for (int e = 0; e < 16; ++e)
            dst[((size_t)(kh * 3 + kw) * CP + ci) * COUT + hc * 16 + e] = my_chunks > 0 ? v[e] + v2[e] : 0.f;
        } else if (row == 3 * CP && kh == 1) {           // ones row x the unshifted-by-rows dy copy
#pragma unroll
          for (int e = 0; e < 16; ++e) dst[9 * CP * COUT + hc * 16 + e] = my_chunks > 0 ? v[e] + v2[e] : 0.f;
        }
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TCOLS));
  }
}

template <int CP, int COUT, int KC>
static int launch_wgradp(int N, int H, int W, const void* x, const void* dy, float* dw, float* db, int* err,
                         WgradBatch* batch, cudaStream_t st) {
  const ConvGeom g = make_geom(N, H, W);
  const long long Lp = planes_positions(N, H, W);
  constexpr int G = CP / 8, GO = COUT / 8, XG = 3 * G + 1;
  const int CH = planes_chunk() < KC ? planes_chunk() : KC;    // K chunks start on multiples of KC
  if (KC % CH) return kPlanesTryNext;
  const int nchx = KC / CH;
  const size_t Pk = (size_t)KC * 16;
  const size_t stage = 2 * XG * Pk + 2 * 3 * GO * Pk;
  // the M = 128 MMA reads 16 row groups from each x half: groups past XG are junk rows (never
  // read back) whose addresses stay inside the stage (x lo is followed by the dy planes)
  const long long over = (long long)(16 - XG) * (long long)Pk - (long long)(6 * GO) * (long long)Pk;
  const size_t tail = (over > 0 ? (size_t)over : 0) + 256;
  int nb = kWpMaxStages;
  while (nb > 1 && 128 + nb * stage + tail > 227 * 1024) --nb;
  if (nb < 2) return kPlanesTryNext;
  const size_t smem = 128 + nb * stage + tail;
  WgradMaps tm;
  for (int k = 0; k < 3; ++k) {
    SEEDRL_TRY_RC(make_plane_map(&tm.x[k], x, Lp, 2 * G, g.PW + k, CH, nchx, G));               // kw = k
    SEEDRL_TRY_RC(make_plane_map(&tm.dy[k], dy, Lp, 2 * GO, (2 - k) * g.PW + 1, CH, nchx, GO));  // kh = k
  }
  static bool attr = false;
  if (!attr) {
    SEEDRL_CUDA(cudaFuncSetAttribute(wgradp_kernel<CP, COUT, KC>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     227 * 1024));
    attr = true;
  }
  static const bool dbg = getenv("SEEDRL_DEBUG_LAUNCH") != nullptr;
  if (dbg) fprintf(stderr, "wgradp<%d,%d,%d> CH=%d nb=%d stage=%zu smem=%zu\n", CP, COUT, KC, CH, nb, stage, smem);
  constexpr int NW = 9 * CP * COUT + COUT;
  WgradpArgs a;
  a.PW = g.PW; a.nchx = nchx; a.nb = nb; a.err = err; a.chunk = CH;
  a.nchunks = (int)((g.Q + g.PW + KC - 1) / KC);        // j = p + (kh - 1) * PW ranges over [0, Q + PW)
  const int grid = a.nchunks < kNumSMs ? a.nchunks : kNumSMs;
  if (!batch || batch->n >= kMaxReduceJobs || batch->used + (size_t)grid * NW > batch->cap_floats)
    return set_error(SEEDRL_ERR_INVALID_ARGUMENT, "wgradp: partial buffer too small");
  a.partial = batch->buf + batch->used;
  batch->used += (size_t)grid * NW;
  wgradp_kernel<CP, COUT, KC><<<grid, kWpThreads, smem, st>>>(tm, a);
  count_launch(PC_CONV_WGRAD, st);
  SEEDRL_CHECK_LAUNCH();
  batch->jobs[batch->n++] = ReduceJob{a.partial, dw, db, grid, 9 * CP * COUT, COUT};
  return SEEDRL_OK;
}

int wgradp(int cin, int cout, int N, int H, int W, const void* x, const void* dy, float* dw, float* db,
           int* err, WgradBatch* batch, cudaStream_t st) {
#define SEEDRL_WP_CASE(CI, CO_)                                                                          \
  if (cin == CI && cout == CO_) {                                                                        \
    int rc = launch_wgradp<CI, CO_, 256>(N, H, W, x, dy, dw, db, err, batch, st);                        \
    if (rc == kPlanesTryNext) rc = launch_wgradp<CI, CO_, 128>(N, H, W, x, dy, dw, db, err, batch, st);  \
    if (rc == kPlanesTryNext) rc = launch_wgradp<CI, CO_, 64>(N, H, W, x, dy, dw, db, err, batch, st);   \
    if (rc == kPlanesTryNext) return set_error(SEEDRL_ERR_INVALID_ARGUMENT, "wgradp: does not fit");     \
    return rc;                                                                                           \
  }
  SEEDRL_WP_CASE(16, 16)
  SEEDRL_WP_CASE(16, 32)
  SEEDRL_WP_CASE(32, 32)
#undef SEEDRL_WP_CASE
  return set_error(SEEDRL_ERR_INVALID_ARGUMENT, "wgradp: unsupported (cin,cout)");
}

// ------------------------------------------------------------------------------------------------
// format kernels (elementwise, HBM-bound; thread = one 16-byte unit = position x 8 channels)
__global__ void to_planes_kernel(ConvGeom g, int Lp, int G, int relu, const float* __restrict__ x,
                                 uint4* __restrict__ out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)Lp * G) return;
  const int s = (int)(i / G), go = (int)(i - (long long)s * G);     // group fastest: a warp reads whole pixels
  const int pix = in_pixel(g, s);
  float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
  if (pix >= 0) {
    const float4* src = reinterpret_cast<const float4*>(x + (size_t)pix * (G * 8) + go * 8);
    a = __ldg(src); b = __ldg(src + 1);
    if (relu) {
      a.x = fmaxf(a.x, 0.f); a.y = fmaxf(a.y, 0.f); a.z = fmaxf(a.z, 0.f); a.w = fmaxf(a.w, 0.f);
      b.x = fmaxf(b.x, 0.f); b.y = fmaxf(b.y, 0.f); b.z = fmaxf(b.z, 0.f); b.w = fmaxf(b.w, 0.f);
    }
  }
  out[(size_t)go * Lp + s] = pack8_bf16(a, b);
  out[(size_t)(G + go) * Lp + s] = pack8_bf16(bf16_resid4(a), bf16_resid4(b));
}

__global__ void from_planes_kernel(ConvGeom g, int Lp, int G, const uint4* __restrict__ in, float* __restrict__ y) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)Lp * G) return;
  const int s = (int)(i / G), go = (int)(i - (long long)s * G);     // group fastest: a warp reads whole pixels
  const int pix = in_pixel(g, s);
  if (pix < 0) return;
  const uint4 h = __ldg(in + (size_t)go * Lp + s), l = __ldg(in + (size_t)(G + go) * Lp + s);
  float4* dst = reinterpret_cast<float4*>(y + (size_t)pix * (G * 8) + go * 8);
  dst[0] = make_float4(bf16lo(h.x) + bf16lo(l.x), bf16hi(h.x) + bf16hi(l.x), bf16lo(h.y) + bf16lo(l.y),
                       bf16hi(h.y) + bf16hi(l.y));
  dst[1] = make_float4(bf16lo(h.z) + bf16lo(l.z), bf16hi(h.z) + bf16hi(l.z), bf16lo(h.w) + bf16lo(l.w),
                       bf16hi(h.w) + bf16hi(l.w));
}

int to_planes(int N, int H, int W, int C, int relu, const float* x, void* out, cudaStream_t st) {
  const ConvGeom g = make_geom(N, H, W);
  const long long Lp = planes_positions(N, H, W);
  const long long n = Lp * (C / 8);
  to_planes_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(g, (int)Lp, C / 8, relu, x,
                                                                reinterpret_cast<uint4*>(out));
  count_launch(PC_MISC, st);
  SEEDRL_CHECK_LAUNCH();
  return SEEDRL_OK;
}
int from_planes(int N, int H, int W, int C, const void* in, float* y, cudaStream_t st) {
  const ConvGeom g = make_geom(N, H, W);
  const long long Lp = planes_positions(N, H, W);
  const long long n = Lp * (C / 8);
  from_planes_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(g, (int)Lp, C / 8,
                                                                  reinterpret_cast<const uint4*>(in), y);
  count_launch(PC_MISC, st);
  SEEDRL_CHECK_LAUNCH();
  return SEEDRL_OK;
}

// Max-pool 3x3 / stride 2, TF 'SAME' (dmlab/networks.py:36-37): fp32 NHWC in -> plane tensors out
// (raw and ReLU'd: the two consumers of the pooled activation, networks.py:52-58) + argmax taps.
__global__ void poolp_fwd_kernel(ConvGeom go_, int Lp, int G, int H, int W, int pt, int pl,
                                 const float* __restrict__ x, uint4* __restrict__ out_raw,
                                 uint4* __restrict__ out_relu, uint8_t* __restrict__ idx) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)Lp * G) return;
  const int s = (int)(i / G), gq = (int)(i - (long long)s * G);
  const int pix = in_pixel(go_, s);           // pooled pixel (n * Ho + ho) * Wo + wo
  float best[8];
  unsigned char arg[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) { best[e] = 0.f; arg[e] = 0; }
  if (pix >= 0) {
    const int Wo = go_.W, Ho = go_.H, C = G * 8;
    const int n = pix / (Ho * Wo), r = pix - n * (Ho * Wo), ho = r / Wo, wo = r - ho * Wo;
#pragma unroll
    for (int e = 0; e < 8; ++e) best[e] = -INFINITY;
#pragma unroll
    for (int kh = 0; kh < 3; ++kh) {
      const int h = ho * 2 - pt + kh;
      if (h < 0 || h >= H) continue;
#pragma unroll
      for (int kw = 0; kw < 3; ++kw) {
        const int w = wo * 2 - pl + kw;
        if (w < 0 || w >= W) continue;
        const float4* src = reinterpret_cast<const float4*>(x + ((size_t)(n * H + h) * W + w) * C + gq * 8);
        const float4 a = __ldg(src), b = __ldg(src + 1);
        const float v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
        const unsigned char t = (unsigned char)(kh * 3 + kw);
#pragma unroll
        for (int e = 0; e < 8; ++e)
          if (v[e] > best[e]) { best[e] = v[e]; arg[e] = t; }
      }
    }
    uint2 packed;
    packed.x = arg[0] | (arg[1] << 8) | (arg[2] << 16) | ((uint32_t)arg[3] << 24);
    packed.y = arg[4] | (arg[5] << 8) | (arg[6] << 16) | ((uint32_t)arg[7] << 24);
    *reinterpret_cast<uint2*>(idx + (size_t)pix * C + gq * 8) = packed;
  }
  const float4 a = make_float4(best[0], best[1], best[2], best[3]), b = make_float4(best[4], best[5], best[6], best[7]);
  out_raw[(size_t)gq * Lp + s] = pack8_bf16(a, b);
  out_raw[(size_t)(G + gq) * Lp + s] = pack8_bf16(bf16_resid4(a), bf16_resid4(b));
  const float4 ra = make_float4(fmaxf(a.x, 0.f), fmaxf(a.y, 0.f), fmaxf(a.z, 0.f), fmaxf(a.w, 0.f));
  const float4 rb = make_float4(fmaxf(b.x, 0.f), fmaxf(b.y, 0.f), fmaxf(b.z, 0.f), fmaxf(b.w, 0.f));
  out_relu[(size_t)gq * Lp + s] = pack8_bf16(ra, rb);
  out_relu[(size_t)(G + gq) * Lp + s] = pack8_bf16(bf16_resid4(ra), bf16_resid4(rb));
}

// dx[n,h,w,c] = sum over the <= 4 windows containing (h, w) whose argmax is (h, w); dy is a plane
// tensor at the pooled resolution, dx either a plane tensor (full resolution) or fp32 NHWC.
template <bool OUT_PLANES>
__global__ void poolp_bwd_kernel(ConvGeom gf, int Lpf, ConvGeom gp, int Lpp, int G, int pt, int pl,
                                 const uint4* __restrict__ dy, const uint8_t* __restrict__ idx,
                                 uint4* __restrict__ dx_planes, float* __restrict__ dx_nhwc) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int H = gf.H, W = gf.W, Ho = gp.H, Wo = gp.W, C = G * 8;
  int gq, s = 0, pix;
  if (OUT_PLANES) {
    if (i >= (long long)Lpf * G) return;
    s = (int)(i / G); gq = (int)(i - (long long)s * G);
    pix = in_pixel(gf, s);
  } else {
    if (i >= (long long)gf.N * H * W * G) return;
    pix = (int)(i / G); gq = (int)(i - (long long)pix * G);
  }
  float acc[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) acc[e] = 0.f;
  if (pix >= 0) {
    const int n = pix / (H * W), r = pix - n * (H * W), h = r / W, w = r - h * W;
    const int hp = h + pt, wp = w + pl;
#pragma unroll
    for (int dh = 0; dh < 2; ++dh) {
      const int ho = (hp >> 1) - dh;
      const int kh = hp - 2 * ho;
      if (ho < 0 || ho >= Ho || kh > 2) continue;
#pragma unroll
      for (int dw = 0; dw < 2; ++dw) {
        const int wo = (wp >> 1) - dw;
        const int kw = wp - 2 * wo;
        if (wo < 0 || wo >= Wo || kw > 2) continue;
        const unsigned tap = (unsigned)(kh * 3 + kw);
        const size_t o = ((size_t)(n * Ho + ho) * Wo + wo) * C + gq * 8;
        const uint2 t = __ldg(reinterpret_cast<const uint2*>(idx + o));
        const int sp = (n * gp.RH + ho + 1) * gp.PW + wo + 1;
        const uint4 hh = __ldg(dy + (size_t)gq * Lpp + sp), ll = __ldg(dy + (size_t)(G + gq) * Lpp + sp);
        const float gv[8] = {bf16lo(hh.x) + bf16lo(ll.x), bf16hi(hh.x) + bf16hi(ll.x), bf16lo(hh.y) + bf16lo(ll.y),
                             bf16hi(hh.y) + bf16hi(ll.y), bf16lo(hh.z) + bf16lo(ll.z), bf16hi(hh.z) + bf16hi(ll.z),
                             bf16lo(hh.w) + bf16lo(ll.w), bf16hi(hh.w) + bf16hi(ll.w)};
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const unsigned te = ((e < 4 ? t.x : t.y) >> (8 * (e & 3))) & 0xFFu;
          if (te == tap) acc[e] += gv[e];
        }
      }
    }
  }
  const float4 a = make_float4(acc[0], acc[1], acc[2], acc[3]), b = make_float4(acc[4], acc[5], acc[6], acc[7]);
  if (OUT_PLANES) {
    dx_planes[(size_t)gq * Lpf + s] = pack8_bf16(a, b);
    dx_planes[(size_t)(G + gq) * Lpf + s] = pack8_bf16(bf16_resid4(a), bf16_resid4(b));
  } else {
    float4* dst = reinterpret_cast<float4*>(dx_nhwc + (size_t)pix * C + gq * 8);
    dst[0] = a; dst[1] = b;
  }
}

static void same_pad3s2(int in, int* out, int* before) {
  *out = (in + 1) / 2;
  const int total = (*out - 1) * 2 + 3 - in;
  *before = total > 0 ? total / 2 : 0;
}

int poolp_forward(int N, int H, int W, int C, const float* x, void* out_raw, void* out_relu, uint8_t* idx,
                  cudaStream_t st) {
  int Ho, Wo, pt, pl;
  same_pad3s2(H, &Ho, &pt);
  same_pad3s2(W, &Wo, &pl);
  const ConvGeom g = make_geom(N, Ho, Wo);
  const long long Lp = planes_positions(N, Ho, Wo);
  const long long n = Lp * (C / 8);
  poolp_fwd_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(g, (int)Lp, C / 8, H, W, pt, pl, x,
                                                                reinterpret_cast<uint4*>(out_raw),
                                                                reinterpret_cast<uint4*>(out_relu), idx);
  count_launch(PC_POOL, st);
  SEEDRL_CHECK_LAUNCH();
  return SEEDRL_OK;
}

int poolp_backward(int N, int H, int W, int C, const void* dy, const uint8_t* idx, void* dx_planes,
                   float* dx_nhwc, cudaStream_t st) {
  int Ho, Wo, pt, pl;
  same_pad3s2(H, &Ho, &pt);
  same_pad3s2(W, &Wo, &pl);
  const ConvGeom gf = make_geom(N, H, W), gp = make_geom(N, Ho, Wo);
  const long long Lpf = planes_positions(N, H, W), Lpp = planes_positions(N, Ho, Wo);
  if (dx_planes) {
    const long long n = Lpf * (C / 8);
    poolp_bwd_kernel<true><<<(unsigned)((n + 255) / 256), 256, 0, st>>>(
        gf, (int)Lpf, gp, (int)Lpp, C / 8, pt, pl, reinterpret_cast<const uint4*>(dy), idx,
        reinterpret_cast<uint4*>(dx_planes), nullptr);
  } else {
    const long long n = (long long)N * H * W * (C / 8);
    poolp_bwd_kernel<false><<<(unsigned)((n + 255) / 256), 256, 0, st>>>(
        gf, (int)Lpf, gp, (int)Lpp, C / 8, pt, pl, reinterpret_cast<const uint4*>(dy), idx, nullptr, dx_nhwc);
  }
  count_launch(PC_POOL, st);
  SEEDRL_CHECK_LAUNCH();
  return SEEDRL_OK;
}

}  // namespace seedrl
