// Dense-layer GEMMs on the 5th-gen tensor cores (tcgen05 + TMEM): the Linear contractions of
// dmlab/networks.py:105-118,157-169 (Dense(256), LSTM input projection, their data and weight
// gradients) with fp32 storage, bf16 (or bf16x3) operands and fp32 accumulation.
//
//   C[M,N] (=|+=) op(A)[M,K] * op(B)[K,N]      row-major fp32, leading dims lda/ldb/ldc
//   TA: A is stored [K,M] (C = A^T B).   TB: B is stored [N,K] (C = A B^T).
//
// Either storage order of either operand is ALREADY a canonical no-swizzle UMMA layout once
// 8 contiguous elements are packed into one 16-byte bf16 unit:
//   contiguous along K  -> K-major :  planes [K/8][rows][8 k],  LBO = plane stride, SBO = 128 B
//   contiguous along MN -> MN-major:  planes [MN/8][k][8 mn],   LBO = 128 B, SBO = plane stride
// so there is no transpose anywhere: TA / TB only flip the major-ness bits of the instruction
// descriptor.  A CTA owns a 128 x BN tile of C (UMMA M = 128, N = BN <= 256) and a slice of K
// (split-K over blockIdx.z); K is walked in blocks of 64 through two shared-memory stages:
// all 8 warps convert fp32 global -> bf16 units of stage s while the tensor core works on
// stage s^1 (its MMAs were committed to that stage's mbarrier).  Split = bf16x3: every
// operand is staged as hi and lo planes and each K-step issues hi*hi + lo*hi + hi*lo.
// Epilogue: tcgen05.ld (thread = one row, 16 columns at a time) -> bias / relu / mask /
// accumulate -> C, or -> the split-K workspace, reduced in slice order by
// gemm_tc_reduce_kernel (deterministic).
#include <cstdlib>

#include "kernels.h"
#include "tc_common.cuh"

namespace seedrl {

constexpr int kGtThreads = 256;
constexpr int kGtBM = 128;
// K elements per staged block = template parameter BK (64, 32 or 16): a smaller block shrinks the
// stage so that several CTAs fit an SM and overlap each other's load / convert / MMA / epilogue phases
// (measured on the learner's nine GEMM shapes: 322 us at BK = 64, 259 us at BK = 32)

struct GemmTcParams {
  int M, N, K;
  const float* A; int lda;
  const float* B; int ldb;
  float* C; int ldc;            // final output (splits == 1) ...
  float* ws;                    // ... or split-K partials [splits][M][N]
  int BN;                       // tile width: multiple of 16, <= 256
  int kblocks_per_split;        // K blocks (of 64) per blockIdx.z
  int vecA, vecB;               // 16-byte aligned rows: float4 loads
  GemmEpi e;
  int* error_flag;
  ConvGather cg;                // GATHER kernels: op(A) = im2col(cg.x)
};

// 8 consecutive elements along the contiguous direction of a row-major matrix (raw fp32, two
// float4).  `row` / `col0` are global coordinates; rows/cols outside [0,R) x [0,Cn) read as zero.
struct RawUnit { float4 a, b; };
__device__ __forceinline__ RawUnit load_raw(const float* __restrict__ P, int ld, int row, int col0, int R, int Cn,
                                            bool vec) {
  RawUnit r;
  r.a = make_float4(0.f, 0.f, 0.f, 0.f);
  r.b = r.a;
  if (row < R && col0 < Cn) {
    const float* src = P + (size_t)row * ld + col0;
    if (vec && col0 + 8 <= Cn) {
      r.a = __ldg(reinterpret_cast<const float4*>(src));
      r.b = __ldg(reinterpret_cast<const float4*>(src) + 1);
    } else {
      float v[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = col0 + j < Cn ? __ldg(src + j) : 0.f;
      r.a = make_float4(v[0], v[1], v[2], v[3]);
      r.b = make_float4(v[4], v[5], v[6], v[7]);
    }
  }
  return r;
}
// 8 consecutive columns (kh, kw, c) of row `pos` = (n, ho, wo) of the im2col matrix, read from the
// NHWC tensor itself: x[n][ho*S + kh][wo*S ...][...] is contiguous over (kw, c) for a fixed kh.
// kc0 % 8 == 0 and KC % 8 == 0 (host-checked), so a unit never straddles two kernel rows.
__device__ __forceinline__ RawUnit load_gather(const ConvGather& g, int pos, int kc0, int R, int Cn) {
  RawUnit r;
  r.a = make_float4(0.f, 0.f, 0.f, 0.f);
  r.b = r.a;
  if (pos < R && kc0 < Cn) {
    const unsigned int t = fast_div((unsigned int)pos, g.wo_mul, g.wo_sh);
    const int wo = pos - (int)t * g.Wo;
    const unsigned int n = fast_div(t, g.ho_mul, g.ho_sh);
    const int ho = (int)t - (int)n * g.Ho;
    const int kh = (int)fast_div((unsigned int)kc0, g.kc_mul, g.kc_sh);
    const int rem = kc0 - kh * g.KC;
    const size_t off = (((size_t)n * g.H + (size_t)(ho * g.S + kh)) * g.W + (size_t)(wo * g.S)) * g.C + rem;
    if (g.u8) {
      const uint2 v = __ldg(reinterpret_cast<const uint2*>(reinterpret_cast<const uint8_t*>(g.x) + off));
      const float k = 1.0f / 255.0f;
      r.a = make_float4((float)(v.x & 0xffu) * k, (float)((v.x >> 8) & 0xffu) * k,
                        (float)((v.x >> 16) & 0xffu) * k, (float)(v.x >> 24) * k);
      r.b = make_float4((float)(v.y & 0xffu) * k, (float)((v.y >> 8) & 0xffu) * k,
                        (float)((v.y >> 16) & 0xffu) * k, (float)(v.y >> 24) * k);
    } else {
      const float4* src = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(g.x) + off);
      r.a = __ldg(src);
      r.b = __ldg(src + 1);
    }
  }
  return r;
}
// -> one bf16x8 unit (+ the residual unit for bf16x3)
template <bool SPLIT>
__device__ __forceinline__ void store_unit(RawUnit r, bool relu, uint4* hi_dst, uint4* lo_dst) {
  if (relu) {
    r.a.x = fmaxf(r.a.x, 0.f); r.a.y = fmaxf(r.a.y, 0.f); r.a.z = fmaxf(r.a.z, 0.f); r.a.w = fmaxf(r.a.w, 0.f);
    r.b.x = fmaxf(r.b.x, 0.f); r.b.y = fmaxf(r.b.y, 0.f); r.b.z = fmaxf(r.b.z, 0.f); r.b.w = fmaxf(r.b.w, 0.f);
  }
  *hi_dst = pack8_bf16(r.a, r.b);
  if (SPLIT) *lo_dst = pack8_bf16(bf16_resid4(r.a), bf16_resid4(r.b));
}

template <bool TA, bool TB, bool SPLIT, int kGtBK, bool GATHER>
__global__ void __launch_bounds__(kGtThreads)
gemm_tc_kernel(const GemmTcParams p) {
  constexpr int S = SPLIT ? 2 : 1;
  constexpr int KG = kGtBK / 8;                 // 16-byte K-groups per row of a block
  extern __shared__ __align__(128) uint8_t smem_raw[];
  const int BN = p.BN;
  // one stage: [A hi | A lo | B hi | B lo].  Plane strides are padded by one 16-byte unit so
  // that the 8 lanes of a store phase (8 consecutive planes, same row) hit 8 different banks.
  constexpr uint32_t a_ps = TA ? (kGtBK + 1) : (kGtBM + 1);            // A plane stride (units)
  constexpr uint32_t a_units = TA ? (kGtBM / 8) * a_ps : (kGtBK / 8) * a_ps;
  const uint32_t b_ps = TB ? (uint32_t)BN + 1 : (uint32_t)kGtBK + 1;    // B plane stride (units)
  const uint32_t b_units = TB ? (kGtBK / 8) * b_ps : (uint32_t)(BN / 8) * b_ps;
  const int b_items = BN * kGtBK / 8;                                  // 16-byte units actually staged
  const uint32_t stage_units = S * (a_units + b_units);
  uint4* s_buf = reinterpret_cast<uint4*>(smem_raw);
  uint64_t* s_bar = reinterpret_cast<uint64_t*>(smem_raw + (size_t)2 * stage_units * 16);   // [2] stage free
  uint64_t* s_done = s_bar + 2;
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(s_done + 1);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int m0 = blockIdx.y * kGtBM, n0 = blockIdx.x * BN;
  const int kb0 = blockIdx.z * p.kblocks_per_split;
  const int nkb_total = (p.K + kGtBK - 1) / kGtBK;
  const int nkb = min(p.kblocks_per_split, nkb_total - kb0);
  const uint32_t tcols = BN <= 32 ? 32u : (BN <= 64 ? 64u : (BN <= 128 ? 128u : 256u));

  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(s_bar)));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(s_bar + 1)));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(s_done)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(s_tmem)),
                 "r"(tcols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(s_tmem);
  // instruction descriptor: bf16 x bf16 -> f32, M = 128, N = BN; bit 15 / 16: A / B MN-major
  const uint32_t idesc = umma_idesc(kGtBM, BN) | (TA ? (1u << 15) : 0u) | (TB ? 0u : (1u << 16));
  bool ok = true;

  for (int kb = 0; kb < nkb; ++kb) {
    const int st = kb & 1;
    const int k0 = (kb0 + kb) * kGtBK;
    uint4* sA = s_buf + (size_t)st * stage_units;
    uint4* sB = sA + (size_t)S * a_units;
    // ---- stage the K-block: every global load is issued before the first conversion -------
    // consecutive threads take consecutive 8-element groups of the SAME row (coalesced).
    constexpr int AI = (kGtBM * kGtBK / 8) / kGtThreads;   // 4 A units per thread
    constexpr int BI = (256 * kGtBK / 8) / kGtThreads;     // <= 8 B units per thread
    RawUnit ra[AI], rb[BI];
#pragma unroll
    for (int r = 0; r < AI; ++r) {
      const int u = tid + r * kGtThreads;
      if (GATHER) {           // rows of the im2col matrix are positions: M of the forward, K of the weight gradient
        if (!TA) ra[r] = load_gather(p.cg, m0 + u / KG, k0 + (u % KG) * 8, p.M, p.K);
        else     ra[r] = load_gather(p.cg, k0 + (u >> 4), m0 + (u & 15) * 8, p.K, p.M);
      } else {
        if (!TA) ra[r] = load_raw(p.A, p.lda, m0 + u / KG, k0 + (u % KG) * 8, p.M, p.K, p.vecA != 0);   // A[M,K]
        else     ra[r] = load_raw(p.A, p.lda, k0 + (u >> 4), m0 + (u & 15) * 8, p.K, p.M, p.vecA != 0);  // A stored [K,M]
      }
    }
    const int bng = BN >> 3;
#pragma unroll
    for (int r = 0; r < BI; ++r) {
      const int u = tid + r * kGtThreads;
      if (u < b_items) {
        if (TB) rb[r] = load_raw(p.B, p.ldb, n0 + u / KG, k0 + (u % KG) * 8, p.N, p.K, p.vecB != 0);  // B stored [N,K]
        else    rb[r] = load_raw(p.B, p.ldb, k0 + u / bng, n0 + (u % bng) * 8, p.K, p.N, p.vecB != 0); // B[K,N]
      }
    }
    // the MMAs that read this stage two K-blocks ago have completed
    if (kb >= 2) ok = mbar_wait_bounded(s_bar + st, (uint32_t)(((kb >> 1) - 1) & 1)) && ok;
#pragma unroll
    for (int r = 0; r < AI; ++r) {
      const int u = tid + r * kGtThreads;
      // K-major planes [8 kg][128 m] / MN-major planes [16 mg][64 k]
      const uint32_t o = !TA ? (uint32_t)(u % KG) * a_ps + u / KG : (uint32_t)(u & 15) * a_ps + (u >> 4);
      store_unit<SPLIT>(ra[r], p.e.a_relu != 0, sA + o, sA + a_units + o);
    }
#pragma unroll
    for (int r = 0; r < BI; ++r) {
      const int u = tid + r * kGtThreads;
      if (u < b_items) {
        // K-major planes [8 kg][BN n] / MN-major planes [BN/8 ng][64 k]
        const uint32_t o = TB ? (uint32_t)(u % KG) * b_ps + u / KG : (uint32_t)(u % bng) * b_ps + u / bng;
        store_unit<SPLIT>(rb[r], false, sB + o, sB + b_units + o);
      }
    }
    // generic-proxy smem writes -> visible to the tensor core's async proxy
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (warp == 0 && elect_one()) {
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t a_addr = smem_u32(sA), b_addr = smem_u32(sB);
#pragma unroll
      for (int ks = 0; ks < kGtBK / 16; ++ks) {
        // K-major: two K-groups = two planes (LBO = plane stride);  MN-major: 16 k-rows = 256 B
        const uint64_t da = TA ? umma_desc(a_addr + ks * 256u, 128u, a_ps * 16u)
                               : umma_desc(a_addr + ks * 2u * a_ps * 16u, a_ps * 16u, 128u);
        const uint64_t db = TB ? umma_desc(b_addr + ks * 2u * b_ps * 16u, b_ps * 16u, 128u)
                               : umma_desc(b_addr + ks * 256u, 128u, b_ps * 16u);
        const uint32_t acc = (kb > 0 || ks > 0) ? 1u : 0u;
        umma_f16(tmem_base, da, db, idesc, acc);
        if (SPLIT) {   // the address field counts 16-byte units
          umma_f16(tmem_base, da + a_units, db, idesc, 1u);
          umma_f16(tmem_base, da, db + b_units, idesc, 1u);
        }
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                       smem_u32(s_bar + st))
                   : "memory");
      if (kb == nkb - 1)
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                         smem_u32(s_done))
                     : "memory");
    }
  }
  if (nkb > 0) ok = mbar_wait_bounded(s_done, 0u) && ok;
  if (!ok && p.error_flag) atomicExch(p.error_flag, 1);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

  // ---- epilogue: TMEM -> registers (thread = row) -> per-warp 32x32 transpose in shared memory
  // -> row-contiguous 128-byte global accesses (a thread-per-row store would touch 32 different
  // sectors per instruction: measured 17 us per 128x256 tile).  The operand stages are free now.
  __syncthreads();
  {
    const int q = warp & 3, half = warp >> 2;          // TMEM lane quadrant, column-block parity
    float* scratch = reinterpret_cast<float*>(smem_raw) + warp * (32 * 33);
    const bool partial = p.ws != nullptr;
    const int row0 = m0 + q * 32;
    for (int cb = half; cb * 32 < BN; cb += 2) {
      float v[32];
      if (nkb > 0) {
        tmem_ld<32>(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(cb * 32), v);
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = 0.f;
      }
#pragma unroll
      for (int j = 0; j < 32; ++j) scratch[lane * 33 + j] = v[j];
      __syncwarp();
      const int gn = n0 + cb * 32 + lane;
      const bool col_ok = gn < p.N && cb * 32 + lane < BN;
      const float bias = (!partial && p.e.bias && col_ok) ? __ldg(p.e.bias + gn) : 0.f;
#pragma unroll 4
      for (int r = 0; r < 32; ++r) {
        const int gm = row0 + r;
        if (gm < p.M && col_ok) {
          float x = scratch[r * 33 + lane];
          if (partial) {
            p.ws[((size_t)blockIdx.z * p.M + gm) * p.N + gn] = x;
          } else {
            x += bias;
            if (p.e.relu) x = fmaxf(x, 0.f);
            if (p.e.mask) x = __ldg(p.e.mask + (size_t)gm * p.e.ldm + gn) > 0.f ? x : 0.f;
            float* c = p.C + (size_t)gm * p.ldc + gn;
            *c = p.e.accumulate ? *c + x : x;
          }
        }
      }
      __syncwarp();
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tcols));
}

// C = epilogue(sum_z ws[z]) in slice order.  VEC: four columns per thread (N % 4 == 0; the
// partial planes are then 16-byte aligned), all `splits` loads of a thread independent.
template <bool VEC>
__global__ void gemm_tc_reduce_kernel(int M, int N, int splits, const float* __restrict__ ws,
                                      float* __restrict__ C, int ldc, GemmEpi e) {
  constexpr int W = VEC ? 4 : 1;
  const int i = (blockIdx.x * blockDim.x + threadIdx.x) * W;
  if (i >= M * N) return;
  const int m = i / N, n = i - m * N;
  float x[W];
#pragma unroll
  for (int j = 0; j < W; ++j) x[j] = 0.f;
  const size_t plane = (size_t)M * N;
#pragma unroll 4
  for (int z = 0; z < splits; ++z) {
    if (VEC) {
      const float4 v = __ldcs(reinterpret_cast<const float4*>(ws + (size_t)z * plane + i));
      x[0] += v.x; x[W > 1 ? 1 : 0] += v.y; x[W > 2 ? 2 : 0] += v.z; x[W > 3 ? 3 : 0] += v.w;
    } else {
      x[0] += __ldcs(ws + (size_t)z * plane + i);
    }
  }
#pragma unroll
  for (int j = 0; j < W; ++j) {
    float v = x[j];
    if (e.bias) v += __ldg(e.bias + n + j);
    if (e.relu) v = fmaxf(v, 0.f);
    if (e.mask) v = __ldg(e.mask + (size_t)m * e.ldm + n + j) > 0.f ? v : 0.f;
    float* c = C + (size_t)m * ldc + n + j;
    *c = e.accumulate ? *c + v : v;
  }
}

// Same reduction for many slices (the im2col weight gradients: 64-128 slices of a small M x N): the slices
// of an output are spread over 8 lanes (lane zl sums z = zl, zl + 8, ... in order), the 8 partial sums are
// combined in lane order through shared memory => still a fixed summation order.  CTA = 32 float4 outputs.
__global__ void __launch_bounds__(256)
gemm_tc_reduce_wide_kernel(int M, int N, int splits, const float* __restrict__ ws, float* __restrict__ C, int ldc,
                           GemmEpi e) {
  __shared__ float4 part[8][32];
  const int o = threadIdx.x & 31, zl = threadIdx.x >> 5;
  const int i = (blockIdx.x * 32 + o) * 4;
  const bool in = i < M * N;
  const size_t plane = (size_t)M * N;
  float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
  if (in) {
    int z = zl;
    for (; z + 8 < splits; z += 16) {
      const float4 v = __ldcs(reinterpret_cast<const float4*>(ws + (size_t)z * plane + i));
      const float4 w = __ldcs(reinterpret_cast<const float4*>(ws + (size_t)(z + 8) * plane + i));
      a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
      b.x += w.x; b.y += w.y; b.z += w.z; b.w += w.w;
    }
    if (z < splits) {
      const float4 v = __ldcs(reinterpret_cast<const float4*>(ws + (size_t)z * plane + i));
      a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
    }
  }
  part[zl][o] = make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w);
  __syncthreads();
  if (zl != 0 || !in) return;
  float x[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const float4 v = part[k][o];
    x[0] += v.x; x[1] += v.y; x[2] += v.z; x[3] += v.w;
  }
  const int m = i / N, n = i - m * N;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    float v = x[j];
    if (e.bias) v += __ldg(e.bias + n + j);
    if (e.relu) v = fmaxf(v, 0.f);
    if (e.mask) v = __ldg(e.mask + (size_t)m * e.ldm + n + j) > 0.f ? v : 0.f;
    float* c = C + (size_t)m * ldc + n + j;
    *c = e.accumulate ? *c + v : v;
  }
}

static int norm_bk(int bk) { return bk == 64 ? 64 : (bk == 16 ? 16 : 32); }
static int g_gemm_bk = norm_bk(getenv("SEEDRL_GEMM_BK") ? atoi(getenv("SEEDRL_GEMM_BK")) : 32);
void gemm_tc_set_bk(int bk) { g_gemm_bk = norm_bk(bk); }

bool gemm_tc_supported(int M, int N, int K) { return M >= 64 && N >= 16 && K >= 32; }

size_t gemm_tc_workspace_bytes() { return (size_t)48 << 20; }

static int g_gemm_gather = getenv("SEEDRL_GEMM_GATHER") ? atoi(getenv("SEEDRL_GEMM_GATHER")) : 1;
void gemm_tc_set_gather(int on) { g_gemm_gather = on; }
bool gemm_tc_gather_enabled() { return g_gemm_gather != 0; }

bool conv_gather_setup(const void* x, int u8, int N, int H, int W, int C, int K, int S, ConvGather* g) {
  const int Ho = (H - K) / S + 1, Wo = (W - K) / S + 1, KC = K * C;
  if (Ho < 2 || Wo < 2 || KC < 8 || (KC & 7)) return false;
  if ((long long)N * Ho * Wo >= (1ll << 31)) return false;                 // fast_div domain
  if (u8 ? (((W * C) & 7) || ((S * C) & 7) || (reinterpret_cast<uintptr_t>(x) & 7))
         : ((C & 3) || (reinterpret_cast<uintptr_t>(x) & 15)))
    return false;                                                            // 8-byte / 16-byte loads
  g->x = x; g->u8 = u8; g->H = H; g->W = W; g->C = C; g->S = S; g->Ho = Ho; g->Wo = Wo; g->KC = KC;
  fast_div_setup((unsigned int)Wo, &g->wo_mul, &g->wo_sh);
  fast_div_setup((unsigned int)Ho, &g->ho_mul, &g->ho_sh);
  fast_div_setup((unsigned int)KC, &g->kc_mul, &g->kc_sh);
  return true;
}

int gemm_tc(bool ta, bool tb, int split, int M, int N, int K, const float* A, int lda, const float* B,
            int ldb, float* C, int ldc, const GemmEpi& e, float* ws, size_t ws_bytes, int* err,
            cudaStream_t st, const ConvGather* cg) {
  if (M <= 0 || N <= 0) return SEEDRL_OK;
  SEEDRL_CHECK_ARG(!cg || (!tb && ((ta ? M : K) & 7) == 0), "gathered operand: tb or a ragged kernel row");
  GemmTcParams p;
  if (cg) p.cg = *cg; else p.cg = ConvGather{};
  p.M = M; p.N = N; p.K = K; p.A = A; p.lda = lda; p.B = B; p.ldb = ldb; p.C = C; p.ldc = ldc;
  p.e = e; p.error_flag = err;
  const int n16 = ((N + 15) / 16) * 16;
  const int BK = g_gemm_bk;
  int bn = split ? 128 : 256;                           // shared memory: 2 stages x (hi + lo)
  if (bn > n16) bn = n16;
  // narrower tiles until the grid can cover the SMs (with split-K below)
  const int nkb_all = ceil_div(K, BK);
  const int kb256 = 256 / BK;                           // K-blocks per 256 elements of K
  while (bn > 64 && ceil_div(M, kGtBM) * ceil_div(N, bn) * (nkb_all >= kb256 ? nkb_all / (kb256 / 2) : 1) < kNumSMs)
    bn >>= 1;
  p.BN = ((bn + 15) / 16) * 16;
  auto al16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
  p.vecA = al16(A) && (lda & 3) == 0;
  p.vecB = al16(B) && (ldb & 3) == 0;
  const int tiles = ceil_div(M, kGtBM) * ceil_div(N, p.BN);
  const int nkb = ceil_div(K, BK);
  // split-K until the grid covers the SMs, keeping >= 128 elements of K per slice and the
  // partials inside the workspace
  int splits = 1;
  static const int waves = getenv("SEEDRL_GEMM_WAVES") ? atoi(getenv("SEEDRL_GEMM_WAVES")) : 1;   // tuning knob
  while (tiles * splits < waves * kNumSMs && nkb / (splits * 2) >= 128 / BK &&
         (size_t)(splits * 2) * M * N * sizeof(float) <= ws_bytes && ws)
    splits *= 2;
  p.kblocks_per_split = ceil_div(nkb, splits);
  splits = ceil_div(nkb, p.kblocks_per_split);
  p.ws = splits > 1 ? ws : nullptr;
  const int S = split ? 2 : 1;
  const size_t a_un = ta ? (size_t)(kGtBM / 8) * (BK + 1) : (size_t)(BK / 8) * (kGtBM + 1);
  const size_t b_un = tb ? (size_t)(BK / 8) * (p.BN + 1) : (size_t)(p.BN / 8) * (BK + 1);
  size_t smem = (size_t)2 * S * (a_un + b_un) * 16 + 64;
  const size_t epi = (size_t)(kGtThreads / 32) * 32 * 33 * 4;      // the epilogue's transpose scratch aliases the stages
  if (smem < epi) smem = epi;
  dim3 grid(ceil_div(N, p.BN), ceil_div(M, kGtBM), splits);
#define SEEDRL_GT_LAUNCH2(TA_, TB_, SP_, BK_, G_)                                               \
  do {                                                                                          \
    static bool attr = false;                                                                   \
    if (!attr) {                                                                                \
      SEEDRL_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<TA_, TB_, SP_, BK_, G_>,                  \
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024)); \
      attr = true;                                                                              \
    }                                                                                           \
    gemm_tc_kernel<TA_, TB_, SP_, BK_, G_><<<grid, kGtThreads, smem, st>>>(p);                  \
  } while (0)
#define SEEDRL_GT_LAUNCH1(TA_, TB_, SP_, BK_)                                                   \
  do {                                                                                          \
    if (cg && !(TB_)) SEEDRL_GT_LAUNCH2(TA_, false, SP_, BK_, true);                            \
    else SEEDRL_GT_LAUNCH2(TA_, TB_, SP_, BK_, false);                                          \
  } while (0)
#define SEEDRL_GT_LAUNCH(TA_, TB_, SP_)                                                         \
  do {                                                                                          \
    if (BK == 32) SEEDRL_GT_LAUNCH1(TA_, TB_, SP_, 32);                                         \
    else if (BK == 16) SEEDRL_GT_LAUNCH1(TA_, TB_, SP_, 16);                                    \
    else SEEDRL_GT_LAUNCH1(TA_, TB_, SP_, 64);                                                  \
  } while (0)
  if (split) {
    if (!ta && !tb) SEEDRL_GT_LAUNCH(false, false, true);
    else if (ta && !tb) SEEDRL_GT_LAUNCH(true, false, true);
    else if (!ta && tb) SEEDRL_GT_LAUNCH(false, true, true);
    else SEEDRL_GT_LAUNCH(true, true, true);
  } else {
    if (!ta && !tb) SEEDRL_GT_LAUNCH(false, false, false);
    else if (ta && !tb) SEEDRL_GT_LAUNCH(true, false, false);
    else if (!ta && tb) SEEDRL_GT_LAUNCH(false, true, false);
    else SEEDRL_GT_LAUNCH(true, true, false);
  }
#undef SEEDRL_GT_LAUNCH
#undef SEEDRL_GT_LAUNCH1
#undef SEEDRL_GT_LAUNCH2
  count_launch(PC_GEMM, st);
  SEEDRL_CHECK_LAUNCH();
  if (splits > 1) {
    if ((N & 3) == 0 && splits >= 16)
      gemm_tc_reduce_wide_kernel<<<ceil_div(M * N / 4, 32), 256, 0, st>>>(M, N, splits, ws, C, ldc, e);
    else if ((N & 3) == 0)
      gemm_tc_reduce_kernel<true><<<ceil_div(M * N / 4, 256), 256, 0, st>>>(M, N, splits, ws, C, ldc, e);
    else
      gemm_tc_reduce_kernel<false><<<ceil_div(M * N, 256), 256, 0, st>>>(M, N, splits, ws, C, ldc, e);
    count_launch(PC_GEMM, st);
    SEEDRL_CHECK_LAUNCH();
  }
  return SEEDRL_OK;
}

}  // namespace seedrl
