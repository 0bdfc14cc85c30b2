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
#include "kernels.h"
#include "tc_common.cuh"

namespace seedrl {

constexpr int kGtThreads = 256;
constexpr int kGtBM = 128;
constexpr int kGtBK = 64;

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
};

// 8 consecutive elements along the contiguous direction of a row-major matrix -> one bf16x8
// unit (+ residual unit).  `row` / `col0` are global coordinates; rows/cols outside
// [0,R) x [0,Cn) read as zero.
template <bool SPLIT>
__device__ __forceinline__ void load_unit(const float* __restrict__ P, int ld, int row, int col0, int R, int Cn,
                                          bool vec, bool relu, uint4* hi, uint4* lo) {
  float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
  if (row < R) {
    const float* src = P + (size_t)row * ld + col0;
    if (vec && col0 + 8 <= Cn) {
      a = __ldg(reinterpret_cast<const float4*>(src));
      b = __ldg(reinterpret_cast<const float4*>(src) + 1);
    } else {
      float v[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = col0 + j < Cn ? __ldg(src + j) : 0.f;
      a = make_float4(v[0], v[1], v[2], v[3]);
      b = make_float4(v[4], v[5], v[6], v[7]);
    }
  }
  if (relu) {
    a.x = fmaxf(a.x, 0.f); a.y = fmaxf(a.y, 0.f); a.z = fmaxf(a.z, 0.f); a.w = fmaxf(a.w, 0.f);
    b.x = fmaxf(b.x, 0.f); b.y = fmaxf(b.y, 0.f); b.z = fmaxf(b.z, 0.f); b.w = fmaxf(b.w, 0.f);
  }
  *hi = pack8_bf16(a, b);
  if (SPLIT) *lo = pack8_bf16(bf16_resid4(a), bf16_resid4(b));
}

template <bool TA, bool TB, bool SPLIT>
__global__ void __launch_bounds__(kGtThreads)
gemm_tc_kernel(const GemmTcParams p) {
  constexpr int S = SPLIT ? 2 : 1;
  extern __shared__ __align__(128) uint8_t smem_raw[];
  const int BN = p.BN;
  // one stage: [A hi | A lo | B hi | B lo]; A = 128 x 64 = 1024 units, B = BN x 64 = 8*BN units
  const uint32_t a_units = kGtBM * kGtBK / 8, b_units = (uint32_t)BN * kGtBK / 8;
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
    // the MMAs that read this stage two K-blocks ago have completed
    if (kb >= 2) ok = mbar_wait_bounded(s_bar + st, (uint32_t)(((kb >> 1) - 1) & 1)) && ok;
    // ---- A tile: 1024 units, 4 per thread -------------------------------------------------
#pragma unroll
    for (int r = 0; r < (kGtBM * kGtBK / 8) / kGtThreads; ++r) {
      const int u = tid + r * kGtThreads;
      uint4 hi, lo;
      if (!TA) {      // A[M,K], k contiguous: K-major planes [8 kg][128 m]
        const int m = u & (kGtBM - 1), kg = u >> 7;
        load_unit<SPLIT>(p.A, p.lda, m0 + m, k0 + kg * 8, p.M, p.K, p.vecA != 0, p.e.a_relu != 0, &hi, &lo);
        sA[kg * kGtBM + m] = hi;
        if (SPLIT) sA[a_units + kg * kGtBM + m] = lo;
      } else {        // A stored [K,M], m contiguous: MN-major planes [16 mg][64 k]
        const int mg = u & 15, k = u >> 4;
        load_unit<SPLIT>(p.A, p.lda, k0 + k, m0 + mg * 8, p.K, p.M, p.vecA != 0, p.e.a_relu != 0, &hi, &lo);
        sA[mg * kGtBK + k] = hi;
        if (SPLIT) sA[a_units + mg * kGtBK + k] = lo;
      }
    }
    // ---- B tile: 8*BN units -------------------------------------------------------------
    for (int u = tid; u < (int)b_units; u += kGtThreads) {
      uint4 hi, lo;
      if (TB) {       // B stored [N,K], k contiguous: K-major planes [8 kg][BN n]
        const int n = u % BN, kg = u / BN;
        load_unit<SPLIT>(p.B, p.ldb, n0 + n, k0 + kg * 8, p.N, p.K, p.vecB != 0, false, &hi, &lo);
        sB[kg * BN + n] = hi;
        if (SPLIT) sB[b_units + kg * BN + n] = lo;
      } else {        // B[K,N], n contiguous: MN-major planes [BN/8 ng][64 k]
        const int ng = u % (BN >> 3), k = u / (BN >> 3);
        load_unit<SPLIT>(p.B, p.ldb, k0 + k, n0 + ng * 8, p.K, p.N, p.vecB != 0, false, &hi, &lo);
        sB[ng * kGtBK + k] = hi;
        if (SPLIT) sB[b_units + ng * kGtBK + k] = lo;
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
        const uint64_t da = TA ? umma_desc(a_addr + ks * 256u, 128u, kGtBK * 16u)
                               : umma_desc(a_addr + ks * 2u * kGtBM * 16u, kGtBM * 16u, 128u);
        const uint64_t db = TB ? umma_desc(b_addr + ks * 2u * (uint32_t)BN * 16u, (uint32_t)BN * 16u, 128u)
                               : umma_desc(b_addr + ks * 256u, 128u, kGtBK * 16u);
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

  // ---- epilogue: thread = row (TMEM lane), 16 columns per tcgen05.ld ----------------------
  {
    const int q = warp & 3, half = warp >> 2;          // lanes 32q.., column half
    const int gm = m0 + q * 32 + lane;
    const int ngroups = BN / 16;
    const bool partial = p.ws != nullptr;
    float* crow = partial ? p.ws + ((size_t)blockIdx.z * p.M + gm) * p.N : p.C + (size_t)gm * p.ldc;
    for (int cg = half; cg < ngroups; cg += 2) {
      float v[16];
      if (nkb > 0) {
        tmem_ld<16>(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(cg * 16), v);
      } else {
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = 0.f;
      }
      if (gm < p.M) {
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const int gn = n0 + cg * 16 + j;
          if (gn < p.N) {
            float x = v[j];
            if (!partial) {
              if (p.e.bias) x += __ldg(p.e.bias + gn);
              if (p.e.relu) x = fmaxf(x, 0.f);
              if (p.e.mask) x = __ldg(p.e.mask + (size_t)gm * p.e.ldm + gn) > 0.f ? x : 0.f;
              if (p.e.accumulate) x += crow[gn];
            }
            crow[gn] = x;
          }
        }
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tcols));
}

// C = epilogue(sum_z ws[z]) in slice order.
__global__ void gemm_tc_reduce_kernel(int M, int N, int splits, const float* __restrict__ ws,
                                      float* __restrict__ C, int ldc, GemmEpi e) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= M * N) return;
  const int m = i / N, n = i - m * N;
  float x = 0.f;
  for (int z = 0; z < splits; ++z) x += ws[(size_t)z * M * N + i];
  if (e.bias) x += __ldg(e.bias + n);
  if (e.relu) x = fmaxf(x, 0.f);
  if (e.mask) x = __ldg(e.mask + (size_t)m * e.ldm + n) > 0.f ? x : 0.f;
  float* c = C + (size_t)m * ldc + n;
  *c = e.accumulate ? *c + x : x;
}

bool gemm_tc_supported(int M, int N, int K) { return M >= 64 && N >= 16 && K >= 32; }

size_t gemm_tc_workspace_bytes() { return (size_t)48 << 20; }

int gemm_tc(bool ta, bool tb, int split, int M, int N, int K, const float* A, int lda, const float* B,
            int ldb, float* C, int ldc, const GemmEpi& e, float* ws, size_t ws_bytes, int* err,
            cudaStream_t st) {
  if (M <= 0 || N <= 0) return SEEDRL_OK;
  GemmTcParams p;
  p.M = M; p.N = N; p.K = K; p.A = A; p.lda = lda; p.B = B; p.ldb = ldb; p.C = C; p.ldc = ldc;
  p.e = e; p.error_flag = err;
  const int n16 = ((N + 15) / 16) * 16;
  const int bn_max = split ? 128 : 256;                 // shared memory: 2 stages x (hi + lo)
  p.BN = n16 < bn_max ? n16 : bn_max;
  auto al16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
  p.vecA = al16(A) && (lda & 3) == 0;
  p.vecB = al16(B) && (ldb & 3) == 0;
  const int tiles = ceil_div(M, kGtBM) * ceil_div(N, p.BN);
  const int nkb = ceil_div(K, kGtBK);
  // split-K until the grid covers the SMs, keeping >= 4 K-blocks per slice and the
  // partials inside the workspace
  int splits = 1;
  while (tiles * splits < kNumSMs && nkb / (splits * 2) >= 4 &&
         (size_t)(splits * 2) * M * N * sizeof(float) <= ws_bytes && ws)
    splits *= 2;
  p.kblocks_per_split = ceil_div(nkb, splits);
  splits = ceil_div(nkb, p.kblocks_per_split);
  p.ws = splits > 1 ? ws : nullptr;
  const int S = split ? 2 : 1;
  const size_t smem = (size_t)2 * S * (kGtBM * kGtBK / 8 + p.BN * kGtBK / 8) * 16 + 64;
  dim3 grid(ceil_div(N, p.BN), ceil_div(M, kGtBM), splits);
#define SEEDRL_GT_LAUNCH(TA_, TB_, SP_)                                                         \
  do {                                                                                          \
    static bool attr = false;                                                                   \
    if (!attr) {                                                                                \
      SEEDRL_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<TA_, TB_, SP_>,                           \
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024)); \
      attr = true;                                                                              \
    }                                                                                           \
    gemm_tc_kernel<TA_, TB_, SP_><<<grid, kGtThreads, smem, st>>>(p);                           \
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
  count_launch(PC_GEMM, st);
  SEEDRL_CHECK_LAUNCH();
  if (splits > 1) {
    gemm_tc_reduce_kernel<<<ceil_div(M * N, 256), 256, 0, st>>>(M, N, splits, ws, C, ldc, e);
    count_launch(PC_GEMM, st);
    SEEDRL_CHECK_LAUNCH();
  }
  return SEEDRL_OK;
}

}  // namespace seedrl
