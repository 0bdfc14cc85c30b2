// tcgen05 / TMEM / mbarrier device helpers shared by the tensor-core kernels (sm_100a).
#pragma once
#include <cuda_bf16.h>
#include <stdint.h>

namespace seedrl {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// One lane of a CONVERGED warp (elect.sync): inside `if (elect_one())` the compiler knows a
// single thread is active, so uniform-datapath instructions (UTCHMMA, UTCBAR) are issued
// directly instead of through a per-instruction election loop.
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred P;\n\t"
      "elect.sync _|P, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, P;\n\t}\n"
      : "=r"(pred));
  return pred != 0;
}

// cute::UMMA::SmemDescriptor (mma_sm100_desc.hpp): start>>4 [0,14), LBO>>4 [16,30),
// SBO>>4 [32,46), version=1 [46,48), layout_type SWIZZLE_NONE=0 [61,64).
__device__ __forceinline__ uint64_t umma_desc(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}

// cute::UMMA::InstrDescriptor: c_format F32 (1<<4), a/b BF16 (1<<7, 1<<10), K-major both,
// n_dim = N>>3 at [17,23), m_dim = M>>4 at [24,29).
__host__ __device__ constexpr uint32_t umma_idesc(int M, int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t a, uint64_t b, uint32_t idesc,
                                         uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p; \n\t"
      "}\n" ::"r"(tmem_d), "l"(a), "l"(b), "r"(idesc), "r"(accumulate));
}

template <int N>
__device__ __forceinline__ void tmem_ld(uint32_t taddr, float* v);
template <>
__device__ __forceinline__ void tmem_ld<16>(uint32_t taddr, float* v) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32"
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
template <>
__device__ __forceinline__ void tmem_ld<8>(uint32_t taddr, float* v) {
  uint32_t r[8];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x8.b32"
      "{%0, %1, %2, %3, %4, %5, %6, %7}, [%8];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}
template <>
__device__ __forceinline__ void tmem_ld<32>(uint32_t taddr, float* v) {
  tmem_ld<16>(taddr, v);
  tmem_ld<16>(taddr + 16, v + 16);
}

__device__ __forceinline__ uint4 pack8_bf16(float4 a, float4 c) {
  __nv_bfloat162 p0 = __floats2bfloat162_rn(a.x, a.y), p1 = __floats2bfloat162_rn(a.z, a.w);
  __nv_bfloat162 p2 = __floats2bfloat162_rn(c.x, c.y), p3 = __floats2bfloat162_rn(c.z, c.w);
  uint4 r;
  r.x = *reinterpret_cast<uint32_t*>(&p0); r.y = *reinterpret_cast<uint32_t*>(&p1);
  r.z = *reinterpret_cast<uint32_t*>(&p2); r.w = *reinterpret_cast<uint32_t*>(&p3);
  return r;
}
// residual of the bf16 rounding: v - float(bf16(v)), componentwise (exact in fp32)
__device__ __forceinline__ float bf16_resid(float v) { return v - __bfloat162float(__float2bfloat16_rn(v)); }
__device__ __forceinline__ float4 bf16_resid4(float4 v) {
  return make_float4(bf16_resid(v.x), bf16_resid(v.y), bf16_resid(v.z), bf16_resid(v.w));
}

// bounded mbarrier wait (never hang the GPU): false if the spin budget ran out
__device__ __forceinline__ bool mbar_wait_bounded(uint64_t* bar, uint32_t parity) {
  uint32_t done = 0;
  int spins = 0;
  while (!done) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}\n"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    if (!done && ++spins > (1 << 22)) return false;
  }
  return true;
}

}  // namespace seedrl
