// Persistent LSTM recurrence (Keras LSTMCell(256) unrolled over time with done-resets,
// dmlab/networks.py:157-169) -- forward and BPTT -- as ONE cooperative kernel each.
//
// The recurrence is latency-bound: per step only B x 256 x 1024 MACs (B = 64) but 2 x T
// dependent steps.  Launching a GEMM + a pointwise kernel per step costs ~40-90 us/step;
// here 128 CTAs stay resident for all T steps, each owning 2 hidden units (8 gate
// columns): its slice of the recurrent matrix U stays in shared memory, the cell state
// (forward) / cell-state gradient (backward) of its units stays on chip, and the only
// per-step global traffic is the [B,256] hidden state (forward) or the [B,1024] gate
// gradient (backward) exchanged through L2 between two grid-wide barriers.
//   forward : z[t] (+)= hprev[t] U ; gates ; c,h ; emits hprev[t+1] = done[t+1] ? 0 : h
//   backward: dh = dH[t] + (done[t+1] ? 0 : dZ[t+1] U^T) ; gate gradients dZ[t]
// Launched with cudaLaunchCooperativeKernel (co-residency guaranteed or the launch fails);
// the grid barrier is a monotonic atomic counter with a bounded spin (sets *err, never hangs).
#include "kernels.h"

namespace seedrl {

constexpr int kLH = 256;            // hidden units
constexpr int kLUnits = 2;          // units per CTA
constexpr int kLGrid = kLH / kLUnits;
constexpr int kLThreads = 256;
constexpr int kLBt = 64;            // batch tile

__device__ __forceinline__ void grid_barrier(unsigned int* counter, unsigned int nblocks,
                                             unsigned int* gen, int* err) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    const unsigned int target = (*gen + 1u) * nblocks;
    atomicAdd(counter, 1u);
    int spins = 0;
    while (*reinterpret_cast<volatile unsigned int*>(counter) < target) {
      if (++spins > (1 << 24)) { if (err) atomicExch(err, 2); break; }
    }
    __threadfence();
  }
  *gen += 1u;
  __syncthreads();
}

struct LstmFwdArgs {
  int T1, B;
  const float* U;          // [256, 1024]
  const uint8_t* done;     // [T1, B]
  float* z;                // [T1, B, 1024]  in: x W + b ; out: activated gates (i,f,g,o)
  const float* h0;         // [B, 256]
  const float* c0;         // [B, 256]
  float* hs;               // [T1, B, 256]
  float* cs;               // [T1, B, 256]
  float* hp;               // [T1, B, 256]  masked recurrent inputs (kept for dU)
  unsigned int* counter;   // zeroed by the host before launch
  int* err;
};

__global__ void __launch_bounds__(kLThreads, 1) lstm_fwd_persistent_kernel(const LstmFwdArgs a) {
  extern __shared__ float sm[];
  float* s_U = sm;                               // [256][8]
  float* s_h = s_U + kLH * 8;                    // [64][260]
  float* s_z = s_h + kLBt * (kLH + 4);           // [64][8]
  float* s_c = s_z + kLBt * 8;                   // [B][2] cell state of this CTA's units
  const int tid = threadIdx.x;
  const int u0 = blockIdx.x * kLUnits;
  const int B = a.B;
  unsigned int gen = 0;
  // column c (0..7) = gate (c>>1) of unit u0 + (c&1)  ->  global column gate*256 + unit
  for (int i = tid; i < kLH * 8; i += kLThreads) {
    const int k = i >> 3, c = i & 7;
    s_U[i] = __ldg(a.U + (size_t)k * 4 * kLH + (c >> 1) * kLH + u0 + (c & 1));
  }
  for (int i = tid; i < B * kLUnits; i += kLThreads)
    s_c[i] = __ldg(a.c0 + (size_t)(i >> 1) * kLH + u0 + (i & 1));
  __syncthreads();

  for (int t = 0; t < a.T1; ++t) {
    const uint8_t* done_t = a.done + (size_t)t * B;
    const uint8_t* done_n = (t + 1 < a.T1) ? a.done + (size_t)(t + 1) * B : nullptr;
    float* hp_t = a.hp + (size_t)t * B * kLH;
    for (int b0 = 0; b0 < B; b0 += kLBt) {
      const int nb = min(kLBt, B - b0);
      // ---- recurrent input of step t for this batch tile -> smem --------------------------
      // (8 independent 16-byte loads in flight per thread: the step is latency-bound)
      for (int i0 = tid; i0 < nb * (kLH / 4); i0 += 8 * kLThreads) {
        float4 v[8];
#pragma unroll
        for (int r = 0; r < 8; ++r) {
          const int i = i0 + r * kLThreads;
          v[r] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (i < nb * (kLH / 4)) {
            const int b = i / (kLH / 4), k4 = i - b * (kLH / 4);
            if (t == 0) {
              if (!done_t[b0 + b]) v[r] = __ldg(reinterpret_cast<const float4*>(a.h0 + (size_t)(b0 + b) * kLH) + k4);
            } else {
              v[r] = __ldcg(reinterpret_cast<const float4*>(hp_t + (size_t)(b0 + b) * kLH) + k4);
            }
          }
        }
#pragma unroll
        for (int r = 0; r < 8; ++r) {
          const int i = i0 + r * kLThreads;
          if (i < nb * (kLH / 4)) {
            const int b = i / (kLH / 4), k4 = i - b * (kLH / 4);
            if (t == 0 && blockIdx.x == 0) reinterpret_cast<float4*>(hp_t + (size_t)(b0 + b) * kLH)[k4] = v[r];
            *reinterpret_cast<float4*>(s_h + b * (kLH + 4) + k4 * 4) = v[r];
          }
        }
      }
      __syncthreads();
      // ---- z[b, 8 cols] = x-part + h . U : thread = (b, column pair) ----------------------
      {
        const int b = tid >> 2, cp = (tid & 3) * 2;
        if (b < nb) {
          float acc0 = 0.f, acc1 = 0.f;
          const float* hrow = s_h + b * (kLH + 4);
#pragma unroll 8
          for (int k = 0; k < kLH; ++k) {
            const float hv = hrow[k];
            const float2 uv = *reinterpret_cast<const float2*>(s_U + k * 8 + cp);
            acc0 = fmaf(hv, uv.x, acc0);
            acc1 = fmaf(hv, uv.y, acc1);
          }
          const float* zrow = a.z + ((size_t)t * B + b0 + b) * 4 * kLH;
          const int g0 = cp >> 1;                 // cp, cp+1 = same gate, units u0 and u0+1
          s_z[b * 8 + cp] = acc0 + __ldg(zrow + g0 * kLH + u0);
          s_z[b * 8 + cp + 1] = acc1 + __ldg(zrow + g0 * kLH + u0 + 1);
        }
      }
      __syncthreads();
      // ---- pointwise: thread = (b, unit) -----------------------------------------------------
      if (tid < nb * kLUnits) {
        const int b = tid >> 1, ul = tid & 1;
        const float gi = sigmoidf_(s_z[b * 8 + 0 + ul]);
        const float gf = sigmoidf_(s_z[b * 8 + 2 + ul]);
        const float gg = tanhf(s_z[b * 8 + 4 + ul]);
        const float go = sigmoidf_(s_z[b * 8 + 6 + ul]);
        const int gb = b0 + b, u = u0 + ul;
        const float cp_ = done_t[gb] ? 0.f : s_c[gb * 2 + ul];
        const float c = gf * cp_ + gi * gg;
        const float h = go * tanhf(c);
        s_c[gb * 2 + ul] = c;
        float* zrow = a.z + ((size_t)t * B + gb) * 4 * kLH;
        zrow[u] = gi; zrow[kLH + u] = gf; zrow[2 * kLH + u] = gg; zrow[3 * kLH + u] = go;
        a.cs[((size_t)t * B + gb) * kLH + u] = c;
        a.hs[((size_t)t * B + gb) * kLH + u] = h;
        if (done_n) a.hp[((size_t)(t + 1) * B + gb) * kLH + u] = done_n[gb] ? 0.f : h;
      }
      __syncthreads();
    }
    if (t + 1 < a.T1) grid_barrier(a.counter, gridDim.x, &gen, a.err);
  }
}

struct LstmBwdArgs {
  int T1, B;
  const float* U;          // [256, 1024]
  const uint8_t* done;     // [T1, B]
  const float* gates;      // [T1, B, 1024] activated gates from the forward
  const float* cs;         // [T1, B, 256]
  const float* c0;         // [B, 256]
  const float* dhs;        // [T1, B, 256]  d loss / d h_t from the heads
  float* dz;               // [T1, B, 1024] out: gate pre-activation gradients
  unsigned int* counter;
  int* err;
};

__global__ void __launch_bounds__(kLThreads, 1) lstm_bwd_persistent_kernel(const LstmBwdArgs a) {
  extern __shared__ float sm[];
  constexpr int KC = 128, KS = KC + 4;           // dZ chunk width (+pad)
  float* s_Ur = sm;                              // [2][1024] rows u0, u0+1 of U
  float* s_dz = s_Ur + kLUnits * 4 * kLH;        // [2][64][KS] double-buffered
  float* s_dh = s_dz + 2 * kLBt * KS;            // [64][2] recurrent part of dh
  float* s_dc = s_dh + kLBt * kLUnits;           // [B][2] dc flowing to the previous step
  const int tid = threadIdx.x;
  const int u0 = blockIdx.x * kLUnits;
  const int B = a.B;
  unsigned int gen = 0;
  for (int i = tid; i < kLUnits * 4 * kLH; i += kLThreads)
    s_Ur[i] = __ldg(a.U + (size_t)(u0 + i / (4 * kLH)) * 4 * kLH + (i % (4 * kLH)));
  for (int i = tid; i < B * kLUnits; i += kLThreads) s_dc[i] = 0.f;
  __syncthreads();

  for (int t = a.T1 - 1; t >= 0; --t) {
    const bool last = (t + 1 == a.T1);
    const uint8_t* done_t = a.done + (size_t)t * B;
    const uint8_t* done_n = last ? nullptr : a.done + (size_t)(t + 1) * B;
    for (int b0 = 0; b0 < B; b0 += kLBt) {
      const int nb = min(kLBt, B - b0);
      // ---- dh_rec[b, 2 units] = dZ[t+1][b, :] . U[u, :]^T, streamed in K chunks of 128 ------
      const int b = tid >> 2, part = tid & 3;
      float acc0 = 0.f, acc1 = 0.f;
      if (!last) {
        const float* dzn = a.dz + ((size_t)(t + 1) * B + b0) * 4 * kLH;
        // chunk c+1 travels L2 -> registers while chunk c is consumed from shared memory
        constexpr int NL = (kLBt * (KC / 4)) / kLThreads;   // 8 x 16-byte loads per thread per chunk
        float4 pre[NL];
        auto load_chunk = [&](int k0) {
#pragma unroll
          for (int r = 0; r < NL; ++r) {
            const int i = tid + r * kLThreads;
            const int rr = i / (KC / 4), k4 = i - rr * (KC / 4);
            pre[r] = rr < nb ? __ldcg(reinterpret_cast<const float4*>(dzn + (size_t)rr * 4 * kLH + k0) + k4)
                             : make_float4(0.f, 0.f, 0.f, 0.f);
          }
        };
        load_chunk(0);
        for (int c = 0; c < (4 * kLH) / KC; ++c) {
          const int k0 = c * KC;
          float* buf = s_dz + (c & 1) * (kLBt * KS);
#pragma unroll
          for (int r = 0; r < NL; ++r) {
            const int i = tid + r * kLThreads;
            const int rr = i / (KC / 4), k4 = i - rr * (KC / 4);
            *reinterpret_cast<float4*>(buf + rr * KS + k4 * 4) = pre[r];
          }
          __syncthreads();
          if (c + 1 < (4 * kLH) / KC) load_chunk(k0 + KC);
          if (b < nb) {
            const float* drow = buf + b * KS + part * (KC / 4);
            const float* u0r = s_Ur + k0 + part * (KC / 4);
            const float* u1r = u0r + 4 * kLH;
#pragma unroll 8
            for (int k = 0; k < KC / 4; ++k) {
              const float d = drow[k];
              acc0 = fmaf(d, u0r[k], acc0);
              acc1 = fmaf(d, u1r[k], acc1);
            }
          }
        }
        __syncthreads();   // the last chunk's buffer is free before the next batch tile / step
        // reduce the 4 K-parts (adjacent lanes)
        acc0 += __shfl_xor_sync(0xffffffffu, acc0, 1); acc0 += __shfl_xor_sync(0xffffffffu, acc0, 2);
        acc1 += __shfl_xor_sync(0xffffffffu, acc1, 1); acc1 += __shfl_xor_sync(0xffffffffu, acc1, 2);
      }
      if (part == 0 && b < nb) { s_dh[b * 2] = acc0; s_dh[b * 2 + 1] = acc1; }
      __syncthreads();
      // ---- pointwise backward: thread = (b, unit) ------------------------------------------
      if (tid < nb * kLUnits) {
        const int bb = tid >> 1, ul = tid & 1;
        const int gb = b0 + bb, u = u0 + ul;
        const float* gr = a.gates + ((size_t)t * B + gb) * 4 * kLH;
        const float gi = __ldg(gr + u), gf = __ldg(gr + kLH + u), gg = __ldg(gr + 2 * kLH + u),
                    go = __ldg(gr + 3 * kLH + u);
        const bool cut = done_n && done_n[gb];
        float dh = __ldg(a.dhs + ((size_t)t * B + gb) * kLH + u);
        if (!last && !cut) dh += s_dh[bb * 2 + ul];
        const float tc = tanhf(__ldg(a.cs + ((size_t)t * B + gb) * kLH + u));
        float dc = dh * go * (1.f - tc * tc);
        if (!last && !cut) dc += s_dc[gb * 2 + ul];
        const float cprev = done_t[gb] ? 0.f
                            : (t == 0 ? __ldg(a.c0 + (size_t)gb * kLH + u)
                                      : __ldg(a.cs + ((size_t)(t - 1) * B + gb) * kLH + u));
        float* dzr = a.dz + ((size_t)t * B + gb) * 4 * kLH;
        dzr[u] = dc * gg * gi * (1.f - gi);
        dzr[kLH + u] = dc * cprev * gf * (1.f - gf);
        dzr[2 * kLH + u] = dc * gi * (1.f - gg * gg);
        dzr[3 * kLH + u] = dh * tc * go * (1.f - go);
        s_dc[gb * 2 + ul] = dc * gf;
      }
      __syncthreads();
    }
    if (t > 0) grid_barrier(a.counter, gridDim.x, &gen, a.err);
  }
}

static size_t lstm_fwd_smem(int B) {
  return ((size_t)kLH * 8 + (size_t)kLBt * (kLH + 4) + kLBt * 8 + (size_t)B * kLUnits) * sizeof(float);
}
static size_t lstm_bwd_smem(int B) {
  return ((size_t)kLUnits * 4 * kLH + (size_t)2 * kLBt * 132 + kLBt * kLUnits + (size_t)B * kLUnits) * sizeof(float);
}

int lstm_forward_persistent(int T1, int B, const float* U, const uint8_t* done, float* z,
                            const float* h0, const float* c0, float* hs, float* cs, float* hp,
                            unsigned int* counter, int* err, cudaStream_t st) {
  const size_t smem = lstm_fwd_smem(B);
  if (smem > 200 * 1024) return set_error(SEEDRL_ERR_INVALID_ARGUMENT, "lstm: batch too large");
  static bool attr = false;
  if (!attr) {
    SEEDRL_CUDA(cudaFuncSetAttribute(lstm_fwd_persistent_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     200 * 1024));
    attr = true;
  }
  SEEDRL_CUDA(cudaMemsetAsync(counter, 0, sizeof(unsigned int), st));
  LstmFwdArgs a{T1, B, U, done, z, h0, c0, hs, cs, hp, counter, err};
  void* args[] = {&a};
  SEEDRL_CUDA(cudaLaunchCooperativeKernel((const void*)lstm_fwd_persistent_kernel, dim3(kLGrid),
                                          dim3(kLThreads), args, smem, st));
  count_launch(PC_LSTM_PW, st);
  return SEEDRL_OK;
}

int lstm_backward_persistent(int T1, int B, const float* U, const uint8_t* done, const float* gates,
                             const float* cs, const float* c0, const float* dhs, float* dz,
                             unsigned int* counter, int* err, cudaStream_t st) {
  const size_t smem = lstm_bwd_smem(B);
  if (smem > 200 * 1024) return set_error(SEEDRL_ERR_INVALID_ARGUMENT, "lstm: batch too large");
  static bool attr = false;
  if (!attr) {
    SEEDRL_CUDA(cudaFuncSetAttribute(lstm_bwd_persistent_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     200 * 1024));
    attr = true;
  }
  SEEDRL_CUDA(cudaMemsetAsync(counter, 0, sizeof(unsigned int), st));
  LstmBwdArgs a{T1, B, U, done, gates, cs, c0, dhs, dz, counter, err};
  void* args[] = {&a};
  SEEDRL_CUDA(cudaLaunchCooperativeKernel((const void*)lstm_bwd_persistent_kernel, dim3(kLGrid),
                                          dim3(kLThreads), args, smem, st));
  count_launch(PC_LSTM_PW, st);
  return SEEDRL_OK;
}

}  // namespace seedrl
