// Per-thread bodies of the R2D2 stack_frames and loss kernels, written
// so that the SAME source text compiles as device code (r2d2_kernels.cu) and, with
// SEEDRL_HD defined as `inline`, as plain host C++ (tests/host_emulation/r2d2_host.cpp): the
// CPU test suite runs these bodies thread by thread against oracle/r2d2_oracle.py.  That
// checks the algorithm and its indexing -- not the GPU execution.
#pragma once
#include <math.h>
#include <stdint.h>
#include <stddef.h>

#ifndef SEEDRL_HD
#define SEEDRL_HD __host__ __device__ __forceinline__
#endif

namespace seedrl {

// Thread (b, p): walks t = 0..T-1 with the last S frames in registers.
//   stacked[t,b,p,i] = ext[t + S-1-i], ext = (S-1 unpacked state frames, oldest first) ++ frames,
//   zeroed when an episode boundary lies in (t-i, t]  <=>  i > a_t, a_t = steps since the most
//   recent done[.] = true at or before t inside this unroll (infinite if none).
//   new_state byte j (LSB = oldest) = masked stacked[T-1,b,p,S-2-j].      (atari/networks.py:57-173)
template <int S>
SEEDRL_HD void r2d2_stack_frames_thread(int T, int B, int P, int b, int p, const uint8_t* frames,
                                        const int32_t* state_in, const uint8_t* done, uint8_t* stacked,
                                        int32_t* state_out) {
  const uint32_t st = (uint32_t)state_in[(size_t)b * P + p];
  // w[i], i >= 1: the frame i steps before the one about to arrive; the newest kept frame is
  // byte S-2 of the packed state (LSB byte = oldest)
  uint32_t w[S], last[S];
  w[0] = 0;
  for (int i = 1; i < S; ++i) w[i] = (st >> (8 * (S - 1 - i))) & 0xFFu;
  for (int i = 0; i < S; ++i) last[i] = 0;
  int age = 1 << 20;                                  // no done seen yet: nothing is masked
  for (int t = 0; t < T; ++t) {
    const uint32_t f = frames[((size_t)t * B + b) * P + p];
    age = done[(size_t)t * B + b] ? 0 : age + 1;
    uint8_t* o = stacked + (((size_t)t * B + b) * P + p) * S;
    o[0] = (uint8_t)f;
    last[0] = f;
    for (int i = 1; i < S; ++i) {
      const uint32_t v = (i <= age) ? w[i] : 0u;     // masking never feeds back into the window
      o[i] = (uint8_t)v;
      last[i] = v;
    }
    for (int i = S - 1; i >= 2; --i) w[i] = w[i - 1];
    w[1] = f;
  }
  uint32_t ns = 0;
  for (int j = 0; j < S - 1; ++j) ns |= last[S - 2 - j] << (8 * j);
  if (T == 0) ns = st;
  state_out[(size_t)b * P + p] = (int32_t)ns;
}

SEEDRL_HD float vf_rescale(float x, float eps) {              // learner.py:180-183
  const float s = (float)((x > 0.f) - (x < 0.f));
  return s * (sqrtf(fabsf(x) + 1.f) - 1.f) + eps * x;
}
SEEDRL_HD float vf_rescale_inv(float x, float eps) {          // learner.py:186-192
  const float s = (float)((x > 0.f) - (x < 0.f));
  const float inner = (sqrtf(1.f + 4.f * eps * (fabsf(x) + 1.f + eps)) - 1.f) / (2.f * eps);
  return s * (inner * inner - 1.f);
}

struct R2d2LossParams {
  int T, B, A, n_steps;
  const float* q_train;     // [T,B,A]
  const float* q_target;    // [T,B,A]
  const int64_t* replay_action;   // [T,B]
  const float* reward;      // [T,B]
  const uint8_t* done;      // [T,B]
  const float* is_weights;  // [B] or null (= 1)
  float gamma, eta, eps;
  float gamma_pow[8];       // fp32(gamma ** k): the reference DIVIDES the padded targets by it
  float* loss;              // [B]
  float* priorities;        // [B]
  float* dq;                // [T,B,A]  d mean_b(w_b loss_b) / d q_train
  float* scratch;           // [B][T + n_steps] Bellman-target work array
};

// Thread b = one sequence (B ~ 64, T ~ 100: negligible next to the two network unrolls).
// Follows n_step_bellman_target literally (padded arrays, n_steps in-place passes) so that the
// fp32 rounding order is the oracle's.
SEEDRL_HD void r2d2_loss_thread(const R2d2LossParams& p, int b) {
  const int T = p.T, B = p.B, A = p.A, n = p.n_steps;
  float* bt = p.scratch + (size_t)b * (T + n);
  // bellman_target = [0, qmax_0 .. qmax_{T-1}, qmax_{T-1}/gamma^1 .. /gamma^{n-1}]   (:241-246)
  bt[0] = 0.f;
  float qlast = 0.f;
  for (int t = 0; t < T; ++t) {
    const float* qt = p.q_train + ((size_t)t * B + b) * A;
    int best = 0;
    float bv = qt[0];
    for (int a = 1; a < A; ++a)
      if (qt[a] > bv) { bv = qt[a]; best = a; }                                  // argmax, first max
    qlast = vf_rescale_inv(p.q_target[((size_t)t * B + b) * A + best], p.eps);   // :303-305
    bt[1 + t] = qlast;
  }
  for (int k = 1; k < n; ++k) bt[T + k] = qlast / p.gamma_pow[k];
  // n passes of  target = r + gamma (1 - done) target[1:]  over the zero-padded r / done (:250-253)
  for (int j = 1; j <= n; ++j) {
    const int len = T + n - j;                       // length after dropping j padded rows
    for (int i = 0; i < len; ++i) {
      const float r = i < T ? p.reward[(size_t)i * B + b] : 0.f;
      const float nd = (i < T && p.done[(size_t)i * B + b]) ? 0.f : 1.f;
      bt[i] = r + p.gamma * nd * bt[i + 1];
    }
  }
  // td_t = h(target[t+1]) - Q(s_t, a_t), t < T-1  (:316-322)
  const float w = (p.is_weights ? p.is_weights[b] : 1.f) / (float)B;             // d mean_b(w_b loss_b)
  float mx = 0.f, sum = 0.f, sq = 0.f;
  for (int t = 0; t < T; ++t) {
    float* dq = p.dq + ((size_t)t * B + b) * A;
    for (int a = 0; a < A; ++a) dq[a] = 0.f;
    if (t + 1 < T) {
      int64_t a = p.replay_action[(size_t)t * B + b];
      a = a < 0 ? 0 : (a >= A ? A - 1 : a);
      const float rq = p.q_train[((size_t)t * B + b) * A + a];
      const float tgt = vf_rescale(bt[t + 1], p.eps);
      const float td = tgt - rq;
      const float ad = fabsf(td);
      mx = fmaxf(mx, ad); sum += ad; sq += ad * ad;
      dq[a] = -w * td;                               // d(0.5 td^2)/d rq = -(tgt - rq), target is stop-gradient
    }
  }
  const int Tm = T - 1 > 0 ? T - 1 : 1;
  p.priorities[b] = p.eta * mx + (1.f - p.eta) * (sum / (float)Tm);             // :325-326
  p.loss[b] = 0.5f * sq;                                                         // :329
}

// ---- prioritized replay sampling (common/utils.py:327-352), phase bodies ---------------------
// phase 1 (thread i): s_cdf[i] = prio_i ^ alpha
SEEDRL_HD void replay_pow_thread(int i, const float* priorities, float priority_exp, float* s_cdf) {
  s_cdf[i] = powf(priorities[i], priority_exp);
}
// phase 2 (one thread): inclusive prefix in index order (deterministic); returns the total
SEEDRL_HD float replay_prefix_serial(int limit, float* s_cdf) {
  float acc = 0.f;
  for (int i = 0; i < limit; ++i) { acc += s_cdf[i]; s_cdf[i] = acc; }
  return acc;
}
// phase 3 (thread j): index = first i with cdf[i] > u_j * total; un-normalised importance weight
SEEDRL_HD float replay_sample_thread(int j, int limit, const float* s_cdf, float total, float is_exp,
                                     const float* uniforms, int64_t* indices, float* weights) {
  // u < total strictly (uniforms[j] * total can round up to total): then the first i with
  // cdf[i] > u exists and has cdf[i] > u >= cdf[i-1], i.e. positive mass -- like
  // tf.random.categorical over log-probabilities, a zero-priority entry is never drawn.
  float u = uniforms[j] * total;
  const float below = total > 0.f ? total * (1.f - 5.9604645e-8f) : 0.f;
  if (u > below) u = below;
  int lo = 0, hi = limit - 1;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (s_cdf[mid] > u) hi = mid; else lo = mid + 1;
  }
  const float pi = (s_cdf[lo] - (lo ? s_cdf[lo - 1] : 0.f)) / total;
  const float wj = powf((1.f / (float)limit) / pi, is_exp);
  indices[j] = lo;
  weights[j] = wj;
  return wj;
}

}  // namespace seedrl
