// (a4) Fused multi-tensor Adam over the flat parameter arena, tf.keras semantics in the
// form TF's fused ApplyAdam kernel uses: m += (g-m)(1-b1); v += (g*g-v)(1-b2);
// var -= (m*lr_t)/(sqrt(v)+eps), all fp32 (incl. the 1-b subtraction)
// (TF 2.4.1 Adam._resource_apply_dense; built at dmlab/vtrace_main.py:46-51,
// applied at agents/vtrace/learner.py:272-273).  Pure HBM stream:
// 4 reads + 3 writes of fp32 per parameter = 28 B/param; float4 vectorised,
// grid = a multiple of the SM count with a grid-stride loop.
#include "common.cuh"

namespace seedrl {

__global__ void __launch_bounds__(256)
adam_kernel(size_t n, float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
            float* __restrict__ v, float lr_t, float b1, float b2, float eps, float gscale,
            long long clamp_index, float clamp_lo, float clamp_hi) {
  const size_t n4 = n >> 2;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const float ob1 = 1.0f - b1, ob2 = 1.0f - b2;
  float4* p4 = reinterpret_cast<float4*>(p);
  const float4* g4 = reinterpret_cast<const float4*>(g);
  float4* m4 = reinterpret_cast<float4*>(m);
  float4* v4 = reinterpret_cast<float4*>(v);
  for (size_t i = tid; i < n4; i += stride) {
    float4 pp = p4[i], gg = g4[i], mm = m4[i], vv = v4[i];
#define SEEDRL_ADAM1(c)                                   \
    {                                                     \
      const float gs = gg.c * gscale;                     \
      mm.c += (gs - mm.c) * ob1;                          \
      vv.c += (gs * gs - vv.c) * ob2;                     \
      pp.c -= (mm.c * lr_t) / (sqrtf(vv.c) + eps);        \
    }
    SEEDRL_ADAM1(x) SEEDRL_ADAM1(y) SEEDRL_ADAM1(z) SEEDRL_ADAM1(w)
#undef SEEDRL_ADAM1
    p4[i] = pp; m4[i] = mm; v4[i] = vv;
  }
  // tail (n % 4) + optional clamp handled by the first threads
  for (size_t i = (n4 << 2) + tid; i < n; i += stride) {
    const float gs = g[i] * gscale;
    const float mm = m[i] + (gs - m[i]) * ob1;
    const float vv = v[i] + (gs * gs - v[i]) * ob2;
    m[i] = mm; v[i] = vv;
    p[i] -= (mm * lr_t) / (sqrtf(vv) + eps);
  }
  (void)clamp_index; (void)clamp_lo; (void)clamp_hi;
}

// The entropy_cost_param constraint (learner.py:229-231) applied after the
// update; a separate 1-thread kernel keeps adam_kernel free of a grid-wide
// ordering hazard on that element.
__global__ void clamp_one_kernel(float* p, long long idx, float lo, float hi) {
  p[idx] = fminf(fmaxf(p[idx], lo), hi);
}

}  // namespace seedrl

using namespace seedrl;

extern "C" int seedrl_adam_apply(size_t n, float* params, const float* grads, float* m, float* v,
                                 float lr_t, float beta1, float beta2, float eps,
                                 float grad_scale, int64_t clamp_index, float clamp_lo,
                                 float clamp_hi, seedrl_stream_t stream) {
  if (n == 0) return SEEDRL_OK;
  SEEDRL_CHECK_ARG(params && grads && m && v, "null pointer");
  SEEDRL_CHECK_ARG(((uintptr_t)params % 16 == 0) && ((uintptr_t)grads % 16 == 0) &&
                       ((uintptr_t)m % 16 == 0) && ((uintptr_t)v % 16 == 0),
                   "arena pointers must be 16-byte aligned");
  SEEDRL_CHECK_ARG(clamp_index < (int64_t)n, "clamp_index out of range");
  const size_t n4 = n >> 2;
  int blocks = (int)ceil_div_sz(n4 > 0 ? n4 : 1, 256);
  const int max_blocks = kNumSMs * 8;
  if (blocks > max_blocks) blocks = max_blocks;
  adam_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(n, params, grads, m, v, lr_t, beta1,
                                                         beta2, eps, grad_scale,
                                                         (long long)clamp_index, clamp_lo,
                                                         clamp_hi);
  count_launch(PC_ADAM, (cudaStream_t)stream);
  SEEDRL_CHECK_LAUNCH();
  if (clamp_index >= 0) {
    clamp_one_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(params, (long long)clamp_index, clamp_lo,
                                                         clamp_hi);
    count_launch(PC_ADAM, (cudaStream_t)stream);
    SEEDRL_CHECK_LAUNCH();
  }
  return SEEDRL_OK;
}
