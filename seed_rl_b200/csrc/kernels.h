// Internal host-side launchers shared between translation units.
#pragma once
#include "common.cuh"

namespace seedrl {

enum { IN_F32 = 0, IN_RELU = 1, IN_U8 = 2 };

// conv_kernels.cu
int conv3x3_forward(int cin, int cout, int in_mode, int N, int H, int W, const void* in,
                    const float* w, const float* bias, const float* mask, const float* res,
                    float* out, cudaStream_t st);
int conv3x3_flip_weights(int cin, int cout, const float* w, float* wt, cudaStream_t st);
size_t conv3x3_wgrad_partial_bytes();
int wgrad_reduce(int nparts, int nw, int nb, const float* partial, float* dw, float* db,
                 cudaStream_t st);
int conv3x3_wgrad(int cin, int cout, int in_mode, int N, int H, int W, const void* x,
                  const float* dy, float* dw, float* db, float* partial, size_t partial_bytes,
                  cudaStream_t st);
int maxpool3s2_forward(int N, int H, int W, int C, const float* x, float* y, uint8_t* idx,
                       cudaStream_t st);
int maxpool3s2_backward(int N, int H, int W, int C, const float* dy, const uint8_t* idx, float* dx,
                        cudaStream_t st);

// convgen_kernels.cu (arbitrary kernel/stride 'valid' conv: shallow net)
int convgen_forward(int N, int H, int W, int cin, int cout, int k, int stride, int in_u8,
                    const void* in, const float* w, const float* bias, int relu, float* out,
                    cudaStream_t st);
int convgen_dgrad(int N, int H, int W, int cin, int cout, int k, int stride, const float* dy,
                  const float* w, const float* mask, float* dx, cudaStream_t st);
int convgen_wgrad(int N, int H, int W, int cin, int cout, int k, int stride, int in_u8,
                  const void* x, const float* dy, float* dw, float* db, float* partial,
                  size_t partial_bytes, cudaStream_t st);

// gemm_kernels.cu
struct GemmEpi {
  const float* bias;
  const float* mask;
  int ldm;
  int relu;
  int accumulate;
  int a_relu;
};
inline GemmEpi epi_none() { return GemmEpi{nullptr, nullptr, 0, 0, 0, 0}; }
int sgemm(bool ta, bool tb, int M, int N, int K, const float* A, int lda, const float* B, int ldb,
          float* C, int ldc, const GemmEpi& e, cudaStream_t st);
int colsum(int M, int N, const float* X, int ld, float* out, cudaStream_t st);
int core_input_tail(int Nrows, int D, int A, const float* reward, const int64_t* prev_action,
                    float* core_in, cudaStream_t st);
int lstm_mask_state(int B, int Hd, const uint8_t* done, const float* h_src, float* h_dst,
                    cudaStream_t st);
int lstm_pointwise_fwd(int B, int Hd, float* z, const float* c_prev_src, const uint8_t* done_t,
                       const uint8_t* done_next, float* c_out, float* h_out, float* hprev_next,
                       cudaStream_t st);
int lstm_pointwise_bwd(int B, int Hd, const float* gates, const float* c_t, const float* c_prev_src,
                       const uint8_t* done_t, const uint8_t* done_next, const float* dh_out,
                       const float* dh_rec, const float* dc_next, float* dz, float* dc_prev_out,
                       cudaStream_t st);
int fill(size_t n, float* p, float v, cudaStream_t st);

}  // namespace seedrl
