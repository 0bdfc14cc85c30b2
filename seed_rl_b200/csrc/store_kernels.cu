// (a7/a8) GPU-resident per-environment state: UnrollStore (common/utils.py:119-257) and
// the row scatter/gather behind Aggregator (common/utils.py:461-543).  Pure byte movement
// (HBM-bound): rows are copied with 16-byte vector accesses when row_bytes and the base
// pointers allow it, one CTA-row pair per (env, step) row otherwise byte-wise.
#include "common.cuh"

namespace seedrl {

__device__ __forceinline__ void copy_row(uint8_t* __restrict__ dst, const uint8_t* __restrict__ src,
                                         size_t bytes) {
  if (((uintptr_t)dst | (uintptr_t)src | bytes) % 16 == 0) {
    const int4* s = reinterpret_cast<const int4*>(src);
    int4* d = reinterpret_cast<int4*>(dst);
    for (size_t i = threadIdx.x; i < bytes / 16; i += blockDim.x) d[i] = s[i];
  } else if (((uintptr_t)dst | (uintptr_t)src | bytes) % 4 == 0) {
    const int* s = reinterpret_cast<const int*>(src);
    int* d = reinterpret_cast<int*>(dst);
    for (size_t i = threadIdx.x; i < bytes / 4; i += blockDim.x) d[i] = s[i];
  } else {
    for (size_t i = threadIdx.x; i < bytes; i += blockDim.x) dst[i] = src[i];
  }
}

// state[env_ids[j], index[env_ids[j]], :] = values[j, :]      utils.py:187-190
__global__ void store_append_kernel(uint8_t* state, const int32_t* __restrict__ index,
                                    const int32_t* __restrict__ env_ids, int full_length,
                                    size_t row_bytes, const uint8_t* __restrict__ values) {
  const int j = blockIdx.x;
  const int env = env_ids[j];
  const int t = index[env];
  if (t < 0 || t >= full_length) return;   // (the reference leaves OOB undefined)
  copy_row(state + ((size_t)env * full_length + t) * row_bytes, values + (size_t)j * row_bytes,
           row_bytes);
}

// index[env]++ ; completed ids compacted in env_ids order (utils.py:194,229-233).
// Single CTA: n is an inference batch (<= a few thousand).
__global__ void store_advance_kernel(int32_t* index, const int32_t* __restrict__ env_ids, int n,
                                     int full_length, int32_t* __restrict__ completed_ids,
                                     int32_t* __restrict__ num_completed) {
  __shared__ int s_base;
  __shared__ int s_warp[32];
  if (threadIdx.x == 0) s_base = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (int j0 = 0; j0 < n; j0 += blockDim.x) {
    const int j = j0 + threadIdx.x;
    int flag = 0, env = 0;
    if (j < n) {
      env = env_ids[j];
      const int v = index[env] + 1;
      index[env] = v;
      flag = (v == full_length);
    }
    const unsigned bal = __ballot_sync(0xffffffffu, flag);
    const int within = __popc(bal & ((1u << lane) - 1));
    if (lane == 0) s_warp[w] = __popc(bal);
    __syncthreads();
    int off = s_base;
    for (int k = 0; k < w; ++k) off += s_warp[k];
    if (flag) completed_ids[off + within] = env;
    __syncthreads();
    if (threadIdx.x == 0) {
      int tot = 0;
      for (int k = 0; k < nw; ++k) tot += s_warp[k];
      s_base += tot;
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) *num_completed = s_base;
}

// unrolls <- state[completed] ; then state[env, :ov+1] = state[env, full-(ov+1):]   utils.py:234-252
// grid = (n_completed, full_length): block (i, t) copies row t of unroll i.
__global__ void store_gather_kernel(const uint8_t* __restrict__ state,
                                    const int32_t* __restrict__ completed_ids, int n_completed,
                                    int full_length, size_t row_bytes, int time_major,
                                    uint8_t* __restrict__ unrolls) {
  const int i = blockIdx.x, t = blockIdx.y;
  const int env = completed_ids[i];
  const size_t dst_row = time_major ? ((size_t)t * n_completed + i) : ((size_t)i * full_length + t);
  copy_row(unrolls + dst_row * row_bytes, state + ((size_t)env * full_length + t) * row_bytes,
           row_bytes);
}
// state[env, r] = unroll_copy[i, full_length - j + r]: the carried rows are read from
// the gathered copy, so source and destination never alias (the reference's
// overlap >= unroll_length/2 case, tests/utils_test.py:191-271).
__global__ void store_carry_kernel(uint8_t* state, const int32_t* __restrict__ completed_ids,
                                   int n_completed, int full_length, size_t row_bytes, int j,
                                   int time_major, const uint8_t* __restrict__ unrolls) {
  const int i = blockIdx.x, r = blockIdx.y;   // r < j
  const int env = completed_ids[i];
  const int t = full_length - j + r;
  const size_t src_row = time_major ? ((size_t)t * n_completed + i) : ((size_t)i * full_length + t);
  copy_row(state + ((size_t)env * full_length + r) * row_bytes, unrolls + src_row * row_bytes,
           row_bytes);
}
__global__ void store_set_index_kernel(int32_t* index, const int32_t* __restrict__ ids, int n, int v) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) index[ids[i]] = v;
}
// state[env, :j] = 0
__global__ void store_zero_rows_kernel(uint8_t* state, const int32_t* __restrict__ ids,
                                       int full_length, size_t row_bytes) {
  const int i = blockIdx.x, r = blockIdx.y;
  uint8_t* p = state + ((size_t)ids[i] * full_length + r) * row_bytes;
  for (size_t k = threadIdx.x; k < row_bytes; k += blockDim.x) p[k] = 0;
}

}  // namespace seedrl

using namespace seedrl;

extern "C" int seedrl_store_append_field(uint8_t* state, const int32_t* index, const int32_t* env_ids,
                                         int n, int full_length, size_t row_bytes,
                                         const uint8_t* values, seedrl_stream_t stream) {
  if (n == 0 || row_bytes == 0) return SEEDRL_OK;
  SEEDRL_CHECK_ARG(state && index && env_ids && values && n > 0 && full_length > 0, "bad argument");
  const int threads = row_bytes >= 4096 ? 256 : (row_bytes >= 512 ? 128 : 32);
  store_append_kernel<<<n, threads, 0, (cudaStream_t)stream>>>(state, index, env_ids, full_length,
                                                               row_bytes, values);
  count_launch(PC_MISC, (cudaStream_t)stream);
  SEEDRL_CHECK_LAUNCH();
  return SEEDRL_OK;
}

extern "C" int seedrl_store_advance(int32_t* index, const int32_t* env_ids, int n, int full_length,
                                    int32_t* completed_ids, int32_t* num_completed,
                                    seedrl_stream_t stream) {
  SEEDRL_CHECK_ARG(index && env_ids && completed_ids && num_completed && n >= 0, "bad argument");
  store_advance_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(index, env_ids, n, full_length,
                                                            completed_ids, num_completed);
  count_launch(PC_MISC, (cudaStream_t)stream);
  SEEDRL_CHECK_LAUNCH();
  return SEEDRL_OK;
}

extern "C" int seedrl_store_gather_field(uint8_t* state, const int32_t* completed_ids,
                                         int n_completed, int full_length, size_t row_bytes,
                                         int overlap, int time_major, uint8_t* unrolls,
                                         seedrl_stream_t stream) {
  if (n_completed == 0 || row_bytes == 0) return SEEDRL_OK;
  SEEDRL_CHECK_ARG(state && completed_ids && unrolls && full_length > 0 && overlap >= 0, "bad argument");
  const int j = overlap + 1;
  SEEDRL_CHECK_ARG(full_length >= j, "num_overlapping_steps + 1 exceeds the unroll");
  const int threads = row_bytes >= 4096 ? 256 : (row_bytes >= 512 ? 128 : 32);
  store_gather_kernel<<<dim3(n_completed, full_length), threads, 0, (cudaStream_t)stream>>>(
      state, completed_ids, n_completed, full_length, row_bytes, time_major, unrolls);
  count_launch(PC_MISC, (cudaStream_t)stream);
  SEEDRL_CHECK_LAUNCH();
  store_carry_kernel<<<dim3(n_completed, j), threads, 0, (cudaStream_t)stream>>>(
      state, completed_ids, n_completed, full_length, row_bytes, j, time_major, unrolls);
  count_launch(PC_MISC, (cudaStream_t)stream);
  SEEDRL_CHECK_LAUNCH();
  return SEEDRL_OK;
}

extern "C" int seedrl_store_finish(int32_t* index, const int32_t* completed_ids, int n_completed,
                                   int overlap, seedrl_stream_t stream) {
  if (n_completed == 0) return SEEDRL_OK;
  store_set_index_kernel<<<ceil_div(n_completed, 128), 128, 0, (cudaStream_t)stream>>>(
      index, completed_ids, n_completed, 1 + overlap);   // utils.py:254-255
  count_launch(PC_MISC, (cudaStream_t)stream);
  SEEDRL_CHECK_LAUNCH();
  return SEEDRL_OK;
}

extern "C" int seedrl_store_reset(uint8_t* state, int32_t* index, const int32_t* env_ids, int n,
                                  int full_length, size_t row_bytes, int overlap,
                                  seedrl_stream_t stream) {
  if (n == 0) return SEEDRL_OK;
  SEEDRL_CHECK_ARG(env_ids, "bad argument");
  if (index) {
    store_set_index_kernel<<<ceil_div(n, 128), 128, 0, (cudaStream_t)stream>>>(index, env_ids, n,
                                                                               overlap);  // :207-208
    count_launch(PC_MISC, (cudaStream_t)stream);
    SEEDRL_CHECK_LAUNCH();
  }
  if (state && overlap > 0 && row_bytes > 0) {                                            // :212-225
    store_zero_rows_kernel<<<dim3(n, overlap), 128, 0, (cudaStream_t)stream>>>(state, env_ids,
                                                                              full_length, row_bytes);
    count_launch(PC_MISC, (cudaStream_t)stream);
    SEEDRL_CHECK_LAUNCH();
  }
  return SEEDRL_OK;
}
