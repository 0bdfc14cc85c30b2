// Library-wide state of libseedrl_b200.so: last-error string, launch counter, ABI version,
// and the optional per-category kernel timing used by bench.py's profiling pass.
#include <vector>

#include "common.cuh"

namespace seedrl {
thread_local std::string g_last_error;
std::atomic<uint64_t> g_launch_count{0};
bool g_prof_on = false;
int g_conv_cat = PC_CONV_FWD;

struct ProfRec { int cat; cudaEvent_t e; };
static std::vector<ProfRec> g_prof;
static cudaEvent_t g_prof_start;

void prof_mark_(int cat, cudaStream_t st) {
  ProfRec r;
  r.cat = cat;
  cudaEventCreate(&r.e);
  cudaEventRecord(r.e, st);
  g_prof.push_back(r);
}
}  // namespace seedrl

static const char* kProfNames[seedrl::PC_COUNT] = {
    "conv3x3_fwd", "conv3x3_dgrad", "conv3x3_wgrad", "maxpool", "sgemm", "lstm_pointwise",
    "vtrace_loss", "adam", "vtrace", "misc"};

extern "C" const char* seedrl_last_error(void) { return seedrl::g_last_error.c_str(); }
extern "C" int seedrl_abi_version(void) { return 1; }
extern "C" uint64_t seedrl_kernel_launch_count(void) {
  return seedrl::g_launch_count.load(std::memory_order_relaxed);
}

extern "C" int seedrl_profile_num_categories(void) { return seedrl::PC_COUNT; }
extern "C" const char* seedrl_profile_category_name(int i) {
  return (i >= 0 && i < seedrl::PC_COUNT) ? kProfNames[i] : "";
}
extern "C" int seedrl_profile_begin(seedrl_stream_t stream) {
  seedrl::g_prof.clear();
  cudaEventCreate(&seedrl::g_prof_start);
  cudaEventRecord(seedrl::g_prof_start, (cudaStream_t)stream);
  seedrl::g_prof_on = true;
  return SEEDRL_OK;
}
extern "C" int seedrl_profile_end(double* ms_per_category, uint64_t* launches_per_category) {
  seedrl::g_prof_on = false;
  if (!ms_per_category || !launches_per_category)
    return seedrl::set_error(SEEDRL_ERR_INVALID_ARGUMENT, "seedrl_profile_end: null pointer");
  if (cudaDeviceSynchronize() != cudaSuccess)
    return seedrl::set_error(SEEDRL_ERR_INTERNAL, "seedrl_profile_end: device sync failed");
  for (int i = 0; i < seedrl::PC_COUNT; ++i) {
    ms_per_category[i] = 0;
    launches_per_category[i] = 0;
  }
  cudaEvent_t prev = seedrl::g_prof_start;
  for (auto& r : seedrl::g_prof) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, prev, r.e) == cudaSuccess) {
      ms_per_category[r.cat] += ms;
      launches_per_category[r.cat] += 1;
    }
    cudaEventDestroy(prev);
    prev = r.e;
  }
  cudaEventDestroy(prev);
  seedrl::g_prof.clear();
  return SEEDRL_OK;
}
