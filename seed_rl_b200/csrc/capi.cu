// Library-wide state of libseedrl_b200.so: last-error string, launch counter, ABI version.
#include "common.cuh"

namespace seedrl {
thread_local std::string g_last_error;
std::atomic<uint64_t> g_launch_count{0};
}  // namespace seedrl

extern "C" const char* seedrl_last_error(void) { return seedrl::g_last_error.c_str(); }
extern "C" int seedrl_abi_version(void) { return 1; }
extern "C" uint64_t seedrl_kernel_launch_count(void) {
  return seedrl::g_launch_count.load(std::memory_order_relaxed);
}
