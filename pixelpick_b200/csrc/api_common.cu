// Error state, version and device query of the C-ABI (include/pixelpick_b200.h).
#include "pp_common.cuh"
#include <string.h>

namespace pp {
static thread_local char g_err[512] = "";
std::atomic<long long> g_launches{0};
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
}  // namespace pp

extern "C" {
int pp_version(void) { return 100; }
long long pp_launch_count(void) { return pp::g_launches.load(std::memory_order_relaxed); }
const char* pp_last_error(void) { return pp::g_err; }
/* host-side views of the two ordering maps of the radix select (pp_common.cuh), for the CPU property tests: the select is
 * exact iff bucket0 is monotone in ord_key */
unsigned int pp_host_ord_key(float score, int largest) { return pp::ord_key(score, largest != 0); }
unsigned int pp_host_bucket0(float score, int largest) { return pp::bucket0(score, largest != 0); }
int pp_device_info(int* sm_count, int* cc_major, int* cc_minor, char* name, int name_len) {
  int dev = 0;
  PP_CUDA(cudaGetDevice(&dev));
  cudaDeviceProp prop;
  PP_CUDA(cudaGetDeviceProperties(&prop, dev));
  if (sm_count) *sm_count = prop.multiProcessorCount;
  if (cc_major) *cc_major = prop.major;
  if (cc_minor) *cc_minor = prop.minor;
  if (name && name_len > 0) {
    strncpy(name, prop.name, (size_t)name_len - 1);
    name[name_len - 1] = 0;
  }
  return PP_OK;
}
}
