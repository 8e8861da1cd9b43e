// capi_common.cu — error reporting, build info and the launch counter of libgsicp_b200.so.
#include <cstring>
#include "host_common.h"

namespace gsicp {
std::atomic<uint64_t> g_launches{0};
static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
}  // namespace gsicp

extern "C" const char* gsicp_last_error(void) { return gsicp::g_err; }

extern "C" const char* gsicp_build_info(void) {
  return "libgsicp_b200 sm_100a (nvcc " __VERSION__ ", built " __DATE__ ")";
}

extern "C" uint64_t gsicp_launch_count(void) { return gsicp::g_launches.load(); }
