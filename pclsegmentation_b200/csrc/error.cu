#include "common.cuh"
#include <string.h>

namespace pcls {
static thread_local char g_err[1024] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
}  // namespace pcls

extern "C" const char* pcls_last_error(void) { return pcls::g_err; }
extern "C" int pcls_abi_version(void) { return PCLS_ABI_VERSION; }
