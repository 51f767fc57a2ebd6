#include "common.cuh"
#include <string.h>
#include <stdlib.h>

namespace pcls {
static thread_local char g_err[1024] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
static int pdl_from_env() {
  const char* e = getenv("PCLS_PDL");
  return (e && e[0] == '0') ? 0 : 1;
}
int pdl_mode = pdl_from_env();
static long long pdl_early_px_from_env() {
  const char* e = getenv("PCLS_PDL_EARLY_PX");
  return e ? atoll(e) : (1ll << 20);
}
long long pdl_early_px = pdl_early_px_from_env();
thread_local int pdl_early_now = 0;

}  // namespace pcls

extern "C" const char* pcls_last_error(void) { return pcls::g_err; }
extern "C" int pcls_abi_version(void) { return PCLS_ABI_VERSION; }
