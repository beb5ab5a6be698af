// Error text + version for the C ABI (include/fhb.h).
#include <stdarg.h>

#include "fhb_common.cuh"

static thread_local char g_err[1024] = "";

void fhb_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

extern "C" const char* fhb_last_error(void) { return g_err; }
extern "C" int fhb_abi_version(void) { return 1; }
