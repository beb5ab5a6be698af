// Error text + version for the C ABI (include/fhb.h).
#include <stdarg.h>
#include <stdlib.h>

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

static int g_pdl = -1;  // -1: not decided yet (off unless FHB_PDL is set in the environment)
bool fhb_pdl_enabled() {
  if (g_pdl < 0) g_pdl = getenv("FHB_PDL") != nullptr ? 1 : 0;
  return g_pdl != 0;
}
// Runtime switch for programmatic dependent launch (returns the previous setting).  Per-kernel CUDA-event timing
// turns it off so that an event pair brackets the WHOLE kernel, prologue included.
extern "C" int fhb_set_pdl(int enabled) {
  const int prev = fhb_pdl_enabled() ? 1 : 0;
  g_pdl = enabled ? 1 : 0;
  return prev;
}
