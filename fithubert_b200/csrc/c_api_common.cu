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

static int g_pdl = -1;  // -1: not decided yet; 0 off, 1 every kernel, 2 (default) only launches hinted as small
static thread_local bool g_pdl_small = false;
bool fhb_pdl_enabled() {
  if (g_pdl < 0) {
    const char* e = getenv("FHB_PDL");
    g_pdl = e == nullptr ? 2 : (e[0] == '2' ? 2 : (e[0] == '0' ? 0 : 1));
  }
  const bool small = g_pdl_small;
  g_pdl_small = false;  // one-shot: consumed by the launch that follows the hint
  return g_pdl == 1 || (g_pdl == 2 && small);
}
// Host functions call this right before launching: `small` = the kernel runs for a few microseconds, so hiding its
// launch latency and prologue behind the predecessor's tail is worth more than the co-residency it costs (mode 2).
void fhb_pdl_hint(bool small) { g_pdl_small = small; }
// Runtime switch for programmatic dependent launch (returns the previous setting: 0 off, 1 all, 2 small kernels only).
// Per-kernel CUDA-event timing turns it off so that an event pair brackets the WHOLE kernel, prologue included.
extern "C" int fhb_set_pdl(int mode) {
  if (g_pdl < 0) fhb_pdl_enabled();
  const int prev = g_pdl;
  g_pdl = mode < 0 ? 0 : (mode > 2 ? 2 : mode);
  return prev;
}

// SMs the persistent kernels leave alone (fhb_num_sms() = device SMs - reserved): while a gradient all-reduce runs
// beside the backward, NCCL's few CTAs get SMs of their own instead of delaying CTAs of a statically scheduled GEMM.
// Process-wide launch configuration like the PDL mode above (the data path itself holds no mutable state).
static int g_reserved_sms = 0;
int fhb_reserved_sms() { return g_reserved_sms; }
extern "C" int fhb_set_reserved_sms(int n) {
  const int prev = g_reserved_sms;
  g_reserved_sms = n < 0 ? 0 : (n > 64 ? 64 : n);
  return prev;
}
