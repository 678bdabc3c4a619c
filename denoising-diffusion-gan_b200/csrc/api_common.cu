// Error plumbing shared by every entry point of libddgan_b200.so.
#include "common.cuh"
#include "ddgan_b200.h"
#include <string.h>

static thread_local char g_err[512] = "";

void ddg_set_last_error(const char* msg) {
  strncpy(g_err, msg ? msg : "", sizeof(g_err) - 1);
  g_err[sizeof(g_err) - 1] = 0;
}

extern "C" const char* ddg_last_error(void) { return g_err; }
extern "C" int ddg_version(void) { return 100; }

// Programmatic dependent launch of the kernels that support it (common.cuh): off unless DDG_PDL=1 / ddg_set_pdl(1).  Measured on the
// graphed CIFAR-10 sampling loop and train step: no difference beyond run-to-run noise (the graph already issues kernels back to back).
#include <stdlib.h>
static int g_pdl = -1;
int ddg_pdl_enabled(void) {
  if (g_pdl < 0) { const char* e = getenv("DDG_PDL"); g_pdl = (e && e[0] == '1') ? 1 : 0; }
  return g_pdl;
}
extern "C" int ddg_set_pdl(int on) { const int old = ddg_pdl_enabled(); g_pdl = on ? 1 : 0; return old; }
