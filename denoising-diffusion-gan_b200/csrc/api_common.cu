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
