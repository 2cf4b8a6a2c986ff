// error reporting + version for the vstar_b200 C-ABI
#include "common.cuh"
#include "vstar_b200.h"

#include <stdarg.h>

static thread_local char g_err[512] = "";

void vsb_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

extern "C" const char* vsb_last_error(void) { return g_err; }
extern "C" int vsb_version(void) { return 100; }
