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

// Batch-invariant mode: the crop frontier is evaluated in batches whose size depends on scheduling (speculation, sharding over
// GPUs), but every crop's result must be a pure function of the crop - otherwise two ranks (or two batch sizes) would walk
// different search trajectories through near-ties.  With the flag set, entry points that would pick a kernel FAMILY by the row
// count (skinny decode GEMMs for M <= 16, block-per-row RMSNorm for <= 32 rows) always take the family used by large batches,
// whose per-element arithmetic does not depend on M (tests/test_kernels_gpu.py::test_rows_invariant_to_batch).
static int g_batch_invariant = 0;
int vsb_batch_invariant() { return g_batch_invariant; }
extern "C" int vsb_set_batch_invariant(int on) {
  g_batch_invariant = on ? 1 : 0;
  return VSB_OK;
}
