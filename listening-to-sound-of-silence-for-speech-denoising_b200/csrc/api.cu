// Error state and version of libsos_b200.so.
#include "common.cuh"
#include "sos_b200.h"
#include <stdarg.h>

namespace {
thread_local char g_err[1024] = "";
}

void sos_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

extern "C" const char* sos_last_error(void) { return g_err; }
extern "C" int sos_version(void) { return 110; }
