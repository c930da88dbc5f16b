// C-ABI plumbing: error reporting and version.
#include "common.cuh"
#include "../../include/jkb200.h"

static thread_local char g_err[1024] = "";

void jk_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

extern "C" const char* jk_last_error(void) { return g_err; }
extern "C" int jk_version(void) { return 100; }
