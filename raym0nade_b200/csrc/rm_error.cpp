// rm_error.cpp — last-error storage for the C ABI (thread-local, no exceptions cross the boundary).
#include "raym0nade_b200.h"
#include "rm_internal.h"

static thread_local char g_err[512] = "";

int rm_fail(int code, const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

extern "C" const char *rm_last_error(void) { return g_err; }
extern "C" const char *rm_version(void) { return "raym0nade_b200 0.1 (sm_100a)"; }
