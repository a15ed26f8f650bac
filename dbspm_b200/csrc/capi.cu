// capi.cu -- version string and thread-local error message of the C ABI (include/dbspm_b200.h).
#include "capi_common.h"

static thread_local char g_err[512] = "";

void dbspm_set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

extern "C" const char *dbspm_last_error(void) { return g_err; }
extern "C" const char *dbspm_version(void) { return "dbspm_b200 0.1.0 sm_100a"; }
