// Library-wide entry points: version and thread-local error message.
#include <stdarg.h>
#include <string.h>

#include "common.cuh"

static thread_local char g_last_error[512] = "";

void hm_set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_last_error, sizeof(g_last_error), fmt, ap);
    va_end(ap);
}

extern "C" {
int hm_version(void) {
 return HM_VERSION; }
const char *hm_last_error(void) { return g_last_error; }
}
