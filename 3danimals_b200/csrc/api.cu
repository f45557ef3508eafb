// api.cu - version + thread-local error reporting of libb2a.so (SURVEY.md §8b error convention: return code + message).
#include <stdarg.h>

#include "common.cuh"

namespace {
thread_local char g_error[512] = "";
}

void b2a_set_error(const char* fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_error, sizeof(g_error), fmt, ap);
    va_end(ap);
}

B2A_API int b2a_version(void) { return B2A_VERSION; }

B2A_API const char* b2a_last_error_string(void) { return g_error; }
