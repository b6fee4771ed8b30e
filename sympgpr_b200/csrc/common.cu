#include "common.cuh"

#include <stdarg.h>

namespace sgp {

static thread_local char g_err[1024] = "";

void set_error(const char* fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

const char* get_error() { return g_err; }

static unsigned long long g_launches = 0ull;
void count_launch(unsigned long long n) { __atomic_fetch_add(&g_launches, n, __ATOMIC_RELAXED); }
unsigned long long launch_count() { return __atomic_load_n(&g_launches, __ATOMIC_RELAXED); }

}  // namespace sgp
