// Library-wide state of the tamtr_b200 C ABI: error string, launch counter, version.
#include <stdarg.h>

#include <atomic>

#include "common.cuh"

namespace tamtr {

static thread_local char g_err[512] = "";
static std::atomic<unsigned long long> g_launches{0};

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

void count_launch(unsigned n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

}  // namespace tamtr

extern "C" int tamtr_abi_version(void) { return TAMTR_B200_ABI_VERSION; }
extern "C" const char *tamtr_last_error(void) { return tamtr::g_err; }
extern "C" unsigned long long tamtr_launch_count(void) { return tamtr::g_launches.load(); }
