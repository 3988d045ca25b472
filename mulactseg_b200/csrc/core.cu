// Error reporting and version query of the C ABI.
#include "common.cuh"

#include <stdarg.h>
#include <atomic>

namespace mas {

static thread_local char g_error[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_error, sizeof(g_error), fmt, ap);
    va_end(ap);
}

static std::atomic<long long> g_launches{0};

void count_launches(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

}  // namespace mas

extern "C" int mas_abi_version(void) { return MAS_ABI_VERSION; }

extern "C" const char* mas_last_error(void) { return mas::g_error; }

extern "C" int64_t mas_kernel_launches(void) { return mas::g_launches.load(std::memory_order_relaxed); }
