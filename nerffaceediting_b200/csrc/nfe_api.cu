// Library-level entry points: version, error text, launch counter.
#include <atomic>
#include <stdarg.h>

#include "nfe_common.cuh"

namespace nfe {

static thread_local char g_error[512] = "";
static std::atomic<uint64_t> g_launches{0};

void set_error(const char* fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_error, sizeof(g_error), fmt, ap);
    va_end(ap);
}

void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

}  // namespace nfe

NFE_EXPORT int nfe_version(void) { return 1; }
NFE_EXPORT const char* nfe_last_error(void) { return nfe::g_error; }
NFE_EXPORT uint64_t nfe_launch_count(void) { return nfe::g_launches.load(std::memory_order_relaxed); }
