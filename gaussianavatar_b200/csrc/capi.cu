// Process-wide plumbing of the C ABI: error string, version, launch counter.
#include <atomic>
#include <cstdarg>

#include "common.cuh"

namespace ga {
namespace {
thread_local char g_err[512] = "";
std::atomic<long long> g_launches{0};
}  // namespace

void set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
}  // namespace ga

extern "C" int ga_version(void) { return 100; }
extern "C" const char *ga_last_error(void) { return ga::g_err; }
extern "C" long long ga_launch_count(void) { return ga::g_launches.load(std::memory_order_relaxed); }
