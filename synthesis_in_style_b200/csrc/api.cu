// Error reporting, version and launch accounting of libsis_b200.
#include "common.cuh"

namespace sis {

static thread_local char g_error[1024] = "";
std::atomic<unsigned long long> g_launches{0};

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_error, sizeof(g_error), fmt, ap);
    va_end(ap);
}

}  // namespace sis

extern "C" const char* sis_last_error(void) { return sis::g_error; }
extern "C" int sis_version(void) { return 100; }
extern "C" uint64_t sis_launch_count(void) { return (uint64_t)sis::g_launches.load(); }
