// Error reporting, version and launch accounting of libsis_b200.
#include <vector>
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

static unsigned int* g_watchdog = nullptr;
unsigned int* watchdog_word() {
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* p = nullptr;
        if (cudaHostAlloc(&p, sizeof(unsigned int), cudaHostAllocMapped | cudaHostAllocPortable) == cudaSuccess) {
            g_watchdog = (unsigned int*)p;      // unified addressing: the same pointer is valid on every device
            *g_watchdog = 0;
        } else {
            cudaGetLastError();
        }
    }
    return g_watchdog;
}

bool g_prof_on = false;
struct ProfEvent { int cat; cudaEvent_t a, b; };
static std::vector<ProfEvent> g_prof_events;
static std::vector<cudaEvent_t> g_prof_open[PROF_NUM];

void prof_begin(int cat, cudaStream_t stream) {
    cudaEvent_t e;
    if (cudaEventCreate(&e) != cudaSuccess) return;
    cudaEventRecord(e, stream);
    g_prof_open[cat].push_back(e);
}
void prof_end(int cat, cudaStream_t stream) {
    if (g_prof_open[cat].empty()) return;
    cudaEvent_t a = g_prof_open[cat].back();
    g_prof_open[cat].pop_back();
    cudaEvent_t b;
    if (cudaEventCreate(&b) != cudaSuccess) { cudaEventDestroy(a); return; }
    cudaEventRecord(b, stream);
    g_prof_events.push_back({cat, a, b});
}

}  // namespace sis

extern "C" int sis_profile_enable(int on) {
    sis::g_prof_on = on != 0;
    return SIS_OK;
}

extern "C" int sis_profile_collect(double* ms_by_category, uint64_t* count_by_category, int n_categories) {
    using namespace sis;
    for (int i = 0; i < n_categories; ++i) { if (ms_by_category) ms_by_category[i] = 0.0; if (count_by_category) count_by_category[i] = 0; }
    for (auto& e : g_prof_events) {
        float ms = 0.0f;
        if (cudaEventSynchronize(e.b) == cudaSuccess && cudaEventElapsedTime(&ms, e.a, e.b) == cudaSuccess && e.cat < n_categories) {
            if (ms_by_category) ms_by_category[e.cat] += ms;
            if (count_by_category) count_by_category[e.cat] += 1;
        }
        cudaEventDestroy(e.a); cudaEventDestroy(e.b);
    }
    g_prof_events.clear();
    return SIS_OK;
}

extern "C" unsigned int sis_watchdog_code(void) {
    unsigned int* w = sis::g_watchdog;
    return w ? *(volatile unsigned int*)w : 0u;
}
extern "C" void sis_watchdog_clear(void) {
    if (sis::g_watchdog) *(volatile unsigned int*)sis::g_watchdog = 0u;
}

extern "C" const char* sis_last_error(void) { return sis::g_error; }
extern "C" int sis_version(void) { return 120; }   // 1.2: contour stage, native PNG writer (round 2)
extern "C" uint64_t sis_launch_count(void) { return (uint64_t)sis::g_launches.load(); }
