// Error plumbing, launch counter, stage timing, memory pool for libcloops_b200.
#include <stdarg.h>
#include <stdlib.h>

#include "common.cuh"

namespace cloops {

thread_local std::string g_err;
std::atomic<long long> g_launches{0};
bool g_profiling = false;
bool g_debug_sync = getenv("CLOOPS_DEBUG_SYNC") != nullptr && getenv("CLOOPS_DEBUG_SYNC")[0] == '1';
thread_local std::vector<StageRec> g_stages;
static thread_local std::vector<cudaEvent_t> g_event_cache;

int fail(int code, const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}

static cudaEvent_t get_event() {
    if (!g_event_cache.empty()) {
        cudaEvent_t e = g_event_cache.back();
        g_event_cache.pop_back();
        return e;
    }
    cudaEvent_t e;
    cudaEventCreate(&e);
    return e;
}

void stages_begin(cudaStream_t s) {
    for (auto& r : g_stages) g_event_cache.push_back(r.ev);
    g_stages.clear();
    if (!g_profiling) return;
    StageRec r{"start", get_event(), 0.f};
    cudaEventRecord(r.ev, s);
    g_stages.push_back(r);
}

void stage_mark(const char* name, cudaStream_t s) {
    if (!g_profiling) return;
    StageRec r{name, get_event(), 0.f};
    cudaEventRecord(r.ev, s);
    g_stages.push_back(r);
}

int stages_end(cudaStream_t s) {
    if (!g_profiling) return 0;
    CU_TRY(cudaStreamSynchronize(s));
    for (size_t i = 1; i < g_stages.size(); ++i) cudaEventElapsedTime(&g_stages[i].ms, g_stages[i - 1].ev, g_stages[i].ev);
    return 0;
}

int pool_init() {
    static thread_local int done_dev = -1;
    int dev = 0;
    CU_TRY(cudaGetDevice(&dev));
    if (done_dev == dev) return 0;
    cudaMemPool_t pool;
    CU_TRY(cudaDeviceGetDefaultMemPool(&pool, dev));
    unsigned long long thr = ~0ull;
    CU_TRY(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr));
    done_dev = dev;
    return 0;
}

}  // namespace cloops

using namespace cloops;

extern "C" {
const char* cloops_last_error(void) { return g_err.c_str(); }
const char* cloops_version(void) { return "cloops_b200 0.1 (sm_100a)"; }
int64_t cloops_kernel_launches(void) { return (int64_t)g_launches.load(); }
void cloops_set_profiling(int on) { g_profiling = on != 0; }
int cloops_stage_count(void) { return g_stages.empty() ? 0 : (int)g_stages.size() - 1; }
const char* cloops_stage_name(int i) { return (i >= 0 && i + 1 < (int)g_stages.size()) ? g_stages[i + 1].name : ""; }
float cloops_stage_ms(int i) { return (i >= 0 && i + 1 < (int)g_stages.size()) ? g_stages[i + 1].ms : 0.f; }
}
