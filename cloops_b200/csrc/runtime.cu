// Error plumbing, launch counter, stage timing, memory pool for libcloops_b200.
#include <stdarg.h>
#include <stdlib.h>

#include <chrono>
#include <map>
#include <memory>
#include <mutex>

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


// ---- per-(thread, device, stream) workspaces (see common.cuh) ---------------------------------------------------------
struct Arena {
    struct Chunk {
        char* p;
        size_t cap;
    };
    int dev = 0;
    std::vector<Chunk> chunks;
    int cur = 0;               // block the bump pointer is in
    size_t off = 0;            // bump pointer inside chunks[cur]
    size_t used = 0;           // bytes handed out and not yet rewound (with the slack left at block ends)
    size_t high = 0;           // largest `used` since the outermost Temp began
    int depth = 0;
    int kind = 0;              // what the outermost Temp of the current call said it is for (WS_*)
};

static const bool g_arena_on = !(getenv("CLOOPS_ARENA") && getenv("CLOOPS_ARENA")[0] == '0');
static const bool g_arena_trace = getenv("CLOOPS_TRACE") != nullptr;
static const size_t ARENA_ALIGN = 256, ARENA_MIN_CHUNK = 64u << 20;
static std::mutex g_arena_mutex;                                   // guards the registry
static std::vector<Arena*> g_arenas;                               // all workspaces of the process (cloops_workspace_release)

struct ArenaSet {                                                  // the calling thread's workspaces
    std::map<std::pair<int, cudaStream_t>, Arena*> by_stream;
    ~ArenaSet() {
        std::lock_guard<std::mutex> l(g_arena_mutex);
        for (auto& kv : by_stream) {
            Arena* a = kv.second;
            for (Arena::Chunk& c : a->chunks)
                if (cudaFree(c.p) != cudaSuccess) cudaGetLastError();          // the runtime may already be gone at exit
            a->chunks.clear();
            for (size_t k = 0; k < g_arenas.size(); ++k)
                if (g_arenas[k] == a) { g_arenas.erase(g_arenas.begin() + k); break; }
            delete a;
        }
    }
};
static thread_local ArenaSet g_my_arenas;

Arena* arena_get(cudaStream_t s) {
    if (!g_arena_on) return nullptr;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    Arena*& a = g_my_arenas.by_stream[std::make_pair(dev, s)];
    if (!a) {
        a = new Arena();
        a->dev = dev;
        std::lock_guard<std::mutex> l(g_arena_mutex);
        g_arenas.push_back(a);
    }
    return a;
}

// largest peak any whole-pass workspace (WS_PASS) of the process has seen, per device: a thread that has only served small
// chromosomes so far sizes its block for the largest one before it meets it, so that all blocks reach their final size in
// the first round instead of one at a time over many calls (a late gigabyte-sized cudaMallocAsync was measured at 0.4 s)
static std::atomic<size_t> g_arena_peak[64];

static size_t arena_target(const Arena* a) {
    if (a->kind != WS_PASS) return 0;
    const size_t peak = g_arena_peak[a->dev & 63].load(std::memory_order_relaxed);
    return peak + peak / 4;
}

ArenaMark arena_enter(Arena* a, cudaStream_t s, int kind) {
    if (a->depth++ == 0) {
        a->kind = kind;
        a->high = a->used;
        const size_t want = arena_target(a);
        if (a->chunks.size() <= 1 && want > (a->chunks.empty() ? 0 : a->chunks[0].cap) && want >= ARENA_MIN_CHUNK) {
            const auto t0 = std::chrono::steady_clock::now();
            if (!a->chunks.empty()) cudaFreeAsync(a->chunks[0].p, s);
            a->chunks.clear();
            void* p = nullptr;
            if (cudaMallocAsync(&p, want, s) == cudaSuccess) a->chunks.push_back(Arena::Chunk{(char*)p, want});
            else cudaGetLastError();
            if (g_arena_trace)
                fprintf(stderr, "[cloops] workspace %p: block of %zu MB for the process-wide peak in %.2f ms\n", (void*)a, want >> 20,
                        std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count());
        }
    }
    return ArenaMark{a->cur, a->off, a->used};
}

int arena_alloc(Arena* a, size_t bytes, void** out, cudaStream_t s) {
    bytes = (bytes + ARENA_ALIGN - 1) / ARENA_ALIGN * ARENA_ALIGN;
    if (a->chunks.empty() || a->off + bytes > a->chunks[a->cur].cap) {
        // the rest of the current block stays unused until the rewind; take a later block that fits, else add one
        int k = a->chunks.empty() ? 0 : a->cur + 1;
        while (k < (int)a->chunks.size() && a->chunks[k].cap < bytes) ++k;
        if (k == (int)a->chunks.size()) {
            size_t cap = a->chunks.empty() ? ARENA_MIN_CHUNK : 2 * a->chunks.back().cap;
            if (cap < bytes) cap = bytes;
            if (cap < arena_target(a)) cap = arena_target(a);
            void* p = nullptr;
            const auto t0 = std::chrono::steady_clock::now();
            cudaError_t e = cudaMallocAsync(&p, cap, s);
            if (e != cudaSuccess) return fail(CLOOPS_ENOMEM, "workspace block of %zu bytes: %s", cap, cudaGetErrorString(e));
            if (g_arena_trace)
                fprintf(stderr, "[cloops] workspace %p: block %d of %zu MB in %.2f ms\n", (void*)a, (int)a->chunks.size(), cap >> 20,
                        std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count());
            a->chunks.push_back(Arena::Chunk{(char*)p, cap});
        }
        a->cur = k;
        a->off = 0;
    }
    *out = a->chunks[a->cur].p + a->off;
    a->off += bytes;
    a->used += bytes;
    if (a->used > a->high) {
        a->high = a->used;
        if (a->kind == WS_PASS) {
            std::atomic<size_t>& peak = g_arena_peak[a->dev & 63];
            size_t seen = peak.load(std::memory_order_relaxed);
            while (a->high > seen && !peak.compare_exchange_weak(seen, a->high, std::memory_order_relaxed)) {}
        }
    }
    return 0;
}

void arena_leave(Arena* a, const ArenaMark& m, cudaStream_t s) {
    a->cur = m.chunk;
    a->off = m.off;
    a->used = m.used;
    if (--a->depth > 0 || a->chunks.size() <= 1) return;
    // the call needed several blocks: one block that holds its peak from now on (freed and allocated in stream order,
    // like the work that used them)
    size_t want = a->high + a->high / 4;
    if (want < arena_target(a)) want = arena_target(a);
    const auto t0 = std::chrono::steady_clock::now();
    for (Arena::Chunk& c : a->chunks) cudaFreeAsync(c.p, s);
    a->chunks.clear();
    a->cur = 0;
    a->off = a->used = 0;
    void* p = nullptr;
    if (cudaMallocAsync(&p, want, s) == cudaSuccess) a->chunks.push_back(Arena::Chunk{(char*)p, want});
    else cudaGetLastError();                                       // the next request starts from an empty list
    if (g_arena_trace)
        fprintf(stderr, "[cloops] workspace %p: one block of %zu MB in %.2f ms\n", (void*)a, want >> 20,
                std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count());
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
int cloops_workspace_release(void) {
    // frees every workspace block of the process; the caller guarantees that no call is in flight on any thread
    std::lock_guard<std::mutex> l(g_arena_mutex);
    int keep = 0;
    if (cudaGetDevice(&keep) != cudaSuccess) { cudaGetLastError(); return 0; }
    for (Arena* a : g_arenas) {
        if (a->chunks.empty()) continue;
        if (a->depth != 0) return fail(CLOOPS_EINVAL, "a workspace is in use");
        cudaSetDevice(a->dev);
        CU_TRY(cudaDeviceSynchronize());
        for (Arena::Chunk& c : a->chunks) CU_TRY(cudaFree(c.p));
        a->chunks.clear();
        a->cur = 0;
        a->off = a->used = 0;
    }
    for (std::atomic<size_t>& p : g_arena_peak) p.store(0);
    cudaSetDevice(keep);
    return 0;
}
float cloops_stage_ms(int i) { return (i >= 0 && i + 1 < (int)g_stages.size()) ? g_stages[i + 1].ms : 0.f; }
}
