// Shared host/device helpers for libcloops_b200 (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <atomic>
#include <string>
#include <vector>

#include "../../include/cloops_b200.h"

typedef unsigned long long u64;
typedef unsigned int u32;

namespace cloops {

// ---- error plumbing ------------------------------------------------------------------------------
extern thread_local std::string g_err;
extern std::atomic<long long> g_launches;
extern bool g_debug_sync;   // CLOOPS_DEBUG_SYNC=1: synchronise and check after every kernel
int fail(int code, const char* fmt, ...);

#define CU_TRY(expr)                                                                                  \
    do {                                                                                              \
        cudaError_t _e = (expr);                                                                      \
        if (_e != cudaSuccess)                                                                        \
            return cloops::fail(CLOOPS_ECUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), \
                                __FILE__, __LINE__);                                                  \
    } while (0)

#define RET_IF(expr)            \
    do {                        \
        int _r = (expr);        \
        if (_r != 0) return _r; \
    } while (0)

// kernel launch with launch counting + error check
#define LAUNCH(kernel, grid, block, smem, stream, ...)                                   \
    do {                                                                                 \
        kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__);                      \
        cloops::g_launches.fetch_add(1, std::memory_order_relaxed);                      \
        CU_TRY(cudaGetLastError());                                                      \
        if (cloops::g_debug_sync) {                                                      \
            cudaError_t _s = cudaStreamSynchronize(stream);                              \
            if (_s != cudaSuccess)                                                       \
                return cloops::fail(CLOOPS_ECUDA, "kernel %s faulted: %s (%s:%d)", #kernel, cudaGetErrorString(_s), \
                                    __FILE__, __LINE__);                                 \
        }                                                                                \
    } while (0)

static inline int cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }

// ---- stage profiling -----------------------------------------------------------------------------
struct StageRec {
    const char* name;
    cudaEvent_t ev;
    float ms;
};
extern bool g_profiling;
extern thread_local std::vector<StageRec> g_stages;
void stages_begin(cudaStream_t s);            // clears, records "start"
void stage_mark(const char* name, cudaStream_t s);  // records end of the stage `name`
int stages_end(cudaStream_t s);               // syncs + resolves ms (profiling only)

// ---- temporary memory ----------------------------------------------------------------------------
// A pass allocates some thirty scratch arrays whose lifetime is one call.  Each (host thread, device, stream) owns a
// workspace: one device block, handed out by a bump pointer and rewound when the Temp that took the memory goes out of
// scope (Temps are stack objects, so they nest).  All work that touches the memory is ordered on that stream, which is
// what makes the rewind safe without a synchronisation.  A request that does not fit opens a further block; when the
// outermost Temp ends the blocks are replaced by one of the size the call needed, so steady-state calls make no
// allocator call at all.  CLOOPS_ARENA=0 goes back to one cudaMallocAsync / cudaFreeAsync per array
// (default pool, release threshold lifted), which is also what every buffer that outlives a call uses.
int pool_init();
struct Arena;
struct ArenaMark {
    int chunk;
    size_t off, used;
};
Arena* arena_get(cudaStream_t s);                                  // nullptr: workspaces are switched off
enum { WS_MISC = 0, WS_PASS = 1 };                                 // WS_PASS: the outermost Temp of a whole per-chromosome pass
ArenaMark arena_enter(Arena* a, cudaStream_t s, int kind);
void arena_leave(Arena* a, const ArenaMark& m, cudaStream_t s);
int arena_alloc(Arena* a, size_t bytes, void** out, cudaStream_t s);

struct Temp {
    cudaStream_t s;
    Arena* a;
    ArenaMark m;
    std::vector<void*> ptrs;
    explicit Temp(cudaStream_t st, int kind = WS_MISC) : s(st), a(arena_get(st)) {
        if (a) m = arena_enter(a, st, kind);
    }
    Temp(const Temp&) = delete;
    Temp& operator=(const Temp&) = delete;
    ~Temp() {
        if (a) arena_leave(a, m, s);
        for (void* p : ptrs) cudaFreeAsync(p, s);
    }
    template <class T>
    int alloc(T** out, size_t count) {
        void* p = nullptr;
        size_t bytes = count * sizeof(T);
        if (bytes == 0) bytes = 16;
        if (a) {
            RET_IF(arena_alloc(a, bytes, &p, s));
            *out = (T*)p;
            return 0;
        }
        cudaError_t e = cudaMallocAsync(&p, bytes, s);
        if (e != cudaSuccess) return fail(CLOOPS_ENOMEM, "cudaMallocAsync(%zu) failed: %s", bytes, cudaGetErrorString(e));
        ptrs.push_back(p);
        *out = (T*)p;
        return 0;
    }
};

// ---- packed (strip, u, vmod) key layout ------------------------------------------------------------
// key = [flag:1][strip:bs][u':bu][vmod:be]   (flag = core point, bit 63; strip = floor(v'/eps);
// u' = (X-Y) - ubase ; v' = (X+Y) - vbase ; ubase, vbase multiples of eps so floor cells are preserved)
struct GridParams {
    int eps;
    int be, bu, bs;        // bit widths
    int sshift;            // be + bu
    u32 emask, umask;      // (1<<be)-1, (1<<bu)-1
    int ubase, vbase;      // multiples of eps
    int ns;                // occupied strip range: strips 0..ns-1 ; strip ns = sentinel for inactive rows
    int n;                 // rows
    int n_act;             // active rows (after cut)
};

#define CORE_FLAG (1ull << 63)
#define KEY_MASK (~CORE_FLAG)

}  // namespace cloops
