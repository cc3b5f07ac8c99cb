// Labels of cDBSCAN (v1) and cDBSCAN2 (v2) over a built index, as order-free data-parallel rules.
//
// Both reference classes share the core set ( n(p) >= minPts, Manhattan, inclusive, self counted:
// cDBSCAN.py:163-166,196-204 ; cDBSCAN2.py:333-334 ) and the components of the core graph; they differ
// in cluster numbering, border ownership and survival:
//   v1 (cDBSCAN.py:128-184): ids ascend with the smallest ROW of a core point of the component (its
//       "seed"); a border point goes to the LARGEST id among clusters whose seed is within eps
//       (:172-173 relabels every neighbour of a new seed), else to the SMALLEST adjacent id (:179-182);
//       clusters that end with < minPts members are removed, ids keep their gaps (:149-152).
//   v2 (cDBSCAN2.py:114-192): clusters are attempted in order of the first-inserted rotated cell that
//       holds one of their core points (:117-140); a border point goes to the lowest-ranked ALIVE
//       adjacent cluster (only label -1 points are ever collected); a cluster with < minPts members is
//       released (:180-185) and its points fall to later clusters; ids are dense over survivors.
#include <limits.h>
#include <stdlib.h>

#include "index.cuh"
#include "scan.cuh"

namespace cloops {

static const bool g_trace = getenv("CLOOPS_TRACE") != nullptr;
static thread_local int g_trace_rounds = 0;

enum : unsigned char { ST_NONE = 0, ST_ALIVE = 1, ST_DEAD = 2, ST_UNDECIDED = 3 };

struct Work {
    int* cnt;        // [n_act] neighbour counts (saturated at minPts)
    int* parent;     // [n_act] union-find over sorted indices (core points only)
    int* rank;       // [n_act] per root: v1 = min row of core point ; v2 = min cell-first-row
    int* ncore;      // [n_act] per root: core points
    int* assigned;   // [n_act] root the point belongs to (or -1)
    int* chead;      // [n_act] v2: sorted index of the first point of the point's rotated cell
    int* cellmin;    // [n_act] v2: per cell head: smallest row in the cell
    int* size;       // [n_act] v1: members per root ; v2: lb
    int* ub;         // [n_act] v2: ub
    unsigned char* status;  // [n_act] per root
    int* flags;      // [n+1] rank-indexed survivor flags, then their exclusive scan
    int* ids;        // [n+1]
    int* list_und;   // [n_act] undecided roots
    int* list_con;   // [n_act] contested border points
    int* counters;   // [8]: 0 n_und, 1 n_con, 2 remaining, 3 n_dead, 4 n_comp
    int* slots_core; // [CTR_SLOTS*CTR_STRIDE] spread partial sums of n_core
    int* slots_lab;  // same for n_labelled
    u32* corebits;   // [CB_WORDS] hashed bitmap of the rotated cells that hold a core point (border pruning)
};

// A non-core point can only be a border point if one of the 3 x 3 rotated floor cells around it holds a core point
// (|du|, |dv| <= eps).  Core points mark their cell in a hashed bitmap (2^26 bits = 8 MB, L2-resident); the border kernels
// probe the nine cells first and skip the neighbour walk when none is marked.  A hash collision can only cause a needless
// walk, never a missed one.  At Hi-C depth (minPts 20-50) core points are confined to the clusters, so most PETs skip.
#define CB_BITS 26
#define CB_WORDS (1u << (CB_BITS - 5))
__device__ __forceinline__ u32 cb_hash(u32 strip, u32 cu) { return (strip * 0x9E3779B1u + cu * 0x85EBCA77u) >> (32 - CB_BITS); }
__device__ __forceinline__ bool cb_any_core_near(const u32* __restrict__ bits, u32 strip, u32 cu) {
    bool any = false;
#pragma unroll
    for (int ds = -1; ds <= 1; ++ds)
#pragma unroll
        for (int dc = -1; dc <= 1; ++dc) {
            const u32 h = cb_hash(strip + (u32)ds, cu + (u32)dc);
            any |= (__ldg(bits + (h >> 5)) >> (h & 31u)) & 1u;
        }
    return any;
}

// Informational totals (n_core, n_labelled): one atomic per CTA, spread over CTR_SLOTS addresses 128 B
// apart -- the L2 atomic unit serialises same-address atomics (~0.5 us per thousand), which made a
// per-warp counter the slowest part of otherwise streaming kernels.
#define CTR_SLOTS 32
#define CTR_STRIDE 32
__device__ __forceinline__ void block_count(int* slots, bool pred) {
    const int total = __syncthreads_count(pred);
    if (threadIdx.x == 0 && total) atomicAdd(&slots[(blockIdx.x % CTR_SLOTS) * CTR_STRIDE], total);
}

// CTA-level compaction: the threads whose predicate holds are queued (by local thread id) so that the
// first *s_n threads of the CTA continue with dense warps.  Kernels that only work on core points (or
// only on non-core points) otherwise run with a third of their lanes active.
__device__ __forceinline__ int cta_compact(bool pred, int* q, int* s_n) {
    if (threadIdx.x == 0) *s_n = 0;
    __syncthreads();
    const unsigned b = __ballot_sync(0xffffffffu, pred);
    const int lane = threadIdx.x & 31;
    int base = 0;
    if (lane == 0 && b) base = atomicAdd(s_n, __popc(b));
    base = __shfl_sync(0xffffffffu, base, 0);
    if (pred) q[base + __popc(b & ((1u << lane) - 1))] = threadIdx.x;
    __syncthreads();
    return *s_n;
}

__global__ void __launch_bounds__(256) fill_int_kernel(int* p, int v, long long n) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

// core flag into bit 63 of the key; union-find and per-root state initialised
__global__ void __launch_bounds__(256) flag_kernel(u64* __restrict__ keys, GridParams P, int minPts, Work W, int want_cells) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    bool core = false;
    if (i < P.n_act) {
        u64 k = keys[i] & KEY_MASK;
        core = W.cnt[i] >= minPts;
        keys[i] = core ? (k | CORE_FLAG) : k;
        W.rank[i] = INT_MAX;
        W.ncore[i] = 0;
        W.size[i] = 0;
        W.status[i] = ST_NONE;
        if (core) {
            const u32 h = cb_hash((u32)(k >> P.sshift), ((u32)(k >> P.be) & P.umask) / (u32)P.eps);
            atomicOr(&W.corebits[h >> 5], 1u << (h & 31u));
        }
        if (want_cells) {
            // rotated floor cell = (strip, floor(u'/eps)) (cDBSCAN2.py:69-70); head = first sorted point of the cell
            bool head = true;
            if (i > 0) {
                u64 kp = keys[i - 1] & KEY_MASK;   // flag bit of a neighbour may or may not be set yet: masked
                u32 cu = ((u32)(k >> P.be) & P.umask) / (u32)P.eps;
                u32 cup = ((u32)(kp >> P.be) & P.umask) / (u32)P.eps;
                head = (k >> P.sshift) != (kp >> P.sshift) || cu != cup;
            }
            W.chead[i] = head ? i : 0;
            W.cellmin[i] = INT_MAX;
        }
    }
    block_count(W.slots_core, core);
}

__global__ void __launch_bounds__(256) cellmin_kernel(const u32* __restrict__ rows, GridParams P, Work W) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P.n_act) return;
    atomicMin(&W.cellmin[W.chead[i]], (int)rows[i]);
}

__device__ __forceinline__ int uf_find(int* parent, int x) {
    int p = parent[x];
    while (p != x) {
        int g = parent[p];
        if (g != p) parent[x] = g;    // path halving; always points at an ancestor
        x = p;
        p = g;
    }
    return x;
}

__device__ __forceinline__ void uf_union(int* parent, int a, int b) {
    while (true) {
        a = uf_find(parent, a);
        b = uf_find(parent, b);
        if (a == b) return;
        if (a < b) { int t = a; a = b; b = t; }        // hook the larger root under the smaller
        int old = atomicCAS(&parent[a], a, b);
        if (old == a) return;
    }
}

// Core graph edges.  Every core-core edge joins points of the same strip or of adjacent strips.
//  * same strip: core points within eps in u form CHAINS (consecutive core points of the strip whose u
//    gaps are <= eps).  A chain is one connected set, so it needs no union-find at all: a core point
//    is a chain head iff it has no core point within eps on its left, and every core point's parent is
//    the latest head at or before it -- one inclusive max-scan over the head indices.
//  * strip s-1: one union per chain of strip s-1 that has a member inside the window passing the v test
//    is enough; after linking, the scan jumps to the next chain head (suffix-min scan of head indices),
//    so a core point costs O(#chains in its window), not O(#points) -- dense Hi-C strips hold hundreds
//    of chain members per window.  Only these cross-strip links go through the lock-free union-find,
//    whose trees start out flat (depth 1).
__global__ void __launch_bounds__(256) chain_head_kernel(const u64* __restrict__ keys, const int* __restrict__ sstart, GridParams P,
                                                         int* __restrict__ head, int* __restrict__ head_or_inf) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P.n_act) return;
    const u64 key = keys[i];
    int h = i;
    if (key >> 63) {
        const PointView p = view(key, P);
        const int lo_s = __ldg(sstart + p.s + 1);
        for (int j = i - 1; j >= lo_s; --j) {
            u64 kq = keys[j];
            if (((u32)(kq >> P.be) & P.umask) < p.ulo) break;
            if (kq >> 63) { h = -1; break; }
        }
    } else {
        h = -1;
    }
    head[i] = h < 0 ? 0 : h;
    head_or_inf[i] = h < 0 ? INT_MAX : h;      // input of the suffix-min that yields "next chain head after i"
}

__global__ void __launch_bounds__(256) clamp_next_head_kernel(int* __restrict__ nh, int na) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < na && nh[i] > na) nh[i] = na;
}

// chain[] = immutable chain head per core point (the scan output); seen[hA] remembers the last chain of
// strip s-1 that chain hA was linked to, so the many members of hA that look at the same chain below
// skip the union (and its pointer chasing) after one cached load.
__global__ void __launch_bounds__(256) union_kernel(const u64* __restrict__ keys, const int* __restrict__ sstart, GridParams P,
                                                    const int* __restrict__ chain, const int* __restrict__ next_head,
                                                    int* __restrict__ seen, int* __restrict__ parent) {
    __shared__ int q[256];
    __shared__ int s_n;
    {
        const int i0 = blockIdx.x * blockDim.x + threadIdx.x;
        const bool core = i0 < P.n_act && (keys[i0] >> 63);
        if ((int)threadIdx.x >= cta_compact(core, q, &s_n)) return;
    }
    const int i = blockIdx.x * blockDim.x + q[threadIdx.x];
    const u64 key = keys[i];
    const PointView p = view(key, P);
    const int lo_s = __ldg(sstart + p.s + 1);
    const int a = __ldg(sstart + p.s);
    if (a < lo_s) {
        const int ha = chain[i];
        u64 base = (u64)(p.s - 1) << P.bu;
        int j = lower_bound_su(keys, a, lo_s, base | p.ulo, P.be);
        u64 top = base | p.uhi;
        for (; j < lo_s; ++j) {
            u64 kq = keys[j];
            if (key_su(kq, P.be) > top) break;
            if (!(kq >> 63)) continue;
            if (((u32)kq & P.emask) >= p.vm) {
                const int hb = chain[j];
                if (seen[ha] != hb) {                                  // racy hint: a stale value only costs a redundant union
                    seen[ha] = hb;
                    uf_union(parent, ha, hb);
                }
                // the rest of this chain adds nothing: continue at the next chain head (dense strips
                // hold hundreds of chain members inside one window)
                j = next_head[j] - 1;
            }
        }
    }
}

// full compression + per-component statistics.  Sorted order is spatially coherent, so lanes of a warp
// mostly share a root: statistics are reduced per (warp, root) group before touching global atomics
// (the giant diagonal component would otherwise serialise millions of same-address atomics).
#define CMP_SLOTS 64
__global__ void __launch_bounds__(256) compress_kernel(const u64* __restrict__ keys, const u32* __restrict__ rows, GridParams P,
                                                       Work W, int variant) {
    // (warp, root) groups reduce in registers, their leaders in a small direct-mapped CTA table; only
    // one partial per (CTA, root) -- or a table collision -- reaches the global atomics
    __shared__ int s_root[CMP_SLOTS], s_cnt[CMP_SLOTS], s_min[CMP_SLOTS];
    if (threadIdx.x < CMP_SLOTS) { s_root[threadIdx.x] = -1; s_cnt[threadIdx.x] = 0; s_min[threadIdx.x] = INT_MAX; }
    __syncthreads();
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    const bool core = (i < P.n_act) && (keys[i] >> 63);
    if (i < P.n_act && !core) W.assigned[i] = -1;
    const unsigned cm = __ballot_sync(0xffffffffu, core);
    if (core) {
        int r = uf_find(W.parent, i);
        W.parent[i] = r;
        W.assigned[i] = r;
        int rk = (variant == CLOOPS_V1) ? (int)rows[i] : W.cellmin[W.chead[i]];
        const unsigned m = __match_any_sync(cm, r);
        const int vmin = __reduce_min_sync(m, rk);
        if ((int)(threadIdx.x & 31) == __ffs(m) - 1) {
            const int slot = r & (CMP_SLOTS - 1);
            const int prev = atomicCAS(&s_root[slot], -1, r);
            if (prev == -1 || prev == r) {
                atomicAdd(&s_cnt[slot], __popc(m));
                atomicMin(&s_min[slot], vmin);
            } else {
                atomicAdd(&W.ncore[r], __popc(m));
                atomicMin(&W.rank[r], vmin);
            }
        }
        if (r == i) atomicAdd(&W.counters[4], 1);
    }
    __syncthreads();
    if (threadIdx.x < CMP_SLOTS && s_root[threadIdx.x] >= 0) {
        atomicAdd(&W.ncore[s_root[threadIdx.x]], s_cnt[threadIdx.x]);
        atomicMin(&W.rank[s_root[threadIdx.x]], s_min[threadIdx.x]);
    }
}

// parent[] of core points may still be one hop short for points compressed before their root was
// hooked?  No: unions finished in the previous kernel, so uf_find() returns final roots.  Kernels below
// nevertheless resolve roots with root_of() which tolerates un-compressed parents.
__device__ __forceinline__ int root_of(const int* parent, int x) {
    int p = parent[x];
    while (p != x) { x = p; p = parent[x]; }
    return x;
}

// Neighbours of one point mostly belong to one component (after compress_kernel their parent IS the
// root), so the (root, rank, status) triple of the previous neighbour is reused when the parent repeats:
// one coalesced-ish load instead of three dependent random ones per core neighbour.
struct RootCache {
    int last_parent = -1, root = -1, rank = INT_MAX;
    unsigned char status = ST_NONE;
    __device__ __forceinline__ void lookup(const Work& W, int j) {
        const int pj = W.parent[j];
        if (pj == last_parent) return;
        last_parent = pj;
        int x = pj, q = W.parent[x];
        while (q != x) { x = q; q = W.parent[x]; }
        root = x;
        rank = W.rank[x];
        status = W.status[x];
    }
};

// ---- v1 border ownership --------------------------------------------------------------------------
__global__ void __launch_bounds__(256) v1_border_kernel(const u64* __restrict__ keys, const int* __restrict__ sstart,
                                                        const u32* __restrict__ rows, GridParams P, Work W) {
    __shared__ int q[256];
    __shared__ int s_n;
    {
        const int i0 = blockIdx.x * blockDim.x + threadIdx.x;
        bool todo = i0 < P.n_act && !(keys[i0] >> 63);
        if (todo) {                                         // no core point in the 3 x 3 cells around it: cannot be a border point
            const u64 k0 = keys[i0];
            todo = cb_any_core_near(W.corebits, (u32)(k0 >> P.sshift), ((u32)(k0 >> P.be) & P.umask) / (u32)P.eps);
        }
        if ((int)threadIdx.x >= cta_compact(todo, q, &s_n)) return;
    }
    const int i = blockIdx.x * blockDim.x + q[threadIdx.x];
    const u64 key = keys[i];
    const PointView p = view(key, P);
    int best_seed_rank = -1, best_seed_root = -1, best_any_rank = INT_MAX, best_any_root = -1;
    RootCache rc;
    for_each_neighbour(keys, sstart, P, i, p, [&](int j, u64 kq) {
        if (!(kq >> 63)) return true;
        rc.lookup(W, j);
        const int r = rc.root, rk = rc.rank;
        if (rk > best_seed_rank && (int)rows[j] == rk) { best_seed_rank = rk; best_seed_root = r; }
        if (rk < best_any_rank) { best_any_rank = rk; best_any_root = r; }
        return true;
    });
    W.assigned[i] = best_seed_root >= 0 ? best_seed_root : best_any_root;
}

// members per root (v1's "< minPts members" deletion).  Index order is spatially coherent: lanes are
// grouped by root first, so the giant diagonal cluster costs one atomic per warp, not one per point.
__global__ void __launch_bounds__(256) size_kernel(GridParams P, Work W) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int a = (i < P.n_act) ? W.assigned[i] : -1;
    const unsigned am = __ballot_sync(0xffffffffu, a >= 0);
    if (a < 0) return;
    const unsigned m = __match_any_sync(am, a);
    if ((int)(threadIdx.x & 31) == __ffs(m) - 1) atomicAdd(&W.size[a], __popc(m));
}

// ---- v2 survival ------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) v2_status_kernel(const u64* __restrict__ keys, GridParams P, int minPts, Work W) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P.n_act) return;
    if (!(keys[i] >> 63) || W.parent[i] != i) return;
    if (W.ncore[i] >= minPts) {
        W.status[i] = ST_ALIVE;
    } else {
        W.status[i] = ST_UNDECIDED;
        W.list_und[atomicAdd(&W.counters[0], 1)] = i;
    }
}

__global__ void __launch_bounds__(256) v2_reset_kernel(Work W, int n_und) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_und) return;
    int r = W.list_und[t];
    if (W.status[r] == ST_UNDECIDED) { W.size[r] = W.ncore[r]; W.ub[r] = W.ncore[r]; }
    if (t == 0) W.counters[2] = 0;
}

// lb(K): border points for which K is the lowest-ranked non-dead adjacent component (K's for sure if K lives)
// ub(K): border points adjacent to K with no decided-alive adjacent component of lower rank
// A border point is adjacent to at most NINE components: its neighbours lie in the 3 x 3 block of rotated floor cells
// around it, and the core points of one cell are mutual neighbours (one component per cell, cDBSCAN2.py:80-83) -- so the
// distinct adjacent components fit a fixed register list and every neighbour is looked at once.
#define ADJ_MAX 9
// The distinct components adjacent to every contested border point, gathered ONCE (one neighbour walk per point): the survival
// rounds and the final re-assignment then work on these short lists (adj[k * n_con + t], adj_n[t]) and only re-read the
// components' status -- round 1 walked the neighbourhood again in every round.
__global__ void __launch_bounds__(128) v2_adjacency_kernel(const u64* __restrict__ keys, const int* __restrict__ sstart, GridParams P,
                                                           Work W, int n_con, int* __restrict__ adj, unsigned char* __restrict__ adj_n) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_con) return;
    const int i = W.list_con[t];
    const PointView p = view(keys[i], P);
    int roots[ADJ_MAX];
    int n_adj = 0;
    RootCache rc;
    for_each_neighbour(keys, sstart, P, i, p, [&](int j, u64 kq) {
        if (!(kq >> 63)) return true;
        const int before = rc.last_parent;
        rc.lookup(W, j);
        if (rc.last_parent == before) return true;          // same component as the previous core neighbour
        bool dup = false;
#pragma unroll
        for (int k = 0; k < ADJ_MAX; ++k) dup |= (k < n_adj && roots[k] == rc.root);
        if (dup || n_adj >= ADJ_MAX) return true;
#pragma unroll
        for (int k = 0; k < ADJ_MAX; ++k)
            if (k == n_adj) roots[k] = rc.root;
        ++n_adj;
        return true;
    });
#pragma unroll
    for (int k = 0; k < ADJ_MAX; ++k)
        if (k < n_adj) adj[(size_t)k * n_con + t] = roots[k];
    adj_n[t] = (unsigned char)n_adj;
}

__global__ void __launch_bounds__(128) v2_accumulate_kernel(Work W, int n_con, const int* __restrict__ adj, const unsigned char* __restrict__ adj_n) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_con) return;
    const int n_adj = adj_n[t];
    int roots[ADJ_MAX], ranks[ADJ_MAX];
    unsigned undecided = 0, live = 0;           // bit k: roots[k] is undecided / not dead
#pragma unroll
    for (int k = 0; k < ADJ_MAX; ++k) {
        if (k >= n_adj) continue;
        const int r = adj[(size_t)k * n_con + t];
        const unsigned char st = W.status[r];
        roots[k] = r;
        ranks[k] = W.rank[r];
        if (st != ST_DEAD) live |= 1u << k;
        if (st == ST_UNDECIDED) undecided |= 1u << k;
    }
    int min_nd_rank = INT_MAX, min_nd = -1, min_alive_rank = INT_MAX;
#pragma unroll
    for (int k = 0; k < ADJ_MAX; ++k) {
        if (k >= n_adj || !((live >> k) & 1u)) continue;
        if (ranks[k] < min_nd_rank) { min_nd_rank = ranks[k]; min_nd = k; }
        if (!((undecided >> k) & 1u) && ranks[k] < min_alive_rank) min_alive_rank = ranks[k];
    }
    if (min_nd < 0) return;
#pragma unroll
    for (int k = 0; k < ADJ_MAX; ++k) {
        if (k >= n_adj || !((undecided >> k) & 1u)) continue;
        if (k == min_nd) atomicAdd(&W.size[roots[k]], 1);
        if (ranks[k] < min_alive_rank) atomicAdd(&W.ub[roots[k]], 1);      // every distinct undecided component ranked below the best alive one
    }
}

// final owner of the contested points once every component is decided: the lowest-ranked adjacent component that lives
__global__ void __launch_bounds__(256) v2_reassign_kernel(Work W, int n_con, const int* __restrict__ adj, const unsigned char* __restrict__ adj_n) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_con) return;
    const int n_adj = adj_n[t];
    int best_rank = INT_MAX, best_root = -1;
#pragma unroll
    for (int k = 0; k < ADJ_MAX; ++k) {
        if (k >= n_adj) continue;
        const int r = adj[(size_t)k * n_con + t];
        if (W.status[r] == ST_DEAD) continue;
        const int rk = W.rank[r];
        if (rk < best_rank) { best_rank = rk; best_root = r; }
    }
    W.assigned[W.list_con[t]] = best_root;
}

__global__ void __launch_bounds__(256) v2_decide_kernel(Work W, int n_und, int minPts) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_und) return;
    int r = W.list_und[t];
    if (W.status[r] != ST_UNDECIDED) return;
    if (W.size[r] >= minPts) W.status[r] = ST_ALIVE;
    else if (W.ub[r] < minPts) { W.status[r] = ST_DEAD; atomicAdd(&W.counters[3], 1); }
    else atomicAdd(&W.counters[2], 1);
}

// v2 ownership: lowest-ranked non-dead adjacent component (cDBSCAN2.py:130,162,352).  The first pass
// (FIX = false, every non-core point, nothing dead yet) also queues the points adjacent to an undecided
// component; after the survival rounds only those are re-evaluated (FIX = true).
template <bool FIX>
__global__ void __launch_bounds__(256) v2_border_kernel(const u64* __restrict__ keys, const int* __restrict__ sstart,
                                                        GridParams P, Work W, int n_items) {
    __shared__ int q[256];
    __shared__ int s_n;
    int i;
    if (FIX) {
        const int t = blockIdx.x * blockDim.x + threadIdx.x;
        if (t >= n_items) return;
        i = W.list_con[t];
    } else {                                             // every non-core point, compacted into dense warps
        const int i0 = blockIdx.x * blockDim.x + threadIdx.x;
        bool todo = i0 < n_items && !(keys[i0] >> 63);
        if (todo) {                                         // no core point in the 3 x 3 cells around it: cannot be a border point
            const u64 k0 = keys[i0];
            todo = cb_any_core_near(W.corebits, (u32)(k0 >> P.sshift), ((u32)(k0 >> P.be) & P.umask) / (u32)P.eps);
        }
        if ((int)threadIdx.x >= cta_compact(todo, q, &s_n)) return;
        i = blockIdx.x * blockDim.x + q[threadIdx.x];
    }
    const u64 key = keys[i];
    const PointView p = view(key, P);
    int best_rank = INT_MAX, best_root = -1;
    bool contested = false;
    RootCache rc;
    for_each_neighbour(keys, sstart, P, i, p, [&](int j, u64 kq) {
        if (!(kq >> 63)) return true;
        rc.lookup(W, j);
        if (rc.status == ST_DEAD) return true;
        if (rc.status == ST_UNDECIDED) contested = true;
        if (rc.rank < best_rank) { best_rank = rc.rank; best_root = rc.root; }
        return true;
    });
    W.assigned[i] = best_root;
    if (!FIX && contested) W.list_con[atomicAdd(&W.counters[1], 1)] = i;
}

// The first ownership pass with G lanes per border point.  One lane per point walks its three strip ranges alone: the
// lanes of a warp then read 32 unrelated places (ncu, chr1 of config 4: 9.5 sectors per load request, 14 of 32 lanes
// active per instruction, 191 us).  Here the CTA's candidates are compacted and dealt to groups of G lanes; a group walks
// one point's ranges with stride G and combines (lowest rank, contested) by shuffles inside the group.
template <int G>
__global__ void __launch_bounds__(256) v2_border_group_kernel(const u64* __restrict__ keys, const int* __restrict__ sstart,
                                                              GridParams P, Work W, int n_items) {
    __shared__ int q[256];
    __shared__ int s_n;
    const int i0 = blockIdx.x * blockDim.x + threadIdx.x;
    bool todo = i0 < n_items && !(keys[i0] >> 63);
    if (todo) {
        const u64 k0 = keys[i0];
        todo = cb_any_core_near(W.corebits, (u32)(k0 >> P.sshift), ((u32)(k0 >> P.be) & P.umask) / (u32)P.eps);
    }
    const int n_todo = cta_compact(todo, q, &s_n);
    const int lg = threadIdx.x % G;
    const unsigned gmask = (G == 32 ? 0xffffffffu : ((1u << G) - 1u)) << ((threadIdx.x & 31) / G * G);
    for (int t = threadIdx.x / G; t < n_todo; t += 256 / G) {
        const int i = blockIdx.x * blockDim.x + q[t];
        const PointView p = view(keys[i], P);
        int best_rank = INT_MAX, best_root = -1;
        int contested = 0;
        RootCache rc;
        for_each_neighbour_strided<G>(keys, sstart, P, i, p, lg, [&](int j, u64 kq) {
            if (!(kq >> 63)) return;
            rc.lookup(W, j);
            if (rc.status == ST_DEAD) return;
            if (rc.status == ST_UNDECIDED) contested = 1;
            if (rc.rank < best_rank) { best_rank = rc.rank; best_root = rc.root; }
        });
#pragma unroll
        for (int d = G / 2; d > 0; d >>= 1) {
            const int o_rank = __shfl_xor_sync(gmask, best_rank, d), o_root = __shfl_xor_sync(gmask, best_root, d);
            contested |= __shfl_xor_sync(gmask, contested, d);
            if (o_rank < best_rank) { best_rank = o_rank; best_root = o_root; }
        }
        if (lg == 0) {
            W.assigned[i] = best_root;
            if (contested) W.list_con[atomicAdd(&W.counters[1], 1)] = i;
        }
    }
}

// ---- numbering ----------------------------------------------------------------------------------------
// flags[rank of root] = 1 for every component that receives an id
__global__ void __launch_bounds__(256) number_flags_kernel(const u64* __restrict__ keys, GridParams P, Work W, int variant) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P.n_act) return;
    if (!(keys[i] >> 63) || W.parent[i] != i) return;
    bool numbered = (variant == CLOOPS_V1) ? true : (W.status[i] != ST_DEAD);
    if (numbered) W.flags[W.rank[i]] = 1;
}

// labels == NULL skips the scatter to row order (callers that work in index order, e.g. the pipeline)
__global__ void __launch_bounds__(256) label_kernel(const u32* __restrict__ rows, GridParams P, Work W, int variant, int minPts,
                                                    int* __restrict__ labels, int* __restrict__ labels_sorted) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    int lab = -1;
    if (i < P.n_act) {
        int a = W.assigned[i];
        if (a >= 0) {
            bool keep = (variant == CLOOPS_V1) ? (W.size[a] >= minPts) : (W.status[a] != ST_DEAD);
            if (keep) lab = W.ids[W.rank[a]];
        }
        if (labels) labels[rows[i]] = lab;
        if (labels_sorted) labels_sorted[i] = lab;
    }
    block_count(W.slots_lab, lab >= 0);
}

int index_dbscan(cloops_index* ix, int minPts, int variant, int* d_labels, int* d_labels_sorted, int64_t* h_info, cudaStream_t st) {
    const GridParams& P = ix->P;
    if (minPts < 1) return fail(CLOOPS_EINVAL, "minPts must be >= 1 (got %d)", minPts);
    if (variant != CLOOPS_V1 && variant != CLOOPS_V2) return fail(CLOOPS_EINVAL, "variant %d not served by the strip index", variant);
    if (h_info) for (int k = 0; k < 8; ++k) h_info[k] = 0;
    if (P.n == 0) return 0;
    if (d_labels && P.n_act < P.n) LAUNCH(fill_int_kernel, cdiv(P.n, 256), 256, 0, st, d_labels, -1, (long long)P.n);
    if (P.n_act == 0) return 0;
    const int na = P.n_act, g = cdiv(na, 256);
    Temp tmp(st);
    Work W;
    RET_IF(tmp.alloc(&W.cnt, na));
    RET_IF(tmp.alloc(&W.parent, na));
    RET_IF(tmp.alloc(&W.rank, na));
    RET_IF(tmp.alloc(&W.ncore, na));
    RET_IF(tmp.alloc(&W.assigned, na));
    RET_IF(tmp.alloc(&W.size, na));
    RET_IF(tmp.alloc(&W.ub, na));
    RET_IF(tmp.alloc(&W.status, na));
    RET_IF(tmp.alloc(&W.flags, (size_t)P.n + 1));
    RET_IF(tmp.alloc(&W.ids, (size_t)P.n + 1));
    RET_IF(tmp.alloc(&W.counters, 8 + 2 * CTR_SLOTS * CTR_STRIDE));
    int* d_scan_tmp;
    RET_IF(tmp.alloc(&d_scan_tmp, scan_tmp_ints((long long)P.n + 1)));
    W.chead = W.cellmin = W.list_und = W.list_con = nullptr;
    const bool v2 = variant == CLOOPS_V2;
    if (v2) {
        RET_IF(tmp.alloc(&W.chead, na));
        RET_IF(tmp.alloc(&W.cellmin, na));
        RET_IF(tmp.alloc(&W.list_und, na));
        RET_IF(tmp.alloc(&W.list_con, na));
    }
    CU_TRY(cudaMemsetAsync(W.counters, 0, (8 + 2 * CTR_SLOTS * CTR_STRIDE) * sizeof(int), st));
    W.slots_core = W.counters + 8;
    W.slots_lab = W.counters + 8 + CTR_SLOTS * CTR_STRIDE;
    CU_TRY(cudaMemsetAsync(W.flags, 0, ((size_t)P.n + 1) * sizeof(int), st));
    RET_IF(tmp.alloc(&W.corebits, CB_WORDS));
    CU_TRY(cudaMemsetAsync(W.corebits, 0, (size_t)CB_WORDS * sizeof(u32), st));

    stage_mark("workspace", st);                 // allocations + the two memsets above, so that "region_query" times the kernel alone
    RET_IF(index_count(ix, minPts, W.cnt, st));
    stage_mark("region_query", st);
    LAUNCH(flag_kernel, g, 256, 0, st, ix->keys, P, minPts, W, v2 ? 1 : 0);
    if (v2) {
        RET_IF((device_scan<SCAN_MAX, true, false>(W.chead, W.chead, na, d_scan_tmp, st)));
        LAUNCH(cellmin_kernel, g, 256, 0, st, ix->rows, P, W);
    }
    stage_mark("flags_cells", st);
    {
        // chains inside strips by scan (parent = latest chain head), cross-strip links by union-find
        int* head = W.assigned;                    // scratch: assigned[] is written later by compress_kernel
        int* chain = W.size;                       // scratch until size_kernel / v2 survival re-initialise it
        int* seen = W.ub;
        int* nh_in = W.ncore;                      // scratch: ncore[] is zeroed again below
        int* next_head;                            // next_head[i] = first chain head with index > i (or n_act)
        RET_IF(tmp.alloc(&next_head, (size_t)na + 1));
        LAUNCH(chain_head_kernel, g, 256, 0, st, ix->keys, ix->sstart, P, head, nh_in);
        RET_IF((device_scan<SCAN_MAX, true, false>(head, chain, na, d_scan_tmp, st)));
        CU_TRY(cudaMemcpyAsync(W.parent, chain, (size_t)na * sizeof(int), cudaMemcpyDeviceToDevice, st));
        {
            // suffix-min over (head ? index : INT_MAX), shifted by one: scan the reversed range
            LAUNCH(fill_int_kernel, 1, 32, 0, st, next_head + na, na, 1LL);
            RET_IF((device_scan<SCAN_MIN, true, true>(nh_in, next_head, na, d_scan_tmp, st)));      // suffix minimum
            LAUNCH(clamp_next_head_kernel, g, 256, 0, st, next_head, na);
            CU_TRY(cudaMemsetAsync(W.ncore, 0, (size_t)na * sizeof(int), st));
        }
        CU_TRY(cudaMemsetAsync(seen, 0xff, (size_t)na * sizeof(int), st));
        stage_mark("chains", st);
        LAUNCH(union_kernel, g, 256, 0, st, ix->keys, ix->sstart, P, chain, next_head + 1, seen, W.parent);
        CU_TRY(cudaMemsetAsync(W.size, 0, (size_t)na * sizeof(int), st));
    }
    stage_mark("union", st);
    LAUNCH(compress_kernel, g, 256, 0, st, ix->keys, ix->rows, P, W, variant);
    stage_mark("compress", st);

    int counters[8] = {0};
    if (!v2) {
        LAUNCH(v1_border_kernel, g, 256, 0, st, ix->keys, ix->sstart, ix->rows, P, W);
        LAUNCH(size_kernel, g, 256, 0, st, P, W);
        stage_mark("border", st);
    } else {
        LAUNCH(v2_status_kernel, g, 256, 0, st, ix->keys, P, minPts, W);
        static const int border_group = getenv("CLOOPS_BORDER_GROUP") ? atoi(getenv("CLOOPS_BORDER_GROUP")) : 4;      // 1: one lane per point
        if (border_group == 4) LAUNCH(v2_border_group_kernel<4>, g, 256, 0, st, ix->keys, ix->sstart, P, W, na);
        else if (border_group == 8) LAUNCH(v2_border_group_kernel<8>, g, 256, 0, st, ix->keys, ix->sstart, P, W, na);
        else if (border_group == 2) LAUNCH(v2_border_group_kernel<2>, g, 256, 0, st, ix->keys, ix->sstart, P, W, na);
        else LAUNCH(v2_border_kernel<false>, g, 256, 0, st, ix->keys, ix->sstart, P, W, na);
        stage_mark("border", st);
        CU_TRY(cudaMemcpyAsync(counters, W.counters, sizeof(counters), cudaMemcpyDeviceToHost, st));
        CU_TRY(cudaStreamSynchronize(st));
        const int n_und = counters[0], n_con = counters[1];
        if (n_und > 0) {
            int* adj = nullptr;
            unsigned char* adj_n = nullptr;
            if (n_con > 0) {
                RET_IF(tmp.alloc(&adj, (size_t)ADJ_MAX * n_con));
                RET_IF(tmp.alloc(&adj_n, (size_t)n_con));
                LAUNCH(v2_adjacency_kernel, cdiv(n_con, 128), 128, 0, st, ix->keys, ix->sstart, P, W, n_con, adj, adj_n);
            }
            for (int round = 0;; ++round) {
                LAUNCH(v2_reset_kernel, cdiv(n_und, 256), 256, 0, st, W, n_und);
                if (n_con > 0) LAUNCH(v2_accumulate_kernel, cdiv(n_con, 128), 128, 0, st, W, n_con, adj, adj_n);
                LAUNCH(v2_decide_kernel, cdiv(n_und, 256), 256, 0, st, W, n_und, minPts);
                CU_TRY(cudaMemcpyAsync(counters, W.counters, sizeof(counters), cudaMemcpyDeviceToHost, st));
                CU_TRY(cudaStreamSynchronize(st));
                g_trace_rounds = round + 1;
                if (counters[2] == 0) break;
                if (round > na) return fail(CLOOPS_ECUDA, "v2 survival did not converge");
            }
            if (n_con > 0 && counters[3] > 0)
                LAUNCH(v2_reassign_kernel, cdiv(n_con, 256), 256, 0, st, W, n_con, adj, adj_n);
            if (g_trace) fprintf(stderr, "[cloops] v2 survival: n_act=%d undecided=%d contested=%d rounds=%d dead=%d\n", na, n_und, n_con, g_trace_rounds, counters[3]);
        }
        stage_mark("survival", st);
    }
    LAUNCH(number_flags_kernel, g, 256, 0, st, ix->keys, P, W, variant);
    RET_IF((device_scan<SCAN_ADD, false, false>(W.flags, W.ids, (long long)P.n + 1, d_scan_tmp, st)));
    LAUNCH(label_kernel, g, 256, 0, st, ix->rows, P, W, variant, minPts, d_labels, d_labels_sorted);
    stage_mark("labels", st);
    if (h_info) {
        int n_clusters = 0;
        static thread_local int slots[2 * CTR_SLOTS * CTR_STRIDE];
        CU_TRY(cudaMemcpyAsync(counters, W.counters, sizeof(counters), cudaMemcpyDeviceToHost, st));
        CU_TRY(cudaMemcpyAsync(slots, W.slots_core, sizeof(slots), cudaMemcpyDeviceToHost, st));
        CU_TRY(cudaMemcpyAsync(&n_clusters, W.ids + P.n, sizeof(int), cudaMemcpyDeviceToHost, st));
        CU_TRY(cudaStreamSynchronize(st));
        long long n_core = 0, n_lab = 0;
        for (int k = 0; k < CTR_SLOTS; ++k) {
            n_core += slots[k * CTR_STRIDE];
            n_lab += slots[(CTR_SLOTS + k) * CTR_STRIDE];
        }
        h_info[0] = P.n_act;
        h_info[1] = n_clusters;
        h_info[2] = counters[4];
        h_info[3] = n_core;
        h_info[4] = counters[3];
        h_info[5] = P.ns;
        h_info[6] = P.be + P.bu + P.bs;
        h_info[7] = n_lab;
    }
    return 0;
}

}  // namespace cloops
