// The resident per-chromosome, per-eps index: PETs sorted by (v-strip, u) as packed 64-bit keys.
#pragma once
#include "common.cuh"

struct cloops_index {
    cloops::GridParams P;
    u64* keys = nullptr;      // [n] sorted packed keys (active rows first, inactive rows in sentinel strip)
    u32* rows = nullptr;      // [n] original row of each sorted position
    int* sstart = nullptr;    // [ns+3] dense strip offsets, entry k = first sorted index of strip k-1
    void* tiles = nullptr;    // [ceil(n_act/1024)] per-tile headers of the region query (index.cu:TileInfo)
    int rmax = 0;             // staging capacity the tile headers were computed for
    int counted_cap = 0;      // cap of the counts currently flagged into the keys (0 = none)
};

namespace cloops {

// device helpers shared by every kernel that walks the index --------------------------------------
__device__ __forceinline__ u64 key_su(u64 key, int be) { return (key & KEY_MASK) >> be; }

// first j in [lo,hi) whose (strip,u') >= target
__device__ __forceinline__ int lower_bound_su(const u64* __restrict__ keys, int lo, int hi, u64 target, int be) {
    while (lo < hi) {
        int mid = (lo + hi) >> 1;
        if (key_su(__ldg(keys + mid), be) < target) lo = mid + 1; else hi = mid;
    }
    return lo;
}

struct PointView {
    int s;        // strip
    u32 up, vm;   // u', v mod eps
    u32 ulo;      // max(up-eps,0)
    u64 uhi;      // min(up+eps, umask)
    bool core;
};

__device__ __forceinline__ PointView view(u64 key, const GridParams& P) {
    PointView p;
    p.core = (key >> 63) != 0;
    key &= KEY_MASK;
    p.s = (int)(key >> P.sshift);
    p.up = (u32)(key >> P.be) & P.umask;
    p.vm = (u32)key & P.emask;
    p.ulo = p.up > (u32)P.eps ? p.up - (u32)P.eps : 0u;
    u64 h = (u64)p.up + (u64)P.eps;
    p.uhi = h < (u64)P.umask ? h : (u64)P.umask;
    return p;
}

// Visit every j != i within Manhattan eps of sorted point i (all three strips).  f(j, key_j) returns
// false to stop early.  Own strip needs no v test; strip s-1 needs vm_q >= vm ; strip s+1 needs vm_q <= vm.
template <class F>
__device__ __forceinline__ void for_each_neighbour(const u64* __restrict__ keys, const int* __restrict__ sstart,
                                                   const GridParams& P, int i, const PointView& p, F&& f) {
    const int lo_s = sstart[p.s + 1], hi_s = sstart[p.s + 2];
    for (int j = i - 1; j >= lo_s; --j) {
        u64 kq = keys[j];
        if (((u32)(kq >> P.be) & P.umask) < p.ulo) break;
        if (!f(j, kq)) return;
    }
    for (int j = i + 1; j < hi_s; ++j) {
        u64 kq = keys[j];
        if ((u64)((u32)(kq >> P.be) & P.umask) > p.uhi) break;
        if (!f(j, kq)) return;
    }
    {
        const int a = sstart[p.s];
        if (a < lo_s) {
            u64 base = (u64)(p.s - 1) << P.bu;
            int j = lower_bound_su(keys, a, lo_s, base | p.ulo, P.be);
            u64 top = base | p.uhi;
            for (; j < lo_s; ++j) {
                u64 kq = keys[j];
                if (key_su(kq, P.be) > top) break;
                if (((u32)kq & P.emask) >= p.vm)
                    if (!f(j, kq)) return;
            }
        }
    }
    {
        const int b = sstart[p.s + 3];
        if (hi_s < b) {
            u64 base = (u64)(p.s + 1) << P.bu;
            int j = lower_bound_su(keys, hi_s, b, base | p.ulo, P.be);
            u64 top = base | p.uhi;
            for (; j < b; ++j) {
                u64 kq = keys[j];
                if (key_su(kq, P.be) > top) break;
                if (((u32)kq & P.emask) <= p.vm)
                    if (!f(j, kq)) return;
            }
        }
    }
}

// The same visit shared by G lanes: lane `lg` of the group takes every G-th candidate of each of the four ranges (the
// candidates of a range are consecutive keys, so the G loads of a step fall into one or two sectors; the range ends are
// monotone in the key, so every lane may stop on its own).  f(j, key_j); no early exit.
template <int G, class F>
__device__ __forceinline__ void for_each_neighbour_strided(const u64* __restrict__ keys, const int* __restrict__ sstart,
                                                           const GridParams& P, int i, const PointView& p, int lg, F&& f) {
    const int lo_s = sstart[p.s + 1], hi_s = sstart[p.s + 2];
    for (int j = i - 1 - lg; j >= lo_s; j -= G) {
        u64 kq = keys[j];
        if (((u32)(kq >> P.be) & P.umask) < p.ulo) break;
        f(j, kq);
    }
    for (int j = i + 1 + lg; j < hi_s; j += G) {
        u64 kq = keys[j];
        if ((u64)((u32)(kq >> P.be) & P.umask) > p.uhi) break;
        f(j, kq);
    }
    {
        const int a = sstart[p.s];
        if (a < lo_s) {
            u64 base = (u64)(p.s - 1) << P.bu;
            int j = lower_bound_su(keys, a, lo_s, base | p.ulo, P.be) + lg;
            u64 top = base | p.uhi;
            for (; j < lo_s; j += G) {
                u64 kq = keys[j];
                if (key_su(kq, P.be) > top) break;
                if (((u32)kq & P.emask) >= p.vm) f(j, kq);
            }
        }
    }
    {
        const int b = sstart[p.s + 3];
        if (hi_s < b) {
            u64 base = (u64)(p.s + 1) << P.bu;
            int j = lower_bound_su(keys, hi_s, b, base | p.ulo, P.be) + lg;
            u64 top = base | p.uhi;
            for (; j < b; j += G) {
                u64 kq = keys[j];
                if (key_su(kq, P.be) > top) break;
                if (((u32)kq & P.emask) <= p.vm) f(j, kq);
            }
        }
    }
}

int index_build(const int32_t* d_x, const int32_t* d_y, int64_t n, int32_t eps, int32_t cut, cloops_index** out,
                cudaStream_t st);
int index_filter(const cloops_index* base, int32_t cut, cloops_index** out, cudaStream_t st);
void index_free(cloops_index* ix, cudaStream_t st);
int index_count(cloops_index* ix, int cap, int* d_counts_sorted, cudaStream_t st);   // region_query.cu
int index_tiles(cloops_index* ix, cudaStream_t st);                                    // region_query.cu
int index_coords(cloops_index* ix, int* d_xs, int* d_ys, cudaStream_t st);

}  // namespace cloops
