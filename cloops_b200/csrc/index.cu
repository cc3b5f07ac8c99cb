// Index build (rotate -> pack -> radix sort -> strip table) and the region-query (neighbour count) kernel.
//
// Geometry (cDBSCAN2.py:66-70): (u,v) = (X-Y, X+Y) turns Manhattan d1 <= eps into
// max(|du|,|dv|) <= eps.  Points are sorted by (strip = floor(v/eps), u); the eps-neighbourhood of a
// point is then three contiguous runs (strips s-1, s, s+1 restricted to u in [u-eps, u+eps]); inside
// the own strip |dv| <= eps-1 holds by construction, in strip s-1 (s+1) the remaining test is
// vmod_q >= vmod_p (vmod_q <= vmod_p).
#include <limits.h>
#include <string.h>

#include <algorithm>
#include <cub/cub.cuh>

#include "index.cuh"

namespace cloops {

struct Extents {
    int umin, umax, vmin, vmax, n_act, overflow;
};

__global__ void extents_init_kernel(Extents* e) {
    e->umin = INT_MAX; e->umax = INT_MIN; e->vmin = INT_MAX; e->vmax = INT_MIN; e->n_act = 0; e->overflow = 0;
}

__global__ void __launch_bounds__(256) extents_kernel(const int* __restrict__ x, const int* __restrict__ y, int n, int cut,
                                                      Extents* out) {
    int umin = INT_MAX, umax = INT_MIN, vmin = INT_MAX, vmax = INT_MIN, cnt = 0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        int xx = __ldg(x + i), yy = __ldg(y + i);
        if (xx < -(1 << 30) || xx >= (1 << 30) || yy < -(1 << 30) || yy >= (1 << 30)) {
            out->overflow = 1;                            // rotated coordinates would leave int32
            continue;
        }
        if (cut > 0 && yy - xx < cut) continue;
        int u = xx - yy, v = xx + yy;
        umin = min(umin, u); umax = max(umax, u); vmin = min(vmin, v); vmax = max(vmax, v);
        ++cnt;
    }
    umin = __reduce_min_sync(0xffffffffu, umin);
    umax = __reduce_max_sync(0xffffffffu, umax);
    vmin = __reduce_min_sync(0xffffffffu, vmin);
    vmax = __reduce_max_sync(0xffffffffu, vmax);
    cnt = __reduce_add_sync(0xffffffffu, cnt);
    if ((threadIdx.x & 31) == 0 && cnt > 0) {
        atomicMin(&out->umin, umin); atomicMax(&out->umax, umax);
        atomicMin(&out->vmin, vmin); atomicMax(&out->vmax, vmax);
        atomicAdd(&out->n_act, cnt);
    }
}

__global__ void __launch_bounds__(256) pack_kernel(const int* __restrict__ x, const int* __restrict__ y, int cut, GridParams P,
                                                   u64* __restrict__ keys, u32* __restrict__ rows) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P.n) return;
    int xx = __ldg(x + i), yy = __ldg(y + i);
    u64 key;
    if (cut > 0 && yy - xx < cut) {
        key = (u64)P.ns << P.sshift;                      // sentinel strip: sorts behind every active row
    } else {
        u32 up = (u32)((xx - yy) - P.ubase);
        u32 vp = (u32)((xx + yy) - P.vbase);
        u32 sv = vp / (u32)P.eps;
        u32 vm = vp - sv * (u32)P.eps;
        key = ((u64)sv << P.sshift) | ((u64)up << P.be) | (u64)vm;
    }
    keys[i] = key;
    rows[i] = (u32)i;
}

// sstart[k] = first sorted index whose strip >= k-1, k in [0, ns+2]
__global__ void __launch_bounds__(256) strip_table_gap_kernel(const u64* __restrict__ keys, GridParams P, int* __restrict__ sstart) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i > P.n_act) return;
    int cur = (i < P.n_act) ? (int)((keys[i] & KEY_MASK) >> P.sshift) : P.ns + 1;
    int prev = (i > 0) ? (int)((keys[i - 1] & KEY_MASK) >> P.sshift) : -2;
    for (int k = prev + 2; k <= cur + 1; ++k) sstart[k] = i;
}

__global__ void __launch_bounds__(256) strip_table_search_kernel(const u64* __restrict__ keys, GridParams P, int* __restrict__ sstart) {
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k > P.ns + 2) return;
    int r;
    if (k == 0) r = 0;
    else if (k - 1 >= P.ns) r = P.n_act;
    else r = lower_bound_su(keys, 0, P.n_act, (u64)(k - 1) << P.bu, P.be);
    sstart[k] = r;
}

// --------------------------------------------------------------------------------------------------
// Region query: neighbour count per point, saturating at cap (cDBSCAN.py:186-205; cDBSCAN2.py:304-346).
//
// A CTA owns TILE consecutive sorted points.  Everything its threads can touch -- the strips of the
// tile plus one strip below and one above -- is ONE contiguous key range R (strips are consecutive in
// the sort order), staged once into shared memory as 32-bit u' and vmod arrays with coalesced loads.
// Phase 1 (every thread): walk left/right inside the own strip; dense points saturate here.
// Phase 2 (compacted): the still unsaturated points are queued in shared memory so that full warps run
// the two binary searches + window scans of strips s-1 and s+1.
// If R does not fit (very long strips: dense Hi-C diagonals, where phase 1 saturates almost at once)
// the CTA falls back to the same walk on global memory through L1.
#define CQ_TILE 256
#define CQ_RMAX 2048

__device__ __forceinline__ int count_point_global(const u64* __restrict__ keys, const int* __restrict__ sstart, const GridParams& P,
                                                  int cap, int i) {
    const PointView p = view(keys[i], P);
    const int lo_s = __ldg(sstart + p.s + 1), hi_s = __ldg(sstart + p.s + 2);
    int c = 1;
    for (int j = i - 1; j >= lo_s && c < cap; --j) {
        if (((u32)(keys[j] >> P.be) & P.umask) < p.ulo) break;
        ++c;
    }
    for (int j = i + 1; j < hi_s && c < cap; ++j) {
        if ((u64)((u32)(keys[j] >> P.be) & P.umask) > p.uhi) break;
        ++c;
    }
    if (c < cap) {
        const int a = __ldg(sstart + p.s);
        if (a < lo_s) {
            u64 base = (u64)(p.s - 1) << P.bu;
            int j = lower_bound_su(keys, a, lo_s, base | p.ulo, P.be);
            u64 top = base | p.uhi;
            for (; j < lo_s && c < cap; ++j) {
                u64 kq = keys[j];
                if (key_su(kq, P.be) > top) break;
                c += (((u32)kq & P.emask) >= p.vm) ? 1 : 0;
            }
        }
    }
    if (c < cap) {
        const int b = __ldg(sstart + p.s + 3);
        if (hi_s < b) {
            u64 base = (u64)(p.s + 1) << P.bu;
            int j = lower_bound_su(keys, hi_s, b, base | p.ulo, P.be);
            u64 top = base | p.uhi;
            for (; j < b && c < cap; ++j) {
                u64 kq = keys[j];
                if (key_su(kq, P.be) > top) break;
                c += (((u32)kq & P.emask) <= p.vm) ? 1 : 0;
            }
        }
    }
    return c;
}

__device__ __forceinline__ int lower_bound_s(const u32* __restrict__ U, int lo, int hi, u32 target) {   // first U >= target
    while (lo < hi) {
        int mid = (lo + hi) >> 1;
        if (U[mid] < target) lo = mid + 1; else hi = mid;
    }
    return lo;
}
__device__ __forceinline__ int upper_bound_s(const u32* __restrict__ U, int lo, int hi, u32 target) {   // first U > target
    while (lo < hi) {
        int mid = (lo + hi) >> 1;
        if (U[mid] <= target) lo = mid + 1; else hi = mid;
    }
    return lo;
}

// window of an adjacent strip [a,b): 4 predicated probes from the lower bound (no divergence), then a
// tail loop for the rare longer windows.  UPPER: strip s-1 needs vmod_q >= vmod_p, strip s+1 vmod_q <= vmod_p.
template <bool PREV>
__device__ __forceinline__ int adjacent_strip_count(const u32* __restrict__ U, const u32* __restrict__ V, int a, int b, int me,
                                                    u32 ulo, u32 uhi, u32 vm, int c, int cap) {
    int j = lower_bound_s(U, a, b, ulo);
    bool in = true;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int jj = j + k;
        in = in && jj < b;
        const int js = in ? jj : me;
        in = in && U[js] <= uhi;
        const u32 vq = V[js];
        c += (in && (PREV ? vq >= vm : vq <= vm)) ? 1 : 0;
    }
    if (in) {
        for (int jj = j + 4; jj < b && c < cap; ++jj) {
            if (U[jj] > uhi) break;
            const u32 vq = V[jj];
            c += (PREV ? vq >= vm : vq <= vm) ? 1 : 0;
        }
    }
    return c;
}

__global__ void __launch_bounds__(CQ_TILE) count_kernel(const u64* __restrict__ keys, const int* __restrict__ sstart, GridParams P,
                                                        int cap, int* __restrict__ cnt) {
    __shared__ u32 U[CQ_RMAX];
    __shared__ u32 V[CQ_RMAX];
    __shared__ int q_pt[CQ_TILE];       // queued points (index relative to R)
    __shared__ int q_c[CQ_TILE];        // their partial counts
    __shared__ int q_s[CQ_TILE];        // their strips
    __shared__ int s_sA, s_sB, s_nq;
    const int tid = threadIdx.x;
    const int t0 = blockIdx.x * CQ_TILE;
    const int t1 = min(t0 + CQ_TILE, P.n_act);
    const int i = t0 + tid;
    int s = 0;
    if (i < t1) {
        s = (int)((__ldg(keys + i) & KEY_MASK) >> P.sshift);
        if (tid == 0) { s_sA = s; s_nq = 0; }
        if (i == t1 - 1) s_sB = s;
    }
    __syncthreads();
    const int r0 = __ldg(sstart + s_sA);          // first point of strip sA-1
    const int r1 = __ldg(sstart + s_sB + 3);      // end of strip sB+1
    if (r1 - r0 > CQ_RMAX) {                      // CTA-uniform
        if (i < t1) cnt[i] = count_point_global(keys, sstart, P, cap, i);
        return;
    }
    for (int j = r0 + tid; j < r1; j += CQ_TILE) {
        const u64 k = __ldg(keys + j);
        U[j - r0] = (u32)(k >> P.be) & P.umask;
        V[j - r0] = (u32)k & P.emask;
    }
    __syncthreads();
    // ---- phase 1: own strip.  Uniform trip counts: the in-window predicate is monotone along the sorted
    // strip, so probing the cap-1 nearest points on each side gives min(count, cap-1) per side.
    int c = 1;
    bool need = false;
    const int me = i - r0;
    if (i < t1) {
        const int lo_s = __ldg(sstart + s + 1) - r0, hi_s = __ldg(sstart + s + 2) - r0;
        const u32 up = U[me];
        const u32 ulo = up > (u32)P.eps ? up - (u32)P.eps : 0u;
        const u64 h = (u64)up + (u64)P.eps;
        const u32 uhi = h < (u64)P.umask ? (u32)h : P.umask;
        if (cap <= 9) {
            for (int k = 1; k < cap; ++k) {
                const int jl = me - k, jr = me + k;
                const bool okl = jl >= lo_s, okr = jr < hi_s;
                const u32 ul = U[okl ? jl : me], ur = U[okr ? jr : me];
                c += (okl && ul >= ulo) ? 1 : 0;
                c += (okr && ur <= uhi) ? 1 : 0;
            }
        } else {
            c = upper_bound_s(U, me + 1, hi_s, uhi) - lower_bound_s(U, lo_s, me, ulo);
        }
        need = c < cap;
        if (!need) cnt[i] = cap;
    }
    // ---- compaction of the unsaturated points
    {
        const unsigned b = __ballot_sync(0xffffffffu, need);
        const int lane = tid & 31;
        int base = 0;
        if (lane == 0 && b) base = atomicAdd(&s_nq, __popc(b));
        base = __shfl_sync(0xffffffffu, base, 0);
        if (need) {
            const int slot = base + __popc(b & ((1u << lane) - 1));
            q_pt[slot] = me;
            q_c[slot] = c;
            q_s[slot] = s;
        }
    }
    __syncthreads();
    // ---- phase 2: strips s-1 and s+1
    if (tid < s_nq) {
        const int pm = q_pt[tid];
        const int ps = q_s[tid];
        c = q_c[tid];
        const u32 up = U[pm], vm = V[pm];
        const u32 ulo = up > (u32)P.eps ? up - (u32)P.eps : 0u;
        const u64 h = (u64)up + (u64)P.eps;
        const u32 uhi = h < (u64)P.umask ? (u32)h : P.umask;
        const int a = __ldg(sstart + ps) - r0, lo_s = __ldg(sstart + ps + 1) - r0;
        const int hi_s = __ldg(sstart + ps + 2) - r0, b = __ldg(sstart + ps + 3) - r0;
        c = adjacent_strip_count<true>(U, V, a, lo_s, pm, ulo, uhi, vm, c, cap);
        if (c < cap) c = adjacent_strip_count<false>(U, V, hi_s, b, pm, ulo, uhi, vm, c, cap);
        cnt[r0 + pm] = c < cap ? c : cap;
    }
}

// (X, Y) of every active PET in index order, decoded from the packed keys
__global__ void __launch_bounds__(256) coords_kernel(const u64* __restrict__ keys, GridParams P, int* __restrict__ xs, int* __restrict__ ys) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P.n_act) return;
    const u64 k = keys[i] & KEY_MASK;
    const long long u = (long long)((u32)(k >> P.be) & P.umask) + P.ubase;
    const long long v = (long long)(k >> P.sshift) * P.eps + ((u32)k & P.emask) + P.vbase;
    xs[i] = (int)((u + v) >> 1);
    ys[i] = (int)((v - u) >> 1);
}

int index_coords(cloops_index* ix, int* d_xs, int* d_ys, cudaStream_t st) {
    const GridParams& P = ix->P;
    if (P.n_act == 0) return 0;
    LAUNCH(coords_kernel, cdiv(P.n_act, 256), 256, 0, st, ix->keys, P, d_xs, d_ys);
    return 0;
}

static int bits_for(u64 v) {  // number of bits needed to represent values 0..v
    int b = 0;
    while (v) { ++b; v >>= 1; }
    return b;
}

static inline long long floor_div(long long a, long long b) {
    long long q = a / b, r = a % b;
    return (r != 0 && ((r < 0) != (b < 0))) ? q - 1 : q;
}

int index_build(const int32_t* d_x, const int32_t* d_y, int64_t n, int32_t eps, int32_t cut, cloops_index** out,
                cudaStream_t st) {
    if (n < 0 || n > 0x7fffff00LL) return fail(CLOOPS_EINVAL, "n=%lld out of range", (long long)n);
    if (eps < 1) return fail(CLOOPS_EINVAL, "eps must be >= 1 (got %d)", eps);
    RET_IF(pool_init());
    cloops_index* ix = new cloops_index();
    GridParams& P = ix->P;
    memset(&P, 0, sizeof(P));
    P.eps = eps;
    P.n = (int)n;
    *out = ix;
    if (n == 0) return 0;

    Temp tmp(st);
    Extents* d_ext;
    RET_IF(tmp.alloc(&d_ext, 1));
    LAUNCH(extents_init_kernel, 1, 1, 0, st, d_ext);
    int grid = std::min(cdiv(n, 256), 148 * 8);
    LAUNCH(extents_kernel, grid, 256, 0, st, d_x, d_y, (int)n, cut, d_ext);
    Extents ext;
    CU_TRY(cudaMemcpyAsync(&ext, d_ext, sizeof(ext), cudaMemcpyDeviceToHost, st));
    CU_TRY(cudaStreamSynchronize(st));
    stage_mark("extents", st);
    if (ext.overflow) return fail(CLOOPS_ERANGE, "coordinates must lie in [-2^30, 2^30)");
    P.n_act = ext.n_act;
    if (P.n_act == 0) {
        P.ns = 0;
        return 0;
    }
    // coordinates must keep u = X-Y and v = X+Y inside int32 (callers guard |X|,|Y| < 2^30)
    long long ubase = floor_div(ext.umin, eps) * (long long)eps;
    long long vbase = floor_div(ext.vmin, eps) * (long long)eps;
    long long uspan = (long long)ext.umax - ubase;     // max u'
    long long vspan = (long long)ext.vmax - vbase;     // max v'
    long long ns = vspan / eps + 1;
    if (ubase < INT_MIN || vbase < INT_MIN || uspan > 0x7fffffffLL || vspan > 0x7fffffffLL)
        return fail(CLOOPS_ERANGE, "coordinate span too large for int32 rotated coordinates");
    if (ns > (1LL << 27))
        return fail(CLOOPS_ERANGE, "eps=%d too small for coordinate span %lld: %lld strips (limit 2^27)", eps, vspan, ns);
    P.ubase = (int)ubase;
    P.vbase = (int)vbase;
    P.ns = (int)ns;
    P.be = bits_for((u64)eps - 1);
    P.bu = std::max(1, bits_for((u64)uspan));
    P.bs = std::max(1, bits_for((u64)ns));             // value ns itself = sentinel strip
    P.sshift = P.be + P.bu;
    if (P.be + P.bu + P.bs > 63)
        return fail(CLOOPS_ERANGE, "packed key needs %d bits (> 63)", P.be + P.bu + P.bs);
    P.emask = (P.be == 0) ? 0u : (u32)((1ull << P.be) - 1);
    P.umask = (u32)((1ull << P.bu) - 1);

    u64 *k0, *k1;
    u32 *r0, *r1;
    CU_TRY(cudaMallocAsync((void**)&k1, n * sizeof(u64), st));
    CU_TRY(cudaMallocAsync((void**)&r1, n * sizeof(u32), st));
    ix->keys = k1;
    ix->rows = r1;
    RET_IF(tmp.alloc(&k0, n));
    RET_IF(tmp.alloc(&r0, n));
    LAUNCH(pack_kernel, cdiv(n, 256), 256, 0, st, d_x, d_y, cut, P, k0, r0);
    stage_mark("pack", st);
    size_t sort_bytes = 0;
    int begin_bit = P.be, end_bit = P.be + P.bu + P.bs;
    CU_TRY(cub::DeviceRadixSort::SortPairs(nullptr, sort_bytes, k0, k1, r0, r1, (int)n, begin_bit, end_bit, st));
    void* d_sort;
    RET_IF(tmp.alloc((char**)&d_sort, sort_bytes));
    CU_TRY(cub::DeviceRadixSort::SortPairs(d_sort, sort_bytes, k0, k1, r0, r1, (int)n, begin_bit, end_bit, st));
    stage_mark("sort", st);
    CU_TRY(cudaMallocAsync((void**)&ix->sstart, (size_t)(P.ns + 3) * sizeof(int), st));
    if ((long long)P.ns > 4LL * P.n_act + 1024) {
        LAUNCH(strip_table_search_kernel, cdiv(P.ns + 3, 256), 256, 0, st, ix->keys, P, ix->sstart);
    } else {
        LAUNCH(strip_table_gap_kernel, cdiv(P.n_act + 1, 256), 256, 0, st, ix->keys, P, ix->sstart);
    }
    stage_mark("strips", st);
    return 0;
}

void index_free(cloops_index* ix, cudaStream_t st) {
    if (!ix) return;
    if (ix->keys) cudaFreeAsync(ix->keys, st);
    if (ix->rows) cudaFreeAsync(ix->rows, st);
    if (ix->sstart) cudaFreeAsync(ix->sstart, st);
    delete ix;
}

int index_count(cloops_index* ix, int cap, int* d_counts_sorted, cudaStream_t st) {
    const GridParams& P = ix->P;
    if (P.n_act == 0) return 0;
    if (cap <= 0) cap = INT_MAX;
    LAUNCH(count_kernel, cdiv(P.n_act, CQ_TILE), CQ_TILE, 0, st, ix->keys, ix->sstart, P, cap, d_counts_sorted);
    return 0;
}

}  // namespace cloops
