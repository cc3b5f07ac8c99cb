// blockDBSCAN (cLoops/blockDBSCAN.py:13-239) as data-parallel rules over a cell-sorted layout.
//
//   cell(p)   = (floor((X-minX)/eps)+1, floor((Y-minY)/eps)+1)                     (:74-86)
//   centroid  = (floor(sumX/n), floor(sumY/n))  (py2 integer division)              (:124-140)
//   8-adjacent cells a,b are CONNECTED iff d1(cen a, cen b) <= eps or some point pair is within eps
//                                                                                   (:204-213,227-238)
//   core cell = own count + counts of connected neighbours >= minPts                (:181,191)
//   clusters  = components of core cells, numbered by the first-inserted core cell  (:142-152)
//   a non-core cell takes the LARGEST id among its connected core neighbours        (:188-198)
//   every point inherits its cell's id; no size filter                              (:154-168)
#include <limits.h>
#include <string.h>

#include <algorithm>
#include <cub/cub.cuh>

#include "common.cuh"
#include "scan.cuh"

namespace cloops {

struct BExt {
    int minx, maxx, miny, maxy, n_act, overflow;
};

struct BParams {
    int eps, minx, miny, bx, n, n_act, ncy;
};

struct Cells {
    u64* ckey;
    int* cstart;     // [ncell+1]
    int* cnt;
    long long* sumx;
    long long* sumy;
    int* cfirst;     // smallest row in the cell (dict insertion order of the reference)
    int* minS; int* maxS; int* minD; int* maxD;   // extremes of x+y and x-y
    int* near;
    int* parent;
    int* rank;       // per root: min cfirst over its core cells
    int* best;       // non-core cells: largest rank among connected core neighbours
};

__global__ void bext_init_kernel(BExt* e) {
    e->minx = INT_MAX; e->maxx = INT_MIN; e->miny = INT_MAX; e->maxy = INT_MIN; e->n_act = 0; e->overflow = 0;
}

__global__ void __launch_bounds__(256) bext_kernel(const int* __restrict__ x, const int* __restrict__ y, int n, int cut, BExt* out) {
    int x0 = INT_MAX, x1 = INT_MIN, y0 = INT_MAX, y1 = INT_MIN, cnt = 0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        int xx = __ldg(x + i), yy = __ldg(y + i);
        if (xx < -(1 << 30) || xx >= (1 << 30) || yy < -(1 << 30) || yy >= (1 << 30)) { out->overflow = 1; continue; }
        if (cut > 0 && yy - xx < cut) continue;
        x0 = min(x0, xx); x1 = max(x1, xx); y0 = min(y0, yy); y1 = max(y1, yy);
        ++cnt;
    }
    x0 = __reduce_min_sync(0xffffffffu, x0); x1 = __reduce_max_sync(0xffffffffu, x1);
    y0 = __reduce_min_sync(0xffffffffu, y0); y1 = __reduce_max_sync(0xffffffffu, y1);
    cnt = __reduce_add_sync(0xffffffffu, cnt);
    if ((threadIdx.x & 31) == 0 && cnt > 0) {
        atomicMin(&out->minx, x0); atomicMax(&out->maxx, x1); atomicMin(&out->miny, y0); atomicMax(&out->maxy, y1);
        atomicAdd(&out->n_act, cnt);
    }
}

__global__ void __launch_bounds__(256) bpack_kernel(const int* __restrict__ x, const int* __restrict__ y, int cut, BParams P,
                                                    u64* __restrict__ keys, u32* __restrict__ rows) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P.n) return;
    int xx = __ldg(x + i), yy = __ldg(y + i);
    u64 key;
    if (cut > 0 && yy - xx < cut) {
        key = (u64)P.ncy << P.bx;
    } else {
        u32 cx = (u32)(xx - P.minx) / (u32)P.eps;
        u32 cy = (u32)(yy - P.miny) / (u32)P.eps;
        key = ((u64)cy << P.bx) | cx;
    }
    keys[i] = key;
    rows[i] = (u32)i;
}

__global__ void __launch_bounds__(256) bhead_kernel(const u64* __restrict__ keys, int n_act, int* __restrict__ head) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_act) return;
    head[i] = (i == 0 || keys[i] != keys[i - 1]) ? 1 : 0;
}

__global__ void __launch_bounds__(256) bcell_init_kernel(Cells C, int ncell) {
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= ncell) return;
    C.cnt[c] = 0; C.sumx[c] = 0; C.sumy[c] = 0; C.cfirst[c] = INT_MAX;
    C.minS[c] = INT_MAX; C.maxS[c] = INT_MIN; C.minD[c] = INT_MAX; C.maxD[c] = INT_MIN;
    C.parent[c] = c; C.rank[c] = INT_MAX; C.best[c] = -1;
}

__global__ void __launch_bounds__(256) bcell_fill_kernel(const u64* __restrict__ keys, const u32* __restrict__ rows,
                                                         const int* __restrict__ x, const int* __restrict__ y,
                                                         const int* __restrict__ cellidx, int n_act, Cells C,
                                                         int* __restrict__ px, int* __restrict__ py) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_act) return;
    int c = cellidx[i] - 1;                     // inclusive scan of head flags
    int r = (int)rows[i];
    int xx = __ldg(x + r), yy = __ldg(y + r);
    px[i] = xx; py[i] = yy;
    if (i == 0 || keys[i] != keys[i - 1]) { C.cstart[c] = i; C.ckey[c] = keys[i]; }
    if (i == n_act - 1) C.cstart[c + 1] = n_act;
    atomicAdd(&C.cnt[c], 1);
    atomicAdd((unsigned long long*)&C.sumx[c], (unsigned long long)(long long)xx);
    atomicAdd((unsigned long long*)&C.sumy[c], (unsigned long long)(long long)yy);
    atomicMin(&C.cfirst[c], r);
    atomicMin(&C.minS[c], xx + yy); atomicMax(&C.maxS[c], xx + yy);
    atomicMin(&C.minD[c], xx - yy); atomicMax(&C.maxD[c], xx - yy);
}

__device__ __forceinline__ long long fdiv(long long a, long long b) {   // floor division, b > 0
    long long q = a / b;
    return (a % b != 0 && a < 0) ? q - 1 : q;
}

__device__ __forceinline__ bool centroid_close(const Cells& C, int a, int b, int eps) {
    long long ax = fdiv(C.sumx[a], C.cnt[a]), ay = fdiv(C.sumy[a], C.cnt[a]);
    long long bx = fdiv(C.sumx[b], C.cnt[b]), by = fdiv(C.sumy[b], C.cnt[b]);
    return llabs(ax - bx) + llabs(ay - by) <= (long long)eps;
}

// forward neighbours (E, NW, N, NE) of every cell; decides centroid / diagonal cases at once and
// queues the E and N pairs that need a point-pair test
__global__ void __launch_bounds__(256) bedges_kernel(Cells C, int ncell, BParams P, int* __restrict__ ea, int* __restrict__ eb,
                                                     int* __restrict__ pa, int* __restrict__ pb, int* __restrict__ counters) {
    int a = blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= ncell) return;
    const u64 key = C.ckey[a];
    const u64 xmask = (1ull << P.bx) - 1;
    const long long cx = (long long)(key & xmask), cy = (long long)(key >> P.bx);
    auto consider = [&](int b, int dir) {       // dir: 0 E, 1 NW, 2 N, 3 NE
        bool conn = centroid_close(C, a, b, P.eps);
        bool pending = false;
        if (!conn) {
            if (dir == 3) conn = (long long)C.minS[b] - (long long)C.maxS[a] <= (long long)P.eps;
            else if (dir == 1) conn = (long long)C.minD[a] - (long long)C.maxD[b] <= (long long)P.eps;
            else pending = true;
        }
        if (conn) { int t = atomicAdd(&counters[0], 1); ea[t] = a; eb[t] = b; }
        if (pending) { int t = atomicAdd(&counters[1], 1); pa[t] = a; pb[t] = b; }
    };
    if (a + 1 < ncell) {
        u64 kn = C.ckey[a + 1];
        if ((long long)(kn >> P.bx) == cy && (long long)(kn & xmask) == cx + 1) consider(a + 1, 0);
    }
    // row cy+1, columns cx-1..cx+1
    long long lo_cx = cx > 0 ? cx - 1 : 0;
    u64 target = ((u64)(cy + 1) << P.bx) | (u64)lo_cx;
    int lo = a + 1, hi = ncell;
    while (lo < hi) { int mid = (lo + hi) >> 1; if (C.ckey[mid] < target) lo = mid + 1; else hi = mid; }
    for (int b = lo; b < ncell && b < lo + 3; ++b) {
        u64 kn = C.ckey[b];
        if ((long long)(kn >> P.bx) != cy + 1) break;
        long long dx = (long long)(kn & xmask) - cx;
        if (dx > 1) break;
        if (dx >= -1) consider(b, dx == -1 ? 1 : (dx == 0 ? 2 : 3));
    }
}

// one warp per queued pair: any p in a, q in b with |dx|+|dy| <= eps ?  (blockDBSCAN.py:204-213)
__global__ void __launch_bounds__(256) bpairs_kernel(Cells C, const int* __restrict__ px, const int* __restrict__ py, int eps,
                                                     const int* __restrict__ pa, const int* __restrict__ pb, int npend,
                                                     int* __restrict__ ea, int* __restrict__ eb, int* __restrict__ counters) {
    int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (w >= npend) return;
    int a = pa[w], b = pb[w];
    int a0 = C.cstart[a], a1 = C.cstart[a + 1], b0 = C.cstart[b], b1 = C.cstart[b + 1];
    if (a1 - a0 < b1 - b0) { int t; t = a0; a0 = b0; b0 = t; t = a1; a1 = b1; b1 = t; }   // lanes over the larger cell
    bool hit = false;
    for (int q = b0; q < b1 && !hit; ++q) {
        int qx = px[q], qy = py[q];
        for (int p = a0 + lane; p - lane < a1; p += 32) {
            bool h = false;
            if (p < a1) h = (long long)abs(px[p] - qx) + (long long)abs(py[p] - qy) <= (long long)eps;
            if (__any_sync(0xffffffffu, h)) { hit = true; break; }
        }
    }
    if (hit && lane == 0) { int t = atomicAdd(&counters[0], 1); ea[t] = pa[w]; eb[t] = pb[w]; }
}

__global__ void __launch_bounds__(256) bnear_init_kernel(Cells C, int ncell) {
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c < ncell) C.near[c] = C.cnt[c];
}

__global__ void __launch_bounds__(256) bnear_kernel(Cells C, const int* __restrict__ ea, const int* __restrict__ eb, int ne) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= ne) return;
    int a = ea[t], b = eb[t];
    atomicAdd(&C.near[a], C.cnt[b]);
    atomicAdd(&C.near[b], C.cnt[a]);
}

__device__ __forceinline__ int bfind(int* parent, int x) {
    int p = parent[x];
    while (p != x) { int g = parent[p]; if (g != p) parent[x] = g; x = p; p = g; }
    return x;
}

// read-only root walk: path halving by concurrent threads may leave parent[] one hop short of the root
__device__ __forceinline__ int broot(const int* parent, int x) {
    int p = parent[x];
    while (p != x) { x = p; p = parent[x]; }
    return x;
}

__global__ void __launch_bounds__(256) bunion_kernel(Cells C, const int* __restrict__ ea, const int* __restrict__ eb, int ne, int minPts) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= ne) return;
    int a = ea[t], b = eb[t];
    if (C.near[a] < minPts || C.near[b] < minPts) return;
    while (true) {
        a = bfind(C.parent, a); b = bfind(C.parent, b);
        if (a == b) return;
        if (a < b) { int s = a; a = b; b = s; }
        if (atomicCAS(&C.parent[a], a, b) == a) return;
    }
}

__global__ void __launch_bounds__(256) bcompress_kernel(Cells C, int ncell, int minPts, int* __restrict__ counters) {
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= ncell) return;
    if (C.near[c] < minPts) return;
    int r = bfind(C.parent, c);
    C.parent[c] = r;
    atomicMin(&C.rank[r], C.cfirst[c]);
    atomicAdd(&counters[3], 1);
    if (r == c) atomicAdd(&counters[2], 1);
}

__global__ void __launch_bounds__(256) bborder_kernel(Cells C, const int* __restrict__ ea, const int* __restrict__ eb, int ne, int minPts,
                                                      int* __restrict__ flags) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= ne) return;
    int a = ea[t], b = eb[t];
    bool ca = C.near[a] >= minPts, cb = C.near[b] >= minPts;
    if (ca && !cb) atomicMax(&C.best[b], C.rank[broot(C.parent, a)]);
    if (cb && !ca) atomicMax(&C.best[a], C.rank[broot(C.parent, b)]);
}

__global__ void __launch_bounds__(256) bflags_kernel(Cells C, int ncell, int minPts, int* __restrict__ flags) {
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= ncell) return;
    if (C.near[c] >= minPts && C.parent[c] == c) flags[C.rank[c]] = 1;
}

__global__ void __launch_bounds__(256) blabel_kernel(Cells C, const int* __restrict__ cellidx, const u32* __restrict__ rows, int n_act,
                                                     int minPts, const int* __restrict__ ids, int* __restrict__ labels,
                                                     int* __restrict__ counters) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_act) return;
    int c = cellidx[i] - 1;
    int lab = -1;
    if (C.near[c] >= minPts) lab = ids[C.rank[broot(C.parent, c)]];
    else if (C.best[c] >= 0) lab = ids[C.best[c]];
    labels[rows[i]] = lab;
    if (lab >= 0) atomicAdd(&counters[4], 1);
}

__global__ void __launch_bounds__(256) bfill_kernel(int* p, int v, long long n) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

static int bits_for(u64 v) { int b = 0; while (v) { ++b; v >>= 1; } return b; }

int block_dbscan(const int32_t* d_x, const int32_t* d_y, int64_t n, int32_t eps, int32_t minPts, int32_t cut, int32_t* d_labels,
                 int64_t* h_info, cudaStream_t st) {
    if (n < 0 || n > 0x7fffff00LL) return fail(CLOOPS_EINVAL, "n=%lld out of range", (long long)n);
    if (eps < 1) return fail(CLOOPS_EINVAL, "eps must be >= 1 (got %d)", eps);
    if (minPts < 1) return fail(CLOOPS_EINVAL, "minPts must be >= 1 (got %d)", minPts);
    if (h_info) for (int k = 0; k < 8; ++k) h_info[k] = 0;
    if (n == 0) return 0;
    RET_IF(pool_init());
    Temp tmp(st);
    LAUNCH(bfill_kernel, cdiv(n, 256), 256, 0, st, d_labels, -1, (long long)n);
    BExt* d_ext;
    RET_IF(tmp.alloc(&d_ext, 1));
    LAUNCH(bext_init_kernel, 1, 1, 0, st, d_ext);
    LAUNCH(bext_kernel, std::min(cdiv(n, 256), 148 * 8), 256, 0, st, d_x, d_y, (int)n, cut, d_ext);
    BExt ext;
    CU_TRY(cudaMemcpyAsync(&ext, d_ext, sizeof(ext), cudaMemcpyDeviceToHost, st));
    CU_TRY(cudaStreamSynchronize(st));
    if (ext.overflow) return fail(CLOOPS_ERANGE, "coordinates must lie in [-2^30, 2^30)");
    if (ext.n_act == 0) return 0;
    stage_mark("extents", st);
    BParams P;
    P.eps = eps; P.minx = ext.minx; P.miny = ext.miny; P.n = (int)n; P.n_act = ext.n_act;
    long long ncx = ((long long)ext.maxx - ext.minx) / eps + 1, ncy = ((long long)ext.maxy - ext.miny) / eps + 1;
    P.bx = std::max(1, bits_for((u64)ncx));
    int by = std::max(1, bits_for((u64)ncy));
    P.ncy = (int)ncy;
    const int na = P.n_act;

    u64 *k0, *k1;
    u32 *r0, *r1;
    RET_IF(tmp.alloc(&k0, n)); RET_IF(tmp.alloc(&k1, n)); RET_IF(tmp.alloc(&r0, n)); RET_IF(tmp.alloc(&r1, n));
    LAUNCH(bpack_kernel, cdiv(n, 256), 256, 0, st, d_x, d_y, cut, P, k0, r0);
    size_t bytes = 0;
    CU_TRY(cub::DeviceRadixSort::SortPairs(nullptr, bytes, k0, k1, r0, r1, (int)n, 0, P.bx + by, st));
    void* d_tmp;
    RET_IF(tmp.alloc((char**)&d_tmp, bytes));
    CU_TRY(cub::DeviceRadixSort::SortPairs(d_tmp, bytes, k0, k1, r0, r1, (int)n, 0, P.bx + by, st));
    stage_mark("sort", st);

    int *head, *cellidx, *px, *py, *counters;
    RET_IF(tmp.alloc(&head, na)); RET_IF(tmp.alloc(&cellidx, na)); RET_IF(tmp.alloc(&px, na)); RET_IF(tmp.alloc(&py, na));
    RET_IF(tmp.alloc(&counters, 8));
    CU_TRY(cudaMemsetAsync(counters, 0, 8 * sizeof(int), st));
    LAUNCH(bhead_kernel, cdiv(na, 256), 256, 0, st, k1, na, head);
    int* d_scan;
    RET_IF(tmp.alloc(&d_scan, scan_tmp_ints((long long)n + 1)));
    RET_IF((device_scan<SCAN_ADD, true, false>(head, cellidx, na, d_scan, st)));
    int ncell = 0;
    CU_TRY(cudaMemcpyAsync(&ncell, cellidx + na - 1, sizeof(int), cudaMemcpyDeviceToHost, st));
    CU_TRY(cudaStreamSynchronize(st));

    Cells C;
    RET_IF(tmp.alloc(&C.ckey, ncell)); RET_IF(tmp.alloc(&C.cstart, (size_t)ncell + 1)); RET_IF(tmp.alloc(&C.cnt, ncell));
    RET_IF(tmp.alloc(&C.sumx, ncell)); RET_IF(tmp.alloc(&C.sumy, ncell)); RET_IF(tmp.alloc(&C.cfirst, ncell));
    RET_IF(tmp.alloc(&C.minS, ncell)); RET_IF(tmp.alloc(&C.maxS, ncell)); RET_IF(tmp.alloc(&C.minD, ncell));
    RET_IF(tmp.alloc(&C.maxD, ncell)); RET_IF(tmp.alloc(&C.near, ncell)); RET_IF(tmp.alloc(&C.parent, ncell));
    RET_IF(tmp.alloc(&C.rank, ncell)); RET_IF(tmp.alloc(&C.best, ncell));
    const int gc = cdiv(ncell, 256);
    LAUNCH(bcell_init_kernel, gc, 256, 0, st, C, ncell);
    LAUNCH(bcell_fill_kernel, cdiv(na, 256), 256, 0, st, k1, r1, d_x, d_y, cellidx, na, C, px, py);
    stage_mark("cells", st);

    int *ea, *eb, *pa, *pb;
    RET_IF(tmp.alloc(&ea, (size_t)4 * ncell)); RET_IF(tmp.alloc(&eb, (size_t)4 * ncell));
    RET_IF(tmp.alloc(&pa, (size_t)4 * ncell)); RET_IF(tmp.alloc(&pb, (size_t)4 * ncell));
    LAUNCH(bedges_kernel, gc, 256, 0, st, C, ncell, P, ea, eb, pa, pb, counters);
    int hc[8];
    CU_TRY(cudaMemcpyAsync(hc, counters, sizeof(hc), cudaMemcpyDeviceToHost, st));
    CU_TRY(cudaStreamSynchronize(st));
    int npend = hc[1];
    if (npend > 0) LAUNCH(bpairs_kernel, cdiv((long long)npend * 32, 256), 256, 0, st, C, px, py, eps, pa, pb, npend, ea, eb, counters);
    CU_TRY(cudaMemcpyAsync(hc, counters, sizeof(hc), cudaMemcpyDeviceToHost, st));
    CU_TRY(cudaStreamSynchronize(st));
    int ne = hc[0];
    stage_mark("edges", st);

    int *flags, *ids;
    RET_IF(tmp.alloc(&flags, (size_t)n + 1)); RET_IF(tmp.alloc(&ids, (size_t)n + 1));
    CU_TRY(cudaMemsetAsync(flags, 0, ((size_t)n + 1) * sizeof(int), st));
    LAUNCH(bnear_init_kernel, gc, 256, 0, st, C, ncell);
    if (ne > 0) {
        LAUNCH(bnear_kernel, cdiv(ne, 256), 256, 0, st, C, ea, eb, ne);
        LAUNCH(bunion_kernel, cdiv(ne, 256), 256, 0, st, C, ea, eb, ne, minPts);
    }
    LAUNCH(bcompress_kernel, gc, 256, 0, st, C, ncell, minPts, counters);
    if (ne > 0) LAUNCH(bborder_kernel, cdiv(ne, 256), 256, 0, st, C, ea, eb, ne, minPts, flags);
    LAUNCH(bflags_kernel, gc, 256, 0, st, C, ncell, minPts, flags);
    RET_IF((device_scan<SCAN_ADD, false, false>(flags, ids, (long long)n + 1, d_scan, st)));
    LAUNCH(blabel_kernel, cdiv(na, 256), 256, 0, st, C, cellidx, r1, na, minPts, ids, d_labels, counters);
    stage_mark("labels", st);
    if (h_info) {
        int ncl = 0;
        CU_TRY(cudaMemcpyAsync(hc, counters, sizeof(hc), cudaMemcpyDeviceToHost, st));
        CU_TRY(cudaMemcpyAsync(&ncl, ids + n, sizeof(int), cudaMemcpyDeviceToHost, st));
        CU_TRY(cudaStreamSynchronize(st));
        h_info[0] = na; h_info[1] = ncl; h_info[2] = hc[2]; h_info[3] = hc[3]; h_info[4] = 0; h_info[5] = ncell;
        h_info[6] = P.bx + by; h_info[7] = hc[4];
    }
    return 0;
}

}  // namespace cloops
