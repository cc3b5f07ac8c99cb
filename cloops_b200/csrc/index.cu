// Index build: rotate -> pack -> order by (strip, u') (counting sort by strip, or radix sort) -> strip table -> tile headers.
// The region-query kernel that runs over the index lives in region_query.cu.
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
#include "scan.cuh"

namespace cloops {

struct Extents {
    int umin, umax, vmin, vmax, n_act, overflow;
};

__global__ void extents_init_kernel(Extents* e) {
    e->umin = INT_MAX; e->umax = INT_MIN; e->vmin = INT_MAX; e->vmax = INT_MIN; e->n_act = 0; e->overflow = 0;
}

__device__ __forceinline__ void extents_point(int xx, int yy, int cut, int& umin, int& umax, int& vmin, int& vmax, int& cnt, int& bad) {
    if (xx < -(1 << 30) || xx >= (1 << 30) || yy < -(1 << 30) || yy >= (1 << 30)) { bad = 1; return; }   // u, v would leave int32
    if (cut > 0 && yy - xx < cut) return;
    const int u = xx - yy, v = xx + yy;
    umin = min(umin, u); umax = max(umax, u); vmin = min(vmin, v); vmax = max(vmax, v);
    ++cnt;
}

// Rows are looked at in tiles of 1024 consecutive rows (four per thread, one 128-bit load each when x and y are 16-byte
// aligned); the number of a tile's rows that pass the cut filter goes to blockcnt[tile] -- the input of the stable compaction
// below.  A CTA takes EX_TILES consecutive tiles, so the five global atomics of the extents are paid once per 8192 rows.
#define EX_ROWS 1024
#define EX_TILES 8
__global__ void __launch_bounds__(256) extents_kernel(const int* __restrict__ x, const int* __restrict__ y, int n, int cut, int vec,
                                                      Extents* out, int* __restrict__ blockcnt) {
    __shared__ int s_cnt[8];
    __shared__ int s_ext[8][4];
    int umin = INT_MAX, umax = INT_MIN, vmin = INT_MAX, vmax = INT_MIN, bad = 0, total = 0;
    const int ntile = (n + EX_ROWS - 1) / EX_ROWS;
    for (int tile = blockIdx.x * EX_TILES; tile < min((int)(blockIdx.x + 1) * EX_TILES, ntile); ++tile) {
        int cnt = 0;
        const int i0 = tile * EX_ROWS + 4 * threadIdx.x;
        if (vec && i0 + 3 < n) {
            const int4 a = __ldg(reinterpret_cast<const int4*>(x + i0)), b = __ldg(reinterpret_cast<const int4*>(y + i0));
            extents_point(a.x, b.x, cut, umin, umax, vmin, vmax, cnt, bad);
            extents_point(a.y, b.y, cut, umin, umax, vmin, vmax, cnt, bad);
            extents_point(a.z, b.z, cut, umin, umax, vmin, vmax, cnt, bad);
            extents_point(a.w, b.w, cut, umin, umax, vmin, vmax, cnt, bad);
        } else {
            for (int i = i0; i < min(i0 + 4, n); ++i) extents_point(__ldg(x + i), __ldg(y + i), cut, umin, umax, vmin, vmax, cnt, bad);
        }
        cnt = __reduce_add_sync(0xffffffffu, cnt);
        if ((threadIdx.x & 31) == 0) s_cnt[threadIdx.x >> 5] = cnt;
        __syncthreads();
        if (threadIdx.x == 0) {
            int tot = 0;
            for (int w = 0; w < 8; ++w) tot += s_cnt[w];
            blockcnt[tile] = tot;
            total += tot;
        }
        __syncthreads();
    }
    if (bad) out->overflow = 1;
    umin = __reduce_min_sync(0xffffffffu, umin);
    umax = __reduce_max_sync(0xffffffffu, umax);
    vmin = __reduce_min_sync(0xffffffffu, vmin);
    vmax = __reduce_max_sync(0xffffffffu, vmax);
    if ((threadIdx.x & 31) == 0) {
        int* e = s_ext[threadIdx.x >> 5];
        e[0] = umin; e[1] = umax; e[2] = vmin; e[3] = vmax;
    }
    __syncthreads();
    if (threadIdx.x == 0 && total > 0) {
        for (int w = 0; w < 8; ++w) {
            umin = min(umin, s_ext[w][0]); umax = max(umax, s_ext[w][1]); vmin = min(vmin, s_ext[w][2]); vmax = max(vmax, s_ext[w][3]);
        }
        atomicMin(&out->umin, umin); atomicMax(&out->umax, umax);
        atomicMin(&out->vmin, vmin); atomicMax(&out->vmax, vmax);
        atomicAdd(&out->n_act, total);
    }
}

// exclusive scan of the per-CTA counts, in place (one CTA; m = rows / 1024 entries)
__global__ void __launch_bounds__(1024) blockcnt_scan_kernel(int* __restrict__ a, int m) {
    __shared__ int s_warp[32];
    __shared__ int s_carry;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    for (int base = 0; base < m; base += 1024) {
        const int i = base + threadIdx.x;
        const int v = i < m ? a[i] : 0;
        int incl = v;
        for (int d = 1; d < 32; d <<= 1) {
            const int y = __shfl_up_sync(0xffffffffu, incl, d);
            if ((threadIdx.x & 31) >= d) incl += y;
        }
        if ((threadIdx.x & 31) == 31) s_warp[threadIdx.x >> 5] = incl;
        __syncthreads();
        if (threadIdx.x < 32) {
            int w = s_warp[threadIdx.x];
            for (int d = 1; d < 32; d <<= 1) {
                const int y = __shfl_up_sync(0xffffffffu, w, d);
                if ((int)threadIdx.x >= d) w += y;
            }
            s_warp[threadIdx.x] = w;
        }
        __syncthreads();
        const int before = s_carry + (threadIdx.x >= 32 ? s_warp[(threadIdx.x >> 5) - 1] : 0) + incl - v;
        if (i < m) a[i] = before;
        __syncthreads();
        if (threadIdx.x == 1023) s_carry = before + v;
        __syncthreads();
    }
}

// Keys of the rows that pass the cut filter only, compacted in ROW ORDER (stable): position = rows of earlier CTAs
// (blockbase, the scanned counts of extents_kernel) + earlier active rows of this CTA.  The radix sort then moves n_act
// rows instead of n -- after the first round of a Hi-C run the cut removes more than half of them.
__global__ void __launch_bounds__(256) pack_compact_kernel(const int* __restrict__ x, const int* __restrict__ y, int cut, GridParams P,
                                                           const int* __restrict__ blockbase, u64* __restrict__ keys,
                                                           u32* __restrict__ rows) {
    __shared__ int s_warp[8];
    const int i0 = blockIdx.x * EX_ROWS + 4 * threadIdx.x;
    u64 key[4];
    unsigned act = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int i = i0 + k;
        if (i >= P.n) continue;
        const int xx = __ldg(x + i), yy = __ldg(y + i);
        if (yy - xx < cut) continue;
        const u32 up = (u32)((xx - yy) - P.ubase), vp = (u32)((xx + yy) - P.vbase);
        const u32 sv = vp / (u32)P.eps, vm = vp - sv * (u32)P.eps;
        key[k] = ((u64)sv << P.sshift) | ((u64)up << P.be) | (u64)vm;
        act |= 1u << k;
    }
    const int mine = __popc(act), lane = threadIdx.x & 31;
    int incl = mine;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += v;
    }
    if (lane == 31) s_warp[threadIdx.x >> 5] = incl;
    __syncthreads();
    int pos = blockbase[blockIdx.x] + incl - mine;
    for (int w = 0; w < (int)(threadIdx.x >> 5); ++w) pos += s_warp[w];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        if (!(act & (1u << k))) continue;
        keys[pos] = key[k];
        rows[pos] = (u32)(i0 + k);
        ++pos;
    }
}

// Packs the key of every row.  RANK: also counts the rows of each strip (cnt[strip + 1], the layout of the strip
// table) and remembers each row's arrival rank inside its strip, so that one exclusive scan of cnt IS the strip
// table and the rows can be placed next to their strip without a radix sort (index_build).  Rows removed by the
// cut filter are ranked behind the active ones through cnt_tail.
template <bool RANK>
__global__ void __launch_bounds__(256) pack_kernel(const int* __restrict__ x, const int* __restrict__ y, int cut, GridParams P,
                                                   u64* __restrict__ keys, u32* __restrict__ rows_or_rank, int* __restrict__ cnt,
                                                   int* __restrict__ cnt_tail) {
    const int i0 = blockIdx.x * 1024 + threadIdx.x;       // four rows per thread, 256 apart: loads and atomics overlap
    int xx[4], yy[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int i = i0 + 256 * k;
        if (i < P.n) { xx[k] = __ldg(x + i); yy[k] = __ldg(y + i); }
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int i = i0 + 256 * k;
        if (i >= P.n) continue;
        u64 key;
        u32 aux = (u32)i;
        if (cut > 0 && yy[k] - xx[k] < cut) {
            key = (u64)P.ns << P.sshift;                  // sentinel strip: sorts behind every active row
            if (RANK) {
                // any distinct place in the tail will do (inactive rows are never looked at again): one atomic per warp,
                // not one per row -- after the first round more than half of all rows are removed by the cut
                const unsigned peers = __activemask();
                const int lane = threadIdx.x & 31, leader = __ffs(peers) - 1;
                int base = 0;
                if (lane == leader) base = atomicAdd(cnt_tail, __popc(peers));
                aux = (u32)(__shfl_sync(peers, base, leader) + __popc(peers & ((1u << lane) - 1)));
            }
        } else {
            u32 up = (u32)((xx[k] - yy[k]) - P.ubase);
            u32 vp = (u32)((xx[k] + yy[k]) - P.vbase);
            u32 sv = vp / (u32)P.eps;
            u32 vm = vp - sv * (u32)P.eps;
            key = ((u64)sv << P.sshift) | ((u64)up << P.be) | (u64)vm;
            if (RANK) aux = (u32)atomicAdd(cnt + sv + 1, 1);
        }
        keys[i] = key;
        rows_or_rank[i] = aux;
    }
}

// out[0] = sum over strips of (rows in the strip)^2 = the work of strip_rank_kernel; out[1] = longest strip (its critical path)
__global__ void __launch_bounds__(256) strip_sumsq_kernel(const int* __restrict__ cnt, int m, unsigned long long* __restrict__ out) {
    unsigned long long acc = 0, mx = 0;
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < m; k += gridDim.x * blockDim.x) {
        const unsigned long long c = (unsigned long long)cnt[k];
        acc += c * c;
        mx = max(mx, c);
    }
    for (int d = 16; d > 0; d >>= 1) {
        acc += __shfl_down_sync(0xffffffffu, acc, d);
        mx = max(mx, __shfl_down_sync(0xffffffffu, mx, d));
    }
    if ((threadIdx.x & 31) == 0 && acc) {
        atomicAdd(out, acc);
        atomicMax(out + 1, mx);
    }
}

// Establishes the destination window of one scatter launch in L2 with full-sector stores, so the 12-byte stores of
// the scatter merge into resident lines instead of each fetching its sector from DRAM first (measured: the sort stage
// of config 2 takes 0.32 ms with it and 0.39 ms without).
__global__ void __launch_bounds__(256) strip_window_clear_kernel(const int* __restrict__ sstart, GridParams P, int s_lo, int s_hi,
                                                                 u64* __restrict__ keys_out, u32* __restrict__ rows_out) {
    const int a = s_lo >= P.ns ? P.n_act : __ldg(sstart + s_lo + 1);
    const int b = s_hi > P.ns ? P.n : __ldg(sstart + s_hi + 1);
    for (int j = a + blockIdx.x * blockDim.x + threadIdx.x; j < b; j += gridDim.x * blockDim.x) {
        keys_out[j] = 0ull;
        rows_out[j] = 0u;
    }
}

// Row i goes to (start of its strip) + (its arrival rank): the strips are contiguous afterwards, in arrival order
// inside.  A scatter over the whole destination would turn every 12-byte store into its own DRAM sector
// read-modify-write (measured: 343 MB written for 120 MB of payload); so the destination is cut into nwin windows
// of a few dozen MB that stay in L2 until their sectors are complete, one launch per window, and each launch
// streams the keys past (evict-first loads) and stores only the rows of its window.
__global__ void __launch_bounds__(256) strip_scatter_kernel(const u64* __restrict__ keys_in, const u32* __restrict__ rank,
                                                            const int* __restrict__ sstart, GridParams P, int s_lo, int s_hi,
                                                            u64* __restrict__ keys_out, u32* __restrict__ rows_out) {
    const int i0 = blockIdx.x * 2048 + threadIdx.x;       // eight rows per thread, 256 apart
    u64 key[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const int i = i0 + 256 * k;
        key[k] = i < P.n ? __ldcs(keys_in + i) : ~0ull;
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const int i = i0 + 256 * k;
        const int s = (int)(key[k] >> P.sshift);
        if (i >= P.n || s < s_lo || s >= s_hi) continue;
        const int dest = (s >= P.ns ? P.n_act : __ldg(sstart + s + 1)) + (int)__ldcs(rank + i);
        keys_out[dest] = key[k];
        rows_out[dest] = (u32)i;
    }
}

// Final position inside the strip = number of rows of the strip that sort before this one by (u', row) -- exactly the
// order a stable radix sort of the (strip, u') bits over rows in row order produces.  A CTA owns 256 consecutive
// positions; the strips they belong to are one contiguous range, staged in shared memory as (u', row); lanes of a
// warp mostly share a strip, so the scan over the strip is a broadcast read.  A strip holds a few dozen rows at
// ChIA-PET / HiChIP density (index_build checks the total work first); ranges that do not fit are scanned in
// global memory.
#define SR_CAP 3072
#define SR_TILE 512
__global__ void __launch_bounds__(256) strip_rank_kernel(const u64* __restrict__ keys_in, const u32* __restrict__ rows_in,
                                                         const int* __restrict__ sstart, GridParams P, u64* __restrict__ keys_out,
                                                         u32* __restrict__ rows_out) {
    __shared__ u32 U[SR_CAP];
    __shared__ u32 R[SR_CAP];
    __shared__ int s_a0, s_b0;
    const int p0 = blockIdx.x * SR_TILE;
    if (p0 >= P.n_act) {                                   // rows behind n_act (cut filter) keep their place
        for (int p = p0 + threadIdx.x; p < min(p0 + SR_TILE, P.n); p += 256) { keys_out[p] = keys_in[p]; rows_out[p] = rows_in[p]; }
        return;
    }
    const int plast = min(p0 + SR_TILE, P.n_act) - 1;
    if (threadIdx.x == 0) s_a0 = __ldg(sstart + (int)(keys_in[p0] >> P.sshift) + 1);
    if (threadIdx.x == 32) s_b0 = __ldg(sstart + (int)(keys_in[plast] >> P.sshift) + 2);
    __syncthreads();
    const int a0 = s_a0, b0 = s_b0;
    const bool staged = b0 - a0 <= SR_CAP;                 // CTA-uniform
    if (staged) {
        for (int j = a0 + threadIdx.x; j < b0; j += 256) {
            U[j - a0] = (u32)(keys_in[j] >> P.be) & P.umask;
            R[j - a0] = rows_in[j];
        }
    }
    __syncthreads();
#pragma unroll
    for (int h = 0; h < SR_TILE / 256; ++h) {
        const int p = p0 + 256 * h + threadIdx.x;
        if (p >= P.n) break;
        const u64 key = keys_in[p];
        const u32 row = rows_in[p];
        int dest = p;
        if (p < P.n_act) {
            const int s = (int)(key >> P.sshift);
            const int a = __ldg(sstart + s + 1), b = __ldg(sstart + s + 2);
            const u32 my = (u32)(key >> P.be) & P.umask;
            int c = 0;
            if (staged) {
                int eq = 0;                                       // rows of the strip with the same u' (this one included)
#pragma unroll 4
                for (int j = a - a0; j < b - a0; ++j) {
                    const u32 uj = U[j];
                    asm("{\n\t.reg .pred p, q;\n\tsetp.lt.u32 p, %2, %3;\n\tsetp.eq.u32 q, %2, %3;\n\t@p add.s32 %0, %0, 1;\n\t@q add.s32 %1, %1, 1;\n\t}"
                        : "+r"(c), "+r"(eq) : "r"(uj), "r"(my));
                }
                if (eq > 1) {                                     // ties are rare: a second pass orders them by row
                    for (int j = a - a0; j < b - a0; ++j)
                        if (U[j] == my && R[j] < row) ++c;
                }
            } else {
                for (int j = a; j < b; ++j) {
                    const u32 uj = (u32)(keys_in[j] >> P.be) & P.umask;
                    if (uj < my) ++c;
                    else if (uj == my && rows_in[j] < row) ++c;
                }
            }
            dest = a + c;
        }
        keys_out[dest] = key;
        rows_out[dest] = row;
    }
}

__global__ void __launch_bounds__(256) iota_kernel(u32* __restrict__ rows, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) rows[i] = (u32)i;
}

// sstart[k] = first sorted index whose strip >= k-1, k in [0, ns+2]  (sparse tables: one binary search per entry)
__global__ void __launch_bounds__(256) strip_table_search_kernel(const u64* __restrict__ keys, GridParams P, int* __restrict__ sstart) {
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k > P.ns + 2) return;
    int r;
    if (k == 0) r = 0;
    else if (k - 1 >= P.ns) r = P.n_act;
    else r = lower_bound_su(keys, 0, P.n_act, (u64)(k - 1) << P.bu, P.be);
    sstart[k] = r;
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
    int* d_blockcnt;
    const int nblk = cdiv(n, EX_ROWS);
    RET_IF(tmp.alloc(&d_ext, 1));
    RET_IF(tmp.alloc(&d_blockcnt, nblk));
    LAUNCH(extents_init_kernel, 1, 1, 0, st, d_ext);
    const int vec = ((((uintptr_t)d_x) | ((uintptr_t)d_y)) & 15) == 0 ? 1 : 0;
    LAUNCH(extents_kernel, cdiv(nblk, EX_TILES), 256, 0, st, d_x, d_y, (int)n, cut, vec, d_ext, d_blockcnt);
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
    // u' is padded by eps on both sides (eps <= u' <= 2^bu-1-eps): a +-eps window around any point stays
    // inside the u field, so kernels may test "same strip and |du| <= eps" with one compare on (strip,u')
    long long ubase = floor_div(ext.umin, eps) * (long long)eps - eps;
    long long vbase = floor_div(ext.vmin, eps) * (long long)eps;
    long long uspan = (long long)ext.umax - ubase + eps;     // max u' + eps
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
    CU_TRY(cudaMallocAsync((void**)&k1, (n + 2) * sizeof(u64), st));   // +2: the region query stages with 128-bit loads
    CU_TRY(cudaMallocAsync((void**)&r1, n * sizeof(u32), st));
    ix->keys = k1;
    ix->rows = r1;
    RET_IF(tmp.alloc(&k0, n));
    RET_IF(tmp.alloc(&r0, n));
    CU_TRY(cudaMallocAsync((void**)&ix->sstart, (size_t)(P.ns + 3) * sizeof(int), st));
    const int begin_bit = P.be, end_bit = P.be + P.bu + P.bs;
    // Order wanted: (strip, u'), ties in row order.  When the strip table is not much larger than the data, a counting
    // sort by strip does it in light passes: pack (+ one atomic per row: strip histogram and arrival rank), an exclusive scan
    // of the histogram (= the strip table), a scatter next to the strip, and a rank inside the strip.
    // The rank pass costs (rows per strip)^2, so its total work is checked first; long strips (Hi-C density) and
    // sparse tables take the radix sort.
    bool counted = false;
    const char* knob = getenv("CLOOPS_INDEX_SORT");                                       // test / measurement knob
    const bool force_radix = knob != nullptr && strcmp(knob, "radix") == 0;
    const bool force_count = knob != nullptr && strcmp(knob, "count") == 0;
    // long strips (Hi-C density: more than ~50 rows per strip on average) go to the radix sort at once: the histogram and
    // arrival ranks of the counting attempt would be thrown away (measured break-even ~ 80 rows per strip, see below)
    const bool long_strips = (long long)P.n_act > 48LL * P.ns;
    if (!force_radix && (force_count || !long_strips) && (long long)P.ns <= 4LL * P.n_act + 1024) {
        u64* k2;
        u32* r2;
        unsigned long long* d_sumsq;
        RET_IF(tmp.alloc(&k2, n));
        RET_IF(tmp.alloc(&r2, n));
        RET_IF(tmp.alloc(&d_sumsq, 3));
        int* cnt = ix->sstart;                                   // scanned in place
        CU_TRY(cudaMemsetAsync(cnt, 0, (size_t)(P.ns + 3) * sizeof(int), st));
        CU_TRY(cudaMemsetAsync(d_sumsq, 0, 3 * sizeof(unsigned long long), st));
        LAUNCH(pack_kernel<true>, cdiv(n, 1024), 256, 0, st, d_x, d_y, cut, P, k0, r0, cnt, reinterpret_cast<int*>(d_sumsq + 2));
        stage_mark("pack", st);
        LAUNCH(strip_sumsq_kernel, std::min(cdiv(P.ns + 3, 256), 148 * 8), 256, 0, st, cnt, P.ns + 3, d_sumsq);
        int* d_scan;
        RET_IF(tmp.alloc(&d_scan, scan_tmp_ints(P.ns + 3)));
        RET_IF((device_scan<SCAN_ADD, false, false>(cnt, cnt, P.ns + 3, d_scan, st)));
        unsigned long long sumsq[2] = {0, 0};
        CU_TRY(cudaMemcpyAsync(sumsq, d_sumsq, sizeof(sumsq), cudaMemcpyDeviceToHost, st));
        CU_TRY(cudaStreamSynchronize(st));
        stage_mark("strips", st);
        // measured on B200: the counting path wins up to ~80 rows per strip (10 M rows, 20 per strip: 0.56 vs 0.83 ms for the
        // whole build; 41: 0.56 vs 0.70; 80: 1.17 vs 1.21; 160: 1.62 vs 1.21)
        if ((force_count || sumsq[0] <= 64ull * (unsigned long long)P.n_act) && sumsq[1] <= 4096ull) {
            // two 60 MB windows for 10 M rows measured best (24 MB: +0.07 ms of re-reads, one 120 MB window: +0.15 ms)
            const long long win_bytes = 60LL << 20;
            const int nwin = (int)std::min<long long>(8, std::max<long long>(1, (12LL * n + win_bytes - 1) / win_bytes));
            for (int w = 0; w < nwin; ++w) {                     // strips are evenly filled: equal strip ranges ~ equal bytes
                const int s_lo = (int)((long long)P.ns * w / nwin);
                const int s_hi = w + 1 == nwin ? P.ns + 1 : (int)((long long)P.ns * (w + 1) / nwin);
                LAUNCH(strip_window_clear_kernel, 148 * 4, 256, 0, st, ix->sstart, P, s_lo, s_hi, k2, r2);
                LAUNCH(strip_scatter_kernel, cdiv(n, 2048), 256, 0, st, k0, r0, ix->sstart, P, s_lo, s_hi, k2, r2);
            }
            LAUNCH(strip_rank_kernel, cdiv(n, SR_TILE), 256, 0, st, k2, r2, ix->sstart, P, k1, r1);
        } else {
            LAUNCH(iota_kernel, cdiv(n, 256), 256, 0, st, r0, (int)n);
            size_t sort_bytes = 0;
            CU_TRY(cub::DeviceRadixSort::SortPairs(nullptr, sort_bytes, k0, k1, r0, r1, (int)n, begin_bit, end_bit, st));
            void* d_sort;
            RET_IF(tmp.alloc((char**)&d_sort, sort_bytes));
            CU_TRY(cub::DeviceRadixSort::SortPairs(d_sort, sort_bytes, k0, k1, r0, r1, (int)n, begin_bit, end_bit, st));
        }
        stage_mark("sort", st);
        counted = true;
    }
    if (!counted) {
        int n_sort = (int)n;
        if (cut > 0 && P.n_act < P.n) {                        // only the rows that pass the cut filter are packed and sorted
            LAUNCH(blockcnt_scan_kernel, 1, 1024, 0, st, d_blockcnt, nblk);
            LAUNCH(pack_compact_kernel, nblk, 256, 0, st, d_x, d_y, cut, P, d_blockcnt, k0, r0);
            n_sort = P.n_act;
        } else {
            LAUNCH(pack_kernel<false>, cdiv(n, 1024), 256, 0, st, d_x, d_y, cut, P, k0, r0, (int*)nullptr, (int*)nullptr);
        }
        stage_mark("pack", st);
        size_t sort_bytes = 0;
        CU_TRY(cub::DeviceRadixSort::SortPairs(nullptr, sort_bytes, k0, k1, r0, r1, n_sort, begin_bit, end_bit, st));
        void* d_sort;
        RET_IF(tmp.alloc((char**)&d_sort, sort_bytes));
        CU_TRY(cub::DeviceRadixSort::SortPairs(d_sort, sort_bytes, k0, k1, r0, r1, n_sort, begin_bit, end_bit, st));
        stage_mark("sort", st);
        LAUNCH(strip_table_search_kernel, cdiv(P.ns + 3, 256), 256, 0, st, ix->keys, P, ix->sstart);
    }
    RET_IF(index_tiles(ix, st));
    stage_mark("strips", st);
    return 0;
}

// ---- index of a cut-filtered round derived from the full (cut = 0) index of the same chromosome and eps -----------------
// The (strip, u', row) order does not depend on the cut: the rows a round keeps (Y - X >= cut, cLoops/pipe.py:59-63) are a
// subsequence of the full index, and Y - X = -u is in the key.  One stable compaction (tile counts, scan, copy) replaces
// extents + pack + radix sort; the -m 3 / -m 4 presets cluster every eps with 2-4 minPts values, each with another cut.
__global__ void __launch_bounds__(256) filter_count_kernel(const u64* __restrict__ keys, GridParams P, int n_in, int cut,
                                                           int* __restrict__ blockcnt) {
    const int i0 = blockIdx.x * EX_ROWS + 4 * threadIdx.x;
    int c = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int i = i0 + k;
        if (i < n_in) c += (-((int)((u32)((keys[i] & KEY_MASK) >> P.be) & P.umask) + P.ubase) >= cut) ? 1 : 0;
    }
    const int tot = __reduce_add_sync(0xffffffffu, c);
    __shared__ int s_w[8];
    if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = tot;
    __syncthreads();
    if (threadIdx.x == 0) {
        int t = 0;
        for (int w = 0; w < 8; ++w) t += s_w[w];
        blockcnt[blockIdx.x] = t;
    }
}

__global__ void __launch_bounds__(256) filter_compact_kernel(const u64* __restrict__ keys_in, const u32* __restrict__ rows_in, GridParams P,
                                                             int n_in, int cut, const int* __restrict__ blockbase,
                                                             u64* __restrict__ keys_out, u32* __restrict__ rows_out) {
    __shared__ int s_warp[8];
    const int i0 = blockIdx.x * EX_ROWS + 4 * threadIdx.x;
    u64 key[4];
    unsigned act = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int i = i0 + k;
        if (i >= n_in) continue;
        key[k] = keys_in[i] & KEY_MASK;                      // a core flag of an earlier clustering of the full index is dropped
        if (-((int)((u32)(key[k] >> P.be) & P.umask) + P.ubase) >= cut) act |= 1u << k;
    }
    const int mine = __popc(act), lane = threadIdx.x & 31;
    int incl = mine;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += v;
    }
    if (lane == 31) s_warp[threadIdx.x >> 5] = incl;
    __syncthreads();
    int pos = blockbase[blockIdx.x] + incl - mine;
    for (int w = 0; w < (int)(threadIdx.x >> 5); ++w) pos += s_warp[w];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        if (!(act & (1u << k))) continue;
        keys_out[pos] = key[k];
        rows_out[pos] = rows_in[i0 + k];
        ++pos;
    }
}

__global__ void filter_total_kernel(const int* __restrict__ blockbase, const int* __restrict__ blockcnt_last, int nblk, int* __restrict__ total) {
    *total = blockbase[nblk - 1] + *blockcnt_last;
}

int index_filter(const cloops_index* base, int32_t cut, cloops_index** out, cudaStream_t st) {
    if (!base) return fail(CLOOPS_EINVAL, "base index is NULL");
    RET_IF(pool_init());
    cloops_index* ix = new cloops_index();
    *out = ix;
    ix->P = base->P;
    GridParams& P = ix->P;
    const int n_in = base->P.n_act;
    if (base->P.n == 0 || n_in == 0) { P.n_act = 0; return 0; }
    if (cut <= 0) cut = INT_MIN;                             // no filter: every row of the full index stays (pipe.py:59 "if cut > 0")
    Temp tmp(st);
    const int nblk = cdiv(n_in, EX_ROWS);
    int *d_cnt, *d_base, *d_total;
    RET_IF(tmp.alloc(&d_cnt, nblk));
    RET_IF(tmp.alloc(&d_base, nblk));
    RET_IF(tmp.alloc(&d_total, 1));
    LAUNCH(filter_count_kernel, nblk, 256, 0, st, base->keys, base->P, n_in, cut, d_cnt);
    CU_TRY(cudaMemcpyAsync(d_base, d_cnt, (size_t)nblk * sizeof(int), cudaMemcpyDeviceToDevice, st));
    LAUNCH(blockcnt_scan_kernel, 1, 1024, 0, st, d_base, nblk);
    LAUNCH(filter_total_kernel, 1, 1, 0, st, d_base, d_cnt + (nblk - 1), nblk, d_total);
    int total = 0;
    CU_TRY(cudaMemcpyAsync(&total, d_total, sizeof(int), cudaMemcpyDeviceToHost, st));
    CU_TRY(cudaStreamSynchronize(st));
    stage_mark("extents", st);
    P.n_act = total;
    if (total == 0) return 0;
    CU_TRY(cudaMallocAsync((void**)&ix->keys, ((size_t)total + 2) * sizeof(u64), st));
    CU_TRY(cudaMallocAsync((void**)&ix->rows, (size_t)total * sizeof(u32), st));
    CU_TRY(cudaMallocAsync((void**)&ix->sstart, (size_t)(P.ns + 3) * sizeof(int), st));
    LAUNCH(filter_compact_kernel, nblk, 256, 0, st, base->keys, base->rows, base->P, n_in, cut, d_base, ix->keys, ix->rows);
    stage_mark("pack", st);
    LAUNCH(strip_table_search_kernel, cdiv(P.ns + 3, 256), 256, 0, st, ix->keys, P, ix->sstart);
    RET_IF(index_tiles(ix, st));
    stage_mark("strips", st);
    return 0;
}

void index_free(cloops_index* ix, cudaStream_t st) {
    if (!ix) return;
    if (ix->keys) cudaFreeAsync(ix->keys, st);
    if (ix->rows) cudaFreeAsync(ix->rows, st);
    if (ix->sstart) cudaFreeAsync(ix->sstart, st);
    if (ix->tiles) cudaFreeAsync(ix->tiles, st);
    delete ix;
}

}  // namespace cloops
