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
#include <type_traits>
#include <cub/cub.cuh>

#include "index.cuh"

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

// vec: x and y are 16-byte aligned -> four rows per 128-bit load
__global__ void __launch_bounds__(256) extents_kernel(const int* __restrict__ x, const int* __restrict__ y, int n, int cut, int vec,
                                                      Extents* out) {
    int umin = INT_MAX, umax = INT_MIN, vmin = INT_MAX, vmax = INT_MIN, cnt = 0, bad = 0;
    const int stride = gridDim.x * blockDim.x, t = blockIdx.x * blockDim.x + threadIdx.x;
    int done = 0;
    if (vec) {
        const int n4 = n >> 2;
        const int4* __restrict__ x4 = reinterpret_cast<const int4*>(x);
        const int4* __restrict__ y4 = reinterpret_cast<const int4*>(y);
        for (int i = t; i < n4; i += stride) {
            const int4 a = __ldg(x4 + i), b = __ldg(y4 + i);
            extents_point(a.x, b.x, cut, umin, umax, vmin, vmax, cnt, bad);
            extents_point(a.y, b.y, cut, umin, umax, vmin, vmax, cnt, bad);
            extents_point(a.z, b.z, cut, umin, umax, vmin, vmax, cnt, bad);
            extents_point(a.w, b.w, cut, umin, umax, vmin, vmax, cnt, bad);
        }
        done = n4 << 2;
    }
    for (int i = done + t; i < n; i += stride) extents_point(__ldg(x + i), __ldg(y + i), cut, umin, umax, vmin, vmax, cnt, bad);
    if (bad) out->overflow = 1;
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
            if (RANK) aux = (u32)atomicAdd(cnt_tail, 1);
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
// the scatter merge into resident lines instead of each fetching its sector from DRAM first.
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
#pragma unroll 4
                for (int j = a - a0; j < b - a0; ++j) {
                    const u32 uj = U[j];
                    c += uj < my ? 1 : 0;
                    if (uj == my) c += R[j] < row ? 1 : 0;
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

// --------------------------------------------------------------------------------------------------
// Region query: neighbour count per point, saturating at cap (cDBSCAN.py:186-205; cDBSCAN2.py:304-346).
// count_point_global is the per-point walk on global memory (through L1) that tiles with very long strips
// fall back to; the production kernel is count_kernel_tiled below.
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

// --------------------------------------------------------------------------------------------------
// Region query, tiled form (the production kernel).
//
// The index pads u' by eps on both sides (index_build), so with the staged 32-bit word
//     W = ((strip - (sA-1)) << bu) | u'
// "same strip AND |du| <= eps" is ONE unsigned compare against W_p -/+ eps (own strip) or
// W_p -/+ 2^bu -/+ eps (adjacent strips), and W is ascending over the whole staged range, so guard
// words replace every bounds test.
//  * A CTA owns 1024 consecutive sorted points, FOUR per thread.  Shared-memory slots keep the global
//    index modulo 4 (slot = G + j - (r0 & ~3)): staging moves key pairs with one 128-bit load and two
//    64-bit shared stores, a thread's own-strip window [p-4, p+8) (or [p-8, p+12) for caps 6..9) is
//    three (five) aligned 128-bit shared loads, and all probes of phase 1 are register compares with
//    compile-time indices: the in-window predicate is monotone along a sorted strip, so the cap-1
//    nearest points on each side give min(count, cap-1) per side.
//  * Phase 2 runs on a compacted queue of the points their own strip did not saturate.  The lower bound
//    in an adjacent strip is a UNIFORM binary search: a CTA-wide step count (from the longest staged
//    strip), no per-lane bounds -- running past the strip's end is harmless because W keeps ascending --
//    so one step is load / compare / predicated add, without divergence.  Then 4 predicated probes and
//    a tail loop for the rare longer windows.
//  * Counts leave the CTA as coalesced 128-bit stores.
// Tiles whose staged range does not fit (very long strips: dense Hi-C diagonals, where the own strip
// saturates at once) fall back to count_point_global.
#define CT_THREADS 256
#define CT_TILE 1024
#define CT_RMAX 2560       // staged points per tile: 8 CTAs of 27 KB per SM.  (Measured: a 4864-point variant at 4 CTAs/SM
                           // is slower on long Hi-C strips than letting those tiles take the global fallback.)
#define CT_G 8             // left guard words; the right side keeps 12
#define CT_SMAX 1024

// c += (a >= b), c += (a <= b): compare + predicated add (the compiler's select form costs a third instruction)
__device__ __forceinline__ void inc_ge(int& c, u32 a, u32 b) {
    asm("{\n\t.reg .pred p;\n\tsetp.ge.u32 p, %1, %2;\n\t@p add.s32 %0, %0, 1;\n\t}" : "+r"(c) : "r"(a), "r"(b));
}
__device__ __forceinline__ void inc_le(int& c, u32 a, u32 b) {
    asm("{\n\t.reg .pred p;\n\tsetp.le.u32 p, %1, %2;\n\t@p add.s32 %0, %0, 1;\n\t}" : "+r"(c) : "r"(a), "r"(b));
}
// f += (w <= thi && v >= vm)  /  f += (w <= thi && v <= vm): two chained compares + predicated add
__device__ __forceinline__ void inc_in_window_ge(int& f, u32 w, u32 thi, u32 v, u32 vm) {
    asm("{\n\t.reg .pred p;\n\tsetp.le.u32 p, %1, %2;\n\tsetp.ge.and.u32 p, %3, %4, p;\n\t@p add.s32 %0, %0, 1;\n\t}"
        : "+r"(f) : "r"(w), "r"(thi), "r"(v), "r"(vm));
}
__device__ __forceinline__ void inc_in_window_le(int& f, u32 w, u32 thi, u32 v, u32 vm) {
    asm("{\n\t.reg .pred p;\n\tsetp.le.u32 p, %1, %2;\n\tsetp.le.and.u32 p, %3, %4, p;\n\t@p add.s32 %0, %0, 1;\n\t}"
        : "+r"(f) : "r"(w), "r"(thi), "r"(v), "r"(vm));
}
// shared-memory word at byte address a + OFF (OFF is folded into the instruction)
template <int OFF>
__device__ __forceinline__ u32 lds_off(u32 a) {
    u32 v;
    asm("ld.shared.u32 %0, [%1+%2];" : "=r"(v) : "r"(a), "n"(OFF));
    return v;
}

// Per-tile header, written once per index by tile_info_kernel: the region query starts with ONE load
// instead of key -> strip -> strip table, and needs no barrier before staging.
struct TileInfo {
    int sA;      // strip of the tile's first point
    int r0, r1;  // staged range: first point of strip sA-1, end of strip sB+1
    int meta;    // nse | nsteps << 16 ; 0 = the tile does not fit shared memory (global fallback)
};

__global__ void __launch_bounds__(128) tile_info_kernel(const u64* __restrict__ keys, const int* __restrict__ sstart, GridParams P,
                                                        int ntiles, int rmax, TileInfo* __restrict__ tiles) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= ntiles) return;
    const int t0 = b * CT_TILE, t1 = min(t0 + CT_TILE, P.n_act);
    TileInfo ti;
    ti.sA = (int)((keys[t0] & KEY_MASK) >> P.sshift);
    const int sB = (int)((keys[t1 - 1] & KEY_MASK) >> P.sshift);
    const int nse = sB - ti.sA + 4;               // strip-table entries sstart[sA .. sB+3]
    ti.r0 = sstart[ti.sA];
    ti.r1 = sstart[sB + 3];
    ti.meta = 0;
    if (ti.r1 - ti.r0 <= rmax && nse <= CT_SMAX && ((u64)nse << P.bu) <= 0xffffffffull) {
        int maxlen = 0, prev = ti.r0;
        for (int k = 1; k < nse; ++k) {
            const int cur = sstart[ti.sA + k];
            maxlen = max(maxlen, cur - prev);
            prev = cur;
        }
        ti.meta = nse | ((32 - __clz(maxlen)) << 16);      // 2^nsteps > longest staged strip
    }
    tiles[b] = ti;
}

// Address of the first word >= t among the words that follow address pa (pa: the largest word known to be < t,
// i.e. one word before the start of the searched strip).  nsteps uniform halving steps cover 2^nsteps - 1
// words; W ascends over the whole staged range and ends in 0xffffffff guards, so running past the strip is
// harmless.  Steps of more than 8 words are clamped to the first guard word; the last four cannot leave the 12
// guard words (pa stays below the first guard) and are always taken (extra steps never hurt).
__device__ __forceinline__ u32 uniform_lower_bound(u32 pa, u32 t, int nsteps, u32 last_a) {
#pragma unroll 1
    for (u32 sb = 2u << nsteps; sb > 32u; sb >>= 1) {             // CTA-uniform trip count (none if nsteps <= 4)
        const u32 na = min(pa + sb, last_a);
        if (lds_off<0>(na) < t) pa = na;
    }
    if (lds_off<32>(pa) < t) pa += 32u;
    if (lds_off<16>(pa) < t) pa += 16u;
    if (lds_off<8>(pa) < t) pa += 8u;
    if (lds_off<4>(pa) < t) pa += 4u;
    return pa + 4u;
}

// points of an adjacent strip with W in [tlo,thi] that pass the v test; counting stops at `room`.
// sa = address of the word before the strip's first W; dv = byte distance from the W array to the V array.
template <bool NEXT>
__device__ __forceinline__ int adjacent_count(u32 sa, u32 dv, u32 tlo, u32 thi, u32 vm, int nsteps, u32 last_a, int room) {
    const u32 ja = uniform_lower_bound(sa, tlo, nsteps, last_a);
    const u32 w0 = lds_off<0>(ja);
    int f = 0;
    if (w0 <= thi) {                               // most windows are empty: their lanes issue no further loads
        const u32 va = ja + dv;
        const u32 w3 = lds_off<12>(ja);
        if (NEXT) {
            f = lds_off<0>(va) <= vm ? 1 : 0;
            inc_in_window_le(f, lds_off<4>(ja), thi, lds_off<4>(va), vm);
            inc_in_window_le(f, lds_off<8>(ja), thi, lds_off<8>(va), vm);
            inc_in_window_le(f, w3, thi, lds_off<12>(va), vm);
        } else {
            f = lds_off<0>(va) >= vm ? 1 : 0;
            inc_in_window_ge(f, lds_off<4>(ja), thi, lds_off<4>(va), vm);
            inc_in_window_ge(f, lds_off<8>(ja), thi, lds_off<8>(va), vm);
            inc_in_window_ge(f, w3, thi, lds_off<12>(va), vm);
        }
        if (w3 <= thi) {                                                   // rare: more than four points in the window
            for (u32 a = ja + 16u; f < room && lds_off<0>(a) <= thi; a += 4u) {
                const u32 v = lds_off<0>(a + dv);
                f += (NEXT ? v <= vm : v >= vm) ? 1 : 0;
            }
        }
    }
    return f;
}

template <int CAPT, int RMAX>
__global__ void __launch_bounds__(CT_THREADS) count_kernel_tiled(const u64* __restrict__ keys, const int* __restrict__ sstart,
                                                                  const TileInfo* __restrict__ tiles, GridParams P, int cap_rt,
                                                                  int* __restrict__ cnt, int vec_ok) {
    constexpr int NV = CAPT == 0 ? 0 : (CAPT > 5 ? 2 : 1);              // 128-bit words of context on each side
    constexpr int NP = CAPT > 0 ? CAPT - 1 : 0;                          // probes on each side
    typedef typename std::conditional<(CAPT > 0), unsigned short, u32>::type QT;   // queue entry: point | count << 10
    __shared__ __align__(16) u32 Wg[RMAX + CT_G + 12 + 4];
    __shared__ __align__(16) u32 Vg[RMAX + CT_G + 12 + 4];
    __shared__ QT Q1[CT_TILE];
    __shared__ int S[CT_SMAX];
    __shared__ int s_nq1;
    const int cap = CAPT > 0 ? CAPT : cap_rt;
    const int tid = threadIdx.x;
    const int t0 = blockIdx.x * CT_TILE;
    const int t1 = min(t0 + CT_TILE, P.n_act);
    const int4 ti = __ldg(reinterpret_cast<const int4*>(tiles) + blockIdx.x);
    const int sA = ti.x, r0 = ti.y, r1 = ti.z, nse = ti.w & 0xffff, nsteps = ti.w >> 16;
    if (nse == 0) {                               // CTA-uniform: the staged range does not fit
        for (int i = t0 + tid; i < t1; i += CT_THREADS) cnt[i] = count_point_global(keys, sstart, P, cap, i);
        return;
    }
    const int be = P.be, bu = P.bu;
    const u32 eps = (u32)P.eps, one = 1u << bu, emask = P.emask;
    const int sl0 = CT_G - (r0 & ~3);             // slot of global index j = sl0 + j ; slot % 4 == j % 4
    if (tid == 0) s_nq1 = 0;
    {
        const u32 base = (u32)((long long)(sA - 1) << bu);      // strip sA-1 -> relative strip 0 (mod 2^32)
        const ulonglong2* __restrict__ k2 = reinterpret_cast<const ulonglong2*>(keys);
        const int p_hi = r1 >> 1;
        const int j2 = ((r0 + 1) >> 1) + tid;                   // whole key pairs inside [r0, r1)
        auto put = [&](int j, const ulonglong2& kk) {
            // the core flag (bit 63) never reaches the low word of key >> be
            const int sl = sl0 + 2 * j;
            *reinterpret_cast<uint2*>(&Wg[sl]) = make_uint2((u32)(kk.x >> be) - base, (u32)(kk.y >> be) - base);
            *reinterpret_cast<uint2*>(&Vg[sl]) = make_uint2((u32)kk.x & emask, (u32)kk.y & emask);
        };
        ulonglong2 kk[3];                                       // all loads of the common case in flight at once
#pragma unroll
        for (int k = 0; k < 3; ++k)
            if (j2 + k * CT_THREADS < p_hi) kk[k] = __ldg(k2 + j2 + k * CT_THREADS);
#pragma unroll 1
        for (int k = tid; k < nse; k += CT_THREADS) S[k] = __ldg(sstart + sA + k) + sl0;   // strip offsets as slots
#pragma unroll
        for (int k = 0; k < 3; ++k)
            if (j2 + k * CT_THREADS < p_hi) put(j2 + k * CT_THREADS, kk[k]);
#pragma unroll 1
        for (int j = j2 + 3 * CT_THREADS; j < p_hi; j += CT_THREADS) put(j, __ldg(k2 + j));
        if (tid >= 64 && tid < 66) {                            // the unpaired first / last point
            const int j = tid == 64 ? r0 : r1 - 1;
            if (j & 1 ? tid == 64 : tid == 65) {
                const u64 k = __ldg(keys + j);
                Wg[sl0 + j] = (u32)(k >> be) - base;
                Vg[sl0 + j] = (u32)k & emask;
            }
        }
        if (tid >= 128 && tid < 128 + CT_G) { Wg[sl0 + r0 - 1 - (tid - 128)] = 0u; Vg[sl0 + r0 - 1 - (tid - 128)] = 0u; }
        if (tid >= 160 && tid < 160 + 12) { Wg[sl0 + r1 + (tid - 160)] = 0xffffffffu; Vg[sl0 + r1 + (tid - 160)] = 0u; }
    }
    __syncthreads();
    const u32 w_a = (u32)__cvta_generic_to_shared(Wg);          // shared byte addresses
    const u32 dv = (u32)__cvta_generic_to_shared(Vg) - w_a;
    const u32 last_a = w_a + 4u * (u32)(sl0 + r1);              // first right guard word
    // ---- phase 1: own strip, four consecutive points per thread; saturated counts are final
    const int i0 = t0 + 4 * tid;
    {
        unsigned nm = 0;                                                 // bit k: point k is not saturated yet
        int c[4] = {0, 0, 0, 0};
        if (i0 < t1) {
            const int s0 = sl0 + i0;                                     // multiple of 4
            u32 w[4 * (2 * NV + 1)];
#pragma unroll
            for (int v = 0; v < 2 * NV + 1; ++v) {
                const uint4 x = *reinterpret_cast<const uint4*>(&Wg[s0 + 4 * (v - NV)]);
                w[4 * v] = x.x; w[4 * v + 1] = x.y; w[4 * v + 2] = x.z; w[4 * v + 3] = x.w;
            }
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const u32 wp = w[4 * NV + k];
                const u32 lo = wp - eps, hi = wp + eps;
                int cc = 1;
                if (CAPT > 0) {
#pragma unroll
                    for (int q = 1; q <= NP; ++q) {
                        inc_ge(cc, w[4 * NV + k - q], lo);
                        inc_le(cc, w[4 * NV + k + q], hi);
                    }
                } else if (cap > 1 && i0 + k < t1) {                     // caps >= 10 and exact counts: two uniform searches
                    const u32 sa = w_a + 4u * (u32)S[wp >> bu] - 4u;     // the word before the own strip
                    cc = (int)(uniform_lower_bound(sa, hi + 1u, nsteps, last_a) - uniform_lower_bound(sa, lo, nsteps, last_a)) >> 2;
                }
                c[k] = cc;
                nm |= (cc < cap && i0 + k < t1) ? (1u << k) : 0u;
            }
            const int4 r = make_int4(min(c[0], cap), min(c[1], cap), min(c[2], cap), min(c[3], cap));
            if (vec_ok && i0 + 3 < t1) {
                *reinterpret_cast<int4*>(cnt + i0) = r;
            } else {
                cnt[i0] = r.x;
                if (i0 + 1 < t1) cnt[i0 + 1] = r.y;
                if (i0 + 2 < t1) cnt[i0 + 2] = r.z;
                if (i0 + 3 < t1) cnt[i0 + 3] = r.w;
            }
        }
        // ---- queue of the points whose own strip did not saturate them (warp scan of the per-thread counts)
        const int lane = tid & 31;
        const int mine = __popc(nm);
        int incl = mine;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int y = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += y;
        }
        const int tot = __shfl_sync(0xffffffffu, incl, 31);
        if (tot) {                                                       // warp-uniform
            int qb = 0;
            if (lane == 0) qb = atomicAdd(&s_nq1, tot);
            qb = __shfl_sync(0xffffffffu, qb, 0) + incl - mine;
            u32 qa = (u32)__cvta_generic_to_shared(Q1) + (u32)sizeof(QT) * (u32)qb;      // running store address
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const u32 e = (u32)(4 * tid + k) | ((u32)c[k] << 10);
                if (CAPT > 0)
                    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %1, 0;\n\t@p st.shared.u16 [%0], %2;\n\t@p add.u32 %0, %0, 2;\n\t}"
                                 : "+r"(qa) : "r"(nm & (1u << k)), "h"((unsigned short)e) : "memory");
                else
                    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %1, 0;\n\t@p st.shared.u32 [%0], %2;\n\t@p add.u32 %0, %0, 4;\n\t}"
                                 : "+r"(qa) : "r"(nm & (1u << k)), "r"(e) : "memory");
            }
        }
    }
    __syncthreads();
    // ---- phase 2: strips s-1 and s+1 of the queued points; their counts overwrite the phase-1 values
    {
        const int nq1 = s_nq1;
#pragma unroll 1
        for (int q = tid; q < nq1; q += CT_THREADS) {
            const u32 e = Q1[q];
            const int pt = (int)(e & 1023u);
            int c = (int)(e >> 10);
            const u32 pa = w_a + 4u * (u32)(sl0 + t0 + pt);
            const u32 wp = lds_off<0>(pa), vm = lds_off<0>(pa + dv);
            const int srel = (int)(wp >> bu);
            c += adjacent_count<false>(w_a + 4u * (u32)S[srel - 1] - 4u, dv, wp - one - eps, wp - one + eps, vm, nsteps, last_a, cap - c);
            if (c < cap) c += adjacent_count<true>(w_a + 4u * (u32)S[srel + 1] - 4u, dv, wp + one - eps, wp + one + eps, vm, nsteps, last_a, cap - c);
            cnt[t0 + pt] = min(c, cap);
        }
    }
}

template <int RMAX>
static int launch_count_tiled(const cloops_index* ix, int cap, int* out, cudaStream_t st) {
    const GridParams& P = ix->P;
    const int grid = cdiv(P.n_act, CT_TILE);
    const int vec_ok = (((uintptr_t)out) & 15) == 0 ? 1 : 0;
    const TileInfo* tiles = reinterpret_cast<const TileInfo*>(ix->tiles);
    switch (cap) {
#define CT_CASE(C) case C: LAUNCH((count_kernel_tiled<C, RMAX>), grid, CT_THREADS, 0, st, ix->keys, ix->sstart, tiles, P, cap, out, vec_ok); break;
        CT_CASE(2) CT_CASE(3) CT_CASE(4) CT_CASE(5) CT_CASE(6) CT_CASE(7) CT_CASE(8) CT_CASE(9)
#undef CT_CASE
        default: LAUNCH((count_kernel_tiled<0, RMAX>), grid, CT_THREADS, 0, st, ix->keys, ix->sstart, tiles, P, cap, out, vec_ok); break;
    }
    return 0;
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
    const int vec = ((((uintptr_t)d_x) | ((uintptr_t)d_y)) & 15) == 0 ? 1 : 0;
    LAUNCH(extents_kernel, grid, 256, 0, st, d_x, d_y, (int)n, cut, vec, d_ext);
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
    // sort by strip does it in three light passes: pack (+ one atomic per row: strip histogram and arrival rank), an
    // exclusive scan of the histogram (= the strip table), a scatter next to the strip, and a rank inside the strip.
    // The rank pass costs (rows per strip)^2, so its total work is checked first; long strips (Hi-C density) and
    // sparse tables take the radix sort.
    bool counted = false;
    const char* knob = getenv("CLOOPS_INDEX_SORT");                                       // test / measurement knob
    const bool force_radix = knob != nullptr && strcmp(knob, "radix") == 0;
    const bool force_count = knob != nullptr && strcmp(knob, "count") == 0;
    if (!force_radix && (long long)P.ns <= 4LL * P.n_act + 1024) {
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
        size_t scan_bytes = 0;
        CU_TRY(cub::DeviceScan::ExclusiveSum(nullptr, scan_bytes, cnt, cnt, P.ns + 3, st));
        void* d_scan;
        RET_IF(tmp.alloc((char**)&d_scan, scan_bytes));
        CU_TRY(cub::DeviceScan::ExclusiveSum(d_scan, scan_bytes, cnt, cnt, P.ns + 3, st));
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
        LAUNCH(pack_kernel<false>, cdiv(n, 1024), 256, 0, st, d_x, d_y, cut, P, k0, r0, (int*)nullptr, (int*)nullptr);
        stage_mark("pack", st);
        size_t sort_bytes = 0;
        CU_TRY(cub::DeviceRadixSort::SortPairs(nullptr, sort_bytes, k0, k1, r0, r1, (int)n, begin_bit, end_bit, st));
        void* d_sort;
        RET_IF(tmp.alloc((char**)&d_sort, sort_bytes));
        CU_TRY(cub::DeviceRadixSort::SortPairs(d_sort, sort_bytes, k0, k1, r0, r1, (int)n, begin_bit, end_bit, st));
        stage_mark("sort", st);
        LAUNCH(strip_table_search_kernel, cdiv(P.ns + 3, 256), 256, 0, st, ix->keys, P, ix->sstart);
    }
    {
        const int ntiles = cdiv(P.n_act, CT_TILE);
        CU_TRY(cudaMallocAsync((void**)&ix->tiles, (size_t)ntiles * sizeof(TileInfo), st));
        ix->rmax = CT_RMAX;
        LAUNCH(tile_info_kernel, cdiv(ntiles, 128), 128, 0, st, ix->keys, ix->sstart, P, ntiles, ix->rmax, reinterpret_cast<TileInfo*>(ix->tiles));
    }
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

int index_count(cloops_index* ix, int cap, int* d_counts_sorted, cudaStream_t st) {
    const GridParams& P = ix->P;
    if (P.n_act == 0) return 0;
    if (cap <= 0) cap = INT_MAX;
    return launch_count_tiled<CT_RMAX>(ix, cap, d_counts_sorted, st);
}

}  // namespace cloops
