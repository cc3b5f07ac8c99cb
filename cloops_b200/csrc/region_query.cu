// Region query over the (strip, u') index: neighbour count per point, saturating at cap
// (cDBSCAN.py:186-205; cDBSCAN2.py:304-346).  See index.cu for the layout of the index.
//
// Geometry (cDBSCAN2.py:66-70): (u,v) = (X-Y, X+Y) turns Manhattan d1 <= eps into max(|du|,|dv|) <= eps.  Points are
// sorted by (strip = floor(v/eps), u); the eps-neighbourhood of a point is three contiguous runs (strips s-1, s, s+1
// restricted to u in [u-eps, u+eps]); inside the own strip |dv| <= eps-1 holds by construction, in strip s-1 (s+1) the
// remaining test is vmod_q >= vmod_p (vmod_q <= vmod_p).
#include <limits.h>

#include <type_traits>

#include "index.cuh"

namespace cloops {

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
//  * The counts phase 1 settles (own strip saturates, or cap reached) leave the CTA as coalesced 128-bit stores straight
//    from registers; phase 2 overwrites its points' counts afterwards (same CTA, ordered by the barrier in between).
//  * One 16-byte header per tile (TileInfo, written once per index) replaces the dependent prologue key -> strip ->
//    strip table: the kernel starts with one broadcast load and has two barriers in all.
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
template <bool NEXT, int PROBES>
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
        u32 wl = w3;
        u32 next_a = ja + 16u;
        if (PROBES == 8 && w3 <= thi) {            // caps >= 10 (Hi-C): windows of 5-8 points are common, four more straight-line probes
            const u32 w7 = lds_off<28>(ja);
            if (NEXT) {
                inc_in_window_le(f, lds_off<16>(ja), thi, lds_off<16>(va), vm);
                inc_in_window_le(f, lds_off<20>(ja), thi, lds_off<20>(va), vm);
                inc_in_window_le(f, lds_off<24>(ja), thi, lds_off<24>(va), vm);
                inc_in_window_le(f, w7, thi, lds_off<28>(va), vm);
            } else {
                inc_in_window_ge(f, lds_off<16>(ja), thi, lds_off<16>(va), vm);
                inc_in_window_ge(f, lds_off<20>(ja), thi, lds_off<20>(va), vm);
                inc_in_window_ge(f, lds_off<24>(ja), thi, lds_off<24>(va), vm);
                inc_in_window_ge(f, w7, thi, lds_off<28>(va), vm);
            }
            wl = w7;
            next_a = ja + 32u;
        }
        if (wl <= thi) {                                                   // rare: more points in the window than were probed
            for (u32 a = next_a; f < room && lds_off<0>(a) <= thi; a += 4u) {
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
            c += adjacent_count<false, (CAPT == 0 ? 8 : 4)>(w_a + 4u * (u32)S[srel - 1] - 4u, dv, wp - one - eps, wp - one + eps, vm, nsteps, last_a, cap - c);
            if (c < cap) c += adjacent_count<true, (CAPT == 0 ? 8 : 4)>(w_a + 4u * (u32)S[srel + 1] - 4u, dv, wp + one - eps, wp + one + eps, vm, nsteps, last_a, cap - c);
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


// per-tile headers of an index (called once by index_build)
int index_tiles(cloops_index* ix, cudaStream_t st) {
    const GridParams& P = ix->P;
    const int ntiles = cdiv(P.n_act, CT_TILE);
    CU_TRY(cudaMallocAsync((void**)&ix->tiles, (size_t)ntiles * sizeof(TileInfo), st));
    ix->rmax = CT_RMAX;
    LAUNCH(tile_info_kernel, cdiv(ntiles, 128), 128, 0, st, ix->keys, ix->sstart, P, ntiles, ix->rmax, reinterpret_cast<TileInfo*>(ix->tiles));
    return 0;
}

int index_count(cloops_index* ix, int cap, int* d_counts_sorted, cudaStream_t st) {
    const GridParams& P = ix->P;
    if (P.n_act == 0) return 0;
    if (cap <= 0) cap = INT_MAX;
    return launch_count_tiled<CT_RMAX>(ix, cap, d_counts_sorted, st);
}

}  // namespace cloops
