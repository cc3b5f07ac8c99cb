// Permuted-local-background range counting (cLoops/cModel.py:31-143).
//
// Coverage model (cModel.py:45-57): the chromosome's PETs sorted by X and, separately, by Y (down to 256-bp
// buckets, see COV_COARSE), so that "X in [lo,hi]" and "Y in [lo,hi]" are contiguous slices (the reference's
// np.searchsorted, :65-66).
// For a candidate (iva, ivb) the kernel evaluates the reference's set algebra as per-PET bit masks:
//   in(W,p) = X_p in W  or  Y_p in W   (union of source and target hits, :73-78,:118-127)
//   ra = #in(A), rb = #in(B), rab = #{X in A and Y in B}  (:72-80)
//   na_i = #in(A_i), nb_j = #in(B_j), C_ij = #{in(A_i) and in(B_j)}  (:118-143), windows per :83-105.
// Only PETs with X or Y inside the hull of all windows can contribute; each is visited exactly once
// (X-sorted slice first, then the Y-sorted slice minus PETs already seen through X).
#include <limits.h>
#include <stdlib.h>
#include <string.h>

#include <cub/cub.cuh>

#include "common.cuh"

struct cloops_coverage {
    int n = 0;
    int* xs_x = nullptr;   // X-sorted: X
    int* xs_y = nullptr;   //           partner Y
    int* ys_y = nullptr;   // Y-sorted: Y
    int* ys_x = nullptr;   //           partner X
};

namespace cloops {

#define NW 11   // window 0 = the anchor itself, 1..10 = the shifted windows in the reference's order

struct Windows {
    int a0[NW], a1[NW], b0[NW], b1[NW];
    int h0[2], h1[2];    // disjoint hull intervals (merged when the two families overlap)
    int nh;
    int fa0, fa1, fb0, fb1;   // hull of the A family / of the B family (anchor included)
    int2 A[NW], B[NW];        // (lo, hi - lo) per window for the one-subtract range test; empty windows -> (INT_MAX, 0)
};

// c in [lo, lo + width]  <=>  (unsigned)(c - lo) <= width   (coordinates are < 2^30 in magnitude, no wrap)
__device__ __forceinline__ unsigned window_mask(const int2* __restrict__ win, int nwin, int c) {
    unsigned m = 0;
#pragma unroll
    for (int w = 0; w < NW; ++w) {
        if (w < nwin) {
            const int2 t = win[w];
            m |= (((unsigned)(c - t.x) <= (unsigned)t.y) ? 1u : 0u) << w;
        }
    }
    return m;
}

__device__ __forceinline__ long long fdiv2(long long a) { return (a >= 0) ? a / 2 : -((-a + 1) / 2); }  // floor(a/2)

__device__ void make_windows(int iva0, int iva1, int ivb0, int ivb1, int win, Windows& W) {
    // cModel.py:89-104 (py2 integer division)
    long long ca = fdiv2((long long)iva0 + iva1), cb = fdiv2((long long)ivb0 + ivb1);
    long long sa = fdiv2((long long)iva1 - iva0), sb = fdiv2((long long)ivb1 - ivb0);
    long long step = fdiv2(sa + sb);
    W.a0[0] = iva0; W.a1[0] = iva1; W.b0[0] = ivb0; W.b1[0] = ivb1;
    int k = 1;
    for (int i = -win; i <= win; ++i) {
        if (i == 0) continue;
        long long v;
        v = ca + i * step - sa; W.a0[k] = (int)max(0LL, min(v, (long long)INT_MAX));
        v = ca + i * step + sa; W.a1[k] = (int)max(0LL, min(v, (long long)INT_MAX));
        v = cb + i * step - sb; W.b0[k] = (int)max(0LL, min(v, (long long)INT_MAX));
        v = cb + i * step + sb; W.b1[k] = (int)max(0LL, min(v, (long long)INT_MAX));
        ++k;
    }
    const int nw = (win > 0) ? NW : 1;
    int ha0 = INT_MAX, ha1 = INT_MIN, hb0 = INT_MAX, hb1 = INT_MIN;
    for (int w = 0; w < nw; ++w) {
        ha0 = min(ha0, W.a0[w]); ha1 = max(ha1, W.a1[w]);
        hb0 = min(hb0, W.b0[w]); hb1 = max(hb1, W.b1[w]);
    }
    W.fa0 = ha0; W.fa1 = ha1; W.fb0 = hb0; W.fb1 = hb1;
    for (int w = 0; w < NW; ++w) {
        W.A[w] = (w < nw && W.a0[w] <= W.a1[w]) ? make_int2(W.a0[w], W.a1[w] - W.a0[w]) : make_int2(INT_MAX, 0);
        W.B[w] = (w < nw && W.b0[w] <= W.b1[w]) ? make_int2(W.b0[w], W.b1[w] - W.b0[w]) : make_int2(INT_MAX, 0);
    }
    if (ha0 > hb0) { int t = ha0; ha0 = hb0; hb0 = t; t = ha1; ha1 = hb1; hb1 = t; }
    if (hb0 <= ha1) { W.nh = 1; W.h0[0] = ha0; W.h1[0] = max(ha1, hb1); }
    else { W.nh = 2; W.h0[0] = ha0; W.h1[0] = ha1; W.h0[1] = hb0; W.h1[1] = hb1; }
}

// COV_COARSE > 0 would order the sorted views by the coordinate's bits >= COV_COARSE only (one radix pass fewer, slices cut
// at bucket granularity; round 1 used 8).  The ranked kernel below needs exact order -- a window's PETs are then ONE
// contiguous range of each view and most counts are differences of ranks -- so the views are sorted completely.
#define COV_COARSE 0
__device__ __forceinline__ int lower_bound_i(const int* __restrict__ a, int n, int v) {   // first idx whose bucket >= bucket(v)
    int lo = 0, hi = n;
    const int bv = v >> COV_COARSE;
    while (lo < hi) { int mid = (lo + hi) >> 1; if ((__ldg(a + mid) >> COV_COARSE) < bv) lo = mid + 1; else hi = mid; }
    return lo;
}
__device__ __forceinline__ int upper_bound_i(const int* __restrict__ a, int n, int v) {   // first idx whose bucket > bucket(v)
    int lo = 0, hi = n;
    const int bv = v >> COV_COARSE;
    while (lo < hi) { int mid = (lo + hi) >> 1; if ((__ldg(a + mid) >> COV_COARSE) <= bv) lo = mid + 1; else hi = mid; }
    return lo;
}

// ---- ranked form (round 2; the production kernel) -------------------------------------------------------------------
// With exactly sorted views the PETs whose X (or Y) lies in a window are one contiguous range of the X (Y) view, found by
// two binary searches.  Per candidate (one warp):
//   #in(W)  = |X range of W| + |Y range of W| - #{t in X range of W : Y_t in W}                  (cModel.py:72-80,118-127)
//   rab     = #{t in X range of A : Y_t in B}                                                    (cModel.py:79)
//   C_ij    = #{p in in(A_i) : X_p in B_j or Y_p in B_j}, in(A_i) enumerated as the X range of A_i plus the PETs of
//             the Y range of A_i whose X is not in A_i                                           (cModel.py:128-143)
// so only contiguous, coalesced ranges are walked (each PET costs one or two compares unless its partner lies in the other
// family's hull), lanes count in registers and the warp reduces once per window: no shared-memory atomics, no per-PET masks
// over all 22 windows.  The legacy form (below) walked both hull slices and updated up to 123 counters per PET.
#define RC_WARPS 4
#define RQ_NW (2 * NW)

__device__ __forceinline__ int lb_exact(const int* __restrict__ a, int lo, int hi, int v) {   // first idx in [lo,hi) with a[idx] >= v
    while (lo < hi) { const int mid = (lo + hi) >> 1; if (__ldg(a + mid) < v) lo = mid + 1; else hi = mid; }
    return lo;
}
__device__ __forceinline__ int ub_exact(const int* __restrict__ a, int lo, int hi, int v) {   // first idx in [lo,hi) with a[idx] > v
    while (lo < hi) { const int mid = (lo + hi) >> 1; if (__ldg(a + mid) <= v) lo = mid + 1; else hi = mid; }
    return lo;
}

template <int WIN>
__global__ void __launch_bounds__(32 * RC_WARPS) range_count_ranked_kernel(const int* __restrict__ xs_x, const int* __restrict__ xs_y,
                                                                           const int* __restrict__ ys_y, const int* __restrict__ ys_x, int n,
                                                                           const int* __restrict__ cand, long long ncand,
                                                                           const int* __restrict__ d_ncand, int* __restrict__ out) {
    constexpr int NOUT = (WIN > 0) ? 123 : 3;
    constexpr int NWIN = (WIN > 0) ? NW : 1;
    __shared__ Windows Ws[RC_WARPS];
    __shared__ int rng[RC_WARPS][4 * RQ_NW];         // per window (A: 0..10, B: 11..21): X range lo, hi, Y range lo, hi
    __shared__ int res[RC_WARPS][NOUT];
    const int wl = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long m = (long long)blockIdx.x * RC_WARPS + wl;
    if (m >= ncand || (d_ncand && m >= *d_ncand)) return;      // whole warp
    Windows& W = Ws[wl];
    int* R = rng[wl];
    int* O = res[wl];
    if (lane == 0) {
        const int4 c = __ldg(reinterpret_cast<const int4*>(cand) + m);
        make_windows(c.x, c.y, c.z, c.w, WIN, W);
    }
    __syncwarp();
    // 1. ranges: task = (window 0..2*NWIN-1, view).  The searches of a family stay inside the slice of its hull.
    {
        int hs = 0;                                            // lanes 0..3: X lo, X hi, Y lo, Y hi of the hull of all windows
        const int g0 = W.h0[0], g1 = W.h1[W.nh - 1];
        if (lane == 0) hs = lb_exact(xs_x, 0, n, g0);
        else if (lane == 1) hs = ub_exact(xs_x, 0, n, g1);
        else if (lane == 2) hs = lb_exact(ys_y, 0, n, g0);
        else if (lane == 3) hs = ub_exact(ys_y, 0, n, g1);
        const int sx0 = __shfl_sync(0xffffffffu, hs, 0), sx1 = __shfl_sync(0xffffffffu, hs, 1);
        const int sy0 = __shfl_sync(0xffffffffu, hs, 2), sy1 = __shfl_sync(0xffffffffu, hs, 3);
        for (int task = lane; task < 4 * NWIN; task += 32) {
            const int view = task & 1, wi = task >> 1;         // wi: 0..NWIN-1 = A windows, NWIN..2*NWIN-1 = B windows
            const int fam = wi >= NWIN, w = fam ? wi - NWIN : wi;
            const int w0 = fam ? W.b0[w] : W.a0[w], w1 = fam ? W.b1[w] : W.a1[w];
            int lo = 0, hi = 0;
            if (w0 <= w1) {
                const int* __restrict__ a = view ? ys_y : xs_x;
                lo = lb_exact(a, view ? sy0 : sx0, view ? sy1 : sx1, w0);
                hi = ub_exact(a, lo, view ? sy1 : sx1, w1);
            }
            R[4 * (fam * NW + w) + 2 * view] = lo;
            R[4 * (fam * NW + w) + 2 * view + 1] = hi;
        }
    }
    __syncwarp();
    // 2. window sizes: |X range| + |Y range| - PETs of the X range whose partner lies in the same window; rab on the way
    for (int wi = 0; wi < 2 * NWIN; ++wi) {
        const int fam = wi >= NWIN, w = fam ? wi - NWIN : wi;
        const int* r = R + 4 * (fam * NW + w);
        const int lo = r[0], hi = r[1];
        const int2 own = fam ? W.B[w] : W.A[w];
        const int2 b0w = W.B[0];
        int both = 0, rab = 0;
        for (int t = lo + lane; t < hi; t += 32) {
            const int y = __ldg(xs_y + t);
            both += ((unsigned)(y - own.x) <= (unsigned)own.y) ? 1 : 0;
            if (wi == 0) rab += ((unsigned)(y - b0w.x) <= (unsigned)b0w.y) ? 1 : 0;
        }
        both = __reduce_add_sync(0xffffffffu, both);
        if (wi == 0) rab = __reduce_add_sync(0xffffffffu, rab);
        if (lane == 0) {
            const int v = (hi - lo) + (r[3] - r[2]) - both;
            if (w == 0) { O[fam] = v; if (fam == 0) O[2] = rab; }
            else O[(fam ? 13 : 3) + w - 1] = v;
        }
    }
    // 3. joint table C_ij = #{p : in(A_i, p) and in(B_j, p)}: only PETs that touch BOTH families matter.  Walk the X ranges of
    //    the hull(s) (a PET whose X lies outside every hull can only count through its Y, and then Y must lie in a window of
    //    each family: second walk, over the Y range of the intersection of the two family hulls, usually empty).
    if (WIN > 0) {
        int* J = O + 23;
        for (int t = lane; t < 100; t += 32) J[t] = 0;
        __syncwarp();
        const int fa0 = W.fa0, fa1 = W.fa1, fb0 = W.fb0, fb1 = W.fb1;
        const int nh = W.nh;
        const int i0 = max(fa0, fb0), i1 = min(fa1, fb1);                   // intersection of the family hulls
        for (int pass = 0; pass <= nh; ++pass) {
            int lo, hi;
            const bool via_y = pass == nh;
            if (!via_y) {
                lo = lb_exact(xs_x, 0, n, W.h0[pass]);
                hi = ub_exact(xs_x, lo, n, W.h1[pass]);
            } else {
                if (i0 > i1) break;
                lo = lb_exact(ys_y, 0, n, i0);
                hi = ub_exact(ys_y, lo, n, i1);
            }
            const int* __restrict__ px = via_y ? ys_x : xs_x;
            const int* __restrict__ py = via_y ? ys_y : xs_y;
            int t = lo + lane;
            int xn = 0, yn = 0;
            if (t < hi) { xn = __ldg(px + t); yn = __ldg(py + t); }
            for (; t < hi; t += 32) {
                const int x = xn, y = yn;
                if (t + 32 < hi) { xn = __ldg(px + t + 32); yn = __ldg(py + t + 32); }
                bool xa = x >= fa0 && x <= fa1, xb = x >= fb0 && x <= fb1;
                if (via_y) {
                    bool seen = false;                                       // X inside a hull: met in an earlier pass
                    for (int g = 0; g < nh; ++g) seen |= (x >= W.h0[g] && x <= W.h1[g]);
                    if (seen) continue;
                    xa = xb = false;
                }
                const bool ya = y >= fa0 && y <= fa1, yb = y >= fb0 && y <= fb1;
                if (!(xa || ya) || !(xb || yb)) continue;
                unsigned ma = 0, mb = 0;
                if (xa) ma |= window_mask(W.A, NW, x);
                if (ya) ma |= window_mask(W.A, NW, y);
                ma >>= 1;                                                    // the shifted windows only
                if (ma == 0) continue;
                if (xb) mb |= window_mask(W.B, NW, x);
                if (yb) mb |= window_mask(W.B, NW, y);
                mb >>= 1;
                for (unsigned r = ma; r; r &= r - 1) {
                    const int i = __ffs(r) - 1;
                    for (unsigned q = mb; q; q &= q - 1) atomicAdd(&J[10 * i + (__ffs(q) - 1)], 1);
                }
            }
        }
    }
    __syncwarp();
    for (int t = lane; t < NOUT; t += 32) out[m * NOUT + t] = O[t];
}

// ---- the same counts with one CTA per candidate (WIN = 5) -------------------------------------------------------------
// One warp per candidate leaves the launch waiting for its largest candidates (a loop whose anchors hold 10^5 PETs keeps a
// single warp busy for milliseconds while the SMs drain).  Here RT_THREADS threads share a candidate: the rank searches are
// spread over the threads, every range is walked with a CTA-wide stride, and the counters are shared-memory atomics fed by
// warp-reduced partial sums.  Same arithmetic as range_count_ranked_kernel.
template <int WIN, int RT_THREADS>
__global__ void __launch_bounds__(RT_THREADS) range_count_team_kernel(const int* __restrict__ xs_x, const int* __restrict__ xs_y,
                                                                      const int* __restrict__ ys_y, const int* __restrict__ ys_x, int n,
                                                                      const int* __restrict__ cand, long long ncand,
                                                                      const int* __restrict__ d_ncand, int* __restrict__ out) {
    constexpr int NOUT = (WIN > 0) ? 123 : 3;
    constexpr int NWIN = (WIN > 0) ? NW : 1;
    __shared__ Windows W;
    __shared__ int R[4 * RQ_NW];
    __shared__ int both[RQ_NW];
    __shared__ int O[NOUT];
    __shared__ int hs[4 + 4];
    const int tid = threadIdx.x, lane = tid & 31;
    const long long m = blockIdx.x;
    if (m >= ncand || (d_ncand && m >= *d_ncand)) return;      // whole CTA
    if (tid == 0) {
        const int4 c = __ldg(reinterpret_cast<const int4*>(cand) + m);
        make_windows(c.x, c.y, c.z, c.w, WIN, W);
    }
    for (int t = tid; t < NOUT; t += RT_THREADS) O[t] = 0;
    if (tid < RQ_NW) both[tid] = 0;
    __syncthreads();
    const int nh = W.nh;
    if (tid < 4) {                                             // X lo, X hi, Y lo, Y hi of the span of all windows
        const int g0 = W.h0[0], g1 = W.h1[nh - 1];
        hs[tid] = tid == 0 ? lb_exact(xs_x, 0, n, g0) : (tid == 1 ? ub_exact(xs_x, 0, n, g1) : (tid == 2 ? lb_exact(ys_y, 0, n, g0) : ub_exact(ys_y, 0, n, g1)));
    } else if (WIN > 0 && tid >= 32 && tid < 32 + 2 * nh) {    // X ranges of the hull(s) for the joint walk
        const int h = (tid - 32) >> 1;
        hs[4 + (tid - 32)] = ((tid - 32) & 1) ? ub_exact(xs_x, 0, n, W.h1[h]) : lb_exact(xs_x, 0, n, W.h0[h]);
    }
    __syncthreads();
    if (tid < 4 * NWIN) {
        const int view = tid & 1, wi = tid >> 1;
        const int fam = wi >= NWIN, w = fam ? wi - NWIN : wi;
        const int w0 = fam ? W.b0[w] : W.a0[w], w1 = fam ? W.b1[w] : W.a1[w];
        int lo = 0, hi = 0;
        if (w0 <= w1) {
            const int* __restrict__ a = view ? ys_y : xs_x;
            lo = lb_exact(a, view ? hs[2] : hs[0], view ? hs[3] : hs[1], w0);
            hi = ub_exact(a, lo, view ? hs[3] : hs[1], w1);
        }
        R[4 * (fam * NW + w) + 2 * view] = lo;
        R[4 * (fam * NW + w) + 2 * view + 1] = hi;
    }
    __syncthreads();
    // window sizes: PETs of the X range whose partner lies in the same window; rab on the way
    for (int wi = 0; wi < 2 * NWIN; ++wi) {
        const int fam = wi >= NWIN, w = fam ? wi - NWIN : wi;
        const int* r = R + 4 * (fam * NW + w);
        const int lo = r[0], hi = r[1];
        const int2 own = fam ? W.B[w] : W.A[w];
        const int2 b0w = W.B[0];
        int cb = 0, rab = 0;
        for (int t = lo + tid; t < hi; t += RT_THREADS) {
            const int y = __ldg(xs_y + t);
            cb += ((unsigned)(y - own.x) <= (unsigned)own.y) ? 1 : 0;
            if (wi == 0) rab += ((unsigned)(y - b0w.x) <= (unsigned)b0w.y) ? 1 : 0;
        }
        cb = __reduce_add_sync(0xffffffffu, cb);
        if (lane == 0 && cb) atomicAdd(&both[fam * NW + w], cb);
        if (wi == 0) {
            rab = __reduce_add_sync(0xffffffffu, rab);
            if (lane == 0 && rab) atomicAdd(&O[2], rab);
        }
    }
    // joint table (as range_count_ranked_kernel, CTA-wide stride)
    if (WIN > 0) {
        int* J = O + 23;
        const int fa0 = W.fa0, fa1 = W.fa1, fb0 = W.fb0, fb1 = W.fb1;
        const int i0 = max(fa0, fb0), i1 = min(fa1, fb1);
        for (int pass = 0; pass <= nh; ++pass) {
            int lo, hi;
            const bool via_y = pass == nh;
            if (!via_y) {
                lo = hs[4 + 2 * pass];
                hi = hs[4 + 2 * pass + 1];
            } else {
                if (i0 > i1) break;
                lo = lb_exact(ys_y, hs[2], hs[3], i0);
                hi = ub_exact(ys_y, lo, hs[3], i1);
            }
            const int* __restrict__ px = via_y ? ys_x : xs_x;
            const int* __restrict__ py = via_y ? ys_y : xs_y;
            for (int t = lo + tid; t < hi; t += RT_THREADS) {
                const int x = __ldg(px + t), y = __ldg(py + t);
                bool xa = x >= fa0 && x <= fa1, xb = x >= fb0 && x <= fb1;
                if (via_y) {
                    bool seen = false;
                    for (int g = 0; g < nh; ++g) seen |= (x >= W.h0[g] && x <= W.h1[g]);
                    if (seen) continue;
                    xa = xb = false;
                }
                const bool ya = y >= fa0 && y <= fa1, yb = y >= fb0 && y <= fb1;
                if (!(xa || ya) || !(xb || yb)) continue;
                unsigned ma = 0, mb = 0;
                if (xa) ma |= window_mask(W.A, NW, x);
                if (ya) ma |= window_mask(W.A, NW, y);
                ma >>= 1;
                if (ma == 0) continue;
                if (xb) mb |= window_mask(W.B, NW, x);
                if (yb) mb |= window_mask(W.B, NW, y);
                mb >>= 1;
                for (unsigned rr = ma; rr; rr &= rr - 1) {
                    const int i = __ffs(rr) - 1;
                    for (unsigned q = mb; q; q &= q - 1) atomicAdd(&J[10 * i + (__ffs(q) - 1)], 1);
                }
            }
        }
    }
    __syncthreads();
    for (int t = tid; t < NOUT; t += RT_THREADS) {
        int v = O[t];
        if (t < 2) v = (R[4 * (t * NW) + 1] - R[4 * (t * NW)]) + (R[4 * (t * NW) + 3] - R[4 * (t * NW) + 2]) - both[t * NW];
        else if (t >= 3 && t < 23) {
            const int fam = t >= 13, w = fam ? t - 12 : t - 2;
            const int* r = R + 4 * (fam * NW + w);
            v = (r[1] - r[0]) + (r[3] - r[2]) - both[fam * NW + w];
        }
        out[m * NOUT + t] = v;
    }
}

// ---- legacy form (round 1; CLOOPS_RC=legacy for A/B runs) ------------------------------------------------------------
// One WARP per candidate (4 candidates per CTA), no CTA-wide barriers: lane 0 derives the windows, up to
// 8 lanes run the slice binary searches, then the 32 lanes stride over the X-sorted and Y-sorted slices.
// out row: ra, rb, rab, na[10], nb[10], C[10][10]
template <int WIN>
__global__ void __launch_bounds__(32 * RC_WARPS) range_count_kernel(const int* __restrict__ xs_x, const int* __restrict__ xs_y,
                                                                    const int* __restrict__ ys_y, const int* __restrict__ ys_x, int n,
                                                                    const int* __restrict__ cand, long long ncand, const int* __restrict__ d_ncand,
                                                                    int* __restrict__ out) {
    constexpr int NOUT = (WIN > 0) ? 123 : 3;
    constexpr int NWIN = (WIN > 0) ? NW : 1;
    __shared__ Windows Ws[RC_WARPS];
    __shared__ int accs[RC_WARPS][NOUT];
    const int wl = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long m = (long long)blockIdx.x * RC_WARPS + wl;
    if (m >= ncand || (d_ncand && m >= *d_ncand)) return;      // whole warp
    Windows& W = Ws[wl];
    int* acc = accs[wl];
    for (int t = lane; t < NOUT; t += 32) acc[t] = 0;
    if (lane == 0) {
        const int4 c = __ldg(reinterpret_cast<const int4*>(cand) + m);
        make_windows(c.x, c.y, c.z, c.w, WIN, W);
    }
    __syncwarp();
    const int nh = W.nh;
    int myseg = 0;
    if (lane < 4 * nh) {
        const int h = lane >> 2, which = lane & 3;
        if (which == 0) myseg = lower_bound_i(xs_x, n, W.h0[h]);
        else if (which == 1) myseg = upper_bound_i(xs_x, n, W.h1[h]);
        else if (which == 2) myseg = lower_bound_i(ys_y, n, W.h0[h]);
        else myseg = upper_bound_i(ys_y, n, W.h1[h]);
    }
    for (int pass = 0; pass < 2 * nh; ++pass) {
        const int h = pass >> 1, via_y = pass & 1;
        const int lo = __shfl_sync(0xffffffffu, myseg, 4 * h + 2 * via_y);
        const int hi = __shfl_sync(0xffffffffu, myseg, 4 * h + 2 * via_y + 1);
        const int* __restrict__ pa = via_y ? ys_x : xs_x;      // x coordinate source
        const int* __restrict__ pb = via_y ? ys_y : xs_y;      // y coordinate source
        int t = lo + lane;
        int xn = 0, yn = 0;
        if (t < hi) { xn = __ldg(pa + t); yn = __ldg(pb + t); }
        for (; t < hi; t += 32) {
            const int x = xn, y = yn;
            if (t + 32 < hi) { xn = __ldg(pa + t + 32); yn = __ldg(pb + t + 32); }   // prefetch the next PET
            if (via_y) {
                bool seen = false;                             // already visited through its X
                for (int g = 0; g < nh; ++g) seen |= (x >= W.h0[g] && x <= W.h1[g]);
                if (seen || y < W.h0[h] || y > W.h1[h]) continue;   // (the slice is cut at bucket granularity)
            } else if (x < W.h0[h] || x > W.h1[h]) continue;
            // a coordinate can only fall into a window of a family whose hull contains it: the X-sorted
            // slice is inside one hull by construction, and the partner coordinate is usually far away
            unsigned ma = 0, mb = 0;
            const bool xa = x >= W.fa0 && x <= W.fa1, ya = y >= W.fa0 && y <= W.fa1;
            const bool xb = x >= W.fb0 && x <= W.fb1, yb = y >= W.fb0 && y <= W.fb1;
            if (xa) ma |= window_mask(W.A, NWIN, x);
            if (ya) ma |= window_mask(W.A, NWIN, y);
            if (xb) mb |= window_mask(W.B, NWIN, x);
            if (yb) mb |= window_mask(W.B, NWIN, y);
            if ((ma | mb) == 0) continue;
            if (ma & 1u) atomicAdd(&acc[0], 1);
            if (mb & 1u) atomicAdd(&acc[1], 1);
            if (x >= W.a0[0] && x <= W.a1[0] && y >= W.b0[0] && y <= W.b1[0]) atomicAdd(&acc[2], 1);
            if (WIN > 0) {
                unsigned wa = ma >> 1, wb = mb >> 1;
                for (unsigned r = wa; r; r &= r - 1) atomicAdd(&acc[3 + (__ffs(r) - 1)], 1);
                for (unsigned r = wb; r; r &= r - 1) atomicAdd(&acc[13 + (__ffs(r) - 1)], 1);
                for (unsigned r = wa; r; r &= r - 1) {
                    int i = __ffs(r) - 1;
                    for (unsigned q = wb; q; q &= q - 1) atomicAdd(&acc[23 + 10 * i + (__ffs(q) - 1)], 1);
                }
            }
        }
    }
    __syncwarp();
    for (int t = lane; t < NOUT; t += 32) out[m * NOUT + t] = acc[t];
}

// Algorithmic work of range_count_kernel for a candidate list: PETs with X inside the hull of all windows and
// PETs with Y inside it, summed over the candidates -- out[0], out[1].  Exact slices (not the 256-bp bucket slices the
// count kernel cuts): per hull interval, #{c : lo <= c <= hi} = #{bucket < b(hi)} - #{bucket < b(lo)} corrected by an exact
// scan of the two boundary buckets.
__device__ __forceinline__ int exact_rank(const int* __restrict__ a, int n, int v, bool upper) {
    // number of entries < v (upper = false) or <= v (upper = true) in an array sorted by bucket only
    int lo = lower_bound_i(a, n, v), hi = upper_bound_i(a, n, v);
    int r = lo;
    for (int j = lo; j < hi; ++j) r += upper ? (__ldg(a + j) <= v) : (__ldg(a + j) < v);
    return r;
}
__global__ void __launch_bounds__(128) range_work_kernel(const int* __restrict__ xs_x, const int* __restrict__ ys_y, int n,
                                                         const int* __restrict__ cand, long long ncand, int win,
                                                         unsigned long long* __restrict__ out) {
    const long long m = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long vx = 0, vy = 0;
    if (m < ncand) {
        Windows W;
        const int4 c = __ldg(reinterpret_cast<const int4*>(cand) + m);
        make_windows(c.x, c.y, c.z, c.w, win, W);
        for (int h = 0; h < W.nh; ++h) {
            vx += (unsigned long long)max(0, exact_rank(xs_x, n, W.h1[h], true) - exact_rank(xs_x, n, W.h0[h], false));
            vy += (unsigned long long)max(0, exact_rank(ys_y, n, W.h1[h], true) - exact_rank(ys_y, n, W.h0[h], false));
        }
    }
    for (int d = 16; d > 0; d >>= 1) { vx += __shfl_down_sync(0xffffffffu, vx, d); vy += __shfl_down_sync(0xffffffffu, vy, d); }
    if ((threadIdx.x & 31) == 0 && (vx | vy)) { atomicAdd(out, vx); atomicAdd(out + 1, vy); }
}

}  // namespace cloops

using namespace cloops;

namespace cloops {
int coverage_build(const int32_t* d_x, const int32_t* d_y, int64_t n, cloops_coverage** out, cudaStream_t st) {
    if (n < 0 || n > 0x7fffff00LL) return fail(CLOOPS_EINVAL, "n=%lld out of range", (long long)n);
    RET_IF(pool_init());
    cloops_coverage* cov = new cloops_coverage();
    *out = cov;
    cov->n = (int)n;
    if (n == 0) return 0;
    CU_TRY(cudaMallocAsync((void**)&cov->xs_x, n * sizeof(int), st));
    CU_TRY(cudaMallocAsync((void**)&cov->xs_y, n * sizeof(int), st));
    CU_TRY(cudaMallocAsync((void**)&cov->ys_y, n * sizeof(int), st));
    CU_TRY(cudaMallocAsync((void**)&cov->ys_x, n * sizeof(int), st));
    Temp tmp(st);
    size_t bytes = 0;
    CU_TRY(cub::DeviceRadixSort::SortPairs(nullptr, bytes, d_x, cov->xs_x, d_y, cov->xs_y, (int)n, COV_COARSE, 32, st));
    void* d_tmp;
    RET_IF(tmp.alloc((char**)&d_tmp, bytes));
    CU_TRY(cub::DeviceRadixSort::SortPairs(d_tmp, bytes, d_x, cov->xs_x, d_y, cov->xs_y, (int)n, COV_COARSE, 32, st));
    CU_TRY(cub::DeviceRadixSort::SortPairs(d_tmp, bytes, d_y, cov->ys_y, d_x, cov->ys_x, (int)n, COV_COARSE, 32, st));
    return 0;
}

// d_ncand != NULL: the candidate count is read on the device (no host round trip); ncand is then an upper bound
// CLOOPS_RC=legacy selects the round-1 walk for A/B runs.  Measured on the config-4 step (437 k scored candidates of 1.23 M,
// anchors ~19 kb wide, mostly overlapping A / B hulls): 123 integers 260 -> 148 ms per step, (ra, rb, rab) of all candidates
// 60.5 -> 8.9 ms per step (profiles/r02_bench_c4_ranked.json / _legacy.json).
static bool rc_team() {                       // CLOOPS_RC=warp: one warp per candidate also for the 123 integers (A/B)
    static const bool v = !(getenv("CLOOPS_RC") != nullptr && strcmp(getenv("CLOOPS_RC"), "warp") == 0);
    return v;
}
static int launch_team(const cloops_coverage* cov, const int32_t* d_cand, int64_t ncand, const int* d_ncand, int32_t* d_out, cudaStream_t st) {
    static const int threads = getenv("CLOOPS_RC_TEAM") ? atoi(getenv("CLOOPS_RC_TEAM")) : 256;      // measurement knob: 128 / 256 / 512
#define RT_LAUNCH(T) LAUNCH((range_count_team_kernel<5, T>), (unsigned)ncand, T, 0, st, cov->xs_x, cov->xs_y, cov->ys_y, cov->ys_x, cov->n, d_cand, (long long)ncand, d_ncand, d_out)
    if (threads == 128) RT_LAUNCH(128);
    else if (threads == 512) RT_LAUNCH(512);
    else RT_LAUNCH(256);
#undef RT_LAUNCH
    return 0;
}
static bool rc_legacy(int win = 5) {
    static const bool v = getenv("CLOOPS_RC") != nullptr && strcmp(getenv("CLOOPS_RC"), "legacy") == 0;
    (void)win;
    return v;
}

int range_counts_dev(const cloops_coverage* cov, const int32_t* d_cand, int64_t ncand, const int* d_ncand, int32_t* d_out,
                     cudaStream_t st) {
    if (ncand <= 0) return 0;
    if (cov->n == 0) { CU_TRY(cudaMemsetAsync(d_out, 0, (size_t)ncand * 123 * sizeof(int), st)); return 0; }
    if (!rc_legacy()) {
        if (rc_team()) RET_IF(launch_team(cov, d_cand, ncand, d_ncand, d_out, st));
        else LAUNCH(range_count_ranked_kernel<5>, (unsigned)((ncand + RC_WARPS - 1) / RC_WARPS), 32 * RC_WARPS, 0, st, cov->xs_x, cov->xs_y,
                    cov->ys_y, cov->ys_x, cov->n, d_cand, (long long)ncand, d_ncand, d_out);
        return 0;
    }
    LAUNCH(range_count_kernel<5>, (unsigned)((ncand + RC_WARPS - 1) / RC_WARPS), 32 * RC_WARPS, 0, st, cov->xs_x, cov->xs_y, cov->ys_y,
           cov->ys_x, cov->n, d_cand, (long long)ncand, d_ncand, d_out);
    return 0;
}
}  // namespace cloops

extern "C" {

int cloops_coverage_build(const int32_t* d_x, const int32_t* d_y, int64_t n, cloops_coverage** out, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (!out) return fail(CLOOPS_EINVAL, "out is NULL");
    *out = nullptr;
    stages_begin(st);
    int rc = coverage_build(d_x, d_y, n, out, st);
    if (rc != 0) {
        cloops_coverage_free(*out);
        *out = nullptr;
        return rc;
    }
    stage_mark("coverage_sort", st);
    return stages_end(st);
}

void cloops_coverage_release(cloops_coverage* cov, void* stream) {
    if (!cov) return;
    cudaStream_t st = (cudaStream_t)stream;
    if (cov->xs_x) cudaFreeAsync(cov->xs_x, st);
    if (cov->xs_y) cudaFreeAsync(cov->xs_y, st);
    if (cov->ys_y) cudaFreeAsync(cov->ys_y, st);
    if (cov->ys_x) cudaFreeAsync(cov->ys_x, st);
    delete cov;
}

void cloops_coverage_free(cloops_coverage* cov) { cloops_coverage_release(cov, 0); }

int cloops_range_counts(const cloops_coverage* cov, const int32_t* d_cand, int64_t m, int32_t* d_out, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (!cov) return fail(CLOOPS_EINVAL, "coverage is NULL");
    if (m < 0 || m > 0x7fffffffLL) return fail(CLOOPS_EINVAL, "m out of range");
    stages_begin(st);
    if (m > 0) {
        if (cov->n == 0) CU_TRY(cudaMemsetAsync(d_out, 0, (size_t)m * 123 * sizeof(int), st));
        else if (!rc_legacy() && rc_team()) RET_IF(launch_team(cov, d_cand, m, (const int*)nullptr, d_out, st));
        else if (!rc_legacy()) LAUNCH(range_count_ranked_kernel<5>, (unsigned)((m + RC_WARPS - 1) / RC_WARPS), 32 * RC_WARPS, 0, st, cov->xs_x, cov->xs_y, cov->ys_y, cov->ys_x, cov->n, d_cand, (long long)m, (const int*)nullptr, d_out);
        else LAUNCH(range_count_kernel<5>, (unsigned)((m + RC_WARPS - 1) / RC_WARPS), 32 * RC_WARPS, 0, st, cov->xs_x, cov->xs_y, cov->ys_y, cov->ys_x, cov->n, d_cand, (long long)m, (const int*)nullptr, d_out);
    }
    stage_mark("range_counts", st);
    return stages_end(st);
}

/* Measurement aid (bench.py roofline of the range-count kernel): h_work[0] = sum over candidates of PETs with X
 * inside the hull of the candidate's windows, h_work[1] = same for Y; win = 5 (range_counts) or 0 (region_pets).  Synchronises. */
int cloops_range_work(const cloops_coverage* cov, const int32_t* d_cand, int64_t m, int32_t win, uint64_t* h_work, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (!cov || !h_work) return fail(CLOOPS_EINVAL, "NULL argument");
    h_work[0] = h_work[1] = 0;
    if (m <= 0 || cov->n == 0) return 0;
    Temp tmp(st);
    unsigned long long* d_w;
    RET_IF(tmp.alloc(&d_w, 2));
    CU_TRY(cudaMemsetAsync(d_w, 0, 2 * sizeof(unsigned long long), st));
    LAUNCH(range_work_kernel, (unsigned)((m + 127) / 128), 128, 0, st, cov->xs_x, cov->ys_y, cov->n, d_cand, (long long)m, (int)win, d_w);
    unsigned long long w[2];
    CU_TRY(cudaMemcpyAsync(w, d_w, sizeof(w), cudaMemcpyDeviceToHost, st));
    CU_TRY(cudaStreamSynchronize(st));
    h_work[0] = w[0];
    h_work[1] = w[1];
    return 0;
}

int cloops_region_pets(const cloops_coverage* cov, const int32_t* d_cand, int64_t m, int32_t* d_out, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (!cov) return fail(CLOOPS_EINVAL, "coverage is NULL");
    if (m < 0 || m > 0x7fffffffLL) return fail(CLOOPS_EINVAL, "m out of range");
    stages_begin(st);
    if (m > 0) {
        if (cov->n == 0) CU_TRY(cudaMemsetAsync(d_out, 0, (size_t)m * 3 * sizeof(int), st));
        else if (!rc_legacy(0)) LAUNCH(range_count_ranked_kernel<0>, (unsigned)((m + RC_WARPS - 1) / RC_WARPS), 32 * RC_WARPS, 0, st, cov->xs_x, cov->xs_y, cov->ys_y, cov->ys_x, cov->n, d_cand, (long long)m, (const int*)nullptr, d_out);
        else LAUNCH(range_count_kernel<0>, (unsigned)((m + RC_WARPS - 1) / RC_WARPS), 32 * RC_WARPS, 0, st, cov->xs_x, cov->xs_y, cov->ys_y, cov->ys_x, cov->n, d_cand, (long long)m, (const int*)nullptr, d_out);
    }
    stage_mark("region_pets", st);
    return stages_end(st);
}

}  // extern "C"
