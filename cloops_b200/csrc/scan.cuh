// Prefix scans over int arrays (sum, running max, running min; forward or from the back), written for this library:
// three launches -- per-tile totals, one CTA scanning the totals, per-tile scan with the carried prefix.  They replace
// the cub::DeviceScan calls of round 1 (strip table, rotated-cell heads, chain heads, "next chain head", cluster
// numbering, candidate positions); the stable radix sort is the only CUB algorithm left.
#pragma once
#include <limits.h>

#include "common.cuh"

namespace cloops {

enum : int { SCAN_ADD = 0, SCAN_MAX = 1, SCAN_MIN = 2 };

template <int OP> __device__ __forceinline__ int scan_identity() { return OP == SCAN_ADD ? 0 : (OP == SCAN_MAX ? INT_MIN : INT_MAX); }
template <int OP> __device__ __forceinline__ int scan_apply(int a, int b) { return OP == SCAN_ADD ? a + b : (OP == SCAN_MAX ? max(a, b) : min(a, b)); }

#define SCAN_THREADS 256
#define SCAN_ITEMS 8
#define SCAN_TILE (SCAN_THREADS * SCAN_ITEMS)

// element i of the scan order: forward = i, reverse = n-1-i
template <bool REVERSE> __device__ __forceinline__ long long scan_index(long long i, long long n) { return REVERSE ? n - 1 - i : i; }

template <int OP>
__device__ __forceinline__ int scan_cta_inclusive(int v, int* s_warp, int& cta_total) {     // inclusive scan of one value per thread over the CTA
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int y = __shfl_up_sync(0xffffffffu, v, d);
        if (lane >= d) v = scan_apply<OP>(v, y);
    }
    if (lane == 31) s_warp[warp] = v;
    __syncthreads();
    int before = scan_identity<OP>();
    int total = scan_identity<OP>();
#pragma unroll
    for (int w = 0; w < SCAN_THREADS / 32; ++w) {
        const int t = s_warp[w];
        if (w < warp) before = scan_apply<OP>(before, t);
        total = scan_apply<OP>(total, t);
    }
    __syncthreads();
    cta_total = total;
    return scan_apply<OP>(before, v);
}

template <int OP, bool REVERSE>
__global__ void __launch_bounds__(SCAN_THREADS) scan_totals_kernel(const int* __restrict__ in, long long n, int* __restrict__ totals) {
    __shared__ int s_warp[SCAN_THREADS / 32];
    const long long base = (long long)blockIdx.x * SCAN_TILE;
    int v = scan_identity<OP>();
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) {
        const long long i = base + (long long)k * SCAN_THREADS + threadIdx.x;
        if (i < n) v = scan_apply<OP>(v, __ldg(in + scan_index<REVERSE>(i, n)));
    }
    int total;
    scan_cta_inclusive<OP>(v, s_warp, total);
    if (threadIdx.x == 0) totals[blockIdx.x] = total;
}

// exclusive scan of the tile totals in place (one CTA)
template <int OP>
__global__ void __launch_bounds__(SCAN_THREADS) scan_spine_kernel(int* __restrict__ totals, int m) {
    __shared__ int s_warp[SCAN_THREADS / 32];
    __shared__ int s_incl[SCAN_THREADS];
    int carry = scan_identity<OP>();
    for (int base = 0; base < m; base += SCAN_THREADS) {              // uniform trip count: every thread meets every barrier
        const int i = base + threadIdx.x;
        const int v = i < m ? totals[i] : scan_identity<OP>();
        int total;
        const int incl = scan_cta_inclusive<OP>(v, s_warp, total);
        s_incl[threadIdx.x] = incl;
        __syncthreads();
        const int excl = threadIdx.x ? s_incl[threadIdx.x - 1] : scan_identity<OP>();
        if (i < m) totals[i] = scan_apply<OP>(carry, excl);
        __syncthreads();
        carry = scan_apply<OP>(carry, total);
    }
}

// A thread owns SCAN_ITEMS consecutive items of the scan order (8 ints = two 128-bit loads / stores when the tile is whole and
// the arrays are 16-byte aligned; in reverse order the same 32 bytes, read back to front).
template <int OP, bool INCLUSIVE, bool REVERSE>
__global__ void __launch_bounds__(SCAN_THREADS) scan_tiles_kernel(const int* __restrict__ in, int* __restrict__ out, long long n,
                                                                  const int* __restrict__ prefix, int vec) {
    __shared__ int s_warp[SCAN_THREADS / 32];
    __shared__ int s_incl[SCAN_THREADS];
    const long long base = (long long)blockIdx.x * SCAN_TILE + (long long)threadIdx.x * SCAN_ITEMS;
    const bool whole = vec && base + SCAN_ITEMS <= n && (!REVERSE || ((n & 3) == 0));
    int x[SCAN_ITEMS];
    if (whole) {
        // memory positions of items base .. base+7: forward [base, base+8) ; reverse [n-8-base, n-base), item k at n-1-base-k
        const long long m0 = REVERSE ? n - SCAN_ITEMS - base : base;
        const int4 a = __ldg(reinterpret_cast<const int4*>(in + m0)), b = __ldg(reinterpret_cast<const int4*>(in + m0 + 4));
        if (REVERSE) { x[0] = b.w; x[1] = b.z; x[2] = b.y; x[3] = b.x; x[4] = a.w; x[5] = a.z; x[6] = a.y; x[7] = a.x; }
        else { x[0] = a.x; x[1] = a.y; x[2] = a.z; x[3] = a.w; x[4] = b.x; x[5] = b.y; x[6] = b.z; x[7] = b.w; }
    } else {
#pragma unroll
        for (int k = 0; k < SCAN_ITEMS; ++k) {
            const long long i = base + k;
            x[k] = i < n ? __ldg(in + scan_index<REVERSE>(i, n)) : scan_identity<OP>();
        }
    }
    int mine = scan_identity<OP>();
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) mine = scan_apply<OP>(mine, x[k]);
    int total;
    const int incl = scan_cta_inclusive<OP>(mine, s_warp, total);
    s_incl[threadIdx.x] = incl;
    __syncthreads();
    int run = scan_apply<OP>(prefix[blockIdx.x], threadIdx.x ? s_incl[threadIdx.x - 1] : scan_identity<OP>());
    int y[SCAN_ITEMS];
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) {
        const int after = scan_apply<OP>(run, x[k]);
        y[k] = INCLUSIVE ? after : run;
        run = after;
    }
    if (whole) {
        const long long m0 = REVERSE ? n - SCAN_ITEMS - base : base;
        if (REVERSE) {
            *reinterpret_cast<int4*>(out + m0) = make_int4(y[7], y[6], y[5], y[4]);
            *reinterpret_cast<int4*>(out + m0 + 4) = make_int4(y[3], y[2], y[1], y[0]);
        } else {
            *reinterpret_cast<int4*>(out + m0) = make_int4(y[0], y[1], y[2], y[3]);
            *reinterpret_cast<int4*>(out + m0 + 4) = make_int4(y[4], y[5], y[6], y[7]);
        }
    } else {
#pragma unroll
        for (int k = 0; k < SCAN_ITEMS; ++k) {
            const long long i = base + k;
            if (i < n) out[scan_index<REVERSE>(i, n)] = y[k];
        }
    }
}

// out may alias in.  tmp: int[scan_tmp_ints(n)] device scratch.
static inline size_t scan_tmp_ints(long long n) { return (size_t)((n + SCAN_TILE - 1) / SCAN_TILE) + 1; }

template <int OP, bool INCLUSIVE, bool REVERSE>
static int device_scan(const int* in, int* out, long long n, int* tmp, cudaStream_t st) {
    if (n <= 0) return 0;
    const int m = (int)((n + SCAN_TILE - 1) / SCAN_TILE);
    LAUNCH((scan_totals_kernel<OP, REVERSE>), m, SCAN_THREADS, 0, st, in, n, tmp);
    LAUNCH((scan_spine_kernel<OP>), 1, SCAN_THREADS, 0, st, tmp, m);
    const int vec = ((((uintptr_t)in) | ((uintptr_t)out)) & 15) == 0 ? 1 : 0;
    LAUNCH((scan_tiles_kernel<OP, INCLUSIVE, REVERSE>), m, SCAN_THREADS, 0, st, in, out, n, tmp, vec);
    return 0;
}

}  // namespace cloops
