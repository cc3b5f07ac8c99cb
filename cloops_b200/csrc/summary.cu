// Cluster -> candidate record reduction (cLoops/pipe.py:76-109): per-cluster bounding box and size,
// zero-extent drop (:83-85), inter/self split (:97) and per-row membership of dis / dss (:106-109).
#include <limits.h>

#include "common.cuh"

namespace cloops {

__global__ void __launch_bounds__(256) summary_init_kernel(int* __restrict__ bbox, int* __restrict__ size, long long k) {
    long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= k) return;
    reinterpret_cast<int4*>(bbox)[c] = make_int4(INT_MAX, INT_MIN, INT_MAX, INT_MIN);
    size[c] = 0;
}

// Rows arrive in file order, so labels inside a CTA are unrelated in general, and one giant cluster
// (the Hi-C / self-ligation diagonal) can own half of all rows.  Lanes of a warp that carry the same
// label first reduce in registers; the group leaders then aggregate in a CTA-wide open-addressing hash
// table in shared memory (1024 slots for the at most 1024 rows a CTA visits; a probe sequence longer
// than 64 falls back to global atomics); only one partial per (CTA, label) reaches the global atomics.
#define SUM_SLOTS 1024
#define SUM_TILES 4
__global__ void __launch_bounds__(256) summary_accumulate_kernel(const int* __restrict__ x, const int* __restrict__ y,
                                                                 const int* __restrict__ labels, long long n, long long k,
                                                                 int* __restrict__ bbox, int* __restrict__ size) {
    __shared__ int s_lab[SUM_SLOTS], s_x0[SUM_SLOTS], s_x1[SUM_SLOTS], s_y0[SUM_SLOTS], s_y1[SUM_SLOTS], s_cnt[SUM_SLOTS];
    for (int t = threadIdx.x; t < SUM_SLOTS; t += blockDim.x) {
        s_lab[t] = -1; s_cnt[t] = 0;
        s_x0[t] = INT_MAX; s_x1[t] = INT_MIN; s_y0[t] = INT_MAX; s_y1[t] = INT_MIN;
    }
    __syncthreads();
    for (int tile = 0; tile < SUM_TILES; ++tile) {
        long long i = ((long long)blockIdx.x * SUM_TILES + tile) * blockDim.x + threadIdx.x;
        int lab = (i < n) ? __ldg(labels + i) : -1;
        if (lab >= k) lab = -1;
        const unsigned lm = __ballot_sync(0xffffffffu, lab >= 0);
        if (lab < 0) continue;
        int xx = __ldg(x + i), yy = __ldg(y + i);
        const unsigned m = __match_any_sync(lm, lab);
        const int x0 = __reduce_min_sync(m, xx), x1 = __reduce_max_sync(m, xx);
        const int y0 = __reduce_min_sync(m, yy), y1 = __reduce_max_sync(m, yy);
        if ((int)(threadIdx.x & 31) != __ffs(m) - 1) continue;
        int slot = (int)(((unsigned)lab * 2654435761u) >> 22);            // 10 bits
        bool found = false;
        for (int probe = 0; probe < 64; ++probe) {
            const int prev = atomicCAS(&s_lab[slot], -1, lab);
            if (prev == -1 || prev == lab) { found = true; break; }
            slot = (slot + 1) & (SUM_SLOTS - 1);
        }
        if (found) {
            atomicMin(&s_x0[slot], x0); atomicMax(&s_x1[slot], x1);
            atomicMin(&s_y0[slot], y0); atomicMax(&s_y1[slot], y1);
            atomicAdd(&s_cnt[slot], __popc(m));
        } else {
            int* b = bbox + 4LL * lab;
            atomicMin(b + 0, x0); atomicMax(b + 1, x1); atomicMin(b + 2, y0); atomicMax(b + 3, y1);
            atomicAdd(size + lab, __popc(m));
        }
    }
    __syncthreads();
    for (int t = threadIdx.x; t < SUM_SLOTS; t += blockDim.x) {
        const int l = s_lab[t];
        if (l < 0) continue;
        int* b = bbox + 4LL * l;
        atomicMin(b + 0, s_x0[t]); atomicMax(b + 1, s_x1[t]); atomicMin(b + 2, s_y0[t]); atomicMax(b + 3, s_y1[t]);
        atomicAdd(size + l, s_cnt[t]);
    }
}

__global__ void __launch_bounds__(256) summary_kind_kernel(const int* __restrict__ bbox, const int* __restrict__ size, long long k,
                                                           unsigned char* __restrict__ kind) {
    long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= k) return;
    unsigned char r = 0;
    if (size[c] > 0) {
        int4 b = reinterpret_cast<const int4*>(bbox)[c];
        if (b.x != b.y && b.z != b.w) r = (b.y < b.z) ? 1 : 2;
    }
    kind[c] = r;
}

__global__ void __launch_bounds__(256) summary_rowkind_kernel(const int* __restrict__ labels, const unsigned char* __restrict__ kind,
                                                              long long n, long long k, unsigned char* __restrict__ row_kind) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int lab = __ldg(labels + i);
    row_kind[i] = (lab >= 0 && lab < k) ? kind[lab] : 0;
}

int cluster_summary(const int32_t* d_x, const int32_t* d_y, const int32_t* d_labels, int64_t n, int64_t k, int32_t* d_bbox,
                    int32_t* d_size, uint8_t* d_kind, uint8_t* d_row_kind, cudaStream_t st) {
    if (n < 0 || k < 0) return fail(CLOOPS_EINVAL, "negative size");
    if (k > 0) LAUNCH(summary_init_kernel, cdiv(k, 256), 256, 0, st, d_bbox, d_size, (long long)k);
    if (n > 0 && k > 0)
        LAUNCH(summary_accumulate_kernel, cdiv(n, 256 * SUM_TILES), 256, 0, st, d_x, d_y, d_labels, (long long)n, (long long)k, d_bbox, d_size);
    if (k > 0) LAUNCH(summary_kind_kernel, cdiv(k, 256), 256, 0, st, d_bbox, d_size, (long long)k, d_kind);
    if (n > 0 && d_row_kind) {
        if (k > 0) LAUNCH(summary_rowkind_kernel, cdiv(n, 256), 256, 0, st, d_labels, d_kind, (long long)n, (long long)k, d_row_kind);
        else CU_TRY(cudaMemsetAsync(d_row_kind, 0, n, st));
    }
    return 0;
}

int row_kinds(const int32_t* d_labels, int64_t n, const uint8_t* d_kind, int64_t k, uint8_t* d_row_kind, cudaStream_t st) {
    if (n <= 0) return 0;
    if (k > 0) LAUNCH(summary_rowkind_kernel, cdiv(n, 256), 256, 0, st, d_labels, d_kind, (long long)n, (long long)k, d_row_kind);
    else CU_TRY(cudaMemsetAsync(d_row_kind, 0, n, st));
    return 0;
}

}  // namespace cloops
