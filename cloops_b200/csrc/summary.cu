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

// Rows arrive in file order, so labels inside a warp are unrelated in general; a warp whose lanes all
// carry the same label (giant Hi-C diagonal clusters) reduces in registers and issues one atomic set.
__global__ void __launch_bounds__(256) summary_accumulate_kernel(const int* __restrict__ x, const int* __restrict__ y,
                                                                 const int* __restrict__ labels, long long n, long long k,
                                                                 int* __restrict__ bbox, int* __restrict__ size) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    int lab = (i < n) ? __ldg(labels + i) : -1;
    if (lab >= k) lab = -1;
    int xx = 0, yy = 0;
    if (lab >= 0) { xx = __ldg(x + i); yy = __ldg(y + i); }
    const unsigned full = 0xffffffffu;
    int lab0 = __shfl_sync(full, lab, 0);
    if (__all_sync(full, lab == lab0)) {
        if (lab0 < 0) return;
        int x0 = __reduce_min_sync(full, xx), x1 = __reduce_max_sync(full, xx);
        int y0 = __reduce_min_sync(full, yy), y1 = __reduce_max_sync(full, yy);
        if ((threadIdx.x & 31) == 0) {
            int* b = bbox + 4LL * lab0;
            atomicMin(b + 0, x0); atomicMax(b + 1, x1); atomicMin(b + 2, y0); atomicMax(b + 3, y1);
            atomicAdd(size + lab0, 32);
        }
        return;
    }
    if (lab < 0) return;
    int* b = bbox + 4LL * lab;
    atomicMin(b + 0, xx); atomicMax(b + 1, xx); atomicMin(b + 2, yy); atomicMax(b + 3, yy);
    atomicAdd(size + lab, 1);
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
        LAUNCH(summary_accumulate_kernel, cdiv(n, 256), 256, 0, st, d_x, d_y, d_labels, (long long)n, (long long)k, d_bbox, d_size);
    if (k > 0) LAUNCH(summary_kind_kernel, cdiv(k, 256), 256, 0, st, d_bbox, d_size, (long long)k, d_kind);
    if (n > 0 && d_row_kind) {
        if (k > 0) LAUNCH(summary_rowkind_kernel, cdiv(n, 256), 256, 0, st, d_labels, d_kind, (long long)n, (long long)k, d_row_kind);
        else CU_TRY(cudaMemsetAsync(d_row_kind, 0, n, st));
    }
    return 0;
}

}  // namespace cloops
