// extern "C" surface of libcloops_b200 (declared in include/cloops_b200.h).
#include <limits.h>

#include "index.cuh"

namespace cloops {
int index_dbscan(cloops_index* ix, int minPts, int variant, int* d_labels, int* d_labels_sorted, int64_t* h_info, cudaStream_t st);
int row_kinds(const int32_t* d_labels, int64_t n, const uint8_t* d_kind, int64_t k, uint8_t* d_row_kind, cudaStream_t st);
int block_dbscan(const int32_t* d_x, const int32_t* d_y, int64_t n, int32_t eps, int32_t minPts, int32_t cut, int32_t* d_labels,
                 int64_t* h_info, cudaStream_t st);
int cluster_summary(const int32_t* d_x, const int32_t* d_y, const int32_t* d_labels, int64_t n, int64_t k, int32_t* d_bbox,
                    int32_t* d_size, uint8_t* d_kind, uint8_t* d_row_kind, cudaStream_t st);

__global__ void __launch_bounds__(256) scatter_counts_kernel(const int* __restrict__ cnt, const u32* __restrict__ rows, int n_act,
                                                             int* __restrict__ out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_act) out[rows[i]] = cnt[i];
}

// int64 [n,3] (id, X, Y) -> int32 x, y with range check
__global__ void __launch_bounds__(256) split_mat_kernel(const long long* __restrict__ mat, long long n, int* __restrict__ x,
                                                        int* __restrict__ y, int* __restrict__ bad) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    long long xx = mat[3 * i + 1], yy = mat[3 * i + 2];
    if (xx < -(1LL << 30) || xx >= (1LL << 30) || yy < -(1LL << 30) || yy >= (1LL << 30)) *bad = 1;
    x[i] = (int)xx;
    y[i] = (int)yy;
}

__global__ void __launch_bounds__(256) widen_labels_kernel(const int* __restrict__ in, long long n, long long* __restrict__ out) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = in[i];
}
}  // namespace cloops

using namespace cloops;

extern "C" {

int cloops_index_build(const int32_t* d_x, const int32_t* d_y, int64_t n, int32_t eps, int32_t cut, cloops_index** out,
                       void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (!out) return fail(CLOOPS_EINVAL, "out is NULL");
    *out = nullptr;
    stages_begin(st);
    int rc = index_build(d_x, d_y, n, eps, cut, out, st);
    if (rc != 0) {
        index_free(*out, st);
        *out = nullptr;
        return rc;
    }
    return stages_end(st);
}

void cloops_index_free(cloops_index* ix) { index_free(ix, 0); }

void cloops_index_release(cloops_index* ix, void* stream) { index_free(ix, (cudaStream_t)stream); }

int64_t cloops_index_n_active(const cloops_index* ix) { return ix ? ix->P.n_act : 0; }

int cloops_index_count(cloops_index* ix, int32_t cap, int32_t* d_counts_sorted, void* stream) {
    if (!ix) return fail(CLOOPS_EINVAL, "index is NULL");
    return index_count(ix, cap, d_counts_sorted, (cudaStream_t)stream);
}

int cloops_index_dbscan(cloops_index* ix, int32_t minPts, int32_t variant, int32_t* d_labels, int32_t* d_labels_sorted,
                        int64_t* h_info, void* stream) {
    if (!ix) return fail(CLOOPS_EINVAL, "index is NULL");
    cudaStream_t st = (cudaStream_t)stream;
    stages_begin(st);
    RET_IF(index_dbscan(ix, minPts, variant, d_labels, d_labels_sorted, h_info, st));
    return stages_end(st);
}

int cloops_index_coords(cloops_index* ix, int32_t* d_xs, int32_t* d_ys, void* stream) {
    if (!ix) return fail(CLOOPS_EINVAL, "index is NULL");
    return index_coords(ix, d_xs, d_ys, (cudaStream_t)stream);
}

int cloops_row_kinds(const int32_t* d_labels, int64_t n, const uint8_t* d_kind, int64_t n_clusters, uint8_t* d_row_kind, void* stream) {
    return row_kinds(d_labels, n, d_kind, n_clusters, d_row_kind, (cudaStream_t)stream);
}

int cloops_dbscan(const int32_t* d_x, const int32_t* d_y, int64_t n, int32_t eps, int32_t minPts, int32_t cut, int32_t variant,
                  int32_t* d_labels, int64_t* h_info, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (minPts < 1) return fail(CLOOPS_EINVAL, "minPts must be >= 1 (got %d)", minPts);
    if (variant == CLOOPS_BLOCK) {
        stages_begin(st);
        RET_IF(block_dbscan(d_x, d_y, n, eps, minPts, cut, d_labels, h_info, st));
        return stages_end(st);
    }
    if (variant != CLOOPS_V1 && variant != CLOOPS_V2) return fail(CLOOPS_EINVAL, "unknown variant %d", variant);
    stages_begin(st);
    cloops_index* ix = nullptr;
    int rc = index_build(d_x, d_y, n, eps, cut, &ix, st);
    if (rc == 0) rc = index_dbscan(ix, minPts, variant, d_labels, nullptr, h_info, st);
    index_free(ix, st);
    if (rc != 0) return rc;
    return stages_end(st);
}

int cloops_dbscan_host(const int64_t* h_mat, int64_t n, int32_t eps, int32_t minPts, int32_t variant, int64_t* h_labels,
                       int64_t* h_info) {
    if (n < 0) return fail(CLOOPS_EINVAL, "n < 0");
    if (h_info) for (int k = 0; k < 8; ++k) h_info[k] = 0;
    if (n == 0) return 0;
    RET_IF(pool_init());
    cudaStream_t st = 0;
    Temp tmp(st);
    long long* d_mat;
    int *d_x, *d_y, *d_lab, *d_bad;
    RET_IF(tmp.alloc(&d_mat, (size_t)3 * n));
    RET_IF(tmp.alloc(&d_x, n));
    RET_IF(tmp.alloc(&d_y, n));
    RET_IF(tmp.alloc(&d_lab, n));
    RET_IF(tmp.alloc(&d_bad, 1));
    CU_TRY(cudaMemsetAsync(d_bad, 0, sizeof(int), st));
    CU_TRY(cudaMemcpyAsync(d_mat, h_mat, (size_t)3 * n * sizeof(long long), cudaMemcpyHostToDevice, st));
    LAUNCH(split_mat_kernel, cdiv(n, 256), 256, 0, st, d_mat, (long long)n, d_x, d_y, d_bad);
    int bad = 0;
    CU_TRY(cudaMemcpyAsync(&bad, d_bad, sizeof(int), cudaMemcpyDeviceToHost, st));
    CU_TRY(cudaStreamSynchronize(st));
    if (bad) return fail(CLOOPS_ERANGE, "coordinates must lie in [-2^30, 2^30)");
    RET_IF(cloops_dbscan(d_x, d_y, n, eps, minPts, 0, variant, d_lab, h_info, st));
    LAUNCH(widen_labels_kernel, cdiv(n, 256), 256, 0, st, d_lab, (long long)n, d_mat);   // reuse d_mat as int64 out
    CU_TRY(cudaMemcpyAsync(h_labels, d_mat, (size_t)n * sizeof(long long), cudaMemcpyDeviceToHost, st));
    CU_TRY(cudaStreamSynchronize(st));
    return 0;
}

int cloops_neighbour_counts(const int32_t* d_x, const int32_t* d_y, int64_t n, int32_t eps, int32_t cap, int32_t cut,
                            int32_t* d_counts, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    stages_begin(st);
    cloops_index* ix = nullptr;
    int rc = index_build(d_x, d_y, n, eps, cut, &ix, st);
    if (rc == 0 && n > 0) {
        Temp tmp(st);
        int* d_cnt = nullptr;
        rc = cudaMemsetAsync(d_counts, 0, (size_t)n * sizeof(int), st) == cudaSuccess ? 0 : fail(CLOOPS_ECUDA, "memset failed");
        if (rc == 0 && ix->P.n_act > 0) {
            rc = tmp.alloc(&d_cnt, ix->P.n_act);
            if (rc == 0) rc = index_count(ix, cap, d_cnt, st);
            stage_mark("region_query", st);
            if (rc == 0) {
                scatter_counts_kernel<<<cdiv(ix->P.n_act, 256), 256, 0, st>>>(d_cnt, ix->rows, ix->P.n_act, d_counts);
                g_launches.fetch_add(1);
            }
        }
    }
    index_free(ix, st);
    if (rc != 0) return rc;
    return stages_end(st);
}

int cloops_cluster_summary(const int32_t* d_x, const int32_t* d_y, const int32_t* d_labels, int64_t n, int64_t n_clusters,
                           int32_t* d_bbox, int32_t* d_size, uint8_t* d_kind, uint8_t* d_row_kind, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    stages_begin(st);
    RET_IF(cluster_summary(d_x, d_y, d_labels, n, n_clusters, d_bbox, d_size, d_kind, d_row_kind, st));
    stage_mark("summary", st);
    return stages_end(st);
}

}  // extern "C"
