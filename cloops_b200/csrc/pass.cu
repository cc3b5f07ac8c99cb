// One pass of the whole hot path over one chromosome as ONE C-ABI call (cloops_pass_*):
// cluster (cLoops/pipe.py:52-75) -> per-cluster candidate records and dis/dss membership (:76-109) ->
// coverage model (cModel.py:45-57) -> permuted-background range counts of every inter-ligation
// candidate (cModel.py:118-143).  Everything stays in HBM; the host sees sizes only.  The coverage
// build (two radix sorts, independent of the clustering) runs on a side stream.
#include <limits.h>

#include "index.cuh"
#include "scan.cuh"

struct cloops_coverage;

namespace cloops {
int index_dbscan(cloops_index* ix, int minPts, int variant, int* d_labels, int* d_labels_sorted, int64_t* h_info, cudaStream_t st);
int block_dbscan(const int32_t* d_x, const int32_t* d_y, int64_t n, int32_t eps, int32_t minPts, int32_t cut, int32_t* d_labels,
                 int64_t* h_info, cudaStream_t st);
int cluster_summary(const int32_t* d_x, const int32_t* d_y, const int32_t* d_labels, int64_t n, int64_t k, int32_t* d_bbox,
                    int32_t* d_size, uint8_t* d_kind, uint8_t* d_row_kind, cudaStream_t st);
int row_kinds(const int32_t* d_labels, int64_t n, const uint8_t* d_kind, int64_t k, uint8_t* d_row_kind, cudaStream_t st);
int coverage_build(const int32_t* d_x, const int32_t* d_y, int64_t n, cloops_coverage** out, cudaStream_t st);
int pass_distance_stats(const int* xs, const int* ys, const unsigned char* member_kind, int n_members, const unsigned char* kind, int k,
                        const int* raw_x, const int* raw_y, int n_raw, int cut, int* d_hist, double* d_mom, cudaStream_t st);
int range_counts_dev(const cloops_coverage* cov, const int32_t* d_cand, int64_t ncand, const int* d_ncand, int32_t* d_out,
                     cudaStream_t st);
}  // namespace cloops

extern "C" void cloops_coverage_release(cloops_coverage* cov, void* stream);

struct cloops_pass {
    int64_t info[8] = {0};
    int n = 0;            // rows
    int n_members = 0;    // length of the member arrays (active PETs in index order; rows for blockDBSCAN)
    int k = 0;            // cluster ids (max id + 1)
    int m = 0;            // inter-ligation candidates
    int scored = 0;
    int *d_x = nullptr, *d_y = nullptr;          // owned copies when the pass was started from host buffers
    int *bbox = nullptr, *size = nullptr;        // [k,4], [k]
    uint8_t* kind = nullptr;                     // [k]
    int *xs = nullptr, *ys = nullptr, *labels = nullptr;   // [n_members]
    uint8_t* member_kind = nullptr;              // [n_members]
    int *cand = nullptr, *counts = nullptr, *d_m = nullptr;  // [k,4], [k,123] (first m rows used), [1]
    bool xs_owned = true;
};

namespace cloops {

__global__ void __launch_bounds__(256) cand_flag_kernel(const unsigned char* __restrict__ kind, int k, int* __restrict__ flag) {
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c < k) flag[c] = kind[c] == 1 ? 1 : 0;
}

// candidates in ascending cluster id (pipe.py:78-102 iterates ids in order), clamped at 0 (cModel.py:281-282)
__global__ void __launch_bounds__(256) cand_scatter_kernel(const unsigned char* __restrict__ kind, const int* __restrict__ pos,
                                                           const int* __restrict__ bbox, int k, int* __restrict__ cand,
                                                           int* __restrict__ d_m) {
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= k) return;
    if (kind[c] == 1) {
        int4 b = reinterpret_cast<const int4*>(bbox)[c];
        reinterpret_cast<int4*>(cand)[pos[c]] = make_int4(max(b.x, 0), b.y, max(b.z, 0), b.w);
    }
    if (c == k - 1) *d_m = pos[c] + (kind[c] == 1 ? 1 : 0);
}

// side stream + fork/join events, one set per (host thread, device)
#define MAX_DEVICES 64
struct Side {
    cudaStream_t stream = nullptr;
    cudaEvent_t fork = nullptr, join = nullptr;
};
static thread_local Side g_sides[MAX_DEVICES];

static int side_get(Side** out) {
    int dev = 0;
    CU_TRY(cudaGetDevice(&dev));
    if (dev < 0 || dev >= MAX_DEVICES) return fail(CLOOPS_EINVAL, "device ordinal %d out of range", dev);
    Side& s = g_sides[dev];
    if (!s.stream) {
        CU_TRY(cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking));
        CU_TRY(cudaEventCreateWithFlags(&s.fork, cudaEventDisableTiming));
        CU_TRY(cudaEventCreateWithFlags(&s.join, cudaEventDisableTiming));
    }
    *out = &s;
    return 0;
}

template <class T>
static int dalloc(T** p, size_t count, cudaStream_t st) {
    size_t bytes = count * sizeof(T);
    if (bytes == 0) bytes = 16;
    cudaError_t e = cudaMallocAsync((void**)p, bytes, st);
    if (e != cudaSuccess) return fail(CLOOPS_ENOMEM, "cudaMallocAsync(%zu) failed: %s", bytes, cudaGetErrorString(e));
    return 0;
}

static void pass_release(cloops_pass* p, cudaStream_t st) {
    if (!p) return;
    void* ptrs[] = {p->d_x, p->d_y, p->bbox, p->size, p->kind, p->xs_owned ? p->xs : nullptr, p->xs_owned ? p->ys : nullptr,
                    p->labels, p->member_kind, p->cand, p->counts, p->d_m};
    for (void* q : ptrs)
        if (q) cudaFreeAsync(q, st);
    delete p;
}

static int pass_run(cloops_pass* p, const int32_t* d_x, const int32_t* d_y, int64_t n, int32_t eps, int32_t minPts, int32_t cut,
                    int32_t variant, int32_t score, cudaStream_t st, int32_t* d_hist = nullptr, double* d_mom = nullptr,
                    const cloops_index* base = nullptr) {
    // every failure inside leaves through the common clean-up at the bottom (index, coverage model, join with the side stream)
#define CU_BRK(expr)                                                                                                    \
    {                                                                                                                   \
        cudaError_t _e = (expr);                                                                                        \
        if (_e != cudaSuccess) { rc = fail(CLOOPS_ECUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); break; } \
    }
    p->n = (int)n;
    if (n == 0) return 0;
    RET_IF(pool_init());
    Temp scope(st, WS_PASS);                       // the scratch arrays of every stage below come from one workspace, settled once per pass
    cloops_coverage* cov = nullptr;
    Side* side = nullptr;
    int rc = 0;
    cloops_index* ix = nullptr;
    bool forked = false;
    do {
        if (score) {                               // fork: coverage sorts on the side stream
            if ((rc = side_get(&side))) break;
            CU_BRK(cudaEventRecord(side->fork, st));
            CU_BRK(cudaStreamWaitEvent(side->stream, side->fork, 0));
            forked = true;
            rc = coverage_build(d_x, d_y, n, &cov, side->stream);
            cudaEventRecord(side->join, side->stream);
            if (rc) break;
        }
        if (variant == CLOOPS_BLOCK) {             // no strip index: row order throughout
            p->n_members = (int)n;
            p->xs_owned = false;
            p->xs = const_cast<int*>(d_x);
            p->ys = const_cast<int*>(d_y);
            if ((rc = dalloc(&p->labels, n, st))) break;
            if ((rc = block_dbscan(d_x, d_y, n, eps, minPts, cut, p->labels, p->info, st))) break;
        } else {
            if ((rc = base ? index_filter(base, cut, &ix, st) : index_build(d_x, d_y, n, eps, cut, &ix, st))) break;
            p->n_members = ix->P.n_act;
            if ((rc = dalloc(&p->labels, p->n_members, st))) break;
            if ((rc = index_dbscan(ix, minPts, variant, nullptr, p->labels, p->info, st))) break;
            stage_mark("cluster", st);
            if ((rc = dalloc(&p->xs, p->n_members, st)) || (rc = dalloc(&p->ys, p->n_members, st))) break;
            if ((rc = index_coords(ix, p->xs, p->ys, st))) break;
        }
        p->k = (int)p->info[1];
        const int k = p->k;
        if ((rc = dalloc(&p->bbox, (size_t)4 * k, st)) || (rc = dalloc(&p->size, k, st)) || (rc = dalloc(&p->kind, k, st))) break;
        if ((rc = dalloc(&p->member_kind, p->n_members, st))) break;
        if ((rc = cluster_summary(p->xs, p->ys, p->labels, p->n_members, k, p->bbox, p->size, p->kind, nullptr, st))) break;
        if ((rc = row_kinds(p->labels, p->n_members, p->kind, k, p->member_kind, st))) break;
        stage_mark("summary", st);
        if (d_hist && d_mom) {
            if ((rc = pass_distance_stats(p->xs, p->ys, p->member_kind, p->n_members, p->kind, k, d_x, d_y, (int)n, cut, d_hist, d_mom, st))) break;
            stage_mark("distance_stats", st);
        }
        if (score && k > 0) {
            Temp tmp(st);
            int *flag, *pos;
            if ((rc = tmp.alloc(&flag, k)) || (rc = tmp.alloc(&pos, k))) break;
            if ((rc = dalloc(&p->cand, (size_t)4 * k, st)) || (rc = dalloc(&p->counts, (size_t)123 * k, st)) || (rc = dalloc(&p->d_m, 1, st))) break;
            cand_flag_kernel<<<cdiv(k, 256), 256, 0, st>>>(p->kind, k, flag);
            g_launches.fetch_add(1);
            int* d_scan;
            if ((rc = tmp.alloc(&d_scan, scan_tmp_ints(k)))) break;
            if ((rc = device_scan<SCAN_ADD, false, false>(flag, pos, k, d_scan, st))) break;
            cand_scatter_kernel<<<cdiv(k, 256), 256, 0, st>>>(p->kind, pos, p->bbox, k, p->cand, p->d_m);
            g_launches.fetch_add(1);
            CU_BRK(cudaStreamWaitEvent(st, side->join, 0));      // join: coverage model ready
            if ((rc = range_counts_dev(cov, p->cand, k, p->d_m, p->counts, st))) break;
            CU_BRK(cudaMemcpyAsync(&p->m, p->d_m, sizeof(int), cudaMemcpyDeviceToHost, st));
            CU_BRK(cudaStreamSynchronize(st));
            p->scored = 1;
            stage_mark("range_counts", st);
        }
    } while (0);
#undef CU_BRK
    if (score) {
        if (forked) cudaStreamWaitEvent(st, side->join, 0);       // never free the model while the side stream builds it
        cloops_coverage_release(cov, st);
        if (rc == 0) p->scored = 1;
    }
    index_free(ix, st);
    return rc;
}

}  // namespace cloops

using namespace cloops;

extern "C" {

int cloops_pass_run(const int32_t* d_x, const int32_t* d_y, int64_t n, int32_t eps, int32_t minPts, int32_t cut, int32_t variant,
                    int32_t score, cloops_pass** out, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (!out) return fail(CLOOPS_EINVAL, "out is NULL");
    *out = nullptr;
    if (n < 0 || n > 0x7fffff00LL) return fail(CLOOPS_EINVAL, "n=%lld out of range", (long long)n);
    if (minPts < 1) return fail(CLOOPS_EINVAL, "minPts must be >= 1 (got %d)", minPts);
    if (variant != CLOOPS_V1 && variant != CLOOPS_V2 && variant != CLOOPS_BLOCK) return fail(CLOOPS_EINVAL, "unknown variant %d", variant);
    stages_begin(st);
    cloops_pass* p = new cloops_pass();
    int rc = pass_run(p, d_x, d_y, n, eps, minPts, cut, variant, score, st);
    if (rc != 0) {
        pass_release(p, st);
        return rc;
    }
    *out = p;
    return stages_end(st);
}

int cloops_pass_run_stats(const int32_t* d_x, const int32_t* d_y, int64_t n, int32_t eps, int32_t minPts, int32_t cut, int32_t variant,
                          int32_t score, int32_t* d_hist, double* d_mom, cloops_pass** out, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (!out) return fail(CLOOPS_EINVAL, "out is NULL");
    *out = nullptr;
    if (n < 0 || n > 0x7fffff00LL) return fail(CLOOPS_EINVAL, "n=%lld out of range", (long long)n);
    if (minPts < 1) return fail(CLOOPS_EINVAL, "minPts must be >= 1 (got %d)", minPts);
    if (variant != CLOOPS_V1 && variant != CLOOPS_V2 && variant != CLOOPS_BLOCK) return fail(CLOOPS_EINVAL, "unknown variant %d", variant);
    if (!d_hist || !d_mom) return fail(CLOOPS_EINVAL, "round accumulators are NULL");
    stages_begin(st);
    cloops_pass* p = new cloops_pass();
    int rc = pass_run(p, d_x, d_y, n, eps, minPts, cut, variant, score, st, d_hist, d_mom);
    if (rc != 0) {
        pass_release(p, st);
        return rc;
    }
    *out = p;
    return stages_end(st);
}

int cloops_pass_run_base(const cloops_index* base, const int32_t* d_x, const int32_t* d_y, int64_t n, int32_t minPts, int32_t cut,
                         int32_t variant, int32_t score, int32_t* d_hist, double* d_mom, cloops_pass** out, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (!out) return fail(CLOOPS_EINVAL, "out is NULL");
    *out = nullptr;
    if (!base) return fail(CLOOPS_EINVAL, "base index is NULL");
    if (n != base->P.n) return fail(CLOOPS_EINVAL, "the base index was built for %d rows, not %lld", base->P.n, (long long)n);
    if (minPts < 1) return fail(CLOOPS_EINVAL, "minPts must be >= 1 (got %d)", minPts);
    if (variant != CLOOPS_V1 && variant != CLOOPS_V2) return fail(CLOOPS_EINVAL, "variant %d has no strip index", variant);
    if ((d_hist == nullptr) != (d_mom == nullptr)) return fail(CLOOPS_EINVAL, "round accumulators: both or none");
    stages_begin(st);
    cloops_pass* p = new cloops_pass();
    int rc = pass_run(p, d_x, d_y, n, base->P.eps, minPts, cut, variant, score, st, d_hist, d_mom, base);
    if (rc != 0) {
        pass_release(p, st);
        return rc;
    }
    *out = p;
    return stages_end(st);
}

int cloops_pass_run_host(const int32_t* h_x, const int32_t* h_y, int64_t n, int32_t eps, int32_t minPts, int32_t cut,
                         int32_t variant, int32_t score, cloops_pass** out, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (!out) return fail(CLOOPS_EINVAL, "out is NULL");
    *out = nullptr;
    if (n < 0 || n > 0x7fffff00LL) return fail(CLOOPS_EINVAL, "n=%lld out of range", (long long)n);
    if (minPts < 1) return fail(CLOOPS_EINVAL, "minPts must be >= 1 (got %d)", minPts);
    if (variant != CLOOPS_V1 && variant != CLOOPS_V2 && variant != CLOOPS_BLOCK) return fail(CLOOPS_EINVAL, "unknown variant %d", variant);
    RET_IF(pool_init());
    stages_begin(st);
    cloops_pass* p = new cloops_pass();
    int rc = 0;
    if (n > 0) {
        rc = dalloc(&p->d_x, n, st);
        if (rc == 0) rc = dalloc(&p->d_y, n, st);
        if (rc == 0 && cudaMemcpyAsync(p->d_x, h_x, n * sizeof(int), cudaMemcpyHostToDevice, st) != cudaSuccess) rc = fail(CLOOPS_ECUDA, "H2D copy of X failed");
        if (rc == 0 && cudaMemcpyAsync(p->d_y, h_y, n * sizeof(int), cudaMemcpyHostToDevice, st) != cudaSuccess) rc = fail(CLOOPS_ECUDA, "H2D copy of Y failed");
        stage_mark("h2d", st);
    }
    if (rc == 0) rc = pass_run(p, p->d_x, p->d_y, n, eps, minPts, cut, variant, score, st);
    if (rc != 0) {
        pass_release(p, st);
        return rc;
    }
    *out = p;
    return stages_end(st);
}

void cloops_pass_free(cloops_pass* p, void* stream) { pass_release(p, (cudaStream_t)stream); }

/* sizes[0..5] = n_members, n_clusters, n_candidates (-> rows of counts), scored, n_rows, reserved; info as cloops_dbscan */
int cloops_pass_sizes(const cloops_pass* p, int64_t* sizes, int64_t* h_info) {
    if (!p) return fail(CLOOPS_EINVAL, "pass is NULL");
    if (sizes) { sizes[0] = p->n_members; sizes[1] = p->k; sizes[2] = p->scored ? p->m : 0; sizes[3] = p->scored; sizes[4] = p->n; sizes[5] = 0; }
    if (h_info) for (int i = 0; i < 8; ++i) h_info[i] = p->info[i];
    return 0;
}

/* device views, valid until cloops_pass_free: which = 0 bbox[k,4], 1 size[k], 2 kind[k] (u8), 3 xs, 4 ys, 5 labels (index
 * order, [n_members]), 6 member_kind[n_members] (u8), 7 cand[m,4], 8 counts[m,123] */
const void* cloops_pass_device_ptr(const cloops_pass* p, int which) {
    if (!p) return nullptr;
    switch (which) {
        case 0: return p->bbox; case 1: return p->size; case 2: return p->kind; case 3: return p->xs; case 4: return p->ys;
        case 5: return p->labels; case 6: return p->member_kind; case 7: return p->cand; case 8: return p->counts;
        default: return nullptr;
    }
}

/* Copies results to HOST buffers (any may be NULL) and synchronises the stream once:
 * h_bbox int32[k,4], h_kind u8[k], h_member_kind u8[n_members], h_counts int32[m,123]. */
int cloops_pass_fetch(const cloops_pass* p, int32_t* h_bbox, uint8_t* h_kind, uint8_t* h_member_kind, int32_t* h_counts, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (!p) return fail(CLOOPS_EINVAL, "pass is NULL");
    if (h_bbox && p->k) CU_TRY(cudaMemcpyAsync(h_bbox, p->bbox, (size_t)p->k * 4 * sizeof(int), cudaMemcpyDeviceToHost, st));
    if (h_kind && p->k) CU_TRY(cudaMemcpyAsync(h_kind, p->kind, (size_t)p->k, cudaMemcpyDeviceToHost, st));
    if (h_member_kind && p->n_members) CU_TRY(cudaMemcpyAsync(h_member_kind, p->member_kind, (size_t)p->n_members, cudaMemcpyDeviceToHost, st));
    if (h_counts && p->scored && p->m) CU_TRY(cudaMemcpyAsync(h_counts, p->counts, (size_t)p->m * 123 * sizeof(int), cudaMemcpyDeviceToHost, st));
    CU_TRY(cudaStreamSynchronize(st));
    return 0;
}

/* Candidate-record view of a pass for the host (pipe.py:76-102): h_bbox int32[k,4], h_size int32[k], h_kind u8[k]
 * (any may be NULL); one synchronisation. */
int cloops_pass_fetch_records(const cloops_pass* p, int32_t* h_bbox, int32_t* h_size, uint8_t* h_kind, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (!p) return fail(CLOOPS_EINVAL, "pass is NULL");
    if (h_bbox && p->k) CU_TRY(cudaMemcpyAsync(h_bbox, p->bbox, (size_t)p->k * 4 * sizeof(int), cudaMemcpyDeviceToHost, st));
    if (h_size && p->k) CU_TRY(cudaMemcpyAsync(h_size, p->size, (size_t)p->k * sizeof(int), cudaMemcpyDeviceToHost, st));
    if (h_kind && p->k) CU_TRY(cudaMemcpyAsync(h_kind, p->kind, (size_t)p->k, cudaMemcpyDeviceToHost, st));
    CU_TRY(cudaStreamSynchronize(st));
    return 0;
}

}  // extern "C"
