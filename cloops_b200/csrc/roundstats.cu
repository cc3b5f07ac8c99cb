// Distance statistics of one clustering round, reduced on the device (cLoops/pipe.py:59-63,106-109,120-127,259 and
// cLoops/ests.py:36-61).  The reference pools two Python lists over all chromosomes -- dis: Y-X of the members of
// inter-ligation clusters; dss: Y-X of the rows removed by the cut filter plus the members of self-ligation clusters --
// and derives the next round's distance cut-off from the mean / std / median of their log2.  Here every chromosome adds
// to one accumulator per round: an exact histogram of the positive self-ligation distances (the median is an order
// statistic; log2 is monotone) and (count, sum, sum of squares) of log2|d| in float64 for the inter-ligation distances.
// The same three moments of the self-ligation distances are NOT accumulated per PET: they follow from the histogram
// (sum over bins of count * log2(bin)), evaluated once per round in cloops_round_middle -- 2^20 logarithms per round
// instead of one per removed / self-ligation PET per chromosome (ncu: the per-PET float64 log2 made the kernel
// instruction-bound, 100 M warp instructions for 23 M rows); only distances beyond the histogram (>= 2^20) keep their
// per-PET terms.  Across GPUs the accumulators are all-reduced (NCCL) -- 4 MB instead of the distance lists themselves.
//
// A chromosome without inter-ligation clusters contributes nothing, not even its dss (pipe.py:121-122): the kernels
// read the chromosome's inter-ligation cluster count on the device and return early.
#include "common.cuh"

namespace cloops {

#define RS_BINS CLOOPS_ROUND_HIST_BINS
#define RS_LOCAL 4096            // distances below this are histogrammed in shared memory first (self-ligation lump)
#define RS_GRID 592              // 4 CTAs per SM
#define RS_NQ 8

__global__ void __launch_bounds__(256) count_inter_kernel(const unsigned char* __restrict__ kind, int k, int* __restrict__ n_inter) {
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    const int t = __syncthreads_count(c < k && kind[c] == 1);
    if (threadIdx.x == 0 && t) atomicAdd(n_inter, t);
}

// members: [n_members] (X, Y, kind) of the clustered PETs (index order, or row order for blockDBSCAN);
// raw: [n_raw] the chromosome's rows, scanned for the ones the cut filter removed (signed Y-X < cut), only when cut > 0.
template <int THREADS>
__global__ void __launch_bounds__(THREADS) dist_stats_kernel(const int* __restrict__ xs, const int* __restrict__ ys,
                                                         const unsigned char* __restrict__ member_kind, int n_members,
                                                         const int* __restrict__ raw_x, const int* __restrict__ raw_y, int n_raw, int cut,
                                                         const int* __restrict__ n_inter, int* __restrict__ hist,
                                                         double* __restrict__ partial) {
    __shared__ int s_hist[RS_LOCAL];
    __shared__ double s_red[THREADS / 32][RS_NQ];
    if (*n_inter == 0) return;                                     // pipe.py:121-122
    for (int t = threadIdx.x; t < RS_LOCAL; t += blockDim.x) s_hist[t] = 0;
    __syncthreads();
    double q[RS_NQ] = {0, 0, 0, 0, 0, 0, 0, 0};                    // n_i, S_i, Q_i, n_s, S_s, Q_s, raw_i, raw_s
    const long long total = (long long)n_members + (cut > 0 ? n_raw : 0);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        int d, which;                                              // which: 1 inter, 2 self, 0 neither
        if (i < n_members) {
            which = member_kind[i];
            d = ys[i] - xs[i];
        } else {
            const long long j = i - n_members;
            d = raw_y[j] - raw_x[j];
            which = d < cut ? 2 : 0;
        }
        if (which == 0) continue;
        const unsigned a = (unsigned)(d < 0 ? -d : d);              // ests.py:40-41 np.abs
        if (which == 1) q[6] += 1.0; else q[7] += 1.0;
        if (a == 0) continue;                                      // ests.py:44-45: zero distances are dropped
        if (which == 2) {
            if (a < RS_LOCAL) { atomicAdd(&s_hist[a], 1); continue; }
            if (a < RS_BINS) { atomicAdd(&hist[a], 1); continue; }
            atomicAdd(&hist[RS_BINS], 1);                          // beyond the histogram: counted there, moments per PET
        }
        const double lg = log2((double)a);
        if (which == 1) { q[0] += 1.0; q[1] += lg; q[2] += lg * lg; }
        else { q[3] += 1.0; q[4] += lg; q[5] += lg * lg; }
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < RS_NQ; ++k) {
        double v = q[k];
        for (int off = 16; off > 0; off >>= 1) v += __shfl_down_sync(0xffffffffu, v, off);
        if (lane == 0) s_red[warp][k] = v;
    }
    __syncthreads();
    if (threadIdx.x < RS_NQ) {
        double v = 0;
        for (int w = 0; w < THREADS / 32; ++w) v += s_red[w][threadIdx.x];    // fixed order: the sums are reproducible
        partial[blockIdx.x * RS_NQ + threadIdx.x] = v;
    }
    for (int t = threadIdx.x; t < RS_LOCAL; t += blockDim.x)
        if (s_hist[t]) atomicAdd(&hist[t], s_hist[t]);
}

// mom: 0 n_i, 1 S_i, 2 Q_i, 3..5 the same for the self-ligation distances >= 2^20 only (the rest is in the histogram),
// 6 len(dis), 7 len(dss), 8 chromosomes with inter-ligation clusters
__global__ void __launch_bounds__(32 * RS_NQ) dist_stats_commit_kernel(const double* __restrict__ partial, int nblocks, const int* __restrict__ n_inter,
                                                                       double* __restrict__ mom) {
    if (*n_inter == 0) return;
    const int k = threadIdx.x >> 5, lane = threadIdx.x & 31;       // one warp per quantity; fixed summation order
    double v = 0;
    for (int b = lane; b < nblocks; b += 32) v += partial[b * RS_NQ + k];
    for (int off = 16; off > 0; off >>= 1) v += __shfl_down_sync(0xffffffffu, v, off);
    if (lane == 0) atomicAdd(&mom[k], v);                           // passes of several chromosomes may run on different streams
    if (threadIdx.x == 0) atomicAdd(&mom[8], 1.0);
}

// Order statistics and moments of the histogrammed self-ligation distances.  Two launches: per 1024-bin chunk the count
// and the two log2 sums (coalesced, fixed summation order); then one CTA scans the chunk counts, finds the chunk of each
// middle rank and scans that chunk, and adds up the chunk sums.
// out[0] = value of rank (k-1)/2, out[1] = value of rank k/2 (0-based; k = number of histogrammed values, overflow bin
// included); -1 when k == 0, RS_BINS when the rank lies in the overflow bin.  out[2] = k.  sums[0..1] = sum of log2 and of
// log2^2 over the values below 2^20.
#define RS_CHUNK 1024
#define RS_NCHUNK ((RS_BINS + 1 + RS_CHUNK - 1) / RS_CHUNK)
__global__ void __launch_bounds__(256) hist_chunk_kernel(const int* __restrict__ hist, long long* __restrict__ chunk, double* __restrict__ csum) {
    __shared__ long long s_w[8];
    __shared__ double s_s[8], s_q[8];
    long long v = 0;
    double sl = 0, sq = 0;
    for (int t = threadIdx.x; t < RS_CHUNK; t += 256) {
        const int b = blockIdx.x * RS_CHUNK + t;
        if (b > RS_BINS) continue;
        const int c = hist[b];
        v += c;
        if (c && b >= 1 && b < RS_BINS) {
            const double lg = log2((double)b);
            sl += c * lg;
            sq += c * (lg * lg);
        }
    }
    for (int off = 16; off > 0; off >>= 1) {
        v += __shfl_down_sync(0xffffffffu, v, off);
        sl += __shfl_down_sync(0xffffffffu, sl, off);
        sq += __shfl_down_sync(0xffffffffu, sq, off);
    }
    if ((threadIdx.x & 31) == 0) { s_w[threadIdx.x >> 5] = v; s_s[threadIdx.x >> 5] = sl; s_q[threadIdx.x >> 5] = sq; }
    __syncthreads();
    if (threadIdx.x == 0) {
        long long tot = 0;
        double a = 0, b = 0;
        for (int w = 0; w < 8; ++w) { tot += s_w[w]; a += s_s[w]; b += s_q[w]; }
        chunk[blockIdx.x] = tot;
        csum[2 * blockIdx.x] = a;
        csum[2 * blockIdx.x + 1] = b;
    }
}

// inclusive scan of s[0..1023] in place (1024 threads); the caller synchronises before reading
__device__ __forceinline__ void scan1024(long long* s) {
    for (int d = 1; d < 1024; d <<= 1) {
        const long long add = (int)threadIdx.x >= d ? s[threadIdx.x - d] : 0;
        __syncthreads();
        s[threadIdx.x] += add;
        __syncthreads();
    }
}

__global__ void __launch_bounds__(1024) hist_middle_kernel(const int* __restrict__ hist, const long long* __restrict__ chunk,
                                                           const double* __restrict__ csum, long long* __restrict__ out,
                                                           double* __restrict__ sums) {
    __shared__ long long s_incl[1024];                       // inclusive counts of chunks 0..1023 (chunk 1024 = the overflow bin alone)
    __shared__ long long s_bins[RS_CHUNK];
    const int tid = threadIdx.x;
    s_incl[tid] = chunk[tid];
    __syncthreads();
    scan1024(s_incl);
    const long long k = s_incl[1023] + chunk[RS_NCHUNK - 1];
    if (tid < 64) {                                          // chunk sums: warp 0 the log2 sums, warp 1 the squares; fixed order
        const int which = tid >> 5, lane = tid & 31;
        double v = 0;
        for (int c = lane; c < RS_NCHUNK; c += 32) v += csum[2 * c + which];
        for (int off = 16; off > 0; off >>= 1) v += __shfl_down_sync(0xffffffffu, v, off);
        if (lane == 0) sums[which] = v;
    }
    if (tid == 0) out[2] = k;
    if (k == 0) {
        if (tid == 0) { out[0] = -1; out[1] = -1; }
        return;
    }
    for (int w = 0; w < 2; ++w) {
        const long long r = w == 0 ? (k - 1) / 2 : k / 2;
        int lo = 0, hi = 1024;                               // first chunk whose inclusive count exceeds r (1024: the overflow bin)
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (s_incl[mid] <= r) lo = mid + 1; else hi = mid;
        }
        const int c = lo;
        if (c == 1024) {
            if (tid == 0) out[w] = RS_BINS;
            continue;
        }
        const long long before = c ? s_incl[c - 1] : 0;
        __syncthreads();
        s_bins[tid] = hist[c * RS_CHUNK + tid];
        __syncthreads();
        scan1024(s_bins);
        const long long incl = s_bins[tid], excl = tid ? s_bins[tid - 1] : 0;
        if (before + excl <= r && r < before + incl) out[w] = (long long)c * RS_CHUNK + tid;
    }
}

int pass_distance_stats(const int* xs, const int* ys, const unsigned char* member_kind, int n_members, const unsigned char* kind, int k,
                        const int* raw_x, const int* raw_y, int n_raw, int cut, int* d_hist, double* d_mom, cudaStream_t st) {
    Temp tmp(st);
    int* d_ninter;
    double* d_partial;
    RET_IF(tmp.alloc(&d_ninter, 1));
    RET_IF(tmp.alloc(&d_partial, (size_t)RS_GRID * RS_NQ));
    CU_TRY(cudaMemsetAsync(d_ninter, 0, sizeof(int), st));
    if (k > 0) LAUNCH(count_inter_kernel, cdiv(k, 256), 256, 0, st, kind, k, d_ninter);
    // every CTA ends by adding its non-empty shared bins to the global histogram: fewer, larger CTAs mean fewer of those
    // atomics (CLOOPS_RS_WIDE=0: 592 CTAs of 256 threads, the form measured first)
    static const bool wide = !(getenv("CLOOPS_RS_WIDE") && getenv("CLOOPS_RS_WIDE")[0] == '0');
    const int grid = wide ? RS_GRID / 4 : RS_GRID;
    if (wide) LAUNCH(dist_stats_kernel<1024>, grid, 1024, 0, st, xs, ys, member_kind, n_members, raw_x, raw_y, n_raw, cut, d_ninter, d_hist, d_partial);
    else LAUNCH(dist_stats_kernel<256>, grid, 256, 0, st, xs, ys, member_kind, n_members, raw_x, raw_y, n_raw, cut, d_ninter, d_hist, d_partial);
    LAUNCH(dist_stats_commit_kernel, 1, 32 * RS_NQ, 0, st, d_partial, grid, d_ninter, d_mom);
    return 0;
}

}  // namespace cloops

using namespace cloops;

extern "C" int cloops_round_middle(const int32_t* d_hist, const double* d_mom, int64_t* h_middle, double* h_mom, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (!d_hist || !d_mom || !h_middle || !h_mom) return fail(CLOOPS_EINVAL, "NULL argument");
    RET_IF(pool_init());
    Temp tmp(st);
    long long *d_out, *d_chunk;
    double *d_csum, *d_sums;
    RET_IF(tmp.alloc(&d_out, 3));
    RET_IF(tmp.alloc(&d_chunk, RS_NCHUNK));
    RET_IF(tmp.alloc(&d_csum, 2 * RS_NCHUNK));
    RET_IF(tmp.alloc(&d_sums, 2));
    LAUNCH(hist_chunk_kernel, RS_NCHUNK, 256, 0, st, d_hist, d_chunk, d_csum);
    LAUNCH(hist_middle_kernel, 1, 1024, 0, st, d_hist, d_chunk, d_csum, d_out, d_sums);
    long long out[3];
    double sums[2];
    CU_TRY(cudaMemcpyAsync(out, d_out, sizeof(out), cudaMemcpyDeviceToHost, st));
    CU_TRY(cudaMemcpyAsync(sums, d_sums, sizeof(sums), cudaMemcpyDeviceToHost, st));
    CU_TRY(cudaMemcpyAsync(h_mom, d_mom, CLOOPS_ROUND_MOM * sizeof(double), cudaMemcpyDeviceToHost, st));
    CU_TRY(cudaStreamSynchronize(st));
    h_middle[0] = out[0];
    h_middle[1] = out[1];
    // the self-ligation moments: every positive distance is in the histogram (out[2] of them); the accumulators hold the
    // log2 terms of the ones beyond it, the histogram gives the rest
    h_mom[3] = (double)out[2];
    h_mom[4] += sums[0];
    h_mom[5] += sums[1];
    return 0;
}
