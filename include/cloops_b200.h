/*
 * cloops_b200 -- C ABI of the B200-native cLoops hot path (libcloops_b200.so).
 *
 * Plain pointers and sizes only.  Pointers named d_* are DEVICE pointers on the current CUDA device,
 * h_* are HOST pointers.  `stream` is a cudaStream_t passed as void* (NULL = legacy default stream).
 * Every function returns 0 on success or a negative CLOOPS_E* code; cloops_last_error() returns the
 * message of the last failure on the calling thread.  There is no CPU fallback anywhere.
 *
 * Reference interfaces replaced (paths relative to the cLoops 0.93 tree):
 *   cloops_dbscan*           cLoops/cDBSCAN2.py:13-35 (v2, default, pipe.py:42), cLoops/cDBSCAN.py:12-40 (v1),
 *                            cLoops/blockDBSCAN.py:13-41 (block): `DBSCAN(mat, eps, minPts).labels`
 *   cloops_neighbour_counts  the eps-neighbourhood query: cDBSCAN.py:186-205 regionQuery,
 *                            cDBSCAN2.py:304-346 getSparseCellNeighbor + :364-378 binSearchAdjPt
 *   cloops_cluster_summary   cLoops/pipe.py:76-109 (per-cluster bbox, zero-extent drop, inter/self split,
 *                            dis/dss membership)
 *   cloops_range_counts      cLoops/cModel.py:60-143 (getCounts / getPETsforRegions /
 *                            getNearbyPairRegions / the counting half of getMultiplePsFdr)
 *   cloops_pass_run*         cLoops/pipe.py:52-110 singleDBSCAN for one chromosome and round (cut filter, clusterer, records,
 *                            dis / dss membership) [+ cModel.py:262-295 getIntSig's counting half with score = 1]
 *   cloops_pass_run_stats / _base, cloops_round_middle
 *                            the pooled dis / dss lists of cLoops/pipe.py:113-127 and their reduction by
 *                            cLoops/ests.py:36-61 estIntSelCutFrag (mean / std / median of log2 distances)
 *   cloops_combine_rounds    cLoops/pipe.py:155-174 combineTwice (host C++)
 *   cloops_remove_dup        cLoops/cModel.py:198-259 removeDup (host C++)
 *   cloops_bedpe_*           cLoops/io.py:30-59,62-189,192-203 BEDPE ingest: PET, parseRawBedpe / parseRawBedpe2, txt2jd (host C++)
 */
#ifndef CLOOPS_B200_H
#define CLOOPS_B200_H

#include <stdint.h>

#if defined(__GNUC__)
#define CLOOPS_API __attribute__((visibility("default")))
#else
#define CLOOPS_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

#define CLOOPS_OK 0
#define CLOOPS_EINVAL (-1)   /* bad argument (eps < 1, minPts < 1, n < 0, unknown variant ...) */
#define CLOOPS_ERANGE (-2)   /* coordinates do not fit the packed 63-bit (strip,u,v) key */
#define CLOOPS_ECUDA (-3)    /* CUDA runtime error; message in cloops_last_error() */
#define CLOOPS_ENOMEM (-4)

/* clusterer variants: same three classes the reference ships */
#define CLOOPS_V1 1          /* cLoops.cDBSCAN.cDBSCAN        */
#define CLOOPS_V2 2          /* cLoops.cDBSCAN2.cDBSCAN       */
#define CLOOPS_BLOCK 3       /* cLoops.blockDBSCAN.blockDBSCAN */

CLOOPS_API const char* cloops_last_error(void);
CLOOPS_API const char* cloops_version(void);
/* number of kernels launched by this library in this process so far (bench.py's gpu_launches) */
CLOOPS_API int64_t cloops_kernel_launches(void);

/* ---- stage timing (CUDA events on the caller's stream) ----------------------------------------
 * With profiling on, every library call records per-stage events; after the call
 * cloops_stage_count()/cloops_stage_name(i)/cloops_stage_ms(i) describe the stages of the LAST
 * call on this thread (the call synchronises the stream when profiling is on). */
CLOOPS_API void cloops_set_profiling(int on);
CLOOPS_API int cloops_stage_count(void);
CLOOPS_API const char* cloops_stage_name(int i);
CLOOPS_API float cloops_stage_ms(int i);
/* Scratch memory of a call comes from a workspace per (host thread, device, stream) that is kept between calls (no allocator
 * call in steady state; CLOOPS_ARENA=0 in the environment: one cudaMallocAsync per array instead).  This returns every
 * workspace block of the process to the driver; no call may be in flight on any thread. */
CLOOPS_API int cloops_workspace_release(void);

/* ---- clustering -------------------------------------------------------------------------------
 * d_x, d_y: int32[n] PET anchor coordinates in ROW order (row = position in the reference's `mat`).
 * cut > 0 drops rows with y - x < cut before clustering (pipe.py:59-63); they get label -1.
 * d_labels: int32[n], row order; cluster ids EXACTLY as the reference numbers them, -1 = the point
 * is absent from the reference's `labels` dict.
 * h_info (may be NULL): int64[8] = {n_active, n_clusters (max id + 1), n_components, n_core,
 *                                   n_dead (v2 released clusters), n_strips, key_bits, n_labelled}. */
CLOOPS_API int cloops_dbscan(const int32_t* d_x, const int32_t* d_y, int64_t n, int32_t eps, int32_t minPts,
                  int32_t cut, int32_t variant, int32_t* d_labels, int64_t* h_info, void* stream);

/* Host-buffer form of the reference boundary `DBSCAN(mat, eps, minPts).labels`:
 * h_mat int64[n,3] rows [pointId, X, Y] (pipe.py:70); h_labels int64[n] row order, -1 = absent.
 * Copies in, clusters on the GPU, copies out. */
CLOOPS_API int cloops_dbscan_host(const int64_t* h_mat, int64_t n, int32_t eps, int32_t minPts, int32_t variant,
                       int64_t* h_labels, int64_t* h_info);

/* Region query alone: d_counts[row] = min(cap, #{q : |dX|+|dY| <= eps}) counting the point itself
 * (cDBSCAN.py:196-204).  cap <= 0 means no saturation. Rows removed by `cut` get 0. */
CLOOPS_API int cloops_neighbour_counts(const int32_t* d_x, const int32_t* d_y, int64_t n, int32_t eps, int32_t cap,
                            int32_t cut, int32_t* d_counts, void* stream);

/* ---- resident index, for sweeps and for timing the region-query kernel in isolation -------------
 * The index is the HBM layout of one chromosome for one eps: points sorted by (v-strip, u) as packed
 * 64-bit keys, the row permutation and the dense strip-offset table. */
typedef struct cloops_index cloops_index;
CLOOPS_API int cloops_index_build(const int32_t* d_x, const int32_t* d_y, int64_t n, int32_t eps, int32_t cut,
                       cloops_index** out, void* stream);
CLOOPS_API void cloops_index_free(cloops_index* ix);                 /* frees on the legacy default stream */
CLOOPS_API void cloops_index_release(cloops_index* ix, void* stream);  /* stream-ordered free, no host sync */
CLOOPS_API int64_t cloops_index_n_active(const cloops_index* ix);
/* one launch of the region-query kernel over the index; d_counts_sorted int32[n_active] in index order */
CLOOPS_API int cloops_index_count(cloops_index* ix, int32_t cap, int32_t* d_counts_sorted, void* stream);
/* labels for one minPts over a built index (v1/v2 only).  d_labels (may be NULL): row order, as
 * cloops_dbscan; d_labels_sorted (may be NULL): int32[n_active], the same labels in index order.  Passing
 * d_labels = NULL skips the scatter back to row order for callers that stay in index order. */
CLOOPS_API int cloops_index_dbscan(cloops_index* ix, int32_t minPts, int32_t variant, int32_t* d_labels,
                        int32_t* d_labels_sorted, int64_t* h_info, void* stream);
/* X, Y of the active PETs in index order (int32[n_active] each), decoded from the packed keys.  Index
 * order is spatially coherent, which makes cloops_cluster_summary over (xs, ys, labels_sorted) cheap. */
CLOOPS_API int cloops_index_coords(cloops_index* ix, int32_t* d_xs, int32_t* d_ys, void* stream);

/* ---- cluster -> candidate records (pipe.py:76-109) ---------------------------------------------
 * n_clusters = max id + 1.  d_bbox int32[n_clusters,4] = minX,maxX,minY,maxY ; d_size int32[n_clusters]
 * (0 for unused ids); d_kind uint8[n_clusters]: 0 unused/dropped (zero extent, pipe.py:83-85),
 * 1 inter-ligation (maxX < minY, pipe.py:97), 2 self-ligation; d_row_kind uint8[n] (may be NULL):
 * kind of the row's cluster, 0 for unlabelled rows (membership of dis / dss, pipe.py:106-109). */
CLOOPS_API int cloops_cluster_summary(const int32_t* d_x, const int32_t* d_y, const int32_t* d_labels, int64_t n,
                           int64_t n_clusters, int32_t* d_bbox, int32_t* d_size, uint8_t* d_kind,
                           uint8_t* d_row_kind, void* stream);

/* d_row_kind[i] = d_kind[d_labels[i]] (0 for unlabelled rows): membership of dis / dss (pipe.py:106-109) */
CLOOPS_API int cloops_row_kinds(const int32_t* d_labels, int64_t n, const uint8_t* d_kind, int64_t n_clusters,
                     uint8_t* d_row_kind, void* stream);

/* ---- permuted-local-background range counts (cModel.py:60-143) ---------------------------------
 * A coverage model is the chromosome's PETs sorted once by X and once by Y (the reference's
 * getGenomeCoverage, cModel.py:45-57). */
typedef struct cloops_coverage cloops_coverage;
CLOOPS_API int cloops_coverage_build(const int32_t* d_x, const int32_t* d_y, int64_t n, cloops_coverage** out, void* stream);
CLOOPS_API void cloops_coverage_free(cloops_coverage* cov);
CLOOPS_API void cloops_coverage_release(cloops_coverage* cov, void* stream);  /* stream-ordered free */
/* d_cand int32[m,4] = iva0, iva1, ivb0, ivb1 (already clamped at 0, cModel.py:281-282);
 * d_out int32[m,123] = ra, rb, rab, na[10], nb[10], C[10][10] row-major (i over A windows). */
CLOOPS_API int cloops_range_counts(const cloops_coverage* cov, const int32_t* d_cand, int64_t m, int32_t* d_out, void* stream);
/* only ra, rb, rab (getPETsforRegions, cModel.py:72-80): d_out int32[m,3] */
/* measurement aid: algorithmic work of the range-count kernel for a candidate list: h_work[0] / h_work[1] = PETs
 * with X / with Y inside the hull of each candidate's windows, summed; win = 5 (cloops_range_counts) or 0 (cloops_region_pets) */
CLOOPS_API int cloops_range_work(const cloops_coverage* cov, const int32_t* d_cand, int64_t m, int32_t win, uint64_t* h_work, void* stream);
CLOOPS_API int cloops_region_pets(const cloops_coverage* cov, const int32_t* d_cand, int64_t m, int32_t* d_out, void* stream);

/* ---- one pass of the whole hot path over one chromosome, as ONE call ----------------------------------
 * cluster (pipe.py:52-75) -> candidate records + dis/dss membership (pipe.py:76-109) -> coverage model
 * (cModel.py:45-57) -> range counts of every inter-ligation candidate (cModel.py:118-143, skipped when
 * score = 0).  Results stay in HBM inside the opaque pass object; the coverage build overlaps the
 * clustering on an internal side stream.  cloops_pass_run_host starts from HOST coordinates (int32[n]
 * each, pinned or pageable) and copies them in first: it is the end-to-end entry point. */
typedef struct cloops_pass cloops_pass;
CLOOPS_API int cloops_pass_run(const int32_t* d_x, const int32_t* d_y, int64_t n, int32_t eps, int32_t minPts, int32_t cut,
                    int32_t variant, int32_t score, cloops_pass** out, void* stream);
/* The same pass, additionally adding this chromosome's distance statistics to the accumulators of the current clustering
 * round (replaces the dis / dss lists of cLoops/pipe.py:59-63,106-109 pooled at pipe.py:120-127 and reduced by
 * cLoops/ests.py:36-61): d_hist int32[CLOOPS_ROUND_HIST_BINS + 1] = exact histogram of the positive self-ligation distances
 * (last bin = overflow), d_mom double[CLOOPS_ROUND_MOM] = n, sum, sum of squares of log2|d| for the inter-ligation set (0..2)
 * and for the self-ligation distances BEYOND the histogram only (3..5; the moments of the histogrammed ones follow from the
 * histogram and are added by cloops_round_middle), len(dis), len(dss) (6, 7), chromosomes that contributed (8).  A chromosome
 * without inter-ligation clusters adds nothing (pipe.py:121-122).  The caller zeroes the accumulators per round and sums them
 * over GPUs. */
#define CLOOPS_ROUND_HIST_BINS (1 << 20)
#define CLOOPS_ROUND_MOM 16
CLOOPS_API int cloops_pass_run_stats(const int32_t* d_x, const int32_t* d_y, int64_t n, int32_t eps, int32_t minPts, int32_t cut,
                          int32_t variant, int32_t score, int32_t* d_hist, double* d_mom, cloops_pass** out, void* stream);
/* The same pass for a round that shares eps with earlier rounds: `base` is the index of the chromosome built ONCE for this eps
 * with cut = 0 (cloops_index_build); the rows this round keeps (Y - X >= cut, cLoops/pipe.py:59-63) are a subsequence of it, so
 * the round's index is one stable compaction of the base instead of a sort.  d_hist / d_mom may both be NULL.  The presets
 * -m 3 / -m 4 (cLoops/pipe.py:337-344) cluster every eps with 2-4 minPts values, each with the cut of the round before. */
CLOOPS_API int cloops_pass_run_base(const cloops_index* base, const int32_t* d_x, const int32_t* d_y, int64_t n, int32_t minPts,
                         int32_t cut, int32_t variant, int32_t score, int32_t* d_hist, double* d_mom, cloops_pass** out, void* stream);
/* The two middle order statistics of the histogram (ranks (k-1)/2 and k/2 of the k positive self-ligation distances it
 * holds; -1 if k = 0, CLOOPS_ROUND_HIST_BINS if a rank lies in the overflow bin) and the round's moments on the host: h_mom =
 * d_mom with the self-ligation entries completed from the histogram (h_mom[3] = k, h_mom[4], h_mom[5] += sum over bins of
 * count * log2(bin) and count * log2(bin)^2, fixed summation order): what estIntSelCutFrag (cLoops/ests.py:36-61) needs.
 * Synchronises the stream. */
CLOOPS_API int cloops_round_middle(const int32_t* d_hist, const double* d_mom, int64_t* h_middle, double* h_mom, void* stream);
CLOOPS_API int cloops_pass_run_host(const int32_t* h_x, const int32_t* h_y, int64_t n, int32_t eps, int32_t minPts, int32_t cut,
                         int32_t variant, int32_t score, cloops_pass** out, void* stream);
/* sizes int64[6] = n_members, n_clusters, n_candidates, scored, n_rows, 0 ; h_info int64[8] as cloops_dbscan.
 * Members are the clustered PETs in index order (rows in file order for CLOOPS_BLOCK). */
CLOOPS_API int cloops_pass_sizes(const cloops_pass* p, int64_t* sizes, int64_t* h_info);
/* device views valid until cloops_pass_free.  which: 0 bbox int32[k,4] (minX,maxX,minY,maxY per cluster id),
 * 1 size int32[k], 2 kind u8[k] (0 dropped, 1 inter-ligation, 2 self-ligation), 3 xs int32[n_members],
 * 4 ys int32[n_members], 5 labels int32[n_members], 6 member_kind u8[n_members], 7 cand int32[m,4] (clamped
 * iva0,iva1,ivb0,ivb1 in ascending cluster id), 8 counts int32[m,123] */
CLOOPS_API const void* cloops_pass_device_ptr(const cloops_pass* p, int which);
/* copy results into HOST buffers (each may be NULL) and synchronise the stream once */
CLOOPS_API int cloops_pass_fetch(const cloops_pass* p, int32_t* h_bbox, uint8_t* h_kind, uint8_t* h_member_kind, int32_t* h_counts,
                      void* stream);
/* candidate records for the host (pipe.py:76-102): h_bbox int32[k,4], h_size int32[k], h_kind u8[k]; any may be NULL; one sync */
CLOOPS_API int cloops_pass_fetch_records(const cloops_pass* p, int32_t* h_bbox, int32_t* h_size, uint8_t* h_kind, void* stream);
CLOOPS_API void cloops_pass_free(cloops_pass* p, void* stream);

/* ---- removeDup (cLoops/cModel.py:198-259) for the loops of one chromosome: host C++ --------------------------------
 * a0,a1,b0,b1 int64[n]: the two anchors of each loop in key order; bp double[n]: binomial p; dens double[n]: rab/ra/rb.
 * keep int64[n] receives the surviving loops in the reference's output order (unique loops in key order, then one winner
 * per overlap group in leader order); an entry -(g+1) stands for tie group g, whose eligible members are
 * tie_members[tie_start[g] .. tie_start[g+1]) (several share the maximum): the reference's winner then depends on the sort
 * pandas uses, and the caller resolves it with the reference's own expression.  tie_start int64[n+1], tie_members int64[n]. */
CLOOPS_API int cloops_remove_dup(const int64_t* a0, const int64_t* a1, const int64_t* b0, const int64_t* b1, const double* bp,
                      const double* dens, int64_t n, double bpcut, int64_t* keep, int64_t* n_keep, int64_t* tie_start,
                      int64_t* tie_members, int64_t* n_ties);

/* combineTwice (cLoops/pipe.py:155-174) over all clustering rounds of one chromosome at once: rows int32[n,4] = candidate
 * boxes of every round in round order, round int32[n]; keep u8[n] = 0 iff the same box came from an EARLIER round.  Host C++. */
CLOOPS_API int cloops_combine_rounds(const int32_t* rows, const int32_t* round, int64_t n, uint8_t* keep);

/* ---- BEDPE ingest (cLoops/io.py:132-189 parseRawBedpe2 / :62-129 parseRawBedpe, with the PET arithmetic of :30-59 and the
 * text -> matrix step of txt2jd :192-203): host C++, one reader thread (zlib for *.gz, io.py:148-151) + `threads` tokenizers.
 * paths: the replicate files in command-line order; chroms: the wanted chromosomes (n_chroms = 0: all, io.py:171);
 * cut: initial distance cut-off (io.py:174).  Per chromosome, in order of first appearance, the accepted cis PETs in file
 * order: cA, cB int64 (floor centres, left <= right), opposite u8 (strandA != strandB, the PETs parseRawBedpe collects for
 * estFragSize, io.py:126-127), line_no int64 (0-based over all files).  Lines whose coordinates python's int() may still
 * accept but that are not plain decimal integers are NOT decided here: they are returned verbatim ("odd" lines) with their
 * line number and the caller applies the reference's own expression; bare_cr counts lines holding a carriage return that
 * is not part of "\r\n" (python 3 splits there, python 2 does not): callers re-read such files line by line. */
typedef struct cloops_bedpe cloops_bedpe;
CLOOPS_API int cloops_bedpe_parse(const char* const* paths, int n_paths, const char* const* chroms, int n_chroms, int64_t cut,
                       int threads, cloops_bedpe** out);
CLOOPS_API int64_t cloops_bedpe_lines(const cloops_bedpe* h);        /* lines read: the reference's counter i (io.py:153) */
CLOOPS_API int64_t cloops_bedpe_bare_cr(const cloops_bedpe* h);
CLOOPS_API int cloops_bedpe_n_chroms(const cloops_bedpe* h);
/* name (not NUL-terminated beyond name_len bytes of payload) and PET count of chromosome k */
CLOOPS_API const char* cloops_bedpe_chrom(const cloops_bedpe* h, int k, int64_t* name_len, int64_t* n_pets);
CLOOPS_API int cloops_bedpe_fetch(const cloops_bedpe* h, int k, int64_t* cA, int64_t* cB, uint8_t* opposite, int64_t* line_no);
CLOOPS_API int64_t cloops_bedpe_n_odd(const cloops_bedpe* h);
CLOOPS_API const char* cloops_bedpe_odd(const cloops_bedpe* h, int64_t k, int64_t* line_no, int64_t* len);
CLOOPS_API void cloops_bedpe_free(cloops_bedpe* h);

#ifdef __cplusplus
}
#endif
#endif
