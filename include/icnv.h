/*
 * libicnv — B200 (sm_100a) CNV-inference hot path behind a plain C ABI.
 *
 * The reference (icbi-lab/infercnvpy @ 89aac1e) is pure Python and has no FFI of
 * its own; these entry points are what a ctypes binding inside
 * src/infercnvpy/tl/_infercnv.py would call in place of the numpy code cited on
 * each function (file:line into /root/reference/src/infercnvpy/).  See
 * INTEGRATION.md for the binding stub.
 *
 * Conventions
 *   - every function returns 0 on success, a negative ICNV_E* code otherwise;
 *     icnv_last_error() returns a thread-local human-readable message;
 *   - all data pointers are DEVICE pointers unless the name ends in _host;
 *   - `stream` is a cudaStream_t passed as void* (NULL = default stream);
 *   - nothing here knows about torch, AnnData or NCCL: cross-GPU reduction of
 *     the column sums is done by the caller between icnv_colsum_* and
 *     icnv_plan_set_reference (one all-reduce).
 */
#ifndef ICNV_H
#define ICNV_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ICNV_OK 0
#define ICNV_EINVAL (-1)      /* bad argument */
#define ICNV_ECUDA (-2)       /* CUDA runtime error */
#define ICNV_EUNSUPPORTED (-3) /* shape outside what the kernels cover (message says why) */

typedef struct icnv_plan icnv_plan;

const char* icnv_last_error(void);
int icnv_version(void);

/* ------------------------------------------------------------------ plan ----
 * Gene-axis plan = everything that depends on var/window/step but not on the
 * cells: position-sorted gene order per chromosome, window grid, output
 * offsets.  Replaces the per-chunk pandas work of
 * tl/_infercnv.py:327-351 (_running_mean_by_chromosome /
 * _running_mean_for_chromosome) and the window bookkeeping of :205-218.
 *
 *   n_genes    columns of the expression matrix (before any masking)
 *   n_seg      chromosomes that take part, in output order (natural sort)
 *   gene_idx_host[seg_off_host[n_seg]]  original column of the p-th gene in
 *              position order, segments concatenated (integer permutation
 *              computed by the host with the reference's own pandas call so
 *              tie order matches)
 *   seg_off_host[n_seg+1]
 *   window, step   tl/_infercnv.py:25-26
 */
int icnv_plan_create(int device, int32_t n_genes, int32_t n_seg, const int32_t* gene_idx_host,
                     const int32_t* seg_off_host, int32_t window, int32_t step, icnv_plan** out);
void icnv_plan_destroy(icnv_plan* plan);

/* Output width K and per-segment first output column (== chr_pos values of
 * tl/_infercnv.py:335-337).  out_off_host has n_seg+1 entries. */
int icnv_plan_out_width(const icnv_plan* plan, int64_t* K);
int icnv_plan_out_offsets(const icnv_plan* plan, int64_t* out_off_host);
/* Which smoothing kernel the plan runs (with the reference dtype installed by icnv_plan_set_reference):
 *   0 = grouped kernel with compile-time (window, step) in {(100,10), (250,10)}; window 100 stages cell rows in pairs
 *       when two rows fit in shared memory (dense input), one row otherwise and for CSR input;
 *   1 = grouped kernel with runtime weights (any step that divides the window);
 *   2 = direct-form kernel staged in parts: any (window, step), float64 centring, any gene-axis length. */
int icnv_plan_kernel_tier(const icnv_plan* plan);
/* Cell rows the smoothing kernel chosen for dense input stages per CTA iteration (1, or 2 = row pairs: every gather-table
 * entry is read once for two rows); < 0 on error. */
int icnv_plan_rows_per_iteration(const icnv_plan* plan);

/* ------------------------------------------------------- reference profile --
 * Column sums for the reference profile: tl/_infercnv.py:385 (all cells) and
 * :400 (one mean per category).  Accumulates in float64.
 *   row_cat   [n_rows] category of each row, -1 = not a reference cell;
 *             NULL = every row belongs to category 0
 *   sums      [n_cat, G] float64, OVERWRITTEN with this shard's sums
 *   counts    [n_cat] int64, OVERWRITTEN with this shard's row counts
 */
int icnv_colsum_dense_f32(const float* X, int64_t n_rows, int64_t ldx, int32_t G, const int32_t* row_cat,
                          int32_t n_cat, double* sums, int64_t* counts, void* stream);
int icnv_colsum_csr_f32(const int64_t* indptr, const int32_t* indices, const float* data, int64_t n_rows,
                        int32_t G, const int32_t* row_cat, int32_t n_cat, double* sums, int64_t* counts,
                        void* stream);
/* mean = sums / counts; written as float32 (the dtype numpy gives for a float32
 * matrix) or, with out_is_f64 != 0, as float64 (integer / float64 matrices). */
int icnv_mean_from_sums(const double* sums, const int64_t* counts, int32_t n_cat, int32_t G, void* ref_out,
                        int32_t out_is_f64, void* stream);
/* indptr[0] = 0, indptr[i+1] = sum(row_nnz[0..i]) (int64); n_rows + 1 entries. */
int icnv_nnz_to_indptr(const int32_t* row_nnz, int64_t n_rows, int64_t* indptr, void* stream);

/* Install the reference profile [n_cat, n_genes] (row-major, device) into the
 * plan: builds the position-sorted per-gene bounds used for centring,
 * tl/_infercnv.py:422-432 (n_cat == 1: x - ref; n_cat > 1: bounded difference
 * against min/max over categories).  ref_is_f64 != 0 selects float64 centring
 * (numpy promotes float32 - float64). */
int icnv_plan_set_reference(icnv_plan* plan, const void* ref, int32_t n_cat, int32_t ref_is_f64, void* stream);

/* ------------------------------------------------------------- smoothing ----
 * Steps 1-3 of tl/_infercnv.py:411-440 for n_rows cells: centre (:422-432), clip to +-lfc_clip (:436),
 * per-chromosome pyramid running mean decimated by `step` (:179-244, :301-356).  Results do not depend on which
 * kernel tier runs, on the launch geometry or on how the rows are split into calls (bit-reproducible).
 * icnv_smooth_csr_f32 densifies every row on load like the reference (:423); tier 2 takes dense input only.
 *   tmp   [n_rows, ld_tmp >= icnv_plan_tmp_width()] float64: the smoothed rows in the kernel's warp-tile
 *         column order (every store a full line); icnv_center_rows turns them into the natural matrix.
 */
int icnv_plan_tmp_width(const icnv_plan* plan, int64_t* ld_tmp);
int icnv_smooth_dense_f32(icnv_plan* plan, const float* X, int64_t n_rows, int64_t ldx, double lfc_clip,
                          double* tmp, int64_t ld_tmp, void* stream);
int icnv_smooth_csr_f32(icnv_plan* plan, const int64_t* indptr, const int32_t* indices, const float* data,
                        int64_t n_rows, double lfc_clip, double* tmp, int64_t ld_tmp, void* stream);

/* Step 4, tl/_infercnv.py:442: subtract the exact row median (np.median: mean of the two middle values
 * for even K).  tmp -> out [n_rows, ldo >= K] float32 (out_is_f64 == 0) or float64, natural column order;
 * row_stats [n_rows, 2] float64 = sum and sum of squares of the centred row (inputs of :450). */
int icnv_center_rows(icnv_plan* plan, const double* tmp, int64_t n_rows, int64_t ld_tmp, void* out,
                     int32_t out_is_f64, int64_t ldo, double* row_stats, void* stream);

/* Step 5, tl/_infercnv.py:449-451.  Rows are cut into consecutive chunks of chunk_rows (the reference's
 * `chunksize`, :123); thr[c] = dyn_thr * population-std over every element of chunk c.  thr has
 * ceil(n_rows / chunk_rows) entries. */
int icnv_chunk_threshold(const double* row_stats, int64_t n_rows, int64_t K, int64_t chunk_rows, double dyn_thr,
                         double* thr, void* stream);
/* Zero |v| < thr[chunk of row] in place (strict, :451); also emits per row sum|v| and the number of
 * non-zeros (inputs of cnv_score, tl/_scores.py:66, and of the CSR conversion, tl/_infercnv.py:455).
 * thr == NULL: no zeroing, statistics only. */
int icnv_apply_threshold(void* out, int32_t out_is_f64, int64_t n_rows, int64_t K, int64_t ldo,
                         int64_t chunk_rows, const double* thr, double* row_abs_sum, int32_t* row_nnz,
                         void* stream);

/* The same filter WITHOUT rewriting the matrix (the CSR path of tl.infercnv never needs the filtered dense block):
 * icnv_filter_count only counts -- per row the number and the sum|v| of the values that survive |v| < thr (strict,
 * :451; zeros never count) -- and icnv_filter_to_csr applies the same predicate again while compacting every row into
 * (indices, data) (:455; indptr = icnv_nnz_to_indptr of the counts).  data is float32 or float64 (data_is_f64; a float64
 * matrix needs float64 data).  thr == NULL: every non-zero survives.  Against icnv_apply_threshold + icnv_dense_to_csr
 * this saves one write and one read of the [n_rows, K] matrix. */
int icnv_filter_count(const void* out, int32_t out_is_f64, int64_t n_rows, int64_t K, int64_t ldo, int64_t chunk_rows,
                      const double* thr, double* row_abs_sum, int32_t* row_nnz, void* stream);
int icnv_filter_to_csr(const void* out, int32_t out_is_f64, int64_t n_rows, int64_t K, int64_t ldo, int64_t chunk_rows,
                       const double* thr, const int64_t* indptr, int32_t* indices, void* data, int32_t data_is_f64,
                       void* stream);

/* Per-gene layer of calculate_gene_values=True, tl/_infercnv.py:141-151, :214-223, :238-242, :247-291, :443-444, :452-453.
 * From the smoothed rows in `tmp` (icnv_smooth_*): value(gene) = np.mean (numpy's pairwise order) of the kept windows that
 * contain the gene (:278-287; the flat mean on chromosomes not longer than the window, :240), minus the median of the
 * row's covered genes (:444), zeroed where |v| < thr[chunk of row] (the WINDOW matrix's threshold, :453; thr == NULL:
 * no filter).  gene_out [n_rows, ldg >= n_genes] float64 in the matrix's own column order; genes no kept window
 * covers and genes outside the plan's chromosomes are NaN (:146).  icnv_plan_gene_coverage: number of non-NaN columns. */
int icnv_gene_values(icnv_plan* plan, const double* tmp, int64_t n_rows, int64_t ld_tmp, int64_t chunk_rows,
                     const double* thr, double* gene_out, int64_t ldg, void* stream);
int icnv_plan_gene_coverage(const icnv_plan* plan, int32_t* n_covered);

/* Dense [n_rows, K] -> CSR (tl/_infercnv.py:455).  indptr [n_rows+1] int64 must
 * already hold the exclusive prefix sum of row_nnz; indices int32; data float32
 * or float64 following out_is_f64. */
int icnv_dense_to_csr(const void* out, int32_t out_is_f64, int64_t n_rows, int64_t K, int64_t ldo,
                      const int64_t* indptr, int32_t* indices, void* data, void* stream);

/* ------------------------------------------------------------- cnv_score ----
 * tl/_scores.py:65-68: per label mean(abs(X_cnv[rows of label, :])).
 *   row_abs_sum [n_rows] float64;  labels [n_rows] int32 in [0, n_labels)
 *   label_sum   [n_labels] float64, label_rows [n_labels] int64 (OVERWRITTEN;
 *   all-reduce both across shards, then score = label_sum / (label_rows * K)). */
int icnv_rowabs_csr(const int64_t* indptr, const void* data, int32_t data_is_f64, int64_t n_rows,
                    double* row_abs_sum, void* stream);
int icnv_rowabs_dense(const void* X, int32_t is_f64, int64_t n_rows, int64_t K, int64_t ld, double* row_abs_sum,
                      void* stream);
int icnv_label_sums(const double* row_abs_sum, const int32_t* labels, int64_t n_rows, int32_t n_labels,
                    double* label_sum, int64_t* label_rows, void* stream);

/* ------------------------------------------------------------ ITH scores ----
 * tl/_scores.py:77-221 (ithgex :130-141, ithcna :203-214): np.corrcoef over the rows (cells) of ONE group.
 *   X     [n_rows, ld >= K] float64 row-major, the group's dense block
 *   corr  [n_rows, ldc >= n_rows] float64, OVERWRITTEN: Pearson correlation of every pair of rows, clipped to
 *         [-1, 1] like np.corrcoef; rows without variance give NaN
 *   work  [2 * n_rows] float64 scratch (row means, inverse norms)
 * The score is the inter-quartile range of ALL n_rows^2 entries (np.percentile, linear interpolation). */
int icnv_row_corrcoef_f64(const double* X, int64_t n_rows, int64_t ld, int32_t K, double* corr, int64_t ldc, double* work,
                          void* stream);

/* --------------------------------------------- pca / neighbors / leiden ----
 * The reference hands these steps to scanpy (tl/__init__.py:13-75, pp/__init__.py:8-43 ->
 * scikit-learn TruncatedSVD(arpack), exact/approximate kNN + umap-learn fuzzy_simplicial_set,
 * leidenalg); the entry points below are the device pieces our wrappers are built from.
 * Parity is unpinned (no reference test asserts anything here). */
/* CSR (float32 / float64 values) -> dense float32 [n_rows, ld] */
int icnv_csr_to_dense_f32(const int64_t* indptr, const int32_t* indices, const void* data, int32_t data_is_f64,
                          int64_t n_rows, int32_t K, float* dense, int64_t ld, void* stream);
/* C [K, K] float64 = X^T X of this row shard (OVERWRITTEN; all-reduce across shards) */
int icnv_gram_f32(const float* X, int64_t n_rows, int64_t ld, int32_t K, double* C, void* stream);
/* Y [n_rows, n_comp] float32 = (X - mu) V;  V [K, n_comp] float64 row-major, mu [K] float64 or NULL */
int icnv_project_f32(const float* X, int64_t n_rows, int64_t ld, int32_t K, const double* V, int32_t n_comp,
                     const double* mu, float* Y, void* stream);
/* Exact euclidean kNN of rows [q0, q0+nq) of P [n_all, d <= 64] float32 (row pitch ld) against all rows, as a tensor-core
 * distance GEMM (tcgen05.mma kind::tf32, 3xTF32 split, accumulators in TMEM) + exact fp64 re-rank of k + slack candidates
 * per query (csrc/icnv_knn.cu).  k <= 20.  knn_idx / knn_d2 are [nq, out_ld]: the k nearest
 * by ascending squared distance, ties by index (column 0 is the query itself).  workspace: icnv_knn_workspace_bytes. */
int64_t icnv_knn_workspace_bytes(int64_t n_all, int64_t nq);
int icnv_knn_f32(const float* P, int64_t n_all, int32_t d, int64_t ld, int64_t q0, int64_t nq, int32_t k, int32_t* knn_idx,
                 float* knn_d2, int32_t out_ld, void* workspace, void* stream);
/* umap-learn smooth_knn_dist + membership strengths per row; dist/idx/vals are [n, k] */
int icnv_fuzzy_rows(const float* dist, const int32_t* idx, int64_t n, int32_t k, int64_t row0, float mean_all,
                    float* vals, float* sigma, float* rho, void* stream);
/* Weighted degree of a CSR graph (both directions stored), and ONE synchronous sweep of the Leiden scheme on it
 * (csrc/icnv_graph.cu) with the RB-configuration quality (gamma * k_i * K_c / 2m).  bound == NULL: local moving (nodes
 * move to the best neighbouring community of `comm`); bound != NULL: refinement — only singletons of `comm` that are well
 * connected to their community `bound` move, and only to sub-communities of the same bound with a positive gain.
 * Half of the nodes (hash of node and sweep) may move per sweep.  work: icnv_community_sweep_work_bytes(n) bytes of
 * scratch; stats: int32[3] = {nodes moved, nodes handled by the hash-table path, table overflow flag (then the sweep is
 * incomplete: treat as ICNV_EUNSUPPORTED)}.  Any node degree is handled exactly. */
int icnv_weighted_degree(const int64_t* indptr, const float* w, int64_t n, double* kdeg, void* stream);
int64_t icnv_community_sweep_work_bytes(int64_t n);
int icnv_community_sweep(const int64_t* indptr, const int32_t* indices, const float* w, const double* kdeg, const int32_t* comm,
                         const int32_t* bound, int64_t n, double two_m, double gamma, int32_t sweep, void* work, int32_t* comm_new,
                         int32_t* stats, void* stream);

/* Launch geometry of the smoothing kernel chosen for this plan (for the bench
 * and the ncu notes): CTAs per SM, threads, dynamic shared memory bytes. */
int icnv_plan_launch_info(icnv_plan* plan, int32_t* ctas_per_sm, int32_t* threads, int32_t* smem_bytes,
                          int32_t* n_sm);

/* Average number of shared-memory wavefronts one warp-level gather of the smoothing kernel costs with this plan's
 * gather schedule (1.0 = conflict-free; csrc/icnv_schedule.cu).  0 when the plan has no grouped layout. */
int icnv_plan_gather_cost(const icnv_plan* plan, double* wavefronts_per_gather);

/* Host-only (no device needed): the gather schedule of csrc/icnv_schedule.cu on a caller-supplied group table
 * gcol [n_groups, gs] (matrix column of every group element, -1 = pad).  Fills slot_group_out [nsets*32] (group of every
 * lane slot, -1 = none) and order_out [nsets*32*gs] (element read at every step); returns wavefronts per gather, < 0 if
 * nsets*32 slots cannot hold the groups.  Used by the CPU tests. */
double icnv_host_schedule_gathers(const int32_t* gcol, int32_t n_groups, int32_t gs, int32_t n_genes, int32_t nsets,
                                  int32_t permute, int32_t* slot_group_out, uint8_t* order_out);

/* Developer aid (tools/timeline.py): when dev_buf != NULL the smoothing kernel writes clock64 stamps
 * [grid][rows_per_cta][16] for the first rows_per_cta rows of every CTA.  NULL switches it off. */
int icnv_debug_set_timeline(long long* dev_buf, int rows_per_cta);

/* ------------------------------------------------------------- embeddings ----
 * cnv.tl.umap / cnv.tl.tsne (tl/__init__.py:78-144: wrappers around scanpy.tl.umap / scanpy.tl.tsne; parity unpinned).
 * icnv_umap_epochs: epochs [epoch0, epoch0 + n_run) of umap-learn's optimize_layout_euclidean on a 2-D embedding
 *   emb [n_vertices, 2] float32 (updated in place).  head / tail [n_edges] int32: the directed edges of the symmetric
 *   graph; epochs_per_sample [n_edges] = max weight / weight; next_sample / next_negative [n_edges] are the per-edge
 *   schedules (initialise to epochs_per_sample and epochs_per_sample / neg_rate; carried between calls).
 * icnv_tsne_affinities: joint probabilities P [n, n] float32 of exact t-SNE from X [n, d <= 64] (per-row perplexity
 *   search, symmetrised, floor 1e-12).
 * icnv_tsne_iterations: n_iter gradient-descent steps (scikit-learn's gains / momentum rule) on Y [n, 2]; vel, gains
 *   [n, 2] carried between calls; work: icnv_tsne_work_floats(n) floats. */
int icnv_umap_epochs(const int32_t* head, const int32_t* tail, int64_t n_edges, float* emb, int32_t n_vertices,
                     const float* epochs_per_sample, float* next_sample, float* next_negative, float a, float b, float gamma,
                     float alpha0, int32_t n_epochs, int32_t epoch0, int32_t n_run, int32_t neg_rate, uint32_t seed, void* stream);
int icnv_tsne_affinities(const float* X, int32_t n, int32_t d, int64_t ld, float perplexity, float* P, void* stream);
int64_t icnv_tsne_work_floats(int32_t n);
int icnv_tsne_iterations(const float* P, float* Y, float* vel, float* gains, float* work, int32_t n, int32_t n_iter,
                         float exaggeration, float momentum, float learning_rate, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* ICNV_H */
