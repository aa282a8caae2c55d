// Shared declarations for libicnv (sm_100a).  See include/icnv.h for the C ABI.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>
#include <type_traits>
#include <vector>

namespace icnv {

constexpr int NT = 512;         // threads per CTA of the smoothing kernel
constexpr int NW = NT / 32;     // warps per CTA
// threads per CTA of the smoothing kernel that stages `rows` cell rows per iteration
// the double-buffered single-row kernel has no hand-over barriers: extra warps are pure gather capacity (ICNV_DBUF_THREADS)
#ifndef ICNV_DBUF_THREADS
#define ICNV_DBUF_THREADS 768
#endif
__host__ __device__ constexpr int smooth_threads(int rows, bool dbuf = false) { return dbuf ? ICNV_DBUF_THREADS : NT + 0 * rows; }
// groups per lane of one phase-2 work unit (table layout [unit][step][lane][u < width]) for a kernel staging `rows` rows
// Row pairs of a window whose pyramid peak falls inside a group (third partial sum per group) keep half units: 12 fp64
// accumulators per lane instead of 16 + 8.
#define ICNV_UNIT_WIDTH(rows, peak_group) (((rows) == 2 && !(peak_group)) ? 4 : 2)
// gather walk of the templated kernels: 0 = conflict-aware permuted walk everywhere (entries carry j / m_j, decoded per
// entry), 1 = natural walk for windows with a peak group (window 250), 2 = natural walk everywhere.
// Same-box A/B, 100k x 20k (gpurun_out/r2b/ab_walk.log): window 250 single-row kernel 2.90 ms -> 2.51 ms with the natural
// walk (the decode of j and m_j costs 8 of ~21 instructions per gene when nothing amortises it), window 100 row pairs
// 1.865 ms permuted vs 1.90 ms natural -> default 1.
#ifndef ICNV_NATURAL_WALK
#define ICNV_NATURAL_WALK 1
#endif
#ifndef ICNV_LOUT
#define ICNV_LOUT 9
#endif
constexpr int LOUT = ICNV_LOUT;         // consecutive outputs owned by one task (9*16 B stride -> conflict-free LDS.128)
constexpr int CAND_CAP = 32;    // median: size of the final exact candidate set
constexpr int PAD_GROUPS = 12;  // slack groups after the last one (sliding window over-read, >= LOUT)

// pyramid weight j of an n-wide window: min(j + 1, n - j)  (tl/_infercnv.py:206-207)
__host__ __device__ constexpr int pyr(int n, int j) { return (j + 1) < (n - j) ? (j + 1) : (n - j); }

// One task = up to LOUT consecutive outputs of one chromosome (kind 0), or the single
// flat-mean output of a chromosome not longer than the window (kind 1).
//   kind 0: x = first group (tiers 0/1) or first sorted gene (tier 2); y = first output column;
//           z = number of outputs (1..LOUT)
//   kind 1: x as above; y = output column; z = number of groups (tiers 0/1) or genes (tier 2);
//           w >> 8 = index into flat_inv
struct Task {
    int32_t x, y, z, w;
};

struct SmoothParams {
    // ---- input matrix: dense (X != null) or CSR
    const float* X;
    int64_t ldx;
    const int64_t* indptr;
    const int32_t* indices;
    const float* data;
    int64_t n_rows;
    int32_t G;        // columns of X
    int32_t Gpad;     // staged floats (>= G+1, multiple of 4); slot G is the zero pad
    int32_t use_tma;  // row base and pitch 16-byte aligned -> cp.async.bulk
    int32_t split_rows;   // row pairs: give every staged row its own group warps in phase 3
    int32_t l2_prefetch;  // prefetch the next iteration's rows into L2 while this one is gathered
    // ---- grouped tiers (0/1)
    int32_t gs;       // genes per group (== step)
    int32_t NG;       // groups
    int32_t NGpad;    // multiple of 4
    int32_t NQ;       // groups per window
    int32_t qstar;    // group-in-window holding the pyramid peak (non-linear weights), -1 if none
    // Per-gene tables in the order the kernel walks them: [work unit of 32 lanes x 2 groups][step t < gs][lane][u < 2];
    // entry (unit, t, lane, u) belongs to the element of group grp_w[(unit*32 + lane)*2 + u] that the lane reads at
    // step t (host-optimised assignment and walk order that minimise shared-memory bank collisions of the gathers,
    // icnv_schedule.cu).  Two consecutive units form one warp-block of 32 quads (the slot sets 4*wb .. 4*wb+3).
    const int32_t* grp_w;   // [unit][lane][u] group index (NGpad = unused slot)
    const uint32_t* off_w;  // shared-window BYTE address of the gene inside the staged raw row (zero pad slot = gene G)
    uint32_t raw_base;      // shared-window address of raw[0] baked into off_w (checked by the kernel)
    const float* lo_w;      // reference lower bound (== ref when one category)
    const float* hi_w;      // upper bound (only read when BOUNDED)
    const double* alpha;    // [NQ] weight of A_g = sum_j x
    const double* beta;     // [NQ] weight of B_g = sum_j j*x
    const double* cw;       // [gs] weights inside the peak group (C_g = sum_j cw_j x)
    // Row pairs with a peak group: the third partial sum C'_g = sum_j (cw_j - cw_0) x_j of both staged rows does not fit
    // in shared memory next to 2 x (row + A/B partials); it goes through a per-CTA global scratch that stays in L2,
    // double-buffered by iteration parity: [grid][2 parities][2 rows][NGpad + PAD_GROUPS]
    double* c_scratch;
    // ---- common
    double clip;
    float clipf;
    double inv_sumw;
    const double* flat_inv; // 1 / (genes of flat segment)
    const Task* tasks;
    int32_t n_tasks;
    int32_t K;
    // smoothed rows [n_rows, ldo] fp64 in warp-tile order: value i of task t sits at (t/32)*32*LOUT + i*32 + t%32
    // (+inf in unused slots), so every store instruction of a warp writes 256 contiguous bytes;
    // icnv_center_rows un-permutes while it centres
    double* out;
    int64_t ldo;
    // optional developer timeline: [grid][dbg_rows][16] clock64 stamps (nullptr = off)
    long long* dbg;
    int32_t dbg_rows;
};

// general direct-form smoothing (icnv_direct.cu)
struct DirectParams {
    const float* X;
    int64_t ldx, n_rows;
    const int32_t* idx_lin;  // [n_sorted] matrix column of the s-th position-sorted gene
    const void* lo_lin;      // [n_sorted] float or double (C64) lower bound (== reference when one category)
    const void* hi_lin;
    const double* wdir;      // [window] pyramid weights
    int32_t window, step;
    const Task* tasks;       // direct layout: x = first sorted gene
    int32_t n_tasks;
    const int4* parts;       // (first sorted gene, end, first tile, end tile): what is staged together
    int32_t n_parts;
    double clip;
    float clipf;
    double inv_sumw;
    const double* flat_inv;
    double* out;             // [n_rows, ldo] fp64, warp-tile order + tile moments (same as SmoothParams::out)
    int64_t ldo;
};

// per-gene layer (icnv_genevals.cu)
struct GeneValParams {
    const double* tmp;       // [n_rows, ld] smoothing kernel output (tile order)
    int64_t n_rows, ld;
    const int32_t* kaddr;    // [K] position of output column k inside a tmp row
    int32_t K;
    const int32_t* first;    // [n_cov] first window column of the covered gene (position order)
    const int32_t* cnt;      // [n_cov] number of windows covering it
    int32_t n_cov;
    const int32_t* inv;      // [G] natural column -> covered index, -1 = NaN
    int32_t G;
    int64_t chunk_rows;
    const double* thr;       // [n_chunks] or nullptr (dynamic_threshold=None)
    double* scratch;         // [grid, n_cov] when the means do not fit in shared memory, else nullptr
    double* out;             // [n_rows, ldo] float64
    int64_t ldo;
    int32_t k_in_smem, v_in_smem;
    int32_t hist_off;        // byte offset of the histogram-selection work area in dynamic shared memory, -1 = none
};

// ---------------------------------------------------------------- PTX helpers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(unsigned long long* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
// 1-D bulk async copy global -> shared (TMA engine, SASS UBLKCP), completes on an mbarrier.
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, unsigned long long* bar,
                                         uint64_t policy) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::
            "r"(smem_u32(dst)),
        "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
        : "memory");
}
// same without an L2 cache hint
__device__ __forceinline__ void bulk_g2s_plain(void* dst, const void* src, uint32_t bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ float ldg_stream_f32(const float* p) {
    float v;
    asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p));
    return v;
}
// named barriers: `count` threads (a multiple of 32) take part; arrive does not block
__device__ __forceinline__ void named_bar_sync(int id, int count) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory");
}
__device__ __forceinline__ void named_bar_arrive(int id, int count) {
    asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(count) : "memory");
}
// Ask the TMA engine to pull [src, src + bytes) into L2 (no shared-memory destination, no completion to wait for).
__device__ __forceinline__ void bulk_prefetch_l2(const void* src, uint32_t bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ float4 ldg_nc_f4(const float* p) {
    float4 v;
    asm volatile("ld.global.nc.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}
__device__ __forceinline__ uint4 ldg_nc_u4(const void* p) {
    uint4 v;
    asm volatile("ld.global.nc.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}
__device__ __forceinline__ float2 ldg_nc_f2(const float* p) {
    float2 v;
    asm volatile("ld.global.nc.v2.f32 {%0,%1}, [%2];" : "=f"(v.x), "=f"(v.y) : "l"(p));
    return v;
}
__device__ __forceinline__ uint2 ldg_nc_u2(const void* p) {
    uint2 v;
    asm volatile("ld.global.nc.v2.u32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p));
    return v;
}
__device__ __forceinline__ float4 ldg_stream_f4(const float* p) {
    float4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                 : "l"(p));
    return v;
}

// ---------------------------------------------------------------- error plumbing (host)
void set_error(const std::string& msg);
int cuda_fail(cudaError_t e, const char* what);
#define ICNV_CUDA(expr)                                          \
    do {                                                         \
        cudaError_t _e = (expr);                                 \
        if (_e != cudaSuccess) return ::icnv::cuda_fail(_e, #expr); \
    } while (0)

// Launchers implemented in icnv_smooth.cu
int smooth_launch(int tier, int nwin, int gs, bool bounded, bool c64, int tpt, int rows, bool dbuf, const SmoothParams& p, int grid,
                  size_t smem, cudaStream_t stream);
int smooth_occupancy(int tier, int nwin, int gs, bool bounded, bool c64, int tpt, int rows, bool dbuf, size_t smem, int* ctas_per_sm);

int direct_launch(const DirectParams& p, bool bounded, bool c64, int grid, size_t smem, cudaStream_t st);
int center_wide_launch(const double* tmp, int64_t n_rows, int64_t ld, const int32_t* kaddr, int K, void* out, bool f64, int64_t ldo,
                       double* row_stats, int n_sm, cudaStream_t st);
// icnv_schedule.cu (host only)
double schedule_gathers(const std::vector<int32_t>& gcol, int NG, int gs, int n_genes, int nsets, bool permute,
                        std::vector<int32_t>& slot_group, std::vector<uint8_t>& order);
int genevals_launch(const GeneValParams& p, int grid, size_t smem, cudaStream_t st);
// icnv_sparse.cu: sparse-aware smoothing of CSR input + deterministic CSR column sums
bool sparse_supported(int nwin, int gs);
size_t sparse_smem_bytes(int DP, int NGpad, bool peak_group);
int sparse_zrow_launch(const int32_t* slot_col, int n_slots, const int4* col_tab, float clipf, bool bounded, float* zrow, cudaStream_t st);
int sparse_col_table_launch(const void* ref, bool ref_f64, int n_cat, int G, const int32_t* col_slot, int4* col_tab, cudaStream_t st);
int sparse_smooth_launch(int nwin, int gs, bool bounded, const int64_t* indptr, const int32_t* indices, const float* data, int64_t n_rows,
                         const int4* col_tab, const float* zrow, int DP, int NG, int NGpad, double inv_sumw, const double* flat_inv,
                         const Task* tasks, int n_tasks, float clipf, double* out, int64_t ldo, int n_sm, cudaStream_t st);
// icnv_sparse_delta.cu: CSR smoothing with work proportional to the stored entries (one category)
bool sparse_delta_supported(int nwin, int gs, int NGpad, int n_tasks);
size_t sparse_delta_smem_bytes(int G, int NGpad, bool peak_group);
int sparse_delta_tables_launch(const int4* col_tab, int G, int gs, uint16_t* gj, float* ref, cudaStream_t st);
int sparse_delta_launch(int nwin, int gs, const int64_t* indptr, const int32_t* indices, const float* data, int64_t n_rows,
                        const uint16_t* gj, const float* ref, int G, int NG, int NGpad, const double* base, double inv_sumw,
                        const double* flat_inv, const Task* tasks, int n_tasks, float clipf, double* out, int64_t ldo, int n_sm,
                        cudaStream_t st);
int sparse_colsum_splits(int64_t n_rows, int n_sm);
int sparse_colsum_launch(const int64_t* indptr, const int32_t* indices, const float* data, int64_t n_rows, int G, const int32_t* row_cat,
                         int n_cat, double* partial, int n_split, cudaStream_t st);

}  // namespace icnv
