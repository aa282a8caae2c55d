// C ABI of libicnv (include/icnv.h) and the gene-axis plan.
#include <algorithm>
#include <memory>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <numeric>

#include "../../include/icnv.h"
#include "icnv_common.cuh"

namespace icnv {

static thread_local std::string g_err;
void set_error(const std::string& msg) { g_err = msg; }
int cuda_fail(cudaError_t e, const char* what) {
    g_err = std::string("CUDA error: ") + cudaGetErrorString(e) + " at " + what;
    return ICNV_ECUDA;
}

// aux launchers (icnv_aux.cu)
int aux_colsum_dense(const float*, int64_t, int64_t, int, const int32_t*, int, double*, int64_t*, double*, int, cudaStream_t);
int aux_colsum_dense_slots(int* slots);
int aux_colsum_csr(const int64_t*, const int32_t*, const float*, int64_t, int, const int32_t*, int, double*, int64_t*, double*, int, cudaStream_t);
int aux_mean_from_sums(const double*, const int64_t*, int, int, void*, bool, cudaStream_t);
int aux_nnz_to_indptr(const int32_t*, int64_t, int64_t*, cudaStream_t);
int aux_build_bounds(const void*, bool, int, int, const int32_t*, int64_t, void*, void*, bool, cudaStream_t);
int aux_chunk_threshold(const double*, int64_t, int64_t, int64_t, double, double*, cudaStream_t);
int aux_center_rows(const double*, int64_t, int64_t, const Task*, int, int, void*, bool, int64_t, double*, cudaStream_t);
int aux_apply_threshold(void*, bool, int64_t, int64_t, int64_t, int64_t, const double*, double*, int32_t*, cudaStream_t);
int umap_epochs(const int32_t*, const int32_t*, int64_t, float*, int32_t, const float*, float*, float*, float, float, float, float, int, int,
                int, int, uint32_t, cudaStream_t);
int tsne_affinities(const float*, int, int, int64_t, float, float*, cudaStream_t);
int tsne_iterations(const float*, float*, float*, float*, float*, int, int, float, float, float, cudaStream_t);
int filter_count(const void*, bool, int64_t, int64_t, int64_t, int64_t, const double*, double*, int32_t*, cudaStream_t);
int filter_to_csr(const void*, bool, int64_t, int64_t, int64_t, int64_t, const double*, const int64_t*, int32_t*, void*, bool, cudaStream_t);
int smooth_raw_base(uint32_t* base);
int aux_dense_to_csr(const void*, bool, int64_t, int64_t, int64_t, const int64_t*, int32_t*, void*, cudaStream_t);
int aux_rowabs_csr(const int64_t*, const void*, bool, int64_t, double*, cudaStream_t);
int aux_rowabs_dense(const void*, bool, int64_t, int64_t, int64_t, double*, cudaStream_t);
int aux_label_sums(const double*, const int32_t*, int64_t, int, double*, int64_t*, cudaStream_t);
size_t smooth_scratch_bytes();
int aux_row_corrcoef(const double*, int64_t, int64_t, int, double*, int64_t, double*, cudaStream_t);
int graph_csr_to_dense(const int64_t*, const int32_t*, const void*, bool, int64_t, int, float*, int64_t, cudaStream_t);
int graph_gram(const float*, int64_t, int64_t, int, double*, cudaStream_t);
int graph_project(const float*, int64_t, int64_t, int, const double*, int, const double*, float*, cudaStream_t);
int knn_launch(const float*, int64_t, int, int64_t, int64_t, int64_t, int, int, int32_t*, float*, void*, cudaStream_t);
size_t knn_workspace_bytes(int64_t, int64_t);
int graph_fuzzy_rows(const float*, const int32_t*, int64_t, int, int64_t, float, float*, float*, float*, cudaStream_t);
int graph_weighted_degree(const int64_t*, const float*, int64_t, double*, cudaStream_t);
int graph_community_sweep(const int64_t*, const int32_t*, const float*, const double*, const int32_t*, const int32_t*, int64_t, double,
                          double, int, void*, int32_t*, int32_t*, cudaStream_t);

template <typename T>
struct DevBuf {
    T* ptr = nullptr;
    size_t n = 0;
    int upload(const std::vector<T>& h) {
        n = h.size();
        if (n == 0) return 0;
        ICNV_CUDA(cudaMalloc(&ptr, n * sizeof(T)));
        ICNV_CUDA(cudaMemcpy(ptr, h.data(), n * sizeof(T), cudaMemcpyHostToDevice));
        return 0;
    }
    int alloc(size_t count) {
        n = count;
        if (n == 0) return 0;
        ICNV_CUDA(cudaMalloc(&ptr, n * sizeof(T)));
        return 0;
    }
    void release() {
        if (ptr) cudaFree(ptr);
        ptr = nullptr;
        n = 0;
    }
};

}  // namespace icnv

using namespace icnv;

struct icnv_plan {
    int device = 0;
    int n_sm = 148;
    int32_t G = 0, n_seg = 0, window = 0, step = 0;
    std::vector<int32_t> seg_off, gene_idx;
    std::vector<int64_t> out_off;  // n_seg + 1
    int64_t K = 0;
    double inv_sumw = 1.0;

    // ---- grouped layout (tiers 0/1); valid when group_ok
    bool group_ok = false;
    int base_tier = 2;
    int32_t gs = 0, NG = 0, NGpad = 0, NQ = 0, qstar = -1, Gpad = 0;
    uint32_t raw_base = 0;
    int32_t n_tasks_g = 0;
    // gather tables in the order a kernel with unit width uw walks them ([unit][step][lane][u < uw]).  Set 0 serves the
    // kernel chosen for dense input; when that is the row-pair kernel (uw = 4), set 1 (uw = 2) serves the single-row
    // kernel that CSR input runs (densify-on-load has no TMA to overlap, two CTAs per SM hide its barriers better).
    struct GatherTables {
        int uw = 0;
        DevBuf<uint32_t> off_w;  // shared-window byte address of the gene (| element position << 24 when permuted)
        DevBuf<int32_t> cols_w;  // same layout, column index, pad = -1 (input of the bounds kernel)
        DevBuf<int32_t> grp_w;   // [unit][lane][u] group whose partial sums this slot produces
        DevBuf<float> lo_w, hi_w;
        void release() {
            off_w.release();
            cols_w.release();
            grp_w.release();
            lo_w.release();
            hi_w.release();
        }
    } tab[2];
    DevBuf<double> alpha, beta, cw;
    DevBuf<Task> tasks_g;
    // ---- sparse-aware CSR smoothing (icnv_sparse.cu): slot = group * gs + element in position order
    bool sparse_ok = false;
    int32_t sp_DP = 0;
    DevBuf<int32_t> sp_slot_col, sp_col_slot;  // [DP] column of a slot (-1 pad); [G] slot of a column (-1: takes no part)
    DevBuf<int4> sp_col_tab;                   // [G] {slot, lo, hi, 0}, filled by icnv_plan_set_reference
    DevBuf<float> sp_zrow;                     // [DP] what a zero entry becomes, filled per smoothing call (depends on lfc_clip)
    // delta kernel (icnv_sparse_delta.cu): one category, work proportional to the stored entries
    bool delta_ok = false;
    DevBuf<uint16_t> sp_gj;                    // [G padded to 8] (group << 4) | element per column, 0xFFFF = takes no part
    DevBuf<float> sp_ref;                      // [G] reference value per column
    DevBuf<double> sp_base;                    // one tmp row: the smoothed constant row, filled per smoothing call
    DevBuf<int64_t> sp_indptr0;                // {0, 0}: the empty row that yields sp_base
    DevBuf<double> c_scratch;  // row pairs with a peak group: third partial sums, [n_sm][2 parities][2 rows][NGpad + PAD_GROUPS]

    // ---- direct layout (tier 2); always built
    int32_t n_sorted = 0, n_tasks_d = 0;
    DevBuf<int32_t> idx_lin;
    DevBuf<unsigned char> lo_lin, hi_lin;  // float or double
    DevBuf<double> wdir;
    DevBuf<Task> tasks_d;

    DevBuf<double> flat_inv;
    std::vector<Task> tasks_d_host;
    // parts of the direct kernel (runs of whole task tiles whose genes fit in shared memory), per staging dtype
    DevBuf<int32_t> parts[2];
    int32_t n_parts[2] = {0, 0};
    size_t parts_smem[2] = {0, 0};

    int rows = 1;                    // cell rows the grouped kernel stages per iteration (table layout depends on it)
    bool permuted = false;           // gather tables carry the element position (bits 24..27 of off_w)
    double gather_wavefronts = 0.0;  // shared-memory wavefronts per gather instruction of the schedule

    // ---- reference state
    bool have_ref = false, bounded = false, c64 = false;

    // ---- workspace for column sums
    DevBuf<double> colsum_partial;

    // ---- per-gene layer (calculate_gene_values=True): coverage of every position-sorted gene by the kept windows
    int32_t n_cov = 0;
    DevBuf<int32_t> gv_first, gv_cnt, gv_inv, gv_kaddr;
    DevBuf<double> gv_scratch;

    ~icnv_plan() {
        tab[0].release();
        tab[1].release();
        alpha.release();
        beta.release();
        cw.release();
        tasks_g.release();
        c_scratch.release();
        sp_slot_col.release();
        sp_col_slot.release();
        sp_col_tab.release();
        sp_zrow.release();
        idx_lin.release();
        lo_lin.release();
        hi_lin.release();
        wdir.release();
        tasks_d.release();
        flat_inv.release();
        parts[0].release();
        parts[1].release();
        colsum_partial.release();
        gv_first.release();
        gv_cnt.release();
        gv_inv.release();
        gv_kaddr.release();
        gv_scratch.release();
    }
};

namespace {

constexpr size_t SMEM_MAX = 232448;  // 227 KB opt-in limit per CTA on sm_100
#ifndef ICNV_SMOOTH_ROWS_DEFAULT
#define ICNV_SMOOTH_ROWS_DEFAULT 2
#endif

size_t smem_grouped(const icnv_plan& p, int tier, int rows = 1, bool dbuf = false) {
    size_t s = smooth_scratch_bytes();
    s += (size_t)rows * p.Gpad * 4;
    const size_t nbuf = dbuf ? 2 : rows;  // partial-sum buffers: one per staged row, or two parities of a single row
    s += nbuf * (p.NGpad + PAD_GROUPS) * 16;
    // the third partial sum of a window with a peak group: shared memory for one row, global (L2) scratch for pairs
    if (p.qstar >= 0 && !(tier == 0 && rows == 2)) s += nbuf * (p.NGpad + PAD_GROUPS) * 8;
    if (tier == 1) s += (size_t)p.NQ * 16 + (size_t)p.gs * 8;
    return (s + 15) / 16 * 16;
}
size_t smem_direct(const icnv_plan& p, bool c64) {
    size_t s = smooth_scratch_bytes();
    s += (size_t)p.window * 8;
    s += (size_t)(p.n_sorted + 4) * (c64 ? 8 : 4);
    return (s + 15) / 16 * 16;
}

struct Choice {
    int tier, nwin, gs, tpt;
    size_t smem;
    int rows = 1;  // cell rows staged together per CTA iteration (icnv_smooth.cu, ROWS)
    bool dbuf = false;  // single row with double-buffered partial sums (icnv_smooth.cu, DBUF)
};

// Row pairs (ROWS = 2) exist for the templated window-100 kernel; ICNV_SMOOTH_ROWS=1|2 overrides the default.
int smooth_rows_env() {  // 0 = not set
    const char* e = std::getenv("ICNV_SMOOTH_ROWS");
    return (e && e[0] == '2') ? 2 : ((e && e[0] == '1') ? 1 : 0);
}
int smooth_rows_default() {
    const int v = smooth_rows_env();
    return v ? v : ICNV_SMOOTH_ROWS_DEFAULT;
}

// ICNV_SMOOTH_DBUF=0 (developer A/B): single-row kernels keep one partial-sum buffer (two CTAs per SM when they fit)
bool smooth_dbuf_default() {
    static int v = -1;
    if (v < 0) {
        const char* e = std::getenv("ICNV_SMOOTH_DBUF");
        v = (e && e[0] == '0') ? 0 : 1;
    }
    return v == 1;
}

// ICNV_CSR_SPARSE=0 (developer A/B): CSR input is densified into the staged row of the dense kernels
bool sparse_csr_default() {
    static int v = -1;
    if (v < 0) {
        const char* e = std::getenv("ICNV_CSR_SPARSE");
        v = (e && e[0] == '0') ? 0 : 1;
    }
    return v == 1;
}

// ICNV_CSR_DELTA=0 (developer A/B): CSR input of one-category plans takes the staged-row kernel instead of the delta kernel
bool sparse_delta_default() {
    static int v = -1;
    if (v < 0) {
        const char* e = std::getenv("ICNV_CSR_DELTA");
        v = (e && e[0] == '0') ? 0 : 1;
    }
    return v == 1;
}

// which kernel instantiation runs for this plan + reference dtype
int choose(const icnv_plan& p, bool c64, Choice* c) {
    if (p.group_ok && !c64) {
        const bool templ = (p.step == 10 && (p.window == 100 || p.window == 250));
        if (templ && p.n_tasks_g <= NT && smem_grouped(p, 0) <= SMEM_MAX) {
            *c = {0, p.window, p.gs, 1, smem_grouped(p, 0)};
            if (p.rows == 2) {  // decided at plan creation (the gather tables are laid out for it)
                c->rows = 2;
                c->smem = smem_grouped(p, 0, 2);
            } else if (smooth_dbuf_default() && smem_grouped(p, 0, 1, true) <= SMEM_MAX) {
                c->dbuf = true;
                c->smem = smem_grouped(p, 0, 1, true);
            }
            return 0;
        }
        if (smem_grouped(p, 1) <= SMEM_MAX && p.n_tasks_g <= 4 * NT) {  // TPT 1 or 4
            *c = {1, 0, 0, p.n_tasks_g <= NT ? 1 : 4, smem_grouped(p, 1)};
            return 0;
        }
    }
    // direct form: any gene axis / window / step (the row is staged in parts, icnv_direct.cu)
    *c = {2, 0, 0, 1, std::min(smem_direct(p, c64), SMEM_MAX - 1024), 1};
    return 0;
}

// Parts of the direct kernel for one staging dtype: greedy runs of whole task tiles whose position-sorted genes fit
// into `cap` staged elements.  Tasks are position-ordered, so a run's genes are one contiguous range.
int build_parts(icnv_plan& p, bool c64) {
    const int k = c64 ? 1 : 0;
    if (p.n_parts[k] > 0 || p.n_tasks_d == 0) return 0;
    const size_t elem = c64 ? 8 : 4;
    const size_t avail = SMEM_MAX - 1024 - (size_t)p.window * 8;
    const int64_t cap = (int64_t)(avail / elem) - 4;
    const int n_tiles = (p.n_tasks_d + 31) / 32;
    auto task_end = [&](const Task& t) -> int64_t {
        return (t.w & 0xFF) ? (int64_t)t.x + t.z : (int64_t)t.x + (int64_t)(t.z - 1) * p.step + p.window;
    };
    std::vector<int32_t> parts;
    int64_t max_len = 0;
    int tile = 0;
    while (tile < n_tiles) {
        const int64_t s0 = p.tasks_d_host[(size_t)tile * 32].x;
        int64_t s1 = s0;
        int t1 = tile;
        while (t1 < n_tiles) {
            int64_t e = s1;
            for (int ti = t1 * 32; ti < std::min(p.n_tasks_d, (t1 + 1) * 32); ++ti) e = std::max(e, task_end(p.tasks_d_host[ti]));
            if (e - s0 > cap) break;
            s1 = e;
            ++t1;
        }
        if (t1 == tile) {
            set_error("direct kernel: the genes of one tile of 32 tasks do not fit in shared memory (window " +
                      std::to_string(p.window) + ", step " + std::to_string(p.step) + ")");
            return ICNV_EUNSUPPORTED;
        }
        parts.push_back((int32_t)s0);
        parts.push_back((int32_t)s1);
        parts.push_back(tile);
        parts.push_back(t1);
        max_len = std::max(max_len, s1 - s0);
        tile = t1;
    }
    p.n_parts[k] = (int32_t)(parts.size() / 4);
    p.parts_smem[k] = ((size_t)p.window * 8 + (size_t)(max_len + 4) * elem + 15) / 16 * 16;
    if (p.parts[k].upload(parts)) return ICNV_ECUDA;
    return 0;
}

}  // namespace

// grow-only per-device workspace for the [split][cat][G] partial column sums (stream-ordered use only)
static int colsum_workspace(size_t need, double** out) {
    static double* ws[16] = {nullptr};
    static size_t ws_bytes[16] = {0};
    int devi = 0;
    ICNV_CUDA(cudaGetDevice(&devi));
    if (devi < 0 || devi >= 16) {
        set_error("column sums: device index out of range");
        return ICNV_EINVAL;
    }
    if (ws_bytes[devi] < need) {
        if (ws[devi]) {
            ICNV_CUDA(cudaDeviceSynchronize());  // an earlier launch may still read the old buffer
            ICNV_CUDA(cudaFree(ws[devi]));
        }
        ICNV_CUDA(cudaMalloc(&ws[devi], need));
        ws_bytes[devi] = need;
    }
    *out = ws[devi];
    return 0;
}

extern "C" {

const char* icnv_last_error(void) { return g_err.c_str(); }
int icnv_version(void) { return 100; }

int icnv_plan_create(int device, int32_t n_genes, int32_t n_seg, const int32_t* gene_idx_host,
                     const int32_t* seg_off_host, int32_t window, int32_t step, icnv_plan** out) {
    if (!out || n_genes <= 0 || n_seg < 0 || window < 1 || step < 1 || (n_seg > 0 && (!gene_idx_host || !seg_off_host))) {
        set_error("icnv_plan_create: bad argument");
        return ICNV_EINVAL;
    }
    ICNV_CUDA(cudaSetDevice(device));
    std::unique_ptr<icnv_plan> p(new icnv_plan());
    p->device = device;
    ICNV_CUDA(cudaDeviceGetAttribute(&p->n_sm, cudaDevAttrMultiProcessorCount, device));
    p->G = n_genes;
    p->n_seg = n_seg;
    p->window = window;
    p->step = step;
    p->seg_off.assign(seg_off_host, seg_off_host + n_seg + 1);
    if (n_seg > 0 && p->seg_off[0] != 0) {
        set_error("icnv_plan_create: seg_off[0] must be 0");
        return ICNV_EINVAL;
    }
    const int32_t n_sorted = n_seg > 0 ? p->seg_off[n_seg] : 0;
    p->gene_idx.assign(gene_idx_host, gene_idx_host + n_sorted);
    for (int32_t i = 0; i < n_sorted; ++i)
        if (p->gene_idx[i] < 0 || p->gene_idx[i] >= n_genes) {
            set_error("icnv_plan_create: gene_idx out of range");
            return ICNV_EINVAL;
        }
    p->n_sorted = n_sorted;

    // ---- window grid per segment (tl/_infercnv.py:205-218, :227-236, :335-337)
    const int n = window, s = step;
    std::vector<int64_t> n_out(n_seg);
    std::vector<char> is_flat(n_seg);
    p->out_off.assign(n_seg + 1, 0);
    for (int c = 0; c < n_seg; ++c) {
        const int64_t Gc = p->seg_off[c + 1] - p->seg_off[c];
        if (Gc <= 0) {
            set_error("icnv_plan_create: empty segment");
            return ICNV_EINVAL;
        }
        is_flat[c] = !(n < Gc);
        n_out[c] = is_flat[c] ? 1 : (Gc - n) / s + 1;
        p->out_off[c + 1] = p->out_off[c] + n_out[c];
    }
    p->K = p->out_off[n_seg];
    {
        long long sw = 0;
        for (int j = 0; j < n; ++j) sw += pyr(n, j);
        p->inv_sumw = 1.0 / (double)sw;
    }

    std::vector<double> flat_inv;
    // ---- direct layout
    {
        std::vector<Task> tasks;
        for (int c = 0; c < n_seg; ++c) {
            const int32_t s0 = p->seg_off[c];
            const int32_t Gc = p->seg_off[c + 1] - s0;
            if (is_flat[c]) {
                tasks.push_back({s0, (int32_t)p->out_off[c], Gc, 1 | ((int32_t)flat_inv.size() << 8)});
                flat_inv.push_back(1.0 / (double)Gc);
            } else {
                for (int64_t t = 0; t * LOUT < n_out[c]; ++t)
                    tasks.push_back({(int32_t)(s0 + t * LOUT * s), (int32_t)(p->out_off[c] + t * LOUT),
                                     (int32_t)std::min<int64_t>(LOUT, n_out[c] - t * LOUT), 0});
            }
        }
        p->n_tasks_d = (int32_t)tasks.size();
        p->tasks_d_host = tasks;
        std::vector<double> w(n);
        for (int j = 0; j < n; ++j) w[j] = (double)pyr(n, j);
        if (p->idx_lin.upload(p->gene_idx) || p->tasks_d.upload(tasks) || p->wdir.upload(w)) return ICNV_ECUDA;
    }

    // ---- grouped layout: groups of gs = step genes when step | window
    p->group_ok = (n % s == 0) && (n_genes < (1 << 28));
    if (p->group_ok) {
        const int gs = s;
        p->gs = gs;
        p->NQ = n / gs;
        std::vector<int32_t> gbase(n_seg + 1, 0);
        for (int c = 0; c < n_seg; ++c) {
            const int64_t Gc = p->seg_off[c + 1] - p->seg_off[c];
            const int64_t ng = is_flat[c] ? (Gc + gs - 1) / gs : ((n_out[c] - 1) * s + n) / gs;
            gbase[c + 1] = gbase[c] + (int32_t)ng;
        }
        p->NG = gbase[n_seg];
        p->NGpad = (p->NG + 3) / 4 * 4;
        if (p->NGpad == 0) p->NGpad = 4;
        p->Gpad = (n_genes + 1 + 3) / 4 * 4;
        // ---- weights: within a group the pyramid is linear in j except (at most) the group holding the peak
        std::vector<double> alpha(p->NQ, 0.0), beta(p->NQ, 0.0), cw(gs, 0.0);
        p->qstar = -1;
        for (int q = 0; q < p->NQ; ++q) {
            const int a = pyr(n, gs * q);
            const int b = gs > 1 ? pyr(n, gs * q + 1) - a : 0;
            bool linear = true;
            for (int j = 0; j < gs; ++j) linear = linear && (pyr(n, gs * q + j) == a + b * j);
            if (linear) {
                alpha[q] = a;
                beta[q] = b;
            } else {
                if (p->qstar >= 0) {
                    set_error("internal: two non-linear weight groups");
                    return ICNV_EINVAL;
                }
                p->qstar = q;
                for (int j = 0; j < gs; ++j) cw[j] = pyr(n, gs * q + j);
            }
        }
        std::vector<Task> tasks;
        for (int c = 0; c < n_seg; ++c) {
            if (is_flat[c]) {
                int fid = 0;  // same order as the direct layout
                for (int cc = 0; cc < c; ++cc) fid += is_flat[cc];
                tasks.push_back({gbase[c], (int32_t)p->out_off[c], gbase[c + 1] - gbase[c], 1 | (fid << 8)});
            } else {
                for (int64_t t = 0; t * LOUT < n_out[c]; ++t)
                    tasks.push_back({(int32_t)(gbase[c] + t * LOUT), (int32_t)(p->out_off[c] + t * LOUT),
                                     (int32_t)std::min<int64_t>(LOUT, n_out[c] - t * LOUT), 0});
            }
        }
        p->n_tasks_g = (int32_t)tasks.size();

        // ---- assign groups to (warp-block, u, lane) slots and order every lane's walk over its group: icnv_schedule.cu.
        // The permuted walk needs the element position j in the table entry (bits 24..27) and is what the templated
        // kernel without a peak group decodes (tier 0, window 100); every other kernel walks j = 0..gs-1.
        const int nquads = p->NGpad / 4;
        const int n_wb = (nquads + 31) / 32;
        const int nsets = n_wb * 4;
        std::vector<int32_t> gcol((size_t)p->NG * gs, -1);  // column of element j of group g, -1 = pad
        for (int c = 0; c < n_seg; ++c) {
            const int32_t s0 = p->seg_off[c];
            const int32_t Gc = p->seg_off[c + 1] - s0;
            for (int32_t g = gbase[c]; g < gbase[c + 1]; ++g)
                for (int j = 0; j < gs; ++j) {
                    const int32_t pos = (g - gbase[c]) * gs + j;
                    if (pos < Gc) gcol[(size_t)g * gs + j] = p->gene_idx[s0 + pos];
                }
        }
        const bool templ_ws = (s == 10 && (n == 100 || n == 250));  // the templated tier-0 instantiations
        p->permuted = templ_ws && gs <= 16 && p->n_tasks_g <= NT && smem_grouped(*p, 0) <= SMEM_MAX;
        // non-linear part of the peak group's weights, m_j = cw_j - cw_0 (bits 28..31 of a permuted table entry)
        std::vector<int> mj(gs, 0);
        if (p->qstar >= 0)
            for (int j = 0; j < gs; ++j) {
                mj[j] = pyr(n, gs * p->qstar + j) - pyr(n, gs * p->qstar);
                if (mj[j] < 0 || mj[j] > 15) p->permuted = false;
            }
        std::vector<int32_t> slot_group;
        std::vector<uint8_t> order;
        // ICNV_GATHER_PERM=0 (developer A/B): keep the natural walk; the entries still carry j
        const char* perm_env = std::getenv("ICNV_GATHER_PERM");
        const bool natural_walk = ICNV_NATURAL_WALK == 2 || (ICNV_NATURAL_WALK == 1 && p->qstar >= 0);  // kernels take j from the loop
        const bool optimise_walk = p->permuted && !natural_walk && !(perm_env && perm_env[0] == '0');
        p->gather_wavefronts = schedule_gathers(gcol, p->NG, gs, n_genes, nsets, optimise_walk, slot_group, order);
        if (p->gather_wavefronts < 0) {
            set_error("internal: group slots exhausted");
            return ICNV_EINVAL;
        }
        // tables in kernel order: entry ((unit*gs + t)*32 + lane)*2 + (u & 1), unit = 2*wb + u/2  <->  the element the lane in
        // slot (wb, u, lane) reads at step t
        const size_t n_entries = (size_t)n_wb * gs * 32 * 4;
        // the kernel that will run decides the unit width: row pairs (templated window 100 that fits twice) walk whole
        // warp-blocks, everything else half warp-blocks
        // row pairs pay for windows without a peak group (window 100: 0.72 of the HBM roofline vs 0.61-0.66 single);
        // with one (window 250) the pair kernel measured 3.8 ms / 100k cells vs 2.9 single: phase 3 is 2.5x heavier and
        // the half-width units leave the gathers latency-bound -> single row with double-buffered partials instead
        // (ICNV_SMOOTH_ROWS=2 forces pairs)
        const bool want_pairs = p->qstar < 0 ? smooth_rows_default() == 2 : smooth_rows_env() == 2;
        const bool pairs = p->permuted && want_pairs && smem_grouped(*p, 0, 2) <= SMEM_MAX;
        p->rows = pairs ? 2 : 1;
        uint32_t raw_base = 0;
        if (smooth_raw_base(&raw_base)) return ICNV_ECUDA;
        p->raw_base = raw_base;
        if (p->permuted && (raw_base + (uint32_t)p->Gpad * 4u * 2u) >= (1u << 24)) {
            set_error("internal: shared window address does not fit the table entry");
            return ICNV_EINVAL;
        }
        for (int ts = 0; ts < (pairs ? 2 : 1); ++ts) {
        const int uw = ts == 0 ? ICNV_UNIT_WIDTH(pairs ? 2 : 1, p->qstar >= 0) : ICNV_UNIT_WIDTH(1, p->qstar >= 0);
        std::vector<uint32_t> off(n_entries, raw_base + (uint32_t)n_genes * 4u);
        std::vector<int32_t> cols(n_entries, -1);
        std::vector<int32_t> grp((size_t)n_wb * 32 * 4, p->NGpad);  // empty slots store their zeros to a pad group
        for (int wb = 0; wb < n_wb; ++wb)
            for (int lane = 0; lane < 32; ++lane)
                for (int u = 0; u < 4; ++u) {
                    const int32_t g = slot_group[((size_t)wb * 4 + u) * 32 + lane];
                    if (g < 0) continue;
                    // work unit = 32 lanes x uw groups (uw = 4: a whole warp-block, uw = 2: half of one)
                    const size_t unit = (size_t)wb * (4 / uw) + (u / uw);
                    grp[(unit * 32 + lane) * uw + (u % uw)] = g;
                    for (int t = 0; t < gs; ++t) {
                        const int j = order[(((size_t)wb * 4 + u) * 32 + lane) * gs + t];  // element read at step t
                        const int32_t col = gcol[(size_t)g * gs + j];
                        const size_t e = ((unit * gs + t) * 32 + lane) * uw + (u % uw);
                        if (col >= 0) {
                            off[e] = raw_base + (uint32_t)col * 4u;
                            cols[e] = col;
                        }
                        if (p->permuted && !natural_walk) off[e] |= ((uint32_t)j << 24) | ((uint32_t)mj[j] << 28);  // pads too (x = 0 either way)
                    }
                }
        auto& T = p->tab[ts];
        T.uw = uw;
        if (T.off_w.upload(off) || T.cols_w.upload(cols) || T.grp_w.upload(grp) || T.lo_w.alloc(n_entries) || T.hi_w.alloc(n_entries))
            return ICNV_ECUDA;
        }  // table sets
        if (p->alpha.upload(alpha) || p->beta.upload(beta) || p->cw.upload(cw) || p->tasks_g.upload(tasks)) return ICNV_ECUDA;
        // sparse-aware CSR path: per-slot column and per-column slot of the position-ordered padded row
        p->sp_DP = p->NGpad * gs;
        p->sparse_ok = p->permuted && sparse_supported(n, gs) && sparse_smem_bytes(p->sp_DP, p->NGpad, p->qstar >= 0) <= SMEM_MAX;
        if (p->sparse_ok) {
            std::vector<int32_t> slot_col((size_t)p->sp_DP, -1), col_slot((size_t)n_genes, -1);
            for (size_t i = 0; i < gcol.size(); ++i) {
                slot_col[i] = gcol[i];
                if (gcol[i] >= 0) col_slot[gcol[i]] = (int32_t)i;
            }
            if (p->sp_slot_col.upload(slot_col) || p->sp_col_slot.upload(col_slot) || p->sp_col_tab.alloc((size_t)n_genes) ||
                p->sp_zrow.alloc((size_t)p->sp_DP))
                return ICNV_ECUDA;
        }
        p->delta_ok = p->sparse_ok && sparse_delta_supported(n, gs, p->NGpad, p->n_tasks_g) &&
                      sparse_delta_smem_bytes(n_genes, p->NGpad, p->qstar >= 0) <= SMEM_MAX;
        if (p->delta_ok) {
            const size_t tmp_w = (size_t)((p->n_tasks_g + 31) / 32) * (32 * LOUT + 1);
            if (p->sp_gj.alloc(((size_t)n_genes + 7) & ~(size_t)7) || p->sp_ref.alloc((size_t)n_genes) || p->sp_base.alloc(tmp_w) ||
                p->sp_indptr0.upload(std::vector<int64_t>(2, 0)))
                return ICNV_ECUDA;
        }
        if (p->permuted && p->qstar >= 0 && p->c_scratch.alloc((size_t)p->n_sm * 8 * (p->NGpad + PAD_GROUPS))) return ICNV_ECUDA;
    }
    if (flat_inv.empty()) flat_inv.push_back(1.0);
    if (p->flat_inv.upload(flat_inv)) return ICNV_ECUDA;
    if (p->lo_lin.alloc((size_t)(n_sorted + 4) * 8) || p->hi_lin.alloc((size_t)(n_sorted + 4) * 8)) return ICNV_ECUDA;

    // ---- per-gene layer tables (tl/_infercnv.py:214-223, :238-242, :278-287): the gene at sorted position pos of a
    // regular chromosome lies in the kept windows k with k*step <= pos < k*step + window; every gene of a flat
    // chromosome takes its single column.  Task index -> tile-order address is the same for both task lists.
    {
        std::vector<int32_t> first, cnt, inv(n_genes, -1), kaddr((size_t)p->K, 0);
        int32_t ti = 0;
        for (int c = 0; c < n_seg; ++c) {
            const int32_t s0 = p->seg_off[c];
            const int32_t Gc = p->seg_off[c + 1] - s0;
            const int64_t no = n_out[c];
            for (int64_t t = 0; t * LOUT < no; ++t, ++ti)
                for (int64_t i = 0; i < std::min<int64_t>(LOUT, no - t * LOUT); ++i)
                    kaddr[(size_t)(p->out_off[c] + t * LOUT + i)] = (ti / 32) * (32 * LOUT) + (int32_t)i * 32 + (ti % 32);
            for (int32_t pos = 0; pos < Gc; ++pos) {
                int64_t k_lo = 0, k_hi = 0;
                if (!is_flat[c]) {
                    k_lo = pos - n + 1 <= 0 ? 0 : (pos - n + 1 + s - 1) / s;
                    k_hi = std::min<int64_t>(no - 1, pos / s);
                    if (k_lo > k_hi) continue;
                }
                inv[p->gene_idx[s0 + pos]] = (int32_t)first.size();
                first.push_back((int32_t)(p->out_off[c] + k_lo));
                cnt.push_back((int32_t)(k_hi - k_lo + 1));
            }
        }
        p->n_cov = (int32_t)first.size();
        if (p->gv_first.upload(first) || p->gv_cnt.upload(cnt) || p->gv_inv.upload(inv) || p->gv_kaddr.upload(kaddr)) return ICNV_ECUDA;
    }

    Choice ch;
    p->base_tier = choose(*p, false, &ch) == 0 ? ch.tier : -1;
    *out = p.release();
    return ICNV_OK;
}

void icnv_plan_destroy(icnv_plan* plan) { delete plan; }

int icnv_plan_out_width(const icnv_plan* plan, int64_t* K) {
    if (!plan || !K) return ICNV_EINVAL;
    *K = plan->K;
    return ICNV_OK;
}
int icnv_plan_out_offsets(const icnv_plan* plan, int64_t* out_off_host) {
    if (!plan || !out_off_host) return ICNV_EINVAL;
    std::copy(plan->out_off.begin(), plan->out_off.end(), out_off_host);
    return ICNV_OK;
}
int icnv_plan_tmp_width(const icnv_plan* plan, int64_t* ld_tmp) {
    if (!plan || !ld_tmp) return ICNV_EINVAL;
    Choice ch;
    int rc = choose(*plan, plan->c64, &ch);
    if (rc) return rc;
    const int n_tasks = ch.tier < 2 ? plan->n_tasks_g : plan->n_tasks_d;
    *ld_tmp = (int64_t)((n_tasks + 31) / 32) * (32 * LOUT + 1);  // values + one float2 of moments per tile
    return ICNV_OK;
}
int icnv_plan_kernel_tier(const icnv_plan* plan) {
    if (!plan) return ICNV_EINVAL;
    Choice ch;
    if (choose(*plan, plan->c64, &ch)) return ICNV_EUNSUPPORTED;
    return ch.tier;
}

int icnv_plan_rows_per_iteration(const icnv_plan* plan) {
    if (!plan) return ICNV_EINVAL;
    Choice ch;
    if (choose(*plan, plan->c64, &ch)) return ICNV_EUNSUPPORTED;
    return ch.rows;
}

int icnv_plan_launch_info(icnv_plan* plan, int32_t* ctas_per_sm, int32_t* threads, int32_t* smem_bytes, int32_t* n_sm) {
    if (!plan) return ICNV_EINVAL;
    Choice ch;
    int rc = choose(*plan, plan->c64, &ch);
    if (rc) return rc;
    int occ = 1;
    if (ch.tier < 2) {
        rc = smooth_occupancy(ch.tier, ch.nwin, ch.gs, plan->bounded, plan->c64, ch.tpt, ch.rows, ch.dbuf, ch.smem, &occ);
        if (rc) return rc;
    }
    if (ctas_per_sm) *ctas_per_sm = occ;
    if (threads) *threads = smooth_threads(ch.rows, ch.dbuf);
    if (smem_bytes) *smem_bytes = (int32_t)ch.smem;
    if (n_sm) *n_sm = plan->n_sm;
    return ICNV_OK;
}

int icnv_colsum_dense_f32(const float* X, int64_t n_rows, int64_t ldx, int32_t G, const int32_t* row_cat, int32_t n_cat,
                          double* sums, int64_t* counts, void* stream) {
    if (!X || !sums || !counts || n_rows < 0 || G <= 0 || n_cat <= 0 || ldx < G) {
        set_error("icnv_colsum_dense_f32: bad argument");
        return ICNV_EINVAL;
    }
    // row splits: exactly one wave of resident CTAs (a partial second wave idles most of the machine), each with a
    // decent run of rows
    int slots = 0;
    if (aux_colsum_dense_slots(&slots)) return ICNV_ECUDA;
    const int64_t per_split = (int64_t)std::max(1, (G + 1023) / 1024) * n_cat;
    int n_split = (int)std::min<int64_t>(std::max<int64_t>(1, n_rows / 64), std::max<int64_t>(1, slots / per_split));
    double* ws = nullptr;
    const int rcw = colsum_workspace(sizeof(double) * (size_t)n_split * n_cat * G, &ws);
    if (rcw) return rcw;
    return aux_colsum_dense(X, n_rows, ldx, G, row_cat, n_cat, sums, counts, ws, n_split, (cudaStream_t)stream);
}

int icnv_colsum_csr_f32(const int64_t* indptr, const int32_t* indices, const float* data, int64_t n_rows, int32_t G,
                        const int32_t* row_cat, int32_t n_cat, double* sums, int64_t* counts, void* stream) {
    if (!indptr || !sums || !counts || n_rows < 0 || G <= 0 || n_cat <= 0) {
        set_error("icnv_colsum_csr_f32: bad argument");
        return ICNV_EINVAL;
    }
    int devi = 0, n_sm = 148;
    ICNV_CUDA(cudaGetDevice(&devi));
    ICNV_CUDA(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, devi));
    const int n_split = sparse_colsum_splits(n_rows, n_sm);
    double* ws = nullptr;
    int rc = colsum_workspace(sizeof(double) * (size_t)n_split * n_cat * G, &ws);
    if (rc) return rc;
    return aux_colsum_csr(indptr, indices, data, n_rows, G, row_cat, n_cat, sums, counts, ws, n_split, (cudaStream_t)stream);
}

int icnv_mean_from_sums(const double* sums, const int64_t* counts, int32_t n_cat, int32_t G, void* ref_out, int32_t out_is_f64,
                        void* stream) {
    if (!sums || !counts || !ref_out) return ICNV_EINVAL;
    return aux_mean_from_sums(sums, counts, n_cat, G, ref_out, out_is_f64 != 0, (cudaStream_t)stream);
}
int icnv_nnz_to_indptr(const int32_t* row_nnz, int64_t n_rows, int64_t* indptr, void* stream) {
    if (!indptr || n_rows < 0 || (n_rows > 0 && !row_nnz)) return ICNV_EINVAL;
    return aux_nnz_to_indptr(row_nnz, n_rows, indptr, (cudaStream_t)stream);
}

int icnv_plan_set_reference(icnv_plan* plan, const void* ref, int32_t n_cat, int32_t ref_is_f64, void* stream) {
    if (!plan || !ref || n_cat < 1) {
        set_error("icnv_plan_set_reference: bad argument");
        return ICNV_EINVAL;
    }
    cudaStream_t st = (cudaStream_t)stream;
    const bool c64 = ref_is_f64 != 0;
    Choice ch;
    int rc = choose(*plan, c64, &ch);
    if (rc) return rc;
    if (ch.tier < 2) {
        for (int ts = 0; ts < 2 && !rc; ++ts)
            if (plan->tab[ts].uw)
                rc = aux_build_bounds(ref, false, n_cat, plan->G, plan->tab[ts].cols_w.ptr, (int64_t)plan->tab[ts].cols_w.n,
                                      plan->tab[ts].lo_w.ptr, plan->tab[ts].hi_w.ptr, false, st);
        if (!rc && plan->sparse_ok) rc = sparse_col_table_launch(ref, false, n_cat, plan->G, plan->sp_col_slot.ptr, plan->sp_col_tab.ptr, st);
        if (!rc && plan->delta_ok && n_cat == 1)
            rc = sparse_delta_tables_launch(plan->sp_col_tab.ptr, plan->G, plan->gs, plan->sp_gj.ptr, plan->sp_ref.ptr, st);
    } else {
        rc = aux_build_bounds(ref, c64, n_cat, plan->G, plan->idx_lin.ptr, plan->n_sorted, plan->lo_lin.ptr,
                              plan->hi_lin.ptr, c64, st);
    }
    if (rc) return rc;
    plan->have_ref = true;
    plan->bounded = n_cat > 1;
    plan->c64 = c64;
    return ICNV_OK;
}

static long long* g_dbg_ptr = nullptr;
static int g_dbg_rows = 0;
extern "C" int icnv_debug_set_timeline(long long* dev_buf, int rows_per_cta) {
    g_dbg_ptr = dev_buf;
    g_dbg_rows = rows_per_cta;
    return ICNV_OK;
}

static int smooth_common(icnv_plan* plan, SmoothParams& sp, double lfc_clip, double* out, int64_t ldo, void* stream) {
    if (!plan->have_ref) {
        set_error("smooth: icnv_plan_set_reference has not been called");
        return ICNV_EINVAL;
    }
    if (!out || !(lfc_clip >= 0)) {
        set_error("smooth: bad argument");
        return ICNV_EINVAL;
    }
    if (sp.n_rows == 0 || plan->K == 0) return ICNV_OK;
    Choice ch;
    int rc = choose(*plan, plan->c64, &ch);
    if (rc) return rc;
    {
        const int n_tasks = ch.tier < 2 ? plan->n_tasks_g : plan->n_tasks_d;
        if (ldo < (int64_t)((n_tasks + 31) / 32) * (32 * LOUT + 1)) {
            set_error("smooth: intermediate pitch smaller than icnv_plan_tmp_width");
            return ICNV_EINVAL;
        }
    }
    if (ch.tier == 2 && !sp.X) {
        set_error("smooth: CSR input is only fused into the grouped kernels; densify first for this (window, step)");
        return ICNV_EUNSUPPORTED;
    }
    if (ch.tier == 2) {
        const int k = plan->c64 ? 1 : 0;
        rc = build_parts(*plan, plan->c64);
        if (rc) return rc;
        DirectParams dp;
        memset(&dp, 0, sizeof(dp));
        dp.X = sp.X;
        dp.ldx = sp.ldx;
        dp.n_rows = sp.n_rows;
        dp.idx_lin = plan->idx_lin.ptr;
        dp.lo_lin = plan->lo_lin.ptr;
        dp.hi_lin = plan->hi_lin.ptr;
        dp.wdir = plan->wdir.ptr;
        dp.window = plan->window;
        dp.step = plan->step;
        dp.tasks = plan->tasks_d.ptr;
        dp.n_tasks = plan->n_tasks_d;
        dp.parts = reinterpret_cast<const int4*>(plan->parts[k].ptr);
        dp.n_parts = plan->n_parts[k];
        dp.clip = plan->c64 ? lfc_clip : (double)(float)lfc_clip;
        dp.clipf = (float)lfc_clip;
        dp.inv_sumw = plan->inv_sumw;
        dp.flat_inv = plan->flat_inv.ptr;
        dp.out = out;
        dp.ldo = ldo;
        const int grid = (int)std::min<int64_t>(sp.n_rows, (int64_t)plan->n_sm);
        return direct_launch(dp, plan->bounded, plan->c64, grid, plan->parts_smem[k], (cudaStream_t)stream);
    }
    if (!sp.X && ch.tier == 0 && plan->sparse_ok && sparse_csr_default()) {
        // CSR input: scatter the stored entries over the constant row instead of densifying in matrix order
        rc = sparse_zrow_launch(plan->sp_slot_col.ptr, plan->sp_DP, plan->sp_col_tab.ptr, (float)lfc_clip, plan->bounded, plan->sp_zrow.ptr,
                                (cudaStream_t)stream);
        if (rc) return rc;
        if (plan->delta_ok && !plan->bounded && sparse_delta_default()) {
            // smooth(row) = smooth(constant row) + smooth(deltas of the stored entries): the constant row's result comes from
            // the staged-row kernel on ONE empty row, the deltas from the entry-proportional kernel
            rc = sparse_smooth_launch(plan->window, plan->gs, false, plan->sp_indptr0.ptr, sp.indices, sp.data, 1, plan->sp_col_tab.ptr,
                                      plan->sp_zrow.ptr, plan->sp_DP, plan->NG, plan->NGpad, plan->inv_sumw, plan->flat_inv.ptr,
                                      plan->tasks_g.ptr, plan->n_tasks_g, (float)lfc_clip, plan->sp_base.ptr, (int64_t)plan->sp_base.n,
                                      plan->n_sm, (cudaStream_t)stream);
            if (rc) return rc;
            return sparse_delta_launch(plan->window, plan->gs, sp.indptr, sp.indices, sp.data, sp.n_rows, plan->sp_gj.ptr, plan->sp_ref.ptr,
                                       plan->G, plan->NG, plan->NGpad, plan->sp_base.ptr, plan->inv_sumw, plan->flat_inv.ptr,
                                       plan->tasks_g.ptr, plan->n_tasks_g, (float)lfc_clip, out, ldo, plan->n_sm, (cudaStream_t)stream);
        }
        return sparse_smooth_launch(plan->window, plan->gs, plan->bounded, sp.indptr, sp.indices, sp.data, sp.n_rows, plan->sp_col_tab.ptr,
                                    plan->sp_zrow.ptr, plan->sp_DP, plan->NG, plan->NGpad, plan->inv_sumw, plan->flat_inv.ptr,
                                    plan->tasks_g.ptr, plan->n_tasks_g, (float)lfc_clip, out, ldo, plan->n_sm, (cudaStream_t)stream);
    }
    sp.G = plan->G;
    sp.Gpad = plan->Gpad;
    sp.gs = plan->gs;
    sp.NG = plan->NG;
    sp.NGpad = plan->NGpad;
    sp.NQ = plan->NQ;
    sp.qstar = plan->qstar;
    // CSR input runs the single-row kernel (two CTAs per SM) even when dense input takes row pairs
    if (ch.tier == 0 && ch.rows == 2 && !sp.X && plan->tab[1].uw) {
        ch.rows = 1;
        ch.smem = smem_grouped(*plan, 0, 1);
    }
    const auto& T = plan->tab[(ch.tier == 0 && plan->rows == 2 && ch.rows == 1) ? 1 : 0];
    sp.off_w = T.off_w.ptr;
    sp.raw_base = plan->raw_base;
    sp.grp_w = T.grp_w.ptr;
    sp.lo_w = T.lo_w.ptr;
    sp.hi_w = T.hi_w.ptr;
    sp.alpha = plan->alpha.ptr;
    sp.beta = plan->beta.ptr;
    sp.cw = plan->cw.ptr;
    sp.c_scratch = plan->c_scratch.ptr;
    sp.clip = plan->c64 ? lfc_clip : (double)(float)lfc_clip;
    sp.clipf = (float)lfc_clip;
    sp.inv_sumw = plan->inv_sumw;
    sp.flat_inv = plan->flat_inv.ptr;
    sp.tasks = ch.tier < 2 ? plan->tasks_g.ptr : plan->tasks_d.ptr;
    sp.n_tasks = ch.tier < 2 ? plan->n_tasks_g : plan->n_tasks_d;
    sp.K = (int32_t)plan->K;
    sp.out = out;
    sp.ldo = ldo;
    sp.dbg = g_dbg_ptr;
    sp.dbg_rows = g_dbg_rows;
    sp.use_tma = sp.X && (sp.ldx % 4 == 0) && ((reinterpret_cast<uintptr_t>(sp.X) & 15) == 0) && (plan->G % 4 == 0);
    {
        const char* e = std::getenv("ICNV_L2_PREFETCH");  // developer A/B switch
        sp.l2_prefetch = (e && e[0] == '0') ? 0 : 1;
        const char* e2 = std::getenv("ICNV_SPLIT_ROWS");
        sp.split_rows = (e2 && e2[0] == '0') ? 0 : 1;
    }
    if (ch.dbuf && !sp.use_tma) {  // the double-buffered kernel overlaps TMA fills with phase 3; other inputs keep two CTAs per SM
        ch.dbuf = false;
        ch.smem = smem_grouped(*plan, 0, 1);
    }
    int occ = 0;
    rc = smooth_occupancy(ch.tier, ch.nwin, ch.gs, plan->bounded, plan->c64, ch.tpt, ch.rows, ch.dbuf, ch.smem, &occ);
    if (rc) return rc;
    if (occ < 1) {
        set_error("smooth: kernel does not fit on an SM");
        return ICNV_EUNSUPPORTED;
    }
    const int grid = (int)std::min<int64_t>((sp.n_rows + ch.rows - 1) / ch.rows, (int64_t)plan->n_sm * occ);
    return smooth_launch(ch.tier, ch.nwin, ch.gs, plan->bounded, plan->c64, ch.tpt, ch.rows, ch.dbuf, sp, grid, ch.smem, (cudaStream_t)stream);
}

int icnv_smooth_dense_f32(icnv_plan* plan, const float* X, int64_t n_rows, int64_t ldx, double lfc_clip, double* tmp,
                          int64_t ld_tmp, void* stream) {
    if (!plan || !X || n_rows < 0 || ldx < plan->G) {
        set_error("icnv_smooth_dense_f32: bad argument");
        return ICNV_EINVAL;
    }
    SmoothParams sp;
    memset(&sp, 0, sizeof(sp));
    sp.X = X;
    sp.ldx = ldx;
    sp.n_rows = n_rows;
    return smooth_common(plan, sp, lfc_clip, tmp, ld_tmp, stream);
}

int icnv_smooth_csr_f32(icnv_plan* plan, const int64_t* indptr, const int32_t* indices, const float* data, int64_t n_rows,
                        double lfc_clip, double* tmp, int64_t ld_tmp, void* stream) {
    if (!plan || !indptr || n_rows < 0) {
        set_error("icnv_smooth_csr_f32: bad argument");
        return ICNV_EINVAL;
    }
    SmoothParams sp;
    memset(&sp, 0, sizeof(sp));
    sp.indptr = indptr;
    sp.indices = indices;
    sp.data = data;
    sp.n_rows = n_rows;
    return smooth_common(plan, sp, lfc_clip, tmp, ld_tmp, stream);
}

int icnv_center_rows(icnv_plan* plan, const double* tmp, int64_t n_rows, int64_t ld_tmp, void* out, int32_t out_is_f64, int64_t ldo,
                     double* row_stats, void* stream) {
    if (!plan || !tmp || !out || !row_stats || ldo < plan->K) {
        set_error("icnv_center_rows: bad argument");
        return ICNV_EINVAL;
    }
    if (n_rows == 0 || plan->K == 0) return ICNV_OK;
    Choice ch;
    int rc = choose(*plan, plan->c64, &ch);
    if (rc) return rc;
    const Task* tasks = ch.tier < 2 ? plan->tasks_g.ptr : plan->tasks_d.ptr;
    const int n_tasks = ch.tier < 2 ? plan->n_tasks_g : plan->n_tasks_d;
    if (ld_tmp < (int64_t)((n_tasks + 31) / 32) * (32 * LOUT + 1)) {
        set_error("icnv_center_rows: intermediate pitch smaller than icnv_plan_tmp_width");
        return ICNV_EINVAL;
    }
    if ((n_tasks + 31) / 32 > 28)  // wider than the warp-per-row kernel's byte counters: CTA-per-row selection
        return center_wide_launch(tmp, n_rows, ld_tmp, plan->gv_kaddr.ptr, (int)plan->K, out, out_is_f64 != 0, ldo, row_stats,
                                  plan->n_sm, (cudaStream_t)stream);
    return aux_center_rows(tmp, n_rows, ld_tmp, tasks, n_tasks, (int)plan->K, out, out_is_f64 != 0, ldo, row_stats, (cudaStream_t)stream);
}

int icnv_gene_values(icnv_plan* plan, const double* tmp, int64_t n_rows, int64_t ld_tmp, int64_t chunk_rows, const double* thr,
                     double* gene_out, int64_t ldg, void* stream) {
    if (!plan || !tmp || !gene_out || chunk_rows < 1 || ldg < plan->G || n_rows < 0) {
        set_error("icnv_gene_values: bad argument");
        return ICNV_EINVAL;
    }
    if (n_rows == 0) return ICNV_OK;
    Choice ch;
    int rc = choose(*plan, plan->c64, &ch);
    if (rc) return rc;
    const int n_tasks = ch.tier < 2 ? plan->n_tasks_g : plan->n_tasks_d;
    if (ld_tmp < (int64_t)((n_tasks + 31) / 32) * (32 * LOUT + 1)) {
        set_error("icnv_gene_values: intermediate pitch smaller than icnv_plan_tmp_width");
        return ICNV_EINVAL;
    }
    GeneValParams gp;
    memset(&gp, 0, sizeof(gp));
    gp.tmp = tmp;
    gp.n_rows = n_rows;
    gp.ld = ld_tmp;
    gp.kaddr = plan->gv_kaddr.ptr;
    gp.K = (int32_t)plan->K;
    gp.first = plan->gv_first.ptr;
    gp.cnt = plan->gv_cnt.ptr;
    gp.n_cov = plan->n_cov;
    gp.inv = plan->gv_inv.ptr;
    gp.G = plan->G;
    gp.chunk_rows = chunk_rows;
    gp.thr = thr;
    gp.out = gene_out;
    gp.ldo = ldg;
    // shared memory: window values first (small), then the per-gene means if they still fit
    const size_t budget = SMEM_MAX - 4096;
    size_t smem = 0;
    if ((size_t)plan->K * 8 <= budget) {
        gp.k_in_smem = 1;
        smem += (size_t)plan->K * 8;
    }
    if (smem + (size_t)plan->n_cov * 8 <= budget) {
        gp.v_in_smem = 1;
        smem += (size_t)plan->n_cov * 8;
    }
    {   // work area of the histogram selection (16 per-warp histograms of 512 bins + 2048 candidates) when it still fits
        const size_t work = (size_t)16 * 512 * 4 + (size_t)2048 * 8;
        const size_t at = (smem + 15) / 16 * 16;
        gp.hist_off = -1;
        if (at + work <= budget) {
            gp.hist_off = (int32_t)at;
            smem = at + work;
        }
    }
    const int grid = (int)std::min<int64_t>(n_rows, (int64_t)plan->n_sm);  // 118 registers x 512 threads: one CTA per SM
    if (!gp.v_in_smem) {
        const size_t need = (size_t)grid * std::max(plan->n_cov, 1);
        if (plan->gv_scratch.n < need) {
            ICNV_CUDA(cudaStreamSynchronize((cudaStream_t)stream));  // an earlier launch may still use the old buffer
            plan->gv_scratch.release();
            if (plan->gv_scratch.alloc(need)) return ICNV_ECUDA;
        }
        gp.scratch = plan->gv_scratch.ptr;
    }
    return genevals_launch(gp, grid, smem < 16 ? 16 : smem, (cudaStream_t)stream);
}

int icnv_plan_gather_cost(const icnv_plan* plan, double* wavefronts_per_gather) {
    if (!plan || !wavefronts_per_gather) return ICNV_EINVAL;
    *wavefronts_per_gather = plan->group_ok ? plan->gather_wavefronts : 0.0;
    return ICNV_OK;
}

int icnv_plan_gene_coverage(const icnv_plan* plan, int32_t* n_covered) {
    if (!plan || !n_covered) return ICNV_EINVAL;
    *n_covered = plan->n_cov;
    return ICNV_OK;
}

int icnv_chunk_threshold(const double* row_stats, int64_t n_rows, int64_t K, int64_t chunk_rows, double dyn_thr, double* thr,
                         void* stream) {
    if (!row_stats || !thr || chunk_rows < 1 || K < 1) {
        set_error("icnv_chunk_threshold: bad argument");
        return ICNV_EINVAL;
    }
    return aux_chunk_threshold(row_stats, n_rows, K, chunk_rows, dyn_thr, thr, (cudaStream_t)stream);
}

int icnv_apply_threshold(void* out, int32_t out_is_f64, int64_t n_rows, int64_t K, int64_t ldo, int64_t chunk_rows,
                         const double* thr, double* row_abs_sum, int32_t* row_nnz, void* stream) {
    if (!out || chunk_rows < 1 || ldo < K) {
        set_error("icnv_apply_threshold: bad argument");
        return ICNV_EINVAL;
    }
    return aux_apply_threshold(out, out_is_f64 != 0, n_rows, K, ldo, chunk_rows, thr, row_abs_sum, row_nnz, (cudaStream_t)stream);
}

int icnv_filter_count(const void* out, int32_t out_is_f64, int64_t n_rows, int64_t K, int64_t ldo, int64_t chunk_rows,
                      const double* thr, double* row_abs_sum, int32_t* row_nnz, void* stream) {
    if ((n_rows > 0 && !out) || chunk_rows < 1 || ldo < K) {
        set_error("icnv_filter_count: bad argument");
        return ICNV_EINVAL;
    }
    return filter_count(out, out_is_f64 != 0, n_rows, K, ldo, chunk_rows, thr, row_abs_sum, row_nnz, (cudaStream_t)stream);
}

int icnv_filter_to_csr(const void* out, int32_t out_is_f64, int64_t n_rows, int64_t K, int64_t ldo, int64_t chunk_rows,
                       const double* thr, const int64_t* indptr, int32_t* indices, void* data, int32_t data_is_f64, void* stream) {
    if ((n_rows > 0 && (!out || !indptr)) || chunk_rows < 1 || ldo < K || (out_is_f64 && !data_is_f64)) {
        set_error("icnv_filter_to_csr: bad argument");
        return ICNV_EINVAL;
    }
    return filter_to_csr(out, out_is_f64 != 0, n_rows, K, ldo, chunk_rows, thr, indptr, indices, data, data_is_f64 != 0, (cudaStream_t)stream);
}

int icnv_dense_to_csr(const void* out, int32_t out_is_f64, int64_t n_rows, int64_t K, int64_t ldo, const int64_t* indptr,
                      int32_t* indices, void* data, void* stream) {
    if (!out || !indptr) return ICNV_EINVAL;
    return aux_dense_to_csr(out, out_is_f64 != 0, n_rows, K, ldo, indptr, indices, data, (cudaStream_t)stream);
}

int icnv_rowabs_csr(const int64_t* indptr, const void* data, int32_t data_is_f64, int64_t n_rows, double* row_abs_sum,
                    void* stream) {
    if (!indptr || !row_abs_sum) return ICNV_EINVAL;
    return aux_rowabs_csr(indptr, data, data_is_f64 != 0, n_rows, row_abs_sum, (cudaStream_t)stream);
}
int icnv_rowabs_dense(const void* X, int32_t is_f64, int64_t n_rows, int64_t K, int64_t ld, double* row_abs_sum, void* stream) {
    if (!X || !row_abs_sum) return ICNV_EINVAL;
    return aux_rowabs_dense(X, is_f64 != 0, n_rows, K, ld, row_abs_sum, (cudaStream_t)stream);
}
int icnv_label_sums(const double* row_abs_sum, const int32_t* labels, int64_t n_rows, int32_t n_labels, double* label_sum,
                    int64_t* label_rows, void* stream) {
    if (!row_abs_sum || !labels || !label_sum || !label_rows) return ICNV_EINVAL;
    return aux_label_sums(row_abs_sum, labels, n_rows, n_labels, label_sum, label_rows, (cudaStream_t)stream);
}

int icnv_row_corrcoef_f64(const double* X, int64_t n_rows, int64_t ld, int32_t K, double* corr, int64_t ldc, double* work,
                          void* stream) {
    if (!X || !corr || !work || n_rows < 0 || K < 1 || ld < K || ldc < n_rows) {
        set_error("icnv_row_corrcoef_f64: bad argument");
        return ICNV_EINVAL;
    }
    if ((n_rows + 63) / 64 > 60000) {
        set_error("icnv_row_corrcoef_f64: group too large");
        return ICNV_EUNSUPPORTED;
    }
    return aux_row_corrcoef(X, n_rows, ld, K, corr, ldc, work, (cudaStream_t)stream);
}

int icnv_csr_to_dense_f32(const int64_t* indptr, const int32_t* indices, const void* data, int32_t data_is_f64, int64_t n_rows,
                          int32_t K, float* dense, int64_t ld, void* stream) {
    if (!indptr || !dense || ld < K) return ICNV_EINVAL;
    return graph_csr_to_dense(indptr, indices, data, data_is_f64 != 0, n_rows, K, dense, ld, (cudaStream_t)stream);
}
int icnv_gram_f32(const float* X, int64_t n_rows, int64_t ld, int32_t K, double* C, void* stream) {
    if (!X || !C || K < 1 || ld < K) return ICNV_EINVAL;
    return graph_gram(X, n_rows, ld, K, C, (cudaStream_t)stream);
}
int icnv_project_f32(const float* X, int64_t n_rows, int64_t ld, int32_t K, const double* V, int32_t n_comp, const double* mu,
                     float* Y, void* stream) {
    if (!X || !V || !Y || n_comp < 1) return ICNV_EINVAL;
    return graph_project(X, n_rows, ld, K, V, n_comp, mu, Y, (cudaStream_t)stream);
}
int64_t icnv_knn_workspace_bytes(int64_t n_all, int64_t nq) {
    if (n_all < 0 || nq < 0) return -1;
    return (int64_t)knn_workspace_bytes(n_all, nq);
}
int icnv_knn_f32(const float* P, int64_t n_all, int32_t d, int64_t ld, int64_t q0, int64_t nq, int32_t k, int32_t* knn_idx, float* knn_d2,
                 int32_t out_ld, void* workspace, void* stream) {
    if (!P || !knn_idx || !knn_d2 || !workspace || q0 < 0 || q0 + nq > n_all || d < 1 || ld < d || k < 1 || out_ld < k) {
        set_error("icnv_knn_f32: bad argument");
        return ICNV_EINVAL;
    }
    return knn_launch(P, n_all, d, ld, q0, nq, k, out_ld, knn_idx, knn_d2, workspace, (cudaStream_t)stream);
}
int icnv_fuzzy_rows(const float* dist, const int32_t* idx, int64_t n, int32_t k, int64_t row0, float mean_all, float* vals,
                    float* sigma, float* rho, void* stream) {
    if (!dist || !idx || !vals || !sigma || !rho || k < 2) return ICNV_EINVAL;
    return graph_fuzzy_rows(dist, idx, n, k, row0, mean_all, vals, sigma, rho, (cudaStream_t)stream);
}
int icnv_weighted_degree(const int64_t* indptr, const float* w, int64_t n, double* kdeg, void* stream) {
    if (!indptr || !kdeg) return ICNV_EINVAL;
    return graph_weighted_degree(indptr, w, n, kdeg, (cudaStream_t)stream);
}
int64_t icnv_community_sweep_work_bytes(int64_t n) { return n < 0 ? -1 : (int64_t)(n * 24 + 64); }
int icnv_community_sweep(const int64_t* indptr, const int32_t* indices, const float* w, const double* kdeg, const int32_t* comm,
                         const int32_t* bound, int64_t n, double two_m, double gamma, int32_t sweep, void* work, int32_t* comm_new,
                         int32_t* stats, void* stream) {
    if (!indptr || !kdeg || !comm || !work || !comm_new || !stats || !(two_m > 0)) {
        set_error("icnv_community_sweep: bad argument");
        return ICNV_EINVAL;
    }
    return graph_community_sweep(indptr, indices, w, kdeg, comm, bound, n, two_m, gamma, sweep, work, comm_new, stats, (cudaStream_t)stream);
}

int icnv_umap_epochs(const int32_t* head, const int32_t* tail, int64_t n_edges, float* emb, int32_t n_vertices,
                     const float* epochs_per_sample, float* next_sample, float* next_negative, float a, float b, float gamma,
                     float alpha0, int32_t n_epochs, int32_t epoch0, int32_t n_run, int32_t neg_rate, uint32_t seed, void* stream) {
    if (n_edges < 0 || n_vertices < 1 || n_epochs < 1 || neg_rate < 1 || !emb ||
        (n_edges > 0 && (!head || !tail || !epochs_per_sample || !next_sample || !next_negative))) {
        set_error("icnv_umap_epochs: bad argument");
        return ICNV_EINVAL;
    }
    return umap_epochs(head, tail, n_edges, emb, n_vertices, epochs_per_sample, next_sample, next_negative, a, b, gamma, alpha0, n_epochs,
                       epoch0, n_run, neg_rate, seed, (cudaStream_t)stream);
}
int icnv_tsne_affinities(const float* X, int32_t n, int32_t d, int64_t ld, float perplexity, float* P, void* stream) {
    if (!X || !P || n < 2 || d < 1 || d > 64 || ld < d || !(perplexity > 0.f) || perplexity >= (float)n) {
        set_error("icnv_tsne_affinities: bad argument (2 <= n, 1 <= d <= 64, 0 < perplexity < n)");
        return ICNV_EINVAL;
    }
    return tsne_affinities(X, n, d, ld, perplexity, P, (cudaStream_t)stream);
}
int64_t icnv_tsne_work_floats(int32_t n) { return n < 0 ? -1 : 5 * (int64_t)n + 4; }
int icnv_tsne_iterations(const float* P, float* Y, float* vel, float* gains, float* work, int32_t n, int32_t n_iter, float exaggeration,
                         float momentum, float learning_rate, void* stream) {
    if (!P || !Y || !vel || !gains || !work || n < 2 || n_iter < 0) {
        set_error("icnv_tsne_iterations: bad argument");
        return ICNV_EINVAL;
    }
    return tsne_iterations(P, Y, vel, gains, work, n, n_iter, exaggeration, momentum, learning_rate, (cudaStream_t)stream);
}

}  // extern "C"
