// Per-gene CNV layer of calculate_gene_values=True
// (/root/reference/src/infercnvpy/tl/_infercnv.py:141-151, :214-223, :238-242, :247-291, :443-444, :452-453).
//
// For every cell row, from the smoothed (not yet centred) window values the smoothing kernel left in `tmp`:
//   value(gene) = np.mean of the kept windows that contain the gene (:278-287), the flat mean for chromosomes not
//                 longer than the window (:240); genes no kept window covers and masked genes are NaN (:146);
//   the layer is centred on ITS OWN row median (:444) and filtered with the WINDOW matrix's chunk threshold (:453).
// One CTA per row (persistent, grid-stride).  The window values (K doubles) and the per-gene means (n_cov doubles) live
// in shared memory when they fit (bench shape: 14 KB + 160 KB), otherwise the means go to an L2-resident scratch row
// per CTA and the windows are read through the tile-order address table.  The median is exact: a histogram bracket around
// the row mean + exact ranking of the middle bin (icnv_select.cuh, cta_median_hist) when shared memory has room for its
// work area, else the radix selection on the order-preserving 64-bit pattern; the final sweep walks the natural gene columns
// so every store is coalesced (streaming, fp64) — the layer is a dense [n_rows, n_genes] float64 matrix (160 KB per cell
// at 20k genes), which is what bounds this kernel: 8 * n_genes bytes written per cell.
#include "icnv_common.cuh"
#include "icnv_select.cuh"

namespace icnv {

constexpr int GV_NT = 512;


// numpy's add.reduce over a contiguous float64 vector (pairwise summation, blocks of 128, 8 accumulators), which is what
// np.mean of the per-gene list runs (:287): identical operation order => identical bits.
template <typename F>
__device__ double np_pairwise_block(F get, int a0, int n) {
    if (n < 8) {
        double res = 0.0;
        for (int i = 0; i < n; ++i) res += get(a0 + i);
        return res;
    }
    double r[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) r[j] = get(a0 + j);
    int i = 8;
    for (; i < n - (n % 8); i += 8) {
#pragma unroll
        for (int j = 0; j < 8; ++j) r[j] += get(a0 + i + j);
    }
    double res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
    for (; i < n; ++i) res += get(a0 + i);
    return res;
}
template <typename F>
__device__ double np_pairwise_sum(F get, int a0, int n) {
    if (n <= 128) return np_pairwise_block(get, a0, n);
    // explicit post-order walk of numpy's recursion: split at n/2 rounded down to a multiple of 8
    struct Frame {
        int a0, n, stage;
        double left;
    };
    Frame st[28];
    int sp = 0;
    st[0] = {a0, n, 0, 0.0};
    double ret = 0.0;
    while (sp >= 0) {
        Frame& f = st[sp];
        if (f.n <= 128) {
            ret = np_pairwise_block(get, f.a0, f.n);
            --sp;
            continue;
        }
        int n2 = f.n / 2;
        n2 -= n2 % 8;
        if (f.stage == 0) {
            f.stage = 1;
            st[sp + 1] = {f.a0, n2, 0, 0.0};
            ++sp;
        } else if (f.stage == 1) {
            f.left = ret;
            f.stage = 2;
            st[sp + 1] = {f.a0 + n2, f.n - n2, 0, 0.0};
            ++sp;
        } else {
            ret = f.left + ret;
            --sp;
        }
    }
    return ret;
}

__global__ void __launch_bounds__(GV_NT, 1) gene_values_kernel(const GeneValParams p) {
    extern __shared__ __align__(16) unsigned char gv_smem[];
    __shared__ SelectSmem sel;
    __shared__ HistSmem hsel;

    double* sk = p.k_in_smem ? reinterpret_cast<double*>(gv_smem) : nullptr;
    double* sv = p.v_in_smem ? reinterpret_cast<double*>(gv_smem) + (p.k_in_smem ? p.K : 0)
                             : p.scratch + (size_t)blockIdx.x * p.n_cov;
    const int tid = threadIdx.x, lane = tid & 31;
    const int n = p.n_cov;

    for (int64_t row = blockIdx.x; row < p.n_rows; row += gridDim.x) {
        const double* srow = p.tmp + row * p.ld;
        // ---- window values in natural column order
        if (sk)
            for (int k = tid; k < p.K; k += GV_NT) sk[k] = __ldg(srow + p.kaddr[k]);
        __syncthreads();
        // ---- per-gene means (position order, covered genes only)
        for (int i = tid; i < n; i += GV_NT) {
            const int f = __ldg(p.first + i), c = __ldg(p.cnt + i);
            double s;
            if (sk) {
                s = np_pairwise_sum([&](int k) { return sk[k]; }, f, c);
            } else {
                s = np_pairwise_sum([&](int k) { return __ldg(srow + p.kaddr[k]); }, f, c);
            }
            sv[i] = s / (double)c;
        }
        __syncthreads();  // (global scratch: writes by this CTA are visible to it after the barrier)

        // ---- exact np.median of the covered genes (:444)
        const double m = p.hist_off >= 0 ? cta_median_hist<GV_NT>([&](int i) { return sv[i]; }, n, sel, hsel, gv_smem + p.hist_off)
                                         : cta_median<GV_NT>([&](int i) { return sv[i]; }, n, sel);

        // ---- centre, filter, natural-order write
        const double t = p.thr ? p.thr[row / p.chunk_rows] : -1.0;
        double* orow = p.out + row * p.ldo;
        const double nan = __longlong_as_double(0x7FF8000000000000ll);
        for (int g = tid; g < p.G; g += GV_NT) {
            const int i = __ldg(p.inv + g);
            double c = nan;
            if (i >= 0) {
                c = sv[i] - m;
                if (fabs(c) < t) c = 0.0;
            }
            __stcs(orow + g, c);
        }
        __syncthreads();  // sk / sv / selection state are reused by the next row
    }
}

int genevals_launch(const GeneValParams& p, int grid, size_t smem, cudaStream_t st) {
    ICNV_CUDA(cudaFuncSetAttribute(gene_values_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    gene_values_kernel<<<grid, GV_NT, smem, st>>>(p);
    ICNV_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace icnv
