// General fall-back kernels: any gene-axis length, any (window, step), any output width.
//
//   smooth_direct_kernel   steps 1-3 of /root/reference/src/infercnvpy/tl/_infercnv.py:411-440 evaluated literally: the
//                          position-sorted, centred and clipped genes of a PART of the row are staged in shared memory
//                          (float, or double when numpy would centre in float64) and every output is the direct
//                          `window`-term pyramid sum (:206-212) or the flat mean (:227-236).  A part is a run of whole
//                          task tiles (32 tasks of <= LOUT outputs) whose genes fit in shared memory; windows never
//                          cross chromosomes and tasks are position-ordered, so a part's genes are one contiguous range
//                          of the sorted axis.  Rows longer than shared memory (the 58k-gene var of the reference's
//                          tutorial dataset) simply take several parts.
//   center_rows_wide_kernel  step 4 (:442) for rows wider than the warp-per-row kernel of icnv_aux.cu handles
//                          (K > 28 tiles, e.g. step = 1): one CTA per row, exact radix selection (icnv_select.cuh).
//
// Both write / read the same intermediate layout as the fast kernels (fp64 rows in warp-tile order + tile moments), so
// every later stage is shared.  They are correctness paths: gathers come straight from global memory (the row stays in
// L2), nothing is tuned.
#include "icnv_common.cuh"
#include "icnv_select.cuh"

namespace icnv {

template <bool BOUNDED, bool C64>
__global__ void __launch_bounds__(NT, 1) smooth_direct_kernel(const DirectParams p) {
    extern __shared__ __align__(16) unsigned char dsm[];
    double* wdir = reinterpret_cast<double*>(dsm);
    using TS = typename std::conditional<C64, double, float>::type;
    TS* buf = reinterpret_cast<TS*>(dsm + (size_t)p.window * 8);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < p.window; i += NT) wdir[i] = p.wdir[i];
    const int n_tiles = (p.n_tasks + 31) >> 5;
    const float clipf = p.clipf;

    for (int64_t row = blockIdx.x; row < p.n_rows; row += gridDim.x) {
        const float* xrow = p.X + row * p.ldx;
        double* orow = p.out + (size_t)row * p.ldo;
        for (int part = 0; part < p.n_parts; ++part) {
            const int4 pt = __ldg(p.parts + part);  // (first sorted gene, end, first tile, end tile)
            __syncthreads();                        // the previous part's readers are done with buf (and wdir is set)
            // ---- centre + clip the part's genes in position order (:422-436)
            for (int s = pt.x + tid; s < pt.y; s += NT) {
                const float x = __ldg(xrow + __ldg(p.idx_lin + s));
                if constexpr (C64) {
                    const double lo = reinterpret_cast<const double*>(p.lo_lin)[s];
                    double d;
                    if constexpr (BOUNDED) {
                        const double hi = reinterpret_cast<const double*>(p.hi_lin)[s];
                        // the bounded result is written into an array of the matrix dtype (:428): round to fp32
                        d = (double)x > hi ? (double)(float)((double)x - hi) : ((double)x < lo ? (double)(float)((double)x - lo) : 0.0);
                        d = (double)fminf(fmaxf((float)d, -clipf), clipf);
                    } else {
                        d = (double)x - lo;
                        d = fmin(fmax(d, -p.clip), p.clip);
                    }
                    buf[s - pt.x] = d;
                } else {
                    const float lo = reinterpret_cast<const float*>(p.lo_lin)[s];
                    float d;
                    if constexpr (BOUNDED) {
                        const float hi = reinterpret_cast<const float*>(p.hi_lin)[s];
                        d = x > hi ? x - hi : (x < lo ? x - lo : 0.f);
                    } else {
                        d = x - lo;
                    }
                    buf[s - pt.x] = fminf(fmaxf(d, -clipf), clipf);
                }
            }
            __syncthreads();
            // ---- one warp per tile of 32 tasks
            for (int tile = pt.z + warp; tile < pt.w; tile += NW) {
                const int ti = tile * 32 + lane;
                double v[LOUT];
#pragma unroll
                for (int i = 0; i < LOUT; ++i) v[i] = INFINITY;
                int nv = 0;
                if (ti < p.n_tasks) {
                    const int4 t = __ldg(reinterpret_cast<const int4*>(p.tasks) + ti);
                    const TS* B0 = buf + (t.x - pt.x);
                    if ((t.w & 0xFF) == 0) {
                        nv = t.z;
                        for (int i = 0; i < t.z; ++i) {
                            const TS* B = B0 + i * p.step;
                            double acc = 0.0;
                            for (int j = 0; j < p.window; ++j) acc = fma(wdir[j], (double)B[j], acc);
                            v[i] = acc * p.inv_sumw;
                        }
                    } else {
                        // chromosome not longer than the window: one flat mean (:227-236)
                        nv = 1;
                        double acc = 0.0;
                        for (int j = 0; j < t.z; ++j) acc += (double)B0[j];
                        v[0] = acc * p.flat_inv[t.w >> 8];
                    }
                }
                // tile moments (steer the median bracket of icnv_center_rows) + tile-order values
                double s1 = 0.0, s2 = 0.0;
#pragma unroll
                for (int i = 0; i < LOUT; ++i)
                    if (i < nv) {
                        s1 += v[i];
                        s2 = fma(v[i], v[i], s2);
                    }
                float f1 = (float)s1, f2 = (float)s2;
#pragma unroll
                for (int sh = 16; sh > 0; sh >>= 1) {
                    f1 += __shfl_xor_sync(0xffffffffu, f1, sh);
                    f2 += __shfl_xor_sync(0xffffffffu, f2, sh);
                }
                if (lane == 0) reinterpret_cast<float2*>(orow + (size_t)n_tiles * (32 * LOUT))[tile] = make_float2(f1, f2);
                double* o = orow + (size_t)tile * (32 * LOUT) + lane;
#pragma unroll
                for (int i = 0; i < LOUT; ++i) o[i * 32] = v[i];
            }
        }
    }
}

int direct_launch(const DirectParams& p, bool bounded, bool c64, int grid, size_t smem, cudaStream_t st) {
#define ICNV_DL(B, C)                                                                                              \
    do {                                                                                                           \
        ICNV_CUDA(cudaFuncSetAttribute(smooth_direct_kernel<B, C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        smooth_direct_kernel<B, C><<<grid, NT, smem, st>>>(p);                                                     \
    } while (0)
    if (bounded && c64)
        ICNV_DL(true, true);
    else if (bounded)
        ICNV_DL(true, false);
    else if (c64)
        ICNV_DL(false, true);
    else
        ICNV_DL(false, false);
#undef ICNV_DL
    ICNV_CUDA(cudaGetLastError());
    return 0;
}

// ------------------------------------------------------------------------------------------------
constexpr int CW_NT = 512;

template <typename TO>
__global__ void __launch_bounds__(CW_NT, 1) center_rows_wide_kernel(const double* __restrict__ tmp, int64_t n_rows, int64_t ld,
                                                                   const int32_t* __restrict__ kaddr, int K, int in_smem,
                                                                   TO* __restrict__ out, int64_t ldo,
                                                                   double* __restrict__ row_stats) {
    extern __shared__ __align__(16) unsigned char cw_smem[];
    __shared__ SelectSmem sel;
    __shared__ double red[2][CW_NT / 32];
    double* sv = reinterpret_cast<double*>(cw_smem);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int64_t row = blockIdx.x; row < n_rows; row += gridDim.x) {
        const double* srow = tmp + row * ld;
        if (in_smem)
            for (int k = tid; k < K; k += CW_NT) sv[k] = __ldg(srow + kaddr[k]);
        __syncthreads();
        double m;
        if (in_smem)
            m = cta_median<CW_NT>([&](int k) { return sv[k]; }, K, sel);
        else
            m = cta_median<CW_NT>([&](int k) { return __ldg(srow + kaddr[k]); }, K, sel);
        double s = 0.0, ss = 0.0;
        for (int k = tid; k < K; k += CW_NT) {
            const double c = (in_smem ? sv[k] : __ldg(srow + kaddr[k])) - m;
            s += c;
            ss = fma(c, c, ss);
            out[row * ldo + k] = (TO)c;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            s += __shfl_xor_sync(0xffffffffu, s, o);
            ss += __shfl_xor_sync(0xffffffffu, ss, o);
        }
        if (lane == 0) {
            red[0][warp] = s;
            red[1][warp] = ss;
        }
        __syncthreads();
        if (tid == 0) {
            double a = 0.0, b = 0.0;
            for (int w = 0; w < CW_NT / 32; ++w) {  // fixed order: deterministic
                a += red[0][w];
                b += red[1][w];
            }
            row_stats[2 * row] = a;
            row_stats[2 * row + 1] = b;
        }
        __syncthreads();  // sv / red are reused by the next row
    }
}

int center_wide_launch(const double* tmp, int64_t n_rows, int64_t ld, const int32_t* kaddr, int K, void* out, bool f64, int64_t ldo,
                       double* row_stats, int n_sm, cudaStream_t st) {
    if (n_rows == 0) return 0;
    const size_t budget = 200 * 1024;
    const int in_smem = (size_t)K * 8 <= budget;
    const size_t smem = in_smem ? (size_t)K * 8 : 16;
    const int grid = (int)(n_rows < n_sm ? n_rows : n_sm);
    if (f64) {
        ICNV_CUDA(cudaFuncSetAttribute(center_rows_wide_kernel<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        center_rows_wide_kernel<double><<<grid, CW_NT, smem, st>>>(tmp, n_rows, ld, kaddr, K, in_smem, (double*)out, ldo, row_stats);
    } else {
        ICNV_CUDA(cudaFuncSetAttribute(center_rows_wide_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        center_rows_wide_kernel<float><<<grid, CW_NT, smem, st>>>(tmp, n_rows, ld, kaddr, K, in_smem, (float*)out, ldo, row_stats);
    }
    ICNV_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace icnv
