// Noise filter + CSR compaction of /root/reference/src/infercnvpy/tl/_infercnv.py:449-455 without rewriting the
// dense matrix:
//   filter_count_kernel   x[abs(x) < thr] = 0 (:451) is only COUNTED: per row nnz and sum|x| of what survives
//   indptr_scan_kernel    row counts -> CSR indptr (multi-CTA, one launch, fixed order)
//   filter_to_csr_kernel  the same predicate again while compacting the row into (indices, data) (:455)
// Against apply_threshold + dense_to_csr (icnv_aux.cu) this drops one write and one read of the [n, K] matrix:
// per cell 2 reads of 4 K bytes + 12 bytes per surviving value.  The dense matrix keeps its UNFILTERED values.
#include "icnv_common.cuh"

namespace icnv {

namespace {

constexpr int FW = 8;  // warps per CTA, one row per warp

__device__ __forceinline__ float ld_keep(const float* p) { return __ldg(p); }
__device__ __forceinline__ double ld_keep(const double* p) { return __ldg(p); }

// survives the filter: not zero and not below the chunk's threshold (strict <, NaN thresholds keep everything)
template <typename T>
__device__ __forceinline__ bool keeps(T v, double t) {
    const double av = fabs((double)v);
    return !(av < t) && av != 0.0;
}

template <typename T>
__global__ void __launch_bounds__(32 * FW) filter_count_kernel(const T* __restrict__ out, int64_t n_rows, int64_t K, int64_t ldo,
                                                               int64_t chunk_rows, const double* __restrict__ thr,
                                                               double* __restrict__ row_abs, int32_t* __restrict__ row_nnz) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    constexpr int U = 8;
    for (int64_t r = warp; r < n_rows; r += n_warps) {
        const double t = thr ? thr[r / chunk_rows] : -1.0;
        const T* row = out + r * ldo;
        double a = 0.0;
        int nz = 0;
        int64_t c = lane;
        for (; c + (U - 1) * 32 < K; c += U * 32) {
            T v[U];
#pragma unroll
            for (int u = 0; u < U; ++u) v[u] = ld_keep(row + c + u * 32);
#pragma unroll
            for (int u = 0; u < U; ++u)
                if (keeps(v[u], t)) {
                    a += fabs((double)v[u]);
                    nz += 1;
                }
        }
        for (; c < K; c += 32) {
            const T v = ld_keep(row + c);
            if (keeps(v, t)) {
                a += fabs((double)v);
                nz += 1;
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            a += __shfl_xor_sync(0xffffffffu, a, o);
            nz += __shfl_xor_sync(0xffffffffu, nz, o);
        }
        if (lane == 0) {
            if (row_abs) row_abs[r] = a;
            if (row_nnz) row_nnz[r] = nz;
        }
    }
}

// indptr[i + 1] = sum of row_nnz[0..i].  CTA b owns rows [b * SCAN_ROWS, (b + 1) * SCAN_ROWS): it first adds up every count
// before its block (coalesced, from L2: at most 4 bytes per row of the whole matrix) and then scans its own block, so one
// launch of independent CTAs suffices and the result does not depend on scheduling.
constexpr int SCAN_T = 1024, SCAN_PER = 8, SCAN_ROWS = SCAN_T * SCAN_PER;

__global__ void __launch_bounds__(SCAN_T) indptr_scan_kernel(const int32_t* __restrict__ row_nnz, int64_t n_rows, int64_t* __restrict__ indptr) {
    __shared__ long long warp_tot[32];
    __shared__ long long base_s;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t r0 = (int64_t)blockIdx.x * SCAN_ROWS;
    // ---- everything before this block
    long long before = 0;
    for (int64_t i = threadIdx.x; i < r0; i += SCAN_T) before += (long long)__ldg(row_nnz + i);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) before += __shfl_xor_sync(0xffffffffu, before, o);
    if (lane == 0) warp_tot[warp] = before;
    __syncthreads();
    if (warp == 0) {
        long long w = warp_tot[lane];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) w += __shfl_xor_sync(0xffffffffu, w, o);
        if (lane == 0) base_s = w;
    }
    __syncthreads();
    const long long base = base_s;
    __syncthreads();
    // ---- own block: thread t owns rows r0 + t * SCAN_PER .. + SCAN_PER - 1
    const int64_t first = r0 + (int64_t)threadIdx.x * SCAN_PER;
    long long v[SCAN_PER];
    long long run = 0;
#pragma unroll
    for (int k = 0; k < SCAN_PER; ++k) {
        const int64_t i = first + k;
        run += i < n_rows ? (long long)__ldg(row_nnz + i) : 0;
        v[k] = run;
    }
    long long incl = run;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const long long t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) warp_tot[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        long long w = warp_tot[lane];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const long long t = __shfl_up_sync(0xffffffffu, w, o);
            if (lane >= o) w += t;
        }
        warp_tot[lane] = w;
    }
    __syncthreads();
    const long long excl = base + (incl - run) + (warp > 0 ? warp_tot[warp - 1] : 0);
#pragma unroll
    for (int k = 0; k < SCAN_PER; ++k) {
        const int64_t i = first + k;
        if (i < n_rows) indptr[i + 1] = excl + v[k];
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) indptr[0] = 0;
}

// One warp per row: the row is taken in slabs of 32 x U columns whose loads are issued together; every 32-column group
// is compacted with a ballot, so the column order is kept.
template <typename T, typename TD>
__global__ void __launch_bounds__(32 * FW) filter_to_csr_kernel(const T* __restrict__ out, int64_t n_rows, int64_t K, int64_t ldo,
                                                                int64_t chunk_rows, const double* __restrict__ thr,
                                                                const int64_t* __restrict__ indptr, int32_t* __restrict__ indices,
                                                                TD* __restrict__ data) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const unsigned below = (1u << lane) - 1u;
    constexpr int U = 8;
    for (int64_t r = warp; r < n_rows; r += n_warps) {
        const double t = thr ? thr[r / chunk_rows] : -1.0;
        const T* row = out + r * ldo;
        int64_t base = __ldg(indptr + r);
        if (__ldg(indptr + r + 1) == base) continue;  // nothing survives in this row
        for (int64_t c0 = 0; c0 < K; c0 += U * 32) {
            T v[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int64_t c = c0 + u * 32 + lane;
                v[u] = c < K ? ld_keep(row + c) : (T)0;
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const bool k = keeps(v[u], t);
                const unsigned m = __ballot_sync(0xffffffffu, k);
                if (k) {
                    const int64_t pos = base + __popc(m & below);
                    indices[pos] = (int32_t)(c0 + u * 32 + lane);
                    data[pos] = (TD)v[u];
                }
                base += __popc(m);
            }
        }
    }
}

int grid_for(int64_t n_rows) {
    int64_t g = (n_rows + FW - 1) / FW;
    if (g < 1) g = 1;
    if (g > 148 * 8) g = 148 * 8;
    return (int)g;
}

}  // namespace

int filter_count(const void* out, bool f64, int64_t n_rows, int64_t K, int64_t ldo, int64_t chunk_rows, const double* thr,
                 double* row_abs, int32_t* row_nnz, cudaStream_t st) {
    if (n_rows == 0) return 0;
    if (f64)
        filter_count_kernel<double><<<grid_for(n_rows), 32 * FW, 0, st>>>((const double*)out, n_rows, K, ldo, chunk_rows, thr, row_abs, row_nnz);
    else
        filter_count_kernel<float><<<grid_for(n_rows), 32 * FW, 0, st>>>((const float*)out, n_rows, K, ldo, chunk_rows, thr, row_abs, row_nnz);
    ICNV_CUDA(cudaGetLastError());
    return 0;
}

int indptr_scan(const int32_t* row_nnz, int64_t n_rows, int64_t* indptr, cudaStream_t st) {
    const int64_t grid = n_rows > 0 ? (n_rows + SCAN_ROWS - 1) / SCAN_ROWS : 1;
    indptr_scan_kernel<<<(unsigned)grid, SCAN_T, 0, st>>>(row_nnz, n_rows, indptr);
    ICNV_CUDA(cudaGetLastError());
    return 0;
}

int filter_to_csr(const void* out, bool f64, int64_t n_rows, int64_t K, int64_t ldo, int64_t chunk_rows, const double* thr,
                  const int64_t* indptr, int32_t* indices, void* data, bool data_f64, cudaStream_t st) {
    if (n_rows == 0) return 0;
    const int g = grid_for(n_rows);
    if (f64) {
        if (!data_f64) {
            set_error("icnv_filter_to_csr: float64 matrix needs float64 data");
            return -1;
        }
        filter_to_csr_kernel<double, double><<<g, 32 * FW, 0, st>>>((const double*)out, n_rows, K, ldo, chunk_rows, thr, indptr, indices, (double*)data);
    } else if (data_f64) {
        filter_to_csr_kernel<float, double><<<g, 32 * FW, 0, st>>>((const float*)out, n_rows, K, ldo, chunk_rows, thr, indptr, indices, (double*)data);
    } else {
        filter_to_csr_kernel<float, float><<<g, 32 * FW, 0, st>>>((const float*)out, n_rows, K, ldo, chunk_rows, thr, indptr, indices, (float*)data);
    }
    ICNV_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace icnv
