// Auxiliary kernels of libicnv: reference profile (column sums), per-gene bound tables, per-chunk
// noise threshold, thresholding + row statistics, CSR compaction, cnv_score reductions.
// All are streaming / reduction kernels bound by HBM bandwidth; see DESIGN.md for bytes per unit.
#include "icnv_common.cuh"

namespace icnv {

// ---------------------------------------------------------------------------------------------
// K0: column sums per category.  /root/reference/src/infercnvpy/tl/_infercnv.py:385,400.
// grid = (column tiles, row splits, categories).  A thread owns 4 columns (one float4 when the
// matrix is 16-byte aligned) and walks its row range, skipping rows of other categories, with 4 rows
// in flight.  fp64 accumulation; partials [split][cat][G] are reduced in a fixed order afterwards so
// the result is run-to-run deterministic.
template <bool VEC>
__global__ void __launch_bounds__(256) colsum_dense_kernel(const float* __restrict__ X, int64_t n_rows, int64_t ldx,
                                                           int G, const int32_t* __restrict__ row_cat,
                                                           double* __restrict__ partial) {
    const int cat = blockIdx.z;
    const int n_split = gridDim.y;
    const int64_t rows_per = (n_rows + n_split - 1) / n_split;
    const int64_t r0 = (int64_t)blockIdx.y * rows_per;
    const int64_t r1 = min(n_rows, r0 + rows_per);
    double acc[4] = {0, 0, 0, 0};
    int col[4];
    if (VEC) {
        const int c0 = (blockIdx.x * 256 + threadIdx.x) * 4;
#pragma unroll
        for (int u = 0; u < 4; ++u) col[u] = c0 + u;
    } else {
#pragma unroll
        for (int u = 0; u < 4; ++u) col[u] = blockIdx.x * 1024 + u * 256 + threadIdx.x;
    }
    const bool any = col[0] < G;
    if (any) {
        int64_t r = r0;
        if (VEC && col[3] < G) {
            constexpr int UR = 8;  // rows in flight per thread
            for (; r + UR <= r1; r += UR) {
                float4 x[UR];
                bool use[UR];
#pragma unroll
                for (int k = 0; k < UR; ++k) {
                    use[k] = row_cat ? (row_cat[r + k] == cat) : true;
                    if (use[k]) x[k] = ldg_stream_f4(X + (r + k) * ldx + col[0]);
                }
#pragma unroll
                for (int k = 0; k < UR; ++k)
                    if (use[k]) {
                        acc[0] += (double)x[k].x;
                        acc[1] += (double)x[k].y;
                        acc[2] += (double)x[k].z;
                        acc[3] += (double)x[k].w;
                    }
            }
        }
        for (; r < r1; ++r) {
            if (row_cat && row_cat[r] != cat) continue;
#pragma unroll
            for (int u = 0; u < 4; ++u)
                if (col[u] < G) acc[u] += (double)__ldg(X + r * ldx + col[u]);
        }
    }
    double* dst = partial + ((size_t)blockIdx.y * gridDim.z + cat) * G;
#pragma unroll
    for (int u = 0; u < 4; ++u)
        if (col[u] < G) dst[col[u]] = acc[u];
}

// CSR variant: one warp per row chunk, atomics into a [cat][G] fp64 accumulator per split.
__global__ void __launch_bounds__(256) colsum_csr_kernel(const int64_t* __restrict__ indptr,
                                                         const int32_t* __restrict__ indices,
                                                         const float* __restrict__ data, int64_t n_rows, int G,
                                                         const int32_t* __restrict__ row_cat, int n_cat,
                                                         double* __restrict__ sums) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t r = warp; r < n_rows; r += n_warps) {
        const int cat = row_cat ? row_cat[r] : 0;
        if (cat < 0 || cat >= n_cat) continue;
        const int64_t e0 = indptr[r], e1 = indptr[r + 1];
        for (int64_t e = e0 + lane; e < e1; e += 32) atomicAdd(sums + (size_t)cat * G + indices[e], (double)data[e]);
    }
}

__global__ void reduce_partials_kernel(const double* __restrict__ partial, int n_split, int64_t n, double* __restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double s = 0.0;
    for (int k = 0; k < n_split; ++k) s += partial[(size_t)k * n + i];
    out[i] = s;
}

__global__ void count_rows_kernel(const int32_t* __restrict__ row_cat, int64_t n_rows, int n_cat,
                                  unsigned long long* __restrict__ counts) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int cat = -1;
    if (i < n_rows) cat = row_cat ? row_cat[i] : 0;
    for (int c = 0; c < n_cat; ++c) {
        const unsigned m = __ballot_sync(0xffffffffu, cat == c);
        if ((threadIdx.x & 31) == 0 && m) atomicAdd(counts + c, (unsigned long long)__popc(m));
    }
}

template <typename T>
__global__ void mean_from_sums_kernel(const double* __restrict__ sums, const int64_t* __restrict__ counts, int n_cat, int G,
                                      T* __restrict__ ref) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int64_t)n_cat * G) return;
    const int c = (int)(i / G);
    ref[i] = (T)(sums[i] / (double)counts[c]);
}

// row_nnz -> CSR indptr: single CTA, chunked inclusive scan (n_rows is at most a few million)
__global__ void __launch_bounds__(1024) nnz_to_indptr_kernel(const int32_t* __restrict__ row_nnz, int64_t n_rows,
                                                             int64_t* __restrict__ indptr) {
    __shared__ long long warp_tot[32];
    __shared__ long long carry_s;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) {
        carry_s = 0;
        indptr[0] = 0;
    }
    __syncthreads();
    for (int64_t base = 0; base < n_rows; base += 1024) {
        const int64_t i = base + threadIdx.x;
        long long v = i < n_rows ? (long long)row_nnz[i] : 0;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const long long t = __shfl_up_sync(0xffffffffu, v, o);
            if (lane >= o) v += t;
        }
        if (lane == 31) warp_tot[warp] = v;
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
        const long long carry = carry_s;
        const long long incl = v + (warp > 0 ? warp_tot[warp - 1] : 0) + carry;
        if (i < n_rows) indptr[i + 1] = incl;
        __syncthreads();
        if (threadIdx.x == 1023) carry_s = incl;
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------------
// Per-gene centring bounds in the layouts the smoothing kernel reads.
// n_cat == 1 -> lo == hi == ref; otherwise min / max over the reference rows
// (/root/reference/src/infercnvpy/tl/_infercnv.py:425-426).
template <typename TR, typename TO>
__global__ void build_bounds_kernel(const TR* __restrict__ ref, int n_cat, int G, const int32_t* __restrict__ cols,
                                    int64_t n, TO* __restrict__ lo, TO* __restrict__ hi) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int c = cols[i];
    TO l = 0, h = 0;
    if (c >= 0 && c < G) {
        TR mn = ref[c], mx = ref[c];
        for (int k = 1; k < n_cat; ++k) {
            const TR v = ref[(size_t)k * G + c];
            mn = v < mn ? v : mn;
            mx = v > mx ? v : mx;
        }
        l = (TO)mn;
        h = (TO)mx;
    }
    lo[i] = l;
    hi[i] = h;
}

// ---------------------------------------------------------------------------------------------
// Step 5 threshold: /root/reference/src/infercnvpy/tl/_infercnv.py:450 — np.std over every element of
// a chunk of `chunk_rows` rows.  One CTA per chunk, fixed-order tree => deterministic.
__global__ void __launch_bounds__(256) chunk_threshold_kernel(const double* __restrict__ row_stats, int64_t n_rows,
                                                              int64_t K, int64_t chunk_rows, double dyn,
                                                              double* __restrict__ thr) {
    __shared__ double sh[2][256];
    const int64_t r0 = (int64_t)blockIdx.x * chunk_rows;
    const int64_t r1 = min(n_rows, r0 + chunk_rows);
    double s = 0.0, ss = 0.0;
    for (int64_t r = r0 + threadIdx.x; r < r1; r += 256) {
        s += row_stats[2 * r];
        ss += row_stats[2 * r + 1];
    }
    sh[0][threadIdx.x] = s;
    sh[1][threadIdx.x] = ss;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) {
            sh[0][threadIdx.x] += sh[0][threadIdx.x + o];
            sh[1][threadIdx.x] += sh[1][threadIdx.x + o];
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        const double n = (double)(r1 - r0) * (double)K;
        const double mean = sh[0][0] / n;
        const double var = fmax(sh[1][0] / n - mean * mean, 0.0);
        thr[blockIdx.x] = dyn * sqrt(var);
    }
}

// Step 5 proper: read the smoothing kernel's intermediate (warp-tile order, see SmoothParams::out), zero
// |v| < thr (strict, :451), write the matrix in natural column order and emit per-row sum|v| and nnz.
// One warp per row; a tile's 32 tasks cover one contiguous column range, so the un-permute goes through a
// 32*LOUT-value staging slab per warp and both the read and the write are fully coalesced.
template <typename T>
__global__ void __launch_bounds__(256) finalize_kernel(const T* __restrict__ tmp, int64_t n_rows, int64_t ld_tmp,
                                                       const Task* __restrict__ tasks, int n_tasks, int64_t chunk_rows,
                                                       const double* __restrict__ thr, T* __restrict__ out, int64_t ldo,
                                                       double* __restrict__ row_abs, int32_t* __restrict__ row_nnz) {
    __shared__ T stage[8][32 * LOUT];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int n_tiles = (n_tasks + 31) >> 5;
    for (int64_t r = warp; r < n_rows; r += n_warps) {
        const double t = thr ? thr[r / chunk_rows] : -1.0;
        double a = 0.0;
        int nz = 0;
        for (int tile = 0; tile < n_tiles; ++tile) {
            const int ti = tile * 32 + lane;
            int col0 = 0, cnt = 0;
            if (ti < n_tasks) {
                const int4 tk = __ldg(reinterpret_cast<const int4*>(tasks) + ti);
                col0 = tk.y;
                cnt = (tk.w & 0xFF) ? 1 : tk.z;
            }
            const int base = __shfl_sync(0xffffffffu, col0, 0);
            const int end = __reduce_max_sync(0xffffffffu, col0 + cnt);
            const T* src = tmp + r * ld_tmp + (int64_t)tile * (32 * LOUT) + lane;
#pragma unroll
            for (int i = 0; i < LOUT; ++i) {
                T x = src[i * 32];
                if (i < cnt) {
                    const double av = fabs((double)x);
                    if (av < t) {
                        x = (T)0;
                    } else if (av != 0.0) {
                        a += av;
                        nz += 1;
                    }
                    stage[wib][col0 - base + i] = x;
                }
            }
            __syncwarp();
            T* dst = out + r * ldo + base;
            for (int k = lane; k < end - base; k += 32) dst[k] = stage[wib][k];
            __syncwarp();
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

// Dense -> CSR (:455): warp per row, ballot compaction keeps column order.
template <typename T>
__global__ void __launch_bounds__(256) dense_to_csr_kernel(const T* __restrict__ out, int64_t n_rows, int64_t K, int64_t ldo,
                                                           const int64_t* __restrict__ indptr, int32_t* __restrict__ indices,
                                                           T* __restrict__ data) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t r = warp; r < n_rows; r += n_warps) {
        const T* row = out + r * ldo;
        int64_t base = indptr[r];
        for (int64_t c0 = 0; c0 < K; c0 += 32) {
            const int64_t c = c0 + lane;
            const T v = c < K ? row[c] : (T)0;
            const bool nz = v != (T)0;
            const unsigned m = __ballot_sync(0xffffffffu, nz);
            if (nz) {
                const int64_t pos = base + __popc(m & ((1u << lane) - 1u));
                indices[pos] = (int32_t)c;
                data[pos] = v;
            }
            base += __popc(m);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// cnv_score pieces: /root/reference/src/infercnvpy/tl/_scores.py:65-68
template <typename T>
__global__ void __launch_bounds__(256) rowabs_csr_kernel(const int64_t* __restrict__ indptr, const T* __restrict__ data,
                                                         int64_t n_rows, double* __restrict__ row_abs) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t r = warp; r < n_rows; r += n_warps) {
        double a = 0.0;
        for (int64_t e = indptr[r] + lane; e < indptr[r + 1]; e += 32) a += fabs((double)data[e]);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
        if (lane == 0) row_abs[r] = a;
    }
}
template <typename T>
__global__ void __launch_bounds__(256) rowabs_dense_kernel(const T* __restrict__ X, int64_t n_rows, int64_t K, int64_t ld,
                                                           double* __restrict__ row_abs) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t r = warp; r < n_rows; r += n_warps) {
        double a = 0.0;
        for (int64_t c = lane; c < K; c += 32) a += fabs((double)X[r * ld + c]);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
        if (lane == 0) row_abs[r] = a;
    }
}
// one CTA per label, fixed-order tree => deterministic
__global__ void __launch_bounds__(256) label_sums_kernel(const double* __restrict__ row_abs, const int32_t* __restrict__ labels,
                                                         int64_t n_rows, double* __restrict__ label_sum,
                                                         int64_t* __restrict__ label_rows) {
    __shared__ double sh[256];
    __shared__ long long shn[256];
    const int lab = blockIdx.x;
    double s = 0.0;
    long long n = 0;
    for (int64_t r = threadIdx.x; r < n_rows; r += 256)
        if (labels[r] == lab) {
            s += row_abs[r];
            n += 1;
        }
    sh[threadIdx.x] = s;
    shn[threadIdx.x] = n;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) {
            sh[threadIdx.x] += sh[threadIdx.x + o];
            shn[threadIdx.x] += shn[threadIdx.x + o];
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        label_sum[lab] = sh[0];
        label_rows[lab] = shn[0];
    }
}

// ---------------------------------------------------------------------------------------------
// host-side launch wrappers (called from icnv_api.cu)
static int grid_for_rows(int64_t n_rows) {
    int64_t g = (n_rows + 7) / 8;  // 8 warps per 256-thread CTA
    if (g < 1) g = 1;
    if (g > 148 * 8) g = 148 * 8;
    return (int)g;
}

int aux_colsum_dense(const float* X, int64_t n_rows, int64_t ldx, int G, const int32_t* row_cat, int n_cat, double* sums,
                     int64_t* counts, double* partial, int n_split, cudaStream_t st) {
    const bool vec = (ldx % 4 == 0) && ((reinterpret_cast<uintptr_t>(X) & 15) == 0);
    dim3 grid((G + 1023) / 1024, n_split, n_cat);
    if (vec)
        colsum_dense_kernel<true><<<grid, 256, 0, st>>>(X, n_rows, ldx, G, row_cat, partial);
    else
        colsum_dense_kernel<false><<<grid, 256, 0, st>>>(X, n_rows, ldx, G, row_cat, partial);
    ICNV_CUDA(cudaGetLastError());
    const int64_t n = (int64_t)n_cat * G;
    reduce_partials_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(partial, n_split, n, sums);
    ICNV_CUDA(cudaGetLastError());
    ICNV_CUDA(cudaMemsetAsync(counts, 0, sizeof(int64_t) * n_cat, st));
    count_rows_kernel<<<(unsigned)((n_rows + 255) / 256), 256, 0, st>>>(row_cat, n_rows, n_cat,
                                                                      reinterpret_cast<unsigned long long*>(counts));
    ICNV_CUDA(cudaGetLastError());
    return 0;
}

int aux_colsum_csr(const int64_t* indptr, const int32_t* indices, const float* data, int64_t n_rows, int G,
                   const int32_t* row_cat, int n_cat, double* sums, int64_t* counts, cudaStream_t st) {
    ICNV_CUDA(cudaMemsetAsync(sums, 0, sizeof(double) * (size_t)n_cat * G, st));
    colsum_csr_kernel<<<grid_for_rows(n_rows), 256, 0, st>>>(indptr, indices, data, n_rows, G, row_cat, n_cat, sums);
    ICNV_CUDA(cudaGetLastError());
    ICNV_CUDA(cudaMemsetAsync(counts, 0, sizeof(int64_t) * n_cat, st));
    count_rows_kernel<<<(unsigned)((n_rows + 255) / 256), 256, 0, st>>>(row_cat, n_rows, n_cat,
                                                                      reinterpret_cast<unsigned long long*>(counts));
    ICNV_CUDA(cudaGetLastError());
    return 0;
}

int aux_mean_from_sums(const double* sums, const int64_t* counts, int n_cat, int G, void* ref, bool f64, cudaStream_t st) {
    const int64_t n = (int64_t)n_cat * G;
    if (f64)
        mean_from_sums_kernel<double><<<(unsigned)((n + 255) / 256), 256, 0, st>>>(sums, counts, n_cat, G, (double*)ref);
    else
        mean_from_sums_kernel<float><<<(unsigned)((n + 255) / 256), 256, 0, st>>>(sums, counts, n_cat, G, (float*)ref);
    ICNV_CUDA(cudaGetLastError());
    return 0;
}
int aux_nnz_to_indptr(const int32_t* row_nnz, int64_t n_rows, int64_t* indptr, cudaStream_t st) {
    nnz_to_indptr_kernel<<<1, 1024, 0, st>>>(row_nnz, n_rows, indptr);
    ICNV_CUDA(cudaGetLastError());
    return 0;
}

int aux_build_bounds(const void* ref, bool ref_f64, int n_cat, int G, const int32_t* cols, int64_t n, void* lo, void* hi,
                     bool out_f64, cudaStream_t st) {
    const unsigned grid = (unsigned)((n + 255) / 256);
    if (n == 0) return 0;
    if (ref_f64 && out_f64)
        build_bounds_kernel<double, double><<<grid, 256, 0, st>>>((const double*)ref, n_cat, G, cols, n, (double*)lo, (double*)hi);
    else if (ref_f64 && !out_f64)
        build_bounds_kernel<double, float><<<grid, 256, 0, st>>>((const double*)ref, n_cat, G, cols, n, (float*)lo, (float*)hi);
    else if (!ref_f64 && out_f64)
        build_bounds_kernel<float, double><<<grid, 256, 0, st>>>((const float*)ref, n_cat, G, cols, n, (double*)lo, (double*)hi);
    else
        build_bounds_kernel<float, float><<<grid, 256, 0, st>>>((const float*)ref, n_cat, G, cols, n, (float*)lo, (float*)hi);
    ICNV_CUDA(cudaGetLastError());
    return 0;
}

int aux_chunk_threshold(const double* row_stats, int64_t n_rows, int64_t K, int64_t chunk_rows, double dyn, double* thr,
                        cudaStream_t st) {
    const int64_t n_chunks = (n_rows + chunk_rows - 1) / chunk_rows;
    if (n_chunks == 0) return 0;
    chunk_threshold_kernel<<<(unsigned)n_chunks, 256, 0, st>>>(row_stats, n_rows, K, chunk_rows, dyn, thr);
    ICNV_CUDA(cudaGetLastError());
    return 0;
}

int aux_finalize(const void* tmp, bool f64, int64_t n_rows, int64_t ld_tmp, const Task* tasks, int n_tasks, int64_t chunk_rows,
                 const double* thr, void* out, int64_t ldo, double* row_abs, int32_t* row_nnz, cudaStream_t st) {
    if (n_rows == 0) return 0;
    if (f64)
        finalize_kernel<double><<<grid_for_rows(n_rows), 256, 0, st>>>((const double*)tmp, n_rows, ld_tmp, tasks, n_tasks, chunk_rows, thr,
                                                                      (double*)out, ldo, row_abs, row_nnz);
    else
        finalize_kernel<float><<<grid_for_rows(n_rows), 256, 0, st>>>((const float*)tmp, n_rows, ld_tmp, tasks, n_tasks, chunk_rows, thr,
                                                                     (float*)out, ldo, row_abs, row_nnz);
    ICNV_CUDA(cudaGetLastError());
    return 0;
}

int aux_dense_to_csr(const void* out, bool f64, int64_t n_rows, int64_t K, int64_t ldo, const int64_t* indptr, int32_t* indices,
                     void* data, cudaStream_t st) {
    if (n_rows == 0) return 0;
    if (f64)
        dense_to_csr_kernel<double><<<grid_for_rows(n_rows), 256, 0, st>>>((const double*)out, n_rows, K, ldo, indptr, indices, (double*)data);
    else
        dense_to_csr_kernel<float><<<grid_for_rows(n_rows), 256, 0, st>>>((const float*)out, n_rows, K, ldo, indptr, indices, (float*)data);
    ICNV_CUDA(cudaGetLastError());
    return 0;
}

int aux_rowabs_csr(const int64_t* indptr, const void* data, bool f64, int64_t n_rows, double* row_abs, cudaStream_t st) {
    if (n_rows == 0) return 0;
    if (f64)
        rowabs_csr_kernel<double><<<grid_for_rows(n_rows), 256, 0, st>>>(indptr, (const double*)data, n_rows, row_abs);
    else
        rowabs_csr_kernel<float><<<grid_for_rows(n_rows), 256, 0, st>>>(indptr, (const float*)data, n_rows, row_abs);
    ICNV_CUDA(cudaGetLastError());
    return 0;
}
int aux_rowabs_dense(const void* X, bool f64, int64_t n_rows, int64_t K, int64_t ld, double* row_abs, cudaStream_t st) {
    if (n_rows == 0) return 0;
    if (f64)
        rowabs_dense_kernel<double><<<grid_for_rows(n_rows), 256, 0, st>>>((const double*)X, n_rows, K, ld, row_abs);
    else
        rowabs_dense_kernel<float><<<grid_for_rows(n_rows), 256, 0, st>>>((const float*)X, n_rows, K, ld, row_abs);
    ICNV_CUDA(cudaGetLastError());
    return 0;
}
int aux_label_sums(const double* row_abs, const int32_t* labels, int64_t n_rows, int n_labels, double* label_sum,
                   int64_t* label_rows, cudaStream_t st) {
    if (n_labels <= 0) return 0;
    label_sums_kernel<<<n_labels, 256, 0, st>>>(row_abs, labels, n_rows, label_sum, label_rows);
    ICNV_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace icnv
