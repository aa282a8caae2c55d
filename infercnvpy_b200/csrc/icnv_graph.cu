// Kernels behind cnv.tl.pca / cnv.pp.neighbors / cnv.tl.leiden on the CNV matrix.
// The reference delegates these three steps to scanpy -> scikit-learn / pynndescent / umap-learn /
// leidenalg (/root/reference/src/infercnvpy/tl/__init__.py:13-75, pp/__init__.py:8-43); none of that
// arithmetic is in the reference tree and its tests pin nothing there, so parity for
// this file is "unpinned": the tests compare against scikit-learn / scipy restatements.
#include "icnv_common.cuh"

namespace icnv {

// ---------------------------------------------------------------------------------------------
// CSR (float32 or float64 values) -> dense float32 [n, ld]
template <typename T>
__global__ void __launch_bounds__(256) csr_to_dense_kernel(const int64_t* __restrict__ indptr, const int32_t* __restrict__ indices,
                                                           const T* __restrict__ data, int64_t n_rows, int K,
                                                           float* __restrict__ dense, int64_t ld) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t r = warp; r < n_rows; r += n_warps) {
        float* row = dense + r * ld;
        for (int c = lane; c < K; c += 32) row[c] = 0.f;
        __syncwarp();
        for (int64_t e = indptr[r] + lane; e < indptr[r + 1]; e += 32) row[indices[e]] = (float)data[e];
    }
}

// ---------------------------------------------------------------------------------------------
// Gram matrix C = X^T X (upper-triangular 64x64 tiles, fp64 accumulation on B200's 1:2 fp64 pipe).
// TruncatedSVD(arpack) of the reference works on X directly; the top eigenpairs of C are the same
// right singular subspace (sigma^2, V) and fp64 keeps the squared condition number harmless.
constexpr int GT = 64;  // tile edge
constexpr int GR = 16;  // rows per shared-memory chunk
__global__ void __launch_bounds__(256) gram_kernel(const float* __restrict__ X, int64_t n_rows, int64_t ld, int K,
                                                   double* __restrict__ C) {
    __shared__ double As[GR][GT], Bs[GR][GT];
    // linear tile index -> (ti <= tj)
    const int nt = (K + GT - 1) / GT;
    int ti = 0, rem = blockIdx.x;
    while (rem >= nt - ti) {
        rem -= nt - ti;
        ++ti;
    }
    const int tj = ti + rem;
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    double acc[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) acc[a][b] = 0.0;
    for (int64_t r0 = 0; r0 < n_rows; r0 += GR) {
        for (int e = threadIdx.x; e < GR * GT; e += 256) {
            const int r = e / GT, c = e % GT;
            const int64_t row = r0 + r;
            const int ca = ti * GT + c, cb = tj * GT + c;
            As[r][c] = (row < n_rows && ca < K) ? (double)__ldg(X + row * ld + ca) : 0.0;
            Bs[r][c] = (row < n_rows && cb < K) ? (double)__ldg(X + row * ld + cb) : 0.0;
        }
        __syncthreads();
#pragma unroll
        for (int r = 0; r < GR; ++r) {
            double a[4], b[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                a[q] = As[r][ty * 4 + q];
                b[q] = Bs[r][tx * 4 + q];
            }
#pragma unroll
            for (int p = 0; p < 4; ++p)
#pragma unroll
                for (int q = 0; q < 4; ++q) acc[p][q] = fma(a[p], b[q], acc[p][q]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int p = 0; p < 4; ++p)
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int ci = ti * GT + ty * 4 + p, cj = tj * GT + tx * 4 + q;
            if (ci < K && cj < K) {
                C[(size_t)ci * K + cj] = acc[p][q];
                C[(size_t)cj * K + ci] = acc[p][q];
            }
        }
}

// Y[n, nc] = (X - mu) V  (mu optional), V [K, nc] fp64 row-major; fp64 accumulate, fp32 result.
constexpr int PR = 64;   // rows per CTA
constexpr int PK = 32;   // k chunk
constexpr int PC = 64;   // max components
__global__ void __launch_bounds__(256) project_kernel(const float* __restrict__ X, int64_t n_rows, int64_t ld, int K,
                                                      const double* __restrict__ V, int nc, const double* __restrict__ mu,
                                                      float* __restrict__ Y) {
    __shared__ float Xs[PR][PK + 1];
    __shared__ double Vs[PK][PC];
    const int r = threadIdx.x >> 2, cg = threadIdx.x & 3;
    const int64_t row0 = (int64_t)blockIdx.x * PR;
    double acc[PC / 4];
#pragma unroll
    for (int c = 0; c < PC / 4; ++c) acc[c] = 0.0;
    for (int k0 = 0; k0 < K; k0 += PK) {
        for (int e = threadIdx.x; e < PR * PK; e += 256) {
            const int rr = e / PK, kk = e % PK;
            const int64_t row = row0 + rr;
            float x = 0.f;
            if (row < n_rows && k0 + kk < K) {
                x = __ldg(X + row * ld + k0 + kk);
                if (mu) x = (float)((double)x - mu[k0 + kk]);
            }
            Xs[rr][kk] = x;
        }
        for (int e = threadIdx.x; e < PK * PC; e += 256) {
            const int kk = e / PC, c = e % PC;
            Vs[kk][c] = (k0 + kk < K && c < nc) ? V[(size_t)(k0 + kk) * nc + c] : 0.0;
        }
        __syncthreads();
#pragma unroll 4
        for (int kk = 0; kk < PK; ++kk) {
            const double x = (double)Xs[r][kk];
#pragma unroll
            for (int c = 0; c < PC / 4; ++c) acc[c] = fma(x, Vs[kk][cg + 4 * c], acc[c]);
        }
        __syncthreads();
    }
    const int64_t row = row0 + r;
    if (row < n_rows) {
#pragma unroll
        for (int c = 0; c < PC / 4; ++c)
            if (cg + 4 * c < nc) Y[row * nc + cg + 4 * c] = (float)acc[c];
    }
}

// ---------------------------------------------------------------------------------------------
// Exact k nearest neighbours (euclidean) on the PCA coordinates: P [n_all, d] float32 (d <= 64),
// queries are rows [q0, q0 + nq).  One warp per query: lanes stride over the candidates, every lane
// keeps its own sorted top-k in registers, then the 32 lists are merged through shared memory.
// scanpy's neighbors (pp/__init__.py:43 -> sc.pp.neighbors, method "umap") is exact below 4096 cells and
// approximate (pynndescent) above; this is exact at every size.
constexpr int KNN_MAXK = 32;
constexpr int KNN_D = 64;
template <int KK>
__global__ void __launch_bounds__(128) knn_kernel(const float* __restrict__ P, int64_t n_all, int d, int64_t q0, int64_t nq,
                                                  int32_t* __restrict__ knn_idx, float* __restrict__ knn_d2) {
    extern __shared__ __align__(16) unsigned char knn_smem[];
    float* q_s = reinterpret_cast<float*>(knn_smem);                 // [4 warps][KNN_D]
    float* md = q_s + 4 * KNN_D;                                       // [4 warps][32 * KK]
    int32_t* mi = reinterpret_cast<int32_t*>(md + 4 * 32 * KK);        // [4 warps][32 * KK]
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int64_t q = (int64_t)blockIdx.x * 4 + w;
    if (q >= nq) return;
    const int64_t qi = q0 + q;
    for (int c = lane; c < KNN_D; c += 32) q_s[w * KNN_D + c] = c < d ? P[qi * d + c] : 0.f;
    __syncwarp();
    float bd[KK];
    int32_t bi[KK];
#pragma unroll
    for (int t = 0; t < KK; ++t) {
        bd[t] = INFINITY;
        bi[t] = -1;
    }
    for (int64_t j = lane; j < n_all; j += 32) {
        const float* pj = P + j * d;
        float s = 0.f;
        for (int c = 0; c < d; ++c) {
            const float diff = q_s[w * KNN_D + c] - __ldg(pj + c);
            s = fmaf(diff, diff, s);
        }
        if (s < bd[KK - 1] || (s == bd[KK - 1] && (int32_t)j < bi[KK - 1])) {
            // insertion into the sorted list (ties: smaller index first)
            float cd = s;
            int32_t ci = (int32_t)j;
#pragma unroll
            for (int t = 0; t < KK; ++t) {
                const bool before = cd < bd[t] || (cd == bd[t] && ci < bi[t]);
                const float td = bd[t];
                const int32_t tix = bi[t];
                if (before) {
                    bd[t] = cd;
                    bi[t] = ci;
                    cd = td;
                    ci = tix;
                }
            }
        }
    }
    // merge: dump the 32 sorted lists, then KK rounds of "warp-wide minimum of the list heads"
    float* mdw = md + (size_t)w * 32 * KK;
    int32_t* miw = mi + (size_t)w * 32 * KK;
#pragma unroll
    for (int t = 0; t < KK; ++t) {
        mdw[lane * KK + t] = bd[t];
        miw[lane * KK + t] = bi[t];
    }
    __syncwarp();
    int head = 0;
    for (int t = 0; t < KK; ++t) {
        float hd = head < KK ? mdw[lane * KK + head] : INFINITY;
        int32_t hi = head < KK ? miw[lane * KK + head] : 0x7fffffff;
        float bdv = hd;
        int32_t biv = hi;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float od = __shfl_xor_sync(0xffffffffu, bdv, o);
            const int32_t oi = __shfl_xor_sync(0xffffffffu, biv, o);
            if (od < bdv || (od == bdv && oi < biv)) {
                bdv = od;
                biv = oi;
            }
        }
        if (hd == bdv && hi == biv) ++head;  // exactly one lane owns that (distance, index) pair
        if (lane == 0) {
            knn_idx[q * KK + t] = biv;
            knn_d2[q * KK + t] = bdv;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// umap-learn's fuzzy_simplicial_set pieces for one row each (smooth_knn_dist + membership strengths),
// restated from the published algorithm (umap-learn 0.5, umap_.py: smooth_knn_dist /
// compute_membership_strengths; local_connectivity = 1, bandwidth = 1, 64 bisection steps, tolerance 1e-5).
// dist [n, k] float32 sorted ascending with the query itself in column 0 (distance 0).
__global__ void __launch_bounds__(256) fuzzy_rows_kernel(const float* __restrict__ dist, const int32_t* __restrict__ idx, int64_t n,
                                                         int k, int64_t row0, float mean_all, float* __restrict__ vals,
                                                         float* __restrict__ sigma_out, float* __restrict__ rho_out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float* di = dist + i * k;
    const float target = log2f((float)k);
    float rho = 0.f;
    {
        // local_connectivity = 1: rho = smallest non-zero distance (max if there is none)
        bool found = false;
        float mx = 0.f;
        for (int j = 0; j < k; ++j) {
            const float dj = di[j];
            if (dj > 0.f) {
                if (!found) {
                    rho = dj;
                    found = true;
                }
                mx = fmaxf(mx, dj);
            }
        }
        (void)mx;
    }
    float lo = 0.f, hi = INFINITY, mid = 1.f;
    for (int it = 0; it < 64; ++it) {
        float psum = 0.f;
        for (int j = 1; j < k; ++j) {
            const float dd = di[j] - rho;
            psum += dd > 0.f ? expf(-(dd / mid)) : 1.f;
        }
        if (fabsf(psum - target) < 1e-5f) break;
        if (psum > target) {
            hi = mid;
            mid = (lo + hi) / 2.f;
        } else {
            lo = mid;
            if (hi == INFINITY)
                mid *= 2.f;
            else
                mid = (lo + hi) / 2.f;
        }
    }
    float sigma = mid;
    float mean_i = 0.f;
    for (int j = 0; j < k; ++j) mean_i += di[j];
    mean_i /= (float)k;
    const float floor_v = 1e-3f * (rho > 0.f ? mean_i : mean_all);
    if (sigma < floor_v) sigma = floor_v;
    sigma_out[i] = sigma;
    rho_out[i] = rho;
    for (int j = 0; j < k; ++j) {
        float v;
        if ((int64_t)idx[i * k + j] == row0 + i)
            v = 0.f;
        else if (di[j] - rho <= 0.f || sigma == 0.f)
            v = 1.f;
        else
            v = expf(-((di[j] - rho) / sigma));
        vals[i * k + j] = v;
    }
}

// ---------------------------------------------------------------------------------------------
// Community detection on the symmetric weighted CNV neighbourhood graph (CSR, both directions stored):
// one synchronous local-moving sweep of modularity optimisation with the RB-configuration null model
// (gamma * k_i * K_c / 2m), the quality leidenalg maximises for scanpy's tl.leiden.  One thread per node;
// degrees of this graph are ~15-60 so the per-node neighbour-community table lives in registers/local.
// A node only moves to a community with a smaller label on gain ties, and only half of the nodes
// (by parity of a hash of the sweep) are allowed to move per sweep, which keeps the synchronous
// update from oscillating.
constexpr int LV_MAXDEG = 96;
__global__ void __launch_bounds__(128) louvain_sweep_kernel(const int64_t* __restrict__ indptr, const int32_t* __restrict__ indices,
                                                            const float* __restrict__ w, const double* __restrict__ kdeg,
                                                            const int32_t* __restrict__ comm, const double* __restrict__ ctot,
                                                            int64_t n, double two_m, double gamma, int sweep,
                                                            int32_t* __restrict__ comm_new, int32_t* __restrict__ n_moved) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int32_t ci = comm[i];
    comm_new[i] = ci;
    // checkerboard: hash(i, sweep) parity decides who may move in this sweep
    uint32_t h = (uint32_t)i * 2654435761u + (uint32_t)sweep * 40503u;
    h ^= h >> 15;
    if ((h & 1u) != 0u) return;
    const int64_t e0 = indptr[i], e1 = indptr[i + 1];
    const int deg = (int)((e1 - e0) < (int64_t)LV_MAXDEG ? (e1 - e0) : (int64_t)LV_MAXDEG);
    int32_t cs[LV_MAXDEG];
    float ws[LV_MAXDEG];
    int nc = 0;
    float w_own = 0.f;
    for (int e = 0; e < deg; ++e) {
        const int32_t j = indices[e0 + e];
        if (j == (int32_t)i) continue;
        const int32_t cj = comm[j];
        const float wj = w[e0 + e];
        if (cj == ci) {
            w_own += wj;
            continue;
        }
        int t = 0;
        for (; t < nc; ++t)
            if (cs[t] == cj) break;
        if (t == nc) {
            cs[nc] = cj;
            ws[nc] = 0.f;
            ++nc;
        }
        ws[t] += wj;
    }
    const double ki = kdeg[i];
    // gain of staying (relative to being isolated) vs. joining c
    double best = (double)w_own - gamma * ki * (ctot[ci] - ki) / two_m;
    int32_t best_c = ci;
    for (int t = 0; t < nc; ++t) {
        const double g = (double)ws[t] - gamma * ki * ctot[cs[t]] / two_m;
        if (g > best + 1e-12 || (fabs(g - best) <= 1e-12 && cs[t] < best_c)) {
            best = g;
            best_c = cs[t];
        }
    }
    if (best_c != ci) {
        comm_new[i] = best_c;
        atomicAdd(n_moved, 1);
    }
}

__global__ void louvain_ctot_kernel(const int32_t* __restrict__ comm, const double* __restrict__ kdeg, int64_t n,
                                    double* __restrict__ ctot) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) atomicAdd(ctot + comm[i], kdeg[i]);
}

__global__ void weighted_degree_kernel(const int64_t* __restrict__ indptr, const float* __restrict__ w, int64_t n,
                                       double* __restrict__ kdeg) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double s = 0.0;
    for (int64_t e = indptr[i]; e < indptr[i + 1]; ++e) s += (double)w[e];
    kdeg[i] = s;
}

// ---------------------------------------------------------------------------------------------
// host launchers
int graph_csr_to_dense(const int64_t* indptr, const int32_t* indices, const void* data, bool f64, int64_t n, int K, float* dense,
                       int64_t ld, cudaStream_t st) {
    if (n == 0) return 0;
    const int grid = (int)(((n + 7) / 8) < (int64_t)(148 * 8) ? ((n + 7) / 8) : (int64_t)(148 * 8));
    if (f64)
        csr_to_dense_kernel<double><<<grid, 256, 0, st>>>(indptr, indices, (const double*)data, n, K, dense, ld);
    else
        csr_to_dense_kernel<float><<<grid, 256, 0, st>>>(indptr, indices, (const float*)data, n, K, dense, ld);
    ICNV_CUDA(cudaGetLastError());
    return 0;
}
int graph_gram(const float* X, int64_t n, int64_t ld, int K, double* C, cudaStream_t st) {
    const int nt = (K + GT - 1) / GT;
    gram_kernel<<<nt * (nt + 1) / 2, 256, 0, st>>>(X, n, ld, K, C);
    ICNV_CUDA(cudaGetLastError());
    return 0;
}
int graph_project(const float* X, int64_t n, int64_t ld, int K, const double* V, int nc, const double* mu, float* Y, cudaStream_t st) {
    if (n == 0) return 0;
    if (nc > PC) {
        set_error("icnv_project_f32: at most 64 components");
        return -3;
    }
    project_kernel<<<(unsigned)((n + PR - 1) / PR), 256, 0, st>>>(X, n, ld, K, V, nc, mu, Y);
    ICNV_CUDA(cudaGetLastError());
    return 0;
}
int graph_knn(const float* P, int64_t n_all, int d, int64_t q0, int64_t nq, int k, int32_t* idx, float* d2, cudaStream_t st) {
    if (nq == 0) return 0;
    if (d > KNN_D || k > KNN_MAXK || k < 1) {
        set_error("icnv_knn_f32: needs d <= 64 and 1 <= k <= 32");
        return -3;
    }
    const unsigned grid = (unsigned)((nq + 3) / 4);
#define ICNV_KNN(KK)                                                                                   \
    do {                                                                                               \
        const size_t smem = 4 * KNN_D * 4 + (size_t)4 * 32 * KK * 8;                                   \
        ICNV_CUDA(cudaFuncSetAttribute(knn_kernel<KK>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        knn_kernel<KK><<<grid, 128, smem, st>>>(P, n_all, d, q0, nq, idx, d2);                         \
    } while (0)
    if (k <= 16) {
        if (k != 16) {
            // the kernel is instantiated for 16 and 32; smaller k is served by the next size up and truncated by the caller
        }
        ICNV_KNN(16);
    } else {
        ICNV_KNN(32);
    }
#undef ICNV_KNN
    ICNV_CUDA(cudaGetLastError());
    return 0;
}
int graph_fuzzy_rows(const float* dist, const int32_t* idx, int64_t n, int k, int64_t row0, float mean_all, float* vals, float* sigma,
                     float* rho, cudaStream_t st) {
    if (n == 0) return 0;
    fuzzy_rows_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(dist, idx, n, k, row0, mean_all, vals, sigma, rho);
    ICNV_CUDA(cudaGetLastError());
    return 0;
}
int graph_weighted_degree(const int64_t* indptr, const float* w, int64_t n, double* kdeg, cudaStream_t st) {
    if (n == 0) return 0;
    weighted_degree_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(indptr, w, n, kdeg);
    ICNV_CUDA(cudaGetLastError());
    return 0;
}
int graph_louvain_sweep(const int64_t* indptr, const int32_t* indices, const float* w, const double* kdeg, const int32_t* comm,
                        double* ctot, int64_t n, double two_m, double gamma, int sweep, int32_t* comm_new, int32_t* n_moved,
                        cudaStream_t st) {
    if (n == 0) return 0;
    ICNV_CUDA(cudaMemsetAsync(ctot, 0, sizeof(double) * n, st));
    louvain_ctot_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(comm, kdeg, n, ctot);
    ICNV_CUDA(cudaMemsetAsync(n_moved, 0, sizeof(int32_t), st));
    louvain_sweep_kernel<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(indptr, indices, w, kdeg, comm, ctot, n, two_m, gamma, sweep,
                                                                      comm_new, n_moved);
    ICNV_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace icnv
