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
// Community detection on the symmetric weighted CNV neighbourhood graph (CSR, both directions stored): the Leiden scheme
// (Traag, Waltman & van Eck 2019 — what leidenalg runs for scanpy's tl.leiden, /root/reference/src/infercnvpy/tl/
// __init__.py:24-30) with the RB-configuration quality  sum_c [ w_in(c) - gamma * K_c^2 / 2m ]:
//   1. local moving   nodes move to the neighbouring community with the best gain (a singleton joins another singleton
//                     only towards the smaller label)
//   2. refinement     inside every community of step 1 ("bound") the nodes start again as singletons; a singleton that is
//                     well connected to its bound ( w(v, C - v) >= gamma * k_v * (K_C - k_v) / 2m ) may join a
//                     neighbouring sub-community OF THE SAME BOUND if that has a positive gain -> sub-communities are
//                     connected by construction
//   3. aggregation    the sub-communities become the nodes of the next level and start it in their bound's community
// Both sweeps are synchronous and data-parallel (one thread per node; a CTA with a shared-memory hash table for nodes with
// more than LV_MAXDEG edges — hubs, and the nodes of aggregated levels).  A hash of (node, sweep) lets only half of the
// nodes move per sweep, gain ties go to the smaller label, and in the refinement a singleton only joins another singleton
// with a smaller label: no two nodes can swap, so the synchronous update does not oscillate.  Sums run in a fixed order
// (thread path) or in 2^-20 fixed point (hash path): the result is run-to-run deterministic.
constexpr int LV_MAXDEG = 96;
constexpr int LV_HASH = 8192;       // slots of the heavy path's table (distinct neighbouring communities <= 3/4 of it)
constexpr double LV_FIX = 1048576.0;  // 2^20

struct SweepArgs {
    const int64_t* indptr;
    const int32_t* indices;
    const float* w;
    const double* kdeg;
    const int32_t* comm;    // current assignment (refinement: the sub-community)
    const int32_t* bound;   // refinement only: community of step 1; nullptr = local moving
    const double* ctot;     // total degree per community of `comm`
    const double* btot;     // refinement: total degree per bound
    const int32_t* csize;   // refinement: members per sub-community
    int64_t n;
    double two_m, gamma;
    int sweep;
    int32_t* comm_new;
    int32_t* stats;         // [0] nodes moved, [1] heavy nodes queued, [2] hash overflow flag
    int32_t* heavy;         // queue of nodes with more than LV_MAXDEG edges
};

__device__ __forceinline__ bool sweep_active(int64_t i, int sweep) {
    uint32_t h = (uint32_t)i * 2654435761u + (uint32_t)sweep * 40503u;
    h ^= h >> 15;
    return (h & 1u) == 0u;
}

// decision for node i given its aggregated neighbourhood: (cs[t], ws[t]) for t < nc = weight to every other candidate
// community, w_own = weight to its own community (local moving) / w_bound = weight to its bound without itself (refinement)
__device__ __forceinline__ int32_t sweep_decide(const SweepArgs& a, int64_t i, int32_t ci, const int32_t* cs, const double* ws, int nc,
                                                double w_own, double w_bound) {
    const double ki = a.kdeg[i];
    double best;
    int32_t best_c = ci;
    if (a.bound == nullptr) {
        best = w_own - a.gamma * ki * (a.ctot[ci] - ki) / a.two_m;  // gain of staying, relative to being isolated
    } else {
        const double Kc = a.btot[a.bound[i]];
        if (w_bound < a.gamma * ki * (Kc - ki) / a.two_m) return ci;  // not well connected to its bound: stays alone
        best = 1e-12;                                                  // a singleton only moves for a positive gain
    }
    for (int t = 0; t < nc; ++t) {
        const int32_t c = cs[t];
        // two singletons: only the larger label moves (refinement: always; local moving: when the node itself is still a
        // singleton -- without it, pairs of singletons of a fresh level swap or chase each other under the synchronous
        // update and a 1M-node kNN graph is still moving thousands of nodes after 200 sweeps)
        if (a.csize[c] == 1 && c > ci && (a.bound != nullptr || a.csize[ci] == 1)) continue;
        const double g = ws[t] - a.gamma * ki * a.ctot[c] / a.two_m;
        if (g > best + 1e-12 || (best_c != ci && fabs(g - best) <= 1e-12 && c < best_c)) {
            best = g;
            best_c = c;
        }
    }
    return best_c;
}

__global__ void __launch_bounds__(128) community_sweep_kernel(const SweepArgs a) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.n) return;
    const int32_t ci = a.comm[i];
    a.comm_new[i] = ci;
    if (!sweep_active(i, a.sweep)) return;
    const bool refine = a.bound != nullptr;
    if (refine && a.csize[ci] != 1) return;  // only singletons move in the refinement
    const int64_t e0 = a.indptr[i], e1 = a.indptr[i + 1];
    if (e1 - e0 > LV_MAXDEG) {
        a.heavy[atomicAdd(a.stats + 1, 1)] = (int32_t)i;
        return;
    }
    const int deg = (int)(e1 - e0);
    const int32_t bi = refine ? a.bound[i] : 0;
    int32_t cs[LV_MAXDEG];
    double ws[LV_MAXDEG];
    int nc = 0;
    double w_own = 0.0, w_bound = 0.0;
    for (int e = 0; e < deg; ++e) {
        const int32_t j = a.indices[e0 + e];
        if (j == (int32_t)i) continue;
        const double wj = (double)a.w[e0 + e];
        if (refine) {
            if (a.bound[j] != bi) continue;
            w_bound += wj;
        }
        const int32_t cj = a.comm[j];
        if (cj == ci) {
            w_own += wj;
            continue;
        }
        int t = 0;
        for (; t < nc; ++t)
            if (cs[t] == cj) break;
        if (t == nc) {
            cs[nc] = cj;
            ws[nc] = 0.0;
            ++nc;
        }
        ws[t] += wj;
    }
    const int32_t best_c = sweep_decide(a, i, ci, cs, ws, nc, w_own, w_bound);
    if (best_c != ci) {
        a.comm_new[i] = best_c;
        atomicAdd(a.stats, 1);
    }
}

// heavy nodes: one CTA per queued node, neighbouring communities aggregated in a shared-memory hash table (fixed point)
__global__ void __launch_bounds__(256) community_sweep_heavy_kernel(const SweepArgs a) {
    extern __shared__ __align__(16) unsigned char lv_smem[];
    unsigned long long* vals = reinterpret_cast<unsigned long long*>(lv_smem);  // [LV_HASH]
    int32_t* keys = reinterpret_cast<int32_t*>(vals + LV_HASH);                  // [LV_HASH]
    __shared__ unsigned long long s_own, s_bound;
    __shared__ double s_best[256];
    __shared__ int32_t s_bestc[256];
    const int n_heavy = a.stats[1];
    const bool refine = a.bound != nullptr;
    for (int h = blockIdx.x; h < n_heavy; h += gridDim.x) {
        const int64_t i = a.heavy[h];
        const int32_t ci = a.comm[i];
        const int32_t bi = refine ? a.bound[i] : 0;
        for (int t = threadIdx.x; t < LV_HASH; t += 256) {
            keys[t] = -1;
            vals[t] = 0ull;
        }
        if (threadIdx.x == 0) s_own = s_bound = 0ull;
        __syncthreads();
        const int64_t e0 = a.indptr[i], e1 = a.indptr[i + 1];
        for (int64_t e = e0 + threadIdx.x; e < e1; e += 256) {
            const int32_t j = a.indices[e];
            if (j == (int32_t)i) continue;
            const unsigned long long wj = (unsigned long long)llrint((double)a.w[e] * LV_FIX);
            if (refine) {
                if (a.bound[j] != bi) continue;
                atomicAdd(&s_bound, wj);
            }
            const int32_t cj = a.comm[j];
            if (cj == ci) {
                atomicAdd(&s_own, wj);
                continue;
            }
            uint32_t slot = ((uint32_t)cj * 2654435761u) >> (32 - 13);  // LV_HASH = 2^13
            int probes = 0;
            while (true) {
                const int32_t prev = atomicCAS(&keys[slot], -1, cj);
                if (prev == -1 || prev == cj) {
                    atomicAdd(&vals[slot], wj);
                    break;
                }
                slot = (slot + 1) & (LV_HASH - 1);
                if (++probes >= LV_HASH) {  // table full: more distinct neighbouring communities than slots
                    atomicExch(a.stats + 2, 1);
                    break;
                }
            }
        }
        __syncthreads();
        // every thread scans its share of the table with the same rule as the thread path, then a fixed-order reduction
        const double ki = a.kdeg[i];
        double best = -1e300;
        int32_t best_c = ci;
        bool allowed = true;
        if (!refine) {
            best = (double)s_own / LV_FIX - a.gamma * ki * (a.ctot[ci] - ki) / a.two_m;
        } else {
            const double Kc = a.btot[bi];
            allowed = (double)s_bound / LV_FIX >= a.gamma * ki * (Kc - ki) / a.two_m;
            best = 1e-12;
        }
        const double base = best;
        for (int t = threadIdx.x; t < LV_HASH && allowed; t += 256) {
            const int32_t c = keys[t];
            if (c < 0) continue;
            if (a.csize[c] == 1 && c > ci && (refine || a.csize[ci] == 1)) continue;
            const double g = (double)vals[t] / LV_FIX - a.gamma * ki * a.ctot[c] / a.two_m;
            if (g > best + 1e-12 || (best_c != ci && fabs(g - best) <= 1e-12 && c < best_c)) {
                best = g;
                best_c = c;
            }
        }
        s_best[threadIdx.x] = best;
        s_bestc[threadIdx.x] = best_c;
        __syncthreads();
        if (threadIdx.x == 0) {
            double gb = base;
            int32_t cb = ci;
            for (int t = 0; t < 256; ++t) {
                const int32_t c = s_bestc[t];
                if (c == ci) continue;
                const double g = s_best[t];
                if (g > gb + 1e-12 || (cb != ci && fabs(g - gb) <= 1e-12 && c < cb)) {
                    gb = g;
                    cb = c;
                }
            }
            if (cb != ci) {
                a.comm_new[i] = cb;
                atomicAdd(a.stats, 1);
            }
        }
        __syncthreads();
    }
}

// totals per community: degree sum (ctot) and member count (csize, optional)
__global__ void community_totals_kernel(const int32_t* __restrict__ comm, const double* __restrict__ kdeg, int64_t n,
                                        double* __restrict__ ctot, int32_t* __restrict__ csize) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    // fixed point so that the totals do not depend on the order of the atomics
    atomicAdd(reinterpret_cast<unsigned long long*>(ctot) + comm[i], (unsigned long long)llrint(kdeg[i] * LV_FIX));
    if (csize) atomicAdd(csize + comm[i], 1);
}
__global__ void community_totals_finish_kernel(double* __restrict__ ctot, int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) ctot[i] = (double)reinterpret_cast<unsigned long long*>(ctot)[i] / LV_FIX;
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
static int community_totals(const int32_t* comm, const double* kdeg, int64_t n, double* ctot, int32_t* csize, cudaStream_t st) {
    ICNV_CUDA(cudaMemsetAsync(ctot, 0, sizeof(double) * n, st));
    if (csize) ICNV_CUDA(cudaMemsetAsync(csize, 0, sizeof(int32_t) * n, st));
    community_totals_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(comm, kdeg, n, ctot, csize);
    community_totals_finish_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(ctot, n);
    ICNV_CUDA(cudaGetLastError());
    return 0;
}

// One synchronous sweep.  bound == nullptr: local moving; else refinement inside `bound`.  work = [ctot n f64 | btot n f64 |
// csize n i32 | heavy n i32]; stats = int32[3] {moved, heavy nodes, overflow}.
int graph_community_sweep(const int64_t* indptr, const int32_t* indices, const float* w, const double* kdeg, const int32_t* comm,
                          const int32_t* bound, int64_t n, double two_m, double gamma, int sweep, void* work, int32_t* comm_new,
                          int32_t* stats, cudaStream_t st) {
    if (n == 0) return 0;
    double* ctot = reinterpret_cast<double*>(work);
    double* btot = ctot + n;
    int32_t* csize = reinterpret_cast<int32_t*>(btot + n);
    int32_t* heavy = csize + n;
    int rc = community_totals(comm, kdeg, n, ctot, csize, st);
    if (rc) return rc;
    if (bound) {
        rc = community_totals(bound, kdeg, n, btot, nullptr, st);
        if (rc) return rc;
    }
    ICNV_CUDA(cudaMemsetAsync(stats, 0, 3 * sizeof(int32_t), st));
    SweepArgs a;
    a.indptr = indptr;
    a.indices = indices;
    a.w = w;
    a.kdeg = kdeg;
    a.comm = comm;
    a.bound = bound;
    a.ctot = ctot;
    a.btot = btot;
    a.csize = csize;
    a.n = n;
    a.two_m = two_m;
    a.gamma = gamma;
    a.sweep = sweep;
    a.comm_new = comm_new;
    a.stats = stats;
    a.heavy = heavy;
    community_sweep_kernel<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(a);
    ICNV_CUDA(cudaGetLastError());
    constexpr size_t heavy_smem = (size_t)LV_HASH * 12;
    ICNV_CUDA(cudaFuncSetAttribute(community_sweep_heavy_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)heavy_smem));
    community_sweep_heavy_kernel<<<(unsigned)std::min<int64_t>(n, 148 * 2), 256, heavy_smem, st>>>(a);
    ICNV_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace icnv
