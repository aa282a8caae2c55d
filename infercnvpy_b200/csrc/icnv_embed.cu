// Low-dimensional embeddings of the CNV neighbourhood graph: cnv.tl.umap / cnv.tl.tsne
// (/root/reference/src/infercnvpy/tl/__init__.py:78-144, thin wrappers around scanpy.tl.umap / scanpy.tl.tsne ->
// umap-learn's optimize_layout_euclidean and scikit-learn's TSNE).  None of that arithmetic is in the reference tree and
// its tests pin nothing here: parity unpinned; both optimisers are randomised upstream.  The kernels
// restate the published algorithms:
//   umap_epoch_kernel   one thread per directed edge: when the edge is due in this epoch it pulls its two ends together
//                       (gradient of the low-dimensional membership 1 / (1 + a d^2b), clipped to +-4) and pushes the head away
//                       from `negative_sample_rate` random vertices; updates are float atomics on the [n, 2] embedding.
//   tsne_*              exact t-SNE on <= 32768 points: conditional P from a per-row perplexity search over all pairwise
//                       distances, gradient = 4 (exaggeration * sum_j p_ij w_ij (y_i - y_j) - sum_j w_ij^2 / Z (y_i - y_j)),
//                       one warp per point sweeping all other points, momentum + gains update.
#include "icnv_common.cuh"

namespace icnv {

namespace {

__device__ __forceinline__ uint32_t mix32(uint32_t x) {  // murmur3 finaliser
    x ^= x >> 16;
    x *= 0x85ebca6bu;
    x ^= x >> 13;
    x *= 0xc2b2ae35u;
    x ^= x >> 16;
    return x;
}

__device__ __forceinline__ float clip4(float v) { return fminf(fmaxf(v, -4.f), 4.f); }

__global__ void __launch_bounds__(256) umap_epoch_kernel(const int32_t* __restrict__ head, const int32_t* __restrict__ tail, int64_t n_edges,
                                                         float* __restrict__ emb, int32_t n_vertices,
                                                         const float* __restrict__ epochs_per_sample, float* __restrict__ next_sample,
                                                         float* __restrict__ next_negative, float a, float b, float gamma, float alpha,
                                                         int epoch, int neg_rate, uint32_t seed) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n_edges) return;
    const float due = next_sample[e];
    if (due > (float)epoch) return;
    const int j = head[e], k = tail[e];
    float cx = emb[2 * j], cy = emb[2 * j + 1];
    {
        const float ox = emb[2 * k], oy = emb[2 * k + 1];
        const float dx = cx - ox, dy = cy - oy;
        const float d2 = dx * dx + dy * dy;
        float coeff = 0.f;
        if (d2 > 0.f) {
            const float pb = __powf(d2, b);
            coeff = -2.f * a * b * (pb / d2) / (a * pb + 1.f);
        }
        const float gx = clip4(coeff * dx) * alpha, gy = clip4(coeff * dy) * alpha;
        atomicAdd(emb + 2 * j, gx);
        atomicAdd(emb + 2 * j + 1, gy);
        atomicAdd(emb + 2 * k, -gx);
        atomicAdd(emb + 2 * k + 1, -gy);
        cx += gx;
        cy += gy;
    }
    const float eps = epochs_per_sample[e];
    next_sample[e] = due + eps;
    const float eps_neg = eps / (float)neg_rate;
    const float nn = next_negative[e];
    const int n_neg = (int)(((float)epoch - nn) / eps_neg);
    float ax = 0.f, ay = 0.f;
    for (int p = 0; p < n_neg; ++p) {
        const uint32_t r = mix32(seed ^ mix32((uint32_t)e * 0x9e3779b9u + (uint32_t)epoch * 0x632be5abu + (uint32_t)p));
        const int o = (int)(r % (uint32_t)n_vertices);
        if (o == j) continue;
        const float dx = cx - emb[2 * o], dy = cy - emb[2 * o + 1];
        const float d2 = dx * dx + dy * dy;
        float gx = 4.f, gy = 4.f;
        if (d2 > 0.f) {
            const float coeff = 2.f * gamma * b / ((0.001f + d2) * (a * __powf(d2, b) + 1.f));
            gx = clip4(coeff * dx);
            gy = clip4(coeff * dy);
        }
        ax += gx * alpha;
        ay += gy * alpha;
        cx += gx * alpha;
        cy += gy * alpha;
    }
    if (n_neg > 0) {
        atomicAdd(emb + 2 * j, ax);
        atomicAdd(emb + 2 * j + 1, ay);
        next_negative[e] = nn + (float)n_neg * eps_neg;
    }
}

// ------------------------------------------------------------------------------------------------ exact t-SNE
// squared euclidean distances of all pairs, [n, n] float32 (d <= 64 coordinates held in registers per point pair tile)
__global__ void __launch_bounds__(256) tsne_sqdist_kernel(const float* __restrict__ X, int n, int d, int64_t ld, float* __restrict__ D) {
    __shared__ float xi[16][65], xj[16][65];
    const int ti = blockIdx.y * 16, tj = blockIdx.x * 16;
    const int t = threadIdx.x;
    for (int q = t; q < 16 * d; q += 256) {
        const int r = q / d, c = q % d;
        xi[r][c] = ti + r < n ? X[(int64_t)(ti + r) * ld + c] : 0.f;
        xj[r][c] = tj + r < n ? X[(int64_t)(tj + r) * ld + c] : 0.f;
    }
    __syncthreads();
    const int a = t / 16, bb = t % 16;
    double s = 0.0;
    for (int c = 0; c < d; ++c) {
        const double df = (double)xi[a][c] - (double)xj[bb][c];
        s = fma(df, df, s);
    }
    if (ti + a < n && tj + bb < n) D[(int64_t)(ti + a) * n + tj + bb] = (float)s;
}

// conditional probabilities p_{j|i} with the row's own bandwidth: binary search of beta = 1 / (2 sigma^2) until the
// entropy of the row equals log(perplexity) (scikit-learn's _binary_search_perplexity: 100 steps, tolerance 1e-5).
// One warp per row; P overwrites D in place.
__global__ void __launch_bounds__(256) tsne_perplexity_kernel(float* __restrict__ D, int n, float log_perp) {
    const int lane = threadIdx.x & 31;
    const int row = (int)(((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    if (row >= n) return;
    float* d = D + (int64_t)row * n;
    float beta = 1.f, lo = -INFINITY, hi = INFINITY;
    float sum_p = 0.f;
    for (int it = 0; it < 100; ++it) {
        float sp = 0.f, sdp = 0.f;
        for (int j = lane; j < n; j += 32)
            if (j != row) {
                const float p = __expf(-d[j] * beta);
                sp += p;
                sdp += d[j] * p;
            }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            sp += __shfl_xor_sync(0xffffffffu, sp, o);
            sdp += __shfl_xor_sync(0xffffffffu, sdp, o);
        }
        sp = fmaxf(sp, 1e-30f);
        sum_p = sp;
        const float entropy = __logf(sp) + beta * sdp / sp;
        const float diff = entropy - log_perp;
        if (fabsf(diff) <= 1e-5f) break;
        if (diff > 0.f) {
            lo = beta;
            beta = isinf(hi) ? beta * 2.f : 0.5f * (beta + hi);
        } else {
            hi = beta;
            beta = isinf(lo) ? beta * 0.5f : 0.5f * (beta + lo);
        }
    }
    {   // normaliser of the final bandwidth (the search may have moved beta after its last evaluation)
        float sp = 0.f;
        for (int j = lane; j < n; j += 32)
            if (j != row) sp += __expf(-d[j] * beta);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) sp += __shfl_xor_sync(0xffffffffu, sp, o);
        sum_p = fmaxf(sp, 1e-30f);
    }
    for (int j = lane; j < n; j += 32) d[j] = j == row ? 0.f : __expf(-d[j] * beta) / sum_p;
}

// P <- max((P + P^T) / (2 n), 1e-12) on the upper triangle, mirrored
__global__ void __launch_bounds__(256) tsne_symmetrize_kernel(float* __restrict__ P, int n) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int i = (int)(idx / n), j = (int)(idx % n);
    if (i >= n || j <= i) return;
    const float v = fmaxf((P[(int64_t)i * n + j] + P[(int64_t)j * n + i]) / (2.f * (float)n), 1e-12f);
    P[(int64_t)i * n + j] = v;
    P[(int64_t)j * n + i] = v;
}

// per point: attractive and repulsive sums of the exact gradient, and its share of Z = sum_{k != l} w_kl
__global__ void __launch_bounds__(256) tsne_forces_kernel(const float* __restrict__ P, const float* __restrict__ Y, int n,
                                                          float* __restrict__ attr, float* __restrict__ rep, float* __restrict__ zpart) {
    const int lane = threadIdx.x & 31;
    const int i = (int)(((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    if (i >= n) return;
    const float yx = Y[2 * i], yy = Y[2 * i + 1];
    const float* p = P + (int64_t)i * n;
    float ax = 0.f, ay = 0.f, rx = 0.f, ry = 0.f, z = 0.f;
    for (int j = lane; j < n; j += 32) {
        if (j == i) continue;
        const float dx = yx - Y[2 * j], dy = yy - Y[2 * j + 1];
        const float w = 1.f / (1.f + dx * dx + dy * dy);
        const float pw = p[j] * w;
        ax += pw * dx;
        ay += pw * dy;
        const float w2 = w * w;
        rx += w2 * dx;
        ry += w2 * dy;
        z += w;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        ax += __shfl_xor_sync(0xffffffffu, ax, o);
        ay += __shfl_xor_sync(0xffffffffu, ay, o);
        rx += __shfl_xor_sync(0xffffffffu, rx, o);
        ry += __shfl_xor_sync(0xffffffffu, ry, o);
        z += __shfl_xor_sync(0xffffffffu, z, o);
    }
    if (lane == 0) {
        attr[2 * i] = ax;
        attr[2 * i + 1] = ay;
        rep[2 * i] = rx;
        rep[2 * i + 1] = ry;
        zpart[i] = z;
    }
}

// deterministic sum of zpart (one CTA, fixed tree)
__global__ void __launch_bounds__(1024) tsne_zsum_kernel(const float* __restrict__ zpart, int n, double* __restrict__ z) {
    __shared__ double sh[1024];
    double s = 0.0;
    for (int i = threadIdx.x; i < n; i += 1024) s += (double)zpart[i];
    sh[threadIdx.x] = s;
    __syncthreads();
    for (int o = 512; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) *z = sh[0];
}

// scikit-learn's _gradient_descent step: gains (+0.2 / x0.8, floor 0.01), momentum, learning rate
__global__ void __launch_bounds__(256) tsne_update_kernel(float* __restrict__ Y, float* __restrict__ vel, float* __restrict__ gains,
                                                          const float* __restrict__ attr, const float* __restrict__ rep,
                                                          const double* __restrict__ z, int n, float exaggeration, float momentum, float lr) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= 2 * n) return;
    const float g = 4.f * (exaggeration * attr[idx] - rep[idx] / (float)(*z));
    float gn = gains[idx];
    const float u = vel[idx];
    gn = (u * g < 0.f) ? gn + 0.2f : gn * 0.8f;
    gn = fmaxf(gn, 0.01f);
    gains[idx] = gn;
    const float un = momentum * u - lr * gn * g;
    vel[idx] = un;
    Y[idx] += un;
}

}  // namespace

int umap_epochs(const int32_t* head, const int32_t* tail, int64_t n_edges, float* emb, int32_t n_vertices, const float* eps,
                float* next_sample, float* next_negative, float a, float b, float gamma, float alpha0, int n_epochs, int epoch0, int n_run,
                int neg_rate, uint32_t seed, cudaStream_t st) {
    if (n_edges == 0 || n_run <= 0) return 0;
    const unsigned grid = (unsigned)((n_edges + 255) / 256);
    for (int ep = epoch0; ep < epoch0 + n_run && ep < n_epochs; ++ep) {
        const float alpha = alpha0 * (1.f - (float)ep / (float)n_epochs);
        umap_epoch_kernel<<<grid, 256, 0, st>>>(head, tail, n_edges, emb, n_vertices, eps, next_sample, next_negative, a, b, gamma, alpha,
                                               ep, neg_rate, seed);
    }
    ICNV_CUDA(cudaGetLastError());
    return 0;
}

int tsne_affinities(const float* X, int n, int d, int64_t ld, float perplexity, float* P, cudaStream_t st) {
    if (d > 64) {
        set_error("icnv_tsne_affinities: at most 64 coordinates");
        return -1;
    }
    dim3 grid((n + 15) / 16, (n + 15) / 16);
    tsne_sqdist_kernel<<<grid, 256, 0, st>>>(X, n, d, ld, P);
    tsne_perplexity_kernel<<<(unsigned)(((int64_t)n * 32 + 255) / 256), 256, 0, st>>>(P, n, logf(perplexity));
    tsne_symmetrize_kernel<<<(unsigned)(((int64_t)n * n + 255) / 256), 256, 0, st>>>(P, n);
    ICNV_CUDA(cudaGetLastError());
    return 0;
}

int tsne_iterations(const float* P, float* Y, float* vel, float* gains, float* work, int n, int n_iter, float exaggeration, float momentum,
                    float lr, cudaStream_t st) {
    // work: attr [2n] | rep [2n] | zpart [n] | z (double, 8-byte aligned slot at the end)
    float* attr = work;
    float* rep = work + 2 * (size_t)n;
    float* zpart = work + 4 * (size_t)n;
    double* z = reinterpret_cast<double*>(work + ((5 * (size_t)n + 1) & ~(size_t)1));
    for (int it = 0; it < n_iter; ++it) {
        tsne_forces_kernel<<<(unsigned)(((int64_t)n * 32 + 255) / 256), 256, 0, st>>>(P, Y, n, attr, rep, zpart);
        tsne_zsum_kernel<<<1, 1024, 0, st>>>(zpart, n, z);
        tsne_update_kernel<<<(2 * n + 255) / 256, 256, 0, st>>>(Y, vel, gains, attr, rep, z, n, exaggeration, momentum, lr);
    }
    ICNV_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace icnv
