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
__global__ void __launch_bounds__(256, 4) colsum_dense_kernel(const float* __restrict__ X, int64_t n_rows, int64_t ldx,
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

// (the CSR column sums live in icnv_sparse.cu: deterministic, no atomics)
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

// row_nnz -> CSR indptr: indptr_scan_kernel (icnv_filter.cu)
int indptr_scan(const int32_t* row_nnz, int64_t n_rows, int64_t* indptr, cudaStream_t st);

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

// Step 4: exact row median + centring (/root/reference/src/infercnvpy/tl/_infercnv.py:442) and the row moments
// that feed the per-chunk std of :450.  Input: the smoothing kernel's fp64 rows in warp-tile order (value i of task t
// at (t/32)*32*LOUT + i*32 + t%32, unused slots = +inf) followed by one float2 (sum, sum of squares) per tile.
// One WARP per row, no block-level synchronisation: an exact selection is a long dependent chain, and with one row
// per warp ~3500 independent chains are in flight per GPU.
//
// Selection: the row is read once from HBM and turned into 16-bit order-preserving keys scaled to mean +- 1.02 sigma
// (from the tile moments; the median is always inside; +inf saturates to the top key), kept in the warp's shared
// memory.  The same sweep fills a 64-bin histogram (per-lane byte counters, so no atomics); the bin holding the two
// middle ranks has ~1 % of the row, i.e. usually <= 32 candidates, which are re-read (L2) and ranked exactly in fp64
// (np.median: mean of the two middle values for even K).  Crowded bins are refined (64 sub-bins per level); ties
// beyond that fall back to an exact bitwise selection.  The centring sweep reads the row a second time — it was
// fetched with an evict-last policy a few microseconds earlier, so this is L2 traffic.
__device__ __forceinline__ double warp_sum_dd(double x) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
    return x;
}
__device__ __forceinline__ unsigned long long ordered_bits64(double v) {
    const unsigned long long b = (unsigned long long)__double_as_longlong(v);
    return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
__device__ __forceinline__ double ldg_keep_f64(const double* p, uint64_t pol) {
    double v;
    asm volatile("ld.global.nc.L2::cache_hint.f64 %0, [%1], %2;" : "=d"(v) : "l"(p), "l"(pol));
    return v;
}
__device__ __forceinline__ uint16_t key16_sat(double v) {  // floor, saturating to [0, 65535]
    uint16_t k;
    asm("cvt.rmi.u16.f64 %0, %1;" : "=h"(k) : "d"(v));
    return k;
}

constexpr int CENTER_WARPS = 8;
constexpr int TILE_V = 32 * LOUT;
constexpr int HIST_BINS = 64;

struct CenterSmem {  // per-warp carve-up, shared by the kernel and its launcher
    size_t key_bytes, per_warp, geom_bytes, total;
    __host__ __device__ CenterSmem(int n_tiles, int warps) {
        // the output tile (<= 8 bytes per value) is staged over the keys, which are dead by then
        const size_t kb = (size_t)((n_tiles + 1) / 2) * TILE_V * 4;  // two tiles' keys per 32-bit word
        key_bytes = kb > (size_t)TILE_V * 8 ? kb : (size_t)TILE_V * 8;
        per_warp = (CAND_CAP * 8 + HIST_BINS * 32 + key_bytes + 15) / 16 * 16;
        geom_bytes = ((size_t)n_tiles * 32 * sizeof(int2) + 15) / 16 * 16;
        total = geom_bytes + per_warp * warps;
    }
};

template <typename TO>
__global__ void __launch_bounds__(32 * CENTER_WARPS, 3) center_rows_kernel(const double* __restrict__ tmp, int64_t n_rows, int64_t ld,
                                                                           const Task* __restrict__ tasks, int n_tasks, int K,
                                                                           TO* __restrict__ out, int64_t ldo,
                                                                           double* __restrict__ row_stats,
                                                                           unsigned long long* __restrict__ next_row) {
    extern __shared__ __align__(16) unsigned char cr_smem[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int n_tiles = (n_tasks + 31) >> 5;
    const int row_elems = n_tiles * TILE_V;
    const CenterSmem lay(n_tiles, CENTER_WARPS);
    int2* geom = reinterpret_cast<int2*>(cr_smem);  // CTA: (column, count) of every task
    unsigned char* wbase = cr_smem + lay.geom_bytes + wib * lay.per_warp;
    double* cand_w = reinterpret_cast<double*>(wbase);
    // histogram: word (b >> 2) * 32 + lane holds lane's byte counters of bins 4*(b>>2) .. +3: every access is a
    // conflict-free 32-bit access to the lane's own bank
    uint32_t* hist = reinterpret_cast<uint32_t*>(wbase + CAND_CAP * 8);
    TO* stage_w = reinterpret_cast<TO*>(hist + HIST_BINS * 8);
    // keys: word (pair * LOUT + i) * 32 + lane = key of (tile 2*pair, slot i) | key of (tile 2*pair+1, slot i) << 16
    uint32_t* keys_w = reinterpret_cast<uint32_t*>(hist + HIST_BINS * 8) + lane;
    uint32_t* hist_l = hist + lane;
    auto hist_inc = [&](uint32_t b) { hist_l[(b >> 2) * 32] += 1u << ((b & 3u) << 3); };
    auto hist_get = [&](int b) { return (int)((hist_l[(b >> 2) * 32] >> ((b & 3) << 3)) & 0xFFu); };
    for (int ti = threadIdx.x; ti < n_tiles * 32; ti += blockDim.x) {
        int2 g = make_int2(0, 0);
        if (ti < n_tasks) {
            const int4 tk = __ldg(reinterpret_cast<const int4*>(tasks) + ti);
            g = make_int2(tk.y, (tk.w & 0xFF) ? 1 : tk.z);
        }
        geom[ti] = g;
    }
    __syncthreads();
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    // rows are handed out dynamically (refinement levels make their cost uneven, and a static split leaves a tail);
    // a warp fetches its next row index while it works on the current one
    auto grab = [&]() {
        unsigned long long r = 0;
        if (lane == 0) r = atomicAdd(next_row, 1ull);
        return (int64_t)__shfl_sync(0xffffffffu, r, 0);
    };
    const int r1 = (K - 1) >> 1, r2 = K >> 1;
    const float invK = 1.f / (float)K;
    uint4* hist_clear = reinterpret_cast<uint4*>(hist) + lane;  // 4 x 16 bytes per lane, lanes contiguous
    // lane L totals bins 2L and 2L+1: bytes (2L & 3) and +1 of row L >> 1, all 32 lanes' words (skewed 16-byte reads)
    const uint4* hist_row = reinterpret_cast<const uint4*>(hist + (lane >> 1) * 32);
    const uint32_t sel0 = (lane & 1) ? 0x00010000u : 0x00000001u, sel1 = sel0 << 8;

#define ICNV_LOAD_TILE(X, T) \
    _Pragma("unroll") for (int i = 0; i < LOUT; ++i) X[i] = ldg_keep_f64(src + (T) * TILE_V + i * 32, pol)

    int64_t row_next = grab();
    while (row_next < n_rows) {
        const int64_t row = row_next;
        row_next = grab();
        const double* src = tmp + row * ld + lane;
        // ---- location / scale (fp32, steers the bracket only)
        float f1 = 0.f, f2 = 0.f;
        if (lane < n_tiles) {
            const float2 mo = __ldg(reinterpret_cast<const float2*>(tmp + row * ld + row_elems) + lane);
            f1 = mo.x;
            f2 = mo.y;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            f1 += __shfl_xor_sync(0xffffffffu, f1, o);
            f2 += __shfl_xor_sync(0xffffffffu, f2, o);
        }
        const float mean = f1 * invK;
        const float var = fmaxf(f2 * invK - mean * mean, 0.f);
        const float half = fmaxf(1.02f * sqrtf(var) + 1e-6f * fabsf(mean), 1e-20f);
        const double kbase = (double)(mean - half);
        const double kscale = (double)(32768.f / half);
        __syncwarp();
#pragma unroll
        for (int q = 0; q < 4; ++q) hist_clear[q * 32] = make_uint4(0u, 0u, 0u, 0u);
        __syncwarp();
        // ---- keys + level-0 histogram; two tiles of loads in flight per lane
        {
            double xa[LOUT], xb[LOUT];
            ICNV_LOAD_TILE(xa, 0);
            if (1 < n_tiles) ICNV_LOAD_TILE(xb, 1);
            for (int t = 0; t < n_tiles; t += 2) {
                uint32_t kk[LOUT];
#pragma unroll
                for (int i = 0; i < LOUT; ++i) {
                    kk[i] = key16_sat((xa[i] - kbase) * kscale);
                    hist_inc(kk[i] >> 10);
                }
                if (t + 2 < n_tiles) ICNV_LOAD_TILE(xa, t + 2);
                if (t + 1 < n_tiles) {
#pragma unroll
                    for (int i = 0; i < LOUT; ++i) {
                        const uint32_t kb = key16_sat((xb[i] - kbase) * kscale);
                        hist_inc(kb >> 10);
                        kk[i] |= kb << 16;
                    }
                    if (t + 3 < n_tiles) ICNV_LOAD_TILE(xb, t + 3);
                } else {
#pragma unroll
                    for (int i = 0; i < LOUT; ++i) kk[i] |= 0xFFFF0000u;
                }
#pragma unroll
                for (int i = 0; i < LOUT; ++i) keys_w[((t >> 1) * LOUT + i) * 32] = kk[i];
            }
        }

        uint32_t klo = 0, ksplit = 0;
        int shift = 10, below = 0, state = -1, b1 = -1, b2 = -1;
        while (true) {
            __syncwarp();
            // bin totals: lane L owns bins 2L, 2L+1
            uint32_t c0 = 0, c1 = 0;
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const uint4 v = hist_row[(q + (lane >> 1)) & 7];
                c0 = __dp4a(v.x, sel0, c0);
                c0 = __dp4a(v.y, sel0, c0);
                c0 = __dp4a(v.z, sel0, c0);
                c0 = __dp4a(v.w, sel0, c0);
                c1 = __dp4a(v.x, sel1, c1);
                c1 = __dp4a(v.y, sel1, c1);
                c1 = __dp4a(v.z, sel1, c1);
                c1 = __dp4a(v.w, sel1, c1);
            }
            int incl = (int)(c0 + c1);
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += t;
            }
            const int before = below + incl - (int)(c0 + c1);  // values below bin 2L
            const int cum0 = before + (int)c0, cum1 = cum0 + (int)c1;
            // first bin whose cumulative count exceeds the rank
            const unsigned m1 = __ballot_sync(0xffffffffu, cum1 > r1), m2 = __ballot_sync(0xffffffffu, cum1 > r2);
            if (m1 == 0u || m2 == 0u) {
                state = 3;  // cannot happen with a consistent bracket: exact bitwise selection over the whole row
                klo = 0;
                shift = 32;
                below = 0;
                break;
            }
            const int l1 = __ffs(m1) - 1, l2 = __ffs(m2) - 1;
            const int odd1 = __shfl_sync(0xffffffffu, cum0 > r1 ? 0 : 1, l1), odd2 = __shfl_sync(0xffffffffu, cum0 > r2 ? 0 : 1, l2);
            b1 = 2 * l1 + odd1;
            b2 = 2 * l2 + odd2;
            const int below1 = __shfl_sync(0xffffffffu, odd1 ? cum0 : before, l1);
            const int n1 = __shfl_sync(0xffffffffu, odd1 ? (int)c1 : (int)c0, l1);
            const int n2 = __shfl_sync(0xffffffffu, odd2 ? (int)c1 : (int)c0, l2);
            const int n_in = b1 == b2 ? n1 : n1 + n2;
            if (n_in <= CAND_CAP) {
                below = below1;
                state = 1;
                break;
            }
            if (b1 != b2) {
                state = 2;
                ksplit = klo + ((uint32_t)b2 << shift);
                break;
            }
            klo += (uint32_t)b1 << shift;
            below = below1;
            if (shift == 0) {
                state = 3;
                break;
            }
            shift = shift >= 6 ? shift - 6 : 0;
            // refine: histogram of the crowded bin's keys
            __syncwarp();
#pragma unroll
            for (int q = 0; q < 4; ++q) hist_clear[q * 32] = make_uint4(0u, 0u, 0u, 0u);
            __syncwarp();
            for (int t = 0; t < n_tiles; t += 2)
#pragma unroll
                for (int i = 0; i < LOUT; ++i) {
                    const uint32_t w = keys_w[((t >> 1) * LOUT + i) * 32];
                    const uint32_t da = ((w & 0xFFFFu) - klo) >> shift, db = ((w >> 16) - klo) >> shift;
                    if (da < (uint32_t)HIST_BINS) hist_inc(da);
                    if (db < (uint32_t)HIST_BINS && t + 1 < n_tiles) hist_inc(db);
                }
        }

        double m;
        if (state == 1) {
            // candidates = members of the bin(s) holding the two middle ranks
            const int mine_n = hist_get(b1) + (b2 != b1 ? hist_get(b2) : 0);
            int incl = mine_n;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += t;
            }
            const int n = __shfl_sync(0xffffffffu, incl, 31);
            int slot = incl - mine_n;
            if (mine_n)
                for (int t = 0; t < n_tiles; t += 2)
#pragma unroll
                    for (int i = 0; i < LOUT; ++i) {
                        const uint32_t w = keys_w[((t >> 1) * LOUT + i) * 32];
                        const uint32_t da = ((w & 0xFFFFu) - klo) >> shift, db = ((w >> 16) - klo) >> shift;
                        if (da == (uint32_t)b1 || da == (uint32_t)b2) cand_w[slot++] = ldg_keep_f64(src + t * TILE_V + i * 32, pol);
                        if ((db == (uint32_t)b1 || db == (uint32_t)b2) && t + 1 < n_tiles)
                            cand_w[slot++] = ldg_keep_f64(src + (t + 1) * TILE_V + i * 32, pol);
                    }
            __syncwarp();
            const double mine = lane < n ? cand_w[lane] : INFINITY;
            int rank = 0;
            for (int j = 0; j < n; ++j) {
                const double o = cand_w[j];
                rank += (o < mine) || (o == mine && j < lane);
            }
            const unsigned q1 = __ballot_sync(0xffffffffu, lane < n && rank == r1 - below);
            const unsigned q2 = __ballot_sync(0xffffffffu, lane < n && rank == r2 - below);
            const double lo = __shfl_sync(0xffffffffu, mine, (__ffs(q1) - 1) & 31);
            const double hi = __shfl_sync(0xffffffffu, mine, (__ffs(q2) - 1) & 31);
            m = (lo + hi) / 2.0;
        } else if (state == 2) {
            // the two middle ranks sit in different crowded bins: largest value below the split, smallest above
            double lo = -INFINITY, hi = INFINITY;
            for (int t = 0; t < n_tiles; ++t)
#pragma unroll
                for (int i = 0; i < LOUT; ++i) {
                    const double x = ldg_keep_f64(src + t * TILE_V + i * 32, pol);
                    const uint32_t w = keys_w[((t >> 1) * LOUT + i) * 32];
                    if (((t & 1) ? (w >> 16) : (w & 0xFFFFu)) < ksplit)
                        lo = fmax(lo, x);
                    else
                        hi = fmin(hi, x);
                }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                lo = fmax(lo, __shfl_xor_sync(0xffffffffu, lo, o));
                hi = fmin(hi, __shfl_xor_sync(0xffffffffu, hi, o));
            }
            m = (lo + hi) / 2.0;
        } else {
            // ties beyond the candidate capacity: exact bitwise selection on the order-preserving 64-bit pattern
            double res[2] = {0.0, 0.0};
            for (int which = 0; which < 2; ++which) {
                if (which == 1 && r2 == r1) {
                    res[1] = res[0];
                    break;
                }
                int rr = (which == 0 ? r1 : r2) - below;
                unsigned long long prefix = 0;
                for (int bit = 63; bit >= 0; --bit) {
                    int local = 0;
                    for (int t = 0; t < n_tiles; ++t) {
                        const int cnt = geom[t * 32 + lane].y;
#pragma unroll
                        for (int i = 0; i < LOUT; ++i)
                            if (i < cnt) {
                                const uint32_t w = keys_w[((t >> 1) * LOUT + i) * 32];
                                const uint32_t kx = (t & 1) ? (w >> 16) : (w & 0xFFFFu);
                                const bool in_set = shift >= 32 ? true : (((kx - klo) >> shift) == 0u);
                                const unsigned long long ob = ordered_bits64(ldg_keep_f64(src + t * TILE_V + i * 32, pol));
                                const bool same = bit == 63 ? true : ((ob >> (bit + 1)) == (prefix >> (bit + 1)));
                                local += in_set && same && !((ob >> bit) & 1ull);
                            }
                    }
                    const int zeros = __reduce_add_sync(0xffffffffu, local);
                    if (rr >= zeros) {
                        rr -= zeros;
                        prefix |= 1ull << bit;
                    }
                }
                const unsigned long long bb = (prefix >> 63) ? (prefix & 0x7FFFFFFFFFFFFFFFull) : ~prefix;
                res[which] = __longlong_as_double((long long)bb);
            }
            m = (res[0] + res[1]) / 2.0;
        }

        // ---- centre, row moments, natural-order write (second read of the row: L2)
        double s = 0.0, ss = 0.0;
        __syncwarp();  // keys are dead: their space becomes the output staging tile
        {
            double xa[LOUT], xb[LOUT];
            auto emit = [&](const double (&x)[LOUT], int t) {
                const int2 g = geom[t * 32 + lane];
                const int tb = __shfl_sync(0xffffffffu, g.x, 0);
                const int te = __reduce_max_sync(0xffffffffu, g.x + g.y);
#pragma unroll
                for (int i = 0; i < LOUT; ++i)
                    if (i < g.y) {
                        const double c = x[i] - m;
                        s += c;
                        ss = fma(c, c, ss);
                        stage_w[g.x - tb + i] = (TO)c;
                    }
                __syncwarp();
                TO* dst = out + row * ldo + tb;
                for (int k = lane; k < te - tb; k += 32) __stcs(dst + k, stage_w[k]);
                __syncwarp();
            };
            ICNV_LOAD_TILE(xa, 0);
            for (int t = 0; t < n_tiles; t += 2) {
                if (t + 1 < n_tiles) ICNV_LOAD_TILE(xb, t + 1);
                emit(xa, t);
                if (t + 1 < n_tiles) {
                    if (t + 2 < n_tiles) ICNV_LOAD_TILE(xa, t + 2);
                    emit(xb, t + 1);
                }
            }
        }
        s = warp_sum_dd(s);
        ss = warp_sum_dd(ss);
        if (lane == 0) {
            row_stats[2 * row] = s;
            row_stats[2 * row + 1] = ss;
        }
    }
#undef ICNV_LOAD_TILE
}

// Step 5: zero |v| < thr (strict, :451) in place on the natural-order matrix; per-row sum|v| and nnz.
template <typename T>
__global__ void __launch_bounds__(256) apply_threshold_kernel(T* __restrict__ out, int64_t n_rows, int64_t K, int64_t ldo,
                                                              int64_t chunk_rows, const double* __restrict__ thr,
                                                              double* __restrict__ row_abs, int32_t* __restrict__ row_nnz) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t r = warp; r < n_rows; r += n_warps) {
        const double t = thr ? thr[r / chunk_rows] : -1.0;
        T* row = out + r * ldo;
        double a = 0.0;
        int nz = 0;
        for (int64_t c = lane; c < K; c += 32) {
            const T vv = row[c];
            const double av = fabs((double)vv);
            if (av < t) {
                row[c] = (T)0;
            } else if (av != 0.0) {
                a += av;
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
// ITH scores: /root/reference/src/infercnvpy/tl/_scores.py:77-221 — np.corrcoef over the ROWS (cells) of one group.
// row_center_kernel: per row the mean and 1/sqrt(sum (x - mean)^2) (warp per row, two passes, fp64);
// row_corr_kernel: 64x64 tiles of cells (upper triangle, mirrored), fp64 FMAs over the features, then
// c_ij * inv_i * inv_j clipped to [-1, 1] like np.corrcoef; a row without variance gives 0 * inf = NaN like numpy's 0/0.
__global__ void __launch_bounds__(256) row_center_kernel(const double* __restrict__ X, int64_t n, int64_t ld, int K,
                                                         double* __restrict__ mean, double* __restrict__ inv) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t r = warp; r < n; r += n_warps) {
        const double* row = X + r * ld;
        double s = 0.0;
        for (int c = lane; c < K; c += 32) s += row[c];
        s = warp_sum_dd(s);
        const double m = s / (double)K;
        double ss = 0.0;
        for (int c = lane; c < K; c += 32) {
            const double d = row[c] - m;
            ss = fma(d, d, ss);
        }
        ss = warp_sum_dd(ss);
        if (lane == 0) {
            mean[r] = m;
            inv[r] = 1.0 / sqrt(ss);
        }
    }
}

constexpr int CT = 64;  // tile edge (cells)
constexpr int CR = 16;  // features per shared-memory chunk
__global__ void __launch_bounds__(256) row_corr_kernel(const double* __restrict__ X, int64_t n, int64_t ld, int K,
                                                       const double* __restrict__ mean, const double* __restrict__ inv,
                                                       double* __restrict__ C, int64_t ldc) {
    __shared__ double As[CR][CT + 1], Bs[CR][CT + 1];
    const int nt = (int)((n + CT - 1) / CT);
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
    for (int k0 = 0; k0 < K; k0 += CR) {
        for (int e = threadIdx.x; e < CR * CT; e += 256) {
            const int c = e / CR, r = e % CR;  // consecutive threads walk the features of one cell
            const int64_t ra = (int64_t)ti * CT + c, rb = (int64_t)tj * CT + c;
            const int k = k0 + r;
            As[r][c] = (ra < n && k < K) ? X[ra * ld + k] - mean[ra] : 0.0;
            Bs[r][c] = (rb < n && k < K) ? X[rb * ld + k] - mean[rb] : 0.0;
        }
        __syncthreads();
#pragma unroll
        for (int r = 0; r < CR; ++r) {
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
            const int64_t ci = (int64_t)ti * CT + ty * 4 + p, cj = (int64_t)tj * CT + tx * 4 + q;
            if (ci < n && cj < n && ci <= cj) {  // diagonal tiles: one writer per symmetric pair
                double v = acc[p][q] * (inv[ci] * inv[cj]);
                if (!isnan(v)) v = fmin(fmax(v, -1.0), 1.0);  // np.clip keeps NaN (fmin / fmax would drop it)
                C[ci * ldc + cj] = v;
                C[cj * ldc + ci] = v;
            }
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

// resident CTAs of the dense column-sum kernel on the current device (cached per device)
int aux_colsum_dense_slots(int* slots) {
    static int cached[16] = {0};
    int dev = 0;
    ICNV_CUDA(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 16 || cached[dev] == 0) {
        int n_sm = 0, occ = 0;
        ICNV_CUDA(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
        ICNV_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, colsum_dense_kernel<true>, 256, 0));
        const int v = n_sm * (occ > 0 ? occ : 1);
        if (dev < 0 || dev >= 16) {
            *slots = v;
            return 0;
        }
        cached[dev] = v;
    }
    *slots = cached[dev];
    return 0;
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
                   const int32_t* row_cat, int n_cat, double* sums, int64_t* counts, double* partial, int n_split, cudaStream_t st) {
    int rc = sparse_colsum_launch(indptr, indices, data, n_rows, G, row_cat, n_cat, partial, n_split, st);
    if (rc) return rc;
    const int64_t n = (int64_t)n_cat * G;
    reduce_partials_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(partial, n_split, n, sums);
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
    return indptr_scan(row_nnz, n_rows, indptr, st);
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

int aux_center_rows(const double* tmp, int64_t n_rows, int64_t ld, const Task* tasks, int n_tasks, int K, void* out, bool f64,
                    int64_t ldo, double* row_stats, cudaStream_t st) {
    if (n_rows == 0) return 0;
    const int n_tiles = (n_tasks + 31) / 32;
    if (n_tiles > 28) {
        set_error("icnv_center_rows: more than 28 tiles of tasks per row (byte counters would overflow)");
        return -3;
    }
    const size_t smem = CenterSmem(n_tiles, CENTER_WARPS).total;
    int dev = 0, n_sm = 148;
    ICNV_CUDA(cudaGetDevice(&dev));
    ICNV_CUDA(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
    int per_sm = (int)((size_t)227 * 1024 / (smem + 1024));
    if (per_sm > 3) per_sm = 3;  // 24 warps per SM
    if (per_sm < 1) per_sm = 1;
    const int64_t want = (n_rows + CENTER_WARPS - 1) / CENTER_WARPS;
    const int grid = (int)(want < (int64_t)n_sm * per_sm ? want : (int64_t)n_sm * per_sm);
    // per-device row counter (stream-ordered use, like the column-sum workspace)
    static unsigned long long* counter[16] = {nullptr};
    if (dev < 0 || dev >= 16) {
        set_error("icnv_center_rows: device index out of range");
        return -1;
    }
    if (!counter[dev]) ICNV_CUDA(cudaMalloc(&counter[dev], sizeof(unsigned long long)));
    ICNV_CUDA(cudaMemsetAsync(counter[dev], 0, sizeof(unsigned long long), st));
    if (f64) {
        ICNV_CUDA(cudaFuncSetAttribute(center_rows_kernel<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        center_rows_kernel<double><<<grid, 32 * CENTER_WARPS, smem, st>>>(tmp, n_rows, ld, tasks, n_tasks, K, (double*)out, ldo, row_stats,
                                                                        counter[dev]);
    } else {
        ICNV_CUDA(cudaFuncSetAttribute(center_rows_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        center_rows_kernel<float><<<grid, 32 * CENTER_WARPS, smem, st>>>(tmp, n_rows, ld, tasks, n_tasks, K, (float*)out, ldo, row_stats,
                                                                       counter[dev]);
    }
    ICNV_CUDA(cudaGetLastError());
    return 0;
}

int aux_apply_threshold(void* out, bool f64, int64_t n_rows, int64_t K, int64_t ldo, int64_t chunk_rows, const double* thr,
                        double* row_abs, int32_t* row_nnz, cudaStream_t st) {
    if (n_rows == 0) return 0;
    if (f64)
        apply_threshold_kernel<double><<<grid_for_rows(n_rows), 256, 0, st>>>((double*)out, n_rows, K, ldo, chunk_rows, thr, row_abs, row_nnz);
    else
        apply_threshold_kernel<float><<<grid_for_rows(n_rows), 256, 0, st>>>((float*)out, n_rows, K, ldo, chunk_rows, thr, row_abs, row_nnz);
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

int aux_row_corrcoef(const double* X, int64_t n, int64_t ld, int K, double* corr, int64_t ldc, double* work, cudaStream_t st) {
    if (n == 0) return 0;
    row_center_kernel<<<grid_for_rows(n), 256, 0, st>>>(X, n, ld, K, work, work + n);
    ICNV_CUDA(cudaGetLastError());
    const int64_t nt = (n + CT - 1) / CT;
    row_corr_kernel<<<(unsigned)(nt * (nt + 1) / 2), 256, 0, st>>>(X, n, ld, K, work, work + n, corr, ldc);
    ICNV_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace icnv
