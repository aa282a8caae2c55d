// CSR input without densifying the gene axis in matrix order (the reference's CSR branch is
// /root/reference/src/infercnvpy/tl/_infercnv.py:115-116 + :423, which calls .toarray() on every chunk).
//
// A zero entry of the matrix centres to the per-gene constant z_g = clip(centre(0, ref_g)), the same for every cell.  The
// kernel keeps a POSITION-SORTED row D[slot] (slot = group * gs + element: the order the windows are taken in) in shared
// memory that the TMA engine pre-fills with the constant row z (80 KB from L2, no SM instructions); the row's stored
// entries are then centred, clipped and scattered over it through a per-COLUMN table (slot, reference bounds).  Every
// group's partial sums A = sum x, B = sum j*x (C' = sum m_j*x for a window with a peak group) are then plain sequential
// reads of 10 consecutive floats — no gather tables, no bank conflicts — and the windows are the same fp64 sliding sums as
// in icnv_smooth.cu.  HBM traffic per cell: 8 * nnz + 8 (indptr) + the tile-order fp64 row, instead of 4 * G.
//
//   per CTA iteration (one cell row, persistent CTAs, 512 threads):
//     wait   mbarrier of D[cur]                       (z row landed; the refill was issued when the buffer died)
//     S      all warps: entries (col, val) -> table[col] -> d = clip(centre(val)) -> D[cur][slot]
//     bar A  (CTA)                                    group warps arrive here after P of the previous row
//     G      all warps: thread per group, 5 x LDS.64 -> A, B (, C') in fp64 -> AB[g]
//     bar B  (CTA)   D[cur] dead -> one thread re-arms its mbarrier and issues the TMA refill for iteration + 2
//     P      warps that own outputs: windows + tile-order stores; the others go on to S of the next row (D[next])
//   AB is single-buffered: G of the next row comes after bar A, which the group warps only reach after their P.
//
// Also here: the deterministic CSR column sums (one row at a time per CTA into a private accumulator -> no atomics, fixed
// order) that replace the fp64 atomicAdd version.
#include "icnv_common.cuh"

namespace icnv {

struct SparseParams {
    const int64_t* indptr;
    const int32_t* indices;
    const float* data;
    int64_t n_rows;
    const int4* col_tab;   // [G] {slot (-1: gene takes no part), lo bits, hi bits, 0}
    const float* zrow;     // [DP] constant row z in slot order (pads 0)
    int32_t DP;            // staged floats per row: NGpad * gs, multiple of 4
    int32_t NG, NGpad;
    double inv_sumw;
    const double* flat_inv;
    const Task* tasks;
    int32_t n_tasks;
    float clipf;
    double* out;           // tile-order fp64 rows + tile moments, like SmoothParams::out
    int64_t ldo;
};

struct __align__(16) SparseScratch {
    unsigned long long mbar[2];
};

__device__ __forceinline__ float centre_clip(float xv, float lo, float hi, float clipf, bool bounded) {
    float d;
    if (bounded)
        d = xv > hi ? xv - hi : (xv < lo ? xv - lo : 0.f);
    else
        d = xv - lo;
    return fminf(fmaxf(d, -clipf), clipf);
}

// z row in slot order: what a zero of the matrix becomes after centring and clipping (pads stay 0)
__global__ void zrow_kernel(const int32_t* __restrict__ slot_col, int n_slots, const int4* __restrict__ col_tab, float clipf,
                            int bounded, float* __restrict__ zrow) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_slots) return;
    const int c = slot_col[i];
    float z = 0.f;
    if (c >= 0) {
        const int4 t = col_tab[c];
        z = centre_clip(0.f, __int_as_float(t.y), __int_as_float(t.z), clipf, bounded != 0);
    }
    zrow[i] = z;
}

// per-column table from the reference rows: min / max over the categories (== the reference when there is one)
template <typename TR>
__global__ void col_table_kernel(const TR* __restrict__ ref, int n_cat, int G, const int32_t* __restrict__ col_slot,
                                 int4* __restrict__ col_tab) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= G) return;
    TR mn = ref[c], mx = ref[c];
    for (int k = 1; k < n_cat; ++k) {
        const TR v = ref[(size_t)k * G + c];
        mn = v < mn ? v : mn;
        mx = v > mx ? v : mx;
    }
    col_tab[c] = make_int4(col_slot[c], __float_as_int((float)mn), __float_as_int((float)mx), 0);
}

// DBL: two staged rows D[2] and one CTA per SM (windows with a peak group: the third partial sum leaves no room for a
// second CTA); otherwise one staged row per CTA and TWO CTAs per SM, whose phases interleave.
template <int NWIN, int GS, bool BOUNDED, bool DBL>
__global__ void __launch_bounds__(NT, DBL ? 1 : 2) smooth_csr_kernel(const SparseParams p) {
    extern __shared__ __align__(16) unsigned char smem[];
    constexpr int NQ_C = NWIN / GS;
    constexpr bool M3_C = (NWIN / 2) % GS != 0;
    constexpr int QSTAR_C = M3_C ? (NWIN / 2) / GS : -1;
    static_assert(NWIN % 2 == 0 && NWIN % GS == 0 && GS % 2 == 0, "templated even windows, even groups (LDS.64 walks)");
    const int ABS = p.NGpad + PAD_GROUPS;
    SparseScratch* sc = reinterpret_cast<SparseScratch*>(smem);
    constexpr int NBUF = DBL ? 2 : 1;
    float* D = reinterpret_cast<float*>(smem + 16);                       // [NBUF][DP]
    double2* AB = reinterpret_cast<double2*>(D + NBUF * (size_t)p.DP);     // [ABS]
    double* Cp = M3_C ? reinterpret_cast<double*>(AB + ABS) : nullptr;  // [ABS]

    // same warp numbering as smooth_kernel: the warps that own outputs take the highest physical ids
    const int lane = threadIdx.x & 31;
    const int warp = NW - 1 - (int)(threadIdx.x >> 5);
    const int tid = warp * 32 + lane;
    constexpr int ISSUER = NT - 32;
    const uint32_t row_bytes = (uint32_t)p.DP * 4u;
    const uint64_t pol = l2_policy_evict_first();  // only used as "some policy": the z row is re-read by every CTA
    (void)pol;

    for (int i = p.NG + tid; i < ABS; i += NT) {
        AB[i] = make_double2(0.0, 0.0);
        if (M3_C) Cp[i] = 0.0;
    }
    if (tid == 0) {
        mbar_init(&sc->mbar[0], 1);
        mbar_init(&sc->mbar[1], 1);
        mbar_fence_init();
    }
    __syncthreads();
    auto refill = [&](int buf) {
        mbar_expect_tx(&sc->mbar[buf], row_bytes);
        constexpr uint32_t CH = 16384;
        char* dst = reinterpret_cast<char*>(D + (size_t)buf * p.DP);
        const char* src = reinterpret_cast<const char*>(p.zrow);
        for (uint32_t off = 0; off < row_bytes; off += CH)
            bulk_g2s_plain(dst + off, src + off, min(CH, row_bytes - off), &sc->mbar[buf]);
    };
    // the entries of an upcoming row, pulled into L2 ahead of time (16-byte granules, clamped to the arrays)
    const int64_t nnz_all = __ldg(p.indptr + p.n_rows);
    auto prefetch_entries = [&](int64_t r) {
        const int64_t e0 = __ldg(p.indptr + r) & ~(int64_t)3, e1 = min((__ldg(p.indptr + r + 1) + 3) & ~(int64_t)3, nnz_all & ~(int64_t)3);
        if (e1 <= e0) return;
        constexpr int64_t CHE = 4096;  // entries per prefetch instruction (16 KB)
        for (int64_t e = e0; e < e1; e += CHE) {
            const uint32_t bytes = (uint32_t)(min(CHE, e1 - e) * 4);
            bulk_prefetch_l2(p.indices + e, bytes);
            bulk_prefetch_l2(p.data + e, bytes);
        }
    };
    const int64_t first = blockIdx.x;
    if (tid == ISSUER) {
        if (first < p.n_rows) refill(0);
        if (DBL && first + gridDim.x < p.n_rows) refill(1);
        if (first + gridDim.x < p.n_rows) prefetch_entries(first + gridDim.x);
    }

    const int n_group = ((p.n_tasks + 31) >> 5) << 5;
    const bool in_group = tid < n_group;
    int4 task = make_int4(0, 0, 0, 0);
    if (tid < p.n_tasks) task = __ldg(reinterpret_cast<const int4*>(p.tasks) + tid);
    const float clipf = p.clipf;

    int it = 0;
    for (int64_t row = first; row < p.n_rows; row += gridDim.x, ++it) {
        const int cur = DBL ? (it & 1) : 0;
        float* Dc = D + (size_t)cur * p.DP;
        // ======================= S: scatter the row's entries over the constant row =======================
        mbar_wait(&sc->mbar[cur], (uint32_t)((DBL ? (it >> 1) : it) & 1));
        {
            const int64_t e0 = __ldg(p.indptr + row);
            const int nnz = (int)(__ldg(p.indptr + row + 1) - e0);
            const int32_t* ip = p.indices + e0;
            const float* vp = p.data + e0;
            constexpr int U = 8;  // entries in flight per thread: the phase is two dependent L2 round trips per entry
            int e = tid;
            for (; e + (U - 1) * NT < nnz; e += U * NT) {
                int c[U];
                float v[U];
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    c[u] = __ldg(ip + e + u * NT);
                    v[u] = ldg_stream_f32(vp + e + u * NT);
                }
                if constexpr (BOUNDED) {
                    int4 t[U];
#pragma unroll
                    for (int u = 0; u < U; ++u) t[u] = __ldg(p.col_tab + c[u]);
#pragma unroll
                    for (int u = 0; u < U; ++u)
                        if (t[u].x >= 0) Dc[t[u].x] = centre_clip(v[u], __int_as_float(t[u].y), __int_as_float(t[u].z), clipf, true);
                } else {
                    int2 t[U];
#pragma unroll
                    for (int u = 0; u < U; ++u) t[u] = __ldg(reinterpret_cast<const int2*>(p.col_tab + c[u]));
#pragma unroll
                    for (int u = 0; u < U; ++u)
                        if (t[u].x >= 0) Dc[t[u].x] = centre_clip(v[u], __int_as_float(t[u].y), 0.f, clipf, false);
                }
            }
            {   // tail: up to U - 1 entries per thread, still issued together
                int c[U];
                float v[U];
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const bool ok = e + u * NT < nnz;
                    c[u] = ok ? __ldg(ip + e + u * NT) : -1;
                    v[u] = ok ? ldg_stream_f32(vp + e + u * NT) : 0.f;
                }
                int4 t[U];
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    t[u] = make_int4(-1, 0, 0, 0);
                    if (c[u] >= 0) {
                        if constexpr (BOUNDED) {
                            t[u] = __ldg(p.col_tab + c[u]);
                        } else {
                            const int2 t2 = __ldg(reinterpret_cast<const int2*>(p.col_tab + c[u]));
                            t[u].x = t2.x;
                            t[u].y = t2.y;
                        }
                    }
                }
#pragma unroll
                for (int u = 0; u < U; ++u)
                    if (t[u].x >= 0) Dc[t[u].x] = centre_clip(v[u], __int_as_float(t[u].y), __int_as_float(t[u].z), clipf, BOUNDED);
            }
        }
        __syncthreads();  // bar A: row complete in D[cur]; the previous row's windows have been read
        // ======================= G: group partial sums (sequential reads, compile-time weights) =======================
        for (int g = tid; g < p.NG; g += NT) {
            const float2* src = reinterpret_cast<const float2*>(Dc + (size_t)g * GS);
            double a = 0.0, b = 0.0, c = 0.0;
#pragma unroll
            for (int h = 0; h < GS / 2; ++h) {
                const float2 x = src[h];
                const double x0 = (double)x.x, x1 = (double)x.y;
                const int j0 = 2 * h, j1 = 2 * h + 1;
                if (h == 0) {
                    a = x0 + x1;
                    b = x1;  // j0 == 0
                } else {
                    a += x0;
                    a += x1;
                    b = fma((double)j0, x0, b);
                    b = fma((double)j1, x1, b);
                }
                if (M3_C) {
                    const int m0 = pyr(NWIN, GS * (QSTAR_C < 0 ? 0 : QSTAR_C) + j0) - pyr(NWIN, GS * (QSTAR_C < 0 ? 0 : QSTAR_C));
                    const int m1 = pyr(NWIN, GS * (QSTAR_C < 0 ? 0 : QSTAR_C) + j1) - pyr(NWIN, GS * (QSTAR_C < 0 ? 0 : QSTAR_C));
                    if (m0 != 0) c = fma((double)m0, x0, c);
                    if (m1 != 0) c = fma((double)m1, x1, c);
                }
            }
            AB[g] = make_double2(a, b);
            if (M3_C) Cp[g] = c;
        }
        __syncthreads();  // bar B: partials visible, D[cur] dead
        if (tid == ISSUER) {
            const int64_t nxt = row + (DBL ? 2 : 1) * (int64_t)gridDim.x;  // the row this buffer serves next
            if (nxt < p.n_rows) refill(cur);
            if (row + 2 * (int64_t)gridDim.x < p.n_rows) prefetch_entries(row + 2 * (int64_t)gridDim.x);
        }
        if (!in_group) continue;

        // ======================= P: windows (same arithmetic as smooth_kernel, tier 0) =======================
        double v[LOUT];
        int nv = 0;
#pragma unroll
        for (int i = 0; i < LOUT; ++i) v[i] = INFINITY;
        if (tid < p.n_tasks) {
            const int4 t = task;
            if ((t.w & 0xFF) == 0) {
                nv = t.z;
                double acc[LOUT];
#pragma unroll
                for (int i = 0; i < LOUT; ++i) acc[i] = 0.0;
                const double2* P = AB + t.x;
#pragma unroll
                for (int q = 0; q < NQ_C + LOUT - 1; ++q) {
                    const double2 ab = P[q];
#pragma unroll
                    for (int i = 0; i < LOUT; ++i) {
                        const int w = q - i;
                        if (w >= 0 && w < NQ_C) {
                            const int al = pyr(NWIN, GS * w);
                            const int be = (GS > 1 && w != QSTAR_C) ? pyr(NWIN, GS * w + 1) - al : 0;
                            acc[i] = fma((double)al, ab.x, acc[i]);
                            if (be == 1)
                                acc[i] += ab.y;
                            else if (be == -1)
                                acc[i] -= ab.y;
                        }
                    }
                }
                if (M3_C) {
#pragma unroll
                    for (int i = 0; i < LOUT; ++i) acc[i] += Cp[t.x + QSTAR_C + i];
                }
#pragma unroll
                for (int i = 0; i < LOUT; ++i)
                    if (i < t.z) v[i] = acc[i] * p.inv_sumw;
            } else {
                nv = 1;  // chromosome not longer than the window: one flat mean (_infercnv.py:227-236)
                double acc = 0.0;
                for (int g = 0; g < t.z; ++g) acc += AB[t.x + g].x;
                v[0] = acc * p.flat_inv[t.w >> 8];
            }
        }
        // tile-order fp64 row + tile moments (layout of SmoothParams::out)
        {
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
            double* orow = p.out + (size_t)row * p.ldo;
            if (lane == 0) reinterpret_cast<float2*>(orow + (size_t)((p.n_tasks + 31) >> 5) * (32 * LOUT))[tid >> 5] = make_float2(f1, f2);
            double* o = orow + (size_t)(tid >> 5) * (32 * LOUT) + (tid & 31);
#pragma unroll
            for (int i = 0; i < LOUT; ++i) o[i * 32] = v[i];
        }
    }
}

size_t sparse_smem_bytes(int DP, int NGpad, bool peak_group) {
    size_t s = 16 + (size_t)(peak_group ? 2 : 1) * DP * 4 + (size_t)(NGpad + PAD_GROUPS) * 16;
    if (peak_group) s += (size_t)(NGpad + PAD_GROUPS) * 8;
    return (s + 15) / 16 * 16;
}

int sparse_zrow_launch(const int32_t* slot_col, int n_slots, const int4* col_tab, float clipf, bool bounded, float* zrow, cudaStream_t st) {
    if (n_slots == 0) return 0;
    zrow_kernel<<<(n_slots + 255) / 256, 256, 0, st>>>(slot_col, n_slots, col_tab, clipf, bounded ? 1 : 0, zrow);
    ICNV_CUDA(cudaGetLastError());
    return 0;
}

int sparse_col_table_launch(const void* ref, bool ref_f64, int n_cat, int G, const int32_t* col_slot, int4* col_tab, cudaStream_t st) {
    if (ref_f64)
        col_table_kernel<double><<<(G + 255) / 256, 256, 0, st>>>((const double*)ref, n_cat, G, col_slot, col_tab);
    else
        col_table_kernel<float><<<(G + 255) / 256, 256, 0, st>>>((const float*)ref, n_cat, G, col_slot, col_tab);
    ICNV_CUDA(cudaGetLastError());
    return 0;
}

template <int NWIN, int GS, bool BOUNDED>
static int sparse_launch_one(const SparseParams& p, int n_sm, size_t smem, cudaStream_t st) {
    constexpr bool DBL = (NWIN / 2) % GS != 0;
    auto k = smooth_csr_kernel<NWIN, GS, BOUNDED, DBL>;
    ICNV_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int occ = 1;
    ICNV_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k, NT, smem));
    if (occ < 1) {
        set_error("sparse smoothing kernel does not fit on an SM");
        return -3;
    }
    const int grid = (int)std::min<int64_t>(p.n_rows, (int64_t)n_sm * occ);
    k<<<grid, NT, smem, st>>>(p);
    ICNV_CUDA(cudaGetLastError());
    return 0;
}

bool sparse_supported(int nwin, int gs) { return gs == 10 && (nwin == 100 || nwin == 250); }

int sparse_smooth_launch(int nwin, int gs, bool bounded, const int64_t* indptr, const int32_t* indices, const float* data, int64_t n_rows,
                         const int4* col_tab, const float* zrow, int DP, int NG, int NGpad, double inv_sumw, const double* flat_inv,
                         const Task* tasks, int n_tasks, float clipf, double* out, int64_t ldo, int n_sm, cudaStream_t st) {
    SparseParams p;
    p.indptr = indptr;
    p.indices = indices;
    p.data = data;
    p.n_rows = n_rows;
    p.col_tab = col_tab;
    p.zrow = zrow;
    p.DP = DP;
    p.NG = NG;
    p.NGpad = NGpad;
    p.inv_sumw = inv_sumw;
    p.flat_inv = flat_inv;
    p.tasks = tasks;
    p.n_tasks = n_tasks;
    p.clipf = clipf;
    p.out = out;
    p.ldo = ldo;
    const size_t smem = sparse_smem_bytes(DP, NGpad, (nwin / 2) % gs != 0);
    if (nwin == 100 && gs == 10)
        return bounded ? sparse_launch_one<100, 10, true>(p, n_sm, smem, st) : sparse_launch_one<100, 10, false>(p, n_sm, smem, st);
    if (nwin == 250 && gs == 10)
        return bounded ? sparse_launch_one<250, 10, true>(p, n_sm, smem, st) : sparse_launch_one<250, 10, false>(p, n_sm, smem, st);
    set_error("sparse_smooth_launch: no kernel instantiation for this (window, step)");
    return -3;
}

// ------------------------------------------------------------------------------------------------
// Deterministic CSR column sums.  A CTA owns a contiguous block of rows and a private fp64 accumulator per column
// (shared memory when [G] doubles fit and there is one category, else its own slice of a global workspace that stays
// in L2).  It walks its rows ONE AT A TIME: inside a canonical CSR row every column occurs once, so the 512 threads
// update distinct accumulators with plain read-modify-writes, and a CTA barrier separates consecutive rows.  Every
// column is therefore summed in row order — no atomics, bit-reproducible — and the per-CTA partials are added in CTA
// order by reduce_partials_kernel.  The next row's entries are loaded into registers before the barrier.
template <bool SMEM>
__global__ void __launch_bounds__(512, 1) colsum_csr_det_kernel(const int64_t* __restrict__ indptr, const int32_t* __restrict__ indices,
                                                                const float* __restrict__ data, int64_t n_rows, int G,
                                                                const int32_t* __restrict__ row_cat, int n_cat,
                                                                double* __restrict__ partial) {
    extern __shared__ __align__(16) unsigned char smem[];
    double* acc_s = reinterpret_cast<double*>(smem);
    const int n_split = gridDim.x;
    const int64_t rows_per = (n_rows + n_split - 1) / n_split;
    const int64_t r0 = (int64_t)blockIdx.x * rows_per, r1 = min(n_rows, r0 + rows_per);
    double* mine = partial + (size_t)blockIdx.x * n_cat * G;  // [n_cat][G], zeroed by the host (global variant) / written at the end
    if (SMEM) {
        for (int i = threadIdx.x; i < G; i += 512) acc_s[i] = 0.0;
        __syncthreads();
    }
    constexpr int ENT = 12;  // entries per thread held for the next row (rows up to 6144 entries are fully prefetched)
    int pc[ENT];
    float pv[ENT];
    auto row_ok = [&](int64_t r) { return row_cat ? (row_cat[r] >= 0 && row_cat[r] < n_cat) : true; };
    auto prefetch = [&](int64_t r) {
        const int64_t e0 = indptr[r], e1 = indptr[r + 1];
#pragma unroll
        for (int k = 0; k < ENT; ++k) {
            const int64_t e = e0 + threadIdx.x + (int64_t)k * 512;
            pc[k] = -1;
            if (e < e1) {
                pc[k] = __ldg(indices + e);
                pv[k] = ldg_stream_f32(data + e);
            }
        }
    };
    int64_t r = r0;
    while (r < r1 && !row_ok(r)) ++r;
    if (r < r1) prefetch(r);
    while (r < r1) {
        const int cat = row_cat ? row_cat[r] : 0;
        double* acc = SMEM ? acc_s : mine + (size_t)cat * G;
        int cc[ENT];
        float cv[ENT];
#pragma unroll
        for (int k = 0; k < ENT; ++k) {
            cc[k] = pc[k];
            cv[k] = pv[k];
        }
        const int64_t e0 = indptr[r], e1 = indptr[r + 1];
        int64_t rn = r + 1;
        while (rn < r1 && !row_ok(rn)) ++rn;
        if (rn < r1) prefetch(rn);
#pragma unroll
        for (int k = 0; k < ENT; ++k)
            if (cc[k] >= 0) {
                if (SMEM)
                    acc[cc[k]] += (double)cv[k];
                else
                    __stcg(acc + cc[k], __ldcg(acc + cc[k]) + (double)cv[k]);
            }
        for (int64_t e = e0 + threadIdx.x + (int64_t)ENT * 512; e < e1; e += 512) {  // very long rows
            const int c = __ldg(indices + e);
            const double v = (double)__ldg(data + e);
            if (SMEM)
                acc[c] += v;
            else
                __stcg(acc + c, __ldcg(acc + c) + v);
        }
        __syncthreads();
        r = rn;
    }
    if (SMEM) {
        for (int i = threadIdx.x; i < G; i += 512) mine[i] = acc_s[i];
    }
}

int sparse_colsum_splits(int64_t n_rows, int n_sm) { return (int)std::max<int64_t>(1, std::min<int64_t>(n_sm, (n_rows + 63) / 64)); }

int sparse_colsum_launch(const int64_t* indptr, const int32_t* indices, const float* data, int64_t n_rows, int G, const int32_t* row_cat,
                         int n_cat, double* partial, int n_split, cudaStream_t st) {
    const size_t smem = (size_t)G * 8;
    const bool in_smem = n_cat == 1 && smem <= (size_t)220 * 1024;
    if (in_smem) {
        ICNV_CUDA(cudaFuncSetAttribute(colsum_csr_det_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        colsum_csr_det_kernel<true><<<n_split, 512, smem, st>>>(indptr, indices, data, n_rows, G, row_cat, n_cat, partial);
    } else {
        ICNV_CUDA(cudaMemsetAsync(partial, 0, sizeof(double) * (size_t)n_split * n_cat * G, st));
        colsum_csr_det_kernel<false><<<n_split, 512, 0, st>>>(indptr, indices, data, n_rows, G, row_cat, n_cat, partial);
    }
    ICNV_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace icnv
