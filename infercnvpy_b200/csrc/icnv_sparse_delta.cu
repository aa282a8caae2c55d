// Sparse-aware smoothing of CSR input, work proportional to the STORED entries (the reference's CSR
// branch, /root/reference/src/infercnvpy/tl/_infercnv.py:115-116 + :423, densifies every chunk).
//
// Smoothing is linear after centring + clipping, and a zero of the matrix centres to the per-gene constant
// z_g = clip(0 - ref_g).  So   smooth(row) = smooth(z) + smooth(delta),   delta_g = clip(x_g - ref_g) - z_g  for the stored
// entries only.  smooth(z) (`base`, one row, the same for every cell) comes from smooth_csr_kernel (icnv_sparse.cu) on an
// empty row; this kernel adds the entries' deltas to the group partial sums  A_g = sum delta,  B_g = sum j * delta
// (C'_g = sum m_j * delta for a window with a peak group) and slides the same windows over them as icnv_smooth.cu.
//
// Entries arrive in column order, the groups are in genomic order: the adds are scattered, and several threads may hit one
// group at once.  To stay deterministic (bit-reproducible, any scheduling) the partial sums are INTEGERS: delta is taken in
// fixed point with 48 fractional bits (exact for every float32 pair with |value| >= 2^-24, else rounded at 2^-49 -- four
// orders of magnitude below the float64 rounding of a window sum), split into a 27-bit low limb and a signed high limb and
// added with native 32-bit shared-memory atomics (a group takes at most `step` adds, so no limb can overflow).
//
//   one persistent CTA per SM, 1024 threads; shared memory: per-column tables (group/element u16 + reference f32, 6 bytes
//   per gene) + two partial-sum buffers.
//   producer warps : entries (col, val) -> table -> delta -> 4 (6) atomic adds into buffer[it & 1]        -> arrive FULL
//   consumer warps : wait FULL -> limbs -> float64 A, B -> windows -> zero the buffer -> arrive FREE -> base + ... -> store
// HBM traffic per cell: 8 bytes per stored entry + the tile-order float64 row; nothing is proportional to the gene count.
#include "icnv_common.cuh"

namespace icnv {

namespace {

constexpr int DT = 1024;          // threads per CTA
constexpr int DW = DT / 32;
constexpr int FX_BITS = 48;       // fractional bits of the fixed-point deltas
constexpr int LIMB = 27;          // low limb width
constexpr int BAR_FULL = 1, BAR_FREE = 3, BAR_CONS = 5;

struct DeltaParams {
    const int64_t* indptr;
    const int32_t* indices;
    const float* data;
    int64_t n_rows;
    const uint16_t* gj;   // [G] (group << 4) | element, 0xFFFF: the column takes no part
    const float* ref;     // [G] reference profile (one category)
    int32_t G, Gpad2;     // Gpad2: bytes of the u16 table padded to 16
    int32_t NG, NGpad;
    const double* base;   // smooth(z) in tile order (one tmp row)
    double scale;         // 2^-FX_BITS / sum of the window weights
    double fx_inv;        // 2^-FX_BITS
    const double* flat_inv;
    const Task* tasks;
    int32_t n_tasks;
    float clipf;
    double* out;
    int64_t ldo;
};

__device__ __forceinline__ double limbs_to_double(int lo, int hi) {
    // exact: |hi| < 2^31, 0 <= lo < 2^31  ->  hi * 2^27 + lo  (|value| < 2^58, rounded once to float64)
    return fma((double)hi, 134217728.0, (double)lo);
}

template <int NWIN, int GS>
__global__ void __launch_bounds__(DT, 1) smooth_csr_delta_kernel(const DeltaParams p) {
    extern __shared__ __align__(16) unsigned char smem[];
    constexpr int NQ_C = NWIN / GS;
    constexpr bool M3_C = (NWIN / 2) % GS != 0;
    constexpr int QSTAR_C = M3_C ? (NWIN / 2) / GS : -1;
    const int ABS = p.NGpad + PAD_GROUPS;
    uint16_t* gj_s = reinterpret_cast<uint16_t*>(smem);
    float* ref_s = reinterpret_cast<float*>(smem + p.Gpad2);
    // limb arrays, structure of arrays so that a warp's scattered adds spread over all 32 banks (bank = group % 32):
    // L[buffer][limb][group], limbs = {A lo, A hi, B lo, B hi (, C lo, C hi for a window with a peak group)}
    constexpr int NL = M3_C ? 6 : 4;
    int* L = reinterpret_cast<int*>(smem + p.Gpad2 + (size_t)((p.G + 3) & ~3) * 4);

    const int lane = threadIdx.x & 31;
    const int warp = DW - 1 - (int)(threadIdx.x >> 5);  // consumers take the highest physical warp ids
    const int tid = warp * 32 + lane;
    const int n_cons = ((p.n_tasks + 31) >> 5) << 5;    // consumer threads: one per task (whole warps)
    const int n_prod = DT - n_cons;
    const bool consumer = tid < n_cons;

    // ---- one-time: tables into shared memory, buffers zeroed
    {
        const uint4* g4 = reinterpret_cast<const uint4*>(p.gj);
        uint4* d4 = reinterpret_cast<uint4*>(gj_s);
        for (int i = threadIdx.x; i < p.Gpad2 / 16; i += DT) d4[i] = __ldg(g4 + i);
        for (int i = threadIdx.x; i < p.G; i += DT) ref_s[i] = __ldg(p.ref + i);
        int4* L4 = reinterpret_cast<int4*>(L);
        for (int i = threadIdx.x; i < 2 * NL * ABS / 4; i += DT) L4[i] = make_int4(0, 0, 0, 0);
    }
    __syncthreads();

    const int64_t first = blockIdx.x;
    const int64_t n_it = first < p.n_rows ? (p.n_rows - first + gridDim.x - 1) / gridDim.x : 0;
    const float clipf = p.clipf;

    if (!consumer) {
        // =========================== producers ===========================
        const int pt = tid - n_cons;
        constexpr int U = 8;
        const int64_t nnz_all = __ldg(p.indptr + p.n_rows);
        for (int64_t it = 0; it < n_it; ++it) {
            const int64_t row = first + it * gridDim.x;
            const int b = (int)(it & 1);
            const int64_t e0 = __ldg(p.indptr + row);
            const int nnz = (int)(__ldg(p.indptr + row + 1) - e0);
            const int32_t* ip = p.indices + e0 + pt;
            const float* vp = p.data + e0 + pt;
            int* Lb = L + b * NL * ABS;
            int c[U];
            float v[U];
            int left = nnz - pt;  // entries from this thread's first one to the end of the row
            auto load = [&]() {
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    c[u] = -1;
                    v[u] = 0.f;
                    if (u * n_prod < left) {
                        c[u] = __ldg(ip + u * n_prod);
                        v[u] = ldg_stream_f32(vp + u * n_prod);
                    }
                }
            };
            load();  // in flight while the buffer is still being read by the consumers
            if (pt == 0 && it + 2 < n_it) {  // pull the entries of the row after the next one into L2
                const int64_t r2 = row + 2 * (int64_t)gridDim.x;
                const int64_t a0 = __ldg(p.indptr + r2) & ~(int64_t)3, a1 = min((__ldg(p.indptr + r2 + 1) + 3) & ~(int64_t)3, nnz_all & ~(int64_t)3);
                for (int64_t a = a0; a < a1; a += 4096) {
                    const uint32_t bytes = (uint32_t)(min((int64_t)4096, a1 - a) * 4);
                    bulk_prefetch_l2(p.indices + a, bytes);
                    bulk_prefetch_l2(p.data + a, bytes);
                }
            }
            if (it >= 2) named_bar_sync(BAR_FREE + b, DT);  // the consumers have read and zeroed this buffer
            while (true) {
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    if (u * n_prod >= left + lane) continue;  // warp-uniform: none of the warp's 32 slots holds an entry
                    const int cc = c[u] < 0 ? 0 : c[u];
                    const uint32_t gjv = gj_s[cc];
                    const float r = ref_s[cc];
                    const bool act = c[u] >= 0 && gjv != 0xFFFFu;
                    const float d = fminf(fmaxf(v[u] - r, -clipf), clipf);
                    const float z = fminf(fmaxf(-r, -clipf), clipf);
                    const long long dfx = __float2ll_rn(d * 281474976710656.0f) - __float2ll_rn(z * 281474976710656.0f);
                    const int g = (int)(gjv >> 4), j = (int)(gjv & 15u);
                    const long long bfx = dfx * j;
                    int* slot = Lb + g;
                    if (act) {
                        atomicAdd(slot, (int)((unsigned)dfx & ((1u << LIMB) - 1u)));
                        atomicAdd(slot + ABS, (int)(dfx >> LIMB));
                        if (j != 0) {
                            atomicAdd(slot + 2 * ABS, (int)((unsigned)bfx & ((1u << LIMB) - 1u)));
                            atomicAdd(slot + 3 * ABS, (int)(bfx >> LIMB));
                        }
                        if constexpr (M3_C) {
                            // non-linear part of the peak group's weights: m_j = w(q*, j) - w(q*, 0)
                            const int pk = (NWIN / 2) - GS * (M3_C ? QSTAR_C : 0);  // elements before the peak inside the group
                            const int m = j < pk ? j : 2 * pk - 1 - j;
                            if (m != 0) {
                                const long long cfx = dfx * m;
                                atomicAdd(slot + 4 * ABS, (int)((unsigned)cfx & ((1u << LIMB) - 1u)));
                                atomicAdd(slot + 5 * ABS, (int)(cfx >> LIMB));
                            }
                        }
                    }
                }
                left -= U * n_prod;
                if (left + pt <= 0) break;  // the row is exhausted for every producer at the same batch
                ip += U * n_prod;
                vp += U * n_prod;
                load();
            }
            named_bar_arrive(BAR_FULL + b, DT);
        }
        return;
    }

    // =========================== consumers ===========================
    int4 task = make_int4(0, 0, 0, 0);
    if (tid < p.n_tasks) task = __ldg(reinterpret_cast<const int4*>(p.tasks) + tid);
    const size_t tile_off = (size_t)(tid >> 5) * (32 * LOUT) + (tid & 31);
    const size_t mom_off = (size_t)((p.n_tasks + 31) >> 5) * (32 * LOUT);
    for (int64_t it = 0; it < n_it; ++it) {
        const int64_t row = first + it * gridDim.x;
        const int b = (int)(it & 1);
        int* Lb = L + b * NL * ABS;
        named_bar_sync(BAR_FULL + b, DT);  // every entry of the row has been added
        double v[LOUT];
        int nv = 0;
#pragma unroll
        for (int i = 0; i < LOUT; ++i) v[i] = 0.0;
        double scale = p.scale;
        if (tid < p.n_tasks) {
            const int4 t = task;
            if ((t.w & 0xFF) == 0) {
                nv = t.z;
                const int* P = Lb + t.x;
#pragma unroll
                for (int q = 0; q < NQ_C + LOUT - 1; ++q) {
                    const double A = limbs_to_double(P[q], P[q + ABS]), B = limbs_to_double(P[q + 2 * ABS], P[q + 3 * ABS]);
#pragma unroll
                    for (int i = 0; i < LOUT; ++i) {
                        const int w = q - i;
                        if (w >= 0 && w < NQ_C) {
                            const int al = pyr(NWIN, GS * w);
                            const int be = (GS > 1 && w != QSTAR_C) ? pyr(NWIN, GS * w + 1) - al : 0;
                            v[i] = fma((double)al, A, v[i]);
                            if (be == 1)
                                v[i] += B;
                            else if (be == -1)
                                v[i] -= B;
                        }
                    }
                }
                if constexpr (M3_C) {
#pragma unroll
                    for (int i = 0; i < LOUT; ++i) {
                        v[i] += limbs_to_double(P[QSTAR_C + i + 4 * ABS], P[QSTAR_C + i + 5 * ABS]);
                    }
                }
            } else {
                nv = 1;  // chromosome not longer than the window: one flat mean (_infercnv.py:227-236)
                double acc = 0.0;
                for (int g = 0; g < t.z; ++g) acc += limbs_to_double(Lb[t.x + g], Lb[t.x + g + ABS]);
                v[0] = acc;
                scale = p.fx_inv * p.flat_inv[t.w >> 8];
            }
        }
        named_bar_sync(BAR_CONS, n_cons);  // every consumer has read its partial sums
        {
            int4* Z = reinterpret_cast<int4*>(Lb);
            for (int i = tid; i < NL * ABS / 4; i += n_cons) Z[i] = make_int4(0, 0, 0, 0);
        }
        if (it + 2 < n_it) named_bar_arrive(BAR_FREE + b, DT);  // matched by the producers' wait two rows from now
        // ---- base + delta, tile moments, tile-order float64 row (layout of SmoothParams::out)
        double s1 = 0.0, s2 = 0.0;
#pragma unroll
        for (int i = 0; i < LOUT; ++i) {
            if (i < nv) {
                v[i] = fma(v[i], scale, __ldg(p.base + tile_off + (size_t)i * 32));
                s1 += v[i];
                s2 = fma(v[i], v[i], s2);
            } else {
                v[i] = INFINITY;
            }
        }
        float f1 = (float)s1, f2 = (float)s2;
#pragma unroll
        for (int sh = 16; sh > 0; sh >>= 1) {
            f1 += __shfl_xor_sync(0xffffffffu, f1, sh);
            f2 += __shfl_xor_sync(0xffffffffu, f2, sh);
        }
        double* orow = p.out + (size_t)row * p.ldo;
        if (lane == 0) reinterpret_cast<float2*>(orow + mom_off)[tid >> 5] = make_float2(f1, f2);
        double* o = orow + tile_off;
#pragma unroll
        for (int i = 0; i < LOUT; ++i) o[i * 32] = v[i];
    }
}

template <int NWIN, int GS>
int delta_launch_one(const DeltaParams& p, int n_sm, size_t smem, cudaStream_t st) {
    auto k = smooth_csr_delta_kernel<NWIN, GS>;
    ICNV_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int grid = (int)std::min<int64_t>(p.n_rows, (int64_t)n_sm);
    k<<<grid, DT, smem, st>>>(p);
    ICNV_CUDA(cudaGetLastError());
    return 0;
}

// per-column tables of the delta kernel: (group << 4) | element and the reference value
__global__ void delta_tables_kernel(const int4* __restrict__ col_tab, int G, int gs, uint16_t* __restrict__ gj, float* __restrict__ ref) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= G) return;
    const int4 t = col_tab[c];
    gj[c] = t.x < 0 ? (uint16_t)0xFFFFu : (uint16_t)(((t.x / gs) << 4) | (t.x % gs));
    ref[c] = __int_as_float(t.y);
}

}  // namespace

size_t sparse_delta_smem_bytes(int G, int NGpad, bool peak_group) {
    const size_t gpad2 = ((size_t)G * 2 + 15) / 16 * 16;
    return gpad2 + (size_t)((G + 3) & ~3) * 4 + 2 * (size_t)(NGpad + PAD_GROUPS) * (peak_group ? 24 : 16);
}

// one category, templated (window, step), group index fits 12 bits, one consumer thread per task
bool sparse_delta_supported(int nwin, int gs, int NGpad, int n_tasks) {
    return sparse_supported(nwin, gs) && NGpad + PAD_GROUPS < 4095 && gs <= 16 && ((n_tasks + 31) / 32) * 32 <= DT - 256;
}

int sparse_delta_tables_launch(const int4* col_tab, int G, int gs, uint16_t* gj, float* ref, cudaStream_t st) {
    delta_tables_kernel<<<(G + 255) / 256, 256, 0, st>>>(col_tab, G, gs, gj, ref);
    ICNV_CUDA(cudaGetLastError());
    return 0;
}

int sparse_delta_launch(int nwin, int gs, const int64_t* indptr, const int32_t* indices, const float* data, int64_t n_rows,
                        const uint16_t* gj, const float* ref, int G, int NG, int NGpad, const double* base, double inv_sumw,
                        const double* flat_inv, const Task* tasks, int n_tasks, float clipf, double* out, int64_t ldo, int n_sm,
                        cudaStream_t st) {
    DeltaParams p;
    p.indptr = indptr;
    p.indices = indices;
    p.data = data;
    p.n_rows = n_rows;
    p.gj = gj;
    p.ref = ref;
    p.G = G;
    p.Gpad2 = (int32_t)(((size_t)G * 2 + 15) / 16 * 16);
    p.NG = NG;
    p.NGpad = NGpad;
    p.base = base;
    p.fx_inv = 1.0 / 281474976710656.0;
    p.scale = p.fx_inv * inv_sumw;
    p.flat_inv = flat_inv;
    p.tasks = tasks;
    p.n_tasks = n_tasks;
    p.clipf = clipf;
    p.out = out;
    p.ldo = ldo;
    const size_t smem = sparse_delta_smem_bytes(G, NGpad, (nwin / 2) % gs != 0);
    if (nwin == 100 && gs == 10) return delta_launch_one<100, 10>(p, n_sm, smem, st);
    if (nwin == 250 && gs == 10) return delta_launch_one<250, 10>(p, n_sm, smem, st);
    set_error("sparse_delta_launch: no kernel instantiation for this (window, step)");
    return -3;
}

}  // namespace icnv
