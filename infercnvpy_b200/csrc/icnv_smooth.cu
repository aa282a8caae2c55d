// Smoothing kernel: steps 1-4 of /root/reference/src/infercnvpy/tl/_infercnv.py:411-442
// (centre, clip, per-chromosome pyramid running mean decimated by `step`, row-median centring)
// for one cell row per CTA iteration, persistent CTAs.
//
// Data flow per row (tiers 0/1, "grouped"):
//   HBM --cp.async.bulk (TMA, mbarrier)--> smem raw row [G] fp32
//   phase 2: every thread owns 4 position-ordered groups of `gs` (= step) genes; it gathers its genes
//            from the raw row through a pre-scaled byte-offset table (warp-coalesced 16-byte table
//            loads, immediate offsets), centres + clips in fp32 exactly like numpy, and accumulates
//            per-group partial sums in fp64:
//              A_g = sum_j x_j,  B_g = sum_j j*x_j,  C_g = sum_j cw_j*x_j (peak group only)
//            Within a step-aligned group the pyramid weights are linear in j, so every window is
//              out_k = sum_q alpha_q*A_{k+q} + beta_q*B_{k+q}  (+ C_{k+q*})
//            i.e. 2*window/step FMAs instead of `window` — and nothing is computed for the 90 % of
//            windows the reference computes and then drops (:215-218).
//   phase 3: one thread per LOUT=9 consecutive outputs slides over the partials (fp64 FMAs, weights
//            are compile-time immediates in tier 0), and accumulates the row's sum / sum of squares.
//   median : exact selection on order-preserving 32-bit keys held in registers: 8-bin counting
//            passes with packed 4-bit counters + warp REDUX (one barrier per pass, every warp
//            derives the next bracket redundantly), then an exact fp64 ranking of <= 32
//            candidates (np.median semantics: mean of the two middle values for even K).
//   write  : out[row, :] = v - median (fp32 or fp64); row statistics follow from the sums.
// Tier 2 ("direct") evaluates the reference formula literally from a position-sorted centred row in
// smem; it covers every (window, step) and float64 centring and is the slow general fallback.
#include "icnv_common.cuh"

namespace icnv {

struct __align__(16) Scratch {
    uint4 wcnt[2][NW];
    double wred[NW][2];
    double cand[CAND_CAP];
    double med[2];
    unsigned long long mbar;
    int cand_n;
    int bcnt[NW];
    int btotal;
    int next_wb[2];  // phase-2 work stealing: next warp-block of quads, per row parity
};
constexpr int SCRATCH_BYTES = (sizeof(Scratch) + 15) / 16 * 16;

__device__ __forceinline__ double warp_sum_d(double x) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
    return x;
}
__device__ __forceinline__ double warp_max_d(double x) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x = fmax(x, __shfl_xor_sync(0xffffffffu, x, o));
    return x;
}
__device__ __forceinline__ double warp_min_d(double x) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x = fmin(x, __shfl_xor_sync(0xffffffffu, x, o));
    return x;
}

// barrier 1 over the first `nthreads` threads of the CTA (the warps that own output values); the other
// warps never touch it and run ahead into the next row
__device__ __forceinline__ void group_sync(int nthreads) { asm volatile("bar.sync 1, %0;" ::"r"(nthreads) : "memory"); }
__device__ __forceinline__ float lds_f32(uint32_t addr) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
    return v;
}

// total of a per-thread int over the group, returned to every thread (2 barriers; slow path only)
__device__ __noinline__ int block_count(int local, Scratch* sc, int lane, int warp, int nthreads) {
    int w = __reduce_add_sync(0xffffffffu, local);
    if (lane == 0) sc->bcnt[warp] = w;
    group_sync(nthreads);
    if (warp == 0) {
        int t = lane < (nthreads >> 5) ? sc->bcnt[lane] : 0;
        t = __reduce_add_sync(0xffffffffu, t);
        if (lane == 0) sc->btotal = t;
    }
    group_sync(nthreads);
    return sc->btotal;
}

__device__ __forceinline__ unsigned long long ordered_bits(double v) {
    unsigned long long b = (unsigned long long)__double_as_longlong(v);
    return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}

// Exact median of the K finite values spread over the CTA's registers (np.median semantics,
// /root/reference/src/infercnvpy/tl/_infercnv.py:442).  Unused slots hold +inf (they sort last and
// never reach the middle ranks).  S1 / S2 = sum and sum of squares of the K values, known to every
// thread; they only steer the first bracket — exactness never depends on them.  Returns the median to
// every thread.  Must be called by exactly the first `nthreads` threads of the CTA (whole warps).
template <int VPT>
__device__ double block_median(const double (&v)[VPT], int K, double S1, double S2, Scratch* sc, int lane, int warp,
                               int nthreads) {
    const int nwarps = nthreads >> 5;
    // the median lies within one standard deviation of the mean; 2 % slack for the fp32 arithmetic
    const float invK = 1.f / (float)K;
    const float mean = (float)S1 * invK;
    const float var = fmaxf((float)S2 * invK - mean * mean, 0.f);
    const float half = fmaxf(1.02f * sqrtf(var) + 1e-6f * fabsf(mean), 1e-20f);
    const double kbase = (double)(mean - half);
    const double kscale = (double)(2147483648.f / half);
    // key(v) = saturating floor((v - kbase) * kscale): monotone in v, 32 bits; +inf -> 0xFFFFFFFF
    uint32_t key[VPT];
#pragma unroll
    for (int i = 0; i < VPT; ++i) key[i] = __double2uint_rd((v[i] - kbase) * kscale);

    const int r1 = (K - 1) >> 1, r2 = K >> 1;
    uint32_t klo = 0, ksplit = 0;
    int shift = 29, below = 0, state = 0, buf = 0;
    while (true) {
        uint32_t w0 = 0, w1 = 0, w2 = 0, w3 = 0;  // 16-bit fields: bins (0,1) (2,3) (4,5) (6,7)
#pragma unroll
        for (int g0 = 0; g0 < VPT; g0 += LOUT) {
            uint32_t c4 = 0;  // eight 4-bit counters (<= LOUT each)
#pragma unroll
            for (int i = g0; i < g0 + LOUT; ++i) {
                const uint32_t b = (key[i] - klo) >> shift;
                c4 += (b < 8u) ? (1u << (4u * b)) : 0u;
            }
            w0 += (c4 & 0xFu) | ((c4 & 0xF0u) << 12);
            w1 += ((c4 >> 8) & 0xFu) | ((c4 & 0xF000u) << 4);
            w2 += ((c4 >> 16) & 0xFu) | ((c4 >> 4) & 0xF0000u);
            w3 += ((c4 >> 24) & 0xFu) | ((c4 >> 12) & 0xF0000u);
        }
        w0 = __reduce_add_sync(0xffffffffu, w0);
        w1 = __reduce_add_sync(0xffffffffu, w1);
        w2 = __reduce_add_sync(0xffffffffu, w2);
        w3 = __reduce_add_sync(0xffffffffu, w3);
        if (lane == 0) sc->wcnt[buf][warp] = make_uint4(w0, w1, w2, w3);
        group_sync(nthreads);
        // every warp derives the decision itself (no second barrier): lane b < 8 owns bin b
        uint4 c = lane < nwarps ? sc->wcnt[buf][lane] : make_uint4(0, 0, 0, 0);
        buf ^= 1;
        c.x = __reduce_add_sync(0xffffffffu, c.x);
        c.y = __reduce_add_sync(0xffffffffu, c.y);
        c.z = __reduce_add_sync(0xffffffffu, c.z);
        c.w = __reduce_add_sync(0xffffffffu, c.w);
        const uint32_t word = (lane & 4) ? ((lane & 2) ? c.w : c.z) : ((lane & 2) ? c.y : c.x);
        const int cb = lane < 8 ? (int)((lane & 1) ? (word >> 16) : (word & 0xFFFFu)) : 0;
        int incl = cb;
#pragma unroll
        for (int o = 1; o < 8; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        incl += below;
        const unsigned m1 = __ballot_sync(0xffffffffu, lane < 8 && incl > r1);
        const unsigned m2 = __ballot_sync(0xffffffffu, lane < 8 && incl > r2);
        const int b1 = __ffs(m1) - 1, b2 = __ffs(m2) - 1;
        const int below1 = __shfl_sync(0xffffffffu, incl - cb, b1 & 31);
        const int n_in = __shfl_sync(0xffffffffu, cb, b1 & 31);
        if (b1 < 0 || b2 < 0) {  // cannot happen for finite input; take the exact slow path over everything
            state = 3;
            klo = 0;
            shift = 32;
            below = 0;
            break;
        }
        if (b1 != b2) {
            state = 2;
            ksplit = klo + ((uint32_t)b2 << shift);
            break;
        }
        klo += (uint32_t)b1 << shift;
        below = below1;
        if (n_in <= CAND_CAP) {
            state = 1;
            break;
        }
        if (shift == 0) {
            state = 3;
            break;
        }
        shift = shift >= 3 ? shift - 3 : 0;
    }

    double m;
    if (state == 1) {
        // <= CAND_CAP values share the final key range: rank them exactly in fp64
#pragma unroll
        for (int i = 0; i < VPT; ++i)
            if (((key[i] - klo) >> shift) == 0u) {
                const int slot = atomicAdd(&sc->cand_n, 1);
                sc->cand[slot] = v[i];
            }
        group_sync(nthreads);
        if (warp == 0) {
            const int n = sc->cand_n;
            const double mine = lane < n ? sc->cand[lane] : 0.0;
            int rank = 0;
            for (int j = 0; j < n; ++j) {
                const double o = sc->cand[j];
                rank += (o < mine) || (o == mine && j < lane);
            }
            if (lane < n) {
                if (rank == r1 - below) sc->med[0] = mine;
                if (rank == r2 - below) sc->med[1] = mine;
            }
        }
        group_sync(nthreads);
        m = (sc->med[0] + sc->med[1]) / 2.0;
    } else if (state == 2) {
        // the two middle ranks sit on either side of a bin boundary
        double lo = -INFINITY, hi = INFINITY;
#pragma unroll
        for (int i = 0; i < VPT; ++i) {
            if (key[i] < ksplit)
                lo = fmax(lo, v[i]);
            else
                hi = fmin(hi, v[i]);
        }
        lo = warp_max_d(lo);
        hi = warp_min_d(hi);
        if (lane == 0) {
            sc->wred[warp][0] = lo;
            sc->wred[warp][1] = hi;
        }
        group_sync(nthreads);
        lo = lane < nwarps ? sc->wred[lane][0] : -INFINITY;
        hi = lane < nwarps ? sc->wred[lane][1] : INFINITY;
        lo = warp_max_d(lo);
        hi = warp_min_d(hi);
        m = (lo + hi) / 2.0;
    } else {
        // more than CAND_CAP values collapse onto one 32-bit key (ties / degenerate rows):
        // exact radix select on the order-preserving 64-bit pattern, one bit per step
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
#pragma unroll
                for (int i = 0; i < VPT; ++i) {
                    const bool in_set = shift >= 32 ? true : (((key[i] - klo) >> shift) == 0u);
                    const unsigned long long ob = ordered_bits(v[i]);
                    const bool same_prefix = bit == 63 ? true : ((ob >> (bit + 1)) == (prefix >> (bit + 1)));
                    local += in_set && same_prefix && !((ob >> bit) & 1ull);
                }
                const int zeros = block_count(local, sc, lane, warp, nthreads);
                if (rr >= zeros) {
                    rr -= zeros;
                    prefix |= 1ull << bit;
                }
            }
            const unsigned long long b = (prefix >> 63) ? (prefix & 0x7FFFFFFFFFFFFFFFull) : ~prefix;
            res[which] = __longlong_as_double((long long)b);
        }
        m = (res[0] + res[1]) / 2.0;
    }
    return m;
}

// ------------------------------------------------------------------------------------------------
template <int TIER, int NWIN, int GS, bool BOUNDED, bool C64, int TPT>
__global__ void __launch_bounds__(NT, (TIER == 0 && TPT == 1) ? 2 : 1) smooth_kernel(const SmoothParams p) {
    extern __shared__ __align__(16) unsigned char smem[];
    Scratch* sc = reinterpret_cast<Scratch*>(smem);
    unsigned char* carve = smem + SCRATCH_BYTES;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr bool GROUPED = TIER < 2;
    constexpr int NQ_C = (TIER == 0) ? NWIN / GS : 0;
    constexpr bool M3_C = (TIER == 0) && ((NWIN / 2) % GS != 0);
    constexpr int QSTAR_C = M3_C ? (NWIN / 2) / GS : -1;
    static_assert(TIER != 0 || NWIN % 2 == 0, "tier 0 instantiations use even windows");
    constexpr int VPT = TPT * LOUT;

    // ---- carve shared memory
    float* raw = nullptr;
    double2* AB = nullptr;
    double* Cp = nullptr;
    double* w_alpha = nullptr;
    double* w_beta = nullptr;
    double* w_c = nullptr;
    void* buf = nullptr;
    double* wdir = nullptr;
    const int gs = (TIER == 0) ? GS : p.gs;
    const int NQ = (TIER == 0) ? NQ_C : p.NQ;
    const int qstar = (TIER == 0) ? QSTAR_C : p.qstar;
    if constexpr (GROUPED) {
        raw = reinterpret_cast<float*>(carve);
        carve += (size_t)p.Gpad * 4;
        AB = reinterpret_cast<double2*>(carve);
        carve += (size_t)(p.NGpad + PAD_GROUPS) * 16;
        if ((TIER == 0) ? M3_C : (p.qstar >= 0)) {
            Cp = reinterpret_cast<double*>(carve);
            carve += (size_t)(p.NGpad + PAD_GROUPS) * 8;
        }
        if constexpr (TIER == 1) {
            w_alpha = reinterpret_cast<double*>(carve);
            carve += (size_t)NQ * 8;
            w_beta = reinterpret_cast<double*>(carve);
            carve += (size_t)NQ * 8;
            w_c = reinterpret_cast<double*>(carve);
            carve += (size_t)gs * 8;
        }
    } else {
        wdir = reinterpret_cast<double*>(carve);
        carve += (size_t)p.window * 8;
        buf = carve;
    }

    // ---- one-time setup
    const bool dense = p.X != nullptr;
    if constexpr (GROUPED) {
        for (int i = p.G + tid; i < p.Gpad; i += NT) raw[i] = 0.f;
        for (int i = p.NGpad + tid; i < p.NGpad + PAD_GROUPS; i += NT) {
            AB[i] = make_double2(0.0, 0.0);
            if (Cp) Cp[i] = 0.0;
        }
        if constexpr (TIER == 1) {
            for (int i = tid; i < NQ; i += NT) {
                w_alpha[i] = p.alpha[i];
                w_beta[i] = p.beta[i];
            }
            for (int i = tid; i < gs; i += NT) w_c[i] = p.cw[i];
        }
        if (tid == 0) {
            mbar_init(&sc->mbar, 1);
            mbar_fence_init();
        }
    } else {
        for (int i = tid; i < p.window; i += NT) wdir[i] = p.wdir[i];
    }
    if (tid == 0) {
        sc->next_wb[0] = 0;
        sc->next_wb[1] = 0;
        sc->cand_n = 0;
    }
    __syncthreads();

    const uint64_t pol = l2_policy_evict_first();
    uint32_t parity = 0;
    const uint32_t row_bytes = (uint32_t)p.G * 4u;
    auto issue_row = [&](int64_t r) {
        const char* src = reinterpret_cast<const char*>(p.X + r * p.ldx);
        mbar_expect_tx(&sc->mbar, row_bytes);
        constexpr uint32_t CH = 16384;
        for (uint32_t off = 0; off < row_bytes; off += CH)
            bulk_g2s(reinterpret_cast<char*>(raw) + off, src + off, min(CH, row_bytes - off), &sc->mbar, pol);
    };

    int64_t row = blockIdx.x;
    const bool tma = GROUPED && dense && p.use_tma;
    if (tma && tid == 0 && row < p.n_rows) issue_row(row);

    const float clipf = p.clipf;
    const int nquads = p.NGpad >> 2;
    const int n_wb = (nquads + 31) >> 5;
    const double Kd = (double)p.K;
    const uint32_t raw_s = GROUPED ? smem_u32(raw) : 0u;
    // Only the warps that own output values ("group") take part in phase 3 / median / write-out; the others
    // go straight to the next row and work ahead on its gathers (barrier 1 = group only, barrier 0 = CTA).
    const int n_group = min(NT, ((p.n_tasks + 31) >> 5) << 5);
    const bool in_group = tid < n_group;
    int it = 0;

    for (; row < p.n_rows; row += gridDim.x, ++it) {
        // ======================= stage the raw row =======================
        if constexpr (GROUPED) {
            if (tma) {
                mbar_wait(&sc->mbar, parity);
                parity ^= 1u;
            } else if (dense) {
                const float* src = p.X + row * p.ldx;
                for (int i = tid; i < p.G; i += NT) raw[i] = __ldg(src + i);
                __syncthreads();
            } else {
                // CSR: densify on load (the reference densifies too, _infercnv.py:423)
                float4* r4 = reinterpret_cast<float4*>(raw);
                for (int i = tid; i < (p.Gpad >> 2); i += NT) r4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
                __syncthreads();
                const int64_t e0 = p.indptr[row], e1 = p.indptr[row + 1];
                for (int64_t e = e0 + tid; e < e1; e += NT) raw[__ldg(p.indices + e)] = __ldg(p.data + e);
                __syncthreads();
            }
        }

        // ======================= centre + clip + partial sums =======================
        if constexpr (GROUPED) {
            // warp-blocks of 32 quads are handed out dynamically: warps that are not in the group arrive
            // here early (they skipped the median of the previous row) and take most of them
            int* next_wb = &sc->next_wb[it & 1];
            while (true) {
                int wb = 0;
                if (lane == 0) wb = atomicAdd(next_wb, 1);
                wb = __shfl_sync(0xffffffffu, wb, 0);
                if (wb >= n_wb) break;
                const int quad = (wb << 5) + lane;
                if (quad >= nquads) continue;
                double a[4] = {0, 0, 0, 0}, b[4] = {0, 0, 0, 0}, c[4] = {0, 0, 0, 0};
                // table entry (wb, j, lane, u): ((wb*gs + j)*32 + lane)*4 + u
                const size_t tbase = ((size_t)wb * gs * 32 + lane) * 4;
                const uint32_t* ip = p.off_w + tbase;
                const float* lp = p.lo_w + tbase;
                const float* hp = p.hi_w + tbase;
                auto body = [&](int j, double cwj) {
                    const uint4 id = ldg_nc_u4(ip + j * 128);
                    const float4 lo = ldg_nc_f4(lp + j * 128);
                    float4 hi = lo;
                    if constexpr (BOUNDED) hi = ldg_nc_f4(hp + j * 128);
                    const float x[4] = {lds_f32(raw_s + id.x), lds_f32(raw_s + id.y), lds_f32(raw_s + id.z), lds_f32(raw_s + id.w)};
                    const float l4[4] = {lo.x, lo.y, lo.z, lo.w};
                    const float h4[4] = {hi.x, hi.y, hi.z, hi.w};
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        float d;
                        if constexpr (BOUNDED)
                            d = x[u] > h4[u] ? x[u] - h4[u] : (x[u] < l4[u] ? x[u] - l4[u] : 0.f);
                        else
                            d = x[u] - l4[u];
                        d = fminf(fmaxf(d, -clipf), clipf);
                        const double dd = (double)d;
                        a[u] += dd;
                        if (j > 0) b[u] = fma((double)j, dd, b[u]);
                        if (qstar >= 0) c[u] = fma(cwj, dd, c[u]);
                    }
                };
                if constexpr (TIER == 0) {
#pragma unroll
                    for (int j = 0; j < GS; ++j) body(j, M3_C ? (double)pyr(NWIN, GS * (QSTAR_C < 0 ? 0 : QSTAR_C) + j) : 0.0);
                } else {
                    for (int j = 0; j < gs; ++j) body(j, w_c[j]);
                }
                // thread's u-th group is u*nquads + quad: for fixed u the lanes store consecutive 16 B
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    AB[u * nquads + quad] = make_double2(a[u], b[u]);
                    if (qstar >= 0) Cp[u * nquads + quad] = c[u];
                }
            }
            __syncthreads();  // gathers done: raw row is dead, partials visible
            if (tid == 0) {
                *next_wb = 0;  // used again two rows from now
                if (tma && row + gridDim.x < p.n_rows) issue_row(row + gridDim.x);
            }
        } else {
            // direct tier: position-sorted centred row (float, or double for float64 centring)
            for (int s = tid; s < p.n_sorted; s += NT) {
                const float x = __ldg(p.X + row * p.ldx + p.idx_lin[s]);
                if constexpr (C64) {
                    const double lo = reinterpret_cast<const double*>(p.lo_lin)[s];
                    double d;
                    if constexpr (BOUNDED) {
                        const double hi = reinterpret_cast<const double*>(p.hi_lin)[s];
                        // bounded result is written into an array of the matrix dtype (:428): round to fp32
                        d = (double)x > hi ? (double)(float)((double)x - hi)
                                           : ((double)x < lo ? (double)(float)((double)x - lo) : 0.0);
                        d = (double)fminf(fmaxf((float)d, -clipf), clipf);
                    } else {
                        d = (double)x - lo;
                        d = fmin(fmax(d, -p.clip), p.clip);
                    }
                    reinterpret_cast<double*>(buf)[s] = d;
                } else {
                    const float lo = reinterpret_cast<const float*>(p.lo_lin)[s];
                    float d;
                    if constexpr (BOUNDED) {
                        const float hi = reinterpret_cast<const float*>(p.hi_lin)[s];
                        d = x > hi ? x - hi : (x < lo ? x - lo : 0.f);
                    } else {
                        d = x - lo;
                    }
                    reinterpret_cast<float*>(buf)[s] = fminf(fmaxf(d, -clipf), clipf);
                }
            }
            __syncthreads();
        }

        if (!in_group) {
            __syncthreads();  // partials / sorted row consumed by the group: safe to start the next row
            continue;
        }

        // ======================= windows =======================
        double v[VPT];
        int nv[TPT], col0[TPT];
        double s1 = 0.0, s2 = 0.0;
#pragma unroll
        for (int tt = 0; tt < TPT; ++tt) {
            nv[tt] = 0;
            col0[tt] = 0;
#pragma unroll
            for (int i = 0; i < LOUT; ++i) v[tt * LOUT + i] = INFINITY;
            const int ti = tid + tt * NT;
            if (ti < p.n_tasks) {
                const int4 t = __ldg(reinterpret_cast<const int4*>(p.tasks) + ti);
                col0[tt] = t.y;
                if ((t.w & 0xFF) == 0) {
                    nv[tt] = t.z;
                    if constexpr (TIER == 0) {
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
                                if (w >= 0 && w < NQ_C && w != QSTAR_C) {
                                    const int al = pyr(NWIN, GS * w);
                                    const int be = GS > 1 ? pyr(NWIN, GS * w + 1) - al : 0;
                                    acc[i] = fma((double)al, ab.x, acc[i]);
                                    if (be == 1)
                                        acc[i] += ab.y;
                                    else if (be == -1)
                                        acc[i] -= ab.y;
                                }
                            }
                        }
                        if constexpr (M3_C) {
#pragma unroll
                            for (int i = 0; i < LOUT; ++i) acc[i] += Cp[t.x + QSTAR_C + i];
                        }
#pragma unroll
                        for (int i = 0; i < LOUT; ++i)
                            if (i < t.z) v[tt * LOUT + i] = acc[i] * p.inv_sumw;
                    } else if constexpr (TIER == 1) {
#pragma unroll
                        for (int i = 0; i < LOUT; ++i) {
                            if (i < t.z) {
                                double acc = 0.0;
                                const double2* P = AB + t.x + i;
                                for (int q = 0; q < NQ; ++q) {
                                    const double2 ab = P[q];
                                    acc = fma(w_alpha[q], ab.x, acc);
                                    acc = fma(w_beta[q], ab.y, acc);
                                }
                                if (qstar >= 0) acc += Cp[t.x + i + qstar];
                                v[tt * LOUT + i] = acc * p.inv_sumw;
                            }
                        }
                    } else {
#pragma unroll
                        for (int i = 0; i < LOUT; ++i) {
                            if (i < t.z) {
                                double acc = 0.0;
                                const int s0 = t.x + i * p.step;
                                if constexpr (C64) {
                                    const double* B = reinterpret_cast<const double*>(buf) + s0;
                                    for (int j = 0; j < p.window; ++j) acc = fma(wdir[j], B[j], acc);
                                } else {
                                    const float* B = reinterpret_cast<const float*>(buf) + s0;
                                    for (int j = 0; j < p.window; ++j) acc = fma(wdir[j], (double)B[j], acc);
                                }
                                v[tt * LOUT + i] = acc * p.inv_sumw;
                            }
                        }
                    }
                } else {
                    // chromosome not longer than the window: one flat mean (_infercnv.py:227-236)
                    nv[tt] = 1;
                    double acc = 0.0;
                    if constexpr (GROUPED) {
                        for (int g = 0; g < t.z; ++g) acc += AB[t.x + g].x;
                    } else if constexpr (C64) {
                        for (int j = 0; j < t.z; ++j) acc += reinterpret_cast<const double*>(buf)[t.x + j];
                    } else {
                        for (int j = 0; j < t.z; ++j) acc += (double)reinterpret_cast<const float*>(buf)[t.x + j];
                    }
                    v[tt * LOUT] = acc * p.flat_inv[t.w >> 8];
                }
#pragma unroll
                for (int i = 0; i < LOUT; ++i)
                    if (i < nv[tt]) {
                        const double x = v[tt * LOUT + i];
                        s1 += x;
                        s2 = fma(x, x, s2);
                    }
            }
        }

        // ======================= row sums (one fp64 reduction), median, centring =======================
        s1 = warp_sum_d(s1);
        s2 = warp_sum_d(s2);
        if (lane == 0) {
            sc->wred[warp][0] = s1;
            sc->wred[warp][1] = s2;
            if (warp == 0) sc->cand_n = 0;
        }
        __syncthreads();  // CTA-wide: also tells the run-ahead warps that the partials have been read
        s1 = lane < (n_group >> 5) ? sc->wred[lane][0] : 0.0;
        s2 = lane < (n_group >> 5) ? sc->wred[lane][1] : 0.0;
        s1 = warp_sum_d(s1);
        s2 = warp_sum_d(s2);

        const double m = block_median<VPT>(v, p.K, s1, s2, sc, lane, warp, n_group);

#pragma unroll
        for (int tt = 0; tt < TPT; ++tt) {
            if (p.out_f64) {
                double* o = reinterpret_cast<double*>(p.out) + row * p.ldo + col0[tt];
#pragma unroll
                for (int i = 0; i < LOUT; ++i)
                    if (i < nv[tt]) o[i] = v[tt * LOUT + i] - m;
            } else {
                float* o = reinterpret_cast<float*>(p.out) + row * p.ldo + col0[tt];
#pragma unroll
                for (int i = 0; i < LOUT; ++i)
                    if (i < nv[tt]) o[i] = (float)(v[tt * LOUT + i] - m);
            }
        }
        if (tid == 0) {
            // sum(v - m) and sum((v - m)^2) from the raw moments
            p.row_stats[2 * row] = s1 - Kd * m;
            p.row_stats[2 * row + 1] = fma(Kd * m, m, fma(-2.0 * m, s1, s2));
        }
        // the next row's barriers order the reuse of `sc`
    }
}

// ------------------------------------------------------------------------------------------------
// instantiation table
template <int TIER, int NWIN, int GS, bool BOUNDED, bool C64, int TPT>
static int launch_one(const SmoothParams& p, int grid, size_t smem, cudaStream_t stream) {
    auto k = smooth_kernel<TIER, NWIN, GS, BOUNDED, C64, TPT>;
    ICNV_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k<<<grid, NT, smem, stream>>>(p);
    ICNV_CUDA(cudaGetLastError());
    return 0;
}
template <int TIER, int NWIN, int GS, bool BOUNDED, bool C64, int TPT>
static int occ_one(size_t smem, int* out) {
    auto k = smooth_kernel<TIER, NWIN, GS, BOUNDED, C64, TPT>;
    ICNV_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    ICNV_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(out, k, NT, smem));
    return 0;
}

#define ICNV_DISPATCH(FN, ...)                                                                             \
    do {                                                                                                   \
        if (tier == 0 && nwin == 100 && gs == 10 && tpt == 1) {                                            \
            return bounded ? FN<0, 100, 10, true, false, 1>(__VA_ARGS__) : FN<0, 100, 10, false, false, 1>(__VA_ARGS__); \
        }                                                                                                  \
        if (tier == 0 && nwin == 250 && gs == 10 && tpt == 1) {                                            \
            return bounded ? FN<0, 250, 10, true, false, 1>(__VA_ARGS__) : FN<0, 250, 10, false, false, 1>(__VA_ARGS__); \
        }                                                                                                  \
        if (tier == 1 && tpt == 1) {                                                                       \
            return bounded ? FN<1, 0, 0, true, false, 1>(__VA_ARGS__) : FN<1, 0, 0, false, false, 1>(__VA_ARGS__); \
        }                                                                                                  \
        if (tier == 1 && tpt == 4) {                                                                       \
            return bounded ? FN<1, 0, 0, true, false, 4>(__VA_ARGS__) : FN<1, 0, 0, false, false, 4>(__VA_ARGS__); \
        }                                                                                                  \
        if (tier == 2 && tpt == 1) {                                                                       \
            if (c64) return bounded ? FN<2, 0, 0, true, true, 1>(__VA_ARGS__) : FN<2, 0, 0, false, true, 1>(__VA_ARGS__); \
            return bounded ? FN<2, 0, 0, true, false, 1>(__VA_ARGS__) : FN<2, 0, 0, false, false, 1>(__VA_ARGS__); \
        }                                                                                                  \
        if (tier == 2 && tpt == 4) {                                                                       \
            if (c64) return bounded ? FN<2, 0, 0, true, true, 4>(__VA_ARGS__) : FN<2, 0, 0, false, true, 4>(__VA_ARGS__); \
            return bounded ? FN<2, 0, 0, true, false, 4>(__VA_ARGS__) : FN<2, 0, 0, false, false, 4>(__VA_ARGS__); \
        }                                                                                                  \
    } while (0)

int smooth_launch(int tier, int nwin, int gs, bool bounded, bool c64, int tpt, const SmoothParams& p, int grid,
                  size_t smem, cudaStream_t stream) {
    ICNV_DISPATCH(launch_one, p, grid, smem, stream);
    set_error("smooth_launch: no kernel instantiation for this configuration");
    return -3;
}
int smooth_occupancy(int tier, int nwin, int gs, bool bounded, bool c64, int tpt, size_t smem, int* ctas_per_sm) {
    ICNV_DISPATCH(occ_one, smem, ctas_per_sm);
    set_error("smooth_occupancy: no kernel instantiation for this configuration");
    return -3;
}

size_t smooth_scratch_bytes() { return SCRATCH_BYTES; }

}  // namespace icnv
