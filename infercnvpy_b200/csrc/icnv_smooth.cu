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
    uint4 wcnt[2][NW];      // generic 8-bin passes: per-warp packed counts (double-buffered by pass parity)
    uint4 fcnt[NW];         // fast first pass: 16 one-byte counters per warp
    int2 fmeta[NW];         // fast first pass: {below | in-region << 16, byte-overflow flag}
    double wred[2][NW][2];  // per-warp exact (sum, sum of squares) of the row, by row parity
    double wmm[NW][2];      // straddle path: per-warp max-below / min-above
    double cand[CAND_CAP];  // final candidates (+inf padded)
    double mrow[2];         // median of the row, by row parity (row statistics are finished one row later)
    long long prow[2];      // row index belonging to mrow / wred
    float2 wsumf[NW];       // per-warp fp32 (sum, sum of squares): steers the first bracket only
    unsigned long long mbar;
    int cand_n;
    int bcnt[NW];
    int btotal;
    int next_wb[2];         // phase-2 work stealing: next warp-block of quads, per row parity
};
constexpr int SCRATCH_BYTES = (sizeof(Scratch) + 15) / 16 * 16;

__device__ __forceinline__ double warp_sum_d(double x) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
    return x;
}
__device__ __forceinline__ float warp_sum_f(float x) {
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

// Exact median of the K finite values spread over the group's registers (np.median semantics,
// /root/reference/src/infercnvpy/tl/_infercnv.py:442).  Unused slots hold +inf (they sort last and
// never reach the middle ranks).  `mean` / `var` (fp32 estimates of the K values) only steer the first
// bracket — exactness never depends on them.  Returns the median to every thread.  Must be called by
// exactly the first `nthreads` threads of the CTA (whole warps); sc->cand must be +inf, sc->cand_n 0.
//
// Keys: key(v) = saturating floor((v - (mean - h)) * 2^31 / h), h = 1.02 sigma: 32 bits, monotone in v; the
// median always lies inside (|median - mean| <= sigma).  Pass 1 (one barrier) counts the 16 bins of width
// h/32 around the mean plus everything below them; for near-symmetric rows that already isolates <= 32
// candidates, which every warp then ranks exactly in fp64.  Anything else (median outside those bins,
// a crowded bin, ties) continues with generic 8-bin passes and, ultimately, an exact bitwise select.
template <int VPT>
__device__ double block_median(const double (&v)[VPT], int K, float mean, float var, Scratch* sc, int lane, int warp,
                               int nthreads, long long* dbg) {
#define MED_STAMP(k)                                   \
    do {                                               \
        if (dbg != nullptr) dbg[(k)] = clock64();      \
    } while (0)
    const int nwarps = nthreads >> 5;
    const float half = fmaxf(1.02f * sqrtf(var) + 1e-6f * fabsf(mean), 1e-20f);
    const double kbase = (double)(mean - half);
    const double kscale = (double)(2147483648.f / half);
    uint32_t key[VPT];
#pragma unroll
    for (int i = 0; i < VPT; ++i) key[i] = __double2uint_rd((v[i] - kbase) * kscale);

    MED_STAMP(8);
    const int r1 = (K - 1) >> 1, r2 = K >> 1;
    uint32_t klo = 0, ksplit = 0;
    int shift = 29, below = 0, state = -1, buf = 0;

#ifdef ICNV_FASTPASS  // measured slower on B200 (tools/ab.py, gpurun_out/ab1.log): off by default
    if constexpr (VPT == LOUT) {
        // ---- pass 1: 16 fine bins t = (key >> 26) - 24 in [0, 16), i.e. mean +- h/4
        uint32_t cA = 0, cB = 0;  // 4-bit counters for t = 0..7 and 8..15
        int nb = 0, nreg = 0;
#pragma unroll
        for (int i = 0; i < VPT; ++i) {
            const int t = (int)(key[i] >> 26) - 24;
            const bool in = (unsigned)t < 16u;
            const uint32_t inc = 1u << ((t & 7) << 2);
            cA += (in && t < 8) ? inc : 0u;
            cB += (in && t >= 8) ? inc : 0u;
            nb += t < 0;
            nreg += in;
        }
        // bytes: a0 = t 0,2,4,6; a1 = t 1,3,5,7; b0 = t 8,10,12,14; b1 = t 9,11,13,15
        uint32_t a0 = cA & 0x0F0F0F0Fu, a1 = (cA >> 4) & 0x0F0F0F0Fu, b0 = cB & 0x0F0F0F0Fu, b1w = (cB >> 4) & 0x0F0F0F0Fu;
        a0 = __reduce_add_sync(0xffffffffu, a0);
        a1 = __reduce_add_sync(0xffffffffu, a1);
        b0 = __reduce_add_sync(0xffffffffu, b0);
        b1w = __reduce_add_sync(0xffffffffu, b1w);
        const int meta = __reduce_add_sync(0xffffffffu, nb | (nreg << 16));
        if (lane == 0) {
            // a byte counter wraps if > 255 values of this warp share one bin (ties): detected by the byte sum
            const int bytesum = __dp4a(a0, 0x01010101u, 0u) + __dp4a(a1, 0x01010101u, 0u) + __dp4a(b0, 0x01010101u, 0u) +
                                __dp4a(b1w, 0x01010101u, 0u);
            sc->fcnt[warp] = make_uint4(a0, a1, b0, b1w);
            sc->fmeta[warp] = make_int2(meta, bytesum != (meta >> 16));
        }
        MED_STAMP(14);
        group_sync(nthreads);
        MED_STAMP(9);
        // every warp derives the decision itself: lane t < 16 owns fine bin t
        const int widx = ((lane >> 3) & 1) * 2 + (lane & 1), kb = ((lane & 7) >> 1) * 8;
        int f = 0, nbt = 0, bad = 0;
        for (int w = 0; w < nwarps; ++w) {
            const uint32_t word = reinterpret_cast<const uint32_t*>(&sc->fcnt[w])[widx];
            const int2 mt = sc->fmeta[w];
            f += (word >> kb) & 0xFFu;
            nbt += mt.x & 0xFFFF;
            bad |= mt.y;
        }
        if (lane >= 16) f = 0;
        int incl = f;
#pragma unroll
        for (int o = 1; o < 16; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        incl += nbt;
        const unsigned m1 = __ballot_sync(0xffffffffu, lane < 16 && incl > r1);
        const unsigned m2 = __ballot_sync(0xffffffffu, lane < 16 && incl > r2);
        const int b1 = __ffs(m1) - 1, b2 = __ffs(m2) - 1;
        if (!bad && r1 >= nbt && b1 >= 0 && b2 >= 0) {
            if (b1 != b2) {
                state = 2;
                ksplit = (uint32_t)(24 + b2) << 26;
            } else {
                klo = (uint32_t)(24 + b1) << 26;
                below = __shfl_sync(0xffffffffu, incl - f, b1);
                const int n_in = __shfl_sync(0xffffffffu, f, b1);
                if (n_in <= CAND_CAP) {
                    state = 1;
                    shift = 26;
                } else {
                    shift = 23;  // crowded bin: generic passes inside it
                }
            }
        }
    }
#endif

    while (state < 0) {
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

    MED_STAMP(10);
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
        MED_STAMP(11);
        // every warp ranks the (+inf padded) candidate list itself: lane i owns candidate i
        const int n = sc->cand_n;
        const double mine = sc->cand[lane];
        int rank = 0;
#pragma unroll
        for (int j = 0; j < CAND_CAP; ++j) {
            const double o = sc->cand[j];
            rank += (o < mine) || (o == mine && j < lane);
        }
        const unsigned q1 = __ballot_sync(0xffffffffu, lane < n && rank == r1 - below);
        const unsigned q2 = __ballot_sync(0xffffffffu, lane < n && rank == r2 - below);
        const double lo = __shfl_sync(0xffffffffu, mine, (__ffs(q1) - 1) & 31);
        const double hi = __shfl_sync(0xffffffffu, mine, (__ffs(q2) - 1) & 31);
        m = (lo + hi) / 2.0;
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
            sc->wmm[warp][0] = lo;
            sc->wmm[warp][1] = hi;
        }
        group_sync(nthreads);
        lo = lane < nwarps ? sc->wmm[lane][0] : -INFINITY;
        hi = lane < nwarps ? sc->wmm[lane][1] : INFINITY;
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
    MED_STAMP(12);
#undef MED_STAMP
    return m;
}

// ------------------------------------------------------------------------------------------------
template <int TIER, int NWIN, int GS, bool BOUNDED, bool C64, int TPT>
#ifdef ICNV_AB_OCC1
__global__ void __launch_bounds__(NT, 1) smooth_kernel(const SmoothParams p) {
#else
__global__ void __launch_bounds__(NT, (TIER == 0 && TPT == 1) ? 2 : 1) smooth_kernel(const SmoothParams p) {
#endif
    extern __shared__ __align__(16) unsigned char smem[];
    Scratch* sc = reinterpret_cast<Scratch*>(smem);
    unsigned char* carve = smem + SCRATCH_BYTES;

    // Logical warp numbering is reversed: the warps that own output values (the per-row critical path:
    // windows, median, write-out) are the HIGHEST physical warp ids, which the warp scheduler favours;
    // the run-ahead gather warps take the low ids.
    const int lane = threadIdx.x & 31;
#ifdef ICNV_AB_NOFLIP
    const int warp = (int)(threadIdx.x >> 5);
#else
    const int warp = NW - 1 - (int)(threadIdx.x >> 5);
#endif
    const int tid = warp * 32 + lane;
    constexpr int ISSUER = NT - 32;  // lane 0 of the last logical warp (a run-ahead warp) drives the TMA
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
        for (int i = p.NG + tid; i < p.NGpad + PAD_GROUPS; i += NT) {
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
        if (GROUPED && (smem_u32(raw) & 0xFFFFFFu) != p.raw_base) __trap();  // host baked a different base
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
    if (tma && tid == ISSUER && row < p.n_rows) issue_row(row);

    const float clipf = p.clipf;
    const int nquads = p.NGpad >> 2;
    const int n_wb = (nquads + 31) >> 5;
    const double Kd = (double)p.K;
    // Only the warps that own output values ("group") take part in phase 3 / median / write-out; the others
    // go straight to the next row and work ahead on its gathers (barrier 1 = group only, barrier 0 = CTA).
    const int n_group = min(NT, ((p.n_tasks + 31) >> 5) << 5);
    const bool in_group = tid < n_group;
    // task descriptors are row-invariant: fetch them once
    int4 task[TPT];
#pragma unroll
    for (int tt = 0; tt < TPT; ++tt) {
        const int ti = tid + tt * NT;
        task[tt] = ti < p.n_tasks ? __ldg(reinterpret_cast<const int4*>(p.tasks) + ti) : make_int4(0, 0, 0, 0);
    }
    // sum(v - m) and sum((v - m)^2) of a finished row from its raw moments (thread 0, one row late)
    auto finish_row_stats = [&](int par) {
        double S1 = 0.0, S2 = 0.0;
        for (int w = 0; w < (n_group >> 5); ++w) {
            S1 += sc->wred[par][w][0];
            S2 += sc->wred[par][w][1];
        }
        const double mm = sc->mrow[par];
        const long long r = sc->prow[par];
        p.row_stats[2 * r] = S1 - Kd * mm;
        p.row_stats[2 * r + 1] = fma(Kd * mm, mm, fma(-2.0 * mm, S1, S2));
    };
    int it = 0;
    // developer timeline (tools/timeline.py): stamps by thread 0 (group) and by the last warp (run-ahead)
#define ICNV_STAMP(slot)                                                                                  \
    do {                                                                                                  \
        if (p.dbg != nullptr && it < p.dbg_rows && tid == 0)                                              \
            p.dbg[((size_t)blockIdx.x * p.dbg_rows + it) * 16 + (slot)] = clock64();                       \
    } while (0)

    for (; row < p.n_rows; row += gridDim.x, ++it) {
        ICNV_STAMP(0);
        // ======================= stage the raw row =======================
        if constexpr (GROUPED) {
            if (tma) {
                mbar_wait(&sc->mbar, parity);
                parity ^= 1u;
                ICNV_STAMP(1);
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
#ifdef ICNV_AB_GROUPNOSTEAL
            while (!in_group || n_group == NT || !tma) {
#else
            while (true) {
#endif
                int wb = 0;
                if (lane == 0) wb = atomicAdd(next_wb, 1);
                wb = __shfl_sync(0xffffffffu, wb, 0);
                if (wb >= n_wb) break;
                const int quad = (wb << 5) + lane;  // slot index; every slot of every warp-block is valid
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
                    // table entries are complete shared-window addresses (raw base baked in by the host)
                    const float x[4] = {lds_f32(id.x), lds_f32(id.y), lds_f32(id.z), lds_f32(id.w)};
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
                // lane l owns groups with g % 8 == l % 8: a quarter-warp's 16-byte stores hit 8 bank groups
                const int4 gid = __ldg(reinterpret_cast<const int4*>(p.grp_w) + quad);
                const int gq[4] = {gid.x, gid.y, gid.z, gid.w};
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    AB[gq[u]] = make_double2(a[u], b[u]);
                    if (qstar >= 0) Cp[gq[u]] = c[u];
                }
            }
            ICNV_STAMP(2);
                __syncthreads();  // gathers done: raw row is dead, partials visible
            ICNV_STAMP(3);
                if (tid == ISSUER) {
                *next_wb = 0;  // used again two rows from now
                if (tma && row + gridDim.x < p.n_rows) issue_row(row + gridDim.x);
            }
            if (tid == 0 && it > 0) finish_row_stats((it - 1) & 1);
            ICNV_STAMP(13);
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
            if (tid == 0 && it > 0) finish_row_stats((it - 1) & 1);
        }

        if (!in_group) {
            __syncthreads();  // partials / sorted row consumed by the group: safe to start the next row
                continue;
        }

        // ======================= windows =======================
        double v[VPT];
        int nv[TPT];
        double s1 = 0.0, s2 = 0.0;
#pragma unroll
        for (int tt = 0; tt < TPT; ++tt) {
            nv[tt] = 0;
#pragma unroll
            for (int i = 0; i < LOUT; ++i) v[tt * LOUT + i] = INFINITY;
            const int ti = tid + tt * NT;
            if (ti < p.n_tasks) {
                const int4 t = task[tt];
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

        ICNV_STAMP(15);
        // ======================= bracket estimate, median, centring =======================
        {
            const float f1 = warp_sum_f((float)s1), f2 = warp_sum_f((float)s2);
            if (lane == 0) sc->wsumf[warp] = make_float2(f1, f2);
            if (warp == 0) {
                sc->cand[lane] = INFINITY;
                if (lane == 0) sc->cand_n = 0;
            }
        }
        ICNV_STAMP(4);
        __syncthreads();  // CTA-wide: also tells the run-ahead warps that the partials have been read
        float t1 = 0.f, t2 = 0.f;
        for (int w = 0; w < (n_group >> 5); ++w) {
            const float2 q = sc->wsumf[w];
            t1 += q.x;
            t2 += q.y;
        }
        const float invK = 1.f / (float)p.K;
        const float mean_f = t1 * invK;
        const float var_f = fmaxf(t2 * invK - mean_f * mean_f, 0.f);
        ICNV_STAMP(5);
        long long* mdbg = (p.dbg != nullptr && it < p.dbg_rows && tid == 0) ? p.dbg + ((size_t)blockIdx.x * p.dbg_rows + it) * 16 : nullptr;
        const double m = block_median<VPT>(v, p.K, mean_f, var_f, sc, lane, warp, n_group, mdbg);
        ICNV_STAMP(6);

#pragma unroll
        for (int tt = 0; tt < TPT; ++tt) {
            // tile order: the 32 lanes of a warp write 32 consecutive values per instruction
            const int ti = tid + tt * NT;
            const size_t pos = (size_t)row * p.ldo + (size_t)(ti >> 5) * (32 * LOUT) + (ti & 31);
            if (ti < ((p.n_tasks + 31) & ~31)) {
                if (p.out_f64) {
                    double* o = reinterpret_cast<double*>(p.out) + pos;
#pragma unroll
                    for (int i = 0; i < LOUT; ++i) o[i * 32] = v[tt * LOUT + i] - m;
                } else {
                    float* o = reinterpret_cast<float*>(p.out) + pos;
#pragma unroll
                    for (int i = 0; i < LOUT; ++i) o[i * 32] = (float)(v[tt * LOUT + i] - m);
                }
            }
        }
        // exact row moments: reduced per warp here, finished by thread 0 after the next CTA-wide barrier
        {
            const double e1 = warp_sum_d(s1), e2 = warp_sum_d(s2);
            if (lane == 0) {
                sc->wred[it & 1][warp][0] = e1;
                sc->wred[it & 1][warp][1] = e2;
                if (warp == 0) {
                    sc->mrow[it & 1] = m;
                    sc->prow[it & 1] = row;
                }
            }
        }
        ICNV_STAMP(7);
        // the next row's barriers order the reuse of `sc`
    }
#undef ICNV_STAMP
    __syncthreads();
    if (tid == 0 && it > 0) finish_row_stats((it - 1) & 1);
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

// Shared-window address of raw[0] for the grouped kernels (dynamic smem base + scratch block).  The host
// bakes it into the gather table; the kernels trap if their own value differs.
__global__ void probe_smem_kernel(uint32_t* out) {
    extern __shared__ __align__(16) unsigned char smem[];
    if (threadIdx.x == 0) *out = smem_u32(smem) & 0xFFFFFFu;
}
int smooth_raw_base(uint32_t* base) {
    static uint32_t cached = 0;
    static bool have = false;
    if (!have) {
        uint32_t* d = nullptr;
        ICNV_CUDA(cudaMalloc(&d, sizeof(uint32_t)));
        probe_smem_kernel<<<1, 32, 64>>>(d);
        ICNV_CUDA(cudaGetLastError());
        uint32_t h = 0;
        ICNV_CUDA(cudaMemcpy(&h, d, sizeof(uint32_t), cudaMemcpyDeviceToHost));
        cudaFree(d);
        cached = h + (uint32_t)SCRATCH_BYTES;
        have = true;
    }
    *base = cached;
    return 0;
}

}  // namespace icnv
