// Smoothing kernel: steps 1-3 of /root/reference/src/infercnvpy/tl/_infercnv.py:411-440
// (centre, clip, per-chromosome pyramid running mean decimated by `step`) for one cell row per CTA
// iteration, persistent CTAs.  Step 4 (row median) and step 5 (noise filter) are icnv_aux.cu's
// center_rows_kernel / threshold kernel: an exact median is a long serial chain per row, and inside this
// kernel it held every row's CTA for ~7k cycles (ablation in profiles/ablation_r1.txt: 42 % -> 58 % of the HBM
// roofline without it); one warp per row with thousands of independent warps hides that latency.
//
// Data flow per row (tiers 0/1, "grouped"):
//   HBM --cp.async.bulk (TMA, mbarrier)--> smem raw row [G] fp32
//   phase 2: every thread owns 4 position-ordered groups of `gs` (= step) genes; it gathers its genes
//            from the raw row through a table of pre-baked shared-memory addresses (warp-coalesced
//            16-byte table loads, immediate offsets), centres + clips in fp32 exactly like numpy, and
//            accumulates per-group partial sums in fp64:
//              A_g = sum_j x_j,  B_g = sum_j j*x_j,  C_g = sum_j cw_j*x_j (peak group only)
//            Within a step-aligned group the pyramid weights are linear in j, so every window is
//              out_k = sum_q alpha_q*A_{k+q} + beta_q*B_{k+q}  (+ C_{k+q*})
//            i.e. 2*window/step FMAs instead of `window` — and nothing is computed for the 90 % of
//            windows the reference computes and then drops (:215-218).
//            Warp-blocks of 32x4 groups are handed out through an smem atomic; warps that own no
//            outputs skip phase 3 and run ahead into the next row's gathers.
//   phase 3: one thread per LOUT=9 consecutive outputs slides over the partials (fp64 FMAs, weights
//            are compile-time immediates in tier 0) and writes them in warp-tile order (256-byte lines).
// Tier 2 ("direct": every other (window, step), float64 centring, gene axes longer than shared memory) is
// icnv_direct.cu.
#include "icnv_common.cuh"

namespace icnv {

struct __align__(16) Scratch {
    unsigned long long mbar;
    int next_wb[2];  // phase-2 work stealing: next warp-block of quads, per row parity
    int pad[2];
};
constexpr int SCRATCH_BYTES = (sizeof(Scratch) + 15) / 16 * 16;


__device__ __forceinline__ float lds_f32(uint32_t addr) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
    return v;
}

// ------------------------------------------------------------------------------------------------
// ROWS = cell rows staged and gathered together per CTA iteration (grouped tiers only).  ROWS == 2 reads every
// gather-table entry once for two rows — the tables (160 KB per sweep, L2 -> L1) are the largest removable share of the
// kernel's l1tex work — at the price of one CTA per SM (2 x 80 KB of staged rows + 2 x 32 KB of partials).
// DBUF (single staged row only): the partial sums are double-buffered by iteration parity, so warps that own no outputs
// gather the WHOLE next row while the group warps are still in phase 3 — no hand-over barriers at all.  One CTA per SM
// (row + 2 x partials), which is how the window-250 kernel (third partial sum per group, no room for row pairs) runs.
template <int TIER, int NWIN, int GS, bool BOUNDED, bool C64, int TPT, int ROWS, bool DBUF>
__global__ void __launch_bounds__(smooth_threads(ROWS, DBUF), (TIER == 0 && TPT == 1 && ROWS == 1 && !DBUF && (NWIN / 2) % (GS > 0 ? GS : 1) == 0) ? 2 : 1)
    smooth_kernel(const SmoothParams p) {
    static_assert(!DBUF || ROWS == 1, "double-buffered partial sums exist for the single-row kernel");
    extern __shared__ __align__(16) unsigned char smem[];
    constexpr int NTH = smooth_threads(ROWS, DBUF), NWH = NTH / 32;  // row pairs run 32 warps (64 registers each)
    Scratch* sc = reinterpret_cast<Scratch*>(smem);
    unsigned char* carve = smem + SCRATCH_BYTES;

    // Logical warp numbering is reversed: the warps that own output values (the per-row critical path:
    // windows, median, write-out) are the HIGHEST physical warp ids, which the warp scheduler favours;
    // the run-ahead gather warps take the low ids.
    const int lane = threadIdx.x & 31;
    const int warp = NWH - 1 - (int)(threadIdx.x >> 5);
    const int tid = warp * 32 + lane;
    constexpr int ISSUER = NTH - 32;  // lane 0 of the last logical warp (a run-ahead warp) drives the TMA
    static_assert(TIER < 2 && !C64, "the direct form (tier 2, float64 centring) lives in icnv_direct.cu");
    constexpr int NQ_C = (TIER == 0) ? NWIN / GS : 0;
    constexpr bool M3_C = (TIER == 0) && ((NWIN / 2) % GS != 0);
    constexpr int QSTAR_C = M3_C ? (NWIN / 2) / GS : -1;
    static_assert(TIER != 0 || NWIN % 2 == 0, "tier 0 instantiations use even windows");
    constexpr int VPT = TPT * LOUT;
    // groups per lane in one phase-2 work unit.  Row pairs keep whole warp-blocks: a run-ahead warp can only gather ONE
    // unit before it has to wait for the partial-sum buffer, so bigger units overlap more of phase 3.
    constexpr int UW = ICNV_UNIT_WIDTH(ROWS, M3_C);
    // permuted walk (icnv_schedule.cu): step t of a lane reads element j = (entry >> 24) & 15 of its group, not element t;
    // with a peak group bits 28..31 carry m_j = cw_j - cw_0, the non-linear part of the weights inside that group
    // ICNV_NATURAL_WALK (icnv_common.cuh): 1 = kernels with a peak group walk their groups in natural order (j and m_j are
    // compile-time immediates: no per-entry decode, at the price of ~2 instead of ~1.2 wavefronts per gather), 2 = all
    constexpr bool PERM = (TIER == 0) && !(ICNV_NATURAL_WALK == 2 || (ICNV_NATURAL_WALK == 1 && M3_C));
    // row pairs with a peak group: the third partial sum lives in a global (L2) scratch, see SmoothParams::c_scratch
    constexpr bool CGLOBAL = M3_C && ROWS == 2;
#define ICNV_ABS (p.NGpad + PAD_GROUPS) /* partial-sum slots per staged row */

    // ---- carve shared memory
    float* raw = nullptr;
    double2* AB = nullptr;
    double* Cp = nullptr;
    double* w_alpha = nullptr;
    double* w_beta = nullptr;
    double* w_c = nullptr;
    const int gs = (TIER == 0) ? GS : p.gs;
    const int NQ = (TIER == 0) ? NQ_C : p.NQ;
    const int qstar = (TIER == 0) ? QSTAR_C : p.qstar;
    {
        raw = reinterpret_cast<float*>(carve);
        carve += (size_t)ROWS * p.Gpad * 4;
        AB = reinterpret_cast<double2*>(carve);
        carve += (size_t)(DBUF ? 2 : ROWS) * ICNV_ABS * 16;
        if constexpr (CGLOBAL) {
            Cp = p.c_scratch + (size_t)blockIdx.x * 2 * ROWS * ICNV_ABS;
        } else if ((TIER == 0) ? M3_C : (p.qstar >= 0)) {
            Cp = reinterpret_cast<double*>(carve);
            carve += (size_t)(DBUF ? 2 : ROWS) * ICNV_ABS * 8;
        }
        if constexpr (TIER == 1) {
            w_alpha = reinterpret_cast<double*>(carve);
            carve += (size_t)NQ * 8;
            w_beta = reinterpret_cast<double*>(carve);
            carve += (size_t)NQ * 8;
            w_c = reinterpret_cast<double*>(carve);
            carve += (size_t)gs * 8;
        }
    }

    // ---- one-time setup
    const bool dense = p.X != nullptr;
    {
#pragma unroll
        for (int rr = 0; rr < ROWS; ++rr) {
            for (int i = p.G + tid; i < p.Gpad; i += NTH) raw[rr * p.Gpad + i] = 0.f;
            for (int i = p.NG + tid; i < ICNV_ABS; i += NTH) {
                AB[(ROWS > 1 ? rr * ICNV_ABS : 0) + i] = make_double2(0.0, 0.0);
                if (Cp) Cp[(ROWS > 1 ? rr * ICNV_ABS : 0) + i] = 0.0;
                if constexpr (CGLOBAL) Cp[(ROWS + rr) * ICNV_ABS + i] = 0.0;  // second parity buffer
                if constexpr (DBUF) {
                    AB[ICNV_ABS + i] = make_double2(0.0, 0.0);
                    if (Cp) Cp[ICNV_ABS + i] = 0.0;
                }
            }
        }
        if constexpr (TIER == 1) {
            for (int i = tid; i < NQ; i += NTH) {
                w_alpha[i] = p.alpha[i];
                w_beta[i] = p.beta[i];
            }
            for (int i = tid; i < gs; i += NTH) w_c[i] = p.cw[i];
        }
        if (tid == 0) {
            mbar_init(&sc->mbar, 1);
            mbar_fence_init();
        }
    }
    if (tid == 0) {
        sc->next_wb[0] = 0;
        sc->next_wb[1] = 0;
        if ((smem_u32(raw) & 0xFFFFFFu) != p.raw_base) __trap();  // host baked a different base
    }
    __syncthreads();

    const uint64_t pol = l2_policy_evict_first();
    uint32_t parity = 0;
    const uint32_t row_bytes = (uint32_t)p.G * 4u;
#define ICNV_ROW_OFF ((uint32_t)p.Gpad * 4u) /* byte distance between the staged rows of a pair */
    auto issue_row = [&](int64_t r) {
        const int nvalid = (int)min((int64_t)ROWS, p.n_rows - r);
        mbar_expect_tx(&sc->mbar, row_bytes * (uint32_t)nvalid);
        constexpr uint32_t CH = 16384;
        for (int rr = 0; rr < nvalid; ++rr) {
            const char* src = reinterpret_cast<const char*>(p.X + (r + rr) * p.ldx);
            char* dst = reinterpret_cast<char*>(raw) + (size_t)rr * ICNV_ROW_OFF;
            for (uint32_t off = 0; off < row_bytes; off += CH)
                bulk_g2s(dst + off, src + off, min(CH, row_bytes - off), &sc->mbar, pol);
        }
    };

    // The staged rows are single-buffered (no room for a second set), so the fill of iteration i+1 cannot start before
    // the gathers of iteration i are done.  HBM is kept streaming anyway by prefetching the rows of iteration i+1 into L2
    // while iteration i is still being gathered; the later shared-memory fill is an L2 hit.
    auto prefetch_row = [&](int64_t r) {
        const int nvalid = (int)min((int64_t)ROWS, p.n_rows - r);
        constexpr uint32_t CH = 16384;
        for (int rr = 0; rr < nvalid; ++rr) {
            const char* src = reinterpret_cast<const char*>(p.X + (r + rr) * p.ldx);
            for (uint32_t off = 0; off < row_bytes; off += CH) bulk_prefetch_l2(src + off, min(CH, row_bytes - off));
        }
    };

    int64_t row = (int64_t)blockIdx.x * ROWS;  // first row of this iteration's group of ROWS
    const bool tma = dense && p.use_tma;
    if (tma && tid == ISSUER && row < p.n_rows) {
        issue_row(row);
        if (p.l2_prefetch && row + (int64_t)gridDim.x * ROWS < p.n_rows) prefetch_row(row + (int64_t)gridDim.x * ROWS);
    }

    const float clipf = p.clipf;
    const int nquads = p.NGpad >> 2;
    const int n_units = ((nquads + 31) >> 5) * (4 / UW);  // phase-2 work units handed out through next_wb
    // Only the warps that own output values ("group") take part in phase 3 / write-out; the others go straight to
    // the next iteration and work ahead on its gathers.  With row pairs and enough warps every staged row has its own
    // set of group warps (split_rows), which halves the length of phase 3.
    //   barrier 0 (CTA)      : gathers of this iteration done -> partials complete, staged rows dead
    //   barrier 1 (CTA)      : group warps ARRIVE when they have read the partials of iteration i; the other warps SYNC
    //                          on it before their first partial-sum store of iteration i+1
    //   barrier 2 (group)    : the same hand-over among the group warps themselves
    const int n_group1 = ((p.n_tasks + 31) >> 5) << 5;
    const bool split_rows = ROWS > 1 && TPT == 1 && ROWS * n_group1 <= NTH && p.split_rows;
    const int n_group = min(NTH, split_rows ? ROWS * n_group1 : n_group1);
    const bool in_group = tid < n_group;
    const int my_rr = split_rows ? tid / n_group1 : 0;       // staged row whose outputs this thread computes
    const int tix = split_rows ? tid - my_rr * n_group1 : tid;  // its first task
    // task descriptors are row-invariant: fetch them once
    int4 task[TPT];
#pragma unroll
    for (int tt = 0; tt < TPT; ++tt) {
        const int ti = tix + tt * NTH;
        task[tt] = ti < p.n_tasks ? __ldg(reinterpret_cast<const int4*>(p.tasks) + ti) : make_int4(0, 0, 0, 0);
    }
    int it = 0;
    // developer timeline (tools/timeline.py): stamps by thread 0 (group) and by the last warp (run-ahead)
#define ICNV_STAMP(slot)                                                                                  \
    do {                                                                                                  \
        if (p.dbg != nullptr && it < p.dbg_rows && tid == 0)                                              \
            p.dbg[((size_t)blockIdx.x * p.dbg_rows + it) * 16 + (slot)] = clock64();                       \
    } while (0)

    for (; row < p.n_rows; row += (int64_t)gridDim.x * ROWS, ++it) {
        ICNV_STAMP(0);
        // ======================= stage the raw row =======================
        {
            if (tma) {
                mbar_wait(&sc->mbar, parity);
                parity ^= 1u;
                ICNV_STAMP(1);
            } else if (dense) {
                for (int rr = 0; rr < ROWS && row + rr < p.n_rows; ++rr) {
                    const float* src = p.X + (row + rr) * p.ldx;
                    for (int i = tid; i < p.G; i += NTH) raw[rr * p.Gpad + i] = __ldg(src + i);
                }
                __syncthreads();
            } else {
                // CSR: densify on load (the reference densifies too, _infercnv.py:423)
                float4* r4 = reinterpret_cast<float4*>(raw);
                for (int i = tid; i < ROWS * (p.Gpad >> 2); i += NTH) r4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
                __syncthreads();
                for (int rr = 0; rr < ROWS && row + rr < p.n_rows; ++rr) {
                    const int64_t e0 = p.indptr[row + rr], e1 = p.indptr[row + rr + 1];
                    for (int64_t e = e0 + tid; e < e1; e += NTH) raw[rr * p.Gpad + __ldg(p.indices + e)] = __ldg(p.data + e);
                }
                __syncthreads();
            }
        }

        // ======================= centre + clip + partial sums =======================
        {
            // work units are handed out dynamically: warps that are not in the group arrive here early (they skipped
            // phase 3 of the previous iteration) and take most of them
            int* next_wb = &sc->next_wb[it & 1];
            bool handed_over = DBUF || in_group || it == 0;  // group warps synchronise on barrier 2 after their phase 3
            const int pbuf = DBUF ? (it & 1) * ICNV_ABS : 0;  // partial-sum buffer of this iteration
            while (true) {
                int wb = 0;
                if (lane == 0) wb = atomicAdd(next_wb, 1);
                wb = __shfl_sync(0xffffffffu, wb, 0);
                if (wb >= n_units) break;
                // `wb` is a work unit: 32 lanes x UW groups (a whole warp-block of quads for row pairs, half of one else)
                const int pairslot = (wb << 5) + lane;  // slot index; every slot of every unit is valid
                double a[ROWS][UW], b[ROWS][UW], c[ROWS][UW];
#pragma unroll
                for (int rr = 0; rr < ROWS; ++rr)
#pragma unroll
                    for (int u = 0; u < UW; ++u) a[rr][u] = b[rr][u] = c[rr][u] = 0.0;
                // table entry (unit, j, lane, u): ((unit*gs + j)*32 + lane)*UW + u
                const size_t tbase = ((size_t)wb * gs * 32 + lane) * UW;
                const uint32_t* ip = p.off_w + tbase;
                const float* lp = p.lo_w + tbase;
                const float* hp = p.hi_w + tbase;
                auto body = [&](int j, double cwj) {
                    // table entries are complete shared-window addresses (raw base baked in by the host); the second
                    // row of a pair sits row_off bytes further
                    float l4[UW], h4[UW];
                    uint32_t ad[UW];
                    if constexpr (UW == 4) {
                        const uint4 id = ldg_nc_u4(ip + j * (32 * UW));
                        const float4 lo = ldg_nc_f4(lp + j * (32 * UW));
                        float4 hi = lo;
                        if constexpr (BOUNDED) hi = ldg_nc_f4(hp + j * (32 * UW));
                        ad[0] = id.x, ad[1] = id.y, ad[2] = id.z, ad[3] = id.w;
                        l4[0] = lo.x, l4[1] = lo.y, l4[2] = lo.z, l4[3] = lo.w;
                        h4[0] = hi.x, h4[1] = hi.y, h4[2] = hi.z, h4[3] = hi.w;
                    } else {
                        const uint2 id = ldg_nc_u2(ip + j * (32 * UW));
                        const float2 lo = ldg_nc_f2(lp + j * (32 * UW));
                        float2 hi = lo;
                        if constexpr (BOUNDED) hi = ldg_nc_f2(hp + j * (32 * UW));
                        ad[0] = id.x, ad[1] = id.y;
                        l4[0] = lo.x, l4[1] = lo.y;
                        h4[0] = hi.x, h4[1] = hi.y;
                    }
                    float x[ROWS][UW];
                    double jd[UW], md[UW];
                    if constexpr (PERM) {
#pragma unroll
                        for (int u = 0; u < UW; ++u) {
                            // exact int -> double without I2F: 2^52 + j carries j in the low mantissa bits
                            jd[u] = __hiloint2double(0x43300000, (int)((ad[u] >> 24) & 15u)) - 4503599627370496.0;
                            if constexpr (M3_C) md[u] = __hiloint2double(0x43300000, (int)(ad[u] >> 28)) - 4503599627370496.0;
                            ad[u] &= 0x00FFFFFFu;
                        }
                    }
#pragma unroll
                    for (int rr = 0; rr < ROWS; ++rr) {
                        const uint32_t ro = rr ? ICNV_ROW_OFF : 0u;
#pragma unroll
                        for (int u = 0; u < UW; ++u) x[rr][u] = lds_f32(ad[u] + ro);
                    }
#pragma unroll
                    for (int rr = 0; rr < ROWS; ++rr)
#pragma unroll
                        for (int u = 0; u < UW; ++u) {
                            const float xv = x[rr][u];
                            float d;
                            if constexpr (BOUNDED)
                                d = xv > h4[u] ? xv - h4[u] : (xv < l4[u] ? xv - l4[u] : 0.f);
                            else
                                d = xv - l4[u];
                            d = fminf(fmaxf(d, -clipf), clipf);
                            const double dd = (double)d;
                            if (TIER == 0 && j == 0) {  // first step of the (unrolled) walk: no add to zero
                                a[rr][u] = dd;
                                if constexpr (PERM) b[rr][u] = jd[u] * dd;
                                if constexpr (PERM && M3_C) c[rr][u] = md[u] * dd;
                                else if (qstar >= 0) c[rr][u] = cwj * dd;
                            } else {
                                a[rr][u] += dd;
                                if constexpr (PERM)
                                    b[rr][u] = fma(jd[u], dd, b[rr][u]);
                                else if (j > 0)
                                    b[rr][u] = fma((double)j, dd, b[rr][u]);
                                if constexpr (PERM && M3_C) c[rr][u] = fma(md[u], dd, c[rr][u]);
                                else if (qstar >= 0 && (TIER != 0 || cwj != 0.0)) c[rr][u] = fma(cwj, dd, c[rr][u]);
                            }
                        }
                };
                if constexpr (TIER == 0) {
#pragma unroll
                    for (int j = 0; j < GS; ++j)  // cwj: the peak group's non-linear weight m_j (natural walk only)
                        body(j, M3_C ? (double)(pyr(NWIN, GS * (M3_C ? QSTAR_C : 0) + j) - pyr(NWIN, GS * (M3_C ? QSTAR_C : 0))) : 0.0);
                } else {
                    for (int j = 0; j < gs; ++j) body(j, w_c[j]);
                }
                // lane l owns groups with g % 8 == l % 8: a quarter-warp's 16-byte stores hit 8 bank groups
                int gq[UW];
                if constexpr (UW == 4) {
                    const int4 gid = __ldg(reinterpret_cast<const int4*>(p.grp_w) + pairslot);
                    gq[0] = gid.x, gq[1] = gid.y, gq[2] = gid.z, gq[3] = gid.w;
                } else {
                    const int2 gid = __ldg(reinterpret_cast<const int2*>(p.grp_w) + pairslot);
                    gq[0] = gid.x, gq[1] = gid.y;
                }
                if (!handed_over) {
                    named_bar_sync(1, NTH);  // the group has read the previous iteration's partials
                    handed_over = true;
                }
#pragma unroll
                for (int rr = 0; rr < ROWS; ++rr)
#pragma unroll
                    for (int u = 0; u < UW; ++u) {
                        AB[(ROWS > 1 ? rr * ICNV_ABS : pbuf) + gq[u]] = make_double2(a[rr][u], b[rr][u]);
                        if constexpr (CGLOBAL)
                            __stcg(Cp + ((it & 1) * ROWS + rr) * ICNV_ABS + gq[u], c[rr][u]);
                        else if (qstar >= 0)
                            Cp[(ROWS > 1 ? rr * ICNV_ABS : pbuf) + gq[u]] = c[rr][u];
                    }
            }
            if (!handed_over) named_bar_sync(1, NTH);  // took no warp-block this time: keep the barrier count whole
            ICNV_STAMP(2);
            __syncthreads();  // gathers done: raw row is dead, partials visible
            ICNV_STAMP(3);
                if (tid == ISSUER) {
                *next_wb = 0;  // used again two rows from now
                if (tma && row + (int64_t)gridDim.x * ROWS < p.n_rows) {
                    issue_row(row + (int64_t)gridDim.x * ROWS);
                    if (p.l2_prefetch && row + 2 * (int64_t)gridDim.x * ROWS < p.n_rows)
                        prefetch_row(row + 2 * (int64_t)gridDim.x * ROWS);
                }
            }
            ICNV_STAMP(13);
        }

        if (!in_group) {
            continue;  // run ahead into the next iteration's gathers (hand-over on barrier 1)
        }

#pragma unroll 1
        for (int rr = my_rr; rr < ROWS; rr += (split_rows ? ROWS : 1)) {
        const bool last_rr = split_rows || rr == ROWS - 1;
        const bool row_exists = ROWS == 1 || row + rr < p.n_rows;  // odd tail: the pair's second row does not exist
        const int pbuf3 = DBUF ? (it & 1) * ICNV_ABS : 0;
        const double2* ABr = AB + (ROWS > 1 ? rr * ICNV_ABS : pbuf3);
        const double* Cpr = Cp + (CGLOBAL ? ((it & 1) * ROWS + rr) * ICNV_ABS : (ROWS > 1 ? rr * ICNV_ABS : pbuf3));
        // ======================= windows =======================
        double v[VPT];
        int nv[TPT];
#pragma unroll
        for (int tt = 0; tt < TPT; ++tt) {
            nv[tt] = 0;
#pragma unroll
            for (int i = 0; i < LOUT; ++i) v[tt * LOUT + i] = INFINITY;
            const int ti = tix + tt * NTH;
            if (ti < p.n_tasks && row_exists) {
                const int4 t = task[tt];
                if ((t.w & 0xFF) == 0) {
                    nv[tt] = t.z;
                    if constexpr (TIER == 0) {
                        double acc[LOUT];
#pragma unroll
                        for (int i = 0; i < LOUT; ++i) {
                            // the peak group's non-linear part C' (its linear part c0 * A rides in the loop below);
                            // from the L2 scratch these loads are issued first and consumed last
                            acc[i] = 0.0;
                        }
                        // the peak group's non-linear part C' (its linear part c0 * A rides in the loop below): from the
                        // L2 scratch these loads are issued first and consumed after the loop
                        double cpk[M3_C ? LOUT : 1];
                        if constexpr (M3_C) {
#pragma unroll
                            for (int i = 0; i < LOUT; ++i) cpk[i] = CGLOBAL ? __ldcg(Cpr + t.x + QSTAR_C + i) : Cpr[t.x + QSTAR_C + i];
                        }
                        const double2* P = ABr + t.x;
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
                        if constexpr (M3_C) {
#pragma unroll
                            for (int i = 0; i < LOUT; ++i) acc[i] += cpk[i];
                        }
#pragma unroll
                        for (int i = 0; i < LOUT; ++i)
                            if (i < t.z) v[tt * LOUT + i] = acc[i] * p.inv_sumw;
                    } else if constexpr (TIER == 1) {
#pragma unroll
                        for (int i = 0; i < LOUT; ++i) {
                            if (i < t.z) {
                                double acc = 0.0;
                                const double2* P = ABr + t.x + i;
                                for (int q = 0; q < NQ; ++q) {
                                    const double2 ab = P[q];
                                    acc = fma(w_alpha[q], ab.x, acc);
                                    acc = fma(w_beta[q], ab.y, acc);
                                }
                                if (qstar >= 0) acc += Cpr[t.x + i + qstar];
                                v[tt * LOUT + i] = acc * p.inv_sumw;
                            }
                        }
                    }
                } else {
                    // chromosome not longer than the window: one flat mean (_infercnv.py:227-236)
                    nv[tt] = 1;
                    double acc = 0.0;
                    for (int g = 0; g < t.z; ++g) acc += ABr[t.x + g].x;
                    v[tt * LOUT] = acc * p.flat_inv[t.w >> 8];
                }
            }
        }

        ICNV_STAMP(15);
        {
            if (last_rr && !DBUF) {  // all partials this thread needs are in registers: hand the buffer over
                named_bar_arrive(1, NTH);
                named_bar_sync(2, n_group);
            }
        }
        if (!row_exists) continue;
        // ======================= write the smoothed row (tile order, fp64) =======================
#pragma unroll
        for (int tt = 0; tt < TPT; ++tt) {
            // the 32 lanes of a warp write 32 consecutive values per instruction (256-byte lines)
            const int ti = tix + tt * NTH;
            if (ti < ((p.n_tasks + 31) & ~31)) {
                // tile moments behind the values: they steer the median bracket of center_rows (fp32 is plenty)
                double s1 = 0.0, s2 = 0.0;
#pragma unroll
                for (int i = 0; i < LOUT; ++i)
                    if (i < nv[tt]) {
                        s1 += v[tt * LOUT + i];
                        s2 = fma(v[tt * LOUT + i], v[tt * LOUT + i], s2);
                    }
                float f1 = (float)s1, f2 = (float)s2;
#pragma unroll
                for (int sh = 16; sh > 0; sh >>= 1) {
                    f1 += __shfl_xor_sync(0xffffffffu, f1, sh);
                    f2 += __shfl_xor_sync(0xffffffffu, f2, sh);
                }
                double* orow = reinterpret_cast<double*>(p.out) + (size_t)(row + rr) * p.ldo;
                if (lane == 0)
                    reinterpret_cast<float2*>(orow + (size_t)((p.n_tasks + 31) >> 5) * (32 * LOUT))[ti >> 5] = make_float2(f1, f2);
                double* o = orow + (size_t)(ti >> 5) * (32 * LOUT) + (ti & 31);
#pragma unroll
                for (int i = 0; i < LOUT; ++i) o[i * 32] = v[tt * LOUT + i];
            }
        }
        }  // rr
        ICNV_STAMP(4);
        ICNV_STAMP(7);
    }
#undef ICNV_STAMP
#undef ICNV_ABS
#undef ICNV_ROW_OFF
}

// ------------------------------------------------------------------------------------------------
// instantiation table
template <int TIER, int NWIN, int GS, bool BOUNDED, bool C64, int TPT, int ROWS = 1, bool DBUF = false>
static int launch_one(const SmoothParams& p, int grid, size_t smem, cudaStream_t stream) {
    auto k = smooth_kernel<TIER, NWIN, GS, BOUNDED, C64, TPT, ROWS, DBUF>;
    ICNV_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k<<<grid, smooth_threads(ROWS, DBUF), smem, stream>>>(p);
    ICNV_CUDA(cudaGetLastError());
    return 0;
}
template <int TIER, int NWIN, int GS, bool BOUNDED, bool C64, int TPT, int ROWS = 1, bool DBUF = false>
static int occ_one(size_t smem, int* out) {
    auto k = smooth_kernel<TIER, NWIN, GS, BOUNDED, C64, TPT, ROWS, DBUF>;
    ICNV_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    ICNV_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(out, k, smooth_threads(ROWS, DBUF), smem));
    return 0;
}

#define ICNV_DISPATCH(FN, ...)                                                                             \
    do {                                                                                                   \
        if (tier == 0 && nwin == 100 && gs == 10 && tpt == 1 && rows == 2) {                               \
            return bounded ? FN<0, 100, 10, true, false, 1, 2>(__VA_ARGS__) : FN<0, 100, 10, false, false, 1, 2>(__VA_ARGS__); \
        }                                                                                                  \
        if (tier == 0 && nwin == 250 && gs == 10 && tpt == 1 && rows == 2) {                               \
            return bounded ? FN<0, 250, 10, true, false, 1, 2>(__VA_ARGS__) : FN<0, 250, 10, false, false, 1, 2>(__VA_ARGS__); \
        }                                                                                                  \
        if (rows != 1) break;                                                                              \
        if (tier == 0 && nwin == 100 && gs == 10 && tpt == 1 && dbuf) {                                    \
            return bounded ? FN<0, 100, 10, true, false, 1, 1, true>(__VA_ARGS__) : FN<0, 100, 10, false, false, 1, 1, true>(__VA_ARGS__); \
        }                                                                                                  \
        if (tier == 0 && nwin == 250 && gs == 10 && tpt == 1 && dbuf) {                                    \
            return bounded ? FN<0, 250, 10, true, false, 1, 1, true>(__VA_ARGS__) : FN<0, 250, 10, false, false, 1, 1, true>(__VA_ARGS__); \
        }                                                                                                  \
        if (dbuf) break;                                                                                   \
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
    } while (0)

int smooth_launch(int tier, int nwin, int gs, bool bounded, bool c64, int tpt, int rows, bool dbuf, const SmoothParams& p, int grid,
                  size_t smem, cudaStream_t stream) {
    ICNV_DISPATCH(launch_one, p, grid, smem, stream);
    set_error("smooth_launch: no kernel instantiation for this configuration");
    return -3;
}
int smooth_occupancy(int tier, int nwin, int gs, bool bounded, bool c64, int tpt, int rows, bool dbuf, size_t smem, int* ctas_per_sm) {
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
