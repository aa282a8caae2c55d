// EXPERIMENTAL, OFF BY DEFAULT (env ICNV_SMOOTH_BANDS=2 at plan creation) — written after the last GPU session of
// round 1 and NOT YET RUN on hardware; the product path is icnv_smooth.cu.  DESIGN.md §7.1 has the reasoning.
//
// Row-pair smoothing kernel (window 100, step 10, permuted walk) with the group axis cut into TWO BANDS so that the
// sliding windows of one band (phase 3) always overlap the gathers of the other (phase 2):
//   * the task list is cut at a tile boundary; band X's groups are every group its tasks read (groups read by both
//     bands are stored twice), so the two ranges of the partial-sum array are disjoint and nobody ever waits for a
//     hand-over of the buffer;
//   * warps 0-7 own phase 3 of band A (4 per staged row), warps 8-15 of band B; every warp gathers work units of
//     both bands, drawn from one shared-memory counter per band.
// Per pair of rows:  wait TMA -> band-A units -> [A-owners: windows of band A] | [others: band-B units] -> barrier 0
// (staged rows dead, next TMA) -> [B-owners: windows of band B] | [others: next pair].
//   barrier 5  every warp has stored its band-A units          (A-owners sync, B-owners arrive)
//   barrier 1  A-owners have read range A                       (A-owners arrive; B-owners sync before their first
//                                                                store into range A of the NEXT pair)
//   barrier 2  the same among the A-owners
//   barrier 3 / 4  likewise for band B (B-owners arrive / sync among themselves; A-owners sync on 3)
//   barrier 0  all units of the pair stored (__syncthreads)
// Arithmetic is identical to smooth_kernel<0, 100, 10, *, false, 1, 2>: same partial sums, same window formula, same
// tile-order output, so every later stage and every parity test applies unchanged.
#include "icnv_common.cuh"

namespace icnv {

namespace {

struct __align__(16) BandScratch {  // same size as icnv_smooth.cu's Scratch: the baked shared-window base stays valid
    unsigned long long mbar;
    int next_unit[2][2];  // [row-pair parity][band]
};
static_assert(sizeof(BandScratch) == 32, "scratch block must keep its size");

__device__ __forceinline__ float bd_lds_f32(uint32_t addr) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
    return v;
}

constexpr int B_NWIN = 100, B_GS = 10, B_NQ = B_NWIN / B_GS, B_ROWS = 2, B_UW = 4;

}  // namespace

template <bool BOUNDED>
__global__ void __launch_bounds__(NT, 1) smooth_banded_kernel(const SmoothParams p) {
    extern __shared__ __align__(16) unsigned char smem[];
    BandScratch* sc = reinterpret_cast<BandScratch*>(smem);
    float* raw = reinterpret_cast<float*>(smem + sizeof(BandScratch));
    const int ABS = p.NGpad + PAD_GROUPS;  // partial-sum slots per staged row (both bands + the gap between them)
    double2* AB = reinterpret_cast<double2*>(reinterpret_cast<unsigned char*>(raw) + (size_t)B_ROWS * p.Gpad * 4);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int ISSUER = NT - 32;      // lane 0 of warp 15 (a band-B owner) drives the TMA
    const int band = warp >> 3;          // whose phase 3 this warp owns
    const int my_rr = (warp >> 2) & 1;   // ... for which staged row
    const int my_tile = p.band_tile0[band] + (warp & 3);
    const bool has_tile = my_tile < p.band_tile0[band] + p.band_tiles[band];
    const int ti = my_tile * 32 + lane;
    const int4 task = (has_tile && ti < p.n_tasks) ? __ldg(reinterpret_cast<const int4*>(p.tasks) + ti) : make_int4(0, 0, 0, 1 << 16);
    const bool has_task = has_tile && ti < p.n_tasks;
    const int n_tiles = (p.n_tasks + 31) >> 5;

    // ---- one-time setup: zero pads and the WHOLE partial-sum array (the gap between the bands is never written)
    for (int rr = 0; rr < B_ROWS; ++rr) {
        for (int i = p.G + tid; i < p.Gpad; i += NT) raw[rr * p.Gpad + i] = 0.f;
        for (int i = tid; i < ABS; i += NT) AB[rr * ABS + i] = make_double2(0.0, 0.0);
    }
    if (tid == 0) {
        mbar_init(&sc->mbar, 1);
        mbar_fence_init();
        sc->next_unit[0][0] = sc->next_unit[0][1] = sc->next_unit[1][0] = sc->next_unit[1][1] = 0;
        if ((smem_u32(raw) & 0xFFFFFFu) != p.raw_base) __trap();  // host baked a different base
    }
    __syncthreads();

    const uint64_t pol = l2_policy_evict_first();
    uint32_t parity = 0;
    const uint32_t row_bytes = (uint32_t)p.G * 4u;
    const uint32_t row_off = (uint32_t)p.Gpad * 4u;
    auto issue_rows = [&](int64_t r) {
        const int nvalid = (int)min((int64_t)B_ROWS, p.n_rows - r);
        mbar_expect_tx(&sc->mbar, row_bytes * (uint32_t)nvalid);
        constexpr uint32_t CH = 16384;
        for (int rr = 0; rr < nvalid; ++rr) {
            const char* src = reinterpret_cast<const char*>(p.X + (r + rr) * p.ldx);
            char* dst = reinterpret_cast<char*>(raw) + (size_t)rr * row_off;
            for (uint32_t off = 0; off < row_bytes; off += CH) bulk_g2s(dst + off, src + off, min(CH, row_bytes - off), &sc->mbar, pol);
        }
    };
    auto prefetch_rows = [&](int64_t r) {
        const int nvalid = (int)min((int64_t)B_ROWS, p.n_rows - r);
        constexpr uint32_t CH = 16384;
        for (int rr = 0; rr < nvalid; ++rr) {
            const char* src = reinterpret_cast<const char*>(p.X + (r + rr) * p.ldx);
            for (uint32_t off = 0; off < row_bytes; off += CH) bulk_prefetch_l2(src + off, min(CH, row_bytes - off));
        }
    };
    const int64_t step_rows = (int64_t)gridDim.x * B_ROWS;
    int64_t row = (int64_t)blockIdx.x * B_ROWS;
    if (tid == ISSUER && row < p.n_rows) {
        issue_rows(row);
        if (p.l2_prefetch && row + step_rows < p.n_rows) prefetch_rows(row + step_rows);
    }
    const float clipf = p.clipf;

    // ---- phase 2: the units of one band.  `first_unit` is the band's first unit in the tables; the hand-over barrier
    //      `bar_id` is passed before this warp's first store into the band's range (not at all when `handed`).
    auto gather_band = [&](int* counter, int first_unit, int n_units, int bar_id, bool handed) {
        while (true) {
            int u = 0;
            if (lane == 0) u = atomicAdd(counter, 1);
            u = __shfl_sync(0xffffffffu, u, 0);
            if (u >= n_units) break;
            const int unit = first_unit + u;
            double a[B_ROWS][B_UW], b[B_ROWS][B_UW];
            const size_t tbase = ((size_t)unit * B_GS * 32 + lane) * B_UW;
            const uint32_t* ip = p.off_w + tbase;
            const float* lp = p.lo_w + tbase;
            const float* hp = p.hi_w + tbase;
#pragma unroll
            for (int j = 0; j < B_GS; ++j) {
                const uint4 id = ldg_nc_u4(ip + j * (32 * B_UW));
                const float4 lo = ldg_nc_f4(lp + j * (32 * B_UW));
                float4 hi = lo;
                if constexpr (BOUNDED) hi = ldg_nc_f4(hp + j * (32 * B_UW));
                uint32_t ad[B_UW] = {id.x, id.y, id.z, id.w};
                const float l4[B_UW] = {lo.x, lo.y, lo.z, lo.w};
                const float h4[B_UW] = {hi.x, hi.y, hi.z, hi.w};
                double jd[B_UW];
#pragma unroll
                for (int q = 0; q < B_UW; ++q) {
                    jd[q] = __hiloint2double(0x43300000, (int)(ad[q] >> 24)) - 4503599627370496.0;  // element position
                    ad[q] &= 0x00FFFFFFu;
                }
                float x[B_ROWS][B_UW];
#pragma unroll
                for (int rr = 0; rr < B_ROWS; ++rr)
#pragma unroll
                    for (int q = 0; q < B_UW; ++q) x[rr][q] = bd_lds_f32(ad[q] + (rr ? row_off : 0u));
#pragma unroll
                for (int rr = 0; rr < B_ROWS; ++rr)
#pragma unroll
                    for (int q = 0; q < B_UW; ++q) {
                        const float xv = x[rr][q];
                        float d;
                        if constexpr (BOUNDED)
                            d = xv > h4[q] ? xv - h4[q] : (xv < l4[q] ? xv - l4[q] : 0.f);
                        else
                            d = xv - l4[q];
                        d = fminf(fmaxf(d, -clipf), clipf);
                        const double dd = (double)d;
                        if (j == 0) {
                            a[rr][q] = dd;
                            b[rr][q] = jd[q] * dd;
                        } else {
                            a[rr][q] += dd;
                            b[rr][q] = fma(jd[q], dd, b[rr][q]);
                        }
                    }
            }
            const int4 gid = __ldg(reinterpret_cast<const int4*>(p.grp_w) + ((unit << 5) + lane));
            const int gq[B_UW] = {gid.x, gid.y, gid.z, gid.w};
            if (!handed) {
                named_bar_sync(bar_id, NT);  // the band's owners have read the previous pair's partials
                handed = true;
            }
#pragma unroll
            for (int rr = 0; rr < B_ROWS; ++rr)
#pragma unroll
                for (int q = 0; q < B_UW; ++q) AB[rr * ABS + gq[q]] = make_double2(a[rr][q], b[rr][q]);
        }
        if (!handed) named_bar_sync(bar_id, NT);  // took no unit of this band: keep the barrier count whole
    };

    // ---- phase 3 of this warp's tile (band `band`, staged row `my_rr`): windows from the partials, hand-over, stores
    auto windows_and_store = [&](int64_t row0, int bar_arrive_id, int bar_group_id) {
        const bool row_exists = row0 + my_rr < p.n_rows;  // odd tail: the pair's second row does not exist
        double v[LOUT];
#pragma unroll
        for (int i = 0; i < LOUT; ++i) v[i] = INFINITY;
        int nv = 0;
        if (has_task && row_exists) {
            const double2* P = AB + my_rr * ABS + task.x;
            if ((task.w & 0xFF) == 0) {
                nv = task.z;
                double acc[LOUT];
#pragma unroll
                for (int i = 0; i < LOUT; ++i) acc[i] = 0.0;
#pragma unroll
                for (int q = 0; q < B_NQ + LOUT - 1; ++q) {
                    const double2 ab = P[q];
#pragma unroll
                    for (int i = 0; i < LOUT; ++i) {
                        const int w = q - i;
                        if (w >= 0 && w < B_NQ) {
                            const int al = pyr(B_NWIN, B_GS * w);
                            const int be = pyr(B_NWIN, B_GS * w + 1) - al;
                            acc[i] = fma((double)al, ab.x, acc[i]);
                            if (be == 1)
                                acc[i] += ab.y;
                            else if (be == -1)
                                acc[i] -= ab.y;
                        }
                    }
                }
#pragma unroll
                for (int i = 0; i < LOUT; ++i)
                    if (i < task.z) v[i] = acc[i] * p.inv_sumw;
            } else {
                // chromosome not longer than the window: one flat mean (_infercnv.py:227-236)
                nv = 1;
                double acc = 0.0;
                for (int g = 0; g < task.z; ++g) acc += P[g].x;
                v[0] = acc * p.flat_inv[task.w >> 8];
            }
        }
        named_bar_arrive(bar_arrive_id, NT);  // this thread's partials are in registers
        named_bar_sync(bar_group_id, NT / 2);
        if (!has_tile || !row_exists) return;
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
        double* orow = reinterpret_cast<double*>(p.out) + (size_t)(row0 + my_rr) * p.ldo;
        if (lane == 0) reinterpret_cast<float2*>(orow + (size_t)n_tiles * (32 * LOUT))[my_tile] = make_float2(f1, f2);
        double* o = orow + (size_t)my_tile * (32 * LOUT) + lane;
#pragma unroll
        for (int i = 0; i < LOUT; ++i) o[i * 32] = v[i];
    };

    int it = 0;
    for (; row < p.n_rows; row += step_rows, ++it) {
        mbar_wait(&sc->mbar, parity);
        parity ^= 1u;
        int* cnt = sc->next_unit[it & 1];
        // band A: its owners synchronised among themselves (barrier 2) when they last read range A
        gather_band(cnt + 0, 0, p.band_units[0], 1, it == 0 || band == 0);
        if (band == 0) {
            named_bar_sync(5, NT);  // every warp has left band A: range A is complete
            windows_and_store(row, 1, 2);
        } else {
            named_bar_arrive(5, NT);
        }
        gather_band(cnt + 1, p.band_units[0], p.band_units[1], 3, it == 0 || band == 1);
        __syncthreads();  // every unit of the pair is stored: staged rows are dead, range B is complete
        if (tid == ISSUER) {
            cnt[0] = 0;  // used again two pairs from now
            cnt[1] = 0;
            if (row + step_rows < p.n_rows) {
                issue_rows(row + step_rows);
                if (p.l2_prefetch && row + 2 * step_rows < p.n_rows) prefetch_rows(row + 2 * step_rows);
            }
        }
        if (band == 1) windows_and_store(row, 3, 4);
    }
}

int smooth_banded_launch(bool bounded, const SmoothParams& p, int grid, size_t smem, cudaStream_t stream) {
    if (bounded) {
        ICNV_CUDA(cudaFuncSetAttribute(smooth_banded_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        smooth_banded_kernel<true><<<grid, NT, smem, stream>>>(p);
    } else {
        ICNV_CUDA(cudaFuncSetAttribute(smooth_banded_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        smooth_banded_kernel<false><<<grid, NT, smem, stream>>>(p);
    }
    ICNV_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace icnv
