// Exact euclidean kNN on the PCA coordinates as a tensor-core distance GEMM (tcgen05 + TMEM + TMA bulk copies).
// Replaces the neighbour search behind /root/reference/src/infercnvpy/pp/__init__.py:43 (scanpy.pp.neighbors ->
// sklearn brute force below 4096 cells, pynndescent above): north_star's "tensor cores ... if the neighbors step is cast
// as a dense cells x cells GEMM".
//
//   d2(q, c) = |q|^2 + |c|^2 - 2 q.c      the cells x cells matrix of q.c is the GEMM; |q|^2 does not change a query's
//                                         ranking, so candidates are ranked by key = |c|^2 - 2 q.c
//
// Three kernels:
//   knn_pack_kernel    fp32 [n, d] -> per tile of 128 points a 64 KB block that IS the shared-memory image the MMA wants:
//                      [hi | lo][k chunk of 4 floats (16 of them, d padded to 64)][128 rows][16 B] — the canonical K-major
//                      no-swizzle UMMA layout (core matrix = 8 rows x 16 B contiguous; SBO = 128 B between 8-row groups,
//                      LBO = 2048 B between k chunks).  hi = x with the 13 low mantissa bits cleared (exact TF32),
//                      lo = x - hi (exact in fp32; the tensor core reads its top 19 bits).  Also |x|^2 per point
//                      (+inf for the padding rows, which therefore never rank).
//   knn_mma_kernel     one CTA per tile of 128 QUERIES, looping over all candidate tiles.  Warp 0: one thread streams the
//                      candidate blocks with cp.async.bulk (2 stages x 64 KB, mbarrier full/empty).  Warp 1: one thread
//                      issues tcgen05.mma.kind::tf32 (M = 128, N = 128, K = 8): q.c ~= qh.ch + qh.cl + ql.ch ("3xTF32",
//                      ~2^-21 relative: fp32-class) = 3 x n_ksteps MMAs per tile into one of two 128-column TMEM
//                      accumulators, then tcgen05.commit -> the stage is free again and the accumulator is full.
//                      Warps 2..5: thread = query row = TMEM lane; tcgen05.ld 32 columns at a time, key = |c|^2 - 2 dot,
//                      compare with the row's current KK-th best, rare insertion into the row's candidate list in shared
//                      memory ([slot][thread]: conflict-free).
//   knn_rerank_kernel  the KK = k + slack candidates of every query are re-ranked with exact distances
//                      (sum (q - c)^2 accumulated in fp64 from the fp32 coordinates), ties by index -> top k.
//
// The GEMM is 2 n^2 d flops (1e14 at 1M cells x 50), but the epilogue looks at n^2 keys: with ~4 instructions per key on
// 4 warps it takes about as long as the 3 x 7 MMAs of a tile, which is why the accumulator is double-buffered.
#include <cfloat>

#include "icnv_common.cuh"

namespace icnv {

constexpr int KNN_TILE = 128;           // points per tile (M and N of the MMA)
constexpr int KNN_DPAD = 64;            // coordinates padded to 64 floats
constexpr int KNN_CHUNKS = KNN_DPAD / 4;  // 16-byte k chunks
constexpr int KNN_HALF_BYTES = KNN_CHUNKS * KNN_TILE * 16;  // 32 KB: hi or lo of one tile
constexpr int KNN_BLOCK_BYTES = 2 * KNN_HALF_BYTES;         // 64 KB per tile
constexpr int KNN_KK = 24;              // candidates kept per query before the exact re-rank (k <= 16 -> 8 of slack)
constexpr int KNN_STAGES = 2;
constexpr int KNN_THREADS = 192;
constexpr int KNN_STATE_WORDS = 2 * KNN_KK * KNN_TILE + KNN_TILE;  // per query tile: keys, ids, meta (32-bit words)
// candidate tiles per launch: 1024 x 64 KB = 64 MB stays in the 126 MB L2 while every query tile streams it (one launch
// over 1M points re-read 950 GB from HBM at a 65 % L2 hit rate: ncu, profiles/ncu_knn_1m_r2.csv)
constexpr int64_t KNN_SUPER_TILES = 1024;

// ------------------------------------------------------------------------------------------------ pack
__global__ void __launch_bounds__(256) knn_pack_kernel(const float* __restrict__ P, int64_t n, int d, int64_t ld, int64_t n_pad,
                                                       float* __restrict__ packed, float* __restrict__ norms /* nullable */) {
    // thread = (point, k chunk): 16 threads per point read 4 consecutive floats each
    const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t pt = gid / KNN_CHUNKS;
    const int ch = (int)(gid % KNN_CHUNKS);
    if (pt >= n_pad) return;
    float x[4] = {0.f, 0.f, 0.f, 0.f};
    if (pt < n) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int c = ch * 4 + u;
            if (c < d) x[u] = __ldg(P + pt * ld + c);
        }
    }
    float hi[4], lo[4];
    double s = 0.0;
#pragma unroll
    for (int u = 0; u < 4; ++u) {
        hi[u] = __uint_as_float(__float_as_uint(x[u]) & 0xFFFFE000u);
        lo[u] = x[u] - hi[u];
        s += (double)x[u] * (double)x[u];
    }
    const int64_t tile = pt / KNN_TILE;
    const int row = (int)(pt % KNN_TILE);
    float* blk = packed + tile * (KNN_BLOCK_BYTES / 4);
    const size_t off = ((size_t)ch * KNN_TILE + row) * 4;  // floats: chunk-major, then row, 16 B per (chunk, row)
    *reinterpret_cast<float4*>(blk + off) = make_float4(hi[0], hi[1], hi[2], hi[3]);
    *reinterpret_cast<float4*>(blk + KNN_HALF_BYTES / 4 + off) = make_float4(lo[0], lo[1], lo[2], lo[3]);
    // |x|^2: sum over the 16 chunk threads of a point (they are consecutive lanes of one half-warp)
#pragma unroll
    for (int sh = 8; sh > 0; sh >>= 1) s += __shfl_xor_sync(0xffffffffu, s, sh);
    if (ch == 0 && norms != nullptr) norms[pt] = pt < n ? (float)s : INFINITY;
}

// ------------------------------------------------------------------------------------------------ PTX helpers
__device__ __forceinline__ void mbar_arrive(unsigned long long* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(unsigned long long* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// K-major, no swizzle: start address, LBO (between k chunks), SBO (between 8-row groups), all >> 4; version 1 (sm_100)
__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFFu);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
    d |= (uint64_t)1 << 46;  // descriptor version for Blackwell
    return d;                // base offset 0, lbo mode 0, layout type 0 (SWIZZLE_NONE)
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, "
        "%19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
          "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
          "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// Rare path of the epilogue, kept out of line: 128 inlined copies per tile made the kernel 4800 instructions long and the
// epilogue warps stalled on instruction fetch ("no_inst" on every reconvergence point, ncu r2).  Everything it changes
// lives in shared memory and the new bar comes back BY VALUE: reference parameters put `thr` on the stack and every
// compare of the hot loop then waited for a local-memory load.
//   lmeta[row] = entries in the list (<= KK) | position of the worst entry << 8
__device__ __noinline__ float knn_insert(float* lkey, int32_t* lidx, int32_t* lmeta, int row, float key, int32_t id) {
    const int meta = lmeta[row];
    int cnt = meta & 0xFF;
    const int slot = cnt < KNN_KK ? cnt : (meta >> 8);
    lkey[slot * KNN_TILE + row] = key;
    lidx[slot * KNN_TILE + row] = id;
    if (cnt < KNN_KK) ++cnt;
    float thr = INFINITY;
    int maxpos = 0;
    if (cnt == KNN_KK) {  // (re)locate the worst entry: it is the bar the next candidates must pass
        float v[KNN_KK];
#pragma unroll
        for (int i = 0; i < KNN_KK; ++i) v[i] = lkey[i * KNN_TILE + row];
        float m = v[0];
#pragma unroll
        for (int i = 1; i < KNN_KK; ++i)
            if (v[i] > m) {
                m = v[i];
                maxpos = i;
            }
        thr = m;
    }
    lmeta[row] = cnt | (maxpos << 8);
    return thr;
}

struct __align__(16) KnnSmem {
    unsigned long long full[KNN_STAGES], empty[KNN_STAGES], tmem_full[2], tmem_empty[2], a_full;
    uint32_t tmem_base;
    uint32_t pad[3];
};

// ------------------------------------------------------------------------------------------------ main kernel
__global__ void __launch_bounds__(KNN_THREADS, 1) knn_mma_kernel(const float* __restrict__ packed, const float* __restrict__ packed_q,
                                                                 const float* __restrict__ norms, int64_t n_cand_tiles, int n_ksteps,
                                                                 int32_t* __restrict__ cand_idx /* [n_q_tiles*128, KK] */,
                                                                 int64_t t_off, float* __restrict__ state, int resume) {
    // Candidates are streamed in L2-sized super-blocks (knn_launch): `packed` / `norms` point at candidate tile t_off, this
    // launch covers n_cand_tiles of them, and a query tile's candidate lists travel between launches through `state`
    // ([q_tile][keys KK*128 | ids KK*128 | meta 128]; resume != 0: start from it).
    extern __shared__ __align__(1024) unsigned char smem[];
    // layout: [A hi|lo 64 KB][B stage 0 64 KB][B stage 1 64 KB][lists key KK*128*4][lists idx KK*128*4][norms 4 x 128][meta 128][KnnSmem]
    unsigned char* sA = smem;
    unsigned char* sB = smem + KNN_BLOCK_BYTES;
    float* lkey = reinterpret_cast<float*>(smem + (size_t)(1 + KNN_STAGES) * KNN_BLOCK_BYTES);
    int32_t* lidx = reinterpret_cast<int32_t*>(lkey + KNN_KK * KNN_TILE);
    // |c|^2 of the candidate tiles in flight: a ring of 4 (2 stages + 2 accumulators: tile t + 2 is only loaded after the
    // MMAs of tile t, which wait for the epilogue of tile t - 2, so slots t - 1, t, t + 1 are never overwritten)
    float* cnorm = reinterpret_cast<float*>(lidx + KNN_KK * KNN_TILE);
    int32_t* lmeta = reinterpret_cast<int32_t*>(cnorm + 4 * KNN_TILE);
    KnnSmem* S = reinterpret_cast<KnnSmem*>(lmeta + KNN_TILE);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t q_tile = blockIdx.x;

    if (threadIdx.x == 0) {
        for (int s = 0; s < KNN_STAGES; ++s) {
            mbar_init(&S->full[s], 1);
            mbar_init(&S->empty[s], 1);
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(&S->tmem_full[b], 1);
            mbar_init(&S->tmem_empty[b], 128);
        }
        mbar_init(&S->a_full, 1);
        mbar_fence_init();
    }
    if (warp == 1) {  // TMEM: 2 accumulators of 128 fp32 columns
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&S->tmem_base)), "r"(256u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = S->tmem_base;

    if (warp == 0) {
        // ===================== producer: one thread streams the query block once, then every candidate block =====================
        if (lane == 0) {
            constexpr uint32_t CH = 16384;
            mbar_expect_tx(&S->a_full, KNN_BLOCK_BYTES);
            const char* srcA = reinterpret_cast<const char*>(packed_q) + (size_t)q_tile * KNN_BLOCK_BYTES;
            for (uint32_t off = 0; off < KNN_BLOCK_BYTES; off += CH) bulk_g2s_plain(sA + off, srcA + off, CH, &S->a_full);
            for (int64_t t = 0; t < n_cand_tiles; ++t) {
                const int s = (int)(t % KNN_STAGES);
                const uint32_t use = (uint32_t)(t / KNN_STAGES);
                if (use > 0) mbar_wait(&S->empty[s], (use - 1) & 1u);
                mbar_expect_tx(&S->full[s], KNN_BLOCK_BYTES + KNN_TILE * 4);
                bulk_g2s_plain(cnorm + (t & 3) * KNN_TILE, norms + t * KNN_TILE, KNN_TILE * 4, &S->full[s]);
                const char* src = reinterpret_cast<const char*>(packed) + (size_t)t * KNN_BLOCK_BYTES;
                unsigned char* dst = sB + (size_t)s * KNN_BLOCK_BYTES;
                for (uint32_t off = 0; off < KNN_BLOCK_BYTES; off += CH) bulk_g2s_plain(dst + off, src + off, CH, &S->full[s]);
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer: one thread =====================
        if (lane == 0) {
            // instruction descriptor: D = F32 (bits 4..5 = 1), A = B = TF32 (bits 7..9 / 10..12 = 2), K-major both,
            // N = 128 (bits 17..22 = N >> 3), M = 128 (bits 24..28 = M >> 4)
            const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(KNN_TILE >> 3) << 17) | ((uint32_t)(KNN_TILE >> 4) << 24);
            constexpr uint32_t LBO = KNN_TILE * 16, SBO = 128;
            const uint32_t a_hi = smem_u32(sA), a_lo = a_hi + KNN_HALF_BYTES;
            mbar_wait(&S->a_full, 0);
            for (int64_t t = 0; t < n_cand_tiles; ++t) {
                const int s = (int)(t % KNN_STAGES);
                const int b = (int)(t & 1);
                const uint32_t use_s = (uint32_t)(t / KNN_STAGES), use_b = (uint32_t)(t >> 1);
                mbar_wait(&S->full[s], use_s & 1u);
                if (use_b > 0) mbar_wait(&S->tmem_empty[b], (use_b - 1) & 1u);
                tc_fence_after();
                const uint32_t b_hi = smem_u32(sB + (size_t)s * KNN_BLOCK_BYTES), b_lo = b_hi + KNN_HALF_BYTES;
                const uint32_t tmem_d = tmem_base + (uint32_t)b * KNN_TILE;
                uint32_t acc = 0;
                for (int pass = 0; pass < 3; ++pass) {  // qh.ch, qh.cl, ql.ch
                    const uint32_t abase = pass == 2 ? a_lo : a_hi;
                    const uint32_t bbase = pass == 1 ? b_lo : b_hi;
                    for (int ks = 0; ks < n_ksteps; ++ks) {  // K = 8 floats = 2 chunks per instruction
                        umma_tf32(tmem_d, umma_desc(abase + (uint32_t)ks * 2 * LBO, LBO, SBO), umma_desc(bbase + (uint32_t)ks * 2 * LBO, LBO, SBO),
                                  idesc, acc);
                        acc = 1;
                    }
                }
                tc_commit(&S->empty[s]);      // the stage may be refilled once these MMAs have read it
                tc_commit(&S->tmem_full[b]);  // ... and the accumulator is complete
            }
        }
    } else {
        // ===================== epilogue: thread = query row = TMEM lane =====================
        const int quarter = warp & 3;             // the TMEM lanes this warp may touch: 32 * (warp % 4) ..
        const int row = quarter * 32 + lane;      // query row inside the tile
        float thr = INFINITY;                     // current worst key of a full list
        float* st_key = state ? state + (size_t)q_tile * KNN_STATE_WORDS : nullptr;
        int32_t* st_idx = reinterpret_cast<int32_t*>(st_key + KNN_KK * KNN_TILE);
        int32_t* st_meta = st_idx + KNN_KK * KNN_TILE;
        if (resume) {
            const int meta = st_meta[row];
            lmeta[row] = meta;
            for (int i = 0; i < KNN_KK; ++i) {
                lkey[i * KNN_TILE + row] = st_key[i * KNN_TILE + row];
                lidx[i * KNN_TILE + row] = st_idx[i * KNN_TILE + row];
            }
            if ((meta & 0xFF) == KNN_KK) thr = lkey[(meta >> 8) * KNN_TILE + row];
        } else {
            lmeta[row] = 0;
            for (int i = 0; i < KNN_KK; ++i) lidx[i * KNN_TILE + row] = -1;
        }
        for (int64_t t = 0; t < n_cand_tiles; ++t) {
            const int b = (int)(t & 1);
            mbar_wait(&S->tmem_full[b], (uint32_t)(t >> 1) & 1u);
            tc_fence_after();
            const float* cn = cnorm + (t & 3) * KNN_TILE;  // landed before the MMAs that filled this accumulator were issued
#pragma unroll 1
            for (int c0 = 0; c0 < KNN_TILE; c0 += 32) {
                uint32_t r[32];
                tmem_ld32(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(b * KNN_TILE + c0), r);
                // 32 independent keys, one min tree, ONE branch per chunk: the single epilogue warp of a scheduler has
                // nobody to hide latencies behind, so a compare-and-branch per key (dependent FFMA -> FSETP -> BRA) cost
                // ~35 cycles per key (ncu r2); the per-key tests only run for the chunks that hold a candidate
                float key[32];
#pragma unroll
                for (int j4 = 0; j4 < 32; j4 += 4) {
                    const float4 n4 = *reinterpret_cast<const float4*>(cn + c0 + j4);
                    key[j4 + 0] = fmaf(-2.f, __uint_as_float(r[j4 + 0]), n4.x);
                    key[j4 + 1] = fmaf(-2.f, __uint_as_float(r[j4 + 1]), n4.y);
                    key[j4 + 2] = fmaf(-2.f, __uint_as_float(r[j4 + 2]), n4.z);
                    key[j4 + 3] = fmaf(-2.f, __uint_as_float(r[j4 + 3]), n4.w);
                }
                float mn[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) mn[j] = fminf(key[j], key[j + 16]);
#pragma unroll
                for (int w = 8; w > 0; w >>= 1)
#pragma unroll
                    for (int j = 0; j < w; ++j) mn[j] = fminf(mn[j], mn[j + w]);
                if (mn[0] < thr) {
#pragma unroll
                    for (int j = 0; j < 32; ++j)
                        if (key[j] < thr) thr = knn_insert(lkey, lidx, lmeta, row, key[j], (int32_t)((t + t_off) * KNN_TILE + c0 + j));
                }
            }
            tc_fence_before();
            mbar_arrive(&S->tmem_empty[b]);
        }
        int32_t* dst = cand_idx + ((int64_t)blockIdx.x * KNN_TILE + row) * KNN_KK;
        for (int i = 0; i < KNN_KK; ++i) dst[i] = lidx[i * KNN_TILE + row];
        if (st_key) {  // hand the lists to the launch that covers the next candidate super-block
            st_meta[row] = lmeta[row];
            for (int i = 0; i < KNN_KK; ++i) {
                st_key[i * KNN_TILE + row] = lkey[i * KNN_TILE + row];
                st_idx[i * KNN_TILE + row] = lidx[i * KNN_TILE + row];
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(256u) : "memory");
    }
}

// ------------------------------------------------------------------------------------------------ exact re-rank
// One warp per query: exact squared distances to its KK candidates (fp64 accumulation of fp32 differences), then the
// k smallest by (distance, index) are written in order.
__global__ void __launch_bounds__(256) knn_rerank_kernel(const float* __restrict__ P, int64_t n_all, int d, int64_t ld, int64_t q0, int64_t nq,
                                                         const int32_t* __restrict__ cand_idx, int k, int out_ld,
                                                         int32_t* __restrict__ knn_idx, float* __restrict__ knn_d2) {
    const int lane = threadIdx.x & 31;
    const int64_t q = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (q >= nq) return;
    const float* xq = P + (q0 + q) * ld;
    int32_t id = lane < KNN_KK ? cand_idx[q * KNN_KK + lane] : -1;
    if (id >= n_all) id = -1;
    double d2 = INFINITY;
    if (id >= 0) {
        const float* xc = P + (int64_t)id * ld;
        double s = 0.0;
        for (int c = 0; c < d; ++c) {
            const double df = (double)__ldg(xq + c) - (double)__ldg(xc + c);
            s = fma(df, df, s);
        }
        d2 = s;
    }
    // rank of this lane's candidate among the 32 lanes by (d2, id); empty slots sort last
    int rank = 0;
    for (int o = 0; o < 32; ++o) {
        const double od = __shfl_sync(0xffffffffu, d2, o);
        const int32_t oi = __shfl_sync(0xffffffffu, id, o);
        const bool o_valid = oi >= 0, me_valid = id >= 0;
        bool before;
        if (o_valid != me_valid)
            before = o_valid;
        else if (!o_valid)
            before = o < lane;
        else
            before = (od < d2) || (od == d2 && oi < id);
        rank += before ? 1 : 0;
    }
    if (rank < k) {
        knn_idx[q * out_ld + rank] = id;
        knn_d2[q * out_ld + rank] = id >= 0 ? (float)d2 : INFINITY;
    }
}

// ------------------------------------------------------------------------------------------------ launchers
size_t knn_smem_bytes() { return (size_t)(1 + KNN_STAGES) * KNN_BLOCK_BYTES + (size_t)2 * KNN_KK * KNN_TILE * 4 + 5 * KNN_TILE * 4 + sizeof(KnnSmem) + 1024; }

size_t knn_workspace_bytes(int64_t n_all, int64_t nq) {
    const int64_t n_tiles = (n_all + KNN_TILE - 1) / KNN_TILE;
    const int64_t q_tiles = (nq + KNN_TILE - 1) / KNN_TILE;
    return (size_t)(n_tiles + q_tiles) * KNN_BLOCK_BYTES + (size_t)n_tiles * KNN_TILE * 4 + (size_t)q_tiles * KNN_TILE * KNN_KK * 4 +
           (n_tiles > KNN_SUPER_TILES + KNN_SUPER_TILES / 2 ? (size_t)q_tiles * KNN_STATE_WORDS * 4 : 0) + 4096;
}

int knn_launch(const float* P, int64_t n_all, int d, int64_t ld, int64_t q0, int64_t nq, int k, int out_ld, int32_t* knn_idx, float* knn_d2,
               void* workspace, cudaStream_t st) {
    if (d > KNN_DPAD) {
        set_error("icnv_knn_f32: at most 64 dimensions");
        return -3;
    }
    if (k > KNN_KK - 4) {
        set_error("icnv_knn_f32: k too large for the candidate lists (k <= 20)");
        return -3;
    }
    if (nq == 0) return 0;
    const int64_t n_tiles = (n_all + KNN_TILE - 1) / KNN_TILE;
    const int64_t n_pad = n_tiles * KNN_TILE;
    char* ws = reinterpret_cast<char*>((reinterpret_cast<uintptr_t>(workspace) + 1023) & ~(uintptr_t)1023);
    const int64_t q_tiles = (nq + KNN_TILE - 1) / KNN_TILE;
    float* packed = reinterpret_cast<float*>(ws);                                        // candidates: every point
    float* packed_q = reinterpret_cast<float*>(ws + (size_t)n_tiles * KNN_BLOCK_BYTES);  // queries: rows q0 .. q0 + nq
    float* norms = reinterpret_cast<float*>(ws + (size_t)(n_tiles + q_tiles) * KNN_BLOCK_BYTES);
    int32_t* cand = reinterpret_cast<int32_t*>(reinterpret_cast<char*>(norms) + (size_t)n_tiles * KNN_TILE * 4);
    {
        const int64_t threads = n_pad * KNN_CHUNKS;
        knn_pack_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, st>>>(P, n_all, d, ld, n_pad, packed, norms);
        ICNV_CUDA(cudaGetLastError());
        const int64_t threads_q = q_tiles * KNN_TILE * KNN_CHUNKS;
        knn_pack_kernel<<<(unsigned)((threads_q + 255) / 256), 256, 0, st>>>(P + q0 * ld, nq, d, ld, q_tiles * KNN_TILE, packed_q, nullptr);
        ICNV_CUDA(cudaGetLastError());
    }
    const size_t smem = knn_smem_bytes();
    ICNV_CUDA(cudaFuncSetAttribute(knn_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (n_tiles <= KNN_SUPER_TILES + KNN_SUPER_TILES / 2) {
        knn_mma_kernel<<<(unsigned)q_tiles, KNN_THREADS, smem, st>>>(packed, packed_q, norms, n_tiles, (d + 7) / 8, cand, 0, nullptr, 0);
    } else {
        float* state = reinterpret_cast<float*>(reinterpret_cast<char*>(cand) + (size_t)q_tiles * KNN_TILE * KNN_KK * 4);
        for (int64_t t0 = 0; t0 < n_tiles; t0 += KNN_SUPER_TILES) {
            int64_t t1 = std::min(n_tiles, t0 + KNN_SUPER_TILES);
            if (n_tiles - t1 < KNN_SUPER_TILES / 2) t1 = n_tiles;  // no short tail launch
            knn_mma_kernel<<<(unsigned)q_tiles, KNN_THREADS, smem, st>>>(packed + (size_t)t0 * (KNN_BLOCK_BYTES / 4), packed_q,
                                                                         norms + t0 * KNN_TILE, t1 - t0, (d + 7) / 8, cand, t0, state, t0 > 0);
            if (t1 == n_tiles) break;
        }
    }
    ICNV_CUDA(cudaGetLastError());
    knn_rerank_kernel<<<(unsigned)((nq * 32 + 255) / 256), 256, 0, st>>>(P, n_all, d, ld, q0, nq, cand, k, out_ld, knn_idx, knn_d2);
    ICNV_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace icnv
