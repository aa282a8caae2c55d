// Exact np.median of n float64 values by a whole CTA: radix selection on the order-preserving 64-bit pattern
// (8 passes of 8 bits, most significant byte first, shared-memory histogram), then the second middle value for even n
// (np.median returns the mean of the two middle values, /root/reference/src/infercnvpy/tl/_infercnv.py:442,444).
// General-purpose path (any n, values reachable through a functor): the per-gene layer and the wide-K row centring use
// it; the bench-shaped row centring has its own warp-per-row kernel in icnv_aux.cu.
#pragma once
#include <stdint.h>

namespace icnv {

struct SelectSmem {
    int hist[256];
    unsigned long long prefix, min_gt;
    int rank, cnt_le;
};

__device__ __forceinline__ unsigned long long sel_ordered(double v) {
    const unsigned long long b = (unsigned long long)__double_as_longlong(v);
    return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
__device__ __forceinline__ double sel_unordered(unsigned long long k) {
    const unsigned long long b = (k >> 63) ? (k & 0x7FFFFFFFFFFFFFFFull) : ~k;
    return __longlong_as_double((long long)b);
}

// Every thread of the CTA (NTHREADS of them) must call this with the same arguments; get(i) returns value i.  The
// values must be visible to the whole CTA on entry (caller synchronises).  Returns 0 for n == 0.
template <int NTHREADS, typename F>
__device__ double cta_median(F get, int n, SelectSmem& ss) {
    const int tid = threadIdx.x, lane = tid & 31;
    if (n <= 0) return 0.0;
    if (tid == 0) {
        ss.prefix = 0ull;
        ss.rank = (n - 1) >> 1;
    }
    for (int pass = 0; pass < 8; ++pass) {
        const int shift = 56 - 8 * pass;
        if (tid < 256) ss.hist[tid] = 0;
        __syncthreads();
        const unsigned long long prefix = ss.prefix;
        for (int i = tid; i < n; i += NTHREADS) {
            const unsigned long long key = sel_ordered(get(i));
            if (pass == 0 || (key >> (shift + 8)) == (prefix >> (shift + 8))) atomicAdd(&ss.hist[(int)((key >> shift) & 255ull)], 1);
        }
        __syncthreads();
        if (tid < 32) {
            int c[8], tot = 0;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                c[j] = ss.hist[lane * 8 + j];
                tot += c[j];
            }
            int incl = tot;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += t;
            }
            const int rank = ss.rank;
            __syncwarp();
            const unsigned mm = __ballot_sync(0xffffffffu, incl > rank);
            const int owner = __ffs(mm) - 1;  // always found: the candidate set holds the rank
            if (lane == owner) {
                int below = incl - tot;
                int b = 0;
                for (; b < 7; ++b) {
                    if (below + c[b] > rank) break;
                    below += c[b];
                }
                ss.prefix = prefix | ((unsigned long long)(lane * 8 + b) << shift);
                ss.rank = rank - below;
            }
        }
        __syncthreads();
    }
    const unsigned long long k1 = ss.prefix;
    double v1 = sel_unordered(k1), v2 = v1;
    if ((n & 1) == 0) {
        // second middle value: v1 again if enough copies of it, else the smallest value above
        if (tid == 0) {
            ss.cnt_le = 0;
            ss.min_gt = ~0ull;
        }
        __syncthreads();
        int le = 0;
        unsigned long long mg = ~0ull;
        for (int i = tid; i < n; i += NTHREADS) {
            const unsigned long long key = sel_ordered(get(i));
            if (key <= k1)
                ++le;
            else if (key < mg)
                mg = key;
        }
        le = __reduce_add_sync(0xffffffffu, le);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const unsigned long long t = __shfl_xor_sync(0xffffffffu, mg, o);
            mg = t < mg ? t : mg;
        }
        if (lane == 0) {
            atomicAdd(&ss.cnt_le, le);
            atomicMin(&ss.min_gt, mg);
        }
        __syncthreads();
        if ((n >> 1) >= ss.cnt_le) v2 = sel_unordered(ss.min_gt);
    }
    __syncthreads();  // ss may be reused by the next call
    return (v1 + v2) / 2.0;
}

}  // namespace icnv
