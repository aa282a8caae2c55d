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


// Faster exact median for rows whose values cluster around their mean (the per-gene layer): one pass for the moments, one
// pass into per-warp histograms of HB uniform bins over mean +- 4 sigma (per-warp copies: the central bins hold ~1 % of the
// values each, a single shared histogram would serialise on them), then the few values of the bin(s) holding the two
// middle ranks are collected and ranked exactly.  Falls back to the radix selection above whenever the shortcut does not
// apply (tiny / constant rows, a middle rank in a clamp bin, too many values in the middle bins).  Same result as
// cta_median: both return np.median of the values.
constexpr int SEL_HB = 512;     // histogram bins
constexpr int SEL_CAP = 2048;   // candidates ranked exactly
struct HistSmem {
    double sum[32], sq[32];
    double lo, inv_w, v1, v2;
    int b1, b2, below, n_cand, ok;
};
// work: int[NTHREADS / 32][SEL_HB] per-warp histograms followed by double[SEL_CAP] candidates
template <int NTHREADS, typename F>
__device__ double cta_median_hist(F get, int n, SelectSmem& ss, HistSmem& hs, unsigned char* work) {
    constexpr int NWARPS = NTHREADS / 32;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (n < 4 * SEL_HB) return cta_median<NTHREADS>(get, n, ss);
    int* whist = reinterpret_cast<int*>(work);
    double* cand = reinterpret_cast<double*>(work + (size_t)NWARPS * SEL_HB * sizeof(int));
    // ---- moments
    double s = 0.0, q = 0.0;
    for (int i = tid; i < n; i += NTHREADS) {
        const double x = get(i);
        s += x;
        q = fma(x, x, q);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        s += __shfl_xor_sync(0xffffffffu, s, o);
        q += __shfl_xor_sync(0xffffffffu, q, o);
    }
    if (lane == 0) {
        hs.sum[warp] = s;
        hs.sq[warp] = q;
    }
    for (int i = tid; i < NWARPS * SEL_HB; i += NTHREADS) whist[i] = 0;
    __syncthreads();
    if (tid == 0) {
        double S = 0.0, Q = 0.0;
        for (int w = 0; w < NWARPS; ++w) {
            S += hs.sum[w];
            Q += hs.sq[w];
        }
        const double mean = S / n, var = Q / n - mean * mean;
        const double sd = var > 0.0 ? sqrt(var) : 0.0;
        hs.ok = (sd > 0.0 && isfinite(sd)) ? 1 : 0;
        hs.lo = mean - 4.0 * sd;
        hs.inv_w = hs.ok ? (double)SEL_HB / (8.0 * sd) : 0.0;
        hs.n_cand = 0;
    }
    __syncthreads();
    if (!hs.ok) return cta_median<NTHREADS>(get, n, ss);
    const double lo = hs.lo, inv_w = hs.inv_w;
    auto bin_of = [&](double x) {
        const double t = (x - lo) * inv_w;
        return t < 0.0 ? 0 : (t >= (double)(SEL_HB - 1) ? SEL_HB - 1 : (int)t);
    };
    for (int i = tid; i < n; i += NTHREADS) atomicAdd(&whist[warp * SEL_HB + bin_of(get(i))], 1);
    __syncthreads();
    for (int b = tid; b < SEL_HB; b += NTHREADS) {
        int t = 0;
        for (int w = 0; w < NWARPS; ++w) t += whist[w * SEL_HB + b];
        whist[b] = t;  // row 0 now holds the totals (each thread only touches its own columns)
    }
    __syncthreads();
    if (tid < 32) {
        constexpr int PER = SEL_HB / 32;
        int c[PER], tot = 0;
#pragma unroll
        for (int j = 0; j < PER; ++j) {
            c[j] = whist[lane * PER + j];
            tot += c[j];
        }
        int incl = tot;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        const int r1 = (n - 1) >> 1, r2 = n >> 1;
        int below = incl - tot;
        for (int j = 0; j < PER; ++j) {
            if (r1 >= below && r1 < below + c[j]) {
                hs.b1 = lane * PER + j;
                hs.below = below;
            }
            if (r2 >= below && r2 < below + c[j]) hs.b2 = lane * PER + j;
            below += c[j];
        }
    }
    __syncthreads();
    const int b1 = hs.b1, b2 = hs.b2;
    int m = 0;
    for (int b = b1; b <= b2; ++b) m += whist[b];
    if (b1 == 0 || b2 == SEL_HB - 1 || m > SEL_CAP) {
        __syncthreads();
        return cta_median<NTHREADS>(get, n, ss);
    }
    for (int i = tid; i < n; i += NTHREADS) {
        const double x = get(i);
        const int b = bin_of(x);
        if (b >= b1 && b <= b2) cand[atomicAdd(&hs.n_cand, 1)] = x;
    }
    __syncthreads();
    const int k1 = ((n - 1) >> 1) - hs.below, k2 = (n >> 1) - hs.below;
    for (int i = tid; i < m; i += NTHREADS) {
        const double x = cand[i];
        int less = 0;
        for (int j = 0; j < m; ++j) {
            const double y = cand[j];
            less += (y < x || (y == x && j < i)) ? 1 : 0;
        }
        if (less == k1) hs.v1 = x;
        if (less == k2) hs.v2 = x;
    }
    __syncthreads();
    const double res = (hs.v1 + hs.v2) / 2.0;
    __syncthreads();  // hs / work may be reused by the next call
    return res;
}

}  // namespace icnv
