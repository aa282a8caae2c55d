// Host-side schedule of the smoothing kernel's shared-memory gathers (no device code in this file).
//
// The kernel stages a cell's raw row [G] fp32 in shared memory and every lane gathers the `gs` genes of its
// position-ordered groups from it (tl/_infercnv.py:350-351 makes the gene order a data-dependent permutation, so the
// addresses are scattered).  One warp-level LDS for slot set (warp-block wb, u) at step t reads one gene per lane; it
// costs as many shared-memory wavefronts as the largest number of distinct words that fall into one of the 32 banks.
// With groups and elements in natural order that is ~3.5 wavefronts per LDS.  Two degrees of freedom are free:
//   (1) which 32 groups share a slot set (any, as long as lane l takes groups with g % 8 == l % 8, which keeps the
//       16-byte partial-sum stores of a quarter-warp on 8 different bank groups);
//   (2) the ORDER in which a lane walks the gs genes of its group — the partial sums A = sum x, B = sum j*x only need
//       the position j of each gene, which travels in spare bits of the table entry.
// (2) turns the problem into edge colouring of the bipartite multigraph lanes x banks of a set: every lane has degree
// gs, so by Koenig's theorem gs conflict-free steps exist iff no bank holds more than gs of the set's genes.  (1) is
// therefore used to flatten the bank histogram of every set (greedy), and the steps are extracted one at a time as a
// matching that covers every lane and every bank that is "tight" (remaining genes == remaining steps); banks that are
// over-full give up their excess in the current step (multiplicity 2) so the remaining steps stay conflict-free.
// Result on the bench gene axis (20k genes, 22 chromosomes): 3.5 -> ~1.1 wavefronts per gather instruction.
#include <algorithm>
#include <cstdint>
#include <functional>
#include <vector>

namespace icnv {

namespace {

constexpr int PAD_BANK = 32;  // pseudo bank of the zero pad word (one shared word: never conflicts with itself)

inline int bank_of(int32_t col) { return col < 0 ? PAD_BANK : (col & 31); }

// wavefronts of one gather step: distinct real words per bank, the pad word counted once in its real bank
int step_wavefronts(const int32_t* cols, int n_genes) {
    int cnt[32] = {0};
    bool pad = false;
    for (int l = 0; l < 32; ++l) {
        if (cols[l] < 0)
            pad = true;
        else
            cnt[cols[l] & 31] += 1;
    }
    if (pad) cnt[n_genes & 31] += 1;
    int m = 1;
    for (int b = 0; b < 32; ++b) m = std::max(m, cnt[b]);
    return m;
}

}  // namespace

// gcol[g * gs + j]: matrix column of element j of group g (-1 = pad).  Fills
//   slot_group[set * 32 + lane] = group handled by that lane (-1 = none), sets = 4 per warp-block of quads;
//   order[(set * 32 + lane) * gs + t] = element j the lane reads at step t;
// returns the average number of wavefronts per gather instruction.
double schedule_gathers(const std::vector<int32_t>& gcol, int NG, int gs, int n_genes, int nsets, bool permute,
                        std::vector<int32_t>& slot_group, std::vector<uint8_t>& order) {
    slot_group.assign((size_t)nsets * 32, -1);
    order.assign((size_t)nsets * 32 * gs, 0);
    for (size_t i = 0; i < order.size(); ++i) order[i] = (uint8_t)(i % gs);
    auto col = [&](int32_t g, int j) { return gcol[(size_t)g * gs + j]; };

    std::vector<std::vector<int32_t>> pool(8);
    for (int32_t g = 0; g < NG; ++g) pool[g & 7].push_back(g);

    if (!permute) {
        // natural element order: take, for every lane, the group that adds the fewest same-step bank collisions
        std::vector<int> cnt((size_t)gs * 32), mx(gs);
        for (int s = 0; s < nsets; ++s) {
            std::fill(cnt.begin(), cnt.end(), 0);
            std::fill(mx.begin(), mx.end(), 0);
            for (int lane = 0; lane < 32; ++lane) {
                auto& cand = pool[lane & 7];
                if (cand.empty()) continue;
                long best = -1;
                size_t best_k = 0;
                for (size_t k = 0; k < cand.size(); ++k) {
                    long add = 0, load = 0;
                    for (int j = 0; j < gs; ++j) {
                        const int32_t c = col(cand[k], j);
                        const int nc = cnt[(size_t)j * 32 + ((c < 0 ? n_genes : c) & 31)] + 1;
                        add += nc > mx[j] ? nc - mx[j] : 0;
                        load += nc;
                    }
                    const long score = add * 4096 + load;
                    if (best < 0 || score < best) {
                        best = score;
                        best_k = k;
                    }
                }
                const int32_t g = cand[best_k];
                cand[best_k] = cand.back();
                cand.pop_back();
                slot_group[(size_t)s * 32 + lane] = g;
                for (int j = 0; j < gs; ++j) {
                    const int32_t c = col(g, j);
                    int& c2 = cnt[(size_t)j * 32 + ((c < 0 ? n_genes : c) & 31)];
                    c2 += 1;
                    mx[j] = std::max(mx[j], c2);
                }
            }
        }
    } else {
        // (1) flat bank histogram per set
        for (int s = 0; s < nsets; ++s) {
            int deg[32] = {0};
            for (int lane = 0; lane < 32; ++lane) {
                auto& cand = pool[lane & 7];
                if (cand.empty()) continue;
                long best = -1;
                size_t best_k = 0;
                for (size_t k = 0; k < cand.size(); ++k) {
                    int d[32];
                    std::copy(deg, deg + 32, d);
                    for (int j = 0; j < gs; ++j) {
                        const int b = bank_of(col(cand[k], j));
                        if (b != PAD_BANK) d[b] += 1;
                    }
                    long mxd = 0, sq = 0;
                    for (int b = 0; b < 32; ++b) {
                        mxd = std::max<long>(mxd, d[b]);
                        sq += (long)d[b] * d[b];
                    }
                    const long score = mxd * (1L << 24) + sq;
                    if (best < 0 || score < best) {
                        best = score;
                        best_k = k;
                    }
                }
                const int32_t g = cand[best_k];
                cand[best_k] = cand.back();
                cand.pop_back();
                slot_group[(size_t)s * 32 + lane] = g;
                for (int j = 0; j < gs; ++j) {
                    const int b = bank_of(col(g, j));
                    if (b != PAD_BANK) deg[b] += 1;
                }
            }
        }
        // (2) steps of every set, one matching at a time
        std::vector<char> used((size_t)32 * gs);
        for (int s = 0; s < nsets; ++s) {
            const int32_t* grp = &slot_group[(size_t)s * 32];
            std::fill(used.begin(), used.end(), 0);
            for (int t = 0; t < gs; ++t) {
                const int R = gs - t;
                int deg[32] = {0};
                for (int l = 0; l < 32; ++l)
                    if (grp[l] >= 0)
                        for (int j = 0; j < gs; ++j)
                            if (!used[l * gs + j]) {
                                const int b = bank_of(col(grp[l], j));
                                if (b != PAD_BANK) deg[b] += 1;
                            }
                // unit nodes: a bank offers max(1, excess) seats in this step; seats of banks with deg >= R are required
                int ubase[33];
                ubase[0] = 0;
                std::vector<int> unit_bank;
                std::vector<char> required;
                for (int b = 0; b < 32; ++b) {
                    const int lb = std::max(0, deg[b] - (R - 1));
                    const int cap = deg[b] == 0 ? 0 : std::max(1, lb);
                    for (int k = 0; k < cap; ++k) {
                        unit_bank.push_back(b);
                        required.push_back(k < lb);
                    }
                    ubase[b + 1] = ubase[b] + cap;
                }
                const int NU = (int)unit_bank.size();
                // lane_has[l][b]: an unused element of lane l in bank b (its index), -1 if none
                int lane_el[32][32];
                for (int l = 0; l < 32; ++l) {
                    std::fill(lane_el[l], lane_el[l] + 32, -1);
                    if (grp[l] < 0) continue;
                    for (int j = gs - 1; j >= 0; --j)
                        if (!used[l * gs + j]) {
                            const int b = bank_of(col(grp[l], j));
                            if (b != PAD_BANK) lane_el[l][b] = j;
                        }
                }
                std::vector<int> matchU(NU, -1);
                int matchL[32];
                std::fill(matchL, matchL + 32, -1);
                std::vector<char> visL(32), visU(NU);
                // augment from a unit (right) node
                std::function<bool(int)> try_unit = [&](int u) -> bool {
                    const int b = unit_bank[u];
                    for (int l = 0; l < 32; ++l) {
                        if (lane_el[l][b] < 0 || visL[l]) continue;
                        visL[l] = 1;
                        if (matchL[l] < 0 || try_unit(matchL[l])) {
                            matchL[l] = u;
                            matchU[u] = l;
                            return true;
                        }
                    }
                    return false;
                };
                // augment from a lane (left) node
                std::function<bool(int)> try_lane = [&](int l) -> bool {
                    for (int b = 0; b < 32; ++b) {
                        if (lane_el[l][b] < 0) continue;
                        for (int u = ubase[b]; u < ubase[b + 1]; ++u) {
                            if (visU[u]) continue;
                            visU[u] = 1;
                            if (matchU[u] < 0 || try_lane(matchU[u])) {
                                matchU[u] = l;
                                matchL[l] = u;
                                return true;
                            }
                        }
                    }
                    return false;
                };
                for (int u = 0; u < NU; ++u)
                    if (required[u] && matchU[u] < 0) {
                        std::fill(visL.begin(), visL.end(), 0);
                        try_unit(u);
                    }
                for (int l = 0; l < 32; ++l)
                    if (grp[l] >= 0 && matchL[l] < 0) {
                        std::fill(visU.begin(), visU.end(), 0);
                        try_lane(l);
                    }
                // commit; lanes without a seat take a pad if they still have one, else the least loaded bank
                int load[32] = {0};
                int pick[32];
                for (int l = 0; l < 32; ++l) {
                    pick[l] = -1;
                    if (grp[l] >= 0 && matchL[l] >= 0) {
                        const int b = unit_bank[matchL[l]];
                        pick[l] = lane_el[l][b];
                        load[b] += 1;
                    }
                }
                for (int l = 0; l < 32; ++l) {
                    if (grp[l] < 0 || pick[l] >= 0) continue;
                    int best_j = -1, best_load = 1 << 30;
                    for (int j = 0; j < gs; ++j) {
                        if (used[l * gs + j]) continue;
                        const int b = bank_of(col(grp[l], j));
                        const int ld = b == PAD_BANK ? -1 : load[b];
                        if (ld < best_load) {
                            best_load = ld;
                            best_j = j;
                        }
                    }
                    pick[l] = best_j;
                    const int b = bank_of(col(grp[l], best_j));
                    if (b != PAD_BANK) load[b] += 1;
                }
                for (int l = 0; l < 32; ++l)
                    if (grp[l] >= 0) {
                        used[l * gs + pick[l]] = 1;
                        order[((size_t)s * 32 + l) * gs + t] = (uint8_t)pick[l];
                    }
            }
        }
    }
    for (int c8 = 0; c8 < 8; ++c8)
        if (!pool[c8].empty()) return -1.0;  // more groups than slots: caller sized nsets wrongly

    // cost of the schedule
    long total = 0;
    for (int s = 0; s < nsets; ++s)
        for (int t = 0; t < gs; ++t) {
            int32_t cols[32];
            for (int l = 0; l < 32; ++l) {
                const int32_t g = slot_group[(size_t)s * 32 + l];
                cols[l] = g < 0 ? -1 : col(g, order[((size_t)s * 32 + l) * gs + t]);
            }
            total += step_wavefronts(cols, n_genes);
        }
    return nsets > 0 ? (double)total / ((double)nsets * gs) : 0.0;
}

}  // namespace icnv

// Host-only test hook (no device, no CUDA call): runs the schedule on a caller-supplied group table so that the CPU test
// suite can check its invariants.  slot_group_out [nsets*32], order_out [nsets*32*gs] (may be NULL).
extern "C" double icnv_host_schedule_gathers(const int32_t* gcol, int32_t n_groups, int32_t gs, int32_t n_genes, int32_t nsets,
                                             int32_t permute, int32_t* slot_group_out, uint8_t* order_out) {
    std::vector<int32_t> g(gcol, gcol + (size_t)n_groups * gs), slot;
    std::vector<uint8_t> order;
    const double cost = icnv::schedule_gathers(g, n_groups, gs, n_genes, nsets, permute != 0, slot, order);
    if (slot_group_out) std::copy(slot.begin(), slot.end(), slot_group_out);
    if (order_out) std::copy(order.begin(), order.end(), order_out);
    return cost;
}
