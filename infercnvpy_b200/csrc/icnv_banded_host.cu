// Host-only layout of the EXPERIMENTAL banded row-pair kernel (icnv_smooth_banded.cu; off by default).
// The task list is cut at a tile boundary; band X keeps every group its tasks read (groups read by both bands are
// duplicated), band B's groups are re-based behind a gap of PAD_GROUPS zero groups, every band gets its own gather
// schedule (icnv_schedule.cu) and band A's work units come first in the tables.
#include <algorithm>

#include "icnv_common.cuh"

namespace icnv {

int banded_layout(const std::vector<int32_t>& gcol, int NG, int gs, int NQ, const std::vector<Task>& tasks, int n_genes,
                  uint32_t raw_base, bool optimise_walk, BandedLayout& L) {
    L = BandedLayout();
    const int n_tasks = (int)tasks.size();
    const int n_tiles = (n_tasks + 31) / 32;
    if (n_tiles < 2 || n_tiles > 8) return 0;
    const int TA = (n_tiles + 1) / 2;
    const int tA = 32 * TA;  // < n_tasks because n_tiles >= 2
    // groups a task reads: z windows of NQ groups starting at x, x+1, ... (flat chromosomes: their z groups)
    auto need_end = [&](const Task& t) { return (t.w & 0xFF) ? t.x + t.z : t.x + (t.z - 1) + NQ; };
    int32_t gA_end = 0;
    for (int t = 0; t < tA; ++t) gA_end = std::max(gA_end, need_end(tasks[t]));
    const int32_t gB_start = tasks[tA].x;
    const int32_t baseB = (gA_end + PAD_GROUPS + 7) / 8 * 8;  // multiple of 8: a lane keeps its bank-group residue
    const int32_t first_group[2] = {0, gB_start}, count[2] = {gA_end, NG - gB_start}, phys0[2] = {0, baseB};
    L.NG = baseB + count[1];
    L.NGpad = (L.NG + 3) / 4 * 4;
    const int32_t dump[2] = {gA_end, L.NGpad};  // where empty slots store their zeros: the gap / the tail pad
    int n_wb[2];
    std::vector<int32_t> slot[2];
    std::vector<uint8_t> order[2];
    for (int b = 0; b < 2; ++b) {
        const int nquads = (count[b] + 3) / 4;
        n_wb[b] = std::max(1, (nquads + 31) / 32);
        std::vector<int32_t> sub(gcol.begin() + (size_t)first_group[b] * gs, gcol.begin() + (size_t)(first_group[b] + count[b]) * gs);
        if (schedule_gathers(sub, count[b], gs, n_genes, n_wb[b] * 4, optimise_walk, slot[b], order[b]) < 0) return -1;
    }
    const size_t n_entries = (size_t)(n_wb[0] + n_wb[1]) * gs * 32 * 4;
    L.off.assign(n_entries, raw_base + (uint32_t)n_genes * 4u);
    L.cols.assign(n_entries, -1);
    L.grp.assign((size_t)(n_wb[0] + n_wb[1]) * 32 * 4, 0);
    for (int b = 0; b < 2; ++b)
        for (int wb = 0; wb < n_wb[b]; ++wb) {
            const size_t unit = (size_t)(b ? n_wb[0] : 0) + wb;
            for (int lane = 0; lane < 32; ++lane)
                for (int u = 0; u < 4; ++u) {
                    const int32_t gl = slot[b][((size_t)wb * 4 + u) * 32 + lane];  // group inside the band, -1 = none
                    L.grp[(unit * 32 + lane) * 4 + u] = gl < 0 ? dump[b] : phys0[b] + gl;
                    for (int t = 0; t < gs; ++t) {
                        const size_t e = ((unit * gs + t) * 32 + lane) * 4 + u;
                        int j = t;
                        if (gl >= 0) {
                            j = order[b][(((size_t)wb * 4 + u) * 32 + lane) * gs + t];  // element read at step t
                            const int32_t col = gcol[(size_t)(first_group[b] + gl) * gs + j];
                            if (col >= 0) {
                                L.off[e] = raw_base + (uint32_t)col * 4u;
                                L.cols[e] = col;
                            }
                        }
                        L.off[e] |= (uint32_t)j << 24;
                    }
                }
        }
    L.tasks = tasks;
    for (int t = tA; t < n_tasks; ++t) L.tasks[t].x = baseB + (L.tasks[t].x - gB_start);
    L.units[0] = n_wb[0];
    L.units[1] = n_wb[1];
    L.tile0[0] = 0;
    L.tile0[1] = TA;
    L.tiles[0] = TA;
    L.tiles[1] = n_tiles - TA;
    L.on = true;
    return 0;
}

}  // namespace icnv

// Host-only test hook: the layout above on caller-supplied groups and tasks.  meta_out [9] = {on, NG, NGpad, units A, units B,
// tile0 B, tiles A, tiles B, n_entries}; the arrays receive at most cap_* elements (nothing is written when too small).
extern "C" int icnv_host_banded_layout(const int32_t* gcol, int32_t n_groups, int32_t gs, int32_t nq, const int32_t* tasks4,
                                       int32_t n_tasks, int32_t n_genes, uint32_t raw_base, int32_t* meta_out, uint32_t* off_out,
                                       int32_t* cols_out, int64_t cap_entries, int32_t* grp_out, int64_t cap_grp,
                                       int32_t* tasks4_out) {
    std::vector<int32_t> g(gcol, gcol + (size_t)n_groups * gs);
    std::vector<icnv::Task> t(n_tasks);
    for (int i = 0; i < n_tasks; ++i) t[i] = {tasks4[4 * i], tasks4[4 * i + 1], tasks4[4 * i + 2], tasks4[4 * i + 3]};
    icnv::BandedLayout L;
    const int rc = icnv::banded_layout(g, n_groups, gs, nq, t, n_genes, raw_base, true, L);
    if (rc) return rc;
    const int32_t meta[9] = {L.on, L.NG, L.NGpad, L.units[0], L.units[1], L.tile0[1], L.tiles[0], L.tiles[1], (int32_t)L.off.size()};
    std::copy(meta, meta + 9, meta_out);
    if (!L.on) return 0;
    if ((int64_t)L.off.size() > cap_entries || (int64_t)L.grp.size() > cap_grp) return -2;
    std::copy(L.off.begin(), L.off.end(), off_out);
    std::copy(L.cols.begin(), L.cols.end(), cols_out);
    std::copy(L.grp.begin(), L.grp.end(), grp_out);
    for (int i = 0; i < n_tasks; ++i) {
        tasks4_out[4 * i] = L.tasks[i].x;
        tasks4_out[4 * i + 1] = L.tasks[i].y;
        tasks4_out[4 * i + 2] = L.tasks[i].z;
        tasks4_out[4 * i + 3] = L.tasks[i].w;
    }
    return 0;
}
