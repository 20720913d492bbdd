// Host-side construction of the fused front end's constant tables (plain C++, shared by
// tables.cu and the CPU emulation in tests/emu): window phase factors and the static schedule of
// the sparse mel projection.
//
// Mel schedule.  The (64 x 601) Slaney mel matrix has 1165 non-zeros (5..61 per filter).  Filter
// j is split into two halves handled by adjacent lanes (j, part); 16 filters x 2 parts form one
// warp-task, 4 warp-tasks cover a frame.  All lanes of a warp-task run the same number of
// iterations (zero-weight padding), entries are stored [warp-task][iteration][lane] so the entry
// load is one coalesced 256-byte access, and the order of a lane's entries is chosen so that the
// 8 lanes of every quarter-warp hit distinct 16-byte slots of a 128-byte shared-memory line
// whenever possible (the V values are gathered with LDS.128).
#pragma once
#include <math.h>
#include <string.h>

#include <vector>

#include "frontend_core.cuh"

namespace ady {

struct MelSchedule {
    int it0[4], nit[4];                 // first iteration row / iteration count of each warp-task
    long cost[4];                       // LDS.128 quarter-wavefronts per V load of each warp-task
    int total_rows;
    std::vector<MelEntry> ent;          // [total_rows][32]
};

// pos_of_bin: V position (25 t + k2) holding FFT bin k (0..600)
inline void pfa_pos_of_bin(int (&pos_of_bin)[NBIN]) {
    for (int k = 0; k < NBIN; ++k) pos_of_bin[k] = -1;
    for (int t = 0; t < 25; ++t)
        for (int k2 = 0; k2 < 25; ++k2) {
            int k = pfa_out(t, k2);
            if (k > 600) k = 1200 - k;
            if (pos_of_bin[k] < 0) pos_of_bin[k] = 25 * t + k2;
        }
}

// mel: dense (64 x 601) row-major.  Returns false if the schedule does not fit MEL_MAXROWS.
inline bool build_mel_schedule(const float* mel, MelSchedule& s) {
    int pos_of_bin[NBIN];
    pfa_pos_of_bin(pos_of_bin);
    s.ent.clear();
    int row = 0;
    for (int wt = 0; wt < 4; ++wt) {
        // lane lists: lane = 2*(j - 16 wt) + part
        std::vector<std::vector<MelEntry>> lanes(32);
        int nit = 0;
        for (int jj = 0; jj < 16; ++jj) {
            const int j = 16 * wt + jj;
            std::vector<MelEntry> all;
            for (int k = 0; k < NBIN; ++k)
                if (mel[(size_t)j * NBIN + k] != 0.f) all.push_back(MelEntry{pos_of_bin[k], mel[(size_t)j * NBIN + k]});
            const int h = ((int)all.size() + 1) / 2;
            lanes[2 * jj].assign(all.begin(), all.begin() + h);
            lanes[2 * jj + 1].assign(all.begin() + h, all.end());
            if (h > nit) nit = h;
        }
        s.it0[wt] = row;
        s.nit[wt] = nit;
        s.ent.resize((size_t)(row + nit) * 32, MelEntry{0, 0.f});
        // Randomised greedy with restarts (deterministic LCG).  Per iteration and quarter-warp the
        // lanes with the least slack choose first; a lane takes an entry whose 16-byte slot
        // (pos mod 8) is still free in this quarter, may idle (padding) while it has slack, and
        // otherwise takes the least loaded slot.  Cost = sum over quarter-iterations of the
        // maximum slot multiplicity == LDS.128 wavefronts.
        unsigned long long rng = 0x9E3779B97F4A7C15ull + wt;
        auto next = [&rng]() { rng = rng * 6364136223846793005ull + 1442695040888963407ull; return (unsigned)(rng >> 33); };
        long best_cost = -1;
        std::vector<int> best((size_t)nit * 32, -1), cur((size_t)nit * 32);
        for (int attempt = 0; attempt < 300; ++attempt) {
            std::vector<std::vector<int>> rem(32);
            for (int l = 0; l < 32; ++l) {
                rem[l].resize(lanes[l].size());
                for (size_t e = 0; e < lanes[l].size(); ++e) rem[l][e] = (int)e;
                for (size_t e = rem[l].size(); e > 1; --e) { size_t o = next() % e; int t = rem[l][e - 1]; rem[l][e - 1] = rem[l][o]; rem[l][o] = t; }
            }
            std::fill(cur.begin(), cur.end(), -1);
            long cost = 0;
            bool ok = true;
            for (int it = 0; it < nit && ok; ++it) {
                const int left = nit - it;
                for (int q = 0; q < 4; ++q) {
                    int load[8] = {0, 0, 0, 0, 0, 0, 0, 0};
                    int order[8];
                    for (int i = 0; i < 8; ++i) order[i] = 8 * q + i;
                    for (int i = 7; i > 0; --i) { int o = next() % (i + 1); int t = order[i]; order[i] = order[o]; order[o] = t; }
                    for (int a2 = 0; a2 < 8; ++a2)            // stable-ish sort: most remaining first
                        for (int b2 = a2 + 1; b2 < 8; ++b2)
                            if (rem[order[b2]].size() > rem[order[a2]].size()) { int t = order[a2]; order[a2] = order[b2]; order[b2] = t; }
                    for (int a2 = 0; a2 < 8; ++a2) {
                        const int l = order[a2];
                        if (rem[l].empty()) continue;
                        const int slack = left - (int)rem[l].size();
                        int pick = -1, pick_score = -1;
                        for (size_t e = 0; e < rem[l].size(); ++e) {
                            const int slot = lanes[l][rem[l][e]].pos & 7;
                            if (load[slot]) continue;
                            int score = 0;                      // prefer the slot most common in the own list
                            for (int e2 : rem[l]) score += ((lanes[l][e2].pos & 7) == slot);
                            if (score > pick_score) { pick_score = score; pick = (int)e; }
                        }
                        if (pick < 0) {
                            if (slack > 0) continue;            // idle this step (zero-weight padding)
                            int bl = 1 << 30;
                            for (size_t e = 0; e < rem[l].size(); ++e) {
                                const int ld = load[lanes[l][rem[l][e]].pos & 7];
                                if (ld < bl) { bl = ld; pick = (int)e; }
                            }
                        }
                        const int eidx = rem[l][pick];
                        rem[l].erase(rem[l].begin() + pick);
                        load[lanes[l][eidx].pos & 7]++;
                        cur[(size_t)it * 32 + l] = eidx;
                    }
                    int mx = 1;
                    for (int i = 0; i < 8; ++i) mx = load[i] > mx ? load[i] : mx;
                    cost += mx;
                }
            }
            for (int l = 0; l < 32; ++l) ok = ok && rem[l].empty();
            if (ok && (best_cost < 0 || cost < best_cost)) { best_cost = cost; best = cur; }
        }
        if (best_cost < 0) return false;
        for (int it = 0; it < nit; ++it)
            for (int l = 0; l < 32; ++l) {
                const int e = best[(size_t)it * 32 + l];
                if (e >= 0) s.ent[(size_t)(row + it) * 32 + l] = lanes[l][e];
                else {
                    // padding: weight 0 at a position already read by a neighbour lane of the same
                    // quarter (identical address -> broadcast, no extra wavefront)
                    int p = 0;
                    for (int l2 = 8 * (l / 8); l2 < 8 * (l / 8) + 8; ++l2)
                        if (best[(size_t)it * 32 + l2] >= 0) { p = lanes[l2][best[(size_t)it * 32 + l2]].pos; break; }
                    s.ent[(size_t)(row + it) * 32 + l] = MelEntry{p, 0.f};
                }
            }
        s.cost[wt] = best_cost;
        row += nit;
    }
    s.total_rows = row;
    return row <= MEL_MAXROWS;
}

}  // namespace ady
