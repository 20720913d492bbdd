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
        // greedy conflict-avoiding order: per iteration and quarter-warp, lanes with the fewest
        // choices pick first, taking an entry whose slot (pos mod 8) is still free
        std::vector<std::vector<char>> used(32);
        for (int l = 0; l < 32; ++l) used[l].assign(lanes[l].size(), 0);
        for (int it = 0; it < nit; ++it) {
            for (int q = 0; q < 4; ++q) {
                bool slot_taken[8] = {false, false, false, false, false, false, false, false};
                int order[8];
                for (int i = 0; i < 8; ++i) order[i] = 8 * q + i;
                // lanes with fewer remaining entries first (less freedom)
                for (int a = 0; a < 8; ++a)
                    for (int b = a + 1; b < 8; ++b) {
                        int ra = 0, rb = 0;
                        for (char u : used[order[a]]) ra += !u;
                        for (char u : used[order[b]]) rb += !u;
                        if (rb < ra) { int t = order[a]; order[a] = order[b]; order[b] = t; }
                    }
                for (int a = 0; a < 8; ++a) {
                    const int l = order[a];
                    int pick = -1;
                    for (size_t e = 0; e < lanes[l].size(); ++e)
                        if (!used[l][e] && !slot_taken[lanes[l][e].pos & 7]) { pick = (int)e; break; }
                    if (pick < 0)
                        for (size_t e = 0; e < lanes[l].size(); ++e)
                            if (!used[l][e]) { pick = (int)e; break; }
                    if (pick < 0) continue;  // lane exhausted: zero-weight padding stays
                    // must finish within nit iterations: a lane may not idle while it still has
                    // more entries than iterations left (never happens: sizes <= nit)
                    used[l][pick] = 1;
                    slot_taken[lanes[l][pick].pos & 7] = true;
                    s.ent[(size_t)(row + it) * 32 + l] = lanes[l][pick];
                }
            }
        }
        // safety: every entry placed
        for (int l = 0; l < 32; ++l)
            for (char u : used[l])
                if (!u) return false;
        row += nit;
    }
    s.total_rows = row;
    return row <= MEL_MAXROWS;
}

}  // namespace ady
