// Host-side construction of the fe2 kernel's constant tables (plain C++, shared by tables.cu and the CPU
// emulation in tests/emu): stage-A window phases, stage-C twiddles and the static schedule of the sparse mel
// projection.
//
// Mel schedule.  The (64 x 601) Slaney mel matrix (librosa.filters.mel, datasets.py:203) has 1165 non-zeros,
// 5..61 per filter.  Filter j is cut into ceil(nnz_j / 7) lane-jobs of nearly equal length (191 jobs for the
// 224 threads of a CTA); a job walks its <= 7 non-zeros once and accumulates BOTH frames of the tile, then
// leaves a 64-byte partial record in shared memory; the epilogue thread of (frame, mel) adds the records of
// that filter's jobs in a fixed order (deterministic, unlike shared-memory atomics).  Entries are stored
// [row][job]; the order of a job's entries is chosen by a randomised descent so that the 8 lanes of every
// quarter-warp gather their 16-byte V records from distinct 16-byte bank groups whenever possible.
#pragma once
#include <math.h>
#include <string.h>

#include <vector>

#include "fe2_core.cuh"

namespace ady {
namespace fe2 {

struct MelPlan {
    std::vector<MelEnt> ent;     // [MEL_L][NJOBS]
    int njobs[NMEL];
    int rec_off[REC_MAXJOBS + 1];   // job (mel j, chunk i) = lane j + rec_off[i]
    int job_mel[NJOBS];
    long cost, ideal;            // LDS.128 wavefronts per tile-iteration of the mel phase (both records), and its lower bound
};

// LDS.128 wavefronts of one (row, quarter-warp) cell of the gather: distinct addresses that fall into the same
// 16-byte bank group serialise, equal addresses broadcast.  Both records (offa, offb) are counted.
inline int mel_cell_cost(const std::vector<MelEnt>& ent, int row, int quarter) {
    int cost = 0;
    for (int which = 0; which < 2; ++which) {
        int cnt[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        int seen[8][8];
        for (int l = 0; l < 8; ++l) {
            const MelEnt& e = ent[(size_t)row * NJOBS + 8 * quarter + l];
            const int off = which ? e.offb : e.offa, g = (off >> 4) & 7;
            bool dup = false;
            for (int i = 0; i < cnt[g]; ++i) dup = dup || seen[g][i] == off;
            if (!dup) seen[g][cnt[g]++] = off;
        }
        int mx = 1;
        for (int g = 0; g < 8; ++g) mx = cnt[g] > mx ? cnt[g] : mx;
        cost += mx;
    }
    return cost;
}
inline long mel_plan_cost(const std::vector<MelEnt>& ent) {
    long cost = 0;
    for (int row = 0; row < MEL_L; ++row)
        for (int q = 0; q < NJOBS / 8; ++q) cost += mel_cell_cost(ent, row, q);
    return cost;
}

// mel: dense (64 x 601) row-major.  Returns false if the matrix needs more than NJOBS jobs of MEL_L entries.
inline bool build_mel_plan(const float* mel, MelPlan& p) {
    int offa[NBIN], offb[NBIN];
    v_offsets_of_bins(offa, offb);
    for (int k = 0; k < NBIN; ++k)
        if (offa[k] < 0) return false;
    // chunks of every filter, then the chunk-major job order: all first chunks (64 jobs), then the second chunks of the
    // filters that have one, ...  The number of chunks must not decrease along the mel axis (Slaney filters widen).
    std::vector<std::vector<std::vector<MelEnt>>> chunks(NMEL);
    int max_nj = 0;
    for (int j = 0; j < NMEL; ++j) {
        std::vector<MelEnt> all;
        for (int k = 0; k < NBIN; ++k)
            if (mel[(size_t)j * NBIN + k] != 0.f) all.push_back(MelEnt{(uint16_t)offa[k], (uint16_t)offb[k], mel[(size_t)j * NBIN + k]});
        if (all.empty()) return false;
        const int nj = ((int)all.size() + MEL_L - 1) / MEL_L;
        if (nj > REC_MAXJOBS || (j > 0 && nj < p.njobs[j - 1])) return false;
        p.njobs[j] = nj;
        max_nj = nj > max_nj ? nj : max_nj;
        for (int q = 0; q < nj; ++q) {
            const size_t a = all.size() * q / nj, b = all.size() * (q + 1) / nj;
            chunks[j].emplace_back(all.begin() + a, all.begin() + b);
        }
    }
    std::vector<std::vector<MelEnt>> jobs;
    std::vector<int> job_mel;
    for (int i = 0; i <= REC_MAXJOBS; ++i) p.rec_off[i] = 0;
    for (int i = 0; i < max_nj; ++i) {
        int first = 0;
        while (p.njobs[first] <= i) ++first;             // filters first .. 63 have a chunk i
        p.rec_off[i] = (int)jobs.size() - first;
        for (int j = first; j < NMEL; ++j) { jobs.push_back(chunks[j][i]); job_mel.push_back(j); }
    }
    if ((int)jobs.size() > NJOBS) return false;
    for (int q = 0; q < NJOBS; ++q) p.job_mel[q] = q < (int)jobs.size() ? job_mel[q] : -1;
    // padding entries: weight 0 at the job's own first record (idle jobs: record 0)
    p.ent.assign((size_t)MEL_L * NJOBS, MelEnt{0, 0, 0.f});
    for (int q = 0; q < NJOBS; ++q) {
        const MelEnt pad = q < (int)jobs.size() ? MelEnt{jobs[q][0].offa, jobs[q][0].offb, 0.f} : MelEnt{0, 0, 0.f};
        for (int r = 0; r < MEL_L; ++r) p.ent[(size_t)r * NJOBS + q] = r < (q < (int)jobs.size() ? (int)jobs[q].size() : 0) ? jobs[q][r] : pad;
    }
    // Randomised descent (deterministic LCG).  Moves: swap two rows of one job, or swap the entries at (row1, job1)
    // and (row2, job2) of two jobs of the SAME mel filter (any partition of a filter's non-zeros over its jobs is
    // valid; a padding entry may move too, it only has to stay a zero-weight entry).  Only the touched cells are re-costed.
    unsigned long long rng = 0x9E3779B97F4A7C15ull;
    auto next = [&rng]() { rng = rng * 6364136223846793005ull + 1442695040888963407ull; return (unsigned)(rng >> 33); };
    const int njobs_used = (int)jobs.size();
    for (int iter = 0; iter < 1600000; ++iter) {
        const int q1 = next() % njobs_used, r1 = next() % MEL_L, r2 = next() % MEL_L;
        int q2 = q1;
        if (next() & 1) {
            const int j = p.job_mel[q1];
            q2 = j + p.rec_off[next() % p.njobs[j]];
        }
        if (q1 == q2 && r1 == r2) continue;
        const int cells[4][2] = {{r1, q1 / 8}, {r2, q2 / 8}, {r1, q2 / 8}, {r2, q1 / 8}};
        auto touched = [&]() {
            long c = mel_cell_cost(p.ent, cells[0][0], cells[0][1]);
            if (cells[1][0] != cells[0][0] || cells[1][1] != cells[0][1]) c += mel_cell_cost(p.ent, cells[1][0], cells[1][1]);
            return c;
        };
        const long before = touched();
        MelEnt& a = p.ent[(size_t)r1 * NJOBS + q1];
        MelEnt& b = p.ent[(size_t)r2 * NJOBS + q2];
        MelEnt t = a; a = b; b = t;
        if (touched() > before) { t = a; a = b; b = t; }
    }
    p.cost = mel_plan_cost(p.ent);
    p.ideal = (long)MEL_L * (NJOBS / 8) * 2;
    return true;
}

// Column permutation of the staged audio (fe2_core.cuh "staging map").  Lane l is an edge between its reader half-warp
// l / 16 and its writer half-warp (16 l mod 75) / 16 (the copy thread rem = 16 l mod 75 writes lane l's column); a
// proper edge colouring of this bipartite multigraph with 16 colours (degree <= 16: Koenig) gives every lane a bank
// pair that is unique within both half-warps, and a colour class is a matching of <= 5 edges, so the lanes of one
// colour take the columns colour, colour + 16, ... < 80.
inline bool build_col_perm(uint8_t perm[80]) {
#ifdef ADY_FE2_LINEAR_STAGING
    // Measurement variant (tools/build_variant.py lin -DADY_FE2_LINEAR_STAGING): the layout a bulk / TMA copy would produce,
    // sample N of the tile at row N / 75, column N mod 75.  Lane l then reads column 16 l mod 75, and because the row
    // pitch is a multiple of 16 samples the lanes of a half-warp fall into bank pairs 0,0,0,0,0,5,5,5,5,5,10,... (5-way
    // conflicts on every stage-A sample load); numbers in DESIGN.md section 4.1.
    for (int i = 0; i < 80; ++i) perm[i] = 0;
    for (int l = 0; l < 75; ++l) perm[l] = (uint8_t)((16 * l) % 75);
    return true;
#endif
    int colour[75], at_r[5][16], at_w[5][16];           // at_x[node][colour] = lane using that colour at the node, or -1
    for (int n = 0; n < 5; ++n)
        for (int c = 0; c < 16; ++c) at_r[n][c] = at_w[n][c] = -1;
    auto rnode = [](int l) { return l / 16; };
    auto wnode = [](int l) { return ((16 * l) % 75) / 16; };
    for (int l = 0; l < 75; ++l) {
        const int u = rnode(l), v = wnode(l);
        int a = 0, b = 0;
        while (a < 16 && at_r[u][a] >= 0) ++a;          // free at the reader node
        while (b < 16 && at_w[v][b] >= 0) ++b;          // free at the writer node
        if (a == 16 || b == 16) return false;
        if (at_w[v][a] >= 0) {
            // a is taken at v: swap a <-> b along the alternating path that starts at v with colour a
            int node = v, side = 1 /* writer */, want = a, other = b;
            int path[80], np = 0;
            while (true) {
                const int e = side ? at_w[node][want] : at_r[node][want];
                if (e < 0) break;
                path[np++] = e;
                node = side ? rnode(e) : wnode(e);
                side ^= 1;
                const int t = want; want = other; other = t;
            }
            for (int i = 0; i < np; ++i) {               // detach, then re-attach with the other colour
                const int e = path[i];
                at_r[rnode(e)][colour[e]] = -1;
                at_w[wnode(e)][colour[e]] = -1;
            }
            for (int i = 0; i < np; ++i) {
                const int e = path[i];
                colour[e] = colour[e] == a ? b : a;
                at_r[rnode(e)][colour[e]] = e;
                at_w[wnode(e)][colour[e]] = e;
            }
        }
        if (at_r[u][a] >= 0 || at_w[v][a] >= 0) return false;
        colour[l] = a;
        at_r[u][a] = l;
        at_w[v][a] = l;
    }
    int used[16] = {0};
    for (int i = 0; i < 80; ++i) perm[i] = 0;
    for (int l = 0; l < 75; ++l) {
        const int col = colour[l] + 16 * used[colour[l]]++;
        if (col >= 80) return false;
        perm[l] = (uint8_t)col;
    }
    // verify: a permutation into [0, 80), bank pairs distinct within every reader and writer half-warp
    bool seen[80] = {false};
    for (int l = 0; l < 75; ++l) {
        if (seen[perm[l]]) return false;
        seen[perm[l]] = true;
        for (int k = 0; k < l; ++k)
            if ((rnode(k) == rnode(l) || wnode(k) == wnode(l)) && perm[k] % 16 == perm[l] % 16) return false;
    }
    return true;
}

inline void fill_tables(const float* mel, Tables& t, MelPlan& plan, bool& ok) {
    memset(&t, 0, sizeof(t));
    if (!build_col_perm(t.col_perm)) { ok = false; return; }
    for (int m = 0; m < 16; ++m)                  // 44 of 75 columns: the other lanes read the mirror image (stage_a_const)
        for (int l = 0; l < WIN_DIRECT; ++l)
            t.win[m * WIN_P + l] = (float)((0.5 - 0.5 * cos(2.0 * M_PI * (double)((16 * l) % 75 + 75 * m) / 1200.0)) / 65536.0);
    for (int c = 0; c < 15; ++c)
        for (int b = 1; b < 5; ++b) {
            const double ang = -2.0 * M_PI * (double)(b * c) / 75.0;
            float* e = t.tw75 + (c * 4 + (b - 1)) * 2;
            e[0] = (float)cos(ang);
            e[1] = (float)sin(ang);
        }
    ok = build_mel_plan(mel, plan);
    if (!ok) return;
    for (size_t i = 0; i < plan.ent.size(); ++i) t.ent[i] = plan.ent[i];
    for (int j = 0; j < NMEL; ++j) t.mel_njobs[j] = (uint8_t)plan.njobs[j];
    for (int i = 0; i <= REC_MAXJOBS; ++i) t.rec_off[i] = (int16_t)plan.rec_off[i];
    for (int q = 0; q < NJOBS; ++q) t.job_mel[q] = (uint8_t)(plan.job_mel[q] < 0 ? 0 : plan.job_mel[q]);
}

}  // namespace fe2
}  // namespace ady
