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
    int job0[NMEL], njobs[NMEL];
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
    std::vector<std::vector<MelEnt>> jobs;
    for (int j = 0; j < NMEL; ++j) {
        std::vector<MelEnt> all;
        for (int k = 0; k < NBIN; ++k)
            if (mel[(size_t)j * NBIN + k] != 0.f) all.push_back(MelEnt{(uint16_t)offa[k], (uint16_t)offb[k], mel[(size_t)j * NBIN + k]});
        if (all.empty()) return false;
        const int nj = ((int)all.size() + MEL_L - 1) / MEL_L;
        if (nj > REC_MAXJOBS) return false;
        p.job0[j] = (int)jobs.size();
        p.njobs[j] = nj;
        for (int q = 0; q < nj; ++q) {
            const size_t a = all.size() * q / nj, b = all.size() * (q + 1) / nj;
            jobs.emplace_back(all.begin() + a, all.begin() + b);
        }
    }
    if ((int)jobs.size() > NJOBS) return false;
    for (int q = 0; q < NJOBS; ++q) p.job_mel[q] = -1;
    for (int j = 0; j < NMEL; ++j)
        for (int q = 0; q < p.njobs[j]; ++q) p.job_mel[p.job0[j] + q] = j;
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
    for (int iter = 0; iter < 400000; ++iter) {
        const int q1 = next() % njobs_used, r1 = next() % MEL_L, r2 = next() % MEL_L;
        int q2 = q1;
        if (next() & 1) {
            const int j = p.job_mel[q1];
            q2 = p.job0[j] + (int)(next() % p.njobs[j]);
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

inline void fill_tables(const float* mel, Tables& t, MelPlan& plan, bool& ok) {
    memset(&t, 0, sizeof(t));
    for (int n16 = 0; n16 < 16; ++n16)
        for (int l = 0; l < 75; ++l)
            t.win[n16 * 80 + l] = (float)((0.5 - 0.5 * cos(2.0 * M_PI * (double)stage_a_sample(l, n16) / 1200.0)) / 65536.0);
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
    for (int j = 0; j < NMEL; ++j) { t.mel_job0[j] = (uint8_t)plan.job0[j]; t.mel_njobs[j] = (uint8_t)plan.njobs[j]; }
    for (int q = 0; q < NJOBS; ++q) t.job_mel[q] = (uint8_t)(plan.job_mel[q] < 0 ? 0 : plan.job_mel[q]);
}

}  // namespace fe2
}  // namespace ady
