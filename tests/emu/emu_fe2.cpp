// CPU emulation of fe2_foa_kernel (TEST INFRASTRUCTURE): runs the same per-thread functions
// (ad-yolo_b200/csrc/fe2_core.cuh) thread by thread, phase by phase, over a fake shared memory, so the index
// logic (staging map, PFA / Cooley-Tukey maps, in-place V layout, mel schedule) is checked against the oracle
// without a GPU.  Mirrors the phase structure of fe2.cu::fe2_kernel.
#include <math.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "../../ad-yolo_b200/csrc/fe2_tables.h"
using namespace ady;
using namespace ady::fe2;

static void stage_copy(unsigned char* samp, const int16_t* clip /* first sample of the clip */, int t0, int nf, const uint8_t* col_perm) {
    for (int tid = 0; tid < NT_AB; ++tid) {
        const int h = tid / 80, rem = tid % 80;
        if (rem >= 75) continue;
        const int col = col_perm[stage_col(rem)];
        for (int i = 0; i < 12; ++i) {
            const int J = 12 * h + i;
            if (J >= 8 * (nf + 1)) continue;
            long long s = 600LL * (t0 - 1) + 75 * J + rem;
            if (s < 0) s = -s;                                   // reflect padding of frame 0
            memcpy(samp + (J * ROWP + col) * 8, clip + s * 4, 8);
        }
    }
}

// rot: per-clip RotationAug combination bits (0 = none), see frontend_core.cuh::rot_bits
extern "C" int emu_fe2_features_foa(const int16_t* audio, int B, long long N, const float* mel_dense /*64x601*/,
                                    const float* mean, const float* istd, float dc_offset, float top_db,
                                    const unsigned char* rot_bits_per_clip, float* out /*B,7,T,64*/, long* plan_cost /*[2] or NULL*/) {
    const int T = (int)(N / HOP);
    const int tpc = (T + TFR - 1) / TFR;
    static Tables tab;
    static MelPlan plan;
    static bool built = false;
    if (!built) {
        bool ok;
        fill_tables(mel_dense, tab, plan, ok);
        if (!ok) return -1;
        built = true;
    }
    if (plan_cost) { plan_cost[0] = plan.cost; plan_cost[1] = plan.ideal; }
    std::vector<unsigned char> smem(SmemLayout::total, 0);
    unsigned char* s_samp = smem.data() + SmemLayout::off_samples;
    unsigned char* s_x = smem.data() + SmemLayout::off_x;
    const unsigned char* s_tw = reinterpret_cast<const unsigned char*>(tab.tw75);
    const float dc0 = dc_offset * 300.0f, dc1 = -dc_offset * 150.0f;
    std::vector<float> gmax((size_t)B * 4, -INFINITY);
    for (int tile = 0; tile < B * tpc; ++tile) {
        const int b = tile / tpc, t0 = (tile % tpc) * TFR, nf = std::min(TFR, T - t0);
        const unsigned rb = rot_bits_per_clip ? rot_bits_per_clip[b] : 0u;
        stage_copy(s_samp, audio + (long long)b * N * 4, t0, nf, tab.col_perm);
        for (int tid = 0; tid < NT_AB; ++tid) {                   // stage A
            const int f = tid / 80, l = tid % 80;
            if (l < 75 && f < nf) stage_a(s_samp, tab.win, s_x, f, l, stage_a_const(l, tab.col_perm[l]));
        }
        for (int tid = 0; tid < NT_AB; ++tid) {                   // stage B
            const int f = tid / 80, u = tid % 80;
            if (f < nf) stage_b(s_x, f, u);
        }
        for (int tid = 0; tid < 2 * NREG; ++tid) {                // stage C: the 224 regular pair-tasks, one per thread
            const int f = tid >= NREG, task = tid - f * NREG;
            if (f < nf) stage_c_foa<false>(s_x + f * X_STRIDE, s_tw, task, dc0, dc1, rb);
        }
        for (int tid = 0; tid < 2 * NC0; ++tid) {                 // ... and the 18 c = 0 pair-tasks
            const int f = tid >= NC0, task = tid - f * NC0;
            if (f < nf) stage_c_foa<true>(s_x + f * X_STRIDE, s_tw, task, dc0, dc1, rb);
        }
        std::vector<f2> accs((size_t)NJOBS * TFR * 4);             // mel jobs (all V reads happen before any record is written)
        for (int q = 0; q < NJOBS; ++q) {
            f2 acc[TFR][4];
            mel_job<true>(s_x, tab.ent + q, acc);
            for (int f = 0; f < TFR; ++f)
                for (int i = 0; i < 4; ++i) accs[(q * TFR + f) * 4 + i] = acc[f][i];
        }
        for (int q = 0; q < NJOBS; ++q)                           // partial records over the frame buffers, 4 planes
            for (int f = 0; f < TFR; ++f) {
                const f2* a = &accs[(q * TFR + f) * 4];
                st_f4(s_x + (2 * f) * REC_PLANE + q * 16, lo2(a[0]), hi2(a[0]), lo2(a[1]), hi2(a[1]));
                st_f4(s_x + (2 * f + 1) * REC_PLANE + q * 16, lo2(a[2]), hi2(a[2]), lo2(a[3]), hi2(a[3]));
            }
        for (int e = 0; e < 128; ++e) {                           // epilogue: thread = (frame, mel)
            const int f = e >> 6, j = e & 63;
            if (f >= nf) continue;
            float v[8] = {0, 0, 0, 0, 0, 0, 0, 0};
            for (int i = 0; i < tab.mel_njobs[j]; ++i) {
                float r[8];
                memcpy(r, s_x + (2 * f) * REC_PLANE + (j + tab.rec_off[i]) * 16, 16);
                memcpy(r + 4, s_x + (2 * f + 1) * REC_PLANE + (j + tab.rec_off[i]) * 16, 16);
                for (int c = 0; c < 8; ++c) v[c] += r[c];
            }
            // record order (|W|^2,|Z|^2,|Y|^2,|X|^2, I_Y, I_Z, I_X, -) -> channels mel W,Y,Z,X, iv Y,Z,X
            float ch[7] = {power_to_db(v[0]), power_to_db(v[2]), power_to_db(v[1]), power_to_db(v[3]), v[4], v[5], v[6]};
            if (rb & 1u) ch[4] = -ch[4];
            if (rb & 2u) ch[5] = -ch[5];
            if (rb & 4u) ch[6] = -ch[6];
            if (rb & 8u) { std::swap(ch[1], ch[3]); std::swap(ch[4], ch[6]); }
            const long long tt = t0 + f;
            for (int c = 0; c < 7; ++c) {
                if (c < 4) gmax[b * 4 + c] = std::max(gmax[b * 4 + c], ch[c]);
                const float mu = mean ? mean[c * NMEL + j] : 0.f, is = istd ? istd[c * NMEL + j] : 1.f;
                out[(((long long)b * 7 + c) * T + tt) * NMEL + j] = fmaf(ch[c], is, -mu * is);
            }
        }
    }
    for (int b = 0; b < B; ++b)                                   // top_db clamp
        for (int c = 0; c < 4; ++c) {
            const float thr = gmax[b * 4 + c] - top_db;
            for (int t = 0; t < T; ++t)
                for (int j = 0; j < NMEL; ++j) {
                    const float mu = mean ? mean[c * NMEL + j] : 0.f, is = istd ? istd[c * NMEL + j] : 1.f;
                    float& o = out[(((long long)b * 7 + c) * T + t) * NMEL + j];
                    o = std::max(o, fmaf(thr, is, -mu * is));
                }
        }
    return 0;
}

// MIC format, CPU emulation of stages A / B / C (stage_c_mic): one clip -> unit phasors (T, 608, 4 channels, re/im) as
// float32 (the device packs them as half2) and the position -> FFT-bin map the GCC-PHAT B table is built from.
extern "C" int emu_fe2_mic_phasors(const int16_t* audio, long long N, const float* mel_dense, float dc_offset,
                                   float* phasors /* T x 608 x 8 */, int* bin_of_pos /* 608 */) {
    const int T = (int)(N / HOP);
    const int tpc = (T + TFR - 1) / TFR;
    static Tables tab;
    static MelPlan plan;
    bool ok;
    fill_tables(mel_dense, tab, plan, ok);
    if (!ok) return -1;
    int bop[PH_K];
    bins_of_phasor_pos(bop);
    for (int p = 0; p < PH_K; ++p) bin_of_pos[p] = bop[p];
    std::vector<unsigned char> smem(SmemLayout::total, 0);
    unsigned char* s_samp = smem.data() + SmemLayout::off_samples;
    unsigned char* s_x = smem.data() + SmemLayout::off_x;
    const unsigned char* s_tw = reinterpret_cast<const unsigned char*>(tab.tw75);
    const float dc0 = dc_offset * 300.0f, dc1 = -dc_offset * 150.0f;
    memset(phasors, 0, sizeof(float) * (size_t)T * PH_K * 8);
    for (int tile = 0; tile < tpc; ++tile) {
        const int t0 = tile * TFR, nf = std::min(TFR, T - t0);
        stage_copy(s_samp, audio, t0, nf, tab.col_perm);
        for (int tid = 0; tid < NT_AB; ++tid) { const int f = tid / 80, l = tid % 80; if (l < 75 && f < nf) stage_a(s_samp, tab.win, s_x, f, l, stage_a_const(l, tab.col_perm[l])); }
        for (int tid = 0; tid < NT_AB; ++tid) { const int f = tid / 80, u = tid % 80; if (f < nf) stage_b(s_x, f, u); }
        for (int f = 0; f < nf; ++f) {
            uint4* ph = reinterpret_cast<uint4*>(phasors + ((size_t)(t0 + f) * PH_K) * 8);   // 8 floats per position
            for (int task = 0; task < NREG; ++task) stage_c_mic<false>(s_x + f * X_STRIDE, s_tw, task, dc0, dc1, ph);
            for (int task = 0; task < NC0; ++task) stage_c_mic<true>(s_x + f * X_STRIDE, s_tw, task, dc0, dc1, ph);
        }
    }
    return 0;
}
