// CPU emulation of frontend_foa_kernel (TEST INFRASTRUCTURE): runs the same per-thread
// functions (frontend_core.cuh) thread by thread, phase by phase, over a fake shared memory,
// so the index logic can be checked against the oracle without a GPU.
#include <math.h>
#include <string.h>
#include <vector>
#include <algorithm>
#include "../../ad-yolo_b200/csrc/frontend_tables.h"
using namespace ady;

extern "C" int emu_features_foa(const int16_t* audio, int B, long long N, const float* mel_dense /*64x601*/,
                                const float* mean, const float* istd, float dc_offset, float top_db,
                                float* out /*B,7,T,64*/) {
  const int T = (int)(N / HOP);
  const int tpc = (T + TF - 1) / TF;
  std::vector<unsigned char> smem(SmemLayout::total, 0);
  uint32_t* s_samples = (uint32_t*)(smem.data() + SmemLayout::off_samples);
  float2* s_x1 = (float2*)(smem.data() + SmemLayout::off_x1);
  MelEntry* s_melent = (MelEntry*)(smem.data() + SmemLayout::off_melent);
  MelSchedule sch;
  if (!build_mel_schedule(mel_dense, sch)) return -1;
  for (size_t i = 0; i < sch.ent.size(); ++i) s_melent[i] = sch.ent[i];
  float wcs[50];
  for (int n2 = 0; n2 < 25; ++n2) { wcs[2 * n2] = (float)cos(2.0 * M_PI * n2 / 25.0); wcs[2 * n2 + 1] = (float)sin(2.0 * M_PI * n2 / 25.0); }
  const float dc0 = dc_offset * 300.0f, dc1 = -dc_offset * 150.0f;
  std::vector<float> gmax(B * 4, -INFINITY);
  for (int tile = 0; tile < B * tpc; ++tile) {
    const int b = tile / tpc, tb = tile % tpc, t0 = tb * TF, nf = std::min(TF, T - t0);
    // ---- copy (same element mapping as issue_tile_copy)
    for (int tid = 0; tid < COPY_THREADS; ++tid) {
      const int pair = (tid >> 4) & 1, i0 = (tid >> 5) * 16 + (tid & 15);
      const int16_t* clip = audio + (long long)b * N * 4 + pair * 2;
      for (int f = 0; f < nf; ++f) {
        const int t = t0 + f;
        uint32_t* dst = s_samples + splane_base(q_of_g(2 * f + pair)) + skew(i0);
        for (int i = 0; i < 15; ++i) {
          const int idx = i0 + 80 * i;
          long long m = t > 0 ? (long long)(t - 1) * HOP + idx : (idx < HOP ? HOP - idx : idx - HOP);
          memcpy(dst + 85 * i, clip + m * 4, 4);
        }
      }
    }
    // ---- stage 1
    for (int tid = 0; tid < NACT; ++tid) { int q = tid / 25, n2 = tid % 25; int f1 = q < TF ? q : q - TF; if (f1 < nf) stage1_task(s_samples, s_x1, q, n2, wcs[2 * n2], wcs[2 * n2 + 1]); }
    // ---- stage 2a (all threads, then "barrier")
    std::vector<Stage2Regs> R(NTHREADS);
    for (int tid = 0; tid < NTHREADS; ++tid) { int L = std::min(tid, NACT - 1); stage2a_task(s_x1, L / 50, (L % 50) >> 1, L & 1, dc0, dc1, 1.f, 1.f, R[tid]); }
    // ---- stage 2b: lane pairs exchange
    for (int tid = 0; tid < NTHREADS; tid += 2) {
      for (int k2 = 0; k2 < 25; ++k2) {
        SlotMine m[2]; SlotOut o[2];
        for (int s = 0; s < 2; ++s) { int L = std::min(tid + s, NACT - 1); int r = L & 1; slot_split(R[tid + s].P[k2], R[tid + s].Q[(25 - k2) % 25], r == 0 ? 1.0f : 1.0f / 3.0f, m[s], o[s]); }
        for (int s = 0; s < 2; ++s) {
          int L = std::min(tid + s, NACT - 1); int f2 = L / 50, t2 = (L % 50) >> 1, r2 = L & 1;
          float iva, ivb; slot_finish(m[s], o[s], o[1 - s], r2, iva, ivb);
          if (tid + s < NACT && f2 < nf) slot_store(s_x1 + v_base(f2) + 50 * t2 + r2, k2, m[s].P0, m[s].P1, iva, ivb);
        }
      }
    }
    // ---- mel: static schedule, lanes of a warp-task emulated one by one, pairs summed
    for (int code = 0; code < TF * 4; ++code) {
      const int f = code >> 2, wt = code & 3;
      if (f >= nf) continue;
      float accs[32][8];
      for (int lane = 0; lane < 32; ++lane)
        mel_task((const float4*)(s_x1 + v_base(f)), s_melent + sch.it0[wt] * 32 + lane, sch.nit[wt], accs[lane]);
      for (int lane = 0; lane < 32; lane += 2) {
        float acc[8];
        for (int c = 0; c < 8; ++c) acc[c] = accs[lane][c] + accs[lane + 1][c];
        const int j = 16 * wt + (lane >> 1);
        const long long tt = t0 + f;
        for (int c = 0; c < 4; ++c) {
          float db = power_to_db_unclamped(acc[c]);
          gmax[b * 4 + c] = std::max(gmax[b * 4 + c], db);
          float mu = mean ? mean[c * NMEL + j] : 0.f, is = istd ? istd[c * NMEL + j] : 1.f;
          out[(((long long)b * 7 + c) * T + tt) * NMEL + j] = fmaf(db, is, -mu * is);
        }
        for (int c = 4; c < 7; ++c) {
          float mu = mean ? mean[c * NMEL + j] : 0.f, is = istd ? istd[c * NMEL + j] : 1.f;
          out[(((long long)b * 7 + c) * T + tt) * NMEL + j] = fmaf(acc[c + 1], is, -mu * is);
        }
      }
    }
  }
  // ---- top_db clamp
  for (int b = 0; b < B; ++b) for (int c = 0; c < 4; ++c) {
    float thr = gmax[b * 4 + c] - top_db;
    for (int t = 0; t < T; ++t) for (int j = 0; j < NMEL; ++j) {
      float mu = mean ? mean[c * NMEL + j] : 0.f, is = istd ? istd[c * NMEL + j] : 1.f;
      float& o = out[(((long long)b * 7 + c) * T + t) * NMEL + j];
      o = std::max(o, fmaf(thr, is, -mu * is));
    }
  }
  return 0;
}
