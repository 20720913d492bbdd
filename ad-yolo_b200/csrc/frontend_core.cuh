// Per-thread building blocks of the fused SELD front-end kernel (FOA: STFT -> log-mel + IV).
//
// Replaces, per tile of TF feature frames of one clip (reference: /root/reference/src/
// datasets.py:252-292 get_stft_spectrogram / get_logmel_spectrogram /
// get_melscale_foa_intensity_vectors / get_feature, == utils/utility.py:142-215):
//
//   stage 1   : windowed samples (int16 pairs (W,Y) / (Z,X) packed as one complex signal)
//               -> 25 x DFT-48 per packed FFT            [thread = (fft, n2)]
//   stage 2a  : 25 x (2 x DFT-25) per packed FFT          [thread = (frame, residue pair t, role)]
//   stage 2b  : split packed spectra into the two real channels' spectra using the partner
//               bin N-k held by the SAME thread, |X|^2, FOA intensity (A/B lanes swap W and
//               energy with one shuffle step), write 7 values per bin to shared memory
//   mel       : sparse (CSR) mel projection, 10 log10, standardise, store (B,7,T,64)
//
// The 1200-point transform is a Good-Thomas prime-factor FFT (48 x 25, coprime): no twiddle
// factors between the stages, only compile-time index permutations.  All functions are
// ADY_HD so that tests/emu can run the identical code on the CPU.
#pragma once
#include <stdint.h>

#include "fft_codelets.cuh"

namespace ady {

// ---------------------------------------------------------------- fixed geometry (hyp_data_DCASE20xx.yaml:5-12)
constexpr int NFFT = 1200;
constexpr int HOP = 600;
constexpr int NBIN = 601;
constexpr int NMEL = 64;
constexpr int NCH_IN = 4;
constexpr int NCH_FOA = 7;

// ---------------------------------------------------------------- tiling
// ADY_TILE_FRAMES = 3: two CTAs of 5 warps per SM (150 of 160 lanes active in the FFT stages);
// ADY_TILE_FRAMES = 7: one CTA of 11 warps per SM (350 of 352 lanes), the four schedulers hold 3/3/3/2
// warps instead of 3/3/2/2, 231.5 KB of shared memory, 168 x 352 registers.
#ifndef ADY_TILE_FRAMES
#define ADY_TILE_FRAMES 3
#endif
constexpr int TF = ADY_TILE_FRAMES;        // frames per tile
constexpr int NFFT_TILE = 2 * TF;          // packed complex FFTs per tile
constexpr int NACT = NFFT_TILE * 25;       // threads with an FFT task (one per (packed FFT, n2) / (frame, residue pair, role))
constexpr int NTHREADS = ((NACT + 31) / 32) * 32;
constexpr int CTAS_PER_SM = TF == 3 ? 2 : 1;
constexpr int COPY_THREADS = 160;          // the staging map covers a frame with 5 warps x 16-sample columns x 15 steps
static_assert(TF == 3 || TF == 7, "mel_assign() holds a unit table for 3- and 7-frame tiles");

// sample planes: one per (frame, pair); word(idx) = idx + idx/16  (skew -> conflict-free
// stride-48 reads); 1275 words per plane == 51*25 so that lane L = 25*fft + n2 reads word
// 51*L + const.
constexpr int SPLANE = 1275;
constexpr int FS = 1256;                 // float2 per packed FFT in the exchange buffer (== 8 mod 16)
constexpr int XFRAME = 2 * FS + 25;      // float2 per frame block (2537 == 9 mod 16, see x1_base)
constexpr int NPOS = 625;                // V is stored in PFA position order p = 25*t + k2 (t, k2 < 25)
constexpr int MEL_MAXROWS = 56;          // rows of 32 lanes in the static mel schedule
constexpr int MEL_MAXNNZ = 1216;

// Stage 1 walks the packed FFTs in the lane-group order g = 0,2,4,1,3,5 (the three (W,Y) FFTs,
// then the three (Z,X) FFTs) and sample planes are stored in that order, so that lane L of the CTA
// reads sample word 51*L + const and writes exchange word 2*(25*k1) + 2*L' + const with the bank
// sequence continuing across lane groups:
//   x1_base(g+2) - x1_base(g) == 9 (mod 16 float2)   [25 float2 of the previous group]
//   x1_base(2f+1) - x1_base(2f) == 8 (mod 16 float2) [stage 2a: adjacent A/B lanes, disjoint banks]
// sample plane q starts at q*1275 words, the three (Z,X) planes shifted by 31 words so that the
// A and B planes of one frame are 16 banks apart (conflict-free 2 x 16-lane cp.async stores)
ADY_HD constexpr int splane_base(int q) { return q * SPLANE + (q >= TF ? 31 : 0); }
ADY_HD constexpr int g_of_q(int q) { return q < TF ? 2 * q : 2 * (q - TF) + 1; }
ADY_HD constexpr int q_of_g(int g) { return (g & 1) ? TF + (g >> 1) : (g >> 1); }
ADY_HD constexpr int x1_base(int g) { return (g >> 1) * XFRAME + (g & 1) * FS; }
ADY_HD constexpr int v_base(int f) { return ((f * XFRAME + 1) >> 1) << 1; }   // float2 units, 16-byte aligned

struct SmemLayout {
    static constexpr int off_samples = 0;                                   // uint32 [6][1275]
    static constexpr int off_x1 = ((off_samples + (NFFT_TILE * SPLANE + 31) * 4 + 15) / 16) * 16;  // float2 [TF][XFRAME]
    static constexpr int off_melent = ((off_x1 + TF * XFRAME * 8 + 15) / 16) * 16;  // MelEntry [56][32]
    static constexpr int off_melhdr = off_melent + MEL_MAXROWS * 32 * 8;     // int32 [8]: it0[4], nit[4]
    static constexpr int off_scale = off_melhdr + 8 * 4;                     // float2 [7][64]: (istd, -mean*istd)
    static constexpr int total = ((off_scale + NCH_FOA * NMEL * 8 + 15) / 16) * 16;
};
static_assert(SmemLayout::off_x1 % 16 == 0 && SmemLayout::off_melent % 16 == 0, "alignment");
static_assert(v_base(0) + 2 * NPOS * 2 <= XFRAME && x1_base(1) + 1200 <= XFRAME, "V / FFT data must fit in a frame block");

struct MelEntry {   // one non-zero of the mel matrix: V position of its FFT bin + weight
    int pos;
    float w;
};

// ---------------------------------------------------------------- helpers
ADY_HD int skew(int idx) { return idx + (idx >> 4); }

// order-preserving float <-> uint key (for atomicMax/Min on floats of any sign)
ADY_HD uint32_t f2key(float f) {
    union { float f; uint32_t u; } c; c.f = f;
    return (c.u & 0x80000000u) ? ~c.u : (c.u | 0x80000000u);
}
ADY_HD float key2f(uint32_t k) {
    union { float f; uint32_t u; } c;
    c.u = (k & 0x80000000u) ? (k & 0x7fffffffu) : ~k;
    return c.f;
}

// ---------------------------------------------------------------- stage 1
// Periodic Hann window at n = (25 n1 + 48 n2) mod 1200, scaled by 2^-16 (int16 -> [-1,1) and the
// 1/2 of the channel split):  w = 2^-16 (0.5 - 0.5 cos(2 pi n1/48 + 2 pi n2/25))
//                               = W0 + CA[n1] * cos(2 pi n2/25) + SA[n1] * sin(2 pi n2/25)
// CA/SA are compile-time immediates, (cB, sB) two per-thread registers: two FMAs replace a
// shared-memory table lookup (shared-memory bandwidth is the scarcer resource in this kernel).
// thread (q = lane group 0..5 -> packed fft g = g_of_q(q), n2 = 0..24).  samples: uint32 words
// (lo int16 = re channel, hi int16 = im channel), plane q.
ADY_HD void stage1_task(const uint32_t* __restrict__ samples, float2* __restrict__ x1, int q, int n2,
                        float cB, float sB) {
    constexpr float WIN_W0 = 0.5f / 65536.0f;
    constexpr float WIN_CA[48] = {-7.6293945312e-06f, -7.5641240034e-06f, -7.3694292167e-06f, -7.0486414529e-06f, -6.6072494796e-06f, -6.0528056358e-06f, -5.3947966094e-06f, -4.6444811173e-06f, -3.8146972656e-06f, -2.9196428861e-06f, -1.9746326073e-06f, -9.9583581711e-07f, -4.6716567961e-22f, 9.9583581711e-07f, 1.9746326073e-06f, 2.9196428861e-06f, 3.8146972656e-06f, 4.6444811173e-06f, 5.3947966094e-06f, 6.0528056358e-06f, 6.6072494796e-06f, 7.0486414529e-06f, 7.3694292167e-06f, 7.5641240034e-06f, 7.6293945312e-06f, 7.5641240034e-06f, 7.3694292167e-06f, 7.0486414529e-06f, 6.6072494796e-06f, 6.0528056358e-06f, 5.3947966094e-06f, 4.6444811173e-06f, 3.8146972656e-06f, 2.9196428861e-06f, 1.9746326073e-06f, 9.9583581711e-07f, 1.4014970388e-21f, -9.9583581711e-07f, -1.9746326073e-06f, -2.9196428861e-06f, -3.8146972656e-06f, -4.6444811173e-06f, -5.3947966094e-06f, -6.0528056358e-06f, -6.6072494796e-06f, -7.0486414529e-06f, -7.3694292167e-06f, -7.5641240034e-06f};
    constexpr float WIN_SA[48] = {0.0000000000e+00f, 9.9583581711e-07f, 1.9746326073e-06f, 2.9196428861e-06f, 3.8146972656e-06f, 4.6444811173e-06f, 5.3947966094e-06f, 6.0528056358e-06f, 6.6072494796e-06f, 7.0486414529e-06f, 7.3694292167e-06f, 7.5641240034e-06f, 7.6293945312e-06f, 7.5641240034e-06f, 7.3694292167e-06f, 7.0486414529e-06f, 6.6072494796e-06f, 6.0528056358e-06f, 5.3947966094e-06f, 4.6444811173e-06f, 3.8146972656e-06f, 2.9196428861e-06f, 1.9746326073e-06f, 9.9583581711e-07f, 9.3433135921e-22f, -9.9583581711e-07f, -1.9746326073e-06f, -2.9196428861e-06f, -3.8146972656e-06f, -4.6444811173e-06f, -5.3947966094e-06f, -6.0528056358e-06f, -6.6072494796e-06f, -7.0486414529e-06f, -7.3694292167e-06f, -7.5641240034e-06f, -7.6293945312e-06f, -7.5641240034e-06f, -7.3694292167e-06f, -7.0486414529e-06f, -6.6072494796e-06f, -6.0528056358e-06f, -5.3947966094e-06f, -4.6444811173e-06f, -3.8146972656e-06f, -2.9196428861e-06f, -1.9746326073e-06f, -9.9583581711e-07f};
    const uint32_t* sp = samples + splane_base(q) + 51 * n2;
    const int thr = 1200 - 48 * n2;  // wrap when 25*n1 >= thr
    cx<float> x[48];
#pragma unroll
    for (int n1 = 0; n1 < 48; ++n1) {
        const int c = 25 * n1;
        const int off = c + (c >> 4);
        const uint32_t word = sp[off - (c >= thr ? SPLANE : 0)];
        const float w = WIN_W0 + WIN_CA[n1] * cB + WIN_SA[n1] * sB;
        const float lo = (float)(int16_t)(word & 0xffffu);
        const float hi = (float)(int16_t)(word >> 16);
        x[n1] = {lo * w, hi * w};
    }
    dft48(x);
    float2* xo = x1 + x1_base(g_of_q(q)) + n2;
#pragma unroll
    for (int k1 = 0; k1 < 48; ++k1) xo[k1 * 25] = make_float2(x[k1].re, x[k1].im);
}

// ---------------------------------------------------------------- rotation augmentation
// The 16 FOA channel sign / swap combinations of the reference (utils/augmentations.py:46-70),
// packed per combination as bit0 = Y negated, bit1 = Z negated, bit2 = X negated, bit3 = X<->Y swap.
// A sign flip of a real channel flips the sign of its spectrum: powers are unchanged, the
// intensity component of that channel changes sign, and the "+1e-8" DC term (added after the
// augmentation in datasets.py:146-147) keeps its sign; the swap is an output channel permutation.
ADY_HD constexpr unsigned rot_bits(int comb) {
    // yzx_weight / xy_swap columns of the table, in order
    constexpr unsigned T[16] = {0x0, 0x2, 0x1, 0x3, 0x5, 0x7, 0x4, 0x6, 0x9, 0xB, 0x8, 0xA, 0xC, 0xE, 0xD, 0xF};
    return T[comb & 15];
}
ADY_HD unsigned rot_bits_rt(int comb) {
    // same table from a 64-bit immediate (no constant-memory array in device code)
    return (unsigned)((0xFDECA8B964753120ull >> (4 * (comb & 15))) & 15ull);
}

// ---------------------------------------------------------------- stage 2a
struct Stage2Regs {
    cx<float> P[25], Q[25];
};

// thread (frame f, residue pair t = 0..24 -> rows (t, (48-t)%48), role r: 0 = (W,Y) fft, 1 = (Z,X) fft)
// dc: analytic contribution of the "+1e-8" DC offset of datasets.py:147 (pre-scaled like the
// window): bin 0 gets +dc0 on both packed components, bins +-1 get dc1.
// (sre, sim) = +-1: sign of the real / imaginary channel of this packed FFT under the rotation
// augmentation; the DC term of a sign-flipped channel enters with the opposite sign so that
// sign * (FFT(x) + sign*dc) == FFT(sign*x) + dc.
ADY_HD void stage2a_task(const float2* __restrict__ x1, int f, int t, int r, float dc0, float dc1,
                         float sre, float sim, Stage2Regs& s) {
    const int g = 2 * f + r;
    const float2* pa = x1 + x1_base(g) + t * 25;
    const float2* pb = x1 + x1_base(g) + ((48 - t) % 48) * 25;
#pragma unroll
    for (int n2 = 0; n2 < 25; ++n2) {
        float2 a = pa[n2], b = pb[n2];
        s.P[n2] = {a.x, a.y};
        s.Q[n2] = {b.x, b.y};
    }
    dft25(s.P);
    dft25(s.Q);
    if (t == 0) {  // bin 0 <-> (k1,k2) = (0,0); P and Q are the same row here
        s.P[0].re += sre * dc0; s.P[0].im += sim * dc0;
        s.Q[0].re += sre * dc0; s.Q[0].im += sim * dc0;
    }
    if (t == 1) {  // bin 1 <-> (1,1) in P; bin 1199 <-> (47,24) in Q
        s.P[1].re += sre * dc1; s.P[1].im += sim * dc1;
        s.Q[24].re += sre * dc1; s.Q[24].im += sim * dc1;
    }
}

// ---------------------------------------------------------------- stage 2b (per slot)
struct SlotOut {          // what a lane sends to its partner lane (role A <-> role B)
    float s0re, s0im, e;
};
struct SlotMine {
    cx<float> S0, S1;     // role A: (W, Y)   role B: (Z, X)
    float P0, P1;
};

ADY_HD void slot_split(const cx<float> a, const cx<float> bq, float c0, SlotMine& m, SlotOut& o) {
    // a = A[k], bq = A[N-k] of the packed FFT (already halved by the window scale):
    // S0 = (a + conj bq), S1 = (a - conj bq)/i
    m.S0 = {a.re + bq.re, a.im - bq.im};
    m.S1 = {a.im + bq.im, bq.re - a.re};
    m.P0 = m.S0.re * m.S0.re + m.S0.im * m.S0.im;
    m.P1 = m.S1.re * m.S1.re + m.S1.im * m.S1.im;
    o.s0re = m.S0.re;
    o.s0im = m.S0.im;
    o.e = c0 * m.P0 + m.P1 * (1.0f / 3.0f);
}

// E = eps + |W|^2 + (|Y|^2+|Z|^2+|X|^2)/3  (datasets.py:272);  I = Re(conj(W) X_c) / E (:271,274)
ADY_HD void slot_finish(const SlotMine& m, const SlotOut& mine, const SlotOut& other, int r,
                        float& iva, float& ivb) {
    const float wre = r == 0 ? m.S0.re : other.s0re;
    const float wim = r == 0 ? m.S0.im : other.s0im;
    const float E = 1e-8f + (mine.e + other.e);
#if defined(__CUDA_ARCH__)
    const float rE = __fdividef(1.0f, E);   // MUFU.RCP (2 ulp) - far inside the 1e-3 tolerance
#else
    const float rE = 1.0f / E;
#endif
    iva = (wre * m.S0.re + wim * m.S0.im) * rE;   // role B: IV_Z        (role A: unused)
    ivb = (wre * m.S1.re + wim * m.S1.im) * rE;   // role A: IV_Y, role B: IV_X
}

// bin index (folded to 0..600) of PFA output (k1 = t, k2)
ADY_HD int slot_bin(int kt /* = 625*t % 1200 */, int k2) {
    int k = kt + (576 * k2) % 1200;
    k = k >= 1200 ? k - 1200 : k;
    return k > 600 ? 1200 - k : k;
}

// V layout per frame (float4 units, PFA position order p = 25 t + k2):
//   V4a[p] = (|W|^2,|Y|^2,|Z|^2,|X|^2),  V4b[p] = (junk, IV_Y, IV_Z, IV_X)
// vbase = frame base + 50 t + r (float2 units): every slot offset is a compile-time constant.
ADY_HD void slot_store(float2* __restrict__ vbase, int k2, float P0, float P1, float iva, float ivb) {
    vbase[2 * k2] = make_float2(P0, P1);
    vbase[2 * NPOS + 2 * k2] = make_float2(iva, ivb);
}

// ---------------------------------------------------------------- mel phase
// warp-task wt (16 filters x 2 parts = 32 lanes): lane walks its column of the static schedule
// ent[it0 + it][lane]; every lane of the warp runs nit iterations (zero-weight padding).
template <bool WITH_B = true>
ADY_HD void mel_task(const float4* __restrict__ vframe4, const MelEntry* __restrict__ ent_col, int nit,
                     float (&acc)[8]) {
#pragma unroll
    for (int c = 0; c < 8; ++c) acc[c] = 0.f;
#pragma unroll 2
    for (int it = 0; it < nit; ++it) {
        const MelEntry e = ent_col[it * 32];
        const float4 a = vframe4[e.pos];
        acc[0] += e.w * a.x; acc[1] += e.w * a.y; acc[2] += e.w * a.z; acc[3] += e.w * a.w;
        if (WITH_B) {
            const float4 b = vframe4[NPOS + e.pos];
            acc[4] += e.w * b.x; acc[5] += e.w * b.y; acc[6] += e.w * b.z; acc[7] += e.w * b.w;
        }
    }
}

// Balanced assignment of the 12 (frame, warp-task) pairs of a tile to the 5 warps (iteration
// counts ~ 3 / 6 / 14 / 31): code = 4*frame + warp-task, -1 = none.  Makespan 34 vs 51 round-robin.
// 15 nibbles [warp][slot], 15 = none:  w0: (f0,wt3) (f0,wt0) | w1: (f1,wt3) (f1,wt0) | w2: (f2,wt3) (f2,wt0)
//                                      | w3: (f0,wt2) (f1,wt2) (f0,wt1) | w4: (f2,wt2) (f1,wt1) (f2,wt1)
constexpr unsigned long long MEL_ASSIGN_PACK = 0x95a162f8bf47f03ull;
// 7-frame tile, 11 warps, 28 units (total 378 iterations, ideal 34.4 per warp; makespan 37):
//   w0: (f0,wt3) (f0,wt1) | w1..w6: (fi,wt3) (fi,wt0) | w7: (f0,wt2) (f1,wt2) (f1,wt1) | w8: (f2,wt2) (f3,wt2) (f2,wt1)
//   w9: (f4,wt2) (f5,wt2) (f3,wt1) | w10: (f6,wt2) (f4,wt1) (f5,wt1) (f6,wt1) (f0,wt0)
constexpr int MEL_SLOTS = TF == 3 ? 3 : 5;
ADY_HD int mel_assign(int warp, int slot) {
    if (TF == 3) {
        const int v = (int)((MEL_ASSIGN_PACK >> (4 * (warp * 3 + slot))) & 15ull);
        return v == 15 ? -1 : v;
    }
    // (closed form instead of a table: a local array would live in local memory)
    if (warp < 7) return slot == 0 ? 4 * warp + 3 : (slot == 1 ? (warp == 0 ? 1 : 4 * warp) : -1);
    if (warp < 10) {
        const int f = 2 * (warp - 7);
        return slot == 0 ? 4 * f + 2 : (slot == 1 ? 4 * (f + 1) + 2 : (slot == 2 ? 4 * (warp - 6) + 1 : -1));
    }
    if (warp == 10) return slot == 0 ? 26 : (slot == 1 ? 17 : (slot == 2 ? 21 : (slot == 3 ? 25 : 0)));
    return -1;
}

// librosa.power_to_db(ref=1, amin=1e-10) before the top_db clamp (datasets.py:265)
ADY_HD float power_to_db_unclamped(float s) {
#if defined(__CUDA_ARCH__)
    float l;   // 10 log10(s) = 3.0103 log2(s); argument >= 1e-10 is never denormal -> bare MUFU.LG2
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l) : "f"(fmaxf(s, 1e-10f)));
    return 3.0102999566398120f * l;
#else
    return 10.0f * __builtin_log10f(s > 1e-10f ? s : 1e-10f);
#endif
}

}  // namespace ady
