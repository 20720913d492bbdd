// Front end, second generation ("fe2"): per-thread building blocks of the fused SELD feature kernel.
//
// Replaces, per tile of 2 feature frames of one clip, the reference's
//   get_stft_spectrogram / get_logmel_spectrogram / get_melscale_foa_intensity_vectors / get_feature
//   (/root/reference/src/datasets.py:252-292 == src/utils/utility.py:142-215) and the normalisation
//   of datasets.py:147.
//
// What changed against the round-1 kernel (frontend_core.cuh, kept for the compatibility paths):
//   * every thread carries BOTH packed complex FFTs of a frame -- (W + iY) and (Z + iX) -- as the two
//     halves of sm_100a's packed FP32 registers: one FADD2 / FMUL2 / FFMA2 does the same butterfly
//     step of both transforms, so the FFT costs half the issue slots, and after the last stage a
//     thread owns all four channel spectra of its bins: the intensity vectors need no lane exchange;
//   * 1200 = 16 x 75 (Good-Thomas, no twiddles) with 75 = 15 x 5 (Cooley-Tukey, 4 twiddles per DFT-5)
//     and 15 = 3 x 5 (Good-Thomas): three register-resident stages of 16 / 15 / 5 points instead of
//     two of 48 / 25, i.e. ~100 registers per thread instead of 168 and 15-20 resident warps per SM;
//   * the exchange buffer holds one 16-byte element {re(A), re(B), im(A), im(B)} per point: every
//     shared-memory access of the FFT stages is a 128-bit one, all conflict-free by construction;
//   * stages B and C work in place, so one 19 KB buffer per frame serves X1, X2 and V;
//   * the mel projection handles both frames of a tile per schedule entry and is balanced as 156
//     lane-jobs of <= 9 non-zeros (one per thread) instead of 4 warp-tasks of 3..31 iterations.
//
// Everything is ADY_HD so that tests/emu runs the identical index logic on the CPU.
#pragma once
#include <stdint.h>
#include <string.h>

#include "fft_codelets.cuh"

#if !defined(__CUDACC__)
struct uint4 { unsigned x, y, z, w; };      // plain C++ build (tests/emu)
#endif

namespace ady {
namespace fe2 {

// ---------------------------------------------------------------- packed pair of floats
#if defined(__CUDA_ARCH__)
struct f2 {
    unsigned long long v;
};
__device__ __forceinline__ f2 mk2(float lo, float hi) { f2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r.v) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ f2 dup2(float a) { return mk2(a, a); }
__device__ __forceinline__ float lo2(f2 a) { float l; asm("{ .reg .b32 t; mov.b64 {%0, t}, %1; }" : "=f"(l) : "l"(a.v)); return l; }
__device__ __forceinline__ float hi2(f2 a) { float h; asm("{ .reg .b32 t; mov.b64 {t, %0}, %1; }" : "=f"(h) : "l"(a.v)); return h; }
__device__ __forceinline__ f2 operator+(f2 a, f2 b) { f2 r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v)); return r; }
__device__ __forceinline__ f2 operator-(f2 a, f2 b) { f2 r; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v)); return r; }
__device__ __forceinline__ f2 operator*(f2 a, f2 b) { f2 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v)); return r; }
__device__ __forceinline__ f2 fma2(f2 a, f2 b, f2 c) { f2 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r.v) : "l"(a.v), "l"(b.v), "l"(c.v)); return r; }
#else
struct f2 {
    float lo, hi;
};
static inline f2 mk2(float lo, float hi) { return {lo, hi}; }
static inline f2 dup2(float a) { return {a, a}; }
static inline float lo2(f2 a) { return a.lo; }
static inline float hi2(f2 a) { return a.hi; }
static inline f2 operator+(f2 a, f2 b) { return {a.lo + b.lo, a.hi + b.hi}; }
static inline f2 operator-(f2 a, f2 b) { return {a.lo - b.lo, a.hi - b.hi}; }
static inline f2 operator*(f2 a, f2 b) { return {a.lo * b.lo, a.hi * b.hi}; }
static inline f2 fma2(f2 a, f2 b, f2 c) { return {__builtin_fmaf(a.lo, b.lo, c.lo), __builtin_fmaf(a.hi, b.hi, c.hi)}; }
#endif

struct c2 {   // one complex value of each of the two packed FFTs: lo half = FFT A (W + iY), hi half = FFT B (Z + iX)
    f2 re, im;
};
#define ADY_K2(x) dup2((float)(x))

// ---------------------------------------------------------------- codelets on c2 (negation-free: add / sub / mul / fma only)
ADY_HD void p_dft3(c2& x0, c2& x1, c2& x2) {
    const f2 mh = ADY_K2(-0.5), s = ADY_K2(0.86602540378443864676), ms = ADY_K2(-0.86602540378443864676);
    const c2 t = {x1.re + x2.re, x1.im + x2.im}, d = {x1.re - x2.re, x1.im - x2.im};
    const c2 m = {fma2(t.re, mh, x0.re), fma2(t.im, mh, x0.im)};
    x0 = {x0.re + t.re, x0.im + t.im};
    x1 = {fma2(d.im, s, m.re), fma2(d.re, ms, m.im)};    // m - i s d
    x2 = {fma2(d.im, ms, m.re), fma2(d.re, s, m.im)};    // m + i s d
}

ADY_HD void p_dft4(c2& x0, c2& x1, c2& x2, c2& x3) {
    const c2 a = {x0.re + x2.re, x0.im + x2.im}, b = {x0.re - x2.re, x0.im - x2.im};
    const c2 c = {x1.re + x3.re, x1.im + x3.im}, d = {x1.re - x3.re, x1.im - x3.im};
    x0 = {a.re + c.re, a.im + c.im};
    x2 = {a.re - c.re, a.im - c.im};
    x1 = {b.re + d.im, b.im - d.re};   // b - i d
    x3 = {b.re - d.im, b.im + d.re};   // b + i d
}

ADY_HD void p_dft5(c2& x0, c2& x1, c2& x2, c2& x3, c2& x4) {
    const f2 c1 = ADY_K2(0.30901699437494742410), c2_ = ADY_K2(-0.80901699437494742410);
    const f2 s1 = ADY_K2(0.95105651629515357212), s2 = ADY_K2(0.58778525229247312917), ms1 = ADY_K2(-0.95105651629515357212);
    const c2 t1 = {x1.re + x4.re, x1.im + x4.im}, t2 = {x2.re + x3.re, x2.im + x3.im};
    const c2 t3 = {x1.re - x4.re, x1.im - x4.im}, t4 = {x2.re - x3.re, x2.im - x3.im};
    const c2 a1 = {fma2(t2.re, c2_, fma2(t1.re, c1, x0.re)), fma2(t2.im, c2_, fma2(t1.im, c1, x0.im))};
    const c2 a2 = {fma2(t2.re, c1, fma2(t1.re, c2_, x0.re)), fma2(t2.im, c1, fma2(t1.im, c2_, x0.im))};
    const c2 b1 = {fma2(t4.re, s2, t3.re * s1), fma2(t4.im, s2, t3.im * s1)};
    const c2 b2 = {fma2(t4.re, ms1, t3.re * s2), fma2(t4.im, ms1, t3.im * s2)};
    x0 = {x0.re + (t1.re + t2.re), x0.im + (t1.im + t2.im)};
    x1 = {a1.re + b1.im, a1.im - b1.re};   // a1 - i b1
    x4 = {a1.re - b1.im, a1.im + b1.re};
    x2 = {a2.re + b2.im, a2.im - b2.re};
    x3 = {a2.re - b2.im, a2.im + b2.re};
}

// multiply by W16^E = exp(-2 pi i E / 16)
template <int E> ADY_HD c2 p_tw16(c2 v) {
    const f2 c8 = ADY_K2(0.92387953251128675613), s8 = ADY_K2(0.38268343236508977173), r2 = ADY_K2(0.70710678118654752440);
    const f2 nc8 = ADY_K2(-0.92387953251128675613), ns8 = ADY_K2(-0.38268343236508977173), nr2 = ADY_K2(-0.70710678118654752440);
    if constexpr (E == 0) return v;
    else if constexpr (E == 1) return {fma2(v.im, s8, v.re * c8), fma2(v.re, ns8, v.im * c8)};    // (c8 - i s8) v
    else if constexpr (E == 2) return {(v.re + v.im) * r2, (v.im - v.re) * r2};
    else if constexpr (E == 3) return {fma2(v.im, c8, v.re * s8), fma2(v.re, nc8, v.im * s8)};    // (s8 - i c8) v
    else if constexpr (E == 6) return {(v.im - v.re) * r2, (v.re + v.im) * nr2};
    else { static_assert(E == 9, "unsupported W16 power"); return {fma2(v.im, ns8, v.re * nc8), fma2(v.re, s8, v.im * nc8)}; }  // (-c8 + i s8) v
}

// DFT-16 = 4 x 4 Cooley-Tukey.  Natural order in; the result X[c + 4 d] is left in slot 4 c + d
// (callers use dft16_slot_k() instead of paying for the transpose).
ADY_HD void p_dft16(c2 (&x)[16]) {
#pragma unroll
    for (int b = 0; b < 4; ++b) p_dft4(x[b], x[4 + b], x[8 + b], x[12 + b]);
    x[5] = p_tw16<1>(x[5]);   x[6] = p_tw16<2>(x[6]);   x[7] = p_tw16<3>(x[7]);
    x[9] = p_tw16<2>(x[9]);
    { const c2 v = x[10]; x[10] = {v.im, ADY_K2(0.0) - v.re}; }                                   // W16^4 = -i
    x[11] = p_tw16<6>(x[11]);
    x[13] = p_tw16<3>(x[13]); x[14] = p_tw16<6>(x[14]); x[15] = p_tw16<9>(x[15]);
#pragma unroll
    for (int c = 0; c < 4; ++c) p_dft4(x[4 * c], x[4 * c + 1], x[4 * c + 2], x[4 * c + 3]);
}
ADY_HD constexpr int dft16_slot_k(int s) { return (s >> 2) + 4 * (s & 3); }

// DFT-15 = 3 x 5 Good-Thomas: input a = (5 a3 + 3 a5) mod 15, output c = (10 c3 + 6 c5) mod 15; natural in / out.
ADY_HD void p_dft15(c2 (&x)[15]) {
    c2 y[3][5];
#pragma unroll
    for (int a5 = 0; a5 < 5; ++a5) {
        c2 p0 = x[(3 * a5) % 15], p1 = x[(5 + 3 * a5) % 15], p2 = x[(10 + 3 * a5) % 15];
        p_dft3(p0, p1, p2);
        y[0][a5] = p0; y[1][a5] = p1; y[2][a5] = p2;
    }
#pragma unroll
    for (int c3 = 0; c3 < 3; ++c3) {
        p_dft5(y[c3][0], y[c3][1], y[c3][2], y[c3][3], y[c3][4]);
#pragma unroll
        for (int c5 = 0; c5 < 5; ++c5) x[(10 * c3 + 6 * c5) % 15] = y[c3][c5];
    }
}

// order-preserving float <-> uint key (atomicMax / atomicMin on floats of any sign)
ADY_HD uint32_t f2key(float f) {
    union { float f; uint32_t u; } c; c.f = f;
    return (c.u & 0x80000000u) ? ~c.u : (c.u | 0x80000000u);
}

// RotationAug combination number (utils/augmentations.py:46-70) -> bit0 = Y negated, bit1 = Z negated, bit2 = X negated,
// bit3 = X <-> Y swap (the table of frontend_core.cuh::rot_bits as a 64-bit immediate)
ADY_HD unsigned rot_bits_rt(int comb) { return (unsigned)((0xFDECA8B964753120ull >> (4 * (comb & 15))) & 15ull); }

// ---------------------------------------------------------------- geometry
constexpr int NFFT = 1200, HOP = 600, NBIN = 601, NMEL = 64;
constexpr int TFR = 2;                       // frames per tile
#ifndef ADY_FE2_NT
#define ADY_FE2_NT 160
#endif
constexpr int NT = ADY_FE2_NT;               // threads per CTA: 160 (5 warps) or 192 (6 warps: shorter mel jobs, 18 warps / SM)
constexpr int NT_AB = 160;                   // stages A and B and the staging copy: 80 lanes per frame
constexpr int ROWP = 80;                     // pitch of a 75-sample row of the staged audio, in samples (8 bytes each)
constexpr int WIN_P = 64;                    // pitch of a window-table row: a multiple of 32, so the bank of an entry is its column
constexpr int WIN_DIRECT = 44;               // lanes 0..43 own a column; lanes 44..74 read the mirror image (columns 31..1)
constexpr int NROWS = 8 * (TFR + 1);         // 3 hops = 24 rows
constexpr int SAMP_BYTES = NROWS * ROWP * 8; // 15 360
constexpr int XSLOTS = 1210;                 // 16-byte slots per frame: 1200 points + 2 x 5 (Vb of the self-mirror tasks)
constexpr int X_BYTES = XSLOTS * 16;
#ifdef ADY_FE2_ALIAS_X
constexpr int X_STRIDE = 0;
#else
constexpr int X_STRIDE = X_BYTES;            // byte distance between the two frame buffers of a tile
#endif
constexpr int MEL_L = NT >= 192 ? 7 : 9;     // schedule rows: non-zeros per lane-job
constexpr int NJOBS = NT >= 192 ? 192 : 160; // lane-jobs of the mel projection (191 / 156 used), one per thread
constexpr int NREG = 112;                    // regular pair-tasks per frame in stage C (c = 1..7, k16 = 0..15)
constexpr int NC0 = 9;                       // c = 0 pair-tasks per frame (k16 = 1..7 and the two self-mirror rows)
constexpr int REC_MAXJOBS = 9;               // <= 9 jobs per mel filter (61 non-zeros / 7)
constexpr int REC_PLANE = NJOBS * 16;        // partial record of job q: slot q of 4 planes (frame 0 | frame 1) x (powers | intensities)

struct MelEnt {            // one non-zero: byte offsets of the bin's two V records inside a frame's buffer + weight
    uint16_t offa, offb;
    float w;
};

struct Tables {            // device-resident constants of the fe2 kernel (built on the host, tables.cu)
    float win[16 * WIN_P];         // stage A: 2^-16 x periodic Hann at sample r + 75 m of lane l < 44 (r = 16 l mod 75), [m][l]; lanes
                                   // l >= 44 read the mirror image w[n] = w[1200 - n]: entry [15 - m][75 - l]  (stage_a_const)
    float tw75[15 * 4 * 2];        // stage C: W75^{b c} as (wr, wi) for c = 0..14, b = 1..4
    alignas(16) MelEnt ent[MEL_L * NJOBS];     // [row][job]  (copied / read as 8-byte words)
    // lane-jobs, chunk-major: the i-th job of mel j (i < mel_njobs[j]) is lane / record slot j + rec_off[i] -- filters with
    // more than i jobs form a suffix of the mel axis, so the 8 epilogue threads of a quarter-warp read 8 consecutive records
    int16_t rec_off[REC_MAXJOBS + 1];
    uint8_t mel_njobs[NMEL];
    uint8_t job_mel[NJOBS];        // (diagnostics / emulation)
    uint8_t col_perm[80];          // staged-audio column of the samples stage-A lane l consumes (see "staging map")
};

// Occupancy.  GROUPS independent 5-warp tile pipelines ("groups", each with its own staging + frame buffers and its own
// named barrier) share one CTA and with it ONE copy of the read-only tables in shared memory.  Default: 1 CTA per SM with
// 4 groups = 20 warps per SM.  The alternative spelling of the same occupancy, 4 CTAs per SM with one group each
// (-DADY_FE2_GROUPS=1 -DADY_FE2_CTAS=4), has no room for four table copies and reads them through an L1 that the 4 x 55 KB
// of shared memory have shrunk to a few KB, i.e. from L2 (measured: 0.446 ms against 0.41 ms).
#ifndef ADY_FE2_GROUPS
#define ADY_FE2_GROUPS 4
#endif
#ifndef ADY_FE2_CTAS
#define ADY_FE2_CTAS 1
#endif
constexpr int GROUPS = ADY_FE2_GROUPS;
constexpr bool TABLES_IN_SMEM = GROUPS > 1 || ADY_FE2_CTAS < 4;

struct SmemLayout {
    static constexpr int smem_max = 232448;                                // 227 KB opt-in limit per CTA
    // shared by the groups of a CTA
    static constexpr int off_meljobs = 0;                                 // uint8 mel_njobs[64] | int32 record byte offset of chunk i [16]
    static constexpr int off_ent = off_meljobs + 2 * NMEL;                // MelEnt [MEL_L][NJOBS]            (TABLES_IN_SMEM)
    static constexpr int off_win = off_ent + (TABLES_IN_SMEM ? MEL_L * NJOBS * 8 : 0);       // float [16][WIN_P]
    static constexpr int off_scale = off_win + (TABLES_IN_SMEM ? 16 * WIN_P * 4 : 0);        // float2 [7][64]: (istd, -mean*istd)
    static constexpr int group_bytes = SAMP_BYTES + TFR * X_BYTES;        // staged audio | TFR frame buffers
    // the standardisation constants join the tables when they fit (they do not next to 4 groups: read through L1 / L2)
    static constexpr bool scale_in_smem = TABLES_IN_SMEM && off_scale + 7 * NMEL * 8 + GROUPS * group_bytes <= smem_max / ADY_FE2_CTAS - 1024;
    static constexpr int off_group = ((off_scale + (scale_in_smem ? 7 * NMEL * 8 : 0) + 15) / 16) * 16;
    static constexpr int off_samples = off_group;                         // (group 0)
    static constexpr int off_x = off_group + SAMP_BYTES;
    static constexpr int total = off_group + GROUPS * group_bytes;
};
static_assert(4 * REC_PLANE <= TFR * X_BYTES, "partial records alias the frame buffers");
static_assert(SmemLayout::off_group % 16 == 0 && SmemLayout::off_ent % 16 == 0 && SmemLayout::off_win % 16 == 0 &&
              SmemLayout::group_bytes % 16 == 0, "alignment");
static_assert(SmemLayout::total <= SmemLayout::smem_max, "shared memory");

// ---------------------------------------------------------------- small memory helpers (same code on host and device)
ADY_HD void ld_c2(const unsigned char* p, c2& v) {
    float4 q;
#if defined(__CUDA_ARCH__)
    q = *reinterpret_cast<const float4*>(p);
#else
    memcpy(&q, p, 16);
#endif
    v.re = mk2(q.x, q.y);
    v.im = mk2(q.z, q.w);
}
ADY_HD void st_c2(unsigned char* p, const c2& v) {
    float4 q = make_float4(lo2(v.re), hi2(v.re), lo2(v.im), hi2(v.im));
#if defined(__CUDA_ARCH__)
    *reinterpret_cast<float4*>(p) = q;
#else
    memcpy(p, &q, 16);
#endif
}
ADY_HD void st_f4(unsigned char* p, float a, float b, float c, float d) {
    float4 q = make_float4(a, b, c, d);
#if defined(__CUDA_ARCH__)
    *reinterpret_cast<float4*>(p) = q;
#else
    memcpy(p, &q, 16);
#endif
}

// ---------------------------------------------------------------- staging map
// Tile sample N (0 .. 1799; clip sample 600 (t0 - 1) + N, mirrored about 0 for the reflect padding of
// librosa.stft(center=True)) lives at row N / 75, column col_perm[61 (N mod 75) mod 75] of the staged buffer: the
// stage-A lane l (task r = 16 l mod 75, i.e. samples n == r mod 75) finds all its samples in one column,
// col_perm[l].  The row pitch is 80 samples = 640 bytes == 0 mod 128, so the 8-byte bank pair of an access is its
// column mod 16 whatever row it is in.  col_perm (fe2_tables.h) is chosen so that BOTH sides are conflict-free:
// the 16 lanes of every half-warp of stage A read 16 different bank pairs, and so do the 16 threads of every
// half-warp of the copy (thread rem writes the column of lane 61 rem mod 75; with col_perm = identity those
// columns step by -14 mod 75 and a half-warp's cp.async needed 1.8 wavefronts instead of 1).
ADY_HD constexpr int stage_col(int rem) { return (61 * rem) % 75; }   // the stage-A lane that consumes residue rem

// ---------------------------------------------------------------- stage A: window + DFT-16 over n16 (75 tasks / frame)
// PFA input map n = (75 n16 + 16 n75) mod 1200.  Lane l handles n75 = l, i.e. the samples n == r (mod 75),
// r = 16 l mod 75; sample n16 of the task is n = r + 75 ((n16 + c_r) & 15) with c_r = (16 - 3 r) & 15.
// Window: periodic Hann scaled by 2^-16 (int16 -> [-1, 1) and the 1/2 of the channel split), read from a table of
// correctly rounded values (half a table: the window is symmetric, see stage_a_const).  (Evaluating 0.5 - 0.5 cos(theta_l + 2 pi n16 / 16) from a per-lane
// (cos, sin) pair costs two FMAs instead of one LDS but leaves a fixed ~2^-41 error pattern in the window whose
// leakage shows up in the 'harsh' fixture: 2.4e-3 instead of 1.5e-3 on the standardised intensity channels.)
struct StageAConst {
    int col_off;     // byte offset of column col_perm[l] in a row of the staged buffer
    int cr;          // DFT input n16 is sample m = (n16 + cr) & 15 of the lane, staged in row m of the frame
    int win_off;     // byte offset into the window table of sample m = 0; sample m adds m * win_step
    int win_step;    // + 4 WIN_P for the lanes that own a column, - 4 WIN_P for the mirrored lanes
};
ADY_HD StageAConst stage_a_const(int l, int col /* col_perm[l] */) {
    const int r = (16 * l) % 75, cr = (16 - (3 * r) % 16) % 16;
    // sample m of lane l is n = r + 75 m; w[n] = w[1200 - n] = sample 15 - m of lane 75 - l
    // The split at 44 (not at 38) makes the columns of every warp distinct modulo 32 -- warp 1 reads columns 32..43 and
    // 31..12 --, so the window loads are conflict-free although every lane reads its own row.
    const bool mir = l >= WIN_DIRECT;
    return {col * 8, cr, 4 * (mir ? 15 * WIN_P + (75 - l) : l), mir ? -4 * WIN_P : 4 * WIN_P};
}
ADY_HD int stage_a_sample(int l, int n16) {   // frame sample index lane l loads as DFT-16 input n16 (host-side table builder)
    const int r = (16 * l) % 75, cr = (16 - (3 * r) % 16) % 16;
    return r + 75 * ((n16 + cr) & 15);
}

ADY_HD void stage_a(const unsigned char* __restrict__ samp, const float* __restrict__ win, unsigned char* __restrict__ xf, int f, int l,
                    const StageAConst k) {
    const unsigned char* sp = samp + f * (8 * ROWP * 8) + k.col_off;
    c2 x[16];
#pragma unroll
    for (int n16 = 0; n16 < 16; ++n16) {
        const int m = (n16 + k.cr) & 15;                  // one wrapped row index serves the sample and the window address
        const int off = m * (ROWP * 8);
        uint32_t wy, zx;
#if defined(__CUDA_ARCH__)
        const uint2 v = *reinterpret_cast<const uint2*>(sp + off);
        wy = v.x; zx = v.y;
#else
        memcpy(&wy, sp + off, 4); memcpy(&zx, sp + off + 4, 4);
#endif
        const float w = *reinterpret_cast<const float*>(reinterpret_cast<const unsigned char*>(win) + (k.win_off + m * k.win_step));
#if defined(__CUDA_ARCH__)
        // int16 -> float without the XU pipe (64 I2F.S16 per task kept it busy for 8 cycles each and made it the
        // bottleneck of this stage): flip the sign bit (offset binary u = x + 32768), plant u in the mantissa of
        // 2^23 with one byte permute, and subtract 2^23 + 32768 -- exact -- with one packed add per channel pair.
        const uint32_t a = wy ^ 0x80008000u, c = zx ^ 0x80008000u;
        const f2 magic = ADY_K2(-8421376.0);                                          // -(2^23 + 2^15)
        const f2 re = mk2(__uint_as_float(__byte_perm(a, 0x4B000000u, 0x7410)), __uint_as_float(__byte_perm(c, 0x4B000000u, 0x7410))) + magic;
        const f2 im = mk2(__uint_as_float(__byte_perm(a, 0x4B000000u, 0x7432)), __uint_as_float(__byte_perm(c, 0x4B000000u, 0x7432))) + magic;
        x[n16].re = re * dup2(w);                                                     // (W, Z)
        x[n16].im = im * dup2(w);                                                     // (Y, X)
#else
        const float sW = (float)(int16_t)(wy & 0xffffu), sY = (float)(int16_t)(wy >> 16);
        const float sZ = (float)(int16_t)(zx & 0xffffu), sX = (float)(int16_t)(zx >> 16);
        x[n16].re = mk2(sW * w, sZ * w);
        x[n16].im = mk2(sY * w, sX * w);
#endif
    }
    p_dft16(x);
    unsigned char* xo = xf + f * X_STRIDE + l * 16;
#pragma unroll
    for (int s = 0; s < 16; ++s) st_c2(xo + dft16_slot_k(s) * (75 * 16), x[s]);   // X1 slot (k16, n75) = 75 k16 + n75
}

// ---------------------------------------------------------------- stage B: DFT-15 over a (n75 = 5 a + b), in place (80 tasks / frame)
// lane u: k16 = u & 15, b = u >> 4 (k16 fastest: the 8 lanes of a quarter-warp are 75 slots == 3 mod 8 apart).
// Output row c is stored at group pi(c) = 2 c mod 15 of the thread's own 15 slots: X2 slot (k16, c, b) = 75 k16 + 5 pi(c) + b.
// With the identity map, the V records of CONSECUTIVE bins (k16 + 1, c + 1) would sit 80 slots apart, i.e. in the same
// 16-byte bank group, and the mel gather (lanes = neighbouring bins) would serialise; with pi they are 85 slots apart.
ADY_HD constexpr int pi15(int c) { return (2 * c) % 15; }
ADY_HD void stage_b(unsigned char* __restrict__ xf, int f, int u) {
    unsigned char* p = xf + f * X_STRIDE + (75 * (u & 15) + (u >> 4)) * 16;
    c2 x[15];
#pragma unroll
    for (int a = 0; a < 15; ++a) ld_c2(p + a * 80, x[a]);
    p_dft15(x);
#pragma unroll
    for (int c = 0; c < 15; ++c) st_c2(p + pi15(c) * 80, x[c]);
}

#if defined(__CUDACC__)
// Stage-C twiddles in constant memory (filled by fe2.cu::get_fe2_tables): a warp's 32 pair-tasks span two values of c,
// so an indexed LDC replays twice -- and stays off the shared-memory pipe that bounds the kernel (8 LDS.64 per task).
static __constant__ float2 c_tw75[15 * 4];
#endif

// ---------------------------------------------------------------- stage C: twiddle W75^{b c}, DFT-5 over b, channel split, |X|^2, intensity
// Pair-task = row (k16, c) ("P") with its mirror row ((16 - k16) & 15, (15 - c) % 15) ("Q"): the thread owns
// bin k and bin 1200 - k of both packed FFTs, i.e. the four channel spectra of 5 bins.
//   regular tasks  tau = 0..111: k16 = tau & 15, c = (tau >> 4) + 1   (c = 1..7, mirror c' = 8..14), Q index 4 - d
//   c = 0 tasks    i   = 0..8  : i < 7: k16 = i + 1 with mirror 15 - i; i = 7: row (0,0), i = 8: row (8,0)
//                                (self-mirror rows: Q == P), Q index (5 - d) % 5
// Output bin of (row, d): k = (225 k16 + 976 (c + 15 d)) mod 1200  (k == k16 mod 16, k == c + 15 d mod 75).
// V records (16 bytes each) are written in place of the points just consumed:
//   Va = (|W|^2, |Z|^2, |Y|^2, |X|^2) over P's slot d,  Vb = (I_Y, I_Z, I_X, 0) / E over Q's slot d
//   (self-mirror rows: Vb goes to the spare slots 1200 + 5 (i - 7) + d).
// "+1e-8" of datasets.py:147, added analytically: bin 0 gets dc0 = 300 eps, bins +-1 get dc1 = -150 eps on the real
// AND imaginary part of each packed FFT (both of its channels carry the offset).  Under the rotation augmentation
// (rb bit0 = Y, bit1 = Z, bit2 = X negated) a sign-flipped channel is carried un-flipped through the kernel, so its
// offset enters with the opposite sign:  sign * (FFT(x) + sign * dc) == FFT(sign * x) + dc.
ADY_HD void add_dc(c2& v, float dc, unsigned rb) {
    v.re = v.re + mk2(dc, (rb & 2u) ? -dc : dc);                          // (W, Z)
    v.im = v.im + mk2((rb & 1u) ? -dc : dc, (rb & 4u) ? -dc : dc);        // (Y, X)
}

ADY_HD int bin_of(int k16, int c, int d) {
    const int k = (225 * k16 + 976 * (c + 15 * d)) % 1200;
    return k > 600 ? 1200 - k : k;
}

template <bool C0>
ADY_HD void stage_c_offsets(int task, int& pbase, int& qbase, int& vbbase, int& c, int& cq) {
    if (!C0) {
        const int k16 = task & 15;
        c = (task >> 4) + 1;
        cq = 15 - c;
        pbase = (75 * k16 + 5 * pi15(c)) * 16;
        qbase = (75 * ((16 - k16) & 15) + 5 * pi15(cq)) * 16;
        vbbase = qbase;
    } else {
        c = cq = 0;
        if (task < 7) { pbase = 75 * (task + 1) * 16; qbase = 75 * (15 - task) * 16; vbbase = qbase; }
        else          { pbase = qbase = (task == 7 ? 0 : 600) * 16; vbbase = (1200 + 5 * (task - 7)) * 16; }
    }
}

// E = eps + |W|^2 + (|Y|^2 + |Z|^2 + |X|^2) / 3 (datasets.py:272); I = Re(conj(W) X_c) / E (:271,274)
ADY_HD void split_power_iv(const c2 a, const c2 bq, float (&va)[4], float (&vb)[4], c2& s0, c2& s1) {
    // a = Z[k], bq = Z[N-k] of both packed FFTs (already halved by the window scale):
    // first channel S0 = a + conj(bq), second channel S1 = (a - conj(bq)) / i
    s0 = {a.re + bq.re, a.im - bq.im};          // (W, Z) spectra
    s1 = {a.im + bq.im, bq.re - a.re};          // (Y, X) spectra
    const f2 p0 = fma2(s0.re, s0.re, s0.im * s0.im), p1 = fma2(s1.re, s1.re, s1.im * s1.im);
    const float pw = lo2(p0), pz = hi2(p0), py = lo2(p1), px = hi2(p1);
    const float wre = lo2(s0.re), wim = lo2(s0.im);
    const float E = (pw + 1e-8f) + ((py + pz) + px) * (1.0f / 3.0f);
#if defined(__CUDA_ARCH__)
    const float rE = __fdividef(1.0f, E);       // MUFU.RCP (2 ulp): far inside the tolerance
#else
    const float rE = 1.0f / E;
#endif
    va[0] = pw; va[1] = pz; va[2] = py; va[3] = px;
    vb[0] = (wre * lo2(s1.re) + wim * lo2(s1.im)) * rE;    // I_Y
    vb[1] = (wre * hi2(s0.re) + wim * hi2(s0.im)) * rE;    // I_Z
    vb[2] = (wre * hi2(s1.re) + wim * hi2(s1.im)) * rE;    // I_X
    vb[3] = 0.f;
}

template <bool C0>
ADY_HD void stage_c_load(const unsigned char* __restrict__ xb, const unsigned char* __restrict__ tw, int pbase, int qbase, int c, int cq,
                         c2 (&P)[5], c2 (&Q)[5]) {
#pragma unroll
    for (int b = 0; b < 5; ++b) { ld_c2(xb + pbase + 16 * b, P[b]); ld_c2(xb + qbase + 16 * b, Q[b]); }
    if (!C0) {
#pragma unroll
        for (int b = 1; b < 5; ++b) {
            float2 tp, tq;                               // (wr, wi): the packed operations broadcast the scalar
#if defined(__CUDA_ARCH__)
            (void)tw;
            tp = c_tw75[c * 4 + (b - 1)];
            tq = c_tw75[cq * 4 + (b - 1)];
#else
            memcpy(&tp, tw + (c * 4 + (b - 1)) * 8, 8);
            memcpy(&tq, tw + (cq * 4 + (b - 1)) * 8, 8);
#endif
            const c2 p = P[b], q = Q[b];
            P[b] = {p.re * dup2(tp.x) - p.im * dup2(tp.y), fma2(p.re, dup2(tp.y), p.im * dup2(tp.x))};
            Q[b] = {q.re * dup2(tq.x) - q.im * dup2(tq.y), fma2(q.re, dup2(tq.y), q.im * dup2(tq.x))};
        }
    }
    p_dft5(P[0], P[1], P[2], P[3], P[4]);
    p_dft5(Q[0], Q[1], Q[2], Q[3], Q[4]);
}
template <bool C0> ADY_HD constexpr int mirror_d(int d) { return C0 ? (5 - d) % 5 : 4 - d; }

// FOA: one pair-task of frame buffer xb.  dc0 / dc1 applied to the rows holding bins 0 and +-1.
template <bool C0>
ADY_HD void stage_c_foa(unsigned char* __restrict__ xb, const unsigned char* __restrict__ tw, int task, float dc0, float dc1, unsigned rb) {
    int pbase, qbase, vbbase, c, cq;
    stage_c_offsets<C0>(task, pbase, qbase, vbbase, c, cq);
    c2 P[5], Q[5];
    stage_c_load<C0>(xb, tw, pbase, qbase, c, cq, P, Q);
    if (C0 && task == 7) {          // row (0,0): bin 0 is its own mirror (P == Q)
        add_dc(P[0], dc0, rb);
        add_dc(Q[0], dc0, rb);
    }
    if (!C0 && task == 1) {         // row (1,1) d = 0 is bin 1, its mirror row (15,14) d = 4 is bin 1199
        add_dc(P[0], dc1, rb);
        add_dc(Q[4], dc1, rb);
    }
#pragma unroll
    for (int d = 0; d < 5; ++d) {
        float va[4], vb[4];
        c2 s0, s1;
        split_power_iv(P[d], Q[mirror_d<C0>(d)], va, vb, s0, s1);
        st_f4(xb + pbase + 16 * d, va[0], va[1], va[2], va[3]);
        st_f4(xb + vbbase + 16 * d, vb[0], vb[1], vb[2], vb[3]);
    }
}

// ---------------------------------------------------------------- MIC format: |X|^2 for the log-mel + unit phasors for GCC-PHAT
// The GCC-PHAT cross spectrum of a microphone pair is R / |R| = conj(u_m) u_n with u_c = X_c / |X_c|, so the front end
// hands the lag-transform kernel one unit phasor per channel and bin as half2 (16 bytes per bin instead of the 32 bytes of
// four complex64 spectra; a vanishing channel is sent as (0, 0): every pair it takes part in has R = 0, phase 0).
// Phasor position of (pair-task, d), d-major: regular task tau -> 112 d + tau, c = 0 task i -> 560 + 9 d + i (605 positions,
// K = 608): the lanes of a warp hold consecutive tasks, so each of the five 16-byte stores of a task covers 512 contiguous
// bytes per warp (task-major positions put the lanes 80 bytes apart: every store touched twice the sectors).
constexpr int PH_K = 608;
ADY_HD int phasor_pos(bool c0, int task, int d) { return c0 ? 5 * NREG + d * NC0 + task : d * NREG + task; }
ADY_HD int raw_bin_of(int k16, int c, int d) { return (225 * k16 + 976 * (c + 15 * d)) % 1200; }

#if defined(__CUDA_ARCH__)
__device__ __forceinline__ uint32_t pack_half2(float re, float im) {
    uint32_t r;
    asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(im), "f"(re));   // upper half <- first source
    return r;
}
// unit phasor (re, im) / sqrt(p) of one channel as a half2 word, p = re^2 + im^2.  A vanishing channel (p below the
// smallest normal: rsqrt.ftz would return inf) is sent as the word 0 = (+0, +0) EXACTLY -- the lag-transform kernel
// recognises a vanishing channel by that word, and (-0) * 0 products would otherwise leave sign bits in it.
__device__ __forceinline__ uint32_t phasor_word(float re, float im, float p) {
    float r;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(p));
    const uint32_t w = pack_half2(re * r, im * r);
    return p >= 1.17549435e-38f ? w : 0u;
}
#endif

template <bool C0>
ADY_HD void stage_c_mic(unsigned char* __restrict__ xb, const unsigned char* __restrict__ tw, int task, float dc0, float dc1,
                        uint4* __restrict__ ph_frame /* this frame's PH_K phasor records, or host emulation buffer */) {
    int pbase, qbase, vbbase, c, cq;
    stage_c_offsets<C0>(task, pbase, qbase, vbbase, c, cq);
    c2 P[5], Q[5];
    stage_c_load<C0>(xb, tw, pbase, qbase, c, cq, P, Q);
    if (C0 && task == 7) { add_dc(P[0], dc0, 0u); add_dc(Q[0], dc0, 0u); }
    if (!C0 && task == 1) { add_dc(P[0], dc1, 0u); add_dc(Q[4], dc1, 0u); }
    const int k16 = C0 ? (task < 7 ? task + 1 : (task == 7 ? 0 : 8)) : (task & 15);
#pragma unroll
    for (int d = 0; d < 5; ++d) {
        float va[4], vb[4];
        c2 s0, s1;
        split_power_iv(P[d], Q[mirror_d<C0>(d)], va, vb, s0, s1);     // (the intensity part is dead code here)
        st_f4(xb + pbase + 16 * d, va[0], va[1], va[2], va[3]);
        // channel spectra at raw bin k_d; k_d > 600 holds the conjugate of bin 1200 - k_d
        const float sg = raw_bin_of(k16, c, d) > 600 ? -1.f : 1.f;
#if defined(__CUDA_ARCH__)
        uint4 rec;                                                     // va = (|W|^2, |Z|^2, |Y|^2, |X|^2)
        rec.x = phasor_word(lo2(s0.re), sg * lo2(s0.im), va[0]);        // channel 0 (W)
        rec.y = phasor_word(lo2(s1.re), sg * lo2(s1.im), va[2]);        // channel 1 (Y)
        rec.z = phasor_word(hi2(s0.re), sg * hi2(s0.im), va[1]);        // channel 2 (Z)
        rec.w = phasor_word(hi2(s1.re), sg * hi2(s1.im), va[3]);        // channel 3 (X)
        ph_frame[phasor_pos(C0, task, d)] = rec;
#else
        float* o = reinterpret_cast<float*>(ph_frame) + phasor_pos(C0, task, d) * 8;   // emulation: 8 floats per position
        const float pw[4] = {va[0], va[2], va[1], va[3]};
        const float re[4] = {lo2(s0.re), lo2(s1.re), hi2(s0.re), hi2(s1.re)}, im[4] = {lo2(s0.im), lo2(s1.im), hi2(s0.im), hi2(s1.im)};
        for (int ch = 0; ch < 4; ++ch) {
            const float r = pw[ch] > 0.f ? 1.0f / __builtin_sqrtf(pw[ch]) : 0.f;
            o[2 * ch] = re[ch] * r;
            o[2 * ch + 1] = sg * im[ch] * r;
        }
#endif
    }
    if (C0 && task == 8) {                                            // K padding 605..607: finite zeros
#if defined(__CUDA_ARCH__)
        for (int i = 605; i < PH_K; ++i) ph_frame[i] = make_uint4(0u, 0u, 0u, 0u);
#endif
    }
}

// host-side: folded FFT bin of phasor position p, or -1 (duplicate of a self-mirror row / padding)
inline void bins_of_phasor_pos(int (&bin_of_pos)[PH_K]) {
    bool seen[NBIN];
    for (int k = 0; k < NBIN; ++k) seen[k] = false;
    for (int p = 0; p < PH_K; ++p) bin_of_pos[p] = -1;
    for (int task = 0; task < NREG; ++task)
        for (int d = 0; d < 5; ++d) {
            const int k = bin_of(task & 15, (task >> 4) + 1, d);
            if (!seen[k]) { seen[k] = true; bin_of_pos[phasor_pos(false, task, d)] = k; }
        }
    for (int task = 0; task < NC0; ++task)
        for (int d = 0; d < 5; ++d) {
            const int k = bin_of(task < 7 ? task + 1 : (task == 7 ? 0 : 8), 0, d);
            if (!seen[k]) { seen[k] = true; bin_of_pos[phasor_pos(true, task, d)] = k; }
        }
}

// host-side: V record offsets of bin k (first occurrence), for the mel schedule
inline void v_offsets_of_bins(int (&offa)[NBIN], int (&offb)[NBIN]) {
    for (int k = 0; k < NBIN; ++k) offa[k] = offb[k] = -1;
    for (int task = 0; task < NREG; ++task) {
        int pb, qb, vb, c, cq;
        stage_c_offsets<false>(task, pb, qb, vb, c, cq);
        for (int d = 0; d < 5; ++d) {
            const int k = bin_of(task & 15, c, d);
            if (offa[k] < 0) { offa[k] = pb + 16 * d; offb[k] = vb + 16 * d; }
        }
    }
    for (int task = 0; task < NC0; ++task) {
        int pb, qb, vb, c, cq;
        stage_c_offsets<true>(task, pb, qb, vb, c, cq);
        const int k16 = task < 7 ? task + 1 : (task == 7 ? 0 : 8);
        for (int d = 0; d < 5; ++d) {
            const int k = bin_of(k16, 0, d);
            if (offa[k] < 0) { offa[k] = pb + 16 * d; offb[k] = vb + 16 * d; }
        }
    }
}

// ---------------------------------------------------------------- mel projection: one lane-job, both frames of the tile
// acc[f][0..3] = partial sums of (|W|^2,|Z|^2), (|Y|^2,|X|^2), (I_Y, I_Z), (I_X, -) as packed pairs
template <bool WITH_IV>
ADY_HD void mel_job(const unsigned char* __restrict__ x0, const MelEnt* __restrict__ ent_col, f2 (&acc)[TFR][4]) {
    // Straight-line code, no frame-count branch: with a one-frame tile the second frame's buffer holds stale (finite
    // or not) values whose sums are simply never stored; this lets the compiler issue all entry loads, then all
    // gathers, instead of one dependent shared-memory round trip after the other.
#pragma unroll
    for (int f = 0; f < TFR; ++f)
#pragma unroll
        for (int i = 0; i < 4; ++i) acc[f][i] = ADY_K2(0.0);
    MelEnt e[MEL_L];
#pragma unroll
    for (int it = 0; it < MEL_L; ++it) {
#if defined(__CUDA_ARCH__)
        // one 8-byte load per entry in every instantiation: without the intensity records (MIC) the compiler would
        // otherwise fetch offa and w with two narrow loads of two wavefronts each
        const uint2 raw = *reinterpret_cast<const uint2*>(&ent_col[it * NJOBS]);
        e[it].offa = (uint16_t)(raw.x & 0xffffu);
        e[it].offb = (uint16_t)(raw.x >> 16);
        e[it].w = __uint_as_float(raw.y);
#else
        e[it] = ent_col[it * NJOBS];
#endif
    }
#pragma unroll
    for (int it = 0; it < MEL_L; ++it) {
        const f2 w = dup2(e[it].w);
#pragma unroll
        for (int f = 0; f < TFR; ++f) {
            c2 a, b;
            ld_c2(x0 + f * X_STRIDE + e[it].offa, a);
            acc[f][0] = fma2(a.re, w, acc[f][0]);
            acc[f][1] = fma2(a.im, w, acc[f][1]);
            if (WITH_IV) {
                ld_c2(x0 + f * X_STRIDE + e[it].offb, b);
                acc[f][2] = fma2(b.re, w, acc[f][2]);
                acc[f][3] = fma2(b.im, w, acc[f][3]);
            }
        }
    }
}

// librosa.power_to_db(ref=1, amin=1e-10) before the top_db clamp (datasets.py:265)
ADY_HD float power_to_db(float s) {
#if defined(__CUDA_ARCH__)
    float l;   // 10 log10(s) = 3.0103 log2(s); argument >= 1e-10 is never denormal -> bare MUFU.LG2
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l) : "f"(fmaxf(s, 1e-10f)));
    return 3.0102999566398120f * l;
#else
    return 10.0f * __builtin_log10f(s > 1e-10f ? s : 1e-10f);
#endif
}

}  // namespace fe2
}  // namespace ady
