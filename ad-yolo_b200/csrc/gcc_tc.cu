// adyolo_gcc_from_stft: NOT in the reference; upstream seld-dcase2022 _get_gcc semantics,
//   cc = irfft(exp(1j * angle(conj(X_m) X_n))), lags [-32, 32), 6 microphone pairs.
//
// GCC-PHAT lag transform on the 5th-generation tensor cores (tcgen05.mma kind::tf32, accumulators
// in TMEM).  This is the one dense contraction on the path: for every (frame, mic
// pair) "item" the 64 lags are   cc[l] = (1/N) sum_k c_k Re( P_k exp(+2 pi i k l / N) ),
// i.e. D[item][lag] = A[item][k'] . B[k'][lag] with k' = (2k, 2k+1) <-> (Re P_k, Im P_k) and
// B[2k][l] = c_k cos(2 pi k l/N), B[2k+1][l] = -c_k sin(2 pi k l/N)  (c_0 = c_600 = 1/2, else 1;
// the common factor 2/N is applied in FP32 in the epilogue so that lag 0 is exact in TF32).
//
// One CTA = 64 frames = 384 items = 3 UMMA M-tiles of 128 rows (tile = pair / 2, row = 2 frame +
// (pair & 1)), N = 64 lags, K = 1216 (608 bins), 2 CTAs per SM.  Per chunk of 8 bins (K = 16):
//   - cp.async stages the raw spectra two chunks ahead (3 stages); every thread copies exactly the
//     16-byte pieces (channel pair x bin x frame) it later consumes, so they need no barrier
//   - a lane normalises its two channels to unit modulus (one rsqrt.approx each; rare slow path for
//     |x|^2 outside the FP32 range or zero), swaps them with its partner lane (the other two channels
//     of the same bin) and forms three of the six cross spectra; they are rounded to TF32 and stored as
//     the A operand in the canonical K-major, no-swizzle UMMA layout (core matrix = 8 rows x 16 bytes).
//     The k-chunk pitch is skewed by 32 bytes and the two lanes of a bin always store pairs of opposite
//     row parity, so the 16 stores of a half-warp hit 16 distinct 8-byte slots
//   - the B chunk (4 KB, precomputed in the same layout) arrives by cp.async as well
//   - fence.proxy.async + barrier; one thread issues 2 (k-steps) x 3 (M-tiles) tcgen05.mma and a
//     tcgen05.commit on the A stage's mbarrier.  Two A stages: chunk c+1 is generated while the tensor
//     core consumes chunk c; a stage is only waited for when it is about to be overwritten
// Epilogue: all 8 warps tcgen05.ld (32x32b.x32) a 32-row x 32-lag block each, stage it through
// shared memory and store standardised 128-byte row segments.
// The MMA truncates FP32 operands to TF32, which biases a coherent peak low by ~6e-4; both operands
// are therefore rounded to nearest first: measured max error 2.9e-5 absolute against the float64
// oracle (gate 1e-3); all-(1,0) cross spectra (digital silence) give cc[0] = 1 exactly.
// B200 (same box, CUDA events): 128 x 5-s clips 0.57 ms (CUDA-core kernel this replaces) -> 0.19 ms;
// 512 clips: 0.775 ms with a one-chunk register prefetch -> 0.606 ms = 3.5 TB/s of spectra + output
// (profiles/r01_ncu_gcc_tc.txt).
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <atomic>
#include <mutex>
#include <vector>

#include "common.cuh"
#include "frontend_core.cuh"
#include <cuda_fp16.h>

#include "fe2_core.cuh"
#include "frontend_host.h"

namespace ady {

struct OutStrides {
    long long sb, sc, st, sj;
};

constexpr int GT_FRAMES = 64;                 // frames per CTA
constexpr int GT_ITEMS = GT_FRAMES * 6;       // 384 rows = 3 M-tiles; row = pair * 64 + frame
constexpr int GT_BINS = 8;                    // bins per chunk -> K = 16 per chunk (2 MMA k-steps)
constexpr int GT_CHUNKS = 76;                 // 608 bins >= 601
constexpr int GT_THREADS = 256;
constexpr int GT_STAGES = 2;                  // A-operand stages (generation of chunk c+1 overlaps the MMAs of chunk c)
constexpr int GT_KCH = GT_BINS * 2 / 4;       // 16-byte k-chunks per stage (4)
// Two input formats (template parameter PH):
//   PH = false: raw channel spectra, complex64 (B, T, 601, 4) -- the compatibility entry adyolo_gcc_from_stft;
//               a 16-byte piece is one channel PAIR of a (frame, bin); partner lanes swap the other pair
//   PH = true : per-channel unit phasors as half2, (B, T, 608) x 16 bytes in the fe2 kernel's position order
//               (fe2_core.cuh::phasor_pos) -- half the bytes, no rsqrt and no lane exchange here: a 16-byte
//               piece is ALL four channels of a (frame, position) and its lane forms all six cross spectra
// A operand: the k-chunk pitch (= LBO) is skewed (32 B / 16 B) so that the 8-byte stores of a half-warp hit 16 distinct
// 8-byte slots (see the store sites)
template <bool PH> struct GtCfg {
    static constexpr int A_LBO = (GT_ITEMS / 8) * 128 + (PH ? 16 : 32);       // 6160 | 6176
    static constexpr int A_BYTES = GT_KCH * A_LBO;
    static constexpr int RAW_BYTES = GT_FRAMES * GT_BINS * (PH ? 16 : 32);    // 8192 | 16384
    static constexpr int PIECES = RAW_BYTES / 16 / GT_THREADS;                // 16-byte pieces per thread and chunk (2 | 4)
    static constexpr int OFF_B = GT_STAGES * A_BYTES;
    static constexpr int OFF_RAW = OFF_B + 4 * 4096;
    static constexpr int SMEM = OFF_RAW + 3 * RAW_BYTES;                       // <= 114944 B: two CTAs per SM
};
constexpr int GT_B_LBO = (64 / 8) * 128;                               // 1024
constexpr int GT_B_BYTES = GT_KCH * GT_B_LBO;                          // 4096
constexpr int GT_B_SLOTS = 4;                 // B chunk c+2 is copied while the MMAs of chunk c-1 may still read theirs
constexpr int GT_RAW_STAGES = 3;              // raw input: chunks c+1 and c+2 in flight while chunk c is consumed
constexpr uint32_t GT_TMEM_COLS = 256;        // 3 x 64 fp32 columns, power of two

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// K-major, SWIZZLE_NONE shared-memory matrix descriptor (sm_100 format: version field = 1).
// LBO = byte distance between core matrices adjacent in K, SBO = between 8-row groups in M/N.
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3fffu);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3fffu) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3fffu) << 32;
    d |= 1ull << 46;
    return d;
}

// instruction descriptor: D = F32, A = B = TF32, both K-major, N = 64, M = 128
constexpr uint32_t GT_IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((64u >> 3) << 17) | ((128u >> 4) << 24);

__device__ __forceinline__ void mbar_wait(uint32_t mb, uint32_t parity) {
    uint32_t done = 0;
    while (!done) {
        asm volatile(
            "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done) : "r"(mb), "r"(parity) : "memory");
    }
}

// The MMA reads only the upper 19 bits of an FP32 operand (truncation).  Adding half a TF32 ulp to
// the bit pattern first turns that into round-to-nearest (ties away), like cvt.rna.tf32.f32 but in
// one integer add; operands here are bounded by 1, so the carry cannot reach infinity.
__device__ __forceinline__ float tf32_rn_bits(float x) { return __uint_as_float(__float_as_uint(x) + 0x1000u); }

__device__ __forceinline__ float rsqrt_ftz(float x) {
    float r;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

// x / |x| per channel.  The value is brought to [0.5, 1) by an exact power of two before the rsqrt,
// so no input magnitude over- or underflows.  A vanishing channel sets `zero`: every pair it takes
// part in has a vanishing cross spectrum, whose phase upstream is np.angle(0) = 0 -> (1, 0).
__device__ __forceinline__ float2 unit(float2 a, bool& zero) {
    const float mx = fmaxf(fabsf(a.x), fabsf(a.y));
    zero = mx < 1.17549435e-38f;
    const int e = (int)(__float_as_uint(mx) >> 23);
    const float sc = __uint_as_float((uint32_t)max(253 - e, 1) << 23);
    const float re = a.x * sc, im = a.y * sc;
    const float r = rsqrt_ftz(re * re + im * im);
    return make_float2(re * r, im * r);
}

// Piece ownership (PH = false): piece q = tid + 256 j of a chunk is (frame q / 16, bin (q % 16) / 2, channel half
// q % 2), i.e. consecutive lanes copy consecutive 16-byte pieces (one frame's 8 bins x 4 channels are
// 256 contiguous bytes) and consume exactly what they copied; the partner holding the other two
// channels of the same (frame, bin) is lane ^ 1.
__device__ __forceinline__ int gt_half(int tid) { return tid & 1; }
__device__ __forceinline__ int gt_kb(int tid) { return (tid & 15) >> 1; }
__device__ __forceinline__ int gt_frame(int tid, int j) { return (tid + GT_THREADS * j) >> 4; }
// PH = true: piece q = tid + 256 j is (position q % 8 of the chunk, frame slot q / 8); the four frame slots of a warp are
// frames 4 w + {0, 2, 1, 3}: the two frames of a half-warp are 2 apart = 4 rows = 64 bytes in the A operand, which
// together with the 16-byte k-chunk skew makes the 16 float2 stores of a half-warp conflict-free.
__device__ __forceinline__ int ph_kb(int tid) { return tid & 7; }
__device__ __forceinline__ int ph_frame(int tid, int j) {
    const int g = (tid + GT_THREADS * j) >> 3, i = g & 3;
    return (g & ~3) + ((i & 1) << 1) + (i >> 1);
}

// Asynchronous staging of one chunk: every thread copies the 16-byte pieces that it will itself consume, so the raw
// data need no barrier at all (cp.async.wait_group only), plus one piece of the B operand chunk.  Pieces past the
// last frame / bin 600 are zero-filled (src-size 0).
template <bool PH>
__device__ __forceinline__ void issue_chunk_copy(unsigned char* smem, const void* __restrict__ in, const float* __restrict__ btab,
                                                 long long f0, long long n_frames, int ch, int tid) {
    using Cfg = GtCfg<PH>;
    if (ch < GT_CHUNKS) {
        const uint32_t raw = smem_u32(smem + Cfg::OFF_RAW + (ch % GT_RAW_STAGES) * Cfg::RAW_BYTES) + tid * 16;
#pragma unroll
        for (int j = 0; j < Cfg::PIECES; ++j) {
            const void* src;
            bool live;
            if (PH) {
                const long long fr = f0 + ph_frame(tid, j);
                live = fr < n_frames;
                src = live ? (const void*)(reinterpret_cast<const uint4*>(in) + fr * 608 + ch * GT_BINS + ph_kb(tid)) : in;
            } else {
                const int f = gt_frame(tid, j), bin = ch * GT_BINS + gt_kb(tid);
                const long long fr = f0 + f;
                live = fr < n_frames && bin < NBIN;
                src = live ? (const void*)(reinterpret_cast<const float2*>(in) + (fr * NBIN + bin) * 4 + gt_half(tid) * 2) : in;
            }
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(raw + j * (GT_THREADS * 16)), "l"(src), "r"(live ? 16 : 0) : "memory");
        }
        const uint32_t d = smem_u32(smem + Cfg::OFF_B + (ch % GT_B_SLOTS) * GT_B_BYTES) + tid * 16;
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(btab + (size_t)ch * (GT_B_BYTES / 4) + tid * 4) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
}

// ---- epilogue (shared by the TF32 and the FP16 kernels).  A warp may read TMEM lanes 32 (w % 4) .. +31: warps w and
// w + 4 share a lane quarter and split the 64 lag columns.  TMEM lane = row of the M-tile = 2 * frame + (pair & 1),
// pair = 2 * tile + (row & 1): a quarter holds 16 frames x the tile's two pairs.  Natural output layout: the 32 x 32
// block goes through shared memory so that global stores are 128-byte row segments; any other stride set is stored
// directly.
__device__ __forceinline__ void gcc_epilogue(uint32_t tmem, unsigned char* smem, int warp, int lane, long long f0, long long n_frames,
                                             int T, const float* __restrict__ mean, const float* __restrict__ istd,
                                             float* __restrict__ out, const OutStrides& os) {
    {
        constexpr float kScale = 2.0f / NFFT;               // B holds cos / -sin (x 1/2 at bins 0 and 600): exact at lag 0
        const int q = warp & 3, h = warp >> 2;
        const bool vec = os.sj == 1 && ((os.sb | os.sc | os.st) & 3) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0;
        float* stage = reinterpret_cast<float*>(smem) + warp * (32 * 36);      // 32 rows, 144-byte pitch: conflict-free
        // vec: this lane's 8 (row, 16-byte column) targets, row = 4 it + lane / 8; otherwise row = lane
        const int pp = vec ? (lane >> 3) & 1 : lane & 1;                       // pair parity of the lane's row(s)
        // frame of target `it`: f0 + 16 q + lane / 16 + 2 it (vec) or f0 + 16 q + lane / 2 (all eight the same).  One
        // division for the first target, then (clip, frame-in-clip) advance by two frames with a carry (eight 64-bit
        // divisions per lane were 19 % of the kernel's instructions)
        long long off[8];
        {
            const long long fr0 = f0 + 16 * q + (vec ? lane >> 4 : lane >> 1);
            long long b = fr0 / T;
            long long t = fr0 - b * T;
#pragma unroll
            for (int it = 0; it < 8; ++it) {
                const long long fr = fr0 + (vec ? 2 * it : 0);
                off[it] = fr < n_frames ? b * os.sb + t * os.st + pp * os.sc + (vec ? h * 32 + (lane & 7) * 4 : 0) : -1;
                if (vec) {
                    t += 2;
                    while (t >= T) { t -= T; ++b; }      // T = 1 moves two clips on
                }
            }
        }
#pragma unroll 1
        for (int mt = 0; mt < 3; ++mt) {
            uint32_t r[32];
            const uint32_t taddr = tmem + ((uint32_t)(q * 32) << 16) + mt * 64 + h * 32;
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                : "r"(taddr));
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            const int p = mt * 2 + pp;
            const float* mu = mean ? mean + p * NMEL + h * 32 : nullptr;
            const float* is = istd ? istd + p * NMEL + h * 32 : nullptr;
            if (vec) {
#pragma unroll
                for (int j4 = 0; j4 < 8; ++j4)
                    *reinterpret_cast<float4*>(stage + lane * 36 + j4 * 4) =
                        make_float4(__uint_as_float(r[4 * j4]), __uint_as_float(r[4 * j4 + 1]), __uint_as_float(r[4 * j4 + 2]), __uint_as_float(r[4 * j4 + 3]));
                __syncwarp();
                const int c4 = (lane & 7) * 4;
                const float4 m4 = mu ? *reinterpret_cast<const float4*>(mu + c4) : make_float4(0.f, 0.f, 0.f, 0.f);
                const float4 i4 = is ? *reinterpret_cast<const float4*>(is + c4) : make_float4(1.f, 1.f, 1.f, 1.f);
#pragma unroll
                for (int it = 0; it < 8; ++it) {
                    float4 v = *reinterpret_cast<const float4*>(stage + (4 * it + (lane >> 3)) * 36 + c4);
                    v.x = (v.x * kScale - m4.x) * i4.x; v.y = (v.y * kScale - m4.y) * i4.y;
                    v.z = (v.z * kScale - m4.z) * i4.z; v.w = (v.w * kScale - m4.w) * i4.w;
                    if (off[it] >= 0) *reinterpret_cast<float4*>(out + off[it] + (2 * mt) * os.sc) = v;
                }
                __syncwarp();
            } else if (off[0] >= 0) {
                float* o = out + off[0] + (2 * mt) * os.sc + (long long)(h * 32) * os.sj;
#pragma unroll
                for (int j = 0; j < 32; ++j)
                    o[j * os.sj] = (__uint_as_float(r[j]) * kScale - (mu ? mu[j] : 0.f)) * (is ? is[j] : 1.f);
            }
        }
    }
}

template <bool PH>
__global__ void __launch_bounds__(GT_THREADS, 2)
gcc_tc_kernel(const void* __restrict__ spec, long long n_frames, int T, const float* __restrict__ btab,
              const float* __restrict__ mean, const float* __restrict__ istd, float* __restrict__ out, OutStrides os) {
    using Cfg = GtCfg<PH>;
    constexpr int GT_A_LBO = Cfg::A_LBO, GT_A_BYTES = Cfg::A_BYTES, GT_OFF_B = Cfg::OFF_B, GT_OFF_RAW = Cfg::OFF_RAW,
                  GT_RAW_BYTES = Cfg::RAW_BYTES, GT_PIECES = Cfg::PIECES;
    extern __shared__ __align__(128) unsigned char smem[];      // A x2 | B x4 | raw x3
    __shared__ __align__(8) unsigned long long mbar[GT_STAGES];
    __shared__ uint32_t tmem_base_s;
    static_assert(GT_B_BYTES == GT_THREADS * 16, "one 16-byte B piece per thread");

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const long long f0 = (long long)blockIdx.x * GT_FRAMES;

    issue_chunk_copy<PH>(smem, spec, btab, f0, n_frames, 0, tid);
    issue_chunk_copy<PH>(smem, spec, btab, f0, n_frames, 1, tid);

    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(GT_TMEM_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < GT_STAGES; ++s) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&mbar[s])));
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tmem = tmem_base_s;

    const int half = gt_half(tid), kb = gt_kb(tid);

    for (int ch = 0; ch < GT_CHUNKS; ++ch) {
        const int stg = ch & (GT_STAGES - 1);
        unsigned char* sA = smem + stg * GT_A_BYTES;
        // A stage (and the B slot chunk ch+2 is about to overwrite) free once the MMAs of chunk ch-2 are done
        if (ch >= GT_STAGES) mbar_wait(smem_u32(&mbar[stg]), (uint32_t)(((ch - GT_STAGES) / GT_STAGES) & 1));
        issue_chunk_copy<PH>(smem, spec, btab, f0, n_frames, ch + 2, tid);
        asm volatile("cp.async.wait_group 2;" ::: "memory");            // this thread's pieces of chunk ch have landed

        const float4* raw = reinterpret_cast<const float4*>(smem + GT_OFF_RAW + (ch % GT_RAW_STAGES) * GT_RAW_BYTES) + tid;
        if (PH) {
            // phasor input: this lane holds all four unit phasors of (frame f, position kb) and forms the six cross
            // spectra conj(u_m) u_n; a zero phasor marks a vanishing channel -> (1, 0).  Pair p = 0..5 = (0,1) (0,2) (0,3)
            // (1,2) (1,3) (2,3) goes to M-tile p / 2, row 2 f + (p & 1).
#pragma unroll
            for (int j = 0; j < GT_PIECES; ++j) {
                const int f = ph_frame(tid, j), pkb = ph_kb(tid);
                const uint4 v = *reinterpret_cast<const uint4*>(&raw[j * GT_THREADS]);
                float2 u[4];
                {
                    const __half2 h0 = *reinterpret_cast<const __half2*>(&v.x), h1 = *reinterpret_cast<const __half2*>(&v.y);
                    const __half2 h2 = *reinterpret_cast<const __half2*>(&v.z), h3 = *reinterpret_cast<const __half2*>(&v.w);
                    u[0] = __half22float2(h0); u[1] = __half22float2(h1); u[2] = __half22float2(h2); u[3] = __half22float2(h3);
                }
                const bool z0 = v.x == 0u, z1 = v.y == 0u, z2 = v.z == 0u, z3 = v.w == 0u;     // (+0, +0): written by the front end
                const bool zz[4] = {z0, z1, z2, z3};
                unsigned char* base = sA + (pkb >> 1) * GT_A_LBO + (pkb & 1) * 8;
#pragma unroll
                for (int p = 0; p < 6; ++p) {
                    constexpr int PM[6] = {0, 0, 0, 1, 1, 2}, PN[6] = {1, 2, 3, 2, 3, 3};
                    const float2 a = u[PM[p]], b2 = u[PN[p]];
                    float2 P = make_float2(a.x * b2.x + a.y * b2.y, a.x * b2.y - a.y * b2.x);
                    if (zz[PM[p]] || zz[PN[p]]) P = make_float2(1.f, 0.f);
                    const int r = 2 * f + (p & 1);
                    *reinterpret_cast<float2*>(base + (p >> 1) * 2048 + (r >> 3) * 128 + (r & 7) * 16) =
                        make_float2(tf32_rn_bits(P.x), tf32_rn_bits(P.y));
                }
            }
        } else {
#pragma unroll
        for (int j = 0; j < GT_PIECES; ++j) {
            const int f = gt_frame(tid, j);
            const float4 v = raw[j * GT_THREADS];
            // this lane's two channels (2 half, 2 half + 1) -> unit modulus
            float2 ua = make_float2(v.x, v.y), ub = make_float2(v.z, v.w);
            const float sa = fmaf(ua.x, ua.x, ua.y * ua.y), sb = fmaf(ub.x, ub.x, ub.y * ub.y);
            unsigned zbits = 0;
            if (sa > 1e-30f && sa < 1e30f && sb > 1e-30f && sb < 1e30f) {
                const float ra = rsqrt_ftz(sa), rb = rsqrt_ftz(sb);
                ua.x *= ra; ua.y *= ra; ub.x *= rb; ub.y *= rb;
            } else {   // rare: silence, padding, extreme magnitudes, NaN
                bool za, zb;
                ua = unit(ua, za); ub = unit(ub, zb);
                zbits = (za ? 1u : 0u) | (zb ? 2u : 0u);
            }
            // the partner lane holds the other two channels of the same (frame, bin)
            const float2 oa = make_float2(__shfl_xor_sync(0xffffffffu, ua.x, 1), __shfl_xor_sync(0xffffffffu, ua.y, 1));
            const float2 ob = make_float2(__shfl_xor_sync(0xffffffffu, ub.x, 1), __shfl_xor_sync(0xffffffffu, ub.y, 1));
            const unsigned zo = __shfl_xor_sync(0xffffffffu, zbits, 1);
            // even lane (channels 0,1): pairs (0,1) (0,2) (0,3) = p 0, 1, 2 = conj(ua) ub, conj(ua) oa, conj(ua) ob
            // odd lane  (channels 2,3): pairs (2,3) (1,3) (1,2) = p 5, 4, 3 = conj(ua) ub, conj(ob) ub, conj(ob) ua
            // (in this order the two lanes of a (frame, bin) always store pairs of opposite parity, see below)
            const float2 X = half ? ob : ua, Y1 = half ? ub : oa, Y2 = half ? ua : ob;
            const bool zX = half ? (zo & 2u) : (zbits & 1u), zY1 = half ? (zbits & 2u) : (zo & 1u), zY2 = half ? (zbits & 1u) : (zo & 2u);
            float2 P0 = make_float2(ua.x * ub.x + ua.y * ub.y, ua.x * ub.y - ua.y * ub.x);
            float2 P1 = make_float2(X.x * Y1.x + X.y * Y1.y, X.x * Y1.y - X.y * Y1.x);
            float2 P2 = make_float2(X.x * Y2.x + X.y * Y2.y, X.x * Y2.y - X.y * Y2.x);
            if (zbits | zo) {
                if (zbits) P0 = make_float2(1.f, 0.f);
                if (zX || zY1) P1 = make_float2(1.f, 0.f);
                if (zX || zY2) P2 = make_float2(1.f, 0.f);
            }
            // canonical K-major layout: [k-chunk = kb/2][m-group = row/8] core matrices of 8 rows x 16 B.
            // M-tile = p / 2, row in the tile = 2 f + (p & 1): 64-bit stores are served per half-warp (one frame:
            // 8 bins x 2 channel halves); the two halves store pairs of opposite parity = adjacent rows, so with
            // the 32-byte k-chunk skew the 16 lanes hit 16 distinct 8-byte slots.
            const int r0 = 2 * f + half, r1 = 2 * f + 1 - half;            // rows of (p0) and of (p1, p2)
            unsigned char* d0 = sA + (kb >> 1) * GT_A_LBO + (kb & 1) * 8 + (r0 >> 3) * 128 + (r0 & 7) * 16;
            unsigned char* d1 = sA + (kb >> 1) * GT_A_LBO + (kb & 1) * 8 + (r1 >> 3) * 128 + (r1 & 7) * 16;
            // tiles: p0 = 0 | 5 -> tile 0 | 2;  p1 = 1 | 4 -> tile 0 | 2;  p2 = 2 | 3 -> tile 1 | 1
            *reinterpret_cast<float2*>(d0 + (half ? 2 : 0) * 2048) = make_float2(tf32_rn_bits(P0.x), tf32_rn_bits(P0.y));
            *reinterpret_cast<float2*>(d1 + (half ? 2 : 0) * 2048) = make_float2(tf32_rn_bits(P1.x), tf32_rn_bits(P1.y));
            *reinterpret_cast<float2*>(d0 + 2048) = make_float2(tf32_rn_bits(P2.x), tf32_rn_bits(P2.y));
        }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy / cp.async writes -> async proxy (UMMA)
        __syncthreads();
        if (tid == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;");
            const uint32_t a0 = smem_u32(sA), b0 = smem_u32(smem + GT_OFF_B + (ch % GT_B_SLOTS) * GT_B_BYTES);
#pragma unroll
            for (int ks = 0; ks < GT_KCH / 2; ++ks) {                   // K = 8 per MMA = 2 k-chunks
                const uint64_t bd = umma_desc(b0 + ks * 2 * GT_B_LBO, GT_B_LBO, 128);
#pragma unroll
                for (int mt = 0; mt < 3; ++mt) {
                    const uint64_t ad = umma_desc(a0 + ks * 2 * GT_A_LBO + mt * 16 * 128, GT_A_LBO, 128);
                    const uint32_t acc = (ch | ks) ? 1u : 0u;           // first MMA overwrites the accumulator
                    asm volatile(
                        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                        ::"r"(tmem + mt * 64), "l"(ad), "l"(bd), "r"(GT_IDESC), "r"(acc) : "memory");
                }
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&mbar[stg])) : "memory");
        }
    }
    asm volatile("cp.async.wait_all;" ::: "memory");
    // the last commit covers every MMA issued before it
    mbar_wait(smem_u32(&mbar[(GT_CHUNKS - 1) & (GT_STAGES - 1)]), (uint32_t)(((GT_CHUNKS - 1) / GT_STAGES) & 1));
    asm volatile("tcgen05.fence::after_thread_sync;");

    gcc_epilogue(tmem, smem, warp, lane, f0, n_frames, T, mean, istd, out, os);
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(GT_TMEM_COLS));
}

// B operand table: [GT_CHUNKS][GT_KCH k-chunks][8 lag-groups][8 lags][4 k'] floats (canonical K-major layout).
// K index = FFT bin (PH = false) or the fe2 kernel's phasor position (PH = true; duplicate / padding positions
// get all-zero rows).
static int get_gcc_btab(const float** dev_tab, bool ph) {
    static std::mutex mu;
    static float* cache[2][64] = {{nullptr}, {nullptr}};
    int dev = 0;
    ADY_CUDA_CHECK(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64) return set_error(ADY_ERR_INVALID, "device index %d out of range", dev);
    std::lock_guard<std::mutex> lk(mu);
    if (!cache[ph][dev]) {
        int bin_of_pos[fe2::PH_K];
        fe2::bins_of_phasor_pos(bin_of_pos);
        std::vector<float> h((size_t)GT_CHUNKS * GT_B_BYTES / 4, 0.f);
        for (int ch = 0; ch < GT_CHUNKS; ++ch)
            for (int kp = 0; kp < 2 * GT_BINS; ++kp) {            // k' within the chunk
                const int kidx = ch * GT_BINS + (kp >> 1);
                const int bin = ph ? bin_of_pos[kidx] : (kidx < NBIN ? kidx : -1);
                if (bin < 0) continue;
                const double ck = (bin == 0 || bin == 600) ? 0.5 : 1.0;    // x 2/N in the epilogue
                for (int n = 0; n < 64; ++n) {
                    const int lag = n - 32;                       // output order cc[-32:], cc[:32]
                    const double ang = 2.0 * M_PI * (double)((bin * (long long)((lag + NFFT) % NFFT)) % NFFT) / NFFT;
                    const double v = (kp & 1) ? -ck * sin(ang) : ck * cos(ang);
                    const size_t off = (size_t)ch * (GT_B_BYTES / 4) + ((size_t)(kp >> 2) * 8 + (n >> 3)) * 32 + (n & 7) * 4 + (kp & 3);
                    static_assert(GT_B_LBO == 8 * 128, "B k-chunk pitch");
                    float fv = (float)v;                          // round to TF32 (nearest, ties away) like cvt.rna
                    uint32_t u;
                    memcpy(&u, &fv, 4);
                    u = (u + 0x1000u) & ~0x1fffu;
                    memcpy(&fv, &u, 4);
                    h[off] = fv;
                }
            }
        float* d = nullptr;
        ADY_CUDA_CHECK(cudaMalloc(&d, h.size() * sizeof(float)));
        ADY_CUDA_CHECK(cudaMemcpy(d, h.data(), h.size() * sizeof(float), cudaMemcpyHostToDevice));
        cache[ph][dev] = d;
    }
    *dev_tab = cache[ph][dev];
    return ADY_OK;
}

template <bool PH>
static int launch_gcc(const void* in, int B, long long T, const float* mean, const float* istd, float* out, OutStrides os,
                      cudaStream_t stream) {
    const float* btab = nullptr;
    int rc = get_gcc_btab(&btab, PH);
    if (rc) return rc;
    const long long n_frames = (long long)B * T;
    const long long blocks = (n_frames + GT_FRAMES - 1) / GT_FRAMES;
    if (blocks > 0x7fffffffLL) return set_error(ADY_ERR_INVALID, "gcc: too many frames");
    static std::atomic<unsigned long long> configured{0};      // per (instantiation, device); the call is idempotent
    int dev = 0;
    ADY_CUDA_CHECK(cudaGetDevice(&dev));
    const int shmem = GtCfg<PH>::SMEM;
    if (!((configured.load(std::memory_order_acquire) >> (dev & 63)) & 1ull)) {
        ADY_CUDA_CHECK(cudaFuncSetAttribute(gcc_tc_kernel<PH>, cudaFuncAttributeMaxDynamicSharedMemorySize, shmem));
        configured.fetch_or(1ull << (dev & 63), std::memory_order_release);
    }
    gcc_tc_kernel<PH><<<(unsigned)blocks, GT_THREADS, shmem, stream>>>(in, n_frames, (int)T, btab, mean, istd, out, os);
    ADY_LAUNCH_CHECK("gcc_tc_kernel");
    return ADY_OK;
}

// ------------------------------------------------------------------------------------------------------------------
// FP16 variant for the phasor input (the product path of adyolo_features_mic_gcc).  The unit phasors arrive as half2,
// so the six cross spectra conj(u_m) u_n are formed with packed half arithmetic on TWO positions at a time
// (HMUL2 / HFMA2: 4 instructions per microphone pair and position pair instead of 8 FP32 ones plus 4 roundings to
// TF32) and stored as the FP16 A operand of tcgen05.mma kind::f16: half the shared-memory bytes and half the MMAs of
// the TF32 kernel above.  Rounding: two half roundings per component (<= 5e-4 each, zero-mean) average out over the
// 1202-term sum with its 2 / N factor: ~1e-5 on an output of magnitude <= 1 (measured against the float64 oracle in
// tests/test_gpu_features.py, gate 1e-3).
//
// One chunk = 16 positions = K 32 = four 16-byte k-chunks; a lane pairs position i with position i + 8 of the chunk, so
// k-chunk q holds [re p, re p+8, im p, im p+8] for p = 2 q and again for p = 2 q + 1 (the B table is built in that
// order); 38 chunks.  Thread roles inside a warp (the warp owns frames 8 w .. 8 w + 7 of the CTA):
//   copy   : piece (j, lane) = frame 8 w + 2 j + lane / 16, position lane % 16 -- a half-warp copies the 256 contiguous
//            bytes of one frame, destinations linear in the lane (a permuted LDGSTS destination serialises into 32
//            shared-memory wavefronts: measured, profiles/r02_gcc_ph16_*)
//   consume: item (jj, lane) = frame 8 w + 4 jj + {0, 2, 1, 3}[lane / 8], positions i and i + 8 with i = lane % 8: two
//            conflict-free LDS.128, six 8-byte stores; the two frames of a half-warp are 2 apart = 64 bytes in the A
//            operand, which with the 16-byte k-chunk skew makes the 16 stores of a half-warp hit 16 distinct slots
// The pieces a warp consumes are the ones it copied: cp.async.wait_group + __syncwarp, no block barrier for the input.
constexpr int H_POS = 16;                               // positions per chunk
constexpr int H_THREADS = GT_THREADS + 32;              // 8 producer warps + the MMA warp
constexpr int H_CHUNKS = fe2::PH_K / H_POS;             // 38
constexpr int H_A_LBO = (GT_ITEMS / 8) * 128 + 16;      // 6160
constexpr int H_A_BYTES = GT_KCH * H_A_LBO;             // 24 640
constexpr int H_RAW_BYTES = GT_FRAMES * H_POS * 16;     // 16 384
constexpr int H_OFF_B = GT_STAGES * H_A_BYTES;
constexpr int H_OFF_RAW = H_OFF_B + GT_B_SLOTS * GT_B_BYTES;
constexpr int H_SMEM = H_OFF_RAW + GT_RAW_STAGES * H_RAW_BYTES;   // 114 816 B: two CTAs per SM
static_assert(fe2::PH_K % H_POS == 0, "whole chunks");
// instruction descriptor: D = F32, A = B = F16, both K-major, N = 64, M = 128
constexpr uint32_t H_IDESC = (1u << 4) | (0u << 7) | (0u << 10) | ((64u >> 3) << 17) | ((128u >> 4) << 24);

__device__ __forceinline__ void h_issue_chunk_copy(unsigned char* smem, const uint4* __restrict__ ph, const uint4* __restrict__ btab,
                                                   long long f0, long long n_frames, int ch, int tid) {
    if (ch < H_CHUNKS) {
        const int w = tid >> 5, lane = tid & 31, pos = lane & 15;
        const uint32_t raw = smem_u32(smem + H_OFF_RAW + (ch % GT_RAW_STAGES) * H_RAW_BYTES) + pos * 16;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int f = 8 * w + 2 * j + (lane >> 4);
            const long long fr = f0 + f;
            const bool live = fr < n_frames;
            const uint4* src = live ? ph + fr * fe2::PH_K + ch * H_POS + pos : ph;
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(raw + f * 256), "l"(src), "r"(live ? 16 : 0) : "memory");
        }
        const uint32_t d = smem_u32(smem + H_OFF_B + (ch % GT_B_SLOTS) * GT_B_BYTES) + tid * 16;
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(btab + (size_t)ch * (GT_B_BYTES / 16) + tid) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
}

__device__ __forceinline__ uint32_t h2_mul(uint32_t a, uint32_t b) { uint32_t r; asm("mul.rn.f16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
__device__ __forceinline__ uint32_t h2_fma(uint32_t a, uint32_t b, uint32_t c) { uint32_t r; asm("fma.rn.f16x2 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
__device__ __forceinline__ uint32_t h2_neg(uint32_t a) { uint32_t r; asm("neg.f16x2 %0, %1;" : "=r"(r) : "r"(a)); return r; }

__global__ void __launch_bounds__(H_THREADS, 2)
gcc_ph16_kernel(const uint4* __restrict__ ph, long long n_frames, int T, const uint4* __restrict__ btab,
                const float* __restrict__ mean, const float* __restrict__ istd, float* __restrict__ out, OutStrides os) {
    extern __shared__ __align__(128) unsigned char smem[];      // A x2 | B x4 | raw x3
    __shared__ __align__(8) unsigned long long mbar[GT_STAGES];        // "empty": the MMAs that read A stage s (and its B slot) are done
    __shared__ __align__(8) unsigned long long mfull[GT_STAGES];       // "full" : the 8 producer warps have written their rows of A stage s
    __shared__ uint32_t tmem_base_s;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const long long f0 = (long long)blockIdx.x * GT_FRAMES;

    if (tid < GT_THREADS) {                                            // producer warps only (the MMA warp copies nothing)
        h_issue_chunk_copy(smem, ph, btab, f0, n_frames, 0, tid);
        h_issue_chunk_copy(smem, ph, btab, f0, n_frames, 1, tid);
    }

    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(GT_TMEM_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < GT_STAGES; ++s) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&mbar[s])));
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&mfull[s])), "n"(GT_THREADS / 32));
        }
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tmem = tmem_base_s;

    // consume role
    const int i = lane & 7, a = lane >> 3, fsub = ((a & 1) << 1) + (a >> 1);
    const int raw_off = (8 * warp + fsub) * 256 + i * 16;                                     // + jj * 4 * 256; position i + 8: + 128
    const int a_off = (i >> 1) * H_A_LBO + (i & 1) * 8 + (2 * warp) * 128 + (2 * fsub) * 16;  // + jj * 128, + (p & 1) * 16 + (p >> 1) * 2048

    if (warp == GT_THREADS / 32) {
        // ---- MMA warp: waits for a stage to be full, issues the chunk's six MMAs and commits them to the stage's "empty"
        // barrier.  The 8 producer warps never wait for each other or for the issue (a __syncthreads per chunk with
        // thread 0 issuing made every warp wait for warp 0 in every chunk: 16 % of the stall samples).
        for (int ch = 0; ch < H_CHUNKS; ++ch) {
            const int stg = ch & (GT_STAGES - 1);
            mbar_wait(smem_u32(&mfull[stg]), (uint32_t)((ch / GT_STAGES) & 1));
            if (lane == 0) {
                asm volatile("tcgen05.fence::after_thread_sync;");
                const uint32_t a0 = smem_u32(smem + stg * H_A_BYTES), b0 = smem_u32(smem + H_OFF_B + (ch % GT_B_SLOTS) * GT_B_BYTES);
#pragma unroll
                for (int ks = 0; ks < GT_KCH / 2; ++ks) {                   // K = 16 halves per MMA = 2 k-chunks
                    const uint64_t bd = umma_desc(b0 + ks * 2 * GT_B_LBO, GT_B_LBO, 128);
#pragma unroll
                    for (int mt = 0; mt < 3; ++mt) {
                        const uint64_t ad = umma_desc(a0 + ks * 2 * H_A_LBO + mt * 16 * 128, H_A_LBO, 128);
                        const uint32_t acc = (ch | ks) ? 1u : 0u;           // first MMA overwrites the accumulator
                        asm volatile(
                            "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                            "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                            ::"r"(tmem + mt * 64), "l"(ad), "l"(bd), "r"(H_IDESC), "r"(acc) : "memory");
                    }
                }
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&mbar[stg])) : "memory");
            }
            __syncwarp();
        }
    } else {
    for (int ch = 0; ch < H_CHUNKS; ++ch) {
        const int stg = ch & (GT_STAGES - 1);
        unsigned char* sA = smem + stg * H_A_BYTES + a_off;
        if (ch >= GT_STAGES) mbar_wait(smem_u32(&mbar[stg]), (uint32_t)(((ch - GT_STAGES) / GT_STAGES) & 1));
        h_issue_chunk_copy(smem, ph, btab, f0, n_frames, ch + 2, tid);
        asm volatile("cp.async.wait_group 2;" ::: "memory");            // this thread's pieces of chunk ch have landed ...
        __syncwarp();                                                   // ... and so have those of the other lanes of the warp

        const unsigned char* raw = smem + H_OFF_RAW + (ch % GT_RAW_STAGES) * H_RAW_BYTES + raw_off;
#pragma unroll
        for (int jj = 0; jj < 2; ++jj) {
            const uint4 v0 = *reinterpret_cast<const uint4*>(raw + jj * 1024);          // position i     : (re, im) x 4 channels
            const uint4 v1 = *reinterpret_cast<const uint4*>(raw + jj * 1024 + 128);    // position i + 8
            const uint32_t c0[4] = {v0.x, v0.y, v0.z, v0.w}, c1[4] = {v1.x, v1.y, v1.z, v1.w};
            uint32_t re[4], im[4];
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                re[c] = __byte_perm(c0[c], c1[c], 0x5410);      // (re p, re p+8)
                im[c] = __byte_perm(c0[c], c1[c], 0x7632);      // (im p, im p+8)
            }
            // a zero phasor (+0, +0; the front end writes exactly that) marks a vanishing channel: every pair it takes
            // part in has a vanishing cross spectrum, whose phase upstream is np.angle(0) = 0 -> (1, 0)
            const uint32_t mn = min(min(min(v0.x, v0.y), min(v0.z, v0.w)), min(min(v1.x, v1.y), min(v1.z, v1.w)));
            uint32_t pr[6], pi[6];
#pragma unroll
            for (int p = 0; p < 6; ++p) {
                constexpr int PM[6] = {0, 0, 0, 1, 1, 2}, PN[6] = {1, 2, 3, 2, 3, 3};
                const int m = PM[p], n = PN[p];
                pr[p] = h2_fma(re[m], re[n], h2_mul(im[m], im[n]));                 // Re conj(u_m) u_n
                pi[p] = h2_fma(re[m], im[n], h2_neg(h2_mul(im[m], re[n])));         // Im
            }
            if (__any_sync(0xffffffffu, mn == 0u)) {                                // rare (digital silence): warp-uniform branch
                uint32_t zm[4];
#pragma unroll
                for (int c = 0; c < 4; ++c) zm[c] = (c0[c] == 0u ? 0x0000ffffu : 0u) | (c1[c] == 0u ? 0xffff0000u : 0u);
#pragma unroll
                for (int p = 0; p < 6; ++p) {
                    constexpr int PM[6] = {0, 0, 0, 1, 1, 2}, PN[6] = {1, 2, 3, 2, 3, 3};
                    const uint32_t z = zm[PM[p]] | zm[PN[p]];
                    pr[p] = (pr[p] & ~z) | (0x3c003c00u & z);
                    pi[p] &= ~z;
                }
            }
            unsigned char* dst = sA + jj * 128;
#pragma unroll
            for (int p = 0; p < 6; ++p) *reinterpret_cast<uint2*>(dst + (p & 1) * 16 + (p >> 1) * 2048) = make_uint2(pr[p], pi[p]);
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy / cp.async writes -> async proxy (UMMA)
        __syncwarp();
        if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&mfull[stg])) : "memory");
    }
    }
    asm volatile("cp.async.wait_all;" ::: "memory");
    mbar_wait(smem_u32(&mbar[(H_CHUNKS - 1) & (GT_STAGES - 1)]), (uint32_t)(((H_CHUNKS - 1) / GT_STAGES) & 1));
    asm volatile("tcgen05.fence::after_thread_sync;");
    // The epilogue stages its tiles in the A-operand memory.  The mbarrier chain already orders every warp's last operand
    // stores before this point (full -> MMA -> commit -> the wait above); the block barrier states the same in a form
    // that compute-sanitizer's racecheck models (it reported the two writes as a hazard without it).  Once per CTA.
    __syncthreads();

    if (warp < GT_THREADS / 32) gcc_epilogue(tmem, smem, warp, lane, f0, n_frames, T, mean, istd, out, os);
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(GT_TMEM_COLS));
}

// FP16 B table: [H_CHUNKS][4 k-chunks][8 lag-groups][8 lags][8 halves], K order inside k-chunk q
// (re p, re p+8, im p, im p+8) for p = 2 q, then the same for p = 2 q + 1 (positions of the chunk)
static int get_gcc_btab16(const uint4** dev_tab) {
    static std::mutex mu;
    static uint4* cache[64] = {nullptr};
    int dev = 0;
    ADY_CUDA_CHECK(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64) return set_error(ADY_ERR_INVALID, "device index %d out of range", dev);
    std::lock_guard<std::mutex> lk(mu);
    if (!cache[dev]) {
        int bin_of_pos[fe2::PH_K];
        fe2::bins_of_phasor_pos(bin_of_pos);
        std::vector<__half> h((size_t)H_CHUNKS * GT_B_BYTES / 2, __float2half_rn(0.f));
        for (int ch = 0; ch < H_CHUNKS; ++ch)
            for (int kk = 0; kk < 2 * H_POS; ++kk) {              // K index within the chunk
                const int q = kk >> 3, e = kk & 7;
                const int pos = ch * H_POS + 2 * q + (e >> 2) + 8 * (e & 1), comp = (e >> 1) & 1;
                const int bin = bin_of_pos[pos];
                if (bin < 0) continue;
                const double ck = (bin == 0 || bin == 600) ? 0.5 : 1.0;    // x 2/N in the epilogue
                for (int n = 0; n < 64; ++n) {
                    const int lag = n - 32;                       // output order cc[-32:], cc[:32]
                    const double ang = 2.0 * M_PI * (double)((bin * (long long)((lag + NFFT) % NFFT)) % NFFT) / NFFT;
                    const double v = comp ? -ck * sin(ang) : ck * cos(ang);
                    const size_t off = (size_t)ch * (GT_B_BYTES / 2) + ((size_t)q * 8 + (n >> 3)) * 64 + (n & 7) * 8 + e;
                    h[off] = __double2half(v);
                }
            }
        uint4* d = nullptr;
        ADY_CUDA_CHECK(cudaMalloc(&d, h.size() * sizeof(__half)));
        ADY_CUDA_CHECK(cudaMemcpy(d, h.data(), h.size() * sizeof(__half), cudaMemcpyHostToDevice));
        cache[dev] = d;
    }
    *dev_tab = cache[dev];
    return ADY_OK;
}

static int launch_gcc_ph16(const void* in, int B, long long T, const float* mean, const float* istd, float* out, OutStrides os,
                           cudaStream_t stream) {
    const uint4* btab = nullptr;
    int rc = get_gcc_btab16(&btab);
    if (rc) return rc;
    const long long n_frames = (long long)B * T;
    const long long blocks = (n_frames + GT_FRAMES - 1) / GT_FRAMES;
    if (blocks > 0x7fffffffLL) return set_error(ADY_ERR_INVALID, "gcc: too many frames");
    static std::atomic<unsigned long long> configured{0};      // per device; the call is idempotent
    int dev = 0;
    ADY_CUDA_CHECK(cudaGetDevice(&dev));
    if (!((configured.load(std::memory_order_acquire) >> (dev & 63)) & 1ull)) {
        ADY_CUDA_CHECK(cudaFuncSetAttribute(gcc_ph16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, H_SMEM));
        configured.fetch_or(1ull << (dev & 63), std::memory_order_release);
    }
    gcc_ph16_kernel<<<(unsigned)blocks, H_THREADS, H_SMEM, stream>>>(reinterpret_cast<const uint4*>(in), n_frames, (int)T, btab, mean,
                                                                       istd, out, os);
    ADY_LAUNCH_CHECK("gcc_ph16_kernel");
    return ADY_OK;
}

int launch_gcc_from_stft(const float2* spec, int B, long long T, const float* mean, const float* istd, float* out,
                         OutStrides os, cudaStream_t stream) {
    return launch_gcc<false>(spec, B, T, mean, istd, out, os, stream);
}

// phasors: half2 x 4 channels per (frame, fe2 position), (B, T, 608) x 16 bytes, written by launch_features_mic_fe2
int launch_gcc_from_phasors(const void* phasors, int B, long long T, const float* mean, const float* istd, float* out,
                            OutStrides os, cudaStream_t stream) {
    // ADYOLO_GCC=tf32 selects the round-2a TF32 kernel (A/B measurements); the default is the FP16 kernel
    static const bool tf32 = [] { const char* e = getenv("ADYOLO_GCC"); return e && !strcmp(e, "tf32"); }();
    if (tf32) return launch_gcc<true>(phasors, B, T, mean, istd, out, os, stream);
    return launch_gcc_ph16(phasors, B, T, mean, istd, out, os, stream);
}

}  // namespace ady
