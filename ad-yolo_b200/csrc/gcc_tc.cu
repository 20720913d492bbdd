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
// One CTA = 64 frames = 384 items (row = pair * 64 + frame) = 3 UMMA M-tiles of 128 rows, N = 64
// lags, K = 1216 (608 bins).  Per chunk of 16 bins (K = 32):
//   - all threads take the 4 channel spectra of (64 frames x 16 bins) from registers (loaded one
//     chunk ahead), normalise each channel to unit modulus, form the 6 cross spectra and store them
//     as the A operand in shared memory in the canonical K-major, no-swizzle UMMA layout (core
//     matrix = 8 rows x 16 bytes); the B chunk (8 KB, precomputed in that layout) arrives by cp.async
//   - fence.proxy.async + barrier; one thread issues 4 (k-steps) x 3 (M-tiles) tcgen05.mma and a
//     tcgen05.commit on the stage's mbarrier.  Two stages: chunk c+1 is generated while the tensor
//     core consumes chunk c; a stage is only waited for when it is about to be overwritten
// Epilogue: all 8 warps tcgen05.ld (32x32b.x32) a 32-item x 32-lag block each, stage it through
// shared memory and store standardised 128-byte row segments.
// The MMA truncates FP32 operands to TF32, which biases a coherent peak low by ~6e-4; both operands
// are therefore rounded to nearest first: measured max error 2.9e-5 absolute against the float64
// oracle (gate 1e-3); all-(1,0) cross spectra (digital silence) give cc[0] = 1 exactly.
// B200, 128 x 5-s clips: 0.57 ms (CUDA-core kernel this replaces) -> 0.22-0.24 ms; at 512 clips the
// kernel streams the 1.97 GB of spectra at 2.7 TB/s (profiles/r01_ncu_gcc_tc.txt).
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <mutex>
#include <vector>

#include "common.cuh"
#include "frontend_core.cuh"
#include "frontend_host.h"

namespace ady {

struct OutStrides {
    long long sb, sc, st, sj;
};

constexpr int GT_FRAMES = 64;                 // frames per CTA
constexpr int GT_ITEMS = GT_FRAMES * 6;       // 384 rows = 3 M-tiles; row = pair * 64 + frame
constexpr int GT_BINS = 16;                   // bins per chunk -> K = 32 per chunk
constexpr int GT_CHUNKS = 38;                 // 608 bins >= 601
constexpr int GT_THREADS = 256;
constexpr int GT_STAGES = 2;
constexpr int GT_A_BYTES = (GT_BINS * 2 / 4) * (GT_ITEMS / 8) * 128;   // 8 k-chunks x 48 m-groups x 128 B = 49152
constexpr int GT_B_BYTES = (GT_BINS * 2 / 4) * (64 / 8) * 128;         // 8 x 8 x 128 = 8192
constexpr int GT_STAGE_BYTES = GT_A_BYTES + GT_B_BYTES;                // 56 KB
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

struct SpecRegs {
    float4 v01[4], v23[4];
};

__device__ __forceinline__ void load_chunk(SpecRegs& R, const float2* __restrict__ spec, long long f0, long long n_frames,
                                           int ch, int warp, int f_lo, int kb_lo) {
#pragma unroll
    for (int it = 0; it < 4; ++it) {
        const int combo = it * 8 + warp, f = (combo & 3) * 16 + f_lo, kb = (combo >> 2) * 2 + kb_lo;
        const int bin = ch * GT_BINS + kb;
        const long long fr = f0 + f;
        R.v01[it] = R.v23[it] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (fr < n_frames && bin < NBIN) {
            const float4* s = reinterpret_cast<const float4*>(spec + (fr * NBIN + bin) * 4);
            R.v01[it] = __ldg(s);
            R.v23[it] = __ldg(s + 1);
        }
    }
}

__global__ void __launch_bounds__(GT_THREADS, 2)
gcc_tc_kernel(const float2* __restrict__ spec, long long n_frames, int T, const float* __restrict__ btab,
              const float* __restrict__ mean, const float* __restrict__ istd, float* __restrict__ out, OutStrides os) {
    extern __shared__ __align__(128) unsigned char smem[];      // GT_STAGES x (A 48 KB | B 8 KB)
    __shared__ __align__(8) unsigned long long mbar[GT_STAGES];
    __shared__ uint32_t tmem_base_s;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const long long f0 = (long long)blockIdx.x * GT_FRAMES;

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

    // generation mapping: lane = (frame low 4 bits, bin parity); (warp, iteration) = (frame group, bin pair).
    // Rows are pair-major, so the 16 frames of a warp store 16 consecutive 16-byte rows: conflict-free.
    const int kb_lo = lane & 1, f_lo = lane >> 1;

    SpecRegs cur;
    load_chunk(cur, spec, f0, n_frames, 0, warp, f_lo, kb_lo);
    for (int ch = 0; ch < GT_CHUNKS; ++ch) {
        const int stg = ch & (GT_STAGES - 1);
        unsigned char* sA = smem + stg * GT_STAGE_BYTES;
        unsigned char* sB = sA + GT_A_BYTES;
        if (ch >= GT_STAGES) mbar_wait(smem_u32(&mbar[stg]), (uint32_t)(((ch - GT_STAGES) / GT_STAGES) & 1));

        {   // B chunk (canonical layout, TF32-rounded on the host): asynchronous copy, 2 x 16 B per thread
            const float4* src = reinterpret_cast<const float4*>(btab + (size_t)ch * (GT_B_BYTES / 4));
            const uint32_t d = smem_u32(sB) + tid * 16;
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(src + tid) : "memory");
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d + GT_THREADS * 16), "l"(src + tid + GT_THREADS) : "memory");
        }
        SpecRegs nxt;                                      // next chunk's spectra are in flight during this chunk's math
        if (ch + 1 < GT_CHUNKS) load_chunk(nxt, spec, f0, n_frames, ch + 1, warp, f_lo, kb_lo);
#pragma unroll
        for (int it = 0; it < 4; ++it) {
            const int combo = it * 8 + warp, f = (combo & 3) * 16 + f_lo, kb = (combo >> 2) * 2 + kb_lo;
            const float2 x[4] = {make_float2(cur.v01[it].x, cur.v01[it].y), make_float2(cur.v01[it].z, cur.v01[it].w),
                                 make_float2(cur.v23[it].x, cur.v23[it].y), make_float2(cur.v23[it].z, cur.v23[it].w)};
            // fast path: |x|^2 comfortably inside the FP32 range for all four channels
            float2 u[4];
            bool ok = true;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const float sq = fmaf(x[c].x, x[c].x, x[c].y * x[c].y);
                ok = ok && sq > 1e-30f && sq < 1e30f;
                const float r = rsqrt_ftz(sq);
                u[c] = make_float2(x[c].x * r, x[c].y * r);
            }
            bool z[4] = {false, false, false, false};
            if (!ok) {   // rare: silence, padding (rows / bins past the end load zeros), extreme magnitudes, NaN
#pragma unroll
                for (int c = 0; c < 4; ++c) u[c] = unit(x[c], z[c]);
            }
            // canonical layout: [k-chunk = kb/2][m-group = row/8] core matrices of 8 rows x 16 B
            unsigned char* dst = sA + ((kb >> 1) * (GT_ITEMS / 8) + (f >> 3)) * 128 + (f & 7) * 16 + (kb & 1) * 8;
            int p = 0;
#pragma unroll
            for (int m = 0; m < 4; ++m)
#pragma unroll
                for (int n = m + 1; n < 4; ++n) {
                    float re = u[m].x * u[n].x + u[m].y * u[n].y, im = u[m].x * u[n].y - u[m].y * u[n].x;   // conj(u_m) u_n
                    if (!ok && (z[m] || z[n])) { re = 1.f; im = 0.f; }
                    *reinterpret_cast<float2*>(dst + p * (GT_FRAMES / 8) * 128) = make_float2(tf32_rn_bits(re), tf32_rn_bits(im));   // row = p * 64 + f
                    ++p;
                }
        }
        cur = nxt;
        asm volatile("cp.async.wait_all;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> async proxy (UMMA)
        __syncthreads();
        if (tid == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;");
            const uint32_t a0 = smem_u32(sA), b0 = smem_u32(sB);
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {                            // K = 8 per MMA = 2 k-chunks
                const uint64_t bd = umma_desc(b0 + ks * 2 * (64 / 8) * 128, (64 / 8) * 128, 128);
#pragma unroll
                for (int mt = 0; mt < 3; ++mt) {
                    const uint64_t ad = umma_desc(a0 + ks * 2 * (GT_ITEMS / 8) * 128 + mt * 16 * 128, (GT_ITEMS / 8) * 128, 128);
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
    // the last commit covers every MMA issued before it
    mbar_wait(smem_u32(&mbar[(GT_CHUNKS - 1) & (GT_STAGES - 1)]), (uint32_t)(((GT_CHUNKS - 1) / GT_STAGES) & 1));
    asm volatile("tcgen05.fence::after_thread_sync;");

    // ---- epilogue.  A warp may read TMEM lanes 32 (w % 4) .. +31: warps w and w + 4 share a lane
    // quarter and split the 64 lag columns.  TMEM lane = row of the M-tile = pair (mt*2 + q/2), frame
    // (q & 1) * 32 + lane.  Natural output layout: the 32 x 32 block goes through shared memory so
    // that global stores are 128-byte row segments; any other stride set is stored directly.
    {
        constexpr float kScale = 2.0f / NFFT;               // B holds cos / -sin (x 1/2 at bins 0 and 600): exact at lag 0
        const int q = warp & 3, h = warp >> 2;
        const bool vec = os.sj == 1 && ((os.sb | os.sc | os.st) & 3) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0;
        float* stage = reinterpret_cast<float*>(smem) + warp * (32 * 36);      // 32 rows, 144-byte pitch: conflict-free
        long long off[8];                                                     // vec: this lane's 8 (row, 16-byte column) targets
#pragma unroll
        for (int it = 0; it < 8; ++it) {
            const int f = (q & 1) * 32 + (vec ? 4 * it + (lane >> 3) : lane);
            const long long fr = f0 + f;
            const long long b = fr / T;
            off[it] = fr < n_frames ? b * os.sb + (fr - b * T) * os.st + (vec ? h * 32 + (lane & 7) * 4 : 0) : -1;
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
            const int p = mt * 2 + (q >> 1);
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
                    if (off[it] >= 0) *reinterpret_cast<float4*>(out + off[it] + p * os.sc) = v;
                }
                __syncwarp();
            } else if (off[0] >= 0) {
                float* o = out + off[0] + p * os.sc + (long long)(h * 32) * os.sj;
#pragma unroll
                for (int j = 0; j < 32; ++j)
                    o[j * os.sj] = (__uint_as_float(r[j]) * kScale - (mu ? mu[j] : 0.f)) * (is ? is[j] : 1.f);
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(GT_TMEM_COLS));
}

// B operand table: [38 chunks][8 k-chunks][8 lag-groups][8 lags][4 k'] floats (canonical K-major layout)
static int get_gcc_btab(const float** dev_tab) {
    static std::mutex mu;
    static float* cache[64] = {nullptr};
    int dev = 0;
    ADY_CUDA_CHECK(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64) return set_error(ADY_ERR_INVALID, "device index %d out of range", dev);
    std::lock_guard<std::mutex> lk(mu);
    if (!cache[dev]) {
        std::vector<float> h((size_t)GT_CHUNKS * GT_B_BYTES / 4, 0.f);
        for (int ch = 0; ch < GT_CHUNKS; ++ch)
            for (int kp = 0; kp < 32; ++kp) {                     // k' within the chunk
                const int bin = ch * GT_BINS + (kp >> 1);
                if (bin >= NBIN) continue;
                const double ck = (bin == 0 || bin == 600) ? 0.5 : 1.0;    // x 2/N in the epilogue
                for (int n = 0; n < 64; ++n) {
                    const int lag = n - 32;                       // output order cc[-32:], cc[:32]
                    const double ang = 2.0 * M_PI * (double)((bin * (long long)((lag + NFFT) % NFFT)) % NFFT) / NFFT;
                    const double v = (kp & 1) ? -ck * sin(ang) : ck * cos(ang);
                    const size_t off = (size_t)ch * (GT_B_BYTES / 4) + ((size_t)(kp >> 2) * 8 + (n >> 3)) * 32 + (n & 7) * 4 + (kp & 3);
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
        cache[dev] = d;
    }
    *dev_tab = cache[dev];
    return ADY_OK;
}

int launch_gcc_from_stft(const float2* spec, int B, long long T, const float* mean, const float* istd, float* out,
                         OutStrides os, cudaStream_t stream) {
    const float* btab = nullptr;
    int rc = get_gcc_btab(&btab);
    if (rc) return rc;
    const long long n_frames = (long long)B * T;
    const long long blocks = (n_frames + GT_FRAMES - 1) / GT_FRAMES;
    if (blocks > 0x7fffffffLL) return set_error(ADY_ERR_INVALID, "gcc: too many frames");
    static int configured_dev = -1;
    int dev = 0;
    ADY_CUDA_CHECK(cudaGetDevice(&dev));
    const int shmem = GT_STAGES * GT_STAGE_BYTES;
    if (configured_dev != dev) {
        ADY_CUDA_CHECK(cudaFuncSetAttribute(gcc_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, shmem));
        configured_dev = dev;
    }
    gcc_tc_kernel<<<(unsigned)blocks, GT_THREADS, shmem, stream>>>(spec, n_frames, (int)T, btab, mean, istd, out, os);
    ADY_LAUNCH_CHECK("gcc_tc_kernel");
    return ADY_OK;
}

}  // namespace ady
