// Fused SELD front end for FOA clips on sm_100a, second generation: int16 PCM -> standardised log-mel +
// intensity vectors in one persistent kernel (+ the sparse top_db fix-up pass of frontend.cu).
//
// Reference behaviour replaced: /root/reference/src/datasets.py:147 (normalisation), :252-292
// (get_stft_spectrogram, get_logmel_spectrogram, get_melscale_foa_intensity_vectors, get_feature) and the
// (C,T,F) stacking of :158-160; == src/utils/utility.py:142-215.  The maths and the per-thread building blocks
// are in fe2_core.cuh; this file is the CTA-level choreography.
//
// One CTA per SM runs GROUPS = 4 independent tile pipelines ("groups" of 160 threads with their own staging + frame buffers
// and their own named barrier) next to ONE shared-memory copy of the read-only tables.  Per group and tile of 2 frames
// of one clip:
//   global int16 (N,4) --cp.async(8 B / sample)--> 24 rows x 75 samples, columns permuted so that stage A is
//                                                  bank-conflict free (3 hops, shared by the 2 frames)
//   stage A  75 tasks / frame : window + DFT-16 of both packed FFTs (f32x2)          -> X1 (16 B / point)
//   stage B  80 tasks / frame : DFT-15, in place                                      -> X2
//   stage C 121 tasks / frame : twiddle + 2 x DFT-5, channel split, |X|^2, I / E, in place -> V  (242 pair-tasks over two rounds)
//   mel     156 lane-jobs     : <= 9 non-zeros each, both frames per entry            -> 64-byte partial records
//   epilogue 128 threads      : (frame, mel): add the filter's records, 10 log10, standardise, coalesced store
// The copy of the next tile is issued right after stage A and overlaps everything else.
#include <atomic>
#include <mutex>

#include "common.cuh"
#include "fe2_core.cuh"
#include "frontend_host.h"

namespace ady {
int build_fe2_tables_host(fe2::Tables* t);   // tables.cu
namespace fe2 {

constexpr int CTAS_PER_SM = ADY_FE2_CTAS;   // x GROUPS tile pipelines each (fe2_core.cuh)
constexpr int NTB = GROUPS * NT;            // threads per CTA

// barrier of one group (bar 0 = __syncthreads is the whole CTA).  The barrier number is re-derived from %tid.x at every
// use: kept live across the tile loop it is the first value ptxas spills, and with 227 KB of the SM's unified storage
// carved out as shared memory a spill reload is mostly an L2 round trip in front of the barrier.
__device__ __forceinline__ void group_sync() {
    if (GROUPS == 1) __syncthreads();
    else {
        unsigned t;
        asm volatile("mov.u32 %0, %%tid.x;" : "=r"(t));
        asm volatile("bar.sync %0, %1;" ::"r"(t / NT + 1), "n"(NT) : "memory");
    }
}

__device__ __forceinline__ void cp_async8(void* smem_dst, const void* gsrc) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(s), "l"(gsrc));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;\n" ::: "memory"); }

// Stage the 3 hops of frames t0, t0+1 of the clip starting at sample `clip0` of `audio` (see stage_col()).
// thread (h = tid / 80, rem = tid % 80 < 75) copies column rem of rows 12 h .. 12 h + 11.
// Interior tiles (two frames, no reflect padding: all but the first and possibly the last tile of a clip) take the
// fast path: one 64-bit source address per thread and 12 copies with immediate offsets on both sides.  The general
// path needs ~10 integer instructions per copy (row guard, 64-bit reflect select, 64-bit address), which was 11 % of
// all instructions of the kernel (ncu source page, round 2).
template <int I>
__device__ __forceinline__ void cp_async8_imm(unsigned sdst, const void* gsrc) {
    asm volatile("cp.async.ca.shared.global [%0+%2], [%1+%3], 8;\n" ::"r"(sdst), "l"(gsrc), "n"(I * ROWP * 8), "n"(I * 75 * 8));
}
__device__ __forceinline__ void issue_tile_copy(unsigned char* samp, const int16_t* __restrict__ audio, long long clip0, int t0, int nf,
                                                int tid, int dst_off /* (12 h * ROWP + col_perm[stage_col(rem)]) * 8 or -1 */) {
    if (dst_off < 0) return;
    const int h = tid >= 80, rem = tid - 80 * h;        // (dst_off < 0 for tid >= NT_AB)
    const int16_t* clip = audio + clip0 * 4;
    if (t0 > 0 && nf == TFR) {
        const int16_t* src = clip + (600LL * (t0 - 1) + 900 * h + rem) * 4;
        const unsigned d = (unsigned)__cvta_generic_to_shared(samp + dst_off);
        cp_async8_imm<0>(d, src); cp_async8_imm<1>(d, src); cp_async8_imm<2>(d, src);  cp_async8_imm<3>(d, src);
        cp_async8_imm<4>(d, src); cp_async8_imm<5>(d, src); cp_async8_imm<6>(d, src);  cp_async8_imm<7>(d, src);
        cp_async8_imm<8>(d, src); cp_async8_imm<9>(d, src); cp_async8_imm<10>(d, src); cp_async8_imm<11>(d, src);
        return;
    }
    const int nrows = 8 * (nf + 1) - 12 * h;                      // rows of this half that the tile needs
    long long s = 600LL * (t0 - 1) + 900 * h + rem;               // clip sample of row 12 h
    unsigned char* dst = samp + dst_off;
#pragma unroll 1
    for (int i = 0; i < 12; ++i) {
        if (i < nrows) {
            const long long m = s < 0 ? -s : s;                   // reflect padding (frame 0 only)
            cp_async8(dst + i * (ROWP * 8), clip + m * 4);
        }
        s += 75;
    }
}

// tile -> (clip, first frame) without an integer division per tile and thread: magic = min(floor(2^32 / tiles_per_clip),
// 2^32 - 1) undershoots the quotient by at most one for tile < 2^31
__device__ __forceinline__ void tile_coords(int tile, int tiles_per_clip, unsigned magic, int& b, int& t0) {
    int q = (int)__umulhi((unsigned)tile, magic), r = tile - q * tiles_per_clip;
    if (r >= tiles_per_clip) { ++q; r -= tiles_per_clip; }
    b = q;
    t0 = r * TFR;
}

// ROT : fused rotation augmentation (utils/augmentations.py:46-111), one combination per clip.
// VIEW: clip b starts at sample clip_off[b] of one resident buffer (on-the-fly chunking, preprocess.py:13-48).
// MIC : microphone-array format (no counterpart in the reference): the 4 log-mel channels go into a
//       (B, 10, T, 64) tensor and stage C writes one half2 unit phasor per channel and bin to `phasor`
//       (B, T, 608) x 16 bytes for the GCC-PHAT lag transform (gcc_tc.cu); no intensity vectors (ROT must be false).
template <bool ROT, bool VIEW, bool MIC>
__global__ void __launch_bounds__(NTB, CTAS_PER_SM)
fe2_foa_kernel(const int16_t* __restrict__ audio, long long N, int T, int tiles_per_clip, unsigned tile_magic, int ntiles,
               const Tables* __restrict__ tab, const float* __restrict__ mean, const float* __restrict__ istd,
               float dc0, float dc1, const int8_t* __restrict__ rot, const long long* __restrict__ clip_off,
               float* __restrict__ out, uint4* __restrict__ phasor, int* __restrict__ flags, uint32_t* __restrict__ kext, int B) {
    constexpr int NCH = MIC ? 10 : 7;
    extern __shared__ __align__(16) unsigned char smem[];
    const int g = GROUPS == 1 ? 0 : (int)threadIdx.x / NT;                          // group (warp-uniform: NT is a multiple of 32)
    const int tid = (int)threadIdx.x - g * NT;                                      // thread within the group
    unsigned char* s_samp = smem + SmemLayout::off_group + g * SmemLayout::group_bytes;
    unsigned char* s_x = s_samp + SAMP_BYTES;
    MelEnt* s_ent = reinterpret_cast<MelEnt*>(smem + SmemLayout::off_ent);           // (only with TABLES_IN_SMEM)
    const unsigned char* s_tw = nullptr;                                             // stage-C twiddles: constant memory on the device (c_tw75)
    float* s_win = reinterpret_cast<float*>(smem + SmemLayout::off_win);
    float2* s_scale = reinterpret_cast<float2*>(smem + SmemLayout::off_scale);       // (only with scale_in_smem)
    uint8_t* s_meljobs = smem + SmemLayout::off_meljobs;

    // fixed roles
    const bool ab = tid < NT_AB;                                                    // warps 0..4 run stages A / B and the copies
    const int fA = tid >= 80, lA = tid - 80 * fA;                                  // stages A / B: frame, lane
    const StageAConst ka = stage_a_const(ab && lA < 75 ? lA : 0, ab && lA < 75 ? tab->col_perm[lA] : 0);
    const int copy_dst = ab && lA < 75 ? (12 * fA * ROWP + tab->col_perm[stage_col(lA)]) * 8 : -1;

    const int tile_step = (int)gridDim.x * GROUPS;
    int tile = (int)blockIdx.x * GROUPS + g;
    if (tile < ntiles) {
        int b, t0;
        tile_coords(tile, tiles_per_clip, tile_magic, b, t0);
        issue_tile_copy(s_samp, audio, VIEW ? clip_off[b] : (long long)b * N, t0, min(TFR, T - t0), tid, copy_dst);
    }
    cp_async_commit();
    // constant tables -> smem (once per persistent CTA, by all its groups)
    {
        const int t = (int)threadIdx.x;
        if (TABLES_IN_SMEM) {
            const uint2* src = reinterpret_cast<const uint2*>(tab->ent);
            uint2* dst = reinterpret_cast<uint2*>(s_ent);
            for (int i = t; i < MEL_L * NJOBS; i += NTB) dst[i] = src[i];
            for (int i = t; i < 16 * WIN_P; i += NTB) s_win[i] = tab->win[i];
        }
        if (SmemLayout::scale_in_smem) {
            for (int i = t; i < 7 * NMEL; i += NTB) {   // standardisation as one FMA: x * is + (-mu * is)
                const float mu = mean ? mean[i] : 0.f, is = istd ? istd[i] : 1.f;
                s_scale[i] = make_float2(is, -mu * is);
            }
        }
        if (t < NMEL) s_meljobs[t] = tab->mel_njobs[t];
        if (t <= REC_MAXJOBS) reinterpret_cast<int*>(s_meljobs + NMEL)[t] = tab->rec_off[t] * 16;
    }
    if (GROUPS > 1) __syncthreads();                                                // tables visible to every group
    const int rec_off = tid * 16;                                                   // this lane-job's record slot

    for (; tile < ntiles; tile += tile_step) {
        int b, t0;
        tile_coords(tile, tiles_per_clip, tile_magic, b, t0);
        const int nf = min(TFR, T - t0);
        const unsigned rb = ROT ? rot_bits_rt(rot[b]) : 0u;
        cp_async_wait_all();
        group_sync();                                   // samples landed; previous tile's epilogue is done with X

        // ---- stage A
        {
            // the 32 per-sample offsets of stage A depend on loop-invariant lane constants only; left alone, ptxas computes
            // them once and keeps them in local memory across the tile loop (33 spill loads per task).  An opaque copy of
            // the row rotation per tile keeps the four integer instructions per sample in the loop instead.
            StageAConst kt = ka;
            asm volatile("" : "+r"(kt.cr));
            if (ab && lA < 75 && fA < nf) stage_a(s_samp, TABLES_IN_SMEM ? s_win : tab->win, s_x, fA, lA, kt);
        }
        group_sync();

        // samples are consumed: prefetch the next tile while the rest of this one runs
        {
            const int nt = tile + tile_step;
            if (nt < ntiles) {
                int nb, nt0;
                tile_coords(nt, tiles_per_clip, tile_magic, nb, nt0);
                issue_tile_copy(s_samp, audio, VIEW ? clip_off[nb] : (long long)nb * N, nt0, min(TFR, T - nt0), tid, copy_dst);
            }
            cp_async_commit();
        }

        // ---- stage B (in place)
        if (ab && fA < nf) stage_b(s_x, fA, lA);
        group_sync();

        // ---- stage C (in place): 224 regular pair-tasks + 18 c = 0 pair-tasks over two rounds of NT threads; the
        // c = 0 tasks run in round 1 on the last warp, next to the tail of the regular tasks on the first warps
#pragma unroll 1
        for (int rd = 0; rd < 2; ++rd) {
            const int slot = rd * NT + tid;
            if (slot < 2 * NREG) {
                const int f = slot >= NREG, task = slot - f * NREG;
                if (f < nf) {
                    if (MIC) stage_c_mic<false>(s_x + f * X_STRIDE, s_tw, task, dc0, dc1, phasor + ((long long)b * T + t0 + f) * PH_K);
                    else stage_c_foa<false>(s_x + f * X_STRIDE, s_tw, task, dc0, dc1, rb);
                }
            } else if (rd == 1 && tid >= NT - 32 && tid < NT - 32 + 2 * NC0) {
                const int i = tid - (NT - 32), f = i >= NC0, task = i - f * NC0;
                if (f < nf) {
                    if (MIC) stage_c_mic<true>(s_x + f * X_STRIDE, s_tw, task, dc0, dc1, phasor + ((long long)b * T + t0 + f) * PH_K);
                    else stage_c_foa<true>(s_x + f * X_STRIDE, s_tw, task, dc0, dc1, rb);
                }
            }
        }
        group_sync();

        // ---- mel projection: lane-job tid, both frames
        f2 acc[TFR][4];
        if (tid < NJOBS) mel_job<!MIC>(s_x, (TABLES_IN_SMEM ? s_ent : tab->ent) + tid, acc);
        group_sync();                                   // every V read is done -> records may overwrite the frame buffers
        if (tid < NJOBS) {
#pragma unroll
            for (int f = 0; f < TFR; ++f) {
                st_f4(s_x + (2 * f) * REC_PLANE + rec_off, lo2(acc[f][0]), hi2(acc[f][0]), lo2(acc[f][1]), hi2(acc[f][1]));
                if (!MIC) st_f4(s_x + (2 * f + 1) * REC_PLANE + rec_off, lo2(acc[f][2]), hi2(acc[f][2]), lo2(acc[f][3]), hi2(acc[f][3]));
            }
        }
        group_sync();

        // ---- epilogue: thread = (frame, mel)
        if (tid < TFR * NMEL && (tid >> 6) < nf) {
            const int f = tid >> 6, j = tid & 63;
            // running extrema of this clip, requested before the record sums so that the L2 round trip is over when they
            // are compared (volatile: the loads must not sink to their use)
            // Only for long clips (> 600 tiles = 30 s): there every tile of the launch queues on the same few addresses
            // (8 x 60 s: 0.230 -> 0.178 ms), while for short clips the eight extra loads cost more than the atomics they
            // save (256 x 5 s: +1.4 %); without the filter the comparison values let every atomic through.
            uint32_t curmx[4] = {0u, 0u, 0u, 0u}, curmn[4] = {~0u, ~0u, ~0u, ~0u};
            if (tiles_per_clip > 600) {
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    asm volatile("ld.global.cg.u32 %0, [%1];" : "=r"(curmx[c]) : "l"(kext + b * 4 + c));
                    asm volatile("ld.global.cg.u32 %0, [%1];" : "=r"(curmn[c]) : "l"(kext + (B + b) * 4 + c));
                }
            }
            const int nq = s_meljobs[j];
            const int* rec_off = reinterpret_cast<const int*>(s_meljobs + NMEL);
            const unsigned char* ra = s_x + (2 * f) * REC_PLANE + j * 16;
            float4 pa = make_float4(0.f, 0.f, 0.f, 0.f), pb = pa;
            for (int i = 0; i < nq; ++i) {                  // fixed order: the result does not depend on scheduling
                const unsigned char* rp = ra + rec_off[i];  // chunk-major slots: the quarter-warp reads 8 consecutive records
                const float4 a = *reinterpret_cast<const float4*>(rp);
                pa.x += a.x; pa.y += a.y; pa.z += a.z; pa.w += a.w;
                if (!MIC) {
                    const float4 c = *reinterpret_cast<const float4*>(rp + REC_PLANE);
                    pb.x += c.x; pb.y += c.y; pb.z += c.z;
                }
            }
            // record order (|W|^2, |Z|^2, |Y|^2, |X|^2 | I_Y, I_Z, I_X) -> channels mel W,Y,Z,X, iv Y,Z,X
            float v[7] = {power_to_db(pa.x), power_to_db(pa.z), power_to_db(pa.y), power_to_db(pa.w), pb.x, pb.y, pb.z};
            if (!MIC && !(pb.x == pb.x && pb.y == pb.y && pb.z == pb.z)) atomicOr(flags, 1);   // datasets.py:277
            if (ROT) {
                if (rb & 1u) v[4] = -v[4];
                if (rb & 2u) v[5] = -v[5];
                if (rb & 4u) v[6] = -v[6];
                if (rb & 8u) { float t = v[1]; v[1] = v[3]; v[3] = t; t = v[4]; v[4] = v[6]; v[6] = t; }   // X <-> Y
            }
            // per-(clip, log-mel channel) extrema of the un-clamped dB values for the top_db pass (power_to_db's global
            // max, datasets.py:265): one warp-wide integer reduction per channel on order-preserving keys
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const uint32_t key = f2key(v[c]);
                const uint32_t hi = __reduce_max_sync(0xffffffffu, key), lo = __reduce_min_sync(0xffffffffu, key);
                // the running extrema only ever move outwards, so a (possibly stale) read that already covers this warp's
                // values makes the atomic redundant: with few long clips every tile of the launch used to queue on the same
                // 8 addresses per clip (4 800 same-address atomics per 60-s clip serialise into ~0.2 ms in L2)
                if ((tid & 31) == 0) {
                    if (hi > curmx[c]) atomicMax(kext + b * 4 + c, hi);
                    if (lo < curmn[c]) atomicMin(kext + (B + b) * 4 + c, lo);
                }
            }
            float* o = out + (((long long)b * NCH) * T + t0 + f) * NMEL + j;
            const long long T64 = (long long)T * NMEL;
#pragma unroll
            for (int c = 0; c < (MIC ? 4 : 7); ++c) {
                float2 k;
                if (SmemLayout::scale_in_smem) k = s_scale[c * NMEL + j];
                else {
                    const float mu = mean ? mean[c * NMEL + j] : 0.f, is = istd ? istd[c * NMEL + j] : 1.f;
                    k = make_float2(is, -mu * is);
                }
                o[c * T64] = fmaf(v[c], k.x, k.y);
            }
        }
        // the barrier at the top of the next iteration orders these record reads before its stage A
    }
    cp_async_wait_all();
}

// ------------------------------------------------------------------------------------------------ host side
static int get_fe2_tables(const Tables** dev_tables) {
    static std::mutex mu;
    static Tables* cache[64] = {nullptr};
    int dev = 0;
    ADY_CUDA_CHECK(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64) return set_error(ADY_ERR_INVALID, "device index %d out of range", dev);
    std::lock_guard<std::mutex> lk(mu);
    if (!cache[dev]) {
        static Tables host;
        static bool host_ok = false;
        if (!host_ok) {
            int rc = build_fe2_tables_host(&host);
            if (rc) return rc;
            host_ok = true;
        }
        Tables* d = nullptr;
        ADY_CUDA_CHECK(cudaMalloc(&d, sizeof(Tables)));
        ADY_CUDA_CHECK(cudaMemcpy(d, &host, sizeof(Tables), cudaMemcpyHostToDevice));
        ADY_CUDA_CHECK(cudaMemcpyToSymbol(c_tw75, host.tw75, sizeof(host.tw75)));     // this translation unit's copy, per device
        cache[dev] = d;
    }
    *dev_tables = cache[dev];
    return ADY_OK;
}

template <bool ROT, bool VIEW, bool MIC>
static int launch_inst(int grid, cudaStream_t stream, const int16_t* audio, long long N, int T, int tpc, int ntiles, const Tables* tab,
                       const float* mean, const float* istd, float dc0, float dc1, const int8_t* rot, const long long* clip_off,
                       float* out, uint4* phasor, int* flags, int B) {
    // once per (instantiation, device); a racing second thread at worst repeats the idempotent call
    static std::atomic<unsigned long long> configured{0};
    int dev = 0;
    ADY_CUDA_CHECK(cudaGetDevice(&dev));
    if (!((configured.load(std::memory_order_acquire) >> (dev & 63)) & 1ull)) {
        ADY_CUDA_CHECK(cudaFuncSetAttribute(fe2_foa_kernel<ROT, VIEW, MIC>, cudaFuncAttributeMaxDynamicSharedMemorySize, SmemLayout::total));
        ADY_CUDA_CHECK(cudaFuncSetAttribute(fe2_foa_kernel<ROT, VIEW, MIC>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
        configured.fetch_or(1ull << (dev & 63), std::memory_order_release);
    }
    const unsigned long long m64 = 0x100000000ull / (unsigned long long)tpc;
    const unsigned magic = m64 > 0xffffffffull ? 0xffffffffu : (unsigned)m64;
    fe2_foa_kernel<ROT, VIEW, MIC><<<grid, NTB, SmemLayout::total, stream>>>(audio, N, T, tpc, magic, ntiles, tab, mean, istd, dc0, dc1, rot,
                                                                          clip_off, out, phasor, flags,
                                                                          reinterpret_cast<uint32_t*>(flags) + 16, B);
    ADY_LAUNCH_CHECK("fe2_foa_kernel");
    return ADY_OK;
}

}  // namespace fe2

static int launch_fe2(const int16_t* audio, int B, long long N, const float* mean, const float* istd, float dc_offset,
                      const int8_t* rot, const long long* clip_off, float* out, uint4* phasor, void* ws, cudaStream_t stream) {
    using namespace fe2;
    const long long T = N / HOP;
    if (B <= 0 || T <= 0) return set_error(ADY_ERR_INVALID, "features: need B>0 and at least %d samples", HOP);
    if (N <= HOP) return set_error(ADY_ERR_INVALID, "features: reflect padding needs N > %d samples", HOP);
    const Tables* tab = nullptr;
    int rc = get_fe2_tables(&tab);
    if (rc) return rc;
    const long long tpc = (T + TFR - 1) / TFR;
    const long long ntiles = (long long)B * tpc;
    if (ntiles > 0x7fffffffLL) return set_error(ADY_ERR_INVALID, "features: too many tiles");
    // workspace: [flags | pad to 64 B] [max keys (B,4)] [min keys (B,4)]  (frontend_workspace_bytes)
    int* flags = reinterpret_cast<int*>(ws);
    ADY_CUDA_CHECK(cudaMemsetAsync(ws, 0, 64 + (size_t)B * 16, stream));
    ADY_CUDA_CHECK(cudaMemsetAsync(reinterpret_cast<char*>(ws) + 64 + (size_t)B * 16, 0xff, (size_t)B * 16, stream));
    int dev = 0, sms = 0;
    ADY_CUDA_CHECK(cudaGetDevice(&dev));
    ADY_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const long long want = (ntiles + GROUPS - 1) / GROUPS;                          // one tile per group to start with
    const int grid = (int)(want < (long long)CTAS_PER_SM * sms ? want : (long long)CTAS_PER_SM * sms);
    // window scale: 2^-15 (int16 -> [-1,1)) * 1/2 (channel split), DC terms scaled by the same 1/2
    const float dc0 = dc_offset * 300.0f, dc1 = -dc_offset * 150.0f;
#define ADY_FE2_GO(R, V, M) return launch_inst<R, V, M>(grid, stream, audio, N, (int)T, (int)tpc, (int)ntiles, tab, mean, istd, dc0, dc1, rot, clip_off, out, phasor, flags, B)
    if (phasor) ADY_FE2_GO(false, false, true);
    if (rot && clip_off) ADY_FE2_GO(true, true, false);
    if (rot) ADY_FE2_GO(true, false, false);
    if (clip_off) ADY_FE2_GO(false, true, false);
    ADY_FE2_GO(false, false, false);
#undef ADY_FE2_GO
}

int launch_features_foa_fe2(const int16_t* audio, int B, long long N, const float* mean, const float* istd, float dc_offset,
                            const int8_t* rot, const long long* clip_off, float* out, void* ws, cudaStream_t stream) {
    return launch_fe2(audio, B, N, mean, istd, dc_offset, rot, clip_off, out, nullptr, ws, stream);
}

// MIC format, first half: 4 log-mel channels of out (B, 10, T, 64) + unit phasors (B, T, 608) x 16 bytes
int launch_features_mic_fe2(const int16_t* audio, int B, long long N, const float* mean, const float* istd, float dc_offset,
                            float* out, void* phasor, void* ws, cudaStream_t stream) {
    if (!phasor) return set_error(ADY_ERR_INVALID, "features_mic: phasor scratch is NULL");
    return launch_fe2(audio, B, N, mean, istd, dc_offset, nullptr, nullptr, out, reinterpret_cast<uint4*>(phasor), ws, stream);
}

}  // namespace ady
