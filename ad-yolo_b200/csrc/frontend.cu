// Fused SELD front end for FOA clips on sm_100a: int16 PCM -> standardised log-mel + intensity
// vectors, one persistent kernel + a sparse top_db fix-up pass.
//
// Reference behaviour replaced: /root/reference/src/datasets.py:147 (normalisation),
// :252-292 (get_stft_spectrogram, get_logmel_spectrogram, get_melscale_foa_intensity_vectors,
// get_feature) and the (C,T,F) stacking of :158-160; == src/utils/utility.py:142-215.
//
// Data flow per tile of TF=3 frames of one clip (see frontend_core.cuh for the maths):
//   global int16 (N,4) --cp.async(4B)--> skewed, de-interleaved sample planes in smem
//   stage 1 (25 x DFT-48 / packed FFT) -> smem exchange -> stage 2 (DFT-25 pairs, channel split,
//   |X|^2, IV) -> smem V -> sparse mel, log10, standardise -> global (B,7,T,64) f32
// The per-(clip,channel) global max needed by librosa.power_to_db(top_db=80) (datasets.py:265)
// is taken by clamp_topdb_kernel over the freshly written (L2-resident) log-mel planes, which
// then rewrites only the rows that fall below max-80 dB.
#include <stdlib.h>

#include "common.cuh"
#include "frontend_core.cuh"
#include "frontend_host.h"

namespace ady {

// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void cp_async4(void* smem_dst, const void* gsrc) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(s), "l"(gsrc));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;\n" ::: "memory"); }

// Stage the samples of frames t0 .. t0+nf-1 of clip b: plane (frame, pair), word = skew(idx).
// Frame t covers padded positions [600 t, 600 t + 1200) of the reflect-padded clip
// (librosa.stft center=True, pad 600): y index = 600 t - 600 + idx, mirrored for t = 0, idx < 600.
__device__ __forceinline__ void issue_tile_copy(uint32_t* samples, const int16_t* __restrict__ audio,
                                                long long clip_base, int t0, int nf) {
    const int tid = threadIdx.x;
    if (tid >= COPY_THREADS) return;
    // element (idx = i0 + 80 i, pair), i = 0..14: lanes 0-15 of a warp copy the (W,Y) words of 16
    // consecutive samples, lanes 16-31 the (Z,X) words of the same samples (one 128-byte line)
    const int pair = (tid >> 4) & 1, i0 = (tid >> 5) * 16 + (tid & 15);
    const int16_t* clip = audio + clip_base * 4 + pair * 2;   // clip_base = first sample of the clip
    for (int f = 0; f < nf; ++f) {
        const int t = t0 + f;
        uint32_t* dst = samples + splane_base(q_of_g(2 * f + pair)) + skew(i0);  // skew(i0 + 80 i) = skew(i0) + 85 i
        if (t > 0) {
            const int16_t* src = clip + ((long long)(t - 1) * HOP + i0) * 4;
#pragma unroll
            for (int i = 0; i < 15; ++i) cp_async4(dst + 85 * i, src + 320 * i);
        } else {
#pragma unroll
            for (int i = 0; i < 15; ++i) {
                const int idx = i0 + 80 * i;
                const int m = idx < HOP ? HOP - idx : idx - HOP;
                cp_async4(dst + 85 * i, clip + (long long)m * 4);
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// ROT : compile the fused rotation augmentation in (training with --augment) or out (no overhead).
// VIEW: clips are views into one resident buffer, clip b starting at sample clip_off[b] (on-the-fly
//       chunking: overlapping 20-s windows of resident 60-s files, preprocess.py:13-48) instead of
//       a dense (B, N, 4) batch.
// MIC : microphone-array format: 4 log-mel channels into a (B, 10, T, 64) tensor and the four
//       channel spectra written to `spec` (B, T, 601, 4) complex64 for the GCC-PHAT kernel; no
//       intensity vectors (ROT must be false).
template <bool ROT, bool VIEW, bool MIC>
__global__ void __launch_bounds__(NTHREADS, CTAS_PER_SM)
frontend_foa_kernel(const int16_t* __restrict__ audio, long long N, int T, int tiles_per_clip, int ntiles,
                    const FrontendTables* __restrict__ tab, const float* __restrict__ mean,
                    const float* __restrict__ istd, float dc0, float dc1, const int8_t* __restrict__ rot,
                    const long long* __restrict__ clip_off, float* __restrict__ out, float2* __restrict__ spec,
                    int* __restrict__ flags) {
    constexpr int NCH_OUT = MIC ? 10 : NCH_FOA;
    extern __shared__ __align__(16) unsigned char smem[];
    uint32_t* s_samples = reinterpret_cast<uint32_t*>(smem + SmemLayout::off_samples);
    float2* s_x1 = reinterpret_cast<float2*>(smem + SmemLayout::off_x1);
    MelEntry* s_melent = reinterpret_cast<MelEntry*>(smem + SmemLayout::off_melent);
    int* s_melhdr = reinterpret_cast<int*>(smem + SmemLayout::off_melhdr);
    float2* s_scale = reinterpret_cast<float2*>(smem + SmemLayout::off_scale);

    const int tid = threadIdx.x;
    int tile = blockIdx.x;
    if (tile < ntiles) {
        const int b = tile / tiles_per_clip, t0 = (tile % tiles_per_clip) * TF;
        issue_tile_copy(s_samples, audio, VIEW ? clip_off[b] : (long long)b * N, t0, min(TF, T - t0));
    }
    cp_async_commit();
    // constant tables -> smem (once per persistent CTA)
    for (int i = tid; i < MEL_MAXROWS * 32; i += NTHREADS) s_melent[i] = MelEntry{tab->mel_pos[i], tab->mel_w[i]};
    if (tid < 8) s_melhdr[tid] = tab->mel_hdr[tid];
    for (int i = tid; i < NCH_FOA * NMEL; i += NTHREADS) {   // standardisation as one FMA: x*is + (-mu*is)
        const float mu = mean ? mean[i] : 0.f, is = istd ? istd[i] : 1.f;
        s_scale[i] = make_float2(is, -mu * is);
    }

    // fixed roles
    const int q1 = tid / 25, n2 = tid - 25 * q1;             // stage 1 (tid < 150): lane group q1 -> fft g_of_q(q1)
    const int f1 = q1 < TF ? q1 : q1 - TF;                   // its frame
    const float cB = tab->wcs[2 * (n2 % 25)], sB = tab->wcs[2 * (n2 % 25) + 1];
    const int L = min(tid, NACT - 1);                        // stage 2: lane pairs (A,B) adjacent
    const int f2 = L / 50, t2 = (L % 50) >> 1, r2 = L & 1;
    const float c0 = r2 == 0 ? 1.0f : (1.0f / 3.0f);
    const int kt = (625 * t2) % 1200;
    float2* const vbase = s_x1 + v_base(f2) + 50 * t2 + r2;
    const int warp = tid >> 5, lane = tid & 31;

    for (; tile < ntiles; tile += gridDim.x) {
        const int b = tile / tiles_per_clip, tb = tile % tiles_per_clip;
        const int t0 = tb * TF, nf = min(TF, T - t0);
        const unsigned rb = ROT ? rot_bits_rt(rot[b]) : 0u;   // rotation augmentation of this clip (0 = none)
        cp_async_wait_all();
        __syncthreads();

        // ---- stage 1
        {
            // keep the 48 window values from being hoisted out of the tile loop (they would be
            // spilled to local memory and re-loaded, which costs more than the two FMAs each)
            float cBt = cB, sBt = sB;
            asm volatile("" : "+f"(cBt), "+f"(sBt));
            if (tid < NACT && f1 < nf) stage1_task(s_samples, s_x1, q1, n2, cBt, sBt);
        }
        __syncthreads();

        // samples are free: prefetch the next tile while stage 2 / mel run
        {
            const int nt = tile + gridDim.x;
            if (nt < ntiles) {
                const int nb = nt / tiles_per_clip, nt0 = (nt % tiles_per_clip) * TF;
                issue_tile_copy(s_samples, audio, VIEW ? clip_off[nb] : (long long)nb * N, nt0, min(TF, T - nt0));
            }
            cp_async_commit();
        }

        // ---- stage 2a
        Stage2Regs R;
        {
            // channel signs of this lane's packed FFT: role A = (W, Y), role B = (Z, X)
            const float sre = (ROT && r2 && (rb & 2u)) ? -1.f : 1.f;
            const float sim = (ROT && (((r2 ? (rb >> 2) : rb) & 1u))) ? -1.f : 1.f;
            stage2a_task(s_x1, f2, t2, r2, dc0, dc1, sre, sim, R);
        }
        __syncthreads();  // every X1 read is done -> V may overwrite the exchange buffer

        // ---- stage 2b
        {
            const bool valid = tid < NACT && f2 < nf;
            float2* const sp = MIC ? spec + ((long long)b * T + (t0 + f2)) * (NBIN * 4) + 2 * r2 : nullptr;
            int ktt = kt;
            if (MIC) asm volatile("" : "+r"(ktt));   // keep the 25 bin indices from being hoisted (and spilled)
#pragma unroll
            for (int k2 = 0; k2 < 25; ++k2) {
                SlotMine m;
                SlotOut mine, other;
                slot_split(R.P[k2], R.Q[(25 - k2) % 25], c0, m, mine);
                if (!MIC) {
                    other.s0re = __shfl_xor_sync(0xffffffffu, mine.s0re, 1);
                    other.s0im = __shfl_xor_sync(0xffffffffu, mine.s0im, 1);
                    other.e = __shfl_xor_sync(0xffffffffu, mine.e, 1);
                    float iva, ivb;
                    slot_finish(m, mine, other, r2, iva, ivb);
                    if (valid) slot_store(vbase, k2, m.P0, m.P1, iva, ivb);
                } else if (valid) {
                    vbase[2 * k2] = make_float2(m.P0, m.P1);
                    // spectra of this lane's two channels at bin k (k > 600 holds the conjugate of
                    // bin 1200 - k); the four channels of a bin fill exactly one 32-byte sector
                    int k = ktt + (576 * k2) % 1200;
                    k = k >= 1200 ? k - 1200 : k;
                    const float sg = k > 600 ? -1.f : 1.f;
                    const int kb = k > 600 ? 1200 - k : k;
                    // (lane's two channels are adjacent and 16-byte aligned: one 128-bit store)
                    *reinterpret_cast<float4*>(sp + kb * 4) = make_float4(m.S0.re, sg * m.S0.im, m.S1.re, sg * m.S1.im);
                }
            }
        }
        __syncthreads();

        // ---- mel projection + log + standardise + store (static schedule, balanced over warps)
        float* const out_tile = out + (((long long)b * NCH_OUT) * T + t0) * NMEL;
#pragma unroll 1
        for (int qq = 0; qq < MEL_SLOTS; ++qq) {
            const int code = mel_assign(warp, qq);
            if (code < 0) break;
            const int f = code >> 2, wt = code & 3;
            if (f >= nf) continue;
            float acc[8];
            mel_task<!MIC>(reinterpret_cast<const float4*>(s_x1 + v_base(f)), s_melent + s_melhdr[wt] * 32 + lane,
                           s_melhdr[4 + wt], acc);
#pragma unroll
            for (int c = 0; c < 8; ++c) acc[c] += __shfl_xor_sync(0xffffffffu, acc[c], 1);
            // uniform epilogue: lane 0 of a pair owns the 4 log-mel channels, lane 1 the 3 IV channels
            const int j = 16 * wt + (lane >> 1), part = lane & 1;
            float v[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float db = power_to_db_unclamped(acc[i]);
                v[i] = part ? acc[5 + (i < 3 ? i : 0)] : db;
            }
            if (part && !(v[0] == v[0] && v[1] == v[1] && v[2] == v[2])) atomicOr(flags, 1);  // datasets.py:277
            if (ROT && part) {                          // rotation: intensity sign follows its channel's sign
                v[0] = (rb & 1u) ? -v[0] : v[0];        // Y
                v[1] = (rb & 2u) ? -v[1] : v[1];        // Z
                v[2] = (rb & 4u) ? -v[2] : v[2];        // X
            }
            const int T64 = T * NMEL;
            if (ROT) {
                const bool swap = rb & 8u;              // X <-> Y: mel ch 1<->3, IV ch 4<->6
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    if (i == 3 && part) break;
                    int c = part * 4 + i;
                    if (swap) c = part ? (i == 0 ? 6 : (i == 2 ? 4 : c)) : (i == 1 ? 3 : (i == 3 ? 1 : c));
                    const float2 k = s_scale[c * NMEL + j];
                    out_tile[(c * T + f) * NMEL + j] = fmaf(v[i], k.x, k.y);
                }
            } else if (!MIC || part == 0) {
                const int c0i = part * 4;
                float* o = out_tile + (c0i * T + f) * NMEL + j;
                const float2* sc = s_scale + c0i * NMEL + j;
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    if (i == 3 && part) break;
                    const float2 k = sc[i * NMEL];
                    o[i * T64] = fmaf(v[i], k.x, k.y);
                }
            }
        }
        // the barrier at the top of the next iteration orders these V reads before its stage 1
    }
    cp_async_wait_all();
}

// ------------------------------------------------------------------------------------------------
// librosa.power_to_db top_db (datasets.py:265): out = max(out, max_over_clip_channel - top_db), in
// the standardised domain (x -> (x-mean)*istd is monotone, so max commutes with it exactly).
// One block per (clip, log-mel channel): pass 1 takes the max of the (T,64) plane (still L2-resident
// from the front-end kernel), pass 2 rewrites only the values below the threshold.
__global__ void __launch_bounds__(256)
clamp_topdb_kernel(float* __restrict__ out, const float* __restrict__ mean, const float* __restrict__ istd,
                   int T, float top_db, int nch) {
    __shared__ float s_max[8];
    const int b = blockIdx.x >> 2, c = blockIdx.x & 3;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float4* base = reinterpret_cast<float4*>(out + ((long long)b * nch + c) * T * NMEL);
    const int n4 = T * (NMEL / 4);
    const int j4 = threadIdx.x & 15;                      // this thread always sees mel bins 4*j4..4*j4+3
    float mu[4], is[4], sd[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        mu[q] = mean ? mean[c * NMEL + 4 * j4 + q] : 0.f;
        is[q] = istd ? istd[c * NMEL + 4 * j4 + q] : 1.f;
        sd[q] = 1.0f / is[q];                              // un-standardise with a multiply, not a divide per element
    }
    // pass 1: max of dB = v/istd + mean  (un-standardise; exact enough: the clamp value itself is
    // recomputed from the dB max below, and values are compared in the standardised domain)
    float mx = -INFINITY;
    for (int i = threadIdx.x; i < n4; i += 256) {          // 256 % 16 == 0 -> j4 is loop-invariant
        const float4 v = base[i];
        mx = fmaxf(mx, fmaxf(fmaxf(fmaf(v.x, sd[0], mu[0]), fmaf(v.y, sd[1], mu[1])),
                             fmaxf(fmaf(v.z, sd[2], mu[2]), fmaf(v.w, sd[3], mu[3]))));
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if (lane == 0) s_max[warp] = mx;
    __syncthreads();
    mx = s_max[0];
#pragma unroll
    for (int w = 1; w < 8; ++w) mx = fmaxf(mx, s_max[w]);
    const float thr = mx - top_db;
    float th[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) th[q] = fmaf(thr, is[q], -mu[q] * is[q]);   // same form as the front-end epilogue
    for (int i = threadIdx.x; i < n4; i += 256) {
        float4 v = base[i];
        if (v.x < th[0] || v.y < th[1] || v.z < th[2] || v.w < th[3]) {
            v.x = fmaxf(v.x, th[0]); v.y = fmaxf(v.y, th[1]); v.z = fmaxf(v.z, th[2]); v.w = fmaxf(v.w, th[3]);
            base[i] = v;
        }
    }
}

// ------------------------------------------------------------------------------------------------
size_t frontend_workspace_bytes(int B, long long N) {
    (void)N;
    // [flags int x4 | pad to 64] [max key uint32 (B,4)] [min key uint32 (B,4)]: per-(clip, log-mel channel) extrema of
    // the un-clamped dB values, left by the fe2 kernel for the top_db pass (order-preserving float keys, f2key)
    return 64 + (size_t)(B > 0 ? B : 0) * 32;
}

// top_db pass of the fe2 path: the fused kernel already knows every (clip, channel) maximum and minimum (its epilogue
// reduces them per warp and keeps them in the workspace), so a plane whose minimum is within top_db of its maximum --
// the usual case -- costs one block that returns at once, and the others are rewritten in one pass (no max pass).
__global__ void __launch_bounds__(256)
clamp_topdb_known_kernel(float* __restrict__ out, const float* __restrict__ mean, const float* __restrict__ istd,
                         int T, float top_db, int nch, const uint32_t* __restrict__ kmax, const uint32_t* __restrict__ kmin) {
    const int b = blockIdx.x >> 2, c = blockIdx.x & 3;
    const float mx = key2f(kmax[blockIdx.x]), mn = key2f(kmin[blockIdx.x]);
    const float thr = mx - top_db;
    if (!(mn < thr)) return;
    float4* base = reinterpret_cast<float4*>(out + ((long long)b * nch + c) * T * NMEL);
    const int n4 = T * (NMEL / 4);
    const int j4 = threadIdx.x & 15;
    float th[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const float mu = mean ? mean[c * NMEL + 4 * j4 + q] : 0.f, is = istd ? istd[c * NMEL + 4 * j4 + q] : 1.f;
        th[q] = fmaf(thr, is, -mu * is);                                   // same form as the front-end epilogue
    }
    for (int i = threadIdx.x; i < n4; i += 256) {                          // 256 % 16 == 0 -> j4 is loop-invariant
        float4 v = base[i];
        if (v.x < th[0] || v.y < th[1] || v.z < th[2] || v.w < th[3]) {
            v.x = fmaxf(v.x, th[0]); v.y = fmaxf(v.y, th[1]); v.z = fmaxf(v.z, th[2]); v.w = fmaxf(v.w, th[3]);
            base[i] = v;
        }
    }
}

static bool frontend_is_v1() {
    static const bool v1 = [] { const char* e = getenv("ADYOLO_FRONTEND"); return e && e[0] == 'v' && e[1] == '1'; }();
    return v1;
}

// ws: the workspace the preceding fe2 launch left the extrema in (NULL: recompute the maxima from `out`)
int launch_features_clamp_nch(float* out, int B, long long N, const float* mean, const float* istd, float top_db, int nch,
                               const void* ws, cudaStream_t stream) {
    const long long T = N / HOP;
    if (B <= 0 || T <= 0) return set_error(ADY_ERR_INVALID, "features_clamp: empty input");
    if (ws && !frontend_is_v1()) {
        const uint32_t* kmax = reinterpret_cast<const uint32_t*>(reinterpret_cast<const char*>(ws) + 64);
        clamp_topdb_known_kernel<<<B * 4, 256, 0, stream>>>(out, mean, istd, (int)T, top_db, nch, kmax, kmax + (size_t)B * 4);
        ADY_LAUNCH_CHECK("clamp_topdb_known_kernel");
        return ADY_OK;
    }
    clamp_topdb_kernel<<<B * 4, 256, 0, stream>>>(out, mean, istd, (int)T, top_db, nch);
    ADY_LAUNCH_CHECK("clamp_topdb_kernel");
    return ADY_OK;
}

int launch_features_foa_clamp(float* out, int B, long long N, const float* mean, const float* istd, float top_db,
                              void* ws, cudaStream_t stream) {
    return launch_features_clamp_nch(out, B, N, mean, istd, top_db, NCH_FOA, ws, stream);
}

template <bool ROT, bool VIEW, bool MIC>
static int launch_frontend_inst(int grid, cudaStream_t stream, const int16_t* audio, long long N, int T, int tpc, int ntiles,
                                const FrontendTables* tab, const float* mean, const float* istd, float dc0, float dc1,
                                const int8_t* rot, const long long* clip_off, float* out, float2* spec, int* flags) {
    // once per (instantiation, device); a racing second thread at worst repeats the idempotent call
    static std::atomic<unsigned long long> configured{0};
    int dev = 0;
    ADY_CUDA_CHECK(cudaGetDevice(&dev));
    if (!((configured.load(std::memory_order_acquire) >> (dev & 63)) & 1ull)) {
        ADY_CUDA_CHECK(cudaFuncSetAttribute(frontend_foa_kernel<ROT, VIEW, MIC>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            SmemLayout::total));
        configured.fetch_or(1ull << (dev & 63), std::memory_order_release);
    }
    frontend_foa_kernel<ROT, VIEW, MIC><<<grid, NTHREADS, SmemLayout::total, stream>>>(
        audio, N, T, tpc, ntiles, tab, mean, istd, dc0, dc1, rot, clip_off, out, spec, flags);
    ADY_LAUNCH_CHECK("frontend_foa_kernel");
    return ADY_OK;
}

int launch_features_foa(const int16_t* audio, int B, long long N, const float* mean, const float* istd,
                        float dc_offset, float top_db, int apply_topdb, const int8_t* rot, const long long* clip_off,
                        float* out, void* ws, cudaStream_t stream) {
    const long long T = N / HOP;
    if (B <= 0 || T <= 0) return set_error(ADY_ERR_INVALID, "features_foa: need B>0 and at least %d samples", HOP);
    if (N <= HOP) return set_error(ADY_ERR_INVALID, "features_foa: reflect padding needs N > %d samples", HOP);
    // The product path is the second-generation kernel (fe2.cu).  ADYOLO_FRONTEND=v1 selects the round-1 kernel of
    // this file for A/B measurements (both are hand-written sm_100a kernels; neither is a fallback).
    if (!frontend_is_v1()) {
        int rc2 = launch_features_foa_fe2(audio, B, N, mean, istd, dc_offset, rot, clip_off, out, ws, stream);
        if (rc2) return rc2;
        if (apply_topdb) return launch_features_foa_clamp(out, B, N, mean, istd, top_db, ws, stream);
        return ADY_OK;
    }
    const FrontendTables* tab = nullptr;
    int rc = get_frontend_tables(&tab);
    if (rc) return rc;
    const long long tpc = (T + TF - 1) / TF;
    const long long ntiles = (long long)B * tpc;
    if (ntiles > 0x7fffffffLL) return set_error(ADY_ERR_INVALID, "features_foa: too many tiles");
    int* flags = reinterpret_cast<int*>(ws);
    ADY_CUDA_CHECK(cudaMemsetAsync(ws, 0, 16, stream));

    int dev = 0, sms = 0;
    ADY_CUDA_CHECK(cudaGetDevice(&dev));
    ADY_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const int grid = (int)(ntiles < (long long)CTAS_PER_SM * sms ? ntiles : (long long)CTAS_PER_SM * sms);
    // window scale: 2^-15 (int16 -> [-1,1)) * 1/2 (channel split), DC terms scaled by the same 1/2
    const float dc0 = dc_offset * 300.0f, dc1 = -dc_offset * 150.0f;
    if (rot && clip_off)
        rc = launch_frontend_inst<true, true, false>(grid, stream, audio, N, (int)T, (int)tpc, (int)ntiles, tab, mean, istd, dc0, dc1, rot, clip_off, out, nullptr, flags);
    else if (rot)
        rc = launch_frontend_inst<true, false, false>(grid, stream, audio, N, (int)T, (int)tpc, (int)ntiles, tab, mean, istd, dc0, dc1, rot, clip_off, out, nullptr, flags);
    else if (clip_off)
        rc = launch_frontend_inst<false, true, false>(grid, stream, audio, N, (int)T, (int)tpc, (int)ntiles, tab, mean, istd, dc0, dc1, rot, clip_off, out, nullptr, flags);
    else
        rc = launch_frontend_inst<false, false, false>(grid, stream, audio, N, (int)T, (int)tpc, (int)ntiles, tab, mean, istd, dc0, dc1, rot, clip_off, out, nullptr, flags);
    if (rc) return rc;
    if (apply_topdb) return launch_features_foa_clamp(out, B, N, mean, istd, top_db, ws, stream);
    return ADY_OK;
}

// MIC format, first half: 4 log-mel channels of a (B, 10, T, 64) tensor + the channel spectra
// (B, T, 601, 4) complex64 for launch_gcc_from_stft.
int launch_features_mic_logmel(const int16_t* audio, int B, long long N, const float* mean, const float* istd,
                               float dc_offset, float top_db, int apply_topdb, float* out, float2* spec, void* ws,
                               cudaStream_t stream) {
    const long long T = N / HOP;
    if (B <= 0 || T <= 0 || N <= HOP) return set_error(ADY_ERR_INVALID, "features_mic: need N > %d samples", HOP);
    const FrontendTables* tab = nullptr;
    int rc = get_frontend_tables(&tab);
    if (rc) return rc;
    const long long tpc = (T + TF - 1) / TF;
    const long long ntiles = (long long)B * tpc;
    if (ntiles > 0x7fffffffLL) return set_error(ADY_ERR_INVALID, "features_mic: too many tiles");
    int* flags = reinterpret_cast<int*>(ws);
    ADY_CUDA_CHECK(cudaMemsetAsync(ws, 0, 16, stream));
    int dev = 0, sms = 0;
    ADY_CUDA_CHECK(cudaGetDevice(&dev));
    ADY_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const int grid = (int)(ntiles < (long long)CTAS_PER_SM * sms ? ntiles : (long long)CTAS_PER_SM * sms);
    rc = launch_frontend_inst<false, false, true>(grid, stream, audio, N, (int)T, (int)tpc, (int)ntiles, tab, mean, istd,
                                                  dc_offset * 300.0f, -dc_offset * 150.0f, nullptr, nullptr, out, spec, flags);
    if (rc) return rc;
    if (apply_topdb) {
        clamp_topdb_kernel<<<B * 4, 256, 0, stream>>>(out, mean, istd, (int)T, top_db, 10);
        ADY_LAUNCH_CHECK("clamp_topdb_kernel");
    }
    return ADY_OK;
}


}  // namespace ady
