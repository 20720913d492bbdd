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
// is reduced with atomicMax while the unclamped values are written; clamp_topdb_kernel then
// rewrites only the tiles whose minimum falls below max-80 dB.
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
                                                long long N, int b, int t0, int nf) {
    const int tid = threadIdx.x;
    const int pair = tid & 1, i0 = tid >> 1;  // element (idx = i0 + 80 i, pair), i = 0..14
    const int16_t* clip = audio + (long long)b * N * 4 + pair * 2;
    for (int f = 0; f < nf; ++f) {
        const int t = t0 + f;
        uint32_t* dst = samples + (2 * f + pair) * SPLANE + skew(i0);  // skew(i0 + 80 i) = skew(i0) + 85 i
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

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ float warp_min(float v) {
#pragma unroll
    for (int o = 16; o; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(NTHREADS, 2)
frontend_foa_kernel(const int16_t* __restrict__ audio, long long N, int T, int tiles_per_clip, int ntiles,
                    const FrontendTables* __restrict__ tab, const float* __restrict__ mean,
                    const float* __restrict__ istd, float dc0, float dc1, float* __restrict__ out,
                    uint32_t* __restrict__ gmax, float* __restrict__ tmin, int* __restrict__ flags) {
    extern __shared__ __align__(16) unsigned char smem[];
    uint32_t* s_samples = reinterpret_cast<uint32_t*>(smem + SmemLayout::off_samples);
    float2* s_x1 = reinterpret_cast<float2*>(smem + SmemLayout::off_x1);
    float* s_win = reinterpret_cast<float*>(smem + SmemLayout::off_win);
    float* s_melw = reinterpret_cast<float*>(smem + SmemLayout::off_melw);
    int16_t* s_melidx = reinterpret_cast<int16_t*>(smem + SmemLayout::off_melidx);
    uint32_t* s_red = reinterpret_cast<uint32_t*>(smem + SmemLayout::off_red);

    const int tid = threadIdx.x;
    int tile = blockIdx.x;
    if (tile < ntiles) {
        const int b = tile / tiles_per_clip, t0 = (tile % tiles_per_clip) * TF;
        issue_tile_copy(s_samples, audio, N, b, t0, min(TF, T - t0));
    }
    cp_async_commit();
    // constant tables -> smem (once per persistent CTA)
    for (int i = tid; i < 25 * WROW; i += NTHREADS) s_win[i] = tab->win[i];
    for (int i = tid; i < MEL_MAXNNZ; i += NTHREADS) s_melw[i] = tab->melw[i];
    for (int i = tid; i < 3 * NMEL; i += NTHREADS) s_melidx[i] = tab->melidx[i];

    // fixed roles
    const int g1 = tid / 25, n2 = tid - 25 * g1;             // stage 1 (tid < 150)
    const int L = min(tid, 6 * 25 - 1);                      // stage 2: lane pairs (A,B) adjacent
    const int f2 = L / 50, t2 = (L % 50) >> 1, r2 = L & 1;
    const int kt = (625 * t2) % 1200;
    const float c0 = r2 == 0 ? 1.0f : (1.0f / 3.0f);

    for (; tile < ntiles; tile += gridDim.x) {
        const int b = tile / tiles_per_clip, tb = tile % tiles_per_clip;
        const int t0 = tb * TF, nf = min(TF, T - t0);
        cp_async_wait_all();
        __syncthreads();
        if (tid < 8) s_red[tid] = tid < 4 ? 0u : 0xffffffffu;  // max keys, min keys

        // ---- stage 1
        if (tid < 150 && (g1 >> 1) < nf) stage1_task(s_samples, s_win, s_x1, g1, n2);
        __syncthreads();

        // samples are free: prefetch the next tile while stage 2 / mel run
        {
            const int nt = tile + gridDim.x;
            if (nt < ntiles) {
                const int nb = nt / tiles_per_clip, nt0 = (nt % tiles_per_clip) * TF;
                issue_tile_copy(s_samples, audio, N, nb, nt0, min(TF, T - nt0));
            }
            cp_async_commit();
        }

        // ---- stage 2a
        Stage2Regs R;
        stage2a_task(s_x1, f2, t2, r2, dc0, dc1, R);
        __syncthreads();  // every X1 read is done -> V may overwrite the exchange buffer

        // ---- stage 2b
        {
            const bool valid = tid < 150 && f2 < nf;
            float2* vframe = s_x1 + f2 * VFRAME;
#pragma unroll
            for (int k2 = 0; k2 < 25; ++k2) {
                SlotMine m;
                SlotOut mine, other;
                slot_split(R.P[k2], R.Q[(25 - k2) % 25], c0, m, mine);
                other.s0re = __shfl_xor_sync(0xffffffffu, mine.s0re, 1);
                other.s0im = __shfl_xor_sync(0xffffffffu, mine.s0im, 1);
                other.e = __shfl_xor_sync(0xffffffffu, mine.e, 1);
                float iva, ivb;
                slot_finish(m, mine, other, r2, iva, ivb);
                if (valid) slot_store(vframe, slot_bin(kt, k2), r2, m.P0, m.P1, iva, ivb);
            }
        }
        __syncthreads();

        // ---- mel projection + log + standardise + store
#pragma unroll 1
        for (int rr = 0; rr < 3; ++rr) {
            const int task = rr * NTHREADS + tid;     // (f, half, j), j fastest; warp-uniform (f, half)
            if (task >= TF * 2 * NMEL) break;
            const int f = task >> 7, half = (task >> 6) & 1, j = task & 63;
            if (f >= nf) continue;
            float acc[4];
            mel_task(reinterpret_cast<const float4*>(s_x1 + f * VFRAME), s_melw, s_melidx, half, j, acc);
            const long long tt = t0 + f;
            if (half == 0) {
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const float db = power_to_db_unclamped(acc[c]);
                    const float mx = warp_max(db), mn = warp_min(db);
                    if ((tid & 31) == 0) {
                        atomicMax(&s_red[c], f2key(mx));
                        atomicMin(&s_red[4 + c], f2key(mn));
                    }
                    const float mu = mean ? __ldg(mean + c * NMEL + j) : 0.f;
                    const float is = istd ? __ldg(istd + c * NMEL + j) : 1.f;
                    out[(((long long)b * NCH_FOA + c) * T + tt) * NMEL + j] = (db - mu) * is;
                }
            } else {
                bool bad = false;
#pragma unroll
                for (int c = 4; c < 7; ++c) {
                    const float v = acc[c - 3];
                    bad |= !(v == v);
                    const float mu = mean ? __ldg(mean + c * NMEL + j) : 0.f;
                    const float is = istd ? __ldg(istd + c * NMEL + j) : 1.f;
                    out[(((long long)b * NCH_FOA + c) * T + tt) * NMEL + j] = (v - mu) * is;
                }
                if (bad) atomicOr(flags, 1);  // reference prints + exit() on NaN (datasets.py:277)
            }
        }
        __syncthreads();
        if (tid < 4) {
            atomicMax(&gmax[b * 4 + tid], s_red[tid]);
            tmin[((long long)b * 4 + tid) * tiles_per_clip + tb] = key2f(s_red[4 + tid]);
        }
    }
    cp_async_wait_all();
}

// ------------------------------------------------------------------------------------------------
// top_db: out = max(out, standardise(max_db - top_db)) for the tiles that need it.
// grid = B*4 blocks (one per clip-channel), 8 warps; warp w walks tiles w, w+8, ...
__global__ void __launch_bounds__(256)
clamp_topdb_kernel(float* __restrict__ out, const uint32_t* __restrict__ gmax, const float* __restrict__ tmin,
                   const float* __restrict__ mean, const float* __restrict__ istd, int T, int tiles_per_clip,
                   float top_db) {
    const int b = blockIdx.x >> 2, c = blockIdx.x & 3;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const float thr = key2f(gmax[blockIdx.x]) - top_db;
    const float th0 = (thr - (mean ? mean[c * NMEL + lane] : 0.f)) * (istd ? istd[c * NMEL + lane] : 1.f);
    const float th1 = (thr - (mean ? mean[c * NMEL + lane + 32] : 0.f)) * (istd ? istd[c * NMEL + lane + 32] : 1.f);
    float* base = out + ((long long)b * NCH_FOA + c) * T * NMEL;
    for (int tb = warp; tb < tiles_per_clip; tb += 8) {
        if (tmin[(long long)blockIdx.x * tiles_per_clip + tb] >= thr) continue;
        const int t0 = tb * TF, nf = min(TF, T - t0);
        for (int f = 0; f < nf; ++f) {
            float* row = base + (long long)(t0 + f) * NMEL;
            row[lane] = fmaxf(row[lane], th0);
            row[lane + 32] = fmaxf(row[lane + 32], th1);
        }
    }
}

// ------------------------------------------------------------------------------------------------
size_t frontend_workspace_bytes(int B, long long N) {
    const long long T = N / HOP;
    const long long tpc = (T + TF - 1) / TF;
    return (size_t)(16 + (long long)B * 4 * 4 + (long long)B * 4 * tpc * 4 + 64);
}

int launch_features_foa(const int16_t* audio, int B, long long N, const float* mean, const float* istd,
                        float dc_offset, float top_db, int apply_topdb, float* out, void* ws, cudaStream_t stream) {
    const long long T = N / HOP;
    if (B <= 0 || T <= 0) return set_error(ADY_ERR_INVALID, "features_foa: need B>0 and at least %d samples", HOP);
    if (N <= HOP) return set_error(ADY_ERR_INVALID, "features_foa: reflect padding needs N > %d samples", HOP);
    const FrontendTables* tab = nullptr;
    int rc = get_frontend_tables(&tab);
    if (rc) return rc;
    const long long tpc = (T + TF - 1) / TF;
    const long long ntiles = (long long)B * tpc;
    if (ntiles > 0x7fffffffLL) return set_error(ADY_ERR_INVALID, "features_foa: too many tiles");
    // workspace: [flags int x4][gmax u32 B*4][tmin f32 B*4*tpc]
    int* flags = reinterpret_cast<int*>(ws);
    uint32_t* gmax = reinterpret_cast<uint32_t*>(flags + 4);
    float* tmin = reinterpret_cast<float*>(gmax + (size_t)B * 4);
    ADY_CUDA_CHECK(cudaMemsetAsync(ws, 0, 16 + (size_t)B * 16, stream));

    static int configured_dev = -1;
    int dev = 0, sms = 0;
    ADY_CUDA_CHECK(cudaGetDevice(&dev));
    ADY_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    if (configured_dev != dev) {
        ADY_CUDA_CHECK(cudaFuncSetAttribute(frontend_foa_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            SmemLayout::total));
        configured_dev = dev;
    }
    const int grid = (int)(ntiles < 2LL * sms ? ntiles : 2LL * sms);
    // window scale: 2^-15 (int16 -> [-1,1)) * 1/2 (channel split), DC terms scaled by the same 1/2
    frontend_foa_kernel<<<grid, NTHREADS, SmemLayout::total, stream>>>(
        audio, N, (int)T, (int)tpc, (int)ntiles, tab, mean, istd, dc_offset * 300.0f, -dc_offset * 150.0f, out,
        gmax, tmin, flags);
    ADY_LAUNCH_CHECK("frontend_foa_kernel");
    if (apply_topdb) {
        clamp_topdb_kernel<<<B * 4, 256, 0, stream>>>(out, gmax, tmin, mean, istd, (int)T, (int)tpc, top_db);
        ADY_LAUNCH_CHECK("clamp_topdb_kernel");
    }
    return ADY_OK;
}

int launch_features_foa_clamp(float* out, int B, long long N, const float* mean, const float* istd, float top_db,
                              void* ws, cudaStream_t stream) {
    const long long T = N / HOP;
    if (B <= 0 || T <= 0) return set_error(ADY_ERR_INVALID, "features_foa_clamp: empty input");
    const long long tpc = (T + TF - 1) / TF;
    int* flags = reinterpret_cast<int*>(ws);
    uint32_t* gmax = reinterpret_cast<uint32_t*>(flags + 4);
    float* tmin = reinterpret_cast<float*>(gmax + (size_t)B * 4);
    clamp_topdb_kernel<<<B * 4, 256, 0, stream>>>(out, gmax, tmin, mean, istd, (int)T, (int)tpc, top_db);
    ADY_LAUNCH_CHECK("clamp_topdb_kernel");
    return ADY_OK;
}

}  // namespace ady
