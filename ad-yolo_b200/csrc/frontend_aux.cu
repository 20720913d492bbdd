// Un-fused front-end stages behind the reference's per-function call surface (the MIC GCC-PHAT
// lag transform lives in gcc_tc.cu).  These materialise the STFT in HBM, so they are the compatibility / offline
// route; the training hot path is the fused kernel in frontend.cu.
//
//   stft_kernel            utility.py:142-165 audio2stft  == datasets.py:252-258 get_stft_spectrogram
//   logmel_from_stft       utility.py:168-191 stft2melscale == datasets.py:260-267
//   iv_from_stft           utility.py:194-215 stft2iv       == datasets.py:269-279
//   clamp_strided          librosa.power_to_db top_db clamp with the global max per (clip, channel)
#include "common.cuh"
#include "frontend_core.cuh"
#include "frontend_host.h"

namespace ady {

struct OutStrides {
    long long sb, sc, st, sj;  // element strides of (batch, channel, frame, mel/lag)
};

// ------------------------------------------------------------------------------------------------
// one block per (clip, frame, channel pair); 32 threads, 25 active. out (B,T,601,4) complex64.
template <typename SampleT>
__global__ void __launch_bounds__(32)
stft_kernel(const SampleT* __restrict__ audio, long long N, int T, const FrontendTables* __restrict__ tab,
            float dc, float2* __restrict__ out) {
    __shared__ float2 x1[48 * 25];
    const int pair = blockIdx.x & 1;
    const long long bt = blockIdx.x >> 1;
    const long long b = bt / T;
    const int t = (int)(bt - b * T);
    const int tid = threadIdx.x;
    const SampleT* clip = audio + b * N * 4 + pair * 2;
    if (tid < 25) {
        cx<float> x[48];
#pragma unroll
        for (int n1 = 0; n1 < 48; ++n1) {
            const int n = (25 * n1 + 48 * tid) % 1200;
            long long m = (long long)t * HOP - HOP + n;   // librosa center=True, reflect pad 600
            if (m < 0) m = -m;
            if (m >= N) m = 2 * (N - 1) - m;
            float a, c;
            if constexpr (sizeof(SampleT) == 2) {
                a = (float)clip[m * 4] * (1.0f / 32768.0f) + dc;
                c = (float)clip[m * 4 + 1] * (1.0f / 32768.0f) + dc;
            } else {
                a = (float)clip[m * 4];
                c = (float)clip[m * 4 + 1];
            }
            const float w = 0.5f * tab->hann[n];
            x[n1] = {a * w, c * w};
        }
        dft48(x);
#pragma unroll
        for (int k1 = 0; k1 < 48; ++k1) x1[k1 * 25 + tid] = make_float2(x[k1].re, x[k1].im);
    }
    __syncthreads();
    if (tid < 25) {
        cx<float> P[25], Q[25];
        const int ra = tid, rb = (48 - tid) % 48;
#pragma unroll
        for (int n2 = 0; n2 < 25; ++n2) {
            const float2 a = x1[ra * 25 + n2], c = x1[rb * 25 + n2];
            P[n2] = {a.x, a.y};
            Q[n2] = {c.x, c.y};
        }
        dft25(P);
        dft25(Q);
        const int kt = (625 * tid) % 1200;
        float2* o = out + (b * T + t) * (long long)NBIN * 4 + pair * 2;
#pragma unroll
        for (int k2 = 0; k2 < 25; ++k2) {
            const cx<float> a = P[k2], q = Q[(25 - k2) % 25];
            int k = kt + (576 * k2) % 1200;
            k = k >= 1200 ? k - 1200 : k;
            const float s = k > 600 ? -1.f : 1.f;   // bin k > 600 holds conj of bin 1200-k
            const int kb = k > 600 ? 1200 - k : k;
            o[(long long)kb * 4] = make_float2(a.re + q.re, s * (a.im - q.im));
            o[(long long)kb * 4 + 1] = make_float2(a.im + q.im, s * (q.re - a.re));
        }
    }
}

// ------------------------------------------------------------------------------------------------
// block per (clip, frame); thread (j = tid & 63, c = tid >> 6), C <= 4.  spec (B,T,601,Cs) complex64.
__global__ void __launch_bounds__(256)
logmel_from_stft_kernel(const float2* __restrict__ spec, int T, int C, int Cs, const FrontendTables* __restrict__ tab,
                        const float* __restrict__ mean, const float* __restrict__ istd, float* __restrict__ out,
                        OutStrides os, uint32_t* __restrict__ gmax) {
    const long long bt = blockIdx.x;
    const long long b = bt / T;
    const int t = (int)(bt - b * T);
    const int j = threadIdx.x & 63, c = threadIdx.x >> 6;
    float db = -INFINITY;
    if (c < C) {
        const int start = tab->melidx[j], len = tab->melidx[NMEL + j], off = tab->melidx[2 * NMEL + j];
        const float2* s = spec + (bt * NBIN + start) * Cs + c;
        float acc = 0.f;
        for (int i = 0; i < len; ++i) {
            const float2 v = s[(long long)i * Cs];
            acc += tab->melw[off + i] * (v.x * v.x + v.y * v.y);
        }
        db = power_to_db_unclamped(acc);
        const float mu = mean ? mean[c * NMEL + j] : 0.f, is = istd ? istd[c * NMEL + j] : 1.f;
        out[b * os.sb + c * os.sc + t * os.st + j * os.sj] = (db - mu) * is;
    }
    // max over the 64 mel bins of this (frame, channel): two warps per channel
    float m = db;
#pragma unroll
    for (int o = 16; o; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0 && c < C) atomicMax(&gmax[b * C + c], f2key(m));
}

__global__ void __launch_bounds__(256)
clamp_strided_kernel(float* __restrict__ out, int T, int C, const float* __restrict__ mean,
                     const float* __restrict__ istd, OutStrides os, const uint32_t* __restrict__ gmax, float top_db) {
    const long long bt = blockIdx.x;
    const long long b = bt / T;
    const int t = (int)(bt - b * T);
    const int j = threadIdx.x & 63, c = threadIdx.x >> 6;
    if (c >= C) return;
    const float thr = key2f(gmax[b * C + c]) - top_db;
    const float mu = mean ? mean[c * NMEL + j] : 0.f, is = istd ? istd[c * NMEL + j] : 1.f;
    float* p = out + b * os.sb + c * os.sc + t * os.st + j * os.sj;
    *p = fmaxf(*p, (thr - mu) * is);
}

// block per (clip, frame); thread (j, c in 0..2) (+1 idle group).  spec (B,T,601,4), channel 0 = W.
__global__ void __launch_bounds__(256)
iv_from_stft_kernel(const float2* __restrict__ spec, int T, const FrontendTables* __restrict__ tab,
                    const float* __restrict__ mean, const float* __restrict__ istd, float* __restrict__ out,
                    OutStrides os, int* __restrict__ flags) {
    const long long bt = blockIdx.x;
    const long long b = bt / T;
    const int t = (int)(bt - b * T);
    const int j = threadIdx.x & 63, c = threadIdx.x >> 6;
    if (c >= 3) return;
    const int start = tab->melidx[j], len = tab->melidx[NMEL + j], off = tab->melidx[2 * NMEL + j];
    const float2* s = spec + (bt * NBIN + start) * 4;
    float acc = 0.f;
    for (int i = 0; i < len; ++i) {
        const float2 w = s[i * 4], y = s[i * 4 + 1], z = s[i * 4 + 2], x = s[i * 4 + 3];
        const float E = 1e-8f + ((w.x * w.x + w.y * w.y) +
                                 ((y.x * y.x + y.y * y.y) + (z.x * z.x + z.y * z.y) + (x.x * x.x + x.y * x.y)) * (1.0f / 3.0f));
        const float2 v = c == 0 ? y : (c == 1 ? z : x);
        acc += tab->melw[off + i] * ((w.x * v.x + w.y * v.y) / E);
    }
    if (!(acc == acc)) atomicOr(flags, 1);
    const float mu = mean ? mean[c * NMEL + j] : 0.f, is = istd ? istd[c * NMEL + j] : 1.f;
    out[b * os.sb + c * os.sc + t * os.st + j * os.sj] = (acc - mu) * is;
}

// ------------------------------------------------------------------------------------------------
int launch_stft(const void* audio, int dtype, int B, long long N, float dc, float2* out, cudaStream_t stream) {
    const long long T = N / HOP;
    if (B <= 0 || T <= 0 || N <= HOP) return set_error(ADY_ERR_INVALID, "stft: need N > %d samples", HOP);
    const FrontendTables* tab = nullptr;
    int rc = get_frontend_tables(&tab);
    if (rc) return rc;
    const long long nblk = (long long)B * T * 2;
    if (nblk > 0x7fffffffLL) return set_error(ADY_ERR_INVALID, "stft: too many frames");
    if (dtype == 0)
        stft_kernel<int16_t><<<(unsigned)nblk, 32, 0, stream>>>((const int16_t*)audio, N, (int)T, tab, dc, out);
    else if (dtype == 1)
        stft_kernel<float><<<(unsigned)nblk, 32, 0, stream>>>((const float*)audio, N, (int)T, tab, 0.f, out);
    else
        return set_error(ADY_ERR_INVALID, "stft: audio_dtype must be 0 (int16) or 1 (float32)");
    ADY_LAUNCH_CHECK("stft_kernel");
    return ADY_OK;
}

int launch_logmel_from_stft(const float2* spec, int B, long long T, int C, int Cs, const float* mean, const float* istd,
                            float* out, OutStrides os, uint32_t* gmax_ws, float top_db, int apply_topdb,
                            cudaStream_t stream) {
    if (C < 1 || C > 4 || Cs < C) return set_error(ADY_ERR_INVALID, "logmel_from_stft: 1..4 channels");
    const FrontendTables* tab = nullptr;
    int rc = get_frontend_tables(&tab);
    if (rc) return rc;
    ADY_CUDA_CHECK(cudaMemsetAsync(gmax_ws, 0, (size_t)B * C * 4, stream));
    logmel_from_stft_kernel<<<(unsigned)(B * T), 256, 0, stream>>>(spec, (int)T, C, Cs, tab, mean, istd, out, os, gmax_ws);
    ADY_LAUNCH_CHECK("logmel_from_stft_kernel");
    if (apply_topdb) {
        clamp_strided_kernel<<<(unsigned)(B * T), 256, 0, stream>>>(out, (int)T, C, mean, istd, os, gmax_ws, top_db);
        ADY_LAUNCH_CHECK("clamp_strided_kernel");
    }
    return ADY_OK;
}

int launch_iv_from_stft(const float2* spec, int B, long long T, const float* mean, const float* istd, float* out,
                        OutStrides os, int* flags, cudaStream_t stream) {
    const FrontendTables* tab = nullptr;
    int rc = get_frontend_tables(&tab);
    if (rc) return rc;
    iv_from_stft_kernel<<<(unsigned)(B * T), 256, 0, stream>>>(spec, (int)T, tab, mean, istd, out, os, flags);
    ADY_LAUNCH_CHECK("iv_from_stft_kernel");
    return ADY_OK;
}

}  // namespace ady
