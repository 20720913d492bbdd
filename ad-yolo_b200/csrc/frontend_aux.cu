// Un-fused front-end stages behind the reference's per-function call surface, plus the MIC
// (GCC-PHAT) path.  These materialise the STFT in HBM, so they are the compatibility / offline
// route; the training hot path is the fused kernel in frontend.cu.
//
//   stft_kernel            utility.py:142-165 audio2stft  == datasets.py:252-258 get_stft_spectrogram
//   logmel_from_stft       utility.py:168-191 stft2melscale == datasets.py:260-267
//   iv_from_stft           utility.py:194-215 stft2iv       == datasets.py:269-279
//   gcc_from_stft          NOT in the reference; upstream seld-dcase2022 _get_gcc semantics:
//                          cc = irfft(exp(1j*angle(conj(X_m) X_n))), lags [-L/2, L/2)
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

// GCC-PHAT, 64 lags x 6 microphone pairs per frame.  One block (4 warps) per (clip, frame):
//   1. PHAT-normalised cross spectra of the 6 pairs -> smem, laid out [bin][pair] so that one
//      bin's six values are three 16-byte loads that every lane of a warp shares (broadcast)
//   2. lane = lag magnitude l (0..31), warp = quarter of the bins.  cc[+l] and cc[-l] share
//        C = sum_k Re(P_k) cos(2 pi k l / N),  S = sum_k Im(P_k) sin(2 pi k l / N):
//        cc[+l] = (e + 2 (C - S)) / N,  cc[-l] = (e + 2 (C + S)) / N
//      (cos, sin) advance by a per-lane rotation each bin (restarted exactly every quarter);
//      lag magnitude 32 (only cc[-32] is needed) is summed with lane = bin residue
//   3. partial sums of the 4 warps are combined through smem; standardise; store
__global__ void __launch_bounds__(128)
gcc_from_stft_kernel(const float2* __restrict__ spec, int T, const float* __restrict__ mean,
                     const float* __restrict__ istd, float* __restrict__ out, OutStrides os) {
    __shared__ __align__(16) float2 P[NBIN][6];           // 28.8 KB
    __shared__ float red[4][13][32];                       // per warp: C[6], S[6] per lane (+ row 12: lag-32 partials)
    const long long bt = blockIdx.x;
    const long long b = bt / T;
    const int t = (int)(bt - b * T);
    const float2* s = spec + bt * NBIN * 4;
    for (int k = threadIdx.x; k < NBIN; k += 128) {
        const float4 v01 = *reinterpret_cast<const float4*>(s + k * 4), v23 = *reinterpret_cast<const float4*>(s + k * 4 + 2);
        const float2 x[4] = {make_float2(v01.x, v01.y), make_float2(v01.z, v01.w), make_float2(v23.x, v23.y), make_float2(v23.z, v23.w)};
        int p = 0;
#pragma unroll
        for (int m = 0; m < 4; ++m)
#pragma unroll
            for (int n = m + 1; n < 4; ++n) {
                float re = x[m].x * x[n].x + x[m].y * x[n].y, im = x[m].x * x[n].y - x[m].y * x[n].x;   // conj(x_m) x_n
                const float mx = fmaxf(fabsf(re), fabsf(im));
                if (mx == 0.f) { re = 1.f; im = 0.f; }                                                // np.angle(0) = 0
                else {
                    re /= mx; im /= mx;
                    const float r = rsqrtf(re * re + im * im);
                    re *= r; im *= r;
                }
                P[k][p++] = make_float2(re, im);
            }
    }
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int k0 = 1 + warp * 150, k1 = min(600, k0 + 150);       // bins 1..599 in four quarters
    float C[6] = {0, 0, 0, 0, 0, 0}, S[6] = {0, 0, 0, 0, 0, 0};
    float cs, sn, rc, rs;
    sincospif((float)((k0 * lane) % NFFT) * (1.0f / 600.0f), &sn, &cs);   // angle of bin k0 for lag `lane`
    sincospif((float)lane * (1.0f / 600.0f), &rs, &rc);                   // per-bin rotation
#pragma unroll 2
    for (int k = k0; k < k1; ++k) {
        const float4 a = *reinterpret_cast<const float4*>(&P[k][0]), c = *reinterpret_cast<const float4*>(&P[k][2]),
                     e = *reinterpret_cast<const float4*>(&P[k][4]);
        C[0] += a.x * cs; S[0] += a.y * sn; C[1] += a.z * cs; S[1] += a.w * sn;
        C[2] += c.x * cs; S[2] += c.y * sn; C[3] += c.z * cs; S[3] += c.w * sn;
        C[4] += e.x * cs; S[4] += e.y * sn; C[5] += e.z * cs; S[5] += e.w * sn;
        const float nc = cs * rc - sn * rs;
        sn = sn * rc + cs * rs;
        cs = nc;
    }
    // lag magnitude 32: lane takes the bins k0 + lane, k0 + lane + 32, ... of this quarter
    float L32[6] = {0, 0, 0, 0, 0, 0};
    for (int k = k0 + lane; k < k1; k += 32) {
        float s32, c32;
        sincospif((float)((k * 32) % NFFT) * (1.0f / 600.0f), &s32, &c32);
#pragma unroll
        for (int p = 0; p < 6; ++p) L32[p] += P[k][p].x * c32 + P[k][p].y * s32;     // cc[-32]: cos + sin
    }
#pragma unroll
    for (int p = 0; p < 6; ++p) {
#pragma unroll
        for (int o = 16; o; o >>= 1) L32[p] += __shfl_xor_sync(0xffffffffu, L32[p], o);
        red[warp][p][lane] = C[p];
        red[warp][6 + p][lane] = S[p];
    }
    if (lane < 6) red[warp][12][lane] = L32[lane];
    __syncthreads();
    // 6 pairs x 64 outputs = 384 values over 128 threads
    for (int o = threadIdx.x; o < 6 * NMEL; o += 128) {
        const int p = o >> 6, j = o & 63;
        const int lag = j - 32;                                   // output order: cc[-32:], cc[:32]
        const int l = lag < 0 ? -lag : lag;
        const float edge = P[0][p].x + ((l & 1) ? -P[600][p].x : P[600][p].x);
        float v;
        if (l == 32) v = red[0][12][p] + red[1][12][p] + red[2][12][p] + red[3][12][p];
        else {
            const float c = red[0][p][l] + red[1][p][l] + red[2][p][l] + red[3][p][l];
            const float sgm = red[0][6 + p][l] + red[1][6 + p][l] + red[2][6 + p][l] + red[3][6 + p][l];
            v = lag < 0 ? c + sgm : c - sgm;
        }
        const float cc = (2.f * v + edge) * (1.0f / NFFT);
        const float mu = mean ? mean[p * NMEL + j] : 0.f, is = istd ? istd[p * NMEL + j] : 1.f;
        out[b * os.sb + p * os.sc + t * os.st + j * os.sj] = (cc - mu) * is;
    }
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

int launch_gcc_from_stft(const float2* spec, int B, long long T, const float* mean, const float* istd, float* out,
                         OutStrides os, cudaStream_t stream) {
    gcc_from_stft_kernel<<<(unsigned)(B * T), 128, 0, stream>>>(spec, (int)T, mean, istd, out, os);
    ADY_LAUNCH_CHECK("gcc_from_stft_kernel");
    return ADY_OK;
}

}  // namespace ady
