// Streaming statistics for the `scaler` action (reference: /root/reference/src/preprocess.py:116-127,
// which concatenates every training frame in host RAM and calls numpy mean/std/max/min).
// Here: FP64 {sum, sum of squares} and {max, min} partials per (channel, mel bin), accumulated over
// any number of feature batches; ranks combine them with one NCCL all-reduce (host side).
#include "common.cuh"
#include "frontend_core.cuh"
#include "frontend_host.h"

namespace ady {

__device__ __forceinline__ void atomic_max_double(double* addr, double v) {
    unsigned long long* a = reinterpret_cast<unsigned long long*>(addr);
    unsigned long long old = *a;
    while (__longlong_as_double((long long)old) < v) {
        const unsigned long long assumed = old;
        old = atomicCAS(a, assumed, (unsigned long long)__double_as_longlong(v));
        if (old == assumed) break;
    }
}
__device__ __forceinline__ void atomic_min_double(double* addr, double v) {
    unsigned long long* a = reinterpret_cast<unsigned long long*>(addr);
    unsigned long long old = *a;
    while (__longlong_as_double((long long)old) > v) {
        const unsigned long long assumed = old;
        old = atomicCAS(a, assumed, (unsigned long long)__double_as_longlong(v));
        if (old == assumed) break;
    }
}

// grid (nchunk, C); block 256 = 4 row lanes x 64 mel bins. feats (B, C, T, 64).
__global__ void __launch_bounds__(256)
scaler_partials_kernel(const float* __restrict__ feats, int B, int C, long long T, long long rows_per_chunk,
                       double* __restrict__ sum, double* __restrict__ sumsq, double* __restrict__ maxv,
                       double* __restrict__ minv) {
    __shared__ double sh_s[4][NMEL], sh_q[4][NMEL];
    __shared__ float sh_mx[4][NMEL], sh_mn[4][NMEL];
    const int c = blockIdx.y, j = threadIdx.x & 63, rl = threadIdx.x >> 6;
    const long long nrows = (long long)B * T;
    const long long r0 = (long long)blockIdx.x * rows_per_chunk;
    const long long r1 = min(nrows, r0 + rows_per_chunk);
    double s = 0.0, q = 0.0;
    float mx = -INFINITY, mn = INFINITY;
    // (clip, frame) of the first row by one division, then a carry per step (a 64-bit division per element was most of
    // this kernel's instructions); four independent loads per iteration: the loop is otherwise a chain of DRAM round trips
    long long b = (r0 + rl) / T, t = (r0 + rl) - b * T;
    for (long long r = r0 + rl; r < r1; r += 16) {
        float v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const bool live = r + 4 * u < r1;
            v[u] = live ? feats[((b * C + c) * T + t) * NMEL + j] : 0.f;
            t += 4;
            while (t >= T) { t -= T; ++b; }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            if (r + 4 * u < r1) {
                s += (double)v[u];
                q += (double)v[u] * (double)v[u];
                mx = fmaxf(mx, v[u]);
                mn = fminf(mn, v[u]);
            }
        }
    }
    sh_s[rl][j] = s; sh_q[rl][j] = q; sh_mx[rl][j] = mx; sh_mn[rl][j] = mn;
    __syncthreads();
    if (rl == 0 && r0 < r1) {
        for (int k = 1; k < 4; ++k) {
            s += sh_s[k][j]; q += sh_q[k][j];
            mx = fmaxf(mx, sh_mx[k][j]); mn = fminf(mn, sh_mn[k][j]);
        }
        atomicAdd(&sum[c * NMEL + j], s);
        atomicAdd(&sumsq[c * NMEL + j], q);
        atomic_max_double(&maxv[c * NMEL + j], (double)mx);
        atomic_min_double(&minv[c * NMEL + j], (double)mn);
    }
}

int launch_scaler_partials(const float* feats, int B, int C, long long T, double* sum, double* sumsq,
                           double* maxv, double* minv, cudaStream_t stream) {
    if (B <= 0 || C <= 0 || T <= 0) return set_error(ADY_ERR_INVALID, "scaler_partials: empty input");
    const long long nrows = (long long)B * T;
    // 128 rows per chunk = 8 iterations of 4 independent loads per thread; few chunks per channel keep the same-address
    // FP64 atomics short (600 chunks per address serialised into ~25 us, 64 dependent loads per thread into ~35 us)
    long long nchunk = (nrows + 127) / 128;
    if (nchunk > 1184) nchunk = 1184;  // 8 x 148
    const long long rpc = (nrows + nchunk - 1) / nchunk;
    dim3 grid((unsigned)nchunk, (unsigned)C);
    scaler_partials_kernel<<<grid, 256, 0, stream>>>(feats, B, C, T, rpc, sum, sumsq, maxv, minv);
    ADY_LAUNCH_CHECK("scaler_partials_kernel");
    return ADY_OK;
}

}  // namespace ady
