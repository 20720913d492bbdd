// Microbenchmark: scalar FFMA vs packed FFMA2 (fma.rn.f32x2, sm_100a) at equal FLOPs, at high and
// low occupancy.  nvcc -gencode arch=compute_100a,code=sm_100a -O3 ffma2_bench.cu -o ffma2_bench
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>   // 0: 8 x FFMA per iteration; 1: 4 x FFMA2 per iteration; 2: 8 x FADD; 3: 4 x FADD2
__global__ void k(float* out, int iters, float a, float b) {
    float x[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = threadIdx.x * 0.001f + i;
    unsigned long long p[4], A, B;
    {
        float2 t = make_float2(a, a), u = make_float2(b, b);
        A = *reinterpret_cast<unsigned long long*>(&t);
        B = *reinterpret_cast<unsigned long long*>(&u);
#pragma unroll
        for (int i = 0; i < 4; ++i) { float2 v = make_float2(x[2 * i], x[2 * i + 1]); p[i] = *reinterpret_cast<unsigned long long*>(&v); }
    }
    for (int it = 0; it < iters; ++it) {
        if (MODE == 0) {
#pragma unroll
            for (int i = 0; i < 8; ++i) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(x[i]) : "f"(a), "f"(b));
        } else if (MODE == 1) {
#pragma unroll
            for (int i = 0; i < 4; ++i) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[i]) : "l"(A), "l"(B));
        } else if (MODE == 2) {
#pragma unroll
            for (int i = 0; i < 8; ++i) asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(x[i]) : "f"(b));
        } else {
#pragma unroll
            for (int i = 0; i < 4; ++i) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(B));
        }
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += x[i];
#pragma unroll
    for (int i = 0; i < 4; ++i) { float2 v = *reinterpret_cast<float2*>(&p[i]); s += v.x + v.y; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
float run(int blocks, int threads, int iters, float* out) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<blocks, threads>>>(out, iters, 1.0001f, 0.5f);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    k<MODE><<<blocks, threads>>>(out, iters, 1.0001f, 0.5f);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    return ms;
}

int main() {
    float* out;
    cudaMalloc(&out, 148 * 8 * 1024 * sizeof(float));
    const int iters = 1 << 15;
    const char* names[4] = {"FFMA  x8", "FFMA2 x4", "FADD  x8", "FADD2 x4"};
    for (int cfg = 0; cfg < 3; ++cfg) {
        const int threads = cfg == 0 ? 1024 : (cfg == 1 ? 320 : 128);     // warps per SM: 32 / 10 / 4
        const int blocks = 148;
        float ms[4] = {run<0>(blocks, threads, iters, out), run<1>(blocks, threads, iters, out),
                       run<2>(blocks, threads, iters, out), run<3>(blocks, threads, iters, out)};
        for (int m = 0; m < 4; ++m) {
            const double lane_ops = (double)blocks * threads * iters * 8;          // scalar-equivalent ops
            printf("%2d warps/SM  %s  %.3f ms  %.1f G lane-ops/s  (%.2f per clk per SM at 1.965 GHz)\n", threads / 32, names[m],
                   ms[m], lane_ops / ms[m] / 1e6, lane_ops / ms[m] / 1e6 / 148 / 1.965);
        }
    }
    return 0;
}
