// Register-resident DFT codelets for the 1200-point prime-factor FFT (48 x 25).
//
// Everything here is straight-line, fully unrolled code on small arrays that live in
// registers (all indexing is compile-time after unrolling).  The same header compiles as
// plain C++ (tests/emu builds it with g++ to check the codelets and the kernel's index
// logic on the CPU) and as CUDA device code.
//
//   dft3 / dft4 / dft5          : small Winograd-style kernels
//   dft16 = Cooley-Tukey 4 x 4   (constant twiddles W16^{jk})
//   dft48 = Good-Thomas 3 x 16   (coprime -> no twiddles, compile-time index maps)
//   dft25 = Cooley-Tukey 5 x 5   (16 constant twiddles W25^{bc})
//
// Forward transforms use exp(-2 pi i n k / N), natural order in, natural order out.
#pragma once

#if defined(__CUDACC__)
#define ADY_HD __host__ __device__ __forceinline__
#else
#define ADY_HD inline __attribute__((always_inline))
#endif

#if !defined(__CUDACC__)
// minimal vector types so the same headers build as plain C++ (tests/emu)
struct float2 { float x, y; };
struct float4 { float x, y, z, w; };
static inline float2 make_float2(float x, float y) { return {x, y}; }
static inline float4 make_float4(float x, float y, float z, float w) { return {x, y, z, w}; }
#endif

namespace ady {

template <typename T>
struct cx {
    T re, im;
};

template <typename T> ADY_HD cx<T> cadd(cx<T> a, cx<T> b) { return {a.re + b.re, a.im + b.im}; }
template <typename T> ADY_HD cx<T> csub(cx<T> a, cx<T> b) { return {a.re - b.re, a.im - b.im}; }
// a * (-i)  and  a * (+i)
template <typename T> ADY_HD cx<T> cmul_mi(cx<T> a) { return {a.im, -a.re}; }
template <typename T> ADY_HD cx<T> cmul_pi(cx<T> a) { return {-a.im, a.re}; }
template <typename T> ADY_HD cx<T> cmul(cx<T> a, T wr, T wi) {
    return {a.re * wr - a.im * wi, a.re * wi + a.im * wr};
}

// ---------------------------------------------------------------------------------- DFT-2/3/4/5
template <typename T> ADY_HD void dft2(cx<T>& a, cx<T>& b) {
    cx<T> t = a;
    a = cadd(t, b);
    b = csub(t, b);
}

template <typename T> ADY_HD void dft3(cx<T>& x0, cx<T>& x1, cx<T>& x2) {
    const T s = T(0.86602540378443864676372317075294);  // sin(2pi/3)
    cx<T> t1 = cadd(x1, x2);
    cx<T> d = csub(x1, x2);
    cx<T> m1 = {x0.re - T(0.5) * t1.re, x0.im - T(0.5) * t1.im};
    x0 = cadd(x0, t1);
    x1 = {m1.re + s * d.im, m1.im - s * d.re};  // m1 - i s d
    x2 = {m1.re - s * d.im, m1.im + s * d.re};  // m1 + i s d
}

template <typename T> ADY_HD void dft4(cx<T>& x0, cx<T>& x1, cx<T>& x2, cx<T>& x3) {
    cx<T> a = cadd(x0, x2), b = csub(x0, x2);
    cx<T> c = cadd(x1, x3), d = cmul_mi(csub(x1, x3));
    x0 = cadd(a, c);
    x1 = cadd(b, d);
    x2 = csub(a, c);
    x3 = csub(b, d);
}

template <typename T> ADY_HD void dft5(cx<T>& x0, cx<T>& x1, cx<T>& x2, cx<T>& x3, cx<T>& x4) {
    const T c1 = T(0.30901699437494742410229341718282);   // cos(2pi/5)
    const T c2 = T(-0.80901699437494742410229341718282);  // cos(4pi/5)
    const T s1 = T(0.95105651629515357211643933337938);   // sin(2pi/5)
    const T s2 = T(0.58778525229247312916870595463907);   // sin(4pi/5)
    cx<T> t1 = cadd(x1, x4), t2 = cadd(x2, x3), t3 = csub(x1, x4), t4 = csub(x2, x3);
    cx<T> a1 = {x0.re + c1 * t1.re + c2 * t2.re, x0.im + c1 * t1.im + c2 * t2.im};
    cx<T> a2 = {x0.re + c2 * t1.re + c1 * t2.re, x0.im + c2 * t1.im + c1 * t2.im};
    cx<T> b1 = {s1 * t3.re + s2 * t4.re, s1 * t3.im + s2 * t4.im};
    cx<T> b2 = {s2 * t3.re - s1 * t4.re, s2 * t3.im - s1 * t4.im};
    x0 = {x0.re + t1.re + t2.re, x0.im + t1.im + t2.im};
    x1 = {a1.re + b1.im, a1.im - b1.re};  // a1 - i b1
    x4 = {a1.re - b1.im, a1.im + b1.re};  // a1 + i b1
    x2 = {a2.re + b2.im, a2.im - b2.re};
    x3 = {a2.re - b2.im, a2.im + b2.re};
}

// ---------------------------------------------------------------------------------- DFT-16 (4 x 4)
// n = 4a + b, k = c + 4d:  X[c+4d] = sum_b W4^{bd} ( W16^{bc} sum_a x[4a+b] W4^{ac} )
template <typename T, int E> ADY_HD cx<T> tw16(cx<T> v) {
    // multiply by W16^E = exp(-2 pi i E / 16), E in {0,1,2,3,4,6,9}
    const T c8 = T(0.92387953251128675612818318939679);  // cos(pi/8)
    const T s8 = T(0.38268343236508977172845998403040);  // sin(pi/8)
    const T r2 = T(0.70710678118654752440084436210485);
    if constexpr (E == 0) return v;
    else if constexpr (E == 1) return cmul(v, c8, -s8);
    else if constexpr (E == 2) return {r2 * (v.re + v.im), r2 * (v.im - v.re)};
    else if constexpr (E == 3) return cmul(v, s8, -c8);
    else if constexpr (E == 4) return cmul_mi(v);
    else if constexpr (E == 6) return {r2 * (v.im - v.re), -r2 * (v.re + v.im)};
    else { static_assert(E == 9, "unsupported W16 power"); return cmul(v, -c8, s8); }
}

template <typename T> ADY_HD void dft16(cx<T> (&x)[16]) {
    // stage 1: for each b, DFT-4 over a (elements b, 4+b, 8+b, 12+b) -> y[b][c] stored in place
#pragma unroll
    for (int b = 0; b < 4; ++b) dft4(x[b], x[4 + b], x[8 + b], x[12 + b]);
    // now x[4c + b] = Y[b][c]; twiddle by W16^{bc}
    x[4 * 1 + 1] = tw16<T, 1>(x[4 * 1 + 1]);
    x[4 * 1 + 2] = tw16<T, 2>(x[4 * 1 + 2]);
    x[4 * 1 + 3] = tw16<T, 3>(x[4 * 1 + 3]);
    x[4 * 2 + 1] = tw16<T, 2>(x[4 * 2 + 1]);
    x[4 * 2 + 2] = tw16<T, 4>(x[4 * 2 + 2]);
    x[4 * 2 + 3] = tw16<T, 6>(x[4 * 2 + 3]);
    x[4 * 3 + 1] = tw16<T, 3>(x[4 * 3 + 1]);
    x[4 * 3 + 2] = tw16<T, 6>(x[4 * 3 + 2]);
    x[4 * 3 + 3] = tw16<T, 9>(x[4 * 3 + 3]);
    // stage 2: for each c, DFT-4 over b (elements 4c+0..3) -> X[c + 4d] at slot 4c + d
#pragma unroll
    for (int c = 0; c < 4; ++c) dft4(x[4 * c + 0], x[4 * c + 1], x[4 * c + 2], x[4 * c + 3]);
    // un-permute: slot 4c+d holds X[c+4d]  (a 4x4 transpose)
#pragma unroll
    for (int c = 0; c < 4; ++c)
#pragma unroll
        for (int d = c + 1; d < 4; ++d) {
            cx<T> t = x[4 * c + d];
            x[4 * c + d] = x[4 * d + c];
            x[4 * d + c] = t;
        }
}

// ---------------------------------------------------------------------------------- DFT-48 (PFA 3 x 16)
// input map n = (16a + 3b) mod 48, output map k = (16c + 33d) mod 48   (a,c<3; b,d<16)
template <typename T> ADY_HD void dft48(cx<T> (&x)[48]) {
    cx<T> y[3][16];
#pragma unroll
    for (int b = 0; b < 16; ++b) {
        cx<T> p0 = x[(16 * 0 + 3 * b) % 48], p1 = x[(16 * 1 + 3 * b) % 48], p2 = x[(16 * 2 + 3 * b) % 48];
        dft3(p0, p1, p2);
        y[0][b] = p0;
        y[1][b] = p1;
        y[2][b] = p2;
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        dft16(y[c]);
#pragma unroll
        for (int d = 0; d < 16; ++d) x[(16 * c + 33 * d) % 48] = y[c][d];
    }
}

// ---------------------------------------------------------------------------------- DFT-25 (CT 5 x 5)
// n = 5a + b, k = c + 5d.  TW25[j] = exp(-2 pi i j / 25), j = b*c in {1,2,3,4,6,8,9,12,16}
template <typename T, int J> ADY_HD cx<T> tw25(cx<T> v) {
    if constexpr (J == 0) return v;
    else if constexpr (J == 1) return cmul(v, T(0.96858316112863108), T(-0.24868988716485479));
    else if constexpr (J == 2) return cmul(v, T(0.87630668004386358), T(-0.48175367410171527));
    else if constexpr (J == 3) return cmul(v, T(0.72896862742141155), T(-0.68454710592868862));
    else if constexpr (J == 4) return cmul(v, T(0.53582679497899666), T(-0.84432792550201508));
    else if constexpr (J == 6) return cmul(v, T(0.06279051952931337), T(-0.99802672842827156));
    else if constexpr (J == 8) return cmul(v, T(-0.42577929156507272), T(-0.90482705246601958));
    else if constexpr (J == 9) return cmul(v, T(-0.63742398974868975), T(-0.77051324277578925));
    else if constexpr (J == 12) return cmul(v, T(-0.99211470131447788), T(-0.12533323356430426));
    else { static_assert(J == 16, "unsupported W25 power"); return cmul(v, T(-0.63742398974868975), T(0.77051324277578925)); }
}

template <typename T, int B> ADY_HD void tw25_row(cx<T>& y1, cx<T>& y2, cx<T>& y3, cx<T>& y4) {
    y1 = tw25<T, B * 1>(y1);
    y2 = tw25<T, B * 2>(y2);
    y3 = tw25<T, B * 3>(y3);
    y4 = tw25<T, B * 4>(y4);
}

template <typename T> ADY_HD void dft25(cx<T> (&x)[25]) {
    // stage 1: for each b, DFT-5 over a on x[5a+b] -> Y[b][c] left in slot 5c + b
#pragma unroll
    for (int b = 0; b < 5; ++b) dft5(x[b], x[5 + b], x[10 + b], x[15 + b], x[20 + b]);
    tw25_row<T, 1>(x[5 * 1 + 1], x[5 * 2 + 1], x[5 * 3 + 1], x[5 * 4 + 1]);
    tw25_row<T, 2>(x[5 * 1 + 2], x[5 * 2 + 2], x[5 * 3 + 2], x[5 * 4 + 2]);
    tw25_row<T, 3>(x[5 * 1 + 3], x[5 * 2 + 3], x[5 * 3 + 3], x[5 * 4 + 3]);
    tw25_row<T, 4>(x[5 * 1 + 4], x[5 * 2 + 4], x[5 * 3 + 4], x[5 * 4 + 4]);
    // stage 2: for each c, DFT-5 over b on slots 5c + b -> X[c + 5d] at slot 5c + d
#pragma unroll
    for (int c = 0; c < 5; ++c) dft5(x[5 * c], x[5 * c + 1], x[5 * c + 2], x[5 * c + 3], x[5 * c + 4]);
#pragma unroll
    for (int c = 0; c < 5; ++c)
#pragma unroll
        for (int d = c + 1; d < 5; ++d) {
            cx<T> t = x[5 * c + d];
            x[5 * c + d] = x[5 * d + c];
            x[5 * d + c] = t;
        }
}

// ---------------------------------------------------------------------------------- 1200 = 48 x 25 PFA maps
// input  n = (25 n1 + 48 n2) mod 1200      (n1 < 48, n2 < 25)
// output k = (625 k1 + 576 k2) mod 1200    (k == k1 mod 48, k == k2 mod 25)
constexpr int PFA_N = 1200, PFA_N1 = 48, PFA_N2 = 25;
ADY_HD constexpr int pfa_in(int n1, int n2) { return (25 * n1 + 48 * n2) % 1200; }
ADY_HD constexpr int pfa_out(int k1, int k2) { return (625 * k1 + 576 * k2) % 1200; }

}  // namespace ady
