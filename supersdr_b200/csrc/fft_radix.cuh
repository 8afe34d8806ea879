// Register-resident radix-2/4/8/16 DIF butterflies, DESIGN.md section 4.2.
//
// Every line is one IEEE float32 operation in a fixed order (the translation unit is compiled with
// -fmad=false; the only fused multiply-adds are the explicit __fmaf_rn below), so the results are
// bit-identical to the scalar statement in oracle/c/ssdr_oracle.c.
#pragma once
#include <cuda_runtime.h>

namespace ssdr {

#define SSDR_DEV __device__ __forceinline__

constexpr float kA = 0.92387953251128674f;  // cos(pi/8)
constexpr float kB = 0.70710678118654752f;  // sqrt(1/2)
constexpr float kC = 0.38268343236508977f;  // sin(pi/8)

SSDR_DEV float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
SSDR_DEV float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
// (u.re + i u.im)(w.re + i w.im): re = fma(u.re, w.re, -(u.im*w.im)), im = fma(u.re, w.im, u.im*w.re)
SSDR_DEV float2 cmul(float2 u, float2 w) {
    float t0 = u.y * w.y;
    float t1 = u.y * w.x;
    return make_float2(__fmaf_rn(u.x, w.x, -t0), __fmaf_rn(u.x, w.y, t1));
}
SSDR_DEV float2 mul_mi(float2 u) { return make_float2(u.y, -u.x); }                              // * (-i)
SSDR_DEV float2 mul_w8(float2 u) { return make_float2((u.x + u.y) * kB, (u.y - u.x) * kB); }      // * B(1-i)
SSDR_DEV float2 mul_w83(float2 u) { return make_float2((u.y - u.x) * kB, (u.x + u.y) * (-kB)); }  // * -B(1+i)

SSDR_DEV void dft2(float2& x0, float2& x1) {
    float2 a = x0, b = x1;
    x0 = cadd(a, b);
    x1 = csub(a, b);
}

SSDR_DEV void dft4(float2& x0, float2& x1, float2& x2, float2& x3) {
    float2 a = cadd(x0, x2), b = csub(x0, x2);
    float2 c = cadd(x1, x3), d = mul_mi(csub(x1, x3));
    x0 = cadd(a, c);
    x2 = csub(a, c);
    x1 = cadd(b, d);
    x3 = csub(b, d);
}

// Natural-order in, natural-order out: x[q] = sum_m x[m] W_R^(m q).
template <int R>
SSDR_DEV void dft(float2 (&x)[R]);

template <>
SSDR_DEV void dft<2>(float2 (&x)[2]) { dft2(x[0], x[1]); }

template <>
SSDR_DEV void dft<4>(float2 (&x)[4]) { dft4(x[0], x[1], x[2], x[3]); }

template <>
SSDR_DEV void dft<8>(float2 (&x)[8]) {
    // stage 1: pairs (m0, m0+4) -> u[p][m0]; kept in x[m0] (p=0) and x[m0+4] (p=1)
#pragma unroll
    for (int m0 = 0; m0 < 4; ++m0) dft2(x[m0], x[m0 + 4]);
    x[5] = mul_w8(x[5]);
    x[6] = mul_mi(x[6]);
    x[7] = mul_w83(x[7]);
    dft4(x[0], x[1], x[2], x[3]);  // p = 0 -> outputs q = 0,2,4,6 in x[0..3]
    dft4(x[4], x[5], x[6], x[7]);  // p = 1 -> outputs q = 1,3,5,7 in x[4..7]
    float2 y[8];
#pragma unroll
    for (int s = 0; s < 4; ++s) { y[2 * s] = x[s]; y[2 * s + 1] = x[4 + s]; }
#pragma unroll
    for (int q = 0; q < 8; ++q) x[q] = y[q];
}

template <>
SSDR_DEV void dft<16>(float2 (&x)[16]) {
    // stage 1: for each m0, radix-4 across (m0, m0+4, m0+8, m0+12); result p lives in x[m0 + 4p]
#pragma unroll
    for (int m0 = 0; m0 < 4; ++m0) dft4(x[m0], x[m0 + 4], x[m0 + 8], x[m0 + 12]);
    // internal twiddles W16^(m0*p), u[p][m0] = x[m0 + 4p]
    const float2 w1 = make_float2(kA, -kC), w3 = make_float2(kC, -kA), w9 = make_float2(-kA, kC);
    x[5] = cmul(x[5], w1);   x[6] = mul_w8(x[6]);    x[7] = cmul(x[7], w3);
    x[9] = mul_w8(x[9]);     x[10] = mul_mi(x[10]);  x[11] = mul_w83(x[11]);
    x[13] = cmul(x[13], w3); x[14] = mul_w83(x[14]); x[15] = cmul(x[15], w9);
    // stage 2: for each p, radix-4 across m0; output s lives in x[4p + s] -> q = p + 4s
#pragma unroll
    for (int p = 0; p < 4; ++p) dft4(x[4 * p], x[4 * p + 1], x[4 * p + 2], x[4 * p + 3]);
    float2 y[16];
#pragma unroll
    for (int p = 0; p < 4; ++p)
#pragma unroll
        for (int s = 0; s < 4; ++s) y[p + 4 * s] = x[4 * p + s];
#pragma unroll
    for (int q = 0; q < 16; ++q) x[q] = y[q];
}

// Twiddle powers w[2..R-1] from w[1]: fixed minimum-depth chain, DESIGN.md section 4.4.
template <int R>
SSDR_DEV void tw_chain(float2 (&w)[R]) {
    if constexpr (R > 2) {
        w[2] = cmul(w[1], w[1]);
        w[3] = cmul(w[2], w[1]);
    }
    if constexpr (R > 4) {
        w[4] = cmul(w[2], w[2]);
        w[5] = cmul(w[4], w[1]);
        w[6] = cmul(w[3], w[3]);
        w[7] = cmul(w[4], w[3]);
    }
    if constexpr (R > 8) {
        w[8] = cmul(w[4], w[4]);
        w[9] = cmul(w[8], w[1]);
        w[10] = cmul(w[5], w[5]);
        w[11] = cmul(w[8], w[3]);
        w[12] = cmul(w[6], w[6]);
        w[13] = cmul(w[8], w[5]);
        w[14] = cmul(w[7], w[7]);
        w[15] = cmul(w[8], w[7]);
    }
}

// cos/sin(2 pi m / 16) from {1, A, B, C, 0} by symmetry; m is a compile-time constant after unrolling.
SSDR_DEV void unit16(int m, float& c, float& s) {
    const float q[5] = {1.0f, kA, kB, kC, 0.0f};
    int mm = m & 15, quad = mm >> 2, r = mm & 3;
    float cr = q[r], sr = q[4 - r];
    switch (quad) {
        case 0: c = cr; s = sr; break;
        case 1: c = -sr; s = cr; break;
        case 2: c = -cr; s = -sr; break;
        default: c = sr; s = -cr; break;
    }
}

}  // namespace ssdr
