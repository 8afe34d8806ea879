// Register-resident radix-2/4/8/16/32 DIF butterflies on packed float2 arithmetic, DESIGN.md 4.2.
//
// A complex value is one 64-bit register pair and every complex add / subtract / multiply is
// issued as sm_100 packed-fp32 instructions (FADD2 / FMUL2 / FFMA2: two IEEE float32 operations per
// lane per issue slot, with the swap / negate / broadcast operand modifiers doing the "multiply by
// -i", the conjugations and the real-by-complex products for free).  Each packed instruction is
// exactly the two scalar float32 operations the spec states, in the stated order, so the results are
// bit-identical to the scalar statement in oracle/c/ssdr_oracle.c.  The translation unit is compiled
// with -fmad=false: the only fused multiply-adds are the explicit ones written here.
#pragma once
#include <cuda_runtime.h>

namespace ssdr {

#define SSDR_DEV __device__ __forceinline__

constexpr float kA = 0.92387953251128674f;  // cos(pi/8)
constexpr float kB = 0.70710678118654752f;  // sqrt(1/2)
constexpr float kC = 0.38268343236508977f;  // sin(pi/8)

// cos(2 pi r / 32), r = 0..8 (first octant + the two axis values); the rest follows by symmetry.
__device__ constexpr float kQ32[9] = {1.0f, 0.98078528040323043f, 0.92387953251128674f, 0.83146961230254524f,
                                       0.70710678118654752f, 0.55557023301960218f, 0.38268343236508977f,
                                       0.19509032201612825f, 0.0f};

// cos / sin (2 pi m / 32); m is a compile-time constant after unrolling.
SSDR_DEV constexpr float unit32_cos(int m) {
    const int mm = m & 31, quad = mm >> 3, r = mm & 7;
    return quad == 0 ? kQ32[r] : quad == 1 ? -kQ32[8 - r] : quad == 2 ? -kQ32[r] : kQ32[8 - r];
}
SSDR_DEV constexpr float unit32_sin(int m) {
    const int mm = m & 31, quad = mm >> 3, r = mm & 7;
    return quad == 0 ? kQ32[8 - r] : quad == 1 ? kQ32[r] : quad == 2 ? -kQ32[8 - r] : -kQ32[r];
}

SSDR_DEV float2 cadd(float2 a, float2 b) { return __fadd2_rn(a, b); }
SSDR_DEV float2 csub(float2 a, float2 b) { return __fadd2_rn(a, make_float2(-b.x, -b.y)); }
// (u.re + i u.im)(w.re + i w.im):  re = fma(u.re, w.re, -(u.im * w.im)),  im = fma(u.im, w.re, u.re * w.im).
// Written so that both instructions take the twiddle as a BROADCAST scalar (w.im for the FMUL2, w.re for the
// FFMA2): a compile-time twiddle is then a 32-bit immediate / a single register instead of a 64-bit register
// pair that has to be materialised, and a table twiddle is used straight from the pair it was loaded into.
// The swap of u sits on the FMUL2 operand and the half negation on the FFMA2 addend (operand modifiers).
SSDR_DEV float2 cmul(float2 u, float2 w) {
    const float2 t = __fmul2_rn(make_float2(u.y, u.x), make_float2(w.y, w.y));      // {im*w.im, re*w.im}
    return __ffma2_rn(u, make_float2(w.x, w.x), make_float2(-t.x, t.y));
}
SSDR_DEV float2 mul_mi(float2 u) { return make_float2(u.y, -u.x); }                      // * (-i), exact
// Odd eighth turns are ordinary complex multiplies by the rounded constants.  (A separate scale by B
// after an add would be an FMUL2 feeding FADD2s, which ptxas contracts into FFMA2 even for mul.rn /
// add.rn -- one rounding less than the spec.  Every multiply in this file therefore feeds an explicit
// fused multiply-add addend, where no further contraction is possible.)
SSDR_DEV float2 mul_w8(float2 u) { return cmul(u, make_float2(kB, -kB)); }               // * B(1-i)  = W8^1
SSDR_DEV float2 mul_w83(float2 u) { return cmul(u, make_float2(-kB, -kB)); }             // * -B(1+i) = W8^3

// ---- first butterfly level: (x[m], x[m + R/2]) <- (sum, difference), m < R/2 ------------------------
template <int R>
SSDR_DEV void l1(float2 (&x)[R]) {
#pragma unroll
    for (int m = 0; m < R / 2; ++m) {
        const float2 a = x[m], b = x[m + R / 2];
        x[m] = cadd(a, b);
        x[m + R / 2] = csub(a, b);
    }
}

// Windowed first level (first pass only).  The pair (x[m], x[m+H]) are the samples n and n + N/2 of the
// frame and the periodic Hann window satisfies w[n + N/2] = 1 - w[n], so with w = w[n]
//     x[m]   <- fma(x[m] - x[m+H], w,  x[m+H])   ( = x[m] w + x[m+H] (1 - w) )
//     x[m+H] <- fma(x[m] + x[m+H], w, -x[m+H])   ( = x[m] w - x[m+H] (1 - w) )
template <int R>
SSDR_DEV void l1_window(float2 (&x)[R], const float (&w)[R / 2]) {
    constexpr int H = R / 2;
#pragma unroll
    for (int m = 0; m < H; ++m) {
        const float2 a = x[m], b = x[m + H];
        const float2 d = csub(a, b), s = cadd(a, b);
        const float2 ww = make_float2(w[m], w[m]);
        x[m] = __ffma2_rn(d, ww, b);
        x[m + H] = __ffma2_rn(s, ww, make_float2(-b.x, -b.y));
    }
}

SSDR_DEV void dft2(float2& x0, float2& x1) {
    const float2 a = x0, b = x1;
    x0 = cadd(a, b);
    x1 = csub(a, b);
}

// second level of a radix-4: in (a, c, b, e) = (x0+x2, x1+x3, x0-x2, x1-x3) at (x0, x1, x2, x3)
SSDR_DEV void dft4_rest(float2& x0, float2& x1, float2& x2, float2& x3) {
    const float2 a = x0, c = x1, b = x2, e = x3;
    x0 = cadd(a, c);
    x2 = csub(a, c);
    x1 = __fadd2_rn(b, make_float2(e.y, -e.x));    // b + (-i) e
    x3 = __fadd2_rn(b, make_float2(-e.y, e.x));    // b - (-i) e
}

SSDR_DEV void dft4(float2& x0, float2& x1, float2& x2, float2& x3) {
    const float2 a = cadd(x0, x2), b = csub(x0, x2);
    const float2 c = cadd(x1, x3), e = csub(x1, x3);
    x0 = a; x1 = c; x2 = b; x3 = e;
    dft4_rest(x0, x1, x2, x3);
}

// after the first level: x0..x3 sums, x4..x7 differences; natural order out
SSDR_DEV void dft8_rest(float2& x0, float2& x1, float2& x2, float2& x3, float2& x4, float2& x5, float2& x6, float2& x7) {
    x5 = mul_w8(x5);
    x6 = mul_mi(x6);
    x7 = mul_w83(x7);
    dft4(x0, x1, x2, x3);   // outputs q = 0, 2, 4, 6
    dft4(x4, x5, x6, x7);   // outputs q = 1, 3, 5, 7
    const float2 y1 = x4, y2 = x1, y3 = x5, y4 = x2, y5 = x6, y6 = x3;
    x1 = y1; x2 = y2; x3 = y3; x4 = y4; x5 = y5; x6 = y6;
}

// Remainder of a radix-R butterfly after its first level; natural order out: x[q] = sum_m x[m] W_R^(m q).
template <int R>
SSDR_DEV void dft_rest(float2 (&x)[R]);

template <>
SSDR_DEV void dft_rest<2>(float2 (&x)[2]) {}

template <>
SSDR_DEV void dft_rest<4>(float2 (&x)[4]) { dft4_rest(x[0], x[1], x[2], x[3]); }

template <>
SSDR_DEV void dft_rest<8>(float2 (&x)[8]) { dft8_rest(x[0], x[1], x[2], x[3], x[4], x[5], x[6], x[7]); }

template <>
SSDR_DEV void dft_rest<16>(float2 (&x)[16]) {
    // radix-4 across (m0, m0+4, m0+8, m0+12) with its first level (pairs m, m + 8) done; result p in x[m0 + 4p]
#pragma unroll
    for (int m0 = 0; m0 < 4; ++m0) dft4_rest(x[m0], x[m0 + 4], x[m0 + 8], x[m0 + 12]);
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

template <>
SSDR_DEV void dft_rest<32>(float2 (&x)[32]) {
    // 32 = 4 x 8, m = m0 + 8 m1.  radix-4 across m1 with its first level (pairs m, m + 16) done; result p in x[m0 + 8p]
#pragma unroll
    for (int m0 = 0; m0 < 8; ++m0) dft4_rest(x[m0], x[m0 + 8], x[m0 + 16], x[m0 + 24]);
    // internal twiddles W32^(m0*p)
#pragma unroll
    for (int p = 1; p < 4; ++p)
#pragma unroll
        for (int m0 = 1; m0 < 8; ++m0) {
            // W32^e = (-i)^(e / 8) W32^(e % 8): the quarter turns are exact swaps / negations (operand
            // modifiers of the consuming add), so only seven distinct constants are live.  Bit-identical
            // to multiplying by the full-circle constant because the unit32 table is built by symmetry.
            const int e = m0 * p, r = e & 7, k = e >> 3;
            float2 v = x[m0 + 8 * p];
            if (r) v = cmul(v, make_float2(unit32_cos(r), -unit32_sin(r)));
            if (k == 1) v = mul_mi(v);
            else if (k == 2) v = make_float2(-v.x, -v.y);
            x[m0 + 8 * p] = v;
        }
    // stage 2: radix-8 across m0 for each p; output s lives in x[8p + s] -> q = p + 4s
#pragma unroll
    for (int p = 0; p < 4; ++p) {
        dft2(x[8 * p], x[8 * p + 4]); dft2(x[8 * p + 1], x[8 * p + 5]); dft2(x[8 * p + 2], x[8 * p + 6]); dft2(x[8 * p + 3], x[8 * p + 7]);
        dft8_rest(x[8 * p], x[8 * p + 1], x[8 * p + 2], x[8 * p + 3], x[8 * p + 4], x[8 * p + 5], x[8 * p + 6], x[8 * p + 7]);
    }
    float2 y[32];
#pragma unroll
    for (int p = 0; p < 4; ++p)
#pragma unroll
        for (int s = 0; s < 8; ++s) y[p + 4 * s] = x[8 * p + s];
#pragma unroll
    for (int q = 0; q < 32; ++q) x[q] = y[q];
}

template <int R>
SSDR_DEV void dft(float2 (&x)[R]) {
    l1<R>(x);
    dft_rest<R>(x);
}

// Twiddles of a "chain" pass: output q = 4a + b is multiplied by W^(j q) as
// (x * A[a]) * B[b] with B[1] = w1 = W^j, B[2] = w1 w1, B[3] = B[2] w1, A[1] = B[2] B[2], A[2] = A[1] A[1],
// A[3] = A[2] A[1]: six live twiddles instead of R - 1, the same number of complex multiplies.
template <int R>
SSDR_DEV void tw_two_level(float2 (&x)[R], float2 w1) {
    static_assert(R <= 16, "chain passes have radix <= 16 (the plan keeps radix-32 passes on tables)");
    if constexpr (R == 2) {
        x[1] = cmul(x[1], w1);
    } else {
        float2 B[4], A[4];
        B[1] = w1;
        B[2] = cmul(w1, w1);
        B[3] = cmul(B[2], w1);
        if constexpr (R > 4) A[1] = cmul(B[2], B[2]);
        if constexpr (R > 8) { A[2] = cmul(A[1], A[1]); A[3] = cmul(A[2], A[1]); }
#pragma unroll
        for (int q = 1; q < R; ++q) {
            if (q >> 2) x[q] = cmul(x[q], A[q >> 2]);
            if (q & 3) x[q] = cmul(x[q], B[q & 3]);
        }
    }
}

}  // namespace ssdr
