// v2 butterflies: radix-2 levels with the twiddle BEFORE the add, natural order in, bit-reversed
// order out, every butterfly fused into THREE packed fused multiply-adds (inter-pass twiddles included):
//     t = a + i Im(T) (swap b) ,  X = t + Re(T) b = a + T b ,  Y = 2 a - X = a - T b.
// The plan of the v1 butterflies (fft_radix.cuh: adds first, twiddles after: 2 FADD2 + a 2-instruction complex
// multiply per butterfly, separate inter-pass twiddle multiplies) costs 770 packed instructions per 32 points of a
// 16384-point frame; this one 624.  Bit-exact CPU statement: oracle/c/ssdr_oracle.c fft_v2().
//
// A pass holds 2^L elements x[m] of one frame in registers (element m at position base + m * stride) and runs L
// consecutive levels in place: level l (1..L) pairs x[m], x[m + h], h = 2^(L - l), inside local blocks mu = m / (2 h).
// With B the index of the pass's elements among the blocks of the levels above it (s0 levels done before), the
// twiddle of (l, mu) is the table entry W_N^e,
//     e = (bitrev_{l-1}(mu) 2^s0 + bitrev_{s0}(B)) (N >> (s0 + l)),
// and blocks come in pairs: mu = 2 g uses T, mu = 2 g + 1 uses -i T (the table is built by exact symmetry), so a pass
// needs 1 + 1 + 2 + 4 + 8 = 16 twiddles for L = 5: index 0 for l = 1, 2^(l-2) + g for l >= 2.
//   * first pass (s0 = 0, B = 0): compile-time constants; blocks 0 / 1 of every level are the trivial 1 / -i.
//   * middle pass: B = the warp's 1024-point block -> a small shared-memory table, warp-uniform (broadcast) loads.
//   * last pass: B = the thread's 32-point block -> 16 twiddles per thread, loop invariant over the frames, parked in
//     TENSOR MEMORY (thread-private columns written once with tcgen05.st, read with tcgen05.ld).
// Every packed instruction is exactly the two scalar float32 operations the spec states (-fmad=false for this TU).
#pragma once
#include <cuda_runtime.h>

#include "fft_radix.cuh"

namespace ssdr {

SSDR_DEV constexpr int bitrev_c(int v, int bits) {
    int r = 0;
    for (int i = 0; i < bits; ++i) r |= ((v >> i) & 1) << (bits - 1 - i);
    return r;
}

// ---- butterflies ---------------------------------------------------------------------------------
SSDR_DEV void bf_one(float2& a, float2& b) {                       // T = 1 (first-pass levels only)
    const float2 x = cadd(a, b), y = csub(a, b);
    a = x; b = y;
}
SSDR_DEV void bf_mi(float2& a, float2& b) {                        // T = -i: T b = (b.im, -b.re)
    const float2 x = __fadd2_rn(a, make_float2(b.y, -b.x)), y = __fadd2_rn(a, make_float2(-b.y, b.x));
    a = x; b = y;
}
// T = (c, s) (table entry: c = cos, s = -sin):
//   t = (fma(-s, b.im, a.re), fma(s, b.re, a.im));  X = (fma(c, b.re, t.re), fma(c, b.im, t.im));  Y = fma(2, a, -X)
// packed: t2 = fma(swap(b), s, (-a.re, a.im)) = (-t.re, t.im) exactly (sign symmetry of round-to-nearest)
SSDR_DEV void bf_even(float2& a, float2& b, float c, float s) {
    const float2 t2 = __ffma2_rn(make_float2(b.y, b.x), make_float2(s, s), make_float2(-a.x, a.y));
    const float2 x = __ffma2_rn(b, make_float2(c, c), make_float2(-t2.x, t2.y));
    const float2 y = __ffma2_rn(a, make_float2(2.0f, 2.0f), make_float2(-x.x, -x.y));
    a = x; b = y;
}
// T' = -i T = (s, -c):  t = (fma(c, b.im, a.re), fma(-c, b.re, a.im));  X = (fma(s, b.re, t.re), fma(s, b.im, t.im))
// packed: t2 = fma(swap(b), c, (a.re, -a.im)) = (t.re, -t.im)
SSDR_DEV void bf_odd(float2& a, float2& b, float c, float s) {
    const float2 t2 = __ffma2_rn(make_float2(b.y, b.x), make_float2(c, c), make_float2(a.x, -a.y));
    const float2 x = __ffma2_rn(b, make_float2(s, s), make_float2(t2.x, -t2.y));
    const float2 y = __ffma2_rn(a, make_float2(2.0f, 2.0f), make_float2(-x.x, -x.y));
    a = x; b = y;
}

// ---- first pass: compile-time twiddles (s0 = 0, B = 0), levels FIRST..L of an L-level pass ---------------------
// (level 1 is done by the caller when the window is folded into it).  The level number is a template parameter so that
// every trip count and register index below is a compile-time constant.
template <int L, int l>
SSDR_DEV void level_const(float2 (&x)[1 << L]) {
    constexpr int h = 1 << (L - l), NB = 1 << (l - 1);
#pragma unroll
    for (int mu = 0; mu < NB; ++mu) {
        const int g = mu >> 1;
        // W_{2^l}^{bitrev_{l-1}(2 g)} = W_32^{m32},  m32 = bitrev_{l-2}(g) (32 >> l)
        const int m32 = (l >= 2) ? bitrev_c(g, l - 2) * (32 >> l) : 0;
        const float c = unit32_cos(m32), s = -unit32_sin(m32);
#pragma unroll
        for (int j = 0; j < h; ++j) {
            float2& a = x[mu * 2 * h + j];
            float2& b = x[mu * 2 * h + j + h];
            if (g == 0) { if (mu & 1) bf_mi(a, b); else bf_one(a, b); }
            else if (mu & 1) bf_odd(a, b, c, s);
            else bf_even(a, b, c, s);
        }
    }
}
template <int L, int FIRST>
SSDR_DEV void levels_const(float2 (&x)[1 << L]) {
    if constexpr (FIRST <= L) {
        level_const<L, FIRST>(x);
        levels_const<L, FIRST + 1>(x);
    }
}

// ---- table passes: twiddle index 0 for level 1, 2^(l-2) + g for level l >= 2 ----------------------------------
// GET(idx) returns the float2 table entry; every entry is fetched once.
template <int L, int l, class GET>
SSDR_DEV void level_table(float2 (&x)[1 << L], GET& get) {
    constexpr int h = 1 << (L - l);
    if constexpr (l == 1) {
        const float2 T = get(0);
#pragma unroll
        for (int j = 0; j < h; ++j) bf_even(x[j], x[j + h], T.x, T.y);
    } else {
        constexpr int NG = 1 << (l - 2);
#pragma unroll
        for (int g = 0; g < NG; ++g) {
            const float2 T = get(NG + g);
#pragma unroll
            for (int j = 0; j < h; ++j) {
                bf_even(x[(2 * g) * 2 * h + j], x[(2 * g) * 2 * h + j + h], T.x, T.y);
                bf_odd(x[(2 * g + 1) * 2 * h + j], x[(2 * g + 1) * 2 * h + j + h], T.x, T.y);
            }
        }
    }
}
template <int L, int LFIRST, int LLAST, class GET>
SSDR_DEV void levels_table(float2 (&x)[1 << L], GET get) {
    if constexpr (LFIRST <= LLAST) {
        level_table<L, LFIRST>(x, get);
        levels_table<L, LFIRST + 1, LLAST>(x, get);
    }
}

// Table entry of (level l, pair g) for a pass that starts after s0 levels with outer block index B:
// exponent e = (bitrev_{l-2}(g) 2^s0 + bitrev_{s0}(B)) (N >> (s0 + l))   (mu = 2 g: bitrev_{l-1}(2 g) = bitrev_{l-2}(g))
SSDR_DEV int tw_exponent(int lgN, int s0, unsigned B, int l, int g) {
    const unsigned brB = s0 ? (__brev(B) >> (32 - s0)) : 0u;
    const unsigned brg = (l >= 3) ? (__brev((unsigned)g) >> (32 - (l - 2))) : 0u;
    return (int)(((brg << s0) + brB) << (lgN - s0 - l));
}
// idx 0..15 -> (l, g)
SSDR_DEV void tw_index_to_lg(int idx, int& l, int& g) {
    if (idx == 0) { l = 1; g = 0; }
    else { l = 33 - __clz(idx); g = idx - (1 << (l - 2)); }      // idx in [2^(l-2), 2^(l-1))
}

// ---- tensor memory as thread-private scratch ---------------------------------------------------------
// tcgen05.st / tcgen05.ld .32x32b: lane i of the warp <-> TMEM lane (32 (warp % 4) + i), consecutive registers <->
// consecutive columns.  Warps w, w + 4, w + 8, .. share a lane quarter and take disjoint column ranges.
SSDR_DEV unsigned tmem_alloc_cols(unsigned* slot_smem, int cols) {      // one warp calls; cols: power of two >= 32
    const unsigned sa = (unsigned)__cvta_generic_to_shared(slot_smem);
    switch (cols) {
        case 32: asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 32;" ::"r"(sa)); break;
        case 64: asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 64;" ::"r"(sa)); break;
        case 128: asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 128;" ::"r"(sa)); break;
        case 256: asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"(sa)); break;
        default: asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(sa)); break;
    }
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    return 0;
}
SSDR_DEV void tmem_free_cols(unsigned base, int cols) {
    switch (cols) {
        case 32: asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 32;" ::"r"(base)); break;
        case 64: asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 64;" ::"r"(base)); break;
        case 128: asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 128;" ::"r"(base)); break;
        case 256: asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" ::"r"(base)); break;
        default: asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(base)); break;
    }
}
SSDR_DEV void tmem_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
SSDR_DEV void tmem_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
SSDR_DEV void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
SSDR_DEV void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

SSDR_DEV void tmem_st8(unsigned addr, const unsigned (&v)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                 ::"r"(addr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]) : "memory");
}
SSDR_DEV void tmem_ld8(unsigned addr, unsigned (&v)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]) : "r"(addr) : "memory");
}
SSDR_DEV void tmem_st16(unsigned addr, const unsigned (&v)[16]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
                 ::"r"(addr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
                 "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]) : "memory");
}
SSDR_DEV void tmem_ld16(unsigned addr, unsigned (&v)[16]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
                   "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]) : "r"(addr) : "memory");
}

}  // namespace ssdr
