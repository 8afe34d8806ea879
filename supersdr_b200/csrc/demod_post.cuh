// Demodulator back end shared by the two FIR engines (demod_kernels.cu: FFMA2 FIR; demod_tc_kernels.cu: tcgen05 FIR):
// one warp holds one 512-sample frame of FIR output, 16 consecutive samples per LOGICAL lane, and runs
// magnitude / RSSI -> AM / SSB / CW / NBFM detector -> AGC -> float32 + int16 PCM, carrying the per-channel streaming
// state (parameter model utils_supersdr.py:936-945,1022-1029; SND header s-meter utils_supersdr.py:1068-1069).
//
// The two engines differ in which physical lane holds which 16 samples, so every lane-order-dependent step (prefix
// scans, carries, the sample index) goes through a lane map LM; order-free reductions use xor shuffles directly.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

#include "demod_host.h"

namespace ssdr {

constexpr int kDemodSpl = SSDR_FRAME / 32;    // 16 samples per lane

// MUFU approximations (relative error ~1e-7, far inside the 1e-5 RMS tolerance of the demodulator): no denormal / range
// fix-up code and no slow-path calls on the per-sample path
__device__ __forceinline__ float ex2_approx(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float lg2_approx(float x) { float y; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float sqrt_approx(float x) { float y; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

// atan2 for the NBFM detector: octant reduction (one approximate division), odd polynomial of degree 15 fitted on [0, 1]
// (Chebyshev interpolation of atan(sqrt s) / sqrt s, degree 7 in s = t^2): max abs error 1.5e-7 rad in float32 (checked over
// 2 M points), + 2 ulp of the division -- far inside the demodulator's 1e-5 RMS tolerance; ~20 instructions against the
// library's ~45 (which made NBFM the dearest mode of the bank: 129 against 186 Gsamples/s for USB).  (0, 0) is handled by the caller.
__device__ __forceinline__ float atan2_fast(float y, float x) {
    const float ax = fabsf(x), ay = fabsf(y);
    const float mx = fmaxf(ax, ay), mn = fminf(ax, ay);
    const float t = __fdividef(mn, mx);
    const float s = t * t;
    float p = -0.00455979211255908f;
    p = fmaf(p, s, 0.023780519142746925f);
    p = fmaf(p, s, -0.05882975459098816f);
    p = fmaf(p, s, 0.09868865460157394f);
    p = fmaf(p, s, -0.14003290235996246f);
    p = fmaf(p, s, 0.19966961443424225f);
    p = fmaf(p, s, -0.3333181142807007f);
    p = fmaf(p, s, 0.9999998807907104f);
    float r = p * t;
    if (ay > ax) r = 1.57079637f - r;
    if (x < 0.0f) r = 3.14159274f - r;
    return copysignf(r, y);
}

// Two samples at once, scaled by K = 32767 / pi (the NBFM detector's output unit): the reduction to t = min / max in [0, 1]
// is scalar (one MUFU.RCP each), the polynomial runs as packed fp32x2 FFMA2 on (t0, t1) with the coefficients pre-multiplied
// by K, the octant fix-ups use K pi / 2 and K pi.  Same approximation as atan2_fast (the scaled coefficients round once more:
// +1 ulp).  (0, 0) -> 0 is handled here: min = max = 0 gives t = 0 * inf = NaN, so the caller's zero test selects 0.
__device__ __forceinline__ float2 atan2_fast2_scaled(float y0, float x0, float y1, float x1) {
    constexpr float K = 32767.0f / 3.14159265358979f;
    const float mx0 = fmaxf(fabsf(x0), fabsf(y0)), mn0 = fminf(fabsf(x0), fabsf(y0));
    const float mx1 = fmaxf(fabsf(x1), fabsf(y1)), mn1 = fminf(fabsf(x1), fabsf(y1));
    float rc0, rc1;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rc0) : "f"(mx0));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rc1) : "f"(mx1));
    const float2 t = __fmul2_rn(make_float2(mn0, mn1), make_float2(rc0, rc1));
    const float2 q = __fmul2_rn(t, t);
    float2 p = make_float2(-0.00455979211255908f * K, -0.00455979211255908f * K);
    p = __ffma2_rn(p, q, make_float2(0.023780519142746925f * K, 0.023780519142746925f * K));
    p = __ffma2_rn(p, q, make_float2(-0.05882975459098816f * K, -0.05882975459098816f * K));
    p = __ffma2_rn(p, q, make_float2(0.09868865460157394f * K, 0.09868865460157394f * K));
    p = __ffma2_rn(p, q, make_float2(-0.14003290235996246f * K, -0.14003290235996246f * K));
    p = __ffma2_rn(p, q, make_float2(0.19966961443424225f * K, 0.19966961443424225f * K));
    p = __ffma2_rn(p, q, make_float2(-0.3333181142807007f * K, -0.3333181142807007f * K));
    p = __ffma2_rn(p, q, make_float2(0.9999998807907104f * K, 0.9999998807907104f * K));
    float2 r = __fmul2_rn(p, t);
    if (fabsf(y0) > fabsf(x0)) r.x = 1.57079637f * K - r.x;
    if (fabsf(y1) > fabsf(x1)) r.y = 1.57079637f * K - r.y;
    if (x0 < 0.0f) r.x = 3.14159274f * K - r.x;
    if (x1 < 0.0f) r.y = 3.14159274f * K - r.y;
    return make_float2(copysignf(r.x, y0), copysignf(r.y, y1));
}

// cos / sin of a 32-bit phase (2 pi phase / 2^32), MUFU path: abs error ~4e-7
__device__ __forceinline__ void nco(unsigned ph, float& c, float& s) {
    float a = (float)(int)ph * 1.4629180792671596e-9f;   // 2 pi / 2^32
    __sincosf(a, &s, &c);
}

// Complex products as two packed fp32x2 instructions (FMUL2 + FFMA2; the half swap, the broadcast and the half negation are
// operand modifiers on sm_100a) instead of four scalar ones: the NCO rotations and the mixer are a quarter of the
// demodulator's per-sample instructions.
__device__ __forceinline__ float2 cmul2(float2 u, float2 w) {            // u w
    const float2 t = __fmul2_rn(make_float2(u.y, u.x), make_float2(w.y, w.y));
    return __ffma2_rn(u, make_float2(w.x, w.x), make_float2(-t.x, t.y));
}
__device__ __forceinline__ float2 cmulc2(float2 u, float2 w) {           // u conj(w)
    const float2 t = __fmul2_rn(make_float2(u.y, u.x), make_float2(w.y, w.y));
    return __ffma2_rn(u, make_float2(w.x, w.x), make_float2(t.x, -t.y));
}

// logical lane = physical lane
struct LanesNatural {
    __device__ __forceinline__ static int logical(int lane) { return lane; }
    template <class V> __device__ __forceinline__ static V up(V v, int d) { return __shfl_up_sync(0xffffffffu, v, d); }
    template <class V> __device__ __forceinline__ static V from(V v, int l) { return __shfl_sync(0xffffffffu, v, l); }
};
// tcgen05 engine: TMEM lane i < 16 holds the real parts of 32-sample block i, lane 16 + i its imaginary parts; after the
// pair exchange physical lane p holds the complex samples of logical lane 2 (p & 15) + (p >> 4)
struct LanesPaired {
    __device__ __forceinline__ static int logical(int lane) { return ((lane & 15) << 1) | (lane >> 4); }
    __device__ __forceinline__ static int phys(int l) { return ((l >> 1) & 15) | ((l & 1) << 4); }
    template <class V> __device__ __forceinline__ static V up(V v, int d) {        // callers ignore the result when logical < d
        return __shfl_sync(0xffffffffu, v, phys((logical(threadIdx.x & 31) - d) & 31));
    }
    template <class V> __device__ __forceinline__ static V from(V v, int l) { return __shfl_sync(0xffffffffu, v, phys(l)); }
};

// per-channel streaming state held in registers while a warp walks the frames of its channel
struct DemodRegs {
    unsigned ph1, ph2;
    float e_in;
    double dc;
    float2 zprev;
    unsigned blk;
    float ring;               // physical lane < SSDR_HANG_BLOCKS holds one slot of the hang ring
};

__device__ __forceinline__ void demod_regs_load(DemodRegs& st, const DemodState* stp, int lane) {
    // L2 loads: within one launch of the FFMA engine the state of a channel may have been written by another SM (time slices)
    st.ph1 = __ldcg(&stp->ph1); st.ph2 = __ldcg(&stp->ph2);
    st.e_in = __ldcg(&stp->e_in);
    st.dc = __ldcg(&stp->dc);
    st.zprev = make_float2(__ldcg(&stp->zprev_re), __ldcg(&stp->zprev_im));
    st.blk = __ldcg(&stp->blk);
    st.ring = (lane < SSDR_HANG_BLOCKS) ? __ldcg(&stp->ring[lane]) : 0.0f;
}
__device__ __forceinline__ void demod_regs_store(const DemodRegs& st, DemodState* stp, int lane) {
    if (lane < SSDR_HANG_BLOCKS) stp->ring[lane] = st.ring;
    if (lane == 0) {
        stp->ph1 = st.ph1; stp->ph2 = st.ph2; stp->e_in = st.e_in; stp->dc = st.dc;
        stp->zprev_re = st.zprev.x; stp->zprev_im = st.zprev.y; stp->blk = st.blk;
    }
}

// acc[r] = FIR output sample 16 L + r of frame b of channel ch (L = logical lane); s0 = index of the frame's first sample
// in the pcm arrays.  Advances st by one frame.
template <class LM>
__device__ __forceinline__ void demod_frame_tail(const float2 (&acc)[kDemodSpl], const DemodChan& cp, const DemodKernelParams& kp,
                                                 int ch, int b, size_t s0, DemodRegs& st) {
    constexpr int SPL = kDemodSpl, FR = SSDR_FRAME;
    const int lane = LM::logical(threadIdx.x & 31);
    // ---- power, RSSI, block peak ---------------------------------------------------------------
    // Everything that only compares or takes logarithms of |z| works on the POWER |z|^2 (round 2): the square root is
    // monotonic, so the block peak is one sqrt of the largest power, and the AGC (below) takes log2 of powers; only the AM
    // detector needs the 16 magnitudes.
    float pw[SPL];
    float pmax = 0.f;
    float2 ps2 = make_float2(0.f, 0.f);
#pragma unroll
    for (int r = 0; r < SPL; r += 2) {
        pw[r] = acc[r].x * acc[r].x + acc[r].y * acc[r].y;
        pw[r + 1] = acc[r + 1].x * acc[r + 1].x + acc[r + 1].y * acc[r + 1].y;
        ps2 = __fadd2_rn(ps2, make_float2(pw[r], pw[r + 1]));
        pmax = fmaxf(pmax, fmaxf(pw[r], pw[r + 1]));
    }
    float psum = ps2.x + ps2.y;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) pmax = fmaxf(pmax, __shfl_xor_sync(0xffffffffu, pmax, o));
    const float bmax = sqrt_approx(pmax);
    if (kp.rssi) {                                           // warp-uniform; no divergent library call on the frame's critical path
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) psum += __shfl_xor_sync(0xffffffffu, psum, o);
        // 10 log10(mean |z|^2 / FS^2) + FS_DBM with FS = 2^15; lg2.approx: ~1e-6 dB (the tests allow 1e-3)
        static_assert(SSDR_FS == 32768.0f, "RSSI and AGC constants assume FS = 2^15");
        const float db = fmaf(lg2_approx(fmaxf(psum * (1.0f / FR), 1e-30f)) - 30.0f, 3.0102999566398120f, kDemodFsDbm);
        if (lane == 0) kp.rssi[(size_t)ch * (kp.pitch / FR) + b] = db;
    }
    // ---- detector --------------------------------------------------------------------------
    float a[SPL];
    if (cp.mode == SSDR_MODE_NBFM) {
        float2 last = acc[SPL - 1];
        float2 prv = make_float2(LM::up(last.x, 1), LM::up(last.y, 1));
        if (lane == 0) prv = st.zprev;
#ifndef SSDR_DEMOD_LIB_ATAN2
#pragma unroll
        for (int r = 0; r < SPL; r += 2) {                 // two samples per step: packed products and polynomial
            const float2 d0 = cmulc2(acc[r], prv), d1 = cmulc2(acc[r + 1], acc[r]);      // z conj(prev) = (re, im)
            const float2 ph = atan2_fast2_scaled(d0.y, d0.x, d1.y, d1.x);
            // a zero product (first sample of a stream, or silence) demodulates to 0, not +-pi
            a[r] = (d0.x == 0.0f && d0.y == 0.0f) ? 0.0f : ph.x;
            a[r + 1] = (d1.x == 0.0f && d1.y == 0.0f) ? 0.0f : ph.y;
            prv = acc[r + 1];
        }
#else
#pragma unroll
        for (int r = 0; r < SPL; ++r) {
            float2 z = acc[r];
            float re = z.x * prv.x + z.y * prv.y;     // z * conj(prev)
            float im = z.y * prv.x - z.x * prv.y;
            a[r] = (re == 0.0f && im == 0.0f) ? 0.0f : atan2f(im, re) * (32767.0f / 3.14159265358979f);
            prv = z;
        }
#endif
    } else if (cp.mode == SSDR_MODE_AM) {
        // carrier tracker dc[k] = dc[k-1] + beta (mag[k] - dc[k-1]): lane-local recurrence from a zero (lane 0: true) carry-in,
        // then an affine warp scan.  SSDR_DEMOD_AM_F64 = 1 keeps the float64 recurrence of round 1; the default is float32
        // (round 2): the tracker is a leaky integrator (it forgets rounding errors with its time constant), its steady-state
        // error is ~4e-7 of the carrier, the same order as the float32 rounding of |z2| that enters it -- and the float64
        // conversions and DFMA chains made AM 1.4 x as dear as USB.
#ifndef SSDR_DEMOD_AM_F64
#define SSDR_DEMOD_AM_F64 0
#endif
        float mag[SPL];
#pragma unroll
        for (int r = 0; r < SPL; ++r) mag[r] = sqrt_approx(pw[r]);
#if SSDR_DEMOD_AM_F64
        double B = (lane == 0) ? st.dc : 0.0;
#pragma unroll
        for (int r = 0; r < SPL; ++r) B = B + kDemodAmBeta * ((double)mag[r] - B);
#pragma unroll
        for (int s = 0; s < 5; ++s) {
            double up = LM::up(B, 1 << s);
            if (lane >= (1 << s)) B = B + kp.am_pow16[s] * up;   // (om^16)^(2^s)
        }
        double carry = LM::up(B, 1);
        if (lane == 0) carry = st.dc;
        st.dc = LM::from(B, 31);
        // B after the scan is the carrier at the end of this lane's segment; replay with the carry-in
        double d = carry;
#pragma unroll
        for (int r = 0; r < SPL; ++r) {
            d = d + kDemodAmBeta * ((double)mag[r] - d);
            a[r] = (float)((double)mag[r] - d);
        }
#else
        constexpr float beta = (float)kDemodAmBeta;
        float B = (lane == 0) ? (float)st.dc : 0.0f;
#pragma unroll
        for (int r = 0; r < SPL; ++r) B = fmaf(beta, mag[r] - B, B);
#pragma unroll
        for (int s = 0; s < 5; ++s) {
            const float up = LM::up(B, 1 << s);
            if (lane >= (1 << s)) B = fmaf((float)kp.am_pow16[s], up, B);   // (om^16)^(2^s)
        }
        float carry = LM::up(B, 1);
        if (lane == 0) carry = (float)st.dc;
        st.dc = (double)LM::from(B, 31);
        float d = carry;
#pragma unroll
        for (int r = 0; r < SPL; ++r) {
            d = fmaf(beta, mag[r] - d, d);
            a[r] = mag[r] - d;
        }
#endif
    } else {
        // one NCO evaluation per four samples (the phase accumulator is exact), the other three by rotation
        float2 r2;
        nco(cp.inc2, r2.x, r2.y);
#pragma unroll
        for (int r4 = 0; r4 < SPL; r4 += 4) {
            float2 w;
            nco(st.ph2 + (unsigned)(SPL * lane + r4) * cp.inc2, w.x, w.y);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                a[r4 + i] = acc[r4 + i].x * w.x - acc[r4 + i].y * w.y;   // Re(z * exp(+j theta2))
                if (i < 3) w = cmul2(w, r2);
            }
        }
    }
    // ---- AGC ---------------------------------------------------------------------------------
    float out[SPL];
    if (cp.mode == SSDR_MODE_NBFM) {
#pragma unroll
        for (int r = 0; r < SPL; ++r) out[r] = a[r];
    } else if (!cp.agc_on) {
#pragma unroll
        for (int r = 0; r < SPL; r += 2) {
            const float2 o = __fmul2_rn(make_float2(a[r], a[r + 1]), make_float2(cp.man_gain, cp.man_gain));
            out[r] = o.x; out[r + 1] = o.y;
        }
    } else {
        // Envelope and gain in the log2-of-POWER domain (round 2; the linear form -- u[k] = hm[k] 2^(k c2), prefix max, e[k] =
        // M[k] 2^(-k c2), gain = AGC_OUT 2^(max(log2(e / FS), knee2) (slope/100 - 1)) -- cost 13 instructions per sample,
        // this one 9, with the same two MUFU per sample and no square root):
        //   Q[k] = log2(hm[k]^2),  U[k] = Q[k] + 2 k c2,  M[k] = max(prefix max U, 2 log2(e_in) - 2 c2),  2 log2 e[k] = M[k] - 2 k c2,
        //   gain = 2^(max(2 log2 e, 2 (knee2 + 15)) (slope/100 - 1) / 2 + log2(AGC_OUT) - 15 (slope/100 - 1)).
        // k = 16 lane + r: the lane part of 2 k c2 is added for the cross-lane scan only.  Silence (log2 0 = -inf) stays -inf
        // through max / fma and ends at the knee; no inf - inf can occur.
        static_assert(kDemodAgcOut == 0.5f, "log2(AGC_OUT) = -1 below");
        float q[SPL];                                        // hm^2: hang = max(max(ring), prefix max of |z|)
        if (cp.agc_hang) {
            float hb = st.ring;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) hb = fmaxf(hb, __shfl_xor_sync(0xffffffffu, hb, o));
            float run = 0.f;
#pragma unroll
            for (int r = 0; r < SPL; ++r) { run = fmaxf(run, pw[r]); q[r] = run; }
            float excl = run;                                // inclusive scan of lane maxima
#pragma unroll
            for (int s = 0; s < 5; ++s) {
                float up = LM::up(excl, 1 << s);
                if (lane >= (1 << s)) excl = fmaxf(excl, up);
            }
            excl = LM::up(excl, 1);
            if (lane == 0) excl = 0.f;
            excl = fmaxf(excl, hb * hb);
#pragma unroll
            for (int r = 0; r < SPL; ++r) q[r] = fmaxf(q[r], excl);
        } else {
#pragma unroll
            for (int r = 0; r < SPL; ++r) q[r] = pw[r];
        }
        const float c2d = 2.0f * cp.c2;
        const float ninf = __int_as_float(0xff800000);
        float W[SPL];
        float w = ninf;
#pragma unroll
        for (int r = 0; r < SPL; ++r) {
            w = fmaxf(w, fmaf((float)r, c2d, lg2_approx(q[r])));
            W[r] = w;
        }
        const float base = (float)(SPL * lane) * c2d;
        float pre = w + base;
#pragma unroll
        for (int s = 0; s < 5; ++s) {
            float up = LM::up(pre, 1 << s);
            if (lane >= (1 << s)) pre = fmaxf(pre, up);
        }
        pre = LM::up(pre, 1);
        if (lane == 0) pre = ninf;
        pre = fmaxf(pre, fmaf(2.0f, lg2_approx(st.e_in), -c2d)) - base;
        const float knee = 2.0f * (cp.knee2 + 15.0f), gs = 0.5f * cp.slope_m1, gc = fmaf(-15.0f, cp.slope_m1, -1.0f);
        float e2_last = ninf;
#pragma unroll
        for (int r = 0; r < SPL; r += 2) {
            const float e2a = fmaf(-(float)r, c2d, fmaxf(W[r], pre)), e2b = fmaf(-(float)(r + 1), c2d, fmaxf(W[r + 1], pre));
            const float2 o = __fmul2_rn(make_float2(a[r], a[r + 1]), make_float2(ex2_approx(fmaf(fmaxf(e2a, knee), gs, gc)),
                                                                                   ex2_approx(fmaf(fmaxf(e2b, knee), gs, gc))));
            out[r] = o.x; out[r + 1] = o.y;
            e2_last = e2b;
        }
        st.e_in = ex2_approx(0.5f * LM::from(e2_last, 31));
    }
    // ---- outputs: 16 consecutive samples per lane ----------------------------------------------
    const size_t o0 = s0 + (size_t)SPL * lane;
    if (kp.pcm_f32) {
        float4* p = reinterpret_cast<float4*>(kp.pcm_f32 + o0);
#pragma unroll
        for (int r = 0; r < SPL; r += 4) __stcs(p + r / 4, make_float4(out[r], out[r + 1], out[r + 2], out[r + 3]));
    }
    if (kp.pcm_i16) {
        unsigned pk[SPL / 2];
#pragma unroll
        for (int r = 0; r < SPL; r += 2) {
            short v0, v1;
            asm("cvt.rni.sat.s16.f32 %0, %1;" : "=h"(v0) : "f"(out[r]));
            asm("cvt.rni.sat.s16.f32 %0, %1;" : "=h"(v1) : "f"(out[r + 1]));
            pk[r / 2] = ((unsigned)(unsigned short)v0) | ((unsigned)(unsigned short)v1 << 16);
        }
        uint4* p = reinterpret_cast<uint4*>(kp.pcm_i16 + o0);
        __stcs(p, make_uint4(pk[0], pk[1], pk[2], pk[3]));
        __stcs(p + 1, make_uint4(pk[4], pk[5], pk[6], pk[7]));
    }
    // ---- per-frame state ---------------------------------------------------------------------------
    st.zprev = make_float2(LM::from(acc[SPL - 1].x, 31), LM::from(acc[SPL - 1].y, 31));
    if ((threadIdx.x & 31) == (int)(st.blk % SSDR_HANG_BLOCKS)) st.ring = bmax;
    st.blk++;
    st.ph1 += (unsigned)FR * cp.inc1;
    st.ph2 += (unsigned)FR * cp.inc2;
}

// sample load (K6 fused): complex64 or Kiwi big-endian int16 pairs (kiwi/client.py:449-453)
template <int FMT>
__device__ __forceinline__ float2 demod_ld_iq(const void* base, size_t idx) {
    if constexpr (FMT == SSDR_IQ_CF32) {
        return __ldcs(reinterpret_cast<const float2*>(base) + idx);
    } else {
        unsigned v = __ldcs(reinterpret_cast<const unsigned*>(base) + idx);
        const unsigned sw = __byte_perm(v, 0u, 0x2301);   // swap the bytes of both 16-bit halves
        const int i = (int)(short)(sw & 0xffffu), q = (int)sw >> 16;
        return make_float2((float)i, (float)q);
    }
}

// four consecutive samples (idx a multiple of 4: 32 / 16 bytes, aligned)
template <int FMT>
__device__ __forceinline__ void demod_ld_iq4(const void* base, size_t idx, float2* x) {
    if constexpr (FMT == SSDR_IQ_CF32) {
        const float4* p = reinterpret_cast<const float4*>(reinterpret_cast<const float2*>(base) + idx);
        const float4 a = __ldcs(p), b = __ldcs(p + 1);
        x[0] = make_float2(a.x, a.y); x[1] = make_float2(a.z, a.w); x[2] = make_float2(b.x, b.y); x[3] = make_float2(b.z, b.w);
    } else {
        const uint4 v = __ldcs(reinterpret_cast<const uint4*>(reinterpret_cast<const unsigned*>(base) + idx));
        const unsigned w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const unsigned sw = __byte_perm(w[i], 0u, 0x2301);
            x[i] = make_float2((float)(int)(short)(sw & 0xffffu), (float)((int)sw >> 16));
        }
    }
}

}  // namespace ssdr
