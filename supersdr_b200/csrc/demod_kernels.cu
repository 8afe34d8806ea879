// Fused demodulator (K4): IQ @12 kHz -> NCO mix -> 127-tap FIR band-pass (real low-pass prototype
// around the pass-band centre) -> AM / USB / LSB / CW / NBFM detector -> AGC -> float32 + int16 PCM
// and the per-frame RSSI.  One kernel, no intermediate HBM traffic.
//
// Replaces the remote KiwiSDR SND computation that the reference only parametrises
// (utils_supersdr.py:1022-1029) and receives as int16 PCM (utils_supersdr.py:1044-1076).
// Arithmetic spec: DESIGN.md section 4.5; float64 oracle: oracle/tier_u.py demod().
//
// Mapping: ONE WARP OWNS ONE CHANNEL and walks its 512-sample frames in order, so all streaming
// state (NCO phases, FIR history, AGC envelope, hang ring, AM carrier, FM previous sample) stays
// in that warp and only __syncwarp() is needed.  Within a frame, the mixer uses a lane-strided
// (coalesced) mapping; the FIR, detector and AGC use a lane-contiguous mapping (16 consecutive
// samples per lane), which turns the AGC peak tracker and the AM carrier tracker into lane-local
// recurrences stitched by warp-shuffle scans.  The FIR is register-tiled: 16 float2 accumulators
// per lane, taps broadcast from shared memory, samples from a padded (conflict-free) shared tile.
#include <cuda_runtime.h>

#include <cstdint>

#include "common.cuh"
#include "demod_host.h"

namespace ssdr {

namespace {

constexpr int T = SSDR_FIR_TAPS;          // 127
constexpr int H = T - 1;                  // 126 history samples
constexpr int FR = SSDR_FRAME;            // 512
constexpr int SPL = FR / 32;              // 16 samples per lane
constexpr int ZLEN = 1 + H + FR;          // 639: [0] = permanent zero, [1..126] history, [127..638] frame
constexpr int ZPAD = ZLEN + (ZLEN >> 4) + 2;
constexpr int UT = 8;                     // taps per FIR step
constexpr int WARPS = 4;
constexpr int TAP_PAD = 128;

__device__ __forceinline__ int zpos(int i) { return i + (i >> 4); }   // 1 pad per 16: lane stride 17

template <int FMT>
__device__ __forceinline__ float2 ld_iq(const void* base, size_t idx) {
    if constexpr (FMT == SSDR_IQ_CF32) {
        return __ldcs(reinterpret_cast<const float2*>(base) + idx);
    } else {
        unsigned v = __ldcs(reinterpret_cast<const unsigned*>(base) + idx);
        const unsigned sw = __byte_perm(v, 0u, 0x2301);   // swap the bytes of both 16-bit halves
        const int i = (int)(short)(sw & 0xffffu), q = (int)sw >> 16;
        return make_float2((float)i, (float)q);
    }
}

// MUFU approximations (relative error ~1e-7, far inside the 1e-5 RMS tolerance of the demodulator): no denormal / range
// fix-up code and no slow-path calls on the per-sample path
__device__ __forceinline__ float ex2_approx(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float lg2_approx(float x) { float y; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float sqrt_approx(float x) { float y; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

// cos / sin of a 32-bit phase (2 pi phase / 2^32), MUFU path: abs error ~4e-7
__device__ __forceinline__ void nco(unsigned ph, float& c, float& s) {
    float a = (float)(int)ph * 1.4629180792671596e-9f;   // 2 pi / 2^32
    __sincosf(a, &s, &c);
}

// One FIR step = 8 taps t = 32*s4 + 8*PH + u.  Sample (r, t) sits K = 127 + r - t past the lane base
// (buffer coordinates) and is kept in register slot K & 31, so consecutive steps reuse 15 of the 23
// window samples and load only 8 new ones; PH makes every slot index a compile-time constant.
template <int PH>
__device__ __forceinline__ void fir_step(float2 (&W)[32], float2 (&acc)[SPL], const float2* zl,
                                         const float2* taps2, int s4) {
    const float2* zs = zl - 34 * s4;                       // K -> K - 32*s4: offset (K + K>>4) - 34*s4
#pragma unroll
    for (int i = 0; i < UT; ++i) {
        constexpr int K0 = 120 - 8 * PH;
        W[(K0 + i) & 31] = zs[(K0 + i) + ((K0 + i) >> 4)];
    }
    const float2* tp = taps2 + 32 * s4 + 8 * PH;
#pragma unroll
    for (int u = 0; u < UT; ++u) {
        const float2 hh = tp[u];
#pragma unroll
        for (int r = 0; r < SPL; ++r) acc[r] = __ffma2_rn(hh, W[(127 + r - u - 8 * PH) & 31], acc[r]);
    }
}

struct WarpSmem {
    float2 z[ZPAD];
    float2 taps2[TAP_PAD];   // (h, h) pairs for packed FMA
};

template <int FMT>
__global__ void __launch_bounds__(WARPS * 32, 3)
demod_kernel(const DemodKernelParams kp) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    WarpSmem* ws = reinterpret_cast<WarpSmem*>(smem_raw) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    const int warp_global = blockIdx.x * WARPS + (threadIdx.x >> 5);
    const int warps_total = gridDim.x * WARPS;
    const int nblk = kp.n_samples / FR;

    for (int ch = warp_global; ch < kp.batch; ch += warps_total) {
        const DemodChan cp = kp.chan[ch];
        DemodState* stp = kp.state + ch;
        // ---- load per-channel state ---------------------------------------------------------
        unsigned ph1 = stp->ph1, ph2 = stp->ph2;
        float e_in = stp->e_in;
        double dc = stp->dc;
        float2 zprev = make_float2(stp->zprev_re, stp->zprev_im);
        unsigned blk = stp->blk;
        float ring = (lane < SSDR_HANG_BLOCKS) ? stp->ring[lane] : 0.0f;
        __syncwarp();
        for (int i = lane; i < H; i += 32) ws->z[zpos(1 + i)] = kp.hist[(size_t)ch * H + i];
        if (lane == 0) ws->z[0] = make_float2(0.f, 0.f);
        for (int i = lane; i < TAP_PAD; i += 32) { float h = (i < T) ? kp.taps[(size_t)ch * T + i] : 0.f; ws->taps2[i] = make_float2(h, h); }
        __syncwarp();

        for (int b = 0; b < nblk; ++b) {
            const size_t s0 = (size_t)ch * kp.pitch + (size_t)b * FR;
            // ---- mixer: lane-strided, coalesced; all sixteen loads of the frame in flight together ----------
            float2 xin[SPL];
#pragma unroll
            for (int r = 0; r < SPL; ++r) xin[r] = ld_iq<FMT>(kp.iq, s0 + lane + 32 * r);
#pragma unroll
            for (int r = 0; r < SPL; ++r) {
                const int k = lane + 32 * r;
                const float2 x = xin[r];
                float c, s;
                nco(ph1 + (unsigned)k * cp.inc1, c, s);
                // x * exp(-j theta): (xr + j xi)(c - j s)
                float2 y = make_float2(x.x * c + x.y * s, x.y * c - x.x * s);
                ws->z[zpos(1 + H + k)] = y;
            }
            __syncwarp();
            // ---- FIR: out[n] = sum_t h[t] z[n - t], lane owns n = 16*lane .. 16*lane + 15 -----
            float2 acc[SPL];
#pragma unroll
            for (int r = 0; r < SPL; ++r) acc[r] = make_float2(0.f, 0.f);
            {
                const float2* zl = ws->z + 17 * lane;      // zpos(16*lane + K) = 17*lane + K + (K >> 4)
                float2 W[32];                              // sliding sample window, slot = K & 31
#pragma unroll
                for (int K = 128; K <= 142; ++K) W[K & 31] = zl[K + (K >> 4)];
#pragma unroll 1
                for (int s4 = 0; s4 < 4; ++s4) {
                    fir_step<0>(W, acc, zl, ws->taps2, s4);
                    fir_step<1>(W, acc, zl, ws->taps2, s4);
                    fir_step<2>(W, acc, zl, ws->taps2, s4);
                    fir_step<3>(W, acc, zl, ws->taps2, s4);
                }
            }
            __syncwarp();
            // ---- slide the history: last 126 mixed samples move to the front -------------------
            {
                float2 tmp[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) { int k = lane + 32 * i; tmp[i] = (k < H) ? ws->z[zpos(FR + 1 + k)] : make_float2(0.f, 0.f); }
                __syncwarp();
#pragma unroll
                for (int i = 0; i < 4; ++i) { int k = lane + 32 * i; if (k < H) ws->z[zpos(1 + k)] = tmp[i]; }
            }
            // ---- magnitude, RSSI -----------------------------------------------------------------
            float mag[SPL];
            float psum = 0.f, bmax = 0.f;
#pragma unroll
            for (int r = 0; r < SPL; ++r) {
                float p = acc[r].x * acc[r].x + acc[r].y * acc[r].y;
                psum += p;
                mag[r] = sqrt_approx(p);
                bmax = fmaxf(bmax, mag[r]);
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                psum += __shfl_xor_sync(0xffffffffu, psum, o);
                bmax = fmaxf(bmax, __shfl_xor_sync(0xffffffffu, bmax, o));
            }
            if (kp.rssi && lane == 0) {
                float mp = fmaxf(psum * (1.0f / FR), 1e-30f);
                kp.rssi[(size_t)ch * (kp.pitch / FR) + b] = 10.0f * log10f(mp * (1.0f / (SSDR_FS * SSDR_FS))) + kDemodFsDbm;
            }
            // ---- detector --------------------------------------------------------------------------
            float a[SPL];
            if (cp.mode == SSDR_MODE_NBFM) {
                float2 last = acc[SPL - 1];
                float2 prv = make_float2(__shfl_up_sync(0xffffffffu, last.x, 1), __shfl_up_sync(0xffffffffu, last.y, 1));
                if (lane == 0) prv = zprev;
#pragma unroll
                for (int r = 0; r < SPL; ++r) {
                    float2 z = acc[r];
                    float re = z.x * prv.x + z.y * prv.y;     // z * conj(prev)
                    float im = z.y * prv.x - z.x * prv.y;
                    // a zero product (first sample of a stream, or silence) demodulates to 0, not +-pi
                    a[r] = (re == 0.0f && im == 0.0f) ? 0.0f : atan2f(im, re) * (32767.0f / 3.14159265358979f);
                    prv = z;
                }
            } else if (cp.mode == SSDR_MODE_AM) {
                // carrier tracker dc[k] = dc[k-1] + beta (mag[k] - dc[k-1]) in float64: lane-local
                // recurrence from a zero (lane 0: true) carry-in, then an affine warp scan.
                double B = (lane == 0) ? dc : 0.0;
#pragma unroll
                for (int r = 0; r < SPL; ++r) B = B + kDemodAmBeta * ((double)mag[r] - B);
#pragma unroll
                for (int s = 0; s < 5; ++s) {
                    double up = __shfl_up_sync(0xffffffffu, B, 1 << s);
                    if (lane >= (1 << s)) B = B + kp.am_pow16[s] * up;   // (om^16)^(2^s)
                }
                double carry = __shfl_up_sync(0xffffffffu, B, 1);
                if (lane == 0) carry = dc;
                dc = __shfl_sync(0xffffffffu, B, 31);
                // B after the scan is the carrier at the end of this lane's segment; replay with the carry-in
                double d = carry;
#pragma unroll
                for (int r = 0; r < SPL; ++r) {
                    d = d + kDemodAmBeta * ((double)mag[r] - d);
                    a[r] = (float)((double)mag[r] - d);
                }
            } else {
#pragma unroll
                for (int r = 0; r < SPL; ++r) {
                    const int k = SPL * lane + r;
                    float c, s;
                    nco(ph2 + (unsigned)k * cp.inc2, c, s);
                    a[r] = acc[r].x * c - acc[r].y * s;       // Re(z * exp(+j theta2))
                }
            }
            // ---- AGC ---------------------------------------------------------------------------------
            float out[SPL];
            if (cp.mode == SSDR_MODE_NBFM) {
#pragma unroll
                for (int r = 0; r < SPL; ++r) out[r] = a[r];
            } else if (!cp.agc_on) {
#pragma unroll
                for (int r = 0; r < SPL; ++r) out[r] = a[r] * cp.man_gain;
            } else {
                // hang: hm[k] = max(max(ring), prefix max of mag)
                float hb = ring;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) hb = fmaxf(hb, __shfl_xor_sync(0xffffffffu, hb, o));
                float m[SPL];
                float run = 0.f;
#pragma unroll
                for (int r = 0; r < SPL; ++r) { run = fmaxf(run, mag[r]); m[r] = cp.agc_hang ? run : mag[r]; }
                if (cp.agc_hang) {
                    float excl = run;                          // inclusive scan of lane maxima
#pragma unroll
                    for (int s = 0; s < 5; ++s) {
                        float up = __shfl_up_sync(0xffffffffu, excl, 1 << s);
                        if (lane >= (1 << s)) excl = fmaxf(excl, up);
                    }
                    excl = __shfl_up_sync(0xffffffffu, excl, 1);
                    if (lane == 0) excl = 0.f;
                    excl = fmaxf(excl, hb);
#pragma unroll
                    for (int r = 0; r < SPL; ++r) m[r] = fmaxf(m[r], excl);
                }
                // u[k] = hm[k] 2^(k c2); M = prefix max with seed e_in 2^(-c2); e[k] = M[k] 2^(-k c2).  k = 16 lane + r:
                // 2^(+-k c2) = 2^(+-16 lane c2) * (2^(+-c2))^r, the second factor by a running product (16 steps)
                const float up1 = ex2_approx(cp.c2), dn1 = ex2_approx(-cp.c2);
                float upk = ex2_approx((float)(SPL * lane) * cp.c2), dnk = ex2_approx(-(float)(SPL * lane) * cp.c2);
                float mrun = 0.f;
                float u[SPL];
                float dn[SPL];
#pragma unroll
                for (int r = 0; r < SPL; ++r) {
                    u[r] = m[r] * upk;
                    dn[r] = dnk;
                    upk *= up1; dnk *= dn1;
                    mrun = fmaxf(mrun, u[r]);
                    u[r] = mrun;
                }
                float pre = mrun;
#pragma unroll
                for (int s = 0; s < 5; ++s) {
                    float up = __shfl_up_sync(0xffffffffu, pre, 1 << s);
                    if (lane >= (1 << s)) pre = fmaxf(pre, up);
                }
                pre = __shfl_up_sync(0xffffffffu, pre, 1);
                if (lane == 0) pre = 0.f;
                pre = fmaxf(pre, e_in * dn1);
                float e_last = 0.f;
#pragma unroll
                for (int r = 0; r < SPL; ++r) {
                    float e = fmaxf(u[r], pre) * dn[r];
                    float m2 = lg2_approx(e * (1.0f / SSDR_FS));
                    float g = kDemodAgcOut * ex2_approx(fmaxf(m2, cp.knee2) * cp.slope_m1);
                    out[r] = a[r] * g;
                    e_last = e;
                }
                e_in = __shfl_sync(0xffffffffu, e_last, 31);
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
            zprev = make_float2(__shfl_sync(0xffffffffu, acc[SPL - 1].x, 31), __shfl_sync(0xffffffffu, acc[SPL - 1].y, 31));
            if (lane == (int)(blk % SSDR_HANG_BLOCKS)) ring = bmax;
            blk++;
            ph1 += (unsigned)FR * cp.inc1;
            ph2 += (unsigned)FR * cp.inc2;
            __syncwarp();
        }
        // ---- store per-channel state ---------------------------------------------------------------
        for (int i = lane; i < H; i += 32) kp.hist[(size_t)ch * H + i] = ws->z[zpos(1 + i)];
        if (lane < SSDR_HANG_BLOCKS) stp->ring[lane] = ring;
        if (lane == 0) {
            stp->ph1 = ph1; stp->ph2 = ph2; stp->e_in = e_in; stp->dc = dc;
            stp->zprev_re = zprev.x; stp->zprev_im = zprev.y; stp->blk = blk;
        }
        __syncwarp();
    }
}

}  // namespace

int demod_launch(const DemodLaunch& a, cudaStream_t st) {
    DemodKernelParams kp;
    kp.iq = a.iq; kp.chan = a.chan; kp.state = a.state; kp.hist = a.hist; kp.taps = a.taps;
    kp.pcm_f32 = a.pcm_f32; kp.pcm_i16 = a.pcm_i16; kp.rssi = a.rssi;
    kp.batch = a.batch; kp.n_samples = a.n_samples; kp.pitch = a.pitch ? a.pitch : a.n_samples;
    for (int s = 0; s < 5; ++s) kp.am_pow16[s] = a.am_pow16[s];
    const size_t smem = sizeof(WarpSmem) * WARPS;
    auto kern = (a.iq_format == SSDR_IQ_CF32) ? demod_kernel<SSDR_IQ_CF32> : demod_kernel<SSDR_IQ_S16BE>;
    SSDR_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int occ = 0;
    SSDR_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, WARPS * 32, smem));
    if (occ < 1) occ = 1;
    int grid = sm_count() * occ;
    const int need = (a.batch + WARPS - 1) / WARPS;
    if (grid > need) grid = need;
    kern<<<grid, WARPS * 32, smem, st>>>(kp);
    count_launch();
    SSDR_CUDA(cudaGetLastError());
    return SSDR_OK;
}

}  // namespace ssdr
