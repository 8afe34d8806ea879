// Waterfall hot path: IQ frames -> Hann window -> mixed-radix DIF FFT in shared memory -> |X|^2 ->
// Kiwi byte (dBm + 255) -> time-binning accumulate over n_avg frames -> dB cal + 40th-percentile
// auto-scale + colour row + uint8 pixels, all in ONE kernel (K1+K2+K3 of SURVEY.md section 2).
//
// Replaces (per channel) the remote KiwiSDR W/F computation whose uint8 lines the reference receives
// (utils_supersdr.py:780-785), kiwi_waterfall.run's averaging (:881-886) and
// kiwi_waterfall.spectrum_db2col (:787-813).  Arithmetic spec: DESIGN.md section 4; bit-exact CPU
// statement: oracle/c/ssdr_oracle.c.
//
// Layout: one frame group of G = N/EPT threads owns one channel at a time and walks its n_avg
// frames; FPC groups share a CTA for small N.  The frame lives in shared memory (XOR-swizzled
// float2[N]); every pass is in place, so one barrier per pass; the per-bin byte sums live in
// shared memory as uint16 in FFT *position* order (digit-reversed), and are permuted to bin order
// only once per channel when the row is written.
#include <cuda_runtime.h>

#include <cstdint>
#include <cstring>
#include <vector>

#include "common.cuh"
#include "fft_radix.cuh"
#include "wf_host.h"

namespace ssdr {

// ---------------------------------------------------------------------------------------------
// compile-time plan (same rule as oracle/c/ssdr_oracle.c so_fft_plan)
// ---------------------------------------------------------------------------------------------
struct PlanC {
    int np;
    int r[5];
};
constexpr PlanC make_plan(int lg) {
    PlanC p{0, {0, 0, 0, 0, 0}};
    int rem = lg;
    while (rem >= 6 || rem == 4) { p.r[p.np++] = 16; rem -= 4; }
    if (rem == 5) { p.r[p.np++] = 8; p.r[p.np++] = 4; }
    else if (rem == 3) p.r[p.np++] = 8;
    else if (rem == 2) p.r[p.np++] = 4;
    return p;
}
constexpr int kTablePassMax = 1024;   // a pass uses exact table twiddles iff (L/R)*(R-1) <= this

constexpr int ilog2(int v) { int l = 0; while ((1 << l) < v) ++l; return l; }

template <int LG>
struct Cfg {
    static constexpr int N = 1 << LG;
    static constexpr PlanC plan = make_plan(LG);
    static constexpr int NP = plan.np;
    static constexpr int EPT = (LG >= 13) ? 32 : 16;      // elements per thread
    static constexpr int G = N / EPT;                     // threads per frame group
    static constexpr int THREADS = (G >= 256) ? G : 256;
    static constexpr int FPC = THREADS / G;               // frame groups per CTA
    static constexpr int R0 = plan.r[0];
    static constexpr int RL = plan.r[plan.np - 1];        // radix of the last pass
    static constexpr int M0 = N / R0;
    static constexpr int NCHUNK = N / RL;                 // accumulator chunks (RL uint16 each)
    static constexpr int CSH = ilog2(M0 / RL);            // chunk >> CSH == first-pass digit q0
    static constexpr int CMASK = (1 << (CSH < 4 ? CSH : 4)) - 1;   // swizzle only bits below the q0 field
    // sub-transform length before pass p
    static constexpr int radix(int p) { return make_plan(LG).r[p]; }   // usable with a runtime index in device code
    static constexpr int Lof(int p) { int L = N; for (int i = 0; i < p; ++i) L /= radix(i); return L; }
    static constexpr bool table_pass(int p) { int L = Lof(p), R = radix(p); return (L / R) * (R - 1) <= kTablePassMax; }
    static constexpr int table_off(int p) {               // float2 offset of pass p's table
        int off = 0;
        for (int i = 0; i < p; ++i) { int L = Lof(i), R = radix(i); if (table_pass(i) && L / R > 1) off += (L / R) * (R - 1); }
        return off;
    }
    static constexpr int TABLE_ELEMS = table_off(plan.np);
    // dynamic shared memory layout (bytes)
    static constexpr size_t SM_DATA = 0;
    static constexpr size_t SM_ACC = SM_DATA + (size_t)FPC * N * sizeof(float2);
    static constexpr size_t SM_TW = SM_ACC + (size_t)FPC * N * sizeof(uint16_t);
    static constexpr size_t SM_THR = SM_TW + (size_t)TABLE_ELEMS * sizeof(float2);
    static constexpr size_t SM_RED = SM_THR + 260 * sizeof(float);
    static constexpr size_t SM_BYTES = SM_RED + (size_t)FPC * 8 * sizeof(int);
};

struct WfKernelParams {
    const void* iq;                 // [batch][n_avg][N] samples
    const float2* wtab;             // master twiddle table, N entries
    const float* thr;               // 257 thresholds (thr[256] = +inf)
    ssdr_wf_display_t* disp;        // [batch], low_clip_db/dynamic_range updated when auto_scale
    uint8_t* pixels;                // [batch][N] or null
    float* colour;                  // [batch][N] or null
    float* spectrum;                // [batch][N] or null
    ssdr_wf_scalars_t* scalars;     // [batch] or null
    const uint8_t* lines;           // colorrow entry: [batch][n_avg][N] uint8 lines (else null)
    int batch, n_avg;
    int p_lo;
    float p_gamma;
    float est_c1, est_c0;           // byte estimate = floor(log2(P) * c1 + c0)
};

SSDR_DEV int swz(int a) { return a ^ ((a >> 4) & 15); }

// ---- sample load (K6 fused): complex64 or Kiwi big-endian int16 pairs (kiwi/client.py:449-453) --
template <int FMT>
SSDR_DEV float2 load_iq(const void* base, size_t idx) {
    if constexpr (FMT == SSDR_IQ_CF32) {
        return __ldcs(reinterpret_cast<const float2*>(base) + idx);
    } else {
        unsigned v = __ldcs(reinterpret_cast<const unsigned*>(base) + idx);
        const unsigned sw = __byte_perm(v, 0u, 0x2301);   // swap the bytes of both 16-bit halves
        const int i = (int)(short)(sw & 0xffffu), q = (int)sw >> 16;
        return make_float2((float)i, (float)q);
    }
}

// ---- |X|^2 -> Kiwi byte: count thresholds (exact), starting from a log2 estimate -----------------
SSDR_DEV int quantise(float P, const float* thr, float c1, float c0) {
    float est = __fmaf_rn(__log2f(P), c1, c0);          // -inf for P == 0, NaN never (P >= 0)
    int k = (int)fminf(fmaxf(est, 0.0f), 255.0f);
    while (k < 255 && P >= thr[k + 1]) ++k;
    while (k > 0 && P < thr[k]) --k;
    return k;
}

// ---------------------------------------------------------------------------------------------
// one FFT pass (in place in shared memory; the first pass reads global memory)
// ---------------------------------------------------------------------------------------------
template <class C, int P, int FMT, bool WINDOW>
SSDR_DEV void fft_pass(float2* d, uint16_t* acc, const float2* tws, const float* thr, int t,
                       const void* src, size_t src_off, const WfKernelParams& kp, bool first_frame) {
    constexpr int N = C::N, R = C::plan.r[P], L = C::Lof(P), M = L / R, G = C::G;
    constexpr bool FIRST = (P == 0), LAST = (P == C::NP - 1);
    constexpr bool TABLE = C::table_pass(P);
    constexpr int NB = (N / R) / G;   // butterflies per thread
    static_assert(NB >= 1, "group too large for this radix");
    const float2* tw = tws + C::table_off(P);
#pragma unroll 1
    for (int i = 0; i < NB; ++i) {
        const int u = t + i * G;
        const int j = u & (M - 1);
        const int base = (u / M) * L;
        float2 x[R];
        float2 w[R];
        if constexpr (FIRST) {
#pragma unroll
            for (int m = 0; m < R; ++m) x[m] = load_iq<FMT>(src, src_off + (size_t)(j + m * M));
            if constexpr (WINDOW || !TABLE) w[1] = __ldg(kp.wtab + j);
            if constexpr (WINDOW) {
                const float c = w[1].x, dd = w[1].y;
#pragma unroll
                for (int m = 0; m < R; ++m) {
                    float Cm, Sm;
                    unit16(m * (16 / R), Cm, Sm);
                    float tt = dd * Sm;
                    float cm = __fmaf_rn(c, Cm, tt);
                    float wv = __fmaf_rn(-0.5f, cm, 0.5f);
                    x[m].x = x[m].x * wv;
                    x[m].y = x[m].y * wv;
                }
            }
        } else {
#pragma unroll
            for (int m = 0; m < R; ++m) x[m] = d[swz(base + j + m * M)];
        }
        dft<R>(x);
        if constexpr (M > 1) {
            if constexpr (TABLE) {
#pragma unroll
                for (int q = 1; q < R; ++q) w[q] = tw[(q - 1) * M + j];
            } else {
                if constexpr (!FIRST) w[1] = __ldg(kp.wtab + j * (N / L));
                tw_chain<R>(w);
            }
#pragma unroll
            for (int q = 1; q < R; ++q) x[q] = cmul(x[q], w[q]);
        }
        if constexpr (!LAST) {
#pragma unroll
            for (int q = 0; q < R; ++q) d[swz(base + j + q * M)] = x[q];
        } else {
            // epilogue: positions u*R + q.  Power -> byte -> accumulate uint16 sums (position order).
            unsigned pk[R / 2];
#pragma unroll
            for (int q = 0; q < R; q += 2) {
                float t0 = x[q].y * x[q].y;
                float P0 = __fmaf_rn(x[q].x, x[q].x, t0);
                float t1 = x[q + 1].y * x[q + 1].y;
                float P1 = __fmaf_rn(x[q + 1].x, x[q + 1].x, t1);
                unsigned b0 = (unsigned)quantise(P0, thr, kp.est_c1, kp.est_c0);
                unsigned b1 = (unsigned)quantise(P1, thr, kp.est_c1, kp.est_c0);
                pk[q / 2] = b0 | (b1 << 16);
            }
            const int cs = u ^ ((u >> C::CSH) & C::CMASK);
            unsigned* a32 = reinterpret_cast<unsigned*>(acc + (size_t)cs * R);
            if (!first_frame) {
#pragma unroll
                for (int q = 0; q < R / 2; ++q) pk[q] += a32[q];   // two uint16 lanes, no carry (<= 25500)
            }
#pragma unroll
            for (int q = 0; q < R / 2; ++q) a32[q] = pk[q];
        }
    }
}

template <class C, int P, int FMT, bool WINDOW>
SSDR_DEV void fft_all_passes(float2* d, uint16_t* acc, const float2* tws, const float* thr, int t,
                             const void* src, size_t src_off, const WfKernelParams& kp, bool first_frame,
                             bool active) {
    if constexpr (P < C::NP) {
        if (P == 0) __syncthreads();          // previous frame's last pass has finished reading d
        if (active) fft_pass<C, P, FMT, WINDOW>(d, acc, tws, thr, t, src, src_off, kp, first_frame);
        if (P < C::NP - 1) __syncthreads();
        fft_all_passes<C, P + 1, FMT, WINDOW>(d, acc, tws, thr, t, src, src_off, kp, first_frame, active);
    }
}

// ---------------------------------------------------------------------------------------------
// group reductions (a frame group is G threads: a half warp, a warp, or several warps)
// ---------------------------------------------------------------------------------------------
template <int G>
SSDR_DEV unsigned group_mask(int lane) {
    if constexpr (G >= 32) return 0xffffffffu;
    else return ((1u << G) - 1u) << (lane & ~(G - 1));
}

// Colour stage for one channel, executed by its frame group.  `red` = 8 ints of scratch per group.
// keys live in acc (uint16 sums) in chunk order; CHUNK_ID maps the k-order index idx to the chunk.
template <class C, bool FFT_ORDER>
SSDR_DEV void colour_stage(uint16_t* acc, int* red, int t, int ch, bool active, const WfKernelParams& kp) {
    constexpr int N = C::N, G = C::G, RL = C::RL, NCH = C::NCHUNK;
    constexpr int CPT = NCH / G;      // chunks per thread
    const int lane = threadIdx.x & 31;
    const unsigned gmask = group_mask<G>(lane);
    const bool leader = (G >= 32) ? (lane == 0) : ((lane & (G - 1)) == 0);

    // chunk (position order, swizzled) holding bins k = idx + NCH*q, q = 0..RL-1
    auto chunk_of = [](int idx) -> int {
        if constexpr (!FFT_ORDER) {
            return idx;
        } else {
            int c = 0, rem = idx, Mi = N;
#pragma unroll
            for (int p = 0; p < C::NP - 1; ++p) {
                const int R = C::radix(p);
                Mi /= R;
                int q = rem & (R - 1);
                rem >>= ilog2(R);
                c += q * (Mi / RL);
            }
            return c ^ ((c >> C::CSH) & C::CMASK);
        }
    };

    ssdr_wf_display_t dp;
    if (active) dp = kp.disp[ch];
    else { dp.zoom = 0; dp.auto_scale = 0; dp.delta_low_db = 0; dp.delta_high_db = 0; dp.low_clip_db = 0.f; dp.dynamic_range = 40.f; }

    // wf_db[0] = wf_db[1] (utils_supersdr.py:791): output bin o = k ^ N/2 (FFT order) or o = k.
    __syncthreads();
    if (active && t == 0) {
        // red[4] keeps the raw sum of bin 0: kiwi_waterfall.spectrum itself is not patched
        if constexpr (FFT_ORDER) {
            // o = 0 <-> k = N/2: idx 0, q = RL/2;  o = 1 <-> k = N/2 + 1: idx 1, q = RL/2
            red[4] = acc[(size_t)chunk_of(0) * RL + RL / 2];
            acc[(size_t)chunk_of(0) * RL + RL / 2] = acc[(size_t)chunk_of(1) * RL + RL / 2];
        } else {
            red[4] = acc[0];
            acc[0] = acc[1];   // plain order: chunk 0 elements 0 and 1 (RL >= 2)
        }
    }
    if (leader && active) { red[0] = 0; red[1] = 0; red[2] = 0x7fffffff; red[3] = 0; }
    __syncthreads();

    // keys of this thread: CPT chunks x RL
    unsigned keys[CPT * RL / 2];
    int kmax = 0;
#pragma unroll
    for (int i = 0; i < CPT; ++i) {
        const int idx = t + i * G;
        const unsigned* a32 = reinterpret_cast<const unsigned*>(acc + (size_t)chunk_of(idx) * RL);
#pragma unroll
        for (int q = 0; q < RL / 2; ++q) {
            unsigned v = active ? a32[q] : 0u;
            keys[i * (RL / 2) + q] = v;
            kmax = max(kmax, (int)max(v & 0xffffu, v >> 16));
        }
    }
    float low_clip = dp.low_clip_db, high_clip = 0.f, dyn = dp.dynamic_range;
    const float fn = (float)kp.n_avg, z3 = (float)(3 * dp.zoom);
    auto wfdb = [&](float s) { return ((__fdiv_rn(s, fn) - 255.0f) - 13.0f) + z3; };

    // ---- max and the two order statistics (ranks p_lo, p_lo + 1) by counting bisection ----------
    kmax = __reduce_max_sync(gmask, kmax);
    if (leader && active) atomicMax(&red[3], kmax);
    __syncthreads();
    const int vmax = red[3];
    int lo = 0, hi = vmax;
    const int want = kp.p_lo + 1;                 // smallest v with count(keys <= v) >= want
    int cnt_lo = 0;
    // fixed trip count (uniform across groups sharing the CTA): 15 bits cover 255 * 100
#pragma unroll 1
    for (int it = 0; it < 15; ++it) {
        const int mid = (lo + hi) >> 1;
        int c = 0;
#pragma unroll
        for (int i = 0; i < CPT * RL / 2; ++i) {
            c += ((int)(keys[i] & 0xffffu) <= mid) + ((int)(keys[i] >> 16) <= mid);
        }
        c = __reduce_add_sync(gmask, c);
        int* slot = &red[it & 1];
        if (leader && active) atomicAdd(slot, c);
        __syncthreads();
        const int total = *slot;
        if (total >= want) hi = mid; else lo = mid + 1;
        if (leader && active) red[(it + 1) & 1] = 0;   // the other slot is free after this barrier
        __syncthreads();
    }
    const int v_lo = hi;                           // == lo
    {   // count(keys <= v_lo) and min{key > v_lo}
        int c = 0, mn = 0x7fffffff;
#pragma unroll
        for (int i = 0; i < CPT * RL / 2; ++i) {
            int a = (int)(keys[i] & 0xffffu), b = (int)(keys[i] >> 16);
            c += (a <= v_lo) + (b <= v_lo);
            if (a > v_lo) mn = min(mn, a);
            if (b > v_lo) mn = min(mn, b);
        }
        c = __reduce_add_sync(gmask, c);
        mn = __reduce_min_sync(gmask, mn);
        // both slots were zeroed: red[(15)&1] by the last iteration, and red[(14)&1]... reset here
        __syncthreads();
        if (leader && active) { red[0] = 0; }
        __syncthreads();
        if (leader && active) { atomicAdd(&red[0], c); atomicMin(&red[2], mn); }
        __syncthreads();
        cnt_lo = red[0];
        const int v_hi = (cnt_lo >= want + 1 || red[2] == 0x7fffffff) ? v_lo : red[2];
        if (dp.auto_scale) {
            // numpy _lerp in float32
            const float a = wfdb((float)v_lo), b = wfdb((float)v_hi), g = kp.p_gamma;
            const float dba = b - a;
            float p;
            if (g >= 0.5f) { float tt = 1.0f - g; tt = dba * tt; p = b - tt; }
            else { float tt = dba * g; p = a + tt; }
            low_clip = p;
            high_clip = wfdb((float)vmax);
            const float dd = high_clip - low_clip;
            dyn = dd > 40.0f ? dd : 40.0f;
        }
    }
    const float low = low_clip + (float)dp.delta_low_db;
    const float nf = dyn + (float)dp.delta_high_db;
    const float den = nf - (float)dp.delta_low_db;
    if (active && t == 0) {
        if (dp.auto_scale) { kp.disp[ch].low_clip_db = low_clip; kp.disp[ch].dynamic_range = dyn; }
        if (kp.scalars) {
            ssdr_wf_scalars_t s;
            s.low_clip_db = low_clip; s.high_clip_db = dp.auto_scale ? high_clip : wfdb((float)vmax);
            s.dynamic_range = dyn; s.wf_min_db = low - z3; s.wf_max_db = (low_clip + nf) - z3;
            kp.scalars[ch] = s;
        }
    }
    // ---- colour row: permute to bin order, coalesced byte / float stores ------------------------
    if (active) {
        const size_t row = (size_t)ch * N;
#pragma unroll
        for (int i = 0; i < CPT; ++i) {
            const int idx = t + i * G;
#pragma unroll
            for (int q = 0; q < RL; ++q) {
                const unsigned pair = keys[i * (RL / 2) + (q >> 1)];
                const float s = (float)((q & 1) ? (pair >> 16) : (pair & 0xffffu));
                const int k = FFT_ORDER ? (idx + NCH * q) : (idx * RL + q);
                const int o = FFT_ORDER ? (k ^ (N / 2)) : k;
                const float m = __fdiv_rn(s, fn);
                const float w = ((m - 255.0f) - 13.0f) + z3;
                float c = __fdiv_rn(w - low, den);
                c = fminf(fmaxf(c, 0.0f), 1.0f);
                c = c * 254.0f;
                c = fminf(fmaxf(c, 0.0f), 255.0f);
                if (kp.spectrum) kp.spectrum[row + o] = (o == 0) ? __fdiv_rn((float)red[4], fn) : m;
                if (kp.colour) kp.colour[row + o] = c;
                if (kp.pixels) kp.pixels[row + o] = (uint8_t)__float2int_rn(c);
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// the fused waterfall kernel
// ---------------------------------------------------------------------------------------------
template <int LG, int FMT, bool WINDOW>
__global__ void __launch_bounds__(Cfg<LG>::THREADS, (LG >= 13) ? 1 : 2)
wf_fft_kernel(const WfKernelParams kp) {
    using C = Cfg<LG>;
    constexpr int N = C::N, G = C::G, FPC = C::FPC;
    extern __shared__ __align__(16) unsigned char smem[];
    float2* data = reinterpret_cast<float2*>(smem + C::SM_DATA);
    uint16_t* accs = reinterpret_cast<uint16_t*>(smem + C::SM_ACC);
    float2* tws = reinterpret_cast<float2*>(smem + C::SM_TW);
    float* thr = reinterpret_cast<float*>(smem + C::SM_THR);
    int* reds = reinterpret_cast<int*>(smem + C::SM_RED);

    const int slot = threadIdx.x / G, t = threadIdx.x % G;
    float2* d = data + (size_t)slot * N;
    uint16_t* acc = accs + (size_t)slot * N;
    int* red = reds + slot * 8;

    // one-time tables: thresholds and the exact-table twiddles of the small passes
    for (int i = threadIdx.x; i < 257; i += blockDim.x) thr[i] = kp.thr[i];
    {
        int L = N;
#pragma unroll
        for (int p = 0; p < C::NP; ++p) {
            const int R = C::radix(p), M = L / R;
            if (C::table_pass(p) && M > 1) {
                float2* tw = tws + C::table_off(p);
                for (int e = threadIdx.x; e < M * (R - 1); e += blockDim.x) {
                    int q = e / M + 1, j = e - (q - 1) * M;
                    tw[e] = kp.wtab[j * q * (N / L)];
                }
            }
            L = M;
        }
    }
    __syncthreads();

    const int n_groups = (kp.batch + FPC - 1) / FPC;
    for (int g = blockIdx.x; g < n_groups; g += gridDim.x) {
        const int ch = g * FPC + slot;
        const bool active = ch < kp.batch;
        for (int f = 0; f < kp.n_avg; ++f) {
            const size_t off = ((size_t)ch * kp.n_avg + f) * N;
            fft_all_passes<C, 0, FMT, WINDOW>(d, acc, tws, thr, t, kp.iq, off, kp, f == 0, active);
        }
        colour_stage<C, true>(acc, red, t, ch, active, kp);
    }
}

// Tier-P entry: finished uint8 lines in, same colour stage (no FFT).  utils_supersdr.py:783-813,881-886
template <int LG>
__global__ void __launch_bounds__(Cfg<LG>::THREADS)
wf_colorrow_kernel(const WfKernelParams kp) {
    using C = Cfg<LG>;
    constexpr int N = C::N, G = C::G, FPC = C::FPC;
    extern __shared__ __align__(16) unsigned char smem[];
    uint16_t* accs = reinterpret_cast<uint16_t*>(smem);
    int* reds = reinterpret_cast<int*>(smem + (size_t)FPC * N * sizeof(uint16_t));
    const int slot = threadIdx.x / G, t = threadIdx.x % G;
    uint16_t* acc = accs + (size_t)slot * N;
    int* red = reds + slot * 8;
    const int n_groups = (kp.batch + FPC - 1) / FPC;
    for (int g = blockIdx.x; g < n_groups; g += gridDim.x) {
        const int ch = g * FPC + slot;
        const bool active = ch < kp.batch;
        __syncthreads();
        if (active) {
            // each thread sums 4 adjacent bins per step: coalesced 32-bit loads of the byte lines
            for (int i = t * 4; i < N; i += G * 4) {
                unsigned s0 = 0, s1 = 0, s2 = 0, s3 = 0;
                for (int f = 0; f < kp.n_avg; ++f) {
                    unsigned v = __ldcs(reinterpret_cast<const unsigned*>(kp.lines + ((size_t)ch * kp.n_avg + f) * N + i));
                    s0 += v & 0xff; s1 += (v >> 8) & 0xff; s2 += (v >> 16) & 0xff; s3 += v >> 24;
                }
                unsigned* a32 = reinterpret_cast<unsigned*>(acc + i);
                a32[0] = s0 | (s1 << 16);
                a32[1] = s2 | (s3 << 16);
            }
        }
        colour_stage<C, false>(acc, red, t, ch, active, kp);
    }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
template <int LG>
static int launch_fft(const WfKernelParams& kp, int fmt, int window, cudaStream_t st) {
    using C = Cfg<LG>;
    auto launch = [&](auto kern) -> int {
        SSDR_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SM_BYTES));
        int occ = 0;
        SSDR_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, C::THREADS, C::SM_BYTES));
        if (occ < 1) { set_error("waterfall kernel does not fit (smem %zu)", (size_t)C::SM_BYTES); return SSDR_E_CUDA; }
        const int n_groups = (kp.batch + C::FPC - 1) / C::FPC;
        int grid = sm_count() * occ;
        if (grid > n_groups) grid = n_groups;
        kern<<<grid, C::THREADS, C::SM_BYTES, st>>>(kp);
        count_launch();
        SSDR_CUDA(cudaGetLastError());
        return SSDR_OK;
    };
    if (fmt == SSDR_IQ_CF32) return window ? launch(wf_fft_kernel<LG, SSDR_IQ_CF32, true>) : launch(wf_fft_kernel<LG, SSDR_IQ_CF32, false>);
    return window ? launch(wf_fft_kernel<LG, SSDR_IQ_S16BE, true>) : launch(wf_fft_kernel<LG, SSDR_IQ_S16BE, false>);
}

template <int LG>
static int launch_colorrow(const WfKernelParams& kp, cudaStream_t st) {
    using C = Cfg<LG>;
    const size_t smem = (size_t)C::FPC * C::N * sizeof(uint16_t) + (size_t)C::FPC * 8 * sizeof(int);
    SSDR_CUDA(cudaFuncSetAttribute(wf_colorrow_kernel<LG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int n_groups = (kp.batch + C::FPC - 1) / C::FPC;
    int grid = sm_count() * 4;
    if (grid > n_groups) grid = n_groups;
    wf_colorrow_kernel<LG><<<grid, C::THREADS, smem, st>>>(kp);
    count_launch();
    SSDR_CUDA(cudaGetLastError());
    return SSDR_OK;
}

int wf_plan(int nfft, int* radices) {
    int lg = ilog2(nfft);
    if ((1 << lg) != nfft || lg < 8 || lg > 14) return -1;
    PlanC p = make_plan(lg);
    for (int i = 0; i < p.np; ++i) radices[i] = p.r[i];
    return p.np;
}

int wf_launch(const WfLaunch& a, cudaStream_t st) {
    WfKernelParams kp;
    std::memset(&kp, 0, sizeof(kp));
    kp.iq = a.iq; kp.wtab = reinterpret_cast<const float2*>(a.wtab); kp.thr = a.thr; kp.disp = a.disp;
    kp.pixels = a.pixels; kp.colour = a.colour; kp.spectrum = a.spectrum; kp.scalars = a.scalars;
    kp.lines = a.lines; kp.batch = a.batch; kp.n_avg = a.n_avg; kp.p_lo = a.p_lo; kp.p_gamma = a.p_gamma;
    kp.est_c1 = a.est_c1; kp.est_c0 = a.est_c0;
    const int lg = ilog2(a.nfft);
    if (a.lines) {
        switch (lg) {
            case 8: return launch_colorrow<8>(kp, st);
            case 9: return launch_colorrow<9>(kp, st);
            case 10: return launch_colorrow<10>(kp, st);
            case 11: return launch_colorrow<11>(kp, st);
            case 12: return launch_colorrow<12>(kp, st);
            case 13: return launch_colorrow<13>(kp, st);
            case 14: return launch_colorrow<14>(kp, st);
        }
    } else {
        switch (lg) {
            case 8: return launch_fft<8>(kp, a.iq_format, a.window, st);
            case 9: return launch_fft<9>(kp, a.iq_format, a.window, st);
            case 10: return launch_fft<10>(kp, a.iq_format, a.window, st);
            case 11: return launch_fft<11>(kp, a.iq_format, a.window, st);
            case 12: return launch_fft<12>(kp, a.iq_format, a.window, st);
            case 13: return launch_fft<13>(kp, a.iq_format, a.window, st);
            case 14: return launch_fft<14>(kp, a.iq_format, a.window, st);
        }
    }
    set_error("unsupported nfft %d", a.nfft);
    return SSDR_E_ARG;
}

}  // namespace ssdr
