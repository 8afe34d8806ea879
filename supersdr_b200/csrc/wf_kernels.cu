// Waterfall hot path: IQ frames -> Hann window -> mixed-radix DIF FFT in shared memory -> |X|^2 ->
// Kiwi byte (dBm + 255) -> time-binning accumulate over n_avg frames -> dB cal + 40th-percentile
// auto-scale + colour row + uint8 pixels, all in ONE kernel (K1+K2+K3 of SURVEY.md section 2).
//
// Replaces (per channel) the remote KiwiSDR W/F computation whose uint8 lines the reference receives
// (utils_supersdr.py:780-785), kiwi_waterfall.run's averaging (:881-886) and
// kiwi_waterfall.spectrum_db2col (:787-813).  Arithmetic spec: DESIGN.md section 4; bit-exact CPU
// statement: oracle/c/ssdr_oracle.c.
//
// Layout: a frame group of G = N/32 threads owns one channel at a time and
// walks its n_avg frames; every thread handles 32 points of every pass.  The plan is one first pass
// of radix N/32^k (reads HBM with coalesced 8-byte loads, applies the window, writes shared memory)
// followed by k <= 2 radix-32 passes, so a 16384-point frame makes only TWO round trips through
// shared memory.  The frame lives in shared memory as float2[N + N/32] (one pad element per 32:
// every pass reads and writes with immediate offsets from one per-thread base and is bank-conflict
// free).  The per-bin byte sums over the n_avg frames never leave registers (16 packed uint16 pairs
// per thread); the colour stage selects the order statistics from those registers and transposes the
// finished row through the (then idle) frame buffer so that HBM sees only coalesced row stores.
#include <cuda.h>                // CUtensorMap (types only: the encoder is fetched through cudaGetDriverEntryPoint, no libcuda link)
#include <cuda_runtime.h>

#include <cstdint>
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <type_traits>
#include <vector>

#include "common.cuh"
#include "fft_radix.cuh"
#include "tmem_scratch.cuh"
#include "wf_host.h"

// Developer timing experiments (scripts/exp_variants.sh): SSDR_EXP is a bit mask that removes one phase of the
// kernel at a time to measure its marginal cost.  Results are WRONG when any bit is set; the product build has 0.
#ifndef SSDR_EXP
#define SSDR_EXP 0
#endif

// Developer timeline (scripts/wf_trace.py): -DSSDR_TRACE records clock64() at the phase boundaries of every warp of CTA 0
// for the first frames of a launch.  Not in the product build.
#ifdef SSDR_TRACE
#define TRACE_FRAMES 40
#define TRACE_PTS 24
__device__ long long g_wf_trace[TRACE_FRAMES * 16 * TRACE_PTS];
__device__ int g_wf_trace_frames;
extern "C" int ssdr_debug_wf_trace(long long* out, int* frames) {
    cudaMemcpyFromSymbol(out, g_wf_trace, sizeof(long long) * TRACE_FRAMES * 16 * TRACE_PTS);
    cudaMemcpyFromSymbol(frames, g_wf_trace_frames, sizeof(int));
    return 0;
}
// time stamps go to shared memory (a global store in the middle of the frame would sit in front of the proxy fence of the
// staging hook and distort what it measures); the warp's row is flushed to global memory at the end of the frame
__device__ __forceinline__ long long* trace_slots() { __shared__ long long s[16 * TRACE_PTS]; return s; }
#define TRACE(pt) do { if ((threadIdx.x & 31) == 0) trace_slots()[(threadIdx.x >> 5) * TRACE_PTS + (pt)] = clock64(); } while (0)
#define TRACE_FLUSH() do { if (blockIdx.x == 0 && (threadIdx.x & 31) == 0 && tr_frame < TRACE_FRAMES) { for (int k_ = 0; k_ < TRACE_PTS; ++k_) \
    g_wf_trace[(tr_frame * 16 + (threadIdx.x >> 5)) * TRACE_PTS + k_] = trace_slots()[(threadIdx.x >> 5) * TRACE_PTS + k_]; } } while (0)
#else
#define TRACE(pt) do { } while (0)
#endif

namespace ssdr {

// ---------------------------------------------------------------------------------------------
// compile-time plan (same rule as oracle/c/ssdr_oracle.c so_fft_plan)
// ---------------------------------------------------------------------------------------------
struct PlanC {
    int np;
    int r[4];
};
constexpr PlanC make_plan(int lg) {
    PlanC p{0, {0, 0, 0, 0}};
    int k = lg / 5, r = lg - 5 * k;
    if (r == 0) { r = 5; k -= 1; }
    p.r[p.np++] = 1 << r;
    for (int i = 0; i < k; ++i) p.r[p.np++] = 32;
    return p;
}

constexpr int ilog2(int v) { int l = 0; while ((1 << l) < v) ++l; return l; }

template <int LG>
struct Cfg {
    static constexpr int N = 1 << LG;
    static constexpr PlanC plan = make_plan(LG);
    static constexpr int NP = plan.np;                    // 2 (N <= 1024) or 3
    static constexpr int R0 = plan.r[0];
    static constexpr int M0 = N / R0;                     // 32 (NP == 2) or 1024 (NP == 3)
    static constexpr int G = N / 32;                      // threads per frame group
    static constexpr int THREADS = (G >= 256) ? G : 256;
    static constexpr int FPC = THREADS / G;               // frame groups per CTA
    static constexpr int NB0 = 32 / R0;                   // first-pass butterflies per thread
    static constexpr int PADN = N + N / 16;               // padded frame, float2 elements (two pad elements per 32)
    static constexpr int TW0 = (M0 == 32) ? 32 * (R0 - 1) : 0;   // first-pass twiddle table (NP == 2)
    static constexpr int TW1 = (NP == 3) ? 32 * 31 : 0;          // middle-pass twiddle table
    static constexpr int SWZ = (NP == 3) ? ilog2(R0) : -1;       // staging swizzle shift (colour stage)
    static constexpr int MIN_CTAS = (LG >= 14) ? 1 : 2;
#ifdef SSDR_NO_SPLIT      // developer check: plain group barriers, for compute-sanitizer racecheck (which does not model mbarriers)
    static constexpr bool SPLIT = false;
#else
    static constexpr bool SPLIT = (FPC == 1 && G > 32);     // split-phase frame-buffer hand-off (mbarrier)
#endif
    // Warps start their warp-local passes STAGGER cycles apart per level (4 levels): leaving the pass-1 barrier in lock
    // step, all warps of an SM hit the shared-memory pipe and then the fp32 pipe together; staggered, one warp's loads
    // overlap another's butterflies (measured: 16384: 2.26 -> 2.09 ms on config 2; 8192: +15 %; 2048/4096: +3 %;
    // smaller sizes lose).
    static constexpr int STAGGER = (LG == 14) ? 500 : (LG == 13) ? 300 : (LG >= 11) ? 100 : 0;     // cycles per level, measured optima
    static_assert(NP == 2 || NP == 3, "supported sizes: 64 .. 16384");
    static_assert(M0 == 32 || M0 == 1024, "first pass leaves 32 or 1024 sub-transforms");
    // dynamic shared memory layout (bytes)
    static constexpr size_t SM_DATA = 0;
    static constexpr size_t SM_TW0 = SM_DATA + (size_t)FPC * PADN * sizeof(float2);
    static constexpr size_t SM_TW1 = SM_TW0 + (size_t)TW0 * sizeof(float2);
    static constexpr size_t SM_ACC = SM_TW1 + (size_t)TW1 * sizeof(float2);        // byte sums, thread-private uint4[4][THREADS]
    // W_N^j of the first-pass butterflies (chain passes only): LG 14 parks this thread's two values in a thread-private
    // shared-memory column, LG 13 keeps its four in registers (the column would cost the second CTA per SM), smaller
    // sizes share one table of M0 = 1024 entries
    // LG 14 (round 2): no chain -- the 2 x 15 first-pass twiddles W_N^(j q) of a thread are loop invariant and live in
    // TENSOR MEMORY (tmem_scratch.cuh), read back with tcgen05.ld: 30 complex multiplies per thread and frame instead of
    // 58, no shared-memory column, and the loads read no registers
    static constexpr bool TW_DIRECT = (LG == 14);
    static constexpr bool TM_MID = TW_DIRECT;                                 // middle-pass twiddles in tensor memory as well
    static constexpr int TM_WORDS = TM_MID ? 128 : 64;                        // per thread: NB0 x 32 words (15 float2 + pad) [+ 64: 31 float2 + pad]
    static constexpr int TM_COLS = (THREADS / 128) * TM_WORDS;                // 4 warps share a lane quarter: 256 / 512 columns
    static constexpr int W1_MODE = (NP == 2) ? 0 : (LG == 14) ? 0 : (LG == 13) ? 2 : 3;
    static constexpr size_t W1_BYTES = (W1_MODE == 1) ? (size_t)THREADS * NB0 * sizeof(float2) : (W1_MODE == 3) ? (size_t)M0 * sizeof(float2) : 0;
    static constexpr size_t SM_W1 = SM_ACC + (size_t)THREADS * 16 * sizeof(unsigned);
    static constexpr size_t SM_WIN = SM_W1 + W1_BYTES;                                   // first half of the Hann window, N/2 floats
    static constexpr size_t SM_RED = SM_WIN + (size_t)(N / 2) * sizeof(float);
    static constexpr size_t SM_MBAR = SM_RED + (size_t)FPC * 8 * sizeof(int);
    static constexpr size_t SM_TBAR = SM_MBAR + 16;        // mbarrier + the tensor-memory base address slot
    static constexpr size_t SM_GBAR = SM_TBAR + 32 * 8;    // STAGED: one transaction barrier per warp (N >= 2048) / per frame group (N <= 1024)
    static constexpr size_t SM_BYTES = SM_GBAR + 8 * 8;    // STAGED, several multi-warp frame groups per CTA (N = 2048, 4096): one "consumed" barrier per group
    // STAGED (LG 14, local input): the raw samples of a frame are bulk-copied by the TMA engine into the frame buffer
    // itself -- each warp's own 1024-point region, free from the moment the warp has loaded its last-pass inputs -- one
    // frame ahead; no global load in the frame loop.  Column mapping: warp w owns first-pass columns 64 w .. 64 w + 63.
    // staged input: 8 KB tile per warp = its region (LG 13, 14) / the whole frame of a frame group (LG 9, 10).  Measured (B200, 65536
    // channels x 10 frames, fraction of the HBM bandwidth, staged / direct): 1024: 0.677 / 0.528, 512: 0.658 / 0.609, 256: 0.580 / 0.601
    // (four groups per warp: the tensor copy is a uniform-datapath instruction, ptxas serialises the four issuing lanes) -> 256 stays direct
    static constexpr bool CAN_STAGE = (LG >= 9);
    // tile rows wider than 256 tensor-map elements (N = 2048, 4096: 512 / 256 samples per row) are fetched as R0 plain bulk copies
    static constexpr bool STAGE_BULK = (LG == 11 || LG == 12);
};

struct WfKernelParams {
    const void* iq;                 // [batch][n_avg][N] samples
    const float2* wtab;             // master twiddle table, N entries
    const float* win;               // first half of the periodic Hann window, N/2 entries
    const float* thr;               // 257 thresholds (thr[256] = +inf)
    ssdr_wf_display_t* disp;        // [batch], low_clip_db/dynamic_range updated when auto_scale
    uint8_t* pixels;                // [batch][N] or null
    float* colour;                  // [batch][N] or null
    float* spectrum;                // [batch][N] or null
    ssdr_wf_scalars_t* scalars;     // [batch] or null
    const uint8_t* lines;           // colorrow entry: [batch][n_avg][N] uint8 lines (else null)
    uint16_t* sums;                 // large-N path: per-bin byte sums out [batch / rf][rf * N] (no colour stage), else null
    int rf;                         // large-N path: front radix (virtual channel = channel * rf + q)
    int prefetch;                   // 1: bulk-prefetch the next frame into L2 (local input); 0 for peer (NVLink) input, where the
                                    // bulk prefetch is pathologically slow (measured 120x) and the loads stream at link rate anyway
    int batch, n_avg;
    int p_lo;
    float p_gamma;
    float est_c1, est_c0;           // byte value ~= log2(P) * c1 + c0   (rounded to nearest = the byte)
    int key_bits;                   // bits needed for 255 * n_avg (selection iterations)
    int stagger;                    // cycles between the warp groups' starts of the warp-local passes (0: the size's default)
};

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

// Ask the memory system to bring [p, p + bytes) into L2 (one bulk request, no destination).
SSDR_DEV void prefetch_l2(const void* p, unsigned bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}

// ---- |X|^2 -> Kiwi byte ------------------------------------------------------------------------
// Spec: byte = largest k with P >= T[k].  Fast path: v = log2(P) c1 + c0 estimates
// the un-rounded byte value to ~1e-4; its nearest integer IS the byte unless v lies within kQEps of
// a rounding boundary, and only then the thresholds are consulted (exact, rare).
constexpr float kQMagic = 12582912.0f;    // 1.5 * 2^23: (v + magic) rounds v to nearest in the low mantissa bits
constexpr unsigned kQMagicBits = 0x4b400000u;
constexpr float kQEps = 1.25e-4f;         // estimate error <= 8.7e-5: lg2.approx 2^-22 relative (|log2 P| <= 100) x c1, fma rounding, c0/c1 rounding

SSDR_DEV float lg2_ftz(float x) {          // MUFU.LG2; a subnormal power is far below T[1], so flushing it to 0 is exact
    float y;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// exact bytes of two bins (byte0 | byte1 << 16) from the thresholds, starting at the estimates
__device__ __noinline__ unsigned quantise_pair_exact(float P0, float P1, float v0, float v1, const float* thr) {
    unsigned k[2];
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const float P = i ? P1 : P0;
        const float v = fminf(fmaxf(i ? v1 : v0, 0.0f), 255.0f);      // P == 0 -> -inf -> 0; NaN -> 0 like the oracle
        int kk = __float2int_rn(v);
        while (kk < 255 && P >= __ldg(thr + kk + 1)) ++kk;
        while (kk > 0 && P < __ldg(thr + kk)) --kk;
        k[i] = (unsigned)kk;
    }
    return k[0] | k[1] << 16;
}

// Eight bins -> four words byte0 | byte1 << 16.  The eight estimates are in flight together (instruction-level
// parallelism across the MUFU latency) and ONE test covers all of them: some |v - rint(v)| within kQEps of a
// rounding boundary, or some v + magic outside [magic, magic + 255] (estimate out of range, infinite or NaN:
// no clamps on the fast path).  Only then the thresholds are consulted (exact, rare).
// NP pairs of bins (NP = 4: eight bins, NP = 8: sixteen) -> NP words byte0 | byte1 << 16; one boundary / range test and one
// branch for all of them.
template <int NP>
SSDR_DEV void quantise_pairs(const float2* x, unsigned (&r)[NP], const WfKernelParams& kp) {
    float P[2 * NP];
#pragma unroll
    for (int i = 0; i < 2 * NP; ++i) {
        const float t = x[i].y * x[i].y;
        P[i] = __fmaf_rn(x[i].x, x[i].x, t);
    }
    float2 v[NP], m[NP];
    float worst = 0.0f;
    unsigned range = 0u;
    const float2 c1 = make_float2(kp.est_c1, kp.est_c1), c0 = make_float2(kp.est_c0, kp.est_c0);
#pragma unroll
    for (int i = 0; i < NP; ++i) {
        v[i] = __ffma2_rn(make_float2(lg2_ftz(P[2 * i]), lg2_ftz(P[2 * i + 1])), c1, c0);
        m[i] = __fadd2_rn(v[i], make_float2(kQMagic, kQMagic));
        const float2 nf = __fadd2_rn(m[i], make_float2(-kQMagic, -kQMagic));
        const float2 dd = __fadd2_rn(v[i], make_float2(-nf.x, -nf.y));
        worst = fmaxf(worst, fmaxf(fabsf(dd.x), fabsf(dd.y)));
        range |= (__float_as_uint(m[i].x) ^ kQMagicBits) | (__float_as_uint(m[i].y) ^ kQMagicBits);   // < 256 iff both in range
    }
#pragma unroll
    for (int i = 0; i < NP; ++i) r[i] = __byte_perm(__float_as_uint(m[i].x), __float_as_uint(m[i].y), 0x5410);   // low 16 bits of m = nearest integer
    if (worst > 0.5f - kQEps || range >= 256u) {
#pragma unroll
        for (int i = 0; i < NP; ++i) {
            const float2 nf = __fadd2_rn(m[i], make_float2(-kQMagic, -kQMagic));
            const float2 dd = __fadd2_rn(v[i], make_float2(-nf.x, -nf.y));
            const unsigned rg = (__float_as_uint(m[i].x) ^ kQMagicBits) | (__float_as_uint(m[i].y) ^ kQMagicBits);
            if (!(fmaxf(fabsf(dd.x), fabsf(dd.y)) <= 0.5f - kQEps) || rg >= 256u)
                r[i] = quantise_pair_exact(P[2 * i], P[2 * i + 1], v[i].x, v[i].y, kp.thr);
        }
    }
}
SSDR_DEV uint4 quantise8(const float2* x, const WfKernelParams& kp) {
    unsigned r[4];
    quantise_pairs<4>(x, r, kp);
    return make_uint4(r[0], r[1], r[2], r[3]);
}

// ---------------------------------------------------------------------------------------------
// FFT passes.  Element at logical position p of the frame lives at d[p + (p >> 5)].
// ---------------------------------------------------------------------------------------------
// First pass, part 1: the R0 samples of first-pass butterfly i of this thread -> registers (coalesced
// 8-byte loads).  Butterfly 0 is loaded before the barrier that frees the frame buffer, so its latency
// hides behind the other warps; the others are in flight while the preceding butterfly computes.
template <class C, int FMT>
SSDR_DEV void first_load(float2 (&x)[C::R0], int i, int t, const void* src, size_t off) {
#pragma unroll
#if SSDR_EXP & 32
    for (int m = 0; m < C::R0; ++m) x[m] = make_float2((float)(t + m), (float)(i - m) + (float)off);
#else
    for (int m = 0; m < C::R0; ++m) x[m] = load_iq<FMT>(src, off + (size_t)(t + i * C::G + m * C::M0));
#endif
}

// First pass, part 2: window, radix-R0 butterfly and twiddles of butterfly i, in registers.
template <class C, bool WINDOW>
SSDR_DEV void first_math(float2 (&x)[C::R0], int i, const float2* tw0, const float* win, int t, float2 w1, unsigned tm_tw = 0u, int jcol = -1) {
    constexpr int R = C::R0, M = C::M0, G = C::G;
    constexpr bool TABLE = (M == 32);
    const int j = (jcol >= 0) ? jcol : t + i * G;
#if SSDR_EXP & 256
    return;
#endif
    if constexpr (WINDOW) {
        // Hann values of the samples j + m M, m < R/2 (all < N/2), from the shared-memory table; the other half
        // of the butterfly uses w[n + N/2] = 1 - w[n], folded into the first level
        float wv[R / 2];
#pragma unroll
        for (int m = 0; m < R / 2; ++m) wv[m] = win[j + m * M];
        l1_window<R>(x, wv);
    } else {
        l1<R>(x);
    }
    dft_rest<R>(x);
    if constexpr (TABLE) {
#pragma unroll
        for (int q = 1; q < R; ++q) x[q] = cmul(x[q], tw0[(q - 1) * 32 + j]);
    } else if constexpr (C::TW_DIRECT) {
        // table twiddles W_N^(j q), q = 1..15, from this thread's tensor-memory words (two loads of eight twiddles)
        unsigned w[16];
        tmem_ld16(tm_tw + 32u * (unsigned)i, w);
        tmem_wait_ld();
#pragma unroll
        for (int q = 1; q <= 8; ++q) x[q] = cmul(x[q], make_float2(__uint_as_float(w[2 * q - 2]), __uint_as_float(w[2 * q - 1])));
        tmem_ld16(tm_tw + 32u * (unsigned)i + 16u, w);
        tmem_wait_ld();
#pragma unroll
        for (int q = 9; q < R; ++q) x[q] = cmul(x[q], make_float2(__uint_as_float(w[2 * q - 18]), __uint_as_float(w[2 * q - 17])));
    } else {
        tw_two_level<R>(x, w1);
    }
}

// First pass, part 3: scatter the outputs of butterfly i into the frame buffer.
template <class C>
SSDR_DEV void first_store(const float2 (&x)[C::R0], int i, float2* d, int t, int jcol = -1) {
    constexpr int R = C::R0, M = C::M0;
    const int j = (jcol >= 0) ? jcol : t + i * C::G;
    float2* o = (M == 32) ? d + j : d + j + 2 * (j >> 5);
#pragma unroll
    for (int q = 0; q < R; ++q) o[q * (M + M / 16)] = x[q];
}

// ---------------------------------------------------------------------------------------------
// Split-phase hand-off of the frame buffer (one-CTA-per-frame sizes: LG 13, 14).  The buffer may be overwritten by
// the next frame's first pass once every warp has LOADED its last-pass inputs: each warp arrives on an mbarrier
// right after those loads and only waits just before its first store of the next frame -- after the butterfly,
// quantiser and the next frame's first HBM loads and first butterfly, which need no shared memory.
// ---------------------------------------------------------------------------------------------
SSDR_DEV void mbar_init(unsigned long long* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(count) : "memory");
}
SSDR_DEV void mbar_arrive(unsigned long long* bar) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.release.cta.shared::cta.b64 st, [%0];\n\t}" ::"r"((unsigned)__cvta_generic_to_shared(bar)) : "memory");
}
SSDR_DEV void mbar_wait(unsigned long long* bar, unsigned parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.acquire.cta.shared::cta.b64 p, [%0], %1;\n\t"
        "@!p bra WAIT_%=;\n\t}" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(parity) : "memory");
}
// (A guarded variant of this wait -- trap after 2^32 cycles instead of spinning for ever on a lost tile copy -- was measured: it
// costs 1.5 % on config 2 and 3 % at 1024 points even when only the tile waits carry it, so the waits stay plain; the tile
// geometry is fixed at compile time per size and covered by the parity tests at every size and both sample formats.)

// ---- TMA staging of the raw frame (STAGED kernels) ----------------------------------------------
// expect `bytes` of bulk-copy traffic on the barrier and arrive once
SSDR_DEV void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.release.cta.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(bytes) : "memory");
}
// one bulk copy global -> shared memory by the TMA engine; completion is counted in bytes on the mbarrier
SSDR_DEV void bulk_g2s(void* dst_smem, const void* src, unsigned bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"((unsigned)__cvta_generic_to_shared(dst_smem)), "l"(src), "r"(bytes), "r"((unsigned)__cvta_generic_to_shared(bar)) : "memory");
}
SSDR_DEV void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
// one 2-D tile global -> shared memory (UTMALDG): box of the tensor map at element coordinates (x, y)
SSDR_DEV void tma_tile_2d(void* dst_smem, const CUtensorMap* tmap, int x, int y, unsigned long long* bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"((unsigned)__cvta_generic_to_shared(dst_smem)), "l"(tmap), "r"(x), "r"(y), "r"((unsigned)__cvta_generic_to_shared(bar)) : "memory");
}

// sample m (row) of first-pass butterfly i of this lane from the warp's staged tile [16 rows][64 samples]
template <int FMT, int ROW>      // ROW = samples per tile row (32 per first-pass butterfly of a lane)
SSDR_DEV float2 staged_sample(const unsigned char* region, int m, int i, int lane) {
    if constexpr (FMT == SSDR_IQ_CF32) {
        return reinterpret_cast<const float2*>(region)[m * ROW + i * 32 + lane];
    } else {
        const unsigned v = reinterpret_cast<const unsigned*>(region)[m * ROW + i * 32 + lane];
        const unsigned sw = __byte_perm(v, 0u, 0x2301);
        const int re = (int)(short)(sw & 0xffffu), im = (int)sw >> 16;
        return make_float2((float)re, (float)im);
    }
}

// radix-32 pass over sub-transforms of length 1024 (NP == 3 only): table twiddles W_1024^(j q)
// TM_MID (LG 14, round 2): the 31 twiddles of this lane come from the thread's tensor-memory words (tm_mid .. + 64) in
// four chunks of eight, the next chunk in flight while the current one multiplies, instead of 31 shared-memory loads.
template <class C, bool TM_MID = false>
SSDR_DEV void pass_mid(float2* d, const float2* tw1, int t, unsigned tm_mid = 0u) {
    const int j = t & 31;
    float2* p = d + (t >> 5) * (1024 + 64) + j;
    float2 x[32];
#pragma unroll
    for (int m = 0; m < 32; ++m) x[m] = p[34 * m];
    if constexpr (TM_MID) {
        unsigned wa[16], wb[16];
        tmem_ld16(tm_mid, wa);                       // twiddles q = 1..8 arrive during the butterfly
#if !(SSDR_EXP & 8)
        dft<32>(x);
#endif
        tmem_wait_ld();
        tmem_ld16(tm_mid + 16u, wb);                 // q = 9..16
#pragma unroll
        for (int k = 0; k < 8; ++k) x[1 + k] = cmul(x[1 + k], make_float2(__uint_as_float(wa[2 * k]), __uint_as_float(wa[2 * k + 1])));
        tmem_wait_ld();
        tmem_ld16(tm_mid + 32u, wa);                 // q = 17..24
#pragma unroll
        for (int k = 0; k < 8; ++k) x[9 + k] = cmul(x[9 + k], make_float2(__uint_as_float(wb[2 * k]), __uint_as_float(wb[2 * k + 1])));
        tmem_wait_ld();
        tmem_ld16(tm_mid + 48u, wb);                 // q = 25..31 (+ padding)
#pragma unroll
        for (int k = 0; k < 8; ++k) x[17 + k] = cmul(x[17 + k], make_float2(__uint_as_float(wa[2 * k]), __uint_as_float(wa[2 * k + 1])));
        tmem_wait_ld();
#pragma unroll
        for (int k = 0; k < 7; ++k) x[25 + k] = cmul(x[25 + k], make_float2(__uint_as_float(wb[2 * k]), __uint_as_float(wb[2 * k + 1])));
    } else {
#if !(SSDR_EXP & 8)
        dft<32>(x);
#endif
#pragma unroll
        for (int q = 1; q < 32; ++q) x[q] = cmul(x[q], tw1[(q - 1) * 32 + j]);
    }
#pragma unroll
    for (int q = 0; q < 32; ++q) p[34 * q] = x[q];
}

// last radix-32 pass (sub-transform length 32, no twiddles) + power + byte + accumulate
// The byte sums over the n_avg frames live in a thread-private uint4[4] column of shared memory (two
// uint16 lanes per word, no carry: <= 25500), touched with 128-bit accesses only.
// |X|^2 -> byte -> accumulate for the 32 outputs of a last-pass butterfly.  The byte sums over the n_avg
// frames live in a thread-private uint4[4] column of shared memory (two uint16 lanes per word, no carry:
// <= 25500), touched with 128-bit accesses only.  STRIDE = threads per CTA (column stride in uint4).
#ifndef SSDR_QGROUP
#define SSDR_QGROUP 8           // bins per quantiser group (one boundary test + branch each).  Measured (B200, config 2): 8: 1.582 ms, 16: 1.600, 32: 1.890
#endif
template <int STRIDE>
SSDR_DEV void last_epilogue(float2 (&x)[32], uint4* accs, bool first_frame, const WfKernelParams& kp, int tr_frame = 1 << 30) {
#if SSDR_EXP & 1
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        uint4 a = make_uint4(0u, 0u, 0u, 0u);
        if (!first_frame) a = accs[c * STRIDE];
        a.x += (__float_as_uint(x[8 * c].x) ^ __float_as_uint(x[8 * c + 1].y)) & 0xffu;
        a.y += (__float_as_uint(x[8 * c + 2].x) ^ __float_as_uint(x[8 * c + 3].y)) & 0xffu;
        a.z += (__float_as_uint(x[8 * c + 4].x) ^ __float_as_uint(x[8 * c + 5].y)) & 0xffu;
        a.w += (__float_as_uint(x[8 * c + 6].x) ^ __float_as_uint(x[8 * c + 7].y)) & 0xffu;
        accs[c * STRIDE] = a;
    }
#else
    constexpr int NP = SSDR_QGROUP / 2, NG = 32 / SSDR_QGROUP, WPG = NP / 4;      // pairs per group, groups, uint4 words per group
#pragma unroll
    for (int g = 0; g < NG; ++g) {
        if (g == 1) TRACE(12);
        if (g == 2) TRACE(13);
        if (g == 3) TRACE(14);
        uint4 a[WPG];
#pragma unroll
        for (int w = 0; w < WPG; ++w) a[w] = first_frame ? make_uint4(0u, 0u, 0u, 0u) : accs[(g * WPG + w) * STRIDE];
        unsigned r[NP];
        quantise_pairs<NP>(x + SSDR_QGROUP * g, r, kp);
#pragma unroll
        for (int w = 0; w < WPG; ++w) {
            a[w].x += r[4 * w]; a[w].y += r[4 * w + 1]; a[w].z += r[4 * w + 2]; a[w].w += r[4 * w + 3];
            accs[(g * WPG + w) * STRIDE] = a[w];
        }
    }
#endif
}

// last radix-32 pass (sub-transform length 32, no twiddles) + power + byte + accumulate
struct NoHook { SSDR_DEV void operator()() const {} };
template <class C, class Hook = NoHook>
SSDR_DEV void pass_last(const float2* d, int t, uint4* accs, bool first_frame, const WfKernelParams& kp, unsigned long long* bar, Hook after_loads = Hook(), int tr_frame = 1 << 30) {
    const float4* p = reinterpret_cast<const float4*>(d + 34 * t);       // 272-byte thread stride: 16-byte aligned, conflict-free
    float2 x[32];
#pragma unroll
    for (int m = 0; m < 16; ++m) {
        const float4 v = p[m];
        x[2 * m] = make_float2(v.x, v.y);
        x[2 * m + 1] = make_float2(v.z, v.w);
    }
    if constexpr (!std::is_same<Hook, NoHook>::value) {
        // STAGED: the first butterfly level consumes every loaded value, so when the warp meets behind it all its reads of
        // the region have COMPLETED (register dependences, no proxy fence needed) -> the hook bulk-copies the next frame's
        // samples into the region
        l1<32>(x);
        after_loads();
        TRACE(10);
        dft_rest<32>(x);
    } else {
        if constexpr (C::SPLIT) {             // this warp no longer needs the frame buffer
            __syncwarp();
            if ((threadIdx.x & 31) == 0) mbar_arrive(bar);
        }
        TRACE(10);
#if !(SSDR_EXP & 4)
        dft<32>(x);
#endif
    }
    TRACE(11);
    last_epilogue<C::THREADS>(x, accs, first_frame, kp, tr_frame);
}

// ---------------------------------------------------------------------------------------------
// group synchronisation: a frame group is G threads = part of a warp, a warp, several warps or the CTA
// ---------------------------------------------------------------------------------------------
template <int G>
SSDR_DEV unsigned group_mask(int lane) {
    if constexpr (G >= 32) return 0xffffffffu;
    else return ((1u << G) - 1u) << (lane & ~(G - 1));
}

template <class C>
SSDR_DEV void group_sync(int slot) {
    if constexpr (C::G <= 32) __syncwarp(group_mask<C::G>(threadIdx.x & 31));
    else if constexpr (C::FPC == 1) __syncthreads();
    else asm volatile("bar.sync %0, %1;" ::"r"(slot + 1), "n"(C::G) : "memory");
}

// ---------------------------------------------------------------------------------------------
// IEEE-754 round-to-nearest division a / b by a loop-invariant divisor (Markstein): with r = RN(1/b),
// q0 = RN(a r), e = a - q0 b (exact, one fma), q = RN(q0 + e r) is the correctly rounded quotient provided
// b is normal, its significand is not all ones and nothing under/overflows -- div_ok() checks the divisor,
// the dividends here (|a| < 2^9 or 0) are safe.  tests/test_oracle_tier_p.py checks the sequence against
// IEEE division on the CPU (every sum / n_avg pair, random colour quotients).
// ---------------------------------------------------------------------------------------------
struct Divisor {
    float b, r;
    bool ok;
};
SSDR_DEV Divisor make_divisor(float b) {
    Divisor d;
    d.b = b;
    d.r = __frcp_rn(b);
    const unsigned u = __float_as_uint(b), ex = (u >> 23) & 0xffu;
    d.ok = (ex >= 64u && ex <= 190u) && ((u & 0x7fffffu) != 0x7fffffu);
    return d;
}
template <bool FAST>
SSDR_DEV float div_rn(float a, const Divisor& d) {
    if constexpr (FAST) {
        const float q0 = __fmul_rn(a, d.r);
        const float e = __fmaf_rn(-q0, d.b, a);
        return __fmaf_rn(e, d.r, q0);
    } else {
        return __fdiv_rn(a, d.b);
    }
}

// ---------------------------------------------------------------------------------------------
// colour stage: order statistics from the register-resident sums, then the row through a transpose
// ---------------------------------------------------------------------------------------------
// Thread t of a group holds the sums of bins k = kbase(t) + G q, q = 0..31, as acc[q/2] halves
// (FFT entry) or k = 32 t + q (colorrow entry, LINEAR).  Output index o = k ^ N/2 (FFT) or k.
// Groups are independent: every barrier below is group-scoped.
template <class C, bool LINEAR>
SSDR_DEV void colour_stage(float* stage, int* red, int slot, int t, int ch, unsigned (&acc)[16], const WfKernelParams& kp,
                            int kb, int tB, int stage_words) {
    constexpr int N = C::N, G = C::G;
    const int lane = threadIdx.x & 31;
    const unsigned gmask = group_mask<G>(lane);
    const bool leader = (lane == 0);                    // used only when G > 32 (whole warps)

    const ssdr_wf_display_t dp = kp.disp[ch];
    TRACE(16);

    // wf_db[0] = wf_db[1] (utils_supersdr.py:791).  Output index o of (t, q):
    //   FFT order:  o = kbase(t) + G (q ^ 16), kbase = (t >> 5) + R0 (t & 31)  (NP == 3)  or  t  (NP == 2)
    //   linear:     o = 32 t + q
    // o == 0 is (t = 0, q = 16) / (t = 0, q = 0); o == 1 is (t = tB, q = 16) / (t = 0, q = 1).
    constexpr int q0 = LINEAR ? 0 : 16, q1 = LINEAR ? 1 : 16;
    auto key_at = [&](int q) -> unsigned { return (q & 1) ? (acc[q >> 1] >> 16) : (acc[q >> 1] & 0xffffu); };
    if (t == tB) red[5] = (int)key_at(q1);
    if (t == 0) { red[0] = 0; red[1] = 0; red[2] = 0x7fffffff; red[3] = 0; red[4] = 0x7fffffff; }
    group_sync<C>(slot);
    unsigned raw0 = 0;
    if (t == 0) {
        raw0 = key_at(q0);          // kiwi_waterfall.spectrum[0] itself is not patched
        const unsigned v1 = (unsigned)red[5];
        if (q0 & 1) acc[q0 >> 1] = (acc[q0 >> 1] & 0x0000ffffu) | (v1 << 16);
        else acc[q0 >> 1] = (acc[q0 >> 1] & 0xffff0000u) | v1;
    }

    float low_clip = dp.low_clip_db, high_clip = 0.f, dyn = dp.dynamic_range;
    const float fn = (float)kp.n_avg, z3 = (float)(3 * dp.zoom);
    const Divisor dfn = make_divisor(fn);               // n_avg = 1..100: always a valid Markstein divisor
    auto wfdb = [&](float s) { return ((div_rn<true>(s, dfn) - 255.0f) - 13.0f) + z3; };

    auto group_sum = [&](int v, int rslot) -> int {
        v = __reduce_add_sync(gmask, v);
        if constexpr (G <= 32) return v;
        else {
            if (leader) atomicAdd(&red[rslot], v);
            group_sync<C>(slot);
            return red[rslot];
        }
    };

    // ---- max ---------------------------------------------------------------------------------
    int kmax = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) kmax = max(kmax, (int)max(acc[i] & 0xffffu, acc[i] >> 16));
    kmax = __reduce_max_sync(gmask, kmax);
    if constexpr (G > 32) {
        if (leader) atomicMax(&red[3], kmax);
        group_sync<C>(slot);
        kmax = red[3];
    }
    const int vmax = kmax;
    TRACE(17);

    if (dp.auto_scale) {          // group-uniform
        const int want = kp.p_lo + 1;
        int v_lo, cnt_lo, mn = 0x7fffffff;
        const int nbins = 255 * kp.n_avg + 1;
        // Histogram or bisection?  Groups of at least 64 threads: histogram whenever it fits.  A warp-sized group (1024 bins): in
        // the FFT kernel up to 16 bins per thread (n_avg <= 2; beyond that the bisection, pure ALU work, overlaps the other warps'
        // transforms better than shared-memory atomics: 0.540 / 0.636 of HBM at n_avg = 4 / 8 against 0.524 / 0.609), in the
        // line-entry kernel, where the row stage is all there is, whenever it fits.  Measured against the old rule (groups of
        // >= 256 threads only; random lines, B200): line entry 4096 bins x 10 lines 0.44 -> 0.55 of HBM, 1024 bins x 1 line (the
        // reference's default averaging_n = 1, utils_supersdr.py:615) 414 -> 469 Mlines/s; FFT entry at n_avg = 1: 1024 / 2048 /
        // 4096 points 0.31 / 0.27 / 0.27 -> 0.35 / 0.31 / 0.32.
#ifndef SSDR_HIST_SMALL_GROUPS
#define SSDR_HIST_SMALL_GROUPS 1      // 0: histogram only for groups of >= 256 threads (the rule before round 2's last session; comparison builds)
#endif
        if (G >= 32 && (G >= 256 || (SSDR_HIST_SMALL_GROUPS && (G >= 64 || LINEAR || nbins <= 16 * G))) && nbins + 32 <= stage_words) {
            // ---- rank p_lo (0-based) from a histogram of the keys in the (idle) frame buffer: 32 shared-memory
            // atomics per thread, one scan.  The first barrier above already ordered every warp's last FFT pass.
            unsigned* hist = reinterpret_cast<unsigned*>(stage);
            int kmin = 0x7fffffff;
#pragma unroll
            for (int i = 0; i < 16; ++i) kmin = min(kmin, (int)min(acc[i] & 0xffffu, acc[i] >> 16));
            kmin = __reduce_min_sync(gmask, kmin);
            if (leader) atomicMin(&red[2], kmin);              // red[2] was set to INT_MAX below (see init)
            for (int i = t; i < nbins + 32; i += G) hist[i] = 0u;
            group_sync<C>(slot);
            kmin = red[2];
            if (kmin == vmax) {                                 // a constant row: every key is the same value
                v_lo = vmax; cnt_lo = N;
            } else {
#pragma unroll
                for (int q = 0; q < 32; ++q) atomicAdd(&hist[key_at(q)], 1u);
                group_sync<C>(slot);
                TRACE(18);
                const int cpt = ((nbins + G - 1) / G) | 1;      // odd chunk length: conflict-free chunk walks
                const int b0 = t * cpt, b1 = min(b0 + cpt, nbins);
                int sum = 0;
                for (int b = b0; b < b1; ++b) sum += (int)hist[b];
                int incl = sum;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) { const int u = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += u; }
                if (lane == 31) hist[nbins + (t >> 5)] = (unsigned)incl;
                group_sync<C>(slot);
                int prefix = incl - sum;
                for (int w = 0; w < (t >> 5); ++w) prefix += (int)hist[nbins + w];
                if (prefix < want && want <= prefix + sum) {    // exactly one thread owns the rank
                    int c = prefix, b = b0;
                    for (; b < b1; ++b) { c += (int)hist[b]; if (c >= want) break; }
                    red[0] = b; red[1] = c;
                }
                group_sync<C>(slot);
                v_lo = red[0]; cnt_lo = red[1];
            }
            // min{key > v_lo}
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                const int a = (int)(acc[i] & 0xffffu), b = (int)(acc[i] >> 16);
                if (a > v_lo) mn = min(mn, a);
                if (b > v_lo) mn = min(mn, b);
            }
            mn = __reduce_min_sync(gmask, mn);
            if (leader) atomicMin(&red[4], mn);
            group_sync<C>(slot);
            mn = red[4];
        } else {
        // ---- rank p_lo (0-based) by bisection on the key bits: smallest v with count(keys <= v) >= p_lo + 1.
        // Packed count: keys < 2^15, so (mid + 0x8000 - key) has bit 15 set iff key <= mid, per 16-bit half.
        if constexpr (G > 32) { if (t == 0) red[2] = 0; group_sync<C>(slot); }     // red[2] doubles as a bisection slot
        int lo = 0, hi = (1 << kp.key_bits) - 1;
#pragma unroll 1
        for (int it = 0; it < kp.key_bits; ++it) {
            const int mid = (lo + hi) >> 1;
            const unsigned M = ((unsigned)mid | 0x8000u) * 0x10001u;
            unsigned c = 0;
#pragma unroll
            for (int i = 0; i < 16; ++i) c += ((M - acc[i]) & 0x80008000u) >> 15;
            const int total = group_sum((int)((c & 0xffffu) + (c >> 16)), it % 3);
            if (total >= want) hi = mid; else lo = mid + 1;
            if constexpr (G > 32) { if (t == 0) red[(it + 2) % 3] = 0; }   // read by everyone before the previous barrier
        }
        v_lo = hi;
        // count(keys <= v_lo) and min{key > v_lo}
        int c = 0;
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            const int a = (int)(acc[i] & 0xffffu), b = (int)(acc[i] >> 16);
            c += (a <= v_lo) + (b <= v_lo);
            if (a > v_lo) mn = min(mn, a);
            if (b > v_lo) mn = min(mn, b);
        }
        mn = __reduce_min_sync(gmask, mn);
        if constexpr (G > 32) {
            group_sync<C>(slot);                               // all reads of the bisection slots are done
            if (t == 0) red[0] = 0;
            group_sync<C>(slot);
            if (leader) atomicMin(&red[4], mn);
        }
        cnt_lo = group_sum(c, 0);
        if constexpr (G > 32) mn = red[4];
        }
        const int v_hi = (cnt_lo >= want + 1 || mn == 0x7fffffff) ? v_lo : mn;
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
    TRACE(19);
    const float low = low_clip + (float)dp.delta_low_db;
    const float nf = dyn + (float)dp.delta_high_db;
    const float den = nf - (float)dp.delta_low_db;
    if (t == 0) {
        if (dp.auto_scale) { kp.disp[ch].low_clip_db = low_clip; kp.disp[ch].dynamic_range = dyn; }
        if (kp.scalars) {
            ssdr_wf_scalars_t s;
            s.low_clip_db = low_clip; s.high_clip_db = dp.auto_scale ? high_clip : wfdb((float)vmax);
            s.dynamic_range = dyn; s.wf_min_db = low - z3; s.wf_max_db = (low_clip + nf) - z3;
            kp.scalars[ch] = s;
        }
    }

    // ---- the row: colour value per bin, transposed to bin order through shared memory ---------------
    auto out_index = [&](int q) -> int {
        if constexpr (LINEAR) return 32 * t + q;
        else return kb + G * (q ^ 16);
    };
    auto sidx = [](int o) -> int {
        if constexpr (LINEAR) return o + (o >> 5);                       // thread-contiguous writes: pad, not swizzle
        else if constexpr (C::SWZ >= 0) return o ^ ((o >> C::SWZ) & 31);
        else return o;
    };
    const size_t row = (size_t)ch * N;
    group_sync<C>(slot);                                                 // every warp of the group is past its last FFT pass
    const Divisor dden = make_divisor(den);
    auto colour_of = [&](float s, auto fast) -> float {
        const float m = div_rn<true>(s, dfn);
        const float w = ((m - 255.0f) - 13.0f) + z3;
        float c = div_rn<decltype(fast)::value>(w - low, dden);
        c = fminf(fmaxf(c, 0.0f), 1.0f);
        c = c * 254.0f;
        return fminf(fmaxf(c, 0.0f), 255.0f);
    };
    // The colour value depends on the bin only through its integer key: when there are (many) fewer possible keys than
    // bins, the group fills a key -> colour table once (behind the row stage in the frame buffer) and every bin looks
    // its value up instead of repeating the two divisions.
    const int nkeys = 255 * kp.n_avg + 1;
    constexpr int row_words = LINEAR ? N + N / 32 : N;          // extent of the row stage (padded when LINEAR)
    // (worth it when filling the table costs at most half of the per-bin evaluations: nkeys / G <= 16 -- large groups at any
    // n_avg <= ~25, and a 1024-bin warp-sized group at n_avg <= 2, the reference's default averaging_n = 1, utils_supersdr.py:615)
    const bool use_lut = (SSDR_HIST_SMALL_GROUPS ? (2 * nkeys <= 32 * G) : (G >= 256)) && (4 * nkeys <= N) && (row_words + nkeys <= stage_words);
    if (use_lut) {
        float* lut = stage + row_words;
        auto fill = [&](auto fast) { for (int k = t; k < nkeys; k += G) lut[k] = colour_of((float)k, fast); };
        if (dden.ok) fill(std::true_type{}); else fill(std::false_type{});
        group_sync<C>(slot);
#pragma unroll
        for (int q = 0; q < 32; ++q) stage[sidx(out_index(q))] = lut[key_at(q)];
    } else {
        auto emit_row = [&](auto fast) {
#pragma unroll
            for (int q = 0; q < 32; ++q) stage[sidx(out_index(q))] = colour_of((float)key_at(q), fast);
        };
        if (dden.ok) emit_row(std::true_type{}); else emit_row(std::false_type{});     // group-uniform
    }
    group_sync<C>(slot);
    TRACE(20);
    // four consecutive outputs per thread and step: one 16-byte read of the (swizzled) stage, one 4-byte pixel store /
    // 16-byte colour store (a warp writes 128 / 512 contiguous bytes) instead of thirty-two 1-byte stores per thread
    auto stage4 = [&](int o4) -> float4 {
        if constexpr (LINEAR || C::SWZ == 0 || C::SWZ == 1) {          // padded layout / swizzle that varies inside the group
            return make_float4(stage[sidx(o4)], stage[sidx(o4 + 1)], stage[sidx(o4 + 2)], stage[sidx(o4 + 3)]);
        } else if constexpr (C::SWZ >= 2) {
            const int sw = (o4 >> C::SWZ) & 31;                         // constant over the aligned group of four
            float4 v = *reinterpret_cast<const float4*>(stage + (o4 ^ (sw & ~3)));
            if (sw & 1) { float u = v.x; v.x = v.y; v.y = u; u = v.z; v.z = v.w; v.w = u; }
            if (sw & 2) { float u = v.x; v.x = v.z; v.z = u; u = v.y; v.y = v.w; v.w = u; }
            return v;
        } else {
            return *reinterpret_cast<const float4*>(stage + o4);
        }
    };
#pragma unroll 4
    for (int i = 0; i < 8; ++i) {
        const int o4 = 4 * (t + i * G);
        const float4 c = stage4(o4);
        if (kp.colour) *reinterpret_cast<float4*>(kp.colour + row + o4) = c;
        if (kp.pixels) {
            // rint(c), 0 <= c <= 255, as the low byte of c + 1.5 * 2^23 (round to nearest even, exactly cvt.rni): the
            // conversion unit runs at a quarter of the fp32 rate and 16384 conversions per row were a tenth of the row stage
            const unsigned b0 = __float_as_uint(__fadd_rn(c.x, kQMagic)), b1 = __float_as_uint(__fadd_rn(c.y, kQMagic));
            const unsigned b2 = __float_as_uint(__fadd_rn(c.z, kQMagic)), b3 = __float_as_uint(__fadd_rn(c.w, kQMagic));
            const unsigned px = __byte_perm(__byte_perm(b0, b1, 0x0040), __byte_perm(b2, b3, 0x0040), 0x5410);
            *reinterpret_cast<unsigned*>(kp.pixels + row + o4) = px;
        }
    }
    TRACE(21);
    if (kp.spectrum) {                                                    // kiwi_waterfall.spectrum (optional output)
        group_sync<C>(slot);
#pragma unroll
        for (int q = 0; q < 32; ++q) {
            const float s = (t == 0 && q == q0) ? (float)raw0 : (float)key_at(q);
            stage[sidx(out_index(q))] = div_rn<true>(s, dfn);
        }
        group_sync<C>(slot);
#pragma unroll 4
        for (int i = 0; i < 8; ++i) {
            const int o4 = 4 * (t + i * G);
            *reinterpret_cast<float4*>(kp.spectrum + row + o4) = stage4(o4);
        }
    }
    // the caller's next barrier (before the frame buffer is written again) orders these reads
}

// Large-N path (N_total = rf * N): this group has the byte sums of sub-transform q of channel ch = vc / rf.
// Global bin q + rf k goes to output index (q + rf k) ^ (N_total / 2) = q + rf (k ^ N/2): transpose the sums
// to k order through shared memory exactly like the colour row, then store them rf apart (2 bytes each).
template <class C>
SSDR_DEV void sums_stage(unsigned* stage, int slot, int t, int vc, unsigned (&acc)[16], const WfKernelParams& kp, int kb) {
    constexpr int N = C::N, G = C::G;
    auto key_at = [&](int q) -> unsigned { return (q & 1) ? (acc[q >> 1] >> 16) : (acc[q >> 1] & 0xffffu); };
    auto sidx = [](int o) -> int {
        if constexpr (C::SWZ >= 0) return o ^ ((o >> C::SWZ) & 31);
        else return o;
    };
    group_sync<C>(slot);                                   // every warp of the group is past its last FFT pass
#pragma unroll
    for (int q = 0; q < 32; ++q) stage[sidx(kb + G * (q ^ 16))] = key_at(q);
    group_sync<C>(slot);
    const int ch = vc / kp.rf, qf = vc - ch * kp.rf;
    uint16_t* out = kp.sums + (size_t)ch * kp.rf * N + qf;
#pragma unroll 4
    for (int i = 0; i < 32; ++i) {
        const int o = t + i * G;
        out[(size_t)o * kp.rf] = (uint16_t)stage[sidx(o)];
    }
}

// ---------------------------------------------------------------------------------------------
// the fused waterfall kernel
// ---------------------------------------------------------------------------------------------
template <int LG, int FMT, bool WINDOW, bool STAGED = false>
__global__ void __launch_bounds__(Cfg<LG>::THREADS, Cfg<LG>::MIN_CTAS)
wf_fft_kernel(const WfKernelParams kp, const __grid_constant__ CUtensorMap tmap) {
    using C = Cfg<LG>;
    constexpr int N = C::N, G = C::G, FPC = C::FPC;
    static_assert(!STAGED || (C::CAN_STAGE && ((C::G > 32 && C::NB0 * C::R0 == 32 && C::M0 == 1024) || (C::G <= 32 && C::NP == 2))),
                  "staged input: 1024-point regions per warp (N >= 2048), or the whole frame per frame group (N <= 1024)");
    extern __shared__ __align__(16) unsigned char smem[];
    float2* data = reinterpret_cast<float2*>(smem + C::SM_DATA);
    float2* tw0 = reinterpret_cast<float2*>(smem + C::SM_TW0);
    float2* tw1 = reinterpret_cast<float2*>(smem + C::SM_TW1);
    uint4* accs = reinterpret_cast<uint4*>(smem + C::SM_ACC) + threadIdx.x;
    int* reds = reinterpret_cast<int*>(smem + C::SM_RED);

    const int slot = threadIdx.x / G, t = threadIdx.x % G;
    float2* d = data + (size_t)slot * C::PADN;
    int* red = reds + slot * 8;

    // one-time tables: twiddles of the table passes W_L^(j q), j < 32
    for (int e = threadIdx.x; e < C::TW0; e += blockDim.x) { const int q = e / 32 + 1, j = e & 31; tw0[e] = kp.wtab[j * q]; }
    for (int e = threadIdx.x; e < C::TW1; e += blockDim.x) { const int q = e / 32 + 1, j = e & 31; tw1[e] = kp.wtab[j * q * (N / 1024)]; }
    float* win = reinterpret_cast<float*>(smem + C::SM_WIN);
    if constexpr (WINDOW) {
        for (int e = threadIdx.x; e < N / 2; e += blockDim.x) win[e] = kp.win[e];
    }
    __syncthreads();

    // (cos, -sin)(2 pi j / N) of this thread's first-pass butterflies (window + first twiddle level):
    // loop invariant, parked in a thread-private shared-memory column (short, fixed latency)
    float2* w1s = reinterpret_cast<float2*>(smem + C::SM_W1);
    float2 w1r[C::W1_MODE == 2 ? C::NB0 : 1];
    if constexpr (C::W1_MODE == 1) {
#pragma unroll
        for (int i = 0; i < C::NB0; ++i) w1s[threadIdx.x + i * C::THREADS] = __ldg(kp.wtab + (STAGED ? (t >> 5) * (32 * C::NB0) + i * 32 + (t & 31) : t + i * G));
    } else if constexpr (C::W1_MODE == 2) {
#pragma unroll
        for (int i = 0; i < C::NB0; ++i) w1r[i] = __ldg(kp.wtab + (STAGED ? (t >> 5) * (32 * C::NB0) + i * 32 + (t & 31) : t + i * G));
    } else if constexpr (C::W1_MODE == 3) {
        for (int e = threadIdx.x; e < C::M0; e += blockDim.x) w1s[e] = kp.wtab[e];
        __syncthreads();
    }
    // first-pass column of butterfly i of this thread
    auto col_of = [&](int i) -> int {
        if constexpr (STAGED) return (t >> 5) * (32 * C::NB0) + i * 32 + (t & 31);      // warp w owns columns 32 NB0 w ..
        else return t + i * G;
    };
    auto w1_of = [&](int i) -> float2 {
        if constexpr (C::W1_MODE == 1) return w1s[threadIdx.x + i * C::THREADS];
        else if constexpr (C::W1_MODE == 2) return w1r[i];
        else if constexpr (C::W1_MODE == 3) return w1s[col_of(i)];
        else return make_float2(1.f, 0.f);
    };

    unsigned long long* bar = reinterpret_cast<unsigned long long*>(smem + C::SM_MBAR);
    unsigned long long* tbar = reinterpret_cast<unsigned long long*>(smem + C::SM_TBAR) + (C::G <= 32 ? threadIdx.x / G : threadIdx.x >> 5);   // STAGED: this warp's / group's
    unsigned frames_done = 0;                 // frames this group has finished (mbarrier phase counter)
    if constexpr (STAGED && C::G > 32) bar = reinterpret_cast<unsigned long long*>(smem + C::SM_GBAR) + slot;      // this frame group's
    if constexpr (C::SPLIT || STAGED) {
        if constexpr (STAGED && C::G > 32) { if (t == 0) mbar_init(bar, G / 32); }
        else if constexpr (C::SPLIT) { if (threadIdx.x == 0) mbar_init(bar, G / 32); }
        if constexpr (STAGED) { if ((C::G <= 32 ? t : (threadIdx.x & 31)) == 0) mbar_init(tbar, 1); }
        __syncthreads();
    }
    unsigned tm_tw = 0u;                      // this thread's tensor-memory words (TW_DIRECT)
    unsigned* tm_slot = reinterpret_cast<unsigned*>(smem + C::SM_MBAR) + 2;
    if constexpr (C::TW_DIRECT) {
        const int warp = threadIdx.x >> 5;
        if (warp == 0) tmem_alloc_cols(tm_slot, C::TM_COLS);
        tmem_fence_before();
        __syncthreads();
        tmem_fence_after();
        tm_tw = *tm_slot + ((unsigned)(32 * (warp & 3)) << 16) + (unsigned)((warp >> 2) * C::TM_WORDS);
#pragma unroll
        for (int i = 0; i < C::NB0; ++i) {
            const int j = col_of(i);
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                unsigned v[16];
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const int q = 8 * half + k + 1;                       // q = 1..16 (16: padding)
                    const float2 w = (q < C::R0) ? __ldg(kp.wtab + ((j * q) & (N - 1))) : make_float2(0.f, 0.f);
                    v[2 * k] = __float_as_uint(w.x); v[2 * k + 1] = __float_as_uint(w.y);
                }
                tmem_st16(tm_tw + 32u * (unsigned)i + 16u * (unsigned)half, v);
            }
        }
        if constexpr (C::TM_MID) {
#pragma unroll
            for (int c4 = 0; c4 < 4; ++c4) {
                unsigned v[16];
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const int q = 8 * c4 + k + 1;                         // q = 1..32 (32: padding)
                    const float2 w = (q < 32) ? __ldg(kp.wtab + (t & 31) * q * (N / 1024)) : make_float2(0.f, 0.f);
                    v[2 * k] = __float_as_uint(w.x); v[2 * k + 1] = __float_as_uint(w.y);
                }
                tmem_st16(tm_tw + 64u + 16u * (unsigned)c4, v);
            }
        }
        tmem_wait_st();
    }

    constexpr unsigned sample_bytes = (FMT == SSDR_IQ_CF32) ? 8u : 4u;
    const int ch_stride = (int)gridDim.x * FPC;
#ifdef SSDR_TRACE
    int tr_frame = 0;
#endif
    if constexpr (STAGED && C::G <= 32) {
        // ---- TMA-staged frame loop, one frame group (a warp or part of one) per channel: 256 .. 1024 points; 1024 is the
        // reference's waterfall size (utils_supersdr.py:596).  The group's whole frame is ONE tile [R0 rows][32 samples] that lands
        // in the group's own frame buffer while the previous frame's last pass runs; no barrier beyond the group's own.
        unsigned char* region = reinterpret_cast<unsigned char*>(d);
        const unsigned char* iq8 = static_cast<const unsigned char*>(kp.iq);
        auto stage_issue = [&](int fr) {
            if (t == 0) {
                mbar_expect_tx(tbar, (unsigned)(C::R0 * 32) * sample_bytes);
                tma_tile_2d(region, &tmap, 0, fr * C::R0, tbar);
            }
        };
        for (int ch = blockIdx.x * FPC + slot; ch < kp.batch; ch += ch_stride) {
            size_t off = (size_t)ch * kp.n_avg * N;
            int fr = ch * kp.n_avg;
            if (t == 0) fence_proxy_async();            // the row stage wrote the buffer through the generic proxy
            stage_issue(fr);
#pragma unroll 1
            for (int f = 0; f < kp.n_avg; ++f) {
                if (kp.prefetch && t == 0) {
                    const bool last = (f + 1 == kp.n_avg);
                    const size_t nxt = last ? (size_t)(ch + ch_stride) * kp.n_avg * N : off + N;
                    if (!last || ch + ch_stride < kp.batch) prefetch_l2(iq8 + nxt * sample_bytes, (unsigned)N * sample_bytes);
                }
                float2 x[C::NB0][C::R0];
                mbar_wait(tbar, frames_done & 1u);
#pragma unroll
                for (int i = 0; i < C::NB0; ++i)
#pragma unroll
                    for (int m = 0; m < C::R0; ++m) x[i][m] = staged_sample<FMT, 32>(region, m, 0, t + i * G);
                group_sync<C>(slot);                    // every thread of the group has its samples: the outputs may overwrite the tile
#pragma unroll
                for (int i = 0; i < C::NB0; ++i) {
                    first_math<C, WINDOW>(x[i], i, tw0, win, t, make_float2(1.f, 0.f));
                    first_store<C>(x[i], i, d, t);
                }
                group_sync<C>(slot);
                const bool more = (f + 1 < kp.n_avg);
                ++fr;
                pass_last<C>(d, t, accs, f == 0, kp, bar, [&]() { group_sync<C>(slot); if (more) stage_issue(fr); });
                ++frames_done;
                off += N;
            }
            unsigned acc[16];
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const uint4 a = accs[c * C::THREADS];
                acc[4 * c] = a.x; acc[4 * c + 1] = a.y; acc[4 * c + 2] = a.z; acc[4 * c + 3] = a.w;
            }
            if (kp.sums) sums_stage<C>(reinterpret_cast<unsigned*>(d), slot, t, ch, acc, kp, t);
            else colour_stage<C, false>(reinterpret_cast<float*>(d), red, slot, t, ch, acc, kp, t, 1, 2 * C::PADN);
            group_sync<C>(slot);                        // the row stage has been read before the next channel's tile lands
        }
    } else
    if constexpr (STAGED) {
        // ---- TMA-staged frame loop (8192 / 16384 points, local input) ---------------------------------
        // Per frame and warp: wait for the warp's tile (copied by the TMA engine into the warp's own 1024-point region of the
        // frame buffer while the previous frame's last pass and quantiser ran), pull the NB0 x R0 samples of each lane into
        // registers, tell the CTA the region is consumed (mbarrier `bar`, one arrival per warp), all first-pass butterflies,
        // wait until EVERY region is consumed (only then may first-pass outputs overwrite them), stores, CTA barrier,
        // warp-local passes.  Behind the first butterfly level of the last pass the region is free again and one lane issues
        // the tile copy of the next frame (one UTMALDG).
        const int warp = t >> 5, lane = t & 31;         // warp within the frame group = its 1024-point region
        unsigned char* region = reinterpret_cast<unsigned char*>(d + (size_t)warp * (1024 + 64));      // 8704 bytes >= the 8 KB tile
        const unsigned char* iq8 = static_cast<const unsigned char*>(kp.iq);
        // one tile [R0 rows][32 NB0 samples] (8 KB) of frame `fr` (global frame number): rows R0 fr .. of the input seen as
        // [frames x R0][1024 samples], columns 32 NB0 w ..; one UTMALDG by one lane -- or, where a row is wider than a tensor-map
        // box may be (N = 2048, 4096), R0 plain bulk copies of one row each by the same lane
        constexpr int inner_per_sample = (FMT == SSDR_IQ_CF32) ? 2 : 1;      // tensor-map elements are 32-bit words
        constexpr int ROW = 32 * C::NB0;                                     // samples per tile row
        auto stage_issue = [&](int fr) {
            if (lane == 0) {
                mbar_expect_tx(tbar, (unsigned)(C::R0 * ROW) * sample_bytes);
                if constexpr (C::STAGE_BULK) {
#pragma unroll
                    for (int m = 0; m < C::R0; ++m)
                        bulk_g2s(region + (size_t)m * ROW * sample_bytes, iq8 + ((size_t)fr * N + (size_t)m * 1024 + (size_t)warp * ROW) * sample_bytes,
                                 (unsigned)ROW * sample_bytes, tbar);
                } else {
                    tma_tile_2d(region, &tmap, warp * ROW * inner_per_sample, fr * C::R0, tbar);
                }
            }
        };
        int jc[C::NB0];
#pragma unroll
        for (int i = 0; i < C::NB0; ++i) jc[i] = col_of(i);
        for (int ch = blockIdx.x * FPC + slot; ch < kp.batch; ch += ch_stride) {
            size_t off = (size_t)ch * kp.n_avg * N;
            int fr = ch * kp.n_avg;                     // global frame number
            if (lane == 0) fence_proxy_async();         // the row stage wrote the buffer through the generic proxy
            stage_issue(fr);                            // frame 0: the buffer is free (kernel start / barrier after the row stage)
#pragma unroll 1
            for (int f = 0; f < kp.n_avg; ++f) {
                TRACE(0);
                if (kp.prefetch && t == 0) {            // the frame after next (or the next channel's first frame) -> L2
                    const bool last = (f + 1 == kp.n_avg);
                    const size_t nxt = last ? (size_t)(ch + ch_stride) * kp.n_avg * N : off + N;
                    if (!last || ch + ch_stride < kp.batch) prefetch_l2(iq8 + nxt * sample_bytes, (unsigned)N * sample_bytes);
                }
                float2 x[C::NB0][C::R0];
                mbar_wait(tbar, frames_done & 1u);
#pragma unroll
                for (int i = 0; i < C::NB0; ++i)
#pragma unroll
                    for (int m = 0; m < C::R0; ++m) x[i][m] = staged_sample<FMT, ROW>(region, m, i, lane);
                TRACE(1);
                __syncwarp();
                if (lane == 0) mbar_arrive(bar);        // this warp's region is consumed (release: the loads above are ordered before)
                // all first-pass butterflies BEFORE the wait: a warp that finished the previous frame early does its first-pass
                // arithmetic while the late warps still run their last pass; after the wait only the stores are left
#pragma unroll
                for (int i = 0; i < C::NB0; ++i) first_math<C, WINDOW>(x[i], i, tw0, win, t, w1_of(i), tm_tw, jc[i]);
                TRACE(2);
                mbar_wait(bar, frames_done & 1u);       // ... and so is everybody's
                TRACE(3);
#pragma unroll
                for (int i = 0; i < C::NB0; ++i) first_store<C>(x[i], i, d, t, jc[i]);
                TRACE(5);
                group_sync<C>(slot);
                TRACE(6);
                {   // stagger of the warp-local passes (section 5.1); the staged kernel is flat between 200 and 400 cycles per level
                    const int lvl = (LG >= 13) ? ((threadIdx.x >> (LG == 14 ? 7 : 6)) & 3) : ((threadIdx.x >> 5) & 3);
                    const int stg = kp.stagger > 0 ? kp.stagger : (LG >= 13 ? 300 : C::STAGGER);
                    if (lvl && stg > 1) { const long long c0 = clock64(); while (clock64() - c0 < lvl * stg) { } }
                }
                TRACE(7);
                pass_mid<C, C::TM_MID>(d, tw1, t, tm_tw + 64u);
                __syncwarp();
                TRACE(8);
                const bool more = (f + 1 < kp.n_avg);
                ++fr;
#ifdef SSDR_TRACE
                pass_last<C>(d, t, accs, f == 0, kp, bar, [&]() { __syncwarp(); if (more) stage_issue(fr); }, tr_frame);
#else
                pass_last<C>(d, t, accs, f == 0, kp, bar, [&]() { __syncwarp(); if (more) stage_issue(fr); });
#endif
                TRACE(9);
#ifdef SSDR_TRACE
                TRACE_FLUSH();
                ++tr_frame;
                if (threadIdx.x == 0 && blockIdx.x == 0) g_wf_trace_frames = tr_frame;
#endif
                ++frames_done;
                off += N;
            }
            unsigned acc[16];
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const uint4 a = accs[c * C::THREADS];
                acc[4 * c] = a.x; acc[4 * c + 1] = a.y; acc[4 * c + 2] = a.z; acc[4 * c + 3] = a.w;
            }
            if (kp.sums) sums_stage<C>(reinterpret_cast<unsigned*>(d), slot, t, ch, acc, kp, (t >> 5) + C::R0 * (t & 31));
            else colour_stage<C, false>(reinterpret_cast<float*>(d), red, slot, t, ch, acc, kp, (t >> 5) + C::R0 * (t & 31), 32, 2 * C::PADN);
            group_sync<C>(slot);                        // the row stage has been read before the next channel's tile lands
        }
    } else {
    for (int ch = blockIdx.x * FPC + slot; ch < kp.batch; ch += ch_stride) {
        float2 x0[C::R0];
        size_t off = (size_t)ch * kp.n_avg * N;
        first_load<C, FMT>(x0, 0, t, kp.iq, off);
#pragma unroll 1
        for (int f = 0; f < kp.n_avg; ++f) {
            TRACE(0);
            // the frame after next (or the first frame of this group's next channel) -> L2
            if (kp.prefetch && t == 0) {
                const bool last = (f + 1 == kp.n_avg);
                const size_t nxt = last ? (size_t)(ch + ch_stride) * kp.n_avg * N : off + N;
                if (!last || ch + ch_stride < kp.batch)
                    prefetch_l2(static_cast<const unsigned char*>(kp.iq) + nxt * sample_bytes, (unsigned)N * sample_bytes);
            }
            // Before the first store of this frame every thread of the group must have finished reading the previous
            // frame (or row).  SPLIT: the first butterfly (and the loads of the second) come first, then the wait on
            // the warps' arrivals of the previous frame; the row of the previous channel is covered by the barrier
            // after the colour stage.  Otherwise: one group barrier.
            auto buffer_free = [&]() {
#if !(SSDR_EXP & 64)
                if constexpr (C::SPLIT) { if (frames_done) mbar_wait(bar, (frames_done - 1) & 1u); }
                else group_sync<C>(slot);
#endif
            };
            if constexpr (C::NB0 == 1) {
                first_math<C, WINDOW>(x0, 0, tw0, win, t, w1_of(0), tm_tw);
                buffer_free();
                first_store<C>(x0, 0, d, t);
            } else {
                // software pipeline over this thread's first-pass butterflies: load i + 1 while i computes
                float2 xa[C::R0], xb[C::R0];
#pragma unroll
                for (int m = 0; m < C::R0; ++m) xa[m] = x0[m];
#pragma unroll
                for (int i = 0; i < C::NB0; i += 2) {
                    first_load<C, FMT>(xb, i + 1, t, kp.iq, off);
                    first_math<C, WINDOW>(xa, i, tw0, win, t, w1_of(i), tm_tw);
                    TRACE(1);
                    if (i == 0) buffer_free();
                    TRACE(2);
                    first_store<C>(xa, i, d, t);
                    if (i + 2 < C::NB0) first_load<C, FMT>(xa, i + 2, t, kp.iq, off);
                    TRACE(3);
                    first_math<C, WINDOW>(xb, i + 1, tw0, win, t, w1_of(i + 1), tm_tw);
                    TRACE(4);
                    first_store<C>(xb, i + 1, d, t);
                }
            }
            TRACE(5);
#if !(SSDR_EXP & 128)
            group_sync<C>(slot);
#endif
            TRACE(6);
            // from here each warp owns a contiguous 1024-point (NP == 3) / 32-point sub-transform: warp-local
            if constexpr (C::STAGGER > 0) {
                const int lvl = (LG == 14) ? ((threadIdx.x >> 7) & 3) : (LG == 13) ? ((threadIdx.x >> 6) & 3) : ((threadIdx.x >> 5) & 3);     // four levels
                const int stg = kp.stagger > 0 ? kp.stagger : C::STAGGER;
                if (lvl) { const long long c0 = clock64(); while (clock64() - c0 < lvl * stg) { } }   // (__nanosleep is too coarse: 2.08 ms)
            }
            TRACE(7);
#if !(SSDR_EXP & 16)
            if constexpr (C::NP == 3) {
                pass_mid<C, C::TM_MID>(d, tw1, t, tm_tw + 64u);
                __syncwarp();
            }
#endif
            TRACE(8);
            pass_last<C>(d, t, accs, f == 0, kp, bar);
            TRACE(9);
#ifdef SSDR_TRACE
            TRACE_FLUSH();
            ++tr_frame;
            if (threadIdx.x == 0 && blockIdx.x == 0) g_wf_trace_frames = tr_frame;
#endif
            ++frames_done;
            off += N;
            if (f + 1 < kp.n_avg) first_load<C, FMT>(x0, 0, t, kp.iq, off);
        }
        unsigned acc[16];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const uint4 a = accs[c * C::THREADS];
            acc[4 * c] = a.x; acc[4 * c + 1] = a.y; acc[4 * c + 2] = a.z; acc[4 * c + 3] = a.w;
        }
#if SSDR_EXP & 2
        if (acc[0] == 0x12345678u) kp.pixels[ch] = (uint8_t)acc[1];
#else
        if (kp.sums) sums_stage<C>(reinterpret_cast<unsigned*>(d), slot, t, ch, acc, kp, (C::NP == 3) ? ((t >> 5) + C::R0 * (t & 31)) : t);
        else
        colour_stage<C, false>(reinterpret_cast<float*>(d), red, slot, t, ch, acc, kp,
                               (C::NP == 3) ? ((t >> 5) + C::R0 * (t & 31)) : t, (C::NP == 3) ? 32 : 1, 2 * C::PADN);
#endif
        if constexpr (C::SPLIT) group_sync<C>(slot);     // the row stage has been read before the next channel's first store
    }
    }
    if constexpr (C::TW_DIRECT) {
        tmem_fence_before();
        __syncthreads();
        if ((threadIdx.x >> 5) == 0) tmem_free_cols(*tm_slot, C::TM_COLS);
    }
}

// Tier-P entry: finished uint8 lines in, same colour stage (no FFT).  utils_supersdr.py:783-813,881-886
template <int LG>
__global__ void __launch_bounds__(Cfg<LG>::THREADS)
wf_colorrow_kernel(const WfKernelParams kp) {
    using C = Cfg<LG>;
    constexpr int N = C::N, G = C::G, FPC = C::FPC;
    extern __shared__ __align__(16) unsigned char smem[];
    float* stages = reinterpret_cast<float*>(smem);
    int* reds = reinterpret_cast<int*>(smem + (size_t)FPC * C::PADN * sizeof(float));
    const int slot = threadIdx.x / G, t = threadIdx.x % G;
    float* stage = stages + (size_t)slot * C::PADN;
    int* red = reds + slot * 8;
    for (int ch = blockIdx.x * FPC + slot; ch < kp.batch; ch += (int)gridDim.x * FPC) {
        unsigned acc[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) acc[i] = 0u;
        // thread t owns bins 32 t .. 32 t + 31: two 16-byte loads per line, four lines (eight loads) in flight per thread
        const uint4* src0 = reinterpret_cast<const uint4*>(kp.lines + (size_t)ch * kp.n_avg * N + 32 * t);
        for (int f0 = 0; f0 < kp.n_avg; f0 += 4) {
            uint4 v[4][2];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                if (f0 + j < kp.n_avg) {
                    v[j][0] = __ldcs(src0 + (size_t)(f0 + j) * (N / 16));
                    v[j][1] = __ldcs(src0 + (size_t)(f0 + j) * (N / 16) + 1);
                } else {
                    v[j][0] = v[j][1] = make_uint4(0u, 0u, 0u, 0u);
                }
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const unsigned w[4] = {v[j][h].x, v[j][h].y, v[j][h].z, v[j][h].w};
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        acc[h * 8 + 2 * k] += __byte_perm(w[k], 0u, 0x4140);       // bytes 0, 1 -> two uint16 lanes
                        acc[h * 8 + 2 * k + 1] += __byte_perm(w[k], 0u, 0x4342);   // bytes 2, 3
                    }
                }
            }
        }
        // the lines of this group's next channel -> L2 while the colour stage runs (one bulk request)
        if (t == 0) {
            const int nx = ch + (int)gridDim.x * FPC;
            const size_t bytes = (size_t)kp.n_avg * N;
            if (nx < kp.batch && bytes <= (1u << 20) && (bytes & 15u) == 0) prefetch_l2(kp.lines + (size_t)nx * bytes, (unsigned)bytes);
        }
        group_sync<C>(slot);          // the previous row's transposed reads are done
        colour_stage<C, true>(stage, red, slot, t, ch, acc, kp, 0, 0, C::PADN);
    }
}

// ---------------------------------------------------------------------------------------------
// Large frames (N = 32768, 65536: more than one SM's shared memory).  Plan rf x 16 x 32 x 32:
//   1. wf_front_kernel: the radix-rf front pass -- window, butterfly, chain twiddles W_N^(j q) -- over HBM:
//      reads the frame once, writes the rf sub-frames y_q[j] (j < 16384) of every frame to a scratch buffer laid
//      out as rf "virtual channels" per channel;
//   2. wf_fft_kernel<14> without window on the scratch, in sums mode (per-bin byte sums, no colour stage);
//   3. wf_colour_big_kernel: rank selection (histogram) + colour row over the N sums of a channel.
// HBM traffic: 8 + 8 + 8 bytes per sample instead of 8 (stated in the roofline of these sizes).
// ---------------------------------------------------------------------------------------------
template <int RF, int FMT, bool WINDOW>
__global__ void __launch_bounds__(256) wf_front_kernel(const void* __restrict__ iq, float2* __restrict__ scratch,
                                                       const float2* __restrict__ wtab, const float* __restrict__ win,
                                                       int batch, int n_avg) {
    constexpr int M = 16384, N = RF * M;
    const size_t total = (size_t)batch * n_avg * M;
    for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
        const int j = (int)(e % M);
        const size_t fr = e / M;                           // ch * n_avg + f
        const size_t ch = fr / n_avg, f = fr - ch * n_avg;
        float2 x[RF];
#pragma unroll
        for (int b = 0; b < RF; ++b) x[b] = load_iq<FMT>(iq, fr * N + (size_t)(j + b * M));
        if constexpr (WINDOW) {
            float wv[RF / 2];
#pragma unroll
            for (int m = 0; m < RF / 2; ++m) wv[m] = __ldg(win + j + m * M);
            l1_window<RF>(x, wv);
        } else {
            l1<RF>(x);
        }
        dft_rest<RF>(x);
        tw_two_level<RF>(x, __ldg(wtab + j));
#pragma unroll
        for (int q = 0; q < RF; ++q) __stcs(scratch + ((ch * RF + q) * n_avg + f) * M + j, x[q]);
    }
}

// ---------------------------------------------------------------------------------------------
// Large frames, fused (opt-in, SSDR_WF_BIG=fused; measured slower than the three-kernel path, see wf_big_fused()): ONE persistent kernel does the front pass and the rf 16384-point sub-transforms of
// a frame back to back on the same SM, so the rf sub-frames travel through a per-CTA scratch (rf x 128 KB) that never
// leaves L2 -- HBM sees the samples once (8 B/sample) instead of three times (read, scratch write, scratch read).
// The byte sums of the rf sub-transforms (rf x 16 packed words per thread) live in TENSOR MEMORY (tmem_scratch.cuh):
// the SM's registers and shared memory are full (the 16384-point plan), its tensor memory is idle.  Arithmetic:
// exactly the three-kernel path's (same device functions, same order), so the results are bit-identical.
// ---------------------------------------------------------------------------------------------
SSDR_DEV void tmem_st4(unsigned addr, const uint4& v) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
SSDR_DEV uint4 tmem_ld4(unsigned addr) {
    uint4 v;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
    return v;
}

// last radix-32 pass + power + byte + accumulate into the thread's tensor-memory words (16 per sub-transform)
template <class C>
SSDR_DEV void pass_last_tmem(const float2* d, int t, unsigned tm_acc, bool first_frame, const WfKernelParams& kp, unsigned long long* bar) {
    const float4* p = reinterpret_cast<const float4*>(d + 34 * t);
    float2 x[32];
#pragma unroll
    for (int m = 0; m < 16; ++m) {
        const float4 v = p[m];
        x[2 * m] = make_float2(v.x, v.y);
        x[2 * m + 1] = make_float2(v.z, v.w);
    }
    if constexpr (C::SPLIT) {
        __syncwarp();
        if ((threadIdx.x & 31) == 0) mbar_arrive(bar);
    }
    dft<32>(x);
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        uint4 a = make_uint4(0u, 0u, 0u, 0u);
        if (!first_frame) { a = tmem_ld4(tm_acc + 4 * c); tmem_wait_ld(); }
        const uint4 k8 = quantise8(x + 8 * c, kp);
        a.x += k8.x; a.y += k8.y; a.z += k8.z; a.w += k8.w;
        tmem_st4(tm_acc + 4 * c, a);
    }
    tmem_wait_st();
}

struct WfBigParams {
    WfKernelParams kp;              // kp.wtab = 16384-point table, kp.sums = [batch][rf * 16384], kp.rf, kp.batch = channels
    const float2* wtab_big;         // N-point master table (front-pass chain twiddles W_N^j)
    const float* win_big;           // first half of the N-point Hann window (global memory)
    float2* scratch;                // [gridDim.x][rf][16384] complex64, L2-resident
};

template <int RF, int FMT, bool WINDOW>
__global__ void __launch_bounds__(512, 1)
wf_big_fused_kernel(const WfBigParams bp) {
    using C = Cfg<14>;
    constexpr int M = 16384, G = C::G;
    const WfKernelParams& kp = bp.kp;
    extern __shared__ __align__(16) unsigned char smem[];
    float2* d = reinterpret_cast<float2*>(smem + C::SM_DATA);
    float2* tw1 = reinterpret_cast<float2*>(smem + C::SM_TW1);
    unsigned* tm_slot = reinterpret_cast<unsigned*>(smem + C::SM_RED);
    unsigned long long* bar = reinterpret_cast<unsigned long long*>(smem + C::SM_MBAR);
    const int t = threadIdx.x, warp = t >> 5;

    for (int e = t; e < C::TW1; e += blockDim.x) { const int q = e / 32 + 1, j = e & 31; tw1[e] = kp.wtab[j * q * (M / 1024)]; }
    // tensor memory per thread: 64 words of first-pass twiddles (as wf_fft_kernel<14>) + RF x 16 words of byte sums;
    // four warps share a lane quarter: 4 x (64 + 16 RF) = 384 / 512 columns -> the whole tensor memory
    constexpr int TW_WORDS = 64;                                 // first-pass twiddles only (the middle pass keeps its shared-memory table here)
    constexpr int TM_PER_THREAD = TW_WORDS + 16 * RF;
    constexpr int TM_COLS = 512;
    static_assert(4 * TM_PER_THREAD <= 512, "tensor memory budget");
    if (warp == 0) tmem_alloc_cols(tm_slot, TM_COLS);
    tmem_fence_before();
    if (t == 0) mbar_init(bar, G / 32);
    __syncthreads();
    tmem_fence_after();
    // this thread's tensor-memory words: lane quarter (warp & 3), TM_PER_THREAD columns per warp of the quarter
    const unsigned tm_tw = *tm_slot + ((unsigned)(32 * (warp & 3)) << 16) + (unsigned)((warp >> 2) * TM_PER_THREAD);
    const unsigned tm_base = tm_tw + (unsigned)TW_WORDS;
#pragma unroll
    for (int i = 0; i < C::NB0; ++i) {
        const int j = t + i * G;
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            unsigned v[16];
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const int q = 8 * half + k + 1;
                const float2 w = (q < C::R0) ? __ldg(kp.wtab + ((j * q) & (M - 1))) : make_float2(0.f, 0.f);
                v[2 * k] = __float_as_uint(w.x); v[2 * k + 1] = __float_as_uint(w.y);
            }
            tmem_st16(tm_tw + 32u * (unsigned)i + 16u * (unsigned)half, v);
        }
    }
    tmem_wait_st();
    unsigned frames_done = 0;
    float2* scr = bp.scratch + (size_t)blockIdx.x * RF * M;
    constexpr unsigned sample_bytes = (FMT == SSDR_IQ_CF32) ? 8u : 4u;

    for (int ch = blockIdx.x; ch < kp.batch; ch += (int)gridDim.x) {
#pragma unroll 1
        for (int f = 0; f < kp.n_avg; ++f) {
            const size_t fr = (size_t)ch * kp.n_avg + f;
            // the next frame -> L2 while this one is transformed (one bulk request)
            if (kp.prefetch && t == 0 && (f + 1 < kp.n_avg || ch + (int)gridDim.x < kp.batch)) {
                const size_t nxt = (f + 1 < kp.n_avg) ? fr + 1 : (size_t)(ch + gridDim.x) * kp.n_avg;
                prefetch_l2(static_cast<const unsigned char*>(kp.iq) + nxt * (size_t)RF * M * sample_bytes, (unsigned)(RF * M) * sample_bytes);
            }
            // ---- front pass (radix RF over stride 16384): window, butterfly, chain twiddles W_N^(j q) -> scratch --------
#pragma unroll 4
            for (int i = 0; i < M / 512; ++i) {
                const int j = t + 512 * i;
                float2 x[RF];
#pragma unroll
                for (int b = 0; b < RF; ++b) x[b] = load_iq<FMT>(kp.iq, fr * (size_t)(RF * M) + (size_t)(j + b * M));
                if constexpr (WINDOW) {
                    float wv[RF / 2];
#pragma unroll
                    for (int m = 0; m < RF / 2; ++m) wv[m] = __ldg(bp.win_big + j + m * M);
                    l1_window<RF>(x, wv);
                } else {
                    l1<RF>(x);
                }
                dft_rest<RF>(x);
                tw_two_level<RF>(x, __ldg(bp.wtab_big + j));
#pragma unroll
                for (int q = 0; q < RF; ++q) __stcg(scr + (size_t)q * M + j, x[q]);
            }
            __syncthreads();                                      // the sub-frames are in L2, visible to the whole CTA
            // ---- the RF 16384-point sub-transforms (no window), byte sums of sub-transform q in tensor memory ---------------
#pragma unroll 1
            for (int q = 0; q < RF; ++q) {
                const float2* src = scr + (size_t)q * M;
                float2 xa[C::R0], xb[C::R0];
#pragma unroll
                for (int m = 0; m < C::R0; ++m) xa[m] = __ldcg(src + t + m * C::M0);
#pragma unroll
                for (int m = 0; m < C::R0; ++m) xb[m] = __ldcg(src + t + G + m * C::M0);
                first_math<C, false>(xa, 0, nullptr, nullptr, t, make_float2(1.f, 0.f), tm_tw);
                if (frames_done) mbar_wait(bar, (frames_done - 1) & 1u);
                first_store<C>(xa, 0, d, t);
                first_math<C, false>(xb, 1, nullptr, nullptr, t, make_float2(1.f, 0.f), tm_tw);
                first_store<C>(xb, 1, d, t);
                __syncthreads();
                {
                    const int lvl = (t >> 7) & 3;
                    if (lvl) { const long long c0 = clock64(); while (clock64() - c0 < lvl * C::STAGGER) { } }
                }
                pass_mid<C>(d, tw1, t);
                __syncwarp();
                pass_last_tmem<C>(d, t, tm_base + 16 * q, f == 0, kp, bar);
                ++frames_done;
            }
            // every sub-transform's first-pass loads were followed by a CTA barrier: the next front pass may overwrite the scratch
        }
        // ---- byte sums of the RF sub-transforms -> global sums in output order (rf apart), as the three-kernel path --------
#pragma unroll 1
        for (int q = 0; q < RF; ++q) {
            unsigned acc[16];
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const uint4 a = tmem_ld4(tm_base + 16 * q + 4 * c);
                acc[4 * c] = a.x; acc[4 * c + 1] = a.y; acc[4 * c + 2] = a.z; acc[4 * c + 3] = a.w;
            }
            tmem_wait_ld();
            sums_stage<C>(reinterpret_cast<unsigned*>(d), 0, t, ch * RF + q, acc, kp, (t >> 5) + C::R0 * (t & 31));
            __syncthreads();                                      // the stage is the frame buffer
        }
    }
    tmem_fence_before();
    __syncthreads();
    if (warp == 0) tmem_free_cols(*tm_slot, TM_COLS);
}

// One CTA per channel: sums uint16[N] (output order) -> scalars, colour row, pixels, spectrum.
// Same arithmetic as colour_stage (utils_supersdr.py:787-813); the keys are re-read from
// global memory (L2) instead of living in registers.
__global__ void __launch_bounds__(1024) wf_colour_big_kernel(const WfKernelParams kp, int N) {
    extern __shared__ __align__(16) unsigned char smem[];
    unsigned* hist = reinterpret_cast<unsigned*>(smem);
    const int nbins = 255 * kp.n_avg + 1;
    int* red = reinterpret_cast<int*>(hist + nbins + 32);
    const int t = threadIdx.x, lane = t & 31, G = blockDim.x;
    for (int ch = blockIdx.x; ch < kp.batch; ch += gridDim.x) {
        const uint16_t* keys = kp.sums + (size_t)ch * N;
        const ssdr_wf_display_t dp = kp.disp[ch];
        for (int i = t; i < nbins + 32; i += G) hist[i] = 0u;
        if (t == 0) { red[0] = 0; red[1] = 0; red[2] = 0x7fffffff; red[3] = 0; red[4] = 0x7fffffff; }
        __syncthreads();
        const unsigned key1 = keys[1];                      // wf_db[0] = wf_db[1]
        int kmax = 0, kmin = 0x7fffffff;
        for (int o = t; o < N; o += G) {
            const int k = (int)(o == 0 ? key1 : (unsigned)keys[o]);
            atomicAdd(&hist[k], 1u);
            kmax = max(kmax, k); kmin = min(kmin, k);
        }
        kmax = __reduce_max_sync(0xffffffffu, kmax);
        kmin = __reduce_min_sync(0xffffffffu, kmin);
        if (lane == 0) { atomicMax(&red[3], kmax); atomicMin(&red[2], kmin); }
        __syncthreads();
        const int vmax = red[3];
        kmin = red[2];

        float low_clip = dp.low_clip_db, high_clip = 0.f, dyn = dp.dynamic_range;
        const float fn = (float)kp.n_avg, z3 = (float)(3 * dp.zoom);
        const Divisor dfn = make_divisor(fn);
        auto wfdb = [&](float s) { return ((div_rn<true>(s, dfn) - 255.0f) - 13.0f) + z3; };
        if (dp.auto_scale) {
            const int want = kp.p_lo + 1;
            const int cpt = ((nbins + G - 1) / G) | 1;
            const int b0 = t * cpt, b1 = min(b0 + cpt, nbins);
            int sum = 0, first = 0x7fffffff;
            for (int b = b0; b < b1; ++b) sum += (int)hist[b];
            int incl = sum;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int u = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += u; }
            if (lane == 31) hist[nbins + (t >> 5)] = (unsigned)incl;
            __syncthreads();
            int prefix = incl - sum;
            for (int w = 0; w < (t >> 5); ++w) prefix += (int)hist[nbins + w];
            if (prefix < want && want <= prefix + sum) {
                int c = prefix, b = b0;
                for (; b < b1; ++b) { c += (int)hist[b]; if (c >= want) break; }
                red[0] = b; red[1] = c;
            }
            __syncthreads();
            const int v_lo = red[0], cnt_lo = red[1];
            for (int b = max(b0, v_lo + 1); b < b1; ++b) if (hist[b]) { first = b; break; }     // min{key > v_lo}
            first = __reduce_min_sync(0xffffffffu, first);
            if (lane == 0) atomicMin(&red[4], first);
            __syncthreads();
            const int mn = red[4];
            const int v_hi = (cnt_lo >= want + 1 || mn == 0x7fffffff) ? v_lo : mn;
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
        const float low = low_clip + (float)dp.delta_low_db;
        const float nf = dyn + (float)dp.delta_high_db;
        const float den = nf - (float)dp.delta_low_db;
        if (t == 0) {
            if (dp.auto_scale) { kp.disp[ch].low_clip_db = low_clip; kp.disp[ch].dynamic_range = dyn; }
            if (kp.scalars) {
                ssdr_wf_scalars_t sc;
                sc.low_clip_db = low_clip; sc.high_clip_db = dp.auto_scale ? high_clip : wfdb((float)vmax);
                sc.dynamic_range = dyn; sc.wf_min_db = low - z3; sc.wf_max_db = (low_clip + nf) - z3;
                kp.scalars[ch] = sc;
            }
        }
        const Divisor dden = make_divisor(den);
        const size_t row = (size_t)ch * N;
        for (int o = t; o < N; o += G) {
            const unsigned raw = keys[o];
            const float s = (float)(o == 0 ? key1 : raw);
            const float m = div_rn<true>(s, dfn);
            const float w = ((m - 255.0f) - 13.0f) + z3;
            float c = dden.ok ? div_rn<true>(w - low, dden) : div_rn<false>(w - low, dden);
            c = fminf(fmaxf(c, 0.0f), 1.0f);
            c = c * 254.0f;
            c = fminf(fmaxf(c, 0.0f), 255.0f);
            if (kp.colour) kp.colour[row + o] = c;
            if (kp.pixels) kp.pixels[row + o] = (uint8_t)__float2int_rn(c);
            if (kp.spectrum) kp.spectrum[row + o] = div_rn<true>((float)raw, dfn);
        }
        __syncthreads();                                     // hist / red are reused by the next channel
    }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
// Tensor map of the input seen as [frames x 16 rows][1024 samples] (32-bit words), box = 16 rows x 64 samples: the tile one
// warp of the staged 16384-point kernel pulls per frame.  The driver's encoder is reached through the runtime
// (cudaGetDriverEntryPoint), so the library does not link libcuda.
#ifndef SSDR_TMAP_L2
#define SSDR_TMAP_L2 CU_TENSOR_MAP_L2_PROMOTION_L2_128B      // measured: NONE / 128B / 256B make no difference
#endif
static int make_stage_tmap(CUtensorMap* tm, const void* iq, int fmt, size_t frames, int rows_per_frame, int cols_per_warp, int row_len) {
    typedef CUresult (*encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static encode_fn encode = [] {
        void* f = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) f = nullptr;
        return (encode_fn)f;
    }();
    if (!encode) { set_error("cuTensorMapEncodeTiled is not available from this driver"); return SSDR_E_CUDA; }
    const cuuint32_t words = (fmt == SSDR_IQ_CF32) ? 2u : 1u;             // 32-bit words per sample
    const cuuint64_t gdim[2] = {(cuuint64_t)row_len * words, (cuuint64_t)frames * (cuuint64_t)rows_per_frame};      // row_len = M0 samples
    const cuuint64_t gstride[1] = {(cuuint64_t)row_len * words * 4ull};
    const cuuint32_t box[2] = {(cuuint32_t)cols_per_warp * words, (cuuint32_t)rows_per_frame};      // 16 x 64 (N = 16384) or 8 x 128 (N = 8192) samples: 8 KB as complex64
    const cuuint32_t estr[2] = {1u, 1u};
    const CUresult r = encode(tm, CU_TENSOR_MAP_DATA_TYPE_UINT32, 2, const_cast<void*>(iq), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                              CU_TENSOR_MAP_SWIZZLE_NONE, SSDR_TMAP_L2, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed (%d)", (int)r); return SSDR_E_CUDA; }
    return SSDR_OK;
}

template <int LG>
static int launch_fft(const WfKernelParams& kp, int fmt, int window, cudaStream_t st) {
    using C = Cfg<LG>;
    CUtensorMap tmap;
    memset(&tmap, 0, sizeof(tmap));
    auto launch = [&](auto kern) -> int {
        SSDR_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SM_BYTES));
        int occ = 0;
        SSDR_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, C::THREADS, C::SM_BYTES));
        if (occ < 1) { set_error("waterfall kernel does not fit (smem %zu)", (size_t)C::SM_BYTES); return SSDR_E_CUDA; }
        const int n_groups = (kp.batch + C::FPC - 1) / C::FPC;
        int grid = sm_count() * occ;
        if (grid > n_groups) grid = n_groups;
        kern<<<grid, C::THREADS, C::SM_BYTES, st>>>(kp, tmap);
        count_launch();
        SSDR_CUDA(cudaGetLastError());
        return SSDR_OK;
    };
    if constexpr (C::CAN_STAGE) {
        // local input: the TMA-staged kernel.  Peer (NVLink) input keeps the direct-load kernel -- bulk
        // requests on peer addresses are pathologically slow (section 7).  SSDR_WF_STAGED=0 selects the direct-load kernel
        // (the comparison arm of profiles/).
        static const bool staged_on = [] { const char* e = getenv("SSDR_WF_STAGED"); return !(e && e[0] == '0'); }();
        bool staged = staged_on && kp.prefetch && ((uintptr_t)kp.iq & 15u) == 0;
        if constexpr (!C::STAGE_BULK) {
            // a driver without cuTensorMapEncodeTiled (or an input the encoder refuses) takes the direct-load GPU kernel
            if (staged && make_stage_tmap(&tmap, kp.iq, fmt, (size_t)kp.batch * kp.n_avg, C::R0, (C::G <= 32) ? 32 : 32 * C::NB0, C::M0) != SSDR_OK) staged = false;
        }
        if (staged) {
            if (fmt == SSDR_IQ_CF32) return window ? launch(wf_fft_kernel<LG, SSDR_IQ_CF32, true, true>) : launch(wf_fft_kernel<LG, SSDR_IQ_CF32, false, true>);
            return window ? launch(wf_fft_kernel<LG, SSDR_IQ_S16BE, true, true>) : launch(wf_fft_kernel<LG, SSDR_IQ_S16BE, false, true>);
        }
    }
    if (fmt == SSDR_IQ_CF32) return window ? launch(wf_fft_kernel<LG, SSDR_IQ_CF32, true>) : launch(wf_fft_kernel<LG, SSDR_IQ_CF32, false>);
    return window ? launch(wf_fft_kernel<LG, SSDR_IQ_S16BE, true>) : launch(wf_fft_kernel<LG, SSDR_IQ_S16BE, false>);
}

template <int LG>
static int launch_colorrow(const WfKernelParams& kp, cudaStream_t st) {
    using C = Cfg<LG>;
    const size_t smem = (size_t)C::FPC * C::PADN * sizeof(float) + (size_t)C::FPC * 8 * sizeof(int);
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
    if ((1 << lg) != nfft || lg < 8 || lg > 16) return -1;
    int n = 0;
    if (lg > 14) { radices[n++] = 1 << (lg - 14); lg = 14; }      // front pass of the large-N path
    PlanC p = make_plan(lg);
    for (int i = 0; i < p.np; ++i) radices[n++] = p.r[i];
    return n;
}

// Large-N path selection: the three-kernel path through an HBM scratch (default) or the fused kernel (SSDR_WF_BIG=fused).
// Measured (profiles/r2f_sweep_big_*.jsonl, 32768 points x 10 frames): three kernels 182 Gsamples/s, fused 162 -- the fused
// kernel moves a third of the bytes over HBM but serialises the latency-bound front pass with the sub-transforms on
// each SM (no registers / shared memory left for a producer warp or a staging buffer); interleaving the next frame's
// front pass with the sub-transforms was slower still (131).  The three-kernel path stays the default.
bool wf_big_fused() {
    static const bool fused = [] { const char* e = std::getenv("SSDR_WF_BIG"); return e && !std::strcmp(e, "fused"); }();
    return fused;
}
size_t wf_big_scratch_bytes(int nfft, int n_avg, int channels) {
    if (nfft <= 16384) return 0;
    return wf_big_fused() ? (size_t)sm_count() * nfft * 8 : (size_t)channels * n_avg * nfft * 8;
}

// fused: front pass + sub-transforms in one persistent kernel (scratch stays in L2) -> colour kernel (two launches)
static int launch_big_fused(const WfLaunch& a, WfKernelParams kp, cudaStream_t st) {
    const int rf = a.nfft / 16384;
    using C = Cfg<14>;
    WfBigParams bp;
    bp.kp = kp;
    bp.kp.iq = a.iq; bp.kp.wtab = reinterpret_cast<const float2*>(a.wtab_sub); bp.kp.batch = a.batch; bp.kp.sums = a.sums; bp.kp.rf = rf;
    bp.kp.pixels = nullptr; bp.kp.colour = nullptr; bp.kp.spectrum = nullptr; bp.kp.scalars = nullptr;
    bp.wtab_big = reinterpret_cast<const float2*>(a.wtab); bp.win_big = a.win; bp.scratch = reinterpret_cast<float2*>(a.scratch);
    const int grid = std::min(a.batch, sm_count());
    auto launch = [&](auto kern) -> int {
        SSDR_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SM_BYTES));
        kern<<<grid, 512, C::SM_BYTES, st>>>(bp);
        count_launch();
        SSDR_CUDA(cudaGetLastError());
        return SSDR_OK;
    };
    int rc;
#define SSDR_BIGF(RF, FMT) (a.window ? launch(wf_big_fused_kernel<RF, FMT, true>) : launch(wf_big_fused_kernel<RF, FMT, false>))
    if (rf == 2) rc = (a.iq_format == SSDR_IQ_CF32) ? SSDR_BIGF(2, SSDR_IQ_CF32) : SSDR_BIGF(2, SSDR_IQ_S16BE);
    else rc = (a.iq_format == SSDR_IQ_CF32) ? SSDR_BIGF(4, SSDR_IQ_CF32) : SSDR_BIGF(4, SSDR_IQ_S16BE);
#undef SSDR_BIGF
    if (rc) return rc;
    WfKernelParams k3 = kp;
    k3.sums = a.sums;
    const size_t smem = ((size_t)255 * a.n_avg + 1 + 32 + 8) * sizeof(unsigned);
    SSDR_CUDA(cudaFuncSetAttribute(wf_colour_big_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int g3 = std::min(a.batch, sm_count() * 2);
    wf_colour_big_kernel<<<g3, 1024, smem, st>>>(k3, a.nfft);
    count_launch();
    SSDR_CUDA(cudaGetLastError());
    return SSDR_OK;
}

// N = 32768 / 65536: front pass -> 16384-point kernel in sums mode -> colour kernel (three launches)
static int launch_big(const WfLaunch& a, WfKernelParams kp, cudaStream_t st) {
    const int rf = a.nfft / 16384;
    if (!a.scratch || !a.sums || !a.wtab_sub) { set_error("large-N launch without scratch buffers"); return SSDR_E_STATE; }
    if (wf_big_fused()) return launch_big_fused(a, kp, st);
    const size_t pts = (size_t)a.batch * a.n_avg * 16384;
    int grid = (int)std::min<size_t>((pts + 255) / 256, (size_t)sm_count() * 32);
    float2* scr = reinterpret_cast<float2*>(a.scratch);
    const float2* wt = reinterpret_cast<const float2*>(a.wtab);
#define SSDR_FRONT(RF, FMT, W) wf_front_kernel<RF, FMT, W><<<grid, 256, 0, st>>>(a.iq, scr, wt, a.win, a.batch, a.n_avg)
    if (rf == 2) {
        if (a.iq_format == SSDR_IQ_CF32) { if (a.window) SSDR_FRONT(2, SSDR_IQ_CF32, true); else SSDR_FRONT(2, SSDR_IQ_CF32, false); }
        else { if (a.window) SSDR_FRONT(2, SSDR_IQ_S16BE, true); else SSDR_FRONT(2, SSDR_IQ_S16BE, false); }
    } else {
        if (a.iq_format == SSDR_IQ_CF32) { if (a.window) SSDR_FRONT(4, SSDR_IQ_CF32, true); else SSDR_FRONT(4, SSDR_IQ_CF32, false); }
        else { if (a.window) SSDR_FRONT(4, SSDR_IQ_S16BE, true); else SSDR_FRONT(4, SSDR_IQ_S16BE, false); }
    }
#undef SSDR_FRONT
    count_launch();
    SSDR_CUDA(cudaGetLastError());
    // 16384-point sub-transforms of the rf virtual channels per channel
    WfKernelParams k2 = kp;
    k2.iq = a.scratch; k2.wtab = reinterpret_cast<const float2*>(a.wtab_sub); k2.batch = a.batch * rf; k2.sums = a.sums; k2.rf = rf;
    k2.pixels = nullptr; k2.colour = nullptr; k2.spectrum = nullptr; k2.scalars = nullptr;
    int rc = launch_fft<14>(k2, SSDR_IQ_CF32, 0, st);
    if (rc) return rc;
    // colour row over the N sums of each channel
    WfKernelParams k3 = kp;
    k3.sums = a.sums;
    const size_t smem = ((size_t)255 * a.n_avg + 1 + 32 + 8) * sizeof(unsigned);
    SSDR_CUDA(cudaFuncSetAttribute(wf_colour_big_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int g3 = std::min(a.batch, sm_count() * 2);
    wf_colour_big_kernel<<<g3, 1024, smem, st>>>(k3, a.nfft);
    count_launch();
    SSDR_CUDA(cudaGetLastError());
    return SSDR_OK;
}

int wf_launch(const WfLaunch& a, cudaStream_t st) {
    WfKernelParams kp;
    std::memset(&kp, 0, sizeof(kp));
    kp.iq = a.iq; kp.wtab = reinterpret_cast<const float2*>(a.wtab); kp.win = a.win; kp.thr = a.thr; kp.disp = a.disp;
    kp.pixels = a.pixels; kp.colour = a.colour; kp.spectrum = a.spectrum; kp.scalars = a.scalars;
    kp.lines = a.lines; kp.batch = a.batch; kp.n_avg = a.n_avg; kp.p_lo = a.p_lo; kp.p_gamma = a.p_gamma;
    kp.est_c1 = a.est_c1; kp.est_c0 = a.est_c0; kp.prefetch = a.remote_input ? 0 : 1;
    kp.key_bits = 1;
    if (const char* e = std::getenv("SSDR_WF_STAGGER")) kp.stagger = std::atoi(e);      // developer knob (scripts/exp_stagger.sh)
    while ((1 << kp.key_bits) <= 255 * a.n_avg) ++kp.key_bits;
    const int lg = ilog2(a.nfft);
    if (!a.lines && lg > 14) return launch_big(a, kp, st);
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
