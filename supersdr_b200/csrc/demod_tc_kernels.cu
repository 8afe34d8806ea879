// Fused demodulator, tensor-core engine: the 127-tap FIR of the demodulator as a Toeplitz GEMM on tcgen05 (sm_100a),
// everything else (K6 unpack, NCO mix, detector, AGC, PCM, RSSI) as in demod_kernels.cu -- the back end is the same code
// (demod_post.cuh).  Replaces the remote KiwiSDR SND computation the reference only parametrises
// (utils_supersdr.py:1022-1029) and receives as PCM (utils_supersdr.py:1044-1076).  Spec: DESIGN.md 4.5.
//
// Formulation.  Cut the mixed signal z of a channel into 32-sample blocks.  The 32 FIR outputs of block i are
//     y[32 i + n] = sum_k Z_i[k] T[k][n],   Z_i[k] = z[32 (i - 4) + k],  k < 160,   T[k][n] = h[128 + n - k]  (0 outside 0..126)
// i.e. D[128 x 32] = A[128 x 160] B[160 x 32] with one ROW per (block, re | im) and the taps as a Toeplitz B operand.
// A tile step covers one 512-sample frame of FOUR channels that share a filter (host-side grouping): 4 channels x
// {re, im} x 16 blocks = 128 rows = one M = 128 tile, and warp w reads back exactly its own channel (TMEM lanes 32 w ..).
// A CTA runs TILES = 4 such tiles side by side (4 warps each, own operand buffer, own TMEM columns, own barriers) on one
// shared B operand: the mixer / detector / AGC code is latency-bound, so the SM needs the 16 warps (128 registers each).
//
//   * One copy of the signal serves all five K chunks: K chunk c of row i is block i - 4 + c, i.e. the same array read
//     c rows further up.  The operand is stored as "mini-streams" of 12 rows (4 history blocks + 8 blocks) per 8-row
//     group, 128-byte swizzle applied on absolute address bits, and the descriptor start address moves by c * 128 bytes
//     with SBO = 12 rows (scripts/ubench/tcgen05_shift_probe.cu pins that this is what the hardware computes).
//   * float32 accuracy from TF32 tensor cores: x = hi + lo (cvt.rna.tf32), D = A_hi [B_hi | B_lo] + A_lo B_hi.  The first
//     product is one kind::tf32 MMA per K step of 8 (N = 64; the two column halves are added after the TMEM read-back).
//     The correction A_lo B_hi is ~2^-12 of the result and needs only a few bits: both factors go to the tensor core
//     as bfloat16 (kind::f16, K step of 16, same fp32 accumulator) -- half the operand bytes of a TF32 MMA and a
//     12 KB instead of a 24 KB copy of the rows.  Relative RMS error vs float64 ~1e-6 (tolerance 1e-5).
//     tests/test_tc_split_model.py restates this arithmetic on the CPU (error budget; why bfloat16 and not float16).
//   * No warp waits for another one to hand a frame over: each warp publishes its rows and counts itself in, the LAST
//     of the tile's four warps to arrive issues the tile's 30 MMAs (one thread); completion comes back through
//     tcgen05.commit -> mbarrier.  Two TMEM accumulators per tile: while the MMAs of frame b run, the tile's warps run
//     the back end of frame b - 1, and the other tiles of the CTA mix / detect as well.
//   * Rounds (four quads of one filter) are pulled from a global counter in the host's order: dearest detector first,
//     the last partial wave as narrower rounds (capi.cu: demod_plan_rounds).
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdlib>

#include "common.cuh"
#include "demod_host.h"
#include "demod_post.cuh"

// Developer switches for timing experiments (scripts/exp_tc_variants.sh); the defaults are the measured optimum.
#ifndef SSDR_TC_LASTARRIVER
#define SSDR_TC_LASTARRIVER 1     // 1: the last warp of a tile to arrive issues the MMAs; 0: tile barrier, fixed issuer thread
#endif
#ifndef SSDR_TC_ISSUE_BLOCK
#define SSDR_TC_ISSUE_BLOCK 1     // 1: the frame's 30 MMAs + commit as one asm block (one elect); 0: one asm statement per MMA
#endif
#ifndef SSDR_TC_WAIT_NS
#define SSDR_TC_WAIT_NS 0           // 0: mbarrier.try_wait with a suspend-time hint; > 0: test_wait + __nanosleep(ns) back-off
#endif
#ifndef SSDR_TC_TILES
#define SSDR_TC_TILES 4           // tiles (groups of four warps) per CTA: 16 warps at 128 registers
#endif
#ifndef SSDR_TC_L1_PREFETCH
#define SSDR_TC_L1_PREFETCH 0     // 1: the next frame's IQ lines are pulled into L1 just before the back end of the previous frame, so the
                                  //    mixer's loads after the MMA wait would hit L1.  Measured (round 2): 181 instead of 186 Gsamples/s -- 0.
#endif
#ifndef SSDR_TC_L2_AHEAD
#define SSDR_TC_L2_AHEAD (SSDR_TC_L1_PREFETCH ? 2 : 1)      // frames ahead of the L2 prefetch
#endif
#ifndef SSDR_TC_STAGGER
#define SSDR_TC_STAGGER 0         // developer knob: cycles between the starts of a CTA's tiles within a round (measured: no effect, the tiles are not in lock step)
#endif
#ifndef SSDR_TC_EARLYMIX
#define SSDR_TC_EARLYMIX 0        // 1: mixer arithmetic of frame b + 1 before the wait for the MMAs of frame b (parked in
                                  //    shared memory across the wait); 0: after it.  Measured: 1 is 13 % slower.
#endif

// Developer timeline (scripts/demod_trace.py): -DSSDR_TRACE records clock64() at the phase boundaries of every warp of CTA 0.
#ifdef SSDR_TRACE
#define DTRACE_FRAMES 96
#define DTRACE_PTS 8
__device__ long long g_demod_trace[DTRACE_FRAMES * 16 * DTRACE_PTS];
extern "C" int ssdr_debug_demod_trace(long long* out) {
    cudaMemcpyFromSymbol(out, g_demod_trace, sizeof(long long) * DTRACE_FRAMES * 16 * DTRACE_PTS);
    return 0;
}
__device__ __forceinline__ long long* dtrace_slots() { __shared__ long long s[16 * DTRACE_PTS]; return s; }
#define DTRACE(pt) do { if ((threadIdx.x & 31) == 0) dtrace_slots()[(threadIdx.x >> 5) * DTRACE_PTS + (pt)] = clock64(); } while (0)
#define DTRACE_FLUSH() do { if (blockIdx.x == 0 && (threadIdx.x & 31) == 0 && tr_n < DTRACE_FRAMES) { for (int k_ = 0; k_ < DTRACE_PTS; ++k_) \
    g_demod_trace[(tr_n * 16 + (threadIdx.x >> 5)) * DTRACE_PTS + k_] = dtrace_slots()[(threadIdx.x >> 5) * DTRACE_PTS + k_]; } ++tr_n; } while (0)
#else
#define DTRACE(pt) do { } while (0)
#define DTRACE_FLUSH() do { } while (0)
#endif

namespace ssdr {

namespace {

constexpr int T = SSDR_FIR_TAPS;          // 127
constexpr int H = T - 1;                  // 126 history samples kept per channel
constexpr int FR = SSDR_FRAME;            // 512
constexpr int SPL = kDemodSpl;            // 16
constexpr int WARPS = 4;                  // channels (warps) per tile
constexpr int TILES = SSDR_TC_TILES;      // tiles per CTA
constexpr int KCH = 5;                    // K chunks of 32 samples (160 = 128 history + 32)
constexpr unsigned ROWB = 128;            // one 32-sample block of floats
constexpr unsigned GROWS = 12;            // rows per mini-stream: 4 history blocks + 8 blocks
constexpr unsigned GRPB = GROWS * ROWB;   // 1536 bytes = SBO of the A operand
constexpr unsigned CHB = 4 * GRPB;        // per channel: re octets 0, 1 then im octets 0, 1
constexpr unsigned A_BYTES = WARPS * CHB; // 24576: the hi (TF32) part of the rows, 128-byte swizzle
// the lo part of the rows as bfloat16, no swizzle: 16-byte K chunks (8 samples) in four planes, row r of a plane at r * 16
constexpr unsigned A16_ROWS = 4 * WARPS * GROWS;          // 192 rows per tile
// plane pitch 3104 = 32 bytes (8 banks) mod 128: every shared-memory access to the planes is then conflict-free --
//   * the mixer's 8-byte stores (half-warp = 2 rows x 4 planes x 2 halves): word offsets 8 plane + 4 row + 2 half, 16 distinct;
//   * the 16-byte history copies (quarter-warp = 4 planes x 2 rows): 8 plane + 4 row, 8 distinct 4-word groups.
// (The r1 pitch of 3088 made plane and row collide, 3136 made planes p and p + 2 collide: ncu counted 2 x the ideal
// wavefronts on these instructions, a quarter of the kernel's shared-memory wavefronts -- profiles/r2e_demod_tc_ncu_summary.txt.)
constexpr unsigned A16_LBO = A16_ROWS * 16 + 32;
constexpr unsigned A16_BYTES = 13 * 1024;                 // 4 planes, rounded so that the next tile's hi part stays 1024-aligned
constexpr unsigned TILE_BYTES = A_BYTES + A16_BYTES;
constexpr unsigned B16_LBO = 4 * 128;                     // B_hi as bfloat16, no swizzle: K chunk of 8 -> 32 rows x 16 bytes
constexpr unsigned B16_BYTES = 20 * B16_LBO;              // 10240
constexpr unsigned B_ATOM = 8 * 1024;     // per K chunk: 32 rows of B_hi (4 groups) then 32 rows of B_lo
constexpr unsigned B_BYTES = KCH * B_ATOM;
constexpr unsigned PARK_BYTES = SSDR_TC_EARLYMIX ? 32 * SPL * 8 : 0;      // per warp: the mixer output of the next frame, thread-private
constexpr unsigned SMEM_BYTES = B_BYTES + B16_BYTES + TILES * (TILE_BYTES + WARPS * PARK_BYTES) + 1024;   // + alignment slack
constexpr unsigned TMEM_COLS = TILES > 2 ? 512 : 256;       // two accumulators of 64 columns per tile, allocation is a power of two

__device__ __forceinline__ unsigned swz(unsigned off) { return off ^ (((off >> 7) & 7u) << 4); }   // off from a 1024-aligned base
// x rounded to TF32 (10 mantissa bits), nearest with ties away from zero = cvt.rna.tf32.f32 for every finite x: half an ulp
// added to the magnitude bits, low 13 bits cleared.  ptxas expands the cvt to four instructions (add, |x| >= inf test, select,
// mask) to keep inf / NaN payloads; the operands here are finite (int16-scale IQ times unit-gain taps), so two suffice --
// 64 instructions per lane and frame less on the mixer's store path, bit-identical results.
#ifndef SSDR_TC_CVT_RNA
#define SSDR_TC_CVT_RNA 0
#endif
__device__ __forceinline__ float tf32_hi(float x) {
#if SSDR_TC_CVT_RNA
    unsigned u; asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x)); return __uint_as_float(u);
#else
    return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xffffe000u);
#endif
}

__device__ __forceinline__ uint64_t desc_sw128(unsigned addr, unsigned sbo) {
    uint64_t d = 0;
    d |= (uint64_t)((addr >> 4) & 0x3fff);                   // start address
    d |= (uint64_t)1 << 16;                                   // LBO (unused: swizzled K-major)
    d |= (uint64_t)((sbo >> 4) & 0x3fff) << 32;               // SBO: next 8-row group
    d |= (uint64_t)1 << 46;                                   // descriptor version (sm_100)
    d |= (uint64_t)2 << 61;                                   // SWIZZLE_128B
    return d;
}
__device__ __forceinline__ uint64_t desc_none(unsigned addr, unsigned lbo, unsigned sbo) {     // no swizzle, K-major
    uint64_t d = 0;
    d |= (uint64_t)((addr >> 4) & 0x3fff);
    d |= (uint64_t)((lbo >> 4) & 0x3fff) << 16;               // LBO: next 16-byte K chunk
    d |= (uint64_t)((sbo >> 4) & 0x3fff) << 32;               // SBO: next 8-row group
    d |= (uint64_t)1 << 46;
    return d;
}
// instruction descriptor: D = F32, A = B = TF32 (2) or BF16 (1), both K-major, N >> 3 at bit 17, M >> 4 at bit 24
__device__ __forceinline__ constexpr unsigned idesc_tf32(int m, int n) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((unsigned)(n >> 3) << 17) | ((unsigned)(m >> 4) << 24);
}
__device__ __forceinline__ constexpr unsigned idesc_bf16(int m, int n) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((unsigned)(n >> 3) << 17) | ((unsigned)(m >> 4) << 24);
}
#if !SSDR_TC_ISSUE_BLOCK
// MMAs for a warp-UNIFORM issue section (every lane holds the same operands, one elected lane issues): descriptors
// as (lo, hi) words so that stepping through the operand is one 32-bit add on the low word (start-address field).
__device__ __forceinline__ void mma_tf32_elect(unsigned tmem, unsigned da_lo, unsigned da_hi, unsigned db_lo, unsigned db_hi, unsigned idesc, unsigned accumulate) {
    asm volatile("{\n\t.reg .pred p, q;\n\t.reg .b64 da, db;\n\telect.sync _|p, 0xffffffff;\n\tsetp.ne.b32 q, %6, 0;\n\tmov.b64 da, {%1, %2};\n\tmov.b64 db, {%3, %4};\n\t"
                 "@p tcgen05.mma.cta_group::1.kind::tf32 [%0], da, db, %5, q;\n\t}"
                 ::"r"(tmem), "r"(da_lo), "r"(da_hi), "r"(db_lo), "r"(db_hi), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void mma_f16_elect(unsigned tmem, unsigned da_lo, unsigned da_hi, unsigned db_lo, unsigned db_hi, unsigned idesc, unsigned accumulate) {
    asm volatile("{\n\t.reg .pred p, q;\n\t.reg .b64 da, db;\n\telect.sync _|p, 0xffffffff;\n\tsetp.ne.b32 q, %6, 0;\n\tmov.b64 da, {%1, %2};\n\tmov.b64 db, {%3, %4};\n\t"
                 "@p tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, q;\n\t}"
                 ::"r"(tmem), "r"(da_lo), "r"(da_hi), "r"(db_lo), "r"(db_hi), "r"(idesc), "r"(accumulate) : "memory");
}
#endif
__device__ __forceinline__ unsigned pack_bf16(float a, float b) {
    const __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<const unsigned*>(&v);
}
__device__ __forceinline__ void tmem_ld32(unsigned taddr, float (&v)[32]) {
    unsigned r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
        "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
          "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
          "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// The 30 MMAs + commit of one frame as ONE instruction block (round 2): one elect.sync for all of them, the descriptor
// low words stepped by immediates -- issued as separate asm statements each MMA carried its own ELECT / VOTEU pair and
// re-materialised its constants (~14 instructions per MMA).  The immediates are the operand geometry in 16-byte units:
//   TF32 step (c, j): A + 8 c + 2 j (row shift c x 128 bytes, K step j x 32 bytes), B + 512 c + 2 j (K chunk c x B_ATOM);
//   bf16 step (c, j): A16 + 2 j x A16_LBO / 16 + c, B16 + (4 c + 2 j) x B16_LBO / 16.
static_assert(ROWB == 128 && B_ATOM == 8192 && A16_LBO == 3104 && B16_LBO == 512 && KCH == 5, "immediates of tc_issue_block");
__device__ __forceinline__ void tc_issue_block(unsigned td, unsigned a_lo, unsigned a_hi, unsigned b_lo, unsigned b_hi, unsigned a16_lo,
                                               unsigned a16_hi, unsigned b16_lo, unsigned b16_hi, unsigned i64, unsigned i16, unsigned barp) {
    asm volatile(
        "{\n\t.reg .pred p, acc;\n\t.reg .b64 da, db;\n\t.reg .b32 ta, tb;\n\t"
        "elect.sync _|p, 0xffffffff;\n\t"
        "setp.ne.b32 acc, 0, 0;\n\t"
        "add.u32 ta, %1, 0;\n\tadd.u32 tb, %3, 0;\n\tmov.b64 da, {ta, %2};\n\tmov.b64 db, {tb, %4};\n\t"
        "@p tcgen05.mma.cta_group::1.kind::tf32 [%0], da, db, %9, acc;\n\t"
        "setp.eq.b32 acc, 0, 0;\n\t"
        "add.u32 ta, %1, 2;\n\tadd.u32 tb, %3, 2;\n\tmov.b64 da, {ta, %2};\n\tmov.b64 db, {tb, %4};\n\t"
        "@p tcgen05.mma.cta_group::1.kind::tf32 [%0], da, db, %9, acc;\n\t"
        "add.u32 ta, %1, 4;\n\tadd.u32 tb, %3, 4;\n\tmov.b64 da, {ta, %2};\n\tmov.b64 db, {tb, %4};\n\t"
        "@p tcgen05.mma.cta_group::1.kind::tf32 [%0], da, db, %9, acc;\n\t"
        "add.u32 ta, %1, 6;\n\tadd.u32 tb, %3, 6;\n\tmov.b64 da, {ta, %2};\n\tmov.b64 db, {tb, %4};\n\t"
        "@p tcgen05.mma.cta_group::1.kind::tf32 [%0], da, db, %9, acc;\n\t"
        "add.u32 ta, %5, 0;\n\tadd.u32 tb, %7, 0;\n\tmov.b64 da, {ta, %6};\n\tmov.b64 db, {tb, %8};\n\t"
        "@p tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %10, acc;\n\t"
        "add.u32 ta, %5, 388;\n\tadd.u32 tb, %7, 64;\n\tmov.b64 da, {ta, %6};\n\tmov.b64 db, {tb, %8};\n\t"
        "@p tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %10, acc;\n\t"
        "add.u32 ta, %1, 8;\n\tadd.u32 tb, %3, 512;\n\tmov.b64 da, {ta, %2};\n\tmov.b64 db, {tb, %4};\n\t"
        "@p tcgen05.mma.cta_group::1.kind::tf32 [%0], da, db, %9, acc;\n\t"
        "add.u32 ta, %1, 10;\n\tadd.u32 tb, %3, 514;\n\tmov.b64 da, {ta, %2};\n\tmov.b64 db, {tb, %4};\n\t"
        "@p tcgen05.mma.cta_group::1.kind::tf32 [%0], da, db, %9, acc;\n\t"
        "add.u32 ta, %1, 12;\n\tadd.u32 tb, %3, 516;\n\tmov.b64 da, {ta, %2};\n\tmov.b64 db, {tb, %4};\n\t"
        "@p tcgen05.mma.cta_group::1.kind::tf32 [%0], da, db, %9, acc;\n\t"
        "add.u32 ta, %1, 14;\n\tadd.u32 tb, %3, 518;\n\tmov.b64 da, {ta, %2};\n\tmov.b64 db, {tb, %4};\n\t"
        "@p tcgen05.mma.cta_group::1.kind::tf32 [%0], da, db, %9, acc;\n\t"
        "add.u32 ta, %5, 1;\n\tadd.u32 tb, %7, 128;\n\tmov.b64 da, {ta, %6};\n\tmov.b64 db, {tb, %8};\n\t"
        "@p tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %10, acc;\n\t"
        "add.u32 ta, %5, 389;\n\tadd.u32 tb, %7, 192;\n\tmov.b64 da, {ta, %6};\n\tmov.b64 db, {tb, %8};\n\t"
        "@p tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %10, acc;\n\t"
        "add.u32 ta, %1, 16;\n\tadd.u32 tb, %3, 1024;\n\tmov.b64 da, {ta, %2};\n\tmov.b64 db, {tb, %4};\n\t"
        "@p tcgen05.mma.cta_group::1.kind::tf32 [%0], da, db, %9, acc;\n\t"
        "add.u32 ta, %1, 18;\n\tadd.u32 tb, %3, 1026;\n\tmov.b64 da, {ta, %2};\n\tmov.b64 db, {tb, %4};\n\t"
        "@p tcgen05.mma.cta_group::1.kind::tf32 [%0], da, db, %9, acc;\n\t"
        "add.u32 ta, %1, 20;\n\tadd.u32 tb, %3, 1028;\n\tmov.b64 da, {ta, %2};\n\tmov.b64 db, {tb, %4};\n\t"
        "@p tcgen05.mma.cta_group::1.kind::tf32 [%0], da, db, %9, acc;\n\t"
        "add.u32 ta, %1, 22;\n\tadd.u32 tb, %3, 1030;\n\tmov.b64 da, {ta, %2};\n\tmov.b64 db, {tb, %4};\n\t"
        "@p tcgen05.mma.cta_group::1.kind::tf32 [%0], da, db, %9, acc;\n\t"
        "add.u32 ta, %5, 2;\n\tadd.u32 tb, %7, 256;\n\tmov.b64 da, {ta, %6};\n\tmov.b64 db, {tb, %8};\n\t"
        "@p tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %10, acc;\n\t"
        "add.u32 ta, %5, 390;\n\tadd.u32 tb, %7, 320;\n\tmov.b64 da, {ta, %6};\n\tmov.b64 db, {tb, %8};\n\t"
        "@p tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %10, acc;\n\t"
        "add.u32 ta, %1, 24;\n\tadd.u32 tb, %3, 1536;\n\tmov.b64 da, {ta, %2};\n\tmov.b64 db, {tb, %4};\n\t"
        "@p tcgen05.mma.cta_group::1.kind::tf32 [%0], da, db, %9, acc;\n\t"
        "add.u32 ta, %1, 26;\n\tadd.u32 tb, %3, 1538;\n\tmov.b64 da, {ta, %2};\n\tmov.b64 db, {tb, %4};\n\t"
        "@p tcgen05.mma.cta_group::1.kind::tf32 [%0], da, db, %9, acc;\n\t"
        "add.u32 ta, %1, 28;\n\tadd.u32 tb, %3, 1540;\n\tmov.b64 da, {ta, %2};\n\tmov.b64 db, {tb, %4};\n\t"
        "@p tcgen05.mma.cta_group::1.kind::tf32 [%0], da, db, %9, acc;\n\t"
        "add.u32 ta, %1, 30;\n\tadd.u32 tb, %3, 1542;\n\tmov.b64 da, {ta, %2};\n\tmov.b64 db, {tb, %4};\n\t"
        "@p tcgen05.mma.cta_group::1.kind::tf32 [%0], da, db, %9, acc;\n\t"
        "add.u32 ta, %5, 3;\n\tadd.u32 tb, %7, 384;\n\tmov.b64 da, {ta, %6};\n\tmov.b64 db, {tb, %8};\n\t"
        "@p tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %10, acc;\n\t"
        "add.u32 ta, %5, 391;\n\tadd.u32 tb, %7, 448;\n\tmov.b64 da, {ta, %6};\n\tmov.b64 db, {tb, %8};\n\t"
        "@p tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %10, acc;\n\t"
        "add.u32 ta, %1, 32;\n\tadd.u32 tb, %3, 2048;\n\tmov.b64 da, {ta, %2};\n\tmov.b64 db, {tb, %4};\n\t"
        "@p tcgen05.mma.cta_group::1.kind::tf32 [%0], da, db, %9, acc;\n\t"
        "add.u32 ta, %1, 34;\n\tadd.u32 tb, %3, 2050;\n\tmov.b64 da, {ta, %2};\n\tmov.b64 db, {tb, %4};\n\t"
        "@p tcgen05.mma.cta_group::1.kind::tf32 [%0], da, db, %9, acc;\n\t"
        "add.u32 ta, %1, 36;\n\tadd.u32 tb, %3, 2052;\n\tmov.b64 da, {ta, %2};\n\tmov.b64 db, {tb, %4};\n\t"
        "@p tcgen05.mma.cta_group::1.kind::tf32 [%0], da, db, %9, acc;\n\t"
        "add.u32 ta, %1, 38;\n\tadd.u32 tb, %3, 2054;\n\tmov.b64 da, {ta, %2};\n\tmov.b64 db, {tb, %4};\n\t"
        "@p tcgen05.mma.cta_group::1.kind::tf32 [%0], da, db, %9, acc;\n\t"
        "add.u32 ta, %5, 4;\n\tadd.u32 tb, %7, 512;\n\tmov.b64 da, {ta, %6};\n\tmov.b64 db, {tb, %8};\n\t"
        "@p tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %10, acc;\n\t"
        "add.u32 ta, %5, 392;\n\tadd.u32 tb, %7, 576;\n\tmov.b64 da, {ta, %6};\n\tmov.b64 db, {tb, %8};\n\t"
        "@p tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %10, acc;\n\t"
        "@p tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%11];\n\t}"
        ::"r"(td), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(a16_lo), "r"(a16_hi), "r"(b16_lo), "r"(b16_hi), "r"(i64), "r"(i16), "r"(barp)
        : "memory");
}

struct alignas(8) TcShared {
    unsigned long long bar[TILES];     // MMAs of the tile's current frame have completed (tcgen05.commit)
    unsigned arrived[TILES];           // warps of the tile whose operand rows of the next frame are in place (running count)
    unsigned tmem_base;
    int round;                         // the round this CTA is working on (dynamic scheduling)
};

template <int FMT>
__global__ void __launch_bounds__(TILES * WARPS * 32, 1)
demod_tc_kernel(const DemodKernelParams kp, const int4* __restrict__ quad_ch, const int* __restrict__ quad_fid, int n_rounds,
                int* __restrict__ round_ctr, int tile_stagger) {
    extern __shared__ unsigned char smem_raw[];
    __shared__ TcShared sh;
    const unsigned raw = (unsigned)__cvta_generic_to_shared(smem_raw);
    unsigned char* base = smem_raw + (((raw + 1023u) & ~1023u) - raw);          // 1024-aligned: swizzle phase = offset bits
    const int tid = threadIdx.x, lane = tid & 31, tile = tid >> 7, warp = (tid >> 5) & 3;     // warp = TMEM lane quarter
    unsigned char* sB = base;
    unsigned char* sB16 = base + B_BYTES;
    unsigned char* sAh = base + B_BYTES + B16_BYTES + (unsigned)tile * TILE_BYTES;
    unsigned char* sA16 = sAh + A_BYTES;
#if SSDR_TC_EARLYMIX
    float2* park = reinterpret_cast<float2*>(base + B_BYTES + B16_BYTES + TILES * TILE_BYTES + (unsigned)(tid >> 5) * PARK_BYTES) + lane;
#endif
    const unsigned aB = (unsigned)__cvta_generic_to_shared(sB), aB16 = aB + B_BYTES, aAh = aB16 + B16_BYTES + (unsigned)tile * TILE_BYTES,
                   aA16 = aAh + A_BYTES;
    const unsigned barp = (unsigned)__cvta_generic_to_shared(&sh.bar[tile]);
    const bool issuer = (tid & 127) == 0;                   // initialises the tile's barrier

    if (issuer) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(barp));
        sh.arrived[tile] = 0u;
    }
    if (tid < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(&sh.tmem_base)), "n"(TMEM_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const unsigned tm = sh.tmem_base + (unsigned)tile * 128u;
    unsigned phase = 0;
#ifdef SSDR_TRACE
    int tr_n = 0;
#endif

    const int nblk = kp.n_samples / FR;
    const unsigned wch = (unsigned)warp * CHB;              // this warp's channel slot in the array of 128-byte rows
    int cur_fid = -1;

    // off = byte offset of a sample in the (unswizzled) array of 128-byte rows; its bfloat16 twin lives in plane
    // (sample >> 3) at row * 16 + (sample & 7) * 2
    auto a16 = [](unsigned off) -> unsigned { return ((off >> 5) & 3u) * A16_LBO + (off >> 7) * 16u + ((off >> 2) & 7u) * 2u; };
    auto put = [&](unsigned off, float x) {                 // x -> hi (TF32) into the swizzled rows, lo (bfloat16) into the planes
        const float hi = tf32_hi(x);
        *reinterpret_cast<float*>(sAh + swz(off)) = hi;
        *reinterpret_cast<__nv_bfloat16*>(sA16 + a16(off)) = __float2bfloat16_rn(x - hi);      // x - hi is exact
    };

    // Rounds differ in cost (detector / AGC variant), so CTAs pull them from a counter in the host's order (dearest first)
    for (;;) {
        if (tid == 0) sh.round = atomicAdd(round_ctr, 1);
        __syncthreads();                                    // the barrier that ends a round protects sh.round
        const int rd = sh.round;
        if (rd >= n_rounds) break;
        // ---- the round's filter as a Toeplitz B operand (rebuilt only when the filter changes) ---------------------
        const int fid = quad_fid[rd * TILES];
        if (fid != cur_fid) {                               // CTA-uniform; every tile has finished the previous round
            cur_fid = fid;
            const float* taps = kp.taps + (size_t)quad_ch[rd * TILES].x * T;
            for (int e = tid; e < 32 * 160; e += TILES * WARPS * 32) {
                const int n = e / 160, k = e - n * 160, tau = 128 + n - k;
                const float x = (tau >= 0 && tau < T) ? __ldg(taps + tau) : 0.f, hi = tf32_hi(x);
                const int c = k >> 5, kk = k & 31;
                const unsigned off = (unsigned)c * B_ATOM + (unsigned)(n >> 3) * 1024u + (unsigned)(n & 7) * 128u +
                                     (unsigned)(((kk >> 2) ^ (n & 7)) << 4) + (unsigned)(kk & 3) * 4u;
                *reinterpret_cast<float*>(sB + off) = hi;
                *reinterpret_cast<float*>(sB + off + 4096u) = x - hi;
                *reinterpret_cast<__nv_bfloat16*>(sB16 + (unsigned)(k >> 3) * B16_LBO + (unsigned)(n >> 3) * 128u + (unsigned)(n & 7) * 16u +
                                                  (unsigned)(k & 7) * 2u) = __float2bfloat16_rn(x);
            }
            __syncthreads();                                // the per-step fence below publishes B to the tensor core
        }
        const int4 q4 = quad_ch[rd * TILES + tile];
        const int ch = (warp == 0) ? q4.x : (warp == 1) ? q4.y : (warp == 2) ? q4.z : q4.w;
        const bool active = ch >= 0;
        const bool tile_active = q4.x >= 0;                 // slot 0 of a quad is filled first
        // ---- per-channel state; FIR history = blocks -4 .. -1 of the re / im mini-streams of octet 0 ------------
        DemodChan cp = {};
        DemodState* stp = nullptr;
        DemodRegs st = {};
        if (active) {
            cp = kp.chan[ch];
            stp = kp.state + ch;
            demod_regs_load(st, stp, lane);
#pragma unroll
            for (int row = 0; row < 4; ++row) {
                const int idx = 32 * row + lane - 2;        // sample -128 + 32 row + lane; the 126 kept samples start at -126
                const float2 z = (idx >= 0) ? kp.hist[(size_t)ch * H + idx] : make_float2(0.f, 0.f);
                const unsigned off = wch + (unsigned)row * ROWB + (unsigned)lane * 4u;
                put(off, z.x);
                put(off + 2 * GRPB, z.y);
            }
        }
        // ---- frame pipeline of a warp: rows of frame b in place -> (last warp of the tile issues the MMAs of frame b) ->
        // back end of frame b - 1 (its accumulator is the other TMEM buffer) and the mixer arithmetic of frame b + 1 while
        // the MMAs run -> wait for them -> history rows -> store the rows of frame b + 1
        unsigned ph1_mix = st.ph1;                          // the back end advances st.ph1 one frame later than the mixer
        // mixer, part 1: every lane takes four consecutive samples of each quarter frame (coalesced 16 / 32-byte loads):
        // y[4 q + i] = sample 128 q + 4 lane + i.  One NCO evaluation per four samples, the other three by rotation.
        float2 r1;
        nco(cp.inc1, r1.x, r1.y);
        auto mix_compute = [&](int b, float2 (&y)[SPL]) {
            const size_t s0 = (size_t)ch * kp.pitch + (size_t)b * FR;
            float2 xin[SPL];
#pragma unroll
            for (int q = 0; q < 4; ++q) demod_ld_iq4<FMT>(kp.iq, s0 + 128 * q + 4 * lane, &xin[4 * q]);
            if (b + SSDR_TC_L2_AHEAD < nblk) {              // a later frame -> L2 (one 128-byte line per lane)
                const char* nx = static_cast<const char*>(kp.iq) + (s0 + SSDR_TC_L2_AHEAD * FR) * (FMT == SSDR_IQ_CF32 ? 8 : 4) + lane * 128;
                if (FMT == SSDR_IQ_CF32 || lane < 16) asm volatile("prefetch.global.L2 [%0];" ::"l"(nx));
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                float2 w;
                nco(ph1_mix + (unsigned)(128 * q + 4 * lane) * cp.inc1, w.x, w.y);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    y[4 * q + i] = cmulc2(xin[4 * q + i], w);                              // x exp(-j theta)
                    if (i < 3) w = cmul2(w, r1);
                }
            }
            ph1_mix += (unsigned)FR * cp.inc1;
        };
        // mixer, part 2: (hi, lo) split into the operand rows -- only after the MMAs that read the previous frame are done.
        // Four samples = one 16-byte chunk of a row: block 4 q + (lane >> 3), chunk lane & 7.
        auto put4 = [&](unsigned off, float a0, float a1, float a2, float a3) {
            const float h0 = tf32_hi(a0), h1 = tf32_hi(a1), h2 = tf32_hi(a2), h3 = tf32_hi(a3);
            *reinterpret_cast<float4*>(sAh + swz(off)) = make_float4(h0, h1, h2, h3);
            *reinterpret_cast<uint2*>(sA16 + a16(off)) = make_uint2(pack_bf16(a0 - h0, a1 - h1), pack_bf16(a2 - h2, a3 - h3));
        };
        auto mix_store = [&](const float2 (&y)[SPL]) {
            const unsigned row = (unsigned)lane >> 3, chunk = ((unsigned)lane & 7u) * 16u;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const unsigned off = wch + (unsigned)(q >> 1) * GRPB + (4u + 4u * (q & 1) + row) * ROWB + chunk;
                put4(off, y[4 * q].x, y[4 * q + 1].x, y[4 * q + 2].x, y[4 * q + 3].x);
                put4(off + 2 * GRPB, y[4 * q].y, y[4 * q + 1].y, y[4 * q + 2].y, y[4 * q + 3].y);
                if (q == 1) {                               // blocks 4..7 are also the history of octet 1
                    const unsigned offh = wch + GRPB + row * ROWB + chunk;
                    put4(offh, y[4].x, y[5].x, y[6].x, y[7].x);
                    put4(offh + 2 * GRPB, y[4].y, y[5].y, y[6].y, y[7].y);
                }
            }
        };
        auto back_end = [&](int b) {
            // read back: lane i < 16 has the real parts of block i, lane 16 + i its imaginary parts; every warp of the tile
            // takes part in the (warp-collective) TMEM loads
            float v[32];
            {
                float w[32];
                const unsigned taddr = tm + (unsigned)(b & 1) * 64u + ((unsigned)(warp * 32) << 16);
                tmem_ld32(taddr, v);
                tmem_ld32(taddr + 32u, w);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                for (int i = 0; i < 32; i += 2) {           // consecutive columns = consecutive registers: packed adds
                    const float2 t = __fadd2_rn(make_float2(v[i], v[i + 1]), make_float2(w[i], w[i + 1]));
                    v[i] = t.x; v[i + 1] = t.y;
                }
            }
            if (!active) return;
            // pair exchange: lanes p and p ^ 16 swap the halves they do not keep
            float2 acc[SPL];
            const bool lo16 = lane < 16;
#pragma unroll
            for (int i = 0; i < SPL; ++i) {
                const float send = lo16 ? v[16 + i] : v[i];
                const float recv = __shfl_xor_sync(0xffffffffu, send, 16);
                acc[i] = lo16 ? make_float2(v[i], recv) : make_float2(recv, v[16 + i]);
            }
            demod_frame_tail<LanesPaired>(acc, cp, kp, ch, b, (size_t)ch * kp.pitch + (size_t)b * FR, st);
        };
        float2 y[SPL];
        if (tile_active && active) {
            mix_compute(0, y);
            mix_store(y);
#if SSDR_TC_EARLYMIX
#pragma unroll
            for (int i = 12; i < SPL; ++i) park[32 * i] = y[i];       // the slots always hold the newest frame's tail (history)
#endif
        }
        // Tile stagger (round 2, found with the phase timeline scripts/demod_trace.py): the four tiles of a CTA leave the round
        // barrier together and stay in lock step, so their MMA batches reach the tensor pipe at the same time; the pipe
        // time-slices them, every batch takes four times as long (issue -> commit 6.7 k cycles instead of 1.4 k) and each warp
        // waits 2.7 k cycles per frame behind its back end.  Starting tile k of a round k x tile_stagger cycles late spreads
        // the batches over the frame period: each tile then has the tensor pipe to itself.  Paid once per round.
        if (tile && tile_stagger > 0) { const long long c0 = clock64(); while (clock64() - c0 < (long long)tile * tile_stagger) { } }
        for (int b = 0; tile_active && b < nblk; ++b) {
            DTRACE(0);
            // This warp's rows of frame b are in place: publish them to the tensor core (async proxy), order this warp's
            // earlier TMEM reads before the MMAs, and count the warp in.  No warp waits for another one here: the LAST
            // of the tile's four warps to arrive issues the tile's MMAs.
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
#if SSDR_TC_LASTARRIVER
            __syncwarp();
            unsigned old = 0u;
            if (lane == 0)
                asm volatile("atom.acq_rel.cta.shared::cta.add.u32 %0, [%1], 1;" : "=r"(old) : "r"((unsigned)__cvta_generic_to_shared(&sh.arrived[tile])) : "memory");
            // The issue section is warp-uniform (round 2): the whole warp of the last arriver enters it, every lane computes the
            // same descriptor words with plain 32-bit adds and ONE elected lane issues each MMA.  The per-thread form (inside
            // `if (lane == 0)`) cost ~21 instructions per MMA -- ptxas wraps every tcgen05.mma of a divergent region in an
            // ELECT / 5 x R2UR.BROADCAST / branch loop -- 3.2 k cycles per frame on the critical path of the tile's slowest warp
            // (phase timeline, scripts/demod_trace.py).
            if (__shfl_sync(0xffffffffu, (int)((old & 3u) == 3u), 0)) {
                {
#else
            asm volatile("bar.sync %0, 128;" ::"r"(tile + 1) : "memory");        // the tile's four warps
            {
                if (warp == 0) {
#endif
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    constexpr unsigned i64 = idesc_tf32(128, 64), i16 = idesc_bf16(128, 32);
                    const unsigned td = tm + (unsigned)(b & 1) * 64u;
                    const uint64_t dA = desc_sw128(aAh, GRPB), dB = desc_sw128(aB, 1024u);
                    const uint64_t dA16 = desc_none(aA16, A16_LBO, GROWS * 16u), dB16 = desc_none(aB16, B16_LBO, 128u);
                    const unsigned a_lo = (unsigned)dA, a_hi = (unsigned)(dA >> 32), b_lo = (unsigned)dB, b_hi = (unsigned)(dB >> 32);
                    const unsigned a16_lo = (unsigned)dA16, a16_hi = (unsigned)(dA16 >> 32), b16_lo = (unsigned)dB16, b16_hi = (unsigned)(dB16 >> 32);
#if SSDR_TC_ISSUE_BLOCK
                    tc_issue_block(td, a_lo, a_hi, b_lo, b_hi, a16_lo, a16_hi, b16_lo, b16_hi, i64, i16, barp);
#else
#pragma unroll
                    for (int c = 0; c < KCH; ++c) {
#pragma unroll
                        for (int j = 0; j < 4; ++j) {       // TF32 K step of 8 samples = 32 bytes inside the swizzled row
                            const unsigned oa = (unsigned)c * ROWB + (unsigned)j * 32u, ob = (unsigned)c * B_ATOM + (unsigned)j * 32u;
                            mma_tf32_elect(td, a_lo + (oa >> 4), a_hi, b_lo + (ob >> 4), b_hi, i64, (c | j) != 0);     // A_hi [B_hi | B_lo] -> columns 0..63
                        }
#pragma unroll
                        for (int j = 0; j < 2; ++j) {       // bfloat16 K step of 16 samples = two planes; rows shifted by c like the hi part
                            mma_f16_elect(td, a16_lo + (((unsigned)(2 * j) * A16_LBO + (unsigned)c * 16u) >> 4), a16_hi,
                                          b16_lo + (((unsigned)(4 * c + 2 * j) * B16_LBO) >> 4), b16_hi, i16, 1u);               // A_lo B_hi -> columns 0..31
                        }
                    }
                    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\t@p tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}" ::"r"(barp) : "memory");
#endif
                }
            }
            __syncwarp();
#if SSDR_TC_L1_PREFETCH
            if (active && b + 1 < nblk) {                   // frame b + 1 -> L1 while the back end and the MMAs run
                const char* nx = static_cast<const char*>(kp.iq) + ((size_t)ch * kp.pitch + (size_t)(b + 1) * FR) * (FMT == SSDR_IQ_CF32 ? 8 : 4) + lane * 128;
                if (FMT == SSDR_IQ_CF32 || lane < 16) asm volatile("prefetch.global.L1 [%0];" ::"l"(nx));
            }
#endif
            DTRACE(1);
            if (b > 0) back_end(b - 1);                     // overlaps the MMAs of frame b
            DTRACE(2);
#if SSDR_TC_EARLYMIX
            if (active && b + 1 < nblk) {                   // ... and so does the mixer of frame b + 1; its output waits in shared
                mix_compute(b + 1, y);                      // memory (thread-private slots), not in registers
#pragma unroll
                for (int i = 0; i < SPL; ++i) park[32 * i] = y[i];
            }
#endif
            // suspend-time hint: the warp sleeps in hardware until the commit arrives (or 2 us pass) instead of spinning
            // through try_wait / yield / branch -- a fifth of the kernel's issued instructions were this loop
#if SSDR_TC_WAIT_NS == 0
            asm volatile(
                "{\n\t.reg .pred p;\n\t"
                "WAIT_%=:\n\t"
                "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n\t"
                "@!p bra WAIT_%=;\n\t}" ::"r"(barp), "r"(phase), "r"(2000u) : "memory");
#else
            // plain poll with a fixed back-off: every mbarrier poll is a shared-memory transaction, and the suspend form above
            // wakes ~38 times per frame (ncu: 78 SYNCS per warp and frame) on a shared-memory pipe that is 85 % busy
            for (;;) {
                unsigned done;
                asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                             : "=r"(done) : "r"(barp), "r"(phase) : "memory");
                if (done) break;
                __nanosleep(SSDR_TC_WAIT_NS);
            }
#endif
            phase ^= 1u;
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            DTRACE(3);
            if (active) {
                // the frame's last four blocks become the history of the next frame (the MMAs of frame b have completed)
                {
                    const unsigned row = (unsigned)lane >> 3, chunk = (unsigned)lane & 7u;
#pragma unroll
                    for (int part = 0; part < 2; ++part) {
                        const unsigned src = swz(wch + (unsigned)part * 2 * GRPB + GRPB + (8 + row) * ROWB + chunk * 16u);
                        const unsigned dst = swz(wch + (unsigned)part * 2 * GRPB + row * ROWB + chunk * 16u);
                        *reinterpret_cast<float4*>(sAh + dst) = *reinterpret_cast<const float4*>(sAh + src);
                    }
                }
                {                                           // bfloat16 planes: 2 parts x 4 rows x 4 planes = one 16-byte chunk per lane
                    const unsigned part = (unsigned)lane >> 4, row = ((unsigned)lane >> 2) & 3u, plane = (unsigned)lane & 3u;
                    const unsigned g0 = (unsigned)warp * 4u + part * 2u;
                    *reinterpret_cast<uint4*>(sA16 + plane * A16_LBO + (g0 * GROWS + row) * 16u) =
                        *reinterpret_cast<const uint4*>(sA16 + plane * A16_LBO + ((g0 + 1u) * GROWS + 8u + row) * 16u);
                }
                __syncwarp();
                DTRACE(4);
                if (b + 1 < nblk) {
#if SSDR_TC_EARLYMIX
#pragma unroll
                    for (int i = 0; i < SPL; ++i) y[i] = park[32 * i];
#else
                    mix_compute(b + 1, y);
#endif
                    DTRACE(5);
                    mix_store(y);
                }
            }
            DTRACE(6);
            DTRACE_FLUSH();
        }
        if (tile_active) back_end(nblk - 1);
        if (active) {
            // ---- store per-channel state: the last 126 mixed samples, exact, from the mixer output of the last frame
            // (y[12 + i] = sample 384 + 4 lane + i)
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int idx = 4 * lane + i - 2;
#if SSDR_TC_EARLYMIX
                if (idx >= 0) kp.hist[(size_t)ch * H + idx] = park[32 * (12 + i)];
#else
                if (idx >= 0) kp.hist[(size_t)ch * H + idx] = y[12 + i];
#endif
            }
            demod_regs_store(st, stp, lane);
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();                                    // tiles re-converge: B may change, TMEM reads are complete
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    }
    if (tid < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(sh.tmem_base), "n"(TMEM_COLS));
}

}  // namespace

int demod_tc_tiles() { return TILES; }

int demod_tc_launch(const DemodLaunch& a, const int4* quad_ch, const int* quad_fid, int n_rounds, int* round_ctr, cudaStream_t st) {
    DemodKernelParams kp;
    kp.iq = a.iq; kp.chan = a.chan; kp.state = a.state; kp.hist = a.hist; kp.taps = a.taps;
    kp.pcm_f32 = a.pcm_f32; kp.pcm_i16 = a.pcm_i16; kp.rssi = a.rssi;
    kp.batch = a.batch; kp.n_samples = a.n_samples; kp.pitch = a.pitch ? a.pitch : a.n_samples;
    for (int s = 0; s < 5; ++s) kp.am_pow16[s] = a.am_pow16[s];
    auto kern = (a.iq_format == SSDR_IQ_CF32) ? demod_tc_kernel<SSDR_IQ_CF32> : demod_tc_kernel<SSDR_IQ_S16BE>;
    SSDR_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES));
    int grid = sm_count();                                  // one CTA per SM (shared memory), TILES x 4 warps
    if (grid > n_rounds) grid = n_rounds;
    if (grid < 1) return SSDR_OK;
    SSDR_CUDA(cudaMemsetAsync(round_ctr, 0, sizeof(int), st));
    static const int tile_stagger = [] { const char* e = getenv("SSDR_TC_STAGGER"); return e ? atoi(e) : SSDR_TC_STAGGER; }();   // developer knob
    kern<<<grid, TILES * WARPS * 32, SMEM_BYTES, st>>>(kp, quad_ch, quad_fid, n_rounds, round_ctr, tile_stagger);
    count_launch();
    SSDR_CUDA(cudaGetLastError());
    return SSDR_OK;
}

}  // namespace ssdr
