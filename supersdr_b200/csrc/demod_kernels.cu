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
#include <cstdlib>

#include "common.cuh"
#include "demod_host.h"
#include "demod_post.cuh"

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

// Work distribution (round 2).  One warp runs one channel at a time, and channels are pulled from a global counter: with
// static striding 4096 channels on 1776 resident warps left the SMs that hold the low-numbered CTAs with 3 channels per warp
// and the others with 2 (ncu: 9.4 of 12 warps active); pulled dynamically every SM runs until the work is gone (83 -> 98
// Gsamples/s on config 3's shape).  A task can also be a (time slice, channel) pair in slice-major order (`n_slices` > 1): a
// slice's state travels through global memory exactly as it does from call to call (bit-identical output), ordered by a
// per-channel progress word -- the writer fences and its lane 0 releases `slice + 1`, the reader's lane 0 acquires it; the
// task that carries a channel's previous slice was handed out `batch` tasks earlier to a running warp, so the (bounded)
// wait cannot deadlock.  Measured slower than whole channels (see demod_launch), so it is a developer knob.
template <int FMT>
__global__ void __launch_bounds__(WARPS * 32, 3)
demod_kernel(const DemodKernelParams kp, int n_slices, int slice_frames, int* sched) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    WarpSmem* ws = reinterpret_cast<WarpSmem*>(smem_raw) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    const int nblk_all = kp.n_samples / FR;
    const int n_tasks = kp.batch * n_slices;
    int* const progress = sched + 1;

    for (;;) {
        int task = 0;
        if (lane == 0) task = atomicAdd(sched, 1);
        task = __shfl_sync(0xffffffffu, task, 0);
        if (task >= n_tasks) break;
        const int slice = task / kp.batch, ch = task - slice * kp.batch;
        const int b0 = slice * slice_frames, nblk = min(nblk_all, b0 + slice_frames);
        if (slice > 0) {
            if (lane == 0) {
                const long long t0 = clock64();
                int got;
                do {
                    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(got) : "l"(progress + ch) : "memory");
                } while (got < slice && clock64() - t0 < 4000000000ll);
            }
            __syncwarp();
        }
        const DemodChan cp = kp.chan[ch];
        DemodState* stp = kp.state + ch;
        // ---- load per-channel state ---------------------------------------------------------
        DemodRegs st;
        demod_regs_load(st, stp, lane);
        __syncwarp();
        for (int i = lane; i < H; i += 32) ws->z[zpos(1 + i)] = __ldcg(kp.hist + (size_t)ch * H + i);
        if (lane == 0) ws->z[0] = make_float2(0.f, 0.f);
        for (int i = lane; i < TAP_PAD; i += 32) { float h = (i < T) ? kp.taps[(size_t)ch * T + i] : 0.f; ws->taps2[i] = make_float2(h, h); }
        __syncwarp();

        for (int b = b0; b < nblk; ++b) {
            const size_t s0 = (size_t)ch * kp.pitch + (size_t)b * FR;
            // ---- mixer: lane-strided, coalesced; all sixteen loads of the frame in flight together ----------
            float2 xin[SPL];
#pragma unroll
            for (int r = 0; r < SPL; ++r) xin[r] = demod_ld_iq<FMT>(kp.iq, s0 + lane + 32 * r);
#pragma unroll
            for (int r = 0; r < SPL; ++r) {
                const int k = lane + 32 * r;
                const float2 x = xin[r];
                float c, s;
                nco(st.ph1 + (unsigned)k * cp.inc1, c, s);
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
            // ---- magnitude / RSSI, detector, AGC, outputs, per-frame state ------------------------------
            demod_frame_tail<LanesNatural>(acc, cp, kp, ch, b, s0, st);
            __syncwarp();
        }
        // ---- store per-channel state ---------------------------------------------------------------
        for (int i = lane; i < H; i += 32) kp.hist[(size_t)ch * H + i] = ws->z[zpos(1 + i)];
        demod_regs_store(st, stp, lane);
        if (n_slices > 1) {
            __threadfence();
            __syncwarp();
            if (lane == 0) asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(progress + ch), "r"(slice + 1) : "memory");
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
    // Tasks = whole channels by default (slices = 1).  Measured on a B200 (gpurun_out/ffma_slices_a7.txt, 4096 USB channels x 64
    // frames): static channel striding 83, dynamic tasks 98.3 Gsamples/s; 2 / 4 / 8 time slices per channel 77.6 / 74.3 / 75.5 --
    // bit-identical, better balanced on paper and slower in fact, so slicing stays a developer knob (SSDR_FFMA_SLICES).
    SSDR_ARG(a.sched != nullptr, "demod_launch: scheduler words missing");
    const int nblk = a.n_samples / FR;
    int slices = 1;
    static const int force = [] { const char* e = getenv("SSDR_FFMA_SLICES"); return e ? atoi(e) : 0; }();      // developer knob
    if (force > 0 && nblk % force == 0) slices = force;
    SSDR_CUDA(cudaMemsetAsync(a.sched, 0, sizeof(int) * (size_t)(1 + a.batch), st));
    kern<<<grid, WARPS * 32, smem, st>>>(kp, slices, nblk / slices, a.sched);
    count_launch();
    SSDR_CUDA(cudaGetLastError());
    return SSDR_OK;
}

}  // namespace ssdr
