// Device-visible per-channel demodulator structures and the launch descriptor.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

#include "../../include/ssdr_b200.h"

namespace ssdr {

constexpr float kDemodFsDbm = -10.0f;     // an IQ tone of amplitude FS reads -10 dBm
constexpr float kDemodAgcOut = 0.5f;      // AGC_OUT
constexpr double kDemodAmBeta = 1.0 / (12000.0 * 0.1);   // AM carrier tracker, tau = 0.1 s

// derived per-channel constants (computed on the host from ssdr_demod_params_t)
struct DemodChan {
    int mode;
    int agc_on, agc_hang;
    unsigned inc1, inc2;      // 32-bit phase increments: (f_off + fc)/fs, fc/fs
    float c2;                 // log2(e) / (fs * decay_s): envelope decay exponent per sample
    float knee2;              // (thresh - FS_DBM)/20 * log2(10)
    float slope_m1;           // slope/100 - 1
    float man_gain;           // 10^(manGain/20)
};

struct DemodState {
    unsigned ph1, ph2;
    float e_in;
    float zprev_re, zprev_im;
    unsigned blk;
    float ring[SSDR_HANG_BLOCKS];
    double dc;
};

struct DemodKernelParams {
    const void* iq;
    const DemodChan* chan;
    DemodState* state;
    float2* hist;             // [batch][126] mixed samples
    const float* taps;        // [batch][127]
    float* pcm_f32;
    int16_t* pcm_i16;
    float* rssi;
    int batch, n_samples;
    int pitch;                // samples between consecutive channels in iq / pcm (>= n_samples); rssi rows are pitch/512 apart
    double am_pow16[5];       // ((1-beta)^16)^(2^s)
};

struct DemodLaunch {
    const void* iq = nullptr;
    int iq_format = SSDR_IQ_CF32;
    const DemodChan* chan = nullptr;
    DemodState* state = nullptr;
    float2* hist = nullptr;
    const float* taps = nullptr;
    float* pcm_f32 = nullptr;
    int16_t* pcm_i16 = nullptr;
    float* rssi = nullptr;
    int batch = 0, n_samples = 0;
    int pitch = 0;            // 0 = n_samples (dense)
    double am_pow16[5] = {0, 0, 0, 0, 0};
    int* sched = nullptr;     // FFMA engine: task counter + per-channel progress words, int[1 + batch] (zeroed by the launch)
};

int demod_launch(const DemodLaunch& a, cudaStream_t st);

// tcgen05 engine (demod_tc_kernels.cu).  quad_ch[q] = four channels that share one filter (slots filled in order, unused
// slots -1), quad_fid[q] = id of that filter.  A CTA takes demod_tc_tiles() consecutive quads per round on one B operand:
// the quads of a round share the filter (a round is padded with empty quads).  CTAs pull rounds from *round_ctr (zeroed
// by the launch) in array order, so the host puts the dearest rounds first.
int demod_tc_tiles();
int demod_tc_launch(const DemodLaunch& a, const int4* quad_ch, const int* quad_fid, int n_rounds, int* round_ctr, cudaStream_t st);

}  // namespace ssdr
