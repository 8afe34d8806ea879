// Launch descriptors of the small kernels (interpolator, unpack, synthetic IQ).
#pragma once
#include <cuda_runtime.h>

#include <cstddef>
#include <cstdint>

#include "../../include/ssdr_b200.h"

namespace ssdr {

struct InterpKernelParams {
    const int16_t* pcm;       // [batch][n]
    const float* volume;      // [batch] percent (kiwi_sound.volume)
    const float* balance;     // [batch] (kiwi_sound.audio_balance)
    const double* taps;       // [n_taps]
    const double* hist_in;    // [batch][(n_taps-1)/ratio] scaled samples carried from the last call
    double* hist_out;         // new history (a different buffer than hist_in)
    int16_t* stereo;          // [batch][ratio*n][2]
    double* mono;             // optional [batch][ratio*n]
    int n, ratio, n_taps;
};

struct InterpLaunch {
    InterpKernelParams kp;
    int batch;
};

struct ResampleKernelParams {
    const int16_t* pcm;       // [batch][n]
    const float* volume;      // [batch] percent
    const float* balance;     // [batch]
    const double* h;          // [n_h] polyphase filter as resample_poly builds it (pre-padded, scaled by up)
    int16_t* stereo;          // [batch][n_keep][2]
    double* mono;             // optional [batch][n_keep]
    int n, n_h, up, down, first, n_keep;
};

int image_rgb_launch(const float* ring, const uint8_t* pal, uint8_t* rgb, int batch, int H, int W, int head, cudaStream_t st);
int image_data_launch(const float* ring, double* out, int batch, int H, int W, int head, cudaStream_t st);
int image_trace_launch(const float* ring, double* v, int* y, int batch, int H, int W, int head, int t_avg, int sh, cudaStream_t st);
int adpcm_launch(const uint8_t* data, int batch, int n_bytes, int* state, int16_t* pcm, cudaStream_t st);
int interp_launch(const InterpLaunch& a, cudaStream_t st);
int resample_line_launch(const ResampleKernelParams& kp, int batch, cudaStream_t st);
int fir_valid_launch(const double* x, const double* h, int T, double* out, size_t n_out, cudaStream_t st);
int unpack_launch(const void* in, float* out, size_t n, cudaStream_t st);
int synth_launch(void* out, int fmt, int batch, int frames, int nfft, unsigned seed, cudaStream_t st);

}  // namespace ssdr
