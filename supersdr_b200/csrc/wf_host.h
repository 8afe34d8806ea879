// Host-visible launch descriptors shared between capi.cu and the kernel translation units.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

#include "../../include/ssdr_b200.h"

namespace ssdr {

struct WfLaunch {
    const void* iq = nullptr;          // device, [batch][n_avg][nfft]
    int iq_format = SSDR_IQ_CF32;
    int remote_input = 0;              // iq lives in a peer GPU's memory (read over NVLink): no L2 bulk prefetch
    const uint8_t* lines = nullptr;    // device, colorrow entry (iq ignored)
    const float* wtab = nullptr;       // device, 2*nfft floats
    const float* win = nullptr;        // device, nfft/2 floats (first half of the Hann window)
    const float* thr = nullptr;        // device, 257 floats
    // large-N path (nfft > 16384): twiddle table of the 16384-point sub-transforms, front-pass scratch
    // [batch * rf][n_avg][16384] complex64 and byte sums [batch][nfft] uint16
    const float* wtab_sub = nullptr;
    void* scratch = nullptr;
    uint16_t* sums = nullptr;
    ssdr_wf_display_t* disp = nullptr; // device, [batch]
    uint8_t* pixels = nullptr;
    float* colour = nullptr;
    float* spectrum = nullptr;
    ssdr_wf_scalars_t* scalars = nullptr;
    int nfft = 0, batch = 0, n_avg = 1, window = 1;
    int p_lo = 0;
    float p_gamma = 0.f;
    float est_c1 = 0.f, est_c0 = 0.f;
};

int wf_plan(int nfft, int* radices);
bool wf_big_fused();                                                   // large-N path: fused kernel (default) or three kernels
size_t wf_big_scratch_bytes(int nfft, int n_avg, int channels);        // scratch the large-N path needs for `channels` channels
int wf_launch(const WfLaunch& a, cudaStream_t st);

struct DemodLaunch;
struct InterpLaunch;

}  // namespace ssdr
