// Small kernels: the Tier-P audio interpolator (K5), the IQ wire-format unpack (K6) and the
// synthetic IQ generator used by the bench and the full-size tests.
#include <cuda_runtime.h>

#include <algorithm>

#include <cstdint>

#include "common.cuh"
#include "misc_host.h"

namespace ssdr {

namespace {

// ---------------------------------------------------------------------------------------------
// K5: kiwi_sound.play_buffer integer-ratio path, utils_supersdr.py:1121-1138, in polyphase form.
//   z = concat(hist, zero-stuffed x * (volume/100)); out[j] = ratio * sum_t h[t] z[j + H - t]
// Only taps with (j - t) % ratio == 0 meet non-zero samples, so output j = ratio*k + ph reads
// xs[k + HS - i] * h[ph + ratio*i] (HS = H / ratio scaled history samples).  float64, separate
// multiply and add in ascending tap order (oracle/c/ssdr_oracle.c so_play_buffer).
// ---------------------------------------------------------------------------------------------
constexpr int IT = 256;   // input samples per CTA tile

__global__ void __launch_bounds__(IT)
interp_kernel(const InterpKernelParams kp) {
    __shared__ double xs[IT + SSDR_INTERP_TAPS_MAX];
    __shared__ double hs[SSDR_INTERP_TAPS_MAX + 1];
    const int ch = blockIdx.y;
    const int k0 = blockIdx.x * IT;
    const int HS = (kp.n_taps - 1) / kp.ratio;
    const double scale = (double)kp.volume[ch] / 100.0;
    const int16_t* x = kp.pcm + (size_t)ch * kp.n;
    for (int i = threadIdx.x; i < IT + HS; i += IT) {
        const int k = k0 - HS + i;            // input index, negative = history
        double v = 0.0;
        if (k < 0) { if (HS + k >= 0) v = kp.hist_in[(size_t)ch * HS + (HS + k)]; }
        else if (k < kp.n) v = (double)x[k] * scale;
        xs[i] = v;
    }
    for (int i = threadIdx.x; i < kp.n_taps; i += IT) hs[i] = kp.taps[i];
    __syncthreads();
    const int k = k0 + threadIdx.x;
    if (k >= kp.n) return;
    const double bal = (double)kp.balance[ch];
    double lv = fmin(1.0 - bal, 1.0), rv = fmin(1.0 + bal, 1.0);
    lv = lv * lv; rv = rv * rv;
    // new history = the last HS scaled input samples of this call
    if (k >= kp.n - HS) kp.hist_out[(size_t)ch * HS + (k - (kp.n - HS))] = xs[threadIdx.x + HS];
    for (int ph = 0; ph < kp.ratio; ++ph) {
        double acc = 0.0;
        for (int i = 0; ph + kp.ratio * i < kp.n_taps; ++i) {
            double p = __dmul_rn(hs[ph + kp.ratio * i], xs[threadIdx.x + HS - i]);
            acc = __dadd_rn(acc, p);
        }
        acc = __dmul_rn(acc, (double)kp.ratio);
        const size_t j = (size_t)ch * kp.n * kp.ratio + (size_t)k * kp.ratio + ph;
        if (kp.mono) kp.mono[j] = acc;
        // numpy astype(int16): C truncation toward zero, then wrap to 16 bits
        const int l = __double2int_rz(__dmul_rn(acc, lv)), r = __double2int_rz(__dmul_rn(acc, rv));
        reinterpret_cast<unsigned*>(kp.stereo)[j] = ((unsigned)l & 0xffffu) | ((unsigned)r << 16);
    }
}

// kiwi_sound.play_buffer, non-integer ratio (utils_supersdr.py:1125-1126): scipy.signal.resample_poly(x, up, down,
// padtype="line")[:-1], stateless per block.  The polyphase filter h (zero-padded, scaled by up) is designed by the
// host exactly as resample_poly designs it; this is scipy's upfirdn: output y uses phase t = (y down) mod up and the
// input window ending at x_idx = (y down) div up, the signal extended on both sides along the line through its first
// and last sample, products accumulated in ascending time order with separate float64 multiplies and adds.
__global__ void resample_line_kernel(const ResampleKernelParams kp) {
    const int ch = blockIdx.y;
    const int16_t* x = kp.pcm + (size_t)ch * kp.n;
    const double scale = (double)kp.volume[ch] / 100.0;
    const double bal = (double)kp.balance[ch];
    double lv = fmin(1.0 - bal, 1.0), rv = fmin(1.0 + bal, 1.0);
    lv = lv * lv; rv = rv * rv;
    const int hpp = (kp.n_h + kp.up - 1) / kp.up;              // coefficients per phase (h zero-padded to a multiple of up)
    const double x_first = __dmul_rn((double)x[0], scale), x_last = __dmul_rn((double)x[kp.n - 1], scale);
    const double slope = (kp.n > 1) ? __ddiv_rn(__dadd_rn(x_last, -x_first), (double)(kp.n - 1)) : 0.0;
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < kp.n_keep; j += gridDim.x * blockDim.x) {
        const long long yd = (long long)(j + kp.first) * kp.down;
        const int t = (int)(yd % kp.up), x_idx = (int)(yd / kp.up);
        double acc = 0.0;
        for (int i = 0; i < hpp; ++i) {
            const int xi = x_idx - hpp + 1 + i;
            const int hi = (hpp - 1 - i) * kp.up + t;
            const double hv = (hi < kp.n_h) ? kp.h[hi] : 0.0;
            double xv;
            if (xi < 0) xv = __dadd_rn(x_first, __dmul_rn((double)xi, slope));
            else if (xi >= kp.n) xv = __dadd_rn(x_last, __dmul_rn((double)(xi - kp.n + 1), slope));
            else xv = __dmul_rn((double)x[xi], scale);
            acc = __dadd_rn(acc, __dmul_rn(xv, hv));
        }
        const size_t o = (size_t)ch * kp.n_keep + j;
        if (kp.mono) kp.mono[o] = acc;
        const int l = __double2int_rz(__dmul_rn(acc, lv)), r = __double2int_rz(__dmul_rn(acc, rv));
        reinterpret_cast<unsigned*>(kp.stereo)[o] = ((unsigned)l & 0xffffu) | ((unsigned)r << 16);
    }
}

// filtering.lowpass, utils_supersdr.py:346-348: out[j] = sum_t h[t] x[j + T - 1 - t]  ("valid")
__global__ void fir_valid_kernel(const double* __restrict__ x, const double* __restrict__ h, int T, double* __restrict__ out, size_t n_out) {
    for (size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x; j < n_out; j += (size_t)gridDim.x * blockDim.x) {
        double acc = 0.0;
        for (int t = 0; t < T; ++t) acc = __dadd_rn(acc, __dmul_rn(h[t], x[j + (size_t)(T - 1 - t)]));
        out[j] = acc;
    }
}

// ---------------------------------------------------------------------------------------------
// K6: kiwi/client.py:449-453 -- big-endian int16 I,Q -> complex64 (unscaled counts)
// ---------------------------------------------------------------------------------------------
__global__ void unpack_kernel(const unsigned* __restrict__ in, float2* __restrict__ out, size_t n) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        unsigned v = __ldcs(in + i);
        const unsigned sw = __byte_perm(v, 0u, 0x2301);
        const int a = (int)(short)(sw & 0xffffu), b = (int)sw >> 16;
        __stcs(out + i, make_float2((float)a, (float)b));
    }
}

// ---------------------------------------------------------------------------------------------
// synthetic IQ in HBM: 3 tones + noise per channel, hash-based, deterministic
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned mix32(unsigned x) {
    x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
    return x;
}

template <int FMT>
__global__ void synth_kernel(void* out, int batch, int frames, int nfft, unsigned seed) {
    const size_t per_ch = (size_t)frames * nfft;
    const size_t total = (size_t)batch * per_ch;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const unsigned ch = (unsigned)(i / per_ch);
        const unsigned n = (unsigned)(i - (size_t)ch * per_ch);
        float re = 0.f, im = 0.f;
        const float amp[3] = {0.5f, 0.05f, 0.005f};
#pragma unroll
        for (int tn = 0; tn < 3; ++tn) {
            unsigned hb = mix32(seed * 0x9e3779b9u + ch * 3u + tn);
            // frequency in cycles/sample with 1/2^20 resolution; phase exact in 32-bit arithmetic
            unsigned fcw = hb & 0xfffff000u;
            unsigned ph = fcw * n + mix32(hb);
            float s, c;
            sincospif((float)(int)ph * 4.656612873077393e-10f, &s, &c);
            re += amp[tn] * c; im += amp[tn] * s;
        }
        unsigned r0 = mix32((unsigned)i * 2u + 0x1234567u + seed), r1 = mix32((unsigned)i * 2u + 1u + seed * 77u);
        // sum of four 8-bit uniforms per component: ~Gaussian, std = 147.8 counts/2^8... scaled below
        float nr = (float)((r0 & 255) + ((r0 >> 8) & 255) + ((r0 >> 16) & 255) + (r0 >> 24)) - 510.0f;
        float ni = (float)((r1 & 255) + ((r1 >> 8) & 255) + ((r1 >> 16) & 255) + (r1 >> 24)) - 510.0f;
        const float ns = 1e-3f * 0.70710678f / 147.8f;       // sigma 1e-3 FS (complex)
        re = (re + nr * ns) * SSDR_FS; im = (im + ni * ns) * SSDR_FS;
        if constexpr (FMT == SSDR_IQ_CF32) {
            reinterpret_cast<float2*>(out)[i] = make_float2(re, im);
        } else {
            int a = max(-32768, min(32767, __float2int_rn(re))), b = max(-32768, min(32767, __float2int_rn(im)));
            unsigned ua = (unsigned)a & 0xffffu, ub = (unsigned)b & 0xffffu;
            // big-endian pairs: bytes I_hi I_lo Q_hi Q_lo
            reinterpret_cast<unsigned*>(out)[i] = (ua >> 8) | ((ua & 0xff) << 8) | ((ub >> 8) << 16) | ((ub & 0xff) << 24);
        }
    }
}

}  // namespace

// ---- IMA-ADPCM (kiwi/client.py:33-87): 4-bit codes -> int16, low nibble first; the step index and the
// previous sample are the per-stream state.  Sequential per stream: one thread per stream, 16 input bytes per load.
__constant__ short kAdpcmStep[89] = {
    7, 8, 9, 10, 11, 12, 13, 14, 16, 17, 19, 21, 23, 25, 28, 31, 34, 37, 41, 45, 50, 55, 60, 66, 73, 80, 88, 97, 107, 118, 130,
    143, 157, 173, 190, 209, 230, 253, 279, 307, 337, 371, 408, 449, 494, 544, 598, 658, 724, 796, 876, 963, 1060, 1166, 1282,
    1411, 1552, 1707, 1878, 2066, 2272, 2499, 2749, 3024, 3327, 3660, 4026, 4428, 4871, 5358, 5894, 6484, 7132, 7845, 8630,
    9493, 10442, 11487, 12635, 13899, 15289, 16818, 18500, 20350, 22385, 24623, 27086, 29794, 32767};

__device__ __forceinline__ int adpcm_sample(int code, int& index, int& prev) {
    const int step = kAdpcmStep[index];
    index = min(max(index + ((code & 4) ? 2 * (code & 3) + 2 : -1), 0), 88);      // indexAdjustTable: -1 x4, 2 4 6 8
    int diff = step >> 3;
    if (code & 1) diff += step >> 2;
    if (code & 2) diff += step >> 1;
    if (code & 4) diff += step;
    if (code & 8) diff = -diff;
    prev = min(max(prev + diff, -32768), 32767);
    return prev;
}

__global__ void adpcm_kernel(const uint8_t* __restrict__ data, int batch, int n_bytes, int* __restrict__ state, int16_t* __restrict__ pcm) {
    const int ch = blockIdx.x * blockDim.x + threadIdx.x;
    if (ch >= batch) return;
    int index = min(max(state[2 * ch], 0), 88), prev = state[2 * ch + 1];
    const uint8_t* src = data + (size_t)ch * n_bytes;
    int16_t* dst = pcm + (size_t)ch * n_bytes * 2;
    int i = 0;
    if ((((size_t)src) & 15) == 0 && (((size_t)dst) & 15) == 0) {
        for (; i + 16 <= n_bytes; i += 16) {
            const uint4 v = *reinterpret_cast<const uint4*>(src + i);
            const unsigned w[4] = {v.x, v.y, v.z, v.w};
            unsigned o[16];
#pragma unroll
            for (int k = 0; k < 16; ++k) {
                const unsigned b = (w[k >> 2] >> (8 * (k & 3))) & 0xffu;
                const int s0 = adpcm_sample((int)(b & 15u), index, prev);
                const int s1 = adpcm_sample((int)(b >> 4), index, prev);
                o[k] = ((unsigned)s0 & 0xffffu) | ((unsigned)s1 << 16);
            }
            uint4* q = reinterpret_cast<uint4*>(dst + 2 * i);
            q[0] = make_uint4(o[0], o[1], o[2], o[3]); q[1] = make_uint4(o[4], o[5], o[6], o[7]);
            q[2] = make_uint4(o[8], o[9], o[10], o[11]); q[3] = make_uint4(o[12], o[13], o[14], o[15]);
        }
    }
    for (; i < n_bytes; ++i) {
        const unsigned b = src[i];
        dst[2 * i] = (int16_t)adpcm_sample((int)(b & 15u), index, prev);
        dst[2 * i + 1] = (int16_t)adpcm_sample((int)(b >> 4), index, prev);
    }
    state[2 * ch] = index; state[2 * ch + 1] = prev;
}

int adpcm_launch(const uint8_t* data, int batch, int n_bytes, int* state, int16_t* pcm, cudaStream_t st) {
    adpcm_kernel<<<(batch + 127) / 128, 128, 0, st>>>(data, batch, n_bytes, state, pcm);
    count_launch();
    SSDR_CUDA(cudaGetLastError());
    return SSDR_OK;
}

// ---- display epilogues (8f.3) -----------------------------------------------------------
// The waterfall image of the reference is wf_data[H][W] scrolled down one line per row (utils_supersdr.py:896-897,
// an O(H W) copy) behind a 3-deep delay deque (:893).  Here it is a ring of float32 rows per channel: display line y
// is ring slot (head + y) mod H.
__global__ void image_rgb_kernel(const float* __restrict__ ring, const uint8_t* __restrict__ pal, uint8_t* __restrict__ rgb,
                                 int batch, int H, int W, int head) {
    const size_t total = (size_t)batch * H * W;
    for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
        const int x = (int)(e % W);
        const size_t r = e / W;
        const int y = (int)(r % H), ch = (int)(r / H);
        const float v = ring[((size_t)ch * H + (head + y) % H) * W + x];
        const int idx = min(max(__float2int_rn(v), 0), 255);     // pixel = uint8(rint(wf_color)), DESIGN.md 1
        rgb[3 * e] = pal[3 * idx]; rgb[3 * e + 1] = pal[3 * idx + 1]; rgb[3 * e + 2] = pal[3 * idx + 2];
    }
}

__global__ void image_data_kernel(const float* __restrict__ ring, double* __restrict__ out, int batch, int H, int W, int head) {
    const size_t total = (size_t)batch * H * W;
    for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
        const int x = (int)(e % W);
        const size_t r = e / W;
        const int y = (int)(r % H), ch = (int)(r / H);
        out[e] = (double)ring[((size_t)ch * H + (head + y) % H) * W + x];
    }
}

// display_stuff.plot_spectrum (utils_supersdr.py:1678-1679): v = nanmean of the newest t_avg lines per bin (float64;
// the sum of <= 15 float32 values is exact in float64), y = SH - 1 - int(v / 255 * SH).
__global__ void image_trace_kernel(const float* __restrict__ ring, double* __restrict__ v_out, int* __restrict__ y_out,
                                   int batch, int H, int W, int head, int t_avg, int spectrum_height) {
    const size_t total = (size_t)batch * W;
    for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
        const int x = (int)(e % W), ch = (int)(e / W);
        double sum = 0.0;
        int cnt = 0;
        for (int y = 0; y < t_avg && y < H; ++y) {
            const float f = ring[((size_t)ch * H + (head + y) % H) * W + x];
            if (f == f) { sum = __dadd_rn(sum, (double)f); ++cnt; }
        }
        const double v = cnt ? __ddiv_rn(sum, (double)cnt) : __longlong_as_double(0x7ff8000000000000LL);
        if (v_out) v_out[e] = v;
        if (y_out) y_out[e] = (v == v) ? spectrum_height - 1 - __double2int_rz(__dmul_rn(__ddiv_rn(v, 255.0), (double)spectrum_height)) : -1;
    }
}

int image_rgb_launch(const float* ring, const uint8_t* pal, uint8_t* rgb, int batch, int H, int W, int head, cudaStream_t st) {
    const size_t total = (size_t)batch * H * W;
    const unsigned blocks = (unsigned)std::min<size_t>((total + 255) / 256, (size_t)sm_count() * 16);
    image_rgb_kernel<<<blocks, 256, 0, st>>>(ring, pal, rgb, batch, H, W, head);
    count_launch();
    SSDR_CUDA(cudaGetLastError());
    return SSDR_OK;
}

int image_data_launch(const float* ring, double* out, int batch, int H, int W, int head, cudaStream_t st) {
    const size_t total = (size_t)batch * H * W;
    const unsigned blocks = (unsigned)std::min<size_t>((total + 255) / 256, (size_t)sm_count() * 16);
    image_data_kernel<<<blocks, 256, 0, st>>>(ring, out, batch, H, W, head);
    count_launch();
    SSDR_CUDA(cudaGetLastError());
    return SSDR_OK;
}

int image_trace_launch(const float* ring, double* v, int* y, int batch, int H, int W, int head, int t_avg, int sh, cudaStream_t st) {
    const size_t total = (size_t)batch * W;
    const unsigned blocks = (unsigned)std::min<size_t>((total + 255) / 256, (size_t)sm_count() * 16);
    image_trace_kernel<<<blocks, 256, 0, st>>>(ring, v, y, batch, H, W, head, t_avg, sh);
    count_launch();
    SSDR_CUDA(cudaGetLastError());
    return SSDR_OK;
}

int interp_launch(const InterpLaunch& a, cudaStream_t st) {
    dim3 grid((a.kp.n + IT - 1) / IT, a.batch);
    interp_kernel<<<grid, IT, 0, st>>>(a.kp);
    count_launch();
    SSDR_CUDA(cudaGetLastError());
    return SSDR_OK;
}

int resample_line_launch(const ResampleKernelParams& kp, int batch, cudaStream_t st) {
    dim3 grid((kp.n_keep + 255) / 256, batch);
    resample_line_kernel<<<grid, 256, 0, st>>>(kp);
    count_launch();
    SSDR_CUDA(cudaGetLastError());
    return SSDR_OK;
}

int fir_valid_launch(const double* x, const double* h, int T, double* out, size_t n_out, cudaStream_t st) {
    if (n_out == 0) return SSDR_OK;
    size_t blocks = (n_out + 255) / 256;
    const size_t cap = (size_t)sm_count() * 16;
    if (blocks > cap) blocks = cap;
    fir_valid_kernel<<<(unsigned)blocks, 256, 0, st>>>(x, h, T, out, n_out);
    count_launch();
    SSDR_CUDA(cudaGetLastError());
    return SSDR_OK;
}

int unpack_launch(const void* in, float* out, size_t n, cudaStream_t st) {
    if (n == 0) return SSDR_OK;
    size_t blocks = (n + 255) / 256;
    const size_t cap = (size_t)sm_count() * 16;
    if (blocks > cap) blocks = cap;
    unpack_kernel<<<(unsigned)blocks, 256, 0, st>>>(reinterpret_cast<const unsigned*>(in), reinterpret_cast<float2*>(out), n);
    count_launch();
    SSDR_CUDA(cudaGetLastError());
    return SSDR_OK;
}

int synth_launch(void* out, int fmt, int batch, int frames, int nfft, unsigned seed, cudaStream_t st) {
    const int blocks = sm_count() * 8;
    if (fmt == SSDR_IQ_CF32) synth_kernel<SSDR_IQ_CF32><<<blocks, 256, 0, st>>>(out, batch, frames, nfft, seed);
    else synth_kernel<SSDR_IQ_S16BE><<<blocks, 256, 0, st>>>(out, batch, frames, nfft, seed);
    count_launch();
    SSDR_CUDA(cudaGetLastError());
    return SSDR_OK;
}

}  // namespace ssdr
