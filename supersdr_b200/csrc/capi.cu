// C ABI of libssdr_b200.so (include/ssdr_b200.h): handles, device buffers, streams, and the
// host<->device pipelines around the kernels in wf_kernels.cu / demod_kernels.cu / misc_kernels.cu.
// No CPU fallback: every compute entry point fails with SSDR_E_CUDA when no device is usable.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <vector>

#include "common.cuh"
#include "demod_host.h"
#include "misc_host.h"
#include "wf_host.h"

namespace ssdr {

static thread_local char g_err[512] = "";
std::atomic<uint64_t> g_launches{0};
static int g_sm_count = 0;
// The CUDA device is a per-thread setting of the runtime.  ssdr_init records the process's selection; every handle
// records the device it was created on; every entry point binds the calling thread to that device first, so a handle
// (and the stateless entry points) work from any thread -- the reference's waterfall / sound / PortAudio-callback
// threads (utils_supersdr.py:879,1106,1150) included -- not only from the thread that called ssdr_init.
std::atomic<int> g_device{-1};

int bind_device(int device) {
    if (device < 0) return SSDR_OK;              // ssdr_init was never called: the runtime's default (device 0)
    int cur = -1;
    cudaError_t e = cudaGetDevice(&cur);
    if (e == cudaSuccess && cur != device) e = cudaSetDevice(device);
    if (e != cudaSuccess) { set_error("cannot bind the calling thread to device %d: %s", device, cudaGetErrorString(e)); return SSDR_E_CUDA; }
    return SSDR_OK;
}
#define SSDR_BIND(dev)                                       \
    do {                                                     \
        int rc_bind__ = ssdr::bind_device(dev);              \
        if (rc_bind__) return rc_bind__;                     \
    } while (0)
#define SSDR_BIND_H(h) SSDR_BIND((h)->device)
#define SSDR_BIND_G() SSDR_BIND(ssdr::g_device.load())

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int sm_count() {
    if (g_sm_count == 0) {
        int dev = 0;
        cudaDeviceProp p;
        if (cudaGetDevice(&dev) == cudaSuccess && cudaGetDeviceProperties(&p, dev) == cudaSuccess) g_sm_count = p.multiProcessorCount;
        else g_sm_count = 148;
    }
    return g_sm_count;
}

template <class Tp>
static int dev_alloc(Tp** p, size_t count) {
    *p = nullptr;
    if (count == 0) return SSDR_OK;
    cudaError_t e = cudaMalloc(reinterpret_cast<void**>(p), count * sizeof(Tp));
    if (e != cudaSuccess) { set_error("cudaMalloc(%zu bytes) failed: %s", count * sizeof(Tp), cudaGetErrorString(e)); return e == cudaErrorMemoryAllocation ? SSDR_E_NOMEM : SSDR_E_CUDA; }
    return SSDR_OK;
}

}  // namespace ssdr

using namespace ssdr;

// =============================================================================================
// handles
// =============================================================================================
struct ssdr_wf {
    int device = -1;                  // the device this handle lives on (bound at every entry point)
    int nfft = 0, batch = 0, n_avg = 1, window = 1, p_lo = 0;
    int remote_input = 0;             // device inputs live in a peer GPU's memory (ssdr_wf_set_remote_input)
    float p_gamma = 0.f;
    double cal_db = 0.0;
    float est_c1 = 0.f, est_c0 = 0.f;
    std::vector<float> h_wtab, h_thr, h_win;
    float* d_wtab = nullptr;
    float* d_win = nullptr;
    // large-N path (nfft > 16384)
    float* d_wtab_sub = nullptr;      // 16384-point twiddle table of the sub-transforms
    void* d_scratch = nullptr;        // front-pass output, sized for scratch_ch channels
    size_t scratch_bytes = 0;
    uint16_t* d_sums = nullptr;       // [batch][nfft]
    float* d_thr = nullptr;
    ssdr_wf_display_t* d_disp = nullptr;
    cudaStream_t compute = nullptr, copy = nullptr;
    cudaEvent_t ev_copied[2] = {nullptr, nullptr}, ev_done[2] = {nullptr, nullptr}, ev_t0 = nullptr, ev_t1 = nullptr;
    // host-API staging (lazily allocated): double-buffered input chunks + full-size outputs
    int chunk_ch = 0;
    void* d_in[2] = {nullptr, nullptr};
    size_t in_bytes = 0;
    uint8_t* d_px = nullptr;
    float* d_col = nullptr;
    float* d_spec = nullptr;
    ssdr_wf_scalars_t* d_sc = nullptr;
};

struct ssdr_demod {
    int device = -1;                  // the device this handle lives on (bound at every entry point)
    int batch = 0, max_samples = 0;
    DemodChan* d_chan = nullptr;
    DemodState* d_state = nullptr;
    float2* d_hist = nullptr;
    float* d_taps = nullptr;
    cudaStream_t compute = nullptr, copy_in = nullptr, copy_out = nullptr;
    cudaEvent_t ev_t0 = nullptr, ev_t1 = nullptr;
    cudaEvent_t ev_copied = nullptr, ev_done = nullptr;
    double am_pow16[5];
    void* d_in = nullptr;          // device staging of the host API, [batch][max_samples] complex64
    size_t in_bytes = 0;
    float* d_f32 = nullptr;
    int16_t* d_i16 = nullptr;
    float* d_rssi = nullptr;
    // FIR engine (ssdr_demod_set_engine) and, for the tcgen05 engine, the channels grouped in quads that share a filter
    int engine = SSDR_DEMOD_ENGINE_AUTO;
    double quad_fill = 0.0;        // channels / (4 x non-empty quads): how full the tensor-core tiles are
    std::vector<float> h_taps;     // host mirror of d_taps, [batch][127]
    std::vector<int> h_work;       // per channel: detector / AGC variant (channels of one quad should cost the same)
    bool quads_dirty = true;
    int n_quads = 0;               // rounds of demod_tc_tiles() quads
    int used_quads = 0;            // non-empty quads (tiles that carry at least one channel) over all rounds
    int4* d_quad_ch = nullptr;
    int* d_quad_fid = nullptr;
    int* d_round_ctr = nullptr;    // work counter of the tcgen05 kernel (dynamic round scheduling)
    int* d_sched = nullptr;        // FFMA engine: task counter + per-channel progress words, int[1 + batch]
};

struct ssdr_interp {
    int device = -1;                  // the device this handle lives on (bound at every entry point)
    int batch = 0, ratio = 0, n_taps = 0, max_samples = 0, hs = 0, cur = 0;
    double* d_taps = nullptr;
    double* d_hist[2] = {nullptr, nullptr};
    cudaStream_t compute = nullptr;
    int16_t* d_in = nullptr;
    float* d_vol = nullptr;
    float* d_bal = nullptr;
    int16_t* d_out = nullptr;
    double* d_mono = nullptr;
};

struct ssdr_wf_image {
    int device = -1;                  // the device this handle lives on (bound at every entry point)
    int batch = 0, H = 0, W = 0, head = 0;
    long long run_index = 0;
    float* d_ring = nullptr;          // [batch][H][W]
    float* d_delay = nullptr;         // [3][batch][W]
    float* d_row = nullptr;           // one row of 255s (white flag)
    uint8_t* d_pal = nullptr;         // [256][3]
    uint8_t* d_rgb = nullptr;
    double* d_f64 = nullptr;
    int* d_y = nullptr;
    std::vector<int> deque;           // delay slots, newest first (utils_supersdr.py:893, maxlen 3)
    cudaStream_t st = nullptr;
};

extern "C" {

// =============================================================================================
// library / device
// =============================================================================================
int ssdr_abi_version(void) { return SSDR_ABI_VERSION; }
const char* ssdr_last_error(void) { return g_err; }
uint64_t ssdr_launch_count(void) { return g_launches.load(); }

int ssdr_init(int device) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) { set_error("no CUDA device: %s", cudaGetErrorString(e)); return SSDR_E_CUDA; }
    SSDR_ARG(device >= 0 && device < n, "device %d out of range (0..%d)", device, n - 1);
    SSDR_CUDA(cudaSetDevice(device));
    cudaDeviceProp p;
    SSDR_CUDA(cudaGetDeviceProperties(&p, device));
    if (p.major != 10) { set_error("libssdr_b200 is built for sm_100a only; device %d is sm_%d%d", device, p.major, p.minor); return SSDR_E_CUDA; }
    g_sm_count = p.multiProcessorCount;
    SSDR_CUDA(cudaFree(0));
    g_device.store(device);
    return SSDR_OK;
}

int ssdr_device_info(int* sms, int* cc_major, int* cc_minor, size_t* hbm_bytes, char* name, int name_len) {
    SSDR_BIND_G();
    int dev = 0;
    SSDR_CUDA(cudaGetDevice(&dev));
    cudaDeviceProp p;
    SSDR_CUDA(cudaGetDeviceProperties(&p, dev));
    if (sms) *sms = p.multiProcessorCount;
    if (cc_major) *cc_major = p.major;
    if (cc_minor) *cc_minor = p.minor;
    if (hbm_bytes) *hbm_bytes = p.totalGlobalMem;
    if (name && name_len > 0) { std::strncpy(name, p.name, (size_t)name_len - 1); name[name_len - 1] = 0; }
    return SSDR_OK;
}

int ssdr_device_pci_bus_id(char* bus_id, int len) {
    SSDR_ARG(bus_id && len >= 16, "bus_id buffer of at least 16 bytes");
    SSDR_BIND_G();
    int dev = 0;
    SSDR_CUDA(cudaGetDevice(&dev));
    SSDR_CUDA(cudaDeviceGetPCIBusId(bus_id, len, dev));
    return SSDR_OK;
}

int ssdr_dev_alloc(void** dev, size_t bytes) {
    SSDR_ARG(dev != nullptr, "null pointer");
    SSDR_BIND_G();
    return dev_alloc(reinterpret_cast<unsigned char**>(dev), bytes);
}
int ssdr_dev_free(void* dev) { SSDR_BIND_G(); SSDR_CUDA(cudaFree(dev)); return SSDR_OK; }
int ssdr_host_alloc(void** host, size_t bytes) { SSDR_BIND_G(); SSDR_ARG(host != nullptr, "null pointer"); SSDR_CUDA(cudaMallocHost(host, bytes)); return SSDR_OK; }
int ssdr_host_free(void* host) { SSDR_BIND_G(); SSDR_CUDA(cudaFreeHost(host)); return SSDR_OK; }
int ssdr_memcpy_h2d(void* dev, const void* host, size_t bytes) { SSDR_BIND_G(); SSDR_CUDA(cudaMemcpy(dev, host, bytes, cudaMemcpyHostToDevice)); return SSDR_OK; }
int ssdr_memcpy_d2h(void* host, const void* dev, size_t bytes) { SSDR_BIND_G(); SSDR_CUDA(cudaMemcpy(host, dev, bytes, cudaMemcpyDeviceToHost)); return SSDR_OK; }
int ssdr_dev_memset(void* dev, int value, size_t bytes) { SSDR_BIND_G(); SSDR_CUDA(cudaMemset(dev, value, bytes)); return SSDR_OK; }
int ssdr_device_sync(void) { SSDR_BIND_G(); SSDR_CUDA(cudaDeviceSynchronize()); return SSDR_OK; }

int ssdr_ipc_export(void* dev, void* handle64) {
    SSDR_ARG(dev && handle64, "null argument");
    SSDR_BIND_G();
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
    cudaIpcMemHandle_t h;
    SSDR_CUDA(cudaIpcGetMemHandle(&h, dev));
    std::memcpy(handle64, &h, 64);
    return SSDR_OK;
}

int ssdr_ipc_open(const void* handle64, void** dev) {
    SSDR_ARG(handle64 && dev, "null argument");
    SSDR_BIND_G();
    cudaIpcMemHandle_t h;
    std::memcpy(&h, handle64, 64);
    SSDR_CUDA(cudaIpcOpenMemHandle(dev, h, cudaIpcMemLazyEnablePeerAccess));
    return SSDR_OK;
}

int ssdr_ipc_close(void* dev) {
    SSDR_ARG(dev != nullptr, "null argument");
    SSDR_BIND_G();
    SSDR_CUDA(cudaIpcCloseMemHandle(dev));
    return SSDR_OK;
}

int ssdr_synth_iq_dev(void* iq_dev, int iq_format, int batch, int frames, int nfft, uint32_t seed) {
    SSDR_ARG(iq_dev && batch > 0 && frames > 0 && nfft > 0, "bad synth arguments");
    SSDR_BIND_G();
    SSDR_ARG(iq_format == SSDR_IQ_CF32 || iq_format == SSDR_IQ_S16BE, "bad iq_format %d", iq_format);
    int rc = synth_launch(iq_dev, iq_format, batch, frames, nfft, seed, 0);
    if (rc) return rc;
    SSDR_CUDA(cudaDeviceSynchronize());
    return SSDR_OK;
}

// =============================================================================================
// waterfall
// =============================================================================================
// Master twiddle table W_N^k = (float(cos), float(-sin))(2 pi k / N), evaluated in double
static void fill_twiddles(int N, float* tab) {
    for (int k = 0; k < N; ++k) {
        const double a = 2.0 * 3.14159265358979323846 * (double)k / (double)N;
        tab[2 * k] = (float)std::cos(a);
        tab[2 * k + 1] = (float)(-std::sin(a));
    }
}

static size_t iq_sample_bytes(int fmt) { return fmt == SSDR_IQ_CF32 ? 8 : 4; }
static bool aligned16(const void* p) { return ((uintptr_t)p & 15u) == 0; }

int ssdr_wf_create(ssdr_wf_t* out, int nfft, int batch, int n_avg, int window, double cal_db, int p_lo, float p_gamma) {
    SSDR_ARG(out != nullptr, "null handle pointer");
    *out = nullptr;
    int radices[8];
    SSDR_ARG(wf_plan(nfft, radices) > 0, "nfft %d unsupported (power of two 256..65536)", nfft);
    SSDR_ARG(batch >= 1, "batch %d < 1", batch);
    SSDR_ARG(n_avg >= 1 && n_avg <= 100, "n_avg %d outside 1..100 (supersdr.py:376-385)", n_avg);
    SSDR_ARG(p_lo >= 0 && p_lo < nfft && p_gamma >= 0.f && p_gamma < 1.f, "bad percentile index (%d, %g)", p_lo, (double)p_gamma);
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { set_error("no CUDA device (libssdr_b200 has no CPU fallback)"); return SSDR_E_CUDA; }
    SSDR_BIND_G();
    ssdr_wf* h = new ssdr_wf();
    if (cudaGetDevice(&h->device) != cudaSuccess) h->device = -1;
    h->nfft = nfft; h->batch = batch; h->n_avg = n_avg; h->window = window ? 1 : 0; h->cal_db = cal_db;
    h->p_lo = p_lo; h->p_gamma = p_gamma;
    // spec tables: twiddles W_N^k = (cos, -sin)(2 pi k / N) and the byte thresholds
    h->h_wtab.resize(2 * (size_t)nfft);
    fill_twiddles(nfft, h->h_wtab.data());
    h->h_win.resize((size_t)nfft / 2);          // first half of the periodic Hann window
    for (int n = 0; n < nfft / 2; ++n)
        h->h_win[n] = (float)(0.5 - 0.5 * std::cos(2.0 * 3.14159265358979323846 * (double)n / (double)nfft));
    h->h_thr.resize(257);
    double ref = (double)nfft * 32768.0 * 0.5;
    ref = ref * ref;
    h->h_thr[0] = 0.0f;
    for (int k = 1; k < 256; ++k) h->h_thr[k] = (float)(ref * std::pow(10.0, ((double)k - 0.5 - 255.0 - cal_db) / 10.0));
    h->h_thr[256] = std::numeric_limits<float>::infinity();
    h->est_c1 = (float)(10.0 * std::log10(2.0));
    h->est_c0 = (float)(-10.0 * std::log10(ref) + cal_db + 255.0);
    int rc = SSDR_OK;
    auto fail = [&](int code) { ssdr_wf_destroy(h); return code; };
    if ((rc = dev_alloc(&h->d_wtab, 2 * (size_t)nfft))) return fail(rc);
    if ((rc = dev_alloc(&h->d_thr, 257))) return fail(rc);
    if ((rc = dev_alloc(&h->d_win, (size_t)nfft / 2))) return fail(rc);
    if (nfft > 16384) {
        std::vector<float> sub(2 * 16384);
        fill_twiddles(16384, sub.data());
        if ((rc = dev_alloc(&h->d_wtab_sub, sub.size()))) return fail(rc);
        if ((rc = dev_alloc(&h->d_sums, (size_t)batch * nfft))) return fail(rc);
        if (cudaMemcpy(h->d_wtab_sub, sub.data(), sizeof(float) * sub.size(), cudaMemcpyHostToDevice) != cudaSuccess) { set_error("table upload failed"); return fail(SSDR_E_CUDA); }
    }
    if ((rc = dev_alloc(&h->d_disp, (size_t)batch))) return fail(rc);
    if ((rc = dev_alloc(&h->d_sc, (size_t)batch))) return fail(rc);
    if (cudaMemcpy(h->d_wtab, h->h_wtab.data(), sizeof(float) * 2 * nfft, cudaMemcpyHostToDevice) != cudaSuccess ||
        cudaMemcpy(h->d_win, h->h_win.data(), sizeof(float) * (nfft / 2), cudaMemcpyHostToDevice) != cudaSuccess ||
        cudaMemcpy(h->d_thr, h->h_thr.data(), sizeof(float) * 257, cudaMemcpyHostToDevice) != cudaSuccess) {
        set_error("table upload failed");
        return fail(SSDR_E_CUDA);
    }
    std::vector<ssdr_wf_display_t> disp((size_t)batch);
    for (auto& d : disp) { d.zoom = 0; d.auto_scale = 1; d.delta_low_db = 0; d.delta_high_db = 0; d.low_clip_db = -120.f; d.dynamic_range = 40.f; }
    if (cudaMemcpy(h->d_disp, disp.data(), sizeof(ssdr_wf_display_t) * batch, cudaMemcpyHostToDevice) != cudaSuccess) { set_error("display upload failed"); return fail(SSDR_E_CUDA); }
    if (cudaStreamCreateWithFlags(&h->compute, cudaStreamNonBlocking) != cudaSuccess ||
        cudaStreamCreateWithFlags(&h->copy, cudaStreamNonBlocking) != cudaSuccess) { set_error("stream creation failed"); return fail(SSDR_E_CUDA); }
    for (int i = 0; i < 2; ++i) {
        if (cudaEventCreateWithFlags(&h->ev_copied[i], cudaEventDisableTiming) != cudaSuccess ||
            cudaEventCreateWithFlags(&h->ev_done[i], cudaEventDisableTiming) != cudaSuccess) { set_error("event creation failed"); return fail(SSDR_E_CUDA); }
    }
    if (cudaEventCreate(&h->ev_t0) != cudaSuccess || cudaEventCreate(&h->ev_t1) != cudaSuccess) { set_error("event creation failed"); return fail(SSDR_E_CUDA); }
    *out = h;
    return SSDR_OK;
}

int ssdr_wf_destroy(ssdr_wf_t h) {
    if (!h) return SSDR_OK;
    bind_device(h->device);
    if (h->compute) cudaStreamSynchronize(h->compute);
    if (h->copy) cudaStreamSynchronize(h->copy);
    cudaFree(h->d_wtab); cudaFree(h->d_win); cudaFree(h->d_wtab_sub); cudaFree(h->d_scratch); cudaFree(h->d_sums); cudaFree(h->d_thr); cudaFree(h->d_disp); cudaFree(h->d_in[0]); cudaFree(h->d_in[1]);
    cudaFree(h->d_px); cudaFree(h->d_col); cudaFree(h->d_spec); cudaFree(h->d_sc);
    for (int i = 0; i < 2; ++i) { if (h->ev_copied[i]) cudaEventDestroy(h->ev_copied[i]); if (h->ev_done[i]) cudaEventDestroy(h->ev_done[i]); }
    if (h->ev_t0) cudaEventDestroy(h->ev_t0);
    if (h->ev_t1) cudaEventDestroy(h->ev_t1);
    if (h->compute) cudaStreamDestroy(h->compute);
    if (h->copy) cudaStreamDestroy(h->copy);
    delete h;
    return SSDR_OK;
}

int ssdr_wf_set_display(ssdr_wf_t h, int first, int count, const ssdr_wf_display_t* params) {
    SSDR_ARG(h && params, "null argument");
    SSDR_BIND_H(h);
    SSDR_ARG(first >= 0 && count >= 0 && first + count <= h->batch, "channel range [%d, %d) outside batch %d", first, first + count, h->batch);
    SSDR_CUDA(cudaStreamSynchronize(h->compute));
    SSDR_CUDA(cudaMemcpy(h->d_disp + first, params, sizeof(ssdr_wf_display_t) * (size_t)count, cudaMemcpyHostToDevice));
    return SSDR_OK;
}

int ssdr_wf_set_remote_input(ssdr_wf_t h, int remote) {
    SSDR_ARG(h != nullptr, "null handle");
    h->remote_input = remote ? 1 : 0;
    return SSDR_OK;
}

int ssdr_wf_get_tables(ssdr_wf_t h, float* twiddles, float* thresholds, int* radices) {
    SSDR_ARG(h != nullptr, "null handle");
    if (twiddles) std::memcpy(twiddles, h->h_wtab.data(), sizeof(float) * 2 * (size_t)h->nfft);
    if (thresholds) std::memcpy(thresholds, h->h_thr.data(), sizeof(float) * 256);
    int r[8];
    int np = wf_plan(h->nfft, r);
    if (radices) for (int i = 0; i < np; ++i) radices[i] = r[i];
    return np;
}

int ssdr_wf_get_window(ssdr_wf_t h, float* window_half) {
    SSDR_ARG(h && window_half, "null argument");
    std::memcpy(window_half, h->h_win.data(), sizeof(float) * (size_t)(h->nfft / 2));
    return SSDR_OK;
}

static WfLaunch wf_base(ssdr_wf_t h) {
    WfLaunch a;
    a.wtab = h->d_wtab; a.win = h->d_win; a.thr = h->d_thr; a.nfft = h->nfft; a.n_avg = h->n_avg; a.window = h->window;
    a.p_lo = h->p_lo; a.p_gamma = h->p_gamma; a.est_c1 = h->est_c1; a.est_c0 = h->est_c0;
    a.wtab_sub = h->d_wtab_sub; a.scratch = h->d_scratch; a.sums = h->d_sums;
    return a;
}

// large-N path: scratch of the front pass -- per SM for the fused kernel (L2-resident), per channel for the three-kernel path
static int wf_ensure_scratch(ssdr_wf_t h, int channels) {
    if (h->nfft <= 16384) return SSDR_OK;
    const size_t need = wf_big_scratch_bytes(h->nfft, h->n_avg, channels);
    if (h->scratch_bytes >= need) return SSDR_OK;
    SSDR_CUDA(cudaStreamSynchronize(h->compute));
    cudaFree(h->d_scratch); h->d_scratch = nullptr; h->scratch_bytes = 0;
    int rc = dev_alloc(reinterpret_cast<unsigned char**>(&h->d_scratch), need);
    if (rc) return rc;
    h->scratch_bytes = need;
    return SSDR_OK;
}

int ssdr_wf_process_dev(ssdr_wf_t h, const void* iq_dev, int iq_format, uint8_t* pixels_dev, float* colour_dev,
                        float* spectrum_dev, ssdr_wf_scalars_t* scalars_dev) {
    SSDR_ARG(h && iq_dev, "null argument");
    SSDR_BIND_H(h);
    SSDR_ARG(iq_format == SSDR_IQ_CF32 || iq_format == SSDR_IQ_S16BE, "bad iq_format %d", iq_format);
    SSDR_ARG(((uintptr_t)iq_dev & 15u) == 0, "iq_dev must be 16-byte aligned (bulk prefetch / TMA tile copies)");
    int rcs = wf_ensure_scratch(h, h->batch);
    if (rcs) return rcs;
    WfLaunch a = wf_base(h);
    a.iq = iq_dev; a.iq_format = iq_format; a.disp = h->d_disp; a.batch = h->batch;
    a.remote_input = h->remote_input;
    a.pixels = pixels_dev; a.colour = colour_dev; a.spectrum = spectrum_dev; a.scalars = scalars_dev ? scalars_dev : h->d_sc;
    return wf_launch(a, h->compute);
}

int ssdr_wf_colorrow_u8_dev(ssdr_wf_t h, const uint8_t* lines_dev, uint8_t* pixels_dev, float* colour_dev,
                            float* spectrum_dev, ssdr_wf_scalars_t* scalars_dev) {
    SSDR_ARG(h && lines_dev, "null argument");
    SSDR_BIND_H(h);
    SSDR_ARG(h->nfft <= 16384, "the uint8-line entry supports nfft <= 16384 (got %d)", h->nfft);
    WfLaunch a = wf_base(h);
    a.lines = lines_dev; a.disp = h->d_disp; a.batch = h->batch;
    a.pixels = pixels_dev; a.colour = colour_dev; a.spectrum = spectrum_dev; a.scalars = scalars_dev ? scalars_dev : h->d_sc;
    return wf_launch(a, h->compute);
}

int ssdr_wf_sync(ssdr_wf_t h) {
    SSDR_ARG(h != nullptr, "null handle");
    SSDR_BIND_H(h);
    SSDR_CUDA(cudaStreamSynchronize(h->compute));
    SSDR_CUDA(cudaStreamSynchronize(h->copy));
    return SSDR_OK;
}

// lazily allocate the host-API staging buffers
static int wf_ensure_staging(ssdr_wf_t h, size_t bytes_per_channel, bool want_col, bool want_spec) {
    if (h->chunk_ch == 0 || h->in_bytes < bytes_per_channel * (size_t)h->chunk_ch) {
        cudaFree(h->d_in[0]); cudaFree(h->d_in[1]);
        h->d_in[0] = h->d_in[1] = nullptr;
        // ~256 MiB per chunk: large enough to fill the GPU, small enough to overlap copy and compute
        // (SSDR_WF_CHUNK_MB: developer override for the end-to-end experiments of DESIGN.md section 7)
        size_t chunk_mb = 256;
        if (const char* e = std::getenv("SSDR_WF_CHUNK_MB")) { const long v = std::atol(e); if (v >= 1 && v <= 4096) chunk_mb = (size_t)v; }
        size_t ch = std::max<size_t>(1, (chunk_mb << 20) / bytes_per_channel);
        const size_t wave = (size_t)sm_count();
        if (ch >= wave) ch = ch / wave * wave;
        ch = std::min<size_t>(ch, (size_t)h->batch);
        h->chunk_ch = (int)ch;
        h->in_bytes = bytes_per_channel * ch;
        int rc;
        for (int i = 0; i < 2; ++i)
            if ((rc = dev_alloc(reinterpret_cast<unsigned char**>(&h->d_in[i]), h->in_bytes))) return rc;
    }
    int rc;
    if (!h->d_px && (rc = dev_alloc(&h->d_px, (size_t)h->batch * h->nfft))) return rc;
    if (want_col && !h->d_col && (rc = dev_alloc(&h->d_col, (size_t)h->batch * h->nfft))) return rc;
    if (want_spec && !h->d_spec && (rc = dev_alloc(&h->d_spec, (size_t)h->batch * h->nfft))) return rc;
    return SSDR_OK;
}

static int wf_process_host(ssdr_wf_t h, const void* in_host, size_t bytes_per_channel, int iq_format, bool lines,
                           uint8_t* pixels, float* colour, float* spectrum, ssdr_wf_scalars_t* scalars) {
    int rc = wf_ensure_staging(h, bytes_per_channel, colour != nullptr, spectrum != nullptr);
    if (rc) return rc;
    if (!lines && (rc = wf_ensure_scratch(h, h->chunk_ch))) return rc;
    const size_t N = (size_t)h->nfft;
    int slot = 0;
    for (int c0 = 0; c0 < h->batch; c0 += h->chunk_ch, slot ^= 1) {
        const int nch = std::min(h->chunk_ch, h->batch - c0);
        // the kernel that last read this slot must be done before we overwrite it
        SSDR_CUDA(cudaStreamWaitEvent(h->copy, h->ev_done[slot], 0));
        SSDR_CUDA(cudaMemcpyAsync(h->d_in[slot], static_cast<const unsigned char*>(in_host) + (size_t)c0 * bytes_per_channel,
                                  (size_t)nch * bytes_per_channel, cudaMemcpyHostToDevice, h->copy));
        SSDR_CUDA(cudaEventRecord(h->ev_copied[slot], h->copy));
        SSDR_CUDA(cudaStreamWaitEvent(h->compute, h->ev_copied[slot], 0));
        WfLaunch a = wf_base(h);
        if (lines) a.lines = static_cast<const uint8_t*>(h->d_in[slot]);
        else { a.iq = h->d_in[slot]; a.iq_format = iq_format; }
        a.disp = h->d_disp + c0; a.batch = nch;
        if (a.sums) a.sums += (size_t)c0 * N;
        a.pixels = h->d_px + (size_t)c0 * N;
        a.colour = colour ? h->d_col + (size_t)c0 * N : nullptr;
        a.spectrum = spectrum ? h->d_spec + (size_t)c0 * N : nullptr;
        a.scalars = h->d_sc + c0;
        if ((rc = wf_launch(a, h->compute))) return rc;
        SSDR_CUDA(cudaEventRecord(h->ev_done[slot], h->compute));
        // results of this chunk go back on the compute stream (small next to the input)
        if (pixels) SSDR_CUDA(cudaMemcpyAsync(pixels + (size_t)c0 * N, a.pixels, (size_t)nch * N, cudaMemcpyDeviceToHost, h->compute));
        if (colour) SSDR_CUDA(cudaMemcpyAsync(colour + (size_t)c0 * N, a.colour, (size_t)nch * N * sizeof(float), cudaMemcpyDeviceToHost, h->compute));
        if (spectrum) SSDR_CUDA(cudaMemcpyAsync(spectrum + (size_t)c0 * N, a.spectrum, (size_t)nch * N * sizeof(float), cudaMemcpyDeviceToHost, h->compute));
        if (scalars) SSDR_CUDA(cudaMemcpyAsync(scalars + c0, a.scalars, (size_t)nch * sizeof(ssdr_wf_scalars_t), cudaMemcpyDeviceToHost, h->compute));
    }
    SSDR_CUDA(cudaStreamSynchronize(h->compute));
    SSDR_CUDA(cudaStreamSynchronize(h->copy));
    return SSDR_OK;
}

int ssdr_wf_process(ssdr_wf_t h, const void* iq_host, int iq_format, uint8_t* pixels, float* colour, float* spectrum,
                    ssdr_wf_scalars_t* scalars) {
    SSDR_ARG(h && iq_host, "null argument");
    SSDR_BIND_H(h);
    SSDR_ARG(iq_format == SSDR_IQ_CF32 || iq_format == SSDR_IQ_S16BE, "bad iq_format %d", iq_format);
    return wf_process_host(h, iq_host, (size_t)h->n_avg * h->nfft * iq_sample_bytes(iq_format), iq_format, false,
                           pixels, colour, spectrum, scalars);
}

int ssdr_wf_colorrow_u8(ssdr_wf_t h, const uint8_t* lines_host, uint8_t* pixels, float* colour, float* spectrum,
                        ssdr_wf_scalars_t* scalars) {
    SSDR_ARG(h && lines_host, "null argument");
    SSDR_BIND_H(h);
    SSDR_ARG(h->nfft <= 16384, "the uint8-line entry supports nfft <= 16384 (got %d)", h->nfft);
    return wf_process_host(h, lines_host, (size_t)h->n_avg * h->nfft, 0, true, pixels, colour, spectrum, scalars);
}

int ssdr_wf_time_dev(ssdr_wf_t h, const void* iq_dev, int iq_format, uint8_t* pixels_dev, int iters, float* total_ms) {
    SSDR_ARG(h && iq_dev && total_ms && iters >= 1, "bad argument");
    SSDR_BIND_H(h);
    SSDR_CUDA(cudaEventRecord(h->ev_t0, h->compute));
    for (int i = 0; i < iters; ++i) {
        int rc = ssdr_wf_process_dev(h, iq_dev, iq_format, pixels_dev, nullptr, nullptr, nullptr);
        if (rc) return rc;
    }
    SSDR_CUDA(cudaEventRecord(h->ev_t1, h->compute));
    SSDR_CUDA(cudaEventSynchronize(h->ev_t1));
    SSDR_CUDA(cudaEventElapsedTime(total_ms, h->ev_t0, h->ev_t1));
    return SSDR_OK;
}

// =============================================================================================
// demodulator
// =============================================================================================
int ssdr_demod_create(ssdr_demod_t* out, int batch, int max_samples) {
    SSDR_ARG(out != nullptr, "null handle pointer");
    *out = nullptr;
    SSDR_ARG(batch >= 1, "batch %d < 1", batch);
    SSDR_ARG(max_samples >= SSDR_FRAME && max_samples % SSDR_FRAME == 0, "max_samples_per_call %d must be a positive multiple of %d", max_samples, SSDR_FRAME);
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { set_error("no CUDA device (libssdr_b200 has no CPU fallback)"); return SSDR_E_CUDA; }
    SSDR_BIND_G();
    ssdr_demod* h = new ssdr_demod();
    if (cudaGetDevice(&h->device) != cudaSuccess) h->device = -1;
    h->batch = batch; h->max_samples = max_samples;
    double om16 = std::pow(1.0 - kDemodAmBeta, 16.0);
    for (int s = 0; s < 5; ++s) { h->am_pow16[s] = om16; om16 *= om16; }
    int rc;
    auto fail = [&](int code) { ssdr_demod_destroy(h); return code; };
    if ((rc = dev_alloc(&h->d_chan, (size_t)batch))) return fail(rc);
    if ((rc = dev_alloc(&h->d_state, (size_t)batch))) return fail(rc);
    if ((rc = dev_alloc(&h->d_hist, (size_t)batch * (SSDR_FIR_TAPS - 1)))) return fail(rc);
    if ((rc = dev_alloc(&h->d_taps, (size_t)batch * SSDR_FIR_TAPS))) return fail(rc);
    if (cudaStreamCreateWithFlags(&h->compute, cudaStreamNonBlocking) != cudaSuccess ||
        cudaStreamCreateWithFlags(&h->copy_in, cudaStreamNonBlocking) != cudaSuccess ||
        cudaStreamCreateWithFlags(&h->copy_out, cudaStreamNonBlocking) != cudaSuccess) { set_error("stream creation failed"); return fail(SSDR_E_CUDA); }
    if (cudaEventCreate(&h->ev_t0) != cudaSuccess || cudaEventCreate(&h->ev_t1) != cudaSuccess) { set_error("event creation failed"); return fail(SSDR_E_CUDA); }
    if (cudaEventCreateWithFlags(&h->ev_copied, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&h->ev_done, cudaEventDisableTiming) != cudaSuccess) { set_error("event creation failed"); return fail(SSDR_E_CUDA); }
    cudaMemset(h->d_chan, 0, sizeof(DemodChan) * (size_t)batch);
    cudaMemset(h->d_taps, 0, sizeof(float) * (size_t)batch * SSDR_FIR_TAPS);
    h->h_taps.assign((size_t)batch * SSDR_FIR_TAPS, 0.0f);
    h->h_work.assign((size_t)batch, 0);
    if (const char* e = std::getenv("SSDR_DEMOD_ENGINE")) {      // developer override of the default engine
        if (!std::strcmp(e, "tcgen05")) h->engine = SSDR_DEMOD_ENGINE_TCGEN05;
        else if (!std::strcmp(e, "ffma")) h->engine = SSDR_DEMOD_ENGINE_FFMA;
        else if (!std::strcmp(e, "auto")) h->engine = SSDR_DEMOD_ENGINE_AUTO;
    }
    *out = h;
    if ((rc = ssdr_demod_reset(h))) { *out = nullptr; return fail(rc); }
    return SSDR_OK;
}

int ssdr_demod_destroy(ssdr_demod_t h) {
    if (!h) return SSDR_OK;
    bind_device(h->device);
    if (h->compute) cudaStreamSynchronize(h->compute);
    if (h->copy_in) cudaStreamSynchronize(h->copy_in);
    if (h->copy_out) cudaStreamSynchronize(h->copy_out);
    cudaFree(h->d_chan); cudaFree(h->d_state); cudaFree(h->d_hist); cudaFree(h->d_taps);
    cudaFree(h->d_in); cudaFree(h->d_f32); cudaFree(h->d_i16); cudaFree(h->d_rssi);
    cudaFree(h->d_quad_ch); cudaFree(h->d_quad_fid); cudaFree(h->d_round_ctr); cudaFree(h->d_sched);
    if (h->ev_t0) cudaEventDestroy(h->ev_t0);
    if (h->ev_t1) cudaEventDestroy(h->ev_t1);
    if (h->ev_copied) cudaEventDestroy(h->ev_copied);
    if (h->ev_done) cudaEventDestroy(h->ev_done);
    if (h->compute) cudaStreamDestroy(h->compute);
    if (h->copy_in) cudaStreamDestroy(h->copy_in);
    if (h->copy_out) cudaStreamDestroy(h->copy_out);
    delete h;
    return SSDR_OK;
}

int ssdr_demod_reset(ssdr_demod_t h) {
    SSDR_ARG(h != nullptr, "null handle");
    SSDR_BIND_H(h);
    SSDR_CUDA(cudaStreamSynchronize(h->compute));
    SSDR_CUDA(cudaMemset(h->d_state, 0, sizeof(DemodState) * (size_t)h->batch));
    SSDR_CUDA(cudaMemset(h->d_hist, 0, sizeof(float2) * (size_t)h->batch * (SSDR_FIR_TAPS - 1)));
    return SSDR_OK;
}

static unsigned phase_inc(double f_hz) {
    double v = std::nearbyint(f_hz / (double)SSDR_KIWI_RATE * 4294967296.0);
    long long iv = (long long)v;
    return (unsigned)((unsigned long long)iv & 0xffffffffull);
}

int ssdr_demod_set(ssdr_demod_t h, int first, int count, const ssdr_demod_params_t* p) {
    SSDR_ARG(h && p, "null argument");
    SSDR_BIND_H(h);
    SSDR_ARG(first >= 0 && count >= 0 && first + count <= h->batch, "channel range [%d, %d) outside batch %d", first, first + count, h->batch);
    std::vector<DemodChan> chan((size_t)count);
    std::vector<float> taps((size_t)count * SSDR_FIR_TAPS);
    for (int i = 0; i < count; ++i) {
        const ssdr_demod_params_t& q = p[i];
        SSDR_ARG(q.mode >= SSDR_MODE_AM && q.mode <= SSDR_MODE_NBFM, "channel %d: unknown mode %d", first + i, q.mode);
        SSDR_ARG(q.high_cut_hz > q.low_cut_hz, "channel %d: high_cut %g <= low_cut %g", first + i, (double)q.high_cut_hz, (double)q.low_cut_hz);
        // the AGC scan forms 2^(+-k c2), k < 512, in float32: a decay under 1 ms would overflow it (inf * 0 = NaN, which
        // would then poison the channel's envelope state).  The reference UI keeps decay in 400..8000 ms (utils_supersdr.py:1009-1019).
        SSDR_ARG(q.agc_decay_ms >= 1.0f && q.agc_decay_ms <= 1.0e6f, "channel %d: AGC decay %g ms outside 1 .. 1e6 ms", first + i, (double)q.agc_decay_ms);
        SSDR_ARG(std::isfinite(q.agc_thresh_dbm) && std::isfinite(q.agc_slope_db) && std::isfinite(q.agc_man_gain_db) &&
                 std::isfinite(q.freq_offset_hz), "channel %d: non-finite AGC / tuning parameter", first + i);
        DemodChan& c = chan[(size_t)i];
        const double fc = ((double)q.low_cut_hz + (double)q.high_cut_hz) / 2.0;
        c.mode = q.mode; c.agc_on = q.agc_on ? 1 : 0; c.agc_hang = q.agc_hang ? 1 : 0;
        c.inc1 = phase_inc((double)q.freq_offset_hz + fc);
        c.inc2 = phase_inc(fc);
        c.c2 = (float)(1.4426950408889634 / ((double)SSDR_KIWI_RATE * (double)q.agc_decay_ms / 1000.0));
        c.knee2 = (float)(((double)q.agc_thresh_dbm - (double)kDemodFsDbm) / 20.0 * 3.321928094887362);
        c.slope_m1 = (float)((double)q.agc_slope_db / 100.0 - 1.0);
        c.man_gain = (float)std::pow(10.0, (double)q.agc_man_gain_db / 20.0);
        std::memcpy(&taps[(size_t)i * SSDR_FIR_TAPS], q.taps, sizeof(float) * SSDR_FIR_TAPS);
    }
    SSDR_CUDA(cudaStreamSynchronize(h->compute));
    SSDR_CUDA(cudaMemcpy(h->d_chan + first, chan.data(), sizeof(DemodChan) * (size_t)count, cudaMemcpyHostToDevice));
    SSDR_CUDA(cudaMemcpy(h->d_taps + (size_t)first * SSDR_FIR_TAPS, taps.data(), sizeof(float) * taps.size(), cudaMemcpyHostToDevice));
    std::memcpy(h->h_taps.data() + (size_t)first * SSDR_FIR_TAPS, taps.data(), sizeof(float) * taps.size());
    for (int i = 0; i < count; ++i) h->h_work[(size_t)(first + i)] = chan[(size_t)i].mode * 4 + chan[(size_t)i].agc_on * 2 + chan[(size_t)i].agc_hang;
    h->quads_dirty = true;
    return SSDR_OK;
}

int ssdr_demod_set_engine(ssdr_demod_t h, int engine) {
    SSDR_ARG(h != nullptr, "null handle");
    SSDR_ARG(engine == SSDR_DEMOD_ENGINE_FFMA || engine == SSDR_DEMOD_ENGINE_TCGEN05 || engine == SSDR_DEMOD_ENGINE_AUTO,
             "unknown demodulator engine %d", engine);
    h->engine = engine;
    return SSDR_OK;
}

// tcgen05 engine: one M = 128 tile is four channels x one frame and all four share the B operand (the taps), so channels
// are grouped by filter (bitwise-equal taps) into quads; a filter with n channels takes ceil(n / 4) quads, the last one
// padded with -1.  `tiles` quads of one filter make a round (padded with empty quads).  Pure host code (no device):
// ssdr_demod_plan exposes it to the CPU tests.  The streaming state stays per channel, so regrouping between calls is free.
static void demod_plan_rounds(const float* t, const int* work, int B, size_t tiles, size_t nsm, bool split_tail,
                              std::vector<int4>& qc, std::vector<int>& qf, double* fill) {
    const size_t tb = sizeof(float) * SSDR_FIR_TAPS;
    const int4 empty = make_int4(-1, -1, -1, -1);
    std::vector<int> order((size_t)B);
    for (int i = 0; i < B; ++i) order[(size_t)i] = i;
    // by filter, then by detector / AGC variant: the four warps of a tile advance in lock step, so a quad of equal cost
    // wastes nothing
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) {
        const int c = std::memcmp(t + (size_t)a * SSDR_FIR_TAPS, t + (size_t)b * SSDR_FIR_TAPS, tb);
        return c != 0 ? c < 0 : work[a] < work[b];
    });
    qc.clear(); qf.clear();
    int fid = -1, fill4 = 4;
    for (int i = 0; i < B; ++i) {
        const int ch = order[(size_t)i];
        const bool same = i > 0 && !std::memcmp(t + (size_t)ch * SSDR_FIR_TAPS, t + (size_t)order[(size_t)i - 1] * SSDR_FIR_TAPS, tb);
        if (!same) {                                         // a new filter starts a new round
            while (qc.size() % tiles) { qc.push_back(empty); qf.push_back(fid); }
            ++fid; fill4 = 4;
        }
        if (fill4 == 4) { qc.push_back(empty); qf.push_back(fid); fill4 = 0; }
        int4& q = qc.back();
        (fill4 == 0 ? q.x : fill4 == 1 ? q.y : fill4 == 2 ? q.z : q.w) = ch;
        ++fill4;
    }
    while (qc.size() % tiles) { qc.push_back(empty); qf.push_back(fid); }
    size_t used = 0;
    for (const int4& q : qc) used += q.x >= 0;
    *fill = used ? (double)B / (4.0 * (double)used) : 0.0;
    // dearest rounds first (measured per mode, profiles/r2final_demod_modes.jsonl: AM 207 < LSB / USB / CW 215 < NBFM 217 Gsamples/s
    // since the detectors were trimmed in round 2), the cheap ones fill the tail
    {
        const size_t nr = qc.size() / tiles;
        auto cost = [&](size_t r) { const int m = work[qc[r * tiles].x] / 4; return m == SSDR_MODE_AM ? 2 : m == SSDR_MODE_NBFM ? 0 : 1; };
        std::vector<size_t> ro(nr);
        for (size_t r = 0; r < nr; ++r) ro[r] = r;
        std::stable_sort(ro.begin(), ro.end(), [&](size_t a, size_t b) { return cost(a) > cost(b); });
        std::vector<int4> qc2(qc.size());
        std::vector<int> qf2(qf.size());
        for (size_t r = 0; r < nr; ++r)
            for (size_t k = 0; k < tiles; ++k) { qc2[r * tiles + k] = qc[ro[r] * tiles + k]; qf2[r * tiles + k] = qf[ro[r] * tiles + k]; }
        qc.swap(qc2); qf.swap(qf2);
    }
    // The last, partial wave: with one CTA per SM, n_rounds = w * n_sm + tail leaves n_sm - tail SMs idle while `tail`
    // CTAs run full rounds.  Spread the quads of those rounds over more, narrower rounds (fewer tiles per CTA finish
    // sooner: the tiles of a CTA share the tensor pipe and the issue slots).
    const size_t nr = qc.size() / tiles, tail = nsm ? nr % nsm : 0;
    if (split_tail && tail > 0 && nr > tail) {
        std::vector<int4> tq;
        std::vector<int> tf;
        for (size_t i = (nr - tail) * tiles; i < qc.size(); ++i)
            if (qc[i].x >= 0) { tq.push_back(qc[i]); tf.push_back(qf[i]); }
        const size_t per = (tq.size() + nsm - 1) / nsm;           // quads per narrow round
        if (per < tiles) {
            std::vector<int4> nq;
            std::vector<int> nf;
            size_t i = 0;
            while (i < tq.size()) {
                size_t k = 0;
                const int f = tf[i];
                for (; k < per && i < tq.size() && tf[i] == f; ++k, ++i) { nq.push_back(tq[i]); nf.push_back(f); }
                for (; k < tiles; ++k) { nq.push_back(empty); nf.push_back(f); }
            }
            // A round cannot mix filters, so with many filters the narrow rounds can outnumber the SMs (27 filters x 16 quads at 3
            // per round: 162 rounds on 148 SMs) -- a third, mostly idle wave instead of the one the split was meant to fill
            // (measured, scripts/demod_hetero.py: 64 channels per filter ran at 171 instead of 215 Gsamples/s).  Split only when the
            // narrow rounds fit one wave.
            if (nq.size() / tiles <= nsm) {
                qc.resize((nr - tail) * tiles); qf.resize((nr - tail) * tiles);
                qc.insert(qc.end(), nq.begin(), nq.end());
                qf.insert(qf.end(), nf.begin(), nf.end());
            }
        }
    }
}

int ssdr_demod_plan(const float* taps, const int32_t* work, int batch, int n_sm, int32_t* quad_ch, int32_t* quad_fid, int cap_quads,
                    int* n_quads, int* tiles_per_round, float* fill) {
    SSDR_ARG(taps && work && n_quads && batch >= 1 && n_sm >= 1, "bad argument");
    std::vector<int4> qc;
    std::vector<int> qf;
    double f = 0.0;
    demod_plan_rounds(taps, work, batch, (size_t)demod_tc_tiles(), (size_t)n_sm, true, qc, qf, &f);
    *n_quads = (int)qc.size();
    if (tiles_per_round) *tiles_per_round = demod_tc_tiles();
    if (fill) *fill = (float)f;
    if (quad_ch && quad_fid) {
        SSDR_ARG((size_t)cap_quads >= qc.size(), "capacity %d < %zu quads", cap_quads, qc.size());
        for (size_t i = 0; i < qc.size(); ++i) {
            quad_ch[4 * i] = qc[i].x; quad_ch[4 * i + 1] = qc[i].y; quad_ch[4 * i + 2] = qc[i].z; quad_ch[4 * i + 3] = qc[i].w;
            quad_fid[i] = qf[i];
        }
    }
    return SSDR_OK;
}

static int demod_build_quads(ssdr_demod_t h) {
    std::vector<int4> qc;
    std::vector<int> qf;
    const size_t tiles = (size_t)demod_tc_tiles();
    demod_plan_rounds(h->h_taps.data(), h->h_work.data(), h->batch, tiles, (size_t)sm_count(),
                      std::getenv("SSDR_DEMOD_NO_TAIL_SPLIT") == nullptr, qc, qf, &h->quad_fill);
    SSDR_CUDA(cudaStreamSynchronize(h->compute));
    cudaFree(h->d_quad_ch); cudaFree(h->d_quad_fid);
    h->d_quad_ch = nullptr; h->d_quad_fid = nullptr;
    int rc;
    if ((rc = dev_alloc(&h->d_quad_ch, qc.size()))) return rc;
    if ((rc = dev_alloc(&h->d_quad_fid, qf.size()))) return rc;
    if (!h->d_round_ctr && (rc = dev_alloc(&h->d_round_ctr, (size_t)1))) return rc;
    SSDR_CUDA(cudaMemcpy(h->d_quad_ch, qc.data(), sizeof(int4) * qc.size(), cudaMemcpyHostToDevice));
    SSDR_CUDA(cudaMemcpy(h->d_quad_fid, qf.data(), sizeof(int) * qf.size(), cudaMemcpyHostToDevice));
    h->n_quads = (int)(qc.size() / tiles);                  // rounds
    h->used_quads = 0;
    for (const int4& q : qc) h->used_quads += q.x >= 0;
    h->quads_dirty = false;
    return SSDR_OK;
}

// AUTO rule.  The tcgen05 engine shares ONE Toeplitz operand per CTA round (four tiles of four channels), so a bank with few
// channels per filter runs rounds with empty tiles and warps.  Measured on a B200 (scripts/demod_hetero.py, 4096 USB channels x
// 32 frames, profiles/r2final_demod_hetero.jsonl): a round of k tiles over F frames costs 25 + F (1.22 + 0.734 k) us of
// one SM (k = 1, 2, 4: 88, 111, 158 us at F = 32), the FFMA engine 0.781 F us of one SM per channel (97 Gsamples/s) -- 2 channels per filter:
// 58 against 97 Gsamples/s (FFMA wins), 4: 111 against 97, 16: 196, one filter: 216.  Tiles at least half full as before; and once
// there are more rounds than SMs (throughput, not latency, decides) the cheaper total wins.
static bool demod_auto_prefers_tc(const ssdr_demod_t h, int n_samples) {
    if (h->quad_fill < 0.5) return false;
    if (h->n_quads <= sm_count()) return true;
    const double F = (double)n_samples / SSDR_FRAME;
    const double tc = (double)h->n_quads * (25.0 + 1.22 * F) + 0.734 * F * (double)h->used_quads;
    const double ff = 0.781 * F * (double)h->batch;
    return tc < ff;
}

static int demod_launch_block(ssdr_demod_t h, const void* iq_dev, int iq_format, int n_samples, int pitch,
                              float* pcm_f32_dev, int16_t* pcm_i16_dev, float* rssi_dev) {
    DemodLaunch a;
    a.iq = iq_dev; a.iq_format = iq_format; a.chan = h->d_chan; a.state = h->d_state; a.hist = h->d_hist; a.taps = h->d_taps;
    a.pcm_f32 = pcm_f32_dev; a.pcm_i16 = pcm_i16_dev; a.rssi = rssi_dev; a.batch = h->batch; a.n_samples = n_samples; a.pitch = pitch;
    for (int s = 0; s < 5; ++s) a.am_pow16[s] = h->am_pow16[s];
    if (h->engine != SSDR_DEMOD_ENGINE_FFMA && h->quads_dirty) { int rc = demod_build_quads(h); if (rc) return rc; }
    if (h->engine == SSDR_DEMOD_ENGINE_TCGEN05 || (h->engine == SSDR_DEMOD_ENGINE_AUTO && demod_auto_prefers_tc(h, n_samples))) {
        return demod_tc_launch(a, h->d_quad_ch, h->d_quad_fid, h->n_quads, h->d_round_ctr, h->compute);
    }
    if (!h->d_sched) { int rc = dev_alloc(&h->d_sched, (size_t)h->batch + 1); if (rc) return rc; }
    a.sched = h->d_sched;
    return demod_launch(a, h->compute);
}

int ssdr_demod_process_dev(ssdr_demod_t h, const void* iq_dev, int iq_format, int n_samples, float* pcm_f32_dev,
                           int16_t* pcm_i16_dev, float* rssi_dev) {
    SSDR_ARG(h && iq_dev, "null argument");
    SSDR_BIND_H(h);
    SSDR_ARG(iq_format == SSDR_IQ_CF32 || iq_format == SSDR_IQ_S16BE, "bad iq_format %d", iq_format);
    SSDR_ARG(n_samples > 0 && n_samples % SSDR_FRAME == 0, "n_samples %d must be a positive multiple of %d", n_samples, SSDR_FRAME);
    // the kernels move IQ and PCM as 16-byte vectors
    SSDR_ARG(aligned16(iq_dev) && aligned16(pcm_f32_dev) && aligned16(pcm_i16_dev) && aligned16(rssi_dev),
             "device pointers of ssdr_demod_process_dev must be 16-byte aligned");
    return demod_launch_block(h, iq_dev, iq_format, n_samples, n_samples, pcm_f32_dev, pcm_i16_dev, rssi_dev);
}

// Host-buffer entry.  The call is cut into TIME blocks (all channels x a few frames): the per-channel streaming state
// carries from block to block exactly as it does from call to call, every kernel covers the whole batch, and the strided
// H2D copy of block k+1 (copy-in stream), the kernel of block k (compute stream) and the D2H copies of block k-1
// (copy-out stream) overlap -- PCIe is full duplex.
int ssdr_demod_process(ssdr_demod_t h, const void* iq_host, int iq_format, int n_samples, float* pcm_f32, int16_t* pcm_i16,
                       float* rssi_dbm) {
    SSDR_ARG(h && iq_host, "null argument");
    SSDR_BIND_H(h);
    SSDR_ARG(iq_format == SSDR_IQ_CF32 || iq_format == SSDR_IQ_S16BE, "bad iq_format %d", iq_format);
    SSDR_ARG(n_samples > 0 && n_samples % SSDR_FRAME == 0 && n_samples <= h->max_samples,
             "n_samples %d must be a multiple of %d and <= %d", n_samples, SSDR_FRAME, h->max_samples);
    const size_t cap = (size_t)h->batch * h->max_samples, sb = iq_sample_bytes(iq_format);
    int rc;
    if (!h->d_in) {
        h->in_bytes = cap * 8;
        if ((rc = dev_alloc(reinterpret_cast<unsigned char**>(&h->d_in), h->in_bytes))) return rc;
    }
    if (pcm_f32 && !h->d_f32 && (rc = dev_alloc(&h->d_f32, cap))) return rc;
    if (pcm_i16 && !h->d_i16 && (rc = dev_alloc(&h->d_i16, cap))) return rc;
    if (rssi_dbm && !h->d_rssi && (rc = dev_alloc(&h->d_rssi, cap / SSDR_FRAME))) return rc;
    // block length: ~64 MiB of input per block, whole frames, at least one frame
    const int nblk_total = n_samples / SSDR_FRAME;
    int fpb = (int)std::max<size_t>(1, ((size_t)64 << 20) / ((size_t)h->batch * SSDR_FRAME * sb));
    fpb = std::min(fpb, nblk_total);
    const size_t pitch_b = (size_t)n_samples;                     // device staging keeps the caller's [batch][n_samples] layout
    unsigned char* d_iq = static_cast<unsigned char*>(h->d_in);
    for (int f0 = 0; f0 < nblk_total; f0 += fpb) {
        const int nf = std::min(fpb, nblk_total - f0);
        const size_t s0 = (size_t)f0 * SSDR_FRAME, ns = (size_t)nf * SSDR_FRAME;
        SSDR_CUDA(cudaMemcpy2DAsync(d_iq + s0 * sb, pitch_b * sb, static_cast<const unsigned char*>(iq_host) + s0 * sb, pitch_b * sb,
                                    ns * sb, (size_t)h->batch, cudaMemcpyHostToDevice, h->copy_in));
        SSDR_CUDA(cudaEventRecord(h->ev_copied, h->copy_in));
        SSDR_CUDA(cudaStreamWaitEvent(h->compute, h->ev_copied, 0));
        if ((rc = demod_launch_block(h, d_iq + s0 * sb, iq_format, (int)ns, n_samples, pcm_f32 ? h->d_f32 + s0 : nullptr,
                                     pcm_i16 ? h->d_i16 + s0 : nullptr, rssi_dbm ? h->d_rssi + f0 : nullptr))) return rc;
        SSDR_CUDA(cudaEventRecord(h->ev_done, h->compute));
        SSDR_CUDA(cudaStreamWaitEvent(h->copy_out, h->ev_done, 0));
        if (pcm_f32) SSDR_CUDA(cudaMemcpy2DAsync(pcm_f32 + s0, pitch_b * sizeof(float), h->d_f32 + s0, pitch_b * sizeof(float),
                                                 ns * sizeof(float), (size_t)h->batch, cudaMemcpyDeviceToHost, h->copy_out));
        if (pcm_i16) SSDR_CUDA(cudaMemcpy2DAsync(pcm_i16 + s0, pitch_b * sizeof(int16_t), h->d_i16 + s0, pitch_b * sizeof(int16_t),
                                                 ns * sizeof(int16_t), (size_t)h->batch, cudaMemcpyDeviceToHost, h->copy_out));
    }
    SSDR_CUDA(cudaStreamSynchronize(h->compute));
    if (rssi_dbm) SSDR_CUDA(cudaMemcpyAsync(rssi_dbm, h->d_rssi, (size_t)h->batch * nblk_total * sizeof(float), cudaMemcpyDeviceToHost, h->copy_out));
    SSDR_CUDA(cudaStreamSynchronize(h->copy_out));
    SSDR_CUDA(cudaStreamSynchronize(h->copy_in));
    return SSDR_OK;
}

int ssdr_demod_sync(ssdr_demod_t h) {
    SSDR_ARG(h != nullptr, "null handle");
    SSDR_BIND_H(h);
    SSDR_CUDA(cudaStreamSynchronize(h->compute));
    return SSDR_OK;
}

int ssdr_demod_time_dev(ssdr_demod_t h, const void* iq_dev, int iq_format, int n_samples, float* pcm_f32_dev,
                        int16_t* pcm_i16_dev, int iters, float* total_ms) {
    SSDR_ARG(h && iq_dev && total_ms && iters >= 1, "bad argument");
    SSDR_BIND_H(h);
    SSDR_CUDA(cudaEventRecord(h->ev_t0, h->compute));
    for (int i = 0; i < iters; ++i) {
        int rc = ssdr_demod_process_dev(h, iq_dev, iq_format, n_samples, pcm_f32_dev, pcm_i16_dev, nullptr);
        if (rc) return rc;
    }
    SSDR_CUDA(cudaEventRecord(h->ev_t1, h->compute));
    SSDR_CUDA(cudaEventSynchronize(h->ev_t1));
    SSDR_CUDA(cudaEventElapsedTime(total_ms, h->ev_t0, h->ev_t1));
    return SSDR_OK;
}

// =============================================================================================
// interpolator
// =============================================================================================
int ssdr_interp_create(ssdr_interp_t* out, int batch, int ratio, const double* taps, int n_taps, int max_samples) {
    SSDR_ARG(out != nullptr, "null handle pointer");
    *out = nullptr;
    SSDR_ARG(batch >= 1 && ratio >= 1 && taps && max_samples >= 1, "bad argument");
    SSDR_ARG(batch <= 65535, "batch %d > 65535 channels per interpolator handle (grid y dimension)", batch);
    SSDR_ARG(n_taps >= 1 && n_taps <= SSDR_INTERP_TAPS_MAX && (n_taps - 1) % ratio == 0,
             "n_taps %d must be <= %d with (n_taps - 1) a multiple of the ratio %d", n_taps, SSDR_INTERP_TAPS_MAX, ratio);
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { set_error("no CUDA device (libssdr_b200 has no CPU fallback)"); return SSDR_E_CUDA; }
    SSDR_BIND_G();
    ssdr_interp* h = new ssdr_interp();
    if (cudaGetDevice(&h->device) != cudaSuccess) h->device = -1;
    h->batch = batch; h->ratio = ratio; h->n_taps = n_taps; h->max_samples = max_samples; h->hs = (n_taps - 1) / ratio;
    int rc;
    auto fail = [&](int code) { ssdr_interp_destroy(h); return code; };
    if ((rc = dev_alloc(&h->d_taps, (size_t)n_taps))) return fail(rc);
    for (int i = 0; i < 2; ++i) if ((rc = dev_alloc(&h->d_hist[i], (size_t)batch * std::max(h->hs, 1)))) return fail(rc);
    if ((rc = dev_alloc(&h->d_vol, (size_t)batch))) return fail(rc);
    if ((rc = dev_alloc(&h->d_bal, (size_t)batch))) return fail(rc);
    if (cudaMemcpy(h->d_taps, taps, sizeof(double) * (size_t)n_taps, cudaMemcpyHostToDevice) != cudaSuccess) { set_error("tap upload failed"); return fail(SSDR_E_CUDA); }
    if (cudaStreamCreateWithFlags(&h->compute, cudaStreamNonBlocking) != cudaSuccess) { set_error("stream creation failed"); return fail(SSDR_E_CUDA); }
    *out = h;
    if ((rc = ssdr_interp_reset(h))) { *out = nullptr; return fail(rc); }
    return SSDR_OK;
}

int ssdr_interp_destroy(ssdr_interp_t h) {
    if (!h) return SSDR_OK;
    bind_device(h->device);
    if (h->compute) cudaStreamSynchronize(h->compute);
    cudaFree(h->d_taps); cudaFree(h->d_hist[0]); cudaFree(h->d_hist[1]); cudaFree(h->d_in); cudaFree(h->d_vol);
    cudaFree(h->d_bal); cudaFree(h->d_out); cudaFree(h->d_mono);
    if (h->compute) cudaStreamDestroy(h->compute);
    delete h;
    return SSDR_OK;
}

int ssdr_interp_reset(ssdr_interp_t h) {
    SSDR_ARG(h != nullptr, "null handle");
    SSDR_BIND_H(h);
    SSDR_CUDA(cudaStreamSynchronize(h->compute));
    for (int i = 0; i < 2; ++i) SSDR_CUDA(cudaMemset(h->d_hist[i], 0, sizeof(double) * (size_t)h->batch * std::max(h->hs, 1)));
    return SSDR_OK;
}

int ssdr_interp_process_dev(ssdr_interp_t h, const int16_t* pcm_dev, int n, const float* volume_dev, const float* balance_dev,
                            int16_t* stereo_dev, double* mono_dev) {
    SSDR_ARG(h && pcm_dev && volume_dev && balance_dev && stereo_dev, "null argument");
    SSDR_BIND_H(h);
    SSDR_ARG(n >= h->hs, "n %d shorter than the filter history %d", n, h->hs);
    SSDR_ARG(((uintptr_t)stereo_dev & 3u) == 0 && ((uintptr_t)pcm_dev & 1u) == 0 && (!mono_dev || ((uintptr_t)mono_dev & 7u) == 0),
             "ssdr_interp_process_dev: stereo_dev must be 4-byte, pcm_dev 2-byte, mono_dev 8-byte aligned");
    InterpLaunch a;
    a.kp.pcm = pcm_dev; a.kp.volume = volume_dev; a.kp.balance = balance_dev; a.kp.taps = h->d_taps;
    a.kp.hist_in = h->d_hist[h->cur]; a.kp.hist_out = h->d_hist[h->cur ^ 1];
    a.kp.stereo = stereo_dev; a.kp.mono = mono_dev; a.kp.n = n; a.kp.ratio = h->ratio; a.kp.n_taps = h->n_taps;
    a.batch = h->batch;
    int rc = interp_launch(a, h->compute);
    if (rc) return rc;
    h->cur ^= 1;
    return SSDR_OK;
}

int ssdr_interp_process(ssdr_interp_t h, const int16_t* pcm_host, int n, const float* volume, const float* balance,
                        int16_t* stereo_out, double* mono_f64) {
    SSDR_ARG(h && pcm_host && volume && balance && stereo_out, "null argument");
    SSDR_BIND_H(h);
    SSDR_ARG(n >= 1 && n <= h->max_samples, "n %d outside 1..%d", n, h->max_samples);
    const size_t cap = (size_t)h->batch * h->max_samples, tot = (size_t)h->batch * n;
    int rc;
    if (!h->d_in && (rc = dev_alloc(&h->d_in, cap))) return rc;
    if (!h->d_out && (rc = dev_alloc(&h->d_out, cap * h->ratio * 2))) return rc;
    if (mono_f64 && !h->d_mono && (rc = dev_alloc(&h->d_mono, cap * h->ratio))) return rc;
    SSDR_CUDA(cudaMemcpyAsync(h->d_in, pcm_host, tot * sizeof(int16_t), cudaMemcpyHostToDevice, h->compute));
    SSDR_CUDA(cudaMemcpyAsync(h->d_vol, volume, sizeof(float) * (size_t)h->batch, cudaMemcpyHostToDevice, h->compute));
    SSDR_CUDA(cudaMemcpyAsync(h->d_bal, balance, sizeof(float) * (size_t)h->batch, cudaMemcpyHostToDevice, h->compute));
    if ((rc = ssdr_interp_process_dev(h, h->d_in, n, h->d_vol, h->d_bal, h->d_out, mono_f64 ? h->d_mono : nullptr))) return rc;
    SSDR_CUDA(cudaMemcpyAsync(stereo_out, h->d_out, tot * h->ratio * 2 * sizeof(int16_t), cudaMemcpyDeviceToHost, h->compute));
    if (mono_f64) SSDR_CUDA(cudaMemcpyAsync(mono_f64, h->d_mono, tot * h->ratio * sizeof(double), cudaMemcpyDeviceToHost, h->compute));
    SSDR_CUDA(cudaStreamSynchronize(h->compute));
    return SSDR_OK;
}

int ssdr_interp_sync(ssdr_interp_t h) {
    SSDR_ARG(h != nullptr, "null handle");
    SSDR_BIND_H(h);
    SSDR_CUDA(cudaStreamSynchronize(h->compute));
    return SSDR_OK;
}

int ssdr_resample_line(const int16_t* pcm_host, int batch, int n, const float* volume, const float* balance, const double* h,
                       int n_h, int up, int down, int first, int n_keep, int16_t* stereo_out, double* mono_f64) {
    SSDR_ARG(pcm_host && volume && balance && h && stereo_out, "null argument");
    SSDR_BIND_G();
    SSDR_ARG(batch >= 1 && n >= 1 && n_h >= 1 && up >= 1 && down >= 1 && first >= 0 && n_keep >= 0, "bad argument");
    SSDR_ARG(batch <= 65535, "batch %d > 65535 channels per call (grid y dimension)", batch);
    // the kept samples must exist: upfirdn produces ((n - 1) up + n_h - 1) / down + 1 samples
    SSDR_ARG((long long)(first + n_keep) <= ((long long)(n - 1) * up + n_h - 1) / down + 1, "first + n_keep exceeds the upfirdn output length");
    if (n_keep == 0) return SSDR_OK;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { set_error("no CUDA device (libssdr_b200 has no CPU fallback)"); return SSDR_E_CUDA; }
    int16_t *d_x = nullptr, *d_o = nullptr;
    float *d_v = nullptr, *d_b = nullptr;
    double *d_h = nullptr, *d_m = nullptr;
    const size_t nin = (size_t)batch * n, nout = (size_t)batch * n_keep;
    int rc = SSDR_OK;
    auto cleanup = [&]() { cudaFree(d_x); cudaFree(d_o); cudaFree(d_v); cudaFree(d_b); cudaFree(d_h); cudaFree(d_m); };
    if ((rc = dev_alloc(&d_x, nin)) || (rc = dev_alloc(&d_o, nout * 2)) || (rc = dev_alloc(&d_v, (size_t)batch)) ||
        (rc = dev_alloc(&d_b, (size_t)batch)) || (rc = dev_alloc(&d_h, (size_t)n_h)) || (mono_f64 && (rc = dev_alloc(&d_m, nout)))) { cleanup(); return rc; }
    cudaError_t e = cudaMemcpy(d_x, pcm_host, nin * sizeof(int16_t), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(d_v, volume, sizeof(float) * batch, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(d_b, balance, sizeof(float) * batch, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(d_h, h, sizeof(double) * n_h, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) {
        ResampleKernelParams kp;
        kp.pcm = d_x; kp.volume = d_v; kp.balance = d_b; kp.h = d_h; kp.stereo = d_o; kp.mono = d_m;
        kp.n = n; kp.n_h = n_h; kp.up = up; kp.down = down; kp.first = first; kp.n_keep = n_keep;
        rc = resample_line_launch(kp, batch, 0);
        if (!rc) e = cudaMemcpy(stereo_out, d_o, nout * 2 * sizeof(int16_t), cudaMemcpyDeviceToHost);
        if (!rc && e == cudaSuccess && mono_f64) e = cudaMemcpy(mono_f64, d_m, nout * sizeof(double), cudaMemcpyDeviceToHost);
    }
    cleanup();
    if (rc) return rc;
    if (e != cudaSuccess) return cuda_fail(e, "resample copy", __FILE__, __LINE__);
    return SSDR_OK;
}

int ssdr_fir_valid_f64(const double* x_host, size_t n, const double* taps, int n_taps, double* out_host) {
    SSDR_ARG(x_host && taps && out_host && n_taps >= 1, "bad argument");
    SSDR_BIND_G();
    if (n < (size_t)n_taps) return SSDR_OK;
    const size_t n_out = n - (size_t)n_taps + 1;
    double *d_x = nullptr, *d_h = nullptr, *d_o = nullptr;
    int rc;
    if ((rc = dev_alloc(&d_x, n)) || (rc = dev_alloc(&d_h, (size_t)n_taps)) || (rc = dev_alloc(&d_o, n_out))) {
        cudaFree(d_x); cudaFree(d_h); cudaFree(d_o);
        return rc;
    }
    cudaError_t e = cudaMemcpy(d_x, x_host, n * sizeof(double), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(d_h, taps, (size_t)n_taps * sizeof(double), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) { rc = fir_valid_launch(d_x, d_h, n_taps, d_o, n_out, 0); if (!rc) e = cudaMemcpy(out_host, d_o, n_out * sizeof(double), cudaMemcpyDeviceToHost); }
    cudaFree(d_x); cudaFree(d_h); cudaFree(d_o);
    if (rc) return rc;
    if (e != cudaSuccess) return cuda_fail(e, "fir_valid copy", __FILE__, __LINE__);
    return SSDR_OK;
}

// =============================================================================================
// display epilogues
// =============================================================================================
static const int kWfDelay = 3;        // kiwi_waterfall.wf_buffer_len (utils_supersdr.py:604)

int ssdr_wf_image_create(ssdr_wf_image_t* out, int batch, int height, int width, const uint8_t* palette_rgb) {
    SSDR_ARG(out != nullptr, "null handle pointer");
    *out = nullptr;
    SSDR_ARG(batch >= 1 && height >= 1 && width >= 1 && palette_rgb, "bad argument");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { set_error("no CUDA device (libssdr_b200 has no CPU fallback)"); return SSDR_E_CUDA; }
    SSDR_BIND_G();
    ssdr_wf_image* h = new ssdr_wf_image();
    if (cudaGetDevice(&h->device) != cudaSuccess) h->device = -1;
    h->batch = batch; h->H = height; h->W = width;
    int rc;
    auto fail = [&](int code) { ssdr_wf_image_destroy(h); return code; };
    if ((rc = dev_alloc(&h->d_ring, (size_t)batch * height * width))) return fail(rc);
    if ((rc = dev_alloc(&h->d_delay, (size_t)kWfDelay * batch * width))) return fail(rc);
    if ((rc = dev_alloc(&h->d_row, (size_t)width))) return fail(rc);
    if ((rc = dev_alloc(&h->d_pal, (size_t)768))) return fail(rc);
    std::vector<float> white((size_t)width, 255.0f);
    if (cudaMemset(h->d_ring, 0, sizeof(float) * (size_t)batch * height * width) != cudaSuccess ||
        cudaMemcpy(h->d_row, white.data(), sizeof(float) * width, cudaMemcpyHostToDevice) != cudaSuccess ||
        cudaMemcpy(h->d_pal, palette_rgb, 768, cudaMemcpyHostToDevice) != cudaSuccess ||
        cudaStreamCreateWithFlags(&h->st, cudaStreamNonBlocking) != cudaSuccess) { set_error("image setup failed"); return fail(SSDR_E_CUDA); }
    *out = h;
    return SSDR_OK;
}

int ssdr_wf_image_destroy(ssdr_wf_image_t h) {
    if (!h) return SSDR_OK;
    bind_device(h->device);
    if (h->st) cudaStreamSynchronize(h->st);
    cudaFree(h->d_ring); cudaFree(h->d_delay); cudaFree(h->d_row); cudaFree(h->d_pal); cudaFree(h->d_rgb); cudaFree(h->d_f64); cudaFree(h->d_y);
    if (h->st) cudaStreamDestroy(h->st);
    delete h;
    return SSDR_OK;
}

static int wf_image_push_any(ssdr_wf_image_t h, const float* src, cudaMemcpyKind kind) {
    const size_t row = (size_t)h->batch * h->W;
    // deque(maxlen = 3).appendleft: a full deque silently drops its oldest (rightmost) entry first
    int slot;
    if ((int)h->deque.size() == kWfDelay) { slot = h->deque.back(); h->deque.pop_back(); }
    else { bool used[kWfDelay] = {false, false, false}; for (int s : h->deque) used[s] = true; slot = 0; while (used[slot]) ++slot; }
    SSDR_CUDA(cudaMemcpyAsync(h->d_delay + (size_t)slot * row, src, row * sizeof(float), kind, h->st));
    h->deque.insert(h->deque.begin(), slot);
    h->run_index++;
    if (h->run_index > kWfDelay) {                       // scroll one line down, top line <- oldest delayed row
        const int s = h->deque.back(); h->deque.pop_back();
        h->head = (h->head + h->H - 1) % h->H;
        SSDR_CUDA(cudaMemcpy2DAsync(h->d_ring + (size_t)h->head * h->W, sizeof(float) * (size_t)h->H * h->W,
                                    h->d_delay + (size_t)s * row, sizeof(float) * h->W, sizeof(float) * h->W,
                                    (size_t)h->batch, cudaMemcpyDeviceToDevice, h->st));
    }
    SSDR_CUDA(cudaStreamSynchronize(h->st));
    return SSDR_OK;
}

int ssdr_wf_image_push(ssdr_wf_image_t h, const float* colour_host) {
    SSDR_ARG(h && colour_host, "null argument");
    SSDR_BIND_H(h);
    return wf_image_push_any(h, colour_host, cudaMemcpyHostToDevice);
}

int ssdr_wf_image_push_dev(ssdr_wf_image_t h, const float* colour_dev) {
    SSDR_ARG(h && colour_dev, "null argument");
    SSDR_BIND_H(h);
    return wf_image_push_any(h, colour_dev, cudaMemcpyDeviceToDevice);
}

int ssdr_wf_image_white(ssdr_wf_image_t h) {
    SSDR_ARG(h != nullptr, "null handle");
    SSDR_BIND_H(h);
    for (int ch = 0; ch < h->batch; ++ch)
        SSDR_CUDA(cudaMemcpyAsync(h->d_ring + ((size_t)ch * h->H + h->head) * h->W, h->d_row, sizeof(float) * h->W,
                                  cudaMemcpyDeviceToDevice, h->st));
    SSDR_CUDA(cudaStreamSynchronize(h->st));
    return SSDR_OK;
}

int ssdr_wf_image_get(ssdr_wf_image_t h, uint8_t* rgb, double* wf_data) {
    SSDR_ARG(h != nullptr, "null handle");
    SSDR_BIND_H(h);
    const size_t px = (size_t)h->batch * h->H * h->W;
    int rc;
    if (rgb) {
        if (!h->d_rgb && (rc = dev_alloc(&h->d_rgb, px * 3))) return rc;
        if ((rc = image_rgb_launch(h->d_ring, h->d_pal, h->d_rgb, h->batch, h->H, h->W, h->head, h->st))) return rc;
        SSDR_CUDA(cudaMemcpyAsync(rgb, h->d_rgb, px * 3, cudaMemcpyDeviceToHost, h->st));
    }
    if (wf_data) {
        if (!h->d_f64 && (rc = dev_alloc(&h->d_f64, px))) return rc;
        if ((rc = image_data_launch(h->d_ring, h->d_f64, h->batch, h->H, h->W, h->head, h->st))) return rc;
        SSDR_CUDA(cudaMemcpyAsync(wf_data, h->d_f64, px * sizeof(double), cudaMemcpyDeviceToHost, h->st));
    }
    SSDR_CUDA(cudaStreamSynchronize(h->st));
    return SSDR_OK;
}

int ssdr_wf_image_trace(ssdr_wf_image_t h, int t_avg, int spectrum_height, double* v, int32_t* y) {
    SSDR_ARG(h != nullptr && t_avg >= 1 && spectrum_height >= 1, "bad argument");
    SSDR_BIND_H(h);
    const size_t n = (size_t)h->batch * h->W;
    int rc;
    if (!h->d_f64 && (rc = dev_alloc(&h->d_f64, (size_t)h->batch * h->H * h->W))) return rc;
    if (!h->d_y && (rc = dev_alloc(&h->d_y, n))) return rc;
    if ((rc = image_trace_launch(h->d_ring, h->d_f64, h->d_y, h->batch, h->H, h->W, h->head, t_avg, spectrum_height, h->st))) return rc;
    if (v) SSDR_CUDA(cudaMemcpyAsync(v, h->d_f64, n * sizeof(double), cudaMemcpyDeviceToHost, h->st));
    if (y) SSDR_CUDA(cudaMemcpyAsync(y, h->d_y, n * sizeof(int), cudaMemcpyDeviceToHost, h->st));
    SSDR_CUDA(cudaStreamSynchronize(h->st));
    return SSDR_OK;
}

int ssdr_adpcm_decode(const uint8_t* data_host, int batch, int n_bytes, int32_t* state, int16_t* pcm_out) {
    SSDR_ARG(data_host && state && pcm_out && batch >= 1 && n_bytes >= 0, "bad argument");
    SSDR_BIND_G();
    if (n_bytes == 0) return SSDR_OK;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { set_error("no CUDA device (libssdr_b200 has no CPU fallback)"); return SSDR_E_CUDA; }
    uint8_t* d_in = nullptr; int* d_st = nullptr; int16_t* d_out = nullptr;
    const size_t nin = (size_t)batch * n_bytes;
    int rc;
    auto cleanup = [&]() { cudaFree(d_in); cudaFree(d_st); cudaFree(d_out); };
    if ((rc = dev_alloc(&d_in, nin)) || (rc = dev_alloc(&d_st, (size_t)batch * 2)) || (rc = dev_alloc(&d_out, nin * 2))) { cleanup(); return rc; }
    cudaError_t e = cudaMemcpy(d_in, data_host, nin, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(d_st, state, sizeof(int) * 2 * batch, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) {
        rc = adpcm_launch(d_in, batch, n_bytes, d_st, d_out, 0);
        if (!rc) e = cudaMemcpy(pcm_out, d_out, nin * 2 * sizeof(int16_t), cudaMemcpyDeviceToHost);
        if (!rc && e == cudaSuccess) e = cudaMemcpy(state, d_st, sizeof(int) * 2 * batch, cudaMemcpyDeviceToHost);
    }
    cleanup();
    if (rc) return rc;
    if (e != cudaSuccess) return cuda_fail(e, "adpcm copy", __FILE__, __LINE__);
    return SSDR_OK;
}

// =============================================================================================
// IQ unpack
// =============================================================================================
int ssdr_unpack_iq_s16be_dev(const void* s16be_dev, float* cf32_dev, size_t n_complex) {
    SSDR_ARG(s16be_dev && cf32_dev, "null argument");
    SSDR_BIND_G();
    int rc = unpack_launch(s16be_dev, cf32_dev, n_complex, 0);
    if (rc) return rc;
    SSDR_CUDA(cudaStreamSynchronize(0));
    return SSDR_OK;
}

int ssdr_unpack_iq_s16be(const void* s16be_host, float* cf32_host, size_t n_complex) {
    SSDR_ARG(s16be_host && cf32_host, "null argument");
    SSDR_BIND_G();
    if (n_complex == 0) return SSDR_OK;
    unsigned char* d_in = nullptr;
    float* d_out = nullptr;
    int rc;
    if ((rc = dev_alloc(&d_in, n_complex * 4))) return rc;
    if ((rc = dev_alloc(&d_out, n_complex * 2))) { cudaFree(d_in); return rc; }
    cudaError_t e = cudaMemcpy(d_in, s16be_host, n_complex * 4, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) { rc = unpack_launch(d_in, d_out, n_complex, 0); if (!rc) e = cudaMemcpy(cf32_host, d_out, n_complex * 8, cudaMemcpyDeviceToHost); }
    cudaFree(d_in); cudaFree(d_out);
    if (rc) return rc;
    if (e != cudaSuccess) return cuda_fail(e, "unpack copy", __FILE__, __LINE__);
    return SSDR_OK;
}

}  // extern "C"
