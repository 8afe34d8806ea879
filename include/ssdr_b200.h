/*
 * ssdr_b200.h -- C ABI of libssdr_b200.so, the B200-native (sm_100a) implementation of the
 * IQ-sample DSP behind SuperSDR's waterfall and audio classes.
 *
 * The reference (mcogoni/supersdr) has no FFI/plugin layer: supersdr.py pokes the attributes of
 * two Python classes.  The boundary is therefore this C ABI, called
 * through ctypes by the Python classes in supersdr_b200/ that keep the reference's duck-typed
 * surface (kiwi_waterfall utils_supersdr.py:592-898, kiwi_sound utils_supersdr.py:901-1186,
 * filtering utils_supersdr.py:333-348).  Each entry point cites the reference code it replaces.
 * INTEGRATION.md shows the binding a maintainer of the reference would add.
 *
 * Conventions: every function returns 0 on success or a negative SSDR_E_* code and records a
 * message retrievable with ssdr_last_error() (thread-local).  Handles are opaque, own their device
 * buffers, per-channel state and one CUDA stream; a handle may be used from any thread but not
 * concurrently (every entry point binds the calling thread to the handle's device -- the device current when the
 * handle was created, i.e. the one ssdr_init selected -- so this holds on every GPU, not only on device 0).
 * Device pointers passed to the *_dev entry points must be 16-byte aligned (ssdr_dev_alloc returns 256-byte
 * aligned memory; offset sub-buffers must keep 16-byte alignment), else SSDR_E_ARG.  "host" pointers are ordinary (ideally pinned) host memory, "dev" pointers are
 * device memory obtained from ssdr_dev_alloc() or owned by the caller.  There is NO CPU fallback:
 * without a CUDA device every compute call fails with SSDR_E_CUDA.
 */
#ifndef SSDR_B200_H
#define SSDR_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SSDR_ABI_VERSION 1

/* error codes */
#define SSDR_OK          0
#define SSDR_E_ARG      -1   /* bad argument / unsupported size */
#define SSDR_E_CUDA     -2   /* CUDA runtime error (incl. no device) */
#define SSDR_E_STATE    -3   /* call sequence error */
#define SSDR_E_NOMEM    -4

/* IQ sample formats */
#define SSDR_IQ_CF32     0   /* interleaved float32 I,Q in int16-count units (complex64)          */
#define SSDR_IQ_S16BE    1   /* Kiwi wire format: big-endian int16 I,Q pairs, kiwi/client.py:449-453 */

/* demodulator modes: utils_supersdr.py:859-873 (AM/USB/LSB/CW) + NBFM kiwi/client.py:237-239 */
#define SSDR_MODE_AM     0
#define SSDR_MODE_USB    1
#define SSDR_MODE_LSB    2
#define SSDR_MODE_CW     3
#define SSDR_MODE_NBFM   4

/* fixed constants of the builder-defined (Tier U) spec, DESIGN.md section 4 */
#define SSDR_FS            32768.0f
#define SSDR_WF_CAL_DB     (-10.0)
#define SSDR_KIWI_RATE     12000      /* utils_supersdr.py:906 */
#define SSDR_FRAME         512        /* KIWI_SAMPLES_PER_FRAME utils_supersdr.py:909 */
#define SSDR_FIR_TAPS      127
#define SSDR_HANG_BLOCKS   11
#define SSDR_INTERP_TAPS_MAX 64

typedef struct ssdr_wf*     ssdr_wf_t;
typedef struct ssdr_demod*  ssdr_demod_t;
typedef struct ssdr_interp* ssdr_interp_t;

/* ---------------------------------------------------------------------------------------------
 * library / device
 * ------------------------------------------------------------------------------------------- */
int         ssdr_abi_version(void);
const char* ssdr_last_error(void);
/* Select the CUDA device for this process (one process per GPU). */
int ssdr_init(int device);
int ssdr_device_info(int* sm_count, int* cc_major, int* cc_minor, size_t* hbm_bytes, char* name, int name_len);
/* PCI bus id ("0000:1b:00.0") of the selected device: lets a host process place its pinned buffers and threads on the
 * GPU's NUMA node (/sys/bus/pci/devices/<id>/numa_node) -- supersdr_b200.numa_bind(). */
int ssdr_device_pci_bus_id(char* bus_id, int len);
/* Number of kernels this library has launched since load (all handles). */
uint64_t ssdr_launch_count(void);

/* device / pinned-host memory for callers that keep data resident (bench, multi-GPU scatter) */
int ssdr_dev_alloc(void** dev, size_t bytes);
int ssdr_dev_free(void* dev);
int ssdr_host_alloc(void** host, size_t bytes);          /* pinned */
int ssdr_host_free(void* host);
int ssdr_memcpy_h2d(void* dev, const void* host, size_t bytes);
int ssdr_memcpy_d2h(void* host, const void* dev, size_t bytes);
int ssdr_dev_memset(void* dev, int value, size_t bytes);
int ssdr_device_sync(void);

/* Peer ingest (multi-GPU, one process per GPU): export a device allocation of this process so that another rank's
 * kernels can read it in place over NVLink -- the root-ingest deployment without a staging copy: every rank's waterfall
 * kernel loads its shard straight from the root's HBM, the transfer overlapping the butterflies.
 * handle64: 64 opaque bytes (cudaIpcMemHandle_t) to pass to the other process by any means. */
int ssdr_ipc_export(void* dev, void* handle64);
int ssdr_ipc_open(const void* handle64, void** dev);     /* maps the peer allocation; enables peer access */
int ssdr_ipc_close(void* dev);

/* ---------------------------------------------------------------------------------------------
 * multi-GPU plumbing (one process per GPU): NCCL over NVLink, used only to move INPUT shards from a root rank to the
 * ranks that own the channels, and the small pixel rows back -- channels are independent, there is no collective in the
 * math.  The reference is a single-receiver client and has no counterpart; the shard of a rank is a
 * contiguous block of channels, i.e. of the [batch][n_avg][nfft] input of ssdr_wf_process_dev.  libnccl is loaded with
 * dlopen at the first call (ssdr_nccl_available() == 0 where it is missing).  All calls are asynchronous on the
 * communicator's stream; ssdr_nccl_sync waits.
 * ------------------------------------------------------------------------------------------- */
typedef struct ssdr_comm* ssdr_comm_t;
#define SSDR_NCCL_ID_BYTES 128
int ssdr_nccl_available(void);                        /* NCCL version code (> 0) or 0 */
int ssdr_nccl_unique_id(void* id128);                 /* rank 0: 128 opaque bytes to hand to every rank by any means */
int ssdr_nccl_init(ssdr_comm_t* c, const void* id128, int rank, int world);      /* collective over the world */
int ssdr_nccl_destroy(ssdr_comm_t c);
/* The root's buffer holds rank r's shard at byte offsets[r], counts[r] bytes (arrays of `world` entries, the same on
 * every rank); every rank receives its shard into recv_dev (the root by a device copy).  One grouped ncclSend/ncclRecv. */
int ssdr_nccl_scatter(ssdr_comm_t c, const void* send_root_dev, const size_t* offsets, const size_t* counts,
                      void* recv_dev, int root);
/* The reverse: every rank's send_dev (counts[rank] bytes) lands at offsets[rank] of the root's buffer. */
int ssdr_nccl_gather(ssdr_comm_t c, const void* send_dev, void* recv_root_dev, const size_t* offsets,
                     const size_t* counts, int root);
int ssdr_nccl_allreduce_max_f64(ssdr_comm_t c, double* value);   /* host scalar in/out: max over ranks (timing) */
int ssdr_nccl_barrier(ssdr_comm_t c);
int ssdr_nccl_sync(ssdr_comm_t c);

/* Deterministic synthetic IQ written directly in HBM (bench / full-size tests): per channel three
 * tones (0.5, 0.05, 0.005 FS at hashed bins) + uniform-sum noise of sigma ~1e-3 FS, SURVEY 8d.
 * iq_dev: [batch][frames][nfft] in the given format. */
int ssdr_synth_iq_dev(void* iq_dev, int iq_format, int batch, int frames, int nfft, uint32_t seed);

/* ---------------------------------------------------------------------------------------------
 * waterfall:  IQ frames -> FFT -> |X|^2 -> Kiwi byte line -> time-binning mean -> colour row
 *
 * Replaces, per channel: the (remote) KiwiSDR W/F computation that delivers uint8 lines
 * (utils_supersdr.py:780-785), kiwi_waterfall.run's averaging (:881-886) and
 * kiwi_waterfall.spectrum_db2col (:787-813); pixel = uint8(rint(wf_color)).
 * ------------------------------------------------------------------------------------------- */

/* per-channel display parameters = the attributes spectrum_db2col reads (utils_supersdr.py:592-620) */
typedef struct {
    int32_t zoom;             /* kiwi_waterfall.zoom                       */
    int32_t auto_scale;       /* wf_auto_scaling                           */
    int32_t delta_low_db;     /* delta_low_db                              */
    int32_t delta_high_db;    /* delta_high_db                             */
    float   low_clip_db;      /* kept value used when auto_scale == 0      */
    float   dynamic_range;    /* kept value used when auto_scale == 0      */
} ssdr_wf_display_t;

/* per-channel scalar results of one row */
typedef struct {
    float low_clip_db, high_clip_db, dynamic_range, wf_min_db, wf_max_db;
} ssdr_wf_scalars_t;

/* nfft: power of two 256..65536 (WF_BINS; 32768 and 65536 take a three-kernel path through an HBM scratch
 * buffer); batch: channels; n_avg: averaging_n 1..100;
 * window: 1 = Hann, 0 = rectangular; cal_db: dBFS->dBm offset (SSDR_WF_CAL_DB).
 * p_lo/p_gamma: lower index and float32 weight of numpy's 40th-percentile interpolation for
 * nfft points, computed by the caller with numpy's own expression. */
int ssdr_wf_create(ssdr_wf_t* h, int nfft, int batch, int n_avg, int window, double cal_db,
                   int p_lo, float p_gamma);
int ssdr_wf_destroy(ssdr_wf_t h);
int ssdr_wf_set_display(ssdr_wf_t h, int first_channel, int count, const ssdr_wf_display_t* params);
/* Tell the handle that the device inputs of ssdr_wf_process_dev live in a PEER GPU's memory (ssdr_ipc_open): the kernel
 * then reads them in place over NVLink and skips its L2 bulk prefetch, which is pathologically slow on peer addresses. */
int ssdr_wf_set_remote_input(ssdr_wf_t h, int remote);
/* Spec tables the handle uses (for parity tests): twiddles float32[2*nfft], thresholds float32[256],
 * radix plan (returns number of passes). */
int ssdr_wf_get_tables(ssdr_wf_t h, float* twiddles, float* thresholds, int* radices);
/* First half of the periodic Hann window the handle uses, float32[nfft/2]. */
int ssdr_wf_get_window(ssdr_wf_t h, float* window_half);

/* One row per channel from HOST IQ [batch][n_avg][nfft]: H2D copy (pipelined in channel chunks
 * on the handle's streams), kernel, D2H of the requested outputs (NULL = not wanted):
 * pixels uint8[batch][nfft], colour float32[batch][nfft] (wf_color), spectrum float32[batch][nfft]
 * (averaged line in Kiwi byte units = kiwi_waterfall.spectrum), scalars[batch]. Synchronous. */
int ssdr_wf_process(ssdr_wf_t h, const void* iq_host, int iq_format, uint8_t* pixels, float* colour,
                    float* spectrum, ssdr_wf_scalars_t* scalars);
/* Same with DEVICE pointers; asynchronous on the handle's stream (ssdr_wf_sync to wait). */
int ssdr_wf_process_dev(ssdr_wf_t h, const void* iq_dev, int iq_format, uint8_t* pixels_dev,
                        float* colour_dev, float* spectrum_dev, ssdr_wf_scalars_t* scalars_dev);
/* Tier-P entry: the reference's own input.  lines uint8[batch][n_avg][nfft] are finished Kiwi
 * W/F lines (utils_supersdr.py:783-784); computes mean + spectrum_db2col + pixels. Host buffers. */
int ssdr_wf_colorrow_u8(ssdr_wf_t h, const uint8_t* lines_host, uint8_t* pixels, float* colour,
                        float* spectrum, ssdr_wf_scalars_t* scalars);
int ssdr_wf_colorrow_u8_dev(ssdr_wf_t h, const uint8_t* lines_dev, uint8_t* pixels_dev, float* colour_dev,
                            float* spectrum_dev, ssdr_wf_scalars_t* scalars_dev);
int ssdr_wf_sync(ssdr_wf_t h);
/* CUDA-event timing on the handle's compute stream: time `iters` back-to-back launches of
 * ssdr_wf_process_dev with the given device buffers; returns total milliseconds. */
int ssdr_wf_time_dev(ssdr_wf_t h, const void* iq_dev, int iq_format, uint8_t* pixels_dev, int iters,
                     float* total_ms);

/* ---------------------------------------------------------------------------------------------
 * demodulator:  IQ @12 kHz -> NCO mix -> FIR band-pass -> AM/SSB/CW/NBFM detect -> AGC -> PCM
 *
 * Replaces the (remote) KiwiSDR SND computation that the reference parametrises with
 * "SET mod= low_cut= high_cut= freq=" (utils_supersdr.py:1026-1029) and
 * "SET agc= hang= thresh= slope= decay= manGain=" (:1022-1024) and receives as int16 PCM
 * (kiwi_sound.process_audio_stream :1044-1076).
 * ------------------------------------------------------------------------------------------- */
typedef struct {
    int32_t mode;             /* SSDR_MODE_*                    kiwi_sound.radio_mode */
    float   low_cut_hz;       /* kiwi_sound.lc                                          */
    float   high_cut_hz;      /* kiwi_sound.hc                                          */
    float   freq_offset_hz;   /* tuning offset inside the IQ band (0 = carrier at DC)   */
    int32_t agc_on;           /* kiwi_sound.on                                          */
    int32_t agc_hang;         /* .hang                                                  */
    float   agc_thresh_dbm;   /* .thresh                                                */
    float   agc_slope_db;     /* .slope                                                 */
    float   agc_decay_ms;     /* .decay                                                 */
    float   agc_man_gain_db;  /* .gain (manGain)                                        */
    float   taps[SSDR_FIR_TAPS]; /* real low-pass prototype, designed by the host (float32)   */
} ssdr_demod_params_t;

int ssdr_demod_create(ssdr_demod_t* h, int batch, int max_samples_per_call);
int ssdr_demod_destroy(ssdr_demod_t h);
int ssdr_demod_set(ssdr_demod_t h, int first_channel, int count, const ssdr_demod_params_t* params);
int ssdr_demod_reset(ssdr_demod_t h);   /* zero all per-channel streaming state */
/* FIR engine of the fused kernel.  FFMA: direct form on the fp32 pipe (one warp per channel).  TCGEN05: the FIR as a
 * Toeplitz GEMM on the 5th-generation tensor cores (TF32 + bfloat16 split, accumulators in TMEM); four channels that share a
 * filter (bitwise-equal taps) make one tile, so it pays when channels share pass-band widths.  AUTO (default): TCGEN05
 * when at least half of the tile rows would carry a channel and -- for banks of more rounds than SMs -- its measured cost
 * model (rounds share one filter: few channels per filter leave tiles empty) beats the FFMA engine's, else FFMA.  Same per-channel state, same outputs to the
 * demodulator's tolerance (1e-5 relative RMS); engines may be switched between calls.  No reference counterpart (the
 * DSP is remote, utils_supersdr.py:1022-1029). */
#define SSDR_DEMOD_ENGINE_FFMA    0
#define SSDR_DEMOD_ENGINE_TCGEN05 1
#define SSDR_DEMOD_ENGINE_AUTO    2
int ssdr_demod_set_engine(ssdr_demod_t h, int engine);
/* Work plan of the TCGEN05 engine, host only (needs no device; exposed for tests and capacity planning).  taps
 * [batch][SSDR_FIR_TAPS], work [batch] = mode * 4 + agc_on * 2 + agc_hang.  Channels are grouped by bitwise-equal taps, then
 * by `work`, into quads (one tensor-core tile: four channels, unused slots -1); `*tiles_per_round` consecutive quads of one
 * filter form a round (padded with empty quads), the dearest detectors first, and the last partial wave over `n_sm` SMs is
 * spread over narrower rounds.  Writes quad_ch [n_quads][4] and quad_fid [n_quads] (filter id) when both are non-NULL
 * (capacity `cap_quads` entries), always *n_quads; *fill = batch / (4 x non-empty quads). */
int ssdr_demod_plan(const float* taps, const int32_t* work, int batch, int n_sm, int32_t* quad_ch, int32_t* quad_fid,
                    int cap_quads, int* n_quads, int* tiles_per_round, float* fill);
/* n_samples per channel, multiple of SSDR_FRAME.  iq [batch][n_samples]; outputs (NULL = skip):
 * pcm_f32 [batch][n_samples], pcm_i16 [batch][n_samples] (rint + saturate),
 * rssi_dbm [batch][n_samples/512] (what the SND header's s-meter carries, utils:1068-1069). */
int ssdr_demod_process(ssdr_demod_t h, const void* iq_host, int iq_format, int n_samples,
                       float* pcm_f32, int16_t* pcm_i16, float* rssi_dbm);
int ssdr_demod_process_dev(ssdr_demod_t h, const void* iq_dev, int iq_format, int n_samples,
                           float* pcm_f32_dev, int16_t* pcm_i16_dev, float* rssi_dev);
int ssdr_demod_sync(ssdr_demod_t h);
int ssdr_demod_time_dev(ssdr_demod_t h, const void* iq_dev, int iq_format, int n_samples,
                        float* pcm_f32_dev, int16_t* pcm_i16_dev, int iters, float* total_ms);

/* ---------------------------------------------------------------------------------------------
 * audio interpolator:  int16 PCM @12 kHz -> x`ratio` zero-stuff + FIR low-pass -> stereo int16
 *
 * Replaces kiwi_sound.play_buffer's integer-ratio path (utils_supersdr.py:1121-1138) with the
 * filter of `filtering` (utils_supersdr.py:333-348); float64 arithmetic as in the reference.
 * ------------------------------------------------------------------------------------------- */
int ssdr_interp_create(ssdr_interp_t* h, int batch, int ratio, const double* taps, int n_taps,
                       int max_samples_per_call);
int ssdr_interp_destroy(ssdr_interp_t h);
int ssdr_interp_reset(ssdr_interp_t h);
/* pcm int16[batch][n]; volume[batch] (kiwi_sound.volume, percent), balance[batch]
 * (audio_balance); stereo_out int16[batch][ratio*n][2]; mono_f64 (optional) float64[batch][ratio*n]
 * = the pre-cast buffer. */
int ssdr_interp_process(ssdr_interp_t h, const int16_t* pcm_host, int n, const float* volume,
                        const float* balance, int16_t* stereo_out, double* mono_f64);
int ssdr_interp_process_dev(ssdr_interp_t h, const int16_t* pcm_dev, int n, const float* volume_dev,
                            const float* balance_dev, int16_t* stereo_dev, double* mono_dev);
int ssdr_interp_sync(ssdr_interp_t h);

/* kiwi_sound.play_buffer, non-integer sample ratio ("high bandwidth kiwis", utils_supersdr.py:1125-1126):
 * scipy.signal.resample_poly(pcm * volume/100, up, down, padtype="line"), stateless per block, then balance and the
 * int16 cast of utils_supersdr.py:1136-1138.  h[n_h] is the polyphase filter exactly as resample_poly builds it
 * (firwin(2*10*max(up,down)+1, 1/max(up,down), window=("kaiser", 5.0)) * up, zero-padded in front); output sample j
 * is upfirdn sample first + j, j < n_keep (the reference keeps n_out - 1 samples).  Host buffers; pcm int16[batch][n],
 * stereo_out int16[batch][n_keep][2], mono_f64 (optional) float64[batch][n_keep]. */
int ssdr_resample_line(const int16_t* pcm_host, int batch, int n, const float* volume, const float* balance,
                       const double* h, int n_h, int up, int down, int first, int n_keep,
                       int16_t* stereo_out, double* mono_f64);

/* filtering.lowpass (utils_supersdr.py:346-348): np.convolve(signal, h, "valid") in float64.
 * x[n] -> out[n - n_taps + 1]; separate multiply / add in ascending tap order. */
int ssdr_fir_valid_f64(const double* x_host, size_t n, const double* taps, int n_taps, double* out_host);

/* ---------------------------------------------------------------------------------------------
 * display epilogues:  colour rows -> scrolling waterfall image -> RGB, and the spectrum trace
 *
 * Replaces kiwi_waterfall.run's image bookkeeping (utils_supersdr.py:893-897: 3-deep delay deque, wf_data scrolled
 * one line per row -- here a ring, no O(H W) copy), the palette look-up pygame does for
 * make_surface(wf_data.T).set_palette(palRGB) (supersdr.py:929-930, create_cm utils_supersdr.py:1391-1412;
 * pixel index = uint8(rint(value)), parity unpinned: the cast happens inside pygame) and
 * display_stuff.plot_spectrum's trace (utils_supersdr.py:1678-1679).  All channels of a handle advance together.
 * ------------------------------------------------------------------------------------------- */
typedef struct ssdr_wf_image* ssdr_wf_image_t;
/* palette_rgb: uint8[256][3] (entry i colours pixel value i). */
int ssdr_wf_image_create(ssdr_wf_image_t* h, int batch, int height, int width, const uint8_t* palette_rgb);
int ssdr_wf_image_destroy(ssdr_wf_image_t h);
/* One wf_color row per channel, float32[batch][width] (host, or device with _dev): delay deque + scroll. */
int ssdr_wf_image_push(ssdr_wf_image_t h, const float* colour_host);
int ssdr_wf_image_push_dev(ssdr_wf_image_t h, const float* colour_dev);
/* kiwi_waterfall.set_white_flag (utils_supersdr.py:875-877): display line 0 <- 255. */
int ssdr_wf_image_white(ssdr_wf_image_t h);
/* rgb uint8[batch][height][width][3]; wf_data float64[batch][height][width] (= kiwi_waterfall.wf_data), NULL = skip. */
int ssdr_wf_image_get(ssdr_wf_image_t h, uint8_t* rgb, double* wf_data);
/* v float64[batch][width] = nanmean of the newest t_avg lines; y int32[batch][width] = spectrum_height - 1 -
 * int(v / 255 * spectrum_height) (-1 where v is NaN); NULL = skip. */
int ssdr_wf_image_trace(ssdr_wf_image_t h, int t_avg, int spectrum_height, double* v, int32_t* y);

/* ---------------------------------------------------------------------------------------------
 * IMA-ADPCM decode (kiwi/client.py:33-87 ImaAdpcmDecoder; Kiwis with audio compression enabled -- SuperSDR itself
 * sends compression=0, utils_supersdr.py:978): data uint8[batch][n_bytes], two 4-bit codes per byte, low nibble
 * first -> pcm int16[batch][2*n_bytes].  state int32[batch][2] = (step index, previous sample) per stream, read
 * and updated, so consecutive calls stream (all zero = a fresh decoder).  Host buffers.
 * ------------------------------------------------------------------------------------------- */
int ssdr_adpcm_decode(const uint8_t* data_host, int batch, int n_bytes, int32_t* state, int16_t* pcm_out);

/* ---------------------------------------------------------------------------------------------
 * IQ wire-format unpack (kiwi/client.py:443-454): big-endian int16 I,Q -> complex64, unscaled
 * ------------------------------------------------------------------------------------------- */
int ssdr_unpack_iq_s16be(const void* s16be_host, float* cf32_host, size_t n_complex);
int ssdr_unpack_iq_s16be_dev(const void* s16be_dev, float* cf32_dev, size_t n_complex);

#ifdef __cplusplus
}
#endif
#endif /* SSDR_B200_H */
