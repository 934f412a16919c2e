/*
 * dy4_b200.h — C ABI of libdy4b200.so: the B200 (sm_100a) implementation of the
 * 3DY4 FM receiver's data-parallel hot path.
 *
 * The reference has no FFI: its "operator interface" for this path is the set
 * of free C++ functions in /root/reference/include/filter.h:17-34, called only
 * from src/project.cpp (frontend :72-93, backend :95-134, main :262-273,:310).
 * This header is what a binding of that interface binds to.  Two tiers:
 *
 *  (1) Compatibility tier — one entry point per filter.h prototype, plain host
 *      pointers and sizes, same argument meaning, state updated in place.  The
 *      C++ shim csrc/filter_shim.cpp re-exports them under the reference's own
 *      C++ signatures so project.cpp links against this library unchanged
 *      (INTEGRATION.md).  Each call copies host->device, runs a CUDA kernel and
 *      copies back: correct and drop-in, not fast.
 *
 *  (2) Throughput tier — a batched receiver ("pipeline") over n_streams
 *      independent streams laid out [stream][time], state resident on the
 *      device, fused sm_100a kernels.  This replaces the reference's block loop
 *      (project.cpp:289-318), its stdin reader (iofunc.cpp:113-120) and its
 *      threadSafeQ hand-off (threadSafeQ.cpp:18-55).
 *
 * All functions return 0 on success or a negative dy4 error code; the text of
 * the last error on the calling thread is available from dy4_last_error().
 * There is no CPU fallback: without a CUDA device every compute entry point
 * fails with DY4_ERR_CUDA.
 */
#ifndef DY4_B200_H
#define DY4_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DY4_OK 0
#define DY4_ERR_ARG (-1)      /* invalid argument (mode, sizes, alignment, NULL) */
#define DY4_ERR_CUDA (-2)     /* CUDA runtime error, see dy4_last_error()        */
#define DY4_ERR_NOMEM (-3)

const char* dy4_last_error(void);
int dy4_version(void);

/* ---- mode table: reference src/project.cpp:178-238 ----------------------- */
typedef struct {
    float rf_Fs;           /* IQ pairs per second in                          */
    int   rf_decim;        /* front-end decimation                            */
    float if_Fs;           /* IF rate ("audio_Fs" in the reference)           */
    int   audio_decim;     /* D of the audio resampler                        */
    int   audio_upsample;  /* U of the audio resampler                        */
    int   audio_taps;      /* 101 * U                                         */
    int   block_size;      /* bytes (interleaved uint8 I,Q) per block         */
    int   if_per_block;    /* IF samples per block                            */
    int   audio_per_block; /* audio samples per block and channel             */
} dy4_mode_params_t;

int dy4_mode_params(int mode, dy4_mode_params_t* out);

/* ---- tap design (host side, one-time) ------------------------------------ */
/* filter.h:17 impulseResponseLPF(Fs, Fc, num_taps, h, upFactor); h has num_taps floats */
int dy4_lpf_taps(float Fs, float Fc, unsigned short num_taps, int up_factor, float* h);
/* filter.h:27 impulseResponseBPF(Fs, Fb, Fe, num_taps, h, upFactor) */
int dy4_bpf_taps(float Fs, float Fb, float Fe, unsigned short num_taps, int up_factor, float* h);

/* RDS path (reference: Python model only): scipy.signal.firwin(num_taps, [lo,hi], window=..., pass_zero=False) —
 * lo <= 0 gives the low-pass firwin(num_taps, hi, window=...); window 0 = 'hann', 1 = 'hamming' (scipy's default);
 * frequencies normalised to Nyquist — and model/fmRRC.py:13 impulseResponseRootRaisedCosine(Fs, N_taps).
 * Double precision, as the model. */
int dy4_firwin(int num_taps, double lo, double hi, int window, double* h);
int dy4_rrc_taps(double Fs, int num_taps, double* h);

/* ---- compatibility tier: host pointers, one stream ----------------------- */
/* iofunc.cpp:117-119: out[k] = (raw[k]-128)/128 */
int dy4_iq_to_float(const uint8_t* raw, size_t n, float* out);
/* filter.h:18 convolveFIR: y has nx+nh-1 floats */
int dy4_convolve_fir(float* y, const float* x, size_t nx, const float* h, size_t nh);
/* filter.h:19 blockConvolveFIR: y has nx floats; state (nstate floats) is replaced by the tail of x */
int dy4_block_fir(float* y, const float* x, size_t nx, const float* h, size_t nh, float* state, size_t nstate);
/* filter.h:25 downsampleBlockConvolveFIR: y has nx/factor floats */
int dy4_decim_fir(int factor, float* y, const float* x, size_t nx, const float* h, size_t nh, float* state, size_t nstate);
/* filter.h:26 resampleBlockConvolveFIR: y has (nx/down)*up floats; *ny receives that count */
int dy4_resample_fir(int up, int down, float* y, size_t* ny, const float* x, size_t nx, const float* h, size_t nh, float* state, size_t nstate);
/* filter.h:20 fmDemodArctan */
int dy4_fm_demod(const float* I, const float* Q, size_t n, float* prev_I, float* prev_Q, float* fm_demod);
/* filter.h:29 fmPLL; the six float& of the reference in its argument order */
int dy4_pll(const float* pll_in, size_t n, float freq, float Fs, float nco_scale, float phase_adjust, float norm_bandwidth,
            float* nco_out, float* feedbackI, float* feedbackQ, float* integrator, float* phaseEst, float* trigOffset, float* nco_state);
/* filter.h:22-23 downsample / upsample */
int dy4_downsample(const float* data, size_t n, size_t factor, float* out, size_t* n_out);
int dy4_upsample(const float* data, size_t n, size_t factor, float* out, size_t* n_out);
/* filter.h:30 delayBlock */
int dy4_delay_block(const float* in, size_t n, float* state, size_t nstate, float* out);
/* filter.h:31-33: multiply carries the reference's gain of 2 (filter.cpp:264) */
int dy4_pointwise_multiply(const float* a, size_t na, const float* b, size_t nb, float* out, size_t* n_out);
int dy4_pointwise_add(const float* a, const float* b, size_t n, float* out);
int dy4_pointwise_subtract(const float* a, const float* b, size_t n, float* out);
/* filter.h:34 interleave: out has nl+nr floats */
int dy4_interleave(const float* left, size_t nl, const float* right, size_t nr, float* out);

/* ---- Fourier diagnostics: reference include/fourier.h, src/fourier.cpp ---- */
/* Not on the receiver's path (project.cpp never calls them); SURVEY.md §8f rank 3.  Complex vectors are interleaved
 * (re, im) floats.  The reference's DFT uses twiddles exp(i * float(-2 PI k m / n)) — the angle is narrowed to float
 * BEFORE cosf/sinf — and sums over k in float; these entry points reproduce exactly that (twiddle table built on the
 * host with the same libm, unfused float accumulation on the device), not an FFT.  n, nfft <= 2048.
 * fourier.h:20 DFT(x, Xf): n real samples -> n complex bins */
int dy4_dft(const float* x, size_t n, float* Xf);
/* fourier.h:38 IDFT(Xf, x): n complex bins -> n complex samples, divided by n */
int dy4_idft(const float* Xf, size_t n, float* x);
/* fourier.h:31 estimatePSD(samples, nFFT, Fs, freq, psd_est): floor(n/nfft) Hann-windowed segments, DFT of each,
 * 10 log10((1/(Fs nfft/2)) 2 |X|^2) per bin, averaged over the segments; freq and psd receive nfft/2 floats
 * (freq may be NULL).  Host pointers. */
int dy4_estimate_psd(const float* samples, size_t n, int nfft, int Fs, float* freq, float* psd);
/* The same PSD for n_streams rows of n samples on the DEVICE (rows row_stride floats apart), e.g. the IF or audio
 * rows of a dy4_pipeline_process call: d_psd[n_streams][nfft/2], rows psd_stride floats apart, asynchronously on
 * `stream` (a cudaStream_t). */
int dy4_psd_batch(const float* d_samples, size_t row_stride, int n_streams, size_t n, int nfft, int Fs,
                  float* d_psd, size_t psd_stride, void* stream);
/* fourier.h:35-41 — the reference's three radix-2 FFTs and its twiddle table (src/fourier.cpp:125-211).  complex64 vectors as
 * interleaved (re, im) floats, n a power of two <= 2048.  All three run the same butterflies; they differ in where a butterfly's
 * twiddle comes from: DY4_FFT_RECURSIVE computes exp(i * float(-2 PI float(k) / size)) at every level (:152); DY4_FFT_IMPROVED
 * (:181) and DY4_FFT_OPTIMIZED (:203) read twiddles[k * n/size] from the caller's table, which must hold NFFT/2 entries with
 * NFFT == n (compute_twiddles, :125).  Results are bit-identical to the reference's.
 * dy4_compute_twiddles: twiddles[k] = exp(i * float(-2 PI float(k) / nfft)), k < n_twiddles (the reference fixes nfft = NFFT = 512).
 * dy4_fft: host pointers, one vector.  dy4_fft_batch: n_rows complex rows on the DEVICE (strides in complex elements). */
#define DY4_FFT_RECURSIVE 0
#define DY4_FFT_IMPROVED 1
#define DY4_FFT_OPTIMIZED 2
int dy4_compute_twiddles(size_t n_twiddles, int nfft, float* twiddles);
int dy4_fft(const float* x, size_t n, int variant, const float* twiddles, size_t n_twiddles, float* Xf);
int dy4_fft_batch(const float* d_x, size_t x_stride, int n_rows, size_t n, int variant, const float* twiddles, size_t n_twiddles,
                  float* d_X, size_t X_stride, void* stream);

/* ---- throughput tier: batched receiver ----------------------------------- */
typedef struct dy4_pipeline dy4_pipeline_t;

#define DY4_FLAG_EXACT_AUDIO 1u   /* also run the stages the PLL never sees (stereo-band BPF, resamplers; in a MONO
                                     receiver the RF front end too) with unfused multiply-add: every output is then
                                     bit-identical to the reference instead of within ~1e-7 (default: fused where
                                     nothing chaotic is downstream; a stereo receiver's IF is always bit-identical) */

#define DY4_FLAG_DEBUG_ROWS 2u    /* keep each process call in ONE sub-chunk so that dy4_pipeline_debug_buffers() returns
                                     whole pilot / NCO rows (diagnostics; disables the PLL/FIR overlap) */

#define DY4_FLAG_RDS 4u           /* mode 0 stereo only: also run the RDS filtering front end of the Python model
                                     (model/fmMonoBlock.py:673-691): 54-60 kHz band-pass, squaring, 113.5-114.5 kHz
                                     band-pass, 114 kHz PLL (ncoScale 0.5, bandwidth 0.001), 50-sample delay, I/Q mix,
                                     19/120 resampler (1919 taps) and 101-tap RRC -> 38 kS/s.  Read the result of the
                                     last process call with dy4_pipeline_rds_read. */

#define DY4_FLAG_PIPELINED 8u     /* stereo, dy4_pipeline_process only: consecutive calls OVERLAP on the device.  A call returns once
                                     it is queued and is NOT joined to `stream` when it ends: its FIR kernels run while the PLL's
                                     serial loops of the call before are still going (on SMs of their own where CUDA green contexts
                                     are available).  The caller must call dy4_pipeline_flush (or synchronise the device) before it
                                     reads the outputs or overwrites the input of ANY earlier call; results are the same bits.
                                     dy4_pipeline_process_host overlaps the same way (whole calls alternate between two staging sets; the
                                     host arrays are valid after dy4_pipeline_sync).  Every other entry point (state, reset, RDS read /
                                     drain) drains the queue itself. */

/* Create a receiver for `n_streams` independent streams in `mode` (0..3), mono (stereo=0)
 * or stereo (stereo=1), on CUDA device `device`.  All carried state starts as in
 * project.cpp:240-255 (zero history, PLL at feedbackI=1, nco_state=1). */
int dy4_pipeline_create(int mode, int stereo, int n_streams, int device, unsigned flags, dy4_pipeline_t** out);
int dy4_pipeline_destroy(dy4_pipeline_t* p);
/* back to the initial state (start of a new set of streams) */
int dy4_pipeline_reset(dy4_pipeline_t* p);

/*
 * Process `n_blocks` whole blocks of every stream.  DEVICE pointers.
 *   d_iq        uint8 [n_streams][row_stride_bytes], interleaved I,Q; the first
 *               n_blocks*block_size bytes of each row are consumed.  Rows and the
 *               base must be 16-byte aligned.
 *   d_pcm       int16 [n_streams][n_blocks*audio_per_block*channels]     or NULL
 *   d_audio     float, same shape (stereo: interleaved L,R)              or NULL
 *   d_if        float [n_streams][n_blocks*if_per_block]                 or NULL
 *   stream      a cudaStream_t (as void*), NULL = the default stream
 * Successive calls continue the same streams (state is carried on the device),
 * exactly as successive blocks do in the reference.  Asynchronous on `stream`;
 * successive calls on one pipeline must be issued on the same stream (or ordered
 * by the caller): the carried state is read and written in stream order.
 */
int dy4_pipeline_process(dy4_pipeline_t* p, const uint8_t* d_iq, size_t row_stride_bytes, int n_blocks,
                         int16_t* d_pcm, float* d_audio, float* d_if, void* stream);
/* DY4_FLAG_PIPELINED: make `stream` wait for every call queued so far (outputs complete, inputs no longer read).
 * Without the flag calls are joined to their stream when they end and this is a no-op. */
int dy4_pipeline_flush(dy4_pipeline_t* p, void* stream);
/* Block the calling thread until everything this pipeline has queued is done, on every stream of its own (device calls, uploads,
 * downloads).  With DY4_FLAG_PIPELINED dy4_pipeline_process_host returns when the call is QUEUED — its upload runs beside the
 * kernels of the call before — and the host arrays are valid after this. */
int dy4_pipeline_sync(dy4_pipeline_t* p);

/*
 * Same, HOST pointers (pinned memory recommended): the input is uploaded in
 * sub-chunks on a copy stream, each sub-chunk's kernels wait only for their own
 * bytes, and each sub-chunk's PCM/audio goes back device->host on a third stream
 * as soon as it exists — the replacement for the reference's stdin reader + queue.
 * chunk_blocks > 0 caps the blocks resident in device staging at once (default:
 * DY4_STAGE_BYTES of input, 4 GiB).  Synchronous: waits for whatever dy4_pipeline_process
 * calls are still queued on the device, and returns when all outputs are in host memory
 * (DY4_FLAG_PIPELINED: returns when queued, see dy4_pipeline_sync).
 */
int dy4_pipeline_process_host(dy4_pipeline_t* p, const uint8_t* h_iq, size_t row_stride_bytes, int n_blocks,
                              int16_t* h_pcm, float* h_audio, int chunk_blocks);

/* Page-locked host memory for the buffers handed to dy4_pipeline_process_host (pageable memory works too, slower). */
void* dy4_pinned_alloc(size_t bytes);
void dy4_pinned_free(void* p);

/* RRC-filtered RDS baseband (in-phase and quadrature, 38 kS/s) produced by the LAST process call: *n_samples of
 * them per stream (about 19/120 of the call's IF samples; the exact count follows the absolute sample index), copied
 * to DEVICE rows d_rrc_i / d_rrc_q (either may be NULL) of `row_stride` floats, asynchronously on `stream`. */
int dy4_pipeline_rds_read(dy4_pipeline_t* p, float* d_rrc_i, float* d_rrc_q, size_t row_stride, int* n_samples, void* stream);

/* RDS back half (reference: the Python model only — model/fmSupportLib.py:209-247 manchesterEncoded, model/fmMonoBlock.py
 * :78-122 find_pattern / decode, :157-284 get_window / frame_sync_receiver, glued as in :699-730): every process call
 * appends its in-phase RRC samples to a per-stream accumulator; each whole MODEL block of 3 040 samples (190 symbols) is
 * decoded on the device with the model's block-to-block state.  Blocks 0-4 only recover symbol timing, blocks 5-9 pick
 * the Manchester pairing, from block 10 on bits are decoded and the 26-bit syndrome frame synchroniser runs.
 * dy4_pipeline_rds_bounds: upper bounds (per stream) of what has accumulated since the last drain.
 * dy4_pipeline_rds_drain: waits for the RDS work queued so far, copies per stream the Manchester symbols (0/1, one
 * int8 each), the decoded bits (0/1, int8), the frame-sync events (4 x int32: block type A,B,C,C',D = 0..4, bit
 * position, false-positive flag, the 16-bit information word) and the complete groups the model's main loop hands to
 * its application layer (fmMonoBlock.py:716-730; 4 x int32: the A, B, C, D words; at most max_events of them) to HOST
 * rows of the given strides (in elements / events / groups; any pointer may be NULL), h_counts[n_streams][4] =
 * symbols, bits, events, groups; then empties the device rows. */
int dy4_pipeline_rds_bounds(dy4_pipeline_t* p, int* max_symbols, int* max_bits, int* max_events);
int dy4_pipeline_rds_drain(dy4_pipeline_t* p, int8_t* h_symbols, size_t sym_stride, int8_t* h_bits, size_t bits_stride,
                           int32_t* h_events, size_t ev_stride, int32_t* h_groups, size_t grp_stride, int32_t* h_counts);

/* Diagnostics (valid after a stereo process call): device pointers to the
 * last sub-chunk's pilot and NCO rows, their stride in floats and length. */
int dy4_pipeline_debug_buffers(dy4_pipeline_t* p, const float** d_pilot, const float** d_nco, size_t* stride, int* n_if);

/* Per-kernel timing with CUDA events on the launch stream.  enable!=0 turns it on.
 * get: ms[] and launches[] have DY4_NUM_KERNELS entries, accumulated since the last reset. */
#define DY4_NUM_KERNELS 9
#define DY4_K_FRONTEND 0
#define DY4_K_BPF 1
#define DY4_K_PLL 2           /* the serial loop only */
#define DY4_K_AUDIO 3
#define DY4_K_TAILS 4
#define DY4_K_RDS_BPF 5       /* RDS: the two band-pass launches */
#define DY4_K_RDS_PLL 6       /* RDS: carrier PLL + NCO rows */
#define DY4_K_RDS_BASEBAND 7  /* RDS: mix + 19/120 resampler + RRC + carry + symbol/bit/frame decoding */
#define DY4_K_PLL_AUX 8       /* the PLL's data-parallel passes (reciprocals before, NCO row after), on the main stream */
int dy4_pipeline_profile(dy4_pipeline_t* p, int enable);
int dy4_pipeline_profile_get(dy4_pipeline_t* p, double* ms, long long* launches, int reset);
/* Census of the PLL's transcendental evaluations whose double result lay within 2 double-ulps of a float rounding boundary when
 * it was narrowed to float (filter.cpp:200,216-217 narrow atan2 / cos / sin to float): the only evaluations on which a libm other
 * than the reference's glibc could produce a different float.  One count per stream (table-driven PLL loop; streams on the direct
 * loop and mono receivers report 0), accumulated since create / the last call with reset != 0.  h_counts: int32[n_streams]. */
int dy4_pipeline_pll_risk(dy4_pipeline_t* p, int32_t* h_counts, int reset);
/* SM partition of a stereo pipeline (opt-in, environment DY4_LOOP_SMS=<multiple of 8>; CUDA green contexts): SMs that run only
 * the PLL's serial loops / SMs for everything else.  Both 0 when the pipeline runs unpartitioned (the default, or green
 * contexts unavailable) or has not processed a stereo call yet. */
int dy4_pipeline_sm_partition(dy4_pipeline_t* p, int* loop_sms, int* rest_sms);
/* total kernels launched by this library in this process */
long long dy4_launch_count(void);

/* Checkpoint of the carried per-stream state (what project.cpp:25-53 holds in its structs):
 * dy4_pipeline_state_size gives the byte count, get/set copy it device<->host. */
size_t dy4_pipeline_state_size(const dy4_pipeline_t* p);
int dy4_pipeline_get_state(dy4_pipeline_t* p, void* host_buf);
int dy4_pipeline_set_state(dy4_pipeline_t* p, const void* host_buf);

#ifdef __cplusplus
}
#endif
#endif /* DY4_B200_H */
