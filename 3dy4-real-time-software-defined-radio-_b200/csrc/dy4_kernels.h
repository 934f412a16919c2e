// dy4_kernels.h — internal launch interface between the pipeline (dy4_pipeline.cu)
// and the kernel translation units.  Not part of the public C ABI (include/dy4_b200.h).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

struct TapPairs;

// All row pointers are [n_streams][stride] with time contiguous; strides in elements.
struct Dy4FrontendArgs {
    const uint8_t* iq; long long row_stride;   // interleaved uint8 I,Q of this chunk; bytes between streams
    const uint8_t* iq_tail;                     // [n_streams][DY4_IQ_TAIL] bytes preceding the chunk
    float* if_out; long long if_stride;         // IF samples out
    int n_if, n_streams, rf_decim, exact;
    const float* taps_g;                        // device copy of the 101 RF taps
    int mode;                                   // selects the per-mode tap table in constant memory
    unsigned long long neg_zero2;               // bit pattern of (-0.f,-0.f)
};

struct Dy4BpfArgs {
    const float* if_in; long long if_stride;    // IF samples of this chunk
    const float* if_tail;                       // [n_streams][DY4_IF_TAIL] IF samples preceding the chunk
    float* pilot; float* sband; long long out_stride;
    int n_if, n_streams, mode;
    unsigned long long neg_zero2;
};

struct Dy4PllArgs {
    const float* in; long long in_stride;       // pilot
    double* inv; double* theta; long long wide_stride;  // scratch rows of doubles: 1/x per input sample; trigArg after each sample
    float* nco0;                                // scratch [n_streams]: NCO value that opens this launch's row
    float* nco; long long nco_stride;
    float* state;                               // [n_streams][8]: fbI fbQ integ phase trigOffset nco_state pad pad
    int n, n_streams;
    float freq, Fs, ncoScale, phaseAdjust, normBandwidth;
};

struct Dy4AudioArgs {
    const float* if_in; long long if_stride; const float* if_tail;
    const float* nco; const float* sband; long long bb_stride;  // stereo only (else NULL)
    const float* mix_tail;                                      // [n_streams][DY4_MIX_TAIL] of nco*sband*2 before the chunk
    float* audio; long long audio_stride;                       // may be NULL; mono: [n] ; stereo: interleaved L,R [2n]
    int16_t* pcm; long long pcm_stride;                         // may be NULL
    int n_if, n_audio, n_streams, stereo, up, down, exact;
    int mode;                                                   // U==1: selects the (h,h) audio low-pass table
    const float* taps_poly;                                     // U>1: device [101][up_pad] polyphase taps, taps_poly[j*up_pad+phase] = h[phase + j*up]
    int up_pad;
    unsigned long long neg_zero2;
};

struct Dy4TailArgs {
    const uint8_t* iq; long long row_stride; long long row_bytes; uint8_t* iq_tail;
    const float* if_in; long long if_stride; int n_if; float* if_tail;
    const float* nco; const float* sband; long long bb_stride; float* mix_tail;  // NULL in mono
    int n_streams;
};

cudaError_t dy4_launch_frontend(const Dy4FrontendArgs& a, cudaStream_t st);
cudaError_t dy4_launch_bpf(const Dy4BpfArgs& a, cudaStream_t st);
cudaError_t dy4_launch_pll(const Dy4PllArgs& a, cudaStream_t st);
cudaError_t dy4_launch_audio(const Dy4AudioArgs& a, cudaStream_t st);
cudaError_t dy4_launch_tails(const Dy4TailArgs& a, cudaStream_t st);
// one-time (per device) upload of the four per-mode tap tables
cudaError_t dy4_upload_taps_frontend(const TapPairs* rf4);
cudaError_t dy4_upload_taps_bpf(const TapPairs* bpf4);
cudaError_t dy4_upload_taps_audio(const TapPairs* audio4);

// Generic single-op kernels for the filter.h compatibility tier (any tap count, any factor).
cudaError_t dy4_launch_generic_fir(const float* x_ext, int n_hist, int n_out, int step, const float* h, int nh, float* y, cudaStream_t st);
cudaError_t dy4_launch_generic_resample(const float* x_ext, int n_hist, int n_out, int up, int down, const float* h, int nh, float* y, cudaStream_t st);
cudaError_t dy4_launch_generic_demod(const float* I, const float* Q, int n, float prev_I, float prev_Q, float* out, cudaStream_t st);
cudaError_t dy4_launch_u8_to_float(const uint8_t* raw, long long n, float* out, cudaStream_t st);
cudaError_t dy4_launch_pointwise(int op, const float* a, const float* b, int n, float* out, cudaStream_t st);
