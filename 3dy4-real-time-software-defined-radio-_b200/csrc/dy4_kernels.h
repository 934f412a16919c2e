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
    const float* if_tail;                       // [n_streams][DY4_IF_TAIL] IF samples preceding the chunk ...
    long long if_tail_stride = 0;               // ... rows if_tail_stride floats apart (0: DY4_IF_TAIL).  With if_tail = if_in - DY4_IF_TAIL and the stride of if_in the
                                                // history is simply the samples that precede the chunk in the same (call-wide) rows.
    float* pilot; float* sband; long long out_stride;
    int n_if, n_streams, mode;                  // mode 0..3: (pilot, stereo) table; 4, 5: RDS band-pass / carrier tables
    int variant;                                // 0: exact (pilot, stereo) pair of one stream; 1: ONE fused (h,h) filter over two streams per CTA -> pilot rows;
                                                // 2: as 1 on the SQUARED input
    unsigned long long neg_zero2;
};

struct Dy4PllArgs {
    const float* in; long long in_stride;       // pilot
    double* inv; double* theta; long long wide_stride;  // scratch rows of doubles: 1/x per input sample; trigArg after each sample
    float* tstart = nullptr;                            // scratch [n_streams] (table-driven loop): sample counter at the start of the launch
    const double* pred_in = nullptr; double* pred_out = nullptr; int pred_carry = 0;   // table-driven loop: the predictor's own state [n_streams][8] before / after this launch
    double* need = nullptr;                             // [n_streams]: whole turns between prediction and true phaseEst (written by the loop of this set, read two launches later)
    int fresh = 0; float* nco0b = nullptr;              // table-driven loop, first launch of a stream: samples left to the direct loop; second NCO carry row
    int* risk = nullptr;                                // optional [n_streams]: census of near-tie narrowings in the table's evaluations (dy4_near_float_tie)
    float4* tab = nullptr; long long tab_stride = 0;    // optional [n_streams][tab_stride] 16-byte words, 2 per sample: selects the table-driven loop (dy4_plltab.h)
    float* nco0;                                // scratch [n_streams]: NCO value that opens this launch's row
    float* nco; long long nco_stride;
    float* state;                               // [n_streams][8]: fbI fbQ integ phase trigOffset nco_state pad pad
    int n, n_streams;
    float freq, Fs, ncoScale, phaseAdjust, normBandwidth;
};

struct Dy4AudioArgs {
    const float* if_in; long long if_stride; const float* if_tail;
    long long if_tail_stride = 0;                               // as in Dy4BpfArgs
    const float* nco; const float* sband; long long bb_stride;  // stereo only (else NULL)
    const float* mix_tail;                                      // [n_streams][DY4_MIX_TAIL] of nco*sband*2 before the chunk
    float* audio; long long audio_stride;                       // may be NULL; mono: [n] ; stereo: interleaved L,R [2n]
    int16_t* pcm; long long pcm_stride;                         // may be NULL
    int n_if, n_audio, n_streams, stereo, up, down, exact;
    int mode;                                                   // U==1: selects the (h,h) audio low-pass table
    const float* taps_poly;                                     // U>1: device [101][up_pad] polyphase taps, taps_poly[j*up_pad+phase] = h[phase + j*up]
    int up_pad;
    int poly_variant = 0;                                       // U>1: 0 the better of the two per ratio, 1 the table-in-shared-memory kernel, 2 the tap-stationary kernel
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
cudaError_t dy4_launch_bpf_mixed(const Dy4BpfArgs& a, cudaStream_t st);   // pilot exact, stereo band fused (modes 0..3, both outputs)
cudaError_t dy4_launch_pll(const Dy4PllArgs& a, cudaStream_t st);
enum { DY4_PLL_PREP = 1, DY4_PLL_LOOP = 2, DY4_PLL_NCO = 4 };   // k_pll_prep (reciprocals), k_pll (serial loop), k_nco (NCO row)
cudaError_t dy4_launch_pll_parts(const Dy4PllArgs& a, cudaStream_t st, int parts);
cudaError_t dy4_launch_audio(const Dy4AudioArgs& a, cudaStream_t st);
cudaError_t dy4_launch_tails(const Dy4TailArgs& a, cudaStream_t st);
// one-time (per device) upload of the four per-mode tap tables
cudaError_t dy4_upload_taps_frontend(const TapPairs* rf4);
cudaError_t dy4_upload_taps_bpf(const TapPairs* bpf6);
cudaError_t dy4_upload_taps_audio(const TapPairs* audio4);

// ---- RDS filtering front end (Python model only: model/fmMonoBlock.py:673-696) -------------------------------------
struct Dy4RdsArgs {
    const float* rds_f; long long stride;       // 54-60 kHz band-pass output of this chunk (IF rate)
    const float* rds_tail;                      // [n_streams][DY4_IF_TAIL] of it before the chunk
    const float* carrier;                       // 113.5-114.5 kHz band-pass of rds_f^2
    double* theta; long long wide_stride;       // scratch: PLL phase argument after each sample
    float* nco_i; float* nco_q;                 // scratch rows (stride)
    double* pll_state;                          // [n_streams][8]: integ, phase, trigOffset, reduced phase, ncoI_state, ncoQ_state
    float* mix_tail;                            // [n_streams][2][DY4_MIX_TAIL]: mixed I / Q before the chunk
    float* lp; long long lp_stride;             // scratch [n_streams][2][lp_stride]: resampler output I, Q of this chunk
    float* lp_tail;                             // [n_streams][2][DY4_MIX_TAIL] of it before the chunk
    float* out_i; float* out_q; long long out_stride;   // RRC-filtered I / Q (38 kS/s) of this chunk
    const float* taps_poly; int up_pad;         // [101][up_pad] polyphase low-pass taps (gain up), up = 19
    const float* taps_rrc;                      // 101 RRC taps
    int n_if, n_streams;
    long long if_abs;                           // absolute IF index of the chunk's first sample (resampler phase)
    long long m_first; int n_out;               // absolute index of the first resampler output of this chunk and their count
    double w, Kp, Ki, nco_scale, phase_adjust;
    int up, down;
    int span_max;                               // set by the launcher: pairs of shared memory reserved for a tile's samples
};
cudaError_t dy4_upload_taps_rrc(const float* rrc);                            // 101 RRC taps -> constant memory (per device)
cudaError_t dy4_launch_rds_pll(const Dy4RdsArgs& a, cudaStream_t st);        // carrier -> theta -> nco_i / nco_q
cudaError_t dy4_launch_rds_resample(const Dy4RdsArgs& a, cudaStream_t st);   // delay, mix, 19/120 resampler, RRC, tails

// RDS back half (model/fmSupportLib.py:209-247, model/fmMonoBlock.py:78-284, 699-730): one thread per stream walks whole
// model blocks of DY4_RDS_BLOCK in-phase RRC samples: symbol timing, Manchester + differential decoding, frame sync.
constexpr int DY4_RDS_BLOCK = 3040;            // 16 samples/symbol x 190 symbols = 19 200 IF samples x 19/120
constexpr int DY4_RDS_STATE_INTS = 20;
struct Dy4RdsDecodeArgs {
    const float* acc; long long acc_stride; int n_blocks;      // n_blocks whole model blocks at the start of every row
    int* state;                                                // [n_streams][DY4_RDS_STATE_INTS]
    int* counts;                                               // [n_streams][4]: symbols, bits, events, groups written so far
    int8_t* sym; long long sym_stride; int sym_cap;
    int8_t* bits; long long bits_stride; int bits_cap;
    int* events; long long ev_stride; int ev_cap;              // 4 ints per event: type (A,B,C,C',D = 0..4), bit position, false-positive flag, 16-bit word
    int* groups; long long grp_stride; int grp_cap;            // 4 ints per complete group handed to the application layer: the A, B, C, D words
    int n_streams;
};
cudaError_t dy4_launch_rds_append(const float* rrc_i, long long rrc_stride, int n_new, float* acc, long long acc_stride,
                                  int consumed, int left, int n_streams, cudaStream_t st);
cudaError_t dy4_launch_rds_decode(const Dy4RdsDecodeArgs& a, cudaStream_t st);

// Generic single-op kernels for the filter.h compatibility tier (any tap count, any factor).
cudaError_t dy4_launch_generic_fir(const float* x_ext, int n_hist, int n_out, int step, const float* h, int nh, float* y, cudaStream_t st);
cudaError_t dy4_launch_generic_resample(const float* x_ext, int n_hist, int n_out, int up, int down, const float* h, int nh, float* y, cudaStream_t st);
cudaError_t dy4_launch_generic_demod(const float* I, const float* Q, int n, float prev_I, float prev_Q, float* out, cudaStream_t st);
cudaError_t dy4_launch_u8_to_float(const uint8_t* raw, long long n, float* out, cudaStream_t st);
cudaError_t dy4_launch_pointwise(int op, const float* a, const float* b, int n, float* out, cudaStream_t st);

// SM partition (dy4_smpart.cu): green contexts that keep the PLL's serial loops on SMs of their own
int dy4_sm_partition(int device, int loop_sms, int* n_loop, int* n_rest);
cudaError_t dy4_sm_partition_stream(int device, int loop_sms, int which, int priority, cudaStream_t* s);
