/*
 * dy4_oracle.h — C API shared by the two CPU checkers under oracle/.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product:
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * `--impl reference` leg may load these libraries, and only as the checker
 * or the timed CPU baseline — never as a fallback for the CUDA path.
 *
 * Two libraries export the SAME functions under two prefixes:
 *   libdy4oracle.so  (prefix dy4o_)  plain-C restatement of the reference
 *                                    algorithm, oracle/dy4_oracle.c
 *   _ref/libdy4ref.so (prefix dy4r_) the reference's own filter.cpp /
 *                                    project.cpp, compiled in place from
 *                                    /root/reference by oracle/Makefile and
 *                                    driven by oracle/ref_replay.cpp
 * The prefix is selected with DY4_ORACLE_PREFIX before including this file.
 *
 * Parity pinning: the reference ships no golden vectors for this path
 * (SURVEY.md §4, §8c), so the restatement is pinned against outputs of the
 * reference itself run in the build container: tests/golden/*.npz are minted
 * by tests/golden/make_golden.py from _ref/libdy4ref.so and both libraries
 * are checked against them bit-for-bit in tests/test_oracle.py.
 */
#ifndef DY4_ORACLE_H
#define DY4_ORACLE_H

#include <stdint.h>

#ifndef DY4_ORACLE_PREFIX
#define DY4_ORACLE_PREFIX dy4o_
#endif
#define DY4_CAT2(a, b) a##b
#define DY4_CAT(a, b) DY4_CAT2(a, b)
#define DY4_FN(name) DY4_CAT(DY4_ORACLE_PREFIX, name)

#ifdef __cplusplus
extern "C" {
#endif

/* One row of the mode table, reference src/project.cpp:178-238. */
typedef struct {
    float rf_Fs;          /* RF sample rate, IQ pairs per second            */
    int   rf_decim;       /* front-end decimation                            */
    float if_Fs;          /* "audio_Fs" in the reference: the IF rate        */
    int   audio_decim;    /* D of the audio resampler                        */
    int   audio_upsample; /* U of the audio resampler                        */
    int   audio_taps;     /* 101 * U                                         */
    int   block_size;     /* bytes (= interleaved uint8 I,Q) per block       */
    int   if_per_block;   /* block_size / 2 / rf_decim                       */
    int   audio_per_block;/* if_per_block * U / D                            */
} dy4_mode_t;

int  DY4_FN(mode_params)(int mode, dy4_mode_t* out);

void DY4_FN(lpf_taps)(float Fs, float Fc, unsigned short num_taps, int up, float* h);
void DY4_FN(bpf_taps)(float Fs, float Fb, float Fe, unsigned short num_taps, int up, float* h);

void DY4_FN(iq_to_float)(const uint8_t* raw, long n, float* out);
void DY4_FN(block_fir)(const float* x, int nx, const float* h, int nh, float* state, int nstate, float* y);
void DY4_FN(decim_fir)(int factor, const float* x, int nx, const float* h, int nh, float* state, int nstate, float* y);
int  DY4_FN(resample_fir)(int up, int down, const float* x, int nx, const float* h, int nh, float* state, int nstate, float* y);
void DY4_FN(fm_demod)(const float* I, const float* Q, int n, float* prev_I, float* prev_Q, float* out);
/* pll_state = {feedbackI, feedbackQ, integrator, phaseEst, trigOffset, nco_state} */
void DY4_FN(pll)(const float* in, int n, float freq, float Fs, float ncoScale, float phaseAdjust,
                 float normBandwidth, float* nco, float* pll_state);
void DY4_FN(delay_block)(const float* in, int n, float* state, int nstate, float* out);
void DY4_FN(pcm16)(const float* x, long n, int16_t* out);

/*
 * One whole receiver over one stream: what `project <mode> <mono|stereo>`
 * does to `nbytes` of stdin (whole blocks only; a trailing partial block is
 * dropped, project.cpp:293-296).  Returns the number of blocks processed.
 * Any output pointer may be NULL.  Sizes per block: if_out/pilot_out/nco_out
 * if_per_block floats; audio_out audio_per_block floats (mono) or
 * 2*audio_per_block floats interleaved L,R (stereo); pcm_out likewise int16.
 */
long DY4_FN(pipeline)(int mode, int stereo, const uint8_t* iq, long nbytes,
                      float* if_out, float* audio_out, int16_t* pcm_out,
                      float* pilot_out, float* nco_out);

/*
 * Fourier diagnostics, reference src/fourier.cpp.  Complex vectors are interleaved (re, im) floats.
 *   dft           :14-23   naive O(n^2) DFT of a real vector, float twiddles exp(i * float(-2 PI k m / n))
 *   idft          :98-107  naive inverse DFT of a complex vector, divided by n
 *   estimate_psd  :37-94   Hann-windowed segments of nfft samples, DFT, 10 log10 of the scaled power, segment average;
 *                          freq and psd hold nfft/2 floats
 */
void DY4_FN(dft)(const float* x, int n, float* Xf);
void DY4_FN(idft)(const float* Xf, int n, float* x);
void DY4_FN(estimate_psd)(const float* samples, long n, int nfft, int Fs, float* freq, float* psd);
/* fourier.cpp:125-211: compute_twiddles (NFFT = 512) and the three radix-2 FFTs (variant 0 recursive, 1 improved, 2 optimized);
 * complex vectors as interleaved (re, im) floats, n a power of two (n == 512 for variants 1 and 2) */
void DY4_FN(compute_twiddles)(int n_tw, float* out);
void DY4_FN(fft)(const float* x, int n, int variant, float* Xf);

#ifdef __cplusplus
}
#endif
#endif
