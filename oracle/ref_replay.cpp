/*
 * ref_replay.cpp — drives the REFERENCE'S OWN CODE and exposes it through the
 * C API of dy4_oracle.h under the prefix dy4r_.
 *
 * TEST INFRASTRUCTURE ONLY.  Built by oracle/Makefile into oracle/_ref/
 * (git-ignored) from the sources where they lie under /root/reference; no
 * reference source is copied into this repository.  project.cpp is pulled in
 * textually with its main() renamed so that frontend() (project.cpp:72-93) and
 * backend() (project.cpp:95-134) run verbatim; filter.cpp and iofunc.cpp are
 * compiled as their own translation units with the reference's flags
 * (-O3 -std=c++17, plus -include cstdint which the reference needs on g++ 13).
 *
 * What this file adds is only glue: the mode table is read off
 * project.cpp:178-238 (it is local to the reference's main(), so it cannot be
 * called; tests/golden/make_golden.py cross-checks it by running the real
 * `project` binary on the same bytes and comparing the PCM bit for bit).
 */
#define DY4_ORACLE_PREFIX dy4r_
#include "dy4_oracle.h"

#define main dy4_reference_main
#include "project.cpp"
#undef main

#include <cstring>
#include <streambuf>

namespace {
struct membuf : std::streambuf {
    membuf(const uint8_t* p, size_t n) {
        char* c = const_cast<char*>(reinterpret_cast<const char*>(p));
        setg(c, c, c + n);
    }
};
}

extern "C" {

int dy4r_mode_params(int mode, dy4_mode_t* m)
{
    switch (mode) {
    case 0: m->rf_Fs = 2.4e6;  m->rf_decim = 10; m->if_Fs = 240e3; m->audio_decim = 5;    m->audio_upsample = 1;   m->block_size = 1024 * 5 * 10 * 2; break;
    case 1: m->rf_Fs = 1.44e6; m->rf_decim = 5;  m->if_Fs = 288e3; m->audio_decim = 8;    m->audio_upsample = 1;   m->block_size = 1024 * 8 * 5 * 2;  break;
    case 2: m->rf_Fs = 2.4e6;  m->rf_decim = 10; m->if_Fs = 240e3; m->audio_decim = 800;  m->audio_upsample = 147; m->block_size = 10 * 800 * 10 * 2;  break;
    case 3: m->rf_Fs = 1.92e6; m->rf_decim = 5;  m->if_Fs = 384e3; m->audio_decim = 1280; m->audio_upsample = 147; m->block_size = 10 * 1280 * 5 * 2;  break;
    default: return -1;
    }
    m->audio_taps = 101 * m->audio_upsample;
    m->if_per_block = m->block_size / 2 / m->rf_decim;
    m->audio_per_block = m->if_per_block * m->audio_upsample / m->audio_decim;
    return 0;
}

void dy4r_lpf_taps(float Fs, float Fc, unsigned short num_taps, int up, float* h)
{
    std::vector<float> v;
    impulseResponseLPF(Fs, Fc, num_taps, v, up);
    std::memcpy(h, v.data(), sizeof(float) * v.size());
}

void dy4r_bpf_taps(float Fs, float Fb, float Fe, unsigned short num_taps, int up, float* h)
{
    std::vector<float> v;
    impulseResponseBPF(Fs, Fb, Fe, num_taps, v, up);
    std::memcpy(h, v.data(), sizeof(float) * v.size());
}

void dy4r_iq_to_float(const uint8_t* raw, long n, float* out)
{
    membuf mb(raw, (size_t)n);
    std::streambuf* old = std::cin.rdbuf(&mb);
    std::cin.clear();
    std::vector<float> v((size_t)n);
    readStdinBlockData((unsigned int)n, 0, v);
    std::cin.rdbuf(old);
    std::memcpy(out, v.data(), sizeof(float) * (size_t)n);
}

void dy4r_block_fir(const float* x, int nx, const float* h, int nh, float* state, int nstate, float* y)
{
    std::vector<float> vx(x, x + nx), vh(h, h + nh), vs(state, state + nstate), vy;
    blockConvolveFIR(vy, vx, vh, vs);
    std::memcpy(y, vy.data(), sizeof(float) * vy.size());
    std::memcpy(state, vs.data(), sizeof(float) * vs.size());
}

void dy4r_decim_fir(int factor, const float* x, int nx, const float* h, int nh, float* state, int nstate, float* y)
{
    std::vector<float> vx(x, x + nx), vh(h, h + nh), vs(state, state + nstate), vy;
    downsampleBlockConvolveFIR(factor, vy, vx, vh, vs);
    std::memcpy(y, vy.data(), sizeof(float) * vy.size());
    std::memcpy(state, vs.data(), sizeof(float) * vs.size());
}

int dy4r_resample_fir(int up, int down, const float* x, int nx, const float* h, int nh, float* state, int nstate, float* y)
{
    std::vector<float> vx(x, x + nx), vh(h, h + nh), vs(state, state + nstate), vy;
    resampleBlockConvolveFIR(up, down, vy, vx, vh, vs);
    std::memcpy(y, vy.data(), sizeof(float) * vy.size());
    std::memcpy(state, vs.data(), sizeof(float) * vs.size());
    return (int)vy.size();
}

void dy4r_fm_demod(const float* I, const float* Q, int n, float* prev_I, float* prev_Q, float* out)
{
    std::vector<float> vi(I, I + n), vq(Q, Q + n), vo;
    fmDemodArctan(vi, vq, *prev_I, *prev_Q, vo);
    std::memcpy(out, vo.data(), sizeof(float) * vo.size());
}

void dy4r_pll(const float* in, int n, float freq, float Fs, float ncoScale, float phaseAdjust,
              float normBandwidth, float* nco, float* st)
{
    std::vector<float> vin(in, in + n), vo;
    fmPLL(vin, freq, Fs, ncoScale, phaseAdjust, normBandwidth, vo, st[0], st[1], st[2], st[3], st[4], st[5]);
    std::memcpy(nco, vo.data(), sizeof(float) * vo.size());
}

void dy4r_delay_block(const float* in, int n, float* state, int nstate, float* out)
{
    std::vector<float> vin(in, in + n), vs(state, state + nstate), vo;
    delayBlock(vin, vs, vo);
    std::memcpy(out, vo.data(), sizeof(float) * vo.size());
    std::memcpy(state, vs.data(), sizeof(float) * vs.size());
}

void dy4r_pcm16(const float* x, long n, int16_t* out)
{
    /* project.cpp:313-316 is inline in main(); restated here and cross-checked
       against the real binary's stdout by make_golden.py */
    for (long k = 0; k < n; k++) {
        if (std::isnan(x[k])) out[k] = 0;
        else out[k] = static_cast<short int>(x[k] * 16384);
    }
}

long dy4r_pipeline(int mode, int stereo, const uint8_t* iq, long nbytes,
                   float* if_out, float* audio_out, int16_t* pcm_out,
                   float* pilot_out, float* nco_out)
{
    dy4_mode_t m;
    if (dy4r_mode_params(mode, &m) != 0) return -1;
    const short num_taps = 101;
    const float rf_Fc = 100e3, audio_Fc = 16e3;
    const float rf_decim = (float)m.rf_decim, audio_Fs = m.if_Fs;
    const float audio_decim = (float)m.audio_decim, audio_upsample = (float)m.audio_upsample;
    const short audio_taps = (short)m.audio_taps;
    const bool mono = !stereo;
    const int block_size = m.block_size;

    /* state sizing and coefficient generation exactly as project.cpp:240-273 */
    RFState rf_states;
    rf_states.i_state_rf.resize(num_taps - 1, 0.0);
    rf_states.q_state_rf.resize(num_taps - 1, 0.0);
    AudioState audio_states;
    audio_states.state_audio.resize(num_taps - 1, 0.0);
    audio_states.pilot_state.resize(num_taps - 1, 0.0);
    audio_states.stereo_state.resize(num_taps - 1, 0.0);
    audio_states.stereo_lowpass_state.resize(num_taps - 1, 0.0);
    audio_states.mono_delay_state.resize(num_taps / 2, 0.0);
    AudioFilter audio_filters;
    PLLState pll_states;
    frontendVectors FrontVectors;
    backendVectors BackVectors;
    std::vector<float> rf_coeff;
    impulseResponseLPF(m.rf_Fs, rf_Fc, num_taps, rf_coeff, 1);
    impulseResponseLPF(audio_Fs * audio_upsample, audio_Fc, audio_taps, audio_filters.audio_coeff, audio_upsample);
    impulseResponseBPF(audio_Fs, 18.5e3, 19.5e3, num_taps, audio_filters.pilot_coeff, 1);
    impulseResponseBPF(audio_Fs, 22e3, 54e3, num_taps, audio_filters.stereo_coeff, 1);

    /* shadow state, used only to expose pilot / NCO (backend() keeps them private) */
    std::vector<float> sh_pilot_state(num_taps - 1, 0.0f), sh_pilot, sh_nco;
    PLLState sh_pll;

    threadSafeQ q;
    std::vector<float> audio_block, left_block, right_block, processed;
    std::vector<float> iq_data(block_size);
    long nblocks = nbytes / block_size;
    const int nif = m.if_per_block, na = m.audio_per_block, nch = stereo ? 2 : 1;

    membuf mb(iq, (size_t)(nblocks * block_size));
    std::streambuf* old = std::cin.rdbuf(&mb);
    std::cin.clear();

    for (long b = 0; b < nblocks; b++) {
        readStdinBlockData(block_size, (unsigned int)b, iq_data);               /* iofunc.cpp:113 */
        frontend(rf_decim, rf_coeff, rf_states, iq_data, FrontVectors, q);      /* project.cpp:72 */
        std::vector<float> fm = q.dequeue();
        if (if_out) std::memcpy(if_out + b * nif, fm.data(), sizeof(float) * nif);
        if (stereo && (pilot_out || nco_out)) {
            blockConvolveFIR(sh_pilot, fm, audio_filters.pilot_coeff, sh_pilot_state);
            fmPLL(sh_pilot, 19e3, audio_Fs, 2.0, 0, 0.01, sh_nco, sh_pll.feedbackI, sh_pll.feedbackQ,
                  sh_pll.integrator, sh_pll.phaseEst, sh_pll.trigOffset, sh_pll.nco_state);
            if (pilot_out) std::memcpy(pilot_out + b * nif, sh_pilot.data(), sizeof(float) * nif);
            if (nco_out) std::memcpy(nco_out + b * nif, sh_nco.data(), sizeof(float) * nif);
        }
        q.enqueue(fm);
        backend(audio_Fs, (int)audio_decim, (int)audio_upsample, audio_filters, audio_states, pll_states,
                audio_block, left_block, right_block, BackVectors, mono, q);    /* project.cpp:95 */
        if (mono) processed.assign(audio_block.begin(), audio_block.end());
        else { processed.resize(2 * na); interleave(left_block, right_block, processed); }
        if (audio_out) std::memcpy(audio_out + b * na * nch, processed.data(), sizeof(float) * na * nch);
        if (pcm_out) dy4r_pcm16(processed.data(), (long)na * nch, pcm_out + b * na * nch);
    }
    std::cin.rdbuf(old);
    std::cin.clear();
    return nblocks;
}

/* ---- Fourier diagnostics: the reference's own fourier.cpp ---------------------------------------------------- */
void dy4r_dft(const float* x, int n, float* Xf)
{
    std::vector<float> xv(x, x + n);
    std::vector<std::complex<float>> X;
    DFT(xv, X);
    for (int i = 0; i < n; i++) { Xf[2 * i] = X[i].real(); Xf[2 * i + 1] = X[i].imag(); }
}
void dy4r_idft(const float* Xf, int n, float* x)
{
    std::vector<std::complex<float>> X(n), xv;
    for (int i = 0; i < n; i++) X[i] = std::complex<float>(Xf[2 * i], Xf[2 * i + 1]);
    IDFT(X, xv);
    for (int i = 0; i < n; i++) { x[2 * i] = xv[i].real(); x[2 * i + 1] = xv[i].imag(); }
}
void dy4r_estimate_psd(const float* samples, long n, int nfft, int Fs, float* freq, float* psd)
{
    std::vector<float> sv(samples, samples + n), f, p;
    estimatePSD(sv, nfft, Fs, f, p);
    std::memcpy(freq, f.data(), sizeof(float) * f.size());
    std::memcpy(psd, p.data(), sizeof(float) * p.size());
}

/* variant 0: FFT_recursive, 1: FFT_improved (level 1), 2: FFT_optimized; twiddles from the reference's compute_twiddles (NFFT/2) */
void dy4r_fft(const float* x, int n, int variant, float* Xf)
{
    std::vector<std::complex<float>> xv(n), X(n), tw(NFFT / 2);
    for (int i = 0; i < n; i++) xv[i] = std::complex<float>(x[2 * i], x[2 * i + 1]);
    compute_twiddles(tw);
    if (variant == 0) FFT_recursive(xv, X);
    else if (variant == 1) FFT_improved(xv, X, tw, 1);
    else FFT_optimized(xv, X, tw);
    for (int i = 0; i < n; i++) { Xf[2 * i] = X[i].real(); Xf[2 * i + 1] = X[i].imag(); }
}
void dy4r_compute_twiddles(int n_tw, float* out)
{
    std::vector<std::complex<float>> tw(n_tw);
    compute_twiddles(tw);
    for (int i = 0; i < n_tw; i++) { out[2 * i] = tw[i].real(); out[2 * i + 1] = tw[i].imag(); }
}

} /* extern "C" */
