/*
 * dy4_oracle.c — plain-C CPU restatement of the reference receiver's hot path.
 *
 * TEST INFRASTRUCTURE ONLY (see dy4_oracle.h).  This is the checker the CUDA
 * path is compared with; it is never a fallback for it.
 *
 * Every function restates one function of the reference and cites it.
 * Citations are relative to /root/reference/.  What is being restated is the
 * ARITHMETIC (operation order, where float narrows and where double widens),
 * because that is what bit-level parity depends on:
 *   - FIR sums are float, ascending tap index, unfused multiply then add
 *     (the reference's Makefile flags emit no FMA; build this file with
 *     -ffp-contract=off and no -march so the same holds here).
 *   - tap design and the PLL call the DOUBLE libm functions on float-widened
 *     arguments and narrow the results back to float.
 *
 * Parity status: PINNED against tests/golden/*.npz, which are outputs of the
 * reference's own compiled code (oracle/_ref/libdy4ref.so) — the reference
 * itself ships no golden vectors for this path.
 */
#include "dy4_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

#define DY4_PI 3.14159265358979323846 /* include/dy4.h:14 — a double literal */
#ifndef DY4_LOG10
#define DY4_LOG10(v) log10((double)(v))   /* fourier.cpp:75 calls the unqualified ::log10 on a float (see tests/test_oracle.py: checked against the reference build) */
#endif
#define DY4_NUM_TAPS 101              /* src/project.cpp:142 */

/* ---- mode table: src/project.cpp:178-238 -------------------------------- */
int DY4_FN(mode_params)(int mode, dy4_mode_t* m)
{
    switch (mode) {
    case 0: m->rf_Fs = 2.4e6f;  m->rf_decim = 10; m->if_Fs = 240e3f; m->audio_decim = 5;    m->audio_upsample = 1;   break;
    case 1: m->rf_Fs = 1.44e6f; m->rf_decim = 5;  m->if_Fs = 288e3f; m->audio_decim = 8;    m->audio_upsample = 1;   break;
    case 2: m->rf_Fs = 2.4e6f;  m->rf_decim = 10; m->if_Fs = 240e3f; m->audio_decim = 800;  m->audio_upsample = 147; break;
    case 3: m->rf_Fs = 1.92e6f; m->rf_decim = 5;  m->if_Fs = 384e3f; m->audio_decim = 1280; m->audio_upsample = 147; break;
    default: return -1;
    }
    m->audio_taps = DY4_NUM_TAPS * m->audio_upsample;
    /* modes 0/1: 1024*D*rf_decim*2 ; modes 2/3: 10*D*rf_decim*2 */
    m->block_size = (m->audio_upsample == 1 ? 1024 : 10) * m->audio_decim * m->rf_decim * 2;
    m->if_per_block = m->block_size / 2 / m->rf_decim;
    m->audio_per_block = (int)((m->if_per_block / (float)m->audio_decim) * m->audio_upsample);
    return 0;
}

/* ---- impulseResponseLPF: src/filter.cpp:14-29 --------------------------- */
void DY4_FN(lpf_taps)(float Fs, float Fc, unsigned short num_taps, int up, float* h)
{
    /* float divide, widened afterwards (:18) */
    double norm_cutoff = Fc / (Fs / 2);
    int mid = (num_taps - 1) / 2;
    for (int i = 0; i < num_taps; i++) {
        float hi;
        if (i == mid) {
            hi = (float)norm_cutoff;                                         /* :21 */
        } else {
            double arg = DY4_PI * norm_cutoff * (i - ((float)num_taps - 1.0) / 2.0); /* :24 */
            hi = (float)(norm_cutoff * sin(arg) / arg);                      /* :25 */
        }
        /* second narrowing through float (:27); window = sin^2 */
        h[i] = (float)(hi * pow(sin(i * DY4_PI / num_taps), 2) * (float)up);
    }
}

/* ---- impulseResponseBPF: src/filter.cpp:31-49 --------------------------- */
void DY4_FN(bpf_taps)(float Fs, float Fb, float Fe, unsigned short num_taps, int up, float* h)
{
    double norm_center = ((Fe + Fb) / 2) / (Fs / 2);                         /* :35 float math */
    double norm_pass = (Fe - Fb) / (Fs / 2);                                 /* :36 */
    int mid = (num_taps - 1) / 2;
    for (int i = 0; i < num_taps; i++) {
        float hi;
        if (i == mid) {
            hi = (float)norm_pass;                                           /* :40 */
        } else {
            double arg = DY4_PI * norm_pass / 2 * (i - ((float)num_taps - 1.0) / 2.0); /* :43 */
            hi = (float)(norm_pass * sin(arg) / arg);                        /* :44 */
        }
        hi = (float)(hi * cos((i - mid) * DY4_PI * norm_center));            /* :46 */
        h[i] = (float)(hi * pow(sin(i * DY4_PI / num_taps), 2) * (float)up); /* :47 */
    }
}

/* ---- readStdinBlockData arithmetic: src/iofunc.cpp:117-119 -------------- */
void DY4_FN(iq_to_float)(const uint8_t* raw, long n, float* out)
{
    for (long k = 0; k < n; k++)
        out[k] = (float)(((int)raw[k] - 128) / 128.0);
}

/* One FIR tap.  The reference (and this oracle as built by oracle/Makefile)
 * rounds the product and the sum separately.  -DDY4_ORACLE_FIR_FMA builds an
 * EXPERIMENT variant with a fused multiply-add, used only by
 * tools/fma_sensitivity.py to show why the CUDA path keeps the unfused order. */
#ifdef DY4_ORACLE_FIR_FMA
#define DY4_TAP(acc, hv, xv) fmaf((hv), (xv), (acc))
#else
static inline float dy4_tap(float acc, float hv, float xv) { float prod = hv * xv; return acc + prod; }
#define DY4_TAP(acc, hv, xv) dy4_tap((acc), (hv), (xv))
#endif

/* value of the block-extended signal (state followed by x) at index i, i may be negative */
static inline float ext(const float* x, const float* state, int nstate, int i)
{
    return i >= 0 ? x[i] : state[nstate + i];
}

/* ---- blockConvolveFIR: src/filter.cpp:66-83 ------------------------------ */
void DY4_FN(block_fir)(const float* x, int nx, const float* h, int nh, float* state, int nstate, float* y)
{
    for (int n = 0; n < nx; n++) {
        float acc = 0.0f;
        for (int k = 0; k < nh; k++) {
            acc = DY4_TAP(acc, h[k], ext(x, state, nstate, n - k));
        }
        y[n] = acc;
    }
    memcpy(state, x + nx - nstate, sizeof(float) * (size_t)nstate);          /* :82 */
}

/* ---- downsampleBlockConvolveFIR: src/filter.cpp:123-140 ------------------ */
void DY4_FN(decim_fir)(int factor, const float* x, int nx, const float* h, int nh, float* state, int nstate, float* y)
{
    for (int n = 0; n < nx; n += factor) {
        float acc = 0.0f;
        for (int k = 0; k < nh; k++) {
            acc = DY4_TAP(acc, h[k], ext(x, state, nstate, n - k));
        }
        y[n / factor] = acc;
    }
    memcpy(state, x + nx - nstate, sizeof(float) * (size_t)nstate);          /* :139 */
}

/* ---- resampleBlockConvolveFIR: src/filter.cpp:142-173 -------------------- */
int DY4_FN(resample_fir)(int up, int down, const float* x, int nx, const float* h, int nh, float* state, int nstate, float* y)
{
    int ny = (int)((nx / (float)down) * up);                                 /* :149 */
    for (int i = 0; i < ny; i++) y[i] = 0.0f;
    for (long n = 0; n < (long)nx * up; n += down) {                         /* :158 */
        int phase = (int)(n % up);
        float acc = 0.0f;
        for (long k = phase; k < nh; k += up) {                              /* only the non-zero taps of the expanded signal */
            long d = n - k;                                                  /* multiple of `up` by construction */
            float xv = d >= 0 ? x[d / up] : state[nstate - (int)((k - n) / up)];
            acc = DY4_TAP(acc, h[k], xv);
        }
        y[n / down] = acc;
    }
    memcpy(state, x + nx - nstate, sizeof(float) * (size_t)nstate);          /* :169 */
    return ny;
}

/* ---- fmDemodArctan: src/filter.cpp:85-102 (a derivative discriminator) --- */
void DY4_FN(fm_demod)(const float* I, const float* Q, int n, float* prev_I, float* prev_Q, float* out)
{
    float pi = *prev_I, pq = *prev_Q;
    for (int k = 0; k < n; k++) {
        /* std::pow(float,int) promotes to double; the sum narrows to float (:88) */
        float denom = (float)(pow((double)I[k], 2.0) + pow((double)Q[k], 2.0));
        if (denom == 0) {
            out[k] = 0.0f;
        } else {
            float dq = Q[k] - pq, di = I[k] - pi;
            float a = I[k] * dq;
            float b = Q[k] * di;
            float num = a - b;
            out[k] = num / denom;                                            /* :94 / :97 */
        }
        pi = I[k]; pq = Q[k];                                                /* k-1 sample, also when denom==0 */
    }
    *prev_I = I[n - 1];
    *prev_Q = Q[n - 1];
}

/* ---- fmPLL: src/filter.cpp:174-228 --------------------------------------- */
void DY4_FN(pll)(const float* in, int n, float freq, float Fs, float ncoScale, float phaseAdjust,
                 float normBandwidth, float* nco, float* st)
{
    float Cp = 2.666f, Ci = 3.555f;                                          /* :175-176 */
    float Kp = normBandwidth * Cp;
    float Ki = normBandwidth * normBandwidth * Ci;
    float fbI = st[0], fbQ = st[1], integ = st[2], phase = st[3], trigOffset = st[4], nco_state = st[5];
    float ratio = freq / Fs;                                                 /* float divide inside :214 */

    nco[0] = nco_state;                                                      /* :184 */
    for (int k = 0; k < n; k++) {
        float errI = (in[k] == 0 ? 1 : in[k]) * fbI;                         /* :192 */
        float errQ = in[k] * (-1 * fbQ);                                     /* :193 */
        float errD = (float)atan2((double)errQ, (double)errI);               /* :200 double atan2 */
        float t0 = Ki * errD;
        integ = integ + t0;                                                  /* :207 */
        float t1 = Kp * errD;
        float t2 = t1 + integ;
        phase = phase + t2;                                                  /* :210 */
        trigOffset = trigOffset + 1.0f;                                      /* :213 float counter */
        double arg = 2 * DY4_PI * (double)ratio * (double)trigOffset + (double)phase;
        float trigArg = (float)arg;                                          /* :214 narrowed to float */
        fbI = (float)cos((double)trigArg);                                   /* :216 */
        fbQ = (float)sin((double)trigArg);                                   /* :217 */
        float narg = trigArg * ncoScale + phaseAdjust;                       /* float */
        float nv = (float)cos((double)narg);
        if (k == n - 1) nco_state = nv; else nco[k + 1] = nv;                /* :218-222 */
    }
    st[0] = fbI; st[1] = fbQ; st[2] = integ; st[3] = phase; st[4] = trigOffset; st[5] = nco_state;
}

/* ---- delayBlock: src/filter.cpp:229-251 ---------------------------------- */
void DY4_FN(delay_block)(const float* in, int n, float* state, int nstate, float* out)
{
    memcpy(out, state, sizeof(float) * (size_t)nstate);
    memcpy(out + nstate, in, sizeof(float) * (size_t)(n - nstate));
    memcpy(state, in + n - nstate, sizeof(float) * (size_t)nstate);
}

/* ---- float -> int16: src/project.cpp:313-316 ----------------------------- */
void DY4_FN(pcm16)(const float* x, long n, int16_t* out)
{
    for (long k = 0; k < n; k++) {
        if (isnan(x[k])) out[k] = 0;
        else out[k] = (int16_t)(int32_t)(x[k] * 16384);  /* truncation toward zero; UB in the reference if |x| >= 2 */
    }
}

/* ---- the block loop: src/project.cpp:240-318, frontend :72-93, backend :95-134 */
long DY4_FN(pipeline)(int mode, int stereo, const uint8_t* iq, long nbytes,
                      float* if_out, float* audio_out, int16_t* pcm_out,
                      float* pilot_out, float* nco_out)
{
    dy4_mode_t m;
    if (DY4_FN(mode_params)(mode, &m) != 0) return -1;
    const int NT = DY4_NUM_TAPS, NS = NT - 1, ND = NT / 2;
    const int bs = m.block_size, np = bs / 2, nif = m.if_per_block, na = m.audio_per_block;
    long nblocks = nbytes / bs;

    float* rf_h = malloc(sizeof(float) * NT);
    float* audio_h = malloc(sizeof(float) * (size_t)m.audio_taps);
    float* pilot_h = malloc(sizeof(float) * NT);
    float* stereo_h = malloc(sizeof(float) * NT);
    DY4_FN(lpf_taps)(m.rf_Fs, 100e3f, NT, 1, rf_h);                                              /* :262 */
    DY4_FN(lpf_taps)(m.if_Fs * (float)m.audio_upsample, 16e3f, (unsigned short)m.audio_taps, m.audio_upsample, audio_h); /* :265 */
    DY4_FN(bpf_taps)(m.if_Fs, 18.5e3f, 19.5e3f, NT, 1, pilot_h);                                 /* :272 */
    DY4_FN(bpf_taps)(m.if_Fs, 22e3f, 54e3f, NT, 1, stereo_h);                                    /* :273 */

    /* carried state, all zero at start (:240-255) */
    float st_i[100] = {0}, st_q[100] = {0}, prev_I = 0, prev_Q = 0;
    float st_mono[100] = {0}, st_diff[100] = {0}, st_pilot[100] = {0}, st_stereo[100] = {0}, st_delay[50] = {0};
    float pll_st[6] = {1.0f, 0.0f, 0.0f, 0.0f, 0.0f, 1.0f};

    float* xf = malloc(sizeof(float) * (size_t)bs);
    float* xi = malloc(sizeof(float) * (size_t)np);
    float* xq = malloc(sizeof(float) * (size_t)np);
    float* di = malloc(sizeof(float) * (size_t)nif);
    float* dq = malloc(sizeof(float) * (size_t)nif);
    float* fm = malloc(sizeof(float) * (size_t)nif);
    float* delayed = malloc(sizeof(float) * (size_t)nif);
    float* pilot = malloc(sizeof(float) * (size_t)nif);
    float* sband = malloc(sizeof(float) * (size_t)nif);
    float* nco = malloc(sizeof(float) * (size_t)nif);
    float* mixed = malloc(sizeof(float) * (size_t)nif);
    float* mono = malloc(sizeof(float) * (size_t)na);
    float* diff = malloc(sizeof(float) * (size_t)na);
    float* outb = malloc(sizeof(float) * (size_t)na * 2);
    const int nch = stereo ? 2 : 1;

    for (long b = 0; b < nblocks; b++) {
        DY4_FN(iq_to_float)(iq + b * bs, bs, xf);                                               /* :292 */
        for (int i = 0; i < np; i++) { xi[i] = xf[2 * i]; xq[i] = xf[2 * i + 1]; }              /* :78-81 */
        DY4_FN(decim_fir)(m.rf_decim, xi, np, rf_h, NT, st_i, NS, di);                          /* :86 */
        DY4_FN(decim_fir)(m.rf_decim, xq, np, rf_h, NT, st_q, NS, dq);                          /* :87 */
        DY4_FN(fm_demod)(di, dq, nif, &prev_I, &prev_Q, fm);                                    /* :90 */
        if (if_out) memcpy(if_out + b * nif, fm, sizeof(float) * (size_t)nif);

        DY4_FN(delay_block)(fm, nif, st_delay, ND, delayed);                                    /* :114 */
        DY4_FN(resample_fir)(m.audio_upsample, m.audio_decim, delayed, nif, audio_h, m.audio_taps, st_mono, NS, mono); /* :116 */
        if (stereo) {
            DY4_FN(block_fir)(fm, nif, pilot_h, NT, st_pilot, NS, pilot);                       /* :120 */
            DY4_FN(block_fir)(fm, nif, stereo_h, NT, st_stereo, NS, sband);                     /* :121 */
            DY4_FN(pll)(pilot, nif, 19e3f, m.if_Fs, 2.0f, 0.0f, 0.01f, nco, pll_st);            /* :123 */
            for (int i = 0; i < nif; i++) mixed[i] = nco[i] * sband[i] * 2;                     /* :126, filter.cpp:264 */
            DY4_FN(resample_fir)(m.audio_upsample, m.audio_decim, mixed, nif, audio_h, m.audio_taps, st_diff, NS, diff); /* :129 */
            for (int i = 0; i < na; i++) {                                                      /* :131-132, :310 */
                outb[2 * i] = mono[i] + diff[i];
                outb[2 * i + 1] = mono[i] - diff[i];
            }
            if (pilot_out) memcpy(pilot_out + b * nif, pilot, sizeof(float) * (size_t)nif);
            if (nco_out) memcpy(nco_out + b * nif, nco, sizeof(float) * (size_t)nif);
        } else {
            memcpy(outb, mono, sizeof(float) * (size_t)na);                                     /* :308 */
        }
        if (audio_out) memcpy(audio_out + b * na * nch, outb, sizeof(float) * (size_t)na * nch);
        if (pcm_out) DY4_FN(pcm16)(outb, (long)na * nch, pcm_out + b * na * nch);               /* :313-316 */
    }

    free(rf_h); free(audio_h); free(pilot_h); free(stereo_h);
    free(xf); free(xi); free(xq); free(di); free(dq); free(fm); free(delayed); free(pilot);
    free(sband); free(nco); free(mixed); free(mono); free(diff); free(outb);
    return nblocks;
}


/* ------------------------------------------------------------------------------------------------------------------
 * Fourier diagnostics — reference src/fourier.cpp (not on the receiver's hot path).
 * std::exp(std::complex<float>(0, a)) is cexpf(0 + ia) = (cosf(a), sinf(a)); float * complex<float> multiplies both
 * parts by the float; complex<float> * complex<float> is libgcc's __mulsc3 (four products, a difference and a sum).
 * ---------------------------------------------------------------------------------------------------------------- */
void DY4_FN(dft)(const float* x, int n, float* Xf)
{
    for (int m = 0; m < n; m++) {                                  /* fourier.cpp:16 */
        float re = 0.0f, im = 0.0f;
        for (int k = 0; k < n; k++) {                              /* :17 */
            const float a = (float)(-2 * DY4_PI * (k * m) / (double)(size_t)n);   /* :18: double expression narrowed by the complex<float> constructor */
            const float wr = cosf(a), wi = sinf(a);
            const float pr = wr * x[k], pi_ = wi * x[k];           /* :19 */
            re = re + pr; im = im + pi_;
        }
        Xf[2 * m] = re; Xf[2 * m + 1] = im;
    }
}

void DY4_FN(idft)(const float* Xf, int n, float* x)
{
    for (unsigned k = 0; k < (unsigned)n; k++) {                   /* :100 */
        float re = 0.0f, im = 0.0f;
        for (unsigned m = 0; m < (unsigned)n; m++) {               /* :101 */
            const float a = (float)(2 * DY4_PI * (k * m) / (double)(size_t)n);    /* :102 */
            const float wr = cosf(a), wi = sinf(a);
            const float ar = Xf[2 * m], ai = Xf[2 * m + 1];
            const float ac = ar * wr, bd = ai * wi, ad = ar * wi, bc = ai * wr;   /* :103, __mulsc3 */
            re = re + (ac - bd); im = im + (ad + bc);
        }
        x[2 * k] = re / (float)n; x[2 * k + 1] = im / (float)n;    /* :105 */
    }
}

/* compute_twiddles, fourier.cpp:125-130 (NFFT = 512, include/dy4.h:18): the angle is a double expression of float(k), narrowed by
 * the complex<float> constructor */
void DY4_FN(compute_twiddles)(int n_tw, float* out)
{
    for (int k = 0; k < n_tw; k++) {
        const float a = (float)(-2 * DY4_PI * (float)k / 512);
        out[2 * k] = cosf(a); out[2 * k + 1] = sinf(a);
    }
}

/* FFT_recursive :132-160 (variant 0), FFT_improved :162-187 (1), FFT_optimized :189-211 (2) — one restatement: the three run
 * the same radix-2 decimation-in-time butterflies (even/odd recursion == bit-reversal + levels) and differ only in the value
 * of a butterfly's twiddle: exp(i float(-2 PI float(k) / size)) computed per level (:152), or twiddles[k * 512/size] (:181, :203). */
void DY4_FN(fft)(const float* x, int n, int variant, float* Xf)
{
    int levels = 0;
    while ((1 << levels) < n) levels++;
    for (int i = 0; i < n; i++) {                                   /* :194-196 / the recursion's even-odd splits */
        int r = 0;
        for (int b = 0; b < levels; b++) if (i & (1 << b)) r |= 1 << (levels - 1 - b);
        Xf[2 * i] = x[2 * r]; Xf[2 * i + 1] = x[2 * r + 1];
    }
    for (int l = 0; l < levels; l++) {
        const int half = 1 << l, size = 2 * half;
        for (int p = 0; p < n; p += size)
            for (int j = 0; j < half; j++) {
                float a;
                if (variant == 0) a = (float)(-2 * DY4_PI * (float)j / (size_t)size);               /* :152 */
                else a = (float)(-2 * DY4_PI * (float)(j * (n / size)) / 512);                      /* :127 at index j * 2^(level-1) */
                const float wr = cosf(a), wi = sinf(a);
                const int k = p + j;
                const float orr = Xf[2 * (k + half)], oi = Xf[2 * (k + half) + 1];
                const float ac = wr * orr, bd = wi * oi, ad = wr * oi, bc = wi * orr;              /* __mulsc3 */
                const float tr = ac - bd, ti = ad + bc;
                const float er = Xf[2 * k], ei = Xf[2 * k + 1];
                Xf[2 * k] = er + tr; Xf[2 * k + 1] = ei + ti;                                        /* :154 */
                Xf[2 * (k + half)] = er - tr; Xf[2 * (k + half) + 1] = ei - ti;                      /* :155 */
            }
    }
}

void DY4_FN(estimate_psd)(const float* samples, long n, int nfft, int Fs, float* freq, float* psd)
{
    const float df = (float)Fs / (float)nfft;                      /* :40 */
    const int half = nfft / 2;
    for (int i = 0; i < half; i++) freq[i] = df * (float)i;        /* :43-45 */
    float* hann = (float*)malloc(sizeof(float) * (size_t)nfft);
    for (int i = 0; i < nfft; i++) hann[i] = (float)pow(sin((float)i * DY4_PI / (float)nfft), 2);   /* :49-51 */
    const int nseg = (int)floorf((float)n / (float)nfft);          /* :54 */
    float* win = (float*)malloc(sizeof(float) * (size_t)nfft);
    float* Xf = (float*)malloc(sizeof(float) * 2 * (size_t)nfft);
    float* list = (float*)malloc(sizeof(float) * (size_t)half * (size_t)(nseg > 0 ? nseg : 1));
    for (int s = 0; s < nseg; s++) {                               /* :58 */
        for (int j = 0; j < nfft; j++) win[j] = samples[(long)s * nfft + j] * hann[j];   /* :61-63 */
        DY4_FN(dft)(win, nfft, Xf);                                /* :67 */
        for (int j = 0; j < half; j++) {                           /* :73-76 */
            const float mag = hypotf(Xf[2 * j], Xf[2 * j + 1]);    /* std::abs(complex<float>) */
            float v = (float)((1.0 / ((float)Fs * (float)nfft / 2.0)) * 2.0 * pow((double)mag, 2));
            v = (float)(10.0 * DY4_LOG10(v));
            list[(size_t)s * half + j] = v;
        }
    }
    for (int i = 0; i < half; i++) {                               /* :84-89 */
        float acc = 0.0f;
        for (int s = 0; s < nseg; s++) acc = acc + list[(size_t)i + (size_t)s * half];
        psd[i] = acc / (float)nseg;
    }
    free(hann); free(win); free(Xf); free(list);
}
