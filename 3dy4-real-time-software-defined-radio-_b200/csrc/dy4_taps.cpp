// dy4_taps.cpp — host-side impulse-response design and the mode table.
//
// Restates, for the product, reference src/filter.cpp:14-29 (impulseResponseLPF)
// and :31-49 (impulseResponseBPF), and the switch in src/project.cpp:178-238.
// One-time host work; kept on the host and in the reference's exact mix of
// float and double (float ratios widened to double, double libm sin/cos, each
// tap narrowed through float twice) so the taps are bit-identical — they are
// inputs to kernels whose outputs must be bit-identical.  Build without
// -ffast-math / -march (no FMA contraction).
#include "../../include/dy4_b200.h"

#include <cmath>

namespace {
const double kPi = 3.14159265358979323846;   // include/dy4.h:14

// Hann-like window of the reference: sin^2(i*pi/N), applied in double then narrowed
inline float windowed(float tap, int i, int num_taps, int up_factor)
{
    const double s = std::sin(i * kPi / num_taps);
    return (float)(tap * (s * s) * (float)up_factor);
}
}  // namespace

extern "C" int dy4_lpf_taps(float Fs, float Fc, unsigned short num_taps, int up_factor, float* h)
{
    if (!h || num_taps == 0) return DY4_ERR_ARG;
    const float half_fs = Fs / 2;
    const double norm = Fc / half_fs;                       // float divide, then widened (filter.cpp:18)
    const int centre = (num_taps - 1) / 2;
    const double centre_f = ((float)num_taps - 1.0) / 2.0;  // filter.cpp:24 uses the non-integer centre here
    for (int i = 0; i < num_taps; i++) {
        float tap;
        if (i == centre) {
            tap = (float)norm;
        } else {
            const double arg = kPi * norm * (i - centre_f);
            tap = (float)(norm * std::sin(arg) / arg);
        }
        h[i] = windowed(tap, i, num_taps, up_factor);
    }
    return DY4_OK;
}

extern "C" int dy4_bpf_taps(float Fs, float Fb, float Fe, unsigned short num_taps, int up_factor, float* h)
{
    if (!h || num_taps == 0) return DY4_ERR_ARG;
    const float half_fs = Fs / 2;
    const float mid_f = (Fe + Fb) / 2;
    const double norm_centre = mid_f / half_fs;             // filter.cpp:35
    const double norm_pass = (Fe - Fb) / half_fs;           // filter.cpp:36
    const int centre = (num_taps - 1) / 2;
    const double centre_f = ((float)num_taps - 1.0) / 2.0;
    for (int i = 0; i < num_taps; i++) {
        float tap;
        if (i == centre) {
            tap = (float)norm_pass;
        } else {
            const double arg = kPi * norm_pass / 2 * (i - centre_f);
            tap = (float)(norm_pass * std::sin(arg) / arg);
        }
        tap = (float)(tap * std::cos((i - centre) * kPi * norm_centre));   // filter.cpp:46, integer centre
        h[i] = windowed(tap, i, num_taps, up_factor);
    }
    return DY4_OK;
}

extern "C" int dy4_mode_params(int mode, dy4_mode_params_t* m)
{
    if (!m) return DY4_ERR_ARG;
    int blocks_of;   // audio samples per block before the U factor
    switch (mode) {
    case 0: m->rf_Fs = 2.4e6f;  m->rf_decim = 10; m->if_Fs = 240e3f; m->audio_decim = 5;    m->audio_upsample = 1;   blocks_of = 1024; break;
    case 1: m->rf_Fs = 1.44e6f; m->rf_decim = 5;  m->if_Fs = 288e3f; m->audio_decim = 8;    m->audio_upsample = 1;   blocks_of = 1024; break;
    case 2: m->rf_Fs = 2.4e6f;  m->rf_decim = 10; m->if_Fs = 240e3f; m->audio_decim = 800;  m->audio_upsample = 147; blocks_of = 10;   break;
    case 3: m->rf_Fs = 1.92e6f; m->rf_decim = 5;  m->if_Fs = 384e3f; m->audio_decim = 1280; m->audio_upsample = 147; blocks_of = 10;   break;
    default: return DY4_ERR_ARG;
    }
    m->audio_taps = 101 * m->audio_upsample;
    m->block_size = blocks_of * m->audio_decim * m->rf_decim * 2;
    m->if_per_block = blocks_of * m->audio_decim;
    m->audio_per_block = blocks_of * m->audio_upsample;
    return DY4_OK;
}


// ---- RDS path taps: the Python model's (model/fmMonoBlock.py:488-514, model/fmRRC.py:13-49), in double --------------
// scipy.signal.firwin(numtaps, [lo, hi] or cutoff, window='hann'): ideal band response (difference of sincs) times a
// symmetric Hann (window 0) or Hamming (window 1, scipy's default) window, scaled to unit gain at the band centre (band-pass) or at DC (low-pass).  Frequencies are
// normalised to Nyquist (1.0 = Fs/2), as firwin takes them.  lo <= 0 means low-pass with cutoff hi.
extern "C" int dy4_firwin(int num_taps, double lo, double hi, int window, double* h)
{
    if (!h || num_taps < 2 || hi <= 0.0 || hi >= 1.0 || lo >= hi || window < 0 || window > 1) return DY4_ERR_ARG;
    const bool lowpass = lo <= 0.0;
    const double left = lowpass ? 0.0 : lo, right = hi;
    const double alpha = 0.5 * (num_taps - 1);
    auto sinc = [](double x) { return x == 0.0 ? 1.0 : std::sin(kPi * x) / (kPi * x); };
    for (int i = 0; i < num_taps; i++) {
        const double mm = i - alpha;
        double v = right * sinc(right * mm) - left * sinc(left * mm);
        const double cw = std::cos(2.0 * kPi * i / (num_taps - 1));
        v *= window == 0 ? 0.5 - 0.5 * cw : 0.54 - 0.46 * cw;           // scipy 'hann' / 'hamming' (symmetric)
        h[i] = v;
    }
    const double scale_frequency = lowpass ? 0.0 : 0.5 * (left + right);
    double sum = 0.0;
    for (int i = 0; i < num_taps; i++) sum += h[i] * std::cos(kPi * (i - alpha) * scale_frequency);
    for (int i = 0; i < num_taps; i++) h[i] /= sum;
    return DY4_OK;
}

// model/fmRRC.py: root-raised-cosine, beta 0.90, T = 1/2375 s, centred on tap N/2 (float division)
extern "C" int dy4_rrc_taps(double Fs, int num_taps, double* h)
{
    if (!h || num_taps < 1 || Fs <= 0) return DY4_ERR_ARG;
    const double T = 1 / 2375.0, beta = 0.90;
    for (int k = 0; k < num_taps; k++) {
        const double t = (k - num_taps / 2.0) / Fs;
        if (t == 0.0) h[k] = 1.0 + beta * ((4 / kPi) - 1);
        else if (t == -T / (4 * beta) || t == T / (4 * beta))
            h[k] = (beta / std::sqrt(2.0)) * (((1 + 2 / kPi) * std::sin(kPi / (4 * beta))) + ((1 - 2 / kPi) * std::cos(kPi / (4 * beta))));
        else
            h[k] = (std::sin(kPi * t * (1 - beta) / T) + 4 * beta * (t / T) * std::cos(kPi * t * (1 + beta) / T)) /
                   (kPi * t * (1 - (4 * beta * t / T) * (4 * beta * t / T)) / T);
    }
    return DY4_OK;
}
