// dy4_plltab.h — the PLL recurrence (src/filter.cpp:174-228) with its transcendental work moved OFF the serial chain.
//
// What the recurrence really depends on.  The detector input is eI = x*fbI, eQ = x*(-fbQ) with (fbI,fbQ) =
// (cos,sin)(trigArg) — so errorD = atan2(eQ,eI) (filter.cpp:200) is a function of TWO things only: the input sample
// x[k] (known before the loop runs) and the previous float trigArg.  And trigArg is a FLOAT (filter.cpp:188,214): it
// lives on a grid of spacing u = ulp(trigArg) (2^-11 rad after 8 k samples, 2^-3 after 10 s).  A cheap predictor —
// the same loop in plain double arithmetic with the detector replaced by its closed form  wrap(pi*[x<0] - trigArg),
// no transcendental at all — stays within ONE grid point of the true float trajectory (the loop is contracting; the
// only thing the predictor lacks is the float rounding of trigArg, |q| <= u/2, which the loop filter attenuates).
//
// So the work is split three ways (dy4_pll.cu):
//   1. k_pll_predict   time-parallel (segments with a warm-up), serial only in cheap double adds: predicted trigArg
//   2. k_pll_table     fully parallel: for every sample the EXACT errorD of the next step for the two neighbouring floats
//                      the prediction lies between (the dy4_pllmath.h sincos + detector, unchanged arithmetic)
//   3. k_pll_tab       the serial loop: per sample ONE float compare of phaseEst with a precomputed threshold picks the
//                      grid point the true trigArg falls on, and two speculative loop-filter updates (float adds) are
//                      selected from — tens of cycles instead of ~445.  Whenever the pick is not certain (outside the
//                      two candidates, within a rounding-error guard band of the threshold, binade edges, start-up) the
//                      thread evaluates that step directly with dy4_pllmath.h, so the result is the reference's bit for
//                      bit by construction.
//
// Everything here is a fixed sequence of IEEE operations compiled for host and device; tests/host/plltab_host.c runs the
// three parts on the host against the reference recurrence with glibc (tests/test_host_logic.py).
#pragma once
#include "dy4_pllmath.h"

#if defined(__CUDA_ARCH__)
#define DY4_FMULF(a, b) __fmul_rn((a), (b))
#define DY4_FADDF(a, b) __fadd_rn((a), (b))
#define DY4_D2F(a) __double2float_rn(a)
#else
#define DY4_FMULF(a, b) ((float)((float)(a) * (float)(b)))   /* host: -ffp-contract=off */
#define DY4_FADDF(a, b) ((float)((float)(a) + (float)(b)))
#define DY4_D2F(a) ((float)(a))
#endif

// Samples of a stream evaluated directly before the table takes over: while the loop acquires lock the detector
// crosses +-pi, where one ulp decides the sign of a 2*pi jump, and nothing predicts that.
#define DY4_TAB_EARLY 1536

// One table row per sample k of a launch: what the serial loop needs to go from state_k to state_{k+1}.
// 32 bytes (the first eight fields), read as two 16-byte words.  The prediction th_hat lies between two neighbouring floats
// lo < hi = lo + u, and so — almost always — does the true trigArg.  With t = (lo - RN_d(w*T_k)) + u/2:
// trigArg_k = RN_f(RN_d(w*T_k) + phase_k) is lo for phase_k in (t - u, t) and hi for phase_k in (t, t + u).
typedef struct {
    float t, hu;      // the threshold rounded to float (NaN: row not usable); u/2
    float hm, pad;    // hm = u/2 - m, m the guard band (rounding-error budget of the pick): certain iff | |phase - t| - u/2 | < hm
    float a_lo, a_hi; // Ki*errorD of step k+1 if trigArg_k = lo, hi   (filter.cpp:207's product)
    float b_lo, b_hi; // Kp*errorD                                     (filter.cpp:210's product)
    float lo, u;      // (not stored on the device) the lower candidate, grid spacing of its binade
} dy4_tabrow_t;

// float counter of filter.cpp:213 as a double: exact below 2^24, sticks there (16777217 rounds back to 16777216)
DY4_HD double dy4_pll_count(double T0, int steps) { return fmin(T0 + (double)steps, 16777216.0); }

// loop filter + phase accumulator, filter.cpp:207,210 (float, unfused, this order); a = Ki*errorD, b = Kp*errorD
DY4_HD void dy4_pll_filter_ab(float a, float b, float* integ, float* phase)
{
    *integ = DY4_FADDF(*integ, a);
    *phase = DY4_FADDF(*phase, DY4_FADDF(b, *integ));
}
DY4_HD void dy4_pll_filter(float eD, float Kp, float Ki, float* integ, float* phase)
{
    dy4_pll_filter_ab(DY4_FMULF(Ki, eD), DY4_FMULF(Kp, eD), integ, phase);
}

// exact trigArg of filter.cpp:214 as a float value
DY4_HD float dy4_pll_trigarg(double w, double T, float phase) { return DY4_D2F(DY4_ADD(DY4_MUL(w, T), (double)phase)); }

DY4_HD int dy4_fast_ok(float x) { const float ax = fabsf(x); return ax > 1e-20f && ax < 1e20f; }

// ---- 1. predictor: one step in plain double, no transcendental -----------------------------------------------------
// th_prev: (predicted) trigArg of the previous step; returns the predicted trigArg of this step.
DY4_HD double dy4_pred_step(float x, double th_prev, double wT, double Kp, double Ki, double* integ, double* phase)
{
    const double PI_ = 3.14159265358979323846, TWO_PI = 6.28318530717958647692, INV_2PI = 0.15915494309189533577;
    double a = (x < 0.0f ? PI_ : 0.0) - th_prev;                 // angle of x*exp(-i*th_prev)
    const double big = 6755399441055744.0;                       // 1.5*2^52: round to nearest integer
    const double n = (a * INV_2PI + big) - big;
    a = fma(-n, TWO_PI, a);                                      // into (-pi, pi]; accuracy ~1e-10 is plenty for a prediction
    // x == 0: the reference's detector sees (1*fbI, +-0) (filter.cpp:192-193): 0 in the right half plane, +-pi in the left
    if (x == 0.0f) a = (fabs(a) < 0.5 * PI_) ? 0.0 : (a > 0.0 ? PI_ : -PI_);
    *integ = fma(Ki, a, *integ);
    *phase = *phase + fma(Kp, a, *integ);
    return wT + *phase;
}

// ---- 2. table row --------------------------------------------------------------------------------------------------
// errorD of the step that follows trigArg = th (a float value) when its input sample is x: the arithmetic of
// pll_step_fast / detector_libm in dy4_pll.cu, i.e. of filter.cpp:192-200,216-217.
DY4_HD float dy4_next_errorD(double th, float x)
{
    dy4_nco_t o;
    dy4_sincos_nco_v(th, x < 0.0f, &o, 0);
    const float fbI = DY4_D2F(o.c), fbQ = DY4_D2F(o.s);                                   // :216-217
    if (dy4_fast_ok(x)) {
        const float eI = DY4_FMULF(x, fbI);                                               // :192 (x != 0)
        const float eQ = DY4_FMULF(x, -fbQ);                                              // :193
        return DY4_D2F(dy4_detector_atan2((double)eQ, (double)eI, &o, dy4_recip(x)));     // :200
    }
    const float eI = DY4_FMULF((x == 0.0f ? 1.0f : x), fbI);
    const float eQ = DY4_FMULF(x, -fbQ);
    return DY4_D2F(atan2((double)eQ, (double)eI));
}

#if defined(__CUDA_ARCH__)
DY4_HD int dy4_f2i_bits(float v) { return __float_as_int(v); }
DY4_HD float dy4_i2f_bits(int v) { return __int_as_float(v); }
#else
DY4_HD int dy4_f2i_bits(float v) { int b; memcpy(&b, &v, 4); return b; }
DY4_HD float dy4_i2f_bits(int v) { float f; memcpy(&f, &v, 4); return f; }
#endif

// th_hat: predicted trigArg of this sample; wT = RN_d(w*T_k);
// x_next: input of step k+1 (has_next == 0 for the last sample of a launch: products unused).
// `force_invalid`: rows the serial loop must evaluate directly whatever the prediction says.
DY4_HD void dy4_tab_make_row(double th_hat, double wT, float x_next, int has_next, int force_invalid, float Kp, float Ki, dy4_tabrow_t* r)
{
    const float c = DY4_D2F(th_hat);
    const int bits = dy4_f2i_bits(c);
    const int expo = (bits >> 23) & 0xff, mant = bits & 0x7fffff;
    // usable: positive normal float with both neighbours in the same binade and u in a sane range (2^-40 .. 2^40)
    const int ok = !force_invalid && bits > 0 && expo >= 110 && expo <= 190 && mant >= 2 && mant <= 0x7ffffd;
    const float u = dy4_i2f_bits((ok ? expo - 23 : 127) << 23);
    const float lo = (th_hat >= (double)c) ? c : DY4_FADDF(c, -u), hi = DY4_FADDF(lo, u);
    const double hu = DY4_MUL(0.5, (double)u);
    const float t = DY4_D2F(DY4_ADD(DY4_SUB((double)lo, wT), hu));
    r->lo = lo; r->u = u; r->hu = DY4_D2F(hu); r->pad = 0.0f;
    r->t = ok ? t : dy4_i2f_bits(0x7fc00000);
    // t is within 2^-24|t| (half an ulp) of the exact threshold; the pick's two float subtractions add at most 2^-24 u, the
    // reference's double add 2^-29 u: the guard band is two half-ulps of t plus 2^-23 u.
    const float m = DY4_FADDF(DY4_FMULF(1.1920928955078125e-07f, fabsf(t)), DY4_FMULF(1.1920928955078125e-07f, u));
    r->hm = DY4_FADDF(r->hu, -m);
    r->a_lo = r->a_hi = r->b_lo = r->b_hi = 0.0f;
    if (ok && has_next) {
        const float e_lo = dy4_next_errorD((double)lo, x_next);
        const float e_hi = dy4_next_errorD((double)hi, x_next);
        r->a_lo = DY4_FMULF(Ki, e_lo); r->a_hi = DY4_FMULF(Ki, e_hi);
        r->b_lo = DY4_FMULF(Kp, e_lo); r->b_hi = DY4_FMULF(Kp, e_hi);
    }
}

// ---- 3. the pick ----------------------------------------------------------------------------------------------------
// Which grid point is trigArg_k = RN_f(RN_d(w*T_k) + phase_k)?  Returns 1 and *up (0: lo, 1: hi) when that is certain:
// phase is farther than m from the threshold t between the two and closer than u - m (their far ends), i.e.
// | |phase - t| - u/2 | < u/2 - m.  0: the serial loop has to evaluate the step directly.  A NaN row fails the comparison.
DY4_HD int dy4_tab_pick(float phase, float t, float hu, float hm, int* up)
{
    const float w = DY4_FADDF(phase, -t);
    const float v = DY4_FADDF(fabsf(w), -hu);
    *up = phase > t;
    return fabsf(v) < hm;
}
