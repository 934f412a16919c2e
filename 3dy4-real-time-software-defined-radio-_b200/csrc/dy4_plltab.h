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
//   1. k_pll_predict    time-parallel (segments with a warm-up), serial only in cheap double adds: predicted trigArg
//   2. k_pll_table_ops  fully parallel: for every sample the EXACT errorD of the next step for the two neighbouring floats the
//                       prediction lies between (the dy4_pllmath.h sincos + detector, unchanged arithmetic), as a chain-ready row
//   3. k_pll_sel        the serial loop: per sample ONE float compare of phaseEst with a precomputed threshold picks the grid
//                       point the true trigArg falls on, and two speculative loop-filter updates (float adds) are selected
//                       from.  Whether each pick was CERTAIN under its rounding-error budget is decided after every 32 steps,
//                       one lane per step; a step that is not (outside the two candidates, within the guard band of the
//                       threshold, binade edges, start-up) is evaluated directly with dy4_pllmath.h, so the result is the
//                       reference's bit for bit by construction.
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

#if defined(__CUDA_ARCH__)
DY4_HD unsigned long long dy4_d2u_bits(double v) { return (unsigned long long)__double_as_longlong(v); }
DY4_HD double dy4_u2d_bits(unsigned long long v) { return __longlong_as_double((long long)v); }
#else
DY4_HD unsigned long long dy4_d2u_bits(double v) { unsigned long long b; memcpy(&b, &v, 8); return b; }
DY4_HD double dy4_u2d_bits(unsigned long long v) { double f; memcpy(&f, &v, 8); return f; }
#endif

// A double that is about to be narrowed to float lies within 2 double-ulps of a float rounding boundary (a tie of round-to-
// nearest: the 29 bits that are dropped read 1000...0).  dy4_pllmath.h is not glibc: its results may differ from glibc's in
// the last double ulp, and only at such a value can that change the float — so this is the census of evaluations on which
// bit-identity with the reference rests on more than the arithmetic contract (dy4_pipeline_pll_risk).
DY4_HD int dy4_near_float_tie(double v)
{
    const int d = (int)(unsigned)(dy4_d2u_bits(v) & 0x1fffffffull) - 0x10000000;
    return d >= -2 && d <= 2;
}

// ---- 2. table row --------------------------------------------------------------------------------------------------
// errorD of the step that follows trigArg = th (a float value) when its input sample is x: the arithmetic of
// pll_step_fast / detector_libm in dy4_pll.cu, i.e. of filter.cpp:192-200,216-217.
DY4_HD float dy4_next_errorD_r(double th, float x, int* risk)
{
    dy4_nco_t o;
    dy4_sincos_nco_v(th, x < 0.0f, &o, 0);
    const float fbI = DY4_D2F(o.c), fbQ = DY4_D2F(o.s);                                   // :216-217
    if (risk) *risk += dy4_near_float_tie(o.c) + dy4_near_float_tie(o.s);
    if (dy4_fast_ok(x)) {
        const float eI = DY4_FMULF(x, fbI);                                               // :192 (x != 0)
        const float eQ = DY4_FMULF(x, -fbQ);                                              // :193
        const double a = dy4_detector_atan2((double)eQ, (double)eI, &o, dy4_recip(x));
        if (risk) *risk += dy4_near_float_tie(a);
        return DY4_D2F(a);                                                                // :200
    }
    const float eI = DY4_FMULF((x == 0.0f ? 1.0f : x), fbI);
    const float eQ = DY4_FMULF(x, -fbQ);
    return DY4_D2F(atan2((double)eQ, (double)eI));
}
DY4_HD float dy4_next_errorD(double th, float x) { return dy4_next_errorD_r(th, x, (int*)0); }

#if defined(__CUDA_ARCH__)
DY4_HD int dy4_f2i_bits(float v) { return __float_as_int(v); }
DY4_HD float dy4_i2f_bits(int v) { return __int_as_float(v); }
#else
DY4_HD int dy4_f2i_bits(float v) { int b; memcpy(&b, &v, 4); return b; }
DY4_HD float dy4_i2f_bits(int v) { float f; memcpy(&f, &v, 4); return f; }
#endif

// ---- 2b. the row: two candidates and the threshold between them ------------------------------------------------------
// The prediction th_hat lies between two neighbouring floats lo < hi = lo + u, and so - almost always - does the true trigArg.
//   t    double: threshold between lo and hi in the phaseEst domain, t = (lo - RN_d(w*T)) + u/2: trigArg = RN_f(RN_d(w*T) + phase)
//        is lo for phase in (t - u, t) and hi for phase in (t, t + u).  The low 9 mantissa bits carry  bit 0: the candidate nearest
//        to the prediction is hi;  bits 1..8: biased float exponent of u.  NaN: row unusable.
//   e_p, e_o   errorD of the next step if trigArg is the predicted / the other candidate (filter.cpp:200)
// k_pll_table_ops turns this into the 32-byte row the loop reads (products with Ki, Kp; float threshold; the cells' centres and
// certified half-width, dy4_spec_fast_row).  dy4_spec_check is the certificate in double (budget 2^-26 u + 2^-40 |t|): the host
// tests hold the loop's float certificate to it.
// th_hat: predicted trigArg of this sample; wT = RN_d(w*T_k); x_next: input of step k+1 (has_next == 0 for the last sample of a
// launch: errorD unused); force_invalid: rows the serial loop must evaluate directly whatever the prediction says.
typedef struct { double t; float e_p, e_o; } dy4_row16_t;


// *lo_out / *u_out (host checks only, may be NULL): the lower candidate and the grid spacing.
DY4_HD void dy4_tab_make_row16_r(double th_hat, double wT, float x_next, int has_next, int force_invalid, dy4_row16_t* r, float* lo_out, float* u_out, int* risk)
{
    const float c = DY4_D2F(th_hat);
    const int bits = dy4_f2i_bits(c);
    const int expo = (bits >> 23) & 0xff, mant = bits & 0x7fffff;
    const int ok = !force_invalid && bits > 0 && expo >= 110 && expo <= 190 && mant >= 2 && mant <= 0x7ffffd;
    const float u = dy4_i2f_bits((ok ? expo - 23 : 127) << 23);
    const int pred_hi = !(th_hat >= (double)c);                    // c is the candidate nearest to the prediction
    const float lo = pred_hi ? DY4_FADDF(c, -u) : c, hi = DY4_FADDF(lo, u);
    const double t = DY4_ADD(DY4_SUB((double)lo, wT), DY4_MUL(0.5, (double)u));
    const unsigned long long tb = (dy4_d2u_bits(t) & ~0x1ffull) | (unsigned long long)pred_hi | ((unsigned long long)(expo - 23) << 1);
    r->t = ok ? dy4_u2d_bits(tb) : dy4_u2d_bits(0x7ff8000000000000ull);
    r->e_p = r->e_o = 0.0f;
    if (lo_out) *lo_out = lo;
    if (u_out) *u_out = u;
    if (ok && has_next) {
        r->e_p = dy4_next_errorD_r((double)(pred_hi ? hi : lo), x_next, risk);
        r->e_o = dy4_next_errorD_r((double)(pred_hi ? lo : hi), x_next, risk);
    }
}
DY4_HD void dy4_tab_make_row16(double th_hat, double wT, float x_next, int has_next, int force_invalid, dy4_row16_t* r, float* lo_out, float* u_out)
{
    dy4_tab_make_row16_r(th_hat, wT, x_next, has_next, force_invalid, r, lo_out, u_out, (int*)0);
}

// Is trigArg = RN_f(RN_d(w*T) + phase) CERTAINLY the predicted candidate of this row (other == 0) / the other one (other == 1)?
// With d = phase - t:  hi iff d in (0, u), lo iff d in (-u, 0); certain when d keeps a distance m from the ends, where
// m = 2^-26 u + 2^-40 |t| bounds everything that is not exact on the way: the 9 borrowed mantissa bits of t (2^-43 |t|),
// the two double roundings in t (2^-52 |t|), this subtraction (2^-53 |d|) and the reference's own RN_d (2^-29 u).
// A NaN row fails both comparisons.
DY4_HD int dy4_spec_check(float phase, double t, int other)
{
    const unsigned long long b = dy4_d2u_bits(t);
    const int side_hi = (int)(b & 1) ^ other;
    const double u = dy4_u2d_bits((((b >> 1) & 0xff) + 896ull) << 52);      // 2^(e - 127) as a double: e - 127 + 1023
    const double d = DY4_SUB((double)phase, t);
    const double dd = side_hi ? d : -d;
    const double m = DY4_ADD(DY4_MUL(1.490116119384765625e-08, u), DY4_MUL(9.094947017729282379e-13, fabs(t)));
    return dd > m && dd < DY4_SUB(u, m);
}

// The loop's certificate, in float, one lane per step: the cell's centre tc = t -+ u/2 rounded to float and
// a half-width hm = u/2 - m_f with m_f = 2^-22 (|t| + u), which covers the float rounding of t (2^-24 |t|), of tc
// (2^-24 (|t| + u/2)) and of the loop's own subtraction (2^-24 u) on top of dy4_spec_check's budget, twice over.
// |phase - tc| < hm  =>  dy4_spec_check holds.  Where it fails the loop evaluates the step directly.
// q = (tc of the predicted candidate, hm, tc of the other candidate); a NaN row gives NaN everywhere (never certain).
DY4_HD void dy4_spec_fast_row(double t, float* tc_p, float* hm, float* tc_o)
{
    const unsigned long long b = dy4_d2u_bits(t);
    const float u = dy4_i2f_bits((int)((b >> 1) & 0xff) << 23);
    const float tf = DY4_D2F(t), hu = DY4_FMULF(0.5f, u);
    const float up = DY4_FADDF(tf, hu), dn = DY4_FADDF(tf, -hu);
    *tc_p = (b & 1) ? up : dn;
    *tc_o = (b & 1) ? dn : up;
    *hm = DY4_FADDF(hu, -DY4_FMULF(2.384185791015625e-07f, DY4_FADDF(fabsf(tf), u)));
}
DY4_HD int dy4_spec_fast_check(float phase, float tc, float hm) { return fabsf(DY4_FADDF(phase, -tc)) < hm; }
