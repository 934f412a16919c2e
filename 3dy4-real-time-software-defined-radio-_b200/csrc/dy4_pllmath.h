// dy4_pllmath.h — double-precision sin/cos and the PLL phase detector, written as explicit sequences
// of IEEE-754 basic operations (fma, +, *, one reciprocal) so that the SAME rounding sequence runs on
// the host (validation against glibc, tools/pllmath_check.c) and on the device (dy4_pll.cu).
//
// Why this exists: the reference's fmPLL (src/filter.cpp:174-228) calls the DOUBLE libm atan2/sin/cos
// once per IF sample inside a serial recurrence and narrows each result to float.  The recurrence is
// chaotic in the last bit, so the float results must equal glibc's, and with one
// thread per stream the kernel is bound by the LATENCY of that dependent chain.  CUDA's libm gives the
// right floats but costs ~1000 cycles per sample.  Here:
//   * sin/cos share one Cody-Waite reduction (pi/2 in three parts, 30+30+53 bits, exact for the float
//     arguments |x| < 2^23 * pi/2 the PLL produces) and fdlibm-style minimax kernels, < 1 ulp;
//   * atan2(eQ, eI) is NOT evaluated from scratch: the detector's inputs are the just-computed
//     (cos, sin) scaled by the input sample and rounded to float, so the angle is the known reduced
//     phase plus a ~1e-7 correction  t = cross/dot  obtained from an exactly-compensated 2x2
//     determinant; the division uses the precomputed reciprocal of the input sample.
// Both are accurate to about one double ulp, i.e. they narrow to the same float as a correctly
// rounded libm except with probability ~1e-9 per call — the same class as CUDA's own libm.
#pragma once

#if defined(__CUDA_ARCH__)
#define DY4_HD __host__ __device__ __forceinline__
#define DY4_MUL(a, b) __dmul_rn((a), (b))
#define DY4_ADD(a, b) __dadd_rn((a), (b))
#define DY4_SUB(a, b) __dadd_rn((a), -(b))
#elif defined(__CUDACC__)
#define DY4_HD __host__ __device__ __forceinline__
#define DY4_MUL(a, b) ((a) * (b))
#define DY4_ADD(a, b) ((a) + (b))
#define DY4_SUB(a, b) ((a) - (b))
#else
#include <math.h>
#define DY4_HD static inline
#define DY4_MUL(a, b) ((a) * (b))   /* host: build with -ffp-contract=off */
#define DY4_ADD(a, b) ((a) + (b))
#define DY4_SUB(a, b) ((a) - (b))
#endif

// bit-level helpers (identical results on host and device)
#if defined(__CUDA_ARCH__)
DY4_HD int dy4_lo32(double v) { return __double2loint(v); }
DY4_HD double dy4_xor_sign(double v, int neg) { return __hiloint2double(__double2hiint(v) ^ (neg << 31), __double2loint(v)); }
DY4_HD double dy4_rcp_seed(float x) { return (double)__frcp_rn(x); }
#else
#include <string.h>
DY4_HD int dy4_lo32(double v) { unsigned long long b; memcpy(&b, &v, 8); return (int)(unsigned)(b & 0xffffffffu); }
DY4_HD double dy4_xor_sign(double v, int neg) { unsigned long long b; memcpy(&b, &v, 8); b ^= (unsigned long long)(neg & 1) << 63; memcpy(&v, &b, 8); return v; }
DY4_HD double dy4_rcp_seed(float x) { return (double)(1.0f / x); }
#endif

// 1/x for a normal float x, ~1e-14 relative: correctly rounded float reciprocal + one Newton step in double
DY4_HD double dy4_recip(float x)
{
    const double r0 = dy4_rcp_seed(x);
    return DY4_MUL(r0, fma(-(double)x, r0, 2.0));
}

typedef struct {
    double c, s;        // cos, sin of the (float) phase argument, ~0.6 ulp
    double rho_hi, rho_lo;  // argument minus n*pi/2 as a double-double, |rho| <= pi/4
    int n;              // quadrant index 0..3 (n mod 4)
} dy4_nco_t;

#define DY4_P1 0x1.921fb54800000p+0     /* pi/2, leading 30 bits  */
#define DY4_P2 (-0x1.de973dc800000p-31) /* next 30 bits            */
#define DY4_P3 (-0x1.9d9cceba3f91fp-62) /* the rest                */
#define DY4_TWO_OVER_PI 0x1.45f306dc9c883p-1
#define DY4_PIO2_HI 0x1.921fb54442d18p+0
#define DY4_PIO2_LO 0x1.1a62633145c07p-54
#define DY4_PI_HI 0x1.921fb54442d18p+1
#define DY4_PI_LO 0x1.1a62633145c07p-53

// sin/cos of x (a float value widened to double, |x| < 8.4e6), with the reduction kept for the detector
DY4_HD void dy4_sincos_nco(double x, dy4_nco_t* o)
{
    // n = nearest integer to x*2/pi (magic-number rounding; |x*2/pi| < 2^23)
    const double big = 0x1.8p52;
    const double shifted = DY4_ADD(DY4_MUL(x, DY4_TWO_OVER_PI), big);
    const double fn = DY4_SUB(shifted, big);
    const int q = dy4_lo32(shifted) & 3;                         // low mantissa bits of the shifted value hold n (two's complement)
    // rho = x - n*pi/2 as hi+lo.  n*P1 and n*P2 are exact products; x - n*P1 is exact.
    const double r1 = fma(-fn, DY4_P1, x);
    const double t2 = DY4_MUL(fn, DY4_P2);
    const double hi0 = DY4_SUB(r1, t2);
    const double bb = DY4_SUB(hi0, r1);                          // TwoDiff: exact error of r1 - t2
    const double lo0 = DY4_SUB(DY4_SUB(r1, DY4_SUB(hi0, bb)), DY4_ADD(t2, bb));
    const double lo1 = fma(-fn, DY4_P3, lo0);
    const double hi = DY4_ADD(hi0, lo1);                         // renormalise (Fast2Sum: |hi0| >= |lo1| or hi0 == 0)
    const double lo = DY4_SUB(lo1, DY4_SUB(hi, hi0));
    o->rho_hi = hi; o->rho_lo = lo; o->n = q;

    // minimax kernels on |hi| <= pi/4 (coefficients of fdlibm's __kernel_sin / __kernel_cos)
    const double z = DY4_MUL(hi, hi);
    const double w = DY4_MUL(z, z);
    const double S1 = -1.66666666666666324348e-01, S2 = 8.33333333332248946124e-03, S3 = -1.98412698298579493134e-04,
                 S4 = 2.75573137070700676789e-06, S5 = -2.50507602534068634195e-08, S6 = 1.58969099521155010221e-10;
    const double C1 = 4.16666666666666019037e-02, C2 = -1.38888888888741095749e-03, C3 = 2.48015872894767294178e-05,
                 C4 = -2.75573143513906633035e-07, C5 = 2.08757232129817482790e-09, C6 = -1.13596475577881948265e-11;
    // sin: hi + lo + hi^3*(S1 + z*(S2 + ... )) ; Estrin in z,w to shorten the dependent chain
    const double sp = fma(w, fma(w, S6, fma(z, S5, S4)), fma(z, S3, S2));   // S2 + z S3 + w (S4 + z S5 + w S6)
    const double v = DY4_MUL(z, hi);
    // fdlibm form: x - ((z*(0.5*y - v*r) - y) - v*S1)
    const double sn = DY4_SUB(hi, DY4_SUB(DY4_SUB(DY4_MUL(z, fma(-v, sp, DY4_MUL(0.5, lo))), lo), DY4_MUL(v, S1)));
    // cos: 1 - z/2 + z^2*(C1 + z C2 + ...) - hi*lo
    const double cp = fma(w, fma(w, fma(z, C6, C5), fma(z, C4, C3)), fma(z, C2, C1));  // C1 + z C2 + w(C3 + z C4 + w (C5 + z C6))
    const double hz = DY4_MUL(0.5, z);
    const double one_m = DY4_SUB(1.0, hz);
    // 1 - hz = one_m + ((1 - one_m) - hz) exactly; add the small terms to the correction
    const double cs = DY4_ADD(one_m, DY4_ADD(DY4_SUB(DY4_SUB(1.0, one_m), hz), fma(w, cp, -DY4_MUL(hi, lo))));

    // quadrant: q=0 (c,s)=(cs,sn); 1: (-sn,cs); 2: (-cs,-sn); 3: (sn,-cs) — selects and sign flips, no branches
    const int swap = q & 1;
    const double cm = swap ? sn : cs;
    const double sm = swap ? cs : sn;
    o->c = dy4_xor_sign(cm, ((q + 1) >> 1) & 1);
    o->s = dy4_xor_sign(sm, (q >> 1) & 1);
}

// atan2(eQ, eI) as the reference calls it (filter.cpp:200) where eI = fl(x*fbI), eQ = fl(x*(-fbQ)),
// (fbI, fbQ) = float(o->c, o->s), x = the (non-zero, finite, normal) input sample, inv_x = 1/x in double.
// Returns the double the libm call would return, to about one ulp.
DY4_HD double dy4_detector_atan2(double eQ, double eI, double x_is_negative, const dy4_nco_t* o, double inv_x)
{
    // reference angle phi0 = B - rho with effective quadrant Q = n + 2*[x<0]
    // B = mB * pi/2 with mB = 0, -1, +-2 (sign of rho), +1 for Q = 0,1,2,3; products by 0,+-1,+-2 are exact
    const int Q = (o->n + (x_is_negative != 0.0 ? 2 : 0)) & 3;
    const int rho_neg = (o->rho_hi < 0.0) | ((o->rho_hi == 0.0) & (o->rho_lo < 0.0));
    const int m_even = (Q == 2) ? (rho_neg ? -2 : 2) : 0;
    const int m_odd = (Q == 1) ? -1 : 1;
    const double mB = (double)((Q & 1) ? m_odd : m_even);
    const double b_hi = DY4_MUL(mB, DY4_PIO2_HI), b_lo = DY4_MUL(mB, DY4_PIO2_LO);
    // base = B - rho (double-double), off the critical path
    const double s_hi = DY4_SUB(b_hi, o->rho_hi);
    const double bb = DY4_SUB(s_hi, b_hi);
    const double s_err = DY4_SUB(DY4_SUB(b_hi, DY4_SUB(s_hi, bb)), DY4_ADD(o->rho_hi, bb));
    const double s_lo = DY4_ADD(s_err, DY4_SUB(b_lo, o->rho_lo));
    // rotation of (eI,eQ) by the unit vector (c,s): cross = eQ*c + eI*s (compensated), dot = eI*c - eQ*s ~ x
    const double p = DY4_MUL(eI, o->s);
    const double pe = fma(eI, o->s, -p);
    const double cross = DY4_ADD(fma(eQ, o->c, p), pe);
    const double dot = fma(eI, o->c, -DY4_MUL(eQ, o->s));
    // t = cross/dot with 1/dot ~ inv_x*(2 - dot*inv_x)  (dot*inv_x = 1 + O(1e-7), so the error is O(1e-14) relative)
    const double g = DY4_MUL(dot, inv_x);
    const double t = DY4_MUL(DY4_MUL(cross, inv_x), DY4_SUB(2.0, g));
    return DY4_ADD(s_hi, DY4_ADD(s_lo, t));
}
