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

// Numeric constants live in one table: in __constant__ memory on the device so that the double-precision
// instructions take them straight from the constant bank (no per-use materialisation in registers).
#define DY4_KTAB_INIT { \
    0x1.45f306dc9c883p-1,      /* 0  2/pi                 */ \
    0x1.8p52,                  /* 1  rounding shifter     */ \
    0x1.921fb54800000p+0,      /* 2  P1                   */ \
    -0x1.de973dc800000p-31,    /* 3  P2                   */ \
    -0x1.9d9cceba3f91fp-62,    /* 4  P3                   */ \
    -1.66666666666666324348e-01, 8.33333333332248946124e-03, -1.98412698298579493134e-04,   /* 5..7   S1 S2 S3 */ \
    2.75573137070700676789e-06, -2.50507602534068634195e-08, 1.58969099521155010221e-10,    /* 8..10  S4 S5 S6 */ \
    4.16666666666666019037e-02, -1.38888888888741095749e-03, 2.48015872894767294178e-05,    /* 11..13 C1 C2 C3 */ \
    -2.75573143513906633035e-07, 2.08757232129817482790e-09, -1.13596475577881948265e-11,   /* 14..16 C4 C5 C6 */ \
    0x1.921fb54442d18p+0, 0x1.1a62633145c07p-54 /* 17,18 pi/2 hi, lo */ }
#if defined(__CUDACC__)
static __constant__ double dy4_ktab_dev[19] = DY4_KTAB_INIT;
#endif
static const double dy4_ktab_host[19] = DY4_KTAB_INIT;
#if defined(__CUDA_ARCH__)
#define DY4_K(i) dy4_ktab_dev[i]
#else
#define DY4_K(i) dy4_ktab_host[i]
#endif

// bit-level helpers (identical results on host and device)
#if defined(__CUDA_ARCH__)
DY4_HD int dy4_lo32(double v) { return __double2loint(v); }
DY4_HD double dy4_xor_sign(double v, int neg) { return __hiloint2double(__double2hiint(v) ^ (neg << 31), __double2loint(v)); }
// 1/x correctly rounded, for NORMAL x well inside the float range (callers check 1e-20 < |x| < 1e20):
// MUFU.RCP followed by one FMA Newton step is the fast path of rcp.rn — no range branches.
DY4_HD double dy4_rcp_seed(float x)
{
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    const float e = -fmaf(r, x, -1.0f);
    return (double)fmaf(r, e, r);
}
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
    double c, s;            // cos, sin of the (float) phase argument, < 1 ulp
    double base_hi, base_lo;// reference angle of the NEXT detector call as a double-double:  B - rho,
                            // rho = argument - n*pi/2, B = 0, -pi/2, +-pi, +pi/2 by quadrant and input sign
} dy4_nco_t;

// sin/cos of x (a float value widened to double, |x| < 8.4e6) and, from the same reduction, the reference
// angle the phase detector will need for the next input sample (whose sign is already known).
//
// Latency matters more than operation count here (one thread per stream, serial recurrence), so:
//  * the polynomials start from the first-stage remainder hi0 = (x - n*P1) - n*P2, two operations after
//    the quadrant is known; the rest of the reduction (exact rounding error of that subtraction and the
//    n*P3 term, together y, |y| < 3e-12) runs beside them and enters as the first-order terms
//    y*cos(hi0) / y*sin(hi0) inside the final sums, so each result still has ONE significant rounding;
//  * Estrin evaluation, depth 5 after z = hi0^2.
DY4_HD void dy4_sincos_nco_v(double x, int next_input_negative, dy4_nco_t* o, const int select_b)
{
    // n = nearest integer to x*2/pi (magic-number rounding; |x*2/pi| < 2^23), q = n mod 4 from the mantissa bits
    const double big = DY4_K(1);
    const double shifted = DY4_ADD(DY4_MUL(x, DY4_K(0)), big);
    const double fn = DY4_SUB(shifted, big);
    const int q = dy4_lo32(shifted) & 3;
    // n*P1 and n*P2 are exact products (30-bit constants, n < 2^23); x - n*P1 is exact
    const double r1 = fma(-fn, DY4_K(2), x);
    const double t2 = DY4_MUL(fn, DY4_K(3));
    const double hi0 = DY4_SUB(r1, t2);
    // --- beside the polynomials: tail y, and the next detector's reference angle -----------------------
    const double bb = DY4_SUB(hi0, r1);                                      // TwoDiff: exact error of r1 - t2
    const double e0 = DY4_SUB(DY4_SUB(r1, DY4_SUB(hi0, bb)), DY4_ADD(t2, bb));
    const double y = fma(-fn, DY4_K(4), e0);                                 // rho = hi0 + y
    {
        const int Q = (q + (next_input_negative ? 2 : 0)) & 3;
        const int rho_sign = (DY4_ADD(hi0, y) < 0.0) ? -1 : 1;               // rho == 0 only for a zero argument: B = +pi
        double b_hi, b_lo;
        if (select_b) {
            // B = 0, -pi/2, +-pi (sign of rho), +pi/2 for Q = 0,1,2,3, picked with selects (no conversions, no table)
            const double k_hi = (Q & 1) ? DY4_K(17) : DY4_ADD(DY4_K(17), DY4_K(17));   // pi/2 or pi (doubling is exact)
            const double k_lo = (Q & 1) ? DY4_K(18) : DY4_ADD(DY4_K(18), DY4_K(18));
            const int neg = (Q & 1) ? (Q == 1) : (rho_sign < 0);
            b_hi = (Q == 0) ? 0.0 : dy4_xor_sign(k_hi, neg);
            b_lo = (Q == 0) ? 0.0 : dy4_xor_sign(k_lo, neg);
        } else {
            // B = m*pi/2 with m = 0, -1, +-2, +1: products by 0, +-1, +-2 are exact
            const int m = (Q & 1) ? (Q - 2) : Q * rho_sign;
            const double mB = (double)m;
            b_hi = DY4_MUL(mB, DY4_K(17)); b_lo = DY4_MUL(mB, DY4_K(18));
        }
        const double s_hi = DY4_SUB(b_hi, hi0);                              // Fast2Sum: |b_hi| >= pi/2 > |hi0|, or b_hi == 0
        const double s_err = DY4_SUB(DY4_SUB(b_hi, s_hi), hi0);
        o->base_hi = s_hi;
        o->base_lo = DY4_ADD(s_err, DY4_SUB(b_lo, y));
    }
    // --- minimax kernels on |hi0| <= pi/4 (coefficients of fdlibm's __kernel_sin / __kernel_cos) ------------
    const double S1 = DY4_K(5), S2 = DY4_K(6), S3 = DY4_K(7), S4 = DY4_K(8), S5 = DY4_K(9), S6 = DY4_K(10);
    const double C1 = DY4_K(11), C2 = DY4_K(12), C3 = DY4_K(13), C4 = DY4_K(14), C5 = DY4_K(15), C6 = DY4_K(16);
    const double z = DY4_MUL(hi0, hi0);
    const double w = DY4_MUL(z, z);
    const double v = DY4_MUL(z, hi0);
    // first-order tail terms: y*cos(hi0), y*sin(hi0) with short series (relative error < 4e-6, times |y| < 3e-12)
    const double cm = fma(w, fma(z, -1.0 / 720, 1.0 / 24), fma(z, -0.5, 1.0));
    const double sm = DY4_MUL(hi0, fma(w, fma(z, -1.0 / 5040, 1.0 / 120), fma(z, -1.0 / 6, 1.0)));
    const double yc = DY4_MUL(y, cm), ys = DY4_MUL(y, sm);
    // sin(hi0+y) = hi0 + [ v*(S1 + z*S2 + z^2*S3 + ...) + y*cos ]
    const double sA = fma(z, S5, S4), sB = fma(z, S3, S2);
    const double sC = fma(w, S6, sA);
    const double sE = fma(z, sB, S1);
    const double zw = DY4_MUL(z, w);
    const double sD = fma(zw, sC, sE);                                       // S1 + z S2 + w S3 + zw S4 + w^2 S5 + w zw S6
    const double sn = DY4_ADD(hi0, fma(v, sD, yc));
    // cos(hi0+y) = (1 - z/2) + [ rounding error of (1 - z/2) + w*(C1 + z C2 + ...) - y*sin ]
    const double cF = fma(z, C2, C1), cG = fma(z, C4, C3), cH = fma(z, C6, C5);
    const double cI = fma(w, cH, cG);
    const double w2 = DY4_MUL(w, w);
    const double cJ = fma(w, cF, -ys);
    const double ct = fma(w2, cI, cJ);
    const double hz = DY4_MUL(0.5, z);
    const double one_m = DY4_SUB(1.0, hz);
    const double comp = DY4_SUB(DY4_SUB(1.0, one_m), hz);                    // exact error of 1 - hz
    const double cs = DY4_ADD(one_m, DY4_ADD(comp, ct));
    // quadrant: q=0 (c,s)=(cs,sn); 1: (-sn,cs); 2: (-cs,-sn); 3: (sn,-cs) — selects and sign flips, no branches
    const int swap = q & 1;
    const double cq = swap ? sn : cs;
    const double sq = swap ? cs : sn;
    o->c = dy4_xor_sign(cq, ((q + 1) >> 1) & 1);
    o->s = dy4_xor_sign(sq, (q >> 1) & 1);
}

DY4_HD void dy4_sincos_nco(double x, int next_input_negative, dy4_nco_t* o) { dy4_sincos_nco_v(x, next_input_negative, o, 0); }

// atan2(eQ, eI) as the reference calls it (filter.cpp:200) where eI = fl(x*fbI), eQ = fl(x*(-fbQ)),
// (fbI, fbQ) = float(o->c, o->s), x = the (non-zero, finite, normal) input sample whose sign was given to
// dy4_sincos_nco, inv_x = 1/x in double (~1e-14).  Returns the double the libm call would return, to ~1 ulp:
// the angle of (eI,eQ) is the reference angle plus the small rotation t = cross/dot against the unit vector (c,s).
DY4_HD double dy4_detector_atan2(double eQ, double eI, const dy4_nco_t* o, double inv_x)
{
    // cross = eQ*c + eI*s with the rounding error of the first product compensated; dot = eI*c - eQ*s ~ x
    const double p = DY4_MUL(eI, o->s);
    const double pe = fma(eI, o->s, -p);
    const double cross = DY4_ADD(fma(eQ, o->c, p), pe);
    const double dot = fma(eI, o->c, -DY4_MUL(eQ, o->s));
    // 1/dot ~ inv_x*(2 - dot*inv_x): dot*inv_x = 1 + O(1e-7), so the relative error is O(1e-14)
    const double g = DY4_MUL(dot, inv_x);
    const double t = DY4_MUL(DY4_MUL(cross, inv_x), DY4_SUB(2.0, g));
    return DY4_ADD(o->base_hi, DY4_ADD(o->base_lo, t));
}

// RN32 of a positive double `a` known to lie in [lim, 2*lim), lim a power of two, without leaving the double
// pipe: adding and subtracting magic = 1.5 * lim * 2^29 rounds to the float grid of that binade, ties to even.
DY4_HD double dy4_round_to_float_grid(double a, double magic) { return DY4_SUB(DY4_ADD(a, magic), magic); }
