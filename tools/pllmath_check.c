/* pllmath_check.c — validates csrc/dy4_pllmath.h against glibc's double libm on the host.
 * The header is a fixed sequence of IEEE-754 basic operations, so what passes here is what the
 * device computes.  Build: gcc -O2 -ffp-contract=off -mfma -o pllmath_check tools/pllmath_check.c -lm
 *   ./pllmath_check sincos N     random float arguments in [0, 8.3e6): count float-narrowed mismatches vs sin/cos
 *   ./pllmath_check pll N_STREAMS N_SAMPLES   run the PLL recurrence both ways on noisy pilots, report mismatches
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "../3dy4-real-time-software-defined-radio-_b200/csrc/dy4_pllmath.h"

static uint64_t rng_state = 88172645463325252ull;
static inline uint64_t rnd(void) { rng_state ^= rng_state << 13; rng_state ^= rng_state >> 7; rng_state ^= rng_state << 17; return rng_state; }
static inline double urand(void) { return (rnd() >> 11) * (1.0 / 9007199254740992.0); }

typedef struct { float fbI, fbQ, integ, phase, trigOffset; } regs_t;

/* reference arithmetic (oracle/dy4_oracle.c: dy4o_pll), one step; returns nco value */
static float step_ref(float x, regs_t* s, double w, float Kp, float Ki, float ncoScale, float phaseAdjust)
{
    float eI = (x == 0 ? 1 : x) * s->fbI;
    float eQ = x * (-1 * s->fbQ);
    float eD = (float)atan2((double)eQ, (double)eI);
    float t0 = Ki * eD; s->integ = s->integ + t0;
    float t1 = Kp * eD; float t2 = t1 + s->integ; s->phase = s->phase + t2;
    s->trigOffset = s->trigOffset + 1.0f;
    float trigArg = (float)(w * (double)s->trigOffset + (double)s->phase);
    s->fbI = (float)cos((double)trigArg); s->fbQ = (float)sin((double)trigArg);
    float narg = trigArg * ncoScale + phaseAdjust;
    return (float)cos((double)narg);
}

int main(int argc, char** argv)
{
    if (getenv("SEED")) rng_state ^= 0x9E3779B97F4A7C15ull * (uint64_t)atoll(getenv("SEED"));
    if (argc >= 3 && !strcmp(argv[1], "sincos")) {
        long n = atol(argv[2]), bad_s = 0, bad_c = 0, d1 = 0;
        double maxulp = 0;
        for (long i = 0; i < n; i++) {
            float xf = (float)(urand() * (i & 1 ? 8.3e6 : 2000.0));
            if (i % 7 == 0) xf = (float)(urand() * 3.0);
            dy4_nco_t o; dy4_sincos_nco((double)xf, 0, &o);
            double rs = sin((double)xf), rc = cos((double)xf);
            if ((float)o.s != (float)rs) bad_s++;
            if ((float)o.c != (float)rc) bad_c++;
            if (o.s != rs || o.c != rc) d1++;
            double us = fabs(o.s - rs) / (nextafter(fabs(rs), INFINITY) - fabs(rs));
            double uc = fabs(o.c - rc) / (nextafter(fabs(rc), INFINITY) - fabs(rc));
            if (us > maxulp) maxulp = us; if (uc > maxulp) maxulp = uc;
        }
        printf("sincos: %ld args, float mismatches sin %ld cos %ld, double-level differences %ld (%.2f%%), max diff %.2f ulp\n",
               n, bad_s, bad_c, d1, 100.0 * d1 / n, maxulp);
        return 0;
    }
    if (argc >= 4 && !strcmp(argv[1], "pll")) {
        int ns = atoi(argv[2]); long n = atol(argv[3]);
        float Kp = 0.01f * 2.666f, Ki = 0.01f * 0.01f * 3.555f;
        float ratio = 19e3f / 240e3f; double w = 2 * 3.14159265358979323846 * (double)ratio;
        long total_mis = 0, streams_bad = 0, det_mis = 0, sc_mis = 0, grid_mis = 0;
        for (int s = 0; s < ns; s++) {
            regs_t a = {1, 0, 0, 0, 0}, b = {1, 0, 0, 0, 0};
            dy4_nco_t o; float xn = 0; int have_o = 0;
            double ph0 = urand() * 6.28, amp = 0.02 + 0.1 * urand(), df = 19e3 * (1 + 2e-5 * (urand() - 0.5));
            long first_bad = -1;
            xn = (float)(amp * sin(ph0) + 0.003 * (urand() - 0.5));
            for (long k = 0; k < n; k++) {
                float x = xn;
                xn = (float)(amp * sin(2 * 3.14159265358979323846 * df / 240e3 * (k + 1) + ph0) + 0.003 * (urand() - 0.5));
                step_ref(x, &a, w, Kp, Ki, 2.0f, 0.0f);
                /* fast path, same float ops around the custom double math */
                float eI = (x == 0 ? 1 : x) * b.fbI;
                float eQ = x * (-1 * b.fbQ);
                float eD;
                if (have_o && fabsf(x) > 1e-20f && fabsf(x) < 1e20f) {
                    double inv_x = dy4_recip(x);
                    eD = (float)dy4_detector_atan2((double)eQ, (double)eI, &o, inv_x);
                    float eDr = (float)atan2((double)eQ, (double)eI);
                    if (eD != eDr) det_mis++;
                } else eD = (float)atan2((double)eQ, (double)eI);
                float t0 = Ki * eD; b.integ = b.integ + t0;
                float t1 = Kp * eD; float t2 = t1 + b.integ; b.phase = b.phase + t2;
                b.trigOffset = b.trigOffset + 1.0f;
                double argd = w * (double)b.trigOffset + (double)b.phase;
                float trigArg = (float)argd;
                /* the magic-number rounding the device uses inside a binade must agree with the float conversion */
                if (argd >= 1.0) { int ex; frexp(argd, &ex); double lim = ldexp(1.0, ex - 1); double mg = 1.5 * lim * 536870912.0;
                    if (dy4_round_to_float_grid(argd, mg) != (double)trigArg) grid_mis++; }
                dy4_sincos_nco((double)trigArg, xn < 0, &o); have_o = 1;
                b.fbI = (float)o.c; b.fbQ = (float)o.s;
                if (b.fbI != (float)cos((double)trigArg) || b.fbQ != (float)sin((double)trigArg)) sc_mis++;
                if (first_bad < 0 && (a.fbI != b.fbI || a.fbQ != b.fbQ || a.phase != b.phase || a.integ != b.integ)) first_bad = k;
            }
            if (first_bad >= 0) { streams_bad++; total_mis++; printf("  stream %d diverged at sample %ld\n", s, first_bad); }
        }
        printf("pll: %d streams x %ld samples: %ld diverged; per-call float mismatches: detector %ld, sincos %ld, grid-rounding %ld (of %ld calls each)\n",
               ns, n, streams_bad, det_mis, sc_mis, grid_mis, (long)ns * n);
        return 0;
    }
    fprintf(stderr, "usage: %s sincos N | pll STREAMS SAMPLES\n", argv[0]);
    return 2;
}
