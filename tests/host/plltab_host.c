/* plltab_host.c — host build of csrc/dy4_plltab.h: the three parts of the table-driven PLL (predict, table rows, serial
 * loop with its certificate) run exactly as the kernels of dy4_pll.cu run them, so that the construction can be checked against the
 * reference recurrence on the CPU (tests/test_host_logic.py).  Test infrastructure only.
 * Build: gcc -O2 -ffp-contract=off -mfma -shared -fPIC -o libplltab_host.so plltab_host.c -lm */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include "../../3dy4-real-time-software-defined-radio-_b200/csrc/dy4_plltab.h"

#define SEG 512
#define WARM 1024

/* the reference recurrence (filter.cpp:174-228) with glibc, same outputs */
void pllref_launch(const float* x, int n, float* state, double w, float Kp, float Ki, float* theta_out)
{
    float fbI = state[0], fbQ = state[1], integ = state[2], phase = state[3], T = state[4];
    for (int k = 0; k < n; k++) {
        const float eI = (x[k] == 0 ? 1 : x[k]) * fbI, eQ = x[k] * (-1 * fbQ);
        const float eD = (float)atan2((double)eQ, (double)eI);
        const float t0 = Ki * eD; integ = integ + t0;
        const float t1 = Kp * eD; const float t2 = t1 + integ; phase = phase + t2;
        T = T + 1.0f;
        const float trigArg = (float)(w * (double)T + (double)phase);
        fbI = (float)cos((double)trigArg); fbQ = (float)sin((double)trigArg);
        theta_out[k] = trigArg;
    }
    state[0] = fbI; state[1] = fbQ; state[2] = integ; state[3] = phase; state[4] = T;
}

static unsigned long long fz_state;
static unsigned long long fz_next(void) { fz_state ^= fz_state << 13; fz_state ^= fz_state >> 7; fz_state ^= fz_state << 17; return fz_state; }
static double fz_uni(void) { return (double)(fz_next() >> 11) * (1.0 / 9007199254740992.0); }

/* ---- k_pll_sel (csrc/dy4_pll.cu): chain-ready rows from k_pll_table_ops, groups of G steps, the pick `phaseEst > t` on the
 * chain, the float certificate of every step afterwards, an uncertain step evaluated directly.  The control flow below is the
 * kernel's, lane by lane.
 * stats[0] += steps taken on the certificate, [1] += direct evaluations, [2] += certified-but-wrong steps (must stay 0),
 * [3] += groups run. */
void pllspec_launch_carry(const float* x, int n, float* state, double w, float Kp, float Ki, float* theta_out, long* stats, double* pred, int carry, int G)
{
    float* rt = (float*)malloc(sizeof(float) * (size_t)(n + 64));            /* per row: t, hm, tc_lo, tc_hi, e_lo, e_hi */
    float* rhm = (float*)malloc(sizeof(float) * (size_t)(n + 64));
    float* rtl = (float*)malloc(sizeof(float) * (size_t)(n + 64));
    float* rth = (float*)malloc(sizeof(float) * (size_t)(n + 64));
    float* rel = (float*)malloc(sizeof(float) * (size_t)(n + 64));
    float* reh = (float*)malloc(sizeof(float) * (size_t)(n + 64));
    float* lo = (float*)malloc(sizeof(float) * (size_t)(n + 64));
    float* uu = (float*)malloc(sizeof(float) * (size_t)(n + 64));
    double* th_hat = (double*)malloc(sizeof(double) * (size_t)n);
    float* y = (float*)malloc(sizeof(float) * (size_t)n);            /* phaseEst after every sample */
    const double T0 = carry ? pred[2] : (double)state[4];
    const double g_integ = carry ? pred[0] : (double)state[2], g_phase = carry ? pred[1] : (double)state[3];
    for (int s0 = 0; s0 < n; s0 += SEG) {
        int kw = s0 - WARM; if (kw < 0) kw = 0;
        double integ = g_integ, phase = g_phase;
        double th_prev = w * dy4_pll_count(T0, kw) + phase;
        for (int k = kw; k < s0 + SEG && k < n; k++) {
            th_prev = dy4_pred_step(x[k], th_prev, DY4_MUL(w, dy4_pll_count(T0, k + 1)), (double)Kp, (double)Ki, &integ, &phase);
            if (k >= s0) th_hat[k] = DY4_MUL(w, dy4_pll_count(T0, k + 1)) + (double)(float)phase;      /* k_pll_predict stores phaseEst as a float; k_pll_table_ops adds RN_d(w*T) back */
        }
        if (s0 + SEG >= n) { pred[0] = integ; pred[1] = phase; pred[2] = dy4_pll_count(T0, n); pred[3] = T0; }
    }
    for (int k = 0; k < n + 64; k++) {                               /* k_pll_table_ops */
        if (k >= n) { rt[k] = rhm[k] = rtl[k] = rth[k] = rel[k] = reh[k] = NAN; lo[k] = uu[k] = 0; continue; }
        dy4_row16_t r;
        dy4_tab_make_row16(th_hat[k], DY4_MUL(w, dy4_pll_count(T0, k + 1)), k + 1 < n ? x[k + 1] : 0.0f, k + 1 < n, T0 + (double)k < (double)DY4_TAB_EARLY,
                           &r, &lo[k], &uu[k]);
        const int ph = (int)(dy4_d2u_bits(r.t) & 1);
        float tc_p, hm, tc_o;
        dy4_spec_fast_row(r.t, &tc_p, &hm, &tc_o);
        rel[k] = ph ? r.e_o : r.e_p; reh[k] = ph ? r.e_p : r.e_o;
        rt[k] = (float)r.t; rhm[k] = hm; rtl[k] = ph ? tc_o : tc_p; rth[k] = ph ? tc_p : tc_o;
    }
    float fbI = state[0], fbQ = state[1], integ = state[2], phase = state[3];
    {
        const float eI = DY4_FMULF((x[0] == 0.0f ? 1.0f : x[0]), fbI), eQ = DY4_FMULF(x[0], -fbQ);
        dy4_pll_filter(DY4_D2F(atan2((double)eQ, (double)eI)), Kp, Ki, &integ, &phase);
    }
    const int n_rows = n - 1;                  /* the kernel's kd (direct start-up part) is 0 here: early rows are NaN and are evaluated directly */
    int r = 0;
    float gi = integ, gp = phase;
    while (r < n_rows) {
        float ci[33], cp[33];
        int ups[33];
        float i0 = gi, p0 = gp;
        int nv = n_rows - r; if (nv > G) nv = G;
        for (int i = 0; i < G; i++) {                                 /* the chain (all G steps, as the kernel) */
            const int k = r + i;
            ci[i] = i0; cp[i] = p0;
            const float a_lo = DY4_FMULF(Ki, rel[k]), a_hi = DY4_FMULF(Ki, reh[k]), b_lo = DY4_FMULF(Kp, rel[k]), b_hi = DY4_FMULF(Kp, reh[k]);
            const float i_lo = DY4_FADDF(i0, a_lo), i_hi = DY4_FADDF(i0, a_hi);
            const float p_lo = DY4_FADDF(p0, DY4_FADDF(b_lo, i_lo)), p_hi = DY4_FADDF(p0, DY4_FADDF(b_hi, i_hi));
            const int up = p0 > rt[k];
            ups[i] = up;
            i0 = up ? i_hi : i_lo; p0 = up ? p_hi : p_lo;
        }
        ci[G] = i0; cp[G] = p0;
        stats[3]++;
        int j = nv;
        for (int lane = 0; lane < nv; lane++) {                       /* the certificate, lane i for step i */
            const int k = r + lane;
            const float myph = cp[lane];
            const float tc = myph > rt[k] ? rth[k] : rtl[k];
            if (!dy4_spec_fast_check(myph, tc, rhm[k])) { j = lane; break; }
        }
        for (int lane = 0; lane < j; lane++) {
            const int k = r + lane;
            y[k] = cp[lane];
            const float cand = ups[lane] ? lo[k] + uu[k] : lo[k];
            if (dy4_pll_trigarg(w, dy4_pll_count(T0, k + 1), cp[lane]) != cand) stats[2]++;
            stats[0]++;
        }
        if (j == nv) { r += nv; gi = ci[nv]; gp = cp[nv]; continue; }
        const int k = r + j;
        float ni = ci[j], np = cp[j];
        y[k] = np;
        dy4_pll_filter(dy4_next_errorD((double)dy4_pll_trigarg(w, dy4_pll_count(T0, k + 1), np), x[k + 1]), Kp, Ki, &ni, &np);
        stats[1]++;
        gi = ni; gp = np; r = k + 1;
    }
    integ = gi; phase = gp;
    y[n - 1] = phase;
    for (int k = 0; k < n; k++) theta_out[k] = dy4_pll_trigarg(w, dy4_pll_count(T0, k + 1), y[k]);
    {
        const float th = theta_out[n - 1];
        dy4_nco_t o; dy4_sincos_nco_v((double)th, 0, &o, 0);
        state[0] = DY4_D2F(o.c); state[1] = DY4_D2F(o.s);
    }
    state[2] = integ; state[3] = phase; state[4] = (float)dy4_pll_count(T0, n);
    free(rt); free(rhm); free(rtl); free(rth); free(rel); free(reh); free(lo); free(uu); free(th_hat); free(y);
}

/* Randomised check of both certificates of the speculative loop (dy4_spec_fast_check, dy4_spec_check): phaseEst values
 * placed within a few float ulps of every boundary of the two candidates' cells (and at random) — a step declared certain
 * must name exactly RN_f(RN_d(w*T) + phaseEst).  out[0] += probes, [1] += float-certain, [2] += double-certain,
 * [3] += certain-but-wrong (must stay 0), [4] += float-certain but not double-certain (must stay 0). */
void pllspec_fuzz(long n, unsigned long long seed, double w, long* out)
{
    fz_state = seed * 0x9E3779B97F4A7C15ull + 88172645463325252ull;
    for (long it = 0; it < n; it++) {
        const double T = floor(fz_uni() * fz_uni() * 16000000.0) + 1.0;
        const float phase0 = (float)((fz_uni() - 0.5) * (it % 5 == 0 ? 200.0 : 8.0));
        const double wT = DY4_MUL(w, T);
        const float th_true = dy4_pll_trigarg(w, T, phase0);
        const float u = nextafterf(th_true, INFINITY) - th_true;
        const double th_hat = (double)th_true + (fz_uni() - 0.5) * 2.6 * (double)u;
        dy4_row16_t r; float lo, uq;
        dy4_tab_make_row16(th_hat, wT, 0.01f, 0, 0, &r, &lo, &uq);
        if (r.t != r.t) continue;
        float tc_p, hm, tc_o;
        dy4_spec_fast_row(r.t, &tc_p, &hm, &tc_o);
        const int pred_hi = (int)(dy4_d2u_bits(r.t) & 1);
        const float tf = (float)r.t;
        for (int probe = 0; probe < 48; probe++) {
            const int kind = probe % 4;
            const float base = kind == 0 ? tf : kind == 1 ? tf - uq : kind == 2 ? tf + uq : phase0;
            float ph = base;
            const int steps = (int)(fz_next() % 17) - 8;
            for (int s = 0; s < (steps < 0 ? -steps : steps); s++) ph = nextafterf(ph, steps < 0 ? -INFINITY : INFINITY);
            if (kind == 3) ph = (float)((double)phase0 + (fz_uni() - 0.5) * 3.0 * (double)uq);
            const float want = dy4_pll_trigarg(w, T, ph);
            for (int other = 0; other < 2; other++) {
                const float cand = (pred_hi ^ other) ? lo + uq : lo;
                const int fc = dy4_spec_fast_check(ph, other ? tc_o : tc_p, hm), dc = dy4_spec_check(ph, r.t, other);
                out[0]++;
                if (fc) out[1]++;
                if (dc) out[2]++;
                if ((fc || dc) && want != cand) out[3]++;
                if (fc && !dc) out[4]++;
            }
        }
    }
}

static int g_spec_G = 16;
void pllspec_set_G(int G) { g_spec_G = G; }
void pllspec_launch(const float* x, int n, float* state, double w, float Kp, float Ki, float* theta_out, long* stats)
{
    double pred[4];
    pllspec_launch_carry(x, n, state, w, Kp, Ki, theta_out, stats, pred, 0, g_spec_G);
}
