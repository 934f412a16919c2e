/* plltab_host.c — host build of csrc/dy4_plltab.h: the three parts of the table-driven PLL (predict, table, serial
 * pick) run exactly as the kernels of dy4_pll.cu run them, so that the construction can be checked against the
 * reference recurrence on the CPU (tests/test_host_logic.py).  Test infrastructure only.
 * Build: gcc -O2 -ffp-contract=off -mfma -shared -fPIC -o libplltab_host.so plltab_host.c -lm */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include "../../3dy4-real-time-software-defined-radio-_b200/csrc/dy4_plltab.h"

#define SEG 256
#define WARM 1024

/* One launch over n samples of one stream.  state: fbI fbQ integ phase trigOffset (as the reference's PLLState).
 * theta_out[n]: float trigArg after each sample.  stats[0] += picks, stats[1] += direct evaluations, stats[2] += wrong picks (must stay 0). */
void plltab_launch_carry(const float* x, int n, float* state, double w, float Kp, float Ki, float* theta_out, long* stats, double* pred, int carry);
void plltab_launch(const float* x, int n, float* state, double w, float Kp, float Ki, float* theta_out, long* stats)
{
    double pred[4];
    plltab_launch_carry(x, n, state, w, Kp, Ki, theta_out, stats, pred, 0);
}

/* pred: the predictor's own state (integ, phase, sample counter, -) carried from launch to launch as k_pll_predict does:
 * with carry != 0 the prediction starts from it instead of the exact PLL state. */
void plltab_launch_carry(const float* x, int n, float* state, double w, float Kp, float Ki, float* theta_out, long* stats, double* pred, int carry)
{
    dy4_tabrow_t* rows = (dy4_tabrow_t*)malloc(sizeof(dy4_tabrow_t) * (size_t)n);
    double* th_hat = (double*)malloc(sizeof(double) * (size_t)n);
    const double T0 = carry ? pred[2] : (double)state[4];
    const double g_integ = carry ? pred[0] : (double)state[2], g_phase = carry ? pred[1] : (double)state[3];
    /* 1. predict: segment i covers [i*SEG, (i+1)*SEG), warm-up from the launch-start state */
    for (int s0 = 0; s0 < n; s0 += SEG) {
        int kw = s0 - WARM; if (kw < 0) kw = 0;
        double integ = g_integ, phase = g_phase;
        double th_prev = w * dy4_pll_count(T0, kw) + phase;       /* trigArg of step kw-1 (guess unless kw == 0) */
        for (int k = kw; k < s0 + SEG && k < n; k++) {
            th_prev = dy4_pred_step(x[k], th_prev, DY4_MUL(w, dy4_pll_count(T0, k + 1)), (double)Kp, (double)Ki, &integ, &phase);
            if (k >= s0) th_hat[k] = th_prev;
        }
        if (s0 + SEG >= n) { pred[0] = integ; pred[1] = phase; pred[2] = dy4_pll_count(T0, n); pred[3] = T0; }
    }
    /* 2. table */
    for (int k = 0; k < n; k++)
        dy4_tab_make_row(th_hat[k], DY4_MUL(w, dy4_pll_count(T0, k + 1)), k + 1 < n ? x[k + 1] : 0.0f, k + 1 < n, T0 + (double)k < (double)DY4_TAB_EARLY, Kp, Ki, &rows[k]);
    /* 3. serial */
    float fbI = state[0], fbQ = state[1], integ = state[2], phase = state[3];
    {
        const float eI = DY4_FMULF((x[0] == 0.0f ? 1.0f : x[0]), fbI), eQ = DY4_FMULF(x[0], -fbQ);
        dy4_pll_filter(DY4_D2F(atan2((double)eQ, (double)eI)), Kp, Ki, &integ, &phase);
    }
    for (int k = 0; k + 1 < n; k++) {
        const dy4_tabrow_t* r = &rows[k];
        int up;
        const float th = dy4_pll_trigarg(w, dy4_pll_count(T0, k + 1), phase);     /* what k_nco_phase derives from the stored phase */
        theta_out[k] = th;
        if (dy4_tab_pick(phase, r->t, r->hu, r->hm, &up)) {
            if (th != (up ? r->lo + r->u : r->lo)) stats[2]++;                     /* a certain pick that is wrong: must never happen */
            dy4_pll_filter_ab(up ? r->a_hi : r->a_lo, up ? r->b_hi : r->b_lo, &integ, &phase);
            stats[0]++;
        } else {
            dy4_pll_filter(dy4_next_errorD((double)th, x[k + 1]), Kp, Ki, &integ, &phase);
            stats[1]++;
        }
    }
    {
        const float th = dy4_pll_trigarg(w, dy4_pll_count(T0, n), phase);
        dy4_nco_t o; dy4_sincos_nco_v((double)th, 0, &o, 0);
        theta_out[n - 1] = th;
        state[0] = DY4_D2F(o.c); state[1] = DY4_D2F(o.s);
    }
    state[2] = integ; state[3] = phase; state[4] = (float)dy4_pll_count(T0, n);
    free(rows); free(th_hat);
}

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

/* Randomised check of the pick's certificate (dy4_tab_pick): for random sample counters, loop phases and predictions, and
 * for phaseEst values placed within a few ulps of every boundary the row defines (the threshold t, the far ends t -+ u) as
 * well as at random, a pick declared certain must name exactly RN_f(RN_d(w*T) + phaseEst).
 * out[0] += cases, out[1] += certain picks, out[2] += certain-but-wrong picks (must stay 0), out[3] += uncertain. */
static unsigned long long fz_state;
static unsigned long long fz_next(void) { fz_state ^= fz_state << 13; fz_state ^= fz_state >> 7; fz_state ^= fz_state << 17; return fz_state; }
static double fz_uni(void) { return (double)(fz_next() >> 11) * (1.0 / 9007199254740992.0); }

void plltab_fuzz(long n, unsigned long long seed, double w, long* out)
{
    fz_state = seed * 0x9E3779B97F4A7C15ull + 88172645463325252ull;
    for (long it = 0; it < n; it++) {
        const double T = floor(fz_uni() * fz_uni() * 16000000.0) + 1.0;                  /* more weight on small counters (small u) */
        const float phase0 = (float)((fz_uni() - 0.5) * (it % 5 == 0 ? 200.0 : 8.0));
        const double wT = DY4_MUL(w, T);
        const float th_true = dy4_pll_trigarg(w, T, phase0);
        const float u = nextafterf(th_true, INFINITY) - th_true;
        const double th_hat = (double)th_true + (fz_uni() - 0.5) * 2.6 * (double)u;      /* prediction within +-1.3 grid points */
        dy4_tabrow_t r;
        dy4_tab_make_row(th_hat, wT, 0.01f, 0, 0, 0.02666f, 0.0003555f, &r);
        if (r.t != r.t) continue;                                                         /* unusable row (binade edge) */
        for (int probe = 0; probe < 40; probe++) {
            float ph;
            const int kind = probe % 4;
            const float base = kind == 0 ? r.t : kind == 1 ? r.t - r.u : kind == 2 ? r.t + r.u : phase0;
            ph = base;
            const int steps = (int)(fz_next() % 17) - 8;                                  /* -8 .. +8 ulps around the boundary */
            for (int s = 0; s < (steps < 0 ? -steps : steps); s++) ph = nextafterf(ph, steps < 0 ? -INFINITY : INFINITY);
            if (kind == 3) ph = (float)((double)phase0 + (fz_uni() - 0.5) * 3.0 * (double)r.u);
            int up;
            out[0]++;
            if (dy4_tab_pick(ph, r.t, r.hu, r.hm, &up)) {
                const float want = dy4_pll_trigarg(w, T, ph);
                const float got = up ? r.lo + r.u : r.lo;
                out[1]++;
                if (want != got) out[2]++;
            } else out[3]++;
        }
    }
}

/* ---- the speculative loop (k_pll_spec, csrc/dy4_pll.cu): 16-byte rows, groups of G steps run on the predicted candidate,
 * certified afterwards by the float certificate, resumed at the first step that is not certain.  The control flow below is
 * the kernel's (SpecLoop / spec_trip / recover), lane by lane.
 * stats[0] += steps taken on a float certificate, [1] += direct evaluations, [2] += certified-but-wrong steps (must stay 0),
 * [3] += groups run, [4] += flips, [5] += steps taken on the double certificate. */
typedef struct { int r, forced, pr, pn, pforced; float integ, phase; } spec_loop_t;

void pllspec_launch_carry(const float* x, int n, float* state, double w, float Kp, float Ki, float* theta_out, long* stats, double* pred, int carry, int G)
{
    dy4_row16_t* rows = (dy4_row16_t*)malloc(sizeof(dy4_row16_t) * (size_t)(n + 64));
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
            if (k >= s0) th_hat[k] = th_prev;
        }
        if (s0 + SEG >= n) { pred[0] = integ; pred[1] = phase; pred[2] = dy4_pll_count(T0, n); pred[3] = T0; }
    }
    for (int k = 0; k < n; k++)
        dy4_tab_make_row16(th_hat[k], DY4_MUL(w, dy4_pll_count(T0, k + 1)), k + 1 < n ? x[k + 1] : 0.0f, k + 1 < n, T0 + (double)k < (double)DY4_TAB_EARLY,
                           &rows[k], &lo[k], &uu[k]);
    for (int k = n; k < n + 64; k++) { rows[k].t = NAN; rows[k].e_p = rows[k].e_o = NAN; lo[k] = uu[k] = 0; }
    float fbI = state[0], fbQ = state[1], integ = state[2], phase = state[3];
    {
        const float eI = DY4_FMULF((x[0] == 0.0f ? 1.0f : x[0]), fbI), eQ = DY4_FMULF(x[0], -fbQ);
        dy4_pll_filter(DY4_D2F(atan2((double)eQ, (double)eI)), Kp, Ki, &integ, &phase);
    }
    const int n_rows = n - 1;                  /* the kernel's kd (direct start-up part) is 0 here: early rows are NaN and go the careful way */
    float cap[2][40][2];
    spec_loop_t L = {0, 0, 0, 0, 0, integ, phase};
    int buf = 0;
    /* a step from (integ, phase) with errorD = eD; checks a certified candidate against the true trigArg */
#define CHECK_CAND(k, ph, other) do { \
        const int hi_ = (int)(dy4_d2u_bits(rows[k].t) & 1) ^ (other); \
        const float cand_ = hi_ ? lo[k] + uu[k] : lo[k]; \
        if (dy4_pll_trigarg(w, dy4_pll_count(T0, (k) + 1), (ph)) != cand_) stats[2]++; } while (0)
    while (L.r < n_rows || L.pn > 0) {
        /* certify the pending group from cap[buf ^ 1] */
        int j = L.pn;
        for (int lane = 0; lane < L.pn; lane++) {
            float tc_p, hm, tc_o;
            dy4_spec_fast_row(rows[L.pr + lane].t, &tc_p, &hm, &tc_o);
            const float myph = cap[buf ^ 1][lane][1];
            if (!dy4_spec_fast_check(myph, (lane == 0 && L.pforced) ? tc_o : tc_p, hm)) { j = lane; break; }
        }
        for (int lane = 0; lane < j; lane++) {
            const float myph = cap[buf ^ 1][lane][1];
            y[L.pr + lane] = myph;
            CHECK_CAND(L.pr + lane, myph, (lane == 0 && L.pforced) ? 1 : 0);
            stats[0]++;
        }
        /* the chain of the group at L.r (runs whether or not the pending group holds, as the kernel's does) */
        float si = L.integ, sp = L.phase;
        for (int i = 0; i < G; i++) {
            cap[buf][i][0] = si; cap[buf][i][1] = sp;
            const dy4_row16_t* rw = &rows[L.r + i < n + 64 ? L.r + i : n + 63];
            const float e = (i == 0 && L.forced) ? rw->e_o : rw->e_p;
            dy4_pll_filter_ab(DY4_FMULF(Ki, e), DY4_FMULF(Kp, e), &si, &sp);
        }
        cap[buf][G][0] = si; cap[buf][G][1] = sp;
        stats[3]++;
        if (j < L.pn) {                                                  /* recover */
            const int k = L.pr + j;
            const float ig = cap[buf ^ 1][j][0], ph = cap[buf ^ 1][j][1];
            if (j == 0 && L.pforced) {
                float eD;
                if (dy4_spec_check(ph, rows[k].t, 0)) { eD = rows[k].e_p; CHECK_CAND(k, ph, 0); stats[5]++; }
                else if (dy4_spec_check(ph, rows[k].t, 1)) { eD = rows[k].e_o; CHECK_CAND(k, ph, 1); stats[5]++; }
                else { eD = dy4_next_errorD((double)dy4_pll_trigarg(w, dy4_pll_count(T0, k + 1), ph), x[k + 1]); stats[1]++; }
                float ni = ig, np = ph;
                dy4_pll_filter(eD, Kp, Ki, &ni, &np);
                y[k] = ph;
                L.integ = ni; L.phase = np; L.r = k + 1; L.forced = 0;
            } else { L.integ = ig; L.phase = ph; L.r = k; L.forced = 1; stats[4]++; }
            L.pn = 0; L.pr = L.r; L.pforced = 0;
            continue;                                                   /* (the kernel restarts into buffer 0; which buffer is immaterial) */
        }
        int nv = n_rows - L.r; if (nv > G) nv = G; if (nv < 0) nv = 0;
        L.pr = L.r; L.pn = nv; L.pforced = L.forced; L.forced = 0;
        if (nv < G) { si = cap[buf][nv][0]; sp = cap[buf][nv][1]; }
        L.r += nv; L.integ = si; L.phase = sp;
        buf ^= 1;
    }
    integ = L.integ; phase = L.phase;
    y[n - 1] = phase;
    for (int k = 0; k < n; k++) theta_out[k] = dy4_pll_trigarg(w, dy4_pll_count(T0, k + 1), y[k]);
    {
        const float th = theta_out[n - 1];
        dy4_nco_t o; dy4_sincos_nco_v((double)th, 0, &o, 0);
        state[0] = DY4_D2F(o.c); state[1] = DY4_D2F(o.s);
    }
    state[2] = integ; state[3] = phase; state[4] = (float)dy4_pll_count(T0, n);
    free(rows); free(lo); free(uu); free(th_hat); free(y);
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
