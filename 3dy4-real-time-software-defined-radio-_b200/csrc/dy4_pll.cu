// dy4_pll.cu — per-stream phase-locked loop + NCO, one thread per stream.
//
// Replaces fmPLL, src/filter.cpp:174-228 (called from backend(), project.cpp:123).
// The loop is a true recurrence (feedbackI/Q, integrator, phaseEst, trigOffset)
// so it stays sequential inside a stream and is parallel only ACROSS streams.
// Arithmetic mirrors the reference's mixed precision exactly: all state is
// float; atan2/sin/cos are DOUBLE evaluations of float-widened arguments
// narrowed back to float; trigArg is a double expression narrowed to float
// (filter.cpp:214); trigOffset is a float counter.  None of it may be fused or
// reassociated: the float trigArg quantises phase to up to 0.125 rad late in a
// stream, which makes the trajectory chaotic in the last bit.
//
// With one thread per stream the kernel is bound by the LATENCY of that dependent
// chain, so the double-precision math is not libm's: dy4_pllmath.h evaluates
// sin/cos with one shared reduction and obtains atan2 as "known reduced phase +
// small correction", as fixed sequences of IEEE operations that were validated
// on the host against glibc (tools/pllmath_check.c) and narrow to the same floats.
//
// Memory: each lane walks its own row; loads are issued four samples ahead as
// one 16-byte load and NCO values leave as 16-byte stores, both off the
// dependent chain.
#include "dy4_common.cuh"
#include "dy4_kernels.h"
#include "dy4_internal.h"
#include "dy4_pllmath.h"
#include "dy4_plltab.h"

#include <algorithm>
#include <cstdlib>
#include <string>

namespace {

struct PllConst { double w; float Kp, Ki, ncoScale, phaseAdjust; };

// float state of the reference (filter.cpp:174 signature) plus the binade tracker for trigArg's rounding
struct PllRegs { float fbI, fbQ, integ, phase; double trigOffset; double lim, lim2, magic; };   // trigOffset: float VALUE kept in a double

__device__ __forceinline__ float nco_value(float trigArg, float ncoScale, float phaseAdjust)
{
    const float narg = __fadd_rn(__fmul_rn(trigArg, ncoScale), phaseAdjust);   // float, as filter.cpp:219/:221
    dy4_nco_t on;
    dy4_sincos_nco((double)narg, 0, &on);
    return __double2float_rn(on.c);
}

// Everything after the phase detector: loop filter, phase accumulator, NCO phase (filter.cpp:206-217).
// Returns the new trigArg as a double holding a float value (the NCO output is a pure function of it and
// is evaluated by k_nco).  `o` receives sin/cos of it and the reference angle for the next detector call.
template <int GRID, bool SELB>
__device__ __forceinline__ double pll_advance(float eD, PllRegs& s, const PllConst& c, dy4_nco_t& o, bool next_neg)
{
    s.integ = __fadd_rn(s.integ, __fmul_rn(c.Ki, eD));                         // :207
    s.phase = __fadd_rn(s.phase, __fadd_rn(__fmul_rn(c.Kp, eD), s.integ));     // :210
    s.trigOffset = fmin(s.trigOffset + 1.0, 16777216.0);                       // :213 float counter: exact below 2^24, sticks there (16777217 rounds back)
    const double arg = __dadd_rn(__dmul_rn(c.w, s.trigOffset), (double)s.phase);            // :214, in double
    // :214 narrows to float.  Inside the tracked binade [lim, 2 lim) that is an add/subtract of a magic
    // constant on the double pipe (ties to even, same grid) instead of two conversions.
    // GRID_SAFE: the caller has shown that arg stays inside the binade for this step, no check needed.
    double th;
    if (GRID == 0) th = (double)__double2float_rn(arg);
    else th = dy4_round_to_float_grid(arg, s.magic);
    if (GRID == 1 && !(arg >= s.lim && arg < s.lim2)) {                        // binade changed (24 times per stream) or start-up
        th = (double)__double2float_rn(arg);
        const int hi = __double2hiint(th);
        if (hi >= 0x38100000 && hi < 0x47e00000) {                             // positive, comfortably inside float's normal range
            s.lim = __hiloint2double(hi & 0x7ff00000, 0);
            s.lim2 = s.lim + s.lim;
            s.magic = s.lim * 805306368.0;                                     // 1.5 * 2^29
        } else { s.lim = 0.0; s.lim2 = 0.0; s.magic = 0.0; }
    }
    dy4_sincos_nco_v(th, next_neg, &o, SELB);
    s.fbI = __double2float_rn(o.c);                                            // :216
    s.fbQ = __double2float_rn(o.s);                                            // :217
    return th;
}

// The libm phase detector: used for the first sample of a launch (the carried feedbackI/Q come from the
// caller) and for inputs the fast detector does not cover (0, denormal, inf, NaN).
__device__ __noinline__ float detector_libm(float x, float fbI, float fbQ)
{
    const float eI = __fmul_rn((x == 0.0f ? 1.0f : x), fbI);                   // filter.cpp:192
    const float eQ = __fmul_rn(x, -fbQ);                                       // :193
    return __double2float_rn(atan2((double)eQ, (double)eI));                   // :200
}

__device__ __forceinline__ bool fast_ok(float x) { const float ax = fabsf(x); return ax > 1e-20f && ax < 1e20f; }

// Fast step: x is a normal, finite, non-zero float and `o` describes the current feedbackI/Q.  Straight-line code.
template <int GRID, bool SELB>
__device__ __forceinline__ double pll_step_fast(float x, double inv_x, float x_next, PllRegs& s, const PllConst& c, dy4_nco_t& o)
{
    const float eI = __fmul_rn(x, s.fbI);                                      // filter.cpp:192 (x != 0)
    const float eQ = __fmul_rn(x, -s.fbQ);                                     // :193
    const float eD = __double2float_rn(dy4_detector_atan2((double)eQ, (double)eI, &o, inv_x));  // :200
    return pll_advance<GRID, SELB>(eD, s, c, o, x_next < 0.0f);
}

// `o_matches_x`: o.base_* was prepared for the sign of this x (true after any step that was given x as x_next)
template <bool SELB>
__device__ __forceinline__ double pll_step_any(float x, float x_next, PllRegs& s, const PllConst& c, dy4_nco_t& o, bool o_valid)
{
    if (o_valid && fast_ok(x)) return pll_step_fast<1, SELB>(x, dy4_recip(x), x_next, s, c, o);
    return pll_advance<1, SELB>(detector_libm(x, s.fbI, s.fbQ), s, c, o, x_next < 0.0f);
}

template <int GRID, bool SELB>
__global__ void __launch_bounds__(128)
k_pll(const float* __restrict__ in, long long in_stride, const double* __restrict__ inv, double* __restrict__ theta, long long wide_stride,
      float* __restrict__ nco0, float* __restrict__ state, int n, int n_streams, PllConst c)
{
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n_streams) return;
    float* st = state + (long long)s * 8;
    PllRegs r = {st[0], st[1], st[2], st[3], (double)st[4], 0.0, 0.0, 0.0};
    nco0[s] = st[5];                             // nco_state opens this launch's NCO row (filter.cpp:184)
    const float* x = in + (long long)s * in_stride;
    const double* ix = inv + (long long)s * wide_stride;      // 1/x per sample from k_pll_prep, 0 where the fast detector does not apply
    double* y = theta + (long long)s * wide_stride;
    dy4_nco_t o;
    o.c = 1.0; o.s = 0.0; o.base_hi = 0.0; o.base_lo = 0.0;
    if (n <= 0) return;
    // first sample: the carried feedbackI/Q are whatever the caller holds, so the detector is libm's
    double th = pll_advance<1, SELB>(detector_libm(x[0], r.fbI, r.fbQ), r, c, o, n > 1 && x[1] < 0.0f);
    y[0] = th;
    int k = 1;
    for (; k < n && (k & 3); k++) { th = pll_step_any<SELB>(x[k], k + 1 < n ? x[k + 1] : 0.0f, r, c, o, true); y[k] = th; }
    // two groups of four samples (and their reciprocals) are kept in registers ahead of the recurrence (the step
    // that closes a group needs the sign of the next group's first sample), so no load is ever waited on
    const float4 z4 = make_float4(0, 0, 0, 0);
    const double2 zd = make_double2(0, 0);
    float4 v = k + 4 <= n ? *reinterpret_cast<const float4*>(x + k) : z4;
    double2 ia = k + 4 <= n ? *reinterpret_cast<const double2*>(ix + k) : zd, ib = k + 4 <= n ? *reinterpret_cast<const double2*>(ix + k + 2) : zd;
    float4 v2 = k + 8 <= n ? *reinterpret_cast<const float4*>(x + k + 4) : z4;
    double2 ia2 = k + 8 <= n ? *reinterpret_cast<const double2*>(ix + k + 4) : zd, ib2 = k + 8 <= n ? *reinterpret_cast<const double2*>(ix + k + 6) : zd;
    for (; k + 4 <= n; k += 4) {
        const float4 cur = v;
        const double2 ca = ia, cb = ib;
        v = v2; ia = ia2; ib = ib2;
        if (k + 12 <= n) {
            v2 = *reinterpret_cast<const float4*>(x + k + 8);
            ia2 = *reinterpret_cast<const double2*>(ix + k + 8);
            ib2 = *reinterpret_cast<const double2*>(ix + k + 10);
        }
        float nx = v.x;                                                       // the sample after this group
        if (k + 8 > n) nx = (k + 4 < n) ? x[k + 4] : 0.0f;
        // Per step the phase argument moves by w + (Kp*eD + integ): |Kp*eD| <= 0.0267*pi and the integrator drifts by
        // < 1.2e-3 per step, so with |integ| < 0.1 four steps move it by at most 4*(w+0.2) and at least 4*min(0,w-0.2).
        // If that range stays inside the tracked binade the narrowing to float needs no per-step check.
        const double up4 = 4.0 * (c.w + 0.2), dn4 = fmin(0.0, 4.0 * (c.w - 0.2)) + fmin(0.0, c.w - 0.2);
        const bool grid_safe = GRID != 2 || ((th + dn4 >= r.lim) && (th + up4 < r.lim2) && (fabsf(r.integ) < 0.1f) && (c.Kp <= 0.0267f));
        const bool fast4 = (ca.x != 0.0) && (ca.y != 0.0) && (cb.x != 0.0) && (cb.y != 0.0);
        double t0, t1, t2, t3;
        if (fast4 && grid_safe) {
            t0 = pll_step_fast<GRID, SELB>(cur.x, ca.x, cur.y, r, c, o);
            t1 = pll_step_fast<GRID, SELB>(cur.y, ca.y, cur.z, r, c, o);
            t2 = pll_step_fast<GRID, SELB>(cur.z, cb.x, cur.w, r, c, o);
            t3 = pll_step_fast<GRID, SELB>(cur.w, cb.y, nx, r, c, o);
        } else {
            t0 = pll_step_any<SELB>(cur.x, cur.y, r, c, o, true);
            t1 = pll_step_any<SELB>(cur.y, cur.z, r, c, o, true);
            t2 = pll_step_any<SELB>(cur.z, cur.w, r, c, o, true);
            t3 = pll_step_any<SELB>(cur.w, nx, r, c, o, true);
        }
        th = t3;
        *reinterpret_cast<double2*>(y + k) = make_double2(t0, t1);
        *reinterpret_cast<double2*>(y + k + 2) = make_double2(t2, t3);
    }
    for (; k < n; k++) { th = pll_step_any<SELB>(x[k], k + 1 < n ? x[k + 1] : 0.0f, r, c, o, true); y[k] = th; }
    st[0] = r.fbI; st[1] = r.fbQ; st[2] = r.integ; st[3] = r.phase; st[4] = (float)r.trigOffset;
    st[5] = nco_value((float)th, c.ncoScale, c.phaseAdjust);      // nco_state for the next launch (filter.cpp:218-219)
}


// ====================================================================================================================
// Table-driven PLL (dy4_plltab.h): predict -> table -> serial pick.  Bit-identical to k_pll by construction (every step
// is either a pick among exactly evaluated candidates, certain under its error budget, or a direct evaluation).
// ====================================================================================================================
constexpr int PRED_SEG = 512;      // samples per predictor thread: 256 -> 7.20, 512 -> 6.91, 1024 -> 7.67 ms per step (latency of the thread's own chain vs redundant warm-up work)
constexpr int PRED_WARM = 1024;    // warm-up steps before a segment: the loop forgets its state as 0.98657^k (1e-6 after 1024)

// 1. predicted phaseEst of every sample (a float: |phaseEst| is a few radians, so its 2^-24 relative precision is far below the
// float grid of trigArg = w*T + phaseEst it is used to locate) -> the theta row.  One thread per (stream, segment); the first WARM samples
// of a launch start from the exact carried state, later segments from that state as a guess plus the warm-up.
// `pred_in` / `pred_out`: [n_streams][8] doubles (integ, phase, sample counter after / before the launch, turns) — the
// predictor's own state at the start / end of the launch.  carry == 0: the launch starts from the exact carried PLL
// state (the loop of the previous launch has finished).  carry != 0: from where the previous launch's prediction ended,
// so that prediction and table of sub-chunk c+1 do not wait for the serial loop of sub-chunk c (dy4_pipeline.cu).
// The loop locks to the pilot modulo 2*pi, and after a loss of lock the predictor may settle a whole number of turns
// away from the true phaseEst — a table built there never matches.  So every serial launch reports how many turns its
// prediction was off (`need`, k_pll_sel), and with carry == 2 the launch two later shifts its start by that much.
__global__ void __launch_bounds__(128)
k_pll_predict(const float* __restrict__ in, long long in_stride, const float* __restrict__ state, const double* __restrict__ pred_in,
              double* __restrict__ pred_out, const double* __restrict__ need, int carry, float* __restrict__ ph_hat, long long ph_stride,
              int n, PllConst c)
{
    const int s = blockIdx.x;
    const int s0 = (blockIdx.y * blockDim.x + threadIdx.x) * PRED_SEG;
    if (s0 >= n) return;
    const float* st = state + (long long)s * 8;
    const float* x = in + (long long)s * in_stride;
    float* y = ph_hat + (long long)s * ph_stride;          // predicted phaseEst, as a float: the table kernel adds RN_d(w*T) back
    const double Kp = (double)c.Kp, Ki = (double)c.Ki;
    double integ, phase, T0, turns = 0.0;
    if (carry) {
        integ = pred_in[8 * s]; phase = pred_in[8 * s + 1]; T0 = pred_in[8 * s + 2]; turns = pred_in[8 * s + 4];
        if (carry == 2) {
            const double d = need[s] - turns;
            if (fabs(d) < 1e6) { phase = fma(d, 6.28318530717958647692, phase); turns = need[s]; }
        }
    } else { integ = (double)st[2]; phase = (double)st[3]; T0 = (double)st[4]; }
    const int kw = max(0, s0 - PRED_WARM), k1 = min(n, s0 + PRED_SEG);
    double th_prev = c.w * dy4_pll_count(T0, kw) + phase;
    int k = kw;
    for (; k + 4 <= k1; k += 4) {                                   // kw, s0 are multiples of 4; rows are 16-byte aligned
        const float4 v = __ldg(reinterpret_cast<const float4*>(x + k));
        th_prev = dy4_pred_step(v.x, th_prev, __dmul_rn(c.w, dy4_pll_count(T0, k + 1)), Kp, Ki, &integ, &phase); const float p0 = (float)phase;
        th_prev = dy4_pred_step(v.y, th_prev, __dmul_rn(c.w, dy4_pll_count(T0, k + 2)), Kp, Ki, &integ, &phase); const float p1 = (float)phase;
        th_prev = dy4_pred_step(v.z, th_prev, __dmul_rn(c.w, dy4_pll_count(T0, k + 3)), Kp, Ki, &integ, &phase); const float p2 = (float)phase;
        th_prev = dy4_pred_step(v.w, th_prev, __dmul_rn(c.w, dy4_pll_count(T0, k + 4)), Kp, Ki, &integ, &phase); const float p3 = (float)phase;
        if (k >= s0) *reinterpret_cast<float4*>(y + k) = make_float4(p0, p1, p2, p3);
    }
    for (; k < k1; k++) {
        th_prev = dy4_pred_step(__ldg(x + k), th_prev, __dmul_rn(c.w, dy4_pll_count(T0, k + 1)), Kp, Ki, &integ, &phase);
        if (k >= s0) y[k] = (float)phase;
    }
    if (k1 == n) {                                                  // the thread of the last segment: state for the next launch
        pred_out[8 * s] = integ; pred_out[8 * s + 1] = phase; pred_out[8 * s + 2] = dy4_pll_count(T0, n); pred_out[8 * s + 3] = T0;
        pred_out[8 * s + 4] = turns;
    }
}

constexpr int TAB_EARLY = DY4_TAB_EARLY;

__device__ __forceinline__ unsigned tab_smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void tab_mbar_init(unsigned long long* bar)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(tab_smem_u32(bar)));
}
template <int BYTES>
__device__ __forceinline__ void tab_bulk_load(void* dst, const void* src, unsigned long long* bar)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(tab_smem_u32(bar)), "n"(BYTES) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(tab_smem_u32(dst)), "l"(src), "n"(BYTES), "r"(tab_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool tab_mbar_try(unsigned long long* bar, unsigned parity)
{
    unsigned ok;
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                 : "=r"(ok) : "r"(tab_smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}

// Samples k0+1 .. k1 through the steps of the direct loop (pll_step_fast: straight-line groups of four whose independent
// work overlaps the dependent chain, ~450 cycles per sample).  *rp / *op hold the state after sample k0 on entry, after
// sample k1 on return; y[k] = phaseEst after sample k for k in [k0, k1).
__device__ __noinline__ void tab_direct_span(const float* __restrict__ x, float* __restrict__ y, int n, int k0, int k1, PllRegs* rp, dy4_nco_t* op,
                                             double w, float Kp, float Ki)
{
    PllConst c; c.w = w; c.Kp = Kp; c.Ki = Ki; c.ncoScale = 0.0f; c.phaseAdjust = 0.0f;
    PllRegs r = *rp;
    dy4_nco_t o = *op;
    auto xat = [&](int i) { return x[min(i, n - 1)]; };
    int k = k0;                                              // r holds the state after sample k
    for (; k < k1 && (k & 3); k++) { y[k] = r.phase; pll_step_any<false>(xat(k + 1), xat(k + 2), r, c, o, true); }
    float v0 = xat(k + 1), v1 = xat(k + 2), v2 = xat(k + 3), v3 = xat(k + 4), v4 = xat(k + 5);
#pragma unroll 1
    for (; k + 4 <= k1; k += 4) {
        const float c0 = v0, c1 = v1, c2 = v2, c3 = v3, c4 = v4;
        v0 = c4; v1 = xat(k + 6); v2 = xat(k + 7); v3 = xat(k + 8); v4 = xat(k + 9);
        float q0, q1, q2, q3;
        if (fast_ok(c0) && fast_ok(c1) && fast_ok(c2) && fast_ok(c3)) {
            q0 = r.phase; pll_step_fast<1, false>(c0, dy4_recip(c0), c1, r, c, o);
            q1 = r.phase; pll_step_fast<1, false>(c1, dy4_recip(c1), c2, r, c, o);
            q2 = r.phase; pll_step_fast<1, false>(c2, dy4_recip(c2), c3, r, c, o);
            q3 = r.phase; pll_step_fast<1, false>(c3, dy4_recip(c3), c4, r, c, o);
        } else {
            q0 = r.phase; pll_step_any<false>(c0, c1, r, c, o, true);
            q1 = r.phase; pll_step_any<false>(c1, c2, r, c, o, true);
            q2 = r.phase; pll_step_any<false>(c2, c3, r, c, o, true);
            q3 = r.phase; pll_step_any<false>(c3, c4, r, c, o, true);
        }
        *reinterpret_cast<float4*>(y + k) = make_float4(q0, q1, q2, q3);
    }
    for (; k < k1; k++) { y[k] = r.phase; pll_step_any<false>(xat(k + 1), xat(k + 2), r, c, o, true); }
    *rp = r; *op = o;
}

// ====================================================================================================================
// 16-byte rows (dy4_plltab.h 2b) and the serial loop on them, k_pll_sel.
// ====================================================================================================================
constexpr int SPEC_SG = 128;                         // rows per bulk copy
constexpr int SPEC_SLOTS = 4;
constexpr int SPEC_R = SPEC_SG * SPEC_SLOTS;         // ring size in rows

// The table kernel for k_pll_sel: one CHAIN-READY 32-byte row per sample, so the serial warp spends nothing on preparing operands:
//   (Ki e_lo, Ki e_hi, Kp e_lo, Kp e_hi)   the products of filter.cpp:207,210 for the lower / upper candidate of trigArg
//   (t, tc_lo, tc_hi, hm)                  float threshold between the two in the phaseEst domain, centres of the two cells and the
//                                          certified half-width of a cell (dy4_spec_fast_row)
// (A 16-byte row (t, hm, e_lo, e_hi) with the four products formed inside the loop was measured too: the four extra FMUL per
// step cost the lone warp 39.8 instead of 26.5 ns per sample — issue slots, not bytes, are what the serial loop is short of.)
__global__ void __launch_bounds__(128)
k_pll_table_ops(const float* __restrict__ in, long long in_stride, const double* __restrict__ pred_out,
                const float* __restrict__ ph_hat, long long ph_stride, float4* __restrict__ tab, long long tab_stride, int* __restrict__ risk, int n, PllConst c)
{
    const int s = blockIdx.x;
    const int k = blockIdx.y * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const double T0 = pred_out[8 * s + 3];                          // sample counter at the start of this launch (k_pll_predict)
    const float* x = in + (long long)s * in_stride;
    dy4_row16_t r;
    int rk = 0;                                                     // evaluations of this row that narrow a near-tie (dy4_near_float_tie)
    const double wT = __dmul_rn(c.w, dy4_pll_count(T0, k + 1));
    dy4_tab_make_row16_r(wT + (double)__ldg(ph_hat + (long long)s * ph_stride + k), wT,
                       k + 1 < n ? __ldg(x + k + 1) : 0.0f, k + 1 < n, T0 + (double)k < (double)DY4_TAB_EARLY, &r, nullptr, nullptr, &rk);
    if (rk && risk) atomicAdd(risk + s, rk);                        // ~1e-8 per evaluation
    const bool ph = (dy4_d2u_bits(r.t) & 1ull) != 0;                // the predicted candidate is the upper one
    const float e_lo = ph ? r.e_o : r.e_p, e_hi = ph ? r.e_p : r.e_o;
    float tc_p, hm, tc_o;
    dy4_spec_fast_row(r.t, &tc_p, &hm, &tc_o);
    float4* o = tab + (long long)s * tab_stride + 2 * (long long)k;
    o[0] = make_float4(__fmul_rn(c.Ki, e_lo), __fmul_rn(c.Ki, e_hi), __fmul_rn(c.Kp, e_lo), __fmul_rn(c.Kp, e_hi));
    o[1] = make_float4(__double2float_rn(r.t), ph ? tc_o : tc_p, ph ? tc_p : tc_o, hm);
}

__device__ int g_spec_stats[4];      // rows, groups, direct steps, - (development counters, DY4_PLL_STATS=1)

__device__ __forceinline__ float4 spec_lds128(unsigned a) { float4 v; asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a)); return v; }

// one step evaluated directly from state (integ, phase) after sample k: filter.cpp:192-214 through dy4_pllmath.h
__device__ __noinline__ float2 spec_direct_step(const float* __restrict__ x, int n, int k, double T0, double w, float Kp, float Ki, float integ, float phase)
{
    const float th = dy4_pll_trigarg(w, dy4_pll_count(T0, k + 1), phase);
    dy4_pll_filter(dy4_next_errorD((double)th, x[min(k + 1, n - 1)]), Kp, Ki, &integ, &phase);
    return make_float2(integ, phase);
}

// integ before step j of the group whose rows start at shared address `base`, from the state (gi, gp) before its step 0: the
// chain's own arithmetic, step by step (runs only when a step of the group was not certain)
__device__ __noinline__ float spec_replay_integ(unsigned base, int j, float gi, float gp)
{
    for (int i = 0; i < j; i++) {
        const float4 A = spec_lds128(base + 32u * i);
        float T;
        asm volatile("ld.shared.f32 %0, [%1];" : "=f"(T) : "r"(base + 32u * i + 16u));
        const float i_lo = __fadd_rn(gi, A.x), i_hi = __fadd_rn(gi, A.y);
        const float p_lo = __fadd_rn(gp, __fadd_rn(A.z, i_lo)), p_hi = __fadd_rn(gp, __fadd_rn(A.w, i_hi));
        const bool up = gp > T;
        gi = up ? i_hi : i_lo;
        gp = up ? p_hi : p_lo;
    }
    return gi;
}

// ====================================================================================================================
// k_pll_sel: the serial loop.  Both candidates carried, the pick on the chain, its certificate OFF it.
// One warp per stream, all lanes run the same chain.  Per step: both candidates' loop-filter updates (six float adds, the
// table's products as operands), ONE compare of phaseEst with the row's float threshold and two selects — nothing else
// sits between two steps.  Whether that pick was CERTAIN (phaseEst at least the rounding budget away from the threshold
// and from the far ends of the two cells, dy4_spec_fast_check) is decided afterwards: lane i parks phaseEst of "before
// step i" in a register on the way (one LOP3), and after B steps every lane takes the certificate of its own step at
// once — one vote for the group.  An uncertain step (the first ~0.1 s of a stream, where the float threshold's half-ulp
// is a visible fraction of the cell; binade edges; a loop out of lock) is evaluated directly (dy4_pllmath.h) and the loop
// resumes behind it: the result is the reference's by construction.
// Rows arrive chain-ready (k_pll_table_ops) by one bulk asynchronous copy per 128 samples straight into the ring the
// chain reads: the warp neither waits on global memory nor spends issue slots on it.
// ====================================================================================================================
#ifndef DY4_SEL_PACKED
#define DY4_SEL_PACKED 1
#endif
template <int B>
__global__ void __launch_bounds__(32)
k_pll_sel(const float* __restrict__ in, long long in_stride, const float4* __restrict__ tab, long long tab_stride,
          float* __restrict__ phase_out, long long phase_stride, float* __restrict__ nco0, float* __restrict__ tstart,
          const double* __restrict__ pred, double* __restrict__ need, float* __restrict__ state, int* __restrict__ stats, int n, int n_streams, PllConst c)
{
    static_assert(B <= 32 && 2 * B <= SPEC_SG, "lane i certifies step i of a group");
    constexpr int RQ = 2;                                         // 16-byte words per row
    __shared__ __align__(128) float4 ring[RQ * (SPEC_R + 32)];    // the last 32 rows mirror the first 32 (a group may straddle the end)
    __shared__ __align__(8) unsigned long long bars[SPEC_SLOTS];
    const int lane = threadIdx.x;
    const int s = blockIdx.x;
    if (s >= n_streams || n <= 0) return;
    float* st = state + (long long)s * 8;
    float fbI = st[0], fbQ = st[1], integ = st[2], phase = st[3];
    const double T0 = (double)st[4];
    const float nco_carry = st[5];
    const float* x = in + (long long)s * in_stride;
    const float4* rows = tab + (long long)s * tab_stride;
    float* y = phase_out + (long long)s * phase_stride;
    const int n_pick = n - 1;                        // steps k = 0 .. n-2 go through the table (row k, input x[k+1])
    int kd = 0;                                      // leading samples evaluated directly (the direct loop's steps)
    if (T0 < (double)TAB_EARLY) kd = min(n_pick, ((int)((double)TAB_EARLY - T0) + 3) & ~3);
    if (pred[8 * s + 3] != T0) kd = n_pick;          // a table built for another sample counter: nothing of it is used
    const int n_rows = n_pick - kd;
    const int n_chunks = (n_rows + SPEC_SG - 1) / SPEC_SG;
    auto issue = [&](int cc) {                       // chunk cc -> slot cc % SPEC_SLOTS; slot 0 also refreshes the mirror (lane 0)
        const int slot = cc % SPEC_SLOTS;
        const float4* src = rows + RQ * ((long long)kd + (long long)cc * SPEC_SG);
        constexpr int CB = SPEC_SG * 16 * RQ, MB = 32 * 16 * RQ;
        if (slot == 0) {
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(tab_smem_u32(&bars[0])), "n"(CB + MB) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(tab_smem_u32(ring)), "l"(src), "n"(CB), "r"(tab_smem_u32(&bars[0])) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(tab_smem_u32(ring + RQ * SPEC_R)), "l"(src), "n"(MB), "r"(tab_smem_u32(&bars[0])) : "memory");
        } else tab_bulk_load<CB>(ring + RQ * slot * SPEC_SG, src, &bars[slot]);
    };
    if (lane == 0) {
#pragma unroll
        for (int i = 0; i < SPEC_SLOTS; i++) tab_mbar_init(&bars[i]);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        for (int i = 0; i < SPEC_SLOTS && i < n_chunks; i++) issue(i);
    }
    __syncwarp();
    // first sample: the carried feedbackI/Q are whatever the caller holds, so the detector is libm's
    if (kd == 0) dy4_pll_filter(detector_libm(x[0], fbI, fbQ), c.Kp, c.Ki, &integ, &phase);
    else {
        PllRegs rg = {fbI, fbQ, integ, phase, T0, 0.0, 0.0, 0.0};
        dy4_nco_t o;
        o.c = 1.0; o.s = 0.0; o.base_hi = 0.0; o.base_lo = 0.0;
        pll_advance<1, false>(detector_libm(x[0], rg.fbI, rg.fbQ), rg, c, o, n > 1 && x[1] < 0.0f);
        tab_direct_span(x, y, n, 0, kd, &rg, &o, c.w, c.Kp, c.Ki);
        integ = rg.integ; phase = rg.phase;
    }
    int ready = 0;                                    // chunks that have landed (and been waited for)
    int directs = 0, groups = 0;
    bool bailed = false;
    unsigned sm_ring = tab_smem_u32(ring);
    asm volatile("mov.b32 %0, %0;" : "+r"(sm_ring));  // opaque: keeps ptxas from re-deriving the shared window address in the loop
    unsigned mask[B];                                 // lane i: all ones in mask[i]
#pragma unroll
    for (int i = 0; i < B; i++) { const unsigned m = lane == i ? 0xffffffffu : 0u; asm volatile("mov.b32 %0, %1;" : "=r"(mask[i]) : "r"(m)); }
    int r = 0;
    float gi = integ, gp = phase;
    float* yp = y + kd + lane;                        // phaseEst row: lane i writes the sample of step i
#pragma unroll 1
    while (r < n_rows) {
        // rows r .. r+B-1 must have landed.  The first time chunk cc is needed the loop is still inside chunk cc-1 (B < SG), so
        // chunk cc-2 is done with: its slot takes chunk cc+2.
        while (ready < n_chunks && ready * SPEC_SG < r + B) {
            const int cc = ready, slot = cc % SPEC_SLOTS;
            const unsigned parity = (unsigned)((cc / SPEC_SLOTS) & 1);
            while (!tab_mbar_try(&bars[slot], parity)) { }
            ready++;
            if (cc >= 2 && cc + 2 < n_chunks && lane == 0) {
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                issue(cc + 2);
            }
        }
        const int nv = min(B, n_rows - r);
        const unsigned base = sm_ring + 16u * RQ * (unsigned)(r & (SPEC_R - 1));
        const float4 q = spec_lds128(base + 32u * (unsigned)lane + 16u);      // (t, tc_lo, tc_hi, hm) of this lane's step
        unsigned cph = 0u;
        float i0 = gi, p0 = gp;
#pragma unroll
        for (int i = 0; i < B; i++) {
            const float4 A = spec_lds128(base + 32u * i);                                          // (a_lo, a_hi, b_lo, b_hi)
            float T;
            asm volatile("ld.shared.f32 %0, [%1];" : "=f"(T) : "r"(base + 32u * i + 16u));
            cph |= __float_as_uint(p0) & mask[i];
#if DY4_SEL_PACKED
            // both candidates at once: packed float adds (add.rn.f32x2 rounds each half like add.rn.f32)
            float i_lo, i_hi, t_lo, t_hi;
            asm("{\n.reg .b64 ii, axy, azw, ic, tc;\nmov.b64 ii, {%4, %4};\nmov.b64 axy, {%5, %6};\nmov.b64 azw, {%7, %8};\n"
                "add.rn.f32x2 ic, ii, axy;\nadd.rn.f32x2 tc, azw, ic;\nmov.b64 {%0, %1}, ic;\nmov.b64 {%2, %3}, tc;\n}\n"
                : "=f"(i_lo), "=f"(i_hi), "=f"(t_lo), "=f"(t_hi) : "f"(i0), "f"(A.x), "f"(A.y), "f"(A.z), "f"(A.w));
            const float p_lo = __fadd_rn(p0, t_lo), p_hi = __fadd_rn(p0, t_hi);
#else
            const float i_lo = __fadd_rn(i0, A.x), i_hi = __fadd_rn(i0, A.y);                          // filter.cpp:207, both candidates
            const float p_lo = __fadd_rn(p0, __fadd_rn(A.z, i_lo)), p_hi = __fadd_rn(p0, __fadd_rn(A.w, i_hi));     // :210
#endif
            const bool up = p0 > T;                                                                    // trigArg rounds to the upper candidate
            i0 = up ? i_hi : i_lo;
            p0 = up ? p_hi : p_lo;
        }
        const float myph = __uint_as_float(cph);
        const float tc = myph > q.x ? q.z : q.y, hm = q.w;
        const unsigned bad = ~__ballot_sync(0xffffffffu, (lane >= nv) | (dy4_spec_fast_check(myph, tc, hm) != 0));
        groups++;
        if (__builtin_expect(bad == 0u && nv == B, 1)) {                       // every pick certain: straight on
            *yp = myph;
            yp += B; r += B; gi = i0; gp = p0;
            continue;
        }
        const int j = bad ? __ffs(bad) - 1 : nv;                               // steps 0 .. j-1 are certain
        if (lane < j) *yp = myph;
        // (integ, phaseEst) before step j: phaseEst is parked in lane j; integ is replayed from the group's start (rare path)
        const float pj = __uint_as_float(__shfl_sync(0xffffffffu, cph, j & 31));
        const float ij = spec_replay_integ(base, j, gi, gp);
        if (j == nv) { yp += j; r += j; gi = (j == B) ? i0 : ij; gp = (j == B) ? p0 : pj; continue; }      // the launch's last rows (nv < B)
        // step j directly, then on from behind it
        const int k = r + j;
        const float2 nx = spec_direct_step(x, n, kd + k, T0, c.w, c.Kp, c.Ki, ij, pj);
        yp += j;
        if (lane == 0) *yp = pj;
        gi = nx.x; gp = nx.y; r = k + 1; yp += 1;
        directs++;
        // a stream whose loop is not in lock lands here all the time: hand the rest of the launch to the direct loop
        if (r >= 2 * SPEC_SG && 4 * directs > r) { bailed = true; break; }
    }
    integ = gi; phase = gp;
    // no bulk copy may be in flight when the CTA retires: wait for every chunk that was issued and not consumed
    {
        const int issued = min(n_chunks, max(SPEC_SLOTS, ready + 2));
        for (int cc = ready; cc < issued; cc++) {
            const unsigned parity = (unsigned)((cc / SPEC_SLOTS) & 1);
            while (!tab_mbar_try(&bars[cc % SPEC_SLOTS], parity)) { }
        }
    }
    if (bailed && kd + r < n_pick) {                                          // the rest directly, from state_k
        const int k = kd + r;
        const float th = dy4_pll_trigarg(c.w, dy4_pll_count(T0, k + 1), phase);
        dy4_nco_t o;
        dy4_sincos_nco_v((double)th, x[min(k + 1, n - 1)] < 0.0f, &o, 0);
        PllRegs rg = {__double2float_rn(o.c), __double2float_rn(o.s), integ, phase, dy4_pll_count(T0, k + 1), 0.0, 0.0, 0.0};
        tab_direct_span(x, y, n, k, n_pick, &rg, &o, c.w, c.Kp, c.Ki);
        integ = rg.integ; phase = rg.phase;
    }
    // last sample of the launch: trigArg and feedbackI/Q directly (they are carried to the next launch)
    const float th = dy4_pll_trigarg(c.w, dy4_pll_count(T0, n), phase);
    dy4_nco_t o;
    dy4_sincos_nco_v((double)th, 0, &o, 0);
    const float nco_next = nco_value(th, c.ncoScale, c.phaseAdjust);          // nco_state for the next launch (filter.cpp:218-219)
    if (lane == 0) {
        y[n - 1] = phase;
        nco0[s] = nco_carry;                         // nco_state opens this launch's NCO row (filter.cpp:184)
        tstart[s] = (float)T0;                       // k_nco_phase needs the sample counter this launch started from
        st[0] = __double2float_rn(o.c); st[1] = __double2float_rn(o.s); st[2] = integ; st[3] = phase;
        st[4] = (float)dy4_pll_count(T0, n);
        st[5] = nco_next;
        // how many whole turns the prediction of this launch ended away from the true phaseEst (see k_pll_predict)
        need[s] = pred[8 * s + 4] + rint(((double)phase - pred[8 * s + 1]) * 0.15915494309189533577);
        if (stats) { atomicAdd(stats + 0, n_rows); atomicAdd(stats + 1, groups); atomicAdd(stats + 2, directs); atomicAdd(stats + 3, directs); }
    }
}

// NCO row from the phaseEst row of k_pll_sel: trigArg[k-1] = RN_f(RN_d(w*T) + phase[k-1]) (filter.cpp:214), then as k_nco
__global__ void __launch_bounds__(256)
k_nco_phase(const float* __restrict__ phase, long long phase_stride, const float* __restrict__ nco0, const float* __restrict__ tstart,
            float* __restrict__ nco, long long nco_stride, int n, PllConst c)
{
    const int k = blockIdx.y * blockDim.x + threadIdx.x;
    const int s = blockIdx.x;
    if (k >= n) return;
    float v;
    if (k == 0) v = nco0[s];
    else v = nco_value(dy4_pll_trigarg(c.w, dy4_pll_count((double)tstart[s], k), __ldg(phase + (long long)s * phase_stride + k - 1)), c.ncoScale, c.phaseAdjust);
    nco[(long long)s * nco_stride + k] = v;
}

// NCO row from the phase row: nco[0] = carried nco_state, nco[k] = cos(trigArg[k-1]*ncoScale + phaseAdjust)
// (filter.cpp:184,219-221).  Not part of the recurrence, so it runs as a plain data-parallel pass.
__global__ void __launch_bounds__(256)
k_nco(const double* __restrict__ theta, long long theta_stride, const float* __restrict__ nco0,
      float* __restrict__ nco, long long nco_stride, int n, float ncoScale, float phaseAdjust)
{
    const int k = blockIdx.y * blockDim.x + threadIdx.x;
    const int s = blockIdx.x;                       // streams on grid.x (no 65535 limit)
    if (k >= n) return;
    float v;
    if (k == 0) v = nco0[s];
    else v = nco_value((float)__ldg(theta + (long long)s * theta_stride + k - 1), ncoScale, phaseAdjust);
    nco[(long long)s * nco_stride + k] = v;
}

// Reciprocal of every PLL input sample, in double (dy4_recip: exact-rounded float reciprocal + one Newton step).
// The detector divides by the input sample; the inputs are known before the recurrence runs, so this is a plain
// data-parallel pre-pass.  0 marks samples the fast detector does not cover (0, denormal, huge, inf, NaN).
__global__ void __launch_bounds__(256)
k_pll_prep(const float* __restrict__ in, long long in_stride, double* __restrict__ inv, long long inv_stride, int n)
{
    // four samples per thread: one 16-byte load, two 16-byte stores (rows are 16-byte aligned, see dy4_pipeline.cu)
    const int k = 4 * (blockIdx.y * blockDim.x + threadIdx.x);
    if (k >= n) return;
    const float* src = in + (long long)blockIdx.x * in_stride + k;
    double* dst = inv + (long long)blockIdx.x * inv_stride + k;
    if (k + 4 <= n) {
        const float4 x = __ldg(reinterpret_cast<const float4*>(src));
        reinterpret_cast<double2*>(dst)[0] = make_double2(fast_ok(x.x) ? dy4_recip(x.x) : 0.0, fast_ok(x.y) ? dy4_recip(x.y) : 0.0);
        reinterpret_cast<double2*>(dst)[1] = make_double2(fast_ok(x.z) ? dy4_recip(x.z) : 0.0, fast_ok(x.w) ? dy4_recip(x.w) : 0.0);
    } else {
        for (int j = 0; k + j < n; j++) { const float x = __ldg(src + j); dst[j] = fast_ok(x) ? dy4_recip(x) : 0.0; }
    }
}

bool g_spec_stats_on = std::getenv("DY4_PLL_STATS") != nullptr;

}  // namespace

// development counters of the speculative loop: rows, iterations, flips, direct steps since the last call
extern "C" int dy4_debug_pll_stats(long long* out4)
{
    int h[4] = {0, 0, 0, 0}, z[4] = {0, 0, 0, 0};
    if (cudaMemcpyFromSymbol(h, g_spec_stats, sizeof(h)) != cudaSuccess) return -1;
    cudaMemcpyToSymbol(g_spec_stats, z, sizeof(z));
    for (int i = 0; i < 4; i++) out4[i] = h[i];
    return 0;
}

cudaError_t dy4_launch_pll(const Dy4PllArgs& a, cudaStream_t st) { return dy4_launch_pll_parts(a, st, DY4_PLL_PREP | DY4_PLL_LOOP | DY4_PLL_NCO); }

// The three passes can be queued on different streams: only the serial loop belongs on the PLL stream, its
// data-parallel pre-pass (reciprocals) and post-pass (NCO row) ride with the FIR kernels (dy4_pipeline.cu).
cudaError_t dy4_launch_pll_parts(const Dy4PllArgs& a, cudaStream_t st, int parts)
{
    if (a.n <= 0 || a.n_streams <= 0) return cudaSuccess;
    PllConst c;
    const float Cp = 2.666f, Ci = 3.555f;                       // filter.cpp:175-176
    c.Kp = a.normBandwidth * Cp;                                // :178
    const float bw2 = a.normBandwidth * a.normBandwidth;
    c.Ki = bw2 * Ci;                                            // :179
    const float ratio = a.freq / a.Fs;                          // float divide inside :214
    c.w = 2 * 3.14159265358979323846 * (double)ratio;           // (2*PI)*(freq/Fs), left to right in double
    c.ncoScale = a.ncoScale;
    c.phaseAdjust = a.phaseAdjust;
    // Table-driven loop (dy4_plltab.h) when the caller provides the row buffer: predict -> table -> serial pick.
    if (a.tab && a.fresh > 0 && a.n > a.fresh + 64) {
        // First launch of a stream.  While the loop acquires lock the detector crosses +-pi, where one ulp decides the
        // sign of a 2*pi jump: the predictor cannot know on which turn phaseEst settles, and a table built around the
        // wrong turn never matches.  So the first `fresh` samples run through the direct loop (k_pll), and prediction,
        // table and picks start from the exact state it leaves, on the rest of the launch.
        const int E = a.fresh;                               // a multiple of 4: every row offset stays 16-byte aligned
        Dy4PllArgs A = a, B = a;
        A.tab = nullptr; A.n = E; A.fresh = 0;
        B.in = a.in + E; B.nco = a.nco + E; B.theta = a.theta + E; B.inv = a.inv + E; B.n = a.n - E; B.fresh = 0; B.pred_carry = 0;
        B.nco0 = a.nco0b;
        cudaError_t e = cudaSuccess;
        if (parts & DY4_PLL_PREP) e = dy4_launch_pll_parts(A, st, DY4_PLL_PREP);
        if (e == cudaSuccess && (parts & DY4_PLL_LOOP)) {
            e = dy4_launch_pll_parts(A, st, DY4_PLL_LOOP);
            if (e == cudaSuccess) e = dy4_launch_pll_parts(B, st, DY4_PLL_PREP | DY4_PLL_LOOP);
        }
        if (e == cudaSuccess && (parts & DY4_PLL_NCO)) {
            e = dy4_launch_pll_parts(A, st, DY4_PLL_NCO);
            if (e == cudaSuccess) e = dy4_launch_pll_parts(B, st, DY4_PLL_NCO);
        }
        return e;
    }
    if (a.tab) {
        if (parts & DY4_PLL_PREP) {                          // time-parallel: may run beside the serial loop of the previous launch
            const int nseg = (a.n + PRED_SEG - 1) / PRED_SEG;
            k_pll_predict<<<dim3(a.n_streams, (nseg + 127) / 128), 128, 0, st>>>(a.in, a.in_stride, a.state, a.pred_in, a.pred_out, a.need, a.pred_carry,
                                                                                 reinterpret_cast<float*>(a.theta), 2 * a.wide_stride, a.n, c);
            k_pll_table_ops<<<dim3(a.n_streams, (a.n + 127) / 128), 128, 0, st>>>(a.in, a.in_stride, a.pred_out, reinterpret_cast<const float*>(a.theta), 2 * a.wide_stride, a.tab, a.tab_stride, a.risk, a.n, c);
            g_dy4_launches += 2;
            cudaError_t e = cudaGetLastError();
            if (e != cudaSuccess) return e;
        }
        if (parts & DY4_PLL_LOOP) {
            float* ph = reinterpret_cast<float*>(a.inv);     // phaseEst row (the reciprocal row of the direct loop is free in this mode)
            int* stats = nullptr;
            if (g_spec_stats_on) cudaGetSymbolAddress(reinterpret_cast<void**>(&stats), g_spec_stats);
            k_pll_sel<32><<<a.n_streams, 32, 0, st>>>(a.in, a.in_stride, a.tab, a.tab_stride, ph, 2 * a.wide_stride, a.nco0, a.tstart, a.pred_out, a.need, a.state,
                                                      stats, a.n, a.n_streams, c);
            g_dy4_launches++;
            cudaError_t e = cudaGetLastError();
            if (e != cudaSuccess) return e;
        }
        if (parts & DY4_PLL_NCO) {
            dim3 grid(a.n_streams, (a.n + 255) / 256);
            k_nco_phase<<<grid, 256, 0, st>>>(reinterpret_cast<const float*>(a.inv), 2 * a.wide_stride, a.nco0, a.tstart, a.nco, a.nco_stride, a.n, c);
            g_dy4_launches++;
        }
        return cudaGetLastError();
    }
    if (parts & DY4_PLL_PREP) {
        dim3 gp(a.n_streams, ((a.n + 3) / 4 + 255) / 256);
        k_pll_prep<<<gp, 256, 0, st>>>(a.in, a.in_stride, a.inv, a.wide_stride, a.n);
        g_dy4_launches++;
        cudaError_t e0 = cudaGetLastError();
        if (e0 != cudaSuccess) return e0;
    }
    if (parts & DY4_PLL_LOOP) {
        k_pll<2, false><<<(a.n_streams + 31) / 32, 32, 0, st>>>(a.in, a.in_stride, a.inv, a.theta, a.wide_stride, a.nco0, a.state, a.n, a.n_streams, c);
        g_dy4_launches++;
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return e;
    }
    if (parts & DY4_PLL_NCO) {
        dim3 grid(a.n_streams, (a.n + 255) / 256);
        k_nco<<<grid, 256, 0, st>>>(a.theta, a.wide_stride, a.nco0, a.nco, a.nco_stride, a.n, a.ncoScale, a.phaseAdjust);
        g_dy4_launches++;
    }
    return cudaGetLastError();
}
