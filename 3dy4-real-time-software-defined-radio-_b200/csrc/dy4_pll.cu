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
constexpr int PRED_SEG = 256;      // samples per predictor thread
constexpr int PRED_WARM = 1024;    // warm-up steps before a segment: the loop forgets its state as 0.98657^k (1e-6 after 1024)

// 1. predicted trigArg of every sample (double) -> theta row.  One thread per (stream, segment); the first WARM samples
// of a launch start from the exact carried state, later segments from that state as a guess plus the warm-up.
// `pred_in` / `pred_out`: [n_streams][8] doubles (integ, phase, sample counter after / before the launch, turns) — the
// predictor's own state at the start / end of the launch.  carry == 0: the launch starts from the exact carried PLL
// state (the loop of the previous launch has finished).  carry != 0: from where the previous launch's prediction ended,
// so that prediction and table of sub-chunk c+1 do not wait for the serial loop of sub-chunk c (dy4_pipeline.cu).
// The loop locks to the pilot modulo 2*pi, and after a loss of lock the predictor may settle a whole number of turns
// away from the true phaseEst — a table built there never matches.  So every serial launch reports how many turns its
// prediction was off (`need`, k_pll_tab), and with carry == 2 the launch two later shifts its start by that much.
__global__ void __launch_bounds__(128)
k_pll_predict(const float* __restrict__ in, long long in_stride, const float* __restrict__ state, const double* __restrict__ pred_in,
              double* __restrict__ pred_out, const double* __restrict__ need, int carry, double* __restrict__ th_hat, long long wide_stride,
              int n, PllConst c)
{
    const int s = blockIdx.x;
    const int s0 = (blockIdx.y * blockDim.x + threadIdx.x) * PRED_SEG;
    if (s0 >= n) return;
    const float* st = state + (long long)s * 8;
    const float* x = in + (long long)s * in_stride;
    double* y = th_hat + (long long)s * wide_stride;
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
        const double t0 = th_prev = dy4_pred_step(v.x, th_prev, __dmul_rn(c.w, dy4_pll_count(T0, k + 1)), Kp, Ki, &integ, &phase);
        const double t1 = th_prev = dy4_pred_step(v.y, th_prev, __dmul_rn(c.w, dy4_pll_count(T0, k + 2)), Kp, Ki, &integ, &phase);
        const double t2 = th_prev = dy4_pred_step(v.z, th_prev, __dmul_rn(c.w, dy4_pll_count(T0, k + 3)), Kp, Ki, &integ, &phase);
        const double t3 = th_prev = dy4_pred_step(v.w, th_prev, __dmul_rn(c.w, dy4_pll_count(T0, k + 4)), Kp, Ki, &integ, &phase);
        if (k >= s0) {
            *reinterpret_cast<double2*>(y + k) = make_double2(t0, t1);
            *reinterpret_cast<double2*>(y + k + 2) = make_double2(t2, t3);
        }
    }
    for (; k < k1; k++) {
        th_prev = dy4_pred_step(__ldg(x + k), th_prev, __dmul_rn(c.w, dy4_pll_count(T0, k + 1)), Kp, Ki, &integ, &phase);
        if (k >= s0) y[k] = th_prev;
    }
    if (k1 == n) {                                                  // the thread of the last segment: state for the next launch
        pred_out[8 * s] = integ; pred_out[8 * s + 1] = phase; pred_out[8 * s + 2] = dy4_pll_count(T0, n); pred_out[8 * s + 3] = T0;
        pred_out[8 * s + 4] = turns;
    }
}

// 2. one table row per sample: the exact errorD of the next step for the two float grid points the prediction lies between
__global__ void __launch_bounds__(128)
k_pll_table(const float* __restrict__ in, long long in_stride, const double* __restrict__ pred_out,
            const double* __restrict__ th_hat, long long wide_stride, float4* __restrict__ tab, long long tab_stride, int n, PllConst c)
{
    const int s = blockIdx.x;
    const int k = blockIdx.y * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const double T0 = pred_out[8 * s + 3];                          // sample counter at the start of this launch (k_pll_predict)
    const float* x = in + (long long)s * in_stride;
    dy4_tabrow_t r;
    dy4_tab_make_row(__ldg(th_hat + (long long)s * wide_stride + k), __dmul_rn(c.w, dy4_pll_count(T0, k + 1)),
                     k + 1 < n ? __ldg(x + k + 1) : 0.0f, k + 1 < n, T0 + (double)k < (double)DY4_TAB_EARLY, c.Kp, c.Ki, &r);
    float4* o = tab + (long long)s * tab_stride + 2 * (long long)k;
    o[0] = make_float4(r.t, r.hu, r.hm, 0.0f);
    o[1] = make_float4(r.a_lo, r.a_hi, r.b_lo, r.b_hi);
}

// 3. the serial loop.  One lane per stream, `lanes` streams per warp (few: a direct evaluation stalls the whole warp).
// Measured on a B200 (tools/ubench_pick.cu): a dependent FADD 4.9 cycles, FSETP -> FSEL 8.9 (both at half issue rate) —
// and ~55 cycles for every BRANCH a lone warp executes.  So:
//  * table rows arrive through a per-lane shared-memory ring, ONE bulk asynchronous copy (cp.async.bulk, completing on
//    the lane's own mbarrier) per super-group of 64 samples, three super-groups ahead: no waiting on global memory and
//    no issue slots spent on it;
//  * a super-group is straight-line code: per sample two speculative loop-filter updates (float adds of the table's
//    precomputed products), ONE compare of phaseEst with the row's threshold, two selects, and a three-instruction
//    guard — with ONE branch per super-group on "every pick was certain"; if not, the super-group is redone step by
//    step from its saved state (tab_redo, out of line).
// Output: phaseEst after every sample (float); trigArg and the NCO follow from it elementwise in k_nco_phase.
constexpr int TAB_LANES = 4;                       // most streams per warp
constexpr int TAB_ROW_Q = 2;                       // 16-byte words per row
// template parameters of k_pll_tab: TAB_SG samples per super-group (64; 32 / 128 through the A/B knob DY4_PLL_SG), TAB_SLOTS ring slots
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

// state_k -> state_{k+1} with no branch.  q0 = (t, u/2, hm, -), q1 = (a_lo, a_hi, b_lo, b_hi).
// `ok` stays true while every pick was certain; if not, integ/phase are garbage and the caller redoes the super-group.
// On the serial chain: one compare of phaseEst with the threshold and one select (8.9 cycles, tools/ubench_pick.cu);
// the two speculative updates and the three-instruction guard run beside it.
__device__ __forceinline__ void tab_step_spec(const float4 q0, const float4 q1, float& integ, float& phase, bool& ok)
{
    const float i_lo = __fadd_rn(integ, q1.x), i_hi = __fadd_rn(integ, q1.y);
    const float p_lo = __fadd_rn(phase, __fadd_rn(q1.z, i_lo));
    const float p_hi = __fadd_rn(phase, __fadd_rn(q1.w, i_hi));
    const bool up = phase > q0.x;
    const float v = __fadd_rn(fabsf(__fadd_rn(phase, -q0.x)), -q0.y);
    ok = ok && (fabsf(v) < q0.z);
    integ = up ? i_hi : i_lo;
    phase = up ? p_hi : p_lo;
}

// A super-group again, carefully: a pick where it is certain, else that step directly (dy4_pllmath.h), as k_pll does.
// Out of line and looping: it runs for a fraction of a percent of the super-groups once the loop is in lock.
__device__ __noinline__ int tab_redo(const float4* src, const float* x_next, float* y, double T0, int k, int count, double w, float Kp, float Ki,
                                     float* integ_io, float* phase_io)
{
    float integ = *integ_io, phase = *phase_io;
    int n_direct = 0;
#pragma unroll 1
    for (int r = 0; r < count; r++) {
        const float4 q0 = src[2 * r], q1 = src[2 * r + 1];
        y[r] = phase;
        int up;
        if (dy4_tab_pick(phase, q0.x, q0.y, q0.z, &up))
            dy4_pll_filter_ab(up ? q1.y : q1.x, up ? q1.w : q1.z, &integ, &phase);
        else {
            dy4_pll_filter(dy4_next_errorD((double)dy4_pll_trigarg(w, dy4_pll_count(T0, k + r + 1), phase), x_next[r]), Kp, Ki, &integ, &phase);
            n_direct++;
        }
    }
    *integ_io = integ; *phase_io = phase;
    return n_direct;                                          // steps that had to be evaluated directly
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

template <bool FENCE, int TAB_SG, int TAB_SLOTS>
__global__ void __launch_bounds__(32)
k_pll_tab(const float* __restrict__ in, long long in_stride, const float4* __restrict__ tab, long long tab_stride,
          float* __restrict__ phase_out, long long phase_stride, float* __restrict__ nco0, float* __restrict__ tstart,
          const double* __restrict__ pred, double* __restrict__ need, float* __restrict__ state, int n, int n_streams, PllConst c, int lanes)
{
    constexpr int TAB_SG_BYTES = TAB_SG * TAB_ROW_Q * 16;
    constexpr int TAB_LANE_Q = TAB_SG * TAB_ROW_Q + 1;   // lane stride in 16-byte words: +1 spreads the lanes over the banks
    __shared__ __align__(16) float4 ring[TAB_SLOTS * TAB_LANES * TAB_LANE_Q];
    __shared__ __align__(8) unsigned long long bars[TAB_SLOTS * TAB_LANES];
    const int lane = threadIdx.x;
    const int s = blockIdx.x * lanes + lane;
    if (lane >= lanes || s >= n_streams || n <= 0) return;
    float* st = state + (long long)s * 8;
    float fbI = st[0], fbQ = st[1], integ = st[2], phase = st[3];
    const double T0 = (double)st[4];
    nco0[s] = st[5];                                 // nco_state opens this launch's NCO row (filter.cpp:184)
    tstart[s] = st[4];                               // k_nco_phase needs the sample counter this launch started from
    const float* x = in + (long long)s * in_stride;
    const float4* rows = tab + (long long)s * tab_stride;
    float* y = phase_out + (long long)s * phase_stride;
    const int n_pick = n - 1;                        // steps k = 0 .. n-2 go through the table (row k, input x[k+1])
    // Start of a stream: while the loop acquires lock the detector crosses +-pi, where one ulp decides the sign of a
    // 2*pi jump — nothing predicts that, so those samples are evaluated directly (as k_pll does) and the table starts
    // after them.  k_pll_table leaves their rows empty.
    int kd = 0;
    if (T0 < (double)TAB_EARLY) kd = min(n_pick, ((int)((double)TAB_EARLY - T0) + 3) & ~3);
    // A pick certifies "trigArg = RN_f(RN_d(w*T) + phaseEst) is this grid point" for the sample counter T the TABLE was
    // built with: it must be this stream's own counter.  It always is when the launches are ordered as dy4_pipeline.cu
    // orders them; if it ever is not, nothing of the table is used.
    if (pred[8 * s + 3] != T0) kd = n_pick;
    const int n_sg = (n_pick - kd) / TAB_SG;         // whole super-groups after the direct part
#pragma unroll
    for (int i = 0; i < TAB_SLOTS; i++) tab_mbar_init(&bars[i * TAB_LANES + lane]);   // each lane owns its barriers: no CTA-wide sync
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    auto issue = [&](int i) {                        // super-group i -> slot i % TAB_SLOTS
        if (i < n_sg) {
            const int slot = i % TAB_SLOTS;
            tab_bulk_load<TAB_SG_BYTES>(ring + (slot * TAB_LANES + lane) * TAB_LANE_Q, rows + TAB_ROW_Q * ((long long)kd + (long long)i * TAB_SG), &bars[slot * TAB_LANES + lane]);
        }
    };
    for (int i = 0; i < TAB_SLOTS - 1; i++) issue(i);
    // first sample: the carried feedbackI/Q are whatever the caller holds, so the detector is libm's
    if (kd == 0) dy4_pll_filter(detector_libm(x[0], fbI, fbQ), c.Kp, c.Ki, &integ, &phase);
    else {
        // direct part (tab_direct_span), samples 0 .. kd; it ends holding state_kd
        PllRegs r = {fbI, fbQ, integ, phase, T0, 0.0, 0.0, 0.0};
        dy4_nco_t o;
        o.c = 1.0; o.s = 0.0; o.base_hi = 0.0; o.base_lo = 0.0;
        pll_advance<1, false>(detector_libm(x[0], r.fbI, r.fbQ), r, c, o, n > 1 && x[1] < 0.0f);
        tab_direct_span(x, y, n, 0, kd, &r, &o, c.w, c.Kp, c.Ki);
        integ = r.integ; phase = r.phase;
    }
    int directs = 0, sg_done = n_sg;
    bool bailed = false;
#pragma unroll 1
    for (int i = 0; i < n_sg; i++) {
        const int slot = i % TAB_SLOTS;
        // slot (i-1) % TAB_SLOTS was read in the previous trip: order those reads before the async write that refills it.
        // (Every value read from it has been consumed by then — `ok` depends on all of them — so DY4_PLL_FENCE=0 drops
        // the fence for an A/B measurement; the default keeps it.)
        if (FENCE) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        issue(i + TAB_SLOTS - 1);
        unsigned long long* bar = &bars[slot * TAB_LANES + lane];
        const unsigned parity = (unsigned)((i / TAB_SLOTS) & 1);
        if (!tab_mbar_try(bar, parity)) { while (!tab_mbar_try(bar, parity)) { } }     // super-group i has landed (normally long ago)
        const float4* src = ring + (slot * TAB_LANES + lane) * TAB_LANE_Q;
        const int k = kd + i * TAB_SG;
        float si = integ, sp = phase;
        bool ok = true;
#pragma unroll
        for (int r = 0; r < TAB_SG; r += 4) {
            float ph[4];
#pragma unroll
            for (int q = 0; q < 4; q++) {
                ph[q] = sp;
                tab_step_spec(src[2 * (r + q)], src[2 * (r + q) + 1], si, sp, ok);
            }
            *reinterpret_cast<float4*>(y + k + r) = make_float4(ph[0], ph[1], ph[2], ph[3]);   // (rewritten by tab_redo if a pick was not certain)
        }
        if (!ok) {
            si = integ; sp = phase;
            directs += tab_redo(src, x + k + 1, y + k, T0, k, TAB_SG, c.w, c.Kp, c.Ki, &si, &sp);
            // A stream whose loop is not in lock (no pilot, noise) misses all the time, and a direct evaluation inside the
            // redo costs four times a step of the direct loop: once more than a quarter of the samples so far had to be
            // evaluated directly, the rest of this launch goes through the direct loop.
            if (i >= 2 && 4 * directs > (i + 1) * TAB_SG) { integ = si; phase = sp; sg_done = i + 1; bailed = true; break; }
        }
        integ = si; phase = sp;
    }
    {
        const int k = kd + sg_done * TAB_SG;
        if (bailed && k < n_pick) {                              // the rest directly, from state_k
            const float th = dy4_pll_trigarg(c.w, dy4_pll_count(T0, k + 1), phase);
            dy4_nco_t o;
            dy4_sincos_nco_v((double)th, x[min(k + 1, n - 1)] < 0.0f, &o, 0);
            PllRegs r = {__double2float_rn(o.c), __double2float_rn(o.s), integ, phase, dy4_pll_count(T0, k + 1), 0.0, 0.0, 0.0};
            tab_direct_span(x, y, n, k, n_pick, &r, &o, c.w, c.Kp, c.Ki);
            integ = r.integ; phase = r.phase;
        } else if (k < n_pick) {
            // tail (fewer than TAB_SG samples): its rows are first copied into ring slot 0 with all loads in flight at once —
            // read one by one from global memory inside the step loop they would cost a DRAM latency per sample
            float4* dst = ring + lane * TAB_LANE_Q;
            const float4* src = rows + TAB_ROW_Q * (long long)k;
            const int nq = TAB_ROW_Q * (n_pick - k);
            for (int q = 0; q < nq; q += 8) {
                float4 v[8];
#pragma unroll
                for (int j = 0; j < 8; j++) v[j] = __ldg(src + min(q + j, nq - 1));
#pragma unroll
                for (int j = 0; j < 8; j++) if (q + j < nq) dst[q + j] = v[j];
            }
            tab_redo(dst, x + k + 1, y + k, T0, k, n_pick - k, c.w, c.Kp, c.Ki, &integ, &phase);
        }
    }
    // last sample of the launch: trigArg and feedbackI/Q directly (they are carried to the next launch)
    y[n - 1] = phase;
    const float th = dy4_pll_trigarg(c.w, dy4_pll_count(T0, n), phase);
    dy4_nco_t o;
    dy4_sincos_nco_v((double)th, 0, &o, 0);
    st[0] = __double2float_rn(o.c); st[1] = __double2float_rn(o.s); st[2] = integ; st[3] = phase;
    st[4] = (float)dy4_pll_count(T0, n);
    st[5] = nco_value(th, c.ncoScale, c.phaseAdjust);          // nco_state for the next launch (filter.cpp:218-219)
    // how many whole turns the prediction of this launch ended away from the true phaseEst (see k_pll_predict)
    need[s] = pred[8 * s + 4] + rint(((double)phase - pred[8 * s + 1]) * 0.15915494309189533577);
}

// ====================================================================================================================
// Speculative serial loop (dy4_plltab.h §2b): 16-byte rows (predicted candidate, other candidate, double threshold).
// One WARP per stream.  The dependent chain of a step is three float adds of the predicted candidate's products — no
// compare, no select: the predictor names the right candidate ~96 % of the time.  All 32 lanes run the same chain;
// G steps are straight-line code, every step parks (integ, phaseEst) in shared memory, and afterwards lane i certifies
// step i (dy4_spec_check, in double, off the chain).  A ballot finds the first step that is not certainly the predicted
// candidate: the loop resumes THERE, from the parked state, with that step forced to the other candidate (certified the
// same way); if neither candidate is certain the step is evaluated directly (dy4_pllmath.h), so the result is the
// reference's by construction, as in k_pll_tab.  Rows arrive by one bulk asynchronous copy per 64 samples into a ring
// of 4 slots; the lanes turn each landed chunk into the chain's operands (Ki*e, Kp*e of both candidates) in one pass.
// ====================================================================================================================
constexpr int SPEC_SG = 128;                         // rows per bulk copy
constexpr int SPEC_SLOTS = 4;
constexpr int SPEC_R = SPEC_SG * SPEC_SLOTS;         // ring size in rows

__global__ void __launch_bounds__(128)
k_pll_table16(const float* __restrict__ in, long long in_stride, const double* __restrict__ pred_out,
              const double* __restrict__ th_hat, long long wide_stride, dy4_row16_t* __restrict__ tab, long long tab_stride, int n, PllConst c)
{
    const int s = blockIdx.x;
    const int k = blockIdx.y * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const double T0 = pred_out[8 * s + 3];                          // sample counter at the start of this launch (k_pll_predict)
    const float* x = in + (long long)s * in_stride;
    dy4_row16_t r;
    dy4_tab_make_row16(__ldg(th_hat + (long long)s * wide_stride + k), __dmul_rn(c.w, dy4_pll_count(T0, k + 1)),
                       k + 1 < n ? __ldg(x + k + 1) : 0.0f, k + 1 < n, T0 + (double)k < (double)DY4_TAB_EARLY, &r, nullptr, nullptr);
    float4 v;
    v.x = __int_as_float(__double2loint(r.t)); v.y = __int_as_float(__double2hiint(r.t)); v.z = r.e_p; v.w = r.e_o;
    reinterpret_cast<float4*>(tab + (long long)s * tab_stride)[k] = v;
}

// One step at row k, carefully: dy4_spec_check (double) for the predicted, then the other candidate; if neither is
// certain the step is evaluated directly (filter.cpp:192-214 through dy4_pllmath.h).  Returns (integ, phase); *direct = 1
// if it came to that.  Out of line: it runs where the float certificate of the loop is too coarse (the first ~0.1 s of
// a stream), at binade edges of trigArg, and where the loop is not in lock.
__device__ __noinline__ float2 spec_slow_step(double t, float e_p, float e_o, const float* __restrict__ x, int n, int k, double T0, double w, float Kp, float Ki,
                                              float integ, float phase, int* direct)
{
    float eD;
    *direct = 0;
    if (dy4_spec_check(phase, t, 0)) eD = e_p;
    else if (dy4_spec_check(phase, t, 1)) eD = e_o;
    else {
        eD = dy4_next_errorD((double)dy4_pll_trigarg(w, dy4_pll_count(T0, k + 1), phase), x[min(k + 1, n - 1)]);
        *direct = 1;
    }
    dy4_pll_filter(eD, Kp, Ki, &integ, &phase);
    return make_float2(integ, phase);
}

__device__ int g_spec_stats[4];      // rows, groups run, flips, direct steps (development counters, DY4_PLL_STATS=1)

// Per-warp state of the speculative loop.  `r`: first row not yet run (rows before it are run, the last group of them
// possibly not yet certified); (integ, phase): state before row r.  The group [pr, pr + pn) is run but not certified.
struct SpecLoop {
    int r, forced;                   // forced: row r must take the OTHER candidate (it was certainly not the predicted one)
    int pr, pn, pforced;             // pending group: first row, rows, whether its first step took the other candidate
    float integ, phase;
};

// One trip: certify the pending group (its states are in cap_prev) while the chain of the next group runs from the
// state the pending group ended in.  ab: the chain's operands (Ki e, Kp e) of rows r .. r+G-1, already in registers;
// abn receives those of rows r+G .. r+2G-1 for the next trip.  Returns the number of certain steps of the pending group
// (== L.pn: all of them; the caller then makes the new group the pending one).
template <int G>
__device__ __forceinline__ int spec_trip(const SpecLoop& L, const float4* __restrict__ conv, const float4* __restrict__ chk,
                                         const float2* __restrict__ cap_prev, float2* __restrict__ cap_cur, const float2 (&ab)[G], float2 (&abn)[G],
                                         float* __restrict__ y_prev, int lane, float& si, float& sp)
{
    // certificate of the pending group, lane i for step i (float form; dy4_spec_fast_row)
    const float myph = cap_prev[min(lane, G)].y;
    const float4 q = chk[(L.pr + lane) & (SPEC_R - 1)];
    const float tc = (lane == 0 && L.pforced) ? q.z : q.x;
    const bool good = (lane >= L.pn) | (dy4_spec_fast_check(myph, tc, q.y) != 0);
    // operands of the group after this one (optimistic: this group will be found certain)
    const float4* cvn = conv + ((L.r + G) & (SPEC_R - 1));
#pragma unroll
    for (int i = 0; i < G; i++) abn[i] = *reinterpret_cast<const float2*>(cvn + i);
    // the chain: three dependent float adds per step (filter.cpp:207,210), state parked before every step
    si = L.integ; sp = L.phase;
#pragma unroll
    for (int i = 0; i < G; i++) {
        cap_cur[i] = make_float2(si, sp);
        si = __fadd_rn(si, ab[i].x);
        sp = __fadd_rn(sp, __fadd_rn(ab[i].y, si));
    }
    cap_cur[G] = make_float2(si, sp);
    const unsigned bad = ~__ballot_sync(0xffffffffu, good);
    const int j = bad ? __ffs(bad) - 1 : L.pn;
    if (lane < j) y_prev[lane] = myph;                 // phaseEst after each certain sample (k_nco_phase turns it into the NCO row)
    return j;
}

template <int G>
__global__ void __launch_bounds__(32)
k_pll_spec(const float* __restrict__ in, long long in_stride, const dy4_row16_t* __restrict__ tab, long long tab_stride,
           float* __restrict__ phase_out, long long phase_stride, float* __restrict__ nco0, float* __restrict__ tstart,
           const double* __restrict__ pred, double* __restrict__ need, float* __restrict__ state, int* __restrict__ stats, int n, int n_streams, PllConst c)
{
    static_assert(G >= 4 && G <= 32 && 3 * G <= SPEC_SG, "one lane certifies one step; three groups must fit a chunk");
    __shared__ __align__(16) dy4_row16_t raw[SPEC_R];
    __shared__ __align__(16) float4 conv[SPEC_R + 32];            // (Ki e_p, Kp e_p, Ki e_o, Kp e_o); the last 32 mirror the first 32
    __shared__ __align__(16) float4 chk[SPEC_R];                  // (tc_p, hm, tc_o, -), dy4_spec_fast_row
    __shared__ __align__(8) float2 cap0[G + 1], cap1[G + 1];      // (integ, phaseEst) before every step of a group, double-buffered
    __shared__ __align__(8) unsigned long long bars[SPEC_SLOTS];
    const int lane = threadIdx.x;
    const int s = blockIdx.x;
    if (s >= n_streams || n <= 0) return;
    float* st = state + (long long)s * 8;
    float fbI = st[0], fbQ = st[1], integ = st[2], phase = st[3];
    const double T0 = (double)st[4];
    const float nco_carry = st[5];
    const float* x = in + (long long)s * in_stride;
    const dy4_row16_t* rows = tab + (long long)s * tab_stride;
    float* y = phase_out + (long long)s * phase_stride;
    const int n_pick = n - 1;                        // steps k = 0 .. n-2 go through the table (row k, input x[k+1])
    int kd = 0;                                      // leading samples evaluated directly (see k_pll_tab)
    if (T0 < (double)TAB_EARLY) kd = min(n_pick, ((int)((double)TAB_EARLY - T0) + 3) & ~3);
    if (pred[8 * s + 3] != T0) kd = n_pick;          // a table built for another sample counter: nothing of it is used
    const int n_rows = n_pick - kd;
    const int n_chunks = (n_rows + SPEC_SG - 1) / SPEC_SG;
    auto issue = [&](int cc) {                       // chunk cc -> slot cc % SPEC_SLOTS (lane 0)
        const int slot = cc % SPEC_SLOTS;
        tab_bulk_load<SPEC_SG * 16>(raw + slot * SPEC_SG, rows + (long long)kd + (long long)cc * SPEC_SG, &bars[slot]);
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
    int conv_chunks = 0, directs = 0, trips = 0, flips = 0;
    bool bailed = false;
    // rows up to `upto` (exclusive) must be in conv / chk: turn the chunks that have landed into the loop's operands
    // and keep two more chunks in flight.  Chunk cc-2 is consumed by then: the oldest row the loop can come back to is
    // the pending group's first, r - G, and conversion of chunk cc is asked for when r + 2G > cc*SG, i.e. r - G > (cc-1)*SG.
    auto cover = [&](int upto) {
        while (conv_chunks < n_chunks && conv_chunks * SPEC_SG < upto) {
            const int cc = conv_chunks, slot = cc % SPEC_SLOTS;
            const unsigned parity = (unsigned)((cc / SPEC_SLOTS) & 1);
            while (!tab_mbar_try(&bars[slot], parity)) { }
#pragma unroll
            for (int p = 0; p < SPEC_SG / 32; p++) {
                const int idx = slot * SPEC_SG + p * 32 + lane;
                const dy4_row16_t row = raw[idx];
                const float4 q = make_float4(__fmul_rn(c.Ki, row.e_p), __fmul_rn(c.Kp, row.e_p), __fmul_rn(c.Ki, row.e_o), __fmul_rn(c.Kp, row.e_o));
                conv[idx] = q;
                if (idx < 32) conv[SPEC_R + idx] = q;
                float4 v;
                dy4_spec_fast_row(row.t, &v.x, &v.y, &v.z);
                v.w = 0.0f;
                chk[idx] = v;
            }
            conv_chunks++;
            __syncwarp();
            if (cc >= 2 && cc + 2 < n_chunks && lane == 0) {
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                issue(cc + 2);
            }
        }
    };
    SpecLoop L;
    L.r = 0; L.forced = 0; L.pr = 0; L.pn = 0; L.pforced = 0; L.integ = integ; L.phase = phase;
    float2 ab0[G], ab1[G];
    // (re)start at row L.r from (L.integ, L.phase): nothing pending, operands of rows r .. r+G-1 into ab0
    auto restart = [&]() {
        cover(L.r + 2 * G);
        L.pn = 0; L.pr = L.r; L.pforced = 0;
        const float4* cv = conv + (L.r & (SPEC_R - 1));
        const float4 q0 = cv[0];
        ab0[0] = L.forced ? make_float2(q0.z, q0.w) : make_float2(q0.x, q0.y);
#pragma unroll
        for (int i = 1; i < G; i++) ab0[i] = *reinterpret_cast<const float2*>(cv + i);
    };
    // the pending group failed at its step j (< L.pn): resume there
    auto recover = [&](int j, const float2* cap_prev) {
        const int k = L.pr + j;                                  // row that is not certainly the predicted candidate
        const float2 sj = cap_prev[j];
        if (j == 0 && L.pforced) {                               // ... and not certainly the other one either (float certificate)
            const dy4_row16_t row = raw[k & (SPEC_R - 1)];
            int direct;
            const float2 nx = spec_slow_step(row.t, row.e_p, row.e_o, x, n, kd + k, T0, c.w, c.Kp, c.Ki, sj.x, sj.y, &direct);
            if (lane == 0) y[kd + k] = sj.y;
            L.integ = nx.x; L.phase = nx.y; L.r = k + 1; L.forced = 0;
            directs += direct;
        } else {
            // a row with no usable threshold (NaN) goes straight to the careful step
            L.integ = sj.x; L.phase = sj.y; L.r = k; L.forced = 1;
            flips++;
        }
    };
    restart();
#pragma unroll 1
    while (L.r < n_rows || L.pn > 0) {
        float si, sp;
        int j;
        // ---- even trip: pending states in cap1, this group's into cap0, operands ab0 -> prefetch ab1
        cover(L.r + 2 * G);
        j = spec_trip<G>(L, conv, chk, cap1, cap0, ab0, ab1, y + kd + L.pr, lane, si, sp);
        trips++;
        if (j < L.pn) {
            recover(j, cap1);
            if (L.r >= 2 * SPEC_SG && 4 * directs > L.r) { bailed = true; L.pn = 0; break; }
            restart();
            continue;
        }
        {
            const int nv = max(0, min(G, n_rows - L.r));
            L.pr = L.r; L.pn = nv; L.pforced = L.forced; L.forced = 0;
            if (nv < G) { const float2 e = cap0[nv]; si = e.x; sp = e.y; }   // the launch's last rows: state after the last valid one
            L.r += nv; L.integ = si; L.phase = sp;
        }
        // ---- odd trip: roles of the buffers swapped
        cover(L.r + 2 * G);
        j = spec_trip<G>(L, conv, chk, cap0, cap1, ab1, ab0, y + kd + L.pr, lane, si, sp);
        trips++;
        if (j < L.pn) {
            recover(j, cap0);
            if (L.r >= 2 * SPEC_SG && 4 * directs > L.r) { bailed = true; L.pn = 0; break; }
            restart();
            continue;
        }
        {
            const int nv = max(0, min(G, n_rows - L.r));
            L.pr = L.r; L.pn = nv; L.pforced = L.forced; L.forced = 0;
            if (nv < G) { const float2 e = cap1[nv]; si = e.x; sp = e.y; }
            L.r += nv; L.integ = si; L.phase = sp;
        }
    }
    integ = L.integ; phase = L.phase;
    // no bulk copy may be in flight when the CTA retires: wait for every chunk that was issued and not consumed
    {
        const int issued = min(n_chunks, max(SPEC_SLOTS, conv_chunks + 2));
        for (int cc = conv_chunks; cc < issued; cc++) {
            const unsigned parity = (unsigned)((cc / SPEC_SLOTS) & 1);
            while (!tab_mbar_try(&bars[cc % SPEC_SLOTS], parity)) { }
        }
    }
    if (bailed && kd + L.r < n_pick) {                                        // the rest directly, from state_k
        const int k = kd + L.r;
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
        if (stats) { atomicAdd(stats + 0, n_rows); atomicAdd(stats + 1, trips); atomicAdd(stats + 2, flips); atomicAdd(stats + 3, directs); }
    }
}

// Variant B of the speculative loop: the group is certified right after its own chain (nothing runs ahead on an
// uncertified state, so a flip wastes only the steps behind it).  Lane i catches phaseEst before step i in a register
// on the way (one LOP3 with a lane mask per step), so the certificate is one subtract, one compare and a vote after the
// last step; (integ, phaseEst) of every step are also parked in shared memory for the resume.
template <int G>
__global__ void __launch_bounds__(32)
k_pll_spec2(const float* __restrict__ in, long long in_stride, const dy4_row16_t* __restrict__ tab, long long tab_stride,
            float* __restrict__ phase_out, long long phase_stride, float* __restrict__ nco0, float* __restrict__ tstart,
            const double* __restrict__ pred, double* __restrict__ need, float* __restrict__ state, int* __restrict__ stats, int n, int n_streams, PllConst c)
{
    static_assert(G >= 4 && G <= 32 && 3 * G <= SPEC_SG, "one lane certifies one step");
    __shared__ __align__(16) dy4_row16_t raw[SPEC_R];
    __shared__ __align__(16) float4 conv[SPEC_R + 32];
    __shared__ __align__(16) float4 chk[SPEC_R];
    __shared__ __align__(16) float2 cap[G + 2];
    __shared__ __align__(8) unsigned long long bars[SPEC_SLOTS];
    const int lane = threadIdx.x;
    const int s = blockIdx.x;
    if (s >= n_streams || n <= 0) return;
    float* st = state + (long long)s * 8;
    float fbI = st[0], fbQ = st[1], integ = st[2], phase = st[3];
    const double T0 = (double)st[4];
    const float nco_carry = st[5];
    const float* x = in + (long long)s * in_stride;
    const dy4_row16_t* rows = tab + (long long)s * tab_stride;
    float* y = phase_out + (long long)s * phase_stride;
    const int n_pick = n - 1;
    int kd = 0;
    if (T0 < (double)TAB_EARLY) kd = min(n_pick, ((int)((double)TAB_EARLY - T0) + 3) & ~3);
    if (pred[8 * s + 3] != T0) kd = n_pick;
    const int n_rows = n_pick - kd;
    const int n_chunks = (n_rows + SPEC_SG - 1) / SPEC_SG;
    auto issue = [&](int cc) {
        const int slot = cc % SPEC_SLOTS;
        tab_bulk_load<SPEC_SG * 16>(raw + slot * SPEC_SG, rows + (long long)kd + (long long)cc * SPEC_SG, &bars[slot]);
    };
    if (lane == 0) {
#pragma unroll
        for (int i = 0; i < SPEC_SLOTS; i++) tab_mbar_init(&bars[i]);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        for (int i = 0; i < SPEC_SLOTS && i < n_chunks; i++) issue(i);
    }
    __syncwarp();
    if (kd == 0) dy4_pll_filter(detector_libm(x[0], fbI, fbQ), c.Kp, c.Ki, &integ, &phase);
    else {
        PllRegs rg = {fbI, fbQ, integ, phase, T0, 0.0, 0.0, 0.0};
        dy4_nco_t o;
        o.c = 1.0; o.s = 0.0; o.base_hi = 0.0; o.base_lo = 0.0;
        pll_advance<1, false>(detector_libm(x[0], rg.fbI, rg.fbQ), rg, c, o, n > 1 && x[1] < 0.0f);
        tab_direct_span(x, y, n, 0, kd, &rg, &o, c.w, c.Kp, c.Ki);
        integ = rg.integ; phase = rg.phase;
    }
    int conv_chunks = 0, directs = 0, trips = 0, flips = 0;
    bool bailed = false;
    auto cover = [&](int upto) {
        while (conv_chunks < n_chunks && conv_chunks * SPEC_SG < upto) {
            const int cc = conv_chunks, slot = cc % SPEC_SLOTS;
            const unsigned parity = (unsigned)((cc / SPEC_SLOTS) & 1);
            while (!tab_mbar_try(&bars[slot], parity)) { }
#pragma unroll
            for (int p = 0; p < SPEC_SG / 32; p++) {
                const int idx = slot * SPEC_SG + p * 32 + lane;
                const dy4_row16_t row = raw[idx];
                const float4 q = make_float4(__fmul_rn(c.Ki, row.e_p), __fmul_rn(c.Kp, row.e_p), __fmul_rn(c.Ki, row.e_o), __fmul_rn(c.Kp, row.e_o));
                conv[idx] = q;
                if (idx < 32) conv[SPEC_R + idx] = q;
                float4 v;
                dy4_spec_fast_row(row.t, &v.x, &v.y, &v.z);
                v.w = 0.0f;
                chk[idx] = v;
            }
            conv_chunks++;
            __syncwarp();
            if (cc >= 2 && cc + 2 < n_chunks && lane == 0) {
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                issue(cc + 2);
            }
        }
    };
    unsigned mask[G];                                 // lane i: all ones in mask[i]
#pragma unroll
    for (int i = 0; i < G; i++) { const unsigned m = lane == i ? 0xffffffffu : 0u; asm volatile("mov.b32 %0, %1;" : "=r"(mask[i]) : "r"(m)); }
    int r = 0, forced = 0;
    float2 ab[G], abn[G];
    auto load_ops = [&](float2 (&dst)[G], int row, int f) {
        const float4* cv = conv + (row & (SPEC_R - 1));
        const float4 q0 = cv[0];
        dst[0] = f ? make_float2(q0.z, q0.w) : make_float2(q0.x, q0.y);
#pragma unroll
        for (int i = 1; i < G; i++) dst[i] = *reinterpret_cast<const float2*>(cv + i);
    };
    cover(2 * G);
    load_ops(ab, 0, 0);
#pragma unroll 1
    while (r < n_rows) {
        cover(r + 2 * G);
        const int nv = min(G, n_rows - r);
        const float4 q = chk[(r + lane) & (SPEC_R - 1)];
        const float tc = (lane == 0 && forced) ? q.z : q.x;
        load_ops(abn, r + G, 0);                      // operands of the next group, should this one hold
        float si = integ, sp = phase;
        unsigned mine = 0;
#pragma unroll
        for (int i = 0; i < G; i++) {
            cap[i] = make_float2(si, sp);
            mine |= __float_as_uint(sp) & mask[i];
            si = __fadd_rn(si, ab[i].x);
            sp = __fadd_rn(sp, __fadd_rn(ab[i].y, si));
        }
        cap[G] = make_float2(si, sp);
        const float myph = __uint_as_float(mine);
        const bool good = (lane >= nv) | (dy4_spec_fast_check(myph, tc, q.y) != 0);
        const unsigned bad = ~__ballot_sync(0xffffffffu, good);
        const int j = bad ? __ffs(bad) - 1 : nv;
        if (lane < j) y[kd + r + lane] = myph;
        trips++;
        if (j == G) {                                 // the whole group is certain: straight on
            r += G; integ = si; phase = sp; forced = 0;
#pragma unroll
            for (int i = 0; i < G; i++) ab[i] = abn[i];
            continue;
        }
        const float2 sj = cap[j];
        if (j == nv) { r += nv; integ = sj.x; phase = sj.y; break; }       // the launch's last rows
        const int k = r + j;
        if (j == 0 && forced) {
            const dy4_row16_t row = raw[k & (SPEC_R - 1)];
            int direct;
            const float2 nx = spec_slow_step(row.t, row.e_p, row.e_o, x, n, kd + k, T0, c.w, c.Kp, c.Ki, sj.x, sj.y, &direct);
            if (lane == 0) y[kd + k] = sj.y;
            integ = nx.x; phase = nx.y; r = k + 1; forced = 0;
            directs += direct;
            if (r >= 2 * SPEC_SG && 4 * directs > r) { bailed = true; break; }
        } else { integ = sj.x; phase = sj.y; r = k; forced = 1; flips++; }
        cover(r + 2 * G);
        load_ops(ab, r, forced);
    }
    {
        const int issued = min(n_chunks, max(SPEC_SLOTS, conv_chunks + 2));
        for (int cc = conv_chunks; cc < issued; cc++) {
            const unsigned parity = (unsigned)((cc / SPEC_SLOTS) & 1);
            while (!tab_mbar_try(&bars[cc % SPEC_SLOTS], parity)) { }
        }
    }
    if (bailed && kd + r < n_pick) {
        const int k = kd + r;
        const float th = dy4_pll_trigarg(c.w, dy4_pll_count(T0, k + 1), phase);
        dy4_nco_t o;
        dy4_sincos_nco_v((double)th, x[min(k + 1, n - 1)] < 0.0f, &o, 0);
        PllRegs rg = {__double2float_rn(o.c), __double2float_rn(o.s), integ, phase, dy4_pll_count(T0, k + 1), 0.0, 0.0, 0.0};
        tab_direct_span(x, y, n, k, n_pick, &rg, &o, c.w, c.Kp, c.Ki);
        integ = rg.integ; phase = rg.phase;
    }
    const float th = dy4_pll_trigarg(c.w, dy4_pll_count(T0, n), phase);
    dy4_nco_t o;
    dy4_sincos_nco_v((double)th, 0, &o, 0);
    const float nco_next = nco_value(th, c.ncoScale, c.phaseAdjust);
    if (lane == 0) {
        y[n - 1] = phase;
        nco0[s] = nco_carry;
        tstart[s] = (float)T0;
        st[0] = __double2float_rn(o.c); st[1] = __double2float_rn(o.s); st[2] = integ; st[3] = phase;
        st[4] = (float)dy4_pll_count(T0, n);
        st[5] = nco_next;
        need[s] = pred[8 * s + 4] + rint(((double)phase - pred[8 * s + 1]) * 0.15915494309189533577);
        if (stats) { atomicAdd(stats + 0, n_rows); atomicAdd(stats + 1, trips); atomicAdd(stats + 2, flips); atomicAdd(stats + 3, directs); }
    }
}

// NCO row from the phaseEst row of k_pll_tab: trigArg[k-1] = RN_f(RN_d(w*T) + phase[k-1]) (filter.cpp:214), then as k_nco
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
    static const int threads = std::getenv("DY4_PLL_THREADS") ? atoi(std::getenv("DY4_PLL_THREADS")) : 32;   // tuning knob
    // Table-driven loop (dy4_plltab.h) when the caller provides the row buffer: predict -> table -> serial pick.
    static const int tab_lanes_env = std::getenv("DY4_PLL_LANES") ? atoi(std::getenv("DY4_PLL_LANES")) : 0;
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
                                                                                 a.theta, a.wide_stride, a.n, c);
            // DY4_PLL_TABLE_SMEM (bytes of unused dynamic shared memory per CTA) caps the resident CTAs of the table kernel: fewer
            // warps contending with the serial loops it runs beside (A/B knob)
            static const int tab_smem = std::getenv("DY4_PLL_TABLE_SMEM") ? atoi(std::getenv("DY4_PLL_TABLE_SMEM")) : 0;
            if (a.spec) k_pll_table16<<<dim3(a.n_streams, (a.n + 127) / 128), 128, 0, st>>>(a.in, a.in_stride, a.pred_out, a.theta, a.wide_stride,
                                                                                              reinterpret_cast<dy4_row16_t*>(a.tab), a.tab_stride, a.n, c);
            else k_pll_table<<<dim3(a.n_streams, (a.n + 127) / 128), 128, tab_smem, st>>>(a.in, a.in_stride, a.pred_out, a.theta, a.wide_stride, a.tab, a.tab_stride, a.n, c);
            g_dy4_launches += 2;
            cudaError_t e = cudaGetLastError();
            if (e != cudaSuccess) return e;
        }
        if ((parts & DY4_PLL_LOOP) && a.spec) {
            float* ph = reinterpret_cast<float*>(a.inv);     // phaseEst row (the reciprocal row of the direct loop is free in this mode)
            static const int g = std::getenv("DY4_PLL_G") ? atoi(std::getenv("DY4_PLL_G")) : 16;
            int* stats = nullptr;
            if (g_spec_stats_on) cudaGetSymbolAddress(reinterpret_cast<void**>(&stats), g_spec_stats);
#define DY4_SPEC_ARGS a.in, a.in_stride, reinterpret_cast<const dy4_row16_t*>(a.tab), a.tab_stride, ph, 2 * a.wide_stride, a.nco0, a.tstart, a.pred_out, a.need, a.state, stats, a.n, a.n_streams, c
            static const int variant = std::getenv("DY4_PLL_VARIANT") ? atoi(std::getenv("DY4_PLL_VARIANT")) : 2;
            if (variant == 1) {
                if (g <= 8) k_pll_spec<8><<<a.n_streams, 32, 0, st>>>(DY4_SPEC_ARGS);
                else if (g <= 12) k_pll_spec<12><<<a.n_streams, 32, 0, st>>>(DY4_SPEC_ARGS);
                else if (g <= 16) k_pll_spec<16><<<a.n_streams, 32, 0, st>>>(DY4_SPEC_ARGS);
                else if (g <= 24) k_pll_spec<24><<<a.n_streams, 32, 0, st>>>(DY4_SPEC_ARGS);
                else k_pll_spec<32><<<a.n_streams, 32, 0, st>>>(DY4_SPEC_ARGS);
            } else {
                if (g <= 8) k_pll_spec2<8><<<a.n_streams, 32, 0, st>>>(DY4_SPEC_ARGS);
                else if (g <= 12) k_pll_spec2<12><<<a.n_streams, 32, 0, st>>>(DY4_SPEC_ARGS);
                else if (g <= 16) k_pll_spec2<16><<<a.n_streams, 32, 0, st>>>(DY4_SPEC_ARGS);
                else if (g <= 24) k_pll_spec2<24><<<a.n_streams, 32, 0, st>>>(DY4_SPEC_ARGS);
                else k_pll_spec2<32><<<a.n_streams, 32, 0, st>>>(DY4_SPEC_ARGS);
            }
#undef DY4_SPEC_ARGS
            g_dy4_launches++;
            cudaError_t e = cudaGetLastError();
            if (e != cudaSuccess) return e;
        } else if (parts & DY4_PLL_LOOP) {
            int lanes = tab_lanes_env > 0 ? tab_lanes_env : (a.n_streams + 591) / 592;       // one warp per SM sub-partition while they last
            lanes = std::max(1, std::min(lanes, TAB_LANES));
            static const bool fence = !(std::getenv("DY4_PLL_FENCE") && atoi(std::getenv("DY4_PLL_FENCE")) == 0);
            const int grid = (a.n_streams + lanes - 1) / lanes;
            float* ph = reinterpret_cast<float*>(a.inv);     // phaseEst row (the reciprocal row of the direct loop is free in this mode)
            static const int sg = std::getenv("DY4_PLL_SG") ? atoi(std::getenv("DY4_PLL_SG")) : 64;
#define DY4_TAB_ARGS a.in, a.in_stride, a.tab, a.tab_stride, ph, 2 * a.wide_stride, a.nco0, a.tstart, a.pred_out, a.need, a.state, a.n, a.n_streams, c, lanes
            if (sg == 32) k_pll_tab<true, 32, 4><<<grid, 32, 0, st>>>(DY4_TAB_ARGS);
            else if (sg == 128) k_pll_tab<true, 128, 2><<<grid, 32, 0, st>>>(DY4_TAB_ARGS);
            else if (fence) k_pll_tab<true, 64, 4><<<grid, 32, 0, st>>>(DY4_TAB_ARGS);
            else k_pll_tab<false, 64, 4><<<grid, 32, 0, st>>>(DY4_TAB_ARGS);
#undef DY4_TAB_ARGS
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
    // DY4_PLL_NARROW=f2f selects the plain double->float->double narrowing of trigArg instead of the
    // magic-constant rounding inside a tracked binade (A/B knob; results are identical, see tests).
    static const bool f2f = std::getenv("DY4_PLL_NARROW") && std::string(std::getenv("DY4_PLL_NARROW")) == "f2f";
    if (parts & DY4_PLL_LOOP) {
        const dim3 g((a.n_streams + threads - 1) / threads);
        if (f2f) k_pll<0, false><<<g, threads, 0, st>>>(a.in, a.in_stride, a.inv, a.theta, a.wide_stride, a.nco0, a.state, a.n, a.n_streams, c);
        else k_pll<2, false><<<g, threads, 0, st>>>(a.in, a.in_stride, a.inv, a.theta, a.wide_stride, a.nco0, a.state, a.n, a.n_streams, c);
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
