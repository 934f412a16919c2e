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

namespace {

struct PllConst { double w; float Kp, Ki, ncoScale, phaseAdjust; };

struct PllRegs { float fbI, fbQ, integ, phase, trigOffset; };

// Everything after the phase detector: loop filter, phase accumulator, NCO (filter.cpp:206-222).
// `o` receives sin/cos of the new trigArg together with its reduction, for the next detector call.
__device__ __forceinline__ float nco_value(float trigArg, float ncoScale, float phaseAdjust)
{
    const float narg = __fadd_rn(__fmul_rn(trigArg, ncoScale), phaseAdjust);   // float, as filter.cpp:219/:221
    dy4_nco_t on;
    dy4_sincos_nco((double)narg, &on);
    return __double2float_rn(on.c);
}

// returns the new trigArg (the NCO output is a pure function of it and is evaluated by k_nco)
__device__ __forceinline__ float pll_advance(float eD, PllRegs& s, const PllConst& c, dy4_nco_t& o)
{
    s.integ = __fadd_rn(s.integ, __fmul_rn(c.Ki, eD));                         // :207
    s.phase = __fadd_rn(s.phase, __fadd_rn(__fmul_rn(c.Kp, eD), s.integ));     // :210
    s.trigOffset = __fadd_rn(s.trigOffset, 1.0f);                              // :213
    const float trigArg = __double2float_rn(__dadd_rn(__dmul_rn(c.w, (double)s.trigOffset), (double)s.phase)); // :214
    dy4_sincos_nco((double)trigArg, &o);
    s.fbI = __double2float_rn(o.c);                                            // :216
    s.fbQ = __double2float_rn(o.s);                                            // :217
    return trigArg;
}

// One PLL step with the libm phase detector: used for the first sample of a launch (the carried
// feedbackI/Q come from the caller) and for inputs the fast detector does not cover (0, denormal, inf, NaN).
__device__ __noinline__ float detector_libm(float x, const PllRegs& s)
{
    const float eI = __fmul_rn((x == 0.0f ? 1.0f : x), s.fbI);                 // filter.cpp:192
    const float eQ = __fmul_rn(x, -s.fbQ);                                     // :193
    return __double2float_rn(atan2((double)eQ, (double)eI));                   // :200
}

__device__ __forceinline__ float pll_step(float x, double inv_x, PllRegs& s, const PllConst& c, dy4_nco_t& o, bool have_o)
{
    float eD;
    const float ax = fabsf(x);
    if (have_o && ax > 1e-20f && ax < 1e20f) {
        const float eI = __fmul_rn(x, s.fbI);
        const float eQ = __fmul_rn(x, -s.fbQ);
        eD = __double2float_rn(dy4_detector_atan2((double)eQ, (double)eI, x < 0.0f ? 1.0 : 0.0, &o, inv_x));
    } else {
        eD = detector_libm(x, s);
    }
    return pll_advance(eD, s, c, o);
}

__global__ void __launch_bounds__(32)
k_pll(const float* __restrict__ in, long long in_stride, float* __restrict__ theta, long long theta_stride,
      float* __restrict__ nco0, float* __restrict__ state, int n, int n_streams, PllConst c)
{
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n_streams) return;
    float* st = state + (long long)s * 8;
    PllRegs r = {st[0], st[1], st[2], st[3], st[4]};
    nco0[s] = st[5];                             // nco_state opens this launch's NCO row (filter.cpp:184)
    const float* x = in + (long long)s * in_stride;
    float* y = theta + (long long)s * theta_stride;
    dy4_nco_t o;
    o.c = 1.0; o.s = 0.0; o.rho_hi = 0.0; o.rho_lo = 0.0; o.n = 0;
    bool have_o = false;
    float last = 0.0f;
    const int n4 = n & ~3;
    float4 v = n4 > 0 ? *reinterpret_cast<const float4*>(x) : make_float4(0, 0, 0, 0);
    for (int k = 0; k < n4; k += 4) {
        const float4 cur = v;
        if (k + 4 < n4) v = *reinterpret_cast<const float4*>(x + k + 4);
        // reciprocals of the inputs: known ahead of the recurrence, so off its critical path
        const double i0 = dy4_recip(cur.x), i1 = dy4_recip(cur.y), i2 = dy4_recip(cur.z), i3 = dy4_recip(cur.w);
        float4 out;
        out.x = pll_step(cur.x, i0, r, c, o, have_o);
        have_o = true;
        out.y = pll_step(cur.y, i1, r, c, o, true);
        out.z = pll_step(cur.z, i2, r, c, o, true);
        out.w = pll_step(cur.w, i3, r, c, o, true);
        last = out.w;
        *reinterpret_cast<float4*>(y + k) = out;
    }
    for (int k = n4; k < n; k++) {
        last = pll_step(x[k], dy4_recip(x[k]), r, c, o, have_o);
        y[k] = last;
        have_o = true;
    }
    st[0] = r.fbI; st[1] = r.fbQ; st[2] = r.integ; st[3] = r.phase; st[4] = r.trigOffset;
    st[5] = nco_value(last, c.ncoScale, c.phaseAdjust);      // nco_state for the next launch (filter.cpp:218-219)
}

// NCO row from the phase row: nco[0] = carried nco_state, nco[k] = cos(trigArg[k-1]*ncoScale + phaseAdjust)
// (filter.cpp:184,219-221).  Not part of the recurrence, so it runs as a plain data-parallel pass.
__global__ void __launch_bounds__(256)
k_nco(const float* __restrict__ theta, long long theta_stride, const float* __restrict__ nco0,
      float* __restrict__ nco, long long nco_stride, int n, float ncoScale, float phaseAdjust)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    const int s = blockIdx.y;
    if (k >= n) return;
    float v;
    if (k == 0) v = nco0[s];
    else v = nco_value(__ldg(theta + (long long)s * theta_stride + k - 1), ncoScale, phaseAdjust);
    nco[(long long)s * nco_stride + k] = v;
}

}  // namespace

cudaError_t dy4_launch_pll(const Dy4PllArgs& a, cudaStream_t st)
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
    const int threads = 32;
    k_pll<<<(a.n_streams + threads - 1) / threads, threads, 0, st>>>(a.in, a.in_stride, a.theta, a.theta_stride, a.nco0, a.state, a.n, a.n_streams, c);
    g_dy4_launches++;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    dim3 grid((a.n + 255) / 256, a.n_streams);
    k_nco<<<grid, 256, 0, st>>>(a.theta, a.theta_stride, a.nco0, a.nco, a.nco_stride, a.n, a.ncoScale, a.phaseAdjust);
    g_dy4_launches++;
    return cudaGetLastError();
}
