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
// Memory: each lane walks its own row; loads are issued four samples ahead as
// one 16-byte load and NCO values leave as 16-byte stores, both off the
// dependent chain that bounds this kernel (FP64 libm latency, not bandwidth).
#include "dy4_common.cuh"
#include "dy4_kernels.h"
#include "dy4_internal.h"

namespace {

struct PllConst { double w; float Kp, Ki, ncoScale, phaseAdjust; };

struct PllRegs { float fbI, fbQ, integ, phase, trigOffset; };

__device__ __forceinline__ float pll_step(float x, PllRegs& s, const PllConst& c)
{
    const float eI = __fmul_rn((x == 0.0f ? 1.0f : x), s.fbI);                 // filter.cpp:192
    const float eQ = __fmul_rn(x, -s.fbQ);                                     // :193
    const float eD = __double2float_rn(atan2((double)eQ, (double)eI));         // :200
    s.integ = __fadd_rn(s.integ, __fmul_rn(c.Ki, eD));                         // :207
    s.phase = __fadd_rn(s.phase, __fadd_rn(__fmul_rn(c.Kp, eD), s.integ));     // :210
    s.trigOffset = __fadd_rn(s.trigOffset, 1.0f);                              // :213
    const float trigArg = __double2float_rn(__dadd_rn(__dmul_rn(c.w, (double)s.trigOffset), (double)s.phase)); // :214
    double sn, cs;
    sincos((double)trigArg, &sn, &cs);
    s.fbI = __double2float_rn(cs);                                             // :216
    s.fbQ = __double2float_rn(sn);                                             // :217
    const float narg = __fadd_rn(__fmul_rn(trigArg, c.ncoScale), c.phaseAdjust);
    return __double2float_rn(cos((double)narg));                               // :219/:221
}

__global__ void __launch_bounds__(32)
k_pll(const float* __restrict__ in, long long in_stride, float* __restrict__ nco, long long nco_stride,
      float* __restrict__ state, int n, int n_streams, PllConst c)
{
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n_streams) return;
    float* st = state + (long long)s * 8;
    PllRegs r = {st[0], st[1], st[2], st[3], st[4]};
    float pend = st[5];                          // nco_state: the value that opens the next block (filter.cpp:184)
    const float* x = in + (long long)s * in_stride;
    float* y = nco + (long long)s * nco_stride;
    const int n4 = n & ~3;
    float4 v = n4 > 0 ? *reinterpret_cast<const float4*>(x) : make_float4(0, 0, 0, 0);
    for (int k = 0; k < n4; k += 4) {
        const float4 cur = v;
        if (k + 4 < n4) v = *reinterpret_cast<const float4*>(x + k + 4);
        float4 o;
        o.x = pend;
        o.y = pll_step(cur.x, r, c);
        o.z = pll_step(cur.y, r, c);
        o.w = pll_step(cur.z, r, c);
        pend = pll_step(cur.w, r, c);
        *reinterpret_cast<float4*>(y + k) = o;
    }
    for (int k = n4; k < n; k++) { y[k] = pend; pend = pll_step(x[k], r, c); }
    st[0] = r.fbI; st[1] = r.fbQ; st[2] = r.integ; st[3] = r.phase; st[4] = r.trigOffset; st[5] = pend;
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
    k_pll<<<(a.n_streams + threads - 1) / threads, threads, 0, st>>>(a.in, a.in_stride, a.nco, a.nco_stride, a.state, a.n, a.n_streams, c);
    g_dy4_launches++;
    return cudaGetLastError();
}
