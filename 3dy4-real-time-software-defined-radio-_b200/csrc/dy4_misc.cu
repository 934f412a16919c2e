// dy4_misc.cu — carried-history update and the generic single-op kernels behind the
// filter.h compatibility tier.
//
// k_tails persists, per stream, exactly what the reference carries between blocks
// (RFState/AudioState, src/project.cpp:25-38) in input-history form: the last
// bytes of IQ, the last IF samples and the last mixed (nco*stereo-band*2) samples.
// Because every FIR/demod/delay state in the reference is INPUT history
// (filter.cpp:82,139,169,100-101,239), absolute indexing with these tails
// reproduces its block-carried state for any chunking.
//
// The generic kernels take any tap count / factor (one thread per output, taps
// ascending, unfused float multiply-add like filter.cpp) and exist so that every
// prototype of include/filter.h:17-34 has a CUDA implementation; they are the
// link-compatibility tier, not the throughput path.
#include "dy4_common.cuh"
#include "dy4_kernels.h"
#include "dy4_internal.h"

namespace {

__global__ void k_tails(Dy4TailArgs a)
{
    const int s = blockIdx.x, t = threadIdx.x;
    if (a.iq_tail && a.iq) {
        const uint8_t* row = a.iq + (long long)s * a.row_stride + a.row_bytes - DY4_IQ_TAIL;
        if (t < DY4_IQ_TAIL) a.iq_tail[(long long)s * DY4_IQ_TAIL + t] = row[t];
    }
    if (a.if_tail && a.if_in) {
        const float* row = a.if_in + (long long)s * a.if_stride + a.n_if - DY4_IF_TAIL;
        if (t < DY4_IF_TAIL) a.if_tail[(long long)s * DY4_IF_TAIL + t] = row[t];
    }
    if (a.mix_tail && a.nco) {
        const long long o = (long long)s * a.bb_stride + a.n_if - DY4_MIX_TAIL;
        if (t < DY4_MIX_TAIL) a.mix_tail[(long long)s * DY4_MIX_TAIL + t] = __fmul_rn(__fmul_rn(a.nco[o + t], a.sband[o + t]), 2.0f);
    }
}

// y[m] = sum_{k<nh} h[k] * xe[n_hist + m*step - k], xe = history followed by the block; out-of-range reads are 0
__global__ void k_generic_fir(const float* __restrict__ xe, int n_hist, int n_out, int step, const float* __restrict__ h, int nh, float* __restrict__ y)
{
    const int m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= n_out) return;
    float acc = 0.0f;
    const long long c = (long long)n_hist + (long long)m * step;
    for (int k = 0; k < nh; k++) {
        const long long i = c - k;
        const float xv = i >= 0 ? xe[i] : 0.0f;
        acc = __fadd_rn(acc, __fmul_rn(h[k], xv));
    }
    y[m] = acc;
}

// filter.cpp:158-167 with n = m*down
__global__ void k_generic_resample(const float* __restrict__ xe, int n_hist, int n_out, int up, int down, const float* __restrict__ h, int nh, float* __restrict__ y)
{
    const int m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= n_out) return;
    const long long n = (long long)m * down;
    const int phase = (int)(n % up);
    float acc = 0.0f;
    for (long long k = phase; k < nh; k += up) {
        const long long d = n - k;                         // multiple of up
        const long long i = n_hist + (d >= 0 ? d / up : -((k - n) / up));
        const float xv = i >= 0 ? xe[i] : 0.0f;
        acc = __fadd_rn(acc, __fmul_rn(h[k], xv));
    }
    y[m] = acc;
}

__global__ void k_generic_demod(const float* __restrict__ I, const float* __restrict__ Q, int n, float prev_I, float prev_Q, float* __restrict__ out)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const float i = I[k], q = Q[k];
    const float pi = k > 0 ? I[k - 1] : prev_I, pq = k > 0 ? Q[k - 1] : prev_Q;
    const float den = __double2float_rn(fma((double)i, (double)i, (double)q * (double)q));
    const float num = __fsub_rn(__fmul_rn(i, __fsub_rn(q, pq)), __fmul_rn(q, __fsub_rn(i, pi)));
    out[k] = den == 0.0f ? 0.0f : __fdiv_rn(num, den);
}

__global__ void k_u8_to_float(const uint8_t* __restrict__ raw, long long n, float* __restrict__ out)
{
    const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n) out[k] = __fmul_rn((float)((int)raw[k] - 128), 0.0078125f);   // exact: k/128
}

__global__ void k_pointwise(int op, const float* __restrict__ a, const float* __restrict__ b, int n, float* __restrict__ out)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    if (op == 0) out[k] = __fmul_rn(__fmul_rn(a[k], b[k]), 2.0f);          // filter.cpp:264
    else if (op == 1) out[k] = __fadd_rn(a[k], b[k]);                       // :276
    else out[k] = __fsub_rn(a[k], b[k]);                                    // :288
}

inline int cdiv(long long a, int b) { return (int)((a + b - 1) / b); }

}  // namespace

cudaError_t dy4_launch_tails(const Dy4TailArgs& a, cudaStream_t st)
{
    if (a.n_streams <= 0) return cudaSuccess;
    k_tails<<<a.n_streams, 256, 0, st>>>(a);
    g_dy4_launches++;
    return cudaGetLastError();
}
cudaError_t dy4_launch_generic_fir(const float* xe, int n_hist, int n_out, int step, const float* h, int nh, float* y, cudaStream_t st)
{
    if (n_out <= 0) return cudaSuccess;
    k_generic_fir<<<cdiv(n_out, 128), 128, 0, st>>>(xe, n_hist, n_out, step, h, nh, y);
    g_dy4_launches++;
    return cudaGetLastError();
}
cudaError_t dy4_launch_generic_resample(const float* xe, int n_hist, int n_out, int up, int down, const float* h, int nh, float* y, cudaStream_t st)
{
    if (n_out <= 0) return cudaSuccess;
    k_generic_resample<<<cdiv(n_out, 128), 128, 0, st>>>(xe, n_hist, n_out, up, down, h, nh, y);
    g_dy4_launches++;
    return cudaGetLastError();
}
cudaError_t dy4_launch_generic_demod(const float* I, const float* Q, int n, float prev_I, float prev_Q, float* out, cudaStream_t st)
{
    if (n <= 0) return cudaSuccess;
    k_generic_demod<<<cdiv(n, 128), 128, 0, st>>>(I, Q, n, prev_I, prev_Q, out);
    g_dy4_launches++;
    return cudaGetLastError();
}
cudaError_t dy4_launch_u8_to_float(const uint8_t* raw, long long n, float* out, cudaStream_t st)
{
    if (n <= 0) return cudaSuccess;
    k_u8_to_float<<<cdiv(n, 256), 256, 0, st>>>(raw, n, out);
    g_dy4_launches++;
    return cudaGetLastError();
}
cudaError_t dy4_launch_pointwise(int op, const float* a, const float* b, int n, float* out, cudaStream_t st)
{
    if (n <= 0) return cudaSuccess;
    k_pointwise<<<cdiv(n, 256), 256, 0, st>>>(op, a, b, n, out);
    g_dy4_launches++;
    return cudaGetLastError();
}
