// dy4_bpf.cu — pilot (18.5-19.5 kHz) and stereo-band (22-54 kHz) band-pass filters
// computed together from ONE staged tile of IF samples.
//
// Replaces the two blockConvolveFIR calls of backend(), src/project.cpp:120-121
// (src/filter.cpp:66-83): y[n] = sum_k h[k] x[n-k], 101 taps, full rate.
// The two filters share the input, so the pair (pilot, stereo-band) rides in one
// packed f32x2 accumulator: x is staged in shared memory duplicated as (x,x) and
// multiplies the constant-bank pair (h_pilot[k], h_stereo[k]).  R consecutive
// outputs per thread; the input index descends so taps are applied in ascending
// k for every output.  EXACT arithmetic (dy4_common.cuh): the pilot output feeds
// the PLL and must be bit-identical to the reference.
#include "dy4_common.cuh"
#include "dy4_kernels.h"
#include "dy4_internal.h"

namespace {

__constant__ TapPairs c_bpf2[6];   // (pilot[k], stereo-band[k]) pairs per mode; [4],[5]: RDS band-pass and RDS carrier band-pass, (h,h)

// TWOSTREAMS (the single-filter launches of the RDS path, (h,h) tap tables): the two halves of the packed registers
// carry two STREAMS (2b, 2b+1) through the same filter instead of one stream through two filters, so no lane is wasted;
// both outputs go to `pilot` rows, `sband` is unused.
template <int R, int NT, bool EXACT, bool SQUARE, bool TWOSTREAMS>
__global__ void __launch_bounds__(NT)
k_twin_bpf(const float* __restrict__ if_in, long long if_stride, const float* __restrict__ if_tail, long long tail_stride,
           float* __restrict__ pilot, float* __restrict__ sband, long long out_stride, int n_if,
           u64 nz, int mode, int n_streams)
{
    constexpr int T = NT * R;
    constexpr int HALO = 128;                       // >= 100, keeps 16-byte alignment of global loads
    constexpr int NP = T + HALO;                    // logical (x,x) pairs per tile
    __shared__ __align__(16) float2 sm[NP + 2 * (NP / R) + 8];
    const int tid = threadIdx.x;
    const int n0 = blockIdx.y * T;                    // streams on grid.x (no 65535 limit), tiles on grid.y
    const int sa = TWOSTREAMS ? 2 * blockIdx.x : blockIdx.x;
    const int sb = TWOSTREAMS ? min(sa + 1, n_streams - 1) : sa;            // odd stream count: the last CTA filters its stream twice
    const float* row = if_in + (long long)sa * if_stride;
    const float* tail = if_tail + (long long)sa * tail_stride;
    const float* row_b = if_in + (long long)sb * if_stride;
    const float* tail_b = if_tail + (long long)sb * tail_stride;

    // stage 4 samples per step as 4 duplicated pairs; logical pair p <-> sample n0 - HALO + p
    for (int u = tid; u < NP / 4; u += NT) {
        const int i = n0 - HALO + 4 * u;            // multiple of 4, never straddles 0
        float4 v;
        if (i < 0) v = *reinterpret_cast<const float4*>(tail + DY4_IF_TAIL + i);
        else if (i + 3 < n_if) v = __ldg(reinterpret_cast<const float4*>(row + i));
        else { v.x = i < n_if ? row[i] : 0.f; v.y = i + 1 < n_if ? row[i + 1] : 0.f; v.z = i + 2 < n_if ? row[i + 2] : 0.f; v.w = 0.f; }
        if (SQUARE) { v.x *= v.x; v.y *= v.y; v.z *= v.z; v.w *= v.w; }     // RDS carrier recovery: the filter input is the squared signal
        float4 vb = v;
        if (TWOSTREAMS) {
            if (i < 0) vb = *reinterpret_cast<const float4*>(tail_b + DY4_IF_TAIL + i);
            else if (i + 3 < n_if) vb = __ldg(reinterpret_cast<const float4*>(row_b + i));
            else { vb.x = i < n_if ? row_b[i] : 0.f; vb.y = i + 1 < n_if ? row_b[i + 1] : 0.f; vb.z = i + 2 < n_if ? row_b[i + 2] : 0.f; vb.w = 0.f; }
            if (SQUARE) { vb.x *= vb.x; vb.y *= vb.y; vb.z *= vb.z; vb.w *= vb.w; }
        }
        const int p = 4 * u;
        float4* d = reinterpret_cast<float4*>(&sm[p + 2 * (p / R)]);
        d[0] = make_float4(v.x, vb.x, v.y, vb.y);
        d[1] = make_float4(v.z, vb.z, v.w, vb.w);
    }
    __syncthreads();

    // output r of this thread is sample n0 + tid*R + r; tap k reads logical pair HALO + tid*R + r - k
    // = tid*R + q with q = (HALO-100) + r + 100 - k
    constexpr int C0 = HALO - (DY4_NTAPS - 1);
    const u64* w = reinterpret_cast<const u64*>(sm) + (R + 2) * tid;
    const u64* hh = reinterpret_cast<const u64*>(c_bpf2[mode].t);
    u64 acc[R];
#pragma unroll
    for (int r = 0; r < R; r++) acc[r] = 0ull;
#pragma unroll
    for (int q = R - 1 + (DY4_NTAPS - 1); q >= 0; q--) {
        const u64 x = w[C0 + q + 2 * ((C0 + q) / R)];
#pragma unroll
        for (int r = 0; r < R; r++) {
            const int k = r + (DY4_NTAPS - 1) - q;
            if (k >= 0 && k < DY4_NTAPS) acc[r] = tap2<EXACT>(acc[r], x, hh[k], nz);
        }
    }

    float op[R], os[R];
#pragma unroll
    for (int r = 0; r < R; r++) upk2(acc[r], op[r], os[r]);
    const long long o = (long long)sa * out_stride + n0 + tid * R;
    float* second = TWOSTREAMS ? (sb != sa ? pilot + (long long)(sb - sa) * out_stride : nullptr) : sband;   // where the high halves go
    const int left = n_if - (n0 + tid * R);
    if (left >= R) {
#pragma unroll
        for (int r = 0; r < R; r += 4) {
            *reinterpret_cast<float4*>(pilot + o + r) = make_float4(op[r], op[r + 1], op[r + 2], op[r + 3]);
            if (second) *reinterpret_cast<float4*>(second + o + r) = make_float4(os[r], os[r + 1], os[r + 2], os[r + 3]);
        }
    } else {
#pragma unroll
        for (int r = 0; r < R; r++) if (r < left) { pilot[o + r] = op[r]; if (second) second[o + r] = os[r]; }
    }
}

// Default (non-exact-audio) receivers: only the PILOT output reaches the PLL and has to be bit-exact; the stereo band
// feeds the audio resampler, where a fused multiply-add is within tolerance.  Scalar accumulators then cost three FP32
// instructions per tap and output (FMUL + FADD for the pilot, FFMA for the stereo band) instead of the packed pair's
// four pipe cycles, the tile is staged once (no duplicated (x,x) pairs), taps come from the constant bank as scalars.
__constant__ float c_pilot[4][DY4_NTAPS + 3], c_stereo[4][DY4_NTAPS + 3];

template <int MODE, int R, int NT>
__global__ void __launch_bounds__(NT)
k_bpf_mixed(const float* __restrict__ if_in, long long if_stride, const float* __restrict__ if_tail, long long tail_stride,
            float* __restrict__ pilot, float* __restrict__ sband, long long out_stride, int n_if)
{
    constexpr int T = NT * R, HALO = 128, NP = T + HALO;
    __shared__ float sm[NP + NP / R + 8];                        // one pad float per R: stride R+1 between threads' windows
    const int tid = threadIdx.x, n0 = blockIdx.y * T;
    const float* row = if_in + (long long)blockIdx.x * if_stride;
    const float* tail = if_tail + (long long)blockIdx.x * tail_stride;
    for (int u = tid; u < NP / 4; u += NT) {
        const int i = n0 - HALO + 4 * u;                          // multiple of 4, never straddles 0
        float4 v;
        if (i < 0) v = *reinterpret_cast<const float4*>(tail + DY4_IF_TAIL + i);
        else if (i + 3 < n_if) v = __ldg(reinterpret_cast<const float4*>(row + i));
        else { v.x = i < n_if ? row[i] : 0.f; v.y = i + 1 < n_if ? row[i + 1] : 0.f; v.z = i + 2 < n_if ? row[i + 2] : 0.f; v.w = 0.f; }
        const int p = 4 * u;                                       // R % 4 == 0: the four samples share one pad offset
        float* d = &sm[p + p / R];
        d[0] = v.x; d[1] = v.y; d[2] = v.z; d[3] = v.w;
    }
    __syncthreads();
    constexpr int C0 = HALO - (DY4_NTAPS - 1);
    const float* w = sm + (R + 1) * tid;
    float ap[R], as[R];
#pragma unroll
    for (int r = 0; r < R; r++) { ap[r] = 0.f; as[r] = 0.f; }
#pragma unroll
    for (int q = R - 1 + (DY4_NTAPS - 1); q >= 0; q--) {         // window index descends: taps ascend for every output
        const float x = w[C0 + q + (C0 + q) / R];
#pragma unroll
        for (int r = 0; r < R; r++) {
            const int k = r + (DY4_NTAPS - 1) - q;
            if (k >= 0 && k < DY4_NTAPS) {
                ap[r] = __fadd_rn(ap[r], __fmul_rn(c_pilot[MODE][k], x));      // filter.cpp:75, unfused: feeds the PLL
                as[r] = fmaf(c_stereo[MODE][k], x, as[r]);
            }
        }
    }
    const long long o = (long long)blockIdx.x * out_stride + n0 + tid * R;
    const int left = n_if - (n0 + tid * R);
    if (left >= R) {
#pragma unroll
        for (int r = 0; r < R; r += 4) {
            *reinterpret_cast<float4*>(pilot + o + r) = make_float4(ap[r], ap[r + 1], ap[r + 2], ap[r + 3]);
            *reinterpret_cast<float4*>(sband + o + r) = make_float4(as[r], as[r + 1], as[r + 2], as[r + 3]);
        }
    } else {
#pragma unroll
        for (int r = 0; r < R; r++) if (r < left) { pilot[o + r] = ap[r]; sband[o + r] = as[r]; }
    }
}

}  // namespace

cudaError_t dy4_launch_bpf_mixed(const Dy4BpfArgs& a, cudaStream_t st)
{
    if (a.n_if <= 0 || a.n_streams <= 0) return cudaSuccess;
    constexpr int R = 8, NT = 128;
    const long long ts = a.if_tail_stride ? a.if_tail_stride : DY4_IF_TAIL;
    dim3 grid(a.n_streams, (a.n_if + NT * R - 1) / (NT * R));
    switch (a.mode) {
    case 0: k_bpf_mixed<0, R, NT><<<grid, NT, 0, st>>>(a.if_in, a.if_stride, a.if_tail, ts, a.pilot, a.sband, a.out_stride, a.n_if); break;
    case 1: k_bpf_mixed<1, R, NT><<<grid, NT, 0, st>>>(a.if_in, a.if_stride, a.if_tail, ts, a.pilot, a.sband, a.out_stride, a.n_if); break;
    case 2: k_bpf_mixed<2, R, NT><<<grid, NT, 0, st>>>(a.if_in, a.if_stride, a.if_tail, ts, a.pilot, a.sband, a.out_stride, a.n_if); break;
    case 3: k_bpf_mixed<3, R, NT><<<grid, NT, 0, st>>>(a.if_in, a.if_stride, a.if_tail, ts, a.pilot, a.sband, a.out_stride, a.n_if); break;
    default: return cudaErrorInvalidValue;
    }
    g_dy4_launches++;
    return cudaGetLastError();
}

cudaError_t dy4_launch_bpf(const Dy4BpfArgs& a, cudaStream_t st)
{
    if (a.n_if <= 0 || a.n_streams <= 0) return cudaSuccess;
    constexpr int R = 8, NT = 128;
    const long long ts = a.if_tail_stride ? a.if_tail_stride : DY4_IF_TAIL;
    dim3 grid(a.variant == 0 ? a.n_streams : (a.n_streams + 1) / 2, (a.n_if + NT * R - 1) / (NT * R));
    if (a.variant == 0) k_twin_bpf<R, NT, true, false, false><<<grid, NT, 0, st>>>(a.if_in, a.if_stride, a.if_tail, ts, a.pilot, a.sband, a.out_stride, a.n_if, a.neg_zero2, a.mode, a.n_streams);
    else if (a.variant == 1) k_twin_bpf<R, NT, false, false, true><<<grid, NT, 0, st>>>(a.if_in, a.if_stride, a.if_tail, ts, a.pilot, nullptr, a.out_stride, a.n_if, a.neg_zero2, a.mode, a.n_streams);
    else k_twin_bpf<R, NT, false, true, true><<<grid, NT, 0, st>>>(a.if_in, a.if_stride, a.if_tail, ts, a.pilot, nullptr, a.out_stride, a.n_if, a.neg_zero2, a.mode, a.n_streams);
    g_dy4_launches++;
    return cudaGetLastError();
}

cudaError_t dy4_upload_taps_bpf(const TapPairs* bpf6)
{
    static float hp[4][DY4_NTAPS + 3], hs[4][DY4_NTAPS + 3];
    for (int m = 0; m < 4; m++)
        for (int k = 0; k < DY4_NTAPS + 3; k++) { hp[m][k] = bpf6[m].t[k].x; hs[m][k] = bpf6[m].t[k].y; }
    cudaError_t e = cudaMemcpyToSymbol(c_pilot, hp, sizeof(hp));
    if (e == cudaSuccess) e = cudaMemcpyToSymbol(c_stereo, hs, sizeof(hs));
    if (e != cudaSuccess) return e;
    return cudaMemcpyToSymbol(c_bpf2, bpf6, sizeof(TapPairs) * 6);
}
