// dy4_frontend.cu — fused RF front end: packed uint8 IQ -> IF (FM-demodulated) samples.
//
// Replaces, for a batch of streams, the reference's
//   readStdinBlockData arithmetic     src/iofunc.cpp:117-119   ((b-128)/128)
//   de-interleave in frontend()       src/project.cpp:78-81
//   downsampleBlockConvolveFIR x2     src/filter.cpp:123-140   (I and Q, 101 taps, keep every D-th)
//   fmDemodArctan                     src/filter.cpp:85-102    (derivative discriminator)
// in one pass: only the kept output phases are computed, I and Q ride in the two
// halves of one packed f32x2 register (same taps, same indices), the block's
// input plus its 100+D-sample history is staged once in shared memory as float
// pairs, and the discriminator runs as the FIR epilogue.  Arithmetic is the
// reference's exactly (see dy4_common.cuh), so IF is bit-identical.
//
// Absolute indexing: history before the chunk comes from iq_tail (zeros, i.e.
// byte 128, at stream start), which reproduces the reference's carried
// i_state_rf/q_state_rf/prev_I/prev_Q (project.cpp:25-30) for any block split.
#include "dy4_common.cuh"
#include "dy4_kernels.h"
#include "dy4_internal.h"

namespace {

__constant__ TapPairs c_rf2[4];   // (h,h) pairs of the RF low-pass, per mode

template <int D, int R, int NT, bool EXACT>
__global__ void __launch_bounds__(NT, 3)
k_frontend(const uint8_t* __restrict__ iq, long long row_stride, const uint8_t* __restrict__ iq_tail,
           float* __restrict__ if_out, long long if_stride, int n_if,
           const float* __restrict__ taps_g, u64 nz, int mode)
{
    extern __shared__ __align__(16) uint32_t sm[];   // one word per IQ sample: bf16(I) | bf16(Q) << 16 (exact for k/128)
    __shared__ float2 s_last[NT];
    constexpr int T = NT * R;               // IF samples per tile
    constexpr int CH = D * R;               // input samples per thread chunk
    constexpr int HALO = DY4_IQ_TAIL / 2;   // 112 samples of history in front of the tile
    const int tid = threadIdx.x;
    const int m0 = blockIdx.x * T;
    const uint8_t* row = iq + (long long)blockIdx.y * row_stride;
    const uint8_t* tail = iq_tail + (long long)blockIdx.y * DY4_IQ_TAIL;
    const long long row_bytes = 2LL * D * n_if;

    // ---- stage: 8 bytes (4 IQ samples) per thread per step -> one 16-byte shared store -------------
    // tile-relative byte 0 is absolute byte 2*D*m0 - 224; an 8-byte unit never straddles byte 0.
    // Branch-free so that the unrolled loads are all in flight together: the address is always valid (history
    // bytes come from the tail buffer, bytes past the end of the row are re-read from its last unit and masked).
    constexpr int UNITS = (2 * D * T + DY4_IQ_TAIL) / 8;
    const long long b0 = 2LL * D * m0 - DY4_IQ_TAIL;
    const long long last_unit = row_bytes - 8;
#pragma unroll 5
    for (int u = tid; u < UNITS; u += NT) {
        const long long b = b0 + 8LL * u;
        const uint8_t* src = b < 0 ? tail + (DY4_IQ_TAIL + b) : row + (b < last_unit ? b : last_unit);
        uint2 w = __ldg(reinterpret_cast<const uint2*>(src));
        if (b > last_unit) w = make_uint2(0x80808080u, 0x80808080u);
        // (b-128)/128 exactly: 0x4B0000bb is 2^23+b; (2^23+b)*2^-7 - 65537 = (b-128)/128, no rounding anywhere;
        // the result has at most 8 significant bits, so its top 16 bits are the exact bf16
        uint4 o;
        const uint32_t i0 = __float_as_uint(fmaf(__uint_as_float(__byte_perm(w.x, 0x4B000000u, 0x7440)), 0.0078125f, -65537.0f));
        const uint32_t q0 = __float_as_uint(fmaf(__uint_as_float(__byte_perm(w.x, 0x4B000000u, 0x7441)), 0.0078125f, -65537.0f));
        const uint32_t i1 = __float_as_uint(fmaf(__uint_as_float(__byte_perm(w.x, 0x4B000000u, 0x7442)), 0.0078125f, -65537.0f));
        const uint32_t q1 = __float_as_uint(fmaf(__uint_as_float(__byte_perm(w.x, 0x4B000000u, 0x7443)), 0.0078125f, -65537.0f));
        const uint32_t i2 = __float_as_uint(fmaf(__uint_as_float(__byte_perm(w.y, 0x4B000000u, 0x7440)), 0.0078125f, -65537.0f));
        const uint32_t q2 = __float_as_uint(fmaf(__uint_as_float(__byte_perm(w.y, 0x4B000000u, 0x7441)), 0.0078125f, -65537.0f));
        const uint32_t i3 = __float_as_uint(fmaf(__uint_as_float(__byte_perm(w.y, 0x4B000000u, 0x7442)), 0.0078125f, -65537.0f));
        const uint32_t q3 = __float_as_uint(fmaf(__uint_as_float(__byte_perm(w.y, 0x4B000000u, 0x7443)), 0.0078125f, -65537.0f));
        o.x = __byte_perm(i0, q0, 0x7632); o.y = __byte_perm(i1, q1, 0x7632);
        o.z = __byte_perm(i2, q2, 0x7632); o.w = __byte_perm(i3, q3, 0x7632);
        const int p = 4 * u;
        *reinterpret_cast<uint4*>(&sm[p + 4 * (p / CH)]) = o;
    }
    __syncthreads();

    // ---- FIR: R consecutive (I,Q) outputs per thread -------------------------------------------------
    u64 acc[R];
    const u64* hh = reinterpret_cast<const u64*>(c_rf2[mode].t);
    pair_decim_fir_bf16<D, R, EXACT, HALO - (DY4_NTAPS - 1)>(sm + (CH + 4) * tid, hh, nz, acc);

    // ---- the sample before the tile (for the discriminator's first difference) ---------------------
    // One extra output, m0-1, on thread 0: same taps from the constant bank, same ascending order, window
    // offsets known at compile time (logical sample HALO - D - k).
    float pI, pQ;
    upk2(acc[R - 1], pI, pQ);
    s_last[tid] = make_float2(pI, pQ);
    if (tid == 0) {
        u64 a0 = 0ull;
#pragma unroll
        for (int k = 0; k < DY4_NTAPS; k++) {
            const int p = HALO - D - k;
            const uint32_t wv = sm[p + 4 * (p / CH)];
            a0 = tap2<EXACT>(a0, pk2(__uint_as_float(wv << 16), __uint_as_float(wv & 0xffff0000u)), hh[k], nz);
        }
        upk2(a0, pI, pQ);
    }
    __syncthreads();
    if (tid > 0) { const float2 l = s_last[tid - 1]; pI = l.x; pQ = l.y; }

    // ---- discriminator, reference arithmetic: double sum of squares narrowed to float, float rest ---
    float out[R];
#pragma unroll
    for (int r = 0; r < R; r++) {
        float I, Q;
        upk2(acc[r], I, Q);
        const float den = __double2float_rn(fma((double)I, (double)I, (double)Q * (double)Q));
        const float num = __fsub_rn(__fmul_rn(I, __fsub_rn(Q, pQ)), __fmul_rn(Q, __fsub_rn(I, pI)));
        out[r] = (den == 0.0f) ? 0.0f : __fdiv_rn(num, den);
        pI = I; pQ = Q;
    }
    float* dst = if_out + (long long)blockIdx.y * if_stride + m0 + tid * R;
    const int left = n_if - (m0 + tid * R);
    if (left >= R) {
#pragma unroll
        for (int r = 0; r < R; r += 4) *reinterpret_cast<float4*>(dst + r) = make_float4(out[r], out[r + 1], out[r + 2], out[r + 3]);
    } else {
#pragma unroll
        for (int r = 0; r < R; r++) if (r < left) dst[r] = out[r];
    }
}

template <int D, int R, int NT, bool EXACT>
cudaError_t launch(const Dy4FrontendArgs& a, cudaStream_t st)
{
    constexpr int T = NT * R;
    const size_t smem = sizeof(uint32_t) * dy4_padded_words_bf16(D, R, NT);
    auto kern = k_frontend<D, R, NT, EXACT>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    dim3 grid((a.n_if + T - 1) / T, a.n_streams);
    kern<<<grid, NT, smem, st>>>(a.iq, a.row_stride, a.iq_tail, a.if_out, a.if_stride, a.n_if, a.taps_g, a.neg_zero2, a.mode);
    g_dy4_launches++;
    return cudaGetLastError();
}

}  // namespace

cudaError_t dy4_launch_frontend(const Dy4FrontendArgs& a, cudaStream_t st)
{
    if (a.n_if <= 0 || a.n_streams <= 0) return cudaSuccess;
    if (a.rf_decim == 10) return launch<10, 8, 128, true>(a, st);
    if (a.rf_decim == 5) return launch<5, 8, 128, true>(a, st);
    return cudaErrorInvalidValue;
}

cudaError_t dy4_upload_taps_frontend(const TapPairs* rf4) { return cudaMemcpyToSymbol(c_rf2, rf4, sizeof(TapPairs) * 4); }
