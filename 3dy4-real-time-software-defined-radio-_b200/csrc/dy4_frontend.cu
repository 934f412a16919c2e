// dy4_frontend.cu — fused RF front end: packed uint8 IQ -> IF (FM-demodulated) samples.
//
// Replaces, for a batch of streams, the reference's
//   readStdinBlockData arithmetic     src/iofunc.cpp:117-119   ((b-128)/128)
//   de-interleave in frontend()       src/project.cpp:78-81
//   downsampleBlockConvolveFIR x2     src/filter.cpp:123-140   (I and Q, 101 taps, keep every D-th)
//   fmDemodArctan                     src/filter.cpp:85-102    (derivative discriminator)
// in one pass, with I and Q riding in the two halves of one packed f32x2 register (same taps, same indices) and the
// discriminator as the FIR's epilogue.  Arithmetic is the reference's exactly (see dy4_common.cuh), so IF is bit-identical.
//
// Absolute indexing: history before the chunk comes from iq_tail (zeros, i.e.
// byte 128, at stream start), which reproduces the reference's carried
// i_state_rf/q_state_rf/prev_I/prev_Q (project.cpp:25-30) for any block split.
#include "dy4_common.cuh"
#include "dy4_kernels.h"
#include "dy4_internal.h"

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <string>

namespace {


// The RF taps (scaled by 1/128: the (b-128)/128 of iofunc.cpp:117 folded in) as compile-time constants, i.e. immediate
// operands: kRfTapsScaled[mode][k], and a host copy for the start-up check that they are what dy4_lpf_taps() designs here
#define DY4_TAPS_QUAL __device__
#define DY4_TAPS_NAME kRfTapsScaled
#include "dy4_rf_taps.inc"
#undef DY4_TAPS_QUAL
#undef DY4_TAPS_NAME
#define DY4_TAPS_QUAL
#define DY4_TAPS_NAME kRfTapsScaledHost
#include "dy4_rf_taps.inc"
#undef DY4_TAPS_QUAL
#undef DY4_TAPS_NAME
template <int MODE, int K> __device__ __forceinline__ u64 rf_tap() { return pk2(kRfTapsScaled[MODE][K], kRfTapsScaled[MODE][K]); }

// (I,Q) of the IQ sample held in the 16-bit half `h` (0 or 1) of word w, as exact integers b-128 in a packed pair
__device__ __forceinline__ u64 unpack_iq(uint32_t w, int h, u64 neg_bias)
{
    const float fi = __uint_as_float(__byte_perm(w, 0x4B000000u, h ? 0x7442 : 0x7440));   // 2^23 + I byte
    const float fq = __uint_as_float(__byte_perm(w, 0x4B000000u, h ? 0x7443 : 0x7441));   // 2^23 + Q byte
    return fadd2(pk2(fi, fq), neg_bias);                                                  // - (2^23 + 128): exact
}


// ------------------------------------------------------------------------------------------------------------
// One thread walks one SEGMENT of a stream's outputs, backwards in time, with the FIR in transposed
// form.  The reference sums h[0]x[Dm] + h[1]x[Dm-1] + ... in that order, so an output's accumulator must meet its
// samples newest first: walking the input downwards, every sample feeds all ~101/D accumulators whose windows cover
// it (tap k = D(m-M)+c is a compile-time constant for a position c in the loop body), an accumulator is born at tap
// 0 and retires after tap 100.  Per sample that is ONE unpack for 10.1 (D=10) or 20.2 (D=5) multiply-adds — the
// windowed (tile-staged) kernels of round 1 managed 4.7 — with no shared memory, no barriers and no tile edges: the body of the loop is
// 40 samples = five aligned 16-byte loads = 40/D outputs; the NSLOT = 40/D + 100/D live accumulators are renamed by
// 40/D slots per body.  Threads of a warp read addresses one segment apart; each 32-byte sector is consumed whole by
// its thread (two back-to-back loads) and L1 is the staging buffer, so DRAM traffic stays at the algorithmic 2 bytes
// per sample.  The price of the uniform body is a ramp of 100/D + 1 outputs per segment (accumulators of the
// neighbouring segment are computed and thrown away), hence long segments (seg_len, chosen per launch).
// Taps are compile-time constant-bank addresses (one table per mode) and the -0 addend sits in a vector register, so
// ptxas feeds the taps to FFMA2 through uniform registers: no tap ever occupies the vector register file.
// ------------------------------------------------------------------------------------------------------------
template <int MODE, int D, bool EXACT, int UNPACK, int MINB>
__global__ void __launch_bounds__(128, MINB)
k_frontend_stream(const uint8_t* __restrict__ iq, long long row_stride, const uint8_t* __restrict__ iq_tail,
                  float* __restrict__ if_out, long long if_stride, int n_if, u64 nz, int seg_len, int segs_per_stream,
                  long long n_threads)
{
    constexpr int BODY = 40;                            // samples per loop body: 5 aligned 16-byte groups
    constexpr int NOUT = BODY / D;                      // outputs retired per body
    constexpr int SPAN = (DY4_NTAPS - 1) / D;           // 10 or 20
    constexpr int NSLOT = NOUT + SPAN;                  // live accumulators; slot e holds output M + e - NOUT
    constexpr int TOP = NSLOT - NOUT;                   // discriminator outputs of a body: M+TOP down to M+TOP-NOUT+1
    constexpr int LO = TOP - NOUT + 1;
    constexpr int A0 = (LO + 3) & ~3;                   // first 4-aligned output offset
    constexpr int NC = A0 - LO;                         // outputs carried to the next body to complete a 16-byte store
    static_assert(BODY % D == 0 && (DY4_NTAPS - 1) % D == 0 && NOUT % 4 == 0, "body geometry");
    __shared__ u64 s_nz[128];
    const int tid = threadIdx.x;
    s_nz[tid] = nz;
    const u64 nzv = *reinterpret_cast<volatile u64*>(&s_nz[tid]);     // (-0,-0) in a VECTOR register (see above)
    const long long gt = (long long)blockIdx.x * 128 + tid;
    if (gt >= n_threads) return;
    const int s = (int)(gt / segs_per_stream), seg = (int)(gt - (long long)s * segs_per_stream);
    const int m_lo = seg * seg_len, m_hi = min(m_lo + seg_len, n_if);
    const uint8_t* row = iq + (long long)s * row_stride;
    const uint8_t* tail = iq_tail + (long long)s * DY4_IQ_TAIL;
    float* orow = if_out + (long long)s * if_stride;
    const u64 neg_bias = pk2(-8388736.0f, -8388736.0f);

    // 16-byte group `i` (0..4, ascending time) of the body that ends below output index M: samples D*M-40+8i .. +7.
    // Only the first segment of a stream ever reaches below the chunk (offsets < 0: carried tail, then don't-care);
    // everywhere else the load is a plain pointer + immediate offset from a per-body base.
    auto load_group_checked = [&](int M, int i) -> uint4 {
        const long long off = 2LL * D * M - 2 * BODY + 16 * i;
        if (off >= 0) return __ldg(reinterpret_cast<const uint4*>(row + off));
        if (off >= -DY4_IQ_TAIL) return __ldg(reinterpret_cast<const uint4*>(tail + DY4_IQ_TAIL + off));
        return make_uint4(0x80808080u, 0x80808080u, 0x80808080u, 0x80808080u);            // feeds discarded accumulators only
    };

    u64 acc[NSLOT];
#pragma unroll
    for (int e = 0; e < NSLOT; e++) acc[e] = 0ull;
    float o[NOUT + NC];
#pragma unroll
    for (int i = 0; i < NOUT + NC; i++) o[i] = 0.f;
    float cI = 0.f, cQ = 0.f;                           // (I,Q) of the output retired last (one index above the next)
    // groups are consumed newest first (4, 3, .. 0); PF of them are kept in flight ahead of the one being filtered
    constexpr int PF = 3;
    uint4 q[PF + 1];
#pragma unroll
    for (int j = 0; j <= PF; j++) q[j] = load_group_checked(m_hi, 4 - j);

    const uint4* base = reinterpret_cast<const uint4*>(row + 2LL * D * m_hi - 2 * BODY);   // group 0 of the current body
    for (int M = m_hi; M + TOP >= m_lo; M -= NOUT, base -= 2 * BODY / 16) {
        const bool plain = M >= 2 * NOUT;               // every group this body asks for lies inside the chunk
#pragma unroll
        for (int gi = 4; gi >= 0; gi--) {
            const uint4 gv = q[0];
#pragma unroll
            for (int j = 0; j < PF; j++) q[j] = q[j + 1];
            {   // refill the far end of the queue: PF+1 groups below the current one (wrapping into the next body)
                const int ahead = gi - (PF + 1);
                if (plain) q[PF] = __ldg(base + ahead);
                else q[PF] = ahead >= 0 ? load_group_checked(M, ahead) : load_group_checked(M - NOUT, ahead + 5);
            }
#pragma unroll
            for (int l = 7; l >= 0; l--) {                  // samples of the group, newest first
                const int p = 8 * gi + l;                   // position in the 40-sample body, ascending time
                const int c = BODY - p;                     // the sample is D*M - c
                const uint32_t w = (l / 2) == 0 ? gv.x : (l / 2) == 1 ? gv.y : (l / 2) == 2 ? gv.z : gv.w;
                u64 x;
                if (UNPACK == 1) {
                    const uint32_t sgn = w ^ 0x80808080u;
                    x = (l & 1) ? pk2((float)(int8_t)((sgn >> 16) & 0xffu), (float)(int8_t)(sgn >> 24))
                                : pk2((float)(int8_t)(sgn & 0xffu), (float)(int8_t)((sgn >> 8) & 0xffu));
                } else {
                    x = unpack_iq(w, l & 1, neg_bias);
                }
#pragma unroll
                for (int e = 0; e < NSLOT; e++) {
                    const int k = D * (e - NOUT) + c;       // tap index of this sample in output M + e - NOUT
                    // tap 0 of the Hann-windowed design is exactly 0 and opens the sum: +0 + (+-0) = +0, nothing to do
                    if (k == 0 && kRfTapsScaled[MODE][0] == 0.0f) continue;
                    if (k >= 0 && k < DY4_NTAPS) acc[e] = tap2<EXACT>(acc[e], x, pk2(kRfTapsScaled[MODE][k], kRfTapsScaled[MODE][k]), nzv);
                }
            }
        }
        // ---- retire NOUT outputs (slots NSLOT-1 .. NSLOT-NOUT, descending index) through the discriminator:
        // out[m+1] needs (I,Q) of m+1 (kept from the previous retirement) and of m (just finished)
#pragma unroll
        for (int i = 0; i < NOUT; i++) {
            float pI, pQ;
            upk2(acc[NSLOT - 1 - i], pI, pQ);
            const float den = __double2float_rn(fma((double)cI, (double)cI, (double)cQ * (double)cQ));
            const float num = __fsub_rn(__fmul_rn(cI, __fsub_rn(cQ, pQ)), __fmul_rn(cQ, __fsub_rn(cI, pI)));
            o[NOUT - 1 - i] = (den == 0.0f) ? 0.0f : __fdiv_rn(num, den);     // offset TOP - i  ->  index TOP - i - LO
            cI = pI; cQ = pQ;
        }
#pragma unroll
        for (int j = 0; j < NOUT / 4; j++) {
            const int a = M + A0 + 4 * j;
            if (a >= m_lo && a + 3 < m_hi)
                *reinterpret_cast<float4*>(orow + a) = make_float4(o[NC + 4 * j], o[NC + 4 * j + 1], o[NC + 4 * j + 2], o[NC + 4 * j + 3]);
        }
#pragma unroll
        for (int i = 0; i < NC; i++) o[NOUT + i] = o[i];
#pragma unroll
        for (int e = NSLOT - 1; e >= NOUT; e--) acc[e] = acc[e - NOUT];
#pragma unroll
        for (int e = 0; e < NOUT; e++) acc[e] = 0ull;
    }
}

template <int MODE, int D, bool EXACT, int UNPACK, int MINB>
cudaError_t launch_stream(const Dy4FrontendArgs& a, cudaStream_t st)
{
    auto kern = k_frontend_stream<MODE, D, EXACT, UNPACK, MINB>;
    static int sms = 0;
    if (!sms) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev); }
    // Segment length.  The ramp costs 100/D + 1 outputs per segment, so segments should be long; but the grid should
    // also be WHOLE waves of resident CTAs (a last, nearly empty wave leaves most SMs idle while it drains):
    // give every stream the largest number of segments that still fits those waves, never going below 64 outputs per
    // segment (small launches then simply do not fill the machine).  A multiple of 8 keeps segment
    // ends on 16-byte input and output groups.
    static int ctas_per_sm = 0;
    if (!ctas_per_sm) {
        cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas_per_sm, kern, 128, 0);
        if (e != cudaSuccess) return e;
        if (ctas_per_sm < 1) ctas_per_sm = 1;
    }
    const long long resident = (long long)sms * ctas_per_sm * 128;
    // whole waves: with the longest useful segment (1024 outputs) the launch needs `waves` waves of resident threads;
    // fill exactly that many with as many (hence as short as necessary) segments per stream as fit
    const long long threads_min = (long long)a.n_streams * ((a.n_if + 1023) / 1024);
    const long long waves = std::max<long long>(1, (threads_min + resident - 1) / resident);
    const long long segs_fit = std::max<long long>(1, waves * resident / a.n_streams);
    int seg = (int)((a.n_if + segs_fit - 1) / segs_fit);
    seg = std::max(64, (seg + 7) & ~7);
    const int segs = (a.n_if + seg - 1) / seg;
    const long long n_threads = (long long)segs * a.n_streams;
    const long long blocks = (n_threads + 127) / 128;
    if (blocks > 0x7fffffffLL) return cudaErrorInvalidValue;
    kern<<<(unsigned)blocks, 128, 0, st>>>(a.iq, a.row_stride, a.iq_tail, a.if_out, a.if_stride, a.n_if, a.neg_zero2, seg, segs, n_threads);
    g_dy4_launches++;
    return cudaGetLastError();
}

template <int UNPACK, int MINB>
cudaError_t launch_stream_mode(const Dy4FrontendArgs& a, cudaStream_t st)
{
    if (a.n_if % 8) return cudaErrorInvalidValue;      // whole blocks only (1024-multiples in every mode)
    constexpr int MINB5 = MINB > 4 ? 4 : MINB;        // D = 5 keeps 28 accumulators (56 registers): four CTAs per SM
    // a.exact == 0 (mono receivers without DY4_FLAG_EXACT_AUDIO): nothing downstream is chaotic, so the multiply-add is
    // fused — one FFMA2 per tap and I/Q pair instead of two instructions, ~1e-7 from the reference's IF.
    switch (a.mode * 2 + (a.exact ? 1 : 0)) {
    case 0: return launch_stream<0, 10, false, UNPACK, MINB>(a, st);
    case 1: return launch_stream<0, 10, true, UNPACK, MINB>(a, st);
    case 2: return launch_stream<1, 5, false, UNPACK, MINB5>(a, st);
    case 3: return launch_stream<1, 5, true, UNPACK, MINB5>(a, st);
    case 4: return launch_stream<2, 10, false, UNPACK, MINB>(a, st);
    case 5: return launch_stream<2, 10, true, UNPACK, MINB>(a, st);
    case 6: return launch_stream<3, 5, false, UNPACK, MINB5>(a, st);
    case 7: return launch_stream<3, 5, true, UNPACK, MINB5>(a, st);
    }
    return cudaErrorInvalidValue;
}
}  // namespace

cudaError_t dy4_launch_frontend(const Dy4FrontendArgs& a, cudaStream_t st)
{
    if (a.n_if <= 0 || a.n_streams <= 0) return cudaSuccess;
    return launch_stream_mode<1, 6>(a, st);          // I2F unpack, six CTAs of 128 threads per SM
}

// The kernels carry the taps as immediates (dy4_rf_taps.inc, generated at build time from dy4_taps.cpp): refuse to run on a build
// whose constants are not what dy4_lpf_taps() designs on this machine.
cudaError_t dy4_upload_taps_frontend(const TapPairs* rf4)
{
    for (int m = 0; m < 4; m++)
        for (int k = 0; k < DY4_NTAPS; k++) {
            const float scaled = rf4[m].t[k].x * 0.0078125f;                      // exact: power of two
            if (std::memcmp(&scaled, &kRfTapsScaledHost[m][k], sizeof(float)) != 0) {
                dy4_set_error("compiled-in RF taps (dy4_rf_taps.inc) differ from dy4_lpf_taps(): rebuild the library");
                return cudaErrorInvalidValue;
            }
        }
    return cudaSuccess;
}
