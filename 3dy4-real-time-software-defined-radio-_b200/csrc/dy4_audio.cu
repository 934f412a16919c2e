// dy4_audio.cu — audio back end: mono delay, 38 kHz mix, the two rational
// resamplers, L/R combine and int16 conversion, in one kernel.
//
// Replaces, from backend() src/project.cpp:95-134 and main() :307-317:
//   delayBlock            filter.cpp:229-251  (50-sample delay = an index offset here)
//   pointwiseMultiply     filter.cpp:253-266  (nco * stereo-band * 2)
//   resampleBlockConvolveFIR x2  filter.cpp:142-173  (mono path and L-R path)
//   pointwiseAdd/Subtract filter.cpp:267-290, interleave :291-301
//   float -> int16        project.cpp:313-316 (truncate toward zero, NaN -> 0)
// y[m] = sum_j h[phase + j*U] * x[floor(m*D/U) - j], phase = (m*D) mod U: only the
// non-zero taps of the zero-stuffed signal are visited, only kept outputs computed.
//
// U == 1 (modes 0,1): a decimating 101-tap FIR.  The mono path (delayed IF) and
// the difference path (mixed signal) use the same taps and indices, so the pair
// (IF[i-50], mixed[i]) rides in one packed f32x2 exactly like (I,Q) in the front end.
// U == 147 (modes 2,3): per-output phase, taps read from a transposed
// [tap][phase] table; one thread per output.
#include "dy4_common.cuh"
#include "dy4_kernels.h"
#include "dy4_internal.h"

#include <algorithm>

namespace {

__constant__ TapPairs c_audio2[4];   // (h,h) pairs of the 101-tap audio low-pass, modes with U == 1

__device__ __forceinline__ int16_t pcm16(float x)
{
    // project.cpp:314-315: NaN -> 0, else static_cast<short>(x*16384): cvttss2si then low 16 bits
    if (x != x) return 0;
    return (int16_t)(uint16_t)(__float2int_rz(__fmul_rn(x, 16384.0f)) & 0xffff);
}

// history-extended reads: index i is relative to the chunk start, may be negative
__device__ __forceinline__ float if_at(const float* row, const float* tail, int n, int i)
{
    return i < 0 ? tail[DY4_IF_TAIL + i] : (i < n ? __ldg(row + i) : 0.0f);
}
__device__ __forceinline__ float mix_at(const float* nco, const float* sband, const float* tail, int n, int i)
{
    if (i < 0) return tail[DY4_MIX_TAIL + i];
    if (i >= n) return 0.0f;
    return __fmul_rn(__fmul_rn(__ldg(nco + i), __ldg(sband + i)), 2.0f);       // filter.cpp:264
}

template <int D, int R, int NT, bool EXACT, bool STEREO>
__global__ void __launch_bounds__(NT)
k_audio_u1(const float* __restrict__ if_in, long long if_stride, const float* __restrict__ if_tail, long long tail_stride,
           const float* __restrict__ nco, const float* __restrict__ sband, long long bb_stride,
           const float* __restrict__ mix_tail, float* __restrict__ audio, long long audio_stride,
           int16_t* __restrict__ pcm, long long pcm_stride, int n_if, int n_audio,
           u64 nz, int mode)
{
    extern __shared__ __align__(16) float2 sm[];
    constexpr int T = NT * R;
    constexpr int CH = D * R;
    constexpr int HALO = 112;
    constexpr int DELAY = DY4_NTAPS / 2;            // project.cpp:251: mono_delay_state has num_taps/2 = 50 entries
    const int tid = threadIdx.x;
    const int m0 = blockIdx.y * T;
    const int s = blockIdx.x;                       // streams on grid.x (no 65535 limit), tiles on grid.y
    const float* row = if_in + (long long)s * if_stride;
    const float* itail = if_tail + (long long)s * tail_stride;
    const float* nrow = STEREO ? nco + (long long)s * bb_stride : nullptr;
    const float* srow = STEREO ? sband + (long long)s * bb_stride : nullptr;
    const float* mtail = STEREO ? mix_tail + (long long)s * DY4_MIX_TAIL : nullptr;

    // logical pair p <-> IF-rate index i = D*m0 - HALO + p ; pair = (delayed IF, mixed).
    // Staged four pairs at a time: D*m0 - HALO is a multiple of 4, so the NCO and stereo-band rows are read with aligned
    // 16-byte loads and the delayed IF (offset -50: 8-byte aligned) with two 8-byte loads; units that touch the carried
    // history or the end of the row take the element-wise path.  A unit never straddles a pad (CH % 4 == 0).
    constexpr int NP = D * T + HALO;
    static_assert(NP % 4 == 0 && CH % 4 == 0 && (D * T) % 4 == 0 && HALO % 4 == 0 && DELAY % 2 == 0, "vector staging geometry");
    for (int u = tid; u < NP / 4; u += NT) {
        const int p = 4 * u;
        const int i = D * m0 - HALO + p;
        float x0, x1, x2, x3, y0 = 0.f, y1 = 0.f, y2 = 0.f, y3 = 0.f;
        if (i - DELAY >= 0 && i + 3 < n_if) {
            const float2 a = __ldg(reinterpret_cast<const float2*>(row + i - DELAY));
            const float2 b = __ldg(reinterpret_cast<const float2*>(row + i - DELAY + 2));
            x0 = a.x; x1 = a.y; x2 = b.x; x3 = b.y;
            if (STEREO) {
                const float4 nv = __ldg(reinterpret_cast<const float4*>(nrow + i));
                const float4 sv = __ldg(reinterpret_cast<const float4*>(srow + i));
                y0 = __fmul_rn(__fmul_rn(nv.x, sv.x), 2.0f); y1 = __fmul_rn(__fmul_rn(nv.y, sv.y), 2.0f);     // filter.cpp:264
                y2 = __fmul_rn(__fmul_rn(nv.z, sv.z), 2.0f); y3 = __fmul_rn(__fmul_rn(nv.w, sv.w), 2.0f);
            }
        } else {
            x0 = if_at(row, itail, n_if, i - DELAY); x1 = if_at(row, itail, n_if, i + 1 - DELAY);
            x2 = if_at(row, itail, n_if, i + 2 - DELAY); x3 = if_at(row, itail, n_if, i + 3 - DELAY);
            if (STEREO) {
                y0 = mix_at(nrow, srow, mtail, n_if, i); y1 = mix_at(nrow, srow, mtail, n_if, i + 1);
                y2 = mix_at(nrow, srow, mtail, n_if, i + 2); y3 = mix_at(nrow, srow, mtail, n_if, i + 3);
            }
        }
        float4* d = reinterpret_cast<float4*>(&sm[p + 2 * (p / CH)]);
        d[0] = make_float4(x0, y0, x1, y1);
        d[1] = make_float4(x2, y2, x3, y3);
    }
    __syncthreads();

    u64 acc[R];
    const u64* w = reinterpret_cast<const u64*>(sm) + (CH + 2) * tid;
    pair_decim_fir<D, R, EXACT, HALO - (DY4_NTAPS - 1)>(w, reinterpret_cast<const u64*>(c_audio2[mode].t), nz, acc);

    const int m = m0 + tid * R;
    const int left = n_audio - m;
    if (left <= 0) return;
    if (STEREO) {
        float o[2 * R];
#pragma unroll
        for (int r = 0; r < R; r++) {
            float mono, diff;
            upk2(acc[r], mono, diff);
            o[2 * r] = __fadd_rn(mono, diff);       // project.cpp:131  left
            o[2 * r + 1] = __fsub_rn(mono, diff);   // project.cpp:132  right
        }
        if (left >= R) {
            if (audio) {
                float4* d = reinterpret_cast<float4*>(audio + (long long)s * audio_stride + 2LL * m);
#pragma unroll
                for (int r = 0; r < 2 * R; r += 4) d[r / 4] = make_float4(o[r], o[r + 1], o[r + 2], o[r + 3]);
            }
            if (pcm) {
                uint4* d = reinterpret_cast<uint4*>(pcm + (long long)s * pcm_stride + 2LL * m);
#pragma unroll
                for (int r = 0; r < 2 * R; r += 8) {
                    uint4 q;
                    q.x = (uint16_t)pcm16(o[r]) | ((uint32_t)(uint16_t)pcm16(o[r + 1]) << 16);
                    q.y = (uint16_t)pcm16(o[r + 2]) | ((uint32_t)(uint16_t)pcm16(o[r + 3]) << 16);
                    q.z = (uint16_t)pcm16(o[r + 4]) | ((uint32_t)(uint16_t)pcm16(o[r + 5]) << 16);
                    q.w = (uint16_t)pcm16(o[r + 6]) | ((uint32_t)(uint16_t)pcm16(o[r + 7]) << 16);
                    d[r / 8] = q;
                }
            }
        } else {
            for (int r = 0; r < 2 * R; r++) if (r < 2 * left) {
                if (audio) audio[(long long)s * audio_stride + 2LL * m + r] = o[r];
                if (pcm) pcm[(long long)s * pcm_stride + 2LL * m + r] = pcm16(o[r]);
            }
        }
    } else {
        float o[R];
#pragma unroll
        for (int r = 0; r < R; r++) { float d; upk2(acc[r], o[r], d); }
        if (left >= R && R % 8 == 0) {
            if (audio) {
                float4* d = reinterpret_cast<float4*>(audio + (long long)s * audio_stride + m);
#pragma unroll
                for (int r = 0; r < R; r += 4) d[r / 4] = make_float4(o[r], o[r + 1], o[r + 2], o[r + 3]);
            }
            if (pcm) {
                uint4* d = reinterpret_cast<uint4*>(pcm + (long long)s * pcm_stride + m);
#pragma unroll
                for (int r = 0; r < R; r += 8) {
                    uint4 q;
                    q.x = (uint16_t)pcm16(o[r]) | ((uint32_t)(uint16_t)pcm16(o[r + 1]) << 16);
                    q.y = (uint16_t)pcm16(o[r + 2]) | ((uint32_t)(uint16_t)pcm16(o[r + 3]) << 16);
                    q.z = (uint16_t)pcm16(o[r + 4]) | ((uint32_t)(uint16_t)pcm16(o[r + 5]) << 16);
                    q.w = (uint16_t)pcm16(o[r + 6]) | ((uint32_t)(uint16_t)pcm16(o[r + 7]) << 16);
                    d[r / 8] = q;
                }
            }
        } else {
            for (int r = 0; r < R; r++) if (r < left) {
                if (audio) audio[(long long)s * audio_stride + m + r] = o[r];
                if (pcm) pcm[(long long)s * pcm_stride + m + r] = pcm16(o[r]);
            }
        }
    }
}

// ---- U > 1: one thread per output, tile of NT outputs per block ---------------------------------------
// The span of input the tile needs is staged in shared memory as (delayed IF, mixed) pairs, and so is the whole
// transposed tap table [tap j][phase] (101 x up_pad floats, 59 KB for U = 147): consecutive outputs differ in phase by
// D mod U, so a warp's tap reads fall on distinct banks, where the same reads through the read-only path cost several
// L1 wavefronts each.  Both filters of an output advance with one packed FFMA2 per tap (fused: the PLL never sees
// this kernel's output); EXACT keeps them unfused.
// Lane mapping: consecutive outputs read inputs D/U = 5.44 (8.71) samples apart, a stride that piles a warp's 8-byte
// sample reads onto a few banks.  Lane l of a warp therefore takes every K-th output (K = 9 for 147/800, 7 for
// 147/1280: K*D/U is within 0.05 of an ODD integer, 49 and 61), K warps interleave to cover 32*K consecutive outputs,
// and both the sample reads and the tap reads (phase stride K*D mod U = -3, -7) are bank-conflict free.
template <int NTMAX, bool EXACT, bool STEREO>
__global__ void __launch_bounds__(NTMAX)
k_audio_poly(const float* __restrict__ if_in, long long if_stride, const float* __restrict__ if_tail, long long tail_stride,
             const float* __restrict__ nco, const float* __restrict__ sband, long long bb_stride,
             const float* __restrict__ mix_tail, float* __restrict__ audio, long long audio_stride,
             int16_t* __restrict__ pcm, long long pcm_stride, int n_if, int n_audio, int up, int down,
             const float* __restrict__ taps_poly, int up_pad, int span_max, u64 nz, int K, int tiles_per_stream, int n_tiles)
{
    // Persistent CTAs: the 59 KB tap table is loaded ONCE per CTA and then tiles (stream, NT consecutive outputs) are
    // walked in a grid-stride loop — re-loading the table per tile moved more bytes through L2 than the signal itself.
    const int NT = blockDim.x;                                              // 32 * K * groups outputs per tile
    extern __shared__ __align__(16) float sm_f[];
    constexpr int DELAY = DY4_NTAPS / 2;
    float* s_taps = sm_f;                                                   // [DY4_NTAPS][up_pad]
    float2* s_x = reinterpret_cast<float2*>(sm_f + DY4_NTAPS * up_pad);     // (delayed IF, mixed) over the tile's span
    const int tid = threadIdx.x;
    for (int t = tid; t < DY4_NTAPS * up_pad / 4; t += NT)                  // up_pad is a multiple of 4
        reinterpret_cast<float4*>(s_taps)[t] = __ldg(reinterpret_cast<const float4*>(taps_poly) + t);
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int s = tile / tiles_per_stream;
    const int m0 = (tile - s * tiles_per_stream) * NT;
    __syncthreads();                                                        // table visible; the previous tile's span no longer read
    const float* row = if_in + (long long)s * if_stride;
    const float* itail = if_tail + (long long)s * tail_stride;
    const float* nrow = STEREO ? nco + (long long)s * bb_stride : nullptr;
    const float* srow = STEREO ? sband + (long long)s * bb_stride : nullptr;
    const float* mtail = STEREO ? mix_tail + (long long)s * DY4_MIX_TAIL : nullptr;

    const int m_last = min(m0 + NT, n_audio) - 1;
    const int i_lo = ((int)(((long long)m0 * down) / up) - (DY4_NTAPS - 1)) & ~3;   // aligned down: 16-byte loads below
    const int i_hi = (int)(((long long)m_last * down) / up);
    const int span = i_hi - i_lo + 1;
    for (int p = 4 * tid; p < span; p += 4 * NT) {
        const int i = i_lo + p;
        float x0, x1, x2, x3, y0 = 0.f, y1 = 0.f, y2 = 0.f, y3 = 0.f;
        if (i - DELAY >= 0 && i + 3 < n_if) {
            const float2 a = __ldg(reinterpret_cast<const float2*>(row + i - DELAY));
            const float2 b = __ldg(reinterpret_cast<const float2*>(row + i - DELAY + 2));
            x0 = a.x; x1 = a.y; x2 = b.x; x3 = b.y;
            if (STEREO) {
                const float4 nv = __ldg(reinterpret_cast<const float4*>(nrow + i));
                const float4 sv = __ldg(reinterpret_cast<const float4*>(srow + i));
                y0 = __fmul_rn(__fmul_rn(nv.x, sv.x), 2.0f); y1 = __fmul_rn(__fmul_rn(nv.y, sv.y), 2.0f);     // filter.cpp:264
                y2 = __fmul_rn(__fmul_rn(nv.z, sv.z), 2.0f); y3 = __fmul_rn(__fmul_rn(nv.w, sv.w), 2.0f);
            }
        } else {
            x0 = if_at(row, itail, n_if, i - DELAY); x1 = if_at(row, itail, n_if, i + 1 - DELAY);
            x2 = if_at(row, itail, n_if, i + 2 - DELAY); x3 = if_at(row, itail, n_if, i + 3 - DELAY);
            if (STEREO) {
                y0 = mix_at(nrow, srow, mtail, n_if, i); y1 = mix_at(nrow, srow, mtail, n_if, i + 1);
                y2 = mix_at(nrow, srow, mtail, n_if, i + 2); y3 = mix_at(nrow, srow, mtail, n_if, i + 3);
            }
        }
        float4* d = reinterpret_cast<float4*>(s_x + p);
        d[0] = make_float4(x0, y0, x1, y1);
        d[1] = make_float4(x2, y2, x3, y3);
    }
    __syncthreads();

    const int warp = tid >> 5, lane = tid & 31;
    const int m = m0 + 32 * K * (warp / K) + lane * K + (warp % K);
    if (m >= n_audio) continue;
    const long long n = (long long)m * down;
    const int phase = (int)(n % up);
    const int base = (int)(n / up) - i_lo;          // position of x[floor(mD/U)] in the span
    const u64* xs = reinterpret_cast<const u64*>(s_x) + base;
    const float* hp = s_taps + phase;
    u64 acc = 0ull;
#pragma unroll 4
    for (int j = 0; j < DY4_NTAPS; j++) {
        const float h = hp[j * up_pad];
        acc = tap2<EXACT>(acc, xs[-j], pk2(h, h), nz);
    }
    float mono, diff;
    upk2(acc, mono, diff);
    if (STEREO) {
        const float l = __fadd_rn(mono, diff), r = __fsub_rn(mono, diff);
        if (audio) *reinterpret_cast<float2*>(audio + (long long)s * audio_stride + 2LL * m) = make_float2(l, r);
        if (pcm) *reinterpret_cast<uint32_t*>(pcm + (long long)s * pcm_stride + 2LL * m) = (uint16_t)pcm16(l) | ((uint32_t)(uint16_t)pcm16(r) << 16);
    } else {
        if (audio) audio[(long long)s * audio_stride + m] = mono;
        if (pcm) pcm[(long long)s * pcm_stride + m] = pcm16(mono);
    }
    }   // tile loop
}

// ---- U > 1, tap-stationary (the default) -------------------------------------------------------------------------
// The kernel above reads TWO things from shared memory per multiply-add pair — a sample pair and a tap — and is bound by those
// wavefronts.  But an output's phase, (m D) mod U, repeats every U outputs: a thread that only ever computes outputs m, m + S,
// m + 2S, ... (S a multiple of U) needs ONE set of 101 taps, and keeps it in registers for the lifetime of a persistent CTA.
// Shared memory then serves nothing but samples: one 8-byte (delayed IF, mixed) read per tap for both filters of an output.
// Tile = P groups of S consecutive outputs of one stream; thread <-> slot in [0, S), visited P times; the span of input the tile
// needs (S P D/U + 100 samples) is staged once.  Slots are dealt to lanes with the same K-interleave as above (lane l of a
// warp takes every K-th output, K D/U within 0.05 of an odd integer), so the 8-byte reads of a half-warp fall on distinct bank
// pairs; slots beyond 32 K G sit in one extra, partly filled warp.
template <int NTMAX, bool EXACT, bool STEREO>
__global__ void __launch_bounds__(NTMAX, 1)
k_audio_poly_ts(const float* __restrict__ if_in, long long if_stride, const float* __restrict__ if_tail, long long tail_stride,
                const float* __restrict__ nco, const float* __restrict__ sband, long long bb_stride,
                const float* __restrict__ mix_tail, float* __restrict__ audio, long long audio_stride,
                int16_t* __restrict__ pcm, long long pcm_stride, int n_if, int n_audio, int up, int down,
                const float* __restrict__ taps_poly, int up_pad, int K, int KG, int S, int P, int tiles_per_stream, int n_tiles)
{
    const int NT = blockDim.x;
    extern __shared__ __align__(16) float sm_f[];
    constexpr int DELAY = DY4_NTAPS / 2;
    float2* s_x = reinterpret_cast<float2*>(sm_f);                          // (delayed IF, mixed) over the tile's span
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int slot = warp < KG ? 32 * K * (warp / K) + lane * K + (warp % K) : 32 * KG + (tid - 32 * KG);
    const bool live = slot < S;
    // this thread's phase and taps: (slot D) mod U, h[phase + j U] (taps_poly is the transposed table [tap j][phase])
    float h[DY4_NTAPS];
    {
        const int phase = live ? (int)(((long long)slot * down) % up) : 0;
#pragma unroll
        for (int j = 0; j < DY4_NTAPS; j++) h[j] = live ? __ldg(taps_poly + (long long)j * up_pad + phase) : 0.0f;
    }
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int s = tile / tiles_per_stream;
        const int m0 = (tile - s * tiles_per_stream) * S * P;
        __syncthreads();                                                    // the previous tile's span is no longer read
        const float* row = if_in + (long long)s * if_stride;
        const float* itail = if_tail + (long long)s * tail_stride;
        const float* nrow = STEREO ? nco + (long long)s * bb_stride : nullptr;
        const float* srow = STEREO ? sband + (long long)s * bb_stride : nullptr;
        const float* mtail = STEREO ? mix_tail + (long long)s * DY4_MIX_TAIL : nullptr;
        const int m_last = min(m0 + S * P, n_audio) - 1;
        const int i_lo = ((int)(((long long)m0 * down) / up) - (DY4_NTAPS - 1)) & ~3;   // aligned down: 16-byte loads below
        const int i_hi = (int)(((long long)m_last * down) / up);
        const int span = i_hi - i_lo + 1;
        for (int p = 4 * tid; p < span; p += 4 * NT) {
            const int i = i_lo + p;
            float x0, x1, x2, x3, y0 = 0.f, y1 = 0.f, y2 = 0.f, y3 = 0.f;
            if (i - DELAY >= 0 && i + 3 < n_if) {
                const float2 a = __ldg(reinterpret_cast<const float2*>(row + i - DELAY));
                const float2 b = __ldg(reinterpret_cast<const float2*>(row + i - DELAY + 2));
                x0 = a.x; x1 = a.y; x2 = b.x; x3 = b.y;
                if (STEREO) {
                    const float4 nv = __ldg(reinterpret_cast<const float4*>(nrow + i));
                    const float4 sv = __ldg(reinterpret_cast<const float4*>(srow + i));
                    y0 = __fmul_rn(__fmul_rn(nv.x, sv.x), 2.0f); y1 = __fmul_rn(__fmul_rn(nv.y, sv.y), 2.0f);     // filter.cpp:264
                    y2 = __fmul_rn(__fmul_rn(nv.z, sv.z), 2.0f); y3 = __fmul_rn(__fmul_rn(nv.w, sv.w), 2.0f);
                }
            } else {
                x0 = if_at(row, itail, n_if, i - DELAY); x1 = if_at(row, itail, n_if, i + 1 - DELAY);
                x2 = if_at(row, itail, n_if, i + 2 - DELAY); x3 = if_at(row, itail, n_if, i + 3 - DELAY);
                if (STEREO) {
                    y0 = mix_at(nrow, srow, mtail, n_if, i); y1 = mix_at(nrow, srow, mtail, n_if, i + 1);
                    y2 = mix_at(nrow, srow, mtail, n_if, i + 2); y3 = mix_at(nrow, srow, mtail, n_if, i + 3);
                }
            }
            float4* d = reinterpret_cast<float4*>(s_x + p);
            d[0] = make_float4(x0, y0, x1, y1);
            d[1] = make_float4(x2, y2, x3, y3);
        }
        __syncthreads();
        if (!live) continue;
        for (int q = 0; q < P; q++) {
            const int m = m0 + q * S + slot;
            if (m >= n_audio) break;
            const int base = (int)(((long long)m * down) / up) - i_lo;      // position of x[floor(m D/U)] in the span
            const float2* xs = s_x + base;
            float mono = 0.0f, diff = 0.0f;
#pragma unroll
            for (int j = 0; j < DY4_NTAPS; j++) {                           // ascending taps, as filter.cpp:159-165
                const float2 v = xs[-j];
                if (EXACT) { mono = __fadd_rn(mono, __fmul_rn(h[j], v.x)); if (STEREO) diff = __fadd_rn(diff, __fmul_rn(h[j], v.y)); }
                else { mono = fmaf(h[j], v.x, mono); if (STEREO) diff = fmaf(h[j], v.y, diff); }
            }
            if (STEREO) {
                const float l = __fadd_rn(mono, diff), r = __fsub_rn(mono, diff);
                if (audio) *reinterpret_cast<float2*>(audio + (long long)s * audio_stride + 2LL * m) = make_float2(l, r);
                if (pcm) *reinterpret_cast<uint32_t*>(pcm + (long long)s * pcm_stride + 2LL * m) = (uint16_t)pcm16(l) | ((uint32_t)(uint16_t)pcm16(r) << 16);
            } else {
                if (audio) audio[(long long)s * audio_stride + m] = mono;
                if (pcm) pcm[(long long)s * pcm_stride + m] = pcm16(mono);
            }
        }
    }
}

template <int D, int R, int NT, bool EXACT, bool STEREO>
cudaError_t launch_u1(const Dy4AudioArgs& a, cudaStream_t st)
{
    constexpr int T = NT * R;
    const size_t smem = sizeof(float2) * dy4_padded_pairs(D, R, NT);
    auto kern = k_audio_u1<D, R, NT, EXACT, STEREO>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    dim3 grid(a.n_streams, (a.n_audio + T - 1) / T);
    kern<<<grid, NT, smem, st>>>(a.if_in, a.if_stride, a.if_tail, a.if_tail_stride ? a.if_tail_stride : (long long)DY4_IF_TAIL, a.nco, a.sband, a.bb_stride, a.mix_tail, a.audio, a.audio_stride,
                                 a.pcm, a.pcm_stride, a.n_if, a.n_audio, a.neg_zero2, a.mode);
    g_dy4_launches++;
    return cudaGetLastError();
}

template <int D>
cudaError_t dispatch_u1(const Dy4AudioArgs& a, cudaStream_t st)
{
    constexpr int R = 8, NT = 128;
    if (a.stereo) return a.exact ? launch_u1<D, R, NT, true, true>(a, st) : launch_u1<D, R, NT, false, true>(a, st);
    return a.exact ? launch_u1<D, R, NT, true, false>(a, st) : launch_u1<D, R, NT, false, false>(a, st);
}

template <bool EXACT, bool STEREO>
cudaError_t launch_poly(const Dy4AudioArgs& a, cudaStream_t st)
{
    constexpr int NTMAX = 640;
    if (a.up_pad % 4) return cudaErrorInvalidValue;
    // K: the output interleave that minimises shared-memory bank conflicts (see the kernel's header), found by counting
    // them for a few sample warps: 8-byte sample reads are served per half-warp over 16 bank pairs, tap reads over 32 banks
    static int k_up = 0, k_down = 0, k_best = 1;
    if (k_up != a.up || k_down != a.down) {
        long long best = -1;
        for (int k = 1; k <= 10; k++) {
            long long cost = 0;
            for (int w = 0; w < 64; w++) {
                int cx[2][16] = {}, ct[32] = {};
                for (int l = 0; l < 32; l++) {
                    const long long n = (long long)(w * 37 + 5 + l * k) * a.down;
                    cx[l >> 4][(n / a.up) & 15]++;
                    ct[(n % a.up) & 31]++;
                }
                int m0 = 0, m1 = 0, mt = 0;
                for (int i = 0; i < 16; i++) { m0 = std::max(m0, cx[0][i]); m1 = std::max(m1, cx[1][i]); }
                for (int i = 0; i < 32; i++) mt = std::max(mt, ct[i]);
                cost += m0 + m1 + mt;
            }
            if (best < 0 || cost < best) { best = cost; k_best = k; }
        }
        k_up = a.up; k_down = a.down;
    }
    const int K = k_best;
    // measured (1 024 streams x 12 blocks, whole-job launch): 147/800 1.07 ms against 1.15 ms for the table-in-shared-memory kernel;
    // 147/1280 1.52 against 1.33 ms (its tile holds one group only: staging is not overlapped with anything) — so each ratio takes its best
    const bool prefer_ts = a.poly_variant == 0 ? (long long)a.down * 100 < (long long)a.up * 700 : a.poly_variant == 2;
    if (a.n_if % a.down == 0 && (a.n_audio % a.up) == 0 && prefer_ts) {
        // tap-stationary kernel: whole periods per chunk (every chunk of whole blocks is: a block is 10 periods in both modes)
        constexpr int NTS = 512;
        int best_L = 0, best_G = 0, best_extra = 0, best_nt = 1 << 30;
        for (int L = 1; L <= 4; L++) {
            const int Sx = a.up * L;
            const int G = Sx / (32 * K), rest = Sx - 32 * K * G;
            const int extra = rest > 0 ? (rest <= 32 ? 1 : -1) : 0;
            const int nt = extra < 0 ? 32 * K * (G + 1) : 32 * (K * G + extra);
            const int g_eff = extra < 0 ? G + 1 : G;
            if (nt > NTS) continue;
            // waste = idle threads per output slot
            if (best_L == 0 || (long long)(nt - Sx) * (a.up * best_L) < (long long)(best_nt - a.up * best_L) * Sx) { best_L = L; best_G = g_eff; best_extra = extra > 0; best_nt = nt; }
        }
        if (best_L) {
            const int S = a.up * best_L, KG = K * best_G, NT = best_nt;
            int P = std::max(1, (int)(6400LL * a.up / ((long long)S * a.down)));          // ~6 400 staged samples per tile (51 KB)
            const int span_max = ((int)(((long long)(S * P) * a.down) / a.up) + DY4_NTAPS + 3 + 4 + 7) & ~3;
            const size_t smem = sizeof(float2) * (size_t)span_max;
            auto kern = k_audio_poly_ts<NTS, EXACT, STEREO>;
            cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return e;
            static int sms = 0;
            if (!sms) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev); }
            int per_sm = 1;
            e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, NT, smem);
            if (e != cudaSuccess) return e;
            const int tiles_per_stream = (a.n_audio + S * P - 1) / (S * P);
            const long long n_tiles = (long long)tiles_per_stream * a.n_streams;
            if (n_tiles > 0x7fffffffLL) return cudaErrorInvalidValue;
            const int grid = (int)std::min<long long>(n_tiles, (long long)sms * std::max(per_sm, 1));
            kern<<<grid, NT, smem, st>>>(a.if_in, a.if_stride, a.if_tail, a.if_tail_stride ? a.if_tail_stride : (long long)DY4_IF_TAIL, a.nco, a.sband, a.bb_stride, a.mix_tail, a.audio, a.audio_stride, a.pcm, a.pcm_stride,
                                         a.n_if, a.n_audio, a.up, a.down, a.taps_poly, a.up_pad, K, KG, S, P, tiles_per_stream, (int)n_tiles);
            g_dy4_launches++;
            return cudaGetLastError();
        }
    }
    const int NT = 32 * K * 2;                         // one 59 KB tap table per CTA, amortised over 64*K outputs
    const int span_max = ((int)(((long long)(NT - 1) * a.down) / a.up) + DY4_NTAPS + 3 + 4 + 7) & ~3;
    const size_t smem = sizeof(float) * ((size_t)DY4_NTAPS * a.up_pad + 2 * (size_t)span_max);
    auto kern = k_audio_poly<NTMAX, EXACT, STEREO>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    static int sms = 0;
    if (!sms) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev); }
    int per_sm = 1;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, NT, smem);
    if (e != cudaSuccess) return e;
    const int tiles_per_stream = (a.n_audio + NT - 1) / NT;
    const long long n_tiles = (long long)tiles_per_stream * a.n_streams;
    if (n_tiles > 0x7fffffffLL) return cudaErrorInvalidValue;
    const int grid = (int)std::min<long long>(n_tiles, (long long)sms * std::max(per_sm, 1));
    kern<<<grid, NT, smem, st>>>(a.if_in, a.if_stride, a.if_tail, a.if_tail_stride ? a.if_tail_stride : (long long)DY4_IF_TAIL, a.nco, a.sband, a.bb_stride, a.mix_tail,
                                 a.audio, a.audio_stride, a.pcm, a.pcm_stride, a.n_if, a.n_audio,
                                 a.up, a.down, a.taps_poly, a.up_pad, span_max, a.neg_zero2, K, tiles_per_stream, (int)n_tiles);
    g_dy4_launches++;
    return cudaGetLastError();
}

}  // namespace

cudaError_t dy4_launch_audio(const Dy4AudioArgs& a, cudaStream_t st)
{
    if (a.n_audio <= 0 || a.n_streams <= 0) return cudaSuccess;
    if (a.up == 1 && a.down == 5) return dispatch_u1<5>(a, st);
    if (a.up == 1 && a.down == 8) return dispatch_u1<8>(a, st);
    if (a.up > 1) {
        if (a.stereo) return a.exact ? launch_poly<true, true>(a, st) : launch_poly<false, true>(a, st);
        return a.exact ? launch_poly<true, false>(a, st) : launch_poly<false, false>(a, st);
    }
    return cudaErrorInvalidValue;
}

cudaError_t dy4_upload_taps_audio(const TapPairs* audio4) { return cudaMemcpyToSymbol(c_audio2, audio4, sizeof(TapPairs) * 4); }
