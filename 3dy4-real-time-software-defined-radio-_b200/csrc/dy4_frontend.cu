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

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <string>

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
    const int m0 = blockIdx.y * T;                    // streams on grid.x (no 65535 limit), tiles on grid.y
    const uint8_t* row = iq + (long long)blockIdx.x * row_stride;
    const uint8_t* tail = iq_tail + (long long)blockIdx.x * DY4_IQ_TAIL;
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
    float* dst = if_out + (long long)blockIdx.x * if_stride + m0 + tid * R;
    const int left = n_if - (m0 + tid * R);
    if (left >= R) {
#pragma unroll
        for (int r = 0; r < R; r += 4) *reinterpret_cast<float4*>(dst + r) = make_float4(out[r], out[r + 1], out[r + 2], out[r + 3]);
    } else {
#pragma unroll
        for (int r = 0; r < R; r++) if (r < left) dst[r] = out[r];
    }
}


// ------------------------------------------------------------------------------------------------------------
// TMA-staged persistent variant (default).  The tile's raw bytes (history + 2*D*T bytes, 20.7 KB for D=10) are
// brought into shared memory by ONE bulk asynchronous copy (cp.async.bulk -> UBLKCP in SASS) that completes on
// an mbarrier; two buffers, so the copy for tile i+2 is issued as soon as tile i has been filtered and lands
// while tile i+1 is being filtered.  No staging pass at all: the FIR reads the packed uint8 pairs straight from
// shared memory with 128-bit loads (8 IQ samples each) and unpacks in registers — two byte-permutes build the
// floats 2^23+b, one packed add removes 2^23+128 — while the 1/128 of (b-128)/128 is folded into the taps (a power
// of two: products and roundings are unchanged).  Each CTA walks tiles t = blockIdx.x, blockIdx.x + gridDim.x, ...
// (stream-major tile order), grid = SMs x resident CTAs.
// Variants measured on a B200 and not kept: three stages with mbarrier-only hand-offs between
// warps (register-capped, spills), one tile per CTA with per-output tap tables (constant-cache thrash).
// ------------------------------------------------------------------------------------------------------------
__constant__ TapPairs c_rf2s[4];   // (h/128, h/128) pairs of the RF low-pass, per mode
// the same values as compile-time constants (immediate operands): kRfTapsScaled[mode][k], and a host copy for the
// start-up check that they are what dy4_lpf_taps() designs on this machine
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

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, unsigned long long* bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WAIT_DONE;\n"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}

// (I,Q) of the IQ sample held in the 16-bit half `h` (0 or 1) of word w, as exact integers b-128 in a packed pair
__device__ __forceinline__ u64 unpack_iq(uint32_t w, int h, u64 neg_bias)
{
    const float fi = __uint_as_float(__byte_perm(w, 0x4B000000u, h ? 0x7442 : 0x7440));   // 2^23 + I byte
    const float fq = __uint_as_float(__byte_perm(w, 0x4B000000u, h ? 0x7443 : 0x7441));   // 2^23 + Q byte
    return fadd2(pk2(fi, fq), neg_bias);                                                  // - (2^23 + 128): exact
}

template <int D, int R, int NT, bool EXACT>
__global__ void __launch_bounds__(NT, 3)
k_frontend_tma(const uint8_t* __restrict__ iq, long long row_stride, const uint8_t* __restrict__ iq_tail,
               float* __restrict__ if_out, long long if_stride, int n_if, u64 nz, int mode,
               int tiles_per_stream, int n_tiles)
{
    constexpr int T = NT * R;                          // IF samples per tile
    constexpr int NW = NT / 32;
    constexpr int CH = D * R;                          // input samples per thread
    constexpr int HALO = DY4_IQ_TAIL / 2;              // 112 samples of history in front of the tile
    constexpr int TILE_BYTES = 2 * D * T + DY4_IQ_TAIL;
    constexpr int BUF_BYTES = (TILE_BYTES + 64 + 127) / 128 * 128;   // +64: the last thread's final 16-byte load runs past the tile
    constexpr int QMAX = D * (R - 1) + (DY4_NTAPS - 1);
    constexpr int C0 = HALO - (DY4_NTAPS - 1);         // 12: first window sample of thread 0
    constexpr int CA = C0 & ~7;                        // window start aligned down to a 16-byte group (8 samples)
    constexpr int NG = (C0 - CA + QMAX) / 8 + 1;       // 16-byte groups per thread window
    static_assert(TILE_BYTES % 16 == 0 && (2 * CH) % 16 == 0, "bulk copies and window loads are 16-byte granular");
    extern __shared__ __align__(128) uint8_t smraw[];
    __shared__ __align__(8) unsigned long long mbar[2];
    __shared__ float2 s_last[NT];
    __shared__ float2 s_prev;
    const int tid = threadIdx.x;
    const long long row_bytes = 2LL * D * n_if;
    const u64* hh = reinterpret_cast<const u64*>(c_rf2s[mode].t);
    const u64 neg_bias = pk2(-8388736.0f, -8388736.0f);

    auto issue = [&](int t, int b) {                   // thread 0: start the bulk copy of tile t into buffer b
        const int s = t / tiles_per_stream, m0 = (t - s * tiles_per_stream) * T;
        const uint8_t* row = iq + (long long)s * row_stride;
        uint8_t* dst = smraw + b * BUF_BYTES;
        const long long start = 2LL * D * m0 - DY4_IQ_TAIL;
        const long long begin = start < 0 ? 0 : start;
        long long len = (start + TILE_BYTES) - begin;
        if (begin + len > row_bytes) len = row_bytes - begin;
        if (len < 0) len = 0;
        const uint32_t head = start < 0 ? (uint32_t)DY4_IQ_TAIL : 0u;   // first tile of the chunk: history from the carried tail
        mbar_expect_tx(&mbar[b], head + (uint32_t)len);
        if (head) bulk_g2s(dst, iq_tail + (long long)s * DY4_IQ_TAIL, head, &mbar[b]);
        if (len > 0) bulk_g2s(dst + head, row + begin, (uint32_t)len, &mbar[b]);
    };

    if (tid == 0) {
        mbar_init(&mbar[0], 1);
        mbar_init(&mbar[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const int first = blockIdx.x, stride = gridDim.x;
    if (tid == 0) {
        if (first < n_tiles) issue(first, 0);
        if (first + stride < n_tiles) issue(first + stride, 1);
    }

    int it = 0;
    for (int t = first; t < n_tiles; t += stride, it++) {
        const int b = it & 1;
        const int s = t / tiles_per_stream, m0 = (t - s * tiles_per_stream) * T;
        mbar_wait(&mbar[b], (it >> 1) & 1);
        const uint8_t* buf = smraw + b * BUF_BYTES;

        // ---- FIR: R consecutive (I,Q) outputs per thread; window index descends so taps ascend ------------
        u64 acc[R];
#pragma unroll
        for (int r = 0; r < R; r++) acc[r] = 0ull;
        const uint4* win = reinterpret_cast<const uint4*>(buf + 2 * (CH * tid + CA));
#pragma unroll
        for (int g = NG - 1; g >= 0; g--) {
            const uint4 v = win[g];
            const uint32_t ws[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int j = 7; j >= 0; j--) {
                const int q = 8 * g + j - (C0 - CA);
                if (q < 0 || q > QMAX) continue;
                const u64 x = unpack_iq(ws[j >> 1], j & 1, neg_bias);
#pragma unroll
                for (int r = 0; r < R; r++) {
                    const int k = D * r + (DY4_NTAPS - 1) - q;
                    if (k >= 0 && k < DY4_NTAPS) acc[r] = tap2<EXACT>(acc[r], x, hh[k], nz);
                }
            }
        }

        // ---- the output before the tile (m0-1), for the discriminator's first difference: 101 extra taps on one
        // lane; the warp that pays rotates with the tile so no scheduler is always the slow one.
        float pI, pQ;
        upk2(acc[R - 1], pI, pQ);
        s_last[tid] = make_float2(pI, pQ);
        if (tid == 32 * (it % NW)) {
            u64 a0 = 0ull;
            const uint4* w0 = reinterpret_cast<const uint4*>(buf);
            constexpr int PMAX = HALO - D;                     // tap k reads sample PMAX - k
#pragma unroll
            for (int g = PMAX / 8; g >= 0; g--) {
                const uint4 v = w0[g];
                const uint32_t ws[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                for (int j = 7; j >= 0; j--) {
                    const int k = PMAX - (8 * g + j);
                    if (k >= 0 && k < DY4_NTAPS) a0 = tap2<EXACT>(a0, unpack_iq(ws[j >> 1], j & 1, neg_bias), hh[k], nz);
                }
            }
            float eI, eQ;
            upk2(a0, eI, eQ);
            s_prev = make_float2(eI, eQ);
        }
        __syncthreads();                                       // edge values visible; every read of buffer b is done
        if (tid == 0 && t + 2 * stride < n_tiles) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            issue(t + 2 * stride, b);
        }
        { const float2 l = tid > 0 ? s_last[tid - 1] : s_prev; pI = l.x; pQ = l.y; }

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
        float* dst = if_out + (long long)s * if_stride + m0 + tid * R;
        const int left = n_if - (m0 + tid * R);
        if (left >= R) {
#pragma unroll
            for (int r = 0; r < R; r += 4) *reinterpret_cast<float4*>(dst + r) = make_float4(out[r], out[r + 1], out[r + 2], out[r + 3]);
        } else {
#pragma unroll
            for (int r = 0; r < R; r++) if (r < left) dst[r] = out[r];
        }
        __syncthreads();                                       // the edge slots are rewritten by the next tile
    }
}

// ------------------------------------------------------------------------------------------------------------
// v3: same staging (one bulk copy per tile, two buffers), three changes aimed at the FMA pipe's idle time:
//  * every CTA walks a CONTIGUOUS range of tiles, so the discriminator's "sample before the tile" is simply the
//    last output of the CTA's previous tile (kept in shared memory); the 101 extra taps on one lane are paid only at
//    the start of a range or of a stream, and the second barrier per tile goes away (edge slots double-buffered);
//  * the -0 addend of the unfused product lives in a VECTOR register (read back from shared memory), which leaves
//    the uniform operand slot of FFMA2 to the taps: ptxas then streams them through uniform registers (LDCU)
//    instead of parking them in ~100 vector registers, and more CTAs fit per SM;
//  * UNPACK = 1 converts the bytes with I2F.S8 (byte select in the instruction, after one XOR 0x80808080 per
//    word), which takes the bias subtraction off the FMA pipe.
// ------------------------------------------------------------------------------------------------------------
template <int UNPACK>
__device__ __forceinline__ void unpack_word(uint32_t w, u64 neg_bias, u64& x0, u64& x1)
{
    if (UNPACK == 1) {
        const uint32_t sgn = w ^ 0x80808080u;                                  // bytes become two's-complement b-128
        x0 = pk2((float)(int8_t)(sgn & 0xffu), (float)(int8_t)((sgn >> 8) & 0xffu));
        x1 = pk2((float)(int8_t)((sgn >> 16) & 0xffu), (float)(int8_t)(sgn >> 24));
    } else {
        x0 = unpack_iq(w, 0, neg_bias);
        x1 = unpack_iq(w, 1, neg_bias);
    }
}

template <int MODE, int D, int R, int NT, bool EXACT, int UNPACK, int MINB>
__global__ void __launch_bounds__(NT, MINB)
k_frontend_v3(const uint8_t* __restrict__ iq, long long row_stride, const uint8_t* __restrict__ iq_tail,
              float* __restrict__ if_out, long long if_stride, int n_if, u64 nz,
              int tiles_per_stream, int n_tiles)
{
    constexpr int T = NT * R;
    constexpr int CH = D * R;
    constexpr int HALO = DY4_IQ_TAIL / 2;
    constexpr int TILE_BYTES = 2 * D * T + DY4_IQ_TAIL;
    constexpr int BUF_BYTES = (TILE_BYTES + 64 + 127) / 128 * 128;
    constexpr int QMAX = D * (R - 1) + (DY4_NTAPS - 1);
    constexpr int C0 = HALO - (DY4_NTAPS - 1);
    constexpr int CA = C0 & ~7;
    constexpr int NG = (C0 - CA + QMAX) / 8 + 1;
    static_assert(TILE_BYTES % 16 == 0 && (2 * CH) % 16 == 0, "bulk copies and window loads are 16-byte granular");
    extern __shared__ __align__(128) uint8_t smraw[];
    __shared__ __align__(8) unsigned long long mbar[2];
    __shared__ float2 s_last[2][NT];
    __shared__ u64 s_nz[NT];
    const int tid = threadIdx.x;
    const long long row_bytes = 2LL * D * n_if;
    const u64 neg_bias = pk2(-8388736.0f, -8388736.0f);

    auto issue = [&](int t, int b) {
        const int s = t / tiles_per_stream, m0 = (t - s * tiles_per_stream) * T;
        const uint8_t* row = iq + (long long)s * row_stride;
        uint8_t* dst = smraw + b * BUF_BYTES;
        const long long start = 2LL * D * m0 - DY4_IQ_TAIL;
        const long long begin = start < 0 ? 0 : start;
        long long len = (start + TILE_BYTES) - begin;
        if (begin + len > row_bytes) len = row_bytes - begin;
        if (len < 0) len = 0;
        const uint32_t head = start < 0 ? (uint32_t)DY4_IQ_TAIL : 0u;
        mbar_expect_tx(&mbar[b], head + (uint32_t)len);
        if (head) bulk_g2s(dst, iq_tail + (long long)s * DY4_IQ_TAIL, head, &mbar[b]);
        if (len > 0) bulk_g2s(dst + head, row + begin, (uint32_t)len, &mbar[b]);
    };

    const int t_begin = (int)((long long)n_tiles * blockIdx.x / gridDim.x);
    const int t_end = (int)((long long)n_tiles * (blockIdx.x + 1) / gridDim.x);
    if (tid == 0) {
        mbar_init(&mbar[0], 1);
        mbar_init(&mbar[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    s_nz[tid] = nz;
    __syncthreads();
    if (tid == 0) {
        if (t_begin < t_end) issue(t_begin, 0);
        if (t_begin + 1 < t_end) issue(t_begin + 1, 1);
    }
    const u64 nzv = *reinterpret_cast<volatile u64*>(&s_nz[tid]);  // (-0,-0) in a vector register: see the header comment

    int it = 0;
    for (int t = t_begin; t < t_end; t++, it++) {
        const int b = it & 1;
        const int s = t / tiles_per_stream, m0 = (t - s * tiles_per_stream) * T;
        mbar_wait(&mbar[b], (it >> 1) & 1);
        const uint8_t* buf = smraw + b * BUF_BYTES;

        u64 acc[R];
#pragma unroll
        for (int r = 0; r < R; r++) acc[r] = 0ull;
        const uint4* win = reinterpret_cast<const uint4*>(buf + 2 * (CH * tid + CA));
#pragma unroll
        for (int g = NG - 1; g >= 0; g--) {
            const uint4 v = win[g];
            const uint32_t ws[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int w = 3; w >= 0; w--) {
                if (8 * g + 2 * w + 1 - (C0 - CA) < 0 || 8 * g + 2 * w - (C0 - CA) > QMAX) continue;
                u64 xs[2];
                unpack_word<UNPACK>(ws[w], neg_bias, xs[0], xs[1]);
#pragma unroll
                for (int h = 1; h >= 0; h--) {
                    const int q = 8 * g + 2 * w + h - (C0 - CA);
                    if (q < 0 || q > QMAX) continue;
#pragma unroll
                    for (int r = 0; r < R; r++) {
                        const int k = D * r + (DY4_NTAPS - 1) - q;
                        if (k >= 0 && k < DY4_NTAPS) acc[r] = tap2<EXACT>(acc[r], xs[h], pk2(kRfTapsScaled[MODE][k], kRfTapsScaled[MODE][k]), nzv);
                    }
                }
            }
        }

        float pI, pQ;
        upk2(acc[R - 1], pI, pQ);
        s_last[b][tid] = make_float2(pI, pQ);
        const bool edge = it == 0 || m0 == 0;                  // no previous tile of this stream in this CTA
        if (edge && tid == 0) {                                // output m0-1 from the history in front of the tile
            u64 a0 = 0ull;
            const uint4* w0 = reinterpret_cast<const uint4*>(buf);
            constexpr int PMAX = HALO - D;
#pragma unroll
            for (int g = PMAX / 8; g >= 0; g--) {
                const uint4 v = w0[g];
                const uint32_t ws[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                for (int j = 7; j >= 0; j--) {
                    const int k = PMAX - (8 * g + j);
                    if (k >= 0 && k < DY4_NTAPS) a0 = tap2<EXACT>(a0, unpack_iq(ws[j >> 1], j & 1, neg_bias), pk2(kRfTapsScaled[MODE][k], kRfTapsScaled[MODE][k]), nzv);
                }
            }
            upk2(a0, pI, pQ);
        }
        __syncthreads();                                       // edge slots visible; every read of buffer b is done
        if (tid == 0 && t + 2 < t_end) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            issue(t + 2, b);
        }
        if (tid > 0) { const float2 l = s_last[b][tid - 1]; pI = l.x; pQ = l.y; }
        else if (!edge) { const float2 l = s_last[b ^ 1][NT - 1]; pI = l.x; pQ = l.y; }

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
        float* dst = if_out + (long long)s * if_stride + m0 + tid * R;
        const int left = n_if - (m0 + tid * R);
        if (left >= R) {
#pragma unroll
            for (int r = 0; r < R; r += 4) *reinterpret_cast<float4*>(dst + r) = make_float4(out[r], out[r + 1], out[r + 2], out[r + 3]);
        } else {
#pragma unroll
            for (int r = 0; r < R; r++) if (r < left) dst[r] = out[r];
        }
    }
}

template <int MODE, int D, int R, int NT, bool EXACT, int UNPACK, int MINB>
cudaError_t launch_v3(const Dy4FrontendArgs& a, cudaStream_t st)
{
    constexpr int T = NT * R;
    constexpr int TILE_BYTES = 2 * D * T + DY4_IQ_TAIL;
    constexpr int BUF_BYTES = (TILE_BYTES + 64 + 127) / 128 * 128;
    const size_t smem = 2 * BUF_BYTES;
    auto kern = k_frontend_v3<MODE, D, R, NT, EXACT, UNPACK, MINB>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    static int ctas_per_sm = 0, sms = 0;
    if (!ctas_per_sm) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas_per_sm, kern, NT, smem);
        if (e != cudaSuccess) return e;
        if (ctas_per_sm < 1) ctas_per_sm = 1;
    }
    const int tiles_per_stream = (a.n_if + T - 1) / T;
    const long long n_tiles = (long long)tiles_per_stream * a.n_streams;
    if (n_tiles > 0x7fffffffLL) return cudaErrorInvalidValue;
    const int grid = (int)std::min<long long>(n_tiles, (long long)sms * ctas_per_sm);
    kern<<<grid, NT, smem, st>>>(a.iq, a.row_stride, a.iq_tail, a.if_out, a.if_stride, a.n_if, a.neg_zero2,
                                 tiles_per_stream, (int)n_tiles);
    g_dy4_launches++;
    return cudaGetLastError();
}

template <int UNPACK, int MINB>
cudaError_t launch_v3_mode(const Dy4FrontendArgs& a, cudaStream_t st)
{
    switch (a.mode) {                                  // rf_decim follows the mode (project.cpp:178-238)
    case 0: return launch_v3<0, 10, 8, 128, true, UNPACK, MINB>(a, st);
    case 1: return launch_v3<1, 5, 8, 128, true, UNPACK, MINB>(a, st);
    case 2: return launch_v3<2, 10, 8, 128, true, UNPACK, MINB>(a, st);
    case 3: return launch_v3<3, 5, 8, 128, true, UNPACK, MINB>(a, st);
    }
    return cudaErrorInvalidValue;
}

// ------------------------------------------------------------------------------------------------------------
// v4 "stream": one thread walks one SEGMENT of a stream's outputs, backwards in time, with the FIR in transposed
// form.  The reference sums h[0]x[Dm] + h[1]x[Dm-1] + ... in that order, so an output's accumulator must meet its
// samples newest first: walking the input downwards, every sample feeds all ~101/D accumulators whose windows cover
// it (tap k = D(m-M)+c is a compile-time constant for a position c in the loop body), an accumulator is born at tap
// 0 and retires after tap 100.  Per sample that is ONE unpack for 10.1 (D=10) or 20.2 (D=5) multiply-adds — the
// windowed kernels above manage 4.7 — with no shared memory, no barriers and no tile edges: the body of the loop is
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
    static const int forced = std::getenv("DY4_FE_SEG") ? atoi(std::getenv("DY4_FE_SEG")) : 0;
    const long long resident = (long long)sms * ctas_per_sm * 128;
    // whole waves: with the longest useful segment (1024 outputs) the launch needs `waves` waves of resident threads;
    // fill exactly that many with as many (hence as short as necessary) segments per stream as fit
    const long long threads_min = (long long)a.n_streams * ((a.n_if + 1023) / 1024);
    const long long waves = std::max<long long>(1, (threads_min + resident - 1) / resident);
    const long long segs_fit = std::max<long long>(1, waves * resident / a.n_streams);
    int seg = (int)((a.n_if + segs_fit - 1) / segs_fit);
    seg = std::max(64, (seg + 7) & ~7);
    if (forced >= 8) seg = forced & ~7;
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

template <int D, int R, int NT, bool EXACT>
cudaError_t launch_tma(const Dy4FrontendArgs& a, cudaStream_t st)
{
    constexpr int T = NT * R;
    constexpr int TILE_BYTES = 2 * D * T + DY4_IQ_TAIL;
    constexpr int BUF_BYTES = (TILE_BYTES + 64 + 127) / 128 * 128;
    const size_t smem = 2 * BUF_BYTES;
    auto kern = k_frontend_tma<D, R, NT, EXACT>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    static int ctas_per_sm = 0, sms = 0;
    if (!ctas_per_sm) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas_per_sm, kern, NT, smem);
        if (e != cudaSuccess) return e;
        if (ctas_per_sm < 1) ctas_per_sm = 1;
    }
    const int tiles_per_stream = (a.n_if + T - 1) / T;
    const long long n_tiles = (long long)tiles_per_stream * a.n_streams;
    if (n_tiles > 0x7fffffffLL) return cudaErrorInvalidValue;
    const int grid = (int)std::min<long long>(n_tiles, (long long)sms * ctas_per_sm);
    kern<<<grid, NT, smem, st>>>(a.iq, a.row_stride, a.iq_tail, a.if_out, a.if_stride, a.n_if, a.neg_zero2, a.mode,
                                 tiles_per_stream, (int)n_tiles);
    g_dy4_launches++;
    return cudaGetLastError();
}

template <int D, int R, int NT, bool EXACT>
cudaError_t launch(const Dy4FrontendArgs& a, cudaStream_t st)
{
    constexpr int T = NT * R;
    const size_t smem = sizeof(uint32_t) * dy4_padded_words_bf16(D, R, NT);
    auto kern = k_frontend<D, R, NT, EXACT>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    dim3 grid(a.n_streams, (a.n_if + T - 1) / T);
    kern<<<grid, NT, smem, st>>>(a.iq, a.row_stride, a.iq_tail, a.if_out, a.if_stride, a.n_if, a.taps_g, a.neg_zero2, a.mode);
    g_dy4_launches++;
    return cudaGetLastError();
}

}  // namespace

cudaError_t dy4_launch_frontend(const Dy4FrontendArgs& a, cudaStream_t st)
{
    if (a.n_if <= 0 || a.n_streams <= 0) return cudaSuccess;
    // DY4_FRONTEND=staged selects the earlier converted-to-bf16 staging kernel (A/B knob; identical results)
    static const std::string which = std::getenv("DY4_FRONTEND") ? std::getenv("DY4_FRONTEND") : "";
    static const bool staged = which == "staged";
    if (which == "" || which == "streami6") return launch_stream_mode<1, 6>(a, st);      // default
    if (which == "stream6") return launch_stream_mode<0, 6>(a, st);
    if (which == "streami4") return launch_stream_mode<1, 4>(a, st);
    if (which == "streami8") return launch_stream_mode<1, 8>(a, st);
    if (which == "v3") return launch_v3_mode<0, 4>(a, st);
    if (which == "v3i") return launch_v3_mode<1, 4>(a, st);
    // which == "tma": the windowed TMA-staged kernel of the first half of round 1 (kept for A/B)
    if (a.rf_decim == 10) return staged ? launch<10, 8, 128, true>(a, st) : launch_tma<10, 8, 128, true>(a, st);
    if (a.rf_decim == 5) return staged ? launch<5, 8, 128, true>(a, st) : launch_tma<5, 8, 128, true>(a, st);
    return cudaErrorInvalidValue;
}

cudaError_t dy4_upload_taps_frontend(const TapPairs* rf4)
{
    cudaError_t e = cudaMemcpyToSymbol(c_rf2, rf4, sizeof(TapPairs) * 4);
    if (e != cudaSuccess) return e;
    static TapPairs scaled[4];
    for (int m = 0; m < 4; m++)
        for (int k = 0; k < DY4_NTAPS + 3; k++) {
            scaled[m].t[k] = make_float2(rf4[m].t[k].x * 0.0078125f, rf4[m].t[k].y * 0.0078125f);   // exact: power of two
            // the kernels carry these as immediates (dy4_rf_taps.inc): refuse to run on a build whose constants are stale
            if (k < DY4_NTAPS && std::memcmp(&scaled[m].t[k].x, &kRfTapsScaledHost[m][k], sizeof(float)) != 0) {
                dy4_set_error("compiled-in RF taps (dy4_rf_taps.inc) differ from dy4_lpf_taps(): rebuild the library");
                return cudaErrorInvalidValue;
            }
        }
    return cudaMemcpyToSymbol(c_rf2s, scaled, sizeof(TapPairs) * 4);
}
