// dy4_common.cuh — shared device helpers and layout constants for the dy4 B200 kernels.
//
// Arithmetic contract: every value that can reach the PLL input
// (front-end FIR, discriminator, pilot band-pass) is computed exactly as the
// reference's CPU code does — float product rounded, then float sum rounded,
// taps in ascending order — because the PLL's float phase accumulator makes the
// stereo output chaotic in the last ulp of its input (tools/fma_sensitivity.py).
// On sm_100a that is the packed pair  FFMA2(x, h, -0)  ->  FADD2(acc, p):
// ptxas contracts mul.rn.f32x2+add.rn.f32x2 into one FFMA2 even with explicit
// .rn, so the product is written as an fma with an opaque -0 addend instead.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

typedef unsigned long long u64;

constexpr int DY4_NTAPS = 101;      // reference src/project.cpp:142
constexpr int DY4_IQ_TAIL = 224;    // bytes of input history kept per stream (>= 2*(100+rf_decim), 16 B multiple)
constexpr int DY4_IF_TAIL = 256;    // floats of IF history kept per stream (>= 100 + 50 delay)
constexpr int DY4_MIX_TAIL = 128;   // floats of history for full-rate back-end signals (>= 100)

// Tap pairs live in __constant__ memory, one table per mode (the taps are a pure function of the mode,
// project.cpp:260-273), uploaded once per device; kernels index them with compile-time offsets so ptxas
// feeds them to FFMA2 from uniform registers (LDCU), never from the vector register file.
struct TapPairs { float2 t[DY4_NTAPS + 3]; };   // padded to 16 B multiple

__device__ __forceinline__ u64 pk2(float lo, float hi) { u64 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void upk2(u64 v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ u64 ffma2(u64 a, u64 b, u64 c) { u64 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ u64 fadd2(u64 a, u64 b) { u64 d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ u64 f2_as_u64(float2 v) { return pk2(v.x, v.y); }

// One FIR tap on a packed pair of accumulators.
//   EXACT: acc = RN(acc + RN(x*h))   (nz must hold (-0.f,-0.f) and be opaque to ptxas)
//   else : acc = RN(acc + x*h)       (fused; ~1e-7 relative from the reference)
template <bool EXACT>
__device__ __forceinline__ u64 tap2(u64 acc, u64 x, u64 h, u64 nz)
{
    if (EXACT) return fadd2(acc, ffma2(x, h, nz));
    return ffma2(x, h, acc);
}

// Decimating 101-tap FIR over packed pairs, R consecutive outputs per thread.
// `w` points at the thread's window in shared memory; logical window index
// q = D*r - k + 100 (output r, tap k), physical = q + c0 + 2*((c0+q)/(D*R)):
// two pad pairs per D*R pairs keep 128-bit loads of neighbouring threads
// (stride D*R+2 pairs = 4*odd words) on distinct banks.
// Input index descends, so every accumulator sees its taps in ascending k.
template <int D, int R, bool EXACT, int C0>
__device__ __forceinline__ void pair_decim_fir(const u64* __restrict__ w, const u64* __restrict__ hh, u64 nz, u64 (&acc)[R])
{
    constexpr int CH = D * R;
    constexpr int QMAX = D * (R - 1) + (DY4_NTAPS - 1);
#pragma unroll
    for (int r = 0; r < R; r++) acc[r] = 0ull;
#pragma unroll
    for (int q = QMAX; q >= 0; q--) {
        const u64 x = w[C0 + q + 2 * ((C0 + q) / CH)];
#pragma unroll
        for (int r = 0; r < R; r++) {
            const int k = D * r + (DY4_NTAPS - 1) - q;
            if (k >= 0 && k < DY4_NTAPS) acc[r] = tap2<EXACT>(acc[r], x, hh[k], nz);
        }
    }
}

// Same FIR over a window stored as bf16 pairs (one 32-bit word per (I,Q) sample: I in the low half, Q in the high
// half).  The front end's samples are k/128 with |k| <= 128 — exact in bf16 — so halving the shared-memory
// footprint (more resident CTAs to hide the staging loads) costs two ALU-pipe instructions per loaded sample and
// no rounding.  Four pad words per D*R keep 128-bit loads of neighbouring threads (stride D*R+4 words = 4*odd)
// on distinct banks.  C0 must be a multiple of 4.
template <int D, int R, bool EXACT, int C0>
__device__ __forceinline__ void pair_decim_fir_bf16(const uint32_t* __restrict__ w, const u64* __restrict__ hh, u64 nz, u64 (&acc)[R])
{
    constexpr int CH = D * R;
    constexpr int QMAX = D * (R - 1) + (DY4_NTAPS - 1);
    static_assert(C0 % 4 == 0 && CH % 4 == 0, "window groups must not straddle a pad");
#pragma unroll
    for (int r = 0; r < R; r++) acc[r] = 0ull;
#pragma unroll
    for (int g = QMAX / 4; g >= 0; g--) {
        const int p = C0 + 4 * g;
        const uint4 v = *reinterpret_cast<const uint4*>(w + p + 4 * (p / CH));
        const uint32_t ws[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int j = 3; j >= 0; j--) {
            const int q = 4 * g + j;
            if (q > QMAX) continue;
            const u64 x = pk2(__uint_as_float(ws[j] << 16), __uint_as_float(ws[j] & 0xffff0000u));
#pragma unroll
            for (int r = 0; r < R; r++) {
                const int k = D * r + (DY4_NTAPS - 1) - q;
                if (k >= 0 && k < DY4_NTAPS) acc[r] = tap2<EXACT>(acc[r], x, hh[k], nz);
            }
        }
    }
}

__host__ __device__ constexpr int dy4_padded_words_bf16(int D, int R, int NT) { return (D * R + 4) * NT + 256; }
__host__ __device__ constexpr int dy4_padded_pairs(int D, int R, int NT) { return (D * R + 2) * NT + 128; }
