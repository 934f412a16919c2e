// ubench_pick.cu — latency of the dependent chains the table-driven PLL loop is made of (one warp, one lane active
// pattern irrelevant: fixed-latency ALU).  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench_pick ubench_pick.cu
#include <cstdio>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

template <int MODE>
__global__ void k(float* out, long long* cyc, float seed, float t_lo, float t_hi, float a0, float a1, float a2, float m, float um, int n)
{
    float x = seed, integ = seed * 0.5f;
    bool okall = true;
    long long c0 = clock64();
#pragma unroll 1
    for (int i = 0; i < n; i++) {
#pragma unroll
        for (int r = 0; r < 8; r++) {
            if (MODE == 0) x = __fadd_rn(x, a1);                                             // FADD chain
            if (MODE == 1) { const bool p = x > t_hi; x = p ? __fadd_rn(a0, 0.f) : a1; x = __fadd_rn(x, 0.f); }   // FSETP -> FSEL -> FADD
            if (MODE == 2) { const bool p = x > t_hi; x = p ? a0 : a1; }                    // FSETP -> FSEL
            if (MODE == 3) {                                                                 // the 3-candidate step without guard
                const float i0 = __fadd_rn(integ, a0), i1 = __fadd_rn(integ, a1), i2 = __fadd_rn(integ, a2);
                const float p0 = __fadd_rn(x, __fadd_rn(a0, i0)), p1 = __fadd_rn(x, __fadd_rn(a1, i1)), p2 = __fadd_rn(x, __fadd_rn(a2, i2));
                const bool neg = x < t_lo, pos = x > t_hi;
                integ = neg ? i0 : (pos ? i2 : i1);
                x = neg ? p0 : (pos ? p2 : p1);
            }
            if (MODE == 4) {                                                                 // the same with the guard
                const float i0 = __fadd_rn(integ, a0), i1 = __fadd_rn(integ, a1), i2 = __fadd_rn(integ, a2);
                const float p0 = __fadd_rn(x, __fadd_rn(a0, i0)), p1 = __fadd_rn(x, __fadd_rn(a1, i1)), p2 = __fadd_rn(x, __fadd_rn(a2, i2));
                const bool neg = x < t_lo, pos = x > t_hi;
                const float d_lo = __fadd_rn(x, -t_lo), d_hi = __fadd_rn(x, -t_hi);
                okall &= (fabsf(d_lo) > m) && (fabsf(d_hi) > m) && (d_lo > -um) && (d_hi < um);
                integ = neg ? i0 : (pos ? i2 : i1);
                x = neg ? p0 : (pos ? p2 : p1);
            }
            if (MODE == 5) {                                                                 // single candidate + guard
                integ = __fadd_rn(integ, a1);
                okall &= fabsf(__fadd_rn(x, -t_lo)) < um;
                x = __fadd_rn(x, __fadd_rn(a0, integ));
            }
            if (MODE == 6) {                                                                 // single candidate, per-step branch
                if (fabsf(__fadd_rn(x, -t_lo)) < um) { integ = __fadd_rn(integ, a1); x = __fadd_rn(x, __fadd_rn(a0, integ)); }
                else { integ = __fadd_rn(integ, a2) * 1.0001f; x = __fadd_rn(x, integ) * 0.999f; }
            }
        }
    }
    long long c1 = clock64();
    if (threadIdx.x == 0) { *cyc = c1 - c0; out[0] = x + integ + (okall ? 1.f : 0.f); }
}

int main()
{
    float* out; long long* cyc;
    CK(cudaMallocManaged(&out, 64)); CK(cudaMallocManaged(&cyc, 8));
    const int n = 1 << 16;
    const char* names[] = {"FADD chain", "FSETP->FSEL->FADD", "FSETP->FSEL", "3-candidate step", "3-candidate step + guard", "1-candidate step + guard", "1-candidate, branch per step"};
#define RUN(M) { k<M><<<1, 32>>>(out, cyc, 0.3f, -1e9f, 1e9f, 1e-9f, 2e-9f, 3e-9f, 1e-12f, 1e30f, n); CK(cudaDeviceSynchronize()); k<M><<<1, 32>>>(out, cyc, 0.3f, -1e9f, 1e9f, 1e-9f, 2e-9f, 3e-9f, 1e-12f, 1e30f, n); CK(cudaDeviceSynchronize()); printf("%-32s %7.2f cycles per step\n", names[M], (double)*cyc / n / 8); }
    RUN(0) RUN(1) RUN(2) RUN(3) RUN(4) RUN(5) RUN(6)
    return 0;
}
