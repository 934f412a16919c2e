// Instruction-throughput / latency microbenchmarks for B200 (sm_100a).
// Purpose: measure the FP32 issue ceilings the FIR kernels are judged against
// (scalar FFMA, unfused FMUL+FADD, packed f32x2 forms) and the FP64 / libm
// latencies that bound the per-stream PLL recurrence.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench tools/ubench.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

constexpr int ITER = 4096;
constexpr int NCH = 16;

__global__ void k_ffma(float* out, float a, float b) {
    float acc[NCH];
#pragma unroll
    for (int i = 0; i < NCH; i++) acc[i] = threadIdx.x + i;
    for (int it = 0; it < ITER; it++) {
#pragma unroll
        for (int i = 0; i < NCH; i++) acc[i] = fmaf(acc[i], a, b);
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < NCH; i++) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// 3 distinct register sources per FFMA (x*h + acc) as in a FIR
__global__ void k_ffma3(float* out, float a, float b) {
    float acc[NCH], x[NCH];
#pragma unroll
    for (int i = 0; i < NCH; i++) { acc[i] = threadIdx.x + i; x[i] = a * i + threadIdx.x; }
    float h0 = a, h1 = b;
    for (int it = 0; it < ITER / 2; it++) {
#pragma unroll
        for (int i = 0; i < NCH; i++) acc[i] = fmaf(x[i], h0, acc[i]);
#pragma unroll
        for (int i = 0; i < NCH; i++) acc[i] = fmaf(x[(i + 1) % NCH], h1, acc[i]);
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < NCH; i++) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void k_fmul_fadd(float* out, float a, float b) {
    float acc[NCH], x[NCH];
#pragma unroll
    for (int i = 0; i < NCH; i++) { acc[i] = threadIdx.x + i; x[i] = a * i + threadIdx.x; }
    float h0 = a, h1 = b;
    for (int it = 0; it < ITER / 2; it++) {
#pragma unroll
        for (int i = 0; i < NCH; i++) acc[i] = __fadd_rn(acc[i], __fmul_rn(x[i], h0));
#pragma unroll
        for (int i = 0; i < NCH; i++) acc[i] = __fadd_rn(acc[i], __fmul_rn(x[(i + 1) % NCH], h1));
#pragma unroll
        for (int i = 0; i < NCH; i++) x[i] = acc[(i + 5) % NCH];      // refresh the operands: otherwise the products are loop-invariant and hoisted
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < NCH; i++) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__device__ __forceinline__ unsigned long long pk(float lo, float hi) {
    unsigned long long r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ unsigned long long ffma2(unsigned long long a, unsigned long long b, unsigned long long c) {
    unsigned long long d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ unsigned long long fmul2(unsigned long long a, unsigned long long b) {
    unsigned long long d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ unsigned long long fadd2(unsigned long long a, unsigned long long b) {
    unsigned long long d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}

__global__ void k_ffma2(float* out, float a, float b) {
    unsigned long long acc[NCH], x[NCH];
#pragma unroll
    for (int i = 0; i < NCH; i++) { acc[i] = pk(threadIdx.x + i, i); x[i] = pk(a * i, threadIdx.x); }
    unsigned long long h0 = pk(a, a), h1 = pk(b, b);
    for (int it = 0; it < ITER / 2; it++) {
#pragma unroll
        for (int i = 0; i < NCH; i++) acc[i] = ffma2(x[i], h0, acc[i]);
#pragma unroll
        for (int i = 0; i < NCH; i++) acc[i] = ffma2(x[(i + 1) % NCH], h1, acc[i]);
    }
    unsigned long long s = 0;
#pragma unroll
    for (int i = 0; i < NCH; i++) s ^= acc[i];
    ((unsigned long long*)out)[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void k_fmul2_fadd2(float* out, float a, float b) {
    unsigned long long acc[NCH], x[NCH];
#pragma unroll
    for (int i = 0; i < NCH; i++) { acc[i] = pk(threadIdx.x + i, i); x[i] = pk(a * i, threadIdx.x); }
    unsigned long long h0 = pk(a, a), h1 = pk(b, b);
    for (int it = 0; it < ITER / 2; it++) {
#pragma unroll
        for (int i = 0; i < NCH; i++) acc[i] = fadd2(acc[i], fmul2(x[i], h0));
#pragma unroll
        for (int i = 0; i < NCH; i++) acc[i] = fadd2(acc[i], fmul2(x[(i + 1) % NCH], h1));
    }
    unsigned long long s = 0;
#pragma unroll
    for (int i = 0; i < NCH; i++) s ^= acc[i];
    ((unsigned long long*)out)[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// the bit-exact MAC of the FIR kernels: p = RN(x*h) as FFMA2(x, h, -0) with -0 in a vector register, acc = RN(acc + p).
// x is refreshed from another accumulator chain every iteration so the products cannot be hoisted out of the loop.
template <int MODE>   // 0: taps in registers, 1: taps as immediates, 2: FADD2 only
__global__ void k_exact2(float* out, float a, float b) {
    unsigned long long acc[NCH], x[NCH];
    __shared__ unsigned long long snz[256];
    snz[threadIdx.x] = 0x8000000080000000ull;
    unsigned long long nz = *(volatile unsigned long long*)&snz[threadIdx.x];
#pragma unroll
    for (int i = 0; i < NCH; i++) { acc[i] = pk(threadIdx.x + i, i); x[i] = pk(a * i, threadIdx.x); }
    unsigned long long h0 = pk(a, a), h1 = pk(b, b);
    for (int it = 0; it < ITER / 2; it++) {
#pragma unroll
        for (int i = 0; i < NCH; i++) {
            if (MODE == 0) acc[i] = fadd2(acc[i], ffma2(x[i], h0, nz));
            else if (MODE == 2) acc[i] = fadd2(acc[i], x[i]);
            else acc[i] = fadd2(acc[i], ffma2(x[i], pk(0.00123f + 0.001f * i, 0.00123f + 0.001f * i), nz));
        }
#pragma unroll
        for (int i = 0; i < NCH; i++) {
            if (MODE == 0) acc[i] = fadd2(acc[i], ffma2(x[(i + 1) % NCH], h1, nz));
            else if (MODE == 2) acc[i] = fadd2(acc[i], x[(i + 1) % NCH]);
            else acc[i] = fadd2(acc[i], ffma2(x[(i + 1) % NCH], pk(-0.0456f + 0.002f * i, -0.0456f + 0.002f * i), nz));
        }
#pragma unroll
        for (int i = 0; i < NCH; i++) x[i] = acc[(i + 5) % NCH];          // register renaming only after unrolling by the compiler
    }
    unsigned long long s = 0;
#pragma unroll
    for (int i = 0; i < NCH; i++) s ^= acc[i];
    ((unsigned long long*)out)[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// Where do the adds of the exact MAC go?  ncu shows FFMA2/FADD2 on the fmaheavy pipe only; scalar FADD can also use fmalite.
// SCALAR_OF_16 of the 16 chains per half-iteration do their add as two scalar FADDs (I and Q apart), the rest as one FADD2.
template <int SCALAR_OF_16>
__global__ void k_exact_mix(float* out, float a, float b) {
    unsigned long long acc[NCH], x[NCH];
    __shared__ unsigned long long snz[256];
    snz[threadIdx.x] = 0x8000000080000000ull;
    unsigned long long nz = *(volatile unsigned long long*)&snz[threadIdx.x];
#pragma unroll
    for (int i = 0; i < NCH; i++) { acc[i] = pk(threadIdx.x + i, i); x[i] = pk(a * i, threadIdx.x); }
    for (int it = 0; it < ITER / 2; it++) {
#pragma unroll
        for (int h = 0; h < 2; h++) {
#pragma unroll
            for (int i = 0; i < NCH; i++) {
                const float t = (h ? -0.0456f : 0.00123f) + 0.001f * i;
                const unsigned long long p = ffma2(x[(i + h) % NCH], pk(t, t), nz);
                if (i < SCALAR_OF_16) {
                    float al, ah, pl, ph;
                    asm("mov.b64 {%0, %1}, %2;" : "=f"(al), "=f"(ah) : "l"(acc[i]));
                    asm("mov.b64 {%0, %1}, %2;" : "=f"(pl), "=f"(ph) : "l"(p));
                    acc[i] = pk(__fadd_rn(al, pl), __fadd_rn(ah, ph));
                } else {
                    acc[i] = fadd2(acc[i], p);
                }
            }
        }
#pragma unroll
        for (int i = 0; i < NCH; i++) x[i] = acc[(i + 5) % NCH];
    }
    unsigned long long s = 0;
#pragma unroll
    for (int i = 0; i < NCH; i++) s ^= acc[i];
    ((unsigned long long*)out)[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void k_dfma(double* out, double a, double b) {
    double acc[NCH];
#pragma unroll
    for (int i = 0; i < NCH; i++) acc[i] = threadIdx.x + i;
    for (int it = 0; it < ITER / 4; it++) {
#pragma unroll
        for (int i = 0; i < NCH; i++) acc[i] = fma(acc[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NCH; i++) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// ---- latency chains: one warp, one dependent chain, clock64 ----
template <int OP>
__global__ void k_lat(double* out, long long* cyc, double seed, int n) {
    double v = seed;
    float f = (float)seed;
    long long t0 = clock64();
    for (int i = 0; i < n; i++) {
        if (OP == 0) v = fma(v, 1.0000001, 1e-9);                         // DFMA
        if (OP == 1) f = fmaf(f, 1.0000001f, 1e-9f);                       // FFMA
        if (OP == 2) v = atan2(v, 1.5) + 0.7;                              // double atan2
        if (OP == 3) { double s, c; sincos(v, &s, &c); v = s + c + 1.3; }  // sincos small arg
        if (OP == 4) { double s, c; sincos(v, &s, &c); v = s + c + 4.0e5; } // sincos large arg (> 105615)
        if (OP == 5) { f = (float)((double)f * 1.0000001 + 1e-9); }        // F2F round trip + DFMA
        if (OP == 6) v = 1.0 / (v + 2.0);                                  // double divide
        if (OP == 7) v = cos(v) + 1.3;                                     // cos small
        if (OP == 8) v = cos(v) + 4.0e5;                                   // cos large
        if (OP == 9) f = __fdividef(1.0f, f + 2.0f);                        // fast float div
    }
    long long t1 = clock64();
    out[threadIdx.x] = v + f;
    if (threadIdx.x == 0) *cyc = t1 - t0;
}

// issue interval of double-precision instructions for ONE warp (what the per-stream PLL sees: one warp per SM):
// NCHAIN independent DFMA/DADD chains, so latency is hidden and the pipe's per-warp issue rate shows.
template <int NCHAIN, int MIX>
__global__ void k_dp_issue(double* out, long long* cyc, double seed, int n) {
    double v[NCHAIN];
#pragma unroll
    for (int i = 0; i < NCHAIN; i++) v[i] = seed + i + threadIdx.x;
    long long t0 = clock64();
    for (int it = 0; it < n; it++) {
#pragma unroll
        for (int i = 0; i < NCHAIN; i++) {
            if (MIX == 0) v[i] = fma(v[i], 1.0000001, 1e-9);
            else if (MIX == 1) v[i] = __dadd_rn(v[i], 1e-9);
            else v[i] = __dmul_rn(v[i], 1.0000001);
        }
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int i = 0; i < NCHAIN; i++) s += v[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

// ping-pong between two warps of one CTA through shared memory (flag + 8-byte payload per lane): the cost of handing a
// value to a helper warp on another SMSP and getting one back — what a cross-warp split of the PLL step would pay twice.
template <int FENCE>
__global__ void k_pingpong(double* out, long long* cyc, int n) {
    __shared__ volatile int s_req, s_ack;
    __shared__ volatile double s_a[32], s_b[32];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) { s_req = 0; s_ack = 0; }
    __syncthreads();
    double v = lane;
    long long t0 = clock64();
    if (warp == 0) {
        for (int i = 1; i <= n; i++) {
            s_a[lane] = v;
            __syncwarp();
            if (lane == 0) { if (FENCE) __threadfence_block(); s_req = i; }
            while (s_ack < i) { }
            v = s_b[lane] + 1.0;
        }
    } else if (warp == 1) {
        for (int i = 1; i <= n; i++) {
            while (s_req < i) { }
            const double x = s_a[lane];
            s_b[lane] = x * 1.0000001;
            __syncwarp();
            if (lane == 0) { if (FENCE) __threadfence_block(); s_ack = i; }
        }
    }
    long long t1 = clock64();
    out[threadIdx.x] = v;
    if (threadIdx.x == 0) *cyc = t1 - t0;
}

template <typename K, typename T>
void run_tput(const char* name, K kern, T* buf, double ops_per_thread, int flops_per_op) {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    int blocks = p.multiProcessorCount * 8, threads = 256;
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    for (int w = 0; w < 3; w++) kern<<<blocks, threads>>>(buf, 1.0001f, 0.5f);
    CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < 5; r++) {
        CK(cudaEventRecord(e0));
        kern<<<blocks, threads>>>(buf, 1.0001f, 0.5f);
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
    }
    double ops = ops_per_thread * blocks * threads;
    printf("%-16s %8.3f ms  %8.2f Ginstr(thread)/s  %8.2f TFLOP/s  (%.1f lane-ops/ns/SM)\n", name, best,
           ops / best * 1e-6, ops * flops_per_op / best * 1e-9, ops / best * 1e-6 / p.multiProcessorCount);
}

int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    int clk = 0; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    printf("device %s SMs %d clock %d kHz smem/SM %zu\n", p.name, p.multiProcessorCount, clk, p.sharedMemPerMultiprocessor);
    float* buf; CK(cudaMalloc(&buf, 64 << 20));
    double n = (double)ITER * NCH;
    run_tput("ffma(imm-ish)", k_ffma, buf, n, 2);
    run_tput("ffma 3reg", k_ffma3, buf, n, 2);
    run_tput("fmul+fadd", k_fmul_fadd, buf, n, 2);       // counts MACs (2 instr each)
    run_tput("ffma2 packed", k_ffma2, buf, n, 4);          // one instr = 2 MAC
    run_tput("fmul2+fadd2", k_fmul2_fadd2, buf, n, 4);    // 2 instr = 2 MAC
    run_tput("dfma", k_dfma, (double*)buf, n / 4, 2);
    // exact MAC pairs: one op = FFMA2 + FADD2 = 2 MACs = 4 flop (the FIR kernels' arithmetic); ceiling = half the FFMA2 rate
    run_tput("exact2 reg taps", k_exact2<0>, buf, n, 4);
    run_tput("exact2 imm taps", k_exact2<1>, buf, n, 4);
    run_tput("fadd2 only", k_exact2<2>, buf, n, 2);
    run_tput("exact mix 0/16", k_exact_mix<0>, buf, n, 4);
    run_tput("exact mix 4/16", k_exact_mix<4>, buf, n, 4);
    run_tput("exact mix 8/16", k_exact_mix<8>, buf, n, 4);
    run_tput("exact mix 11/16", k_exact_mix<11>, buf, n, 4);
    run_tput("exact mix 16/16", k_exact_mix<16>, buf, n, 4);

    long long* cyc; CK(cudaMallocManaged(&cyc, 8));
    const char* names[] = {"DFMA", "FFMA", "atan2(double)", "sincos small", "sincos large", "F2F+DFMA+F2F", "ddiv", "cos small", "cos large", "fdividef"};
    int nrep = 2000;
#define LAT(OP) { k_lat<OP><<<1, 32>>>((double*)buf, cyc, 0.3, nrep); CK(cudaDeviceSynchronize()); k_lat<OP><<<1, 32>>>((double*)buf, cyc, 0.3, nrep); CK(cudaDeviceSynchronize()); printf("lat %-16s %8.1f cycles/iter\n", names[OP], (double)*cyc / nrep); }
#define DPI(NC, MIX, THREADS, NAME) { k_dp_issue<NC, MIX><<<1, THREADS>>>((double*)buf, cyc, 0.3, nrep); CK(cudaDeviceSynchronize()); k_dp_issue<NC, MIX><<<1, THREADS>>>((double*)buf, cyc, 0.3, nrep); CK(cudaDeviceSynchronize()); printf("dp issue %-28s %6.2f cycles per instruction per warp\n", NAME, (double)*cyc / nrep / NC); }
    DPI(16, 0, 32, "DFMA 1 warp x16 chains") DPI(16, 1, 32, "DADD 1 warp x16 chains") DPI(16, 2, 32, "DMUL 1 warp x16 chains")
    DPI(16, 0, 8, "DFMA 8 lanes x16 chains") DPI(16, 0, 64, "DFMA 2 warps (2 SMSPs)") DPI(16, 0, 160, "DFMA 5 warps (2 on SMSP0)") DPI(4, 0, 32, "DFMA 1 warp x4 chains")
    { k_pingpong<1><<<1, 64>>>((double*)buf, cyc, nrep); CK(cudaDeviceSynchronize()); k_pingpong<1><<<1, 64>>>((double*)buf, cyc, nrep); CK(cudaDeviceSynchronize());
      printf("warp-to-warp round trip through shared memory  %8.1f cycles (fenced)\n", (double)*cyc / nrep);
      k_pingpong<0><<<1, 64>>>((double*)buf, cyc, nrep); CK(cudaDeviceSynchronize());
      printf("warp-to-warp round trip through shared memory  %8.1f cycles (volatile only)\n", (double)*cyc / nrep); }
    LAT(0) LAT(1) LAT(2) LAT(3) LAT(4) LAT(5) LAT(6) LAT(7) LAT(8) LAT(9)
    return 0;
}
