// Microbenchmark: the serial PLL loop's step (k_pll_sel: LDS.128 + LDS + FSETP + 2 FADD2 + FSEL + FADD + predicated FADD + LOP3) with
// NCH independent chains statically interleaved in ONE warp, W warps per SM: cycles per step per chain.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench_chains tools/ubench_chains.cu && tools/ubench_chains
#include <cstdio>
#include <cuda_runtime.h>

template <int NCH>
__global__ void __launch_bounds__(32) k(const float4* __restrict__ rows, float* out, long long* cyc, int groups)
{
    __shared__ __align__(16) float4 ring[NCH][2 * 64];
    const int lane = threadIdx.x;
    for (int c = 0; c < NCH; c++)
        for (int i = lane; i < 128; i += 32) ring[c][i] = rows[(blockIdx.x * NCH + c) * 128 + i];
    __syncwarp();
    unsigned mask[32];
#pragma unroll
    for (int i = 0; i < 32; i++) { const unsigned m = lane == i ? 0xffffffffu : 0u; asm volatile("mov.b32 %0, %1;" : "=r"(mask[i]) : "r"(m)); }
    float i0[NCH], p0[NCH];
    unsigned cph[NCH];
    for (int c = 0; c < NCH; c++) { i0[c] = 0.001f * c; p0[c] = 0.01f * c; cph[c] = 0; }
    unsigned base[NCH];
    for (int c = 0; c < NCH; c++) { base[c] = (unsigned)__cvta_generic_to_shared(&ring[c][0]); asm volatile("mov.b32 %0, %0;" : "+r"(base[c])); }
    const long long t0 = clock64();
#pragma unroll 1
    for (int g = 0; g < groups; g++) {
#pragma unroll
        for (int i = 0; i < 32; i++) {
#pragma unroll
            for (int c = 0; c < NCH; c++) {
                float4 A; float T;
                asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(A.x), "=f"(A.y), "=f"(A.z), "=f"(A.w) : "r"(base[c] + 32u * i));
                asm volatile("ld.shared.f32 %0, [%1];" : "=f"(T) : "r"(base[c] + 32u * i + 16u));
                cph[c] |= __float_as_uint(p0[c]) & mask[i];
                float il, ih, tl, th;
                asm("{\n.reg .b64 ii, axy, azw, ic, tc;\nmov.b64 ii, {%4, %4};\nmov.b64 axy, {%5, %6};\nmov.b64 azw, {%7, %8};\n"
                    "add.rn.f32x2 ic, ii, axy;\nadd.rn.f32x2 tc, azw, ic;\nmov.b64 {%0, %1}, ic;\nmov.b64 {%2, %3}, tc;\n}\n"
                    : "=f"(il), "=f"(ih), "=f"(tl), "=f"(th) : "f"(i0[c]), "f"(A.x), "f"(A.y), "f"(A.z), "f"(A.w));
                const float pl = __fadd_rn(p0[c], tl), ph = __fadd_rn(p0[c], th);
                const bool up = p0[c] > T;
                i0[c] = up ? ih : il;
                p0[c] = up ? ph : pl;
            }
        }
        // the group's vote, as in the kernel
        unsigned bad = 0;
        for (int c = 0; c < NCH; c++) bad |= ~__ballot_sync(0xffffffffu, __uint_as_float(cph[c]) < 1e30f);
        if (bad) break;
        for (int c = 0; c < NCH; c++) cph[c] = 0;
    }
    const long long t1 = clock64();
    float acc = 0;
    for (int c = 0; c < NCH; c++) acc += i0[c] + p0[c];
    if (lane == 0) { out[blockIdx.x] = acc; cyc[blockIdx.x] = t1 - t0; }
}

template <int NCH>
void run(int blocks, const char* what)
{
    const int groups = 2000;
    float4* rows; float* out; long long* cyc;
    cudaMalloc(&rows, sizeof(float4) * 128 * blocks * NCH); cudaMalloc(&out, 4 * blocks); cudaMalloc(&cyc, 8 * blocks);
    float4* h = new float4[128 * blocks * NCH];
    for (int i = 0; i < 128 * blocks * NCH; i++) h[i] = (i & 1) ? make_float4(0.5f, 1e-3f, -1e-3f, 0.1f) : make_float4(1e-7f, 2e-7f, -1e-4f, 1e-4f);
    cudaMemcpy(rows, h, sizeof(float4) * 128 * blocks * NCH, cudaMemcpyHostToDevice);
    k<NCH><<<blocks, 32>>>(rows, out, cyc, groups);
    k<NCH><<<blocks, 32>>>(rows, out, cyc, groups);
    cudaDeviceSynchronize();
    long long* hc = new long long[blocks];
    cudaMemcpy(hc, cyc, 8 * blocks, cudaMemcpyDeviceToHost);
    double mx = 0, sum = 0;
    for (int i = 0; i < blocks; i++) { mx = hc[i] > mx ? hc[i] : mx; sum += hc[i]; }
    printf("%-44s %6.2f cycles per step per chain (mean), %6.2f (slowest warp)  [%s]\n", what, sum / blocks / (groups * 32.0), mx / (groups * 32.0), cudaGetErrorString(cudaGetLastError()));
    cudaFree(rows); cudaFree(out); cudaFree(cyc); delete[] h; delete[] hc;
}

int main()
{
    int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    run<1>(sms * 1, "1 chain / warp, 1 warp / SM");
    run<1>(sms * 4, "1 chain / warp, 4 warps / SM");
    run<1>(sms * 8, "1 chain / warp, 8 warps / SM");
    run<2>(sms * 4, "2 chains / warp, 4 warps / SM");
    run<2>(sms * 2, "2 chains / warp, 2 warps / SM");
    run<3>(sms * 4, "3 chains / warp, 4 warps / SM");
    run<4>(sms * 4, "4 chains / warp, 4 warps / SM");
    run<2>(sms * 8, "2 chains / warp, 8 warps / SM");
    return 0;
}
