// SM partition for the stereo pipeline: two CUDA green contexts on one device — a few SMs that run NOTHING but
// the PLL's serial loops (one warp per stream, bound by the issue rate of a lone warp on a dependent chain: any other warp on
// the same SM sub-partition takes issue slots from it), and the remaining SMs for every data-parallel kernel.
// The driver entry points are looked up at run time (cudaGetDriverEntryPoint): the library links no libcuda, and loads on a
// host without a driver.  Where green contexts are not available the pipeline runs unpartitioned, as before.
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <mutex>

#include "dy4_kernels.h"

namespace {

struct Part {
    bool tried = false, ok = false;
    CUgreenCtx g_loop = nullptr, g_rest = nullptr;
    int loop_sms = 0, rest_sms = 0;
};
constexpr int MAX_DEV = 64, MAX_SIZES = 16;          // loop partitions of 8, 16, ... 128 SMs
Part g_part[MAX_DEV][MAX_SIZES];
std::mutex g_mu;

template <typename F>
bool drv(const char* name, F& f)
{
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint(name, &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess || !p) { cudaGetLastError(); return false; }
    f = reinterpret_cast<F>(p);
    return true;
}

struct Api {
    CUresult (*DeviceGet)(CUdevice*, int) = nullptr;
    CUresult (*DeviceGetDevResource)(CUdevice, CUdevResource*, CUdevResourceType) = nullptr;
    CUresult (*DevSmResourceSplitByCount)(CUdevResource*, unsigned int*, const CUdevResource*, CUdevResource*, unsigned int, unsigned int) = nullptr;
    CUresult (*DevResourceGenerateDesc)(CUdevResourceDesc*, CUdevResource*, unsigned int) = nullptr;
    CUresult (*GreenCtxCreate)(CUgreenCtx*, CUdevResourceDesc, CUdevice, unsigned int) = nullptr;
    CUresult (*GreenCtxDestroy)(CUgreenCtx) = nullptr;
    CUresult (*GreenCtxStreamCreate)(CUstream*, CUgreenCtx, unsigned int, int) = nullptr;
    bool load()
    {
        return drv("cuDeviceGet", DeviceGet) && drv("cuDeviceGetDevResource", DeviceGetDevResource) &&
               drv("cuDevSmResourceSplitByCount", DevSmResourceSplitByCount) && drv("cuDevResourceGenerateDesc", DevResourceGenerateDesc) &&
               drv("cuGreenCtxCreate", GreenCtxCreate) && drv("cuGreenCtxDestroy", GreenCtxDestroy) && drv("cuGreenCtxStreamCreate", GreenCtxStreamCreate);
    }
};
Api g_api;

bool make_partition(int device, int loop_sms, Part& pt)
{
    if (!g_api.GreenCtxCreate && !g_api.load()) return false;
    if (cudaSetDevice(device) != cudaSuccess || cudaFree(nullptr) != cudaSuccess) { cudaGetLastError(); return false; }   // the primary context exists
    CUdevice dev;
    if (g_api.DeviceGet(&dev, device) != CUDA_SUCCESS) return false;
    CUdevResource all, loop, rest;
    if (g_api.DeviceGetDevResource(dev, &all, CU_DEV_RESOURCE_TYPE_SM) != CUDA_SUCCESS) return false;
    if (loop_sms <= 0 || (unsigned)loop_sms + 8 > all.sm.smCount) return false;
    unsigned int groups = 1;
    if (g_api.DevSmResourceSplitByCount(&loop, &groups, &all, &rest, 0, (unsigned)loop_sms) != CUDA_SUCCESS || groups != 1) return false;
    if (rest.sm.smCount < 8) return false;
    CUdevResourceDesc d_loop, d_rest;
    if (g_api.DevResourceGenerateDesc(&d_loop, &loop, 1) != CUDA_SUCCESS || g_api.DevResourceGenerateDesc(&d_rest, &rest, 1) != CUDA_SUCCESS) return false;
    if (g_api.GreenCtxCreate(&pt.g_loop, d_loop, dev, CU_GREEN_CTX_DEFAULT_STREAM) != CUDA_SUCCESS) return false;
    if (g_api.GreenCtxCreate(&pt.g_rest, d_rest, dev, CU_GREEN_CTX_DEFAULT_STREAM) != CUDA_SUCCESS) {
        g_api.GreenCtxDestroy(pt.g_loop); pt.g_loop = nullptr;
        return false;
    }
    pt.loop_sms = (int)loop.sm.smCount; pt.rest_sms = (int)rest.sm.smCount;
    return true;
}

}  // namespace

// One partition per device and loop size (a multiple of 8 SMs) for the life of the process, created on first use.
// Returns 1 and fills the SM counts when the partition exists, 0 when the device runs unpartitioned.
int dy4_sm_partition(int device, int loop_sms, int* n_loop, int* n_rest)
{
    if (device < 0 || device >= MAX_DEV || loop_sms < 8 || loop_sms % 8 || loop_sms / 8 > MAX_SIZES) return 0;
    std::lock_guard<std::mutex> lk(g_mu);
    Part& pt = g_part[device][loop_sms / 8 - 1];
    if (!pt.tried) {
        pt.tried = true;
        pt.ok = make_partition(device, loop_sms, pt);
        if (!pt.ok && std::getenv("DY4_PLL_STATS")) std::fprintf(stderr, "dy4: no SM partition on device %d (green contexts unavailable or refused): running unpartitioned\n", device);
    }
    if (pt.ok) { if (n_loop) *n_loop = pt.loop_sms; if (n_rest) *n_rest = pt.rest_sms; }
    return pt.ok ? 1 : 0;
}

// A non-blocking stream of the given priority that runs its kernels on the loop SMs (which = 0) or on the rest (which = 1)
cudaError_t dy4_sm_partition_stream(int device, int loop_sms, int which, int priority, cudaStream_t* s)
{
    if (device < 0 || device >= MAX_DEV || loop_sms < 8 || loop_sms % 8 || loop_sms / 8 > MAX_SIZES) return cudaErrorNotSupported;
    const Part& pt = g_part[device][loop_sms / 8 - 1];
    if (!pt.ok) return cudaErrorNotSupported;
    CUstream cs = nullptr;
    const CUresult r = g_api.GreenCtxStreamCreate(&cs, which == 0 ? pt.g_loop : pt.g_rest, CU_STREAM_NON_BLOCKING, priority);
    if (r != CUDA_SUCCESS) return cudaErrorUnknown;
    *s = reinterpret_cast<cudaStream_t>(cs);
    return cudaSuccess;
}
