// dy4_internal.h — host-side helpers shared by the library's translation units.
#pragma once
#include <atomic>
#include <string>
#include <cuda_runtime.h>

extern std::atomic<long long> g_dy4_launches;   // kernels launched by this library in this process
void dy4_set_error(const std::string& s);
int dy4_cuda_fail(cudaError_t e, const char* what);
