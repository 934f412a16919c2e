// dy4_capi.cu — compatibility tier of include/dy4_b200.h: one entry point per prototype of
// the reference's include/filter.h:17-34, host pointers in and out, one stream per call.
// Each call stages its operands on the device, runs the generic CUDA kernel of
// dy4_misc.cu (or the PLL kernel) and copies the result back; carried state is updated the
// way the reference function updates it.  Pure data movement that the reference also does on
// the host (the `state.assign` tail copy, delayBlock, interleave, up/downsample) is done by
// device-side copies so that no arithmetic and no sample ever takes a CPU path.
#include "../../include/dy4_b200.h"
#include "dy4_common.cuh"
#include "dy4_kernels.h"
#include "dy4_internal.h"

#include <cstdlib>
#include <cstring>
#include <mutex>
#include <vector>

#define CU(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) return dy4_cuda_fail(e_, #x); } while (0)

namespace {

// Scratch arena on the current device (device buffer + stream).  filter.h functions may be called from any host thread — the
// reference's main() spawns two short-lived threads per block (project.cpp:299-305) — so a thread LEASES an arena from a
// process-wide pool on its first call and hands it back when it exits: a live stream cycles through two arenas for ever
// instead of leaking a buffer and a stream per thread.
struct Arena {
    char* base = nullptr; size_t cap = 0, used = 0; cudaStream_t st = nullptr;
    int reserve(size_t bytes)
    {
        if (!st) CU(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
        if (bytes > cap) {
            if (base) { CU(cudaStreamSynchronize(st)); CU(cudaFree(base)); base = nullptr; cap = 0; }
            size_t want = bytes + (bytes >> 2) + 4096;
            CU(cudaMalloc(&base, want));
            cap = want;
        }
        used = 0;
        return DY4_OK;
    }
    template <typename T> T* take(size_t n) { used = (used + 255) & ~(size_t)255; T* p = (T*)(base + used); used += n * sizeof(T); return p; }
};
struct ArenaPool { std::mutex mu; std::vector<Arena*> free_list; };
ArenaPool& arena_pool() { static ArenaPool* p = new ArenaPool(); return *p; }      // never destroyed: threads may outlive static destructors
struct ArenaLease {
    Arena* a = nullptr;
    Arena& get()
    {
        if (!a) {
            ArenaPool& p = arena_pool();
            std::lock_guard<std::mutex> lk(p.mu);
            if (!p.free_list.empty()) { a = p.free_list.back(); p.free_list.pop_back(); }
            else a = new Arena();
        }
        return *a;
    }
    ~ArenaLease()
    {
        if (!a) return;
        ArenaPool& p = arena_pool();                  // every entry point synchronises its stream before returning: nothing is in flight
        std::lock_guard<std::mutex> lk(p.mu);
        p.free_list.push_back(a);
    }
};
thread_local ArenaLease t_lease;
#define t_arena (t_lease.get())

inline size_t al(size_t bytes) { return (bytes + 255 + 256) & ~(size_t)255; }

// y = FIR over (state || x): shared by blockConvolveFIR / downsampleBlockConvolveFIR / convolveFIR
int fir_common(float* y, size_t n_out, int step, const float* x, size_t nx, const float* h, size_t nh, const float* state, size_t nstate, size_t zero_pad_after, const float** d_x_out = nullptr)
{
    Arena& a = t_arena;
    int rc = a.reserve(al((nstate + nx + zero_pad_after) * 4) + al(nh * 4) + al(n_out * 4));
    if (rc) return rc;
    float* d_xe = a.take<float>(nstate + nx + zero_pad_after);
    float* d_h = a.take<float>(nh);
    float* d_y = a.take<float>(n_out);
    if (nstate) CU(cudaMemcpyAsync(d_xe, state, nstate * 4, cudaMemcpyHostToDevice, a.st));
    if (nx) CU(cudaMemcpyAsync(d_xe + nstate, x, nx * 4, cudaMemcpyHostToDevice, a.st));
    if (zero_pad_after) CU(cudaMemsetAsync(d_xe + nstate + nx, 0, zero_pad_after * 4, a.st));
    CU(cudaMemcpyAsync(d_h, h, nh * 4, cudaMemcpyHostToDevice, a.st));
    CU(dy4_launch_generic_fir(d_xe, (int)nstate, (int)n_out, step, d_h, (int)nh, d_y, a.st));
    if (n_out) CU(cudaMemcpyAsync(y, d_y, n_out * 4, cudaMemcpyDeviceToHost, a.st));
    CU(cudaStreamSynchronize(a.st));
    if (d_x_out) *d_x_out = d_xe + nstate;
    return DY4_OK;
}

// state <- last nstate samples of x, via the device copy (the reference's state.assign, filter.cpp:82,139,169)
int carry_tail(float* state, size_t nstate, const float* d_x, size_t nx, cudaStream_t st)
{
    if (!nstate) return DY4_OK;
    if (nx < nstate) { dy4_set_error("block shorter than the carried state"); return DY4_ERR_ARG; }
    CU(cudaMemcpyAsync(state, d_x + (nx - nstate), nstate * 4, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    return DY4_OK;
}

}  // namespace

extern "C" int dy4_iq_to_float(const uint8_t* raw, size_t n, float* out)
{
    if (!raw || !out) return DY4_ERR_ARG;
    Arena& a = t_arena;
    int rc = a.reserve(al(n) + al(n * 4));
    if (rc) return rc;
    uint8_t* d_r = a.take<uint8_t>(n);
    float* d_o = a.take<float>(n);
    CU(cudaMemcpyAsync(d_r, raw, n, cudaMemcpyHostToDevice, a.st));
    CU(dy4_launch_u8_to_float(d_r, (long long)n, d_o, a.st));
    CU(cudaMemcpyAsync(out, d_o, n * 4, cudaMemcpyDeviceToHost, a.st));
    CU(cudaStreamSynchronize(a.st));
    return DY4_OK;
}

extern "C" int dy4_convolve_fir(float* y, const float* x, size_t nx, const float* h, size_t nh)
{
    if (!y || !x || !h || !nh) return DY4_ERR_ARG;
    // filter.cpp:53-64: full convolution = FIR over x followed by nh-1 zeros, no history
    return fir_common(y, nx + nh - 1, 1, x, nx, h, nh, nullptr, 0, nh - 1);
}

extern "C" int dy4_block_fir(float* y, const float* x, size_t nx, const float* h, size_t nh, float* state, size_t nstate)
{
    if (!y || !x || !h || (nstate && !state)) return DY4_ERR_ARG;
    const float* d_x = nullptr;
    int rc = fir_common(y, nx, 1, x, nx, h, nh, state, nstate, 0, &d_x);
    if (rc) return rc;
    return carry_tail(state, nstate, d_x, nx, t_arena.st);
}

extern "C" int dy4_decim_fir(int factor, float* y, const float* x, size_t nx, const float* h, size_t nh, float* state, size_t nstate)
{
    if (!y || !x || !h || factor <= 0 || (nstate && !state)) return DY4_ERR_ARG;
    const float* d_x = nullptr;
    int rc = fir_common(y, nx / (size_t)factor, factor, x, nx, h, nh, state, nstate, 0, &d_x);
    if (rc) return rc;
    return carry_tail(state, nstate, d_x, nx, t_arena.st);
}

extern "C" int dy4_resample_fir(int up, int down, float* y, size_t* ny, const float* x, size_t nx, const float* h, size_t nh, float* state, size_t nstate)
{
    if (!y || !x || !h || up <= 0 || down <= 0 || (nstate && !state)) return DY4_ERR_ARG;
    const size_t n_out = (size_t)(((float)nx / (float)down) * (float)up);    // filter.cpp:149 (float arithmetic)
    Arena& a = t_arena;
    int rc = a.reserve(al((nstate + nx) * 4) + al(nh * 4) + al(n_out * 4));
    if (rc) return rc;
    float* d_xe = a.take<float>(nstate + nx);
    float* d_h = a.take<float>(nh);
    float* d_y = a.take<float>(n_out);
    if (nstate) CU(cudaMemcpyAsync(d_xe, state, nstate * 4, cudaMemcpyHostToDevice, a.st));
    CU(cudaMemcpyAsync(d_xe + nstate, x, nx * 4, cudaMemcpyHostToDevice, a.st));
    CU(cudaMemcpyAsync(d_h, h, nh * 4, cudaMemcpyHostToDevice, a.st));
    CU(dy4_launch_generic_resample(d_xe, (int)nstate, (int)n_out, up, down, d_h, (int)nh, d_y, a.st));
    if (n_out) CU(cudaMemcpyAsync(y, d_y, n_out * 4, cudaMemcpyDeviceToHost, a.st));
    if (ny) *ny = n_out;
    return carry_tail(state, nstate, d_xe + nstate, nx, a.st);
}

extern "C" int dy4_fm_demod(const float* I, const float* Q, size_t n, float* prev_I, float* prev_Q, float* out)
{
    if (!I || !Q || !prev_I || !prev_Q || !out || !n) return DY4_ERR_ARG;
    Arena& a = t_arena;
    int rc = a.reserve(3 * al(n * 4));
    if (rc) return rc;
    float* d_i = a.take<float>(n); float* d_q = a.take<float>(n); float* d_o = a.take<float>(n);
    CU(cudaMemcpyAsync(d_i, I, n * 4, cudaMemcpyHostToDevice, a.st));
    CU(cudaMemcpyAsync(d_q, Q, n * 4, cudaMemcpyHostToDevice, a.st));
    CU(dy4_launch_generic_demod(d_i, d_q, (int)n, *prev_I, *prev_Q, d_o, a.st));
    CU(cudaMemcpyAsync(out, d_o, n * 4, cudaMemcpyDeviceToHost, a.st));
    CU(cudaMemcpyAsync(prev_I, d_i + n - 1, 4, cudaMemcpyDeviceToHost, a.st));   // filter.cpp:100-101
    CU(cudaMemcpyAsync(prev_Q, d_q + n - 1, 4, cudaMemcpyDeviceToHost, a.st));
    CU(cudaStreamSynchronize(a.st));
    return DY4_OK;
}

extern "C" int dy4_pll(const float* pll_in, size_t n, float freq, float Fs, float nco_scale, float phase_adjust, float norm_bandwidth,
                       float* nco_out, float* feedbackI, float* feedbackQ, float* integrator, float* phaseEst, float* trigOffset, float* nco_state)
{
    if (!pll_in || !nco_out || !feedbackI || !feedbackQ || !integrator || !phaseEst || !trigOffset || !nco_state || !n) return DY4_ERR_ARG;
    Arena& a = t_arena;
    int rc = a.reserve(2 * al(n * 4) + 2 * al(n * 8) + 2 * al(64));
    if (rc) return rc;
    float* d_in = a.take<float>(n); float* d_nco = a.take<float>(n); double* d_th = a.take<double>(n); double* d_inv = a.take<double>(n); float* d_st = a.take<float>(8); float* d_n0 = a.take<float>(1);
    float st[8] = {*feedbackI, *feedbackQ, *integrator, *phaseEst, *trigOffset, *nco_state, 0.f, 0.f};
    CU(cudaMemcpyAsync(d_in, pll_in, n * 4, cudaMemcpyHostToDevice, a.st));
    CU(cudaMemcpyAsync(d_st, st, sizeof(st), cudaMemcpyHostToDevice, a.st));
    Dy4PllArgs pa;
    pa.in = d_in; pa.in_stride = (long long)n; pa.nco = d_nco; pa.nco_stride = (long long)n; pa.theta = d_th; pa.inv = d_inv; pa.wide_stride = (long long)n; pa.nco0 = d_n0; pa.state = d_st;
    pa.n = (int)n; pa.n_streams = 1; pa.freq = freq; pa.Fs = Fs; pa.ncoScale = nco_scale; pa.phaseAdjust = phase_adjust; pa.normBandwidth = norm_bandwidth;
    CU(dy4_launch_pll(pa, a.st));
    CU(cudaMemcpyAsync(nco_out, d_nco, n * 4, cudaMemcpyDeviceToHost, a.st));
    CU(cudaMemcpyAsync(st, d_st, sizeof(st), cudaMemcpyDeviceToHost, a.st));
    CU(cudaStreamSynchronize(a.st));
    *feedbackI = st[0]; *feedbackQ = st[1]; *integrator = st[2]; *phaseEst = st[3]; *trigOffset = st[4]; *nco_state = st[5];
    return DY4_OK;
}

// ---- pure re-arrangements: strided device copies ------------------------------------------------------
extern "C" int dy4_downsample(const float* data, size_t n, size_t factor, float* out, size_t* n_out)
{
    if (!data || !out || !factor) return DY4_ERR_ARG;
    const size_t m = (n + factor - 1) / factor;                       // filter.cpp:107: i = 0, factor, 2*factor ... < n
    Arena& a = t_arena;
    int rc = a.reserve(al(n * 4) + al(m * 4));
    if (rc) return rc;
    float* d_x = a.take<float>(n); float* d_y = a.take<float>(m);
    CU(cudaMemcpyAsync(d_x, data, n * 4, cudaMemcpyHostToDevice, a.st));
    if (m) CU(cudaMemcpy2DAsync(d_y, 4, d_x, factor * 4, 4, m, cudaMemcpyDeviceToDevice, a.st));
    if (m) CU(cudaMemcpyAsync(out, d_y, m * 4, cudaMemcpyDeviceToHost, a.st));
    CU(cudaStreamSynchronize(a.st));
    if (n_out) *n_out = m;
    return DY4_OK;
}

extern "C" int dy4_upsample(const float* data, size_t n, size_t factor, float* out, size_t* n_out)
{
    if (!data || !out || !factor) return DY4_ERR_ARG;
    const size_t m = n * factor;                                      // filter.cpp:115-120: each sample then factor-1 zeros
    Arena& a = t_arena;
    int rc = a.reserve(al(n * 4) + al(m * 4));
    if (rc) return rc;
    float* d_x = a.take<float>(n); float* d_y = a.take<float>(m);
    CU(cudaMemcpyAsync(d_x, data, n * 4, cudaMemcpyHostToDevice, a.st));
    if (m) CU(cudaMemsetAsync(d_y, 0, m * 4, a.st));
    if (n) CU(cudaMemcpy2DAsync(d_y, factor * 4, d_x, 4, 4, n, cudaMemcpyDeviceToDevice, a.st));
    if (m) CU(cudaMemcpyAsync(out, d_y, m * 4, cudaMemcpyDeviceToHost, a.st));
    CU(cudaStreamSynchronize(a.st));
    if (n_out) *n_out = m;
    return DY4_OK;
}

extern "C" int dy4_delay_block(const float* in, size_t n, float* state, size_t nstate, float* out)
{
    if (!in || !out || (nstate && !state) || n < nstate) return DY4_ERR_ARG;
    Arena& a = t_arena;
    int rc = a.reserve(al((nstate + n) * 4));
    if (rc) return rc;
    float* d = a.take<float>(nstate + n);                             // state || in ; out = first n, new state = last nstate
    if (nstate) CU(cudaMemcpyAsync(d, state, nstate * 4, cudaMemcpyHostToDevice, a.st));
    CU(cudaMemcpyAsync(d + nstate, in, n * 4, cudaMemcpyHostToDevice, a.st));
    CU(cudaMemcpyAsync(out, d, n * 4, cudaMemcpyDeviceToHost, a.st));
    if (nstate) CU(cudaMemcpyAsync(state, d + n, nstate * 4, cudaMemcpyDeviceToHost, a.st));
    CU(cudaStreamSynchronize(a.st));
    return DY4_OK;
}

static int pointwise(int op, const float* x, const float* y, size_t n, float* out)
{
    if (!x || !y || !out) return DY4_ERR_ARG;
    Arena& a = t_arena;
    int rc = a.reserve(3 * al(n * 4));
    if (rc) return rc;
    float* d_a = a.take<float>(n); float* d_b = a.take<float>(n); float* d_o = a.take<float>(n);
    CU(cudaMemcpyAsync(d_a, x, n * 4, cudaMemcpyHostToDevice, a.st));
    CU(cudaMemcpyAsync(d_b, y, n * 4, cudaMemcpyHostToDevice, a.st));
    CU(dy4_launch_pointwise(op, d_a, d_b, (int)n, d_o, a.st));
    if (n) CU(cudaMemcpyAsync(out, d_o, n * 4, cudaMemcpyDeviceToHost, a.st));
    CU(cudaStreamSynchronize(a.st));
    return DY4_OK;
}

extern "C" int dy4_pointwise_multiply(const float* x, size_t na, const float* y, size_t nb, float* out, size_t* n_out)
{
    const size_t n = na < nb ? na : nb;                               // filter.cpp:260
    if (n_out) *n_out = n;
    return pointwise(0, x, y, n, out);
}
extern "C" int dy4_pointwise_add(const float* x, const float* y, size_t n, float* out) { return pointwise(1, x, y, n, out); }
extern "C" int dy4_pointwise_subtract(const float* x, const float* y, size_t n, float* out) { return pointwise(2, x, y, n, out); }

extern "C" int dy4_interleave(const float* left, size_t nl, const float* right, size_t nr, float* out)
{
    if (!left || !right || !out) return DY4_ERR_ARG;
    const size_t total = nl + nr;                                     // filter.cpp:292-300: even slots from left, odd from right
    const size_t n_even = (total + 1) / 2, n_odd = total / 2;
    if (n_even > nl || n_odd > nr) { dy4_set_error("dy4_interleave: channel lengths differ by more than one"); return DY4_ERR_ARG; }
    Arena& a = t_arena;
    int rc = a.reserve(al(nl * 4) + al(nr * 4) + al(total * 4));
    if (rc) return rc;
    float* d_l = a.take<float>(nl); float* d_r = a.take<float>(nr); float* d_o = a.take<float>(total);
    if (nl) CU(cudaMemcpyAsync(d_l, left, nl * 4, cudaMemcpyHostToDevice, a.st));
    if (nr) CU(cudaMemcpyAsync(d_r, right, nr * 4, cudaMemcpyHostToDevice, a.st));
    if (n_even) CU(cudaMemcpy2DAsync(d_o, 8, d_l, 4, 4, n_even, cudaMemcpyDeviceToDevice, a.st));
    if (n_odd) CU(cudaMemcpy2DAsync(d_o + 1, 8, d_r, 4, 4, n_odd, cudaMemcpyDeviceToDevice, a.st));
    if (total) CU(cudaMemcpyAsync(out, d_o, total * 4, cudaMemcpyDeviceToHost, a.st));
    CU(cudaStreamSynchronize(a.st));
    return DY4_OK;
}
