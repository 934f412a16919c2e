// dy4_fourier.cu — the reference's Fourier diagnostics (src/fourier.cpp) on the device:
//   DFT          fourier.cpp:14-23    naive O(n^2) DFT of a real vector
//   IDFT         fourier.cpp:98-107   naive inverse DFT of a complex vector
//   estimatePSD  fourier.cpp:37-94    Hann window, DFT per segment, 10 log10 of the scaled power, segment average
// and a batched PSD over [stream][time] rows for the receiver's IF / audio buffers; and its three radix-2 FFTs
//   compute_twiddles  fourier.cpp:125-130   exp(i * float(-2 PI float(k) / NFFT))
//   FFT_recursive     fourier.cpp:132-160   decimation in time, twiddles exp(i * float(-2 PI float(k) / size)) per level
//   FFT_improved      fourier.cpp:162-187   the same recursion on a precomputed twiddle table, stride 2^(level-1)
//   FFT_optimized     fourier.cpp:189-211   the same butterflies iteratively after a bit-reversal permutation
// (one kernel: the three are the same butterfly graph and differ only in where a butterfly's twiddle VALUE comes from).
//
// Parity means the REFERENCE'S numbers, and its DFT is not an accurate one: every twiddle is exp(i a) with the angle
// a = -2 PI k m / n narrowed to float before cosf/sinf — k m reaches n^2, so late twiddles are off by 1e-4 rad — and the
// sum runs over k in float.  An FFT, however exact, would differ from it in the fourth digit.  So the kernels do what the
// reference does: the n x n twiddle table is built on the HOST with the same libm (cexpf of the float angle), and the
// device accumulates x[k] * w[k][m] over ascending k with unfused float multiply and add (packed: re and im of an output
// bin ride in one f32x2 register).  Bins are independent, rows (segments, streams) are independent: one thread per bin.
// Results: DFT / IDFT bit-identical to the reference; PSD bit-identical except where CUDA's double log10 and glibc's
// differ in the last place of a double (tests assert one float ulp).
#include "../../include/dy4_b200.h"
#include "dy4_common.cuh"
#include "dy4_internal.h"

#include <algorithm>
#include <cmath>
#include <complex>
#include <map>
#include <mutex>
#include <vector>

#define CU(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) return dy4_cuda_fail(e_, #x); } while (0)

namespace {

constexpr double kPi = 3.14159265358979323846;          // include/dy4.h:14
constexpr int kMaxN = 2048;                              // twiddle table = n^2 x 8 bytes (32 MB at 2048)

struct Tables { float2* fwd = nullptr; float2* inv = nullptr; float* hann = nullptr; };
std::mutex g_mu;
std::map<std::pair<int, int>, Tables> g_tables;          // (device, n) -> tables

// forward: w[k][m] = exp(i * float(-2 PI (k m) / n))   (fourier.cpp:18);  inverse: +2 PI (fourier.cpp:102)
int get_tables(int n, bool need_inv, bool need_hann, Tables* out)
{
    int dev = 0;
    CU(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lk(g_mu);
    Tables& t = g_tables[{dev, n}];
    auto build = [&](double sign, float2** dst) -> int {
        std::vector<float2> h((size_t)n * n);
        for (int k = 0; k < n; k++)
            for (int m = 0; m < n; m++) {
                const std::complex<float> expval(0, (float)(sign * 2 * kPi * (k * m) / (double)(size_t)n));
                const std::complex<float> w = std::exp(expval);
                h[(size_t)k * n + m] = make_float2(w.real(), w.imag());
            }
        CU(cudaMalloc(dst, h.size() * sizeof(float2)));
        CU(cudaMemcpy(*dst, h.data(), h.size() * sizeof(float2), cudaMemcpyHostToDevice));
        return DY4_OK;
    };
    int rc;
    if (!t.fwd && (rc = build(-1.0, &t.fwd))) return rc;
    if (need_inv && !t.inv && (rc = build(+1.0, &t.inv))) return rc;
    if (need_hann && !t.hann) {
        std::vector<float> h(n);
        for (int i = 0; i < n; i++) h[i] = (float)std::pow(std::sin((float)i * kPi / (float)n), 2);   // fourier.cpp:50
        CU(cudaMalloc(&t.hann, h.size() * sizeof(float)));
        CU(cudaMemcpy(t.hann, h.data(), h.size() * sizeof(float), cudaMemcpyHostToDevice));
    }
    *out = t;
    return DY4_OK;
}

// Xf[row][m] = sum_k x[row][k] * w[k][m], m < n_out; one thread per bin, the row staged in shared memory
__global__ void __launch_bounds__(128)
k_dft(const float* __restrict__ x, long long x_stride, int n, int n_out, const float2* __restrict__ w,
      float2* __restrict__ Xf, long long X_stride, u64 nz)
{
    extern __shared__ float s_x[];
    const float* row = x + (long long)blockIdx.x * x_stride;
    for (int i = threadIdx.x; i < n; i += blockDim.x) s_x[i] = row[i];
    __syncthreads();
    const int m = blockIdx.y * blockDim.x + threadIdx.x;
    if (m >= n_out) return;
    u64 acc = 0ull;
    for (int k = 0; k < n; k++) {
        const float2 wv = __ldg(w + (size_t)k * n + m);
        acc = tap2<true>(acc, pk2(wv.x, wv.y), pk2(s_x[k], s_x[k]), nz);      // (w.re x, w.im x) rounded, then added: fourier.cpp:19
    }
    float re, im;
    upk2(acc, re, im);
    Xf[(long long)blockIdx.x * X_stride + m] = make_float2(re, im);
}

// x[row][k] = (sum_m Xf[row][m] * w[k][m]) / n with std::complex<float>'s product (four products, a difference, a sum)
__global__ void __launch_bounds__(128)
k_idft(const float2* __restrict__ Xf, long long X_stride, int n, const float2* __restrict__ w,
       float2* __restrict__ x, long long x_stride)
{
    extern __shared__ float2 s_X[];
    const float2* row = Xf + (long long)blockIdx.x * X_stride;
    for (int i = threadIdx.x; i < n; i += blockDim.x) s_X[i] = row[i];
    __syncthreads();
    const int k = blockIdx.y * blockDim.x + threadIdx.x;
    if (k >= n) return;
    float re = 0.f, im = 0.f;
    for (int m = 0; m < n; m++) {
        const float2 wv = __ldg(w + (size_t)m * n + k);      // the table is symmetric in (k, m): coalesced along k
        const float2 a = s_X[m];
        const float pr = __fsub_rn(__fmul_rn(a.x, wv.x), __fmul_rn(a.y, wv.y));
        const float pi = __fadd_rn(__fmul_rn(a.x, wv.y), __fmul_rn(a.y, wv.x));
        re = __fadd_rn(re, pr); im = __fadd_rn(im, pi);      // fourier.cpp:103
    }
    x[(long long)blockIdx.x * x_stride + k] = make_float2(__fdiv_rn(re, (float)n), __fdiv_rn(im, (float)n));   // :105
}

// One CTA = one stream x 128 bins, ALL segments in order (the reference averages the per-segment dB values
// sequentially, fourier.cpp:84-89): windowed segment into shared memory, DFT of the bins, power in dB, accumulate.
__global__ void __launch_bounds__(128)
k_psd(const float* __restrict__ samples, long long stride, int n_seg, int nfft, const float2* __restrict__ w,
      const float* __restrict__ hann, double scale, float* __restrict__ psd, long long psd_stride, u64 nz)
{
    extern __shared__ float s_x[];
    const float* row = samples + (long long)blockIdx.x * stride;
    const int m = blockIdx.y * blockDim.x + threadIdx.x;
    const int half = nfft / 2;
    float sum = 0.f;
    for (int seg = 0; seg < n_seg; seg++) {
        __syncthreads();
        for (int i = threadIdx.x; i < nfft; i += blockDim.x) s_x[i] = __fmul_rn(row[(long long)seg * nfft + i], hann[i]);   // :62
        __syncthreads();
        if (m < half) {
            u64 acc = 0ull;
            for (int k = 0; k < nfft; k++) {
                const float2 wv = __ldg(w + (size_t)k * nfft + m);
                acc = tap2<true>(acc, pk2(wv.x, wv.y), pk2(s_x[k], s_x[k]), nz);
            }
            float re, im;
            upk2(acc, re, im);
            // std::abs(complex<float>) = hypotf = (float)sqrt of the exact double sum of squares (glibc)
            const float mag = __double2float_rn(sqrt(__dadd_rn(__dmul_rn((double)re, (double)re), __dmul_rn((double)im, (double)im))));
            const float v = __double2float_rn(__dmul_rn(scale, __dmul_rn((double)mag, (double)mag)));          // :74
            const float db = __double2float_rn(__dmul_rn(10.0, log10((double)v)));                             // :75
            sum = __fadd_rn(sum, db);                                                                          // :86
        }
    }
    if (m < half) psd[(long long)blockIdx.x * psd_stride + m] = __fdiv_rn(sum, (float)n_seg);                  // :88
}

const u64 kNegZero2 = 0x8000000080000000ull;

int check_n(size_t n, const char* who)
{
    if (n < 1 || n > (size_t)kMaxN) { dy4_set_error(std::string(who) + ": transform length must be 1.." + std::to_string(kMaxN)); return DY4_ERR_ARG; }
    return DY4_OK;
}

}  // namespace

extern "C" int dy4_psd_batch(const float* d_samples, size_t row_stride, int n_streams, size_t n, int nfft, int Fs,
                             float* d_psd, size_t psd_stride, void* stream)
{
    if (!d_samples || !d_psd || n_streams < 0 || nfft < 2 || (nfft & 1)) { dy4_set_error("dy4_psd_batch: bad arguments"); return DY4_ERR_ARG; }
    int rc = check_n((size_t)nfft, "dy4_psd_batch");
    if (rc) return rc;
    const int n_seg = (int)std::floor((float)n / (float)nfft);                      // fourier.cpp:54
    if (n_seg < 1) { dy4_set_error("dy4_psd_batch: fewer samples than one segment"); return DY4_ERR_ARG; }
    if (n_streams == 0) return DY4_OK;
    Tables t;
    if ((rc = get_tables(nfft, false, true, &t))) return rc;
    const double scale = (1.0 / ((float)Fs * (float)nfft / 2.0)) * 2.0;             // fourier.cpp:74
    dim3 grid(n_streams, (nfft / 2 + 127) / 128);
    k_psd<<<grid, 128, nfft * sizeof(float), (cudaStream_t)stream>>>(d_samples, (long long)row_stride, n_seg, nfft, t.fwd, t.hann, scale,
                                                                     d_psd, (long long)psd_stride, kNegZero2);
    g_dy4_launches++;
    CU(cudaGetLastError());
    return DY4_OK;
}

// ---- FFTs --------------------------------------------------------------------------------------------------------
// One CTA per row.  Stage l (sub-transform size 2^(l+1)) combines Xf[k] and Xf[k + 2^l] exactly as the reference's three
// variants do (fourier.cpp:154-155, :181-182, :203-205): t = twiddle * odd as the four products and two sums of a complex<float>
// multiplication, then even +- t — float, unfused.  tw holds the stage's twiddles consecutively: stage l at offset 2^l - 1.
__device__ __forceinline__ float2 cmul_ref(float2 a, float2 b)
{
    return make_float2(__fadd_rn(__fmul_rn(a.x, b.x), -__fmul_rn(a.y, b.y)), __fadd_rn(__fmul_rn(a.x, b.y), __fmul_rn(a.y, b.x)));
}
__global__ void __launch_bounds__(1024)
k_fft(const float2* __restrict__ x, long long x_stride, int n, int levels, const float2* __restrict__ tw, float2* __restrict__ X, long long X_stride)
{
    extern __shared__ float2 s_v[];
    const float2* xr = x + (long long)blockIdx.x * x_stride;
    for (int i = threadIdx.x; i < n; i += blockDim.x) s_v[i] = xr[__brev((unsigned)i) >> (32 - levels)];     // fourier.cpp:115-123, :194-196
    if (levels == 0 && threadIdx.x == 0) s_v[0] = xr[0];
    __syncthreads();
    for (int l = 0; l < levels; l++) {
        const int half = 1 << l;
        for (int b = threadIdx.x; b < n / 2; b += blockDim.x) {
            const int j = b & (half - 1), k = ((b >> l) << (l + 1)) + j;
            const float2 t = cmul_ref(tw[half - 1 + j], s_v[k + half]);
            const float2 e = s_v[k];
            s_v[k] = make_float2(__fadd_rn(e.x, t.x), __fadd_rn(e.y, t.y));
            s_v[k + half] = make_float2(__fadd_rn(e.x, -t.x), __fadd_rn(e.y, -t.y));
        }
        __syncthreads();
    }
    float2* Xr = X + (long long)blockIdx.x * X_stride;
    for (int i = threadIdx.x; i < n; i += blockDim.x) Xr[i] = s_v[i];
}

// per-stage twiddle values of the chosen variant, stage l at offset 2^l - 1 (n - 1 values in all)
static int fft_stage_twiddles(size_t n, int levels, int variant, const float* twiddles, size_t n_tw, std::vector<float2>* out)
{
    out->assign(n > 1 ? n - 1 : 1, make_float2(1.0f, 0.0f));
    for (int l = 0; l < levels; l++) {
        const size_t half = (size_t)1 << l, size = half * 2, stride = n / size;
        for (size_t j = 0; j < half; j++) {
            float2 w;
            if (variant == DY4_FFT_RECURSIVE) {
                const std::complex<float> expval(0.0, -2 * kPi * float(j) / size);                 // fourier.cpp:152
                const std::complex<float> t = std::exp(expval);
                w = make_float2(t.real(), t.imag());
            } else {                                                                                // :181 / :203: twiddles[j * (n / size)]
                const size_t idx = j * stride;
                if (!twiddles || idx >= n_tw) { dy4_set_error("dy4_fft: twiddle table too short for this length (needs n/2 entries for NFFT = n)"); return DY4_ERR_ARG; }
                w = make_float2(twiddles[2 * idx], twiddles[2 * idx + 1]);
            }
            (*out)[half - 1 + j] = w;
        }
    }
    return DY4_OK;
}

extern "C" int dy4_compute_twiddles(size_t n_twiddles, int nfft, float* twiddles)
{
    if (!twiddles || nfft < 1) { dy4_set_error("dy4_compute_twiddles: bad arguments"); return DY4_ERR_ARG; }
    for (size_t k = 0; k < n_twiddles; k++) {
        const std::complex<float> expval(0.0, -2 * kPi * float(k) / nfft);                          // fourier.cpp:127
        const std::complex<float> t = std::exp(expval);
        twiddles[2 * k] = t.real(); twiddles[2 * k + 1] = t.imag();
    }
    return DY4_OK;
}

extern "C" int dy4_fft_batch(const float* d_x, size_t x_stride, int n_rows, size_t n, int variant, const float* twiddles, size_t n_twiddles,
                             float* d_X, size_t X_stride, void* stream)
{
    if (!d_x || !d_X || n_rows < 0 || variant < DY4_FFT_RECURSIVE || variant > DY4_FFT_OPTIMIZED) { dy4_set_error("dy4_fft_batch: bad arguments"); return DY4_ERR_ARG; }
    int rc = check_n(n, "dy4_fft_batch");
    if (rc) return rc;
    if (n & (n - 1)) { dy4_set_error("dy4_fft_batch: the reference's FFTs are radix-2: length must be a power of two"); return DY4_ERR_ARG; }
    if (n_rows == 0) return DY4_OK;
    int levels = 0;
    while (((size_t)1 << levels) < n) levels++;
    std::vector<float2> tw;
    if ((rc = fft_stage_twiddles(n, levels, variant, twiddles, n_twiddles, &tw))) return rc;
    float2* d_tw = nullptr;
    cudaStream_t st = (cudaStream_t)stream;
    CU(cudaMalloc(&d_tw, tw.size() * sizeof(float2)));
    cudaError_t e = cudaMemcpyAsync(d_tw, tw.data(), tw.size() * sizeof(float2), cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) {
        const int threads = (int)std::min<size_t>(1024, std::max<size_t>(32, n / 2));
        k_fft<<<n_rows, threads, n * sizeof(float2), st>>>(reinterpret_cast<const float2*>(d_x), (long long)x_stride, (int)n, levels, d_tw,
                                                           reinterpret_cast<float2*>(d_X), (long long)X_stride);
        g_dy4_launches++;
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);          // tw (host) and d_tw are released here
    cudaFree(d_tw);
    if (e != cudaSuccess) return dy4_cuda_fail(e, "dy4_fft_batch");
    return DY4_OK;
}

extern "C" int dy4_fft(const float* x, size_t n, int variant, const float* twiddles, size_t n_twiddles, float* Xf)
{
    if (!x || !Xf) { dy4_set_error("dy4_fft: null pointer"); return DY4_ERR_ARG; }
    int rc = check_n(n, "dy4_fft");
    if (rc) return rc;
    float *d_x = nullptr, *d_X = nullptr;
    CU(cudaMalloc(&d_x, n * sizeof(float2)));
    CU(cudaMalloc(&d_X, n * sizeof(float2)));
    cudaError_t e = cudaMemcpy(d_x, x, n * sizeof(float2), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) rc = dy4_fft_batch(d_x, n, 1, n, variant, twiddles, n_twiddles, d_X, n, nullptr);
    if (e == cudaSuccess && rc == DY4_OK) e = cudaMemcpy(Xf, d_X, n * sizeof(float2), cudaMemcpyDeviceToHost);
    cudaFree(d_x); cudaFree(d_X);
    if (e != cudaSuccess) return dy4_cuda_fail(e, "dy4_fft");
    return rc;
}

// ---- compatibility tier: host pointers, one vector per call (include/fourier.h:20, :31, :38) -------------------------
extern "C" int dy4_dft(const float* x, size_t n, float* Xf)
{
    if (!x || !Xf) { dy4_set_error("dy4_dft: null pointer"); return DY4_ERR_ARG; }
    int rc = check_n(n, "dy4_dft");
    if (rc) return rc;
    Tables t;
    if ((rc = get_tables((int)n, false, false, &t))) return rc;
    float* d_x = nullptr; float2* d_X = nullptr;
    CU(cudaMalloc(&d_x, n * sizeof(float)));
    CU(cudaMalloc(&d_X, n * sizeof(float2)));
    CU(cudaMemcpy(d_x, x, n * sizeof(float), cudaMemcpyHostToDevice));
    k_dft<<<dim3(1, (unsigned)((n + 127) / 128)), 128, n * sizeof(float)>>>(d_x, (long long)n, (int)n, (int)n, t.fwd, d_X, (long long)n, kNegZero2);
    g_dy4_launches++;
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaMemcpy(Xf, d_X, n * sizeof(float2), cudaMemcpyDeviceToHost);
    cudaFree(d_x); cudaFree(d_X);
    if (e != cudaSuccess) return dy4_cuda_fail(e, "dy4_dft");
    return DY4_OK;
}

extern "C" int dy4_idft(const float* Xf, size_t n, float* x)
{
    if (!x || !Xf) { dy4_set_error("dy4_idft: null pointer"); return DY4_ERR_ARG; }
    int rc = check_n(n, "dy4_idft");
    if (rc) return rc;
    Tables t;
    if ((rc = get_tables((int)n, true, false, &t))) return rc;
    float2 *d_X = nullptr, *d_x = nullptr;
    CU(cudaMalloc(&d_X, n * sizeof(float2)));
    CU(cudaMalloc(&d_x, n * sizeof(float2)));
    CU(cudaMemcpy(d_X, Xf, n * sizeof(float2), cudaMemcpyHostToDevice));
    k_idft<<<dim3(1, (unsigned)((n + 127) / 128)), 128, n * sizeof(float2)>>>(d_X, (long long)n, (int)n, t.inv, d_x, (long long)n);
    g_dy4_launches++;
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaMemcpy(x, d_x, n * sizeof(float2), cudaMemcpyDeviceToHost);
    cudaFree(d_x); cudaFree(d_X);
    if (e != cudaSuccess) return dy4_cuda_fail(e, "dy4_idft");
    return DY4_OK;
}

extern "C" int dy4_estimate_psd(const float* samples, size_t n, int nfft, int Fs, float* freq, float* psd)
{
    if (!samples || !psd || nfft < 2 || (nfft & 1)) { dy4_set_error("dy4_estimate_psd: bad arguments"); return DY4_ERR_ARG; }
    if (freq) {
        const float df = (float)Fs / (float)nfft;                                    // fourier.cpp:40-45
        for (int i = 0; i < nfft / 2; i++) freq[i] = df * (float)i;
    }
    float *d_s = nullptr, *d_p = nullptr;
    CU(cudaMalloc(&d_s, n * sizeof(float)));
    CU(cudaMalloc(&d_p, (size_t)(nfft / 2) * sizeof(float)));
    CU(cudaMemcpy(d_s, samples, n * sizeof(float), cudaMemcpyHostToDevice));
    int rc = dy4_psd_batch(d_s, n, 1, n, nfft, Fs, d_p, (size_t)(nfft / 2), nullptr);
    cudaError_t e = cudaSuccess;
    if (rc == DY4_OK) e = cudaMemcpy(psd, d_p, (size_t)(nfft / 2) * sizeof(float), cudaMemcpyDeviceToHost);
    cudaFree(d_s); cudaFree(d_p);
    if (rc) return rc;
    if (e != cudaSuccess) return dy4_cuda_fail(e, "dy4_estimate_psd");
    return DY4_OK;
}
