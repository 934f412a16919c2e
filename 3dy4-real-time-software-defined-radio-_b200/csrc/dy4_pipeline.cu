// dy4_pipeline.cu — the batched receiver behind the throughput tier of include/dy4_b200.h.
//
// Replaces the reference's block loop, src/project.cpp:289-318: per block it
// reads stdin, spawns frontend()/backend() threads joined through threadSafeQ,
// converts to int16 and writes stdout.  Here a "chunk" of whole blocks of ALL
// streams goes through five kernels on one CUDA stream
//     front end -> twin BPF -> PLL -> audio -> tails
// with the carried state (project.cpp:25-53) resident on the device, and the
// host-facing variant overlaps host->device copies, compute and device->host
// copies of successive chunks on three streams with events (double-buffered
// staging) instead of a queue between two threads.
#include "../../include/dy4_b200.h"
#include "dy4_common.cuh"
#include "dy4_kernels.h"
#include "dy4_internal.h"

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <new>
#include <vector>

std::atomic<long long> g_dy4_launches{0};
static thread_local std::string t_err;

void dy4_set_error(const std::string& s) { t_err = s; }
int dy4_cuda_fail(cudaError_t e, const char* what)
{
    t_err = std::string(what) + ": " + cudaGetErrorString(e);
    return DY4_ERR_CUDA;
}
extern "C" const char* dy4_last_error(void) { return t_err.c_str(); }
extern "C" int dy4_version(void) { return 100; }
extern "C" long long dy4_launch_count(void) { return g_dy4_launches.load(); }

#define CU(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) return dy4_cuda_fail(e_, #x); } while (0)

struct ProfRec { int k; cudaEvent_t e0, e1; };

struct dy4_pipeline {
    int mode, stereo, n_streams, device;
    unsigned flags;
    dy4_mode_params_t mp;
    float* d_rf_taps = nullptr;
    float* d_taps_poly = nullptr;
    int up_pad = 0;
    // carried state
    uint8_t* iq_tail = nullptr; float* if_tail = nullptr; float* mix_tail = nullptr; float* pll_state = nullptr;
    // workspace for one sub-chunk
    float *ws_if = nullptr, *ws_pilot = nullptr, *ws_sband = nullptr, *ws_nco = nullptr, *ws_nco0 = nullptr;
    double *ws_theta = nullptr, *ws_inv = nullptr;
    size_t ws_stride = 0; int ws_blocks = 0; int last_n_if = 0;
    // host-facing staging
    uint8_t* d_stage[2] = {nullptr, nullptr}; int16_t* d_pcm_stage[2] = {nullptr, nullptr}; float* d_audio_stage[2] = {nullptr, nullptr};
    int stage_blocks = 0; bool stage_audio = false;
    cudaStream_t s_compute = nullptr, s_h2d = nullptr, s_d2h = nullptr;
    cudaEvent_t ev_h2d[2], ev_comp[2], ev_d2h[2];
    bool streams_ready = false;
    // profiling
    bool prof = false;
    std::vector<ProfRec> recs;
    std::vector<cudaEvent_t> pool;
    double acc_ms[DY4_NUM_KERNELS] = {0, 0, 0, 0, 0};
    long long acc_n[DY4_NUM_KERNELS] = {0, 0, 0, 0, 0};
};

namespace {

const unsigned long long kNegZero2 = 0x8000000080000000ull;

// The taps are a pure function of the mode (project.cpp:260-273): build all four tables once per
// device and park them in constant memory.
int upload_tap_tables(int device)
{
    static std::mutex mu;
    static std::vector<int> done;
    std::lock_guard<std::mutex> lk(mu);
    if (std::find(done.begin(), done.end(), device) != done.end()) return DY4_OK;
    static TapPairs rf4[4], bpf4[4], audio4[4];
    std::memset(rf4, 0, sizeof(rf4)); std::memset(bpf4, 0, sizeof(bpf4)); std::memset(audio4, 0, sizeof(audio4));
    for (int mode = 0; mode < 4; mode++) {
        dy4_mode_params_t mp;
        dy4_mode_params(mode, &mp);
        float rf[DY4_NTAPS], pilot[DY4_NTAPS], sb[DY4_NTAPS], au[DY4_NTAPS];
        dy4_lpf_taps(mp.rf_Fs, 100e3f, DY4_NTAPS, 1, rf);
        dy4_bpf_taps(mp.if_Fs, 18.5e3f, 19.5e3f, DY4_NTAPS, 1, pilot);
        dy4_bpf_taps(mp.if_Fs, 22e3f, 54e3f, DY4_NTAPS, 1, sb);
        if (mp.audio_upsample == 1) dy4_lpf_taps(mp.if_Fs, 16e3f, DY4_NTAPS, 1, au);
        else std::memset(au, 0, sizeof(au));
        for (int k = 0; k < DY4_NTAPS; k++) {
            rf4[mode].t[k] = make_float2(rf[k], rf[k]);
            bpf4[mode].t[k] = make_float2(pilot[k], sb[k]);
            audio4[mode].t[k] = make_float2(au[k], au[k]);
        }
    }
    CU(dy4_upload_taps_frontend(rf4));
    CU(dy4_upload_taps_bpf(bpf4));
    CU(dy4_upload_taps_audio(audio4));
    done.push_back(device);
    return DY4_OK;
}

int init_state(dy4_pipeline* p, cudaStream_t st)
{
    const size_t S = (size_t)p->n_streams;
    CU(cudaMemsetAsync(p->iq_tail, 128, S * DY4_IQ_TAIL, st));            // byte 128 = 0.0f: zero RF history (project.cpp:242-243)
    CU(cudaMemsetAsync(p->if_tail, 0, S * DY4_IF_TAIL * sizeof(float), st));
    CU(cudaMemsetAsync(p->mix_tail, 0, S * DY4_MIX_TAIL * sizeof(float), st));
    std::vector<float> h(S * 8, 0.0f);
    for (size_t s = 0; s < S; s++) { h[s * 8 + 0] = 1.0f; h[s * 8 + 5] = 1.0f; }   // PLLState, project.cpp:46-53
    CU(cudaMemcpyAsync(p->pll_state, h.data(), h.size() * sizeof(float), cudaMemcpyHostToDevice, st));
    CU(cudaStreamSynchronize(st));
    return DY4_OK;
}

// IF / pilot / stereo-band / NCO rows for one sub-chunk.  Sized to the job, capped by a byte budget
// (DY4_WS_BYTES, default 3 GiB); longer jobs are cut into sub-chunks, the tails carry the state across.
int ensure_workspace(dy4_pipeline* p, int n_blocks)
{
    size_t budget = 3ull << 30;
    if (const char* e = std::getenv("DY4_WS_BYTES")) budget = std::strtoull(e, nullptr, 10);
    const size_t per_block = (size_t)p->n_streams * p->mp.if_per_block * sizeof(float) * (p->stereo ? 8 : 1);
    int blocks = (int)std::max<size_t>(1, budget / per_block);
    blocks = std::min(blocks, std::max(n_blocks, 1));
    if (const char* e = std::getenv("DY4_SUBCHUNK_BLOCKS")) blocks = std::max(1, atoi(e));
    if (p->ws_blocks >= blocks) return DY4_OK;
    if (p->ws_blocks > 0) {
        CU(cudaDeviceSynchronize());
        cudaFree(p->ws_if); cudaFree(p->ws_pilot); cudaFree(p->ws_sband); cudaFree(p->ws_nco); cudaFree(p->ws_theta); cudaFree(p->ws_inv);
        p->ws_if = p->ws_pilot = p->ws_sband = p->ws_nco = nullptr; p->ws_theta = p->ws_inv = nullptr;
        p->ws_blocks = 0;
    }
    p->ws_stride = (size_t)blocks * p->mp.if_per_block;
    const size_t bytes = (size_t)p->n_streams * p->ws_stride * sizeof(float);
    CU(cudaMalloc(&p->ws_if, bytes));
    if (p->stereo) {
        CU(cudaMalloc(&p->ws_pilot, bytes));
        CU(cudaMalloc(&p->ws_sband, bytes));
        CU(cudaMalloc(&p->ws_nco, bytes));
        CU(cudaMalloc(&p->ws_theta, 2 * bytes));
        CU(cudaMalloc(&p->ws_inv, 2 * bytes));
        if (!p->ws_nco0) CU(cudaMalloc(&p->ws_nco0, (size_t)p->n_streams * sizeof(float)));
    }
    p->ws_blocks = blocks;
    return DY4_OK;
}

struct Timer {
    dy4_pipeline* p; int k; cudaStream_t st; cudaEvent_t e0 = nullptr, e1 = nullptr;
    Timer(dy4_pipeline* p_, int k_, cudaStream_t st_) : p(p_), k(k_), st(st_)
    {
        if (!p->prof) return;
        auto get = [&]() { cudaEvent_t e; if (!p->pool.empty()) { e = p->pool.back(); p->pool.pop_back(); } else cudaEventCreate(&e); return e; };
        e0 = get(); e1 = get();
        cudaEventRecord(e0, st);
    }
    ~Timer() { if (p->prof) { cudaEventRecord(e1, st); p->recs.push_back({k, e0, e1}); } }
};

int run_subchunk(dy4_pipeline* p, const uint8_t* d_iq, size_t row_stride, int nb,
                 int16_t* d_pcm, size_t pcm_stride, float* d_audio, size_t audio_stride,
                 float* d_if, size_t if_out_stride, cudaStream_t st)
{
    const dy4_mode_params_t& m = p->mp;
    const int n_if = nb * m.if_per_block, n_audio = nb * m.audio_per_block;
    const bool exact_audio = (p->flags & DY4_FLAG_EXACT_AUDIO) != 0;

    Dy4FrontendArgs fa;
    fa.iq = d_iq; fa.row_stride = (long long)row_stride; fa.iq_tail = p->iq_tail;
    fa.if_out = p->ws_if; fa.if_stride = (long long)p->ws_stride; fa.n_if = n_if; fa.n_streams = p->n_streams;
    fa.rf_decim = m.rf_decim; fa.exact = 1; fa.taps_g = p->d_rf_taps; fa.mode = p->mode; fa.neg_zero2 = kNegZero2;
    { Timer t(p, DY4_K_FRONTEND, st); CU(dy4_launch_frontend(fa, st)); }

    if (p->stereo) {
        Dy4BpfArgs ba;
        ba.if_in = p->ws_if; ba.if_stride = (long long)p->ws_stride; ba.if_tail = p->if_tail;
        ba.pilot = p->ws_pilot; ba.sband = p->ws_sband; ba.out_stride = (long long)p->ws_stride;
        ba.n_if = n_if; ba.n_streams = p->n_streams; ba.mode = p->mode; ba.neg_zero2 = kNegZero2;
        { Timer t(p, DY4_K_BPF, st); CU(dy4_launch_bpf(ba, st)); }

        Dy4PllArgs pa;
        pa.in = p->ws_pilot; pa.in_stride = (long long)p->ws_stride; pa.nco = p->ws_nco; pa.nco_stride = (long long)p->ws_stride;
        pa.theta = p->ws_theta; pa.inv = p->ws_inv; pa.wide_stride = (long long)p->ws_stride; pa.nco0 = p->ws_nco0;
        pa.state = p->pll_state; pa.n = n_if; pa.n_streams = p->n_streams;
        pa.freq = 19e3f; pa.Fs = m.if_Fs; pa.ncoScale = 2.0f; pa.phaseAdjust = 0.0f; pa.normBandwidth = 0.01f;   // project.cpp:99-102
        { Timer t(p, DY4_K_PLL, st); CU(dy4_launch_pll(pa, st)); }
    }

    Dy4AudioArgs aa;
    aa.if_in = p->ws_if; aa.if_stride = (long long)p->ws_stride; aa.if_tail = p->if_tail;
    aa.nco = p->ws_nco; aa.sband = p->ws_sband; aa.bb_stride = (long long)p->ws_stride; aa.mix_tail = p->mix_tail;
    aa.audio = d_audio; aa.audio_stride = (long long)audio_stride; aa.pcm = d_pcm; aa.pcm_stride = (long long)pcm_stride;
    aa.n_if = n_if; aa.n_audio = n_audio; aa.n_streams = p->n_streams; aa.stereo = p->stereo;
    aa.up = m.audio_upsample; aa.down = m.audio_decim; aa.exact = exact_audio ? 1 : 0;
    aa.mode = p->mode; aa.taps_poly = p->d_taps_poly; aa.up_pad = p->up_pad; aa.neg_zero2 = kNegZero2;
    if (d_audio || d_pcm) { Timer t(p, DY4_K_AUDIO, st); CU(dy4_launch_audio(aa, st)); }

    if (d_if) CU(cudaMemcpy2DAsync(d_if, if_out_stride * sizeof(float), p->ws_if, p->ws_stride * sizeof(float),
                                   (size_t)n_if * sizeof(float), p->n_streams, cudaMemcpyDeviceToDevice, st));

    Dy4TailArgs ta;
    ta.iq = d_iq; ta.row_stride = (long long)row_stride; ta.row_bytes = (long long)nb * m.block_size; ta.iq_tail = p->iq_tail;
    ta.if_in = p->ws_if; ta.if_stride = (long long)p->ws_stride; ta.n_if = n_if; ta.if_tail = p->if_tail;
    ta.nco = p->ws_nco; ta.sband = p->ws_sband; ta.bb_stride = (long long)p->ws_stride; ta.mix_tail = p->stereo ? p->mix_tail : nullptr;
    ta.n_streams = p->n_streams;
    { Timer t(p, DY4_K_TAILS, st); CU(dy4_launch_tails(ta, st)); }
    p->last_n_if = n_if;
    return DY4_OK;
}

int process_device(dy4_pipeline* p, const uint8_t* d_iq, size_t row_stride, int n_blocks,
                   int16_t* d_pcm, float* d_audio, float* d_if, cudaStream_t st,
                   size_t pcm_stride, size_t audio_stride, size_t if_stride)
{
    const dy4_mode_params_t& m = p->mp;
    const int ch = p->stereo ? 2 : 1;
    int rc = ensure_workspace(p, n_blocks);
    if (rc) return rc;
    for (int b = 0; b < n_blocks; b += p->ws_blocks) {
        const int nb = std::min(p->ws_blocks, n_blocks - b);
        rc = run_subchunk(p, d_iq + (size_t)b * m.block_size, row_stride, nb,
                          d_pcm ? d_pcm + (size_t)b * m.audio_per_block * ch : nullptr, pcm_stride,
                          d_audio ? d_audio + (size_t)b * m.audio_per_block * ch : nullptr, audio_stride,
                          d_if ? d_if + (size_t)b * m.if_per_block : nullptr, if_stride, st);
        if (rc) return rc;
    }
    return DY4_OK;
}

}  // namespace

extern "C" int dy4_pipeline_create(int mode, int stereo, int n_streams, int device, unsigned flags, dy4_pipeline_t** out)
{
    if (!out || n_streams <= 0) { dy4_set_error("dy4_pipeline_create: bad arguments"); return DY4_ERR_ARG; }
    dy4_mode_params_t mp;
    if (dy4_mode_params(mode, &mp) != DY4_OK) { dy4_set_error("dy4_pipeline_create: mode must be 0..3"); return DY4_ERR_ARG; }
    CU(cudaSetDevice(device));
    dy4_pipeline* p = new (std::nothrow) dy4_pipeline();
    if (!p) return DY4_ERR_NOMEM;
    p->mode = mode; p->stereo = stereo ? 1 : 0; p->n_streams = n_streams; p->device = device; p->flags = flags; p->mp = mp;

    // coefficient generation as project.cpp:260-273
    float rf[DY4_NTAPS], pilot[DY4_NTAPS], sb[DY4_NTAPS];
    std::vector<float> audio((size_t)mp.audio_taps);
    dy4_lpf_taps(mp.rf_Fs, 100e3f, DY4_NTAPS, 1, rf);
    dy4_lpf_taps(mp.if_Fs * (float)mp.audio_upsample, 16e3f, (unsigned short)mp.audio_taps, mp.audio_upsample, audio.data());
    dy4_bpf_taps(mp.if_Fs, 18.5e3f, 19.5e3f, DY4_NTAPS, 1, pilot);
    dy4_bpf_taps(mp.if_Fs, 22e3f, 54e3f, DY4_NTAPS, 1, sb);
    int rc0 = upload_tap_tables(device);
    if (rc0) return rc0;
    CU(cudaMalloc(&p->d_rf_taps, sizeof(rf)));
    CU(cudaMemcpy(p->d_rf_taps, rf, sizeof(rf), cudaMemcpyHostToDevice));
    if (mp.audio_upsample > 1) {
        const int U = mp.audio_upsample;
        p->up_pad = U + 1;                                  // 148: odd row stride in words spreads phases over banks/sectors
        std::vector<float> poly((size_t)DY4_NTAPS * p->up_pad, 0.0f);
        for (int j = 0; j < DY4_NTAPS; j++)
            for (int ph = 0; ph < U; ph++) poly[(size_t)j * p->up_pad + ph] = audio[(size_t)ph + (size_t)j * U];
        CU(cudaMalloc(&p->d_taps_poly, poly.size() * sizeof(float)));
        CU(cudaMemcpy(p->d_taps_poly, poly.data(), poly.size() * sizeof(float), cudaMemcpyHostToDevice));
    }
    const size_t S = (size_t)n_streams;
    CU(cudaMalloc(&p->iq_tail, S * DY4_IQ_TAIL));
    CU(cudaMalloc(&p->if_tail, S * DY4_IF_TAIL * sizeof(float)));
    CU(cudaMalloc(&p->mix_tail, S * DY4_MIX_TAIL * sizeof(float)));
    CU(cudaMalloc(&p->pll_state, S * 8 * sizeof(float)));
    int rc = init_state(p, nullptr);
    if (rc) return rc;
    *out = p;
    return DY4_OK;
}

extern "C" int dy4_pipeline_reset(dy4_pipeline_t* p)
{
    if (!p) return DY4_ERR_ARG;
    CU(cudaSetDevice(p->device));
    CU(cudaDeviceSynchronize());
    return init_state(p, nullptr);
}

extern "C" int dy4_pipeline_destroy(dy4_pipeline_t* p)
{
    if (!p) return DY4_OK;
    cudaSetDevice(p->device);
    cudaDeviceSynchronize();
    for (auto& r : p->recs) { cudaEventDestroy(r.e0); cudaEventDestroy(r.e1); }
    for (auto e : p->pool) cudaEventDestroy(e);
    cudaFree(p->d_rf_taps); cudaFree(p->d_taps_poly);
    cudaFree(p->iq_tail); cudaFree(p->if_tail); cudaFree(p->mix_tail); cudaFree(p->pll_state);
    cudaFree(p->ws_if); cudaFree(p->ws_pilot); cudaFree(p->ws_sband); cudaFree(p->ws_nco); cudaFree(p->ws_theta); cudaFree(p->ws_inv); cudaFree(p->ws_nco0);
    for (int i = 0; i < 2; i++) { cudaFree(p->d_stage[i]); cudaFree(p->d_pcm_stage[i]); cudaFree(p->d_audio_stage[i]); }
    if (p->streams_ready) {
        cudaStreamDestroy(p->s_compute); cudaStreamDestroy(p->s_h2d); cudaStreamDestroy(p->s_d2h);
        for (int i = 0; i < 2; i++) { cudaEventDestroy(p->ev_h2d[i]); cudaEventDestroy(p->ev_comp[i]); cudaEventDestroy(p->ev_d2h[i]); }
    }
    delete p;
    return DY4_OK;
}

extern "C" int dy4_pipeline_process(dy4_pipeline_t* p, const uint8_t* d_iq, size_t row_stride_bytes, int n_blocks,
                                    int16_t* d_pcm, float* d_audio, float* d_if, void* stream)
{
    if (!p || !d_iq || n_blocks < 0) { dy4_set_error("dy4_pipeline_process: bad arguments"); return DY4_ERR_ARG; }
    if (n_blocks == 0) return DY4_OK;
    const dy4_mode_params_t& m = p->mp;
    if (row_stride_bytes < (size_t)n_blocks * m.block_size || (row_stride_bytes & 15) || ((uintptr_t)d_iq & 15)) {
        dy4_set_error("dy4_pipeline_process: rows must hold n_blocks*block_size bytes and be 16-byte aligned");
        return DY4_ERR_ARG;
    }
    CU(cudaSetDevice(p->device));
    const int ch = p->stereo ? 2 : 1;
    const size_t astride = (size_t)n_blocks * m.audio_per_block * ch;
    return process_device(p, d_iq, row_stride_bytes, n_blocks, d_pcm, d_audio, d_if, (cudaStream_t)stream,
                          astride, astride, (size_t)n_blocks * m.if_per_block);
}

extern "C" int dy4_pipeline_process_host(dy4_pipeline_t* p, const uint8_t* h_iq, size_t row_stride_bytes, int n_blocks,
                                         int16_t* h_pcm, float* h_audio, int chunk_blocks)
{
    if (!p || !h_iq || n_blocks < 0) { dy4_set_error("dy4_pipeline_process_host: bad arguments"); return DY4_ERR_ARG; }
    if (n_blocks == 0) return DY4_OK;
    const dy4_mode_params_t& m = p->mp;
    if (row_stride_bytes < (size_t)n_blocks * m.block_size) { dy4_set_error("dy4_pipeline_process_host: row stride too small"); return DY4_ERR_ARG; }
    CU(cudaSetDevice(p->device));
    const int ch = p->stereo ? 2 : 1;
    const size_t S = (size_t)p->n_streams;
    if (chunk_blocks <= 0) {
        // default: ~64 MiB of input per chunk, at least one block, at most the job
        const size_t per_block = S * m.block_size;
        chunk_blocks = (int)std::max<size_t>(1, (64ull << 20) / per_block);
    }
    chunk_blocks = std::min(chunk_blocks, n_blocks);
    if (!p->streams_ready) {
        CU(cudaStreamCreateWithFlags(&p->s_compute, cudaStreamNonBlocking));
        CU(cudaStreamCreateWithFlags(&p->s_h2d, cudaStreamNonBlocking));
        CU(cudaStreamCreateWithFlags(&p->s_d2h, cudaStreamNonBlocking));
        for (int i = 0; i < 2; i++) {
            CU(cudaEventCreateWithFlags(&p->ev_h2d[i], cudaEventDisableTiming));
            CU(cudaEventCreateWithFlags(&p->ev_comp[i], cudaEventDisableTiming));
            CU(cudaEventCreateWithFlags(&p->ev_d2h[i], cudaEventDisableTiming));
        }
        p->streams_ready = true;
    }
    if (p->stage_blocks < chunk_blocks || (h_audio && !p->stage_audio)) {
        CU(cudaDeviceSynchronize());
        for (int i = 0; i < 2; i++) {
            cudaFree(p->d_stage[i]); cudaFree(p->d_pcm_stage[i]); cudaFree(p->d_audio_stage[i]);
            p->d_stage[i] = nullptr; p->d_pcm_stage[i] = nullptr; p->d_audio_stage[i] = nullptr;
            CU(cudaMalloc(&p->d_stage[i], S * chunk_blocks * m.block_size));
            CU(cudaMalloc(&p->d_pcm_stage[i], S * chunk_blocks * m.audio_per_block * ch * sizeof(int16_t)));
            if (h_audio) CU(cudaMalloc(&p->d_audio_stage[i], S * chunk_blocks * m.audio_per_block * ch * sizeof(float)));
        }
        p->stage_blocks = chunk_blocks;
        p->stage_audio = h_audio != nullptr;
    }
    const size_t total_audio = (size_t)n_blocks * m.audio_per_block * ch;   // per-stream row length of the host outputs
    int c = 0;
    for (int b = 0; b < n_blocks; b += chunk_blocks, c++) {
        const int nb = std::min(chunk_blocks, n_blocks - b), buf = c & 1;
        const size_t in_bytes = (size_t)nb * m.block_size, st_stride = (size_t)p->stage_blocks * m.block_size;
        const size_t na = (size_t)nb * m.audio_per_block * ch, out_stride = (size_t)p->stage_blocks * m.audio_per_block * ch;
        if (c >= 2) CU(cudaStreamWaitEvent(p->s_h2d, p->ev_comp[buf], 0));            // staging buffer free again
        CU(cudaMemcpy2DAsync(p->d_stage[buf], st_stride, h_iq + (size_t)b * m.block_size, row_stride_bytes,
                             in_bytes, S, cudaMemcpyHostToDevice, p->s_h2d));
        CU(cudaEventRecord(p->ev_h2d[buf], p->s_h2d));
        CU(cudaStreamWaitEvent(p->s_compute, p->ev_h2d[buf], 0));
        if (c >= 2) CU(cudaStreamWaitEvent(p->s_compute, p->ev_d2h[buf], 0));         // output staging drained
        int rc = process_device(p, p->d_stage[buf], st_stride, nb, h_pcm ? p->d_pcm_stage[buf] : nullptr,
                                h_audio ? p->d_audio_stage[buf] : nullptr, nullptr, p->s_compute, out_stride, out_stride, 0);
        if (rc) return rc;
        CU(cudaEventRecord(p->ev_comp[buf], p->s_compute));
        CU(cudaStreamWaitEvent(p->s_d2h, p->ev_comp[buf], 0));
        if (h_pcm) CU(cudaMemcpy2DAsync(h_pcm + (size_t)b * m.audio_per_block * ch, total_audio * sizeof(int16_t), p->d_pcm_stage[buf],
                                        out_stride * sizeof(int16_t), na * sizeof(int16_t), S, cudaMemcpyDeviceToHost, p->s_d2h));
        if (h_audio) CU(cudaMemcpy2DAsync(h_audio + (size_t)b * m.audio_per_block * ch, total_audio * sizeof(float), p->d_audio_stage[buf],
                                          out_stride * sizeof(float), na * sizeof(float), S, cudaMemcpyDeviceToHost, p->s_d2h));
        CU(cudaEventRecord(p->ev_d2h[buf], p->s_d2h));
    }
    CU(cudaStreamSynchronize(p->s_h2d));
    CU(cudaStreamSynchronize(p->s_compute));
    CU(cudaStreamSynchronize(p->s_d2h));
    return DY4_OK;
}

extern "C" int dy4_pipeline_debug_buffers(dy4_pipeline_t* p, const float** d_pilot, const float** d_nco, size_t* stride, int* n_if)
{
    if (!p || !p->stereo || !p->ws_pilot) { dy4_set_error("dy4_pipeline_debug_buffers: no stereo sub-chunk processed yet"); return DY4_ERR_ARG; }
    if (d_pilot) *d_pilot = p->ws_pilot;
    if (d_nco) *d_nco = p->ws_nco;
    if (stride) *stride = p->ws_stride;
    if (n_if) *n_if = p->last_n_if;
    return DY4_OK;
}

extern "C" int dy4_pipeline_profile(dy4_pipeline_t* p, int enable)
{
    if (!p) return DY4_ERR_ARG;
    p->prof = enable != 0;
    return DY4_OK;
}

extern "C" int dy4_pipeline_profile_get(dy4_pipeline_t* p, double* ms, long long* launches, int reset)
{
    if (!p) return DY4_ERR_ARG;
    CU(cudaSetDevice(p->device));
    for (auto& r : p->recs) {
        CU(cudaEventSynchronize(r.e1));
        float t = 0.f;
        CU(cudaEventElapsedTime(&t, r.e0, r.e1));
        p->acc_ms[r.k] += t; p->acc_n[r.k] += 1;
        p->pool.push_back(r.e0); p->pool.push_back(r.e1);
    }
    p->recs.clear();
    for (int k = 0; k < DY4_NUM_KERNELS; k++) {
        if (ms) ms[k] = p->acc_ms[k];
        if (launches) launches[k] = p->acc_n[k];
        if (reset) { p->acc_ms[k] = 0; p->acc_n[k] = 0; }
    }
    return DY4_OK;
}

// ---- checkpoint of carried state -------------------------------------------------------------------
extern "C" size_t dy4_pipeline_state_size(const dy4_pipeline_t* p)
{
    if (!p) return 0;
    const size_t S = (size_t)p->n_streams;
    return S * (DY4_IQ_TAIL + (DY4_IF_TAIL + DY4_MIX_TAIL + 8) * sizeof(float));
}

extern "C" int dy4_pipeline_get_state(dy4_pipeline_t* p, void* host_buf)
{
    if (!p || !host_buf) return DY4_ERR_ARG;
    CU(cudaSetDevice(p->device));
    CU(cudaDeviceSynchronize());
    const size_t S = (size_t)p->n_streams;
    char* o = (char*)host_buf;
    CU(cudaMemcpy(o, p->iq_tail, S * DY4_IQ_TAIL, cudaMemcpyDeviceToHost)); o += S * DY4_IQ_TAIL;
    CU(cudaMemcpy(o, p->if_tail, S * DY4_IF_TAIL * sizeof(float), cudaMemcpyDeviceToHost)); o += S * DY4_IF_TAIL * sizeof(float);
    CU(cudaMemcpy(o, p->mix_tail, S * DY4_MIX_TAIL * sizeof(float), cudaMemcpyDeviceToHost)); o += S * DY4_MIX_TAIL * sizeof(float);
    CU(cudaMemcpy(o, p->pll_state, S * 8 * sizeof(float), cudaMemcpyDeviceToHost));
    return DY4_OK;
}

extern "C" int dy4_pipeline_set_state(dy4_pipeline_t* p, const void* host_buf)
{
    if (!p || !host_buf) return DY4_ERR_ARG;
    CU(cudaSetDevice(p->device));
    CU(cudaDeviceSynchronize());
    const size_t S = (size_t)p->n_streams;
    const char* o = (const char*)host_buf;
    CU(cudaMemcpy(p->iq_tail, o, S * DY4_IQ_TAIL, cudaMemcpyHostToDevice)); o += S * DY4_IQ_TAIL;
    CU(cudaMemcpy(p->if_tail, o, S * DY4_IF_TAIL * sizeof(float), cudaMemcpyHostToDevice)); o += S * DY4_IF_TAIL * sizeof(float);
    CU(cudaMemcpy(p->mix_tail, o, S * DY4_MIX_TAIL * sizeof(float), cudaMemcpyHostToDevice)); o += S * DY4_MIX_TAIL * sizeof(float);
    CU(cudaMemcpy(p->pll_state, o, S * 8 * sizeof(float), cudaMemcpyHostToDevice));
    return DY4_OK;
}
