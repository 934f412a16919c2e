// dy4_pipeline.cu — the batched receiver behind the throughput tier of include/dy4_b200.h.
//
// Replaces the reference's block loop, src/project.cpp:289-318: per block it
// reads stdin, spawns frontend()/backend() threads joined through threadSafeQ,
// converts to int16 and writes stdout.  Here a CALL of whole blocks of ALL
// streams goes through
//     front end -> band-pass pair -> PLL (predict, table, serial loop, NCO row) -> audio -> tails
// with the carried state (project.cpp:25-53) resident on the device: the FIR
// kernels over the whole call, everything else over sub-chunks of it, on four
// CUDA streams joined by events (process_device).  Consecutive
// calls can overlap (DY4_FLAG_PIPELINED), the serial loops then on SMs of their
// own (dy4_smpart.cu).  The host-facing variants overlap host->device copies,
// kernels and device->host copies (sub-chunk by sub-chunk, or call by call when
// calls overlap) instead of a queue between two threads.
#include "../../include/dy4_b200.h"
#include "dy4_common.cuh"
#include "dy4_kernels.h"
#include "dy4_internal.h"

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <functional>
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
    // carried state (input history)
    uint8_t* iq_tail = nullptr; float* if_tail = nullptr; float* mix_tail = nullptr; float* pll_state = nullptr;
    long long seq = 0;                               // sub-chunks processed so far: selects the workspace set
    // call rows: IF, pilot, stereo band and NCO of the WHOLE call, [n_streams][c_stride].  The FIR kernels fill them in full-wave
    // launches ahead of the PLL's sub-chunks; a sub-chunk's history is simply what precedes it in the same rows.
    float *c_if = nullptr, *c_pilot = nullptr, *c_sband = nullptr, *c_nco = nullptr;
    size_t c_stride = 0; int c_blocks = 0;
    // per sub-chunk PLL workspace: NSETS sets, sub-chunk c uses set c % NSETS
    static constexpr int NSETS = 3;                  // prediction + table may run two sub-chunks ahead of the serial loop
    struct WorkSet { double *theta = nullptr, *inv = nullptr; float4* tab = nullptr; };
    WorkSet ws[NSETS];
    float* ws_nco0 = nullptr;
    int* pll_risk = nullptr;                         // [n_streams]: near-tie narrowings seen by the PLL table kernel (dy4_pipeline_pll_risk)
    double* pred_state = nullptr;                    // table-driven PLL: the predictor's own state, [NSETS][n_streams][8], then [NSETS][n_streams] turns
    size_t ws_stride = 0; int ws_blocks = 0; int last_n_if = 0; size_t last_off = 0;
    // RDS filtering front end (DY4_FLAG_RDS): its own stream beside the stereo PLL
    float *rds_f = nullptr, *rds_carrier = nullptr, *rds_nco_i = nullptr, *rds_nco_q = nullptr, *rds_lp = nullptr, *rds_out = nullptr;
    double *rds_theta = nullptr, *rds_pll_state = nullptr;
    float *rds_tail = nullptr, *rds_mix_tail = nullptr, *rds_lp_tail = nullptr, *d_rds_poly = nullptr, *d_rds_rrc = nullptr;
    size_t rds_cap = 0, rds_out_cap = 0; int rds_n_out = 0, rds_call_n = 0; long long if_abs = 0;   // rds_out: the whole call's RRC rows; rds_call_n of them so far
    // RDS back half: accumulation rows of in-phase RRC samples, decoder state, growing output rows
    float* rds_acc = nullptr; size_t rds_acc_cap = 0; int rds_left = 0, rds_consumed = 0;
    int *rds_dec_state = nullptr, *rds_counts = nullptr, *rds_events = nullptr, *rds_groups = nullptr;
    int8_t *rds_sym = nullptr, *rds_bits = nullptr;
    size_t rds_sym_cap = 0, rds_bits_cap = 0, rds_ev_cap = 0, rds_grp_cap = 0;
    long long rds_blocks_since_drain = 0;
    cudaStream_t s_rds = nullptr; cudaEvent_t ev_rds = nullptr;
    bool pll_table = false;                          // table-driven PLL loop (dy4_plltab.h)
    bool pll_fresh = true;                           // no sample processed since create / reset: the next PLL launch starts the streams
    cudaStream_t s_pll = nullptr;                    // the serial PLL chain runs here, beside the FIR kernels of the next sub-chunk
    cudaStream_t s_aux = nullptr;                    // the PLL's time-parallel FP64 kernels (prediction, table) run here, beside the FP32-bound FIR kernels
    // SM partition (dy4_smpart.cu): the loops' stream runs on loop_sms SMs of its own, every other stream of a stereo call on the rest;
    // the caller's stream forks into s_main at the start of a call and joins at its end
    cudaStream_t s_main = nullptr; cudaEvent_t ev_out = nullptr; int loop_sms = 0, rest_sms = 0;
    cudaStream_t s_back = nullptr; cudaEvent_t ev_fir_all = nullptr;     // the back halves' stream: not behind the call's FIR launches
    // DY4_FLAG_PIPELINED: consecutive device-path calls overlap.  Two sets of call rows and two IF histories alternate, so that a
    // call's FIR launches can run while the call before is still being read; no join to the caller's stream at the end of a call.
    bool pipelined = false;
    float* rows_alt[4] = {nullptr, nullptr, nullptr, nullptr}; float* if_tail_alt[2] = {nullptr, nullptr};    // with if_tail: a ring of three IF histories
    const float* call_hist = nullptr;                // IF history of the call being queued (if_tail moves on as soon as its FIR work is queued)
    cudaEvent_t ev_call_done[2] = {nullptr, nullptr}, ev_main_done = nullptr;
    int call_parity = 0;
    cudaEvent_t ev_back[NSETS] = {}, ev_pll[NSETS] = {}, ev_prep[NSETS] = {}, ev_in = nullptr, ev_prep1 = nullptr;
    std::vector<cudaEvent_t> ev_fir;                 // one per FIR piece of a call
    // host-facing staging
    uint8_t* d_stage = nullptr; int16_t* d_pcm_stage = nullptr; float* d_audio_stage = nullptr;
    int stage_blocks = 0; bool stage_audio = false;
    cudaStream_t s_compute = nullptr, s_h2d = nullptr, s_d2h = nullptr;
    std::vector<cudaEvent_t> ev_up, ev_done;          // one pair per sub-chunk of a window
    // host path of a DY4_FLAG_PIPELINED pipeline: whole calls alternate between two staging sets, so that the upload of call k+1 runs
    // beside the kernels of call k (process_host_overlapped)
    struct HostSet { uint8_t* d_iq = nullptr; int16_t* d_pcm = nullptr; float* d_audio = nullptr; cudaEvent_t done = nullptr; bool used = false; };
    HostSet hset[2]; int hset_blocks = 0; bool hset_audio = false; int host_parity = 0; cudaEvent_t ev_h_up = nullptr;
    bool streams_ready = false;
    // profiling
    bool prof = false;
    std::vector<ProfRec> recs;
    std::vector<cudaEvent_t> pool;
    double acc_ms[DY4_NUM_KERNELS] = {};
    long long acc_n[DY4_NUM_KERNELS] = {};
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
    static TapPairs rf4[4], bpf4[6], audio4[4];
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
    {   // RDS band-pass filters of the Python model (fmMonoBlock.py:489-499), IF rate 240 kS/s, float32 copies of firwin's doubles
        double h[DY4_NTAPS];
        const double nyq = 240e3 / 2;
        dy4_firwin(DY4_NTAPS, 54e3 / nyq, 60e3 / nyq, 0, h);
        for (int k = 0; k < DY4_NTAPS; k++) bpf4[4].t[k] = make_float2((float)h[k], (float)h[k]);
        dy4_firwin(DY4_NTAPS, 113.5e3 / nyq, 114.5e3 / nyq, 0, h);
        for (int k = 0; k < DY4_NTAPS; k++) bpf4[5].t[k] = make_float2((float)h[k], (float)h[k]);
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
    p->seq = 0;
    p->pll_fresh = true;
    if (p->pred_state) CU(cudaMemsetAsync(p->pred_state, 0, dy4_pipeline::NSETS * S * 9 * sizeof(double), st));
    CU(cudaMemsetAsync(p->mix_tail, 0, S * DY4_MIX_TAIL * sizeof(float), st));
    std::vector<float> h(S * 8, 0.0f);
    for (size_t s = 0; s < S; s++) { h[s * 8 + 0] = 1.0f; h[s * 8 + 5] = 1.0f; }   // PLLState, project.cpp:46-53
    CU(cudaMemcpyAsync(p->pll_state, h.data(), h.size() * sizeof(float), cudaMemcpyHostToDevice, st));
    p->if_abs = 0; p->rds_n_out = 0; p->rds_call_n = 0;
    if (p->flags & DY4_FLAG_RDS) {
        CU(cudaMemsetAsync(p->rds_tail, 0, S * DY4_IF_TAIL * sizeof(float), st));
        CU(cudaMemsetAsync(p->rds_mix_tail, 0, S * 2 * DY4_MIX_TAIL * sizeof(float), st));
        CU(cudaMemsetAsync(p->rds_lp_tail, 0, S * 2 * DY4_MIX_TAIL * sizeof(float), st));
        std::vector<double> d(S * 8, 0.0);
        for (size_t s = 0; s < S; s++) { d[s * 8 + 4] = 1.0; d[s * 8 + 5] = 1.0; }   // ncoState = q_ncoState = 1.0 (fmMonoBlock.py:455,459)
        CU(cudaMemcpyAsync(p->rds_pll_state, d.data(), d.size() * sizeof(double), cudaMemcpyHostToDevice, st));
        std::vector<int> ds(S * DY4_RDS_STATE_INTS, 0);
        for (size_t s = 0; s < S; s++) {
            int* d = &ds[s * DY4_RDS_STATE_INTS];
            d[7] = 24; d[9] = -1;                                  // window_index = 24, offsetState = '' (fmMonoBlock.py:580,592)
            d[15] = d[16] = d[17] = d[18] = -1;                    // msgs.a .. msgs.d = [] (:596-600)
        }
        CU(cudaMemcpyAsync(p->rds_dec_state, ds.data(), ds.size() * sizeof(int), cudaMemcpyHostToDevice, st));
        CU(cudaMemsetAsync(p->rds_counts, 0, S * 4 * sizeof(int), st));
        p->rds_left = 0; p->rds_consumed = 0; p->rds_blocks_since_drain = 0;
    }
    CU(cudaStreamSynchronize(st));
    return DY4_OK;
}

// everything this pipeline has queued, on every stream of its own, is done
int quiesce(dy4_pipeline* p)
{
    for (cudaStream_t s : {p->s_h2d, p->s_compute, p->s_main, p->s_aux, p->s_pll, p->s_back, p->s_rds, p->s_d2h}) if (s) CU(cudaStreamSynchronize(s));
    CU(cudaDeviceSynchronize());
    return DY4_OK;
}

// The PLL's per-sub-chunk rows (predicted phase, phaseEst / reciprocal row, table), NSETS sets.  A stereo job is cut
// into sub-chunks (plan_subchunks) so that the data-parallel kernels of the sub-chunks ahead run beside the serial PLL of
// sub-chunk c; the whole thing is capped by a byte budget (DY4_WS_BYTES, default 8 GiB).  DY4_FLAG_DEBUG_ROWS keeps the job in
// one sub-chunk.
int ensure_workspace(dy4_pipeline* p, int n_blocks)
{
    size_t budget = 8ull << 30;
    if (const char* e = std::getenv("DY4_WS_BYTES")) budget = std::strtoull(e, nullptr, 10);
    // Table-driven PLL (dy4_plltab.h, 32 bytes of table per IF sample) while the stream count leaves the serial loop
    // latency-bound; with many streams the direct loop's FP64 work is already throughput-bound and the table's 3x
    // evaluations would only add to it.  DY4_PLL_TABLE_MAX=0 selects the direct loop always.
    // Measured: 4 096 streams 113 -> 131 G samples/s with the table, 8 192 streams 160 -> 135 without / with.
    // With the RDS branch beside it (its own FP64 carrier PLL and FIRs on another stream) the table's extra FP64 work
    // costs more than it saves at 4 096 streams (69.9 -> 63.8), so the switch-over is lower there.
    int tab_max = (p->flags & DY4_FLAG_RDS) ? 1024 : 4096;
    if (const char* e = std::getenv("DY4_PLL_TABLE_MAX")) tab_max = atoi(e);
    p->pll_table = p->stereo && p->n_streams <= tab_max;
    const size_t per_block = (size_t)p->n_streams * p->mp.if_per_block * (p->stereo ? (p->pll_table ? 48 : 16) * dy4_pipeline::NSETS : 4);
    int blocks = (int)std::max<size_t>(1, budget / per_block);
    blocks = std::min(blocks, std::max(n_blocks, 1));
    const bool whole = (p->flags & DY4_FLAG_DEBUG_ROWS) != 0;
    const int nsub = 3;                                // largest sub-chunk = a third of the job (see plan_subchunks)
    if (p->stereo && !whole && n_blocks >= 8) blocks = std::min(blocks, (n_blocks + nsub - 1) / nsub);
    if (const char* e = std::getenv("DY4_SUBCHUNK_BLOCKS")) blocks = std::max(1, atoi(e));
    if (whole) blocks = std::max(blocks, n_blocks);
    if (p->ws_blocks >= blocks && (p->ws_blocks == blocks || whole || n_blocks < 8 || !p->stereo)) return DY4_OK;
    if (p->ws_blocks > 0) {
        { const int rq = quiesce(p); if (rq) return rq; }
        for (auto& w : p->ws) { cudaFree(w.theta); cudaFree(w.inv); cudaFree(w.tab); w = dy4_pipeline::WorkSet(); }
        p->ws_blocks = 0;
    }
    p->ws_stride = (size_t)blocks * p->mp.if_per_block;
    const size_t bytes = (size_t)p->n_streams * p->ws_stride * sizeof(float);
    if (p->stereo)
        for (auto& w : p->ws) {
            CU(cudaMalloc(&w.theta, 2 * bytes));
            CU(cudaMalloc(&w.inv, 2 * bytes));
            if (p->pll_table) CU(cudaMalloc(&w.tab, 8 * bytes + 512 * sizeof(float4)));      // padded: the serial loop copies whole chunks of 128 rows
        }
    if (p->stereo && !p->ws_nco0) CU(cudaMalloc(&p->ws_nco0, 3 * dy4_pipeline::NSETS * (size_t)p->n_streams * sizeof(float)));   // one row per workspace set: NCO carry, sample counter, second NCO carry
    if (p->pll_table && !p->pll_risk) { CU(cudaMalloc(&p->pll_risk, (size_t)p->n_streams * sizeof(int))); CU(cudaMemset(p->pll_risk, 0, (size_t)p->n_streams * sizeof(int))); }
    if (p->pll_table && !p->pred_state) {              // [NSETS][S][8] predictor state + [NSETS][S] turns
        CU(cudaMalloc(&p->pred_state, dy4_pipeline::NSETS * (size_t)p->n_streams * 9 * sizeof(double)));
        CU(cudaMemset(p->pred_state, 0, dy4_pipeline::NSETS * (size_t)p->n_streams * 9 * sizeof(double)));
    }
    if (p->stereo && !p->s_pll) {
        // SM partition, opt-in (DY4_LOOP_SMS = SMs set aside for the serial loops, a multiple of 8; dy4_smpart.cu).  The loops
        // are bound by the issue rate of ONE warp per stream, and warps of other kernels on the same SM sub-partition take issue
        // slots from it: on SMs of their own, one warp per sub-partition, they run at 14.5 ns per sample instead of 20-25.
        // It is not the default because the other kernels then have fewer SMs and become the bound (DESIGN.md 4.3.3:
        // 256 streams 6.69 ms unpartitioned, 6.92 with 32 loop SMs, 7.47 with 64; 128 streams 5.13 -> 4.68 with 32).
        int want = 0;
        // Pipelined calls (DY4_FLAG_PIPELINED) are the case the partition is made for: the other kernels of the NEXT call run
        // while this call's loops do, so the loops' SMs are never idle and the rest never waits for a ramp.  32 SMs hold 256
        // loop warps two to a sub-partition (18.7 ns per sample), 128 one to a sub-partition (14.5).
        if (p->pipelined && p->pll_table && p->n_streams <= 256) want = std::min(32, std::max(8, (p->n_streams + 31) / 32 * 8));
        if (const char* e = std::getenv("DY4_LOOP_SMS")) want = atoi(e);
        // Stream priorities: a CTA of a higher-priority stream is dispatched before the pending CTAs of a lower one.  Without
        // them the one-block prediction of the first sub-chunk queues behind every CTA of the call's FIR launches.
        int pr_least = 0, pr_greatest = 0;
        CU(cudaDeviceGetStreamPriorityRange(&pr_least, &pr_greatest));
        const int pr_loop = pr_greatest, pr_aux = std::min(pr_least, pr_greatest + 1);
        if (want > 0 && dy4_sm_partition(p->device, want, &p->loop_sms, &p->rest_sms)) {
            CU(dy4_sm_partition_stream(p->device, want, 0, pr_loop, &p->s_pll));
            CU(dy4_sm_partition_stream(p->device, want, 1, pr_aux, &p->s_aux));
            CU(dy4_sm_partition_stream(p->device, want, 1, pr_least, &p->s_main));
            CU(dy4_sm_partition_stream(p->device, want, 1, pr_aux, &p->s_back));
            if ((p->flags & DY4_FLAG_RDS) && !p->s_rds) CU(dy4_sm_partition_stream(p->device, want, 1, pr_least, &p->s_rds));
            CU(cudaEventCreateWithFlags(&p->ev_out, cudaEventDisableTiming));
        } else {
            CU(cudaStreamCreateWithPriority(&p->s_pll, cudaStreamNonBlocking, pr_loop));
            CU(cudaStreamCreateWithPriority(&p->s_aux, cudaStreamNonBlocking, pr_aux));
            CU(cudaStreamCreateWithPriority(&p->s_back, cudaStreamNonBlocking, pr_aux));
            if (p->pipelined) {                            // the FIR launches need a stream that is never joined to the caller's
                CU(cudaStreamCreateWithPriority(&p->s_main, cudaStreamNonBlocking, pr_least));
                CU(cudaEventCreateWithFlags(&p->ev_out, cudaEventDisableTiming));
            }
        }
        CU(cudaEventCreateWithFlags(&p->ev_fir_all, cudaEventDisableTiming));
        if (p->pipelined) {
            for (auto& e : p->ev_call_done) CU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
            CU(cudaEventCreateWithFlags(&p->ev_main_done, cudaEventDisableTiming));
        }
        for (int i = 0; i < dy4_pipeline::NSETS; i++) {
            CU(cudaEventCreateWithFlags(&p->ev_back[i], cudaEventDisableTiming));
            CU(cudaEventCreateWithFlags(&p->ev_prep[i], cudaEventDisableTiming));
            CU(cudaEventCreateWithFlags(&p->ev_pll[i], cudaEventDisableTiming));
        }
        CU(cudaEventCreateWithFlags(&p->ev_in, cudaEventDisableTiming));
        CU(cudaEventCreateWithFlags(&p->ev_prep1, cudaEventDisableTiming));
    }
    if (p->flags & DY4_FLAG_RDS) {
        cudaFree(p->rds_f); cudaFree(p->rds_carrier); cudaFree(p->rds_nco_i); cudaFree(p->rds_nco_q); cudaFree(p->rds_theta);
        cudaFree(p->rds_lp);
        CU(cudaMalloc(&p->rds_f, bytes)); CU(cudaMalloc(&p->rds_carrier, bytes));
        CU(cudaMalloc(&p->rds_nco_i, bytes)); CU(cudaMalloc(&p->rds_nco_q, bytes)); CU(cudaMalloc(&p->rds_theta, 2 * bytes));
        p->rds_cap = (p->ws_stride * 19 + 119) / 120 + 4;
        CU(cudaMalloc(&p->rds_lp, (size_t)p->n_streams * 2 * p->rds_cap * sizeof(float)));
        if (!p->ev_rds) {
            if (!p->s_rds) CU(cudaStreamCreateWithFlags(&p->s_rds, cudaStreamNonBlocking));
            CU(cudaEventCreateWithFlags(&p->ev_rds, cudaEventDisableTiming));
        }
    }
    p->ws_blocks = blocks;
    return DY4_OK;
}

// blocks of one call that fit the call rows' budget (IF, pilot, stereo band, NCO: 16 bytes per IF sample; mono: 4)
int max_call_blocks(const dy4_pipeline* p)
{
    size_t budget = 32ull << 30;
    if (const char* e = std::getenv("DY4_WS_BYTES")) budget = 4 * std::strtoull(e, nullptr, 10);
    const size_t per_block = (size_t)p->n_streams * p->mp.if_per_block * (p->stereo ? (p->pipelined ? 32 : 16) : 4);
    return (int)std::max<size_t>(1, std::min<size_t>(budget / per_block, 1 << 20));
}

int ensure_call_rows(dy4_pipeline* p, int n_blocks)
{
    if (n_blocks <= p->c_blocks) return DY4_OK;
    if (p->c_blocks > 0) { int rq = quiesce(p); if (rq) return rq; }
    cudaFree(p->c_if); cudaFree(p->c_pilot); cudaFree(p->c_sband); cudaFree(p->c_nco);
    for (auto& r : p->rows_alt) { cudaFree(r); r = nullptr; }
    p->c_if = p->c_pilot = p->c_sband = p->c_nco = nullptr; p->c_blocks = 0;
    p->c_stride = (size_t)n_blocks * p->mp.if_per_block;
    const size_t bytes = (size_t)p->n_streams * p->c_stride * sizeof(float);
    CU(cudaMalloc(&p->c_if, bytes));
    if (p->stereo) { CU(cudaMalloc(&p->c_pilot, bytes)); CU(cudaMalloc(&p->c_sband, bytes)); CU(cudaMalloc(&p->c_nco, bytes)); }
    if (p->stereo && p->pipelined) for (auto& r : p->rows_alt) CU(cudaMalloc(&r, bytes));
    p->c_blocks = n_blocks;
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

struct SubChunk {                        // one sub-chunk of a call: blocks [b, b + nb)
    int b, nb;
    int16_t* pcm; float* audio;
    int set;
    int fresh = 0;                       // table-driven PLL: leading samples of a fresh stream left to the direct loop
    int pred_carry = 0;                  // table-driven PLL: the prediction continues from the previous sub-chunk's (not the first of a call)
};

// IF history of the sub-chunk that starts at block b of the call: the carried tail for b == 0, else the samples before it in the call rows
struct Hist { const float* tail; long long stride; };
Hist if_history(const dy4_pipeline* p, int b)
{
    if (b == 0) return {p->call_hist, (long long)DY4_IF_TAIL};
    return {p->c_if + (size_t)b * p->mp.if_per_block - DY4_IF_TAIL, (long long)p->c_stride};
}

// FIR work of blocks [b, b + nb) of the call on `st`: uint8 IQ -> IF -> (pilot, stereo band) into the call rows; leaves the IQ history
int run_fir(dy4_pipeline* p, const uint8_t* d_iq, size_t row_stride, int b, int nb, cudaStream_t st)
{
    const dy4_mode_params_t& m = p->mp;
    const int n_if = nb * m.if_per_block;
    const size_t off = (size_t)b * m.if_per_block;
    const uint8_t* iq = d_iq + (size_t)b * m.block_size;
    Dy4FrontendArgs fa;
    fa.iq = iq; fa.row_stride = (long long)row_stride; fa.iq_tail = p->iq_tail;
    fa.if_out = p->c_if + off; fa.if_stride = (long long)p->c_stride; fa.n_if = n_if; fa.n_streams = p->n_streams;
    fa.rf_decim = m.rf_decim; fa.exact = (p->stereo || (p->flags & DY4_FLAG_EXACT_AUDIO)) ? 1 : 0; /* the PLL needs a bit-exact IF; mono does not */ fa.taps_g = p->d_rf_taps; fa.mode = p->mode; fa.neg_zero2 = kNegZero2;
    { Timer t(p, DY4_K_FRONTEND, st); CU(dy4_launch_frontend(fa, st)); }
    Dy4TailArgs ta{};
    ta.iq = iq; ta.row_stride = (long long)row_stride; ta.row_bytes = (long long)nb * m.block_size; ta.iq_tail = p->iq_tail;
    ta.n_streams = p->n_streams;
    { Timer t(p, DY4_K_TAILS, st); CU(dy4_launch_tails(ta, st)); }
    if (p->stereo) {
        const Hist h = if_history(p, b);
        Dy4BpfArgs ba;
        ba.if_in = p->c_if + off; ba.if_stride = (long long)p->c_stride; ba.if_tail = h.tail; ba.if_tail_stride = h.stride;
        ba.pilot = p->c_pilot + off; ba.sband = p->c_sband + off; ba.out_stride = (long long)p->c_stride;
        ba.n_if = n_if; ba.n_streams = p->n_streams; ba.mode = p->mode; ba.variant = 0; ba.neg_zero2 = kNegZero2;
        // the stereo band only has to be bit-exact when the audio is asked to be (DY4_FLAG_EXACT_AUDIO)
        const bool mixed = !(p->flags & DY4_FLAG_EXACT_AUDIO);
        { Timer t(p, DY4_K_BPF, st); CU(mixed ? dy4_launch_bpf_mixed(ba, st) : dy4_launch_bpf(ba, st)); }
    }
    return DY4_OK;
}

// grow a [n_streams][cap] device array to at least `need` columns, keeping its contents (stream-ordered on st)
template <typename T>
int grow_rows(T*& buf, size_t& cap, size_t need, size_t n_streams, size_t elems, cudaStream_t st)
{
    if (need <= cap) return DY4_OK;
    const size_t ncap = std::max(need, cap * 2);
    T* nb = nullptr;
    CU(cudaMalloc(&nb, n_streams * ncap * elems * sizeof(T)));
    if (buf && cap) CU(cudaMemcpy2DAsync(nb, ncap * elems * sizeof(T), buf, cap * elems * sizeof(T), cap * elems * sizeof(T), n_streams, cudaMemcpyDeviceToDevice, st));
    CU(cudaStreamSynchronize(st));
    cudaFree(buf);
    buf = nb; cap = ncap;
    return DY4_OK;
}

// RDS back half on the samples of this call: append to the accumulation rows, decode every whole model block
int run_rds_decode(dy4_pipeline* p, cudaStream_t st)
{
    const size_t S = (size_t)p->n_streams;
    const int n_new = p->rds_n_out;
    int rc;
    if ((rc = grow_rows(p->rds_acc, p->rds_acc_cap, (size_t)p->rds_left + p->rds_consumed + n_new + 64, S, 1, st))) return rc;
    CU(dy4_launch_rds_append(p->rds_out + (p->rds_call_n - n_new), 2LL * (long long)p->rds_out_cap, n_new, p->rds_acc, (long long)p->rds_acc_cap, p->rds_consumed, p->rds_left,
                             p->n_streams, st));
    const int total = p->rds_left + n_new;
    const int nblk = total / DY4_RDS_BLOCK;
    p->rds_consumed = nblk * DY4_RDS_BLOCK;
    p->rds_left = total - p->rds_consumed;
    if (nblk == 0) return DY4_OK;
    p->rds_blocks_since_drain += nblk;
    const size_t need_sym = (size_t)p->rds_blocks_since_drain * (DY4_RDS_BLOCK / 16), need_bits = (need_sym + 1) / 2 + 1;
    if ((rc = grow_rows(p->rds_sym, p->rds_sym_cap, need_sym, S, 1, st))) return rc;
    if ((rc = grow_rows(p->rds_bits, p->rds_bits_cap, need_bits, S, 1, st))) return rc;
    if ((rc = grow_rows(p->rds_events, p->rds_ev_cap, need_bits, S, 4, st))) return rc;    // at most one event per window, one window per bit
    if ((rc = grow_rows(p->rds_groups, p->rds_grp_cap, need_bits, S, 4, st))) return rc;   // likewise one group per window at most
    Dy4RdsDecodeArgs da{};
    da.acc = p->rds_acc; da.acc_stride = (long long)p->rds_acc_cap; da.n_blocks = nblk;
    da.state = p->rds_dec_state; da.counts = p->rds_counts;
    da.sym = p->rds_sym; da.sym_stride = (long long)p->rds_sym_cap; da.sym_cap = (int)p->rds_sym_cap;
    da.bits = p->rds_bits; da.bits_stride = (long long)p->rds_bits_cap; da.bits_cap = (int)p->rds_bits_cap;
    da.events = p->rds_events; da.ev_stride = (long long)p->rds_ev_cap; da.ev_cap = (int)p->rds_ev_cap;
    da.groups = p->rds_groups; da.grp_stride = (long long)p->rds_grp_cap; da.grp_cap = (int)p->rds_grp_cap;
    da.n_streams = p->n_streams;
    CU(dy4_launch_rds_decode(da, st));
    return DY4_OK;
}

// RDS filtering front end of one sub-chunk on its own stream (fmMonoBlock.py:673-691), beside the stereo PLL
int run_rds(dy4_pipeline* p, const SubChunk& c, cudaStream_t st)
{
    const dy4_mode_params_t& m = p->mp;
    const int n_if = c.nb * m.if_per_block;
    const size_t off = (size_t)c.b * m.if_per_block;
    const Hist h = if_history(p, c.b);
    Dy4BpfArgs ba;
    ba.if_in = p->c_if + off; ba.if_stride = (long long)p->c_stride; ba.if_tail = h.tail; ba.if_tail_stride = h.stride;
    ba.pilot = p->rds_f; ba.sband = nullptr; ba.out_stride = (long long)p->ws_stride;
    ba.n_if = n_if; ba.n_streams = p->n_streams; ba.mode = 4; ba.variant = 1; ba.neg_zero2 = kNegZero2;
    {
        Timer t(p, DY4_K_RDS_BPF, st);
        CU(dy4_launch_bpf(ba, st));                               // RDS channel extraction, 54-60 kHz
        ba.if_in = p->rds_f; ba.if_stride = (long long)p->ws_stride; ba.if_tail = p->rds_tail; ba.if_tail_stride = 0;
        ba.pilot = p->rds_carrier; ba.mode = 5; ba.variant = 2;
        CU(dy4_launch_bpf(ba, st));                               // squaring + 113.5-114.5 kHz carrier extraction
    }
    Dy4RdsArgs ra{};
    ra.rds_f = p->rds_f; ra.stride = (long long)p->ws_stride; ra.rds_tail = p->rds_tail; ra.carrier = p->rds_carrier;
    ra.theta = p->rds_theta; ra.wide_stride = (long long)p->ws_stride; ra.nco_i = p->rds_nco_i; ra.nco_q = p->rds_nco_q;
    ra.pll_state = p->rds_pll_state; ra.mix_tail = p->rds_mix_tail; ra.lp = p->rds_lp; ra.lp_stride = (long long)p->rds_cap;
    ra.lp_tail = p->rds_lp_tail; ra.out_i = p->rds_out + p->rds_call_n; ra.out_q = p->rds_out + p->rds_out_cap + p->rds_call_n; ra.out_stride = 2LL * (long long)p->rds_out_cap;
    ra.taps_poly = p->d_rds_poly; ra.up_pad = 20; ra.taps_rrc = p->d_rds_rrc; ra.n_if = n_if; ra.n_streams = p->n_streams;
    ra.if_abs = p->if_abs; ra.up = 19; ra.down = 120;
    ra.m_first = (19 * p->if_abs + 119) / 120;
    ra.n_out = (int)((19 * (p->if_abs + n_if) + 119) / 120 - ra.m_first);
    const double bw = 0.001;                                      // fmMonoBlock.py:444-447
    ra.w = 2 * 3.141592653589793 * (114e3 / 240e3); ra.Kp = bw * 2.666; ra.Ki = (bw * bw) * 3.555; ra.nco_scale = 0.5; ra.phase_adjust = 0.0;
    { Timer t(p, DY4_K_RDS_PLL, st); CU(dy4_launch_rds_pll(ra, st)); }
    Timer t_bb(p, DY4_K_RDS_BASEBAND, st);                        // until the end of this function: resampler, RRC, carry, decoding
    CU(dy4_launch_rds_resample(ra, st));
    Dy4TailArgs ta{};                                             // history of the band-pass output for the next call
    ta.if_in = p->rds_f; ta.if_stride = (long long)p->ws_stride; ta.n_if = n_if; ta.if_tail = p->rds_tail; ta.n_streams = p->n_streams;
    CU(dy4_launch_tails(ta, st));
    p->rds_n_out = ra.n_out;
    p->rds_call_n += ra.n_out;
    p->if_abs += n_if;
    return run_rds_decode(p, st);
}

// the PLL of one sub-chunk: pilot (call rows) -> NCO row (call rows), scratch in the sub-chunk's workspace set.
// `parts`: prediction + table (or, for the direct loop, the reciprocal pre-pass), the serial loop, the NCO pass — queued on
// different streams (process_device).
int run_pll(dy4_pipeline* p, const SubChunk& c, cudaStream_t st, int parts)
{
    const dy4_mode_params_t& m = p->mp;
    auto& w = p->ws[c.set];
    const size_t off = (size_t)c.b * m.if_per_block;
    Dy4PllArgs pa;
    pa.in = p->c_pilot + off; pa.in_stride = (long long)p->c_stride; pa.nco = p->c_nco + off; pa.nco_stride = (long long)p->c_stride;
    pa.theta = w.theta; pa.inv = w.inv; pa.wide_stride = (long long)p->ws_stride; pa.nco0 = p->ws_nco0 + (size_t)c.set * p->n_streams;
    pa.state = p->pll_state; pa.n = c.nb * m.if_per_block; pa.n_streams = p->n_streams;
    pa.tab = w.tab; pa.tab_stride = 2 * (long long)p->ws_stride; pa.risk = p->pll_risk; pa.tstart = p->ws_nco0 + (size_t)(dy4_pipeline::NSETS + c.set) * p->n_streams;
    if (w.tab) {
        pa.pred_out = p->pred_state + (size_t)c.set * p->n_streams * 8;
        pa.pred_in = p->pred_state + (size_t)((c.set + dy4_pipeline::NSETS - 1) % dy4_pipeline::NSETS) * p->n_streams * 8;
        pa.need = p->pred_state + (size_t)(8 * dy4_pipeline::NSETS + c.set) * p->n_streams;      // written by the loop NSETS launches back, complete by now
        pa.pred_carry = c.pred_carry; pa.fresh = c.fresh; pa.nco0b = p->ws_nco0 + (size_t)(2 * dy4_pipeline::NSETS + c.set) * p->n_streams;
    }
    pa.freq = 19e3f; pa.Fs = m.if_Fs; pa.ncoScale = 2.0f; pa.phaseAdjust = 0.0f; pa.normBandwidth = 0.01f;   // project.cpp:99-102
    { Timer t(p, (parts & DY4_PLL_LOOP) ? DY4_K_PLL : DY4_K_PLL_AUX, st); CU(dy4_launch_pll_parts(pa, st, parts)); }
    return DY4_OK;
}

// back half on the main stream: IF, NCO, stereo band -> audio / PCM; leaves the mixed-signal history
int run_back(dy4_pipeline* p, const SubChunk& c, size_t pcm_stride, size_t audio_stride, cudaStream_t st)
{
    const dy4_mode_params_t& m = p->mp;
    const int n_if = c.nb * m.if_per_block, n_audio = c.nb * m.audio_per_block;
    const size_t off = (size_t)c.b * m.if_per_block;
    const Hist h = if_history(p, c.b);
    Dy4AudioArgs aa;
    aa.if_in = p->c_if + off; aa.if_stride = (long long)p->c_stride; aa.if_tail = h.tail; aa.if_tail_stride = h.stride;
    aa.nco = p->stereo ? p->c_nco + off : nullptr; aa.sband = p->stereo ? p->c_sband + off : nullptr; aa.bb_stride = (long long)p->c_stride; aa.mix_tail = p->mix_tail;
    aa.audio = c.audio; aa.audio_stride = (long long)audio_stride; aa.pcm = c.pcm; aa.pcm_stride = (long long)pcm_stride;
    aa.n_if = n_if; aa.n_audio = n_audio; aa.n_streams = p->n_streams; aa.stereo = p->stereo;
    aa.up = m.audio_upsample; aa.down = m.audio_decim; aa.exact = (p->flags & DY4_FLAG_EXACT_AUDIO) ? 1 : 0;
    aa.mode = p->mode; aa.taps_poly = p->d_taps_poly; aa.up_pad = p->up_pad; aa.neg_zero2 = kNegZero2;
    if (c.audio || c.pcm) { Timer t(p, DY4_K_AUDIO, st); CU(dy4_launch_audio(aa, st)); }
    if (p->stereo) {
        Dy4TailArgs ta{};
        ta.nco = p->c_nco + off; ta.sband = p->c_sband + off; ta.bb_stride = (long long)p->c_stride; ta.n_if = n_if; ta.mix_tail = p->mix_tail;
        ta.n_streams = p->n_streams;
        { Timer t(p, DY4_K_TAILS, st); CU(dy4_launch_tails(ta, st)); }
    }
    p->last_n_if = n_if; p->last_off = off;
    return DY4_OK;
}

// Optional per-sub-chunk hooks of the host-facing path: wait for that sub-chunk's upload before its FIR work,
// start the download of its PCM after its back half.  Arguments: first block and number of blocks of the sub-chunk.
struct Hooks { std::function<int(int, int, int, cudaStream_t)> before_front, after_back; };   // (index, first block, blocks, the stream the sub-chunk's kernels are on)

// The sub-chunks of a stereo job grow geometrically — 1, 2, 4, ... blocks up to the workspace size, then that size —
// so that (a) only one block's FIR work and, on the host path, one block's upload precede the first PLL launch, and
// every later upload (PCIe moves a block ~2x faster than the PLL consumes it) lands before it is needed, while
// (b) most of the job runs in large launches (few PLL launches).  Mono: uniform sub-chunks.
// (Ending the job on small sub-chunks as well — so that little follows the last serial loop — was measured: the two extra
// launches cost more than the shorter tail saves, 7.34 against 7.20 ms per step.)
std::vector<std::pair<int, int>> plan_subchunks(int n_blocks, int sb, bool geometric)
{
    std::vector<std::pair<int, int>> v;
    int b = 0;
    if (geometric && n_blocks >= 8)
        for (int g = 1; g < sb && b + g < n_blocks; g *= 2) { v.push_back({b, g}); b += g; }
    for (; b < n_blocks; b += sb) v.push_back({b, std::min(sb, n_blocks - b)});
    return v;
}

// RDS: RRC rows of the whole call (read back by dy4_pipeline_rds_read), sized before its first window
int rds_begin_call(dy4_pipeline* p, int n_blocks)
{
    if (!(p->flags & DY4_FLAG_RDS)) return DY4_OK;
    const size_t need = ((size_t)n_blocks * p->mp.if_per_block * 19 + 119) / 120 + 8;
    if (need > p->rds_out_cap) {
        if (p->s_rds) CU(cudaStreamSynchronize(p->s_rds));
        cudaFree(p->rds_out); p->rds_out = nullptr;
        CU(cudaMalloc(&p->rds_out, (size_t)p->n_streams * 2 * need * sizeof(float)));
        p->rds_out_cap = need;
    }
    p->rds_call_n = 0;
    return DY4_OK;
}

// One call: n_blocks whole blocks of every stream, input on the device.
int process_device(dy4_pipeline* p, const uint8_t* d_iq, size_t row_stride, int n_blocks,
                   int16_t* d_pcm, float* d_audio, float* d_if, cudaStream_t st,
                   size_t pcm_stride, size_t audio_stride, size_t if_stride, const Hooks* hooks = nullptr)
{
    const dy4_mode_params_t& m = p->mp;
    const int ch = p->stereo ? 2 : 1;
    int rc = ensure_workspace(p, n_blocks);
    if (rc) return rc;
    if ((rc = ensure_call_rows(p, n_blocks))) return rc;
    const bool overlap = p->pipelined && p->stereo && !hooks && p->s_main;        // this call is not joined to the caller's stream
    // the geometric ramp serves a call that starts on an idle device; an overlapped call's first sub-chunks are prepared while
    // the loops of the call before are still running, and uniform sub-chunks keep the predictions two whole launches ahead
    const auto plan = plan_subchunks(n_blocks, p->ws_blocks, p->stereo && !(p->flags & DY4_FLAG_DEBUG_ROWS) && !(overlap && !p->pll_fresh));
    if (overlap) {                                     // the other set of call rows: the previous call's are still being read
        std::swap(p->c_if, p->rows_alt[0]); std::swap(p->c_pilot, p->rows_alt[1]); std::swap(p->c_sband, p->rows_alt[2]); std::swap(p->c_nco, p->rows_alt[3]);
        p->call_parity ^= 1;
    }
    p->call_hist = p->if_tail;
    auto sub = [&](int b, int nb, long long seq) {
        SubChunk c;
        c.b = b; c.nb = nb;
        c.pcm = d_pcm ? d_pcm + (size_t)b * m.audio_per_block * ch : nullptr;
        c.audio = d_audio ? d_audio + (size_t)b * m.audio_per_block * ch : nullptr;
        c.set = p->stereo ? (int)(seq % dy4_pipeline::NSETS) : 0;
        return c;
    };
    // the IF history the NEXT call starts from, and the caller's copy of the IF rows: after everything that reads them
    auto finish_call = [&]() -> int {
        if (d_if) CU(cudaMemcpy2DAsync(d_if, if_stride * sizeof(float), p->c_if, p->c_stride * sizeof(float),
                                       (size_t)n_blocks * m.if_per_block * sizeof(float), p->n_streams, cudaMemcpyDeviceToDevice, st));
        Dy4TailArgs ta{};
        ta.if_in = p->c_if; ta.if_stride = (long long)p->c_stride; ta.n_if = n_blocks * m.if_per_block; ta.if_tail = p->if_tail; ta.n_streams = p->n_streams;
        { Timer t(p, DY4_K_TAILS, st); CU(dy4_launch_tails(ta, st)); }
        return DY4_OK;
    };
    if (!p->stereo) {                                  // mono: no serial stage, one stream
        for (size_t i = 0; i < plan.size(); i++, p->seq++) {
            const int b = plan[i].first;
            const SubChunk c = sub(b, plan[i].second, p->seq);
            if (hooks && (rc = hooks->before_front((int)i, b, c.nb, st))) return rc;
            if ((rc = run_fir(p, d_iq, row_stride, b, c.nb, st))) return rc;
            if ((rc = run_back(p, c, pcm_stride, audio_stride, st))) return rc;
            if (hooks && (rc = hooks->after_back((int)i, b, c.nb, st))) return rc;
        }
        return finish_call();
    }
    // stereo: software pipeline over four streams and NSETS = 3 workspace sets.
    //   main stream:  fir(piece 0) fir(piece 1) ...                               ... if_tail      FIR kernels of the whole call, up front
    //   aux stream:        prep(0)  prep(1)  prep(2) ...                                          prediction + table (FP64)
    //   PLL stream:                loop(0)   loop(1)   loop(2) ...                                the serial loops, back to back
    //   back stream:                        back(0)   back(1)  ...  back(last)                    NCO row + audio
    // FIR PIECES are full-wave launches over many sub-chunks: with the input on the device, piece 0 is the first sub-chunk
    // (one block: the first serial loop starts after 1/48 of the FIR work) and piece 1 everything else (more pieces were
    // measured: their smaller launches cost more than the earlier start of sub-chunks 1..3 saves, 6.97 against 6.69 ms); on the
    // host path (hooks) every sub-chunk is its own piece, gated on ITS upload, and the back halves stay on the main stream.
    // fir(piece) -> prep(c) -> loop(c) -> back(c) by events; prep(c) also waits for back(c - NSETS), the last reader of its
    // workspace set.  The loops' stream has the highest priority, aux and back the next, the FIR launches the lowest.
    std::vector<std::pair<int, int>> pieces;           // (first sub-chunk, sub-chunks)
    if (hooks) for (size_t i = 0; i < plan.size(); i++) pieces.push_back({(int)i, 1});
    else {                                             // {first sub-chunk}, {the rest}
        pieces.push_back({0, 1});
        if (plan.size() > 1) pieces.push_back({1, (int)plan.size() - 1});
    }
    while (p->ev_fir.size() < pieces.size()) { cudaEvent_t e; CU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming)); p->ev_fir.push_back(e); }
    CU(cudaEventRecord(p->ev_in, st));
    CU(cudaStreamWaitEvent(p->s_pll, p->ev_in, 0));     // the PLL stream starts after whatever precedes this call on `st`
    cudaStream_t const caller = st;
    if (p->s_main) {                                   // SM partition / pipelined calls: the FIR launches run on a stream of the pipeline's own
        CU(cudaStreamWaitEvent(p->s_main, p->ev_in, 0));
        st = p->s_main;
        if (overlap) {
            // this call's rows, and the IF history slot it writes for the next call (a ring of three), were last read by the
            // call before the previous one
            CU(cudaStreamWaitEvent(st, p->ev_call_done[p->call_parity], 0));
        }
    }
    const bool fresh_call = p->pll_fresh;              // no sample processed since create / reset: the streams start in this call
    p->pll_fresh = false;
    // back halves still to be queued (at most NSETS - 1 of them)
    struct Pending { SubChunk c; int i; };
    std::vector<Pending> pend;
    // the back halves run on a stream of their own when the input is on the device (on the main stream they would sit behind
    // every FIR launch of the call, and so would the predictions that wait for their workspace set)
    cudaStream_t sb = hooks ? st : p->s_back;
    if (sb != st) CU(cudaStreamWaitEvent(sb, p->ev_in, 0));
    auto flush_back = [&]() -> int {
        const Pending q = pend.front();
        pend.erase(pend.begin());
        CU(cudaStreamWaitEvent(sb, p->ev_pll[q.c.set], 0));
        int r2;
        if ((r2 = run_pll(p, q.c, sb, DY4_PLL_NCO))) return r2;
        if ((r2 = run_back(p, q.c, pcm_stride, audio_stride, sb))) return r2;
        CU(cudaEventRecord(p->ev_back[q.c.set], sb));
        if (hooks && (r2 = hooks->after_back(q.i, q.c.b, q.c.nb, sb))) return r2;
        return DY4_OK;
    };
    std::vector<int> piece_of(plan.size(), 0);
    for (size_t k = 0; k < pieces.size(); k++) for (int j = 0; j < pieces[k].second; j++) piece_of[pieces[k].first + j] = (int)k;
    size_t next_piece = 0;
    auto queue_piece = [&]() -> int {
        const int s0 = pieces[next_piece].first, ns = pieces[next_piece].second;
        const int b0 = plan[s0].first, nb = plan[s0 + ns - 1].first + plan[s0 + ns - 1].second - b0;
        int r2;
        if (hooks && (r2 = hooks->before_front(s0, b0, nb, st))) return r2;
        if ((r2 = run_fir(p, d_iq, row_stride, b0, nb, st))) return r2;
        CU(cudaEventRecord(p->ev_fir[next_piece], st));
        next_piece++;
        return DY4_OK;
    };
    // input on the device: every FIR piece is queued before anything else of the call (nothing they wait for; the back halves
    // queued on the same stream later wait for the loops).  Host path: a piece is queued when its sub-chunk comes up, so that
    // the back halves (and the downloads behind them) are not held up by uploads still to come.
    if (!hooks) while (next_piece < pieces.size()) if ((rc = queue_piece())) return rc;
    if (overlap) {                                     // the IF rows are complete on this stream: the caller's copy, and the next call's history
        if (d_if) CU(cudaMemcpy2DAsync(d_if, if_stride * sizeof(float), p->c_if, p->c_stride * sizeof(float),
                                       (size_t)n_blocks * m.if_per_block * sizeof(float), p->n_streams, cudaMemcpyDeviceToDevice, st));
        Dy4TailArgs ta{};
        ta.if_in = p->c_if; ta.if_stride = (long long)p->c_stride; ta.n_if = n_blocks * m.if_per_block; ta.if_tail = p->if_tail_alt[0]; ta.n_streams = p->n_streams;
        { Timer t(p, DY4_K_TAILS, st); CU(dy4_launch_tails(ta, st)); }
        float* const hist = p->if_tail;                // p->call_hist keeps pointing at it for the rest of this call
        p->if_tail = p->if_tail_alt[0]; p->if_tail_alt[0] = p->if_tail_alt[1]; p->if_tail_alt[1] = hist;
        CU(cudaEventRecord(p->ev_main_done, st));
    }
    const long long seq0 = p->seq;
    for (size_t i = 0; i < plan.size(); i++, p->seq++) {
        SubChunk c = sub(plan[i].first, plan[i].second, p->seq);
        while (next_piece <= (size_t)piece_of[i]) if ((rc = queue_piece())) return rc;
        cudaEvent_t ev_fir = p->ev_fir[piece_of[i]];
        // Table-driven PLL (dy4_pll.cu).  Sub-chunk 0 starts from the exact carried state; at the START of a stream its first
        // samples go through the direct loop while the PLL acquires lock, and sub-chunk 1 is predicted from the exact state
        // too (its prediction waits for the loop of sub-chunk 0, on the PLL stream); from sub-chunk 2 on the prediction
        // carries its own state and runs on the aux stream, beside the serial loop of the sub-chunks before.
        // A CONTINUING stream (not the first call after create / reset) needs none of that: the loop of the previous call has
        // finished, so sub-chunk 0 is predicted from the exact state, sub-chunk 1 carries on from that prediction, and nothing
        // but the serial loops is queued on the PLL stream.  Should the signal have jumped between two calls, the loop notices
        // (its picks stop being certain) and finishes that launch with the direct loop's steps.
        // PIPELINED calls: the loops of the previous call are still running when this call's first prediction is queued, so it
        // too carries on from the previous prediction (which ended where this call starts).
        if (fresh_call) c.pred_carry = i < 2 ? 0 : (i == 2 ? 1 : 2);
        else if (overlap) c.pred_carry = 2;
        else c.pred_carry = i == 0 ? 0 : (i == 1 ? 1 : 2);
        c.fresh = (i == 0 && fresh_call) ? 1536 : 0;                              // DY4_TAB_EARLY (dy4_plltab.h)
        const bool prep_on_pll = p->pll_table && i == 1 && fresh_call;
        if (p->flags & DY4_FLAG_RDS) {                           // the RDS branch needs only the IF rows: beside everything else
            CU(cudaStreamWaitEvent(p->s_rds, ev_fir, 0));
            if ((rc = run_rds(p, c, p->s_rds))) return rc;
            CU(cudaEventRecord(p->ev_rds, p->s_rds));
        }
        const bool reuse = p->seq - seq0 >= dy4_pipeline::NSETS || seq0 > 0;      // this workspace set has been used before: wait for its last reader
        if (p->pll_table && !prep_on_pll) {
            // prediction + table on the AUX stream: FP64-bound, they run beside the FP32-bound FIR kernels instead of in line
            // with them.  They read the pilot rows of their FIR piece; the previous prediction (same stream) and, NSETS
            // launches back, the loop's turn report are complete.
            CU(cudaStreamWaitEvent(p->s_aux, ev_fir, 0));
            if (reuse) CU(cudaStreamWaitEvent(p->s_aux, p->ev_back[c.set], 0));
            if (i == 2 && fresh_call) CU(cudaStreamWaitEvent(p->s_aux, p->ev_prep1, 0));
            if ((rc = run_pll(p, c, p->s_aux, DY4_PLL_PREP))) return rc;
            CU(cudaEventRecord(p->ev_prep[c.set], p->s_aux));
            CU(cudaStreamWaitEvent(p->s_pll, p->ev_prep[c.set], 0));
        } else {
            CU(cudaStreamWaitEvent(p->s_pll, ev_fir, 0));
            if (reuse) CU(cudaStreamWaitEvent(p->s_pll, p->ev_back[c.set], 0));
            if ((rc = run_pll(p, c, p->s_pll, DY4_PLL_PREP))) return rc;          // direct loop: reciprocals; fresh stream: sub-chunk 1 from the exact state
            if (prep_on_pll) CU(cudaEventRecord(p->ev_prep1, p->s_pll));
        }
        if ((rc = run_pll(p, c, p->s_pll, DY4_PLL_LOOP))) return rc;
        CU(cudaEventRecord(p->ev_pll[c.set], p->s_pll));
        pend.push_back({c, (int)i});
        if ((int)pend.size() >= dy4_pipeline::NSETS && (rc = flush_back())) return rc;
    }
    while (!pend.empty()) if ((rc = flush_back())) return rc;
    if (overlap) {                                     // no join: the call is done when its last back half (and RDS sub-chunk) is
        if (p->flags & DY4_FLAG_RDS) CU(cudaStreamWaitEvent(sb, p->ev_rds, 0));
        CU(cudaEventRecord(p->ev_call_done[p->call_parity], sb));
        return DY4_OK;
    }
    if (sb != st) {                                    // join: the call's last back half
        CU(cudaEventRecord(p->ev_fir_all, sb));
        CU(cudaStreamWaitEvent(st, p->ev_fir_all, 0));
    }
    if (p->flags & DY4_FLAG_RDS) CU(cudaStreamWaitEvent(st, p->ev_rds, 0));
    if ((rc = finish_call())) return rc;
    if (st != caller) {
        CU(cudaEventRecord(p->ev_out, st));
        CU(cudaStreamWaitEvent(caller, p->ev_out, 0));
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
    p->pipelined = (flags & DY4_FLAG_PIPELINED) && stereo;

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
    if (flags & DY4_FLAG_RDS) {
        if (mode != 0 || !stereo) { dy4_set_error("DY4_FLAG_RDS: the model defines the RDS path for mode 0 stereo only"); delete p; return DY4_ERR_ARG; }
        const int U = 19, NT = DY4_NTAPS * U;                      // fmMonoBlock.py:61-67,514-515: 1919 taps, Fc 3 kHz, gain 19, scipy default window
        std::vector<double> h(NT), r(DY4_NTAPS);
        dy4_firwin(NT, 0.0, 3e3 / (240e3 * U / 2), 1, h.data());
        std::vector<float> poly((size_t)DY4_NTAPS * 20, 0.0f), rrc(DY4_NTAPS);
        for (int j = 0; j < DY4_NTAPS; j++)
            for (int ph = 0; ph < U; ph++) poly[(size_t)j * 20 + ph] = (float)(h[(size_t)ph + (size_t)j * U] * U);
        dy4_rrc_taps(38000.0, DY4_NTAPS, r.data());               // sps 16 * 2375 (fmMonoBlock.py:66-67,690)
        for (int k = 0; k < DY4_NTAPS; k++) rrc[k] = (float)r[k];
        CU(cudaMalloc(&p->d_rds_poly, poly.size() * sizeof(float)));
        CU(cudaMemcpy(p->d_rds_poly, poly.data(), poly.size() * sizeof(float), cudaMemcpyHostToDevice));
        CU(dy4_upload_taps_rrc(rrc.data()));
        CU(cudaMalloc(&p->d_rds_rrc, rrc.size() * sizeof(float)));
        CU(cudaMemcpy(p->d_rds_rrc, rrc.data(), rrc.size() * sizeof(float), cudaMemcpyHostToDevice));
        CU(cudaMalloc(&p->rds_tail, S * DY4_IF_TAIL * sizeof(float)));
        CU(cudaMalloc(&p->rds_mix_tail, S * 2 * DY4_MIX_TAIL * sizeof(float)));
        CU(cudaMalloc(&p->rds_lp_tail, S * 2 * DY4_MIX_TAIL * sizeof(float)));
        CU(cudaMalloc(&p->rds_pll_state, S * 8 * sizeof(double)));
        CU(cudaMalloc(&p->rds_dec_state, S * DY4_RDS_STATE_INTS * sizeof(int)));
        CU(cudaMalloc(&p->rds_counts, S * 4 * sizeof(int)));
    }
    CU(cudaMalloc(&p->iq_tail, S * DY4_IQ_TAIL));
    CU(cudaMalloc(&p->if_tail, S * DY4_IF_TAIL * sizeof(float)));
    if (p->pipelined) for (auto& t : p->if_tail_alt) CU(cudaMalloc(&t, S * DY4_IF_TAIL * sizeof(float)));
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
    { const int rq = quiesce(p); if (rq) return rq; }
    return init_state(p, nullptr);
}

extern "C" int dy4_debug_pll_stats(long long* out4);

extern "C" int dy4_pipeline_destroy(dy4_pipeline_t* p)
{
    if (!p) return DY4_OK;
    cudaSetDevice(p->device);
    quiesce(p);
    if (p->pll_table && std::getenv("DY4_PLL_STATS")) {                       // development counters of the serial PLL loop
        long long v[4];
        if (dy4_debug_pll_stats(v) == 0 && v[1] > 0)
            fprintf(stderr, "dy4 pll stats: rows %lld groups %lld (%.2f rows/group) direct steps %lld (%.3f%%) %lld\n", v[0], v[1], (double)v[0] / (double)v[1],
                    v[2], 100.0 * (double)v[2] / (double)std::max(1LL, v[0]), v[3]);
    }
    for (auto& r : p->recs) { cudaEventDestroy(r.e0); cudaEventDestroy(r.e1); }
    for (auto e : p->pool) cudaEventDestroy(e);
    cudaFree(p->d_rf_taps); cudaFree(p->d_taps_poly);
    cudaFree(p->rds_f); cudaFree(p->rds_carrier); cudaFree(p->rds_nco_i); cudaFree(p->rds_nco_q); cudaFree(p->rds_theta); cudaFree(p->rds_lp); cudaFree(p->rds_out);
    cudaFree(p->rds_tail); cudaFree(p->rds_mix_tail); cudaFree(p->rds_lp_tail); cudaFree(p->rds_pll_state); cudaFree(p->d_rds_poly); cudaFree(p->d_rds_rrc);
    cudaFree(p->rds_acc); cudaFree(p->rds_dec_state); cudaFree(p->rds_counts); cudaFree(p->rds_events); cudaFree(p->rds_groups); cudaFree(p->rds_sym); cudaFree(p->rds_bits);
    if (p->s_rds) { cudaStreamDestroy(p->s_rds); cudaEventDestroy(p->ev_rds); }
    cudaFree(p->iq_tail); cudaFree(p->if_tail); cudaFree(p->mix_tail); cudaFree(p->pll_state);
    for (auto& w : p->ws) { cudaFree(w.theta); cudaFree(w.inv); cudaFree(w.tab); }
    cudaFree(p->c_if); cudaFree(p->c_pilot); cudaFree(p->c_sband); cudaFree(p->c_nco);
    for (auto r : p->rows_alt) cudaFree(r);
    for (auto t : p->if_tail_alt) cudaFree(t);
    for (auto e : p->ev_fir) cudaEventDestroy(e);
    for (auto e : {p->ev_call_done[0], p->ev_call_done[1], p->ev_main_done}) if (e) cudaEventDestroy(e);
    cudaFree(p->ws_nco0); cudaFree(p->pred_state); cudaFree(p->pll_risk);
    if (p->s_pll) {
        cudaStreamDestroy(p->s_pll);
        if (p->s_aux) cudaStreamDestroy(p->s_aux);
        if (p->s_main) { cudaStreamDestroy(p->s_main); cudaEventDestroy(p->ev_out); }
        if (p->s_back) { cudaStreamDestroy(p->s_back); cudaEventDestroy(p->ev_fir_all); }
        for (int i = 0; i < dy4_pipeline::NSETS; i++) { cudaEventDestroy(p->ev_back[i]); cudaEventDestroy(p->ev_pll[i]); cudaEventDestroy(p->ev_prep[i]); }
        if (p->ev_prep1) cudaEventDestroy(p->ev_prep1);
        cudaEventDestroy(p->ev_in);
    }
    cudaFree(p->d_stage); cudaFree(p->d_pcm_stage); cudaFree(p->d_audio_stage);
    for (auto& h : p->hset) { cudaFree(h.d_iq); cudaFree(h.d_pcm); cudaFree(h.d_audio); if (h.done) cudaEventDestroy(h.done); }
    if (p->ev_h_up) cudaEventDestroy(p->ev_h_up);
    for (auto e : p->ev_up) cudaEventDestroy(e);
    for (auto e : p->ev_done) cudaEventDestroy(e);
    if (p->streams_ready) { cudaStreamDestroy(p->s_compute); cudaStreamDestroy(p->s_h2d); cudaStreamDestroy(p->s_d2h); }
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
    // the call rows (16 bytes per IF sample) hold a window of the call; very large calls go through in several windows
    const int window = std::min(n_blocks, max_call_blocks(p));
    { const int rc = rds_begin_call(p, n_blocks); if (rc) return rc; }
    for (int w0 = 0; w0 < n_blocks; w0 += window) {
        const size_t ao = (size_t)w0 * m.audio_per_block * ch;
        const int rc = process_device(p, d_iq + (size_t)w0 * m.block_size, row_stride_bytes, std::min(window, n_blocks - w0),
                                      d_pcm ? d_pcm + ao : nullptr, d_audio ? d_audio + ao : nullptr, d_if ? d_if + (size_t)w0 * m.if_per_block : nullptr,
                                      (cudaStream_t)stream, astride, astride, (size_t)n_blocks * m.if_per_block);
        if (rc) return rc;
    }
    return DY4_OK;
}

extern "C" int dy4_pipeline_flush(dy4_pipeline_t* p, void* stream)
{
    if (!p) { dy4_set_error("dy4_pipeline_flush: bad arguments"); return DY4_ERR_ARG; }
    if (!p->pipelined || !p->ev_main_done) return DY4_OK;       // calls are joined to their stream when they end
    CU(cudaSetDevice(p->device));
    cudaStream_t st = (cudaStream_t)stream;
    for (cudaEvent_t e : {p->ev_main_done, p->ev_call_done[0], p->ev_call_done[1], p->ev_rds}) if (e) CU(cudaStreamWaitEvent(st, e, 0));
    return DY4_OK;
}

namespace {
int ensure_host_streams(dy4_pipeline* p)
{
    if (p->streams_ready) return DY4_OK;
    CU(cudaStreamCreateWithFlags(&p->s_compute, cudaStreamNonBlocking));
    CU(cudaStreamCreateWithFlags(&p->s_h2d, cudaStreamNonBlocking));
    CU(cudaStreamCreateWithFlags(&p->s_d2h, cudaStreamNonBlocking));
    p->streams_ready = true;
    return DY4_OK;
}

// Host path of a DY4_FLAG_PIPELINED pipeline.  A call (or a window of it) is uploaded WHOLE into one of two staging sets — when
// the caller's rows are contiguous (row stride = row bytes) as ONE 1-D copy, 55 GB/s on this pool where row-pitched copies reach
// 50 — then queued as an overlapped device call (process_device: not joined), and its PCM comes back on the download stream
// once it is done.  Nothing waits for anything but data: the upload of call k+1 runs beside the kernels of call k, so a step
// costs max(upload, kernels) instead of their partly overlapped sum.  The call returns when everything is QUEUED; the host
// arrays are valid after dy4_pipeline_sync (or any other entry point, which all drain the queue).
int process_host_overlapped(dy4_pipeline* p, const uint8_t* h_iq, size_t row_stride_bytes, int n_blocks, int16_t* h_pcm, float* h_audio, int chunk_blocks)
{
    const dy4_mode_params_t& m = p->mp;
    const int ch = 2;
    const size_t S = (size_t)p->n_streams;
    size_t budget = 4ull << 30;
    if (const char* e = std::getenv("DY4_STAGE_BYTES")) budget = std::strtoull(e, nullptr, 10);
    int window = chunk_blocks > 0 ? chunk_blocks : (int)std::max<size_t>(1, budget / (S * m.block_size));
    window = std::min(std::min(window, n_blocks), max_call_blocks(p));
    int rc = ensure_host_streams(p);
    if (rc) return rc;
    if (p->hset_blocks < window || (h_audio && !p->hset_audio)) {
        if ((rc = quiesce(p))) return rc;
        for (auto& h : p->hset) {
            cudaFree(h.d_iq); cudaFree(h.d_pcm); cudaFree(h.d_audio);
            h.d_iq = nullptr; h.d_pcm = nullptr; h.d_audio = nullptr; h.used = false;
            CU(cudaMalloc(&h.d_iq, S * window * m.block_size));
            CU(cudaMalloc(&h.d_pcm, S * window * m.audio_per_block * ch * sizeof(int16_t)));
            if (h_audio) CU(cudaMalloc(&h.d_audio, S * window * m.audio_per_block * ch * sizeof(float)));
            if (!h.done) CU(cudaEventCreateWithFlags(&h.done, cudaEventDisableTiming));
        }
        if (!p->ev_h_up) CU(cudaEventCreateWithFlags(&p->ev_h_up, cudaEventDisableTiming));
        p->hset_blocks = window; p->hset_audio = h_audio != nullptr;
    }
    const size_t total_audio = (size_t)n_blocks * m.audio_per_block * ch;                // host output row length
    if ((rc = rds_begin_call(p, n_blocks))) return rc;
    for (int w0 = 0; w0 < n_blocks; w0 += window) {
        const int wn = std::min(window, n_blocks - w0);
        auto& h = p->hset[p->host_parity ^= 1];
        if (h.used) CU(cudaEventSynchronize(h.done));                                    // the call that last used this set has delivered its PCM
        const size_t in_bytes = (size_t)wn * m.block_size, na = (size_t)wn * m.audio_per_block * ch;
        const uint8_t* src = h_iq + (size_t)w0 * m.block_size;
        if (row_stride_bytes == in_bytes) CU(cudaMemcpyAsync(h.d_iq, src, S * in_bytes, cudaMemcpyHostToDevice, p->s_h2d));
        else CU(cudaMemcpy2DAsync(h.d_iq, in_bytes, src, row_stride_bytes, in_bytes, S, cudaMemcpyHostToDevice, p->s_h2d));
        CU(cudaEventRecord(p->ev_h_up, p->s_h2d));
        CU(cudaStreamWaitEvent(p->s_compute, p->ev_h_up, 0));
        if ((rc = process_device(p, h.d_iq, in_bytes, wn, h_pcm ? h.d_pcm : nullptr, h_audio ? h.d_audio : nullptr, nullptr, p->s_compute, na, na, 0))) return rc;
        if ((rc = dy4_pipeline_flush(p, p->s_d2h))) return rc;                           // the download stream waits for this call (and those before it)
        const size_t ao = (size_t)w0 * m.audio_per_block * ch;
        if (h_pcm) {
            if (total_audio == na) CU(cudaMemcpyAsync(h_pcm, h.d_pcm, S * na * sizeof(int16_t), cudaMemcpyDeviceToHost, p->s_d2h));
            else CU(cudaMemcpy2DAsync(h_pcm + ao, total_audio * sizeof(int16_t), h.d_pcm, na * sizeof(int16_t), na * sizeof(int16_t), S, cudaMemcpyDeviceToHost, p->s_d2h));
        }
        if (h_audio) {
            if (total_audio == na) CU(cudaMemcpyAsync(h_audio, h.d_audio, S * na * sizeof(float), cudaMemcpyDeviceToHost, p->s_d2h));
            else CU(cudaMemcpy2DAsync(h_audio + ao, total_audio * sizeof(float), h.d_audio, na * sizeof(float), na * sizeof(float), S, cudaMemcpyDeviceToHost, p->s_d2h));
        }
        CU(cudaEventRecord(h.done, p->s_d2h));
        h.used = true;
    }
    return DY4_OK;
}
}  // namespace

extern "C" int dy4_pipeline_sync(dy4_pipeline_t* p)
{
    if (!p) { dy4_set_error("dy4_pipeline_sync: bad arguments"); return DY4_ERR_ARG; }
    CU(cudaSetDevice(p->device));
    return quiesce(p);
}

extern "C" int dy4_pipeline_process_host(dy4_pipeline_t* p, const uint8_t* h_iq, size_t row_stride_bytes, int n_blocks,
                                         int16_t* h_pcm, float* h_audio, int chunk_blocks)
{
    if (!p || !h_iq || n_blocks < 0) { dy4_set_error("dy4_pipeline_process_host: bad arguments"); return DY4_ERR_ARG; }
    if (n_blocks == 0) return DY4_OK;
    const dy4_mode_params_t& m = p->mp;
    if (row_stride_bytes < (size_t)n_blocks * m.block_size) { dy4_set_error("dy4_pipeline_process_host: row stride too small"); return DY4_ERR_ARG; }
    CU(cudaSetDevice(p->device));
    if (p->pipelined) return process_host_overlapped(p, h_iq, row_stride_bytes, n_blocks, h_pcm, h_audio, chunk_blocks);
    // device-path calls are asynchronous on the CALLER's stream, this call runs on streams of the pipeline's own: whatever is still
    // queued must be through before it touches the carried state
    { const int rq = quiesce(p); if (rq) return rq; }
    const int ch = p->stereo ? 2 : 1;
    const size_t S = (size_t)p->n_streams;
    // A "window" of blocks is resident in device staging at a time (DY4_STAGE_BYTES of input, default 4 GiB, or
    // chunk_blocks if given).  Inside a window every sub-chunk has its own upload and download: all uploads are queued
    // at once on the copy stream, the front half of sub-chunk c waits only for ITS bytes, and its PCM starts back to
    // the host as soon as its back half is done — so the copies hide under the (PLL-bound) compute of the neighbours.
    size_t budget = 4ull << 30;
    if (const char* e = std::getenv("DY4_STAGE_BYTES")) budget = std::strtoull(e, nullptr, 10);
    int window = chunk_blocks > 0 ? chunk_blocks : (int)std::max<size_t>(1, budget / (S * m.block_size));
    window = std::min(std::min(window, n_blocks), max_call_blocks(p));
    { const int rs = ensure_host_streams(p); if (rs) return rs; }
    if (p->stage_blocks < window || (h_audio && !p->stage_audio)) {
        { const int rq = quiesce(p); if (rq) return rq; }
        cudaFree(p->d_stage); cudaFree(p->d_pcm_stage); cudaFree(p->d_audio_stage);
        p->d_stage = nullptr; p->d_pcm_stage = nullptr; p->d_audio_stage = nullptr;
        CU(cudaMalloc(&p->d_stage, S * window * m.block_size));
        CU(cudaMalloc(&p->d_pcm_stage, S * window * m.audio_per_block * ch * sizeof(int16_t)));
        if (h_audio) CU(cudaMalloc(&p->d_audio_stage, S * window * m.audio_per_block * ch * sizeof(float)));
        p->stage_blocks = window;
        p->stage_audio = h_audio != nullptr;
    }
    const size_t st_stride = (size_t)p->stage_blocks * m.block_size;                     // staging row strides
    const size_t out_stride = (size_t)p->stage_blocks * m.audio_per_block * ch;
    const size_t total_audio = (size_t)n_blocks * m.audio_per_block * ch;                // host output row length

    { const int rc = rds_begin_call(p, n_blocks); if (rc) return rc; }
    for (int w0 = 0; w0 < n_blocks; w0 += window) {
        const int wn = std::min(window, n_blocks - w0);
        int rc = ensure_workspace(p, wn);
        if (rc) return rc;
        const auto plan = plan_subchunks(wn, p->ws_blocks, p->stereo && !(p->flags & DY4_FLAG_DEBUG_ROWS));
        const int nsub = (int)plan.size();
        while ((int)p->ev_up.size() < nsub) {
            cudaEvent_t a, d;
            CU(cudaEventCreateWithFlags(&a, cudaEventDisableTiming));
            CU(cudaEventCreateWithFlags(&d, cudaEventDisableTiming));
            p->ev_up.push_back(a); p->ev_done.push_back(d);
        }
        for (int i = 0; i < nsub; i++) {                                                 // queue every upload of the window
            const int b = plan[i].first, nb = plan[i].second;
            CU(cudaMemcpy2DAsync(p->d_stage + (size_t)b * m.block_size, st_stride, h_iq + (size_t)(w0 + b) * m.block_size, row_stride_bytes,
                                 (size_t)nb * m.block_size, S, cudaMemcpyHostToDevice, p->s_h2d));
            CU(cudaEventRecord(p->ev_up[i], p->s_h2d));
        }
        Hooks hooks;
        hooks.before_front = [&](int i, int, int, cudaStream_t cs) -> int {
            CU(cudaStreamWaitEvent(cs, p->ev_up[i], 0));
            return DY4_OK;
        };
        hooks.after_back = [&](int i, int b, int nb, cudaStream_t cs) -> int {
            const size_t na = (size_t)nb * m.audio_per_block * ch, off = (size_t)b * m.audio_per_block * ch;
            CU(cudaEventRecord(p->ev_done[i], cs));
            CU(cudaStreamWaitEvent(p->s_d2h, p->ev_done[i], 0));
            if (h_pcm) CU(cudaMemcpy2DAsync(h_pcm + (size_t)w0 * m.audio_per_block * ch + off, total_audio * sizeof(int16_t), p->d_pcm_stage + off,
                                            out_stride * sizeof(int16_t), na * sizeof(int16_t), S, cudaMemcpyDeviceToHost, p->s_d2h));
            if (h_audio) CU(cudaMemcpy2DAsync(h_audio + (size_t)w0 * m.audio_per_block * ch + off, total_audio * sizeof(float), p->d_audio_stage + off,
                                              out_stride * sizeof(float), na * sizeof(float), S, cudaMemcpyDeviceToHost, p->s_d2h));
            return DY4_OK;
        };
        rc = process_device(p, p->d_stage, st_stride, wn, h_pcm ? p->d_pcm_stage : nullptr, h_audio ? p->d_audio_stage : nullptr,
                            nullptr, p->s_compute, out_stride, out_stride, 0, &hooks);
        if (rc) return rc;
        CU(cudaStreamSynchronize(p->s_h2d));                                             // the staging is reused by the next window
        CU(cudaStreamSynchronize(p->s_compute));
        CU(cudaStreamSynchronize(p->s_d2h));
    }
    return DY4_OK;
}

extern "C" int dy4_pipeline_debug_buffers(dy4_pipeline_t* p, const float** d_pilot, const float** d_nco, size_t* stride, int* n_if)
{
    if (!p || !p->stereo || !p->c_pilot) { dy4_set_error("dy4_pipeline_debug_buffers: no stereo sub-chunk processed yet"); return DY4_ERR_ARG; }
    CU(cudaSetDevice(p->device));
    { const int rq = quiesce(p); if (rq) return rq; }
    if (d_pilot) *d_pilot = p->c_pilot + p->last_off;
    if (d_nco) *d_nco = p->c_nco + p->last_off;
    if (stride) *stride = p->c_stride;
    if (n_if) *n_if = p->last_n_if;
    return DY4_OK;
}

extern "C" void* dy4_pinned_alloc(size_t bytes)
{
    void* p = nullptr;
    cudaError_t e = cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocDefault);
    if (e != cudaSuccess) { dy4_cuda_fail(e, "dy4_pinned_alloc"); return nullptr; }
    return p;
}
extern "C" void dy4_pinned_free(void* p) { if (p) cudaFreeHost(p); }

extern "C" int dy4_pipeline_rds_read(dy4_pipeline_t* p, float* d_rrc_i, float* d_rrc_q, size_t row_stride, int* n_samples, void* stream)
{
    if (!p || !(p->flags & DY4_FLAG_RDS) || !n_samples) { dy4_set_error("dy4_pipeline_rds_read: pipeline was not created with DY4_FLAG_RDS"); return DY4_ERR_ARG; }
    CU(cudaSetDevice(p->device));
    *n_samples = p->rds_call_n;
    if (p->rds_call_n <= 0) return DY4_OK;
    if (row_stride < (size_t)p->rds_call_n) { dy4_set_error("dy4_pipeline_rds_read: row stride smaller than the sample count"); return DY4_ERR_ARG; }
    cudaStream_t st = (cudaStream_t)stream;
    if (p->pipelined && p->ev_rds) CU(cudaStreamWaitEvent(st, p->ev_rds, 0));       // not joined at the end of the call
    if (d_rrc_i) CU(cudaMemcpy2DAsync(d_rrc_i, row_stride * sizeof(float), p->rds_out, 2 * p->rds_out_cap * sizeof(float),
                                      (size_t)p->rds_call_n * sizeof(float), p->n_streams, cudaMemcpyDeviceToDevice, st));
    if (d_rrc_q) CU(cudaMemcpy2DAsync(d_rrc_q, row_stride * sizeof(float), p->rds_out + p->rds_out_cap, 2 * p->rds_out_cap * sizeof(float),
                                      (size_t)p->rds_call_n * sizeof(float), p->n_streams, cudaMemcpyDeviceToDevice, st));
    return DY4_OK;
}

extern "C" int dy4_pipeline_rds_bounds(dy4_pipeline_t* p, int* max_symbols, int* max_bits, int* max_events)
{
    if (!p || !(p->flags & DY4_FLAG_RDS)) { dy4_set_error("dy4_pipeline_rds_bounds: pipeline was not created with DY4_FLAG_RDS"); return DY4_ERR_ARG; }
    const long long ns = p->rds_blocks_since_drain * (DY4_RDS_BLOCK / 16);
    if (max_symbols) *max_symbols = (int)ns;
    if (max_bits) *max_bits = (int)((ns + 1) / 2 + 1);
    if (max_events) *max_events = (int)((ns + 1) / 2 + 1);
    return DY4_OK;
}

extern "C" int dy4_pipeline_rds_drain(dy4_pipeline_t* p, int8_t* h_symbols, size_t sym_stride, int8_t* h_bits, size_t bits_stride,
                                      int32_t* h_events, size_t ev_stride, int32_t* h_groups, size_t grp_stride, int32_t* h_counts)
{
    if (!p || !(p->flags & DY4_FLAG_RDS) || !h_counts) { dy4_set_error("dy4_pipeline_rds_drain: pipeline was not created with DY4_FLAG_RDS"); return DY4_ERR_ARG; }
    CU(cudaSetDevice(p->device));
    CU(cudaStreamSynchronize(p->s_rds));
    const size_t S = (size_t)p->n_streams;
    int ms, mb, me;
    dy4_pipeline_rds_bounds(p, &ms, &mb, &me);
    if ((h_symbols && sym_stride < (size_t)ms) || (h_bits && bits_stride < (size_t)mb) || (h_events && ev_stride < (size_t)me) ||
        (h_groups && grp_stride < (size_t)me)) {
        dy4_set_error("dy4_pipeline_rds_drain: row strides smaller than dy4_pipeline_rds_bounds");
        return DY4_ERR_ARG;
    }
    std::vector<int> c(S * 4);
    CU(cudaMemcpy(c.data(), p->rds_counts, c.size() * sizeof(int), cudaMemcpyDeviceToHost));
    for (size_t s = 0; s < S; s++) for (int j = 0; j < 4; j++) h_counts[s * 4 + j] = c[s * 4 + j];
    if (p->rds_blocks_since_drain > 0) {
        if (h_symbols) CU(cudaMemcpy2D(h_symbols, sym_stride, p->rds_sym, p->rds_sym_cap, (size_t)ms, S, cudaMemcpyDeviceToHost));
        if (h_bits) CU(cudaMemcpy2D(h_bits, bits_stride, p->rds_bits, p->rds_bits_cap, std::min((size_t)mb, p->rds_bits_cap), S, cudaMemcpyDeviceToHost));
        if (h_events) CU(cudaMemcpy2D(h_events, ev_stride * 16, p->rds_events, p->rds_ev_cap * 16, std::min((size_t)me, p->rds_ev_cap) * 16, S, cudaMemcpyDeviceToHost));
        if (h_groups) CU(cudaMemcpy2D(h_groups, grp_stride * 16, p->rds_groups, p->rds_grp_cap * 16, std::min((size_t)me, p->rds_grp_cap) * 16, S, cudaMemcpyDeviceToHost));
    }
    CU(cudaMemset(p->rds_counts, 0, S * 4 * sizeof(int)));
    p->rds_blocks_since_drain = 0;
    return DY4_OK;
}

extern "C" int dy4_pipeline_sm_partition(dy4_pipeline_t* p, int* loop_sms, int* rest_sms)
{
    if (!p) { dy4_set_error("dy4_pipeline_sm_partition: bad arguments"); return DY4_ERR_ARG; }
    if (loop_sms) *loop_sms = p->s_main ? p->loop_sms : 0;
    if (rest_sms) *rest_sms = p->s_main ? p->rest_sms : 0;
    return DY4_OK;
}

extern "C" int dy4_pipeline_pll_risk(dy4_pipeline_t* p, int32_t* h_counts, int reset)
{
    if (!p || !h_counts) { dy4_set_error("dy4_pipeline_pll_risk: bad arguments"); return DY4_ERR_ARG; }
    CU(cudaSetDevice(p->device));
    { const int rq = quiesce(p); if (rq) return rq; }
    const size_t S = (size_t)p->n_streams;
    if (!p->pll_risk) { std::memset(h_counts, 0, S * sizeof(int32_t)); return DY4_OK; }     // mono, or the direct loop: nothing counted
    CU(cudaMemcpy(h_counts, p->pll_risk, S * sizeof(int32_t), cudaMemcpyDeviceToHost));
    if (reset) CU(cudaMemset(p->pll_risk, 0, S * sizeof(int32_t)));
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
    const bool trace = std::getenv("DY4_TRACE") != nullptr && !p->recs.empty();      // development: the timeline of every timed launch
    for (auto& r : p->recs) {
        CU(cudaEventSynchronize(r.e1));
        float t = 0.f;
        CU(cudaEventElapsedTime(&t, r.e0, r.e1));
        if (trace) {
            float t0 = 0.f;
            CU(cudaEventElapsedTime(&t0, p->recs.front().e0, r.e0));
            std::fprintf(stderr, "dy4-trace k=%d start=%.4f end=%.4f ms\n", r.k, t0, t0 + t);
        }
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
// RDS part of a checkpoint (DY4_FLAG_RDS): a 16-byte header {if_abs, samples held for the decoder}, then per stream the
// band-pass / mixer / resampler histories, the carrier PLL state, the decoder state and the RRC samples the decoder has
// not consumed yet (less than one model block).  Decoded symbols / bits / events are results, not state: drain first.
static size_t rds_state_bytes(const dy4_pipeline_t* p)
{
    if (!(p->flags & DY4_FLAG_RDS)) return 0;
    const size_t S = (size_t)p->n_streams;
    return 16 + S * (DY4_IF_TAIL * sizeof(float) + 4 * DY4_MIX_TAIL * sizeof(float) + 8 * sizeof(double) +
                     DY4_RDS_STATE_INTS * sizeof(int) + DY4_RDS_BLOCK * sizeof(float));
}

extern "C" size_t dy4_pipeline_state_size(const dy4_pipeline_t* p)
{
    if (!p) return 0;
    const size_t S = (size_t)p->n_streams;
    return S * (DY4_IQ_TAIL + (DY4_IF_TAIL + DY4_MIX_TAIL + 8) * sizeof(float)) + rds_state_bytes(p);
}

extern "C" int dy4_pipeline_get_state(dy4_pipeline_t* p, void* host_buf)
{
    if (!p || !host_buf) return DY4_ERR_ARG;
    CU(cudaSetDevice(p->device));
    { const int rq = quiesce(p); if (rq) return rq; }
    const size_t S = (size_t)p->n_streams;
    char* o = (char*)host_buf;
    CU(cudaMemcpy(o, p->iq_tail, S * DY4_IQ_TAIL, cudaMemcpyDeviceToHost)); o += S * DY4_IQ_TAIL;
    CU(cudaMemcpy(o, p->if_tail, S * DY4_IF_TAIL * sizeof(float), cudaMemcpyDeviceToHost)); o += S * DY4_IF_TAIL * sizeof(float);
    CU(cudaMemcpy(o, p->mix_tail, S * DY4_MIX_TAIL * sizeof(float), cudaMemcpyDeviceToHost)); o += S * DY4_MIX_TAIL * sizeof(float);
    CU(cudaMemcpy(o, p->pll_state, S * 8 * sizeof(float), cudaMemcpyDeviceToHost)); o += S * 8 * sizeof(float);
    if (p->flags & DY4_FLAG_RDS) {
        long long hdr[2] = {p->if_abs, (long long)p->rds_left};
        std::memcpy(o, hdr, 16); o += 16;
        CU(cudaMemcpy(o, p->rds_tail, S * DY4_IF_TAIL * sizeof(float), cudaMemcpyDeviceToHost)); o += S * DY4_IF_TAIL * sizeof(float);
        CU(cudaMemcpy(o, p->rds_mix_tail, S * 2 * DY4_MIX_TAIL * sizeof(float), cudaMemcpyDeviceToHost)); o += S * 2 * DY4_MIX_TAIL * sizeof(float);
        CU(cudaMemcpy(o, p->rds_lp_tail, S * 2 * DY4_MIX_TAIL * sizeof(float), cudaMemcpyDeviceToHost)); o += S * 2 * DY4_MIX_TAIL * sizeof(float);
        CU(cudaMemcpy(o, p->rds_pll_state, S * 8 * sizeof(double), cudaMemcpyDeviceToHost)); o += S * 8 * sizeof(double);
        CU(cudaMemcpy(o, p->rds_dec_state, S * DY4_RDS_STATE_INTS * sizeof(int), cudaMemcpyDeviceToHost)); o += S * DY4_RDS_STATE_INTS * sizeof(int);
        std::memset(o, 0, S * DY4_RDS_BLOCK * sizeof(float));
        if (p->rds_left > 0)
            CU(cudaMemcpy2D(o, DY4_RDS_BLOCK * sizeof(float), p->rds_acc + p->rds_consumed, p->rds_acc_cap * sizeof(float),
                            (size_t)p->rds_left * sizeof(float), S, cudaMemcpyDeviceToHost));
    }
    return DY4_OK;
}

extern "C" int dy4_pipeline_set_state(dy4_pipeline_t* p, const void* host_buf)
{
    if (p) p->pll_fresh = false;                     // whatever the restored streams are, they are not at sample 0 by construction
    if (!p || !host_buf) return DY4_ERR_ARG;
    CU(cudaSetDevice(p->device));
    { const int rq = quiesce(p); if (rq) return rq; }
    const size_t S = (size_t)p->n_streams;
    const char* o = (const char*)host_buf;
    CU(cudaMemcpy(p->iq_tail, o, S * DY4_IQ_TAIL, cudaMemcpyHostToDevice)); o += S * DY4_IQ_TAIL;
    CU(cudaMemcpy(p->if_tail, o, S * DY4_IF_TAIL * sizeof(float), cudaMemcpyHostToDevice)); o += S * DY4_IF_TAIL * sizeof(float);
    CU(cudaMemcpy(p->mix_tail, o, S * DY4_MIX_TAIL * sizeof(float), cudaMemcpyHostToDevice)); o += S * DY4_MIX_TAIL * sizeof(float);
    CU(cudaMemcpy(p->pll_state, o, S * 8 * sizeof(float), cudaMemcpyHostToDevice)); o += S * 8 * sizeof(float);
    if (p->flags & DY4_FLAG_RDS) {
        long long hdr[2];
        std::memcpy(hdr, o, 16); o += 16;
        if (hdr[1] < 0 || hdr[1] >= DY4_RDS_BLOCK) { dy4_set_error("dy4_pipeline_set_state: corrupt RDS header"); return DY4_ERR_ARG; }
        p->if_abs = hdr[0];
        CU(cudaMemcpy(p->rds_tail, o, S * DY4_IF_TAIL * sizeof(float), cudaMemcpyHostToDevice)); o += S * DY4_IF_TAIL * sizeof(float);
        CU(cudaMemcpy(p->rds_mix_tail, o, S * 2 * DY4_MIX_TAIL * sizeof(float), cudaMemcpyHostToDevice)); o += S * 2 * DY4_MIX_TAIL * sizeof(float);
        CU(cudaMemcpy(p->rds_lp_tail, o, S * 2 * DY4_MIX_TAIL * sizeof(float), cudaMemcpyHostToDevice)); o += S * 2 * DY4_MIX_TAIL * sizeof(float);
        CU(cudaMemcpy(p->rds_pll_state, o, S * 8 * sizeof(double), cudaMemcpyHostToDevice)); o += S * 8 * sizeof(double);
        CU(cudaMemcpy(p->rds_dec_state, o, S * DY4_RDS_STATE_INTS * sizeof(int), cudaMemcpyHostToDevice)); o += S * DY4_RDS_STATE_INTS * sizeof(int);
        int rc = grow_rows(p->rds_acc, p->rds_acc_cap, (size_t)DY4_RDS_BLOCK + 64, S, 1, p->s_rds ? p->s_rds : (cudaStream_t)0);
        if (rc) return rc;
        p->rds_left = (int)hdr[1]; p->rds_consumed = 0;
        if (p->rds_left > 0)
            CU(cudaMemcpy2D(p->rds_acc, p->rds_acc_cap * sizeof(float), o, DY4_RDS_BLOCK * sizeof(float),
                            (size_t)p->rds_left * sizeof(float), S, cudaMemcpyHostToDevice));
        CU(cudaMemset(p->rds_counts, 0, S * 4 * sizeof(int)));
        p->rds_blocks_since_drain = 0; p->rds_call_n = 0; p->rds_n_out = 0;
    }
    return DY4_OK;
}
