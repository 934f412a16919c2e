// dy4_rds.cu — RDS filtering front end (BASELINE.json configs[3]; SURVEY.md §8d "config 4").
//
// The reference implements RDS only in its Python model (model/fmMonoBlock.py:673-696, float64):
//   convolve(54-60 kHz)  -> squaringNonlinearity -> convolve(113.5-114.5 kHz) -> fmPll(114 kHz, ncoScale 0.5, bw 0.001)
//   delayBlock(50) -> pointwiseMultiply(nco I/Q, delayed, 1) -> resampler(19, 120, 1919-tap low-pass) -> convolve(RRC)
// The two band-pass filters reuse k_twin_bpf (dy4_bpf.cu, fused arithmetic, squared input for the second); this file
// holds the rest.  Arithmetic: float32 FIRs with fused multiply-add, float64 PLL — the model's PLL is float64 throughout,
// so unlike the stereo PLL it is not chaotic and needs no bit-exact inputs; parity is a tolerance, written in the tests.
//
// PLL: with a float64 feedback pair the detector atan2(-x sin t, x cos t) is just the reduced phase, -t (x>0) or
// pi - t (x<0) wrapped to (-pi, pi] — no transcendental in the recurrence at all.  One thread per stream; the NCO
// I/Q rows are an elementwise pass over the stored phase arguments.
#include "dy4_common.cuh"
#include "dy4_kernels.h"
#include "dy4_internal.h"

#include <cstdlib>

namespace {

constexpr double kTwoPiHi = 6.283185307179586, kInvTwoPi = 0.15915494309189535;
constexpr double kPi = 3.141592653589793;

// The detector needs the phase argument reduced to (-pi, pi].  Reducing w*k + phase from scratch every sample (it reaches
// 1e5 rad) would put a Cody-Waite reduction on the serial chain; instead the reduced phase r is part of the carried
// state and advanced incrementally, r += (w mod 2pi) + dphase, wrapped by at most one turn.  Its rounding (1e-16 per
// step) is far inside the path's tolerance — the model itself is float64 — while the unreduced argument theta, which the
// NCO rows are computed from, is still formed as w*k + phase exactly as the model does.  Per sample the dependent chain
// is: select eD -> integ -> dphase -> r -> wrap.
__global__ void __launch_bounds__(128)
k_rds_pll(const float* __restrict__ carrier, long long stride, double* __restrict__ theta, long long wide_stride,
          double* __restrict__ state, int n, int n_streams, double w, double Kp, double Ki)
{
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n_streams) return;
    double* st = state + (long long)s * 8;
    double integ = st[0], phase = st[1], k = st[2];
    double r = st[3];                                             // the reduced phase is CARRIED (0 at stream start), so any chunking gives the same samples
    const double w_red = w - kTwoPiHi * rint(w * kInvTwoPi);      // w is below 2 pi here: exact
    const float* x = carrier + (long long)s * stride;
    double* y = theta + (long long)s * wide_stride;
    auto step = [&](float xv) -> double {
        const double rw = r + w_red, rw2 = rw - kTwoPiHi;          // beside the chain: both wraps of the next phase
        // fmPll: errorD = atan2(-x sin t, x cos t) = -t (x > 0) or pi - t wrapped (x < 0); 0 for x == 0 (fmMonoBlock.py:355-362)
        const double e_neg = (r > 0.0 ? kPi : -kPi) - r;
        const double eD = xv > 0.0f ? -r : (xv < 0.0f ? e_neg : 0.0);
        integ = fma(Ki, eD, integ);                                // :363
        const double dph = fma(Kp, eD, integ);
        phase += dph;                                              // :364
        k += 1.0;                                                  // :367
        const double a = rw + dph, b = rw2 + dph;
        r = a > kPi ? b : a;
        return fma(w, k, phase);                                   // :368
    };
    int i = 0;
    for (; i < n && (i & 3); i++) y[i] = step(x[i]);
    float4 v = i + 4 <= n ? *reinterpret_cast<const float4*>(x + i) : make_float4(0, 0, 0, 0);
    for (; i + 4 <= n; i += 4) {
        const float4 cur = v;
        if (i + 8 <= n) v = *reinterpret_cast<const float4*>(x + i + 4);
        const double t0 = step(cur.x), t1 = step(cur.y), t2 = step(cur.z), t3 = step(cur.w);
        *reinterpret_cast<double2*>(y + i) = make_double2(t0, t1);
        *reinterpret_cast<double2*>(y + i + 2) = make_double2(t2, t3);
    }
    for (; i < n; i++) y[i] = step(x[i]);
    st[0] = integ; st[1] = phase; st[2] = k; st[3] = r;
}

// nco_i[k] = cos(theta[k-1]*scale + adj), nco_q[k] = sin(...); [0] from the carried state (fmMonoBlock.py:353-354,373-376)
__global__ void __launch_bounds__(256)
k_rds_nco(const double* __restrict__ theta, long long wide_stride, double* __restrict__ state, float* __restrict__ nco_i,
          float* __restrict__ nco_q, long long stride, int n, double scale, double adj)
{
    const int k = blockIdx.y * blockDim.x + threadIdx.x;
    const int s = blockIdx.x;
    if (k > n) return;
    double* st = state + (long long)s * 8;
    if (k == 0) { nco_i[(long long)s * stride] = (float)st[4]; nco_q[(long long)s * stride] = (float)st[5]; return; }
    double sn, cs;
    sincos(theta[(long long)s * wide_stride + k - 1] * scale + adj, &sn, &cs);
    if (k == n) { st[6] = cs; st[7] = sn; return; }               // next chunk's [0]; moved into st[4..5] by k_rds_finish
    nco_i[(long long)s * stride + k] = (float)cs;
    nco_q[(long long)s * stride + k] = (float)sn;
}

// 19/120 polyphase resampler of the mixed signals (nco * rds_f delayed by 50), I and Q as one packed pair per tap.
// One thread per output m (absolute index m_first + j): y[m] = sum_j h[phase + up*j] * mixed[floor(m*down/up) - j].
template <int NT>
__global__ void __launch_bounds__(NT)
k_rds_resample(Dy4RdsArgs a)
{
    extern __shared__ __align__(16) float2 sm_mix[];
    const int s = blockIdx.x, tid = threadIdx.x;
    const int j0 = blockIdx.y * NT;
    const int j_last = min(j0 + NT, a.n_out) - 1;
    const long long mA = a.m_first + j0, mB = a.m_first + j_last;
    // IF-rate span needed, as indices relative to the chunk start
    const long long i_lo = (mA * a.down) / a.up - a.if_abs - (DY4_NTAPS - 1);
    const long long i_hi = (mB * a.down) / a.up - a.if_abs;
    const int span = (int)(i_hi - i_lo + 1);
    const float* rf = a.rds_f + (long long)s * a.stride;
    const float* rt = a.rds_tail + (long long)s * DY4_IF_TAIL;
    const float* ni = a.nco_i + (long long)s * a.stride;
    const float* nq = a.nco_q + (long long)s * a.stride;
    const float* mt = a.mix_tail + (long long)s * 2 * DY4_MIX_TAIL;
    constexpr int DELAY = DY4_NTAPS / 2;                              // RDS_delay_state has int(101/2) = 50 entries
    for (int p = tid; p < span; p += NT) {
        const long long i = i_lo + p;
        float2 v;
        if (i < 0) { v.x = mt[DY4_MIX_TAIL + i]; v.y = mt[2 * DY4_MIX_TAIL + i]; }
        else if (i >= a.n_if) v = make_float2(0.f, 0.f);
        else {
            const long long d = i - DELAY;
            const float x = d < 0 ? rt[DY4_IF_TAIL + d] : rf[d];
            v.x = ni[i] * x; v.y = nq[i] * x;                         // pointwiseMultiply(nco, delayed, 1)
        }
        sm_mix[p] = v;
    }
    // the polyphase table [tap][phase] (101 x 20 floats) next to the samples in shared memory
    float* s_taps = reinterpret_cast<float*>(sm_mix + a.span_max);
    for (int t = tid; t < DY4_NTAPS * a.up_pad; t += NT) s_taps[t] = __ldg(a.taps_poly + t);
    __syncthreads();
    const int j = j0 + tid;
    if (j >= a.n_out) return;
    const long long m = a.m_first + j, n = m * a.down;
    const int phase = (int)(n % a.up);
    const int base = (int)(n / a.up - a.if_abs - i_lo);
    const u64* xs = reinterpret_cast<const u64*>(sm_mix) + base;
    const float* hp = s_taps + phase;
    u64 acc = 0ull;
#pragma unroll 8
    for (int t = 0; t < DY4_NTAPS; t++) {
        const float h = hp[t * a.up_pad];
        acc = ffma2(xs[-t], pk2(h, h), acc);
    }
    float yi, yq;
    upk2(acc, yi, yq);
    float* lp = a.lp + (long long)s * 2 * a.lp_stride;
    lp[j] = yi; lp[a.lp_stride + j] = yq;
}

// RRC: 101-tap FIR at the resampler's output rate on the (I, Q) pair; history from lp_tail.  A tile of NT*R outputs
// plus its 100-sample history is staged in shared memory as (I, Q) pairs; every thread keeps R consecutive outputs in
// registers and walks its window once, newest sample first (taps ascending), one packed FFMA2 per tap and output.
__constant__ float2 c_rrc2[DY4_NTAPS + 3];       // (h, h) pairs of the RRC taps

template <int R, int NT>
__global__ void __launch_bounds__(NT)
k_rds_rrc(Dy4RdsArgs a)
{
    constexpr int T = NT * R, HALO = DY4_NTAPS - 1;
    __shared__ __align__(16) float2 sm[T + HALO + 2 * ((T + HALO) / R) + 8];
    const int s = blockIdx.x, tid = threadIdx.x, j0 = blockIdx.y * T;
    const float* lp = a.lp + (long long)s * 2 * a.lp_stride;
    const float* lt = a.lp_tail + (long long)s * 2 * DY4_MIX_TAIL;
    for (int p = tid; p < T + HALO; p += NT) {                    // logical pair p <-> sample j0 - HALO + p
        const int i = j0 - HALO + p;
        float2 v;
        if (i < 0) v = make_float2(lt[DY4_MIX_TAIL + i], lt[2 * DY4_MIX_TAIL + i]);
        else if (i < a.n_out) v = make_float2(lp[i], lp[a.lp_stride + i]);
        else v = make_float2(0.f, 0.f);
        sm[p + 2 * (p / R)] = v;                                  // two pad pairs per R: neighbouring threads' loads on distinct banks
    }
    __syncthreads();
    // output r of this thread is sample j0 + tid*R + r; tap k reads logical pair tid*R + r + HALO - k = tid*R + q, q = r + HALO - k
    const u64* w = reinterpret_cast<const u64*>(sm) + (R + 2) * tid;
    const u64* hh = reinterpret_cast<const u64*>(c_rrc2);
    u64 acc[R];
#pragma unroll
    for (int r = 0; r < R; r++) acc[r] = 0ull;
#pragma unroll
    for (int q = R - 1 + HALO; q >= 0; q--) {
        const u64 x = w[q + 2 * (q / R)];
#pragma unroll
        for (int r = 0; r < R; r++) {
            const int k = r + HALO - q;
            if (k >= 0 && k < DY4_NTAPS) acc[r] = ffma2(x, hh[k], acc[r]);
        }
    }
    float* oi = a.out_i + (long long)s * a.out_stride;
    float* oq = a.out_q + (long long)s * a.out_stride;
#pragma unroll
    for (int r = 0; r < R; r++) {
        const int j = j0 + tid * R + r;
        if (j < a.n_out) { float yi, yq; upk2(acc[r], yi, yq); oi[j] = yi; oq[j] = yq; }
    }
}

// carry: last 128 mixed I/Q samples, last 128 resampler outputs, NCO state
__global__ void __launch_bounds__(128)
k_rds_finish(Dy4RdsArgs a)
{
    const int s = blockIdx.x, t = threadIdx.x;
    constexpr int DELAY = DY4_NTAPS / 2;
    const float* rf = a.rds_f + (long long)s * a.stride;
    const float* rt = a.rds_tail + (long long)s * DY4_IF_TAIL;
    float* mt = a.mix_tail + (long long)s * 2 * DY4_MIX_TAIL;
    float* lt = a.lp_tail + (long long)s * 2 * DY4_MIX_TAIL;
    const float* lp = a.lp + (long long)s * 2 * a.lp_stride;
    {   // mixed tail: samples n_if-128 .. n_if-1 (n_if >= 128 + DELAY is guaranteed: a chunk is at least one block)
        const long long i = (long long)a.n_if - DY4_MIX_TAIL + t;
        const long long d = i - DELAY;
        const float x = d < 0 ? rt[DY4_IF_TAIL + d] : rf[d];
        const float vi = a.nco_i[(long long)s * a.stride + i] * x, vq = a.nco_q[(long long)s * a.stride + i] * x;
        // resampler-output tail: shift in this chunk's outputs (n_out may be smaller than 128)
        float li, lq;
        const int src = a.n_out - DY4_MIX_TAIL + t;
        if (src >= 0) { li = lp[src]; lq = lp[a.lp_stride + src]; }
        else { li = lt[DY4_MIX_TAIL + src]; lq = lt[2 * DY4_MIX_TAIL + src]; }
        __syncthreads();
        mt[t] = vi; mt[DY4_MIX_TAIL + t] = vq;
        lt[t] = li; lt[DY4_MIX_TAIL + t] = lq;
    }
    if (t == 0) { double* st = a.pll_state + (long long)s * 8; st[4] = st[6]; st[5] = st[7]; }
}

// ---- back half: symbol timing, Manchester + differential decoding, frame synchronisation ---------------------------
// model/fmSupportLib.py:209-247 (manchesterEncoded) and model/fmMonoBlock.py:78-122, 157-284, 699-730, per MODEL BLOCK of
// 3 040 RRC samples (190 symbols), with the model's state hand-off between blocks and its quirks kept (see the
// restatement in oracle/rds.py).  Serial, branchy and tiny: one thread per stream.
constexpr int RB_SPS = 16, RB_BLOCK = DY4_RDS_BLOCK, RB_NSYM = RB_BLOCK / RB_SPS, RB_MAXBITS = (RB_NSYM + 1) / 2;
enum { RS_BLOCK = 0, RS_MIDX, RS_FOUND, RS_SYMSTATE, RS_ERR1, RS_ERR2, RS_BITSTATE, RS_WINDEX, RS_SYNCED, RS_OFFSET,
       RS_NUMSYNCED, RS_BITPOS, RS_LASTPOS, RS_WSTATE_LEN, RS_WSTATE, RS_MSG_A, RS_MSG_B, RS_MSG_C, RS_MSG_D, RS_PAD };   // DY4_RDS_STATE_INTS = 20

// rows of the parity-check matrix as masks over the 26-bit window, window element j = bit j (fmMonoBlock.py:183-192)
__constant__ unsigned c_rds_rows[10] = {
    0x39BE401u, 0x337C802u, 0x1F47404u, 0x0730C08u, 0x0E61810u, 0x257D420u, 0x3344C40u, 0x1F37C80u, 0x3E6F900u, 0x3CDF200u};
// syndromes of the offset words A, B, C, C', D as bit masks (syndrome element i = bit i), and the offsets each may follow
__constant__ unsigned c_rds_syn[5] = {0x06Fu, 0x0AFu, 0x0E9u, 0x0CFu, 0x069u};
__constant__ unsigned c_rds_pred[5] = {1u << 4, 1u << 0, 1u << 1, 1u << 1, (1u << 2) | (1u << 3)};

__global__ void __launch_bounds__(32)
k_rds_decode(Dy4RdsDecodeArgs a)
{
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= a.n_streams) return;
    int* st = a.state + (long long)s * DY4_RDS_STATE_INTS;
    int* cnt = a.counts + (long long)s * 4;
    int block_count = st[RS_BLOCK], m_index = st[RS_MIDX], found = st[RS_FOUND], symbol_state = st[RS_SYMSTATE];
    int errors1 = st[RS_ERR1], errors2 = st[RS_ERR2], bit_state = st[RS_BITSTATE], window_index = st[RS_WINDEX];
    int synced = st[RS_SYNCED], offset_state = st[RS_OFFSET], num_synced = st[RS_NUMSYNCED], bit_pos = st[RS_BITPOS];
    int last_pos = st[RS_LASTPOS], wstate_len = st[RS_WSTATE_LEN];
    unsigned wstate = (unsigned)st[RS_WSTATE];                   // the previous bits' last 25, element j = bit j
    int msgs[4] = {st[RS_MSG_A], st[RS_MSG_B], st[RS_MSG_C], st[RS_MSG_D]};   // last A, B, C, D words while in sync; -1 = none (fmMonoBlock.py:596-600)
    int n_sym = cnt[0], n_bits = cnt[1], n_ev = cnt[2], n_grp = cnt[3];
    int* o_grp = a.groups + (long long)s * a.grp_stride * 4;
    int8_t* o_sym = a.sym + (long long)s * a.sym_stride;
    int8_t* o_bits = a.bits + (long long)s * a.bits_stride;
    int* o_ev = a.events + (long long)s * a.ev_stride * 4;
    int8_t sym[RB_NSYM];
    int8_t bits[RB_MAXBITS];

    for (int blk = 0; blk < a.n_blocks; blk++) {
        const float* sig = a.acc + (long long)s * a.acc_stride + (long long)blk * RB_BLOCK;
        // ---- manchesterEncoded (fmSupportLib.py:209-247)
        int idx = m_index;
        bool truncate = false;
        if (!found) {
            float mx = 0.f;
            idx = 0;
            for (int i = 0; i < 2 * RB_SPS; i++) {
                const float v = sig[i];
                if (fabsf(v) > mx) { mx = v; idx = i; found = 1; truncate = true; }    // `max = signal[i]`: signed, as the model
            }
        }
        for (int k = idx; k < RB_BLOCK; k += RB_SPS) sym[k / RB_SPS] = sig[k] < 0.f ? 0 : 1;
        // end-of-block sanity check on output[int(k/sps)] and output[int((k-1)/sps)] at the loop's final k: the second is the
        // same slot unless k is a multiple of sps, and the model writes `abs(x < threshold)`, i.e. a signed comparison
        const int k_last = idx + RB_SPS * ((RB_BLOCK - 1 - idx) / RB_SPS);
        const float out_last = sig[k_last];
        const float out_prev = (k_last % RB_SPS != 0 || k_last == 0) ? out_last : sig[k_last - RB_SPS];
        if (fabs((double)out_last) < 0.05 && (double)out_prev < 0.05) found = 0;
        m_index = k_last % RB_SPS;
        const int first = truncate ? 1 : 0;                      // a block that (re)acquired timing drops its first symbol
        const int ns = RB_NSYM - first;
        for (int i = 0; i < ns; i++) {
            if (n_sym < a.sym_cap) o_sym[n_sym] = sym[first + i];
            n_sym++;
        }
        const int8_t* sy = sym + first;

        if (block_count >= 5) {
            if (block_count < 10) {
                // ---- find_pattern (fmMonoBlock.py:78-92)
                for (int i = 1; i < ns; i += 2) {
                    const int c1 = sy[i], p1 = sy[i - 1], c2 = sy[i - 1], p2 = i != 1 ? sy[i - 2] : symbol_state;
                    errors1 += c1 == p1;
                    errors2 += c2 == p2;
                }
                symbol_state = sy[ns - 1];
            } else {
                // ---- decode (fmMonoBlock.py:97-122): Manchester, then differential
                const int start = errors1 > errors2 ? 0 : 1;
                int nb = 0;
                for (int i = start; i < ns; i += 2) {
                    const int cur = sy[i], prev = i != 0 ? sy[i - 1] : symbol_state;
                    const int b = (cur == 0 && prev == 1) ? 1 : 0;
                    bits[nb] = (int8_t)(b != bit_state);
                    if (n_bits < a.bits_cap) o_bits[n_bits] = bits[nb];
                    n_bits++;
                    nb++;
                    bit_state = b;
                }
                symbol_state = sy[ns - 1];
                // ---- get_window + frame_sync_receiver (fmMonoBlock.py:157-173, 176-284, 711-715)
                int widx = 0;
                while ((synced && widx < nb - 26) || (!synced && widx < nb - 1)) {
                    window_index += synced ? 26 : 1;
                    if (window_index >= nb) window_index -= nb;
                    unsigned w = 0;                                // window element j = bit j
                    if (window_index < 25) {
                        int j = 0;
                        for (int q = window_index; q < wstate_len; q++, j++) w |= ((wstate >> q) & 1u) << j;
                        for (int q = 0; q <= window_index; q++, j++) w |= (unsigned)bits[q] << j;
                    } else {
                        for (int j = 0; j < 26; j++) w |= (unsigned)bits[window_index - 25 + j] << j;
                    }
                    wstate = 0; wstate_len = 25;
                    for (int j = 0; j < 25; j++) wstate |= (unsigned)bits[nb - 25 + j] << j;
                    widx = window_index;

                    unsigned syn = 0;
                    for (int i = 0; i < 10; i++) syn |= (unsigned)(__popc(w & c_rds_rows[i]) & 1) << i;
                    int t = -1, msg_word = -1;                     // msg = [] unless a syndrome matches
                    for (int i = 0; i < 5; i++) if (syn == c_rds_syn[i]) t = i;
                    if (t >= 0) {
                        const int old_offset = offset_state;
                        if ((old_offset >= 0 && ((c_rds_pred[t] >> old_offset) & 1u)) || (old_offset < 0 && !synced)) synced = 1;
                        else if (synced) { synced = 0; num_synced = 0; }
                        const int false_pos = (bit_pos != last_pos + 26 && old_offset >= 0) ? 1 : 0;
                        int msg = 0;
                        for (int j = 0; j < 16; j++) msg = (msg << 1) | (int)((w >> j) & 1u);
                        msg_word = msg;
                        if (n_ev < a.ev_cap) { o_ev[4 * n_ev] = t; o_ev[4 * n_ev + 1] = bit_pos; o_ev[4 * n_ev + 2] = false_pos; o_ev[4 * n_ev + 3] = msg; }
                        n_ev++;
                        offset_state = synced ? t : -1;
                        last_pos = bit_pos;
                    }
                    bit_pos += synced ? 26 : 1;
                    if (num_synced > 3 && !synced) synced = 1;
                    // the glue of the model's main loop (fmMonoBlock.py:716-730): while in sync the word of the CURRENT
                    // offset state replaces that block's slot (an unmatched window empties it; C' never lands in the C
                    // slot), out of sync everything is dropped; a complete A,B,C,D set goes to the application layer
                    if (synced) {
                        const int slot = offset_state == 0 ? 0 : offset_state == 1 ? 1 : offset_state == 2 ? 2 : offset_state == 4 ? 3 : -1;
                        if (slot >= 0) msgs[slot] = msg_word;
                    } else {
                        msgs[0] = msgs[1] = msgs[2] = msgs[3] = -1;
                    }
                    if (msgs[0] >= 0 && msgs[1] >= 0 && msgs[2] >= 0 && msgs[3] >= 0) {
                        if (n_grp < a.grp_cap) { o_grp[4 * n_grp] = msgs[0]; o_grp[4 * n_grp + 1] = msgs[1]; o_grp[4 * n_grp + 2] = msgs[2]; o_grp[4 * n_grp + 3] = msgs[3]; }
                        n_grp++;
                    }
                }
            }
        }
        block_count++;
    }
    st[RS_BLOCK] = block_count; st[RS_MIDX] = m_index; st[RS_FOUND] = found; st[RS_SYMSTATE] = symbol_state;
    st[RS_ERR1] = errors1; st[RS_ERR2] = errors2; st[RS_BITSTATE] = bit_state; st[RS_WINDEX] = window_index;
    st[RS_SYNCED] = synced; st[RS_OFFSET] = offset_state; st[RS_NUMSYNCED] = num_synced; st[RS_BITPOS] = bit_pos;
    st[RS_LASTPOS] = last_pos; st[RS_WSTATE_LEN] = wstate_len; st[RS_WSTATE] = (int)wstate;
    st[RS_MSG_A] = msgs[0]; st[RS_MSG_B] = msgs[1]; st[RS_MSG_C] = msgs[2]; st[RS_MSG_D] = msgs[3];
    cnt[0] = n_sym; cnt[1] = n_bits; cnt[2] = n_ev; cnt[3] = n_grp;
}

// append this call's in-phase RRC samples to the per-stream accumulation row (after moving what the decoder left
// unconsumed — less than one model block — to the front)
__global__ void __launch_bounds__(256)
k_rds_append(const float* __restrict__ rrc_i, long long rrc_stride, int n_new, float* __restrict__ acc, long long acc_stride,
             int consumed, int left)
{
    const int s = blockIdx.x;
    float* row = acc + (long long)s * acc_stride;
    float keep[(DY4_RDS_BLOCK + 255) / 256];                     // left < one model block
    if (consumed > 0) {
#pragma unroll
        for (int j = 0; j < (DY4_RDS_BLOCK + 255) / 256; j++) { const int i = threadIdx.x + 256 * j; keep[j] = i < left ? row[consumed + i] : 0.f; }
        __syncthreads();
#pragma unroll
        for (int j = 0; j < (DY4_RDS_BLOCK + 255) / 256; j++) { const int i = threadIdx.x + 256 * j; if (i < left) row[i] = keep[j]; }
        __syncthreads();
    }
    const float* src = rrc_i + (long long)s * rrc_stride;
    for (int i = threadIdx.x; i < n_new; i += blockDim.x) row[left + i] = src[i];
}

}  // namespace

cudaError_t dy4_launch_rds_append(const float* rrc_i, long long rrc_stride, int n_new, float* acc, long long acc_stride,
                                  int consumed, int left, int n_streams, cudaStream_t st)
{
    if (n_streams <= 0) return cudaSuccess;
    k_rds_append<<<n_streams, 256, 0, st>>>(rrc_i, rrc_stride, n_new, acc, acc_stride, consumed, left);
    g_dy4_launches++;
    return cudaGetLastError();
}

cudaError_t dy4_launch_rds_decode(const Dy4RdsDecodeArgs& a, cudaStream_t st)
{
    if (a.n_streams <= 0 || a.n_blocks <= 0) return cudaSuccess;
    k_rds_decode<<<(a.n_streams + 31) / 32, 32, 0, st>>>(a);
    g_dy4_launches++;
    return cudaGetLastError();
}

cudaError_t dy4_upload_taps_rrc(const float* rrc)
{
    float2 h[DY4_NTAPS + 3] = {};
    for (int k = 0; k < DY4_NTAPS; k++) h[k] = make_float2(rrc[k], rrc[k]);
    return cudaMemcpyToSymbol(c_rrc2, h, sizeof(h));
}

cudaError_t dy4_launch_rds_pll(const Dy4RdsArgs& a, cudaStream_t st)
{
    if (a.n_if <= 0 || a.n_streams <= 0) return cudaSuccess;
    const int threads = 32;                            // one warp per 32 streams: the loop is bound by its dependent chain, not by throughput
    k_rds_pll<<<(a.n_streams + threads - 1) / threads, threads, 0, st>>>(a.carrier, a.stride, a.theta, a.wide_stride, a.pll_state, a.n_if, a.n_streams, a.w, a.Kp, a.Ki);
    g_dy4_launches++;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    dim3 grid(a.n_streams, (a.n_if + 1 + 255) / 256);
    k_rds_nco<<<grid, 256, 0, st>>>(a.theta, a.wide_stride, a.pll_state, a.nco_i, a.nco_q, a.stride, a.n_if, a.nco_scale, a.phase_adjust);
    g_dy4_launches++;
    return cudaGetLastError();
}

cudaError_t dy4_launch_rds_resample(const Dy4RdsArgs& a, cudaStream_t st)
{
    if (a.n_streams <= 0) return cudaSuccess;
    cudaError_t e;
    if (a.n_out > 0) {
        constexpr int NT = 128;
        const int span_max = (int)(((long long)(NT - 1) * a.down) / a.up) + DY4_NTAPS + 4;
        dim3 grid(a.n_streams, (a.n_out + NT - 1) / NT);
        Dy4RdsArgs b = a;
        b.span_max = span_max;
        k_rds_resample<NT><<<grid, NT, span_max * sizeof(float2) + DY4_NTAPS * a.up_pad * sizeof(float), st>>>(b);
        g_dy4_launches++;
        if ((e = cudaGetLastError()) != cudaSuccess) return e;
        constexpr int RR = 4, RNT = 128;
        dim3 g2(a.n_streams, (a.n_out + RR * RNT - 1) / (RR * RNT));
        k_rds_rrc<RR, RNT><<<g2, RNT, 0, st>>>(a);
        g_dy4_launches++;
        if ((e = cudaGetLastError()) != cudaSuccess) return e;
    }
    k_rds_finish<<<a.n_streams, DY4_MIX_TAIL, 0, st>>>(a);
    g_dy4_launches++;
    return cudaGetLastError();
}
