"""The reference's include/filter.h:17-34 by its own function names, on numpy arrays.

Same argument meaning and state handling as the C++ functions (state arrays are
updated in place, scalar state is returned), so parity tests read like calls
into the reference.  Every function runs on the GPU through the compatibility
tier of libdy4b200.so (include/dy4_b200.h); nothing here computes on the CPU.
"""
import ctypes as C

import numpy as np

from ._lib import lib, check


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _p(a):
    return C.c_void_p(a.ctypes.data)


def impulseResponseLPF(Fs, Fc, num_taps, upFactor=1):
    h = np.empty(num_taps, np.float32)
    check(lib.dy4_lpf_taps(Fs, Fc, num_taps, upFactor, _p(h)), "impulseResponseLPF")
    return h


def impulseResponseBPF(Fs, Fb, Fe, num_taps, upFactor=1):
    h = np.empty(num_taps, np.float32)
    check(lib.dy4_bpf_taps(Fs, Fb, Fe, num_taps, upFactor, _p(h)), "impulseResponseBPF")
    return h


def iqToFloat(raw):
    raw = np.ascontiguousarray(raw, np.uint8)
    out = np.empty(raw.size, np.float32)
    check(lib.dy4_iq_to_float(_p(raw), raw.size, _p(out)), "iqToFloat")
    return out


def convolveFIR(x, h):
    x, h = _f32(x), _f32(h)
    y = np.empty(x.size + h.size - 1, np.float32)
    check(lib.dy4_convolve_fir(_p(y), _p(x), x.size, _p(h), h.size), "convolveFIR")
    return y


def blockConvolveFIR(x, h, state):
    x, h = _f32(x), _f32(h)
    y = np.empty(x.size, np.float32)
    check(lib.dy4_block_fir(_p(y), _p(x), x.size, _p(h), h.size, _p(state), state.size), "blockConvolveFIR")
    return y


def downsampleBlockConvolveFIR(factor, x, h, state):
    x, h = _f32(x), _f32(h)
    y = np.empty(x.size // factor, np.float32)
    check(lib.dy4_decim_fir(factor, _p(y), _p(x), x.size, _p(h), h.size, _p(state), state.size), "downsampleBlockConvolveFIR")
    return y


def resampleBlockConvolveFIR(upFactor, downFactor, x, h, state):
    x, h = _f32(x), _f32(h)
    y = np.empty(int((x.size / np.float32(downFactor)) * upFactor) + 1, np.float32)
    n = C.c_size_t(0)
    check(lib.dy4_resample_fir(upFactor, downFactor, _p(y), C.byref(n), _p(x), x.size, _p(h), h.size, _p(state), state.size),
          "resampleBlockConvolveFIR")
    return y[:n.value].copy()


def fmDemodArctan(I, Q, prev_I, prev_Q):
    """-> (fm_demod, prev_I, prev_Q)"""
    I, Q = _f32(I), _f32(Q)
    out = np.empty(I.size, np.float32)
    pi, pq = C.c_float(prev_I), C.c_float(prev_Q)
    check(lib.dy4_fm_demod(_p(I), _p(Q), I.size, C.byref(pi), C.byref(pq), _p(out)), "fmDemodArctan")
    return out, pi.value, pq.value


def fmPLL(PLLin, freq, Fs, ncoScale, phaseAdjust, normBandwidth, state):
    """state: float32[6] = feedbackI, feedbackQ, integrator, phaseEst, trigOffset, nco_state; updated in place."""
    x = _f32(PLLin)
    nco = np.empty(x.size, np.float32)
    c = [C.c_float(float(v)) for v in state]
    check(lib.dy4_pll(_p(x), x.size, freq, Fs, ncoScale, phaseAdjust, normBandwidth, _p(nco), *[C.byref(v) for v in c]), "fmPLL")
    state[:] = [v.value for v in c]
    return nco


def downsample(data, factor):
    x = _f32(data)
    y = np.empty((x.size + factor - 1) // factor, np.float32)
    n = C.c_size_t(0)
    check(lib.dy4_downsample(_p(x), x.size, factor, _p(y), C.byref(n)), "downsample")
    return y[:n.value]


def upsample(data, factor):
    x = _f32(data)
    y = np.empty(x.size * factor, np.float32)
    n = C.c_size_t(0)
    check(lib.dy4_upsample(_p(x), x.size, factor, _p(y), C.byref(n)), "upsample")
    return y[:n.value]


def delayBlock(input_block, state_block):
    x = _f32(input_block)
    y = np.empty(x.size, np.float32)
    check(lib.dy4_delay_block(_p(x), x.size, _p(state_block), state_block.size, _p(y)), "delayBlock")
    return y


def pointwiseMultiply(a, b):
    a, b = _f32(a), _f32(b)
    y = np.empty(min(a.size, b.size), np.float32)
    n = C.c_size_t(0)
    check(lib.dy4_pointwise_multiply(_p(a), a.size, _p(b), b.size, _p(y), C.byref(n)), "pointwiseMultiply")
    return y


def pointwiseAdd(a, b):
    a, b = _f32(a), _f32(b)
    y = np.empty(a.size, np.float32)
    check(lib.dy4_pointwise_add(_p(a), _p(b), a.size, _p(y)), "pointwiseAdd")
    return y


def pointwiseSubtract(a, b):
    a, b = _f32(a), _f32(b)
    y = np.empty(a.size, np.float32)
    check(lib.dy4_pointwise_subtract(_p(a), _p(b), a.size, _p(y)), "pointwiseSubtract")
    return y


def interleave(left, right):
    l, r = _f32(left), _f32(right)
    y = np.empty(l.size + r.size, np.float32)
    check(lib.dy4_interleave(_p(l), l.size, _p(r), r.size, _p(y)), "interleave")
    return y
