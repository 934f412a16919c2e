"""ctypes front end for the two CPU checkers (see oracle/dy4_oracle.h).

``load("oracle")`` -> the plain-C restatement (oracle/dy4_oracle.c, prefix dy4o_)
``load("ref")``    -> the reference's own compiled code (oracle/_ref, prefix dy4r_)

TEST INFRASTRUCTURE ONLY — never imported by the product package.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_PATHS = {
    "oracle": (os.path.join(HERE, "_build", "libdy4oracle.so"), "dy4o_"),
    "ref": (os.path.join(HERE, "_ref", "libdy4ref.so"), "dy4r_"),
}


class ModeParams(C.Structure):
    _fields_ = [("rf_Fs", C.c_float), ("rf_decim", C.c_int), ("if_Fs", C.c_float),
                ("audio_decim", C.c_int), ("audio_upsample", C.c_int), ("audio_taps", C.c_int),
                ("block_size", C.c_int), ("if_per_block", C.c_int), ("audio_per_block", C.c_int)]


def build(ref=None):
    """Compile the checkers.  The restatement always; the reference replay only
    where /root/reference exists (the build container)."""
    subprocess.check_call(["make", "-s", "-C", HERE, "oracle"])
    if ref is None:
        ref = os.path.isdir("/root/reference/src")
    if ref:
        subprocess.check_call(["make", "-s", "-C", HERE, "ref"])


def have_ref():
    return os.path.exists(_PATHS["ref"][0])


def _fp(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


class CpuReceiver:
    """One of the two CPU libraries, numpy in / numpy out."""

    def __init__(self, kind):
        path, prefix = _PATHS[kind]
        if not os.path.exists(path):
            if kind == "oracle":
                build(ref=False)
            else:
                raise FileNotFoundError(path + " (run `make -C oracle ref` where /root/reference exists)")
        self.kind = kind
        self.lib = C.CDLL(path)
        self.p = prefix
        f = self._f
        f("mode_params").restype = C.c_int
        f("resample_fir").restype = C.c_int
        f("pipeline").restype = C.c_long
        f("pipeline").argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_long] + [C.c_void_p] * 5
        f("lpf_taps").argtypes = [C.c_float, C.c_float, C.c_ushort, C.c_int, C.c_void_p]
        f("bpf_taps").argtypes = [C.c_float, C.c_float, C.c_float, C.c_ushort, C.c_int, C.c_void_p]
        f("pll").argtypes = [C.c_void_p, C.c_int, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float,
                             C.c_void_p, C.c_void_p]

    def _f(self, name):
        return getattr(self.lib, self.p + name)

    # ---- mode table -------------------------------------------------------
    def mode_params(self, mode):
        m = ModeParams()
        if self._f("mode_params")(int(mode), C.byref(m)) != 0:
            raise ValueError("mode must be 0..3")
        return m

    # ---- tap design -------------------------------------------------------
    def lpf_taps(self, Fs, Fc, num_taps, up=1):
        h = np.empty(num_taps, np.float32)
        self._f("lpf_taps")(Fs, Fc, num_taps, up, h.ctypes.data)
        return h

    def bpf_taps(self, Fs, Fb, Fe, num_taps, up=1):
        h = np.empty(num_taps, np.float32)
        self._f("bpf_taps")(Fs, Fb, Fe, num_taps, up, h.ctypes.data)
        return h

    # ---- per-op -----------------------------------------------------------
    def iq_to_float(self, raw):
        raw = np.ascontiguousarray(raw, np.uint8)
        out = np.empty(raw.size, np.float32)
        self._f("iq_to_float")(C.c_void_p(raw.ctypes.data), C.c_long(raw.size), C.c_void_p(out.ctypes.data))
        return out

    def block_fir(self, x, h, state):
        x = np.ascontiguousarray(x, np.float32); h = np.ascontiguousarray(h, np.float32)
        y = np.empty(x.size, np.float32)
        self._f("block_fir")(_fp(x), x.size, _fp(h), h.size, _fp(state), state.size, _fp(y))
        return y

    def decim_fir(self, factor, x, h, state):
        x = np.ascontiguousarray(x, np.float32); h = np.ascontiguousarray(h, np.float32)
        y = np.empty(x.size // factor, np.float32)
        self._f("decim_fir")(factor, _fp(x), x.size, _fp(h), h.size, _fp(state), state.size, _fp(y))
        return y

    def resample_fir(self, up, down, x, h, state):
        x = np.ascontiguousarray(x, np.float32); h = np.ascontiguousarray(h, np.float32)
        y = np.empty(int((x.size / np.float32(down)) * up) + 1, np.float32)
        n = self._f("resample_fir")(up, down, _fp(x), x.size, _fp(h), h.size, _fp(state), state.size, _fp(y))
        return y[:n].copy()

    def fm_demod(self, I, Q, prev):
        I = np.ascontiguousarray(I, np.float32); Q = np.ascontiguousarray(Q, np.float32)
        out = np.empty(I.size, np.float32)
        pi, pq = C.c_float(prev[0]), C.c_float(prev[1])
        self._f("fm_demod")(_fp(I), _fp(Q), I.size, C.byref(pi), C.byref(pq), _fp(out))
        prev[0], prev[1] = pi.value, pq.value
        return out

    def pll(self, x, freq, Fs, ncoScale, phaseAdjust, normBandwidth, state):
        """state: float32[6] = fbI, fbQ, integrator, phaseEst, trigOffset, nco_state (updated in place)."""
        x = np.ascontiguousarray(x, np.float32)
        nco = np.empty(x.size, np.float32)
        self._f("pll")(x.ctypes.data, x.size, freq, Fs, ncoScale, phaseAdjust, normBandwidth,
                       nco.ctypes.data, state.ctypes.data)
        return nco

    def delay_block(self, x, state):
        x = np.ascontiguousarray(x, np.float32)
        out = np.empty(x.size, np.float32)
        self._f("delay_block")(_fp(x), x.size, _fp(state), state.size, _fp(out))
        return out

    def pcm16(self, x):
        x = np.ascontiguousarray(x, np.float32)
        out = np.empty(x.size, np.int16)
        self._f("pcm16")(_fp(x), C.c_long(x.size), C.c_void_p(out.ctypes.data))
        return out

    # ---- Fourier diagnostics (src/fourier.cpp) --------------------------------
    def dft(self, x):
        x = np.ascontiguousarray(x, np.float32)
        out = np.empty(x.size, np.complex64)
        self._f("dft")(C.c_void_p(x.ctypes.data), C.c_int(x.size), C.c_void_p(out.ctypes.data))
        return out

    def idft(self, Xf):
        Xf = np.ascontiguousarray(Xf, np.complex64)
        out = np.empty(Xf.size, np.complex64)
        self._f("idft")(C.c_void_p(Xf.ctypes.data), C.c_int(Xf.size), C.c_void_p(out.ctypes.data))
        return out

    def fft(self, x, variant):
        x = np.ascontiguousarray(x, np.complex64)
        out = np.empty(x.size, np.complex64)
        self._f("fft")(C.c_void_p(x.ctypes.data), C.c_int(x.size), C.c_int(variant), C.c_void_p(out.ctypes.data))
        return out

    def compute_twiddles(self, n_tw=256):
        out = np.empty(n_tw, np.complex64)
        self._f("compute_twiddles")(C.c_int(n_tw), C.c_void_p(out.ctypes.data))
        return out

    def estimate_psd(self, samples, nfft, Fs):
        s = np.ascontiguousarray(samples, np.float32)
        freq, psd = np.empty(nfft // 2, np.float32), np.empty(nfft // 2, np.float32)
        self._f("estimate_psd")(C.c_void_p(s.ctypes.data), C.c_long(s.size), C.c_int(nfft), C.c_int(Fs),
                                C.c_void_p(freq.ctypes.data), C.c_void_p(psd.ctypes.data))
        return freq, psd

    # ---- whole receiver over one stream ------------------------------------
    def pipeline(self, mode, stereo, iq, want=("if", "audio", "pcm", "pilot", "nco")):
        iq = np.ascontiguousarray(iq, np.uint8)
        m = self.mode_params(mode)
        nb = iq.size // m.block_size
        nch = 2 if stereo else 1
        out = {}
        if "if" in want: out["if"] = np.zeros(nb * m.if_per_block, np.float32)
        if "audio" in want: out["audio"] = np.zeros(nb * m.audio_per_block * nch, np.float32)
        if "pcm" in want: out["pcm"] = np.zeros(nb * m.audio_per_block * nch, np.int16)
        if stereo and "pilot" in want: out["pilot"] = np.zeros(nb * m.if_per_block, np.float32)
        if stereo and "nco" in want: out["nco"] = np.zeros(nb * m.if_per_block, np.float32)
        ptr = lambda k: C.c_void_p(out[k].ctypes.data) if k in out else None
        n = self._f("pipeline")(mode, int(bool(stereo)), C.c_void_p(iq.ctypes.data), C.c_long(iq.size),
                                ptr("if"), ptr("audio"), ptr("pcm"), ptr("pilot"), ptr("nco"))
        assert n == nb
        out["blocks"] = nb
        return out


_cache = {}


def load(kind="oracle"):
    if kind not in _cache:
        _cache[kind] = CpuReceiver(kind)
    return _cache[kind]
