"""The reference's include/fourier.h by its own function names, on numpy arrays / torch tensors.

DFT, IDFT and estimatePSD reproduce src/fourier.cpp:14-107 (the naive float DFT with float-narrowed twiddle angles,
not an FFT) on the GPU through libdy4b200.so; psd_batch is the same PSD over [stream][time] device rows.
Nothing here computes on the CPU.
"""
import ctypes as C

import numpy as np

from ._lib import lib, check


def _p(a):
    return C.c_void_p(a.ctypes.data)


def DFT(x):
    """fourier.h:20 — n real samples -> complex64[n]."""
    x = np.ascontiguousarray(x, np.float32)
    out = np.empty(x.size, np.complex64)
    check(lib.dy4_dft(_p(x), x.size, _p(out)), "DFT")
    return out


def IDFT(Xf):
    """fourier.h:38 — complex64[n] -> complex64[n], divided by n."""
    Xf = np.ascontiguousarray(Xf, np.complex64)
    out = np.empty(Xf.size, np.complex64)
    check(lib.dy4_idft(_p(Xf), Xf.size, _p(out)), "IDFT")
    return out


def estimatePSD(samples, nFFT, Fs):
    """fourier.h:31 — returns (freq, psd_est), nFFT/2 floats each."""
    s = np.ascontiguousarray(samples, np.float32)
    freq = np.empty(nFFT // 2, np.float32)
    psd = np.empty(nFFT // 2, np.float32)
    check(lib.dy4_estimate_psd(_p(s), s.size, int(nFFT), int(Fs), _p(freq), _p(psd)), "estimatePSD")
    return freq, psd


NFFT = 512      # include/dy4.h:18


def compute_twiddles(n_twiddles=NFFT // 2, nfft=NFFT):
    """fourier.h:35 / fourier.cpp:125 — twiddles[k] = exp(i * float(-2 PI float(k) / NFFT)), complex64[n_twiddles]."""
    out = np.empty(n_twiddles, np.complex64)
    check(lib.dy4_compute_twiddles(n_twiddles, int(nfft), _p(out)), "compute_twiddles")
    return out


def _fft(x, variant, twiddles):
    x = np.ascontiguousarray(x, np.complex64)
    out = np.empty(x.size, np.complex64)
    tw = None if twiddles is None else np.ascontiguousarray(twiddles, np.complex64)
    check(lib.dy4_fft(_p(x), x.size, variant, None if tw is None else _p(tw), 0 if tw is None else tw.size, _p(out)), "FFT")
    return out


def FFT_recursive(x):
    """fourier.h:37 — radix-2 decimation in time, twiddles computed at every level."""
    return _fft(x, 0, None)


def FFT_improved(x, twiddles, recursion_level=1):
    """fourier.h:39 — the same on a precomputed table (len(x) == NFFT of the table; the reference is called with level 1)."""
    assert recursion_level == 1, "the reference's entry call; deeper levels are its own recursion"
    return _fft(x, 1, twiddles)


def FFT_optimized(x, twiddles):
    """fourier.h:41 — bit reversal, then the butterflies level by level."""
    return _fft(x, 2, twiddles)


def psd_batch(rows, nFFT, Fs, stream=None):
    """estimatePSD of every row of a CUDA float32 tensor [n_streams, n]: returns a CUDA tensor [n_streams, nFFT/2]."""
    import torch
    assert rows.is_cuda and rows.dtype == torch.float32 and rows.dim() == 2 and rows.stride(1) == 1
    out = torch.empty((rows.shape[0], nFFT // 2), dtype=torch.float32, device=rows.device)
    if stream is None:
        stream = torch.cuda.current_stream(rows.device)
    with torch.cuda.device(rows.device):
        check(lib.dy4_psd_batch(C.c_void_p(rows.data_ptr()), rows.stride(0), rows.shape[0], rows.shape[1], int(nFFT), int(Fs),
                                C.c_void_p(out.data_ptr()), out.stride(0), C.c_void_p(stream.cuda_stream)), "psd_batch")
    return out
