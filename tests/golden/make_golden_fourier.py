"""Mint the Fourier fixture from the reference's own src/fourier.cpp (compiled in place into oracle/_ref).
Run in the build container:  python tests/golden/make_golden_fourier.py
The reference's own unit tests (test/fft_unittest.cpp, test/idft_unittest.cpp) draw random vectors in [-10, 10] and
compare its FFTs / IDFT with its DFT; they hold no fixed vectors, so this fixture follows their recipe (uniform
[-10, 10], seed below) and records what the reference's DFT, IDFT, estimatePSD, compute_twiddles and three FFTs return."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    ref = oracle.load("ref")
    rng = np.random.Generator(np.random.PCG64(3765))
    x64 = rng.uniform(-10, 10, 64).astype(np.float32)               # fft_unittest.cpp:27-33 style input
    x512 = rng.uniform(-10, 10, 512).astype(np.float32)             # NFFT = 512 (include/dy4.h)
    X64 = ref.dft(x64)
    sig = (0.3 * np.sin(2 * np.pi * 19e3 * np.arange(4096) / 240e3) + rng.normal(0, 0.05, 4096)).astype(np.float32)
    freq, psd = ref.estimate_psd(sig, 512, 240000)
    c512 = x512.astype(np.complex64)                                # fft_unittest.cpp:30-33: the same data as complex
    np.savez_compressed(os.path.join(HERE, "fourier.npz"), x64=x64, X64=X64, x64_back=ref.idft(X64),
                        x512=x512, X512=ref.dft(x512), sig=sig, freq=freq, psd=psd,
                        twiddles=ref.compute_twiddles(256), F512_recursive=ref.fft(c512, 0), F512_improved=ref.fft(c512, 1),
                        F512_optimized=ref.fft(c512, 2), F64_recursive=ref.fft(x64.astype(np.complex64), 0))
    print("round trip error", np.abs(ref.idft(X64).real - x64).max(), "psd peak bin", int(np.argmax(psd)), float(freq[np.argmax(psd)]))


if __name__ == "__main__":
    main()
