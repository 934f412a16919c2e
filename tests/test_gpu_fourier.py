"""GPU suite for the Fourier diagnostics (reference src/fourier.cpp, include/fourier.h).
DFT and IDFT must be BIT-IDENTICAL to the reference (same host-built float twiddles, same unfused float sums);
the PSD may differ where CUDA's double log10 and glibc's differ in the last place: one float ulp."""
import numpy as np
import pytest

from conftest import golden, bits

pytestmark = pytest.mark.gpu


def ulps(a, b):
    return np.abs(bits(np.asarray(a, np.float32)).astype(np.int64) - bits(np.asarray(b, np.float32)).astype(np.int64))


def test_dft_idft_golden(dy4):
    g = golden("fourier.npz")
    F = dy4.fourierh
    assert np.array_equal(F.DFT(g["x64"]).view(np.uint32), g["X64"].view(np.uint32))
    assert np.array_equal(F.DFT(g["x512"]).view(np.uint32), g["X512"].view(np.uint32))
    assert np.array_equal(F.IDFT(g["X64"]).view(np.uint32), g["x64_back"].view(np.uint32))


def test_psd_golden(dy4):
    g = golden("fourier.npz")
    freq, psd = dy4.fourierh.estimatePSD(g["sig"], 512, 240000)
    assert np.array_equal(freq, g["freq"])
    assert ulps(psd, g["psd"]).max() <= 1
    assert int(np.argmax(psd)) == int(np.argmax(g["psd"])) == 41          # the 19 kHz tone


@pytest.mark.parametrize("n", [1, 2, 7, 100, 256, 1024])
def test_dft_idft_against_checker(dy4, checker, n):
    rng = np.random.Generator(np.random.PCG64(100 + n))
    x = rng.uniform(-10, 10, n).astype(np.float32)
    X = checker.dft(x)
    assert np.array_equal(dy4.fourierh.DFT(x).view(np.uint32), X.view(np.uint32))
    assert np.array_equal(dy4.fourierh.IDFT(X).view(np.uint32), checker.idft(X).view(np.uint32))


def test_psd_batch_of_receiver_if_rows(dy4, checker):
    """The batched PSD over the IF rows a pipeline call leaves on the device, against estimatePSD per stream."""
    import torch
    m = dy4.mode_params(0)
    S, nb = 6, 2
    iq = dy4.synth.make_batch(0, S, nb * m.block_size // 2, base_seed=31)
    p = dy4.Pipeline(0, 1, S)
    out = p.process(torch.from_numpy(iq).cuda(), want=("if", "pcm"))
    psd = dy4.fourierh.psd_batch(out["if"], 512, 240000)
    torch.cuda.synchronize()
    psd = psd.cpu().numpy()
    h_if = out["if"].cpu().numpy()
    p.close()
    worst = 0
    for s in range(S):
        _, ref = checker.estimate_psd(h_if[s], 512, 240000)
        worst = max(worst, int(ulps(psd[s], ref).max()))
        assert int(np.argmax(psd[s][30:60])) + 30 == 41                    # the 19 kHz pilot stands out of the multiplex
    assert worst <= 1


def test_fourier_argument_checks(dy4):
    with pytest.raises(Exception):
        dy4.fourierh.DFT(np.zeros(4096, np.float32))                        # above the table limit
    with pytest.raises(Exception):
        dy4.fourierh.estimatePSD(np.zeros(100, np.float32), 512, 48000)     # fewer samples than one segment


def test_fft_variants_golden_and_checker(dy4, checker):
    """The reference's three FFTs (fourier.cpp:132-211) and its twiddle table (:125): bit-identical to the fixture minted from the
    reference's own code and to the checker at other lengths; magnitudes equal to the DFT's as test/fft_unittest.cpp:53-91 asks;
    the twiddle table is required where the reference requires it."""
    g = golden("fourier.npz")
    F = dy4.fourierh
    tw = F.compute_twiddles()
    assert np.array_equal(tw.view(np.uint32), g["twiddles"].view(np.uint32))
    c512 = g["x512"].astype(np.complex64)
    got = {"recursive": F.FFT_recursive(c512), "improved": F.FFT_improved(c512, tw, 1), "optimized": F.FFT_optimized(c512, tw)}
    for name, X in got.items():
        assert np.array_equal(X.view(np.uint32), g["F512_" + name].view(np.uint32)), name
        assert np.abs(np.abs(X) - np.abs(F.DFT(g["x512"]))).max() < 2e-2
    assert np.array_equal(F.FFT_recursive(g["x64"].astype(np.complex64)).view(np.uint32), g["F64_recursive"].view(np.uint32))
    rng = np.random.Generator(np.random.PCG64(77))
    for n in (1, 2, 8, 128, 1024, 2048):
        x = (rng.uniform(-10, 10, n) + 1j * rng.uniform(-10, 10, n)).astype(np.complex64)
        assert np.array_equal(F.FFT_recursive(x).view(np.uint32), checker.fft(x, 0).view(np.uint32)), n
    x = (rng.uniform(-10, 10, 512) + 1j * rng.uniform(-10, 10, 512)).astype(np.complex64)
    assert np.array_equal(F.FFT_improved(x, tw).view(np.uint32), checker.fft(x, 1).view(np.uint32))
    assert np.array_equal(F.FFT_optimized(x, tw).view(np.uint32), checker.fft(x, 2).view(np.uint32))
    with pytest.raises(Exception):
        F.FFT_recursive(np.zeros(100, np.complex64))                  # radix 2 only
    with pytest.raises(Exception):
        F.FFT_optimized(np.zeros(1024, np.complex64), tw)             # table built for NFFT = 512
