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
