"""CPU suite, part 1: the oracle (oracle/dy4_oracle.c) against the golden fixtures minted from
the reference itself (tests/golden/make_golden.py), and against the reference replay when
oracle/_ref is present.  Nothing here touches a GPU."""
import hashlib

import numpy as np
import pytest

from conftest import golden, bits, rel_l2

MODES = [0, 1, 2, 3]


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


@pytest.mark.parametrize("mode", MODES)
def test_oracle_matches_golden_stereo(orc, mode):
    g = golden("mode%d_stereo.npz" % mode)
    out = orc.pipeline(mode, 1, g["iq"])
    for k in ("if", "pilot", "nco", "audio"):
        assert np.array_equal(bits(out[k]), bits(g[k])), k
    assert np.array_equal(out["pcm"], g["pcm"])


@pytest.mark.parametrize("mode", MODES)
def test_oracle_matches_golden_mono(orc, mode):
    iq = golden("mode%d_stereo.npz" % mode)["iq"]
    g = golden("mode%d_mono.npz" % mode)
    out = orc.pipeline(mode, 0, iq)
    assert np.array_equal(bits(out["audio"]), bits(g["audio"]))
    assert np.array_equal(out["pcm"], g["pcm"])


def test_oracle_matches_golden_long_stream(orc, dy4):
    """12 blocks: past ~16.7k IF samples the PLL trajectory is chaotic in the last bit, so this
    only passes if every rounding of the reference is reproduced."""
    g = golden("mode0_stereo_long.npz")
    m = orc.mode_params(0)
    iq = dy4.synth.make_stream(0, int(g["n_blocks"]) * m.block_size // 2, int(g["seed"]))
    assert sha(iq) == str(g["iq_sha256"]), "synthetic generator drifted; regenerate the goldens"
    out = orc.pipeline(0, 1, iq)
    assert sha(out["if"]) == str(g["if_sha256"])
    assert sha(out["pilot"]) == str(g["pilot_sha256"])
    assert sha(out["nco"]) == str(g["nco_sha256"])
    assert np.array_equal(bits(out["nco"][::16]), bits(g["nco_16"]))
    assert np.array_equal(bits(out["audio"]), bits(g["audio"]))
    assert np.array_equal(out["pcm"], g["pcm"])


@pytest.mark.parametrize("mode", MODES)
def test_taps_match_golden(orc, dy4, mode):
    """Both the oracle's and the PRODUCT's host-side tap design (csrc/dy4_taps.cpp) are bit-identical
    to the reference's impulseResponseLPF/BPF output."""
    t = golden("taps.npz")
    m = orc.mode_params(mode)
    fh = dy4.filterh
    cases = [("rf", lambda L: L(m.rf_Fs, 100e3, 101, 1), "lpf"),
             ("audio", lambda L: L(m.if_Fs * m.audio_upsample, 16e3, m.audio_taps, m.audio_upsample), "lpf"),
             ("pilot", lambda L: L(m.if_Fs, 18.5e3, 19.5e3, 101, 1), "bpf"),
             ("stereo", lambda L: L(m.if_Fs, 22e3, 54e3, 101, 1), "bpf")]
    for name, call, kind in cases:
        want = t["%s_%d" % (name, mode)]
        assert np.array_equal(bits(call(orc.lpf_taps if kind == "lpf" else orc.bpf_taps)), bits(want)), name
        assert np.array_equal(bits(call(fh.impulseResponseLPF if kind == "lpf" else fh.impulseResponseBPF)), bits(want)), name
    assert t["audio_%d" % mode].size == 101 * m.audio_upsample


@pytest.mark.parametrize("mode", MODES)
@pytest.mark.parametrize("stereo", [0, 1])
def test_oracle_equals_reference_replay(orc, refl, dy4, mode, stereo):
    m = orc.mode_params(mode)
    iq = dy4.synth.make_stream(mode, 5 * m.block_size // 2, 1234 + mode)
    a, b = orc.pipeline(mode, stereo, iq), refl.pipeline(mode, stereo, iq)
    for k in a:
        if k != "blocks":
            assert np.array_equal(bits(a[k]), bits(b[k])), k


def test_per_op_oracle_equals_reference(orc, refl):
    rng = np.random.default_rng(7)
    x = rng.standard_normal(4000).astype(np.float32)
    h = rng.standard_normal(101).astype(np.float32)
    for name, args in (("block_fir", ()), ("decim_fir", (10,)), ("resample_fir", (147, 800))):
        sa, sb = np.zeros(100, np.float32), np.zeros(100, np.float32)
        hh = rng.standard_normal(101 * 147).astype(np.float32) if name == "resample_fir" else h
        for blk in range(3):
            xb = x[blk * 800:(blk + 1) * 800 + 800]
            ya, yb = getattr(orc, name)(*args, xb, hh, sa), getattr(refl, name)(*args, xb, hh, sb)
            assert np.array_equal(bits(ya), bits(yb)) and np.array_equal(bits(sa), bits(sb)), name
    I, Q = x[:2000].copy(), x[2000:].copy()
    I[5] = Q[5] = 0.0                                      # denominator == 0 branch (filter.cpp:89-92)
    pa, pb = [0.25, -0.5], [0.25, -0.5]
    assert np.array_equal(bits(orc.fm_demod(I, Q, pa)), bits(refl.fm_demod(I, Q, pb))) and pa == pb
    sta = np.array([1, 0, 0, 0, 0, 1], np.float32); stb = sta.copy()
    pil = (0.1 * np.sin(2 * np.pi * 19e3 / 240e3 * np.arange(30000))).astype(np.float32)
    pil[100] = 0.0                                         # the (in==0 ? 1 : in) guard (filter.cpp:192)
    na = np.concatenate([orc.pll(pil[i:i + 10000], 19e3, 240e3, 2.0, 0.0, 0.01, sta) for i in (0, 10000, 20000)])
    nb = np.concatenate([refl.pll(pil[i:i + 10000], 19e3, 240e3, 2.0, 0.0, 0.01, stb) for i in (0, 10000, 20000)])
    assert np.array_equal(bits(na), bits(nb)) and np.array_equal(bits(sta), bits(stb))
    std = np.arange(50, dtype=np.float32); std2 = std.copy()
    assert np.array_equal(orc.delay_block(x[:500], std), refl.delay_block(x[:500], std2)) and np.array_equal(std, std2)
    raw = np.arange(256, dtype=np.uint8)
    assert np.array_equal(bits(orc.iq_to_float(raw)), bits(refl.iq_to_float(raw)))


def test_iq_to_float_is_k_over_128(orc):
    raw = np.arange(256, dtype=np.uint8)
    assert np.array_equal(orc.iq_to_float(raw), ((raw.astype(np.int32) - 128) / 128.0).astype(np.float32))


def test_pcm_conversion(orc):
    x = np.array([0.0, 0.99999, -0.99999, 1.5 / 16384, -1.5 / 16384, np.nan, 0.5, -0.5], np.float32)
    assert orc.pcm16(x).tolist() == [0, 16383, -16383, 1, -1, 0, 8192, -8192]      # truncation toward zero, NaN -> 0


def test_partial_trailing_block_is_dropped(orc, dy4):
    m = orc.mode_params(0)
    iq = dy4.synth.make_stream(0, 2 * m.block_size // 2 + 777, 3)
    full = orc.pipeline(0, 0, iq[:2 * m.block_size])
    part = orc.pipeline(0, 0, iq)
    assert part["blocks"] == 2 and np.array_equal(part["pcm"], full["pcm"])       # project.cpp:293-296
    assert orc.pipeline(0, 0, iq[:100])["blocks"] == 0


def test_silence_gives_zero_if(orc):
    m = orc.mode_params(1)
    iq = np.full(2 * m.block_size, 128, np.uint8)
    out = orc.pipeline(1, 1, iq)
    assert not out["if"].any() and not out["pcm"].any()


def test_block_split_invariance(orc):
    """FIR / demod / delay state is input history, so any block split gives the same samples."""
    rng = np.random.default_rng(11)
    x = rng.standard_normal(6000).astype(np.float32)
    h = rng.standard_normal(101).astype(np.float32)
    s1 = np.zeros(100, np.float32)
    whole = orc.decim_fir(10, x, h, s1)
    s2 = np.zeros(100, np.float32)
    parts = np.concatenate([orc.decim_fir(10, x[a:b], h, s2) for a, b in ((0, 1000), (1000, 1500), (1500, 6000))])
    assert np.array_equal(bits(whole), bits(parts))


def test_decimating_fir_is_fir_then_keep_every_dth(orc):
    rng = np.random.default_rng(12)
    x = rng.standard_normal(5000).astype(np.float32)
    h = rng.standard_normal(101).astype(np.float32)
    a = orc.decim_fir(5, x, h, np.zeros(100, np.float32))
    b = orc.block_fir(x, h, np.zeros(100, np.float32))[::5]
    assert np.array_equal(bits(a), bits(b))
    c = orc.resample_fir(1, 5, x, h, np.zeros(100, np.float32))
    assert np.array_equal(bits(a), bits(c))


def test_pll_locks_to_pilot(orc):
    n = 48000
    ph = 2 * np.pi * 19e3 / 240e3 * np.arange(n) + 0.7
    st = np.array([1, 0, 0, 0, 0, 1], np.float32)
    nco = orc.pll(np.sin(ph).astype(np.float32) * 0.1, 19e3, 240e3, 2.0, 0.0, 0.01, st)
    # locked: the NCO is a unit cosine at 38 kHz, phase-coherent with the doubled pilot
    tail = slice(n - 4000, n)
    coh = np.abs(np.mean(nco[tail] * np.exp(-2j * ph[tail])))
    assert coh > 0.49, coh                                   # 0.5 = perfectly coherent unit cosine
    assert np.abs(np.mean(nco[tail] * np.exp(-2j * 1.01 * ph[tail]))) < 0.1


# ---------------------------------------------------------------- RDS path (reference: the Python model only)
def _rds_golden_if(orc, dy4):
    g = golden("rds_mode0.npz")
    m = dy4.mode_params(0)
    iq = dy4.synth.make_stream(0, int(g["n_blocks"]) * m.block_size // 2, int(g["seed"]), rds=True)
    assert sha(iq) == str(g["iq_sha256"])
    return g, orc.pipeline(0, 1, iq, want=("if",))["if"]


def test_rds_oracle_matches_model_golden(orc, dy4):
    """oracle/rds.py (numpy restatement) against outputs of the model's own functions (make_golden_rds.py)."""
    from oracle import rds
    g, fm = _rds_golden_if(orc, dy4)
    r = rds.rds_front(fm)
    n = len(g["rrc_i64"])
    assert rel_l2(r["rrc_i"][:n], g["rrc_i64"]) <= 1e-9 and rel_l2(r["rrc_q"][:n], g["rrc_q64"]) <= 1e-9
    assert rel_l2(r["rrc_i"], g["rrc_i"]) <= 2e-7 and rel_l2(r["rrc_q"], g["rrc_q"]) <= 2e-7      # fixture stored as float32
    for k in ("rds_f", "carrier", "nco_i", "nco_q"):
        assert rel_l2(r[k][::8], g[k + "_8"]) <= 2e-7
    be = rds.rds_back(r["rrc_i"], r["rrc_q"])
    assert [len(x) for x in be.symbols] == list(g["symbol_counts"])
    assert np.array_equal(np.concatenate([np.array(x, np.int8) for x in be.symbols]), g["symbols"])
    assert np.array_equal(np.array(be.bits, np.int8), g["bits"])
    assert np.array_equal(np.array(be.events, np.int32).reshape(-1, 4), g["events"])
    assert [be.errors1, be.errors2] == list(g["errors"])
    ev = g["events"]
    assert len(ev) >= 16 and set(ev[:, 0]) >= {0, 1, 2, 4} and not ev[:, 2].any()      # the fixture does exercise the frame sync
    # the glue that collects A,B,C,D words and the application layer (the model's own process_rds_data in the fixture)
    assert np.array_equal(np.array(be.groups, np.int32).reshape(-1, 4), g["groups"]) and len(g["groups"]) >= 10
    assert be.app_lines == str(g["app_lines"]).split("\n")
    assert any(l.startswith("PI code: ") for l in be.app_lines) and any(l.startswith("Program type: ") for l in be.app_lines)


def test_rds_application_layer_of_the_package_matches_model(dy4):
    """dy4_b200.rds_app.ApplicationLayer (product side, host strings) fed with the fixture's groups prints what the
    model's process_rds_data printed."""
    g = golden("rds_mode0.npz")
    app = dy4.rds_app.ApplicationLayer()
    assert app.feed_groups(g["groups"]) == str(g["app_lines"]).split("\n")


def test_rds_taps_match_scipy_and_the_library(dy4):
    """firwin / RRC designs: oracle restatement == scipy (what the model calls) == the library's C implementation."""
    import ctypes as C
    from scipy import signal
    from oracle import rds
    t = rds.model_taps()
    nyq = 120e3
    ref = dict(rds=signal.firwin(101, [54e3 / nyq, 60e3 / nyq], window="hann", pass_zero=False),
               carrier=signal.firwin(101, [113.5e3 / nyq, 114.5e3 / nyq], window="hann", pass_zero=False),
               lpf=signal.firwin(1919, 3e3 / (240e3 * 19 / 2)) * 19)
    for k, v in ref.items():
        assert np.abs(t[k] - v).max() <= 1e-15
    lib = dy4._lib.lib
    h = np.empty(101)
    assert lib.dy4_firwin(101, 54e3 / nyq, 60e3 / nyq, 0, C.c_void_p(h.ctypes.data)) == 0
    assert np.abs(h - ref["rds"]).max() <= 1e-15
    h = np.empty(1919)
    assert lib.dy4_firwin(1919, 0.0, 3e3 / (240e3 * 19 / 2), 1, C.c_void_p(h.ctypes.data)) == 0
    assert np.abs(h * 19 - ref["lpf"]).max() <= 1e-15
    h = np.empty(101)
    assert lib.dy4_rrc_taps(38000.0, 101, C.c_void_p(h.ctypes.data)) == 0
    assert np.abs(h - t["rrc"]).max() <= 1e-15
    assert lib.dy4_firwin(101, 0.5, 0.4, 0, C.c_void_p(h.ctypes.data)) != 0          # band edges out of order


def test_synthetic_rds_groups_are_valid(dy4):
    """synth's RDS groups carry correct checkwords: every block's syndrome is its offset word's (model's matrix)."""
    from oracle import rds
    bits = dy4.synth.rds_group_bits(104 * 5, 1234)
    for g in range(5):
        for j, off in enumerate(("A", "B", "C", "D")):
            blk = [int(b) for b in bits[104 * g + 26 * j:104 * g + 26 * (j + 1)]]
            syn = tuple(sum(blk[i] for i in row) % 2 for row in rds.PARITY_ROWS)
            assert rds.SYNDROMES.get(syn) == off
    d = dy4.synth.rds_bitstream(300, 1234)
    assert np.array_equal(np.diff(d) % 2, bits[1:300])                                 # differential encoding


# ---------------------------------------------------------------- Fourier diagnostics (src/fourier.cpp)
def test_fourier_oracle_matches_golden(orc):
    """The C restatement of DFT / IDFT / estimatePSD against what the reference's own fourier.cpp returned."""
    g = golden("fourier.npz")
    assert np.array_equal(orc.dft(g["x64"]).view(np.uint32), g["X64"].view(np.uint32))
    assert np.array_equal(orc.dft(g["x512"]).view(np.uint32), g["X512"].view(np.uint32))
    assert np.array_equal(orc.idft(g["X64"]).view(np.uint32), g["x64_back"].view(np.uint32))
    freq, psd = orc.estimate_psd(g["sig"], 512, 240000)
    assert np.array_equal(freq, g["freq"]) and np.array_equal(bits(psd), bits(g["psd"]))


def test_fourier_oracle_equals_reference_live(orc, refl):
    rng = np.random.Generator(np.random.PCG64(11))
    for n in (1, 2, 7, 16, 100, 256):
        x = rng.uniform(-10, 10, n).astype(np.float32)
        X = refl.dft(x)
        assert np.array_equal(orc.dft(x).view(np.uint32), X.view(np.uint32))
        assert np.array_equal(orc.idft(X).view(np.uint32), refl.idft(X).view(np.uint32))
    s = rng.normal(0, 1, 3000).astype(np.float32)
    for nfft in (64, 256):
        a, b = orc.estimate_psd(s, nfft, 48000), refl.estimate_psd(s, nfft, 48000)
        assert np.array_equal(a[0], b[0]) and np.array_equal(bits(a[1]), bits(b[1]))


def test_fourier_properties(orc):
    """What the reference's own unit tests check (test/idft_unittest.cpp:45-60, fft_unittest.cpp:40-60), with a
    meaningful tolerance: IDFT(DFT(x)) returns x, and the DFT agrees with numpy's FFT to the accuracy its float
    twiddle angles allow."""
    rng = np.random.Generator(np.random.PCG64(5))
    x = rng.uniform(-10, 10, 128).astype(np.float32)
    X = orc.dft(x)
    back = orc.idft(X)
    assert np.abs(back.real - x).max() < 1e-3 and np.abs(back.imag).max() < 1e-3
    assert rel_l2(np.stack([X.real, X.imag]), np.stack([np.fft.fft(x).real, np.fft.fft(x).imag])) < 1e-4


def test_fft_oracle_matches_golden_and_reference(orc):
    """fourier.cpp:125-211: the restatement of compute_twiddles and of the three radix-2 FFTs reproduces the fixture minted from the
    reference's own code, bit for bit; as test/fft_unittest.cpp:53-91 checks, their magnitudes agree with the DFT's; and — where
    oracle/_ref is built — the restatement equals the reference live at other lengths."""
    import oracle
    g = golden("fourier.npz")
    c512 = g["x512"].astype(np.complex64)
    assert np.array_equal(orc.compute_twiddles(256).view(np.uint32), g["twiddles"].view(np.uint32))
    for v, name in enumerate(("recursive", "improved", "optimized")):
        assert np.array_equal(orc.fft(c512, v).view(np.uint32), g["F512_" + name].view(np.uint32)), name
        assert np.abs(np.abs(g["F512_" + name]) - np.abs(g["X512"])).max() < 2e-2          # |FFT| vs |DFT|: the unit test's comparison
    assert np.array_equal(orc.fft(g["x64"].astype(np.complex64), 0).view(np.uint32), g["F64_recursive"].view(np.uint32))
    if oracle.have_ref():
        ref = oracle.load("ref")
        rng = np.random.Generator(np.random.PCG64(9))
        for n in (1, 2, 8, 128, 512, 2048):
            x = (rng.uniform(-10, 10, n) + 1j * rng.uniform(-10, 10, n)).astype(np.complex64)
            assert np.array_equal(orc.fft(x, 0).view(np.uint32), ref.fft(x, 0).view(np.uint32)), n
        x = (rng.uniform(-10, 10, 512) + 1j * rng.uniform(-10, 10, 512)).astype(np.complex64)
        for v in (1, 2):
            assert np.array_equal(orc.fft(x, v).view(np.uint32), ref.fft(x, v).view(np.uint32)), v
