"""CPU suite, part 3: host-side logic — synthetic stream generator, stream sharding, and the
N>1 launch path (world_size 2, gloo) with the oracle standing in for the per-rank compute."""
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_synth_is_deterministic_and_8bit(dy4):
    a = dy4.synth.make_stream(0, 4096, 65)
    b = dy4.synth.make_stream(0, 4096, 65)
    c = dy4.synth.make_stream(0, 4096, 66)
    assert a.dtype == np.uint8 and a.size == 8192 and np.array_equal(a, b) and not np.array_equal(a, c)
    batch = dy4.synth.make_batch(0, 3, 4096, base_seed=65)
    assert np.array_equal(batch[0], a) and np.array_equal(batch[1], c)
    assert 20 < a.min() and a.max() < 236                 # |I|,|Q| ~ 1 scaled by 100 around 128: never clips


def test_synth_is_a_stereo_fm_signal(dy4, orc):
    """Demodulated through the oracle it has the 19 kHz pilot, a non-zero L-R channel and sane levels.
    (SURVEY.md §8d puts the 38 kHz sub-carrier on sin(2wt); the reference's PLL regenerates ~cos, so the
    recovered L-R is weak — the recipe is kept as specified, it is the arithmetic that is under test.)"""
    m = orc.mode_params(0)
    out = orc.pipeline(0, 1, dy4.synth.make_stream(0, 20 * m.block_size // 2, 65))
    spec = np.abs(np.fft.rfft(out["if"][-65536:] * np.hanning(65536)))
    f = np.fft.rfftfreq(65536, 1 / 240e3)
    assert abs(f[np.argmax(spec * (f > 17e3) * (f < 21e3))] - 19e3) < 20
    L, R = out["audio"][0::2], out["audio"][1::2]
    assert np.std(L - R) > 0.01 and np.abs(out["audio"]).max() < 1.0


@pytest.mark.parametrize("n,w", [(256, 1), (256, 2), (256, 8), (10, 4), (3, 8), (0, 2)])
def test_stream_partition(dy4, n, w):
    r = [dy4.shard.stream_range(n, w, k) for k in range(w)]
    assert r[0][0] == 0 and r[-1][1] == n
    assert all(r[i][1] == r[i + 1][0] for i in range(w - 1))
    sizes = [b - a for a, b in r]
    assert max(sizes) - min(sizes) <= 1


_WORKER = r'''
import os, sys
sys.path.insert(0, %(root)r)
import numpy as np, torch, torch.distributed as dist
import dy4_b200, oracle
rank, local, world = dy4_b200.shard.init_process_group("gloo")
S, nb, mode = 5, 2, 0
m = dy4_b200.mode_params(mode)
lo, hi = dy4_b200.shard.stream_range(S, world, rank)
iq = dy4_b200.synth.make_batch(mode, S, nb * m.block_size // 2, base_seed=65)
o = oracle.load("oracle")                       # stand-in for the per-rank GPU pipeline (CPU test)
rows = torch.from_numpy(np.stack([o.pipeline(mode, 1, iq[s])["pcm"] for s in range(lo, hi)]))
dy4_b200.shard.barrier()
t = dy4_b200.shard.max_over_ranks(10.0 + rank)
tot = dy4_b200.shard.sum_over_ranks(hi - lo)
full = dy4_b200.shard.gather_rows_to_rank0(rows, S)
# the timed gather: every rank writes its slice of ONE shared array (on a GPU box: the target of its device->host copies)
sh = dy4_b200.shard.SharedRows("dy4_test_gather_%%d" %% os.getppid(), S, rows.shape[1], np.int16, rank, world)
sh.mine[:] = rows.numpy()
dy4_b200.shard.barrier()
if rank == 0:
    want = np.stack([o.pipeline(mode, 1, iq[s])["pcm"] for s in range(S)])
    assert t == 10.0 + world - 1 and tot == S, (t, tot)
    assert np.array_equal(full.numpy(), want)
    assert np.array_equal(np.asarray(sh.all), want)
    print("OK")
else:
    assert full is None
path = sh.path
sh.close()
assert not os.path.exists(path)
dist.destroy_process_group()
'''


def test_two_rank_launch_gloo(dy4, orc, tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(_WORKER % {"root": ROOT})
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29533", OMP_NUM_THREADS="1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", "29533", str(script)]
    p = subprocess.run(cmd, capture_output=True, text=True, env=env, timeout=300)
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-2000:]
    assert "OK" in p.stdout


# ---- table-driven PLL (csrc/dy4_plltab.h): host build of the three kernels' logic ------------------------------------
def _plltab_lib(tmp_path_factory=None):
    import ctypes as C
    src = os.path.join(os.path.dirname(os.path.abspath(__file__)), "host", "plltab_host.c")
    out = os.path.join(os.path.dirname(src), "_build")
    os.makedirs(out, exist_ok=True)
    so = os.path.join(out, "libplltab_host.so")
    hdrs = [src] + [os.path.join(os.path.dirname(src), "..", "..", "3dy4-real-time-software-defined-radio-_b200", "csrc", h)
                    for h in ("dy4_plltab.h", "dy4_pllmath.h")]
    if not os.path.exists(so) or any(os.path.getmtime(h) > os.path.getmtime(so) for h in hdrs):
        subprocess.check_call(["gcc", "-O2", "-ffp-contract=off", "-mfma", "-shared", "-fPIC", "-o", so, src, "-lm"])
    return C.CDLL(so)


def _plltab_run(lib, pilot, Fs, launches, which="pllspec_launch"):
    import ctypes as C
    w = 2 * 3.14159265358979323846 * float(np.float32(19e3) / np.float32(Fs))
    Kp = np.float32(0.01) * np.float32(2.666)
    Ki = np.float32(np.float32(0.01) * np.float32(0.01)) * np.float32(3.555)
    th = np.zeros(pilot.size, np.float32)
    st = np.array([1, 0, 0, 0, 0, 0, 0, 0], np.float32)              # PLLState, project.cpp:46-53
    stats = (C.c_long * 6)(0, 0, 0, 0, 0, 0)
    a = 0
    for m in launches:
        args = [C.c_void_p(pilot[a:].ctypes.data), int(m), C.c_void_p(st.ctypes.data), C.c_double(w), C.c_float(Kp), C.c_float(Ki),
                C.c_void_p(th[a:].ctypes.data)]
        if which == "pllspec_launch": args.append(stats)
        getattr(lib, which)(*args)
        a += m
    assert a == pilot.size
    assert stats[2] == 0, "a pick declared certain chose the wrong grid point"
    if which == "pllspec_launch": return th, st, tuple(stats)
    return th, st, (stats[0], stats[1])


# ---- k_pll_sel: chain-ready rows, the pick on the chain, its float certificate per lane afterwards -------------------------
@pytest.mark.parametrize("mode", [0, 1, 2, 3])
@pytest.mark.parametrize("G", [16, 32])
def test_sel_pll_host_build_reproduces_golden_nco(mode, G):
    """The control flow of k_pll_table_ops + k_pll_sel (groups of G steps, pick by threshold, float certificate per step,
    direct evaluation of an uncertain step), built for the host: the reference's trigArg bit for bit, whatever the launch split,
    and the group size.  The in-file reference recurrence is held to the golden NCO row minted from the reference's own fmPLL."""
    lib = _plltab_lib()
    lib.pllspec_set_G(G)
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "mode%d_stereo.npz" % mode))
    Fs = {0: 240e3, 1: 288e3, 2: 240e3, 3: 384e3}[mode]
    pilot = np.ascontiguousarray(g["pilot"]); n = pilot.size
    ref_th, ref_st, _ = _plltab_run(lib, pilot, Fs, [n], "pllref_launch")
    nco = np.empty(n, np.float32); nco[0] = 1.0
    nco[1:] = np.cos((ref_th[:-1] * np.float32(2.0)).astype(np.float64)).astype(np.float32)      # filter.cpp:219-221
    assert np.array_equal(nco.view(np.uint32), g["nco"].view(np.uint32))      # the in-file reference recurrence is the reference's
    for launches in ([n], [n // 2, n - n // 2], [1, 2, 3, 5, n - 11], [37] * (n // 37) + ([n % 37] if n % 37 else [])):
        th, st, stats = _plltab_run(lib, pilot, Fs, launches, "pllspec_launch")
        assert np.array_equal(th.view(np.uint32), ref_th.view(np.uint32)), launches[:4]
        assert np.array_equal(st[:5].view(np.uint32), ref_st[:5].view(np.uint32))


def test_sel_pll_host_build_long_stream_and_direct_rate(dy4, orc):
    """24 blocks in the bench's sub-chunks: bit-identical; after start-up essentially no step needs a direct evaluation (that
    is the speed-up).  Then adversarial inputs."""
    lib = _plltab_lib()
    lib.pllspec_set_G(32)
    nb = 24
    iq = dy4.synth.make_stream(0, nb * 51200, 77)
    pilot = np.ascontiguousarray(orc.pipeline(0, True, iq, want=("pilot",))["pilot"])
    launches = [b * 5120 for b in (1, 2, 4, 8, 8, 1)]
    ref_th, ref_st, _ = _plltab_run(lib, pilot, 240e3, launches, "pllref_launch")
    th, st, stats = _plltab_run(lib, pilot, 240e3, launches, "pllspec_launch")
    fast, direct, wrong, groups = stats[:4]
    assert np.array_equal(th.view(np.uint32), ref_th.view(np.uint32))
    assert np.array_equal(st[:5].view(np.uint32), ref_st[:5].view(np.uint32))
    n = pilot.size
    assert direct <= 1536 + 0.004 * n, (fast, direct)
    print("k_pll_sel flow: %d samples, %d groups of 32, direct %d (%.3f%% beyond the 1536 start-up samples)" % (n, groups, direct, 100.0 * (direct - 1536) / n))
    rng = np.random.default_rng(5)
    bad = pilot[:40960].copy()
    bad[1000:1100] = 0.0; bad[5000:5050] = 1e-42; bad[9000:9400] *= -1.0
    bad[20000:22000] = rng.normal(0, 0.03, 2000).astype(np.float32)
    ref_th, ref_st, _ = _plltab_run(lib, bad, 240e3, [5120] * 8, "pllref_launch")
    th, st, _ = _plltab_run(lib, bad, 240e3, [5120] * 8, "pllspec_launch")
    assert np.array_equal(th.view(np.uint32), ref_th.view(np.uint32))
    assert np.array_equal(st[:5].view(np.uint32), ref_st[:5].view(np.uint32))


def test_spec_pll_certificates_randomised():
    """Both certificates of the serial loop's picks (float: dy4_spec_fast_check, the one k_pll_sel uses; double: dy4_spec_check), probed within 8 float
    ulps of every cell boundary of both candidates: nothing certified is wrong, the float certificate implies the double
    one, and the double one leaves (almost) no gap at the boundaries."""
    import ctypes as C
    lib = _plltab_lib()
    for Fs in (240e3, 288e3, 384e3):
        w = 2 * 3.14159265358979323846 * float(np.float32(19e3) / np.float32(Fs))
        out = (C.c_long * 5)(0, 0, 0, 0, 0)
        lib.pllspec_fuzz(C.c_long(35000), C.c_ulonglong(int(Fs)), C.c_double(w), out)
        probes, fc, dc, wrong, fc_not_dc = list(out)
        assert probes > 2_000_000 and wrong == 0 and fc_not_dc == 0, list(out)
        assert dc > fc > 0.1 * probes, list(out)


def test_rate_change_matches_the_model(dy4, tmp_path):
    """model/fmRateChange.py:43-66 re-rates fixture files with scipy.signal.resample_poly; dy4_b200.rate_change restates that
    algorithm in numpy: same samples (1e-12) and the same output bytes, for every rate pair the reference's table offers from
    2.4 MS/s, and through the command line."""
    from scipy import signal
    rc = dy4.rate_change
    rng = np.random.default_rng(11)
    raw = rng.integers(0, 256, 2 * 6000, dtype=np.uint8)
    iq = (raw - 128.0) / 128.0
    for out_id in range(1, 7):
        fs_in, fs_out = 2400 * 1000, rc.SAMPLE_RATE_TABLE[out_id] * 1000
        g = np.gcd(fs_in, fs_out)
        up, down = fs_out // g, fs_in // g
        want_i = signal.resample_poly(iq[0::2], up, down)
        got_i = rc.resample_poly(iq[0::2], up, down)
        assert got_i.shape == want_i.shape and np.abs(got_i - want_i).max() < 1e-12, out_id
        want = np.empty(2 * want_i.size, np.uint8)
        want_q = signal.resample_poly(iq[1::2], up, down)
        for k in range(want_i.size):                                  # the model's own loop, fmRateChange.py:57-59
            want[2 * k] = np.uint8((128 + int(want_i[k] * 127)) & 0xff)
            want[2 * k + 1] = np.uint8((128 + int(want_q[k] * 127)) & 0xff)
        assert np.array_equal(rc.rate_change(raw, out_id, 0), want), out_id
    f = tmp_path / "cap.raw"
    raw.tofile(f)
    assert rc.main(["rate_change", str(f), "4"]) == 0
    assert np.array_equal(np.fromfile(tmp_path / "cap_1440.raw", dtype=np.uint8), rc.rate_change(raw, 4, 0))
