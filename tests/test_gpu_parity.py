"""GPU suite (-m gpu): the CUDA path, called through the C ABI, against the golden fixtures minted
from the reference and against the CPU checker (the reference replay oracle/_ref when it was built,
else the C restatement) on the same seeded inputs.

Tolerances (BASELINE.json north_star): IF and float audio relative L2 <= 1e-5, int16 PCM within
1 LSB.  What is actually asserted is stronger wherever the design makes it so: IF, pilot and NCO
are BIT-EXACT always (they feed the PLL, which is chaotic in the last bit), and with
exact_audio=True audio and PCM are bit-exact too."""
import hashlib

import numpy as np
import pytest

from conftest import golden, bits, rel_l2

pytestmark = pytest.mark.gpu
MODES = [0, 1, 2, 3]
TOL_L2 = 1e-5


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def run_gpu(dy4, mode, stereo, iq, exact_audio=False, want=("pcm", "audio", "if"), host=False, **kw):
    import torch
    iq = np.atleast_2d(iq)
    p = dy4.Pipeline(mode, stereo, iq.shape[0], exact_audio=exact_audio, debug_rows=not host)
    try:
        if host:
            out = p.process_host(iq, want=[w for w in want if w != "if"], **kw)
            return {k: np.asarray(v) for k, v in out.items()}, None
        out = p.process(torch.from_numpy(iq).cuda(), want=want)
        torch.cuda.synchronize()
        dbg = [t.cpu().numpy() for t in p.debug_pilot_nco()] if stereo else None
        return {k: v.cpu().numpy() for k, v in out.items()}, dbg
    finally:
        p.close()


# ---------------------------------------------------------------- golden fixtures (the reference's own outputs)
@pytest.mark.parametrize("mode", MODES)
def test_golden_stereo(dy4, mode):
    g = golden("mode%d_stereo.npz" % mode)
    out, (pilot, nco) = run_gpu(dy4, mode, 1, g["iq"], exact_audio=True)
    assert np.array_equal(bits(out["if"][0]), bits(g["if"]))
    assert np.array_equal(bits(pilot[0]), bits(g["pilot"]))
    assert np.array_equal(bits(nco[0]), bits(g["nco"]))
    assert np.array_equal(bits(out["audio"][0]), bits(g["audio"]))
    assert np.array_equal(out["pcm"][0], g["pcm"])
    out, _ = run_gpu(dy4, mode, 1, g["iq"], exact_audio=False)            # default (fused where the PLL cannot see it)
    assert np.array_equal(bits(out["if"][0]), bits(g["if"]))
    assert rel_l2(out["audio"][0], g["audio"]) <= TOL_L2
    assert np.abs(out["pcm"][0].astype(np.int32) - g["pcm"]).max() <= 1


@pytest.mark.parametrize("mode", MODES)
def test_golden_mono(dy4, mode):
    iq = golden("mode%d_stereo.npz" % mode)["iq"]
    g = golden("mode%d_mono.npz" % mode)
    out, _ = run_gpu(dy4, mode, 0, iq, exact_audio=True)
    assert np.array_equal(bits(out["audio"][0]), bits(g["audio"])) and np.array_equal(out["pcm"][0], g["pcm"])
    out, _ = run_gpu(dy4, mode, 0, iq)
    assert rel_l2(out["audio"][0], g["audio"]) <= TOL_L2
    assert np.abs(out["pcm"][0].astype(np.int32) - g["pcm"]).max() <= 1


def test_golden_long_stream_pll_chaos(dy4):
    g = golden("mode0_stereo_long.npz")
    m = dy4.mode_params(0)
    iq = dy4.synth.make_stream(0, int(g["n_blocks"]) * m.block_size // 2, int(g["seed"]))
    assert sha(iq) == str(g["iq_sha256"])
    out, (pilot, nco) = run_gpu(dy4, 0, 1, iq, exact_audio=True)
    assert sha(out["if"][0]) == str(g["if_sha256"])
    assert sha(pilot[0]) == str(g["pilot_sha256"])
    assert sha(nco[0]) == str(g["nco_sha256"])
    assert np.array_equal(bits(out["audio"][0]), bits(g["audio"])) and np.array_equal(out["pcm"][0], g["pcm"])


# ---------------------------------------------------------------- live checker on seeded batches
@pytest.mark.parametrize("mode", MODES)
@pytest.mark.parametrize("stereo", [0, 1])
def test_batch_against_checker(dy4, checker, mode, stereo):
    m = dy4.mode_params(mode)
    S, nb = 5, 7
    iq = dy4.synth.make_batch(mode, S, nb * m.block_size // 2, base_seed=200 + 10 * mode)
    out, dbg = run_gpu(dy4, mode, stereo, iq)
    oute, _ = run_gpu(dy4, mode, stereo, iq, exact_audio=True)
    for s in range(S):
        ref = checker.pipeline(mode, stereo, iq[s])
        if stereo:
            assert np.array_equal(bits(out["if"][s]), bits(ref["if"]))          # feeds the PLL: always bit-exact
        else:
            assert rel_l2(out["if"][s], ref["if"]) <= TOL_L2                   # mono default: fused front end
        assert np.array_equal(bits(oute["if"][s]), bits(ref["if"]))
        assert rel_l2(out["audio"][s], ref["audio"]) <= TOL_L2
        assert np.abs(out["pcm"][s].astype(np.int32) - ref["pcm"]).max() <= 1
        assert np.array_equal(bits(oute["audio"][s]), bits(ref["audio"])) and np.array_equal(oute["pcm"][s], ref["pcm"])
        if stereo:
            assert np.array_equal(bits(dbg[0][s]), bits(ref["pilot"])) and np.array_equal(bits(dbg[1][s]), bits(ref["nco"]))


def test_two_second_stream(dy4, checker):
    """~2 s of mode 0 stereo (94 blocks, 481k IF samples): deep inside the regime where one differing
    rounding anywhere upstream of the PLL would change the audio at the 1e-3 level."""
    m = dy4.mode_params(0)
    iq = dy4.synth.make_batch(0, 2, 94 * m.block_size // 2, base_seed=500)
    out, (pilot, nco) = run_gpu(dy4, 0, 1, iq)
    for s in range(2):
        ref = checker.pipeline(0, 1, iq[s])
        assert np.array_equal(bits(nco[s]), bits(ref["nco"]))
        assert rel_l2(out["audio"][s], ref["audio"]) <= TOL_L2
        assert np.abs(out["pcm"][s].astype(np.int32) - ref["pcm"]).max() <= 1


# ---------------------------------------------------------------- chunking, host path, state
def test_chunking_and_subchunking_do_not_change_results(dy4, monkeypatch):
    import torch
    m = dy4.mode_params(2)
    iq = dy4.synth.make_batch(2, 3, 9 * m.block_size // 2, base_seed=31)
    one, _ = run_gpu(dy4, 2, 1, iq, exact_audio=True)
    d = torch.from_numpy(iq).cuda()
    p = dy4.Pipeline(2, 1, 3, exact_audio=True)
    parts = [p.process(d[:, a * m.block_size:b * m.block_size], want=("pcm", "audio", "if")) for a, b in ((0, 1), (1, 2), (2, 6), (6, 9))]
    for k in one:
        assert np.array_equal(torch.cat([x[k] for x in parts], 1).cpu().numpy(), one[k]), k
    p.close()
    monkeypatch.setenv("DY4_SUBCHUNK_BLOCKS", "2")           # force the internal time sub-chunking
    sub, _ = run_gpu(dy4, 2, 1, iq, exact_audio=True)
    for k in one:
        assert np.array_equal(sub[k], one[k]), k


def test_call_windows_do_not_change_results(dy4, monkeypatch):
    """A call whose IF / pilot / stereo-band / NCO rows exceed the row budget goes through in several windows (dy4_pipeline_process):
    PCM, audio, IF rows and the RDS baseband of the whole call are those of the single-window call."""
    import torch
    m = dy4.mode_params(0)
    S, nb = 3, 20
    d = dy4.synth.make_batch_torch(0, S, nb * m.block_size // 2, base_seed=91, device="cuda", rds=True)

    def run():
        p = dy4.Pipeline(0, 1, S, rds=True)
        out = p.process(d, want=("pcm", "audio", "if"))
        ri, rq = p.rds_read()
        torch.cuda.synchronize()
        dr = p.rds_drain()
        p.close()
        return out, ri.cpu().numpy(), rq.cpu().numpy(), dr

    one, ri1, rq1, dr1 = run()
    monkeypatch.setenv("DY4_WS_BYTES", str(S * m.if_per_block * 16 * 7 // 4))     # rows for 7 blocks: windows of 7, 7, 6
    win, ri2, rq2, dr2 = run()
    for k in one:
        assert torch.equal(one[k], win[k]), k
    assert ri1.shape == ri2.shape and np.array_equal(bits(ri1), bits(ri2)) and np.array_equal(bits(rq1), bits(rq2))
    for s in range(S):
        for k in ("symbols", "bits", "events", "groups"):
            assert np.array_equal(dr1[s][k], dr2[s][k]), (s, k)


@pytest.mark.parametrize("rds", [False, True])
def test_pipelined_calls_do_not_change_results(dy4, rds):
    """DY4_FLAG_PIPELINED: consecutive calls overlap on the device (two sets of call rows, IF history moved on as soon as a call's
    FIR work is queued, the prediction carried across calls, loops on SMs of their own).  Six calls of uneven length queued
    back to back, outputs read after one flush: the bits of the call-by-call pipeline; state and RDS output as well."""
    import torch
    m = dy4.mode_params(0)
    S = 24
    cuts = [0, 9, 10, 22, 30, 31, 44]
    d = dy4.synth.make_batch_torch(0, S, cuts[-1] * m.block_size // 2, base_seed=411, device="cuda", rds=rds)

    def run(pipelined):
        p = dy4.Pipeline(0, 1, S, rds=rds, pipelined=pipelined)
        outs = [p.process(d[:, a * m.block_size:b * m.block_size], want=("pcm", "audio", "if")) for a, b in zip(cuts, cuts[1:])]
        p.flush()
        torch.cuda.synchronize()
        part = p.sm_partition()
        state = p.get_state().copy()
        dr = p.rds_drain() if rds else None
        p.close()
        return outs, state, dr, part

    one, st1, dr1, part1 = run(False)
    two, st2, dr2, part2 = run(True)
    assert part1 == (0, 0) and part2[0] >= 8, (part1, part2)
    for x, y in zip(one, two):
        for k in x:
            assert torch.equal(x[k], y[k]), k
    assert np.array_equal(st1, st2)
    if rds:
        for s in range(S):
            for k in ("symbols", "bits", "events", "groups"):
                assert np.array_equal(dr1[s][k], dr2[s][k]), (s, k)


@pytest.mark.parametrize("case", ["jumps", "no_partition", "direct_loop", "windows"])
def test_pipelined_calls_edge_cases(dy4, monkeypatch, case):
    """Overlapped calls where the carried prediction is of no use or the schedule differs: unrelated signals from call to call (the
    loop falls back to direct steps and the turn report re-centres the prediction), more than 256 streams (no SM partition), the
    direct PLL loop (no table), and calls cut into windows.  Always the bits of joined calls."""
    import torch
    m = dy4.mode_params(0)
    S = 300 if case == "no_partition" else 20
    nb = 9
    if case == "direct_loop":
        monkeypatch.setenv("DY4_PLL_TABLE_MAX", "0")
    if case == "windows":
        monkeypatch.setenv("DY4_WS_BYTES", str(S * m.if_per_block * 32 * 4 // 4))     # rows for 4 blocks (two sets): windows of 4, 4, 1
    chunks = [dy4.synth.make_batch_torch(0, S, nb * m.block_size // 2, base_seed=500 + (97 * k if case == "jumps" else 0), device="cuda")
              for k in range(4)]
    if case != "jumps":                                            # one continuous signal
        d = dy4.synth.make_batch_torch(0, S, 4 * nb * m.block_size // 2, base_seed=500, device="cuda")
        chunks = [d[:, k * nb * m.block_size:(k + 1) * nb * m.block_size] for k in range(4)]

    def run(pipelined):
        p = dy4.Pipeline(0, 1, S, pipelined=pipelined)
        outs = [p.process(c, want=("pcm", "audio")) for c in chunks]
        p.flush()
        torch.cuda.synchronize()
        part = p.sm_partition()
        state = p.get_state().copy()
        p.close()
        return outs, state, part

    one, st1, _ = run(False)
    two, st2, part = run(True)
    assert (part[0] > 0) == (case in ("jumps", "windows")), (case, part)
    for x, y in zip(one, two):
        for k in x:
            assert torch.equal(x[k], y[k]), (case, k)
    assert np.array_equal(st1, st2)


@pytest.mark.parametrize("pinned", [True, False])
def test_pipelined_host_path(dy4, pinned):
    """process_host on a pipelined pipeline: whole calls through two staging sets, queued back to back (contiguous rows go up as
    one 1-D copy, a column slab of a wider array as a 2-D copy), outputs read after sync().  Bits of the synchronous host path."""
    import torch
    m = dy4.mode_params(0)
    S, nb, calls = 6, 5, 5
    iq = dy4.synth.make_batch(0, S, calls * nb * m.block_size // 2, base_seed=733)
    whole = torch.from_numpy(iq).pin_memory().numpy() if pinned else iq
    chunks = [np.ascontiguousarray(whole[:, k * nb * m.block_size:(k + 1) * nb * m.block_size]) if k % 2 else whole[:, k * nb * m.block_size:(k + 1) * nb * m.block_size]
              for k in range(calls)]                          # odd calls: contiguous rows; even calls: a slab of the wide array

    def run(pipelined):
        p = dy4.Pipeline(0, 1, S, pipelined=pipelined)
        outs = [p.process_host(c, want=("pcm", "audio")) for c in chunks]
        p.sync()
        state = p.get_state().copy()
        p.close()
        return outs, state

    one, st1 = run(False)
    two, st2 = run(True)
    for x, y in zip(one, two):
        assert np.array_equal(x["pcm"], y["pcm"]) and np.array_equal(bits(x["audio"]), bits(y["audio"]))
    assert np.array_equal(st1, st2)


@pytest.mark.parametrize("mode", [1, 2, 3])
def test_pipelined_calls_other_modes(dy4, mode):
    """Overlapped calls in modes 1-3 (other block sizes and IF rates; modes 2 and 3 through the polyphase audio kernels), device and
    host path, exact and default audio arithmetic: the bits of joined calls."""
    import torch
    m = dy4.mode_params(mode)
    S = 10
    cuts = [0, 3, 4, 13, 14, 22]
    d = dy4.synth.make_batch_torch(mode, S, cuts[-1] * m.block_size // 2, base_seed=1200 + mode, device="cuda")
    h = d.cpu().pin_memory()
    for exact in (False, True):
        def run(pipelined):
            p = dy4.Pipeline(mode, 1, S, exact_audio=exact, pipelined=pipelined)
            outs = []
            for k, (a, b) in enumerate(zip(cuts, cuts[1:])):
                sl = slice(a * m.block_size, b * m.block_size)
                o = p.process_host(h[:, sl].numpy(), want=("pcm",)) if k == 2 else p.process(d[:, sl], want=("pcm",))
                outs.append(o["pcm"])
            p.flush()
            p.sync()
            torch.cuda.synchronize()
            outs = [o if isinstance(o, np.ndarray) else o.cpu().numpy() for o in outs]
            p.close()
            return outs
        for k, (x, y) in enumerate(zip(run(False), run(True))):
            assert np.array_equal(x, y), (mode, exact, k)


def test_pipelined_calls_random_lengths(dy4):
    """Thirty overlapped calls of random length (1 .. 20 blocks: one sub-chunk, uniform sub-chunks, row and workspace regrowth on the
    way), device and host path mixed, against the same calls joined one by one."""
    import torch
    m = dy4.mode_params(0)
    S = 33
    rng = np.random.default_rng(12)
    lens = [int(x) for x in rng.integers(1, 21, size=30)]
    total = sum(lens)
    d = dy4.synth.make_batch_torch(0, S, total * m.block_size // 2, base_seed=911, device="cuda")
    h = d.cpu().pin_memory()
    host_calls = set(int(x) for x in rng.choice(30, size=8, replace=False))

    def run(pipelined):
        p = dy4.Pipeline(0, 1, S, pipelined=pipelined)
        outs, b = [], 0
        for k, nb in enumerate(lens):
            sl = slice(b * m.block_size, (b + nb) * m.block_size)
            if k in host_calls:
                outs.append(p.process_host(h[:, sl].numpy(), n_blocks=nb, want=("pcm",))["pcm"])
            else:
                outs.append(p.process(d[:, sl], n_blocks=nb, want=("pcm",))["pcm"])
            b += nb
        p.flush()
        p.sync()
        torch.cuda.synchronize()
        outs = [o if isinstance(o, np.ndarray) else o.cpu().numpy() for o in outs]
        p.close()
        return outs

    one, two = run(False), run(True)
    for k, (x, y) in enumerate(zip(one, two)):
        assert np.array_equal(x, y), (k, lens[k], k in host_calls)


def test_sm_partition_does_not_change_results(dy4, monkeypatch):
    """DY4_LOOP_SMS (opt-in, dy4_smpart.cu): the serial loops on a green context of 32 SMs, every other kernel on the rest; the
    caller's stream forks into the partition's streams and joins at the end of the call.  Same bits out, call after call."""
    import torch
    m = dy4.mode_params(0)
    S, nb = 24, 12
    d = dy4.synth.make_batch_torch(0, S, 2 * nb * m.block_size // 2, base_seed=301, device="cuda")

    def run():
        p = dy4.Pipeline(0, 1, S)
        a = p.process(d[:, :nb * m.block_size], want=("pcm", "audio"))
        b = p.process(d[:, nb * m.block_size:], want=("pcm", "audio"))
        torch.cuda.synchronize()
        part = p.sm_partition()
        p.close()
        return a, b, part

    *one, part = run()
    assert part == (0, 0)
    monkeypatch.setenv("DY4_LOOP_SMS", "32")
    *two, part = run()
    assert part[0] >= 32 and sum(part) == torch.cuda.get_device_properties(0).multi_processor_count, part
    for x, y in zip(one, two):
        for k in x:
            assert torch.equal(x[k], y[k]), k


@pytest.mark.parametrize("mode,stereo", [(0, 1), (1, 0), (3, 1)])
def test_host_path_equals_device_path(dy4, mode, stereo):
    import torch
    m = dy4.mode_params(mode)
    iq = dy4.synth.make_batch(mode, 4, 7 * m.block_size // 2, base_seed=77)
    dev, _ = run_gpu(dy4, mode, stereo, iq)
    host, _ = run_gpu(dy4, mode, stereo, iq, host=True, chunk_blocks=2)
    assert np.array_equal(host["pcm"], dev["pcm"]) and np.array_equal(bits(host["audio"]), bits(dev["audio"]))
    pinned = torch.from_numpy(iq).pin_memory()
    host2, _ = run_gpu(dy4, mode, stereo, pinned.numpy(), host=True)     # default chunking, pinned input
    assert np.array_equal(host2["pcm"], dev["pcm"])


def test_checkpoint_restore(dy4):
    import torch
    m = dy4.mode_params(0)
    iq = dy4.synth.make_batch(0, 2, 6 * m.block_size // 2, base_seed=5)
    d = torch.from_numpy(iq).cuda()
    p = dy4.Pipeline(0, 1, 2)
    full = p.process(d, want=("pcm",))["pcm"].cpu().numpy()
    p.reset()
    p.process(d[:, :3 * m.block_size], want=("pcm",))
    torch.cuda.synchronize()
    ckpt = p.get_state()
    p.close()
    q = dy4.Pipeline(0, 1, 2)
    q.set_state(ckpt)
    rest = q.process(d[:, 3 * m.block_size:], want=("pcm",))["pcm"].cpu().numpy()
    assert np.array_equal(rest, full[:, full.shape[1] // 2:])
    q.close()


def test_ragged_and_empty_inputs(dy4):
    import torch
    m = dy4.mode_params(0)
    iq = dy4.synth.make_batch(0, 2, 2 * m.block_size // 2 + 1234, base_seed=9)   # trailing partial block
    p = dy4.Pipeline(0, 0, 2)
    d = torch.zeros((2, 3 * m.block_size), dtype=torch.uint8, device="cuda")
    d[:, :iq.shape[1]] = torch.from_numpy(iq).cuda()
    out = p.process(d[:, :iq.shape[1] // 16 * 16], want=("pcm",))                 # n_blocks = floor(bytes/block): 2
    assert out["pcm"].shape == (2, 2 * m.audio_per_block)
    p.reset()
    assert p.process(d, n_blocks=0, want=())== {}
    with pytest.raises(dy4.Dy4Error):
        p.process(d[:, 1:], n_blocks=1, want=("pcm",))                           # misaligned rows are refused
    p.close()


def test_silence(dy4):
    m = dy4.mode_params(1)
    iq = np.full((2, 2 * m.block_size), 128, np.uint8)
    out, _ = run_gpu(dy4, 1, 1, iq)
    assert not out["if"].any() and not out["pcm"].any()


# ---------------------------------------------------------------- filter.h compatibility tier, per op
def test_filterh_ops_against_checker(dy4, checker):
    fh = dy4.filterh
    rng = np.random.default_rng(3)
    x = rng.standard_normal(7000).astype(np.float32)
    for nh in (101, 37, 1):
        h = rng.standard_normal(nh).astype(np.float32)
        ns = max(nh - 1, 1)
        sa, sb = rng.standard_normal(ns).astype(np.float32), None
        sb = sa.copy()
        for blk in (x[:3000], x[3000:7000]):
            assert np.array_equal(bits(fh.blockConvolveFIR(blk, h, sa)), bits(checker.block_fir(blk, h, sb)))
            assert np.array_equal(sa, sb)
        sa = np.zeros(ns, np.float32); sb = sa.copy()
        for blk in (x[:3000], x[3000:7000]):
            assert np.array_equal(bits(fh.downsampleBlockConvolveFIR(10, blk, h, sa)), bits(checker.decim_fir(10, blk, h, sb)))
            assert np.array_equal(sa, sb)
    hh = rng.standard_normal(101 * 147).astype(np.float32)
    sa = np.zeros(100, np.float32); sb = sa.copy()
    for blk in (x[:1600], x[1600:6400]):
        assert np.array_equal(bits(fh.resampleBlockConvolveFIR(147, 800, blk, hh, sa)), bits(checker.resample_fir(147, 800, blk, hh, sb)))
        assert np.array_equal(sa, sb)
    I, Q = x[:3000].copy(), x[3000:6000].copy()
    I[7] = Q[7] = 0.0
    pb = [0.5, -0.25]
    got, pi, pq = fh.fmDemodArctan(I, Q, 0.5, -0.25)
    assert np.array_equal(bits(got), bits(checker.fm_demod(I, Q, pb))) and [pi, pq] == pb
    pil = (0.1 * np.sin(2 * np.pi * 19e3 / 240e3 * np.arange(40000) + 1.0)).astype(np.float32)
    pil[50] = 0.0
    for scale, adj, bw in ((2.0, 0.0, 0.01), (0.5, 0.3, 0.001)):
        sa = np.array([1, 0, 0, 0, 0, 1], np.float32); sb = sa.copy()
        for a in (0, 20000):
            assert np.array_equal(bits(fh.fmPLL(pil[a:a + 20000], 19e3, 240e3, scale, adj, bw, sa)),
                                  bits(checker.pll(pil[a:a + 20000], 19e3, 240e3, scale, adj, bw, sb)))
            assert np.array_equal(bits(sa), bits(sb))
    sa = np.arange(50, dtype=np.float32); sb = sa.copy()
    assert np.array_equal(fh.delayBlock(x[:600], sa), checker.delay_block(x[:600], sb)) and np.array_equal(sa, sb)
    raw = rng.integers(0, 256, 5000, dtype=np.uint8)
    assert np.array_equal(bits(fh.iqToFloat(raw)), bits(checker.iq_to_float(raw)))
    a, b = x[:1000], x[1000:1900]
    assert np.array_equal(fh.pointwiseMultiply(a, b), (a[:900] * b * np.float32(2)))
    assert np.array_equal(fh.pointwiseAdd(a, x[2000:3000]), a + x[2000:3000])
    assert np.array_equal(fh.pointwiseSubtract(a, x[2000:3000]), a - x[2000:3000])
    il = fh.interleave(a, x[2000:3000])
    assert np.array_equal(il[0::2], a) and np.array_equal(il[1::2], x[2000:3000])
    assert np.array_equal(fh.downsample(a, 7), a[::7])
    up = fh.upsample(a[:100], 3)
    assert np.array_equal(up[::3], a[:100]) and not up[1::3].any() and not up[2::3].any()
    full = fh.convolveFIR(a, x[5000:5101])
    ref = checker.block_fir(np.concatenate([a, np.zeros(100, np.float32)]), x[5000:5101], np.zeros(100, np.float32))
    assert np.array_equal(bits(full), bits(ref))


# ---------------------------------------------------------------- BASELINE-size properties
def test_config2_size_batch_independence_and_sampled_parity(dy4, checker):
    """configs[1]: mode 0 stereo, 256 streams.  Streams are independent, so (a) a stream's output does not
    depend on its slot or its neighbours, (b) a sample of streams equals the CPU checker."""
    import torch
    m = dy4.mode_params(0)
    S, nb = 256, 6
    base = dy4.synth.make_batch(0, 8, nb * m.block_size // 2, base_seed=65)
    idx = np.arange(S) % 8
    iq = base[idx]
    out, _ = run_gpu(dy4, 0, 1, iq, want=("pcm", "audio"))
    for s in range(8, S):
        assert np.array_equal(out["pcm"][s], out["pcm"][idx[s]])
    for s in (0, 3, 7):
        ref = checker.pipeline(0, 1, base[s])
        assert rel_l2(out["audio"][s], ref["audio"]) <= TOL_L2
        assert np.abs(out["pcm"][s].astype(np.int32) - ref["pcm"]).max() <= 1


def test_linearity_of_fir_stages_on_device(dy4):
    """Size-independent property: the compat-tier FIR is linear within float rounding."""
    fh = dy4.filterh
    rng = np.random.default_rng(8)
    x, y = rng.standard_normal(20000).astype(np.float32), rng.standard_normal(20000).astype(np.float32)
    h = fh.impulseResponseLPF(2.4e6, 100e3, 101)
    z = lambda: np.zeros(100, np.float32)
    a = fh.downsampleBlockConvolveFIR(10, x, h, z()) + fh.downsampleBlockConvolveFIR(10, y, h, z())
    b = fh.downsampleBlockConvolveFIR(10, x + y, h, z())
    assert rel_l2(a, b) < 1e-6


# ---------------------------------------------------------------- the reference's own project.cpp over this library
@pytest.mark.parametrize("mode,stereo", [(0, 1), (2, 1), (1, 0)])
def test_reference_project_binary_linked_against_this_library(dy4, mode, stereo):
    """oracle/_ref/project_b200 = the reference's UNMODIFIED project.cpp/iofunc.cpp objects linked against
    libdy4b200.so instead of the reference's filter.o.  Fed the golden input on stdin it must
    write the golden PCM: every filter.h call it makes runs through the filter.h shim and the CUDA kernels."""
    import os
    import subprocess
    exe = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref", "project_b200")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/project_b200 not built (needs /root/reference at build time)")
    iq = golden("mode%d_stereo.npz" % mode)["iq"]
    want = golden("mode%d_%s.npz" % (mode, "stereo" if stereo else "mono"))["pcm"]
    before = dy4.launch_count()
    p = subprocess.run([exe, str(mode), "stereo" if stereo else "mono"], input=iq.tobytes(), capture_output=True, timeout=300)
    assert p.returncode == 1, p.stderr[-500:]                      # the reference exits 1 at end of input
    got = np.frombuffer(p.stdout, np.int16)
    assert got.size == want.size
    assert np.array_equal(got, want)                               # compat tier is unfused everywhere: bit-exact PCM


def test_more_than_65535_streams(dy4, checker):
    """config 5's upper end: 66 000 streams in one batch (streams ride on grid.x, which has no 65 535 limit)."""
    import torch
    m = dy4.mode_params(1)
    S, nb = 66000, 1
    base = dy4.synth.make_batch(1, 8, nb * m.block_size // 2, base_seed=42)
    iq = torch.from_numpy(base).cuda().repeat(S // 8, 1).contiguous()
    p = dy4.Pipeline(1, 1, S)
    out = p.process(iq, want=("pcm",))["pcm"]
    torch.cuda.synchronize()
    first = out[:8]
    assert bool((out.view(S // 8, 8, -1) == first.unsqueeze(0)).all())
    for s in (0, 5):
        ref = checker.pipeline(1, 1, base[s])
        assert np.abs(first[s].cpu().numpy().astype(np.int32) - ref["pcm"]).max() <= 1
    p.close()


@pytest.mark.parametrize("mode,stereo,per_call", [(0, 1, 1), (0, 0, 3), (2, 1, 2), (1, 1, 5)])
def test_streaming_command_line(dy4, mode, stereo, per_call):
    """dy4_project: the reference's `project <mode> <mono|stereo>` boundary (uint8 IQ on stdin, int16 PCM on stdout,
    partial trailing block dropped, exit status 1 at end of input) over the throughput tier, one stream.  The bytes on
    stdout must be the reference's for any blocks-per-call cadence."""
    import os
    import subprocess
    exe = os.path.join(dy4.PACKAGE_DIR, "dy4_project")
    assert os.path.exists(exe), "dy4_project not built (make -C csrc)"
    iq = golden("mode%d_stereo.npz" % mode)["iq"]
    want = golden("mode%d_%s.npz" % (mode, "stereo" if stereo else "mono"))["pcm"]
    data = iq.tobytes() + b"\x80" * 1000                           # a partial trailing block: dropped (project.cpp:293-296)
    p = subprocess.run([exe, str(mode), "stereo" if stereo else "mono", str(per_call)], input=data, capture_output=True, timeout=300)
    assert p.returncode == 1, p.stderr[-500:]
    assert b"End of input stream reached" in p.stderr
    got = np.frombuffer(p.stdout, np.int16)
    assert got.size == want.size and np.array_equal(got, want)
    bad = subprocess.run([exe, "7", "stereo"], input=b"", capture_output=True, timeout=60)
    assert bad.returncode == 1 and b"Wrong mode" in bad.stderr


# ---------------------------------------------------------------- table-driven PLL through the software pipeline
@pytest.mark.gpu
def test_table_pll_multi_subchunk_calls_with_signal_jump(dy4, checker, monkeypatch):
    """16-block calls are cut into five sub-chunks, so the prediction carries its own state (and its turn correction)
    from the third one on; feeding the same 16 blocks AGAIN continues the streams across a phase jump of the pilot
    (the loop re-acquires lock).  Results must equal the reference run over the concatenated input, bit for bit, and
    the direct loop (DY4_PLL_TABLE_MAX=0) must give the same."""
    import torch
    mode, S, nb = 0, 3, 16
    m = dy4.mode_params(mode)
    iq = dy4.synth.make_batch(mode, S, nb * m.block_size // 2, base_seed=41)
    d = torch.from_numpy(iq).cuda()
    both = np.concatenate([iq, iq], axis=1)
    ref = [checker.pipeline(mode, 1, both[s]) for s in range(S)]

    def run():
        p = dy4.Pipeline(mode, 1, S, exact_audio=True)          # no debug_rows: the real sub-chunk plan
        try:
            outs = [p.process(d, want=("pcm", "audio", "if")) for _ in range(2)]
            torch.cuda.synchronize()
            return {k: torch.cat([o[k] for o in outs], 1).cpu().numpy() for k in outs[0]}
        finally:
            p.close()

    tab = run()
    for s in range(S):
        assert np.array_equal(bits(tab["if"][s]), bits(ref[s]["if"]))
        assert np.array_equal(bits(tab["audio"][s]), bits(ref[s]["audio"])), "stream %d" % s
        assert np.array_equal(tab["pcm"][s], ref[s]["pcm"])
    monkeypatch.setenv("DY4_PLL_TABLE_MAX", "0")
    direct = run()
    for k in tab:
        assert np.array_equal(direct[k], tab[k]), k


def test_table_pll_streams_that_never_lock(dy4, checker):
    """Input without a pilot (uniform noise bytes; an FM carrier with mono audio only): the PLL never locks, the
    prediction is useless, every super-group is redone and the loop falls back to direct evaluation — the results are
    still the reference's, bit for bit."""
    import torch
    mode, nb = 0, 10
    m = dy4.mode_params(mode)
    n_pairs = nb * m.block_size // 2
    rng = np.random.default_rng(11)
    noise = rng.integers(0, 256, size=2 * n_pairs, dtype=np.uint8)
    t = np.arange(n_pairs) / 2.4e6
    ph = 2 * np.pi * 75e3 * np.cumsum(0.5 * np.sin(2 * np.pi * 1e3 * t)) / 2.4e6          # mono programme, no 19 kHz pilot
    mono = np.empty(2 * n_pairs, np.uint8)
    mono[0::2] = np.clip(np.rint(100 * np.cos(ph) + 128), 0, 255)
    mono[1::2] = np.clip(np.rint(100 * np.sin(ph) + 128), 0, 255)
    iq = np.stack([noise, mono, dy4.synth.make_stream(mode, n_pairs, 65)])
    p = dy4.Pipeline(mode, 1, 3, exact_audio=True)
    try:
        outs = [p.process(torch.from_numpy(iq).cuda(), want=("pcm", "audio", "if")) for _ in range(2)]
        torch.cuda.synchronize()
        got = {k: torch.cat([o[k] for o in outs], 1).cpu().numpy() for k in outs[0]}
    finally:
        p.close()
    both = np.concatenate([iq, iq], axis=1)
    for s in range(3):
        ref = checker.pipeline(mode, 1, both[s])
        assert np.array_equal(bits(got["if"][s]), bits(ref["if"])), s
        assert np.array_equal(bits(got["audio"][s]), bits(ref["audio"])), s
        assert np.array_equal(got["pcm"][s], ref["pcm"]), s
