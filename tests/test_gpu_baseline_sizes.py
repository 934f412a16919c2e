"""GPU suite (-m gpu), part 2: parity at the sizes BASELINE.json's configs are quoted on, against the CPU checker
(the reference's own code, oracle/_ref, when it was built, else the C restatement) run on all host cores.

  config 1/2 length   mode 0 stereo, 468 blocks (9.98 s, 2.4 M PLL steps: trigArg's float ulp reaches 2^-3 rad), 16 streams
  config 3            modes 2 and 3 (147/800, 147/1280 polyphase resamplers), 1 024 streams, 16 of them against the checker
  config 4            mode 0 stereo + RDS path, 4 096 streams, 16 of them against oracle/rds.py
  counter saturation  the PLL's float sample counter sticks at 2^24 (filter.cpp:213)

Tolerances as in test_gpu_parity.py: IF / pilot / NCO bit-exact, audio rel-L2 <= 1e-5, PCM within 1 LSB, RDS integers exact."""
import multiprocessing as mp
import os

import numpy as np
import pytest

from conftest import bits, rel_l2

pytestmark = pytest.mark.gpu
TOL_L2 = 1e-5

_JOB = {}


def _cpu_one(s):
    import oracle
    chk = oracle.load(_JOB["kind"])
    return chk.pipeline(_JOB["mode"], _JOB["stereo"], _JOB["iq"][s], want=_JOB["want"])


def cpu_parallel(checker, mode, stereo, iq_rows, want):
    """One reference pipeline per stream on all host cores (processes: the reference replay is not thread-safe)."""
    _JOB.update(kind=checker.kind, mode=mode, stereo=stereo, iq=iq_rows, want=want)
    ctx = mp.get_context("fork")
    with ctx.Pool(min(len(iq_rows), os.cpu_count() or 1)) as pool:
        return pool.map(_cpu_one, range(len(iq_rows)))


def test_config1_length_stereo_ten_seconds(dy4, checker):
    """468 blocks of mode 0 stereo, 16 streams: IF and NCO bit-exact over 2.4 M PLL steps, audio / PCM within tolerance — both
    in one launch per kernel (debug rows, where the pilot / NCO rows can be read back), through the pipelined sub-chunks of one
    call, and as ten overlapped calls (DY4_FLAG_PIPELINED)."""
    import torch
    m = dy4.mode_params(0)
    S, nb = 16, 468
    d = dy4.synth.make_batch_torch(0, S, nb * m.block_size // 2, base_seed=1465, device="cuda")
    iq = d.cpu().numpy()
    refs = cpu_parallel(checker, 0, 1, iq, ("if", "audio", "pcm", "nco"))
    p = dy4.Pipeline(0, 1, S, debug_rows=True)
    out = p.process(d, want=("pcm", "audio", "if"))
    torch.cuda.synchronize()
    pilot, nco = p.debug_pilot_nco()
    p.close()
    q = dy4.Pipeline(0, 1, S)                                       # joined calls: geometric sub-chunks, four CUDA streams
    out2 = q.process(d, want=("pcm", "audio"))
    torch.cuda.synchronize()
    q.close()
    r = dy4.Pipeline(0, 1, S, pipelined=True)                       # the bench's path: overlapped calls of 48 blocks (and a last one of 36)
    parts = [r.process(d[:, b * m.block_size:min(b + 48, nb) * m.block_size], want=("pcm",))["pcm"] for b in range(0, nb, 48)]
    r.flush()
    torch.cuda.synchronize()
    r.close()
    assert torch.equal(torch.cat(parts, 1), out2["pcm"])
    worst = 0.0
    for s in range(S):
        ref = refs[s]
        assert np.array_equal(bits(out["if"][s].cpu().numpy()), bits(ref["if"])), s
        assert np.array_equal(bits(nco[s].cpu().numpy()), bits(ref["nco"])), s
        for o in (out, out2):
            a = o["audio"][s].cpu().numpy()
            worst = max(worst, rel_l2(a, ref["audio"]))
            assert rel_l2(a, ref["audio"]) <= TOL_L2, s
            assert np.abs(o["pcm"][s].cpu().numpy().astype(np.int32) - ref["pcm"]).max() <= 1, s
        assert torch.equal(out["pcm"][s], out2["pcm"][s])
    print("468 blocks x 16 streams: IF and NCO bit-exact, worst audio rel-L2 %.2e" % worst)


@pytest.mark.parametrize("mode", [2, 3])
def test_config3_polyphase_1024_streams(dy4, checker, mode):
    """configs[2]: 1 024 streams through the 147/800 (147/1280) polyphase resamplers, 16 sampled streams against the checker;
    every stream against its twin (the batch holds 512 distinct streams twice)."""
    import torch
    m = dy4.mode_params(mode)
    S, nb = 1024, 12
    d = dy4.synth.make_batch_torch(mode, 512, nb * m.block_size // 2, base_seed=3000 + mode, device="cuda")
    d = d.repeat(2, 1).contiguous()
    sample = sorted(set(list(range(0, S, 67)) + [S - 1]))[:17]
    refs = cpu_parallel(checker, mode, 1, d[sample].cpu().numpy(), ("if", "audio", "pcm"))
    for exact in (False, True):
        p = dy4.Pipeline(mode, 1, S, exact_audio=exact)
        out = p.process(d, want=("pcm", "audio", "if"))
        torch.cuda.synchronize()
        p.close()
        assert torch.equal(out["pcm"][:512], out["pcm"][512:])
        for s, ref in zip(sample, refs):
            assert np.array_equal(bits(out["if"][s].cpu().numpy()), bits(ref["if"])), (s, exact)
            a, pcm = out["audio"][s].cpu().numpy(), out["pcm"][s].cpu().numpy()
            if exact:
                assert np.array_equal(bits(a), bits(ref["audio"])) and np.array_equal(pcm, ref["pcm"]), s
            else:
                assert rel_l2(a, ref["audio"]) <= TOL_L2 and np.abs(pcm.astype(np.int32) - ref["pcm"]).max() <= 1, s


def test_config4_rds_4096_streams(dy4, orc):
    """configs[3]: 4 096 stereo receivers with the RDS path; 16 sampled streams against oracle/rds.py (RRC baseband within
    tolerance, symbols / bits / frame-sync events / groups exact)."""
    import torch
    from oracle import rds
    m = dy4.mode_params(0)
    S, nb = 4096, 60
    d = dy4.synth.make_batch_torch(0, 512, nb * m.block_size // 2, base_seed=4400, device="cuda", rds=True)
    d = d.repeat(8, 1).contiguous()
    p = dy4.Pipeline(0, 1, S, rds=True)
    p.process(d, want=("pcm",))
    ri, rq = p.rds_read()
    torch.cuda.synchronize()
    dr = p.rds_drain()
    p.close()
    sample = sorted(set(list(range(0, S, 271)) + [S - 1]))[:16]
    n_events = 0
    for s in sample:
        iq = d[s].cpu().numpy()
        r = rds.rds_front(orc.pipeline(0, 1, iq, want=("if",))["if"])
        assert rel_l2(ri[s].cpu().numpy(), r["rrc_i"]) <= TOL_L2 and rel_l2(rq[s].cpu().numpy(), r["rrc_q"]) <= TOL_L2, s
        be = rds.rds_back(r["rrc_i"], r["rrc_q"])
        assert np.array_equal(dr[s]["symbols"], np.concatenate([np.array(x, np.int8) for x in be.symbols])), s
        assert np.array_equal(dr[s]["bits"], np.array(be.bits, np.int8)), s
        assert np.array_equal(dr[s]["events"], np.array(be.events, np.int32).reshape(-1, 4)), s
        assert np.array_equal(dr[s]["groups"], np.array(be.groups, np.int32).reshape(-1, 4)), s
        n_events += len(be.events)
    assert n_events >= 4 * len(sample)
    for s in (1, 700, 4095):                                          # twins (512 distinct streams, eight times)
        for k in ("symbols", "bits", "events", "groups"):
            assert np.array_equal(dr[s][k], dr[s % 512][k]), (s, k)


def test_pll_sample_counter_saturation(dy4, checker):
    """filter.cpp:213: trigOffset is a FLOAT counter; at 2^24 `trigOffset++` stops changing it.  Preset the carried counter to
    2^24 - 1000 (69.9 s into a stream) and run 3 blocks through the table-driven loop: NCO row and carried state are the
    reference's, bit for bit, across the saturation point."""
    import torch
    m = dy4.mode_params(0)
    nb = 3
    iq = dy4.synth.make_stream(0, nb * m.block_size // 2, 8123)
    T0 = float(2 ** 24 - 1000)
    p = dy4.Pipeline(0, 1, 1, debug_rows=True)
    st = p.get_state()
    off = 224 + 256 * 4 + 128 * 4                                     # iq_tail, if_tail, mix_tail precede the PLL state (dy4_pipeline_get_state)
    pll = st[off:off + 32].view(np.float32)
    assert list(pll[:6]) == [1.0, 0.0, 0.0, 0.0, 0.0, 1.0]            # PLLState, project.cpp:46-53
    pll[4] = T0
    p.set_state(st)
    p.process(torch.from_numpy(iq[None]).cuda(), want=("pcm",))
    torch.cuda.synchronize()
    pilot, nco = [t.cpu().numpy()[0] for t in p.debug_pilot_nco()]
    st2 = p.get_state()[off:off + 32].view(np.float32).copy()
    p.close()
    ref_state = np.array([1, 0, 0, 0, T0, 1], np.float32)
    ref_nco = checker.pll(pilot, 19e3, 240e3, 2.0, 0.0, 0.01, ref_state)
    assert np.array_equal(bits(nco), bits(ref_nco))
    assert np.array_equal(bits(st2[:6]), bits(ref_state))
    assert st2[4] == 2.0 ** 24                                        # it stuck


def test_filterh_shim_from_a_thousand_short_lived_threads(dy4):
    """The reference's main() runs frontend() and backend() on two fresh std::threads per block (project.cpp:299-305); linked
    against this library, every such thread calls the filter.h shim once.  Scratch buffers and streams are leased from a pool
    and handed back at thread exit: device memory stays flat over 1 000 threads."""
    import threading
    import torch
    fh = dy4.filterh
    rng = np.random.default_rng(1)
    x = rng.standard_normal(5120).astype(np.float32)
    h = rng.standard_normal(101).astype(np.float32)
    want = fh.blockConvolveFIR(x, h, np.zeros(100, np.float32))
    errs = []

    def work():
        try:
            got = fh.blockConvolveFIR(x, h, np.zeros(100, np.float32))
            if not np.array_equal(bits(got), bits(want)):
                errs.append("mismatch")
        except Exception as e:                                        # noqa: BLE001
            errs.append(repr(e))

    def burst(n):
        for _ in range(n // 2):
            ts = [threading.Thread(target=work) for _ in range(2)]
            for t in ts:
                t.start()
            for t in ts:
                t.join()

    burst(50)
    torch.cuda.synchronize()
    free0 = torch.cuda.mem_get_info()[0]
    burst(1000)
    torch.cuda.synchronize()
    free1 = torch.cuda.mem_get_info()[0]
    assert not errs, errs[:3]
    assert free0 - free1 < 8 << 20, (free0, free1)
