"""GPU suite for the RDS path (BASELINE.json configs[3]).  The reference implements RDS only in its
float64 Python model; the fixture tests/golden/rds_mode0.npz holds outputs of the MODEL'S OWN FUNCTIONS
(tests/golden/make_golden_rds.py) and oracle/rds.py is the numpy restatement pinned to it (tests/test_oracle.py).

Tolerances: the RRC-filtered baseband is float32 here and float64 in the model -> relative L2 <= 1e-5 (north_star's float
tolerance); Manchester symbols, decoded bits and frame-sync events are integers -> exact."""
import numpy as np
import pytest

from conftest import golden, rel_l2

pytestmark = pytest.mark.gpu
TOL_L2 = 1e-5


def run_rds(dy4, iq, splits=None):
    """iq [S, bytes] -> (rrc_i, rrc_q, drained) with the stream fed in the given block counts per call."""
    import torch
    iq = np.atleast_2d(iq)
    m = dy4.mode_params(0)
    nb = iq.shape[1] // m.block_size
    splits = splits or [nb]
    assert sum(splits) == nb
    p = dy4.Pipeline(0, 1, iq.shape[0], rds=True)
    try:
        d = torch.from_numpy(iq).cuda()
        ri, rq, b0 = [], [], 0
        for n in splits:
            p.process(d[:, b0 * m.block_size:(b0 + n) * m.block_size].contiguous(), want=("pcm",))
            i_t, q_t = p.rds_read()
            torch.cuda.synchronize()
            ri.append(i_t.cpu().numpy()); rq.append(q_t.cpu().numpy())
            b0 += n
        return np.concatenate(ri, 1), np.concatenate(rq, 1), p.rds_drain()
    finally:
        p.close()


def golden_iq(dy4):
    g = golden("rds_mode0.npz")
    m = dy4.mode_params(0)
    return g, dy4.synth.make_stream(0, int(g["n_blocks"]) * m.block_size // 2, int(g["seed"]), rds=True)


@pytest.mark.parametrize("splits", [None, [7, 13, 1, 20, 19], [15] * 4])
def test_rds_golden(dy4, splits):
    g, iq = golden_iq(dy4)
    ri, rq, dr = run_rds(dy4, iq, splits)
    assert ri.shape[1] == len(g["rrc_i"])
    print("rel-L2 of the RRC baseband vs the model:", rel_l2(ri[0], g["rrc_i"]), rel_l2(rq[0], g["rrc_q"]))
    assert rel_l2(ri[0], g["rrc_i"]) <= TOL_L2 and rel_l2(rq[0], g["rrc_q"]) <= TOL_L2
    n = len(g["rrc_i64"])
    assert rel_l2(ri[0, :n], g["rrc_i64"]) <= TOL_L2 and rel_l2(rq[0, :n], g["rrc_q64"]) <= TOL_L2
    assert np.array_equal(dr[0]["symbols"], g["symbols"])
    assert np.array_equal(dr[0]["bits"], g["bits"])
    assert np.array_equal(dr[0]["events"], g["events"])
    assert np.array_equal(dr[0]["groups"], g["groups"])
    assert dy4.rds_app.ApplicationLayer().feed_groups(dr[0]["groups"]) == str(g["app_lines"]).split("\n")


def test_rds_batch_against_oracle(dy4, orc):
    """8 seeded streams x 60 blocks (16 model blocks): RRC baseband within tolerance, symbols / bits / events exact."""
    from oracle import rds
    m = dy4.mode_params(0)
    S, nb = 8, 60
    iq = dy4.synth.make_batch(0, S, nb * m.block_size // 2, base_seed=4065, rds=True)
    ri, rq, dr = run_rds(dy4, iq, [23, 37])
    n_events = 0
    for s in range(S):
        r = rds.rds_front(orc.pipeline(0, 1, iq[s], want=("if",))["if"])
        assert rel_l2(ri[s], r["rrc_i"]) <= TOL_L2 and rel_l2(rq[s], r["rrc_q"]) <= TOL_L2
        be = rds.rds_back(r["rrc_i"], r["rrc_q"])
        assert np.array_equal(dr[s]["symbols"], np.concatenate([np.array(x, np.int8) for x in be.symbols]))
        assert np.array_equal(dr[s]["bits"], np.array(be.bits, np.int8))
        assert np.array_equal(dr[s]["events"], np.array(be.events, np.int32).reshape(-1, 4))
        assert np.array_equal(dr[s]["groups"], np.array(be.groups, np.int32).reshape(-1, 4))
        assert dy4.rds_app.ApplicationLayer().feed_groups(dr[s]["groups"]) == be.app_lines
        n_events += len(be.events)
    assert n_events >= 8 * S


def test_rds_decoded_words_are_the_transmitted_ones(dy4):
    """End to end: the 16-bit words the frame synchroniser reports are the ones synth put on the sub-carrier."""
    g, iq = golden_iq(dy4)
    _, _, dr = run_rds(dy4, iq)
    src = dy4.synth.rds_group_bits(4000, int(g["seed"]) + 7919)
    words = {(j % 4, int("".join(str(int(b)) for b in src[26 * j:26 * j + 16]), 2)) for j in range(len(src) // 26)}
    code = {0: 0, 1: 1, 2: 2, 4: 3}
    ev = dr[0]["events"]
    assert len(ev) >= 16
    assert all((code[int(t)], int(w)) in words for t, _, _, w in ev)


def test_rds_drain_empties_and_audio_is_unchanged(dy4):
    """The RDS branch runs beside the stereo path and must not disturb it; a second drain returns nothing."""
    import torch
    m = dy4.mode_params(0)
    iq = dy4.synth.make_batch(0, 3, 4 * m.block_size // 2, base_seed=99, rds=True)     # 3 242 RRC samples: one model block
    d = torch.from_numpy(iq).cuda()
    outs = []
    for rds_on in (False, True):
        p = dy4.Pipeline(0, 1, 3, rds=rds_on, debug_rows=True)
        outs.append(p.process(d, want=("pcm",))["pcm"].cpu().numpy())
        if rds_on:
            first = p.rds_drain()
            again = p.rds_drain()
            assert all(len(x["symbols"]) in (189, 190) for x in first) and all(len(x["symbols"]) == 0 for x in again)
        p.close()
    assert np.array_equal(outs[0], outs[1])


def test_rds_needs_mode0_stereo(dy4):
    with pytest.raises(Exception):
        dy4.Pipeline(2, 1, 2, rds=True)
    with pytest.raises(Exception):
        dy4.Pipeline(0, 0, 2, rds=True)


def test_rds_checkpoint_and_host_path(dy4):
    """Checkpoint after 27 blocks, restore into a fresh receiver, feed the rest: the RRC baseband of the second part and
    every symbol / bit / event are those of the uninterrupted run.  The second part goes through process_host()."""
    import torch
    g, iq = golden_iq(dy4)
    m = dy4.mode_params(0)
    cut = 27 * m.block_size
    ri_all, rq_all, dr_all = run_rds(dy4, iq, [27, 33])
    a = dy4.Pipeline(0, 1, 1, rds=True)
    a.process(torch.from_numpy(iq[None, :cut]).cuda(), want=("pcm",))
    first = a.rds_drain()
    state = a.get_state()
    a.close()
    b = dy4.Pipeline(0, 1, 1, rds=True)
    b.set_state(state)
    b.process_host(iq[None, cut:].copy(), want=("pcm",))
    i_t, q_t = b.rds_read()
    torch.cuda.synchronize()
    second = b.rds_drain()
    b.close()
    n2 = i_t.shape[1]
    assert np.array_equal(i_t.cpu().numpy()[0], ri_all[0, -n2:]) and np.array_equal(q_t.cpu().numpy()[0], rq_all[0, -n2:])
    for k in ("symbols", "bits", "events", "groups"):
        assert np.array_equal(np.concatenate([first[0][k], second[0][k]]), dr_all[0][k]), k
    assert np.array_equal(dr_all[0]["bits"], g["bits"])


def test_rds_batch_independence_at_scale(dy4):
    """Towards BASELINE configs[3] (many streams with the RDS path): in a 768-stream batch every stream's RRC baseband,
    symbols, bits, events and groups are exactly those of the same stream processed alone — no cross-stream coupling at
    any grid size.  Most streams reach frame sync (the model's timing recovery is fragile — replicated, not repaired — so
    not all of them do)."""
    import torch
    m = dy4.mode_params(0)
    S, nb = 768, 60
    d = dy4.synth.make_batch_torch(0, S, nb * m.block_size // 2, base_seed=9000, device="cuda", rds=True)
    p = dy4.Pipeline(0, 1, S, rds=True)
    p.process(d, want=("pcm",))
    ri, rq = p.rds_read()
    torch.cuda.synchronize()
    dr = p.rds_drain()
    p.close()
    assert sum(1 for x in dr if len(x["events"]) >= 8 and len(x["groups"]) >= 1) >= 0.8 * S
    for s in (0, 301, S - 1):
        q = dy4.Pipeline(0, 1, 1, rds=True)
        q.process(d[s:s + 1].contiguous(), want=("pcm",))
        qi, qq = q.rds_read()
        torch.cuda.synchronize()
        one = q.rds_drain()[0]
        q.close()
        assert torch.equal(qi[0], ri[s]) and torch.equal(qq[0], rq[s])
        for k in ("symbols", "bits", "events", "groups"):
            assert np.array_equal(one[k], dr[s][k]), (s, k)
