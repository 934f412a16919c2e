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
if rank == 0:
    want = np.stack([o.pipeline(mode, 1, iq[s])["pcm"] for s in range(S)])
    assert t == 10.0 + world - 1 and tot == S, (t, tot)
    assert np.array_equal(full.numpy(), want)
    print("OK")
else:
    assert full is None
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
