"""Small workload that touches every kernel of the library, for compute-sanitizer (memcheck / racecheck / initcheck)
where that tool is available (it is closed on the round-1 GPU pool, so this ran plain there: 470 launches, no error):
    compute-sanitizer --tool memcheck python tools/sanitize_run.py
All four modes mono and stereo (device and host paths, odd stream counts, ragged block counts), the RDS path with its
decoder, the Fourier entry points, the filter.h compatibility tier."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, dy4_b200

for mode in (0, 1, 2, 3):
    m = dy4_b200.mode_params(mode)
    for stereo in (1, 0):
        S, nb = 3, 9 if mode == 0 else 3
        iq = dy4_b200.synth.make_batch(mode, S, nb * m.block_size // 2, base_seed=5 + mode)
        for exact in (False, True):
            p = dy4_b200.Pipeline(mode, stereo, S, exact_audio=exact)
            out = p.process(torch.from_numpy(iq).cuda(), want=("pcm", "audio", "if"))
            out = p.process(torch.from_numpy(iq[:, :2 * m.block_size].copy()).cuda(), want=("pcm",))     # a second, shorter call
            p.process_host(iq, want=("pcm",))
            st = p.get_state(); p.set_state(st)
            p.close()
# overlapped calls (DY4_FLAG_PIPELINED): two sets of call rows, IF history ring, loops on SMs of their own where green contexts exist
m = dy4_b200.mode_params(0)
iq = torch.from_numpy(dy4_b200.synth.make_batch(0, 5, 40 * m.block_size // 2, base_seed=9)).cuda()
p = dy4_b200.Pipeline(0, 1, 5, pipelined=True)
outs = [p.process(iq[:, a * m.block_size:b * m.block_size], want=("pcm", "audio", "if")) for a, b in ((0, 9), (9, 10), (10, 26), (26, 40))]
p.flush(); torch.cuda.synchronize()
st = p.get_state(); p.set_state(st); p.reset()
outs = [p.process(iq[:, :12 * m.block_size], want=("pcm",)) for _ in range(3)]
p.close()
m = dy4_b200.mode_params(0)
iq = dy4_b200.synth.make_batch(0, 3, 19 * m.block_size // 2, base_seed=77, rds=True)
p = dy4_b200.Pipeline(0, 1, 3, rds=True)
for a, b in ((0, 4), (4, 15), (15, 19)):                      # 19 blocks: one whole model block decoded at least four times
    p.process(torch.from_numpy(iq[:, a * m.block_size:b * m.block_size].copy()).cuda(), want=("pcm",))
    p.rds_read()
torch.cuda.synchronize()
p.rds_drain(); st = p.get_state(); p.set_state(st); p.close()
x = np.random.default_rng(0).uniform(-1, 1, 300).astype(np.float32)
X = dy4_b200.fourierh.DFT(x); dy4_b200.fourierh.IDFT(X); dy4_b200.fourierh.estimatePSD(np.tile(x, 8), 128, 48000)
F = dy4_b200.filterh
h = F.impulseResponseLPF(240e3, 16e3, 101)
F.blockConvolveFIR(x, h, np.zeros(100, np.float32))
torch.cuda.synchronize()
print("sanitize_run: done, launches", dy4_b200.launch_count())
