"""Times the PLL kernel variants (DY4_PLL_VARIANT is read once per process, so each runs in a child)."""
import os, subprocess, sys
code = r'''
import os, sys
sys.path.insert(0, %r)
import torch, dy4_b200
m = dy4_b200.mode_params(0)
S, NB = 256, 12
iq = dy4_b200.synth.make_batch_torch(0, S, NB * m.block_size // 2, base_seed=65, device="cuda")
p = dy4_b200.Pipeline(0, 1, S)
p.process(iq, want=("pcm",)); torch.cuda.synchronize()
p.profile(True); p.profile_get()
for _ in range(3): p.process(iq, want=("pcm",))
torch.cuda.synchronize()
r = p.profile_get()
n = NB * m.if_per_block
print("threads %%s variant %%s: pll %%.3f ms/launch  %%.1f ns/sample" %% (os.environ.get("DY4_PLL_THREADS"), os.environ.get("DY4_PLL_VARIANT"), r["pll"]["ms"] / 3, r["pll"]["ms"] / 3 * 1e6 / n))
''' % os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for t in os.environ.get("THREADS", "32").split(","):
    for v in sys.argv[1:] or ["0", "1", "2", "3", "4", "5"]:
        subprocess.run([sys.executable, "-c", code], env=dict(os.environ, DY4_PLL_VARIANT=v, DY4_PLL_THREADS=t))
