import sys, os, time
sys.path.insert(0, '/root/repo')
import numpy as np, torch
import dy4_b200
m = dy4_b200.mode_params(0)
S=4; nb=8
iq = dy4_b200.synth.make_batch(0, S, nb*m.block_size//2, base_seed=65)
d = torch.from_numpy(iq).cuda()
p = dy4_b200.Pipeline(0, 1, S)
p.profile(True)
for rep in range(2):
    t=time.time(); out = p.process(d, want=("pcm",)); torch.cuda.synchronize(); print("process", time.time()-t)
