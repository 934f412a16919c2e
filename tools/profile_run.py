"""Small fixed workload for ncu captures: mode 0 stereo, S streams x NB blocks, REP passes."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, dy4_b200
S = int(os.environ.get("S", 256)); NB = int(os.environ.get("NB", 4)); REP = int(os.environ.get("REP", 2)); MODE = int(os.environ.get("MODE", 0))
m = dy4_b200.mode_params(MODE)
iq = dy4_b200.synth.make_batch_torch(MODE, min(S, 256), NB * m.block_size // 2, base_seed=65, device="cuda", rds=bool(int(os.environ.get("RDS", 0))),
                                     periodic=dy4_b200.synth.whole_cycles(MODE, NB * m.block_size // 2))   # NB = 24, 48, ...: the second pass continues the first
if S > 256: iq = iq.repeat((S + 255) // 256, 1)[:S].contiguous()
RDS = bool(int(os.environ.get("RDS", 0)))
p = dy4_b200.Pipeline(MODE, 1, S, debug_rows=bool(int(os.environ.get("WHOLE", 0))), rds=RDS)   # WHOLE=1: one sub-chunk per call (whole-job launches)
for _ in range(REP):
    out = p.process(iq, want=("pcm",))
torch.cuda.synchronize()
print("ok", out["pcm"].shape, dy4_b200.launch_count())
