"""Why the CUDA path keeps the reference's unfused multiply-add order upstream of the PLL.

Builds a variant of the C oracle whose FIR taps are fused (fmaf) — everything else unchanged — and
compares it with the normal oracle over ~10 s mode-0 stereo streams.  IF and pilot move by ~1e-7 (harmless),
but the PLL's float phase accumulator makes the NCO trajectory diverge after ~16.7k IF samples and the stereo
audio ends up ~5e-3 away (L-R channel ~8e-2, PCM hundreds of LSB): DESIGN.md §3.  CPU only.
"""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import importlib.util
import numpy as np
import oracle
from oracle import cpu

spec = importlib.util.spec_from_file_location("synth", os.path.join(ROOT, "3dy4-real-time-software-defined-radio-_b200", "synth.py"))
synth = importlib.util.module_from_spec(spec); spec.loader.exec_module(synth)

so = "/tmp/libdy4oracle_fma.so"
subprocess.check_call(["gcc", "-O2", "-std=c11", "-ffp-contract=off", "-mfma", "-DDY4_ORACLE_FIR_FMA", "-fPIC", "-shared",
                       "-o", so, os.path.join(ROOT, "oracle", "dy4_oracle.c"), "-lm"])
cpu._PATHS["fma"] = (so, "dy4o_")
o, f = oracle.load("oracle"), oracle.load("fma")
rl2 = lambda a, b: float(np.linalg.norm(a.astype(np.float64) - b) / np.linalg.norm(b.astype(np.float64)))
m = o.mode_params(0)
nb = int(sys.argv[1]) if len(sys.argv) > 1 else 468
for seed in range(65, 69):
    iq = synth.make_stream(0, nb * m.block_size // 2, seed)
    a, b = o.pipeline(0, 1, iq), f.pipeline(0, 1, iq)
    first = np.nonzero(np.abs(a["nco"] - b["nco"]) > 1e-3)[0][:1]
    print("seed %d: IF %.1e pilot %.1e nco %.1e audio %.1e L-R %.1e  max PCM diff %d LSB  nco diverges at IF sample %s"
          % (seed, rl2(b["if"], a["if"]), rl2(b["pilot"], a["pilot"]), rl2(b["nco"], a["nco"]), rl2(b["audio"], a["audio"]),
             rl2(b["audio"][0::2] - b["audio"][1::2], a["audio"][0::2] - a["audio"][1::2]),
             int(np.abs(a["pcm"].astype(int) - b["pcm"]).max()), first))
