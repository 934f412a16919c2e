"""How often does a stream's PLL trajectory leave the reference's?  (development / evidence tool, one GPU)

The PLL is chaotic in the last bit, and dy4_pllmath.h is not glibc: where a double result lies within a
double-ulp or so of a float rounding boundary the two may narrow to different floats.  This run processes N streams x 1.024 s
(mode 0 stereo, table-driven PLL), reads the per-stream census of such near-tie evaluations (dy4_pipeline_pll_risk), and
compares EVERY flagged stream plus a random sample of unflagged ones against the CPU checker (the reference's own code when
oracle/_ref was built): a stream has diverged if its PCM differs from the reference's by more than 1 LSB anywhere.

  python tools/divergence_run.py [n_streams=16384] [blocks=48] > gpurun_out/divergence.json
"""
import json
import multiprocessing as mp
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import dy4_b200
import oracle

_JOB = {}


def _cpu_one(i):
    chk = oracle.load(_JOB["kind"])
    return chk.pipeline(0, 1, _JOB["iq"][i], want=("pcm",))["pcm"]


def main():
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
    nb = int(sys.argv[2]) if len(sys.argv) > 2 else 48
    batch = 4096
    m = dy4_b200.mode_params(0)
    kind = "ref" if oracle.have_ref() else "oracle"
    rng = np.random.default_rng(2)
    tot = dict(streams=0, flagged=0, near_tie_evaluations=0, checked=0, checked_flagged=0, diverged=0, diverged_flagged=0, worst_lsb=0)
    rows = []
    t0 = time.time()
    for b0 in range(0, N, batch):
        S = min(batch, N - b0)
        d = dy4_b200.synth.make_batch_torch(0, S, nb * m.block_size // 2, base_seed=100000 + b0, device="cuda")
        p = dy4_b200.Pipeline(0, 1, S)
        pcm = p.process(d, n_blocks=nb, want=("pcm",))["pcm"]
        torch.cuda.synchronize()
        risk = p.pll_risk()
        p.close()
        flagged = np.nonzero(risk)[0]
        others = rng.choice(np.setdiff1d(np.arange(S), flagged), size=min(32, S - len(flagged)), replace=False)
        idx = np.concatenate([flagged, others]).astype(np.int64)
        _JOB.update(kind=kind, iq=d[torch.from_numpy(idx).cuda()].cpu().numpy())
        with mp.get_context("fork").Pool(os.cpu_count() or 1) as pool:
            refs = pool.map(_cpu_one, range(len(idx)))
        got = pcm[torch.from_numpy(idx).cuda()].cpu().numpy()
        for j, s in enumerate(idx):
            diff = int(np.abs(got[j].astype(np.int64) - refs[j].astype(np.int64)).max())
            is_f = j < len(flagged)
            tot["checked"] += 1
            tot["checked_flagged"] += int(is_f)
            tot["worst_lsb"] = max(tot["worst_lsb"], diff)
            if diff > 1:
                tot["diverged"] += 1
                tot["diverged_flagged"] += int(is_f)
                rows.append({"stream": int(b0 + s), "near_tie": int(risk[s]), "pcm_max_abs_diff": diff})
        tot["streams"] += S
        tot["flagged"] += int(len(flagged))
        tot["near_tie_evaluations"] += int(risk.sum())
        del d, pcm
        torch.cuda.empty_cache()
    per_stream = nb * m.if_per_block * 2 * 3
    out = {
        "workload": "mode 0 stereo, %d streams x %d blocks (%.3f s each), table-driven PLL" % (N, nb, nb * 0.0213333),
        "checker": kind, **tot,
        "evaluations_per_stream": per_stream,
        "near_tie_rate_per_evaluation": tot["near_tie_evaluations"] / (per_stream * tot["streams"]),
        "diverged_streams": rows, "seconds": round(time.time() - t0, 1),
        "note": "flagged = streams whose census is non-zero; every flagged stream and 32 unflagged ones per 4096 were compared; "
                "diverged = PCM more than 1 LSB from the reference's anywhere in the stream",
    }
    print(json.dumps(out))


if __name__ == "__main__":
    main()
