"""Ad-hoc GPU parity check (development aid): CUDA pipeline vs the CPU oracle, all modes."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import dy4_b200, oracle

def rl2(a, b):
    a = a.astype(np.float64); b = b.astype(np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))

o = oracle.load("ref") if oracle.have_ref() else oracle.load("oracle")
print("checker:", o.kind)
nb = int(os.environ.get("NB", "6")); S = int(os.environ.get("S", "3"))
for mode in range(4):
    m = dy4_b200.mode_params(mode)
    iq = dy4_b200.synth.make_batch(mode, S, nb * m.block_size // 2, base_seed=65 + 10 * mode)
    d_iq = torch.from_numpy(iq).cuda()
    for stereo in (0, 1):
        for exact in (0, 1):
            p = dy4_b200.Pipeline(mode, stereo, S, exact_audio=bool(exact), debug_rows=True)
            t = time.time()
            out = p.process(d_iq, want=("pcm", "audio", "if"))
            torch.cuda.synchronize()
            dt = time.time() - t
            res = []
            for s in range(S):
                ref = o.pipeline(mode, stereo, iq[s])
                g_if = out["if"][s].cpu().numpy(); g_a = out["audio"][s].cpu().numpy(); g_p = out["pcm"][s].cpu().numpy()
                res.append((int((g_if.view(np.uint32) != ref["if"].view(np.uint32)).sum()), rl2(g_a, ref["audio"]),
                            int((g_a.view(np.uint32) != ref["audio"].view(np.uint32)).sum()),
                            int(np.abs(g_p.astype(int) - ref["pcm"]).max())))
            extra = ""
            if stereo:
                pil, nco = p.debug_pilot_nco()
                ref = o.pipeline(mode, 1, iq[S - 1])
                extra = " pilot!= %d nco!= %d" % (int((pil[S-1].cpu().numpy().view(np.uint32) != ref["pilot"].view(np.uint32)).sum()),
                                                   int((nco[S-1].cpu().numpy().view(np.uint32) != ref["nco"].view(np.uint32)).sum()))
            print("mode %d %s exact_audio=%d  %.3fs  [IF bit-mismatches, audio relL2, audio bit-mismatches, pcm max|diff|] per stream: %s%s"
                  % (mode, "stereo" if stereo else "mono", exact, dt, res, extra), flush=True)
            p.close()
# chunked processing must equal one-shot processing (state carried in tails)
mode = 0; m = dy4_b200.mode_params(mode)
iq = dy4_b200.synth.make_batch(mode, 2, 8 * m.block_size // 2, base_seed=99)
d_iq = torch.from_numpy(iq).cuda()
p = dy4_b200.Pipeline(mode, 1, 2, exact_audio=True)
one = p.process(d_iq, want=("pcm", "audio", "if"))
p.reset()
parts = [p.process(d_iq[:, b * m.block_size:(b + nbk) * m.block_size].contiguous(), want=("pcm", "audio", "if")) for b, nbk in ((0, 1), (1, 3), (4, 4))]
for k in one:
    cat = torch.cat([x[k] for x in parts], dim=1)
    print("chunked == one-shot", k, bool(torch.equal(cat, one[k])))
# host path
hp = p.process_host(iq, want=("pcm", "audio"), chunk_blocks=3) if (p.reset() or True) else None
print("host path == device path pcm", np.array_equal(hp["pcm"], one["pcm"].cpu().numpy()), "audio", np.array_equal(hp["audio"], one["audio"].cpu().numpy()))
print("launches:", dy4_b200.launch_count())
