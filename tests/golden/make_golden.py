"""Mint the golden fixtures in tests/golden/ from the REFERENCE ITSELF.

Run in the build container (needs /root/reference):  python tests/golden/make_golden.py
It builds oracle/_ref (the reference's own filter.cpp/project.cpp/iofunc.cpp compiled in
place, see oracle/Makefile), runs it on seeded synthetic streams, cross-checks the PCM
against the unmodified `project` binary fed the same bytes on stdin, and stores:

  mode{M}_stereo.npz  iq (2 blocks, uint8) + if, pilot, nco, audio, pcm
  mode{M}_mono.npz    audio, pcm for the same iq
  mode0_stereo_long.npz  12 blocks (past the point where the PLL becomes chaotic in the last
                      bit, ~16.7k IF samples): input is regenerated from its seed (sha256 kept),
                      outputs pcm + audio + every 16th nco sample + sha256 of the full if/nco
  taps.npz            every impulse response project.cpp:262-273 generates, per mode

The reference ships no golden vectors for this path; these are its
outputs, so they pin both oracle/dy4_oracle.c and the CUDA path.
"""
import hashlib
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402
import importlib.util  # noqa: E402

_s = importlib.util.spec_from_file_location("synth", os.path.join(ROOT, "3dy4-real-time-software-defined-radio-_b200", "synth.py"))
synth = importlib.util.module_from_spec(_s)
_s.loader.exec_module(synth)

HERE = os.path.dirname(os.path.abspath(__file__))
SEED = 65


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def binary_pcm(mode, stereo, iq):
    p = subprocess.run([os.path.join(ROOT, "oracle", "_ref", "project"), str(mode), "stereo" if stereo else "mono"],
                       input=iq.tobytes(), capture_output=True)
    assert p.returncode == 1, p.returncode          # the reference exits 1 at end of input (project.cpp:293-296)
    return np.frombuffer(p.stdout, np.int16)


def main():
    oracle.build(ref=True)
    r = oracle.load("ref")
    taps = {}
    for mode in range(4):
        m = r.mode_params(mode)
        iq = synth.make_stream(mode, 2 * m.block_size // 2, SEED + mode)
        st = r.pipeline(mode, 1, iq)
        mo = r.pipeline(mode, 0, iq)
        assert np.array_equal(binary_pcm(mode, 1, iq), st["pcm"]) and np.array_equal(binary_pcm(mode, 0, iq), mo["pcm"])
        np.savez_compressed(os.path.join(HERE, "mode%d_stereo.npz" % mode), iq=iq, **{k: st[k] for k in ("if", "pilot", "nco", "audio", "pcm")})
        np.savez_compressed(os.path.join(HERE, "mode%d_mono.npz" % mode), audio=mo["audio"], pcm=mo["pcm"])
        taps["rf_%d" % mode] = r.lpf_taps(m.rf_Fs, 100e3, 101, 1)
        taps["audio_%d" % mode] = r.lpf_taps(m.if_Fs * m.audio_upsample, 16e3, m.audio_taps, m.audio_upsample)
        taps["pilot_%d" % mode] = r.bpf_taps(m.if_Fs, 18.5e3, 19.5e3, 101, 1)
        taps["stereo_%d" % mode] = r.bpf_taps(m.if_Fs, 22e3, 54e3, 101, 1)
    np.savez_compressed(os.path.join(HERE, "taps.npz"), **taps)

    m = r.mode_params(0)
    nb = 12
    iq = synth.make_stream(0, nb * m.block_size // 2, SEED + 100)
    st = r.pipeline(0, 1, iq)
    assert np.array_equal(binary_pcm(0, 1, iq), st["pcm"])
    np.savez_compressed(os.path.join(HERE, "mode0_stereo_long.npz"), seed=SEED + 100, n_blocks=nb, iq_sha256=sha(iq),
                        pcm=st["pcm"], audio=st["audio"], nco_16=st["nco"][::16].copy(),
                        if_sha256=sha(st["if"]), nco_sha256=sha(st["nco"]), pilot_sha256=sha(st["pilot"]))
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print("%-28s %8d bytes" % (f, os.path.getsize(os.path.join(HERE, f))))


if __name__ == "__main__":
    main()
