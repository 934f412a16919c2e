"""Mint the RDS golden fixture from the reference's PYTHON MODEL (the only RDS implementation it has).

Run in the build container:  python tests/golden/make_golden_rds.py        (about two minutes: the model is pure Python)
The model (model/fmMonoBlock.py) keeps its main loop under `if __name__ == "__main__"` and imports matplotlib, which
is not installed here; an empty stub package on sys.path lets its functions be imported unchanged.  Per SURVEY.md §8c
the Python and C++ front ends differ (float64, firwin taps, atan2 demod), so the oracles are joined at the IF:
the IF comes from the C oracle (bit-identical to the CUDA path), and the MODEL'S OWN FUNCTIONS — convolve,
squaringNonlinearity, delayBlock, fmPll, pointwiseMultiply, resampler, impulseResponseRootRaisedCosine,
manchesterEncoded, find_pattern, decode, get_window, frame_sync_receiver, process_rds_data — run on it in float64 with the model's own
taps, parameters and block size (fmMonoBlock.py:444-447, 488-515, 568, 673-730).  Only the glue of the model's main loop
(which cannot be imported) is restated here, line for line.
"""
import contextlib
import hashlib
import io
import os
import re
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
stub = tempfile.mkdtemp()
os.makedirs(os.path.join(stub, "matplotlib"))
open(os.path.join(stub, "matplotlib", "__init__.py"), "w").close()
open(os.path.join(stub, "matplotlib", "pyplot.py"), "w").close()
sys.path.insert(0, stub)
sys.path.insert(0, "/root/reference/model")
from scipy import signal  # noqa: E402
import fmMonoBlock as M  # noqa: E402
from fmRRC import impulseResponseRootRaisedCosine  # noqa: E402
from fmSupportLib import manchesterEncoded  # noqa: E402
from RDS_Application_layer import process_rds_data  # noqa: E402
import oracle  # noqa: E402
import importlib.util  # noqa: E402

_s = importlib.util.spec_from_file_location("synth", os.path.join(ROOT, "3dy4-real-time-software-defined-radio-_b200", "synth.py"))
synth = importlib.util.module_from_spec(_s)
_s.loader.exec_module(synth)

HERE = os.path.dirname(os.path.abspath(__file__))
SEED, NB = 7065, 60                      # 60 mode-0 blocks = 307 200 IF samples = 16 model blocks of 19 200
TYPE_CODE = {"A": 0, "B": 1, "C": 2, "Cp": 3, "D": 4}


def main():
    o = oracle.load("ref") if oracle.have_ref() else oracle.load("oracle")
    m = o.mode_params(0)
    iq = synth.make_stream(0, NB * m.block_size // 2, SEED, rds=True)
    fm = o.pipeline(0, 1, iq, want=("if",))["if"].astype(np.float64)
    Fs, taps = 240e3, M.rf_taps
    rds_coeff = signal.firwin(taps, [54e3 / (Fs / 2), 60e3 / (Fs / 2)], window=('hann'), pass_zero=False)
    car_coeff = signal.firwin(taps, [113.5e3 / (Fs / 2), 114.5e3 / (Fs / 2)], window=('hann'), pass_zero=False)
    lpf = signal.firwin(M.RDS_taps, M.RDS_Fc / (Fs * M.RDS_upsample / 2)) * M.RDS_upsample
    st = M.EmptyObject()
    st.integrator = 0.0; st.phaseEst = 0.0; st.feedbackI = 1.0; st.feedbackQ = 0.0; st.ncoState = 1.0; st.trigOffset = 0; st.q_ncoState = 1.0
    s_rds, s_car, s_del = np.zeros(taps - 1), np.zeros(taps - 1), np.zeros(int(taps / 2))
    s_lp, s_lpq = np.zeros(M.RDS_taps - 1), np.zeros(M.RDS_taps - 1)
    s_rrc, s_rrcq = np.zeros(int(M.RDS_taps / M.RDS_upsample) - 1), np.zeros(int(M.RDS_taps / M.RDS_upsample) - 1)
    # back-half state, fmMonoBlock.py:528-530, 576-594
    m_index, m_found = 0.0, False
    window_index, synced, window_state = 24, False, []
    symbol_state, errors1, errors2, bit_state = 0, 0, 0, 0
    offsetState, numSynced, bit_pos, last_pos = '', 0, 0, 0
    out = {k: [] for k in ("rds_f", "carrier", "nco_i", "nco_q", "rrc_i", "rrc_q")}
    symbols, sym_counts, bits, events = [], [], [], []
    msgs = M.EmptyObject(); msgs.a = []; msgs.b = []; msgs.c = []; msgs.d = []          # fmMonoBlock.py:596-600
    PTYcode, PIcode, count = '', '', 0                                                  # :523-525
    groups, app_lines = [], []
    blk = M.sps * M.RDS_decim * M.rf_decim * M.audio_decim * 2 * 2 // 2 // M.rf_decim     # block_size (:568) in IF samples
    assert blk == 19200 and len(fm) % blk == 0
    for block_count in range(len(fm) // blk):
        x = fm[block_count * blk:(block_count + 1) * blk]
        rds_f, s_rds = M.convolve(x, rds_coeff, s_rds)                                    # :675
        sq = np.array(M.squaringNonlinearity(rds_f))                                      # :678
        car, s_car = M.convolve(sq, car_coeff, s_car)                                     # :679
        dly, s_del = M.delayBlock(rds_f, s_del)                                           # :682
        nco_i, nco_q = M.fmPll(car, 114e3, Fs, 0.5, 0, 0.001, st)                         # :685
        mix_i = M.pointwiseMultiply(nco_i, dly, 1)                                        # :686
        lp_i, s_lp = M.resampler(M.RDS_upsample, M.RDS_decim, mix_i, lpf, s_lp)           # :689
        rrc = impulseResponseRootRaisedCosine(M.RDS_Fs, int(M.RDS_taps / M.RDS_upsample))  # :692
        rrc_i, s_rrc = M.convolve(lp_i, rrc, s_rrc)                                       # :693
        mix_q = M.pointwiseMultiply(nco_q, dly, 1)                                        # :696
        lp_q, s_lpq = M.resampler(M.RDS_upsample, M.RDS_decim, mix_q, lpf, s_lpq)         # :697
        rrc_q, s_rrcq = M.convolve(lp_q, rrc, s_rrcq)                                     # :698
        for k, v in (("rds_f", rds_f), ("carrier", car), ("nco_i", nco_i), ("nco_q", nco_q), ("rrc_i", rrc_i), ("rrc_q", rrc_q)):
            out[k].append(np.array(v, np.float64))
        log = io.StringIO()
        with contextlib.redirect_stdout(log):
            _, _, RDS_symbols, m_index, m_found = manchesterEncoded(rrc_i, rrc_q, M.sps, m_index, m_found)   # :701
            symbols.append(np.array(RDS_symbols, np.int8)); sym_counts.append(len(RDS_symbols))
            if block_count >= 5:                                                          # :703-730
                if block_count < 10:
                    symbol_state, errors1, errors2 = M.find_pattern(RDS_symbols, symbol_state, errors1, errors2)
                else:
                    decode_start = 0 if errors1 > errors2 else 1
                    decoded_stream, symbol_state, bit_state = M.decode(RDS_symbols, symbol_state, bit_state, decode_start)
                    bits.extend(int(b) for b in decoded_stream)
                    widx = 0
                    while ((synced and widx < len(decoded_stream) - 26) or (not synced and widx < len(decoded_stream) - 1)):
                        window_data, window_index, window_state = M.get_window(decoded_stream, window_index, synced, window_state)
                        widx = window_index
                        mark = log.tell()
                        pos_before = bit_pos
                        synced, msg, offsetState, numSynced, bit_pos, last_pos = M.frame_sync_receiver(
                            window_data, synced, offsetState, numSynced, bit_pos, last_pos)
                        said = log.getvalue()[mark:]
                        found = re.search(r"Block type (\w+) found! Bit position\s+(\d+)", said)
                        if found:
                            assert int(found.group(2)) == pos_before and msg != []
                            word = 0
                            for b in msg:
                                word = (word << 1) | int(b)
                            events.append((TYPE_CODE[found.group(1)], pos_before, int("false positive" in said), word))
                        else:
                            assert msg == []
                        if synced:                                                        # :718-722
                            msgs.a = msg if offsetState == 'A' else msgs.a
                            msgs.b = msg if offsetState == 'B' else msgs.b
                            msgs.c = msg if offsetState == 'C' else msgs.c
                            msgs.d = msg if offsetState == 'D' else msgs.d
                        else:                                                             # :723-727
                            msgs.a = []; msgs.b = []; msgs.c = []; msgs.d = []
                        if msgs.a != [] and msgs.b != [] and msgs.c != [] and msgs.d != []:   # :729-730
                            groups.append([int("".join(str(int(b)) for b in w), 2) for w in (msgs.a, msgs.b, msgs.c, msgs.d)])
                            mark2 = log.tell()
                            PTYcode, PIcode, count = process_rds_data(msgs, PTYcode, PIcode, count)
                            app_lines.extend(log.getvalue()[mark2:].splitlines())
        print("model block", block_count, "symbols", sym_counts[-1], "bits so far", len(bits), "events", len(events),
              "errors1/2", errors1, errors2, "synced", synced, flush=True)
    out = {k: np.concatenate(v) for k, v in out.items()}
    n64 = 4 * 3040
    np.savez_compressed(os.path.join(HERE, "rds_mode0.npz"), seed=SEED, n_blocks=NB,
                        iq_sha256=hashlib.sha256(iq.tobytes()).hexdigest(),
                        rrc_i64=out["rrc_i"][:n64], rrc_q64=out["rrc_q"][:n64],
                        rrc_i=out["rrc_i"].astype(np.float32), rrc_q=out["rrc_q"].astype(np.float32),
                        rds_f_8=out["rds_f"][::8].astype(np.float32), carrier_8=out["carrier"][::8].astype(np.float32),
                        nco_i_8=out["nco_i"][::8].astype(np.float32), nco_q_8=out["nco_q"][::8].astype(np.float32),
                        symbols=np.concatenate(symbols), symbol_counts=np.array(sym_counts, np.int32),
                        bits=np.array(bits, np.int8), events=np.array(events, np.int32).reshape(-1, 4),
                        errors=np.array([errors1, errors2], np.int32),
                        groups=np.array(groups, np.int32).reshape(-1, 4), app_lines=np.array("\n".join(app_lines)))
    print({k: (v.shape, float(np.abs(v).max())) for k, v in out.items()})
    print("bits", len(bits), "events", events[:12])


if __name__ == "__main__":
    main()
