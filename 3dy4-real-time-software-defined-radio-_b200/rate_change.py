"""fmRateChange — the reference's fixture re-rater (model/fmRateChange.py:43-66) as a host utility.

    python -m dy4_b200.rate_change <inputFile> [outFsID=1] [inFsID=0]

Reads raw interleaved 8-bit unsigned I/Q, resamples I and Q from sample_rate_table[inFsID] to sample_rate_table[outFsID]
kS/s with a polyphase FIR and writes <input>_<rate>.raw.  The reference calls scipy.signal.resample_poly; this module
restates that algorithm in numpy (scipy's published design: a Kaiser(5.0)-windowed sinc of 2*10*max(up, down) + 1 taps with
cut-off 1/max(up, down) and gain `up`, zero-phase trimming), so the tool carries no scipy dependency; tests/test_host_logic.py
holds it to scipy's own output.  Not on the hot path: files are prepared once, off line.
"""
import math
import sys

import numpy as np

SAMPLE_RATE_TABLE = [2400, 2880, 2304, 1920, 1440, 1152, 960]          # fmRateChange.py:14, kS/s


def _design(up, down):
    """scipy.signal.resample_poly's default filter: firwin(2 * half_len + 1, 1 / max_rate, window=('kaiser', 5.0)) * up."""
    max_rate = max(up, down)
    half_len = 10 * max_rate
    n = 2 * half_len + 1
    m = np.arange(n) - half_len
    fc = 1.0 / max_rate
    h = fc * np.sinc(fc * m)                                           # firwin: ideal low-pass, cut-off in units of Nyquist
    beta = 5.0
    win = np.i0(beta * np.sqrt(1 - (2.0 * np.arange(n) / (n - 1) - 1.0) ** 2)) / np.i0(beta)
    h = h * win
    h /= h.sum()                                                       # unity gain at DC (firwin scale=True)
    return h * up, half_len


def _out_len(len_h, n_in, up, down):
    return ((n_in - 1) * up + len_h - 1) // down + 1                   # scipy.signal._upfirdn._output_len


def resample_poly(x, up, down):
    """y = resample_poly(x, up, down) as scipy defines it (axis 0, zero padding), float64."""
    x = np.asarray(x, np.float64)
    up, down = int(up), int(down)
    g = math.gcd(up, down)
    up //= g
    down //= g
    if up == down == 1:
        return x.copy()
    n_in = x.size
    n_out = n_in * up // down + bool(n_in * up % down)
    h, half_len = _design(up, down)
    n_pre_pad = down - half_len % down
    n_post_pad = 0
    n_pre_remove = (half_len + n_pre_pad) // down
    while _out_len(h.size + n_pre_pad + n_post_pad, n_in, up, down) < n_out + n_pre_remove:
        n_post_pad += 1
    h = np.concatenate([np.zeros(n_pre_pad), h, np.zeros(n_post_pad)])
    # upfirdn: zero-stuff by `up`, FIR, keep every `down`-th — phase by phase, so nothing is multiplied by a stuffed zero
    n_full = _out_len(h.size, n_in, up, down)
    y = np.zeros(n_full)
    t = np.arange(n_full) * down                                       # index into the up-sampled stream
    for ph in range(up):                                               # taps h[ph::up] meet input samples
        sel = np.nonzero(t % up == ph)[0]
        if sel.size == 0:
            continue
        hp = h[ph::up]
        full = np.convolve(x, hp)                                      # full[j] = sum_k hp[k] x[j - k]
        j = t[sel] // up
        ok = j < full.size
        y[sel[ok]] = full[j[ok]]
    return y[n_pre_remove:n_pre_remove + n_out]


def rate_change(raw_u8, out_fs_id=1, in_fs_id=0):
    """uint8 interleaved I/Q at sample_rate_table[in_fs_id] -> the same at sample_rate_table[out_fs_id] (fmRateChange.py:43-66)."""
    raw = np.asarray(raw_u8, np.uint8)
    fs_in, fs_out = SAMPLE_RATE_TABLE[in_fs_id] * 1e3, SAMPLE_RATE_TABLE[out_fs_id] * 1e3
    iq = (raw - 128.0) / 128.0                                          # :46
    g = math.gcd(int(fs_in), int(fs_out))
    expand, decim = int(fs_out) // g, int(fs_in) // g                   # :49-50
    ri = resample_poly(iq[0::2], expand, decim)                         # :52-53
    rq = resample_poly(iq[1::2], expand, decim)
    out = np.empty(2 * ri.size, np.uint8)
    out[0::2] = (128 + np.trunc(ri * 127)).astype(np.int64).astype(np.uint8)    # :57-59: 128 + int(x * 127), wrapped to uint8
    out[1::2] = (128 + np.trunc(rq * 127)).astype(np.int64).astype(np.uint8)
    return out


def main(argv):
    if len(argv) < 2:
        print("usage: python -m dy4_b200.rate_change <inputFile> [outFsID=1] [inFsID=0]\nsample rate IDs: " +
              ", ".join("%d - %g MS/s" % (i, r / 1e3) for i, r in enumerate(SAMPLE_RATE_TABLE)))
        return 1
    in_fname = argv[1]
    out_id = int(argv[2]) if len(argv) > 2 else 1
    in_id = int(argv[3]) if len(argv) > 3 else 0
    out_fname = in_fname.partition(".")[0] + "_" + str(SAMPLE_RATE_TABLE[out_id]) + ".raw"      # :40
    rate_change(np.fromfile(in_fname, dtype=np.uint8), out_id, in_id).tofile(out_fname)
    print('Written resampled RF data to "%s" in unsigned 8-bit format' % out_fname)
    return 0


if __name__ == "__main__":
    sys.exit(main(sys.argv))
