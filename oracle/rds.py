"""CPU restatement of the reference's RDS path — TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

The reference implements RDS only in its Python model, in float64 (`model/fmMonoBlock.py`, `model/fmSupportLib.py`,
`model/fmRRC.py`; all citations below are into /root/reference/model/).  This module restates that algorithm with
numpy (vectorised FIRs, a plain loop for the PLL and the bit-level state machines) so it finishes in seconds on whole
streams.  It is PINNED against the model itself: `tests/golden/make_golden_rds.py` imports the model's own functions
in the build container and commits their outputs (`tests/golden/rds_mode0.npz`); `tests/test_oracle.py` checks this
restatement against that fixture (filter chain to 1e-9, symbols / bits / frame-sync events exactly).

Two halves, joined at the RRC-filtered baseband:
  rds_front(if_signal)      fmMonoBlock.py:673-696   IF (240 kS/s) -> RRC I/Q at 38 kS/s
  RdsBackEnd.push(i, q)     fmMonoBlock.py:699-730   one model block (3 040 samples) -> symbols, bits, frame-sync events
"""
import math

import numpy as np

IF_FS = 240e3
NTAPS = 101                      # rf_taps, fmMonoBlock.py:48
RDS_UP, RDS_DOWN = 19, 120       # fmMonoBlock.py:61-62
RDS_TAPS = NTAPS * RDS_UP        # :63-64
RDS_FC = 3e3                     # :65
SPS = 16                         # :66
RDS_FS = SPS * 2375              # :67
BLOCK_IF = SPS * RDS_DOWN * 10   # IF samples of one model block: block_size/2/rf_decim, fmMonoBlock.py:568
BLOCK_RDS = BLOCK_IF * RDS_UP // RDS_DOWN   # 3 040 RRC samples = 190 symbols


# ---------------------------------------------------------------- taps
def firwin(numtaps, cutoff, window="hamming", pass_zero=True):
    """scipy.signal.firwin as the model calls it (fmMonoBlock.py:489-499,514): windowed ideal response, unit gain at
    DC (low-pass) or at the band centre (band-pass).  `cutoff` normalised to Nyquist; a pair means band-pass."""
    if np.isscalar(cutoff):
        left, right = 0.0, float(cutoff)
        assert pass_zero
    else:
        left, right = float(cutoff[0]), float(cutoff[1])
        assert not pass_zero
    alpha = 0.5 * (numtaps - 1)
    m = np.arange(numtaps) - alpha
    h = right * np.sinc(right * m) - left * np.sinc(left * m)
    n = np.arange(numtaps)
    if window == "hann":
        w = 0.5 - 0.5 * np.cos(2 * np.pi * n / (numtaps - 1))
    else:
        w = 0.54 - 0.46 * np.cos(2 * np.pi * n / (numtaps - 1))
    h = h * w
    scale_frequency = 0.0 if left == 0.0 else 0.5 * (left + right)
    return h / np.sum(h * np.cos(np.pi * m * scale_frequency))


def rrc_taps(Fs=RDS_FS, n_taps=NTAPS):
    """fmRRC.py:13-49."""
    T, beta = 1 / 2375.0, 0.90
    h = np.empty(n_taps)
    for k in range(n_taps):
        t = float((k - n_taps / 2)) / Fs
        if t == 0.0:
            h[k] = 1.0 + beta * ((4 / math.pi) - 1)
        elif t == -T / (4 * beta) or t == T / (4 * beta):
            h[k] = (beta / np.sqrt(2)) * (((1 + 2 / math.pi) * (math.sin(math.pi / (4 * beta)))) +
                                          ((1 - 2 / math.pi) * (math.cos(math.pi / (4 * beta)))))
        else:
            h[k] = (math.sin(math.pi * t * (1 - beta) / T) + 4 * beta * (t / T) * math.cos(math.pi * t * (1 + beta) / T)) / \
                   (math.pi * t * (1 - (4 * beta * t / T) * (4 * beta * t / T)) / T)
    return h


def model_taps():
    nyq = IF_FS / 2
    return dict(
        rds=firwin(NTAPS, [54e3 / nyq, 60e3 / nyq], "hann", pass_zero=False),            # fmMonoBlock.py:489-491
        carrier=firwin(NTAPS, [113.5e3 / nyq, 114.5e3 / nyq], "hann", pass_zero=False),  # :497-499
        lpf=firwin(RDS_TAPS, RDS_FC / (IF_FS * RDS_UP / 2)) * RDS_UP,                    # :514-515 (scipy's default window)
        rrc=rrc_taps())                                                                  # :690


# ---------------------------------------------------------------- filter chain
def convolve(x, h, state=None):
    """fmMonoBlock.py:320-329 — y[n] = sum_k h[k] x~[n-k], x~ = state || x."""
    L = len(h) - 1
    ext = np.concatenate((np.zeros(L) if state is None else state, x))
    return np.convolve(ext, h)[L:L + len(x)]


def resampler(up, down, x, h, state=None):
    """fmMonoBlock.py:331-341 — y[m] = sum_j h[phase + up j] x~[(m down)//up - j], phase = (m down) % up."""
    per = len(h) // up
    L = per - 1
    ext = np.concatenate((np.zeros(L) if state is None else state, x))
    n = np.arange(0, len(x) * up, down)
    base, phase = n // up, n % up
    hp = np.asarray(h)[:per * up].reshape(per, up).T          # hp[phase, j] = h[phase + up j]
    idx = base[:, None] - np.arange(per)[None, :] + L
    return np.einsum("mj,mj->m", hp[phase], ext[idx])


def fm_pll(x, freq, Fs, nco_scale, phase_adjust, bw, st):
    """fmMonoBlock.py:346-381.  st: dict integrator, phaseEst, feedbackI, feedbackQ, ncoState, q_ncoState, trigOffset."""
    Kp, Ki = bw * 2.666, (bw * bw) * 3.555
    n = len(x)
    nco, qnco = np.empty(n + 1), np.empty(n + 1)
    nco[0], qnco[0] = st["ncoState"], st["q_ncoState"]
    integ, phase, fbI, fbQ, off = st["integrator"], st["phaseEst"], st["feedbackI"], st["feedbackQ"], st["trigOffset"]
    w = 2 * math.pi * (freq / Fs)
    for k in range(n):
        eI = x[k] * (+fbI)
        eQ = x[k] * (-fbQ)
        eD = 0 if eI == 0 else math.atan2(eQ, eI)            # :359-362
        integ = integ + Ki * eD
        phase = phase + Kp * eD + integ
        off += 1
        arg = w * off + phase
        fbI, fbQ = math.cos(arg), math.sin(arg)
        nco[k + 1] = math.cos(arg * nco_scale + phase_adjust)
        qnco[k + 1] = math.sin(arg * nco_scale + phase_adjust)
    st.update(integrator=integ, phaseEst=phase, feedbackI=fbI, feedbackQ=fbQ, trigOffset=off,
              ncoState=nco[n], q_ncoState=qnco[n])
    return nco[:-1], qnco[:-1]


def new_pll_state():
    return dict(integrator=0.0, phaseEst=0.0, feedbackI=1.0, feedbackQ=0.0, ncoState=1.0, q_ncoState=1.0, trigOffset=0)  # :450-459


def rds_front(if_signal, taps=None):
    """fmMonoBlock.py:673-696 over a whole stream (every stage is independent of the block partition)."""
    t = taps or model_taps()
    x = np.asarray(if_signal, np.float64)
    rds_f = convolve(x, t["rds"])                                       # :675
    carrier = convolve(rds_f * rds_f, t["carrier"])                     # :678-679
    d = NTAPS // 2                                                      # RDS_delay_state: int(rf_taps/2), :503
    delayed = np.concatenate((np.zeros(d), rds_f[:-d]))                 # :682
    nco_i, nco_q = fm_pll(carrier, 114e3, IF_FS, 0.5, 0, 0.001, new_pll_state())   # :685
    lp_i = resampler(RDS_UP, RDS_DOWN, nco_i * delayed, t["lpf"])       # :686-689
    lp_q = resampler(RDS_UP, RDS_DOWN, nco_q * delayed, t["lpf"])       # :694-695
    return dict(rds_f=rds_f, carrier=carrier, nco_i=nco_i, nco_q=nco_q,
                rrc_i=convolve(lp_i, t["rrc"]), rrc_q=convolve(lp_q, t["rrc"]))   # :691,696


# ---------------------------------------------------------------- back half
def manchester_encoded(sig, qsig, sps, K, found):
    """fmSupportLib.py:209-247, quirks included (`max = signal[i]` keeps the sign; `abs(output[..] < threshold)`)."""
    idx = K
    threshold = 0.05
    truncate = False
    ns = int(len(sig) / sps)
    out, qout = np.zeros(ns), np.zeros(ns)
    sym = np.zeros(ns, np.int16)
    if not found:
        mx = 0
        idx = 0
        for i in range(sps * 2):
            if abs(sig[i]) > mx:
                mx = sig[i]
                idx = i
                found = True
                truncate = True
    k = 0
    for k in range(idx, len(sig), sps):
        out[int(k / sps)] = sig[k]
        qout[int(k / sps)] = qsig[k]
        sym[int(k / sps)] = 0 if (sig[k] < 0) else 1
    if abs(out[int(k / sps)]) < threshold and abs(out[int((k - 1) / sps)] < threshold):
        found = False
    k = k % sps
    if truncate:
        out, qout, sym = out[1:], qout[1:], sym[1:]
    return out, qout, sym, k, found


SYNDROMES = {  # fmMonoBlock.py:196,213,230,247,264
    (1, 1, 1, 1, 0, 1, 1, 0, 0, 0): "A", (1, 1, 1, 1, 0, 1, 0, 1, 0, 0): "B", (1, 0, 0, 1, 0, 1, 1, 1, 0, 0): "C",
    (1, 1, 1, 1, 0, 0, 1, 1, 0, 0): "Cp", (1, 0, 0, 1, 0, 1, 1, 0, 0, 0): "D"}
PREDECESSORS = {"A": ("D",), "B": ("A",), "C": ("B",), "Cp": ("B",), "D": ("C", "Cp")}
TYPE_CODE = {"A": 0, "B": 1, "C": 2, "Cp": 3, "D": 4}
PARITY_ROWS = [  # the indices summed for each syndrome bit, fmMonoBlock.py:183-192
    (0, 10, 13, 14, 15, 16, 17, 19, 20, 23, 24, 25), (1, 11, 14, 15, 16, 17, 18, 20, 21, 24, 25),
    (2, 10, 12, 13, 14, 18, 20, 21, 22, 23, 24), (3, 10, 11, 16, 17, 20, 21, 22), (4, 11, 12, 17, 18, 21, 22, 23),
    (5, 10, 12, 14, 15, 16, 17, 18, 20, 22, 25), (6, 10, 11, 14, 18, 20, 21, 24, 25),
    (7, 10, 11, 12, 13, 14, 16, 17, 20, 21, 22, 23, 24), (8, 11, 12, 13, 14, 15, 17, 18, 21, 22, 23, 24, 25),
    (9, 12, 13, 14, 15, 16, 18, 19, 22, 23, 24, 25)]


PTY_NAMES = (  # RDS programme types by 5-bit code, as the model's table lists them (RDS_Application_layer.py:11-44)
    "No programme type or undefined", "News", "Current Affairs", "Information", "Sport", "Education", "Drama", "Culture",
    "Science", "Varied", "Pop Music", "Rock Music", "Easy Listening Music", "Light classical", "Serious classical",
    "Other Music", "Weather", "Finance", "Children's programmes", "Social Affairs", "Religion", "Phone In", "Travel",
    "Leisure", "Jazz Music", "Country Music", "National Music", "Oldies Music", "Folk Music", "Documentary", "Alarm Test",
    "Alarm")


def process_rds_data(a, b, c, d, prev_pty, prev_pi, count):
    """RDS_Application_layer.py:1-177 on one complete group (four 16-bit words, lists of bits).  Returns the lines the
    model prints and its (PTYcode, PIcode, count) hand-off.  Kept as the model behaves: the service-name segment array is
    local to the call, and the character table is keyed 'xxxx xxxx' (with a space) while the coded strings have none, so
    no character ever matches — the printed service name is empty."""
    lines = ["A block " + str(a), "B Block " + str(b), "C Block " + str(c), "D Block " + str(d)]
    pi = "".join(hex(sum(a[i + j] * (2 ** (3 - j)) for j in range(4)))[2:].upper() for i in range(0, 16, 4))   # :123-129
    pty = "".join(map(str, b[6:11]))                                                                           # :135
    psns = [""] * 8
    if "".join(map(str, d[0:5])) == "00000":                                                                   # :143-161
        idx = {"00": 0, "01": 2, "10": 4, "11": 6}["".join(map(str, d[0:2]))]
        psns[idx] = ""
        psns[idx + 1] = ""
        count += 1
    if count == 4:                                                                                             # :166-169
        lines.append("Program service: " + "".join(psns))
        count = 0
    if pty != prev_pty and pi != prev_pi and pty != "" and pi != []:                                           # :172-175
        lines.append("PI code: " + pi)
        lines.append("Program type: " + PTY_NAMES[int(pty, 2)])
    return lines, pty, pi, count


class RdsBackEnd:
    """The per-block glue of fmMonoBlock.py:699-730 with its state (:576-601)."""

    def __init__(self):
        self.block_count = 0
        self.m_index, self.m_found = 0, False
        self.symbol_state, self.errors1, self.errors2 = 0, 0, 0
        self.bit_state = 0
        self.window_index, self.synced, self.window_state = 24, False, []
        self.offset_state, self.num_synced, self.bit_pos, self.last_pos = "", 0, 0, 0
        self.symbols, self.bits, self.events = [], [], []     # everything produced so far
        self.msgs = {"A": [], "B": [], "C": [], "D": []}      # :596-600
        self.pty, self.pi, self.count = "", "", 0             # :523-525
        self.groups, self.app_lines = [], []                  # complete groups handed to the application layer; what it printed

    def push(self, rrc_i, rrc_q):
        assert len(rrc_i) == BLOCK_RDS
        _, _, sym, self.m_index, self.m_found = manchester_encoded(rrc_i, rrc_q, SPS, self.m_index, self.m_found)   # :699
        sym = [int(s) for s in sym]
        self.symbols.append(sym)
        if self.block_count >= 5:
            if self.block_count < 10:
                self._find_pattern(sym)                                   # :703-704
            else:
                start = 0 if self.errors1 > self.errors2 else 1          # :706
                bits = self._decode(sym, start)                           # :708
                self.bits.extend(bits)
                widx = 0
                while (self.synced and widx < len(bits) - 26) or (not self.synced and widx < len(bits) - 1):   # :711
                    window = self._get_window(bits)
                    widx = self.window_index
                    msg = self._frame_sync(window)
                    if self.synced:                                       # :718-722 ('Cp' never lands in msgs.c, as in the model)
                        if self.offset_state in self.msgs:
                            self.msgs[self.offset_state] = msg
                    else:                                                 # :723-727
                        self.msgs = {"A": [], "B": [], "C": [], "D": []}
                    if all(self.msgs[k] != [] for k in "ABCD"):           # :729-730
                        m = self.msgs
                        self.groups.append(tuple(int("".join(map(str, m[k])), 2) for k in "ABCD"))
                        lines, self.pty, self.pi, self.count = process_rds_data(m["A"], m["B"], m["C"], m["D"], self.pty, self.pi, self.count)
                        self.app_lines.extend(lines)
        self.block_count += 1

    def _find_pattern(self, s):                                           # :78-92
        for i in range(1, len(s), 2):
            c1, p1 = s[i], s[i - 1]
            c2 = s[i - 1]
            p2 = s[i - 2] if i != 1 else self.symbol_state
            if c1 == p1:
                self.errors1 += 1
            if c2 == p2:
                self.errors2 += 1
        self.symbol_state = s[-1]

    def _decode(self, s, start):                                          # :97-122
        out = []
        for i in range(start, len(s), 2):
            cur = s[i]
            prev = s[i - 1] if i != 0 else self.symbol_state
            b = 1 if (cur == 0 and prev == 1) else 0
            out.append(1 if b != self.bit_state else 0)
            self.bit_state = b
        self.symbol_state = s[-1]
        return out

    def _get_window(self, data):                                          # :157-173
        self.window_index += 26 if self.synced else 1
        if self.window_index >= len(data):
            self.window_index -= len(data)
        i = self.window_index
        window = self.window_state[i:] + data[:i + 1] if i < 25 else data[i - 25:i + 1]
        self.window_state = data[len(data) - 25:]
        return window

    def _frame_sync(self, m):                                             # :176-284
        s = tuple(sum(m[j] for j in row) % 2 for row in PARITY_ROWS)
        t = SYNDROMES.get(s)
        msg_bits = list(m[0:16]) if t is not None else []
        if t is not None:
            if self.offset_state in PREDECESSORS[t] or (self.offset_state == "" and not self.synced):
                self.synced = True        # numSynced += 1 if not synced else 0  -> adds 0 (synced was just set)
            elif self.synced:
                self.synced = False
                self.num_synced = 0
            false_pos = self.bit_pos != self.last_pos + 26 and self.offset_state != ""
            msg = 0
            for b in m[0:16]:
                msg = (msg << 1) | b
            self.events.append((TYPE_CODE[t], self.bit_pos, int(false_pos), msg))
            self.offset_state = t if self.synced else ""
            self.last_pos = self.bit_pos
        self.bit_pos += 26 if self.synced else 1
        if self.num_synced > 3 and not self.synced:
            self.synced = True
        return msg_bits


def rds_back(rrc_i, rrc_q):
    be = RdsBackEnd()
    for b in range(len(rrc_i) // BLOCK_RDS):
        be.push(rrc_i[b * BLOCK_RDS:(b + 1) * BLOCK_RDS], rrc_q[b * BLOCK_RDS:(b + 1) * BLOCK_RDS])
    return be
