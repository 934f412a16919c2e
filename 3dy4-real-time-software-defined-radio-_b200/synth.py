"""Deterministic synthetic FM-broadcast IQ streams (8-bit interleaved I,Q).

Recipe of SURVEY.md §8(d): per stream s with base seed S, rng = PCG64(S+s);
left/right tones in [300, 5000] Hz with amplitudes in [0.2, 0.5]; multiplex
m = 0.45(L+R) + 0.1 sin(wp t + phi) + 0.45 (L-R) sin(2(wp t + phi)), wp = 2 pi 19 kHz;
FM with 75 kHz deviation; additive N(0, 0.01^2) noise on I and Q;
u8 = clip(round(100 x + 128), 0, 255).  With rds=True a 57 kHz sub-carrier (3x the pilot, amplitude 0.05) carries
valid RDS groups: differentially encoded, biphase (Manchester) symbols at 2375 symbols/s.  The format is what the reference reads
on stdin (rtl_sdr output, src/iofunc.cpp:113-120).

Not on the hot path: this only feeds tests and the benchmark.
"""
import numpy as np

RF_FS = {0: 2.4e6, 1: 1.44e6, 2: 2.4e6, 3: 1.92e6}  # reference src/project.cpp:180,192,204,216


def stream_params(seed, period_s=None):
    """period_s: make the stream periodic with that period (tone frequencies moved to the nearest multiple of 1/period_s; the
    19 kHz pilot and its harmonics must be multiples already): replaying the buffer then IS a continuous signal."""
    rng = np.random.Generator(np.random.PCG64(seed))
    p = dict(fL=rng.uniform(300, 5000), fR=rng.uniform(300, 5000),
             aL=rng.uniform(0.2, 0.5), aR=rng.uniform(0.2, 0.5),
             phi=rng.uniform(0, 2 * np.pi), noise_seed=int(rng.integers(0, 2**31)))
    if period_s:
        assert abs(19e3 * period_s - round(19e3 * period_s)) < 1e-9, "the pilot must complete whole cycles in one period"
        p["fL"] = round(p["fL"] * period_s) / period_s
        p["fR"] = round(p["fR"] * period_s) / period_s
    return p


RDS_OFFSET_WORDS = {"A": 0x0FC, "B": 0x198, "C": 0x168, "Cp": 0x350, "D": 0x1B4}   # IEC 62106 annex A
RDS_POLY = 0x5B9                                                                     # x^10+x^8+x^7+x^5+x^4+x^3+1


def rds_block(info16, offset):
    """26-bit RDS block, first transmitted bit first: 16 information bits, then the 10-bit checkword
    (remainder of info * x^10 by the generator polynomial, plus the offset word)."""
    reg = info16 << 10
    for b in range(25, 9, -1):
        if reg & (1 << b):
            reg ^= RDS_POLY << (b - 10)
    word = (info16 << 10) | ((reg & 0x3FF) ^ RDS_OFFSET_WORDS[offset])
    return [(word >> (25 - i)) & 1 for i in range(26)]


def rds_group_bits(n_bits, seed):
    """Source bits of whole RDS groups (blocks A, B, C, D; random information words, one PI code per stream)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    pi_code = int(rng.integers(0, 1 << 16))
    out = []
    while len(out) < n_bits:
        out += rds_block(pi_code, "A")
        for off in ("B", "C", "D"):
            out += rds_block(int(rng.integers(0, 1 << 16)), off)
    return np.array(out[:n_bits], np.int64)


def rds_bitstream(n_bits, seed):
    """What modulates the 57 kHz sub-carrier: the group bits, differentially encoded (0/1)."""
    return np.cumsum(rds_group_bits(n_bits, seed)) % 2


def whole_cycles(mode, n_pairs):
    """True if the 19 kHz pilot completes whole cycles in n_pairs samples (then periodic=True streams of that length exist)."""
    return (n_pairs * 19000) % int(RF_FS[mode]) == 0


def make_stream(mode, n_pairs, seed, rds=False):
    """One stream: uint8[2*n_pairs], interleaved I0 Q0 I1 Q1 ..."""
    fs = RF_FS[mode]
    p = stream_params(seed)
    t = np.arange(n_pairs, dtype=np.float64) / fs
    L = p["aL"] * np.sin(2 * np.pi * p["fL"] * t)
    R = p["aR"] * np.sin(2 * np.pi * p["fR"] * t)
    wp = 2 * np.pi * 19e3 * t + p["phi"]
    m = 0.45 * (L + R) + 0.1 * np.sin(wp) + 0.45 * (L - R) * np.sin(2 * wp)
    if rds:
        nbits = int(np.ceil(t[-1] * 1187.5)) + 2
        bits = rds_bitstream(nbits, seed + 7919) * 2 - 1
        sym = np.floor(t * 2375.0).astype(np.int64)          # Manchester: two half-symbols per bit
        d = bits[sym // 2] * np.where(sym % 2 == 0, 1.0, -1.0)
        m = m + 0.05 * d * np.cos(3 * wp)
    phase = 2 * np.pi * 75e3 * np.cumsum(m) / fs
    rng_n = np.random.Generator(np.random.PCG64(p["noise_seed"]))
    noise = rng_n.normal(0.0, 0.01, size=(2, n_pairs))
    out = np.empty(2 * n_pairs, np.uint8)
    out[0::2] = np.clip(np.rint(100 * (np.cos(phase) + noise[0]) + 128), 0, 255).astype(np.uint8)
    out[1::2] = np.clip(np.rint(100 * (np.sin(phase) + noise[1]) + 128), 0, 255).astype(np.uint8)
    return out


def make_batch(mode, n_streams, n_pairs, base_seed=65, rds=False):
    """uint8[n_streams, 2*n_pairs]; stream s uses seed base_seed + s."""
    return np.stack([make_stream(mode, n_pairs, base_seed + s, rds) for s in range(n_streams)])


def make_batch_torch(mode, n_streams, n_pairs, base_seed=65, device="cuda", chunk_streams=16, rds=False, periodic=False):
    """Same signal model generated on the GPU with torch (benchmark input only:
    the random draws differ from make_batch, the statistics do not).  periodic: every component completes whole cycles
    in the buffer (tones, pilot, FM phase), so feeding the buffer again continues the signal without a jump."""
    import torch
    fs = RF_FS[mode]
    out = torch.empty((n_streams, 2 * n_pairs), dtype=torch.uint8, device=device)
    g = torch.Generator(device=device)
    t = torch.arange(n_pairs, dtype=torch.float64, device=device) / fs
    for s0 in range(0, n_streams, chunk_streams):
        s1 = min(n_streams, s0 + chunk_streams)
        ps = [stream_params(base_seed + s, n_pairs / fs if periodic else None) for s in range(s0, s1)]
        col = lambda k: torch.tensor([p[k] for p in ps], dtype=torch.float64, device=device)[:, None]
        L = col("aL") * torch.sin(2 * np.pi * col("fL") * t)
        R = col("aR") * torch.sin(2 * np.pi * col("fR") * t)
        wp = 2 * np.pi * 19e3 * t + col("phi")
        m = 0.45 * (L + R) + 0.1 * torch.sin(wp) + 0.45 * (L - R) * torch.sin(2 * wp)
        if rds:                                      # same sub-carrier as make_stream(rds=True)
            nbits = int(np.ceil(n_pairs / fs * 1187.5)) + 2
            bits = torch.tensor(np.stack([rds_bitstream(nbits, base_seed + s + 7919) for s in range(s0, s1)]) * 2 - 1,
                                dtype=torch.float64, device=device)
            sym = torch.floor(t * 2375.0).to(torch.int64)
            d = bits[:, sym // 2] * torch.where(sym % 2 == 0, 1.0, -1.0).to(torch.float64)
            m = m + 0.05 * d * torch.cos(3 * wp)
            del bits, sym, d
        del L, R, wp
        if periodic:
            m = m - m.mean(dim=1, keepdim=True)         # the FM phase must return to its start (rounding leaves ~1e-17 per sample)
        phase = (2 * np.pi * 75e3 / fs) * torch.cumsum(m, dim=1)
        del m
        g.manual_seed(base_seed * 1000003 + s0)
        for ch, fn in ((0, torch.cos), (1, torch.sin)):
            x = fn(phase).to(torch.float32)
            x += 0.01 * torch.randn(x.shape, generator=g, device=device, dtype=torch.float32)
            out[s0:s1, ch::2] = torch.clamp(torch.round(100 * x + 128), 0, 255).to(torch.uint8)
        del phase
    return out
