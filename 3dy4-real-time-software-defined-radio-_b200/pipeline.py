"""Batched receiver: host-side handle over the throughput tier of include/dy4_b200.h.

What `project <mode> <mono|stereo>` (reference src/project.cpp:137-330) does to one
stdin stream, done to `n_streams` independent streams at once on one B200.  torch
is used only to own device/pinned memory and streams; the library sees raw pointers.
"""
import ctypes as C

import numpy as np

from ._lib import lib, check
from .modes import mode_params

FLAG_EXACT_AUDIO = 1
FLAG_DEBUG_ROWS = 2
FLAG_RDS = 4
FLAG_PIPELINED = 8
KERNELS = ("frontend", "twin_bpf", "pll", "audio", "tails", "rds_bpf", "rds_pll", "rds_baseband", "pll_aux")


def launch_count():
    """Kernels launched by libdy4b200.so in this process so far."""
    return int(lib.dy4_launch_count())


class Pipeline:
    def __init__(self, mode, stereo, n_streams, device=0, exact_audio=False, debug_rows=False, rds=False, pipelined=False):
        self.mode, self.stereo, self.n_streams, self.device = int(mode), bool(stereo), int(n_streams), int(device)
        self.params = mode_params(mode)
        self.channels = 2 if stereo else 1
        self.pipelined, self._in_flight = bool(pipelined) and bool(stereo), []
        self._h = C.c_void_p()
        check(lib.dy4_pipeline_create(self.mode, int(self.stereo), self.n_streams, self.device,
                                      (FLAG_EXACT_AUDIO if exact_audio else 0) | (FLAG_DEBUG_ROWS if debug_rows else 0) | (FLAG_RDS if rds else 0)
                                      | (FLAG_PIPELINED if pipelined else 0),
                                      C.byref(self._h)), "dy4_pipeline_create")

    def close(self):
        if self._h:
            lib.dy4_pipeline_destroy(self._h)        # drains everything queued first
            self._h = C.c_void_p()
            self._in_flight.clear()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def reset(self):
        check(lib.dy4_pipeline_reset(self._h), "dy4_pipeline_reset")

    # ---- device-resident ---------------------------------------------------------------------
    def process(self, iq, n_blocks=None, want=("pcm",), out=None, stream=None):
        """iq: torch uint8 CUDA tensor [n_streams, >= n_blocks*block_size].  Returns a dict of torch tensors
        for the names in `want` ("pcm", "audio", "if").  Asynchronous on `stream` (default: current stream)."""
        import torch
        p = self.params
        assert iq.is_cuda and iq.dtype == torch.uint8 and iq.dim() == 2 and iq.shape[0] == self.n_streams
        assert iq.stride(1) == 1
        if n_blocks is None:
            n_blocks = iq.shape[1] // p.block_size
        assert 0 <= n_blocks and iq.shape[1] >= n_blocks * p.block_size, "every row must hold n_blocks * block_size bytes"
        dev = iq.device
        self._last_blocks = int(n_blocks)
        out = dict(out or {})
        na = n_blocks * p.audio_per_block * self.channels
        if "pcm" in want and "pcm" not in out:
            out["pcm"] = torch.empty((self.n_streams, na), dtype=torch.int16, device=dev)
        if "audio" in want and "audio" not in out:
            out["audio"] = torch.empty((self.n_streams, na), dtype=torch.float32, device=dev)
        if "if" in want and "if" not in out:
            out["if"] = torch.empty((self.n_streams, n_blocks * p.if_per_block), dtype=torch.float32, device=dev)
        if stream is None:
            stream = torch.cuda.current_stream(dev)
        ptr = lambda k: C.c_void_p(out[k].data_ptr()) if k in out else None
        # a single row may carry any stride (numpy/torch leave it arbitrary for size-1 dims)
        row_stride = iq.stride(0) if self.n_streams > 1 else (n_blocks * p.block_size + 15) // 16 * 16
        check(lib.dy4_pipeline_process(self._h, C.c_void_p(iq.data_ptr()), row_stride, int(n_blocks),
                                       ptr("pcm"), ptr("audio"), ptr("if"), C.c_void_p(stream.cuda_stream)),
              "dy4_pipeline_process")
        return out

    def flush(self, stream=None):
        """pipelined=True (DY4_FLAG_PIPELINED): process() calls overlap on the device and are not joined to their stream; flush()
        makes `stream` (default: the current stream) wait for all of them.  Call it before reading any output, before
        overwriting an input, and keep the output tensors alive until then.  A no-op otherwise."""
        import torch
        if stream is None:
            stream = torch.cuda.current_stream(self.device)
        check(lib.dy4_pipeline_flush(self._h, C.c_void_p(stream.cuda_stream)), "dy4_pipeline_flush")

    def sync(self):
        """Block until everything queued on this pipeline is done (dy4_pipeline_sync).  pipelined=True: process_host() returns when the
        call is queued; read its outputs (and reuse its input) after sync()."""
        check(lib.dy4_pipeline_sync(self._h), "dy4_pipeline_sync")
        self._in_flight.clear()

    # ---- host buffers (H2D / compute / D2H overlapped inside the library) ------------------------
    def process_host(self, iq, n_blocks=None, want=("pcm",), out=None, chunk_blocks=0):
        """iq: uint8 host array/tensor [n_streams, >= n_blocks*block_size] (pinned torch tensor recommended).
        Synchronous — unless the pipeline was made with pipelined=True: then the call returns when it is queued (its upload runs
        beside the kernels of the call before) and outputs / input are the caller's again after sync().
        Returns numpy arrays (or fills the arrays/tensors passed in `out`)."""
        p = self.params
        a = _host_view(iq)
        assert a.dtype == np.uint8 and a.ndim == 2 and a.shape[0] == self.n_streams and a.strides[1] == 1
        if n_blocks is None:
            n_blocks = a.shape[1] // p.block_size
        assert 0 <= n_blocks and a.shape[1] >= n_blocks * p.block_size, "every row must hold n_blocks * block_size bytes"
        self._last_blocks = int(n_blocks)
        na = n_blocks * p.audio_per_block * self.channels
        out = dict(out or {})
        if "pcm" in want and "pcm" not in out:
            out["pcm"] = np.empty((self.n_streams, na), np.int16)
        if "audio" in want and "audio" not in out:
            out["audio"] = np.empty((self.n_streams, na), np.float32)
        hp = lambda k: C.c_void_p(_host_view(out[k]).ctypes.data) if k in out else None
        row_stride = a.strides[0] if self.n_streams > 1 else n_blocks * p.block_size
        check(lib.dy4_pipeline_process_host(self._h, C.c_void_p(a.ctypes.data), row_stride, int(n_blocks),
                                            hp("pcm"), hp("audio"), int(chunk_blocks)), "dy4_pipeline_process_host")
        if self.pipelined:
            self._in_flight.append((a, out))         # the copies are still running: keep the buffers alive until sync()
        return out

    def rds_read(self, stream=None):
        """(rrc_i, rrc_q): the RDS baseband produced by the last process call, torch float32 [n_streams, n] each."""
        import torch
        dev = "cuda:%d" % self.device
        cap = 16 + (self._last_blocks * self.params.if_per_block * 19 + 119) // 120
        i_t = torch.empty((self.n_streams, cap), dtype=torch.float32, device=dev)
        q_t = torch.empty((self.n_streams, cap), dtype=torch.float32, device=dev)
        n = C.c_int()
        if stream is None:
            stream = torch.cuda.current_stream(i_t.device)
        check(lib.dy4_pipeline_rds_read(self._h, C.c_void_p(i_t.data_ptr()), C.c_void_p(q_t.data_ptr()), cap, C.byref(n),
                                        C.c_void_p(stream.cuda_stream)), "dy4_pipeline_rds_read")
        return i_t[:, :n.value], q_t[:, :n.value]

    def rds_drain(self, raw=False):
        """Everything the RDS back half decoded since the last drain: per stream, dict(symbols, bits, events) of numpy
        arrays (events: rows of [block type 0..4 = A,B,C,C',D, bit position, false-positive flag, 16-bit word]; groups:
        rows of the [A, B, C, D] words of every complete group, the input of rds_app.ApplicationLayer).
        raw=True returns the padded arrays and the counts instead: (symbols[S,:], bits[S,:], events[S,:,4], groups[S,:,4], counts[S,4])."""
        import numpy as np
        ms, mb, me = C.c_int(), C.c_int(), C.c_int()
        check(lib.dy4_pipeline_rds_bounds(self._h, C.byref(ms), C.byref(mb), C.byref(me)), "dy4_pipeline_rds_bounds")
        S = self.n_streams
        sym = np.zeros((S, max(ms.value, 1)), np.int8)
        bits = np.zeros((S, max(mb.value, 1)), np.int8)
        ev = np.zeros((S, max(me.value, 1), 4), np.int32)
        grp = np.zeros((S, max(me.value, 1), 4), np.int32)
        cnt = np.zeros((S, 4), np.int32)
        check(lib.dy4_pipeline_rds_drain(self._h, C.c_void_p(sym.ctypes.data), sym.shape[1], C.c_void_p(bits.ctypes.data), bits.shape[1],
                                         C.c_void_p(ev.ctypes.data), ev.shape[1], C.c_void_p(grp.ctypes.data), grp.shape[1],
                                         C.c_void_p(cnt.ctypes.data)), "dy4_pipeline_rds_drain")
        if raw:
            return sym, bits, ev, grp, cnt
        return [dict(symbols=sym[s, :cnt[s, 0]].copy(), bits=bits[s, :cnt[s, 1]].copy(), events=ev[s, :cnt[s, 2]].copy(),
                     groups=grp[s, :cnt[s, 3]].copy()) for s in range(S)]

    # ---- diagnostics ----------------------------------------------------------------------------
    def debug_pilot_nco(self):
        """(pilot, nco) of the last sub-chunk as torch tensors (copies)."""
        import torch
        dp, dn, stride, n = C.c_void_p(), C.c_void_p(), C.c_size_t(), C.c_int()
        check(lib.dy4_pipeline_debug_buffers(self._h, C.byref(dp), C.byref(dn), C.byref(stride), C.byref(n)), "dy4_pipeline_debug_buffers")
        res = []
        for ptr in (dp, dn):
            t = torch.empty((self.n_streams, n.value), dtype=torch.float32, device="cuda:%d" % self.device)
            rc = _cudart().cudaMemcpy2D(C.c_void_p(t.data_ptr()), C.c_size_t(n.value * 4), ptr, C.c_size_t(stride.value * 4),
                                        C.c_size_t(n.value * 4), C.c_size_t(self.n_streams), 3)
            assert rc == 0, rc
            res.append(t)
        return res

    def pll_risk(self, reset=True):
        """int32[n_streams]: PLL evaluations narrowed to float within 2 double-ulps of a rounding boundary (dy4_pipeline_pll_risk)."""
        out = np.zeros(self.n_streams, np.int32)
        check(lib.dy4_pipeline_pll_risk(self._h, C.c_void_p(out.ctypes.data), int(reset)), "dy4_pipeline_pll_risk")
        return out

    def sm_partition(self):
        """(loop SMs, other SMs) of the opt-in SM partition (DY4_LOOP_SMS), (0, 0) when unpartitioned (dy4_pipeline_sm_partition)."""
        a, b = C.c_int(0), C.c_int(0)
        check(lib.dy4_pipeline_sm_partition(self._h, C.byref(a), C.byref(b)), "dy4_pipeline_sm_partition")
        return a.value, b.value

    def profile(self, enable=True):
        check(lib.dy4_pipeline_profile(self._h, int(enable)), "dy4_pipeline_profile")

    def profile_get(self, reset=True):
        ms = (C.c_double * len(KERNELS))()
        n = (C.c_longlong * len(KERNELS))()
        check(lib.dy4_pipeline_profile_get(self._h, ms, n, int(reset)), "dy4_pipeline_profile_get")
        return {k: {"ms": ms[i], "launches": int(n[i])} for i, k in enumerate(KERNELS)}

    def get_state(self):
        buf = np.empty(lib.dy4_pipeline_state_size(self._h), np.uint8)
        check(lib.dy4_pipeline_get_state(self._h, C.c_void_p(buf.ctypes.data)), "dy4_pipeline_get_state")
        return buf

    def set_state(self, buf):
        buf = np.ascontiguousarray(buf, np.uint8)
        assert buf.size == lib.dy4_pipeline_state_size(self._h)
        check(lib.dy4_pipeline_set_state(self._h, C.c_void_p(buf.ctypes.data)), "dy4_pipeline_set_state")


def _host_view(x):
    if isinstance(x, np.ndarray):
        return x
    return x.numpy()      # CPU torch tensor (pinned or not) shares memory with numpy


_rt = None


def _cudart():
    global _rt
    if _rt is None:
        import torch, glob, os
        cands = glob.glob(os.path.join(os.path.dirname(torch.__file__), "lib", "libcudart*.so*")) + \
            glob.glob(os.path.join(os.path.dirname(os.path.dirname(torch.__file__)), "nvidia", "cuda_runtime", "lib", "libcudart.so*")) + \
            glob.glob("/usr/local/cuda/lib64/libcudart.so*")
        _rt = C.CDLL(cands[0])
    return _rt
