"""Loads libdy4b200.so and declares the C ABI of include/dy4_b200.h for ctypes."""
import ctypes as C
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libdy4b200.so")


class Dy4Error(RuntimeError):
    pass


def build_library(force=False):
    """nvcc-compile csrc/ for sm_100a into libdy4b200.so (in-tree)."""
    if force:
        subprocess.check_call(["make", "-s", "-C", os.path.join(HERE, "csrc"), "clean"])
    subprocess.check_call(["make", "-s", "-j8", "-C", os.path.join(HERE, "csrc")])
    return LIB_PATH


class ModeParamsC(C.Structure):
    _fields_ = [("rf_Fs", C.c_float), ("rf_decim", C.c_int), ("if_Fs", C.c_float),
                ("audio_decim", C.c_int), ("audio_upsample", C.c_int), ("audio_taps", C.c_int),
                ("block_size", C.c_int), ("if_per_block", C.c_int), ("audio_per_block", C.c_int)]


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            "dy4_b200: %s is missing — build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "or `make -C %s/csrc`. There is no CPU fallback." % (LIB_PATH, HERE))
    L = C.CDLL(LIB_PATH)
    vp, sz, i, f, us = C.c_void_p, C.c_size_t, C.c_int, C.c_float, C.c_ushort
    pf = C.POINTER(C.c_float)
    psz = C.POINTER(C.c_size_t)
    sigs = {
        "dy4_last_error": (C.c_char_p, []),
        "dy4_version": (i, []),
        "dy4_launch_count": (C.c_longlong, []),
        "dy4_mode_params": (i, [i, C.POINTER(ModeParamsC)]),
        "dy4_lpf_taps": (i, [f, f, us, i, vp]),
        "dy4_bpf_taps": (i, [f, f, f, us, i, vp]),
        "dy4_firwin": (i, [i, C.c_double, C.c_double, i, vp]),
        "dy4_rrc_taps": (i, [C.c_double, i, vp]),
        "dy4_dft": (i, [vp, sz, vp]),
        "dy4_idft": (i, [vp, sz, vp]),
        "dy4_estimate_psd": (i, [vp, sz, i, i, vp, vp]),
        "dy4_psd_batch": (i, [vp, sz, i, sz, i, i, vp, sz, vp]),
        "dy4_iq_to_float": (i, [vp, sz, vp]),
        "dy4_convolve_fir": (i, [vp, vp, sz, vp, sz]),
        "dy4_block_fir": (i, [vp, vp, sz, vp, sz, vp, sz]),
        "dy4_decim_fir": (i, [i, vp, vp, sz, vp, sz, vp, sz]),
        "dy4_resample_fir": (i, [i, i, vp, psz, vp, sz, vp, sz, vp, sz]),
        "dy4_fm_demod": (i, [vp, vp, sz, pf, pf, vp]),
        "dy4_pll": (i, [vp, sz, f, f, f, f, f, vp, pf, pf, pf, pf, pf, pf]),
        "dy4_downsample": (i, [vp, sz, sz, vp, psz]),
        "dy4_upsample": (i, [vp, sz, sz, vp, psz]),
        "dy4_delay_block": (i, [vp, sz, vp, sz, vp]),
        "dy4_pointwise_multiply": (i, [vp, sz, vp, sz, vp, psz]),
        "dy4_pointwise_add": (i, [vp, vp, sz, vp]),
        "dy4_pointwise_subtract": (i, [vp, vp, sz, vp]),
        "dy4_interleave": (i, [vp, sz, vp, sz, vp]),
        "dy4_pipeline_create": (i, [i, i, i, i, C.c_uint, C.POINTER(vp)]),
        "dy4_pipeline_destroy": (i, [vp]),
        "dy4_pipeline_reset": (i, [vp]),
        "dy4_pipeline_process": (i, [vp, vp, sz, i, vp, vp, vp, vp]),
        "dy4_pipeline_process_host": (i, [vp, vp, sz, i, vp, vp, i]),
        "dy4_pinned_alloc": (vp, [sz]),
        "dy4_pinned_free": (None, [vp]),
        "dy4_pipeline_rds_read": (i, [vp, vp, vp, sz, C.POINTER(i), vp]),
        "dy4_pipeline_rds_bounds": (i, [vp, C.POINTER(i), C.POINTER(i), C.POINTER(i)]),
        "dy4_pipeline_rds_drain": (i, [vp, vp, sz, vp, sz, vp, sz, vp, sz, vp]),
        "dy4_pipeline_debug_buffers": (i, [vp, C.POINTER(vp), C.POINTER(vp), psz, C.POINTER(i)]),
        "dy4_compute_twiddles": (i, [sz, i, vp]),
        "dy4_fft": (i, [vp, sz, i, vp, sz, vp]),
        "dy4_fft_batch": (i, [vp, sz, i, sz, i, vp, sz, vp, sz, vp]),
        "dy4_pipeline_pll_risk": (i, [vp, vp, i]),
        "dy4_pipeline_sm_partition": (i, [vp, vp, vp]),
        "dy4_pipeline_flush": (i, [vp, vp]),
        "dy4_pipeline_sync": (i, [vp]),
        "dy4_pipeline_profile": (i, [vp, i]),
        "dy4_pipeline_profile_get": (i, [vp, C.POINTER(C.c_double), C.POINTER(C.c_longlong), i]),
        "dy4_pipeline_state_size": (sz, [vp]),
        "dy4_pipeline_get_state": (i, [vp, vp]),
        "dy4_pipeline_set_state": (i, [vp, vp]),
    }
    for name, (res, args) in sigs.items():
        fn = getattr(L, name)          # AttributeError here = header and library disagree
        fn.restype = res
        fn.argtypes = args
    return L, tuple(sigs)


lib, EXPORTS = _load()


def check(rc, who):
    if rc != 0:
        raise Dy4Error("%s failed (%d): %s" % (who, rc, lib.dy4_last_error().decode()))
