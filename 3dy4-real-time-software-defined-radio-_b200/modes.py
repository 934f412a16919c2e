"""Mode table of the receiver (reference src/project.cpp:178-238), read from the library."""
import collections
import ctypes as C

from ._lib import lib, ModeParamsC, check

ModeParams = collections.namedtuple(
    "ModeParams", "rf_Fs rf_decim if_Fs audio_decim audio_upsample audio_taps block_size if_per_block audio_per_block")


def mode_params(mode):
    m = ModeParamsC()
    check(lib.dy4_mode_params(int(mode), C.byref(m)), "dy4_mode_params")
    return ModeParams(*[getattr(m, f) for f in ModeParams._fields])
