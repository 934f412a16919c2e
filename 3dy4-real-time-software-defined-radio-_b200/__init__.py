"""dy4_b200 — B200 (sm_100a) implementation of the 3DY4 FM receiver's hot path.

Host-side mirror of the C ABI in include/dy4_b200.h (ctypes; no torch types
cross the boundary — tensors are passed as raw device pointers).  Two tiers,
as in the header:

* ``filterh``  — the reference's include/filter.h functions by their own names
  (impulseResponseLPF, blockConvolveFIR, fmPLL, ...) on numpy arrays.
* ``fourierh`` — include/fourier.h (DFT, IDFT, estimatePSD) likewise, plus a batched PSD over device rows.
* ``Pipeline`` — the batched receiver over many independent streams.

There is no CPU fallback: everything that computes calls libdy4b200.so, and
importing this package fails loudly if the library has not been built.
The directory name contains characters Python cannot import; the repo root
holds ``dy4_b200.py`` which loads this package under the name ``dy4_b200``.
"""
import os as _os

from ._lib import lib, Dy4Error, LIB_PATH, build_library  # noqa: F401
from .modes import mode_params, ModeParams  # noqa: F401
from .pipeline import Pipeline, launch_count  # noqa: F401
from . import filterh  # noqa: F401
from . import fourierh  # noqa: F401
from . import rds_app  # noqa: F401
from . import rate_change  # noqa: F401
from . import synth  # noqa: F401
from . import shard  # noqa: F401

PACKAGE_DIR = _os.path.dirname(_os.path.abspath(__file__))
__all__ = ["lib", "Dy4Error", "Pipeline", "mode_params", "ModeParams", "filterh", "fourierh", "rds_app", "synth", "launch_count", "build_library"]
