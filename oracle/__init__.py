"""CPU checkers for the dy4 hot path — TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
``--impl reference`` leg may import this package, and only as the checker or
the timed CPU baseline.  The product (the package next to this directory)
never imports it and has no CPU fallback.
"""
from .cpu import CpuReceiver, load, build, have_ref  # noqa: F401
