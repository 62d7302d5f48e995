"""B200-native per-frame trajectory analysis (centres -> density -> RDF) behind the itp::GmxHandle surface.

The product is the C-ABI library ``libb2a.so`` (CUDA, sm_100a; ``include/b2a.h``) and the C++ drop-in
header ``include/itp/gmx``.  This Python package is the thin host-side mirror used by tests and
``bench.py``: ctypes bindings (``lib``) and the multi-GPU plumbing (``dist``).  Synthetic inputs and file
writers live in ``workloads/`` (harness).
"""
from . import lib  # noqa: F401

__all__ = ["lib", "dist"]
