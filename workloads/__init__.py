"""Test / bench harness (NOT product): synthetic systems of SURVEY.md 8(d) and writers for the file formats the
libgromacs stand-in reads.  The product is gmx2020postanalysis_b200/ (libb2a.so + ctypes mirror) and include/."""
from . import synth, trajio  # noqa: F401
