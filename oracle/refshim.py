"""ctypes bindings of oracle/_ref/libref_handle.so (the reference's OWN GmxHandle loaders compiled
unmodified from /root/reference) -- TEST INFRASTRUCTURE.  ``available()`` is False on boxes where the
reference was not present at build time and no prebuilt copy travelled."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_PATH = os.path.join(_HERE, "_ref", "libref_handle.so")
_lib = None
_P = C.c_void_p


def available() -> bool:
    return os.path.exists(_PATH)


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(_PATH)
        L.refshim_new.restype = _P
        L.refshim_new.argtypes = [C.c_int]
        L.refshim_free.argtypes = [_P]
        L.refshim_set_group.argtypes = [_P, C.c_int, _P, C.c_int, C.c_int, _P, _P]
        L.refshim_set_frame.argtypes = [_P, _P, _P]
        L.refshim_load_position.argtypes = [_P, C.c_int, _P]
        L.refshim_load_position_center.argtypes = [_P, C.c_int, C.c_int, _P]
        L.refshim_set_vf.argtypes = [_P, _P, _P]
        L.refshim_load_field.argtypes = [_P, C.c_int, C.c_int, _P]
        L.refshim_load_field_center.argtypes = [_P, C.c_int, C.c_int, C.c_int, _P]
        L.refshim_periodicity.restype = C.c_double
        L.refshim_periodicity.argtypes = [C.c_double, C.c_double]
        L.refshim_eigen_sum.restype = C.c_double
        L.refshim_eigen_sum.argtypes = [_P, C.c_int]
        L.refshim_rightcols_rowmean.argtypes = [_P, C.c_int, C.c_int, C.c_int, _P]
        L.refshim_full_new.restype = _P
        L.refshim_full_new.argtypes = [_P, C.c_int, C.c_int, _P, _P, _P]
        L.refshim_full_free.argtypes = [_P]
        L.refshim_full_set_frame.argtypes = [_P, _P, _P]
        L.refshim_full_load_position.argtypes = [_P, _P, C.c_int]
        L.refshim_full_load_position_center.argtypes = [_P, C.c_int, _P]
        _lib = L
    return _lib


class RefHandle:
    """One group-0-only itp::GmxHandle populated from arrays."""

    def __init__(self, index, napm, mass, charge):
        self._h = lib().refshim_new(1)
        self.index = np.ascontiguousarray(index, dtype=np.int32)
        self.napm = napm
        self.nmol = len(self.index) // napm
        m = np.ascontiguousarray(mass, dtype=np.float64)
        q = np.ascontiguousarray(charge, dtype=np.float64)
        lib().refshim_set_group(self._h, 0, self.index.ctypes.data, len(self.index), napm, m.ctypes.data, q.ctypes.data)

    def set_frame(self, x, box9):
        self._x = np.ascontiguousarray(x, dtype=np.float32)
        self._b = np.ascontiguousarray(box9, dtype=np.float32)
        lib().refshim_set_frame(self._h, self._x.ctypes.data, self._b.ctypes.data)

    def load_position(self):
        out = np.empty((self.napm, self.nmol, 3), dtype=np.float64)
        lib().refshim_load_position(self._h, 0, out.ctypes.data)
        return out.transpose(1, 0, 2)

    def load_position_center(self, com=0):
        out = np.empty((3, self.nmol), dtype=np.float64)
        lib().refshim_load_position_center(self._h, 0, com, out.ctypes.data)
        return out.T

    def set_vf(self, v, f):
        self._v = np.ascontiguousarray(v, dtype=np.float32)
        self._f = np.ascontiguousarray(f, dtype=np.float32)
        lib().refshim_set_vf(self._h, self._v.ctypes.data, self._f.ctypes.data)

    def load_field(self, force=False):
        out = np.empty((self.napm, self.nmol, 3), dtype=np.float64)
        lib().refshim_load_field(self._h, 0, int(force), out.ctypes.data)
        return out.transpose(1, 0, 2)

    def load_field_center(self, force=False, com=0):
        out = np.empty((3, self.nmol), dtype=np.float64)
        lib().refshim_load_field_center(self._h, 0, int(force), com, out.ctypes.data)
        return out.T

    def __del__(self):
        try:
            lib().refshim_free(self._h)
        except Exception:
            pass


def periodicity(dx, box):
    return lib().refshim_periodicity(dx, box)


def eigen_sum(w):
    w = np.ascontiguousarray(w, dtype=np.float64)
    return lib().refshim_eigen_sum(w.ctypes.data, len(w))


class RefHandleFull:
    """itp::GmxHandleFull (heterogeneous molecule sizes) populated from arrays; one group."""

    def __init__(self, index, napm, mass, charge):
        self.index = np.ascontiguousarray(index, dtype=np.int32)
        self.napm = np.ascontiguousarray(napm, dtype=np.int32)
        self.nmol = len(self.napm)
        m = np.ascontiguousarray(mass, dtype=np.float64)
        q = np.ascontiguousarray(charge, dtype=np.float64)
        self._h = lib().refshim_full_new(self.index.ctypes.data, len(self.index), self.nmol, self.napm.ctypes.data, m.ctypes.data,
                                         q.ctypes.data)

    def set_frame(self, x, box9):
        self._x = np.ascontiguousarray(x, dtype=np.float32)
        self._b = np.ascontiguousarray(box9, dtype=np.float32)
        lib().refshim_full_set_frame(self._h, self._x.ctypes.data, self._b.ctypes.data)

    def load_position(self):
        mx = int(self.napm.max())
        out = np.zeros((mx, self.nmol, 3), dtype=np.float64)
        lib().refshim_full_load_position(self._h, out.ctypes.data, mx)
        return out.transpose(1, 0, 2)

    def load_position_center(self, center=0):
        out = np.empty((3, self.nmol), dtype=np.float64)
        lib().refshim_full_load_position_center(self._h, center, out.ctypes.data)
        return out.T
