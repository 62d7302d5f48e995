"""ctypes bindings of oracle/liboracle.so -- TEST INFRASTRUCTURE, never imported by the product."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_lib = None
_P = C.c_void_p


def lib():
    global _lib
    if _lib is None:
        path = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(path):
            raise RuntimeError(f"{path} missing: run `make -C oracle port` (or __graft_entry__.build())")
        L = C.CDLL(path)
        L.orc_periodicity.restype = C.c_double
        L.orc_periodicity.argtypes = [C.c_double, C.c_double]
        L.orc_eigen_sum.restype = C.c_double
        L.orc_eigen_sum.argtypes = [_P, C.c_int]
        L.orc_load_position.argtypes = [_P, _P, C.c_int, C.c_int, _P, _P]
        L.orc_load_position_center.argtypes = [_P, _P, C.c_int, C.c_int, _P, _P, _P]
        L.orc_dipole_center.argtypes = [_P, _P, C.c_int, C.c_int, _P, _P, _P]
        L.orc_density_accum.argtypes = [_P, C.c_int, C.c_int, C.c_float, C.c_float, C.c_double, C.c_int, C.c_int,
                                        C.c_int, C.c_float, C.c_float, _P, _P]
        L.orc_rdf_default_rm.restype = C.c_float
        L.orc_rdf_default_rm.argtypes = [_P]
        L.orc_rdf_rbin.restype = C.c_int
        L.orc_rdf_rbin.argtypes = [C.c_float]
        L.orc_rdf_accum.argtypes = [_P, C.c_int, _P, C.c_int, _P, C.c_int, C.c_float, C.c_float, C.c_int, C.c_int,
                                    C.c_float, C.c_float, C.c_float, C.c_int, _P, _P]
        L.orc_rdf_normalise.argtypes = [_P, C.c_int, C.c_int, _P]
        L.orc_orient_accum.argtypes = [_P, _P, C.c_int, C.c_int, _P, C.c_int, C.c_float, C.c_float, C.c_float,
                                       C.c_float, C.c_float, C.c_float, C.c_float, C.c_int, C.c_int, C.c_int,
                                       C.c_int, C.c_int, C.c_int, _P, _P, _P]
        L.orc_full_load_position.argtypes = [_P, _P, C.c_int, _P, C.c_int, _P, _P]
        L.orc_full_load_position_center.argtypes = [_P, _P, C.c_int, _P, _P, C.c_int, _P, _P]
        L.orc_field_center.argtypes = [_P, _P, C.c_int, C.c_int, _P, C.c_double, _P]
        L.orc_field_values.argtypes = [_P, _P, C.c_int, C.c_int, _P]
        L.orc_pair_count.restype = C.c_int
        L.orc_pair_count.argtypes = [_P, C.c_int, _P, C.c_int, _P, C.c_float, C.c_int, C.c_float, C.c_float, C.c_int,
                                     C.c_float, C.c_float, C.c_float, C.c_float, _P]
        _lib = L
    return _lib


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def lbox_from_box9(box9):
    """Lbox[k] = double(float box[k][k])  (gmx_handle.ipp:72-74)."""
    b = _f32(box9).reshape(3, 3)
    return np.array([b[0, 0], b[1, 1], b[2, 2]], dtype=np.float64)


def periodicity(dx, box):
    return lib().orc_periodicity(dx, box)


def eigen_sum(w):
    w = _f64(w)
    return lib().orc_eigen_sum(w.ctypes.data, len(w))


def load_position(x, index, napm, Lbox):
    """-> [nmol, napm, 3] float64 view of the boxd buffer (memory order (i + j*nmol)*3 + k)."""
    x, index, Lbox = _f32(x), _i32(index), _f64(Lbox)
    nmol = len(index) // napm
    out = np.empty((napm, nmol, 3), dtype=np.float64)
    lib().orc_load_position(x.ctypes.data, index.ctypes.data, nmol, napm, Lbox.ctypes.data, out.ctypes.data)
    return out.transpose(1, 0, 2)


def load_position_center(x, index, napm, w, Lbox, dipole_guard=False):
    """-> [nmol, 3] float64, Fortran (column-major) memory like itp::matd."""
    x, index, w, Lbox = _f32(x), _i32(index), _f64(w), _f64(Lbox)
    nmol = len(index) // napm
    out = np.empty((3, nmol), dtype=np.float64)
    fn = lib().orc_dipole_center if dipole_guard else lib().orc_load_position_center
    fn(x.ctypes.data, index.ctypes.data, nmol, napm, w.ctypes.data, Lbox.ctypes.data, out.ctypes.data)
    return out.T


def _colmajor(posc):
    """[nmol,3] array -> contiguous (3, nmol) buffer = column-major nmol x 3."""
    return np.ascontiguousarray(np.asarray(posc, dtype=np.float64).T)


def density_accum(posc, dim, low, up, dbin, nbin, hist=None, upper_inclusive=False, dim2=-1, low2=0.0, up2=0.0):
    p = _colmajor(posc)
    if hist is None:
        hist = np.zeros(nbin, dtype=np.uint64)
    dropped = np.zeros(1, dtype=np.uint64)
    lib().orc_density_accum(p.ctypes.data, p.shape[1], dim, low, up, float(dbin), nbin, int(upper_inclusive), dim2,
                            low2, up2, hist.ctypes.data, dropped.ctypes.data)
    return hist, int(dropped[0])


def rdf_default_rm(Lbox):
    Lbox = _f64(Lbox)
    return float(np.float32(lib().orc_rdf_default_rm(Lbox.ctypes.data)))


def rdf_rbin(Rm):
    return int(lib().orc_rdf_rbin(Rm))


def rdf_accum(posc1, posc2, Lbox, nbin, low, up, dim, dim2, low2, up2, Rm, Rbin, rdf=None, stats=None):
    """rdf: uint64 [Rbin, nbin] C-order == (meZ + meR*nbin) column-major nbin x Rbin."""
    p1, p2, Lbox = _colmajor(posc1), _colmajor(posc2), _f64(Lbox)
    if rdf is None:
        rdf = np.zeros((Rbin, nbin), dtype=np.uint64)
    if stats is None:
        stats = np.zeros(3, dtype=np.uint64)
    lib().orc_rdf_accum(p1.ctypes.data, p1.shape[1], p2.ctypes.data, p2.shape[1], Lbox.ctypes.data, nbin, low, up, dim,
                        dim2, low2, up2, Rm, Rbin, rdf.ctypes.data, stats.ctypes.data)
    return rdf, stats


def rdf_normalise(rdf):
    rdf = np.ascontiguousarray(rdf, dtype=np.uint64)
    Rbin, nbin = rdf.shape
    g = np.empty((Rbin, nbin), dtype=np.float64)
    with np.errstate(all="ignore"):
        lib().orc_rdf_normalise(rdf.ctypes.data, nbin, Rbin, g.ctypes.data)
    return g


def orient_accum(pos, posc, Lbox, DIMN, low, up, c1, c2, dr, CR, Cr, nAngle, tp1, tp2, tp3, DIMNOrient, method,
                 orient=None, ncm=None):
    """pos: [nmol,napm,3]; orient: uint64 [nAngle, nbin1] C-order == (meZ + meA*nbin1)."""
    pos = np.asarray(pos, dtype=np.float64)
    nmol, napm, _ = pos.shape
    pbuf = np.ascontiguousarray(pos.transpose(1, 0, 2))
    pc, Lbox = _colmajor(posc), _f64(Lbox)
    nbin1 = int(np.float32(np.float32(CR) - np.float32(Cr)) / np.float32(dr))
    if orient is None:
        orient = np.zeros((nAngle, nbin1), dtype=np.uint64)
    if ncm is None:
        ncm = np.zeros(nbin1, dtype=np.uint64)
    dropped = np.zeros(1, dtype=np.uint64)
    lib().orc_orient_accum(pbuf.ctypes.data, pc.ctypes.data, nmol, napm, Lbox.ctypes.data, DIMN, low, up, c1, c2, dr,
                           CR, Cr, nAngle, tp1, tp2, tp3, DIMNOrient, method, orient.ctypes.data, ncm.ctypes.data,
                           dropped.ctypes.data)
    return orient, ncm, int(dropped[0])


def full_load_position(x, index, napm, Lbox):
    """GmxHandleFull::loadPosition -> [nmol, max_napm, 3] (slots j >= napm[i] are 0)."""
    x, index, napm, Lbox = _f32(x), _i32(index), _i32(napm), _f64(Lbox)
    nmol, mx = len(napm), int(napm.max())
    out = np.zeros((mx, nmol, 3), dtype=np.float64)
    lib().orc_full_load_position(x.ctypes.data, index.ctypes.data, nmol, napm.ctypes.data, mx, Lbox.ctypes.data, out.ctypes.data)
    return out.transpose(1, 0, 2)


def full_load_position_center(x, index, napm, w, Lbox, geometry=False):
    x, index, napm, w, Lbox = _f32(x), _i32(index), _i32(napm), _f64(w), _f64(Lbox)
    nmol = len(napm)
    out = np.empty((3, nmol), dtype=np.float64)
    lib().orc_full_load_position_center(x.ctypes.data, index.ctypes.data, nmol, napm.ctypes.data, w.ctypes.data, int(geometry),
                                        Lbox.ctypes.data, out.ctypes.data)
    return out.T


def field_center(v, index, napm, w):
    """loadVelocityCenter / loadForceCenter -> [nmol, 3] (column-major memory)."""
    v, index, w = _f32(v), _i32(index), _f64(w)
    nmol = len(index) // napm
    out = np.empty((3, nmol), dtype=np.float64)
    lib().orc_field_center(v.ctypes.data, index.ctypes.data, nmol, napm, w.ctypes.data, eigen_sum(w), out.ctypes.data)
    return out.T


def field_values(v, index, napm):
    """loadVelocity / loadForce -> [nmol, napm, 3] view of the boxd buffer."""
    v, index = _f32(v), _i32(index)
    nmol = len(index) // napm
    out = np.empty((napm, nmol, 3), dtype=np.float64)
    lib().orc_field_values(v.ctypes.data, index.ctypes.data, nmol, napm, out.ctypes.data)
    return out.transpose(1, 0, 2)


def pair_count(posc1, posc2, Lbox, Rmin, dim, low, up, cyl=None):
    """calc_pair_dist_bulk.cpp:66-92 (cyl=None) / calc_pair_dist.cpp:80-106 (cyl=(cx, cy, Rlow, Rup)).
    -> (count[n0] int32 with -1 for filtered-out i, numA)."""
    p1, p2, Lbox = _colmajor(posc1), _colmajor(posc2), _f64(Lbox)
    count = np.empty(p1.shape[1], dtype=np.int32)
    cx, cy, rl, ru = cyl if cyl is not None else (0.0, 0.0, 0.0, 0.0)
    numA = lib().orc_pair_count(p1.ctypes.data, p1.shape[1], p2.ctypes.data, p2.shape[1], Lbox.ctypes.data, Rmin, dim, low, up,
                                int(cyl is not None), cx, cy, rl, ru, count.ctypes.data)
    return count, int(numA)
