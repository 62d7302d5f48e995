"""ctypes binding of libb2a.so (include/b2a.h).  Fails loudly when the CUDA library is missing or no GPU is
visible: there is no CPU fallback anywhere in the product path."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("B2A_LIB") or os.path.join(_HERE, "libb2a.so")   # B2A_LIB: an experimental build (A/B runs)
_lib = None
_P = C.c_void_p

STAT_FRAMES, STAT_DROPPED, STAT_EXACT, STAT_MOVED, STAT_SELECTED, STAT_PAIRS, STAT_OVERFLOW, STAT_MISPROVEN = range(8)
STAT_NSLOTS = 8
COM_MASS, COM_GEOMETRY, COM_CHARGE, COM_DIPOLE = range(4)
FIELD_V, FIELD_F = 1, 2

# every symbol include/b2a.h declares (tests check the library exports all of them)
SYMBOLS = [
    "b2a_last_error", "b2a_version", "b2a_device_count", "b2a_create", "b2a_destroy", "b2a_sync", "b2a_stream",
    "b2a_pinned_alloc", "b2a_pinned_free", "b2a_set_group", "b2a_set_group_full", "b2a_set_group_wsum", "b2a_group_nmol",
    "b2a_submit_batch", "b2a_batch_frames", "b2a_centers", "b2a_centers_batch", "b2a_positions",
    "b2a_submit_field", "b2a_field_centers", "b2a_field_centers_batch", "b2a_field_values",
    "b2a_density_create", "b2a_region_create", "b2a_region_read_frames", "b2a_rdf_create", "b2a_orient_create", "b2a_coord_create", "b2a_coord_read_frames",
    "b2a_coord_counts", "b2a_coord_fold", "b2a_accumulate", "b2a_acc_reset", "b2a_acc_size",
    "b2a_acc_device_ptr", "b2a_acc_read", "b2a_rdf_normalise", "b2a_rdf_tune", "b2a_rdf_cells", "b2a_acc_fast", "b2a_invalidate", "b2a_timing_enable",
    "b2a_timing_read",
    "b2a_create_multi", "b2a_num_devices", "b2a_member", "b2a_current_device", "b2a_finalize", "b2a_comm_unique_id", "b2a_comm_init",
    "b2a_set_frame_base", "b2a_series_size", "b2a_series_read", "b2a_acc_public_size", "b2a_centers_device", "b2a_acc_also_centers", "b2a_set_host_buffers", "b2a_join",
]


class RegionSpec(C.Structure):
    _fields_ = [("grp", C.c_int), ("com", C.c_int), ("nfilter", C.c_int), ("fdim", C.c_int * 3), ("flow", C.c_float * 3),
                ("fup", C.c_float * 3), ("fincl", C.c_int * 3), ("wrap", C.c_int), ("cyl", C.c_int), ("cx", C.c_float), ("cy", C.c_float),
                ("R2", C.c_double), ("nbdim", C.c_int), ("bdim", C.c_int * 2), ("blow", C.c_float * 2), ("dbin", C.c_double * 2),
                ("nb", C.c_int * 2), ("per_frame", C.c_int)]


class DensitySpec(C.Structure):
    _fields_ = [("grp", C.c_int), ("com", C.c_int), ("dim", C.c_int), ("low", C.c_float), ("up", C.c_float),
                ("dbin", C.c_double), ("nbin", C.c_int), ("upper_inclusive", C.c_int), ("dim2", C.c_int),
                ("low2", C.c_float), ("up2", C.c_float)]


class RdfSpec(C.Structure):
    _fields_ = [("grp0", C.c_int), ("grp1", C.c_int), ("com0", C.c_int), ("com1", C.c_int), ("nbin", C.c_int),
                ("low", C.c_float), ("up", C.c_float), ("dim", C.c_int), ("dim2", C.c_int), ("low2", C.c_float),
                ("up2", C.c_float), ("Rm", C.c_float), ("Rbin", C.c_int)]


class OrientSpec(C.Structure):
    _fields_ = [("grp", C.c_int), ("com", C.c_int), ("DIMN", C.c_int), ("low", C.c_float), ("up", C.c_float),
                ("c1", C.c_float), ("c2", C.c_float), ("dr", C.c_float), ("CR", C.c_float), ("Cr", C.c_float),
                ("nAngle", C.c_int), ("tp1", C.c_int), ("tp2", C.c_int), ("tp3", C.c_int), ("DIMNOrient", C.c_int),
                ("method", C.c_int)]


class CoordSpec(C.Structure):
    _fields_ = [("grp0", C.c_int), ("grp1", C.c_int), ("com0", C.c_int), ("com1", C.c_int), ("Rmin", C.c_float),
                ("dim", C.c_int), ("low", C.c_float), ("up", C.c_float), ("use_cyl", C.c_int), ("cx", C.c_float),
                ("cy", C.c_float), ("Rlow", C.c_float), ("Rup", C.c_float), ("max_count", C.c_int)]


class B2AError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libb2a error {code}: {msg}")
        self.code = code


def load():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} is missing -- build it with __graft_entry__.build(); "
                               "the B200 path has no CPU fallback")
        L = C.CDLL(LIB_PATH)
        L.b2a_last_error.restype = C.c_char_p
        L.b2a_version.restype = C.c_char_p
        L.b2a_device_count.restype = C.c_int
        L.b2a_create.argtypes = [C.POINTER(_P), C.c_int, C.c_int, C.c_int]
        L.b2a_destroy.argtypes = [_P]
        L.b2a_destroy.restype = None
        L.b2a_sync.argtypes = [_P]
        L.b2a_stream.argtypes = [_P]
        L.b2a_stream.restype = _P
        L.b2a_pinned_alloc.argtypes = [C.c_size_t]
        L.b2a_pinned_alloc.restype = _P
        L.b2a_pinned_free.argtypes = [_P]
        L.b2a_pinned_free.restype = None
        L.b2a_set_group.argtypes = [_P, C.c_int, _P, C.c_int, C.c_int, _P, _P]
        L.b2a_set_group_full.argtypes = [_P, C.c_int, _P, C.c_int, C.c_int, _P, _P, _P]
        L.b2a_set_group_wsum.argtypes = [_P, C.c_int, C.c_int, C.c_double]
        L.b2a_group_nmol.argtypes = [_P, C.c_int]
        L.b2a_submit_batch.argtypes = [_P, C.c_int, _P, _P]
        L.b2a_batch_frames.argtypes = [_P]
        L.b2a_centers.argtypes = [_P, C.c_int, C.c_int, C.c_int, _P]
        L.b2a_centers_batch.argtypes = [_P, C.c_int, C.c_int, _P]
        L.b2a_positions.argtypes = [_P, C.c_int, C.c_int, _P]
        L.b2a_submit_field.argtypes = [_P, C.c_int, C.c_int, _P]
        L.b2a_field_centers.argtypes = [_P, C.c_int, C.c_int, C.c_int, C.c_int, _P]
        L.b2a_field_centers_batch.argtypes = [_P, C.c_int, C.c_int, C.c_int, _P]
        L.b2a_field_values.argtypes = [_P, C.c_int, C.c_int, C.c_int, _P]
        L.b2a_density_create.argtypes = [_P, C.POINTER(DensitySpec), C.POINTER(C.c_int)]
        L.b2a_region_create.argtypes = [_P, C.POINTER(RegionSpec), C.POINTER(C.c_int)]
        L.b2a_region_read_frames.argtypes = [_P, C.c_int, _P]
        L.b2a_rdf_create.argtypes = [_P, C.POINTER(RdfSpec), C.POINTER(C.c_int)]
        L.b2a_orient_create.argtypes = [_P, C.POINTER(OrientSpec), C.POINTER(C.c_int)]
        L.b2a_coord_create.argtypes = [_P, C.POINTER(CoordSpec), C.POINTER(C.c_int)]
        L.b2a_coord_read_frames.argtypes = [_P, C.c_int, _P, _P]
        L.b2a_coord_counts.argtypes = [_P, C.c_int, C.c_int, _P]
        L.b2a_coord_fold.argtypes = [_P, _P, C.c_int, C.c_int, _P]
        L.b2a_accumulate.argtypes = [_P, C.c_int]
        L.b2a_acc_reset.argtypes = [_P, C.c_int]
        L.b2a_acc_size.argtypes = [_P, C.c_int, C.POINTER(C.c_size_t)]
        L.b2a_acc_device_ptr.argtypes = [_P, C.c_int, C.POINTER(_P)]
        L.b2a_acc_read.argtypes = [_P, C.c_int, _P, _P]
        L.b2a_rdf_normalise.argtypes = [_P, C.c_int, C.c_int, _P]
        L.b2a_rdf_tune.argtypes = [_P, C.c_int, C.c_int, C.c_int]
        L.b2a_rdf_cells.argtypes = [_P, C.c_int, C.c_int]
        L.b2a_acc_fast.argtypes = [_P, C.c_int, C.c_int]
        L.b2a_invalidate.argtypes = [_P]
        L.b2a_timing_enable.argtypes = [_P, C.c_int]
        L.b2a_timing_read.argtypes = [_P, _P, _P]
        L.b2a_create_multi.argtypes = [C.POINTER(_P), C.c_int, _P, C.c_int, C.c_int]
        L.b2a_num_devices.argtypes = [_P]
        L.b2a_member.argtypes = [_P, C.c_int]
        L.b2a_member.restype = _P
        L.b2a_current_device.argtypes = [_P]
        L.b2a_finalize.argtypes = [_P]
        L.b2a_comm_unique_id.argtypes = [_P]
        L.b2a_comm_init.argtypes = [_P, C.c_int, C.c_int, _P]
        L.b2a_set_frame_base.argtypes = [_P, C.c_longlong]
        L.b2a_series_size.argtypes = [_P, C.c_int, C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)]
        L.b2a_series_read.argtypes = [_P, C.c_int, _P, _P]
        L.b2a_acc_public_size.argtypes = [_P, C.c_int, C.POINTER(C.c_size_t)]
        L.b2a_centers_device.argtypes = [_P, C.c_int, C.c_int, C.POINTER(_P)]
        L.b2a_acc_also_centers.argtypes = [_P, C.c_int, C.c_int]
        L.b2a_set_host_buffers.argtypes = [_P, C.c_int]
        L.b2a_join.argtypes = [_P]
        _lib = L
    return _lib


def check(rc):
    if rc != 0:
        raise B2AError(rc, load().b2a_last_error().decode())


_PINNED = {}


def pinned_empty(shape, dtype):
    """numpy array over page-locked host memory (cudaMallocHost through the C ABI) for asynchronous H2D copies.
    The allocation lives until pinned_free(array) or the end of the process."""
    L = load()
    n = int(np.prod(shape)) * np.dtype(dtype).itemsize
    p = L.b2a_pinned_alloc(max(n, 1))
    if not p:
        raise B2AError(2, L.b2a_last_error().decode())
    buf = (C.c_char * max(n, 1)).from_address(p)
    _PINNED[p] = buf
    return np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)


def pinned_free(arr):
    """Release an array made by pinned_empty (the caller must not touch it, or any view of it, afterwards)."""
    p = arr.ctypes.data
    if _PINNED.pop(p, None) is not None:
        load().b2a_pinned_free(p)


class Context:
    """One GPU, one stream: owner of the device-resident frame batch, group tables and accumulators."""

    def __init__(self, natoms: int, max_frames: int, device: int = 0, devices=None):
        """devices: list of CUDA devices -> one context driving all of them (b2a_create_multi): batches are dealt to the
        devices in turn, read() returns the totals over the devices."""
        self.L = load()
        self._h = _P()
        if devices is None:
            check(self.L.b2a_create(C.byref(self._h), device, natoms, max_frames))
            self.devices = [device]
        else:
            dv = np.ascontiguousarray(devices, dtype=np.int32)
            check(self.L.b2a_create_multi(C.byref(self._h), len(dv), dv.ctypes.data, natoms, max_frames))
            self.devices, device = [int(d) for d in dv], int(dv[0])
        self.natoms, self.max_frames, self.device = natoms, max_frames, device
        self._keep = []

    # ---- several GPUs / processes ------------------------------------------------------------------------
    def finalize(self):
        """The merge: one 64-bit integer all-reduce per accumulator over the devices (and processes) of the communicator."""
        check(self.L.b2a_finalize(self._h))

    def comm_init(self, nprocs, rank, uid: bytes):
        assert len(uid) == 128
        buf = C.create_string_buffer(uid, 128)
        check(self.L.b2a_comm_init(self._h, nprocs, rank, C.cast(buf, _P)))

    def set_host_buffers(self, n):
        check(self.L.b2a_set_host_buffers(self._h, n))

    def set_frame_base(self, first_frame):
        check(self.L.b2a_set_frame_base(self._h, int(first_frame)))

    def current_device(self):
        return self.L.b2a_current_device(self._h)

    def series(self, acc):
        """-> (frame_index [n] int64, records [n, words] uint64) ordered by global frame index over all devices."""
        n, w = C.c_size_t(), C.c_size_t()
        check(self.L.b2a_series_size(self._h, acc, C.byref(n), C.byref(w)))
        idx = np.zeros(n.value, dtype=np.int64)
        out = np.zeros((n.value, w.value), dtype=np.uint64)
        check(self.L.b2a_series_read(self._h, acc, idx.ctypes.data, out.ctypes.data))
        return idx, out

    def close(self):
        if self._h:
            self.L.b2a_destroy(self._h)
            self._h = _P()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def stream(self) -> int:
        return int(self.L.b2a_stream(self._h) or 0)

    def sync(self):
        check(self.L.b2a_sync(self._h))

    def join(self):
        """Order the context stream behind the accumulators' internal side streams (asynchronous)."""
        check(self.L.b2a_join(self._h))

    def set_group(self, grp, index, napm, mass, charge):
        idx = np.ascontiguousarray(index, dtype=np.int32)
        m = np.ascontiguousarray(mass, dtype=np.float64)
        q = np.ascontiguousarray(charge, dtype=np.float64)
        check(self.L.b2a_set_group(self._h, grp, idx.ctypes.data, len(idx), napm, m.ctypes.data, q.ctypes.data))

    def set_group_full(self, grp, index, napm, mass, charge):
        """Heterogeneous molecules (GmxHandleFull): napm[nmol], mass/charge per index position."""
        idx = np.ascontiguousarray(index, dtype=np.int32)
        na = np.ascontiguousarray(napm, dtype=np.int32)
        m = np.ascontiguousarray(mass, dtype=np.float64)
        q = np.ascontiguousarray(charge, dtype=np.float64)
        check(self.L.b2a_set_group_full(self._h, grp, idx.ctypes.data, len(idx), len(na), na.ctypes.data, m.ctypes.data, q.ctypes.data))

    def set_group_wsum(self, grp, com, wsum):
        check(self.L.b2a_set_group_wsum(self._h, grp, com, wsum))

    def nmol(self, grp):
        return self.L.b2a_group_nmol(self._h, grp)

    def submit(self, x, box9):
        """x: [K, natoms, 3] float32 (C-contiguous), box9: [9] or [K, 9] float32."""
        x = np.ascontiguousarray(x, dtype=np.float32)
        if x.ndim == 2:
            x = x[None]
        K = x.shape[0]
        b = np.ascontiguousarray(box9, dtype=np.float32).reshape(-1, 9)
        if b.shape[0] == 1 and K > 1:
            b = np.ascontiguousarray(np.repeat(b, K, axis=0))
        assert x.shape[1:] == (self.natoms, 3) and b.shape[0] == K
        check(self.L.b2a_submit_batch(self._h, K, x.ctypes.data, b.ctypes.data))
        # the copy is asynchronous: keep this batch's arrays AND the previous one's alive (the library only guarantees that
        # the upload before last has completed when b2a_submit_batch returns)
        self._keep = ([x, b] + getattr(self, "_keep", []))[:8]
        return K

    def centers(self, grp, com=0, frame=0):
        n = max(self.nmol(grp), 0)
        out = np.empty((3, n), dtype=np.float64)
        check(self.L.b2a_centers(self._h, grp, com, frame, out.ctypes.data))
        return out.T            # [nmol, 3], column-major memory like itp::matd

    def centers_device(self, grp, com=0) -> int:
        """Queue the centre kernel for the current batch and return the device address of [K][3][nmol] doubles (no copy)."""
        p = _P()
        check(self.L.b2a_centers_device(self._h, grp, com, C.byref(p)))
        return int(p.value or 0)

    def centers_batch(self, grp, com=0):
        n, K = max(self.nmol(grp), 0), max(self.L.b2a_batch_frames(self._h), 0)
        out = np.empty((K, 3, n), dtype=np.float64)
        check(self.L.b2a_centers_batch(self._h, grp, com, out.ctypes.data))
        return out.transpose(0, 2, 1)

    def positions(self, grp, napm, frame=0):
        n = max(self.nmol(grp), 0)
        out = np.empty((napm, n, 3), dtype=np.float64)
        check(self.L.b2a_positions(self._h, grp, frame, out.ctypes.data))
        return out.transpose(1, 0, 2)   # [nmol, napm, 3]

    # ---- velocities / forces -----------------------------------------------------------------------------
    def submit_field(self, field, data):
        """field: FIELD_V / FIELD_F; data [K, natoms, 3] float32 for the frames of the current batch."""
        d = np.ascontiguousarray(data, dtype=np.float32)
        if d.ndim == 2:
            d = d[None]
        self._keep.append(d)
        check(self.L.b2a_submit_field(self._h, field, d.shape[0], d.ctypes.data))

    def field_centers(self, field, grp, com=0, frame=0):
        n = max(self.nmol(grp), 0)
        out = np.empty((3, n), dtype=np.float64)
        check(self.L.b2a_field_centers(self._h, field, grp, com, frame, out.ctypes.data))
        return out.T

    def field_centers_batch(self, field, grp, com=0):
        n, K = max(self.nmol(grp), 0), max(self.L.b2a_batch_frames(self._h), 0)
        out = np.empty((K, 3, n), dtype=np.float64)
        check(self.L.b2a_field_centers_batch(self._h, field, grp, com, out.ctypes.data))
        return out.transpose(0, 2, 1)

    def field_values(self, field, grp, napm, frame=0):
        n = max(self.nmol(grp), 0)
        out = np.empty((napm, n, 3), dtype=np.float64)
        check(self.L.b2a_field_values(self._h, field, grp, frame, out.ctypes.data))
        return out.transpose(1, 0, 2)

    # ---- accumulators ----------------------------------------------------------------------------------
    def density_create(self, **kw):
        s = DensitySpec(**kw)
        a = C.c_int()
        check(self.L.b2a_density_create(self._h, C.byref(s), C.byref(a)))
        return a.value

    def region_read_frames(self, acc, nframes):
        n = self.acc_size(acc)
        out = np.zeros((nframes, n), dtype=np.uint64)
        check(self.L.b2a_region_read_frames(self._h, acc, out.ctypes.data))
        return out

    def region_create(self, grp, com=0, filters=(), wrap=0, cyl=None, bins=(), per_frame=0):
        """filters: (dim, low, up, inclusive) x 0..3; cyl: (cx, cy, R2) or None; bins: (dim, low, dbin, n) x 0..2."""
        s = RegionSpec()
        s.grp, s.com, s.nfilter, s.wrap, s.nbdim, s.per_frame = grp, com, len(filters), int(wrap), len(bins), int(per_frame)
        for q, (dim, lo, up, incl) in enumerate(filters):
            s.fdim[q], s.flow[q], s.fup[q], s.fincl[q] = dim, lo, up, int(incl)
        if cyl is not None:
            s.cyl, s.cx, s.cy, s.R2 = 1, cyl[0], cyl[1], cyl[2]
        for b, (dim, lo, dbin, n) in enumerate(bins):
            s.bdim[b], s.blow[b], s.dbin[b], s.nb[b] = dim, lo, dbin, n
        a = C.c_int()
        check(self.L.b2a_region_create(self._h, C.byref(s), C.byref(a)))
        return a.value

    def rdf_create(self, **kw):
        s = RdfSpec(**kw)
        a = C.c_int()
        check(self.L.b2a_rdf_create(self._h, C.byref(s), C.byref(a)))
        return a.value

    def orient_create(self, **kw):
        s = OrientSpec(**kw)
        a = C.c_int()
        check(self.L.b2a_orient_create(self._h, C.byref(s), C.byref(a)))
        return a.value

    def coord_create(self, **kw):
        kw.setdefault("use_cyl", 0)
        for k in ("cx", "cy", "Rlow", "Rup"):
            kw.setdefault(k, 0.0)
        s = CoordSpec(**kw)
        a = C.c_int()
        check(self.L.b2a_coord_create(self._h, C.byref(s), C.byref(a)))
        self._coord_max = getattr(self, "_coord_max", {})
        self._coord_max[a.value] = (kw["max_count"], kw["grp0"])
        return a.value

    def coord_read_frames(self, acc):
        """-> (table [frames, max_count] uint32, numA [frames] uint32) of the last accumulated batch."""
        K = max(self.L.b2a_batch_frames(self._h), 0)
        mc = self._coord_max[acc][0]
        table = np.zeros((K, mc), dtype=np.uint32)
        numA = np.zeros(K, dtype=np.uint32)
        check(self.L.b2a_coord_read_frames(self._h, acc, table.ctypes.data, numA.ctypes.data))
        return table, numA

    def coord_counts(self, acc, frame):
        out = np.empty(max(self.nmol(self._coord_max[acc][1]), 0), dtype=np.int32)
        check(self.L.b2a_coord_counts(self._h, acc, frame, out.ctypes.data))
        return out

    def rdf_tune(self, acc, symmetric=0, paranoid=0):
        check(self.L.b2a_rdf_tune(self._h, acc, symmetric, paranoid))

    def rdf_cells(self, acc, mode):
        check(self.L.b2a_rdf_cells(self._h, acc, mode))

    def acc_fast(self, acc, mode):
        """density / orient: 0 literal FP64, 1 FP32 bracket + FP64 queue (default), 2 validate (STAT_MISPROVEN)."""
        check(self.L.b2a_acc_fast(self._h, acc, mode))

    def acc_also_centers(self, acc, com):
        """orient: the same pass also writes the exact centres of kind `com` (cached for the batch)."""
        check(self.L.b2a_acc_also_centers(self._h, acc, com))

    def invalidate(self):
        """Drop cached centres so the next accumulate redoes the whole path on the resident batch."""
        check(self.L.b2a_invalidate(self._h))

    def timing_enable(self, on=True):
        check(self.L.b2a_timing_enable(self._h, int(on)))

    def timing_read(self):
        """-> dict(kernel -> (total_ms, launches)) measured with CUDA events on the context stream; resets."""
        ms = np.zeros(8, dtype=np.float64)
        n = np.zeros(8, dtype=np.int64)
        check(self.L.b2a_timing_read(self._h, ms.ctypes.data, n.ctypes.data))
        names = ["k1_centers", "k3_prep", "k3_rdf", "k2_density", "k4_orient", "h2d", "k5_coord", "other7"]
        return {k: (float(ms[i]), int(n[i])) for i, k in enumerate(names)}

    def accumulate(self, acc):
        check(self.L.b2a_accumulate(self._h, acc))

    def reset(self, acc):
        check(self.L.b2a_acc_reset(self._h, acc))

    def acc_size(self, acc):
        n = C.c_size_t()
        check(self.L.b2a_acc_size(self._h, acc, C.byref(n)))
        return n.value

    def acc_device_ptr(self, acc):
        p = _P()
        check(self.L.b2a_acc_device_ptr(self._h, acc, C.byref(p)))
        return int(p.value)

    def read(self, acc, ncounters=None):
        """b2a_acc_read always writes the accumulator's public layout; the buffer is sized from the library, never from the
        caller (`ncounters`, kept for older call sites, only slices the result)."""
        full = self._public_size(acc)
        counts = np.zeros(full, dtype=np.uint64)
        stats = np.zeros(STAT_NSLOTS, dtype=np.uint64)
        check(self.L.b2a_acc_read(self._h, acc, counts.ctypes.data, stats.ctypes.data))
        return (counts if ncounters is None else counts[:ncounters]), stats

    def _public_size(self, acc):
        n = C.c_size_t()
        check(self.L.b2a_acc_public_size(self._h, acc, C.byref(n)))
        return n.value


def comm_unique_id() -> bytes:
    """128-byte id of a new multi-process communicator (rank 0 calls this and broadcasts it through the launcher's channel)."""
    buf = C.create_string_buffer(128)
    check(load().b2a_comm_unique_id(C.cast(buf, _P)))
    return buf.raw


def coord_fold(table, numA, tot=None):
    """totPair[k] += tmpPair[k] / numA frame by frame (calc_pair_dist_bulk.cpp:89-92)."""
    table = np.ascontiguousarray(table, dtype=np.uint32)
    numA = np.ascontiguousarray(numA, dtype=np.uint32)
    if tot is None:
        tot = np.zeros(table.shape[1], dtype=np.float64)
    check(load().b2a_coord_fold(table.ctypes.data, numA.ctypes.data, table.shape[0], table.shape[1], tot.ctypes.data))
    return tot


def rdf_normalise(counts, nbin, Rbin):
    counts = np.ascontiguousarray(counts, dtype=np.uint64)
    g = np.empty(nbin * Rbin, dtype=np.float64)
    with np.errstate(all="ignore"):
        check(load().b2a_rdf_normalise(counts.ctypes.data, nbin, Rbin, g.ctypes.data))
    return g.reshape(Rbin, nbin)
