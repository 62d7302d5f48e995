"""Synthetic systems: SPC/E water boxes and a 25+7-site ionic liquid.

Test / bench harness, not product: the generator (``workloads/synth_gen.c`` -> ``libsynthgen.so``, plus a
bit-identical pure-numpy restatement that needs no compiled code -- ``B2A_SYNTH_NUMPY=1`` or a missing library
selects it; the reference arm of bench.py uses it so that it maps nothing this repo compiled) and the system
definitions (templates, masses, charges, molecule/residue tables, index groups) that the reference would read
from a .tpr/.ndx (reference include/itp/gmx_bits/gmx_handle.ipp:23-34, 419-468).
"""
from __future__ import annotations

import ctypes
import os
from dataclasses import dataclass, field

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_lib = None


LIB_PATH = os.path.join(_HERE, "libsynthgen.so")


def use_numpy() -> bool:
    return bool(int(os.environ.get("B2A_SYNTH_NUMPY", "0") or 0)) or not os.path.exists(LIB_PATH)


def _load():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} missing: run `python -c 'import __graft_entry__ as g; g.build()'`")
        lib = ctypes.CDLL(LIB_PATH)
        lib.b2a_synth_species.restype = ctypes.c_int
        lib.b2a_synth_species.argtypes = [ctypes.c_uint64, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                          ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
        lib.b2a_synth_hash.restype = ctypes.c_uint64
        lib.b2a_synth_hash.argtypes = [ctypes.c_uint64] * 5
        _lib = lib
    return _lib


# ---- pure-numpy restatement of synth_gen.c (same integer hash, same IEEE double / float operations, no FMA) --------
_U = np.uint64


def _mix64(z):
    z = (z ^ (z >> _U(30))) * _U(0xbf58476d1ce4e5b9)
    z = (z ^ (z >> _U(27))) * _U(0x94d049bb133111eb)
    return z ^ (z >> _U(31))


def _hash_np(seed, frame, species, mol, lane):
    """Vectorised b2a_synth_hash: `mol` and `lane` are uint64 arrays (or scalars), arithmetic wraps mod 2^64."""
    with np.errstate(over="ignore"):
        h = _mix64(_U(seed) + _U(0x9e3779b97f4a7c15))
        h = _mix64(h ^ (_U(frame) * _U(0xd1342543de82ef95) + _U(0x632be59bd9b4e019)))
        h = _mix64(h ^ (_U(species) * _U(0xa0761d6478bd642f) + _U(0xe7037ed1a0b428db)))
        h = _mix64(h ^ (np.asarray(mol, _U) * _U(0x8ebc6af09c88c6e3) + _U(0x589965cc75374cc3)))
        h = _mix64(h ^ (np.asarray(lane, _U) * _U(0x9e3779b97f4a7c15) + _U(0x1d8e4e27c47d124f)))
    return h


def _u01(h):
    return (h >> _U(40)).astype(np.float64) * (1.0 / 16777216.0)


def _species_numpy(seed, frame, species, nmol, tmpl, L, out):
    mol = np.arange(nmol, dtype=_U)
    c = np.empty((nmol, 3), np.float32)
    for k in range(3):
        c[:, k] = _u01(_hash_np(seed, frame, species, mol, k)).astype(np.float32) * np.float32(L[k])
    q = np.empty((nmol, 4), np.float64)
    n2 = np.empty(nmol, np.float64)
    lane = np.full(nmol, 3, dtype=_U)
    todo = np.arange(nmol)
    while todo.size:                                        # rejection from the 4-cube, per molecule
        acc = np.zeros(todo.size, np.float64)
        for a in range(4):
            qa = 2.0 * _u01(_hash_np(seed, frame, species, mol[todo], lane[todo] + _U(a))) - 1.0
            q[todo, a] = qa
            acc = acc + qa * qa
        lane[todo] += _U(4)
        n2[todo] = acc
        todo = todo[(acc > 1.0) | (acc < 1e-6)]
    inv = 1.0 / np.sqrt(n2)
    w, x, y, z = (q[:, a] * inv for a in range(4))
    R = [[1.0 - 2.0 * (y * y + z * z), 2.0 * (x * y - w * z), 2.0 * (x * z + w * y)],
         [2.0 * (x * y + w * z), 1.0 - 2.0 * (x * x + z * z), 2.0 * (y * z - w * x)],
         [2.0 * (x * z - w * y), 2.0 * (y * z + w * x), 1.0 - 2.0 * (x * x + y * y)]]
    t = tmpl.astype(np.float64)
    o = out.reshape(nmol, tmpl.shape[0], 3)
    for j in range(tmpl.shape[0]):
        for k in range(3):
            Lk = np.float32(L[k])
            p = c[:, k].astype(np.float64) + (R[k][0] * t[j, 0] + R[k][1] * t[j, 1] + R[k][2] * t[j, 2])
            f = p.astype(np.float32)
            f = np.where(f < 0, f + Lk, f)
            f = np.where(f >= Lk, f - Lk, f)
            f = np.where(f < 0, np.float32(0), f)
            f = np.where(f >= Lk, np.float32(0), f)
            o[:, j, k] = f


def synth_hash(seed, frame, species, mol, lane) -> int:
    if use_numpy():
        return int(_hash_np(seed, frame, species, mol, lane))
    return int(_load().b2a_synth_hash(seed, frame, species, mol, lane))


@dataclass
class Species:
    name: str
    nmol: int
    tmpl: np.ndarray            # [napm, 3] float32, nm
    mass: np.ndarray            # [napm] float32 (GROMACS t_atom.m is `real`)
    charge: np.ndarray          # [napm] float32
    atomnames: list
    split_atoms: bool = False   # topology lists every atom as its own molecule/residue (O-O RDF)

    @property
    def napm(self) -> int:
        return int(self.tmpl.shape[0])


@dataclass
class System:
    species: list
    L: np.ndarray               # [3] float32 box lengths
    seed: int = 20200101
    groups: dict = field(default_factory=dict)   # name -> int32 atom indices (0-based)

    @property
    def natoms(self) -> int:
        return sum(s.nmol * s.napm for s in self.species)

    def offsets(self):
        off, out = 0, []
        for s in self.species:
            out.append(off)
            off += s.nmol * s.napm
        return out

    def frame(self, f: int, out: np.ndarray | None = None) -> np.ndarray:
        if out is None:
            out = np.empty((self.natoms, 3), dtype=np.float32)
        assert out.dtype == np.float32 and out.flags.c_contiguous and out.shape == (self.natoms, 3)
        L = np.ascontiguousarray(self.L, dtype=np.float32)
        lib = None if use_numpy() else _load()
        for sid, (s, off) in enumerate(zip(self.species, self.offsets())):
            tm = np.ascontiguousarray(s.tmpl, dtype=np.float32)
            view = out[off:off + s.nmol * s.napm]
            if lib is None:
                _species_numpy(self.seed, f, sid, s.nmol, tm, L, view)
                continue
            rc = lib.b2a_synth_species(self.seed, f, sid, s.nmol, s.napm, tm.ctypes.data, L.ctypes.data,
                                       view.ctypes.data)
            if rc != 0:
                raise RuntimeError("b2a_synth_species failed")
        return out

    def frames(self, f0: int, n: int, out: np.ndarray | None = None) -> np.ndarray:
        if out is None:
            out = np.empty((n, self.natoms, 3), dtype=np.float32)
        for k in range(n):
            self.frame(f0 + k, out[k])
        return out

    def box9(self) -> np.ndarray:
        b = np.zeros(9, dtype=np.float32)
        b[0], b[4], b[8] = self.L
        return b

    # ---- per-atom topology tables (what read_tpx_top would deliver) -------------------------------
    def atom_tables(self):
        m, q, resind, typ, names, mols = [], [], [], [], [], [0]
        res = 0
        for sid, s in enumerate(self.species):
            for _ in range(s.nmol):
                for j in range(s.napm):
                    m.append(s.mass[j]); q.append(s.charge[j]); typ.append(sid); names.append(s.atomnames[j])
                    if s.split_atoms:
                        resind.append(res); res += 1
                        mols.append(mols[-1] + 1)
                    else:
                        resind.append(res)
                if not s.split_atoms:
                    res += 1
                    mols.append(mols[-1] + s.napm)
        return (np.asarray(m, np.float32), np.asarray(q, np.float32), np.asarray(resind, np.int32),
                np.asarray(typ, np.int32), names, np.asarray(mols, np.int32))


def group_info(system: System, name: str) -> dict:
    """napm / weights of an index group that is a whole species or the oxygen subset of split-atom water: what
    GmxHandle::init derives from the topology (get_natom_per_mol / get_mass / get_charge, reference ipp:419-468)."""
    idx = np.asarray(system.groups[name], dtype=np.int32)
    for s, off in zip(system.species, system.offsets()):
        n = s.nmol * s.napm
        if idx[0] >= off and idx[-1] < off + n:
            if s.split_atoms:
                j = int((idx[0] - off) % s.napm)
                m = np.array([s.mass[j]], np.float64); q = np.array([s.charge[j]], np.float64)
                return {"index": idx, "napm": 1, "w": {0: m, 1: np.ones(1), 2: q}, "mass": m, "charge": q}
            assert len(idx) == n, "group must be a whole species"
            m = s.mass.astype(np.float64); q = s.charge.astype(np.float64)
            return {"index": idx, "napm": s.napm, "w": {0: m, 1: np.ones(s.napm), 2: q}, "mass": m, "charge": q}
    raise ValueError(name)


# SPC/E: O-H 0.1 nm, H-O-H 109.47 deg
_SPCE_TMPL = np.array([[0.0, 0.0, 0.0],
                       [0.08164904, 0.05773596, 0.0],
                       [-0.08164904, 0.05773596, 0.0]], dtype=np.float32)
_SPCE_MASS = np.array([15.9994, 1.008, 1.008], dtype=np.float32)
_SPCE_Q = np.array([-0.8476, 0.4238, 0.4238], dtype=np.float32)


def water_box(nmol: int, L: float, seed: int = 20200101, split_atoms: bool = False) -> System:
    """SPC/E box.  Groups: SOL (all atoms), OW (oxygens).  ``split_atoms`` gives the topology in which every
    atom is its own molecule so an O-only group passes GmxHandle::init(true) (option 1)."""
    sp = Species("SOL", nmol, _SPCE_TMPL, _SPCE_MASS, _SPCE_Q, ["OW", "HW1", "HW2"], split_atoms)
    sysm = System([sp], np.array([L, L, L], np.float32), seed)
    n = 3 * nmol
    sysm.groups = {"SOL": np.arange(n, dtype=np.int32), "OW": np.arange(0, n, 3, dtype=np.int32)}
    return sysm


def _blob_template(napm: int, radius: float, key: int) -> np.ndarray:
    """Deterministic rigid template: napm sites inside a sphere of the given radius (extent <= 2*radius)."""
    pts, lane = [], 0
    while len(pts) < napm:
        p = np.array([2.0 * ((synth_hash(key, 0, 0, len(pts), lane + a) >> 40) / 16777216.0) - 1.0 for a in range(3)])
        lane += 3
        if p @ p <= 1.0:
            pts.append(p * radius)
    return np.asarray(pts, dtype=np.float32)


def ionic_liquid(npairs: int, L: float, seed: int = 20200103) -> System:
    """31 250 x (25-site cation + 7-site anion) = 1M atoms at npairs=31250.
    Cation total charge +0.8, anion -0.8, molecular extent <= 0.6 nm."""
    cm = np.array(([12.011, 1.008, 1.008, 14.007, 12.011] * 5), dtype=np.float32)
    cq = np.full(25, 0.8 / 25, dtype=np.float32)
    cq[::5] += 0.05; cq[1::5] -= 0.05          # non-uniform but still +0.8 overall (to float rounding)
    am = np.array([10.811, 18.998, 18.998, 18.998, 18.998, 12.011, 14.007], dtype=np.float32)
    aq = np.array([0.9, -0.45, -0.45, -0.45, -0.45, 0.2, -0.1], dtype=np.float32)
    cat = Species("CAT", npairs, _blob_template(25, 0.3, 7001), cm, cq, [f"C{j}" for j in range(25)])
    ani = Species("ANI", npairs, _blob_template(7, 0.2, 7002), am, aq, [f"A{j}" for j in range(7)])
    sysm = System([cat, ani], np.array([L, L, L], np.float32), seed)
    nc = 25 * npairs
    sysm.groups = {"CAT": np.arange(nc, dtype=np.int32), "ANI": np.arange(nc, nc + 7 * npairs, dtype=np.int32),
                   "System": np.arange(nc + 7 * npairs, dtype=np.int32)}
    return sysm


def mixed_box(n_water: int = 1200000, n_cat: int = 8000, n_ani: int = 8000, n_cos: int = 24000, L: float = 34.0,
              seed: int = 20200105) -> System:
    """Config C5: 1.2 M water + 8 000 cation (25 sites) + 8 000 anion (7) + 24 000 six-site cosolvent
    = 4 000 000 atoms at the defaults.  Groups: SOL, CAT, ANI, COS."""
    il = ionic_liquid(1, 1.0)
    cat_t, ani_t = il.species[0], il.species[1]
    water = Species("SOL", n_water, _SPCE_TMPL, _SPCE_MASS, _SPCE_Q, ["OW", "HW1", "HW2"])
    cat = Species("CAT", n_cat, cat_t.tmpl, cat_t.mass, cat_t.charge, cat_t.atomnames)
    ani = Species("ANI", n_ani, ani_t.tmpl, ani_t.mass, ani_t.charge, ani_t.atomnames)
    cm = np.array([12.011, 12.011, 15.9994, 1.008, 1.008, 1.008], dtype=np.float32)
    cq = np.array([0.1, -0.2, -0.5, 0.2, 0.2, 0.2], dtype=np.float32)
    cos = Species("COS", n_cos, _blob_template(6, 0.18, 7003), cm, cq, [f"S{j}" for j in range(6)])
    sysm = System([water, cat, ani, cos], np.array([L, L, L], np.float32), seed)
    off = sysm.offsets()
    sysm.groups = {s.name: np.arange(o, o + s.nmol * s.napm, dtype=np.int32) for s, o in zip(sysm.species, off)}
    return sysm
