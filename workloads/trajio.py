"""Writers for the gmxstub file formats (gmxstub/README.md): raw-float trajectory, binary topology, .ndx."""
from __future__ import annotations

import struct

import numpy as np


def write_trajectory(path, frames, box9, times=None, velocities=None, forces=None):
    """frames: iterable of [natoms,3] float32 arrays (or one [nframes,natoms,3] array); velocities / forces: optional
    arrays of the same shape (the v / f blocks of a TRR-like trajectory)."""
    frames = list(frames) if not isinstance(frames, np.ndarray) else [frames[k] for k in range(frames.shape[0])]
    nat = frames[0].shape[0]
    box9 = np.asarray(box9, dtype=np.float32).reshape(-1, 9)
    flags = 1 | (2 if velocities is not None else 0) | (4 if forces is not None else 0)
    with open(path, "wb") as fp:
        fp.write(b"B2ATRJ01")
        fp.write(struct.pack("<iiii", nat, len(frames), flags, 0))
        for k, x in enumerate(frames):
            t = float(times[k]) if times is not None else float(k)
            fp.write(struct.pack("<f", t))
            fp.write(np.ascontiguousarray(box9[min(k, box9.shape[0] - 1)], dtype="<f4").tobytes())
            fp.write(np.ascontiguousarray(x, dtype="<f4").tobytes())
            if velocities is not None:
                fp.write(np.ascontiguousarray(velocities[k], dtype="<f4").tobytes())
            if forces is not None:
                fp.write(np.ascontiguousarray(forces[k], dtype="<f4").tobytes())


def write_topology(path, system):
    m, q, resind, typ, names, mols = system.atom_tables()
    ntypes = int(typ.max()) + 1 if len(typ) else 1
    with open(path, "wb") as fp:
        fp.write(b"B2ATOP01")
        fp.write(struct.pack("<iiii", len(m), len(mols) - 1, ntypes, 0))
        rec = np.zeros(len(m), dtype=[("m", "<f4"), ("q", "<f4"), ("resind", "<i4"), ("type", "<i4"), ("name", "S8")])
        rec["m"], rec["q"], rec["resind"], rec["type"] = m, q, resind, typ
        rec["name"] = [n.encode() for n in names]
        fp.write(rec.tobytes())
        fp.write(mols.astype("<i4").tobytes())
        lj = np.zeros((ntypes * ntypes, 2), dtype="<f4")
        lj[:, 0], lj[:, 1] = 1e-3, 1e-6
        fp.write(lj.tobytes())


def write_index(path, groups):
    with open(path, "w") as fp:
        for name, idx in groups.items():
            fp.write(f"[ {name} ]\n")
            idx1 = np.asarray(idx, dtype=np.int64) + 1
            for s in range(0, len(idx1), 15):
                fp.write(" ".join(str(v) for v in idx1[s:s + 15]) + "\n")


def write_system(prefix, system, nframes, groups=None, f0=0, velocities=None, forces=None):
    """Writes prefix.b2t / prefix.b2p / prefix.ndx; returns the three paths."""
    trj, top, ndx = prefix + ".b2t", prefix + ".b2p", prefix + ".ndx"
    write_trajectory(trj, system.frames(f0, nframes), system.box9(), velocities=velocities, forces=forces)
    write_topology(top, system)
    write_index(ndx, groups if groups is not None else system.groups)
    return trj, top, ndx
