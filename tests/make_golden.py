"""Generates tests/golden/* from the reference itself (oracle/_ref: the reference's own GmxHandle loaders and
analysis programs compiled unmodified from /root/reference against gmxstub).  Runs only in the dev container;
the fixtures are committed and are what pins the oracle on boxes without /root/reference.

    python tests/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import refprog  # noqa: E402
from workloads import synth  # noqa: E402
from oracle import refshim  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
CASES = __import__("golden_cases")


def loaders():
    out = {}
    for name, system in CASES.loader_systems().items():
        x = system.frame(CASES.LOADER_FRAME)
        box = system.box9()
        for s, off in zip(system.species, system.offsets()):
            idx = np.arange(off, off + s.nmol * s.napm, dtype=np.int32)
            rh = refshim.RefHandle(idx, s.napm, s.mass, s.charge)
            rh.set_frame(x, box)
            out[f"{name}.{s.name}.pos"] = np.ascontiguousarray(rh.load_position())
            for com in (0, 1, 2):
                out[f"{name}.{s.name}.cen{com}"] = np.ascontiguousarray(rh.load_position_center(com))
    # hand-made edge fixture: molecule straddling each face, atom exactly L/2 from its reference
    x, box, idx, napm, mass, charge = CASES.edge_fixture()
    rh = refshim.RefHandle(idx, napm, mass, charge)
    rh.set_frame(x, box)
    out["edge.pos"] = np.ascontiguousarray(rh.load_position())
    for com in (0, 1, 2):
        out[f"edge.cen{com}"] = np.ascontiguousarray(rh.load_position_center(com))
    system, v, fo = CASES.field_fixture()
    for s, off in zip(system.species, system.offsets()):
        idx = np.arange(off, off + s.nmol * s.napm, dtype=np.int32)
        rh = refshim.RefHandle(idx, s.napm, s.mass, s.charge)
        rh.set_frame(system.frame(0), system.box9())
        rh.set_vf(v, fo)
        for force, tag in ((False, "vel"), (True, "force")):
            out[f"field.{s.name}.{tag}"] = np.ascontiguousarray(rh.load_field(force))
            for com in (0, 1, 2):
                out[f"field.{s.name}.{tag}.cen{com}"] = np.ascontiguousarray(rh.load_field_center(force, com))
    system, idx, napm, m, q = CASES.full_fixture()
    rf = refshim.RefHandleFull(idx, napm, m, q)
    rf.set_frame(system.frame(CASES.LOADER_FRAME), system.box9())
    out["full.pos"] = np.ascontiguousarray(rf.load_position())
    out["full.cen0"] = np.ascontiguousarray(rf.load_position_center(0))
    out["full.cen2"] = np.ascontiguousarray(rf.load_position_center(2))
    w = np.array([12.011, 1.008, 1.008, 14.007, 12.011, 15.9994, 1.008, 0.3, 7.7, 1e-3, 5.5], dtype=np.float64)
    out["eigen_sum"] = np.array([refshim.eigen_sum(w[:n]) for n in range(1, len(w) + 1)])
    pd = CASES.periodicity_inputs()
    out["periodicity"] = np.array([refshim.periodicity(a, b) for a, b in pd])
    np.savez_compressed(os.path.join(GOLD, "loaders.npz"), **out)
    print("loaders.npz:", len(out), "arrays")


def programs():
    import tempfile
    for name, case in CASES.program_cases().items():
        with tempfile.TemporaryDirectory() as td:
            lines = refprog.run_program(os.path.join(refprog.REFDIR, case["exe"]), td, case["system"](), case["nframes"],
                                        case["groups"], case["args"], "out.xvg")
        with open(os.path.join(GOLD, name + ".txt"), "w") as fp:
            fp.write("\n".join(lines) + "\n")
        print(name, len(lines), "lines")


if __name__ == "__main__":
    os.makedirs(GOLD, exist_ok=True)
    loaders()
    programs()
