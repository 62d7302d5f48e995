"""Randomised self-consistency of K3: for random systems, boxes, cut-offs and region filters the histogram must be identical
(i) with the FP32 bracket and with every pair re-evaluated in FP64 (paranoid), (ii) all-pairs and cell-binned, (iii) full and
symmetric evaluation.  The paranoid kernel is the reference's arithmetic (pinned against the oracle in tests/); this tool widens
the input space far beyond the fixtures.  One line per case.

    python tests/stress_rdf.py [--cases 40] [--seed 1]
"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import refprog  # noqa: E402
from gmx2020postanalysis_b200 import lib  # noqa: E402
from workloads import synth  # noqa: E402


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--cases", type=int, default=40)
    ap.add_argument("--seed", type=int, default=1)
    a = ap.parse_args(argv)
    rng = np.random.default_rng(a.seed)
    bad = 0
    for case in range(a.cases):
        same = rng.random() < 0.5
        if same:
            n = int(rng.integers(500, 30000))
            L0 = float((n / rng.uniform(10, 40)) ** (1 / 3))
            system = synth.water_box(n, L0, seed=7000 + case, split_atoms=True)
            g = ("OW", "OW")
        else:
            n = int(rng.integers(200, 6000))
            L0 = float((n / rng.uniform(1.5, 6)) ** (1 / 3))
            system = synth.ionic_liquid(n, L0, seed=7000 + case)
            g = ("CAT", "ANI") if rng.random() < 0.7 else ("ANI", "CAT")
        K = int(rng.integers(1, 4))
        box = np.array([L0] * 3, np.float32)
        if rng.random() < 0.4:
            box = (box * rng.uniform(0.7, 1.4, 3)).astype(np.float32)
        x = system.frames(0, K).copy() * (box / np.float32(L0))
        box9 = np.tile(np.diag(box).reshape(1, 9), (K, 1)).astype(np.float32)
        Rm = refprog.f32(float(rng.uniform(0.08, 0.5) * box.min()))
        Rbin = int(np.float32(Rm) / 0.002)
        nbin = int(rng.integers(1, 6))
        dim, dim2 = int(rng.integers(0, 3)), int(rng.integers(0, 3))
        if rng.random() < 0.5:
            low, up, low2, up2 = -100.0, 100.0, -100.0, 100.0
        else:
            low, up = sorted(rng.uniform(-0.1, 1.1, 2) * float(box[dim]))
            low2, up2 = sorted(rng.uniform(-0.1, 1.1, 2) * float(box[dim2]))
        ctx = lib.Context(system.natoms, K)
        for k, name in enumerate(g):
            gi = refprog.group_info(system, name)
            ctx.set_group(k, gi["index"], gi["napm"], gi["mass"], gi["charge"])
        ctx.submit(x, box9)
        res = {}
        for tag, sym, par, cells in (("fp32", 0, 0, 0), ("fp64", 0, 1, 0), ("cells", 0, 0, 1), ("sym", 1, 0, 0), ("cells+sym", 1, 0, 1)):
            if par and system.natoms > 40000:
                continue                                     # paranoid mode is ~50x slower: only on the smaller systems
            acc = ctx.rdf_create(grp0=0, grp1=1, com0=0, com1=0, nbin=nbin, low=float(low), up=float(up), dim=dim, dim2=dim2, low2=float(low2),
                                 up2=float(up2), Rm=Rm, Rbin=Rbin)
            ctx.rdf_tune(acc, sym, par)
            ctx.rdf_cells(acc, cells)
            ctx.accumulate(acc)
            res[tag] = ctx.read(acc, nbin * Rbin)
        ref = res["fp32"]
        ok = True
        for tag, (c, st) in res.items():
            ok = ok and np.array_equal(c, ref[0]) and int(st[lib.STAT_SELECTED]) == int(ref[1][lib.STAT_SELECTED]) \
                and int(st[lib.STAT_DROPPED]) == int(ref[1][lib.STAT_DROPPED])
        bad += 0 if ok else 1
        pairs = {t: int(st[lib.STAT_PAIRS]) for t, (c, st) in res.items()}
        print(f"case {case:3d} {g[0]}-{g[1]} n {n:5d} K {K} box {box[0]:.2f} {box[1]:.2f} {box[2]:.2f} Rm {Rm:.3f} nbin {nbin} counted {int(ref[0].sum())} "
              f"evaluated {pairs} {'ok' if ok else 'FAIL'}", flush=True)
        ctx.close()
    print(f"{a.cases} cases, {bad} failures")
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
