"""Randomised bit-exactness of the loaders against the oracle (test infrastructure: this tool, like tests/, may call oracle/).
Random molecule sizes (1..40 atoms: thread-per-molecule and dimension-parallel streaming K1, direct kernels for index groups
that are not one contiguous run), random masses / charges, cubic and elongated boxes, atoms pushed across faces and whole boxes
away; centres for com 0 / 1 / 2 over the whole batch and unwrapped positions of one frame must equal the oracle's doubles bit for
bit.  One line per case.

    python tests/stress_loaders.py [--cases 60] [--seed 1]
"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gmx2020postanalysis_b200 import lib  # noqa: E402
from workloads import synth  # noqa: E402
from oracle import port  # noqa: E402


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--cases", type=int, default=60)
    ap.add_argument("--seed", type=int, default=1)
    a = ap.parse_args(argv)
    rng = np.random.default_rng(a.seed)
    bad = 0
    for case in range(a.cases):
        napm = int(rng.integers(1, 41))
        nmol = int(rng.integers(50, 20000 // napm + 60))
        box = rng.uniform(2.5, 9.0, 3).astype(np.float32)
        if rng.random() < 0.5:
            box[:] = box[0]
        K = int(rng.integers(1, 6))
        mass = rng.uniform(1.0, 30.0, napm).astype(np.float32)
        charge = rng.uniform(-1.0, 1.0, napm).astype(np.float32)
        if abs(float(charge.astype(np.float64).sum())) < 0.05:
            charge[0] += np.float32(0.5)
        sp = synth.Species("MOL", nmol, synth._blob_template(napm, float(rng.uniform(0.05, 0.5)), 9500 + case), mass, charge,
                           [f"A{j}" for j in range(napm)])
        system = synth.System([sp], np.array([box[0]] * 3, np.float32), 6000 + case)
        x = system.frames(0, K).copy() * (box / box[0])
        sel = rng.random(x.shape[:2]) < 0.03
        x[sel] += (rng.integers(-3, 4, (int(sel.sum()), 3)) * box).astype(np.float32)
        box9 = np.tile(np.diag(box).reshape(1, 9), (K, 1)).astype(np.float32)
        nat = system.natoms
        index = np.arange(nat, dtype=np.int32)
        contiguous = rng.random() < 0.7
        if not contiguous:                                   # whole molecules, but in shuffled order: the direct-load kernels
            order = rng.permutation(nmol)
            index = (order[:, None] * napm + np.arange(napm)[None, :]).astype(np.int32).ravel()
        ctx = lib.Context(nat, K)
        ctx.set_group(0, index, napm, mass.astype(np.float64), charge.astype(np.float64))
        ctx.submit(x, box9)
        ok = True
        for com, w in ((0, mass.astype(np.float64)), (1, np.ones(napm)), (2, charge.astype(np.float64))):
            got = ctx.centers_batch(0, com)
            for f in range(K):
                exp = port.load_position_center(x[f], index, napm, w, port.lbox_from_box9(box9[f]))
                ok = ok and np.array_equal(got[f], exp)
        f = int(rng.integers(0, K))
        gp = ctx.positions(0, napm, f)
        ep = port.load_position(x[f], index, napm, port.lbox_from_box9(box9[f]))
        ok = ok and np.array_equal(gp, ep)
        bad += 0 if ok else 1
        print(f"case {case:3d} napm {napm:2d} nmol {nmol:5d} K {K} box {box[0]:.2f} {box[1]:.2f} {box[2]:.2f} {'contiguous' if contiguous else 'shuffled  '} "
              f"{'ok' if ok else 'FAIL'}", flush=True)
        ctx.close()
    print(f"{a.cases} cases, {bad} failures")
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
