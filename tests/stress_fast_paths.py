"""Randomised validation of the FP32-bracketed fast paths (K2 density, K4 orientation: generic, two-molecule packed, three-site
reuse).  Every case builds a random molecule (2..14 atoms, random masses / charges), a random (possibly elongated) box, frames
with atoms pushed across box faces and whole boxes away, random filters, bins and vector atoms, and runs the accumulator in
mode 0 (literal FP64), 1 (bracket + queue) and 2 (validation: both, counting FP32-proven answers the FP64 code contradicts).
Pass = identical histograms and statistics in all three modes and zero contradictions.  One line per case, summary at the end.

    python tests/stress_fast_paths.py [--cases 60] [--seed 1]
"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gmx2020postanalysis_b200 import lib  # noqa: E402
from workloads import synth  # noqa: E402


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--cases", type=int, default=60)
    ap.add_argument("--seed", type=int, default=1)
    a = ap.parse_args(argv)
    rng = np.random.default_rng(a.seed)
    bad = 0
    for case in range(a.cases):
        napm = int(rng.integers(2, 15))
        nmol = int(rng.integers(3000, 30000))
        box = rng.uniform(3.0, 9.0, 3).astype(np.float32)
        if rng.random() < 0.5:
            box[:] = box[0]
        K = int(rng.integers(1, 5))
        mass = rng.uniform(1.0, 30.0, napm).astype(np.float32)
        charge = rng.uniform(-1.0, 1.0, napm).astype(np.float32)
        sp = synth.Species("MOL", nmol, synth._blob_template(napm, float(rng.uniform(0.05, 0.4)), 9000 + case), mass, charge,
                           [f"A{j}" for j in range(napm)])
        system = synth.System([sp], np.array([box[0]] * 3, np.float32), 5000 + case)
        x = system.frames(0, K).copy() * (box / box[0])
        # atoms across faces / whole boxes away, a frame far outside [0, L), coordinates snapped onto bin edges
        if rng.random() < 0.7:
            sel = rng.random(x.shape[:2]) < 0.02
            x[sel] += (rng.integers(-2, 3, (int(sel.sum()), 3)) * box).astype(np.float32)
        if rng.random() < 0.3:
            x[0] += (3.0 * box).astype(np.float32)
        if rng.random() < 0.3:
            d = int(rng.integers(0, 3))
            x[:, :, d] = np.round(x[:, :, d] / (box[d] / 32)) * np.float32(box[d] / 32)
        box9 = np.tile(np.diag(box).reshape(1, 9), (K, 1)).astype(np.float32)
        nat = system.natoms
        ctx = lib.Context(nat, K)
        ctx.set_group(0, np.arange(nat, dtype=np.int32), napm, mass.astype(np.float64), charge.astype(np.float64))
        ctx.submit(x, box9)
        com = int(rng.integers(0, 2))                       # mass or geometry (charge centres of random molecules may be ill-conditioned)
        specs = []
        dim = int(rng.integers(0, 3))
        lo, up = sorted(rng.uniform(-0.2, 1.2, 2) * box[dim])
        nb = int(rng.integers(3, 200))
        specs.append(("density", dict(grp=0, com=com, dim=dim, low=float(lo), up=float(up), dbin=float(np.float32(np.float32(up - lo) / np.float32(nb))),
                                       nbin=nb, upper_inclusive=int(rng.integers(0, 2)), dim2=int(rng.integers(-1, 3)), low2=float(0.1 * box.min()),
                                       up2=float(0.9 * box.max()))))
        if napm >= 3:
            D = int(rng.integers(0, 3))
            oth = [k for k in range(3) if k != D]
            tp = rng.choice(np.arange(1, napm + 1), 3, replace=False)
            if rng.random() < 0.4:
                tp = np.array([1, 2, 3]) if rng.random() < 0.5 else np.array([1, 3, 2])
            lo, up = sorted(rng.uniform(0.0, 1.0, 2) * box[D])
            specs.append(("orient", dict(grp=0, com=com, DIMN=D, low=float(lo), up=float(up), c1=float(0.5 * box[oth[0]]), c2=float(0.5 * box[oth[1]]),
                                          dr=float(rng.uniform(0.03, 0.5)), CR=float(rng.uniform(0.3, 0.6) * min(box[oth[0]], box[oth[1]])),
                                          Cr=float(rng.uniform(0.0, 0.2)), nAngle=int(rng.choice([18, 45, 90, 180])), tp1=int(tp[0]), tp2=int(tp[1]),
                                          tp3=int(tp[2]), DIMNOrient=int(rng.integers(0, 3)), method=int(rng.integers(1, 3)))))
        line = f"case {case:3d} napm {napm:2d} nmol {nmol:5d} K {K} box {box[0]:.2f} {box[1]:.2f} {box[2]:.2f}"
        for kind, spec in specs:
            res = []
            for mode in (0, 1, 2):
                acc = ctx.density_create(**spec) if kind == "density" else ctx.orient_create(**spec)
                ctx.acc_fast(acc, mode)
                ctx.accumulate(acc)
                res.append(ctx.read(acc, ctx.acc_size(acc)))
            (c0, s0), (c1, s1), (c2, s2) = res
            ok = np.array_equal(c0, c1) and np.array_equal(c0, c2) and int(s2[lib.STAT_MISPROVEN]) == 0
            for st in (lib.STAT_DROPPED, lib.STAT_SELECTED, lib.STAT_FRAMES):
                ok = ok and int(s0[st]) == int(s1[st]) == int(s2[st])
            bad += 0 if ok else 1
            line += f" | {kind}: counted {int(c0.sum())} undecided {int(s1[lib.STAT_EXACT])} misproven {int(s2[lib.STAT_MISPROVEN])} {'ok' if ok else 'FAIL ' + repr(spec)}"
        print(line, flush=True)
        ctx.close()
    print(f"{a.cases} cases, {bad} failures")
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
