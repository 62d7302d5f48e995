"""K1 / K2 streaming rate against atoms per molecule (shared-memory bank conflicts of the one-molecule-per-thread read:
stride 3 * napm words, so even napm conflicts 2- to 8-way).  ~500 000 atoms, K frames resident.  One JSON line per napm.

    python tools/bench_hbm_napm.py [--frames 32] [--napm 2,3,4,6,7,8,25]
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gmx2020postanalysis_b200 import lib  # noqa: E402
from workloads import synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=32)
    ap.add_argument("--reps", type=int, default=10)
    ap.add_argument("--napm", default="2,3,4,6,7,8,25")
    a = ap.parse_args()
    pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
    hbm = json.load(open(pk)).get("hbm_gbs", 6650.0) if os.path.exists(pk) else 6650.0
    for napm in [int(v) for v in a.napm.split(",")]:
        nmol = 500000 // napm
        rng = np.random.default_rng(napm)
        sp = synth.Species("MOL", nmol, synth._blob_template(napm, 0.2, 7100 + napm), rng.uniform(1, 16, napm).astype(np.float32),
                           rng.uniform(-0.5, 0.5, napm).astype(np.float32), [f"A{j}" for j in range(napm)])
        L = 17.0
        system = synth.System([sp], np.array([L, L, L], np.float32), 4400 + napm)
        nat = system.natoms
        K = a.frames
        base = system.frames(0, 2)
        x = np.empty((K, nat, 3), np.float32)
        for k in range(K):
            x[k] = base[k % 2]
        ctx = lib.Context(nat, K)
        ctx.set_group(0, np.arange(nat, dtype=np.int32), napm, sp.mass.astype(np.float64), sp.charge.astype(np.float64))
        ctx.submit(x, system.box9())
        ctx.sync()
        den = ctx.density_create(grp=0, com=0, dim=2, low=0.0, up=L, dbin=L / 100, nbin=100, upper_inclusive=0, dim2=-1, low2=0.0, up2=0.0)

        def timed(fn, slot):
            for _ in range(3):
                fn()
            ctx.sync(); ctx.timing_enable(True); ctx.timing_read()
            for _ in range(a.reps):
                fn()
            t = ctx.timing_read()[slot]
            ctx.timing_enable(False)
            return t[0] / max(t[1], 1)

        def k1():
            ctx.invalidate(); ctx.centers(0, 0, 0)
        ms1 = timed(k1, "k1_centers")
        ms2 = timed(lambda: ctx.accumulate(den), "k2_density")
        b1, b2 = K * (nat * 12 + nmol * 40), K * nat * 12
        row = {"napm": napm, "nmol": nmol, "k1_ms": ms1, "k1_frac": b1 / (ms1 * 1e-3) / 1e9 / hbm, "k2_ms": ms2,
               "k2_frac": b2 / (ms2 * 1e-3) / 1e9 / hbm}
        if napm >= 3:
            # K4 on the generic path: the vector atoms are NOT (reference atom, the other two), so nothing is reused
            ori = ctx.orient_create(grp=0, com=0, DIMN=2, low=-1.0, up=100.0, c1=0.0, c2=0.0, dr=100.0, CR=100.0, Cr=0.0, nAngle=90,
                                    tp1=2, tp2=1, tp3=3, DIMNOrient=2, method=1)
            ms4 = timed(lambda: ctx.accumulate(ori), "k4_orient")
            row.update(k4_ms=ms4, k4_frac=b2 / (ms4 * 1e-3) / 1e9 / hbm)
        print(json.dumps(row))
        ctx.close()


if __name__ == "__main__":
    main()
