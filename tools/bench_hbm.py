"""HBM-roofline measurement of the streaming kernels (K1 centres, K1b positions, K2 density, K4 orientation) on a
config-4-sized water box (166 666 SPC/E = 499 998 atoms), K frames resident.  Algorithmic bytes:
12 B read per atom-frame (+ outputs that are materialised).  Prints one JSON line per kernel.

    python tools/bench_hbm.py [--frames 64] [--reps 20]
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
    ap.add_argument("--frames", type=int, default=64)
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--nmol", type=int, default=166666)
    args = ap.parse_args()
    peaks = {}
    pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pk):
        peaks = json.load(open(pk))
    hbm = peaks.get("hbm_gbs", 6650.0)
    peak_src = "measured (MEASURED_PEAKS.json)" if "hbm_gbs" in peaks else "fallback 6.65 TB/s"

    K, nmol = args.frames, args.nmol
    L = float((nmol / 33.4) ** (1.0 / 3.0))
    system = synth.water_box(nmol, L, seed=20200104)
    nat = system.natoms
    x = np.empty((K, nat, 3), np.float32)
    base = system.frames(0, min(K, 4))
    for k in range(K):                      # distinct frames are irrelevant for a streaming kernel: reuse 4, rolled
        x[k] = np.roll(base[k % base.shape[0]], k, axis=0) if k >= base.shape[0] else base[k]
    ctx = lib.Context(nat, K)
    sp = system.species[0]
    ctx.set_group(0, system.groups["SOL"], 3, sp.mass.astype(np.float64), sp.charge.astype(np.float64))
    ctx.submit(x, system.box9())
    ctx.sync()

    den = ctx.density_create(grp=0, com=0, dim=2, low=0.0, up=L, dbin=L / 100, nbin=100, upper_inclusive=0, dim2=-1, low2=0.0, up2=0.0)
    ori = ctx.orient_create(grp=0, com=0, DIMN=2, low=-1.0, up=100.0, c1=0.0, c2=0.0, dr=100.0, CR=100.0, Cr=0.0, nAngle=90,
                            tp1=1, tp2=2, tp3=3, DIMNOrient=2, method=1)

    def timed(fn, slot):
        for _ in range(3):
            fn()
        ctx.sync()
        ctx.timing_enable(True)
        ctx.timing_read()
        for _ in range(args.reps):
            fn()
        t = ctx.timing_read()[slot]
        ctx.timing_enable(False)
        return t[0] / max(t[1], 1)

    def k1():
        ctx.invalidate()
        ctx.centers(0, 0, 0)                # K1 over all K frames (+ one small D2H of frame 0)

    rows = []
    ms = timed(k1, "k1_centers")
    rows.append(("k1_centers (mass centres, matd + fixed point written)", ms, K * (nat * 12 + nmol * (24 + 16))))
    for mode, tag in ((1, "FP32 bracket + FP64 queue"), (0, "literal FP64")):
        for acc, slot, what in ((den, "k2_density", "k2_density (centre z -> 100-bin histogram, fused)"),
                                (ori, "k4_orient", "k4_orient (centre + bisector angle histogram, fused)")):
            ctx.acc_fast(acc, mode)
            ctx.reset(acc)
            ms = timed(lambda: ctx.accumulate(acc), slot)
            cnt, st = ctx.read(acc, ctx.acc_size(acc))
            rows.append((f"{what} [{tag}; {int(st[lib.STAT_EXACT])} of {int(st[lib.STAT_FRAMES]) * nmol} molecules re-evaluated in FP64]",
                         ms, K * nat * 12))
    for name, ms, nbytes in rows:
        gbs = nbytes / (ms * 1e-3) / 1e9
        print(json.dumps({"kernel": name, "ms_per_launch": ms, "frames": K, "atoms": nat, "algorithmic_bytes": nbytes,
                          "achieved_GBs": gbs, "peak_GBs": hbm, "peak_source": peak_src, "frac": gbs / hbm,
                          "frames_per_s": K / (ms * 1e-3)}))
    ctx.close()


if __name__ == "__main__":
    main()
