"""Throughput of the other BASELINE.json configs on one GPU (the contract benchmark, bench.py, is config 2).

  C1: README template density profile, 3 000 SPC/E, 100 frames, nbin=100, dim=z, COM   (device + end-to-end)
  C3: cation-anion COM RDF, 1 000 000-atom ionic liquid (31 250 x (25 + 7) atoms)       (per-frame, K frames)
  C4: dipole vectors + bisector orientation histogram, 166 666 SPC/E (499 998 atoms)   (K frames)
Prints one JSON line per config.      python tools/bench_configs.py [--c3-frames 4] [--c4-frames 64]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gmx2020postanalysis_b200 import lib  # noqa: E402
from workloads import synth  # noqa: E402


def groups(ctx, system, names):
    for g, name in enumerate(names):
        idx = system.groups[name]
        sp = [s for s, off in zip(system.species, system.offsets()) if off <= idx[0] < off + s.nmol * s.napm][0]
        ctx.set_group(g, idx, sp.napm, sp.mass.astype(np.float64), sp.charge.astype(np.float64))


def timeit(ctx, fn, reps):
    fn(); ctx.sync()
    t = time.perf_counter()
    for _ in range(reps):
        fn()
    ctx.sync()
    return (time.perf_counter() - t) / reps


def c1(reps=20):
    system = synth.water_box(3000, 4.476, seed=20200101)
    K = 100
    x = lib.pinned_empty((K, system.natoms, 3), np.float32)
    system.frames(0, K, out=x)
    ctx = lib.Context(system.natoms, K)
    groups(ctx, system, ["SOL"])
    dbin = float(np.float32(np.float32(4.476) / np.float32(100)))
    acc = ctx.density_create(grp=0, com=0, dim=2, low=0.0, up=4.476, dbin=dbin, nbin=100, upper_inclusive=0, dim2=-1, low2=0.0, up2=0.0)
    ctx.submit(x, system.box9())
    dev = timeit(ctx, lambda: ctx.accumulate(acc), reps)

    def e2e():
        ctx.submit(x, system.box9()); ctx.accumulate(acc); ctx.read(acc, 100)
    tot = timeit(ctx, e2e, reps)
    ctx.close()
    return {"config": "C1 density template, 3000 SPC/E x 100 frames", "device_ms_per_100_frames": dev * 1e3,
            "e2e_ms_per_100_frames": tot * 1e3, "frames_per_s_device": K / dev, "frames_per_s_e2e": K / tot}


def c3(K, reps=3):
    system = synth.ionic_liquid(31250, 22.1, seed=20200103)
    x = system.frames(0, K)
    ctx = lib.Context(system.natoms, K)
    groups(ctx, system, ["CAT", "ANI"])
    Rm = float(np.float32(np.float32(22.1) / 2)); Rbin = int(Rm / 0.002)
    acc = ctx.rdf_create(grp0=0, grp1=1, com0=0, com1=0, nbin=5, low=-100.0, up=100.0, dim=0, dim2=2, low2=-100.0, up2=100.0, Rm=Rm, Rbin=Rbin)
    ctx.submit(x, system.box9())

    def step():
        ctx.invalidate(); ctx.accumulate(acc)
    dev = timeit(ctx, step, reps)
    ctx.timing_enable(True); ctx.timing_read(); step(); tm = ctx.timing_read(); ctx.timing_enable(False)
    cnt, st = ctx.read(acc, 5 * Rbin)
    ctx.close()
    pairs = 31250.0 * 31250.0 * K
    return {"config": "C3 cation-anion COM RDF, 1M-atom ionic liquid", "frames": K, "ms_per_frame": dev * 1e3 / K,
            "frames_per_s": K / dev, "pair_distances_per_s": pairs / dev, "k1_ms": tm["k1_centers"][0], "k3_ms": tm["k3_rdf"][0],
            "prep_ms": tm["k3_prep"][0], "Rbin": Rbin, "fp64_reevaluated_per_frame": int(st[lib.STAT_EXACT]) / (reps + 2) / K}


def c4(K, reps=10):
    system = synth.water_box(166666, 17.09, seed=20200104)
    base = system.frames(0, 4)
    x = lib.pinned_empty((K, system.natoms, 3), np.float32)
    for k in range(K):
        x[k] = base[k % 4]
    ctx = lib.Context(system.natoms, K)
    groups(ctx, system, ["SOL"])
    ori = ctx.orient_create(grp=0, com=0, DIMN=2, low=-1.0, up=100.0, c1=0.0, c2=0.0, dr=100.0, CR=100.0, Cr=0.0, nAngle=90,
                            tp1=1, tp2=2, tp3=3, DIMNOrient=2, method=1)
    ctx.submit(x, system.box9())

    def step():
        ctx.invalidate(); ctx.centers(0, lib.COM_DIPOLE, 0); ctx.accumulate(ori)     # dipole vectors (K1, com 3) + orientation (K4)
    dev = timeit(ctx, step, reps)

    def e2e():
        ctx.submit(x, system.box9()); ctx.centers(0, lib.COM_DIPOLE, 0); ctx.accumulate(ori); ctx.read(ori, 91)
    tot = timeit(ctx, e2e, 3)
    ctx.close()
    return {"config": "C4 dipole + orientation, 166 666 SPC/E (499 998 atoms)", "frames": K, "device_frames_per_s": K / dev,
            "e2e_frames_per_s": K / tot, "device_GBs": K * system.natoms * 12 * 2 / dev / 1e9,
            "e2e_h2d_GBs": K * system.natoms * 12 / tot / 1e9}


def coord(K=4, reps=3):
    """8f-4: coordination numbers (mofs/calc_pair_dist_bulk.cpp) on the C3 ionic liquid: anions within 0.8 nm of every cation."""
    system = synth.ionic_liquid(31250, 22.1, seed=20200103)
    x = system.frames(0, K)
    ctx = lib.Context(system.natoms, K)
    groups(ctx, system, ["CAT", "ANI"])
    ctx.submit(x, system.box9())
    out = {"config": "coordination numbers, 31 250 cations x 31 250 anions, Rmin 0.8 nm", "frames": K}
    for cells, tag in ((0, "brute_force"), (1, "cell_list")):
        acc = ctx.coord_create(grp0=0, grp1=1, com0=0, com1=0, Rmin=0.8, dim=2, low=-100.0, up=100.0, max_count=64)
        ctx.rdf_cells(acc, cells)

        def step():
            ctx.invalidate(); ctx.accumulate(acc)
        dev = timeit(ctx, step, reps)
        ctx.reset(acc)
        ctx.timing_enable(True); ctx.timing_read(); step(); tm = ctx.timing_read(); ctx.timing_enable(False)
        cnt, st = ctx.read(acc, 64)
        out[tag] = {"ms_per_frame": dev * 1e3 / K, "k5_ms_per_frame": tm["k5_coord"][0] / K, "evaluated_pairs_per_frame": int(st[lib.STAT_PAIRS]) / K,
                    "evaluated_pairs_per_s": int(st[lib.STAT_PAIRS]) / (tm["k5_coord"][0] * 1e-3),
                    "reference_equivalent_pairs_per_s": 31250.0 * 31250.0 * K / dev, "fp64_reevaluated_per_frame": int(st[lib.STAT_EXACT]) / K,
                    "mean_coordination": float((cnt * np.arange(64)).sum() / max(cnt.sum(), 1))}
    ctx.close()
    return out


def c5(K=1, reps=2):
    """4M atoms, 4 species, all 10 group pairs, Rm = 3 nm (Rbin 1500): cell-binned K3 vs brute force on the big pair."""
    system = synth.mixed_box()
    x = lib.pinned_empty((K, system.natoms, 3), np.float32)
    t = time.perf_counter(); system.frames(0, K, out=x); gen_s = time.perf_counter() - t
    names = ["SOL", "CAT", "ANI", "COS"]
    ctx = lib.Context(system.natoms, K)
    groups(ctx, system, names)
    Rm = 3.0; Rbin = int(np.float32(3.0) / 0.002)
    accs = {}
    for a in range(4):
        for b in range(a, 4):
            acc = ctx.rdf_create(grp0=a, grp1=b, com0=0, com1=0, nbin=1, low=-100.0, up=100.0, dim=0, dim2=2, low2=-100.0,
                                 up2=100.0, Rm=Rm, Rbin=Rbin)
            ctx.rdf_tune(acc, 1, 0)
            accs[(a, b)] = acc
    ctx.submit(x, system.box9())

    def step():
        ctx.invalidate()
        for acc in accs.values():
            ctx.accumulate(acc)
    dev = timeit(ctx, step, reps)
    per = {}
    for (a, b), acc in accs.items():
        ctx.reset(acc); ctx.sync()
        t = time.perf_counter(); ctx.accumulate(acc); ctx.sync(); dt = time.perf_counter() - t
        cnt, st = ctx.read(acc, Rbin)
        per[f"{names[a]}-{names[b]}"] = {"ms": dt * 1e3, "evaluated_pairs": int(st[lib.STAT_PAIRS]), "counted": int(cnt.sum()),
                                         "all_pairs": int(ctx.nmol(a)) * int(ctx.nmol(b)) * K}
    # brute force (symmetric) on the dominant pair for comparison, one frame
    ww = accs[(0, 0)]
    ctx.rdf_cells(ww, 0); ctx.reset(ww); ctx.sync()
    t = time.perf_counter(); ctx.accumulate(ww); ctx.sync(); brute = time.perf_counter() - t
    cnt_b, _ = ctx.read(ww, Rbin)
    ctx.close()
    return {"config": "C5 4M atoms, 10 group pairs, Rm=3 nm", "frames": K, "ms_per_frame_all_pairs_of_groups": dev * 1e3 / K,
            "frames_per_s": K / dev, "per_pair": per, "water_water_brute_force_sym_ms": brute * 1e3,
            "water_water_counts_equal": int(cnt_b.sum()) == per["SOL-SOL"]["counted"], "host_generation_s": gen_s}


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--c3-frames", type=int, default=4)
    ap.add_argument("--c4-frames", type=int, default=64)
    ap.add_argument("--c5", action="store_true")
    ap.add_argument("--only-c5", action="store_true")
    ap.add_argument("--c5-frames", type=int, default=1)
    a = ap.parse_args()
    ap2 = a
    runs = [c1, lambda: c3(a.c3_frames), lambda: c4(a.c4_frames), coord] if not a.only_c5 else []
    if a.c5 or a.only_c5:
        runs.append(lambda: c5(a.c5_frames))
    for r in runs:
        print(json.dumps(r()))
