"""Quick GPU bring-up: parity of every kernel against the oracle on small systems, then C2-size timing.
Run under gpurun:  python tests/gpu_check.py [--big]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import refprog  # noqa: E402
from gmx2020postanalysis_b200 import lib  # noqa: E402
from workloads import synth  # noqa: E402
from oracle import port  # noqa: E402


def setup_groups(ctx, system, names):
    infos = []
    for g, name in enumerate(names):
        gi = refprog.group_info(system, name)
        ctx.set_group(g, gi["index"], gi["napm"], gi["mass"], gi["charge"])
        infos.append(gi)
    return infos


def check_loaders(system, names, K=3):
    ctx = lib.Context(system.natoms, K)
    infos = setup_groups(ctx, system, names)
    x = system.frames(0, K)
    box = system.box9()
    Lbox = port.lbox_from_box9(box)
    ctx.submit(x, box)
    ok = True
    for g, gi in enumerate(infos):
        for com in (0, 1, 2):
            w = gi["w"][com]
            for f in range(K):
                got = ctx.centers(g, com, f)
                exp = port.load_position_center(x[f], gi["index"], gi["napm"], w, Lbox)
                same = np.array_equal(got, exp, equal_nan=True)
                ok &= same
                if not same:
                    print("  centre mismatch", names[g], com, f, np.nanmax(np.abs(got - exp)))
        got = ctx.centers(g, 3, 0)
        exp = port.load_position_center(x[0], gi["index"], gi["napm"], gi["charge"], Lbox, dipole_guard=True)
        ok &= np.array_equal(got, exp, equal_nan=True)
        pos = ctx.positions(g, gi["napm"], 1)
        ok &= np.array_equal(pos, port.load_position(x[1], gi["index"], gi["napm"], Lbox))
    ctx.close()
    print("loaders", names, "OK" if ok else "FAIL")
    return ok


def check_density(system, name, K=4, nbin=100, dim=2, low=0.0, up=4.476, inclusive=0, dim2=-1, low2=0.0, up2=0.0):
    ctx = lib.Context(system.natoms, K)
    (gi,) = setup_groups(ctx, system, [name])
    x = system.frames(0, K)
    box = system.box9()
    Lbox = port.lbox_from_box9(box)
    dbin = float(np.float32(np.float32(np.float32(up) - np.float32(low)) / np.float32(nbin)))
    acc = ctx.density_create(grp=0, com=0, dim=dim, low=low, up=up, dbin=dbin, nbin=nbin, upper_inclusive=inclusive,
                             dim2=dim2, low2=low2, up2=up2)
    ctx.submit(x, box)
    ctx.accumulate(acc)
    got, st = ctx.read(acc, nbin)
    exp = np.zeros(nbin, np.uint64)
    drop = 0
    for f in range(K):
        pc = port.load_position_center(x[f], gi["index"], gi["napm"], gi["w"][0], Lbox)
        _, d = port.density_accum(pc, dim, refprog.f32(low), refprog.f32(up), dbin, nbin, exp, bool(inclusive), dim2,
                                  refprog.f32(low2), refprog.f32(up2))
        drop += d
    ok = np.array_equal(got, exp) and int(st[lib.STAT_DROPPED]) == drop and int(st[0]) == K
    print("density", name, "OK" if ok else "FAIL", int(got.sum()), int(exp.sum()), "dropped", int(st[1]), drop)
    ctx.close()
    return ok


def check_rdf(system, g0, g1, K=2, paranoid=0, **kw):
    spec = dict(nbin=5, low=-100.0, up=100.0, dim=0, dim2=2, low2=-100.0, up2=100.0, Rm=0.0, com=0)
    spec.update(kw)
    exp, est, Rm, Rbin = refprog.oracle_rdf_region(system, K, g0, g1, return_counts=True, **spec)
    ctx = lib.Context(system.natoms, K)
    setup_groups(ctx, system, [g0, g1])
    acc = ctx.rdf_create(grp0=0, grp1=1, com0=spec["com"], com1=spec["com"], nbin=spec["nbin"], low=spec["low"],
                         up=spec["up"], dim=spec["dim"], dim2=spec["dim2"], low2=spec["low2"], up2=spec["up2"], Rm=Rm,
                         Rbin=Rbin)
    ctx.rdf_tune(acc, 0, paranoid)
    ctx.submit(system.frames(0, K), system.box9())
    ctx.accumulate(acc)
    got, st = ctx.read(acc, spec["nbin"] * Rbin)
    got = got.reshape(Rbin, spec["nbin"])
    ok = np.array_equal(got, exp)
    print(f"rdf {g0}-{g1} K={K} paranoid={paranoid} Rm={Rm:.4f} Rbin={Rbin}:", "OK" if ok else "FAIL",
          "counted", int(got.sum()), int(exp.sum()), "stats", [int(v) for v in st], "oracle", [int(v) for v in est])
    if not ok:
        d = got.astype(np.int64) - exp.astype(np.int64)
        nz = np.argwhere(d != 0)
        print("   differing cells:", len(nz), nz[:10].tolist(), d[d != 0][:10].tolist())
    ctx.close()
    return ok


def time_c2(nmol=72000, L=12.913, K=4, reps=3):
    import ctypes as C
    system = synth.water_box(nmol, L, seed=20200102, split_atoms=True)
    ctx = lib.Context(system.natoms, K)
    setup_groups(ctx, system, ["OW", "OW"])
    Lbox = port.lbox_from_box9(system.box9())
    Rm = port.rdf_default_rm(Lbox)
    Rbin = port.rdf_rbin(Rm)
    acc = ctx.rdf_create(grp0=0, grp1=1, com0=0, com1=0, nbin=5, low=-100.0, up=100.0, dim=0, dim2=2, low2=-100.0,
                         up2=100.0, Rm=Rm, Rbin=Rbin)
    x = system.frames(0, K)
    ctx.submit(x, system.box9())
    ctx.accumulate(acc)
    ctx.sync()
    t = time.perf_counter()
    for _ in range(reps):
        ctx.accumulate(acc)
    ctx.sync()
    dt = (time.perf_counter() - t) / reps
    got, st = ctx.read(acc, 5 * Rbin)
    pairs = nmol * nmol * K
    print(f"C2 timing: N={nmol} K={K}: {dt*1e3/K:.3f} ms/frame, {pairs/dt:.3e} pairs/s, stats {[int(v) for v in st]}")
    ctx.close()


if __name__ == "__main__":
    big = "--big" in sys.argv
    ok = True
    w = synth.water_box(3000, 4.476)
    il = synth.ionic_liquid(400, 5.1)
    ok &= check_loaders(w, ["SOL"])
    ok &= check_loaders(il, ["CAT", "ANI"])
    ok &= check_density(w, "SOL")
    ok &= check_density(il, "CAT", nbin=37, dim=0, low=0.3, up=4.9, inclusive=1, dim2=1, low2=0.5, up2=4.0)
    ws = synth.water_box(4000, 4.93, split_atoms=True)
    ok &= check_rdf(ws, "OW", "OW", K=2)
    ok &= check_rdf(il, "CAT", "ANI", K=3, nbin=3, low=0.5, up=4.6, dim=1, dim2=0, low2=0.3, up2=4.9, Rm=1.9)
    ok &= check_rdf(il, "CAT", "ANI", K=2)
    ok &= check_rdf(synth.water_box(1500, 3.55, split_atoms=True), "OW", "OW", K=1, paranoid=1)
    print("ALL PARITY", "OK" if ok else "FAIL")
    if big:
        time_c2()
    sys.exit(0 if ok else 1)
