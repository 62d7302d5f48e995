"""K4 (orientation) and K1 (dipole centres) on the config-4 box with DISTINCT frames (64 x 6 MB = 384 MB > L2), L2 flushed between
launches: the honest HBM number.  Sweeps the launch knobs through the environment (read by libb2a at every launch).

    python tools/bench_k4.py [--frames 64] [--reps 10] [--sweep] [--one]       (--one: a single launch sequence, for ncu)
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gmx2020postanalysis_b200 import lib  # noqa: E402
from workloads import synth  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--frames", type=int, default=64)
ap.add_argument("--reps", type=int, default=10)
ap.add_argument("--sweep", action="store_true")
ap.add_argument("--one", action="store_true")
ap.add_argument("--timeline", action="store_true")
ap.add_argument("--nmol", type=int, default=166666)
a = ap.parse_args()

system = synth.water_box(a.nmol, 17.09, seed=20200104)
K = a.frames
x = lib.pinned_empty((K, system.natoms, 3), np.float32)
system.frames(0, K, out=x)
ctx = lib.Context(system.natoms, K)
gi = synth.group_info(system, "SOL")
ctx.set_group(0, gi["index"], gi["napm"], gi["mass"], gi["charge"])
kw = dict(grp=0, com=0, DIMN=2, low=-1.0, up=100.0, c1=0.0, c2=0.0, dr=100.0, CR=100.0, Cr=0.0, nAngle=90, tp1=1, tp2=2, tp3=3, DIMNOrient=2, method=1)
ori = ctx.orient_create(**kw)
ctx.submit(x, system.box9())
stream = torch.cuda.ExternalStream(ctx.stream)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
peak = 6558.1
try:
    peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass
nbytes = 12.0 * system.natoms * K


def run(tag, env, dipole=False, reps=None, fused=False):
    ctx.acc_also_centers(ori, lib.COM_DIPOLE if fused else -1)
    for k in ("B2A_K4_THREADS", "B2A_ITEM_FRAMES", "B2A_ORIENT_VAR", "B2A_PARK_NS", "B2A_K4_STAGES", "B2A_K4_FUSE"):
        os.environ.pop(k, None)
    os.environ.update({k: str(v) for k, v in env.items()})
    reps = reps or a.reps
    ctx.reset(ori)
    def step():
        ctx.invalidate()
        if dipole and not fused:
            ctx.centers_device(0, lib.COM_DIPOLE)        # K1 (com 3), then K4: two passes over the coordinates
        ctx.accumulate(ori)
        if dipole and fused:
            ctx.centers_device(0, lib.COM_DIPOLE)        # already written by the orientation kernel: no launch
    for _ in range(2):
        step()
    ctx.sync()
    ctx.timing_enable(True); ctx.timing_read()
    for _ in range(reps):
        with torch.cuda.stream(stream):
            flush.zero_()
        step()
    tm = ctx.timing_read()
    ctx.timing_enable(False)
    k4 = tm["k4_orient"][0] / max(tm["k4_orient"][1], 1)
    k1 = tm["k1_centers"][0] / max(tm["k1_centers"][1], 1)
    c, st = ctx.read(ori)
    out = {"tag": tag, "k4_ms": round(k4, 4), "k4_GBs": round(nbytes / k4 / 1e6, 1), "k4_frac": round(nbytes / k4 / 1e6 / peak, 3),
           "exact": int(st[lib.STAT_EXACT]) // (reps + 2), "counted": int(c[:90].sum()) // (reps + 2)}
    if dipole:
        out.update(k1_ms=round(k1, 4), total_ms=round(k1 + k4, 4))
    print(json.dumps(out), flush=True)
    return out


if a.timeline:
    import ctypes as C
    os.environ["B2A_STREAM_DEBUG"] = "1"
    for rep in range(3):
        with torch.cuda.stream(stream):
            flush.zero_()
        ctx.invalidate(); ctx.accumulate(ori); ctx.sync()
    nb = 4096
    buf = np.zeros(nb * 8, dtype=np.uint64)
    L = lib.load()
    L.b2a_debug_read.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
    assert L.b2a_debug_read(ctx._h, buf.ctypes.data, nb * 8) == 0
    t = buf.reshape(nb, 8).astype(np.float64)
    t = t[t[:, 0] > 0]
    t0 = t[:, 0].min()
    names = ["start", "prologue_done", "first_stage", "stream_end(warp0)", "all_warps_done", "drained", "end"]
    print("blocks", len(t), "queue entries per block: mean %.1f max %d" % (t[:, 7].mean(), t[:, 7].max()))
    for i, nme in enumerate(names):
        col = (t[:, i] - t0) / 1e3
        print(f"{nme:20s} min {col.min():8.2f}  median {np.median(col):8.2f}  max {col.max():8.2f} us")
    print("per block: prologue %.2f, first data %.2f, stream %.2f, wait for slowest warp %.2f, drain %.2f, flush %.2f us (medians)" % tuple(
        np.median(t[:, i + 1] - t[:, i]) / 1e3 for i in range(6)))
    del os.environ["B2A_STREAM_DEBUG"]
    sys.exit(0)
if a.one:
    ctx.invalidate(); ctx.accumulate(ori); ctx.sync()
    sys.exit(0)
run("default", {})
run("default+dipole(K1 com3)", {}, dipole=True)
run("fused dipole + orientation (one pass)", {}, dipole=True, fused=True)
if a.sweep:
    for th in (64, 128, 192, 256):
        for itf in (1, 2, 4, 8):
            run(f"threads={th} item_frames={itf}", {"B2A_K4_THREADS": th, "B2A_ITEM_FRAMES": itf})
    run("var0 (one molecule per thread)", {"B2A_ORIENT_VAR": 0})
    run("park0", {"B2A_PARK_NS": 0})
ctx.close()
