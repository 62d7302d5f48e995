"""Config 1 (README density template: 3000 SPC/E, 100 frames per launch = 10.8 MB): per-block time line of the K2 launch.
    python tools/bench_k2_c1.py [--frames 100] [--nmol 3000]"""
import argparse
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gmx2020postanalysis_b200 import lib  # noqa: E402
from workloads import synth  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--frames", type=int, default=100)
ap.add_argument("--nmol", type=int, default=3000)
a = ap.parse_args()
system = synth.water_box(a.nmol, 4.476, seed=20200101)
K = a.frames
x = lib.pinned_empty((K, system.natoms, 3), np.float32)
system.frames(0, K, out=x)
ctx = lib.Context(system.natoms, K)
gi = synth.group_info(system, "SOL")
ctx.set_group(0, gi["index"], gi["napm"], gi["mass"], gi["charge"])
dbin = float(np.float32(np.float32(4.476) / np.float32(100)))
acc = ctx.density_create(grp=0, com=0, dim=2, low=0.0, up=4.476, dbin=dbin, nbin=100, upper_inclusive=0, dim2=-1, low2=0.0, up2=0.0)
ctx.submit(x, system.box9())
for _ in range(50):
    ctx.accumulate(acc)
ctx.sync()
ctx.timing_enable(True); ctx.timing_read()
for _ in range(200):
    ctx.accumulate(acc)
tm = ctx.timing_read()
print("k2_density: %.2f us per launch of %d frames (%.1f MB)" % (1e3 * tm["k2_density"][0] / tm["k2_density"][1], K, 12e-6 * system.natoms * K))
os.environ["B2A_STREAM_DEBUG"] = "1"
for _ in range(3):
    ctx.accumulate(acc); ctx.sync()
nb = 4096
buf = np.zeros(nb * 8, dtype=np.uint64)
L = lib.load()
L.b2a_debug_read.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
assert L.b2a_debug_read(ctx._h, buf.ctypes.data, nb * 8) == 0
t = buf.reshape(nb, 8).astype(np.float64)
t = t[t[:, 0] > 0]
t0 = t[:, 0].min()
names = ["start", "prologue_done", "first_stage", "stream_end(warp0)", "all_warps_done", "drained", "end"]
print("blocks", len(t))
for i, nme in enumerate(names):
    col = (t[:, i] - t0) / 1e3
    print(f"{nme:20s} min {col.min():8.2f}  median {np.median(col):8.2f}  max {col.max():8.2f} us")
