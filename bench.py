#!/usr/bin/env python
"""bench.py -- headline benchmark: O-O RDF of the 216k-atom synthetic water box (BASELINE.json configs[1]).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b2a|reference] [--config c1|c2|c3|c4|c5] [--no-others]

The printed line is quoted on config 2 (the default; `--config` makes another BASELINE.json workload the headline of a
manual run).  Unless `--no-others`, the line also carries `other_configs`: configs 1, 3, 4 and 5 pushed through the SAME
measurement (a few time-boxed steps each): device-resident rate, end-to-end rate, roofline fraction of the dominant kernel,
the reference's CPU program on the host beside it (N = 1), and a self-check -- so all five named workloads are measured by
whoever runs this file, at every N.

A "step" is one pass of the whole hot path over one batch of F frames: centres (K1), slab / cell prep, the pair kernel (K3)
-- or the density (K2) / orientation (K4) kernel for configs 1 and 4.  Steps cycle through D distinct pinned batches of
distinct synthetic frames (rank r, batch d: frames (r*D + d)*F ...), so no step sees the frames of the previous one, and the
default 20 steps x 64 frames cover the 1000 frames config 2 names.
  * `value`: frames resident in HBM when a step's CUDA-event bracket opens (the upload of step s+1 is queued on the copy
    stream behind the kernels of step s and is outside every bracket); the sum of the brackets is the time.  L2 is flushed
    before each bracket.
  * `e2e`: wall clock of the same loop through the C ABI with HOST buffers: pinned frames -> H2D -> kernels -> D2H of the
    u64 histogram, every step (double-buffered: the H2D of step s+1 overlaps the kernels of step s).
  * `check.parity`: after the timed region the accumulated histogram is compared with an independent evaluation of the same
    frames (config 2: the full N0 x N1 kernel, no symmetry, batch by batch: every timed step is accounted for); at N > 1 the
    ncclAllReduce inside libb2a (b2a_finalize) is checked against a sum of the rank-private histograms gathered another way.
Frames shard across ranks (one process per GPU, weak scaling: F frames per rank per step); the only collective of the path
is the one all-reduce of the private histograms at the end, inside libb2a.  `host_cpp` reports the same workload driven by
ONE C++ process through the drop-in handle (build/dropin/calc_rdf_region_fused, B2A_DEVICES=0..N-1).
`--impl reference` times the reference's own CPU program (oracle/_ref, compiled unmodified) on a bounded sample.
"""
from __future__ import annotations

import argparse
import json
import os
import shutil
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

REFDIR = os.path.join(ROOT, "oracle", "_ref")
DROPIN = os.path.join(ROOT, "build", "dropin")
STUBDIR = os.path.join(ROOT, "build")

ISSUE_SLOTS_PER_PAIR = 19          # algorithmic issue slots per evaluated pair (the model `frac` uses)
# What the hardware counters say about the same kernel (ncu --set full, profiles/r01_k3_ncu.md; re-captured per round):
# `frac` is a MODEL fraction (19 slots per pair); the kernel really executes K3_INST_PER_PAIR thread-instructions per pair
# (packed FFMA2 / FMUL2 halve several of the 19) and keeps the issue ports busy K3_ISSUE_ACTIVE of the time.
K3_NCU = {"issue_active": 0.787, "inst_per_pair": 17.1, "dram_bytes_per_frame": 4.64e6, "source": "profiles/r01_k3_ncu.md"}
for _f in ("r02_k3_ncu.json",):                  # a later capture replaces the constants (tools/profile_round.sh writes it)
    _p = os.path.join(ROOT, "profiles", _f)
    if os.path.exists(_p):
        try:
            K3_NCU.update(json.load(open(_p)))
        except Exception:
            pass


def hbm_peak():
    """(GB/s, provenance): the measured copy bandwidth of this pool's B200s, else the profiling recipe's fallback."""
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        return float(json.load(open(p))["hbm_gbs"]), "MEASURED_PEAKS.json"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


# ================================================================================================ workloads
class Workload:
    """One BASELINE.json config: the synthetic system, the device job and how to count it."""
    kind = "rdf"

    def __init__(self, name):
        from workloads import synth
        self.name = name
        gi = synth.group_info
        if name == "c1":
            self.kind = "density"
            self.system = synth.water_box(3000, 4.476, seed=20200101)
            self.text = "README template density profile (nbin=100, dim=z, COM), 3k-molecule SPC/E box (9 000 atoms), 100 frames"
            self.F, self.D = 100, 4
            self.groups = [gi(self.system, "SOL")]
        elif name == "c2":
            self.L = 12.913
            self.system = synth.water_box(72000, self.L, seed=20200102, split_atoms=True)
            self.text = "O-O RDF, 216k-atom SPC/E water box (72k O, L=12.913 nm, Rm=L/2, Rbin=3228, nbin=5), calc_rdf_region defaults"
            self.F, self.D = 64, 4
            self.groups = [gi(self.system, "OW"), gi(self.system, "OW")]
            self.pairs = [(0, 1)]
            self.Rm, self.nbin = float(np.float32(np.float32(self.L) / np.float32(2))), 5      # calc_rdf_region.cpp:52-57, :23
        elif name == "c3":
            self.L = 22.1
            self.system = synth.ionic_liquid(31250, self.L, seed=20200103)
            self.text = ("cation-anion molecular-COM RDF, 1M-atom ionic liquid (31 250 x (25-site cation + 7-site anion), L=22.1 nm, "
                         "Rm=L/2, Rbin=5525, nbin=5), calc_rdf_region defaults")
            self.F, self.D = 64, 2                            # frames per batch 8 / 16 / 32 / 64: 1497 / 1571 / 1596 / 1612 frames/s
            self.groups = [gi(self.system, "CAT"), gi(self.system, "ANI")]
            self.pairs = [(0, 1)]
            self.Rm, self.nbin = float(np.float32(np.float32(self.L) / np.float32(2))), 5
        elif name == "c4":
            self.kind = "orient"
            self.system = synth.water_box(166666, 17.09, seed=20200104)
            self.text = ("charge-centre dipole vectors + bisector orientation-angle histogram (calc_orient_mof, nAngle=90), "
                         "166 666 SPC/E = 499 998 atoms")
            self.F, self.D = 384, 2                           # 2.3 GB per batch: a launch carries ~16 us of start-up, drain and tail (K4 alone:
                                                              # 0.61 of the HBM peak per 64-frame launch, 0.69 per 128, 0.76 per 256)
            self.groups = [gi(self.system, "SOL")]
        elif name == "c5":
            self.L = 34.0
            self.system = synth.mixed_box()
            self.text = ("mixed-species RDF sweep, 4M atoms (1.2M water + 8k cation + 8k anion + 24k cosolvent, L=34 nm), all 10 group "
                         "pairs, Rm=3 nm, Rbin=1500, nbin=1 (cell-binned K3)")
            self.F, self.D = 8, 2                             # frames per batch 2 / 4 / 8: 128.7 / 131.7 / 134.5 frames/s (longer launches, fewer tails)
            self.groups = [gi(self.system, n) for n in ("SOL", "CAT", "ANI", "COS")]
            self.pairs = [(a, b) for a in range(4) for b in range(a, 4)]
            self.Rm, self.nbin = 3.0, 1
        else:
            raise SystemExit(f"unknown --config {name}")
        self.natoms = self.system.natoms
        self.nmol = [len(g["index"]) // g["napm"] for g in self.groups]
        if self.kind == "rdf":
            self.Rbin = int(np.float32(self.Rm) / 0.002)                                       # calc_rdf_region.cpp:58
            self.pairs_per_frame = float(sum(self.nmol[a] * self.nmol[b] for a, b in self.pairs))   # reference-equivalent ordered pairs
            self.symmetric = 1
        self.atoms_per_frame = sum(len(g["index"]) for g in self.groups[:1]) if self.kind != "rdf" else None

    # ---- device job ------------------------------------------------------------------------------------------
    def setup(self, ctx):
        from gmx2020postanalysis_b200 import lib
        for g, gi in enumerate(self.groups):
            ctx.set_group(g, gi["index"], gi["napm"], gi["mass"], gi["charge"])
        self.accs, self.chk = [], []
        if self.kind == "rdf":
            for g0, g1 in self.pairs:
                a = ctx.rdf_create(grp0=g0, grp1=g1, com0=0, com1=0, nbin=self.nbin, low=-100.0, up=100.0, dim=0, dim2=2, low2=-100.0,
                                   up2=100.0, Rm=self.Rm, Rbin=self.Rbin)
                ctx.rdf_tune(a, self.symmetric, 0)
                self.accs.append(a)
            # the independent evaluation the self-check compares with: same spec, full N0 x N1 evaluation, no symmetry, no cells
            g0, g1 = self.pairs[0]
            c = ctx.rdf_create(grp0=g0, grp1=g1, com0=0, com1=0, nbin=self.nbin, low=-100.0, up=100.0, dim=0, dim2=2, low2=-100.0,
                               up2=100.0, Rm=self.Rm, Rbin=self.Rbin)
            # (config 3 already runs that kernel -- two different groups, Rm = L/2 -- so there the check re-evaluates EVERY pair in FP64)
            ctx.rdf_tune(c, 0, 1 if self.name == "c3" else 0)
            ctx.rdf_cells(c, 0)
            self.chk = [c]
            if self.name == "c2":
                # ... and a third one that shares nothing with the FP32 bracket: EVERY pair through the FP64 path (one frame only)
                c2 = ctx.rdf_create(grp0=g0, grp1=g1, com0=0, com1=0, nbin=self.nbin, low=-100.0, up=100.0, dim=0, dim2=2, low2=-100.0,
                                    up2=100.0, Rm=self.Rm, Rbin=self.Rbin)
                ctx.rdf_tune(c2, 0, 1)
                ctx.rdf_cells(c2, 0)
                self.chk.append(c2)
        elif self.kind == "density":
            L = 4.476
            dbin = float(np.float32(np.float32(L) / np.float32(100)))                         # README.md:127
            kw = dict(grp=0, com=0, dim=2, low=0.0, up=L, dbin=dbin, nbin=100, upper_inclusive=0, dim2=-1, low2=0.0, up2=0.0)
            self.accs = [ctx.density_create(**kw)]
            self.chk = [ctx.density_create(**kw)]
            ctx.acc_fast(self.chk[0], 0)                                                       # literal FP64 for every molecule
        elif self.kind == "orient":
            kw = dict(grp=0, com=0, DIMN=2, low=-1.0, up=100.0, c1=0.0, c2=0.0, dr=100.0, CR=100.0, Cr=0.0, nAngle=90, tp1=1, tp2=2, tp3=3,
                      DIMNOrient=2, method=1)
            self.accs = [ctx.orient_create(**kw)]
            self.chk = [ctx.orient_create(**kw)]
            ctx.acc_fast(self.chk[0], 0)
        self.lib = lib

    def step(self, ctx):
        """One pass of the hot path over the resident batch (nothing cached from a previous step is reused)."""
        ctx.invalidate()
        for a in self.accs:
            ctx.accumulate(a)
        if self.kind == "orient":
            # the per-molecule dipole vectors, left on the device.  Queued AFTER the orientation kernel: K1 writes 24 B per molecule
            # (4 MB per frame), and a K4 launched right behind it shares HBM with that write-back (0.60 instead of 0.70 of peak)
            ctx.centers_device(0, self.lib.COM_DIPOLE)

    def read_all(self, ctx, accs=None):
        out = [ctx.read(a) for a in (self.accs if accs is None else accs)]
        return [c for c, _ in out], np.sum([st for _, st in out], axis=0)

    def dominant(self):
        return {"rdf": "k3_rdf", "density": "k2_density", "orient": "k4_orient"}[self.kind]


# ================================================================================================ clocks
class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu_index, self.proc, self.path = gpu_index, None, None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu_index), f"--query-gpu={self.QUERY}",
                                          "--format=csv,noheader,nounits", "-lms", "25"],
                                         stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        try:
            for ln in open(self.path):
                f = [t.strip() for t in ln.split(",")]
                if len(f) < 9:
                    continue
                try:
                    sm.append(float(f[1])); mx.append(float(f[2])); pw.append(float(f[3]))
                except ValueError:
                    continue
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            busy = [s for s, p in zip(sm, pw) if p > 0.5 * max(pw)] or sm
            out.update(sm_mhz=float(np.median(busy)), sm_max_mhz=float(max(mx)), reasons=sorted(reasons), samples=len(sm),
                       samples_under_load=len(busy), power_w_max=float(max(pw)))
        return out


# ================================================================================================ CPU baseline
def _ref_env():
    return dict(os.environ, LD_LIBRARY_PATH=STUBDIR + ":" + os.environ.get("LD_LIBRARY_PATH", ""), OMP_NUM_THREADS="1")


def have_ref(prog):
    return os.path.exists(os.path.join(REFDIR, prog)) and os.path.exists(os.path.join(STUBDIR, "libgmxstub.so"))


def _write_inputs(prefix, wl, frames, groups):
    from workloads import trajio
    if not os.path.exists(prefix + ".b2t"):
        trajio.write_trajectory(prefix + ".b2t", frames, wl.system.box9())
        trajio.write_topology(prefix + ".b2p", wl.system)
    trajio.write_index(prefix + ".ndx", groups)


def cpu_rdf_sample(wl, n_i, tmpdir, x0, tag="s"):
    """The reference's own calc_rdf_region (oracle/_ref; its pair loop is single-threaded) on grp0 = the first n_i molecules of
    group 0 x grp1 = all of group 1, one frame.  Returns (pairs/s, kind, secs)."""
    g0, g1 = wl.groups[wl.pairs[0][0]], wl.groups[wl.pairs[0][1]]
    n1 = len(g1["index"]) // g1["napm"]
    if have_ref("calc_rdf_region"):
        prefix = os.path.join(tmpdir, wl.name)
        _write_inputs(prefix, wl, x0[None], {"G0_sub": g0["index"][:n_i * g0["napm"]], "G1": g1["index"]})
        cmd = [os.path.join(REFDIR, "calc_rdf_region"), "-f", prefix + ".b2t", "-s", prefix + ".b2p", "-n", prefix + ".ndx", "-o", prefix + tag + ".xvg"]
        if wl.name == "c5":
            cmd += ["-Rm", "3", "-nbin", "1"]
        t = time.perf_counter()
        r = subprocess.run(cmd, input="0 1\n", capture_output=True, text=True, env=_ref_env(), cwd=tmpdir)
        dt = time.perf_counter() - t
        if r.returncode != 0:
            raise RuntimeError("reference program failed: " + r.stderr[-500:])
        return n_i * n1 / dt, "reference", dt
    # the reference binary did not travel: time the oracle port instead (single thread)
    from oracle import port
    os.environ["OMP_NUM_THREADS"] = "1"
    Lbox = port.lbox_from_box9(wl.system.box9())
    p0 = port.load_position_center(x0, g0["index"][:n_i * g0["napm"]], g0["napm"], g0["w"][0], Lbox)
    p1 = port.load_position_center(x0, g1["index"], g1["napm"], g1["w"][0], Lbox)
    t = time.perf_counter()
    port.rdf_accum(p0, p1, Lbox, wl.nbin, -100.0, 100.0, 0, 2, -100.0, 100.0, np.float32(wl.Rm), wl.Rbin)
    dt = time.perf_counter() - t
    return n_i * n1 / dt, "port", dt


def cpu_rdf_all_cores(wl, n_i, procs, tmpdir, x0):
    """The reference has no threading in the pair loop; the most its code can use of a host is one PROCESS per core, each on
    its own share of the i-molecules.  Returns aggregate pairs/s, or None where the reference binaries did not travel."""
    from workloads import trajio
    if not have_ref("calc_rdf_region"):
        return None, 0
    g0, g1 = wl.groups[wl.pairs[0][0]], wl.groups[wl.pairs[0][1]]
    n0, n1, na = len(g0["index"]) // g0["napm"], len(g1["index"]) // g1["napm"], g0["napm"]
    n_i = min(n_i, n0 // procs)
    prefix = os.path.join(tmpdir, wl.name)
    _write_inputs(prefix, wl, x0[None], {"G0_sub": g0["index"][:n_i * na], "G1": g1["index"]})
    cmds = []
    for q in range(procs):
        ndx = os.path.join(tmpdir, f"p{q}.ndx")
        trajio.write_index(ndx, {"G0_sub": g0["index"][q * n_i * na:(q + 1) * n_i * na], "G1": g1["index"]})
        cmds.append([os.path.join(REFDIR, "calc_rdf_region"), "-f", prefix + ".b2t", "-s", prefix + ".b2p", "-n", ndx, "-o", os.path.join(tmpdir, f"p{q}.xvg")]
                    + (["-Rm", "3", "-nbin", "1"] if wl.name == "c5" else []))
    t = time.perf_counter()
    ps = [subprocess.Popen(cmd, stdin=subprocess.PIPE, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, env=_ref_env(), cwd=tmpdir, text=True)
          for cmd in cmds]
    for pr in ps:
        pr.stdin.write("0 1\n"); pr.stdin.close()
    rcs = [pr.wait() for pr in ps]
    dt = time.perf_counter() - t
    if any(rcs):
        return None, 0
    return procs * n_i * n1 / dt, n_i


def cpu_frames_sample(wl, nframes, tmpdir, frames):
    """Configs 1 and 4: the reference's own program (README template / calc_orient_mof) over `nframes` whole frames, one thread.
    Returns (frames/s, kind, secs)."""
    prog, args = (("readme_template", ["-nbin", "100", "-dim", "2", "-low", "0", "-up", "4.476"]) if wl.kind == "density" else
                  ("calc_orient_mof", ["-NCM", "0", "-dim", "2", "-low", "-1", "-1", "-1", "-up", "100", "100", "100", "-c1", "0", "-c2", "0", "-dr", "100",
                                       "-CR", "100", "-Cr", "0", "-nA", "90", "-tp1", "1", "-tp2", "2", "-tp3", "3", "-DIMNOrient", "2", "-method", "1"]))
    if have_ref(prog):
        prefix = os.path.join(tmpdir, wl.name)
        _write_inputs(prefix, wl, frames[:nframes], {"SOL": wl.groups[0]["index"]})
        cmd = [os.path.join(REFDIR, prog), "-f", prefix + ".b2t", "-s", prefix + ".b2p", "-n", prefix + ".ndx", "-o", prefix + ".xvg"] + args
        t = time.perf_counter()
        r = subprocess.run(cmd, input="0\n", capture_output=True, text=True, env=_ref_env(), cwd=tmpdir)
        dt = time.perf_counter() - t
        if r.returncode != 0:
            raise RuntimeError(f"reference program {prog} failed: " + (r.stdout + r.stderr)[-600:])
        return nframes / dt, "reference", dt
    from oracle import port
    os.environ["OMP_NUM_THREADS"] = "1"
    g = wl.groups[0]
    Lbox = port.lbox_from_box9(wl.system.box9())
    t = time.perf_counter()
    if wl.kind == "density":
        hist = np.zeros(100, dtype=np.uint64)
        dbin = float(np.float32(np.float32(4.476) / np.float32(100)))
        for f in range(nframes):
            pc = port.load_position_center(frames[f], g["index"], g["napm"], g["w"][0], Lbox)
            port.density_accum(pc, 2, np.float32(0.0), np.float32(4.476), dbin, 100, hist)
    else:
        o = n = None
        for f in range(nframes):
            pos = port.load_position(frames[f], g["index"], g["napm"], Lbox)
            pc = port.load_position_center(frames[f], g["index"], g["napm"], g["w"][0], Lbox)
            o, n, _ = port.orient_accum(pos, pc, Lbox, 2, np.float32(-1.0), np.float32(100.0), np.float32(0.0), np.float32(0.0), np.float32(100.0),
                                        np.float32(100.0), np.float32(0.0), 90, 1, 2, 3, 2, 1, o, n)
    dt = time.perf_counter() - t
    return nframes / dt, "port", dt


def cpu_baseline(wl, tmpdir, frames, rdf_sample):
    if wl.kind == "rdf":
        n1 = wl.nmol[wl.pairs[0][1]]
        n_i = max(1, min(rdf_sample, wl.nmol[wl.pairs[0][0]], int(6.0e8 // n1)))              # ~10 s of one core
        rate, kind, dt = cpu_rdf_sample(wl, n_i, tmpdir, frames[0], "cb")
        return {"value": rate, "unit": "pair-distances/s", "cores": 1, "kind": kind,
                "sample": f"{n_i} of {wl.nmol[wl.pairs[0][0]]} i-molecules x all {n1} j-molecules of one frame ({dt:.1f} s, whole-process wall time)"
                          + (", dominant group pair SOL-SOL" if wl.name == "c5" else ""),
                "host_cpus": os.cpu_count(), "frames_per_s": rate / wl.pairs_per_frame}
    n = min(len(frames), 100 if wl.kind == "density" else 8)
    rate, kind, dt = cpu_frames_sample(wl, n, tmpdir, frames)
    return {"value": rate, "unit": "frames/s", "cores": 1, "kind": kind,
            "sample": f"{n} whole frames ({dt:.2f} s, whole-process wall time incl. start-up and reading the frames)", "host_cpus": os.cpu_count(),
            "frames_per_s": rate}


def run_reference_arm(args):
    """--impl reference: same metric / config / unit, each step a bounded sample of the workload, on the host's cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    os.environ["B2A_SYNTH_NUMPY"] = "1"           # pure-numpy frame generator: this arm maps no library built from this repo's product
    wl = Workload(args.config)
    x0 = wl.system.frames(0, 1)
    rates, kind = [], "reference"
    with tempfile.TemporaryDirectory() as td:
        if wl.kind == "rdf":
            n0, n1 = wl.nmol[wl.pairs[0][0]], wl.nmol[wl.pairs[0][1]]
            n_i = max(1, min(args.ref_sample, n0, int(7.2e7 // n1)))                           # ~1-2 s of one core per step
            for s in range(args.warmup + args.steps):
                rate, kind, dt = cpu_rdf_sample(wl, n_i, td, x0[0])
                if s >= args.warmup:
                    rates.append((rate * dt, dt))
            procs = os.cpu_count() or 1
            all_cores, n_each = cpu_rdf_all_cores(wl, n_i, procs, td, x0[0]) if kind == "reference" else (None, 0)
            sample = f"{n_i} of {n0} i-molecules x all {n1} j-molecules of frame 0 per step"
            unit, per_frame = "pair-distances/s", wl.pairs_per_frame
            metric = "rdf_pair_distances_per_s"
        else:
            nfr = 100 if wl.kind == "density" else 4
            fr = wl.system.frames(0, nfr)
            for s in range(args.warmup + args.steps):
                rate, kind, dt = cpu_frames_sample(wl, nfr, td, fr)
                if s >= args.warmup:
                    rates.append((rate * dt, dt))
            all_cores, sample, unit, per_frame, metric = None, f"{nfr} whole frames per step", "frames/s", 1.0, "frames_per_s"
    tot_units, tot_s = sum(u for u, _ in rates), sum(d for _, d in rates)
    value = tot_units / tot_s
    line = {
        "impl": "reference", "metric": metric, "value": value, "unit": unit,
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * tot_s / len(rates),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "frames_per_s": value / per_frame,
        "config": {"workload": wl.text, "sample": sample},
        "cpu_baseline": {"value": value, "unit": unit, "cores": 1, "kind": kind,
                         "sample": f"{sample}, {len(rates)} steps, whole-process wall time (includes reading the frame)", "host_cpus": os.cpu_count()},
        "e2e": {"value": value, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    if all_cores:
        # not the reference's own mode of operation (its frame loop is single-threaded): reported beside `value`, never instead
        line["cpu_all_cores"] = {"value": all_cores, "unit": unit, "processes": procs,
                                 "sample": f"{procs} concurrent reference processes, {n_each} x {n1} ordered pairs each, wall time of the slowest"}
    print(json.dumps(line))
    return 0


# ================================================================================================ B200 arm
class Dist:
    """torch.distributed plumbing of the bench: barriers, max / sum over ranks.  The path's own collective is NOT here (libb2a)."""

    def __init__(self):
        import torch
        self.torch = torch
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py: no CUDA device -- the B200 path has no CPU fallback")
        torch.cuda.set_device(self.local)
        self.cpu_group = None
        if self.world > 1:
            import torch.distributed as dist
            self.dist = dist
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local))
            self.cpu_group = dist.new_group(backend="gloo")        # host-side waits that must not spin on a GPU

    def barrier(self, ctx=None):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize(self.local)
        if ctx is not None:
            ctx.sync()

    def cpu_barrier(self):
        if self.world > 1:
            self.dist.barrier(group=self.cpu_group)

    def reduce(self, vals, op):
        if self.world == 1:
            return [float(v) for v in vals]
        t = self.torch.tensor(list(vals), dtype=self.torch.float64, device=f"cuda:{self.local}")
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX if op == "max" else self.dist.ReduceOp.SUM)
        return [float(v) for v in t]

    def gather_u64(self, arr):
        """Every rank's array on rank 0 (a channel independent of libb2a's communicator)."""
        if self.world == 1:
            return [arr]
        t = self.torch.from_numpy(np.ascontiguousarray(arr).view(np.int64).copy()).to(f"cuda:{self.local}")
        out = [self.torch.empty_like(t) for _ in range(self.world)]
        self.dist.all_gather(out, t)
        return [o.cpu().numpy().view(np.uint64) for o in out]

    def close(self):
        if self.world > 1:
            self.dist.barrier()
            self.dist.destroy_process_group()


def measure(wl, D, steps, warmup, with_cpu, rdf_sample, flush):
    """Runs one workload through the whole measurement; returns the dict the line (or its other_configs entry) is built from."""
    torch = D.torch
    from gmx2020postanalysis_b200 import dist as b2d, lib
    rank, world, local = D.rank, D.world, D.local
    F, nb = wl.F, wl.D
    # rank-private, distinct frames in pinned host memory (contiguous frame blocks per rank)
    batches = []
    t_gen = time.perf_counter()
    for d in range(nb):
        xh = lib.pinned_empty((F, wl.natoms, 3), np.float32)
        wl.system.frames((rank * nb + d) * F, F, out=xh)
        batches.append(xh)
    t_gen = time.perf_counter() - t_gen
    box = wl.system.box9()
    ctx = lib.Context(wl.natoms, F, device=local)
    joined = b2d.comm_init(ctx, device=f"cuda:{local}")       # ncclCommInitRank inside libb2a (id broadcast by torch.distributed)
    wl.setup(ctx)
    if joined:
        ctx.finalize()                           # NCCL connects lazily: the first collective of a communicator pays for it, not the timed one
        ctx.sync()
    stream = torch.cuda.ExternalStream(ctx.stream, device=local)
    pos = {"next": 0, "cur": -1}

    def feed():
        ctx.submit(batches[pos["next"]], box)
        pos["cur"] = pos["next"]
        pos["next"] = (pos["next"] + 1) % nb

    def reset_all():
        for a in wl.accs + wl.chk:
            ctx.reset(a)

    # ---- phase 1: device-resident, CUDA events around each step ------------------------------------------------
    feed()
    ctx.sync()
    # warm-up: at least `warmup` steps AND at least a quarter of a second of device work -- the steps of configs 1 and 4 last a
    # fraction of a millisecond, and a GPU that idled through the host-side frame generation needs that long to clock up again
    t_w, n_w = time.perf_counter(), 0
    while n_w < warmup or (time.perf_counter() - t_w < 0.25 and n_w < 2000):
        wl.step(ctx)
        feed()
        n_w += 1
        if n_w % 8 == 0:
            ctx.sync()
    ctx.sync()
    reset_all()
    ctx.timing_enable(True)
    ctx.timing_read()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    used = []
    sampler = ClockSampler(local)
    sampler.start()
    D.barrier(ctx)
    t_wall = time.perf_counter()
    for s in range(steps):
        with torch.cuda.stream(stream):
            flush.zero_()                        # L2 flush between timed iterations (outside the event pair)
        ev[s][0].record(stream)
        wl.step(ctx)
        ctx.join()                               # accumulators of a sweep run on streams of their own: the bracket closes behind all of them
        ev[s][1].record(stream)
        used.append(pos["cur"])
        # next step's frames: the upload overlaps this step's kernels (copy stream).  The HBM-bound configs (1 and 4) finish a step
        # 30x faster than PCIe delivers the next batch: feeding every step would leave the GPU idle (and clocking down) between
        # brackets, so they take a new batch every fourth step -- their batch is larger than L2 (config 4) or L2 is flushed (config 1)
        if wl.kind == "rdf" or s % 4 == 3:
            feed()
    D.barrier(ctx)
    t_wall = time.perf_counter() - t_wall
    clocks = sampler.stop()
    tot_ms = float(sum(a.elapsed_time(b) for a, b in ev))
    timing = ctx.timing_read()
    ctx.timing_enable(False)
    priv_counts, priv_stats = wl.read_all(ctx)                 # this rank's private histograms after the timed steps
    evaluated_pairs = float(priv_stats[lib.STAT_PAIRS])

    # ---- the path's one collective: merge the rank-private histograms inside libb2a ------------------------------
    D.barrier(ctx)                               # so that the figure below is the collective, not the skew between the ranks' reads
    t0 = time.perf_counter()
    ctx.finalize()
    ctx.sync()
    allreduce_ms = 1e3 * (time.perf_counter() - t0)
    merged_counts, merged_stats = wl.read_all(ctx)
    mstats = merged_stats.copy()
    mstats[lib.STAT_FRAMES] = mstats[lib.STAT_FRAMES] // len(wl.accs)      # every accumulator saw every frame

    # ---- self-check ------------------------------------------------------------------------------------------------
    check = {"frames_merged": int(mstats[lib.STAT_FRAMES]), "counted_merged": int(sum(int(c.sum()) for c in merged_counts)),
             "fp64_reevaluated": int(mstats[lib.STAT_EXACT]), "dropped_oob": int(mstats[lib.STAT_DROPPED])}
    problems = []
    # (a) every timed step against an independent evaluation of the same frames (batch by batch, weighted by how often a batch was used)
    uses = np.bincount(np.asarray(used, dtype=np.int64), minlength=nb)
    check_batches = [b for b in range(nb) if uses[b]] if wl.name != "c5" else []
    if check_batches:
        expect = np.zeros_like(priv_counts[0])
        for b in check_batches:
            ctx.submit(batches[b], box)
            ctx.reset(wl.chk[0])
            ctx.accumulate(wl.chk[0])
            hb, _ = ctx.read(wl.chk[0])
            expect += hb * np.uint64(uses[b])
        if not np.array_equal(expect, priv_counts[0]):
            problems.append("timed histogram differs from the independent evaluation of the same frames")
        if wl.name == "c2":
            # the full kernel shares the FP32 bracket with the symmetric one; one frame also goes through the mode that re-evaluates
            # EVERY pair in FP64 (a kernel that skipped its re-evaluations would pass the comparison above and fail this one)
            ctx.submit(batches[0][:1], box)
            for a in (wl.accs[0], wl.chk[1]):
                ctx.reset(a)
                ctx.accumulate(a)
            if not np.array_equal(ctx.read(wl.accs[0])[0], ctx.read(wl.chk[1])[0]):
                problems.append("symmetric histogram of one frame differs from the all-FP64 evaluation")
            check["independent_fp64"] = "one frame: symmetric production kernel == every pair re-evaluated in FP64"
        check["independent"] = {"rdf": "every pair re-evaluated in FP64" if wl.name == "c3" else "full N0 x N1 kernel (no symmetry, no cells)", "density": "literal FP64 kernel",
                                "orient": "literal FP64 kernel"}[wl.kind] + f", {len(check_batches)} batches x {F} frames"
    elif wl.name == "c5":
        # 4M atoms: one frame of the dominant pair, cell-binned + symmetric (production) against brute force (about a second)
        ctx.submit(batches[0][:1], box)
        for a in (wl.accs[0], wl.chk[0]):
            ctx.reset(a)
            ctx.accumulate(a)
        h_prod, h_brute = ctx.read(wl.accs[0])[0], ctx.read(wl.chk[0])[0]
        if not np.array_equal(h_prod, h_brute):
            problems.append("cell-binned histogram differs from brute force")
        check["independent"] = "brute-force N0 x N1 kernel on one frame of the dominant pair (SOL-SOL)"
    # (b) the merge: libb2a's all-reduce against the sum of the private histograms gathered over an independent channel
    if world > 1:
        flat = np.concatenate([c.ravel() for c in priv_counts] + [priv_stats[[lib.STAT_FRAMES, lib.STAT_PAIRS, lib.STAT_EXACT]]])
        parts = D.gather_u64(flat)
        total = np.sum(parts, axis=0)
        mflat = np.concatenate([c.ravel() for c in merged_counts] + [merged_stats[[lib.STAT_FRAMES, lib.STAT_PAIRS, lib.STAT_EXACT]]])
        if not np.array_equal(total, mflat):
            problems.append("merged histogram is not the integer sum of the rank-private ones")
        check["merge"] = f"ncclAllReduce inside libb2a == sum of {world} private histograms (torch all_gather)"
    if int(mstats[lib.STAT_FRAMES]) != world * steps * F:
        problems.append(f"frames_merged {int(mstats[lib.STAT_FRAMES])} != {world * steps * F}")
    bad = D.reduce([len(problems)], "sum")[0]
    check["parity"] = "ok" if bad == 0 else ("FAILED: " + "; ".join(problems) if problems else "FAILED on another rank")

    # ---- phase 2: end to end through the C ABI with host buffers ---------------------------------------------------
    reset_all()
    e2e_steps = steps
    pos["next"] = 0
    feed()
    for _ in range(2):
        wl.step(ctx)
        feed()
        wl.read_all(ctx)
    reset_all()
    D.barrier(ctx)
    t0 = time.perf_counter()
    # every step uploads its own frames from pinned memory and reads its own result back; the upload of step s+1 is queued
    # (double-buffered, on the copy stream) before the host blocks on the result of step s, so it overlaps the kernels of s.
    # (the first upload of this loop happened just above: the resident batch is consumed by step 0)
    for s in range(e2e_steps):
        wl.step(ctx)
        feed()                                   # H2D of step s+1 from the next pinned batch
        wl.read_all(ctx)                         # D2H of step s's results (u64 histograms + stats)
    D.barrier(ctx)
    e2e_s = time.perf_counter() - t0
    h2d = F * wl.natoms * 12 + F * 36 + F * 208
    d2h = sum((ctx.acc_size(a) + lib.STAT_NSLOTS) * 8 for a in wl.accs)

    tot_ms_max, e2e_s_max = D.reduce([tot_ms, e2e_s], "max")
    evaluated_all = D.reduce([evaluated_pairs], "sum")[0]
    res = {"wl": wl, "F": F, "steps": steps, "tot_ms": tot_ms, "tot_ms_max": tot_ms_max, "e2e_s_max": e2e_s_max, "timing": timing, "clocks": clocks,
           "evaluated_pairs": evaluated_pairs, "evaluated_all": evaluated_all, "allreduce_ms": allreduce_ms, "check": check, "h2d": h2d, "d2h": d2h,
           "gen_s": t_gen, "wall_s": t_wall, "nccl_in_libb2a": bool(joined), "cpu_baseline": None}
    if with_cpu and rank == 0 and world == 1:
        with tempfile.TemporaryDirectory() as td:
            res["cpu_baseline"] = cpu_baseline(wl, td, batches[0], rdf_sample)
    ctx.close()
    for xh in batches:
        lib.pinned_free(xh)
    return res


def roofline_of(res, world):
    wl, timing, clocks = res["wl"], res["timing"], res["clocks"]
    kname = wl.dominant()
    k_ms, k_n = timing[kname]
    share = k_ms / res["tot_ms"] if res["tot_ms"] else None
    if wl.kind == "rdf":
        sm_mhz = clocks.get("sm_mhz") or 1965.0
        peak = 148 * 128 * sm_mhz * 1e6                          # FP32 issue ceiling at the SM clock seen in this run
        pairs_per_launch = res["evaluated_pairs"] / max(k_n, 1)
        ach = pairs_per_launch * ISSUE_SLOTS_PER_PAIR / (k_ms / max(k_n, 1) * 1e-3) if k_ms > 0 else 0.0
        concurrent = len(wl.accs) > 1
        if concurrent:
            # a sweep over group pairs: the accumulators' kernels run side by side on their own streams, so the per-launch spans
            # overlap; the honest denominator is the whole step (centres and sorting included)
            ach = res["evaluated_pairs"] * ISSUE_SLOTS_PER_PAIR / (res["tot_ms"] * 1e-3)
            share = None
        rf = {"bound": "fp32_issue", "kernel": "k3_rdf_cellw" if wl.name == "c5" else kname, "achieved": ach / 1e12, "peak": peak / 1e12, "unit": "Tinst/s",
              "frac": ach / peak if peak else None,
              "traffic": K3_NCU["dram_bytes_per_frame"] * res["F"] if wl.name == "c2" else None,
              "traffic_unit": f"bytes per k3_rdf launch (ncu, {K3_NCU['source']})",
              "how": f"evaluated pairs per launch x {ISSUE_SLOTS_PER_PAIR} issue slots (BASELINE.md section 4) / mean CUDA-event duration of the "
                     "k3_rdf launches inside the timed region; peak = 148 SM x 128 lanes x SM clock sampled by nvidia-smi during the timed region. "
                     "`frac` is this MODEL fraction; `issue_active` and `inst_per_pair` are what ncu measured on the same kernel",
              "kernel_ms_per_launch": k_ms / max(k_n, 1), "share_of_step": share}
        if concurrent:
            rf["how"] = (f"evaluated pairs of a step x {ISSUE_SLOTS_PER_PAIR} issue slots / CUDA-event duration of the WHOLE step: the {len(wl.accs)} "
                         "accumulators' k3_rdf_cellw (cell-binned, one tile per warp) launches overlap on their own streams, so no per-launch time exists; centres, cell sorting and slab "
                         "prep are inside the denominator; peak as for config 2")
            rf["kernel_ms_per_launch"] = None
        if wl.name == "c2":
            rf.update(issue_active=K3_NCU["issue_active"], inst_per_pair=K3_NCU["inst_per_pair"], ncu_source=K3_NCU["source"])
        elif wl.name == "c5" and "cellw" in K3_NCU:      # the water-water launch of the sweep, captured alone
            cw = K3_NCU["cellw"]
            rf.update(issue_active=cw["issue_active"], inst_per_pair=cw["inst_per_pair"], ncu_source=K3_NCU["source"],
                      traffic=cw["dram_bytes_per_frame"] * res["F"], traffic_unit="bytes per water-water k3_rdf_cellw launch (ncu)")
        return rf
    peak, prov = hbm_peak()
    bytes_per_launch = 12.0 * wl.atoms_per_frame * res["F"]      # one rvec per atom-frame; the histogram stays on chip
    ach = bytes_per_launch / (k_ms / max(k_n, 1) * 1e-3) / 1e9 if k_ms > 0 else 0.0
    return {"bound": "hbm", "kernel": kname, "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "peak_source": prov,
            "traffic": None, "how": "12 B per atom-frame x atoms x frames of one launch / mean CUDA-event duration of that kernel's launches",
            "kernel_ms_per_launch": k_ms / max(k_n, 1), "share_of_step": share,
            "note": ("the whole batch is %.1f MB: launch-latency-bound, not bandwidth-bound" % (bytes_per_launch / 1e6)) if bytes_per_launch < 64e6 else None}


def entry_of(res, world):
    """other_configs entry (also the basis of the headline line)."""
    wl, steps, F = res["wl"], res["steps"], res["F"]
    frames = world * steps * F
    fps = frames / (res["tot_ms_max"] * 1e-3)
    fps_e2e = frames / res["e2e_s_max"]
    timing = res["timing"]
    e = {"workload": wl.text, "frames_per_step_per_gpu": F, "steps": steps, "ms_per_step": res["tot_ms_max"] / steps, "frames_per_s": fps,
         "e2e": {"frames_per_s": fps_e2e, "ms_per_step": 1e3 * res["e2e_s_max"] / steps, "h2d_bytes_per_step": res["h2d"], "d2h_bytes_per_step": res["d2h"]},
         "roofline": roofline_of(res, world),
         "kernel_ms_per_step": {k: t / max(steps, 1) for k, (t, n) in timing.items() if n},
         "check": res["check"], "allreduce_ms": res["allreduce_ms"]}
    if wl.kind == "rdf":
        e["pair_distances_per_s"] = fps * wl.pairs_per_frame
        e["evaluated_pairs_per_s"] = res["evaluated_all"] / (res["tot_ms_max"] * 1e-3)
        e["e2e"]["pair_distances_per_s"] = fps_e2e * wl.pairs_per_frame
        e["group_pairs"] = len(wl.accs)
    else:
        e["atom_frames_per_s"] = fps * wl.atoms_per_frame
    if res["cpu_baseline"]:
        e["cpu_baseline"] = res["cpu_baseline"]
    return e


def host_cpp_block(D, seconds_budget=40.0):
    """The same workload driven by ONE C++ process through the drop-in handle: build/dropin/calc_rdf_region_fused on a config-2
    trajectory file, with 1 device and (N > 1) with B2A_DEVICES = 0..N-1 -- batches dealt to the devices by libb2a, histograms
    merged by b2a_finalize.  Rank 0 runs it while the other ranks wait on the host (their GPUs idle)."""
    exe = os.path.join(DROPIN, "calc_rdf_region_fused")
    out = None
    D.barrier()                                   # every rank's device work is finished; from here the others wait on the host
    if D.rank == 0 and os.path.exists(exe) and os.path.exists(os.path.join(STUBDIR, "libgmxstub.so")):
        from workloads import trajio
        wl = Workload("c2")
        base = "/dev/shm" if os.path.isdir("/dev/shm") and shutil.disk_usage("/dev/shm").free > (2 << 30) else None
        td = tempfile.mkdtemp(prefix="b2a_cpp_", dir=base)
        try:
            nfile = 64
            prefix = os.path.join(td, "c2")
            trajio.write_trajectory(prefix + ".b2t", wl.system.frames(7000, nfile), wl.system.box9())
            trajio.write_topology(prefix + ".b2p", wl.system)
            trajio.write_index(prefix + ".ndx", {"OW": wl.groups[0]["index"]})
            env = dict(os.environ, LD_LIBRARY_PATH=STUBDIR + ":" + os.path.join(ROOT, "gmx2020postanalysis_b200") + ":" + os.environ.get("LD_LIBRARY_PATH", ""))
            for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK", "MASTER_ADDR", "MASTER_PORT"):
                env.pop(k, None)
            out = {"program": "build/dropin/calc_rdf_region_fused (one process, itp::GmxHandle + libb2a)", "trajectory": f"{nfile} distinct frames replayed (raw float format)"}
            outputs = {}
            for ndev in sorted({1, D.world}):
                rep = int(os.environ.get("B2A_CPP_REPEAT", "32")) * ndev        # 2048 frames per device: long enough that ramp-up and drain of the devices do not dominate
                e2 = dict(env, B2A_BATCH=os.environ.get("B2A_BATCH", "32"), GMXSTUB_REPEAT=str(rep), B2A_DEVICES=",".join(str(d) for d in range(ndev)))
                t = time.perf_counter()
                r = subprocess.run([exe, "-f", prefix + ".b2t", "-s", prefix + ".b2p", "-n", prefix + ".ndx", "-o", prefix + f".{ndev}.xvg"], input="0 0\n",
                                   capture_output=True, text=True, env=e2, cwd=td, timeout=600)
                wall = time.perf_counter() - t
                if r.returncode != 0:
                    out[f"error_{ndev}"] = (r.stdout + r.stderr)[-400:]
                    continue
                if os.environ.get("B2A_TRACE"):
                    out[f"trace_{ndev}gpu"] = [ln for ln in r.stderr.splitlines() if ln.startswith("b2a ")]
                loop = [ln for ln in r.stdout.splitlines() if ln.startswith("frame loop")]
                fps = float(loop[-1].split("=")[1].split()[0]) if loop else None
                out[f"frames_per_s_{ndev}gpu"] = fps
                out[f"frames_{ndev}gpu"] = rep * nfile
                out[f"wall_s_{ndev}gpu"] = wall
                outputs[ndev] = (rep, [ln for ln in open(prefix + f".{ndev}.xvg") if not ln.startswith("# ")])
            if len(outputs) == 2:
                # g(r) is normalised by its own tail mean, so replaying the file more often leaves every printed digit unchanged
                out["output_identical_1_vs_n"] = outputs[1][1] == outputs[D.world][1]
                if out.get("frames_per_s_1gpu") and out.get(f"frames_per_s_{D.world}gpu"):
                    out["speedup"] = out[f"frames_per_s_{D.world}gpu"] / out["frames_per_s_1gpu"]
        except Exception as ex:                                   # the block is informative; it must not take the bench line down
            out = {"error": repr(ex)[:300]}
        finally:
            shutil.rmtree(td, ignore_errors=True)
    D.cpu_barrier()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b2a", choices=["b2a", "reference"])
    ap.add_argument("--frames-per-step", type=int, default=0, help="override the workload's batch size")
    ap.add_argument("--ref-sample", type=int, default=1000, help="i-molecules per reference step (~2 s of one core)")
    ap.add_argument("--cpu-sample", type=int, default=8000, help="i-molecules of the cpu_baseline leg (~10 s of one core)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-others", action="store_true", help="skip the other_configs block")
    ap.add_argument("--no-host-cpp", action="store_true", help="skip the one-process C++ run")
    ap.add_argument("--host-cpp-only", action="store_true", help="only the one-process C++ run over --gpus devices (no torch, no ranks)")
    ap.add_argument("--other-steps", type=int, default=5)
    ap.add_argument("--symmetric", type=int, default=1)
    ap.add_argument("--config", default="c2", choices=["c1", "c2", "c3", "c4", "c5"],
                    help="c2 = headline (BASELINE configs[1]); c1, c3, c4, c5 = configs[0], [2], [3], [4]")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference_arm(args)
    if args.host_cpp_only:
        class _One:
            rank, world = 0, args.gpus
            def barrier(self, ctx=None): pass
            def cpu_barrier(self): pass
        print(json.dumps({"host_cpp": host_cpp_block(_One())}))
        return 0
    args.warmup = max(args.warmup, 3)

    D = Dist()
    torch = D.torch
    from gmx2020postanalysis_b200 import lib
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=f"cuda:{D.local}")   # > 126 MB L2
    wl = Workload(args.config)
    if args.frames_per_step:
        wl.F = args.frames_per_step
    if wl.kind == "rdf":
        wl.symmetric = args.symmetric
    res = measure(wl, D, args.steps, args.warmup, not args.no_cpu_baseline, args.cpu_sample, flush)
    head = entry_of(res, D.world)
    others = {}
    if not args.no_others:
        for name in ("c1", "c3", "c4", "c5"):
            if name == args.config:
                continue
            t0 = time.perf_counter()
            try:
                o = Workload(name)
                r = measure(o, D, max(1, min(args.other_steps, 3 if name == "c5" else args.other_steps)), 3, not args.no_cpu_baseline, args.cpu_sample, flush)
                others[name] = entry_of(r, D.world)
                others[name]["bench_seconds"] = time.perf_counter() - t0
            except Exception as ex:
                if D.world > 1:
                    raise                                         # a rank that drops out would hang the others: fail loudly
                others[name] = {"error": repr(ex)[:300]}
    cpp = None if args.no_host_cpp else host_cpp_block(D)

    if D.rank == 0:
        clocks, timing = res["clocks"], res["timing"]
        rdf = wl.kind == "rdf"
        per_frame = wl.pairs_per_frame if rdf else 1.0
        # kernels of libb2a launched inside the timed region, from the launch sequence of one step (profiles/r02_launches.csv shows the
        # same list for config 2): K1 per group, then per RDF accumulator k_fetch_words + k3_slab_count + k3_slab_scatter + k3_rdf
        # (full-mode launch, plus the symmetric one when both sides are the same molecules) + k_add_frames; configs 1 / 4: the
        # streaming kernel + k_add_frames (+ K1 for the dipole vectors).  Cell-binned sweeps (config 5) sort per group on top: there
        # the count falls back to the number of timed kernel spans.
        if wl.kind == "rdf" and wl.name != "c5":
            per_step = len(wl.groups) + len(wl.accs) * (1 + 2 + (2 if (wl.symmetric and wl.name == "c2") else 1) + 1)
        elif wl.kind == "rdf":
            per_step = int(sum(n for k, (t, n) in timing.items() if k != "h2d")) // max(args.steps, 1) + len(wl.accs)
        else:
            per_step = 2 + (1 if wl.kind == "orient" else 0)
        launches = per_step * args.steps
        line = {
            "metric": "rdf_pair_distances_per_s" if rdf else "frames_per_s", "value": head["frames_per_s"] * per_frame,
            "unit": "pair-distances/s" if rdf else "frames/s", "n_gpus": D.world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": head["ms_per_step"], "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32+f64", "data": "synthetic", "ingest_parity": "unpinned",
            "frames_per_s": head["frames_per_s"], "frames_timed": D.world * args.steps * wl.F,
            "config": {"workload": wl.text, "frames_per_step_per_gpu": wl.F, "distinct_batches_per_gpu": wl.D,
                       "parallelism": f"frame-sharded x{D.world}", "l2": "flushed between timed steps (256 MiB write)",
                       "frames": "every step consumes the next of the distinct pinned batches (distinct synthetic frames per rank and batch)",
                       "pairs": "value counts reference-equivalent ordered pairs (N0*N1 per frame); roofline counts only evaluated pairs"},
            "clocks": {"sm_mhz": clocks.get("sm_mhz"), "sm_max_mhz": clocks.get("sm_max_mhz"), "reasons": clocks.get("reasons", []),
                       "samples": clocks.get("samples"), "samples_under_load": clocks.get("samples_under_load"), "power_w_max": clocks.get("power_w_max")},
            "e2e": {"value": head["e2e"]["frames_per_s"] * per_frame, "unit": "pair-distances/s" if rdf else "frames/s",
                    "h2d_bytes_per_step": res["h2d"], "d2h_bytes_per_step": res["d2h"], "frames_per_s": head["e2e"]["frames_per_s"],
                    "ms_per_step": head["e2e"]["ms_per_step"]},
            "gpu_launches": launches,
            "roofline": head["roofline"],
            "kernel_ms": head["kernel_ms_per_step"],
            "allreduce_ms": res["allreduce_ms"], "collective": "ncclAllReduce(u64, sum) inside libb2a (b2a_finalize)" if res["nccl_in_libb2a"] else "none (one device)",
            "check": res["check"],
        }
        if rdf:
            line["evaluated_pairs_per_s"] = head["evaluated_pairs_per_s"]
            line["config"].update(symmetric_i_lt_j=bool(wl.symmetric) and wl.name in ("c2", "c5"), group_pairs=len(wl.accs))
        if res["cpu_baseline"]:
            line["cpu_baseline"] = res["cpu_baseline"]
        if others:
            line["other_configs"] = others
        if cpp:
            line["host_cpp"] = cpp
        print(json.dumps(line))
    D.close()
    return 0


if __name__ == "__main__":
    sys.exit(main())
