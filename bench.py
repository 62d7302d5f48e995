#!/usr/bin/env python
"""bench.py -- headline benchmark: O-O RDF of the 216k-atom synthetic water box (BASELINE.json configs[1]).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b2a|reference] [--config c2|c3|c5]

`--config` selects another BASELINE.json workload for the runs recorded in profiles/ (c3: cation-anion RDF of the 1M-atom ionic
liquid; c5: all 10 group pairs of the 4M-atom mixed box, cell-binned; use --frames-per-step 2 there); the default c2 is the
headline the driver measures.

A "step" is one pass of the whole hot path over one batch of F frames (default 8): centres of both groups
(K1), slab prep, all-pairs minimum-image histogram (K3).  `value` is measured with the batch already resident
in HBM (cached intermediates are dropped before every step, so nothing is reused across steps); `e2e` goes
through the C ABI with HOST buffers: pinned frames -> H2D -> kernels -> D2H of the u64 histogram, every step.
Frames shard across ranks (one process per GPU, weak scaling: F frames per rank per step); the only collective
is one all-reduce of the rank-private histograms at the end.  `--impl reference` times the reference's own CPU
program (oracle/_ref/calc_rdf_region, compiled unmodified) on a bounded i-subset of the same frames.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

ISSUE_SLOTS_PER_PAIR = 19                           # algorithmic work per evaluated pair
# dram__bytes_read.sum + dram__bytes_write.sum of one k3_rdf launch, per frame of the launch, from the ncu --set full
# capture summarised in profiles/r01_k3_ncu.md (the kernel re-reads its 1.15 MB/frame of fixed-point coordinates from L2,
# not from HBM; nothing is written back but the 26 KB histogram)
K3_DRAM_BYTES_PER_FRAME = 4.64e6


class Workload:
    """The RDF job a bench line is quoted on.  `c2` is the headline (BASELINE.json configs[1], what the driver runs); `c3`
    (configs[2], the frame-sharded multi-GPU case) is selectable for the scaling runs recorded in profiles/."""

    def __init__(self, name):
        from workloads import synth
        self.name = name
        if name == "c2":
            self.L, self.seed = 12.913, 20200102
            self.system = synth.water_box(72000, self.L, seed=self.seed, split_atoms=True)
            self.g0 = self.g1 = "OW"
            self.text = "O-O RDF, 216k-atom SPC/E water box (72k O, L=12.913 nm, Rm=L/2, Rbin=3228, nbin=5), calc_rdf_region defaults"
            self.dram_per_frame = K3_DRAM_BYTES_PER_FRAME
        elif name == "c3":
            self.L, self.seed = 22.1, 20200103
            self.system = synth.ionic_liquid(31250, self.L, seed=self.seed)
            self.g0, self.g1 = "CAT", "ANI"
            self.text = ("cation-anion molecular-COM RDF, 1M-atom ionic liquid (31 250 x (25-site cation + 7-site anion), L=22.1 nm, "
                         "Rm=L/2, Rbin=5525, nbin=5), calc_rdf_region defaults")
            self.dram_per_frame = None
        elif name == "c5":
            self.L, self.seed = 34.0, 20200105
            self.system = synth.mixed_box()
            self.g0 = self.g1 = "SOL"                              # the CPU sample is taken on the dominant pair
            self.text = ("mixed-species RDF sweep, 4M atoms (1.2M water + 8k cation + 8k anion + 24k cosolvent, L=34 nm), all 10 group "
                         "pairs, Rm=3 nm, Rbin=1500, nbin=1 (cell-binned K3)")
            self.dram_per_frame = None
        else:
            raise SystemExit(f"unknown --config {name}")
        import refprog
        self.gi0, self.gi1 = refprog.group_info(self.system, self.g0), refprog.group_info(self.system, self.g1)
        self.n0 = len(self.gi0["index"]) // self.gi0["napm"]
        self.n1 = len(self.gi1["index"]) // self.gi1["napm"]
        self.same = self.g0 == self.g1
        # the device job: groups to register and the (grp0, grp1) pairs to accumulate
        if name == "c5":
            names = ["SOL", "CAT", "ANI", "COS"]
            self.groups = [refprog.group_info(self.system, n) for n in names]
            self.pairs = [(a, b) for a in range(4) for b in range(a, 4)]
            self.Rm, self.nbin = 3.0, 1
        else:
            self.groups = [self.gi0, self.gi1]
            self.pairs = [(0, 1)]
            self.Rm, self.nbin = float(np.float32(np.float32(self.L) / np.float32(2))), 5      # calc_rdf_region.cpp:52-57, :23
        nm = [len(g["index"]) // g["napm"] for g in self.groups]
        self.pairs_per_frame = float(sum(nm[a] * nm[b] for a, b in self.pairs))                # reference-equivalent ordered pairs


# ------------------------------------------------------------------------------------------------ clocks
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
                       power_w_max=float(max(pw)))
        return out


# ------------------------------------------------------------------------------------------------ CPU baseline
def cpu_reference_sample(wl, n_i, tmpdir, x0):
    """Runs the reference's own calc_rdf_region (oracle/_ref, 1 thread in the pair loop -- the reference has no
    threading there) on grp0 = the first n_i molecules of group 0 x grp1 = all of group 1, one frame.  Returns
    (pairs/s, kind, secs)."""
    import refprog
    from workloads import trajio
    system, n1 = wl.system, wl.n1
    i0, i1 = wl.gi0["index"], wl.gi1["index"]
    exe = os.path.join(refprog.REFDIR, "calc_rdf_region")
    if refprog.have_ref("calc_rdf_region"):
        prefix = os.path.join(tmpdir, wl.name)
        if not os.path.exists(prefix + ".b2t"):
            trajio.write_trajectory(prefix + ".b2t", x0[None], system.box9())
            trajio.write_topology(prefix + ".b2p", system)
        trajio.write_index(prefix + ".ndx", {"G0_sub": i0[:n_i * wl.gi0["napm"]], "G1": i1})
        env = dict(os.environ, LD_LIBRARY_PATH=os.path.join(ROOT, "build") + ":" + os.environ.get("LD_LIBRARY_PATH", ""),
                   OMP_NUM_THREADS="1")
        cmd = [exe, "-f", prefix + ".b2t", "-s", prefix + ".b2p", "-n", prefix + ".ndx", "-o", prefix + ".xvg"]
        if wl.name == "c5":
            cmd += ["-Rm", "3", "-nbin", "1"]
        t = time.perf_counter()
        r = subprocess.run(cmd, input="0 1\n", capture_output=True, text=True, env=env, cwd=tmpdir)
        dt = time.perf_counter() - t
        if r.returncode != 0:
            raise RuntimeError("reference program failed: " + r.stderr[-500:])
        return n_i * n1 / dt, "reference", dt
    # the reference binary did not travel: time the oracle port instead (single thread)
    from oracle import port
    os.environ["OMP_NUM_THREADS"] = "1"
    Lbox = port.lbox_from_box9(system.box9())
    Rm = port.rdf_default_rm(Lbox)
    p0 = port.load_position_center(x0, wl.gi0["index"][:n_i * wl.gi0["napm"]], wl.gi0["napm"], wl.gi0["w"][0], Lbox)
    p1 = port.load_position_center(x0, wl.gi1["index"], wl.gi1["napm"], wl.gi1["w"][0], Lbox)
    t = time.perf_counter()
    port.rdf_accum(p0, p1, Lbox, 5, -100.0, 100.0, 0, 2, -100.0, 100.0, Rm, port.rdf_rbin(Rm))
    dt = time.perf_counter() - t
    return n_i * n1 / dt, "port", dt


def cpu_reference_all_cores(wl, n_i, procs, tmpdir, x0):
    """The reference has no threading in the pair loop; the most its code can use of a host is one PROCESS per core, each on
    its own share of the i-molecules (what a user would do by hand with index groups).  Returns aggregate pairs/s, or None
    where the reference binaries did not travel."""
    import refprog
    from workloads import trajio
    if not refprog.have_ref("calc_rdf_region"):
        return None
    system, n1 = wl.system, wl.n1
    i0, i1, na = wl.gi0["index"], wl.gi1["index"], wl.gi0["napm"]
    n_i = min(n_i, wl.n0 // procs)
    exe = os.path.join(refprog.REFDIR, "calc_rdf_region")
    prefix = os.path.join(tmpdir, wl.name)
    if not os.path.exists(prefix + ".b2t"):
        trajio.write_trajectory(prefix + ".b2t", x0[None], system.box9())
        trajio.write_topology(prefix + ".b2p", system)
    env = dict(os.environ, LD_LIBRARY_PATH=os.path.join(ROOT, "build") + ":" + os.environ.get("LD_LIBRARY_PATH", ""), OMP_NUM_THREADS="1")
    cmds = []
    for q in range(procs):
        ndx = os.path.join(tmpdir, f"p{q}.ndx")
        trajio.write_index(ndx, {"G0_sub": i0[q * n_i * na:(q + 1) * n_i * na], "G1": i1})
        cmds.append([exe, "-f", prefix + ".b2t", "-s", prefix + ".b2p", "-n", ndx, "-o", os.path.join(tmpdir, f"p{q}.xvg")]
                    + (["-Rm", "3", "-nbin", "1"] if wl.name == "c5" else []))
    t = time.perf_counter()
    ps = [subprocess.Popen(cmd, stdin=subprocess.PIPE, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, env=env, cwd=tmpdir, text=True)
          for cmd in cmds]
    for pr in ps:
        pr.stdin.write("0 1\n"); pr.stdin.close()
    rcs = [pr.wait() for pr in ps]
    dt = time.perf_counter() - t
    if any(rcs):
        return None
    return procs * n_i * n1 / dt


def run_reference_arm(args):
    """--impl reference: same metric/config/unit, each step a bounded sample (n_i oxygens x all) of the workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    wl = Workload(args.config)
    system = wl.system
    x0 = system.frame(0)
    n_i = max(1, min(args.ref_sample, wl.n0, int(7.2e7 // wl.n1)))      # ~1-2 s of one core per step
    rates, kind = [], "reference"
    with tempfile.TemporaryDirectory() as td:
        for s in range(args.warmup + args.steps):
            rate, kind, dt = cpu_reference_sample(wl, n_i, td, x0)
            if s >= args.warmup:
                rates.append((rate, dt))
        procs = os.cpu_count() or 1
        all_cores = cpu_reference_all_cores(wl, n_i, procs, td, x0) if kind == "reference" else None
    tot_pairs = sum(r * d for r, d in rates)
    tot_s = sum(d for _, d in rates)
    value = tot_pairs / tot_s
    line = {
        "impl": "reference", "metric": "rdf_pair_distances_per_s", "value": value, "unit": "pair-distances/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * tot_s / len(rates),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "frames_per_s": value / wl.pairs_per_frame,
        "config": {"workload": wl.text, "sample": f"{n_i} of {wl.n0} i-molecules x all {wl.n1} j-molecules of frame 0 per step"},
        "cpu_baseline": {"value": value, "unit": "pair-distances/s", "cores": 1, "kind": kind,
                         "sample": f"{n_i} x {wl.n1} ordered pairs per step, {len(rates)} steps, whole-process wall time "
                                   "(includes reading the frame)", "host_cpus": os.cpu_count()},
        "e2e": {"value": value, "unit": "pair-distances/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    if all_cores:
        # not the reference's own mode of operation (its frame loop is single-threaded): reported beside `value`, never instead
        line["cpu_all_cores"] = {"value": all_cores, "unit": "pair-distances/s", "processes": procs,
                                 "sample": f"{procs} concurrent reference processes, {min(n_i, wl.n0 // procs)} x {wl.n1} ordered pairs each, wall time of the slowest"}
    print(json.dumps(line))
    return 0


# ------------------------------------------------------------------------------------------------ B200 arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b2a", choices=["b2a", "reference"])
    ap.add_argument("--frames-per-step", type=int, default=8)
    ap.add_argument("--ref-sample", type=int, default=1000, help="i-molecules per reference step (~2 s of one core)")
    ap.add_argument("--cpu-sample", type=int, default=8000, help="i-molecules of the cpu_baseline leg (~15 s of one core)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--symmetric", type=int, default=1)
    ap.add_argument("--config", default="c2", choices=["c2", "c3", "c5"],
                    help="c2 = headline (BASELINE configs[1]); c3 = configs[2]; c5 = configs[4] (use --frames-per-step 2)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference_arm(args)
    args.warmup = max(args.warmup, 3)

    import torch
    import torch.distributed as dist
    from gmx2020postanalysis_b200 import dist as b2d, lib
    from workloads import synth

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the B200 path has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    F = args.frames_per_step
    wl = Workload(args.config)
    system, LBOX = wl.system, wl.L
    natoms = system.natoms
    # rank-private, distinct frames in pinned host memory (frames shard contiguously across ranks)
    xh = torch.empty((F, natoms, 3), dtype=torch.float32).pin_memory()
    system.frames(rank * F, F, out=xh.numpy())
    xh2 = torch.empty((F, natoms, 3), dtype=torch.float32).pin_memory()      # second host buffer of the end-to-end loop
    xh2.copy_(xh)
    box = system.box9()

    ctx = lib.Context(natoms, F, device=local)
    for g, gi in enumerate(wl.groups):
        ctx.set_group(g, gi["index"], gi["napm"], gi["mass"], gi["charge"])
    Rm = wl.Rm
    Rbin = int(np.float32(Rm) / 0.002)                              # :58
    nbin = wl.nbin
    accs = []
    for g0, g1 in wl.pairs:
        a_ = ctx.rdf_create(grp0=g0, grp1=g1, com0=0, com1=0, nbin=nbin, low=-100.0, up=100.0, dim=0, dim2=2, low2=-100.0,
                            up2=100.0, Rm=Rm, Rbin=Rbin)
        ctx.rdf_tune(a_, args.symmetric, 0)
        accs.append(a_)
    ncounters = nbin * Rbin

    def accumulate_all():
        for a_ in accs:
            ctx.accumulate(a_)

    def read_all():
        out = [ctx.read(a_, ncounters) for a_ in accs]
        return sum(int(c.sum()) for c, _ in out), np.sum([st for _, st in out], axis=0)
    stream = torch.cuda.ExternalStream(ctx.stream, device=local)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=f"cuda:{local}")   # > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(local)
        ctx.sync()

    pairs_step = float(F) * wl.pairs_per_frame   # reference-equivalent ordered pairs per rank per step

    # ---- phase 1: device-resident ------------------------------------------------------------------------
    ctx.submit(xh.numpy(), box)
    ctx.sync()
    for _ in range(args.warmup):
        ctx.invalidate()
        accumulate_all()
    ctx.sync()
    for a_ in accs:
        ctx.reset(a_)
    ctx.timing_enable(True)
    ctx.timing_read()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    sampler = ClockSampler(local)
    sampler.start()
    barrier()
    t_wall = time.perf_counter()
    for s in range(args.steps):
        with torch.cuda.stream(stream):
            flush.zero_()                        # L2 flush between timed iterations (outside the event pair)
        ev[s][0].record(stream)
        ctx.invalidate()
        accumulate_all()
        ev[s][1].record(stream)
    barrier()
    t_wall = time.perf_counter() - t_wall
    clocks = sampler.stop()
    step_ms = [a.elapsed_time(b) for a, b in ev]
    tot_ms = float(sum(step_ms))
    timing = ctx.timing_read()
    ctx.timing_enable(False)
    _, stats_dev = read_all()
    evaluated_pairs = float(stats_dev[lib.STAT_PAIRS])           # physically evaluated by this rank over the timed steps

    # ---- phase 2: end to end through the C ABI with host buffers -------------------------------------------
    def reset_all():
        for a_ in accs:
            ctx.reset(a_)
    reset_all()
    for _ in range(2):
        ctx.submit(xh.numpy(), box)
        accumulate_all()
        read_all()
    reset_all()
    barrier()
    t0 = time.perf_counter()
    # every step uploads its own frames from pinned memory and reads its own result back; the upload of step s+1 is queued
    # (double-buffered, on the copy stream) before the host blocks on the result of step s, so it overlaps the kernels of s
    ctx.submit(xh.numpy(), box)                  # H2D of step 0
    for s in range(args.steps):
        accumulate_all()
        if s + 1 < args.steps:
            ctx.submit(xh2.numpy(), box)         # H2D of step s+1 (second pinned buffer: ping-pong, as the drop-in batcher does)
            xh, xh2 = xh2, xh
        read_all()                               # D2H of step s's results (u64 histograms + stats)
    barrier()
    e2e_s = time.perf_counter() - t0
    h2d = F * natoms * 12 + F * 36 + F * 72
    d2h = sum((ctx.acc_size(a_) + lib.STAT_NSLOTS) * 8 for a_ in accs)

    # ---- the single collective: merge rank-private histograms ------------------------------------------------
    t0 = time.perf_counter()
    for a_ in accs:
        b2d.allreduce_accumulator(ctx, a_)
    allreduce_ms = 1e3 * (time.perf_counter() - t0)
    merged_sum, mstats = read_all()
    mstats[lib.STAT_FRAMES] = mstats[lib.STAT_FRAMES] // len(accs)          # every accumulator saw every frame

    # ---- max over ranks -----------------------------------------------------------------------------------------
    red = torch.tensor([tot_ms, e2e_s, evaluated_pairs], dtype=torch.float64, device=f"cuda:{local}")
    if world > 1:
        mx = red.clone(); dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = red.clone(); dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        tot_ms_max, e2e_s_max, evaluated_all = float(mx[0]), float(mx[1]), float(sm[2])
    else:
        tot_ms_max, e2e_s_max, evaluated_all = tot_ms, e2e_s, evaluated_pairs

    if rank == 0:
        value = world * args.steps * pairs_step / (tot_ms_max * 1e-3)
        e2e_value = world * args.steps * pairs_step / e2e_s_max
        k3_ms, k3_n = timing["k3_rdf"]
        k1_ms, k1_n = timing["k1_centers"]
        prep_ms, prep_n = timing["k3_prep"]
        sm_mhz = clocks.get("sm_mhz") or 1965.0
        peak_inst = 148 * 128 * sm_mhz * 1e6                        # FP32 issue ceiling at the SM clock seen in this run
        k3_pairs_per_launch = evaluated_pairs / max(k3_n, 1)
        achieved_inst = k3_pairs_per_launch * ISSUE_SLOTS_PER_PAIR / (k3_ms / max(k3_n, 1) * 1e-3) if k3_ms > 0 else 0.0
        line = {
            "metric": "rdf_pair_distances_per_s", "value": value, "unit": "pair-distances/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": tot_ms_max / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32+f64", "data": "synthetic",
            "frames_per_s": value / wl.pairs_per_frame,
            "evaluated_pairs_per_s": evaluated_all / (tot_ms_max * 1e-3),
            "config": {"workload": wl.text, "frames_per_step_per_gpu": F, "parallelism": f"frame-sharded x{world}",
                       "l2": "flushed between timed steps (256 MiB write)", "symmetric_i_lt_j": bool(args.symmetric) and wl.same, "group_pairs": len(accs),
                       "pairs": "value counts reference-equivalent ordered pairs (N0*N1 per frame); roofline counts only evaluated pairs"},
            "clocks": {"sm_mhz": clocks.get("sm_mhz"), "sm_max_mhz": clocks.get("sm_max_mhz"), "reasons": clocks.get("reasons", []),
                       "samples": clocks.get("samples"), "power_w_max": clocks.get("power_w_max")},
            "e2e": {"value": e2e_value, "unit": "pair-distances/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "frames_per_s": e2e_value / wl.pairs_per_frame, "ms_per_step": 1e3 * e2e_s_max / args.steps},
            "gpu_launches": int(sum(n for k, (t, n) in timing.items() if k in ("k1_centers", "k3_rdf")) + 2 * prep_n + args.steps),
            "roofline": {"bound": "fp32_issue", "kernel": "k3_rdf", "achieved": achieved_inst / 1e12, "peak": peak_inst / 1e12,
                         "unit": "Tinst/s", "frac": achieved_inst / peak_inst if peak_inst else None,
                         "traffic": wl.dram_per_frame * F if wl.dram_per_frame else None, "traffic_unit": "bytes per k3_rdf launch (ncu, profiles/r01_k3_ncu.md)",
                         "how": f"evaluated pairs per launch x {ISSUE_SLOTS_PER_PAIR} issue slots (BASELINE.md section 4) / mean CUDA-event "
                                "duration of the k3_rdf launches inside the timed region; peak = 148 SM x 128 lanes x SM clock "
                                "sampled by nvidia-smi during the timed region",
                         "kernel_ms_per_launch": k3_ms / max(k3_n, 1), "share_of_step": k3_ms / tot_ms if tot_ms else None},
            "kernel_ms": {"k1_centers": k1_ms / max(args.steps, 1), "k3_prep": prep_ms / max(args.steps, 1),
                          "k3_rdf": k3_ms / max(args.steps, 1)},
            "allreduce_ms": allreduce_ms,
            "check": {"counted_pairs_merged": int(merged_sum), "fp64_reevaluated": int(mstats[lib.STAT_EXACT]),
                      "dropped_oob": int(mstats[lib.STAT_DROPPED]), "frames_merged": int(mstats[lib.STAT_FRAMES])},
        }
        if world == 1 and not args.no_cpu_baseline:
            with tempfile.TemporaryDirectory() as td:
                n_cpu = max(1, min(args.cpu_sample, wl.n0, int(6.0e8 // wl.n1)))    # ~10 s of one core
                rate, kind, dt = cpu_reference_sample(wl, n_cpu, td, xh.numpy()[0])
            line["cpu_baseline"] = {"value": rate, "unit": "pair-distances/s", "cores": 1, "kind": kind,
                                    "sample": f"{n_cpu} of {wl.n0} i-molecules x all {wl.n1} j-molecules of one frame "
                                              f"({dt:.1f} s, whole-process wall time)", "host_cpus": os.cpu_count(),
                                    "frames_per_s": rate / wl.pairs_per_frame}
        print(json.dumps(line))
    ctx.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
