"""End-to-end ingestion benchmark: XTC file -> pinned K-frame batches -> device RDF, on the C2 system
(72 000 SPC/E = 216 000 atoms per frame).

  1. host decode alone: frames/s of the stand-in's XTC decoder vs number of decode threads (K frames per batch);
  2. the ported program build/dropin/calc_rdf_region_fused reading that XTC file (decode of batch k+1 overlaps the device
     work of batch k; ping-pong pinned buffers), against the same program reading the raw-float file.
Prints one JSON line.      python tools/bench_ingest.py [--frames 64]
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from workloads import synth, trajio  # noqa: E402

_P = C.c_void_p


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=64)
    ap.add_argument("--nmol", type=int, default=72000)
    args = ap.parse_args()
    L = C.CDLL(os.path.join(ROOT, "build", "libgmxstub.so"))
    L.gmxstub_write_xtc.argtypes = [C.c_char_p, C.c_int, C.c_int, _P, _P, C.c_int, _P, C.c_float]
    L.gmxstub_read_batches.argtypes = [C.c_char_p, C.c_int, C.c_int, _P, _P, _P]
    box_l = 12.913 * (args.nmol / 72000.0) ** (1.0 / 3.0)
    system = synth.water_box(args.nmol, box_l, seed=20200102, split_atoms=True)
    K = args.frames
    out = {"config": f"XTC ingestion, {system.natoms} atoms/frame, {K} frames", "host_cpus": os.cpu_count()}
    with tempfile.TemporaryDirectory() as td:
        base = system.frames(0, 8)
        x = np.concatenate([base] * ((K + 7) // 8))[:K]
        box = np.ascontiguousarray(system.box9(), np.float32).reshape(1, 9)
        xtc, raw = os.path.join(td, "c2.xtc"), os.path.join(td, "c2.b2t")
        t = time.perf_counter()
        assert L.gmxstub_write_xtc(xtc.encode(), x.shape[1], K, x.ctypes.data, box.ctypes.data, 1, None, 1000.0) == 0
        out["encode_frames_per_s"] = K / (time.perf_counter() - t)
        out["xtc_bytes_per_frame"] = os.path.getsize(xtc) / K
        out["raw_bytes_per_frame"] = x.nbytes / K
        bx = np.zeros_like(x); bb = np.zeros((K, 9), np.float32); bt = np.zeros(K, np.float32)
        dec = {}
        for th in (1, 2, 4, 8, 16):
            if th > (os.cpu_count() or 1):
                continue
            os.environ["GMXSTUB_DECODE_THREADS"] = str(th)
            L.gmxstub_read_batches(xtc.encode(), 16, min(K, 16), bx.ctypes.data, bb.ctypes.data, bt.ctypes.data)   # page in
            t = time.perf_counter()
            n = L.gmxstub_read_batches(xtc.encode(), 16, K, bx.ctypes.data, bb.ctypes.data, bt.ctypes.data)
            dt = time.perf_counter() - t
            dec[str(th)] = {"frames_per_s": n / dt, "decoded_MB_per_s": x.nbytes / dt / 1e6}
        del os.environ["GMXSTUB_DECODE_THREADS"]
        out["decode"] = dec
        # end to end through the drop-in handle: steady-state rate from two trajectory lengths (the fixed cost -- process
        # start, CUDA context, topology, output file -- cancels in the difference)
        exe = os.path.join(ROOT, "build", "dropin", "calc_rdf_region_fused")
        if os.path.exists(exe):
            xq = bx.copy()
            K1 = max(8, K // 8)
            trajio.write_trajectory(raw, xq, system.box9())
            short_raw, short_xtc = os.path.join(td, "s.b2t"), os.path.join(td, "s.xtc")
            trajio.write_trajectory(short_raw, xq[:K1], system.box9())
            assert L.gmxstub_write_xtc(short_xtc.encode(), x.shape[1], K1, x.ctypes.data, box.ctypes.data, 1, None, 1000.0) == 0
            trajio.write_topology(os.path.join(td, "c2.b2p"), system)
            trajio.write_index(os.path.join(td, "c2.ndx"), {"OW": system.groups["OW"]})
            env = dict(os.environ, B2A_BATCH="8",
                       LD_LIBRARY_PATH=os.path.join(ROOT, "build") + ":" + os.path.join(ROOT, "gmx2020postanalysis_b200") + ":"
                       + os.environ.get("LD_LIBRARY_PATH", ""))

            import re

            def run(trj, tag):
                """-> (whole-process wall time, seconds the program itself measured around its frame loop), best of two"""
                best, loop = 1e30, 1e30
                for rep in range(2):
                    t = time.perf_counter()
                    r = subprocess.run([exe, "-f", trj, "-s", os.path.join(td, "c2.b2p"), "-n", os.path.join(td, "c2.ndx"), "-o",
                                        os.path.join(td, tag + ".xvg")], input="0 0\n", capture_output=True, text=True, env=env, cwd=td)
                    best = min(best, time.perf_counter() - t)
                    if r.returncode != 0:
                        raise RuntimeError(r.stderr[-800:])
                    m = re.search(r"frame loop ([0-9.]+) s", r.stdout)
                    if m:
                        loop = min(loop, float(m.group(1)))
                return best, loop
            e2e = {}
            for kind, long_f, short_f in (("xtc", xtc, short_xtc), ("raw", raw, short_raw)):
                (t_short, _), (t_long, loop_long) = run(short_f, kind + "_s"), run(long_f, kind)
                e2e[kind] = {"wall_s_%d_frames" % K1: t_short, "wall_s_%d_frames" % K: t_long,
                             "frame_loop_s": loop_long, "frames_per_s_frame_loop": K / loop_long,
                             "frames_per_s_from_wall_difference": (K - K1) / max(t_long - t_short, 1e-9)}
            e2e["outputs_identical"] = open(os.path.join(td, "xtc.xvg")).read().split("\n")[3:] == open(os.path.join(td, "raw.xvg")).read().split("\n")[3:]
            out["fused_program"] = e2e
            # an HBM-class analysis (the README's density template, K2 over all atoms): the device needs ~0.05 ms per frame, so
            # this program runs at whatever rate the frames arrive (decode threads / file read, then the H2D copy)
            exe2 = os.path.join(ROOT, "build", "dropin", "readme_template_fused")
            if os.path.exists(exe2):
                trajio.write_index(os.path.join(td, "all.ndx"), {"SOL": np.arange(system.natoms, dtype=np.int32)})
                dens = {}
                for batch in ("8", "16", "32"):
                    for kind, trj in (("xtc", xtc), ("raw", raw)):
                        loop = 1e30
                        for rep in range(2):
                            r = subprocess.run([exe2, "-f", trj, "-s", os.path.join(td, "c2.b2p"), "-n", os.path.join(td, "all.ndx"), "-o",
                                                os.path.join(td, "dens_" + kind + ".xvg"), "-nbin", "100", "-dim", "2", "-low", "0", "-up", "%g" % box_l],
                                               input="0\n", capture_output=True, text=True, env=dict(env, B2A_BATCH=batch), cwd=td)
                            if r.returncode != 0:
                                raise RuntimeError(r.stderr[-800:])
                            m = re.search(r"frame loop ([0-9.]+) s", r.stdout)
                            loop = min(loop, float(m.group(1)))
                        dens.setdefault(kind, {})["batch_" + batch] = K / loop
                dens["outputs_identical"] = open(os.path.join(td, "dens_xtc.xvg")).read().split("\n")[3:] == open(os.path.join(td, "dens_raw.xvg")).read().split("\n")[3:]
                out["density_program_frames_per_s"] = dens
    print(json.dumps(out))


if __name__ == "__main__":
    main()
