"""Drop-in sweep: EVERY program of the reference that still compiles is run twice on the same synthetic
trajectory -- once as built from the reference's own headers (oracle/_ref/<prog>, CPU) and once as built from the same
unmodified source against include/itp/gmx + libb2a.so (build/dropin/<prog>, loaders on the GPU) -- and every file the two
runs write must be identical (minus the openWrite provenance header).  This module only describes the runs; it never
reads /root/reference (the option names below were collected from the sources once and are part of the table)."""
from __future__ import annotations

import os

import numpy as np

import refprog
from workloads import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DROPIN = os.path.join(ROOT, "build", "dropin")

# program -> (number of index groups, needs velocities, needs forces, scalar options it understands, rvec options)
# Options get region-like values scaled to the box (see _value); everything else keeps the program's default.
PROGRAMS = {
    "calc_areaDensityDistributionPore": (1, 0, 0, ["-up1", "-low1", "-up2", "-low2", "-up3", "-low3"], []),
    "calc_density_distribution_in_pore": (1, 0, 0, ["-upPos", "-lowPos", "-centerX", "-centerY", "-Rmof"], []),
    "calc_density_region": (1, 0, 0, ["-up", "-low", "-up2", "-low2"], []),
    "calc_density_time": (1, 0, 0, ["-upPos", "-lowPos", "-centerX", "-centerY", "-Rmof"], []),
    "calc_distance": (2, 0, 0, ["-up", "-low"], []),
    "calc_distance_mof": (1, 0, 0, [], []),
    "calc_ecacf": (1, 1, 0, [], []),
    "calc_ecacf-in-pore": (1, 1, 0, [], ["-low", "-up"]),
    "calc_energy": (2, 0, 0, ["-Cr", "-CR", "-low", "-up"], []),
    "calc_energy_time": (2, 0, 0, ["-Cr", "-CR", "-low", "-up"], []),
    "calc_force": (1, 0, 1, [], []),                       # reads atom 10800 of every frame: needs a system that large
    "calc_force_distribution": (1, 0, 1, [], ["-low", "-up"]),
    "calc_free_bond_bulk": (2, 0, 0, ["-Rmin", "-low", "-up"], []),
    "calc_inpore_path": (2, 0, 0, ["-upPos", "-lowPos", "-centerX", "-centerY", "-CR", "-Cr", "-Rmin"], []),
    "calc_inpore_path_freeBond": (2, 0, 0, ["-upPos", "-lowPos", "-centerX", "-centerY", "-CR", "-Cr", "-Rmin"], []),
    "calc_kineticEnergy": (1, 1, 0, ["-up1", "-low1", "-up2", "-low2"], []),
    "calc_msd_bin": (1, 0, 0, ["-up", "-low"], []),
    "calc_msd_mof2": (1, 0, 0, ["-lowx", "-lowy", "-lowz", "-upx", "-upy", "-upz", "-c1", "-c2", "-CR", "-Cr"], []),
    "calc_numdens_pore": (1, 0, 0, [], ["-low", "-up"]),
    "calc_orient_mof": (1, 0, 0, ["-c1", "-c2", "-CR", "-Cr"], ["-low", "-up"]),
    "calc_pair_dist": (2, 0, 0, ["-up", "-low", "-centerX", "-centerY", "-Rmin"], []),
    "calc_pair_dist_bulk": (2, 0, 0, ["-Rmin", "-low", "-up"], []),
    "calc_pull_force": (1, 0, 1, ["-upPos", "-lowPos", "-centerX", "-centerY", "-CR", "-Cr"], []),
    "calc_qmsd_pore3": (1, 0, 0, [], ["-low", "-up"]),
    "calc_qmsd_pore3_thread": (1, 0, 0, [], ["-low", "-up"]),
    "calc_radius_distribution_in_pore": (1, 0, 0, ["-upPos", "-lowPos", "-centerX", "-centerY", "-Rmof"], []),
    "calc_radius_distribution_in_pore_time": (1, 0, 0, ["-centerX", "-centerY", "-Cr", "-CR", "-x1", "-x2", "-y1", "-y2", "-z1", "-z2"], []),
    "calc_radius_integration_distribution_in_pore": (1, 0, 0, ["-centerX", "-centerY", "-Cr", "-CR"], []),
    "calc_rdf_mof_pore": (2, 0, 0, ["-up", "-low", "-centerX", "-centerY", "-Rmof"], []),
    "calc_rdf_region": (2, 0, 0, ["-up", "-low", "-low2", "-up2"], []),
    "density_slit_jhl": (1, 0, 0, ["-up1", "-low1", "-up2", "-low2"], []),
    "template1": (1, 0, 0, ["-up", "-low"], []),
}


def _value(opt, L):
    if opt.startswith("-up") or opt in ("-x2", "-y2", "-z2"):
        return 0.9 * L
    if opt.startswith("-low") or opt in ("-x1", "-y1", "-z1"):
        return 0.1 * L
    if opt in ("-centerX", "-centerY", "-c1", "-c2"):
        return 0.5 * L
    if opt in ("-Rmof", "-CR"):
        return 0.45 * L
    if opt == "-Cr":
        return 0.0
    if opt == "-Rmin":
        return 0.7
    raise KeyError(opt)


# programs whose own arrays are overrun by the generic region values get hand-made arguments (as a function of the box length)
ARGS_OVERRIDE = {
    # no r <= CR test in the program (:88-92): the box must lie inside the circle, or rdf[meZ] is out of bounds and Eigen aborts
    "calc_radius_distribution_in_pore_time": lambda L: ["-centerX", 0.5 * L, "-centerY", 0.5 * L, "-Cr", 0.0, "-CR", 0.45 * L, "-dr", 0.03,
                                                        "-x1", 0.3 * L, "-x2", 0.7 * L, "-y1", 0.3 * L, "-y2", 0.7 * L, "-z1", 0.1 * L, "-z2", 0.9 * L],
}


def case(name):
    """-> (system, group names, nframes, argument list, velocities, forces)"""
    ngrp, need_v, need_f, scalars, rvecs = PROGRAMS[name]
    K = 6
    if name == "calc_force":
        system = synth.water_box(3700, 4.8, seed=9105)
    elif ngrp == 1:
        system = synth.water_box(600, 2.7, seed=9101)
    else:
        system = synth.ionic_liquid(90, 3.4, seed=9102)
    groups = ["SOL"] if ngrp == 1 else ["CAT", "ANI"]
    L = float(np.float32(system.box9()[0]))
    args = []
    for o in scalars:
        args += [o, f"{_value(o, L):.4f}"]
    for o in rvecs:
        args += [o] + [f"{_value(o, L):.4f}"] * 3
    if name in ARGS_OVERRIDE:
        args = [a if isinstance(a, str) else f"{a:.4f}" for a in ARGS_OVERRIDE[name](L)]
    rng = np.random.default_rng(9103)
    v = rng.normal(size=(K, system.natoms, 3)).astype(np.float32) if need_v else None
    f = rng.normal(size=(K, system.natoms, 3)).astype(np.float32) if need_f else None
    return system, groups, K, args, v, f


def run(exe, workdir, name):
    """Runs one build of one program; returns ("ok", {file: lines}) or ("fail", return code)."""
    system, groups, K, args, v, f = case(name)
    os.makedirs(workdir, exist_ok=True)
    before = set(os.listdir(workdir))
    try:
        refprog.run_program(exe, workdir, system, K, groups, args, "o.xvg", velocities=v, forces=f)
    except FileNotFoundError:
        pass                                     # the program writes no -o file (template1) or only files of its own
    except RuntimeError as e:
        return "fail", str(e).split("\n")[0].rsplit("rc=", 1)[-1]
    out = {}
    for fn in sorted(set(os.listdir(workdir)) - before):
        if fn.startswith("sys."):
            continue
        out[fn] = refprog.data_lines(os.path.join(workdir, fn))
    return "ok", out
