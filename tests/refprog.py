"""Helpers shared by the oracle-pinning tests and tests/make_golden.py: run a reference program from
oracle/_ref on a synthetic system and reproduce its output text from oracle.port results."""
from __future__ import annotations

import os
import subprocess

import numpy as np

from workloads import synth, trajio
from oracle import port

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REFDIR = os.path.join(ROOT, "oracle", "_ref")


def have_ref(prog: str) -> bool:
    return os.path.exists(os.path.join(REFDIR, prog)) and os.path.exists(os.path.join(ROOT, "build", "libgmxstub.so"))


def run_program(exe, workdir, system, nframes, group_names, args, out_name, stdin_sel=None, velocities=None, forces=None):
    """Writes the system, runs `exe`, returns the non-comment lines of the output file."""
    os.makedirs(workdir, exist_ok=True)
    groups = {g: system.groups[g] for g in dict.fromkeys(group_names)}
    trj, top, ndx = trajio.write_system(os.path.join(workdir, "sys"), system, nframes, groups=groups, velocities=velocities,
                                        forces=forces)
    names = list(groups)
    sel = stdin_sel if stdin_sel is not None else " ".join(str(names.index(g)) for g in group_names) + "\n"
    out = os.path.join(workdir, out_name)
    cmd = [exe, "-f", trj, "-s", top, "-n", ndx, "-o", out] + [str(a) for a in args]
    env = dict(os.environ)
    env["LD_LIBRARY_PATH"] = (os.path.join(ROOT, "build") + ":" + os.path.join(ROOT, "gmx2020postanalysis_b200") + ":"
                              + env.get("LD_LIBRARY_PATH", ""))
    r = subprocess.run(cmd, input=sel, capture_output=True, text=True, env=env, cwd=workdir)
    if r.returncode != 0:
        raise RuntimeError(f"{exe} failed rc={r.returncode}\n{r.stdout[-2000:]}\n{r.stderr[-2000:]}")
    return data_lines(out)


def data_lines(path):
    """Output lines minus the openWrite provenance header (time stamp, cwd, command line; ipp:378-393)."""
    keep = []
    with open(path) as fp:
        for ln in fp:
            if ln.startswith("# This file was created") or ln.startswith("# Working dir") or ln.startswith("# Command line"):
                continue
            keep.append(ln.rstrip("\n").replace("-nan", "nan"))
    return keep


def f32(v):
    return float(np.float32(v))


def fmt_fixed(v, width, prec):
    if v != v:  # fmt 6.0.0 prints 0/0 as "-nan" LEFT-aligned in the field; data_lines() strips the sign
        return "nan" + " " * (width - 4)
    return f"{v:{width}.{prec}f}"


# ---- text of calc_rdf_region.cpp:117-135 from oracle results ------------------------------------------------
def rdf_region_text(g, nbin, Rbin, dim, low, up):
    low, up = np.float32(low), np.float32(up)
    lines = ["# {} ".format(chr(ord("x") + dim)) +
             "".join(fmt_fixed(float(low + np.float32(i) * (up - low) / np.float32(nbin)), 8, 4) + " " for i in range(nbin + 1))]
    for j in range(Rbin):
        lines.append(fmt_fixed(j * 0.002, 8, 4) + " " + "".join(fmt_fixed(g[j, i], 8, 4) + " " for i in range(nbin)))
    return lines


def oracle_rdf_region(system, nframes, grp0, grp1, nbin=5, low=-100.0, up=100.0, dim=0, dim2=2, low2=-100.0,
                      up2=100.0, Rm=0.0, com=0, return_counts=False):
    """calc_rdf_region.cpp main() restated on top of oracle.port."""
    box = system.box9()
    Lbox = port.lbox_from_box9(box)
    Rm = f32(Rm)
    if Rm < 1e-6:
        Rm = port.rdf_default_rm(Lbox)
    Rbin = port.rdf_rbin(Rm)
    rdf = np.zeros((Rbin, nbin), dtype=np.uint64)
    stats = np.zeros(3, dtype=np.uint64)
    g0, g1 = group_info(system, grp0), group_info(system, grp1)
    for f in range(nframes):
        x = system.frame(f)
        p1 = port.load_position_center(x, g0["index"], g0["napm"], g0["w"][com], Lbox)
        p2 = port.load_position_center(x, g1["index"], g1["napm"], g1["w"][com], Lbox)
        port.rdf_accum(p1, p2, Lbox, nbin, f32(low), f32(up), dim, dim2, f32(low2), f32(up2), Rm, Rbin, rdf, stats)
    if return_counts:
        return rdf, stats, Rm, Rbin
    return rdf_region_text(port.rdf_normalise(rdf), nbin, Rbin, dim, low, up)


group_info = synth.group_info      # lives with the synthetic systems; kept under this name for the tests


# ---- README template.cpp (README.md:129-176) -----------------------------------------------------------------
def oracle_readme_template(system, nframes, grp, nbin=100, dim=2, low=0.0, up=30.0):
    box = system.box9()
    Lbox = port.lbox_from_box9(box)
    low32, up32 = np.float32(low), np.float32(up)
    dbin = float(np.float32(up32 - low32)) / nbin          # (upPos - lowPos) float, / int -> float? see note
    # README.md:127: `double dbin = (upPos - lowPos) / nbin;` -> float / int = float, then widened
    dbin = float(np.float32(np.float32(up32 - low32) / np.float32(nbin)))
    gi = group_info(system, grp)
    hist = np.zeros(nbin, dtype=np.uint64)
    for f in range(nframes):
        x = system.frame(f)
        pc = port.load_position_center(x, gi["index"], gi["napm"], gi["w"][0], Lbox)
        port.density_accum(pc, dim, f32(low), f32(up), dbin, nbin, hist)
    dens = hist.astype(np.float64) / nframes
    lines = []
    for i in range(nbin):
        lines.append(fmt_fixed(float(low32) + dbin / 2 + dbin * i, 8, 5) + " " + fmt_fixed(dens[i], 10, 5))
    return lines, hist


# ---- mofs/calc_orient_mof.cpp main() restated on oracle.port -------------------------------------------------
def oracle_orient_mof(system, nframes, grp, NCM=0, DIMN=2, low=(0.0, 0.0, 0.0), up=(100.0, 100.0, 100.0), c1=0.0, c2=0.0,
                      dr=0.01, CR=0.5, Cr=0.0, nA=90, tp1=1, tp2=2, tp3=3, DIMNOrient=2, method=1, return_counts=False):
    Lbox = port.lbox_from_box9(system.box9())
    gi = group_info(system, grp)
    orient = ncm = None
    for f in range(nframes):
        x = system.frame(f)
        pos = port.load_position(x, gi["index"], gi["napm"], Lbox)
        pc = port.load_position_center(x, gi["index"], gi["napm"], gi["w"][NCM], Lbox)
        orient, ncm, _ = port.orient_accum(pos, pc, Lbox, DIMN, f32(low[DIMN]), f32(up[DIMN]), f32(c1), f32(c2), f32(dr),
                                           f32(CR), f32(Cr), nA, tp1, tp2, tp3, DIMNOrient, method, orient, ncm)
    if return_counts:
        return orient, ncm
    return orient_text(orient, ncm, nA)


def orient_text(orient, ncm, nA):
    """calc_orient_mof.cpp:201-215.  orient: [nAngle, nbin1] counts."""
    dAngle = float(180 // nA)
    lines = []
    for i in range(nA):
        s = f"{(i + 0.5) * dAngle:8.4e}  "
        for j in range(len(ncm)):
            v = float(orient[i, j])
            if ncm[j] > 0:
                v = v / float(ncm[j])
            s += f"{v:8.4e} "
        lines.append(s)
    return lines


def parse_orient_args(args):
    """-flag value list (rvec flags take three values) -> kwargs of oracle_orient_mof."""
    kw, i = {}, 0
    names = {"-NCM": "NCM", "-dim": "DIMN", "-c1": "c1", "-c2": "c2", "-dr": "dr", "-CR": "CR", "-Cr": "Cr", "-nA": "nA",
             "-tp1": "tp1", "-tp2": "tp2", "-tp3": "tp3", "-DIMNOrient": "DIMNOrient", "-method": "method"}
    while i < len(args):
        a = args[i]
        if a in ("-low", "-up"):
            kw[a[1:]] = tuple(float(v) for v in args[i + 1:i + 4])
            i += 4
        elif a == "-dbin":
            i += 4
        else:
            kw[names[a]] = args[i + 1]
            i += 2
    return kw


# ---- mofs/calc_pair_dist_bulk.cpp main() restated on oracle.port ---------------------------------------------
def pair_dist_text(tables, numAs):
    """calc_pair_dist_bulk.cpp:89-107 from per-frame histograms of pairNumber (tables[f][k]) and numA[f]."""
    tot = {}
    for tab, numA in zip(tables, numAs):
        if numA == 0:
            continue
        for k in np.nonzero(tab)[0]:
            tot[int(k)] = tot.get(int(k), 0.0) + float(tab[k]) / int(numA)      # totPair[p.first] += p.second / numA
    total = 0.0
    for k in sorted(tot):
        total += tot[k]
    return [f"{k:8d}\t{tot[k] / total:12.5f}" for k in sorted(tot)]


def oracle_pair_dist_bulk(system, nframes, grp0, grp1, ncm=0, Rmin=0.8, dim=2, low=0.0, up=12.0, return_counts=False):
    Lbox = port.lbox_from_box9(system.box9())
    g0, g1 = group_info(system, grp0), group_info(system, grp1)
    counts, tables, numAs = [], [], []
    for f in range(nframes):
        x = system.frame(f)
        p1 = port.load_position_center(x, g0["index"], g0["napm"], g0["w"][ncm], Lbox)
        p2 = port.load_position_center(x, g1["index"], g1["napm"], g1["w"][ncm], Lbox)
        cnt, numA = port.pair_count(p1, p2, Lbox, f32(Rmin), dim, f32(low), f32(up))
        counts.append(cnt)
        tables.append(np.bincount(cnt[cnt >= 0], minlength=1))
        numAs.append(numA)
    if return_counts:
        return counts, numAs
    return pair_dist_text(tables, numAs)
