"""Drop-in check: the reference's UNMODIFIED program sources, recompiled against include/itp/gmx + libb2a.so
(build/dropin/*), must write the same output files as the same sources compiled against the reference's own
headers (oracle/_ref/*, or the committed golden outputs where those binaries are absent)."""
import os

import pytest

import dropin_sweep as SW
import golden_cases as GC
import refprog

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DROPIN = os.path.join(ROOT, "build", "dropin")
GOLD = os.path.join(ROOT, "tests", "golden")


def _dropin(name):
    exe = os.path.join(DROPIN, name)
    if not os.path.exists(exe):
        pytest.skip(f"{exe} not built (needs /root/reference at build time)")
    return exe


def _golden(name):
    return open(os.path.join(GOLD, name + ".txt")).read().split("\n")[:-1]


@pytest.mark.parametrize("name", sorted(GC.program_cases()))
@pytest.mark.parametrize("batch", ["16", "2"])
def test_unmodified_program_matches_reference_output(name, batch, tmp_path, monkeypatch):
    """batch=2 forces several pinned-batch refills inside one run (H4: per-frame API over K-frame batches)."""
    case = GC.program_cases()[name]
    monkeypatch.setenv("B2A_BATCH", batch)
    got = refprog.run_program(_dropin(case["exe"]), str(tmp_path), case["system"](), case["nframes"], case["groups"],
                              case["args"], "out.xvg")
    assert got == _golden(name)


def test_fused_port_matches_reference_output(tmp_path, monkeypatch):
    for name in ("rdf_water_oo", "rdf_il_slabs"):
        case = GC.program_cases()[name]
        monkeypatch.setenv("B2A_BATCH", "2")
        got = refprog.run_program(_dropin("calc_rdf_region_fused"), str(tmp_path / name), case["system"](), case["nframes"],
                                  case["groups"], case["args"], "out.xvg")
        assert got == _golden(name)


@pytest.mark.parametrize("name,exe", [("density_template_c1", "readme_template_fused"), ("orient_water", "calc_orient_mof_fused"),
                                      ("orient_il_cross", "calc_orient_mof_fused"), ("pair_dist_il", "calc_pair_dist_bulk_fused")])
@pytest.mark.parametrize("batch", ["16", "3"])
def test_other_fused_ports_match_reference_output(name, exe, batch, tmp_path, monkeypatch):
    """The density template, calc_orient_mof and calc_pair_dist_bulk ported to the fused accumulators (K2 / K4 / K5: same
    command line, frame loop replaced by one accumulate per resident batch) write the reference program's output file."""
    case = GC.program_cases()[name]
    monkeypatch.setenv("B2A_BATCH", batch)
    got = refprog.run_program(_dropin(exe), str(tmp_path), case["system"](), case["nframes"], case["groups"], case["args"], "out.xvg")
    assert got == _golden(name)


def _device_env():
    """B2A_DEVICES values to exercise: members sharing device 0 on a one-GPU box (the dealing, merge and series ordering are the
    same code; only the NCCL call is replaced by a local sum), plus the real devices where there are several."""
    from gmx2020postanalysis_b200 import lib
    n = lib.load().b2a_device_count()
    return ["0,0,0"] + (["all"] if n >= 2 else [])


@pytest.mark.parametrize("devices", _device_env())
def test_fused_ports_on_several_devices(devices, tmp_path, monkeypatch):
    """SURVEY 8e behind the C++ handle: the ported programs, unchanged, with B2A_DEVICES naming several devices -- batches dealt
    round-robin, histograms merged by b2a_finalize inside b2aRead, per-frame series gathered in frame order -- write the very
    files of the one-device runs (= the reference's)."""
    monkeypatch.setenv("B2A_DEVICES", devices)
    monkeypatch.setenv("B2A_BATCH", "2")
    for name, exe in (("rdf_water_oo", "calc_rdf_region_fused"), ("rdf_il_slabs", "calc_rdf_region_fused"),
                      ("density_template_c1", "readme_template_fused"), ("orient_water", "calc_orient_mof_fused"),
                      ("pair_dist_il", "calc_pair_dist_bulk_fused")):
        case = GC.program_cases()[name]
        got = refprog.run_program(_dropin(exe), str(tmp_path / name), case["system"](), case["nframes"], case["groups"], case["args"], "out.xvg")
        assert got == _golden(name), name
    # per-frame time series: one line per frame, in trajectory order, whichever device processed the frame
    for name in ("calc_density_time", "calc_radius_distribution_in_pore_time", "calc_pair_dist"):
        ref_exe, new_exe = os.path.join(refprog.REFDIR, name), os.path.join(DROPIN, name + "_fused")
        if not (refprog.have_ref(name) and os.path.exists(new_exe)):
            continue
        st_ref, ref = SW.run(ref_exe, str(tmp_path / ("ref_" + name)), name)
        st_new, new = SW.run(new_exe, str(tmp_path / ("new_" + name)), name)
        assert st_ref == "ok" and st_new == "ok", (name, ref, new)
        assert sorted(new) == sorted(ref) and all(new[fn] == ref[fn] for fn in ref), name
    # an unmodified program (compatibility loaders) keeps working on a multi-device context as well
    case = GC.program_cases()["rdf_water_oo"]
    got = refprog.run_program(_dropin(case["exe"]), str(tmp_path / "compat"), case["system"](), case["nframes"], case["groups"], case["args"], "out.xvg")
    assert got == _golden("rdf_water_oo")
    # ... including the velocity loaders (TRR-like input: the v block of every batch follows its x block to the same device)
    exe = "calc_kineticEnergy"
    if refprog.have_ref(exe) and os.path.exists(os.path.join(DROPIN, exe)):
        import numpy as np
        from workloads import synth
        system, nframes = synth.water_box(900, 3.0, seed=75), 7
        v = np.random.default_rng(76).normal(size=(nframes, system.natoms, 3)).astype(np.float32)
        args = ["-low1", 0.2, "-up1", 2.7, "-low2", 0.1, "-up2", 2.9]
        ref = refprog.run_program(os.path.join(refprog.REFDIR, exe), str(tmp_path / "vref"), system, nframes, ["SOL"], args, "o.xvg", velocities=v)
        got = refprog.run_program(os.path.join(DROPIN, exe), str(tmp_path / "vnew"), system, nframes, ["SOL"], args, "o.xvg", velocities=v)
        assert len(ref) == nframes + 1 and got == ref


def test_density_slit_fused_port_matches_reference_build(tmp_path, monkeypatch):
    """general/density_slit_jhl.cpp ported to K2 (both region tests inclusive, second filter dimension, volume normalisation) against
    the reference build of the original on the same trajectory."""
    from workloads import synth
    exe = "density_slit_jhl"
    if not refprog.have_ref(exe) or not os.path.exists(os.path.join(DROPIN, exe + "_fused")):
        pytest.skip("binaries not built (needs /root/reference at build time)")
    monkeypatch.setenv("B2A_BATCH", "4")
    system = synth.water_box(2500, 4.2, seed=79)
    args = ["-nbin1", 60, "-dim1", 2, "-dim2", 0, "-low1", 0.3, "-up1", 3.9, "-low2", 0.5, "-up2", 3.7, "-yy", 4.2]
    ref = refprog.run_program(os.path.join(refprog.REFDIR, exe), str(tmp_path / "ref"), system, 9, ["SOL"], args, "o.xvg")
    got = refprog.run_program(os.path.join(DROPIN, exe + "_fused"), str(tmp_path / "new"), system, 9, ["SOL"], args, "o.xvg")
    assert len(ref) == 60 and got == ref and any(float(ln.split()[1]) > 0 for ln in ref)


@pytest.mark.parametrize("name", ["calc_density_distribution_in_pore", "calc_areaDensityDistributionPore", "calc_numdens_pore", "calc_pair_dist", "calc_radius_distribution_in_pore", "calc_density_time", "calc_radius_distribution_in_pore_time"])
@pytest.mark.parametrize("batch", ["16", "4"])
def test_region_fused_ports_match_reference_builds(name, batch, tmp_path, monkeypatch):
    """The rest of the density family (z profile inside a cylinder, 2-D map behind three region tests, wrapped box count) ported to
    the general region accumulator: same files as the reference builds of the originals, on the sweep's trajectory and arguments."""
    ref_exe, new_exe = os.path.join(refprog.REFDIR, name), os.path.join(DROPIN, name + "_fused")
    if not (refprog.have_ref(name) and os.path.exists(new_exe)):
        pytest.skip("binaries not built (needs /root/reference at build time)")
    monkeypatch.setenv("B2A_BATCH", batch)
    st_ref, ref = SW.run(ref_exe, str(tmp_path / "ref"), name)
    st_new, new = SW.run(new_exe, str(tmp_path / "new"), name)
    assert st_ref == "ok" and st_new == "ok", (ref, new)
    assert sorted(new) == sorted(ref) and all(new[fn] == ref[fn] for fn in ref)
    body = [ln for fn in ref for ln in ref[fn] if not ln.startswith("#")]
    assert any(any(float(v) != 0.0 for v in ln.split()[(1 if len(ln.split()) > 1 and name != "calc_areaDensityDistributionPore" else 0):]) for ln in body)


def test_live_reference_vs_dropin_other_programs(tmp_path):
    """Programs without a golden file: run both builds here and compare (only where oracle/_ref travelled)."""
    from workloads import synth
    cases = [
        ("density_slit_jhl", synth.water_box(1500, 3.6, seed=71), ["SOL"], 5, []),
        ("calc_pair_dist_bulk", synth.ionic_liquid(150, 3.6, seed=72), ["CAT", "ANI"], 4, []),
        ("template1", synth.water_box(400, 2.4, seed=73), ["SOL"], 3, []),
    ]
    ran = 0
    for exe, system, groups, nframes, args in cases:
        if not refprog.have_ref(exe) or not os.path.exists(os.path.join(DROPIN, exe)):
            continue
        try:
            ref = refprog.run_program(os.path.join(refprog.REFDIR, exe), str(tmp_path / ("r_" + exe)), system, nframes, groups, args, "o.xvg")
        except (RuntimeError, FileNotFoundError):
            continue          # the reference program itself fails on this input (or writes no -o file): nothing to compare
        got = refprog.run_program(os.path.join(DROPIN, exe), str(tmp_path / ("d_" + exe)), system, nframes, groups, args, "o.xvg")
        assert got == ref, exe
        ran += 1
    if ran == 0:
        pytest.skip("no reference binaries on this box")


def test_velocity_program_dropin_matches_reference_build(tmp_path, monkeypatch):
    """phaseChange/calc_kineticEnergy.cpp (TRX_READ_V, loadVelocityCenter + loadPositionCenter): the unmodified source
    built against the drop-in header (velocity centres computed on the device) writes the same file as the reference build."""
    import numpy as np
    from workloads import synth
    exe = "calc_kineticEnergy"
    if not refprog.have_ref(exe) or not os.path.exists(os.path.join(DROPIN, exe)):
        pytest.skip("calc_kineticEnergy binaries not built (needs /root/reference at build time)")
    monkeypatch.setenv("B2A_BATCH", "3")
    system, nframes = synth.water_box(900, 3.0, seed=75), 7
    v = np.random.default_rng(76).normal(size=(nframes, system.natoms, 3)).astype(np.float32)
    args = ["-low1", 0.2, "-up1", 2.7, "-low2", 0.1, "-up2", 2.9]
    ref = refprog.run_program(os.path.join(refprog.REFDIR, exe), str(tmp_path / "ref"), system, nframes, ["SOL"], args, "o.xvg", velocities=v)
    got = refprog.run_program(os.path.join(DROPIN, exe), str(tmp_path / "new"), system, nframes, ["SOL"], args, "o.xvg", velocities=v)
    assert len(ref) == nframes + 1 and got == ref


@pytest.mark.parametrize("batch", ["4", "1"])
def test_dropin_reads_xtc_and_trr(batch, tmp_path, monkeypatch):
    """8f-2: the drop-in program fed by a real XTC file (K frames decoded on K host threads into the pinned batch) and by a
    TRR file gives the same output as when fed the raw-float file holding the same (quantised) coordinates."""
    import ctypes as C
    import subprocess
    import numpy as np
    from workloads import synth, trajio
    import test_trajfmt as TF
    L = C.CDLL(TF.STUB)
    L.gmxstub_write_xtc.argtypes = [C.c_char_p, C.c_int, C.c_int, TF._P, TF._P, C.c_int, TF._P, C.c_float]
    L.gmxstub_write_trr.argtypes = [C.c_char_p, C.c_int, C.c_int, TF._P, TF._P, TF._P, TF._P, C.c_int, TF._P]
    monkeypatch.setenv("B2A_BATCH", batch)
    exe = _dropin("calc_rdf_region")
    system = synth.water_box(700, 2.8, seed=86, split_atoms=True)
    K = 6
    x = system.frames(0, K)
    xq = TF.quantise(x, 1000.0)
    box = np.ascontiguousarray(system.box9(), np.float32).reshape(1, 9)
    outs = {}
    for kind in ("raw", "xtc", "trr"):
        d = tmp_path / kind
        d.mkdir()
        trj = str(d / ("sys." + {"raw": "b2t", "xtc": "xtc", "trr": "trr"}[kind]))
        if kind == "raw":
            trajio.write_trajectory(trj, xq, system.box9())
        elif kind == "xtc":
            TF.write_xtc(L, trj, x, system.box9())
        else:
            assert L.gmxstub_write_trr(trj.encode(), x.shape[1], K, xq.ctypes.data, None, None, box.ctypes.data, 1, None) == 0
        trajio.write_topology(str(d / "sys.b2p"), system)
        trajio.write_index(str(d / "sys.ndx"), {"OW": system.groups["OW"]})
        env = dict(os.environ, LD_LIBRARY_PATH=os.path.join(ROOT, "build") + ":" + os.path.join(ROOT, "gmx2020postanalysis_b200") + ":"
                   + os.environ.get("LD_LIBRARY_PATH", ""))
        r = subprocess.run([exe, "-f", trj, "-s", str(d / "sys.b2p"), "-n", str(d / "sys.ndx"), "-o", str(d / "o.xvg")], input="0 0\n",
                           capture_output=True, text=True, env=env, cwd=str(d))
        assert r.returncode == 0, r.stderr[-800:]
        outs[kind] = refprog.data_lines(str(d / "o.xvg"))
    assert len(outs["raw"]) > 100 and outs["raw"] == outs["xtc"] == outs["trr"]


@pytest.mark.parametrize("center", [0, 2])
def test_handle_full_dropin_matches_reference_build(center, tmp_path, monkeypatch):
    """GmxHandleFull (heterogeneous molecule sizes): the same exerciser source built against the reference's own header
    and against the drop-in header dumps identical centres and unwrapped positions (17 significant digits)."""
    from workloads import synth
    ref_exe, new_exe = os.path.join(refprog.REFDIR, "full_dump"), os.path.join(DROPIN, "full_dump")
    if not (os.path.exists(ref_exe) and os.path.exists(new_exe)):
        pytest.skip("full_dump binaries not built (needs /root/reference at build time)")
    monkeypatch.setenv("B2A_BATCH", "2")
    system = synth.ionic_liquid(60, 3.2, seed=74)
    args = ["-center", center]
    ref = refprog.run_program(ref_exe, str(tmp_path / "ref"), system, 5, ["System", "ANI"], args, "o.dat")
    got = refprog.run_program(new_exe, str(tmp_path / "new"), system, 5, ["System", "ANI"], args, "o.dat")
    assert len(ref) > 600 and got == ref
    # with velocities (TRX_READ_V): GmxHandleFull::loadVelocityCenter, per-molecule totals, on the device
    import numpy as np
    v = np.random.default_rng(77).normal(size=(5, system.natoms, 3)).astype(np.float32)
    ref = refprog.run_program(ref_exe, str(tmp_path / "refv"), system, 5, ["System", "ANI"], args + ["-vel", 1], "o.dat", velocities=v)
    got = refprog.run_program(new_exe, str(tmp_path / "newv"), system, 5, ["System", "ANI"], args + ["-vel", 1], "o.dat", velocities=v)
    assert sum(ln.startswith("v ") for ln in ref) == 5 * (120 + 60) and got == ref


# ---- the whole surface: every live program of the reference, reference build vs drop-in build ------------------------
@pytest.mark.parametrize("name", sorted(SW.PROGRAMS))
def test_every_live_reference_program_as_dropin(name, tmp_path, monkeypatch):
    """SURVEY.md 8(b) "who calls it": all 32 programs under src/general, src/mofs and src/phaseChange that compile against
    the reference's own headers also compile, unmodified, against include/itp/gmx, and write byte-identical files (every
    file the program creates, not just -o) when the loaders run on the GPU.  B2A_BATCH=4 makes the 6-frame run cross a
    pinned-batch boundary."""
    ref_exe, new_exe = os.path.join(refprog.REFDIR, name), os.path.join(DROPIN, name)
    if not (refprog.have_ref(name) and os.path.exists(new_exe)):
        pytest.skip("binaries not built (needs /root/reference at build time)")
    monkeypatch.setenv("B2A_BATCH", "4")
    st_ref, ref = SW.run(ref_exe, str(tmp_path / "ref"), name)
    st_new, new = SW.run(new_exe, str(tmp_path / "new"), name)
    if st_ref != "ok":
        # the reference program itself dies on this input (calc_density_region: fmt::format_error thrown by its own format
        # string, general/calc_density_region.cpp:76; calc_force_distribution: segmentation fault): same fate expected
        assert st_new != "ok", f"{name}: reference build fails (rc={ref}) but the drop-in build succeeds"
        return
    assert st_new == "ok", f"{name}: drop-in build failed rc={new}"
    assert sorted(new) == sorted(ref), (sorted(new), sorted(ref))
    for fn in ref:
        assert new[fn] == ref[fn], f"{name}: {fn} differs"
