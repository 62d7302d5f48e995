"""The oracle (oracle/port/oracle.c) against the REFERENCE: (1) live, where oracle/_ref exists -- the
reference's own loaders and programs compiled unmodified; (2) always, against the committed fixtures under
tests/golden/ that tests/make_golden.py produced from (1)."""
import os

import numpy as np
import pytest

import golden_cases as GC
import refprog
from workloads import synth
from oracle import port, refshim

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
needs_ref = pytest.mark.skipif(not refshim.available(), reason="oracle/_ref not built (no /root/reference on this box)")


# ---- (2) committed fixtures ---------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def gold():
    return np.load(os.path.join(GOLD, "loaders.npz"))


def test_golden_loaders(gold):
    for name, system in GC.loader_systems().items():
        x = system.frame(GC.LOADER_FRAME)
        Lbox = port.lbox_from_box9(system.box9())
        for s, off in zip(system.species, system.offsets()):
            idx = np.arange(off, off + s.nmol * s.napm, dtype=np.int32)
            assert np.array_equal(port.load_position(x, idx, s.napm, Lbox), gold[f"{name}.{s.name}.pos"])
            ws = {0: s.mass.astype(np.float64), 1: np.ones(s.napm), 2: s.charge.astype(np.float64)}
            for com in (0, 1, 2):
                got = port.load_position_center(x, idx, s.napm, ws[com], Lbox)
                assert np.array_equal(got, gold[f"{name}.{s.name}.cen{com}"], equal_nan=True), (name, s.name, com)


def test_golden_edge_fixture(gold):
    x, box, idx, napm, mass, charge = GC.edge_fixture()
    Lbox = port.lbox_from_box9(box)
    pos = port.load_position(x, idx, napm, Lbox)
    assert np.array_equal(pos, gold["edge.pos"])
    # the atom exactly L/2 from its reference atom is NOT shifted (strict >), the one a hair beyond is
    assert pos[3, 1, 0] == np.float32(2.5) and pos[3, 2, 0] < 0.0
    for com, w in ((0, mass), (1, np.ones(3)), (2, charge)):
        assert np.array_equal(port.load_position_center(x, idx, napm, w, Lbox), gold[f"edge.cen{com}"], equal_nan=True)


def test_golden_eigen_sum_and_periodicity(gold):
    w = np.array([12.011, 1.008, 1.008, 14.007, 12.011, 15.9994, 1.008, 0.3, 7.7, 1e-3, 5.5], dtype=np.float64)
    assert [port.eigen_sum(w[:n]) for n in range(1, len(w) + 1)] == list(gold["eigen_sum"])
    assert [port.periodicity(a, b) for a, b in GC.periodicity_inputs()] == list(gold["periodicity"])


def test_golden_handle_full(gold):
    """GmxHandleFull loaders (heterogeneous molecule sizes) against the reference's own outputs."""
    system, idx, napm, m, q = GC.full_fixture()
    x = system.frame(GC.LOADER_FRAME)
    Lbox = port.lbox_from_box9(system.box9())
    assert np.array_equal(port.full_load_position(x, idx, napm, Lbox), gold["full.pos"])
    assert np.array_equal(port.full_load_position_center(x, idx, napm, m, Lbox), gold["full.cen0"], equal_nan=True)
    assert np.array_equal(port.full_load_position_center(x, idx, napm, q, Lbox), gold["full.cen2"], equal_nan=True)


def test_golden_velocity_force_loaders(gold):
    """loadVelocity[Center] / loadForce[Center] (no unwrapping) against the reference's own outputs."""
    system, v, f = GC.field_fixture()
    for s, off in zip(system.species, system.offsets()):
        idx = np.arange(off, off + s.nmol * s.napm, dtype=np.int32)
        ws = {0: s.mass.astype(np.float64), 1: np.ones(s.napm), 2: s.charge.astype(np.float64)}
        for arr, tag in ((v, "vel"), (f, "force")):
            assert np.array_equal(port.field_values(arr, idx, s.napm), gold[f"field.{s.name}.{tag}"])
            for com in (0, 1, 2):
                assert np.array_equal(port.field_center(arr, idx, s.napm, ws[com]), gold[f"field.{s.name}.{tag}.cen{com}"],
                                      equal_nan=True), (s.name, tag, com)


def _golden_lines(name):
    with open(os.path.join(GOLD, name + ".txt")) as fp:
        return fp.read().split("\n")[:-1]


def _oracle_lines(name, case):
    system = case["system"]()
    a = case["args"]
    if case["exe"] == "calc_rdf_region":
        kw = {}
        m = {"-nbin": "nbin", "-low": "low", "-up": "up", "-dim": "dim", "-dim2": "dim2", "-low2": "low2", "-up2": "up2", "-Rm": "Rm"}
        for k in range(0, len(a), 2):
            kw[m[a[k]]] = a[k + 1]
        return refprog.oracle_rdf_region(system, case["nframes"], case["groups"][0], case["groups"][1], **kw)
    if case["exe"] == "readme_template":
        kw = {a[k][1:]: a[k + 1] for k in range(0, len(a), 2)}
        return refprog.oracle_readme_template(system, case["nframes"], case["groups"][0], **kw)[0]
    if case["exe"] == "calc_orient_mof":
        return refprog.oracle_orient_mof(system, case["nframes"], case["groups"][0], **refprog.parse_orient_args(a))
    if case["exe"] == "calc_pair_dist_bulk":
        kw = {a[k][1:]: a[k + 1] for k in range(0, len(a), 2)}
        return refprog.oracle_pair_dist_bulk(system, case["nframes"], case["groups"][0], case["groups"][1], **kw)
    raise AssertionError(case["exe"])


@pytest.mark.parametrize("name", sorted(GC.program_cases()))
def test_golden_program_output(name):
    """Text-identical to what the reference's own program wrote for the same synthetic trajectory."""
    assert _oracle_lines(name, GC.program_cases()[name]) == _golden_lines(name)


# ---- (1) live against oracle/_ref ---------------------------------------------------------------------------
@needs_ref
def test_live_loaders_random():
    rng = np.random.default_rng(3)
    for n in range(1, 30):
        w = rng.normal(size=n) * 10.0 ** rng.integers(-3, 3, size=n)
        assert port.eigen_sum(w) == refshim.eigen_sum(w)
    for system in (synth.water_box(400, 2.3, seed=31), synth.ionic_liquid(80, 3.1, seed=32)):
        for f in range(2):
            x, box = system.frame(f), system.box9()
            Lbox = port.lbox_from_box9(box)
            for s, off in zip(system.species, system.offsets()):
                idx = np.arange(off, off + s.nmol * s.napm, dtype=np.int32)
                rh = refshim.RefHandle(idx, s.napm, s.mass, s.charge)
                rh.set_frame(x, box)
                assert np.array_equal(rh.load_position(), port.load_position(x, idx, s.napm, Lbox))
                ws = {0: s.mass.astype(np.float64), 1: np.ones(s.napm), 2: s.charge.astype(np.float64)}
                for com in (0, 1, 2):
                    assert np.array_equal(rh.load_position_center(com),
                                          port.load_position_center(x, idx, s.napm, ws[com], Lbox), equal_nan=True)


@needs_ref
@pytest.mark.parametrize("name", ["rdf_il_slabs", "orient_water", "pair_dist_il"])
def test_live_program_output(name, tmp_path):
    case = GC.program_cases()[name]
    if not refprog.have_ref(case["exe"]):
        pytest.skip("reference program not built")
    ref = refprog.run_program(os.path.join(refprog.REFDIR, case["exe"]), str(tmp_path), case["system"](), case["nframes"],
                              case["groups"], case["args"], "out.xvg")
    assert ref == _oracle_lines(name, case) == _golden_lines(name)
