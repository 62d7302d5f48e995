"""GPU parity tests proper: every call goes through the C ABI (ctypes -> libb2a.so) and is compared with the
oracle on the same seeded inputs.  Integer histograms and FP64 centres/positions are required to be BIT-EXACT
(the north star's 1e-5 relative tolerance on centres is met with margin 0)."""
import os

import numpy as np
import pytest

import golden_cases as GC
import refprog
from gmx2020postanalysis_b200 import lib
from workloads import synth
from oracle import port

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _groups(ctx, system, names):
    infos = []
    for g, name in enumerate(names):
        gi = refprog.group_info(system, name)
        ctx.set_group(g, gi["index"], gi["napm"], gi["mass"], gi["charge"])
        infos.append(gi)
    return infos


# ---- A3 / A4: loaders ---------------------------------------------------------------------------------------
@pytest.mark.parametrize("which", ["water", "il"])
def test_loaders_bit_exact(gpu_ctx_factory, which):
    system = synth.water_box(2500, 4.2, seed=51) if which == "water" else synth.ionic_liquid(300, 4.6, seed=52)
    names = ["SOL"] if which == "water" else ["CAT", "ANI"]
    K = 3
    ctx = gpu_ctx_factory(system.natoms, K)
    infos = _groups(ctx, system, names)
    x, box = system.frames(0, K), system.box9()
    Lbox = port.lbox_from_box9(box)
    ctx.submit(x, box)
    for g, gi in enumerate(infos):
        for com in (0, 1, 2):
            batch = ctx.centers_batch(g, com)
            for f in range(K):
                exp = port.load_position_center(x[f], gi["index"], gi["napm"], gi["w"][com], Lbox)
                assert np.array_equal(ctx.centers(g, com, f), exp, equal_nan=True)
                assert np.array_equal(batch[f], exp, equal_nan=True)
        exp = port.load_position_center(x[1], gi["index"], gi["napm"], gi["charge"], Lbox, dipole_guard=True)
        assert np.array_equal(ctx.centers(g, lib.COM_DIPOLE, 1), exp, equal_nan=True)
        for f in (0, K - 1):
            assert np.array_equal(ctx.positions(g, gi["napm"], f), port.load_position(x[f], gi["index"], gi["napm"], Lbox))


def test_loaders_golden_from_reference(gpu_ctx_factory):
    """Against outputs of the reference's own loadPosition / loadPositionCenter (tests/golden/loaders.npz)."""
    gold = np.load(os.path.join(GOLD, "loaders.npz"))
    for name, system in GC.loader_systems().items():
        ctx = gpu_ctx_factory(system.natoms, 1)
        ctx.submit(system.frame(GC.LOADER_FRAME), system.box9())
        for g, (s, off) in enumerate(zip(system.species, system.offsets())):
            idx = np.arange(off, off + s.nmol * s.napm, dtype=np.int32)
            ctx.set_group(g, idx, s.napm, s.mass.astype(np.float64), s.charge.astype(np.float64))
            assert np.array_equal(ctx.positions(g, s.napm, 0), gold[f"{name}.{s.name}.pos"])
            for com in (0, 1, 2):
                assert np.array_equal(ctx.centers(g, com, 0), gold[f"{name}.{s.name}.cen{com}"], equal_nan=True)
    x, box, idx, napm, mass, charge = GC.edge_fixture()
    ctx = gpu_ctx_factory(len(x), 1)
    ctx.set_group(0, idx, napm, mass, charge)
    ctx.submit(x, box)
    assert np.array_equal(ctx.positions(0, napm, 0), gold["edge.pos"])
    for com in (0, 1, 2):
        assert np.array_equal(ctx.centers(0, com, 0), gold[f"edge.cen{com}"], equal_nan=True)


def test_noncontiguous_index_group(gpu_ctx_factory):
    """O-only group of a 3-site water box (stride-3 gather, napm = 1): centre == the float coordinate exactly."""
    system = synth.water_box(1000, 3.1, seed=53, split_atoms=True)
    ctx = gpu_ctx_factory(system.natoms, 1)
    (gi,) = _groups(ctx, system, ["OW"])
    x = system.frame(0)
    ctx.submit(x, system.box9())
    assert np.array_equal(ctx.centers(0, 0, 0), x[0::3].astype(np.float64))


def test_error_behaviour(gpu_ctx_factory):
    ctx = gpu_ctx_factory(30, 2)
    with pytest.raises(lib.B2AError, match="whole molecules"):            # gmx_handle.ipp:432
        ctx.set_group(0, np.arange(10), 3, np.ones(3), np.ones(3))
    ctx.set_group(0, np.arange(30), 3, np.ones(3), np.ones(3))
    with pytest.raises(lib.B2AError, match="no frame batch"):
        ctx.centers(0, 0, 0)
    with pytest.raises(lib.B2AError, match=r"grp\(1\) should less than ngrps"):   # gmx_handle.ipp:101
        ctx.centers(1, 0, 0)
    ctx.submit(np.zeros((1, 30, 3), np.float32) + 0.5, np.eye(3, dtype=np.float32).ravel() * 2)
    with pytest.raises(lib.B2AError, match="frame 1 outside"):
        ctx.centers(0, 0, 1)


# ---- A8: density ---------------------------------------------------------------------------------------------
def _density_oracle(x, gi, Lbox, K, com, dim, low, up, dbin, nbin, inclusive, dim2, low2, up2):
    exp, drop = np.zeros(nbin, np.uint64), 0
    for f in range(K):
        pc = port.load_position_center(x[f], gi["index"], gi["napm"], gi["w"][com], Lbox)
        _, d = port.density_accum(pc, dim, refprog.f32(low), refprog.f32(up), dbin, nbin, exp, bool(inclusive), dim2,
                                  refprog.f32(low2), refprog.f32(up2))
        drop += d
    return exp, drop


@pytest.mark.parametrize("case", [
    dict(sys="water", nbin=100, dim=2, low=0.0, up=4.476, inclusive=0, dim2=-1, low2=0.0, up2=0.0, com=0),     # BASELINE config 1
    dict(sys="il", nbin=37, dim=0, low=0.3, up=4.9, inclusive=1, dim2=1, low2=0.5, up2=4.0, com=1),
    dict(sys="water", nbin=7, dim=1, low=1.0, up=2.0, inclusive=0, dim2=-1, low2=0.0, up2=0.0, com=0),
    dict(sys="water", nbin=50000, dim=0, low=0.0, up=4.476, inclusive=0, dim2=-1, low2=0.0, up2=0.0, com=1),   # global-atomic path
])
def test_density_bit_exact(gpu_ctx_factory, case):
    system = synth.water_box(3000, 4.476, seed=20200101) if case["sys"] == "water" else synth.ionic_liquid(400, 5.1, seed=54)
    name = "SOL" if case["sys"] == "water" else "CAT"
    K = 5
    ctx = gpu_ctx_factory(system.natoms, K)
    (gi,) = _groups(ctx, system, [name])
    x, box = system.frames(0, K), system.box9()
    Lbox = port.lbox_from_box9(box)
    dbin = float(np.float32(np.float32(np.float32(case["up"]) - np.float32(case["low"])) / np.float32(case["nbin"])))
    acc = ctx.density_create(grp=0, com=case["com"], dim=case["dim"], low=case["low"], up=case["up"], dbin=dbin,
                             nbin=case["nbin"], upper_inclusive=case["inclusive"], dim2=case["dim2"], low2=case["low2"],
                             up2=case["up2"])
    ctx.submit(x, box)
    ctx.accumulate(acc)
    ctx.submit(x[:2], box)          # second, shorter batch accumulates on top
    ctx.accumulate(acc)
    got, st = ctx.read(acc, case["nbin"])
    e1, d1 = _density_oracle(x, gi, Lbox, K, case["com"], case["dim"], case["low"], case["up"], dbin, case["nbin"],
                             case["inclusive"], case["dim2"], case["low2"], case["up2"])
    e2, d2 = _density_oracle(x[:2], gi, Lbox, 2, case["com"], case["dim"], case["low"], case["up"], dbin, case["nbin"],
                             case["inclusive"], case["dim2"], case["low2"], case["up2"])
    assert np.array_equal(got, e1 + e2)
    assert int(st[lib.STAT_FRAMES]) == K + 2 and int(st[lib.STAT_DROPPED]) == d1 + d2
    assert got.sum() > 0


def test_density_template_matches_reference_program_output(gpu_ctx_factory):
    """BASELINE config 1 (reduced to 10 frames): output text of the README template program, from GPU counts."""
    case = GC.program_cases()["density_template_c1"]
    system, K, nbin = case["system"](), case["nframes"], 100
    ctx = gpu_ctx_factory(system.natoms, K)
    _groups(ctx, system, ["SOL"])
    low, up = np.float32(0.0), np.float32(4.476)
    dbin = float(np.float32(np.float32(up - low) / np.float32(nbin)))
    acc = ctx.density_create(grp=0, com=0, dim=2, low=0.0, up=4.476, dbin=dbin, nbin=nbin, upper_inclusive=0, dim2=-1,
                             low2=0.0, up2=0.0)
    ctx.submit(system.frames(0, K), system.box9())
    ctx.accumulate(acc)
    got, st = ctx.read(acc, nbin)
    dens = got.astype(np.float64) / int(st[lib.STAT_FRAMES])     # density /= hd.nframe (README.md:154)
    lines = [refprog.fmt_fixed(float(low) + dbin / 2 + dbin * i, 8, 5) + " " + refprog.fmt_fixed(dens[i], 10, 5) for i in range(nbin)]
    assert lines == open(os.path.join(GOLD, "density_template_c1.txt")).read().split("\n")[:-1]


# ---- A7: RDF -------------------------------------------------------------------------------------------------
def _rdf_gpu(ctx, system, g0, g1, K, spec, Rm, Rbin, paranoid=0, frames=None, symmetric=0):
    _groups(ctx, system, [g0, g1])
    acc = ctx.rdf_create(grp0=0, grp1=1, com0=spec["com"], com1=spec["com"], nbin=spec["nbin"], low=spec["low"], up=spec["up"],
                         dim=spec["dim"], dim2=spec["dim2"], low2=spec["low2"], up2=spec["up2"], Rm=Rm, Rbin=Rbin)
    ctx.rdf_tune(acc, symmetric, paranoid)
    ctx.submit(system.frames(0, K) if frames is None else frames, system.box9())
    ctx.accumulate(acc)
    got, st = ctx.read(acc, spec["nbin"] * Rbin)
    return got.reshape(Rbin, spec["nbin"]), st


_DEF = dict(nbin=5, low=-100.0, up=100.0, dim=0, dim2=2, low2=-100.0, up2=100.0, Rm=0.0, com=0)


@pytest.mark.parametrize("case", [
    dict(sys=("water", 4000, 4.93), g=("OW", "OW"), K=2),                                                   # C2 shape, reduced
    dict(sys=("il", 400, 5.1), g=("CAT", "ANI"), K=3, nbin=3, low=0.5, up=4.6, dim=1, dim2=0, low2=0.3, up2=4.9, Rm=1.9),
    dict(sys=("il", 400, 5.1), g=("CAT", "ANI"), K=2),                                                      # C3 shape, reduced
    dict(sys=("il", 400, 5.1), g=("ANI", "CAT"), K=2, com=1, nbin=1, low=0.0, up=5.1),
    dict(sys=("water", 3000, 4.476), g=("SOL", "SOL"), K=2, com=2, nbin=2, low=0.0, up=4.476, dim=2, Rm=0.9),   # charge centres: NaN/inf centres
    dict(sys=("water", 1500, 3.55), g=("OW", "OW"), K=1, paranoid=1),                                        # every pair through FP64
    dict(sys=("water", 700, 2.8), g=("OW", "OW"), K=2, Rm=0.4),                                              # tiny Rm: clamp-free, large kmax
    # symmetric evaluation (i < j only, weight 2: what bench.py and the ported calc_rdf_region run) straight against the oracle
    dict(sys=("water", 4000, 4.93), g=("OW", "OW"), K=2, symmetric=1, sym_runs=True),                        # C2 shape: MODE 1, several i-tiles
    dict(sys=("water", 1100, 3.2), g=("OW", "OW"), K=3, symmetric=1, sym_runs=True),                         # one full + one ragged i-tile
    dict(sys=("water", 1500, 3.55), g=("OW", "OW"), K=1, symmetric=1, paranoid=1),                           # every evaluated pair through FP64
    dict(sys=("water", 3000, 4.476), g=("SOL", "SOL"), K=2, symmetric=1, sym_runs=True, com=0, nbin=1, Rm=0.9),   # molecular centres, cell-binned + symmetric
    dict(sys=("water", 3000, 4.476), g=("SOL", "SOL"), K=2, symmetric=1, com=0, nbin=3, low=0.0, up=4.476, dim=2),   # several slabs: ineligible, falls through to the full kernel
    dict(sys=("il", 400, 5.1), g=("CAT", "ANI"), K=2, symmetric=1),                                          # different groups: request ignored
])
def test_rdf_bit_exact(gpu_ctx_factory, case):
    kind, n, L = case["sys"]
    system = synth.water_box(n, L, seed=61, split_atoms=(case["g"][0] == "OW")) if kind == "water" else synth.ionic_liquid(n, L, seed=62)
    spec = dict(_DEF)
    spec.update({k: v for k, v in case.items() if k in _DEF})
    K = case["K"]
    exp, est, Rm, Rbin = refprog.oracle_rdf_region(system, K, case["g"][0], case["g"][1], return_counts=True, **spec)
    ctx = gpu_ctx_factory(system.natoms, K)
    got, st = _rdf_gpu(ctx, system, case["g"][0], case["g"][1], K, spec, Rm, Rbin, case.get("paranoid", 0), symmetric=case.get("symmetric", 0))
    assert np.array_equal(got, exp)
    assert int(got.sum()) == int(est[0]) and int(st[lib.STAT_DROPPED]) == int(est[1])
    assert int(st[lib.STAT_SELECTED]) == int(est[2]) and int(st[lib.STAT_FRAMES]) == K
    if case.get("paranoid") and not case.get("symmetric"):
        assert int(st[lib.STAT_EXACT]) == n * n * K
    if case.get("sym_runs"):
        # the symmetric path really ran: fewer than the N0 * N1 ordered pairs were evaluated
        assert int(st[lib.STAT_PAIRS]) < n * n * K


def test_rdf_matches_reference_program_output(gpu_ctx_factory):
    """The .xvg the reference's calc_rdf_region wrote (tests/golden), reproduced from GPU counts + b2a_rdf_normalise."""
    for name in ("rdf_water_oo", "rdf_il_slabs"):
        case = GC.program_cases()[name]
        system, K = case["system"](), case["nframes"]
        a = case["args"]
        spec = dict(_DEF)
        m = {"-nbin": "nbin", "-low": "low", "-up": "up", "-dim": "dim", "-dim2": "dim2", "-low2": "low2", "-up2": "up2", "-Rm": "Rm"}
        for k in range(0, len(a), 2):
            spec[m[a[k]]] = a[k + 1]
        Lbox = port.lbox_from_box9(system.box9())
        Rm = refprog.f32(spec["Rm"]) if spec["Rm"] >= 1e-6 else port.rdf_default_rm(Lbox)
        Rbin = port.rdf_rbin(Rm)
        ctx = gpu_ctx_factory(system.natoms, K)
        got, _ = _rdf_gpu(ctx, system, case["groups"][0], case["groups"][1], K, spec, Rm, Rbin)
        g = lib.rdf_normalise(got.ravel(), spec["nbin"], Rbin)
        lines = refprog.rdf_region_text(g, spec["nbin"], Rbin, spec["dim"], spec["low"], spec["up"])
        assert lines == open(os.path.join(GOLD, name + ".txt")).read().split("\n")[:-1]


def test_rdf_edge_fixtures(gpu_ctx_factory):
    """Hand-placed pairs: exactly on a bin edge, at r == Rm, at r^2 == 1e-6 boundary, coincident, across the box."""
    L = np.float32(8.0)
    Rm = refprog.f32(3.0)
    Rbin = port.rdf_rbin(Rm)
    base = np.array([1.0, 1.0, 1.0], np.float32)
    offs = [0.0, 0.002, 0.0020000001, 0.001, 0.0009999, 0.0010001, 3.0, 2.9999998, 3.0000002, 1.5, 0.75, 2.998, 4.0, 7.999]
    pts = [base] + [base + np.array([d, 0, 0], np.float32) for d in offs[1:]] + [np.array([7.9995, 1.0, 1.0], np.float32)]
    # a few diagonal pairs whose distance is irrational
    pts += [base + np.array([0.3, 0.4, 1.2], np.float32), base + np.array([1.7320508, 1.7320508, 1.7320508], np.float32)]
    x = np.array(pts, np.float32) % L
    n = len(x)
    box = np.zeros(9, np.float32)
    box[0] = box[4] = box[8] = L
    Lbox = port.lbox_from_box9(box)
    pc = x.astype(np.float64)
    exp, est = port.rdf_accum(pc, pc, Lbox, 1, -100.0, 100.0, 0, 2, -100.0, 100.0, Rm, Rbin)
    ctx = gpu_ctx_factory(n, 1)
    for g in (0, 1):
        ctx.set_group(g, np.arange(n), 1, np.ones(1), np.ones(1))
    ctx.submit(x, box)
    for symmetric in (0, 1):                                  # the full evaluation and the i < j one bench.py runs
        acc = ctx.rdf_create(grp0=0, grp1=1, com0=1, com1=1, nbin=1, low=-100.0, up=100.0, dim=0, dim2=2, low2=-100.0, up2=100.0,
                             Rm=Rm, Rbin=Rbin)
        ctx.rdf_tune(acc, symmetric, 0)
        ctx.accumulate(acc)
        got, st = ctx.read(acc, Rbin)
        assert np.array_equal(got.reshape(Rbin, 1), exp), symmetric
        assert int(st[lib.STAT_DROPPED]) == int(est[1])


def test_rdf_properties_translation_and_transpose(gpu_ctx_factory):
    """Size-independent properties: translating every atom by a box vector leaves the histogram
    unchanged; swapping the groups leaves the radial histogram unchanged when the filter passes everything."""
    system = synth.ionic_liquid(500, 5.4, seed=63)
    K = 2
    Lbox = port.lbox_from_box9(system.box9())
    Rm = port.rdf_default_rm(Lbox)
    Rbin = port.rdf_rbin(Rm)
    spec = dict(_DEF, nbin=1)
    ctx = gpu_ctx_factory(system.natoms, K)
    h_ab, _ = _rdf_gpu(ctx, system, "CAT", "ANI", K, spec, Rm, Rbin)
    ctx2 = gpu_ctx_factory(system.natoms, K)
    h_ba, _ = _rdf_gpu(ctx2, system, "ANI", "CAT", K, spec, Rm, Rbin)
    assert np.array_equal(h_ab, h_ba) and h_ab.sum() > 0
    # Translation by box vectors: every atom moved by its own integer multiples of the box lengths.  Minimum-image
    # distances do not change; with exactly representable shifts (L = 5.5 = 11 * 2^-1, coordinates on a 2^-10 grid) neither do
    # the float coordinates' fractional positions, so every histogram must be IDENTICAL, bin by bin, in both evaluation modes.
    L = 5.5
    system.L = np.array([L, L, L], np.float32)
    x = system.frames(0, K)
    x = (np.round(x * 1024.0) / 1024.0).astype(np.float32)               # 2^-10 grid: x + n L is exact in float for |n| <= 3
    rng = np.random.default_rng(7)
    shift = rng.integers(-3, 4, size=(1, system.natoms, 3)).astype(np.float32) * np.float32(L)
    shift[0, system.groups["CAT"][::25]] = 0.0                           # keep some molecules' reference atoms in place
    xs = (x + shift).astype(np.float32)
    assert np.array_equal((xs - shift).astype(np.float32), x)
    Lbox = port.lbox_from_box9(system.box9())
    Rm = port.rdf_default_rm(Lbox)
    Rbin = port.rdf_rbin(Rm)
    out = []
    for frames in (x, xs):
        c = gpu_ctx_factory(system.natoms, K)
        h, _ = _rdf_gpu(c, system, "CAT", "ANI", K, spec, Rm, Rbin, frames=frames)
        c2 = gpu_ctx_factory(system.natoms, K)
        hs, _ = _rdf_gpu(c2, system, "ANI", "ANI", K, spec, Rm, Rbin, frames=frames, symmetric=1)
        out.append((h, hs))
    assert np.array_equal(out[0][0], out[1][0]) and np.array_equal(out[0][1], out[1][1]) and out[0][1].sum() > 0
    # and the oracle agrees on the translated frames
    g0, g1 = refprog.group_info(system, "CAT"), refprog.group_info(system, "ANI")
    exp = np.zeros((Rbin, 1), dtype=np.uint64)
    for f in range(K):
        p1 = port.load_position_center(xs[f], g0["index"], g0["napm"], g0["w"][0], Lbox)
        p2 = port.load_position_center(xs[f], g1["index"], g1["napm"], g1["w"][0], Lbox)
        port.rdf_accum(p1, p2, Lbox, 1, -100.0, 100.0, 0, 2, -100.0, 100.0, Rm, Rbin, exp)
    assert np.array_equal(out[1][0], exp)


@pytest.mark.parametrize("nonzero_box", [(3.0, 4.0, 5.0)])
def test_rdf_orthorhombic_box(gpu_ctx_factory, nonzero_box):
    """Non-cubic box: per-dimension fixed-point units (the CUBIC=false kernel)."""
    system = synth.ionic_liquid(200, 3.0, seed=64)
    system.L = np.array(nonzero_box, np.float32)
    K = 2
    spec = dict(_DEF, nbin=2, low=0.0, up=3.0)
    exp, est, Rm, Rbin = refprog.oracle_rdf_region(system, K, "CAT", "ANI", return_counts=True, **spec)
    ctx = gpu_ctx_factory(system.natoms, K)
    got, st = _rdf_gpu(ctx, system, "CAT", "ANI", K, spec, Rm, Rbin)
    assert np.array_equal(got, exp) and got.sum() > 0


def test_rdf_full_size_paranoid_free_properties(gpu_ctx_factory):
    """BASELINE config 2 at FULL size (72 000 O atoms, one frame): properties that need no CPU oracle --
    total count == number of ordered pairs within the cut-off as counted by the FP64-only kernel variant on an
    i-subset, histogram symmetric under group swap, frames counter."""
    n, L = 72000, 12.913
    system = synth.water_box(n, L, seed=20200102, split_atoms=True)
    Lbox = port.lbox_from_box9(system.box9())
    Rm = port.rdf_default_rm(Lbox)
    Rbin = port.rdf_rbin(Rm)
    ctx = gpu_ctx_factory(system.natoms, 1)
    x = system.frames(0, 1)
    got, st = _rdf_gpu(ctx, system, "OW", "OW", 1, dict(_DEF), Rm, Rbin, frames=x)
    assert int(st[lib.STAT_PAIRS]) == n * n and int(st[lib.STAT_SELECTED]) == n
    assert got[:, [0, 1, 3, 4]].sum() == 0                      # default slabs: everything lands in meZ == 2
    frac = got.sum() / float(n) / float(n)
    assert abs(frac - np.pi / 6) < 2e-3                         # ideal-gas-like box: sphere of radius L/2 in L^3
    # the oracle on an i-subset of the same frame (first 64 oxygens against all): bit-exact
    sub = x[0][0::3][:64].astype(np.float64)
    allo = x[0][0::3].astype(np.float64)
    exp, _ = port.rdf_accum(sub, allo, Lbox, 5, -100.0, 100.0, 0, 2, -100.0, 100.0, Rm, Rbin)
    ctx2 = gpu_ctx_factory(system.natoms, 1)
    gi = refprog.group_info(system, "OW")
    ctx2.set_group(0, gi["index"][:64], 1, gi["mass"], gi["charge"])
    ctx2.set_group(1, gi["index"], 1, gi["mass"], gi["charge"])
    acc = ctx2.rdf_create(grp0=0, grp1=1, com0=0, com1=0, nbin=5, low=-100.0, up=100.0, dim=0, dim2=2, low2=-100.0,
                          up2=100.0, Rm=Rm, Rbin=Rbin)
    ctx2.submit(x, system.box9())
    ctx2.accumulate(acc)
    got2, _ = ctx2.read(acc, 5 * Rbin)
    assert np.array_equal(got2.reshape(Rbin, 5), exp)


def test_group_redefined_after_accumulator_creation(gpu_ctx_factory):
    """A group may be redefined (b2a_set_group) with MORE molecules after accumulators were created on it: the per-molecule
    scratch sized at create time has to follow (it used to be overrun), symmetric evaluation must notice that the two sides no
    longer hold the same centres (different weights), and counts stay those of the oracle."""
    system = synth.water_box(2600, 4.3, seed=65, split_atoms=True)
    K = 2
    gi = refprog.group_info(system, "OW")
    small = gi["index"][:300]
    ctx = gpu_ctx_factory(system.natoms, K)
    for g in (0, 1):
        ctx.set_group(g, small, 1, gi["mass"], gi["charge"])
    Lbox = port.lbox_from_box9(system.box9())
    Rm = port.rdf_default_rm(Lbox)
    Rbin = port.rdf_rbin(Rm)
    acc = ctx.rdf_create(grp0=0, grp1=1, com0=0, com1=0, nbin=5, low=-100.0, up=100.0, dim=0, dim2=2, low2=-100.0, up2=100.0, Rm=Rm, Rbin=Rbin)
    ctx.rdf_tune(acc, 1, 0)
    coo = ctx.coord_create(grp0=0, grp1=1, com0=0, com1=0, Rmin=0.5, dim=2, low=-100.0, up=100.0, max_count=64)
    x = system.frames(0, K)
    ctx.submit(x, system.box9())
    ctx.accumulate(acc); ctx.accumulate(coo)
    ctx.reset(acc); ctx.reset(coo)
    for g in (0, 1):                                   # now the whole group: 2600 molecules on scratch sized for 300
        ctx.set_group(g, gi["index"], 1, gi["mass"], gi["charge"])
    ctx.submit(x, system.box9())
    ctx.accumulate(acc); ctx.accumulate(coo)
    got, st = ctx.read(acc)
    exp, est, _, _ = refprog.oracle_rdf_region(system, K, "OW", "OW", return_counts=True)
    assert np.array_equal(got.reshape(Rbin, 5), exp) and int(st[lib.STAT_PAIRS]) < 2600 * 2600 * K      # symmetric path, right counts
    cnt = ctx.coord_counts(coo, 1)
    o = x[1][0::3].astype(np.float64)
    ecnt, _ = port.pair_count(o, o, Lbox, refprog.f32(0.5), 2, refprog.f32(-100.0), refprog.f32(100.0))
    assert np.array_equal(cnt, ecnt)
    # same atoms, different weights on one side (geometry centre of a one-atom "molecule" is the same point, but the library cannot
    # know that in general): the symmetric request must be dropped, not mis-applied -- every ordered pair is evaluated
    ctx.set_group(1, gi["index"], 1, gi["mass"] * 2.0, gi["charge"])
    ctx.reset(acc)
    ctx.submit(x, system.box9())
    ctx.accumulate(acc)
    got2, st2 = ctx.read(acc)
    assert np.array_equal(got2.reshape(Rbin, 5), exp) and int(st2[lib.STAT_PAIRS]) == 2600 * 2600 * K


def test_rdf_headline_mode_full_size_against_oracle(gpu_ctx_factory):
    """EXACTLY the configuration bench.py times (BASELINE config 2 at full size, 8 distinct frames per batch,
    b2a_rdf_tune(symmetric=1), cells automatic): (i) the symmetric histogram of the batch equals the full (MODE 0) evaluation of
    the same frames, bin by bin; (ii) the symmetric kernel on one whole frame -- all 72 000 x 72 000 ordered pairs -- is
    bit-identical to oracle.port.rdf_accum (calc_rdf_region.cpp:73-95) of that frame."""
    n, L, K = 72000, 12.913, 8
    system = synth.water_box(n, L, seed=20200102, split_atoms=True)
    Lbox = port.lbox_from_box9(system.box9())
    Rm = port.rdf_default_rm(Lbox)
    Rbin = port.rdf_rbin(Rm)
    x = system.frames(0, K)
    ctx = gpu_ctx_factory(system.natoms, K)
    got_sym, st = _rdf_gpu(ctx, system, "OW", "OW", K, dict(_DEF), Rm, Rbin, frames=x, symmetric=1)
    assert int(st[lib.STAT_FRAMES]) == K and int(st[lib.STAT_SELECTED]) == n * K
    assert int(st[lib.STAT_PAIRS]) < 0.52 * n * n * K            # the i < j evaluation really ran (diagonal tiles in full)
    ctx2 = gpu_ctx_factory(system.natoms, K)
    got_full, st2 = _rdf_gpu(ctx2, system, "OW", "OW", K, dict(_DEF), Rm, Rbin, frames=x, symmetric=0)
    assert int(st2[lib.STAT_PAIRS]) == n * n * K
    assert np.array_equal(got_sym, got_full) and int(st[lib.STAT_DROPPED]) == int(st2[lib.STAT_DROPPED])
    # one whole frame against the oracle (about 20 s of the host's cores)
    f = 3
    o = x[f][0::3].astype(np.float64)
    exp, est = port.rdf_accum(o, o, Lbox, 5, -100.0, 100.0, 0, 2, -100.0, 100.0, Rm, Rbin)
    ctx3 = gpu_ctx_factory(system.natoms, 1)
    got1, st1 = _rdf_gpu(ctx3, system, "OW", "OW", 1, dict(_DEF), Rm, Rbin, frames=x[f:f + 1], symmetric=1)
    assert np.array_equal(got1, exp)
    assert int(got1.sum()) == int(est[0]) and int(st1[lib.STAT_DROPPED]) == int(est[1])


def test_full_size_configs_3_4_5_properties(gpu_ctx_factory):
    """BASELINE configs 3, 4 and 5 at FULL size, one frame each, through properties that need no full-size CPU run:
    C3 (1 M-atom ionic liquid): cation-anion == anion-cation, i-subset bit-exact against the oracle, sphere fraction;
    C4 (499 998-atom water): literal FP64, bracketed and validated orientation histograms identical, every molecule counted once;
    C5 (4 M atoms): cell-binned symmetric water-water == symmetric all-pairs, cross pair cell-binned == all-pairs."""
    # ---- C3
    system = synth.ionic_liquid(31250, 22.1, seed=20200103)
    Lbox = port.lbox_from_box9(system.box9())
    Rm = port.rdf_default_rm(Lbox)
    Rbin = port.rdf_rbin(Rm)
    x = system.frames(0, 1)
    ctx = gpu_ctx_factory(system.natoms, 1)
    h_ab, st = _rdf_gpu(ctx, system, "CAT", "ANI", 1, dict(_DEF, nbin=1), Rm, Rbin, frames=x)
    ctx.close()
    ctx = gpu_ctx_factory(system.natoms, 1)
    h_ba, _ = _rdf_gpu(ctx, system, "ANI", "CAT", 1, dict(_DEF, nbin=1), Rm, Rbin, frames=x)
    ctx.close()
    assert np.array_equal(h_ab, h_ba) and int(st[lib.STAT_PAIRS]) == 31250 * 31250
    assert abs(h_ab.sum() / 31250.0 ** 2 - np.pi / 6) < 3e-3
    gc, ga = refprog.group_info(system, "CAT"), refprog.group_info(system, "ANI")
    pc = port.load_position_center(x[0], gc["index"][:48 * 25], 25, gc["w"][0], Lbox)
    pa = port.load_position_center(x[0], ga["index"], 7, ga["w"][0], Lbox)
    exp, _ = port.rdf_accum(pc, pa, Lbox, 1, -100.0, 100.0, 0, 2, -100.0, 100.0, Rm, Rbin)
    ctx = gpu_ctx_factory(system.natoms, 1)
    ctx.set_group(0, gc["index"][:48 * 25], 25, gc["mass"], gc["charge"])
    ctx.set_group(1, ga["index"], 7, ga["mass"], ga["charge"])
    acc = ctx.rdf_create(grp0=0, grp1=1, com0=0, com1=0, nbin=1, low=-100.0, up=100.0, dim=0, dim2=2, low2=-100.0, up2=100.0, Rm=Rm, Rbin=Rbin)
    ctx.submit(x, system.box9()); ctx.accumulate(acc)
    assert np.array_equal(ctx.read(acc, Rbin)[0].reshape(Rbin, 1), exp)
    ctx.close()
    # ---- C4
    system = synth.water_box(166666, 17.09, seed=20200104)
    x = system.frames(0, 1)
    ctx = gpu_ctx_factory(system.natoms, 1)
    _groups(ctx, system, ["SOL"])
    ctx.submit(x, system.box9())
    res = []
    for mode in (0, 1, 2):
        acc = ctx.orient_create(grp=0, com=0, DIMN=2, low=-1.0, up=100.0, c1=0.0, c2=0.0, dr=100.0, CR=100.0, Cr=0.0, nAngle=90,
                                tp1=1, tp2=2, tp3=3, DIMNOrient=2, method=1)
        ctx.acc_fast(acc, mode); ctx.accumulate(acc)
        res.append(ctx.read(acc, 91))
    (c0, s0), (c1, s1), (c2, s2) = res
    assert np.array_equal(c0, c1) and np.array_equal(c0, c2) and int(s2[lib.STAT_MISPROVEN]) == 0
    assert int(c0[90]) == 166666 and int(c0[:90].sum()) + int(s0[lib.STAT_DROPPED]) == 166666
    mu = ctx.centers_batch(0, lib.COM_DIPOLE)[0]
    assert np.allclose(np.linalg.norm(mu, axis=1), 0.04894, atol=2e-4)          # rigid SPC/E dipole, every molecule
    ctx.close()
    # ---- C5
    system = synth.mixed_box()
    x = system.frames(0, 1)
    Rm5 = refprog.f32(3.0)
    Rbin5 = port.rdf_rbin(Rm5)
    out = {}
    for tag, g, cells, sym in (("ww_cells", ("SOL", "SOL"), 1, 1), ("ww_all", ("SOL", "SOL"), 0, 1), ("wc_cells", ("SOL", "COS"), 1, 0), ("wc_all", ("SOL", "COS"), 0, 0)):
        ctx = gpu_ctx_factory(system.natoms, 1)
        _groups(ctx, system, list(g))
        acc = ctx.rdf_create(grp0=0, grp1=1, com0=0, com1=0, nbin=1, low=-100.0, up=100.0, dim=0, dim2=2, low2=-100.0, up2=100.0, Rm=Rm5, Rbin=Rbin5)
        ctx.rdf_tune(acc, sym, 0); ctx.rdf_cells(acc, cells)
        ctx.submit(x, system.box9()); ctx.accumulate(acc)
        out[tag] = ctx.read(acc, Rbin5)
        ctx.close()
    assert np.array_equal(out["ww_cells"][0], out["ww_all"][0]) and np.array_equal(out["wc_cells"][0], out["wc_all"][0])
    assert out["ww_cells"][0].sum() > 4e9 and int(out["ww_cells"][1][lib.STAT_PAIRS]) * 20 < int(out["ww_all"][1][lib.STAT_PAIRS])


# ---- A6: orientation ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["orient_water", "orient_il_cross"])
def test_orient_bit_exact_and_reference_text(gpu_ctx_factory, name):
    case = GC.program_cases()[name]
    system, K = case["system"](), case["nframes"]
    kw = refprog.parse_orient_args(case["args"])
    exp_o, exp_n = refprog.oracle_orient_mof(system, K, case["groups"][0], return_counts=True, **kw)
    ctx = gpu_ctx_factory(system.natoms, K)
    _groups(ctx, system, case["groups"])
    D = kw["DIMN"]
    acc = ctx.orient_create(grp=0, com=kw["NCM"], DIMN=D, low=kw["low"][D], up=kw["up"][D], c1=kw["c1"], c2=kw["c2"],
                            dr=kw["dr"], CR=kw["CR"], Cr=kw["Cr"], nAngle=kw["nA"], tp1=kw["tp1"], tp2=kw["tp2"],
                            tp3=kw["tp3"], DIMNOrient=kw["DIMNOrient"], method=kw["method"])
    ctx.submit(system.frames(0, K), system.box9())
    ctx.accumulate(acc)
    nbin1 = len(exp_n)
    got, st = ctx.read(acc, nbin1 * kw["nA"] + nbin1)
    o, ncm = got[:nbin1 * kw["nA"]].reshape(kw["nA"], nbin1), got[nbin1 * kw["nA"]:]
    assert np.array_equal(o, exp_o) and np.array_equal(ncm, exp_n) and o.sum() > 0
    assert refprog.orient_text(o, ncm, kw["nA"]) == open(os.path.join(GOLD, name + ".txt")).read().split("\n")[:-1]


@pytest.mark.parametrize("which", ["water", "il"])
def test_orient_fused_centre_output(gpu_ctx_factory, which):
    """b2a_acc_also_centers: the orientation pass also materialises the centres of a second kind (config 4: the dipole vectors).
    The fused centres are bit-identical to b2a_centers of the same kind computed on their own, the histogram is unchanged, and the
    cache is honoured (no second kernel, fresh values after the next batch)."""
    if which == "water":
        system, grp, tp, com2 = synth.water_box(5000, 5.3, seed=20200104), "SOL", (1, 2, 3), lib.COM_DIPOLE
    else:
        system, grp, tp, com2 = synth.ionic_liquid(300, 4.0, seed=25), "CAT", (3, 7, 12), lib.COM_CHARGE
    K = 5
    kw = dict(grp=0, com=0, DIMN=2, low=0.3, up=3.9, c1=2.0, c2=2.0, dr=0.25, CR=2.0, Cr=0.0, nAngle=90, tp1=tp[0], tp2=tp[1], tp3=tp[2],
              DIMNOrient=2, method=1)
    ref = gpu_ctx_factory(system.natoms, K)
    _groups(ref, system, [grp])
    a0 = ref.orient_create(**kw)
    ctx = gpu_ctx_factory(system.natoms, K)
    _groups(ctx, system, [grp])
    a1 = ctx.orient_create(**kw)
    ctx.acc_also_centers(a1, com2)
    for b in range(2):
        x = system.frames(K * b, K)
        ref.submit(x, system.box9()); ref.accumulate(a0)
        ctx.submit(x, system.box9()); ctx.accumulate(a1)
        ctx.timing_enable(True); ctx.timing_read()
        got = ctx.centers_batch(0, com2)                   # served from what the orientation kernel wrote
        assert ctx.timing_read()["k1_centers"][1] == 0     # ... no centre kernel ran
        ctx.timing_enable(False)
        exp = ref.centers_batch(0, com2)
        assert np.array_equal(got, exp, equal_nan=True) and np.isfinite(got).all()
    h0, s0 = ref.read(a0)
    h1, s1 = ctx.read(a1)
    assert np.array_equal(h0, h1) and h0.sum() > 0 and int(s0[lib.STAT_SELECTED]) == int(s1[lib.STAT_SELECTED])


def test_dipole_orientation_config4_shape(gpu_ctx_factory):
    """BASELINE config 4, reduced: charge-centre dipole vectors (guarded) + bisector angle histogram."""
    system = synth.water_box(5000, 5.3, seed=20200104)
    K = 3
    ctx = gpu_ctx_factory(system.natoms, K)
    (gi,) = _groups(ctx, system, ["SOL"])
    x, box = system.frames(0, K), system.box9()
    Lbox = port.lbox_from_box9(box)
    ctx.submit(x, box)
    mu = ctx.centers_batch(0, lib.COM_DIPOLE)
    for f in range(K):
        assert np.array_equal(mu[f], port.load_position_center(x[f], gi["index"], 3, gi["charge"], Lbox, dipole_guard=True))
    # |mu| of rigid SPC/E is constant: 0.4238 e * 2 * 0.1 nm * cos(109.47/2 deg) = 0.0489 e nm
    assert np.allclose(np.linalg.norm(mu[0], axis=1), 0.04894, atol=2e-4)
    kw = dict(NCM=0, DIMN=2, low=(0, 0, -1.0), up=(100, 100, 100.0), c1=0.0, c2=0.0, dr=100.0, CR=100.0, Cr=0.0, nA=90,
              tp1=1, tp2=2, tp3=3, DIMNOrient=2, method=1)
    exp_o, exp_n = refprog.oracle_orient_mof(system, K, "SOL", return_counts=True, **kw)
    acc = ctx.orient_create(grp=0, com=0, DIMN=2, low=-1.0, up=100.0, c1=0.0, c2=0.0, dr=100.0, CR=100.0, Cr=0.0, nAngle=90,
                            tp1=1, tp2=2, tp3=3, DIMNOrient=2, method=1)
    ctx.accumulate(acc)
    got, st = ctx.read(acc, 90 + 1)
    assert np.array_equal(got[:90].reshape(90, 1), exp_o) and got[90] == exp_n[0] == 5000 * K


# ---- 8f-4: velocity / force loaders and coordination numbers --------------------------------------------------------
def test_velocity_force_loaders(gpu_ctx_factory):
    """loadVelocity[Center] / loadForce[Center] on the device, bit-exact vs the oracle and the reference's own outputs."""
    system, v, f = GC.field_fixture()
    gold = np.load(os.path.join(GOLD, "loaders.npz"))
    K = 2
    ctx = gpu_ctx_factory(system.natoms, K)
    infos = _groups(ctx, system, ["CAT", "ANI"])
    ctx.submit(system.frames(0, K), system.box9())
    with pytest.raises(lib.B2AError):
        ctx.field_centers(lib.FIELD_V, 0)              # nothing submitted for this batch yet
    v2, f2 = np.stack([v, v[::-1]]), np.stack([f, 2 * f])
    ctx.submit_field(lib.FIELD_V, v2)
    ctx.submit_field(lib.FIELD_F, f2)
    for g, (gi, sp) in enumerate(zip(infos, system.species)):
        for field, arr, tag in ((lib.FIELD_V, v2, "vel"), (lib.FIELD_F, f2, "force")):
            for com in (0, 1, 2):
                got = ctx.field_centers_batch(field, g, com)
                for k in range(K):
                    assert np.array_equal(got[k], port.field_center(arr[k], gi["index"], gi["napm"], gi["w"][com]), equal_nan=True)
                assert np.array_equal(got[0], gold[f"field.{sp.name}.{tag}.cen{com}"], equal_nan=True)
                assert np.array_equal(ctx.field_centers(field, g, com, 1), got[1], equal_nan=True)
            assert np.array_equal(ctx.field_values(field, g, gi["napm"], 1), port.field_values(arr[1], gi["index"], gi["napm"]))
            assert np.array_equal(ctx.field_values(field, g, gi["napm"], 0), gold[f"field.{sp.name}.{tag}"])
    ctx.submit(system.frames(0, K), system.box9())     # a new batch invalidates the uploaded fields
    with pytest.raises(lib.B2AError):
        ctx.field_values(lib.FIELD_F, 0, infos[0]["napm"], 0)


def test_coord_matches_reference_program_output(gpu_ctx_factory):
    """mofs/calc_pair_dist_bulk.cpp on an ionic liquid: per-molecule counts vs the oracle, text vs the reference program."""
    case = GC.program_cases()["pair_dist_il"]
    system, K = case["system"](), case["nframes"]
    ctx = gpu_ctx_factory(system.natoms, K)
    _groups(ctx, system, case["groups"])
    exp_counts, exp_numA = refprog.oracle_pair_dist_bulk(system, K, "CAT", "ANI", ncm=0, Rmin=0.8, dim=2, low=0.6, up=3.9,
                                                         return_counts=True)
    acc = ctx.coord_create(grp0=0, grp1=1, com0=0, com1=0, Rmin=0.8, dim=2, low=0.6, up=3.9, max_count=64)
    tot = None
    for lo in (0, 2):                                   # two batches of two frames: per-frame tables fold in frame order
        ctx.submit(system.frames(lo, 2), system.box9())
        ctx.accumulate(acc)
        table, numA = ctx.coord_read_frames(acc)
        for k in range(2):
            assert np.array_equal(ctx.coord_counts(acc, k), exp_counts[lo + k])
            assert int(numA[k]) == exp_numA[lo + k]
            assert np.array_equal(table[k], np.bincount(exp_counts[lo + k][exp_counts[lo + k] >= 0], minlength=64))
        tot = lib.coord_fold(table, numA, tot)
    total = 0.0
    for k in np.nonzero(tot)[0]:
        total += tot[k]
    lines = [f"{k:8d}\t{tot[k] / total:12.5f}" for k in np.nonzero(tot)[0]]
    assert lines == open(os.path.join(GOLD, "pair_dist_il.txt")).read().split("\n")[:-1]
    counts, st = ctx.read(acc, 64)
    assert int(counts.sum()) == sum(exp_numA) == int(st[lib.STAT_SELECTED]) and int(st[lib.STAT_DROPPED]) == 0


@pytest.mark.parametrize("case", [
    dict(sys=("water", 20000, 8.43), g=("OW", "OW"), Rmin=0.35, low=-100.0, up=100.0, cyl=None),        # self group: d = 0 excluded by d > 1e-3
    dict(sys=("il", 2500, 9.3), g=("CAT", "ANI"), Rmin=0.75, low=1.0, up=8.0, cyl=(4.6, 4.7, 0.5, 3.9)),  # calc_pair_dist.cpp filter
])
def test_coord_cells_and_brute_force_bit_exact(gpu_ctx_factory, case):
    kind, n, L = case["sys"]
    system = synth.water_box(n, L, seed=61, split_atoms=True) if kind == "water" else synth.ionic_liquid(n, L, seed=62)
    K = 2
    ctx = gpu_ctx_factory(system.natoms, K)
    g0, g1 = _groups(ctx, system, list(case["g"]))
    x, box = system.frames(0, K), system.box9()
    Lbox = port.lbox_from_box9(box)
    ctx.submit(x, box)
    kw = dict(grp0=0, grp1=1, com0=0, com1=0, Rmin=case["Rmin"], dim=2, low=case["low"], up=case["up"], max_count=32)
    if case["cyl"]:
        kw.update(use_cyl=1, cx=case["cyl"][0], cy=case["cyl"][1], Rlow=case["cyl"][2], Rup=case["cyl"][3])
    res = {}
    for cells in (0, 1):
        acc = ctx.coord_create(**kw)
        ctx.rdf_cells(acc, cells)
        ctx.accumulate(acc)
        res[cells] = [ctx.coord_counts(acc, k) for k in range(K)] + [ctx.read(acc, 32)]
    cyl = tuple(refprog.f32(v) for v in case["cyl"]) if case["cyl"] else None
    for k in range(K):
        p1 = port.load_position_center(x[k], g0["index"], g0["napm"], g0["w"][0], Lbox)
        p2 = port.load_position_center(x[k], g1["index"], g1["napm"], g1["w"][0], Lbox)
        exp, numA = port.pair_count(p1, p2, Lbox, refprog.f32(case["Rmin"]), 2, refprog.f32(case["low"]), refprog.f32(case["up"]), cyl)
        assert numA > 0 and exp.max() > 2
        assert np.array_equal(res[0][k], exp) and np.array_equal(res[1][k], exp)
    (c0, s0), (c1, s1) = res[0][K], res[1][K]
    assert np.array_equal(c0, c1) and int(s0[lib.STAT_SELECTED]) == int(s1[lib.STAT_SELECTED])
    if kind == "water":                                                 # (2 500 ion pairs are too few for a cell grid: both runs are brute force)
        assert int(s1[lib.STAT_PAIRS]) < 0.5 * int(s0[lib.STAT_PAIRS])  # the cell list really culled
        assert int(s0[lib.STAT_EXACT]) >= n * K                         # every self pair (d = 0) went through the FP64 path


# ---- FP32-bracketed fast paths of K2 / K4: proven answers must equal the literal FP64 ones -------------------------
@pytest.mark.parametrize("which", ["water", "il", "water_shaken"])
def test_fast_paths_validate_against_fp64(gpu_ctx_factory, which):
    """Mode 2 evaluates every molecule both ways and counts FP32-proven results that the FP64 code contradicts (must be
    0); modes 0 (FP64 only) and 1 (bracket + queue) must give identical histograms and statistics."""
    if which == "il":
        system, name, napm = synth.ionic_liquid(2000, 8.7, seed=91), "CAT", None
    else:
        system, name = synth.water_box(40000, 10.62, seed=77), "SOL"
    K = 4
    x, box = system.frames(0, K), system.box9()
    if which == "water_shaken":          # centres pushed onto bin edges / region bounds: exercises the undecided queue
        L = float(np.float32(box[0]))
        x = x.copy()
        x[:, :, 2] = np.round(x[:, :, 2] / (L / 64)) * np.float32(L / 64)
        x[:, :, 0] += np.float32(3.0 * L)        # whole system far outside [0, L)
        x[:, 1::9, 1] += np.float32(2.0 * L)     # one atom of every third molecule two boxes away: multi-box unwrap
    ctx = gpu_ctx_factory(system.natoms, K)
    _groups(ctx, system, [name])
    L = float(np.float32(box[0]))
    xoff = 3.0 * L if which == "water_shaken" else 0.0
    dens = dict(grp=0, com=0, dim=2, low=0.0, up=L, dbin=float(np.float32(L) / np.float32(64)), nbin=64, upper_inclusive=0,
                dim2=0, low2=0.1 * L + xoff, up2=0.9 * L + xoff)
    ori = dict(grp=0, com=0, DIMN=2, low=0.2 * L, up=0.9 * L, c1=0.5 * L + xoff, c2=0.5 * L, dr=0.05, CR=0.45 * L, Cr=0.3, nAngle=90,
               tp1=1, tp2=2, tp3=3, DIMNOrient=1, method=1)
    ori2 = dict(ori, method=2, nAngle=180, DIMNOrient=0, Cr=0.0, dr=0.011)
    ctx.submit(x, box)
    for kind, spec in (("density", dens), ("orient", ori), ("orient", ori2)):
        res = []
        for mode in (0, 1, 2):
            acc = ctx.density_create(**spec) if kind == "density" else ctx.orient_create(**spec)
            ctx.acc_fast(acc, mode)
            ctx.accumulate(acc)
            res.append(ctx.read(acc, ctx.acc_size(acc)))
        (c0, s0), (c1, s1), (c2, s2) = res
        assert c0.sum() > 0
        assert np.array_equal(c0, c1) and np.array_equal(c0, c2), kind
        assert int(s2[lib.STAT_MISPROVEN]) == 0, kind
        for st in (lib.STAT_DROPPED, lib.STAT_SELECTED, lib.STAT_FRAMES):
            assert int(s0[st]) == int(s1[st]) == int(s2[st]), (kind, st)
        nmol = ctx.nmol(0) * K
        assert int(s1[lib.STAT_EXACT]) == int(s2[lib.STAT_EXACT])
        if which != "water_shaken":
            assert int(s1[lib.STAT_EXACT]) < 0.05 * nmol, (kind, int(s1[lib.STAT_EXACT]), nmol)   # the bracket decides almost all
        else:
            assert int(s1[lib.STAT_EXACT]) > 0


@pytest.mark.parametrize("var", [
    dict(DIMN=2, DIMNOrient=2, method=1, tp=(1, 2, 3)),
    dict(DIMN=0, DIMNOrient=1, method=2, tp=(1, 3, 2)),
    dict(DIMN=1, DIMNOrient=0, method=2, tp=(1, 2, 3), box=(9.3, 7.1, 11.9)),
    dict(DIMN=2, DIMNOrient=0, method=1, tp=(1, 3, 2), box=(6.0, 12.0, 8.0), nAngle=180),
    dict(DIMN=2, DIMNOrient=2, method=1, tp=(2, 1, 3)),           # first vector atom is not the reference atom: generic path
])
def test_orient_three_site_packed_fast_path(gpu_ctx_factory, var):
    """The packed three-site fast path (orient_fast3: vectors = unwrapped offsets, FADD2/FFMA2, loop-free angle bin) against
    the literal FP64 kernel (mode 0), in validation mode (mode 2: proven-but-different must be 0) and against the oracle."""
    box = var.get("box", (8.1, 8.1, 8.1))
    system = synth.water_box(12000, box[0], seed=4242)
    K = 3
    x = system.frames(0, K).copy()
    box9 = np.tile(np.diag(np.array(box, np.float32)).reshape(1, 9), (K, 1)).astype(np.float32)
    if "box" in var:                       # squeeze the cubic synthetic box into the orthorhombic one, then re-wrap per atom
        x *= (np.array(box, np.float32) / np.float32(box[0]))
    x[1] += np.float32(2.0) * np.array(box, np.float32)          # a frame far outside [0, L)
    x[2, 4::9, 2] -= np.float32(box[2])                          # some H atoms one box below their O
    ctx = gpu_ctx_factory(system.natoms, K)
    (gi,) = _groups(ctx, system, ["SOL"])
    L = [float(np.float32(b)) for b in box]
    D = var["DIMN"]
    a_, b_ = [k for k in range(3) if k != D]
    spec = dict(grp=0, com=0, DIMN=D, low=0.1 * L[D], up=0.8 * L[D], c1=0.5 * L[a_], c2=0.5 * L[b_], dr=0.07, CR=0.42 * min(L[a_], L[b_]),
                Cr=0.2, nAngle=var.get("nAngle", 90), tp1=var["tp"][0], tp2=var["tp"][1], tp3=var["tp"][2], DIMNOrient=var["DIMNOrient"],
                method=var["method"])
    ctx.submit(x, box9)
    res = []
    for mode in (0, 1, 2):
        acc = ctx.orient_create(**spec)
        ctx.acc_fast(acc, mode)
        ctx.accumulate(acc)
        res.append(ctx.read(acc, ctx.acc_size(acc)))
    (c0, s0), (c1, s1), (c2, s2) = res
    assert c0.sum() > 0
    assert np.array_equal(c0, c1) and np.array_equal(c0, c2)
    assert int(s2[lib.STAT_MISPROVEN]) == 0
    for st in (lib.STAT_DROPPED, lib.STAT_SELECTED, lib.STAT_FRAMES):
        assert int(s0[st]) == int(s1[st]) == int(s2[st])
    assert 0 < int(s1[lib.STAT_EXACT]) < 0.05 * ctx.nmol(0) * K
    # and the literal kernel itself against the oracle port on the same frames
    nbin1, nA = int((np.float32(spec["CR"]) - np.float32(spec["Cr"])) / np.float32(spec["dr"])), spec["nAngle"]
    f32 = np.float32
    o = ncm = None
    for f in range(K):
        Lbox = port.lbox_from_box9(box9[f:f + 1])
        pos = port.load_position(x[f], gi["index"], 3, Lbox)
        pc = port.load_position_center(x[f], gi["index"], 3, gi["w"][0], Lbox)
        o, ncm, _ = port.orient_accum(pos, pc, Lbox, D, f32(spec["low"]), f32(spec["up"]), f32(spec["c1"]), f32(spec["c2"]), f32(spec["dr"]),
                                      f32(spec["CR"]), f32(spec["Cr"]), nA, spec["tp1"], spec["tp2"], spec["tp3"], spec["DIMNOrient"],
                                      spec["method"], o, ncm)
    assert np.array_equal(c0[:nA * nbin1].reshape(nA, nbin1), o) and np.array_equal(c0[nA * nbin1:], ncm)


# ---- general region accumulator (rest of the density family, A8) ----------------------------------------------------
def test_region_accumulator_against_numpy(gpu_ctx_factory):
    """b2a_region_create: half-open and inclusive filters, cylinder test, 1-D and 2-D bins, wrapped box count, restated in
    numpy on the oracle's centres (the ported programs are compared with the reference builds in test_gpu_dropin.py)."""
    system = synth.ionic_liquid(900, 6.3, seed=333)
    K = 3
    x, box = system.frames(0, K).copy(), system.box9()
    x[1] += np.float32(2.0) * np.float32(box[0])              # one frame two boxes away: exercises the wrap
    ctx = gpu_ctx_factory(system.natoms, K)
    (gi,) = _groups(ctx, system, ["ANI"])
    Lbox = port.lbox_from_box9(box)
    cen = [port.load_position_center(x[f], gi["index"], gi["napm"], gi["w"][0], Lbox) for f in range(K)]
    ctx.submit(x, box)
    f32 = np.float32
    L = float(Lbox[0])

    def run(**kw):
        acc = ctx.region_create(0, 0, **kw)
        ctx.accumulate(acc)
        return ctx.read(acc, ctx.acc_size(acc))

    # (1) z profile inside a cylinder (calc_density_distribution_in_pore.cpp:47-60)
    lo, up, cx, cy, Rm, nb = f32(0.7), f32(5.4), f32(3.1), f32(2.9), f32(2.2), 37
    dbin = float(f32(f32(up - lo) / f32(nb)))
    R2 = float(Rm) * float(Rm)
    exp, sel, drop = np.zeros(nb, np.uint64), 0, 0
    for c in cen:
        ok = (float(lo) <= c[:, 2]) & (c[:, 2] < float(up)) & ((c[:, 0] - float(cx)) ** 2 + (c[:, 1] - float(cy)) ** 2 <= R2)
        me = np.trunc((c[ok, 2] - float(lo)) / dbin).astype(np.int64)
        sel += int(ok.sum()); drop += int((me >= nb).sum())
        np.add.at(exp, me[me < nb], 1)
    got, st = run(filters=[(2, lo, up, 0)], cyl=(cx, cy, R2), bins=[(2, lo, dbin, nb)])
    assert np.array_equal(got, exp) and got.sum() > 0 and int(st[lib.STAT_SELECTED]) == sel and int(st[lib.STAT_DROPPED]) == drop
    # (2) 2-D map behind three filters, one of them inclusive (calc_areaDensityDistributionPore.cpp:56-68; slit variant :60-61)
    db = float(f32(0.21))
    n1, n2 = int(f32(f32(5.0) - f32(1.0)) / f32(0.21)), int(f32(f32(5.5) - f32(0.5)) / f32(0.21))
    exp = np.zeros((n2, n1), np.uint64)
    for c in cen:
        ok = (1.0 <= c[:, 0]) & (c[:, 0] < 5.0) & (0.5 <= c[:, 1]) & (c[:, 1] < 5.5) & (float(f32(0.3)) <= c[:, 2]) & (c[:, 2] <= float(f32(6.0)))
        m1 = np.trunc((c[ok, 0] - 1.0) / db).astype(np.int64); m2 = np.trunc((c[ok, 1] - 0.5) / db).astype(np.int64)
        keep = (m1 < n1) & (m2 < n2)
        np.add.at(exp, (m2[keep], m1[keep]), 1)
    got, st = run(filters=[(0, 1.0, 5.0, 0), (1, 0.5, 5.5, 0), (2, 0.3, 6.0, 1)], bins=[(0, 1.0, db, n1), (1, 0.5, db, n2)])
    assert np.array_equal(got.reshape(n2, n1), exp) and got.sum() > 0
    # (3) wrapped box count (calc_numdens_pore.cpp:12-27): float variable, double box length
    lo3, up3 = (f32(0.4), f32(1.1), f32(0.0)), (f32(5.9), f32(4.7), f32(6.1))
    cnt = 0
    for c in cen:
        ok = np.ones(len(c), bool)
        for m in range(3):
            t = c[:, m].astype(np.float32)
            for _ in range(8):
                t = np.where(t.astype(np.float64) > L, (t.astype(np.float64) - L).astype(np.float32), t)
            for _ in range(8):
                t = np.where(t < f32(0), (t.astype(np.float64) + L).astype(np.float32), t)
            ok &= (lo3[m] <= t) & (t < up3[m])
        cnt += int(ok.sum())
    got, st = run(filters=[(m, lo3[m], up3[m], 0) for m in range(3)], wrap=1)
    assert int(got[0]) == cnt > 0 and int(st[lib.STAT_FRAMES]) == K


# ---- boxes that change from frame to frame (NPT), short batches, degenerate selections -------------------------------
def test_box_changes_per_frame_and_degenerate_selections(gpu_ctx_factory):
    """Every frame of a batch carries its own box (ipp:72-74 re-reads Lbox per frame): RDF (all-pairs, symmetric-ineligible
    and cell mode), density and orientation against the oracle frame by frame; a batch shorter than the context's capacity;
    region filters that select nothing; one-molecule groups."""
    system = synth.ionic_liquid(400, 5.0, seed=4711)
    K = 3
    scale = np.array([[1.0, 1.0, 1.0], [1.03, 0.98, 1.01], [0.96, 1.02, 0.99]], np.float32)
    x = system.frames(0, K) * scale[:, None, :]
    box9 = np.zeros((K, 9), np.float32)
    for f in range(K):
        box9[f, 0::4] = np.float32(5.0) * scale[f]
    ctx = gpu_ctx_factory(system.natoms, 8)               # capacity 8, batches of 3 and 1
    gc, ga = _groups(ctx, system, ["CAT", "ANI"])
    Rm = refprog.f32(1.1)
    Rbin = port.rdf_rbin(Rm)
    f32 = refprog.f32
    spec = dict(nbin=3, low=0.4, up=4.7, dim=1, dim2=0, low2=0.2, up2=4.9)
    exp = {c: np.zeros((Rbin, 3), np.uint64) for c in (0, 1)}
    dexp = np.zeros(40, np.uint64)
    for f in range(K):
        Lbox = port.lbox_from_box9(box9[f])
        pc = port.load_position_center(x[f], gc["index"], gc["napm"], gc["w"][0], Lbox)
        pa = port.load_position_center(x[f], ga["index"], ga["napm"], ga["w"][0], Lbox)
        for c in (0, 1):
            port.rdf_accum(pc, pa, Lbox, 3, f32(spec["low"]), f32(spec["up"]), spec["dim"], spec["dim2"], f32(spec["low2"]), f32(spec["up2"]),
                           Rm, Rbin, exp[c])
        port.density_accum(pa, 2, f32(0.3), f32(4.5), float(np.float32(np.float32(4.2) / np.float32(40))), 40, dexp)
        # loaders see the frame's own box
        assert np.array_equal(ctx_centers_after_submit(ctx, x, box9, f, 0), pc)
    for cells in (0, 1):
        acc = ctx.rdf_create(grp0=0, grp1=1, com0=0, com1=0, Rm=Rm, Rbin=Rbin, **spec)
        ctx.rdf_tune(acc, 1, 0)
        ctx.rdf_cells(acc, cells)
        ctx.submit(x, box9)
        ctx.accumulate(acc)
        ctx.submit(x[:1], box9[:1])                       # a second, shorter batch into the same accumulator
        ctx.accumulate(acc)
        got, st = ctx.read(acc, 3 * Rbin)
        Lbox = port.lbox_from_box9(box9[0])
        extra = np.zeros((Rbin, 3), np.uint64)
        port.rdf_accum(port.load_position_center(x[0], gc["index"], gc["napm"], gc["w"][0], Lbox),
                       port.load_position_center(x[0], ga["index"], ga["napm"], ga["w"][0], Lbox), Lbox, 3, f32(spec["low"]), f32(spec["up"]),
                       spec["dim"], spec["dim2"], f32(spec["low2"]), f32(spec["up2"]), Rm, Rbin, extra)
        assert np.array_equal(got.reshape(Rbin, 3), exp[cells] + extra) and int(st[lib.STAT_FRAMES]) == K + 1
    den = ctx.density_create(grp=1, com=0, dim=2, low=0.3, up=4.5, dbin=float(np.float32(np.float32(4.2) / np.float32(40))), nbin=40,
                             upper_inclusive=0, dim2=-1, low2=0.0, up2=0.0)
    ctx.submit(x, box9)
    ctx.accumulate(den)
    assert np.array_equal(ctx.read(den, 40)[0], dexp)
    # filters that select nothing: zero histograms, zero selected, no error
    empty = ctx.rdf_create(grp0=0, grp1=1, com0=0, com1=0, Rm=Rm, Rbin=Rbin, nbin=2, low=50.0, up=60.0, dim=0, dim2=2, low2=-100.0, up2=100.0)
    ctx.accumulate(empty)
    got, st = ctx.read(empty, 2 * Rbin)
    assert got.sum() == 0 and int(st[lib.STAT_SELECTED]) == 0
    # one-molecule groups: a single pair, wrapped across the box
    ctx1 = gpu_ctx_factory(2, 1)
    for g in (0, 1):
        ctx1.set_group(g, np.array([g], np.int32), 1, np.ones(1), np.ones(1))
    one = ctx1.rdf_create(grp0=0, grp1=1, com0=1, com1=1, Rm=f32(1.0), Rbin=500, nbin=1, low=-100.0, up=100.0, dim=0, dim2=2, low2=-100.0, up2=100.0)
    ctx1.submit(np.array([[[0.05, 1.0, 1.0], [3.95, 1.0, 1.0]]], np.float32), np.array([4, 0, 0, 0, 4, 0, 0, 0, 4], np.float32))
    ctx1.accumulate(one)
    got, _ = ctx1.read(one, 500)
    e1, _ = port.rdf_accum(np.array([[0.05, 1.0, 1.0]], np.float32).astype(np.float64), np.array([[3.95, 1.0, 1.0]], np.float32).astype(np.float64),
                           np.array([4.0, 4.0, 4.0]), 1, -100.0, 100.0, 0, 2, -100.0, 100.0, f32(1.0), 500)
    assert got.sum() == 1 and np.array_equal(got.reshape(500, 1), e1)


def ctx_centers_after_submit(ctx, x, box9, f, grp):
    ctx.submit(x, box9)
    return ctx.centers(grp, 0, f)


# ---- cell-binned RDF: identical counts to the all-pairs evaluation ---------------------------------
def _rdf_gpu_cells(ctx, system, g0, g1, K, spec, Rm, Rbin, cells, paranoid=0, symmetric=0):
    _groups(ctx, system, [g0, g1])
    acc = ctx.rdf_create(grp0=0, grp1=1, com0=spec["com"], com1=spec["com"], nbin=spec["nbin"], low=spec["low"], up=spec["up"],
                         dim=spec["dim"], dim2=spec["dim2"], low2=spec["low2"], up2=spec["up2"], Rm=Rm, Rbin=Rbin)
    ctx.rdf_tune(acc, symmetric, paranoid)
    ctx.rdf_cells(acc, cells)
    ctx.submit(system.frames(0, K), system.box9())
    ctx.accumulate(acc)
    got, st = ctx.read(acc, spec["nbin"] * Rbin)
    return got.reshape(Rbin, spec["nbin"]), st


@pytest.mark.parametrize("case", [
    dict(sys=("water", 4000, 4.93), g=("OW", "OW"), K=2, Rm=0.6),
    dict(sys=("il", 3000, 9.9), g=("CAT", "ANI"), K=2, Rm=0.9, nbin=3, low=0.5, up=9.0, dim=1, dim2=0, low2=0.3, up2=9.5),
    dict(sys=("il", 4000, 10.9), g=("ANI", "ANI"), K=1, Rm=1.4, com=1),
    dict(sys=("water", 3500, 4.7), g=("OW", "OW"), K=1, Rm=0.5, paranoid=1),
    dict(sys=("il", 3500, 9.0), g=("CAT", "ANI"), K=2, Rm=0.8, box=(9.0, 7.0, 11.0)),
    # symmetric evaluation inside cell mode (same group on both sides, one slab): beyond-the-tile runs x 2 + tile diagonals
    dict(sys=("water", 6000, 5.65), g=("OW", "OW"), K=2, Rm=0.7, symmetric=1),
    dict(sys=("il", 4000, 10.9), g=("ANI", "ANI"), K=2, Rm=1.1, com=1, symmetric=1, box=(10.9, 9.0, 12.5)),
    dict(sys=("water", 3500, 4.7), g=("OW", "OW"), K=1, Rm=0.5, paranoid=1, symmetric=1),
    # same, but the region filter spreads the molecules over slabs: frames are not eligible and fall back to the full evaluation
    dict(sys=("water", 5000, 5.3), g=("OW", "OW"), K=2, Rm=0.6, symmetric=1, nbin=4, low=0.4, up=4.9, dim=2),
])
def test_rdf_cell_mode_bit_exact(gpu_ctx_factory, case):
    kind, n, L = case["sys"]
    system = synth.water_box(n, L, seed=81, split_atoms=True) if kind == "water" else synth.ionic_liquid(n, L, seed=82)
    if "box" in case:
        system.L = np.array(case["box"], np.float32)
    spec = dict(_DEF)
    spec.update({k: v for k, v in case.items() if k in _DEF})
    K = case["K"]
    exp, est, Rm, Rbin = refprog.oracle_rdf_region(system, K, case["g"][0], case["g"][1], return_counts=True, **spec)
    ctx = gpu_ctx_factory(system.natoms, K)
    got, st = _rdf_gpu_cells(ctx, system, case["g"][0], case["g"][1], K, spec, Rm, Rbin, cells=1, paranoid=case.get("paranoid", 0),
                             symmetric=case.get("symmetric", 0))
    assert np.array_equal(got, exp)
    assert int(st[lib.STAT_DROPPED]) == int(est[1]) and int(st[lib.STAT_SELECTED]) == int(est[2])
    n0 = len(refprog.group_info(system, case["g"][0])["index"]) // refprog.group_info(system, case["g"][0])["napm"]
    n1 = len(refprog.group_info(system, case["g"][1])["index"]) // refprog.group_info(system, case["g"][1])["napm"]
    assert int(st[lib.STAT_PAIRS]) < n0 * n1 * K          # cells were really used: fewer pairs evaluated than all-pairs


def test_rdf_cells_equal_brute_force_large(gpu_ctx_factory):
    """The reference's own check pattern (src/test/nb_cellSearch_anyBoxSize.cpp:288-299): accelerated structure ==
    brute force, exact integer equality, on a config-5-like box (r_max << L) too large for the CPU oracle."""
    n, L = 150000, 16.5
    system = synth.water_box(n, L, seed=20200105, split_atoms=True)
    Rm, spec = refprog.f32(1.5), dict(_DEF)
    Rbin = port.rdf_rbin(Rm)
    res = {}
    for cells in (0, 1):
        ctx = gpu_ctx_factory(system.natoms, 1)
        res[cells] = _rdf_gpu_cells(ctx, system, "OW", "OW", 1, spec, Rm, Rbin, cells=cells, symmetric=1 - cells)
        ctx.close()
    ctx = gpu_ctx_factory(system.natoms, 1)
    res[2] = _rdf_gpu_cells(ctx, system, "OW", "OW", 1, spec, Rm, Rbin, cells=1, symmetric=1)      # cells AND symmetric
    ctx.close()
    assert np.array_equal(res[0][0], res[1][0]) and np.array_equal(res[0][0], res[2][0]) and res[0][0].sum() > 0
    assert int(res[1][1][lib.STAT_PAIRS]) * 4 < int(res[0][1][lib.STAT_PAIRS])     # cells evaluate far fewer pairs
    assert int(res[2][1][lib.STAT_PAIRS]) * 1.8 < int(res[1][1][lib.STAT_PAIRS])   # ... and half of those when symmetric
    for st in (lib.STAT_DROPPED, lib.STAT_SELECTED):
        assert int(res[0][1][st]) == int(res[1][1][st]) == int(res[2][1][st])


@pytest.mark.parametrize("pair", ["blob-blob", "blob-sparse", "sparse-sparse", "blob-blob-slabs"])
def test_rdf_cells_inhomogeneous(gpu_ctx_factory, pair):
    """Cell mode away from the uniform liquid: a dense blob (columns holding twenty times the mean: a warp loops over several tiles;
    most columns empty), a sparse group (cells far below one molecule each, runs of one or two molecules, last tiles of a few
    molecules) and their cross pair -- against the oracle (bit-exact) and the all-pairs kernel, symmetric and not."""
    rng = np.random.default_rng(20201020)
    L, nA, nB, K = 12.0, 7000, 900, 2
    f32 = np.float32
    x = np.empty((K, nA + nB, 3), np.float32)
    for f in range(K):
        blob = rng.normal(6.0 + 0.3 * f, 0.55, size=(int(0.7 * nA), 3))
        x[f, :nA] = np.concatenate([blob, rng.uniform(0.0, L, size=(nA - len(blob), 3))]).astype(np.float32)
        x[f, nA:] = rng.uniform(-0.2 * L, 1.2 * L, size=(nB, 3)).astype(np.float32)      # some molecules outside the box: wrapped by the centre code
    box9 = np.array([L, 0, 0, 0, L, 0, 0, 0, L], np.float32)
    idx = {"blob": np.arange(nA, dtype=np.int32), "sparse": np.arange(nA, nA + nB, dtype=np.int32)}
    a, b = pair.split("-")[:2]
    slabs = pair.endswith("slabs")
    spec = dict(_DEF)
    if slabs: spec.update(nbin=3, low=2.0, up=10.5, dim=1, dim2=0, low2=1.0, up2=11.0)
    else: spec.update(nbin=1)
    Rm = f32(1.1); Rbin = port.rdf_rbin(Rm)
    Lbox = np.array([L, L, L], np.float64)
    exp = np.zeros((Rbin, spec["nbin"]), np.uint64)
    for f in range(K):
        ca = port.load_position_center(x[f], idx[a], 1, np.ones(1), Lbox)
        cb = port.load_position_center(x[f], idx[b], 1, np.ones(1), Lbox)
        h, _ = port.rdf_accum(ca, cb, Lbox, spec["nbin"], spec["low"], spec["up"], spec["dim"], spec["dim2"], spec["low2"], spec["up2"], Rm, Rbin)
        exp += h.astype(np.uint64)
    res = {}
    for cells, sym in ((0, 0), (1, 0), (1, 1)):
        ctx = gpu_ctx_factory(nA + nB, K)
        ctx.set_group(0, idx[a], 1, np.ones(1), np.ones(1))
        ctx.set_group(1, idx[b], 1, np.ones(1), np.ones(1))
        acc = ctx.rdf_create(grp0=0, grp1=1, com0=0, com1=0, nbin=spec["nbin"], low=spec["low"], up=spec["up"], dim=spec["dim"], dim2=spec["dim2"],
                             low2=spec["low2"], up2=spec["up2"], Rm=Rm, Rbin=Rbin)
        ctx.rdf_tune(acc, sym, 0); ctx.rdf_cells(acc, cells)
        ctx.submit(x, box9)
        ctx.accumulate(acc)
        got, st = ctx.read(acc, spec["nbin"] * Rbin)
        res[(cells, sym)] = (got.reshape(Rbin, spec["nbin"]), st)
        ctx.close()
    for key, (got, st) in res.items():
        assert np.array_equal(got, exp), key
    n0, n1 = len(idx[a]), len(idx[b])
    assert int(res[(1, 0)][1][lib.STAT_PAIRS]) < int(res[(0, 0)][1][lib.STAT_PAIRS]) <= n0 * n1 * K
    if a == b and not slabs:
        assert int(res[(1, 1)][1][lib.STAT_PAIRS]) < 0.6 * int(res[(1, 0)][1][lib.STAT_PAIRS])   # symmetric: about half the pairs


# ---- GmxHandleFull: heterogeneous molecule sizes -----------------------------------------------------
def test_handle_full_loaders_and_accumulators(gpu_ctx_factory):
    gold = np.load(os.path.join(GOLD, "loaders.npz"))
    system, idx, napm, m, q = GC.full_fixture()
    K = 3
    x, box = system.frames(0, K), system.box9()
    Lbox = port.lbox_from_box9(box)
    ctx = gpu_ctx_factory(system.natoms, K)
    ctx.set_group_full(0, idx, napm, m, q)
    ctx.set_group(1, system.groups["ANI"], 7, system.species[1].mass.astype(np.float64), system.species[1].charge.astype(np.float64))
    ctx.submit(x, box)
    nmol, mx = len(napm), int(napm.max())
    # against the reference's own outputs (frame LOADER_FRAME) and the oracle (all frames)
    assert np.array_equal(ctx.positions(0, mx, GC.LOADER_FRAME), gold["full.pos"])
    assert np.array_equal(ctx.centers(0, 0, GC.LOADER_FRAME), gold["full.cen0"], equal_nan=True)
    assert np.array_equal(ctx.centers(0, 2, GC.LOADER_FRAME), gold["full.cen2"], equal_nan=True)
    for f in range(K):
        assert np.array_equal(ctx.centers(0, 0, f), port.full_load_position_center(x[f], idx, napm, m, Lbox))
        assert np.array_equal(ctx.centers(0, 1, f), port.full_load_position_center(x[f], idx, napm, m, Lbox, geometry=True))
        assert np.array_equal(ctx.positions(0, mx, f), port.full_load_position(x[f], idx, napm, Lbox))
    # density and RDF accumulators on the heterogeneous group
    dbin = 0.1
    acc = ctx.density_create(grp=0, com=0, dim=1, low=0.2, up=2.8, dbin=dbin, nbin=26, upper_inclusive=0, dim2=-1, low2=0.0, up2=0.0)
    ctx.accumulate(acc)
    got, st = ctx.read(acc, 26)
    exp = np.zeros(26, np.uint64)
    for f in range(K):
        port.density_accum(port.full_load_position_center(x[f], idx, napm, m, Lbox), 1, refprog.f32(0.2), refprog.f32(2.8), dbin, 26, exp)
    assert np.array_equal(got, exp) and got.sum() > 0
    Rm = port.rdf_default_rm(Lbox); Rbin = port.rdf_rbin(Rm)
    racc = ctx.rdf_create(grp0=0, grp1=1, com0=0, com1=0, nbin=2, low=0.0, up=3.0, dim=2, dim2=0, low2=-100.0, up2=100.0, Rm=Rm, Rbin=Rbin)
    ctx.accumulate(racc)
    rgot, _ = ctx.read(racc, 2 * Rbin)
    rexp = np.zeros((Rbin, 2), np.uint64)
    gi = refprog.group_info(system, "ANI")
    for f in range(K):
        p1 = port.full_load_position_center(x[f], idx, napm, m, Lbox)
        p2 = port.load_position_center(x[f], gi["index"], 7, gi["w"][0], Lbox)
        port.rdf_accum(p1, p2, Lbox, 2, 0.0, 3.0, 2, 0, -100.0, 100.0, Rm, Rbin, rexp)
    assert np.array_equal(rgot.reshape(Rbin, 2), rexp) and rexp.sum() > 0
    # orientation histogram on the heterogeneous group: per molecule its own atom count, weights and total (literal FP64 path),
    # against the oracle fed with GmxHandleFull's own loaders (padded boxd + per-molecule centres)
    for method, tp in ((1, (1, 2, 3)), (2, (2, 5, 7))):
        kw = dict(grp=0, com=0, DIMN=2, low=0.2, up=2.9, c1=1.5, c2=1.5, dr=0.25, CR=2.0, Cr=0.0, nAngle=60, tp1=tp[0], tp2=tp[1], tp3=tp[2],
                  DIMNOrient=1, method=method)
        oacc = ctx.orient_create(**kw)
        ctx.accumulate(oacc)
        ogot, ost = ctx.read(oacc)
        eo = en = None
        for f in range(K):
            eo, en, _ = port.orient_accum(port.full_load_position(x[f], idx, napm, Lbox), port.full_load_position_center(x[f], idx, napm, m, Lbox), Lbox,
                                          2, refprog.f32(0.2), refprog.f32(2.9), refprog.f32(1.5), refprog.f32(1.5), refprog.f32(0.25), refprog.f32(2.0),
                                          refprog.f32(0.0), 60, tp[0], tp[1], tp[2], 1, method, eo, en)
        nb1 = len(en)
        assert np.array_equal(ogot[:nb1 * 60].reshape(60, nb1), eo) and np.array_equal(ogot[nb1 * 60:], en) and eo.sum() > 0
        assert int(ost[lib.STAT_FRAMES]) == K
    with pytest.raises(lib.B2AError, match="tp1..3 must lie in 1..napm"):      # an atom the smallest molecule does not have
        ctx.orient_create(grp=0, com=0, DIMN=2, low=0.0, up=3.0, c1=0.0, c2=0.0, dr=1.0, CR=2.0, Cr=0.0, nAngle=90, tp1=1, tp2=2, tp3=int(napm.min()) + 1,
                          DIMNOrient=2, method=1)
