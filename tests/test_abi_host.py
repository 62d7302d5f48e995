"""CPU-side checks: the C-ABI library loads and exports every symbol include/b2a.h declares, fails loudly without
a GPU (no CPU fallback), and its host-only pieces (normalisation, synthetic generator, file round trip) agree
with the oracle / the reference."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

import refprog
from gmx2020postanalysis_b200 import dist, lib
from workloads import synth, trajio
from oracle import port

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "b2a.h")).read()
    declared = sorted(set(re.findall(r"\b(b2a_[a-z0-9_]+)\s*\(", hdr)))
    assert declared, "no prototypes parsed from include/b2a.h"
    L = lib.load()
    missing = [s for s in declared if not hasattr(L, s)]
    assert not missing, missing
    assert sorted(lib.SYMBOLS) == declared


def test_no_cpu_fallback():
    L = lib.load()
    if L.b2a_device_count() > 0:
        pytest.skip("a GPU is visible")
    h = C.c_void_p()
    rc = L.b2a_create(C.byref(h), 0, 100, 2)
    assert rc == 2 and b"no CPU fallback" in L.b2a_last_error()
    with pytest.raises(lib.B2AError):
        lib.Context(100, 2)


def test_product_never_imports_oracle():
    """The product path (package + C sources + drop-in header) must not reference oracle/."""
    bad = []
    for base in ("gmx2020postanalysis_b200", "include"):
        for dp, _, fns in os.walk(os.path.join(ROOT, base)):
            for fn in fns:
                if fn.endswith((".py", ".cu", ".cuh", ".c", ".h", ".hpp", ".ipp")) or "." not in fn:
                    txt = open(os.path.join(dp, fn), errors="ignore").read()
                    if re.search(r"(from|import)\s+oracle|liboracle|oracle/port|orc_[a-z]", txt):
                        bad.append(os.path.join(dp, fn))
    assert not bad, bad


def test_rdf_normalise_matches_oracle():
    rng = np.random.default_rng(0)
    nbin, Rbin = 4, 300
    counts = rng.integers(0, 5000, size=(Rbin, nbin)).astype(np.uint64)
    counts[:, 1] = 0                       # empty slab -> NaN row, as the reference prints
    g = lib.rdf_normalise(counts.ravel(), nbin, Rbin)
    assert np.array_equal(g, port.rdf_normalise(counts), equal_nan=True)
    assert np.isnan(g[:, 1]).all() and np.isfinite(g[:, 0]).all()


def test_synth_is_deterministic_and_wrapped():
    assert synth.synth_hash(1, 2, 3, 4, 5) == 0x26d9ff6af1517356   # pins the counter hash across hosts
    s = synth.water_box(50, 1.9, seed=7)
    a, b = s.frame(3), s.frame(3)
    assert np.array_equal(a, b) and a.dtype == np.float32
    assert (a >= 0).all() and (a < 1.9).all()
    # molecules straddle faces: some O-H distances exceed half a box before unwrapping
    d = np.abs(a[1::3] - a[0::3]).max()
    assert d > 0.95
    assert not np.array_equal(s.frame(4), a)
    il = synth.ionic_liquid(10, 2.5)
    assert il.natoms == 320 and il.species[0].napm == 25 and il.species[1].napm == 7


def test_numpy_generator_is_bit_identical_to_the_compiled_one(monkeypatch):
    """workloads/synth.py restates synth_gen.c in numpy (what bench.py --impl reference uses, so that the reference arm maps
    no library of this repo): same hash, same coordinates, for water and for the rejection-sampled quaternions of the blobs."""
    for mk in (lambda: synth.water_box(700, 2.9, seed=5), lambda: synth.ionic_liquid(40, 3.1, seed=6)):
        s = mk()
        a = s.frames(2, 2)
        monkeypatch.setenv("B2A_SYNTH_NUMPY", "1")
        assert synth.use_numpy()
        b = s.frames(2, 2)
        h = synth.synth_hash(1, 2, 3, 4, 5)
        monkeypatch.delenv("B2A_SYNTH_NUMPY")
        assert np.array_equal(a, b) and h == 0x26d9ff6af1517356


def test_shard_frames_partition():
    for F in (1, 7, 100, 1000):
        for G in (1, 2, 3, 4, 8):
            blocks = [dist.shard_frames(F, r, G) for r in range(G)]
            assert blocks[0][0] == 0 and blocks[-1][1] == F
            assert all(blocks[i][1] == blocks[i + 1][0] for i in range(G - 1))
            sizes = [b - a for a, b in blocks]
            assert max(sizes) - min(sizes) <= 1


def test_stub_trajectory_roundtrip(tmp_path):
    """gmxstub file formats: what trajio writes is what the reference's template1 program reads (runs to the end)."""
    exe = os.path.join(refprog.REFDIR, "template1")
    if not refprog.have_ref("template1"):
        pytest.skip("oracle/_ref not built")
    s = synth.water_box(64, 1.6)
    trj, top, ndx = trajio.write_system(str(tmp_path / "w"), s, 3, groups={"SOL": s.groups["SOL"]})
    env = dict(os.environ, LD_LIBRARY_PATH=os.path.join(ROOT, "build"))
    r = subprocess.run([exe, "-f", trj, "-s", top, "-n", ndx, "-o", str(tmp_path / "o.xvg")], capture_output=True, text=True,
                       env=env, input="0\n")
    assert r.returncode == 0, r.stderr
    assert "Nr. of molecules in group: 64" in r.stdout


def test_dropin_build_is_complete():
    """Where the reference was present at build time (this container), build/dropin holds every live program of the sweep table, the
    ported programs named in dropin/Makefile, and oracle/_ref the matching reference builds."""
    import re
    import dropin_sweep as SW
    dropin, ref = os.path.join(ROOT, "build", "dropin"), os.path.join(ROOT, "oracle", "_ref")
    if not os.path.isdir(dropin) or not os.path.exists(os.path.join(ref, "calc_rdf_region")):
        pytest.skip("drop-in programs not built (needs /root/reference at build time)")
    for name in SW.PROGRAMS:
        assert os.access(os.path.join(dropin, name), os.X_OK), name
        assert os.access(os.path.join(ref, name), os.X_OK), name
    mk = open(os.path.join(ROOT, "dropin", "Makefile")).read()
    fused = re.search(r"^FUSED\s*:=\s*(.*)$", mk, re.M).group(1).split()
    assert len(fused) >= 12
    for name in fused:
        assert os.access(os.path.join(dropin, name), os.X_OK), name
        assert os.path.exists(os.path.join(ROOT, "dropin", name + ".cpp")), name
