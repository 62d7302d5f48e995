"""Real-trajectory ingestion: the XTC / TRR codecs of the libgromacs stand-in (gmxstub/xdrtraj.hpp).
libgromacs 2020.1 itself is not available and the reference holds no trajectory fixtures, so these formats are pinned
against (1) the codec's own writer: decoding returns exactly the quantised coordinates the format defines, for every code
path of the compression (small systems, water-like runs, huge coordinate ranges), frame by frame and through the
multi-threaded batch reader; (2) an independent pure-Python reader of the published formats (tests/xtc_classic.py, no
shared code, classic algorithm) that must see the same frames in the files the stand-in writes; (3) XTC and double-
precision TRR files packed by hand from the format description, which the stand-in must read; and the reference's own
analysis program gives the same output from an XTC file as from the raw-float file holding the same quantised
coordinates."""
import ctypes as C
import os

import numpy as np
import pytest

import refprog
import xtc_classic
from workloads import synth, trajio

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
STUB = os.path.join(ROOT, "build", "libgmxstub.so")
_P = C.c_void_p
TRX_READ_X, TRX_READ_V, TRX_READ_F = 1 << 0, 1 << 2, 1 << 4     # gmxstub/include/gromacs/fileio/trxio.h


@pytest.fixture(scope="module")
def stub():
    if not os.path.exists(STUB):
        pytest.skip("build/libgmxstub.so not built")
    L = C.CDLL(STUB)
    L.gmxstub_write_xtc.argtypes = [C.c_char_p, C.c_int, C.c_int, _P, _P, C.c_int, _P, C.c_float]
    L.gmxstub_write_trr.argtypes = [C.c_char_p, C.c_int, C.c_int, _P, _P, _P, _P, C.c_int, _P]
    L.gmxstub_read_all.argtypes = [C.c_char_p, C.c_int, C.c_int, _P, _P, _P, _P, _P, C.POINTER(C.c_int)]
    L.gmxstub_read_batches.argtypes = [C.c_char_p, C.c_int, C.c_int, _P, _P, _P]
    return L


def _ptr(a):
    return a.ctypes.data if a is not None else None


def write_xtc(L, path, x, box9, times=None, precision=1000.0):
    x = np.ascontiguousarray(x, np.float32)
    b = np.ascontiguousarray(box9, np.float32).reshape(-1, 9)
    t = np.ascontiguousarray(times, np.float32) if times is not None else None
    assert L.gmxstub_write_xtc(path.encode(), x.shape[1], x.shape[0], _ptr(x), _ptr(b), b.shape[0], _ptr(t), precision) == 0


def read_all(L, path, nmax, natoms, flags=TRX_READ_X):
    x = np.zeros((nmax, natoms, 3), np.float32); v = np.zeros_like(x); f = np.zeros_like(x)
    box = np.zeros((nmax, 9), np.float32); t = np.zeros(nmax, np.float32); na = C.c_int()
    n = L.gmxstub_read_all(path.encode(), nmax, flags, _ptr(x), _ptr(v), _ptr(f), _ptr(box), _ptr(t), C.byref(na))
    return n, na.value, x[:n], v[:n], f[:n], box[:n], t[:n]


def quantise(x, precision):
    """What an XTC file stores: int(x * precision +- 0.5) in float arithmetic, returned as int * (1 / precision)."""
    p = np.float32(precision)
    lf = x.astype(np.float32) * p
    lf = lf + np.where(lf >= 0, np.float32(0.5), np.float32(-0.5)).astype(np.float32)
    li = np.trunc(lf).astype(np.int32)
    return li.astype(np.float32) * (np.float32(1.0) / p)


@pytest.mark.parametrize("case", ["water", "il", "tiny", "huge_range", "precision100", "wide60", "wide66", "wide_runs30", "wide_runs400"])
def test_xtc_round_trip_is_exactly_the_quantised_coordinates(stub, tmp_path, case, monkeypatch):
    prec = 1000.0
    if case == "water":
        system = synth.water_box(1500, 3.6, seed=81); x = system.frames(0, 5); box = system.box9()
    elif case == "il":
        system = synth.ionic_liquid(200, 4.0, seed=82); x = system.frames(0, 3); box = system.box9()
    elif case == "tiny":                       # natoms <= 9: stored as plain floats, no quantisation
        x = np.random.default_rng(1).normal(size=(4, 7, 3)).astype(np.float32); box = np.eye(3, dtype=np.float32).ravel() * 5
    elif case == "huge_range":                 # coordinate span > 2^24 / precision: every component coded with its own bit width
        rng = np.random.default_rng(2)
        x = (rng.uniform(-20000, 20000, size=(2, 64, 3))).astype(np.float32); box = np.eye(3, dtype=np.float32).ravel() * 4e4
    elif case in ("wide60", "wide66"):         # packed triples of 57..64 bits and of more than 64 bits (128-bit unpacking)
        span = 500.0 if case == "wide60" else 2000.0
        x = np.random.default_rng(7).uniform(-span, span, size=(2, 50, 3)).astype(np.float32); box = np.eye(3, dtype=np.float32).ravel() * 2 * span
    elif case.startswith("wide_runs"):         # runs whose small triples are 45..70 bits wide: triplets of atoms scattered far apart
        rng = np.random.default_rng(8); spread = float(case[9:])
        cen = rng.uniform(-2000, 2000, size=(2, 40, 1, 3))
        x = (cen + rng.uniform(-spread, spread, size=(2, 40, 3, 3))).reshape(2, 120, 3).astype(np.float32); box = np.eye(3, dtype=np.float32).ravel() * 5000
    else:
        system = synth.water_box(800, 2.9, seed=83); x = system.frames(0, 2); box = system.box9(); prec = 100.0
    path = str(tmp_path / "t.xtc")
    times = np.arange(x.shape[0], dtype=np.float32) * 2.5
    write_xtc(stub, path, x, box, times, prec)
    n, natoms, got, _, _, gbox, gt = read_all(stub, path, x.shape[0] + 3, x.shape[1])
    assert n == x.shape[0] and natoms == x.shape[1]
    exp = x if case == "tiny" else quantise(x, prec)
    assert np.array_equal(got, exp)
    assert np.array_equal(gt, times) and np.array_equal(gbox, np.repeat(np.asarray(box, np.float32).reshape(1, 9), n, axis=0))
    # an independent reader of the published format (tests/xtc_classic.py: plain bit stream, big-integer triples) sees the
    # same frames: header layout, bit order, run / swap logic and the adaptive width all agree with the stand-in's codec
    indep = xtc_classic.read_frames(path)
    assert len(indep) == n and all(fr["natoms"] == natoms for fr in indep)
    assert np.array_equal(np.stack([fr["x"] for fr in indep]), got)
    assert [fr["time"] for fr in indep] == list(times) and all(np.array_equal(fr["box"], gbox[0]) for fr in indep)
    if case == "water":                        # neighbours are coded as small differences: well under half the raw size
        assert os.path.getsize(path) < 0.45 * x.nbytes
    # the batch reader decodes the frames on several threads: same result, any batch size
    for batch in (1, 2, 16):
        bx = np.zeros_like(x); bb = np.zeros((x.shape[0], 9), np.float32); bt = np.zeros(x.shape[0], np.float32)
        assert stub.gmxstub_read_batches(path.encode(), batch, x.shape[0], _ptr(bx), _ptr(bb), _ptr(bt)) == x.shape[0]
        assert np.array_equal(bx, exp) and np.array_equal(bt, times)
    # the stand-in maps XTC files and decodes in place; the FILE* path (pipes, GMXSTUB_NO_MMAP) gives the same frames
    monkeypatch.setenv("GMXSTUB_NO_MMAP", "1")
    n2, _, got2, _, _, _, gt2 = read_all(stub, path, x.shape[0] + 3, x.shape[1])
    assert n2 == n and np.array_equal(got2, exp) and np.array_equal(gt2, times)
    bx = np.zeros_like(x); bb = np.zeros((x.shape[0], 9), np.float32); bt = np.zeros(x.shape[0], np.float32)
    assert stub.gmxstub_read_batches(path.encode(), 2, x.shape[0], _ptr(bx), _ptr(bb), _ptr(bt)) == x.shape[0]
    assert np.array_equal(bx, exp)


def test_xtc_decoder_survives_corrupt_input(tmp_path):
    """Bit flips in header and stream, truncation, with and without in-place slack, under ASan + UBSan: the decoder either
    decodes or reports failure; it never reads outside the buffer it was given (tests/fuzz_xtc.cpp)."""
    import subprocess
    exe = str(tmp_path / "fuzz_xtc")
    cc = subprocess.run(["g++", "-O1", "-g", "-std=c++17", "-fsanitize=address,undefined", "-fno-sanitize-recover=undefined",
                         "-I" + os.path.join(ROOT, "gmxstub"), os.path.join(ROOT, "tests", "fuzz_xtc.cpp"), "-o", exe],
                        capture_output=True, text=True)
    if cc.returncode != 0:
        pytest.skip("sanitizer build not available: " + cc.stderr[-200:])
    r = subprocess.run([exe, "30000"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    assert "decoded" in r.stdout and "rejected" in r.stdout


def test_trr_round_trip_exact(stub, tmp_path, monkeypatch):
    system = synth.ionic_liquid(50, 2.8, seed=84)
    K = 4
    x = system.frames(0, K)
    rng = np.random.default_rng(3)
    v = rng.normal(size=x.shape).astype(np.float32); f = (rng.normal(size=x.shape) * 500).astype(np.float32)
    box = np.ascontiguousarray(system.box9(), np.float32).reshape(1, 9)
    path = str(tmp_path / "t.trr")
    assert stub.gmxstub_write_trr(path.encode(), x.shape[1], K, _ptr(x), _ptr(v), _ptr(f), _ptr(box), 1, None) == 0
    n, natoms, gx, gv, gf, gbox, gt = read_all(stub, path, K + 2, x.shape[1], TRX_READ_X | TRX_READ_V | TRX_READ_F)
    assert n == K and natoms == x.shape[1]
    assert np.array_equal(gx, x) and np.array_equal(gv, v) and np.array_equal(gf, f)
    assert np.array_equal(gt, np.arange(K, dtype=np.float32))
    indep = xtc_classic.read_trr_frames(path)          # independent reader of the published layout
    assert len(indep) == K and all(fr["natoms"] == x.shape[1] for fr in indep)
    assert all(np.array_equal(fr["x"], x[k]) and np.array_equal(fr["v"], v[k]) and np.array_equal(fr["f"], f[k])
               and np.array_equal(fr["box"], box[0]) and fr["step"] == k for k, fr in enumerate(indep))
    # coordinates only: v / f blocks are skipped
    n, _, gx, gv, _, _, _ = read_all(stub, path, K, x.shape[1], TRX_READ_X)
    assert n == K and np.array_equal(gx, x) and not gv.any()
    # batch reader (coordinates of K frames converted on several threads out of the mapped file), any batch size
    for batch in (1, 3, 16):
        bx = np.zeros_like(x); bb = np.zeros((K, 9), np.float32); bt = np.zeros(K, np.float32)
        assert stub.gmxstub_read_batches(path.encode(), batch, K, _ptr(bx), _ptr(bb), _ptr(bt)) == K
        assert np.array_equal(bx, x) and np.array_equal(bt, np.arange(K, dtype=np.float32)) and np.array_equal(bb, np.repeat(box, K, axis=0))
    # the FILE* path (no mapping) reads the same frames
    monkeypatch.setenv("GMXSTUB_NO_MMAP", "1")
    n, _, gx, gv, gf, _, _ = read_all(stub, path, K + 2, x.shape[1], TRX_READ_X | TRX_READ_V | TRX_READ_F)
    assert n == K and np.array_equal(gx, x) and np.array_equal(gv, v) and np.array_equal(gf, f)
    bx = np.zeros_like(x); bb = np.zeros((K, 9), np.float32); bt = np.zeros(K, np.float32)
    assert stub.gmxstub_read_batches(path.encode(), 3, K, _ptr(bx), _ptr(bb), _ptr(bt)) == K and np.array_equal(bx, x)
    monkeypatch.delenv("GMXSTUB_NO_MMAP")
    # a file without velocities
    path2 = str(tmp_path / "x.trr")
    assert stub.gmxstub_write_trr(path2.encode(), x.shape[1], K, _ptr(x), None, None, _ptr(box), 1, None) == 0
    n, _, gx, _, _, _, _ = read_all(stub, path2, K, x.shape[1], TRX_READ_X | TRX_READ_V)
    assert n == K and np.array_equal(gx, x)


def test_xtc_written_elsewhere_without_runs(stub, tmp_path):
    """An XTC file packed by hand in Python from the published format by the simplest legal writer (every atom as a full
    triple, run flag 0): no stand-in encoder involved, so the stand-in's decoder is checked against the format itself."""
    import struct
    rng = np.random.default_rng(6)
    K, natoms, prec = 2, 40, 1000.0
    ints = rng.integers(-3000, 9000, size=(K, natoms, 3))
    blob = b""
    for k in range(K):
        lo, hi = ints[k].min(axis=0), ints[k].max(axis=0)
        sizes = [int(hi[q] - lo[q] + 1) for q in range(3)]
        nbits = (sizes[0] * sizes[1] * sizes[2]).bit_length()
        stream, count = 0, 0
        for a in range(natoms):
            u = [int(ints[k, a, q] - lo[q]) for q in range(3)]
            val, left = (u[0] * sizes[1] + u[1]) * sizes[2] + u[2], nbits
            while left > 8:                               # least significant piece first
                stream = (stream << 8) | (val & 0xFF); val >>= 8; left -= 8; count += 8
            stream = (stream << left) | val; count += left
            stream <<= 1; count += 1                      # run flag 0
        pad = -count % 8
        data = (stream << pad).to_bytes((count + pad) // 8, "big")
        blob += struct.pack(">iiif", 1995, natoms, k, 1.5 * k) + (np.eye(3) * 12).astype(">f4").tobytes()
        blob += struct.pack(">if", natoms, prec) + struct.pack(">6i", *[int(t) for t in lo], *[int(t) for t in hi])
        blob += struct.pack(">ii", 20, len(data)) + data + b"\0" * (-len(data) % 4)
    path = str(tmp_path / "hand.xtc")
    open(path, "wb").write(blob)
    n, na, gx, _, _, gbox, gt = read_all(stub, path, K + 1, natoms)
    assert n == K and na == natoms
    assert np.array_equal(gx, ints.astype(np.float32) * (np.float32(1.0) / np.float32(prec)))
    assert np.array_equal(gt, np.float32(1.5) * np.arange(K, dtype=np.float32)) and gbox[0][0] == 12 and gbox[0][4] == 12
    assert np.array_equal(np.stack([fr["x"] for fr in xtc_classic.read_frames(path)]), gx)


def test_trr_written_elsewhere_double_precision(stub, tmp_path):
    """A double-precision TRR frame sequence laid out by hand from the published format (no stand-in writer involved):
    header with strlen + 1 and the XDR string, 8-byte reals for time / lambda / box / x / v."""
    import struct
    rng = np.random.default_rng(5)
    K, natoms = 3, 11
    x = rng.normal(size=(K, natoms, 3)); v = rng.normal(size=(K, natoms, 3))
    box = np.diag([3.0, 4.0, 5.0]).ravel()
    blob = b""
    for k in range(K):
        blob += struct.pack(">iii", 1993, 13, 12) + b"GMX_trn_file"
        blob += struct.pack(">13i", 0, 0, 72, 0, 0, 0, 0, natoms * 24, natoms * 24, 0, natoms, 10 * k, 0)
        blob += struct.pack(">dd", 0.5 * k, 0.0) + box.astype(">f8").tobytes() + x[k].astype(">f8").tobytes() + v[k].astype(">f8").tobytes()
    path = str(tmp_path / "d.trr")
    open(path, "wb").write(blob)
    n, na, gx, gv, _, gbox, gt = read_all(stub, path, K + 1, natoms, TRX_READ_X | TRX_READ_V)
    assert n == K and na == natoms
    assert np.array_equal(gx, x.astype(np.float32)) and np.array_equal(gv, v.astype(np.float32))
    assert np.array_equal(gbox[0], box.astype(np.float32)) and np.array_equal(gt, np.float32(0.5) * np.arange(K, dtype=np.float32))
    assert [fr["step"] for fr in xtc_classic.read_trr_frames(path)] == [0, 10, 20]


@pytest.mark.parametrize("select", [[], ["-b", "2", "-e", "7", "-dt", "2"]], ids=["all_frames", "b_e_dt"])
def test_reference_program_reads_xtc_like_raw(stub, tmp_path, select):
    """The reference's own calc_rdf_region (oracle/_ref) on an XTC file == on the raw file of the quantised coordinates,
    also when -b / -e / -dt make the reader skip frames (in the mapped file: header walk without touching the blocks)."""
    if not refprog.have_ref("calc_rdf_region"):
        pytest.skip("oracle/_ref not built")
    import subprocess
    system = synth.water_box(500, 2.5, seed=85, split_atoms=True)
    K = 9 if select else 3
    xq = quantise(system.frames(0, K), 1000.0)
    outs = []
    for kind in ("raw", "xtc"):
        d = tmp_path / kind
        d.mkdir()
        trj = str(d / ("sys.b2t" if kind == "raw" else "sys.xtc"))
        if kind == "raw":
            trajio.write_trajectory(trj, xq, system.box9())
        else:
            write_xtc(stub, trj, system.frames(0, K), system.box9())
        trajio.write_topology(str(d / "sys.b2p"), system)
        trajio.write_index(str(d / "sys.ndx"), {"OW": system.groups["OW"]})
        env = dict(os.environ, LD_LIBRARY_PATH=os.path.join(ROOT, "build") + ":" + os.environ.get("LD_LIBRARY_PATH", ""))
        r = subprocess.run([os.path.join(refprog.REFDIR, "calc_rdf_region"), "-f", trj, "-s", str(d / "sys.b2p"), "-n", str(d / "sys.ndx"),
                            "-o", str(d / "o.xvg")] + select, input="0 0\n", capture_output=True, text=True, env=env, cwd=str(d))
        assert r.returncode == 0, r.stderr[-500:]
        outs.append(refprog.data_lines(str(d / "o.xvg")))
    assert len(outs[0]) > 100 and outs[0] == outs[1]
    if select:       # and the selection really selected: frames 2, 4, 6 only -> not the all-frames answer of the same file
        r = subprocess.run([os.path.join(refprog.REFDIR, "calc_rdf_region"), "-f", trj, "-s", str(d / "sys.b2p"), "-n", str(d / "sys.ndx"),
                            "-o", str(d / "all.xvg")], input="0 0\n", capture_output=True, text=True, env=env, cwd=str(d))
        assert r.returncode == 0 and refprog.data_lines(str(d / "all.xvg")) != outs[1]
