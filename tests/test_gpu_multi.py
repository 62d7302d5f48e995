"""Several GPUs behind one context (b2a_create_multi), the merge (b2a_finalize) and the per-frame series.

The driver's GPU tier has ONE device, so the dealing / merge / ordering logic is exercised with a context whose members all
sit on device 0 (NCCL cannot place two ranks on one device: the merge is then a local sum); the same tests run with distinct
devices and the real ncclAllReduce where the box has them (`gpurun --gpus 2`), and `test_two_processes_nccl` covers the
multi-process communicator (b2a_comm_init) there."""
import os
import subprocess
import sys

import numpy as np
import pytest

import refprog
from gmx2020postanalysis_b200 import lib
from oracle import port
from workloads import synth

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ndev():
    return lib.load().b2a_device_count()


def _device_sets():
    sets = [[0, 0, 0]]
    if _ndev() >= 2:
        sets.append(list(range(min(_ndev(), 4))))
    return sets


def _setup(ctx, system):
    """RDF (symmetric, the headline mode), density, per-frame region counter, coordination numbers on one context."""
    gi = synth.group_info(system, "SOL")
    for g in (0, 1):
        ctx.set_group(g, gi["index"], gi["napm"], gi["mass"], gi["charge"])
    Lbox = port.lbox_from_box9(system.box9())
    Rm = port.rdf_default_rm(Lbox)
    Rbin = port.rdf_rbin(Rm)
    rdf = ctx.rdf_create(grp0=0, grp1=1, com0=0, com1=0, nbin=5, low=-100.0, up=100.0, dim=0, dim2=2, low2=-100.0, up2=100.0, Rm=Rm, Rbin=Rbin)
    ctx.rdf_tune(rdf, 1, 0)
    L = float(system.L[2])
    den = ctx.density_create(grp=0, com=0, dim=2, low=0.0, up=L, dbin=L / 50, nbin=50, upper_inclusive=0, dim2=-1, low2=0.0, up2=0.0)
    reg = ctx.region_create(0, filters=[(2, 0.2 * L, 0.8 * L, 0)], bins=[(0, 0.0, L / 7, 7)], per_frame=1)
    coo = ctx.coord_create(grp0=0, grp1=1, com0=0, com1=0, Rmin=0.45, dim=2, low=0.0, up=L, max_count=64)
    return dict(rdf=rdf, den=den, reg=reg, coo=coo), Rbin


def _run(ctx, system, accs, nframes, K):
    f = 0
    while f < nframes:
        n = min(K, nframes - f)
        ctx.submit(system.frames(f, n), system.box9())
        for a in accs.values():
            ctx.accumulate(a)
        f += n


@pytest.mark.parametrize("devices", _device_sets(), ids=lambda d: "dev" + "".join(map(str, d)))
def test_team_equals_single_device(devices):
    """Batches dealt over the devices, merged by b2a_finalize: counters, stats and per-frame series are those of ONE device
    that saw every frame; series come back in trajectory order; reads are repeatable and accumulation may continue."""
    system = synth.water_box(900, 3.0, seed=91)
    K, nframes = 2, 11                                   # 6 batches (the last one ragged) over 3 members
    one = lib.Context(system.natoms, K, device=0)
    accs1, Rbin = _setup(one, system)
    _run(one, system, accs1, nframes, K)
    team = lib.Context(system.natoms, K, devices=devices)
    accs, _ = _setup(team, system)
    assert accs == accs1                                 # same ids on every member
    _run(team, system, accs, nframes, K)
    # every member took part: before the merge a member's read returns its PRIVATE counts (afterwards: the totals)
    shares = []
    for i in range(len(devices)):
        m = lib.Context.__new__(lib.Context)
        m.L, m._h, m._keep = team.L, lib._P(team.L.b2a_member(team._h, i)), []
        shares.append(int(m.read(accs["rdf"])[1][lib.STAT_FRAMES]))
        m._h = lib._P()                                  # owned by the leader
    assert sum(shares) == nframes and min(shares) >= K
    for name in accs:
        c1, s1 = one.read(accs1[name])
        c2, s2 = team.read(accs[name])
        assert np.array_equal(c1, c2), name
        for slot in (lib.STAT_FRAMES, lib.STAT_DROPPED, lib.STAT_SELECTED, lib.STAT_MISPROVEN):
            assert int(s1[slot]) == int(s2[slot]), (name, slot)
        assert int(s2[lib.STAT_FRAMES]) == nframes
    for name in ("reg", "coo"):
        i1, r1 = one.series(accs1[name])
        i2, r2 = team.series(accs[name])
        assert list(i2) == list(range(nframes)) and np.array_equal(i1, i2)
        assert np.array_equal(r1, r2) and r2.sum() > 0, name
    # a second read is the same; more frames keep accumulating on top of the private counts
    c2b, _ = team.read(accs["rdf"])
    assert np.array_equal(c2b, one.read(accs1["rdf"])[0])
    _run(one, system, accs1, 3, K)
    _run(team, system, accs, 3, K)
    assert np.array_equal(team.read(accs["rdf"])[0], one.read(accs1["rdf"])[0])
    assert int(team.read(accs["den"])[1][lib.STAT_FRAMES]) == nframes + 3
    # reset clears every member and the series
    team.reset(accs["reg"])
    assert team.series(accs["reg"])[0].size == 0 and int(team.read(accs["reg"])[0].sum()) == 0
    one.close()
    team.close()


def test_team_loaders_follow_the_current_batch():
    """The compatibility loaders on a multi-device context read the device that holds the current batch."""
    system = synth.water_box(300, 2.2, seed=92)
    team = lib.Context(system.natoms, 2, devices=[0, 0])
    gi = synth.group_info(system, "SOL")
    team.set_group(0, gi["index"], gi["napm"], gi["mass"], gi["charge"])
    Lbox = port.lbox_from_box9(system.box9())
    for b in range(3):
        x = system.frames(2 * b, 2)
        team.submit(x, system.box9())
        assert team.current_device() == b % 2
        for f in range(2):
            exp = port.load_position_center(x[f], gi["index"], gi["napm"], gi["w"][0], Lbox)
            assert np.array_equal(team.centers(0, 0, f), exp)
    team.close()


def test_single_context_finalize_and_series():
    """b2a_finalize on a plain context is the identity merge; the series of a single device is its accumulation order, or
    the order b2a_set_frame_base dictates."""
    system = synth.water_box(500, 2.6, seed=93)
    ctx = lib.Context(system.natoms, 2, device=0)
    accs, _ = _setup(ctx, system)
    ctx.set_frame_base(100)                               # this process reads frames 100.. of a longer trajectory
    _run(ctx, system, accs, 4, 2)
    before = ctx.read(accs["rdf"])[0].copy()
    ctx.finalize()
    assert np.array_equal(ctx.read(accs["rdf"])[0], before)
    idx, rec = ctx.series(accs["reg"])
    assert list(idx) == [100, 101, 102, 103] and rec.shape == (4, 7)
    per = ctx.region_read_frames(accs["reg"], 2)          # the last batch alone (the per-batch read still works)
    assert np.array_equal(per, rec[2:])
    ctx.close()


WORKER = r'''
import os, sys
import numpy as np
sys.path.insert(0, os.environ["B2A_ROOT"]); sys.path.insert(0, os.path.join(os.environ["B2A_ROOT"], "tests"))
import torch.distributed as dist
from gmx2020postanalysis_b200 import lib, dist as b2d
from workloads import synth
import test_gpu_multi as T
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo")                      # the launcher's channel: only carries the 128-byte id
system = synth.water_box(900, 3.0, seed=91)
K, nframes = 2, 12
ctx = lib.Context(system.natoms, K, device=rank)
b2d.comm_init(ctx)                                   # ncclCommInitRank inside libb2a
accs, Rbin = T._setup(ctx, system)
lo, hi = b2d.shard_frames(nframes, rank, world)
ctx.set_frame_base(lo)
f = lo
while f < hi:
    n = min(K, hi - f)
    ctx.submit(system.frames(f, n), system.box9())
    for a in accs.values(): ctx.accumulate(a)
    f += n
ctx.finalize()                                       # ncclAllReduce + ordered ncclAllGather of the series
out = {k: ctx.read(a)[0] for k, a in accs.items()}
idx, rec = ctx.series(accs["reg"])
np.savez(os.path.join(os.environ["B2A_OUT"], f"rank{rank}.npz"), idx=idx, rec=rec, **out)
ctx.close()
dist.barrier()
'''


@pytest.mark.skipif(_ndev() < 2, reason="needs two GPUs (run with gpurun --gpus 2)")
def test_two_processes_nccl(tmp_path):
    """One process per GPU: b2a_comm_init + b2a_finalize inside libb2a (ncclAllReduce of the histograms, ncclAllGather of the
    per-frame series); every rank ends up with the totals of a single-device run over all frames, series in frame order."""
    system = synth.water_box(900, 3.0, seed=91)
    one = lib.Context(system.natoms, 2, device=0)
    accs1, _ = _setup(one, system)
    _run(one, system, accs1, 12, 2)
    (tmp_path / "w.py").write_text(WORKER)
    env = dict(os.environ, B2A_ROOT=ROOT, B2A_OUT=str(tmp_path))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", "29517", str(tmp_path / "w.py")], capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-3000:]
    i1, r1 = one.series(accs1["reg"])
    for rank in range(2):
        z = np.load(tmp_path / f"rank{rank}.npz")
        for k, a in accs1.items():
            assert np.array_equal(z[k], one.read(a)[0]), (rank, k)
        assert np.array_equal(z["idx"], i1) and np.array_equal(z["rec"], r1)
    one.close()
