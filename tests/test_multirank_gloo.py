"""world_size-2 gloo test of the frame-sharded path on CPU: each rank accumulates its contiguous block of frames
into private u64 histograms, one all-reduce merges them, and the result is bit-identical to the single-rank
run (integer sums commute).  On the CPU the per-rank accumulation is played by the oracle; the sharding and the
hand-over of the communicator id are the product's host plumbing (gmx2020postanalysis_b200/dist.py).  The device
merge itself (b2a_finalize: ncclAllReduce / ncclAllGather inside libb2a) is covered by tests/test_gpu_multi.py."""
import os
import socket
import sys

import numpy as np
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port_no, nframes, q):
    for p in (ROOT, os.path.join(ROOT, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import refprog
    from gmx2020postanalysis_b200 import dist as b2d
    from workloads import synth
    from oracle import port
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port_no)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    system = synth.ionic_liquid(120, 3.4, seed=41)
    Lbox = port.lbox_from_box9(system.box9())
    Rm = port.rdf_default_rm(Lbox)
    Rbin = port.rdf_rbin(Rm)
    g0, g1 = refprog.group_info(system, "CAT"), refprog.group_info(system, "ANI")
    lo, hi = b2d.shard_frames(nframes, rank, world)
    rdf = np.zeros((Rbin, 2), dtype=np.uint64)
    for f in range(lo, hi):
        x = system.frame(f)
        p1 = port.load_position_center(x, g0["index"], g0["napm"], g0["w"][0], Lbox)
        p2 = port.load_position_center(x, g1["index"], g1["napm"], g1["w"][0], Lbox)
        port.rdf_accum(p1, p2, Lbox, 2, 0.0, 3.4, 0, 2, -100.0, 100.0, Rm, Rbin, rdf)
    merged = b2d.allreduce_host_counts(rdf.ravel())
    nfr = b2d.allreduce_host_counts(np.array([hi - lo], dtype=np.uint64))
    # the launcher's other job: carry the 128-byte communicator id from rank 0 to everybody (here a stand-in id: no GPU)
    uid = b2d.broadcast_id(lambda: bytes(range(128)))
    ok = np.array([uid == bytes(range(128))], dtype=np.uint64)
    ok = b2d.allreduce_host_counts(ok)
    if rank == 0:
        q.put((merged, int(nfr[0]), (lo, hi), int(ok[0])))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_merge_is_bit_identical():
    import refprog
    from workloads import synth
    nframes, world = 5, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port_no = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port_no, nframes, q)) for r in range(world)]
    for p in procs:
        p.start()
    merged, nfr, block0, id_ok = q.get(timeout=120)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert nfr == nframes and block0 == (0, 2) and id_ok == world
    system = synth.ionic_liquid(120, 3.4, seed=41)
    exp, _, _, Rbin = refprog.oracle_rdf_region(system, nframes, "CAT", "ANI", nbin=2, low=0.0, up=3.4, dim=0, dim2=2,
                                                return_counts=True)
    assert np.array_equal(merged.reshape(Rbin, 2), exp)
