"""world_size-2 gloo test of the N>1 path: x-slab column sharding of one model + one all-gather of the
dispersion map + MAX-reduction of the flags (mctomo_b200/shard.py).  The per-slab compute is the oracle
here (no GPU in this test); on the B200 box bench.py --config C5 runs the same plumbing over NCCL."""
import os
import socket
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, nx, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, HERE)
    import oracle_lib as orc
    from mctomo_b200 import shard, synth
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    grid = synth.make_grid(nx, 6, 30)
    pts, par = synth.generate_model(grid, 40, 99)
    freqs = synth.freqs(8)
    # every rank grids the model (nuclei are replicated: 48 B each) and solves only its slab
    vp, vs, rho = (np.zeros(grid.shape) for _ in range(3))
    sid = np.zeros(grid.shape, np.int32)
    orc.kdtree_to_grid(pts, par, grid, grid.full_box(), vp, vs, rho, sid)
    vp, rho = orc.vs2vp_rho(vs)
    lo, hi, per = shard.slab_bounds(grid.nx, world, rank)
    if rank == 1:
        vs_bad = vs.copy()  # only rank 1's slab sees the invalid column -> the flag must reach rank 0
    if hi >= lo:
        pv, gv, ie, cnt, _ = orc.surf_dispersion(vp, vs, rho, grid, (lo, hi, 1, grid.ny), freqs, nthreads=1)
        local = torch.from_numpy(pv)
    else:
        local = torch.zeros((0, grid.ny, len(freqs)), dtype=torch.float64)
    full = shard.allgather_map(local, grid.nx, world)
    flags = torch.tensor([1 if rank == 1 else 0, 0], dtype=torch.int32)
    shard.combine_flags(flags)
    if rank == 0:
        ref, _, _, _, _ = orc.surf_dispersion(vp, vs, rho, grid, (1, grid.nx, 1, grid.ny), freqs, nthreads=1)
        q.put((bool(np.array_equal(full.numpy(), ref)), flags.tolist(), list(full.shape)))
    dist.barrier()
    dist.destroy_process_group()


def _run(nx):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, nx, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    return res


def test_even_slabs():
    same, flags, shape = _run(8)
    assert same and flags == [1, 0] and shape == [8, 6, 8]


def test_uneven_last_slab():
    same, flags, shape = _run(7)
    assert same and shape == [7, 6, 8]


def test_slab_bounds_cover_everything():
    from mctomo_b200 import shard
    for nx in (1, 7, 8, 64, 1000, 1024):
        for world in (1, 2, 3, 8):
            cols = []
            for r in range(world):
                lo, hi, per = shard.slab_bounds(nx, world, r)
                cols += list(range(lo, hi + 1))
            assert cols == list(range(1, nx + 1))
