"""tools/shard_hetero.py -- slab vs balanced division of ONE chain's columns (mct_comm_set_mode 0 / 1) on a laterally
HETEROGENEOUS model: the nucleus density grows along x (3/4 of the nuclei in the right half), so contiguous x-slabs
carry very different numbers of layers.  Run under torchrun with 2+ GPUs; rank 0 prints per-rank K2 times."""
import os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, '.')
from mctomo_b200 import capi, synth

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
capi.init(local); capi.comm_init_torch(dist, dev)
grid = synth.make_grid(512, 512, 80); freqs = synth.freqs(30); nout = 30
rng = np.random.default_rng(7)
n = 3000
x = np.where(rng.uniform(size=n) < 0.75, rng.uniform(0, 5, n), rng.uniform(-5, 0, n))
pts = np.column_stack([x, rng.uniform(-5, 5, n), rng.uniform(0, 12, n)])
vs = 2.0 + pts[:, 2] / 3.0
par = np.column_stack([vs * float(np.float32(1.73)), vs, 2.35 + 0.036 * (vs * 1.73 - 3) ** 2])
capi.set_nuclei_batch(pts, par, np.array([0, n], np.int64))
opts = capi.disp_opts()
per = capi.slab_bounds(grid.nx, world, rank)[2]
nn = grid.nx * grid.ny * grid.nz; cols = per * world * grid.ny
z = lambda m, dt: torch.zeros(m, dtype=dt, device=dev)
vp, vs_, rho, sid = z(nn, torch.float64), z(nn, torch.float64), z(nn, torch.float64), z(nn, torch.int32)
pv, gv, ie, fl = z(cols * nout, torch.float64), z(cols * nout, torch.float64), z(cols, torch.int32), z(2, torch.int32)
st = torch.cuda.current_stream().cuda_stream
res = {}
for mode in (0, 1):
    capi.comm_set_mode(mode)
    for it in range(3):
        if it == 1:
            torch.cuda.synchronize(); dist.barrier(); capi.set_profiling(True); capi.kernel_times(reset=True)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True); a.record()
        capi.forward_sharded_dev(grid, freqs, opts, vp.data_ptr(), vs_.data_ptr(), rho.data_ptr(), sid.data_ptr(), pv.data_ptr(),
                                 gv.data_ptr(), ie.data_ptr(), fl.data_ptr(), st)
    b.record(); torch.cuda.synchronize()
    kt = capi.kernel_times(reset=True); capi.set_profiling(False)
    mine = torch.tensor([a.elapsed_time(b) / 2, kt["k2_ms"] / 2], dtype=torch.float64, device=dev)
    allr = [torch.zeros_like(mine) for _ in range(world)]
    dist.all_gather(allr, mine)
    res[mode] = (torch.stack(allr).cpu().numpy(), pv[: grid.nx * grid.ny * nout].clone())
if rank == 0:
    for mode in (0, 1):
        r = res[mode][0]
        print(f"mode {mode} ({'x-slabs' if mode == 0 else 'balanced'}): {r[:, 0].max():8.1f} ms per evaluation; per-rank K2 ms {np.round(r[:, 1], 1).tolist()}")
    print("maps identical:", bool(torch.equal(res[0][1], res[1][1])))
dist.barrier(); capi.comm_destroy(); dist.destroy_process_group(); capi.shutdown()
