"""Worker of tests/test_gpu_sharded.py::test_two_rank_nccl_allgather_in_library (one process per GPU, torchrun)."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, HERE)


def main():
    import torch
    import torch.distributed as dist
    import oracle_lib as orc
    from mctomo_b200 import capi, synth

    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    capi.init(local)
    capi.comm_init_torch(dist, dev)
    info = capi.comm_info()
    assert info["active"] and info["rank"] == rank and info["nranks"] == world
    nx = int(os.environ.get("MCT_TEST_NX", "16"))
    grid = synth.make_grid(nx, 10, 30)
    freqs = synth.freqs(6)
    for mode, pg, bad in ((0, 0, False), (0, 1, False), (0, 0, True), (1, 0, False), (1, 1, False), (1, 0, True)):
        capi.comm_set_mode(mode)   # 0: contiguous x-slabs; 1: balanced (every n-th distinct column of the sorted list)
        opts = capi.disp_opts(raylov=1, phaseGroup=pg, nmodes=0)
        pts, par = synth.generate_model(grid, 50, 321)
        if bad:  # an invalid column in the LAST slab only: every rank must see model_invalid after the all-reduce
            q = np.array([grid.xmax, grid.ymin, grid.zmin])
            par = par.copy()
            par[int(np.argmin(((pts - q) ** 2).sum(1))), 1] = 9.5
        ref = orc.forward_eval(pts, par, grid, freqs, phaseGroup=pg)
        capi.set_nuclei_batch(pts, par, np.array([0, len(pts)], np.int64))
        per = capi.slab_bounds(grid.nx, world, rank)[2]
        ncols = per * world * grid.ny
        nout = len(freqs)
        n = grid.nx * grid.ny * grid.nz
        z = lambda m, dt: torch.zeros(m, dtype=dt, device=dev)  # noqa: E731
        vp, vs, rho, sid = z(n, torch.float64), z(n, torch.float64), z(n, torch.float64), z(n, torch.int32)
        pv, gv, ie, fl = z(ncols * nout, torch.float64), z(ncols * nout, torch.float64), z(ncols, torch.int32), z(2, torch.int32)
        st = torch.cuda.current_stream().cuda_stream
        capi.forward_sharded_dev(grid, freqs, opts, vp.data_ptr(), vs.data_ptr(), rho.data_ptr(), sid.data_ptr(), pv.data_ptr(),
                                 gv.data_ptr(), ie.data_ptr(), fl.data_ptr(), st)
        torch.cuda.synchronize()
        assert bad or capi.comm_last_ms() > 0.0
        nc = grid.nx * grid.ny
        flags = fl.cpu().numpy()
        assert flags[0] == ref["model_invalid"], (rank, flags, ref["model_invalid"])
        if not bad:
            assert np.array_equal(pv.cpu().numpy()[: nc * nout].reshape(grid.nx, grid.ny, nout), ref["pvel"]), f"rank {rank} pvel"
            assert np.array_equal(ie.cpu().numpy()[:nc].reshape(grid.nx, grid.ny), ref["ierr"]), f"rank {rank} ierr"
            if pg:
                assert np.array_equal(gv.cpu().numpy()[: nc * nout].reshape(grid.nx, grid.ny, nout), ref["gvel"]), f"rank {rank} gvel"
    capi.comm_set_mode(0)
    print(f"SHARDED_OK rank {rank}", flush=True)
    dist.barrier()
    capi.comm_destroy()
    dist.destroy_process_group()
    capi.shutdown()


if __name__ == "__main__":
    main()
