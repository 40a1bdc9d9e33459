"""The column-sharded path (BASELINE config 5) on the GPU, through the C ABI.

One GPU: the slab arguments of mct_forward_batch_dev (K1 window, x offset, slab-scoped check_model, outputs indexed
relative to the slab) for 2, 3 and 8 slabs incl. an uneven last slab and an invalid column that only ONE slab sees;
the concatenation must equal the unsharded call and the oracle bit for bit.  Two GPUs (skipped otherwise): the same
through mct_comm_init + mct_forward_sharded_dev, i.e. NCCL in-place all-gather inside the library, launched as two
processes."""
import os
import subprocess
import sys

import numpy as np
import pytest

import oracle_lib as orc
from mctomo_b200 import synth
from mctomo_b200.capi import disp_opts

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _dev_arrays(torch, grid, nout, ncols_out):
    n = grid.nx * grid.ny * grid.nz
    f = lambda m, dt: torch.zeros(m, dtype=dt, device="cuda")  # noqa: E731
    return dict(vp=f(n, torch.float64), vs=f(n, torch.float64), rho=f(n, torch.float64), sid=f(n, torch.int32),
                pv=f(ncols_out * nout, torch.float64), gv=f(ncols_out * nout, torch.float64), ie=f(ncols_out, torch.int32),
                fl=f(2, torch.int32))


@pytest.mark.parametrize("nx,nslabs", [(16, 2), (17, 3), (19, 8), (5, 8)])
@pytest.mark.parametrize("pg", [0, 1])
def test_slabs_concatenate_to_the_unsharded_result(mct, nx, nslabs, pg):
    import torch
    grid = synth.make_grid(nx, 9, 30)
    freqs = synth.freqs(6)
    opts = disp_opts(raylov=1, phaseGroup=pg, nmodes=0)
    nout = len(freqs)
    pts, par = synth.generate_model(grid, 60, 123 + nx)
    ref = orc.forward_eval(pts, par, grid, freqs, phaseGroup=pg)
    st = torch.cuda.current_stream().cuda_stream
    off = np.array([0, len(pts)], np.int64)
    mct.set_nuclei_batch(pts, par, off)
    # unsharded, through the same entry point
    A = _dev_arrays(torch, grid, nout, grid.nx * grid.ny)
    mct.forward_batch_dev(grid, 1, freqs, opts, A["vp"].data_ptr(), A["vs"].data_ptr(), A["rho"].data_ptr(), A["sid"].data_ptr(),
                          A["pv"].data_ptr(), A["gv"].data_ptr(), A["ie"].data_ptr(), A["fl"].data_ptr(), st)
    torch.cuda.synchronize()
    whole_pv = A["pv"].cpu().numpy().reshape(grid.nx, grid.ny, nout)
    assert np.array_equal(whole_pv, ref["pvel"])
    # sharded: every slab into its chunk of ONE full-size buffer, as mct_forward_sharded_dev lays it out
    per = mct.slab_bounds(grid.nx, nslabs, 0)[2]
    B = _dev_arrays(torch, grid, nout, per * nslabs * grid.ny)
    flags = []
    covered = []
    for r in range(nslabs):
        lo, hi, _ = mct.slab_bounds(grid.nx, nslabs, r)
        if hi < lo:
            continue
        covered += list(range(lo, hi + 1))
        c0 = r * per * grid.ny
        mct.forward_batch_dev(grid, 1, freqs, opts, B["vp"].data_ptr(), B["vs"].data_ptr(), B["rho"].data_ptr(), B["sid"].data_ptr(),
                              B["pv"].data_ptr() + c0 * nout * 8, B["gv"].data_ptr() + c0 * nout * 8, B["ie"].data_ptr() + c0 * 4,
                              B["fl"].data_ptr(), st, slab=(lo, hi))
        torch.cuda.synchronize()
        flags.append(B["fl"].cpu().numpy().copy())
    assert covered == list(range(1, grid.nx + 1))
    ncol = grid.nx * grid.ny
    assert np.array_equal(B["pv"].cpu().numpy()[: ncol * nout].reshape(grid.nx, grid.ny, nout), ref["pvel"])
    assert np.array_equal(B["ie"].cpu().numpy()[:ncol].reshape(grid.nx, grid.ny), ref["ierr"])
    if pg:
        assert np.array_equal(B["gv"].cpu().numpy()[: ncol * nout].reshape(grid.nx, grid.ny, nout), ref["gvel"])
    # the gridded model assembled slab by slab is the whole model
    for k in ("vp", "vs", "rho", "sid"):
        assert torch.equal(A[k], B[k]), k
    assert np.array_equal(B["sid"].cpu().numpy().reshape(grid.shape), ref["sites_id"])
    assert all((f == 0).all() for f in flags)


def test_invalid_column_is_seen_only_by_its_slab(mct):
    """check_model is scoped to the slab: the slab holding the offending column raises its flag, the others do not --
    the MAX all-reduce of mct_forward_sharded_dev then makes it global (likelihood_surf.F90:631-646)."""
    import torch
    grid = synth.make_grid(12, 8, 20)
    freqs = synth.freqs(4)
    opts = disp_opts()
    pts, par = synth.generate_model(grid, 30, 9)
    # make the nucleus nearest to the top of column (ix=11, iy=3) faster than everything below it
    q = np.array([grid.xmin + 10 * grid.dx, grid.ymin + 2 * grid.dy, grid.zmin])
    j = int(np.argmin(((pts - q) ** 2).sum(1)))
    par = par.copy()
    par[j, 1] = 9.5
    ref = orc.forward_eval(pts, par, grid, freqs)
    assert ref["model_invalid"] == 1
    st = torch.cuda.current_stream().cuda_stream
    mct.set_nuclei_batch(pts, par, np.array([0, len(pts)], np.int64))
    B = _dev_arrays(torch, grid, len(freqs), grid.nx * grid.ny)
    seen = []
    for r in range(3):
        lo, hi, per = mct.slab_bounds(grid.nx, 3, r)
        mct.forward_batch_dev(grid, 1, freqs, opts, B["vp"].data_ptr(), B["vs"].data_ptr(), B["rho"].data_ptr(), B["sid"].data_ptr(),
                              B["pv"].data_ptr() + r * per * grid.ny * len(freqs) * 8, B["gv"].data_ptr() + r * per * grid.ny * len(freqs) * 8,
                              B["ie"].data_ptr() + r * per * grid.ny * 4, B["fl"].data_ptr(), st, slab=(lo, hi))
        torch.cuda.synchronize()
        seen.append(int(B["fl"][0].item()))
    vs = B["vs"].cpu().numpy().reshape(grid.shape)
    bad_x = sorted(set(np.nonzero((vs[:, :, 1:] < vs[:, :, :1]).any(2))[0] + 1))
    expect = [int(any(mct.slab_bounds(grid.nx, 3, r)[0] <= x <= mct.slab_bounds(grid.nx, 3, r)[1] for x in bad_x)) for r in range(3)]
    assert seen == expect and max(seen) == 1 and min(seen) == 0


def test_sharded_entry_point_single_rank(mct):
    """Without a communicator mct_forward_sharded_dev is the whole-grid evaluation (nranks = 1)."""
    import torch
    grid = synth.make_grid(10, 7, 24)
    freqs = synth.freqs(5)
    opts = disp_opts(phaseGroup=1)
    pts, par = synth.generate_model(grid, 25, 4)
    ref = orc.forward_eval(pts, par, grid, freqs, phaseGroup=1)
    mct.set_nuclei_batch(pts, par, np.array([0, len(pts)], np.int64))
    A = _dev_arrays(torch, grid, len(freqs), grid.nx * grid.ny)
    st = torch.cuda.current_stream().cuda_stream
    mct.forward_sharded_dev(grid, freqs, opts, A["vp"].data_ptr(), A["vs"].data_ptr(), A["rho"].data_ptr(), A["sid"].data_ptr(),
                            A["pv"].data_ptr(), A["gv"].data_ptr(), A["ie"].data_ptr(), A["fl"].data_ptr(), st)
    torch.cuda.synchronize()
    assert np.array_equal(A["pv"].cpu().numpy().reshape(grid.nx, grid.ny, -1), ref["pvel"])
    assert np.array_equal(A["gv"].cpu().numpy().reshape(grid.nx, grid.ny, -1), ref["gvel"])
    assert mct.comm_info()["active"] is False and mct.comm_last_ms() == 0.0


def _gpu_count():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.skipif(_gpu_count() < 2, reason="needs two GPUs (NCCL all-gather inside the library)")
@pytest.mark.parametrize("nx", [16, 13])
def test_two_rank_nccl_allgather_in_library(nx):
    """Two processes, one GPU each: mct_comm_init (id moved by torch.distributed's broadcast, as MPI_Bcast would),
    mct_forward_sharded_dev, then EVERY rank compares the gathered full maps with the oracle."""
    env = dict(os.environ, MCT_TEST_NX=str(nx))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", str(29500 + nx), os.path.join(ROOT, "tests", "sharded_worker.py")]
    r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert r.stdout.count("SHARDED_OK") == 2, r.stdout[-3000:]
