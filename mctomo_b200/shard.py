"""(Test/bench plumbing; the production data plane is mct_comm_* / mct_forward_sharded_dev inside the library.)
Column sharding of ONE model over the ranks of a torch.distributed group (BASELINE config 5).

The reference shards nothing inside a chain except OpenMP loops over x (src/likelihood_surf.F90:197-206);
here the same x axis is cut into contiguous slabs, one per rank: (nz,ny,nx) and (np,ny,nx) arrays make a
slab one contiguous block, a column's result depends on nothing outside itself, so the only exchange is
ONE all-gather of the dispersion map (and a MAX all-reduce of the two status flags).  Backend-agnostic:
NCCL on the GPUs (bench.py), gloo in the CPU tests.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def slab_bounds(nx: int, world: int, rank: int):
    """1-based inclusive x range of `rank`; slabs have equal width ceil(nx/world), the last may be short or empty."""
    per = (nx + world - 1) // world
    lo = rank * per + 1
    hi = min(nx, (rank + 1) * per)
    return lo, hi, per


def allgather_map(local: torch.Tensor, nx: int, world: int, group=None) -> torch.Tensor:
    """local: (wx_rank, ny, nout) slab of this rank -> (nx, ny, nout) on every rank (one all_gather_into_tensor)."""
    per = (nx + world - 1) // world
    ny, nout = local.shape[1], local.shape[2]
    if local.shape[0] != per:  # short last slab: pad so that every rank contributes the same count
        pad = torch.zeros((per - local.shape[0], ny, nout), dtype=local.dtype, device=local.device)
        local = torch.cat([local, pad], 0)
    out = torch.empty((world * per, ny, nout), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, local.contiguous(), group=group)
    return out[:nx]


def combine_flags(flags: torch.Tensor, group=None) -> torch.Tensor:
    """flags int32[2] = {model_invalid, max condition code} of the local slab -> global values (check_model is an
    `any` over the whole grid, src/likelihood_surf.F90:631-646)."""
    dist.all_reduce(flags, op=dist.ReduceOp.MAX, group=group)
    return flags
