"""Multi-GPU partition of one Fisher job: one process per GPU (torch.distributed), independent slices,
one small all-reduce.

The Fisher matrix is a sum over lattice cells (GCsp: flattened (mu,k) cells of every z-bin) or over
multipole bins (photo), so each rank evaluates ALL stencil points on its own slice -- the tables are a few
MB and are uploaded by every rank -- and the only exchange is the (Z x) npar x npar partial Fisher
(<= 4 x 48 x 48 doubles), all-reduced once per probe over NCCL/NVLink.  Quadrature weights
are separable and precomputed per node, so slices need no halo.
"""
from __future__ import annotations

from typing import Tuple

import numpy as np


def split_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced [begin, end) of rank `rank` out of `world` over n units."""
    base, rem = divmod(int(n), int(world))
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def is_distributed() -> bool:
    try:
        import torch.distributed as dist

        return dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1
    except Exception:
        return False


def rank_world() -> Tuple[int, int]:
    if not is_distributed():
        return 0, 1
    import torch.distributed as dist

    return dist.get_rank(), dist.get_world_size()


def shard(fm, rank: int, world: int):
    """Make `fm.compute()` evaluate only rank's slice and all-reduce the partial Fisher."""
    fm._shard = (int(rank), int(world))
    return fm


def allreduce_numpy(a: np.ndarray, device=None) -> np.ndarray:
    """Sum a host array over all ranks (NCCL: staged through CUDA device `device`, default LOCAL_RANK --
    the device of the cfp context, not torch's current device; gloo: CPU)."""
    if not is_distributed():
        return a
    import os

    import torch
    import torch.distributed as dist

    t = torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64))
    if dist.get_backend() == "nccl":
        dev = int(os.environ.get("LOCAL_RANK", "0")) if device is None else int(device)
        t = t.cuda(dev)
        dist.all_reduce(t)
        return t.cpu().numpy()
    dist.all_reduce(t)
    return t.numpy()


def allreduce_plan_result(plan, shape) -> bool:
    """Sum the Fisher result buffer of a cfp plan over all ranks IN PLACE on the device (one NCCL all-reduce over
    NVLink on the plan's own device buffer).  Returns False when there is nothing to do on the device
    (single process, or a gloo group: the caller then reduces the fetched host array)."""
    if not is_distributed():
        return False
    import torch
    import torch.distributed as dist

    if dist.get_backend() != "nccl":
        return False
    ctx = plan.ctx
    ctx.sync()  # the plan ran on the context's stream, the collective runs on torch's
    t = device_view(plan.fisher_device_ptr(), shape, ctx.device)
    with torch.cuda.device(ctx.device):
        dist.all_reduce(t)
        torch.cuda.current_stream().synchronize()
    return True


def barrier():
    if is_distributed():
        import torch.distributed as dist

        dist.barrier()


def allgather_numpy(a: np.ndarray, device=None) -> np.ndarray:
    """Concatenate equally shaped 1-D host arrays of all ranks in rank order (NCCL all_gather staged through the
    cfp device, or gloo on the host).  Used by the sharded likelihood batches."""
    if not is_distributed():
        return a
    import os

    import torch
    import torch.distributed as dist

    t = torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64))
    world = dist.get_world_size()
    if dist.get_backend() == "nccl":
        dev = int(os.environ.get("LOCAL_RANK", "0")) if device is None else int(device)
        t = t.cuda(dev)
        out = torch.empty(world * t.numel(), dtype=torch.float64, device=t.device)
        dist.all_gather_into_tensor(out, t)
        return out.cpu().numpy()
    parts = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(parts, t)
    return torch.cat(parts).numpy()


class _DevArray:
    def __init__(self, ptr: int, shape):
        self.__cuda_array_interface__ = {"shape": tuple(int(s) for s in shape), "typestr": "<f8",
                                         "data": (int(ptr), False), "version": 3, "strides": None}


def device_view(ptr: int, shape, device: int = 0):
    """Zero-copy torch view of a device buffer owned by a cfp plan (e.g. the Fisher result), so that
    torch.distributed can all-reduce it in place on the launch stream."""
    import torch

    return torch.as_tensor(_DevArray(ptr, shape), device=torch.device("cuda", device))
