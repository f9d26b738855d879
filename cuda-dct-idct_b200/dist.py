"""Multi-GPU plumbing: one process per GPU (torchrun), block-row stripes, no collective on
the data path.  torch.distributed is used only for the barrier, the max-over-ranks timing
reduction and the OPTIONAL final gather of the stripes (NCCL over NVLink on GPUs, gloo on
CPU for the tests)."""
from __future__ import annotations

import os

from .stripes import stripe_rows


def env_rank():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))


def init(backend: str | None = None):
    """Initialise torch.distributed from the torchrun environment (no-op for 1 rank).
    Returns (rank, local_rank, world_size)."""
    import torch
    import torch.distributed as dist

    rank, local_rank, world = env_rank()
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29531")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local_rank)
            dist.init_process_group(backend, device_id=torch.device("cuda", local_rank))
        else:
            dist.init_process_group(backend)
    return rank, local_rank, world


def shutdown():
    import torch.distributed as dist

    if dist.is_available() and dist.is_initialized():
        dist.destroy_process_group()


def barrier():
    import torch.distributed as dist

    if dist.is_available() and dist.is_initialized():
        dist.barrier()


def max_over_ranks(value: float, device=None) -> float:
    """MAX all-reduce of a scalar (the slowest rank defines the step time)."""
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()):
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device or ("cuda" if dist.get_backend() == "nccl" else "cpu"))
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value: float, device=None) -> float:
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()):
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device or ("cuda" if dist.get_backend() == "nccl" else "cpu"))
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def all_ranks(value: float, device=None) -> list:
    """The scalar of every rank, in rank order (per-rank timing diagnostics)."""
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()):
        return [float(value)]
    dev = device or ("cuda" if dist.get_backend() == "nccl" else "cpu")
    t = torch.tensor([float(value)], dtype=torch.float64, device=dev)
    out = [torch.empty_like(t) for _ in range(dist.get_world_size())]
    dist.all_gather(out, t)
    return [float(x.item()) for x in out]


def my_stripe(H: int):
    rank, _, world = env_rank()
    return stripe_rows(H, world, rank)


def gather_stripes(stripe, H: int, dst: int = 0):
    """Optional final gather of the per-rank block-row stripes of an H-row image onto rank
    `dst` (reported separately from kernel throughput).  Stripes may differ by one
    block-row, so this is an all_gather of padded stripes followed by trimming."""
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()):
        return stripe
    world, rank = dist.get_world_size(), dist.get_rank()
    spans = [stripe_rows(H, world, r) for r in range(world)]
    max_rows = max(b - a for a, b in spans)
    pad = torch.zeros((max_rows, stripe.shape[1]), dtype=stripe.dtype, device=stripe.device)
    pad[: stripe.shape[0]] = stripe
    parts = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad)
    if rank != dst:
        return None
    return torch.cat([p[: b - a] for p, (a, b) in zip(parts, spans)], 0)


class PeerImage:
    """A full-size image in symmetric (peer-mapped) memory, for the FUSED transform+gather.

    Every rank allocates the same H x W buffer and exchanges handles once
    (torch.distributed._symmetric_memory: CUDA VMM allocations mapped into every rank's
    address space over NVLink/NVSwitch).  `stripe_on(dst, r0, r1)` is then an ordinary CUDA
    tensor view of rows [r0, r1) of rank `dst`'s buffer: handing it to the kernels as the
    OUTPUT plane makes each rank's transform write its finished stripe straight into the
    destination GPU's HBM with plain peer stores -- the gather rides on the transform's own
    stores, tile by tile, instead of running as a separate NCCL all-gather afterwards.
    """

    def __init__(self, H: int, W: int, dtype, device):
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm_mem

        self.H, self.W, self.dtype = H, W, dtype
        self.buf = symm_mem.empty((H, W), dtype=dtype, device=device)
        self.hdl = symm_mem.rendezvous(self.buf, dist.group.WORLD)

    def stripe_on(self, dst_rank: int, r0: int, r1: int):
        return self.hdl.get_buffer(dst_rank, (r1 - r0, self.W), self.dtype, r0 * self.W)

    def local(self):
        return self.buf

    def barrier(self):
        """Stream-ordered barrier over all ranks (signal pads): after it every peer's stores
        issued before its own barrier() are visible."""
        self.hdl.barrier()
