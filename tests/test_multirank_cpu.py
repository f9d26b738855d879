"""World-size-2 gloo test of the N>1 host logic on CPU: stripe partition, barrier,
max-over-ranks timing reduction, optional gather.  The per-stripe transform is the ORACLE
here (test stand-in for the CUDA kernels, which need a GPU); what is under test is that
striping by block-rows with no halo and no data-path collective reproduces the
single-process result."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, H, W, q):
    os.environ.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world),
                      MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    sys.path.insert(0, ROOT)
    import cuda_dct_idct_b200 as m
    from oracle import oracle as o

    r, lr, ws = m.dist.init("gloo")
    assert (r, ws) == (rank, world)
    img = o.rand_image(H, W, 42)                     # every rank can materialise the image
    r0, r1 = m.dist.my_stripe(H)
    mine = o.roundtrip(np.ascontiguousarray(img[r0:r1])) if r1 > r0 else np.zeros((0, W), np.float32)
    m.dist.barrier()
    slowest = m.dist.max_over_ranks(10.0 + rank)      # rank-dependent "time"
    total = m.dist.sum_over_ranks(float((r1 - r0) * W))
    per_rank = m.dist.all_ranks(10.0 + rank)          # bench.py's per-rank timing diagnostics
    assert per_rank == [10.0 + i for i in range(world)]
    # bench.py's parity_all_ranks: every rank checks its own stripe, the failure flags are summed
    ok = np.array_equal(mine.view(np.uint32), o.roundtrip(img)[r0:r1].view(np.uint32))
    assert m.dist.sum_over_ranks(0.0 if ok else 1.0) == 0.0
    assert m.dist.sum_over_ranks(1.0 if rank == 1 else 0.0) == 1.0   # one failing rank is seen by all
    full = m.dist.gather_stripes(torch.from_numpy(mine), H, dst=0)
    if rank == 0:
        q.put((slowest, total, full.numpy()))
    m.dist.barrier()
    torch.distributed.destroy_process_group()


@pytest.mark.parametrize("H", [64, 72, 8])
def test_two_rank_stripes_match_single_process(oracle, H):
    W, world = 96, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, H, W, q)) for r in range(world)]
    for p in procs:
        p.start()
    slowest, total, full = q.get(timeout=120)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert slowest == 11.0 and total == H * W
    want = oracle.roundtrip(oracle.rand_image(H, W, 42))
    assert np.array_equal(full.view(np.uint32), want.view(np.uint32))
