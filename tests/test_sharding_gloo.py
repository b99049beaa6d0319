"""N>1 host path on CPU: world_size-2 gloo run of the batch sharding + result gather used by bench.py."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from helmnet_b200.sharding import gather_results, shard_sizes, shard_slice


def test_shard_slice_covers_batch():
    for batch, world in [(256, 8), (10, 4), (3, 2), (7, 8)]:
        seen = []
        for r in range(world):
            lo, hi = shard_slice(batch, world, r)
            seen += list(range(lo, hi))
        assert seen == list(range(batch))


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    batch, n, k = 5, 8, 3
    lo, hi = shard_slice(batch, world, rank)
    full_wf = torch.arange(batch * 2 * n * n, dtype=torch.float32).reshape(batch, 2, n, n)
    full_rm = torch.arange(k * batch, dtype=torch.float32).reshape(k, batch)
    out = gather_results(full_wf[lo:hi].clone(), full_rm[:, lo:hi].clone(), dst=0)
    # known slice sizes (no size exchange), and the equal-slice fast path (batch 4 over 2 ranks)
    out2 = gather_results(full_wf[lo:hi].clone(), full_rm[:, lo:hi].clone(), dst=0, sizes=shard_sizes(batch, world))
    lo4, hi4 = shard_slice(4, world, rank)
    out3 = gather_results(full_wf[lo4:hi4].clone(), full_rm[:, lo4:hi4].clone(), dst=0, sizes=shard_sizes(4, world))
    if rank == 0:
        ok = torch.equal(out[0], full_wf) and torch.equal(out2[0], full_wf) and torch.equal(out3[0], full_wf[:4])
        q.put((ok, torch.equal(out[1], full_rm) and torch.equal(out2[1], full_rm) and torch.equal(out3[1], full_rm[:, :4])))
    else:
        assert out is None and out2 is None and out3 is None
    dist.destroy_process_group()


def test_gather_results_world2():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    ok = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert ok == (True, True)
