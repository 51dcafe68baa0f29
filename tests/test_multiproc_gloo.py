"""World-size-2 CPU (gloo) test of the multi-GPU host logic: contiguous batch shards, per-rank seeds, and the one
optional collective of the path (sampler.gather_outputs = all_gather of every output tensor along the batch)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from brepgen_b200.sampler import gather_outputs, shard_batch


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, global_batch, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = shard_batch(global_batch, rank, world)
    # each rank "samples" its shard: values identify (global sample index), masks depend on the index
    idx = torch.arange(lo, hi)
    out = {"surfZ": idx.float().view(-1, 1, 1).repeat(1, 3, 4), "surfMask": (idx % 2 == 0).view(-1, 1).repeat(1, 3)}
    g = gather_outputs(out)
    ok = (g["surfZ"][:, 0, 0].tolist() == list(range(global_batch))
          and g["surfMask"].dtype == torch.bool
          and g["surfMask"][:, 0].tolist() == [i % 2 == 0 for i in range(global_batch)])
    q.put((rank, ok, tuple(g["surfZ"].shape)))
    dist.destroy_process_group()


def test_shard_and_gather_world2():
    world, gb = 2, 8
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_worker, args=(r, world, port, gb, q)) for r in range(world)]
    for p in ps:
        p.start()
    res = [q.get(timeout=120) for _ in ps]
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok, _ in res), res
    assert all(shape == (gb, 3, 4) for _, _, shape in res)
