"""CPU tests of the multi-GPU host logic with the gloo backend, world_size 2 (the N>1 path of bench.py / batch.py)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from dexdeform_b200.batch import allreduce_loss_and_grads, gather_scores, pack, partition_envs, unpack


def test_partition_covers_all_envs_once():
    for n in (1, 7, 64, 512, 513):
        for world in (1, 2, 4, 8):
            seen = []
            for r in range(world):
                s, c = partition_envs(n, world, r)
                seen += list(range(s, s + c))
            assert seen == list(range(n))
    with pytest.raises(ValueError):
        partition_envs(8, 2, 2)


def test_pack_unpack_roundtrip():
    g = [torch.arange(6.0).reshape(2, 3), torch.ones(4)]
    buf = pack(2.5, g)
    assert buf.shape == (11,) and buf.dtype == torch.float32
    loss, out = unpack(buf, [(2, 3), (4,)])
    assert float(loss) == 2.5 and torch.equal(out[0], g[0]) and torch.equal(out[1], g[1])


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_envs, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    start, count = partition_envs(n_envs, world, rank)
    # every environment e contributes loss e and gradient e * ones: the reduced result must equal the single-process sum
    loss = torch.tensor(float(sum(range(start, start + count))))
    grad = torch.ones(5, 26) * float(sum(range(start, start + count)))
    l, (g,) = allreduce_loss_and_grads(loss, [grad])
    scores = gather_scores(torch.arange(start, start + count, dtype=torch.float32), n_envs)
    q.put((rank, float(l), g.clone(), scores.clone()))
    dist.destroy_process_group()


def test_allreduce_matches_single_process_sum():
    world, n_envs = 2, 7
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_envs, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    total = float(sum(range(n_envs)))
    for rank, l, g, scores in results:
        assert l == total
        assert torch.equal(g, torch.ones(5, 26) * total)
        assert torch.equal(scores, torch.arange(n_envs, dtype=torch.float32))
