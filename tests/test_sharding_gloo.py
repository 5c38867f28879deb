"""N>1 host path on CPU: utterance sharding + the final token all-gather, world_size 2, gloo."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from b200asr.sharding import gather_tokens, group_by_length, ragged_batches, shard_indices


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, lengths, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = shard_indices(lengths, world, rank)
    toks = [[i * 10 + k for k in range(i % 4 + 1)] for i in mine]          # fake per-clip tokens
    out = gather_tokens(toks, mine, len(lengths), max_len=8)
    q.put((rank, mine, out))
    dist.destroy_process_group()


def test_shard_and_gather_world2():
    lengths = [128000, 64000, 128000, 32000, 128000, 64000, 16000]
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, lengths, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    expect = [[i * 10 + k for k in range(i % 4 + 1)] for i in range(len(lengths))]
    owned = sorted(i for _, mine, _ in results for i in mine)
    assert owned == list(range(len(lengths)))                              # a partition: nothing lost or duplicated
    for _, _, out in results:
        assert out == expect                                               # every rank sees every clip, in order


def test_shard_balance_and_grouping():
    lengths = [480000] * 3 + [128000] * 13
    shards = [shard_indices(lengths, 4, r) for r in range(4)]
    assert sorted(i for s in shards for i in s) == list(range(16))
    assert max(len(s) for s in shards) - min(len(s) for s in shards) <= 1
    loads = [sum(lengths[i] for i in s) for s in shards]
    assert max(loads) - min(loads) <= 480000
    groups = group_by_length(shards[0], lengths)
    assert all(len({lengths[i] for i in idx}) == 1 for _, idx in groups)


def test_gather_single_process():
    out = gather_tokens([[1, 2], [3]], [1, 0], 2, max_len=4)
    assert out == [[3], [1, 2]]


def test_ragged_batches_cover_every_clip_once_and_keep_neighbours_together():
    lengths = [16000, 128000, 8000, 64000, 64160, 127840, 9000]
    batches = ragged_batches(range(len(lengths)), lengths, 3)
    assert sorted(i for b in batches for i in b) == list(range(len(lengths)))
    assert all(len(b) <= 3 for b in batches)
    assert batches[0] == [1, 5, 4]                      # the three longest share a batch
    spans = [max(lengths[i] for i in b) - min(lengths[i] for i in b) for b in batches]
    assert spans[0] < 64000
