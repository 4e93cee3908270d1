"""world_size-2 tests of the N>1 host logic on CPU (gloo): game-id sharding is a partition,
timings reduce with MAX and counters with SUM, and the frozen-weight broadcast makes every rank's
network identical to rank 0's."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from elf_b200 import dist_utils
    from elf_b200.model import PolicyValueNet, broadcast_weights

    assert dist_utils.env_world() == (world, rank, rank)
    G = 64
    ids = set()
    for step in range(3):
        f = dist_utils.shard_first_game_id(step, world, rank, G)
        ids |= set(range(f, f + G))
    times, counts = dist_utils.reduce_timing_and_counts(dist, "cpu", [10.0 + rank, 5.0 - rank], [100 + rank, len(ids)])
    torch.manual_seed(100 + rank)  # different init per rank on purpose
    net = PolicyValueNet(9, num_block=1, dim=8)
    before = float(sum(p.double().sum() for p in net.parameters()))
    broadcast_weights(net, src=0)
    after = float(sum(p.double().sum() for p in net.parameters()))
    allsum = [None] * world
    dist.all_gather_object(allsum, (after, sorted(ids)[:3], len(ids)))
    q.put((rank, times, counts, before, after, allsum))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_reduction_and_weight_broadcast():
    world = 2
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    [p.start() for p in ps]
    out = sorted(q.get(timeout=120) for _ in range(world))
    [p.join(timeout=60) for p in ps]
    assert all(p.exitcode == 0 for p in ps)
    for rank, times, counts, before, after, allsum in out:
        assert times == [11.0, 5.0]            # MAX over ranks
        assert counts == [201, 2 * 3 * 64]     # SUM over ranks; every rank owns 3 x 64 distinct ids
        assert allsum[0][0] == allsum[1][0]    # identical weights after the broadcast
    assert out[0][3] != out[1][3]              # ... and they were different before
    # the two ranks' id windows are disjoint
    a0, a1 = out[0][5][0][1], out[0][5][1][1]
    assert set(a0).isdisjoint(a1)


def test_shard_windows_partition_the_id_space():
    from elf_b200.dist_utils import shard_first_game_id

    G, W = 4096, 8
    seen = np.zeros(3 * W * G, np.int32)
    for step in range(3):
        for r in range(W):
            f = shard_first_game_id(step, W, r, G)
            seen[f:f + G] += 1
    assert (seen == 1).all()
