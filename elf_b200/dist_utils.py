"""Host-side logic of the multi-GPU path (one process per GPU, games sharded, no data-path
collective).  Kept free of CUDA so it can be exercised with the gloo backend on CPU
(tests/test_dist_gloo.py)."""
import os


def env_world():
    return (int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")),
            int(os.environ.get("LOCAL_RANK", "0")))


def shard_first_game_id(step, world, rank, games_per_rank):
    """Global id of the first game rank `rank` plays in step `step`: every (step, rank) pair owns a
    disjoint window of `games_per_rank` consecutive ids (game g -> GPU by contiguous blocks)."""
    return (step * world + rank) * games_per_rank


def reduce_timing_and_counts(dist, device, times_ms, counts):
    """MAX over ranks of the timings, SUM over ranks of the counters (lists of numbers)."""
    import torch

    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return list(times_ms), list(counts)
    t = torch.tensor(list(times_ms), dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    c = torch.tensor(list(counts), dtype=torch.int64, device=device)
    dist.all_reduce(c, op=dist.ReduceOp.SUM)
    return t.tolist(), c.tolist()
