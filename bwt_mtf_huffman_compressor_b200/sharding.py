"""File-batch sharding across GPUs: independent files, one BWT block each, greedy
largest-first assignment, no collective on the data path.  Only the timing reduction (max over
ranks) and the gathering of output sizes use torch.distributed."""


def shard_files(sizes, world_size):
    """Returns `world_size` lists of file indices.  Greedy: largest file first onto the least
    loaded rank (ties: lowest rank); every rank applies the same deterministic rule, so no
    communication is needed to agree on the assignment."""
    order = sorted(range(len(sizes)), key=lambda i: (-sizes[i], i))
    load = [0] * world_size
    out = [[] for _ in range(world_size)]
    for i in order:
        r = min(range(world_size), key=lambda k: (load[k], k))
        out[r].append(i)
        load[r] += sizes[i]
    return out


def max_over_ranks(value, dist=None, device="cpu"):
    """Timing rule of the bench contract: the job is as slow as its slowest rank."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return float(value)
    import torch
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t[0])


def gather_sizes(local_pairs, dist=None):
    """local_pairs: list of (file_index, compressed_size) of this rank -> dict over all ranks."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return dict(local_pairs)
    objs = [None] * dist.get_world_size()
    dist.all_gather_object(objs, list(local_pairs))
    out = {}
    for lst in objs:
        out.update(dict(lst))
    return out
