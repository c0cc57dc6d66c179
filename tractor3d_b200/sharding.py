"""Emitter-ownership sharding over the GPUs of one box.

Emitters are independent (no particle-particle interaction, no cross-emitter state on the device), so the
population is partitioned statically: emitter `e` is owned by rank `e mod world`.  Every rank drives its own
context, slabs and launches; there is no per-frame collective.  The only communication is `reduce_stats`,
once, after the run: time is reduced with MAX (the slowest GPU defines the frame), counters with SUM.
Works with any torch.distributed backend (NCCL on the GPUs, gloo in the CPU tests).
"""
import torch
import torch.distributed as dist


def owner(emitter_id, world):
    return emitter_id % world


def local_emitter_ids(n_emitters, rank, world):
    """Global ids of the emitters rank `rank` owns, in the order it creates and launches them."""
    return list(range(rank, n_emitters, world))


def shard_sizes(n_emitters, world):
    return [len(range(r, n_emitters, world)) for r in range(world)]


def reduce_stats(max_stats, sum_stats, device=None):
    """max_stats / sum_stats: dicts of numbers local to this rank.  Returns (max_dict, sum_dict) over all ranks.
    Without an initialised process group the inputs are returned unchanged (single GPU)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return dict(max_stats), dict(sum_stats)
    dev = device if device is not None else ("cuda" if dist.get_backend() == "nccl" else "cpu")
    mk, sk = sorted(max_stats), sorted(sum_stats)
    out_max, out_sum = {}, {}
    if mk:
        t = torch.tensor([float(max_stats[k]) for k in mk], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        out_max = {k: float(v) for k, v in zip(mk, t.tolist())}
    if sk:
        t = torch.tensor([int(sum_stats[k]) for k in sk], dtype=torch.int64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        out_sum = {k: int(v) for k, v in zip(sk, t.tolist())}
    return out_max, out_sum
