"""The N > 1 path on CPU: two gloo ranks shard a mixed emitter population by ownership, each advances its own
shard (the oracle stands in for the GPU here -- this test is about the host-side partitioning and the end-of-run
reduction), and the reduced counters must equal a single-process run over the whole population."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import cpu_harness as H
import scenarios as S
from tractor3d_b200 import presets, sharding

N_EMITTERS, FRAMES = 11, 40


def _make(lib, eid):
    p, sprite = presets.MIXED[eid % 3]()
    p.particle_count_max = 300 + 17 * eid
    p.emission_rate = 150 + 25 * eid
    e = lib.emitter(p)
    e.set_sprite_frame_coords(*sprite)
    lib.emitter_set_device_rng(e.h, 1, 9000 + eid)   # keyed by the GLOBAL emitter id: results do not depend on the sharding
    e.start()
    return e


def _run_shard(ids):
    lib = H.oracle(fresh=True)
    ems = [_make(lib, eid) for eid in ids]
    updates = 0
    for f in range(FRAMES):
        for e in ems:
            e.update(S.DT)
            e.draw()
            updates += e.count()
    return {eid: e.count() for eid, e in zip(ids, ems)}, updates


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ids = sharding.local_emitter_ids(N_EMITTERS, rank, world)
    counts, updates = _run_shard(ids)
    mx, sm = sharding.reduce_stats({"ms": 10.0 + rank}, {"updates": updates, "live": sum(counts.values()), "emitters": len(ids)})
    q.put((rank, ids, counts, mx, sm))
    dist.barrier()
    dist.destroy_process_group()


def test_ownership_partition_is_exact():
    for world in (1, 2, 3, 4, 8):
        for n in (0, 1, 7, 8, 8192):
            shards = [sharding.local_emitter_ids(n, r, world) for r in range(world)]
            flat = sorted(i for s in shards for i in s)
            assert flat == list(range(n))
            assert [len(s) for s in shards] == sharding.shard_sizes(n, world)
            assert all(sharding.owner(i, world) == r for r, s in enumerate(shards) for i in s)
            assert max(map(len, shards)) - min(map(len, shards)) <= 1


def test_two_gloo_ranks_equal_one_process():
    whole_counts, whole_updates = _run_shard(list(range(N_EMITTERS)))
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    outs = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    merged = {}
    for rank, ids, counts, mx, sm in outs:
        assert ids == sharding.local_emitter_ids(N_EMITTERS, rank, 2)
        merged.update(counts)
        assert mx == {"ms": 11.0}                                   # MAX over ranks
        assert sm == {"updates": whole_updates, "live": sum(whole_counts.values()), "emitters": N_EMITTERS}
    assert merged == whole_counts


def test_reduce_stats_without_a_group_is_identity():
    assert sharding.reduce_stats({"a": 1.5}, {"b": 2}) == ({"a": 1.5}, {"b": 2})
