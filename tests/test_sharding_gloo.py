"""world_size-2 gloo test of the N > 1 host logic: event sharding + result gather."""
import os
import socket

import pytest
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_events, out):
    import torch.distributed as dist

    from acts_b200 import sharding

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = sharding.events_of_rank(n_events, rank, world)
    local = {e: 1000 + 7 * e for e in mine}       # stands for the seed count of event e
    counts = sharding.gather_seed_counts(local, n_events)
    out[rank] = (mine, counts)
    dist.destroy_process_group()


def test_event_sharding_and_gather_world2():
    from acts_b200 import sharding

    n_events, world = 11, 2
    assert sharding.events_of_rank(n_events, 0, 2) == [0, 2, 4, 6, 8, 10]
    assert sorted(sharding.events_of_rank(n_events, 0, 2) + sharding.events_of_rank(n_events, 1, 2)) == list(range(n_events))
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), n_events, out), nprocs=world, join=True)
    expected = [1000 + 7 * e for e in range(n_events)]
    for rank in range(world):
        mine, counts = out[rank]
        assert mine == list(range(rank, n_events, world))
        assert counts == expected


def test_gather_detects_missing_events():
    from acts_b200 import sharding

    with pytest.raises(RuntimeError):
        sharding.gather_seed_counts({0: 1, 2: 3}, 3)


def test_phi_sectors_partition_the_bins():
    from acts_b200 import sharding

    for n_phi, world in ((53, 8), (53, 2), (26, 4), (5, 8), (138, 3)):
        seen = []
        for rank in range(world):
            first, count = sharding.phi_sector_of_rank(n_phi, rank, world)
            seen.extend(range(first, first + count))
        assert seen == list(range(1, n_phi + 1))
