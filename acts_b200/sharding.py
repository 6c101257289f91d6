"""Event sharding across GPUs.

Events are independent (one WhiteBoard per event in the reference's Sequencer,
Examples/Framework/src/Framework/Sequencer.cpp:495-501), so an N-GPU job gives
event ``e`` to rank ``e mod N``; every rank runs whole events on its own device
and only the per-event results are gathered on the host.  No device collective
is on the data path.
"""
from __future__ import annotations


def events_of_rank(n_events: int, rank: int, world: int) -> list[int]:
    return list(range(rank, n_events, world))


def gather_seed_counts(local_counts: dict[int, int], n_events: int, group=None) -> list[int]:
    """All ranks contribute {event: nSeeds}; returns the per-event list on every rank
    (torch.distributed object gather over the host, gloo or nccl backend)."""
    import torch.distributed as dist

    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        merged = dict(local_counts)
    else:
        parts = [None] * dist.get_world_size(group)
        dist.all_gather_object(parts, local_counts, group=group)
        merged = {}
        for p in parts:
            merged.update(p)
    missing = [e for e in range(n_events) if e not in merged]
    if missing:
        raise RuntimeError(f"events without a result: {missing[:8]}")
    return [merged[e] for e in range(n_events)]


def phi_sector_of_rank(n_phi_bins: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous block of middle phi bins (1-based first bin, count) of one rank when a
    single event is split over `world` GPUs (latency mode, BASELINE.json configs[4]).

    With more ranks than phi bins the surplus ranks get count 0: they must seed NOTHING
    (``SeedingEngine.set_phi_sector(empty=True)`` / ``apply_phi_sector`` below) -- a plain
    ``set_phi_sector(first, 0)`` means "all bins" and would duplicate every seed."""
    base, extra = divmod(n_phi_bins, world)
    first = 1 + rank * base + min(rank, extra)
    count = base + (1 if rank < extra else 0)
    return first, count


def apply_phi_sector(engine, n_phi_bins: int, rank: int, world: int) -> tuple[int, int]:
    """Restrict ``engine`` to the sector of ``rank``; ranks without bins are set to the empty sector."""
    first, count = phi_sector_of_rank(n_phi_bins, rank, world)
    if count == 0:
        engine.set_phi_sector(empty=True)
    else:
        engine.set_phi_sector(first, count)
    return first, count
