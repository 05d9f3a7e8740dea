"""Host-side logic of the multi-GPU run: the path shards by independent decoder streams - one
process per GPU, stream s on rank s mod world, no data-path collective.  The only communication is the timing
protocol of bench.py: a barrier on both sides of the timed region, the MAX of the ranks' device times, the SUM of
the units they processed.  Works on any torch.distributed backend (NCCL on the GPUs, gloo in the CPU tests)."""


def streams_of_rank(n_streams, rank, world):
    """Decoder streams owned by `rank`: s with s % world == rank (stream -> GPU s mod G)."""
    if world < 1 or not 0 <= rank < world:
        raise ValueError("rank %d outside world %d" % (rank, world))
    return list(range(rank, n_streams, world))


def seed_of_rank(base_seed, rank):
    """Every rank decodes different content (its own streams): synthetic inputs are seeded per rank."""
    return base_seed + rank


def aggregate(units_local, elapsed_ms_local, dist=None, device=None):
    """Whole-job throughput inputs: (units of all ranks, slowest rank's time in ms).  dist: an initialised
    torch.distributed module or None for a single process."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return float(units_local), float(elapsed_ms_local)
    import torch
    t = torch.tensor([float(elapsed_ms_local)], dtype=torch.float64, device=device)
    u = torch.tensor([float(units_local)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dist.all_reduce(u, op=dist.ReduceOp.SUM)
    return float(u.item()), float(t.item())


def throughput_mpix(units_total, elapsed_ms):
    return units_total / (elapsed_ms * 1e-3) / 1e6
