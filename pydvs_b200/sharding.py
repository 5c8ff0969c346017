"""Multi-GPU sharding of independent camera streams.

Streams never exchange pixels -- a stream's reference / threshold state depends only on its own frames
(``mnist_to_spikes.py:241`` resets it per image; every reference caller owns one ``ref`` per camera) -- so
the path shards by contiguous blocks of streams, one process per GPU, with no data-path collective.  The
only exchange is one all-gather of per-stream event counts (8 bytes per stream) at the end of a run, used
for reporting and for the global CSR offsets of a concatenated event stream.
"""
import torch
import torch.distributed as dist


def shard_streams(total_streams, world_size, rank):
    """Contiguous block of streams owned by ``rank``: returns (first_stream, n_streams).
    The first ``total % world`` ranks own one extra stream."""
    if not (0 <= rank < world_size):
        raise ValueError("rank %d outside world of %d" % (rank, world_size))
    base, extra = divmod(int(total_streams), int(world_size))
    first = rank * base + min(rank, extra)
    return first, base + (1 if rank < extra else 0)


def owner_of_stream(stream, total_streams, world_size):
    base, extra = divmod(int(total_streams), int(world_size))
    boundary = extra * (base + 1)
    if stream < boundary:
        return stream // (base + 1)
    return extra + (stream - boundary) // base if base else world_size - 1


def gather_event_counts(local_counts, total_streams=None, group=None):
    """All-gather the per-stream event counts of every rank (int64 [S_local] -> int64 [S_total] on
    every rank, in global stream order).  NCCL when the tensor is on a GPU, gloo on the CPU."""
    local_counts = local_counts.to(torch.int64).contiguous()
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return local_counts.clone()
    world = dist.get_world_size(group)
    n_local = torch.tensor([local_counts.numel()], dtype=torch.int64, device=local_counts.device)
    sizes = [torch.zeros_like(n_local) for _ in range(world)]
    dist.all_gather(sizes, n_local, group=group)
    sizes = [int(s.item()) for s in sizes]
    width = max(sizes) if sizes else 0
    padded = torch.zeros(width, dtype=torch.int64, device=local_counts.device)
    padded[:local_counts.numel()] = local_counts
    parts = [torch.empty_like(padded) for _ in range(world)]
    dist.all_gather(parts, padded, group=group)
    out = torch.cat([p[:n] for p, n in zip(parts, sizes)])
    if total_streams is not None and out.numel() != total_streams:
        raise RuntimeError("gathered %d stream counts, expected %d" % (out.numel(), total_streams))
    return out


def global_event_offsets(all_counts):
    """Exclusive prefix of the gathered counts: where each stream's events start in a concatenated,
    stream-major global event list (int64 [S_total + 1])."""
    off = torch.zeros(all_counts.numel() + 1, dtype=torch.int64, device=all_counts.device)
    torch.cumsum(all_counts, 0, out=off[1:])
    return off
