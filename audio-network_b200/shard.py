"""Channel sharding across ranks (one process per GPU) and the host-side gather of decoded frames.

Channels are independent: rank r owns a contiguous range of channel ids, runs its
own demodulator on its own GPU, and nothing crosses GPUs on the data path.  Only the decoded frames
(a few bytes per channel-second) are gathered on the host, ordered by (channel, start_sample)."""


def channel_range(rank, world, n_total):
    """Contiguous, balanced partition: first (n_total % world) ranks get one extra channel."""
    base, extra = divmod(n_total, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def to_global(frames, ch0):
    """Frame tuples (channel, start_sample, crc_ok, payload) with rank-local channel ids -> global ids."""
    return [(c + ch0, s, ok, p) for (c, s, ok, p) in frames]


def gather_frames(frames, group=None, dst=0):
    """Gathers every rank's (already global-id) frame list on `dst` over the process group's host
    transport and returns the merged, deterministically ordered list there (None elsewhere)."""
    import torch.distributed as dist

    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return sorted(frames, key=lambda f: (f[0], f[1]))
    rank = dist.get_rank(group)
    bucket = [None] * dist.get_world_size(group) if rank == dst else None
    dist.gather_object(frames, bucket, dst=dst, group=group)
    if rank != dst:
        return None
    merged = [f for part in bucket for f in part]
    merged.sort(key=lambda f: (f[0], f[1]))
    return merged
