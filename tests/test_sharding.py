"""world_size-2 gloo run of the multi-rank host logic: channel partition + host-side frame gather.
The per-rank demodulation is stood in for by the CPU oracle (test infrastructure); the point here
is that sharded + gathered output equals the single-process output."""
import os
import sys

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

import audio_network_b200 as anm
from audio_network_b200 import shard
from oracle_binding import oracle_frames_batch
from sigutil import make_channels


def test_channel_range_partition():
    for n, w in ((8192, 8), (65536, 8), (10, 3), (5, 8), (1, 1)):
        seen = []
        for r in range(w):
            lo, hi = shard.channel_range(r, w, n)
            seen += list(range(lo, hi))
        assert seen == list(range(n))
        sizes = [shard.channel_range(r, w, n)[1] - shard.channel_range(r, w, n)[0] for r in range(w)]
        assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    cfg = anm.config_preset("ref4")
    pcm, _ = make_channels(cfg, 6, 300 * cfg.sym_len, seed=41, snr_db=10.0, offset_max=400)
    lo, hi = shard.channel_range(rank, world, pcm.shape[0])
    local = oracle_frames_batch(cfg, pcm[lo:hi])
    merged = shard.gather_frames(shard.to_global(local, lo))
    if rank == 0:
        q.put(merged)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gather_equals_single_process():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    merged = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    cfg = anm.config_preset("ref4")
    pcm, _ = make_channels(cfg, 6, 300 * cfg.sym_len, seed=41, snr_db=10.0, offset_max=400)
    assert merged == oracle_frames_batch(cfg, pcm)
    assert len(merged) > 0
