#!/usr/bin/env python
"""BASELINE.json configs[2] as written: 65,536 channels x 60 s at 44.1 kHz, AWGN at 10 dB SNR, streamed in 1-second chunks
from HOST memory, sharded over the GPUs of one box (one process per GPU), decoded frames gathered on rank 0.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 tools/stream_cfg3.py [--realtime-factor F]

Per rank: 65,536 / world channels.  Chunk k of every channel is synthesised (GPU transmitter stand-in, seeded per global
channel id), lands in pinned host memory, and is then fed like captured PCM: anm_demod_feed_host_async (H2D copy + kernel) ->
anm_demod_collect_upto(1) -> frames out of the handle's queue.  The generation of chunk k+1 overlaps the feed of chunk k.
With --realtime-factor F the submission of chunks is paced by the reference transmitter's leaky bucket (anm_pacer_*: 1200 ms of
receiver buffer draining F x 1000 ms per second, MulticastAudioOutput.kt:85) -- the flow control a live feed would see.

Checks on rank 0: no queue overflow anywhere; the digest of the gathered records equals the sum of the digests the ranks
computed locally; every frame of the sampled channels equals, byte for byte, what the CPU oracle decodes from PCM rendered by
the oracle's OWN transmitter (oracle/anm_oracle_tx.c) for the whole 60 s.  Prints one JSON line.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for _p in (ROOT, os.path.join(ROOT, "tests")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import bench  # noqa: E402  (workload constants and program builder)

TOTAL_CH = 65536
CHUNK_SYMS = 345          # 44,160 samples = 1.0014 s; 60 chunks = 60.08 s >= 60 s (ragged: 345 = 10 x 32 + 25 symbol periods)
N_CHUNKS = 60


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--channels-total", type=int, default=TOTAL_CH)
    ap.add_argument("--chunks", type=int, default=N_CHUNKS)
    ap.add_argument("--sample-channels", type=int, default=4, help="channels per rank checked against the oracle")
    ap.add_argument("--realtime-factor", type=float, default=0.0, help="pace the feed at F x real time (0 = as fast as possible)")
    ap.add_argument("--out", default="")
    args = ap.parse_args()

    import torch
    import torch.distributed as dist

    import audio_network_b200 as anm

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    numa = bench.bind_to_gpu_numa(torch, local)
    host_group = None
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
        host_group = dist.new_group(backend="gloo")

    cfg = anm.config_preset("ref4")
    n_ch = args.channels_total // world
    ch0 = rank * n_ch
    chunk = CHUNK_SYMS * cfg.sym_len
    stream = torch.cuda.current_stream().cuda_stream

    progs, lens, plist = bench.build_programs(cfg, anm.frame_symbols, anm.tx_params, n_ch, ch0)
    params = anm.tx_params_array(plist)
    d_prog = torch.from_numpy(progs).to(dev)
    d_len = torch.from_numpy(lens).to(dev)
    d_par = torch.from_numpy(params.view(np.uint8).copy()).to(dev)
    d_scratch = [torch.empty((n_ch, chunk), dtype=torch.int16, device=dev) for _ in range(2)]
    host = [torch.empty((n_ch, chunk), dtype=torch.int16).pin_memory() for _ in range(3)]
    gen_stream = torch.cuda.Stream()
    gen_done = [torch.cuda.Event() for _ in range(3)]

    def generate(k):
        """chunk k -> pinned host buffer k % 3 (on its own stream: overlaps the feed of chunk k - 1)"""
        with torch.cuda.stream(gen_stream):
            s = d_scratch[k % 2]
            anm.tx_render_device(cfg, d_prog.data_ptr(), progs.shape[1], d_len.data_ptr(), d_par.data_ptr(), n_ch, k * chunk,
                                 s.data_ptr(), chunk, chunk, gen_stream.cuda_stream)
            host[k % 3].copy_(s, non_blocking=True)
            gen_done[k % 3].record(gen_stream)

    dm = anm.Demod(cfg, n_ch, device=local)
    pacer = anm.Pacer(1200, int(1000 * args.realtime_factor), 0) if args.realtime_factor > 0 else None
    chunk_ms = int(round(1000.0 * chunk / 44100.0))
    recs_all, by_all = [], []
    n_frames = n_ok = n_bytes = 0
    waited_ns = 0

    def consume():
        nonlocal n_frames, n_ok, n_bytes
        r, b = dm.peek_frames()
        if len(r):
            ok, by = anm.frames_summary(r)
            n_frames += len(r)
            n_ok += ok
            n_bytes += by
            recs_all.append(r.copy())
            by_all.append(b.copy())
        dm.drop_frames()

    generate(0)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    t_feed = 0.0
    for k in range(args.chunks):
        if k + 1 < args.chunks:
            if k >= 2:
                dm.wait_input()                     # the H2D copy that read host[(k + 1) % 3] (chunk k - 2) is complete
            generate(k + 1)
        gen_done[k % 3].synchronize()               # chunk k is in host memory: from here on it is "captured PCM"
        if pacer is not None:
            now = int((time.perf_counter() - t0) * 1e9)
            while True:
                w = pacer.try_put(chunk_ms, now)
                if w is None:
                    break
                time.sleep(w * 1e-9)
                waited_ns += w
                now = int((time.perf_counter() - t0) * 1e9)
        f0 = time.perf_counter()
        dm.feed_host_async_ptr(host[k % 3].data_ptr(), chunk, chunk)
        dm.collect_upto(1)
        consume()
        t_feed += time.perf_counter() - f0
    dm.collect()
    consume()
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    overflow = dm.overflowed()
    stats = dm.stats()
    dm.close()

    # ---- host-side gather of the frame records on rank 0 (binary records over gloo) ----
    recs = np.concatenate(recs_all) if recs_all else np.zeros(0, dtype=anm.FRAME_DTYPE)
    offs = np.concatenate([[0], np.cumsum([len(b) for b in by_all])[:-1]]).astype(np.int64) if by_all else np.zeros(0, dtype=np.int64)
    if len(recs):
        recs["offset"] = (recs["offset"].astype(np.int64) + np.repeat(offs, [len(r) for r in recs_all])).astype(np.uint32)
        recs["channel"] += ch0
    by = np.concatenate(by_all) if by_all else np.zeros(0, dtype=np.uint8)
    local_dg = anm.frames_digest(recs, by)
    g0 = time.perf_counter()
    parts = [(recs, by)]
    meta = torch.tensor([recs.nbytes, len(by), local_dg & 0x7FFFFFFFFFFFFFFF, local_dg >> 63, n_frames, n_ok, n_bytes, int(overflow),
                         int(wall * 1e6), int(t_feed * 1e6), int(stats["frames_ok"].sum()), waited_ns // 1000], dtype=torch.int64)
    metas = [meta]
    if world > 1:
        metas = [torch.zeros_like(meta) for _ in range(world)]
        dist.all_gather(metas, meta, group=host_group)
        if rank == 0:
            for r in range(1, world):
                tr = torch.empty(int(metas[r][0]), dtype=torch.uint8)
                tb = torch.empty(int(metas[r][1]), dtype=torch.uint8)
                dist.recv(tr, src=r, group=host_group)
                dist.recv(tb, src=r, group=host_group)
                parts.append((tr.numpy().view(anm.FRAME_DTYPE), tb.numpy()))
        else:
            dist.send(torch.from_numpy(recs.view(np.uint8).reshape(-1)), dst=0, group=host_group)
            dist.send(torch.from_numpy(by), dst=0, group=host_group)
    gather_s = time.perf_counter() - g0

    if rank == 0:
        import oracle_binding as ob

        digest_ok = all(anm.frames_digest(pr, pb) == (int(m[2]) | (int(m[3]) << 63)) for (pr, pb), m in zip(parts, metas))
        # sampled channels against the oracle, whole 60 s, PCM from the oracle's own transmitter
        ocfg = ob.preset("ref4")
        n_total = args.chunks * chunk
        sample = sorted({r * n_ch + int(j * n_ch / args.sample_channels) for r in range(world) for j in range(args.sample_channels)})
        mismatches = 0
        checked_frames = 0
        for gch in sample:
            p_, l_, pl_ = bench.build_programs(ocfg, ob.frame_symbols, ob.tx_params, 1, gch)
            arr = np.zeros(1, dtype=anm.TXPARAMS_DTYPE)
            arr[0] = (pl_[0].seed, pl_[0].start_offset, pl_[0].amplitude_q15, pl_[0].snr_mdb, pl_[0].ppm_x1000, 0)
            pcm = ob.tx_render_batch(ocfg, p_, l_, arr, 0, n_total, 1)
            want = ob.oracle_frames_batch(ocfg, pcm)
            pr, pb = parts[gch // n_ch]
            sel = pr[pr["channel"] == gch]
            got = sorted(((0, int(r["start_sample"]), int(r["crc_ok"]), bytes(pb[int(r["offset"]): int(r["offset"]) + int(r["len"])])) for r in sel),
                         key=lambda f: f[1])
            checked_frames += len(want)
            if got != want:
                mismatches += 1
        tot_frames = sum(int(m[4]) for m in metas)
        tot_ok = sum(int(m[5]) for m in metas)
        tot_bytes = sum(int(m[6]) for m in metas)
        wall_max = max(int(m[8]) for m in metas) * 1e-6
        feed_max = max(int(m[9]) for m in metas) * 1e-6
        samples = args.channels_total * n_total
        line = {
            "what": "BASELINE.json configs[2] as written: %d channels x %d chunks of %d samples (%.2f s of stream each) from pinned host memory, 10 dB SNR, %d GPU(s)"
                    % (args.channels_total, args.chunks, chunk, n_total / 44100.0, world),
            "n_gpus": world, "channels_per_gpu": n_ch, "samples_total": samples, "host_bytes_streamed": samples * 2,
            "wall_s_max_over_ranks": round(wall_max, 3), "Msamples_per_s_wall": round(samples / wall_max / 1e6, 1),
            "feed_s_max_over_ranks": round(feed_max, 3), "Msamples_per_s_feed_only": round(samples / feed_max / 1e6, 1),
            "realtime_factor_achieved": round((n_total / 44100.0) / wall_max, 1),
            "frames": tot_frames, "frames_crc_ok": tot_ok, "payload_bytes_ok": tot_bytes, "decoded_bits_per_s_wall": round(tot_bytes * 8 / wall_max, 1),
            "stats_frames_ok_sum": sum(int(m[10]) for m in metas),
            "overflowed_any_rank": bool(any(int(m[7]) for m in metas)),
            "gather": {"frames_on_rank0": int(sum(len(pr) for pr, _ in parts)), "payload_bytes_on_rank0": int(sum(len(pb) for _, pb in parts)),
                       "seconds": round(gather_s, 3), "digest_ok": bool(digest_ok), "transport": "gloo send/recv of raw records" if world > 1 else "in-process"},
            "oracle_check": {"channels": len(sample), "frames": checked_frames, "channels_mismatching": mismatches,
                             "pcm_source": "oracle/anm_oracle_tx.c (independent of the GPU renderer)"},
            "pacer": None if pacer is None else {"realtime_factor_requested": args.realtime_factor, "capacity_ms": 1200, "chunk_ms": chunk_ms,
                                                 "waited_s_rank_max": round(max(int(m[11]) for m in metas) * 1e-6, 3)},
            "numa": numa,
        }
        print(json.dumps(line), flush=True)
        if args.out:
            with open(args.out, "w") as f:
                f.write(json.dumps(line) + "\n")
        assert not line["overflowed_any_rank"] and digest_ok and mismatches == 0
        assert line["stats_frames_ok_sum"] == tot_ok
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
