#!/usr/bin/env python
"""Throughput of the batched CELT decoder (row f1) on real encoder output: the golden stream of the transmitter's settings (50 stereo 20 ms frames, about 366
bytes each) replicated to --streams streams, decoded by anm_celt_entropy_device / anm_celt_spectrum_device / anm_celt_decode_device; CUDA events around
each call, digest of one replica's PCM checked against the committed golden.  Prints one JSON line.
    python tools/celt_bench.py --streams 4096"""
import argparse
import ctypes
import hashlib
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import audio_network_b200 as anm  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--streams", type=int, default=4096)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--name", default="gold_stereo_20ms")
    ap.add_argument("--in-flight", type=int, default=0, help="also time the full decode with the streams split over K calls on K CUDA streams (K contexts)")
    ap.add_argument("--cpu-baseline", action="store_true", help="also time the REFERENCE decoder (oracle/_ref/libref_opus.so, one core) on the same stream")
    a = ap.parse_args()
    G = np.load(os.path.join(ROOT, "tests", "golden", "celt_entropy.npz"))
    P = np.load(os.path.join(ROOT, "tests", "golden", "celt_pcm.npz"))
    names = [str(n) for n in G["names"]]
    s = names.index(a.name)
    lo, hi = int(G["stream_begin"][s]), int(G["stream_begin"][s + 1])
    fr = G["frames"][lo:hi]
    nf = hi - lo
    jobs1 = np.zeros(nf, dtype=anm.CELT_JOB_DTYPE)
    for k in ("offset", "len", "channels", "lm", "end_band"):
        jobs1[k] = fr[k]
    jobs = np.tile(jobs1, a.streams)
    sb = (np.arange(a.streams + 1) * nf).astype(np.uint32)
    cc = int(fr["channels"].max())
    dev = torch.device("cuda:0")
    L = anm.lib()
    ctx = ctypes.c_void_p()
    assert L.anm_celt_ctx_create(0, ctypes.byref(ctx)) == 0
    d_by = torch.from_numpy(np.ascontiguousarray(G["bytes"])).to(dev)
    d_jobs = torch.from_numpy(jobs.view(np.uint8).reshape(-1).copy()).to(dev)
    d_sb = torch.from_numpy(sb.view(np.uint8).copy()).to(dev)
    d_st = torch.zeros(a.streams * anm.CELT_STREAM_DTYPE.itemsize, dtype=torch.uint8, device=dev)
    d_sy = torch.zeros(a.streams * anm.CELT_SYNTH_DTYPE.itemsize, dtype=torch.uint8, device=dev)
    d_fr = torch.zeros(len(jobs) * anm.CELT_FRAME_DTYPE.itemsize, dtype=torch.uint8, device=dev)
    d_x = torch.zeros(len(jobs) * 1920, dtype=torch.int16, device=dev)
    d_pcm = torch.zeros(len(jobs) * 1920, dtype=torch.int16, device=dev)
    stream = torch.cuda.current_stream().cuda_stream
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]

    def timed(fn):
        t = []
        for r in range(a.reps + 1):
            d_st.zero_()
            d_sy.zero_()
            ev[0].record()
            assert fn() == 0
            ev[1].record()
            torch.cuda.synchronize()
            if r:
                t.append(ev[0].elapsed_time(ev[1]))
        return float(np.median(t))
    args = (ctx, d_jobs.data_ptr(), d_sb.data_ptr(), a.streams, len(jobs), d_by.data_ptr(), 0xFFFFFFFF, d_st.data_ptr())
    t_ent = timed(lambda: L.anm_celt_entropy_device(*args, d_fr.data_ptr(), stream))
    t_spec = timed(lambda: L.anm_celt_spectrum_device(*args, d_fr.data_ptr(), d_x.data_ptr(), 1920, None, stream))
    t_dec = timed(lambda: L.anm_celt_decode_device(*args, d_sy.data_ptr(), d_fr.data_ptr(), d_pcm.data_ptr(), 1920, stream))
    t_split = None
    if a.in_flight > 1:
        # the latency-bound kernels of one call (per-stream energies, de-emphasis: a thread per stream / channel) leave most of the GPU idle; with several
        # calls in flight on streams of their own, one call's thin kernels run beside another's wide ones.  A context serves one call at a time: K contexts.
        K = a.in_flight
        cuts = [a.streams * k // K for k in range(K + 1)]
        ctxs, strs, d_sbs = [], [], []
        for k in range(K):
            cx = ctypes.c_void_p()
            assert L.anm_celt_ctx_create(0, ctypes.byref(cx)) == 0
            ctxs.append(cx)
            strs.append(torch.cuda.Stream())
            d_sbs.append(torch.from_numpy(((np.arange(cuts[k + 1] - cuts[k] + 1)) * nf).astype(np.uint32).view(np.uint8).copy()).to(dev))
        jsz, ssz, ysz, fsz = anm.CELT_JOB_DTYPE.itemsize, anm.CELT_STREAM_DTYPE.itemsize, anm.CELT_SYNTH_DTYPE.itemsize, anm.CELT_FRAME_DTYPE.itemsize
        d_pcm.zero_()

        def split_call():
            cur = torch.cuda.current_stream()
            for k in range(K):
                s0, n_s = cuts[k], cuts[k + 1] - cuts[k]
                j0, n_j = s0 * nf, n_s * nf
                strs[k].wait_stream(cur)
                rc = L.anm_celt_decode_device(ctxs[k], d_jobs.data_ptr() + j0 * jsz, d_sbs[k].data_ptr(), n_s, n_j, d_by.data_ptr(), 0xFFFFFFFF,
                                              d_st.data_ptr() + s0 * ssz, d_sy.data_ptr() + s0 * ysz, d_fr.data_ptr() + j0 * fsz, d_pcm.data_ptr() + j0 * 1920 * 2, 1920,
                                              strs[k].cuda_stream)
                assert rc == 0
            for k in range(K):
                cur.wait_stream(strs[k])
            return 0
        t_split = timed(split_call)
        for cx in ctxs:
            L.anm_celt_ctx_destroy(cx)
    pcm = d_pcm.view(len(jobs), 1920)
    ok = True
    for rep in (0, a.streams // 2, a.streams - 1):
        got = pcm[rep * nf: (rep + 1) * nf].cpu().numpy()
        for k in range(nf):
            ns = (120 << int(fr["lm"][k])) * cc
            d = np.frombuffer(hashlib.sha256(np.ascontiguousarray(got[k, :ns]).tobytes()).digest()[:8], dtype="<u8")[0]
            ok = ok and d == P["cc_native"][lo + k]
    L.anm_celt_ctx_destroy(ctx)
    cpu = None
    if a.cpu_baseline:
        # the reference's own celt_decode_with_ec() (a copy of celt/celt_decoder.c compiled in place, oracle/ref_celt_state_shim.c), one host core
        import time
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import celt_spectrum_binding as sbind
        R = sbind.ref()
        maxlen = int(fr["len"].max())
        buf, lens = np.zeros((nf, maxlen), np.uint8), fr["len"].astype(np.int32)
        for k in range(nf):
            buf[k, :lens[k]] = G["bytes"][fr["offset"][k]: fr["offset"][k] + lens[k]]
        params = np.ascontiguousarray(np.stack([fr["channels"], fr["lm"], fr["end_band"]], axis=1).astype(np.int32))
        states, out = np.zeros(nf, sbind.STATE), np.zeros((nf, 960 * cc), np.int16)
        reps, t0 = 0, time.perf_counter()
        while time.perf_counter() - t0 < 3.0:
            assert R.ref_celt_stream_states(buf.ctypes.data, lens.ctypes.data, nf, maxlen, params.ctypes.data, cc, states.ctypes.data, out.ctypes.data) == nf
            reps += 1
        dt = time.perf_counter() - t0
        cpu = {"kind": "reference", "cores": 1, "frames_per_s": round(reps * nf / dt, 1), "sample": "%d x the %d-frame stream in %.1f s" % (reps, nf, dt)}
    nfr = len(jobs)
    audio_s = nfr * (120 << int(fr["lm"][0])) / 48000.0
    pk_bytes = int(jobs["len"].astype(np.int64).sum())
    out_bytes = nfr * (120 << int(fr["lm"][0])) * cc * 2
    print(json.dumps({"tool": "celt_bench", "stream": a.name, "streams": a.streams, "frames": nfr, "packet_bytes": pk_bytes, "pcm_bytes": out_bytes,
                      "entropy_ms": round(t_ent, 3), "entropy_plus_spectrum_ms": round(t_spec, 3), "full_decode_ms": round(t_dec, 3),
                      "Mframes_per_s_full": round(nfr / t_dec / 1e3, 3), "audio_seconds_per_second": round(audio_s / (t_dec * 1e-3), 1),
                      "algorithmic_GBps_full": round((pk_bytes + out_bytes) / (t_dec * 1e-3) / 1e9, 2), "pcm_digests_equal_reference": bool(ok), "cpu_baseline": cpu,
                      "in_flight": None if t_split is None else {"calls": a.in_flight, "full_decode_ms": round(t_split, 3), "Mframes_per_s_full": round(nfr / t_split / 1e3, 3)},
                      "kernels": "k_celt_entropy, k_celt_energies, k_celt_spectrum, k_celt_blocks, k_celt_overlap, k_celt_deemphasis"}))


if __name__ == "__main__":
    main()
