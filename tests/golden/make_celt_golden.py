#!/usr/bin/env python
"""Regenerates tests/golden/celt_entropy.npz from the REFERENCE's libopus 1.3.1 (oracle/_ref/libref_opus.so): CELT frames and, for each, what
the reference's own decoder read off the range coder --
  final_range   OPUS_GET_FINAL_RANGE after opus_decode() of the frame wrapped as a code-0 packet (public API), and
  per-frame trace of ref_celt_entropy_trace (oracle/ref_celt_shim.c: the reference's internal functions driven in decode order): flags, post-filter
  parameters, tf / spread / trim / intensity / coded bands, per-band PVQ budgets and fine bits, and the band energies after the frame.
Three corpora: `enc` streams of the reference ENCODER (CELT-only, several bitrates / frame sizes / bandwidths / signals incl. silence, transients);
`gold` the frames of tests/golden/opus_packets.json (the transmitter's settings); `rnd` random bytes under every CELT configuration.
    python tests/golden/make_celt_golden.py
"""
import ctypes as C
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import celt_binding as cb  # noqa: E402

R = cb.ref()
R.ref_opus_encode_stream_celt.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int]
ENDS = [13, 17, 19, 21]


def signal(kind, n, ch, rng):
    t = np.arange(n) / 48000.0
    if kind == "silence":
        x = np.zeros((n, ch))
    elif kind == "tones":
        x = sum(a * np.sin(2 * np.pi * f * t + p)[:, None] * np.ones(ch) for a, f, p in [(6000, 440, 0), (3000, 1250, 1), (1500, 5100, 2), (800, 11000, 3)])
        if ch == 2:
            x[:, 1] = np.roll(x[:, 1], 37) * 0.7
    elif kind == "noise":
        x = rng.normal(0, 5000, size=(n, ch))
    elif kind == "clicks":   # transients on a quiet background
        x = rng.normal(0, 30, size=(n, ch))
        for p in rng.integers(0, n - 50, size=max(1, n // 3000)):
            x[p: p + 40] += rng.normal(0, 14000, size=(40, ch))
    else:                     # speech-like bursts
        env = (np.sin(2 * np.pi * 3.1 * t) > 0.2) * (0.5 + 0.5 * np.sin(2 * np.pi * 0.7 * t))
        x = (env * (4000 * np.sin(2 * np.pi * 180 * t) + 2500 * np.sin(2 * np.pi * 2300 * t + 1)))[:, None] * np.ones(ch) + rng.normal(0, 60, size=(n, ch))
    return np.clip(np.round(x), -32768, 32767).astype(np.int16)


def encode(pcm, n_frames, frame_samples, ch, bitrate, max_bw):
    out = np.zeros((n_frames, 1500), dtype=np.uint8)
    lens = np.zeros(n_frames, dtype=np.int32)
    assert R.ref_opus_encode_stream_celt(pcm.ctypes.data, n_frames, frame_samples, ch, bitrate, max_bw, out.ctypes.data, lens.ctypes.data, 1500) == n_frames
    return [bytes(out[i, : lens[i]]) for i in range(n_frames)]


def toc_fields(toc):
    cfgn = toc >> 3
    assert cfgn >= 16, "not CELT-only"
    return 1 + ((toc >> 2) & 1), cfgn & 3, ENDS[(cfgn - 16) >> 2]


streams = []   # (name, decoder channels, [(frame bytes, C, LM, end, toc)])
rng = np.random.default_rng(11)
for si, (kind, fs, ch, br, bw) in enumerate([("tones", 960, 2, 92000, 0), ("noise", 960, 2, 24000, 0), ("clicks", 960, 2, 64000, 0), ("speechy", 480, 1, 16000, 0),
                                             ("silence", 960, 2, 92000, 0), ("clicks", 240, 2, 48000, 0), ("tones", 120, 1, 32000, 0), ("noise", 960, 1, 8000, 0),
                                             ("speechy", 960, 2, 12000, 1103), ("tones", 960, 2, 40000, 1104), ("clicks", 480, 2, 20000, 1101),
                                             ("noise", 960, 2, 510000, 0), ("speechy", 960, 2, 6000, 0)]):
    n_frames = 24
    pcm = signal(kind, n_frames * fs, ch, rng)
    pk = encode(np.ascontiguousarray(pcm), n_frames, fs, ch, br, bw)
    frames = []
    for p in pk:
        c, lm, end = toc_fields(p[0])
        assert (p[0] & 3) == 0
        frames.append((p[1:], c, lm, end, p[0]))
    streams.append(("enc_%02d_%s_%d_%dch_%dbps" % (si, kind, fs, ch, br), ch, frames))
gold = json.load(open(os.path.join(HERE, "opus_packets.json")))
for st in gold["streams"]:
    frames = []
    for pk, pr in zip(st["packets"], st["parse"]):
        b = bytes.fromhex(pk)
        for fr, c, lm, end in cb.frames_of_packet(b, pr):
            frames.append((fr, c, lm, end, b[0]))
    streams.append(("gold_" + st["name"], st["channels"], frames))
rr = np.random.default_rng(12)
for k in range(40):
    frames = []
    for _ in range(30):
        cfgn, stereo = int(rr.integers(16, 32)), int(rr.integers(0, 2))
        ln = int(rr.choice([2, 3, 5, 9, 17, 33, 65, 129, 257, 700, 1275])) if rr.random() < 0.4 else int(rr.integers(2, 180))
        fr = rr.integers(0, 256, size=ln, dtype=np.uint8)
        if rr.random() < 0.2:
            fr[: min(3, ln)] = rr.integers(0, 8, size=min(3, ln))
        frames.append((fr.tobytes(), 1 + stereo, cfgn & 3, ENDS[(cfgn - 16) >> 2], (cfgn << 3) | (stereo << 2)))
    streams.append(("rnd_%02d" % k, 2, frames))

rows, blob, names, sbegin = [], bytearray(), [], [0]
for name, dec_ch, frames in streams:
    # public API: one decoder per stream, frames in order
    maxlen = max(len(f[0]) for f in frames)
    ranges = np.zeros(len(frames), dtype=np.uint32)
    d_err = C.c_int(0)
    R.opus_decoder_create.restype = C.c_void_p
    dec = R.opus_decoder_create(48000, dec_ch, C.byref(d_err))
    pcm = np.zeros(5760 * 2, dtype=np.int16)
    old_e = np.zeros(42, dtype=np.int16)
    for i, (fr, c, lm, end, toc) in enumerate(frames):
        pkt = bytes([toc & 0xFC]) + fr
        r = R.opus_decode(C.c_void_p(dec), pkt, len(pkt), pcm.ctypes.data_as(C.c_void_p), 5760, 0)
        assert r > 0, (name, i, r)
        fr_ = C.c_uint32(0)
        R.opus_decoder_ctl(C.c_void_p(dec), 4031, C.byref(fr_))    # OPUS_GET_FINAL_RANGE_REQUEST
        tr = cb.RefTrace()
        fb = np.frombuffer(fr, dtype=np.uint8)
        assert R.ref_celt_entropy_trace(fb.ctypes.data, len(fr), c, lm, end, old_e.ctypes.data, C.byref(tr)) == 0
        assert tr.rng[7] == fr_.value, (name, i)
        flags = (1 if tr.silence else 0) | (2 if tr.postfilter else 0) | (4 if tr.transient else 0) | (8 if tr.intra else 0) | (16 if tr.dual_stereo else 0) | \
                (32 if tr.anti_collapse_on else 0)
        rows.append((len(blob), len(fr), c, lm, end, fr_.value, tr.tell[7], flags, tr.pf_pitch, tr.pf_qg, tr.pf_tapset, tr.spread, tr.alloc_trim, tr.intensity,
                     tr.coded_bands, list(tr.tf_res), list(tr.fine_quant), list(tr.pulses), list(tr.band_e)))
        blob += fr
    R.opus_decoder_destroy(C.c_void_p(dec))
    names.append(name)
    sbegin.append(len(rows))
dt = np.dtype([("offset", "<u4"), ("len", "<u4"), ("channels", "u1"), ("lm", "u1"), ("end_band", "u1"), ("final_range", "<u4"), ("tell_bits", "<i4"), ("flags", "<u4"),
               ("pf_pitch", "<u2"), ("pf_gain_q", "u1"), ("pf_tapset", "u1"), ("spread", "u1"), ("alloc_trim", "u1"), ("intensity", "u1"), ("coded_bands", "u1"),
               ("tf_res", "i1", (21,)), ("fine_quant", "u1", (21,)), ("pulses", "<i2", (21,)), ("band_e", "<i2", (42,))])
arr = np.zeros(len(rows), dtype=dt)
for i, r in enumerate(rows):
    arr[i] = r
np.savez_compressed(os.path.join(HERE, "celt_entropy.npz"), frames=arr, bytes=np.frombuffer(bytes(blob), dtype=np.uint8), stream_begin=np.array(sbegin, dtype=np.uint32),
                    names=np.array(names))
fl = arr["flags"]
print(len(rows), "frames in", len(names), "streams;", len(blob), "bytes; silence", int((fl & 1 > 0).sum()), "postfilter", int((fl & 2 > 0).sum()), "transient", int((fl & 4 > 0).sum()),
      "intra", int((fl & 8 > 0).sum()), "dual", int((fl & 16 > 0).sum()), "anti-collapse", int((fl & 32 > 0).sum()), "coded bands", sorted(set(arr["coded_bands"].tolist())))
