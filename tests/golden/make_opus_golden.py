#!/usr/bin/env python
"""Regenerates tests/golden/opus_packets.json and opus_synthetic.npz from the REFERENCE's libopus 1.3.1 (oracle/_ref/libref_opus.so, compiled in
place by `make -C oracle ref_opus`):
  streams   packets of the reference's own encoder with the transmitter's settings (OpusEncoder.kt:51-67) for seeded
            test signals, the reference's parse of each packet, and the SHA-256 of the PCM the reference's decoder
            produces for the stream (opus_decode, playback.cpp:115-122) -- the known answer a batched frame decoder
            (row f1 proper) will have to reproduce bit for bit;
  synthetic every TOC byte under every framing code, size / padding / limit edge cases, truncations, mutations, random
            bytes, each with the reference's parse (opus_packet_parse + the opus_packet_get_* family).
    python tests/golden/make_opus_golden.py
"""
import hashlib
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import opus_corpus as oc  # noqa: E402

R = oc.ref_lib()
import numpy as np  # noqa: E402

import audio_network_b200 as anm  # noqa: E402

out = {"version": R.ref_opus_version().decode(), "streams": []}
for name, n_frames, frame_samples, channels, seed in (("stereo_20ms", 50, 960, 2, 1), ("stereo_60ms", 20, 2880, 2, 2), ("mono_20ms", 25, 960, 1, 3),
                                                       ("stereo_10ms", 20, 480, 2, 4), ("stereo_2_5ms", 20, 120, 2, 5)):
    pcm, packets = oc.encode_stream(R, n_frames, frame_samples, channels, seed)
    dec = oc.decode_stream(R, packets, channels, frame_samples)
    out["streams"].append({"name": name, "frame_samples": frame_samples, "channels": channels, "seed": seed,
                           "packets": [p.hex() for p in packets], "parse": [oc.ref_parse(R, p) for p in packets],
                           "input_sha256": hashlib.sha256(pcm.tobytes()).hexdigest(), "decoded_sha256": hashlib.sha256(dec.tobytes()).hexdigest(),
                           "decoded_head": dec[:16].reshape(-1).tolist()})
    print(name, "packet bytes", sum(len(p) for p in packets), "modes", sorted({q["mode"] for q in out["streams"][-1]["parse"]}))
with open(os.path.join(HERE, "opus_packets.json"), "w") as f:
    json.dump(out, f, separators=(",", ":"))
pk, fss, refs = [], [], []
for i, p in enumerate(oc.synthetic_corpus()):
    for fs in ((48000,) if i % 5 else (48000, 16000, 8000)):
        pk.append(p)
        fss.append(fs)
        refs.append(oc.ref_parse(R, p, fs))
ref = np.zeros(len(pk), dtype=anm.OPUS_PACKET_DTYPE)
for i, r in enumerate(refs):
    for k in oc.FIELDS:
        ref[k][i] = r[k]
    ref["size"][i] = r["size"]
off = np.zeros(len(pk) + 1, dtype=np.int64)
off[1:] = np.cumsum([len(p) for p in pk])
np.savez_compressed(os.path.join(HERE, "opus_synthetic.npz"), bytes=np.frombuffer(b"".join(pk), dtype=np.uint8), off=off,
                    fs=np.array(fss, dtype=np.int32), ref=ref)
print(len(pk), "synthetic cases,", int((ref["count"] > 0).sum()), "accepted by the reference")
