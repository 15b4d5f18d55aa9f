#!/usr/bin/env python
"""Regenerates the committed fixtures under tests/golden/.

  pb_messages.json   varint-delimited ip.proto messages produced by the REFERENCE's own
                     nanopb 0.4.5 encoder + generated ip.pb.c (oracle/_ref, compiled from
                     /root/reference by oracle/Makefile) -- the one stage of the path that is
                     pinned by reference code.
  modem_kat.npz      known-answer vectors of the SPEC.md modem: a short multi-channel PCM
                     capture (CPU transmitter stand-in) with the frames, symbols and a tone-
                     energy digest the CPU oracle produced for it when this file was written.
                     They pin the oracle + SPEC against silent drift; they are NOT reference
                     outputs (the reference has no demodulator: parity unpinned).

Run in the build container (needs /root/reference for the first file):
    python tests/golden/make_golden.py
"""
import ctypes as C
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import audio_network_b200 as anm  # noqa: E402
from oracle_binding import REF_LIB, Oracle  # noqa: E402
from sigutil import make_channels  # noqa: E402


def pb_messages():
    L = C.CDLL(REF_LIB)
    L.ref_encode_to_receiver_audio.restype = C.c_size_t
    L.ref_encode_broadcast_request.restype = C.c_size_t
    L.ref_encode_broadcast_response.restype = C.c_size_t
    L.ref_encode_to_transmitter_info.restype = C.c_size_t
    L.ref_encode_to_transmitter_error.restype = C.c_size_t
    out = {}
    buf = (C.c_uint8 * 8192)()
    rng = np.random.default_rng(2024)
    for n in (0, 1, 5, 127, 128, 300, 4096):
        data = rng.integers(0, 256, size=n, dtype=np.uint8).tobytes()
        ln = L.ref_encode_to_receiver_audio(data, C.c_size_t(n), buf, C.c_size_t(8192))
        out["to_receiver_audio_%d" % n] = {"payload": data.hex(), "wire": bytes(buf[:ln]).hex()}
    ln = L.ref_encode_broadcast_request(C.c_uint32(0x2C5DA044), buf, C.c_size_t(8192))
    out["broadcast_request"] = {"magic": 0x2C5DA044, "wire": bytes(buf[:ln]).hex()}
    ln = L.ref_encode_broadcast_response(C.c_uint32(0x2C5DA044), C.c_uint32(1), C.c_uint64(0x24A160123456),
                                         b"Audio-Network Receiver", C.c_int(1), b"libopus 1.3.1-fixed", buf, C.c_size_t(8192))
    out["broadcast_response"] = {"wire": bytes(buf[:ln]).hex(), "protocol_version": 1, "mac": 0x24A160123456,
                                 "device_name": "Audio-Network Receiver", "streaming": 1, "opus_version": "libopus 1.3.1-fixed"}
    ln = L.ref_encode_to_transmitter_info(C.c_uint32(1), C.c_uint64(0x24A160123456), b"Audio-Network Receiver", C.c_int(0),
                                          b"libopus 1.3.1-fixed", C.c_uint32(4096), C.c_uint32(11520), buf, C.c_size_t(8192))
    out["to_transmitter_info"] = {"wire": bytes(buf[:ln]).hex(), "max_enc": 4096, "max_dec": 11520}
    ln = L.ref_encode_to_transmitter_error(C.c_int(1), C.c_int(0), buf, C.c_size_t(8192))
    out["to_transmitter_error"] = {"wire": bytes(buf[:ln]).hex(), "underflow": 1, "decode_error": 0}
    with open(os.path.join(HERE, "pb_messages.json"), "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)
    print("wrote pb_messages.json (%d messages)" % len(out))


def modem_kat():
    arrays = {}
    for name, n_ch, n_sym, snr, ppm in (("ref4", 6, 420, 8.0, 150.0), ("bfsk2", 2, 500, None, 0.0), ("mfsk16", 2, 300, 12.0, 0.0),
                                         ("wide64", 2, 160, 10.0, 0.0)):
        cfg = anm.config_preset(name)
        pcm, _ = make_channels(cfg, n_ch, n_sym * cfg.sym_len, seed=77, snr_db=snr, ppm_max=ppm, offset_max=1500)
        arrays[name + "_pcm"] = pcm
        hops = pcm.shape[1] // cfg.hop
        recs, syms, edig = [], [], []
        for c in range(n_ch):
            o = Oracle(cfg, trace_hops=hops)
            o.feed(pcm[c])
            for (_, start, ok, payload) in o.frames(c):
                recs.append((c, start, ok, payload.hex()))
            syms.append(o.symbols())
            edig.append(hashlib.sha256(o.E.tobytes()).hexdigest())
            if c == 0:
                arrays[name + "_E0"] = o.E[:64].copy()   # first 64 hops of channel 0, raw fp32
        arrays[name + "_frames"] = np.array(json.dumps(recs))
        arrays[name + "_symbols"] = np.array(json.dumps([s.tolist() for s in syms]))
        arrays[name + "_energy_sha256"] = np.array(json.dumps(edig))
        print(name, "frames:", len(recs), "crc ok:", sum(r[2] for r in recs))
    np.savez_compressed(os.path.join(HERE, "modem_kat.npz"), **arrays)
    print("wrote modem_kat.npz")


if __name__ == "__main__":
    if os.path.exists(REF_LIB):
        pb_messages()
    else:
        print("oracle/_ref missing: pb_messages.json not regenerated")
    modem_kat()
