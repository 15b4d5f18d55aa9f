#!/usr/bin/env python
"""Regenerates tests/golden/pb_handshake.json: the verdict and the decoded fields of the REFERENCE's
nanopb 0.4.5 (oracle/_ref, pb_decode_delimited with BroadcastMessage_fields / ToTransmitter_fields) on
the corpus of tests/pb_corpus.py -- valid discovery / hello messages, every wire type on every field,
varint and string limits, oneof switches, duplicate submessages, truncations and byte mutations.
Needs /root/reference (via `make -C oracle`); the committed JSON travels without it.
    python tests/golden/make_pb_handshake.py
"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import pb_corpus as pc  # noqa: E402
from oracle_binding import REF_LIB  # noqa: E402

R = pc.ref_lib(REF_LIB)
out = {"broadcast": [], "to_transmitter": []}
for w in pc.broadcast_corpus():
    out["broadcast"].append({"wire": w.hex(), "ref": pc.ref_decode_broadcast(R, w)})
for w in pc.to_transmitter_corpus():
    out["to_transmitter"].append({"wire": w.hex(), "ref": pc.ref_decode_to_transmitter(R, w)})
with open(os.path.join(HERE, "pb_handshake.json"), "w") as f:
    json.dump(out, f, separators=(",", ":"))
for k, v in out.items():
    print(k, len(v), "messages,", sum(1 for r in v if r["ref"] is not None), "accepted by the reference")
