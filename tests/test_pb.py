"""The hand-off to the protobuf decoder: wire helpers against bytes produced by the REFERENCE's
nanopb encoder (tests/golden/pb_messages.json, made by tests/golden/make_golden.py), and -- when
oracle/_ref is present -- against the reference decoder itself, including decoding THROUGH the
product's pb_istream byte source (the drop-in seam of hardware/src/network.cpp:262-305, 406-411)."""
import ctypes as C
import json
import os

import numpy as np
import pytest

import audio_network_b200 as anm
from oracle_binding import REF_LIB

GOLD = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "pb_messages.json")))


def _lib():
    L = anm.lib()
    L.anm_pb_encode_to_receiver_audio.restype = C.c_size_t
    L.anm_pb_encode_to_receiver_audio.argtypes = [C.c_char_p, C.c_size_t, C.c_void_p, C.c_size_t]
    L.anm_pb_encode_broadcast_request.restype = C.c_size_t
    L.anm_pb_encode_broadcast_request.argtypes = [C.c_uint32, C.c_void_p, C.c_size_t]
    L.anm_pb_scan_to_receiver_audio.restype = C.c_size_t
    L.anm_pb_scan_to_receiver_audio.argtypes = [C.c_char_p, C.c_size_t, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t)]
    L.anm_pb_queue_create.restype = C.c_void_p
    L.anm_pb_queue_push.argtypes = [C.c_void_p, C.c_char_p, C.c_size_t]
    L.anm_pb_queue_destroy.argtypes = [C.c_void_p]
    L.anm_pb_queue_size.restype = C.c_size_t
    L.anm_pb_queue_size.argtypes = [C.c_void_p]
    return L


class PbStream(C.Structure):
    _fields_ = [("callback", C.c_void_p), ("state", C.c_void_p), ("bytes_left", C.c_size_t), ("errmsg", C.c_char_p)]


def test_encoders_match_reference_bytes():
    L = _lib()
    buf = (C.c_uint8 * 8192)()
    for key, rec in GOLD.items():
        if key.startswith("to_receiver_audio_"):
            data = bytes.fromhex(rec["payload"])
            n = L.anm_pb_encode_to_receiver_audio(data, len(data), buf, 8192)
            assert bytes(buf[:n]).hex() == rec["wire"], key
    n = L.anm_pb_encode_broadcast_request(GOLD["broadcast_request"]["magic"], buf, 8192)
    assert bytes(buf[:n]).hex() == GOLD["broadcast_request"]["wire"]
    assert L.anm_pb_encode_to_receiver_audio(b"abc", 3, buf, 4) == 0  # capacity too small


def test_scan_walks_reference_bytes_and_rejects_garbage():
    L = _lib()
    for key, rec in GOLD.items():
        if not key.startswith("to_receiver_audio_"):
            continue
        wire = bytes.fromhex(rec["wire"])
        p, n = C.c_void_p(), C.c_size_t()
        used = L.anm_pb_scan_to_receiver_audio(wire, len(wire), C.byref(p), C.byref(n))
        assert used == len(wire)
        assert C.string_at(p.value, n.value) == bytes.fromhex(rec["payload"])
        # truncated input, as after a dropped frame
        assert L.anm_pb_scan_to_receiver_audio(wire[:-1], len(wire) - 1, C.byref(p), C.byref(n)) == 0
    bad = bytes([0x05, 0x0A, 0xFF, 0xFF, 0xFF, 0xFF])
    p, n = C.c_void_p(), C.c_size_t()
    assert L.anm_pb_scan_to_receiver_audio(bad, len(bad), C.byref(p), C.byref(n)) == 0
    # a ToTransmitter message is not a ToReceiver{audio_data}
    wire = bytes.fromhex(GOLD["to_transmitter_error"]["wire"])
    assert L.anm_pb_scan_to_receiver_audio(wire, len(wire), C.byref(p), C.byref(n)) == 0


@pytest.mark.skipif(not os.path.exists(REF_LIB), reason="oracle/_ref not built (reference tree absent)")
def test_reference_decoder_reads_through_product_stream():
    L = _lib()
    R = C.CDLL(REF_LIB)
    R.ref_decode_to_receiver_from_stream.restype = C.c_long
    R.ref_decode_to_receiver_from_stream.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]
    q = L.anm_pb_queue_create()
    L.anm_pb_istream_from_queue.restype = PbStream
    L.anm_pb_istream_from_queue.argtypes = [C.c_void_p]
    payloads = []
    for key in sorted(GOLD):
        if key.startswith("to_receiver_audio_"):
            wire = bytes.fromhex(GOLD[key]["wire"])
            payloads.append(bytes.fromhex(GOLD[key]["payload"]))
            L.anm_pb_queue_push(q, wire, len(wire))  # three frames' payloads back to back, as the demodulator emits them
    s = L.anm_pb_istream_from_queue(q)
    out = (C.c_uint8 * 8192)()
    for want in payloads:
        n = R.ref_decode_to_receiver_from_stream(s.callback, s.state, out, 8192)
        assert n == len(want) and bytes(out[:n]) == want
    assert L.anm_pb_queue_size(q) == 0
    # the stream is exhausted: the reference decoder must fail like on a closed socket
    assert R.ref_decode_to_receiver_from_stream(s.callback, s.state, out, 8192) == -1
    L.anm_pb_queue_destroy(q)


@pytest.mark.skipif(not os.path.exists(REF_LIB), reason="oracle/_ref not built (reference tree absent)")
def test_reference_decodes_golden_messages():
    R = C.CDLL(REF_LIB)

    class Bc(C.Structure):
        _fields_ = [("magic", C.c_uint32), ("which", C.c_uint32), ("discovery_request", C.c_uint32), ("protocol_version", C.c_uint32),
                    ("mac", C.c_uint64), ("streaming", C.c_uint32), ("pad", C.c_uint32), ("device_name", C.c_char * 128), ("opus_version", C.c_char * 128)]

    b = Bc()
    wire = bytes.fromhex(GOLD["broadcast_request"]["wire"])
    assert R.ref_decode_broadcast(wire, C.c_size_t(len(wire)), C.byref(b)) == 0
    assert b.magic == 0x2C5DA044 and b.which == 2 and b.discovery_request == 1  # protocol/ip.proto:9-18
    wire = bytes.fromhex(GOLD["broadcast_response"]["wire"])
    assert R.ref_decode_broadcast(wire, C.c_size_t(len(wire)), C.byref(b)) == 0
    assert b.which == 3 and b.mac == 0x24A160123456 and b.device_name == b"Audio-Network Receiver"
    # a payload over MAX_ENCODED_FRAME_SIZE (hardware/src/network.cpp:24,223) is refused by the reference callback
    L = _lib()
    buf = (C.c_uint8 * 8192)()
    n = L.anm_pb_encode_to_receiver_audio(bytes(4097), 4097, buf, 8192)
    out = (C.c_uint8 * 8192)()
    R.ref_decode_to_receiver_audio.restype = C.c_long
    assert R.ref_decode_to_receiver_audio(buf, C.c_size_t(n), out, C.c_size_t(8192), None) == -1


def test_frame_payloads_carry_delimited_messages_end_to_end():
    """payload -> frame symbols -> PCM -> oracle -> payload -> scan: the modem is transparent."""
    from oracle_binding import Oracle

    L = _lib()
    cfg = anm.config_preset("ref4")
    opus = bytes(range(40))
    buf = (C.c_uint8 * 256)()
    n = L.anm_pb_encode_to_receiver_audio(opus, len(opus), buf, 256)
    wire = bytes(buf[:n])
    syms = anm.frame_symbols(cfg, wire)
    prog = np.concatenate([np.full(3, 255, np.uint8), syms, np.full(20, 255, np.uint8)])
    pcm = anm.tx_render(cfg, prog, anm.tx_params(seed=1, amplitude=0.5, snr_db=10.0), 0, (len(prog) + 4) * cfg.sym_len)
    o = Oracle(cfg)
    o.feed(pcm)
    (_, _, ok, payload), = o.frames()
    assert ok and payload == wire
    p, ln = C.c_void_p(), C.c_size_t()
    assert L.anm_pb_scan_to_receiver_audio(payload, len(payload), C.byref(p), C.byref(ln)) == len(wire)
    assert C.string_at(p.value, ln.value) == opus
