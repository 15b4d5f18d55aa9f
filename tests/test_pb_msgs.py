"""Discovery / handshake message codec (include/anmodem_pb.h, csrc/anm_pb_msgs.c)
against the REFERENCE's nanopb 0.4.5: the committed verdicts and fields of tests/golden/pb_handshake.json
(made by tests/golden/make_pb_handshake.py from oracle/_ref) and, when oracle/_ref is present, the
reference decoder / encoder themselves on fresh random messages."""
import ctypes as C
import json
import os

import numpy as np
import pytest

import audio_network_b200 as anm
import pb_corpus as pc
from oracle_binding import REF_LIB

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = json.load(open(os.path.join(HERE, "golden", "pb_handshake.json")))
MSGS = json.load(open(os.path.join(HERE, "golden", "pb_messages.json")))


def _name(arr_owner, field):
    return C.string_at(C.addressof(arr_owner) + type(arr_owner).__dict__[field].offset, 128).hex()


def _bc_fields(m):
    d = m.discovery_response
    return {"magic": m.magic_word, "which": m.which, "discovery_request": m.discovery_request, "protocol_version": d.protocol_version,
            "mac": d.mac_address, "streaming": d.currently_streaming, "device_name": _name(d, "device_name"), "opus_version": _name(d, "opus_version")}


def _tt_fields(m):
    d = m.discovery_data
    return {"which": m.which, "protocol_version": d.protocol_version, "mac": d.mac_address, "streaming": d.currently_streaming,
            "max_enc": m.max_encoded_frame_size, "max_dec": m.max_decoded_frame_size, "underflow": m.audio_underflow,
            "decode_error": m.audio_decode_error, "device_name": _name(d, "device_name"), "opus_version": _name(d, "opus_version")}


def _check(kind, wire, ref):
    got = anm.pb_decode_broadcast(wire) if kind == "broadcast" else anm.pb_decode_to_transmitter(wire)
    if ref is None:
        assert got is None, "%s %s: the reference rejects, the product accepts" % (kind, wire.hex())
        return 0
    assert got is not None, "%s %s: the reference accepts, the product rejects" % (kind, wire.hex())
    fields = _bc_fields(got[0]) if kind == "broadcast" else _tt_fields(got[0])
    assert fields == ref, "%s %s" % (kind, wire.hex())
    return 1


def test_decoders_match_the_committed_reference_verdicts():
    n_ok = 0
    for kind in ("broadcast", "to_transmitter"):
        assert len(GOLD[kind]) > 1500
        for rec in GOLD[kind]:
            n_ok += _check(kind, bytes.fromhex(rec["wire"]), rec["ref"])
    assert n_ok > 400


def test_consumed_counts_the_length_prefix_and_the_message_only():
    wire = bytes.fromhex(MSGS["broadcast_response"]["wire"])
    m, used = anm.pb_decode_broadcast(wire + b"\x01\x02\x03")
    assert used == len(wire) and m.which == 3
    assert m.discovery_response.device_name == MSGS["broadcast_response"]["device_name"].encode()
    assert m.discovery_response.mac_address == MSGS["broadcast_response"]["mac"]


def test_encoders_reproduce_the_reference_encoder_bytes():
    m = anm.PbBroadcast(magic_word=0x2C5DA044, which=2, discovery_request=1)
    assert anm.pb_encode_broadcast(m).hex() == MSGS["broadcast_request"]["wire"]
    g = MSGS["broadcast_response"]
    m = anm.PbBroadcast(magic_word=0x2C5DA044, which=3)
    d = m.discovery_response
    d.protocol_version, d.mac_address, d.currently_streaming = g["protocol_version"], g["mac"], g["streaming"]
    d.device_name, d.opus_version = g["device_name"].encode(), g["opus_version"].encode()
    assert anm.pb_encode_broadcast(m).hex() == g["wire"]
    t = anm.PbToTransmitter(which=1, max_encoded_frame_size=4096, max_decoded_frame_size=11520)
    t.discovery_data.protocol_version, t.discovery_data.mac_address = 1, 0x24A160123456
    t.discovery_data.device_name, t.discovery_data.opus_version = b"Audio-Network Receiver", b"libopus 1.3.1-fixed"
    assert anm.pb_encode_to_transmitter(t).hex() == MSGS["to_transmitter_info"]["wire"]
    e = anm.PbToTransmitter(which=2, audio_underflow=1, audio_decode_error=0)
    assert anm.pb_encode_to_transmitter(e).hex() == MSGS["to_transmitter_error"]["wire"]
    # capacity too small / invalid oneof selector
    buf = (C.c_uint8 * 4)()
    assert anm.lib().anm_pb_encode_broadcast(C.byref(m), buf, 4) == 0
    assert anm.lib().anm_pb_encode_to_transmitter(C.byref(anm.PbToTransmitter(which=9)), buf, 4) == 0


def test_firmware_discovery_response_round_trips():
    """network_initialize_discovery_response (hardware/src/network.cpp:356-378): version 1, not streaming, empty name"""
    m = anm.PbBroadcast()
    anm.lib().anm_pb_firmware_discovery(0x665544332211, b"libopus 1.3.1-fixed", C.byref(m))
    wire = anm.pb_encode_broadcast(m)
    back, used = anm.pb_decode_broadcast(wire)
    assert used == len(wire) and _bc_fields(back) == _bc_fields(m)
    assert back.magic_word == 0x2C5DA044 and back.discovery_response.protocol_version == 1
    assert back.discovery_response.device_name == b"" and back.discovery_response.currently_streaming == 0
    assert len(wire) <= 288 + 2  # BroadcastMessage_size, hardware/src/protogen/ip.pb.h:185


@pytest.mark.skipif(not os.path.exists(REF_LIB), reason="oracle/_ref not built (reference tree absent)")
def test_live_differential_against_the_reference():
    R = pc.ref_lib(REF_LIB)
    rng = np.random.default_rng(99)
    buf = (C.c_uint8 * 1024)()
    n_ok = 0
    for i in range(300):  # random valid messages: same bytes from both encoders, same fields from both decoders
        name = bytes(rng.integers(1, 256, int(rng.integers(0, 128)), dtype=np.uint8))
        opus = bytes(rng.integers(1, 256, int(rng.integers(0, 128)), dtype=np.uint8))
        ver, mac = int(rng.integers(0, 2**32)), int(rng.integers(0, 2**63)) * 2 + int(rng.integers(0, 2))
        streaming = int(rng.integers(0, 2))
        n = R.ref_encode_broadcast_response(C.c_uint32(0x2C5DA044), C.c_uint32(ver), C.c_uint64(mac), name, C.c_int(streaming), opus, buf, C.c_size_t(1024))
        wire = bytes(buf[:n])
        m = anm.PbBroadcast(magic_word=0x2C5DA044, which=3)
        d = m.discovery_response
        d.protocol_version, d.mac_address, d.currently_streaming, d.device_name, d.opus_version = ver, mac, streaming, name, opus
        assert anm.pb_encode_broadcast(m) == wire
        n_ok += _check("broadcast", wire, pc.ref_decode_broadcast(R, wire))
        me, md = int(rng.integers(0, 2**32)), int(rng.integers(0, 2**32))
        n = R.ref_encode_to_transmitter_info(C.c_uint32(ver), C.c_uint64(mac), name, C.c_int(streaming), opus, C.c_uint32(me), C.c_uint32(md), buf, C.c_size_t(1024))
        wire = bytes(buf[:n])
        t = anm.PbToTransmitter(which=1, max_encoded_frame_size=me, max_decoded_frame_size=md)
        d = t.discovery_data
        d.protocol_version, d.mac_address, d.currently_streaming, d.device_name, d.opus_version = ver, mac, streaming, name, opus
        assert anm.pb_encode_to_transmitter(t) == wire
        n_ok += _check("to_transmitter", wire, pc.ref_decode_to_transmitter(R, wire))
    assert n_ok == 600
    for seed in (11, 12, 13):  # fresh mutation corpora
        for w in pc.broadcast_corpus(seed):
            _check("broadcast", w, pc.ref_decode_broadcast(R, w))
        for w in pc.to_transmitter_corpus(seed + 100):
            _check("to_transmitter", w, pc.ref_decode_to_transmitter(R, w))


def _handshake_capture(cfg, n_ch=6):
    """one channel per accepted golden message (both kinds), each carried as the payload of one frame"""
    wires = []
    for kind in ("broadcast", "to_transmitter"):
        ok = [r for r in GOLD[kind] if r["ref"] is not None and 0 < len(r["wire"]) // 2 <= 300]
        for r in ok[:: max(1, len(ok) // (n_ch // 2))][: n_ch // 2]:
            wires.append((kind, bytes.fromhex(r["wire"]), r["ref"]))
    n_sym = max(len(anm.frame_symbols(cfg, w)) for _, w, _ in wires) + 12
    pcm = np.zeros((len(wires), n_sym * cfg.sym_len), dtype=np.int16)
    for c, (_, w, _) in enumerate(wires):
        fs = anm.frame_symbols(cfg, w)
        prog = np.concatenate([np.full(3 + c, 255, np.uint8), fs, np.full(n_sym + 8 - len(fs), 255, np.uint8)])  # silence to the end of the capture
        pcm[c] = anm.tx_render(cfg, prog, anm.tx_params(seed=40 + c, amplitude=0.5, snr_db=10.0), 0, pcm.shape[1])
    return wires, pcm


def test_handshake_messages_survive_the_modem_oracle():
    """row f3: discovery / hello messages as known-answer frame payloads (CPU oracle as the demodulator)"""
    from oracle_binding import Oracle

    cfg = anm.config_preset("ref4")
    wires, pcm = _handshake_capture(cfg)
    for c, (kind, w, ref) in enumerate(wires):
        o = Oracle(cfg)
        o.feed(pcm[c])
        (_, _, ok, payload), = o.frames()
        assert ok and payload == w
        assert _check(kind, payload, ref) == 1


@pytest.mark.gpu
def test_handshake_messages_survive_the_modem_gpu():
    """the same capture through the CUDA demodulator (C ABI): payload bytes, then the product decoder's fields
    against the reference nanopb's committed fields"""
    import torch

    cfg = anm.config_preset("ref4")
    wires, pcm = _handshake_capture(cfg)
    dm = anm.Demod(cfg, len(wires), device=0)
    d_pcm = torch.from_numpy(pcm).cuda()
    dm.feed_device(d_pcm.data_ptr(), pcm.shape[1], pcm.shape[1], torch.cuda.current_stream().cuda_stream)
    dm.collect()
    frames = anm.frames_to_list(*dm.read_frames())
    dm.close()
    assert len(frames) == len(wires)
    by_ch = {f[0]: f for f in frames}
    for c, (kind, w, ref) in enumerate(wires):
        ch, _start, ok, payload = by_ch[c]
        assert ok and payload == w
        assert _check(kind, payload, ref) == 1
