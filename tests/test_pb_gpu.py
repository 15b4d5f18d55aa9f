"""The batched GPU deframer (include/anmodem_pb.h) against the REFERENCE's own
nanopb decoder compiled in place (oracle/_ref: pb_decode_delimited(ToReceiver_fields) with the
reference's field callback, hardware/src/network.cpp:212-249, 406-430) on valid, truncated, mutated
and random messages, and end to end behind the CUDA demodulator."""
import ctypes as C
import json
import os

import numpy as np
import pytest

import audio_network_b200 as anm
from oracle_binding import REF_LIB

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not os.path.exists(REF_LIB), reason="oracle/_ref not built (reference tree absent)")]

GOLD = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "pb_messages.json")))


def _ref():
    R = C.CDLL(REF_LIB)
    R.ref_decode_to_receiver_audio.restype = C.c_long
    R.ref_decode_to_receiver_audio.argtypes = [C.c_char_p, C.c_size_t, C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t)]
    return R


def _varint(v):
    out = bytearray()
    while True:
        b = v & 0x7F
        v >>= 7
        out.append(b | (0x80 if v else 0))
        if not v:
            return bytes(out)


def _msg(audio, extra_outer=b"", extra_inner=b"", pre_inner=b""):
    inner = pre_inner + b"\x0a" + _varint(len(audio)) + audio + extra_inner
    outer = b"\x0a" + _varint(len(inner)) + inner + extra_outer
    return _varint(len(outer)) + outer


def _corpus():
    rng = np.random.default_rng(20260101)
    msgs = [bytes.fromhex(r["wire"]) for r in GOLD.values() if isinstance(r, dict) and "wire" in r]
    for n in (0, 1, 2, 127, 128, 300, 4095, 4096, 4097, 5000):
        msgs.append(_msg(bytes(rng.integers(0, 256, n, dtype=np.uint8))))
    a = bytes(range(50))
    unknown = [b"\x10\x05", b"\x10\xff\xff\xff\xff\xff\xff\xff\xff\xff\x01", b"\x19" + bytes(8), b"\x1a\x03abc", b"\x25" + bytes(4),
               b"\x13", b"\x14", b"\x16", b"\x17", b"\x00\x00", b"\x1a\x7f", b"\xf8\xff\xff\xff\x0f\x01", b"\xf8\xff\xff\xff\x1f\x01",
               b"\x10" + b"\x80" * 9 + b"\x00", b"\x10" + b"\x80" * 12 + b"\x00"]
    for u in unknown:
        msgs += [_msg(a, extra_outer=u), _msg(a, extra_inner=u), _msg(a, pre_inner=u)]
    # AudioData field 1 with scalar wire types (the reference hands the raw bytes to its callback), duplicates, missing
    for inner in (b"\x08\x05", b"\x08\xff\xff\xff\xff\xff\xff\xff\xff\xff\x01", b"\x08" + b"\xff" * 10 + b"\x01", b"\x09" + bytes(range(8)),
                  b"\x0d" + bytes(range(4)), b"\x0b", b"\x0a\x02hi\x0a\x03abc", b"", b"\x12\x01x"):
        outer = b"\x0a" + _varint(len(inner)) + inner
        msgs.append(_varint(len(outer)) + outer)
    msgs += [b"\x00", b"", b"\x01", b"\x02\x0a", b"\x02\x0a\x00", b"\x04\x0a\x00\x0a\x00", b"\x08\x0a\x02\x0a\x00\x0a\x02\x12\x00",
             b"\x02\x08\x01", b"\x02\x0d\x01", b"\x80\x80\x80\x80\x80\x80\x80\x80\x80\x80\x00", b"\xff\xff\xff\xff\x0f", b"\x85\x80\x80\x80\x10\x0a\x03\x0a\x01x",
             b"\x85\x80\x80\x80\x00\x0a\x03\x0a\x01x", _msg(a) + b"trailing"]
    base = [m for m in msgs if 4 < len(m) < 400]
    for m in base:                                   # truncations and byte mutations
        for cut in (1, 2, len(m) // 2):
            msgs.append(m[:-cut])
        for _ in range(6):
            b = bytearray(m)
            for _k in range(int(rng.integers(1, 4))):
                b[int(rng.integers(0, len(b)))] = int(rng.integers(0, 256))
            msgs.append(bytes(b))
    for _ in range(300):                             # short random strings biased to protobuf-looking bytes
        n = int(rng.integers(1, 24))
        msgs.append(bytes(rng.choice([0x00, 0x01, 0x02, 0x05, 0x08, 0x0a, 0x0d, 0x10, 0x12, 0x1a, 0x7f, 0x80, 0xff], size=n).astype(np.uint8)))
    return msgs


def test_deframer_matches_reference_nanopb():
    R = _ref()
    msgs = _corpus()
    assert len(msgs) > 900
    recs = np.zeros(len(msgs), dtype=anm.FRAME_DTYPE)
    off = 0
    for i, m in enumerate(msgs):
        recs[i] = (i, len(m), 0, 1, off)
        off += len(m)
    arena = np.frombuffer(b"".join(msgs), dtype=np.uint8)
    spans = anm.pb_deframe(recs, arena)
    out = (C.c_uint8 * 8192)()
    kinds = {0: 0, 1: 0, 2: 0}
    for i, m in enumerate(msgs):
        used = C.c_size_t(0)
        n = R.ref_decode_to_receiver_audio(m, len(m), out, 8192, C.byref(used))
        s = spans[i]
        if n == -1:
            assert s["status"] == anm.ANM_PB_FAIL, (i, m.hex(), s)
        elif n == -2:
            assert s["status"] == anm.ANM_PB_NO_AUDIO and s["consumed"] == used.value, (i, m.hex(), s, used.value)
        else:
            assert s["status"] == anm.ANM_PB_OK and s["consumed"] == used.value and s["audio_len"] == n, (i, m.hex(), s, n, used.value)
            assert arena[s["audio_offset"]: s["audio_offset"] + n].tobytes() == bytes(out[:n]), (i, m.hex())
        kinds[int(s["status"])] += 1
    assert min(kinds.values()) > 10          # the corpus exercises every verdict
    # the host scanner is the same walk (anm_pb_wire.h): it accepts exactly what the GPU deframer reports as OK, with the same span
    L = anm.lib()
    L.anm_pb_scan_to_receiver_audio.restype = C.c_size_t
    L.anm_pb_scan_to_receiver_audio.argtypes = [C.c_char_p, C.c_size_t, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t)]
    for i, m in enumerate(msgs):
        p, ln = C.c_void_p(), C.c_size_t()
        buf = C.create_string_buffer(m, max(1, len(m)))
        used = L.anm_pb_scan_to_receiver_audio(buf, len(m), C.byref(p), C.byref(ln))
        s = spans[i]
        if s["status"] == anm.ANM_PB_OK:
            assert used == s["consumed"] and ln.value == s["audio_len"], (i, m.hex())
            assert p.value - C.addressof(buf) == int(s["audio_offset"]) - int(recs[i]["offset"]), (i, m.hex())
        else:
            assert used == 0, (i, m.hex(), int(s["status"]))


def test_crc_failed_frames_are_not_decoded_and_ring_addressing():
    torch = pytest.importorskip("torch")
    m = _msg(b"opus-bytes-0123456789")
    cap = 64                                          # a 64-byte ring: the message wraps around its end
    ring = np.zeros(cap, dtype=np.uint8)
    start = cap - 7
    for k, b in enumerate(m):
        ring[(start + k) % cap] = b
    recs = np.zeros(2, dtype=anm.FRAME_DTYPE)
    recs[0] = (0, len(m), 0, 1, start)
    recs[1] = (1, len(m), 0, 0, start)
    d_f = torch.from_numpy(recs.view(np.uint8).copy()).cuda()
    d_b = torch.from_numpy(ring).cuda()
    d_o = torch.zeros(2 * 16, dtype=torch.uint8, device="cuda")
    rc = anm.lib().anm_pb_deframe_device(d_f.data_ptr(), 2, d_b.data_ptr(), cap - 1, d_o.data_ptr(), None)
    assert rc == 0
    torch.cuda.synchronize()
    sp = d_o.cpu().numpy().view(anm.PB_SPAN_DTYPE)
    assert sp[0]["status"] == anm.ANM_PB_OK and sp[0]["consumed"] == len(m) and sp[0]["audio_len"] == 21
    got = bytes(ring[(int(sp[0]["audio_offset"]) + k) % cap] for k in range(21))
    assert got == b"opus-bytes-0123456789"
    assert sp[1]["status"] == anm.ANM_PB_CRC
    assert anm.lib().anm_pb_deframe_device(d_f.data_ptr(), 2, d_b.data_ptr(), 100, d_o.data_ptr(), None) == anm.ANM_ERR_ARG


def test_demodulated_frames_deframe_to_the_sent_opus_packets():
    from sigutil import make_program

    L = anm.lib()
    L.anm_pb_encode_to_receiver_audio.restype = C.c_size_t
    L.anm_pb_encode_to_receiver_audio.argtypes = [C.c_char_p, C.c_size_t, C.c_void_p, C.c_size_t]
    cfg = anm.config_preset("ref4")
    rng = np.random.default_rng(7)
    n_ch, n = 8, 2400 * cfg.sym_len
    pcm = np.zeros((n_ch, n), dtype=np.int16)
    sent = []
    buf = (C.c_uint8 * 512)()
    for c in range(n_ch):
        prog, packets = [np.full(5, anm.ANM_SILENCE, np.uint8)], []
        while sum(len(p) for p in prog) < 2400:
            opus = bytes(rng.integers(0, 256, int(rng.integers(1, 120)), dtype=np.uint8))
            k = L.anm_pb_encode_to_receiver_audio(opus, len(opus), buf, 512)
            prog += [anm.frame_symbols(cfg, bytes(buf[:k])), np.full(6, anm.ANM_SILENCE, np.uint8)]
            packets.append(opus)
        sent.append(packets)
        pcm[c] = anm.tx_render(cfg, np.concatenate(prog), anm.tx_params(seed=100 + c, amplitude=0.5, snr_db=9.0), 0, n)
    dm = anm.Demod(cfg, n_ch, device=0)
    dm.feed_host(pcm)
    dm.collect()
    recs, by = dm.read_frames()
    dm.close()
    spans = anm.pb_deframe(recs, by)
    assert len(recs) >= n_ch * 4
    per_ch = {c: [] for c in range(n_ch)}
    for r, s in zip(recs, spans):
        assert r["crc_ok"] == 1 and s["status"] == anm.ANM_PB_OK and s["consumed"] == r["len"]
        per_ch[int(r["channel"])].append(by[s["audio_offset"]: s["audio_offset"] + s["audio_len"]].tobytes())
    for c in range(n_ch):
        assert per_ch[c] == sent[c][: len(per_ch[c])] and len(per_ch[c]) >= 4
