"""Corpus of BroadcastMessage / ToTransmitter byte strings (valid, crafted, truncated, mutated) and the
ctypes view of the REFERENCE's decoders in oracle/_ref (ref_shim.c) -- shared by tests/test_pb_msgs.py
and tests/golden/make_pb_handshake.py.  Test infrastructure."""
import ctypes as C

import numpy as np


class RefBroadcast(C.Structure):  # ref_broadcast_t, oracle/ref_shim.c
    _fields_ = [("magic", C.c_uint32), ("which", C.c_uint32), ("discovery_request", C.c_uint32), ("protocol_version", C.c_uint32),
                ("mac", C.c_uint64), ("streaming", C.c_uint32), ("pad", C.c_uint32), ("device_name", C.c_char * 128), ("opus_version", C.c_char * 128)]


class RefToTransmitter(C.Structure):  # ref_to_transmitter_t
    _fields_ = [("which", C.c_uint32), ("protocol_version", C.c_uint32), ("mac", C.c_uint64), ("streaming", C.c_uint32), ("max_enc", C.c_uint32),
                ("max_dec", C.c_uint32), ("underflow", C.c_uint32), ("decode_error", C.c_uint32), ("pad", C.c_uint32),
                ("device_name", C.c_char * 128), ("opus_version", C.c_char * 128)]


def ref_lib(path):
    R = C.CDLL(path)
    R.ref_decode_broadcast.argtypes = [C.c_char_p, C.c_size_t, C.POINTER(RefBroadcast)]
    R.ref_decode_to_transmitter.argtypes = [C.c_char_p, C.c_size_t, C.POINTER(RefToTransmitter)]
    for n in ("ref_encode_broadcast_request", "ref_encode_broadcast_response", "ref_encode_to_transmitter_info", "ref_encode_to_transmitter_error"):
        getattr(R, n).restype = C.c_size_t
    return R


def raw128(arr):
    """all 128 bytes of a char[128] field (ctypes .value would stop at the first NUL)"""
    return bytes(C.string_at(C.addressof(arr) if not isinstance(arr, bytes) else arr, 128)) if not isinstance(arr, bytes) else arr


def ref_decode_broadcast(R, wire):
    out = RefBroadcast()
    if R.ref_decode_broadcast(wire, len(wire), C.byref(out)) != 0:
        return None
    return {"magic": out.magic, "which": out.which, "discovery_request": out.discovery_request, "protocol_version": out.protocol_version,
            "mac": out.mac, "streaming": out.streaming,
            "device_name": C.string_at(C.addressof(out) + RefBroadcast.device_name.offset, 128).hex(),
            "opus_version": C.string_at(C.addressof(out) + RefBroadcast.opus_version.offset, 128).hex()}


def ref_decode_to_transmitter(R, wire):
    out = RefToTransmitter()
    if R.ref_decode_to_transmitter(wire, len(wire), C.byref(out)) != 0:
        return None
    return {"which": out.which, "protocol_version": out.protocol_version, "mac": out.mac, "streaming": out.streaming, "max_enc": out.max_enc,
            "max_dec": out.max_dec, "underflow": out.underflow, "decode_error": out.decode_error,
            "device_name": C.string_at(C.addressof(out) + RefToTransmitter.device_name.offset, 128).hex(),
            "opus_version": C.string_at(C.addressof(out) + RefToTransmitter.opus_version.offset, 128).hex()}


def varint(v):
    out = bytearray()
    while True:
        b = v & 0x7F
        v >>= 7
        out.append(b | (0x80 if v else 0))
        if not v:
            return bytes(out)


def ld(field, body):
    return varint((field << 3) | 2) + varint(len(body)) + body


def vi(field, v):
    return varint(field << 3) + varint(v)


def discovery(version=1, mac=0x24A160123456, name=b"Audio-Network Receiver", streaming=0, opus=b"libopus 1.3.1-fixed", skip=(), extra=b""):
    parts = {1: vi(1, version), 2: vi(2, mac), 3: ld(3, name), 4: vi(4, streaming), 5: ld(5, opus)}
    return b"".join(v for k, v in parts.items() if k not in skip) + extra


def delimited(body):
    return varint(len(body)) + body


UNKNOWN = [b"\x30\x05", b"\x30\xff\xff\xff\xff\xff\xff\xff\xff\xff\x01", b"\x39" + bytes(8), b"\x3a\x03abc", b"\x3d" + bytes(4), b"\x33", b"\x34",
           b"\x36", b"\x37", b"\x00\x00", b"\x3a\x7f", b"\xf8\xff\xff\xff\x0f\x01", b"\xf8\xff\xff\xff\x1f\x01", b"\x30" + b"\x80" * 9 + b"\x00",
           b"\x30" + b"\x80" * 12 + b"\x00"]
BIG = [0, 1, 127, 128, 0x7FFFFFFF, 0x80000000, 0xFFFFFFFF, 0x100000000, (1 << 63), (1 << 64) - 1]


def broadcast_corpus(seed=1):
    rng = np.random.default_rng(seed)
    magic = vi(1, 0x2C5DA044)
    d = discovery()
    msgs = [delimited(magic + vi(2, 1)), delimited(magic + ld(3, d)), delimited(magic), delimited(vi(2, 1)), delimited(ld(3, d)), delimited(b""),
            delimited(magic + vi(2, 1) + ld(3, d)), delimited(magic + ld(3, d) + vi(2, 0)), delimited(magic + ld(3, d) + vi(2, 1) + ld(3, d)),
            delimited(magic + ld(3, discovery(name=b"first")) + ld(3, discovery(name=b"second-longer", version=7))),
            delimited(magic + ld(3, discovery(name=b"a-long-first-name")) + ld(3, discovery(name=b"xy"))),
            delimited(magic + ld(3, d) + ld(3, discovery(skip=(3,)))), delimited(ld(3, d) + magic + magic), delimited(magic + ld(3, d)) + b"trailing",
            delimited(magic + ld(3, discovery(name=b"x" * 127))), delimited(magic + ld(3, discovery(name=b"x" * 128))),
            delimited(magic + ld(3, discovery(opus=b"y" * 127, name=b""))), delimited(magic + ld(3, discovery(opus=b"y" * 200))),
            delimited(magic + ld(3, discovery(name=b"nul\x00inside"))), delimited(magic + ld(3, discovery(name=b"\xff\xfe bad utf8")))]
    for k in range(1, 6):
        msgs.append(delimited(magic + ld(3, discovery(skip=(k,)))))
    for v in BIG:
        msgs += [delimited(vi(1, v) + vi(2, 1)), delimited(magic + vi(2, v)), delimited(magic + ld(3, discovery(version=v))),
                 delimited(magic + ld(3, discovery(mac=v))), delimited(magic + ld(3, discovery(streaming=v)))]
    for wt in range(8):  # every wire type on every known field
        for f in (1, 2, 3):
            msgs.append(delimited(varint((f << 3) | wt) + b"\x01\x02\x03\x04\x05\x06\x07\x08\x09"))
        for f in (1, 2, 3, 4, 5):
            msgs.append(delimited(magic + ld(3, d + varint((f << 3) | wt) + b"\x01\x02\x03\x04\x05\x06\x07\x08\x09")))
    for u in UNKNOWN:
        msgs += [delimited(magic + vi(2, 1) + u), delimited(u + magic + ld(3, d)), delimited(magic + ld(3, discovery(extra=u))), delimited(magic + ld(3, u + d))]
    msgs += [b"", b"\x00", b"\x01", b"\x05" + magic[:3], b"\x80\x80\x80\x80\x80\x80\x80\x80\x80\x80\x00", b"\xff\xff\xff\xff\x0f", b"\x87\x80\x80\x80\x00" + magic + vi(2, 1),
             b"\x87\x80\x80\x80\x10" + magic + vi(2, 1)]
    return _mutate(msgs, rng)


def to_transmitter_corpus(seed=2):
    rng = np.random.default_rng(seed)
    d = discovery()
    info = ld(1, d) + vi(2, 4096) + vi(3, 11520)
    err = vi(1, 1) + vi(2, 0)
    msgs = [delimited(ld(1, info)), delimited(ld(2, err)), delimited(b""), delimited(ld(1, info) + ld(2, err)), delimited(ld(2, err) + ld(1, info)),
            delimited(ld(1, info) + ld(1, ld(1, discovery(name=b"second")) + vi(2, 1) + vi(3, 2))), delimited(ld(2, err) + ld(2, vi(1, 0) + vi(2, 1))),
            delimited(ld(1, info) + ld(2, err) + ld(1, info)), delimited(ld(1, ld(1, d) + ld(1, discovery(name=b"merged", skip=()))) + vi(2, 5) + vi(3, 6)),
            delimited(ld(1, ld(1, d) + ld(1, discovery(name=b"merged")) + vi(2, 5) + vi(3, 6))), delimited(ld(1, ld(1, d) + ld(1, discovery(skip=(5,))) + vi(2, 5) + vi(3, 6))),
            delimited(ld(1, vi(3, 6) + vi(2, 5) + ld(1, d))), delimited(ld(2, vi(2, 1) + vi(1, 1))), delimited(ld(2, vi(1, 1))), delimited(ld(2, vi(2, 1))), delimited(ld(2, b"")),
            delimited(ld(1, b"")), delimited(ld(1, ld(1, d) + vi(2, 1))), delimited(ld(1, ld(1, d) + vi(3, 1))), delimited(ld(1, vi(2, 1) + vi(3, 1))), delimited(ld(1, info)) + b"more"]
    for v in BIG:
        msgs += [delimited(ld(1, ld(1, d) + vi(2, v) + vi(3, 1))), delimited(ld(1, ld(1, d) + vi(2, 1) + vi(3, v))), delimited(ld(2, vi(1, v) + vi(2, 1))),
                 delimited(ld(2, vi(1, 1) + vi(2, v)))]
    for wt in range(8):
        for f in (1, 2):
            msgs.append(delimited(varint((f << 3) | wt) + b"\x01\x02\x03\x04\x05\x06\x07\x08\x09"))
        for f in (1, 2, 3):
            msgs.append(delimited(ld(1, info + varint((f << 3) | wt) + b"\x01\x02\x03\x04\x05\x06\x07\x08\x09")))
        for f in (1, 2):
            msgs.append(delimited(ld(2, err + varint((f << 3) | wt) + b"\x01\x02\x03\x04\x05\x06\x07\x08\x09")))
    for u in UNKNOWN:
        msgs += [delimited(ld(1, info) + u), delimited(u + ld(2, err)), delimited(ld(1, info + u)), delimited(ld(2, u + err)), delimited(ld(1, ld(1, d + u) + vi(2, 1) + vi(3, 1)))]
    msgs += [b"", b"\x00", b"\x02\x12", b"\x03\x12\x04\x08", b"\xff\xff\xff\xff\x1f"]
    return _mutate(msgs, rng)


def _mutate(msgs, rng):
    out = list(msgs)
    for m in msgs:
        if len(m) < 4:
            continue
        for cut in (1, 2, len(m) // 2):
            out.append(m[:-cut])
        for _ in range(4):
            b = bytearray(m)
            for _k in range(int(rng.integers(1, 4))):
                b[int(rng.integers(0, len(b)))] = int(rng.integers(0, 256))
            out.append(bytes(b))
    for _ in range(200):
        n = int(rng.integers(1, 28))
        out.append(bytes(rng.choice([0x00, 0x01, 0x02, 0x04, 0x08, 0x0a, 0x10, 0x12, 0x18, 0x1a, 0x20, 0x2a, 0x7f, 0x80, 0xff], size=n).astype(np.uint8)))
    return out
