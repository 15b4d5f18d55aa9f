"""Opus packets for the packet-parse tests (row f1, first stage): real packets from the REFERENCE's encoder with the
transmitter's settings, every TOC byte under every framing code, size / padding / limit edge cases, truncations,
mutations and random bytes; plus the ctypes view of oracle/_ref/libref_opus.so (oracle/ref_opus_shim.c).
Test infrastructure."""
import ctypes as C
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_OPUS = os.path.join(ROOT, "oracle", "_ref", "libref_opus.so")
FIELDS = ("count", "toc", "channels", "mode", "bandwidth", "samples_per_frame", "payload_offset", "nb_frames", "nb_samples")


class RefPacket(C.Structure):  # ref_opus_packet_t
    _fields_ = [("count", C.c_int32), ("toc", C.c_uint8), ("channels", C.c_uint8), ("pad", C.c_uint8 * 2), ("mode", C.c_int32),
                ("bandwidth", C.c_int32), ("samples_per_frame", C.c_int32), ("payload_offset", C.c_int32), ("nb_frames", C.c_int32),
                ("nb_samples", C.c_int32), ("size", C.c_int16 * 48)]


def ref_lib():
    R = C.CDLL(REF_OPUS)
    R.ref_opus_parse.argtypes = [C.c_char_p, C.c_int32, C.c_int32, C.POINTER(RefPacket)]
    R.ref_opus_parse.restype = None
    R.ref_opus_encode_stream.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int]
    R.ref_opus_decode_stream.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int]
    R.ref_opus_version.restype = C.c_char_p
    return R


def ref_parse(R, pkt, fs=48000):
    out = RefPacket()
    R.ref_opus_parse(pkt, len(pkt), fs, C.byref(out))
    d = {k: int(getattr(out, k)) for k in FIELDS}
    d["size"] = [int(v) for v in out.size]
    return d


def test_signal(n, channels, seed):
    """a few tones + noise + a transient, interleaved int16 at 48 kHz"""
    rng = np.random.default_rng(seed)
    t = np.arange(n) / 48000.0
    x = np.zeros((n, channels))
    for c in range(channels):
        x[:, c] = 6000 * np.sin(2 * np.pi * (220 * (c + 1)) * t) + 3000 * np.sin(2 * np.pi * 3520 * t + c) + 800 * rng.standard_normal(n)
    x[n // 3: n // 3 + 200] += 12000 * rng.standard_normal((200, channels))
    return np.clip(np.round(x), -32768, 32767).astype(np.int16)


def encode_stream(R, n_frames, frame_samples, channels, seed):
    pcm = test_signal(n_frames * frame_samples, channels, seed)
    max_len = 4096
    out = np.zeros((n_frames, max_len), dtype=np.uint8)
    lens = np.zeros(n_frames, dtype=np.int32)
    rc = R.ref_opus_encode_stream(pcm.ctypes.data, n_frames, frame_samples, channels, out.ctypes.data, lens.ctypes.data, max_len)
    assert rc == n_frames, rc
    return pcm, [bytes(out[i, : lens[i]]) for i in range(n_frames)]


def decode_stream(R, packets, channels, frame_samples):
    max_len = 4096
    buf = np.zeros((len(packets), max_len), dtype=np.uint8)
    lens = np.array([len(p) for p in packets], dtype=np.int32)
    for i, p in enumerate(packets):
        buf[i, : len(p)] = np.frombuffer(p, dtype=np.uint8)
    pcm = np.zeros((len(packets) * frame_samples, channels), dtype=np.int16)
    n = R.ref_opus_decode_stream(buf.ctypes.data, lens.ctypes.data, len(packets), max_len, channels, pcm.ctypes.data, frame_samples)
    assert n == len(packets) * frame_samples, n
    return pcm


def _size_bytes(n):
    n = min(n, 1275)  # the largest size the two-byte form can express
    return bytes([n]) if n < 252 else bytes([252 + (n & 3), (n - (252 + (n & 3))) >> 2])


def synthetic_corpus(seed=7):
    rng = np.random.default_rng(seed)
    body = bytes((i * 7) % 251 for i in range(6000))  # frame contents do not matter to the parser: keep the fixture compressible
    pk = []
    for toc in range(256):
        code = toc & 3
        base = toc & 0xFC
        if code == 0:
            for ln in (0, 1, 1275, 1276):
                pk.append(bytes([base]) + body[:ln])
        elif code == 1:
            for ln in (0, 1, 2, 11, 2550, 2552):
                pk.append(bytes([base | 1]) + body[:ln])
        elif code == 2:
            for first in (0, 100, 251, 252, 255, 1275):
                for rest in (0, 1275, 1276):
                    pk.append(bytes([base | 2]) + _size_bytes(first) + body[: first + rest])
            pk += [bytes([base | 2]), bytes([base | 2, 252]), bytes([base | 2, 10, 1, 2])]
        else:
            for cnt in (0, 1, 3, 6, 24, 48, 49, 63):
                for flags in (0x00, 0x40, 0x80, 0xC0):
                    for ln in (cnt * 7, cnt * 40 + 1):
                        pad = b""
                        if flags & 0x40:
                            pad = [b"\x00", b"\x05", b"\xfe", b"\xff\x00", b"\xff\xff\x03", b"\xff"][int(rng.integers(0, 6))]
                        sizes = b""
                        if flags & 0x80 and cnt:
                            sizes = b"".join(_size_bytes(int(v)) for v in rng.choice([0, 1, 7, 40, 251, 252, 300], size=max(cnt - 1, 0)))
                        pk.append(bytes([base | 3, flags | cnt]) + pad + sizes + body[: ln + int(rng.integers(0, 300))])
            pk.append(bytes([base | 3]))
    out = list(pk)
    for m in pk[::9]:                 # truncations and mutations
        if len(m) > 3:
            out.append(m[: len(m) // 2])
            b = bytearray(m)
            for _ in range(2):
                b[int(rng.integers(0, min(len(b), 8)))] = int(rng.integers(0, 256))
            out.append(bytes(b))
    for _ in range(400):
        out.append(bytes(rng.integers(0, 256, int(rng.integers(1, 40)), dtype=np.uint8)))
    return [p for p in out if len(p) <= 4096]
