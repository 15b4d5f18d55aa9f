"""Bindings for the CELT entropy-decode tests: the host harness of the product header (tests/native), the reference trace shim
(oracle/_ref/libref_opus.so) and helpers that cut golden packets into frames."""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HERE = os.path.join(ROOT, "tests", "native")
BUILD = os.path.join(HERE, "_build")
REF_OPUS = os.path.join(ROOT, "oracle", "_ref", "libref_opus.so")

NB = 21
FRAME_DTYPE = np.dtype([("final_range", "<u4"), ("tell_bits", "<i4"), ("flags", "<u4"), ("pf_pitch", "<u2"), ("pf_gain_q", "u1"), ("pf_tapset", "u1"),
                        ("spread", "u1"), ("alloc_trim", "u1"), ("intensity", "u1"), ("coded_bands", "u1"), ("lm", "u1"), ("channels", "u1"), ("pad", "u1", (2,)),
                        ("pvq_codewords", "<u4"), ("pvq_pulses", "<u4"), ("pvq_index_xor", "<u4"), ("tf_res", "i1", (NB,)), ("fine_quant", "u1", (NB,)),
                        ("pulses", "<i2", (NB,)), ("band_e", "<i2", (2 * NB,))])
JOB_DTYPE = np.dtype([("offset", "<u4"), ("len", "<u4"), ("channels", "u1"), ("lm", "u1"), ("end_band", "u1"), ("flags", "u1")])
TABLES_FIELDS = [("ebands", "<i2", (NB + 1,)), ("logn", "<i2", (NB,)), ("cache_index", "<i2", (5 * NB,)), ("cache_size", "<u2"), ("cache_bits", "u1", (512,)),
                 ("cache_caps", "u1", (4 * 2 * NB,)), ("alloc", "u1", (11 * NB,)), ("e_prob", "u1", (4 * 2 * 42,))]


class RefTrace(C.Structure):
    _fields_ = [("rng", C.c_uint32 * 8), ("tell", C.c_int32 * 8), ("silence", C.c_int32), ("postfilter", C.c_int32), ("pf_pitch", C.c_int32), ("pf_qg", C.c_int32),
                ("pf_tapset", C.c_int32), ("transient", C.c_int32), ("intra", C.c_int32), ("spread", C.c_int32), ("alloc_trim", C.c_int32), ("intensity", C.c_int32),
                ("dual_stereo", C.c_int32), ("coded_bands", C.c_int32), ("anti_collapse_on", C.c_int32), ("balance", C.c_int32), ("tf_res", C.c_int32 * 21),
                ("offsets", C.c_int32 * 21), ("cap", C.c_int32 * 21), ("pulses", C.c_int32 * 21), ("fine_quant", C.c_int32 * 21), ("fine_priority", C.c_int32 * 21),
                ("band_e", C.c_int16 * 42)]


_h = None


def harness():
    """builds (if stale) and loads the host harness of the product's header"""
    global _h
    if _h is not None:
        return _h
    os.makedirs(BUILD, exist_ok=True)
    so = os.path.join(BUILD, "libcelt_harness.so")
    srcs = [os.path.join(HERE, "celt_harness.c"), os.path.join(ROOT, "audio-network_b200", "csrc", "anm_celt_tables.c")]
    deps = srcs + [os.path.join(ROOT, "audio-network_b200", "csrc", h) for h in ("anm_celt_entropy.h", "anm_celt_vec.h", "anm_celt_synth.h")] + \
        [os.path.join(ROOT, "include", "anmodem_opus.h")]
    if not os.path.exists(so) or any(os.path.getmtime(d) > os.path.getmtime(so) for d in deps):
        subprocess.check_call(["gcc", "-O2", "-fPIC", "-shared", "-std=gnu11", "-Wall", "-o", so] + srcs + ["-lm"])
    L = C.CDLL(so)
    L.harness_celt_frame.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    L.anm_celt_tables_build.argtypes = [C.c_void_p]
    _h = L
    return L


_t = None


def tables():
    global _t
    if _t is None:
        L = harness()
        buf = np.zeros(L.harness_sizeof_tables(), dtype=np.uint8)
        assert L.anm_celt_tables_build(buf.ctypes.data) == 0
        _t = buf
    return _t


def ref():
    R = C.CDLL(REF_OPUS)
    R.ref_celt_entropy_trace.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.POINTER(RefTrace)]
    R.ref_celt_mode_tables.argtypes = [C.c_void_p] * 6
    R.ref_opus_frames_final_range.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_uint8, C.c_void_p, C.c_int]
    return R


BW_END = {1101: 13, 1102: 17, 1103: 17, 1104: 19, 1105: 21}   # opus_decoder.c:462-481 (medium band is not a CELT bandwidth)


def frames_of_packet(packet, parse):
    """[(frame bytes, channels, LM, end_band)] of one CELT-only packet, from its reference parse record"""
    out = []
    off = parse["payload_offset"]
    lm = {120: 0, 240: 1, 480: 2, 960: 3}[parse["samples_per_frame"]]
    for i in range(parse["count"]):
        n = parse["size"][i]
        out.append((bytes(packet[off: off + n]), parse["channels"], lm, BW_END[parse["bandwidth"]]))
        off += n
    return out


def decode_frame(frame, ch, lm, end, old_e):
    out = np.zeros(1, dtype=FRAME_DTYPE)
    b = np.frombuffer(frame, dtype=np.uint8) if len(frame) else np.zeros(1, dtype=np.uint8)
    rc = harness().harness_celt_frame(tables().ctypes.data, b.ctypes.data, len(frame), ch, lm, end, old_e.ctypes.data, out.ctypes.data)
    return rc, out[0]
