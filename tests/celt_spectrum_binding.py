"""Bindings for the CELT spectrum tests (row f1, stage 2): the reference's decoder state and spectra through oracle/_ref/libref_opus.so
(oracle/ref_celt_state_shim.c, oracle/ref_celt_shim.c) and the host harness of the product headers (tests/native)."""
import ctypes as C
import hashlib

import numpy as np

import celt_binding as cb

STATE = np.dtype([("rng_before", "<u4"), ("rng_after", "<u4"), ("log_e1_before", "<i2", (42,)), ("log_e2_before", "<i2", (42,)), ("band_e_after", "<i2", (42,)),
                  ("log_e1_after", "<i2", (42,)), ("log_e2_after", "<i2", (42,)), ("ret", "<i4")])                    # ref_celt_state_t
HIST = np.dtype([("log_e1", "<i2", (42,)), ("log_e2", "<i2", (42,)), ("seed", "<u4")])                               # ce_hist_t
STREAM = np.dtype([("old_e", "<i2", (42,)), ("log_e1", "<i2", (42,)), ("log_e2", "<i2", (42,)), ("rng", "<u4"), ("flags", "<u4")])


def ref():
    R = cb.ref()
    R.ref_celt_spectrum_trace2.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.POINTER(cb.RefTrace), C.c_uint32, C.c_int, C.c_void_p,
                                           C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    R.ref_celt_stream_states.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
    return R


def harness():
    L = cb.harness()
    L.harness_celt_frame_full.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                          C.c_void_p]
    L.harness_celt_spectrum.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_uint32, C.c_int, C.c_void_p,
                                        C.c_void_p, C.c_void_p]
    assert L.harness_sizeof_hist() == HIST.itemsize and L.harness_sizeof_stream() == STREAM.itemsize == 260
    return L


def x_digest(x):
    """64-bit digest of a frame's spectrum in the [2][960] comparison layout"""
    return np.frombuffer(hashlib.sha256(np.ascontiguousarray(x, dtype="<i2").tobytes()).digest()[:8], dtype="<u8")[0]


def compare_layout(x_row, channels, lm, end, ebands):
    """the product's [channels][120 << lm] frame -> [2][960] with everything above the end band cleared (what the shims report)"""
    nf, ncmp = 120 << lm, (1 << lm) * int(ebands[end])
    out = np.zeros(1920, np.int16)
    for c in range(channels):
        out[960 * c: 960 * c + ncmp] = x_row[nf * c: nf * c + ncmp]
    return out


def ref_stream(R, frames, params, cc):
    """the reference decoder over one stream: (states, spectra after anti-collapse [n, 1920], collapse masks [n, 42], anti_collapse_on [n])"""
    n = len(frames)
    maxlen = max(len(f) for f in frames)
    buf = np.zeros((n, maxlen), np.uint8)
    lens = np.zeros(n, np.int32)
    for k, f in enumerate(frames):
        buf[k, :len(f)] = np.frombuffer(f, np.uint8)
        lens[k] = len(f)
    params = np.ascontiguousarray(params, np.int32)
    states = np.zeros(n, STATE)
    assert R.ref_celt_stream_states(buf.ctypes.data, lens.ctypes.data, n, maxlen, params.ctypes.data, cc, states.ctypes.data, None) == n
    xs, cms, ac = np.zeros((n, 1920), np.int16), np.zeros((n, 42), np.uint8), np.zeros(n, np.uint8)
    oe = np.zeros(42, np.int16)
    for k in range(n):
        b = buf[k, :lens[k]].copy()
        ch, lm, end = [int(v) for v in params[k]]
        tr, xr, so = cb.RefTrace(), np.zeros(1920, np.int16), C.c_uint32(0)
        p1, p2 = states[k]["log_e1_before"].copy(), states[k]["log_e2_before"].copy()
        assert R.ref_celt_spectrum_trace2(b.ctypes.data, len(b), ch, lm, end, oe.ctypes.data, C.byref(tr), int(states[k]["rng_before"]), 1 if cc == 1 else 0,
                                          xr.ctypes.data, cms[k].ctypes.data, C.byref(so), p1.ctypes.data, p2.ctypes.data, xs[k].ctypes.data) == 0
        assert np.array_equal(oe, states[k]["band_e_after"]) and tr.rng[7] == states[k]["rng_after"]   # the traced sequence IS what the decoder did
        ac[k] = tr.anti_collapse_on
    return states, xs, cms, ac
