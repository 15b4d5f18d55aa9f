#!/usr/bin/env python
"""Randomised differential test of the CELT decode logic (stages 1-3) on the HOST: the product headers compiled by the test harness (tests/native) against the
reference decoder itself (oracle/_ref/libref_opus.so, its own celt_decode_with_ec through oracle/ref_celt_state_shim.c): random bytes as frames, in streams
whose channel count, frame size and bandwidth may change from frame to frame, decoded by a mono or a stereo decoder; every int16 sample must be equal.
Test infrastructure (needs /root/reference to have been present when oracle/_ref was built).  Usage: python tools/celt_fuzz.py [seed] [streams]"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, os.path.join(ROOT, 'tests')); sys.path.insert(0, ROOT)
import celt_binding as cb, celt_spectrum_binding as sbind
import audio_network_b200 as anm
L = sbind.harness(); R = sbind.ref()
L.anm_celt_synth_tables_build.argtypes = [C.c_void_p]
L.harness_celt_decode_frame.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
stb = np.zeros(1, anm.CELT_SYNTH_TABLES_DTYPE); assert L.anm_celt_synth_tables_build(stb.ctypes.data) == 0
t = cb.tables()
rng = np.random.default_rng(int(sys.argv[1]) if len(sys.argv) > 1 else 1)
nstreams = int(sys.argv[2]) if len(sys.argv) > 2 else 300
bad = 0; tot = 0
for s in range(nstreams):
    cc = int(rng.integers(1, 3)); nfr = int(rng.integers(2, 7))
    params = np.array([[int(rng.integers(1, 3)), int(rng.integers(0, 4)), int(rng.choice([13, 17, 19, 21]))] for _ in range(nfr)], np.int32)
    if rng.random() < 0.6: params[:] = params[0]
    frames = [rng.integers(0, 256, int(rng.choice([rng.integers(2, 12), rng.integers(12, 120), rng.integers(120, 500), rng.integers(500, 1276)])), dtype=np.uint8) for _ in range(nfr)]
    for f in frames:
        if rng.random() < 0.7: f[0] &= 0x7F
    maxlen = max(len(f) for f in frames)
    buf = np.zeros((nfr, maxlen), np.uint8); lens = np.array([len(f) for f in frames], np.int32)
    for k, f in enumerate(frames): buf[k, :len(f)] = f
    states = np.zeros(nfr, sbind.STATE); pcm_ref = np.zeros((nfr, 960 * cc), np.int16)
    assert R.ref_celt_stream_states(buf.ctypes.data, lens.ctypes.data, nfr, maxlen, params.ctypes.data, cc, states.ctypes.data, pcm_ref.ctypes.data) == nfr
    st = np.zeros(1, anm.CELT_STREAM_DTYPE); syn = np.zeros(1, anm.CELT_SYNTH_DTYPE)
    for k in range(nfr):
        ch, lm, end = [int(v) for v in params[k]]
        b = frames[k].copy(); out = np.zeros(1, cb.FRAME_DTYPE); pcm = np.zeros(960 * cc, np.int16)
        assert L.harness_celt_decode_frame(t.ctypes.data, stb.ctypes.data, b.ctypes.data, len(b), ch, cc, lm, end, st.ctypes.data, syn.ctypes.data, out.ctypes.data, pcm.ctypes.data) == 0
        n = (120 << lm) * cc; tot += 1
        if states[k]["ret"] < 0: continue
        if not np.array_equal(pcm[:n], pcm_ref[k, :n]):
            bad += 1
            if bad <= 5:
                d = np.nonzero(pcm[:n] != pcm_ref[k, :n])[0]
                print("stream", s, "frame", k, "C", ch, "CC", cc, "lm", lm, "end", end, "len", len(b), "ret", states[k]["ret"], "flags", hex(int(out[0]["flags"])), "ndiff", len(d), "first", d[:3], pcm[d[:3]], pcm_ref[k, d[:3]])
print("frames", tot, "bad", bad)
