#!/usr/bin/env python
"""Regenerates tests/golden/celt_pcm.npz from the REFERENCE's libopus 1.3.1 (oracle/_ref/libref_opus.so): the int16 PCM the reference's own
celt_decode_with_ec() writes for every frame of tests/golden/celt_entropy.npz (oracle/ref_celt_state_shim.c runs a copy of celt/celt_decoder.c compiled
in place), every stream decoded twice -- by a decoder with as many channels as the stream has (`cc_native`) and by a MONO decoder (`cc_mono`: stereo
frames are downmixed, phase inversion is off) -- plus, for the streams that have mono frames only, by a STEREO decoder (`cc_stereo`: the upmix path).
Per frame a 64-bit digest of its (120 << lm) x CC samples; for the first two frames of every stream the samples themselves.
    python tests/golden/make_celt_pcm_golden.py
"""
import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import celt_spectrum_binding as sbind  # noqa: E402

GOLD = np.load(os.path.join(HERE, "celt_entropy.npz"))


def digest(a):
    return np.frombuffer(hashlib.sha256(np.ascontiguousarray(a, dtype="<i2").tobytes()).digest()[:8], dtype="<u8")[0]


def main():
    R = sbind.ref()
    fr, by, sb = GOLD["frames"], GOLD["bytes"], GOLD["stream_begin"]
    n = len(fr)
    out = {k: np.zeros(n, "<u8") for k in ("cc_native", "cc_mono", "cc_stereo")}
    head_idx, head = [], []
    for s in range(len(sb) - 1):
        js = list(range(sb[s], sb[s + 1]))
        frames = [bytes(by[fr[j]["offset"]: fr[j]["offset"] + fr[j]["len"]]) for j in js]
        params = np.array([[int(fr[j]["channels"]), int(fr[j]["lm"]), int(fr[j]["end_band"])] for j in js], np.int32)
        native = int(params[:, 0].max())
        for key, cc in (("cc_native", native), ("cc_mono", 1), ("cc_stereo", 2)):
            maxlen = max(len(f) for f in frames)
            buf, lens = np.zeros((len(js), maxlen), np.uint8), np.zeros(len(js), np.int32)
            for k, f in enumerate(frames):
                buf[k, :len(f)] = np.frombuffer(f, np.uint8)
                lens[k] = len(f)
            states, pcm = np.zeros(len(js), sbind.STATE), np.zeros((len(js), 960 * cc), np.int16)
            assert R.ref_celt_stream_states(buf.ctypes.data, lens.ctypes.data, len(js), maxlen, params.ctypes.data, cc, states.ctypes.data, pcm.ctypes.data) == len(js)
            for k, j in enumerate(js):
                ns = (120 << int(params[k, 1])) * cc
                assert states[k]["ret"] == 120 << int(params[k, 1])
                out[key][j] = digest(pcm[k, :ns])
                if key == "cc_native" and k < 2:
                    head_idx.append(j)
                    row = np.zeros(1920, np.int16)
                    row[:ns] = pcm[k, :ns]
                    head.append(row)
    np.savez_compressed(os.path.join(HERE, "celt_pcm.npz"), head_idx=np.array(head_idx, np.int32), head=np.array(head, "<i2"), **out)
    print("frames", n, "streams", len(sb) - 1)


if __name__ == "__main__":
    main()
