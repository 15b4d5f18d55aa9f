#!/usr/bin/env python
"""Regenerates tests/golden/celt_spectrum.npz from the REFERENCE's libopus 1.3.1 (oracle/_ref/libref_opus.so) for the frames of
tests/golden/celt_entropy.npz: per frame
  x_digest        64-bit digest of the normalised spectrum celt_synthesis() receives (quant_all_bands + anti_collapse; [2][960] layout, zero above
                  the end band) -- from the reference's own functions driven in decode order (oracle/ref_celt_shim.c), with the noise seed and the
                  log-energy histories READ OFF the reference decoder's private state (oracle/ref_celt_state_shim.c compiles celt_decoder.c in place);
  collapse        the 42 collapse masks; seed_in, log_e1 / log_e2 (the histories before the frame), anti_collapse_on;
and x_full: the complete spectra of the first two frames of every stream.
    python tests/golden/make_celt_spectrum_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import celt_spectrum_binding as sbind  # noqa: E402

GOLD = np.load(os.path.join(HERE, "celt_entropy.npz"))


def main():
    R = sbind.ref()
    fr, by, sb = GOLD["frames"], GOLD["bytes"], GOLD["stream_begin"]
    n = len(fr)
    dig, cm, seed = np.zeros(n, "<u8"), np.zeros((n, 42), np.uint8), np.zeros(n, "<u4")
    e1, e2, ac = np.zeros((n, 42), "<i2"), np.zeros((n, 42), "<i2"), np.zeros(n, np.uint8)
    full_idx, full = [], []
    for s in range(len(sb) - 1):
        js = range(sb[s], sb[s + 1])
        frames = [bytes(by[fr[j]["offset"]: fr[j]["offset"] + fr[j]["len"]]) for j in js]
        params = [[int(fr[j]["channels"]), int(fr[j]["lm"]), int(fr[j]["end_band"])] for j in js]
        cc = max(p[0] for p in params)
        states, xs, cms, acs = sbind.ref_stream(R, frames, params, cc)
        for k, j in enumerate(js):
            assert states[k]["rng_after"] == fr[j]["final_range"]
            dig[j], cm[j], seed[j], e1[j], e2[j], ac[j] = sbind.x_digest(xs[k]), cms[k], states[k]["rng_before"], states[k]["log_e1_before"], states[k]["log_e2_before"], acs[k]
            if k < 2:
                full_idx.append(j)
                full.append(xs[k])
    np.savez_compressed(os.path.join(HERE, "celt_spectrum.npz"), x_digest=dig, collapse=cm, seed_in=seed, log_e1=e1, log_e2=e2, anti_collapse_on=ac,
                        full_idx=np.array(full_idx, np.int32), x_full=np.array(full, "<i2"))
    print("frames", n, "with anti-collapse", int(ac.sum()), "nonzero digests", int((dig != sbind.x_digest(np.zeros(1920, np.int16))).sum()))


if __name__ == "__main__":
    main()
