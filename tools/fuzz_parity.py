#!/usr/bin/env python
"""Randomised differential test: CUDA path (through the C ABI) vs the CPU oracle on seeded channels
with random preset, SNR, clock error, offsets, payload lengths and chunking.  Test infrastructure.
Usage: python tools/fuzz_parity.py [seconds] [first_seed] [preset]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)

import audio_network_b200 as anm  # noqa: E402
from sigutil import make_channels  # noqa: E402
from test_gpu_parity import _check_against_oracle  # noqa: E402

budget = float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
only = sys.argv[3] if len(sys.argv) > 3 else None
t0, n, frames_total = time.time(), 0, 0
while time.time() - t0 < budget:
    rng = np.random.default_rng(seed)
    name = ["ref4", "ref4", "bfsk2", "mfsk8", "mfsk16", "wide64"][int(rng.integers(0, 6))]
    if only:
        name = only
    cfg = anm.config_preset(name)
    if name == "ref4" and rng.integers(0, 4) == 0:
        cfg.hops_per_sym = int(rng.choice([2, 8]))
    n_ch = int(rng.integers(1, 9)) if name != "wide64" else int(rng.integers(1, 40))
    n_sym = int(rng.integers(40, 400))
    snr = [None, 12.0, 6.0, 3.0, 1.0, 0.0][int(rng.integers(0, 6))]
    ppm = float(rng.choice([0.0, 50.0, 200.0]))
    pl = (1, int(rng.integers(2, 60)))
    pcm, _ = make_channels(cfg, n_ch, n_sym * cfg.sym_len, seed=seed, snr_db=snr, ppm_max=ppm,
                           offset_max=int(rng.integers(0, 3000)), payload_len=pl, gap=(1, int(rng.integers(2, 30))),
                           amplitude=float(rng.choice([0.05, 0.5, 0.99])))
    chunks = [int(x) for x in rng.integers(1, 70, size=int(rng.integers(1, 5)))]
    frames = _check_against_oracle(cfg, pcm, chunks)
    frames_total += len(frames)
    n += 1
    seed += 1
print("fuzz ok: %d cases, %d frames, seeds up to %d, %.0f s" % (n, frames_total, seed - 1, time.time() - t0))
