"""Synthetic channel construction shared by tests, smoke() and bench.py (host side)."""
import numpy as np

import audio_network_b200 as anm


def make_program(cfg, rng, n_symbols, payload_len=(4, 40), gap=(2, 24), lead=(0, 12)):
    """A symbol program of at least n_symbols entries: [gap][frame][gap][frame]...

    Returns (program uint8 array, [payload bytes, ...])."""
    prog = [np.full(int(rng.integers(lead[0], lead[1] + 1)), anm.ANM_SILENCE, dtype=np.uint8)]
    payloads = []
    total = len(prog[0])
    while total < n_symbols:
        ln = int(rng.integers(payload_len[0], payload_len[1] + 1))
        pl = rng.integers(0, 256, size=ln, dtype=np.uint8).tobytes()
        syms = anm.frame_symbols(cfg, pl)
        g = np.full(int(rng.integers(gap[0], gap[1] + 1)), anm.ANM_SILENCE, dtype=np.uint8)
        prog += [syms, g]
        payloads.append(pl)
        total += len(syms) + len(g)
    return np.concatenate(prog), payloads


def make_channels(cfg, n_ch, n_samples, seed=1, snr_db=None, ppm_max=0.0, amplitude=0.5, offset_max=0,
                  payload_len=(4, 40), gap=(2, 24)):
    """pcm[n_ch, n_samples] int16 rendered by the CPU transmitter stand-in, plus per-channel
    (program, payloads, TxParams)."""
    rng = np.random.default_rng(seed)
    pcm = np.zeros((n_ch, n_samples), dtype=np.int16)
    meta = []
    nsym = n_samples // cfg.sym_len + 4
    for c in range(n_ch):
        prog, payloads = make_program(cfg, rng, nsym, payload_len, gap)
        p = anm.tx_params(
            seed=seed * 1000003 + c,
            start_offset=-int(rng.integers(0, offset_max + 1)),
            amplitude=amplitude,
            snr_db=snr_db,
            ppm=float(rng.uniform(-ppm_max, ppm_max)) if ppm_max else 0.0,
        )
        pcm[c] = anm.tx_render(cfg, prog, p, 0, n_samples)
        meta.append((prog, payloads, p))
    return pcm, meta
