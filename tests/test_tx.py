"""Transmitter stand-in (SPEC 6), CPU renderer."""
import numpy as np

import audio_network_b200 as anm


def test_render_is_deterministic_and_chunkable():
    cfg = anm.config_preset("ref4")
    prog = np.array([0, 1, 2, 3, 255, 2, 1], dtype=np.uint8)
    p = anm.tx_params(seed=5, start_offset=-100, amplitude=0.4, snr_db=6.0, ppm=123.456)
    whole = anm.tx_render(cfg, prog, p, 0, 5000)
    again = anm.tx_render(cfg, prog, p, 0, 5000)
    assert np.array_equal(whole, again)
    parts = np.concatenate([anm.tx_render(cfg, prog, p, 0, 1234), anm.tx_render(cfg, prog, p, 1234, 5000 - 1234)])
    assert np.array_equal(whole, parts)


def test_silence_and_leading_offset():
    cfg = anm.config_preset("ref4")
    p = anm.tx_params(start_offset=-300, amplitude=0.5)
    x = anm.tx_render(cfg, np.array([1], dtype=np.uint8), p, 0, 1000)
    assert not x[:300].any() and x[300:].any()
    y = anm.tx_render(cfg, np.array([255], dtype=np.uint8), anm.tx_params(amplitude=0.5), 0, 1000)
    assert not y.any()


def test_tone_frequency_amplitude_and_phase_continuity():
    cfg = anm.config_preset("ref4")
    N = cfg.sym_len
    x = anm.tx_render(cfg, np.array([2], dtype=np.uint8), anm.tx_params(amplitude=0.5), 0, 8 * N).astype(np.float64)
    spec = np.abs(np.fft.rfft(x[:N]))
    assert int(np.argmax(spec)) == cfg.tone_bin[2]
    assert abs(np.max(np.abs(x)) - 16384) <= 2
    # integer cycles per symbol: every symbol starts at phase 0, so the waveform is N-periodic
    assert np.max(np.abs(x[:N] - x[N:2 * N])) <= 1


def test_snr_is_roughly_as_requested():
    cfg = anm.config_preset("ref4")
    n = 200000
    clean = anm.tx_render(cfg, np.array([1], dtype=np.uint8), anm.tx_params(amplitude=0.25), 0, n).astype(np.float64)
    for snr in (0.0, 10.0, 20.0):
        noisy = anm.tx_render(cfg, np.array([1], dtype=np.uint8), anm.tx_params(seed=3, amplitude=0.25, snr_db=snr), 0, n).astype(np.float64)
        noise = noisy - clean
        got = 10 * np.log10(np.mean(clean ** 2) / np.mean(noise ** 2))
        assert abs(got - snr) < 0.3, (snr, got)
        assert abs(np.mean(noise)) < 20


def test_clock_error_moves_the_tone():
    cfg = anm.config_preset("ref4")
    n = 1 << 18
    x = anm.tx_render(cfg, np.array([0], dtype=np.uint8), anm.tx_params(amplitude=0.5, ppm=200.0), 0, n).astype(np.float64)
    f = np.fft.rfftfreq(n, 1 / 44100.0)
    peak = f[np.argmax(np.abs(np.fft.rfft(x * np.hanning(n))))]
    nominal = cfg.tone_bin[0] * 44100.0 / cfg.sym_len
    assert abs(peak - nominal * (1 + 200e-6)) < 0.2
