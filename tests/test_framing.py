"""Transmit-side framing (SPEC 4) and the twiddle table (SPEC 3)."""
import numpy as np
import pytest

import audio_network_b200 as anm


def test_crc_known_answers():
    # CRC-16/CCITT-FALSE and CRC-8 (poly 0x07) check values of the ASCII string "123456789"
    assert anm.crc16(b"123456789") == 0x29B1
    assert anm.crc8(b"123456789") == 0xF4
    assert anm.crc16(b"") == 0xFFFF
    # incremental == one shot
    assert anm.crc16(b"6789", anm.crc16(b"12345")) == 0x29B1


@pytest.mark.parametrize("name", ["ref4", "bfsk2", "mfsk8", "mfsk16", "wide64"])
def test_frame_symbols_layout(name):
    cfg = anm.config_preset(name)
    b = cfg.bits_per_sym
    payload = bytes([0xA5, 0x00, 0xFF, 0x3C, 0x81])
    syms = anm.frame_symbols(cfg, payload)
    P = cfg.preamble_len
    hdr = -(-24 // b)
    body = -(-(len(payload) + 2) * 8 // b)
    assert len(syms) == P + hdr + body
    assert list(syms[:P]) == list(cfg.preamble[:P])
    assert syms.max() < cfg.n_tones

    def unpack(tones, nbytes):
        bits = []
        for g in tones:
            v = int(g)
            s = 1
            while s < 8:
                v ^= v >> s
                s <<= 1
            bits += [(v >> (b - 1 - i)) & 1 for i in range(b)]
        out = bytearray()
        for i in range(nbytes):
            x = 0
            for k in range(8):
                x = (x << 1) | bits[i * 8 + k]
            out.append(x)
        return bytes(out)

    h = unpack(syms[P:P + hdr], 3)
    assert (h[0] << 8 | h[1]) == len(payload)
    assert h[2] == anm.crc8(h[:2])
    bd = unpack(syms[P + hdr:], len(payload) + 2)
    assert bd[:len(payload)] == payload
    assert (bd[-2] << 8 | bd[-1]) == anm.crc16(payload, anm.crc16(h[:2]))


def test_frame_symbols_rejects_bad_lengths():
    cfg = anm.config_preset("ref4")
    with pytest.raises(anm.AnmError):
        anm.frame_symbols(cfg, b"")
    with pytest.raises(anm.AnmError):
        anm.frame_symbols(cfg, bytes(cfg.max_payload + 1))
    assert len(anm.frame_symbols(cfg, bytes(cfg.max_payload))) > 0


@pytest.mark.parametrize("name", ["ref4", "bfsk2", "mfsk16", "wide64"])
def test_twiddle_table_values_and_symmetry(name):
    cfg = anm.config_preset(name)
    tw = anm.twiddles(cfg)
    N, T = cfg.sym_len, cfg.n_tones
    m = np.arange(N)[:, None]
    bins = np.array(cfg.tone_bin[:T])[None, :]
    ang = 2 * np.pi * ((bins * m) % N) / N
    assert np.allclose(tw[:, :, 0], np.cos(ang), atol=1e-6)
    assert np.allclose(tw[:, :, 1], np.sin(ang), atol=1e-6)
    # exact quarter-period symmetry (SPEC 3): (c - js)[m + N/4] = (-j)^b (c - js)[m]
    q = N // 4
    for k in range(T):
        r = cfg.tone_bin[k] % 4
        c0, s0 = tw[:q, k, 0], tw[:q, k, 1]
        c1, s1 = tw[q:2 * q, k, 0], tw[q:2 * q, k, 1]
        exp_c = [c0, -s0, -c0, s0][r]
        exp_s = [s0, c0, -s0, -c0][r]
        assert np.array_equal(c1, exp_c) and np.array_equal(s1, exp_s)
