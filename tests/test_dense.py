"""SPEC 3b (dense tone sets, T >= 32): the int8 basis, the exact integer oracle path and -- on a
GPU -- the tcgen05 contraction kernel k_demod_tc against it.  The CPU part needs no GPU."""
import numpy as np
import pytest

import audio_network_b200 as anm
from oracle_binding import Oracle
from sigutil import make_channels


def test_dense_rule_and_basis_symmetry():
    assert anm.config_dense(anm.config_preset("wide64"))
    for name in ("ref4", "bfsk2", "mfsk8", "mfsk16"):
        assert not anm.config_dense(anm.config_preset(name))
    cfg = anm.config_preset("wide64")
    b = anm.basis_q7(cfg).astype(np.int32)
    N, T = cfg.sym_len, cfg.n_tones
    assert b.shape == (N, T, 2) and np.abs(b).max() == 127
    # first quarter = round(127 cos / sin) of the exactly reduced angle
    m = np.arange(N // 4)[:, None]
    bins = np.array(cfg.tone_bin[:T])[None, :]
    ang = 2 * np.pi * ((bins * m) % N) / N
    assert np.array_equal(b[: N // 4, :, 0], np.rint(127 * np.cos(ang)).astype(np.int32))
    assert np.array_equal(b[: N // 4, :, 1], np.rint(127 * np.sin(ang)).astype(np.int32))
    # other quarters: (c - j s)[m + q N/4] = (-j)^(bin q) (c - j s)[m], an exact swap / negation
    z0 = b[: N // 4, :, 0] - 1j * b[: N // 4, :, 1]
    for q in range(1, 4):
        zq = b[q * N // 4: (q + 1) * N // 4, :, 0] - 1j * b[q * N // 4: (q + 1) * N // 4, :, 1]
        assert np.array_equal(zq, z0 * (-1j) ** ((bins * q) % 4))


def test_dense_energies_are_exact_integer_window_sums():
    cfg = anm.config_preset("wide64")
    N, S, T, H = cfg.sym_len, cfg.hops_per_sym, cfg.n_tones, cfg.hop
    rng = np.random.default_rng(5)
    x = rng.integers(-32768, 32768, size=6 * N).astype(np.int16)   # full-scale noise: the largest sums
    x[:N] = 32767
    x[N: 2 * N] = -32768
    hops = len(x) // H
    o = Oracle(cfg, trace_hops=hops)
    o.feed(x)
    b = anm.basis_q7(cfg).astype(np.int64)
    xx = np.concatenate([np.zeros(N, dtype=np.int64), x.astype(np.int64)])
    for h in range(hops):
        win = xx[N + (h + 1) * H - N: N + (h + 1) * H]               # the S hops ending with hop h
        m = (np.arange((h + 1) * H - N, (h + 1) * H)) % N
        I = (win[:, None] * b[m, :, 0]).sum(axis=0)
        Q = (win[:, None] * b[m, :, 1]).sum(axis=0)
        assert np.abs(I).max() < 2 ** 31 and np.abs(Q).max() < 2 ** 31
        fi, fq = I.astype(np.float32), Q.astype(np.float32)
        E = (fi.astype(np.float64) ** 2 + (fq * fq).astype(np.float64)).astype(np.float32)   # fma(fi, fi, fq*fq): one rounding
        assert np.array_equal(o.E[h].view(np.uint32), E.view(np.uint32))
        assert o.D[h] == int(np.argmax(o.E[h])) and o.Emax[h] == o.E[h].max()


def test_dense_oracle_decodes_and_is_feed_size_invariant():
    cfg = anm.config_preset("wide64")
    pcm, meta = make_channels(cfg, 2, 260 * cfg.sym_len, seed=31, snr_db=6.0, ppm_max=120.0, offset_max=900)
    for c in range(2):
        a = Oracle(cfg)
        a.feed(pcm[c])
        fr = a.frames(c)
        assert len(fr) > 0 and all(f[2] == 1 and f[3] in meta[c][1] for f in fr)
        b = Oracle(cfg)
        pos, sizes, i = 0, [1, 63, 64, 65, 1000, 257], 0
        while pos < pcm.shape[1]:
            ln = min(sizes[i % len(sizes)], pcm.shape[1] - pos)
            b.feed(pcm[c, pos:pos + ln])
            pos += ln
            i += 1
        assert b.frames(c) == fr and np.array_equal(a.symbols(), b.symbols())


# ------------------------------------------------------------------ GPU: k_demod_tc vs the oracle
def _gpu_vs_oracle(cfg, pcm, chunks):
    from test_gpu_parity import _check_against_oracle

    return _check_against_oracle(cfg, pcm, chunks)


@pytest.mark.gpu
@pytest.mark.parametrize("chunks", [[1], [5, 1, 40, 7], [32], [33, 31], [250]])
def test_dense_chunking_invariance_noisy_drift(chunks):
    cfg = anm.config_preset("wide64")
    # 10 channels: two full CTAs of four channels and a half-empty one
    pcm, _ = make_channels(cfg, 10, 300 * cfg.sym_len, seed=37, snr_db=4.0, ppm_max=200.0, offset_max=1500)
    frames = _gpu_vs_oracle(cfg, pcm, chunks)
    assert sum(f[2] for f in frames) >= 10


@pytest.mark.gpu
def test_dense_extreme_inputs():
    cfg = anm.config_preset("wide64")
    rng = np.random.default_rng(2)
    pcm = np.zeros((5, 150 * cfg.sym_len), dtype=np.int16)
    pcm[1] = rng.integers(-32768, 32768, size=pcm.shape[1])          # full-scale noise
    pcm[2] = 32767
    pcm[3] = -32768
    pcm[4, ::2], pcm[4, 1::2] = 32767, -32768                        # Nyquist square wave
    _gpu_vs_oracle(cfg, pcm, [37])


@pytest.mark.gpu
def test_dense_long_frames():
    cfg = anm.config_preset("wide64")
    pcm, _ = make_channels(cfg, 3, 1800 * cfg.sym_len, seed=41, snr_db=12.0, payload_len=(900, 1024), gap=(1, 5))
    frames = _gpu_vs_oracle(cfg, pcm, [344])
    assert any(len(f[3]) >= 900 and f[2] for f in frames)


@pytest.mark.gpu
def test_dense_config4_at_bench_scale_all_channels_vs_oracle():
    """BASELINE.json configs[3] at the size bench.py's cfg4 leg runs: 4,736 channels (= 148 SMs x 2 resident CTAs x 4
    channels x 4 waves of CTAs) x 2 chunks of 352 symbol periods through k_demod_tc, 10 dB SNR, with clock error and random
    offsets.  Every frame of every channel equals the oracle's: count, CRC verdicts, payload bytes and the
    order-independent digest of oracle/anm_oracle_batch.c; 37 channels also frame for frame."""
    import os

    import torch

    from oracle_binding import frames_digest, oracle_frames_batch, run_batch
    from test_gpu_parity import _gpu_render, _programs

    cfg = anm.config_preset("wide64")
    n_ch, chunk = 4736, 352 * cfg.sym_len
    n = 2 * chunk
    progs, lens, params = _programs(cfg, n_ch, seed=909, snr_db=10.0, ppm_max=150.0, offset_max=3000, payload=(16, 48), max_len=1024)
    d_pcm = _gpu_render(cfg, progs, lens, params, n)
    dm = anm.Demod(cfg, n_ch, device=0)
    stream = torch.cuda.current_stream().cuda_stream
    dm.feed_device(d_pcm.data_ptr(), n, chunk, stream)
    dm.feed_device(d_pcm.data_ptr() + chunk * 2, n, chunk, stream)
    dm.collect()
    fr = anm.frames_to_list(*dm.read_frames(cap=1 << 20, bytes_cap=1 << 26))
    assert not dm.overflowed()
    dm.close()
    pcm = d_pcm.cpu().numpy()
    _sec, ok, bad, nbytes, dg = run_batch(cfg, pcm, os.cpu_count() or 1)
    assert len(fr) == ok + bad and ok > 2 * n_ch
    assert sum(f[2] for f in fr) == ok
    assert sum(len(f[3]) for f in fr if f[2]) == nbytes
    assert frames_digest(fr) == dg
    sample = list(range(0, n_ch, 128))
    want = oracle_frames_batch(cfg, pcm[sample])
    got = [(sample.index(f[0]), f[1], f[2], f[3]) for f in fr if f[0] in set(sample)]
    assert got == want
