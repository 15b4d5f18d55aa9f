"""GPU parity tests proper: the CUDA path (through the C ABI) against the CPU oracle on the
same seeded PCM.  Bit-exact for symbols, frames, CRC verdicts; tone energies are checked
bit-exact AND against north_star's stated tolerance (1e-4 relative, fp32)."""
import numpy as np
import pytest

import audio_network_b200 as anm
from oracle_binding import Oracle, oracle_frames_batch
from sigutil import make_channels

pytestmark = pytest.mark.gpu

PRESETS = ["ref4", "bfsk2", "mfsk8", "mfsk16", "wide64"]


def _torch():
    import torch

    assert torch.cuda.is_available()
    return torch


def _variant(name, S=None, N=None):
    cfg = anm.config_preset(name)
    if S is not None:
        cfg.hops_per_sym = S
    if N is not None:
        cfg.sym_len = N
    return cfg


@pytest.mark.parametrize("preset", PRESETS)
def test_tone_energies_match_oracle(preset):
    torch = _torch()
    cfg = anm.config_preset(preset)
    n_ch, n_sym = 6, 75  # ragged last step (75 = 2*32 + 11)
    n = n_sym * cfg.sym_len
    pcm, _ = make_channels(cfg, n_ch, n, seed=3, snr_db=6.0, offset_max=300)
    hops = n // cfg.hop
    d_pcm = torch.from_numpy(pcm).cuda()
    dE = torch.zeros((n_ch, hops, cfg.n_tones), dtype=torch.float32, device="cuda")
    dD = torch.zeros((n_ch, hops), dtype=torch.uint8, device="cuda")
    dM = torch.zeros((n_ch, hops), dtype=torch.float32, device="cuda")
    anm.tone_energies_device(cfg, d_pcm.data_ptr(), n_ch, n, n, dE.data_ptr(), dD.data_ptr(), dM.data_ptr())
    torch.cuda.synchronize()
    E, D, M = dE.cpu().numpy(), dD.cpu().numpy(), dM.cpu().numpy()
    for c in range(n_ch):
        o = Oracle(cfg, trace_hops=hops)
        o.feed(pcm[c])
        # stated contract: 1e-4 relative in fp32
        denom = np.maximum(np.abs(o.E), 1e-30)
        assert np.max(np.abs(E[c] - o.E) / denom) <= 1e-4
        # achieved: identical bits (SPEC 3 pins the operation order)
        assert np.array_equal(E[c].view(np.uint32), o.E.view(np.uint32))
        assert np.array_equal(D[c], o.D)
        assert np.array_equal(M[c].view(np.uint32), o.Emax.view(np.uint32))


def _run_gpu(cfg, pcm, chunks=None, flags=anm.ANM_FLAG_SYMBOLS):
    torch = _torch()
    n_ch, n = pcm.shape
    d_pcm = torch.from_numpy(pcm).cuda()
    dm = anm.Demod(cfg, n_ch, device=0, flags=flags)
    pos = 0
    chunks = chunks or [n]
    i = 0
    while pos < n:
        ln = min(chunks[i % len(chunks)] * cfg.sym_len, n - pos)
        dm.feed_device(d_pcm.data_ptr() + pos * 2, n, ln, torch.cuda.current_stream().cuda_stream)
        pos += ln
        i += 1
    dm.collect()
    frames = anm.frames_to_list(*dm.read_frames())
    syms = [dm.read_symbols(c) for c in range(n_ch)] if flags & anm.ANM_FLAG_SYMBOLS else None
    stats = dm.stats()
    dm.close()
    return frames, syms, stats


def _check_against_oracle(cfg, pcm, chunks=None):
    frames, syms, stats = _run_gpu(cfg, pcm, chunks)
    want = []
    for c in range(pcm.shape[0]):
        o = Oracle(cfg)
        o.feed(pcm[c])
        want.extend(o.frames(c))
        assert np.array_equal(syms[c], o.symbols()), "decided symbols differ on channel %d" % c
        st = o.stats()
        for k in ("locks", "header_fail", "frames_ok", "frames_bad", "symbols", "trk_moves"):
            assert int(stats[c][k]) == int(st[k]), (c, k, int(stats[c][k]), int(st[k]))
    want.sort(key=lambda f: (f[0], f[1]))
    assert frames == want
    return frames


@pytest.mark.parametrize("preset", PRESETS)
def test_frames_bit_exact_clean(preset):
    cfg = anm.config_preset(preset)
    pcm, meta = make_channels(cfg, 24, 700 * cfg.sym_len, seed=5, offset_max=2000)
    frames = _check_against_oracle(cfg, pcm)
    assert len(frames) >= 24
    assert all(f[2] == 1 for f in frames)
    # every frame fully inside the capture decodes to the payload that was sent
    sent = {c: meta[c][1] for c in range(24)}
    for ch, _start, _ok, payload in frames:
        assert payload in sent[ch]


def test_frames_bit_exact_noisy_drift():
    cfg = anm.config_preset("ref4")
    pcm, _ = make_channels(cfg, 48, 900 * cfg.sym_len, seed=9, snr_db=2.0, ppm_max=200.0, offset_max=4000)
    frames = _check_against_oracle(cfg, pcm)
    assert sum(f[2] for f in frames) > 40


def test_noise_only_and_silence():
    cfg = anm.config_preset("ref4")
    rng = np.random.default_rng(1)
    pcm = np.zeros((4, 200 * cfg.sym_len), dtype=np.int16)
    pcm[1] = rng.integers(-3000, 3000, size=pcm.shape[1])
    pcm[2] = 32767
    pcm[3] = -32768
    _check_against_oracle(cfg, pcm)


@pytest.mark.parametrize("chunks", [[1], [3, 1, 40, 7], [32], [33, 31], [250]])
def test_chunking_invariance(chunks):
    cfg = anm.config_preset("ref4")
    pcm, _ = make_channels(cfg, 16, 500 * cfg.sym_len, seed=13, snr_db=8.0, ppm_max=150.0, offset_max=1000)
    _check_against_oracle(cfg, pcm, chunks)


@pytest.mark.parametrize("S", [2, 8])
def test_other_hop_counts(S):
    cfg = _variant("ref4", S=S)
    pcm, _ = make_channels(cfg, 12, 400 * cfg.sym_len, seed=17, snr_db=10.0, ppm_max=100.0, offset_max=700)
    frames = _check_against_oracle(cfg, pcm, [37])
    assert len(frames) > 0


def test_short_symbols():
    cfg = _variant("ref4", N=64)
    cfg.tone_bin[0], cfg.tone_bin[1], cfg.tone_bin[2], cfg.tone_bin[3] = 5, 7, 9, 11
    pcm, _ = make_channels(cfg, 12, 600 * cfg.sym_len, seed=19, snr_db=12.0, offset_max=300)
    frames = _check_against_oracle(cfg, pcm, [50, 9])
    assert len(frames) > 0


def test_long_frames_and_max_payload():
    cfg = anm.config_preset("ref4")
    pcm, meta = make_channels(cfg, 4, 9000 * cfg.sym_len, seed=23, payload_len=(900, 1024), gap=(1, 5))
    frames = _check_against_oracle(cfg, pcm, [1000])
    assert any(len(f[3]) >= 900 and f[2] for f in frames)


def test_feed_host_equals_feed_device():
    cfg = anm.config_preset("ref4")
    pcm, _ = make_channels(cfg, 32, 300 * cfg.sym_len, seed=29, snr_db=10.0)
    dm = anm.Demod(cfg, 32, device=0)
    dm.feed_host(pcm[:, : 100 * cfg.sym_len])
    dm.feed_host(np.ascontiguousarray(pcm[:, 100 * cfg.sym_len:]))
    dm.collect()
    got = anm.frames_to_list(*dm.read_frames())
    dm.close()
    assert got == oracle_frames_batch(cfg, pcm)


def test_alignment_and_argument_errors():
    torch = _torch()
    cfg = anm.config_preset("ref4")
    dm = anm.Demod(cfg, 2, device=0)
    buf = torch.zeros(4096, dtype=torch.int16, device="cuda")
    with pytest.raises(anm.AnmError) as e:
        dm.feed_device(buf.data_ptr(), 2048, 100, 0)  # not a multiple of sym_len
    assert e.value.code == anm.ANM_ERR_ALIGN
    with pytest.raises(anm.AnmError) as e:
        dm.feed_device(buf.data_ptr() + 2, 2048, 128, 0)  # misaligned base
    assert e.value.code == anm.ANM_ERR_ALIGN
    dm.close()
    bad = anm.config_preset("ref4")
    bad.sym_len = 96
    with pytest.raises(anm.AnmError):
        anm.Demod(bad, 2, device=0)


def test_tx_render_gpu_equals_cpu():
    torch = _torch()
    cfg = anm.config_preset("ref4")
    rng = np.random.default_rng(31)
    n_ch, n = 5, 40 * cfg.sym_len + 13
    plist, progs = [], []
    for c in range(n_ch):
        progs.append(rng.integers(0, 5, size=37 + c).astype(np.uint8))
        progs[-1][progs[-1] == 4] = anm.ANM_SILENCE
        plist.append(anm.tx_params(seed=100 + c, start_offset=-int(rng.integers(0, 900)) + 300 * (c == 0),
                                   amplitude=0.3 + 0.1 * c, snr_db=[None, 10.0, 0.0, 3.0, 20.0][c],
                                   ppm=[0.0, 200.0, -200.0, 37.5, -0.001][c]))
    stride = 64
    P = np.zeros((n_ch, stride), dtype=np.uint8)
    for c in range(n_ch):
        P[c, : len(progs[c])] = progs[c]
    d_prog = torch.from_numpy(P).cuda()
    d_len = torch.tensor([len(p) for p in progs], dtype=torch.int32, device="cuda")
    d_par = torch.from_numpy(anm.tx_params_array(plist).view(np.uint8)).cuda()
    first = 12345
    ch_stride = ((n + 7) // 8) * 8
    d_pcm = torch.zeros((n_ch, ch_stride), dtype=torch.int16, device="cuda")
    anm.tx_render_device(cfg, d_prog.data_ptr(), stride, d_len.data_ptr(), d_par.data_ptr(), n_ch, first, d_pcm.data_ptr(), ch_stride, n)
    torch.cuda.synchronize()
    got = d_pcm.cpu().numpy()[:, :n]
    for c in range(n_ch):
        want = anm.tx_render(cfg, progs[c], plist[c], first, n)
        assert np.array_equal(got[c], want), "channel %d differs" % c


def test_single_channel_firmware_interface():
    import ctypes as C

    cfg = anm.config_preset("ref4")
    pcm, meta = make_channels(cfg, 1, 400 * cfg.sym_len, seed=37, snr_db=10.0, offset_max=100)
    L = anm.lib()
    assert L.demod_initialize(C.byref(cfg)) == 0
    d = L.demod_create()
    assert d
    # arbitrary feed sizes, as firmware would hand over DMA buffers
    pos, sizes, i = 0, [441, 1000, 77, 4096, 5], 0
    x = pcm[0]
    while pos < len(x):
        ln = min(sizes[i % len(sizes)], len(x) - pos)
        seg = np.ascontiguousarray(x[pos: pos + ln])
        assert L.demod_feed(d, seg.ctypes.data, ln) == 0
        pos += ln
        i += 1

    class DF(C.Structure):
        _fields_ = [("sample_offset", C.c_uint64), ("len", C.c_uint32), ("crc_ok", C.c_uint32), ("bytes", C.c_uint8 * 4104)]

    out = (DF * 64)()
    n = L.demod_read_frames(d, out, 64)
    got = [(0, int(out[i].sample_offset), int(out[i].crc_ok), bytes(out[i].bytes[: out[i].len])) for i in range(n)]
    want = oracle_frames_batch(cfg, pcm[:, : (len(x) // cfg.sym_len) * cfg.sym_len])
    assert got == want and len(got) > 0
    syms = np.zeros(1 << 16, dtype=np.uint8)
    ns = L.demod_read_symbols(d, syms.ctypes.data, len(syms))
    o = Oracle(cfg)
    o.feed(pcm[0, : (len(x) // cfg.sym_len) * cfg.sym_len])
    assert np.array_equal(syms[:ns], o.symbols())
    L.demod_destroy(d)


def test_gpu_matches_committed_golden_vectors():
    """The CUDA path against tests/golden/modem_kat.npz directly (no oracle involved)."""
    import json
    import os

    gold = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "modem_kat.npz"))
    for name in ("ref4", "bfsk2", "mfsk16", "wide64"):
        cfg = anm.config_preset(name)
        pcm = np.ascontiguousarray(gold[name + "_pcm"])
        frames, syms, _ = _run_gpu(cfg, pcm, [50, 13])
        want = [(r[0], r[1], r[2], bytes.fromhex(r[3])) for r in json.loads(str(gold[name + "_frames"]))]
        assert frames == want, name
        want_syms = json.loads(str(gold[name + "_symbols"]))
        for c in range(pcm.shape[0]):
            assert syms[c].tolist() == want_syms[c]


def test_pb_istream_seam_on_gpu_frames():
    """demod_as_pb_istream(): frames decoded on the GPU, read back through the byte-source callback."""
    import ctypes as C

    L = anm.lib()
    cfg = anm.config_preset("ref4")
    L.anm_pb_encode_to_receiver_audio.restype = C.c_size_t
    buf = (C.c_uint8 * 512)()
    wires = []
    prog = [np.full(4, 255, np.uint8)]
    for i in range(3):
        opus = bytes((i * 37 + k) & 0xFF for k in range(20 + 7 * i))
        n = L.anm_pb_encode_to_receiver_audio(opus, C.c_size_t(len(opus)), buf, C.c_size_t(512))
        wires.append(bytes(buf[:n]))
        prog += [anm.frame_symbols(cfg, wires[-1]), np.full(6, 255, np.uint8)]
    prog = np.concatenate(prog)
    n = (len(prog) + 8) * cfg.sym_len
    pcm = anm.tx_render(cfg, prog, anm.tx_params(seed=2, amplitude=0.5, snr_db=12.0), 0, n)
    assert L.demod_initialize(C.byref(cfg)) == 0
    d = L.demod_create()
    assert L.demod_feed(d, pcm.ctypes.data, C.c_size_t(n)) == 0

    class PbStream(C.Structure):
        _fields_ = [("callback", C.CFUNCTYPE(C.c_bool, C.c_void_p, C.c_void_p, C.c_size_t)), ("state", C.c_void_p),
                    ("bytes_left", C.c_size_t), ("errmsg", C.c_char_p)]

    L.demod_as_pb_istream.restype = PbStream
    L.demod_as_pb_istream.argtypes = [C.c_void_p]
    s = L.demod_as_pb_istream(d)
    total = b"".join(wires)
    got = (C.c_uint8 * len(total))()
    assert s.callback(C.addressof(s), got, len(total))
    assert bytes(got) == total
    assert not s.callback(C.addressof(s), got, 1)   # drained: fails like a closed socket
    L.demod_destroy(d)


def _gpu_render(cfg, progs, lens, params, n, first=0):
    torch = _torch()
    n_ch = progs.shape[0]
    d_prog = torch.from_numpy(progs).cuda()
    d_len = torch.from_numpy(lens.astype(np.int32)).cuda()
    d_par = torch.from_numpy(params.view(np.uint8).copy()).cuda()
    d_pcm = torch.empty((n_ch, n), dtype=torch.int16, device="cuda")
    anm.tx_render_device(cfg, d_prog.data_ptr(), progs.shape[1], d_len.data_ptr(), d_par.data_ptr(), n_ch, first, d_pcm.data_ptr(), n, n)
    torch.cuda.synchronize()
    return d_pcm


def _programs(cfg, n_ch, seed, snr_db, ppm_max, offset_max, payload=(8, 64), max_len=1024):
    from sigutil import make_program

    progs = np.full((n_ch, max_len), anm.ANM_SILENCE, dtype=np.uint8)
    lens = np.zeros(n_ch, dtype=np.int32)
    plist = []
    rng = np.random.default_rng(seed)
    for c in range(n_ch):
        prog, _ = make_program(cfg, rng, max_len // 2, payload_len=payload, gap=(2, 30))
        prog = prog[:max_len]
        progs[c, : len(prog)] = prog
        lens[c] = len(prog)
        plist.append(anm.tx_params(seed=seed * 7919 + c, start_offset=-int(rng.integers(0, offset_max + 1)), amplitude=0.5,
                                   snr_db=snr_db if not callable(snr_db) else snr_db(c),
                                   ppm=float(rng.uniform(-ppm_max, ppm_max)) if ppm_max else 0.0))
    return progs, lens, anm.tx_params_array(plist)


def test_config2_1024_channels_10s_clean_bit_exact():
    """BASELINE.json configs[1]: 1,024 channels x 10 s (3,445 symbol periods = 440,960 samples) of clean FSK
    at the reference tone set on one B200; every frame and CRC verdict equals the oracle's (compared through
    the order-independent digest of oracle/anm_oracle_batch.c, plus exact counts)."""
    import os

    from oracle_binding import frames_digest, run_batch

    torch = _torch()
    cfg = anm.config_preset("ref4")
    n_ch, n = 1024, 3445 * cfg.sym_len
    progs, lens, params = _programs(cfg, n_ch, seed=101, snr_db=None, ppm_max=0.0, offset_max=3000)
    d_pcm = _gpu_render(cfg, progs, lens, params, n)
    dm = anm.Demod(cfg, n_ch, device=0)
    chunk = 345 * cfg.sym_len
    pos = 0
    while pos < n:
        ln = min(chunk, n - pos)
        dm.feed_device(d_pcm.data_ptr() + pos * 2, n, ln, torch.cuda.current_stream().cuda_stream)
        pos += ln
    dm.collect()
    frames = anm.frames_to_list(*dm.read_frames(cap=1 << 20, bytes_cap=1 << 26))
    assert not dm.overflowed()
    dm.close()
    pcm = d_pcm.cpu().numpy()
    sec, ok, bad, nbytes, dg = run_batch(cfg, pcm, os.cpu_count() or 1)
    assert len(frames) == ok + bad and ok > 15 * n_ch and bad == 0
    assert sum(f[2] for f in frames) == ok
    assert sum(len(f[3]) for f in frames if f[2]) == nbytes
    assert frames_digest(frames) == dg


def test_config5_low_snr_drift_sweep_bit_exact():
    """BASELINE.json configs[4]: 0-3 dB SNR, +/-200 ppm clock error, random frame offsets.  Bit-exact against the
    oracle channel by channel; detection and CRC-pass rates are printed for the record."""
    torch = _torch()
    cfg = anm.config_preset("ref4")
    n_ch, n = 128, 1500 * cfg.sym_len
    progs, lens, params = _programs(cfg, n_ch, seed=55, snr_db=lambda c: float(c % 4), ppm_max=200.0, offset_max=5000, payload=(16, 200))
    d_pcm = _gpu_render(cfg, progs, lens, params, n)
    pcm = d_pcm.cpu().numpy()
    frames = _check_against_oracle(cfg, pcm, [211, 64])
    by_snr = {}
    for ch, _s, ok, _p in frames:
        k = ch % 4
        a, b = by_snr.get(k, (0, 0))
        by_snr[k] = (a + 1, b + ok)
    print("config5 frames detected / CRC ok by SNR dB:", by_snr)
    assert all(b >= 0.9 * a and a > 0 for a, b in by_snr.values())


def test_pipelined_host_feed_equals_oracle_and_queue_wraps():
    """feed_host_async + collect_upto(1): same frames as the oracle, in many small collects (the frame and
    payload queues are rings addressed by free-running counters)."""
    cfg = anm.config_preset("ref4")
    pcm, _ = make_channels(cfg, 40, 1200 * cfg.sym_len, seed=61, snr_db=9.0, offset_max=900, payload_len=(4, 24), gap=(1, 6))
    dm = anm.Demod(cfg, 40, device=0)
    step = 60 * cfg.sym_len
    got = []
    chunks = [np.ascontiguousarray(pcm[:, p: p + step]) for p in range(0, pcm.shape[1], step)]
    for ck in chunks:
        dm.feed_host_async_ptr(ck.ctypes.data, ck.shape[1], ck.shape[1])
        dm.collect_upto(1)
        got += anm.frames_to_list(*dm.read_frames())
    dm.collect()
    got += anm.frames_to_list(*dm.read_frames())
    assert not dm.overflowed()
    dm.close()
    got.sort(key=lambda f: (f[0], f[1]))
    want = oracle_frames_batch(cfg, pcm)
    assert got == want and len(got) > 400


def test_config3_one_gpu_share_streamed_10db():
    """BASELINE.json configs[2] at one GPU's share of the 8-GPU run: 8,192 channels with AWGN at 10 dB SNR, streamed in
    1-second chunks.  At this size the oracle checks a sample of channels frame for frame; the whole batch is held to
    size-independent properties: streaming in chunks == one shot (every frame, byte and CRC verdict), every detected
    frame passes its CRC at 10 dB, and every channel delivers at least the frames its program holds completely."""
    torch = _torch()
    cfg = anm.config_preset("ref4")
    n_ch, chunk, n_chunks = 8192, 344 * cfg.sym_len, 3
    n = chunk * n_chunks
    progs, lens, params = _programs(cfg, n_ch, seed=303, snr_db=10.0, ppm_max=0.0, offset_max=2000, payload=(24, 40), max_len=2200)
    d_pcm = _gpu_render(cfg, progs, lens, params, n)
    stream = torch.cuda.current_stream().cuda_stream

    def run(pieces):
        dm = anm.Demod(cfg, n_ch, device=0)
        pos = 0
        for ln in pieces:
            dm.feed_device(d_pcm.data_ptr() + pos * 2, n, ln, stream)
            pos += ln
        dm.collect()
        recs, by = dm.read_frames(cap=1 << 21, bytes_cap=1 << 27)
        assert not dm.overflowed()
        stats = dm.stats()
        dm.close()
        return recs, by, stats

    recs, by, stats = run([chunk] * n_chunks)
    recs1, by1, stats1 = run([n])
    fr = sorted(anm.frames_to_list(recs, by), key=lambda f: (f[0], f[1]))
    fr1 = sorted(anm.frames_to_list(recs1, by1), key=lambda f: (f[0], f[1]))
    assert fr == fr1 and len(fr) > 5 * n_ch
    assert np.array_equal(stats["frames_ok"], stats1["frames_ok"]) and np.array_equal(stats["symbols"], stats1["symbols"])
    assert all(f[2] == 1 for f in fr)                      # 10 dB: no CRC failure among the detected frames
    per_ch = np.bincount([f[0] for f in fr], minlength=n_ch)
    assert per_ch.min() >= 4                                # 1,032 symbol periods hold at least 4 whole frames of <= 196 + 30 symbols
    sample = list(range(0, n_ch, 128))                      # 64 channels against the oracle frame for frame
    pcm = d_pcm.cpu().numpy()
    want = oracle_frames_batch(cfg, pcm[sample])
    got = [(sample.index(f[0]), f[1], f[2], f[3]) for f in fr if f[0] in set(sample)]
    assert got == want
    # ... and ALL 8,192 channels through the oracle's threaded batch runner: frame count, CRC verdicts, payload bytes and the
    # order-independent digest over (channel, start_sample, len, crc_ok, payload) of every frame
    import os

    from oracle_binding import frames_digest, run_batch

    _sec, ok, bad, nbytes, dg = run_batch(cfg, pcm, os.cpu_count() or 1)
    assert len(fr) == ok + bad and bad == 0
    assert sum(len(f[3]) for f in fr if f[2]) == nbytes
    assert frames_digest(fr) == dg


def test_pb_istream_seam_reference_decoder_reads_gpu_frames():
    """BASELINE.json configs[0] closed end to end: ToReceiver{AudioData} messages -> frames -> PCM -> the CUDA demodulator
    (demod_feed) -> demod_as_pb_istream() -> the REFERENCE's own pb_decode_delimited(&is, ToReceiver_fields, &msg) with its
    audio-data callback (hardware/src/network.cpp:406-430, 212-249; nanopb + ip.pb.c compiled in place as oracle/_ref) ->
    the Opus bytes the transmitter put in."""
    import ctypes as C
    import os

    from oracle_binding import REF_LIB

    if not os.path.exists(REF_LIB):
        pytest.skip("oracle/_ref/libref_nanopb.so not present")
    R = C.CDLL(REF_LIB)
    R.ref_decode_to_receiver_from_stream.restype = C.c_long
    R.ref_decode_to_receiver_from_stream.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]
    R.ref_encode_to_receiver_audio.restype = C.c_size_t
    R.ref_encode_to_receiver_audio.argtypes = [C.c_char_p, C.c_size_t, C.c_void_p, C.c_size_t]
    L = anm.lib()
    cfg = anm.config_preset("ref4")
    rng = np.random.default_rng(77)
    buf = (C.c_uint8 * 2048)()
    sent, prog = [], [np.full(5, 255, np.uint8)]
    for i in range(12):
        opus = rng.integers(0, 256, size=int(rng.integers(1, 900)), dtype=np.uint8).tobytes()
        n = R.ref_encode_to_receiver_audio(opus, len(opus), buf, 2048)      # the reference's own encoder makes the wire bytes
        assert n > 0
        sent.append(opus)
        prog += [anm.frame_symbols(cfg, bytes(buf[:n])), np.full(int(rng.integers(2, 9)), 255, np.uint8)]
    prog = np.concatenate(prog)
    n = (len(prog) + 8) * cfg.sym_len
    pcm = anm.tx_render(cfg, prog, anm.tx_params(seed=4, amplitude=0.5, snr_db=11.0), 0, n)
    assert L.demod_initialize(C.byref(cfg)) == 0
    d = L.demod_create()
    for pos in range(0, n, 50000):                                             # the firmware idiom: arbitrary buffer sizes
        part = np.ascontiguousarray(pcm[pos: pos + 50000])
        assert L.demod_feed(d, part.ctypes.data, C.c_size_t(len(part))) == 0

    class PbStream(C.Structure):
        _fields_ = [("callback", C.c_void_p), ("state", C.c_void_p), ("bytes_left", C.c_size_t), ("errmsg", C.c_char_p)]

    L.demod_as_pb_istream.restype = PbStream
    L.demod_as_pb_istream.argtypes = [C.c_void_p]
    s = L.demod_as_pb_istream(d)
    out = (C.c_uint8 * 4096)()
    for want in sent:
        got = R.ref_decode_to_receiver_from_stream(s.callback, s.state, out, 4096)
        assert got == len(want) and bytes(out[:got]) == want
    assert R.ref_decode_to_receiver_from_stream(s.callback, s.state, out, 4096) == -1      # drained: like a closed socket
    L.demod_destroy(d)


def test_collect_upto_after_collect_hands_nothing_out_twice():
    """feed_async twice, collect(), collect_upto(1), collect() -- the stale snapshot of the older launch
    must not rewind the read position: no duplicates, no false overflow."""
    cfg = anm.config_preset("ref4")
    pcm, _ = make_channels(cfg, 12, 480 * cfg.sym_len, seed=67, snr_db=10.0, offset_max=600, payload_len=(4, 24), gap=(1, 6))
    half = 240 * cfg.sym_len
    a, b = np.ascontiguousarray(pcm[:, :half]), np.ascontiguousarray(pcm[:, half:])
    dm = anm.Demod(cfg, 12, device=0)
    dm.feed_host_async_ptr(a.ctypes.data, half, half)
    dm.feed_host_async_ptr(b.ctypes.data, half, half)
    dm.collect()
    got = anm.frames_to_list(*dm.read_frames())
    assert dm.collect_upto(1) == 0 and dm.collect_upto(0) == 0      # older / same launches: nothing new
    got += anm.frames_to_list(*dm.read_frames())
    dm.collect()
    got += anm.frames_to_list(*dm.read_frames())
    assert not dm.overflowed()
    dm.close()
    want = oracle_frames_batch(cfg, pcm)
    assert sorted(got, key=lambda f: (f[0], f[1])) == want and len(got) == len(set((f[0], f[1]) for f in got))


def test_take_frames_and_one_at_a_time_reads():
    """take_frames moves a whole drain out in arrival order; read_frames(cap=1) pops in (channel, start_sample) order from a
    cursor (no re-sort per call); mixing feeds between reads keeps the order of what is still unread."""
    torch = _torch()
    cfg = anm.config_preset("ref4")
    pcm, _ = make_channels(cfg, 30, 640 * cfg.sym_len, seed=71, snr_db=10.0, offset_max=700, payload_len=(4, 20), gap=(1, 5))
    want = oracle_frames_batch(cfg, pcm)
    d_pcm = torch.from_numpy(pcm).cuda()
    n = pcm.shape[1]
    stream = torch.cuda.current_stream().cuda_stream
    # (1) take_frames
    dm = anm.Demod(cfg, 30, device=0)
    dm.feed_device(d_pcm.data_ptr(), n, n, stream)
    dm.collect()
    recs = np.zeros(1 << 14, dtype=anm.FRAME_DTYPE)
    by = np.zeros(1 << 20, dtype=np.uint8)
    small = np.zeros(3, dtype=anm.FRAME_DTYPE)
    assert dm.take_frames(small, by) == (0, 0)                       # too small: queue untouched
    nf, nb = dm.take_frames(recs, by)
    got = anm.frames_to_list(recs[:nf], by[:nb])
    assert sorted(got, key=lambda f: (f[0], f[1])) == want and nb == sum(len(f[3]) for f in want)
    for c in range(30):                                              # per channel still chronological
        starts = [f[1] for f in got if f[0] == c]
        assert starts == sorted(starts)
    assert anm.frames_digest(recs[:nf], by[:nb]) == __import__("oracle_binding").frames_digest(want)
    assert dm.take_frames(recs, by) == (0, 0)
    dm.close()
    # (2) cap=1 reads interleaved with a second feed
    dm = anm.Demod(cfg, 30, device=0)
    half = 320 * cfg.sym_len
    dm.feed_device(d_pcm.data_ptr(), n, half, stream)
    dm.collect()
    first = [anm.frames_to_list(*dm.read_frames(cap=1))[0] for _ in range(5)]
    dm.feed_device(d_pcm.data_ptr() + half * 2, n, n - half, stream)
    dm.collect()
    rest = []
    while True:
        r = anm.frames_to_list(*dm.read_frames(cap=1))
        if not r:
            break
        rest += r
    assert sorted(first + rest, key=lambda f: (f[0], f[1])) == want
    assert rest == sorted(rest, key=lambda f: (f[0], f[1]))          # what was unread comes out fully ordered
    dm.close()


def test_feed_on_two_streams_is_ordered():
    """A launch on another stream than the handle's previous launch is ordered behind it."""
    torch = _torch()
    cfg = anm.config_preset("ref4")
    pcm, _ = make_channels(cfg, 200, 512 * cfg.sym_len, seed=73, snr_db=10.0, offset_max=500)
    d_pcm = torch.from_numpy(pcm).cuda()
    torch.cuda.synchronize()
    n = pcm.shape[1]
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    dm = anm.Demod(cfg, 200, device=0)
    q = 128 * cfg.sym_len
    for i in range(4):
        dm.feed_device(d_pcm.data_ptr() + i * q * 2, n, q, (s1 if i % 2 == 0 else s2).cuda_stream)
    dm.collect()
    got = anm.frames_to_list(*dm.read_frames())
    dm.close()
    assert got == oracle_frames_batch(cfg, pcm)


@pytest.mark.parametrize("preset,n_ch,n_chunks,syms", [("ref4", 3500, 6, 96), ("ref4", 700, 9, 45), ("bfsk2", 300, 4, 64), ("mfsk16", 200, 5, 64)])
def test_several_chunks_in_one_launch_equal_chunk_by_chunk(preset, n_ch, n_chunks, syms):
    """anm_demod_feed_device_chunks: (chunk, channel) work items from one queue, a channel's chunks ordered by its progress counter -- the frames equal
    those of n_chunks single launches (more channels than resident warps, ragged 45-symbol chunks, noisy drifting signals), twice in a row on one handle
    and mixed with single launches; 40 of the channels against the oracle."""
    torch = _torch()
    cfg = anm.config_preset(preset)
    q = syms * cfg.sym_len
    pcm, _ = make_channels(cfg, n_ch, 2 * n_chunks * q + q, seed=91, snr_db=9.0, offset_max=700, payload_len=(4, 40), gap=(2, 10), ppm_max=150.0)
    d_pcm = torch.from_numpy(pcm).cuda()
    torch.cuda.synchronize()
    n = pcm.shape[1]
    st = torch.cuda.current_stream().cuda_stream
    a = anm.Demod(cfg, n_ch, device=0)
    for c in range(2 * n_chunks + 1):
        a.feed_device(d_pcm.data_ptr() + c * q * 2, n, q, st)
    a.collect()
    want = anm.frames_to_list(*a.read_frames())
    a.close()
    b = anm.Demod(cfg, n_ch, device=0)
    b.feed_device_chunks(d_pcm.data_ptr(), n, q, q, n_chunks, st)                       # chunks 0 .. n_chunks - 1 in one launch
    b.feed_device(d_pcm.data_ptr() + n_chunks * q * 2, n, q, st)                        # one single launch in between
    b.feed_device_chunks(d_pcm.data_ptr() + (n_chunks + 1) * q * 2, n, q, q, n_chunks, st)
    b.collect()
    got = anm.frames_to_list(*b.read_frames())
    assert not b.overflowed()
    b.close()
    assert got == want and len(got) > n_ch
    assert [f for f in got if f[0] < 40] == oracle_frames_batch(cfg, pcm[:40])


def test_chunks_call_on_the_dense_kernel_and_with_symbol_recording_takes_one_launch_per_chunk():
    """anm_demod_feed_device_chunks on a configuration of the tensor-core kernel, and on a handle that records decided symbols: same frames (and
    symbols) as chunk-by-chunk feeding; the call falls back to one launch per chunk there"""
    torch = _torch()
    st = torch.cuda.current_stream().cuda_stream
    for preset, flags in (("wide64", 0), ("ref4", anm.ANM_FLAG_SYMBOLS)):
        cfg = anm.config_preset(preset)
        q, n_chunks, n_ch = 64 * cfg.sym_len, 4, 24
        pcm, _ = make_channels(cfg, n_ch, n_chunks * q, seed=95, snr_db=12.0, offset_max=300, payload_len=(4, 30), gap=(2, 8))
        d_pcm = torch.from_numpy(pcm).cuda()
        torch.cuda.synchronize()
        a = anm.Demod(cfg, n_ch, device=0, flags=flags)
        l0 = a.launch_count()
        a.feed_device_chunks(d_pcm.data_ptr(), pcm.shape[1], q, q, n_chunks, st)
        assert a.launch_count() - l0 == n_chunks
        a.collect()
        got = anm.frames_to_list(*a.read_frames())
        a.close()
        assert got == oracle_frames_batch(cfg, pcm) and len(got) > n_ch


def _multi_devices():
    torch = _torch()
    n = torch.cuda.device_count()
    return [0, 1 % n, 0] if n > 1 else [0, 0, 0]     # a device may appear more than once: three shards, three host threads


def test_multi_device_handle_gathers_frames_in_order():
    """anm_demod_multi_*: channels sharded over several device handles (one host thread each), PCM fed from ONE pinned host
    buffer, frames gathered host-side with global channel ids in (channel, start_sample) order == the oracle; uneven shards
    (50 channels over 3), streamed in chunks with pipelined collects."""
    cfg = anm.config_preset("ref4")
    n_ch, n = 50, 900 * cfg.sym_len
    pcm, _ = make_channels(cfg, n_ch, n, seed=83, snr_db=9.0, offset_max=800, payload_len=(4, 30), gap=(1, 8))
    m = anm.DemodMulti(cfg, n_ch, _multi_devices())
    sh = m.shards()
    assert [s["n_channels"] for s in sh] == [17, 17, 16] and [s["first_channel"] for s in sh] == [0, 17, 34]
    step = 300 * cfg.sym_len
    bufs = [m.alloc_pcm(step) for _ in range(2)]
    got = []
    for i, pos in enumerate(range(0, n, step)):
        b = bufs[i % 2]
        if i >= 2:
            m.wait_input()                         # the copy that read this buffer two feeds ago has completed
        b[:, :] = pcm[:, pos: pos + step]
        m.feed_host(b)
        m.collect_upto(1)
        got += anm.frames_to_list(*m.read_frames())
    m.collect()
    got += anm.frames_to_list(*m.read_frames())
    assert not m.overflowed()
    st = m.stats()
    m.close()
    want = oracle_frames_batch(cfg, pcm)
    assert sorted(got, key=lambda f: (f[0], f[1])) == want and len(want) > 300
    assert int(st["frames_ok"].sum()) == sum(f[2] for f in want)
    # one collect hands out frames globally ordered
    m = anm.DemodMulti(cfg, n_ch, _multi_devices())
    b = m.alloc_pcm(n)
    b[:, :] = pcm
    m.feed_host(b)
    m.collect()
    assert anm.frames_to_list(*m.read_frames()) == want
    m.close()


def test_gpu_transmitter_equals_cpu_transmitters_far_into_a_stream():
    """beyond 2^31 samples (13.5 h at 44.1 kHz) and with large |start_offset| the GPU renderer must still equal the CPU renderers
    (128-bit positions); checked against the product's CPU transmitter and the oracle's own one."""
    import oracle_binding as ob

    cfg = anm.config_preset("ref4")
    rng = np.random.default_rng(97)
    n_ch, n = 6, 4096
    progs = np.full((n_ch, 64), anm.ANM_SILENCE, dtype=np.uint8)
    lens = np.zeros(n_ch, dtype=np.int32)
    plist = []
    for c in range(n_ch):
        ln = int(rng.integers(3, 64))
        progs[c, :ln] = rng.integers(0, 4, size=ln)
        lens[c] = ln
        plist.append(dict(seed=1000 + c, start_offset=int(rng.choice([-(1 << 33), -12345, 0, (1 << 34) + 77])), amplitude=0.5, snr_db=9.0,
                          ppm=float(rng.uniform(-250, 250))))
    params = anm.tx_params_array([anm.tx_params(**k) for k in plist])
    for first in (0, (1 << 31) - 100, (1 << 33) + 12345, (1 << 40) + 5):
        d = _gpu_render(cfg, progs, lens, params, n, first=first).cpu().numpy()
        for c in range(n_ch):
            a = anm.tx_render(cfg, progs[c, : lens[c]], anm.tx_params(**plist[c]), first, n)
            b = ob.tx_render(ob.preset("ref4"), progs[c, : lens[c]], ob.tx_params(**plist[c]), first, n)
            assert np.array_equal(d[c], a) and np.array_equal(a, b), (first, c)
