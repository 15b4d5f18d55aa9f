"""Row f1, stage 3: the batched CELT decode down to PCM (include/anmodem_opus.h, anm_celt_decode_*) against the REFERENCE's libopus 1.3.1 in its fixed-point
build: every int16 sample equals what the reference decoder writes.

Two independent pins: (1) tests/golden/opus_packets.json -- SHA-256 of the PCM the reference's PUBLIC opus_decode() returns for whole streams of the
transmitter's settings; (2) tests/golden/celt_pcm.npz -- per-frame digests of what the reference's celt_decode_with_ec() writes for the 1,687 frames
of the entropy goldens, with a native, a mono (downmix) and a stereo (upmix) decoder (tests/golden/make_celt_pcm_golden.py).  `-m "not gpu"` runs the
product's headers compiled for the host by the test harness (a checker of the logic: the product has no CPU path); `-m gpu` runs the CUDA kernels
through the C ABI."""
import base64
import ctypes as C
import hashlib
import json
import os

import numpy as np
import pytest

import audio_network_b200 as anm
import celt_binding as cb
import celt_spectrum_binding as sbind

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = np.load(os.path.join(HERE, "golden", "celt_entropy.npz"))
PCM = np.load(os.path.join(HERE, "golden", "celt_pcm.npz"))
PACKETS = json.load(open(os.path.join(HERE, "golden", "opus_packets.json")))
HAVE_REF = os.path.exists(cb.REF_OPUS)


def _digest(a):
    return np.frombuffer(hashlib.sha256(np.ascontiguousarray(a, dtype="<i2").tobytes()).digest()[:8], dtype="<u8")[0]


def _harness():
    L = sbind.harness()
    L.anm_celt_synth_tables_build.argtypes = [C.c_void_p]
    L.harness_celt_decode_frame.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                            C.c_void_p]
    assert L.harness_sizeof_synth() == anm.CELT_SYNTH_DTYPE.itemsize == 17376 and L.harness_sizeof_synth_tables() == anm.CELT_SYNTH_TABLES_DTYPE.itemsize
    stb = np.zeros(1, anm.CELT_SYNTH_TABLES_DTYPE)
    assert L.anm_celt_synth_tables_build(stb.ctypes.data) == 0
    return L, stb


def _packet_bytes(p):
    return bytes.fromhex(p) if all(c in "0123456789abcdefABCDEF" for c in p) else base64.b64decode(p)


def _stream_frames(s):
    """the CELT frames of one golden opus stream, in order: [(bytes, channels, lm, end)]"""
    out = []
    for p, parse in zip(s["packets"], s["parse"]):
        out += cb.frames_of_packet(_packet_bytes(p), parse)
    return out


@pytest.mark.skipif(not HAVE_REF, reason="oracle/_ref/libref_opus.so not built (reference tree absent)")
def test_computed_synthesis_tables_equal_the_reference_static_tables():
    """window, MDCT twiddles, FFT twiddles and bit-reversal tables are COMPUTED by anm_celt_tables.c; they must equal the reference's static tables"""
    _, stb = _harness()
    R = sbind.ref()
    R.ref_celt_synth_tables.argtypes = [C.c_void_p] * 4
    w, tr, tw, br = np.zeros(120, np.int16), np.zeros(1800, np.int16), np.zeros(960, np.int16), np.zeros(900, np.int16)
    assert R.ref_celt_synth_tables(w.ctypes.data, tr.ctypes.data, tw.ctypes.data, br.ctypes.data) == 1920
    for k, a in (("window", w), ("trig", tr), ("fft_tw", tw), ("bitrev", br)):
        assert np.array_equal(stb[0][k], a), k


def _host_decode_stream(L, stb, frames, cc):
    t = cb.tables()
    st, syn = np.zeros(1, anm.CELT_STREAM_DTYPE), np.zeros(1, anm.CELT_SYNTH_DTYPE)
    out = []
    for f, ch, lm, end in frames:
        b = np.frombuffer(f, np.uint8).copy() if len(f) else np.zeros(1, np.uint8)
        rec, pcm = np.zeros(1, cb.FRAME_DTYPE), np.zeros(960 * cc, np.int16)
        assert L.harness_celt_decode_frame(t.ctypes.data, stb.ctypes.data, b.ctypes.data, len(f), ch, cc, lm, end, st.ctypes.data, syn.ctypes.data, rec.ctypes.data,
                                           pcm.ctypes.data) == 0
        out.append(pcm[: (120 << lm) * cc].copy())
    return out


def test_host_harness_pcm_equals_the_public_api_decode_of_the_golden_streams():
    L, stb = _harness()
    for s in PACKETS["streams"]:
        pcm = np.concatenate(_host_decode_stream(L, stb, _stream_frames(s), s["channels"]))
        assert hashlib.sha256(pcm.tobytes()).hexdigest() == s["decoded_sha256"], s["name"]
        assert pcm[:16].tolist() == np.array(s["decoded_head"]).reshape(-1)[:16].tolist()


def test_host_harness_pcm_equals_the_reference_decoder_frame_by_frame_native_mono_and_stereo_decoders():
    L, stb = _harness()
    fr, by, sb = GOLD["frames"], GOLD["bytes"], GOLD["stream_begin"]
    for s in range(len(sb) - 1):
        js = range(sb[s], sb[s + 1])
        frames = [(bytes(by[fr[j]["offset"]: fr[j]["offset"] + fr[j]["len"]]), int(fr[j]["channels"]), int(fr[j]["lm"]), int(fr[j]["end_band"])) for j in js]
        native = max(f[1] for f in frames)
        for key, cc in (("cc_native", native), ("cc_mono", 1), ("cc_stereo", 2)):
            if s % 4 and key != "cc_native":
                continue                      # every stream natively, every fourth one through the down- and upmixing decoders too
            for j, pcm in zip(js, _host_decode_stream(L, stb, frames, cc)):
                assert _digest(pcm) == PCM[key][j], (s, key, j)


def _random_streams(rng, n_streams):
    """random bytes as frames in streams whose channel count, frame size and bandwidth may change from frame to frame, decoded by a mono or a stereo
    decoder: [(frames [(bytes, C, lm, end)], cc, reference PCM per frame, reference return codes)]"""
    R = sbind.ref()
    out = []
    for _ in range(n_streams):
        cc, nfr = int(rng.integers(1, 3)), int(rng.integers(2, 7))
        params = np.array([[int(rng.integers(1, 3)), int(rng.integers(0, 4)), int(rng.choice([13, 17, 19, 21]))] for _ in range(nfr)], np.int32)
        if rng.random() < 0.6:
            params[:] = params[0]
        frames = [rng.integers(0, 256, int(rng.choice([rng.integers(2, 12), rng.integers(12, 120), rng.integers(120, 500), rng.integers(500, 1276)])), dtype=np.uint8)
                  for _ in range(nfr)]
        for f in frames:
            if rng.random() < 0.7:
                f[0] &= 0x7F
        maxlen = max(len(f) for f in frames)
        buf, lens = np.zeros((nfr, maxlen), np.uint8), np.array([len(f) for f in frames], np.int32)
        for k, f in enumerate(frames):
            buf[k, :len(f)] = f
        states, pcm = np.zeros(nfr, sbind.STATE), np.zeros((nfr, 960 * cc), np.int16)
        assert R.ref_celt_stream_states(buf.ctypes.data, lens.ctypes.data, nfr, maxlen, params.ctypes.data, cc, states.ctypes.data, pcm.ctypes.data) == nfr
        out.append(([(frames[k].tobytes(), int(params[k, 0]), int(params[k, 1]), int(params[k, 2])) for k in range(nfr)], cc, pcm, states["ret"].copy()))
    return out


@pytest.mark.skipif(not HAVE_REF, reason="oracle/_ref/libref_opus.so not built (reference tree absent)")
def test_host_harness_random_streams_pcm_vs_reference_live():
    """garbage in, the same garbage out: random frames (extreme gains, saturation, every post-filter setting, layout changes inside a stream)"""
    L, stb = _harness()
    rng = np.random.default_rng(int.from_bytes(os.urandom(4), "little"))
    n = 0
    for frames, cc, want, ret in _random_streams(rng, 120):
        for k, pcm in enumerate(_host_decode_stream(L, stb, frames, cc)):
            if ret[k] >= 0:
                assert np.array_equal(pcm, want[k, : len(pcm)]), (k, frames[k][1:], cc, len(frames[k][0]))
                n += 1
    assert n > 300


def test_jobs_from_packets_equals_the_reference_packet_framing():
    """anm_celt_jobs_from_packets (host glue): frame offsets, sizes, lm, end band and channels of every frame of the golden packets (1-frame and
    3-frame packets, code 0 and code 3) from the REFERENCE's parse records; non-CELT and rejected packets give no job"""
    for s in PACKETS["streams"]:
        spans = np.zeros(len(s["packets"]), dtype=anm.PB_SPAN_DTYPE)
        pks = np.zeros(len(s["packets"]), dtype=anm.OPUS_PACKET_DTYPE)
        arena, want = bytearray(), []
        for i, (p, parse) in enumerate(zip(s["packets"], s["parse"])):
            b = _packet_bytes(p)
            spans[i]["audio_offset"], spans[i]["audio_len"] = len(arena), len(b)
            for k in ("count", "toc", "channels", "mode", "bandwidth", "samples_per_frame", "payload_offset", "nb_frames", "nb_samples"):
                pks[i][k] = parse[k]
            pks[i]["size"] = parse["size"]
            off = len(arena) + parse["payload_offset"]
            for f, ch, lm, end in cb.frames_of_packet(b, parse):
                want.append((off, len(f), ch, lm, end, 0))
                off += len(f)
            arena += b
        jobs, first = anm.celt_jobs_from_packets(spans, pks)
        assert [tuple(int(v) for v in j) for j in jobs] == want, s["name"]
        assert first[0] == 0 and (np.diff(first.astype(np.int64)) == [p["count"] for p in s["parse"]][:-1]).all()
    bad = np.zeros(3, dtype=anm.OPUS_PACKET_DTYPE)
    bad["count"], bad["mode"], bad["bandwidth"], bad["samples_per_frame"] = [1, -4, 1], [1000, 1002, 1002], [1103, 1105, 1105], [960, 960, 1920]
    jobs, first = anm.celt_jobs_from_packets(np.zeros(3, dtype=anm.PB_SPAN_DTYPE), bad)
    assert len(jobs) == 0 and (first == 0xFFFFFFFF).all()


# ------------------------------------------------------------------------------------------------ GPU
def _jobs_from_gold(cc_of_stream):
    fr, sb = GOLD["frames"], GOLD["stream_begin"]
    jobs = np.zeros(len(fr), dtype=anm.CELT_JOB_DTYPE)
    for k in ("offset", "len", "channels", "lm", "end_band"):
        jobs[k] = fr[k]
    for s in range(len(sb) - 1):
        if cc_of_stream[s] == 1:
            jobs["flags"][sb[s]: sb[s + 1]] = anm.CELT_JOB_DISABLE_INV
    return jobs


@pytest.mark.gpu
@pytest.mark.parametrize("key", ["cc_native", "cc_mono", "cc_stereo"])
def test_gpu_pcm_equals_the_reference_decoder_on_the_golden_frames(key):
    """k_celt_entropy -> k_celt_energies -> k_celt_spectrum -> k_celt_blocks -> k_celt_overlap -> k_celt_deemphasis through the C ABI: 58 streams / 1,687 frames"""
    fr, sb = GOLD["frames"], GOLD["stream_begin"]
    native = np.array([fr["channels"][sb[s]: sb[s + 1]].max() for s in range(len(sb) - 1)])
    cc = {"cc_native": native, "cc_mono": np.ones_like(native), "cc_stereo": np.full_like(native, 2)}[key]
    rec, _, sy, pcm = anm.celt_decode(_jobs_from_gold(cc), sb, GOLD["bytes"], out_channels=cc)
    assert np.array_equal(rec["final_range"], fr["final_range"]) and np.array_equal(sy["out_channels"], cc)
    for s in range(len(sb) - 1):
        for j in range(sb[s], sb[s + 1]):
            ns = (120 << int(fr["lm"][j])) * int(cc[s])
            assert _digest(pcm[j, :ns]) == PCM[key][j], (key, s, j)
    if key == "cc_native":
        for j, want in zip(PCM["head_idx"], PCM["head"]):
            assert np.array_equal(pcm[j], want)


@pytest.mark.gpu
def test_gpu_pcm_of_the_transmitter_streams_equals_the_public_api_and_state_carries_across_calls():
    """the five golden streams of the transmitter's settings, each decoded in ONE call and in two calls with the stream state (energies, histories, noise
    seed) and the synthesis state (overlap, post-filter memory, de-emphasis) carried through: SHA-256 of the PCM == the reference's opus_decode()"""
    for s in PACKETS["streams"]:
        frames = _stream_frames(s)
        by = np.frombuffer(b"".join(f[0] for f in frames), np.uint8)
        jobs = np.zeros(len(frames), dtype=anm.CELT_JOB_DTYPE)
        off = 0
        for k, (f, ch, lm, end) in enumerate(frames):
            jobs[k] = (off, len(f), ch, lm, end, anm.CELT_JOB_DISABLE_INV if s["channels"] == 1 else 0)
            off += len(f)
        cc = s["channels"]

        def flat(pcm, js):
            return np.concatenate([pcm[k, : (120 << int(js["lm"][k])) * cc] for k in range(len(js))])
        _, _, _, pcm = anm.celt_decode(jobs, [0, len(jobs)], by, out_channels=[cc])
        assert hashlib.sha256(flat(pcm, jobs).tobytes()).hexdigest() == s["decoded_sha256"], s["name"]
        h = len(jobs) // 2
        _, st, sy, pa = anm.celt_decode(jobs[:h], [0, h], by, out_channels=[cc])
        _, _, _, pb = anm.celt_decode(jobs[h:], [0, len(jobs) - h], by, streams=st, synth=sy)
        assert hashlib.sha256(np.concatenate([flat(pa, jobs[:h]), flat(pb, jobs[h:])]).tobytes()).hexdigest() == s["decoded_sha256"], s["name"]


@pytest.mark.gpu
@pytest.mark.skipif(not HAVE_REF, reason="oracle/_ref/libref_opus.so not built (reference tree absent)")
def test_gpu_random_streams_pcm_vs_reference_live():
    """150 streams of random frames with changing layouts in ONE call: every sample the GPU returns equals the reference decoder's"""
    rng = np.random.default_rng(int.from_bytes(os.urandom(4), "little"))
    streams = _random_streams(rng, 150)
    jobs, by, sbeg, ccs = [], bytearray(), [0], []
    for frames, cc, _, _ in streams:
        for f, ch, lm, end in frames:
            jobs.append((len(by), len(f), ch, lm, end, anm.CELT_JOB_DISABLE_INV if cc == 1 else 0))
            by += f
        sbeg.append(len(jobs))
        ccs.append(cc)
    jobs = np.array(jobs, dtype=anm.CELT_JOB_DTYPE)
    _, _, _, pcm = anm.celt_decode(jobs, np.array(sbeg, np.uint32), np.frombuffer(bytes(by), np.uint8), out_channels=ccs)
    n = 0
    for s, (frames, cc, want, ret) in enumerate(streams):
        for k in range(len(frames)):
            if ret[k] >= 0:
                ns = (120 << frames[k][2]) * cc
                assert np.array_equal(pcm[sbeg[s] + k, :ns], want[k, :ns]), (s, k, frames[k][1:], cc)
                n += 1
    assert n > 400


@pytest.mark.gpu
def test_gpu_device_call_with_pcm_rows_of_any_stride_and_alignment():
    """anm_celt_decode_device writes 16 bytes of PCM at a time when the rows allow it and sample by sample when they do not (a stride that is no multiple of
    8 samples, a buffer that starts off a 16-byte boundary): both equal the public API's result, mono and stereo"""
    import torch
    L = anm.lib()
    dev = torch.device("cuda:0")
    for s in [x for x in PACKETS["streams"] if x["name"] in ("stereo_20ms", "mono_20ms")] or PACKETS["streams"][:2]:
        frames = _stream_frames(s)
        by = np.frombuffer(b"".join(f[0] for f in frames), np.uint8)
        jobs = np.zeros(len(frames), dtype=anm.CELT_JOB_DTYPE)
        off = 0
        for k, (f, ch, lm, end) in enumerate(frames):
            jobs[k] = (off, len(f), ch, lm, end, anm.CELT_JOB_DISABLE_INV if s["channels"] == 1 else 0)
            off += len(f)
        cc = s["channels"]
        _, _, _, want = anm.celt_decode(jobs, [0, len(jobs)], by, out_channels=[cc])
        ctx = C.c_void_p()
        assert L.anm_celt_ctx_create(0, C.byref(ctx)) == 0
        d_by = torch.from_numpy(by.copy()).to(dev)
        d_jobs = torch.from_numpy(jobs.view(np.uint8).reshape(-1).copy()).to(dev)
        d_sb = torch.tensor([0, len(jobs)], dtype=torch.int32, device=dev)
        d_fr = torch.zeros(len(jobs) * anm.CELT_FRAME_DTYPE.itemsize, dtype=torch.uint8, device=dev)
        for stride, shift in ((1920, 0), (1924, 0), (1921, 0), (1920, 3), (1928, 8)):
            d_st = torch.zeros(anm.CELT_STREAM_DTYPE.itemsize, dtype=torch.uint8, device=dev)
            sy = np.zeros(1, dtype=anm.CELT_SYNTH_DTYPE)
            sy["out_channels"] = cc
            d_sy = torch.from_numpy(sy.view(np.uint8).reshape(-1).copy()).to(dev)
            d_pcm = torch.zeros(len(jobs) * stride + 16, dtype=torch.int16, device=dev)
            rc = L.anm_celt_decode_device(ctx, d_jobs.data_ptr(), d_sb.data_ptr(), 1, len(jobs), d_by.data_ptr(), 0xFFFFFFFF, d_st.data_ptr(), d_sy.data_ptr(),
                                          d_fr.data_ptr(), d_pcm.data_ptr() + 2 * shift, stride, None)
            assert rc == 0, L.anm_last_error()
            torch.cuda.synchronize()
            got = d_pcm[shift: shift + len(jobs) * stride].view(len(jobs), stride).cpu().numpy()
            for k in range(len(jobs)):
                ns = (120 << int(jobs["lm"][k])) * cc
                assert np.array_equal(got[k, :ns], want[k, :ns]), (s["name"], stride, shift, k)
        L.anm_celt_ctx_destroy(ctx)


@pytest.mark.gpu
def test_gpu_receive_chain_pcm_in_pcm_out():
    """The widened hot path end to end on the GPU: ToReceiver{AudioData{CELT packet}} messages -> modem frames -> modem PCM -> k_demod -> k_pb_deframe ->
    k_opus_parse -> the six CELT kernels: the AUDIO that comes out equals what the reference's opus_decode() returns for the packets that went in."""
    s = [x for x in PACKETS["streams"] if x["name"] == "stereo_20ms"][0]
    L = anm.lib()
    L.anm_pb_encode_to_receiver_audio.restype = C.c_size_t
    cfg = anm.config_preset("ref4")
    buf = (C.c_uint8 * 2048)()
    prog = [np.full(4, 255, np.uint8)]
    for p in s["packets"]:
        opus = _packet_bytes(p)
        n = L.anm_pb_encode_to_receiver_audio(opus, C.c_size_t(len(opus)), buf, C.c_size_t(2048))
        prog += [anm.frame_symbols(cfg, bytes(buf[:n])), np.full(5, 255, np.uint8)]
    prog = np.concatenate(prog)
    n = ((len(prog) + 8 + 31) // 32 * 32) * cfg.sym_len
    pcm_in = anm.tx_render(cfg, prog, anm.tx_params(seed=8, amplitude=0.5, snr_db=12.0), 0, n).reshape(1, -1)
    dm = anm.Demod(cfg, 1, device=0)
    dm.feed_host(np.ascontiguousarray(pcm_in))
    dm.collect()
    recs, pay = dm.read_frames(cap=1 << 12, bytes_cap=1 << 22)
    dm.close()
    assert len(recs) == len(s["packets"]) and (recs["crc_ok"] == 1).all()
    spans = anm.pb_deframe(recs, pay)
    pk = anm.opus_parse(spans, pay)
    assert (pk["count"] == 1).all() and (pk["mode"] == 1002).all()
    jobs, first = anm.celt_jobs_from_packets(spans, pk)                      # the product's own glue from parse records to frame jobs
    assert len(jobs) == len(recs) and np.array_equal(first, np.arange(len(recs))) and (jobs["lm"] == 3).all() and (jobs["end_band"] == 21).all()
    _, _, _, audio = anm.celt_decode(jobs, [0, len(jobs)], pay, out_channels=[2])
    assert hashlib.sha256(audio[:, :1920].tobytes()).hexdigest() == s["decoded_sha256"]
