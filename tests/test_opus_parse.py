"""Batched Opus packet parse (first stage; include/anmodem_opus.h, csrc/anm_opus_gpu.cu) against the
REFERENCE's libopus 1.3.1: committed results (tests/golden/opus_packets.json, opus_synthetic.npz, made by
tests/golden/make_opus_golden.py from oracle/_ref/libref_opus.so) and, when that library is present, the library itself."""
import hashlib
import json
import os

import numpy as np
import pytest

import audio_network_b200 as anm
import opus_corpus as oc

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = json.load(open(os.path.join(HERE, "golden", "opus_packets.json")))
SYN = np.load(os.path.join(HERE, "golden", "opus_synthetic.npz"))
HAVE_REF = os.path.exists(oc.REF_OPUS)


def _spans(packets):
    """packets laid out back to back in one arena, one ANM_PB_OK span each (what pb_deframe reports)"""
    spans = np.zeros(len(packets), dtype=anm.PB_SPAN_DTYPE)
    off = 0
    for i, p in enumerate(packets):
        spans[i] = (anm.ANM_PB_OK, 0, off, len(p))
        off += len(p)
    return spans, np.frombuffer(b"".join(packets) or b"\0", dtype=np.uint8)


def _equal(got, ref):
    for k in oc.FIELDS:
        assert int(got[k]) == int(ref[k]), (k, int(got[k]), int(ref[k]))
    assert [int(v) for v in got["size"]] == [int(v) for v in ref["size"]]


def test_golden_is_from_the_reference_library():
    assert GOLD["version"] == "libopus 1.3.1-fixed"
    modes = {q["mode"] for s in GOLD["streams"] for q in s["parse"]}
    assert modes == {anm.lib() and 1002}  # the transmitter's settings give CELT-only packets
    for s in GOLD["streams"]:
        assert all(q["count"] * q["samples_per_frame"] == s["frame_samples"] for q in s["parse"])


@pytest.mark.skipif(not HAVE_REF, reason="oracle/_ref/libref_opus.so not built (reference tree absent)")
def test_reference_library_reproduces_the_committed_fixtures():
    """the fixtures are what the reference's libopus says today: encoder packets, parse results, decoded-PCM digest"""
    R = oc.ref_lib()
    for s in GOLD["streams"]:
        pcm, packets = oc.encode_stream(R, len(s["packets"]), s["frame_samples"], s["channels"], s["seed"])
        assert hashlib.sha256(pcm.tobytes()).hexdigest() == s["input_sha256"]
        assert [p.hex() for p in packets] == s["packets"]
        assert [oc.ref_parse(R, p) for p in packets] == s["parse"]
        dec = oc.decode_stream(R, packets, s["channels"], s["frame_samples"])
        assert hashlib.sha256(dec.tobytes()).hexdigest() == s["decoded_sha256"]
    off = SYN["off"]
    for i in range(0, len(SYN["fs"]), 37):
        p = SYN["bytes"][off[i]: off[i + 1]].tobytes()
        _equal(SYN["ref"][i], oc.ref_parse(R, p, int(SYN["fs"][i])))


def test_no_cpu_fallback_and_argument_errors():
    import torch

    spans, by = _spans([b"\xfc\x00"])
    L = anm.lib()
    out = np.zeros(1, dtype=anm.OPUS_PACKET_DTYPE)
    assert L.anm_opus_parse_host(spans.ctypes.data, 1, by.ctypes.data, 1, 48000, out.ctypes.data) == anm.ANM_ERR_ARG  # span beyond the arena
    assert L.anm_opus_parse_device(None, 1, None, 0xFFFFFFFF, 48000, None, None) == anm.ANM_ERR_ARG
    assert L.anm_opus_parse_device(1, 1, 1, 0xFFFFFFFF, 44100, 1, None) == anm.ANM_ERR_ARG  # not a decoder rate
    if not torch.cuda.is_available():
        with pytest.raises(anm.AnmError) as e:
            anm.opus_parse(spans, by)
        assert e.value.code == anm.ANM_ERR_CUDA and "no CPU fallback" in str(e.value)


@pytest.mark.gpu
def test_parse_of_reference_encoder_packets():
    for s in GOLD["streams"]:
        packets = [bytes.fromhex(h) for h in s["packets"]]
        got = anm.opus_parse(*_spans(packets))
        for g, ref in zip(got, s["parse"]):
            _equal(g, ref)
        assert int(got["nb_samples"].sum()) == len(packets) * s["frame_samples"]


@pytest.mark.gpu
def test_parse_matches_reference_on_the_synthetic_corpus():
    off, fs_all = SYN["off"], SYN["fs"]
    n_ok = 0
    for fs in (48000, 16000, 8000):
        idx = np.nonzero(fs_all == fs)[0]
        packets = [SYN["bytes"][off[i]: off[i + 1]].tobytes() for i in idx]
        keep = [k for k, p in enumerate(packets) if len(p) > 0]   # empty packets never reach the parser (see header)
        got = anm.opus_parse(*_spans([packets[k] for k in keep]), fs=fs)
        ref = SYN["ref"][idx[keep]]
        for f in oc.FIELDS + ("size",):
            bad = np.nonzero((got[f] != ref[f]).reshape(len(got), -1).any(axis=1))[0]
            assert len(bad) == 0, (f, packets[keep[bad[0]]][:16].hex(), got[f][bad[0]], ref[f][bad[0]])
        n_ok += int((got["count"] > 0).sum())
    assert len(fs_all) > 10000 and n_ok > 2500


@pytest.mark.gpu
def test_spans_without_audio_and_ring_addressing():
    packets = [bytes.fromhex(h) for h in GOLD["streams"][1]["packets"][:6]]
    spans, by = _spans(packets)
    spans["status"][2] = anm.ANM_PB_FAIL
    spans["status"][4] = anm.ANM_PB_CRC
    spans["audio_len"][5] = 0
    got = anm.opus_parse(spans, by)
    for i in (2, 4, 5):
        assert got["count"][i] == anm.ANM_OPUS_BAD_ARG and got["toc"][i] == 0 and not got["size"][i].any()
    for i in (0, 1, 3):
        _equal(got[i], GOLD["streams"][1]["parse"][i])
    # the same packets in a power-of-two ring, the first one wrapping around its end
    import torch

    ring = 1 << 13
    total = sum(len(p) for p in packets)
    assert total < ring
    start = ring - len(packets[0]) // 2
    arena = np.zeros(ring, dtype=np.uint8)
    sp = np.zeros(len(packets), dtype=anm.PB_SPAN_DTYPE)
    pos = start
    for i, p in enumerate(packets):
        for j, b in enumerate(p):
            arena[(pos + j) % ring] = b
        sp[i] = (anm.ANM_PB_OK, 0, pos & 0xFFFFFFFF, len(p))
        pos += len(p)
    d_sp = torch.from_numpy(sp.view(np.uint8)).cuda()
    d_by = torch.from_numpy(arena).cuda()
    d_out = torch.zeros(len(packets) * anm.OPUS_PACKET_DTYPE.itemsize, dtype=torch.uint8, device="cuda")
    rc = anm.lib().anm_opus_parse_device(d_sp.data_ptr(), len(packets), d_by.data_ptr(), ring - 1, 48000, d_out.data_ptr(), None)
    assert rc == 0
    torch.cuda.synchronize()
    res = d_out.cpu().numpy().view(anm.OPUS_PACKET_DTYPE)
    for g, ref in zip(res, GOLD["streams"][1]["parse"]):
        _equal(g, ref)


@pytest.mark.gpu
def test_modem_to_opus_work_list_end_to_end():
    """ToReceiver{AudioData{reference-encoder packet}} as frame payloads -> CUDA demodulator -> GPU deframer -> GPU packet
    parse: the work list equals the reference's parse of the packets that were sent."""
    import ctypes as C

    import torch

    cfg = anm.config_preset("ref4")
    cfg.max_payload = 1024
    s = GOLD["streams"][0]
    packets = [bytes.fromhex(h) for h in s["packets"][:8]]
    L = anm.lib()
    L.anm_pb_encode_to_receiver_audio.restype = C.c_size_t
    buf = (C.c_uint8 * 2048)()
    prog = [np.full(4, 255, np.uint8)]
    for p in packets:
        n = L.anm_pb_encode_to_receiver_audio(p, C.c_size_t(len(p)), buf, C.c_size_t(2048))
        prog += [anm.frame_symbols(cfg, bytes(buf[:n])), np.full(5, 255, np.uint8)]
    prog = np.concatenate(prog)
    n = (len(prog) + 8) * cfg.sym_len
    pcm = anm.tx_render(cfg, prog, anm.tx_params(seed=9, amplitude=0.5, snr_db=10.0), 0, n).reshape(1, -1)
    dm = anm.Demod(cfg, 1, device=0)
    d_pcm = torch.from_numpy(pcm).cuda()
    dm.feed_device(d_pcm.data_ptr(), n, n, torch.cuda.current_stream().cuda_stream)
    dm.collect()
    recs, by = dm.read_frames()
    dm.close()
    assert len(recs) == len(packets) and (recs["crc_ok"] == 1).all()
    spans = anm.pb_deframe(recs, by)
    assert (spans["status"] == anm.ANM_PB_OK).all()
    got = anm.opus_parse(spans, by)
    for g, ref, sp, p in zip(got, s["parse"], spans, packets):
        _equal(g, ref)
        assert bytes(by[sp["audio_offset"]: sp["audio_offset"] + sp["audio_len"]]) == p
