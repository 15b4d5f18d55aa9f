"""audio_network_b200 -- Python (ctypes) mirror of include/anmodem.h.

The product is libanmodem.so (sm_100a CUDA kernels behind a C ABI).  This module only
loads it and mirrors its functions with the same names and argument meaning, so tests and
bench.py read like calls to the C interface.  There is no Python or CPU implementation of
the receive path here: every compute call fails loudly (AnmError) when the shared object
or a CUDA device is missing.

Reference seam being mirrored: SURVEY.md section 8(b) -- the pb_istream_t byte source of
hardware/src/network.cpp:262-305 and the `<module>_initialize()` idiom of
hardware/README.md:10-14.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("ANM_LIB_PATH") or os.path.join(_HERE, "libanmodem.so")   # ANM_LIB_PATH: experiment knob (variant builds)

ANM_MAX_TONES = 64
ANM_MAX_PREAMBLE = 32
ANM_SILENCE = 0xFF
ANM_SNR_CLEAN = 2**31 - 1
ANM_FLAG_SYMBOLS = 1

ANM_OK, ANM_ERR_ARG, ANM_ERR_CUDA, ANM_ERR_NOMEM, ANM_ERR_ALIGN, ANM_ERR_OVERFLOW, ANM_ERR_UNSUPPORTED, ANM_ERR_FORMAT = 0, -1, -2, -3, -4, -5, -6, -7


class AnmError(RuntimeError):
    def __init__(self, code, msg=""):
        super().__init__("anmodem error %d: %s" % (code, msg))
        self.code = code


class Config(C.Structure):
    _fields_ = [
        ("sample_rate", C.c_uint32),
        ("sym_len", C.c_uint32),
        ("hops_per_sym", C.c_uint32),
        ("n_tones", C.c_uint32),
        ("tone_bin", C.c_uint32 * ANM_MAX_TONES),
        ("preamble_len", C.c_uint32),
        ("preamble", C.c_uint8 * ANM_MAX_PREAMBLE),
        ("sync_tol", C.c_uint32),
        ("max_payload", C.c_uint32),
        ("trk_epoch", C.c_uint32),
        ("trk_thresh", C.c_uint32),
    ]

    @property
    def hop(self):
        return self.sym_len // self.hops_per_sym

    @property
    def bits_per_sym(self):
        return int(self.n_tones).bit_length() - 1


class Frame(C.Structure):
    _fields_ = [
        ("channel", C.c_uint32),
        ("len", C.c_uint32),
        ("start_sample", C.c_uint64),
        ("crc_ok", C.c_uint32),
        ("offset", C.c_uint32),
    ]


class ChanStats(C.Structure):
    _fields_ = [
        ("locks", C.c_uint32),
        ("header_fail", C.c_uint32),
        ("frames_ok", C.c_uint32),
        ("frames_bad", C.c_uint32),
        ("symbols", C.c_uint64),
        ("trk_moves", C.c_int32),
        ("reserved", C.c_uint32),
    ]


class TxParams(C.Structure):
    _fields_ = [
        ("seed", C.c_uint64),
        ("start_offset", C.c_int64),
        ("amplitude_q15", C.c_uint32),
        ("snr_mdb", C.c_int32),
        ("ppm_x1000", C.c_int32),
        ("reserved", C.c_uint32),
    ]


FRAME_DTYPE = np.dtype(
    [("channel", "<u4"), ("len", "<u4"), ("start_sample", "<u8"), ("crc_ok", "<u4"), ("offset", "<u4")]
)
TXPARAMS_DTYPE = np.dtype(
    [("seed", "<u8"), ("start_offset", "<i8"), ("amplitude_q15", "<u4"), ("snr_mdb", "<i4"), ("ppm_x1000", "<i4"), ("reserved", "<u4")]
)
STATS_DTYPE = np.dtype(
    [("locks", "<u4"), ("header_fail", "<u4"), ("frames_ok", "<u4"), ("frames_bad", "<u4"), ("symbols", "<u8"), ("trk_moves", "<i4"), ("reserved", "<u4")]
)


class PbDiscovery(C.Structure):
    """anm_pb_discovery_t (include/anmodem_pb.h): DiscoveryResponse of protocol/ip.proto:20-27"""
    _fields_ = [
        ("protocol_version", C.c_uint32),
        ("currently_streaming", C.c_uint8),
        ("pad", C.c_uint8 * 3),
        ("mac_address", C.c_uint64),
        ("device_name", C.c_char * 128),
        ("opus_version", C.c_char * 128),
    ]


class PbBroadcast(C.Structure):
    _fields_ = [
        ("magic_word", C.c_uint32),
        ("which", C.c_uint32),
        ("discovery_request", C.c_uint8),
        ("pad", C.c_uint8 * 7),
        ("discovery_response", PbDiscovery),
    ]


class PbToTransmitter(C.Structure):
    _fields_ = [
        ("which", C.c_uint32),
        ("max_encoded_frame_size", C.c_uint32),
        ("max_decoded_frame_size", C.c_uint32),
        ("audio_underflow", C.c_uint8),
        ("audio_decode_error", C.c_uint8),
        ("pad", C.c_uint8 * 2),
        ("discovery_data", PbDiscovery),
    ]


class Pacer(C.Structure):
    """anm_pacer_t (include/anmodem.h): the reference transmitter's LeakyBucket with an explicit clock"""
    _fields_ = [("capacity", C.c_int64), ("drain_rate_per_second", C.c_int64), ("last_value", C.c_int64), ("last_value_at_ns", C.c_int64)]

    def __init__(self, capacity=1200, drain_rate_per_second=1000, now_ns=0):
        super().__init__()
        _check(lib().anm_pacer_init(C.byref(self), capacity, drain_rate_per_second, now_ns))

    def level(self, now_ns):
        return lib().anm_pacer_level(C.byref(self), now_ns)

    def try_put(self, amount, now_ns):
        """None when added (as LeakyBucket.tryPut), else the nanoseconds to wait; AnmError where the reference throws."""
        r = lib().anm_pacer_try_put(C.byref(self), amount, now_ns)
        if r < 0:
            raise AnmError(int(r), "amount exceeds the bucket capacity")
        return None if r == 0 else int(r)

    def wait_for_capacity(self, amount, now_ns):
        """Virtual-clock waitForCapacity: returns (now_ns after the waits, total ns waited)."""
        t = C.c_int64(now_ns)
        r = lib().anm_pacer_wait_for_capacity(C.byref(self), amount, C.byref(t))
        if r < 0:
            raise AnmError(int(r), "amount exceeds the bucket capacity")
        return t.value, int(r)


# every symbol include/anmodem.h declares (checked by tests/test_abi.py)
EXPORTS = [
    "anm_config_preset", "anm_config_validate", "anm_twiddles", "anm_config_dense", "anm_basis_q7", "anm_config_foldable", "anm_fold_twiddles", "anm_crc16", "anm_crc8",
    "anm_frame_num_symbols", "anm_frame_symbols", "anm_tx_render", "anm_tx_render_device",
    "anm_tone_energies_device", "anm_demod_create", "anm_demod_destroy", "anm_demod_reset",
    "anm_demod_feed_device", "anm_demod_feed_device_chunks", "anm_demod_feed_host", "anm_demod_feed_host_async", "anm_demod_wait_input",
    "anm_demod_collect", "anm_demod_collect_upto", "anm_demod_read_frames", "anm_demod_take_frames", "anm_demod_frame_rings", "demod_create_cfg", "anm_frames_digest", "anm_demod_peek_frames", "anm_demod_drop_frames", "anm_frames_summary",
    "anm_demod_multi_create", "anm_demod_multi_destroy", "anm_demod_multi_reset", "anm_demod_multi_num_devices", "anm_demod_multi_shard",
    "anm_demod_multi_device_handle", "anm_demod_multi_alloc_pcm", "anm_demod_multi_feed_host", "anm_demod_multi_wait_input",
    "anm_demod_multi_collect", "anm_demod_multi_collect_upto", "anm_demod_multi_read_frames", "anm_demod_multi_take_frames",
    "anm_demod_multi_overflowed", "anm_demod_multi_stats",
    "anm_demod_read_symbols", "anm_demod_stats", "anm_demod_launch_count", "anm_demod_overflowed", "anm_demod_last_kernel_ms",
    "anm_last_error", "anm_version", "demod_initialize", "demod_create", "demod_feed",
    "demod_read_symbols", "demod_read_frames", "demod_destroy",
    "anm_pacer_init", "anm_pacer_level", "anm_pacer_try_put", "anm_pacer_wait_for_capacity",
    "anm_opus_parse_device", "anm_opus_parse_host",
    "anm_celt_tables_build", "anm_celt_ctx_create", "anm_celt_ctx_destroy", "anm_celt_entropy_device", "anm_celt_entropy_host", "anm_celt_spectrum_device", "anm_celt_spectrum_host", "anm_celt_decode_device", "anm_celt_decode_host", "anm_celt_synth_tables_build", "anm_celt_jobs_from_packets",
]

_lib = None


def lib():
    """Loads libanmodem.so (building is __graft_entry__.build()'s job); fails loudly if absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise AnmError(ANM_ERR_UNSUPPORTED, "%s is missing: run `python audio-network_b200/build.py` (no fallback exists)" % LIB_PATH)
    L = C.CDLL(LIB_PATH)
    vp, u8p, i16p, f32p, u32p = C.c_void_p, C.POINTER(C.c_uint8), C.POINTER(C.c_int16), C.POINTER(C.c_float), C.POINTER(C.c_uint32)
    cfgp = C.POINTER(Config)
    sig = {
        "anm_config_preset": (C.c_int, [C.c_char_p, cfgp]),
        "anm_config_validate": (C.c_int, [cfgp]),
        "anm_twiddles": (C.c_int, [cfgp, vp]),
        "anm_config_dense": (C.c_int, [cfgp]),
        "anm_config_foldable": (C.c_int, [cfgp]),
        "anm_fold_twiddles": (C.c_int, [cfgp, vp]),
        "anm_basis_q7": (C.c_int, [cfgp, vp]),
        "anm_crc16": (C.c_uint16, [vp, C.c_size_t, C.c_uint16]),
        "anm_crc8": (C.c_uint8, [vp, C.c_size_t, C.c_uint8]),
        "anm_frame_num_symbols": (C.c_size_t, [cfgp, C.c_size_t]),
        "anm_frame_symbols": (C.c_size_t, [cfgp, vp, C.c_size_t, vp, C.c_size_t]),
        "anm_tx_render": (C.c_int, [cfgp, vp, C.c_size_t, C.POINTER(TxParams), C.c_uint64, vp, C.c_size_t]),
        "anm_tx_render_device": (C.c_int, [cfgp, vp, C.c_size_t, vp, vp, C.c_uint32, C.c_uint64, vp, C.c_size_t, C.c_size_t, vp]),
        "anm_tx_params_prepare": (None, [vp, C.c_size_t]),
        "anm_tone_energies_device": (C.c_int, [cfgp, vp, C.c_uint32, C.c_size_t, C.c_size_t, vp, vp, vp, vp]),
        "anm_demod_create": (C.c_int, [cfgp, C.c_uint32, C.c_int, C.c_uint32, C.POINTER(vp)]),
        "anm_demod_destroy": (None, [vp]),
        "anm_demod_reset": (C.c_int, [vp]),
        "anm_demod_feed_device": (C.c_int, [vp, vp, C.c_size_t, C.c_size_t, vp]),
        "anm_demod_feed_device_chunks": (C.c_int, [vp, vp, C.c_size_t, C.c_size_t, C.c_size_t, C.c_uint32, vp]),
        "anm_demod_feed_host": (C.c_int, [vp, vp, C.c_size_t, C.c_size_t]),
        "anm_demod_feed_host_async": (C.c_int, [vp, vp, C.c_size_t, C.c_size_t]),
        "anm_demod_wait_input": (C.c_int, [vp]),
        "anm_demod_collect_upto": (C.c_long, [vp, C.c_uint32]),
        "anm_demod_collect": (C.c_long, [vp]),
        "anm_demod_read_frames": (C.c_size_t, [vp, vp, C.c_size_t, vp, C.c_size_t]),
        "anm_demod_take_frames": (C.c_size_t, [vp, vp, C.c_size_t, vp, C.c_size_t, C.POINTER(C.c_size_t)]),
        "anm_demod_frame_rings": (C.c_int, [vp, C.POINTER(vp), u32p, C.POINTER(vp), u32p]),
        "demod_create_cfg": (vp, [cfgp]),
        "anm_demod_peek_frames": (C.c_size_t, [vp, C.POINTER(vp), C.POINTER(vp), C.POINTER(C.c_size_t)]),
        "anm_demod_drop_frames": (None, [vp]),
        "anm_frames_summary": (None, [vp, C.c_size_t, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
        "anm_demod_multi_create": (C.c_int, [cfgp, C.c_uint32, C.POINTER(C.c_int), C.c_uint32, C.c_uint32, C.POINTER(vp)]),
        "anm_demod_multi_destroy": (None, [vp]),
        "anm_demod_multi_reset": (C.c_int, [vp]),
        "anm_demod_multi_num_devices": (C.c_uint32, [vp]),
        "anm_demod_multi_shard": (C.c_int, [vp, C.c_uint32, C.POINTER(C.c_int), u32p, u32p, C.POINTER(C.c_int)]),
        "anm_demod_multi_device_handle": (vp, [vp, C.c_uint32]),
        "anm_demod_multi_alloc_pcm": (C.c_int, [vp, C.c_size_t, C.POINTER(vp), C.POINTER(C.c_size_t)]),
        "anm_demod_multi_feed_host": (C.c_int, [vp, vp, C.c_size_t, C.c_size_t]),
        "anm_demod_multi_wait_input": (C.c_int, [vp]),
        "anm_demod_multi_collect": (C.c_long, [vp]),
        "anm_demod_multi_collect_upto": (C.c_long, [vp, C.c_uint32]),
        "anm_demod_multi_read_frames": (C.c_size_t, [vp, vp, C.c_size_t, vp, C.c_size_t]),
        "anm_demod_multi_take_frames": (C.c_size_t, [vp, vp, C.c_size_t, vp, C.c_size_t, C.POINTER(C.c_size_t)]),
        "anm_demod_multi_overflowed": (C.c_int, [vp]),
        "anm_demod_multi_stats": (C.c_int, [vp, vp]),
        "anm_frames_digest": (C.c_uint64, [vp, C.c_size_t, vp]),
        "anm_demod_read_symbols": (C.c_size_t, [vp, C.c_uint32, vp, C.c_size_t]),
        "anm_demod_stats": (C.c_int, [vp, vp]),
        "anm_demod_launch_count": (C.c_uint64, [vp]),
        "anm_demod_overflowed": (C.c_int, [vp]),
        "anm_demod_last_kernel_ms": (C.c_float, [vp]),
        "anm_demod_kernel_time": (C.c_int, [vp, C.POINTER(C.c_float)]),
        "anm_demod_launch_geometry": (C.c_int, [vp, u32p, u32p, u32p]),
        "anm_pb_deframe_device": (C.c_int, [vp, C.c_uint32, vp, C.c_uint32, vp, vp]),
        "anm_pb_deframe_host": (C.c_int, [vp, C.c_size_t, vp, C.c_size_t, vp]),
        "anm_opus_parse_device": (C.c_int, [vp, C.c_uint32, vp, C.c_uint32, C.c_int32, vp, vp]),
        "anm_opus_parse_host": (C.c_int, [vp, C.c_size_t, vp, C.c_size_t, C.c_int32, vp]),
        "anm_celt_tables_build": (C.c_int, [vp]),
        "anm_celt_ctx_create": (C.c_int, [C.c_int, C.POINTER(vp)]),
        "anm_celt_ctx_destroy": (None, [vp]),
        "anm_celt_entropy_device": (C.c_int, [vp, vp, vp, C.c_uint32, C.c_uint32, vp, C.c_uint32, vp, vp, vp]),
        "anm_celt_entropy_host": (C.c_int, [vp, vp, C.c_uint32, vp, C.c_size_t, vp, vp]),
        "anm_celt_spectrum_device": (C.c_int, [vp, vp, vp, C.c_uint32, C.c_uint32, vp, C.c_uint32, vp, vp, vp, C.c_uint32, vp, vp]),
        "anm_celt_spectrum_host": (C.c_int, [vp, vp, C.c_uint32, vp, C.c_size_t, vp, vp, vp, C.c_uint32, vp]),
        "anm_celt_decode_device": (C.c_int, [vp, vp, vp, C.c_uint32, C.c_uint32, vp, C.c_uint32, vp, vp, vp, vp, C.c_uint32, vp]),
        "anm_celt_decode_host": (C.c_int, [vp, vp, C.c_uint32, vp, C.c_size_t, vp, vp, vp, vp, C.c_uint32]),
        "anm_celt_synth_tables_build": (C.c_int, [vp]),
        "anm_celt_jobs_from_packets": (C.c_long, [vp, vp, C.c_size_t, vp, C.c_size_t, C.c_uint32, vp]),
        "anm_pb_encode_broadcast": (C.c_size_t, [C.POINTER(PbBroadcast), vp, C.c_size_t]),
        "anm_pb_encode_to_transmitter": (C.c_size_t, [C.POINTER(PbToTransmitter), vp, C.c_size_t]),
        "anm_pb_decode_broadcast": (C.c_int, [vp, C.c_size_t, C.POINTER(PbBroadcast), C.POINTER(C.c_size_t)]),
        "anm_pb_decode_to_transmitter": (C.c_int, [vp, C.c_size_t, C.POINTER(PbToTransmitter), C.POINTER(C.c_size_t)]),
        "anm_pb_firmware_discovery": (None, [C.c_uint64, C.c_char_p, C.POINTER(PbBroadcast)]),
        "anm_pacer_init": (C.c_int, [C.POINTER(Pacer), C.c_int64, C.c_int64, C.c_int64]),
        "anm_pacer_level": (C.c_int64, [C.POINTER(Pacer), C.c_int64]),
        "anm_pacer_try_put": (C.c_int64, [C.POINTER(Pacer), C.c_int64, C.c_int64]),
        "anm_pacer_wait_for_capacity": (C.c_int64, [C.POINTER(Pacer), C.c_int64, C.POINTER(C.c_int64)]),
        "anm_last_error": (C.c_char_p, []),
        "anm_version": (C.c_char_p, []),
        "demod_initialize": (C.c_int, [cfgp]),
        "demod_create": (vp, []),
        "demod_feed": (C.c_int, [vp, vp, C.c_size_t]),
        "demod_read_symbols": (C.c_size_t, [vp, vp, C.c_size_t]),
        "demod_read_frames": (C.c_size_t, [vp, vp, C.c_size_t]),
        "demod_destroy": (None, [vp]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = args
    _lib = L
    return L


def _check(rc):
    if rc < 0:
        raise AnmError(rc, (lib().anm_last_error() or b"").decode())
    return rc


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


# ---------------------------------------------------------------- host helpers
def config_preset(name):
    cfg = Config()
    _check(lib().anm_config_preset(name.encode(), C.byref(cfg)))
    return cfg


def twiddles(cfg):
    out = np.empty((cfg.sym_len, cfg.n_tones, 2), dtype=np.float32)
    _check(lib().anm_twiddles(C.byref(cfg), _ptr(out)))
    return out


PB_SPAN_DTYPE = np.dtype([("status", "<u4"), ("consumed", "<u4"), ("audio_offset", "<u4"), ("audio_len", "<u4")])
ANM_PB_OK, ANM_PB_FAIL, ANM_PB_NO_AUDIO, ANM_PB_CRC = 0, 1, 2, 3


def pb_deframe(recs, payload_bytes):
    """Batched GPU deframer (include/anmodem_pb.h): frame records + byte arena as returned by
    Demod.read_frames() -> one span record per frame locating the Opus bytes of its ToReceiver message."""
    recs = np.ascontiguousarray(recs, dtype=FRAME_DTYPE)
    by = np.ascontiguousarray(payload_bytes, dtype=np.uint8)
    out = np.zeros(len(recs), dtype=PB_SPAN_DTYPE)
    _check(lib().anm_pb_deframe_host(_ptr(recs) if len(recs) else None, len(recs), _ptr(by) if len(by) else None, len(by),
                                     _ptr(out) if len(recs) else None))
    return out


OPUS_PACKET_DTYPE = np.dtype([("count", "<i4"), ("toc", "u1"), ("channels", "u1"), ("pad", "u1", (2,)), ("mode", "<i4"), ("bandwidth", "<i4"),
                              ("samples_per_frame", "<i4"), ("payload_offset", "<i4"), ("nb_frames", "<i4"), ("nb_samples", "<i4"),
                              ("size", "<i2", (48,))])  # anm_opus_packet_t (include/anmodem_opus.h), 128 bytes
ANM_OPUS_BAD_ARG, ANM_OPUS_INVALID_PACKET = -1, -4


def opus_parse(spans, payload_bytes, fs=48000):
    """Batched Opus packet parse on the GPU (include/anmodem_opus.h): span records as returned by pb_deframe() + the byte
    arena -> one anm_opus_packet_t per span (TOC fields, frame count, frame sizes) as libopus' opus_packet_parse gives them."""
    spans = np.ascontiguousarray(spans, dtype=PB_SPAN_DTYPE)
    by = np.ascontiguousarray(payload_bytes, dtype=np.uint8)
    out = np.zeros(len(spans), dtype=OPUS_PACKET_DTYPE)
    _check(lib().anm_opus_parse_host(_ptr(spans) if len(spans) else None, len(spans), _ptr(by) if len(by) else None, len(by), fs,
                                     _ptr(out) if len(spans) else None))
    return out


CELT_BANDS = 21
CELT_JOB_DTYPE = np.dtype([("offset", "<u4"), ("len", "<u4"), ("channels", "u1"), ("lm", "u1"), ("end_band", "u1"), ("flags", "u1")])   # anm_celt_job_t
CELT_FRAME_DTYPE = np.dtype([("final_range", "<u4"), ("tell_bits", "<i4"), ("flags", "<u4"), ("pf_pitch", "<u2"), ("pf_gain_q", "u1"), ("pf_tapset", "u1"),
                             ("spread", "u1"), ("alloc_trim", "u1"), ("intensity", "u1"), ("coded_bands", "u1"), ("lm", "u1"), ("channels", "u1"), ("pad", "u1", (2,)),
                             ("pvq_codewords", "<u4"), ("pvq_pulses", "<u4"), ("pvq_index_xor", "<u4"), ("tf_res", "i1", (CELT_BANDS,)),
                             ("fine_quant", "u1", (CELT_BANDS,)), ("pulses", "<i2", (CELT_BANDS,)), ("band_e", "<i2", (2 * CELT_BANDS,))])    # anm_celt_frame_t
CELT_STREAM_DTYPE = np.dtype([("old_e", "<i2", (2 * CELT_BANDS,)), ("log_e1", "<i2", (2 * CELT_BANDS,)), ("log_e2", "<i2", (2 * CELT_BANDS,)),
                              ("rng", "<u4"), ("flags", "<u4")])                                                                                  # anm_celt_stream_t
CELT_SYNTH_DTYPE = np.dtype([("mem", "<i4", (2, 2048 + 120)), ("preemph_mem", "<i4", (2,)), ("pf_period", "<i4"), ("pf_period_old", "<i4"), ("pf_tapset", "<i4"),
                             ("pf_tapset_old", "<i4"), ("pf_gain", "<i2"), ("pf_gain_old", "<i2"), ("out_channels", "<u4")])                        # anm_celt_synth_t
CELT_SYNTH_TABLES_DTYPE = np.dtype([("window", "<i2", (120,)), ("trig", "<i2", (1800,)), ("fft_tw", "<i2", (960,)), ("bitrev", "<i2", (900,)), ("e_means", "i1", (25,)),
                                    ("pad", "i1", (3,))])                                                                                           # anm_celt_synth_tables_t
CELT_JOB_DISABLE_INV = 1
CELT_X_STRIDE = 1920            # int16 coefficients per frame in celt_spectrum()'s output: [channels][120 << lm], at most 2 x 960


def celt_entropy(jobs, stream_begin, payload_bytes, streams=None):
    """Batched CELT entropy decode on the GPU (include/anmodem_opus.h): jobs (CELT_JOB_DTYPE, stream-major), stream_begin (n_streams + 1 offsets
    into jobs), the byte arena -> (one CELT_FRAME_DTYPE record per job, the streams' carried state)."""
    jobs = np.ascontiguousarray(jobs, dtype=CELT_JOB_DTYPE)
    sb = np.ascontiguousarray(stream_begin, dtype=np.uint32)
    by = np.ascontiguousarray(payload_bytes, dtype=np.uint8)
    n_streams = len(sb) - 1
    st = np.zeros(n_streams, dtype=CELT_STREAM_DTYPE) if streams is None else np.ascontiguousarray(streams, dtype=CELT_STREAM_DTYPE).copy()
    out = np.zeros(len(jobs), dtype=CELT_FRAME_DTYPE)
    _check(lib().anm_celt_entropy_host(_ptr(jobs) if len(jobs) else None, _ptr(sb), n_streams, _ptr(by) if len(by) else None, len(by), _ptr(st),
                                       _ptr(out) if len(jobs) else None))
    return out, st


def celt_spectrum(jobs, stream_begin, payload_bytes, streams=None):
    """Batched CELT decode up to the normalised spectrum (stages 1 + 2, include/anmodem_opus.h anm_celt_spectrum_*): as celt_entropy(), plus
    x[n_jobs, CELT_X_STRIDE] int16 (Q14; frame j's channel c at x[j, c * (120 << lm):][:120 << lm]) and the collapse masks [n_jobs, 42]."""
    jobs = np.ascontiguousarray(jobs, dtype=CELT_JOB_DTYPE)
    sb = np.ascontiguousarray(stream_begin, dtype=np.uint32)
    by = np.ascontiguousarray(payload_bytes, dtype=np.uint8)
    n_streams = len(sb) - 1
    st = np.zeros(n_streams, dtype=CELT_STREAM_DTYPE) if streams is None else np.ascontiguousarray(streams, dtype=CELT_STREAM_DTYPE).copy()
    out = np.zeros(len(jobs), dtype=CELT_FRAME_DTYPE)
    x = np.zeros((max(len(jobs), 1), CELT_X_STRIDE), dtype=np.int16)
    cm = np.zeros((max(len(jobs), 1), 2 * CELT_BANDS), dtype=np.uint8)
    _check(lib().anm_celt_spectrum_host(_ptr(jobs) if len(jobs) else None, _ptr(sb), n_streams, _ptr(by) if len(by) else None, len(by), _ptr(st),
                                        _ptr(out) if len(jobs) else None, _ptr(x), CELT_X_STRIDE, _ptr(cm)))
    return out, st, x[:len(jobs)], cm[:len(jobs)]


def celt_jobs_from_packets(spans, packets, flags=0):
    """parse records -> CELT frame jobs (include/anmodem_opus.h anm_celt_jobs_from_packets): (jobs, first_job per packet)"""
    spans = np.ascontiguousarray(spans, dtype=PB_SPAN_DTYPE)
    packets = np.ascontiguousarray(packets, dtype=OPUS_PACKET_DTYPE)
    n = len(packets)
    first = np.zeros(max(n, 1), dtype=np.uint32)
    need = lib().anm_celt_jobs_from_packets(_ptr(spans) if n else None, _ptr(packets) if n else None, n, None, 0, flags, _ptr(first))
    if need < 0:
        _check(int(need))
    jobs = np.zeros(max(need, 1), dtype=CELT_JOB_DTYPE)
    got = lib().anm_celt_jobs_from_packets(_ptr(spans) if n else None, _ptr(packets) if n else None, n, _ptr(jobs), need, flags, _ptr(first))
    assert got == need
    return jobs[:need], first[:n]


def celt_decode(jobs, stream_begin, payload_bytes, out_channels=None, streams=None, synth=None):
    """Batched CELT decode to PCM (all three stages, include/anmodem_opus.h anm_celt_decode_*): as celt_entropy(), plus pcm[n_jobs, CELT_X_STRIDE] int16
    (frame j: (120 << lm) x out_channels interleaved samples) and the streams' synthesis state.  out_channels: per-stream decoder channels (default: the
    channel count of the stream's first frame); a mono decoder sets CELT_JOB_DISABLE_INV on its jobs itself."""
    jobs = np.ascontiguousarray(jobs, dtype=CELT_JOB_DTYPE)
    sb = np.ascontiguousarray(stream_begin, dtype=np.uint32)
    by = np.ascontiguousarray(payload_bytes, dtype=np.uint8)
    n_streams = len(sb) - 1
    st = np.zeros(n_streams, dtype=CELT_STREAM_DTYPE) if streams is None else np.ascontiguousarray(streams, dtype=CELT_STREAM_DTYPE).copy()
    sy = np.zeros(n_streams, dtype=CELT_SYNTH_DTYPE) if synth is None else np.ascontiguousarray(synth, dtype=CELT_SYNTH_DTYPE).copy()
    if out_channels is not None:
        sy["out_channels"] = np.asarray(out_channels, dtype=np.uint32)
    out = np.zeros(len(jobs), dtype=CELT_FRAME_DTYPE)
    pcm = np.zeros((max(len(jobs), 1), CELT_X_STRIDE), dtype=np.int16)
    _check(lib().anm_celt_decode_host(_ptr(jobs) if len(jobs) else None, _ptr(sb), n_streams, _ptr(by) if len(by) else None, len(by), _ptr(st), _ptr(sy),
                                      _ptr(out) if len(jobs) else None, _ptr(pcm), CELT_X_STRIDE))
    return out, st, sy, pcm[:len(jobs)]


def pb_encode_broadcast(m):
    buf = (C.c_uint8 * 512)()
    n = lib().anm_pb_encode_broadcast(C.byref(m), buf, 512)
    return bytes(buf[:n])


def pb_encode_to_transmitter(m):
    buf = (C.c_uint8 * 512)()
    n = lib().anm_pb_encode_to_transmitter(C.byref(m), buf, 512)
    return bytes(buf[:n])


def pb_decode_broadcast(data):
    """-> (PbBroadcast, consumed) or None where the reference's pb_decode_delimited returns false"""
    m, used = PbBroadcast(), C.c_size_t(0)
    b = (C.c_uint8 * max(1, len(data))).from_buffer_copy(bytes(data) or b"\0")
    rc = lib().anm_pb_decode_broadcast(b, len(data), C.byref(m), C.byref(used))
    return (m, used.value) if rc == ANM_OK else None


def pb_decode_to_transmitter(data):
    m, used = PbToTransmitter(), C.c_size_t(0)
    b = (C.c_uint8 * max(1, len(data))).from_buffer_copy(bytes(data) or b"\0")
    rc = lib().anm_pb_decode_to_transmitter(b, len(data), C.byref(m), C.byref(used))
    return (m, used.value) if rc == ANM_OK else None


def config_dense(cfg):
    """SPEC 3b: True if the configuration uses the dense integer basis (tensor-core contraction path)."""
    return bool(lib().anm_config_dense(C.byref(cfg)))


def config_foldable(cfg):
    """SPEC 3: True if the hop partials are computed by centre folding."""
    return bool(lib().anm_config_foldable(C.byref(cfg)))


def fold_twiddles(cfg):
    out = np.empty((cfg.hop // 2, cfg.n_tones, 2), dtype=np.float32)
    _check(lib().anm_fold_twiddles(C.byref(cfg), _ptr(out)))
    return out


def basis_q7(cfg):
    out = np.empty((cfg.sym_len, cfg.n_tones, 2), dtype=np.int8)
    _check(lib().anm_basis_q7(C.byref(cfg), _ptr(out)))
    return out


def crc16(data, crc=0xFFFF):
    b = np.frombuffer(bytes(data), dtype=np.uint8)
    return lib().anm_crc16(_ptr(b) if len(b) else None, len(b), crc)


def crc8(data, crc=0):
    b = np.frombuffer(bytes(data), dtype=np.uint8)
    return lib().anm_crc8(_ptr(b) if len(b) else None, len(b), crc)


def frame_symbols(cfg, payload):
    """Tone indices (preamble + header + body) of a frame carrying `payload`."""
    pl = np.frombuffer(bytes(payload), dtype=np.uint8)
    n = lib().anm_frame_num_symbols(C.byref(cfg), len(pl))
    if n == 0:
        raise AnmError(ANM_ERR_ARG, "invalid payload length %d" % len(pl))
    out = np.empty(n, dtype=np.uint8)
    got = lib().anm_frame_symbols(C.byref(cfg), _ptr(pl), len(pl), _ptr(out), n)
    if got != n:
        raise AnmError(ANM_ERR_ARG, "frame_symbols failed")
    return out


def tx_params(seed=0, start_offset=0, amplitude=0.5, snr_db=None, ppm=0.0):
    p = TxParams()
    p.seed = seed
    p.start_offset = start_offset
    p.amplitude_q15 = int(round(amplitude * 32768))
    p.snr_mdb = ANM_SNR_CLEAN if snr_db is None else int(round(snr_db * 1000))
    p.ppm_x1000 = int(round(ppm * 1000))
    return p


def tx_render(cfg, program, params, first_sample, n):
    """CPU transmitter stand-in: int16 PCM of rx samples [first_sample, first_sample + n)."""
    prog = np.ascontiguousarray(program, dtype=np.uint8)
    out = np.empty(n, dtype=np.int16)
    _check(lib().anm_tx_render(C.byref(cfg), _ptr(prog), len(prog), C.byref(params), first_sample, _ptr(out), n))
    return out


def tx_params_array(plist):
    arr = np.zeros(len(plist), dtype=TXPARAMS_DTYPE)
    for i, p in enumerate(plist):
        arr[i] = (p.seed, p.start_offset, p.amplitude_q15, p.snr_mdb, p.ppm_x1000, 0)
    lib().anm_tx_params_prepare(_ptr(arr), len(arr))
    return arr


def tx_render_device(cfg, d_programs, prog_stride, d_prog_len, d_params, n_ch, first_sample, d_pcm, ch_stride, n, stream=0):
    """GPU renderer; all d_* are raw device pointers (ints)."""
    _check(lib().anm_tx_render_device(C.byref(cfg), d_programs, prog_stride, d_prog_len, d_params, n_ch, first_sample, d_pcm, ch_stride, n, stream))


def tone_energies_device(cfg, d_pcm, n_ch, ch_stride, n_samples, d_energy=None, d_sym=None, d_emax=None, stream=0):
    _check(lib().anm_tone_energies_device(C.byref(cfg), d_pcm, n_ch, ch_stride, n_samples, d_energy, d_sym, d_emax, stream))


class Demod:
    """Batched streaming demodulator handle (anm_demod_*)."""

    def __init__(self, cfg, n_channels, device=0, flags=0):
        self.cfg = cfg
        self.n_channels = n_channels
        self._h = C.c_void_p()
        _check(lib().anm_demod_create(C.byref(cfg), n_channels, device, flags, C.byref(self._h)))

    def close(self):
        if self._h:
            lib().anm_demod_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def reset(self):
        _check(lib().anm_demod_reset(self._h))

    def feed_device(self, d_pcm, ch_stride, n_samples, stream=0):
        _check(lib().anm_demod_feed_device(self._h, d_pcm, ch_stride, n_samples, stream))

    def feed_device_chunks(self, d_pcm, ch_stride, chunk_stride, n_samples, n_chunks, stream=0):
        """n_chunks resident chunks per channel in one launch (include/anmodem.h anm_demod_feed_device_chunks)"""
        _check(lib().anm_demod_feed_device_chunks(self._h, d_pcm, ch_stride, chunk_stride, n_samples, n_chunks, stream))

    def feed_host(self, pcm):
        """pcm: int16 array [n_channels, n_samples] (C-contiguous rows)."""
        assert pcm.dtype == np.int16 and pcm.ndim == 2 and pcm.shape[0] == self.n_channels
        assert pcm.strides[1] == 2
        _check(lib().anm_demod_feed_host(self._h, _ptr(pcm), pcm.strides[0] // 2, pcm.shape[1]))

    def feed_host_ptr(self, ptr, ch_stride, n_samples):
        _check(lib().anm_demod_feed_host(self._h, ptr, ch_stride, n_samples))

    def feed_host_async_ptr(self, ptr, ch_stride, n_samples):
        """Enqueue copy + kernel and return; the host buffer must stay untouched until wait_input()."""
        _check(lib().anm_demod_feed_host_async(self._h, ptr, ch_stride, n_samples))

    def wait_input(self):
        _check(lib().anm_demod_wait_input(self._h))

    def collect_upto(self, lag):
        return _check(lib().anm_demod_collect_upto(self._h, lag))

    def collect(self):
        return _check(lib().anm_demod_collect(self._h))

    def read_frames(self, cap=1 << 16, bytes_cap=1 << 24):
        """Returns (records ndarray FRAME_DTYPE, bytes ndarray) in (channel, start_sample) order."""
        recs = np.zeros(cap, dtype=FRAME_DTYPE)
        by = np.zeros(bytes_cap, dtype=np.uint8)
        n = lib().anm_demod_read_frames(self._h, _ptr(recs), cap, _ptr(by), bytes_cap)
        recs = recs[:n]
        used = int(recs["len"].sum()) if n else 0
        return recs, by[:used]

    def take_frames(self, recs, by):
        """Everything queued, arrival order, into caller-owned arrays (FRAME_DTYPE records, uint8 bytes); -> (n_frames, n_bytes)."""
        nb = C.c_size_t(0)
        n = lib().anm_demod_take_frames(self._h, _ptr(recs), len(recs), _ptr(by), len(by), C.byref(nb))
        return int(n), int(nb.value)

    def peek_frames(self):
        """Zero-copy views (FRAME_DTYPE records, uint8 bytes) of everything queued, arrival order; valid until the next
        collect / feed.  Call drop_frames() when done."""
        f, b, nb = C.c_void_p(), C.c_void_p(), C.c_size_t(0)
        n = lib().anm_demod_peek_frames(self._h, C.byref(f), C.byref(b), C.byref(nb))
        if n == 0:
            return np.zeros(0, dtype=FRAME_DTYPE), np.zeros(0, dtype=np.uint8)
        recs = np.ctypeslib.as_array(C.cast(f.value, C.POINTER(C.c_uint8)), shape=(n * FRAME_DTYPE.itemsize,)).view(FRAME_DTYPE)
        by = np.ctypeslib.as_array(C.cast(b.value, C.POINTER(C.c_uint8)), shape=(max(nb.value, 1),))[: nb.value]
        return recs, by

    def drop_frames(self):
        lib().anm_demod_drop_frames(self._h)

    def frame_rings(self):
        """(d_frames ptr, frames_mask, d_bytes ptr, bytes_mask) of the device-resident rings."""
        f, b, fm, bm = C.c_void_p(), C.c_void_p(), C.c_uint32(), C.c_uint32()
        _check(lib().anm_demod_frame_rings(self._h, C.byref(f), C.byref(fm), C.byref(b), C.byref(bm)))
        return f.value, fm.value, b.value, bm.value

    def read_symbols(self, channel, cap=1 << 20):
        out = np.zeros(cap, dtype=np.uint8)
        n = lib().anm_demod_read_symbols(self._h, channel, _ptr(out), cap)
        return out[:n]

    def stats(self):
        out = np.zeros(self.n_channels, dtype=STATS_DTYPE)
        _check(lib().anm_demod_stats(self._h, _ptr(out)))
        return out

    def overflowed(self):
        return bool(lib().anm_demod_overflowed(self._h))

    def launch_count(self):
        return int(lib().anm_demod_launch_count(self._h))

    def kernel_time(self):
        """(sum of kernel device time in ms, launches) since the previous call."""
        ms = C.c_float()
        n = _check(lib().anm_demod_kernel_time(self._h, C.byref(ms)))
        return float(ms.value), n

    def launch_geometry(self):
        g, w, s = C.c_uint32(), C.c_uint32(), C.c_uint32()
        _check(lib().anm_demod_launch_geometry(self._h, C.byref(g), C.byref(w), C.byref(s)))
        return g.value, w.value, s.value


def frames_digest(recs, by):
    """anm_frames_digest over read_frames / take_frames output (order-independent)."""
    recs = np.ascontiguousarray(recs, dtype=FRAME_DTYPE)
    by = np.ascontiguousarray(by, dtype=np.uint8)
    return int(lib().anm_frames_digest(_ptr(recs) if len(recs) else None, len(recs), _ptr(by) if len(by) else None))


class DemodMulti:
    """Several GPUs behind one handle (anm_demod_multi_*): channels shard over `devices`, one host thread per device."""

    def __init__(self, cfg, n_channels, devices, flags=0):
        self.cfg, self.n_channels = cfg, n_channels
        devs = (C.c_int * len(devices))(*devices)
        self._h = C.c_void_p()
        _check(lib().anm_demod_multi_create(C.byref(cfg), n_channels, devs, len(devices), flags, C.byref(self._h)))

    def close(self):
        if self._h:
            lib().anm_demod_multi_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def shards(self):
        out = []
        for d in range(lib().anm_demod_multi_num_devices(self._h)):
            dev, first, cnt, node = C.c_int(), C.c_uint32(), C.c_uint32(), C.c_int()
            _check(lib().anm_demod_multi_shard(self._h, d, C.byref(dev), C.byref(first), C.byref(cnt), C.byref(node)))
            out.append({"device": dev.value, "first_channel": first.value, "n_channels": cnt.value, "numa_node": node.value})
        return out

    def alloc_pcm(self, n_samples):
        """Page-locked int16 array [n_channels, n_samples] (a view with the library's row stride), shards placed next to their GPUs."""
        p, st = C.c_void_p(), C.c_size_t()
        _check(lib().anm_demod_multi_alloc_pcm(self._h, n_samples, C.byref(p), C.byref(st)))
        buf = np.ctypeslib.as_array(C.cast(p.value, C.POINTER(C.c_int16)), shape=(self.n_channels, st.value))
        return buf[:, :n_samples]

    def feed_host(self, pcm):
        assert pcm.dtype == np.int16 and pcm.ndim == 2 and pcm.shape[0] == self.n_channels and pcm.strides[1] == 2
        _check(lib().anm_demod_multi_feed_host(self._h, pcm.ctypes.data, pcm.strides[0] // 2, pcm.shape[1]))

    def wait_input(self):
        _check(lib().anm_demod_multi_wait_input(self._h))

    def reset(self):
        _check(lib().anm_demod_multi_reset(self._h))

    def collect(self):
        return _check(lib().anm_demod_multi_collect(self._h))

    def collect_upto(self, lag):
        return _check(lib().anm_demod_multi_collect_upto(self._h, lag))

    def read_frames(self, cap=1 << 16, bytes_cap=1 << 24):
        recs = np.zeros(cap, dtype=FRAME_DTYPE)
        by = np.zeros(bytes_cap, dtype=np.uint8)
        n = lib().anm_demod_multi_read_frames(self._h, _ptr(recs), cap, _ptr(by), bytes_cap)
        recs = recs[:n]
        return recs, by[: int(recs["len"].sum()) if n else 0]

    def take_frames(self, recs, by):
        nb = C.c_size_t(0)
        n = lib().anm_demod_multi_take_frames(self._h, _ptr(recs), len(recs), _ptr(by), len(by), C.byref(nb))
        return int(n), int(nb.value)

    def overflowed(self):
        return bool(lib().anm_demod_multi_overflowed(self._h))

    def stats(self):
        out = np.zeros(self.n_channels, dtype=STATS_DTYPE)
        _check(lib().anm_demod_multi_stats(self._h, _ptr(out)))
        return out


def frames_summary(recs):
    """(CRC-valid frames, their payload bytes) of a record array, one pass in C."""
    recs = np.ascontiguousarray(recs, dtype=FRAME_DTYPE)
    ok, by = C.c_uint64(0), C.c_uint64(0)
    lib().anm_frames_summary(_ptr(recs) if len(recs) else None, len(recs), C.byref(ok), C.byref(by))
    return int(ok.value), int(by.value)


def frames_to_list(recs, by):
    """[(channel, start_sample, crc_ok, payload bytes)] from read_frames output."""
    out = []
    for r in recs:
        o = int(r["offset"])
        out.append((int(r["channel"]), int(r["start_sample"]), int(r["crc_ok"]), bytes(by[o:o + int(r["len"])])))
    return out
