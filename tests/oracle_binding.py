"""ctypes binding of oracle/liboracle.so and oracle/_ref/libref_nanopb.so.

TEST INFRASTRUCTURE: imported only by tests/, __graft_entry__.smoke() and bench.py's
CPU-baseline legs (see oracle/anm_oracle.h).

This module never loads the product library (libanmodem.so): the oracle computes its own
twiddles, presets, frames and test signals (oracle/anm_oracle_tx.c), and the record layouts
below are restated from include/anmodem.h.  `import audio_network_b200` is needed only for
its ctypes / numpy record types, which load nothing; bench.py --impl reference therefore maps
liboracle.so and nothing of the product.
"""
import ctypes as C
import os

import numpy as np

import audio_network_b200 as anm   # types only (lazy loader: importing does not map libanmodem.so)

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_LIB = os.path.join(ROOT, "oracle", "liboracle.so")
REF_LIB = os.path.join(ROOT, "oracle", "_ref", "libref_nanopb.so")

_o = None


def olib():
    global _o
    if _o is None:
        if not os.path.exists(ORACLE_LIB):
            raise RuntimeError("oracle/liboracle.so missing: run `make -C oracle`")
        L = C.CDLL(ORACLE_LIB)
        vp = C.c_void_p
        L.anm_oracle_create.restype = vp
        L.anm_oracle_create.argtypes = [vp, vp]
        L.anm_oracle_reset.argtypes = [vp]
        L.anm_oracle_destroy.argtypes = [vp]
        L.anm_oracle_set_trace.argtypes = [vp, vp, vp, vp, C.c_size_t]
        L.anm_oracle_feed.argtypes = [vp, vp, C.c_size_t]
        for n in ("anm_oracle_num_frames", "anm_oracle_num_bytes", "anm_oracle_num_symbols"):
            getattr(L, n).restype = C.c_size_t
            getattr(L, n).argtypes = [vp]
        for n in ("anm_oracle_frames", "anm_oracle_bytes", "anm_oracle_symbols"):
            getattr(L, n).restype = vp
            getattr(L, n).argtypes = [vp]
        L.anm_oracle_stats.argtypes = [vp, vp]
        L.anm_oracle_run_batch.restype = C.c_double
        L.anm_oracle_run_batch.argtypes = [vp, vp, vp, C.c_uint32, C.c_size_t, C.c_size_t, C.c_uint32,
                                           C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
        L.anm_oracle_preset.argtypes = [C.c_char_p, vp]
        L.anm_oracle_frame_symbols.restype = C.c_size_t
        L.anm_oracle_frame_symbols.argtypes = [vp, vp, C.c_size_t, vp, C.c_size_t]
        L.anm_oracle_tx_render.argtypes = [vp, vp, C.c_size_t, vp, C.c_uint64, vp, C.c_size_t]
        L.anm_oracle_tx_render_batch.argtypes = [vp, vp, C.c_size_t, vp, vp, C.c_uint32, C.c_uint64, vp, C.c_size_t, C.c_size_t, C.c_uint32]
        _o = L
    return _o


class Oracle:
    """One channel of the sequential CPU oracle."""

    def __init__(self, cfg, trace_hops=0, twiddles=None):
        self.cfg = cfg
        self.tw = None if twiddles is None else np.ascontiguousarray(twiddles, dtype=np.float32)   # None: the oracle's own table (SPEC 3)
        self._h = olib().anm_oracle_create(C.byref(cfg), None if self.tw is None else self.tw.ctypes.data_as(C.c_void_p))
        self.trace_hops = trace_hops
        if trace_hops:
            self.E = np.zeros((trace_hops, cfg.n_tones), dtype=np.float32)
            self.D = np.zeros(trace_hops, dtype=np.uint8)
            self.Emax = np.zeros(trace_hops, dtype=np.float32)
            olib().anm_oracle_set_trace(self._h, self.E.ctypes.data, self.D.ctypes.data, self.Emax.ctypes.data, trace_hops)

    def __del__(self):
        if getattr(self, "_h", None):
            olib().anm_oracle_destroy(self._h)
            self._h = None

    def feed(self, pcm):
        pcm = np.ascontiguousarray(pcm, dtype=np.int16)
        olib().anm_oracle_feed(self._h, pcm.ctypes.data, len(pcm))

    def frames(self, channel=0):
        n = olib().anm_oracle_num_frames(self._h)
        if n == 0:
            return []
        recs = np.ctypeslib.as_array(C.cast(olib().anm_oracle_frames(self._h), C.POINTER(C.c_uint8)), shape=(n * anm.FRAME_DTYPE.itemsize,)).view(anm.FRAME_DTYPE).copy()
        nb = olib().anm_oracle_num_bytes(self._h)
        by = np.ctypeslib.as_array(C.cast(olib().anm_oracle_bytes(self._h), C.POINTER(C.c_uint8)), shape=(max(nb, 1),)).copy()
        out = []
        for r in recs:
            o = int(r["offset"])
            out.append((channel, int(r["start_sample"]), int(r["crc_ok"]), bytes(by[o:o + int(r["len"])])))
        return out

    def symbols(self):
        n = olib().anm_oracle_num_symbols(self._h)
        if n == 0:
            return np.zeros(0, dtype=np.uint8)
        return np.ctypeslib.as_array(C.cast(olib().anm_oracle_symbols(self._h), C.POINTER(C.c_uint8)), shape=(n,)).copy()

    def stats(self):
        out = np.zeros(1, dtype=anm.STATS_DTYPE)
        olib().anm_oracle_stats(self._h, out.ctypes.data)
        return out[0]


def oracle_frames_batch(cfg, pcm):
    """Runs the oracle over pcm[n_ch, n]; returns the sorted frame list."""
    out = []
    for c in range(pcm.shape[0]):
        o = Oracle(cfg)
        o.feed(pcm[c])
        out.extend(o.frames(c))
    out.sort(key=lambda f: (f[0], f[1]))
    return out


def run_batch(cfg, pcm, n_threads):
    """(seconds, frames_ok, frames_bad, payload_bytes_ok, digest) of the threaded batch runner."""
    ok, bad, by, dg = C.c_uint64(), C.c_uint64(), C.c_uint64(), C.c_uint64()
    assert pcm.dtype == np.int16 and pcm.strides[1] == 2
    sec = olib().anm_oracle_run_batch(C.byref(cfg), None, pcm.ctypes.data, pcm.shape[0], pcm.strides[0] // 2,
                                      pcm.shape[1], n_threads, C.byref(ok), C.byref(bad), C.byref(by), C.byref(dg))
    return sec, ok.value, bad.value, by.value, dg.value


def frames_digest(frames):
    """Same order-independent digest as anm_oracle_batch.c, from a frame list."""
    MASK = (1 << 64) - 1

    def fnv(h, data):
        for b in data:
            h = ((h ^ b) * 0x100000001B3) & MASK
        return h

    total = 0
    for ch, start, ok, payload in frames:
        d = fnv(0xCBF29CE484222325 ^ ch, int(start).to_bytes(8, "little"))
        d = fnv(d, len(payload).to_bytes(4, "little"))
        d = fnv(d, int(ok).to_bytes(4, "little"))
        d = fnv(d, payload)
        total = (total + d) & MASK
    return total


# ---- the oracle's own transmit side (oracle/anm_oracle_tx.c): SPEC 2 / 4 / 6 restated independently of the product ----
def preset(name):
    cfg = anm.Config()
    if olib().anm_oracle_preset(name.encode(), C.byref(cfg)) != 0:
        raise ValueError("unknown preset %r" % name)
    return cfg


def frame_symbols(cfg, payload):
    pl = np.frombuffer(bytes(payload), dtype=np.uint8)
    out = np.empty(cfg.preamble_len + 8 * (len(pl) + 8), dtype=np.uint8)
    n = olib().anm_oracle_frame_symbols(C.byref(cfg), pl.ctypes.data, len(pl), out.ctypes.data, len(out))
    if n == 0:
        raise ValueError("invalid payload length %d" % len(pl))
    return out[:n].copy()


def tx_params(seed=0, start_offset=0, amplitude=0.5, snr_db=None, ppm=0.0):
    """anm_tx_params_t of include/anmodem.h (the noise scale in `reserved` is a GPU-renderer detail; the oracle derives it)"""
    p = anm.TxParams()
    p.seed, p.start_offset = seed, start_offset
    p.amplitude_q15 = int(round(amplitude * 32768))
    p.snr_mdb = anm.ANM_SNR_CLEAN if snr_db is None else int(round(snr_db * 1000))
    p.ppm_x1000 = int(round(ppm * 1000))
    return p


def tx_render(cfg, program, params, first_sample, n):
    prog = np.ascontiguousarray(program, dtype=np.uint8)
    out = np.empty(n, dtype=np.int16)
    rc = olib().anm_oracle_tx_render(C.byref(cfg), prog.ctypes.data, len(prog), C.byref(params), first_sample, out.ctypes.data, n)
    assert rc == 0
    return out


def tx_render_batch(cfg, progs, lens, params_arr, first_sample, n, n_threads):
    """progs uint8 [n_ch, stride], lens [n_ch], params_arr TXPARAMS_DTYPE [n_ch] -> int16 [n_ch, n]"""
    progs = np.ascontiguousarray(progs, dtype=np.uint8)
    lens = np.ascontiguousarray(lens, dtype=np.uint32)
    params_arr = np.ascontiguousarray(params_arr, dtype=anm.TXPARAMS_DTYPE)
    out = np.empty((progs.shape[0], n), dtype=np.int16)
    rc = olib().anm_oracle_tx_render_batch(C.byref(cfg), progs.ctypes.data, progs.shape[1], lens.ctypes.data, params_arr.ctypes.data,
                                           progs.shape[0], first_sample, out.ctypes.data, n, n, n_threads)
    assert rc == 0
    return out
