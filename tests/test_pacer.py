"""Chunk pacer (include/anmodem.h anm_pacer_*) against a restatement of the
reference transmitter's LeakyBucket (oracle/leaky_bucket.py, LeakyBucket.kt:9-64) on random schedules,
and the MulticastAudioOutput instance: 1200 ms of receiver buffer draining 1000 ms per second."""
import os
import sys

import numpy as np
import pytest

import audio_network_b200 as anm

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
from leaky_bucket import LeakyBucket  # noqa: E402


def test_random_schedules_match_the_restated_reference():
    rng = np.random.default_rng(5)
    for trial in range(200):
        cap = int(rng.integers(1, 5000))
        rate = int(rng.choice([1, 7, 1000, 44100, 48000, 10**6, 10**9]))
        t = int(rng.integers(0, 10**12))
        ref, got = LeakyBucket(cap, rate, t), anm.Pacer(cap, rate, t)
        for _ in range(100):
            t += int(rng.choice([0, 1, 999, 10**6, 20 * 10**6, 10**9, int(rng.integers(0, 3 * 10**9))]))
            amount = int(rng.integers(0, cap + 1))
            assert got.level(t) == ref.current_value(t)
            r, g = ref.try_put(amount, t), got.try_put(amount, t)
            assert g == (r if r != 0 else 1) if r is not None else g is None  # a zero-length wait is reported as 1 ns
            assert (got.last_value, got.last_value_at_ns) == (ref.last_value, ref.last_value_at_nanos)
        with pytest.raises(anm.AnmError):
            got.try_put(cap + 1, t)  # LeakyBucket.tryPut throws IllegalArgumentException (LeakyBucket.kt:37-39)


def test_reference_instance_paces_60ms_frames_to_real_time():
    """MulticastAudioOutput.kt:85-96: capacity 1200 ms, 1000 ms/s, one put of the Opus frame duration (60 ms) per frame:
    the first 20 frames go out at once, after that one frame per 60 ms."""
    p, ref = anm.Pacer(1200, 1000, 0), LeakyBucket(1200, 1000, 0)
    t = tr = 0
    sent_at = []
    for _ in range(120):
        t, _w = p.wait_for_capacity(60, t)
        tr = ref.wait_for_capacity(60, tr)
        assert t == tr
        sent_at.append(t)
    assert sent_at[:20] == [0] * 20
    gaps = np.diff(sent_at[20:])
    assert gaps.min() >= 59_000_000 and gaps.max() <= 61_000_000
    assert abs(sent_at[-1] - 100 * 60_000_000) <= 60_000_000  # 120 frames = 7.2 s of audio sent within 1.2 s of real time
    assert p.level(sent_at[-1]) <= 1200


def test_argument_errors():
    import ctypes as C
    raw = anm.Pacer()
    L = anm.lib()
    assert L.anm_pacer_init(C.byref(raw), -1, 1000, 0) == anm.ANM_ERR_ARG
    assert L.anm_pacer_init(C.byref(raw), 10, 0, 0) == anm.ANM_ERR_ARG
    assert L.anm_pacer_init(None, 10, 10, 0) == anm.ANM_ERR_ARG


import pytest  # noqa: E402


@pytest.mark.gpu
def test_paced_feed_through_the_demodulator():
    """Row f4 on a real feed path: chunks of PCM are submitted to the CUDA demodulator under the transmitter's leaky bucket (1200 ms of receiver
    buffer, MulticastAudioOutput.kt:85) at 25x real time on a virtual clock -- the schedule is the bucket's (a burst of one buffer, then one chunk
    per chunk-duration / 25), the frames are the oracle's."""
    import numpy as np

    import audio_network_b200 as anm
    from oracle_binding import oracle_frames_batch
    from sigutil import make_channels

    cfg = anm.config_preset("ref4")
    chunk = 64 * cfg.sym_len                         # 8,192 samples = 185.76 ms of audio
    chunk_ms = 186
    pcm, _ = make_channels(cfg, 6, 20 * chunk, seed=91, snr_db=10.0, offset_max=500)
    factor = 25
    pacer = anm.Pacer(1200, 1000 * factor, 0)
    dm = anm.Demod(cfg, 6, device=0)
    now, submit_times, got = 0, [], []
    for k in range(20):
        now, _waited = pacer.wait_for_capacity(chunk_ms, now)      # virtual clock: returns the time at which the chunk may go
        submit_times.append(now)
        dm.feed_host(np.ascontiguousarray(pcm[:, k * chunk: (k + 1) * chunk]))
        dm.collect()
        got += anm.frames_to_list(*dm.read_frames())
    dm.close()
    assert sorted(got, key=lambda f: (f[0], f[1])) == oracle_frames_batch(cfg, pcm)
    burst = 1200 // chunk_ms                                        # chunks that fit the empty bucket at once
    assert all(t == 0 for t in submit_times[:burst]) and submit_times[burst] > 0
    steady = np.diff(submit_times[burst + 1:])
    assert np.all(np.abs(steady - chunk_ms * 1e6 / factor) <= 1e5)  # then one chunk per chunk_ms / factor of (virtual) time
