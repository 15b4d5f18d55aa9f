"""The product's presets, framer and transmitter stand-in (audio-network_b200/csrc/anm_config.c, anm_tx.c) against
the oracle's own restatement of SPEC 2 / 4 / 6 (oracle/anm_oracle_tx.c).  Two independently written
implementations of the transmit side: neither parity leg (oracle, CUDA path) depends on signals only the
other side's code could produce."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np
import pytest

import audio_network_b200 as anm
import oracle_binding as ob

PRESETS = ["ref4", "bfsk2", "mfsk8", "mfsk16", "wide64"]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("name", PRESETS)
def test_presets_are_the_same_table(name):
    a, b = anm.config_preset(name), ob.preset(name)
    assert bytes(a) == bytes(b)


@pytest.mark.parametrize("name", PRESETS)
def test_frame_symbols_equal(name):
    cfg = anm.config_preset(name)
    rng = np.random.default_rng(17)
    for ln in [1, 2, 3, 7, 31, 32, 33, 255, 256, 1023, 1024] + [int(x) for x in rng.integers(1, 1025, size=40)]:
        pl = rng.integers(0, 256, size=ln, dtype=np.uint8).tobytes()
        assert np.array_equal(anm.frame_symbols(cfg, pl), ob.frame_symbols(cfg, pl)), (name, ln)
    with pytest.raises(ValueError):
        ob.frame_symbols(cfg, b"")
    with pytest.raises(ValueError):
        ob.frame_symbols(cfg, bytes(cfg.max_payload + 1))


@pytest.mark.parametrize("name", ["ref4", "wide64"])
def test_transmitters_equal_sample_for_sample(name):
    cfg = anm.config_preset(name)
    rng = np.random.default_rng(23)
    for case in range(24):
        prog = rng.integers(0, cfg.n_tones, size=int(rng.integers(1, 60)), dtype=np.uint8)
        prog[rng.random(len(prog)) < 0.2] = anm.ANM_SILENCE
        kw = dict(seed=int(rng.integers(0, 2**63)), start_offset=int(rng.integers(-5000, 5000)), amplitude=float(rng.uniform(0.05, 0.99)),
                  snr_db=None if case % 4 == 0 else float(rng.uniform(-6, 30)), ppm=float(rng.uniform(-300, 300)) if case % 3 else 0.0)
        first = int(rng.integers(0, 1 << 20)) if case % 2 else int(rng.integers(1 << 33, 1 << 40))   # also far beyond 2^31 samples
        n = 3000
        a = anm.tx_render(cfg, prog, anm.tx_params(**kw), first, n)
        b = ob.tx_render(cfg, prog, ob.tx_params(**kw), first, n)
        assert np.array_equal(a, b), (name, case, kw, first)


def test_batch_render_equals_single():
    cfg = ob.preset("ref4")
    rng = np.random.default_rng(5)
    n_ch, n = 9, 2000
    progs = np.full((n_ch, 64), anm.ANM_SILENCE, dtype=np.uint8)
    lens = np.zeros(n_ch, dtype=np.uint32)
    plist = []
    for c in range(n_ch):
        ln = int(rng.integers(1, 64))
        progs[c, :ln] = rng.integers(0, 4, size=ln)
        lens[c] = ln
        plist.append(ob.tx_params(seed=c, start_offset=-int(rng.integers(0, 300)), snr_db=10.0))
    arr = np.zeros(n_ch, dtype=anm.TXPARAMS_DTYPE)
    for i, p in enumerate(plist):
        arr[i] = (p.seed, p.start_offset, p.amplitude_q15, p.snr_mdb, p.ppm_x1000, 0)
    out = ob.tx_render_batch(cfg, progs, lens, arr, 100, n, 3)
    for c in range(n_ch):
        assert np.array_equal(out[c], ob.tx_render(cfg, progs[c, : lens[c]], plist[c], 100, n))


def test_oracle_own_twiddles_equal_the_product_table():
    """SPEC 3: both sides derive the table from the same definition, separately."""
    for name in PRESETS:
        cfg = anm.config_preset(name)
        pcm = ob.tx_render(cfg, np.arange(cfg.n_tones, dtype=np.uint8), ob.tx_params(seed=1, snr_db=6.0), 0, 40 * cfg.sym_len)
        hops = len(pcm) // cfg.hop
        o1 = ob.Oracle(cfg, trace_hops=hops)
        o1.feed(pcm)
        o2 = ob.Oracle(cfg, trace_hops=hops, twiddles=anm.twiddles(cfg))
        o2.feed(pcm)
        assert np.array_equal(o1.E.view(np.uint32), o2.E.view(np.uint32))


def test_cpu_arm_does_not_map_the_product_library():
    """bench.py --impl reference must run on liboracle.so alone (the arm mapped libanmodem.so)."""
    code = (
        "import sys; sys.path[:0] = [%r, %r]\n"
        "import bench\n"
        "line = bench.reference_line(preset='ref4', n_ch=8, steps=1, warmup=0, threads=2)\n"
        "maps = open('/proc/self/maps').read()\n"
        "assert 'liboracle.so' in maps, 'oracle not loaded'\n"
        "assert 'libanmodem' not in maps, 'product library mapped by the CPU arm'\n"
        "assert line['impl'] == 'reference' and line['value'] > 0 and line['frames_ok'] > 0\n"
        "print('ok')\n" % (ROOT, os.path.join(ROOT, "tests"))
    )
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and out.stdout.strip().endswith("ok"), out.stdout + out.stderr
