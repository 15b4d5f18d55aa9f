"""The CPU oracle against the committed known-answer vectors, and its own invariants.
(Parity unpinned w.r.t. the reference -- see oracle/anm_oracle.c; these vectors pin SPEC.md.)"""
import hashlib
import json
import os

import numpy as np
import pytest

import audio_network_b200 as anm
from oracle_binding import Oracle, frames_digest, oracle_frames_batch, run_batch
from sigutil import make_channels

GOLD = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "modem_kat.npz"))


@pytest.mark.parametrize("name", ["ref4", "bfsk2", "mfsk16", "wide64"])
def test_oracle_matches_golden_vectors(name):
    cfg = anm.config_preset(name)
    pcm = GOLD[name + "_pcm"]
    want_frames = [tuple(r) for r in json.loads(str(GOLD[name + "_frames"]))]
    want_syms = json.loads(str(GOLD[name + "_symbols"]))
    want_sha = json.loads(str(GOLD[name + "_energy_sha256"]))
    hops = pcm.shape[1] // cfg.hop
    got = []
    for c in range(pcm.shape[0]):
        o = Oracle(cfg, trace_hops=hops)
        o.feed(pcm[c])
        got += [(c, s, ok, p.hex()) for (_, s, ok, p) in o.frames(c)]
        assert o.symbols().tolist() == want_syms[c]
        assert hashlib.sha256(o.E.tobytes()).hexdigest() == want_sha[c]
        if c == 0:
            assert np.array_equal(o.E[:64].view(np.uint32), GOLD[name + "_E0"].view(np.uint32))
    assert got == want_frames
    assert len(got) > 0 and all(f[2] == 1 for f in got)


def test_generator_reproduces_golden_pcm():
    cfg = anm.config_preset("ref4")
    pcm, _ = make_channels(cfg, 6, 420 * cfg.sym_len, seed=77, snr_db=8.0, ppm_max=150.0, offset_max=1500)
    assert np.array_equal(pcm, GOLD["ref4_pcm"])


def test_oracle_is_invariant_to_feed_sizes():
    cfg = anm.config_preset("ref4")
    pcm, _ = make_channels(cfg, 1, 300 * cfg.sym_len, seed=4, snr_db=5.0, ppm_max=100.0, offset_max=500)
    a = Oracle(cfg)
    a.feed(pcm[0])
    b = Oracle(cfg)
    pos, sizes, i = 0, [1, 7, 31, 32, 33, 1000, 129], 0
    while pos < pcm.shape[1]:
        ln = min(sizes[i % len(sizes)], pcm.shape[1] - pos)
        b.feed(pcm[0, pos:pos + ln])
        pos += ln
        i += 1
    assert a.frames() == b.frames() and np.array_equal(a.symbols(), b.symbols())


def test_tone_energy_of_a_pure_tone():
    cfg = anm.config_preset("ref4")
    N, S = cfg.sym_len, cfg.hops_per_sym
    x = anm.tx_render(cfg, np.array([3], dtype=np.uint8), anm.tx_params(amplitude=0.5), 0, 6 * N)
    o = Oracle(cfg, trace_hops=6 * S)
    o.feed(x)
    E = o.E[S - 1 + S]          # a full window inside the tone
    expect = (16384 * N / 2) ** 2
    assert abs(E[3] / expect - 1) < 2e-3
    assert E[:3].max() < 1e-6 * expect   # orthogonal tones
    assert (o.D[S:] == 3).all()


def test_drift_and_noise_are_survivable():
    cfg = anm.config_preset("ref4")
    pcm, meta = make_channels(cfg, 12, 1200 * cfg.sym_len, seed=21, snr_db=3.0, ppm_max=200.0, offset_max=3000,
                              payload_len=(100, 200))
    frames = oracle_frames_batch(cfg, pcm)
    ok = [f for f in frames if f[2]]
    assert len(ok) >= 0.9 * len(frames) > 0
    for ch, _s, _ok, payload in ok:
        assert payload in meta[ch][1]


def test_batch_runner_matches_sequential_and_digest():
    cfg = anm.config_preset("ref4")
    pcm, _ = make_channels(cfg, 9, 400 * cfg.sym_len, seed=8, snr_db=10.0)
    frames = oracle_frames_batch(cfg, pcm)
    sec, ok, bad, nbytes, dg = run_batch(cfg, pcm, 3)
    assert ok == sum(f[2] for f in frames) and bad == len(frames) - ok
    assert nbytes == sum(len(f[3]) for f in frames if f[2])
    assert dg == frames_digest(frames)
    assert sec > 0


def test_empty_and_tiny_inputs():
    cfg = anm.config_preset("ref4")
    o = Oracle(cfg)
    o.feed(np.zeros(0, dtype=np.int16))
    o.feed(np.zeros(5, dtype=np.int16))
    assert o.frames() == [] and len(o.symbols()) == 0


@pytest.mark.parametrize("name,S", [("ref4", 4), ("ref4", 2), ("ref4", 8), ("mfsk16", 4), ("bfsk2", 4)])
def test_energies_equal_the_sliding_dft(name, S):
    """SPEC 3 (centre-folded or direct form) is the sliding DFT at the tone bins: |W|^2 of a float64
    restatement, within fp32 rounding -- the folding signs and twiddle phases are right for every S."""
    cfg = anm.config_preset(name)
    cfg.hops_per_sym = S
    N, T, H = cfg.sym_len, cfg.n_tones, cfg.sym_len // S
    assert anm.config_foldable(cfg) == (S != 8)          # bins 10, 14 are not multiples of 8/2
    rng = np.random.default_rng(3)
    x = rng.integers(-20000, 20000, size=5 * N).astype(np.int16)
    hops = len(x) // H
    o = Oracle(cfg, trace_hops=hops)
    o.feed(x)
    xx = np.concatenate([np.zeros(N), x.astype(np.float64)])
    bins = np.array(cfg.tone_bin[:T], dtype=np.float64)
    for h in range(hops):
        n = np.arange((h + 1) * H - N, (h + 1) * H)
        w = (xx[N + n][:, None] * np.exp(-2j * np.pi * bins[None, :] * n[:, None] / N)).sum(axis=0)
        e64 = np.abs(w) ** 2
        assert np.max(np.abs(o.E[h] - e64)) <= 1e-5 * e64.max()


def test_spec_revision_and_goldens_are_frozen():
    """SPEC.md revision 2 is frozen: the committed golden files are the ones generated under it (SHA-256 recorded in
    tests/golden/SPEC_VERSION.json), SPEC.md and the library name the same revision."""
    import hashlib
    import json
    import re

    here = os.path.dirname(os.path.abspath(__file__))
    ver = json.load(open(os.path.join(here, "golden", "SPEC_VERSION.json")))
    for name, digest in ver["sha256"].items():
        assert hashlib.sha256(open(os.path.join(here, "golden", name), "rb").read()).hexdigest() == digest, \
            "%s changed: a change of the golden vectors needs a new SPEC revision with the old files kept" % name
    spec = open(os.path.join(os.path.dirname(here), "SPEC.md")).read()
    m = re.search(r"\*\*Revision (\d+) . FROZEN", spec)
    assert m and int(m.group(1)) == ver["spec_revision"]
    assert ("rev %d" % ver["spec_revision"]).encode() in anm.lib().anm_version()
