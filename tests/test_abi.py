"""The C-ABI library loads on a machine without a GPU, exports every symbol the public headers
declare, and refuses compute calls loudly instead of falling back to a CPU path."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import audio_network_b200 as anm

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared(header):
    text = open(os.path.join(ROOT, "include", header)).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    names = set(re.findall(r"\b((?:anm|demod)_[a-z0-9_]+)\s*\(", text))
    return {n for n in names if not n.endswith("_t")}


def test_library_exports_every_declared_symbol():
    L = anm.lib()
    declared = _declared("anmodem.h") | _declared("anmodem_pb.h") | _declared("anmodem_opus.h")
    assert len(declared) >= 35
    missing = [n for n in sorted(declared) if not hasattr(L, n)]
    assert not missing, "declared in include/*.h but not exported: %s" % missing
    for n in anm.EXPORTS:
        assert n in declared


def test_struct_layouts_match_the_header():
    assert C.sizeof(anm.Frame) == 24 and anm.FRAME_DTYPE.itemsize == 24
    assert C.sizeof(anm.ChanStats) == 32 and anm.STATS_DTYPE.itemsize == 32
    assert C.sizeof(anm.TxParams) == 32 and anm.TXPARAMS_DTYPE.itemsize == 32
    assert C.sizeof(anm.Config) == 4 * 4 + 64 * 4 + 4 + 32 + 4 * 4


def test_version_and_error_strings():
    assert b"anmodem" in anm.lib().anm_version()


def test_no_cpu_fallback_without_device():
    import torch

    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    cfg = anm.config_preset("ref4")
    with pytest.raises(anm.AnmError) as e:
        anm.Demod(cfg, 4, device=0)
    assert e.value.code == anm.ANM_ERR_CUDA
    assert "no CPU fallback" in str(e.value)
    pcm = np.zeros(128 * 4, dtype=np.int16)
    rc = anm.lib().anm_tone_energies_device(C.byref(cfg), pcm.ctypes.data, 1, 512, 512, None, None, None, None)
    assert rc in (anm.ANM_ERR_CUDA, anm.ANM_ERR_ALIGN)


def test_config_validation():
    for name in ("ref4", "bfsk2", "mfsk8", "mfsk16", "wide64"):
        cfg = anm.config_preset(name)
        assert anm.lib().anm_config_validate(C.byref(cfg)) == 0
        assert cfg.sym_len % cfg.hops_per_sym == 0
    bad = anm.config_preset("ref4")
    bad.tone_bin[1] = bad.tone_bin[0]
    assert anm.lib().anm_config_validate(C.byref(bad)) == anm.ANM_ERR_ARG
    bad = anm.config_preset("ref4")
    bad.hops_per_sym = 3
    assert anm.lib().anm_config_validate(C.byref(bad)) == anm.ANM_ERR_ARG
    bad = anm.config_preset("ref4")
    bad.tone_bin[0] = 64
    assert anm.lib().anm_config_validate(C.byref(bad)) == anm.ANM_ERR_ARG
    with pytest.raises(anm.AnmError):
        anm.config_preset("nope")
