#!/usr/bin/env python
"""BASELINE.json configs[4]: low-SNR sweep (0-3 dB) with +/-200 ppm clock error and random frame offsets.
Demodulates on the GPU (through the C ABI), counts frames sent / detected / CRC-valid per SNR, and checks
a sample of channels bit for bit against the CPU oracle.  Writes one JSON object.  Test infrastructure.
Usage: python tools/snr_sweep.py [out.json]"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)

import torch  # noqa: E402

import audio_network_b200 as anm  # noqa: E402
from oracle_binding import Oracle  # noqa: E402
from sigutil import make_program  # noqa: E402

cfg = anm.config_preset("ref4")
N = cfg.sym_len
n_ch, n_sym = 4096, 3000
snrs = [-12.0, -9.0, -6.0, -3.0, 0.0, 1.0, 2.0, 3.0]   # configs[4] asks for 0-3 dB; the lower rows show where the modem gives up
rng = np.random.default_rng(2026)
progs = np.full((n_ch, 4096), anm.ANM_SILENCE, dtype=np.uint8)
lens = np.zeros(n_ch, dtype=np.int32)
params, sent = [], []
for c in range(n_ch):
    prog, payloads = make_program(cfg, rng, 3400, payload_len=(16, 200), gap=(2, 24))
    prog = prog[:4096]
    progs[c, : len(prog)] = prog
    lens[c] = len(prog)
    off = int(rng.integers(0, 5000))
    params.append(anm.tx_params(seed=9000 + c, start_offset=-off, amplitude=0.1, snr_db=snrs[c % len(snrs)],
                                ppm=float(rng.uniform(-200.0, 200.0))))
    # payloads of the frames that end at least three symbol periods before the capture does (clock error
    # moves a frame by less than one symbol period over the capture)
    pos, k = [], 0
    i = 0
    while i < len(prog):
        if prog[i] == anm.ANM_SILENCE:
            i += 1
            continue
        ln = len(anm.frame_symbols(cfg, payloads[k]))
        if (i + ln) * N + off <= (n_sym - 3) * N:
            pos.append(payloads[k])
        i += ln
        k += 1
    sent.append(pos)
par = anm.tx_params_array(params)
dev = torch.device("cuda", 0)
d_prog, d_len = torch.from_numpy(progs).to(dev), torch.from_numpy(lens).to(dev)
d_par = torch.from_numpy(par.view(np.uint8).copy()).to(dev)
n = n_sym * N
d_pcm = torch.empty((n_ch, n), dtype=torch.int16, device=dev)
st = torch.cuda.current_stream().cuda_stream
anm.tx_render_device(cfg, d_prog.data_ptr(), progs.shape[1], d_len.data_ptr(), d_par.data_ptr(), n_ch, 0, d_pcm.data_ptr(), n, n, st)
torch.cuda.synchronize()
dm = anm.Demod(cfg, n_ch, device=0)
chunk = 344 * N
for pos in range(0, n, chunk):
    dm.feed_device(d_pcm.data_ptr() + pos * 2, n, min(chunk, n - pos), st)
dm.collect()
recs, by = dm.read_frames(cap=1 << 22, bytes_cap=1 << 28)
frames = anm.frames_to_list(recs, by)
dm.close()
rows = {}
for s_i, s in enumerate(snrs):
    chs = [c for c in range(n_ch) if c % len(snrs) == s_i]
    tx = sum(len(sent[c]) for c in chs)
    det = sum(1 for f in frames if f[0] % len(snrs) == s_i)
    ok = sum(1 for f in frames if f[0] % len(snrs) == s_i and f[2])
    good = {}
    for f in frames:
        if f[0] % len(snrs) == s_i and f[2]:
            good.setdefault(f[0], set()).add(f[3])
    delivered = sum(1 for c in chs for pl in sent[c] if pl in good.get(c, ()))
    rows["%g dB" % s] = {"channels": len(chs), "frames_sent_inside_capture": tx, "frames_delivered_crc_ok": delivered,
                         "delivery_rate": round(delivered / max(tx, 1), 4), "frames_detected_total": det,
                         "crc_pass_rate_of_detected": round(ok / max(det, 1), 4)}
# oracle check on a sample of channels (bit-exact frames)
sample = list(range(0, n_ch, 64))
pcm = d_pcm[sample].cpu().numpy()
mism = 0
for j, c in enumerate(sample):
    o = Oracle(cfg)
    o.feed(pcm[j])
    want = o.frames(c)
    got = [f for f in frames if f[0] == c]
    mism += int(got != want)
out = {"config": "BASELINE configs[4]: preset ref4, %d channels x %d symbol periods, +/-200 ppm, offsets < 5000 samples, payloads 16-200 B" % (n_ch, n_sym),
       "by_snr": rows, "oracle_checked_channels": len(sample), "oracle_mismatching_channels": mism}
print(json.dumps(out, indent=1))
if len(sys.argv) > 1:
    with open(sys.argv[1], "w") as f:
        json.dump(out, f, indent=1)
assert mism == 0
