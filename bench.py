#!/usr/bin/env python
"""bench.py -- throughput of the receive-side demodulation hot path (BASELINE.json metric).

Workload: BASELINE.json configs[2] weak-scaled -- 65,536 channels x 60 s at 44.1 kHz, AWGN at 10 dB SNR, sharded over
8 GPUs = 8,192 channels per GPU, streamed in chunks of 352 symbol periods (45,056 samples, 1.02 s; a multiple of the
kernel's 32-symbol step).  One STEP = this GPU's whole shard of the config: 60 chunks x 8,192 channels (22.1 G samples,
44.3 GB of int16 PCM, all of it resident in HBM, every launch reads 738 MB it has never touched) through the demodulator
(PCM -> tone energies -> sync -> symbols -> frames + CRC), with the decoded frames delivered to the host every step.
PCM is synthetic (transmitter stand-in rendered on the GPU, per-channel payloads seeded by the global channel id).

  python bench.py [--gpus N --steps K --warmup W]          our arm (CUDA, through the C ABI)
  python bench.py --impl reference [...]                    CPU arm: the in-repo C oracle on all host cores (there is no
                                                            reference demodulator); maps liboracle.so only
Prints ONE JSON line on rank 0.
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (ROOT, os.path.join(ROOT, "tests")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

CHUNK_SYMS = int(os.environ.get("ANM_BENCH_CHUNK_SYMS", "352"))   # symbol periods per chunk (1.02 s at N=128): 11 steps of 32
CHUNKS_PER_STEP = int(os.environ.get("ANM_BENCH_CHUNKS", "60"))   # configs[2]: 60 s of stream per channel
CH_PER_GPU = 8192         # 65,536 channels / 8 GPUs
SNR_DB = 10.0
PAYLOAD = 32              # payload bytes per frame
METRIC = "demodulated Msamples/s"
SUSTAIN_S = float(os.environ.get("ANM_BENCH_SUSTAIN_S", "2.5"))   # length of the sustained-clock region
CPU_CHUNKS = 2            # chunks per step of the CPU arm (a bounded sample of the 60-chunk step)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--channels", type=int, default=CH_PER_GPU, help="channels per GPU")
    ap.add_argument("--preset", default="ref4")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-cfg4", action="store_true", help="skip the dense-tone-set (tensor-core) leg")
    ap.add_argument("--no-sustain", action="store_true", help="skip the >= 2.5 s sustained-clock region")
    ap.add_argument("--launch-chunks", type=int, default=0, help="resident chunks per launch of k_demod (anm_demod_feed_device_chunks); 0 = the whole step, "
                                                                 "1 = one launch per chunk as in round 1")
    return ap.parse_args()


def workload_config(preset, n_ch, world):
    """The `config` both arms print (same dict: the CPU arm runs bounded samples of exactly this workload)."""
    chunk = CHUNK_SYMS * (256 if preset == "wide64" else 128)
    return {"workload": "cfg3 weak-scaled: %d channels/GPU x %d chunks of %d samples (%d symbol periods) per step, 10 dB SNR, "
                        "preset %s, %d-byte payload frames" % (n_ch, CHUNKS_PER_STEP, chunk, CHUNK_SYMS, preset, PAYLOAD),
            "channels_per_gpu": n_ch, "channels_total": world * n_ch, "chunk_samples": chunk, "chunks_per_step": CHUNKS_PER_STEP,
            "snr_db": SNR_DB, "preset": preset,
            "l2_policy": "inputs larger than L2 (%.0f MB per launch, %d distinct chunks = %.1f GB resident per GPU)"
                         % (n_ch * chunk * 2 / 1e6, CHUNKS_PER_STEP, n_ch * chunk * 2 * CHUNKS_PER_STEP / 1e9)}


def measured_traffic():
    """DRAM bytes per launch of the dominant kernel from the committed ncu --set full capture."""
    for name in ("r2_k_demod_traffic.json", "r1_k_demod_traffic.json"):
        path = os.path.join(ROOT, "profiles", name)
        if os.path.exists(path):
            with open(path) as f:
                d = json.load(f)
            return int(d["dram_bytes_read"]) + int(d["dram_bytes_write"])
    return None


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def frame_payload(rng):
    """PAYLOAD bytes: one varint-delimited ToReceiver{AudioData{opus_encoded_frame}} (protocol/ip.proto:29-33, 62-64) whose
    Opus bytes are a code-0 CELT fullband stereo 20 ms packet (TOC 0xFC) with random contents -- what the reference's
    receive loop would hand to pb_decode_delimited and then to opus_decode."""
    n = PAYLOAD - 5
    opus = bytes([0xFC]) + rng.integers(0, 256, size=n - 1, dtype=np.uint8).tobytes()
    return bytes([n + 4, 0x0A, n + 2, 0x0A, n]) + opus


def build_programs(cfg, frame_symbols, tx_params, n_ch, ch0, max_len=512):
    """Per-channel cyclic symbol programs: frames of PAYLOAD bytes (frame_payload; seed = global channel id) separated by
    4..16 symbols of silence.  `frame_symbols` / `tx_params` come from the product (our arm) or from the oracle's own
    transmit side (CPU arm): tests/test_oracle_tx.py holds the two against each other."""
    progs = np.full((n_ch, max_len), 0xFF, dtype=np.uint8)
    lens = np.zeros(n_ch, dtype=np.int32)
    params = []
    for c in range(n_ch):
        rng = np.random.default_rng(ch0 + c)
        parts, total = [], 0
        while True:
            pl = frame_payload(rng)
            syms = frame_symbols(cfg, pl)
            gap = int(rng.integers(4, 17))
            if total + len(syms) + gap > max_len:
                break
            parts += [syms, np.full(gap, 0xFF, dtype=np.uint8)]
            total += len(syms) + gap
        p = np.concatenate(parts)
        progs[c, : len(p)] = p
        lens[c] = len(p)
        params.append(tx_params(seed=ch0 + c, start_offset=-int(rng.integers(0, 4 * cfg.sym_len)), amplitude=0.5, snr_db=SNR_DB, ppm=0.0))
    return progs, lens, params


class NvmlSampler:
    """SM clock, power and throttle reasons polled through NVML every ~5 ms DURING the timed regions.  NVML queries take
    driver locks that kernel launches also need, so polling much faster than this can starve the launch loop."""

    def __init__(self, gpu_index):
        self.gpu, self.rows, self.stop_flag, self.ok = gpu_index, [], False, False
        try:
            import pynvml

            self.nv = pynvml
            pynvml.nvmlInit()
            import torch

            uuid = str(torch.cuda.get_device_properties(gpu_index).uuid)   # honour CUDA_VISIBLE_DEVICES: map by UUID
            self.h = None
            for i in range(pynvml.nvmlDeviceGetCount()):
                h = pynvml.nvmlDeviceGetHandleByIndex(i)
                u = pynvml.nvmlDeviceGetUUID(h)
                u = u.decode() if isinstance(u, bytes) else u
                if uuid in u:
                    self.h = h
            if self.h is None:
                self.h = pynvml.nvmlDeviceGetHandleByIndex(gpu_index)
            self.max = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception as e:  # pragma: no cover
            self.err = repr(e)

    def start(self):
        if self.ok:
            self.t = threading.Thread(target=self._run, daemon=True)
            self.t.start()

    def _run(self):
        nv = self.nv
        while not self.stop_flag:
            try:
                c = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
                r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                try:
                    w = nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0
                except Exception:
                    w = float("nan")
                self.rows.append((time.perf_counter(), c, r, w))
            except Exception:
                break
            time.sleep(0.005)

    def stop(self):
        self.stop_flag = True
        if self.ok:
            self.t.join(timeout=1)

    def window(self, t0, t1):
        if not self.ok:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvml unavailable: " + getattr(self, "err", "")]}
        rows = [(c, r, w) for (t, c, r, w) in self.rows if t0 <= t <= t1]
        if not rows:   # region shorter than one poll: take the samples closest to it
            rows = [(c, r, w) for (t, c, r, w) in self.rows[-3:]]
        nv = self.nv
        names = {"hw_slowdown": nv.nvmlClocksEventReasonHwSlowdown, "hw_thermal_slowdown": nv.nvmlClocksEventReasonHwThermalSlowdown,
                 "sw_thermal_slowdown": nv.nvmlClocksEventReasonSwThermalSlowdown, "sw_power_cap": nv.nvmlClocksEventReasonSwPowerCap}
        reasons = sorted(k for k, bit in names.items() if any(r & bit for _, r, _ in rows))
        sm = [c for c, _, _ in rows]
        pw = [w for _, _, w in rows if w == w]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_mhz_min": float(min(sm)) if sm else None, "sm_max_mhz": float(self.max),
                "power_w_median": round(float(np.median(pw)), 1) if pw else None, "power_w_max": round(float(max(pw)), 1) if pw else None,
                "samples": len(sm), "seconds": round(t1 - t0, 3), "reasons": reasons}


def bind_to_gpu_numa(torch, local):
    """Pins this process (and so the first touch of its pinned PCM buffers) to the CPUs of the GPU's NUMA node."""
    try:
        bus = torch.cuda.get_device_properties(local).pci_bus_id
        dom = torch.cuda.get_device_properties(local).pci_domain_id
        dev = torch.cuda.get_device_properties(local).pci_device_id
        path = "/sys/bus/pci/devices/%04x:%02x:%02x.0/numa_node" % (dom, bus, dev)
        node = int(open(path).read().strip())
        if node < 0:
            return {"node": node, "bound": False}
        cpus = set()
        for part in open("/sys/devices/system/node/node%d/cpulist" % node).read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
        return {"node": node, "bound": bool(cpus), "cpus": len(cpus)}
    except Exception as e:
        return {"node": None, "bound": False, "why": repr(e)[:80]}


# --------------------------------------------------------------------------------------------- CPU arm
def reference_line(preset="ref4", n_ch=CH_PER_GPU, steps=20, warmup=3, threads=None, world=1):
    """The in-repo C oracle with all host threads on bounded samples of the workload: a CPU step is CPU_CHUNKS consecutive
    chunks of all n_ch channels (our step is CHUNKS_PER_STEP of them).  Uses liboracle.so only: presets, frames, transmitter
    and demodulator are the oracle's own (oracle/anm_oracle_tx.c, anm_oracle.c); libanmodem.so is never mapped."""
    import oracle_binding as ob
    import audio_network_b200 as anm_types   # record types only; loads nothing

    cfg = ob.preset(preset)
    threads = threads or (os.cpu_count() or 1)
    chunk = CHUNK_SYMS * cfg.sym_len
    n = CPU_CHUNKS * chunk
    progs, lens, plist = build_programs(cfg, ob.frame_symbols, ob.tx_params, n_ch, 0)
    arr = np.zeros(n_ch, dtype=anm_types.TXPARAMS_DTYPE)
    for i, p in enumerate(plist):
        arr[i] = (p.seed, p.start_offset, p.amplitude_q15, p.snr_mdb, p.ppm_x1000, 0)
    pcm = ob.tx_render_batch(cfg, progs, lens, arr, 0, n, threads)
    times, ok_total, by_total = [], 0, 0
    for s in range(warmup + steps):
        sec, ok, bad, by, _ = ob.run_batch(cfg, pcm, threads)
        if s >= warmup:
            times.append(sec)
            ok_total += ok
            by_total += by
    tot = sum(times)
    msps = n_ch * n * len(times) / tot / 1e6
    sample = ("each step = %d of the %d chunks of a step x all %d channels of one GPU's shard, streamed through the in-repo C oracle "
              "(gcc -O2 -mfma, one channel at a time per thread; no reference demodulator exists, SURVEY.md 0)" % (CPU_CHUNKS, CHUNKS_PER_STEP, n_ch))
    return {
        "impl": "reference", "metric": METRIC, "value": round(msps, 3), "unit": "Msamples/s", "n_gpus": world,
        "steps": steps, "warmup": warmup, "ms_per_step": round(1e3 * tot / len(times), 3),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(preset, n_ch, world),
        "decoded_bits_per_s": round(by_total * 8 / tot, 1), "frames_ok": int(ok_total),
        "cpu_baseline": {"value": round(msps, 3), "unit": "Msamples/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": round(msps, 3), "unit": "Msamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "one host's cores whatever --gpus says: at N > 1 this is still the throughput of ONE host, not of N",
    }


def reference_arm(args):
    if int(os.environ.get("RANK", "0")) != 0:
        return
    print(json.dumps(reference_line(args.preset, args.channels, args.steps, args.warmup, None, args.gpus)), flush=True)


def cpu_baseline(cfg_name, pcm, n_threads):
    """The in-repo C oracle on the SAME PCM the GPU arm demodulates (its first chunk, all channels), repeated to ~10 s."""
    import oracle_binding as ob

    cfg = ob.preset(cfg_name)
    sec, ok, bad, by, _ = ob.run_batch(cfg, pcm, n_threads)      # calibration pass (also warms caches)
    reps = int(max(1, min(200, 10.0 / max(sec, 1e-3))))
    sec = 0.0
    for _ in range(reps):
        s1, ok, bad, by, _ = ob.run_batch(cfg, pcm, n_threads)
        sec += s1
    sec /= reps
    msps = pcm.shape[0] * pcm.shape[1] / sec / 1e6
    return {
        "value": round(msps, 3), "unit": "Msamples/s", "cores": n_threads, "kind": "port",
        "sample": "%d channels x %d samples (chunk 0 of the step, all channels of the shard, the same PCM), %d repeats; in-repo C oracle, "
                  "gcc -O2 -mfma, one channel at a time per thread (no reference demodulator exists, SURVEY.md 0)" % (pcm.shape[0], pcm.shape[1], reps),
        "seconds": round(sec * reps, 3), "repeats": reps, "frames_ok": int(ok), "decoded_bits_per_s": round(by * 8 / sec, 1),
    }


# --------------------------------------------------------------------------------------------- GPU legs
def kernel_name(cfg, anm):
    return ("k_demod_tc<%d,%d,%d>" if anm.config_dense(cfg) else "k_demod<%d,%d,%d>") % (cfg.n_tones, cfg.sym_len, cfg.hops_per_sym)


def tensor_profile():
    """Tensor-pipe utilisation of k_demod_tc from the committed ncu --set full capture (profiles/)."""
    for name in ("r2_tc_pipe.json", "r1_tc_pipe.json"):
        path = os.path.join(ROOT, "profiles", name)
        if os.path.exists(path):
            with open(path) as f:
                return json.load(f)
    return None


def decode_chain_leg(anm, torch, dev, recs, by, reps=20):
    """The steps after frame assembly (rows f2 / f1 first stage) on ONE drain of the timed region (one step's frames):
    k_pb_deframe locates the Opus bytes of every frame's ToReceiver message, k_opus_parse reads the packets' framing."""
    n = len(recs)
    if n == 0:
        return None
    d_recs = torch.from_numpy(recs.view(np.uint8).reshape(-1).copy()).to(dev)
    d_by = torch.from_numpy(np.ascontiguousarray(by)).to(dev)
    d_spans = torch.zeros(n * anm.PB_SPAN_DTYPE.itemsize, dtype=torch.uint8, device=dev)
    d_pk = torch.zeros(n * anm.OPUS_PACKET_DTYPE.itemsize, dtype=torch.uint8, device=dev)
    L = anm.lib()
    stream = torch.cuda.current_stream().cuda_stream
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    t_def = t_par = 0.0
    for r in range(reps + 2):
        ev[0].record()
        assert L.anm_pb_deframe_device(d_recs.data_ptr(), n, d_by.data_ptr(), 0xFFFFFFFF, d_spans.data_ptr(), stream) == 0
        ev[1].record()
        assert L.anm_opus_parse_device(d_spans.data_ptr(), n, d_by.data_ptr(), 0xFFFFFFFF, 48000, d_pk.data_ptr(), stream) == 0
        ev[2].record()
        torch.cuda.synchronize()
        if r >= 2:
            t_def += ev[0].elapsed_time(ev[1])
            t_par += ev[1].elapsed_time(ev[2])
    spans = d_spans.cpu().numpy().view(anm.PB_SPAN_DTYPE)
    pk = d_pk.cpu().numpy().view(anm.OPUS_PACKET_DTYPE)
    ok = recs["crc_ok"] == 1
    # row f1 stage 1: the CELT entropy decode of every located packet, one GPU thread per channel (= stream), frames in stream order
    celt = None
    sel = np.flatnonzero((pk["count"] == 1) & (pk["mode"] == 1002))
    if len(sel):
        order = sel[np.lexsort((recs["start_sample"][sel], recs["channel"][sel]))]
        jobs, _ = anm.celt_jobs_from_packets(spans[order], pk[order])   # one 20 ms fullband frame per packet in this workload
        assert len(jobs) == len(order)
        chs, first = np.unique(recs["channel"][order], return_index=True)
        sb = np.concatenate([first, [len(order)]]).astype(np.uint32)
        ctx = ctypes.c_void_p()
        assert L.anm_celt_ctx_create(dev.index or 0, ctypes.byref(ctx)) == 0
        d_jobs = torch.from_numpy(jobs.view(np.uint8).reshape(-1).copy()).to(dev)
        d_sb = torch.from_numpy(sb.view(np.uint8).copy()).to(dev)
        d_st = torch.zeros(len(chs) * anm.CELT_STREAM_DTYPE.itemsize, dtype=torch.uint8, device=dev)
        d_fr = torch.zeros(len(jobs) * anm.CELT_FRAME_DTYPE.itemsize, dtype=torch.uint8, device=dev)
        t_celt = 0.0
        for r in range(reps // 2 + 2):
            d_st.zero_()
            ev[0].record()
            assert L.anm_celt_entropy_device(ctx, d_jobs.data_ptr(), d_sb.data_ptr(), len(chs), len(jobs), d_by.data_ptr(), 0xFFFFFFFF, d_st.data_ptr(), d_fr.data_ptr(), stream) == 0
            ev[1].record()
            torch.cuda.synchronize()
            if r >= 2:
                t_celt += ev[0].elapsed_time(ev[1])
        # rows f1 stages 2 + 3 on a bounded sample of whole streams (15 KB of intermediate data per frame): spectrum, inverse MDCT, overlap / post-filter / de-emphasis
        ns_dec = int(np.searchsorted(sb, 65536, side="right")) - 1 if len(jobs) > 65536 else len(chs)
        ns_dec = max(ns_dec, 1)
        nj_dec = int(sb[ns_dec])
        d_sy = torch.zeros(ns_dec * anm.CELT_SYNTH_DTYPE.itemsize, dtype=torch.uint8, device=dev)
        d_pcm = torch.zeros(nj_dec * 1920, dtype=torch.int16, device=dev)
        t_dec = 0.0
        for r in range(3):
            d_st.zero_()
            d_sy.zero_()
            ev[0].record()
            assert L.anm_celt_decode_device(ctx, d_jobs.data_ptr(), d_sb.data_ptr(), ns_dec, nj_dec, d_by.data_ptr(), 0xFFFFFFFF, d_st.data_ptr(), d_sy.data_ptr(),
                                            d_fr.data_ptr(), d_pcm.data_ptr(), 1920, stream) == 0
            ev[1].record()
            torch.cuda.synchronize()
            if r >= 1:
                t_dec += ev[0].elapsed_time(ev[1]) / 2
        pcm_nonzero = bool((d_pcm != 0).any().item())
        del d_pcm, d_sy
        d_st.zero_()
        assert L.anm_celt_entropy_device(ctx, d_jobs.data_ptr(), d_sb.data_ptr(), len(chs), len(jobs), d_by.data_ptr(), 0xFFFFFFFF, d_st.data_ptr(), d_fr.data_ptr(), stream) == 0
        torch.cuda.synchronize()
        L.anm_celt_ctx_destroy(ctx)
        fr = d_fr.cpu().numpy().view(anm.CELT_FRAME_DTYPE)
        ms = t_celt / (reps // 2)
        celt = {"k_celt_entropy_plus_energies_ms": round(ms, 4), "streams": int(len(chs)), "frames": int(len(jobs)), "Mframes_per_s": round(len(jobs) / (ms * 1e-3) / 1e6, 2),
                "packet_bytes": int(jobs["len"].sum()), "frames_within_budget": int(((fr["flags"] & 1024) == 0).sum()),
                "full_decode": {"kernels": "k_celt_entropy, k_celt_energies, k_celt_spectrum, k_celt_blocks, k_celt_overlap, k_celt_deemphasis", "streams": ns_dec, "frames": nj_dec,
                                "ms": round(t_dec, 3), "Mframes_per_s": round(nj_dec / (t_dec * 1e-3) / 1e6, 3), "audio_x_realtime": round(nj_dec * 0.02 / (t_dec * 1e-3), 1),
                                "pcm_nonzero": pcm_nonzero,
                                "note": "bounded sample of whole streams; PCM parity of these kernels is pinned on real encoder output in tests/test_celt_synth.py, "
                                        "throughput on real frames: tools/celt_bench.py"},
                "note": "payloads of the synthetic workload are random bytes behind a CELT TOC: valid range-coder input, worst case for branch divergence; "
                        "parity of this stage is pinned on real encoder output in tests/test_celt_entropy.py"}
    peak, _src = peaks()
    # algorithmic bytes per frame: k_pb_deframe = 24 (frame record) + 32 (the sector holding the message header) + 16 (span out);
    # k_opus_parse = 16 (span) + 32 (the sector holding the TOC / size bytes) + 128 (packet record out)
    gbs_def = n * 72 / (t_def / reps * 1e-3) / 1e9
    gbs_par = n * 176 / (t_par / reps * 1e-3) / 1e9
    return {"batch_frames": int(n), "payload_bytes": int(len(by)), "k_pb_deframe_ms": round(t_def / reps, 4), "k_opus_parse_ms": round(t_par / reps, 4),
            "k_pb_deframe_hbm": {"algorithmic_bytes_per_frame": 72, "achieved_gbs": round(gbs_def, 1), "frac": round(gbs_def / peak, 4)},
            "k_opus_parse_hbm": {"algorithmic_bytes_per_frame": 176, "achieved_gbs": round(gbs_par, 1), "frac": round(gbs_par / peak, 4)},
            "Mframes_per_s": round(n / ((t_def + t_par) / reps * 1e-3) / 1e6, 1), "gpu_launches": 2 * reps, "celt_entropy": celt,
            "audio_located": int((spans["status"] == anm.ANM_PB_OK).sum()), "crc_ok_frames": int(ok.sum()),
            "opus_packets_parsed": int((pk["count"] == 1).sum()),
            "all_crc_ok_frames_decode": bool(((spans["status"] == anm.ANM_PB_OK) == ok).all() and ((pk["count"] == 1) == ok).all()
                                              and (pk["size"][ok, 0] == PAYLOAD - 6).all())}


def params_array(anm, plist):
    return anm.tx_params_array(plist)


def render_resident(anm, torch, cfg, dev, n_ch, ch0, chunks, chunk, stream):
    progs, lens, plist = build_programs(cfg, anm.frame_symbols, anm.tx_params, n_ch, ch0)
    params = params_array(anm, plist)
    d_prog = torch.from_numpy(progs).to(dev)
    d_len = torch.from_numpy(lens).to(dev)
    d_par = torch.from_numpy(params.view(np.uint8).copy()).to(dev)
    total = chunks * chunk
    d_pcm = torch.empty((n_ch, total), dtype=torch.int16, device=dev)
    anm.tx_render_device(cfg, d_prog.data_ptr(), progs.shape[1], d_len.data_ptr(), d_par.data_ptr(), n_ch, 0,
                         d_pcm.data_ptr(), total, total, stream)
    torch.cuda.synchronize()
    return d_pcm, total


def cfg4_leg(anm, torch, dev, local, stream, steps=12, warmup=3, n_ch=4736):
    """BASELINE config 4: the 64-tone preset through the tcgen05 contraction kernel (k_demod_tc), same chunking and SNR as
    the headline workload; PCM resident in HBM, inputs larger than L2.
    4,736 channels = 148 SMs x 2 resident CTAs x 4 channels per CTA x 4 waves."""
    cfg = anm.config_preset("wide64")
    chunk = CHUNK_SYMS * cfg.sym_len
    resident = min(steps + warmup, 8)
    d_pcm, total = render_resident(anm, torch, cfg, dev, n_ch, 1 << 20, resident, chunk, stream)
    dm = anm.Demod(cfg, n_ch, device=local)
    for i in range(warmup):
        dm.feed_device(d_pcm.data_ptr() + (i % resident) * chunk * 2, total, chunk, stream)
    dm.collect()
    dm.read_frames(cap=1 << 22, bytes_cap=1 << 28)
    dm.kernel_time()
    for i in range(steps):
        dm.feed_device(d_pcm.data_ptr() + ((warmup + i) % resident) * chunk * 2, total, chunk, stream)
    torch.cuda.synchronize()
    k_ms, k_n = dm.kernel_time()
    dm.collect()
    recs, _ = dm.read_frames(cap=1 << 22, bytes_cap=1 << 30)
    avg_ms = k_ms / max(1, k_n)
    peak, _ = peaks()
    gbs = n_ch * chunk * 2 / (avg_ms * 1e-3) / 1e9
    out = {"workload": "cfg4: %d channels x %d-sample chunks, preset wide64 (64 tones, N=256), 10 dB SNR" % (n_ch, chunk),
           "kernel": kernel_name(cfg, anm), "value": round(n_ch * chunk / (avg_ms * 1e-3) / 1e6, 2), "unit": "Msamples/s",
           "avg_kernel_ms": round(avg_ms, 4), "launches_timed": k_n, "hbm_gbs": round(gbs, 1), "hbm_frac": round(gbs / peak, 4),
           "frames_ok": int((recs["crc_ok"] == 1).sum()), "arith": "s8/u8 x s8 -> s32 (tcgen05.mma kind::i8), exact",
           "parity": "tests/test_dense.py::test_dense_config4_at_bench_scale_all_channels_vs_oracle checks this size against the oracle",
           "ncu": tensor_profile()}
    dm.close()
    del d_pcm
    return out


class FrameTotals:
    """What the consumer of a drain keeps in this bench: counts (frames, CRC-valid frames, payload bits)."""

    def __init__(self, anm):
        self.anm = anm
        self.frames = self.ok = self.bits = self.d2h = 0

    def add(self, recs, nbytes):
        ok, by = self.anm.frames_summary(recs)
        self.frames += len(recs)
        self.ok += ok
        self.bits += by * 8
        self.d2h += 16 + recs.nbytes + nbytes


def main():
    args = parse()
    if args.impl == "reference":
        reference_arm(args)
        return
    import torch
    import torch.distributed as dist

    import audio_network_b200 as anm

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback exists)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    numa = bind_to_gpu_numa(torch, local)
    host_group = None
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)       # barrier + timing reductions only
        host_group = dist.new_group(backend="gloo")          # the host-side gather of decoded frames

    cfg = anm.config_preset(args.preset)
    N = cfg.sym_len
    n_ch = args.channels
    chunk = CHUNK_SYMS * N
    CPS = CHUNKS_PER_STEP
    assert 1 <= CPS <= 62, "a step's launches must fit the handle's snapshot window (collect_upto lag < 63)"
    LC = max(1, min(CPS, args.launch_chunks if args.launch_chunks > 0 else CPS))
    if anm.config_dense(cfg):
        LC = 1                                              # the tensor-core kernel takes its chunks one launch at a time
    stream = torch.cuda.current_stream().cuda_stream

    # ---- synthesize this rank's shard in HBM (not timed): all CPS chunks of the step are distinct and resident ----
    d_pcm, total = render_resident(anm, torch, cfg, dev, n_ch, rank * n_ch, CPS, chunk, stream)
    dm = anm.Demod(cfg, n_ch, device=local)
    L = anm.lib()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def run_steps(n_steps, totals):
        """n_steps x CPS launches; the frames of step s are drained (device -> the handle's pinned host queue) and consumed in
        place while the launches of step s+1 are already queued, so the GPU never waits for the host.  The last drain is left
        in the queue for the caller (peek_frames)."""
        for s in range(n_steps):
            l_step = dm.launch_count()
            for c in range(0, CPS, LC):                     # LC resident chunks per launch: work items (chunk, channel) from one queue
                dm.feed_device_chunks(d_pcm.data_ptr() + c * chunk * 2, total, chunk, chunk, min(LC, CPS - c), stream)
            if s > 0:
                dm.collect_upto(dm.launch_count() - l_step)  # everything up to the last launch of step s-1
                recs, by = dm.peek_frames()
                totals.add(recs, len(by))
                dm.drop_frames()
        dm.collect()
        recs, by = dm.peek_frames()
        totals.add(recs, len(by))
        assert not dm.overflowed(), "frame queue overflowed inside the timed region"

    sampler = NvmlSampler(local)
    sampler.start()
    run_steps(args.warmup, FrameTotals(anm))
    dm.drop_frames()
    dm.kernel_time()
    l0 = dm.launch_count()
    barrier()
    tm0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    tot = FrameTotals(anm)
    run_steps(args.steps, tot)
    e1.record()
    barrier()
    tm1 = time.perf_counter()
    kept = tuple(a.copy() for a in dm.peek_frames())        # the last step's frames, for the decode-chain leg (outside the timed region)
    dm.drop_frames()
    ms = e0.elapsed_time(e1)
    launches = dm.launch_count() - l0
    k_ms, k_n = dm.kernel_time()                             # the first 64 launches of the region, one event pair each
    clocks = sampler.window(tm0, tm1)

    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    agg = torch.tensor([float(tot.bits), float(tot.ok), float(launches), float(tot.frames)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(agg, op=dist.ReduceOp.SUM)           # totals only; the frames themselves travel in the e2e leg
    ms_max = float(t.item())
    samples_step = n_ch * chunk * CPS
    value = float(world) * samples_step * args.steps / (ms_max * 1e-3) / 1e6

    # ---- sustained region: the same loop for >= SUSTAIN_S seconds, clocks and power sampled throughout ----
    sustained = None
    if not args.no_sustain:
        n_sus = max(1, int(np.ceil(SUSTAIN_S * 1e3 / (ms / args.steps))))
        barrier()
        ts0 = time.perf_counter()
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record()
        l_sus = dm.launch_count()
        run_steps(n_sus, FrameTotals(anm))
        l_sus = dm.launch_count() - l_sus
        s1.record()
        barrier()
        ts1 = time.perf_counter()
        dm.drop_frames()
        sms = s0.elapsed_time(s1)
        tt = torch.tensor([sms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        sus_clk = sampler.window(ts0, ts1)
        peak, _ = peaks()
        sustained = {"value": round(float(world) * samples_step * n_sus / (float(tt.item()) * 1e-3) / 1e6, 2), "unit": "Msamples/s",
                     "steps": n_sus, "launches": int(l_sus), "chunks": n_sus * CPS, "seconds": round(float(tt.item()) * 1e-3, 3),
                     "hbm_frac_of_step_time": round(samples_step * 2 * n_sus / (float(tt.item()) * 1e-3) / 1e9 / peak, 4),
                     "clocks": sus_clk}
    sampler.stop()

    # ---- e2e: pinned host PCM -> feed_host_async (H2D inside) -> collect + take frames (D2H) -> gather on rank 0 ----
    e2e = None
    if args.e2e_steps > 0:
        nh = 4
        host = [torch.empty((n_ch, chunk), dtype=torch.int16).pin_memory() for _ in range(nh)]
        for j in range(nh):
            host[j].copy_(d_pcm[:, j * chunk: (j + 1) * chunk])
        torch.cuda.synchronize()
        # the box's ceiling for this leg: the same buffers as plain pinned copies on every rank at once
        d_tmp = torch.empty((n_ch, chunk), dtype=torch.int16, device=dev)
        for j in range(2):
            d_tmp.copy_(host[j], non_blocking=True)
        barrier()
        c0 = time.perf_counter()
        ncp = 24
        for j in range(ncp):
            d_tmp.copy_(host[j % nh], non_blocking=True)
        torch.cuda.synchronize()
        tc = torch.tensor([time.perf_counter() - c0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tc, op=dist.ReduceOp.MAX)
        ceiling_gbs = float(world) * ncp * n_ch * chunk * 2 / float(tc.item()) / 1e9
        del d_tmp

        dm2 = anm.Demod(cfg, n_ch, device=local)
        dm2.feed_host_ptr(host[0].data_ptr(), chunk, chunk)   # warm-up (allocates the staging buffers)
        dm2.collect()
        dm2.read_frames(cap=1 << 20, bytes_cap=1 << 26)
        dm2.reset()
        et = FrameTotals(anm)
        gathered = {"frames": 0, "bytes": 0, "digest_ok": None}
        step_recs, step_by = [], []

        def consume():
            recs, by = dm2.peek_frames()
            if len(recs):
                et.add(recs, len(by))
                step_recs.append(recs.copy())                # out of the handle's queue: these travel to rank 0
                step_by.append(by.copy())
            dm2.drop_frames()

        def gather_step():
            """host-side gather of the step's frame records on rank 0: binary records over gloo (no pickles, no NCCL)"""
            if step_recs:
                recs = np.concatenate(step_recs)
                by = np.concatenate(step_by)
                offs = np.concatenate([[0], np.cumsum([len(b) for b in step_by])[:-1]]).astype(np.int64)
                base = np.repeat(offs, [len(r) for r in step_recs])
                recs["offset"] = (recs["offset"].astype(np.int64) + base).astype(np.uint32)   # offsets into the step's byte arena
                recs["channel"] += rank * n_ch                                                # global channel ids
            else:
                recs, by = np.zeros(0, dtype=anm.FRAME_DTYPE), np.zeros(0, dtype=np.uint8)
            step_recs.clear()
            step_by.clear()
            local_dg = anm.frames_digest(recs, by)
            if world == 1:
                gathered["frames"] += len(recs)
                gathered["bytes"] += len(by)
                return recs, by, [local_dg]
            sizes = torch.tensor([recs.nbytes, len(by), local_dg & 0x7FFFFFFFFFFFFFFF, local_dg >> 63], dtype=torch.int64)
            all_sizes = [torch.zeros(4, dtype=torch.int64) for _ in range(world)]
            dist.all_gather(all_sizes, sizes, group=host_group)
            if rank == 0:
                parts_r, parts_b = [recs], [by]
                for r in range(1, world):
                    tr = torch.empty(int(all_sizes[r][0]), dtype=torch.uint8)
                    tb = torch.empty(int(all_sizes[r][1]), dtype=torch.uint8)
                    dist.recv(tr, src=r, group=host_group)
                    dist.recv(tb, src=r, group=host_group)
                    parts_r.append(tr.numpy().view(anm.FRAME_DTYPE))
                    parts_b.append(tb.numpy())
                gathered["frames"] += sum(len(p) for p in parts_r)
                gathered["bytes"] += sum(len(p) for p in parts_b)
                dgs = [int(a[2]) | (int(a[3]) << 63) for a in all_sizes]
                return parts_r, parts_b, dgs
            dist.send(torch.from_numpy(recs.view(np.uint8).reshape(-1)), dst=0, group=host_group)
            dist.send(torch.from_numpy(by), dst=0, group=host_group)
            return None, None, None

        barrier()
        t0 = time.perf_counter()
        last = None
        # pipelined: while chunk j crosses PCIe, the host drains and takes the frames of chunk j-1
        for s in range(args.e2e_steps):
            for j in range(CPS):
                dm2.feed_host_async_ptr(host[j % nh].data_ptr(), chunk, chunk)
                dm2.collect_upto(1)
                consume()
            if s == args.e2e_steps - 1:
                dm2.collect()
                consume()
            last = gather_step()
        torch.cuda.synchronize()
        te = time.perf_counter() - t0
        tt = torch.tensor([te], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        assert not dm2.overflowed()
        if rank == 0 and last is not None and last[0] is not None:
            # outside the timed region: the records that arrived on rank 0 carry the digests the ranks computed locally
            if world == 1:
                gathered["digest_ok"] = bool(anm.frames_digest(last[0], last[1]) == last[2][0])
            else:
                gathered["digest_ok"] = all(anm.frames_digest(pr, pb) == dg for pr, pb, dg in zip(last[0], last[1], last[2]))
        e2e_val = float(world) * samples_step * args.e2e_steps / float(tt.item()) / 1e6
        e2e = {"value": round(e2e_val, 2), "unit": "Msamples/s",
               "h2d_bytes_per_step": n_ch * chunk * 2 * CPS, "d2h_bytes_per_step": int(et.d2h // args.e2e_steps), "steps": args.e2e_steps,
               "seconds": round(float(tt.item()), 3),
               "h2d_ceiling": {"GB_per_s_all_ranks": round(ceiling_gbs, 1), "how": "%d plain cudaMemcpyAsync of the same pinned chunks on every rank at once" % ncp},
               "frac_of_h2d_ceiling": round(e2e_val * 2e6 / 1e9 / ceiling_gbs, 4),
               "gather": {"frames_on_rank0": gathered["frames"], "payload_bytes_on_rank0": gathered["bytes"], "digest_ok": gathered["digest_ok"],
                          "transport": "torch.distributed gloo send/recv of raw anm_frame_t records + payload arena" if world > 1 else "in-process"},
               "numa": numa,
               "note": "pinned host PCM -> H2D copy -> kernel -> frames D2H -> host-side gather on rank 0, every step, through anm_demod_feed_host_async / "
                       "collect_upto / take_frames (two staging buffers: the copy of chunk k+1 overlaps the kernel of chunk k)"}
        dm2.close()
        del host

    # ---- CPU baseline (rank 0, N=1 only): oracle on the same PCM, same config ----
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        sample = np.ascontiguousarray(d_pcm[:, :chunk].cpu().numpy())
        cpu = cpu_baseline(args.preset, sample, os.cpu_count() or 1)

    chain = decode_chain_leg(anm, torch, dev, kept[0], kept[1]) if (rank == 0 and kept is not None) else None

    cfg4 = None
    if rank == 0 and world == 1 and not args.no_cfg4 and args.preset == "ref4":
        del d_pcm
        torch.cuda.empty_cache()
        cfg4 = cfg4_leg(anm, torch, dev, local, stream)

    if rank == 0:
        peak, peak_src = peaks()
        per_launch_bytes = n_ch * chunk * 2 * LC
        avg_ms = (k_ms / k_n) if k_n else ms / max(1, launches)
        achieved = per_launch_bytes / (avg_ms * 1e-3) / 1e9
        grid, wpc, smem = dm.launch_geometry()
        config = workload_config(args.preset, n_ch, world)
        line = {
            "metric": METRIC, "value": round(value, 2), "unit": "Msamples/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": round(ms_max / args.steps, 4), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": config,
            "decoded_bits_per_s": round(float(agg[0].item()) / (ms_max * 1e-3), 1),
            "frames_ok": int(agg[1].item()), "frames": int(agg[3].item()),
            "gpu_launches": int(agg[2].item()),
            "launch": {"grid": grid, "warps_per_cta": wpc, "smem_bytes": smem},
            "roofline": {"bound": "hbm", "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s",
                         "frac": round(achieved / peak, 4), "traffic": (measured_traffic() * LC) if (n_ch == CH_PER_GPU and args.preset == "ref4" and measured_traffic()) else None, "peak_source": peak_src,
                         "kernel": kernel_name(cfg, anm),
                         "avg_kernel_ms": round(avg_ms, 4), "launches_timed": k_n,
                         "algorithmic_bytes_per_launch": per_launch_bytes,
                         "chunks_per_launch": LC,
                         "frac_of_step_time": round(n_ch * chunk * 2 * CPS / (ms / args.steps * 1e-3) / 1e9 / peak, 4),
                         "note": "achieved = bytes / per-launch CUDA-event time of the first %d launches of the timed region (a launch covers chunks_per_launch "
                                 "resident chunks; traffic = the ncu DRAM bytes of a one-chunk launch x chunks_per_launch); frac_of_step_time divides by the whole "
                                 "step instead (launch gaps and the per-step drain included)" % k_n},
            "clocks": clocks,
        }
        if sustained:
            line["sustained"] = sustained
        if e2e:
            line["e2e"] = e2e
        if cpu:
            line["cpu_baseline"] = cpu
        if chain:
            line["decode_chain"] = chain
        if cfg4:
            line["cfg4"] = cfg4
        print(json.dumps(line), flush=True)
    dm.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
