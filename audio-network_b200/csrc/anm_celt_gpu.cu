/*
 * anm_celt_gpu.cu -- the batched CELT decoder on the GPU (include/anmodem_opus.h, anm_celt_entropy_* / _spectrum_* / _decode_*).
 * Stage 1, two passes.  k_celt_entropy, one thread per FRAME: the range decoder is sequential inside a frame, but no symbol of a
 * frame depends on any other frame, so every frame of every stream decodes at once (integer / byte work on a few hundred bytes of
 * packet, latency bound per thread: the batch of frames is what fills the machine).  k_celt_energies, one warp per STREAM: the band
 * energies predict from the previous frame of the same stream (celt/quant_bands.c:427-490) -- a recurrence over the stream's frames in order
 * on the coarse symbols and offsets pass 1 left in a scratch array, band i on lane i (anm_celt_entropy.h holds the decode itself, shared
 * with the host-side test harness).
 *
 * Stage 2 (anm_celt_spectrum_*) adds k_celt_spectrum, one thread per frame again: the frame is picked up where pass 1 left its range decoder in
 * front of the band loop, and the bands are decoded with the spectrum arithmetic switched on (anm_celt_vec.h: PVQ vectors, rotations, folding,
 * reorderings, stereo merge, anti-collapse), seeded and informed by what the per-stream pass left (the noise seed is the previous frame's final range,
 * anti-collapse reads the stream's two log-energy histories).  The threads of a warp work on frames of their own, so what they do differs from band
 * to band; the band and partition walks are written so that the expensive steps sit at one place in the code each and the lanes meet there
 * (anm_celt_entropy.h: ce_partition, ce_band_channels), and the frames are grouped by kind first (k_celt_kind_*: transient or not, frame size,
 * channels, dual stereo, packet bytes) so that frames which do the same thing share a warp -- 14.4 of 32 lanes active per instruction against 3.3
 * for the straight transcription of the reference's recursion.  A warp-per-frame form was measured 3 x slower (anm_celt_vec.h).
 * Stage 3 (anm_celt_decode_*) adds k_celt_blocks (a warp per frame and output channel: denormalisation, the fixed-point FFT of the inverse MDCT in
 * shared memory), k_celt_overlap (a warp per stream and output channel: overlap-add and pitch post-filter with the output history in shared memory)
 * and k_celt_deemphasis (a thread per stream and output channel: the one-pole recurrence to 16-bit PCM).
 *
 * Reference path replaced: playback.cpp:115-122 opus_decode() -> opus_decode_frame (opus_decoder.c:214-626, CELT-only branch)
 * -> celt_decode_with_ec (celt/celt_decoder.c:815-1180), without the concealment of lost frames.
 */
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdlib>
#include <new>

#include "anm_celt_synth.h"
#include "anm_internal.h"

struct anm_celt_ctx {
    int device;
    anm_celt_tables_t *d_tables;
    int16_t *d_scratch; /* per frame: 42 coarse symbols + 42 energy offsets */
    size_t scratch_frames;
    ce_hist_t *d_hist;  /* per frame: the stream's histories before the frame (stage 2) */
    size_t hist_frames;
    ce_resume_t *d_resume; /* per frame: the range decoder in front of the band loop, where stage 2 picks the frame up */
    size_t resume_frames;
    uint32_t *d_order;  /* [2 n + kSortKeys]: the frames grouped by kind (stage 2's thread -> frame map) | each frame's place in its group | group sizes */
    size_t order_words;
    anm_celt_synth_tables_t *d_synth_tables; /* stage 3 */
    int16_t *d_x;       /* per frame: the normalised spectrum (anm_celt_decode_device keeps it to itself) */
    size_t x_frames;
    int32_t *d_raw;     /* per frame: the raw inverse-MDCT blocks of both output channels */
    size_t raw_frames;
};

namespace {

/* pass 1, one thread per FRAME: everything the frame's bits say.  No symbol depends on the stream's history, so all frames of all streams
 * decode at once; what the history needs (coarse symbols, energy offsets) goes to the scratch array. */
__global__ void __launch_bounds__(128) k_celt_entropy(const anm_celt_tables_t *__restrict__ t, const anm_celt_job_t *__restrict__ jobs, uint32_t n_jobs,
                                                      const uint8_t *__restrict__ bytes, uint32_t mask, int16_t *__restrict__ scratch, anm_celt_frame_t *out,
                                                      ce_resume_t *resume) {
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n_jobs) return;
    const anm_celt_job_t job = jobs[j];
    anm_celt_frame_t fr;
    int16_t qi[2 * ANM_CE_NB], eoff[2 * ANM_CE_NB];
    const int rc = anm_celt_frame_symbols(t, bytes, mask, job.offset, job.len, job.channels, job.lm, job.end_band, qi, eoff, &fr, 0, 0, 0, resume ? resume + j : 0);
    if (rc != 0) { /* impossible job description: treated like a lost frame, flagged */
        fr.final_range = 0;
        fr.flags = ANM_CELT_F_LOST | ANM_CELT_F_EC_ERROR;
    }
    int16_t *sc = scratch + (size_t)j * (4 * ANM_CE_NB);
    for (int i = 0; i < 2 * ANM_CE_NB; ++i) {
        sc[i] = qi[i];
        sc[2 * ANM_CE_NB + i] = eoff[i];
    }
    out[j] = fr;
}

/* pass 2, one WARP per STREAM: the band energies predict from the previous frame of the same stream and, inside a frame, from the bands below
 * (celt/quant_bands.c:427-490) -- a recurrence over the stream's frames in order.  Lane i holds band i of both channels; what runs up the bands
 * inside a frame (prev) is a sum of terms that each depend on their own band's symbol only, so it is a warp prefix sum.  The same arithmetic as
 * anm_celt_stream_step (anm_celt_entropy.h: the host-side harness runs that one; the GPU tests hold this kernel's output against the same golden
 * vectors); a thread per stream took 42 dependent steps and as many scattered loads per frame (1.1 ms for 4,096 streams of 50 frames). */
__global__ void __launch_bounds__(128) k_celt_energies(const uint32_t *__restrict__ stream_begin, uint32_t n_streams, const int16_t *__restrict__ scratch,
                                                       anm_celt_stream_t *streams, anm_celt_frame_t *out, ce_hist_t *hist) {
    constexpr unsigned kAll = 0xFFFFFFFFu;
    constexpr int NB = ANM_CE_NB;
    const uint32_t s = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (s >= n_streams) return; /* the whole warp */
    anm_celt_stream_t *st = &streams[s];
    const bool on = lane < NB;
    const int i = on ? lane : 0;
    uint32_t flags = st->flags, rng = st->rng;
    int oe[2], l1[2], l2[2];
    for (int c = 0; c < 2; ++c) {
        oe[c] = st->old_e[c * NB + i];
        l1[c] = st->log_e1[c * NB + i];
        l2[c] = st->log_e2[c * NB + i];
    }
    for (uint32_t j = stream_begin[s]; j < stream_begin[s + 1]; ++j) {
        const int16_t *qi = scratch + (size_t)j * (4 * NB), *eoff = qi + 2 * NB;
        anm_celt_frame_t *fr = &out[j];
        if (!(flags & 1u)) { /* a fresh decoder: both histories at -28 dB */
            l1[0] = l1[1] = l2[0] = l2[1] = -28672;
            flags |= 1u;
        }
        if (hist) {
            if (on)
                for (int c = 0; c < 2; ++c) {
                    hist[j].log_e1[c * NB + i] = (int16_t)l1[c];
                    hist[j].log_e2[c * NB + i] = (int16_t)l2[c];
                }
            if (lane == 0) hist[j].seed = rng;
        }
        const uint32_t ff = fr->flags;
        const bool lost = (ff & ANM_CELT_F_LOST) != 0;
        const int end = fr->pad[0];
        if (!lost) {
            const int C = fr->channels, LM = fr->lm, intra = (ff & ANM_CELT_F_INTRA) != 0;
            const int coef = intra ? 0 : (LM == 0 ? 29440 : LM == 1 ? 26112 : LM == 2 ? 21248 : 16384);
            const int beta = intra ? 4915 : (LM == 0 ? 30147 : LM == 1 ? 22282 : LM == 2 ? 12124 : 6554);
            if (C == 1) oe[0] = oe[0] > oe[1] ? oe[0] : oe[1];
            const bool act = on && i < end;
            for (int c = 0; c < C; ++c) {
                const int32_t q = act ? (int32_t)qi[i + c * NB] * 1024 : 0; /* SHL32(qi, DB_SHIFT) */
                const int32_t term = q * 128 - (int32_t)beta * (int16_t)ce_pshr32(q, 8);
                int32_t run = term; /* inclusive sum up the bands */
                for (int d = 1; d < 32; d <<= 1) {
                    const int32_t v = __shfl_up_sync(kAll, run, d);
                    if (lane >= d) run = (int32_t)((uint32_t)run + (uint32_t)v);
                }
                const int32_t prev = (int32_t)((uint32_t)run - (uint32_t)term);
                if (act) {
                    int e = oe[c];
                    if (e < -9216) e = -9216; /* MAX16(-QCONST16(9, DB_SHIFT), .) */
                    int32_t tmp = ce_pshr32((int32_t)coef * e, 8) + prev + q * 128;
                    if (tmp < -3670016) tmp = -3670016; /* -QCONST32(28, DB_SHIFT + 7) */
                    oe[c] = (int16_t)(ce_pshr32(tmp, 7) + eoff[i + c * NB]);
                }
            }
            if (ff & ANM_CELT_F_SILENCE)
                for (int c = 0; c < C; ++c) oe[c] = -28672; /* -QCONST16(28, DB_SHIFT) */
            if (C == 1) oe[1] = oe[0];
            if (i >= end) oe[0] = oe[1] = 0;
        }
        if (on) {
            fr->band_e[i] = (int16_t)oe[0];
            fr->band_e[NB + i] = (int16_t)oe[1];
        }
        if (lost) continue;
        for (int c = 0; c < 2; ++c) {
            if (!(ff & ANM_CELT_F_TRANSIENT)) {
                l2[c] = l1[c];
                l1[c] = oe[c];
            } else {
                l1[c] = l1[c] < oe[c] ? l1[c] : oe[c];
            }
            if (i >= end) l1[c] = l2[c] = -28672;
        }
        rng = fr->final_range;
    }
    if (on)
        for (int c = 0; c < 2; ++c) {
            st->old_e[c * NB + i] = (int16_t)oe[c];
            st->log_e1[c * NB + i] = (int16_t)l1[c];
            st->log_e2[c * NB + i] = (int16_t)l2[c];
        }
    if (lane == 0) {
        st->rng = rng;
        st->flags = flags;
    }
}

/* Which frames share a warp in stage 2.  The threads of a warp run together only where their frames do the same thing, and what a frame does is
 * largely settled by a few of its header symbols: transient frames (eight short blocks, with the Haar / Hadamard reorderings around every band) and
 * long ones, frame size, channel count, dual stereo, and roughly how many bits there are to spend.  So the frames are grouped by those (a counting sort
 * over kSortKeys kinds, three small kernels; the place of a frame inside its group is whatever order the atomics happen to give -- it only decides
 * which thread decodes the frame, not what comes out) and thread t of k_celt_spectrum takes frame order[t].  The expensive kinds come first. */
constexpr uint32_t kSortKeys = 1024;
__device__ __forceinline__ uint32_t celt_kind(const anm_celt_frame_t &fr, const anm_celt_job_t &job) {
    if (fr.flags & ANM_CELT_F_LOST) return kSortKeys - 1u;
    const uint32_t sz = 15u - min(15u, job.len / 96u);
    return ((fr.flags & ANM_CELT_F_TRANSIENT) ? 0u : 1u) << 8 | (3u - (fr.lm & 3u)) << 6 | (fr.channels == 2 ? 0u : 1u) << 5 |
           ((fr.flags & ANM_CELT_F_DUAL_STEREO) ? 0u : 1u) << 4 | sz;
}
__global__ void __launch_bounds__(256) k_celt_kind_count(const anm_celt_frame_t *__restrict__ recs, const anm_celt_job_t *__restrict__ jobs, uint32_t n_jobs,
                                                         uint32_t *place, uint32_t *count) {
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n_jobs) return;
    /* a batch is a few kinds, so the lanes of a warp mostly want the same counter: one atomic per kind and warp, the lanes take consecutive places */
    const uint32_t kind = celt_kind(recs[j], jobs[j]);
    const unsigned lane = threadIdx.x & 31u, peers = __match_any_sync(__activemask(), kind);
    const int leader = __ffs((int)peers) - 1;
    uint32_t first = 0;
    if ((int)lane == leader) first = atomicAdd(&count[kind], (uint32_t)__popc(peers));
    first = __shfl_sync(peers, first, leader);
    place[j] = first + (uint32_t)__popc(peers & ((1u << lane) - 1u));
}
__global__ void __launch_bounds__(kSortKeys) k_celt_kind_scan(uint32_t *count) { /* group sizes -> group starts, in place */
    __shared__ uint32_t sh[kSortKeys];
    const uint32_t t = threadIdx.x, own = count[t];
    sh[t] = own;
    __syncthreads();
    for (uint32_t d = 1; d < kSortKeys; d <<= 1) {
        const uint32_t v = t >= d ? sh[t - d] : 0u;
        __syncthreads();
        sh[t] += v;
        __syncthreads();
    }
    count[t] = sh[t] - own;
}
__global__ void __launch_bounds__(256) k_celt_kind_place(const anm_celt_frame_t *__restrict__ recs, const anm_celt_job_t *__restrict__ jobs, uint32_t n_jobs,
                                                         const uint32_t *__restrict__ place, const uint32_t *__restrict__ start, uint32_t *order) {
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j < n_jobs) order[start[celt_kind(recs[j], jobs[j])] + place[j]] = j;
}

/* stage 2, one thread per FRAME.  What a frame works on -- the folding source, the band being decoded, the pulse vector and the reordering scratch,
 * 5.1 KB -- sits in thread-local memory: its layout puts the same element of the 32 lanes of a warp side by side, and since the band and partition
 * walks of anm_celt_entropy.h keep the lanes in step, their accesses fall into the same lines.  A finished band is copied to the output once, where
 * working in the output in place wrote every coefficient half a dozen times through the write-through L1 (204,800 frames, before the reordering loops
 * lost their divisions: 32.5 ms with the working storage in a global slab, 30.7 ms thread-local without the band, 26.5 ms with it).  Register budget for
 * 10 / 12 / 14 / 16 resident blocks of 64 threads (96 / 80 / 72 / 64 registers): entropy + spectrum 20.2 / 18.3 / 19.2 / 19.0 ms.  Bound by latency
 * along the dependent chains of a frame and by instruction fetch (the warps of an SM are all somewhere else in 20,000 instructions), not by
 * bandwidth.  A warp-per-frame form with the working set in shared memory was 3 x slower (anm_celt_vec.h). */
#ifndef ANM_CELT_SPEC_MINB
#define ANM_CELT_SPEC_MINB 12 /* resident blocks of 64 threads the register budget is set for */
#endif
__global__ void __launch_bounds__(64, ANM_CELT_SPEC_MINB) k_celt_spectrum(const anm_celt_tables_t *__restrict__ t, const anm_celt_job_t *__restrict__ jobs, uint32_t n_jobs,
                                                      const uint8_t *__restrict__ bytes, uint32_t mask, const anm_celt_frame_t *__restrict__ recs,
                                                      const ce_hist_t *__restrict__ hist, const ce_resume_t *__restrict__ resume, const uint32_t *__restrict__ order,
                                                      int16_t *x, uint32_t x_stride, uint8_t *collapse) {
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x, nthr = gridDim.x * blockDim.x;
    int16_t l_norm[CE_SPEC_NORM], l_tmp[CE_SPEC_TMP], l_band[CE_SPEC_BAND];
    int l_iy[CE_SPEC_IY];
    ce_spec_t sp;
    sp.norm = l_norm;
    sp.tmp = l_tmp;
    sp.iy = l_iy;
    sp.band = l_band;
    sp.lane = 0;
    sp.nl = 1;
    sp.spread = 0;
    for (uint32_t k = tid; k < n_jobs; k += nthr) {
        const uint32_t j = order[k];
        const anm_celt_job_t job = jobs[j];
        if (recs[j].flags & ANM_CELT_F_LOST) continue;
        uint8_t cm[2 * ANM_CE_NB];
        int16_t *X = x + (size_t)j * x_stride;
        const int C = job.channels, NF = 120 << job.lm;
        const int rc = anm_celt_frame_spectrum(t, bytes, mask, job.offset, job.len, C, job.lm, job.end_band, job.flags & ANM_CELT_JOB_DISABLE_INV, &hist[j], &recs[j],
                                               &resume[j], &sp, X, cm);
        if (rc != 0) continue;
        /* zero above the end band (the last band's part served as scratch) */
        const int ncoded = (1 << job.lm) * t->ebands[job.end_band];
        for (int c = 0; c < C; ++c)
            for (int i = ncoded; i < NF; ++i) X[c * NF + i] = 0;
        if (collapse)
            for (int i = 0; i < 2 * ANM_CE_NB; ++i) collapse[(size_t)j * (2 * ANM_CE_NB) + i] = cm[i];
    }
}

/* stage 3, frame-parallel part: one WARP per (frame, output channel) -- denormalisation and the raw inverse-MDCT blocks of that channel; the butterflies
 * of an FFT stage, the rotations and the band loops go over the lanes, the coefficients sit in shared memory (anm_celt_synth.h): 7.5 KB per warp, 28
 * warps per SM (a warp per frame with both channels' 15 KB: 12 warps per SM) */
constexpr uint32_t kBlkWarps = 4;
constexpr uint32_t kBlkWarpBytes = 2u * 960u * 4u; /* freq [960] | raw [960], int32 */
__global__ void __launch_bounds__(kBlkWarps * 32) k_celt_blocks(const anm_celt_tables_t *__restrict__ t, const anm_celt_synth_tables_t *__restrict__ stb,
                                                                const anm_celt_job_t *__restrict__ jobs, const uint32_t *__restrict__ stream_begin,
                                                                uint32_t n_streams, uint32_t n_jobs, const anm_celt_frame_t *__restrict__ recs,
                                                                const anm_celt_synth_t *__restrict__ synth, const int16_t *__restrict__ x, int32_t *raw) {
    extern __shared__ __align__(16) unsigned char blk_smem[];
    const int lane = threadIdx.x & 31;
    const uint32_t w = threadIdx.x >> 5;
    int32_t *fq = reinterpret_cast<int32_t *>(blk_smem + w * kBlkWarpBytes), *rw = fq + 960;
    for (uint32_t it = blockIdx.x * kBlkWarps + w; it < 2u * n_jobs; it += gridDim.x * kBlkWarps) {
        const uint32_t j = it >> 1;
        const int c = (int)(it & 1u);
        const anm_celt_frame_t *fr = &recs[j];
        if (fr->flags & ANM_CELT_F_LOST) continue;
        /* the frame's stream: the last s with stream_begin[s] <= j */
        uint32_t lo = 0, hi = n_streams;
        while (hi - lo > 1) {
            const uint32_t mid = (lo + hi) >> 1;
            if (stream_begin[mid] <= j) lo = mid;
            else hi = mid;
        }
        int CC = (int)synth[lo].out_channels;
        if (CC == 0) CC = jobs[stream_begin[lo]].channels;
        if (c >= CC) continue;
        const int n = 120 << fr->lm;
        cs_channel_blocks(t, stb, x + (size_t)j * 1920, fr->band_e, fr->channels, CC, c, fr->lm, fr->pad[0], (fr->flags & ANM_CELT_F_TRANSIENT) != 0,
                          (fr->flags & ANM_CELT_F_SILENCE) != 0, fq, rw, lane, 32);
        int32_t *ro = raw + (size_t)j * 1920 + (size_t)c * n;
        for (int i = lane; i < n; i += 32) ro[i] = rw[i];
        __syncwarp();
    }
}

/* stage 3, per-stream part: one WARP per (stream, output channel) -- window overlap-add and pitch post-filter with the channel's output history (8.7 KB)
 * in shared memory for the whole call; copies, window mix, saturation and the post-filter go over the lanes.  The filtered samples replace the raw blocks
 * of the frame in `raw`; k_celt_deemphasis turns them into PCM. */
constexpr uint32_t kOvlWarps = 8;
constexpr uint32_t kOvlWarpBytes = (2048u + 120u) * 4u; /* the history */
__global__ void __launch_bounds__(kOvlWarps * 32) k_celt_overlap(const anm_celt_synth_tables_t *__restrict__ stb, const anm_celt_job_t *__restrict__ jobs,
                                                                 const uint32_t *__restrict__ stream_begin, uint32_t n_streams,
                                                                 const anm_celt_frame_t *__restrict__ recs, anm_celt_synth_t *synth, int32_t *raw) {
    extern __shared__ __align__(16) unsigned char ovl_smem[];
    const int lane = threadIdx.x & 31;
    const uint32_t w = threadIdx.x >> 5, id = blockIdx.x * kOvlWarps + w, s = id >> 1;
    const int c = (int)(id & 1u);
    int32_t *mem = reinterpret_cast<int32_t *>(ovl_smem + w * kOvlWarpBytes);
    const bool live = s < n_streams;
    anm_celt_synth_t *sy = live ? &synth[s] : nullptr;
    int CC = 0;
    cs_pf_t pf = {};
    if (live) {
        CC = (int)sy->out_channels;
        if (CC == 0) CC = stream_begin[s + 1] > stream_begin[s] ? jobs[stream_begin[s]].channels : 1;
        cs_pf_load(&pf, sy);
    }
    __syncthreads(); /* both channels of a stream (neighbouring warps of this block) have read the stream's state before either writes it */
    if (!live) return;
    if (c < CC) {
        for (int i = lane; i < 2048 + 120; i += 32) mem[i] = sy->mem[c][i];
        __syncwarp();
        for (uint32_t j = stream_begin[s]; j < stream_begin[s + 1]; ++j) {
            const anm_celt_frame_t *fr = &recs[j];
            if (fr->flags & ANM_CELT_F_LOST) continue;
            const int N = 120 << fr->lm;
            int32_t *rc = raw + (size_t)j * 1920 + (size_t)c * N;
            cs_channel_signal(stb, mem, &pf, fr, rc, lane, 32);
            for (int i = lane; i < N; i += 32) rc[i] = mem[2048 - N + i];
            __syncwarp();
        }
        for (int i = lane; i < 2048 + 120; i += 32) sy->mem[c][i] = mem[i];
    }
    if (c == 0 && lane == 0) {
        cs_pf_store(&pf, sy);
        sy->out_channels = (uint32_t)CC;
    }
}

__device__ __forceinline__ void prefetch_l2(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

/* stage 3, the last step: de-emphasis, one THREAD per (stream, output channel) -- a one-pole recurrence over the channel's samples in order, 16-bit PCM
 * out.  Every step waits for the one before it, so what counts is that nothing else stands in that chain: the samples come in eight at a time (two
 * 16-byte loads, four such groups = one line in flight), and go out eight at a time -- the two channels of a stereo stream sit in neighbouring lanes
 * and trade halves, so that each lane stores 16 bytes of interleaved PCM (vec: the PCM rows are 16-byte aligned; single 16-bit stores otherwise).
 * The loops run to the warp's maxima with the lanes that have nothing left switched off, not gone, so that the exchange is a plain full-warp shuffle.
 * Measured (8,192 channels of 48,000 samples): 2.1 ms, 44 ns a sample where the chain itself is about 10 ns -- with two warps per SM no latency of
 * the memory system is hidden (neither loading further ahead nor prefetching into the L2 moved it); it is what the number of channels in a batch
 * gives.  Inside k_celt_overlap on one lane of a warp per channel the same recurrence took 7 ms.  k_celt_overlap has set out_channels. */
__global__ void __launch_bounds__(64) k_celt_deemphasis(const uint32_t *__restrict__ stream_begin, uint32_t n_streams, const anm_celt_frame_t *__restrict__ recs,
                                                        anm_celt_synth_t *synth, const int32_t *__restrict__ sig, int16_t *pcm, uint32_t pcm_stride, int vec) {
    constexpr unsigned kAll = 0xFFFFFFFFu;
    const uint32_t id = blockIdx.x * blockDim.x + threadIdx.x, s = id >> 1;
    const int c = (int)(id & 1u);
    anm_celt_synth_t *sy = s < n_streams ? &synth[s] : nullptr;
    const int CC = sy ? (int)sy->out_channels : 0;
    const bool mine = c < CC;
    const uint32_t jb = mine ? stream_begin[s] : 0u, nf = mine ? stream_begin[s + 1] - jb : 0u;
    const uint32_t nf_max = __reduce_max_sync(kAll, nf);
    int32_t pm = mine ? sy->preemph_mem[c] : 0;
    for (uint32_t f = 0; f < nf_max; ++f) {
        const uint32_t j = jb + f;
        const bool on = f < nf && !(recs[j].flags & ANM_CELT_F_LOST);
        const int N = on ? 120 << recs[j].lm : 0, ng = N >> 3;
        const int ng_max = (int)__reduce_max_sync(kAll, (unsigned)ng);
        if (ng_max == 0) continue;
        const int4 *in = reinterpret_cast<const int4 *>(sig + (size_t)j * 1920 + (size_t)c * N);
        int16_t *po = pcm + (size_t)j * pcm_stride;
        /* the samples were written by another kernel and are far too many for the L2: the L2 is asked for the lines 1 KB ahead, and for the head of
         * the stream's next frame at the start of this one */
        if (f + 1 < nf) {
            const char *nx = reinterpret_cast<const char *>(sig + (size_t)(j + 1) * 1920 + (size_t)c * (120 << recs[j + 1].lm));
            for (int k = 0; k < 4; ++k) prefetch_l2(nx + 128 * k);
        }
        int4 q[4][2];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            q[k][0] = on ? in[2 * k] : make_int4(0, 0, 0, 0);
            q[k][1] = on ? in[2 * k + 1] : make_int4(0, 0, 0, 0);
        }
        for (int g0 = 0; g0 < ng_max; g0 += 4) {
            if (32 * g0 + 1024 + 512 <= 4 * N) {
#pragma unroll
                for (int k = 0; k < 4; ++k) prefetch_l2(reinterpret_cast<const char *>(in) + 32 * g0 + 1024 + 128 * k);
            }
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int g = g0 + k, i = 8 * g;
                if (g >= ng_max) break;
                const bool act = g < ng;
                const int4 a0 = q[k][0], b0 = q[k][1];
                if (g + 4 < ng) {
                    q[k][0] = in[2 * (g + 4)];
                    q[k][1] = in[2 * (g + 4) + 1];
                }
                uint32_t w[4] = {0u, 0u, 0u, 0u};
                if (act) {
                    int16_t v[8];
                    v[0] = cs_deemphasis_step(a0.x, &pm);
                    v[1] = cs_deemphasis_step(a0.y, &pm);
                    v[2] = cs_deemphasis_step(a0.z, &pm);
                    v[3] = cs_deemphasis_step(a0.w, &pm);
                    v[4] = cs_deemphasis_step(b0.x, &pm);
                    v[5] = cs_deemphasis_step(b0.y, &pm);
                    v[6] = cs_deemphasis_step(b0.z, &pm);
                    v[7] = cs_deemphasis_step(b0.w, &pm);
                    if (!vec)
                        for (int e = 0; e < 8; ++e) po[(i + e) * CC + c] = v[e];
                    for (int e = 0; e < 4; ++e) w[e] = (uint32_t)(uint16_t)v[2 * e] | (uint32_t)(uint16_t)v[2 * e + 1] << 16;
                }
                if (!vec) continue;
                /* stereo: the left lane writes samples i .. i+3 of both channels, the right lane i+4 .. i+7; each hands the other the half it does not write */
                const uint32_t x0 = __shfl_xor_sync(kAll, c ? w[0] : w[2], 1), x1 = __shfl_xor_sync(kAll, c ? w[1] : w[3], 1);
                if (!act) continue;
                if (CC == 1) {
                    *reinterpret_cast<uint4 *>(po + i) = make_uint4(w[0], w[1], w[2], w[3]);
                } else {
                    const uint32_t l0 = c ? x0 : w[0], l1 = c ? x1 : w[1], r0 = c ? w[2] : x0, r1 = c ? w[3] : x1;
                    *reinterpret_cast<uint4 *>(po + (i + 4 * c) * 2) =
                        make_uint4(__byte_perm(l0, r0, 0x5410), __byte_perm(l0, r0, 0x7632), __byte_perm(l1, r1, 0x5410), __byte_perm(l1, r1, 0x7632));
                }
            }
        }
    }
    if (mine) sy->preemph_mem[c] = pm;
}

} /* namespace */

extern "C" int anm_celt_ctx_create(int device, anm_celt_ctx_t **out) {
    if (!out) return ANM_ERR_ARG;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        anm_set_error("no CUDA device: the CELT entropy decoder has no CPU fallback");
        return ANM_ERR_CUDA;
    }
    if (device < 0 || device >= ndev) return ANM_ERR_ARG;
    anm_celt_ctx *c = new (std::nothrow) anm_celt_ctx();
    anm_celt_tables_t *h = new (std::nothrow) anm_celt_tables_t();
    if (!c || !h) { delete c; delete h; return ANM_ERR_NOMEM; }
    c->device = device;
    c->d_tables = nullptr;
    c->d_scratch = nullptr;
    c->scratch_frames = 0;
    c->d_hist = nullptr;
    c->hist_frames = 0;
    c->d_resume = nullptr;
    c->resume_frames = 0;
    c->d_order = nullptr;
    c->order_words = 0;
    c->d_synth_tables = nullptr;
    c->d_x = nullptr;
    c->x_frames = 0;
    c->d_raw = nullptr;
    c->raw_frames = 0;
    int rc = anm_celt_tables_build(h);
    if (rc == ANM_OK && (cudaSetDevice(device) != cudaSuccess || cudaMalloc(&c->d_tables, sizeof *h) != cudaSuccess ||
                         cudaMemcpy(c->d_tables, h, sizeof *h, cudaMemcpyHostToDevice) != cudaSuccess)) {
        anm_set_error("anm_celt_ctx_create: %s", cudaGetErrorString(cudaGetLastError()));
        rc = ANM_ERR_CUDA;
    }
    delete h;
    if (rc == ANM_OK) {
        anm_celt_synth_tables_t *hs = new (std::nothrow) anm_celt_synth_tables_t();
        if (!hs) rc = ANM_ERR_NOMEM;
        else if ((rc = anm_celt_synth_tables_build(hs)) == ANM_OK &&
                 (cudaMalloc(&c->d_synth_tables, sizeof *hs) != cudaSuccess || cudaMemcpy(c->d_synth_tables, hs, sizeof *hs, cudaMemcpyHostToDevice) != cudaSuccess)) {
            anm_set_error("anm_celt_ctx_create: %s", cudaGetErrorString(cudaGetLastError()));
            rc = ANM_ERR_CUDA;
        }
        delete hs;
    }
    if (rc != ANM_OK) {
        cudaFree(c->d_tables);
        cudaFree(c->d_synth_tables);
        delete c;
        return rc;
    }
    /* the band splitting recurses (at most five levels deep); give the threads room for it */
    size_t lim = 0;
    if (cudaDeviceGetLimit(&lim, cudaLimitStackSize) == cudaSuccess && lim < 8192) cudaDeviceSetLimit(cudaLimitStackSize, 8192);
    *out = c;
    return ANM_OK;
}

extern "C" void anm_celt_ctx_destroy(anm_celt_ctx_t *c) {
    if (!c) return;
    cudaSetDevice(c->device);
    cudaFree(c->d_tables);
    cudaFree(c->d_scratch);
    cudaFree(c->d_hist);
    cudaFree(c->d_resume);
    cudaFree(c->d_order);
    cudaFree(c->d_synth_tables);
    cudaFree(c->d_x);
    cudaFree(c->d_raw);
    delete c;
}

/* grows a per-frame / per-thread device array of the context; the stream is drained first: launches in flight may still use the old one */
template <typename T>
static int grow(T **p, size_t *have, size_t want, cudaStream_t s, const char *what) {
    if (want <= *have) return ANM_OK;
    if (cudaStreamSynchronize(s) != cudaSuccess) { anm_set_error("%s: %s", what, cudaGetErrorString(cudaGetLastError())); return ANM_ERR_CUDA; }
    cudaFree(*p);
    *p = nullptr;
    *have = 0;
    if (cudaMalloc(p, want * sizeof(T)) != cudaSuccess) {
        anm_set_error("%s: out of device memory (%zu bytes)", what, want * sizeof(T));
        cudaGetLastError();
        return ANM_ERR_NOMEM;
    }
    *have = want;
    return ANM_OK;
}

static int entropy_impl(anm_celt_ctx_t *c, const anm_celt_job_t *d_jobs, const uint32_t *d_stream_begin, uint32_t n_streams, uint32_t n_jobs,
                        const uint8_t *d_bytes, uint32_t bytes_mask, anm_celt_stream_t *d_streams, anm_celt_frame_t *d_out, ce_hist_t *d_hist, ce_resume_t *d_resume,
                        cudaStream_t s) {
    const int rcg = grow(&c->d_scratch, &c->scratch_frames, (size_t)n_jobs * 4 * ANM_CE_NB, s, "anm_celt_entropy_device");
    if (rcg != ANM_OK) return rcg;
    k_celt_entropy<<<(n_jobs + 127u) / 128u, 128, 0, s>>>(c->d_tables, d_jobs, n_jobs, d_bytes, bytes_mask, c->d_scratch, d_out, d_resume);
    k_celt_energies<<<(n_streams + 3u) / 4u, 128, 0, s>>>(d_stream_begin, n_streams, c->d_scratch, d_streams, d_out, d_hist);
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        anm_set_error("k_celt_entropy launch failed: %s", cudaGetErrorString(e));
        return ANM_ERR_CUDA;
    }
    return ANM_OK;
}

static int check_args(anm_celt_ctx_t *c, const void *d_jobs, const void *d_stream_begin, const void *d_streams, const void *d_out, uint32_t n_streams,
                      uint32_t bytes_mask) {
    if (!c || ((!d_jobs || !d_stream_begin || !d_streams || !d_out) && n_streams)) return ANM_ERR_ARG;
    if (bytes_mask != 0xFFFFFFFFu && (bytes_mask & (bytes_mask + 1u)) != 0u) return ANM_ERR_ARG;
    return ANM_OK;
}

extern "C" int anm_celt_entropy_device(anm_celt_ctx_t *c, const anm_celt_job_t *d_jobs, const uint32_t *d_stream_begin, uint32_t n_streams, uint32_t n_jobs,
                                       const uint8_t *d_bytes, uint32_t bytes_mask, anm_celt_stream_t *d_streams, anm_celt_frame_t *d_out, void *stream) {
    const int rc = check_args(c, d_jobs, d_stream_begin, d_streams, d_out, n_streams, bytes_mask);
    if (rc != ANM_OK) return rc;
    if (n_streams == 0 || n_jobs == 0) return ANM_OK;
    return entropy_impl(c, d_jobs, d_stream_begin, n_streams, n_jobs, d_bytes, bytes_mask, d_streams, d_out, nullptr, nullptr, (cudaStream_t)stream);
}

extern "C" int anm_celt_spectrum_device(anm_celt_ctx_t *c, const anm_celt_job_t *d_jobs, const uint32_t *d_stream_begin, uint32_t n_streams, uint32_t n_jobs,
                                        const uint8_t *d_bytes, uint32_t bytes_mask, anm_celt_stream_t *d_streams, anm_celt_frame_t *d_out, int16_t *d_x,
                                        uint32_t x_stride, uint8_t *d_collapse, void *stream) {
    int rc = check_args(c, d_jobs, d_stream_begin, d_streams, d_out, n_streams, bytes_mask);
    if (rc != ANM_OK) return rc;
    if ((!d_x && n_jobs) || x_stride < 120u) return ANM_ERR_ARG;
    if (n_streams == 0 || n_jobs == 0) return ANM_OK;
    cudaStream_t s = (cudaStream_t)stream;
    if ((rc = grow(&c->d_hist, &c->hist_frames, (size_t)n_jobs, s, "anm_celt_spectrum_device")) != ANM_OK) return rc;
    if ((rc = grow(&c->d_resume, &c->resume_frames, (size_t)n_jobs, s, "anm_celt_spectrum_device")) != ANM_OK) return rc;
    if ((rc = entropy_impl(c, d_jobs, d_stream_begin, n_streams, n_jobs, d_bytes, bytes_mask, d_streams, d_out, c->d_hist, c->d_resume, s)) != ANM_OK) return rc;
    if ((rc = grow(&c->d_order, &c->order_words, 2u * (size_t)n_jobs + kSortKeys, s, "anm_celt_spectrum_device")) != ANM_OK) return rc;
    uint32_t *order = c->d_order, *place = order + n_jobs, *count = place + n_jobs;
    cudaMemsetAsync(count, 0, kSortKeys * sizeof(uint32_t), s);
    k_celt_kind_count<<<(n_jobs + 255u) / 256u, 256, 0, s>>>(d_out, d_jobs, n_jobs, place, count);
    k_celt_kind_scan<<<1, kSortKeys, 0, s>>>(count);
    k_celt_kind_place<<<(n_jobs + 255u) / 256u, 256, 0, s>>>(d_out, d_jobs, n_jobs, place, count, order);
    /* one frame per thread: the hardware balances the very uneven frames block by block */
    k_celt_spectrum<<<(n_jobs + 63u) / 64u, 64, 0, s>>>(c->d_tables, d_jobs, n_jobs, d_bytes, bytes_mask, d_out, c->d_hist, c->d_resume, order, d_x, x_stride,
                                                        d_collapse);
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        anm_set_error("k_celt_spectrum launch failed: %s", cudaGetErrorString(e));
        return ANM_ERR_CUDA;
    }
    return ANM_OK;
}

extern "C" int anm_celt_decode_device(anm_celt_ctx_t *c, const anm_celt_job_t *d_jobs, const uint32_t *d_stream_begin, uint32_t n_streams, uint32_t n_jobs,
                                      const uint8_t *d_bytes, uint32_t bytes_mask, anm_celt_stream_t *d_streams, anm_celt_synth_t *d_synth, anm_celt_frame_t *d_out,
                                      int16_t *d_pcm, uint32_t pcm_stride, void *stream) {
    int rc = check_args(c, d_jobs, d_stream_begin, d_streams, d_out, n_streams, bytes_mask);
    if (rc != ANM_OK) return rc;
    if (((!d_pcm || !d_synth) && n_jobs) || pcm_stride < 120u) return ANM_ERR_ARG;
    if (n_streams == 0 || n_jobs == 0) return ANM_OK;
    cudaStream_t s = (cudaStream_t)stream;
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, c->device);
    if ((rc = grow(&c->d_x, &c->x_frames, (size_t)n_jobs * 1920u, s, "anm_celt_decode_device")) != ANM_OK) return rc;
    if ((rc = grow(&c->d_raw, &c->raw_frames, (size_t)n_jobs * 1920u, s, "anm_celt_decode_device")) != ANM_OK) return rc;
    if ((rc = anm_celt_spectrum_device(c, d_jobs, d_stream_begin, n_streams, n_jobs, d_bytes, bytes_mask, d_streams, d_out, c->d_x, 1920u, nullptr, stream)) != ANM_OK)
        return rc;
    cudaFuncSetAttribute((const void *)k_celt_blocks, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(kBlkWarps * kBlkWarpBytes)); /* per device */
    cudaFuncSetAttribute((const void *)k_celt_overlap, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(kOvlWarps * kOvlWarpBytes));
    const uint32_t blk_blocks = (uint32_t)std::min<uint64_t>((2ull * n_jobs + kBlkWarps - 1u) / kBlkWarps, (uint64_t)sms * 7u);
    k_celt_blocks<<<blk_blocks, kBlkWarps * 32, kBlkWarps * kBlkWarpBytes, s>>>(c->d_tables, c->d_synth_tables, d_jobs, d_stream_begin, n_streams, n_jobs, d_out, d_synth,
                                                                                 c->d_x, c->d_raw);
    k_celt_overlap<<<(2u * n_streams + kOvlWarps - 1u) / kOvlWarps, kOvlWarps * 32, kOvlWarps * kOvlWarpBytes, s>>>(c->d_synth_tables, d_jobs, d_stream_begin, n_streams,
                                                                                                                       d_out, d_synth, c->d_raw);
    const int vec = (reinterpret_cast<uintptr_t>(d_pcm) & 15u) == 0 && (pcm_stride & 7u) == 0;
    k_celt_deemphasis<<<(2u * n_streams + 63u) / 64u, 64, 0, s>>>(d_stream_begin, n_streams, d_out, d_synth, c->d_raw, d_pcm, pcm_stride, vec);
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        anm_set_error("k_celt_blocks / k_celt_overlap / k_celt_deemphasis launch failed: %s", cudaGetErrorString(e));
        return ANM_ERR_CUDA;
    }
    return ANM_OK;
}

static int host_impl(const anm_celt_job_t *jobs, const uint32_t *stream_begin, uint32_t n_streams, const uint8_t *bytes, size_t n_bytes,
                     anm_celt_stream_t *streams, anm_celt_frame_t *out, int16_t *x, uint32_t x_stride, uint8_t *collapse, bool spectrum) {
    if ((!jobs || !stream_begin || !streams || !out) && n_streams) return ANM_ERR_ARG;
    if (spectrum && ((!x && n_streams) || x_stride < 120u)) return ANM_ERR_ARG;
    if (n_streams == 0) return ANM_OK;
    const uint32_t n_jobs = stream_begin[n_streams];
    for (uint32_t i = 0; i < n_jobs; ++i)
        if ((size_t)jobs[i].offset + jobs[i].len > n_bytes) return ANM_ERR_ARG;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) {
        cudaGetLastError();
        anm_set_error("no CUDA device: the CELT decoder stages have no CPU fallback");
        return ANM_ERR_CUDA;
    }
    anm_celt_ctx_t *c = nullptr;
    int rc = anm_celt_ctx_create(dev, &c);
    if (rc != ANM_OK) return rc;
    anm_celt_job_t *d_j = nullptr;
    uint32_t *d_sb = nullptr;
    uint8_t *d_b = nullptr, *d_cm = nullptr;
    anm_celt_stream_t *d_s = nullptr;
    anm_celt_frame_t *d_o = nullptr;
    int16_t *d_x = nullptr;
    const size_t nj = n_jobs ? n_jobs : 1, x_bytes = spectrum ? nj * x_stride * sizeof(int16_t) : 0;
    rc = ANM_ERR_CUDA;
    if (cudaMalloc(&d_j, nj * sizeof *d_j) == cudaSuccess && cudaMalloc(&d_sb, (n_streams + 1) * sizeof *d_sb) == cudaSuccess &&
        cudaMalloc(&d_b, n_bytes ? n_bytes : 1) == cudaSuccess && cudaMalloc(&d_s, n_streams * sizeof *d_s) == cudaSuccess &&
        cudaMalloc(&d_o, nj * sizeof *d_o) == cudaSuccess && (!spectrum || (cudaMalloc(&d_x, x_bytes) == cudaSuccess && cudaMalloc(&d_cm, nj * 42) == cudaSuccess)) &&
        cudaMemcpy(d_j, jobs, n_jobs * sizeof *d_j, cudaMemcpyHostToDevice) == cudaSuccess &&
        cudaMemcpy(d_sb, stream_begin, (n_streams + 1) * sizeof *d_sb, cudaMemcpyHostToDevice) == cudaSuccess &&
        cudaMemcpy(d_b, bytes, n_bytes, cudaMemcpyHostToDevice) == cudaSuccess &&
        cudaMemcpy(d_s, streams, n_streams * sizeof *d_s, cudaMemcpyHostToDevice) == cudaSuccess &&
        (!spectrum || (cudaMemcpy(d_x, x, x_bytes, cudaMemcpyHostToDevice) == cudaSuccess && cudaMemset(d_cm, 0, nj * 42) == cudaSuccess))) {
        rc = spectrum ? anm_celt_spectrum_device(c, d_j, d_sb, n_streams, n_jobs, d_b, 0xFFFFFFFFu, d_s, d_o, d_x, x_stride, d_cm, nullptr)
                      : anm_celt_entropy_device(c, d_j, d_sb, n_streams, n_jobs, d_b, 0xFFFFFFFFu, d_s, d_o, nullptr);
        if (rc == ANM_OK && (cudaDeviceSynchronize() != cudaSuccess || cudaMemcpy(out, d_o, n_jobs * sizeof *d_o, cudaMemcpyDeviceToHost) != cudaSuccess ||
                             cudaMemcpy(streams, d_s, n_streams * sizeof *d_s, cudaMemcpyDeviceToHost) != cudaSuccess ||
                             (spectrum && (cudaMemcpy(x, d_x, x_bytes, cudaMemcpyDeviceToHost) != cudaSuccess ||
                                           (collapse && cudaMemcpy(collapse, d_cm, (size_t)n_jobs * 42, cudaMemcpyDeviceToHost) != cudaSuccess)))))
            rc = ANM_ERR_CUDA;
    }
    if (rc == ANM_ERR_CUDA) anm_set_error("anm_celt_*_host: %s", cudaGetErrorString(cudaGetLastError()));
    cudaFree(d_j);
    cudaFree(d_sb);
    cudaFree(d_b);
    cudaFree(d_s);
    cudaFree(d_o);
    cudaFree(d_x);
    cudaFree(d_cm);
    anm_celt_ctx_destroy(c);
    return rc;
}

extern "C" int anm_celt_entropy_host(const anm_celt_job_t *jobs, const uint32_t *stream_begin, uint32_t n_streams, const uint8_t *bytes, size_t n_bytes,
                                     anm_celt_stream_t *streams, anm_celt_frame_t *out) {
    return host_impl(jobs, stream_begin, n_streams, bytes, n_bytes, streams, out, nullptr, 0, nullptr, false);
}

/* x is read as well as written: coefficients the decode does not touch keep what the caller put there */
extern "C" int anm_celt_spectrum_host(const anm_celt_job_t *jobs, const uint32_t *stream_begin, uint32_t n_streams, const uint8_t *bytes, size_t n_bytes,
                                      anm_celt_stream_t *streams, anm_celt_frame_t *out, int16_t *x, uint32_t x_stride, uint8_t *collapse) {
    return host_impl(jobs, stream_begin, n_streams, bytes, n_bytes, streams, out, x, x_stride, collapse, true);
}

extern "C" int anm_celt_decode_host(const anm_celt_job_t *jobs, const uint32_t *stream_begin, uint32_t n_streams, const uint8_t *bytes, size_t n_bytes,
                                    anm_celt_stream_t *streams, anm_celt_synth_t *synth, anm_celt_frame_t *out, int16_t *pcm, uint32_t pcm_stride) {
    if ((!jobs || !stream_begin || !streams || !synth || !out || !pcm) && n_streams) return ANM_ERR_ARG;
    if (pcm_stride < 120u) return ANM_ERR_ARG;
    if (n_streams == 0) return ANM_OK;
    const uint32_t n_jobs = stream_begin[n_streams];
    for (uint32_t i = 0; i < n_jobs; ++i)
        if ((size_t)jobs[i].offset + jobs[i].len > n_bytes) return ANM_ERR_ARG;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) {
        cudaGetLastError();
        anm_set_error("no CUDA device: the CELT decoder stages have no CPU fallback");
        return ANM_ERR_CUDA;
    }
    anm_celt_ctx_t *c = nullptr;
    int rc = anm_celt_ctx_create(dev, &c);
    if (rc != ANM_OK) return rc;
    anm_celt_job_t *d_j = nullptr;
    uint32_t *d_sb = nullptr;
    uint8_t *d_b = nullptr;
    anm_celt_stream_t *d_s = nullptr;
    anm_celt_synth_t *d_y = nullptr;
    anm_celt_frame_t *d_o = nullptr;
    int16_t *d_p = nullptr;
    const size_t nj = n_jobs ? n_jobs : 1, p_bytes = nj * pcm_stride * sizeof(int16_t);
    rc = ANM_ERR_CUDA;
    if (cudaMalloc(&d_j, nj * sizeof *d_j) == cudaSuccess && cudaMalloc(&d_sb, (n_streams + 1) * sizeof *d_sb) == cudaSuccess &&
        cudaMalloc(&d_b, n_bytes ? n_bytes : 1) == cudaSuccess && cudaMalloc(&d_s, n_streams * sizeof *d_s) == cudaSuccess &&
        cudaMalloc(&d_y, n_streams * sizeof *d_y) == cudaSuccess && cudaMalloc(&d_o, nj * sizeof *d_o) == cudaSuccess && cudaMalloc(&d_p, p_bytes) == cudaSuccess &&
        cudaMemcpy(d_j, jobs, n_jobs * sizeof *d_j, cudaMemcpyHostToDevice) == cudaSuccess &&
        cudaMemcpy(d_sb, stream_begin, (n_streams + 1) * sizeof *d_sb, cudaMemcpyHostToDevice) == cudaSuccess &&
        cudaMemcpy(d_b, bytes, n_bytes, cudaMemcpyHostToDevice) == cudaSuccess &&
        cudaMemcpy(d_s, streams, n_streams * sizeof *d_s, cudaMemcpyHostToDevice) == cudaSuccess &&
        cudaMemcpy(d_y, synth, n_streams * sizeof *d_y, cudaMemcpyHostToDevice) == cudaSuccess && cudaMemset(d_p, 0, p_bytes) == cudaSuccess) {
        rc = anm_celt_decode_device(c, d_j, d_sb, n_streams, n_jobs, d_b, 0xFFFFFFFFu, d_s, d_y, d_o, d_p, pcm_stride, nullptr);
        if (rc == ANM_OK && (cudaDeviceSynchronize() != cudaSuccess || cudaMemcpy(out, d_o, n_jobs * sizeof *d_o, cudaMemcpyDeviceToHost) != cudaSuccess ||
                             cudaMemcpy(streams, d_s, n_streams * sizeof *d_s, cudaMemcpyDeviceToHost) != cudaSuccess ||
                             cudaMemcpy(synth, d_y, n_streams * sizeof *d_y, cudaMemcpyDeviceToHost) != cudaSuccess ||
                             cudaMemcpy(pcm, d_p, (size_t)n_jobs * pcm_stride * sizeof(int16_t), cudaMemcpyDeviceToHost) != cudaSuccess))
            rc = ANM_ERR_CUDA;
    }
    if (rc == ANM_ERR_CUDA) anm_set_error("anm_celt_decode_host: %s", cudaGetErrorString(cudaGetLastError()));
    cudaFree(d_j);
    cudaFree(d_sb);
    cudaFree(d_b);
    cudaFree(d_s);
    cudaFree(d_y);
    cudaFree(d_o);
    cudaFree(d_p);
    anm_celt_ctx_destroy(c);
    return rc;
}
