/*
 * anmodem.h -- C ABI of the B200-native batched audio-modem receive path.
 *
 * What this boundary replaces.  BASELINE.json's north_star names "the firmware
 * demodulator's C interface (sample buffer in, decoded symbols and frames out)"
 * feeding the nanopb ip.proto decoder.  SURVEY.md section 0 establishes that the
 * reference (tmarsteel/audio-network) contains no such demodulator; the nearest
 * real seams are
 *   - the nanopb input stream the firmware decodes ToReceiver messages from:
 *     hardware/lib/nanopb/src/pb_decode.h:28-46 (struct pb_istream_s) and its
 *     socket-backed instance hardware/src/network.cpp:262-305, consumed by
 *     pb_decode_delimited at hardware/src/network.cpp:406-411;
 *   - the module convention `<module>_initialize()` + handle functions,
 *     hardware/README.md:10-14, hardware/include/playback.hpp:15.
 * The functions below are therefore the interface SURVEY.md section 8(b)
 * proposes for that seam: firmware-idiom single-channel calls (demod_*), and the
 * batched form (anm_demod_*) that the hot path actually runs.  The modem itself
 * (tone set, timing, preamble, framing, CRC and the exact fp32 operation order)
 * is defined by SPEC.md in this repository, not by the reference.
 *
 * All entry points are plain C: pointers and sizes only.  Return values are 0 on
 * success or a negative ANM_ERR_* code (the esp_err_t idiom of the reference,
 * hardware/src/playback.cpp:174-191); nothing throws, nothing falls back to a
 * CPU implementation -- without a CUDA device every compute entry point returns
 * ANM_ERR_CUDA.
 */
#ifndef ANMODEM_H_INCLUDED
#define ANMODEM_H_INCLUDED

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* revision of SPEC.md this library implements bit for bit (frozen: tests/golden/SPEC_VERSION.json) */
#define ANM_SPEC_REVISION 2

#define ANM_MAX_TONES 64
#define ANM_MAX_PREAMBLE 32
#define ANM_SILENCE 0xFF /* tx program entry: no tone during this symbol */

enum {
    ANM_OK = 0,
    ANM_ERR_ARG = -1,      /* invalid argument / configuration */
    ANM_ERR_CUDA = -2,     /* CUDA runtime error or no device */
    ANM_ERR_NOMEM = -3,    /* host or device allocation failed */
    ANM_ERR_ALIGN = -4,    /* buffer not 16-byte aligned / length not a multiple of sym_len */
    ANM_ERR_OVERFLOW = -5, /* an output queue overflowed; results were dropped */
    ANM_ERR_UNSUPPORTED = -6,
    ANM_ERR_FORMAT = -7    /* malformed message: the reference's decoder would return false */
};

/* Modem configuration (SPEC.md section 2). */
typedef struct anm_config {
    uint32_t sample_rate;  /* Hz; informational (44100) */
    uint32_t sym_len;      /* N: samples per symbol; power of two, 32..512 */
    uint32_t hops_per_sym; /* S: timing hypotheses per symbol; 2, 4 or 8 */
    uint32_t n_tones;      /* T: power of two, 2..64; bits/symbol = log2 T */
    uint32_t tone_bin[ANM_MAX_TONES]; /* DFT bin of tone k over a sym_len window */
    uint32_t preamble_len; /* P: 8, 16 or 32 symbols */
    uint8_t preamble[ANM_MAX_PREAMBLE]; /* tone index per preamble symbol */
    uint32_t sync_tol;     /* max preamble symbol mismatches accepted */
    uint32_t max_payload;  /* largest payload length accepted, bytes (<= 4104) */
    uint32_t trk_epoch;    /* symbols between timing-tracker decisions */
    uint32_t trk_thresh;   /* |vote sum| needed to move timing by one hop */
} anm_config_t;

/* One decoded frame.  `offset` indexes the byte arena returned alongside. */
typedef struct anm_frame {
    uint32_t channel;
    uint32_t len;          /* payload bytes */
    uint64_t start_sample; /* stream index of the first preamble sample */
    uint32_t crc_ok;       /* 1 if CRC-16 matched */
    uint32_t offset;       /* payload position in the byte arena */
} anm_frame_t;

/* Per-channel counters since create. */
typedef struct anm_chan_stats {
    uint32_t locks;        /* preambles detected */
    uint32_t header_fail;  /* header check failures */
    uint32_t frames_ok;
    uint32_t frames_bad;   /* CRC-16 failures */
    uint64_t symbols;      /* symbols decided while locked */
    int32_t trk_moves;     /* net timing moves, hops */
    uint32_t reserved;
} anm_chan_stats_t;

/* Transmit-side synthetic channel description (SPEC.md section 6). */
typedef struct anm_tx_params {
    uint64_t seed;          /* noise PRNG seed */
    int64_t start_offset;   /* tx sample position at rx sample 0 (may be negative: leading silence) */
    uint32_t amplitude_q15; /* tone peak amplitude, Q15 of full scale */
    int32_t snr_mdb;        /* full-band SNR in milli-dB; ANM_SNR_CLEAN = no noise */
    int32_t ppm_x1000;      /* transmitter clock error, 1/1000 ppm */
    uint32_t reserved;
} anm_tx_params_t;
#define ANM_SNR_CLEAN INT32_MAX

/* ---- configuration / framing helpers (host C, no device) ---------------- */
int anm_config_preset(const char *name, anm_config_t *out); /* "ref4", "bfsk2", "mfsk8", "mfsk16", "wide64" */
int anm_config_validate(const anm_config_t *cfg);
/* twiddle table [sym_len][n_tones][2] = (cos, sin)(2*pi*bin*m/N) rounded to fp32 */
int anm_twiddles(const anm_config_t *cfg, float *out);
/* SPEC 3: 1 if the hop partials are computed by centre folding (every tone bin a multiple of S/2) */
int anm_config_foldable(const anm_config_t *cfg);
/* SPEC 3: folded twiddles out[H/2][n_tones][2] = (cos, sin)(2 pi bin (k + 1/2) / N) */
int anm_fold_twiddles(const anm_config_t *cfg, float *out);
/* SPEC 3b: 1 if the configuration uses the dense integer basis (n_tones >= 32) */
int anm_config_dense(const anm_config_t *cfg);
/* SPEC 3b: int8 basis out[sym_len][n_tones][2] = (round(127 cos), round(127 sin)) with the quarter-period symmetry */
int anm_basis_q7(const anm_config_t *cfg, int8_t *out);
uint16_t anm_crc16(const uint8_t *data, size_t len, uint16_t crc);
uint8_t anm_crc8(const uint8_t *data, size_t len, uint8_t crc);
/* number of symbols (preamble + header + body) of a frame carrying len bytes */
size_t anm_frame_num_symbols(const anm_config_t *cfg, size_t len);
/* writes tone indices; returns count or 0 if cap too small / len invalid */
size_t anm_frame_symbols(const anm_config_t *cfg, const uint8_t *payload, size_t len,
                         uint8_t *syms, size_t cap);

/* ---- transmitter stand-in (integer DDS; bit-identical on CPU and GPU) ---- */
/* CPU render of rx samples [first_sample, first_sample+n) of one channel that
 * cyclically plays `program` (tone indices or ANM_SILENCE). */
int anm_tx_render(const anm_config_t *cfg, const uint8_t *program, size_t prog_len,
                  const anm_tx_params_t *p, uint64_t first_sample, int16_t *out, size_t n);
/* GPU render of n_ch channels into d_pcm[ch * ch_stride + i]; programs are
 * device-resident: d_programs[ch * prog_stride + j], prog_len[ch] entries used. */
int anm_tx_render_device(const anm_config_t *cfg, const uint8_t *d_programs, size_t prog_stride,
                         const uint32_t *d_prog_len, const anm_tx_params_t *d_params,
                         uint32_t n_ch, uint64_t first_sample, int16_t *d_pcm, size_t ch_stride,
                         size_t n, void *stream);

/* ---- stateless tone-energy pass (parity / debug mode) ------------------- */
/* For each channel and each hop h in [0, n_samples/hop): energies of the window
 * of sym_len samples ending with hop h (zero history before sample 0).
 * d_energy [n_ch][n_hops][n_tones] fp32 (may be NULL), d_sym [n_ch][n_hops]
 * argmax tone (may be NULL), d_emax [n_ch][n_hops] (may be NULL). */
int anm_tone_energies_device(const anm_config_t *cfg, const int16_t *d_pcm, uint32_t n_ch,
                             size_t ch_stride, size_t n_samples, float *d_energy,
                             uint8_t *d_sym, float *d_emax, void *stream);

/* ---- batched streaming demodulator --------------------------------------- */
typedef struct anm_demod anm_demod_t;

#define ANM_FLAG_SYMBOLS 1u /* also record decided symbols per channel */

int anm_demod_create(const anm_config_t *cfg, uint32_t n_channels, int device, uint32_t flags,
                     anm_demod_t **out);
void anm_demod_destroy(anm_demod_t *h);
int anm_demod_reset(anm_demod_t *h);
/* PCM already in HBM: d_pcm[ch * ch_stride + i], i < n_samples; n_samples must be
 * a multiple of sym_len, d_pcm and ch_stride*2 multiples of 16 bytes.  Launches
 * on `stream` (a cudaStream_t, NULL = default) and returns without waiting; a launch on another stream than the
 * handle's previous launch is ordered behind it.  Fastest when n_samples is a multiple of 32 * sym_len (the kernel
 * works in steps of 32 symbol periods; a ragged last step costs a whole one). */
int anm_demod_feed_device(anm_demod_t *h, const int16_t *d_pcm, size_t ch_stride,
                          size_t n_samples, void *stream);
/* Several chunks that are already resident, in ONE launch: chunk c of channel ch at d_pcm[ch * ch_stride + c * chunk_stride + i], i < n_samples
 * (chunk_stride * 2 a multiple of 16 bytes; chunk_stride = n_samples for one contiguous run per channel).  Same result as n_chunks calls of
 * anm_demod_feed_device; the kernel's work items are then (chunk, channel) pairs handed out from one queue, so no SM idles through the tail of a
 * chunk while the next one has not been launched (8,192 channels on 2,960 resident warps: 2.77 waves per chunk, 3 are paid).  Configurations on
 * the tensor-core kernel and handles with ANM_FLAG_SYMBOLS take the chunks one launch at a time. */
int anm_demod_feed_device_chunks(anm_demod_t *h, const int16_t *d_pcm, size_t ch_stride, size_t chunk_stride, size_t n_samples, uint32_t n_chunks,
                                 void *stream);
/* PCM in host memory (pinned for full speed): copies to an internal HBM staging
 * buffer, demodulates, and waits for completion. */
int anm_demod_feed_host(anm_demod_t *h, const int16_t *h_pcm, size_t ch_stride, size_t n_samples);
/* Pipelined form: enqueues the copy and the kernel on the handle's stream and returns.  h_pcm must
 * stay untouched until anm_demod_wait_input() (or a later collect) returns.  Together with
 * anm_demod_collect_upto(h, 1) this overlaps the host-side handling of chunk k with the PCIe
 * transfer of chunk k+1. */
int anm_demod_feed_host_async(anm_demod_t *h, const int16_t *h_pcm, size_t ch_stride, size_t n_samples);
int anm_demod_wait_input(anm_demod_t *h);
/* Like anm_demod_collect for frames only, but waits just for the launch that is `lag` launches
 * behind the most recent one (lag 0 = the latest, lag < 63); works after any kind of feed.  A launch whose frames were
 * already collected yields nothing new. */
long anm_demod_collect_upto(anm_demod_t *h, uint32_t lag);
/* Waits for outstanding work and moves newly produced frames/symbols to the
 * host queues.  Returns number of frames now queued, or a negative error. */
long anm_demod_collect(anm_demod_t *h);
/* Pops up to cap frames in (channel, start_sample) order; payload bytes are
 * appended to `bytes` (frame.offset indexes it).  Returns frames written. */
size_t anm_demod_read_frames(anm_demod_t *h, anm_frame_t *out, size_t cap, uint8_t *bytes,
                             size_t bytes_cap);
/* Moves out EVERYTHING that is queued, in arrival order (per channel still chronological), as two plain copies -- the
 * cheapest way to hand a whole drain to a consumer that does not need the global order.  *n_bytes (optional) receives the
 * payload bytes written.  Returns the number of frames; 0 (queue untouched) when a destination is too small or a sorted
 * read is half way through what it had ordered. */
size_t anm_demod_take_frames(anm_demod_t *h, anm_frame_t *out, size_t cap, uint8_t *bytes, size_t bytes_cap, size_t *n_bytes);
/* Zero-copy form of the same: pointers into the handle's (pinned) queue storage, arrival order; valid until the next
 * collect / feed on this handle.  anm_demod_drop_frames() then empties the queue.  Returns the number of frames. */
size_t anm_demod_peek_frames(anm_demod_t *h, const anm_frame_t **frames, const uint8_t **bytes, size_t *n_bytes);
void anm_demod_drop_frames(anm_demod_t *h);
/* CRC-valid frames and their payload bytes among n records (one pass; what a throughput report needs) */
void anm_frames_summary(const anm_frame_t *frames, size_t n, uint64_t *n_ok, uint64_t *payload_bytes_ok);
/* The device-resident frame record / payload rings (powers of two; masks = size - 1) for consumers that stay on the GPU,
 * e.g. anm_pb_deframe_device() directly on the rings (include/anmodem_pb.h). */
int anm_demod_frame_rings(const anm_demod_t *h, const anm_frame_t **d_frames, uint32_t *frames_mask, const uint8_t **d_bytes,
                          uint32_t *bytes_mask);
/* Order-independent 64-bit checksum of frame records + their payload bytes (frame.offset indexes `bytes`): the same value
 * whatever order or grouping the frames are gathered in (anm_config.c). */
uint64_t anm_frames_digest(const anm_frame_t *frames, size_t n, const uint8_t *bytes);
/* Pops up to cap decided symbols (tone indices) of one channel. */
size_t anm_demod_read_symbols(anm_demod_t *h, uint32_t channel, uint8_t *out, size_t cap);
int anm_demod_stats(anm_demod_t *h, anm_chan_stats_t *out /*[n_channels]*/);
/* 1 if an output queue overflowed since create/reset: frames or symbols were dropped.  Queues hold
 * 512 frames / 32 KiB of payload per channel between two collects. */
int anm_demod_overflowed(const anm_demod_t *h);
/* number of kernels this handle has launched (bench.py's gpu_launches) */
uint64_t anm_demod_launch_count(const anm_demod_t *h);
/* duration in ms of the most recent feed's kernel, measured with CUDA events on
 * the launching stream (waits for it). */
float anm_demod_last_kernel_ms(anm_demod_t *h);
/* Sums the device time (ms, CUDA events on the launching stream) of the kernels launched since the
 * previous call (at most 64 are tracked between calls); returns the number of launches summed. */
int anm_demod_kernel_time(anm_demod_t *h, float *sum_ms);
/* persistent-grid geometry chosen for this handle */
int anm_demod_launch_geometry(const anm_demod_t *h, uint32_t *grid, uint32_t *warps_per_cta, uint32_t *smem_bytes);
/* fills anm_tx_params_t.reserved with the Q20 noise scale anm_tx_render_device reads */
void anm_tx_params_prepare(anm_tx_params_t *p, size_t n);
const char *anm_last_error(void);
const char *anm_version(void);

/* ---- several GPUs behind one handle ------------------------------------------
 * Channels shard over the listed devices in contiguous ranges (sizes differing by at most one); every device has its own
 * anm_demod_t and ONE host thread that issues all CUDA work for it (bound to the CPUs of the GPU's NUMA node when sysfs
 * names one).  No collective and no peer traffic: the only exchange is the host-side gather of frame records, which
 * come out with global channel ids in (channel, start_sample) order.  A device may be listed more than once. */
typedef struct anm_demod_multi anm_demod_multi_t;
int anm_demod_multi_create(const anm_config_t *cfg, uint32_t n_channels, const int *devices, uint32_t n_devices, uint32_t flags,
                           anm_demod_multi_t **out);
void anm_demod_multi_destroy(anm_demod_multi_t *m);
int anm_demod_multi_reset(anm_demod_multi_t *m);
uint32_t anm_demod_multi_num_devices(const anm_demod_multi_t *m);
/* shard d: its CUDA device, channel range and the NUMA node its thread is bound to (-1: not bound) */
int anm_demod_multi_shard(const anm_demod_multi_t *m, uint32_t d, int *device, uint32_t *first_channel, uint32_t *n_channels, int *numa_node);
anm_demod_t *anm_demod_multi_device_handle(anm_demod_multi_t *m, uint32_t d); /* for stats / timing queries of one shard */
/* Page-locked PCM buffer [n_channels][*ch_stride] whose shards are first touched next to their GPUs; freed with the handle. */
int anm_demod_multi_alloc_pcm(anm_demod_multi_t *m, size_t n_samples, int16_t **out, size_t *ch_stride);
/* h_pcm[ch * ch_stride + i] for ALL channels; every device's thread enqueues the copy + kernel of its shard
 * (anm_demod_feed_host_async) and the call returns.  The buffer must stay untouched until anm_demod_multi_wait_input()
 * or a collect returns. */
int anm_demod_multi_feed_host(anm_demod_multi_t *m, const int16_t *h_pcm, size_t ch_stride, size_t n_samples);
int anm_demod_multi_wait_input(anm_demod_multi_t *m);
/* every device drains (in parallel), then the frames are gathered into one queue; return = frames queued or an error */
long anm_demod_multi_collect(anm_demod_multi_t *m);
long anm_demod_multi_collect_upto(anm_demod_multi_t *m, uint32_t lag);
size_t anm_demod_multi_read_frames(anm_demod_multi_t *m, anm_frame_t *out, size_t cap, uint8_t *bytes, size_t bytes_cap);
size_t anm_demod_multi_take_frames(anm_demod_multi_t *m, anm_frame_t *out, size_t cap, uint8_t *bytes, size_t bytes_cap, size_t *n_bytes);
int anm_demod_multi_overflowed(const anm_demod_multi_t *m);
int anm_demod_multi_stats(anm_demod_multi_t *m, anm_chan_stats_t *out /*[n_channels]*/);

/* ---- firmware-idiom single-channel interface ------------ */
typedef struct demod demod_t;
typedef struct demod_frame {
    uint64_t sample_offset;
    uint32_t len;
    uint32_t crc_ok;
    uint8_t bytes[4104];
} demod_frame_t;

int demod_initialize(const anm_config_t *cfg); /* sets the DEFAULT configuration of demodulators created later by demod_create() */
demod_t *demod_create(void);                    /* a demodulator with (its own copy of) the default configuration */
demod_t *demod_create_cfg(const anm_config_t *cfg); /* ... or with an explicit one: no process-wide state involved */
int demod_feed(demod_t *d, const int16_t *pcm, size_t n_samples); /* borrowed input, any length */
size_t demod_read_symbols(demod_t *d, uint8_t *out, size_t cap);
size_t demod_read_frames(demod_t *d, demod_frame_t *out, size_t cap);
void demod_destroy(demod_t *d);

/* ---- chunk pacing of a real-time streaming feed -----------------------
 * The reference transmitter's leaky bucket (transmitter/.../LeakyBucket.kt:9-64; instance "1200 ms of
 * receiver buffer draining 1000 ms per second", MulticastAudioOutput.kt:85) with the clock as a
 * parameter.  Amounts are in the caller's unit (the reference: milliseconds of audio). */
typedef struct anm_pacer {
    int64_t capacity, drain_rate_per_second;
    int64_t last_value, last_value_at_ns;
} anm_pacer_t;
int anm_pacer_init(anm_pacer_t *p, int64_t capacity, int64_t drain_rate_per_second, int64_t now_ns);
int64_t anm_pacer_level(const anm_pacer_t *p, int64_t now_ns);              /* LeakyBucket.currentValue */
int64_t anm_pacer_try_put(anm_pacer_t *p, int64_t amount, int64_t now_ns); /* 0 added, > 0 ns to wait, ANM_ERR_ARG */
int64_t anm_pacer_wait_for_capacity(anm_pacer_t *p, int64_t amount, int64_t *now_ns); /* virtual-clock waitForCapacity */

#ifdef __cplusplus
}
#endif
#endif /* ANMODEM_H_INCLUDED */
