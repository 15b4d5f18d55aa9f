/*
 * anmodem_opus.h -- first stage of the batched Opus receive path: what the
 * reference's decoder learns from a packet before it decodes a single frame.
 *
 * Reference chain: network.cpp:428 playback_queue_audio(opus bytes) -> playback.cpp:115-122 opus_decode()
 * -> opus_decode_native (hardware/lib/libopus/src/opus_decoder.c:627-739), which starts with
 *   opus_packet_get_mode / _bandwidth / _samples_per_frame / _nb_channels   opus_decoder.c:206-219, 973-994; opus.c:170-192
 *   opus_packet_parse_impl(data, len, self_delimited = 0, ...)              opus.c:194-345
 * and rejects the packet when the parse fails (opus_decoder.c:666-669).  anm_opus_parse_* does exactly this
 * step for thousands of packets at once on the GPU, directly on the byte arena the deframer
 * (anm_pb_deframe_*) located the Opus bytes in: per packet the TOC fields, the frame count and every
 * frame's size -- the work list a batched frame decoder consumes.  Integer / byte work, results equal to
 * libopus 1.3.1's (tests/test_opus_parse.py, against the reference's libopus compiled in place).
 *
 * Second stage (anm_celt_entropy_*): the ENTROPY DECODE of CELT frames -- everything celt_decode_with_ec
 * (celt/celt_decoder.c:946-1095) reads off the range coder: frame flags, post-filter parameters, coarse / fine / final band
 * energies, time-frequency and spread decisions, dynamic allocation, the bit allocation and, band by band with every split, the
 * PVQ codeword of each partition.  No symbol of a frame depends on another frame, so one GPU thread per FRAME decodes all frames of all
 * streams at once; only the band energies predict from frame to frame and are chained per stream in a second, short pass.  The check is the reference's own: the range coder's final state of every frame
 * equals OPUS_GET_FINAL_RANGE of libopus 1.3.1.
 *
 * Third stage (anm_celt_spectrum_*): the frames' normalised spectra -- PVQ codewords to pulse vectors, normalisation, spreading rotations,
 * folding / noise filling, the Haar and Hadamard reorderings, stereo merge, collapse masks and anti-collapse, in the reference's fixed-point
 * arithmetic: every coefficient equals what the reference's quant_all_bands() + anti_collapse() hand to celt_synthesis().  The synthesis itself
 * (denormalisation, inverse MDCT, post-filter, de-emphasis) and packet-loss concealment are NOT part of this library yet.
 */
#ifndef ANMODEM_OPUS_H_INCLUDED
#define ANMODEM_OPUS_H_INCLUDED

#include "anmodem_pb.h"

#ifdef __cplusplus
extern "C" {
#endif

enum { /* libopus error codes this stage can produce (opus_defines.h:46-58) */
    ANM_OPUS_BAD_ARG = -1,
    ANM_OPUS_INVALID_PACKET = -4
};
enum { ANM_OPUS_MODE_SILK_ONLY = 1000, ANM_OPUS_MODE_HYBRID = 1001, ANM_OPUS_MODE_CELT_ONLY = 1002 }; /* opus_private.h */

typedef struct anm_opus_packet {
    int32_t count;             /* opus_packet_parse(): number of frames (1..48) or ANM_OPUS_INVALID_PACKET */
    uint8_t toc;               /* first byte of the packet */
    uint8_t channels;          /* opus_packet_get_nb_channels: 1 or 2 */
    uint8_t pad[2];
    int32_t mode;              /* ANM_OPUS_MODE_* (opus_packet_get_mode) */
    int32_t bandwidth;         /* OPUS_BANDWIDTH_* 1101..1105 (opus_packet_get_bandwidth) */
    int32_t samples_per_frame; /* opus_packet_get_samples_per_frame(data, Fs) */
    int32_t payload_offset;    /* bytes from the start of the packet to its first frame (count > 0) */
    int32_t nb_frames;         /* opus_packet_get_nb_frames(): differs from count for packets parse() rejects */
    int32_t nb_samples;        /* opus_packet_get_nb_samples(data, len, Fs) or its error */
    int16_t size[48];          /* size of each frame in bytes (count > 0; zero otherwise) */
} anm_opus_packet_t;           /* 128 bytes */

/* One record per span.  Spans whose status is not ANM_PB_OK, and empty packets, give count = ANM_OPUS_BAD_ARG
 * with every other field zero (the firmware never calls opus_decode for them).  d_spans / d_bytes / d_out in
 * device memory; bytes_mask as for anm_pb_deframe_device; Fs = the decoder's sampling rate (48000 in the
 * reference, playback.cpp:112); stream is a cudaStream_t. */
int anm_opus_parse_device(const anm_pb_span_t *d_spans, uint32_t n, const uint8_t *d_bytes, uint32_t bytes_mask,
                          int32_t Fs, anm_opus_packet_t *d_out, void *stream);
/* host arrays: copies in, runs the kernel, copies out (no CPU fallback: ANM_ERR_CUDA without a device) */
int anm_opus_parse_host(const anm_pb_span_t *spans, size_t n, const uint8_t *bytes, size_t n_bytes, int32_t Fs,
                        anm_opus_packet_t *out);


/* ---- CELT frame entropy decode (row f1, stage 1) -------------------------------------------------------------- */
#define ANM_CELT_BANDS 21          /* bands of the 48 kHz standard mode (celt/modes.c:42-45) */
#define ANM_CELT_ALLOC_VECTORS 11  /* rows of the allocation table (celt/modes.c:48-63) */
#define ANM_CELT_PVQ_ROWS 15       /* U(n, k) is kept for min(n, k) <= 14 ... */
#define ANM_CELT_PVQ_COLS 180      /* ... and max(n, k) < 180 (largest band 176 coefficients, + 1) */

/* Static data of the decoder, built by anm_celt_tables_build(): the normative constants of RFC 6716 (band edges, allocation
 * table, Laplace parameters of the coarse energy) and the tables derived from them by the standard's own formulas (logN, the
 * pulse cache and its caps, the PVQ codebook sizes). */
typedef struct anm_celt_tables {
    int16_t ebands[ANM_CELT_BANDS + 1];
    int16_t logn[ANM_CELT_BANDS];
    int16_t cache_index[5 * ANM_CELT_BANDS]; /* [LM + 1][band] -> offset into cache_bits */
    uint16_t cache_size;
    uint8_t cache_bits[512];
    uint8_t cache_caps[4 * 2 * ANM_CELT_BANDS]; /* [LM][C - 1][band] */
    uint8_t alloc[ANM_CELT_ALLOC_VECTORS * ANM_CELT_BANDS];
    uint8_t e_prob[4 * 2 * 42];               /* [LM][intra][2 * min(band, 20) + {p0, decay}] */
    uint32_t pvq_u[ANM_CELT_PVQ_ROWS * ANM_CELT_PVQ_COLS];
} anm_celt_tables_t;
int anm_celt_tables_build(anm_celt_tables_t *out); /* host; deterministic */

enum { /* anm_celt_frame_t.flags */
    ANM_CELT_F_SILENCE = 1, ANM_CELT_F_POSTFILTER = 2, ANM_CELT_F_TRANSIENT = 4, ANM_CELT_F_INTRA = 8, ANM_CELT_F_DUAL_STEREO = 16,
    ANM_CELT_F_ANTI_COLLAPSE = 32,
    ANM_CELT_F_LOST = 256,     /* frame of <= 1 byte: the reference conceals (celt_decode_lost), nothing is decoded */
    ANM_CELT_F_EC_ERROR = 512, /* the range decoder flagged an impossible symbol (ec_get_error) */
    ANM_CELT_F_OVERRUN = 1024  /* more bits consumed than the frame holds: OPUS_INTERNAL_ERROR in the reference */
};

/* what to decode: one CELT frame of a stream.  Frames of a stream must be listed in stream order, streams back to back. */
typedef struct anm_celt_job {
    uint32_t offset;   /* first byte of the frame in the byte arena (after the TOC / length bytes: anm_opus_packet_t.size[]) */
    uint32_t len;      /* bytes of the frame */
    uint8_t channels;  /* 1 or 2: the packet's stereo flag (opus_packet_get_nb_channels) */
    uint8_t lm;        /* log2(frame samples at 48 kHz / 120): 0..3 for 2.5, 5, 10, 20 ms */
    uint8_t end_band;  /* bands coded at the packet's bandwidth: 13 NB, 17 WB, 19 SWB, 21 FB (opus_decoder.c:462-481) */
    uint8_t flags;     /* ANM_CELT_JOB_* */
} anm_celt_job_t;
enum { ANM_CELT_JOB_DISABLE_INV = 1 }; /* the stream is decoded to ONE output channel: no phase inversion of the side (celt_decoder.c:208, st->disable_inv) */

/* Host: expands the parse records of n packets (record i of anm_pb_deframe_* / anm_opus_parse_*, Fs = 48000) into CELT frame jobs -- the frames of a
 * packet back to back, offset = audio_offset + payload_offset + the sizes of the packet's earlier frames, lm from samples_per_frame, end_band from the
 * bandwidth (opus_decoder.c:473-488), `flags` copied into every job.  Packets that are not CELT-only, or that the parse rejected (count <= 0), produce no
 * job.  first_job (optional, [n]): index of packet i's first job, UINT32_MAX for a skipped packet.  Returns the number of jobs (never more than cap are
 * written; a return value above cap says how many were needed), or ANM_ERR_ARG. */
long anm_celt_jobs_from_packets(const anm_pb_span_t *spans, const anm_opus_packet_t *packets, size_t n, anm_celt_job_t *jobs, size_t cap, uint32_t flags,
                                uint32_t *first_job);

typedef struct anm_celt_frame {
    uint32_t final_range;   /* the range coder's rng after the frame = OPUS_GET_FINAL_RANGE (0 for a lost frame) */
    int32_t tell_bits;      /* ec_tell at the end of the frame */
    uint32_t flags;         /* ANM_CELT_F_* */
    uint16_t pf_pitch;      /* post-filter period (celt_decoder.c:978) */
    uint8_t pf_gain_q, pf_tapset;
    uint8_t spread, alloc_trim, intensity, coded_bands, lm, channels, pad[2]; /* pad[0]: the frame's end band */
    uint32_t pvq_codewords; /* partitions that carried a PVQ codeword ... */
    uint32_t pvq_pulses;    /* ... their pulses in total ... */
    uint32_t pvq_index_xor; /* ... and a checksum of the codeword indices */
    int8_t tf_res[ANM_CELT_BANDS];
    uint8_t fine_quant[ANM_CELT_BANDS];
    int16_t pulses[ANM_CELT_BANDS];     /* PVQ bit budget per band, 1/8 bit */
    int16_t band_e[2 * ANM_CELT_BANDS]; /* band log-energies after the frame, Q10 (oldBandE) */
} anm_celt_frame_t;

/* per-stream state carried between calls (zero-initialised = a fresh decoder): the band energies the next frame predicts from, the two
 * log-energy histories anti-collapse looks at (oldLogE, oldLogE2 of the reference's CELTDecoder) and the noise seed (its rng = the final
 * range of the stream's previous frame) */
typedef struct anm_celt_stream {
    int16_t old_e[2 * ANM_CELT_BANDS];
    int16_t log_e1[2 * ANM_CELT_BANDS], log_e2[2 * ANM_CELT_BANDS];
    uint32_t rng;
    uint32_t flags; /* bit 0: log_e1 / log_e2 are valid (clear: both are -28 dB, the decoder's reset value) */
} anm_celt_stream_t;  /* 260 bytes */

/* ---- CELT synthesis (row f1, stage 3): normalised spectrum -> PCM ------------------------------------------------- */
/* static data of the synthesis, built by anm_celt_synth_tables_build() from the standard's formulas (and checked against the reference's static
 * tables): the 120-sample window, the MDCT twiddles of the 1920 / 960 / 480 / 240-point transforms back to back, the FFT twiddles of the 480-point
 * transform (the shorter ones use every 2nd / 4th / 8th), the bit-reversal permutations of the 480 / 240 / 120 / 60-point FFTs, the band mean energies */
typedef struct anm_celt_synth_tables {
    int16_t window[120];
    int16_t trig[1800];
    int16_t fft_tw[2 * 480]; /* (re, im) */
    int16_t bitrev[480 + 240 + 120 + 60];
    int8_t e_means[25];
    int8_t pad[3];
} anm_celt_synth_tables_t;
int anm_celt_synth_tables_build(anm_celt_synth_tables_t *out); /* host; deterministic */

/* per-stream synthesis state carried between calls (zero-initialised = a fresh decoder): the output history of both channels (overlap-add tail and
 * the post-filter's memory), the de-emphasis memory and the post-filter parameters of the last two frames */
typedef struct anm_celt_synth {
    int32_t mem[2][2048 + 120];
    int32_t preemph_mem[2];
    int32_t pf_period, pf_period_old, pf_tapset, pf_tapset_old;
    int16_t pf_gain, pf_gain_old;
    uint32_t out_channels; /* channels of the decoder (1 or 2); 0: the channel count of the stream's first frame in the call */
} anm_celt_synth_t;

/* Opaque device-side context: the tables in HBM and the per-call scratch arrays of the stages (grown on demand).  A context serves ONE call at a
 * time (one stream, one host thread); use one context per concurrent stream. */
typedef struct anm_celt_ctx anm_celt_ctx_t;
int anm_celt_ctx_create(int device, anm_celt_ctx_t **out);
void anm_celt_ctx_destroy(anm_celt_ctx_t *c);
/* stream s owns jobs[stream_begin[s] .. stream_begin[s + 1]) (n_jobs = stream_begin[n_streams]), in stream order; d_* in device memory;
 * d_streams is read and written (zero-initialised for a fresh stream); bytes_mask as for anm_pb_deframe_device; stream is a cudaStream_t.
 * Two launches: every frame's symbols in parallel (k_celt_entropy), then the per-stream energy recurrence (k_celt_energies). */
int anm_celt_entropy_device(anm_celt_ctx_t *c, const anm_celt_job_t *d_jobs, const uint32_t *d_stream_begin, uint32_t n_streams, uint32_t n_jobs,
                            const uint8_t *d_bytes, uint32_t bytes_mask, anm_celt_stream_t *d_streams, anm_celt_frame_t *d_out, void *stream);
/* Stage 2 as well: the frames' normalised spectra, as celt_synthesis() receives them (quant_all_bands + anti_collapse, celt/celt_decoder.c:
 * 1084-1098).  Three launches: the two of anm_celt_entropy_device (d_out is filled as there), then one WARP per frame decodes the frame again
 * WITH its spectrum -- the noise seed of a frame is the final range of the stream's previous frame, the anti-collapse histories come from the
 * per-stream pass.  d_x: n_jobs x x_stride int16 (celt_norm, Q14), frame j's channel c at d_x + j * x_stride + c * (120 << lm), x_stride >=
 * channels * (120 << lm) (1920 always fits); coefficients above the frame's end band are written as zero.  d_collapse (may be NULL): 42 collapse masks per frame, [band * channels + channel].  Frames of <= 1 byte are
 * lost frames: the reference conceals them (celt_decode_lost), which is NOT built -- their spectrum is not written and the stream's noise seed
 * and histories pass through unchanged. */
int anm_celt_spectrum_device(anm_celt_ctx_t *c, const anm_celt_job_t *d_jobs, const uint32_t *d_stream_begin, uint32_t n_streams, uint32_t n_jobs,
                             const uint8_t *d_bytes, uint32_t bytes_mask, anm_celt_stream_t *d_streams, anm_celt_frame_t *d_out, int16_t *d_x,
                             uint32_t x_stride, uint8_t *d_collapse, void *stream);
int anm_celt_spectrum_host(const anm_celt_job_t *jobs, const uint32_t *stream_begin, uint32_t n_streams, const uint8_t *bytes, size_t n_bytes,
                           anm_celt_stream_t *streams, anm_celt_frame_t *out, int16_t *x, uint32_t x_stride, uint8_t *collapse);
/* All three stages: CELT frames -> PCM, sample for sample what the reference's celt_decode_with_ec() writes (fixed-point build; no packet-loss
 * concealment: frames of <= 1 byte produce no PCM and leave the stream's state untouched).  As anm_celt_spectrum_device, then k_celt_blocks (one warp
 * per frame: denormalisation and the inverse MDCT blocks of every output channel), k_celt_overlap (one warp per stream and output channel: window
 * overlap-add and pitch post-filter -- the per-stream recurrences that go over the lanes) and k_celt_deemphasis (one thread per stream and output
 * channel: the one-pole recurrence to 16-bit PCM).  Frames of one stream in one call: at most the caller's memory allows.  d_synth: one
 * anm_celt_synth_t per stream; d_pcm: frame j's (120 << lm) x out_channels interleaved int16 samples at d_pcm + j * pcm_stride (pcm_stride >= 1920
 * always fits; rows that start on 16-byte boundaries -- d_pcm aligned, pcm_stride a multiple of 8 -- are written 16 bytes at a time). */
int anm_celt_decode_device(anm_celt_ctx_t *c, const anm_celt_job_t *d_jobs, const uint32_t *d_stream_begin, uint32_t n_streams, uint32_t n_jobs,
                           const uint8_t *d_bytes, uint32_t bytes_mask, anm_celt_stream_t *d_streams, anm_celt_synth_t *d_synth, anm_celt_frame_t *d_out,
                           int16_t *d_pcm, uint32_t pcm_stride, void *stream);
int anm_celt_decode_host(const anm_celt_job_t *jobs, const uint32_t *stream_begin, uint32_t n_streams, const uint8_t *bytes, size_t n_bytes,
                         anm_celt_stream_t *streams, anm_celt_synth_t *synth, anm_celt_frame_t *out, int16_t *pcm, uint32_t pcm_stride);
/* host arrays: copies in, runs the kernel, copies out (no CPU fallback: ANM_ERR_CUDA without a device) */
int anm_celt_entropy_host(const anm_celt_job_t *jobs, const uint32_t *stream_begin, uint32_t n_streams, const uint8_t *bytes, size_t n_bytes,
                          anm_celt_stream_t *streams, anm_celt_frame_t *out);

#ifdef __cplusplus
}
#endif
#endif
