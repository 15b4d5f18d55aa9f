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
 * The frame decoder itself (SILK / CELT) is NOT part of this library yet.
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

#ifdef __cplusplus
}
#endif
#endif
