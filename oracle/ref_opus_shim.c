/*
 * ref_opus_shim.c -- host glue linked with the REFERENCE's libopus 1.3.1 (fixed-point build, compiled in
 * place from /root/reference/hardware/lib/libopus/src by oracle/Makefile) to form
 * oracle/_ref/libref_opus.so.  TEST INFRASTRUCTURE ONLY: the oracle of SURVEY.md 8(f) row f1 (the
 * reference's real PCM path: playback.cpp:115-122 -> opus_decode).
 *
 *   ref_opus_parse          what opus_decode_native learns from a packet before it decodes a frame
 *                           (opus_decoder.c:661-669: mode, bandwidth, frame size, channels,
 *                           opus_packet_parse_impl) plus opus_packet_get_nb_frames / _nb_samples
 *   ref_opus_encode_stream  packets as the transmitter makes them (OpusEncoder.kt:51-67: application AUDIO,
 *                           92 kbit/s, complexity 10, signal AUTO, max bandwidth FULLBAND)
 *   ref_opus_decode_stream  opus_decode() over a packet sequence (playback.cpp:115-122)
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "opus.h"

typedef struct {
    int32_t count; /* opus_packet_parse: frames, or a negative OPUS_* error */
    uint8_t toc, channels, pad[2];
    int32_t mode, bandwidth, samples_per_frame, payload_offset, nb_frames, nb_samples;
    int16_t size[48];
} ref_opus_packet_t;

/* opus_packet_get_mode is static in the reference (opus_decoder.c:206-219); MODE_* from opus_private.h:124-126 */
static int packet_mode(const unsigned char *data) {
    if (data[0] & 0x80) return 1002;             /* MODE_CELT_ONLY */
    if ((data[0] & 0x60) == 0x60) return 1001;   /* MODE_HYBRID */
    return 1000;                                 /* MODE_SILK_ONLY */
}

void ref_opus_parse(const uint8_t *data, int32_t len, int32_t Fs, ref_opus_packet_t *out) {
    memset(out, 0, sizeof *out);
    int off = 0;
    unsigned char toc = 0;
    out->count = opus_packet_parse(data, len, &toc, NULL, out->size, &off);
    out->nb_frames = opus_packet_get_nb_frames(data, len);
    out->nb_samples = opus_packet_get_nb_samples(data, len, Fs);
    if (len >= 1) {
        out->toc = data[0];
        out->channels = (uint8_t)opus_packet_get_nb_channels(data);
        out->mode = packet_mode(data);
        out->bandwidth = opus_packet_get_bandwidth(data);
        out->samples_per_frame = opus_packet_get_samples_per_frame(data, Fs);
    }
    if (out->count >= 0) out->payload_offset = off;
    else memset(out->size, 0, sizeof out->size); /* sizes of a rejected packet are unspecified */
}

/* pcm: interleaved int16, n_frames * frame_samples * channels; out: n_frames slots of max_len bytes */
int ref_opus_encode_stream(const int16_t *pcm, int n_frames, int frame_samples, int channels, uint8_t *out, int32_t *lens, int max_len) {
    int err = 0;
    OpusEncoder *e = opus_encoder_create(48000, channels, OPUS_APPLICATION_AUDIO, &err);
    if (!e || err != OPUS_OK) return -1;
    opus_encoder_ctl(e, OPUS_SET_BITRATE(92000));
    opus_encoder_ctl(e, OPUS_SET_COMPLEXITY(10));
    opus_encoder_ctl(e, OPUS_SET_SIGNAL(OPUS_AUTO));
    opus_encoder_ctl(e, OPUS_SET_MAX_BANDWIDTH(OPUS_BANDWIDTH_FULLBAND));
    for (int i = 0; i < n_frames; ++i) {
        lens[i] = opus_encode(e, pcm + (size_t)i * frame_samples * channels, frame_samples, out + (size_t)i * max_len, max_len);
        if (lens[i] < 0) { opus_encoder_destroy(e); return lens[i]; }
    }
    opus_encoder_destroy(e);
    return n_frames;
}

/* CELT-only packets at any bitrate (OPUS_APPLICATION_RESTRICTED_LOWDELAY never uses SILK): test corpus for the frame decoder */
int ref_opus_encode_stream_celt(const int16_t *pcm, int n_frames, int frame_samples, int channels, int bitrate, int max_bandwidth, uint8_t *out,
                                int32_t *lens, int max_len) {
    int err = 0;
    OpusEncoder *e = opus_encoder_create(48000, channels, OPUS_APPLICATION_RESTRICTED_LOWDELAY, &err);
    if (!e || err != OPUS_OK) return -1;
    opus_encoder_ctl(e, OPUS_SET_BITRATE(bitrate));
    opus_encoder_ctl(e, OPUS_SET_COMPLEXITY(10));
    if (max_bandwidth) opus_encoder_ctl(e, OPUS_SET_MAX_BANDWIDTH(max_bandwidth));
    for (int i = 0; i < n_frames; ++i) {
        lens[i] = opus_encode(e, pcm + (size_t)i * frame_samples * channels, frame_samples, out + (size_t)i * max_len, max_len);
        if (lens[i] < 0) { opus_encoder_destroy(e); return lens[i]; }
    }
    opus_encoder_destroy(e);
    return n_frames;
}

/* returns the samples per channel decoded in total (negative OPUS_* error of the first failing packet) */
int ref_opus_decode_stream(const uint8_t *packets, const int32_t *lens, int n, int max_len, int channels, int16_t *pcm, int max_frame_samples) {
    int err = 0, total = 0;
    OpusDecoder *d = opus_decoder_create(48000, channels, &err);
    if (!d || err != OPUS_OK) return -1;
    for (int i = 0; i < n; ++i) {
        int r = opus_decode(d, packets + (size_t)i * max_len, lens[i], pcm + (size_t)total * channels, max_frame_samples, 0);
        if (r < 0) { opus_decoder_destroy(d); return r; }
        total += r;
    }
    opus_decoder_destroy(d);
    return total;
}

const char *ref_opus_version(void) { return opus_get_version_string(); }
