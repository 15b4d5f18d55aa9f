/*
 * anmodem_pb.h -- the hand-off from the demodulator to the reference's protobuf decoder.
 *
 * Reference interfaces this file stands in for / plugs into:
 *   - struct pb_istream_s, hardware/lib/nanopb/src/pb_decode.h:28-46: the byte source
 *     pb_decode_delimited() pulls from (callback contract at pb_decode.h:20-27: return
 *     false on I/O error, buf == NULL means skip, state is the callee's).
 *   - its socket-backed instance, hardware/src/network.cpp:262-305, consumed in the
 *     receive loop at hardware/src/network.cpp:406-411.
 * anm_pb_istream_t is layout-compatible with pb_istream_t of nanopb 0.4.5 built WITHOUT
 * PB_BUFFER_ONLY (the reference's configuration), so a maintainer can write
 *     pb_istream_t is; memcpy(&is, &s, sizeof is);      // or cast the pointer
 *     pb_decode_delimited(&is, ToReceiver_fields, &msg);
 * exactly where network_pb_istream_from_socket() is used today.
 *
 * The small encode/scan helpers below build and walk varint-delimited ip.proto messages
 * (protocol/ip.proto:9-64) without nanopb; they exist so that tests and tools can make
 * frame payloads on a machine that does not have the reference tree.  They are checked
 * against the reference's nanopb in tests/test_pb.py.
 */
#ifndef ANMODEM_PB_H_INCLUDED
#define ANMODEM_PB_H_INCLUDED

#include <stdbool.h>
#include <stddef.h>
#include <stdint.h>

#include "anmodem.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct anm_pb_istream anm_pb_istream_t;
struct anm_pb_istream {
    bool (*callback)(anm_pb_istream_t *stream, uint8_t *buf, size_t count);
    void *state;
    size_t bytes_left;
    const char *errmsg;
};

/* A byte queue fed with CRC-valid frame payloads, drained through the stream callback. */
typedef struct anm_pb_queue anm_pb_queue_t;
anm_pb_queue_t *anm_pb_queue_create(void);
void anm_pb_queue_destroy(anm_pb_queue_t *q);
int anm_pb_queue_push(anm_pb_queue_t *q, const uint8_t *bytes, size_t len);
size_t anm_pb_queue_size(const anm_pb_queue_t *q);
/* stream over the queue: reading past the queued bytes fails (callback returns false),
 * which nanopb reports as "io error"/"end-of-stream" just like a closed socket */
anm_pb_istream_t anm_pb_istream_from_queue(anm_pb_queue_t *q);

/* Firmware-idiom: moves every CRC-valid frame decoded so far by `d` into an internal queue
 * and returns a stream over it (SURVEY.md 8(b) seam #1). */
anm_pb_istream_t demod_as_pb_istream(demod_t *d);

/* ---- minimal ip.proto wire helpers (proto2, protocol/ip.proto) ---------------------- */
size_t anm_pb_varint(uint64_t v, uint8_t *out); /* returns bytes written (<= 10) */
/* delimited ToReceiver{audio_data{opus_encoded_frame = data}}; returns length or 0 */
size_t anm_pb_encode_to_receiver_audio(const uint8_t *data, size_t len, uint8_t *out, size_t cap);
/* delimited BroadcastMessage{magic_word, discovery_request = true} */
size_t anm_pb_encode_broadcast_request(uint32_t magic, uint8_t *out, size_t cap);
/* Walks one delimited ToReceiver message: on success returns the total encoded length
 * consumed and sets *payload / *payload_len to the AudioData bytes inside buf; returns 0 on
 * malformed input (bad varint, truncated field, unknown wire type). */
size_t anm_pb_scan_to_receiver_audio(const uint8_t *buf, size_t len, const uint8_t **payload, size_t *payload_len);

/* ---- discovery / handshake messages --------------------------------
 * BroadcastMessage and ToTransmitter of protocol/ip.proto:9-64 as flat structs (the nanopb structs
 * of hardware/src/protogen/ip.pb.h:18-68 without the unions).  decode = the verdict and the fields of
 * pb_decode_delimited(&stream, BroadcastMessage_fields / ToTransmitter_fields, &msg) over a buffer
 * (hardware/src/network.cpp:475); encode = the bytes pb_encode_delimited writes (network.cpp:394). */
#define ANM_PB_MAGIC_WORD 0x2C5DA044u /* protocol/ip.proto:10, network.cpp:369 */
typedef struct anm_pb_discovery {
    uint32_t protocol_version;
    uint8_t currently_streaming;
    uint8_t pad[3];
    uint64_t mac_address;
    char device_name[128]; /* zero-terminated, at most 127 characters (protobuf_ip.options:1-2) */
    char opus_version[128];
} anm_pb_discovery_t;
typedef struct anm_pb_broadcast {
    uint32_t magic_word;
    uint32_t which;            /* 0: oneof not set, 2: discovery_request, 3: discovery_response */
    uint8_t discovery_request; /* valid when which == 2 */
    uint8_t pad[7];
    anm_pb_discovery_t discovery_response; /* valid when which == 3 */
} anm_pb_broadcast_t;
typedef struct anm_pb_to_transmitter {
    uint32_t which; /* 0: not set, 1: receiver_information, 2: error */
    uint32_t max_encoded_frame_size, max_decoded_frame_size; /* which == 1 */
    uint8_t audio_underflow, audio_decode_error;             /* which == 2 */
    uint8_t pad[2];
    anm_pb_discovery_t discovery_data; /* which == 1 */
} anm_pb_to_transmitter_t;
/* return the encoded length (length varint included), 0 if it does not fit or the struct is invalid */
size_t anm_pb_encode_broadcast(const anm_pb_broadcast_t *m, uint8_t *out, size_t cap);
size_t anm_pb_encode_to_transmitter(const anm_pb_to_transmitter_t *m, uint8_t *out, size_t cap);
/* ANM_OK, or ANM_ERR_FORMAT where pb_decode_delimited returns false; *consumed (optional) = bytes read */
int anm_pb_decode_broadcast(const uint8_t *buf, size_t len, anm_pb_broadcast_t *out, size_t *consumed);
int anm_pb_decode_to_transmitter(const uint8_t *buf, size_t len, anm_pb_to_transmitter_t *out, size_t *consumed);
/* the discovery response the firmware sends (network.cpp:356-378): protocol version 1, not streaming */
void anm_pb_firmware_discovery(uint64_t mac, const char *opus_version, anm_pb_broadcast_t *out);

/* ---- batched deframer on the GPU ----------------------------
 * The decode step of the reference's receive loop (hardware/src/network.cpp:406-430) for many
 * frames at once: every CRC-valid frame payload is walked as one varint-delimited ToReceiver
 * message with the semantics of pb_decode_delimited(&stream, ToReceiver_fields, &msg)
 * (hardware/lib/nanopb/src/pb_decode.c:1142-1168) and of the field callback
 * network_pb_callback_audio_data (hardware/src/network.cpp:212-249, 4096-byte limit at :223).
 * Instead of copying the Opus bytes into two heap blocks per frame it reports where they lie. */
enum {
    ANM_PB_OK = 0,       /* decoded, audio_data present: [audio_offset, audio_offset + audio_len) */
    ANM_PB_FAIL = 1,     /* pb_decode_delimited would return false */
    ANM_PB_NO_AUDIO = 2, /* decoded, but the oneof does not hold audio_data */
    ANM_PB_CRC = 3       /* frame failed its CRC-16: never handed to the decoder */
};
typedef struct anm_pb_span {
    uint32_t status;
    uint32_t consumed;     /* bytes of the payload the decoder consumed (length varint + message) */
    uint32_t audio_offset; /* position in the byte arena (same space as anm_frame_t.offset) */
    uint32_t audio_len;
} anm_pb_span_t;
/* frames / bytes / out in device memory; bytes_mask = arena size - 1 for a power-of-two ring,
 * 0xFFFFFFFF for a linear array; stream is a cudaStream_t (NULL = default stream) */
int anm_pb_deframe_device(const anm_frame_t *d_frames, uint32_t n_frames, const uint8_t *d_bytes,
                          uint32_t bytes_mask, anm_pb_span_t *d_out, void *stream);
/* host arrays as returned by anm_demod_read_frames(); copies in, runs the kernel, copies out */
int anm_pb_deframe_host(const anm_frame_t *frames, size_t n_frames, const uint8_t *bytes, size_t n_bytes,
                        anm_pb_span_t *out);

#ifdef __cplusplus
}
#endif
#endif
