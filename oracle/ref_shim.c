/*
 * ref_shim.c -- host glue linked with the REFERENCE's nanopb 0.4.5 and generated
 * ip.pb.c (compiled in place from /root/reference by oracle/Makefile) to form
 * oracle/_ref/libref_nanopb.so.  TEST INFRASTRUCTURE ONLY.
 *
 * ip.pb.c needs the external field callback named in hardware/protobuf_ip.options:3
 * and declared at hardware/src/protogen/ip.pb.h:161-162.  The reference defines it in
 * hardware/src/network.cpp:220-249, a file that cannot be built on a host (Arduino,
 * FreeRTOS, lwIP headers).  The definition below is this repository's own restatement
 * of that callback's decode behaviour (reject > 4096 bytes, allocate, pb_read the whole
 * field, stash pointer+length in opus_encoded_frame.arg); unlike the reference it also
 * implements the encode direction so the shim can build payloads.
 *
 * The flat ref_* functions are what tests call through ctypes:
 *   - ref_encode_*: build varint-delimited messages with pb_encode_ex(PB_ENCODE_DELIMITED)
 *   - ref_decode_*: pb_decode_ex(PB_DECODE_DELIMITED), the call at network.cpp:411/475
 */
#include <pb_decode.h>
#include <pb_encode.h>
#include <stdlib.h>
#include <string.h>

#include "ip.pb.h"

#define MAX_ENCODED_FRAME_SIZE 4096 /* hardware/src/network.cpp:24 */

typedef struct {
    void *data;
    size_t len;
} bytes_ctx;

bool network_pb_callback_audio_data(pb_istream_t *istream, pb_ostream_t *ostream, const pb_field_t *field) {
    if (field->tag != AudioData_opus_encoded_frame_tag) return pb_default_field_callback(istream, ostream, field);
    AudioData *msg = (AudioData *)field->message;
    if (ostream != NULL) { /* encode direction (not in the reference) */
        const bytes_ctx *c = (const bytes_ctx *)msg->opus_encoded_frame.arg;
        if (!c) return true;
        return pb_encode_tag_for_field(ostream, field) && pb_encode_string(ostream, (const pb_byte_t *)c->data, c->len);
    }
    if (istream->bytes_left > MAX_ENCODED_FRAME_SIZE) {
        istream->errmsg = "Encoded frame exceeds max size";
        return false;
    }
    bytes_ctx *c = (bytes_ctx *)malloc(sizeof *c);
    if (!c) return false;
    c->len = istream->bytes_left;
    c->data = malloc(c->len ? c->len : 1);
    if (!c->data || !pb_read(istream, (pb_byte_t *)c->data, c->len)) {
        free(c->data);
        free(c);
        return false;
    }
    msg->opus_encoded_frame.arg = c;
    return true;
}

/* ---- encoders: return encoded length, 0 on failure ------------------------ */
size_t ref_encode_to_receiver_audio(const uint8_t *opus, size_t len, uint8_t *out, size_t cap) {
    ToReceiver m = ToReceiver_init_zero;
    bytes_ctx c = {(void *)opus, len};
    m.which_message = ToReceiver_audio_data_tag;
    m.message.audio_data.opus_encoded_frame.arg = &c;
    pb_ostream_t os = pb_ostream_from_buffer(out, cap);
    return pb_encode_ex(&os, ToReceiver_fields, &m, PB_ENCODE_DELIMITED) ? os.bytes_written : 0;
}

size_t ref_encode_broadcast_request(uint32_t magic, uint8_t *out, size_t cap) {
    BroadcastMessage m = BroadcastMessage_init_zero;
    m.magic_word = magic;
    m.which_message = BroadcastMessage_discovery_request_tag;
    m.message.discovery_request = true;
    pb_ostream_t os = pb_ostream_from_buffer(out, cap);
    return pb_encode_ex(&os, BroadcastMessage_fields, &m, PB_ENCODE_DELIMITED) ? os.bytes_written : 0;
}

static void fill_discovery(DiscoveryResponse *d, uint32_t version, uint64_t mac, const char *name,
                           int streaming, const char *opus) {
    d->protocol_version = version;
    d->mac_address = mac;
    strncpy(d->device_name, name, sizeof d->device_name - 1);
    d->currently_streaming = streaming != 0;
    strncpy(d->opus_version, opus, sizeof d->opus_version - 1);
}

size_t ref_encode_broadcast_response(uint32_t magic, uint32_t version, uint64_t mac, const char *name,
                                     int streaming, const char *opus, uint8_t *out, size_t cap) {
    BroadcastMessage m = BroadcastMessage_init_zero;
    m.magic_word = magic;
    m.which_message = BroadcastMessage_discovery_response_tag;
    fill_discovery(&m.message.discovery_response, version, mac, name, streaming, opus);
    pb_ostream_t os = pb_ostream_from_buffer(out, cap);
    return pb_encode_ex(&os, BroadcastMessage_fields, &m, PB_ENCODE_DELIMITED) ? os.bytes_written : 0;
}

size_t ref_encode_to_transmitter_info(uint32_t version, uint64_t mac, const char *name, int streaming,
                                      const char *opus, uint32_t max_enc, uint32_t max_dec, uint8_t *out,
                                      size_t cap) {
    ToTransmitter m = ToTransmitter_init_zero;
    m.which_message = ToTransmitter_receiver_information_tag;
    fill_discovery(&m.message.receiver_information.discovery_data, version, mac, name, streaming, opus);
    m.message.receiver_information.max_encoded_frame_size = max_enc;
    m.message.receiver_information.max_decoded_frame_size = max_dec;
    pb_ostream_t os = pb_ostream_from_buffer(out, cap);
    return pb_encode_ex(&os, ToTransmitter_fields, &m, PB_ENCODE_DELIMITED) ? os.bytes_written : 0;
}

size_t ref_encode_to_transmitter_error(int underflow, int decode_error, uint8_t *out, size_t cap) {
    ToTransmitter m = ToTransmitter_init_zero;
    m.which_message = ToTransmitter_error_tag;
    m.message.error.audio_underflow = underflow != 0;
    m.message.error.audio_decode_error = decode_error != 0;
    pb_ostream_t os = pb_ostream_from_buffer(out, cap);
    return pb_encode_ex(&os, ToTransmitter_fields, &m, PB_ENCODE_DELIMITED) ? os.bytes_written : 0;
}

/* ---- decoders -------------------------------------------------------------- */
/* returns payload length (>= 0) copied to out, -1 on decode failure, -2 wrong oneof */
long ref_decode_to_receiver_audio(const uint8_t *buf, size_t len, uint8_t *out, size_t cap, size_t *consumed) {
    pb_istream_t is = pb_istream_from_buffer(buf, len);
    ToReceiver m = ToReceiver_init_zero;
    if (!pb_decode_ex(&is, ToReceiver_fields, &m, PB_DECODE_DELIMITED)) return -1;
    if (consumed) *consumed = len - is.bytes_left;
    if (m.which_message != ToReceiver_audio_data_tag) return -2;
    bytes_ctx *c = (bytes_ctx *)m.message.audio_data.opus_encoded_frame.arg;
    if (!c) return -1;
    long n = (long)c->len;
    if (c->len <= cap) memcpy(out, c->data, c->len); else n = -1;
    free(c->data);
    free(c);
    return n;
}

typedef struct {
    uint32_t magic, which, discovery_request, protocol_version;
    uint64_t mac;
    uint32_t streaming, pad;
    char device_name[128];
    char opus_version[128];
} ref_broadcast_t;

int ref_decode_broadcast(const uint8_t *buf, size_t len, ref_broadcast_t *out) {
    pb_istream_t is = pb_istream_from_buffer(buf, len);
    BroadcastMessage m = BroadcastMessage_init_zero;
    if (!pb_decode_ex(&is, BroadcastMessage_fields, &m, PB_DECODE_DELIMITED)) return -1;
    memset(out, 0, sizeof *out);
    out->magic = m.magic_word;
    out->which = m.which_message;
    if (m.which_message == BroadcastMessage_discovery_request_tag) out->discovery_request = m.message.discovery_request;
    if (m.which_message == BroadcastMessage_discovery_response_tag) {
        const DiscoveryResponse *d = &m.message.discovery_response;
        out->protocol_version = d->protocol_version;
        out->mac = d->mac_address;
        out->streaming = d->currently_streaming;
        memcpy(out->device_name, d->device_name, 128);
        memcpy(out->opus_version, d->opus_version, 128);
    }
    return 0;
}

typedef struct {
    uint32_t which, protocol_version;
    uint64_t mac;
    uint32_t streaming, max_enc, max_dec, underflow, decode_error, pad;
    char device_name[128];
    char opus_version[128];
} ref_to_transmitter_t;

int ref_decode_to_transmitter(const uint8_t *buf, size_t len, ref_to_transmitter_t *out) {
    pb_istream_t is = pb_istream_from_buffer(buf, len);
    ToTransmitter m = ToTransmitter_init_zero;
    if (!pb_decode_ex(&is, ToTransmitter_fields, &m, PB_DECODE_DELIMITED)) return -1;
    memset(out, 0, sizeof *out);
    out->which = m.which_message;
    if (m.which_message == ToTransmitter_receiver_information_tag) {
        const ReceiverInformation *r = &m.message.receiver_information;
        out->protocol_version = r->discovery_data.protocol_version;
        out->mac = r->discovery_data.mac_address;
        out->streaming = r->discovery_data.currently_streaming;
        memcpy(out->device_name, r->discovery_data.device_name, 128);
        memcpy(out->opus_version, r->discovery_data.opus_version, 128);
        out->max_enc = r->max_encoded_frame_size;
        out->max_dec = r->max_decoded_frame_size;
    } else if (m.which_message == ToTransmitter_error_tag) {
        out->underflow = m.message.error.audio_underflow;
        out->decode_error = m.message.error.audio_decode_error;
    }
    return 0;
}

/* Decode through a caller-supplied pb_istream_t-compatible callback: this is the drop-in
 * seam of hardware/src/network.cpp:262-305 -- the product's byte source (anm_pb.c) is
 * handed to the reference decoder exactly where the socket stream is today. */
typedef bool (*ref_stream_cb)(pb_istream_t *stream, pb_byte_t *buf, size_t count);
long ref_decode_to_receiver_from_stream(ref_stream_cb cb, void *state, uint8_t *out, size_t cap) {
    pb_istream_t is = {cb, state, SIZE_MAX, NULL};
    ToReceiver m = ToReceiver_init_zero;
    if (!pb_decode_ex(&is, ToReceiver_fields, &m, PB_DECODE_DELIMITED)) return -1;
    if (m.which_message != ToReceiver_audio_data_tag) return -2;
    bytes_ctx *c = (bytes_ctx *)m.message.audio_data.opus_encoded_frame.arg;
    if (!c) return -1;
    long n = (long)c->len;
    if (c->len <= cap) memcpy(out, c->data, c->len); else n = -1;
    free(c->data);
    free(c);
    return n;
}
