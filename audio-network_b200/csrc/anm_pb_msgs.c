/*
 * anm_pb_msgs.c -- discovery / handshake message codec of ip.proto, host C.
 *
 * The messages the reference exchanges around the audio stream, used here as known-answer frame
 * payloads for the modem and by a host service that fronts many receivers:
 *   BroadcastMessage{magic_word, oneof{discovery_request, discovery_response}}   protocol/ip.proto:9-27
 *   ToTransmitter{oneof{receiver_information, error}}                            protocol/ip.proto:41-64
 * built by the firmware at hardware/src/network.cpp:356-378 (discovery response) and :389-394 (hello
 * message), decoded by it at network.cpp:475 (pb_decode_delimited(BroadcastMessage_fields)).
 *
 * The decoders follow, field for field, what the reference's nanopb 0.4.5 does for these static
 * descriptors (hardware/src/protogen/ip.pb.h:114-160):
 *   pb_decode_ex / PB_DECODE_DELIMITED     hardware/lib/nanopb/src/pb_decode.c:1142-1168
 *   pb_decode_inner                        :978-1140   (zero tag, unknown fields skipped, required bitmap)
 *   pb_decode_varint32_eof / pb_decode_varint   :170-260   pb_skip_field :263-315
 *   decode_basic_field "wrong wire type"   :393-462    oneof memset on switch :519-546
 *   pb_dec_varint "integer too large"      :1406-1476  pb_dec_string "string overflow" :1518-1566
 *   pb_dec_submessage (PB_DECODE_NOINIT: a repeated static submessage merges) :1568-1618
 * Written from that behaviour, not copied; tests/test_pb_msgs.py compares every accept / reject
 * verdict and every decoded field with the reference's nanopb (oracle/_ref) on valid, truncated and
 * mutated messages.  The encoders emit the byte sequences pb_encode_delimited produces (fields in tag
 * order, minimal varints).
 */
#include "../../include/anmodem_pb.h"
#include "anm_pb_wire.h"

#include <string.h>

/* wire primitives: anm_pb_wire.h (shared with the GPU deframer and the host scanner) */
typedef anm_wstream_t rs_t;
#define varint32 anm_w_varint32
#define varint64 anm_w_varint64
#define skip_field anm_w_skip_field
#define substream anm_w_substream

static bool next_tag(rs_t *s, uint32_t *tag, uint32_t *wt, bool *done) {
    uint32_t t;
    bool eof = false;
    *done = false;
    if (!varint32(s, &t, &eof)) {
        *done = eof;
        return eof;
    }
    *tag = t >> 3;
    *wt = t & 7u;
    return *tag != 0; /* "zero tag" */
}

static bool dec_u32(rs_t *s, uint32_t wt, uint32_t *out) {
    uint64_t v;
    if (wt != 0) return false; /* "wrong wire type" */
    if (!varint64(s, &v)) return false;
    *out = (uint32_t)v;
    return (uint64_t)*out == v; /* "integer too large" */
}

static bool dec_bool(rs_t *s, uint32_t wt, uint8_t *out) {
    uint32_t v;
    if (wt != 0) return false;
    if (!varint32(s, &v, NULL)) return false;
    *out = v != 0;
    return true;
}

/* pb_dec_string into a char[128]: the terminator is written before the bytes are read */
static bool dec_string(rs_t *s, uint32_t wt, char *dst) {
    uint32_t size;
    if (wt != 2) return false;
    if (!varint32(s, &size, NULL)) return false;
    if (size == 0xFFFFFFFFu) return false;       /* "size too large" */
    if ((size_t)size + 1 > 128) return false;    /* "string overflow" */
    dst[size] = 0;
    if (s->left < size) return false;
    memcpy(dst, s->bytes + s->pos, size); /* linear host buffer (mask = 0xFFFFFFFF) */
    return anm_w_skip(s, size);
}

/* DiscoveryResponse: five required fields; decoded on top of *d (PB_DECODE_NOINIT) */
static bool dec_discovery(rs_t *s, anm_pb_discovery_t *d) {
    uint32_t seen = 0, tag, wt;
    bool done;
    while (s->left) {
        if (!next_tag(s, &tag, &wt, &done)) return false;
        if (done) break;
        bool ok;
        switch (tag) {
        case 1: ok = dec_u32(s, wt, &d->protocol_version); break;
        case 2: ok = wt == 0 && varint64(s, &d->mac_address); break;
        case 3: ok = dec_string(s, wt, d->device_name); break;
        case 4: ok = dec_bool(s, wt, &d->currently_streaming); break;
        case 5: ok = dec_string(s, wt, d->opus_version); break;
        default:
            if (!skip_field(s, wt)) return false;
            continue;
        }
        if (!ok) return false;
        seen |= 1u << (tag - 1);
    }
    return seen == 0x1Fu; /* "missing required field" */
}

int anm_pb_decode_broadcast(const uint8_t *buf, size_t len, anm_pb_broadcast_t *out, size_t *consumed) {
    if (!buf || !out || len > 0xFFFFFFFFu) return ANM_ERR_ARG;
    rs_t top = {buf, 0xFFFFFFFFu, 0u, (uint32_t)len}, s;
    anm_pb_broadcast_t m;
    memset(&m, 0, sizeof m);
    if (!substream(&top, &s)) return ANM_ERR_FORMAT;
    bool have_magic = false, done;
    uint32_t tag, wt;
    while (s.left) {
        if (!next_tag(&s, &tag, &wt, &done)) return ANM_ERR_FORMAT;
        if (done) break;
        if (tag == 1) {
            if (!dec_u32(&s, wt, &m.magic_word)) return ANM_ERR_FORMAT;
            have_magic = true;
        } else if (tag == 2) {
            m.which = 2;
            if (!dec_bool(&s, wt, &m.discovery_request)) return ANM_ERR_FORMAT;
        } else if (tag == 3) {
            rs_t sub;
            if (m.which != 3) memset(&m.discovery_response, 0, sizeof m.discovery_response); /* oneof switch */
            m.which = 3;
            if (wt != 2 || !substream(&s, &sub)) return ANM_ERR_FORMAT;
            if (!dec_discovery(&sub, &m.discovery_response)) return ANM_ERR_FORMAT;
        } else if (!skip_field(&s, wt)) {
            return ANM_ERR_FORMAT;
        }
    }
    if (!have_magic) return ANM_ERR_FORMAT;
    /* what a reader of the union sees: only the active member */
    if (m.which != 2) m.discovery_request = 0;
    if (m.which != 3) memset(&m.discovery_response, 0, sizeof m.discovery_response);
    *out = m;
    if (consumed) *consumed = len - top.left;
    return ANM_OK;
}

int anm_pb_decode_to_transmitter(const uint8_t *buf, size_t len, anm_pb_to_transmitter_t *out, size_t *consumed) {
    if (!buf || !out || len > 0xFFFFFFFFu) return ANM_ERR_ARG;
    rs_t top = {buf, 0xFFFFFFFFu, 0u, (uint32_t)len}, s;
    anm_pb_to_transmitter_t m;
    memset(&m, 0, sizeof m);
    if (!substream(&top, &s)) return ANM_ERR_FORMAT;
    bool done;
    uint32_t tag, wt;
    while (s.left) {
        if (!next_tag(&s, &tag, &wt, &done)) return ANM_ERR_FORMAT;
        if (done) break;
        if (tag == 1) { /* ReceiverInformation */
            rs_t sub;
            if (m.which != 1) {
                memset(&m, 0, sizeof m);
                m.which = 1;
            }
            if (wt != 2 || !substream(&s, &sub)) return ANM_ERR_FORMAT;
            uint32_t seen = 0, t2, w2;
            while (sub.left) {
                if (!next_tag(&sub, &t2, &w2, &done)) return ANM_ERR_FORMAT;
                if (done) break;
                if (t2 == 1) {
                    rs_t dd;
                    if (w2 != 2 || !substream(&sub, &dd)) return ANM_ERR_FORMAT;
                    if (!dec_discovery(&dd, &m.discovery_data)) return ANM_ERR_FORMAT;
                } else if (t2 == 2) {
                    if (!dec_u32(&sub, w2, &m.max_encoded_frame_size)) return ANM_ERR_FORMAT;
                } else if (t2 == 3) {
                    if (!dec_u32(&sub, w2, &m.max_decoded_frame_size)) return ANM_ERR_FORMAT;
                } else {
                    if (!skip_field(&sub, w2)) return ANM_ERR_FORMAT;
                    continue;
                }
                seen |= 1u << (t2 - 1);
            }
            if (seen != 7u) return ANM_ERR_FORMAT;
        } else if (tag == 2) { /* ReceiverError */
            rs_t sub;
            if (m.which != 2) {
                memset(&m, 0, sizeof m);
                m.which = 2;
            }
            if (wt != 2 || !substream(&s, &sub)) return ANM_ERR_FORMAT;
            uint32_t seen = 0, t2, w2;
            while (sub.left) {
                if (!next_tag(&sub, &t2, &w2, &done)) return ANM_ERR_FORMAT;
                if (done) break;
                if (t2 == 1) {
                    if (!dec_bool(&sub, w2, &m.audio_underflow)) return ANM_ERR_FORMAT;
                } else if (t2 == 2) {
                    if (!dec_bool(&sub, w2, &m.audio_decode_error)) return ANM_ERR_FORMAT;
                } else {
                    if (!skip_field(&sub, w2)) return ANM_ERR_FORMAT;
                    continue;
                }
                seen |= 1u << (t2 - 1);
            }
            if (seen != 3u) return ANM_ERR_FORMAT;
        } else if (!skip_field(&s, wt)) {
            return ANM_ERR_FORMAT;
        }
    }
    *out = m;
    if (consumed) *consumed = len - top.left;
    return ANM_OK;
}

/* ---- encoders: what pb_encode_delimited writes for these structs ------------------------------- */
typedef struct {
    uint8_t *p;
    size_t n, cap;
} ws_t;

static void put(ws_t *w, const void *src, size_t n) {
    if (w->p && w->n + n <= w->cap) memcpy(w->p + w->n, src, n);
    w->n += n; /* keeps counting past cap: the caller compares with cap */
}
static void put_varint(ws_t *w, uint64_t v) {
    uint8_t tmp[10];
    put(w, tmp, anm_pb_varint(v, tmp));
}
static void put_tag(ws_t *w, uint32_t field, uint32_t wt) { put_varint(w, (field << 3) | wt); }

static size_t cstr_len(const char *s) { /* a char[128] field: at most 127 characters are encoded */
    size_t n = 0;
    while (n < 127 && s[n]) ++n;
    return n;
}

static void enc_discovery(ws_t *w, const anm_pb_discovery_t *d) {
    put_tag(w, 1, 0);
    put_varint(w, d->protocol_version);
    put_tag(w, 2, 0);
    put_varint(w, d->mac_address);
    put_tag(w, 3, 2);
    put_varint(w, cstr_len(d->device_name));
    put(w, d->device_name, cstr_len(d->device_name));
    put_tag(w, 4, 0);
    put_varint(w, d->currently_streaming ? 1 : 0);
    put_tag(w, 5, 2);
    put_varint(w, cstr_len(d->opus_version));
    put(w, d->opus_version, cstr_len(d->opus_version));
}
static size_t discovery_size(const anm_pb_discovery_t *d) {
    ws_t c = {NULL, 0, 0};
    enc_discovery(&c, d);
    return c.n;
}

static void enc_broadcast_body(ws_t *w, const anm_pb_broadcast_t *m) {
    put_tag(w, 1, 0);
    put_varint(w, m->magic_word);
    if (m->which == 2) {
        put_tag(w, 2, 0);
        put_varint(w, m->discovery_request ? 1 : 0);
    } else if (m->which == 3) {
        put_tag(w, 3, 2);
        put_varint(w, discovery_size(&m->discovery_response));
        enc_discovery(w, &m->discovery_response);
    }
}

size_t anm_pb_encode_broadcast(const anm_pb_broadcast_t *m, uint8_t *out, size_t cap) {
    if (!m || !out || (m->which != 0 && m->which != 2 && m->which != 3)) return 0;
    ws_t c = {NULL, 0, 0};
    enc_broadcast_body(&c, m);
    ws_t w = {out, 0, cap};
    put_varint(&w, c.n);
    enc_broadcast_body(&w, m);
    return w.n <= cap ? w.n : 0;
}

static void enc_to_transmitter_body(ws_t *w, const anm_pb_to_transmitter_t *m) {
    if (m->which == 1) {
        const size_t dd = discovery_size(&m->discovery_data);
        ws_t c = {NULL, 0, 0};
        put_tag(&c, 1, 2);
        put_varint(&c, dd);
        c.n += dd;
        put_tag(&c, 2, 0);
        put_varint(&c, m->max_encoded_frame_size);
        put_tag(&c, 3, 0);
        put_varint(&c, m->max_decoded_frame_size);
        put_tag(w, 1, 2);
        put_varint(w, c.n);
        put_tag(w, 1, 2);
        put_varint(w, dd);
        enc_discovery(w, &m->discovery_data);
        put_tag(w, 2, 0);
        put_varint(w, m->max_encoded_frame_size);
        put_tag(w, 3, 0);
        put_varint(w, m->max_decoded_frame_size);
    } else if (m->which == 2) {
        put_tag(w, 2, 2);
        put_varint(w, 4);
        put_tag(w, 1, 0);
        put_varint(w, m->audio_underflow ? 1 : 0);
        put_tag(w, 2, 0);
        put_varint(w, m->audio_decode_error ? 1 : 0);
    }
}

size_t anm_pb_encode_to_transmitter(const anm_pb_to_transmitter_t *m, uint8_t *out, size_t cap) {
    if (!m || !out || m->which > 2) return 0;
    ws_t c = {NULL, 0, 0};
    enc_to_transmitter_body(&c, m);
    ws_t w = {out, 0, cap};
    put_varint(&w, c.n);
    enc_to_transmitter_body(&w, m);
    return w.n <= cap ? w.n : 0;
}

/* the firmware's discovery response and hello message with its constants (network.cpp:356-378, :389-394) */
void anm_pb_firmware_discovery(uint64_t mac, const char *opus_version, anm_pb_broadcast_t *out) {
    memset(out, 0, sizeof *out);
    out->magic_word = ANM_PB_MAGIC_WORD;
    out->which = 3;
    out->discovery_response.protocol_version = 1;
    out->discovery_response.mac_address = mac;
    out->discovery_response.currently_streaming = 0;
    if (opus_version) strncpy(out->discovery_response.opus_version, opus_version, 127);
}
