/*
 * anm_pb.c -- byte source for the reference's nanopb decoder and minimal ip.proto wire
 * helpers (include/anmodem_pb.h).  Host C.
 *
 * Follows the wire behaviour of the reference's decoder for the messages this path carries:
 * varint (hardware/lib/nanopb/src/pb_decode.c:170-232), tag split (pb_decode.c:288-303),
 * length-delimited substreams (pb_decode.c:359-387) and the delimited wrapper
 * (pb_decode.c:1142-1168); message shapes from protocol/ip.proto:9-64.  The decode-side primitives live in
 * anm_pb_wire.h (transcriptions of nanopb's rules, see the notice there); the encoders are written from the
 * protobuf wire format.
 */
#include "../../include/anmodem_pb.h"
#include "anm_pb_wire.h"

#include <stdlib.h>
#include <string.h>

struct anm_pb_queue {
    uint8_t *buf;
    size_t head, tail, cap;
};

anm_pb_queue_t *anm_pb_queue_create(void) { return (anm_pb_queue_t *)calloc(1, sizeof(anm_pb_queue_t)); }

void anm_pb_queue_destroy(anm_pb_queue_t *q) {
    if (!q) return;
    free(q->buf);
    free(q);
}

int anm_pb_queue_push(anm_pb_queue_t *q, const uint8_t *bytes, size_t len) {
    if (!q || (!bytes && len)) return ANM_ERR_ARG;
    if (q->head == q->tail) q->head = q->tail = 0;
    if (q->tail + len > q->cap) {
        /* compact, then grow */
        memmove(q->buf, q->buf + q->head, q->tail - q->head);
        q->tail -= q->head;
        q->head = 0;
        if (q->tail + len > q->cap) {
            size_t ncap = (q->tail + len) * 2 + 256;
            uint8_t *nb = (uint8_t *)realloc(q->buf, ncap);
            if (!nb) return ANM_ERR_NOMEM;
            q->buf = nb;
            q->cap = ncap;
        }
    }
    memcpy(q->buf + q->tail, bytes, len);
    q->tail += len;
    return ANM_OK;
}

size_t anm_pb_queue_size(const anm_pb_queue_t *q) { return q ? q->tail - q->head : 0; }

/* pb_istream_t callback contract, pb_decode.h:20-27 */
static bool queue_read(anm_pb_istream_t *s, uint8_t *buf, size_t count) {
    anm_pb_queue_t *q = (anm_pb_queue_t *)s->state;
    if (count == 0) return true;
    if (!q || q->tail - q->head < count) {
        s->bytes_left = 0; /* same signal the socket stream gives on a closed peer, network.cpp:288-291 */
        return false;
    }
    if (buf) memcpy(buf, q->buf + q->head, count);
    q->head += count;
    return true;
}

anm_pb_istream_t anm_pb_istream_from_queue(anm_pb_queue_t *q) {
    anm_pb_istream_t s;
    s.callback = queue_read;
    s.state = q;
    s.bytes_left = (size_t)-1; /* SIZE_MAX, like network_pb_istream_from_socket (network.cpp:299-305) */
    s.errmsg = NULL;
    return s;
}

size_t anm_pb_varint(uint64_t v, uint8_t *out) {
    size_t n = 0;
    do {
        uint8_t b = (uint8_t)(v & 0x7F);
        v >>= 7;
        out[n++] = (uint8_t)(b | (v ? 0x80 : 0));
    } while (v);
    return n;
}

static size_t varint_len(uint64_t v) {
    uint8_t tmp[10];
    return anm_pb_varint(v, tmp);
}

size_t anm_pb_encode_to_receiver_audio(const uint8_t *data, size_t len, uint8_t *out, size_t cap) {
    /* ToReceiver{ 1: AudioData{ 1: bytes } } */
    size_t inner = 1 + varint_len(len) + len;       /* AudioData body */
    size_t outer = 1 + varint_len(inner) + inner;   /* ToReceiver body */
    size_t total = varint_len(outer) + outer;
    if ((!data && len) || !out || cap < total) return 0;
    size_t n = 0;
    n += anm_pb_varint(outer, out + n);
    out[n++] = (1u << 3) | 2u;
    n += anm_pb_varint(inner, out + n);
    out[n++] = (1u << 3) | 2u;
    n += anm_pb_varint(len, out + n);
    memcpy(out + n, data, len);
    return n + len;
}

size_t anm_pb_encode_broadcast_request(uint32_t magic, uint8_t *out, size_t cap) {
    /* BroadcastMessage{ 1: uint32 magic_word, 2: bool discovery_request = true } */
    size_t body = 1 + varint_len(magic) + 2;
    size_t total = varint_len(body) + body;
    if (!out || cap < total) return 0;
    size_t n = 0;
    n += anm_pb_varint(body, out + n);
    out[n++] = (1u << 3) | 0u;
    n += anm_pb_varint(magic, out + n);
    out[n++] = (2u << 3) | 0u;
    out[n++] = 1;
    return n;
}

/* Host twin of k_pb_deframe: the same walk (anm_pb_wire.h), so a payload is accepted on the host exactly when the GPU
 * deframer and the reference decoder accept it -- including "wrong wire type" on field 1, the 32-bit varint rules and the
 * callback's 4096-byte limit (hardware/src/network.cpp:223).  A message without audio_data is reported as 0 here. */
size_t anm_pb_scan_to_receiver_audio(const uint8_t *buf, size_t len, const uint8_t **payload, size_t *payload_len) {
    if (!buf || !payload || !payload_len || len > 0xFFFFFFFFu) return 0;
    anm_wstream_t s = {buf, 0xFFFFFFFFu, 0u, (uint32_t)len};
    bool have = false;
    uint32_t a_off = 0, a_len = 0;
    if (!anm_w_to_receiver(&s, &have, &a_off, &a_len) || !have) return 0;
    *payload = buf + a_off;
    *payload_len = a_len;
    return s.pos;
}
