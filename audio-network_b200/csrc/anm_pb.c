/*
 * anm_pb.c -- byte source for the reference's nanopb decoder and minimal ip.proto wire
 * helpers (include/anmodem_pb.h).  Host C.
 *
 * Follows the wire behaviour of the reference's decoder for the messages this path carries:
 * varint (hardware/lib/nanopb/src/pb_decode.c:170-232), tag split (pb_decode.c:288-303),
 * length-delimited substreams (pb_decode.c:359-387) and the delimited wrapper
 * (pb_decode.c:1142-1168); message shapes from protocol/ip.proto:9-64.  Written from the
 * protobuf wire format, not copied from nanopb.
 */
#include "../../include/anmodem_pb.h"

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

/* returns bytes consumed, 0 on error; rejects encodings longer than 10 bytes and, like
 * pb_decode_varint32 (pb_decode.c:206-229), lengths that do not fit 32 bits */
static size_t read_varint(const uint8_t *p, size_t len, uint64_t *out) {
    uint64_t v = 0;
    for (size_t i = 0; i < len && i < 10; ++i) {
        v |= (uint64_t)(p[i] & 0x7F) << (7 * i);
        if (!(p[i] & 0x80)) {
            *out = v;
            return i + 1;
        }
    }
    return 0;
}

static size_t skip_field(const uint8_t *p, size_t len, uint32_t wt) {
    uint64_t v;
    size_t n;
    switch (wt) {
    case 0: return read_varint(p, len, &v);
    case 1: return len >= 8 ? 8 : 0;
    case 2:
        n = read_varint(p, len, &v);
        if (!n || v > len - n) return 0;
        return n + (size_t)v;
    case 5: return len >= 4 ? 4 : 0;
    default: return 0;
    }
}

size_t anm_pb_scan_to_receiver_audio(const uint8_t *buf, size_t len, const uint8_t **payload, size_t *payload_len) {
    if (!buf || !payload || !payload_len) return 0;
    uint64_t mlen;
    size_t n = read_varint(buf, len, &mlen);
    if (!n || mlen > 0xFFFFFFFFull || mlen > len - n) return 0;
    const uint8_t *m = buf + n, *mend = m + mlen;
    const uint8_t *found = NULL;
    size_t found_len = 0;
    while (m < mend) {
        uint64_t tag;
        size_t k = read_varint(m, (size_t)(mend - m), &tag);
        if (!k || (tag >> 3) == 0) return 0;
        m += k;
        if ((tag >> 3) == 1 && (tag & 7) == 2) { /* audio_data submessage */
            uint64_t sl;
            k = read_varint(m, (size_t)(mend - m), &sl);
            if (!k || sl > (uint64_t)(mend - m - k)) return 0;
            const uint8_t *a = m + k, *aend = a + sl;
            m = aend;
            bool have = false;
            while (a < aend) {
                uint64_t t2;
                size_t k2 = read_varint(a, (size_t)(aend - a), &t2);
                if (!k2 || (t2 >> 3) == 0) return 0;
                a += k2;
                if ((t2 >> 3) == 1 && (t2 & 7) == 2) {
                    uint64_t bl;
                    k2 = read_varint(a, (size_t)(aend - a), &bl);
                    if (!k2 || bl > (uint64_t)(aend - a - k2)) return 0;
                    found = a + k2;
                    found_len = (size_t)bl;
                    a += k2 + bl;
                    have = true;
                } else {
                    k2 = skip_field(a, (size_t)(aend - a), (uint32_t)(t2 & 7));
                    if (!k2) return 0;
                    a += k2;
                }
            }
            if (!have) return 0; /* required field missing (pb_decode.c:1100-1138) */
        } else {
            k = skip_field(m, (size_t)(mend - m), (uint32_t)(tag & 7));
            if (!k) return 0;
            m += k;
        }
    }
    if (!found) return 0;
    *payload = found;
    *payload_len = found_len;
    return n + (size_t)mlen;
}
