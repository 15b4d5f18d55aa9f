/*
 * anm_pb_wire.h -- the ONE set of protobuf wire primitives of this library, shared by the host codecs
 * (anm_pb.c, anm_pb_msgs.c; plain C) and the GPU deframer (anm_pb_gpu.cu; device code).
 *
 * TRANSCRIPTION NOTICE.  The functions below restate nanopb 0.4.5 (zlib licence, (c) Petteri Aimonen),
 * as compiled into the reference firmware, rule for rule, because its accept / reject verdicts are
 * defined by these exact overflow and end-of-stream rules and the parity tests compare verdicts:
 *   anm_w_varint32  = pb_decode_varint32_eof   hardware/lib/nanopb/src/pb_decode.c:170-232
 *                     (same bit positions, the `bitpos == 35 && (byte & 0x70)` test and the sign-extension
 *                     rule for 10-byte negative int32 values)
 *   anm_w_varint64  = pb_decode_varint         pb_decode.c:240-260
 *   anm_w_skip_field = pb_skip_field / pb_skip_varint / pb_skip_string   pb_decode.c:262-315
 *   anm_w_substream = pb_make_string_substream (+ the implicit pb_close_string_substream)   pb_decode.c:359-387
 * They are transcriptions of that logic onto this library's stream type (a position in a linear buffer or
 * in a power-of-two ring, so the same code walks host buffers and the device byte arena), not new designs.
 */
#ifndef ANM_PB_WIRE_H_INCLUDED
#define ANM_PB_WIRE_H_INCLUDED

#include <stdbool.h>
#include <stddef.h>
#include <stdint.h>

#ifdef __CUDACC__
#define ANM_W_FN __host__ __device__ static inline
#else
#define ANM_W_FN static inline
#endif

typedef struct anm_wstream {
    const uint8_t *bytes;
    uint32_t mask; /* ring mask of the byte arena; 0xFFFFFFFF for a linear buffer */
    uint32_t pos;  /* position of the next byte (arena position, or index into the buffer) */
    uint32_t left; /* bytes_left */
} anm_wstream_t;

ANM_W_FN bool anm_w_rd(anm_wstream_t *s, uint32_t *b) {
    if (s->left == 0) return false; /* "end-of-stream" */
    *b = s->bytes[s->pos & s->mask];
    ++s->pos;
    --s->left;
    return true;
}
ANM_W_FN bool anm_w_skip(anm_wstream_t *s, uint32_t n) {
    if (s->left < n) return false;
    s->pos += n;
    s->left -= n;
    return true;
}

ANM_W_FN bool anm_w_varint32(anm_wstream_t *s, uint32_t *out, bool *eof) {
    uint32_t byte, result;
    if (!anm_w_rd(s, &byte)) {
        if (eof) *eof = true; /* bytes_left == 0 */
        return false;
    }
    if ((byte & 0x80u) == 0) {
        result = byte;
    } else {
        uint32_t bitpos = 7;
        result = byte & 0x7Fu;
        do {
            if (!anm_w_rd(s, &byte)) return false;
            if (bitpos >= 32) {
                /* trailing 0x80 bytes, or the sign extension of a negative int32 */
                const uint32_t sign_extension = (bitpos < 63) ? 0xFFu : 0x01u;
                const bool valid = ((byte & 0x7Fu) == 0) || ((result >> 31) != 0 && byte == sign_extension);
                if (bitpos >= 64 || !valid) return false; /* "varint overflow" */
            } else {
                result |= (byte & 0x7Fu) << bitpos;
            }
            bitpos += 7;
        } while (byte & 0x80u);
        if (bitpos == 35 && (byte & 0x70u) != 0) return false; /* only 4 bits of the fifth byte fit */
    }
    *out = result;
    return true;
}

ANM_W_FN bool anm_w_varint64(anm_wstream_t *s, uint64_t *out) {
    uint32_t byte, bitpos = 0;
    uint64_t result = 0;
    do {
        if (bitpos >= 64) return false; /* "varint overflow" */
        if (!anm_w_rd(s, &byte)) return false;
        result |= (uint64_t)(byte & 0x7Fu) << bitpos;
        bitpos += 7;
    } while (byte & 0x80u);
    *out = result;
    return true;
}

ANM_W_FN bool anm_w_skip_field(anm_wstream_t *s, uint32_t wt) {
    uint32_t b, len;
    switch (wt) {
    case 0: /* pb_skip_varint: no length limit */
        do {
            if (!anm_w_rd(s, &b)) return false;
        } while (b & 0x80u);
        return true;
    case 1: return anm_w_skip(s, 8);
    case 2: return anm_w_varint32(s, &len, NULL) && anm_w_skip(s, len);
    case 5: return anm_w_skip(s, 4);
    default: return false; /* "invalid wire_type" */
    }
}

/* the parent keeps what follows the substream */
ANM_W_FN bool anm_w_substream(anm_wstream_t *s, anm_wstream_t *sub) {
    uint32_t size;
    if (!anm_w_varint32(s, &size, NULL)) return false;
    if (s->left < size) return false; /* "parent stream too short" */
    *sub = *s;
    sub->left = size;
    s->pos += size;
    s->left -= size;
    return true;
}

/* ---- ToReceiver{audio_data{opus_encoded_frame}} walk (protocol/ip.proto:29-33, 62-64) --------------------
 * pb_decode_delimited(&stream, ToReceiver_fields, &msg) with the reference's field callback
 * network_pb_callback_audio_data (hardware/src/network.cpp:212-249): pb_decode_inner's tag loop, zero tag,
 * unknown-field skipping, oneof / submessage nesting, "wrong wire type" for a known field, the required-field
 * check, the callback's 4096-byte limit (network.cpp:24, 223) and its raw-bytes behaviour for scalar wire
 * types (decode_callback_field, pb_decode.c:743-789).  One implementation for k_pb_deframe and the host scanner. */
#define ANM_W_MAX_ENCODED_FRAME 4096u

/* AudioData: {1: required bytes opus_encoded_frame (callback)} */
ANM_W_FN bool anm_w_audio_data(anm_wstream_t *s, uint32_t *a_off, uint32_t *a_len) {
    bool seen = false;
    while (s->left) {
        uint32_t t;
        bool eof = false;
        if (!anm_w_varint32(s, &t, &eof)) {
            if (eof) break;
            return false;
        }
        const uint32_t tag = t >> 3, wt = t & 7u;
        if (tag == 0) return false; /* "zero tag" */
        if (tag != 1) {
            if (!anm_w_skip_field(s, wt)) return false;
            continue;
        }
        seen = true;
        if (wt == 2) { /* string: the callback sees the whole field */
            anm_wstream_t f;
            if (!anm_w_substream(s, &f)) return false;
            if (f.left > ANM_W_MAX_ENCODED_FRAME) return false; /* "Encoded frame exceeds max size" */
            *a_off = f.pos;
            *a_len = f.left;
        } else { /* scalar wire types reach the callback as their raw bytes (read_raw_value) */
            uint32_t b;
            const uint32_t p0 = s->pos;
            if (wt == 0) {
                uint32_t n = 0;
                do {
                    if (++n > 10) return false; /* "varint overflow" */
                    if (!anm_w_rd(s, &b)) return false;
                } while (b & 0x80u);
                *a_len = n;
            } else if (wt == 1) {
                if (!anm_w_skip(s, 8)) return false;
                *a_len = 8;
            } else if (wt == 5) {
                if (!anm_w_skip(s, 4)) return false;
                *a_len = 4;
            } else {
                return false; /* "invalid wire_type" */
            }
            *a_off = p0;
        }
    }
    return seen; /* "missing required field" */
}

/* One delimited ToReceiver message at the head of `s`.  Returns false where pb_decode_delimited returns false; otherwise
 * *have says whether the oneof holds audio_data, [*a_off, *a_off + *a_len) locates the Opus bytes, and `s` stands behind
 * the message. */
ANM_W_FN bool anm_w_to_receiver(anm_wstream_t *s, bool *have, uint32_t *a_off, uint32_t *a_len) {
    anm_wstream_t m;
    *have = false;
    *a_off = *a_len = 0;
    if (!anm_w_substream(s, &m)) return false; /* PB_DECODE_DELIMITED */
    while (m.left) {
        uint32_t t;
        bool eof = false;
        if (!anm_w_varint32(&m, &t, &eof)) return eof;
        const uint32_t tag = t >> 3, wt = t & 7u;
        if (tag == 0) return false;
        if (tag != 1) {
            if (!anm_w_skip_field(&m, wt)) return false;
            continue;
        }
        if (wt != 2) return false; /* submessage: "wrong wire type" */
        anm_wstream_t a;
        if (!anm_w_substream(&m, &a)) return false;
        *have = true; /* which_message = audio_data */
        if (!anm_w_audio_data(&a, a_off, a_len)) return false;
    }
    return true;
}

#endif /* ANM_PB_WIRE_H_INCLUDED */
