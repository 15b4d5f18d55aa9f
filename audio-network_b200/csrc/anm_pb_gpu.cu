/*
 * anm_pb_gpu.cu -- batched delimited-protobuf deframer: the decode step of the
 * reference's receive loop, hardware/src/network.cpp:406-430, for thousands of frames at once.
 *
 * For every frame the kernel walks the payload as ONE varint-delimited ToReceiver message exactly the
 * way pb_decode_delimited(&stream, ToReceiver_fields, &msg) does with the reference's field callback:
 *   pb_decode_ex / PB_DECODE_DELIMITED      hardware/lib/nanopb/src/pb_decode.c:1142-1168
 *   pb_decode_inner (tag loop, zero tag, unknown fields, required-field bitmap)   :978-1140
 *   pb_decode_varint32_eof (overflow rules)  :170-232      pb_skip_field / pb_skip_varint / pb_skip_string :262-315
 *   pb_make / close_string_substream         :359-387      decode_callback_field (non-string wire types too) :743-789
 *   oneof + submessage decode                :519-560, 1568-1618
 *   network_pb_callback_audio_data           hardware/src/network.cpp:212-249 (rejects > 4096 bytes, :223)
 * and reports where the Opus bytes lie instead of copying them into two mallocs per frame.
 * One thread per frame: frames are short (<= max_payload bytes) and independent.  Written from the
 * behaviour of the code above, not copied from it; parity is checked against the reference's own
 * nanopb compiled in place (oracle/_ref) on valid, truncated and mutated messages (tests/test_pb_gpu.py).
 */
#include <cuda_runtime.h>

#include <vector>

#include "../../include/anmodem_pb.h"
#include "anm_internal.h"

namespace {

constexpr uint32_t kMaxEncodedFrame = 4096; /* MAX_ENCODED_FRAME_SIZE, hardware/src/network.cpp:24 */

struct Stream {
    const uint8_t *bytes;
    uint32_t mask; /* arena ring mask, 0xFFFFFFFF for a linear array */
    uint32_t pos;  /* absolute arena position of the next byte */
    uint32_t left; /* bytes_left */
};

__device__ __forceinline__ bool rd(Stream &s, uint32_t &b) {
    if (s.left == 0) return false; /* "end-of-stream" */
    b = s.bytes[s.pos & s.mask];
    ++s.pos;
    --s.left;
    return true;
}
__device__ __forceinline__ bool skip(Stream &s, uint32_t n) {
    if (s.left < n) return false;
    s.pos += n;
    s.left -= n;
    return true;
}

/* pb_decode_varint32_eof */
__device__ bool varint32(Stream &s, uint32_t &out, bool *eof) {
    uint32_t byte;
    if (!rd(s, byte)) {
        if (eof) *eof = true; /* bytes_left == 0 */
        return false;
    }
    uint32_t result;
    if ((byte & 0x80u) == 0) {
        result = byte;
    } else {
        uint32_t bitpos = 7;
        result = byte & 0x7Fu;
        do {
            if (!rd(s, byte)) return false;
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
    out = result;
    return true;
}

/* pb_skip_field */
__device__ bool skip_field(Stream &s, uint32_t wt) {
    uint32_t b, len;
    switch (wt) {
    case 0: /* pb_skip_varint: no length limit */
        do {
            if (!rd(s, b)) return false;
        } while (b & 0x80u);
        return true;
    case 1: return skip(s, 8);
    case 2: return varint32(s, len, nullptr) && skip(s, len);
    case 5: return skip(s, 4);
    default: return false; /* "invalid wire_type" */
    }
}

/* pb_make_string_substream: the parent keeps what follows the substream */
__device__ bool substream(Stream &s, Stream &sub) {
    uint32_t size;
    if (!varint32(s, size, nullptr)) return false;
    if (s.left < size) return false; /* "parent stream too short" */
    sub = s;
    sub.left = size;
    s.pos += size;
    s.left -= size;
    return true;
}

/* AudioData: pb_decode_inner over {1: required bytes opus_encoded_frame (callback)} */
__device__ bool decode_audio_data(Stream &s, uint32_t &a_off, uint32_t &a_len) {
    bool seen = false;
    while (s.left) {
        uint32_t t;
        bool eof = false;
        if (!varint32(s, t, &eof)) {
            if (eof) break;
            return false;
        }
        const uint32_t tag = t >> 3, wt = t & 7u;
        if (tag == 0) return false; /* "zero tag" */
        if (tag != 1) {
            if (!skip_field(s, wt)) return false;
            continue;
        }
        seen = true;
        if (wt == 2) { /* decode_callback_field, string: the callback sees the whole field */
            Stream f;
            if (!substream(s, f)) return false;
            if (f.left > kMaxEncodedFrame) return false; /* "Encoded frame exceeds max size" */
            a_off = f.pos;
            a_len = f.left;
        } else { /* scalar wire types reach the callback as their raw bytes (read_raw_value) */
            uint32_t b;
            const uint32_t p0 = s.pos;
            if (wt == 0) {
                uint32_t n = 0;
                do {
                    if (++n > 10) return false; /* "varint overflow" */
                    if (!rd(s, b)) return false;
                } while (b & 0x80u);
                a_len = n;
            } else if (wt == 1) {
                if (!skip(s, 8)) return false;
                a_len = 8;
            } else if (wt == 5) {
                if (!skip(s, 4)) return false;
                a_len = 4;
            } else {
                return false; /* "invalid wire_type" */
            }
            a_off = p0;
        }
    }
    return seen; /* "missing required field" */
}

__global__ void k_pb_deframe(const anm_frame_t *frames, uint32_t n, const uint8_t *bytes, uint32_t mask, anm_pb_span_t *out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const anm_frame_t f = frames[i];
    anm_pb_span_t r = {ANM_PB_FAIL, 0u, 0u, 0u};
    if (!f.crc_ok) {
        r.status = ANM_PB_CRC;
        out[i] = r;
        return;
    }
    Stream s = {bytes, mask, f.offset, f.len};
    Stream m;
    bool ok = substream(s, m); /* PB_DECODE_DELIMITED */
    bool have = false;
    uint32_t a_off = 0, a_len = 0;
    if (ok) {
        /* ToReceiver: pb_decode_inner over {oneof message {1: AudioData audio_data}} */
        while (m.left) {
            uint32_t t;
            bool eof = false;
            if (!varint32(m, t, &eof)) {
                ok = eof;
                break;
            }
            const uint32_t tag = t >> 3, wt = t & 7u;
            if (tag == 0) { ok = false; break; }
            if (tag != 1) {
                if (!skip_field(m, wt)) { ok = false; break; }
                continue;
            }
            if (wt != 2) { ok = false; break; } /* submessage: "wrong wire type" */
            Stream a;
            if (!substream(m, a)) { ok = false; break; }
            have = true; /* which_message = audio_data */
            if (!decode_audio_data(a, a_off, a_len)) { ok = false; break; }
        }
    }
    if (ok) {
        r.status = have ? ANM_PB_OK : ANM_PB_NO_AUDIO;
        r.consumed = s.pos - f.offset; /* length varint + message, like len - bytes_left of the buffer stream */
        r.audio_offset = have ? a_off : 0u;
        r.audio_len = have ? a_len : 0u;
    }
    out[i] = r;
}

} /* namespace */

extern "C" int anm_pb_deframe_device(const anm_frame_t *d_frames, uint32_t n_frames, const uint8_t *d_bytes, uint32_t bytes_mask,
                                     anm_pb_span_t *d_out, void *stream) {
    if ((!d_frames || !d_out) && n_frames) return ANM_ERR_ARG;
    if (n_frames == 0) return ANM_OK;
    if (bytes_mask != 0xFFFFFFFFu && (bytes_mask & (bytes_mask + 1u)) != 0u) return ANM_ERR_ARG; /* 2^k - 1 */
    k_pb_deframe<<<(n_frames + 127u) / 128u, 128, 0, (cudaStream_t)stream>>>(d_frames, n_frames, d_bytes, bytes_mask, d_out);
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        anm_set_error("k_pb_deframe launch failed: %s", cudaGetErrorString(e));
        return ANM_ERR_CUDA;
    }
    return ANM_OK;
}

extern "C" int anm_pb_deframe_host(const anm_frame_t *frames, size_t n_frames, const uint8_t *bytes, size_t n_bytes, anm_pb_span_t *out) {
    if ((!frames || !out) && n_frames) return ANM_ERR_ARG;
    if (n_frames == 0) return ANM_OK;
    for (size_t i = 0; i < n_frames; ++i)
        if ((size_t)frames[i].offset + frames[i].len > n_bytes) return ANM_ERR_ARG;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        anm_set_error("no CUDA device: the deframer has no CPU fallback");
        return ANM_ERR_CUDA;
    }
    anm_frame_t *d_f = nullptr;
    uint8_t *d_b = nullptr;
    anm_pb_span_t *d_o = nullptr;
    int rc = ANM_ERR_CUDA;
    if (cudaMalloc(&d_f, n_frames * sizeof(anm_frame_t)) == cudaSuccess && cudaMalloc(&d_b, n_bytes ? n_bytes : 1) == cudaSuccess &&
        cudaMalloc(&d_o, n_frames * sizeof(anm_pb_span_t)) == cudaSuccess &&
        cudaMemcpy(d_f, frames, n_frames * sizeof(anm_frame_t), cudaMemcpyHostToDevice) == cudaSuccess &&
        cudaMemcpy(d_b, bytes, n_bytes, cudaMemcpyHostToDevice) == cudaSuccess) {
        rc = anm_pb_deframe_device(d_f, (uint32_t)n_frames, d_b, 0xFFFFFFFFu, d_o, nullptr);
        if (rc == ANM_OK && cudaMemcpy(out, d_o, n_frames * sizeof(anm_pb_span_t), cudaMemcpyDeviceToHost) != cudaSuccess) rc = ANM_ERR_CUDA;
    }
    if (rc == ANM_ERR_CUDA) anm_set_error("anm_pb_deframe_host: %s", cudaGetErrorString(cudaGetLastError()));
    cudaFree(d_f);
    cudaFree(d_b);
    cudaFree(d_o);
    return rc;
}
