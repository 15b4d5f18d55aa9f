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
 * One thread per frame: frames are short (<= max_payload bytes) and independent.  The wire primitives
 * (anm_pb_wire.h) are transcriptions of nanopb's rules -- see the notice there; the message walk is written
 * from the behaviour of the code above.  Parity is checked against the reference's own nanopb compiled in
 * place (oracle/_ref) on valid, truncated and mutated messages (tests/test_pb_gpu.py).
 */
#include <cuda_runtime.h>

#include <vector>

#include "../../include/anmodem_pb.h"
#include "anm_internal.h"
#include "anm_pb_wire.h"

namespace {

/* wire primitives and the ToReceiver walk: anm_pb_wire.h (one implementation for the device and the host scanner) */
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
    anm_wstream_t s = {bytes, mask, f.offset, f.len};
    bool have = false;
    uint32_t a_off = 0, a_len = 0;
    if (anm_w_to_receiver(&s, &have, &a_off, &a_len)) {
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
