/*
 * anm_opus_gpu.cu -- batched Opus packet parse (first stage; include/anmodem_opus.h).
 *
 * One thread per packet restates what opus_decode_native does before it touches a frame
 * (hardware/lib/libopus/src/opus_decoder.c:661-669):
 *   opus_packet_get_samples_per_frame   opus.c:170-192        opus_packet_get_mode        opus_decoder.c:206-219
 *   opus_packet_get_bandwidth           opus_decoder.c:973-989  opus_packet_get_nb_channels opus_decoder.c:991-994
 *   opus_packet_get_nb_frames / _nb_samples   opus_decoder.c:996-1027
 *   parse_size                          opus.c:153-168
 *   opus_packet_parse_impl, self_delimited = 0   opus.c:194-345: code 0 / 1 / 2 / 3 packets, CBR and VBR,
 *       padding chains, the 120 ms and 1275-byte limits, every OPUS_INVALID_PACKET exit
 * on the byte arena in HBM (ring-addressed like the deframer's).  Packets are short (<= 4096 bytes,
 * network.cpp:24) and independent; the work is byte / integer bookkeeping and bit-exact by construction.
 * TRANSCRIPTION NOTICE: the framing rules are RFC 6716 section 3 and the kernel body follows opus_packet_parse_impl branch for
 * branch (last_size, framesize, count, the padding loop, every error exit) -- a restatement of that function ((c) Xiph.Org /
 * Skype, BSD 3-clause) onto one GPU thread per packet, not a new design; what is this repository's is the batching: the ring-
 * addressed arena, records staged in shared memory and stored one whole record per warp instruction.
 * Parity against the reference's libopus compiled
 * in place (oracle/_ref/libref_opus.so) on packets of its own encoder, on all 256 TOC bytes x framings, and
 * on truncated / mutated / random packets (tests/test_opus_parse.py).
 */
#include <cuda_runtime.h>

#include <cstddef>

#include "../../include/anmodem_opus.h"
#include "anm_internal.h"

namespace {

static_assert(sizeof(anm_opus_packet_t) == 128 && offsetof(anm_opus_packet_t, size) == 32, "anm_opus_packet_t layout");

struct Pkt {
    const uint8_t *bytes;
    uint32_t mask, base;
    __device__ __forceinline__ int at(int i) const { return bytes[(base + (uint32_t)i) & mask]; }
};

__device__ int samples_per_frame(int toc, int Fs) {
    if (toc & 0x80) return (Fs << ((toc >> 3) & 3)) / 400;
    if ((toc & 0x60) == 0x60) return (toc & 0x08) ? Fs / 50 : Fs / 100;
    const int a = (toc >> 3) & 3;
    return a == 3 ? Fs * 60 / 1000 : (Fs << a) / 100;
}

/* parse_size: bytes used (1 or 2) or -1, *size = -1 */
__device__ int parse_size(const Pkt &p, int pos, int len, int &size) {
    if (len < 1) { size = -1; return -1; }
    const int b0 = p.at(pos);
    if (b0 < 252) { size = b0; return 1; }
    if (len < 2) { size = -1; return -1; }
    size = 4 * p.at(pos + 1) + b0;
    return 2;
}

/* One thread per packet; the 128-byte record is built in shared memory (frame sizes are indexed at run time) and the
 * CTA's records leave as one contiguous block, each warp instruction storing one whole 128-byte record. */
constexpr int kParseThreads = 128;

__global__ void __launch_bounds__(kParseThreads) k_opus_parse(const anm_pb_span_t *spans, uint32_t n, const uint8_t *bytes, uint32_t mask, int Fs,
                                                              anm_opus_packet_t *out) {
    constexpr int kWords = (int)(sizeof(anm_opus_packet_t) / 4); /* 32 */
    __shared__ uint32_t recs[kParseThreads][kWords + 1];          /* +1 word: the threads' records fall into different banks */
    const uint32_t idx = blockIdx.x * kParseThreads + threadIdx.x;
#pragma unroll
    for (int i = 0; i < kWords; ++i) recs[threadIdx.x][i] = 0u;
    if (idx < n) {
        anm_opus_packet_t &r = *reinterpret_cast<anm_opus_packet_t *>(recs[threadIdx.x]);
        const anm_pb_span_t sp = spans[idx];
        if (sp.status != ANM_PB_OK || sp.audio_len == 0u) {
            /* len == 0: opus_decode takes its packet-loss path, not the parser (opus_decoder.c:644); nothing to parse */
            r.count = ANM_OPUS_BAD_ARG;
        } else {
            const Pkt p = {bytes, mask, sp.audio_offset};
            int len = (int)sp.audio_len;
            const int len0 = len;
            const int toc = p.at(0);
            r.toc = (uint8_t)toc;
            r.channels = (toc & 0x4) ? 2 : 1;
            r.mode = (toc & 0x80) ? ANM_OPUS_MODE_CELT_ONLY : ((toc & 0x60) == 0x60 ? ANM_OPUS_MODE_HYBRID : ANM_OPUS_MODE_SILK_ONLY);
            if (toc & 0x80) {
                const int bw = 1102 + ((toc >> 5) & 3);     /* OPUS_BANDWIDTH_MEDIUMBAND + ... */
                r.bandwidth = bw == 1102 ? 1101 : bw;       /* MEDIUMBAND -> NARROWBAND */
            } else if ((toc & 0x60) == 0x60) {
                r.bandwidth = (toc & 0x10) ? 1105 : 1104;   /* FULLBAND : SUPERWIDEBAND */
            } else {
                r.bandwidth = 1101 + ((toc >> 5) & 3);
            }
            const int spf = samples_per_frame(toc, Fs);
            r.samples_per_frame = spf;
            /* opus_packet_get_nb_frames / _nb_samples */
            {
                const int c = toc & 3;
                const int nf = c == 0 ? 1 : (c != 3 ? 2 : (len0 < 2 ? ANM_OPUS_INVALID_PACKET : (p.at(1) & 0x3F)));
                r.nb_frames = nf;
                if (nf < 0) r.nb_samples = nf;
                else {
                    const int s = nf * spf;
                    r.nb_samples = (s * 25 > Fs * 3) ? ANM_OPUS_INVALID_PACKET : s;
                }
            }
            /* opus_packet_parse_impl, self_delimited = 0 */
            const int framesize = samples_per_frame(toc, 48000);
            int pos = 1, count = 0, last_size, sz;
            bool bad = false;
            len--;
            last_size = len;
            switch (toc & 3) {
            case 0:
                count = 1;
                break;
            case 1:
                count = 2;
                if (len & 1) bad = true;
                else {
                    last_size = len / 2;
                    r.size[0] = (int16_t)last_size;
                }
                break;
            case 2: {
                count = 2;
                const int used = parse_size(p, pos, len, sz);
                len -= used;
                if (sz < 0 || sz > len) bad = true;
                else {
                    r.size[0] = (int16_t)sz;
                    pos += used;
                    last_size = len - sz;
                }
                break;
            }
            default: {
                if (len < 1) { bad = true; break; }
                const int ch = p.at(pos++);
                count = ch & 0x3F;
                if (count <= 0 || framesize * count > 5760) { bad = true; break; }
                len--;
                if (ch & 0x40) { /* padding */
                    int pb;
                    do {
                        if (len <= 0) { bad = true; break; }
                        pb = p.at(pos++);
                        len--;
                        len -= (pb == 255) ? 254 : pb;
                    } while (pb == 255);
                    if (bad) break;
                }
                if (len < 0) { bad = true; break; }
                if (ch & 0x80) { /* VBR */
                    last_size = len;
                    for (int i = 0; i < count - 1; ++i) {
                        const int used = parse_size(p, pos, len, sz);
                        len -= used;
                        if (sz < 0 || sz > len) { bad = true; break; }
                        r.size[i] = (int16_t)sz;
                        pos += used;
                        last_size -= used + sz;
                    }
                    if (!bad && last_size < 0) bad = true;
                } else { /* CBR */
                    last_size = len / count;
                    if (last_size * count != len) bad = true;
                    else
                        for (int i = 0; i < count - 1; ++i) r.size[i] = (int16_t)last_size;
                }
                break;
            }
            }
            if (!bad && last_size > 1275) bad = true;
            if (bad) {
#pragma unroll
                for (int i = 8; i < kWords; ++i) recs[threadIdx.x][i] = 0u; /* size[48]: words 8..31 */
                r.count = ANM_OPUS_INVALID_PACKET;
            } else {
                r.size[count - 1] = (int16_t)last_size;
                r.count = count;
                r.payload_offset = pos;
            }
        }
    }
    __syncthreads();
    const uint32_t first = blockIdx.x * kParseThreads;
    const uint32_t nrec = min((uint32_t)kParseThreads, n - first);
    uint32_t *dstw = reinterpret_cast<uint32_t *>(out + first);
    for (uint32_t i = threadIdx.x; i < nrec * (uint32_t)kWords; i += kParseThreads) dstw[i] = recs[i / kWords][i % kWords]; /* a warp stores one whole record */
}

} /* namespace */

extern "C" int anm_opus_parse_device(const anm_pb_span_t *d_spans, uint32_t n, const uint8_t *d_bytes, uint32_t bytes_mask, int32_t Fs,
                                     anm_opus_packet_t *d_out, void *stream) {
    if ((!d_spans || !d_out) && n) return ANM_ERR_ARG;
    if (Fs != 8000 && Fs != 12000 && Fs != 16000 && Fs != 24000 && Fs != 48000) return ANM_ERR_ARG; /* opus_decoder_init's rates */
    if (n == 0) return ANM_OK;
    if (bytes_mask != 0xFFFFFFFFu && (bytes_mask & (bytes_mask + 1u)) != 0u) return ANM_ERR_ARG;
    k_opus_parse<<<(n + (uint32_t)kParseThreads - 1u) / (uint32_t)kParseThreads, kParseThreads, 0, (cudaStream_t)stream>>>(d_spans, n, d_bytes, bytes_mask, Fs, d_out);
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        anm_set_error("k_opus_parse launch failed: %s", cudaGetErrorString(e));
        return ANM_ERR_CUDA;
    }
    return ANM_OK;
}

extern "C" int anm_opus_parse_host(const anm_pb_span_t *spans, size_t n, const uint8_t *bytes, size_t n_bytes, int32_t Fs,
                                   anm_opus_packet_t *out) {
    if ((!spans || !out) && n) return ANM_ERR_ARG;
    if (n == 0) return ANM_OK;
    for (size_t i = 0; i < n; ++i)
        if (spans[i].status == ANM_PB_OK && (size_t)spans[i].audio_offset + spans[i].audio_len > n_bytes) return ANM_ERR_ARG;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        anm_set_error("no CUDA device: the Opus packet parser has no CPU fallback");
        return ANM_ERR_CUDA;
    }
    anm_pb_span_t *d_s = nullptr;
    uint8_t *d_b = nullptr;
    anm_opus_packet_t *d_o = nullptr;
    int rc = ANM_ERR_CUDA;
    if (cudaMalloc(&d_s, n * sizeof(anm_pb_span_t)) == cudaSuccess && cudaMalloc(&d_b, n_bytes ? n_bytes : 1) == cudaSuccess &&
        cudaMalloc(&d_o, n * sizeof(anm_opus_packet_t)) == cudaSuccess &&
        cudaMemcpy(d_s, spans, n * sizeof(anm_pb_span_t), cudaMemcpyHostToDevice) == cudaSuccess &&
        cudaMemcpy(d_b, bytes, n_bytes, cudaMemcpyHostToDevice) == cudaSuccess) {
        rc = anm_opus_parse_device(d_s, (uint32_t)n, d_b, 0xFFFFFFFFu, Fs, d_o, nullptr);
        if (rc == ANM_OK && cudaMemcpy(out, d_o, n * sizeof(anm_opus_packet_t), cudaMemcpyDeviceToHost) != cudaSuccess) rc = ANM_ERR_CUDA;
    }
    if (rc == ANM_ERR_CUDA) anm_set_error("anm_opus_parse_host: %s", cudaGetErrorString(cudaGetLastError()));
    cudaFree(d_s);
    cudaFree(d_b);
    cudaFree(d_o);
    return rc;
}
