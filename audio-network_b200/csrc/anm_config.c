/*
 * anm_config.c -- presets, validation, twiddle table, CRCs and transmit-side
 * framing of the SPEC.md modem.  Host C, no device code.
 *
 * Reference context: the payload carried by a frame is one varint-delimited
 * ip.proto message (protocol/ip.proto:9-64; consumer pb_decode_delimited at
 * hardware/src/network.cpp:411).  Everything else here is defined by SPEC.md.
 */
#include "anm_internal.h"

#include <math.h>
#include <string.h>

static const uint8_t PRE4[16] = {3, 0, 1, 2, 3, 2, 0, 2, 1, 0, 3, 1, 2, 0, 2, 3};
static const uint8_t PRE8[16] = {5, 0, 2, 1, 4, 2, 3, 7, 0, 6, 3, 6, 5, 7, 1, 3};
static const uint8_t PRE16[16] = {12, 13, 5, 11, 2, 14, 3, 5, 12, 11, 15, 0, 15, 1, 9, 12};
static const uint8_t PRE64[16] = {58, 3, 29, 22, 23, 11, 32, 4, 9, 10, 2, 57, 1, 35, 31, 34};
static const uint8_t PRE2[32] = {1, 0, 1, 1, 0, 1, 0, 0, 1, 1, 0, 0, 0, 1, 0, 1,
                                 1, 1, 0, 0, 0, 0, 1, 0, 1, 0, 1, 1, 0, 0, 1, 1};

static void base(anm_config_t *c, uint32_t N, uint32_t S, uint32_t T, uint32_t bin0, uint32_t step,
                 const uint8_t *pre, uint32_t P, uint32_t tol) {
    memset(c, 0, sizeof *c);
    c->sample_rate = 44100;
    c->sym_len = N;
    c->hops_per_sym = S;
    c->n_tones = T;
    for (uint32_t k = 0; k < T; ++k) c->tone_bin[k] = bin0 + k * step;
    c->preamble_len = P;
    memcpy(c->preamble, pre, P);
    c->sync_tol = tol;
    c->max_payload = 1024;
    c->trk_epoch = 16;
    c->trk_thresh = 3;
}

int anm_config_preset(const char *name, anm_config_t *out) {
    if (!name || !out) return ANM_ERR_ARG;
    if (!strcmp(name, "ref4")) base(out, 128, 4, 4, 10, 2, PRE4, 16, 2);
    else if (!strcmp(name, "bfsk2")) base(out, 128, 4, 2, 12, 4, PRE2, 32, 3);
    else if (!strcmp(name, "mfsk8")) base(out, 128, 4, 8, 8, 2, PRE8, 16, 2);
    else if (!strcmp(name, "mfsk16")) base(out, 128, 4, 16, 8, 2, PRE16, 16, 2);
    else if (!strcmp(name, "wide64")) base(out, 256, 4, 64, 16, 1, PRE64, 16, 2);
    else return ANM_ERR_ARG;
    return ANM_OK;
}

static int is_pow2(uint32_t v) { return v && !(v & (v - 1)); }

int anm_config_validate(const anm_config_t *c) {
    if (!c) return ANM_ERR_ARG;
    if (!is_pow2(c->sym_len) || c->sym_len < 32 || c->sym_len > 512) return ANM_ERR_ARG;
    if (c->hops_per_sym != 2 && c->hops_per_sym != 4 && c->hops_per_sym != 8) return ANM_ERR_ARG;
    if ((c->sym_len / c->hops_per_sym) % 8) return ANM_ERR_ARG;
    if (!is_pow2(c->n_tones) || c->n_tones < 2 || c->n_tones > ANM_MAX_TONES) return ANM_ERR_ARG;
    for (uint32_t k = 0; k < c->n_tones; ++k) {
        if (c->tone_bin[k] == 0 || c->tone_bin[k] >= c->sym_len / 2) return ANM_ERR_ARG;
        for (uint32_t j = 0; j < k; ++j)
            if (c->tone_bin[j] == c->tone_bin[k]) return ANM_ERR_ARG;
    }
    if (c->preamble_len != 8 && c->preamble_len != 16 && c->preamble_len != 32) return ANM_ERR_ARG;
    for (uint32_t p = 0; p < c->preamble_len; ++p)
        if (c->preamble[p] >= c->n_tones) return ANM_ERR_ARG;
    if (c->sync_tol >= c->preamble_len) return ANM_ERR_ARG;
    if (c->max_payload == 0 || c->max_payload > 4104) return ANM_ERR_ARG;
    if (c->trk_epoch == 0 || c->trk_thresh == 0) return ANM_ERR_ARG;
    /* dense (integer) arithmetic: |window sum| <= N * 32768 * 127 must fit a signed 32-bit integer */
    if (c->n_tones >= 32u && c->sym_len > 512u) return ANM_ERR_ARG;
    return ANM_OK;
}

int anm_twiddles(const anm_config_t *c, float *out) {
    if (anm_config_validate(c) != ANM_OK || !out) return ANM_ERR_ARG;
    const double two_pi = 6.283185307179586476925286766559;
    const uint32_t N = c->sym_len, T = c->n_tones;
    /* First quarter from libm (angle reduced exactly in integers first).  The other quarters
     * follow from e^{-j2pi b (m + q N/4)/N} = (-j)^{b q} e^{-j2pi b m/N}: an exact swap / negation
     * of (cos, sin), which the CUDA path relies on (SPEC 3). */
    for (uint32_t m = 0; m < N / 4; ++m)
        for (uint32_t k = 0; k < T; ++k) {
            uint32_t r0 = (c->tone_bin[k] * m) % N;
            double a = two_pi * (double)r0 / (double)N;
            float co = (float)cos(a), si = (float)sin(a);
            for (uint32_t q = 0; q < 4; ++q) {
                uint32_t r = (c->tone_bin[k] * q) & 3u;
                float cq = r == 0 ? co : r == 1 ? -si : r == 2 ? -co : si;
                float sq = r == 0 ? si : r == 1 ? co : r == 2 ? -si : -co;
                out[((m + q * (N / 4)) * T + k) * 2 + 0] = cq;
                out[((m + q * (N / 4)) * T + k) * 2 + 1] = sq;
            }
        }
    return ANM_OK;
}

/* SPEC 3: the hop partials are computed by centre folding when every tone bin is a multiple of S/2
 * (then the centre of every hop sits at a multiple of a quarter turn of every tone) */
int anm_config_foldable(const anm_config_t *c) {
    if (!c || anm_config_validate(c) != ANM_OK || c->n_tones >= 32u) return 0;
    if ((c->sym_len / c->hops_per_sym) % 16u) return 0;
    for (uint32_t k = 0; k < c->n_tones; ++k)
        if ((2u * c->tone_bin[k]) % c->hops_per_sym) return 0;
    return 1;
}

/* out[k][tone] = (cos, sin)(2 pi bin (k + 1/2) / N), k < H/2: the twiddles of the sample pairs at
 * distance k + 1/2 on either side of a hop centre */
int anm_fold_twiddles(const anm_config_t *c, float *out) {
    if (!anm_config_foldable(c) || !out) return ANM_ERR_ARG;
    const double two_pi = 6.283185307179586476925286766559;
    const uint32_t N = c->sym_len, T = c->n_tones, H2 = N / c->hops_per_sym / 2;
    for (uint32_t k = 0; k < H2; ++k)
        for (uint32_t t = 0; t < T; ++t) {
            uint32_t r = (c->tone_bin[t] * (2u * k + 1u)) % (2u * N); /* angle reduced exactly in integers */
            double a = two_pi * (double)r / (double)(2u * N);
            out[(k * T + t) * 2 + 0] = (float)cos(a);
            out[(k * T + t) * 2 + 1] = (float)sin(a);
        }
    return ANM_OK;
}

/* SPEC 3b: configurations with a dense tone set (T >= 32) compute tone energies from an 8-bit
 * integer basis with exact integer accumulation (the tensor-core contraction of the CUDA path). */
int anm_config_dense(const anm_config_t *c) { return c && c->n_tones >= 32u; }

int anm_basis_q7(const anm_config_t *c, int8_t *out) {
    if (anm_config_validate(c) != ANM_OK || !out) return ANM_ERR_ARG;
    const double two_pi = 6.283185307179586476925286766559;
    const uint32_t N = c->sym_len, T = c->n_tones;
    /* first quarter rounded from libm, the other quarters by the exact (-j)^{b q} symmetry (as anm_twiddles) */
    for (uint32_t m = 0; m < N / 4; ++m)
        for (uint32_t k = 0; k < T; ++k) {
            uint32_t r0 = (c->tone_bin[k] * m) % N;
            double a = two_pi * (double)r0 / (double)N;
            int co = (int)lround(127.0 * cos(a)), si = (int)lround(127.0 * sin(a));
            for (uint32_t q = 0; q < 4; ++q) {
                uint32_t r = (c->tone_bin[k] * q) & 3u;
                int cq = r == 0 ? co : r == 1 ? -si : r == 2 ? -co : si;
                int sq = r == 0 ? si : r == 1 ? co : r == 2 ? -si : -co;
                out[((m + q * (N / 4)) * T + k) * 2 + 0] = (int8_t)cq;
                out[((m + q * (N / 4)) * T + k) * 2 + 1] = (int8_t)sq;
            }
        }
    return ANM_OK;
}

uint16_t anm_crc16(const uint8_t *data, size_t len, uint16_t crc) {
    for (size_t i = 0; i < len; ++i) {
        crc ^= (uint16_t)data[i] << 8;
        for (int b = 0; b < 8; ++b) crc = (crc & 0x8000) ? (uint16_t)((crc << 1) ^ 0x1021) : (uint16_t)(crc << 1);
    }
    return crc;
}

uint8_t anm_crc8(const uint8_t *data, size_t len, uint8_t crc) {
    for (size_t i = 0; i < len; ++i) {
        crc ^= data[i];
        for (int b = 0; b < 8; ++b) crc = (crc & 0x80) ? (uint8_t)((crc << 1) ^ 0x07) : (uint8_t)(crc << 1);
    }
    return crc;
}

uint32_t anm_bits_per_sym(const anm_config_t *c) {
    uint32_t b = 0;
    while ((1u << b) < c->n_tones) ++b;
    return b;
}

size_t anm_frame_num_symbols(const anm_config_t *c, size_t len) {
    if (anm_config_validate(c) != ANM_OK || len == 0 || len > c->max_payload) return 0;
    uint32_t b = anm_bits_per_sym(c);
    return c->preamble_len + (24 + b - 1) / b + ((len + 2) * 8 + b - 1) / b;
}

/* bits MSB first -> b-bit values -> Gray-mapped tone indices, zero padded */
static size_t pack_section(const uint8_t *bytes, size_t nbytes, uint32_t b, uint8_t *syms) {
    size_t nbits = nbytes * 8, ns = (nbits + b - 1) / b;
    for (size_t s = 0; s < ns; ++s) {
        uint32_t v = 0;
        for (uint32_t i = 0; i < b; ++i) {
            size_t bit = s * b + i;
            uint32_t x = bit < nbits ? (bytes[bit >> 3] >> (7 - (bit & 7))) & 1u : 0u;
            v = (v << 1) | x;
        }
        syms[s] = (uint8_t)(v ^ (v >> 1));
    }
    return ns;
}

size_t anm_frame_symbols(const anm_config_t *c, const uint8_t *payload, size_t len, uint8_t *syms,
                         size_t cap) {
    size_t total = anm_frame_num_symbols(c, len);
    if (!total || !payload || !syms || cap < total) return 0;
    uint32_t b = anm_bits_per_sym(c);
    size_t n = 0;
    memcpy(syms, c->preamble, c->preamble_len);
    n += c->preamble_len;
    uint8_t hdr[3] = {(uint8_t)(len >> 8), (uint8_t)len, 0};
    hdr[2] = anm_crc8(hdr, 2, 0);
    n += pack_section(hdr, 3, b, syms + n);
    /* body = payload | crc16(LEN bytes + payload); build in a bounded buffer */
    uint8_t body[4104 + 2];
    memcpy(body, payload, len);
    uint16_t crc = anm_crc16(payload, len, anm_crc16(hdr, 2, 0xFFFF));
    body[len] = (uint8_t)(crc >> 8);
    body[len + 1] = (uint8_t)crc;
    n += pack_section(body, len + 2, b, syms + n);
    return n;
}

/* Order-independent checksum of a set of frames: FNV-1a over (channel-seeded) start_sample, len, crc_ok and the payload of each
 * frame, summed modulo 2^64 over the frames.  A host-side gather (several devices / processes) uses it to check that what
 * arrived is what was decoded; the test oracle computes the same quantity independently (oracle/anm_oracle_batch.c). */
uint64_t anm_frames_digest(const anm_frame_t *frames, size_t n, const uint8_t *bytes) {
    uint64_t total = 0;
    for (size_t i = 0; i < n; ++i) {
        uint64_t h = 0xCBF29CE484222325ull ^ frames[i].channel;
        const uint8_t *parts[4] = {(const uint8_t *)&frames[i].start_sample, (const uint8_t *)&frames[i].len,
                                   (const uint8_t *)&frames[i].crc_ok, bytes ? bytes + frames[i].offset : NULL};
        const size_t lens[4] = {8, 4, 4, bytes ? frames[i].len : 0};
        for (int k = 0; k < 4; ++k)
            for (size_t j = 0; j < lens[k]; ++j) h = (h ^ parts[k][j]) * 0x100000001B3ull;
        total += h;
    }
    return total;
}
