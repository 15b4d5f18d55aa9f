/*
 * anm_oracle.c -- CPU oracle of the SPEC.md receive path.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may build, load or call this file.  The product (audio-network_b200/) never does.
 *
 * PARITY UNPINNED with respect to tmarsteel/audio-network: SURVEY.md section 0/8(c)
 * shows the reference has no demodulator, no golden vectors and no known-answer tests
 * for this path, so there is no reference code for this file to follow.  It is a plain,
 * sequential, hop-by-hop restatement of SPEC.md sections 3-5 (one channel, one hop at a
 * time, the way a firmware demodulator would run), written independently of the CUDA
 * kernels.  The one stage that IS reference-defined -- decoding recovered payload bytes
 * with nanopb (hardware/lib/nanopb/src/pb_decode.c:1142-1168, hardware/src/protogen/
 * ip.pb.c) -- is checked with oracle/_ref (see oracle/Makefile).
 *
 * Build: gcc -O2 -mfma -ffp-contract=off (explicit fmaf only; no reassociation).
 */
#include "anm_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

#define RING 512 /* hop-record history, power of two >= P*S + S + 2 */
#define MAXT 64
#define MAXLVL 3

enum { ST_SEARCH = 0, ST_PEAK = 1, ST_HEADER = 2, ST_BODY = 3 };

struct anm_oracle {
    anm_config_t cfg;
    uint32_t N, S, H, T, b, P, lvl;
    float *tw; /* [N][T][2] */
    /* sample accumulation */
    int16_t hopbuf[512];
    uint32_t nbuf;
    uint64_t hop; /* index of the hop being assembled */
    /* tree history: hist[l][h & 7][k][2]; level 0 = hop partials */
    float hist[MAXLVL + 1][8][MAXT][2];
    /* SPEC 3 (every tone bin a multiple of S/2): centre-folded hop partials */
    int fold;
    float *ftw; /* [H/2][T][2] */
    /* SPEC 3b (dense tone sets, T >= 32): int8 basis and exact integer hop partials / window sums */
    int dense;
    int8_t *bq; /* [N][T][2] */
    int32_t ihist[MAXLVL + 1][8][MAXT][2];
    /* hop records */
    uint8_t rd[RING];
    float re[RING];
    /* state machine */
    int state;
    uint64_t peak_end, best_h, t0, next, prev_hop;
    float best_q;
    uint32_t nsym, total, hdr_syms;
    int32_t acc;
    uint8_t s_prev, s_prev2; /* tones of symbols nsym-1 and nsym-2 */
    uint8_t *fsyms;          /* tones of the current frame */
    /* outputs */
    anm_frame_t *frames;
    size_t nframes, capframes;
    uint8_t *bytes;
    size_t nbytes, capbytes;
    uint8_t *syms;
    size_t nsyms, capsyms;
    anm_chan_stats_t stats;
    /* trace */
    float *trE, *trEmax;
    uint8_t *trD;
    size_t trcap;
};

static uint16_t crc16(const uint8_t *d, size_t n, uint16_t crc) {
    for (size_t i = 0; i < n; ++i) {
        crc ^= (uint16_t)d[i] << 8;
        for (int k = 0; k < 8; ++k) crc = (crc & 0x8000) ? (uint16_t)((crc << 1) ^ 0x1021) : (uint16_t)(crc << 1);
    }
    return crc;
}
static uint8_t crc8(const uint8_t *d, size_t n) {
    uint8_t crc = 0;
    for (size_t i = 0; i < n; ++i) {
        crc ^= d[i];
        for (int k = 0; k < 8; ++k) crc = (crc & 0x80) ? (uint8_t)((crc << 1) ^ 0x07) : (uint8_t)(crc << 1);
    }
    return crc;
}
static uint32_t gray_inv(uint32_t g) {
    uint32_t v = g;
    for (uint32_t s = 1; s < 8; s <<= 1) v ^= v >> s;
    return v;
}

anm_oracle_t *anm_oracle_create(const anm_config_t *cfg, const float *twiddles) {
    anm_oracle_t *o = (anm_oracle_t *)calloc(1, sizeof *o);
    if (!o) return NULL;
    o->cfg = *cfg;
    o->N = cfg->sym_len;
    o->S = cfg->hops_per_sym;
    o->H = o->N / o->S;
    o->T = cfg->n_tones;
    o->P = cfg->preamble_len;
    while ((1u << o->b) < o->T) ++o->b;
    while ((1u << o->lvl) < o->S) ++o->lvl;
    o->hdr_syms = (24 + o->b - 1) / o->b;
    size_t ntw = (size_t)o->N * o->T * 2;
    o->tw = (float *)malloc(ntw * sizeof(float));
    if (twiddles) {
        memcpy(o->tw, twiddles, ntw * sizeof(float)); /* a caller-supplied table (tests that perturb it) */
    } else {
        /* SPEC 3, the oracle's own table: first quarter from libm on the exactly reduced angle, the other three
         * quarters by multiplying (C - j Sn) with -j once per quarter turn the tone advances */
        const double two_pi = 6.283185307179586476925286766559;
        for (uint32_t k = 0; k < o->T; ++k)
            for (uint32_t m = 0; m < o->N / 4; ++m) {
                double a = two_pi * (double)((cfg->tone_bin[k] * m) % o->N) / (double)o->N;
                float c = (float)cos(a), s = (float)sin(a);
                for (uint32_t q = 0; q < 4; ++q) {
                    float cq = c, sq = s;
                    for (uint32_t t = 0; t < ((cfg->tone_bin[k] * q) & 3u); ++t) { /* (c - j s) * (-j) = -s - j c */
                        float nc = -sq, ns = cq;
                        cq = nc;
                        sq = ns;
                    }
                    o->tw[((size_t)(m + q * (o->N / 4)) * o->T + k) * 2 + 0] = cq;
                    o->tw[((size_t)(m + q * (o->N / 4)) * o->T + k) * 2 + 1] = sq;
                }
            }
    }
    o->fsyms = (uint8_t *)malloc(((size_t)cfg->max_payload + 8) * 8 + 64);
    o->dense = o->T >= 32;
    o->fold = !o->dense && (o->H % 16) == 0;
    for (uint32_t k = 0; k < o->T; ++k)
        if ((2 * cfg->tone_bin[k]) % o->S) o->fold = 0;
    if (o->fold) {
        /* twiddles of the sample pairs k + 1/2 away from a hop centre, angle reduced exactly first */
        const double two_pi = 6.283185307179586476925286766559;
        o->ftw = (float *)malloc((size_t)(o->H / 2) * o->T * 2 * sizeof(float));
        for (uint32_t k = 0; k < o->H / 2; ++k)
            for (uint32_t t = 0; t < o->T; ++t) {
                double a = two_pi * (double)((cfg->tone_bin[t] * (2 * k + 1)) % (2 * o->N)) / (double)(2 * o->N);
                o->ftw[(k * o->T + t) * 2 + 0] = (float)cos(a);
                o->ftw[(k * o->T + t) * 2 + 1] = (float)sin(a);
            }
    }
    if (o->dense) {
        /* SPEC 3b: first quarter = round(127 cos), round(127 sin) of the reduced angle; the other
         * quarters by the exact rotation (-j)^(bin q) */
        const double two_pi = 6.283185307179586476925286766559;
        o->bq = (int8_t *)malloc((size_t)o->N * o->T * 2);
        for (uint32_t m = 0; m < o->N / 4; ++m)
            for (uint32_t k = 0; k < o->T; ++k) {
                double a = two_pi * (double)((cfg->tone_bin[k] * m) % o->N) / (double)o->N;
                int co = (int)lround(127.0 * cos(a)), si = (int)lround(127.0 * sin(a));
                for (uint32_t q = 0; q < 4; ++q) {
                    int cq = co, sq = si;
                    for (uint32_t t = 0; t < ((cfg->tone_bin[k] * q) & 3u); ++t) { /* multiply (c - j s) by -j */
                        int nc = -sq, ns = cq;
                        cq = nc;
                        sq = ns;
                    }
                    o->bq[((size_t)(m + q * (o->N / 4)) * o->T + k) * 2 + 0] = (int8_t)cq;
                    o->bq[((size_t)(m + q * (o->N / 4)) * o->T + k) * 2 + 1] = (int8_t)sq;
                }
            }
    }
    anm_oracle_reset(o);
    return o;
}

void anm_oracle_reset(anm_oracle_t *o) {
    o->nbuf = 0;
    o->hop = 0;
    memset(o->hist, 0, sizeof o->hist);
    memset(o->ihist, 0, sizeof o->ihist);
    memset(o->rd, 0xFF, sizeof o->rd);
    memset(o->re, 0, sizeof o->re);
    o->state = ST_SEARCH;
    o->nframes = o->nbytes = o->nsyms = 0;
    memset(&o->stats, 0, sizeof o->stats);
}

void anm_oracle_destroy(anm_oracle_t *o) {
    if (!o) return;
    free(o->tw);
    free(o->bq);
    free(o->ftw);
    free(o->fsyms);
    free(o->frames);
    free(o->bytes);
    free(o->syms);
    free(o);
}

void anm_oracle_set_trace(anm_oracle_t *o, float *E, uint8_t *d, float *emax, size_t cap_hops) {
    o->trE = E;
    o->trD = d;
    o->trEmax = emax;
    o->trcap = cap_hops;
}

/* SPEC 5: match count and quality of the alignment whose last preamble symbol ends at hop h */
static uint32_t match_quality(const anm_oracle_t *o, uint64_t h, float *q) {
    float leaf[ANM_MAX_PREAMBLE];
    uint32_t m = 0;
    for (uint32_t p = 0; p < o->P; ++p) {
        uint64_t back = (uint64_t)(o->P - 1 - p) * o->S;
        leaf[p] = 0.0f;
        if (back > h) continue; /* before the stream: d = 0xFF never matches */
        uint64_t hh = h - back;
        if (o->rd[hh & (RING - 1)] == o->cfg.preamble[p]) {
            ++m;
            leaf[p] = o->re[hh & (RING - 1)];
        }
    }
    for (uint32_t w = 1; w < o->P; w <<= 1)
        for (uint32_t p = 0; p < o->P; p += 2 * w) leaf[p] = leaf[p] + leaf[p + w];
    *q = leaf[0];
    return m;
}

static void push_sym(anm_oracle_t *o, uint8_t s) {
    if (o->nsyms == o->capsyms) {
        o->capsyms = o->capsyms ? o->capsyms * 2 : 4096;
        o->syms = (uint8_t *)realloc(o->syms, o->capsyms);
    }
    o->syms[o->nsyms++] = s;
    o->stats.symbols++;
}

/* tones -> bytes of one section (b-bit Gray-decoded values, MSB first) */
static void unpack_section(const anm_oracle_t *o, const uint8_t *tones, size_t nbytes, uint8_t *out) {
    memset(out, 0, nbytes);
    size_t nbits = nbytes * 8;
    for (size_t bit = 0; bit < nbits; ++bit) {
        size_t s = bit / o->b;
        uint32_t v = gray_inv(tones[s]);
        uint32_t x = (v >> (o->b - 1 - (bit % o->b))) & 1u;
        out[bit >> 3] |= (uint8_t)(x << (7 - (bit & 7)));
    }
}

static void emit_frame(anm_oracle_t *o, uint32_t len) {
    uint8_t hdr[3];
    unpack_section(o, o->fsyms, 3, hdr);
    uint8_t *body = (uint8_t *)malloc(len + 2);
    unpack_section(o, o->fsyms + o->hdr_syms, len + 2, body);
    uint16_t crc = crc16(body, len, crc16(hdr, 2, 0xFFFF));
    uint16_t got = (uint16_t)((body[len] << 8) | body[len + 1]);
    if (o->nframes == o->capframes) {
        o->capframes = o->capframes ? o->capframes * 2 : 64;
        o->frames = (anm_frame_t *)realloc(o->frames, o->capframes * sizeof(anm_frame_t));
    }
    if (o->nbytes + len > o->capbytes) {
        o->capbytes = (o->nbytes + len) * 2 + 1024;
        o->bytes = (uint8_t *)realloc(o->bytes, o->capbytes);
    }
    anm_frame_t *f = &o->frames[o->nframes++];
    f->channel = 0;
    f->len = len;
    f->start_sample = (o->t0 + 1 - (uint64_t)o->P * o->S) * o->H;
    f->crc_ok = crc == got;
    f->offset = (uint32_t)o->nbytes;
    memcpy(o->bytes + o->nbytes, body, len);
    o->nbytes += len;
    if (f->crc_ok) o->stats.frames_ok++; else o->stats.frames_bad++;
    free(body);
}

static void state_step(anm_oracle_t *o, uint64_t h) {
    float q;
    switch (o->state) {
    case ST_SEARCH: {
        uint32_t m = match_quality(o, h, &q);
        if (m >= o->P - o->cfg.sync_tol) {
            o->best_q = q;
            o->best_h = h;
            o->peak_end = h + o->S - 1;
            o->state = ST_PEAK;
        }
        break;
    }
    case ST_PEAK: {
        uint32_t m = match_quality(o, h, &q);
        if (m >= o->P - o->cfg.sync_tol && q > o->best_q) {
            o->best_q = q;
            o->best_h = h;
        }
        if (h == o->peak_end) {
            o->t0 = o->best_h;
            o->next = o->t0 + o->S;
            o->nsym = 0;
            o->acc = 0;
            o->s_prev = o->cfg.preamble[o->P - 1];
            o->s_prev2 = 0xFF;
            o->prev_hop = o->t0;
            o->state = ST_HEADER;
            o->stats.locks++;
        }
        break;
    }
    default: {
        if (h != o->next) break;
        uint8_t s = o->rd[h & (RING - 1)];
        o->fsyms[o->nsym] = s;
        push_sym(o, s);
        if (o->nsym >= 1) {
            uint64_t hj = o->prev_hop;
            uint8_t sj = o->s_prev, sjm = o->s_prev2;
            float e_on = o->re[hj & (RING - 1)];
            float e_early = o->rd[(hj - 1) & (RING - 1)] == sj ? o->re[(hj - 1) & (RING - 1)] : 0.0f;
            float e_late = o->rd[(hj + 1) & (RING - 1)] == sj ? o->re[(hj + 1) & (RING - 1)] : 0.0f;
            o->acc += (int)((s != sj) && e_late > e_on) - (int)((sjm != sj) && e_early > e_on);
        }
        o->s_prev2 = o->s_prev;
        o->s_prev = s;
        o->prev_hop = h;
        o->nsym++;
        o->next += o->S;
        if (o->nsym % o->cfg.trk_epoch == 0) {
            if (o->acc >= (int32_t)o->cfg.trk_thresh) { o->next += 1; o->stats.trk_moves++; }
            else if (o->acc <= -(int32_t)o->cfg.trk_thresh) { o->next -= 1; o->stats.trk_moves--; }
            o->acc = 0;
        }
        if (o->state == ST_HEADER && o->nsym == o->hdr_syms) {
            uint8_t hdr[3];
            unpack_section(o, o->fsyms, 3, hdr);
            uint32_t len = ((uint32_t)hdr[0] << 8) | hdr[1];
            if (len == 0 || len > o->cfg.max_payload || crc8(hdr, 2) != hdr[2]) {
                o->stats.header_fail++;
                o->state = ST_SEARCH;
            } else {
                o->total = o->hdr_syms + ((len + 2) * 8 + o->b - 1) / o->b;
                o->state = ST_BODY;
            }
        } else if (o->state == ST_BODY && o->nsym == o->total) {
            uint8_t hdr[3];
            unpack_section(o, o->fsyms, 3, hdr);
            emit_frame(o, ((uint32_t)hdr[0] << 8) | hdr[1]);
            o->state = ST_SEARCH;
        }
        break;
    }
    }
}

/* SPEC 3b: hop partials, window sums (exact integers) and energies of a dense tone set */
static void dense_energies(anm_oracle_t *o, uint64_t h, float *E) {
    const uint32_t T = o->T, H = o->H;
    const uint32_t m0 = (uint32_t)((h * H) % o->N);
    int32_t(*P)[2] = o->ihist[0][h & 7];
    for (uint32_t k = 0; k < T; ++k) {
        int32_t I = 0, Q = 0;
        for (uint32_t j = 0; j < H; ++j) {
            const int8_t *b = o->bq + ((size_t)(m0 + j) * T + k) * 2;
            I += (int32_t)o->hopbuf[j] * b[0];
            Q += (int32_t)o->hopbuf[j] * b[1];
        }
        P[k][0] = I;
        P[k][1] = Q;
    }
    for (uint32_t l = 1; l <= o->lvl; ++l) {
        uint32_t d = 1u << (l - 1);
        int32_t(*cur)[2] = o->ihist[l][h & 7];
        int32_t(*a)[2] = o->ihist[l - 1][(h - d) & 7];
        int32_t(*bb)[2] = o->ihist[l - 1][h & 7];
        for (uint32_t k = 0; k < T; ++k) {
            cur[k][0] = a[k][0] + bb[k][0];
            cur[k][1] = a[k][1] + bb[k][1];
        }
    }
    int32_t(*W)[2] = o->ihist[o->lvl][h & 7];
    for (uint32_t k = 0; k < T; ++k) {
        float fi = (float)W[k][0], fq = (float)W[k][1]; /* round to nearest */
        E[k] = fmaf(fi, fi, fq * fq);
    }
}

/* SPEC 3: one complete hop of H samples */
static void process_hop(anm_oracle_t *o) {
    const uint64_t h = o->hop;
    const uint32_t T = o->T, H = o->H;
    float Eh[MAXT];
    if (o->dense) {
        dense_energies(o, h, Eh);
    } else {
        float(*P)[2] = o->hist[0][h & 7];
        if (o->fold) {
            /* SPEC 3, centre folding: the samples k + 1/2 after and before the hop centre share a twiddle
             * up to conjugation; their sum and difference are exact in fp32 */
            for (uint32_t k = 0; k < T; ++k) {
                float A = 0.0f, Bq = 0.0f;
                for (uint32_t j = 0; j < H / 2; ++j) {
                    float a = (float)o->hopbuf[H / 2 + j], b = (float)o->hopbuf[H / 2 - 1 - j];
                    const float *tw = o->ftw + ((size_t)j * T + k) * 2;
                    A = fmaf(a + b, tw[0], A);
                    Bq = fmaf(a - b, tw[1], Bq);
                }
                /* relative quarter turns between hop centres: a sign flip on odd hops of "odd" tones */
                if (((2 * o->cfg.tone_bin[k] / o->S) & 1u) && (h & 1u)) { A = -A; Bq = -Bq; }
                P[k][0] = A;
                P[k][1] = Bq;
            }
        } else {
            const uint32_t m0 = (uint32_t)((h * H) % o->N);
            for (uint32_t k = 0; k < T; ++k) {
                const float *tw = o->tw + ((size_t)m0 * T + k) * 2;
                float x = (float)o->hopbuf[0];
                float I = x * tw[0], Q = x * tw[1];
                for (uint32_t j = 1; j < H; ++j) {
                    tw += (size_t)T * 2;
                    x = (float)o->hopbuf[j];
                    I = fmaf(x, tw[0], I);
                    Q = fmaf(x, tw[1], Q);
                }
                P[k][0] = I;
                P[k][1] = Q;
            }
        }
        for (uint32_t l = 1; l <= o->lvl; ++l) {
            uint32_t d = 1u << (l - 1);
            float(*cur)[2] = o->hist[l][h & 7];
            float(*a)[2] = o->hist[l - 1][(h - d) & 7]; /* zero before the stream (hist zero-initialised) */
            float(*bb)[2] = o->hist[l - 1][h & 7];
            for (uint32_t k = 0; k < T; ++k) {
                cur[k][0] = a[k][0] + bb[k][0];
                cur[k][1] = a[k][1] + bb[k][1];
            }
        }
        float(*W)[2] = o->hist[o->lvl][h & 7];
        for (uint32_t k = 0; k < T; ++k) Eh[k] = fmaf(W[k][0], W[k][0], W[k][1] * W[k][1]);
    }
    uint32_t best = 0;
    float emax = 0.0f;
    for (uint32_t k = 0; k < T; ++k) {
        float E = Eh[k];
        if (o->trE && h < o->trcap) o->trE[h * T + k] = E;
        if (k == 0 || E > emax) {
            emax = E;
            best = k;
        }
    }
    o->rd[h & (RING - 1)] = (uint8_t)best;
    o->re[h & (RING - 1)] = emax;
    if (h < o->trcap) {
        if (o->trD) o->trD[h] = (uint8_t)best;
        if (o->trEmax) o->trEmax[h] = emax;
    }
    state_step(o, h);
    o->hop = h + 1;
}

void anm_oracle_feed(anm_oracle_t *o, const int16_t *pcm, size_t n) {
    size_t i = 0;
    while (i < n) {
        size_t take = o->H - o->nbuf;
        if (take > n - i) take = n - i;
        memcpy(o->hopbuf + o->nbuf, pcm + i, take * sizeof(int16_t));
        o->nbuf += (uint32_t)take;
        i += take;
        if (o->nbuf == o->H) {
            process_hop(o);
            o->nbuf = 0;
        }
    }
}

size_t anm_oracle_num_frames(const anm_oracle_t *o) { return o->nframes; }
const anm_frame_t *anm_oracle_frames(const anm_oracle_t *o) { return o->frames; }
const uint8_t *anm_oracle_bytes(const anm_oracle_t *o) { return o->bytes; }
size_t anm_oracle_num_bytes(const anm_oracle_t *o) { return o->nbytes; }
size_t anm_oracle_num_symbols(const anm_oracle_t *o) { return o->nsyms; }
const uint8_t *anm_oracle_symbols(const anm_oracle_t *o) { return o->syms; }
void anm_oracle_stats(const anm_oracle_t *o, anm_chan_stats_t *out) { *out = o->stats; }
