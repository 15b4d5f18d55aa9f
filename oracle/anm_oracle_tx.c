/*
 * anm_oracle_tx.c -- the oracle's OWN transmit side: presets (SPEC.md section 2), frame
 * construction (section 4) and the integer transmitter (section 6).  TEST / BENCH
 * INFRASTRUCTURE ONLY -- see anm_oracle.h.
 *
 * Why it exists: until SPEC rev 2 the oracle's test signals came from the product library
 * (anm_frame_symbols / anm_tx_render in audio-network_b200/csrc), so a bug shared by both
 * sides of a parity test could hide.  This file restates the same sections of SPEC.md a second
 * time, written from the text (bit-serial frame builder, one sample at a time, no tables shared
 * with the product), so that
 *   - tests/test_oracle_tx.py can hold the product's transmitter and framer against it, and
 *   - bench.py --impl reference runs without mapping libanmodem.so at all.
 *
 * PARITY UNPINNED with respect to tmarsteel/audio-network: the reference has no modem
 * (SURVEY.md section 0); the only reference-defined bytes here are frame payloads.
 */
#define _GNU_SOURCE
#include "anm_oracle.h"

#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

/* SPEC 2: the preset table */
static const struct preset_row {
    const char *name;
    uint32_t N, S, T, bin0, bin_step, P, tol;
    uint8_t pre[32];
} k_presets[] = {
    {"ref4", 128, 4, 4, 10, 2, 16, 2, {3, 0, 1, 2, 3, 2, 0, 2, 1, 0, 3, 1, 2, 0, 2, 3}},
    {"bfsk2", 128, 4, 2, 12, 4, 32, 3, {1, 0, 1, 1, 0, 1, 0, 0, 1, 1, 0, 0, 0, 1, 0, 1, 1, 1, 0, 0, 0, 0, 1, 0, 1, 0, 1, 1, 0, 0, 1, 1}},
    {"mfsk8", 128, 4, 8, 8, 2, 16, 2, {5, 0, 2, 1, 4, 2, 3, 7, 0, 6, 3, 6, 5, 7, 1, 3}},
    {"mfsk16", 128, 4, 16, 8, 2, 16, 2, {12, 13, 5, 11, 2, 14, 3, 5, 12, 11, 15, 0, 15, 1, 9, 12}},
    {"wide64", 256, 4, 64, 16, 1, 16, 2, {58, 3, 29, 22, 23, 11, 32, 4, 9, 10, 2, 57, 1, 35, 31, 34}},
};

int anm_oracle_preset(const char *name, anm_config_t *out) {
    if (!name || !out) return -1;
    for (size_t i = 0; i < sizeof k_presets / sizeof k_presets[0]; ++i) {
        const struct preset_row *r = &k_presets[i];
        if (strcmp(name, r->name)) continue;
        memset(out, 0, sizeof *out);
        out->sample_rate = 44100;
        out->sym_len = r->N;
        out->hops_per_sym = r->S;
        out->n_tones = r->T;
        for (uint32_t k = 0; k < r->T; ++k) out->tone_bin[k] = r->bin0 + r->bin_step * k;
        out->preamble_len = r->P;
        memcpy(out->preamble, r->pre, r->P);
        out->sync_tol = r->tol;
        out->max_payload = 1024;
        out->trk_epoch = 16;
        out->trk_thresh = 3;
        return 0;
    }
    return -1;
}

/* ---- SPEC 4: frame construction, bit-serial ---------------------------------------------- */
typedef struct {
    uint8_t *syms;
    size_t n, cap;
    uint32_t b, fill, val;
} symw_t;

static int put_bit(symw_t *w, uint32_t bit) {
    w->val = (w->val << 1) | (bit & 1u);
    if (++w->fill == w->b) {
        if (w->n >= w->cap) return -1;
        w->syms[w->n++] = (uint8_t)(w->val ^ (w->val >> 1)); /* Gray: tone g = v ^ (v >> 1) */
        w->fill = 0;
        w->val = 0;
    }
    return 0;
}
static int put_byte(symw_t *w, uint32_t byte) {
    for (int i = 7; i >= 0; --i)
        if (put_bit(w, byte >> i)) return -1;
    return 0;
}
static int flush_section(symw_t *w) { /* zero padding up to a whole symbol */
    while (w->fill)
        if (put_bit(w, 0)) return -1;
    return 0;
}
/* CRC registers advanced one message bit at a time (no byte tables) */
static uint32_t crc_bits(uint32_t crc, uint32_t byte, uint32_t width, uint32_t poly) {
    const uint32_t top = 1u << (width - 1), mask = (top << 1) - 1u;
    for (int i = 7; i >= 0; --i) {
        const uint32_t fb = ((crc & top) ? 1u : 0u) ^ ((byte >> i) & 1u);
        crc = (crc << 1) & mask;
        if (fb) crc ^= poly;
    }
    return crc;
}

size_t anm_oracle_frame_symbols(const anm_config_t *cfg, const uint8_t *payload, size_t len, uint8_t *syms, size_t cap) {
    if (!cfg || !payload || !syms || len == 0 || len > cfg->max_payload) return 0;
    symw_t w = {syms, 0, cap, 0, 0, 0};
    while ((1u << w.b) < cfg->n_tones) ++w.b;
    if (cap < cfg->preamble_len) return 0;
    for (uint32_t p = 0; p < cfg->preamble_len; ++p) syms[w.n++] = cfg->preamble[p];
    /* header: LEN (16 bit, big endian) | CRC-8(poly 0x07, init 0) of the two LEN bytes */
    const uint32_t hi = (uint32_t)(len >> 8) & 0xffu, lo = (uint32_t)len & 0xffu;
    uint32_t c8 = crc_bits(crc_bits(0, hi, 8, 0x07), lo, 8, 0x07);
    if (put_byte(&w, hi) || put_byte(&w, lo) || put_byte(&w, c8) || flush_section(&w)) return 0;
    /* body: payload | CRC-16/CCITT-FALSE (poly 0x1021, init 0xFFFF) over LEN bytes + payload, big endian */
    uint32_t c16 = crc_bits(crc_bits(0xFFFF, hi, 16, 0x1021), lo, 16, 0x1021);
    for (size_t i = 0; i < len; ++i) {
        c16 = crc_bits(c16, payload[i], 16, 0x1021);
        if (put_byte(&w, payload[i])) return 0;
    }
    if (put_byte(&w, c16 >> 8) || put_byte(&w, c16 & 0xffu) || flush_section(&w)) return 0;
    return w.n;
}

/* ---- SPEC 6: integer transmitter ----------------------------------------------------------- */
typedef unsigned __int128 u128;
typedef __int128 i128;

static int64_t round_half_away_div(int64_t num, int64_t den) { /* den > 0 */
    return num >= 0 ? (num + den / 2) / den : -((-num + den / 2) / den); /* den even */
}

static uint64_t splitmix_finish(uint64_t z) {
    z ^= z >> 30;
    z *= 0xBF58476D1CE4E5B9ull;
    z ^= z >> 27;
    z *= 0x94D049BB133111EBull;
    z ^= z >> 31;
    return z;
}

int anm_oracle_tx_render(const anm_config_t *cfg, const uint8_t *program, size_t prog_len, const anm_tx_params_t *p,
                         uint64_t first_sample, int16_t *out, size_t n) {
    if (!cfg || !program || !prog_len || !p || (!out && n)) return -1;
    int16_t sine[1024]; /* SINE[i] = round(32767 sin(2 pi i / 1024)), half away from zero */
    for (int i = 0; i < 1024; ++i) sine[i] = (int16_t)lround(32767.0 * sin(6.283185307179586476925286766559 * i / 1024.0));
    const uint32_t N = cfg->sym_len;
    uint32_t lgN = 0;
    while ((1u << lgN) < N) ++lgN;
    /* step = 2^32 + round_half_away(ppm_x1000 * 4294.967296 / 1000) = 2^32 + round(ppm_x1000 * 2^32 / 1e9) */
    const int64_t step = ((int64_t)1 << 32) + round_half_away_div((int64_t)p->ppm_x1000 * 4294967296ll, 1000000000ll);
    uint32_t nscale = 0;
    if (p->snr_mdb != ANM_SNR_CLEAN) {
        /* nscale = round(amp / sqrt 2 / 10^(snr/20) / 53509.0 * 2^20) */
        double v = (double)p->amplitude_q15 / sqrt(2.0) / pow(10.0, (double)p->snr_mdb / 20000.0) / 53509.0 * 1048576.0;
        if (v > 4294967295.0) v = 4294967295.0;
        nscale = (uint32_t)floor(v + 0.5);
    }
    for (size_t i = 0; i < n; ++i) {
        const uint64_t nn = first_sample + i;
        const i128 tpos = (i128)p->start_offset * ((i128)1 << 32) + (i128)nn * (i128)step;
        int32_t sig = 0;
        if (tpos >= 0) {
            const u128 up = (u128)tpos;
            const uint64_t whole = (uint64_t)(up >> 32);
            const uint64_t sym = whole / N;
            const uint8_t entry = program[sym % prog_len];
            if (entry != ANM_SILENCE && entry < cfg->n_tones) {
                const u128 pos = up - (((u128)sym * N) << 32); /* Q32 position inside the symbol */
                const uint32_t phase = (uint32_t)(((u128)cfg->tone_bin[entry] * pos) >> lgN);
                sig = ((int32_t)p->amplitude_q15 * (int32_t)sine[phase >> 22]) >> 15;
            }
        }
        int32_t noise = 0;
        if (nscale) {
            int64_t sum = 0;
            for (uint64_t w = 0; w < 2; ++w) { /* four 16-bit uniforms per 64-bit word */
                uint64_t z = splitmix_finish(p->seed + (2 * nn + w) * 0x9E3779B97F4A7C15ull);
                for (int k = 0; k < 4; ++k, z >>= 16) sum += (int64_t)(z & 0xFFFFu);
            }
            noise = (int32_t)(((sum - 262140) * (int64_t)nscale) >> 20);
        }
        int32_t v = sig + noise;
        if (v > 32767) v = 32767;
        if (v < -32768) v = -32768;
        out[i] = (int16_t)v;
    }
    return 0;
}

/* threaded render of many channels (the CPU arm's signal source) */
typedef struct {
    const anm_config_t *cfg;
    const uint8_t *programs;
    size_t prog_stride;
    const uint32_t *prog_len;
    const anm_tx_params_t *params;
    uint32_t n_ch, tid, n_threads;
    uint64_t first;
    int16_t *out;
    size_t ch_stride, n;
} rjob_t;

static void *render_worker(void *arg) {
    rjob_t *j = (rjob_t *)arg;
    for (uint32_t c = j->tid; c < j->n_ch; c += j->n_threads)
        anm_oracle_tx_render(j->cfg, j->programs + (size_t)c * j->prog_stride, j->prog_len[c], &j->params[c], j->first,
                             j->out + (size_t)c * j->ch_stride, j->n);
    return NULL;
}

int anm_oracle_tx_render_batch(const anm_config_t *cfg, const uint8_t *programs, size_t prog_stride, const uint32_t *prog_len,
                               const anm_tx_params_t *params, uint32_t n_ch, uint64_t first_sample, int16_t *out,
                               size_t ch_stride, size_t n, uint32_t n_threads) {
    if (n_threads == 0) n_threads = 1;
    pthread_t *th = (pthread_t *)calloc(n_threads, sizeof *th);
    rjob_t *jobs = (rjob_t *)calloc(n_threads, sizeof *jobs);
    if (!th || !jobs) { free(th); free(jobs); return -1; }
    for (uint32_t t = 0; t < n_threads; ++t) {
        jobs[t] = (rjob_t){cfg, programs, prog_stride, prog_len, params, n_ch, t, n_threads, first_sample, out, ch_stride, n};
        pthread_create(&th[t], NULL, render_worker, &jobs[t]);
    }
    for (uint32_t t = 0; t < n_threads; ++t) pthread_join(th[t], NULL);
    free(th);
    free(jobs);
    return 0;
}
