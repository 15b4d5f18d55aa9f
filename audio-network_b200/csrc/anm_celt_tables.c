/*
 * anm_celt_tables.c -- static data of the CELT frame decoder (include/anmodem_opus.h, anm_celt_tables_t).  Host C.
 *
 * Two kinds of content:
 *  (1) NORMATIVE CONSTANTS of RFC 6716 that have no generating formula and are reproduced from the standard's reference
 *      decoder as shipped with the reference firmware (hardware/lib/libopus/src, BSD 3-clause, (c) Xiph.Org et al.):
 *        band edges of the 5 ms base layout            celt/modes.c:42-45   (eband5ms)
 *        the bit-allocation table                       celt/modes.c:50-63   (band_allocation)
 *        Laplace parameters of the coarse band energy   celt/quant_bands.c:77-138 (e_prob_model)
 *  (2) tables DERIVED from (1) by the standard's formulas, computed here instead of being carried as data, and checked
 *      against the reference's own static tables in tests/test_celt_entropy.py:
 *        logN          = log2 of the band widths in 1/8 bit      (celt/modes.c:296-300 compute_ebands / static logN400)
 *        pulse cache   = bits needed for K pulses in N samples   (celt/rate.c:73-140 compute_pulse_cache, cwrs.c:46-72 log2_frac)
 *        cache caps    = highest useful rate per band            (celt/rate.c:142-244)
 *        PVQ sizes     U(n, k) = U(n-1, k) + U(n, k-1) + U(n-1, k-1)   (celt/cwrs.c:75-207)
 *      and, for the synthesis (anm_celt_synth_tables_t; checked against the reference's static tables in tests/test_celt_synth.py):
 *        window        = sin(pi/2 sin^2(pi/2 (i + 1/2) / 120)), Q15      (celt/modes.c:379)
 *        MDCT twiddles = round(32768 cos(2 pi (i + 1/8) / N)), clamped      (celt/mdct.c:94, the form the static tables were dumped from)
 *        FFT twiddles  = (cos_norm(-i / 480), cos_norm(-i / 480 - 1/4))  (celt/kiss_fft.c:431-437 compute_twiddles, kf_cexp2)
 *        bit reversal  of the mixed-radix FFTs                           (celt/kiss_fft.c:329-359 compute_bitrev_table)
 *      The band mean energies eMeans (celt/quant_bands.c:44-50) are normative constants of kind (1).
 */
#include <math.h>
#include <string.h>

#include "anm_celt_entropy.h" /* cv_cos_norm for the synthesis twiddles */

#define NB ANM_CELT_BANDS
#define BITRES 3

static const int16_t k_ebands[NB + 1] = {0, 1, 2, 3, 4, 5, 6, 7, 8, 10, 12, 14, 16, 20, 24, 28, 34, 40, 48, 60, 78, 100};

static const uint8_t k_alloc[ANM_CELT_ALLOC_VECTORS * NB] = {
    0,   0,   0,   0,   0,   0,   0,   0,   0,   0,   0,   0,   0,   0,   0,   0,   0,   0,   0,   0,   0,
    90,  80,  75,  69,  63,  56,  49,  40,  34,  29,  20,  18,  10,  0,   0,   0,   0,   0,   0,   0,   0,
    110, 100, 90,  84,  78,  71,  65,  58,  51,  45,  39,  32,  26,  20,  12,  0,   0,   0,   0,   0,   0,
    118, 110, 103, 93,  86,  80,  75,  70,  65,  59,  53,  47,  40,  31,  23,  15,  4,   0,   0,   0,   0,
    126, 119, 112, 104, 95,  89,  83,  78,  72,  66,  60,  54,  47,  39,  32,  25,  17,  12,  1,   0,   0,
    134, 127, 120, 114, 103, 97,  91,  85,  78,  72,  66,  60,  54,  47,  41,  35,  29,  23,  16,  10,  1,
    144, 137, 130, 124, 113, 107, 101, 95,  88,  82,  76,  70,  64,  57,  51,  45,  39,  33,  26,  15,  1,
    152, 145, 138, 132, 123, 117, 111, 105, 98,  92,  86,  80,  74,  67,  61,  55,  49,  43,  36,  20,  1,
    162, 155, 148, 142, 133, 127, 121, 115, 108, 102, 96,  90,  84,  77,  71,  65,  59,  53,  46,  30,  1,
    172, 165, 158, 152, 143, 137, 131, 125, 118, 112, 106, 100, 94,  87,  81,  75,  69,  63,  56,  45,  20,
    200, 200, 200, 200, 200, 200, 200, 200, 198, 193, 188, 183, 178, 173, 168, 163, 158, 153, 148, 129, 104,
};

/* [frame size 120, 240, 480, 960][inter, intra][21 x {probability of 0, decay}], Q8 */
static const uint8_t k_e_prob[4 * 2 * 42] = {
    72,  127, 65,  129, 66,  128, 65,  128, 64,  128, 62,  128, 64,  128, 64,  128, 92,  78,  92,  79,  92,  78,  90,  79,  116, 41,  115, 40,
    114, 40,  132, 26,  132, 26,  145, 17,  161, 12,  176, 10,  177, 11,
    24,  179, 48,  138, 54,  135, 54,  132, 53,  134, 56,  133, 55,  132, 55,  132, 61,  114, 70,  96,  74,  88,  75,  88,  87,  74,  89,  66,
    91,  67,  100, 59,  108, 50,  120, 40,  122, 37,  97,  43,  78,  50,
    83,  78,  84,  81,  88,  75,  86,  74,  87,  71,  90,  73,  93,  74,  93,  74,  109, 40,  114, 36,  117, 34,  117, 34,  143, 17,  145, 18,
    146, 19,  162, 12,  165, 10,  178, 7,   189, 6,   190, 8,   177, 9,
    23,  178, 54,  115, 63,  102, 66,  98,  69,  99,  74,  89,  71,  91,  73,  91,  78,  89,  86,  80,  92,  66,  93,  64,  102, 59,  103, 60,
    104, 60,  117, 52,  123, 44,  138, 35,  133, 31,  97,  38,  77,  45,
    61,  90,  93,  60,  105, 42,  107, 41,  110, 45,  116, 38,  113, 38,  112, 38,  124, 26,  132, 27,  136, 19,  140, 20,  155, 14,  159, 16,
    158, 18,  170, 13,  177, 10,  187, 8,   192, 6,   175, 9,   159, 10,
    21,  178, 59,  110, 71,  86,  75,  85,  84,  83,  91,  66,  88,  73,  87,  72,  92,  75,  98,  72,  105, 58,  107, 54,  115, 52,  114, 55,
    112, 56,  129, 51,  132, 40,  150, 33,  140, 29,  98,  35,  77,  42,
    42,  121, 96,  66,  108, 43,  111, 40,  117, 44,  123, 32,  120, 36,  119, 33,  127, 33,  134, 34,  139, 21,  147, 23,  152, 20,  158, 25,
    154, 26,  166, 21,  173, 16,  184, 13,  184, 10,  150, 13,  139, 15,
    22,  178, 63,  114, 74,  82,  84,  83,  92,  82,  103, 62,  96,  72,  96,  67,  101, 73,  107, 72,  113, 55,  118, 52,  125, 52,  118, 52,
    117, 55,  135, 49,  137, 39,  157, 32,  145, 29,  97,  33,  77,  40,
};

static int ilog(uint32_t x) { return x ? 32 - __builtin_clz(x) : 0; }

/* ceil(log2(val) * 2^frac), the fixed-point squaring loop of the standard */
static int log2_frac(uint32_t val, int frac) {
    int l = ilog(val);
    if (val & (val - 1)) {
        if (l > 16) val = ((val - 1) >> (l - 16)) + 1;
        else val <<= 16 - l;
        l = (l - 1) << frac;
        do {
            int b = (int)(val >> 16);
            l += b << frac;
            val = (val + (uint32_t)b) >> b;
            val = (val * val + 0x7FFF) >> 15;
        } while (frac-- > 0);
        return l + (val > 0x8000);
    }
    return (l - 1) << frac;
}

static int get_pulses(int i) { return i < 8 ? i : (8 + (i & 7)) << ((i >> 3) - 1); }

/* U(n, k) for all n, k < LIM with 64-bit saturation; V(n, k) = U(n, k) + U(n, k + 1) */
#define LIM 260
static uint64_t g_u[LIM][LIM];
static const uint64_t SAT = (uint64_t)1 << 40;
static void build_u(void) {
    for (int n = 0; n < LIM; ++n)
        for (int k = 0; k < LIM; ++k) {
            uint64_t v;
            if (k == 0) v = 0;                    /* U(n, 0) = 0 */
            else if (n == 0) v = 0;               /* no dimensions, k > 0 pulses: nothing */
            else if (k == 1) v = 1;               /* U(n, 1) = 1 */
            else if (n == 1) v = 1;               /* U(1, k) = 1 */
            else v = g_u[n - 1][k] + g_u[n][k - 1] + g_u[n - 1][k - 1];
            g_u[n][k] = v > SAT ? SAT : v;
        }
}
static int fits32(int n, int k) { return k + 1 < LIM && n < LIM && g_u[n][k] + g_u[n][k + 1] <= 0xFFFFFFFFull; }

int anm_celt_tables_build(anm_celt_tables_t *t) {
    if (!t) return ANM_ERR_ARG;
    memset(t, 0, sizeof *t);
    memcpy(t->ebands, k_ebands, sizeof k_ebands);
    memcpy(t->alloc, k_alloc, sizeof k_alloc);
    memcpy(t->e_prob, k_e_prob, sizeof k_e_prob);
    build_u();
    for (int j = 0; j < NB; ++j) t->logn[j] = (int16_t)log2_frac((uint32_t)(k_ebands[j + 1] - k_ebands[j]), BITRES);
    for (int r = 0; r < ANM_CELT_PVQ_ROWS; ++r)
        for (int c = 0; c < ANM_CELT_PVQ_COLS; ++c) t->pvq_u[r * ANM_CELT_PVQ_COLS + c] = g_u[r][c] > 0xFFFFFFFFull ? 0xFFFFFFFFu : (uint32_t)g_u[r][c];
    /* ---- pulse cache: one entry list per distinct partition size N = band width << i >> 1, i = 0 .. maxLM + 1 ---- */
    const int LM = 3;
    int entryN[100], entryK[100], entryI[100], nb = 0, curr = 0;
    for (int i = 0; i <= LM + 1; ++i)
        for (int j = 0; j < NB; ++j) {
            const int N = (k_ebands[j + 1] - k_ebands[j]) << i >> 1;
            int found = -1;
            for (int k = 0; k <= i && found < 0; ++k)
                for (int n = 0; n < NB && (k != i || n < j); ++n)
                    if (N == (k_ebands[n + 1] - k_ebands[n]) << k >> 1) {
                        found = t->cache_index[k * NB + n];
                        break;
                    }
            t->cache_index[i * NB + j] = (int16_t)found;
            if (found == -1 && N != 0) {
                int K = 0;
                while (fits32(N, get_pulses(K + 1)) && K < 40) K++; /* MAX_PSEUDO */
                entryN[nb] = N;
                entryK[nb] = K;
                entryI[nb] = curr;
                t->cache_index[i * NB + j] = (int16_t)curr;
                curr += K + 1;
                nb++;
            }
        }
    if (curr > (int)sizeof t->cache_bits) return ANM_ERR_NOMEM;
    t->cache_size = (uint16_t)curr;
    for (int e = 0; e < nb; ++e) {
        uint8_t *ptr = t->cache_bits + entryI[e];
        const int N = entryN[e];
        for (int j = 1; j <= entryK[e]; ++j) {
            const int k = get_pulses(j);
            const int bits = N == 1 ? 1 << BITRES : log2_frac((uint32_t)(g_u[N][k] + g_u[N][k + 1]), BITRES);
            ptr[j] = (uint8_t)(bits - 1);
        }
        ptr[0] = (uint8_t)entryK[e];
    }
    /* ---- caps: the highest rate each band can usefully take, per LM and channel count ---- */
    uint8_t *cap = t->cache_caps;
    for (int i = 0; i <= LM; ++i)
        for (int C = 1; C <= 2; ++C)
            for (int j = 0; j < NB; ++j) {
                int N0 = k_ebands[j + 1] - k_ebands[j], max_bits;
                if (N0 << i == 1) {
                    max_bits = C * (1 + 8) << BITRES; /* a sign bit and MAX_FINE_BITS */
                } else {
                    int LM0 = 0;
                    if (N0 > 2) { N0 >>= 1; LM0--; }       /* even bands above 2 can be split once more */
                    else if (N0 <= 1) { LM0 = i < 1 ? i : 1; N0 <<= LM0; } /* N0 = 1 cannot go below N = 2 */
                    const uint8_t *pc = t->cache_bits + t->cache_index[(LM0 + 1) * NB + j];
                    max_bits = pc[pc[0]] + 1; /* the lowest-level PVQ of a fully split band */
                    int N = N0;
                    for (int k = 0; k < i - LM0; ++k) { /* regular splits */
                        max_bits <<= 1;
                        const int offset = ((t->logn[j] + ((LM0 + k) << BITRES)) >> 1) - 4; /* QTHETA_OFFSET */
                        const int32_t num = 459 * (int32_t)((2 * N - 1) * offset + max_bits);
                        const int32_t den = ((int32_t)(2 * N - 1) << 9) - 459;
                        int qb = (num + (den >> 1)) / den;
                        if (qb > 57) qb = 57;
                        max_bits += qb;
                        N <<= 1;
                    }
                    if (C == 2) { /* the stereo split */
                        max_bits <<= 1;
                        const int offset = ((t->logn[j] + (i << BITRES)) >> 1) - (N == 2 ? 16 : 4);
                        const int ndof = 2 * N - 1 - (N == 2);
                        const int32_t num = (N == 2 ? 512 : 487) * (int32_t)(max_bits + ndof * offset);
                        const int32_t den = ((int32_t)ndof << 9) - (N == 2 ? 512 : 487);
                        int qb = (num + (den >> 1)) / den;
                        const int lim = N == 2 ? 64 : 61;
                        if (qb > lim) qb = lim;
                        max_bits += qb;
                    }
                    /* fine energy bits */
                    const int ndof = C * N + ((C == 2 && N > 2) ? 1 : 0);
                    int offset = ((t->logn[j] + (i << BITRES)) >> 1) - 21; /* FINE_OFFSET */
                    if (N == 2) offset += 1 << BITRES >> 2;
                    const int32_t num = max_bits + ndof * offset;
                    const int32_t den = (ndof - 1) << BITRES;
                    int qb = (num + (den >> 1)) / den;
                    if (qb > 8) qb = 8;
                    max_bits += C * qb << BITRES;
                }
                max_bits = (4 * max_bits / (C * ((k_ebands[j + 1] - k_ebands[j]) << i))) - 64;
                *cap++ = (uint8_t)max_bits;
            }
    return ANM_OK;
}


/* ---------------------------------------------------------------- synthesis tables */
static void bitrev_fill(int fout, int16_t *f, int fstride, const int8_t *factors) {
    const int p = *factors++, m = *factors++;
    if (m == 1) {
        for (int j = 0; j < p; j++) {
            *f = (int16_t)(fout + j);
            f += fstride;
        }
    } else {
        for (int j = 0; j < p; j++) {
            bitrev_fill(fout, f, fstride * p, factors);
            f += fstride;
            fout += m;
        }
    }
}

int anm_celt_synth_tables_build(anm_celt_synth_tables_t *t) {
    static const int8_t k_emeans[25] = {103, 100, 92, 85, 81, 77, 72, 70, 78, 75, 73, 71, 78, 74, 69, 72, 70, 74, 76, 71, 60, 60, 60, 60, 60};
    /* radix / remaining length of every stage of the 480, 240, 120 and 60-point transforms (celt/static_modes_fixed.h:432-498; what kf_factor gives) */
    static const int8_t k_factors[4][10] = {{5, 96, 3, 32, 4, 8, 2, 4, 4, 1}, {5, 48, 3, 16, 4, 4, 4, 1, 0, 0}, {5, 24, 3, 8, 2, 4, 4, 1, 0, 0}, {5, 12, 3, 4, 4, 1, 0, 0, 0, 0}};
    const double pi = 3.14159265358979323846; /* M_PI */
    if (!t) return ANM_ERR_ARG;
    memset(t, 0, sizeof *t);
    for (int i = 0; i < 120; i++) {
        const double s = sin(.5 * pi * (i + .5) / 120);
        const double w = floor(.5 + 32768. * sin(.5 * pi * s * s));
        t->window[i] = (int16_t)(w > 32767 ? 32767 : w);
    }
    {
        int16_t *trig = t->trig;
        int N = 1920, N2 = 960;
        for (int shift = 0; shift <= 3; shift++) {
            for (int i = 0; i < N2; i++) { /* the static tables of the reference were dumped from the floating-point formula (celt/mdct.c:94) */
                double v = floor(.5 + 32768 * cos(2 * pi * (i + .125) / N));
                v = v > 32767 ? 32767 : v < -32767 ? -32767 : v;
                trig[i] = (int16_t)v;
            }
            trig += N2;
            N2 >>= 1;
            N >>= 1;
        }
    }
    for (int i = 0; i < 480; i++) {
        const int32_t phase = (int32_t)((uint32_t)(-i) << 17) / 480; /* DIV32(SHL32(phase, 17), nfft): C division, towards zero */
        t->fft_tw[2 * i] = cv_cos_norm(phase);
        t->fft_tw[2 * i + 1] = cv_cos_norm(phase - 32768);
    }
    {
        int16_t *b = t->bitrev;
        for (int k = 0; k < 4; k++) {
            bitrev_fill(0, b, 1, k_factors[k]);
            b += 480 >> k;
        }
    }
    memcpy(t->e_means, k_emeans, sizeof k_emeans);
    return ANM_OK;
}


/* ---------------------------------------------------------------- parse records -> frame jobs (host glue of the receive chain) */
long anm_celt_jobs_from_packets(const anm_pb_span_t *spans, const anm_opus_packet_t *packets, size_t n, anm_celt_job_t *jobs, size_t cap, uint32_t flags,
                                uint32_t *first_job) {
    if ((!spans || !packets) && n) return ANM_ERR_ARG;
    if (!jobs && cap) return ANM_ERR_ARG;
    size_t nj = 0;
    for (size_t i = 0; i < n; ++i) {
        const anm_opus_packet_t *pk = &packets[i];
        if (first_job) first_job[i] = 0xFFFFFFFFu;
        if (pk->count <= 0 || pk->count > 48 || pk->mode != ANM_OPUS_MODE_CELT_ONLY) continue;
        int lm;
        switch (pk->samples_per_frame) { /* at 48 kHz */
            case 120: lm = 0; break;
            case 240: lm = 1; break;
            case 480: lm = 2; break;
            case 960: lm = 3; break;
            default: continue;
        }
        int end;
        switch (pk->bandwidth) { /* opus_decoder.c:473-488 */
            case 1101: end = 13; break;
            case 1102:
            case 1103: end = 17; break;
            case 1104: end = 19; break;
            case 1105: end = 21; break;
            default: continue;
        }
        if (first_job) first_job[i] = (uint32_t)nj;
        uint32_t off = spans[i].audio_offset + (uint32_t)pk->payload_offset;
        for (int f = 0; f < pk->count; ++f) {
            if (nj < cap) {
                anm_celt_job_t *j = &jobs[nj];
                j->offset = off;
                j->len = (uint32_t)pk->size[f];
                j->channels = pk->channels;
                j->lm = (uint8_t)lm;
                j->end_band = (uint8_t)end;
                j->flags = (uint8_t)flags;
            }
            off += (uint32_t)pk->size[f];
            nj++;
        }
    }
    return (long)nj;
}
