/*
 * anm_celt_entropy.h -- stage 1 of the batched CELT frame decoder: everything the reference's
 * celt_decode_with_ec() reads off the range coder, for one frame, with no spectrum arithmetic.  Host + device code (the CUDA
 * kernel of anm_celt_gpu.cu runs it one thread per stream; tests/native compiles the same header for the host as a CHECKER
 * of the logic -- the product has no CPU path).
 *
 * TRANSCRIPTION NOTICE.  The Opus bit stream is defined by its reference decoder (RFC 6716: "the reference implementation
 * ... is the normative part"): how many bits every symbol consumes is fixed by these exact integer rules, and the test for
 * this file is that the range coder's final state equals OPUS_GET_FINAL_RANGE of the reference's libopus 1.3.1 for every
 * frame.  So the control flow below follows the reference function by function; it is a restatement of (c) Xiph.Org / Skype /
 * Octasic / Jean-Marc Valin / Timothy B. Terriberry / CSIRO / Gregory Maxwell code (BSD 3-clause, hardware/lib/libopus/COPYING):
 *   range decoder (RFC 6716 4.1)            celt/entdec.c:91-245, celt/entcode.c:91-117 (ec_tell_frac), entcode.h:111
 *   Laplace decoder                          celt/laplace.c:42-48, 93-134
 *   frame header, tf, spread, dynalloc, trim celt/celt_decoder.c:441-478, 946-1058
 *   coarse / fine / final band energies      celt/quant_bands.c:427-541 (fixed-point build, DB_SHIFT 10)
 *   bit allocation                           celt/rate.c:248-644, celt/celt.c:272-281 (init_caps)
 *   band loop and band splitting             celt/bands.c:647-676 (compute_qn), 705-906 (compute_theta), 908-945,
 *                                            953-1123 (quant_partition), 1127-1238 (quant_band), 1242-1389 (stereo),
 *                                            1405-1672 (quant_all_bands); celt/rate.h:48-87; celt/mathops.c:43-66
 *   PVQ codeword size V(N, K)                celt/cwrs.c:75-207 (the recurrence; the table itself is rebuilt in anm_celt_tables.c)
 * What is NOT here: anything that touches the normalised spectrum (alg_unquant's vector, folding, collapse masks, the
 * Hadamard / Haar reorderings, stereo merge) -- none of it influences how many bits are read; that is stage 2.
 */
#ifndef ANM_CELT_ENTROPY_H_INCLUDED
#define ANM_CELT_ENTROPY_H_INCLUDED

#include <stdint.h>

#include "../../include/anmodem_opus.h"

#ifdef __CUDACC__
#define ANM_CE_FN __host__ __device__ static inline
#else
#define ANM_CE_FN static inline
#endif

#ifdef __CUDACC__
#define ANM_CE_NOUNROLL _Pragma("unroll 1")
#else
#define ANM_CE_NOUNROLL
#endif

#define ANM_CE_BITRES 3
#define ANM_CE_NB 21 /* bands of the 48 kHz standard mode */

/* ---------------------------------------------------------------- range decoder */
typedef struct anm_ec {
    const uint8_t *bytes;
    uint32_t mask, base, storage;
    uint32_t offs, end_offs, end_window;
    int nend_bits, nbits_total;
    uint32_t rng, val, ext;
    int rem, error;
} anm_ec_t;

ANM_CE_FN int anm_ce_ilog(uint32_t x) { /* EC_ILOG: bits needed for x > 0 */
#ifdef __CUDA_ARCH__
    return 32 - __clz((int)x);
#else
    return 32 - __builtin_clz(x);
#endif
}
ANM_CE_FN int ce_read_byte(anm_ec_t *d) { return d->offs < d->storage ? d->bytes[(d->base + d->offs++) & d->mask] : 0; }
ANM_CE_FN int ce_read_byte_end(anm_ec_t *d) { return d->end_offs < d->storage ? d->bytes[(d->base + d->storage - ++d->end_offs) & d->mask] : 0; }
ANM_CE_FN void ce_normalize(anm_ec_t *d) {
    while (d->rng <= 0x800000u) { /* EC_CODE_BOT */
        d->nbits_total += 8;
        d->rng <<= 8;
        int sym = d->rem;
        d->rem = ce_read_byte(d);
        sym = (sym << 8 | d->rem) >> 1; /* EC_SYM_BITS - EC_CODE_EXTRA = 1 */
        d->val = ((d->val << 8) + (255u & ~(uint32_t)sym)) & 0x7FFFFFFFu;
    }
}
ANM_CE_FN void ce_init(anm_ec_t *d, const uint8_t *bytes, uint32_t mask, uint32_t base, uint32_t len) {
    d->bytes = bytes;
    d->mask = mask;
    d->base = base;
    d->storage = len;
    d->end_offs = 0;
    d->end_window = 0;
    d->nend_bits = 0;
    d->nbits_total = 9; /* EC_CODE_BITS + 1 - ((EC_CODE_BITS - EC_CODE_EXTRA) / EC_SYM_BITS) * EC_SYM_BITS */
    d->offs = 0;
    d->rng = 1u << 7;
    d->rem = ce_read_byte(d);
    d->val = d->rng - 1 - (uint32_t)(d->rem >> 1);
    d->error = 0;
    d->ext = 0;
    ce_normalize(d);
}
ANM_CE_FN int ce_tell(const anm_ec_t *d) { return d->nbits_total - anm_ce_ilog(d->rng); }
ANM_CE_FN uint32_t ce_tell_frac(const anm_ec_t *d) {
    const uint32_t corr[8] = {35733, 38967, 42495, 46340, 50535, 55109, 60097, 65535};
    const uint32_t nbits = (uint32_t)d->nbits_total << ANM_CE_BITRES;
    int l = anm_ce_ilog(d->rng);
    const uint32_t r = d->rng >> (l - 16);
    uint32_t b = (r >> 12) - 8;
    b += r > corr[b];
    l = (l << 3) + (int)b;
    return nbits - (uint32_t)l;
}
ANM_CE_FN uint32_t ce_decode(anm_ec_t *d, uint32_t ft) {
    d->ext = d->rng / ft;
    const uint32_t s = d->val / d->ext;
    return ft - (s + 1 < ft ? s + 1 : ft);
}
ANM_CE_FN uint32_t ce_decode_bin(anm_ec_t *d, uint32_t bits) {
    d->ext = d->rng >> bits;
    const uint32_t s = d->val / d->ext;
    return (1u << bits) - (s + 1u < (1u << bits) ? s + 1u : (1u << bits));
}
ANM_CE_FN void ce_update(anm_ec_t *d, uint32_t fl, uint32_t fh, uint32_t ft) {
    const uint32_t s = d->ext * (ft - fh);
    d->val -= s;
    d->rng = fl > 0 ? d->ext * (fh - fl) : d->rng - s;
    ce_normalize(d);
}
ANM_CE_FN int ce_bit_logp(anm_ec_t *d, uint32_t logp) {
    const uint32_t r = d->rng, v = d->val, s = r >> logp;
    const int ret = v < s;
    if (!ret) d->val = v - s;
    d->rng = ret ? s : r - s;
    ce_normalize(d);
    return ret;
}
ANM_CE_FN int ce_icdf(anm_ec_t *d, const uint8_t *icdf, uint32_t ftb) {
    uint32_t s = d->rng, t;
    const uint32_t v = d->val, r = s >> ftb;
    int ret = -1;
    do {
        t = s;
        s = r * icdf[++ret];
    } while (v < s);
    d->val = v - s;
    d->rng = t - s;
    ce_normalize(d);
    return ret;
}
ANM_CE_FN uint32_t ce_bits(anm_ec_t *d, uint32_t bits) {
    uint32_t window = d->end_window;
    int available = d->nend_bits;
    if ((uint32_t)available < bits) {
        do {
            window |= (uint32_t)ce_read_byte_end(d) << available;
            available += 8;
        } while (available <= 32 - 8);
    }
    const uint32_t ret = window & ((1u << bits) - 1u);
    window >>= bits;
    available -= (int)bits;
    d->end_window = window;
    d->nend_bits = available;
    d->nbits_total += (int)bits;
    return ret;
}
ANM_CE_FN uint32_t ce_uint(anm_ec_t *d, uint32_t ft) { /* ft > 1 */
    ft--;
    int ftb = anm_ce_ilog(ft);
    if (ftb > 8) {
        ftb -= 8;
        const uint32_t f = (ft >> ftb) + 1;
        const uint32_t s = ce_decode(d, f);
        ce_update(d, s, s + 1, f);
        const uint32_t t = s << ftb | ce_bits(d, (uint32_t)ftb);
        if (t <= ft) return t;
        d->error = 1;
        return ft;
    }
    ft++;
    const uint32_t s = ce_decode(d, ft);
    ce_update(d, s, s + 1, ft);
    return s;
}
ANM_CE_FN int ce_laplace(anm_ec_t *d, uint32_t fs, int decay) {
    int val = 0;
    uint32_t fl = 0;
    const uint32_t fm = ce_decode_bin(d, 15);
    if (fm >= fs) {
        val++;
        fl = fs;
        fs = ((32768u - 32u - fs) * (uint32_t)(16384 - decay) >> 15) + 1u; /* ec_laplace_get_freq1 + LAPLACE_MINP */
        while (fs > 1u && fm >= fl + 2 * fs) {
            fs *= 2;
            fl += fs;
            fs = ((fs - 2u) * (uint32_t)decay) >> 15;
            fs += 1u;
            val++;
        }
        if (fs <= 1u) {
            const int di = (int)((fm - fl) >> 1);
            val += di;
            fl += 2u * (uint32_t)di;
        }
        if (fm < fl + fs) val = -val;
        else fl += fs;
    }
    ce_update(d, fl, (fl + fs < 32768u ? fl + fs : 32768u), 32768u);
    return val;
}

/* ---------------------------------------------------------------- small math */
ANM_CE_FN int ce_frac_mul16(int a, int b) { return (16384 + ((int32_t)(int16_t)a * (int16_t)b)) >> 15; }
ANM_CE_FN int ce_bitexact_cos(int x) {
    const int32_t tmp = (4096 + ((int32_t)x * x)) >> 13;
    int x2 = (int16_t)tmp;
    x2 = (int16_t)((32767 - x2) + ce_frac_mul16(x2, (-7651 + ce_frac_mul16(x2, (8277 + ce_frac_mul16(-626, x2))))));
    return 1 + x2;
}
ANM_CE_FN int ce_bitexact_log2tan(int isin, int icos) {
    const int lc = anm_ce_ilog((uint32_t)icos), ls = anm_ce_ilog((uint32_t)isin);
    icos <<= 15 - lc;
    isin <<= 15 - ls;
    return (ls - lc) * (1 << 11) + ce_frac_mul16(isin, ce_frac_mul16(isin, -2597) + 7932) - ce_frac_mul16(icos, ce_frac_mul16(icos, -2597) + 7932);
}
ANM_CE_FN uint32_t ce_isqrt32(uint32_t v) {
    uint32_t g = 0;
    int bshift = (anm_ce_ilog(v) - 1) >> 1;
    uint32_t b = 1u << bshift;
    do {
        const uint32_t t = ((g << 1) + b) << bshift;
        if (t <= v) {
            g += b;
            v -= t;
        }
        b >>= 1;
        bshift--;
    } while (bshift >= 0);
    return g;
}
ANM_CE_FN int ce_imin(int a, int b) { return a < b ? a : b; }
ANM_CE_FN int ce_imax(int a, int b) { return a > b ? a : b; }
ANM_CE_FN int ce_pshr32(int32_t a, int s) { return (a + (1 << (s - 1))) >> s; }

/* ---------------------------------------------------------------- pulse cache / PVQ sizes */
ANM_CE_FN const uint8_t *ce_cache(const anm_celt_tables_t *t, int band, int LM) { return t->cache_bits + t->cache_index[(LM + 1) * ANM_CE_NB + band]; }
ANM_CE_FN int ce_get_pulses(int i) { return i < 8 ? i : (8 + (i & 7)) << ((i >> 3) - 1); }
ANM_CE_FN int ce_bits2pulses(const anm_celt_tables_t *t, int band, int LM, int bits) {
    const uint8_t *cache = ce_cache(t, band, LM);
    int lo = 0, hi = cache[0];
    bits--;
    for (int i = 0; i < 6; ++i) {
        const int mid = (lo + hi + 1) >> 1;
        if ((int)cache[mid] >= bits) hi = mid;
        else lo = mid;
    }
    if (bits - (lo == 0 ? -1 : (int)cache[lo]) <= (int)cache[hi] - bits) return lo;
    return hi;
}
ANM_CE_FN int ce_pulses2bits(const anm_celt_tables_t *t, int band, int LM, int pulses) {
    const uint8_t *cache = ce_cache(t, band, LM);
    return pulses == 0 ? 0 : cache[pulses] + 1;
}
/* U(n, k) of the PVQ codebook: symmetric, stored for min(n, k) <= 14 */
ANM_CE_FN uint32_t ce_pvq_u(const anm_celt_tables_t *t, int n, int k) {
    const int r = n < k ? n : k, c = n < k ? k : n;
    return t->pvq_u[r * ANM_CELT_PVQ_COLS + c];
}

#include "anm_celt_vec.h"

/* ---------------------------------------------------------------- band splitting (stage 1: bits only; stage 2: with the spectrum) */
/* per-frame working storage of stage 2 (NULL in a ce_band_ctx_t = stage 1: no spectrum arithmetic at all) */
typedef struct ce_spec {
    int16_t *norm;         /* [2 * 624] folding source: the decoded bands so far, scaled by sqrt(N) (two channels for dual stereo) */
    int16_t *tmp;          /* [176] Hadamard reordering scratch */
    int *iy;               /* [176] pulse vector of one partition */
    int16_t *band;         /* [3 * 176] or NULL: the band being decoded (two channels) and the folding scratch, when they are to live apart from the
                              output (thread-local on the GPU); the finished band is then copied out */
    uint32_t seed;         /* noise generator (CELTDecoder.rng) */
    int spread, disable_inv;
    int lane, nl;          /* this thread's share of the loops over coefficients (anm_celt_vec.h); the arrays above are shared by the nl lanes */
} ce_spec_t;
#define CE_SPEC_NORM 1248
#define CE_SPEC_TMP 176
#define CE_SPEC_IY 176
#define CE_SPEC_BAND (3 * 176)

typedef struct ce_band_ctx {
    const anm_celt_tables_t *t;
    anm_ec_t *ec;
    int i, intensity, tf_change;
    int32_t remaining_bits;
    anm_celt_frame_t *out; /* pulse statistics */
    ce_spec_t *sp;
} ce_band_ctx_t;

typedef struct ce_split {
    int inv, imid, iside, delta, itheta, qalloc;
} ce_split_t;

ANM_CE_FN int ce_compute_qn(int N, int b, int offset, int pulse_cap, int stereo) {
    const int16_t exp2_table8[8] = {16384, 17866, 19483, 21247, 23170, 25267, 27554, 30048};
    int qn, qb;
    int N2 = 2 * N - 1;
    if (stereo && N == 2) N2--;
    qb = (b + N2 * offset) / N2; /* celt_sudiv: C division, truncating */
    qb = ce_imin(b - pulse_cap - (4 << ANM_CE_BITRES), qb);
    qb = ce_imin(8 << ANM_CE_BITRES, qb);
    if (qb < (1 << ANM_CE_BITRES >> 1)) {
        qn = 1;
    } else {
        qn = exp2_table8[qb & 0x7] >> (14 - (qb >> ANM_CE_BITRES));
        qn = (qn + 1) >> 1 << 1;
    }
    return qn;
}

ANM_CE_FN void ce_compute_theta(ce_band_ctx_t *ctx, ce_split_t *s, int N, int *b, int B, int B0, int LM, int stereo, int *fill) {
    anm_ec_t *ec = ctx->ec;
    int itheta = 0, inv = 0, imid, iside, delta;
    const int pulse_cap = ctx->t->logn[ctx->i] + LM * (1 << ANM_CE_BITRES);
    const int offset = (pulse_cap >> 1) - (stereo && N == 2 ? 16 : 4); /* QTHETA_OFFSET_TWOPHASE : QTHETA_OFFSET */
    int qn = ce_compute_qn(N, *b, offset, pulse_cap, stereo);
    if (stereo && ctx->i >= ctx->intensity) qn = 1;
    const int32_t tell = (int32_t)ce_tell_frac(ec);
    if (qn != 1) {
        if (stereo && N > 2) {
            /* step pdf: probability p0 up to qn / 2, 1 after */
            const int p0 = 3, x0 = qn / 2, ft = p0 * (x0 + 1) + x0;
            const int fs = (int)ce_decode(ec, (uint32_t)ft);
            int x;
            if (fs < (x0 + 1) * p0) x = fs / p0;
            else x = x0 + 1 + (fs - (x0 + 1) * p0);
            ce_update(ec, (uint32_t)(x <= x0 ? p0 * x : (x - 1 - x0) + (x0 + 1) * p0), (uint32_t)(x <= x0 ? p0 * (x + 1) : (x - x0) + (x0 + 1) * p0), (uint32_t)ft);
            itheta = x;
        } else if (B0 > 1 || stereo) {
            itheta = (int)ce_uint(ec, (uint32_t)qn + 1u); /* uniform pdf */
        } else {
            /* triangular pdf */
            const int ft = ((qn >> 1) + 1) * ((qn >> 1) + 1);
            const int fm = (int)ce_decode(ec, (uint32_t)ft);
            int fs, fl;
            if (fm < ((qn >> 1) * ((qn >> 1) + 1) >> 1)) {
                itheta = (int)((ce_isqrt32(8u * (uint32_t)fm + 1u) - 1u) >> 1);
                fs = itheta + 1;
                fl = itheta * (itheta + 1) >> 1;
            } else {
                itheta = (int)((2u * (uint32_t)(qn + 1) - ce_isqrt32(8u * (uint32_t)(ft - fm - 1) + 1u)) >> 1);
                fs = qn + 1 - itheta;
                fl = ft - ((qn + 1 - itheta) * (qn + 2 - itheta) >> 1);
            }
            ce_update(ec, (uint32_t)fl, (uint32_t)(fl + fs), (uint32_t)ft);
        }
        itheta = (int)(((uint32_t)itheta * 16384u) / (uint32_t)qn);
    } else if (stereo) {
        if (*b > 2 << ANM_CE_BITRES && ctx->remaining_bits > 2 << ANM_CE_BITRES) inv = ce_bit_logp(ec, 2);
        else inv = 0;
        if (ctx->sp && ctx->sp->disable_inv) inv = 0; /* a mono decoder never inverts (downmix) */
        itheta = 0;
    }
    const int qalloc = (int)((int32_t)ce_tell_frac(ec) - tell);
    *b -= qalloc;
    if (itheta == 0) {
        imid = 32767;
        iside = 0;
        *fill &= (1 << B) - 1;
        delta = -16384;
    } else if (itheta == 16384) {
        imid = 0;
        iside = 32767;
        *fill &= ((1 << B) - 1) << B;
        delta = 16384;
    } else {
        imid = ce_bitexact_cos((int16_t)itheta);
        iside = ce_bitexact_cos((int16_t)(16384 - itheta));
        delta = ce_frac_mul16((N - 1) << 7, ce_bitexact_log2tan(iside, imid));
    }
    s->inv = inv;
    s->imid = imid;
    s->iside = iside;
    s->delta = delta;
    s->itheta = itheta;
    s->qalloc = qalloc;
}

/* a band of one coefficient: its sign */
ANM_CE_FN unsigned ce_band_n1(ce_band_ctx_t *ctx, int16_t *X, int16_t *Y, int stereo, int16_t *lowband_out) {
    int16_t *x = X;
    for (int c = 0; c < 1 + stereo; ++c) {
        int sign = 0;
        if (ctx->remaining_bits >= 1 << ANM_CE_BITRES) {
            sign = (int)ce_bits(ctx->ec, 1);
            ctx->remaining_bits -= 1 << ANM_CE_BITRES;
        }
        if (ctx->sp && ctx->sp->lane == 0) x[0] = sign ? -16384 : 16384; /* NORM_SCALING */
        x = Y;
    }
    if (ctx->sp) {
        if (ctx->sp->lane == 0 && lowband_out) lowband_out[0] = (int16_t)(X[0] >> 4);
        CV_SYNC();
    }
    return 1;
}

/* a mono partition: splits in two while the budget exceeds what one codeword can carry, then reads the PVQ codeword; with a spectrum
 * (ctx->sp) the codeword becomes the partition's coefficients, an empty partition is folded from `lowband` or filled with noise.
 * Returns the collapse mask (which of the B interleaved blocks received energy).
 * The reference recurses (quant_partition, bands.c:915-1078); here the tree is walked depth first with the pending second halves on a
 * small stack (a partition splits at most four times: LM 3 -> -1), so that the work at the leaves -- codeword to coefficients, the
 * expensive part -- sits at ONE place in the code: the threads of a warp, each on its own frame, go through it together whatever the
 * shape of their trees, where a recursion runs them one call path after the other. */
typedef struct ce_pending { /* the half of a split partition that waits for the other half's subtree */
    int16_t *X, *lowband;
    int N, b, B, LM, fill, shift;
    int first_bits, adj_ok; /* what the first half was given; whether its leftovers may move over (itheta not at its end) */
    int32_t rebalance0;     /* the budget when the first half started */
    int16_t gain;
} ce_pending_t;

ANM_CE_FN unsigned ce_partition(ce_band_ctx_t *ctx, int16_t *X, int N, int b, int B, int16_t *lowband, int LM, int16_t gain, int fill) {
    ce_spec_t *sp = ctx->sp;
    ce_pending_t stack[4];
    int depth = 0, shift = 0; /* shift: where this partition's blocks sit in the mask of the whole */
    unsigned cm_all = 0;
    for (;;) {
        for (;;) { /* down to a leaf, first halves first */
            const uint8_t *cache = ce_cache(ctx->t, ctx->i, LM);
            if (!(LM != -1 && b > cache[cache[0]] + 12 && N > 2)) break;
            ce_split_t s;
            const int B0 = B;
            N >>= 1;
            int16_t *Y = X ? X + N : X;
            LM -= 1;
            if (B == 1) fill = (fill & 1) | (fill << 1);
            B = (B + 1) >> 1;
            ce_compute_theta(ctx, &s, N, &b, B, B0, LM, 0, &fill);
            int delta = s.delta;
            const int itheta = s.itheta;
            const int16_t mid = (int16_t)s.imid, side = (int16_t)s.iside;
            /* more bits to low-energy MDCTs than they would otherwise deserve */
            if (B0 > 1 && (itheta & 0x3fff)) {
                if (itheta > 8192) delta -= delta >> (4 - LM);
                else delta = ce_imin(0, delta + (N << ANM_CE_BITRES >> (5 - LM)));
            }
            const int mbits = ce_imax(0, ce_imin(b, (b - delta) / 2));
            const int sbits = b - mbits;
            ctx->remaining_bits -= s.qalloc;
            int16_t *next_lowband2 = lowband ? lowband + N : lowband;
            const int16_t gm = (int16_t)CV_P15(gain, mid), gs = (int16_t)CV_P15(gain, side);
            ce_pending_t *p = &stack[depth++];
            p->N = N;
            p->B = B;
            p->LM = LM;
            p->rebalance0 = ctx->remaining_bits;
            if (mbits >= sbits) { /* the mid first; the side's blocks are the upper ones of the mask */
                p->X = Y;
                p->lowband = next_lowband2;
                p->b = sbits;
                p->gain = gs;
                p->fill = fill >> B;
                p->shift = shift + (B0 >> 1);
                p->first_bits = mbits;
                p->adj_ok = itheta != 0;
                b = mbits;
                gain = gm;
            } else {
                p->X = X;
                p->lowband = lowband;
                p->b = mbits;
                p->gain = gm;
                p->fill = fill;
                p->shift = shift;
                p->first_bits = sbits;
                p->adj_ok = itheta != 16384;
                X = Y;
                lowband = next_lowband2;
                b = sbits;
                gain = gs;
                fill >>= B;
                shift += B0 >> 1;
            }
        }
        unsigned cm = 0;
        int q = ce_bits2pulses(ctx->t, ctx->i, LM, b);
        int curr_bits = ce_pulses2bits(ctx->t, ctx->i, LM, q);
        ctx->remaining_bits -= curr_bits;
        while (ctx->remaining_bits < 0 && q > 0) { /* never bust the budget */
            ctx->remaining_bits += curr_bits;
            q--;
            curr_bits = ce_pulses2bits(ctx->t, ctx->i, LM, q);
            ctx->remaining_bits -= curr_bits;
        }
        if (q != 0) {
            const int K = ce_get_pulses(q);
            /* decode_pulses: one uniform symbol over the V(N, K) codewords */
            const uint32_t v = ce_pvq_u(ctx->t, N, K) + ce_pvq_u(ctx->t, N, K + 1);
            const uint32_t idx = ce_uint(ctx->ec, v);
            if (ctx->out) {
                ctx->out->pvq_codewords++;
                ctx->out->pvq_pulses += (uint32_t)K;
                ctx->out->pvq_index_xor ^= idx * 2654435761u + (uint32_t)(N * 131 + K);
            }
            if (sp) { /* alg_unquant */
                const int32_t Ryy = cv_cwrsi(ctx->t, N, K, idx, sp->iy, sp->lane);
                cv_normalise_residual(sp->iy, X, N, Ryy, gain, sp->lane, sp->nl);
                cv_exp_rotation_dec(X, N, B, K, sp->spread, sp->lane, sp->nl);
                cm = cv_collapse_mask(sp->iy, N, B, sp->lane, sp->nl);
                CV_SYNC(); /* iy is free for the next partition */
            }
        } else if (sp) {
            /* no pulse: fill the partition anyway */
            const unsigned cm_mask = (unsigned)(1UL << B) - 1;
            fill &= (int)cm_mask;
            if (!fill) {
                for (int j = sp->lane; j < N; j += sp->nl) X[j] = 0;
                CV_SYNC();
            } else {
                CV_SYNC();
                if (lowband == 0) { /* noise: every lane steps the generator, one stores */
                    for (int j = 0; j < N; j++) {
                        sp->seed = cv_lcg(sp->seed);
                        if (sp->lane == 0) X[j] = (int16_t)((int32_t)sp->seed >> 20);
                    }
                    cm = cm_mask;
                } else { /* folded spectrum, plus a little noise about 48 dB below it */
                    for (int j = 0; j < N; j++) {
                        sp->seed = cv_lcg(sp->seed);
                        const int16_t tmp = (sp->seed & 0x8000u) ? 4 : -4; /* QCONST16(1 / 256, 10) */
                        if (sp->lane == 0) X[j] = (int16_t)(lowband[j] + tmp);
                    }
                    cm = (unsigned)fill;
                }
                cv_renormalise(X, N, gain, sp->lane, sp->nl);
            }
        }
        cm_all |= cm << shift;
        if (depth == 0) break;
        /* on to the half that waited: what its sibling's subtree left over moves to it */
        const ce_pending_t *p = &stack[--depth];
        const int32_t rebalance = p->first_bits - (p->rebalance0 - ctx->remaining_bits);
        b = p->b;
        if (rebalance > 3 << ANM_CE_BITRES && p->adj_ok) b += rebalance - (3 << ANM_CE_BITRES);
        X = p->X;
        lowband = p->lowband;
        N = p->N;
        B = p->B;
        LM = p->LM;
        gain = p->gain;
        fill = p->fill;
        shift = p->shift;
    }
    return cm_all;
}

/* one band of one channel (or the mid / side of a stereo band): the time-frequency reshaping around the partition */
ANM_CE_FN unsigned ce_band(ce_band_ctx_t *ctx, int16_t *X, int N, int b, int B, int16_t *lowband, int LM, int16_t *lowband_out, int16_t gain,
                           int16_t *lowband_scratch, int fill) {
    const uint8_t bit_interleave_table[16] = {0, 1, 1, 1, 2, 3, 3, 3, 2, 3, 3, 3, 2, 3, 3, 3};
    const uint8_t bit_deinterleave_table[16] = {0x00, 0x03, 0x0C, 0x0F, 0x30, 0x33, 0x3C, 0x3F, 0xC0, 0xC3, 0xCC, 0xCF, 0xF0, 0xF3, 0xFC, 0xFF};
    ce_spec_t *sp = ctx->sp;
    const int N0 = N;
    int B0 = B, time_divide = 0, recombine = 0, k;
    int tf_change = ctx->tf_change;
    const int long_blocks = B0 == 1;
    if (N == 1) return ce_band_n1(ctx, X, 0, 0, lowband_out);
    int N_B = (int)((uint32_t)N / (uint32_t)B);
    if (tf_change > 0) recombine = tf_change;
    if (sp && lowband_scratch && lowband && (recombine || ((N_B & 1) == 0 && tf_change < 0) || B0 > 1)) {
        CV_SYNC();
        for (k = sp->lane; k < N; k += sp->nl) lowband_scratch[k] = lowband[k];
        CV_SYNC();
        lowband = lowband_scratch;
    }
    for (k = 0; k < recombine; k++) { /* band recombining: more frequency resolution */
        if (sp && lowband) cv_haar1(lowband, N >> k, 1 << k, sp->lane, sp->nl);
        fill = bit_interleave_table[fill & 0xF] | bit_interleave_table[fill >> 4] << 2;
    }
    B >>= recombine;
    N_B <<= recombine;
    while ((N_B & 1) == 0 && tf_change < 0) { /* more time resolution */
        if (sp && lowband) cv_haar1(lowband, N_B, B, sp->lane, sp->nl);
        fill |= fill << B;
        B <<= 1;
        N_B >>= 1;
        time_divide++;
        tf_change++;
    }
    B0 = B;
    const int N_B0 = N_B;
    /* time order instead of frequency order */
    if (sp && B0 > 1 && lowband) cv_deinterleave_hadamard(lowband, sp->tmp, N_B >> recombine, B0 << recombine, long_blocks, sp->lane, sp->nl);
    unsigned cm = ce_partition(ctx, X, N, b, B, lowband, LM, gain, fill);
    if (sp) {
        if (B0 > 1) cv_interleave_hadamard(X, sp->tmp, N_B >> recombine, B0 << recombine, long_blocks, sp->lane, sp->nl);
        /* undo the time-frequency changes */
        N_B = N_B0;
        B = B0;
        for (k = 0; k < time_divide; k++) {
            B >>= 1;
            N_B <<= 1;
            cm |= cm >> B;
            cv_haar1(X, N_B, B, sp->lane, sp->nl);
        }
        for (k = 0; k < recombine; k++) {
            cm = bit_deinterleave_table[cm];
            cv_haar1(X, N0 >> k, 1 << k, sp->lane, sp->nl);
        }
        B <<= recombine;
        if (lowband_out) { /* scaled for later folding */
            const int16_t n = (int16_t)cv_sqrt((int32_t)((uint32_t)N0 << 22));
            CV_SYNC();
            for (k = sp->lane; k < N0; k += sp->nl) lowband_out[k] = (int16_t)CV_Q15(n, X[k]);
            CV_SYNC();
        }
        cm &= (unsigned)(1 << B) - 1;
    }
    return cm;
}

/* the channels of one band: mono (C == 1), the two channels of a dual-stereo band coded one after the other (dual), or a mid / side pair
 * (quant_band_stereo, bands.c:1371-1489).  Whatever the case, it comes down to at most two mono codings -- and they all go through the ONE
 * ce_band below, in a loop that must not be unrolled: the threads of a warp, each on its own frame, then meet inside it whichever case their
 * band is (see ce_partition).  lb / lb2, lbo / lbo2: folding source and folding output per channel; x_cm / y_cm: in, the masks of the folding
 * source; out, the band's. */
typedef struct ce_mono {
    int16_t *X, *lowband, *lowband_out, *scratch;
    int b, fill;
    int16_t gain;
} ce_mono_t;

ANM_CE_FN void ce_band_channels(ce_band_ctx_t *ctx, int16_t *X, int16_t *Y, int C, int dual, int N, int b, int B, int16_t *lb, int16_t *lb2, int LM, int16_t *lbo,
                                int16_t *lbo2, int16_t *lowband_scratch, unsigned *x_cm, unsigned *y_cm) {
    ce_spec_t *sp = ctx->sp;
    ce_mono_t cur, nxt;
    int n = 0;       /* mono codings to do */
    int first_bits = 0, adj_ok = 0, rebalancing = 0;
    int32_t rebalance0 = 0;
    int stereo = 0, sign = 1, c = 0;
    ce_split_t s;
    s.inv = 0;
    s.imid = 0;
    s.iside = 0;
    s.itheta = 0;
    cur.X = X; cur.lowband = lb; cur.lowband_out = lbo; cur.scratch = lowband_scratch; cur.b = b; cur.fill = 0; cur.gain = 32767;
    nxt = cur;
    if (C == 1) {
        cur.fill = (int)(*x_cm | *y_cm);
        n = 1;
    } else if (dual) {
        cur.b = b / 2;
        cur.fill = (int)*x_cm;
        nxt.X = Y; nxt.lowband = lb2; nxt.lowband_out = lbo2; nxt.b = b / 2; nxt.fill = (int)*y_cm;
        n = 2;
    } else if (N == 1) {
        *x_cm = *y_cm = ce_band_n1(ctx, X, Y, 1, lbo);
        return;
    } else {
        stereo = 1;
        int fill = (int)(*x_cm | *y_cm);
        const int orig_fill = fill;
        ce_compute_theta(ctx, &s, N, &b, B, B, LM, 1, &fill);
        const int itheta = s.itheta;
        if (N == 2) {
            /* mid and side are orthogonal: one bit for the side's sign */
            int sbits = 0;
            if (itheta != 0 && itheta != 16384) sbits = 1 << ANM_CE_BITRES;
            c = itheta > 8192;
            ctx->remaining_bits -= s.qalloc + sbits;
            if (sbits) sign = 1 - 2 * (int)ce_bits(ctx->ec, 1);
            /* orig_fill: the side is folded even when itheta == 16384 cleared the low bits of fill */
            cur.X = c ? Y : X;
            cur.b = b - sbits;
            cur.fill = orig_fill;
            n = 1;
        } else {
            const int mbits = ce_imax(0, ce_imin(b, (b - s.delta) / 2));
            const int sbits = b - mbits;
            ctx->remaining_bits -= s.qalloc;
            rebalance0 = ctx->remaining_bits;
            rebalancing = 1;
            /* the mid keeps unit norm (it is the folding source of later bands); a stereo split never folds the side */
            ce_mono_t m = cur, sd = cur;
            m.b = mbits; m.fill = fill;
            sd.X = Y; sd.lowband = 0; sd.lowband_out = 0; sd.scratch = 0; sd.b = sbits; sd.fill = fill >> B; sd.gain = (int16_t)s.iside;
            if (mbits >= sbits) {
                cur = m; nxt = sd;
                first_bits = mbits;
                adj_ok = itheta != 0;
            } else {
                cur = sd; nxt = m;
                first_bits = sbits;
                adj_ok = itheta != 16384;
            }
            n = 2;
        }
    }
    unsigned cm0 = 0, cm1 = 0;
    ANM_CE_NOUNROLL
    for (int k = 0; k < n; k++) {
        if (k == 1 && rebalancing) { /* what the first coding left over goes to the second */
            const int32_t rebalance = first_bits - (rebalance0 - ctx->remaining_bits);
            if (rebalance > 3 << ANM_CE_BITRES && adj_ok) cur.b += rebalance - (3 << ANM_CE_BITRES);
        }
        const unsigned cm = ce_band(ctx, cur.X, N, cur.b, B, cur.lowband, LM, cur.lowband_out, cur.gain, cur.scratch, cur.fill);
        if (k == 0) cm0 = cm;
        else cm1 = cm;
        cur = nxt;
    }
    if (!stereo) {
        *x_cm = cm0;
        *y_cm = dual ? cm1 : cm0;
        return;
    }
    if (sp) {
        const int16_t mid = (int16_t)s.imid, side = (int16_t)s.iside;
        if (N == 2) {
            int16_t *x2 = c ? Y : X, *y2 = c ? X : Y;
            CV_SYNC();
            if (sp->lane == 0) {
                y2[0] = (int16_t)(-sign * x2[1]);
                y2[1] = (int16_t)(sign * x2[0]);
                X[0] = (int16_t)CV_Q15(mid, X[0]);
                X[1] = (int16_t)CV_Q15(mid, X[1]);
                Y[0] = (int16_t)CV_Q15(side, Y[0]);
                Y[1] = (int16_t)CV_Q15(side, Y[1]);
                int16_t tmp = X[0];
                X[0] = (int16_t)CV_S16(tmp, Y[0]);
                Y[0] = CV_A16(tmp, Y[0]);
                tmp = X[1];
                X[1] = (int16_t)CV_S16(tmp, Y[1]);
                Y[1] = CV_A16(tmp, Y[1]);
            }
            CV_SYNC();
        } else {
            cv_stereo_merge(X, Y, mid, N, sp->lane, sp->nl);
        }
        if (s.inv) {
            CV_SYNC();
            for (int j = sp->lane; j < N; j += sp->nl) Y[j] = (int16_t)-Y[j];
            CV_SYNC();
        }
    }
    *x_cm = *y_cm = cm0 | cm1;
}

/* ---------------------------------------------------------------- bit allocation */
ANM_CE_FN int ce_interp_bits2pulses(const anm_celt_tables_t *t, int start, int end, int skip_start, const int *bits1, const int *bits2, const int *thresh,
                                    const int *cap, int32_t total, int32_t *balance_out, int skip_rsv, int *intensity, int intensity_rsv,
                                    int *dual_stereo, int dual_stereo_rsv, int *bits, int *ebits, int *fine_priority, int C, int LM, anm_ec_t *ec) {
    const uint8_t log2_frac_table[24] = {0, 8, 13, 16, 19, 21, 23, 24, 26, 27, 28, 29, 30, 31, 32, 32, 33, 34, 34, 35, 36, 36, 37, 37};
    const int16_t *eb = t->ebands;
    int32_t psum;
    int lo = 0, hi = 1 << 6, j, codedBands, done;
    const int alloc_floor = C << ANM_CE_BITRES, stereo = C > 1, logM = LM << ANM_CE_BITRES;
    int32_t left, percoeff, balance;
    for (int i = 0; i < 6; ++i) { /* ALLOC_STEPS */
        const int mid = (lo + hi) >> 1;
        psum = 0;
        done = 0;
        for (j = end; j-- > start;) {
            const int tmp = bits1[j] + (int)(((int32_t)mid * bits2[j]) >> 6);
            if (tmp >= thresh[j] || done) {
                done = 1;
                psum += ce_imin(tmp, cap[j]);
            } else if (tmp >= alloc_floor) {
                psum += alloc_floor;
            }
        }
        if (psum > total) hi = mid;
        else lo = mid;
    }
    psum = 0;
    done = 0;
    for (j = end; j-- > start;) {
        int tmp = bits1[j] + (int)(((int32_t)lo * bits2[j]) >> 6);
        if (tmp < thresh[j] && !done) {
            tmp = tmp >= alloc_floor ? alloc_floor : 0;
        } else {
            done = 1;
        }
        tmp = ce_imin(tmp, cap[j]);
        bits[j] = tmp;
        psum += tmp;
    }
    /* which bands to skip, working backwards from the end */
    for (codedBands = end;; codedBands--) {
        j = codedBands - 1;
        if (j <= skip_start) {
            total += skip_rsv; /* the bit reserved to end skipping comes back */
            break;
        }
        left = total - psum;
        percoeff = (int32_t)((uint32_t)left / (uint32_t)(eb[codedBands] - eb[start]));
        left -= (eb[codedBands] - eb[start]) * percoeff;
        const int rem = ce_imax((int)left - (eb[j] - eb[start]), 0);
        const int band_width = eb[codedBands] - eb[j];
        int band_bits = (int)(bits[j] + percoeff * band_width + rem);
        if (band_bits >= ce_imax(thresh[j], alloc_floor + (1 << ANM_CE_BITRES))) {
            if (ce_bit_logp(ec, 1)) break;
            psum += 1 << ANM_CE_BITRES; /* a bit was used to skip this band */
            band_bits -= 1 << ANM_CE_BITRES;
        }
        psum -= bits[j] + intensity_rsv; /* reclaim what the band had */
        if (intensity_rsv > 0) intensity_rsv = log2_frac_table[j - start];
        psum += intensity_rsv;
        if (band_bits >= alloc_floor) {
            psum += alloc_floor;
            bits[j] = alloc_floor;
        } else {
            bits[j] = 0;
        }
    }
    if (intensity_rsv > 0) *intensity = start + (int)ce_uint(ec, (uint32_t)(codedBands + 1 - start));
    else *intensity = 0;
    if (*intensity <= start) {
        total += dual_stereo_rsv;
        dual_stereo_rsv = 0;
    }
    if (dual_stereo_rsv > 0) *dual_stereo = ce_bit_logp(ec, 1);
    else *dual_stereo = 0;
    /* the remaining bits */
    left = total - psum;
    percoeff = (int32_t)((uint32_t)left / (uint32_t)(eb[codedBands] - eb[start]));
    left -= (eb[codedBands] - eb[start]) * percoeff;
    for (j = start; j < codedBands; j++) bits[j] += (int)percoeff * (eb[j + 1] - eb[j]);
    for (j = start; j < codedBands; j++) {
        const int tmp = (int)(left < eb[j + 1] - eb[j] ? left : eb[j + 1] - eb[j]);
        bits[j] += tmp;
        left -= tmp;
    }
    balance = 0;
    for (j = start; j < codedBands; j++) {
        const int N0 = eb[j + 1] - eb[j], N = N0 << LM;
        int32_t excess;
        const int32_t bit = (int32_t)bits[j] + balance;
        if (N > 1) {
            excess = bit - cap[j] > 0 ? bit - cap[j] : 0;
            bits[j] = (int)(bit - excess);
            const int den = C * N + ((C == 2 && N > 2 && !*dual_stereo && j < *intensity) ? 1 : 0); /* the extra DoF in stereo */
            const int NClogN = den * (t->logn[j] + logM);
            int offset = (NClogN >> 1) - den * 21; /* FINE_OFFSET */
            if (N == 2) offset += den << ANM_CE_BITRES >> 2;
            if (bits[j] + offset < den * 2 << ANM_CE_BITRES) offset += NClogN >> 2;
            else if (bits[j] + offset < den * 3 << ANM_CE_BITRES) offset += NClogN >> 3;
            ebits[j] = ce_imax(0, (bits[j] + offset + (den << (ANM_CE_BITRES - 1))));
            ebits[j] = (int)((uint32_t)ebits[j] / (uint32_t)den) >> ANM_CE_BITRES;
            if (C * ebits[j] > (bits[j] >> ANM_CE_BITRES)) ebits[j] = bits[j] >> stereo >> ANM_CE_BITRES;
            ebits[j] = ce_imin(ebits[j], 8); /* MAX_FINE_BITS */
            fine_priority[j] = ebits[j] * (den << ANM_CE_BITRES) >= bits[j] + offset;
            bits[j] -= C * ebits[j] << ANM_CE_BITRES;
        } else {
            excess = bit - (C << ANM_CE_BITRES) > 0 ? bit - (C << ANM_CE_BITRES) : 0;
            bits[j] = (int)(bit - excess);
            ebits[j] = 0;
            fine_priority[j] = 1;
        }
        if (excess > 0) {
            const int extra_fine = ce_imin((int)(excess >> (stereo + ANM_CE_BITRES)), 8 - ebits[j]);
            ebits[j] += extra_fine;
            const int extra_bits = extra_fine * C << ANM_CE_BITRES;
            fine_priority[j] = extra_bits >= excess - balance;
            excess -= extra_bits;
        }
        balance = excess;
    }
    *balance_out = balance;
    for (; j < end; j++) { /* skipped bands: everything goes to fine energy */
        ebits[j] = bits[j] >> stereo >> ANM_CE_BITRES;
        bits[j] = 0;
        fine_priority[j] = ebits[j] < 1;
    }
    return codedBands;
}

ANM_CE_FN int ce_compute_allocation(const anm_celt_tables_t *t, int start, int end, const int *offsets, const int *cap, int alloc_trim, int *intensity,
                                    int *dual_stereo, int32_t total, int32_t *balance, int *pulses, int *ebits, int *fine_priority, int C, int LM,
                                    anm_ec_t *ec) {
    const uint8_t log2_frac_table[24] = {0, 8, 13, 16, 19, 21, 23, 24, 26, 27, 28, 29, 30, 31, 32, 32, 33, 34, 34, 35, 36, 36, 37, 37};
    const int16_t *eb = t->ebands;
    int bits1[ANM_CE_NB], bits2[ANM_CE_NB], thresh[ANM_CE_NB], trim_offset[ANM_CE_NB];
    int lo, hi, j, skip_start = start;
    total = total > 0 ? total : 0;
    const int skip_rsv = total >= 1 << ANM_CE_BITRES ? 1 << ANM_CE_BITRES : 0;
    total -= skip_rsv;
    int intensity_rsv = 0, dual_stereo_rsv = 0;
    if (C == 2) {
        intensity_rsv = log2_frac_table[end - start];
        if (intensity_rsv > total) {
            intensity_rsv = 0;
        } else {
            total -= intensity_rsv;
            dual_stereo_rsv = total >= 1 << ANM_CE_BITRES ? 1 << ANM_CE_BITRES : 0;
            total -= dual_stereo_rsv;
        }
    }
    for (j = start; j < end; j++) {
        thresh[j] = ce_imax(C << ANM_CE_BITRES, (3 * (eb[j + 1] - eb[j]) << LM << ANM_CE_BITRES) >> 4);
        trim_offset[j] = C * (eb[j + 1] - eb[j]) * (alloc_trim - 5 - LM) * (end - j - 1) * (1 << (LM + ANM_CE_BITRES)) >> 6;
        if ((eb[j + 1] - eb[j]) << LM == 1) trim_offset[j] -= C << ANM_CE_BITRES;
    }
    lo = 1;
    hi = ANM_CELT_ALLOC_VECTORS - 1;
    do {
        int done = 0, psum = 0;
        const int mid = (lo + hi) >> 1;
        for (j = end; j-- > start;) {
            const int N = eb[j + 1] - eb[j];
            int bitsj = C * N * t->alloc[mid * ANM_CE_NB + j] << LM >> 2;
            if (bitsj > 0) bitsj = ce_imax(0, bitsj + trim_offset[j]);
            bitsj += offsets[j];
            if (bitsj >= thresh[j] || done) {
                done = 1;
                psum += ce_imin(bitsj, cap[j]);
            } else if (bitsj >= C << ANM_CE_BITRES) {
                psum += C << ANM_CE_BITRES;
            }
        }
        if (psum > total) hi = mid - 1;
        else lo = mid + 1;
    } while (lo <= hi);
    hi = lo--;
    for (j = start; j < end; j++) {
        const int N = eb[j + 1] - eb[j];
        int bits1j = C * N * t->alloc[lo * ANM_CE_NB + j] << LM >> 2;
        int bits2j = hi >= ANM_CELT_ALLOC_VECTORS ? cap[j] : C * N * t->alloc[hi * ANM_CE_NB + j] << LM >> 2;
        if (bits1j > 0) bits1j = ce_imax(0, bits1j + trim_offset[j]);
        if (bits2j > 0) bits2j = ce_imax(0, bits2j + trim_offset[j]);
        if (lo > 0) bits1j += offsets[j];
        bits2j += offsets[j];
        if (offsets[j] > 0) skip_start = j;
        bits2j = ce_imax(0, bits2j - bits1j);
        bits1[j] = bits1j;
        bits2[j] = bits2j;
    }
    return ce_interp_bits2pulses(t, start, end, skip_start, bits1, bits2, thresh, cap, total, balance, skip_rsv, intensity, intensity_rsv, dual_stereo,
                                 dual_stereo_rsv, pulses, ebits, fine_priority, C, LM, ec);
}

/* the range decoder and the allocation's balance as they stand in front of the band loop: stage 1 records them, stage 2 resumes there */
typedef struct ce_resume {
    anm_ec_t dec; /* .bytes is not carried (the caller passes the arena again) */
    int32_t balance, band_total;
} ce_resume_t;

/* ---------------------------------------------------------------- the band loop (quant_all_bands) */
/* band_total: the frame's bits in 1/8 bit minus the anti-collapse reserve; balance: what clt_compute_allocation left; pulses / tf_res: per band.
 * sp == NULL: bits only (stage 1); otherwise the spectrum as well (see anm_celt_frame_symbols). */
ANM_CE_FN void ce_all_bands(const anm_celt_tables_t *t, anm_ec_t *dec, int C, int LM, int end, const int *pulses, const int *tf_res, int32_t balance,
                            int32_t band_total, int short_blocks, int spread, int intensity, int dual_stereo, int coded_bands, anm_celt_frame_t *out,
                            ce_spec_t *sp, int16_t *X_, uint8_t *collapse_masks) {
    ce_band_ctx_t ctx;
    ctx.t = t;
    ctx.ec = dec;
    ctx.intensity = intensity;
    ctx.out = out;
    ctx.sp = sp;
    const int M = 1 << LM, start = 0;
    const int16_t *eb = t->ebands;
    const int B = short_blocks ? M : 1;
    int i;
    const int NF = M * 120; /* coefficients per channel */
    int ds = dual_stereo;
    /* folding state: norm holds the bands decoded so far (per channel while dual stereo lasts), up to the last band's start */
    const int norm_offset = M * eb[start], norm_len = M * eb[ANM_CE_NB - 1] - norm_offset;
    int16_t *norm = sp ? sp->norm : 0, *norm2 = sp ? sp->norm + norm_len : 0;
    int16_t *lowband_scratch = sp ? (sp->band ? sp->band + 2 * 176 : X_ + M * eb[ANM_CE_NB - 1]) : 0;
    int lowband_offset = 0, update_lowband = 1;
    if (sp) sp->spread = spread;
    for (i = start; i < end; i++) {
        ctx.i = i;
        const int last = i == end - 1;
        const int N = M * eb[i + 1] - M * eb[i];
        int16_t *X = sp ? (sp->band ? sp->band : X_ + M * eb[i]) : 0, *Y = (sp && C == 2) ? (sp->band ? sp->band + 176 : X_ + NF + M * eb[i]) : 0;
        const int32_t tl = (int32_t)ce_tell_frac(dec);
        if (i != start) balance -= tl;
        const int32_t remaining = band_total - tl - 1;
        ctx.remaining_bits = remaining;
        int b;
        if (i <= coded_bands - 1) {
            const int32_t curr_balance = balance / ce_imin(3, coded_bands - i); /* celt_sudiv */
            b = ce_imax(0, ce_imin(16383, ce_imin((int)remaining + 1, pulses[i] + (int)curr_balance)));
        } else {
            b = 0;
        }
        if ((M * eb[i] - N >= M * eb[start] || i == start + 1) && (update_lowband || lowband_offset == 0)) lowband_offset = i;
        ctx.tf_change = tf_res[i];
        if (last) lowband_scratch = 0;
        /* a conservative estimate of the collapse masks of the bands this one folds from */
        int effective_lowband = -1;
        unsigned x_cm, y_cm;
        if (lowband_offset != 0 && (spread != 3 || B > 1 || tf_res[i] < 0)) { /* SPREAD_AGGRESSIVE */
            effective_lowband = ce_imax(0, M * eb[lowband_offset] - norm_offset - N); /* never repeat spectral content within one band */
            int fold_start = lowband_offset;
            while (M * eb[--fold_start] > effective_lowband + norm_offset) {}
            int fold_end = lowband_offset - 1;
            while (++fold_end < i && M * eb[fold_end] < effective_lowband + norm_offset + N) {}
            x_cm = y_cm = 0;
            if (sp) {
                int fold_i = fold_start;
                do {
                    x_cm |= collapse_masks[fold_i * C + 0];
                    y_cm |= collapse_masks[fold_i * C + C - 1];
                } while (++fold_i < fold_end);
            }
        } else {
            x_cm = y_cm = (1u << B) - 1; /* folding from the noise generator: every block gets energy */
        }
        if (ds && i == intensity) { /* dual stereo switches off to do intensity */
            ds = 0;
            if (sp) {
                CV_SYNC();
                for (int j = sp->lane; j < M * eb[i] - norm_offset; j += sp->nl) norm[j] = (int16_t)(((int32_t)norm[j] + norm2[j]) >> 1);
                CV_SYNC();
            }
        }
        int16_t *lb = (sp && effective_lowband != -1) ? norm + effective_lowband : 0;
        int16_t *lbo = (sp && !last) ? norm + M * eb[i] - norm_offset : 0;
        int16_t *lb2 = (sp && ds && effective_lowband != -1) ? norm2 + effective_lowband : 0;
        int16_t *lbo2 = (sp && ds && !last) ? norm2 + M * eb[i] - norm_offset : 0;
        ce_band_channels(&ctx, X, Y, C, ds, N, b, B, lb, lb2, LM, lbo, lbo2, lowband_scratch, &x_cm, &y_cm);
        if (sp && sp->band) { /* the finished band goes to the output */
            for (int c = 0; c < C; c++) {
                int16_t *dst = X_ + c * NF + M * eb[i];
                const int16_t *src = sp->band + 176 * c;
                for (int k = sp->lane; k < N; k += sp->nl) dst[k] = src[k];
            }
        }
        if (sp) {
            if (sp->lane == 0) {
                collapse_masks[i * C + 0] = (uint8_t)x_cm;
                collapse_masks[i * C + C - 1] = (uint8_t)y_cm;
            }
            CV_SYNC();
        }
        balance += pulses[i] + tl;
        update_lowband = b > (N << ANM_CE_BITRES); /* the folding position moves only while there is 1 bit / sample of depth */
    }
}

/* ---------------------------------------------------------------- one frame */
/* Everything the frame's bits say, WITHOUT the stream's history: the coarse energy symbols qi[c * 21 + band] and the sum of the fine and
 * final energy offsets eoff[c * 21 + band] (Q10) are returned instead of being applied -- no symbol of a frame depends on the band energies,
 * so frames decode independently of each other and only anm_celt_apply_energies() below is sequential per stream.
 * end = coded bands of the packet's bandwidth.  Returns 0, or a negative ANM_OPUS_* code (nothing is read then). */
/* sp != NULL (stage 2): the frame's normalised spectrum as well -- X_: [C][120 << LM] coefficients (celt_norm, Q14: what quant_all_bands leaves in
 * X, before anti-collapse), collapse_masks: [21 * C]; sp->seed / spread / disable_inv are inputs, sp->seed is updated.  The last band's part of
 * X_ doubles as scratch while the earlier bands are decoded, as in the reference. */
ANM_CE_FN int anm_celt_frame_symbols(const anm_celt_tables_t *t, const uint8_t *bytes, uint32_t mask, uint32_t base, uint32_t len, int C, int LM,
                                     int end, int16_t *qi_out, int16_t *eoff, anm_celt_frame_t *out, ce_spec_t *sp, int16_t *X_, uint8_t *collapse_masks,
                                     ce_resume_t *resume) {
    const uint8_t trim_icdf[11] = {126, 124, 119, 109, 87, 41, 19, 9, 4, 2, 0};
    const uint8_t spread_icdf[4] = {25, 23, 2, 0};
    const uint8_t tapset_icdf[3] = {2, 1, 0};
    const uint8_t small_energy_icdf[3] = {2, 1, 0};
    const int8_t tf_select_table[4][8] = {{0, -1, 0, -1, 0, -1, 0, -1}, {0, -1, 0, -2, 1, 0, 1, -1}, {0, -2, 0, -3, 2, 0, 1, -1}, {0, -2, 0, -3, 3, 0, 1, -1}};
    const int start = 0, M = 1 << LM;
    const int16_t *eb = t->ebands;
    anm_ec_t dec;
    int i, c;
    out->final_range = 0;
    out->flags = 0;
    out->pvq_codewords = out->pvq_pulses = out->pvq_index_xor = 0;
    if (len > 1275u || LM < 0 || LM > 3 || (C != 1 && C != 2) || end < 1 || end > ANM_CE_NB) return ANM_OPUS_BAD_ARG;
    if (len <= 1) { /* packet loss concealment in the reference: nothing is decoded, the final range is 0 */
        out->flags = ANM_CELT_F_LOST;
        return 0;
    }
    ce_init(&dec, bytes, mask, base, len);
    for (i = 0; i < 2 * ANM_CE_NB; i++) qi_out[i] = eoff[i] = 0;
    int32_t total_bits = (int32_t)len * 8;
    int32_t tell = ce_tell(&dec);
    int silence;
    if (tell >= total_bits) silence = 1;
    else if (tell == 1) silence = ce_bit_logp(&dec, 15);
    else silence = 0;
    if (silence) {
        tell = (int32_t)len * 8; /* pretend all the remaining bits were read */
        dec.nbits_total += tell - ce_tell(&dec);
    }
    int pf_pitch = 0, pf_qg = 0, pf_tapset = 0, pf = 0;
    if (start == 0 && tell + 16 <= total_bits) {
        if (ce_bit_logp(&dec, 1)) {
            pf = 1;
            const int octave = (int)ce_uint(&dec, 6);
            pf_pitch = (16 << octave) + (int)ce_bits(&dec, (uint32_t)(4 + octave)) - 1;
            pf_qg = (int)ce_bits(&dec, 3);
            if (ce_tell(&dec) + 2 <= total_bits) pf_tapset = ce_icdf(&dec, tapset_icdf, 2);
        }
        tell = ce_tell(&dec);
    }
    int transient = 0;
    if (LM > 0 && tell + 3 <= total_bits) {
        transient = ce_bit_logp(&dec, 3);
        tell = ce_tell(&dec);
    }
    const int short_blocks = transient ? M : 0;
    const int intra = tell + 3 <= total_bits ? ce_bit_logp(&dec, 3) : 0;

    /* ---- coarse energy (unquant_coarse_energy): the symbols only ---- */
    {
        const uint8_t *prob = t->e_prob + (LM * 2 + intra) * 42;
        const int32_t budget = (int32_t)len * 8;
        for (i = start; i < end; i++) {
            for (c = 0; c < C; ++c) {
                int qi;
                tell = ce_tell(&dec);
                if (budget - tell >= 15) {
                    const int pi = 2 * ce_imin(i, 20);
                    qi = ce_laplace(&dec, (uint32_t)prob[pi] << 7, prob[pi + 1] << 6);
                } else if (budget - tell >= 2) {
                    qi = ce_icdf(&dec, small_energy_icdf, 2);
                    qi = (qi >> 1) ^ -(qi & 1);
                } else if (budget - tell >= 1) {
                    qi = -ce_bit_logp(&dec, 1);
                } else {
                    qi = -1;
                }
                qi_out[i + c * ANM_CE_NB] = (int16_t)qi;
            }
        }
    }
    /* ---- tf_decode ---- */
    int tf_res[ANM_CE_NB];
    {
        uint32_t budget = len * 8u;
        uint32_t tl = (uint32_t)ce_tell(&dec);
        int logp = transient ? 2 : 4;
        const int tf_select_rsv = LM > 0 && tl + (uint32_t)logp + 1u <= budget;
        budget -= (uint32_t)tf_select_rsv;
        int tf_changed = 0, curr = 0;
        for (i = start; i < end; i++) {
            if (tl + (uint32_t)logp <= budget) {
                curr ^= ce_bit_logp(&dec, (uint32_t)logp);
                tl = (uint32_t)ce_tell(&dec);
                tf_changed |= curr;
            }
            tf_res[i] = curr;
            logp = transient ? 4 : 5;
        }
        int tf_select = 0;
        if (tf_select_rsv && tf_select_table[LM][4 * transient + 0 + tf_changed] != tf_select_table[LM][4 * transient + 2 + tf_changed])
            tf_select = ce_bit_logp(&dec, 1);
        for (i = start; i < end; i++) tf_res[i] = tf_select_table[LM][4 * transient + 2 * tf_select + tf_res[i]];
    }
    tell = ce_tell(&dec);
    int spread = 2; /* SPREAD_NORMAL */
    if (tell + 4 <= total_bits) spread = ce_icdf(&dec, spread_icdf, 5);
    /* ---- caps, dynamic allocation ---- */
    int cap[ANM_CE_NB], offsets[ANM_CE_NB];
    for (i = 0; i < ANM_CE_NB; i++) {
        const int N = (eb[i + 1] - eb[i]) << LM;
        cap[i] = (t->cache_caps[ANM_CE_NB * (2 * LM + C - 1) + i] + 64) * C * N >> 2;
    }
    int dynalloc_logp = 6;
    total_bits <<= ANM_CE_BITRES;
    tell = (int32_t)ce_tell_frac(&dec);
    for (i = start; i < end; i++) {
        const int width = C * (eb[i + 1] - eb[i]) << LM;
        const int quanta = ce_imin(width << ANM_CE_BITRES, ce_imax(6 << ANM_CE_BITRES, width));
        int loop_logp = dynalloc_logp, boost = 0;
        while (tell + (loop_logp << ANM_CE_BITRES) < total_bits && boost < cap[i]) {
            const int flag = ce_bit_logp(&dec, (uint32_t)loop_logp);
            tell = (int32_t)ce_tell_frac(&dec);
            if (!flag) break;
            boost += quanta;
            total_bits -= quanta;
            loop_logp = 1;
        }
        offsets[i] = boost;
        if (boost > 0) dynalloc_logp = ce_imax(2, dynalloc_logp - 1);
    }
    const int alloc_trim = tell + (6 << ANM_CE_BITRES) <= total_bits ? ce_icdf(&dec, trim_icdf, 7) : 5;
    int32_t bits = (((int32_t)len * 8) << ANM_CE_BITRES) - (int32_t)ce_tell_frac(&dec) - 1;
    const int anti_collapse_rsv = transient && LM >= 2 && bits >= ((LM + 2) << ANM_CE_BITRES) ? (1 << ANM_CE_BITRES) : 0;
    bits -= anti_collapse_rsv;
    int pulses[ANM_CE_NB], fine_quant[ANM_CE_NB], fine_priority[ANM_CE_NB];
    int intensity = 0, dual_stereo = 0;
    int32_t balance = 0;
    for (i = 0; i < ANM_CE_NB; i++) pulses[i] = fine_quant[i] = fine_priority[i] = 0;
    const int coded_bands = ce_compute_allocation(t, start, end, offsets, cap, alloc_trim, &intensity, &dual_stereo, bits, &balance, pulses, fine_quant,
                                                  fine_priority, C, LM, &dec);
    /* ---- fine energy ---- */
    for (i = start; i < end; i++) {
        if (fine_quant[i] <= 0) continue;
        for (c = 0; c < C; ++c) {
            const int q2 = (int)ce_bits(&dec, (uint32_t)fine_quant[i]);
            const int16_t offset = (int16_t)((((int32_t)q2 * 1024 + 512) >> fine_quant[i]) - 512);
            eoff[i + c * ANM_CE_NB] = (int16_t)(eoff[i + c * ANM_CE_NB] + offset); /* 16-bit wrapping adds commute with the coarse value */
        }
    }
    /* ---- the bands (quant_all_bands) ---- */
    const int32_t band_total = (int32_t)len * (8 << ANM_CE_BITRES) - anti_collapse_rsv;
    if (resume && !sp) { /* stage 1 leaves what stage 2 needs to start right here */
        resume->dec = dec;
        resume->balance = balance;
        resume->band_total = band_total;
    }
    ce_all_bands(t, &dec, C, LM, end, pulses, tf_res, balance, band_total, short_blocks, spread, intensity, dual_stereo, coded_bands, out, sp, X_, collapse_masks);
    int anti_collapse_on = 0;
    if (anti_collapse_rsv > 0) anti_collapse_on = (int)ce_bits(&dec, 1);
    /* ---- unquant_energy_finalise ---- */
    {
        int bits_left = (int)len * 8 - ce_tell(&dec);
        for (int prio = 0; prio < 2; prio++) {
            for (i = start; i < end && bits_left >= C; i++) {
                if (fine_quant[i] >= 8 || fine_priority[i] != prio) continue;
                for (c = 0; c < C; ++c) {
                    const int q2 = (int)ce_bits(&dec, 1);
                    const int16_t offset = (int16_t)((int16_t)(q2 * 1024 - 512) >> (fine_quant[i] + 1));
                    eoff[i + c * ANM_CE_NB] = (int16_t)(eoff[i + c * ANM_CE_NB] + offset);
                    bits_left--;
                }
            }
        }
    }
    out->final_range = dec.rng;
    out->tell_bits = ce_tell(&dec);
    out->flags = (uint32_t)(silence ? ANM_CELT_F_SILENCE : 0) | (uint32_t)(pf ? ANM_CELT_F_POSTFILTER : 0) | (uint32_t)(transient ? ANM_CELT_F_TRANSIENT : 0) |
                 (uint32_t)(intra ? ANM_CELT_F_INTRA : 0) | (uint32_t)(dual_stereo ? ANM_CELT_F_DUAL_STEREO : 0) | (uint32_t)(anti_collapse_on ? ANM_CELT_F_ANTI_COLLAPSE : 0) |
                 (uint32_t)(dec.error ? ANM_CELT_F_EC_ERROR : 0) | (uint32_t)(out->tell_bits > (int32_t)len * 8 ? ANM_CELT_F_OVERRUN : 0);
    out->pf_pitch = (uint16_t)pf_pitch;
    out->pf_gain_q = (uint8_t)pf_qg;
    out->pf_tapset = (uint8_t)pf_tapset;
    out->spread = (uint8_t)spread;
    out->alloc_trim = (uint8_t)alloc_trim;
    out->intensity = (uint8_t)intensity;
    out->coded_bands = (uint8_t)coded_bands;
    out->lm = (uint8_t)LM;
    out->channels = (uint8_t)C;
    out->pad[0] = (uint8_t)end;
    for (i = 0; i < ANM_CE_NB; i++) {
        out->tf_res[i] = (int8_t)(i < end ? tf_res[i] : 0);
        out->fine_quant[i] = (uint8_t)fine_quant[i];
        out->pulses[i] = (int16_t)pulses[i];
    }
    return 0;
}

ANM_CE_FN int anm_celt_entropy_symbols(const anm_celt_tables_t *t, const uint8_t *bytes, uint32_t mask, uint32_t base, uint32_t len, int C, int LM,
                                       int end, int16_t *qi_out, int16_t *eoff, anm_celt_frame_t *out) {
    return anm_celt_frame_symbols(t, bytes, mask, base, len, C, LM, end, qi_out, eoff, out, 0, 0, 0, 0);
}

/* The sequential part of a stream: the band energies after a frame from the energies before it (old_e, Q10, [2][21]), the frame's coarse
 * symbols and offsets -- unquant_coarse_energy's prediction (celt/quant_bands.c:427-490, fixed-point build) followed by the fine and final
 * offsets, the mono / silence / band-limit rules of celt_decode_with_ec (celt/celt_decoder.c:941-945, 1100-1104, 1137-1166).  `fr` is the
 * record anm_celt_entropy_symbols() filled; its band_e is written here. */
ANM_CE_FN void anm_celt_apply_energies(anm_celt_frame_t *fr, const int16_t *qi, const int16_t *eoff, int16_t *old_e) {
    const int16_t pred_coef[4] = {29440, 26112, 21248, 16384}, beta_coef[4] = {30147, 22282, 12124, 6554};
    int i, c;
    if (!(fr->flags & ANM_CELT_F_LOST)) {
        const int C = fr->channels, LM = fr->lm, end = fr->pad[0], intra = (fr->flags & ANM_CELT_F_INTRA) != 0;
        const int coef = intra ? 0 : pred_coef[LM], beta = intra ? 4915 : beta_coef[LM];
        int32_t prev[2] = {0, 0};
        if (C == 1)
            for (i = 0; i < ANM_CE_NB; i++) old_e[i] = old_e[i] > old_e[ANM_CE_NB + i] ? old_e[i] : old_e[ANM_CE_NB + i];
        for (i = 0; i < end; i++) {
            for (c = 0; c < C; ++c) {
                const int32_t q = (int32_t)qi[i + c * ANM_CE_NB] * 1024; /* SHL32(qi, DB_SHIFT) */
                int16_t *e = &old_e[i + c * ANM_CE_NB];
                if (*e < -9216) *e = -9216; /* MAX16(-QCONST16(9, DB_SHIFT), .) */
                int32_t tmp = ce_pshr32((int32_t)coef * *e, 8) + prev[c] + q * 128;
                if (tmp < -3670016) tmp = -3670016; /* -QCONST32(28, DB_SHIFT + 7) */
                *e = (int16_t)(ce_pshr32(tmp, 7) + eoff[i + c * ANM_CE_NB]);
                prev[c] = prev[c] + q * 128 - (int32_t)beta * (int16_t)ce_pshr32(q, 8);
            }
        }
        if (fr->flags & ANM_CELT_F_SILENCE)
            for (i = 0; i < C * ANM_CE_NB; i++) old_e[i] = -28672; /* -QCONST16(28, DB_SHIFT) */
        if (C == 1)
            for (i = 0; i < ANM_CE_NB; i++) old_e[ANM_CE_NB + i] = old_e[i];
        for (c = 0; c < 2; ++c)
            for (i = end; i < ANM_CE_NB; i++) old_e[c * ANM_CE_NB + i] = 0;
    }
    for (i = 0; i < 2 * ANM_CE_NB; i++) fr->band_e[i] = old_e[i];
}

/* what stage 2 needs of the stream's history for one frame, as the stream stood BEFORE the frame */
typedef struct ce_hist {
    int16_t log_e1[2 * ANM_CE_NB], log_e2[2 * ANM_CE_NB];
    uint32_t seed;
} ce_hist_t;

/* One frame's step of the per-stream state: hist (may be NULL) receives the histories before the frame, then the band energies are applied
 * and the histories updated as celt_decode_with_ec does after synthesis (celt/celt_decoder.c:1134-1166).  Lost frames change nothing here
 * (the reference's concealment, which does, is not built). */
ANM_CE_FN void anm_celt_stream_step(anm_celt_frame_t *fr, const int16_t *qi, const int16_t *eoff, anm_celt_stream_t *st, ce_hist_t *hist) {
    int i;
    if (!(st->flags & 1u)) {
        for (i = 0; i < 2 * ANM_CE_NB; i++) st->log_e1[i] = st->log_e2[i] = -28672; /* -QCONST16(28, DB_SHIFT) */
        st->flags |= 1u;
    }
    if (hist) {
        for (i = 0; i < 2 * ANM_CE_NB; i++) {
            hist->log_e1[i] = st->log_e1[i];
            hist->log_e2[i] = st->log_e2[i];
        }
        hist->seed = st->rng;
    }
    anm_celt_apply_energies(fr, qi, eoff, st->old_e);
    if (fr->flags & ANM_CELT_F_LOST) return;
    const int end = fr->pad[0];
    if (!(fr->flags & ANM_CELT_F_TRANSIENT)) {
        for (i = 0; i < 2 * ANM_CE_NB; i++) {
            st->log_e2[i] = st->log_e1[i];
            st->log_e1[i] = st->old_e[i];
        }
    } else {
        for (i = 0; i < 2 * ANM_CE_NB; i++) st->log_e1[i] = st->log_e1[i] < st->old_e[i] ? st->log_e1[i] : st->old_e[i];
    }
    for (int c = 0; c < 2; ++c)
        for (i = end; i < ANM_CE_NB; i++) st->log_e1[c * ANM_CE_NB + i] = st->log_e2[c * ANM_CE_NB + i] = -28672;
    st->rng = fr->final_range;
}

/* Stage 2 for one frame whose stage-1 record `rec` (band energies applied) and history `hist` are known: the frame is decoded again, this time
 * with its spectrum, and anti-collapse is applied when the frame asks for it.  X: [C][120 << LM]; cm: 42 collapse masks; sp: working storage.
 * Returns the noise seed after the bands in sp->seed. */
/* resume == NULL: the frame is decoded from its first bit once more (the stand-alone form; tests hold it against the resumed one) */
ANM_CE_FN int anm_celt_frame_spectrum(const anm_celt_tables_t *t, const uint8_t *bytes, uint32_t mask, uint32_t base, uint32_t len, int C, int LM, int end,
                                      int disable_inv, const ce_hist_t *hist, const anm_celt_frame_t *rec, const ce_resume_t *resume, ce_spec_t *sp,
                                      int16_t *X, uint8_t *cm) {
    sp->seed = hist->seed;
    sp->disable_inv = disable_inv;
    for (int i = sp->lane; i < 2 * ANM_CE_NB; i += sp->nl) cm[i] = 0;
    CV_SYNC();
    uint32_t flags;
#ifndef __CUDA_ARCH__
    if (!resume) {
        int16_t qi[2 * ANM_CE_NB], eoff[2 * ANM_CE_NB];
        anm_celt_frame_t again;
        const int rc = anm_celt_frame_symbols(t, bytes, mask, base, len, C, LM, end, qi, eoff, &again, sp, X, cm, 0);
        if (rc != 0) return rc;
        flags = again.flags;
        if (flags & ANM_CELT_F_LOST) return 0;
    } else
#endif
    {
        /* everything in front of the band loop was decoded by stage 1: its symbols are in the record, the decoder's state in *resume */
        flags = rec->flags;
        if (flags & ANM_CELT_F_LOST) return 0;
        anm_ec_t dec = resume->dec;
        dec.bytes = bytes;
        dec.mask = mask;
        dec.base = base;
        int pulses[ANM_CE_NB], tf_res[ANM_CE_NB];
        for (int i = 0; i < ANM_CE_NB; i++) {
            pulses[i] = rec->pulses[i];
            tf_res[i] = rec->tf_res[i];
        }
        ce_all_bands(t, &dec, C, LM, end, pulses, tf_res, resume->balance, resume->band_total, (flags & ANM_CELT_F_TRANSIENT) != 0, rec->spread, rec->intensity,
                     (flags & ANM_CELT_F_DUAL_STEREO) != 0, rec->coded_bands, 0, sp, X, cm);
    }
    if (flags & ANM_CELT_F_ANTI_COLLAPSE)
        cv_anti_collapse(t, X, cm, LM, C, 120 << LM, end, rec->band_e, hist->log_e1, hist->log_e2, rec->pulses, sp->seed, sp->lane, sp->nl);
    return 0;
}

/* both parts for one frame (the host-side test harness; the kernels run them as two passes) */
ANM_CE_FN int anm_celt_entropy_frame(const anm_celt_tables_t *t, const uint8_t *bytes, uint32_t mask, uint32_t base, uint32_t len, int C, int LM,
                                     int end, int16_t *old_e, anm_celt_frame_t *out) {
    int16_t qi[2 * ANM_CE_NB], eoff[2 * ANM_CE_NB];
    const int rc = anm_celt_entropy_symbols(t, bytes, mask, base, len, C, LM, end, qi, eoff, out);
    if (rc != 0) return rc;
    anm_celt_apply_energies(out, qi, eoff, old_e);
    return 0;
}

#endif /* ANM_CELT_ENTROPY_H_INCLUDED */
