/*
 * anm_celt_vec.h -- stage 2 of the batched CELT frame decoder, the arithmetic part: everything
 * quant_all_bands() does to the normalised spectrum of a frame once the symbols are read -- PVQ codeword -> pulse vector,
 * normalisation, spreading rotation, the Haar / Hadamard reorderings, folding and noise filling, stereo merge, collapse masks.
 * Host + device code, included by anm_celt_entropy.h (whose band functions call it when a spectrum is asked for).
 *
 * TRANSCRIPTION NOTICE.  The reference builds libopus in its FIXED_POINT configuration (hardware/lib/libopus/src/config.h), so the
 * spectrum is integer arithmetic with a prescribed rounding at every step, and the test for this file is equality, coefficient by
 * coefficient, with what the reference's quant_all_bands leaves in X (oracle/ref_celt_shim.c: ref_celt_spectrum_trace).  The
 * functions therefore restate, operation by operation, (c) Xiph.Org / Skype / Octasic / Jean-Marc Valin / Timothy B. Terriberry /
 * CSIRO / Gregory Maxwell code (BSD 3-clause, hardware/lib/libopus/COPYING):
 *   fixed-point operators                    celt/fixed_generic.h:37-160 (MULT16_16, *_Q15, *_P15, PSHR32, VSHR32, ...)
 *   celt_rsqrt_norm, celt_sqrt, celt_cos_norm, celt_rcp, celt_div, celt_ilog2     celt/mathops.c:97-207, celt/mathops.h:180-249
 *   cwrsi (codeword index -> pulses)         celt/cwrs.c:464-538
 *   exp_rotation, normalise_residual, extract_collapse_mask, renormalise_vector   celt/vq.c:47-163, 383-407
 *   haar1, (de)interleave_hadamard, stereo_merge                                  celt/bands.c:426-477, 576-645
 *   anti_collapse                            celt/bands.c:268-361; celt_exp2                celt/mathops.h:221-246
 */
#ifndef ANM_CELT_VEC_H_INCLUDED
#define ANM_CELT_VEC_H_INCLUDED

/* ---------------------------------------------------------------- lanes
 * The loops over a band's coefficients are written for `nl` cooperating lanes (lane, lane + nl, ...) with synchronisation points between the phases
 * (CV_SYNC, and CV_SUM / CV_OR for the reductions); in-place recurrences (the spreading rotation) and stores of sequentially generated values
 * (pulses, noise) are done by lane 0.  A warp-per-frame kernel built on this (every lane running the symbol decode redundantly, the frame's 7.4 KB
 * working set in shared memory) was measured 3 x SLOWER than one thread per frame on B200 -- the sequential symbol decode dominates a frame, and a
 * warp per frame leaves 16 frames in flight per SM instead of 900 -- so the kernels and the host harness run with ONE lane and the macros are empty;
 * ANM_CELT_WARP_LANES brings the warp form back for experiments. */
#if defined(__CUDA_ARCH__) && defined(ANM_CELT_WARP_LANES)
#define CV_SYNC() __syncwarp()
#define CV_SUM(x) ((int32_t)__reduce_add_sync(0xffffffffu, (unsigned)(x)))
#define CV_OR(x) __reduce_or_sync(0xffffffffu, (unsigned)(x))
#else
#define CV_SYNC() ((void)0)
#define CV_SUM(x) (x)
#define CV_OR(x) (x)
#endif

/* ---------------------------------------------------------------- fixed-point operators */
#define CV_M16(a, b) ((int32_t)(int16_t)(a) * (int32_t)(int16_t)(b))          /* MULT16_16 */
#define CV_Q15(a, b) (CV_M16(a, b) >> 15)                                     /* MULT16_16_Q15 */
#define CV_P15(a, b) ((16384 + CV_M16(a, b)) >> 15)                           /* MULT16_16_P15 */
#define CV_A16(a, b) ((int16_t)((int16_t)(a) + (int16_t)(b)))                 /* ADD16 */
#define CV_S16(a, b) ((int32_t)(int16_t)(a) - (int32_t)(int16_t)(b))          /* SUB16: not narrowed */
#define CV_PSHR32(a, s) (((int32_t)(a) + ((1 << (s)) >> 1)) >> (s))           /* PSHR32 */

ANM_CE_FN int32_t cv_vshr32(int32_t a, int s) { return s > 0 ? a >> s : (int32_t)((uint32_t)a << -s); }
ANM_CE_FN int cv_ilog2(int32_t x) { return anm_ce_ilog((uint32_t)x) - 1; }
ANM_CE_FN int32_t cv_mult32_32_q31(int32_t a, int32_t b) { return (int32_t)(((int64_t)a * (int64_t)b) >> 31); }

/* reciprocal square root of a Q16 value in [0.25, 1), Q14 */
ANM_CE_FN int16_t cv_rsqrt_norm(int32_t x) {
    const int16_t n = (int16_t)(x - 32768);
    const int16_t r = CV_A16(23557, CV_Q15(n, CV_A16(-13490, CV_Q15(n, 6713))));
    const int16_t r2 = (int16_t)CV_Q15(r, r);
    const int16_t y = (int16_t)((uint16_t)(int16_t)CV_S16(CV_A16(CV_Q15(r2, n), r2), 16384) << 1);
    return CV_A16(r, CV_Q15(r, CV_Q15(y, CV_S16(CV_Q15(y, 12288), 16384))));
}
/* square root, QX in, QX/2 out */
ANM_CE_FN int32_t cv_sqrt(int32_t x) {
    if (x == 0) return 0;
    if (x >= 1073741824) return 32767;
    const int k = (cv_ilog2(x) >> 1) - 7;
    x = cv_vshr32(x, 2 * k);
    const int16_t n = (int16_t)(x - 32768);
    int32_t rt = CV_A16(23175, CV_Q15(n, CV_A16(11561, CV_Q15(n, CV_A16(-3011, CV_Q15(n, CV_A16(1699, CV_Q15(n, -664))))))));
    rt = cv_vshr32(rt, 7 - k);
    return rt;
}
ANM_CE_FN int16_t cv_cos_pi_2(int16_t x) {
    const int16_t x2 = (int16_t)CV_P15(x, x);
    const int32_t v = CV_S16(32767, x2) + CV_P15(x2, -7651 + CV_P15(x2, 8277 + CV_P15(-626, x2)));
    return CV_A16(1, v < 32766 ? v : 32766);
}
/* cos(pi/2 x), x in Q16 turns of a quarter circle (celt_cos_norm) */
ANM_CE_FN int16_t cv_cos_norm(int32_t x) {
    x = x & 0x0001ffff;
    if (x > (1 << 16)) x = (1 << 17) - x;
    if (x & 0x00007fff) {
        if (x < (1 << 15)) return cv_cos_pi_2((int16_t)x);
        return (int16_t)-cv_cos_pi_2((int16_t)(65536 - x));
    }
    if (x & 0x0000ffff) return 0;
    if (x & 0x0001ffff) return -32767;
    return 32767;
}
/* reciprocal, Q15 in, Q16 out */
ANM_CE_FN int32_t cv_rcp(int32_t x) {
    const int i = cv_ilog2(x);
    const int16_t n = (int16_t)(cv_vshr32(x, i - 15) - 32768);
    int16_t r = CV_A16(30840, CV_Q15(-15420, n));
    r = (int16_t)CV_S16(r, CV_Q15(r, CV_A16(CV_Q15(r, n), CV_A16(r, -32768))));
    r = (int16_t)CV_S16(r, CV_A16(1, CV_Q15(r, CV_A16(CV_Q15(r, n), CV_A16(r, -32768)))));
    return cv_vshr32((int32_t)r, i - 16);
}
ANM_CE_FN uint32_t cv_lcg(uint32_t seed) { return 1664525u * seed + 1013904223u; }

/* ---------------------------------------------------------------- PVQ: codeword index -> pulse vector, returns sum of squares */
ANM_CE_FN int32_t cv_cwrsi(const anm_celt_tables_t *t, int n, int k, uint32_t i, int *y, int lane) {
    uint32_t p;
    int s, k0;
    int16_t val;
    int32_t yy = 0;
    while (n > 2) {
        uint32_t q;
        if (k >= n) { /* many pulses */
            p = ce_pvq_u(t, n, k + 1);
            s = -(int)(i >= p);
            i -= p & (uint32_t)s;
            k0 = k;
            q = ce_pvq_u(t, n, n);
            if (q > i) {
                k = n;
                do p = ce_pvq_u(t, --k, n);
                while (p > i);
            } else {
                for (p = ce_pvq_u(t, n, k); p > i; p = ce_pvq_u(t, n, k)) k--;
            }
            i -= p;
            val = (int16_t)((k0 - k + s) ^ s);
            if (lane == 0) *y = val;
            y++;
            yy += CV_M16(val, val);
        } else { /* many dimensions */
            p = ce_pvq_u(t, k, n);
            q = ce_pvq_u(t, k + 1, n);
            if (p <= i && i < q) {
                i -= p;
                if (lane == 0) *y = 0;
                y++;
            } else {
                s = -(int)(i >= q);
                i -= q & (uint32_t)s;
                k0 = k;
                do p = ce_pvq_u(t, --k, n);
                while (p > i);
                i -= p;
                val = (int16_t)((k0 - k + s) ^ s);
                if (lane == 0) *y = val;
                y++;
                yy += CV_M16(val, val);
            }
        }
        n--;
    }
    /* n == 2 */
    p = 2u * (uint32_t)k + 1u;
    s = -(int)(i >= p);
    i -= p & (uint32_t)s;
    k0 = k;
    k = (int)((i + 1) >> 1);
    if (k) i -= 2u * (uint32_t)k - 1u;
    val = (int16_t)((k0 - k + s) ^ s);
    if (lane == 0) *y = val;
    y++;
    yy += CV_M16(val, val);
    /* n == 1 */
    s = -(int)i;
    val = (int16_t)((k + s) ^ s);
    if (lane == 0) *y = val;
    yy += CV_M16(val, val);
    CV_SYNC();
    return yy;
}

/* ---------------------------------------------------------------- vector operations on celt_norm (int16, Q14) */
ANM_CE_FN void cv_exp_rotation1(int16_t *X, int len, int stride, int16_t c, int16_t s) {
    const int16_t ms = (int16_t)-s;
    int16_t *p = X;
    int i;
    for (i = 0; i < len - stride; i++) {
        const int16_t x1 = p[0], x2 = p[stride];
        p[stride] = (int16_t)CV_PSHR32(CV_M16(c, x2) + CV_M16(s, x1), 15);
        *p++ = (int16_t)CV_PSHR32(CV_M16(c, x1) + CV_M16(ms, x2), 15);
    }
    p = &X[len - 2 * stride - 1];
    for (i = len - 2 * stride - 1; i >= 0; i--) {
        const int16_t x1 = p[0], x2 = p[stride];
        p[stride] = (int16_t)CV_PSHR32(CV_M16(c, x2) + CV_M16(s, x1), 15);
        *p-- = (int16_t)CV_PSHR32(CV_M16(c, x1) + CV_M16(ms, x2), 15);
    }
}
/* the decoder's direction (dir = -1) of exp_rotation: in-place recurrences, one lane per interleaved block */
ANM_CE_FN void cv_exp_rotation_dec(int16_t *X, int len, int stride, int K, int spread, int lane, int nl) {
    if (2 * K >= len || spread == 0) return;
    const int factor = spread == 1 ? 15 : spread == 2 ? 10 : 5;
    const int16_t gain = (int16_t)cv_mult32_32_q31(CV_M16(32767, len), cv_rcp(len + factor * K));
    const int16_t theta = (int16_t)(CV_Q15(gain, gain) >> 1);
    const int16_t c = cv_cos_norm((int32_t)theta), s = cv_cos_norm((int32_t)CV_S16(32767, theta));
    int stride2 = 0;
    if (len >= 8 * stride) {
        stride2 = 1;
        while ((stride2 * stride2 + stride2) * stride + (stride >> 2) < len) stride2++;
    }
    len = (int)((uint32_t)len / (uint32_t)stride);
    CV_SYNC();
    for (int i = lane; i < stride; i += nl) {
        if (stride2) cv_exp_rotation1(X + i * len, len, stride2, s, c);
        cv_exp_rotation1(X + i * len, len, 1, c, s);
    }
    CV_SYNC();
}
/* pulses -> unit-norm vector scaled by gain (normalise_residual) */
ANM_CE_FN void cv_normalise_residual(const int *iy, int16_t *X, int N, int32_t Ryy, int16_t gain, int lane, int nl) {
    const int k = cv_ilog2(Ryy) >> 1;
    const int32_t t = cv_vshr32(Ryy, 2 * (k - 7));
    const int16_t g = (int16_t)CV_P15(cv_rsqrt_norm(t), gain);
    CV_SYNC();
    for (int i = lane; i < N; i += nl) X[i] = (int16_t)CV_PSHR32(CV_M16(g, iy[i]), k + 1);
    CV_SYNC();
}
ANM_CE_FN unsigned cv_collapse_mask(const int *iy, int N, int B, int lane, int nl) {
    if (B <= 1) return 1;
    const int N0 = (int)((uint32_t)N / (uint32_t)B);
    unsigned mask = 0;
    CV_SYNC();
    /* b = i / N0, kept up without dividing (here and below: an integer division is some thirty instructions, per element more than the work itself) */
    int b = (int)((uint32_t)lane / (uint32_t)N0), r = lane - b * N0;
    for (int i = lane; i < N; i += nl) {
        mask |= (unsigned)(iy[i] != 0) << b;
        for (r += nl; r >= N0; r -= N0) b++;
    }
    return CV_OR(mask);
}
ANM_CE_FN void cv_renormalise(int16_t *X, int N, int16_t gain, int lane, int nl) {
    int32_t E = 0;
    CV_SYNC();
    for (int i = lane; i < N; i += nl) E += CV_M16(X[i], X[i]);
    E = 1 + CV_SUM(E); /* EPSILON */
    const int k = cv_ilog2(E) >> 1;
    const int32_t t = cv_vshr32(E, 2 * (k - 7));
    const int16_t g = (int16_t)CV_P15(cv_rsqrt_norm(t), gain);
    for (int i = lane; i < N; i += nl) X[i] = (int16_t)CV_PSHR32(CV_M16(g, X[i]), k + 1);
    CV_SYNC();
}
ANM_CE_FN void cv_haar1(int16_t *X, int N0, int stride, int lane, int nl) {
    N0 >>= 1;
    CV_SYNC();
    /* element idx = j * stride + i of the upper halves: its pair sits at p = 2 * j * stride + i and p + stride */
    int j = (int)((uint32_t)lane / (uint32_t)stride), i = lane - j * stride;
    for (int idx = lane; idx < stride * N0; idx += nl) {
        int16_t *p = X + 2 * j * stride + i;
        const int32_t t1 = CV_M16(23170, p[0]), t2 = CV_M16(23170, p[stride]);
        p[0] = (int16_t)CV_PSHR32(t1 + t2, 15);
        p[stride] = (int16_t)CV_PSHR32(t1 - t2, 15);
        for (i += nl; i >= stride; i -= stride) j++;
    }
    CV_SYNC();
}
/* ordery_table (bands.c:580-585) for stride 2, 4, 8, 16, one entry per nibble, entry 0 lowest: {1,0} {3,0,2,1} {7,0,4,3,6,1,5,2}
 * {15,0,8,7,12,3,11,4,14,1,9,6,13,2,10,5} (an array here is built on the stack on every call) */
ANM_CE_FN uint64_t cv_ordery_row(int stride) {
    return stride == 2 ? 0x01ull : stride == 4 ? 0x1203ull : stride == 8 ? 0x25163407ull : 0x5A2D691E4B3C780Full;
}
ANM_CE_FN int cv_ordery(uint64_t row, int i) { return (int)(row >> (4 * i)) & 15; }
/* frequency order -> time order (tmp: N0 * stride entries of scratch) */
ANM_CE_FN void cv_deinterleave_hadamard(int16_t *X, int16_t *tmp, int N0, int stride, int hadamard, int lane, int nl) {
    const int N = N0 * stride;
    CV_SYNC();
    int i = (int)((uint32_t)lane / (uint32_t)N0), j = lane - i * N0;
    const uint64_t row = cv_ordery_row(stride);
    for (int idx = lane; idx < N; idx += nl) {
        const int o = hadamard ? cv_ordery(row, i) : i;
        tmp[o * N0 + j] = X[j * stride + i];
        for (j += nl; j >= N0; j -= N0) i++;
    }
    CV_SYNC();
    for (int i = lane; i < N; i += nl) X[i] = tmp[i];
    CV_SYNC();
}
ANM_CE_FN void cv_interleave_hadamard(int16_t *X, int16_t *tmp, int N0, int stride, int hadamard, int lane, int nl) {
    const int N = N0 * stride;
    CV_SYNC();
    int i = (int)((uint32_t)lane / (uint32_t)N0), j = lane - i * N0;
    const uint64_t row = cv_ordery_row(stride);
    for (int idx = lane; idx < N; idx += nl) {
        const int o = hadamard ? cv_ordery(row, i) : i;
        tmp[j * stride + i] = X[o * N0 + j];
        for (j += nl; j >= N0; j -= N0) i++;
    }
    CV_SYNC();
    for (int i = lane; i < N; i += nl) X[i] = tmp[i];
    CV_SYNC();
}
/* mid / side -> left / right of a band */
ANM_CE_FN void cv_stereo_merge(int16_t *X, int16_t *Y, int16_t mid, int N, int lane, int nl) {
    int32_t xp = 0, side = 0;
    CV_SYNC();
    for (int j = lane; j < N; j += nl) {
        xp += CV_M16(Y[j], X[j]);
        side += CV_M16(Y[j], Y[j]);
    }
    xp = CV_SUM(xp);
    side = CV_SUM(side);
    xp = (int32_t)(((int64_t)mid * xp) >> 15); /* MULT16_32_Q15 */
    const int16_t mid2 = (int16_t)(mid >> 1);
    const int32_t El = CV_M16(mid2, mid2) + side - 2 * xp, Er = CV_M16(mid2, mid2) + side + 2 * xp;
    if (Er < 161061 || El < 161061) { /* QCONST32(6e-4f, 28) */
        for (int j = lane; j < N; j += nl) Y[j] = X[j];
        CV_SYNC();
        return;
    }
    int kl = cv_ilog2(El) >> 1, kr = cv_ilog2(Er) >> 1;
    const int16_t lgain = cv_rsqrt_norm(cv_vshr32(El, (kl - 7) << 1)), rgain = cv_rsqrt_norm(cv_vshr32(Er, (kr - 7) << 1));
    if (kl < 7) kl = 7;
    if (kr < 7) kr = 7;
    for (int j = lane; j < N; j += nl) {
        const int16_t l = (int16_t)CV_P15(mid, X[j]), r = Y[j];
        X[j] = (int16_t)CV_PSHR32(CV_M16(lgain, CV_S16(l, r)), kl + 1);
        Y[j] = (int16_t)CV_PSHR32(CV_M16(rgain, CV_A16(l, r)), kr + 1);
    }
    CV_SYNC();
}

/* 2^x, Q10 in, Q16 out (celt_exp2) */
ANM_CE_FN int32_t cv_exp2(int16_t x) {
    const int integer = x >> 10;
    if (integer > 14) return 0x7f000000;
    if (integer < -15) return 0;
    const int16_t frac = (int16_t)((uint16_t)(int16_t)(x - (int16_t)((uint16_t)(int16_t)integer << 10)) << 4);
    const int16_t v = CV_A16(16383, CV_Q15(frac, CV_A16(22804, CV_Q15(frac, CV_A16(14819, CV_Q15(10204, frac))))));
    return cv_vshr32((int32_t)v, -integer - 2);
}

/* Transient frames: a short block of a band that received no pulse at all is filled with noise at the level the two previous frames
 * suggest, then the band is renormalised.  X_: [C][size] coefficients; log_e: this frame's band energies [2][21] (Q10), prev1 / prev2: the
 * decoder's two log-energy histories before the frame; pulses: the bands' PVQ budgets; seed: the noise generator after the bands. */
ANM_CE_FN void cv_anti_collapse(const anm_celt_tables_t *t, int16_t *X_, const uint8_t *collapse_masks, int LM, int C, int size, int end,
                                const int16_t *log_e, const int16_t *prev1, const int16_t *prev2, const int16_t *pulses, uint32_t seed, int lane, int nl) {
    const int NB = 21;
    for (int i = 0; i < end; i++) {
        const int N0 = t->ebands[i + 1] - t->ebands[i];
        const int depth = (int)((uint32_t)(1 + pulses[i]) / (uint32_t)N0) >> LM; /* in 1/8 bit per coefficient */
        const int32_t thresh32 = cv_exp2((int16_t)-(int16_t)((uint16_t)(int16_t)depth << 7)) >> 1; /* -SHL16(depth, 10 - BITRES) */
        const int16_t thresh = (int16_t)(((int64_t)16384 * (thresh32 < 32767 ? thresh32 : 32767)) >> 15); /* MULT16_32_Q15(0.5, .) */
        int32_t tt = N0 << LM;
        const int shift = cv_ilog2(tt) >> 1;
        tt = (int32_t)((uint32_t)tt << ((7 - shift) << 1));
        const int16_t sqrt_1 = cv_rsqrt_norm(tt);
        for (int c = 0; c < C; c++) {
            int16_t p1 = prev1[c * NB + i], p2 = prev2[c * NB + i];
            if (C == 1) {
                p1 = p1 > prev1[NB + i] ? p1 : prev1[NB + i];
                p2 = p2 > prev2[NB + i] ? p2 : prev2[NB + i];
            }
            int32_t Ediff = (int32_t)log_e[c * NB + i] - (int32_t)(p1 < p2 ? p1 : p2);
            if (Ediff < 0) Ediff = 0;
            int16_t r;
            if (Ediff < 16384) {
                const int32_t r32 = cv_exp2((int16_t)-(int16_t)Ediff) >> 1;
                r = (int16_t)(2 * (r32 < 16383 ? r32 : 16383));
            } else {
                r = 0;
            }
            if (LM == 3) r = (int16_t)(CV_M16(23170, r < 23169 ? r : 23169) >> 14); /* MULT16_16_Q14(23170, MIN32(23169, r)) */
            r = (int16_t)((thresh < r ? thresh : r) >> 1);
            r = (int16_t)(CV_Q15(sqrt_1, r) >> shift);
            int16_t *X = X_ + c * size + (t->ebands[i] << LM);
            int renormalize = 0;
            for (int k = 0; k < 1 << LM; k++) {
                if (!(collapse_masks[i * C + c] & 1 << k)) { /* this short block collapsed */
                    for (int j = 0; j < N0; j++) { /* every lane steps the generator, one stores */
                        seed = cv_lcg(seed);
                        if (lane == 0) X[(j << LM) + k] = (seed & 0x8000u) ? r : (int16_t)-r;
                    }
                    renormalize = 1;
                }
            }
            if (renormalize) cv_renormalise(X, N0 << LM, 32767, lane, nl);
        }
    }
}

#endif /* ANM_CELT_VEC_H_INCLUDED */
