/*
 * anm_celt_synth.h -- stage 3 of the batched CELT frame decoder: from the normalised spectrum to PCM.
 * denormalisation by the band energies, the inverse MDCT (one long block or 1 << LM short blocks) on a fixed-point FFT, the
 * windowed overlap-add, the pitch post-filter and the de-emphasis -- celt_synthesis(), comb_filter() and deemphasis() as
 * celt_decode_with_ec() runs them (celt/celt_decoder.c:1104-1170).  Host + device code.
 *
 * The work splits in two: everything up to the raw inverse-MDCT output of a block (cs_frame_blocks) depends on nothing but the
 * frame itself and runs frame-parallel; the overlap-add, the post-filter (an IIR over the decoder's output history) and the
 * de-emphasis (a one-pole IIR) are a short per-stream recurrence (cs_stream_frame).
 *
 * TRANSCRIPTION NOTICE.  Fixed-point build of the reference: the test for this file is that the int16 PCM equals the reference
 * decoder's sample for sample.  The functions restate, operation by operation, (c) Xiph.Org / Skype / Octasic / Jean-Marc Valin /
 * Timothy B. Terriberry / CSIRO / Gregory Maxwell / Mark Borgerding code (BSD 3-clause, hardware/lib/libopus/COPYING):
 *   denormalise_bands                        celt/bands.c:196-265; celt_exp2_frac celt/mathops.h:227-232
 *   clt_mdct_backward                        celt/mdct.c:241-342
 *   opus_fft_impl, kf_bfly2 / 3 / 4 / 5      celt/kiss_fft.c:47-326, 520-566; celt/_kiss_fft_guts.h:58-104
 *   celt_synthesis, deemphasis               celt/celt_decoder.c:225-254, 260-355, 363-440
 *   comb_filter, comb_filter_const           celt/celt.c:162-249
 *   the decoder's buffer handling and post-filter state   celt/celt_decoder.c:915-923, 1066-1071, 1107-1132, 1168
 */
#ifndef ANM_CELT_SYNTH_H_INCLUDED
#define ANM_CELT_SYNTH_H_INCLUDED

#include "anm_celt_entropy.h"

/* On the GPU a WARP works on a frame (k_celt_blocks) or on a channel of a stream (k_celt_overlap): the loops over coefficients / samples are split
 * over the lanes (lane, lane + nl, ...) with the data in shared memory and CS_SYNC between the phases; what is a recurrence (post-filter, de-emphasis)
 * runs on lane 0.  The host harness runs the same code with one lane. */
#ifdef __CUDA_ARCH__
#define CS_SYNC() __syncwarp()
#else
#define CS_SYNC() ((void)0)
#endif

#define CS_OVERLAP 120
#define CS_BUF 2048 /* DECODE_BUFFER_SIZE */
#define CS_SIG_SAT 300000000

typedef struct cs_cpx {
    int32_t r, i;
} cs_cpx_t;

ANM_CE_FN int32_t cs_smul(int32_t a, int16_t b) { return (int32_t)(((int64_t)b * a) >> 15); } /* S_MUL(a, b) = MULT16_32_Q15(b, a) */
ANM_CE_FN int32_t cs_add(int32_t a, int32_t b) { return (int32_t)((uint32_t)a + (uint32_t)b); } /* ADD32_ovflw */
ANM_CE_FN int32_t cs_sub(int32_t a, int32_t b) { return (int32_t)((uint32_t)a - (uint32_t)b); }
ANM_CE_FN int32_t cs_neg(int32_t a) { return (int32_t)(0u - (uint32_t)a); }
ANM_CE_FN int32_t cs_sat(int32_t x) { return x > CS_SIG_SAT ? CS_SIG_SAT : x < -CS_SIG_SAT ? -CS_SIG_SAT : x; }

/* ---------------------------------------------------------------- denormalise_bands (downsample = 1) */
/* the gain of band i: g * 2^-shift (celt_exp2 of the band's log energy plus its mean, celt/bands.c:203-243) */
ANM_CE_FN void cs_band_gain(const anm_celt_synth_tables_t *st, const int16_t *band_log_e, int i, int16_t *g_out, int *shift_out) {
    int32_t lg32 = (int32_t)band_log_e[i] + (int32_t)((uint32_t)(int32_t)st->e_means[i] << 6);
    const int16_t lg = (int16_t)(lg32 > 32767 ? 32767 : lg32 < -32768 ? -32768 : lg32); /* SATURATE16 */
    int shift = 16 - (lg >> 10);
    int16_t g;
    if (shift > 31) {
        shift = 0;
        g = 0;
    } else {
        const int16_t frac = (int16_t)((uint16_t)(int16_t)(lg & 1023) << 4); /* celt_exp2_frac */
        g = CV_A16(16383, CV_Q15(frac, CV_A16(22804, CV_Q15(frac, CV_A16(14819, CV_Q15(10204, frac))))));
    }
    if (shift <= -2) { /* a cap on extreme gains: only a corrupted stream gets here */
        g = 16384;
        shift = -2;
    }
    *g_out = g;
    *shift_out = shift;
}
ANM_CE_FN void cs_denormalise(const anm_celt_tables_t *t, const anm_celt_synth_tables_t *st, const int16_t *X, int32_t *freq, const int16_t *band_log_e,
                              int end, int M, int silence, int lane, int nl) {
    const int N = M * 120;
    int bound = M * t->ebands[end];
    if (silence) {
        bound = 0;
        end = 0;
    }
    CS_SYNC();
#ifdef __CUDA_ARCH__
    if (nl == 32) {
        /* a warp: lane i works out the gain of band i, then ONE pass over the coefficients with every load in flight at once -- band by band each
         * band's few loads had to come back from memory before the next band's were issued (a third of k_celt_blocks' time) */
        int16_t g = 0;
        int shift = 0;
        if (lane < end) cs_band_gain(st, band_log_e, lane, &g, &shift);
        int b = 0, next = M * t->ebands[1];
#pragma unroll 4
        for (int j0 = 0; j0 < bound; j0 += 32) {
            const int j = j0 + lane;
            const bool act = j < bound;
            const int16_t x = act ? X[j] : (int16_t)0;
            while (act && j >= next) next = M * t->ebands[++b + 1];
            const int gb = __shfl_sync(0xFFFFFFFFu, (int)g, b), sb = __shfl_sync(0xFFFFFFFFu, shift, b);
            if (act) freq[j] = sb < 0 ? (int32_t)((uint32_t)CV_M16(x, gb) << -sb) : CV_M16(x, gb) >> sb;
        }
    } else
#endif
    for (int i = 0; i < end; i++) {
        const int j0 = M * t->ebands[i], band_end = M * t->ebands[i + 1];
        int16_t g;
        int shift;
        cs_band_gain(st, band_log_e, i, &g, &shift);
        if (shift < 0) {
            for (int j = j0 + lane; j < band_end; j += nl) freq[j] = (int32_t)((uint32_t)CV_M16(X[j], g) << -shift);
        } else {
            for (int j = j0 + lane; j < band_end; j += nl) freq[j] = CV_M16(X[j], g) >> shift;
        }
    }
    for (int i = bound + lane; i < N; i += nl) freq[i] = 0;
    CS_SYNC();
}

/* ---------------------------------------------------------------- the FFT of the inverse MDCT (opus_fft_impl) */
#define CS_CMUL(m, a, tr, ti)                                  \
    do {                                                        \
        (m).r = cs_sub(cs_smul((a).r, tr), cs_smul((a).i, ti)); \
        (m).i = cs_add(cs_smul((a).r, ti), cs_smul((a).i, tr)); \
    } while (0)

ANM_CE_FN void cs_bfly2(cs_cpx_t *Fbeg, int m, int N, int lane, int nl) {
    const int16_t tw = 23170; /* QCONST16(0.7071067812f, 15) */
    (void)m;                  /* m == 4: the radix 2 always follows a radix 4 in these transforms */
    for (int i = lane; i < N; i += nl) {
        cs_cpx_t *F = Fbeg + 8 * i, *F2 = F + 4, t;
        t = F2[0];
        F2[0].r = cs_sub(F[0].r, t.r); F2[0].i = cs_sub(F[0].i, t.i);
        F[0].r = cs_add(F[0].r, t.r); F[0].i = cs_add(F[0].i, t.i);
        t.r = cs_smul(cs_add(F2[1].r, F2[1].i), tw);
        t.i = cs_smul(cs_sub(F2[1].i, F2[1].r), tw);
        F2[1].r = cs_sub(F[1].r, t.r); F2[1].i = cs_sub(F[1].i, t.i);
        F[1].r = cs_add(F[1].r, t.r); F[1].i = cs_add(F[1].i, t.i);
        t.r = F2[2].i;
        t.i = -F2[2].r;
        F2[2].r = cs_sub(F[2].r, t.r); F2[2].i = cs_sub(F[2].i, t.i);
        F[2].r = cs_add(F[2].r, t.r); F[2].i = cs_add(F[2].i, t.i);
        t.r = cs_smul(cs_sub(F2[3].i, F2[3].r), tw);
        t.i = cs_smul(cs_neg(cs_add(F2[3].i, F2[3].r)), tw);
        F2[3].r = cs_sub(F[3].r, t.r); F2[3].i = cs_sub(F[3].i, t.i);
        F[3].r = cs_add(F[3].r, t.r); F[3].i = cs_add(F[3].i, t.i);
    }
}
ANM_CE_FN void cs_bfly4(cs_cpx_t *beg, int fstride, const int16_t *tw, int m, int N, int mm, int lane, int nl) {
    if (m == 1) {
        for (int i = lane; i < N; i += nl) {
            cs_cpx_t *Fout = beg + 4 * i;
            cs_cpx_t s0, s1;
            s0.r = cs_sub(Fout[0].r, Fout[2].r); s0.i = cs_sub(Fout[0].i, Fout[2].i);
            Fout[0].r = cs_add(Fout[0].r, Fout[2].r); Fout[0].i = cs_add(Fout[0].i, Fout[2].i);
            s1.r = cs_add(Fout[1].r, Fout[3].r); s1.i = cs_add(Fout[1].i, Fout[3].i);
            Fout[2].r = cs_sub(Fout[0].r, s1.r); Fout[2].i = cs_sub(Fout[0].i, s1.i);
            Fout[0].r = cs_add(Fout[0].r, s1.r); Fout[0].i = cs_add(Fout[0].i, s1.i);
            s1.r = cs_sub(Fout[1].r, Fout[3].r); s1.i = cs_sub(Fout[1].i, Fout[3].i);
            Fout[1].r = cs_add(s0.r, s1.i);
            Fout[1].i = cs_sub(s0.i, s1.r);
            Fout[3].r = cs_sub(s0.r, s1.i);
            Fout[3].i = cs_add(s0.i, s1.r);
        }
        return;
    }
    const int m2 = 2 * m, m3 = 3 * m;
    for (int idx = lane; idx < N * m; idx += nl) {
        const int i = idx / m, j = idx % m;
        cs_cpx_t *F = beg + i * mm + j;
        const int16_t *tw1 = tw + 2 * fstride * j, *tw2 = tw + 2 * fstride * 2 * j, *tw3 = tw + 2 * fstride * 3 * j;
        cs_cpx_t s0, s1, s2, s3, s4, s5;
        CS_CMUL(s0, F[m], tw1[0], tw1[1]);
        CS_CMUL(s1, F[m2], tw2[0], tw2[1]);
        CS_CMUL(s2, F[m3], tw3[0], tw3[1]);
        s5.r = cs_sub(F->r, s1.r); s5.i = cs_sub(F->i, s1.i);
        F->r = cs_add(F->r, s1.r); F->i = cs_add(F->i, s1.i);
        s3.r = cs_add(s0.r, s2.r); s3.i = cs_add(s0.i, s2.i);
        s4.r = cs_sub(s0.r, s2.r); s4.i = cs_sub(s0.i, s2.i);
        F[m2].r = cs_sub(F->r, s3.r); F[m2].i = cs_sub(F->i, s3.i);
        F->r = cs_add(F->r, s3.r); F->i = cs_add(F->i, s3.i);
        F[m].r = cs_add(s5.r, s4.i);
        F[m].i = cs_sub(s5.i, s4.r);
        F[m3].r = cs_sub(s5.r, s4.i);
        F[m3].i = cs_add(s5.i, s4.r);
    }
}
ANM_CE_FN void cs_bfly3(cs_cpx_t *beg, int fstride, const int16_t *tw, int m, int N, int mm, int lane, int nl) {
    const int m2 = 2 * m;
    const int16_t epi3_i = -28378;
    for (int idx = lane; idx < N * m; idx += nl) {
        const int i = idx / m, j = idx % m;
        cs_cpx_t *F = beg + i * mm + j;
        const int16_t *tw1 = tw + 2 * fstride * j, *tw2 = tw + 2 * fstride * 2 * j;
        cs_cpx_t s0, s1, s2, s3;
        CS_CMUL(s1, F[m], tw1[0], tw1[1]);
        CS_CMUL(s2, F[m2], tw2[0], tw2[1]);
        s3.r = cs_add(s1.r, s2.r); s3.i = cs_add(s1.i, s2.i);
        s0.r = cs_sub(s1.r, s2.r); s0.i = cs_sub(s1.i, s2.i);
        F[m].r = cs_sub(F->r, s3.r >> 1);
        F[m].i = cs_sub(F->i, s3.i >> 1);
        s0.r = cs_smul(s0.r, epi3_i);
        s0.i = cs_smul(s0.i, epi3_i);
        F->r = cs_add(F->r, s3.r); F->i = cs_add(F->i, s3.i);
        F[m2].r = cs_add(F[m].r, s0.i);
        F[m2].i = cs_sub(F[m].i, s0.r);
        F[m].r = cs_sub(F[m].r, s0.i);
        F[m].i = cs_add(F[m].i, s0.r);
    }
}
ANM_CE_FN void cs_bfly5(cs_cpx_t *beg, int fstride, const int16_t *tw, int m, int N, int mm, int lane, int nl) {
    const int16_t ya_r = 10126, ya_i = -31164, yb_r = -26510, yb_i = -19261;
    for (int idx = lane; idx < N * m; idx += nl) {
        const int i = idx / m, u = idx % m;
        cs_cpx_t *F0 = beg + i * mm + u, *F1 = F0 + m, *F2 = F0 + 2 * m, *F3 = F0 + 3 * m, *F4 = F0 + 4 * m;
        cs_cpx_t s0, s1, s2, s3, s4, s5, s6, s7, s8, s9, s10, s11, s12;
        s0 = *F0;
        CS_CMUL(s1, *F1, tw[2 * u * fstride], tw[2 * u * fstride + 1]);
        CS_CMUL(s2, *F2, tw[2 * 2 * u * fstride], tw[2 * 2 * u * fstride + 1]);
        CS_CMUL(s3, *F3, tw[2 * 3 * u * fstride], tw[2 * 3 * u * fstride + 1]);
        CS_CMUL(s4, *F4, tw[2 * 4 * u * fstride], tw[2 * 4 * u * fstride + 1]);
        s7.r = cs_add(s1.r, s4.r); s7.i = cs_add(s1.i, s4.i);
        s10.r = cs_sub(s1.r, s4.r); s10.i = cs_sub(s1.i, s4.i);
        s8.r = cs_add(s2.r, s3.r); s8.i = cs_add(s2.i, s3.i);
        s9.r = cs_sub(s2.r, s3.r); s9.i = cs_sub(s2.i, s3.i);
        F0->r = cs_add(F0->r, cs_add(s7.r, s8.r));
        F0->i = cs_add(F0->i, cs_add(s7.i, s8.i));
        s5.r = cs_add(s0.r, cs_add(cs_smul(s7.r, ya_r), cs_smul(s8.r, yb_r)));
        s5.i = cs_add(s0.i, cs_add(cs_smul(s7.i, ya_r), cs_smul(s8.i, yb_r)));
        s6.r = cs_add(cs_smul(s10.i, ya_i), cs_smul(s9.i, yb_i));
        s6.i = cs_neg(cs_add(cs_smul(s10.r, ya_i), cs_smul(s9.r, yb_i)));
        F1->r = cs_sub(s5.r, s6.r); F1->i = cs_sub(s5.i, s6.i);
        F4->r = cs_add(s5.r, s6.r); F4->i = cs_add(s5.i, s6.i);
        s11.r = cs_add(s0.r, cs_add(cs_smul(s7.r, yb_r), cs_smul(s8.r, ya_r)));
        s11.i = cs_add(s0.i, cs_add(cs_smul(s7.i, yb_r), cs_smul(s8.i, ya_r)));
        s12.r = cs_sub(cs_smul(s9.i, ya_i), cs_smul(s10.i, yb_i));
        s12.i = cs_sub(cs_smul(s10.r, yb_i), cs_smul(s9.r, ya_i));
        F2->r = cs_add(s11.r, s12.r); F2->i = cs_add(s11.i, s12.i);
        F3->r = cs_sub(s11.r, s12.r); F3->i = cs_sub(s11.i, s12.i);
    }
}
/* the transform of 480 >> k complex points, k = 0..3 (the four kiss_fft states of the 48 kHz mode, celt/static_modes_fixed.h:432-498); the butterflies of a
 * stage are independent of each other and go over the lanes */
ANM_CE_FN void cs_fft(const anm_celt_synth_tables_t *st, int k, cs_cpx_t *fout, int lane, int nl) {
    const int8_t factors[4][10] = {{5, 96, 3, 32, 4, 8, 2, 4, 4, 1}, {5, 48, 3, 16, 4, 4, 4, 1, 0, 0}, {5, 24, 3, 8, 2, 4, 4, 1, 0, 0}, {5, 12, 3, 4, 4, 1, 0, 0, 0, 0}};
    const int8_t *fac = factors[k];
    const int shift = k; /* st->shift: -1 (taken as 0), 1, 2, 3 */
    int fstride[6], L = 0, m;
    fstride[0] = 1;
    do {
        const int p = fac[2 * L];
        m = fac[2 * L + 1];
        fstride[L + 1] = fstride[L] * p;
        L++;
    } while (m != 1);
    m = fac[2 * L - 1];
    for (int i = L - 1; i >= 0; i--) {
        const int m2 = i != 0 ? fac[2 * i - 1] : 1;
        CS_SYNC();
        switch (fac[2 * i]) {
            case 2: cs_bfly2(fout, m, fstride[i], lane, nl); break;
            case 4: cs_bfly4(fout, fstride[i] << shift, st->fft_tw, m, fstride[i], m2, lane, nl); break;
            case 3: cs_bfly3(fout, fstride[i] << shift, st->fft_tw, m, fstride[i], m2, lane, nl); break;
            case 5: cs_bfly5(fout, fstride[i] << shift, st->fft_tw, m, fstride[i], m2, lane, nl); break;
        }
        m = m2;
    }
    CS_SYNC();
}

/* ---------------------------------------------------------------- inverse MDCT of one block, up to (not including) the window mix */
/* in: the block's N2 = 960 >> shift coefficients, `stride` apart; raw: N2 values -- what clt_mdct_backward holds in out[overlap / 2 .. overlap / 2 + N2)
 * before it mirrors the block's ends against the previous block's tail */
ANM_CE_FN void cs_imdct_raw(const anm_celt_synth_tables_t *st, const int32_t *in, int stride, int shift, int32_t *raw, int lane, int nl) {
    int N = 1920;
    const int16_t *trig = st->trig;
    const int16_t *bitrev = st->bitrev;
    for (int i = 0; i < shift; i++) {
        N >>= 1;
        trig += N;
        bitrev += N >> 1; /* 480, 240, 120 entries */
    }
    const int N2 = N >> 1, N4 = N >> 2;
    CS_SYNC();
    for (int i = lane; i < N4; i += nl) {
        const int32_t xp1 = in[2 * stride * i], xp2 = in[stride * (N2 - 1) - 2 * stride * i];
        const int rev = bitrev[i];
        const int32_t yr = cs_add(cs_smul(xp2, trig[i]), cs_smul(xp1, trig[N4 + i]));
        const int32_t yi = cs_sub(cs_smul(xp1, trig[i]), cs_smul(xp2, trig[N4 + i]));
        raw[2 * rev + 1] = yr; /* real and imaginary swapped: an FFT instead of an IFFT */
        raw[2 * rev] = yi;
    }
    cs_fft(st, shift, (cs_cpx_t *)raw, lane, nl);
    /* post-rotation from both ends at once: pair i touches words 2i, 2i + 1, N2 - 2 - 2i, N2 - 1 - 2i only (N4 is even in all four transforms) */
    for (int i = lane; i < (N4 + 1) >> 1; i += nl) {
        int32_t *yp0 = raw + 2 * i, *yp1 = raw + N2 - 2 - 2 * i;
        int32_t re = yp0[1], im = yp0[0];
        int16_t t0 = trig[i], t1 = trig[N4 + i];
        int32_t yr = cs_add(cs_smul(re, t0), cs_smul(im, t1));
        int32_t yi = cs_sub(cs_smul(re, t1), cs_smul(im, t0));
        re = yp1[1];
        im = yp1[0];
        yp0[0] = yr;
        yp1[1] = yi;
        t0 = trig[N4 - i - 1];
        t1 = trig[N2 - i - 1];
        yr = cs_add(cs_smul(re, t0), cs_smul(im, t1));
        yi = cs_sub(cs_smul(re, t1), cs_smul(im, t0));
        yp1[0] = yr;
        yp0[1] = yi;
    }
    CS_SYNC();
}
/* the window mix of a block whose raw output sits at out + overlap / 2: out[0 .. overlap / 2) still holds the previous block's tail */
ANM_CE_FN void cs_mirror(const anm_celt_synth_tables_t *st, int32_t *out, int lane, int nl) {
    for (int i = lane; i < CS_OVERLAP / 2; i += nl) {
        const int32_t x1 = out[CS_OVERLAP - 1 - i], x2 = out[i];
        const int16_t w1 = st->window[i], w2 = st->window[CS_OVERLAP - 1 - i];
        out[i] = cs_sub(cs_smul(x2, w2), cs_smul(x1, w1));
        out[CS_OVERLAP - 1 - i] = cs_add(cs_smul(x2, w1), cs_smul(x1, w2));
    }
}

/* ---------------------------------------------------------------- frame-parallel part: the raw blocks of every output channel */
/* The raw blocks of ONE output channel c of a frame.  X: [C][N] normalised spectrum (stage 2), band_e: the frame's band energies [2][21]; freq: N words of
 * scratch; rawc: N words -- block b at rawc + b * (N / B) -- which also hold the second channel's coefficients on the way to the mono downmix of a
 * stereo frame.  A mono frame feeds both channels of a stereo output (celt_synthesis, celt_decoder.c:434-489). */
ANM_CE_FN void cs_channel_blocks(const anm_celt_tables_t *t, const anm_celt_synth_tables_t *st, const int16_t *X, const int16_t *band_e, int C, int CC, int c, int LM,
                                 int end, int transient, int silence, int32_t *freq, int32_t *rawc, int lane, int nl) {
    const int M = 1 << LM, N = 120 << LM;
    const int B = transient ? M : 1, NB = transient ? 120 : N, shift = transient ? 3 : 3 - LM;
    if (CC == 1 && C == 2) {
        cs_denormalise(t, st, X, freq, band_e, end, M, silence, lane, nl);
        cs_denormalise(t, st, X + N, rawc, band_e + ANM_CE_NB, end, M, silence, lane, nl);
        for (int i = lane; i < N; i += nl) freq[i] = (freq[i] >> 1) + (rawc[i] >> 1);
    } else {
        const int sc = C == 1 ? 0 : c;
        cs_denormalise(t, st, X + sc * N, freq, band_e + sc * ANM_CE_NB, end, M, silence, lane, nl);
    }
    for (int b = 0; b < B; b++) cs_imdct_raw(st, freq + b, B, shift, rawc + NB * b, lane, nl);
    CS_SYNC();
}
/* all output channels of a frame (the host-side test harness; the kernel runs one warp per channel): raw: [CC][N] */
ANM_CE_FN void cs_frame_blocks(const anm_celt_tables_t *t, const anm_celt_synth_tables_t *st, const int16_t *X, const int16_t *band_e, int C, int CC, int LM,
                               int end, int transient, int silence, int32_t *freq, int32_t *raw, int lane, int nl) {
    for (int c = 0; c < CC; c++) cs_channel_blocks(t, st, X, band_e, C, CC, c, LM, end, transient, silence, freq, raw + c * (120 << LM), lane, nl);
}

/* ---------------------------------------------------------------- per-stream part */
/* The pitch post-filter, in place.  y[i] takes y[i - T .. i - T +- 2] with T >= 15: a recurrence, but one that reaches back at least 13 samples, so
 * runs of min(T0, T1) - 2 outputs are independent of each other and go over the lanes (the reference's sliding registers x0..x4 hold exactly
 * y[i - T1 + 2 .. i - T1 - 2] as they stand when y[i] is computed). */
ANM_CE_FN void cs_comb_filter(const anm_celt_synth_tables_t *st, int32_t *y, int T0, int T1, int N, int16_t g0, int16_t g1, int tapset0, int tapset1, int overlap,
                              int lane, int nl) {
    const int16_t gains[3][3] = {{10048, 7112, 4248}, {15200, 8784, 0}, {26208, 3280, 0}};
    if (g0 == 0 && g1 == 0) return; /* in place: nothing to move */
    T0 = ce_imax(T0, 15);
    T1 = ce_imax(T1, 15);
    const int16_t g00 = (int16_t)CV_P15(g0, gains[tapset0][0]), g01 = (int16_t)CV_P15(g0, gains[tapset0][1]), g02 = (int16_t)CV_P15(g0, gains[tapset0][2]);
    const int16_t g10 = (int16_t)CV_P15(g1, gains[tapset1][0]), g11 = (int16_t)CV_P15(g1, gains[tapset1][1]), g12 = (int16_t)CV_P15(g1, gains[tapset1][2]);
    if (g0 == g1 && T0 == T1 && tapset0 == tapset1) overlap = 0;
    const int run = (T0 < T1 ? T0 : T1) - 2; /* >= 13 */
    CS_SYNC();
    for (int base = 0; base < overlap; base += run) {
        const int lim = base + run < overlap ? base + run : overlap;
        for (int i = base + lane; i < lim; i += nl) {
            const int16_t f = (int16_t)CV_Q15(st->window[i], st->window[i]);
            const int16_t nf = (int16_t)(32767 - f);
            const int32_t v = y[i] + cs_smul(y[i - T0], (int16_t)CV_Q15(nf, g00)) + cs_smul(y[i - T0 + 1] + y[i - T0 - 1], (int16_t)CV_Q15(nf, g01)) +
                              cs_smul(y[i - T0 + 2] + y[i - T0 - 2], (int16_t)CV_Q15(nf, g02)) + cs_smul(y[i - T1], (int16_t)CV_Q15(f, g10)) +
                              cs_smul(y[i - T1 + 1] + y[i - T1 - 1], (int16_t)CV_Q15(f, g11)) + cs_smul(y[i - T1 + 2] + y[i - T1 - 2], (int16_t)CV_Q15(f, g12));
            y[i] = cs_sat(v);
        }
        CS_SYNC();
    }
    if (g1 == 0) return;
    /* the part with the constant filter */
    const int run1 = T1 - 2;
    for (int base = overlap; base < N; base += run1) {
        const int lim = base + run1 < N ? base + run1 : N;
        for (int i = base + lane; i < lim; i += nl) {
            const int32_t v = y[i] + cs_smul(y[i - T1], g10) + cs_smul(y[i - T1 + 1] + y[i - T1 - 1], g11) + cs_smul(y[i - T1 + 2] + y[i - T1 - 2], g12);
            y[i] = cs_sat(v);
        }
        CS_SYNC();
    }
}

/* the post-filter parameters a frame carries (celt/celt_decoder.c:968-983) */
typedef struct cs_pf {
    int32_t period, period_old, tapset, tapset_old;
    int16_t gain, gain_old;
} cs_pf_t;

/* One frame of ONE output channel c of a stream: raw blocks (cs_frame_blocks) -> PCM (cs_channel_frame below), in two parts.  mem: the channel's
 * output history (anm_celt_synth_t.mem[c]); pf: the stream's post-filter state, updated for the next frame (every channel of a stream sees and makes
 * the same updates).  cs_channel_signal, the part in front of the de-emphasis, leaves the frame's N filtered samples at mem + CS_BUF - N (where the
 * next frame's history move picks them up): the copies, the window mix, the saturation and (in runs shorter than its period) the post-filter go over
 * the lanes.  cs_deemphasis is a one-pole recurrence over the samples in order: one lane (the GPU runs it in a kernel of its own, a thread per
 * channel, instead of holding 31 lanes of a warp idle for it). */
ANM_CE_FN void cs_channel_signal(const anm_celt_synth_tables_t *st, int32_t *mem, cs_pf_t *pf, const anm_celt_frame_t *fr, const int32_t *raw_c, int lane, int nl) {
    const int LM = fr->lm, N = 120 << LM, transient = (fr->flags & ANM_CELT_F_TRANSIENT) != 0;
    const int B = transient ? 1 << LM : 1, NB = transient ? 120 : N;
    const int pf_on = (fr->flags & ANM_CELT_F_POSTFILTER) != 0;
    const int pf_pitch = pf_on ? fr->pf_pitch : 0, pf_tapset = pf_on ? fr->pf_tapset : 0;
    const int16_t pf_gain = pf_on ? (int16_t)(3072 * (fr->pf_gain_q + 1)) : 0; /* QCONST16(.09375f, 15) * (qg + 1) */
    /* the history moves up by one frame (half of the overlap is still to be mixed): in chunks of nl, every chunk read before it is written */
    CS_SYNC();
    for (int i0 = 0; i0 < CS_BUF - N + CS_OVERLAP / 2; i0 += nl) {
        const int i = i0 + lane;
        const int32_t v = i < CS_BUF - N + CS_OVERLAP / 2 ? mem[i + N] : 0;
        CS_SYNC();
        if (i < CS_BUF - N + CS_OVERLAP / 2) mem[i] = v;
        CS_SYNC();
    }
    int32_t *out = mem + CS_BUF - N;
    for (int b = 0; b < B; b++) {
        int32_t *ob = out + NB * b;
        const int32_t *rb = raw_c + NB * b;
        for (int i = lane; i < NB; i += nl) ob[CS_OVERLAP / 2 + i] = rb[i];
        CS_SYNC();
        cs_mirror(st, ob, lane, nl);
        CS_SYNC();
    }
    for (int i = lane; i < N; i += nl) out[i] = cs_sat(out[i]);
    CS_SYNC();
    /* pitch post-filter: the first 120 samples fade from the filter before the previous frame's to the previous frame's, the rest to this frame's */
    pf->period = ce_imax(pf->period, 15);
    pf->period_old = ce_imax(pf->period_old, 15);
    cs_comb_filter(st, out, pf->period_old, pf->period, 120, pf->gain_old, pf->gain, pf->tapset_old, pf->tapset, CS_OVERLAP, lane, nl);
    if (LM != 0) cs_comb_filter(st, out + 120, pf->period, pf_pitch, N - 120, pf->gain, pf_gain, pf->tapset, pf_tapset, CS_OVERLAP, lane, nl);
    pf->period_old = pf->period;
    pf->gain_old = pf->gain;
    pf->tapset_old = pf->tapset;
    pf->period = pf_pitch;
    pf->gain = pf_gain;
    pf->tapset = pf_tapset;
    if (LM != 0) {
        pf->period_old = pf->period;
        pf->gain_old = pf->gain;
        pf->tapset_old = pf->tapset;
    }
    CS_SYNC();
}
/* de-emphasis to 16 bits (celt_decoder.c:266-339, the path without downsampling): a one-pole recurrence over the samples of a channel, in order.
 * sig: the N filtered samples; pcm: [N][CC]. */
ANM_CE_FN int16_t cs_deemphasis_step(int32_t x, int32_t *m) {
    const int32_t tmp = x + *m;       /* VERY_SMALL = 0 */
    *m = cs_smul(tmp, 27853);         /* mode->preemph[0] */
    int32_t v = CV_PSHR32(tmp, 12);   /* SIG2WORD16 */
    v = v < -32768 ? -32768 : v > 32767 ? 32767 : v;
    return (int16_t)v;
}
ANM_CE_FN void cs_deemphasis(const int32_t *sig, int N, int32_t *preemph_mem, int16_t *pcm, int CC, int c) {
    int32_t m = *preemph_mem;
    for (int j = 0; j < N; j++) pcm[j * CC + c] = cs_deemphasis_step(sig[j], &m);
    *preemph_mem = m;
}
ANM_CE_FN void cs_channel_frame(const anm_celt_synth_tables_t *st, int32_t *mem, int32_t *preemph_mem, cs_pf_t *pf, const anm_celt_frame_t *fr, const int32_t *raw_c,
                                int CC, int c, int16_t *pcm, int lane, int nl) {
    const int N = 120 << fr->lm;
    cs_channel_signal(st, mem, pf, fr, raw_c, lane, nl);
    if (lane == 0) cs_deemphasis(mem + CS_BUF - N, N, preemph_mem, pcm, CC, c);
    CS_SYNC();
}
ANM_CE_FN void cs_pf_load(cs_pf_t *pf, const anm_celt_synth_t *s) {
    pf->period = s->pf_period; pf->period_old = s->pf_period_old; pf->tapset = s->pf_tapset; pf->tapset_old = s->pf_tapset_old;
    pf->gain = s->pf_gain; pf->gain_old = s->pf_gain_old;
}
ANM_CE_FN void cs_pf_store(const cs_pf_t *pf, anm_celt_synth_t *s) {
    s->pf_period = pf->period; s->pf_period_old = pf->period_old; s->pf_tapset = pf->tapset; s->pf_tapset_old = pf->tapset_old;
    s->pf_gain = pf->gain; s->pf_gain_old = pf->gain_old;
}
/* all channels of a stream's frame (the host-side test harness; the kernel runs one thread per channel) */
ANM_CE_FN void cs_stream_frame(const anm_celt_synth_tables_t *st, anm_celt_synth_t *s, const anm_celt_frame_t *fr, const int32_t *raw, int CC, int16_t *pcm) {
    const int N = 120 << fr->lm;
    cs_pf_t pf0, pf;
    cs_pf_load(&pf0, s);
    pf = pf0;
    for (int c = 0; c < CC; c++) {
        pf = pf0;
        cs_channel_frame(st, s->mem[c], &s->preemph_mem[c], &pf, fr, raw + c * N, CC, c, pcm, 0, 1);
    }
    cs_pf_store(&pf, s);
}

#endif /* ANM_CELT_SYNTH_H_INCLUDED */
