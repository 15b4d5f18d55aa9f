/*
 * ref_celt_shim.c -- stage-by-stage trace of the REFERENCE's CELT entropy decode, linked into oracle/_ref/libref_opus.so
 * (compiled in place from /root/reference/hardware/lib/libopus/src by oracle/Makefile).  TEST INFRASTRUCTURE ONLY.
 *
 * It drives the reference's OWN functions -- ec_dec_*, unquant_coarse_energy, clt_compute_allocation, unquant_fine_energy,
 * quant_all_bands, unquant_energy_finalise -- in the order celt_decode_with_ec calls them (celt/celt_decoder.c:946-1095) and
 * records the range coder's state after each, so that a restatement that goes wrong can be located.  The only logic restated
 * here is what celt_decoder.c keeps static (tf_decode, :441-478) and the glue between the calls.
 * Also exports the static tables of the 48 kHz / 960 mode so that tables computed elsewhere can be compared with them.
 */
#ifdef HAVE_CONFIG_H
#include "config.h"
#endif
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "opus.h"
#include "opus_custom.h"
#include "celt.h"
#include "modes.h"
#include "entdec.h"
#include "quant_bands.h"
#include "rate.h"
#include "bands.h"

typedef struct {
    uint32_t rng[8]; /* after: header flags, coarse energy, tf, spread + dynalloc + trim, allocation, fine energy, bands, finalise */
    int32_t tell[8];
    int32_t silence, postfilter, pf_pitch, pf_qg, pf_tapset, transient, intra, spread, alloc_trim, intensity, dual_stereo, coded_bands, anti_collapse_on;
    int32_t balance;
    int32_t tf_res[21], offsets[21], cap[21], pulses[21], fine_quant[21], fine_priority[21];
    int16_t band_e[42];
} ref_celt_trace_t;

static void tf_decode_(int start, int end, int isTransient, int *tf_res, int LM, ec_dec *dec) {
    int i, curr, tf_select, tf_select_rsv, tf_changed, logp;
    opus_uint32 budget = dec->storage * 8, tell = ec_tell(dec);
    logp = isTransient ? 2 : 4;
    tf_select_rsv = LM > 0 && tell + logp + 1 <= budget;
    budget -= tf_select_rsv;
    tf_changed = curr = 0;
    for (i = start; i < end; i++) {
        if (tell + logp <= budget) {
            curr ^= ec_dec_bit_logp(dec, logp);
            tell = ec_tell(dec);
            tf_changed |= curr;
        }
        tf_res[i] = curr;
        logp = isTransient ? 4 : 5;
    }
    tf_select = 0;
    if (tf_select_rsv && tf_select_table[LM][4 * isTransient + 0 + tf_changed] != tf_select_table[LM][4 * isTransient + 2 + tf_changed])
        tf_select = ec_dec_bit_logp(dec, 1);
    for (i = start; i < end; i++) tf_res[i] = tf_select_table[LM][4 * isTransient + 2 * tf_select + tf_res[i]];
}

#define MARK(k) do { tr->rng[k] = dec.rng; tr->tell[k] = ec_tell(&dec); } while (0)

/* x_out (may be NULL): the normalised spectrum quant_all_bands leaves, [2][960] celt_norm (Q14), channel c at x_out + 960 c, N = 120 << LM
 * coefficients each; cm_out: the 42 collapse masks; seed_in / seed_out: the decoder's noise seed (CELTDecoder.rng) before / after the bands */
static int trace_impl(const uint8_t *data, int len, int C, int LM, int end, int16_t *oldBandE /* [42] in/out */, ref_celt_trace_t *tr, uint32_t seed_in,
                      int disable_inv, int16_t *x_out, uint8_t *cm_out, uint32_t *seed_out, const int16_t *prev1logE, const int16_t *prev2logE,
                      int16_t *x_post_out) {
    int err = 0;
    const CELTMode *mode = opus_custom_mode_create(48000, 960, &err);
    if (!mode || len <= 1 || len > 1275) return -1;
    const int nbEBands = mode->nbEBands, start = 0, M = 1 << LM, N = M * mode->shortMdctSize;
    const opus_int16 *eBands = mode->eBands;
    static const unsigned char trim_icdf_[11] = {126, 124, 119, 109, 87, 41, 19, 9, 4, 2, 0};
    static const unsigned char spread_icdf_[4] = {25, 23, 2, 0};
    static const unsigned char tapset_icdf_[3] = {2, 1, 0};
    ec_dec dec;
    int i, c;
    memset(tr, 0, sizeof *tr);
    ec_dec_init(&dec, (unsigned char *)data, len);
    if (C == 1)
        for (i = 0; i < nbEBands; i++) oldBandE[i] = MAX16(oldBandE[i], oldBandE[nbEBands + i]);
    opus_int32 total_bits = len * 8, tell = ec_tell(&dec), bits, balance;
    int silence;
    if (tell >= total_bits) silence = 1;
    else if (tell == 1) silence = ec_dec_bit_logp(&dec, 15);
    else silence = 0;
    if (silence) {
        tell = len * 8;
        dec.nbits_total += tell - ec_tell(&dec);
    }
    if (start == 0 && tell + 16 <= total_bits) {
        if (ec_dec_bit_logp(&dec, 1)) {
            int octave = ec_dec_uint(&dec, 6);
            tr->postfilter = 1;
            tr->pf_pitch = (16 << octave) + ec_dec_bits(&dec, 4 + octave) - 1;
            tr->pf_qg = ec_dec_bits(&dec, 3);
            if (ec_tell(&dec) + 2 <= total_bits) tr->pf_tapset = ec_dec_icdf(&dec, tapset_icdf_, 2);
        }
        tell = ec_tell(&dec);
    }
    int isTransient = 0;
    if (LM > 0 && tell + 3 <= total_bits) {
        isTransient = ec_dec_bit_logp(&dec, 3);
        tell = ec_tell(&dec);
    }
    const int shortBlocks = isTransient ? M : 0;
    const int intra_ener = tell + 3 <= total_bits ? ec_dec_bit_logp(&dec, 3) : 0;
    MARK(0);
    unquant_coarse_energy(mode, start, end, oldBandE, intra_ener, &dec, C, LM);
    MARK(1);
    int tf_res[21], cap[21], offsets[21], fine_quant[21], pulses[21], fine_priority[21];
    memset(tf_res, 0, sizeof tf_res);
    tf_decode_(start, end, isTransient, tf_res, LM, &dec);
    MARK(2);
    tell = ec_tell(&dec);
    int spread_decision = SPREAD_NORMAL;
    if (tell + 4 <= total_bits) spread_decision = ec_dec_icdf(&dec, spread_icdf_, 5);
    init_caps(mode, cap, LM, C);
    int dynalloc_logp = 6;
    total_bits <<= BITRES;
    tell = ec_tell_frac(&dec);
    for (i = start; i < end; i++) {
        int width = C * (eBands[i + 1] - eBands[i]) << LM;
        int quanta = IMIN(width << BITRES, IMAX(6 << BITRES, width));
        int dynalloc_loop_logp = dynalloc_logp, boost = 0;
        while (tell + (dynalloc_loop_logp << BITRES) < total_bits && boost < cap[i]) {
            int flag = ec_dec_bit_logp(&dec, dynalloc_loop_logp);
            tell = ec_tell_frac(&dec);
            if (!flag) break;
            boost += quanta;
            total_bits -= quanta;
            dynalloc_loop_logp = 1;
        }
        offsets[i] = boost;
        if (boost > 0) dynalloc_logp = IMAX(2, dynalloc_logp - 1);
    }
    int alloc_trim = tell + (6 << BITRES) <= total_bits ? ec_dec_icdf(&dec, trim_icdf_, 7) : 5;
    MARK(3);
    bits = (((opus_int32)len * 8) << BITRES) - ec_tell_frac(&dec) - 1;
    int anti_collapse_rsv = isTransient && LM >= 2 && bits >= ((LM + 2) << BITRES) ? (1 << BITRES) : 0;
    bits -= anti_collapse_rsv;
    int intensity = 0, dual_stereo = 0;
    memset(pulses, 0, sizeof pulses);
    memset(fine_quant, 0, sizeof fine_quant);
    memset(fine_priority, 0, sizeof fine_priority);
    int codedBands = clt_compute_allocation(mode, start, end, offsets, cap, alloc_trim, &intensity, &dual_stereo, bits, &balance, pulses, fine_quant,
                                            fine_priority, C, LM, &dec, 0, 0, 0);
    MARK(4);
    tr->balance = balance;
    unquant_fine_energy(mode, start, end, oldBandE, fine_quant, &dec, C);
    MARK(5);
    celt_norm *X = (celt_norm *)calloc((size_t)2 * N + 64, sizeof(celt_norm));
    unsigned char collapse_masks[42];
    opus_uint32 seed = seed_in;
    memset(collapse_masks, 0, sizeof collapse_masks);
    quant_all_bands(0, mode, start, end, X, C == 2 ? X + N : NULL, collapse_masks, NULL, pulses, shortBlocks, spread_decision, dual_stereo, intensity, tf_res,
                    len * (8 << BITRES) - anti_collapse_rsv, balance, &dec, LM, codedBands, &seed, 0, 0, disable_inv);
    if (x_out) {
        memset(x_out, 0, 2 * 960 * sizeof(int16_t));
        for (c = 0; c < C; c++) memcpy(x_out + 960 * c, X + N * c, (size_t)(M * eBands[end]) * sizeof(int16_t));
    }
    if (cm_out) memcpy(cm_out, collapse_masks, 42);
    if (seed_out) *seed_out = seed;
    MARK(6);
    if (anti_collapse_rsv > 0) tr->anti_collapse_on = ec_dec_bits(&dec, 1);
    unquant_energy_finalise(mode, start, end, oldBandE, fine_quant, fine_priority, len * 8 - ec_tell(&dec), &dec, C);
    MARK(7);
    /* celt_decoder.c:1096-1098: the spectrum synthesis starts from */
    if (x_post_out) {
        if (tr->anti_collapse_on) anti_collapse(mode, X, collapse_masks, LM, C, N, start, end, oldBandE, prev1logE, prev2logE, pulses, seed, 0);
        memset(x_post_out, 0, 2 * 960 * sizeof(int16_t));
        for (c = 0; c < C; c++) memcpy(x_post_out + 960 * c, X + N * c, (size_t)(M * eBands[end]) * sizeof(int16_t));
    }
    free(X);
    if (silence)
        for (i = 0; i < C * nbEBands; i++) oldBandE[i] = -QCONST16(28.f, DB_SHIFT);
    if (C == 1) OPUS_COPY(&oldBandE[nbEBands], oldBandE, nbEBands);
    for (c = 0; c < 2; c++)
        for (i = end; i < nbEBands; i++) oldBandE[c * nbEBands + i] = 0;
    tr->silence = silence;
    tr->transient = isTransient;
    tr->intra = intra_ener;
    tr->spread = spread_decision;
    tr->alloc_trim = alloc_trim;
    tr->intensity = intensity;
    tr->dual_stereo = dual_stereo;
    tr->coded_bands = codedBands;
    for (i = 0; i < 21; i++) {
        tr->tf_res[i] = tf_res[i];
        tr->offsets[i] = i < end ? offsets[i] : 0;
        tr->cap[i] = cap[i];
        tr->pulses[i] = pulses[i];
        tr->fine_quant[i] = fine_quant[i];
        tr->fine_priority[i] = fine_priority[i];
    }
    memcpy(tr->band_e, oldBandE, 42 * sizeof(int16_t));
    return 0;
}

int ref_celt_entropy_trace(const uint8_t *data, int len, int C, int LM, int end, int16_t *oldBandE /* [42] in/out */, ref_celt_trace_t *tr) {
    return trace_impl(data, len, C, LM, end, oldBandE, tr, 0u, 0, NULL, NULL, NULL, NULL, NULL, NULL);
}
/* the same with the spectrum kept: what celt_decode_with_ec holds in X after quant_all_bands (celt/celt_decoder.c:1084-1088) */
int ref_celt_spectrum_trace(const uint8_t *data, int len, int C, int LM, int end, int16_t *oldBandE, ref_celt_trace_t *tr, uint32_t seed_in, int disable_inv,
                            int16_t *x_out, uint8_t *cm_out, uint32_t *seed_out) {
    return trace_impl(data, len, C, LM, end, oldBandE, tr, seed_in, disable_inv, x_out, cm_out, seed_out, NULL, NULL, NULL);
}
/* ... and with anti_collapse() applied (celt/celt_decoder.c:1096-1098) on the log-energy histories given (oldLogE, oldLogE2 of the decoder before
 * the frame: ref_celt_stream_states() reads them off the reference decoder itself): the spectrum celt_synthesis() starts from */
int ref_celt_spectrum_trace2(const uint8_t *data, int len, int C, int LM, int end, int16_t *oldBandE, ref_celt_trace_t *tr, uint32_t seed_in, int disable_inv,
                             int16_t *x_out, uint8_t *cm_out, uint32_t *seed_out, const int16_t *prev1logE, const int16_t *prev2logE, int16_t *x_post_out) {
    return trace_impl(data, len, C, LM, end, oldBandE, tr, seed_in, disable_inv, x_out, cm_out, seed_out, prev1logE, prev2logE, x_post_out);
}

/* static tables of the standard mode: logN[21], cache.index[105], cache.bits[size], cache.caps[168]; returns cache.size */
int ref_celt_mode_tables(int16_t *ebands, int16_t *logn, int16_t *cache_index, uint8_t *cache_bits, uint8_t *cache_caps, uint8_t *alloc) {
    int err = 0;
    const CELTMode *m = opus_custom_mode_create(48000, 960, &err);
    if (!m) return -1;
    memcpy(ebands, m->eBands, 22 * sizeof(int16_t));
    memcpy(logn, m->logN, 21 * sizeof(int16_t));
    memcpy(cache_index, m->cache.index, 105 * sizeof(int16_t));
    memcpy(cache_bits, m->cache.bits, (size_t)m->cache.size);
    memcpy(cache_caps, m->cache.caps, 168);
    memcpy(alloc, m->allocVectors, (size_t)m->nbAllocVectors * 21);
    return m->cache.size;
}

/* V(n, k) of the reference's PVQ codebook, via its own encoder-side size function where exported; here through the decoder:
 * the number of codewords is what ec_dec_uint is called with in decode_pulses (cwrs.c), CELT_PVQ_V(n, k). */


/* per-frame final range through the PUBLIC API: every frame is wrapped as a code-0 packet with the given TOC */
int ref_opus_frames_final_range(const uint8_t *frames, const int32_t *lens, int n, int max_len, int channels, uint8_t toc, uint32_t *ranges, int frame_samples) {
    int err = 0;
    OpusDecoder *d = opus_decoder_create(48000, channels, &err);
    if (!d || err != OPUS_OK) return -1;
    int16_t *pcm = (int16_t *)malloc(sizeof(int16_t) * 5760 * 2);
    uint8_t *pkt = (uint8_t *)malloc((size_t)max_len + 1);
    for (int i = 0; i < n; ++i) {
        pkt[0] = (uint8_t)(toc & 0xFC); /* code 0: one frame */
        memcpy(pkt + 1, frames + (size_t)i * max_len, (size_t)lens[i]);
        int r = opus_decode(d, pkt, lens[i] + 1, pcm, 5760, 0);
        if (r < 0) { ranges[i] = 0xFFFFFFFFu; continue; }
        (void)frame_samples;
        opus_uint32 fr = 0;
        opus_decoder_ctl(d, OPUS_GET_FINAL_RANGE(&fr));
        ranges[i] = fr;
    }
    free(pkt);
    free(pcm);
    opus_decoder_destroy(d);
    return n;
}

/* static synthesis tables of the standard mode: window[120], mdct trig[1800], fft twiddles[480] (re, im), bitrev of the four transforms back to back */
int ref_celt_synth_tables(int16_t *window, int16_t *trig, int16_t *fft_tw, int16_t *bitrev) {
    int err = 0;
    const CELTMode *m = opus_custom_mode_create(48000, 960, &err);
    if (!m) return -1;
    memcpy(window, m->window, 120 * sizeof(int16_t));
    memcpy(trig, m->mdct.trig, 1800 * sizeof(int16_t));
    for (int i = 0; i < 480; i++) {
        fft_tw[2 * i] = m->mdct.kfft[0]->twiddles[i].r;
        fft_tw[2 * i + 1] = m->mdct.kfft[0]->twiddles[i].i;
    }
    for (int k = 0, off = 0; k < 4; k++) {
        memcpy(bitrev + off, m->mdct.kfft[k]->bitrev, (size_t)(480 >> k) * sizeof(int16_t));
        off += 480 >> k;
    }
    return m->mdct.n;
}
