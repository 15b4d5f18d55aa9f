/*
 * ref_celt_state_shim.c -- the REFERENCE's CELT decoder run frame by frame with its private state made visible, linked into
 * oracle/_ref/libref_opus.so.  TEST INFRASTRUCTURE ONLY.
 *
 * CELTDecoder (struct OpusCustomDecoder) is private to celt/celt_decoder.c, and what stage 2 of the batched decoder needs from it --
 * the two log-energy histories anti_collapse() reads, the noise seed, the band energies -- lives behind that struct.  This file
 * therefore COMPILES celt/celt_decoder.c a second time, in place from the reference tree, as part of this translation unit (its
 * external names renamed so that they do not collide with the copy already in the library), and reads the state the reference's own
 * celt_decode_with_ec() leaves.  No line of the reference is restated here.
 */
#ifdef HAVE_CONFIG_H
#include "config.h"
#endif
#define validate_celt_decoder shim2_validate_celt_decoder
#define celt_decoder_get_size shim2_celt_decoder_get_size
#define opus_custom_decoder_get_size shim2_opus_custom_decoder_get_size
#define celt_decoder_init shim2_celt_decoder_init
#define opus_custom_decoder_init shim2_opus_custom_decoder_init
#define opus_custom_decoder_destroy shim2_opus_custom_decoder_destroy
#define celt_decode_with_ec shim2_celt_decode_with_ec
#define opus_custom_decoder_ctl shim2_opus_custom_decoder_ctl
#define deemphasis shim2_deemphasis
#define celt_synthesis shim2_celt_synthesis
#include "celt/celt_decoder.c"

#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
    uint32_t rng_before, rng_after;   /* CELTDecoder.rng: the noise seed the frame starts with / the final range it leaves */
    int16_t log_e1_before[42], log_e2_before[42]; /* oldLogE, oldLogE2 as anti_collapse() of this frame sees them */
    int16_t band_e_after[42], log_e1_after[42], log_e2_after[42];
    int32_t ret;                      /* celt_decode_with_ec's return value */
} ref_celt_state_t;

/* n frames of one stream (frame i: frames + i * max_len, lens[i] bytes; params[3 i ..]: its stream channels C, LM, end band -- set per frame the way
 * opus_decode_frame does, opus_decoder.c:462-489), decoded by a fresh decoder of CC output channels; pcm (may be NULL): [n][960][CC] int16 */
int ref_celt_stream_states(const uint8_t *frames, const int32_t *lens, int n, int max_len, const int32_t *params, int CC, ref_celt_state_t *out,
                           int16_t *pcm) {
    const int size = shim2_celt_decoder_get_size(CC);
    CELTDecoder *st = (CELTDecoder *)calloc(1, (size_t)size);
    int16_t *scratch = (int16_t *)malloc(sizeof(int16_t) * 960 * 2);
    if (!st || !scratch || shim2_celt_decoder_init(st, 48000, CC) != OPUS_OK) { free(st); free(scratch); return -1; }
    const int nb = st->mode->nbEBands;
    opus_val16 *lpc = (opus_val16 *)(st->_decode_mem + (DECODE_BUFFER_SIZE + st->overlap) * st->channels);
    opus_val16 *oldBandE = lpc + st->channels * LPC_ORDER, *oldLogE = oldBandE + 2 * nb, *oldLogE2 = oldLogE + 2 * nb;
    for (int i = 0; i < n; ++i) {
        ref_celt_state_t *o = &out[i];
        const int N = 120 << params[3 * i + 1];
        shim2_opus_custom_decoder_ctl(st, CELT_SET_END_BAND(params[3 * i + 2]));
        shim2_opus_custom_decoder_ctl(st, CELT_SET_CHANNELS(params[3 * i]));
        o->rng_before = st->rng;
        memcpy(o->log_e1_before, oldLogE, 42 * sizeof(int16_t));
        memcpy(o->log_e2_before, oldLogE2, 42 * sizeof(int16_t));
        o->ret = shim2_celt_decode_with_ec(st, frames + (size_t)i * max_len, lens[i], pcm ? pcm + (size_t)i * 960 * CC : scratch, N, NULL, 0);
        o->rng_after = st->rng;
        memcpy(o->band_e_after, oldBandE, 42 * sizeof(int16_t));
        memcpy(o->log_e1_after, oldLogE, 42 * sizeof(int16_t));
        memcpy(o->log_e2_after, oldLogE2, 42 * sizeof(int16_t));
    }
    free(scratch);
    free(st);
    return n;
}
