/*
 * anm_oracle.h -- CPU oracle of the SPEC.md receive path.  TEST INFRASTRUCTURE ONLY:
 * may be used by tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs,
 * never by the product.  PARITY UNPINNED w.r.t. the reference (see anm_oracle.c).
 * Uses the public types of include/anmodem.h (types only; no product code is linked).
 */
#ifndef ANM_ORACLE_H_INCLUDED
#define ANM_ORACLE_H_INCLUDED
#include "../include/anmodem.h"
#ifdef __cplusplus
extern "C" {
#endif
typedef struct anm_oracle anm_oracle_t;
/* twiddles: NULL = the oracle computes its own table from SPEC 3 (the normal case); a caller may pass a
 * [sym_len][n_tones][2] table instead (tests that perturb it) */
anm_oracle_t *anm_oracle_create(const anm_config_t *cfg, const float *twiddles);
void anm_oracle_reset(anm_oracle_t *o);
void anm_oracle_destroy(anm_oracle_t *o);
/* optional per-hop trace buffers (any may be NULL): E[cap][T], d[cap], emax[cap] */
void anm_oracle_set_trace(anm_oracle_t *o, float *E, uint8_t *d, float *emax, size_t cap_hops);
void anm_oracle_feed(anm_oracle_t *o, const int16_t *pcm, size_t n);
size_t anm_oracle_num_frames(const anm_oracle_t *o);
const anm_frame_t *anm_oracle_frames(const anm_oracle_t *o);
const uint8_t *anm_oracle_bytes(const anm_oracle_t *o);
size_t anm_oracle_num_bytes(const anm_oracle_t *o);
size_t anm_oracle_num_symbols(const anm_oracle_t *o);
const uint8_t *anm_oracle_symbols(const anm_oracle_t *o);
void anm_oracle_stats(const anm_oracle_t *o, anm_chan_stats_t *out);

/* Batch runner for the CPU baseline: channel c of pcm[c*ch_stride + i] is demodulated
 * by thread (c mod n_threads), one channel at a time per core.  Returns elapsed seconds
 * (wall clock around all threads); totals are written to the out pointers. */
double anm_oracle_run_batch(const anm_config_t *cfg, const float *twiddles, const int16_t *pcm,
                            uint32_t n_ch, size_t ch_stride, size_t n_samples, uint32_t n_threads,
                            uint64_t *frames_ok, uint64_t *frames_bad, uint64_t *payload_bytes_ok,
                            uint64_t *digest);

/* ---- the oracle's own transmit side (anm_oracle_tx.c): SPEC 2 presets, SPEC 4 frames, SPEC 6 transmitter.
 * Independent restatements; tests hold the product's anm_config_preset / anm_frame_symbols / anm_tx_render
 * against them, and bench.py --impl reference uses nothing else. */
int anm_oracle_preset(const char *name, anm_config_t *out);
size_t anm_oracle_frame_symbols(const anm_config_t *cfg, const uint8_t *payload, size_t len, uint8_t *syms, size_t cap);
int anm_oracle_tx_render(const anm_config_t *cfg, const uint8_t *program, size_t prog_len, const anm_tx_params_t *p,
                         uint64_t first_sample, int16_t *out, size_t n);
int anm_oracle_tx_render_batch(const anm_config_t *cfg, const uint8_t *programs, size_t prog_stride, const uint32_t *prog_len,
                               const anm_tx_params_t *params, uint32_t n_ch, uint64_t first_sample, int16_t *out,
                               size_t ch_stride, size_t n, uint32_t n_threads);
#ifdef __cplusplus
}
#endif
#endif
