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
/* twiddles: [sym_len][n_tones][2] table (SPEC 3), an input of the oracle */
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
#ifdef __cplusplus
}
#endif
#endif
