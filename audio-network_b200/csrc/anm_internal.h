/* anm_internal.h -- declarations shared by the host C and CUDA translation units. */
#ifndef ANM_INTERNAL_H_INCLUDED
#define ANM_INTERNAL_H_INCLUDED

#include "../../include/anmodem.h"

#ifdef __cplusplus
extern "C" {
#endif

uint32_t anm_bits_per_sym(const anm_config_t *c);
/* 1024-entry Q15 sine table of the transmitter DDS (SPEC.md section 6) */
const int16_t *anm_tx_sine_table(void);
/* noise scale (Q20) for an amplitude and SNR; 0 when clean */
uint32_t anm_tx_noise_scale(uint32_t amplitude_q15, int32_t snr_mdb);
/* Q32 tx samples per rx sample */
int64_t anm_tx_step(int32_t ppm_x1000);
void anm_set_error(const char *fmt, ...);

#ifdef __cplusplus
}
#endif
#endif
