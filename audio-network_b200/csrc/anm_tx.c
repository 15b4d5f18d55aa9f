/*
 * anm_tx.c -- CPU transmitter stand-in: renders the int16 PCM a receiver channel
 * would capture (SPEC.md section 6).  Integer-only so that the CUDA renderer in
 * anm_tx.cu produces the same samples bit for bit.
 *
 * Reference context: BASELINE.json configs[0] names a "Java transmitter" that
 * renders ip.proto frames to PCM; SURVEY.md section 0 shows the real transmitter
 * (transmitter/src/main/kotlin/.../MulticastAudioOutput.kt:72-96) sends Opus over
 * TCP and never produces modem PCM, so this generator is authored here.
 */
#include "anm_internal.h"
#define ANM_STR_(x) #x
#define ANM_STR(x) ANM_STR_(x)

#include <math.h>
#include <pthread.h>
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

static int16_t g_sine[1024];
static pthread_once_t g_sine_once = PTHREAD_ONCE_INIT;

static void init_sine(void) {
    const double two_pi = 6.283185307179586476925286766559;
    for (int i = 0; i < 1024; ++i) {
        double v = 32767.0 * sin(two_pi * (double)i / 1024.0);
        g_sine[i] = (int16_t)(v >= 0 ? floor(v + 0.5) : -floor(-v + 0.5));
    }
}

const int16_t *anm_tx_sine_table(void) {
    pthread_once(&g_sine_once, init_sine);
    return g_sine;
}

uint32_t anm_tx_noise_scale(uint32_t amplitude_q15, int32_t snr_mdb) {
    if (snr_mdb == ANM_SNR_CLEAN) return 0;
    double sig_rms = (double)amplitude_q15 / sqrt(2.0);
    double sigma = sig_rms / pow(10.0, (double)snr_mdb / 20000.0);
    double v = sigma / 53509.0 * 1048576.0;
    if (v > 4294967295.0) v = 4294967295.0;
    return (uint32_t)floor(v + 0.5);
}

int64_t anm_tx_step(int32_t ppm_x1000) {
    /* round_half_away(ppm_x1000 * 2^32 / 1e9), exact integer arithmetic (same on the GPU) */
    const int64_t num = (int64_t)ppm_x1000 * 4294967296ll;
    const int64_t r = (num >= 0) ? (num + 500000000ll) / 1000000000ll : -((-num + 500000000ll) / 1000000000ll);
    return ((int64_t)1 << 32) + r;
}

static inline uint64_t mix64(uint64_t z) {
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

int anm_tx_render(const anm_config_t *cfg, const uint8_t *program, size_t prog_len,
                  const anm_tx_params_t *p, uint64_t first_sample, int16_t *out, size_t n) {
    if (anm_config_validate(cfg) != ANM_OK || !program || !prog_len || !p || (!out && n)) return ANM_ERR_ARG;
    const int16_t *sine = anm_tx_sine_table();
    const uint32_t N = cfg->sym_len;
    uint32_t lg = 0;
    while ((1u << lg) < N) ++lg;
    const int64_t step = anm_tx_step(p->ppm_x1000);
    const uint32_t nscale = anm_tx_noise_scale(p->amplitude_q15, p->snr_mdb);
    for (size_t i = 0; i < n; ++i) {
        uint64_t nn = first_sample + i;
        __int128 tpos = ((__int128)p->start_offset << 32) + (__int128)nn * step;
        int32_t sig = 0;
        if (tpos >= 0) {
            uint64_t t = (uint64_t)(tpos >> 32);      /* whole tx samples */
            uint64_t sym = t >> lg;
            uint8_t e = program[sym % prog_len];
            if (e != ANM_SILENCE && e < cfg->n_tones) {
                uint64_t pos = (uint64_t)(tpos - ((__int128)(sym << lg) << 32)); /* Q32, < N*2^32 */
                uint32_t phase = (uint32_t)(((__int128)cfg->tone_bin[e] * pos) >> lg);
                sig = ((int32_t)p->amplitude_q15 * (int32_t)sine[phase >> 22]) >> 15;
            }
        }
        int32_t noise = 0;
        if (nscale) {
            uint64_t a = mix64(p->seed + (2 * nn) * 0x9E3779B97F4A7C15ull);
            uint64_t b = mix64(p->seed + (2 * nn + 1) * 0x9E3779B97F4A7C15ull);
            int64_t s = (int64_t)((a & 0xFFFF) + ((a >> 16) & 0xFFFF) + ((a >> 32) & 0xFFFF) + (a >> 48) +
                                  (b & 0xFFFF) + ((b >> 16) & 0xFFFF) + ((b >> 32) & 0xFFFF) + (b >> 48)) - 262140;
            noise = (int32_t)((s * (int64_t)nscale) >> 20);
        }
        int32_t v = sig + noise;
        out[i] = (int16_t)(v > 32767 ? 32767 : (v < -32768 ? -32768 : v));
    }
    return ANM_OK;
}

/* ---- error string ---------------------------------------------------------- */
static __thread char g_err[256];
void anm_set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
}
const char *anm_last_error(void) { return g_err; }
const char *anm_version(void) { return "anmodem-b200 0.2 (SPEC.md rev " ANM_STR(ANM_SPEC_REVISION) ")"; }
