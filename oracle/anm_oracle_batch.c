/*
 * anm_oracle_batch.c -- runs the CPU oracle over many channels with one pthread per
 * core (the "reference C demodulator timed on the host cores, one channel per core"
 * leg of BASELINE.json).  TEST / BENCH INFRASTRUCTURE ONLY -- see anm_oracle.h.
 */
#define _GNU_SOURCE
#include "anm_oracle.h"

#include <pthread.h>
#include <sched.h>
#include <stdlib.h>
#include <time.h>
#include <unistd.h>

typedef struct {
    const anm_config_t *cfg;
    const float *tw;
    const int16_t *pcm;
    uint32_t n_ch, tid, n_threads;
    size_t ch_stride, n_samples;
    uint64_t ok, bad, bytes, digest;
} job_t;

static uint64_t fnv(uint64_t h, const void *p, size_t n) {
    const unsigned char *c = (const unsigned char *)p;
    for (size_t i = 0; i < n; ++i) h = (h ^ c[i]) * 0x100000001B3ull;
    return h;
}

static void *worker(void *arg) {
    job_t *j = (job_t *)arg;
    cpu_set_t set;
    CPU_ZERO(&set);
    CPU_SET(j->tid % (unsigned)sysconf(_SC_NPROCESSORS_ONLN), &set);
    pthread_setaffinity_np(pthread_self(), sizeof set, &set); /* best effort */
    anm_oracle_t *o = anm_oracle_create(j->cfg, j->tw);
    for (uint32_t c = j->tid; c < j->n_ch; c += j->n_threads) {
        anm_oracle_reset(o);
        anm_oracle_feed(o, j->pcm + (size_t)c * j->ch_stride, j->n_samples);
        size_t nf = anm_oracle_num_frames(o);
        const anm_frame_t *f = anm_oracle_frames(o);
        const uint8_t *by = anm_oracle_bytes(o);
        for (size_t i = 0; i < nf; ++i) {
            uint64_t d = fnv(0xCBF29CE484222325ull ^ c, &f[i].start_sample, 8);
            d = fnv(d, &f[i].len, 4);
            d = fnv(d, &f[i].crc_ok, 4);
            d = fnv(d, by + f[i].offset, f[i].len);
            j->digest += d; /* order-independent sum over frames */
            if (f[i].crc_ok) { j->ok++; j->bytes += f[i].len; } else j->bad++;
        }
    }
    anm_oracle_destroy(o);
    return NULL;
}

double anm_oracle_run_batch(const anm_config_t *cfg, const float *twiddles, const int16_t *pcm,
                            uint32_t n_ch, size_t ch_stride, size_t n_samples, uint32_t n_threads,
                            uint64_t *frames_ok, uint64_t *frames_bad, uint64_t *payload_bytes_ok,
                            uint64_t *digest) {
    if (n_threads == 0) n_threads = 1;
    pthread_t *th = (pthread_t *)calloc(n_threads, sizeof *th);
    job_t *jobs = (job_t *)calloc(n_threads, sizeof *jobs);
    struct timespec a, b;
    clock_gettime(CLOCK_MONOTONIC, &a);
    for (uint32_t t = 0; t < n_threads; ++t) {
        jobs[t] = (job_t){cfg, twiddles, pcm, n_ch, t, n_threads, ch_stride, n_samples, 0, 0, 0, 0};
        pthread_create(&th[t], NULL, worker, &jobs[t]);
    }
    uint64_t ok = 0, bad = 0, bytes = 0, dg = 0;
    for (uint32_t t = 0; t < n_threads; ++t) {
        pthread_join(th[t], NULL);
        ok += jobs[t].ok; bad += jobs[t].bad; bytes += jobs[t].bytes; dg += jobs[t].digest;
    }
    clock_gettime(CLOCK_MONOTONIC, &b);
    if (frames_ok) *frames_ok = ok;
    if (frames_bad) *frames_bad = bad;
    if (payload_bytes_ok) *payload_bytes_ok = bytes;
    if (digest) *digest = dg;
    free(th); free(jobs);
    return (double)(b.tv_sec - a.tv_sec) + 1e-9 * (double)(b.tv_nsec - a.tv_nsec);
}
