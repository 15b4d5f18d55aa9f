/*
 * anm_pacer.c -- send-rate limiter of the streaming feed, host C.
 *
 * The reference's only flow control: the transmitter models the receivers' buffer usage in
 * milliseconds of audio with a leaky bucket of capacity 1200 that drains 1000 per second
 * (transmitter/.../MulticastAudioOutput.kt:79-96) and waits for room before every frame
 * (LeakyBucket.kt:33-64); the receiver side of the same budget is the 40-frame queue of
 * hardware/src/playback.cpp:152.  A host that streams chunked PCM to anm_demod_feed_host in real time
 * paces its chunks with the same arithmetic.  The clock is a parameter (the reference reads
 * System.nanoTime()), so the behaviour is deterministic and testable; all arithmetic is the
 * reference's 64-bit integer arithmetic (truncating division, level clamped at zero).
 */
#include "../../include/anmodem.h"

#include <stdint.h>

#define NANOS_PER_SECOND 1000000000LL

int anm_pacer_init(anm_pacer_t *p, int64_t capacity, int64_t drain_rate_per_second, int64_t now_ns) {
    if (!p || capacity < 0 || capacity > INT64_MAX / 2 || drain_rate_per_second <= 0) return ANM_ERR_ARG; /* level + amount must fit */
    p->capacity = capacity;
    p->drain_rate_per_second = drain_rate_per_second;
    p->last_value = 0;
    p->last_value_at_ns = now_ns;
    return ANM_OK;
}

/* LeakyBucket.currentValue, LeakyBucket.kt:21-26 */
int64_t anm_pacer_level(const anm_pacer_t *p, int64_t now_ns) {
    const int64_t since = now_ns - p->last_value_at_ns;
    /* the reference multiplies in 64 bits (and silently wraps after long idle times at high rates); 128 bits here */
    const __int128 drained = (__int128)p->drain_rate_per_second * since / NANOS_PER_SECOND;
    const __int128 v = (__int128)p->last_value - drained;
    return v < 0 ? 0 : (v > INT64_MAX ? INT64_MAX : (int64_t)v);
}

/* LeakyBucket.tryPut, LeakyBucket.kt:33-51: 0 = added; > 0 = nanoseconds to wait before retrying (the
 * level is left untouched); ANM_ERR_ARG where the reference throws (amount > capacity) */
int64_t anm_pacer_try_put(anm_pacer_t *p, int64_t amount, int64_t now_ns) {
    if (!p || amount < 0 || amount > p->capacity) return ANM_ERR_ARG;
    const int64_t cur = anm_pacer_level(p, now_ns);
    const int64_t nv = cur + amount;
    if (nv > p->capacity) {
        const __int128 wait = (__int128)(nv - p->capacity) * NANOS_PER_SECOND / p->drain_rate_per_second;
        return wait > INT64_MAX ? INT64_MAX : (wait > 0 ? (int64_t)wait : 1); /* a sub-nanosecond overshoot still has to wait */
    }
    p->last_value = nv;
    p->last_value_at_ns = now_ns;
    return 0;
}

/* LeakyBucket.waitForCapacity (LeakyBucket.kt:57-64) on a virtual clock: puts `amount`, advancing *now_ns
 * by the delays the reference would sleep; returns the total nanoseconds waited or ANM_ERR_ARG */
int64_t anm_pacer_wait_for_capacity(anm_pacer_t *p, int64_t amount, int64_t *now_ns) {
    if (!p || !now_ns) return ANM_ERR_ARG;
    int64_t waited = 0;
    for (;;) {
        const int64_t d = anm_pacer_try_put(p, amount, *now_ns);
        if (d < 0) return d;
        if (d == 0) return waited;
        *now_ns += d;
        waited += d;
    }
}
