"""CPU restatement of the reference transmitter's send-rate limiter -- TEST INFRASTRUCTURE (oracle):
imported by tests/ only, never by the product.

Follows transmitter/src/main/kotlin/com/github/tmarsteel/audionetwork/transmitter/LeakyBucket.kt:9-64
(Kotlin Long arithmetic: `/` truncates toward zero, coerceAtLeast(0)); System.nanoTime() is replaced by an
explicit `now` argument so that schedules are reproducible.  No golden vectors exist for it in the
reference's tests: parity unpinned, the restatement is line-by-line.
"""

NANOS_PER_SECOND = 1_000_000_000


def _kdiv(a, b):
    """Kotlin / Java long division (truncation toward zero)."""
    q = abs(a) // abs(b)
    return q if (a >= 0) == (b >= 0) else -q


class LeakyBucket:
    def __init__(self, capacity, drain_rate_per_second, now):
        self.capacity = capacity
        self.drain_rate_per_second = drain_rate_per_second
        self.last_value = 0              # LeakyBucket.kt:16
        self.last_value_at_nanos = now   # LeakyBucket.kt:19

    def current_value(self, now):        # LeakyBucket.kt:21-26
        nanos = now - self.last_value_at_nanos
        drained = _kdiv(self.drain_rate_per_second * nanos, NANOS_PER_SECOND)
        return max(self.last_value - drained, 0)

    def try_put(self, amount, now):      # LeakyBucket.kt:33-51: None = added, else nanoseconds to wait
        cur = self.current_value(now)
        if amount > self.capacity:
            raise ValueError("amount exceeds the bucket capacity")
        new_value = cur + amount
        if new_value > self.capacity:
            return _kdiv((new_value - self.capacity) * NANOS_PER_SECOND, self.drain_rate_per_second)
        self.last_value = cur + amount
        self.last_value_at_nanos = now
        return None

    def wait_for_capacity(self, amount, now):  # LeakyBucket.kt:57-64 on a virtual clock; returns the new time
        while True:
            d = self.try_put(amount, now)
            if d is None:
                return now
            now += max(d, 1)
