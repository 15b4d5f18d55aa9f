/*
 * anm_host_queue.h -- host-side queue of decoded frames (C++ host code of the C-ABI layer).
 *
 * Storage is pinned (page-locked) memory that grows by doubling, so the device-to-host copies of a drain land
 * in the queue itself at full PCIe speed and nothing is copied twice.  Frames are kept in arrival order
 * (the order of the kernels' atomic counters); the (channel, start_sample) order the C ABI promises is built
 * on demand and only for what arrived since the last read:
 *   - one warp owns one channel for a whole launch and launches on a handle are ordered, so the frames of a
 *     channel arrive in increasing start_sample order; a STABLE counting sort by channel is therefore the full
 *     (channel, start_sample) order, in O(frames + channels).  The result is verified in the same pass and the
 *     code falls back to a comparison sort should the assumption ever not hold;
 *   - reads pop from a cursor; nothing is re-sorted or compacted per call, so popping frames one at a time
 *     (demod_read_frames, demod_as_pb_istream) is linear in the number of frames.
 */
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "../../include/anmodem.h"

namespace anm {

template <typename T>
struct PinnedVec {
    T *p = nullptr;
    size_t n = 0, cap = 0;
    bool pinned = false;
    ~PinnedVec() { release(); }
    void release() {
        if (p) {
            if (pinned) cudaFreeHost(p);
            else free(p);
        }
        p = nullptr;
        n = cap = 0;
    }
    /* room for `extra` more elements; false when out of memory */
    bool reserve_extra(size_t extra) {
        if (n + extra <= cap) return true;
        size_t ncap = std::max<size_t>(cap * 2, n + extra);
        ncap = std::max<size_t>(ncap, 4096 / sizeof(T) + 1);
        T *np = nullptr;
        bool pin = true;
        if (cudaHostAlloc((void **)&np, ncap * sizeof(T), cudaHostAllocDefault) != cudaSuccess) {
            cudaGetLastError();
            np = (T *)malloc(ncap * sizeof(T)); /* pageable memory still works, just slower */
            pin = false;
            if (!np) return false;
        }
        if (n) memcpy(np, p, n * sizeof(T));
        const size_t keep = n;
        release();
        p = np;
        n = keep;
        cap = ncap;
        pinned = pin;
        return true;
    }
};

struct FrameQueue {
    PinnedVec<anm_frame_t> frames; /* arrival order; .offset indexes `bytes` */
    PinnedVec<uint8_t> bytes;
    std::vector<uint32_t> order;   /* indices into frames in (channel, start_sample) order, from `cursor` on still unread */
    size_t cursor = 0;
    size_t n_sorted = 0;           /* frames[0 .. n_sorted) are covered by `order` */
    std::vector<uint32_t> hist;    /* counting-sort scratch */

    size_t pending() const { return (order.size() - cursor) + (frames.n - n_sorted); }
    void clear() {
        frames.n = bytes.n = 0;
        order.clear();
        cursor = n_sorted = 0;
    }

    /* appends nf records / nb bytes that the caller fills in place (D2H target); returns false when out of memory */
    bool grow(size_t nf, size_t nb, anm_frame_t **fdst, uint8_t **bdst) {
        if (pending() == 0) clear(); /* everything was read: start over at the front */
        if (!frames.reserve_extra(nf) || !bytes.reserve_extra(nb)) return false;
        *fdst = frames.p + frames.n;
        *bdst = bytes.p + bytes.n;
        frames.n += nf;
        bytes.n += nb;
        return true;
    }

    /* brings `order` up to date: unread old entries first, then the new arrivals, stably sorted by channel */
    void sort_pending(uint32_t n_ch) {
        if (n_sorted == frames.n) return;
        std::vector<uint32_t> src;
        src.reserve(order.size() - cursor + (frames.n - n_sorted));
        src.insert(src.end(), order.begin() + cursor, order.end());
        for (size_t i = n_sorted; i < frames.n; ++i) src.push_back((uint32_t)i);
        order.resize(src.size());
        cursor = 0;
        n_sorted = frames.n;
        bool counting_ok = true;
        hist.assign((size_t)n_ch + 1, 0u);
        for (uint32_t i : src) {
            const uint32_t c = frames.p[i].channel;
            if (c >= n_ch) { counting_ok = false; break; }
            ++hist[c + 1];
        }
        if (counting_ok) {
            for (size_t c = 0; c < n_ch; ++c) hist[c + 1] += hist[c];
            for (uint32_t i : src) order[hist[frames.p[i].channel]++] = i;
            for (size_t k = 1; k < order.size(); ++k) {
                const anm_frame_t &x = frames.p[order[k - 1]], &y = frames.p[order[k]];
                if (x.channel == y.channel && x.start_sample > y.start_sample) { counting_ok = false; break; }
            }
        }
        if (!counting_ok) {
            order = src;
            std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) {
                const anm_frame_t &x = frames.p[a], &y = frames.p[b];
                return x.channel != y.channel ? x.channel < y.channel : x.start_sample < y.start_sample;
            });
        }
    }

    /* pops up to cap frames in (channel, start_sample) order */
    size_t pop_sorted(uint32_t n_ch, anm_frame_t *out, size_t cap, uint8_t *obytes, size_t obytes_cap) {
        sort_pending(n_ch);
        size_t n = 0, bo = 0;
        while (n < cap && cursor < order.size()) {
            const anm_frame_t &f = frames.p[order[cursor]];
            if (bo + f.len > obytes_cap) break;
            out[n] = f;
            out[n].offset = (uint32_t)bo;
            if (obytes && f.len) memcpy(obytes + bo, bytes.p + f.offset, f.len);
            bo += f.len;
            ++cursor;
            ++n;
        }
        if (pending() == 0) clear();
        return n;
    }

    /* moves out EVERYTHING that is queued, in arrival order, as two plain copies (per channel the order is still
     * chronological); returns 0 and leaves the queue untouched when the destination is too small or a sorted read is
     * half way through its batch */
    size_t take_all(anm_frame_t *out, size_t cap, uint8_t *obytes, size_t obytes_cap, size_t *nbytes) {
        if (cursor != 0 || frames.n > cap || bytes.n > obytes_cap) return 0;
        const size_t n = frames.n, nb = bytes.n;
        if (n) memcpy(out, frames.p, n * sizeof(anm_frame_t));
        if (nb && obytes) memcpy(obytes, bytes.p, nb);
        if (nbytes) *nbytes = nb;
        clear();
        return n;
    }
};

} /* namespace anm */
