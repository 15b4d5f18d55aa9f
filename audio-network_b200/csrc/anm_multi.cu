/*
 * anm_multi.cu -- several GPUs behind one handle (include/anmodem.h, anm_demod_multi_*): SURVEY.md 8(e).
 *
 * Channels are independent, so they shard trivially: device d owns a contiguous range of channel ids, has its own
 * anm_demod_t (state, streams, staging buffers, result rings) and ONE host thread that issues every CUDA call for it --
 * copy, kernel, drain -- so the devices' host-side work (enqueueing, draining, offset fix-ups) runs in parallel and each
 * thread can be bound to the CPUs next to its GPU.  There is no collective and no peer traffic on the data path: PCM goes
 * host -> its device, frames come back device -> host, and the only "exchange" is the host-side gather of frame records
 * into one queue, ordered by (global channel, start_sample) like the single-device interface.
 *
 * Reference idiom mirrored: one producer task per resource feeding a queue that one consumer drains
 * (hardware/src/network.cpp:544-582, playback.cpp:174-191).
 */
#include <cuda_runtime.h>
#include <sched.h>
#include <sys/mman.h>
#include <unistd.h>

#include <condition_variable>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <new>
#include <thread>
#include <vector>

#include "anm_host_queue.h"
#include "anm_internal.h"

namespace {

enum Cmd : int { CMD_NONE = 0, CMD_FEED, CMD_COLLECT, CMD_COLLECT_UPTO, CMD_RESET, CMD_WAIT_INPUT, CMD_TOUCH, CMD_QUIT };

struct Worker {
    int device = 0;
    uint32_t first = 0, count = 0; /* channel range */
    anm_demod_t *h = nullptr;
    std::thread th;
    std::mutex mu;
    std::condition_variable cv;
    int cmd = CMD_NONE;   /* posted command */
    bool done = true;     /* the posted command has been executed */
    long rc = 0;
    /* arguments */
    const int16_t *pcm = nullptr;
    size_t ch_stride = 0, n_samples = 0;
    uint32_t lag = 0;
    void *touch_ptr = nullptr;
    size_t touch_bytes = 0;
    int numa_node = -1;
    bool bound = false;
};

/* CPUs of the NUMA node the device hangs off, from sysfs; best effort */
bool bind_thread_to_device_node(int device, int *node_out) {
    char bus[32] = {0};
    if (cudaDeviceGetPCIBusId(bus, sizeof bus, device) != cudaSuccess) { cudaGetLastError(); return false; }
    for (char *c = bus; *c; ++c) *c = (char)tolower(*c);
    char path[128];
    snprintf(path, sizeof path, "/sys/bus/pci/devices/%s/numa_node", bus);
    FILE *f = fopen(path, "r");
    int node = -1;
    if (f) {
        if (fscanf(f, "%d", &node) != 1) node = -1;
        fclose(f);
    }
    if (node_out) *node_out = node;
    if (node < 0) return false;
    snprintf(path, sizeof path, "/sys/devices/system/node/node%d/cpulist", node);
    f = fopen(path, "r");
    if (!f) return false;
    cpu_set_t set;
    CPU_ZERO(&set);
    int a, b, n = 0;
    while (fscanf(f, "%d", &a) == 1) {
        b = a;
        int c = fgetc(f);
        if (c == '-') {
            if (fscanf(f, "%d", &b) != 1) b = a;
            c = fgetc(f);
        }
        for (int i = a; i <= b && i < CPU_SETSIZE; ++i) { CPU_SET(i, &set); ++n; }
        if (c != ',') break;
    }
    fclose(f);
    return n > 0 && sched_setaffinity(0, sizeof set, &set) == 0;
}

void worker_main(Worker *w) {
    cudaSetDevice(w->device);
    w->bound = bind_thread_to_device_node(w->device, &w->numa_node);
    for (;;) {
        std::unique_lock<std::mutex> lk(w->mu);
        w->cv.wait(lk, [&] { return !w->done; });
        const int cmd = w->cmd;
        lk.unlock();
        long rc = 0;
        switch (cmd) {
        case CMD_FEED:
            rc = anm_demod_feed_host_async(w->h, w->pcm + (size_t)w->first * w->ch_stride, w->ch_stride, w->n_samples);
            break;
        case CMD_COLLECT: rc = anm_demod_collect(w->h); break;
        case CMD_COLLECT_UPTO: rc = anm_demod_collect_upto(w->h, w->lag); break;
        case CMD_RESET: rc = anm_demod_reset(w->h); break;
        case CMD_WAIT_INPUT: rc = anm_demod_wait_input(w->h); break;
        case CMD_TOUCH: memset(w->touch_ptr, 0, w->touch_bytes); break; /* first touch from the thread next to the GPU */
        default: break;
        }
        lk.lock();
        w->rc = rc;
        w->done = true;
        lk.unlock();
        w->cv.notify_all();
        if (cmd == CMD_QUIT) return;
    }
}

void post(Worker &w, int cmd) {
    std::lock_guard<std::mutex> lk(w.mu);
    w.cmd = cmd;
    w.done = false;
    w.cv.notify_all();
}
long wait_done(Worker &w) {
    std::unique_lock<std::mutex> lk(w.mu);
    w.cv.wait(lk, [&] { return w.done; });
    return w.rc;
}

} /* namespace */

struct anm_demod_multi {
    anm_config_t cfg;
    uint32_t n_ch = 0;
    std::vector<Worker *> workers;
    anm::FrameQueue q; /* gathered frames, global channel ids */
    int overflow = 0;
    struct Region { void *p; size_t bytes; };
    std::vector<Region> regions; /* PCM buffers handed out by anm_demod_multi_alloc_pcm */
};

/* runs one command on every device at once; the first negative result wins */
static long run_all(anm_demod_multi *m, int cmd) {
    for (Worker *w : m->workers) post(*w, cmd);
    long rc = 0, sum = 0;
    for (Worker *w : m->workers) {
        const long r = wait_done(*w);
        if (r < 0 && rc == 0) rc = r;
        else sum += r;
    }
    return rc < 0 ? rc : sum;
}

extern "C" void anm_demod_multi_destroy(anm_demod_multi_t *m) {
    if (!m) return;
    for (Worker *w : m->workers) {
        if (w->th.joinable()) {
            post(*w, CMD_QUIT);
            w->th.join();
        }
        anm_demod_destroy(w->h);
        delete w;
    }
    for (auto &r : m->regions) {
        cudaHostUnregister(r.p);
        munmap(r.p, r.bytes);
    }
    m->q.frames.release();
    m->q.bytes.release();
    delete m;
}

extern "C" int anm_demod_multi_create(const anm_config_t *cfg, uint32_t n_channels, const int *devices, uint32_t n_devices, uint32_t flags,
                                      anm_demod_multi_t **out) {
    if (!cfg || !out || !devices || n_devices == 0 || n_channels < n_devices) return ANM_ERR_ARG;
    if (flags & ANM_FLAG_SYMBOLS) { anm_set_error("the multi-device handle gathers frames only"); return ANM_ERR_UNSUPPORTED; }
    anm_demod_multi *m = new (std::nothrow) anm_demod_multi();
    if (!m) return ANM_ERR_NOMEM;
    m->cfg = *cfg;
    m->n_ch = n_channels;
    /* contiguous ranges, sizes differing by at most one */
    uint32_t first = 0;
    for (uint32_t d = 0; d < n_devices; ++d) {
        Worker *w = new (std::nothrow) Worker();
        if (!w) { anm_demod_multi_destroy(m); return ANM_ERR_NOMEM; }
        w->device = devices[d];
        w->first = first;
        w->count = n_channels / n_devices + (d < n_channels % n_devices ? 1u : 0u);
        first += w->count;
        m->workers.push_back(w);
        const int rc = anm_demod_create(cfg, w->count, w->device, flags, &w->h);
        if (rc != ANM_OK) { anm_demod_multi_destroy(m); return rc; }
    }
    for (Worker *w : m->workers) w->th = std::thread(worker_main, w);
    *out = m;
    return ANM_OK;
}

extern "C" uint32_t anm_demod_multi_num_devices(const anm_demod_multi_t *m) { return m ? (uint32_t)m->workers.size() : 0u; }

extern "C" int anm_demod_multi_shard(const anm_demod_multi_t *m, uint32_t d, int *device, uint32_t *first_channel, uint32_t *n_channels,
                                     int *numa_node) {
    if (!m || d >= m->workers.size()) return ANM_ERR_ARG;
    const Worker *w = m->workers[d];
    if (device) *device = w->device;
    if (first_channel) *first_channel = w->first;
    if (n_channels) *n_channels = w->count;
    if (numa_node) *numa_node = w->bound ? w->numa_node : -1;
    return ANM_OK;
}

extern "C" anm_demod_t *anm_demod_multi_device_handle(anm_demod_multi_t *m, uint32_t d) {
    return (m && d < m->workers.size()) ? m->workers[d]->h : nullptr;
}

extern "C" int anm_demod_multi_reset(anm_demod_multi_t *m) {
    if (!m) return ANM_ERR_ARG;
    const long rc = run_all(m, CMD_RESET);
    m->q.clear();
    m->overflow = 0;
    return rc < 0 ? (int)rc : ANM_OK;
}

extern "C" int anm_demod_multi_feed_host(anm_demod_multi_t *m, const int16_t *h_pcm, size_t ch_stride, size_t n_samples) {
    if (!m || (!h_pcm && n_samples)) return ANM_ERR_ARG;
    for (Worker *w : m->workers) {
        w->pcm = h_pcm;
        w->ch_stride = ch_stride;
        w->n_samples = n_samples;
    }
    const long rc = run_all(m, CMD_FEED);
    return rc < 0 ? (int)rc : ANM_OK;
}

extern "C" int anm_demod_multi_wait_input(anm_demod_multi_t *m) {
    if (!m) return ANM_ERR_ARG;
    const long rc = run_all(m, CMD_WAIT_INPUT);
    return rc < 0 ? (int)rc : ANM_OK;
}

/* host-side gather: what every device's handle has queued moves into the common queue, device after device (= ascending
 * channel ranges), with global channel ids */
static long gather(anm_demod_multi *m) {
    for (Worker *w : m->workers) {
        const anm_frame_t *f;
        const uint8_t *by;
        size_t nb = 0;
        const size_t n = anm_demod_peek_frames(w->h, &f, &by, &nb);
        if (anm_demod_overflowed(w->h)) m->overflow = 1;
        if (!n) continue;
        anm_frame_t *fdst;
        uint8_t *bdst;
        if (!m->q.grow(n, nb, &fdst, &bdst)) return ANM_ERR_NOMEM;
        const uint32_t b0 = (uint32_t)(m->q.bytes.n - nb);
        memcpy(bdst, by, nb);
        for (size_t i = 0; i < n; ++i) {
            fdst[i] = f[i];
            fdst[i].channel += w->first;
            fdst[i].offset += b0;
        }
        anm_demod_drop_frames(w->h);
    }
    return (long)m->q.pending();
}

extern "C" long anm_demod_multi_collect_upto(anm_demod_multi_t *m, uint32_t lag) {
    if (!m) return ANM_ERR_ARG;
    for (Worker *w : m->workers) w->lag = lag;
    const long rc = run_all(m, CMD_COLLECT_UPTO);
    if (rc < 0) return rc;
    return gather(m);
}

extern "C" long anm_demod_multi_collect(anm_demod_multi_t *m) {
    if (!m) return ANM_ERR_ARG;
    const long rc = run_all(m, CMD_COLLECT);
    if (rc < 0) return rc;
    return gather(m);
}

extern "C" size_t anm_demod_multi_read_frames(anm_demod_multi_t *m, anm_frame_t *out, size_t cap, uint8_t *bytes, size_t bytes_cap) {
    if (!m || !out || !cap) return 0;
    return m->q.pop_sorted(m->n_ch, out, cap, bytes, bytes_cap);
}

extern "C" size_t anm_demod_multi_take_frames(anm_demod_multi_t *m, anm_frame_t *out, size_t cap, uint8_t *bytes, size_t bytes_cap, size_t *n_bytes) {
    if (!m || !out) return 0;
    return m->q.take_all(out, cap, bytes, bytes_cap, n_bytes);
}

extern "C" int anm_demod_multi_overflowed(const anm_demod_multi_t *m) {
    if (!m) return 0;
    int o = m->overflow;
    for (const Worker *w : m->workers) o |= anm_demod_overflowed(w->h);
    return o;
}

extern "C" int anm_demod_multi_stats(anm_demod_multi_t *m, anm_chan_stats_t *out) {
    if (!m || !out) return ANM_ERR_ARG;
    for (Worker *w : m->workers) {
        const int rc = anm_demod_stats(w->h, out + w->first);
        if (rc) return rc;
    }
    return ANM_OK;
}

/* Page-locked PCM buffer for all channels, [channel][n_samples] with the returned stride.  Every shard's pages are touched
 * first by the thread that is bound next to the shard's GPU (so they are allocated on that NUMA node), then the whole
 * range is registered with CUDA for full-speed asynchronous copies. */
extern "C" int anm_demod_multi_alloc_pcm(anm_demod_multi_t *m, size_t n_samples, int16_t **out, size_t *ch_stride) {
    if (!m || !out || !n_samples) return ANM_ERR_ARG;
    const size_t stride = (n_samples + 7u) & ~(size_t)7u; /* 16-byte rows */
    const size_t page = (size_t)sysconf(_SC_PAGESIZE);
    const size_t bytes = ((size_t)m->n_ch * stride * sizeof(int16_t) + page - 1) / page * page;
    void *p = mmap(nullptr, bytes, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
    if (p == MAP_FAILED) { anm_set_error("mmap of %zu bytes failed", bytes); return ANM_ERR_NOMEM; }
    for (Worker *w : m->workers) {
        w->touch_ptr = (char *)p + (size_t)w->first * stride * sizeof(int16_t);
        w->touch_bytes = (size_t)w->count * stride * sizeof(int16_t);
    }
    run_all(m, CMD_TOUCH);
    if (cudaHostRegister(p, bytes, cudaHostRegisterPortable) != cudaSuccess) {
        anm_set_error("cudaHostRegister failed: %s", cudaGetErrorString(cudaGetLastError()));
        munmap(p, bytes);
        return ANM_ERR_CUDA;
    }
    m->regions.push_back({p, bytes});
    *out = (int16_t *)p;
    if (ch_stride) *ch_stride = stride;
    return ANM_OK;
}
