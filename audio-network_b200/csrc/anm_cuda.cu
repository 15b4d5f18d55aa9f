/*
 * anm_cuda.cu -- C-ABI host side of the batched demodulator (include/anmodem.h):
 * handle management, kernel dispatch, result collection.  Host code is C-style C++
 * calling the sm_100a kernels of anm_kernels.cuh; there is no CPU compute fallback.
 *
 * Reference seam: this is the byte source that would stand where the socket-backed
 * pb_istream_t stands today (hardware/src/network.cpp:262-305, consumed at :406-411);
 * the single-channel demod_* functions follow the reference's module idiom
 * (hardware/README.md:10-14, hardware/include/playback.hpp:15).
 */
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <mutex>
#include <new>
#include <vector>

#include "anm_host_queue.h"
#include "anm_kernels_tc.cuh"
#include "../../include/anmodem_pb.h"

using namespace anm;

#define CK(call)                                                                          \
    do {                                                                                  \
        cudaError_t e_ = (call);                                                          \
        if (e_ != cudaSuccess) {                                                          \
            anm_set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
            return ANM_ERR_CUDA;                                                          \
        }                                                                                 \
    } while (0)

namespace {

typedef void (*kern_fn)(const KParams);

struct Variant {
    uint32_t T, N, S;
    kern_fn fn;       /* streaming demodulator */
    kern_fn fn_trace; /* stateless tone-energy pass */
    kern_fn fn_fold, fn_trace_fold; /* the same for foldable configurations (SPEC 3, centre-folded hop partials) */
    uint32_t warp_smem, cta_smem, state_bytes;
    bool dense; /* SPEC 3b: tensor-core contraction kernel, 4 channels per CTA */
};

#define VARIANT(T_, N_, S_)                                                                       \
    {T_, N_, S_, (kern_fn)k_demod<T_, N_, S_, 0, 0>, (kern_fn)k_demod<T_, N_, S_, 1, 0>,                     \
     (kern_fn)k_demod<T_, N_, S_, 0, 1>, (kern_fn)k_demod<T_, N_, S_, 1, 1>, warp_smem_bytes<T_, N_, S_>(), \
     cta_smem_bytes<T_, N_, S_>(), state_bytes<T_, S_>(), false}
#define VARIANT_TC(T_, N_, S_)                                                                    \
    {T_, N_, S_, (kern_fn)k_demod_tc<T_, N_, S_, 0>, (kern_fn)k_demod_tc<T_, N_, S_, 1>, nullptr, nullptr, 0u, tc::smem_bytes<T_, N_, S_>(), \
     state_bytes<T_, S_>(), true}

const Variant kVariants[] = {
    VARIANT(4, 128, 4),  VARIANT(2, 128, 4),  VARIANT(8, 128, 4), VARIANT(16, 128, 4),
    VARIANT_TC(64, 256, 4), VARIANT(4, 128, 8),  VARIANT(4, 128, 2), VARIANT(4, 64, 4),
};

/* the kernel of a configuration: the folded instantiation when SPEC 3's centre folding applies */
kern_fn variant_fn(const Variant *v, const anm_config_t *c, bool trace) {
    const bool fold = !v->dense && anm_config_foldable(c);
    return trace ? (fold ? v->fn_trace_fold : v->fn_trace) : (fold ? v->fn_fold : v->fn);
}

const Variant *find_variant(const anm_config_t *c) {
    for (const Variant &v : kVariants)
        if (v.T == c->n_tones && v.N == c->sym_len && v.S == c->hops_per_sym && v.dense == (anm_config_dense(c) != 0)) return &v;
    return nullptr;
}

struct EvPair {
    cudaEvent_t a, b;
};

constexpr size_t kSnapSlots = 64; /* per-launch snapshots of the queue counters kept in pinned memory */

uint32_t pow2_ceil(uint64_t v) {
    uint64_t p = 1;
    while (p < v) p <<= 1;
    return (uint32_t)std::min<uint64_t>(p, 1ull << 31);
}

} /* namespace */

struct anm_demod {
    anm_config_t cfg;
    const Variant *var;
    int device;
    uint32_t n_ch, flags;
    int num_sms;
    uint32_t warps_per_cta, grid;
    size_t smem_bytes;
    unsigned char *d_state;
    uint8_t *d_fsyms;
    uint32_t fsym_stride, max_frame_syms;
    anm_frame_t *d_frames;
    uint8_t *d_bytes;
    uint32_t *d_counters;
    uint32_t *d_progress;  /* [n_ch] chunks completed per channel: orders a channel's chunks inside a multi-chunk launch */
    uint64_t tickets;      /* work-queue tickets handed out since reset (counters[4]) */
    uint32_t chunks_done;  /* progress[] of every channel between launches */
    uint32_t frames_cap, bytes_cap;
    uint8_t *d_osyms;
    uint32_t osym_cap;
    float2 *d_tw;
    std::vector<float> h_tw;
    uint8_t *d_basis; /* dense tone sets: int8 basis panels */
    /* host feeds: two staging buffers, so that the copy of chunk k+1 crosses PCIe while the kernel of chunk k runs */
    int16_t *d_stage[2];
    size_t stage_cap;
    uint32_t stage_next;
    cudaStream_t own_stream, h2d_stream, last_stream;
    cudaEvent_t stage_free[2]; /* recorded behind the kernel that read the staging buffer */
    cudaEvent_t h2d_done, order_ev;
    uint64_t samples_fed, syms_since_collect;
    uint64_t launches;             /* launches of the streaming kernel since create (bench.py's gpu_launches) */
    uint64_t launches_since_reset; /* the work-queue / snapshot sequence restarts with the state */
    uint64_t drained_upto;         /* launches (since reset) whose frames are in the host queue */
    std::vector<EvPair> evs;
    size_t ev_used;
    uint32_t read_f, read_b;       /* frames / payload bytes consumed by the host (mod 2^32) */
    uint32_t seen_drops;           /* value of the device's drop counter the host has accounted for */
    uint32_t *snap;                /* pinned, device-visible [kSnapSlots][4]: written by each launch's last warp */
    cudaEvent_t snap_ev[kSnapSlots];
    cudaStream_t d2h_stream;
    anm::FrameQueue q;             /* host-side result queue (pinned storage; device-to-host copies land in it) */
    std::vector<std::deque<uint8_t>> q_syms;
    int overflow;
    KParams kp;
};

/* two bits per tone: tone_bin mod 4, the rotation (in quarter turns) a twiddle picks up when the
 * sample position advances by N/4 (anm_twiddles builds the table with exactly this symmetry) */
static void set_tw_sign(const anm_config_t *cfg, KParams *k) {
    k->tw_rot[0] = k->tw_rot[1] = 0;
    k->fold = anm_config_foldable(cfg) ? 1u : 0u;
    k->fold_odd = 0;
    if (k->fold)
        for (uint32_t t = 0; t < cfg->n_tones; ++t)
            if ((2u * cfg->tone_bin[t] / cfg->hops_per_sym) & 1u) k->fold_odd |= 1ull << t;
    memset(k->fold_tw, 0, sizeof k->fold_tw);
    if (k->fold && (cfg->sym_len / cfg->hops_per_sym / 2) * cfg->n_tones <= 256u) {
        std::vector<float> ft((size_t)cfg->sym_len * cfg->n_tones * 2);
        anm_fold_twiddles(cfg, ft.data());
        memcpy(k->fold_tw, ft.data(), (size_t)(cfg->sym_len / cfg->hops_per_sym / 2) * cfg->n_tones * 8u);
    }
    for (uint32_t t = 0; t < 16; ++t) {
        const float sg = ((k->fold_odd >> t) & 1ull) ? -1.0f : 1.0f;
        k->fold_sg[t] = make_float2(sg, sg);
    }
    uint32_t any = 0;
    for (uint32_t t = 0; t < cfg->n_tones; ++t) {
        k->tw_rot[t >> 5] |= (unsigned long long)(cfg->tone_bin[t] & 3u) << (2 * (t & 31));
        any |= cfg->tone_bin[t] & 3u;
    }
    k->rot_mode = (any == 0) ? 0u : ((any & 1u) ? 2u : 1u);
    /* CRC-16 lane constants of the lane-parallel frame check: x^(8(31-lane)+16) mod 0x11021 */
    for (int lane = 0; lane < 32; ++lane) {
        uint32_t v = 1;
        for (int i = 0; i < 8 * (31 - lane) + 16; ++i) v = ((v << 1) ^ ((v & 0x8000u) ? 0x1021u : 0u)) & 0xffffu;
        k->crc_pow[lane] = (uint16_t)v;
    }
    for (uint32_t b = 0; b < 256; ++b) {
        const uint8_t one = (uint8_t)b;
        k->crc8_tab[b] = anm_crc8(&one, 1, 0);
    }
}

/* int8 basis of a dense configuration in the panel order the MMA descriptors of k_demod_tc address:
 * [hop phase q][16-sample K chunk][column][sample in chunk]; columns go by tone pairs, 4 * (tone / 2) + {0: cos of the even tone,
 * 1: cos of the odd tone, 2: sin of the even tone, 3: sin of the odd tone}, so that the epilogue finds the I (and the Q) of two
 * neighbouring tones in neighbouring registers: one packed fp32 operand */
static int upload_basis_panels(const anm_config_t *cfg, uint8_t **d_out, cudaStream_t stream) {
    const uint32_t N = cfg->sym_len, T = cfg->n_tones, S = cfg->hops_per_sym, H = N / S, KC = H / 16;
    if (2u * T != tc::kNcol) return ANM_ERR_UNSUPPORTED;
    std::vector<int8_t> q7((size_t)N * T * 2);
    if (anm_basis_q7(cfg, q7.data()) != ANM_OK) return ANM_ERR_ARG;
    constexpr uint32_t kFullPanel = tc::kNcol * 16u; /* one K chunk of ALL 128 basis rows; a CTA of a pair keeps half of the rows of each */
    std::vector<uint8_t> pan((size_t)S * KC * kFullPanel);
    for (uint32_t q = 0; q < S; ++q)
        for (uint32_t kc = 0; kc < KC; ++kc)
            for (uint32_t n = 0; n < tc::kNcol; ++n)
                for (uint32_t kk = 0; kk < 16; ++kk)
                    pan[((size_t)(q * KC + kc) * tc::kNcol + n) * 16 + kk] =
                        (uint8_t)q7[((size_t)(q * H + kc * 16 + kk) * T + (2u * (n / 4u) + (n & 1u))) * 2 + ((n >> 1) & 1u)];
    CK(cudaMalloc(d_out, pan.size()));
    /* pageable source: the runtime stages it before the call returns, the copy itself is ordered on `stream` */
    CK(cudaMemcpyAsync(*d_out, pan.data(), pan.size(), cudaMemcpyHostToDevice, stream));
    CK(cudaStreamSynchronize(stream));
    return ANM_OK;
}

static int set_device(const anm_demod *h) {
    CK(cudaSetDevice(h->device));
    return ANM_OK;
}

static void choose_launch(anm_demod *h) {
    if (h->var->dense) {
        /* one persistent CTA per SM (it owns all 512 TMEM columns) walks groups of four channels: 8 epilogue + 1 issuer +
         * 4 loader + 8 state-machine warps */
        h->warps_per_cta = tc::kWarps;
        /* two groups in flight per CTA; tc::kPair CTAs (a cluster on one TPC) walk their group pairs in lockstep */
        const uint32_t n_pairs = ((h->n_ch + 3u) / 4u + 1u) / 2u, n_units = (n_pairs + tc::kPair - 1u) / tc::kPair;
        h->grid = tc::kPair * std::min<uint32_t>((uint32_t)h->num_sms / tc::kPair, n_units);
        h->smem_bytes = h->var->cta_smem;
        return;
    }
    const uint32_t per_warp = h->var->warp_smem;
    const uint32_t smem_max = 227u * 1024u - h->var->cta_smem;
    uint32_t wmax = std::min<uint32_t>((uint32_t)kMaxWarps, smem_max / per_warp);
    if (wmax < 1) wmax = 1;
    const uint32_t sms = (uint32_t)h->num_sms;
    uint32_t W = wmax;
    if (h->n_ch <= sms * wmax) {
        W = std::max<uint32_t>(1u, (h->n_ch + sms - 1) / sms);
        h->grid = (h->n_ch + W - 1) / W;
    } else {
        /* channels beyond the first wave are handed out by the kernel's work queue: more resident warps
         * only add latency hiding */
        W = wmax;
        h->grid = sms;
    }
    if (const char *env = getenv("ANM_WARPS")) { /* experiment knob: force warps per CTA */
        const uint32_t w = (uint32_t)atoi(env);
        if (w >= 1 && w <= wmax) { W = w; h->grid = std::min<uint32_t>(sms, (h->n_ch + W - 1) / W); }
    }
    h->warps_per_cta = W;
    h->smem_bytes = (size_t)W * per_warp + h->var->cta_smem;
}

static int create_impl(anm_demod *h, const anm_config_t *cfg, const Variant *var, uint32_t n_channels, int device, uint32_t flags) {
    h->cfg = *cfg;
    h->var = var;
    h->device = device;
    h->n_ch = n_channels;
    h->flags = flags;
    CK(cudaSetDevice(device));
    CK(cudaDeviceGetAttribute(&h->num_sms, cudaDevAttrMultiProcessorCount, device));
    choose_launch(h);
    CK(cudaFuncSetAttribute((const void *)variant_fn(var, cfg, false), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(227 * 1024)));
    const uint32_t b = anm_bits_per_sym(cfg);
    const uint32_t hdr_syms = (24 + b - 1) / b;
    h->max_frame_syms = hdr_syms + ((cfg->max_payload + 2) * 8 + b - 1) / b;
    h->fsym_stride = (h->max_frame_syms + 63u) & ~63u;
    /* frames / payload bytes that may accumulate between two collects (about 4 minutes of back-to-back short frames per
     * channel); overflow drops frames and is reported by anm_demod_overflowed() */
    h->frames_cap = pow2_ceil(std::min<uint64_t>(1ull << 24, std::max<uint64_t>(4096u, (uint64_t)n_channels * 512u)));
    h->bytes_cap = pow2_ceil(std::min<uint64_t>(1ull << 30, std::max<uint64_t>(1u << 20, (uint64_t)n_channels * 32768u)));
    h->osym_cap = (flags & ANM_FLAG_SYMBOLS) ? 4096u : 0u;
    CK(cudaStreamCreateWithFlags(&h->own_stream, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&h->h2d_stream, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&h->d2h_stream, cudaStreamNonBlocking));
    h->last_stream = h->own_stream;
    CK(cudaMalloc(&h->d_state, (size_t)n_channels * var->state_bytes));
    CK(cudaMalloc(&h->d_fsyms, (size_t)n_channels * h->fsym_stride));
    CK(cudaMalloc(&h->d_frames, (size_t)h->frames_cap * sizeof(anm_frame_t)));
    CK(cudaMalloc(&h->d_bytes, h->bytes_cap));
    CK(cudaMalloc(&h->d_counters, 32));
    CK(cudaMalloc(&h->d_progress, (size_t)n_channels * sizeof(uint32_t)));
    if (h->osym_cap) CK(cudaMalloc(&h->d_osyms, (size_t)n_channels * h->osym_cap));
    h->h_tw.resize((size_t)cfg->sym_len * cfg->n_tones * 2);
    if (anm_config_foldable(cfg)) anm_fold_twiddles(cfg, h->h_tw.data()); /* [H/2][T][2], a prefix of the buffer */
    else anm_twiddles(cfg, h->h_tw.data());
    CK(cudaMalloc(&h->d_tw, h->h_tw.size() * sizeof(float)));
    /* uploads go through the handle's own stream: it is non-blocking, so the legacy default stream would not order them
     * before the first kernel (h_tw lives as long as the handle) */
    CK(cudaMemcpyAsync(h->d_tw, h->h_tw.data(), h->h_tw.size() * sizeof(float), cudaMemcpyHostToDevice, h->own_stream));
    h->evs.resize(64);
    for (EvPair &e : h->evs) {
        CK(cudaEventCreate(&e.a));
        CK(cudaEventCreate(&e.b));
    }
    CK(cudaHostAlloc(&h->snap, kSnapSlots * 16, cudaHostAllocDefault));
    memset(h->snap, 0, kSnapSlots * 16);
    for (size_t i = 0; i < kSnapSlots; ++i) CK(cudaEventCreateWithFlags(&h->snap_ev[i], cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&h->h2d_done, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&h->order_ev, cudaEventDisableTiming));
    for (int i = 0; i < 2; ++i) CK(cudaEventCreateWithFlags(&h->stage_free[i], cudaEventDisableTiming));
    h->q_syms.resize((flags & ANM_FLAG_SYMBOLS) ? n_channels : 0);
    /* constant part of the kernel parameters */
    KParams &k = h->kp;
    memset(&k, 0, sizeof k);
    k.n_ch = n_channels;
    k.state = h->d_state;
    k.state_stride = var->state_bytes;
    k.fsyms = h->d_fsyms;
    k.fsym_stride = h->fsym_stride;
    k.max_frame_syms = h->max_frame_syms;
    k.frames = h->d_frames;
    k.bytes = h->d_bytes;
    k.counters = h->d_counters;
    k.progress = h->d_progress;
    k.n_chunks = 1;
    k.frames_cap = h->frames_cap;
    k.bytes_cap = h->bytes_cap;
    k.osyms = h->d_osyms;
    k.osym_cap = h->osym_cap;
    k.P = cfg->preamble_len;
    k.tol = cfg->sync_tol;
    k.max_payload = cfg->max_payload;
    k.trk_epoch = cfg->trk_epoch;
    k.trk_thresh = cfg->trk_thresh;
    k.hdr_syms = hdr_syms;
    for (uint32_t j = 0; j < 7; ++j) {
        uint32_t m = 0;
        for (uint32_t pp = 0; pp < cfg->preamble_len; ++pp) m |= ((cfg->preamble[pp] >> j) & 1u) << pp;
        k.pre_plane[j] = m;
    }
    memcpy(k.preamble, cfg->preamble, ANM_MAX_PREAMBLE);
    k.tw_global = h->d_tw;
    set_tw_sign(cfg, &k);
    if (var->dense) {
        int rcb = upload_basis_panels(cfg, &h->d_basis, h->own_stream);
        if (rcb != ANM_OK) return rcb;
        k.tc_basis = h->d_basis;
    }
    return anm_demod_reset(h); /* synchronises own_stream: every upload above is complete when create returns */
}

extern "C" int anm_demod_create(const anm_config_t *cfg, uint32_t n_channels, int device, uint32_t flags,
                                anm_demod_t **out) {
    if (!cfg || !out || n_channels == 0) return ANM_ERR_ARG;
    if (anm_config_validate(cfg) != ANM_OK) { anm_set_error("invalid configuration"); return ANM_ERR_ARG; }
    const Variant *var = find_variant(cfg);
    if (!var) {
        anm_set_error("no kernel instance for T=%u N=%u S=%u", cfg->n_tones, cfg->sym_len, cfg->hops_per_sym);
        return ANM_ERR_UNSUPPORTED;
    }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        anm_set_error("no CUDA device: the demodulator has no CPU fallback");
        return ANM_ERR_CUDA;
    }
    if (device < 0 || device >= ndev) return ANM_ERR_ARG;
    anm_demod *h = new (std::nothrow) anm_demod();
    if (!h) return ANM_ERR_NOMEM;
    const int rc = create_impl(h, cfg, var, n_channels, device, flags);
    if (rc != ANM_OK) {
        anm_demod_destroy(h); /* releases whatever was allocated before the failure */
        return rc;
    }
    *out = h;
    return ANM_OK;
}

static int init_state(const Variant *var, unsigned char *d_state, uint32_t n_ch, cudaStream_t s) {
    /* zero everything, hop-decision history = 0xFF ("before the stream") */
    const size_t words = (size_t)n_ch * (var->state_bytes / 4u);
    k_init_state<<<(unsigned)((words + 255) / 256), 256, 0, s>>>(d_state, var->state_bytes, n_ch, 32u * var->S);
    CK(cudaGetLastError());
    return ANM_OK;
}

extern "C" int anm_demod_reset(anm_demod_t *h) {
    if (!h) return ANM_ERR_ARG;
    if (set_device(h)) return ANM_ERR_CUDA;
    CK(cudaStreamSynchronize(h->last_stream));
    CK(cudaStreamSynchronize(h->h2d_stream));
    CK(cudaStreamSynchronize(h->d2h_stream));
    int rc = init_state(h->var, h->d_state, h->n_ch, h->own_stream);
    if (rc) return rc;
    CK(cudaMemsetAsync(h->d_counters, 0, 32, h->own_stream));
    CK(cudaMemsetAsync(h->d_progress, 0, (size_t)h->n_ch * sizeof(uint32_t), h->own_stream));
    CK(cudaStreamSynchronize(h->own_stream));
    h->tickets = 0;
    h->chunks_done = 0;
    h->last_stream = h->own_stream;
    h->samples_fed = 0;
    h->syms_since_collect = 0;
    h->ev_used = 0;
    h->read_f = h->read_b = 0;
    h->seen_drops = 0;
    h->launches_since_reset = 0;
    h->drained_upto = 0;
    h->q.clear();
    for (auto &q : h->q_syms) q.clear();
    h->overflow = 0;
    return ANM_OK;
}

extern "C" void anm_demod_destroy(anm_demod_t *h) {
    if (!h) return;
    cudaSetDevice(h->device);
    cudaDeviceSynchronize();
    cudaFree(h->d_state);
    cudaFree(h->d_fsyms);
    cudaFree(h->d_frames);
    cudaFree(h->d_bytes);
    cudaFree(h->d_counters);
    cudaFree(h->d_progress);
    cudaFree(h->d_osyms);
    cudaFree(h->d_tw);
    cudaFree(h->d_basis);
    cudaFree(h->d_stage[0]);
    cudaFree(h->d_stage[1]);
    for (EvPair &e : h->evs) {
        if (e.a) cudaEventDestroy(e.a);
        if (e.b) cudaEventDestroy(e.b);
    }
    for (size_t i = 0; i < kSnapSlots; ++i)
        if (h->snap_ev[i]) cudaEventDestroy(h->snap_ev[i]);
    for (int i = 0; i < 2; ++i)
        if (h->stage_free[i]) cudaEventDestroy(h->stage_free[i]);
    if (h->h2d_done) cudaEventDestroy(h->h2d_done);
    if (h->order_ev) cudaEventDestroy(h->order_ev);
    h->q.frames.release(); /* pinned storage: free it while the context is current */
    h->q.bytes.release();
    if (h->snap) cudaFreeHost(h->snap);
    if (h->d2h_stream) cudaStreamDestroy(h->d2h_stream);
    if (h->h2d_stream) cudaStreamDestroy(h->h2d_stream);
    if (h->own_stream) cudaStreamDestroy(h->own_stream);
    delete h;
}

static int launch(anm_demod *h, KParams &k, cudaStream_t s, bool timed) {
    EvPair *ev = nullptr;
    if (timed && h->ev_used < h->evs.size()) ev = &h->evs[h->ev_used++];
    /* One launch of a handle at a time: a launch on another stream than the previous one is ordered behind it (the
     * per-channel state, the work-queue tickets and the snapshot sequence all assume it). */
    if (s != h->last_stream) {
        CK(cudaEventRecord(h->order_ev, h->last_stream));
        CK(cudaStreamWaitEvent(s, h->order_ev, 0));
    }
    /* frame / byte queues are rings addressed by monotonically increasing counters; the kernel may
     * fill them up to one capacity beyond what the host has consumed so far */
    k.base_f = h->read_f;
    k.base_b = h->read_b;
    const uint64_t seq = h->launches_since_reset;
    const size_t slot = (size_t)(seq % kSnapSlots);
    k.q_base = (uint32_t)h->tickets;
    k.prog_base = h->chunks_done;
    const bool multi = k.n_chunks > 1u;
    h->tickets += (uint64_t)h->n_ch * k.n_chunks + (multi ? (uint64_t)h->grid * h->warps_per_cta : 0u);
    if (multi) h->chunks_done += k.n_chunks;
    k.done_base = (uint32_t)(seq * (uint64_t)h->grid * h->warps_per_cta);
    k.snap = h->snap + slot * 4;
    k.seq1 = (uint32_t)(seq + 1);
    if (ev) CK(cudaEventRecord(ev->a, s));
    void *args[] = {(void *)&k};
    CK(cudaLaunchKernel((const void *)variant_fn(h->var, &h->cfg, false), dim3(h->grid), dim3(h->warps_per_cta * 32), args, h->smem_bytes, s));
    if (ev) CK(cudaEventRecord(ev->b, s));
    /* the kernel's last warp writes the queue counters as they stand after this launch into snap[slot] (pinned host
     * memory); the event tells the host when that has happened.  A later drain can therefore stop exactly behind this
     * launch while newer launches are in flight -- no extra copy in the stream. */
    CK(cudaEventRecord(h->snap_ev[slot], s));
    h->launches++;
    h->launches_since_reset++;
    h->last_stream = s;
    return ANM_OK;
}

/* Moves the frames produced up to (and including) launch number `seq` (counted since reset) into the host queue.  Waits
 * only for that launch; later launches may still be running or queued behind a host->device copy. */
static int drain_frames(anm_demod *h, uint64_t seq) {
    if (seq + 1 <= h->drained_upto) return ANM_OK; /* an older launch than what was already handed out: nothing new */
    if (h->launches_since_reset - seq > kSnapSlots) {
        /* the launch's snapshot slot has been reused: fall back to the state after everything submitted so far */
        seq = h->launches_since_reset - 1;
    }
    const size_t slot = (size_t)(seq % kSnapSlots);
    CK(cudaEventSynchronize(h->snap_ev[slot]));
    volatile const uint32_t *sn = h->snap + slot * 4;
    if (sn[3] != (uint32_t)(seq + 1)) {
        anm_set_error("internal: snapshot of launch %llu holds sequence %u", (unsigned long long)seq, sn[3] - 1u);
        return ANM_ERR_CUDA;
    }
    const uint32_t wf = sn[0], wb = sn[1], drops = sn[2];
    h->drained_upto = seq + 1;
    const uint32_t nf = wf - h->read_f, nb = wb - h->read_b; /* modulo 2^32 */
    if (drops != h->seen_drops || nf > h->frames_cap || nb > h->bytes_cap) {
        /* Records of dropped frames were never written: discard what is pending and resynchronise on the snapshot.  The
         * device's drop counter only ever counts up and is never cleared while launches may be in flight; launches
         * submitted after this point take the new read positions as their base. */
        h->overflow = 1;
        h->seen_drops = drops;
        h->read_f = wf;
        h->read_b = wb;
        return ANM_OK;
    }
    if (nf) {
        anm_frame_t *fdst;
        uint8_t *bdst;
        const size_t b0 = h->q.bytes.n;
        if (!h->q.grow(nf, nb, &fdst, &bdst)) { anm_set_error("out of host memory for %u frames", nf); return ANM_ERR_NOMEM; }
        const size_t b0q = (h->q.bytes.n - nb); /* grow() may have restarted the queue at the front */
        (void)b0;
        const uint32_t fpos = h->read_f & (h->frames_cap - 1), f1 = std::min(nf, h->frames_cap - fpos);
        CK(cudaMemcpyAsync(fdst, h->d_frames + fpos, (size_t)f1 * sizeof(anm_frame_t), cudaMemcpyDeviceToHost, h->d2h_stream));
        if (nf > f1) CK(cudaMemcpyAsync(fdst + f1, h->d_frames, (size_t)(nf - f1) * sizeof(anm_frame_t), cudaMemcpyDeviceToHost, h->d2h_stream));
        if (nb) {
            const uint32_t bpos = h->read_b & (h->bytes_cap - 1), b1 = std::min(nb, h->bytes_cap - bpos);
            CK(cudaMemcpyAsync(bdst, h->d_bytes + bpos, b1, cudaMemcpyDeviceToHost, h->d2h_stream));
            if (nb > b1) CK(cudaMemcpyAsync(bdst + b1, h->d_bytes, nb - b1, cudaMemcpyDeviceToHost, h->d2h_stream));
        }
        CK(cudaStreamSynchronize(h->d2h_stream));
        for (uint32_t i = 0; i < nf; ++i) fdst[i].offset = (uint32_t)b0q + (fdst[i].offset - h->read_b);
        h->read_f = wf;
        h->read_b = wb;
    }
    return ANM_OK;
}

/* n_chunks consecutive chunks of n_samples per channel, chunk c of a channel chunk_stride samples behind chunk c - 1, in ONE launch of k_demod
 * (work items = (chunk, channel): no idle tail between the chunks); configurations on the tensor-core kernel and handles that record decided symbols
 * take the chunks one launch at a time */
extern "C" int anm_demod_feed_device_chunks(anm_demod_t *h, const int16_t *d_pcm, size_t ch_stride, size_t chunk_stride, size_t n_samples, uint32_t n_chunks,
                                            void *stream) {
    if (!h || (!d_pcm && n_samples && n_chunks)) return ANM_ERR_ARG;
    if (n_samples == 0 || n_chunks == 0) return ANM_OK;
    const uint32_t N = h->cfg.sym_len;
    if (n_samples % N) { anm_set_error("n_samples must be a multiple of sym_len=%u", N); return ANM_ERR_ALIGN; }
    if ((reinterpret_cast<uintptr_t>(d_pcm) & 15u) || ((ch_stride * 2) & 15u) || ch_stride < n_samples || (n_chunks > 1 && ((chunk_stride * 2) & 15u))) {
        anm_set_error("pcm base, channel stride and chunk stride must be 16-byte aligned, channel stride >= n_samples");
        return ANM_ERR_ALIGN;
    }
    if (set_device(h)) return ANM_ERR_CUDA;
    const uint64_t nsyms = n_samples / N;
    if (n_chunks > 1 && (h->var->dense || h->osym_cap || (uint64_t)h->n_ch * n_chunks >= (1ull << 31))) {
        for (uint32_t c = 0; c < n_chunks; ++c) {
            const int rc = anm_demod_feed_device_chunks(h, d_pcm + (size_t)c * chunk_stride, ch_stride, 0, n_samples, 1, stream);
            if (rc) return rc;
        }
        return ANM_OK;
    }
    if (h->osym_cap && h->syms_since_collect + nsyms + 8 > h->osym_cap) {
        long rc = anm_demod_collect(h);
        if (rc < 0) return (int)rc;
        if (nsyms + 8 > h->osym_cap) { anm_set_error("chunk longer than the symbol buffer"); return ANM_ERR_ARG; }
    }
    cudaStream_t s = (cudaStream_t)stream; /* NULL is the legacy default stream, as in the CUDA runtime */
    KParams k = h->kp;
    k.pcm = d_pcm;
    k.ch_stride = ch_stride;
    k.chunk_stride = chunk_stride;
    k.n_chunks = n_chunks;
    k.n_syms = (uint32_t)nsyms;
    k.hop_base = h->samples_fed / (N / h->cfg.hops_per_sym);
    int rc = launch(h, k, s, true);
    if (rc) return rc;
    h->samples_fed += n_samples * (uint64_t)n_chunks;
    h->syms_since_collect += nsyms * n_chunks;
    return ANM_OK;
}

extern "C" int anm_demod_feed_device(anm_demod_t *h, const int16_t *d_pcm, size_t ch_stride, size_t n_samples, void *stream) {
    return anm_demod_feed_device_chunks(h, d_pcm, ch_stride, 0, n_samples, 1, stream);
}

static int feed_host_impl(anm_demod_t *h, const int16_t *h_pcm, size_t ch_stride, size_t n_samples, bool async) {
    if (!h || (!h_pcm && n_samples)) return ANM_ERR_ARG;
    if (n_samples == 0) return ANM_OK;
    if (n_samples % h->cfg.sym_len) return ANM_ERR_ALIGN;
    if (ch_stride < n_samples) return ANM_ERR_ARG;
    if (set_device(h)) return ANM_ERR_CUDA;
    const size_t need = (size_t)h->n_ch * n_samples;
    if (need > h->stage_cap) {
        CK(cudaStreamSynchronize(h->last_stream));
        CK(cudaStreamSynchronize(h->h2d_stream));
        for (int i = 0; i < 2; ++i) {
            cudaFree(h->d_stage[i]);
            h->d_stage[i] = nullptr;
        }
        h->stage_cap = 0;
        for (int i = 0; i < 2; ++i) CK(cudaMalloc(&h->d_stage[i], need * sizeof(int16_t)));
        h->stage_cap = need;
        /* fresh events: nothing has read the new buffers */
        for (int i = 0; i < 2; ++i) CK(cudaEventRecord(h->stage_free[i], h->own_stream));
    }
    const uint32_t sb = h->stage_next;
    h->stage_next ^= 1u;
    int16_t *dst = h->d_stage[sb];
    /* the copy runs on its own stream: it only waits for the kernel that last read this staging buffer (two feeds ago),
     * so it overlaps the kernel of the previous chunk */
    CK(cudaStreamWaitEvent(h->h2d_stream, h->stage_free[sb], 0));
    if (ch_stride == n_samples)
        CK(cudaMemcpyAsync(dst, h_pcm, need * sizeof(int16_t), cudaMemcpyHostToDevice, h->h2d_stream));
    else
        CK(cudaMemcpy2DAsync(dst, n_samples * 2, h_pcm, ch_stride * 2, n_samples * 2, h->n_ch, cudaMemcpyHostToDevice, h->h2d_stream));
    CK(cudaEventRecord(h->h2d_done, h->h2d_stream));
    CK(cudaStreamWaitEvent(h->own_stream, h->h2d_done, 0));
    int rc = anm_demod_feed_device(h, dst, n_samples, n_samples, h->own_stream);
    if (rc) return rc;
    CK(cudaEventRecord(h->stage_free[sb], h->own_stream));
    if (!async) CK(cudaStreamSynchronize(h->own_stream));
    return ANM_OK;
}

extern "C" int anm_demod_feed_host(anm_demod_t *h, const int16_t *h_pcm, size_t ch_stride, size_t n_samples) {
    return feed_host_impl(h, h_pcm, ch_stride, n_samples, false);
}

extern "C" int anm_demod_feed_host_async(anm_demod_t *h, const int16_t *h_pcm, size_t ch_stride, size_t n_samples) {
    return feed_host_impl(h, h_pcm, ch_stride, n_samples, true);
}

extern "C" int anm_demod_wait_input(anm_demod_t *h) {
    if (!h) return ANM_ERR_ARG;
    if (set_device(h)) return ANM_ERR_CUDA;
    CK(cudaEventSynchronize(h->h2d_done));
    return ANM_OK;
}

extern "C" long anm_demod_collect_upto(anm_demod_t *h, uint32_t lag) {
    if (!h) return ANM_ERR_ARG;
    if (set_device(h)) return ANM_ERR_CUDA;
    if (lag >= kSnapSlots - 1) return ANM_ERR_ARG;
    if (h->launches_since_reset > lag) {
        int rc = drain_frames(h, h->launches_since_reset - 1 - lag);
        if (rc) return rc;
    }
    return (long)h->q.pending();
}

extern "C" long anm_demod_collect(anm_demod_t *h) {
    if (!h) return ANM_ERR_ARG;
    if (set_device(h)) return ANM_ERR_CUDA;
    cudaStream_t s = h->last_stream;
    CK(cudaStreamSynchronize(s));
    if (h->launches_since_reset) {
        int rc = drain_frames(h, h->launches_since_reset - 1);
        if (rc) return rc;
    }
    if (h->osym_cap && h->syms_since_collect) {
        std::vector<uint32_t> oc(h->n_ch);
        const size_t off = offsetof(ChanScalars, osym_cnt);
        CK(cudaMemcpy2DAsync(oc.data(), 4, h->d_state + off, h->var->state_bytes, 4, h->n_ch, cudaMemcpyDeviceToHost, s));
        CK(cudaStreamSynchronize(s));
        uint32_t mx = 0;
        for (uint32_t c : oc) mx = std::max(mx, std::min(c, h->osym_cap));
        if (mx) {
            std::vector<uint8_t> tmp((size_t)h->n_ch * mx);
            CK(cudaMemcpy2DAsync(tmp.data(), mx, h->d_osyms, h->osym_cap, mx, h->n_ch, cudaMemcpyDeviceToHost, s));
            CK(cudaStreamSynchronize(s));
            for (uint32_t c = 0; c < h->n_ch; ++c) {
                const uint32_t n = std::min(oc[c], h->osym_cap);
                if (oc[c] > h->osym_cap) h->overflow = 1;
                h->q_syms[c].insert(h->q_syms[c].end(), tmp.begin() + (size_t)c * mx, tmp.begin() + (size_t)c * mx + n);
            }
        }
        /* on the launching stream (non-blocking streams are not ordered against the legacy default stream), and complete
         * before the next kernel can read osym_cnt */
        CK(cudaMemset2DAsync(h->d_state + off, h->var->state_bytes, 0, 4, h->n_ch, s));
        CK(cudaStreamSynchronize(s));
    }
    h->syms_since_collect = 0;
    if (h->overflow) { anm_set_error("an output queue overflowed; some frames or symbols were dropped"); }
    return (long)h->q.pending();
}

extern "C" size_t anm_demod_read_frames(anm_demod_t *h, anm_frame_t *out, size_t cap, uint8_t *bytes, size_t bytes_cap) {
    if (!h || !out || !cap) return 0;
    return h->q.pop_sorted(h->n_ch, out, cap, bytes, bytes_cap);
}

extern "C" size_t anm_demod_take_frames(anm_demod_t *h, anm_frame_t *out, size_t cap, uint8_t *bytes, size_t bytes_cap, size_t *n_bytes) {
    if (!h || !out) return 0;
    return h->q.take_all(out, cap, bytes, bytes_cap, n_bytes);
}

extern "C" size_t anm_demod_peek_frames(anm_demod_t *h, const anm_frame_t **frames, const uint8_t **bytes, size_t *n_bytes) {
    if (!h || !frames || h->q.cursor != 0) return 0;
    *frames = h->q.frames.p;
    if (bytes) *bytes = h->q.bytes.p;
    if (n_bytes) *n_bytes = h->q.bytes.n;
    return h->q.frames.n;
}

extern "C" void anm_demod_drop_frames(anm_demod_t *h) {
    if (h) h->q.clear();
}

extern "C" void anm_frames_summary(const anm_frame_t *frames, size_t n, uint64_t *n_ok, uint64_t *payload_bytes_ok) {
    uint64_t ok = 0, by = 0;
    for (size_t i = 0; i < n; ++i) {
        ok += frames[i].crc_ok ? 1u : 0u;
        by += frames[i].crc_ok ? frames[i].len : 0u;
    }
    if (n_ok) *n_ok = ok;
    if (payload_bytes_ok) *payload_bytes_ok = by;
}

extern "C" int anm_demod_frame_rings(const anm_demod_t *h, const anm_frame_t **d_frames, uint32_t *frames_mask, const uint8_t **d_bytes,
                                     uint32_t *bytes_mask) {
    if (!h) return ANM_ERR_ARG;
    if (d_frames) *d_frames = h->d_frames;
    if (frames_mask) *frames_mask = h->frames_cap - 1u;
    if (d_bytes) *d_bytes = h->d_bytes;
    if (bytes_mask) *bytes_mask = h->bytes_cap - 1u;
    return ANM_OK;
}

extern "C" size_t anm_demod_read_symbols(anm_demod_t *h, uint32_t channel, uint8_t *out, size_t cap) {
    if (!h || !out || channel >= h->q_syms.size()) return 0;
    std::deque<uint8_t> &q = h->q_syms[channel];
    const size_t n = std::min(cap, q.size());
    std::copy(q.begin(), q.begin() + n, out);
    q.erase(q.begin(), q.begin() + n);
    return n;
}

extern "C" int anm_demod_stats(anm_demod_t *h, anm_chan_stats_t *out) {
    if (!h || !out) return ANM_ERR_ARG;
    if (set_device(h)) return ANM_ERR_CUDA;
    CK(cudaStreamSynchronize(h->last_stream));
    CK(cudaMemcpy2D(out, sizeof(anm_chan_stats_t), h->d_state + offsetof(ChanScalars, stats), h->var->state_bytes,
                    sizeof(anm_chan_stats_t), h->n_ch, cudaMemcpyDeviceToHost));
    return ANM_OK;
}

extern "C" uint64_t anm_demod_launch_count(const anm_demod_t *h) { return h ? h->launches : 0; }

/* 1 if a frame/symbol queue overflowed since create/reset (frames were dropped) */
extern "C" int anm_demod_overflowed(const anm_demod_t *h) { return h ? h->overflow : 0; }

/* sums the device time of the kernels launched since the last call (at most 64 are
 * tracked between calls); returns the number of launches summed */
extern "C" int anm_demod_kernel_time(anm_demod_t *h, float *sum_ms) {
    if (!h || !sum_ms) return ANM_ERR_ARG;
    if (set_device(h)) return ANM_ERR_CUDA;
    float tot = 0.f;
    const size_t n = h->ev_used;
    for (size_t i = 0; i < n; ++i) {
        float ms = 0.f;
        CK(cudaEventSynchronize(h->evs[i].b));
        CK(cudaEventElapsedTime(&ms, h->evs[i].a, h->evs[i].b));
        tot += ms;
    }
    h->ev_used = 0;
    *sum_ms = tot;
    return (int)n;
}

extern "C" float anm_demod_last_kernel_ms(anm_demod_t *h) {
    if (!h || h->ev_used == 0) return -1.f;
    float ms = -1.f;
    if (cudaSetDevice(h->device) != cudaSuccess) return -1.f;
    const EvPair &e = h->evs[h->ev_used - 1];
    if (cudaEventSynchronize(e.b) != cudaSuccess) return -1.f;
    if (cudaEventElapsedTime(&ms, e.a, e.b) != cudaSuccess) return -1.f;
    return ms;
}

extern "C" int anm_demod_launch_geometry(const anm_demod_t *h, uint32_t *grid, uint32_t *warps_per_cta, uint32_t *smem) {
    if (!h) return ANM_ERR_ARG;
    if (grid) *grid = h->grid;
    if (warps_per_cta) *warps_per_cta = h->warps_per_cta;
    if (smem) *smem = (uint32_t)h->smem_bytes;
    return ANM_OK;
}

/* ---- stateless tone-energy pass ----------------------------------------------- */
extern "C" int anm_tone_energies_device(const anm_config_t *cfg, const int16_t *d_pcm, uint32_t n_ch, size_t ch_stride,
                                        size_t n_samples, float *d_energy, uint8_t *d_sym, float *d_emax, void *stream) {
    if (!cfg || !d_pcm || n_ch == 0) return ANM_ERR_ARG;
    if (anm_config_validate(cfg) != ANM_OK) return ANM_ERR_ARG;
    const Variant *var = find_variant(cfg);
    if (!var) return ANM_ERR_UNSUPPORTED;
    if (n_samples % cfg->sym_len) return ANM_ERR_ALIGN;
    if ((reinterpret_cast<uintptr_t>(d_pcm) & 15u) || ((ch_stride * 2) & 15u)) return ANM_ERR_ALIGN;
    if (n_samples == 0) return ANM_OK;
    int dev = 0, sms = 0;
    CK(cudaGetDevice(&dev));
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    cudaStream_t s = (cudaStream_t)stream;
    unsigned char *d_state = nullptr;
    float2 *d_tw = nullptr;
    std::vector<float> tw((size_t)cfg->sym_len * cfg->n_tones * 2);
    if (anm_config_foldable(cfg)) anm_fold_twiddles(cfg, tw.data());
    else anm_twiddles(cfg, tw.data());
    CK(cudaMalloc(&d_state, (size_t)n_ch * var->state_bytes));
    if (cudaMalloc(&d_tw, tw.size() * sizeof(float)) != cudaSuccess) {
        cudaFree(d_state);
        anm_set_error("tone pass: out of device memory");
        return ANM_ERR_CUDA;
    }
    int rc = init_state(var, d_state, n_ch, s);
    if (rc == ANM_OK && cudaMemcpyAsync(d_tw, tw.data(), tw.size() * sizeof(float), cudaMemcpyHostToDevice, s) != cudaSuccess) rc = ANM_ERR_CUDA;
    if (rc == ANM_OK) {
        KParams k;
        memset(&k, 0, sizeof k);
        k.n_chunks = 1;
        k.pcm = d_pcm;
        k.ch_stride = ch_stride;
        k.n_ch = n_ch;
        k.n_syms = (uint32_t)(n_samples / cfg->sym_len);
        k.hop_base = 0;
        k.state = d_state;
        k.state_stride = var->state_bytes;
        set_tw_sign(cfg, &k);
        k.trE = d_energy;
        k.trD = d_sym;
        k.trEmax = d_emax;
        k.tr_hops = n_samples / (cfg->sym_len / cfg->hops_per_sym);
        k.P = cfg->preamble_len;
        k.tw_global = d_tw;
        uint8_t *d_basis = nullptr;
        uint32_t W, grid;
        size_t smem;
        if (var->dense) {
            rc = upload_basis_panels(cfg, &d_basis, s);
            k.tc_basis = d_basis;
            W = tc::kWarps;
            const uint32_t n_pairs = ((n_ch + 3u) / 4u + 1u) / 2u, n_units = (n_pairs + tc::kPair - 1u) / tc::kPair;
            grid = tc::kPair * std::min<uint32_t>((uint32_t)sms / tc::kPair, n_units);
            smem = var->cta_smem;
        } else {
            W = std::min<uint32_t>(8u, (227u * 1024u - var->cta_smem) / var->warp_smem);
            grid = std::min<uint32_t>((n_ch + W - 1) / W, (uint32_t)sms * 4u);
            smem = (size_t)W * var->warp_smem + var->cta_smem;
        }
        cudaFuncSetAttribute((const void *)variant_fn(var, cfg, true), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(227 * 1024));
        void *args[] = {(void *)&k};
        cudaError_t e = rc == ANM_OK ? cudaLaunchKernel((const void *)variant_fn(var, cfg, true), dim3(grid), dim3(W * 32), args, smem, s) : cudaSuccess;
        if (rc == ANM_OK && e == cudaSuccess) {
            const cudaError_t es = cudaStreamSynchronize(s);
            if (es != cudaSuccess) { anm_set_error("tone pass: %s", cudaGetErrorString(es)); rc = ANM_ERR_CUDA; }
        }
        cudaFree(d_basis);
        if (e != cudaSuccess) { anm_set_error("launch: %s", cudaGetErrorString(e)); rc = ANM_ERR_CUDA; }
    }
    cudaError_t e2 = cudaStreamSynchronize(s);
    if (rc == ANM_OK && e2 != cudaSuccess) { anm_set_error("tone pass: %s", cudaGetErrorString(e2)); rc = ANM_ERR_CUDA; }
    cudaFree(d_state);
    cudaFree(d_tw);
    return rc;
}

/* ---- tx params helper: fills the host-computed noise scale the GPU renderer reads ---- */
extern "C" void anm_tx_params_prepare(anm_tx_params_t *p, size_t n) {
    for (size_t i = 0; i < n; ++i) p[i].reserved = anm_tx_noise_scale(p[i].amplitude_q15, p[i].snr_mdb);
}

/* ---- firmware-idiom single-channel interface ---------------------------------------- */
struct demod {
    anm_demod_t *h;
    std::vector<int16_t> pend; /* samples not yet forming a whole symbol period */
    anm_pb_queue_t *pbq;       /* CRC-valid payloads handed to the protobuf decoder */
    std::vector<uint8_t> buf;  /* one frame's payload on its way out */
};
/* demod_initialize() follows the reference's `<module>_initialize()` idiom (hardware/README.md:10-14): it sets the
 * DEFAULT configuration of demodulators created later with demod_create().  Every demodulator owns a copy of its
 * configuration from then on (anm_demod.cfg); demod_create_cfg() takes one directly and needs no global at all. */
static anm_config_t g_cfg;
static bool g_cfg_set = false;
static std::mutex g_cfg_mu;

extern "C" int demod_initialize(const anm_config_t *cfg) {
    if (!cfg || anm_config_validate(cfg) != ANM_OK) return ANM_ERR_ARG;
    std::lock_guard<std::mutex> lk(g_cfg_mu);
    g_cfg = *cfg;
    g_cfg_set = true;
    return ANM_OK;
}
extern "C" demod_t *demod_create_cfg(const anm_config_t *cfg) {
    if (!cfg) return nullptr;
    demod *d = new (std::nothrow) demod();
    if (!d) return nullptr;
    int dev = 0;
    cudaGetDevice(&dev);
    if (anm_demod_create(cfg, 1, dev, ANM_FLAG_SYMBOLS, &d->h) != ANM_OK) { delete d; return nullptr; }
    d->buf.resize(4104);
    return d;
}
extern "C" demod_t *demod_create(void) {
    anm_config_t cfg;
    {
        std::lock_guard<std::mutex> lk(g_cfg_mu);
        if (!g_cfg_set) { anm_set_error("demod_initialize() was not called"); return nullptr; }
        cfg = g_cfg;
    }
    return demod_create_cfg(&cfg);
}
extern "C" int demod_feed(demod_t *d, const int16_t *pcm, size_t n) {
    if (!d || (!pcm && n)) return ANM_ERR_ARG;
    d->pend.insert(d->pend.end(), pcm, pcm + n);
    const size_t N = d->h->cfg.sym_len;
    const size_t whole = d->pend.size() / N * N;
    if (!whole) return ANM_OK;
    int rc = anm_demod_feed_host(d->h, d->pend.data(), whole, whole);
    if (rc) return rc;
    d->pend.erase(d->pend.begin(), d->pend.begin() + whole);
    long c = anm_demod_collect(d->h);
    return c < 0 ? (int)c : ANM_OK;
}
extern "C" size_t demod_read_symbols(demod_t *d, uint8_t *out, size_t cap) {
    return d ? anm_demod_read_symbols(d->h, 0, out, cap) : 0;
}
extern "C" size_t demod_read_frames(demod_t *d, demod_frame_t *out, size_t cap) {
    if (!d || !out) return 0;
    size_t n = 0;
    while (n < cap) { /* one pop per frame from the queue's cursor: linear in the number of frames */
        anm_frame_t f;
        if (anm_demod_read_frames(d->h, &f, 1, d->buf.data(), d->buf.size()) != 1) break;
        out[n].sample_offset = f.start_sample;
        out[n].len = f.len;
        out[n].crc_ok = f.crc_ok;
        memcpy(out[n].bytes, d->buf.data(), f.len);
        ++n;
    }
    return n;
}
extern "C" void demod_destroy(demod_t *d) {
    if (!d) return;
    anm_demod_destroy(d->h);
    anm_pb_queue_destroy(d->pbq);
    delete d;
}

/* SURVEY.md 8(b) seam #1: the byte source pb_decode_delimited() reads, standing where
 * network_pb_istream_from_socket() (hardware/src/network.cpp:299-305) stands today. */
extern "C" anm_pb_istream_t demod_as_pb_istream(demod_t *d) {
    anm_pb_istream_t none = {nullptr, nullptr, 0, "no demodulator"};
    if (!d) return none;
    if (!d->pbq) d->pbq = anm_pb_queue_create();
    anm_frame_t f;
    while (anm_demod_read_frames(d->h, &f, 1, d->buf.data(), d->buf.size()) == 1)
        if (f.crc_ok) anm_pb_queue_push(d->pbq, d->buf.data(), f.len);
    return anm_pb_istream_from_queue(d->pbq);
}
