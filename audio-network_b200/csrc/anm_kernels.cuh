/*
 * anm_kernels.cuh -- sm_100a kernels of the SPEC.md receive path.
 *
 * One warp owns one channel.  Within a step of 32 symbol periods, lane l owns symbol
 * period l (N samples = S hops), so all per-sample work is lane-private register
 * arithmetic (SPEC 3's FMA chains as packed fma.rn.f32x2), the S-hop window tree needs
 * S-1 shuffles per tone per step, the preamble correlation is ballots + popc over
 * bit-planes of the hop decisions, and symbol slicing / tracking / framing are
 * lane-parallel over up to 32 symbols at a time.  PCM moves HBM -> shared memory with
 * coalesced 16-byte cp.async into an XOR-swizzled double buffer, and is read back with
 * conflict-free LDS.128.
 *
 * There is no reference kernel for any of this; the behaviour is
 * SPEC.md's, the structure is B200-first.
 */
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "anm_internal.h"

namespace anm {

constexpr int kMaxConstTw = 6144; /* float2 entries = 48 KB of constant bank 3 */
__constant__ float2 c_tw[kMaxConstTw]; /* single translation unit: anm_cuda.cu */

enum : uint32_t { ST_SEARCH = 0, ST_PEAK = 1, ST_HEADER = 2, ST_BODY = 3 };

/* Uniform (per-channel) part of the carried state; lane-distributed parts follow it. */
struct ChanScalars {
    uint64_t peak_end, best_h, t0, next, prev_hop;
    float best_q;
    uint32_t state, nsym, total, flen;
    int32_t acc;
    uint32_t s_prev, s_prev2;
    uint32_t osym_cnt;
    anm_chan_stats_t stats; /* 32 bytes */
    uint32_t pad[4];
};
static_assert(sizeof(ChanScalars) == 128, "ChanScalars layout");

struct KParams {
    const int16_t *pcm;
    unsigned long long ch_stride; /* samples */
    uint32_t n_ch;
    uint32_t n_syms;              /* symbol periods per channel in this chunk */
    unsigned long long hop_base;  /* absolute index of the chunk's first hop */
    unsigned char *state;         /* per channel: ChanScalars | lane records | tree carry */
    uint32_t state_stride;        /* bytes */
    uint32_t do_sm;               /* 0: tone energies only (stateless trace pass) */
    uint8_t *fsyms;               /* per channel frame symbol store */
    uint32_t fsym_stride;
    uint32_t max_frame_syms;
    float *trE;                   /* [n_ch][tr_hops][T] or NULL */
    uint8_t *trD;                 /* [n_ch][tr_hops] or NULL */
    float *trEmax;                /* [n_ch][tr_hops] or NULL */
    unsigned long long tr_hops;
    anm_frame_t *frames;
    uint8_t *bytes;
    uint32_t *counters;           /* [0]=n_frames [1]=n_bytes [2]=overflow flags */
    uint32_t frames_cap, bytes_cap;
    uint8_t *osyms;               /* [n_ch][osym_cap] or NULL */
    uint32_t osym_cap;
    uint32_t P, tol, max_payload, trk_epoch, trk_thresh, hdr_syms;
    uint32_t pre_plane[7];        /* bit-planes of the preamble tone indices */
    uint8_t preamble[ANM_MAX_PREAMBLE];
    const float2 *tw_global;      /* [N][T] (cos, sin); used when the table exceeds c_tw */
};

__device__ __forceinline__ float2 ffma2(float a, float2 b, float2 c) {
    unsigned long long rb, rc, rd, ra;
    ra = ((unsigned long long)__float_as_uint(a) << 32) | __float_as_uint(a);
    rb = ((unsigned long long)__float_as_uint(b.y) << 32) | __float_as_uint(b.x);
    rc = ((unsigned long long)__float_as_uint(c.y) << 32) | __float_as_uint(c.x);
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rd) : "l"(ra), "l"(rb), "l"(rc));
    return make_float2(__uint_as_float((uint32_t)rd), __uint_as_float((uint32_t)(rd >> 32)));
}
__device__ __forceinline__ float2 fadd2(float2 a, float2 b) {
    unsigned long long ra, rb, rd;
    ra = ((unsigned long long)__float_as_uint(a.y) << 32) | __float_as_uint(a.x);
    rb = ((unsigned long long)__float_as_uint(b.y) << 32) | __float_as_uint(b.x);
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(rd) : "l"(ra), "l"(rb));
    return make_float2(__uint_as_float((uint32_t)rd), __uint_as_float((uint32_t)(rd >> 32)));
}
__device__ __forceinline__ float2 shfl2(float2 v, int src) {
    return make_float2(__shfl_sync(0xffffffffu, v.x, src), __shfl_sync(0xffffffffu, v.y, src));
}
__device__ __forceinline__ void cp_async16(uint32_t smem_addr, const void *g) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_addr), "l"(g) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int NKEEP>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(NKEEP) : "memory");
}
__device__ __forceinline__ uint4 lds128(uint32_t smem_addr) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(smem_addr));
    return v;
}

/* select element ph (warp-uniform, runtime) of a register array without local memory */
template <int S, typename V>
__device__ __forceinline__ V pick(const V (&r)[S], int ph) {
    V v = r[0];
#pragma unroll
    for (int i = 1; i < S; ++i) v = (ph == i) ? r[i] : v;
    return v;
}
__device__ __forceinline__ uint32_t gray_inv(uint32_t g) {
    g ^= g >> 1;
    g ^= g >> 2;
    g ^= g >> 4;
    return g;
}
__device__ __forceinline__ int floordiv(int a, int b) { return (a >= 0) ? a / b : -((-a + b - 1) / b); }

template <int T>
struct Log2 { static constexpr int v = 1 + Log2<T / 2>::v; };
template <>
struct Log2<1> { static constexpr int v = 0; };

/* bytes of lane-distributed + carry state after ChanScalars */
template <int T, int S>
__host__ __device__ constexpr uint32_t state_bytes() {
    return (uint32_t)sizeof(ChanScalars) + 32u * S * 8u /* d (u32) + emax per lane per phase */ + (uint32_t)(S - 1) * T * 8u;
}
template <int N>
__host__ __device__ constexpr uint32_t stage_bytes() { return 32u * N * 2u; }
template <int T, int N, int S>
__host__ __device__ constexpr uint32_t warp_smem_bytes() { return 2u * stage_bytes<N>() + (uint32_t)(S - 1) * T * 8u; }

template <int T, int N, int S, bool TWC>
__global__ void __launch_bounds__(512) k_demod(const __grid_constant__ KParams p) {
    constexpr int H = N / S;
    constexpr int B = Log2<T>::v;
    constexpr int TG = T < 8 ? T : 8; /* tones per register group */
    constexpr int NG = T / TG;
    constexpr int LV = Log2<S>::v;
    constexpr int CPS = N / 8; /* 16-byte chunks per symbol period */
    constexpr uint32_t FULL = 0xffffffffu;
    static_assert(N >= 64 && (H % 8) == 0, "unsupported geometry");

    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ uint16_t s_crc[256];
    for (int i = threadIdx.x; i < 256; i += blockDim.x) {
        uint32_t c = (uint32_t)i << 8;
        for (int k = 0; k < 8; ++k) c = (c & 0x8000u) ? ((c << 1) ^ 0x1021u) : (c << 1);
        s_crc[i] = (uint16_t)c;
    }
    __syncthreads();

    const int lane = threadIdx.x & 31;
    const int wib = threadIdx.x >> 5;
    const int wpb = blockDim.x >> 5;
    unsigned char *wsm = smem_raw + (size_t)wib * warp_smem_bytes<T, N, S>();
    const uint32_t sbuf0 = (uint32_t)__cvta_generic_to_shared(wsm);
    float2 *carry = reinterpret_cast<float2 *>(wsm + 2 * stage_bytes<N>()); /* [(S-1)*T] */

    const uint32_t n_steps = (p.n_syms + 31u) / 32u;
    const uint32_t total_warps = gridDim.x * wpb;

    for (uint32_t ch = blockIdx.x * wpb + wib; ch < p.n_ch; ch += total_warps) {
        unsigned char *stp = p.state + (size_t)ch * p.state_stride;
        ChanScalars *gsc = reinterpret_cast<ChanScalars *>(stp);
        uint32_t *grec_d = reinterpret_cast<uint32_t *>(stp + sizeof(ChanScalars));
        float *grec_e = reinterpret_cast<float *>(stp + sizeof(ChanScalars) + 32 * S * 4);
        float2 *gcarry = reinterpret_cast<float2 *>(stp + sizeof(ChanScalars) + 32 * S * 8);

        /* ---- restore carried state ---- */
        uint32_t pd[S];
        float pe[S];
#pragma unroll
        for (int i = 0; i < S; ++i) {
            pd[i] = grec_d[i * 32 + lane];
            pe[i] = grec_e[i * 32 + lane];
        }
        for (int i = lane; i < (S - 1) * T; i += 32) carry[i] = gcarry[i];
        ChanScalars sc;
        if (p.do_sm) sc = *gsc; /* uniform loads */
        __syncwarp();

        const int16_t *src = p.pcm + (size_t)ch * p.ch_stride;
        auto issue = [&](uint32_t step) {
            const uint32_t buf = sbuf0 + (step & 1u) * stage_bytes<N>();
            const uint32_t nv = min(32u, p.n_syms - step * 32u);
            const char *g = reinterpret_cast<const char *>(src + (size_t)step * 32u * N);
#pragma unroll
            for (int q = 0; q < CPS; ++q) {
                const uint32_t gi = q * 32 + lane; /* 16-byte chunk index in the step */
                const uint32_t sl = gi / CPS, c = gi % CPS;
                if (sl < nv) cp_async16(buf + sl * (2 * N) + ((c ^ (sl & 7u)) << 4), g + (size_t)gi * 16);
            }
            cp_async_commit();
        };
        if (n_steps) issue(0);

        for (uint32_t step = 0; step < n_steps; ++step) {
            if (step + 1 < n_steps) {
                issue(step + 1);
                cp_async_wait<1>();
            } else {
                cp_async_wait<0>();
            }
            __syncwarp();
            const uint32_t buf = sbuf0 + (step & 1u) * stage_bytes<N>();
            const int nvalid = (int)min(32u, p.n_syms - step * 32u);
            const bool active = lane < nvalid;
            const unsigned long long hbs = p.hop_base + (unsigned long long)step * 32u * S;

            /* ================= tone energies (SPEC 3) ================= */
            uint32_t dc[S];
            float ec[S];
#pragma unroll
            for (int i = 0; i < S; ++i) { dc[i] = 0xFFu; ec[i] = 0.0f; }

#pragma unroll 1
            for (int g = 0; g < NG; ++g) {
                float2 Pp[S][TG];
                if (active) {
#pragma unroll
                    for (int i = 0; i < S; ++i) {
                        float2 acc[TG];
#pragma unroll
                        for (int t = 0; t < TG; ++t) acc[t] = make_float2(0.0f, 0.0f);
#pragma unroll
                        for (int c = 0; c < H / 8; ++c) {
                            const int cc = i * (H / 8) + c;
                            const uint4 v = lds128(buf + lane * (2 * N) + (((uint32_t)cc ^ (lane & 7u)) << 4));
                            const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                            for (int q = 0; q < 4; ++q) {
                                const float x0 = (float)(short)(w[q] & 0xffffu);
                                const float x1 = (float)((int)w[q] >> 16);
                                const int m = cc * 8 + q * 2;
#pragma unroll
                                for (int t = 0; t < TG; ++t) {
                                    const float2 t0 = TWC ? c_tw[m * T + g * TG + t] : __ldg(&p.tw_global[m * T + g * TG + t]);
                                    acc[t] = ffma2(x0, t0, acc[t]);
                                }
#pragma unroll
                                for (int t = 0; t < TG; ++t) {
                                    const float2 t1 = TWC ? c_tw[(m + 1) * T + g * TG + t] : __ldg(&p.tw_global[(m + 1) * T + g * TG + t]);
                                    acc[t] = ffma2(x1, t1, acc[t]);
                                }
                            }
                        }
#pragma unroll
                        for (int t = 0; t < TG; ++t) Pp[i][t] = acc[t];
                    }
                } else {
#pragma unroll
                    for (int i = 0; i < S; ++i)
#pragma unroll
                        for (int t = 0; t < TG; ++t) Pp[i][t] = make_float2(0.0f, 0.0f);
                }
                /* window tree: lane 0 takes the previous step's tail from the carry buffer */
#pragma unroll
                for (int t = 0; t < TG; ++t) {
                    const int tg = g * TG + t;
                    float2 cin[S - 1];
#pragma unroll
                    for (int i = 0; i < S - 1; ++i) cin[i] = (lane == 0) ? carry[tg * (S - 1) + i] : make_float2(0.f, 0.f);
                    __syncwarp();
                    float2 L[S];
#pragma unroll
                    for (int i = 0; i < S; ++i) L[i] = Pp[i][t];
#pragma unroll
                    for (int lv = 1; lv <= LV; ++lv) {
                        const int d = 1 << (lv - 1);
                        float2 Nw[S];
#pragma unroll
                        for (int i = 0; i < S; ++i) {
                            float2 a;
                            if (i >= d) {
                                a = L[i - d];
                            } else {
                                const float2 tail = L[S - d + i];
                                a = shfl2(tail, (lane + 31) & 31);
                                if (lane == 0) a = cin[d - 1 + i];
                                if (lane == nvalid - 1) carry[tg * (S - 1) + d - 1 + i] = tail;
                            }
                            Nw[i] = fadd2(a, L[i]);
                        }
#pragma unroll
                        for (int i = 0; i < S; ++i) L[i] = Nw[i];
                    }
#pragma unroll
                    for (int i = 0; i < S; ++i) {
                        const float E = __fmaf_rn(L[i].x, L[i].x, __fmul_rn(L[i].y, L[i].y));
                        if (p.trE && active) {
                            const size_t hop = ((size_t)step * 32 + lane) * S + i;
                            p.trE[((size_t)ch * p.tr_hops + hop) * T + tg] = E;
                        }
                        if (tg == 0 || E > ec[i]) { ec[i] = E; dc[i] = (uint32_t)tg; }
                    }
                }
            }
            if (!active) {
#pragma unroll
                for (int i = 0; i < S; ++i) { dc[i] = 0xFFu; ec[i] = 0.0f; }
            }
            if (p.trD && active) {
#pragma unroll
                for (int i = 0; i < S; ++i) p.trD[(size_t)ch * p.tr_hops + ((size_t)step * 32 + lane) * S + i] = (uint8_t)dc[i];
            }
            if (p.trEmax && active) {
#pragma unroll
                for (int i = 0; i < S; ++i) p.trEmax[(size_t)ch * p.tr_hops + ((size_t)step * 32 + lane) * S + i] = ec[i];
            }

            /* ================= sync / slicing / framing (SPEC 5) ================= */
            if (p.do_sm) {
                const int endh = nvalid * S;
                int cur = 0;
                bool have_cand = false;
                uint32_t cand[S];
#pragma unroll
                for (int i = 0; i < S; ++i) cand[i] = 0;

                /* quality of the alignment ending at (slot sh, phase ph): SPEC 5 q(h) */
                auto quality = [&](int sh, int ph) -> float {
                    const int ss = sh - (int)(p.P - 1) + lane; /* slot of preamble symbol `lane` */
                    const uint32_t dcur_ = pick<S>(dc, ph), dprev_ = pick<S>(pd, ph);
                    const float ecur_ = pick<S>(ec, ph), eprev_ = pick<S>(pe, ph);
                    const uint32_t d1 = __shfl_sync(FULL, dcur_, ss & 31), d0 = __shfl_sync(FULL, dprev_, ss & 31);
                    const float e1 = __shfl_sync(FULL, ecur_, ss & 31), e0 = __shfl_sync(FULL, eprev_, ss & 31);
                    const uint32_t dv = ss >= 0 ? d1 : d0;
                    const float ev = ss >= 0 ? e1 : e0;
                    float leaf = (lane < (int)p.P && dv == (uint32_t)p.preamble[lane & 31]) ? ev : 0.0f;
                    for (uint32_t w = 1; w < p.P; w <<= 1) leaf = __fadd_rn(leaf, __shfl_xor_sync(FULL, leaf, w));
                    return __shfl_sync(FULL, leaf, 0);
                };

                while (cur < endh) {
                    if (sc.state == ST_SEARCH || sc.state == ST_PEAK) {
                        if (!have_cand) {
                            /* preamble correlation on bit-planes of the hop decisions */
#pragma unroll
                            for (int i = 0; i < S; ++i) {
                                uint32_t mism = 0;
                                const int sh = 32 + lane - (int)(p.P - 1); /* bit of preamble symbol 0 in the 64-bit history */
#pragma unroll
                                for (int j = 0; j <= B; ++j) {
                                    const uint32_t bc = (j < B) ? ((dc[i] >> j) & 1u) : (dc[i] > (uint32_t)(T - 1));
                                    const uint32_t bp = (j < B) ? ((pd[i] >> j) & 1u) : (pd[i] > (uint32_t)(T - 1));
                                    const unsigned long long hist = ((unsigned long long)__ballot_sync(FULL, bc) << 32) | __ballot_sync(FULL, bp);
                                    const uint32_t w = (uint32_t)(hist >> sh);
                                    mism |= (j < B) ? (w ^ p.pre_plane[j]) : w;
                                }
                                const uint32_t pmask = (p.P >= 32) ? 0xffffffffu : ((1u << p.P) - 1u);
                                const uint32_t m = p.P - __popc(mism & pmask);
                                cand[i] = __ballot_sync(FULL, active && m >= p.P - p.tol);
                            }
                            have_cand = true;
                        }
                        if (sc.state == ST_SEARCH) {
                            int h0 = 0x7fffffff;
#pragma unroll
                            for (int i = 0; i < S; ++i) {
                                const int smin = (cur > i) ? (cur - i + S - 1) / S : 0;
                                const uint32_t mk = (smin >= 32) ? 0u : (cand[i] & (0xffffffffu << smin));
                                if (mk) h0 = min(h0, (__ffs(mk) - 1) * S + i);
                            }
                            if (h0 == 0x7fffffff) { cur = endh; break; }
                            sc.best_q = quality(h0 / S, h0 % S);
                            sc.best_h = hbs + h0;
                            sc.peak_end = hbs + h0 + S - 1;
                            sc.state = ST_PEAK;
                            cur = h0 + 1;
                        } else {
                            const long long pend = (long long)(sc.peak_end - hbs);
                            while (cur < endh && cur <= pend) {
                                const int sl = cur / S, ph = cur % S;
                                if ((pick<S>(cand, ph) >> sl) & 1u) {
                                    const float q = quality(sl, ph);
                                    if (q > sc.best_q) { sc.best_q = q; sc.best_h = hbs + cur; }
                                }
                                ++cur;
                            }
                            if (cur > pend) {
                                sc.t0 = sc.best_h;
                                sc.next = sc.t0 + S;
                                sc.nsym = 0;
                                sc.acc = 0;
                                sc.s_prev = p.preamble[p.P - 1];
                                sc.s_prev2 = 0xFFu;
                                sc.prev_hop = sc.t0;
                                sc.state = ST_HEADER;
                                sc.stats.locks++;
                            }
                        }
                    } else {
                        /* ---- locked: slice up to 32 symbols at once ---- */
                        const long long firstl = (long long)(sc.next - hbs);
                        if (firstl >= endh) { cur = endh; break; }
                        const int first = (int)firstl;
                        const uint32_t until_epoch = p.trk_epoch - (sc.nsym % p.trk_epoch);
                        const uint32_t until_evt = (sc.state == ST_HEADER ? p.hdr_syms : sc.total) - sc.nsym;
                        uint32_t cnt = min(until_epoch, until_evt);
                        cnt = min(cnt, (uint32_t)((endh - 1 - first) / S + 1));
                        const int phi = first % S, s0 = first / S;
                        const int e = lane - s0;
                        const bool part = e >= 0 && e < (int)cnt;
                        const uint32_t sym = pick<S>(dc, phi);
                        uint8_t *fs = p.fsyms + (size_t)ch * p.fsym_stride;
                        if (part) {
                            fs[sc.nsym + e] = (uint8_t)sym;
                            if (p.osyms) {
                                const uint32_t oi = sc.osym_cnt + e;
                                if (oi < p.osym_cap) p.osyms[(size_t)ch * p.osym_cap + oi] = (uint8_t)sym;
                            }
                        }
                        /* tracker votes (SPEC 5): lane e votes for symbol nsym+e-1 */
                        {
                            /* records of the previous symbol for lanes e >= 1: slot-1 */
                            const int phe = (phi >= 1) ? phi - 1 : S - 1, ke = (phi >= 1) ? 1 : 2;
                            const int phl = (phi + 1 < S) ? phi + 1 : 0, kl = (phi + 1 < S) ? 1 : 0;
                            auto fetch = [&](int k, int ph, uint32_t &dv, float &ev) {
                                const uint32_t dcur_ = pick<S>(dc, ph), dprev_ = pick<S>(pd, ph);
                                const float ecur_ = pick<S>(ec, ph), eprev_ = pick<S>(pe, ph);
                                const uint32_t ds = (lane >= 32 - k) ? dprev_ : dcur_;
                                const float es = (lane >= 32 - k) ? eprev_ : ecur_;
                                dv = __shfl_sync(FULL, ds, (lane - k) & 31);
                                ev = __shfl_sync(FULL, es, (lane - k) & 31);
                            };
                            uint32_t d_on, d_ea, d_la;
                            float e_on, e_ea, e_la;
                            fetch(1, phi, d_on, e_on);
                            fetch(ke, phe, d_ea, e_ea);
                            fetch(kl, phl, d_la, e_la);
                            uint32_t sj = __shfl_up_sync(FULL, sym, 1);
                            uint32_t sjm = __shfl_up_sync(FULL, sym, 2);
                            if (e == 1) sjm = sc.s_prev;
                            /* lane e == 0: previous symbol is at the carried hop prev_hop */
                            {
                                const int r = (int)((long long)(sc.prev_hop - hbs));
                                uint32_t dd[3];
                                float ee[3];
#pragma unroll
                                for (int z = 0; z < 3; ++z) {
                                    const int rr = r - 1 + z;
                                    const int sl = floordiv(rr, S), ph = rr - sl * S;
                                    const uint32_t dsel = sl < 0 ? pick<S>(pd, ph) : pick<S>(dc, ph);
                                    const float esel = sl < 0 ? pick<S>(pe, ph) : pick<S>(ec, ph);
                                    dd[z] = __shfl_sync(FULL, dsel, sl & 31);
                                    ee[z] = __shfl_sync(FULL, esel, sl & 31);
                                }
                                if (e == 0) {
                                    d_ea = dd[0]; e_ea = ee[0];
                                    d_on = dd[1]; e_on = ee[1];
                                    d_la = dd[2]; e_la = ee[2];
                                    sj = sc.s_prev;
                                    sjm = sc.s_prev2;
                                }
                            }
                            (void)d_on;
                            const bool voter = part && (sc.nsym + e >= 1);
                            const float ve = (d_ea == sj) ? e_ea : 0.0f;
                            const float vl = (d_la == sj) ? e_la : 0.0f;
                            const uint32_t bl = __ballot_sync(FULL, voter && (sym != sj) && vl > e_on);
                            const uint32_t be = __ballot_sync(FULL, voter && (sjm != sj) && ve > e_on);
                            sc.acc += __popc(bl) - __popc(be);
                        }
                        const int last = s0 + (int)cnt - 1;
                        const uint32_t ns1 = __shfl_sync(FULL, sym, last);
                        const uint32_t ns2 = __shfl_sync(FULL, sym, (last - 1) & 31);
                        sc.s_prev2 = (cnt >= 2) ? ns2 : sc.s_prev;
                        sc.s_prev = ns1;
                        sc.prev_hop = hbs + first + (cnt - 1) * S;
                        sc.nsym += cnt;
                        sc.osym_cnt += cnt;
                        sc.next += (unsigned long long)cnt * S;
                        sc.stats.symbols += cnt;
                        cur = first + (int)(cnt - 1) * S + 1;
                        if (sc.nsym % p.trk_epoch == 0) {
                            if (sc.acc >= (int)p.trk_thresh) { sc.next += 1; sc.stats.trk_moves++; }
                            else if (sc.acc <= -(int)p.trk_thresh) { sc.next -= 1; sc.stats.trk_moves--; }
                            sc.acc = 0;
                        }
                        if (sc.state == ST_HEADER && sc.nsym == p.hdr_syms) {
                            __syncwarp();
                            uint32_t hdr = 0;
#pragma unroll
                            for (int bit = 0; bit < 24; ++bit) {
                                const uint32_t v = gray_inv(fs[bit / B]);
                                hdr = (hdr << 1) | ((v >> (B - 1 - (bit % B))) & 1u);
                            }
                            const uint32_t len = hdr >> 8;
                            uint32_t c8 = 0;
#pragma unroll
                            for (int z = 0; z < 2; ++z) {
                                c8 ^= (hdr >> (16 - 8 * z)) & 0xffu;
#pragma unroll
                                for (int k = 0; k < 8; ++k) c8 = (c8 & 0x80u) ? (((c8 << 1) ^ 0x07u) & 0xffu) : ((c8 << 1) & 0xffu);
                            }
                            if (len == 0 || len > p.max_payload || c8 != (hdr & 0xffu)) {
                                sc.stats.header_fail++;
                                sc.state = ST_SEARCH;
                            } else {
                                sc.flen = len;
                                sc.total = p.hdr_syms + ((len + 2) * 8 + B - 1) / B;
                                sc.state = ST_BODY;
                            }
                        } else if (sc.state == ST_BODY && sc.nsym == sc.total) {
                            __syncwarp();
                            const uint32_t len = sc.flen;
                            uint32_t fidx = 0, boff = 0;
                            if (lane == 0) {
                                fidx = atomicAdd(&p.counters[0], 1u);
                                boff = atomicAdd(&p.counters[1], len);
                            }
                            fidx = __shfl_sync(FULL, fidx, 0);
                            boff = __shfl_sync(FULL, boff, 0);
                            const bool fits = fidx < p.frames_cap && boff + len <= p.bytes_cap;
                            /* bits -> bytes, lane-parallel over body bytes (payload + CRC16) */
                            const uint8_t *bs = fs + p.hdr_syms;
                            uint32_t crc_rx = 0;
                            for (uint32_t byi = lane; byi < len + 2; byi += 32) {
                                uint32_t v8 = 0;
#pragma unroll
                                for (int k = 0; k < 8; ++k) {
                                    const uint32_t bit = byi * 8 + k;
                                    const uint32_t v = gray_inv(bs[bit / B]);
                                    v8 = (v8 << 1) | ((v >> (B - 1 - (bit % B))) & 1u);
                                }
                                if (byi < len) { if (fits) p.bytes[boff + byi] = (uint8_t)v8; }
                                else crc_rx |= v8 << (8 * (len + 1 - byi));
                            }
                            crc_rx |= __shfl_xor_sync(FULL, crc_rx, 16);
                            crc_rx |= __shfl_xor_sync(FULL, crc_rx, 8);
                            crc_rx |= __shfl_xor_sync(FULL, crc_rx, 4);
                            crc_rx |= __shfl_xor_sync(FULL, crc_rx, 2);
                            crc_rx |= __shfl_xor_sync(FULL, crc_rx, 1);
                            __syncwarp();
                            uint32_t ok = 0;
                            if (fits) {
                                uint32_t crc = 0xFFFFu;
                                if (lane == 0) {
                                    crc = ((crc << 8) ^ s_crc[((crc >> 8) ^ (len >> 8)) & 0xffu]) & 0xffffu;
                                    crc = ((crc << 8) ^ s_crc[((crc >> 8) ^ (len & 0xffu)) & 0xffu]) & 0xffffu;
                                    for (uint32_t z = 0; z < len; ++z)
                                        crc = ((crc << 8) ^ s_crc[((crc >> 8) ^ p.bytes[boff + z]) & 0xffu]) & 0xffffu;
                                    ok = crc == crc_rx;
                                    anm_frame_t f;
                                    f.channel = ch;
                                    f.len = len;
                                    f.start_sample = (sc.t0 + 1 - (unsigned long long)p.P * S) * H;
                                    f.crc_ok = ok;
                                    f.offset = boff;
                                    p.frames[fidx] = f;
                                }
                                ok = __shfl_sync(FULL, ok, 0);
                                if (ok) sc.stats.frames_ok++; else sc.stats.frames_bad++;
                            } else if (lane == 0) {
                                atomicOr(&p.counters[2], 1u);
                            }
                            sc.state = ST_SEARCH;
                        }
                    }
                }
            }

            /* ---- this step's records become the history of the next ---- */
            if (nvalid == 32) {
#pragma unroll
                for (int i = 0; i < S; ++i) { pd[i] = dc[i]; pe[i] = ec[i]; }
            } else {
#pragma unroll
                for (int i = 0; i < S; ++i) {
                    const uint32_t a = __shfl_sync(FULL, pd[i], (lane + nvalid) & 31), b = __shfl_sync(FULL, dc[i], (lane + nvalid) & 31);
                    const float fa = __shfl_sync(FULL, pe[i], (lane + nvalid) & 31), fb = __shfl_sync(FULL, ec[i], (lane + nvalid) & 31);
                    pd[i] = (lane < 32 - nvalid) ? a : b;
                    pe[i] = (lane < 32 - nvalid) ? fa : fb;
                }
            }
            __syncwarp();
        }

        /* ---- save carried state ---- */
#pragma unroll
        for (int i = 0; i < S; ++i) {
            grec_d[i * 32 + lane] = pd[i];
            grec_e[i * 32 + lane] = pe[i];
        }
        for (int i = lane; i < (S - 1) * T; i += 32) gcarry[i] = carry[i];
        if (p.do_sm && lane == 0) *gsc = sc;
        __syncwarp();
    }
}

} /* namespace anm */
