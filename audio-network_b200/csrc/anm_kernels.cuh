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
    uint32_t ep_left;       /* symbols until the next tracker epoch boundary */
    uint32_t pad[3];
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
    uint32_t tw_sign;             /* bit k: tone_bin[k] is odd, i.e. twiddle[m+N/2][k] = -twiddle[m][k] (first 32 tones) */
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
    unsigned long long tw_sign_hi; /* same for tones 32..63 */
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

/* MODE 0: streaming demodulator (sync, slicing, framing; no trace output).
 * MODE 1: stateless tone-energy pass (trace outputs only; parity / debug). */
template <int T, int N, int S, bool TWC, int MODE>
__global__ void __launch_bounds__(512) k_demod(const __grid_constant__ KParams p) {
    constexpr int H = N / S;
    constexpr int B = Log2<T>::v;
    constexpr int TG = T < 8 ? T : 8; /* tones per register group */
    constexpr int NG = T / TG;
    constexpr int LV = Log2<S>::v;
    constexpr int HS = S / 2;  /* hop-pair iterations: hops i and i+S/2 share twiddles up to sign */
    constexpr int CPH = H / 8; /* 16-byte chunks per hop */
    constexpr int CPS = N / 8; /* 16-byte chunks per symbol period */
    constexpr uint32_t FULL = 0xffffffffu;
    static_assert(N >= 64 && (H % 8) == 0 && S >= 2, "unsupported geometry");

    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ uint16_t s_crc[256];
    if (MODE == 0) {
        for (int i = threadIdx.x; i < 256; i += blockDim.x) {
            uint32_t c = (uint32_t)i << 8;
            for (int k = 0; k < 8; ++k) c = (c & 0x8000u) ? ((c << 1) ^ 0x1021u) : (c << 1);
            s_crc[i] = (uint16_t)c;
        }
        __syncthreads();
    }

    const int lane = threadIdx.x & 31;
    const int wib = threadIdx.x >> 5;
    const int wpb = blockDim.x >> 5;
    unsigned char *wsm = smem_raw + (size_t)wib * warp_smem_bytes<T, N, S>();
    const uint32_t sbuf0 = (uint32_t)__cvta_generic_to_shared(wsm);
    float2 *carry = reinterpret_cast<float2 *>(wsm + 2 * stage_bytes<N>()); /* [(S-1)*T] */

    /* lane-constant pieces of the swizzled addresses */
    const uint32_t sw = lane & 7u;
    constexpr uint32_t LOWM = (uint32_t)(CPH - 1) & 7u, HIM = 7u & ~LOWM;
    uint32_t swc[CPH];
#pragma unroll
    for (int c = 0; c < CPH; ++c) swc[c] = (((uint32_t)c ^ (sw & LOWM))) << 4;
    const uint32_t swhi = sw & HIM;
    /* cp.async: lane copies 16-byte chunk (q*32 + lane) of the step; destination slot and
     * chunk-in-slot are lane constants up to a per-q constant */
    constexpr int LPS = (CPS >= 32) ? 1 : 32 / CPS; /* symbol slots covered by one cp.async instruction */
    const uint32_t cp_slot = (CPS >= 32) ? 0u : (uint32_t)lane / (uint32_t)CPS;
    const uint32_t cp_chunk = (uint32_t)lane % (uint32_t)CPS;

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
        if (MODE == 0) sc = *gsc; /* uniform loads */
        __syncwarp();

        const char *src = reinterpret_cast<const char *>(p.pcm + (size_t)ch * p.ch_stride);
        auto issue = [&](uint32_t step) {
            const uint32_t buf = sbuf0 + (step & 1u) * stage_bytes<N>();
            const uint32_t nv = min(32u, p.n_syms - step * 32u);
            const char *g = src + (size_t)step * (32u * N * 2u) + (size_t)lane * 16u;
            if (CPS >= 32) {
                /* one instruction covers 512 bytes of one slot (N >= 256) */
                constexpr uint32_t IPS = (CPS >= 32) ? CPS / 32 : 1; /* instructions per slot */
#pragma unroll 4
                for (uint32_t q = 0; q < (uint32_t)CPS; ++q) {
                    const uint32_t sl = q / IPS, c = (q % IPS) * 32u + lane;
                    if (sl < nv) cp_async16(buf + sl * (2 * N) + ((c ^ (sl & 7u)) << 4), g + (size_t)q * 512u);
                }
            } else if (nv == 32u) {
#pragma unroll
                for (int q = 0; q < CPS; ++q) {
                    /* slot = q*LPS + cp_slot; (slot & 7) = ((q*LPS) & 7) ^ cp_slot since LPS is a power of two > cp_slot */
                    const uint32_t sl7 = ((uint32_t)(q * LPS) & 7u) | (cp_slot & 7u);
                    cp_async16(buf + (uint32_t)(q * LPS) * (2 * N) + cp_slot * (2 * N) + ((cp_chunk ^ sl7) << 4), g + (size_t)q * 512u);
                }
            } else {
#pragma unroll 4
                for (int q = 0; q < CPS; ++q) {
                    const uint32_t sl = (uint32_t)(q * LPS) + cp_slot;
                    if (sl < nv) cp_async16(buf + sl * (2 * N) + ((cp_chunk ^ (sl & 7u)) << 4), g + (size_t)q * 512u);
                }
            }
            cp_async_commit();
        };
        if (n_steps) issue(0);

#pragma unroll 1
        for (uint32_t step = 0; step < n_steps; ++step) {
            if (step + 1 < n_steps) {
                issue(step + 1);
                cp_async_wait<1>();
            } else {
                cp_async_wait<0>();
            }
            __syncwarp();
            const uint32_t row = sbuf0 + (step & 1u) * stage_bytes<N>() + (uint32_t)lane * (2 * N);
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
                float2 PA[HS][TG], PB[HS][TG]; /* PA[i] = P[hop i], PB[i] = P[hop i + S/2] */
#pragma unroll
                for (int i = 0; i < HS; ++i)
#pragma unroll
                    for (int t = 0; t < TG; ++t) { PA[i][t] = make_float2(0.f, 0.f); PB[i][t] = make_float2(0.f, 0.f); }
                /* all 32 lanes run the arithmetic (warp-uniform control flow keeps the twiddle loads on
                 * the uniform datapath); lanes beyond a ragged chunk end read stale shared memory and
                 * their results are discarded below */
                {
                    /* fully unrolled: static twiddle offsets become LDCU.128 with immediate addresses */
#pragma unroll
                    for (int it = 0; it < HS; ++it) {
                        float2 aA[TG], aB[TG];
#pragma unroll
                        for (int t = 0; t < TG; ++t) { aA[t] = make_float2(0.f, 0.f); aB[t] = make_float2(0.f, 0.f); }
                        const uint32_t baseA = (uint32_t)it * CPH, baseB = baseA + (uint32_t)HS * CPH;
                        const uint32_t hiA = row + ((baseA ^ swhi) << 4), hiB = row + ((baseB ^ swhi) << 4);
                        const int twb = it * (H * T) + g * TG; /* uniform twiddle base of this hop */
#pragma unroll
                        for (int c = 0; c < CPH; ++c) {
                            const uint4 vA = lds128(hiA + swc[c]);
                            const uint4 vB = lds128(hiB + swc[c]);
                            const uint32_t wA[4] = {vA.x, vA.y, vA.z, vA.w};
                            const uint32_t wB[4] = {vB.x, vB.y, vB.z, vB.w};
#pragma unroll
                            for (int q = 0; q < 4; ++q) {
                                const float xA0 = (float)(short)(wA[q] & 0xffffu), xA1 = (float)((int)wA[q] >> 16);
                                const float xB0 = (float)(short)(wB[q] & 0xffffu), xB1 = (float)((int)wB[q] >> 16);
                                const int m = (c * 8 + q * 2) * T;
#pragma unroll
                                for (int t = 0; t < TG; ++t) {
                                    const float2 w0 = TWC ? c_tw[twb + m + t] : __ldg(&p.tw_global[twb + m + t]);
                                    aA[t] = ffma2(xA0, w0, aA[t]);
                                    aB[t] = ffma2(xB0, w0, aB[t]);
                                }
#pragma unroll
                                for (int t = 0; t < TG; ++t) {
                                    const float2 w1 = TWC ? c_tw[twb + m + T + t] : __ldg(&p.tw_global[twb + m + T + t]);
                                    aA[t] = ffma2(xA1, w1, aA[t]);
                                    aB[t] = ffma2(xB1, w1, aB[t]);
                                }
                            }
                        }
                        /* twiddle[m + N/2][k] = -twiddle[m][k] for odd bins: negate exactly */
                        const unsigned long long sgn = ((unsigned long long)p.tw_sign | (p.tw_sign_hi << 32)) >> (g * TG);
#pragma unroll
                        for (int t = 0; t < TG; ++t) {
                            const uint32_t mk = ((uint32_t)(sgn >> t) & 1u) << 31;
                            aB[t].x = __uint_as_float(__float_as_uint(aB[t].x) ^ mk);
                            aB[t].y = __uint_as_float(__float_as_uint(aB[t].y) ^ mk);
                        }
#pragma unroll
                        for (int j = 0; j + 1 < HS; ++j)
#pragma unroll
                            for (int t = 0; t < TG; ++t) { PA[j][t] = PA[j + 1][t]; PB[j][t] = PB[j + 1][t]; }
#pragma unroll
                        for (int t = 0; t < TG; ++t) { PA[HS - 1][t] = aA[t]; PB[HS - 1][t] = aB[t]; }
                    }
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
                    for (int i = 0; i < S; ++i) L[i] = (i < HS) ? PA[i % HS][t] : PB[i % HS][t];
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
                        if (MODE == 1) {
                            if (p.trE && active) {
                                const size_t hop = ((size_t)step * 32 + lane) * S + i;
                                p.trE[((size_t)ch * p.tr_hops + hop) * T + tg] = E;
                            }
                        }
                        if (tg == 0 || E > ec[i]) { ec[i] = E; dc[i] = (uint32_t)tg; }
                    }
                }
            }
            if (!active) {
#pragma unroll
                for (int i = 0; i < S; ++i) { dc[i] = 0xFFu; ec[i] = 0.0f; }
            }
            if (MODE == 1) {
                if (p.trD && active) {
#pragma unroll
                    for (int i = 0; i < S; ++i) p.trD[(size_t)ch * p.tr_hops + ((size_t)step * 32 + lane) * S + i] = (uint8_t)dc[i];
                }
                if (p.trEmax && active) {
#pragma unroll
                    for (int i = 0; i < S; ++i) p.trEmax[(size_t)ch * p.tr_hops + ((size_t)step * 32 + lane) * S + i] = ec[i];
                }
            }

            /* ================= sync / slicing / framing (SPEC 5) ================= */
            if (MODE == 0) {
                const int endh = nvalid * S;
                int cur = 0;
                bool have_cand = false;
                uint32_t cand[S];
#pragma unroll
                for (int i = 0; i < S; ++i) cand[i] = 0;

                /* record (d, emax) of relative hop rr (warp-uniform, may lie in the previous step) */
                auto rec_at = [&](int rr, uint32_t &dv, float &ev) {
                    const int sl = rr >> LV, ph = rr & (S - 1);
                    const uint32_t dsel = sl < 0 ? pick<S>(pd, ph) : pick<S>(dc, ph);
                    const float esel = sl < 0 ? pick<S>(pe, ph) : pick<S>(ec, ph);
                    dv = __shfl_sync(FULL, dsel, sl & 31);
                    ev = __shfl_sync(FULL, esel, sl & 31);
                };
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

#pragma unroll 1
                while (cur < endh) {
                    if (sc.state == ST_SEARCH || sc.state == ST_PEAK) {
                        if (!have_cand) {
                            /* preamble correlation on bit-planes of the hop decisions */
                            const int sh = 32 + lane - (int)(p.P - 1); /* bit of preamble symbol 0 in the 64-bit history */
                            const uint32_t pmask = (p.P >= 32) ? 0xffffffffu : ((1u << p.P) - 1u);
#pragma unroll
                            for (int i = 0; i < S; ++i) {
                                uint32_t mism = 0;
#pragma unroll
                                for (int j = 0; j <= B; ++j) {
                                    const uint32_t bc = (j < B) ? ((dc[i] >> j) & 1u) : (dc[i] > (uint32_t)(T - 1));
                                    const uint32_t bp = (j < B) ? ((pd[i] >> j) & 1u) : (pd[i] > (uint32_t)(T - 1));
                                    const unsigned long long hist = ((unsigned long long)__ballot_sync(FULL, bc) << 32) | __ballot_sync(FULL, bp);
                                    const uint32_t w = (uint32_t)(hist >> sh);
                                    mism |= (j < B) ? (w ^ p.pre_plane[j]) : w;
                                }
                                const uint32_t m = p.P - __popc(mism & pmask);
                                cand[i] = __ballot_sync(FULL, active && m >= p.P - p.tol);
                            }
                            have_cand = true;
                        }
                        if (sc.state == ST_SEARCH) {
                            int h0 = 0x7fffffff;
#pragma unroll
                            for (int i = 0; i < S; ++i) {
                                const int smin = (cur > i) ? ((cur - i + S - 1) >> LV) : 0;
                                const uint32_t mk = (smin >= 32) ? 0u : (cand[i] & (0xffffffffu << smin));
                                if (mk) h0 = min(h0, ((__ffs(mk) - 1) << LV) + i);
                            }
                            if (h0 == 0x7fffffff) { cur = endh; break; }
                            sc.best_q = quality(h0 >> LV, h0 & (S - 1));
                            sc.best_h = hbs + h0;
                            sc.peak_end = hbs + h0 + S - 1;
                            sc.state = ST_PEAK;
                            cur = h0 + 1;
                        } else {
                            const long long pend = (long long)(sc.peak_end - hbs);
                            while (cur < endh && cur <= pend) {
                                const int sl = cur >> LV, ph = cur & (S - 1);
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
                                sc.ep_left = p.trk_epoch;
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
                        const uint32_t until_evt = (sc.state == ST_HEADER ? p.hdr_syms : sc.total) - sc.nsym;
                        uint32_t cnt = min(until_evt, (uint32_t)(((endh - 1 - first) >> LV) + 1));
                        const int phi = first & (S - 1), s0 = first >> LV;
                        const int e = lane - s0;
                        const uint32_t sym = pick<S>(dc, phi);
                        uint8_t *fs = p.fsyms + (size_t)ch * p.fsym_stride;
                        if (e >= 0 && e < (int)cnt) {
                            fs[sc.nsym + e] = (uint8_t)sym;
                            if (p.osyms) {
                                const uint32_t oi = sc.osym_cnt + e;
                                if (oi < p.osym_cap) p.osyms[(size_t)ch * p.osym_cap + oi] = (uint8_t)sym;
                            }
                        }
                        /* tracker votes (SPEC 5): lane e votes for symbol nsym+e-1.  For e >= 1 that
                         * symbol sits one slot back at the same phase; for e == 0 at the carried hop. */
                        uint32_t bl, be;
                        {
                            const int r0 = (int)((long long)(sc.prev_hop - hbs));
                            uint32_t d0[3];
                            float e0[3];
#pragma unroll
                            for (int z = 0; z < 3; ++z) rec_at(r0 - 1 + z, d0[z], e0[z]);
                            /* same-phase neighbours, fetched with one rotate-by-one each */
                            const int phe = (phi >= 1) ? phi - 1 : S - 1, phl = (phi + 1 < S) ? phi + 1 : 0;
                            const uint32_t von_d = pick<S>(dc, phi);
                            const float von_e = pick<S>(ec, phi);
                            (void)von_d;
                            const uint32_t ea_d = pick<S>(dc, phe), la_d = pick<S>(dc, phl);
                            const float ea_e = pick<S>(ec, phe), la_e = pick<S>(ec, phl);
                            /* e >= 1 implies lane >= 1 (and >= 2 when two slots back are needed only if e >= 1 and
                             * phi == 0, where slot-2 >= s0-1 >= -1 can be the previous step: use the carried copy) */
                            float e_on = __shfl_up_sync(FULL, von_e, 1);
                            uint32_t d_ea;
                            float e_ea;
                            if (phi >= 1) {
                                d_ea = __shfl_up_sync(FULL, ea_d, 1);
                                e_ea = __shfl_up_sync(FULL, ea_e, 1);
                            } else {
                                const uint32_t pd_l = pick<S>(pd, S - 1);
                                const float pe_l = pick<S>(pe, S - 1);
                                const uint32_t dsrc = (lane >= 30) ? pd_l : ea_d;
                                const float esrc = (lane >= 30) ? pe_l : ea_e;
                                d_ea = __shfl_sync(FULL, dsrc, (lane - 2) & 31);
                                e_ea = __shfl_sync(FULL, esrc, (lane - 2) & 31);
                            }
                            uint32_t d_la = (phi + 1 < S) ? __shfl_up_sync(FULL, la_d, 1) : la_d;
                            float e_la = (phi + 1 < S) ? __shfl_up_sync(FULL, la_e, 1) : la_e;
                            uint32_t sj = __shfl_up_sync(FULL, sym, 1);
                            uint32_t sjm = __shfl_up_sync(FULL, sym, 2);
                            if (e == 1) sjm = sc.s_prev;
                            if (e == 0) {
                                d_ea = d0[0]; e_ea = e0[0];
                                e_on = e0[1];
                                d_la = d0[2]; e_la = e0[2];
                                sj = sc.s_prev;
                                sjm = sc.s_prev2;
                            }
                            const bool voter = e >= 0 && e < (int)cnt && (sc.nsym + e >= 1);
                            const float ve = (d_ea == sj) ? e_ea : 0.0f;
                            const float vl = (d_la == sj) ? e_la : 0.0f;
                            bl = __ballot_sync(FULL, voter && (sym != sj) && vl > e_on);
                            be = __ballot_sync(FULL, voter && (sjm != sj) && ve > e_on);
                        }
                        /* tracker epochs inside the run: only an actual timing move ends the run early */
                        int adj = 0;
                        {
                            uint32_t pos = 0;
                            while (true) {
                                const uint32_t eb = pos + sc.ep_left; /* symbols of the run up to the next boundary */
                                const uint32_t hi = min(eb, cnt);
                                /* lanes [s0+pos, s0+hi) */
                                const uint32_t lo_m = 0xffffffffu << (s0 + pos);
                                const uint32_t hi_m = (s0 + hi >= 32u) ? 0xffffffffu : ((1u << (s0 + hi)) - 1u);
                                sc.acc += __popc(bl & lo_m & hi_m) - __popc(be & lo_m & hi_m);
                                if (eb > cnt) { sc.ep_left -= (cnt - pos); break; }
                                pos = eb;
                                sc.ep_left = p.trk_epoch;
                                if (sc.acc >= (int)p.trk_thresh) adj = 1;
                                else if (sc.acc <= -(int)p.trk_thresh) adj = -1;
                                sc.acc = 0;
                                if (adj) { cnt = pos; break; }
                                if (pos == cnt) break;
                            }
                        }
                        const int last = s0 + (int)cnt - 1;
                        const uint32_t ns1 = __shfl_sync(FULL, sym, last);
                        const uint32_t ns2 = __shfl_sync(FULL, sym, (last - 1) & 31);
                        sc.s_prev2 = (cnt >= 2) ? ns2 : sc.s_prev;
                        sc.s_prev = ns1;
                        sc.prev_hop = hbs + first + ((cnt - 1) << LV);
                        sc.nsym += cnt;
                        sc.osym_cnt += cnt;
                        sc.next += ((unsigned long long)cnt << LV) + adj;
                        sc.stats.symbols += cnt;
                        sc.stats.trk_moves += adj;
                        cur = first + (int)((cnt - 1) << LV) + 1;
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
#pragma unroll 1
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
#pragma unroll 1
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
        if (MODE == 0 && lane == 0) *gsc = sc;
        __syncwarp();
    }
}

} /* namespace anm */
