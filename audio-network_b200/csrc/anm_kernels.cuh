/*
 * anm_kernels.cuh -- sm_100a kernels of the SPEC.md receive path.
 *
 * One warp owns one channel.  Within a step of 32 symbol periods, lane l owns symbol
 * period l (N samples = S hops), so all per-sample work is lane-private register
 * arithmetic (SPEC 3: centre-folded hop partials, one packed fma.rn.f32x2 per tone and
 * sample pair; the direct FMA chains for tone sets that do not fold), the tails of the
 * S-hop window tree travel to the next lane through the dead PCM stage, the preamble
 * correlation is ballots + popc over bit-planes of the hop decisions, and symbol
 * slicing / tracking / framing (sm_step, shared with the tensor-core kernel of
 * anm_kernels_tc.cuh) are lane-parallel over up to 32 symbols at a time.  PCM moves
 * HBM -> shared memory with coalesced 16-byte cp.async into a stage of padded rows
 * (refilled for the next step as soon as the arithmetic of this step has consumed it)
 * and is read back conflict-free.
 *
 * There is no reference kernel for any of this; the behaviour is
 * SPEC.md's, the structure is B200-first.
 */
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "anm_internal.h"

/* which of the four "before the centre" samples of an 8-sample group are converted on the ALU pipe (PRMT + I2FP, two issue
 * slots) instead of the XU pipe (I2F.S16, one issue slot at a quarter of the rate); the "after" samples always take XU */
/* 1: the folded loop of a single tone group reads the PCM stage with 16-byte loads (LDS.128, conflict-free with the stage's
 * odd-chunk row pitch) instead of 8-byte ones (two lanes of every half-warp on the same bank) */
#ifndef ANM_LDS128
#define ANM_LDS128 1
#endif
#ifndef ANM_CVT_ALU_MASK
#define ANM_CVT_ALU_MASK 0xF
#endif

namespace anm {

enum : uint32_t { ST_SEARCH = 0, ST_PEAK = 1, ST_HEADER = 2, ST_BODY = 3 };

/* Uniform (per-channel) part of the carried state; lane-distributed parts follow it.  Hop indices are
 * 32-bit wrapping counters: every difference the state machine takes is far below 2^31 hops.  The
 * first 48 bytes are the working set of every step (three 16-byte loads / stores); the rest is touched
 * on events only. */
struct ChanScalars {
    uint32_t state, nsym, total, flen;
    uint32_t next, prev_hop;
    int32_t acc;
    uint32_t ep_left;  /* symbols until the next tracker epoch boundary */
    uint32_t s_prev, s_prev2, osym_cnt, pad0;
    uint32_t best_h, peak_end; /* SEARCH / PEAK only */
    float best_q;
    uint32_t pad1;
    unsigned long long t0; /* absolute hop of the lock (frame start_sample) */
    uint32_t pad2[2];
    anm_chan_stats_t stats; /* 32 bytes */
    uint32_t pad3[4];
};
static_assert(sizeof(ChanScalars) == 128, "ChanScalars layout");

/* loads that bypass L1: per-channel data another SM may have written earlier in the SAME launch (multi-chunk launches) */
__device__ __forceinline__ uint32_t ld_cg_u8(const uint8_t *p) { return (uint32_t)__ldcg(p); }
__device__ __forceinline__ uint32_t ld_cg_u32(const uint32_t *p) { return __ldcg(p); }

/* hop record: energy of the strongest tone and its index (0xFF before the stream) */
struct HopRec {
    float e;
    uint32_t d;
};

struct KParams {
    const int16_t *pcm;
    unsigned long long ch_stride; /* samples */
    uint32_t n_ch;
    uint32_t n_syms;              /* symbol periods per channel in this chunk */
    unsigned long long hop_base;  /* absolute index of the chunk's first hop */
    unsigned char *state;         /* per channel: ChanScalars | lane records | tree carry */
    uint32_t state_stride;        /* bytes */
    uint32_t rot_mode;            /* 0: all tone bins = 0 mod 4; 1: all bins even; 2: general */
    uint8_t *fsyms;               /* per channel frame symbol store */
    uint32_t fsym_stride;
    uint32_t max_frame_syms;
    float *trE;                   /* [n_ch][tr_hops][T] or NULL */
    uint8_t *trD;                 /* [n_ch][tr_hops] or NULL */
    float *trEmax;                /* [n_ch][tr_hops] or NULL */
    unsigned long long tr_hops;
    anm_frame_t *frames;
    uint8_t *bytes;
    uint32_t *counters;           /* [0]=n_frames [1]=n_bytes [2]=frames dropped (queue full) [4]=channels handed out [5]=warps done;
                                   * all free-running (mod 2^32) since create / reset: nothing is re-armed between launches */
    uint32_t frames_cap, bytes_cap; /* powers of two: the queues are rings */
    uint32_t base_f, base_b;        /* counters as of what the host has consumed (mod 2^32) */
    uint32_t q_base, done_base;     /* values of counters[4] / counters[5] when this launch starts (launch index x n_ch / x warps) */
    uint32_t n_chunks;              /* k_demod: chunks of n_syms symbol periods per channel in this launch (>= 1), chunk c of a channel chunk_stride samples behind c - 1 */
    uint32_t prog_base;             /* progress[] of every channel when this launch starts */
    unsigned long long chunk_stride;
    uint32_t *progress;             /* [n_ch] chunks completed per channel (free-running): orders the chunks of a channel inside a multi-chunk launch */
    uint32_t *snap;                 /* pinned host memory: {n_frames, n_bytes, dropped, launch seq + 1} as they stand when this launch
                                     * ends, written by the last warp to leave (a stream-ordered snapshot without a copy) */
    uint32_t seq1;
    uint8_t *osyms;               /* [n_ch][osym_cap] or NULL */
    uint32_t osym_cap;
    uint32_t P, tol, max_payload, trk_epoch, trk_thresh, hdr_syms;
    uint32_t pre_plane[7];        /* bit-planes of the preamble tone indices */
    uint8_t preamble[ANM_MAX_PREAMBLE];
    const float2 *tw_global;      /* [N][T] (cos, sin) twiddle table in HBM; its first N/NQ rows are staged per CTA */
    unsigned long long tw_rot[2];  /* 2 bits per tone: tone_bin mod 4 (quarter-period rotation code) */
    uint16_t crc_pow[32];          /* x^(8(31-lane)+16) mod the CRC-16 polynomial, per lane */
    uint8_t crc8_tab[256];         /* CRC-8 (poly 0x07) of one byte, for the 2-byte header check */
    uint32_t fold;                 /* 1: centre-folded hop partials (SPEC 3); tw_global then holds the folded twiddles [H/2][T] */
    unsigned long long fold_odd;   /* bit per tone: 2*tone_bin/S is odd (odd hops of that tone change sign) */
    float2 fold_tw[256];           /* folded twiddles [H/2][T] when they fit (constant-bank operands of the arithmetic loop) */
    float2 fold_sg[16];            /* (-1, -1) for those tones, (1, 1) otherwise (and always when not folding): the factor the
                                    * first tree level applies to its odd-hop operand */
    const uint8_t *tc_basis;       /* dense tone sets: int8 basis panels [group][K chunk][32 columns][16] (anm_kernels_tc.cuh) */
};

/* packed pairs of fp32 travel as 64-bit registers; mov.b64 keeps the halves in a register pair (no shifts / ORs) */
__device__ __forceinline__ unsigned long long pk2(float lo, float hi) {
    unsigned long long r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ float2 upk2(unsigned long long v) {
    float2 r;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(v));
    return r;
}
__device__ __forceinline__ float2 ffma2(float a, float2 b, float2 c) {
    unsigned long long rd;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rd) : "l"(pk2(a, a)), "l"(pk2(b.x, b.y)), "l"(pk2(c.x, c.y)));
    return upk2(rd);
}
__device__ __forceinline__ float2 ffma2vv(float2 a, float2 b, float2 c) {
    unsigned long long rd;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rd) : "l"(pk2(a.x, a.y)), "l"(pk2(b.x, b.y)), "l"(pk2(c.x, c.y)));
    return upk2(rd);
}
__device__ __forceinline__ float2 fmul2(float2 a, float2 b) {
    unsigned long long rd;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(rd) : "l"(pk2(a.x, a.y)), "l"(pk2(b.x, b.y)));
    return upk2(rd);
}
__device__ __forceinline__ float2 fadd2(float2 a, float2 b) {
    unsigned long long rd;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(rd) : "l"(pk2(a.x, a.y)), "l"(pk2(b.x, b.y)));
    return upk2(rd);
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

/* int16 half HI of a packed word -> fp32 in one instruction (I2F.S16 Rd, Rs.H0/.H1) */
template <int HI>
__device__ __forceinline__ float cvt_s16(uint32_t w) {
    float f;
    if (HI) asm("{ .reg .b16 lo, hi; mov.b32 {lo, hi}, %1; cvt.rn.f32.s16 %0, hi; }" : "=f"(f) : "r"(w));
    else asm("{ .reg .b16 lo, hi; mov.b32 {lo, hi}, %1; cvt.rn.f32.s16 %0, lo; }" : "=f"(f) : "r"(w));
    return f;
}

/* the same conversion on the ALU pipe (PRMT sign extension + I2FP): two issue slots instead of one, but it
 * takes load off the quarter-rate XU pipe that I2F.S16 runs on */
template <int HI>
__device__ __forceinline__ float cvt_s16_alu(uint32_t w) {
    uint32_t x;
    float f;
    if (HI) asm("prmt.b32 %0, %1, 0, 0xBB32;" : "=r"(x) : "r"(w));
    else asm("prmt.b32 %0, %1, 0, 0x9910;" : "=r"(x) : "r"(w));
    asm("cvt.rn.f32.s32 %0, %1;" : "=f"(f) : "r"(x));
    return f;
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

/* The hop-record ring of a channel in shared memory: HopRec[64 * S], addressed by hop index & (64 * S - 1).  Two layouts.  PM (phase-major,
 * [hop % S][symbol period % 64]): every access of the state machine is one hop phase over the 32 symbol periods of the lanes, so consecutive
 * lanes read consecutive 8-byte records (2 wavefronts) instead of records S * 8 bytes apart (8 wavefronts for S = 4) -- used by k_demod_tc, whose
 * tensor-core operand reads compete for shared-memory bandwidth.  Linear ([hop]): three instructions less per access -- used by k_demod, which is
 * bound by instruction issue and does not notice the bank conflicts (measured both ways: conflicts 8.6 M -> 1.9 M wavefronts per launch, time
 * unchanged, +2.4 % instructions). */
template <int S, bool PM>
__device__ __forceinline__ uint32_t ring_off(uint32_t idx) { /* idx < 64 * S; byte offset of the record */
    return PM ? ((((idx & (uint32_t)(S - 1)) << 6) | (idx >> Log2<S>::v)) << 3) : (idx << 3);
}

/* Per-channel state in HBM: ChanScalars | HopRec[32 slots][S] | tree carry */
template <int T, int S>
__host__ __device__ constexpr uint32_t state_carry_offset() { return (uint32_t)sizeof(ChanScalars) + 32u * S * 8u; }
template <int T, int S>
__host__ __device__ constexpr uint32_t state_bytes() { return state_carry_offset<T, S>() + (uint32_t)(S - 1) * T * 8u; }
/* PCM stage: one row per lane (= symbol period), padded to an odd number of 16-byte chunks so that
 * the lanes' LDS.128 of the same chunk index fall into different banks */
template <int N>
__host__ __device__ constexpr uint32_t stage_row_bytes() { return 2u * N + 16u; }
template <int N>
__host__ __device__ constexpr uint32_t stage_bytes() { return 32u * stage_row_bytes<N>(); }
/* Per-warp shared memory: PCM stage | HopRec ring[64 slots][S] | scalars | tree carry */
template <int T, int N, int S>
__host__ __device__ constexpr uint32_t warp_smem_bytes() {
    return stage_bytes<N>() + 64u * S * 8u + 128u + (uint32_t)(S - 1) * T * 8u;
}
/* Hops that share one pass over the twiddle table: the table holds the first 1/NQ of a symbol
 * period; hop offsets that differ by N/NQ rotate every twiddle by an exact multiple of 90 degrees. */
template <int S>
__host__ __device__ constexpr int quad_hops() { return S >= 4 ? 4 : 2; }
/* CTA-shared twiddle table: (N / NQ) positions x T tones x (cos, sin) */
template <int T, int N, int S>
__host__ __device__ constexpr uint32_t cta_smem_bytes() { return (uint32_t)(N / quad_hops<S>()) * T * 8u; }

/* ================= sync / slicing / framing (SPEC 5) =================
 * One call per step (32 symbol periods) and channel, by the whole warp.  The step's hop records are
 * already in the ring `sr` (HopRec[64*S], indexed by hop-in-chunk & RM); dc[] are this lane's own
 * decisions of the step (lane = symbol period). */
template <int T, int N, int S, bool PM>
__device__ __forceinline__ void sm_step(const KParams &p, const uint32_t ch, const int lane, const uint32_t sr,
                                        const uint32_t hic, const int nvalid, const bool active,
                                        const uint32_t (&dc)[S], const uint32_t ssa, const uint32_t crc_k, const unsigned long long hop_base) {
    constexpr int H = N / S;
    constexpr int B = Log2<T>::v;
    constexpr int LV = Log2<S>::v;
    constexpr uint32_t RM = 64u * S - 1u;
    constexpr uint32_t FULL = 0xffffffffu;
    const uint32_t hb = (uint32_t)hop_base + hic; /* wrapping index of the step's first hop */
    /* hop record by hop index r relative to the step start, r in [-32S, 32S): .x = emax bits, .y = d */
    auto REC = [&](int r) -> uint2 {
        uint2 v;
        asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(sr + ring_off<S, PM>((hic + (uint32_t)r) & RM)) : "memory");
        return v;
    };
    /* warp-uniform working set: three broadcast LDS.128; everything else of ChanScalars is read / written
     * in shared memory (address ssa) when an event needs it */
    struct {
        uint32_t state, nsym, total, flen, next, prev_hop;
        int32_t acc;
        uint32_t ep_left, s_prev, s_prev2, osym_cnt, pad0;
    } sc;
    {
        uint32_t *w = reinterpret_cast<uint32_t *>(&sc);
#pragma unroll
        for (int i = 0; i < 3; ++i)
            asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(w[4 * i]), "=r"(w[4 * i + 1]), "=r"(w[4 * i + 2]), "=r"(w[4 * i + 3]) : "r"(ssa + 16u * i) : "memory");
    }
    auto cold_ld = [&](uint32_t off) -> uint32_t {
        uint32_t v;
        asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(ssa + off) : "memory");
        return v;
    };
    auto cold_st = [&](uint32_t off, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(ssa + off), "r"(v) : "memory"); }; /* every lane, same value */
    constexpr uint32_t O_BEST_H = offsetof(ChanScalars, best_h), O_PEAK_END = offsetof(ChanScalars, peak_end), O_BEST_Q = offsetof(ChanScalars, best_q),
                       O_T0 = offsetof(ChanScalars, t0), O_STATS = offsetof(ChanScalars, stats);
    constexpr uint32_t O_LOCKS = O_STATS + offsetof(anm_chan_stats_t, locks), O_HFAIL = O_STATS + offsetof(anm_chan_stats_t, header_fail),
                       O_FOK = O_STATS + offsetof(anm_chan_stats_t, frames_ok), O_FBAD = O_STATS + offsetof(anm_chan_stats_t, frames_bad),
                       O_SYMS = O_STATS + offsetof(anm_chan_stats_t, symbols), O_TRK = O_STATS + offsetof(anm_chan_stats_t, trk_moves);
    uint32_t syms_step = 0; /* symbols decided in this step, folded into stats.symbols at the end */
    const int endh = nvalid * S;
    int cur = 0;
    bool have_cand = false;
    uint32_t cand[S];
#pragma unroll
    for (int i = 0; i < S; ++i) cand[i] = 0;

    /* quality of the alignment whose last preamble symbol ends at relative hop h: SPEC 5 q(h) */
    auto quality = [&](int h) -> float {
        const int pl = min(lane, (int)p.P - 1);
        const uint2 rv = REC(h - ((int)p.P - 1 - pl) * S);
        float leaf = (lane < (int)p.P && rv.y == (uint32_t)p.preamble[pl]) ? __uint_as_float(rv.x) : 0.0f;
#pragma unroll
        for (uint32_t w = 1; w < (uint32_t)ANM_MAX_PREAMBLE; w <<= 1) {
            const float o = __shfl_xor_sync(FULL, leaf, w);
            if (w < p.P) leaf = __fadd_rn(leaf, o);
        }
        return __shfl_sync(FULL, leaf, 0);
    };

#pragma unroll 1
    while (cur < endh) {
        if (sc.state <= ST_PEAK) {
            if (!have_cand) {
                /* preamble correlation on bit-planes of the hop decisions */
                const int sh = 32 + lane - (int)(p.P - 1); /* bit of preamble symbol 0 in the 64-bit history */
                const uint32_t pmask = (p.P >= 32) ? 0xffffffffu : ((1u << p.P) - 1u);
#pragma unroll
                for (int i = 0; i < S; ++i) {
                    const uint32_t dprev = REC((lane - 32) * S + i).y; /* same lane, previous step */
                    uint32_t mism = 0;
                    auto plane = [&](int j) {
                        const uint32_t bc = (j < B) ? ((dc[i] >> j) & 1u) : (dc[i] > (uint32_t)(T - 1));
                        const uint32_t bp = (j < B) ? ((dprev >> j) & 1u) : (dprev > (uint32_t)(T - 1));
                        const unsigned long long hist = ((unsigned long long)__ballot_sync(FULL, bc) << 32) | __ballot_sync(FULL, bp);
                        const uint32_t w = (uint32_t)(hist >> sh);
                        mism |= (j < B) ? (w ^ p.pre_plane[j]) : w;
                    };
                    if (B >= 4) {
                        /* Many bits per symbol: mismatches only accumulate from plane to plane, so when the two lowest planes
                         * already rule out every alignment of this phase (random symbols agree with the preamble in two bits one
                         * time in four) the other planes need not be looked at.  Same candidates, fewer ballots. */
                        plane(0);
                        plane(1);
                        if (__ballot_sync(FULL, (uint32_t)__popc(mism & pmask) <= p.tol) == 0u) {
                            cand[i] = 0u;
                            continue;
                        }
#pragma unroll
                        for (int j = 2; j <= B; ++j) plane(j);
                    } else {
#pragma unroll
                        for (int j = 0; j <= B; ++j) plane(j);
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
                if (h0 == 0x7fffffff) break;
                cold_st(O_BEST_Q, __float_as_uint(quality(h0)));
                cold_st(O_BEST_H, hb + (uint32_t)h0);
                cold_st(O_PEAK_END, hb + (uint32_t)(h0 + S - 1));
                sc.state = ST_PEAK;
                cur = h0 + 1;
            } else {
                const int pend = (int)(cold_ld(O_PEAK_END) - hb);
                float best_q = __uint_as_float(cold_ld(O_BEST_Q));
                uint32_t best_h = cold_ld(O_BEST_H);
                while (cur < endh && cur <= pend) {
                    if ((pick<S>(cand, cur & (S - 1)) >> (cur >> LV)) & 1u) {
                        const float q = quality(cur);
                        if (q > best_q) { best_q = q; best_h = hb + (uint32_t)cur; }
                    }
                    ++cur;
                }
                cold_st(O_BEST_Q, __float_as_uint(best_q));
                cold_st(O_BEST_H, best_h);
                if (cur > pend) {
                    const unsigned long long t0 = hop_base + (unsigned long long)hic + (long long)(int)(best_h - hb);
                    cold_st(O_T0, (uint32_t)t0);
                    cold_st(O_T0 + 4u, (uint32_t)(t0 >> 32));
                    sc.next = best_h + S;
                    sc.nsym = 0;
                    sc.acc = 0;
                    sc.ep_left = p.trk_epoch;
                    sc.s_prev = p.preamble[p.P - 1];
                    sc.s_prev2 = 0xFFu;
                    sc.prev_hop = best_h;
                    sc.state = ST_HEADER;
                    cold_st(O_LOCKS, cold_ld(O_LOCKS) + 1u);
                }
            }
        } else {
            /* ---- locked: slice up to 32 symbols at once (lane = symbol) ---- */
            const int first = (int)(sc.next - hb);
            if (first >= endh) break;
            const uint32_t until_evt = (sc.state == ST_HEADER ? p.hdr_syms : sc.total) - sc.nsym;
            uint32_t cnt = min(until_evt, (uint32_t)(((endh - 1 - first) >> LV) + 1));
            const int s0 = first >> LV;
            const int e = lane - s0;                 /* index of this lane's symbol in the run */
            const int hr = first + (e << LV);        /* its relative hop (ring-addressed for every lane) */
            const bool part = (uint32_t)e < cnt;
            const uint32_t sym = REC(hr).y;
            uint8_t *fs = p.fsyms + (size_t)ch * p.fsym_stride;
            if (part) {
                fs[sc.nsym + e] = (uint8_t)sym;
                if (p.osyms) {
                    const uint32_t oi = sc.osym_cnt + e;
                    if (oi < p.osym_cap) p.osyms[(size_t)ch * p.osym_cap + oi] = (uint8_t)sym;
                }
            }
            /* tracker votes (SPEC 5): the lane of symbol n votes for symbol n-1, whose hop is one
             * symbol back -- or, for the first symbol of the run, the carried prev_hop.  Its tone
             * s_j is the decision at that hop (the nominal s_{-1} never is the subject of a vote). */
            const int hj = (e <= 0) ? (int)(sc.prev_hop - hb) : hr - S;
            const uint2 rj = REC(hj), rje = REC(hj - 1), rjl = REC(hj + 1);
            const uint32_t sj2 = REC(hr - 2 * S).y;
            const uint32_t sj = rj.y;
            /* tone of symbol n-2: for the first symbols of a run the carried values */
            const uint32_t sjm = (e >= 2) ? sj2 : ((e == 1) ? sc.s_prev : sc.s_prev2);
            const float e_on = __uint_as_float(rj.x);
            const bool early = rje.y == sj && __uint_as_float(rje.x) > e_on;
            const bool late = rjl.y == sj && __uint_as_float(rjl.x) > e_on;
            const bool voter = part && (sc.nsym + (uint32_t)e >= 1u);
            const uint32_t bl = __ballot_sync(FULL, voter && (sym != sj) && late);
            const uint32_t be = __ballot_sync(FULL, voter && (sjm != sj) && early);
            /* tracker epochs inside the run: only an actual timing move ends the run early */
            int adj = 0;
            if ((bl | be) == 0u && sc.acc < (int)p.trk_thresh && sc.acc > -(int)p.trk_thresh) {
                /* no vote in this run and the carried sum cannot trip: epochs just tick over */
                if (cnt < sc.ep_left) {
                    sc.ep_left -= cnt;
                } else {
                    uint32_t rem = cnt - sc.ep_left;
                    while (rem >= p.trk_epoch) rem -= p.trk_epoch;
                    sc.ep_left = p.trk_epoch - rem;
                    sc.acc = 0;
                }
            } else {
                uint32_t pos = 0;
                while (true) {
                    const uint32_t eb = pos + sc.ep_left; /* symbols of the run up to the next boundary */
                    const uint32_t hi = min(eb, cnt);
                    const uint32_t lo_m = 0xffffffffu << ((uint32_t)s0 + pos);
                    const uint32_t hi_m = ((uint32_t)s0 + hi >= 32u) ? 0xffffffffu : ((1u << ((uint32_t)s0 + hi)) - 1u);
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
            const int lasth = first + (int)((cnt - 1) << LV);
            sc.s_prev2 = (cnt >= 2) ? REC(lasth - S).y : sc.s_prev;
            sc.s_prev = REC(lasth).y;
            sc.prev_hop = hb + (uint32_t)lasth;
            sc.nsym += cnt;
            sc.osym_cnt += cnt;
            sc.next += (cnt << LV) + (uint32_t)adj;
            syms_step += cnt;
            if (adj) cold_st(O_TRK, cold_ld(O_TRK) + (uint32_t)adj);
            cur = lasth + 1;
            if (sc.state == ST_HEADER && sc.nsym == p.hdr_syms) {
                __syncwarp();
                /* 24 header bits from hdr_syms symbols, one symbol per lane, OR-reduced */
                uint32_t contrib = 0;
                if (lane < (int)p.hdr_syms) {
                    const uint32_t v = gray_inv(ld_cg_u8(fs + lane));
                    const int pos = 24 - B * (lane + 1);
                    contrib = (pos >= 0) ? (v << pos) : (v >> (-pos));
                }
                const uint32_t hdr = __reduce_or_sync(FULL, contrib);
                const uint32_t len = hdr >> 8;
                /* CRC-8 of the two LEN bytes: two look-ups in the host-built byte table */
                const uint32_t c8 = p.crc8_tab[p.crc8_tab[(hdr >> 16) & 0xffu] ^ ((hdr >> 8) & 0xffu)];
                if (len == 0 || len > p.max_payload || c8 != (hdr & 0xffu)) {
                    cold_st(O_HFAIL, cold_ld(O_HFAIL) + 1u);
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
                const bool fits = (fidx - p.base_f) < p.frames_cap && (boff + len - p.base_b) <= p.bytes_cap;
                const uint8_t *bs = fs + p.hdr_syms;
                /* body byte byi (payload, then the two CRC bytes) from its symbols */
                auto body_byte = [&](uint32_t byi) -> uint32_t {
                    uint32_t v8 = 0;
                    if (B == 2 && (p.hdr_syms & 3u) == 0u) {
                        /* four 2-bit symbols in one aligned word: Gray-decode all four at once */
                        const uint32_t wv = ld_cg_u32(reinterpret_cast<const uint32_t *>(bs + byi * 4u));
                        const uint32_t v = wv ^ ((wv >> 1) & 0x01010101u);
                        v8 = ((v & 3u) << 6) | (((v >> 8) & 3u) << 4) | (((v >> 16) & 3u) << 2) | ((v >> 24) & 3u);
                    } else if (8 % B == 0) {
                        constexpr int SPB = (8 % B == 0) ? 8 / B : 1;
#pragma unroll
                        for (int j = 0; j < SPB; ++j) v8 |= gray_inv(ld_cg_u8(bs + byi * SPB + j)) << (B * (SPB - 1 - j));
                    } else {
#pragma unroll
                        for (int k = 0; k < 8; ++k) {
                            const uint32_t bit = byi * 8 + k;
                            const uint32_t v = gray_inv(ld_cg_u8(bs + bit / B));
                            v8 = (v8 << 1) | ((v >> (B - 1 - (bit % B))) & 1u);
                        }
                    }
                    return v8;
                };
                /* bits -> bytes and CRC-16, both lane-parallel.  The message (LEN bytes +
                 * payload) is consumed 32 bytes per round, right-aligned: lane l holds the byte
                 * that is 31-l positions from the end of the round, multiplies it by
                 * x^(8(31-l)+16) mod p (crc_k, a lane constant) in GF(2)[x], and the round is the
                 * XOR of all lanes; the running CRC enters the next round through its first two
                 * bytes (a CRC register R equals XORing R into the next two message bytes). */
                const uint32_t mlen = len + 2; /* CRC'd bytes: LEN hi, LEN lo, payload */
                uint32_t crc = 0xFFFFu;
                uint32_t done = 0;
                uint32_t take = mlen & 31u; /* first (short) round */
                if (take == 0) take = 32;
#pragma unroll 1
                while (done < mlen) {
                    const int li = lane - (32 - (int)take); /* index within the round */
                    uint32_t v8 = 0;
                    if (li >= 0) {
                        const uint32_t mi = done + li; /* index in the CRC'd message */
                        if (mi < 2) {
                            v8 = (mi == 0) ? (len >> 8) : (len & 0xffu);
                        } else {
                            v8 = body_byte(mi - 2);
                            if (fits) p.bytes[(boff + mi - 2) & (p.bytes_cap - 1u)] = (uint8_t)v8;
                        }
                        if (li == 0) v8 ^= crc >> 8;
                        if (li == 1) v8 ^= crc & 0xffu;
                    }
                    /* v8 * crc_k mod p, Horner over the 8 bits */
                    uint32_t acc = 0;
#pragma unroll
                    for (int k = 7; k >= 0; --k) {
                        acc = ((acc << 1) ^ ((acc & 0x8000u) ? 0x1021u : 0u)) & 0xffffu;
                        if ((v8 >> k) & 1u) acc ^= crc_k;
                    }
                    acc = __reduce_xor_sync(FULL, acc);
                    /* a 1-byte round has no second byte to carry the register's low byte:
                     * R_lo * x^(8n) with n = 1 is R_lo << 8 */
                    if (take == 1) acc ^= (crc & 0xffu) << 8;
                    crc = acc;
                    done += take;
                    take = 32;
                }
                /* received CRC-16: the two bytes after the payload */
                const uint32_t rx = (lane < 2) ? (body_byte(len + lane) << (8 * (1 - lane))) : 0u;
                const uint32_t crc_rx = __reduce_or_sync(FULL, rx);
                const uint32_t ok = crc == crc_rx;
                if (fits) {
                    if (lane == 0) {
                        anm_frame_t f;
                        f.channel = ch;
                        f.len = len;
                        const unsigned long long t0 = (unsigned long long)cold_ld(O_T0) | ((unsigned long long)cold_ld(O_T0 + 4u) << 32);
                        f.start_sample = (t0 + 1 - (unsigned long long)p.P * S) * H;
                        f.crc_ok = ok;
                        f.offset = boff;
                        p.frames[fidx & (p.frames_cap - 1u)] = f;
                    }
                    cold_st(ok ? O_FOK : O_FBAD, cold_ld(ok ? O_FOK : O_FBAD) + 1u);
                } else if (lane == 0) {
                    atomicAdd(&p.counters[2], 1u);
                }
                sc.state = ST_SEARCH;
            }
        }
    }
    if (syms_step) {
        const unsigned long long n = ((unsigned long long)cold_ld(O_SYMS) | ((unsigned long long)cold_ld(O_SYMS + 4u) << 32)) + syms_step;
        cold_st(O_SYMS, (uint32_t)n);
        cold_st(O_SYMS + 4u, (uint32_t)(n >> 32));
    }
    {
        const uint32_t *w = reinterpret_cast<const uint32_t *>(&sc);
#pragma unroll
        for (int i = 0; i < 3; ++i)
            asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(ssa + 16u * i), "r"(w[4 * i]), "r"(w[4 * i + 1]), "r"(w[4 * i + 2]), "r"(w[4 * i + 3]) : "memory");
    }
    __syncwarp();
}

/* The last warp of a launch to get here (every warp of the grid calls this once, lane 0) writes the queue counters as they
 * stand at the end of the launch into pinned host memory.  All frame records and payload bytes of the launch were written
 * before the counting atomics that precede this call, and the host reads the snapshot only after the launch's event. */
__device__ __forceinline__ void publish_snapshot(const KParams &p, uint32_t total_warps) {
    __threadfence();
    if (atomicAdd(&p.counters[5], 1u) - p.done_base == total_warps - 1u) {
        __threadfence();
        volatile uint32_t *c = p.counters;
        volatile uint32_t *s = p.snap;
        s[0] = c[0];
        s[1] = c[1];
        s[2] = c[2];
        s[3] = p.seq1;
        __threadfence_system();
    }
}

constexpr int kMaxWarps = 20; /* registers are allocated per 4 warps: 20 warps x 96 registers fit the file; 24 would cap at 80 */

/* MODE 0: streaming demodulator (sync, slicing, framing; no trace output).
 * MODE 1: stateless tone-energy pass (trace outputs only; parity / debug).
 * FOLD 1: centre-folded hop partials (configurations for which anm_config_foldable holds), 0: direct form. */
template <int T, int N, int S, int MODE, int FOLD>
__global__ void __launch_bounds__(kMaxWarps * 32) k_demod(const __grid_constant__ KParams p) {
    constexpr int H = N / S;
    constexpr int NQ = quad_hops<S>(); /* hops per table pass */
    constexpr int TL = N / NQ;         /* table positions */
    constexpr int GR = S / NQ;         /* passes per symbol period (hop = pass + q*GR) */
    constexpr int TG = (T * NQ <= 16) ? T : (16 / NQ); /* tones per register group: NQ*TG accumulators */
    constexpr int NG = T / TG;
    constexpr int LV = Log2<S>::v;
    constexpr int CPH = H / 8; /* 16-byte chunks per hop */
    
    constexpr uint32_t RM = 64u * S - 1u; /* record ring mask (hops) */
    constexpr uint32_t FULL = 0xffffffffu;
    static_assert(N >= 64 && (H % 8) == 0 && S >= 2, "unsupported geometry");

    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31;
    const int wib = threadIdx.x >> 5;
    const int wpb = blockDim.x >> 5;
    /* CTA-shared twiddle table (first 1/NQ of a symbol period), read back as broadcast LDS.128 */
    {
        const float4 *gsrc = reinterpret_cast<const float4 *>(p.tw_global);
        float4 *dst = reinterpret_cast<float4 *>(smem_raw);
        const int n4 = FOLD ? (H / 2) * T / 2 : TL * T / 2; /* folded twiddles [H/2][T] or first quarter period [TL][T] */
        for (int i = threadIdx.x; i < n4; i += blockDim.x) dst[i] = __ldg(&gsrc[i]);
        __syncthreads();
    }
    const uint32_t stw = (uint32_t)__cvta_generic_to_shared(smem_raw);
    unsigned char *wsm = smem_raw + cta_smem_bytes<T, N, S>() + (size_t)wib * warp_smem_bytes<T, N, S>();
    const uint32_t stage = (uint32_t)__cvta_generic_to_shared(wsm);
    const uint32_t sr = stage + stage_bytes<N>(); /* HopRec ring [64*S] */
    ChanScalars *ssc = reinterpret_cast<ChanScalars *>(wsm + stage_bytes<N>() + 64u * S * 8u);
    float2 *carry = reinterpret_cast<float2 *>(wsm + stage_bytes<N>() + 64u * S * 8u + 128u); /* [(S-1)*T] */

    constexpr uint32_t RS = stage_row_bytes<N>();
    constexpr int CPS = N / 8; /* 16-byte chunks per symbol period */
    /* cp.async: per instruction the warp copies 32 consecutive 16-byte chunks (512 contiguous bytes);
     * lane l lands in row cp_row (+ rows per instruction), chunk cp_col of the padded stage */
    constexpr int LPS = (CPS >= 32) ? 1 : 32 / CPS; /* stage rows covered by one cp.async instruction */
    const uint32_t cp_row = (CPS >= 32) ? 0u : (uint32_t)lane / (uint32_t)CPS;
    const uint32_t cp_col = (uint32_t)lane % (uint32_t)CPS;

    /* CRC-16 lane constant: x^(8(31-lane)+16) mod p (see frame assembly) */
    const uint32_t crc_k = (MODE == 0) ? (uint32_t)p.crc_pow[lane] : 0u; /* computed once on the host */

    const uint32_t n_steps = (p.n_syms + 31u) / 32u;
    const uint32_t total_warps = gridDim.x * wpb;

    /* Work items = (chunk, channel), chunk-major: the first one static, further ones from the work queue.  With several chunks in one launch
     * (anm_demod_feed_device_chunks) a warp that finishes a channel's chunk goes on with whatever comes next instead of idling through the
     * tail of a wave; chunk c of a channel waits for its chunk c - 1 -- handed out n_ch tickets earlier, so almost always long finished. */
    const uint32_t n_items = p.n_ch * p.n_chunks;
    const bool multi = MODE == 0 && p.n_chunks > 1u;
    uint32_t item = blockIdx.x * wpb + wib;
    if (multi) {
        /* every item from the queue, the first one too: a warp then only ever waits for items that RUNNING warps hold (tickets go out in
         * order), whatever part of the grid is resident */
        uint32_t nx = 0;
        if (lane == 0) nx = atomicAdd(&p.counters[4], 1u);
        item = __shfl_sync(0xffffffffu, nx, 0) - p.q_base;
    }
    while (item < n_items) {
        const uint32_t chunk = multi ? item / p.n_ch : 0u, ch = item - chunk * p.n_ch;
        const unsigned long long hop_base = p.hop_base + (unsigned long long)chunk * p.n_syms * S;
        unsigned char *stp = p.state + (size_t)ch * p.state_stride;
        uint2 *grec = reinterpret_cast<uint2 *>(stp + sizeof(ChanScalars));
        float2 *gcarry = reinterpret_cast<float2 *>(stp + state_carry_offset<T, S>());
        if constexpr (MODE == 0) if (multi && chunk > 0u) {
            if (lane == 0) {
                uint32_t done;
                for (;;) {
                    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(done) : "l"(p.progress + ch) : "memory");
                    if (done - p.prog_base >= chunk) break;
                    __nanosleep(200);
                }
            }
            __syncwarp();
        }

        /* ---- restore carried state: the last 32 symbol slots go to ring slots 32..63 (L1 bypassed: see ld_cg_u8) ---- */
        __syncwarp();
#pragma unroll
        for (int i = 0; i < S; ++i) {
            const uint2 rv = __ldcg(&grec[lane * S + i]);
            asm volatile("st.shared.v2.u32 [%0], {%1,%2};" ::"r"(sr + ring_off<S, false>((uint32_t)((32 + lane) * S + i))), "r"(rv.x), "r"(rv.y) : "memory");
        }
        for (int i = lane; i < (S - 1) * T; i += 32) carry[i] = __ldcg(&gcarry[i]);
        if (MODE == 0) reinterpret_cast<uint32_t *>(ssc)[lane] = __ldcg(&reinterpret_cast<const uint32_t *>(stp)[lane]);
        __syncwarp();

        const char *src = reinterpret_cast<const char *>(p.pcm + (size_t)ch * p.ch_stride + (size_t)chunk * p.chunk_stride);
        auto issue = [&](uint32_t step) {
            const uint32_t nv = min(32u, p.n_syms - step * 32u);
            const char *g = src + (size_t)step * (32u * N * 2u) + (size_t)lane * 16u;
            if (CPS >= 32) {
                constexpr uint32_t IPS = (CPS >= 32) ? CPS / 32 : 1; /* instructions per row */
#pragma unroll 4
                for (uint32_t q = 0; q < 32u * IPS; ++q) {
                    const uint32_t r = q / IPS, c = (q % IPS) * 32u + lane;
                    if (r < nv) cp_async16(stage + r * RS + (c << 4), g + (size_t)q * 512u);
                }
            } else {
                const uint32_t d0 = stage + cp_row * RS + (cp_col << 4);
                if (nv == 32u) {
#pragma unroll
                    for (int q = 0; q < CPS; ++q) cp_async16(d0 + (uint32_t)(q * LPS) * RS, g + (size_t)q * 512u);
                } else {
#pragma unroll 4
                    for (int q = 0; q < CPS; ++q)
                        if ((uint32_t)(q * LPS) + cp_row < nv) cp_async16(d0 + (uint32_t)(q * LPS) * RS, g + (size_t)q * 512u);
                }
            }
            cp_async_commit();
        };
        if (n_steps) issue(0);

#pragma unroll 1
        for (uint32_t step = 0; step < n_steps; ++step) {
            cp_async_wait<0>();
            __syncwarp();
            const uint32_t row = stage + (uint32_t)lane * RS;
            const int nvalid = (int)min(32u, p.n_syms - step * 32u);
            const bool active = lane < nvalid;
            const uint32_t hic = step * 32u * S; /* hop index of the step start within the chunk (ring position) */

            /* ================= tone energies (SPEC 3) ================= */
            uint32_t dc[S];
            float ec[S];
#pragma unroll
            for (int i = 0; i < S; ++i) { dc[i] = 0xFFu; ec[i] = 0.0f; }

#pragma unroll 1
            for (int g = 0; g < NG; ++g) {
                float2 Pp[S][TG]; /* hop partials of this tone group */
                /* All 32 lanes run the arithmetic (lanes beyond a ragged chunk end read stale shared
                 * memory; their results are discarded below).  One pass accumulates the NQ hops whose
                 * offsets differ by N/NQ: they see the same twiddle sequence up to an exact rotation by
                 * multiples of 90 degrees, applied once after the chain. */
                if ((H % 16) == 0 && FOLD) {
                    /* Centre folding (SPEC 3): the samples k + 1/2 after and before a hop centre share a twiddle
                     * up to conjugation, so their exact sum and difference feed ONE packed FMA per tone:
                     * (A, Bq) += (a + b, a - b) * (cos, sin).  Every hop uses the same H/2 twiddles. */
#pragma unroll
                    for (int pass = 0; pass < GR; ++pass) {
                        float2 acc[NQ][TG];
#pragma unroll
                        for (int q = 0; q < NQ; ++q)
#pragma unroll
                            for (int t = 0; t < TG; ++t) acc[q][t] = make_float2(0.f, 0.f);
                        uint32_t twa = stw + (uint32_t)(g * TG) * 8u;
                        uint32_t fwd = row + (uint32_t)(pass * 2 * H + H), bwd = fwd; /* walk away from the hop centres */
#if ANM_LDS128
                        uint4 wf[NQ], wb4[NQ];
#endif
#pragma unroll(NG == 1 ? H / 8 : 1) /* one tone group: straight-line loop, every twiddle address an immediate */
                        for (int i = 0; i < H / 8; ++i, twa += 4 * T * 8, fwd += 8u) {
                            bwd -= 8u;
                            uint2 vf[NQ], vb[NQ]; /* four samples after / before the centre of each hop */
#if ANM_LDS128
                            /* 16-byte loads (conflict-free with the odd-chunk row pitch): every second iteration fetches the
                             * eight samples on either side that this and the next iteration consume */
#endif
#pragma unroll
                            for (int q = 0; q < NQ; ++q) {
#if ANM_LDS128
                                if ((i & 1) == 0) {
                                    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(wf[q].x), "=r"(wf[q].y), "=r"(wf[q].z), "=r"(wf[q].w) : "r"(fwd + (uint32_t)(q * GR * 2 * H)));
                                    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(wb4[q].x), "=r"(wb4[q].y), "=r"(wb4[q].z), "=r"(wb4[q].w) : "r"(bwd - 8u + (uint32_t)(q * GR * 2 * H)));
                                    vf[q] = make_uint2(wf[q].x, wf[q].y);
                                    vb[q] = make_uint2(wb4[q].z, wb4[q].w);
                                } else {
                                    vf[q] = make_uint2(wf[q].z, wf[q].w);
                                    vb[q] = make_uint2(wb4[q].x, wb4[q].y);
                                }
#else
                                asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(vf[q].x), "=r"(vf[q].y) : "r"(fwd + (uint32_t)(q * GR * 2 * H)));
                                asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(vb[q].x), "=r"(vb[q].y) : "r"(bwd + (uint32_t)(q * GR * 2 * H)));
#endif
                            }
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                float2 xs[NQ];
#pragma unroll
                                for (int q = 0; q < NQ; ++q) {
                                    const uint32_t wa = (j < 2) ? vf[q].x : vf[q].y;
                                    const uint32_t wb = (j < 2) ? vb[q].y : vb[q].x;
                                    const float a = (j & 1) ? cvt_s16<1>(wa) : cvt_s16<0>(wa);               /* XU pipe */
                                    const float b = ((ANM_CVT_ALU_MASK >> j) & 1) ? ((j & 1) ? cvt_s16_alu<0>(wb) : cvt_s16_alu<1>(wb))   /* element 3 - j, ALU pipe */
                                                                                  : ((j & 1) ? cvt_s16<0>(wb) : cvt_s16<1>(wb));
                                    /* (a + b, a - b) as one packed FMA: (a, b) * (1, -1) + (b, a); both halves exact */
                                    xs[q] = ffma2vv(make_float2(a, b), make_float2(1.0f, -1.0f), make_float2(b, a));
                                }
#pragma unroll
                                for (int t = 0; t < TG; t += 2) {
                                    float4 w2;
                                    if ((H / 2) * T <= 256 && NG == 1) {
                                        /* warp-uniform twiddles straight from the kernel parameters (constant bank / uniform registers) */
                                        const float2 wa = p.fold_tw[(i * 4 + j) * T + g * TG + t], wb = p.fold_tw[(i * 4 + j) * T + g * TG + t + 1];
                                        w2 = make_float4(wa.x, wa.y, wb.x, wb.y);
                                    } else {
                                        asm("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];"
                                            : "=f"(w2.x), "=f"(w2.y), "=f"(w2.z), "=f"(w2.w)
                                            : "r"(twa + (uint32_t)(j * T + t) * 8u));
                                    }
#pragma unroll
                                    for (int q = 0; q < NQ; ++q) {
                                        if (NG == 1 && i == 0 && j == 0) {
                                            /* straight-line form: the chain starts with a product instead of 0 + product (the only
                                             * difference, the sign of an all-zero sum, cannot reach an energy: E = fma(I, I, Q * Q)) */
                                            acc[q][t] = fmul2(xs[q], make_float2(w2.x, w2.y));
                                            acc[q][t + 1] = fmul2(xs[q], make_float2(w2.z, w2.w));
                                        } else {
                                            acc[q][t] = ffma2vv(xs[q], make_float2(w2.x, w2.y), acc[q][t]);
                                            acc[q][t + 1] = ffma2vv(xs[q], make_float2(w2.z, w2.w), acc[q][t + 1]);
                                        }
                                    }
                                }
                            }
                        }
                        /* relative quarter turns between hop centres: a sign flip on odd hops of "odd" tones; the first
                         * level of the window tree applies it (every level-1 add pairs one odd and one even hop) */
#pragma unroll
                        for (int q = 0; q < NQ; ++q)
#pragma unroll
                            for (int t = 0; t < TG; ++t) Pp[pass + q * GR][t] = acc[q][t];
                    }
                } else
#pragma unroll
                for (int pass = 0; pass < GR; ++pass) {
                    float2 acc[NQ][TG];
#pragma unroll
                    for (int q = 0; q < NQ; ++q)
#pragma unroll
                        for (int t = 0; t < TG; ++t) acc[q][t] = make_float2(0.f, 0.f);
                    uint32_t twa = stw + (uint32_t)((pass * H) * T + g * TG) * 8u;
#pragma unroll 1
                    for (int c = 0; c < TL / 8 / GR; ++c, twa += 8 * T * 8) {
                        uint4 v[NQ];
#pragma unroll
                        for (int q = 0; q < NQ; ++q) v[q] = lds128(row + (uint32_t)(((pass + q * GR) * CPH + c) * 16));
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            float x[NQ];
#pragma unroll
                            for (int q = 0; q < NQ; ++q) {
                                const uint32_t w = (j < 2) ? v[q].x : (j < 4) ? v[q].y : (j < 6) ? v[q].z : v[q].w;
                                x[q] = (j & 1) ? cvt_s16<1>(w) : cvt_s16<0>(w);
                            }
#pragma unroll
                            for (int t = 0; t < TG; t += 2) {
                                float4 w2;
                                asm("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];"
                                    : "=f"(w2.x), "=f"(w2.y), "=f"(w2.z), "=f"(w2.w)
                                    : "r"(twa + (uint32_t)(j * T + t) * 8u));
#pragma unroll
                                for (int q = 0; q < NQ; ++q) {
                                    acc[q][t] = ffma2(x[q], make_float2(w2.x, w2.y), acc[q][t]);
                                    acc[q][t + 1] = ffma2(x[q], make_float2(w2.z, w2.w), acc[q][t + 1]);
                                }
                            }
                        }
                    }
                    /* rotate hop q by (-j)^(bin*q*4/NQ): exact swap / negate (SPEC 3, table symmetry).
                     * rot_mode 0: every rotation is the identity; 1: all bins even, sign flips only. */
                    const unsigned long long rots = (g * TG < 32) ? (p.tw_rot[0] >> (2 * g * TG)) : (p.tw_rot[1] >> (2 * (g * TG - 32)));
                    if (p.rot_mode == 0u) {
#pragma unroll
                        for (int q = 0; q < NQ; ++q)
#pragma unroll
                            for (int t = 0; t < TG; ++t) Pp[pass + q * GR][t] = acc[q][t];
                    } else if (p.rot_mode == 1u) {
#pragma unroll
                        for (int q = 0; q < NQ; ++q) {
#pragma unroll
                            for (int t = 0; t < TG; ++t) {
                                /* bins even: r = (bin mod 4) * m mod 4 with m = q*4/NQ is 2 iff m is odd and bin = 2 mod 4 */
                                if (((q * (4 / NQ)) & 1) == 0) {
                                    Pp[pass + q * GR][t] = acc[q][t];
                                } else {
                                    const uint32_t mk = (((uint32_t)(rots >> (2 * t)) & 2u) != 0u) ? 0x80000000u : 0u;
                                    Pp[pass + q * GR][t] = make_float2(__uint_as_float(__float_as_uint(acc[q][t].x) ^ mk),
                                                                       __uint_as_float(__float_as_uint(acc[q][t].y) ^ mk));
                                }
                            }
                        }
                    } else {
#pragma unroll
                        for (int q = 0; q < NQ; ++q) {
#pragma unroll
                            for (int t = 0; t < TG; ++t) {
                                const uint32_t r = (((uint32_t)(rots >> (2 * t)) & 3u) * (uint32_t)(q * (4 / NQ))) & 3u;
                                const float a = acc[q][t].x, b2 = acc[q][t].y;
                                const float ni = (r & 1u) ? b2 : a;  /* r=1: I=-Q', r=3: I=Q' */
                                const float nq = (r & 1u) ? a : b2;  /* r=1: Q=I',  r=3: Q=-I' */
                                const uint32_t sI = (r == 1u || r == 2u) ? 0x80000000u : 0u;
                                const uint32_t sQ = (r == 2u || r == 3u) ? 0x80000000u : 0u;
                                Pp[pass + q * GR][t] = make_float2(__uint_as_float(__float_as_uint(ni) ^ sI),
                                                                   __uint_as_float(__float_as_uint(nq) ^ sQ));
                            }
                        }
                    }
                }
                /* ---- window tree (SPEC 3).  Each lane needs the S-1 tail values of the lane before it
                 * (lane 0: of the previous step, kept in `carry`).  With one tone group the PCM stage is
                 * dead by now and serves as the exchange buffer; otherwise the tails travel by shuffle.
                 * Once the last group has consumed the stage, the next step's PCM starts streaming in. */
                float2 pin[TG][S - 1];
                /* V[t][lv][i]: level-lv tree value of hop i.  Entries with i >= 2^lv - 1 depend on this lane
                 * only and are computed now (their top ones are the tails the next lane needs); the others
                 * follow after the exchange, from the previous lane's tails.  Every add has the operands and
                 * the order of SPEC 3's tree. */
                float2 V[TG][LV + 1][S];
                {
                    float2 tails[TG][S - 1];
#pragma unroll
                    for (int t = 0; t < TG; ++t) {
#pragma unroll
                        for (int i = 0; i < S; ++i) V[t][0][i] = Pp[i][t];
#pragma unroll
                        for (int lv = 1; lv <= LV; ++lv) {
                            const int d = 1 << (lv - 1);
#pragma unroll
                            for (int j = 0; j < d; ++j) tails[t][d - 1 + j] = V[t][lv - 1][S - d + j];
#pragma unroll
                            for (int i = (1 << lv) - 1; i < S; ++i)
                                V[t][lv][i] = (lv == 1) ? ffma2vv(V[t][0][(i & 1) ? i : i - 1], p.fold_sg[g * TG + t], V[t][0][(i & 1) ? i - 1 : i])
                                                        : fadd2(V[t][lv - 1][i - d], V[t][lv - 1][i]);
                        }
                    }
                    float2 *cg = carry + g * TG * (S - 1);
                    if (NG == 1) {
                        constexpr uint32_t TB = (uint32_t)TG * (S - 1) * 8u;         /* tail bytes per lane */
                        constexpr uint32_t XS = ((TB / 16u) & 1u) ? TB : TB + 16u;  /* odd multiple of 16: conflict-free */
                        static_assert(33u * XS <= stage_bytes<N>(), "exchange buffer exceeds the stage");
                        __syncwarp(); /* every lane is done reading PCM */
                        float4 *mine = reinterpret_cast<float4 *>(&tails[0][0]);
#pragma unroll
                        for (uint32_t k = 0; k < TB / 16u; ++k)
                            asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(stage + (uint32_t)(lane + 1) * XS + k * 16u),
                                         "f"(mine[k].x), "f"(mine[k].y), "f"(mine[k].z), "f"(mine[k].w) : "memory");
                        __syncwarp();
                        const uint32_t rsrc = (lane == 0) ? (uint32_t)__cvta_generic_to_shared(cg) : stage + (uint32_t)lane * XS;
                        float4 *pv = reinterpret_cast<float4 *>(&pin[0][0]);
#pragma unroll
                        for (uint32_t k = 0; k < TB / 16u; ++k)
                            asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(pv[k].x), "=f"(pv[k].y), "=f"(pv[k].z), "=f"(pv[k].w)
                                         : "r"(rsrc + k * 16u) : "memory");
                        __syncwarp(); /* exchange buffer and old carry consumed */
                    } else {
#pragma unroll
                        for (int t = 0; t < TG; ++t)
#pragma unroll
                            for (int k = 0; k < S - 1; ++k) {
                                const float2 c0 = (lane == 0) ? cg[t * (S - 1) + k] : make_float2(0.f, 0.f);
                                const float2 sh = shfl2(tails[t][k], (lane + 31) & 31);
                                pin[t][k] = (lane == 0) ? c0 : sh;
                            }
                        __syncwarp();
                    }
                    if (lane == nvalid - 1) {
#pragma unroll
                        for (int t = 0; t < TG; ++t)
#pragma unroll
                            for (int k = 0; k < S - 1; ++k) cg[t * (S - 1) + k] = tails[t][k];
                    }
                    if (g == NG - 1 && step + 1 < n_steps) issue(step + 1);
                }
#pragma unroll
                for (int t = 0; t < TG; ++t) {
                    const int tg = g * TG + t;
                    float2 L[S];
#pragma unroll
                    for (int lv = 1; lv <= LV; ++lv) {
                        const int d = 1 << (lv - 1);
#pragma unroll
                        for (int i = 0; i < (1 << lv) - 1 && i < S; ++i)
                            V[t][lv][i] = (lv == 1) ? ffma2vv(pin[t][0], p.fold_sg[g * TG + t], V[t][0][0]) /* the previous lane's hop S-1 is odd */
                                                    : fadd2((i >= d) ? V[t][lv - 1][(i >= d) ? i - d : 0] : pin[t][d - 1 + i], V[t][lv - 1][i]);
                    }
#pragma unroll
                    for (int i = 0; i < S; ++i) L[i] = V[t][LV][i];
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
            /* publish this step's hop records in the ring (slots of lanes past a ragged end keep their
             * older content: they are never addressed) */
            if (active) {
                const uint32_t a0 = sr + (((hic + (uint32_t)(lane * S)) & RM) << 3); /* S consecutive ring entries (linear layout, ring_off<S, false>) */
                if (S % 2 == 0) {
#pragma unroll
                    for (int i = 0; i < S; i += 2)
                        asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(a0 + (uint32_t)i * 8u), "r"(__float_as_uint(ec[i])), "r"(dc[i]),
                                     "r"(__float_as_uint(ec[(i + 1) % S])), "r"(dc[(i + 1) % S]) : "memory");
                } else {
#pragma unroll
                    for (int i = 0; i < S; ++i)
                        asm volatile("st.shared.v2.u32 [%0], {%1,%2};" ::"r"(a0 + (uint32_t)i * 8u), "r"(__float_as_uint(ec[i])), "r"(dc[i]) : "memory");
                }
            }
            __syncwarp();
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

            if (MODE == 0) sm_step<T, N, S, false>(p, ch, lane, sr, hic, nvalid, active, dc, (uint32_t)__cvta_generic_to_shared(ssc), crc_k, hop_base);
        }

        /* ---- save carried state: the last 32 symbol slots of the chunk ---- */
        __syncwarp();
#pragma unroll
        for (int i = 0; i < S; ++i) {
            const uint32_t idx = ((p.n_syms - 32u + (uint32_t)lane) * S + (uint32_t)i) & RM;
            uint2 rv;
            asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(rv.x), "=r"(rv.y) : "r"(sr + ring_off<S, false>(idx)) : "memory");
            grec[lane * S + i] = rv;
        }
        for (int i = lane; i < (S - 1) * T; i += 32) gcarry[i] = carry[i];
        if (MODE == 0) reinterpret_cast<uint32_t *>(stp)[lane] = reinterpret_cast<const uint32_t *>(ssc)[lane];
        __syncwarp();
        if (MODE == 0) {
            if (multi) { /* this channel's next chunk may start (the state above is visible before the counter) */
                __threadfence();
                __syncwarp();
                if (lane == 0) asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p.progress + ch), "r"(p.prog_base + chunk + 1u) : "memory");
            }
            /* every processed item takes one ticket (and, in a multi-chunk launch, every warp one more at the start), so a launch advances the
             * counter by exactly n_ch (n_ch * n_chunks + warps): the host knows the value it starts from (q_base) and nothing is reset between launches */
            uint32_t nx = 0;
            if (lane == 0) nx = atomicAdd(&p.counters[4], 1u);
            item = (multi ? 0u : total_warps) + (__shfl_sync(FULL, nx, 0) - p.q_base);
        } else {
            item += total_warps;
        }
    }
    if (MODE == 0 && lane == 0) publish_snapshot(p, total_warps);
}

/* fresh per-channel state: everything zero, hop records "before the stream" (d = 0xFF, emax = 0) */
__global__ void k_init_state(unsigned char *state, uint32_t state_stride, uint32_t n_ch, uint32_t n_rec) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t words = state_stride / 4u;
    if (i >= (size_t)n_ch * words) return;
    const uint32_t w = (uint32_t)(i % words);
    uint32_t v = 0u;
    if (w >= sizeof(ChanScalars) / 4u && w < sizeof(ChanScalars) / 4u + 2u * n_rec && (w & 1u)) v = 0xFFu;
    reinterpret_cast<uint32_t *>(state)[i] = v;
}

} /* namespace anm */
