/*
 * anm_kernels_tc.cuh -- dense tone sets (SPEC 3b, T >= 32): the windows-by-basis contraction on the
 * 5th-generation tensor cores (tcgen05.mma kind::i8, accumulators in TMEM).
 *
 * A CTA owns four channels; two warps serve each channel c = 4*blockIdx.x + (w & 3): within a step of
 * 32 symbol periods lane l owns symbol period l (the same ownership as k_demod, so the hop-record
 * ring, the carried state and the whole sync / slicing / framing state machine sm_step are shared).
 * Both warps of a channel read the same TMEM lanes (rows) and split the tones of every group; warp
 * c ("front") merges the two argmax candidates and runs the state machine while warp c + 4 ("back")
 * already loads and byte-splits the next step's PCM.
 *
 *   PCM (int16, HBM) --LDG.128, coalesced--> byte split (PRMT): high bytes (s8), low bytes (u8)
 *        --> A operand panels in shared memory, K-major, no swizzle: row = symbol period (128 rows =
 *            4 warps x 32 lanes), K = the H samples of one hop; one panel set per hop phase q and plane
 *   basis (int8, one panel set per hop phase) --> B operand panels, 16 tones (32 columns) per group
 *   D[q][plane] (128 x 32, s32, TMEM) = A[q][plane] . B^T        2 x tcgen05.mma (K = 32 each) per D
 *   hop partial = 256 * D[q][hi] + D[q][lo]                       exact integer (x = 256 hi + lo)
 *   window sums (exact integer adds, tails of the previous symbol period by shuffle / carry),
 *   E = fma(fI, fI, fQ fQ), argmax over tones.
 *
 * There is no reference kernel for this; behaviour is SPEC.md's.
 */
#pragma once
#include "anm_kernels.cuh"

namespace anm {
namespace tc {

constexpr uint32_t kRows = 128;                    /* MMA M: symbol periods per CTA step */
constexpr uint32_t kPanel = kRows * 16u + 16u;     /* one 16-byte K chunk of all rows; +16: spreads the panels over the banks */
constexpr int kTG = 16;                            /* tones per MMA group */
constexpr uint32_t kNcol = 2u * kTG;               /* MMA N: (cos, sin) columns of a group */
constexpr uint32_t kBPanel = kNcol * 16u;          /* one 16-byte K chunk of a group's basis rows */
constexpr uint32_t kTmemCols = 256;                /* 4 hop phases x 2 byte planes x 32 columns */

template <int N, int S>
__host__ __device__ constexpr uint32_t a_bytes() { return 2u * S * (uint32_t)(N / S / 16) * kPanel; }
/* basis panels [hop phase q][tone group][K chunk]: the basis of hop phase q is the first-quarter basis
 * rotated by (-j)^(bin q); keeping all S phases removes every rotation from the epilogue */
template <int T, int N, int S>
__host__ __device__ constexpr uint32_t b_bytes() { return (uint32_t)S * (uint32_t)(T / kTG) * (uint32_t)(N / S / 16) * kBPanel; }
template <int T, int S>
__host__ __device__ constexpr uint32_t warp_bytes() { return 64u * S * 8u + 128u + (uint32_t)(S - 1) * T * 8u; }
template <int T, int N, int S>
__host__ __device__ constexpr uint32_t smem_bytes() { return a_bytes<N, S>() + b_bytes<T, N, S>() + 4u * warp_bytes<T, S>() + 16u; }

/* shared-memory matrix descriptor: K-major, no swizzle; LBO = stride between the two 16-byte K chunks
 * of an MMA, SBO = stride between groups of 8 rows (cute::UMMA::SmemDescriptor, version 1) */
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((addr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo_bytes >> 4) << 16) | ((uint64_t)(sbo_bytes >> 4) << 32) | (1ull << 46);
}
/* instruction descriptor of kind::i8: D = s32, B = s8, A = s8 (a_signed) or u8, both K-major */
__host__ __device__ constexpr uint32_t idesc_i8(bool a_signed) {
    return (2u << 4) | ((a_signed ? 1u : 0u) << 7) | (1u << 10) | ((kNcol >> 3) << 17) | ((kRows >> 4) << 24);
}
__device__ __forceinline__ void mma_i8(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %6, %7, %8}, p;\n\t"
        "}\n" ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(0u), "r"(0u), "r"(0u), "r"(0u)
        : "memory");
}
__device__ __forceinline__ void mma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld4(uint32_t taddr, int32_t (&v)[4]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]) : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, int32_t (&v)[16]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
                   "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                 : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, int32_t (&v)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                 : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    do {
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                     : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    } while (!ok);
}
__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel) {
    uint32_t d;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(sel));
    return d;
}

} /* namespace tc */

/* MODE 0: streaming demodulator; MODE 1: stateless tone-energy pass (trace outputs). */
template <int T, int N, int S, int MODE>
__global__ void __launch_bounds__(256, 2) k_demod_tc(const __grid_constant__ KParams p) {
    using namespace tc;
    constexpr int H = N / S;
    constexpr int KC = H / 16;          /* 16-byte K chunks per hop */
    constexpr int KS = H / 32;          /* MMAs (K = 32) per hop */
    constexpr int NG = T / kTG;         /* tone groups */
    constexpr int CPS = N / 8;          /* 16-byte PCM chunks per symbol period */
    constexpr int TH = kTG / 2;         /* tones of a group per warp of the pair */
    constexpr int TN = 4;               /* tones per epilogue iteration */
    constexpr uint32_t RM = 64u * S - 1u;
    constexpr uint32_t FULL = 0xffffffffu;
    static_assert(S == 4 && (H % 32) == 0 && (T % kTG) == 0 && (TH % TN) == 0, "unsupported dense geometry");
    static_assert(2u * S * kNcol == kTmemCols, "TMEM column budget");

    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31;
    const int w = threadIdx.x >> 5;
    const int c4 = w & 3;               /* channel slot of the CTA = TMEM lane quadrant */
    const bool front = w < 4;           /* front: merge + state machine; back: next step's PCM */
    const uint32_t sA = (uint32_t)__cvta_generic_to_shared(smem_raw);
    const uint32_t sB = sA + a_bytes<N, S>();
    unsigned char *wsm = smem_raw + a_bytes<N, S>() + b_bytes<T, N, S>() + (size_t)c4 * warp_bytes<T, S>();
    const uint32_t sr = (uint32_t)__cvta_generic_to_shared(wsm); /* HopRec ring [64*S] of the channel */
    ChanScalars *ssc = reinterpret_cast<ChanScalars *>(wsm + 64u * S * 8u);
    int2 *carry = reinterpret_cast<int2 *>(wsm + 64u * S * 8u + 128u); /* [T][S-1] suffix sums of the last symbol period */
    unsigned char *tail = smem_raw + a_bytes<N, S>() + b_bytes<T, N, S>() + 4u * warp_bytes<T, S>();
    const uint32_t mbar = (uint32_t)__cvta_generic_to_shared(tail);
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(tail + 8);

    const uint32_t crc_k = (MODE == 0) ? (uint32_t)p.crc_pow[lane] : 0u;
    const uint32_t n_steps = (p.n_syms + 31u) / 32u;
    const uint32_t ch = blockIdx.x * 4u + (uint32_t)c4;
    const bool have_ch = ch < p.n_ch;
    unsigned char *stp = p.state + (size_t)(have_ch ? ch : 0u) * p.state_stride;
    uint2 *grec = reinterpret_cast<uint2 *>(stp + sizeof(ChanScalars));
    int2 *gcarry = reinterpret_cast<int2 *>(stp + state_carry_offset<T, S>());
    const char *src = reinterpret_cast<const char *>(p.pcm + (size_t)(have_ch ? ch : 0u) * p.ch_stride);

    /* PCM of one step -> byte planes in the A panels (rows 32*c4 ..).  All loads are issued before the
     * first split so that the DRAM / L2 latency is paid once; the following step is pulled into L2. */
    auto load_step = [&](uint32_t step) {
        const int nv = (int)min(32u, p.n_syms - step * 32u);
        const char *g = src + (size_t)step * (32u * N * 2u) + (size_t)lane * 16u;
        constexpr int IT = CPS; /* CPS * 32 chunks of 16 bytes per step and channel, 32 per instruction */
        static_assert(CPS % 32 == 0 || 32 % CPS == 0, "chunk geometry");
        constexpr int BATCH = 16;
        static_assert(IT % BATCH == 0, "load batch");
#pragma unroll 1
        for (int b0 = 0; b0 < IT; b0 += BATCH) {
            uint4 v[BATCH];
#pragma unroll
            for (int j = 0; j < BATCH; ++j) {
                const uint32_t idx = (uint32_t)(b0 + j) * 32u + (uint32_t)lane;
                v[j] = make_uint4(0u, 0u, 0u, 0u);
                if ((int)(idx / (uint32_t)CPS) < nv) v[j] = __ldg(reinterpret_cast<const uint4 *>(g + (size_t)(b0 + j) * 512u));
            }
            if (b0 == 0 && step + 1 < n_steps) {
                /* next step: 32 * N * 2 bytes per channel = N / 2 lines of 128 bytes, N / 64 per lane */
                const char *nx = src + (size_t)(step + 1) * (32u * N * 2u) + (size_t)lane * 128u;
#pragma unroll
                for (int j = 0; j < N / 64; ++j) asm volatile("prefetch.global.L2 [%0];" ::"l"(nx + (size_t)j * 4096u));
            }
#pragma unroll
            for (int j = 0; j < BATCH; ++j) {
                const uint32_t idx = (uint32_t)(b0 + j) * 32u + (uint32_t)lane;
                const uint32_t r = idx / (uint32_t)CPS, c = idx % (uint32_t)CPS; /* symbol period in the step, chunk in it */
                const uint32_t q = c / (uint32_t)(H / 8), hc = c % (uint32_t)(H / 8);
                const uint32_t off = (q * KC + (hc >> 1)) * kPanel + ((uint32_t)(32 * c4) + r) * 16u + (hc & 1u) * 8u;
                const uint32_t lo0 = prmt(v[j].x, v[j].y, 0x6420u), lo1 = prmt(v[j].z, v[j].w, 0x6420u);
                const uint32_t hi0 = prmt(v[j].x, v[j].y, 0x7531u), hi1 = prmt(v[j].z, v[j].w, 0x7531u);
                asm volatile("st.shared.v2.u32 [%0], {%1,%2};" ::"r"(sA + off), "r"(hi0), "r"(hi1) : "memory");
                asm volatile("st.shared.v2.u32 [%0], {%1,%2};" ::"r"(sA + (uint32_t)(S * KC) * kPanel + off), "r"(lo0), "r"(lo1) : "memory");
            }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); /* panels -> visible to the MMA's async proxy */
    };

    /* ---- one-time setup: basis panels, mbarrier, TMEM, carried state, first step's PCM ---- */
    {
        const uint4 *gsrc = reinterpret_cast<const uint4 *>(p.tc_basis);
        uint4 *dst = reinterpret_cast<uint4 *>(smem_raw + a_bytes<N, S>());
        for (uint32_t i = threadIdx.x; i < b_bytes<T, N, S>() / 16u; i += blockDim.x) dst[i] = __ldg(&gsrc[i]);
    }
    if (threadIdx.x == 0) {
        mbar_init(mbar, 1u);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (w == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(tmem_slot)), "r"(kTmemCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (have_ch && front) {
#pragma unroll
        for (int i = 0; i < S; ++i) {
            const uint2 rv = grec[lane * S + i];
            asm volatile("st.shared.v2.u32 [%0], {%1,%2};" ::"r"(sr + (uint32_t)((32 + lane) * S + i) * 8u), "r"(rv.x), "r"(rv.y) : "memory");
        }
        for (int i = lane; i < (S - 1) * T; i += 32) carry[i] = gcarry[i];
        if (MODE == 0) reinterpret_cast<uint32_t *>(ssc)[lane] = reinterpret_cast<const uint32_t *>(stp)[lane];
    }
    if (have_ch && !front && n_steps) load_step(0);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); /* basis panels */
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t tmem_lane = tmem_base + ((uint32_t)(32 * c4) << 16) + (front ? 0u : (uint32_t)(2 * TH)); /* this warp's columns of every accumulator */
    uint32_t mph = 0;

#pragma unroll 1
    for (uint32_t step = 0; step < n_steps; ++step) {
        const int nvalid = (int)min(32u, p.n_syms - step * 32u);
        const bool active = have_ch && lane < nvalid;
        const uint32_t hic = step * 32u * S;

        uint32_t dc[S];
        float ec[S];
#pragma unroll
        for (int i = 0; i < S; ++i) { dc[i] = 0xFFu; ec[i] = 0.0f; }

#pragma unroll 1
        for (int g = 0; g < NG; ++g) {
            /* ---- contraction of tone group g: D[q][plane] = A[q][plane] . B[q][g]^T (one thread of a back warp) ---- */
            if (w == 4 + (g & 3) && lane == 0) {
                tc_fence_after();
                /* descriptors differ only in their start-address field: base + a compile-time offset */
                const uint64_t a0 = smem_desc(sA, kPanel, 128u);
                const uint64_t b0 = smem_desc(sB + (uint32_t)(g * KC) * kBPanel, kBPanel, 128u);
#pragma unroll
                for (int q = 0; q < S; ++q)
#pragma unroll
                    for (int pl = 0; pl < 2; ++pl)
#pragma unroll
                        for (int ks = 0; ks < KS; ++ks)
                            mma_i8(tmem_base + (uint32_t)(q * 2 + pl) * kNcol, a0 + (uint64_t)(((uint32_t)((pl * S + q) * KC + 2 * ks) * kPanel) >> 4),
                                   b0 + (uint64_t)(((uint32_t)(q * NG * KC + 2 * ks) * kBPanel) >> 4), idesc_i8(pl == 0), ks > 0 ? 1u : 0u);
                mma_commit(mbar);
            }
            mbar_wait(mbar, mph);
            mph ^= 1u;
            tc_fence_after();

            /* ---- epilogue: this warp's TH tones of the group, TN at a time ---- */
            if (have_ch) {
#pragma unroll 1
                for (int tb = 0; tb < TH / TN; ++tb) {
                    int32_t v[S][2][2 * TN];
#pragma unroll
                    for (int q = 0; q < S; ++q)
#pragma unroll
                        for (int pl = 0; pl < 2; ++pl) tmem_ld8(tmem_lane + (uint32_t)((q * 2 + pl) * (int)kNcol + 2 * TN * tb), v[q][pl]);
                    tmem_ld_wait();
                    const int tone0 = g * kTG + (front ? 0 : TH) + tb * TN;
                    /* hop partials, their suffix sums (hops i..S-1) and the window sums:
                     * W_i = (suffix sum of the previous symbol period from hop i+1) + (prefix sum to hop i) */
                    int32_t cI[TN][S - 1], cQ[TN][S - 1]; /* this lane's suffix sums, next step's carry */
#pragma unroll
                    for (int tt = 0; tt < TN; ++tt) {
                        int32_t PI[S], PQ[S];
#pragma unroll
                        for (int q = 0; q < S; ++q) {
                            PI[q] = v[q][0][2 * tt] * 256 + v[q][1][2 * tt];
                            PQ[q] = v[q][0][2 * tt + 1] * 256 + v[q][1][2 * tt + 1];
                        }
                        cI[tt][S - 2] = PI[S - 1];
                        cQ[tt][S - 2] = PQ[S - 1];
#pragma unroll
                        for (int i = S - 2; i >= 1; --i) { cI[tt][i - 1] = PI[i] + cI[tt][i]; cQ[tt][i - 1] = PQ[i] + cQ[tt][i]; }
                        int32_t fI = 0, fQ = 0;
#pragma unroll
                        for (int i = 0; i < S; ++i) {
                            fI += PI[i];
                            fQ += PQ[i];
                            int32_t wI = fI, wQ = fQ;
                            if (i < S - 1) {
                                int32_t pI = __shfl_up_sync(FULL, cI[tt][i], 1), pQ = __shfl_up_sync(FULL, cQ[tt][i], 1);
                                if (lane == 0) { const int2 cv = carry[(tone0 + tt) * (S - 1) + i]; pI = cv.x; pQ = cv.y; }
                                wI += pI;
                                wQ += pQ;
                            }
                            const float xI = (float)wI, xQ = (float)wQ;
                            const float E = __fmaf_rn(xI, xI, __fmul_rn(xQ, xQ));
                            if (MODE == 1) {
                                if (p.trE && active) {
                                    const size_t hop = ((size_t)step * 32 + lane) * S + i;
                                    p.trE[((size_t)ch * p.tr_hops + hop) * T + tone0 + tt] = E;
                                }
                            }
                            if ((g == 0 && tb == 0 && tt == 0) || E > ec[i]) { ec[i] = E; dc[i] = (uint32_t)(tone0 + tt); }
                        }
                    }
                    __syncwarp(); /* lane 0 has read the old carry */
                    if (lane == nvalid - 1) {
#pragma unroll
                        for (int tt = 0; tt < TN; ++tt)
#pragma unroll
                            for (int i = 0; i < S - 1; ++i) carry[(tone0 + tt) * (S - 1) + i] = make_int2(cI[tt][i], cQ[tt][i]);
                    }
                }
            }
            /* TMEM reads of this group done before the next group's MMAs overwrite the accumulators */
            tc_fence_before();
            __syncthreads();
        }

        /* ---- the back warp hands its argmax candidates to the front warp through the ring slots of this
         * step (their old content, two steps back, is dead) and moves on to the next step's PCM ---- */
        const uint32_t a0 = sr + (((hic + (uint32_t)(lane * S)) & RM) << 3);
        if (have_ch && !front && active) {
#pragma unroll
            for (int i = 0; i < S; i += 2)
                asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(a0 + (uint32_t)i * 8u), "r"(__float_as_uint(ec[i])), "r"(dc[i]),
                             "r"(__float_as_uint(ec[i + 1])), "r"(dc[i + 1]) : "memory");
        }
        __syncthreads();
        if (have_ch && front) {
            if (active) {
#pragma unroll
                for (int i = 0; i < S; i += 2) {
                    uint32_t e0, d0, e1, d1;
                    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(e0), "=r"(d0), "=r"(e1), "=r"(d1) : "r"(a0 + (uint32_t)i * 8u) : "memory");
                    /* lowest tone index wins a tie (SPEC 3): the back warp's tones of a group are the higher ones,
                     * but a later group of the front warp is higher still */
                    const float f0 = __uint_as_float(e0), f1 = __uint_as_float(e1);
                    if (f0 > ec[i] || (f0 == ec[i] && d0 < dc[i])) { ec[i] = f0; dc[i] = d0; }
                    if (f1 > ec[i + 1] || (f1 == ec[i + 1] && d1 < dc[i + 1])) { ec[i + 1] = f1; dc[i + 1] = d1; }
                }
#pragma unroll
                for (int i = 0; i < S; i += 2)
                    asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(a0 + (uint32_t)i * 8u), "r"(__float_as_uint(ec[i])), "r"(dc[i]),
                                 "r"(__float_as_uint(ec[i + 1])), "r"(dc[i + 1]) : "memory");
            } else {
#pragma unroll
                for (int i = 0; i < S; ++i) { dc[i] = 0xFFu; ec[i] = 0.0f; }
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
            if (MODE == 0) sm_step<T, N, S>(p, ch, lane, sr, hic, nvalid, active, dc, (uint32_t)__cvta_generic_to_shared(ssc), crc_k);
        }
        if (have_ch && !front && step + 1 < n_steps) load_step(step + 1);
        __syncthreads(); /* next step's panels complete; the ring is the state machine's again */
    }

    /* ---- save carried state ---- */
    if (have_ch && front) {
        __syncwarp();
#pragma unroll
        for (int i = 0; i < S; ++i) {
            const uint32_t idx = ((p.n_syms - 32u + (uint32_t)lane) * S + (uint32_t)i) & RM;
            uint2 rv;
            asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(rv.x), "=r"(rv.y) : "r"(sr + idx * 8u) : "memory");
            grec[lane * S + i] = rv;
        }
        for (int i = lane; i < (S - 1) * T; i += 32) gcarry[i] = carry[i];
        if (MODE == 0) reinterpret_cast<uint32_t *>(stp)[lane] = reinterpret_cast<const uint32_t *>(ssc)[lane];
    }
    tc_fence_before();
    __syncthreads();
    if (w == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
}

} /* namespace anm */
