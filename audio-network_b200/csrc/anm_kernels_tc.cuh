/*
 * anm_kernels_tc.cuh -- dense tone sets (SPEC 3b, T >= 32): the windows-by-basis contraction on the
 * 5th-generation tensor cores (tcgen05.mma kind::i8, accumulators in TMEM).
 *
 * A CTA owns four channels and steps through them 32 symbol periods at a time.  The MMA rows are
 * interleaved, row = 4 * symbol period + channel, so that "the previous symbol period of the same
 * channel" is always four rows up: starting an A descriptor 64 bytes earlier shifts every row by one
 * symbol period (the four rows in front of each panel hold the last symbol period of the previous
 * step).  That puts the sliding window itself into the contraction:
 *
 *   PCM (int16, HBM) --LDG.128, coalesced--> byte split (PRMT): high bytes (s8), low bytes (u8)
 *        --> A operand panels in shared memory, K-major, no swizzle, one panel per 16-sample K chunk,
 *            per hop j of the symbol period and per byte plane
 *   basis (int8, one panel set per hop phase) --> B operand panels, 32 tones (64 columns) per group
 *   W[i][plane] (128 x 64, s32, TMEM) = sum_{j<=i} A[j][plane] . B[j]^T  +  sum_{j>i} A_prev[j][plane] . B[j]^T
 *        window ending with hop i of every symbol period: 8 x tcgen05.mma (K = 32) per W and plane.  A round
 *        is one window of one tone group (2 planes x 64 TMEM columns); two accumulator sets alternate and a
 *        ninth warp does nothing but issue, so the contraction of round r+1 runs under the epilogue of
 *        round r (full: tcgen05.commit -> mbarrier; empty: one arrival per worker warp)
 *   W = 256 * W[hi] + W[lo] (exact integer, x = 256 hi + lo), E = fma(fI, fI, fQ fQ), argmax over tones.
 *
 * Two worker warps share each TMEM lane quadrant and split the tones of every group.  The epilogue lanes
 * therefore serve (symbol period, channel) pairs, not one channel per warp; hop records go to the
 * per-channel rings, and after the step front warp c reads channel c's decisions back, lane = symbol
 * period, and runs the shared state machine (sm_step) while back warp c already loads and byte-splits
 * channel c's next 32 symbol periods.
 *
 * There is no reference kernel for this; behaviour is SPEC.md's.
 */
#pragma once
#include "anm_kernels.cuh"

namespace anm {
namespace tc {

constexpr uint32_t kRows = 128;                    /* MMA M: 4 channels x 32 symbol periods, row = 4 * symbol period + channel */
constexpr uint32_t kCarryRows = 4;                 /* the last symbol period of the previous step, one row per channel */
constexpr uint32_t kPanel = (kCarryRows + kRows) * 16u + 16u; /* one 16-byte K chunk of all rows; +16: spreads the panels over the banks */
constexpr int kTG = 32;                            /* tones per MMA group */
constexpr uint32_t kNcol = 2u * kTG;               /* MMA N: (cos, sin) columns of a group */
constexpr uint32_t kBPanel = kNcol * 16u;          /* one 16-byte K chunk of a group's basis rows */
constexpr uint32_t kTmemBuf = 2u * kNcol;          /* one accumulator set: 2 byte planes x 64 columns */
constexpr uint32_t kTmemCols = 2u * kTmemBuf;      /* two sets: the contraction of round r+1 runs under the epilogue of round r */
constexpr int kWorkerWarps = 8;                    /* two per TMEM lane quadrant; warp 8 only issues the MMAs */

template <int N, int S>
__host__ __device__ constexpr uint32_t a_bytes() { return 2u * S * (uint32_t)(N / S / 16) * kPanel; }
/* basis panels [hop phase q][tone group][K chunk]: the basis of hop phase q is the first-quarter basis
 * rotated by (-j)^(bin q); keeping all S phases removes every rotation from the epilogue */
template <int T, int N, int S>
__host__ __device__ constexpr uint32_t b_bytes() { return (uint32_t)S * (uint32_t)(T / kTG) * (uint32_t)(N / S / 16) * kBPanel; }
template <int T, int S>
__host__ __device__ constexpr uint32_t warp_bytes() { return 64u * S * 8u + 128u; } /* per channel: HopRec ring | scalars */
template <int T, int N, int S>
__host__ __device__ constexpr uint32_t smem_bytes() { return a_bytes<N, S>() + b_bytes<T, N, S>() + 4u * warp_bytes<T, S>() + 48u; }

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
__device__ __forceinline__ uint32_t idesc_i8_n(bool a_signed, uint32_t ncol) {
    return (2u << 4) | ((a_signed ? 1u : 0u) << 7) | (1u << 10) | ((ncol >> 3) << 17) | ((kRows >> 4) << 24);
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
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    do {
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                     : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    } while (!ok);
}
/* one lane of a converged warp; unlike `lane == 0` the compiler knows the elected predicate is safe for
 * the warp-uniform tcgen05 instructions and does not wrap each of them in an election loop */
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0u;
}
__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel) {
    uint32_t d;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(sel));
    return d;
}

} /* namespace tc */

/* MODE 0: streaming demodulator; MODE 1: stateless tone-energy pass (trace outputs). */
template <int T, int N, int S, int MODE>
__global__ void __launch_bounds__(288, 2) k_demod_tc(const __grid_constant__ KParams p) {
    using namespace tc;
    constexpr int H = N / S;
    constexpr int KC = H / 16;          /* 16-byte K chunks per hop */
    constexpr int KS = H / 32;          /* MMAs (K = 32) per hop */
    constexpr int NG = T / kTG;         /* tone groups */
    constexpr int CPS = N / 8;          /* 16-byte PCM chunks per symbol period */
    constexpr int TH = kTG / 2;         /* tones of a group per warp of the pair */
    constexpr int TN = 8;               /* tones per epilogue iteration */
    constexpr int R = NG * S;           /* rounds per step: tone group x window */
    constexpr uint32_t RM = 64u * S - 1u;
    constexpr uint32_t CUR = kCarryRows * 16u; /* byte offset of the current rows inside a panel */
    static_assert(S == 4 && (H % 32) == 0 && (T % kTG) == 0 && (TH % TN) == 0 && (R % 2) == 0, "unsupported dense geometry");
    static_assert(2u * kNcol == kTmemBuf, "TMEM column budget");
    static_assert(2u * S * KC * 16u <= (uint32_t)(S - 1) * T * 8u, "carry rows must fit the state's carry area");

    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31;
    const int w = threadIdx.x >> 5;
    const bool issuer = w == kWorkerWarps;
    const int c4 = w & 3;               /* TMEM lane quadrant; also the channel slot this warp loads / runs the state machine for */
    const bool front = w < 4;           /* front: merge + state machine; back: next step's PCM */
    const bool back = w >= 4 && !issuer;
    const uint32_t sA = (uint32_t)__cvta_generic_to_shared(smem_raw);
    const uint32_t sB = sA + a_bytes<N, S>();
    unsigned char *chsm = smem_raw + a_bytes<N, S>() + b_bytes<T, N, S>(); /* per channel: HopRec ring [64*S] | ChanScalars */
    const uint32_t sr0 = (uint32_t)__cvta_generic_to_shared(chsm);
    const uint32_t sr = sr0 + (uint32_t)c4 * warp_bytes<T, S>();             /* ring of channel slot c4 */
    ChanScalars *ssc = reinterpret_cast<ChanScalars *>(chsm + (size_t)c4 * warp_bytes<T, S>() + 64u * S * 8u);
    unsigned char *tail = chsm + 4u * warp_bytes<T, S>();
    const uint32_t bar_full = (uint32_t)__cvta_generic_to_shared(tail); /* [2]: accumulator set written (tcgen05.commit) */
    const uint32_t bar_empty = bar_full + 16u;                          /* [2]: accumulator set read by all worker warps */
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(tail + 32);

    const uint32_t crc_k = (MODE == 0) ? (uint32_t)p.crc_pow[lane] : 0u;
    const uint32_t n_steps = (p.n_syms + 31u) / 32u;
    /* role 1: this warp's own channel (loads, carried state, state machine) */
    const uint32_t ch = blockIdx.x * 4u + (uint32_t)c4;
    const bool have_ch = ch < p.n_ch && !issuer;
    unsigned char *stp = p.state + (size_t)(have_ch ? ch : 0u) * p.state_stride;
    uint2 *grec = reinterpret_cast<uint2 *>(stp + sizeof(ChanScalars));
    uint4 *gcarry = reinterpret_cast<uint4 *>(stp + state_carry_offset<T, S>()); /* carry rows: [plane][hop][K chunk] x 16 bytes */
    const char *src = reinterpret_cast<const char *>(p.pcm + (size_t)(have_ch ? ch : 0u) * p.ch_stride);
    /* role 2: the (symbol period, channel) pair of this lane's TMEM row 32 * c4 + lane = 4 * esp + ec4 */
    const int esp = 8 * c4 + (lane >> 2);
    const int ec4 = lane & 3;
    const uint32_t ech = blockIdx.x * 4u + (uint32_t)ec4;
    const bool ehave = ech < p.n_ch;
    const uint32_t esr = sr0 + (uint32_t)ec4 * warp_bytes<T, S>();

    /* PCM of one step of this warp's channel -> byte planes in the A panels (rows 4 * sp + c4).  All loads
     * are issued before the first split so that the DRAM / L2 latency is paid once; the following step
     * is pulled into L2. */
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
                const uint32_t off = (q * KC + (hc >> 1)) * kPanel + CUR + (4u * r + (uint32_t)c4) * 16u + (hc & 1u) * 8u;
                const uint32_t lo0 = prmt(v[j].x, v[j].y, 0x6420u), lo1 = prmt(v[j].z, v[j].w, 0x6420u);
                const uint32_t hi0 = prmt(v[j].x, v[j].y, 0x7531u), hi1 = prmt(v[j].z, v[j].w, 0x7531u);
                asm volatile("st.shared.v2.u32 [%0], {%1,%2};" ::"r"(sA + off), "r"(hi0), "r"(hi1) : "memory");
                asm volatile("st.shared.v2.u32 [%0], {%1,%2};" ::"r"(sA + (uint32_t)(S * KC) * kPanel + off), "r"(lo0), "r"(lo1) : "memory");
            }
        }
    };

    /* ---- one-time setup: basis panels, mbarriers, TMEM, carried state, first step's PCM ---- */
    {
        const uint4 *gsrc = reinterpret_cast<const uint4 *>(p.tc_basis);
        uint4 *dst = reinterpret_cast<uint4 *>(smem_raw + a_bytes<N, S>());
        for (uint32_t i = threadIdx.x; i < b_bytes<T, N, S>() / 16u; i += blockDim.x) dst[i] = __ldg(&gsrc[i]);
    }
    if (threadIdx.x == 0) {
        mbar_init(bar_full, 1u);
        mbar_init(bar_full + 8u, 1u);
        mbar_init(bar_empty, (uint32_t)kWorkerWarps);
        mbar_init(bar_empty + 8u, (uint32_t)kWorkerWarps);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (w == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(tmem_slot)), "r"(kTmemCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (front && have_ch) {
#pragma unroll
        for (int i = 0; i < S; ++i) {
            const uint2 rv = grec[lane * S + i];
            asm volatile("st.shared.v2.u32 [%0], {%1,%2};" ::"r"(sr + (uint32_t)((32 + lane) * S + i) * 8u), "r"(rv.x), "r"(rv.y) : "memory");
        }
        if (MODE == 0) reinterpret_cast<uint32_t *>(ssc)[lane] = reinterpret_cast<const uint32_t *>(stp)[lane];
    }
    if (back) {
        /* carry row of this channel in every panel: lane = panel (2 planes x S hops x KC chunks = 32) */
        static_assert(2 * S * KC == 32, "one panel per lane");
        uint4 cv = make_uint4(0u, 0u, 0u, 0u);
        if (have_ch) cv = gcarry[lane];
        asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(sA + (uint32_t)lane * kPanel + (uint32_t)c4 * 16u), "r"(cv.x), "r"(cv.y), "r"(cv.z), "r"(cv.w) : "memory");
        if (have_ch && n_steps) load_step(0);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); /* panels -> visible to the MMA's async proxy */
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t tmem_lane = tmem_base + ((uint32_t)(32 * c4) << 16) + (front ? 0u : (uint32_t)(2 * TH)); /* this warp's columns of every accumulator */
    uint32_t ph_full[2] = {0u, 0u};  /* workers: parity of the next completion of full[b] */
    uint32_t ph_empty[2] = {1u, 1u}; /* issuer: parity to wait for on empty[b]; the first use of a set passes at once */

#pragma unroll 1
    for (uint32_t step = 0; step < n_steps; ++step) {
        const int nvalid = (int)min(32u, p.n_syms - step * 32u);
        const bool eactive = ehave && esp < nvalid && !issuer;
        const uint32_t hic = step * 32u * S;

        uint32_t dc[S];
        float ec[S];
#pragma unroll
        for (int i = 0; i < S; ++i) { dc[i] = 0xFFu; ec[i] = 0.0f; }

        if (issuer) {
            /* ---- the contractions of the step: round rnd = (tone group g, window i) into accumulator set rnd & 1,
             * W[plane] = sum over the S hops ending with hop i ---- */
            const uint64_t a0 = smem_desc(sA, kPanel, 128u);
#pragma unroll 1
            for (int g = 0; g < NG; ++g) {
                const uint64_t b0 = smem_desc(sB + (uint32_t)(g * KC) * kBPanel, kBPanel, 128u);
#pragma unroll
                for (int i = 0; i < S; ++i) {
                    const int b = i & 1; /* S is even: set = rnd & 1 = i & 1 */
                    mbar_wait(bar_empty + 8u * b, ph_empty[b]);
                    ph_empty[b] ^= 1u;
                    tc_fence_after();
                    const uint32_t d0 = tmem_base + (uint32_t)b * kTmemBuf;
                    if (elect_one()) {
#pragma unroll
                        for (int pl = 0; pl < 2; ++pl)
#pragma unroll
                            for (int j = 0; j < S; ++j)
#pragma unroll
                                for (int ks = 0; ks < KS; ++ks) {
                                    /* hops up to the window's last one come from this symbol period, later ones from the
                                     * previous symbol period of the same channel: four rows (64 bytes) up */
                                    const uint32_t aoff = (uint32_t)((pl * S + j) * KC + 2 * ks) * kPanel;
                                    const uint64_t ad = a0 + (uint64_t)(aoff >> 4) + (uint64_t)((j <= i) ? (CUR >> 4) : 0u);
                                    const uint64_t bd = b0 + (uint64_t)(((uint32_t)(j * NG * KC + 2 * ks) * kBPanel) >> 4);
                                    mma_i8(d0 + (uint32_t)pl * kNcol, ad, bd, idesc_i8_n(pl == 0, kNcol), (j > 0 || ks > 0) ? 1u : 0u);
                                }
                        mma_commit(bar_full + 8u * b);
                    }
                    __syncwarp();
                }
            }
            __syncwarp();
        } else {
#pragma unroll 1
            for (int g = 0; g < NG; ++g) {
#pragma unroll
                for (int i = 0; i < S; ++i) {
                    const int b = i & 1;
                    mbar_wait(bar_full + 8u * b, ph_full[b]);
                    ph_full[b] ^= 1u;
                    tc_fence_after();
                    /* ---- epilogue of window i, tone group g: this warp's TH tones, TN at a time ---- */
#pragma unroll 1
                    for (int tb = 0; tb < TH / TN; ++tb) {
                        int32_t v[2][2 * TN];
#pragma unroll
                        for (int pl = 0; pl < 2; ++pl) tmem_ld16(tmem_lane + (uint32_t)b * kTmemBuf + (uint32_t)(pl * (int)kNcol + 2 * TN * tb), v[pl]);
                        tmem_ld_wait();
                        const int tone0 = g * kTG + (front ? 0 : TH) + tb * TN;
#pragma unroll
                        for (int tt = 0; tt < TN; ++tt) {
                            const float xI = (float)(v[0][2 * tt] * 256 + v[1][2 * tt]);
                            const float xQ = (float)(v[0][2 * tt + 1] * 256 + v[1][2 * tt + 1]);
                            const float E = __fmaf_rn(xI, xI, __fmul_rn(xQ, xQ));
                            if (MODE == 1) {
                                if (p.trE && eactive) {
                                    const size_t hop = ((size_t)step * 32 + esp) * S + i;
                                    p.trE[((size_t)ech * p.tr_hops + hop) * T + tone0 + tt] = E;
                                }
                            }
                            if ((g == 0 && tb == 0 && tt == 0) || E > ec[i]) { ec[i] = E; dc[i] = (uint32_t)(tone0 + tt); }
                        }
                    }
                    /* this warp's TMEM reads of the set are complete (tcgen05.wait::ld): hand it back */
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(bar_empty + 8u * b);
                }
            }
        }
        /* ---- a worker that has waited for the last round knows every contraction of the step is complete: the
         * A panels are free.  The back warps move this step's last symbol period into the carry rows, hand their
         * argmax candidates to their front warp through the ring slots of the step (their old content, two
         * steps back, was last read by the previous step's state machine, which every front warp has left by
         * now: no worker is more than two rounds ahead of another) and load the next step's PCM ---- */
        const uint32_t a0r = esr + (((hic + (uint32_t)(esp * S)) & RM) << 3);
        if (back) {
            {
                uint4 cv;
                const uint32_t pa = sA + (uint32_t)lane * kPanel;
                asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(cv.x), "=r"(cv.y), "=r"(cv.z), "=r"(cv.w)
                             : "r"(pa + CUR + (uint32_t)(4 * (nvalid - 1) + c4) * 16u) : "memory");
                asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(pa + (uint32_t)c4 * 16u), "r"(cv.x), "r"(cv.y), "r"(cv.z), "r"(cv.w) : "memory");
            }
            if (eactive) {
#pragma unroll
                for (int i = 0; i < S; i += 2)
                    asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(a0r + (uint32_t)i * 8u), "r"(__float_as_uint(ec[i])), "r"(dc[i]),
                                 "r"(__float_as_uint(ec[i + 1])), "r"(dc[i + 1]) : "memory");
            }
            asm volatile("bar.arrive %0, 64;" ::"r"(3 + c4) : "memory"); /* candidates of this quadrant are in the rings */
        }
        if (front) {
            asm volatile("bar.sync %0, 64;" ::"r"(3 + c4) : "memory");
            if (eactive) {
#pragma unroll
                for (int i = 0; i < S; i += 2) {
                    uint32_t e0, d0, e1, d1;
                    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(e0), "=r"(d0), "=r"(e1), "=r"(d1) : "r"(a0r + (uint32_t)i * 8u) : "memory");
                    /* lowest tone index wins a tie (SPEC 3): the back warp's tones of a group are the higher ones,
                     * but a later group of the front warp is higher still */
                    const float f0 = __uint_as_float(e0), f1 = __uint_as_float(e1);
                    if (f0 > ec[i] || (f0 == ec[i] && d0 < dc[i])) { ec[i] = f0; dc[i] = d0; }
                    if (f1 > ec[i + 1] || (f1 == ec[i + 1] && d1 < dc[i + 1])) { ec[i + 1] = f1; dc[i + 1] = d1; }
                }
#pragma unroll
                for (int i = 0; i < S; i += 2)
                    asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(a0r + (uint32_t)i * 8u), "r"(__float_as_uint(ec[i])), "r"(dc[i]),
                                 "r"(__float_as_uint(ec[i + 1])), "r"(dc[i + 1]) : "memory");
                if (MODE == 1) {
                    if (p.trD) {
#pragma unroll
                        for (int i = 0; i < S; ++i) p.trD[(size_t)ech * p.tr_hops + ((size_t)step * 32 + esp) * S + i] = (uint8_t)dc[i];
                    }
                    if (p.trEmax) {
#pragma unroll
                        for (int i = 0; i < S; ++i) p.trEmax[(size_t)ech * p.tr_hops + ((size_t)step * 32 + esp) * S + i] = ec[i];
                    }
                }
            }
            /* every front warp has published its rows: front warp c now owns channel c, lane = symbol period */
            asm volatile("bar.sync 1, 128;" ::: "memory");
            if (MODE == 0 && have_ch) {
                const bool active = lane < nvalid;
                uint32_t mydc[S];
#pragma unroll
                for (int i = 0; i < S; ++i) mydc[i] = 0xFFu;
                if (active) {
                    const uint32_t a1 = sr + (((hic + (uint32_t)(lane * S)) & RM) << 3);
#pragma unroll
                    for (int i = 0; i < S; i += 2) {
                        uint32_t e0, d0, e1, d1;
                        asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(e0), "=r"(d0), "=r"(e1), "=r"(d1) : "r"(a1 + (uint32_t)i * 8u) : "memory");
                        mydc[i] = d0;
                        mydc[i + 1] = d1;
                    }
                }
                sm_step<T, N, S>(p, ch, lane, sr, hic, nvalid, active, mydc, (uint32_t)__cvta_generic_to_shared(ssc), crc_k);
            }
        } else if (back) {
            if (have_ch && step + 1 < n_steps) load_step(step + 1);
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); /* carry rows and panels -> async proxy */
        }
        /* next step's panels complete: only the loaders and the issuer meet here; the front warps join the
         * next step's rounds whenever their state machine is done (full / empty mbarriers keep them in step) */
        if (!front) asm volatile("bar.sync 2, 160;" ::: "memory");
    }
    __syncthreads();

    /* ---- save carried state ---- */
    if (have_ch) {
        if (front) {
            __syncwarp();
#pragma unroll
            for (int i = 0; i < S; ++i) {
                const uint32_t idx = ((p.n_syms - 32u + (uint32_t)lane) * S + (uint32_t)i) & RM;
                uint2 rv;
                asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(rv.x), "=r"(rv.y) : "r"(sr + idx * 8u) : "memory");
                grec[lane * S + i] = rv;
            }
            if (MODE == 0) reinterpret_cast<uint32_t *>(stp)[lane] = reinterpret_cast<const uint32_t *>(ssc)[lane];
        } else {
            uint4 cv;
            asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(cv.x), "=r"(cv.y), "=r"(cv.z), "=r"(cv.w)
                         : "r"(sA + (uint32_t)lane * kPanel + (uint32_t)c4 * 16u) : "memory");
            gcarry[lane] = cv;
        }
    }
    if (MODE == 0 && lane == 0) publish_snapshot(p, gridDim.x * (blockDim.x >> 5));
    tc_fence_before();
    __syncthreads();
    if (w == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
}

} /* namespace anm */
