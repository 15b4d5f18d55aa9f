/*
 * anm_kernels_tc.cuh -- dense tone sets (SPEC 3b, T = 64): the windows-by-basis contraction on the
 * 5th-generation tensor cores (tcgen05.mma kind::i8, accumulators in TMEM).  Generation 3 of this kernel.
 *
 * One persistent CTA per SM -- two of them, on the two SMs of a TPC, share every MMA as a CTA pair (cta_group::2, M = 256) -- works
 * through groups of four channels, 32 symbol periods (a "job") at a time.  The MMA rows of a CTA are interleaved,
 * row = 4 * symbol period + channel, so that "the previous symbol period of the same channel" is always four rows up: starting an
 * A descriptor 64 bytes earlier shifts every row by one symbol period (four carry rows in front of each panel hold the symbol
 * period before the job).  That puts the sliding window itself into the contraction:
 *
 *   W[i][plane] (128 x 128 per CTA, s32, TMEM) = sum_{j<=i} A[j][plane] . B[j]^T  +  sum_{j>i} A_prev[j][plane] . B[j]^T
 *        window ending with hop i of every symbol period, all 64 tones (128 columns) at once: 8 x tcgen05.mma
 *        (M 256 over the pair, N 128, K 32) per window and byte plane; each CTA of the pair supplies its own 128 rows of A and
 *        HALF of the basis B.
 *   W = 256 * W[hi plane] + W[lo plane] (exact integer, x = 256 hi + lo), E = fma(fI, fI, fQ fQ), argmax over tones.
 *
 * Twenty-one warps in four roles, decoupled by mbarriers only (no block-wide barrier inside the stream of jobs); two
 * groups of four channels are in flight per CTA and alternate job by job.  Warp ids go by priority (the warp schedulers
 * prefer the highest id among the eligible warps): state machines lowest, epilogue and issuer highest.
 *   8 state-machine warps (one per     merge the epilogue's candidates (lowest tone wins a tie), write the channel's hop-record
 *     channel of both groups)          ring and run sm_step (sync / slicing / framing, shared with k_demod); a state machine is
 *                                      one latency-bound warp, so each gets two job times per job
 *   4 loader warps (one per channel)   PCM (int16, HBM) --LDG.128, coalesced--> byte split (PRMT) --> A operand panels
 *                                      in shared memory (K-major, no swizzle, one panel per 16-sample K chunk, hop and
 *                                      byte plane), DOUBLE BUFFERED: job k+1 is loaded and split while job k is contracted
 *   8 epilogue warps (2 per TMEM lane  tcgen05.ld, 256 hi + lo, int -> fp32, energies (packed fp32 pairs), running argmax by
 *     quadrant, 32 tones each)         groups of four tones; per job they leave (emax, d) candidates per row and hop
 *   1 issuer warp (leader CTA only)    4 windows x 16 MMAs per job into two alternating accumulator sets (2 x 256 of the
 *                                      512 TMEM columns); multicast tcgen05.commit -> "accumulator full" / "A buffer free"
 * so the contraction of job k+1 runs under the epilogue of job k, which runs under the state machines of jobs k-1 and k-2.
 * DESIGN.md 3b has the measurements behind each of these choices and the ones that were tried and dropped.
 *
 * There is no reference kernel for this; behaviour is SPEC.md's.
 */
#pragma once
#include "anm_kernels.cuh"

/* 1: two CTAs on the two SMs of a TPC form a pair (cluster of 2, tcgen05 cta_group::2): one MMA covers both CTAs' rows (M = 256) and every
 * CTA keeps and reads only HALF of the basis (the B operand) -- a quarter less shared-memory operand traffic per SM, which is what bounds
 * this kernel.  0: every CTA on its own (cta_group::1). */
#ifndef ANM_TC_PAIR
#define ANM_TC_PAIR 1
#endif

/* nanoseconds a waiting warp sleeps between two polls of its barrier (the roles share the warp schedulers) */
#ifndef ANM_TC_ISSUER_NS
#define ANM_TC_ISSUER_NS 32
#endif
#ifndef ANM_TC_LOADER_NS
#define ANM_TC_LOADER_NS 128
#endif
#ifndef ANM_TC_SM_NS
#define ANM_TC_SM_NS 256
#endif

namespace anm {
namespace tc {

constexpr uint32_t kRows = 128;                    /* MMA M: 4 channels x 32 symbol periods, row = 4 * symbol period + channel */
constexpr uint32_t kCarryRows = 4;                 /* the symbol period before the job, one row per channel */
constexpr uint32_t kPanel = (kCarryRows + kRows) * 16u + 16u; /* one 16-byte K chunk of all rows; +16: spreads the panels over the banks */
constexpr uint32_t kNcol = 128;                    /* MMA N: (cos, sin) columns of all 64 tones */
constexpr uint32_t kPair = ANM_TC_PAIR ? 2u : 1u;   /* CTAs that share one MMA */
constexpr uint32_t kBPanel = kNcol / kPair * 16u;  /* one 16-byte K chunk of this CTA's share of the basis rows */
constexpr uint32_t kAccCols = 2u * kNcol;          /* one accumulator set: 2 byte planes x 128 columns */
constexpr uint32_t kTmemCols = 2u * kAccCols;      /* two sets = all 512 TMEM columns of the SM */
/* Warp roles, lowest priority first: the SM's warp schedulers prefer the highest warp id among the eligible warps, and the epilogue is the
 * critical path -- polling loops of the other roles must not win issue slots against it. */
constexpr int kSmWarp0 = 0;                        /* warps 0..7: group slot = warp >> 2, channel slot = warp & 3 */
#ifndef ANM_TC_LOADERS
#define ANM_TC_LOADERS 4
#endif
#ifndef ANM_TC_EPI_PARTS
#define ANM_TC_EPI_PARTS 2
#endif
constexpr int kLoaderWarp0 = 8;                    /* channel slot = (warp - 8) & 3; with 8 loader warps (warp - 8) >> 2 says which 16 of the job's 32 symbol periods */
constexpr int kLoaderWarps = ANM_TC_LOADERS;
constexpr int kLoaderRows = 128 / kLoaderWarps;    /* symbol periods of one channel per loader warp and job */
constexpr int kEpiParts = ANM_TC_EPI_PARTS;        /* epilogue warps per TMEM lane quadrant (= per warp scheduler): each a range of the tones */
constexpr int kEpiWarp0 = kLoaderWarp0 + kLoaderWarps; /* TMEM lane quadrant = warp & 3, tone range = (warp - kEpiWarp0) >> 2 */
constexpr int kEpiWarps = 4 * kEpiParts;
constexpr int kIssuerWarp = kEpiWarp0 + kEpiWarps;
constexpr int kWarps = kIssuerWarp + 1;            /* 25: 72 registers per thread (one scheduler holds 7 warps) */
static_assert(kEpiParts == 2 || kEpiParts == 3, "tone ranges");
static_assert(kLoaderWarps == 4 || kLoaderWarps == 8, "one or two loader warps per channel slot");                         /* 800 threads x 80 registers = 64,000 of the SM's 65,536 */

template <int N, int S>
__host__ __device__ constexpr uint32_t a_bytes() { return 2u * S * (uint32_t)(N / S / 16) * kPanel; } /* one buffer */
/* basis panels [hop phase q][K chunk][128 columns][16]: the basis of hop phase q is the first-quarter basis
 * rotated by (-j)^(bin q); keeping all S phases removes every rotation from the epilogue */
template <int T, int N, int S>
__host__ __device__ constexpr uint32_t b_bytes() { return (uint32_t)S * (uint32_t)(N / S / 16) * kBPanel; }
template <int T, int S>
__host__ __device__ constexpr uint32_t warp_bytes() { return 64u * S * 8u + 128u; } /* per channel: HopRec ring | scalars */
/* candidates [job parity][tone range][hop][channel slot: 36 entries, 32 used][symbol period] x {e, d}: an epilogue warp writes rows
 * 32 * quadrant + lane (8 symbol periods x 4 channels), a state-machine warp reads one channel's 32 symbol periods; the 36 keeps both
 * free of bank conflicts (the channels' 8-entry pieces of a store land 32 bytes apart modulo 128) */
constexpr uint32_t kCandPlane = 4u * 36u * 8u;
__host__ __device__ constexpr uint32_t cand_pos(uint32_t row) { return ((row & 3u) * 36u + (row >> 2)) * 8u; }
template <int S>
__host__ __device__ constexpr uint32_t cand_bytes() { return 2u * (uint32_t)kEpiParts * (uint32_t)S * kCandPlane; }
template <int T, int N, int S>
__host__ __device__ constexpr uint32_t smem_bytes() {
    return 2u * a_bytes<N, S>() + b_bytes<T, N, S>() + 8u * warp_bytes<T, S>() + cand_bytes<S>() + 128u; /* tail: mbarriers, TMEM address */
}

/* shared-memory matrix descriptor: K-major, no swizzle; LBO = stride between the two 16-byte K chunks
 * of an MMA, SBO = stride between groups of 8 rows (cute::UMMA::SmemDescriptor, version 1) */
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((addr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo_bytes >> 4) << 16) | ((uint64_t)(sbo_bytes >> 4) << 32) | (1ull << 46);
}
/* instruction descriptor of kind::i8: D = s32, B = s8, A = s8 (a_signed) or u8, both K-major, M = 128 rows per CTA of the group, N = ncol */
__device__ __forceinline__ uint32_t idesc_i8_n(bool a_signed, uint32_t ncol) {
    return (2u << 4) | ((a_signed ? 1u : 0u) << 7) | (1u << 10) | ((ncol >> 3) << 17) | (((kRows * kPair) >> 4) << 24);
}
__device__ __forceinline__ void mma_i8(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
#if ANM_TC_PAIR
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::i8 [%0], %1, %2, %3, {%5, %6, %7, %8, %9, %10, %11, %12}, p;\n\t"
        "}\n" ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(0u), "r"(0u), "r"(0u), "r"(0u), "r"(0u), "r"(0u), "r"(0u), "r"(0u)
        : "memory");
#else
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %6, %7, %8}, p;\n\t"
        "}\n" ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(0u), "r"(0u), "r"(0u), "r"(0u)
        : "memory");
#endif
}
/* completion of every MMA issued so far -> one arrival on the barrier at this shared-memory offset (in both CTAs of a pair) */
__device__ __forceinline__ void mma_commit(uint32_t bar) {
#if ANM_TC_PAIR
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar), "h"((uint16_t)3) : "memory");
#else
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
#endif
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
#if ANM_TC_PAIR
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
#else
    return 0u;
#endif
}
__device__ __forceinline__ void cluster_sync_all() {
#if ANM_TC_PAIR
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
#else
    __syncthreads();
#endif
}
/* Arrivals on the barrier at the same shared-memory offset in the pair's LEADER CTA (rank 0): the issuer there waits for both CTAs.
 * RELEASE = false (what the kernel uses): the barrier's default semantics, as CUTLASS' ClusterBarrier::arrive(cta_id).  What the
 * arrival protects is ordered by its own fences: the loaders' A panels by fence.proxy.async (each CTA's panels are read by the tensor
 * core of its own SM), the epilogue's TMEM reads by tcgen05.wait::ld + tcgen05.fence::before_thread_sync.  RELEASE = true makes the
 * arrival a release at cluster scope: a MEMBAR + ERRBAR per arrival, 11 % of the kernel's stall samples when every arrival had it. */
template <bool RELEASE>
__device__ __forceinline__ void mbar_arrive_leader(uint32_t bar) {
#if ANM_TC_PAIR
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(bar), "r"(0u));
    if (RELEASE) asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(r) : "memory");
    else asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(r) : "memory");
#else
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
#endif
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
__device__ __forceinline__ void tmem_ldn(uint32_t taddr, int32_t (&v)[8]) { tmem_ld8(taddr, v); }
__device__ __forceinline__ void tmem_ldn(uint32_t taddr, int32_t (&v)[16]) { tmem_ld16(taddr, v); }
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
/* the same for a warp that mostly waits (the MMA issuer): it sleeps between polls instead of spinning on the issue slots
 * of the scheduler it shares with two epilogue warps.  ACQUIRE = true pairs with the loaders' cluster-scope release in a CTA pair. */
template <int NS, bool ACQUIRE>
__device__ __forceinline__ void mbar_wait_issuer(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    for (;;) {
#if ANM_TC_PAIR
        if (ACQUIRE)
            asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                         : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
        else
#endif
            asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                         : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
        if (ok) break;
        __nanosleep(NS);
    }
}
template <int NS>
__device__ __forceinline__ void mbar_wait_relaxed(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    for (;;) {
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                     : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
        if (ok) break;
        __nanosleep(NS);
    }
}
__device__ __forceinline__ float fmax3(float a, float b, float c) {
    float r;
    asm("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
    return r;
}
/* running argmax step: if (E > em) { em = E; dm = t; } as one compare and two predicated moves (ptxas issues predicated
 * moves on the FMA pipe as IMAD.MOV, the ALU pipe -- conversions, compares -- is the busy one in the epilogue) */
__device__ __forceinline__ void argmax_step(float &em, uint32_t &dm, float E, uint32_t t) {
    asm("{\n\t.reg .pred p;\n\tsetp.gt.f32 p, %2, %0;\n\t@p mov.f32 %0, %2;\n\t@p mov.u32 %1, %3;\n\t}" : "+f"(em), "+r"(dm) : "f"(E), "r"(t));
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
__global__ void
#if ANM_TC_PAIR
__cluster_dims__(2, 1, 1)
#endif
__launch_bounds__(tc::kWarps * 32, 1) k_demod_tc(const __grid_constant__ KParams p) {
    using namespace tc;
    constexpr int H = N / S;
    constexpr int KC = H / 16;          /* 16-byte K chunks per hop */
    constexpr int KS = H / 32;          /* MMAs (K = 32) per hop */
    constexpr int CPS = N / 8;          /* 16-byte PCM chunks per symbol period */
    /* Tone ranges of the kEpiParts epilogue warps of a lane quadrant, [kTB[part], kTB[part + 1]).  Three warps: the schedulers' issue slots
     * were a third idle with two -- each warp alone is bound by its dependency chains -- and the first (lowest warp id = lowest priority
     * of the three) gets the short range. */
    constexpr int kTB1 = kEpiParts == 3 ? 16 : T / 2, kTB2 = kEpiParts == 3 ? 40 : T;
#ifndef ANM_TC_TN
#define ANM_TC_TN 8
#endif
    constexpr int TN = ANM_TC_TN;       /* tones per tcgen05.ld batch (2 * TN columns per byte plane); the running argmax goes by groups of four */
    constexpr uint32_t RM = 64u * S - 1u;
    constexpr uint32_t CUR = kCarryRows * 16u; /* byte offset of the current rows inside a panel */
    static_assert(T == 64 && 2 * T == (int)kNcol, "the dense kernel contracts all 64 tones (128 columns) per MMA");
    static_assert(S == 4 && (H % 32) == 0 && CPS == 32, "unsupported dense geometry");
    static_assert(2 * S * KC == 32, "one (plane, hop, K chunk) panel per lane");
    static_assert(2u * S * KC * 16u <= (uint32_t)(S - 1) * T * 8u, "carry rows must fit the state's carry area");
    static_assert((kTB1 % (2 * TN)) == 0 && (kTB2 % (2 * TN)) == 0 && (T % (2 * TN)) == 0, "two tone batches per trip of the epilogue loop");

    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31;
    const int w = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0); /* through a shuffle: the compiler then treats it (and the TMEM addresses built on it) as warp-uniform */
    const uint32_t sA = (uint32_t)__cvta_generic_to_shared(smem_raw);       /* two A buffers */
    const uint32_t sB = sA + 2u * a_bytes<N, S>();
    unsigned char *chsm = smem_raw + 2u * a_bytes<N, S>() + b_bytes<T, N, S>(); /* per channel slot: HopRec ring [64*S] | ChanScalars */
    const uint32_t sr0 = (uint32_t)__cvta_generic_to_shared(chsm);
    const uint32_t sCand = sr0 + 8u * warp_bytes<T, S>();
    unsigned char *tail = chsm + 8u * warp_bytes<T, S>() + cand_bytes<S>();
    const uint32_t bars = (uint32_t)__cvta_generic_to_shared(tail);
    /* mbarriers (8 bytes each): [kind][buffer] */
    const uint32_t bar_a_full = bars, bar_a_empty = bars + 16u, bar_acc_full = bars + 32u, bar_acc_empty = bars + 48u,
                   bar_cand_full = bars + 64u, bar_cand_empty = bars + 80u;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(tail + 96);

    const uint32_t n_steps = (p.n_syms + 31u) / 32u;
    const uint32_t n_groups = (p.n_ch + 3u) / 4u;
    const uint32_t n_pairs = (n_groups + 1u) / 2u;
    /* A pair of CTAs walks the group pairs in lockstep: unit u of the grid's n_units holds kPair group pairs, CTA rank r takes pair kPair * u + r.  A job
     * exists for the pair when rank 0's group exists; a CTA whose own group does not exist still runs the job's barriers (on no channels). */
    const uint32_t cta_rank = cluster_ctarank();
    const bool leader = cta_rank == 0u;
    const uint32_t n_units = (n_pairs + kPair - 1u) / kPair, unit0 = blockIdx.x / kPair, unit_step = gridDim.x / kPair;
    /* Two groups are in flight per CTA, group slot g2 of a pair always in A buffer / candidate buffer g2: their jobs alternate
     * (pair, step, slot 0), (pair, step, slot 1), ... so that every channel's state machine -- one warp, latency bound -- has two
     * job times for one job.  A missing second group (odd group count) is skipped by every role alike. */

    /* ---- one-time setup: basis panels, mbarriers, TMEM ---- */
    {
        /* global panels [hop phase][K chunk][128 columns][16]; this CTA keeps columns [rank * 128 / kPair, (rank + 1) * 128 / kPair) of every panel */
        const uint4 *gsrc = reinterpret_cast<const uint4 *>(p.tc_basis);
        uint4 *dst = reinterpret_cast<uint4 *>(smem_raw + 2u * a_bytes<N, S>());
        constexpr uint32_t CPP = kNcol / kPair; /* 16-byte column entries per local panel */
        for (uint32_t i = threadIdx.x; i < b_bytes<T, N, S>() / 16u; i += blockDim.x)
            dst[i] = __ldg(&gsrc[(i / CPP) * kNcol + cta_rank * CPP + (i % CPP)]);
    }
    if (threadIdx.x == 0) {
        for (uint32_t b = 0; b < 2; ++b) {
            mbar_init(bar_a_full + 8u * b, (uint32_t)kLoaderWarps * kPair); /* one arrival per loader warp of the pair (the leader's barrier is the one in use) */
            mbar_init(bar_a_empty + 8u * b, 1u);                    /* tcgen05.commit */
            mbar_init(bar_acc_full + 8u * b, 1u);                   /* tcgen05.commit */
            mbar_init(bar_acc_empty + 8u * b, (uint32_t)kEpiWarps * kPair); /* every epilogue warp of the pair, on the leader's barrier */
            mbar_init(bar_cand_full + 8u * b, (uint32_t)kEpiWarps);
            mbar_init(bar_cand_empty + 8u * b, 4u);                 /* one arrival per state-machine warp */
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (w == 0) { /* in a pair, warp 0 of BOTH CTAs issues the allocation */
#if ANM_TC_PAIR
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(tmem_slot)), "r"(kTmemCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
#else
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(tmem_slot)), "r"(kTmemCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
#endif
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); /* basis panels -> visible to the MMA's async proxy */
    tc_fence_before();
    __syncthreads();
    cluster_sync_all(); /* barriers initialised and basis in place in both CTAs before anything crosses over */
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    if (w >= kLoaderWarp0 && w < kLoaderWarp0 + kLoaderWarps) {
        /* =================== loader: PCM of one job of this warp's channel -> byte planes in the A panels =================== */
        const int c4 = (w - kLoaderWarp0) & 3, rh = (w - kLoaderWarp0) >> 2; /* two warps per channel: symbol periods [16 * rh, 16 * rh + 16) of the job */
        /* Every load instruction of the warp covers ONE symbol period (32 lanes x 16 bytes = 512 bytes = N samples): the lane is the
         * 16-byte chunk, i.e. (hop q, K chunk kc, which half of the 16-sample K chunk) are lane constants. */
        const uint32_t q = (uint32_t)lane / (uint32_t)(H / 8), hc = (uint32_t)lane % (uint32_t)(H / 8);
        const uint32_t lane_off = (q * KC + (hc >> 1)) * kPanel + (uint32_t)c4 * 16u + (hc & 1u) * 8u; /* carry row of this channel */
        const uint32_t plane_lo = (uint32_t)(S * KC) * kPanel;
        const uint32_t st_hi = (q * KC + (hc >> 1)) * 16u + (hc & 1u) * 8u, st_lo = st_hi + (uint32_t)(S * KC) * 16u; /* state carry: [plane][hop][K chunk] x 16 bytes */
        uint32_t use[2] = {0u, 0u};
        for (uint32_t unit = unit0; unit < n_units; unit += unit_step)
            for (uint32_t step = 0; step < n_steps; ++step)
#pragma unroll
            for (uint32_t b = 0; b < 2; ++b) {
                if (2u * (kPair * unit) + b >= n_groups) continue; /* no such job for the pair */
                const uint32_t grp = 2u * (kPair * unit + cta_rank) + b;
                const uint32_t ch = grp * 4u + (uint32_t)c4;
                const bool have = grp < n_groups && ch < p.n_ch;
                const char *src = reinterpret_cast<const char *>(p.pcm + (size_t)(have ? ch : 0u) * p.ch_stride) + (size_t)lane * 16u;
                unsigned char *gst = p.state + (size_t)(have ? ch : 0u) * p.state_stride + state_carry_offset<T, S>();
                const uint32_t k = use[b]++;
                const int nv = (int)min(32u, p.n_syms - step * 32u);
                mbar_wait_relaxed<ANM_TC_LOADER_NS>(bar_a_empty + 8u * b, (k & 1u) ^ 1u); /* the contraction that read this buffer two jobs ago is complete */
                if (have) {
                    const uint32_t base = sA + b * a_bytes<N, S>() + lane_off;
                    const char *g = src + (size_t)step * (32u * N * 2u);
                    /* carry row: the symbol period before the job (step 0: the planes saved by the previous chunk) */
                    if (rh != 0) {
                    } else if (step == 0) {
                        const uint2 chi = *reinterpret_cast<const uint2 *>(gst + st_hi);
                        const uint2 clo = *reinterpret_cast<const uint2 *>(gst + st_lo);
                        asm volatile("st.shared.v2.u32 [%0], {%1,%2};" ::"r"(base), "r"(chi.x), "r"(chi.y) : "memory");
                        asm volatile("st.shared.v2.u32 [%0], {%1,%2};" ::"r"(base + plane_lo), "r"(clo.x), "r"(clo.y) : "memory");
                    } else {
                        const uint4 v = __ldg(reinterpret_cast<const uint4 *>(g - 512));
                        asm volatile("st.shared.v2.u32 [%0], {%1,%2};" ::"r"(base), "r"(prmt(v.x, v.y, 0x7531u)), "r"(prmt(v.z, v.w, 0x7531u)) : "memory");
                        asm volatile("st.shared.v2.u32 [%0], {%1,%2};" ::"r"(base + plane_lo), "r"(prmt(v.x, v.y, 0x6420u)), "r"(prmt(v.z, v.w, 0x6420u)) : "memory");
                    }
                    constexpr int BATCH = 8; /* loads in flight before the first split (21 warps leave 80 registers per thread) */
#pragma unroll 1
                    for (int r0 = kLoaderRows * rh; r0 < kLoaderRows * rh + kLoaderRows; r0 += BATCH) {
                        uint4 v[BATCH];
                        if (nv == 32) {
#pragma unroll
                            for (int j = 0; j < BATCH; ++j) v[j] = __ldg(reinterpret_cast<const uint4 *>(g + (size_t)(r0 + j) * 512u));
                        } else {
                            /* ragged last job of a chunk: rows past the end repeat the last symbol period (their results are never used) */
#pragma unroll
                            for (int j = 0; j < BATCH; ++j) v[j] = __ldg(reinterpret_cast<const uint4 *>(g + (size_t)min(r0 + j, nv - 1) * 512u));
                        }
                        if (r0 == kLoaderRows * rh && step + 1 < n_steps) {
                            /* next job of this channel: 32 * N * 2 bytes = N / 2 lines of 128 bytes, N / 64 per lane, pulled into L2 */
                            const char *nx = reinterpret_cast<const char *>(p.pcm + (size_t)ch * p.ch_stride) + (size_t)(step + 1) * (32u * N * 2u) + (size_t)lane * 128u +
                                             (size_t)rh * (N * kLoaderRows / 2048) * 4096u; /* each of the channel's warps its share */
#pragma unroll
                            for (int j = 0; j < N * kLoaderRows / 2048; ++j) asm volatile("prefetch.global.L2 [%0];" ::"l"(nx + (size_t)j * 4096u));
                        }
#pragma unroll
                        for (int j = 0; j < BATCH; ++j) {
                            const uint32_t a = base + CUR + (uint32_t)(r0 + j) * 64u; /* row 4 * (r0 + j) + c4 */
                            const uint32_t hi0 = prmt(v[j].x, v[j].y, 0x7531u), hi1 = prmt(v[j].z, v[j].w, 0x7531u);
                            const uint32_t lo0 = prmt(v[j].x, v[j].y, 0x6420u), lo1 = prmt(v[j].z, v[j].w, 0x6420u);
                            asm volatile("st.shared.v2.u32 [%0], {%1,%2};" ::"r"(a), "r"(hi0), "r"(hi1) : "memory");
                            asm volatile("st.shared.v2.u32 [%0], {%1,%2};" ::"r"(a + plane_lo), "r"(lo0), "r"(lo1) : "memory");
                        }
                    }
                    if (rh == 0 && step + 1 == n_steps) { /* the warp that read the old carry at step 0 */
                        /* the chunk's last symbol period is the next chunk's carry row (read back from L2: once per chunk) */
                        const uint4 v = __ldg(reinterpret_cast<const uint4 *>(g + (size_t)(nv - 1) * 512u));
                        *reinterpret_cast<uint2 *>(gst + st_hi) = make_uint2(prmt(v.x, v.y, 0x7531u), prmt(v.z, v.w, 0x7531u));
                        *reinterpret_cast<uint2 *>(gst + st_lo) = make_uint2(prmt(v.x, v.y, 0x6420u), prmt(v.z, v.w, 0x6420u));
                    }
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); /* generic-proxy stores -> the MMA's async proxy */
                }
                __syncwarp();
                if (lane == 0) mbar_arrive_leader<false>(bar_a_full + 8u * b);
            }
    } else if (w == kIssuerWarp) {
        if (leader) {
        /* =================== issuer: 4 windows x (2 planes x 4 hops x 2 K steps) MMAs per job =================== */
        uint32_t use[2] = {0u, 0u}, rnd = 0;
        const uint32_t id_hi = idesc_i8_n(true, kNcol), id_lo = idesc_i8_n(false, kNcol);
        const uint64_t b0 = smem_desc(sB, kBPanel, 128u);
        for (uint32_t unit = unit0; unit < n_units; unit += unit_step)
            for (uint32_t step = 0; step < n_steps; ++step)
#pragma unroll
            for (uint32_t b = 0; b < 2; ++b) {
                if (2u * (kPair * unit) + b >= n_groups) continue;
                const uint32_t k = use[b]++;
                mbar_wait_issuer<ANM_TC_ISSUER_NS, false>(bar_a_full + 8u * b, k & 1u);
                tc_fence_after();
                const uint64_t a0 = smem_desc(sA + b * a_bytes<N, S>(), kPanel, 128u);
#pragma unroll 1 /* rolled: the four roles share the instruction cache, keep every role's loop body small */
                for (int i = 0; i < S; ++i, ++rnd) {
                    const uint32_t set = rnd & 1u, u = rnd >> 1;
                    mbar_wait_issuer<ANM_TC_ISSUER_NS, false>(bar_acc_empty + 8u * set, (u & 1u) ^ 1u);
                    tc_fence_after();
                    const uint32_t d0 = tmem_base + set * kAccCols;
                    if (elect_one()) {
#pragma unroll
                        for (int pl = 0; pl < 2; ++pl)
#pragma unroll
                            for (int j = 0; j < S; ++j)
#pragma unroll
                                for (int ks = 0; ks < KS; ++ks) {
                                    /* hops up to the window's last one come from this symbol period, later ones from the
                                     * previous symbol period of the same channel: four rows (64 bytes) up */
                                    const uint32_t aoff = (uint32_t)((pl * S + j) * KC + 2 * ks) * kPanel + ((j <= i) ? CUR : 0u);
                                    const uint32_t boff = (uint32_t)(j * KC + 2 * ks) * kBPanel;
#ifndef ANM_TC_DEBUG_SKIP_MMA /* timing experiment only: results are garbage */
                                    mma_i8(d0 + (uint32_t)pl * kNcol, a0 + (uint64_t)(aoff >> 4), b0 + (uint64_t)(boff >> 4), pl == 0 ? id_hi : id_lo,
                                           (j > 0 || ks > 0) ? 1u : 0u);
#endif
                                }
                        mma_commit(bar_acc_full + 8u * set);
                        if (i == S - 1) mma_commit(bar_a_empty + 8u * b); /* every contraction that reads this A buffer is complete */
                    }
                    __syncwarp();
                }
            }
        }
    } else if (w >= kEpiWarp0) {
        /* =================== epilogue: this warp's 32 tones of its TMEM lane quadrant =================== */
        const int c4 = w & 3, hf = (w - kEpiWarp0) >> 2;
        const int esp = 8 * c4 + (lane >> 2), ec4 = lane & 3; /* TMEM lane 32 * c4 + lane = row 4 * esp + ec4 */
        const int tlo = hf == 0 ? 0 : hf == 1 ? kTB1 : kTB2, thi = hf == 0 ? kTB1 : hf == 1 ? kTB2 : T; /* this warp's tones */
        const uint32_t tmem_lane = tmem_base + ((uint32_t)(32 * c4) << 16) + (uint32_t)(2 * tlo);
        const int n_tb = (thi - tlo) / TN; /* tone batches of this warp per window */
        const uint32_t row = (uint32_t)(32 * c4 + lane);
        uint32_t use[2] = {0u, 0u}, rnd = 0;
        for (uint32_t unit = unit0; unit < n_units; unit += unit_step)
            for (uint32_t step = 0; step < n_steps; ++step)
#pragma unroll
            for (uint32_t b = 0; b < 2; ++b) {
                if (2u * (kPair * unit) + b >= n_groups) continue;
                const uint32_t grp = 2u * (kPair * unit + cta_rank) + b;
                const uint32_t ech = grp < n_groups ? grp * 4u + (uint32_t)ec4 : 0xFFFFFFFFu;
                const uint32_t k = use[b]++;
                /* candidates of this tone half: [job parity][half][hop][row] x {e, d} */
                mbar_wait(bar_cand_empty + 8u * b, (k & 1u) ^ 1u); /* the state machines have read what job - 2 left here */
                const uint32_t ca = sCand + ((b * (uint32_t)kEpiParts + (uint32_t)hf) * (uint32_t)S) * kCandPlane + cand_pos(row);
#pragma unroll 1 /* rolled: see the issuer */
                for (int i = 0; i < S; ++i, ++rnd) {
                    const uint32_t set = rnd & 1u, u = rnd >> 1;
                    mbar_wait(bar_acc_full + 8u * set, u & 1u);
                    tc_fence_after();
                    const uint32_t t0 = tmem_lane + set * kAccCols;
                    float em = -1.0f, w0 = 0.0f, w1 = 0.0f, w2 = 0.0f; /* running maximum; the first three energies of the group of four it came from */
                    uint32_t dq = 0u;                                   /* first tone of that group */
                    int32_t vh[2][2 * TN], vl[2][2 * TN];
                    tmem_ldn(t0, vh[0]);
                    tmem_ldn(t0 + kNcol, vl[0]);
                    /* two batches per trip (ping-pong registers): a short loop body -- the roles share the instruction cache */
#pragma unroll 1
#ifdef ANM_TC_DEBUG_SKIP_EPI /* timing experiment only: results are garbage */
                    for (int tb2 = 0; tb2 < 0; tb2 += 2) {
#else
                    for (int tb2 = 0; tb2 < n_tb; tb2 += 2) {
#endif
#pragma unroll
                        for (int h2 = 0; h2 < 2; ++h2) {
                            const int tb = tb2 + h2;
                            tmem_ld_wait();
                            { /* the next batch travels while this one is evaluated; behind the last batch the same one again, rather than a
                               * branch: the trip stays one basic block and ptxas overlaps the two batches' dependency chains */
                                const uint32_t nb = (uint32_t)(2 * TN) * (uint32_t)min(tb + 1, n_tb - 1);
                                tmem_ldn(t0 + nb, vh[h2 ^ 1]);
                                tmem_ldn(t0 + kNcol + nb, vl[h2 ^ 1]);
                            }
                            const int32_t(&xh)[2 * TN] = vh[h2];
                            const int32_t(&xl)[2 * TN] = vl[h2];
                            const int tone0 = tlo + tb * TN;
#pragma unroll
                            for (int tq = 0; tq < TN; tq += 4) {
                                float E[4];
#pragma unroll
                                for (int pr = 0; pr < 2; ++pr) {
                                    /* columns 4 * (tone / 2) + {I even, I odd, Q even, Q odd}: two tones per packed fp32 operation; each
                                     * component sees exactly fma(fI, fI, fQ * fQ) of SPEC 3b on fI = RN(256 * hi + lo) */
                                    const int c = 2 * tq + 4 * pr;
                                    float2 fi, fq;
                                    fi = make_float2((float)(xh[c] * 256 + xl[c]), (float)(xh[c + 1] * 256 + xl[c + 1]));
                                    fq = make_float2((float)(xh[c + 2] * 256 + xl[c + 2]), (float)(xh[c + 3] * 256 + xl[c + 3]));
                                    const float2 e2 = ffma2vv(fi, fi, fmul2(fq, fq));
                                    E[2 * pr] = e2.x;
                                    E[2 * pr + 1] = e2.y;
                                }
                                if (MODE == 1) {
                                    if (p.trE && ech < p.n_ch && esp < (int)min(32u, p.n_syms - step * 32u)) {
                                        const size_t hop = ((size_t)step * 32 + esp) * S + i;
                                        float *o = p.trE + ((size_t)ech * p.tr_hops + hop) * T + tone0 + tq;
#pragma unroll
                                        for (int j = 0; j < 4; ++j) o[j] = E[j];
                                    }
                                }
                                /* the running argmax by groups of four tones: strictly greater replaces (the first group always wins against
                                 * em = -1: energies are >= 0), so the lowest tone among equal maxima is kept, as tone by tone */
                                const float m = fmaxf(fmax3(E[0], E[1], E[2]), E[3]);
                                const bool gt = m > em;
                                em = gt ? m : em;
                                w0 = gt ? E[0] : w0;
                                w1 = gt ? E[1] : w1;
                                w2 = gt ? E[2] : w2;
                                dq = gt ? (uint32_t)(tone0 + tq) : dq;
                            }
                        }
                    }
                    const uint32_t dm = dq + (w0 == em ? 0u : w1 == em ? 1u : w2 == em ? 2u : 3u);
                    tmem_ld_wait(); /* the repeated load of the last batch */
                    /* this warp's TMEM reads of the set are complete (tcgen05.wait::ld): hand it back */
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive_leader<false>(bar_acc_empty + 8u * set);
                    asm volatile("st.shared.v2.u32 [%0], {%1,%2};" ::"r"(ca + (uint32_t)i * kCandPlane), "r"(__float_as_uint(em)), "r"(dm) : "memory");
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_cand_full + 8u * b);
            }
    } else {
        /* =================== state machine: channel slot c4, lane = symbol period =================== */
        const int c4 = (w - kSmWarp0) & 3;
        const uint32_t b = (uint32_t)(w - kSmWarp0) >> 2; /* group slot of the pair = A / candidate buffer this warp serves */
        const uint32_t sr = sr0 + (uint32_t)(w - kSmWarp0) * warp_bytes<T, S>();
        ChanScalars *ssc = reinterpret_cast<ChanScalars *>(chsm + (size_t)(w - kSmWarp0) * warp_bytes<T, S>() + 64u * S * 8u);
        const uint32_t crc_k = (MODE == 0) ? (uint32_t)p.crc_pow[lane] : 0u;
        const uint32_t row = (uint32_t)(4 * lane + c4);
        uint32_t k = 0;
        for (uint32_t unit = unit0; unit < n_units; unit += unit_step) {
            if (2u * (kPair * unit) + b >= n_groups) continue;
            const uint32_t grp = 2u * (kPair * unit + cta_rank) + b;
            const uint32_t ch = grp * 4u + (uint32_t)c4;
            const bool have = grp < n_groups && ch < p.n_ch;
            unsigned char *stp = p.state + (size_t)(have ? ch : 0u) * p.state_stride;
            uint2 *grec = reinterpret_cast<uint2 *>(stp + sizeof(ChanScalars));
            if (have) {
                /* carried state: the last 32 symbol slots go to ring slots 32..63 */
#pragma unroll
                for (int i = 0; i < S; ++i) {
                    const uint2 rv = grec[lane * S + i];
                    asm volatile("st.shared.v2.u32 [%0], {%1,%2};" ::"r"(sr + ring_off<S, true>((uint32_t)((32 + lane) * S + i))), "r"(rv.x), "r"(rv.y) : "memory");
                }
                if (MODE == 0) reinterpret_cast<uint32_t *>(ssc)[lane] = reinterpret_cast<const uint32_t *>(stp)[lane];
            }
            __syncwarp();
            for (uint32_t step = 0; step < n_steps; ++step, ++k) {
                const int nvalid = (int)min(32u, p.n_syms - step * 32u);
                const bool active = lane < nvalid;
                const uint32_t hic = step * 32u * S;
                mbar_wait_relaxed<ANM_TC_SM_NS>(bar_cand_full + 8u * b, k & 1u);
                uint32_t dc[S];
                float ec[S];
                {
                    const uint32_t c0 = sCand + ((b * (uint32_t)kEpiParts) * (uint32_t)S) * kCandPlane + cand_pos(row);
#pragma unroll
                    for (int i = 0; i < S; ++i) {
                        uint32_t e0, d0;
                        asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(e0), "=r"(d0) : "r"(c0 + (uint32_t)i * kCandPlane) : "memory");
#pragma unroll
                        for (int part = 1; part < kEpiParts; ++part) {
                            uint32_t f0, g0;
                            asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(f0), "=r"(g0) : "r"(c0 + (uint32_t)(part * S + i) * kCandPlane) : "memory");
                            /* lowest tone index wins a tie (SPEC 3): later ranges hold the higher tones */
                            const bool t0 = __uint_as_float(f0) > __uint_as_float(e0);
                            e0 = t0 ? f0 : e0;
                            d0 = t0 ? g0 : d0;
                        }
                        ec[i] = __uint_as_float(e0);
                        dc[i] = d0;
                    }
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_cand_empty + 8u * b);
                if (!active) {
#pragma unroll
                    for (int i = 0; i < S; ++i) { dc[i] = 0xFFu; ec[i] = 0.0f; }
                }
                if (have) {
                    if (active) {
                        const uint32_t i0 = (hic + (uint32_t)(lane * S)) & RM;
#pragma unroll
                        for (int i = 0; i < S; ++i)
                            asm volatile("st.shared.v2.u32 [%0], {%1,%2};" ::"r"(sr + ring_off<S, true>(i0 + (uint32_t)i)), "r"(__float_as_uint(ec[i])), "r"(dc[i]) : "memory");
                        if (MODE == 1) {
                            if (p.trD) {
#pragma unroll
                                for (int i = 0; i < S; ++i) p.trD[(size_t)ch * p.tr_hops + ((size_t)step * 32 + lane) * S + i] = (uint8_t)dc[i];
                            }
                            if (p.trEmax) {
#pragma unroll
                                for (int i = 0; i < S; ++i) p.trEmax[(size_t)ch * p.tr_hops + ((size_t)step * 32 + lane) * S + i] = ec[i];
                            }
                        }
                    }
                    __syncwarp();
                    if (MODE == 0) sm_step<T, N, S, true>(p, ch, lane, sr, hic, nvalid, active, dc, (uint32_t)__cvta_generic_to_shared(ssc), crc_k, p.hop_base);
                }
            }
            /* ---- save carried state: the last 32 symbol slots of the chunk ---- */
            __syncwarp();
            if (have && n_steps) {
#pragma unroll
                for (int i = 0; i < S; ++i) {
                    const uint32_t idx = ((p.n_syms - 32u + (uint32_t)lane) * S + (uint32_t)i) & RM;
                    uint2 rv;
                    asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(rv.x), "=r"(rv.y) : "r"(sr + ring_off<S, true>(idx)) : "memory");
                    grec[lane * S + i] = rv;
                }
                if (MODE == 0) reinterpret_cast<uint32_t *>(stp)[lane] = reinterpret_cast<const uint32_t *>(ssc)[lane];
            }
            __syncwarp();
        }
    }

    if (MODE == 0 && lane == 0) publish_snapshot(p, gridDim.x * (blockDim.x >> 5));
    tc_fence_before();
    __syncthreads();
    cluster_sync_all(); /* the peer's shared memory and barriers stay alive until both CTAs are done */
    if (w == 0) {
#if ANM_TC_PAIR
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
#else
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
#endif
    }
}

} /* namespace anm */
