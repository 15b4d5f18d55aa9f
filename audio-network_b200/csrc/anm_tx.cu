/*
 * anm_tx.cu -- CUDA renderer of the transmitter stand-in (SPEC.md section 6).  Same
 * integer arithmetic as anm_tx.c, one thread per 8 output samples (one 16-byte store),
 * so bulk synthetic workloads (thousands of channels x seconds) are produced in HBM
 * without crossing PCIe.  Not on the receive hot path.
 */
#include <cuda_runtime.h>

#include "anm_internal.h"

namespace {

__constant__ int16_t c_sine[1024];
bool g_sine_loaded[64] = {false};

struct TxK {
    const uint8_t *programs;
    unsigned long long prog_stride;
    const uint32_t *prog_len;
    const anm_tx_params_t *params;
    uint32_t n_ch;
    unsigned long long first_sample;
    int16_t *pcm;
    unsigned long long ch_stride;
    unsigned long long n;
    uint32_t lg, n_tones;
    uint32_t bins[ANM_MAX_TONES];
};

__device__ __forceinline__ unsigned long long mix64(unsigned long long z) {
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

__global__ void __launch_bounds__(256) k_tx_render(const __grid_constant__ TxK k) {
    const unsigned long long groups = (k.n + 7) / 8;
    const uint32_t ch = blockIdx.x;
    const anm_tx_params_t prm = k.params[ch];
    const uint32_t nscale = prm.reserved; /* host-computed Q20 noise scale */
    long long step = (1ll << 32);
    {
        /* round_half_away(ppm_x1000 * 4294.967296 / 1000) in exact integer arithmetic:
         * 4294.967296/1000 = 2^32 / 1e9 */
        const long long num = (long long)prm.ppm_x1000 * 4294967296ll; /* |.| < 2^63 for |ppm_x1000| < 2^30 */
        const long long q = (num >= 0) ? (num + 500000000ll) / 1000000000ll : -((-num + 500000000ll) / 1000000000ll);
        step += q;
    }
    const uint8_t *prog = k.programs + (size_t)ch * k.prog_stride;
    const uint32_t plen = k.prog_len[ch];
    for (unsigned long long gidx = blockIdx.y * (unsigned long long)blockDim.x + threadIdx.x; gidx < groups;
         gidx += (unsigned long long)gridDim.y * blockDim.x) {
        int16_t v8[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const unsigned long long i = gidx * 8 + j;
            const unsigned long long nn = k.first_sample + i;
            /* 128-bit like the CPU renderer (anm_tx.c): start_offset * 2^32 + n * step leaves 64 bits after 2^31 samples */
            const __int128 tpos = (__int128)prm.start_offset * ((__int128)1 << 32) + (__int128)nn * (__int128)step;
            int sig = 0;
            if (tpos >= 0) {
                const unsigned long long t = (unsigned long long)(tpos >> 32);
                const unsigned long long sym = t >> k.lg;
                const uint32_t e = prog[sym % plen];
                if (e != ANM_SILENCE && e < k.n_tones) {
                    const unsigned long long pos = (unsigned long long)(tpos - ((__int128)(sym << k.lg) << 32)); /* Q32, < N * 2^32 */
                    const uint32_t phase = (uint32_t)(((unsigned long long)k.bins[e] * pos) >> k.lg);
                    sig = ((int)prm.amplitude_q15 * (int)c_sine[phase >> 22]) >> 15;
                }
            }
            int noise = 0;
            if (nscale) {
                const unsigned long long a = mix64(prm.seed + (2 * nn) * 0x9E3779B97F4A7C15ull);
                const unsigned long long b = mix64(prm.seed + (2 * nn + 1) * 0x9E3779B97F4A7C15ull);
                const long long s = (long long)((a & 0xFFFF) + ((a >> 16) & 0xFFFF) + ((a >> 32) & 0xFFFF) + (a >> 48) +
                                                (b & 0xFFFF) + ((b >> 16) & 0xFFFF) + ((b >> 32) & 0xFFFF) + (b >> 48)) - 262140;
                noise = (int)((s * (long long)nscale) >> 20);
            }
            const int v = sig + noise;
            v8[j] = (int16_t)(v > 32767 ? 32767 : (v < -32768 ? -32768 : v));
        }
        int16_t *dst = k.pcm + (size_t)ch * k.ch_stride + gidx * 8;
        if (gidx * 8 + 8 <= k.n && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0)) {
            *reinterpret_cast<uint4 *>(dst) = *reinterpret_cast<const uint4 *>(v8);
        } else {
            for (int j = 0; j < 8 && gidx * 8 + j < k.n; ++j) dst[j] = v8[j];
        }
    }
}

} /* namespace */

extern "C" int anm_tx_render_device(const anm_config_t *cfg, const uint8_t *d_programs, size_t prog_stride,
                                    const uint32_t *d_prog_len, const anm_tx_params_t *d_params, uint32_t n_ch,
                                    uint64_t first_sample, int16_t *d_pcm, size_t ch_stride, size_t n, void *stream) {
    if (anm_config_validate(cfg) != ANM_OK || !d_programs || !d_prog_len || !d_params || !d_pcm) return ANM_ERR_ARG;
    if (n_ch == 0 || n == 0) return ANM_OK;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) { anm_set_error("no CUDA device"); return ANM_ERR_CUDA; }
    if (dev < 64 && !g_sine_loaded[dev]) {
        if (cudaMemcpyToSymbol(c_sine, anm_tx_sine_table(), 1024 * sizeof(int16_t)) != cudaSuccess) {
            anm_set_error("sine table upload failed: %s", cudaGetErrorString(cudaGetLastError()));
            return ANM_ERR_CUDA;
        }
        g_sine_loaded[dev] = true;
    }
    TxK k;
    k.programs = d_programs;
    k.prog_stride = prog_stride;
    k.prog_len = d_prog_len;
    k.params = d_params;
    k.n_ch = n_ch;
    k.first_sample = first_sample;
    k.pcm = d_pcm;
    k.ch_stride = ch_stride;
    k.n = n;
    k.lg = 0;
    while ((1u << k.lg) < cfg->sym_len) ++k.lg;
    k.n_tones = cfg->n_tones;
    for (uint32_t i = 0; i < ANM_MAX_TONES; ++i) k.bins[i] = cfg->tone_bin[i];
    const unsigned long long groups = (n + 7) / 8;
    unsigned gx = (unsigned)((groups + 255) / 256);
    if (gx > 4096) gx = 4096;
    dim3 grid(n_ch, gx);
    k_tx_render<<<grid, 256, 0, (cudaStream_t)stream>>>(k);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { anm_set_error("tx render launch: %s", cudaGetErrorString(e)); return ANM_ERR_CUDA; }
    return ANM_OK;
}
