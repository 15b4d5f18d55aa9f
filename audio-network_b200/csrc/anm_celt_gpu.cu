/*
 * anm_celt_gpu.cu -- batched CELT entropy decode on the GPU (include/anmodem_opus.h, anm_celt_entropy_*; SURVEY.md 8(f) row f1,
 * stage 1).  Two passes.  k_celt_entropy, one thread per FRAME: the range decoder is sequential inside a frame, but no symbol of a
 * frame depends on any other frame, so every frame of every stream decodes at once (integer / byte work on a few hundred bytes of
 * packet, latency bound per thread: the batch of frames is what fills the machine).  k_celt_energies, one thread per STREAM: the band
 * energies predict from the previous frame of the same stream (celt/quant_bands.c:427-490) -- a recurrence of about a hundred integer
 * operations per frame over the coarse symbols and offsets pass 1 left in a scratch array (anm_celt_entropy.h holds the decode itself,
 * shared with the host-side test harness).
 *
 * Reference path replaced: playback.cpp:115-122 opus_decode() -> opus_decode_frame (opus_decoder.c:214-626, CELT-only branch)
 * -> celt_decode_with_ec (celt/celt_decoder.c:815-1095), up to and including unquant_energy_finalise.
 */
#include <cuda_runtime.h>

#include <new>

#include "anm_celt_entropy.h"
#include "anm_internal.h"

struct anm_celt_ctx {
    int device;
    anm_celt_tables_t *d_tables;
    int16_t *d_scratch; /* per frame: 42 coarse symbols + 42 energy offsets */
    size_t scratch_frames;
};

namespace {

/* pass 1, one thread per FRAME: everything the frame's bits say.  No symbol depends on the stream's history, so all frames of all streams
 * decode at once; what the history needs (coarse symbols, energy offsets) goes to the scratch array. */
__global__ void __launch_bounds__(128) k_celt_entropy(const anm_celt_tables_t *__restrict__ t, const anm_celt_job_t *__restrict__ jobs, uint32_t n_jobs,
                                                      const uint8_t *__restrict__ bytes, uint32_t mask, int16_t *__restrict__ scratch, anm_celt_frame_t *out) {
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n_jobs) return;
    const anm_celt_job_t job = jobs[j];
    anm_celt_frame_t fr;
    int16_t qi[2 * ANM_CE_NB], eoff[2 * ANM_CE_NB];
    const int rc = anm_celt_entropy_symbols(t, bytes, mask, job.offset, job.len, job.channels, job.lm, job.end_band, qi, eoff, &fr);
    if (rc != 0) { /* impossible job description: treated like a lost frame, flagged */
        fr.final_range = 0;
        fr.flags = ANM_CELT_F_LOST | ANM_CELT_F_EC_ERROR;
    }
    int16_t *sc = scratch + (size_t)j * (4 * ANM_CE_NB);
    for (int i = 0; i < 2 * ANM_CE_NB; ++i) {
        sc[i] = qi[i];
        sc[2 * ANM_CE_NB + i] = eoff[i];
    }
    out[j] = fr;
}

/* pass 2, one thread per STREAM: the band energies predict from frame to frame (celt/quant_bands.c:427-490) -- a short recurrence over the
 * stream's frames in order, about a hundred integer operations per frame */
__global__ void __launch_bounds__(128) k_celt_energies(const uint32_t *__restrict__ stream_begin, uint32_t n_streams, const int16_t *__restrict__ scratch,
                                                       anm_celt_stream_t *streams, anm_celt_frame_t *out) {
    const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n_streams) return;
    anm_celt_stream_t st = streams[s];
    for (uint32_t j = stream_begin[s]; j < stream_begin[s + 1]; ++j) {
        const int16_t *sc = scratch + (size_t)j * (4 * ANM_CE_NB);
        anm_celt_apply_energies(&out[j], sc, sc + 2 * ANM_CE_NB, st.old_e);
    }
    streams[s] = st;
}

} /* namespace */

extern "C" int anm_celt_ctx_create(int device, anm_celt_ctx_t **out) {
    if (!out) return ANM_ERR_ARG;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        anm_set_error("no CUDA device: the CELT entropy decoder has no CPU fallback");
        return ANM_ERR_CUDA;
    }
    if (device < 0 || device >= ndev) return ANM_ERR_ARG;
    anm_celt_ctx *c = new (std::nothrow) anm_celt_ctx();
    anm_celt_tables_t *h = new (std::nothrow) anm_celt_tables_t();
    if (!c || !h) { delete c; delete h; return ANM_ERR_NOMEM; }
    c->device = device;
    c->d_tables = nullptr;
    c->d_scratch = nullptr;
    c->scratch_frames = 0;
    int rc = anm_celt_tables_build(h);
    if (rc == ANM_OK && (cudaSetDevice(device) != cudaSuccess || cudaMalloc(&c->d_tables, sizeof *h) != cudaSuccess ||
                         cudaMemcpy(c->d_tables, h, sizeof *h, cudaMemcpyHostToDevice) != cudaSuccess)) {
        anm_set_error("anm_celt_ctx_create: %s", cudaGetErrorString(cudaGetLastError()));
        rc = ANM_ERR_CUDA;
    }
    delete h;
    if (rc != ANM_OK) {
        cudaFree(c->d_tables);
        delete c;
        return rc;
    }
    /* the band splitting recurses (at most five levels deep); give the threads room for it */
    size_t lim = 0;
    if (cudaDeviceGetLimit(&lim, cudaLimitStackSize) == cudaSuccess && lim < 8192) cudaDeviceSetLimit(cudaLimitStackSize, 8192);
    *out = c;
    return ANM_OK;
}

extern "C" void anm_celt_ctx_destroy(anm_celt_ctx_t *c) {
    if (!c) return;
    cudaSetDevice(c->device);
    cudaFree(c->d_tables);
    cudaFree(c->d_scratch);
    delete c;
}

extern "C" int anm_celt_entropy_device(anm_celt_ctx_t *c, const anm_celt_job_t *d_jobs, const uint32_t *d_stream_begin, uint32_t n_streams, uint32_t n_jobs,
                                       const uint8_t *d_bytes, uint32_t bytes_mask, anm_celt_stream_t *d_streams, anm_celt_frame_t *d_out, void *stream) {
    if (!c || ((!d_jobs || !d_stream_begin || !d_streams || !d_out) && n_streams)) return ANM_ERR_ARG;
    if (n_streams == 0 || n_jobs == 0) return ANM_OK;
    if (bytes_mask != 0xFFFFFFFFu && (bytes_mask & (bytes_mask + 1u)) != 0u) return ANM_ERR_ARG;
    cudaStream_t s = (cudaStream_t)stream;
    if (n_jobs > c->scratch_frames) {
        if (cudaStreamSynchronize(s) != cudaSuccess) { anm_set_error("anm_celt_entropy_device: %s", cudaGetErrorString(cudaGetLastError())); return ANM_ERR_CUDA; }
        cudaFree(c->d_scratch);
        c->d_scratch = nullptr;
        c->scratch_frames = 0;
        if (cudaMalloc(&c->d_scratch, (size_t)n_jobs * 4 * ANM_CE_NB * sizeof(int16_t)) != cudaSuccess) {
            anm_set_error("anm_celt_entropy_device: out of device memory for %u frames", n_jobs);
            cudaGetLastError();
            return ANM_ERR_NOMEM;
        }
        c->scratch_frames = n_jobs;
    }
    k_celt_entropy<<<(n_jobs + 127u) / 128u, 128, 0, s>>>(c->d_tables, d_jobs, n_jobs, d_bytes, bytes_mask, c->d_scratch, d_out);
    k_celt_energies<<<(n_streams + 127u) / 128u, 128, 0, s>>>(d_stream_begin, n_streams, c->d_scratch, d_streams, d_out);
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        anm_set_error("k_celt_entropy launch failed: %s", cudaGetErrorString(e));
        return ANM_ERR_CUDA;
    }
    return ANM_OK;
}

extern "C" int anm_celt_entropy_host(const anm_celt_job_t *jobs, const uint32_t *stream_begin, uint32_t n_streams, const uint8_t *bytes, size_t n_bytes,
                                     anm_celt_stream_t *streams, anm_celt_frame_t *out) {
    if ((!jobs || !stream_begin || !streams || !out) && n_streams) return ANM_ERR_ARG;
    if (n_streams == 0) return ANM_OK;
    const uint32_t n_jobs = stream_begin[n_streams];
    for (uint32_t i = 0; i < n_jobs; ++i)
        if ((size_t)jobs[i].offset + jobs[i].len > n_bytes) return ANM_ERR_ARG;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) {
        cudaGetLastError();
        anm_set_error("no CUDA device: the CELT entropy decoder has no CPU fallback");
        return ANM_ERR_CUDA;
    }
    anm_celt_ctx_t *c = nullptr;
    int rc = anm_celt_ctx_create(dev, &c);
    if (rc != ANM_OK) return rc;
    anm_celt_job_t *d_j = nullptr;
    uint32_t *d_sb = nullptr;
    uint8_t *d_b = nullptr;
    anm_celt_stream_t *d_s = nullptr;
    anm_celt_frame_t *d_o = nullptr;
    rc = ANM_ERR_CUDA;
    if (cudaMalloc(&d_j, (n_jobs ? n_jobs : 1) * sizeof *d_j) == cudaSuccess && cudaMalloc(&d_sb, (n_streams + 1) * sizeof *d_sb) == cudaSuccess &&
        cudaMalloc(&d_b, n_bytes ? n_bytes : 1) == cudaSuccess && cudaMalloc(&d_s, n_streams * sizeof *d_s) == cudaSuccess &&
        cudaMalloc(&d_o, (n_jobs ? n_jobs : 1) * sizeof *d_o) == cudaSuccess &&
        cudaMemcpy(d_j, jobs, n_jobs * sizeof *d_j, cudaMemcpyHostToDevice) == cudaSuccess &&
        cudaMemcpy(d_sb, stream_begin, (n_streams + 1) * sizeof *d_sb, cudaMemcpyHostToDevice) == cudaSuccess &&
        cudaMemcpy(d_b, bytes, n_bytes, cudaMemcpyHostToDevice) == cudaSuccess &&
        cudaMemcpy(d_s, streams, n_streams * sizeof *d_s, cudaMemcpyHostToDevice) == cudaSuccess) {
        rc = anm_celt_entropy_device(c, d_j, d_sb, n_streams, n_jobs, d_b, 0xFFFFFFFFu, d_s, d_o, nullptr);
        if (rc == ANM_OK && (cudaDeviceSynchronize() != cudaSuccess || cudaMemcpy(out, d_o, n_jobs * sizeof *d_o, cudaMemcpyDeviceToHost) != cudaSuccess ||
                             cudaMemcpy(streams, d_s, n_streams * sizeof *d_s, cudaMemcpyDeviceToHost) != cudaSuccess))
            rc = ANM_ERR_CUDA;
    }
    if (rc == ANM_ERR_CUDA) anm_set_error("anm_celt_entropy_host: %s", cudaGetErrorString(cudaGetLastError()));
    cudaFree(d_j);
    cudaFree(d_sb);
    cudaFree(d_b);
    cudaFree(d_s);
    cudaFree(d_o);
    anm_celt_ctx_destroy(c);
    return rc;
}
