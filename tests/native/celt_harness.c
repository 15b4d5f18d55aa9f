/*
 * celt_harness.c -- TEST INFRASTRUCTURE: compiles the product's entropy-decode header (audio-network_b200/csrc/
 * anm_celt_entropy.h) and table builder for the HOST so that the logic can be checked against the reference without a GPU
 * (`-m "not gpu"` tests).  The product itself runs this code only inside the CUDA kernel of anm_celt_gpu.cu; this library is
 * built by tests/test_celt_entropy.py into tests/native/_build and is never loaded by the package.
 */
#include "../../audio-network_b200/csrc/anm_celt_entropy.h"

int harness_celt_frame(const anm_celt_tables_t *t, const uint8_t *bytes, uint32_t len, int C, int LM, int end, int16_t *old_e, anm_celt_frame_t *out) {
    return anm_celt_entropy_frame(t, bytes, 0xFFFFFFFFu, 0u, len, C, LM, end, old_e, out);
}
uint32_t harness_sizeof_tables(void) { return (uint32_t)sizeof(anm_celt_tables_t); }
uint32_t harness_sizeof_frame(void) { return (uint32_t)sizeof(anm_celt_frame_t); }

/* the host runs the spectrum code with ONE lane */
static void spec_init(ce_spec_t *sp) {
    static int16_t norm[CE_SPEC_NORM], tmp[CE_SPEC_TMP];
    static int iy[CE_SPEC_IY];
    static int16_t band[CE_SPEC_BAND];
    sp->band = band; /* as the kernel runs it */
    sp->norm = norm;
    sp->tmp = tmp;
    sp->iy = iy;
    sp->lane = 0;
    sp->nl = 1;
    sp->seed = 0;
    sp->spread = 0;
    sp->disable_inv = 0;
}

/* stage 2: the frame's normalised spectrum (X: [2][960], channel c at X + 960 c ... repacked from the product's [C][120 << LM] layout) */
int harness_celt_spectrum(const anm_celt_tables_t *t, const uint8_t *bytes, uint32_t len, int C, int LM, int end, int16_t *old_e, anm_celt_frame_t *out,
                          uint32_t seed_in, int disable_inv, int16_t *x_out, uint8_t *cm_out, uint32_t *seed_out) {
    ce_spec_t sp;
    static int16_t X[2 * 960];
    int16_t qi[2 * ANM_CE_NB], eoff[2 * ANM_CE_NB];
    uint8_t cm[2 * ANM_CE_NB];
    const int NF = 120 << LM;
    spec_init(&sp);
    for (int i = 0; i < 2 * 960; i++) X[i] = 0;
    for (int i = 0; i < 2 * ANM_CE_NB; i++) cm[i] = 0;
    sp.seed = seed_in;
    sp.disable_inv = disable_inv;
    const int rc = anm_celt_frame_symbols(t, bytes, 0xFFFFFFFFu, 0u, len, C, LM, end, qi, eoff, out, &sp, X, cm, 0);
    if (rc != 0) return rc;
    anm_celt_apply_energies(out, qi, eoff, old_e);
    const int ncmp = (1 << LM) * t->ebands[end];
    for (int i = 0; i < 2 * 960; i++) x_out[i] = 0;
    for (int c = 0; c < C; c++)
        for (int i = 0; i < ncmp; i++) x_out[960 * c + i] = X[NF * c + i];
    for (int i = 0; i < 2 * ANM_CE_NB; i++) cm_out[i] = cm[i];
    *seed_out = sp.seed;
    return 0;
}

/* stages 1 + 2 for one frame of a stream, exactly as the kernels chain them: symbols -> per-stream step (histories) -> spectrum + anti-collapse.
 * x_post: [2][960] repacked like x_out above; hist_out: the histories the frame saw (ce_hist_t). */
int harness_celt_frame_full(const anm_celt_tables_t *t, const uint8_t *bytes, uint32_t len, int C, int LM, int end, int disable_inv, anm_celt_stream_t *st,
                            anm_celt_frame_t *out, int16_t *x_post, uint8_t *cm_out, ce_hist_t *hist_out) {
    ce_spec_t sp;
    static int16_t X[2 * 960];
    int16_t qi[2 * ANM_CE_NB], eoff[2 * ANM_CE_NB];
    uint8_t cm[2 * ANM_CE_NB];
    const int NF = 120 << LM;
    spec_init(&sp);
    ce_resume_t resume;
    int rc = anm_celt_frame_symbols(t, bytes, 0xFFFFFFFFu, 0u, len, C, LM, end, qi, eoff, out, 0, 0, 0, &resume);
    if (rc != 0) return rc;
    anm_celt_stream_step(out, qi, eoff, st, hist_out);
    for (int i = 0; i < 2 * 960; i++) X[i] = 0;
    rc = anm_celt_frame_spectrum(t, bytes, 0xFFFFFFFFu, 0u, len, C, LM, end, disable_inv, hist_out, out, &resume, &sp, X, cm);
    if (rc != 0) return rc;
    { /* the same frame decoded from its first bit instead of from the resume point: must not differ */
        static int16_t X2[2 * 960];
        uint8_t cm2[2 * ANM_CE_NB];
        for (int i = 0; i < 2 * 960; i++) X2[i] = 0;
        rc = anm_celt_frame_spectrum(t, bytes, 0xFFFFFFFFu, 0u, len, C, LM, end, disable_inv, hist_out, out, 0, &sp, X2, cm2);
        if (rc != 0) return rc;
        if (!(out->flags & ANM_CELT_F_LOST)) {
            const int ncmp2 = (1 << LM) * t->ebands[end];
            for (int c = 0; c < C; c++)
                for (int i = 0; i < ncmp2; i++)
                    if (X2[NF * c + i] != X[NF * c + i]) return -99;
            for (int i = 0; i < C * end; i++)
                if (cm2[i] != cm[i]) return -98;
        }
    }
    const int ncmp = (1 << LM) * t->ebands[end];
    for (int i = 0; i < 2 * 960; i++) x_post[i] = 0;
    for (int c = 0; c < C; c++)
        for (int i = 0; i < ncmp; i++) x_post[960 * c + i] = X[NF * c + i];
    for (int i = 0; i < 2 * ANM_CE_NB; i++) cm_out[i] = cm[i];
    return 0;
}
uint32_t harness_sizeof_stream(void) { return (uint32_t)sizeof(anm_celt_stream_t); }
uint32_t harness_sizeof_hist(void) { return (uint32_t)sizeof(ce_hist_t); }

/* ---- stage 3 ---- */
#include "../../audio-network_b200/csrc/anm_celt_synth.h"
uint32_t harness_sizeof_synth(void) { return (uint32_t)sizeof(anm_celt_synth_t); }
uint32_t harness_sizeof_synth_tables(void) { return (uint32_t)sizeof(anm_celt_synth_tables_t); }

/* all three stages for one frame of a stream, chained as the kernels chain them; pcm: [120 << LM][CC] */
int harness_celt_decode_frame(const anm_celt_tables_t *t, const anm_celt_synth_tables_t *stb, const uint8_t *bytes, uint32_t len, int C, int CC, int LM, int end,
                              anm_celt_stream_t *st, anm_celt_synth_t *syn, anm_celt_frame_t *out, int16_t *pcm) {
    ce_spec_t sp;
    static int16_t X[2 * 960];
    static int32_t freq[2 * 960], raw[2 * 960];
    spec_init(&sp);
    int16_t qi[2 * ANM_CE_NB], eoff[2 * ANM_CE_NB];
    uint8_t cm[2 * ANM_CE_NB];
    ce_hist_t hist;
    ce_resume_t resume;
    int rc = anm_celt_frame_symbols(t, bytes, 0xFFFFFFFFu, 0u, len, C, LM, end, qi, eoff, out, 0, 0, 0, &resume);
    if (rc != 0) return rc;
    anm_celt_stream_step(out, qi, eoff, st, &hist);
    if (out->flags & ANM_CELT_F_LOST) return 0;
    for (int i = 0; i < 2 * 960; i++) X[i] = 0;
    rc = anm_celt_frame_spectrum(t, bytes, 0xFFFFFFFFu, 0u, len, C, LM, end, CC == 1, &hist, out, &resume, &sp, X, cm);
    if (rc != 0) return rc;
    cs_frame_blocks(t, stb, X, out->band_e, C, CC, LM, end, (out->flags & ANM_CELT_F_TRANSIENT) != 0, (out->flags & ANM_CELT_F_SILENCE) != 0, freq, raw, 0, 1);
    cs_stream_frame(stb, syn, out, raw, CC, pcm);
    return 0;
}
