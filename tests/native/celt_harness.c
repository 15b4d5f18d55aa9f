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
