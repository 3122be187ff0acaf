/* cenn_debug.h -- bring-up / measurement probes exported by libcenn.so for tools/*probe*.py.  NOT part of the drop-in boundary
 * (include/cenn.h): no reference interface corresponds to them and nothing on the product path calls them. */
#ifndef CENN_DEBUG_H
#define CENN_DEBUG_H
#include "cenn.h"
#ifdef __cplusplus
extern "C" {
#endif
/* one conv / GEMM primitive on zero operands: mean duration and per-role cycle counters (tools/gemm_probe.py, tools/gemm_big_probe.py) */
CENN_API int cenn_debug_gemm_probe(cenn_state *s, int kind, int N, int h, int w, int Cs, int Cl, int with_stats, int act,
                                   int iters, float *ms_out, unsigned long long *dbg_out);
/* UMMA shared-memory descriptor semantics (tools/desc_probe.py) */
CENN_API int cenn_debug_desc_probe(cenn_state *s, const uint16_t *a_host, int start_row, int base_off, int sbo_bytes, float *out_host);
/* CTA-pair (tcgen05.mma.cta_group::2) plain GEMM (tools/gemm2sm_probe.py) */
CENN_API int cenn_debug_gemm2sm_probe(cenn_state *s, const uint16_t *a_host, const uint16_t *b_host, int M, int N, int K, int iters,
                                      float *c_host, float *ms_out);
/* one TMA box of the implicit-im2col map of a bordered thin tensor, raw shared-memory bytes (tools/tma_thin_probe.py) */
CENN_API int cenn_debug_tma_thin_probe(cenn_state *s, const uint16_t *lpad_host, int N, int H2, int W2, int Cp, int bw, int bh, int bn,
                                       int c1, int x0, int y0, int n0, uint8_t *out_host);
#ifdef __cplusplus
}
#endif
#endif
