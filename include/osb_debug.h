/*
 * osb_debug.h — developer hooks of libosb.so: timing probes, phase-stamp traces and A/B switches between kernel
 * variants.  They set process-wide state and are meant for tools/ and for tests that compare kernel variants;
 * the product path (the Python package) and bench.py never call them.  Not part of the stable ABI.
 */
#ifndef OSB_DEBUG_H_
#define OSB_DEBUG_H_

#ifdef __cplusplus
extern "C" {
#endif

void osb_debug_set_head_splitk(int on);            /* 0: small batches take the persistent forward kernel too */
int  osb_debug_head_splitk_max_clusters(void);     /* clusters of 8 CTAs resident at once (split-K forward) */
void osb_debug_set_head_dbg(int bits);             /* dX probes: 1 stale operands, 2 no TMEM drain, 8 no stores */
void osb_debug_set_head_trace(void* device_buf);   /* [grid][64] int64 phase stamps of the head kernels, or NULL */
void osb_debug_set_dw_trace(void* device_buf);     /* [grid][8] int64 stamps of the dW GEMM alone (tools/timeline.py), or NULL */
void osb_debug_set_allreduce_timeout(long long cycles);   /* give-up time of the peer kernels, 0 = default (~2 s) */
void osb_debug_set_exchange_timeout(long long cycles);    /* same for osb_oscr_exchange */
void osb_debug_set_curve_trace(void* device_buf);  /* [grid][32] int64 SM clock stamps of oscr_curve_kernel, or NULL */
void osb_debug_set_records_mode(int bits);         /* stage-A probes: 1 no records emitted, 2 no row loop */

#ifdef __cplusplus
}
#endif
#endif /* OSB_DEBUG_H_ */
