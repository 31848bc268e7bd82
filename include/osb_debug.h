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

/*
 * Environment switches read once per process (A/B measurements of launch forms; every variant gives the same bits):
 *   OSB_NO_PDL=1          no programmatic dependent launches (plain stream order between the kernels of a step)
 *   OSB_NO_COOP_LAUNCH=1  the persistent forward without the cooperative-launch attribute (Nsight Compute cannot replay
 *                         a launch that is both cooperative and clustered)
 *   OSB_NO_DIRECT_DW=1    dW always as split-K partials + reduce kernel, never written by the GEMM itself
 *   OSB_NO_DW_PREZERO=1   two-split dW by the GEMM as "split 0 stores, split 1 adds after it" even when dX and dW come from one
 *                         call (default there: dX's kernel clears dW, both splits add)
 *   OSB_NO_FUSED_DX=1     osb_head_forward_dx always as two kernels
 *   OSB_NO_ROW_FLAGS=1    OSB_FWD_PUBLISH_ROW_TILES / OSB_BWD_DZ_FROM_PREV_FORWARD are ignored
 */

#ifdef __cplusplus
}
#endif
#endif /* OSB_DEBUG_H_ */
