// Blackwell (sm_100a) primitives used by the fused head: mbarrier, TMA (cp.async.bulk.tensor),
// tcgen05.mma / TMEM, and the shared-memory / instruction descriptors, as thin inline-PTX wrappers.
//
// Conventions used throughout libosb:
//   * operands live in shared memory in the 128-byte-swizzle layout TMA writes (CU_TENSOR_MAP_SWIZZLE_128B);
//     a TMA box is always 64 bf16 (128 bytes) wide in its contiguous dimension and R rows tall, i.e. R rows
//     of 128 bytes, 8-row swizzle atoms of 1024 bytes, box base 1024-byte aligned;
//   * the same box is a K-major operand tile (rows = M/N index, 64 contiguous K elements) or an MN-major
//     operand tile (rows = K index, 64 contiguous M/N elements) — only the descriptor differs;
//   * accumulators are fp32 in TMEM, lane = output row, column = output column (UMMA M = 128, cta_group::1).
#pragma once

#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace osb {
namespace tc {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---- mbarrier --------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.b32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) { }
}
// Whole-warp wait for long waits (an epilogue warp waiting for a tile): ONE lane polls, with a back-off, and the rest
// of the warp parks on the warp barrier.  128+ threads spinning on try_wait compete with the tensor core's operand
// reads for shared-memory bandwidth (measured: a k-block of 4 MMAs took 727 cycles instead of 573).
__device__ __forceinline__ void mbar_wait_warp(uint64_t* bar, uint32_t parity) {
    if ((threadIdx.x & 31) == 0) {
        while (!mbar_try_wait(bar, parity)) __nanosleep(40);
    }
    __syncwarp();
}

// ---- TMA -------------------------------------------------------------------------------------
__device__ __forceinline__ void prefetch_tensormap(const CUtensorMap* m) {
    asm volatile("prefetch.tensormap [%0];" :: "l"(m) : "memory");
}
// global -> shared, 2-D tile at element coordinates (c0 = contiguous dim, c1 = row); completes on `bar`.
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        :: "r"(smem_u32(smem_dst)), "l"(m), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
// 3-D variant (c2 = plane): the grad-logits planes [plane][row][class]
__device__ __forceinline__ void tma_load_3d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        :: "r"(smem_u32(smem_dst)), "l"(m), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
// global -> L2 only (no shared-memory destination, no completion): warms L2 for a tile that will be loaded later
__device__ __forceinline__ void tma_prefetch_l2_2d(const CUtensorMap* m, int c0, int c1) {
    asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global.tile [%0, {%1, %2}];" :: "l"(m), "r"(c0), "r"(c1) : "memory");
}
// shared -> global, 2-D tile (bulk async group).
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* m, const void* smem_src, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                 :: "l"(m), "r"(smem_u32(smem_src)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* m, const void* smem_src, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
                 :: "l"(m), "r"(smem_u32(smem_src)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
// shared -> global with an fp32 ADD at the destination (performed at L2): the split-K partner's contribution
__device__ __forceinline__ void tma_reduce_add_2d(const CUtensorMap* m, const void* smem_src, int c0, int c1) {
    asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.bulk_group [%0, {%2, %3}], [%1];"
                 :: "l"(m), "r"(smem_u32(smem_src)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void tma_store_wait_read() {
    asm volatile("cp.async.bulk.wait_group.read %0;" :: "n"(N) : "memory");
}
template <int N> __device__ __forceinline__ void tma_store_wait() {
    asm volatile("cp.async.bulk.wait_group %0;" :: "n"(N) : "memory");
}

// ---- TMEM ------------------------------------------------------------------------------------
// One full warp allocates `ncols` (power of two, 32..512) columns; the base address lands in *smem_dst.
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                 :: "r"(smem_u32(smem_dst)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// 32 lanes x 32 consecutive fp32 columns: thread t of the warp receives lane (taddr.lane + t).
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
          "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
          "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr) : "memory");
}
// registers -> TMEM, same shape as tmem_ld32 (used to park exp(z - max) between epilogue sweeps)
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
        :: "r"(taddr),
           "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]),
           "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]),
           "r"(v[16]), "r"(v[17]), "r"(v[18]), "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]),
           "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31])
        : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// TMEM address of (lane, column) relative to an allocation base.
__device__ __forceinline__ uint32_t tmem_addr(uint32_t base, uint32_t lane, uint32_t col) {
    return base + (lane << 16) + col;
}

// ---- tcgen05.mma -----------------------------------------------------------------------------
// D[tmem] (+)= A[smem] * B[smem]; bf16 inputs, fp32 accumulate; issued by ONE thread.
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, bool accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        :: "r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"((uint32_t)accumulate) : "memory");
}
// Arrive on `bar` when every MMA issued so far by this thread has completed (implies fence::before_thread_sync).
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];"
                 :: "r"(smem_u32(bar)) : "memory");
}

// ---- CTA pair (cta_group::2): two CTAs of a cluster share one M = 256 MMA --------------------------------------
// CTA r of the pair keeps rows [128 r, 128 r + 128) of the A tile and its half of the B tile (N / 2 rows) in ITS
// shared memory at the same offsets; the leader (cluster rank 0) issues the MMAs, which read both CTAs' shared
// memory, and each CTA receives its 128 accumulator rows x N columns in its own TMEM.  Per CTA that is 2/3 of the
// shared-memory traffic of the single-CTA M = 128 path (TMA writes + operand reads), which is what bounds it.
constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;       // clears the CTA-rank bit of a shared::cluster address: -> leader CTA

__device__ __forceinline__ uint32_t cluster_ctarank_() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// one warp of EACH CTA of the pair executes the allocation / deallocation
__device__ __forceinline__ void tmem_alloc2(uint32_t* smem_dst, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(smem_dst)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" :: "r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma2_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, bool accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        :: "r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"((uint32_t)accumulate) : "memory");
}
// Arrive on the barrier at this shared-memory offset in BOTH CTAs of the pair once every MMA issued so far is done.
__device__ __forceinline__ void umma2_commit_mc(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 :: "r"(smem_u32(bar)), "h"((uint16_t)3) : "memory");
}
// 2-D tile load into THIS CTA's shared memory; the bytes are counted on the LEADER CTA's barrier (same offset).
__device__ __forceinline__ void tma_load_2d_2sm(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 :: "r"(smem_u32(smem_dst)), "l"(m), "r"(smem_u32(bar) & kPeerBitMask), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_load_3d_2sm(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                 :: "r"(smem_u32(smem_dst)), "l"(m), "r"(smem_u32(bar) & kPeerBitMask), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
// Arrive on the LEADER CTA's barrier at this offset (from either CTA of the pair).  RELAXED: the barrier only hands a TMEM
// accumulator back to the MMA warp, and the tcgen05 loads that read it are ordered by tcgen05.wait::ld +
// tcgen05.fence::before_thread_sync.  The default .release.cluster compiles to MEMBAR.ALL.GPU in front of the arrive,
// which waits for every global store the warp has in flight: ~2 us per tile in the epilogue warps (measured).
__device__ __forceinline__ void mbar_arrive_leader(uint64_t* bar) {
    asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" :: "r"(smem_u32(bar) & kPeerBitMask) : "memory");
}

// Instruction descriptor, kind::f16: bf16 x bf16 -> fp32, M x N tile, operand majors.
__host__ __device__ constexpr uint32_t make_idesc_bf16(int M, int N, bool a_mn_major, bool b_mn_major) {
    return (1u << 4)                       // D format: F32
         | (1u << 7)                       // A format: BF16
         | (1u << 10)                      // B format: BF16
         | ((a_mn_major ? 1u : 0u) << 15)
         | ((b_mn_major ? 1u : 0u) << 16)
         | ((uint32_t)(N >> 3) << 17)
         | ((uint32_t)(M >> 4) << 24);
}

// Shared-memory matrix descriptors for 128-byte-swizzled tiles (layout_type 2, descriptor version 1).
//   K-major : rows of 64 K-elements; 8-row atoms 1024 B apart (SBO); LBO unused.
//             Advance along K by 16 elements = +32 bytes on the start address.
//   MN-major: rows of 64 M/N-elements, one row per K index; 8-K atoms 1024 B apart (SBO); the next group
//             of 64 M/N elements is `lbo_bytes` away (the next TMA box).  Advance along K by 16 = +2048 bytes.
__device__ __forceinline__ uint64_t make_desc_sw128(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
    d |= (uint64_t)1 << 46;                // descriptor version (Blackwell)
    d |= (uint64_t)2 << 61;                // SWIZZLE_128B
    return d;
}
__device__ __forceinline__ uint64_t make_desc_kmajor(uint32_t smem_addr) { return make_desc_sw128(smem_addr, 0, 1024); }
__device__ __forceinline__ uint64_t make_desc_mnmajor(uint32_t smem_addr, uint32_t lbo_bytes) {
    return make_desc_sw128(smem_addr, lbo_bytes, 1024);
}

__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "elect.sync _|p, 0xffffffff;\n\t"
        "selp.b32 %0, 1, 0, p;\n\t}"
        : "=r"(pred));
    return pred != 0;
}

}  // namespace tc

// ---- host: TMA descriptor encode through the driver entry point (no link-time libcuda dependency) ----
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static inline PFN_encodeTiled get_encode_tiled() {
    static PFN_encodeTiled fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<PFN_encodeTiled>(p);
    }
    return fn;
}

// 2-D bf16 tensor [rows, cols] with row pitch `ld` elements; box = 64 columns x box_rows rows, SWIZZLE_128B,
// out-of-bounds elements read as zero.  Returns 0 on success.  Encoding costs ~1-2 us of host time, which is
// what bounds the small-batch configurations, so the last few descriptors are kept in a thread-local cache
// (a descriptor depends only on the arguments below).
static inline int make_tmap_bf16(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t ld,
                                 uint32_t box_rows, uint32_t box_cols = 64) {
    struct Entry { const void* base; uint64_t rows, cols, ld; uint32_t box_rows, box_cols; bool valid; CUtensorMap map; };
    constexpr int kSlots = 32;
    static thread_local Entry cache[kSlots];
    const uint64_t h = (reinterpret_cast<uint64_t>(base) >> 8) ^ (rows * 0x9E3779B97F4A7C15ull) ^ (cols << 20) ^ (ld << 40) ^ box_rows;
    Entry& e = cache[(h ^ (h >> 17) ^ (h >> 31)) % kSlots];
    if (e.valid && e.base == base && e.rows == rows && e.cols == cols && e.ld == ld && e.box_rows == box_rows && e.box_cols == box_cols) {
        *out = e.map;
        return 0;
    }
    PFN_encodeTiled enc = get_encode_tiled();
    if (!enc) return -1;
    cuuint64_t dims[2] = {cols, rows};
    cuuint64_t strides[1] = {ld * 2};
    cuuint32_t box[2] = {box_cols, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return (int)r;
    e.base = base; e.rows = rows; e.cols = cols; e.ld = ld; e.box_rows = box_rows; e.box_cols = box_cols; e.map = *out; e.valid = true;
    return 0;
}

// 3-D bf16 tensor [planes, rows, cols] (row pitch `ld`, plane pitch `plane_ld` elements); box = 64 columns x box_rows
// rows x 1 plane, SWIZZLE_128B: the grad-logits planes (plane 1 = bf16 rounding residual of plane 0).  Cached like the
// 2-D descriptors.
static inline int make_tmap_bf16_3d(CUtensorMap* out, const void* base, uint64_t planes, uint64_t rows, uint64_t cols, uint64_t ld,
                                    uint64_t plane_ld, uint32_t box_rows) {
    struct Entry { const void* base; uint64_t planes, rows, cols, ld, plane_ld; uint32_t box_rows; bool valid; CUtensorMap map; };
    constexpr int kSlots = 16;
    static thread_local Entry cache[kSlots];
    const uint64_t h = (reinterpret_cast<uint64_t>(base) >> 8) ^ (rows * 0x9E3779B97F4A7C15ull) ^ (cols << 20) ^ (ld << 40) ^ box_rows ^ (planes << 7);
    Entry& e = cache[(h ^ (h >> 17) ^ (h >> 31)) % kSlots];
    if (e.valid && e.base == base && e.planes == planes && e.rows == rows && e.cols == cols && e.ld == ld && e.plane_ld == plane_ld &&
        e.box_rows == box_rows) {
        *out = e.map;
        return 0;
    }
    PFN_encodeTiled enc = get_encode_tiled();
    if (!enc) return -1;
    cuuint64_t dims[3] = {cols, rows, planes};
    cuuint64_t strides[2] = {ld * 2, plane_ld * 2};
    cuuint32_t box[3] = {64, box_rows, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return (int)r;
    e.base = base; e.planes = planes; e.rows = rows; e.cols = cols; e.ld = ld; e.plane_ld = plane_ld; e.box_rows = box_rows;
    e.map = *out; e.valid = true;
    return 0;
}

// 2-D fp32 tensor [rows, cols] (row pitch `ld` elements); box = 32 columns (128 bytes) x box_rows rows, SWIZZLE_128B: dX in fp32
static inline int make_tmap_f32_2d(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows) {
    PFN_encodeTiled enc = get_encode_tiled();
    if (!enc) return -1;
    cuuint64_t dims[2] = {cols, rows};
    cuuint64_t strides[1] = {ld * 4};
    cuuint32_t box[2] = {32, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : (int)r;
}

// 3-D fp32 tensor [planes, rows, cols] (row pitch `ld` elements, plane pitch rows * ld); box = 32 columns (128 bytes) x
// box_rows rows x 1 plane, SWIZZLE_128B: the split-K partial matrices of the dW GEMM, stored tile by tile.
static inline int make_tmap_f32_3d(CUtensorMap* out, const void* base, uint64_t planes, uint64_t rows, uint64_t cols, uint64_t ld,
                                   uint32_t box_rows) {
    PFN_encodeTiled enc = get_encode_tiled();
    if (!enc) return -1;
    cuuint64_t dims[3] = {cols, rows, planes};
    cuuint64_t strides[2] = {ld * 4, rows * ld * 4};
    cuuint32_t box[3] = {32, box_rows, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<void*>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : (int)r;
}

}  // namespace osb
