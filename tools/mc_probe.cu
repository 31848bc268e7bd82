// Operand-fill probe for B200 (developer tool): what does the L2 -> SM path deliver when every SM streams GEMM operand
// tiles (no MMAs at all), and does TMA multicast of the operand two CTAs share raise it?
// Clusters of 4 CTAs (= two CTA pairs of the head kernels).  Per pipeline step every CTA receives 16 KB of a SHARED
// operand (CTA r and CTA r ^ 2 of the cluster want the same bytes) and 16 KB of its own:
//   mode 0: both fetch the shared 16 KB themselves (unicast; what the head kernels do today),
//   mode 1: each of the two fetches one 8 KB half and multicasts it to both (half the L2 reads of the shared operand),
//   mode 2: only the 16 KB of the private operand (reference: no sharing possible),
//   mode 3: like 1 but shared by all four CTAs of the cluster (each fetches 4 KB, multicast to four).
// Prints delivered TB/s (bytes written into shared memory) and the L2 read rate for several grid sizes.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o build/mc_probe tools/mc_probe.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s failed: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

constexpr int kStages = 4;
constexpr uint32_t kChunk = 16384;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_remote(uint64_t* bar, uint32_t cta) {
    uint32_t ra;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(smem_u32(bar)), "r"(cta));
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" :: "r"(ra) : "memory");
}
__device__ __forceinline__ bool mbar_try(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void bulk_g2s_mc(void* dst, const void* src, uint32_t bytes, uint64_t* bar, uint16_t mask) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;"
                 :: "r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)), "h"(mask) : "memory");
}
__device__ __forceinline__ void cluster_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}

__global__ void __cluster_dims__(4, 1, 1) __launch_bounds__(64, 1)
fill_kernel(const uint8_t* __restrict__ shared_src, const uint8_t* __restrict__ own_src, int iters, int mode,
            size_t shared_mask, size_t own_mask) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + (size_t)kStages * 2 * kChunk);
    uint64_t* empty = full + kStages;
    uint32_t rank;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
    const uint32_t cluster = blockIdx.x >> 2;
    const int sharers = mode == 1 ? 2 : (mode == 3 ? 4 : 1);
    if (threadIdx.x == 0) {
        for (int s = 0; s < kStages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], sharers); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    cluster_sync();
    const uint32_t total = (mode == 2 ? 1u : 2u) * kChunk;
    if (threadIdx.x == 0) {
        for (int it = 0; it < iters; ++it) {
            const int s = it % kStages;
            const uint32_t ph = (it / kStages) & 1;
            { uint32_t spins = 0; while (!mbar_try(&empty[s], ph ^ 1)) { if (++spins > (1u << 24)) __trap(); } }
            mbar_expect(&full[s], total);
            uint8_t* st = smem + (size_t)s * 2 * kChunk;
            // the shared operand: the same address for CTA r and r ^ 2 (mode 0 / 1), for all four (mode 3)
            const size_t grp = (mode == 3) ? 0 : (rank & 1);
            const size_t a_off = ((((size_t)cluster * iters + it) * 2 + grp) * kChunk) & shared_mask;
            const size_t b_off = (((size_t)blockIdx.x * iters + it) * kChunk) & own_mask;
            if (mode == 0) {
                bulk_g2s(st, shared_src + a_off, kChunk, &full[s]);
            } else if (mode == 1) {
                const uint32_t half = rank >> 1, piece = kChunk / 2;
                bulk_g2s_mc(st + half * piece, shared_src + a_off + half * piece, piece, &full[s], (uint16_t)((1u << rank) | (1u << (rank ^ 2))));
            } else if (mode == 3) {
                const uint32_t piece = kChunk / 4;
                bulk_g2s_mc(st + rank * piece, shared_src + a_off + rank * piece, piece, &full[s], (uint16_t)0xF);
            }
            bulk_g2s(st + kChunk, own_src + b_off, kChunk, &full[s]);
        }
    } else if (threadIdx.x == 32) {
        for (int it = 0; it < iters; ++it) {
            const int s = it % kStages;
            const uint32_t ph = (it / kStages) & 1;
            { uint32_t spins = 0; while (!mbar_try(&full[s], ph)) { if (++spins > (1u << 24)) __trap(); } }
            if (mode == 1) { mbar_arrive(&empty[s]); mbar_arrive_remote(&empty[s], rank ^ 2); }
            else if (mode == 3) { for (uint32_t c = 0; c < 4; ++c) { if (c == rank) mbar_arrive(&empty[s]); else mbar_arrive_remote(&empty[s], c); } }
            else mbar_arrive(&empty[s]);
        }
    }
    __syncthreads();
    cluster_sync();          // nobody leaves while a partner may still multicast into it / arrive on its barriers
}

int main(int argc, char** argv) {
    const int iters = argc > 1 ? atoi(argv[1]) : 4000;
    const size_t shared_bytes = 32u << 20, own_bytes = 64u << 20;     // both stay L2 resident (126 MB)
    uint8_t *a, *b;
    CK(cudaMalloc(&a, shared_bytes + kChunk * 4));
    CK(cudaMalloc(&b, own_bytes + kChunk * 4));
    CK(cudaMemset(a, 1, shared_bytes + kChunk * 4));
    CK(cudaMemset(b, 2, own_bytes + kChunk * 4));
    const size_t smem = (size_t)kStages * 2 * kChunk + 2 * kStages * 8 + 1024;
    CK(cudaFuncSetAttribute(fill_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(148); cfg.blockDim = dim3(64); cfg.dynamicSmemBytes = smem;
        cudaLaunchAttribute at; at.id = cudaLaunchAttributeClusterDimension; at.val.clusterDim.x = 4; at.val.clusterDim.y = 1; at.val.clusterDim.z = 1;
        cfg.attrs = &at; cfg.numAttrs = 1;
        int n = 0;
        // the head kernels keep one CTA per SM: ask with their shared-memory footprint
        CK(cudaFuncSetAttribute(fill_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 215 * 1024));
        cfg.dynamicSmemBytes = 215 * 1024;
        CK(cudaOccupancyMaxActiveClusters(&n, fill_kernel, &cfg));
        printf("co-resident clusters of 4 at one CTA per SM (215 KB shared memory): %d (= %d SMs)\n", n, 4 * n);
    }
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    const int grids[] = {16, 64, 128, 144};
    for (int gi = 0; gi < 4; ++gi) {
        for (int mode = 0; mode < 4; ++mode) {
            const int grid = grids[gi];
            float best = 1e30f;
            for (int rep = 0; rep < 4; ++rep) {
                CK(cudaEventRecord(e0));
                fill_kernel<<<grid, 64, 215 * 1024>>>(a, b, iters, mode, shared_bytes - 1, own_bytes - 1);
                CK(cudaEventRecord(e1));
                CK(cudaEventSynchronize(e1));
                CK(cudaGetLastError());
                float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
                if (rep > 0 && ms < best) best = ms;
            }
            const double per_cta = (mode == 2 ? 1.0 : 2.0) * kChunk;
            const double delivered = per_cta * grid * iters;
            const double l2reads = (mode == 0 ? 2.0 : mode == 1 ? 1.5 : mode == 3 ? 1.25 : 1.0) * kChunk * grid * iters;
            printf("grid %3d mode %d: %8.3f ms  delivered %6.2f TB/s  (L2 reads %6.2f TB/s)  %6.1f ns per 32 KB step\n", grid, mode, best,
                   delivered / best * 1e-9, l2reads / best * 1e-9, best * 1e6 / iters);
        }
    }
    return 0;
}
