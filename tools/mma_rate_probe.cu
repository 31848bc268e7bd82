// Issue-rate probe of tcgen05.mma kind::f16 (bf16 -> fp32) by operand layout (developer tool).
// Operands sit in shared memory (contents irrelevant); one thread issues `iters` blocks of 4 MMAs (one 64-element
// k-block) back to back and the kernel reports SM cycles per MMA.  Layouts: K-major and MN-major 128B-swizzled tiles
// exactly as the head kernels build them (tc05.cuh), single CTA (M = 128) and CTA pair (M = 256), N = 256.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I openset_imagenet_b200/csrc -o build/mma_rate_probe tools/mma_rate_probe.cu
#include "tc05.cuh"
#include <stdio.h>
#include <stdlib.h>

using namespace osb::tc;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s failed: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

constexpr int kBox = 8192;        // 64 rows x 128 B
constexpr int kStage = 16384 + 32768;
constexpr int kSmem = 4 * kStage + 1024 + 256;

template <bool PAIR, bool A_MN, bool B_MN>
__global__ void __launch_bounds__(64, 1) rate_kernel(int iters, int N, long long* out, int data, int mode) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 4 * kStage);
    uint64_t* bars2 = bar + 2;      // [8] per-stage barriers (mode bits 1, 2)
    uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 1);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 4 * kStage / 4; i += blockDim.x) {
        uint32_t v = 0;
        if (data) {     // two pseudo-random bf16 in (-2, 2): sign, exponent 126..127, random mantissa
            uint32_t h = (i + 1u) * 2654435761u + blockIdx.x * 40503u; h ^= h >> 15; h *= 2246822519u; h ^= h >> 13;
            const uint32_t lo = (h & 0x807Fu) | (0x3F00u + ((h >> 8) & 0x80u)), hi = ((h >> 16) & 0x807Fu) | (0x3F00u + ((h >> 24) & 0x80u));
            v = lo | (hi << 16);
            if (data == 2) {   // full-range values: bf16 of two uniform floats in (-1, 1), scaled by 1e-3 for odd words
                const float f0 = ((int)(h & 0xFFFF) - 32768) / 32768.f, f1 = ((int)(h >> 16) - 32768) / 32768.f * 1e-3f;
                v = (__float_as_uint(f0) >> 16) | (__float_as_uint(f1) & 0xFFFF0000u);
            }
        }
        reinterpret_cast<uint32_t*>(smem)[i] = v;
    }
    if (threadIdx.x == 0) { mbar_init(bar, 1); for (int i = 0; i < 8; ++i) mbar_init(&bars2[i], 1); fence_barrier_init(); }
    fence_proxy_async_smem();
    if (warp == 1) { if (PAIR) tmem_alloc2(slot, 512); else tmem_alloc(slot, 512); }
    tc_fence_before_sync();
    if (PAIR) cluster_sync_(); else __syncthreads();
    tc_fence_after_sync();
    const uint32_t tbase = *slot;
    const uint32_t rank = PAIR ? cluster_ctarank_() : 0;
    if (warp == 1 && lane == 0 && rank == 0) {
        const uint32_t idesc = make_idesc_bf16(PAIR ? 256 : 128, N, A_MN, B_MN);
        const uint32_t a_addr = smem_u32(smem), b_addr = a_addr + 16384;
        const long long t0 = clock64();
        for (int it = 0; it < iters; ++it) {
            // mode bit 2: wait on a barrier whose awaited phase completed long ago (the full-barrier wait of a fed pipeline)
            if (mode & 2) { mbar_wait(&bars2[4 + (it & 3)], 1); tc_fence_after_sync(); }
            const uint32_t so = (mode & 4) ? (it & 3) * kStage : 0;        // mode bit 4: cycle through 4 stage buffers
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const uint64_t ad = A_MN ? make_desc_mnmajor(a_addr + so + k * 2048, kBox) : make_desc_kmajor(a_addr + so + k * 32);
                const uint64_t bd = B_MN ? make_desc_mnmajor(b_addr + so + k * 2048, kBox) : make_desc_kmajor(b_addr + so + k * 32);
                // mode bit 8: "tiles" of 16 k-blocks accumulate into one buffer (first MMA overwrites), buffers alternate per tile
                const uint32_t tb = tbase + (((mode & 8) ? (it >> 4) : it) & 1) * 256;
                const bool acc = (mode & 8) ? ((it & 15) | k) != 0 : true;
                if (PAIR) umma2_bf16(tb, ad, bd, idesc, acc);
                else umma_bf16(tb, ad, bd, idesc, acc);
            }
            // mode bit 1: commit after every k-block (the empty-barrier arrive of a fed pipeline); nobody waits on these
            if (mode & 1) { if (PAIR) umma2_commit_mc(&bars2[it & 3]); else umma_commit(&bars2[it & 3]); }
            // mode bit 8: all MMAs of a 16-k-block "tile" accumulate into one buffer, switching every 16 (instead of every block)
        }
        if (PAIR) umma2_commit_mc(bar); else umma_commit(bar);
        mbar_wait(bar, 0);
        const long long t1 = clock64();
        out[blockIdx.x] = t1 - t0;
    }
    tc_fence_before_sync();
    if (PAIR) cluster_sync_(); else __syncthreads();
    if (warp == 1) { if (PAIR) tmem_dealloc2(tbase, 512); else tmem_dealloc(tbase, 512); }
}

template <bool PAIR, bool A_MN, bool B_MN>
static void run(const char* name, int grid, int N, int data = 0, int mode = 0) {
    const int iters = 2000;
    long long* out;
    CK(cudaMalloc(&out, 1024 * sizeof(long long)));
    CK(cudaMemset(out, 0, 1024 * sizeof(long long)));
    auto k = rate_kernel<PAIR, A_MN, B_MN>;
    CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(64); cfg.dynamicSmemBytes = kSmem;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = PAIR ? 2 : 1; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    for (int rep = 0; rep < 2; ++rep) {
        CK(cudaEventRecord(e0));
        CK(cudaLaunchKernelEx(&cfg, k, iters, N, out, data, mode));
        CK(cudaEventRecord(e1));
        CK(cudaDeviceSynchronize());
    }
    float ms;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    long long h[1024];
    CK(cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost));
    const double cyc = (double)h[0] / (iters * 4.0);
    const double flops = 2.0 * (PAIR ? 256 : 128) * N * 16 * 4.0 * iters * (PAIR ? grid / 2 : grid);
    printf("%-34s mode %d %s grid %3d N %3d: %7.1f cycles per MMA, %7.1f us, %7.1f TF/s, SM clock %.0f MHz\n", name, mode, data == 2 ? "wide  " : data ? "random" : "zeros ", grid, N, cyc, ms * 1e3,
           flops / (ms * 1e-3) / 1e12, (double)h[0] / (ms * 1e3));
    CK(cudaFree(out));
}

int main() {
    for (int grid : {2, 148}) {
        run<false, false, false>("1 CTA  M=128: A K-major, B K-major", grid, 256);
        run<false, false, true>("1 CTA  M=128: A K-major, B MN-major", grid, 256);
        run<false, true, true>("1 CTA  M=128: A MN-major, B MN-major", grid, 256);
        run<true, false, false>("pair   M=256: A K-major, B K-major", grid, 256);
        run<true, false, true>("pair   M=256: A K-major, B MN-major", grid, 256);
        run<true, true, true>("pair   M=256: A MN-major, B MN-major", grid, 256);
        run<true, true, false>("pair   M=256: A MN-major, B K-major", grid, 256);
    }
    for (int grid : {2, 148}) {
        run<false, false, false>("1 CTA  M=128: A K-major, B K-major", grid, 256, 1);
        run<true, false, false>("pair   M=256: A K-major, B K-major", grid, 256, 1);
        run<true, false, true>("pair   M=256: A K-major, B MN-major", grid, 256, 1);
        run<true, true, true>("pair   M=256: A MN-major, B MN-major", grid, 256, 1);
    }
    for (int mode : {15}) {
        run<true, false, false>("pair   M=256: A K-major, B K-major", 148, 256, 2, mode);
        run<true, false, true>("pair   M=256: A K-major, B MN-major", 148, 256, 2, mode);
        run<true, true, true>("pair   M=256: A MN-major, B MN-major", 148, 256, 2, mode);
        run<false, true, true>("1 CTA  M=128: A MN-major, B MN-major", 148, 256, 2, mode);
    }
    run<false, false, false>("1 CTA  M=128: A K-major, B K-major", 148, 256, 1, 7);
    run<false, false, false>("1 CTA  M=128 N=128 K/K", 148, 128);
    run<true, false, false>("pair   M=256 N=128 K/K", 148, 128);
    return 0;
}
