// Probe of the 2-CTA tensor-core path (tcgen05 cta_group::2) on B200 (developer tool).
//   D[M,N] (fp32) = A[M,K] (bf16, K-major) x B[N,K]^T (bf16, K-major)
// One cluster of 2 CTAs per 256 x 256 output tile: CTA r holds rows [128 r, 128 r + 128) of the A tile and rows
// [128 r, 128 r + 128) of the B tile in ITS shared memory; the leader (rank 0) issues M = 256 MMAs that read both CTAs'
// shared memory; each CTA ends up with its 128 rows x 256 columns in its own TMEM.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I openset_imagenet_b200/csrc -o build/umma2_probe tools/umma2_probe.cu
#include "tc05.cuh"
#include <stdio.h>
#include <stdlib.h>
#include <math.h>
#include <vector>
#include <cuda_bf16.h>

using namespace osb::tc;
using osb::make_tmap_bf16;

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s failed: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

constexpr int BM = 128, BN_HALF = 128, BK = 64, STAGES = 4;
constexpr int kABytes = BM * BK * 2, kBBytes = BN_HALF * BK * 2, kStageBytes = kABytes + kBBytes;      // 32 KB per CTA per stage
constexpr uint32_t kPeerMask = 0xFEFFFFFFu;

__device__ __forceinline__ uint32_t cluster_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_alloc2(uint32_t* smem_dst, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(smem_dst)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" :: "r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma2_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, bool acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 :: "r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"((uint32_t)acc) : "memory");
}
// arrive on the barrier at this shared-memory offset in BOTH CTAs of the pair when all MMAs issued so far are done
__device__ __forceinline__ void umma2_commit_mc(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 :: "r"(smem_u32(bar)), "h"((uint16_t)3) : "memory");
}
// 2-D tile load into THIS CTA's shared memory; the bytes are counted on the LEADER CTA's barrier
__device__ __forceinline__ void tma_load_2d_2sm(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 :: "r"(smem_u32(smem_dst)), "l"(m), "r"(smem_u32(bar) & kPeerMask), "r"(c0), "r"(c1) : "memory");
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(192, 1)
gemm2_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b, float* __restrict__ D, int M, int N, int K) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * kStageBytes);
    uint64_t* full = bars; uint64_t* empty = full + STAGES; uint64_t* t_full = empty + STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(t_full + 1);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_rank();
    const int pair = blockIdx.x / 2;
    const int n_tiles_n = N / 256;
    const int tm = pair / n_tiles_n, tn = pair % n_tiles_n;
    const int n_kb = K / BK;
    if (threadIdx.x == 0) {
        for (int i = 0; i < STAGES; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
        mbar_init(t_full, 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc2(tmem_slot, 256);
    tc_fence_before_sync();
    cluster_sync();
    tc_fence_after_sync();
    const uint32_t tbase = *tmem_slot;

    if (warp == 0 && lane == 0) {
        for (int kb = 0; kb < n_kb; ++kb) {
            const int s = kb % STAGES; const uint32_t ph = (kb / STAGES) & 1;
            mbar_wait(&empty[s], ph ^ 1);
            if (rank == 0) mbar_arrive_expect_tx(&full[s], 2 * kStageBytes);       // both CTAs' bytes land on the leader's barrier
            uint8_t* st = smem + s * kStageBytes;
            tma_load_2d_2sm(st, &tm_a, &full[s], kb * BK, tm * 256 + (int)rank * 128);
            tma_load_2d_2sm(st + kABytes, &tm_b, &full[s], kb * BK, tn * 256 + (int)rank * 128);
        }
    } else if (warp == 1 && lane == 0 && rank == 0) {
        const uint32_t idesc = make_idesc_bf16(256, 256, false, false);
        for (int kb = 0; kb < n_kb; ++kb) {
            const int s = kb % STAGES; const uint32_t ph = (kb / STAGES) & 1;
            mbar_wait(&full[s], ph);
            tc_fence_after_sync();
            const uint32_t a_addr = smem_u32(smem + s * kStageBytes);
            const uint32_t b_addr = a_addr + kABytes;
#pragma unroll
            for (int k = 0; k < BK / 16; ++k)
                umma2_bf16(tbase, make_desc_kmajor(a_addr + k * 32), make_desc_kmajor(b_addr + k * 32), idesc, (kb | k) != 0);
            umma2_commit_mc(&empty[s]);
        }
        umma2_commit_mc(t_full);
    } else if (warp >= 2) {
        const int q = warp & 3;
        mbar_wait(t_full, 0);
        tc_fence_after_sync();
        const int row = tm * 256 + (int)rank * 128 + q * 32 + lane;
        const uint32_t t_row = tmem_addr(tbase, q * 32, 0);
        for (int c0 = 0; c0 < 256; c0 += 32) {
            uint32_t v[32];
            tmem_ld32(t_row + c0, v);
            tmem_ld_wait();
            if (row < M)
                for (int i = 0; i < 32; ++i) D[(size_t)row * N + tn * 256 + c0 + i] = __uint_as_float(v[i]);
        }
    }
    tc_fence_before_sync();
    cluster_sync();
    if (warp == 1) tmem_dealloc2(tbase, 256);
}

static float bf(float x) { return __bfloat162float(__float2bfloat16(x)); }

int main(int argc, char** argv) {
    const int M = argc > 1 ? atoi(argv[1]) : 512, N = argc > 2 ? atoi(argv[2]) : 512, K = argc > 3 ? atoi(argv[3]) : 256;
    printf("umma2 probe M=%d N=%d K=%d\n", M, N, K);
    std::vector<__nv_bfloat16> hA((size_t)M * K), hB((size_t)N * K);
    std::vector<float> fA((size_t)M * K), fB((size_t)N * K);
    srand(1);
    for (size_t i = 0; i < hA.size(); ++i) { float v = (rand() % 2001 - 1000) / 1000.f; hA[i] = __float2bfloat16(v); fA[i] = bf(v); }
    for (size_t i = 0; i < hB.size(); ++i) { float v = (rand() % 2001 - 1000) / 1000.f; hB[i] = __float2bfloat16(v); fB[i] = bf(v); }
    __nv_bfloat16 *dA, *dB; float* dD;
    CK(cudaMalloc(&dA, hA.size() * 2)); CK(cudaMalloc(&dB, hB.size() * 2)); CK(cudaMalloc(&dD, (size_t)M * N * 4));
    CK(cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice));
    CK(cudaMemset(dD, 0xff, (size_t)M * N * 4));
    CUtensorMap ta, tb;
    if (make_tmap_bf16(&ta, dA, M, K, K, 128) || make_tmap_bf16(&tb, dB, N, K, K, 128)) { printf("tensor map failed\n"); return 1; }
    const size_t smem = STAGES * kStageBytes + 256 + 1024;
    CK(cudaFuncSetAttribute(gemm2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int pairs = (M / 256) * (N / 256);
    gemm2_kernel<<<pairs * 2, 192, smem>>>(ta, tb, dD, M, N, K);
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
    std::vector<float> hD((size_t)M * N);
    CK(cudaMemcpy(hD.data(), dD, hD.size() * 4, cudaMemcpyDeviceToHost));
    double max_err = 0; int bad = 0;
    const int step_r = M > 1024 ? 37 : 1, step_c = N > 1024 ? 29 : 1;
    for (int r = 0; r < M; r += step_r)
        for (int c = 0; c < N; c += step_c) {
            double acc = 0;
            for (int k = 0; k < K; ++k) acc += (double)fA[(size_t)r * K + k] * fB[(size_t)c * K + k];
            const double e = fabs(acc - hD[(size_t)r * N + c]);
            if (!(e <= 1e-2 * (1 + fabs(acc)))) { if (bad < 5) printf("  mismatch at (%d,%d): got %f want %f\n", r, c, hD[(size_t)r * N + c], acc); ++bad; }
            if (e > max_err) max_err = e;
        }
    printf("max abs err %.3e, mismatches %d -> %s\n", max_err, bad, bad ? "FAIL" : "OK");
    if (!bad) {
        cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
        float best = 1e30f;
        for (int i = 0; i < 10; ++i) {
            CK(cudaEventRecord(a)); gemm2_kernel<<<pairs * 2, 192, smem>>>(ta, tb, dD, M, N, K); CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b));
            float ms; CK(cudaEventElapsedTime(&ms, a, b)); if (ms < best) best = ms;
        }
        printf("time %.1f us, %.1f TFLOP/s (non-persistent, no epilogue overlap)\n", best * 1e3, 2.0 * M * N * K / (best * 1e-3) / 1e12);
    }
    return bad ? 2 : 0;
}
