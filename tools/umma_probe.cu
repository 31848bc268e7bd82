// Hardware probe for the tcgen05 shared-memory descriptor conventions libosb relies on (tc05.cuh):
// one CTA computes D[128 x 256] = A[128 x 128] * B[256 x 128]^T with TMA-loaded, 128B-swizzled operands in
// every combination of K-major / MN-major A and B, and checks against a host reference.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o umma_probe tools/umma_probe.cu && ./umma_probe
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include "../openset_imagenet_b200/csrc/tc05.cuh"

using namespace osb;
using namespace osb::tc;

constexpr int M = 128, N = 256, KD = 128, KB = 64;       // two k-blocks of 64
constexpr int A_STAGE = M * KB * 2, B_STAGE = N * KB * 2;

struct Variant { int a_mn, b_mn; uint32_t lbo_a, sbo_a, lbo_b, sbo_b; int a_fp16; };

__device__ int g_timeout;

__device__ bool wait_bounded(uint64_t* bar, uint32_t parity) {
    for (long i = 0; i < 20000000L; ++i)
        if (mbar_try_wait(bar, parity)) return true;
    g_timeout = 1;
    return false;
}

__global__ void __launch_bounds__(128, 1)
probe_kernel(const __grid_constant__ CUtensorMap ta, const __grid_constant__ CUtensorMap tb, Variant v, float* out) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sa = smem;                       // 2 stages of A
    uint8_t* sb = smem + 2 * A_STAGE;         // 2 stages of B
    __shared__ uint64_t full[2], done;
    __shared__ uint32_t tmem_base;
    const int warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) {
        mbar_init(&full[0], 1); mbar_init(&full[1], 1); mbar_init(&done, 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(&tmem_base, 256);
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tbase = tmem_base;
    if (threadIdx.x == 0) {
        for (int s = 0; s < 2; ++s) {
            mbar_arrive_expect_tx(&full[s], A_STAGE + B_STAGE);
            if (!v.a_mn) tma_load_2d(sa + s * A_STAGE, &ta, &full[s], s * KB, 0);                // box 64(K) x 128 rows
            else for (int j = 0; j < M / 64; ++j) tma_load_2d(sa + s * A_STAGE + j * 8192, &ta, &full[s], j * 64, s * KB);
            if (!v.b_mn) tma_load_2d(sb + s * B_STAGE, &tb, &full[s], s * KB, 0);                // box 64(K) x 256 rows
            else for (int j = 0; j < N / 64; ++j) tma_load_2d(sb + s * B_STAGE + j * 8192, &tb, &full[s], j * 64, s * KB);
        }
        uint32_t idesc = make_idesc_bf16(M, N, v.a_mn, v.b_mn);
        if (v.a_fp16) idesc &= ~(7u << 7);          // A format field = F16 (0), B stays BF16
        bool ok = true;
        for (int s = 0; s < 2 && ok; ++s) {
            ok = wait_bounded(&full[s], 0);
            tc_fence_after_sync();
            for (int k = 0; k < KB / 16 && ok; ++k) {
                const uint32_t a_addr = smem_u32(sa + s * A_STAGE) + (v.a_mn ? k * 2048 : k * 32);
                const uint32_t b_addr = smem_u32(sb + s * B_STAGE) + (v.b_mn ? k * 2048 : k * 32);
                umma_bf16(tbase, make_desc_sw128(a_addr, v.lbo_a, v.sbo_a), make_desc_sw128(b_addr, v.lbo_b, v.sbo_b),
                          idesc, (s | k) != 0);
            }
        }
        umma_commit(&done);
    }
    __syncthreads();
    bool ok = wait_bounded(&done, 0);
    tc_fence_after_sync();
    if (ok) {
        const int row = warp * 32 + (threadIdx.x & 31);
        for (int c = 0; c < N; c += 32) {
            uint32_t r[32];
            tmem_ld32(tmem_addr(tbase, warp * 32, c), r);
            tmem_ld_wait();
            for (int i = 0; i < 32; ++i) out[row * N + c + i] = __uint_as_float(r[i]);
        }
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tbase, 256);
}

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorName(e), __FILE__, __LINE__); return 2; } } while (0)

int main() {
    std::vector<float> A(M * KD), B(N * KD), ref(M * N);
    srand(1);
    for (auto& x : A) x = (float)(rand() % 7 - 3);
    for (auto& x : B) x = (float)(rand() % 7 - 3);
    for (int m = 0; m < M; ++m)
        for (int n = 0; n < N; ++n) {
            float s = 0;
            for (int k = 0; k < KD; ++k) s += A[m * KD + k] * B[n * KD + k];
            ref[m * N + n] = s;
        }
    std::vector<__nv_bfloat16> Ak(M * KD), At(KD * M), Bk(N * KD), Bt(KD * N);
    std::vector<__half> Akh(M * KD), Ath(KD * M);
    for (int m = 0; m < M; ++m) for (int k = 0; k < KD; ++k) { Ak[m * KD + k] = __float2bfloat16(A[m * KD + k]); At[k * M + m] = Ak[m * KD + k]; Akh[m * KD + k] = __float2half(A[m * KD + k]); Ath[k * M + m] = Akh[m * KD + k]; }
    for (int n = 0; n < N; ++n) for (int k = 0; k < KD; ++k) { Bk[n * KD + k] = __float2bfloat16(B[n * KD + k]); Bt[k * N + n] = Bk[n * KD + k]; }
    __nv_bfloat16 *dAk, *dAt, *dBk, *dBt; float* dout; __half *dAkh, *dAth;
    CK(cudaMalloc(&dAkh, Akh.size() * 2)); CK(cudaMalloc(&dAth, Ath.size() * 2));
    CK(cudaMemcpy(dAkh, Akh.data(), Akh.size() * 2, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dAth, Ath.data(), Ath.size() * 2, cudaMemcpyHostToDevice));
    CK(cudaMalloc(&dAk, Ak.size() * 2)); CK(cudaMalloc(&dAt, At.size() * 2));
    CK(cudaMalloc(&dBk, Bk.size() * 2)); CK(cudaMalloc(&dBt, Bt.size() * 2));
    CK(cudaMalloc(&dout, M * N * 4));
    CK(cudaMemcpy(dAk, Ak.data(), Ak.size() * 2, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dAt, At.data(), At.size() * 2, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dBk, Bk.data(), Bk.size() * 2, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dBt, Bt.data(), Bt.size() * 2, cudaMemcpyHostToDevice));
    CUtensorMap tAk, tAt, tBk, tBt, tAkh, tAth;
    if (make_tmap_bf16(&tAkh, dAkh, M, KD, KD, 128) || make_tmap_bf16(&tAth, dAth, KD, M, M, 64)) { printf("tensor map encode failed\n"); return 3; }
    if (make_tmap_bf16(&tAk, dAk, M, KD, KD, 128) || make_tmap_bf16(&tAt, dAt, KD, M, M, 64) ||
        make_tmap_bf16(&tBk, dBk, N, KD, KD, 256) || make_tmap_bf16(&tBt, dBt, KD, N, N, 64)) {
        printf("tensor map encode failed\n");
        return 3;
    }
    const int smem_bytes = 2 * A_STAGE + 2 * B_STAGE + 1024;
    CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
    struct Named { const char* name; Variant v; };
    Named variants[] = {
        {"A=K  B=K   (lbo 0, sbo 1024)", {0, 0, 0, 1024, 0, 1024, 0}},
        {"A=MN B=K   (lbo 8192, sbo 1024)", {1, 0, 8192, 1024, 0, 1024, 0}},
        {"A=K  B=MN  (lbo 8192, sbo 1024)", {0, 1, 0, 1024, 8192, 1024, 0}},
        {"A=MN B=MN  (lbo 8192, sbo 1024)", {1, 1, 8192, 1024, 8192, 1024, 0}},
        {"A=MN B=K   SWAPPED (lbo 1024, sbo 8192)", {1, 0, 1024, 8192, 0, 1024, 0}},
        {"A=K  B=MN  SWAPPED (lbo 1024, sbo 8192)", {0, 1, 0, 1024, 1024, 8192, 0}},
        {"A=K  B=K   (lbo 16, sbo 1024)", {0, 0, 16, 1024, 16, 1024, 0}},
        {"A=K(fp16)  B=K(bf16)  mixed formats", {0, 0, 0, 1024, 0, 1024, 1}},
        {"A=MN(fp16) B=MN(bf16) mixed formats", {1, 1, 8192, 1024, 8192, 1024, 1}},
    };
    std::vector<float> got(M * N);
    int fails = 0;
    for (auto& nv : variants) {
        CK(cudaMemset(dout, 0xff, M * N * 4));
        int zero = 0;
        CK(cudaMemcpyToSymbol(g_timeout, &zero, sizeof(int)));
        probe_kernel<<<1, 128, smem_bytes>>>(nv.v.a_fp16 ? (nv.v.a_mn ? tAth : tAkh) : (nv.v.a_mn ? tAt : tAk), nv.v.b_mn ? tBt : tBk, nv.v, dout);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("%-44s CUDA ERROR %s\n", nv.name, cudaGetErrorName(e)); return 4; }
        int to = 0;
        CK(cudaMemcpyFromSymbol(&to, g_timeout, sizeof(int)));
        CK(cudaMemcpy(got.data(), dout, M * N * 4, cudaMemcpyDeviceToHost));
        double maxerr = 0; int bad = 0;
        for (int i = 0; i < M * N; ++i) {
            double d = fabs((double)got[i] - ref[i]);
            if (!(d <= 1e-3)) { ++bad; }
            if (d > maxerr || d != d) maxerr = d;
        }
        printf("%-44s timeout=%d mismatches=%d/%d maxerr=%g  %s\n", nv.name, to, bad, M * N, maxerr, bad == 0 && !to ? "PASS" : "FAIL");
        if (bad || to) ++fails;
    }
    printf("probe done, %d variant(s) failed\n", fails);
    return 0;
}
