// Streaming-read probe for B200 (developer tool): how fast can one pass over a large float array go with
//   (a) cp.async.bulk (1-D TMA) into shared memory, S stages of B bytes per CTA, P CTAs per SM, K sub-copies per stage
//   (b) plain 128-bit loads, U independent loads in flight per thread
// Every variant folds the data into a max so that nothing is optimised away.  Prints GB/s per variant.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o build/stream_probe tools/stream_probe.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s failed: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// (a) bulk-copy pipeline.  Each CTA takes chunks blockIdx.x, blockIdx.x + grid, ...
__global__ void bulk_kernel(const float* __restrict__ src, size_t n_chunks, uint32_t B, int S, int K, int touch, float* out, uint32_t LB = 0) {
    extern __shared__ __align__(128) uint8_t smem[];
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + (size_t)S * (B + LB));
    const int tid = threadIdx.x;
    if (tid == 0) {
        for (int s = 0; s < S; ++s) mbar_init(&full[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const uint32_t sub = B / K;
    auto issue = [&](size_t chunk, int s) {        // threads 0..K-1 each copy one slice; thread 0 arms the barrier
        if (chunk >= n_chunks) return;
        if (tid == 0) mbar_expect(&full[s], B + LB);
        if (tid == 0 && LB) bulk_g2s(smem + (size_t)S * B + (size_t)s * LB, reinterpret_cast<const uint8_t*>(src) + chunk * LB, LB, &full[s]);
        __syncwarp();
        if (tid < K) bulk_g2s(smem + (size_t)s * B + (size_t)tid * sub, reinterpret_cast<const uint8_t*>(src) + chunk * B + (size_t)tid * sub, sub, &full[s]);
    };
    if (tid < 32) for (int k = 0; k < S; ++k) issue(blockIdx.x + (size_t)k * gridDim.x, k);
    float m = -1e30f;
    int s = 0; uint32_t ph = 0;
    for (size_t chunk = blockIdx.x; chunk < n_chunks; chunk += gridDim.x) {
        while (!mbar_try(&full[s], ph)) { }
        const float4* p = reinterpret_cast<const float4*>(smem + (size_t)s * B);
        const int n4 = B / 16;
        if (touch) { for (int i = tid; i < n4; i += blockDim.x) { float4 v = p[i]; m = fmaxf(m, fmaxf(fmaxf(v.x, v.y), fmaxf(v.z, v.w))); } }
        else { float4 v = p[tid % n4]; m = fmaxf(m, v.x); }
        __syncthreads();
        if (tid < 32) issue(chunk + (size_t)S * gridDim.x, s);
        if (++s == S) { s = 0; ph ^= 1; }
    }
    if (m == 12345.f) out[0] = m;
}

// (b) direct loads, U float4 in flight per thread
template <int U>
__global__ void ldg_kernel(const float4* __restrict__ src, size_t n4, float* out) {
    float m = -1e30f;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + (U - 1) * stride < n4; i += U * stride) {
        float4 v[U];
#pragma unroll
        for (int u = 0; u < U; ++u)
            asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];"
                         : "=f"(v[u].x), "=f"(v[u].y), "=f"(v[u].z), "=f"(v[u].w) : "l"(src + i + u * stride));
#pragma unroll
        for (int u = 0; u < U; ++u) m = fmaxf(m, fmaxf(fmaxf(v[u].x, v[u].y), fmaxf(v[u].z, v[u].w)));
    }
    for (; i < n4; i += stride) { float4 v = src[i]; m = fmaxf(m, v.x); }
    if (m == 12345.f) out[0] = m;
}

template <typename F>
static float time_us(F launch, int iters = 10) {
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    launch(); launch();
    CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int i = 0; i < iters; ++i) {
        CK(cudaEventRecord(a)); launch(); CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b));
        float ms; CK(cudaEventElapsedTime(&ms, a, b));
        if (ms < best) best = ms;
    }
    return best * 1e3f;
}

int main() {
    const size_t bytes = (size_t)464 << 20;         // ~486 MB, like 1M x 116 floats
    float* src; float* out;
    CK(cudaMalloc(&src, bytes)); CK(cudaMalloc(&out, 256));
    CK(cudaMemset(src, 0, bytes));
    int sms = 0; CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    printf("SMs %d, bytes %.1f MB\n", sms, bytes / 1e6);
    CK(cudaFuncSetAttribute(bulk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    struct V { uint32_t B; int S; int K; int P; int threads; int touch; };
    const V vs[] = {
        {29696, 6, 1, 1, 256, 0}, {29696, 6, 1, 1, 256, 1}, {29696, 7, 1, 1, 512, 1}, {59392, 3, 1, 1, 512, 1},
        {29696, 3, 1, 2, 256, 1}, {14848, 6, 1, 2, 256, 1}, {14848, 4, 1, 3, 256, 1}, {7424, 12, 1, 2, 256, 1},
        {29696, 6, 4, 1, 256, 1}, {29696, 6, 16, 1, 256, 1}, {59392, 3, 16, 1, 512, 1}, {29696, 3, 8, 2, 256, 1},
        {16384, 13, 1, 1, 256, 1}, {32768, 6, 1, 1, 256, 1}, {65536, 3, 1, 1, 512, 1}, {32768, 6, 8, 1, 256, 1},
    };
    for (const V& v : vs) {
        const size_t n_chunks = bytes / v.B;
        const size_t smem = (size_t)v.S * v.B + 128;
        const int grid = sms * v.P;
        float us = time_us([&] { bulk_kernel<<<grid, v.threads, smem>>>(src, n_chunks, v.B, v.S, v.K, v.touch, out); });
        CK(cudaGetLastError());
        printf("bulk  B=%6u S=%2d K=%2d ctas/SM=%d thr=%4d touch=%d : %7.1f us  %7.1f GB/s\n", v.B, v.S, v.K, v.P, v.threads, v.touch,
               us, (double)n_chunks * v.B / us / 1e3);
    }
    for (uint32_t LB : {0u, 512u, 2048u}) {
        const uint32_t B = 29696; const int S = 3, P = 2;
        const size_t n_chunks = bytes / B;
        float us = time_us([&] { bulk_kernel<<<sms * P, 256, (size_t)S * (B + LB) + 128>>>(src, n_chunks, B, S, 1, 1, out, LB); });
        printf("bulk  B=%6u S=%2d ctas/SM=%d + label copy %4u B : %7.1f us  %7.1f GB/s\n", B, S, P, LB, us, (double)n_chunks * B / us / 1e3);
    }
    const size_t n4 = bytes / 16;
    for (int occ : {2, 4, 8}) {
        float us;
        us = time_us([&] { ldg_kernel<4><<<sms * occ, 256>>>(reinterpret_cast<const float4*>(src), n4, out); });
        printf("ldg   U=4  ctas/SM=%d thr=256 : %7.1f us  %7.1f GB/s\n", occ, us, (double)bytes / us / 1e3);
        us = time_us([&] { ldg_kernel<8><<<sms * occ, 256>>>(reinterpret_cast<const float4*>(src), n4, out); });
        printf("ldg   U=8  ctas/SM=%d thr=256 : %7.1f us  %7.1f GB/s\n", occ, us, (double)bytes / us / 1e3);
        us = time_us([&] { ldg_kernel<16><<<sms * occ, 256>>>(reinterpret_cast<const float4*>(src), n4, out); });
        printf("ldg   U=16 ctas/SM=%d thr=256 : %7.1f us  %7.1f GB/s\n", occ, us, (double)bytes / us / 1e3);
    }
    CK(cudaDeviceSynchronize());
    return 0;
}
