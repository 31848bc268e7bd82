// All-reduce of the head's weight gradient over NVLink peer memory, fused with the split-K reduction.
//
// The dW GEMM (head_dw_kernel) leaves `splits` fp32 partial matrices in every rank's workspace.  Instead of
// reducing them locally and handing the result to NCCL, ONE kernel per rank does a two-shot all-reduce straight
// from the partials: rank r owns the r-th slice of the C*K elements, reads that slice of EVERY rank's partials
// with peer loads (fixed order: rank 0 split 0, rank 0 split 1, ..., so the sum is deterministic and all ranks end
// up with bit-identical gradients), scales by g and writes the result into every rank's dW with peer stores.
// A small `extra` vector per rank (db, loss numerator) is summed by every rank for itself.  Cross-GPU ordering uses
// epoch-stamped flags in peer memory (st.release.sys / ld.acquire.sys); epochs only grow, nothing is reset.  The epoch
// itself lives in the flag area and is advanced by the kernel, so a launch has no per-step argument and can be replayed
// from a CUDA graph.
//
// All buffers must be peer-mapped (torch.distributed._symmetric_memory).  The kernel runs on a side stream with a
// small grid so that it overlaps the dX GEMM.  The reference has no counterpart.
#include "osb_common.cuh"

namespace osb {

constexpr int kArMaxWorld = 16;
constexpr int kArThreads = 1024;

struct ArParams {
    const float* partials[kArMaxWorld];     // [splits][n] per rank
    float* out[kArMaxWorld];                // [n] per rank (dW)
    const float* extra_in[kArMaxWorld];     // [n_extra] per rank, nullable
    unsigned int* flags[kArMaxWorld];       // per rank: [0..W) ready stamps, [W..2W) done stamps, [2W] error word, [2W+1] exit
                                            // ticket, [2W+2] epochs completed so far (local)
    float* extra_out;                       // local [n_extra]
    const float* mc_partials;               // nullable: MULTICAST address of the partial buffers (NVSwitch, multimem.ld_reduce)
    float* mc_out;                          // nullable: multicast address of the dW buffers (multimem.st)
    const float* gscale;                    // nullable device scalar
    const double* den;                      // global loss denominator (device)
    long long n;                            // elements of dW (multiple of 4)
    long long split_stride;                 // elements between consecutive splits
    int n_extra, n_plain, splits, rank, world;
    long long timeout_cycles;
};

__device__ __forceinline__ void st_release_sys(unsigned int* p, unsigned int v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" :: "l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned int ld_acquire_sys(const unsigned int* p) {
    unsigned int v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ float4 ld_peer_f4(const float4* p) {
    float4 r;       // peer memory is not cached in the local L2; keep it out of L1 as well
    asm volatile("ld.global.relaxed.sys.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p) : "memory");
    return r;
}

// Wait until flags[i] >= epoch for every i < world (threads 0..world-1 poll one flag each).  Gives up after
// timeout_cycles and records the failure instead of hanging the GPU; returns false (block-uniform) in that case.
__device__ __forceinline__ bool wait_stamps(const unsigned int* stamps, unsigned int* err, unsigned int epoch, const ArParams& p) {
    int ok = 1;
    if (threadIdx.x < p.world) {
        const long long t0 = clock64();
        while ((int)(ld_acquire_sys(stamps + threadIdx.x) - epoch) < 0) {
            if (clock64() - t0 > p.timeout_cycles) { atomicExch(err, epoch); ok = 0; break; }
        }
    }
    return __syncthreads_and(ok) != 0;
}

__global__ void __launch_bounds__(kArThreads)
dw_allreduce_kernel(const ArParams p) {
    unsigned int* my_flags = p.flags[p.rank];
    unsigned int* err = my_flags + 2 * p.world;
    __shared__ unsigned int s_last;
    // the epoch of this launch: one more than the launches completed so far (the word is advanced by the last block out,
    // after every block has read it, and the next launch on this stream starts after this one has ended)
    const unsigned int epoch = *reinterpret_cast<volatile unsigned int*>(my_flags + 2 * p.world + 2) + 1u;
    // 1. my partials are complete (stream order: this kernel follows the dW GEMM): tell every rank
    if (blockIdx.x == 0 && threadIdx.x < p.world) st_release_sys(p.flags[threadIdx.x] + p.rank, epoch);
    // 2. wait for every rank's partials.  A rank that does not show up poisons the result instead of corrupting it:
    //    nothing is reduced or stored, the plain extras (loss) become NaN and the error word keeps the epoch.
    const bool alive = wait_stamps(my_flags, err, epoch, p);
    // 3. reduce my slice and publish it everywhere
    const float gs = (float)((p.gscale ? (double)p.gscale[0] : 1.0) / p.den[0]);
    const long long n4 = p.n / 4;
    const long long per = (n4 + p.world - 1) / p.world;
    const long long lo = per * p.rank, hi = (lo + per < n4) ? lo + per : n4;
    const long long stride4 = p.split_stride / 4;
    // All sources (rank 0 split 0, rank 0 split 1, ..., rank 1 split 0, ...) are listed once in shared memory; a thread
    // takes kElems elements and kBatch sources at a time: 16 independent 16-byte peer loads in flight per thread (a
    // peer load takes ~2 us), while the sum itself stays in that fixed source order.
    constexpr int kElems = 2, kBatch = 8;
    __shared__ const float4* s_src[kArMaxWorld * 16];
    const int n_src = p.world * p.splits;
    for (int q = threadIdx.x; q < n_src; q += kArThreads)
        s_src[q] = reinterpret_cast<const float4*>(p.partials[q / p.splits]) + (long long)(q % p.splits) * stride4;
    __syncthreads();
    const long long tstride = (long long)gridDim.x * kArThreads;
    if (p.mc_partials != nullptr) {
        // NVSwitch path: ONE multimem.ld_reduce per element and split returns the sum over all ranks, added inside the switch
        // (1 / world of the peer-load traffic and round trips); ONE multimem.st writes the result into every rank's dW.
        // Each element is still reduced once, by its owner, so all ranks hold bit-identical gradients.
        // The switch answers a multimem.ld_reduce in a few microseconds, so throughput = bytes in flight / that latency:
        // every thread keeps kMc elements of (up to) TWO splits in flight at once -- 8 x 16 bytes, 128 KB per CTA.
        constexpr int kMc = 4;
        const float4* mc_in = reinterpret_cast<const float4*>(p.mc_partials);
        float4* mc_o = reinterpret_cast<float4*>(p.mc_out);
        auto ldr = [](const float4* a) {
            float4 v;
            asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
                         : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(a) : "memory");
            return v;
        };
        for (long long i0 = lo + (long long)blockIdx.x * kArThreads + threadIdx.x; alive && i0 < hi; i0 += tstride * kMc) {
            float4 v0[kMc], v1[kMc];
#pragma unroll
            for (int e = 0; e < kMc; ++e) {
                const long long i = i0 + e * tstride;
                v0[e] = i < hi ? ldr(mc_in + i) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
            for (int e = 0; e < kMc; ++e) {
                const long long i = i0 + e * tstride;
                v1[e] = (p.splits > 1 && i < hi) ? ldr(mc_in + stride4 + i) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
            for (int e = 0; e < kMc; ++e) { v0[e].x += v1[e].x; v0[e].y += v1[e].y; v0[e].z += v1[e].z; v0[e].w += v1[e].w; }
            for (int sp = 2; sp < p.splits; ++sp) {             // more than two splits (not used by the Python layer)
#pragma unroll
                for (int e = 0; e < kMc; ++e) {
                    const long long i = i0 + e * tstride;
                    if (i < hi) { const float4 t = ldr(mc_in + (long long)sp * stride4 + i); v0[e].x += t.x; v0[e].y += t.y; v0[e].z += t.z; v0[e].w += t.w; }
                }
            }
#pragma unroll
            for (int e = 0; e < kMc; ++e) {
                const long long i = i0 + e * tstride;
                if (i < hi)
                    asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};"
                                 :: "l"(mc_o + i), "f"(v0[e].x * gs), "f"(v0[e].y * gs), "f"(v0[e].z * gs), "f"(v0[e].w * gs) : "memory");
            }
        }
    } else
    for (long long i0 = lo + (long long)blockIdx.x * kArThreads + threadIdx.x; alive && i0 < hi; i0 += tstride * kElems) {
        float4 acc[kElems];
#pragma unroll
        for (int e = 0; e < kElems; ++e) acc[e] = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int q0 = 0; q0 < n_src; q0 += kBatch) {
            float4 v[kElems][kBatch];
#pragma unroll
            for (int j = 0; j < kBatch; ++j) {
                const float4* src = s_src[q0 + j < n_src ? q0 + j : 0];
#pragma unroll
                for (int e = 0; e < kElems; ++e) {
                    const long long i = i0 + e * tstride;
                    v[e][j] = (i < hi && q0 + j < n_src) ? src[i] : make_float4(0.f, 0.f, 0.f, 0.f);
                }
            }
#pragma unroll
            for (int e = 0; e < kElems; ++e) {
#pragma unroll
                for (int j = 0; j < kBatch; ++j) { acc[e].x += v[e][j].x; acc[e].y += v[e][j].y; acc[e].z += v[e][j].z; acc[e].w += v[e][j].w; }
            }
        }
#pragma unroll
        for (int e = 0; e < kElems; ++e) {
            const long long i = i0 + e * tstride;
            if (i < hi) {
                acc[e].x *= gs; acc[e].y *= gs; acc[e].z *= gs; acc[e].w *= gs;
                for (int r = 0; r < p.world; ++r) reinterpret_cast<float4*>(p.out[r])[i] = acc[e];
            }
        }
    }
    if (blockIdx.x == 0 && p.n_extra > 0) {
        for (int i = threadIdx.x; i < p.n_extra; i += kArThreads) {
            float acc = 0.f;
            for (int r = 0; r < p.world; ++r) {
                float v;
                asm volatile("ld.global.relaxed.sys.f32 %0, [%1];" : "=f"(v) : "l"(p.extra_in[r] + i) : "memory");
                acc += v;
            }
            const bool plain = i >= p.n_extra - p.n_plain;            // loss numerators / counts: plain sums
            p.extra_out[i] = alive ? (plain ? acc : acc * gs) : __int_as_float(0x7fc00000);
        }
    }
    // 4. all my blocks are done: tell every rank (the last block out does it)
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned int* ticket = my_flags + 2 * p.world + 1;
        const unsigned int t = atomicAdd(ticket, 1u);
        s_last = (t == gridDim.x - 1) ? 1u : 0u;
        if (s_last) *ticket = 0u;
    }
    __syncthreads();
    if (s_last) {
        __threadfence_system();
        if (threadIdx.x < p.world) st_release_sys(p.flags[threadIdx.x] + p.world + p.rank, epoch);
        // 5. my dW is complete once every owner has published its slice; the extras of the other ranks must also
        //    stay untouched until everybody has read them, which the same stamps guarantee
        const bool done = wait_stamps(my_flags + p.world, err, epoch, p);
        if (!done && threadIdx.x == 0 && p.n_extra > 0)
            for (int i = p.n_extra - p.n_plain; i < p.n_extra; ++i) p.extra_out[i] = __int_as_float(0x7fc00000);
        if (threadIdx.x == 0) my_flags[2 * p.world + 2] = epoch;
    }
}

// ---- sum of a few doubles over the ranks (loss denominator of a batch-sharded softmax / garbage step) ----------------
// slot area per rank: [2 (epoch parity)][world][kPsMax][2] (32 bits, epoch) words, then the local epoch counter and the
// error word.  Two parities: a rank can run at most one call ahead of a peer (it needs that peer's values to finish), so
// its stores for call e + 1 must not land on the words the peer still polls for call e.
constexpr int kPsMax = 8;
struct PsParams {
    uint2* slots[kArMaxWorld];
    const double* in; double* out;
    int n, rank, world;
    long long timeout_cycles;
};
__device__ __forceinline__ void st_ll_sys(uint2* p, unsigned int v, unsigned int epoch) {
    asm volatile("st.relaxed.sys.global.v2.b32 [%0], {%1, %2};" :: "l"(p), "r"(v), "r"(epoch) : "memory");
}
__device__ __forceinline__ uint2 ld_ll_sys(const uint2* p) {
    uint2 v;
    asm volatile("ld.relaxed.sys.global.v2.b32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p) : "memory");
    return v;
}
__global__ void __launch_bounds__(kArMaxWorld * kPsMax)
peer_sum_kernel(const PsParams p) {
    // dependents launched programmatically behind this kernel (the split-K forward) may start now
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    uint2* mine = p.slots[p.rank];
    unsigned int* ctr = reinterpret_cast<unsigned int*>(mine + (size_t)2 * kArMaxWorld * kPsMax * 2);
    const unsigned int epoch = *reinterpret_cast<volatile unsigned int*>(ctr) + 1u;
    const size_t par = (size_t)(epoch & 1u) * kArMaxWorld * kPsMax * 2;
    const int r = threadIdx.x / kPsMax, j = threadIdx.x % kPsMax;
    __shared__ double s_val[kArMaxWorld][kPsMax];
    __shared__ int s_bad;
    if (threadIdx.x == 0) s_bad = 0;
    __syncthreads();
    // thread (r, j): store my value j into rank r's area (8-byte words are single-copy atomic: a reader that sees the epoch
    // has the 32 bits stored with it), then fetch rank r's value j from my own area
    if (r < p.world && j < p.n) {
        const unsigned long long bits = (unsigned long long)__double_as_longlong(p.in[j]);
        uint2* dst = p.slots[r] + par + ((size_t)p.rank * kPsMax + j) * 2;
        st_ll_sys(dst + 0, (unsigned int)bits, epoch);
        st_ll_sys(dst + 1, (unsigned int)(bits >> 32), epoch);
        const uint2* src = mine + par + ((size_t)r * kPsMax + j) * 2;
        const long long t0 = clock64();
        uint2 a, b;
        bool ok = false;
        do {
            a = ld_ll_sys(src + 0); b = ld_ll_sys(src + 1);
            ok = (a.y == epoch) && (b.y == epoch);
        } while (!ok && clock64() - t0 <= p.timeout_cycles);
        if (!ok) { s_bad = 1; ctr[1] = epoch; }
        s_val[r][j] = __longlong_as_double((long long)(((unsigned long long)b.x << 32) | a.x));
    }
    __syncthreads();
    if (threadIdx.x < p.n) {
        double acc = 0.0;
        for (int q = 0; q < p.world; ++q) acc += s_val[q][threadIdx.x];        // rank order: bit-identical on every rank
        p.out[threadIdx.x] = s_bad ? __longlong_as_double(0x7ff8000000000000LL) : acc;
    }
    if (threadIdx.x == 0) *ctr = epoch;
}

}  // namespace osb

using namespace osb;

static long long g_ar_timeout = 0;
// developer hook (osb_debug.h): shorter give-up time for the tests, 0 = default (~2 s)
extern "C" void osb_debug_set_allreduce_timeout(long long cycles) { g_ar_timeout = cycles; }

extern "C" size_t osb_head_dw_allreduce_flag_bytes(int world) { return world < 1 ? 0 : sizeof(unsigned int) * (2 * (size_t)world + 8); }

extern "C" int osb_head_dw_allreduce(const void* const* peer_partials, void* const* peer_dw, const void* const* peer_extra,
                                     void* const* peer_flags, int rank, int world, int splits, int64_t n, int64_t split_stride,
                                     float* extra_out, int n_extra, int n_plain, const float* gscale, const double* den,
                                     const void* mc_partials, void* mc_dw, int max_ctas, void* stream) {
    OSB_ENTRY_GUARD();
    OSB_REQUIRE((mc_partials == nullptr) == (mc_dw == nullptr) && (uintptr_t)mc_partials % 16 == 0 && (uintptr_t)mc_dw % 16 == 0, OSB_ERR_ARG,
                "osb_head_dw_allreduce: the multicast addresses come as a 16-byte aligned pair, or not at all");
    OSB_REQUIRE(peer_partials && peer_dw && peer_flags && den, OSB_ERR_ARG, "osb_head_dw_allreduce: null pointer");
    OSB_REQUIRE(world >= 1 && world <= kArMaxWorld && rank >= 0 && rank < world, OSB_ERR_ARG, "osb_head_dw_allreduce: bad rank %d / world %d", rank, world);
    OSB_REQUIRE(splits <= 16, OSB_ERR_ARG, "osb_head_dw_allreduce: at most 16 splits (got %d)", splits);
    OSB_REQUIRE(splits >= 1 && n > 0 && n % 4 == 0 && split_stride >= n && split_stride % 4 == 0, OSB_ERR_ARG,
                "osb_head_dw_allreduce: bad sizes splits=%d n=%lld stride=%lld", splits, (long long)n, (long long)split_stride);
    OSB_REQUIRE(n_extra >= 0 && n_plain >= 0 && n_plain <= n_extra, OSB_ERR_ARG, "osb_head_dw_allreduce: bad n_extra=%d / n_plain=%d", n_extra, n_plain);
    OSB_REQUIRE(n_extra == 0 || (peer_extra && extra_out), OSB_ERR_ARG, "osb_head_dw_allreduce: extras without buffers");
    ArParams p = {};
    for (int r = 0; r < world; ++r) {
        OSB_REQUIRE(peer_partials[r] && peer_dw[r] && peer_flags[r] && (n_extra == 0 || peer_extra[r]), OSB_ERR_ARG,
                    "osb_head_dw_allreduce: null peer pointer for rank %d", r);
        OSB_REQUIRE((uintptr_t)peer_partials[r] % 16 == 0 && (uintptr_t)peer_dw[r] % 16 == 0, OSB_ERR_ARG,
                    "osb_head_dw_allreduce: peer buffers must be 16-byte aligned");
        p.partials[r] = reinterpret_cast<const float*>(peer_partials[r]);
        p.out[r] = reinterpret_cast<float*>(peer_dw[r]);
        p.extra_in[r] = n_extra ? reinterpret_cast<const float*>(peer_extra[r]) : nullptr;
        p.flags[r] = reinterpret_cast<unsigned int*>(peer_flags[r]);
    }
    p.mc_partials = reinterpret_cast<const float*>(mc_partials); p.mc_out = reinterpret_cast<float*>(mc_dw);
    p.extra_out = extra_out; p.gscale = gscale; p.den = den; p.n = n; p.split_stride = split_stride;
    p.n_extra = n_extra; p.n_plain = n_plain; p.splits = splits; p.rank = rank; p.world = world;
    p.timeout_cycles = g_ar_timeout > 0 ? g_ar_timeout : 4000000000LL;        // ~2 s at 1.9 GHz
    int grid = max_ctas > 0 ? max_ctas : 32;
    const long long work = (n / 4 / world + kArThreads - 1) / kArThreads;
    if (grid > work) grid = (int)(work < 1 ? 1 : work);
    dw_allreduce_kernel<<<grid, kArThreads, 0, (cudaStream_t)stream>>>(p);
    OSB_LAUNCH_CHECK();
    return OSB_OK;
}

extern "C" int osb_head_dw_allreduce_error(const void* flags_local, int world, uint32_t* epoch_out) {
    OSB_ENTRY_GUARD();
    OSB_REQUIRE(flags_local && epoch_out && world >= 1 && world <= kArMaxWorld, OSB_ERR_ARG, "osb_head_dw_allreduce_error: bad argument");
    OSB_CUDA(cudaMemcpy(epoch_out, reinterpret_cast<const unsigned int*>(flags_local) + 2 * world, sizeof(uint32_t), cudaMemcpyDeviceToHost));
    return OSB_OK;
}

extern "C" size_t osb_peer_sum_bytes(int world) {
    (void)world;
    return sizeof(uint2) * (size_t)2 * kArMaxWorld * kPsMax * 2 + 64;
}

extern "C" int osb_peer_sum(const double* in, double* out, int n, void* const* peer_slots, int rank, int world, void* stream) {
    OSB_ENTRY_GUARD();
    OSB_REQUIRE(in && out && peer_slots, OSB_ERR_ARG, "osb_peer_sum: null pointer");
    OSB_REQUIRE(n >= 1 && n <= kPsMax, OSB_ERR_ARG, "osb_peer_sum: n=%d outside [1,%d]", n, kPsMax);
    OSB_REQUIRE(world >= 1 && world <= kArMaxWorld && rank >= 0 && rank < world, OSB_ERR_ARG, "osb_peer_sum: bad rank %d / world %d", rank, world);
    PsParams p = {};
    for (int r = 0; r < world; ++r) {
        OSB_REQUIRE(peer_slots[r] && (uintptr_t)peer_slots[r] % 8 == 0, OSB_ERR_ARG, "osb_peer_sum: bad slot pointer for rank %d", r);
        p.slots[r] = reinterpret_cast<uint2*>(peer_slots[r]);
    }
    p.in = in; p.out = out; p.n = n; p.rank = rank; p.world = world;
    p.timeout_cycles = g_ar_timeout > 0 ? g_ar_timeout : 4000000000LL;
    peer_sum_kernel<<<1, kArMaxWorld * kPsMax, 0, (cudaStream_t)stream>>>(p);
    OSB_LAUNCH_CHECK();
    return OSB_OK;
}
