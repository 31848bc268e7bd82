// Row-sharded OSCR over the GPUs of one box: all-gather of the compacted stage-A records through NVLink peer memory.
//
// Stage A (oscr_records_kernel) reduces a rank's rows to two compacted record arrays.  Thresholds are the GLOBAL set of
// unique known target scores (util.py:113), so every rank needs every record: ONE kernel per rank
//   1. publishes its four stage-A totals to every peer as (32 bits, epoch) words,
//   2. waits for everybody's, which gives it the rank-ordered offsets of its own records in the gathered arrays,
//   3. copies its records into EVERY rank's gathered arrays with peer stores (posted writes: bandwidth, not latency),
//   4. stamps "done" everywhere and waits for everybody's stamp.
// Afterwards every rank runs the same sort + scan (oscr_curve_kernel) on identical inputs, so the curve is replicated and
// bit-identical to the single-GPU result without any further exchange.  The reference has no counterpart.
// Re-use across calls: a rank can start call e + 1 only after every peer's done stamp of call e, i.e. after every peer has
// read the count words of call e; and it can write records of call e + 1 only after every peer's count words of call
// e + 1, i.e. after every peer has finished sorting call e.  Epochs live in the slot area (graph capturable).
#include "osb_common.cuh"

namespace osb {

constexpr int kOxMaxWorld = 16;
constexpr int kOxThreads = 512;

struct OxParams {
    uint64_t* dst_unk[kOxMaxWorld];
    uint64_t* dst_known[kOxMaxWorld];
    uint2* slots[kOxMaxWorld];              // per rank: [world][4] count words, [world] done stamps (as uint2.x), epoch, error, ticket
    const uint64_t* rec_unk; const uint64_t* rec_known;
    const unsigned long long* totals;       // local stage-A totals: n_known, n_correct, n_unk, n_bad
    long long* totals_global;               // local out [4]: summed over the ranks
    long long cap;
    int rank, world;
    long long timeout_cycles;
};

__device__ __forceinline__ void ox_st_ll(uint2* p, unsigned int v, unsigned int epoch) {
    asm volatile("st.relaxed.sys.global.v2.b32 [%0], {%1, %2};" :: "l"(p), "r"(v), "r"(epoch) : "memory");
}
__device__ __forceinline__ uint2 ox_ld_ll(const uint2* p) {
    uint2 v;
    asm volatile("ld.relaxed.sys.global.v2.b32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void ox_st_release(unsigned int* p, unsigned int v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" :: "l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned int ox_ld_acquire(const unsigned int* p) {
    unsigned int v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void ox_st_peer_u64(uint64_t* p, uint64_t v) {
    asm volatile("st.relaxed.sys.global.u64 [%0], %1;" :: "l"(p), "l"(v) : "memory");
}

__global__ void __launch_bounds__(kOxThreads)
oscr_exchange_kernel(const OxParams p) {
    const int W = p.world;
    uint2* mine = p.slots[p.rank];
    unsigned int* done = reinterpret_cast<unsigned int*>(mine + (size_t)kOxMaxWorld * 4);          // [world]
    unsigned int* ctl = done + kOxMaxWorld;                                                        // [0] epoch, [1] error, [2] ticket
    const unsigned int epoch = *reinterpret_cast<volatile unsigned int*>(ctl) + 1u;
    __shared__ unsigned int s_cnt[kOxMaxWorld][4];
    __shared__ int s_ok;
    __shared__ unsigned int s_last;
    if (threadIdx.x == 0) s_ok = 1;
    __syncthreads();
    // 1. my totals -> every rank (block 0 only), 2. everybody's totals <- my own slot area (every block for itself)
    if (blockIdx.x == 0 && threadIdx.x < W * 4) {
        const int r = threadIdx.x / 4, j = threadIdx.x % 4;
        ox_st_ll(p.slots[r] + (size_t)p.rank * 4 + j, (unsigned int)p.totals[j], epoch);
    }
    if (threadIdx.x < W * 4) {
        const int r = threadIdx.x / 4, j = threadIdx.x % 4;
        const long long t0 = clock64();
        uint2 v;
        bool ok;
        do {
            v = ox_ld_ll(mine + (size_t)r * 4 + j);
            ok = v.y == epoch;
        } while (!ok && clock64() - t0 <= p.timeout_cycles);
        if (!ok) { s_ok = 0; atomicExch(ctl + 1, epoch); }
        s_cnt[r][j] = v.x;
    }
    __syncthreads();
    const bool alive = s_ok != 0;
    // rank-ordered layout of the gathered arrays
    long long off_known = 0, off_unk = 0, tot[4] = {0, 0, 0, 0};
    for (int r = 0; r < W; ++r) {
        if (r < p.rank) { off_known += s_cnt[r][0]; off_unk += s_cnt[r][2]; }
#pragma unroll
        for (int j = 0; j < 4; ++j) tot[j] += s_cnt[r][j];
    }
    const long long n_known = s_cnt[p.rank][0], n_unk = s_cnt[p.rank][2];
    const bool fits = tot[0] <= p.cap && tot[2] <= p.cap;
    if (blockIdx.x == 0 && threadIdx.x < 4)
        p.totals_global[threadIdx.x] = alive && fits ? tot[threadIdx.x] : (threadIdx.x == 3 ? (1LL << 40) : 0);   // failure: no records, n_bad poisoned
    // 3. my records -> every rank's gathered arrays (peers first, so that the NVLink writes start at once)
    if (alive && fits) {
        const long long tstride = (long long)gridDim.x * kOxThreads;
        const long long t = (long long)blockIdx.x * kOxThreads + threadIdx.x;
        for (int d = 1; d <= W; ++d) {
            const int r = (p.rank + d) % W;
            uint64_t* dk = p.dst_known[r] + off_known;
            uint64_t* du = p.dst_unk[r] + off_unk;
            for (long long i = t; i < n_known; i += tstride) ox_st_peer_u64(dk + i, p.rec_known[i]);
            for (long long i = t; i < n_unk; i += tstride) ox_st_peer_u64(du + i, p.rec_unk[i]);
        }
    }
    // 4. every block's stores are out: the last block stamps "done" everywhere and waits for everybody's stamp
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned int tk = atomicAdd(ctl + 2, 1u);
        s_last = (tk == gridDim.x - 1) ? 1u : 0u;
        if (s_last) ctl[2] = 0u;
    }
    __syncthreads();
    if (s_last) {
        __threadfence_system();
        if (threadIdx.x < W) {
            unsigned int* peer_done = reinterpret_cast<unsigned int*>(p.slots[threadIdx.x] + (size_t)kOxMaxWorld * 4);
            ox_st_release(peer_done + p.rank, epoch);
            const long long t0 = clock64();
            while ((int)(ox_ld_acquire(done + threadIdx.x) - epoch) < 0) {
                if (clock64() - t0 > p.timeout_cycles) { atomicExch(ctl + 1, epoch); p.totals_global[3] = 1LL << 40; break; }
            }
        }
        __syncthreads();
        if (threadIdx.x == 0) ctl[0] = epoch;
    }
}

}  // namespace osb

using namespace osb;

static long long g_ox_timeout = 0;
extern "C" void osb_debug_set_exchange_timeout(long long cycles) { g_ox_timeout = cycles; }

extern "C" size_t osb_oscr_exchange_slot_bytes(int world) {
    (void)world;
    return sizeof(uint2) * (size_t)kOxMaxWorld * 4 + sizeof(unsigned int) * (kOxMaxWorld + 8);
}

extern "C" int osb_oscr_exchange(const uint64_t* rec_unk, const uint64_t* rec_known, const int64_t* totals,
                                 void* const* peer_rec_unk, void* const* peer_rec_known, void* const* peer_slots,
                                 int rank, int world, int64_t cap, int64_t* totals_global, int max_ctas, void* stream) {
    OSB_ENTRY_GUARD();
    OSB_REQUIRE(rec_unk && rec_known && totals && peer_rec_unk && peer_rec_known && peer_slots && totals_global, OSB_ERR_ARG,
                "osb_oscr_exchange: null pointer");
    OSB_REQUIRE(world >= 1 && world <= kOxMaxWorld && rank >= 0 && rank < world && cap >= 0 && cap < (1LL << 30), OSB_ERR_ARG,
                "osb_oscr_exchange: bad rank %d / world %d / cap %lld", rank, world, (long long)cap);
    OxParams p = {};
    for (int r = 0; r < world; ++r) {
        OSB_REQUIRE(peer_rec_unk[r] && peer_rec_known[r] && peer_slots[r], OSB_ERR_ARG, "osb_oscr_exchange: null peer pointer for rank %d", r);
        p.dst_unk[r] = reinterpret_cast<uint64_t*>(peer_rec_unk[r]);
        p.dst_known[r] = reinterpret_cast<uint64_t*>(peer_rec_known[r]);
        p.slots[r] = reinterpret_cast<uint2*>(peer_slots[r]);
    }
    p.rec_unk = rec_unk; p.rec_known = rec_known;
    p.totals = reinterpret_cast<const unsigned long long*>(totals);
    p.totals_global = reinterpret_cast<long long*>(totals_global);
    p.cap = cap; p.rank = rank; p.world = world;
    p.timeout_cycles = g_ox_timeout > 0 ? g_ox_timeout : 4000000000LL;
    int grid = max_ctas > 0 ? max_ctas : 32;
    if (grid > sm_count()) grid = sm_count();
    oscr_exchange_kernel<<<grid, kOxThreads, 0, (cudaStream_t)stream>>>(p);
    OSB_LAUNCH_CHECK();
    return OSB_OK;
}
