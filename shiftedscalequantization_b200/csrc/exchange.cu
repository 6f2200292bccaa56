// The one exchange step of the path: SUM of the unit's flat alpha gradient over the ranks every
// iteration (reference intent: link.allreduce(p.grad) per parameter, quant/block_recon.py:100-102), followed by
// optimizer.step (:103). An all-reduce followed by Adam on every rank moves 2(R-1)/R n floats per rank over the wire and
// then streams all five Adam arrays on EVERY rank. Here it is ONE kernel over peer memory (NVLink 5 / NVSwitch):
//
//   reduce-scatter   rank r sums elements [r n/R, (r+1) n/R) of every rank's gradient buffer with 128-bit loads from the
//                    peers' HBM (fixed rank order 0..R-1: the sum does not depend on who computes it),
//   Adam             on that shard only — exp_avg / exp_avg_sq exist only for the shard, 1/R of the optimizer traffic,
//   all-gather       the updated alphas are stored straight into every rank's parameter buffer,
//
// tile by tile, so the loads of the next vectors are in flight while the current ones are stepped and stored. The flat
// parameter and gradient buffers live in symmetric memory (torch.distributed._symmetric_memory: same offset on every
// rank, peer pointers exchanged once per unit); no NCCL call is on the path.
//
// Synchronisation: CTA b of every rank meets CTA b of all peers twice through a pad of 32-bit flags in symmetric memory
// (put = one st.release.sys of the launch's epoch into the peer's flag, wait = ld.acquire.sys polling of the own flag until
// it reaches that epoch; epochs only grow, nothing is reset). Arrival 1: "my gradient buffer is complete" (it was written by the previous kernel
// on this stream). Arrival 2: "my stores into your parameter buffer are done". A kernel completes only after all its CTAs
// have passed arrival 2, hence after every CTA of every rank has finished storing: the next kernel on any rank reads
// consistent alphas, and nobody overwrites a gradient buffer that is still being read. The grid is the same on every rank
// (it depends on n and R only) and small enough to be co-resident.
#include "ssq_common.cuh"
#include <stdlib.h>

namespace ssq {

constexpr int XCHG_MAX_WORLD = 16;
constexpr int XCHG_MAX_CTAS = 128;                // <= one CTA per SM: co-resident by construction
constexpr uint32_t XCHG_SPIN_LIMIT = 1u << 25;   // flag polls before a rendezvous gives up (tens of seconds): never hang the GPU

struct PeerTable {
    float* flat[XCHG_MAX_WORLD];           // parameter buffers (symmetric)
    const float* grad[XCHG_MAX_WORLD];     // gradient buffers (symmetric)
    uint32_t* pad[XCHG_MAX_WORLD];         // flag pads [XCHG_MAX_CTAS * 2][XCHG_MAX_WORLD]
};

// flags carry an epoch (the launch number of this CTA row, the same on every rank): a put is ONE one-way store over NVLink,
// a wait polls local memory; nothing is ever reset
__device__ __forceinline__ void flag_put(uint32_t* addr, uint32_t epoch, bool release) {
    if (release) asm volatile("st.global.release.sys.u32 [%0], %1;" :: "l"(addr), "r"(epoch) : "memory");
    else asm volatile("st.global.relaxed.sys.u32 [%0], %1;" :: "l"(addr), "r"(epoch) : "memory");
}
__device__ __forceinline__ bool flag_wait(const uint32_t* addr, uint32_t epoch) {
    uint32_t v, spins = 0;
    do {
        asm volatile("ld.global.acquire.sys.u32 %0, [%1];" : "=r"(v) : "l"(addr) : "memory");
    } while ((int32_t)(v - epoch) < 0 && ++spins < XCHG_SPIN_LIMIT);
    return (int32_t)(v - epoch) >= 0;
}
// all ranks' CTA `blockIdx.x` meet; `phase` selects one of the two flag rows of this CTA
__device__ __forceinline__ void cta_rendezvous(const PeerTable& P, int rank, int world, int phase, uint32_t epoch, unsigned int* timeouts) {
    __syncthreads();
    const int row = (blockIdx.x * 2 + phase) * XCHG_MAX_WORLD;
    if ((int)threadIdx.x < world && (int)threadIdx.x != rank) {
        // phase 0 publishes nothing this kernel wrote (the gradient buffer was completed by the previous kernel on this stream,
        // i.e. it is in L2, where peer loads are served): a relaxed flag. Phase 1 publishes this CTA's stores into the peers'
        // buffers: release at system scope (cumulative over the CTA's stores through the barrier above).
        flag_put(P.pad[threadIdx.x] + row + rank, epoch, phase != 0);       // tell peer t: rank `rank` reached `epoch`
        if (!flag_wait(P.pad[rank] + row + threadIdx.x, epoch))             // wait for peer t
            atomicAdd(timeouts, 1u);                                        // it never showed up: the host raises after the run
    }
    __syncthreads();
}

// Peer data moves with weak L2-coherent accesses (.cg: no L1 allocation). Ordering comes from the rendezvous: the producer's
// stores are fenced (fence.sc.sys via __threadfence_system) before its release.sys flag, the consumer's loads follow its
// acquire.sys on that flag, and no line of these buffers can sit in this SM's L1 from before the acquire.
// `mode` (SSQ_XCHG_MODE, experiments): bit 0 = relaxed.sys loads, bit 1 = relaxed.sys stores; 0 = weak .cg accesses
__device__ __forceinline__ float4 ld_peer4(const float* p, int mode) {
    float4 v;
    if (mode & 1) asm volatile("ld.global.relaxed.sys.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
    else asm volatile("ld.global.cg.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_peer4(float* p, const float4& v, int mode) {
    if (mode & 2) asm volatile("st.global.relaxed.sys.v4.f32 [%0], {%1,%2,%3,%4};" :: "l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
    else asm volatile("st.global.cg.v4.f32 [%0], {%1,%2,%3,%4};" :: "l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// n4 = number of float4 vectors of the flat buffers (padded to a multiple of 4 floats by the engine);
// shard of rank r = vectors [r*per, min((r+1)*per, n4)), per = ceil(n4 / world).
// W = ranks whose gradient vectors are loaded together (world <= W, or passes of W for larger worlds); U = vectors per thread
// per trip: U * W 128-bit loads in flight per thread, most of them crossing NVLink (latency ~2 us: the kernel needs megabytes
// in flight to fill the links).
template <int W, int U>
__global__ void __launch_bounds__(SSQ_THREADS)
exchange_adam_kernel(const __grid_constant__ PeerTable P, int rank, int world, int64_t n4,
                     float* __restrict__ m_shard, float* __restrict__ v_shard,
                     const float* __restrict__ lr_dev, int64_t* step_dev, unsigned int* ticket, unsigned int* timeouts,
                     double beta1d, double beta2d, double epsd, float* __restrict__ reduced_out, int mode,
                     uint32_t* __restrict__ epochs /* [XCHG_MAX_CTAS] local launch counters, one per CTA row */) {
    __shared__ AdamConst s_c;
    __shared__ uint32_t s_epoch;
    if (threadIdx.x == 0) {
        s_c = adam_const(beta1d, beta2d, epsd, (double)(*step_dev + 1), __ldg(lr_dev));
        s_epoch = epochs[blockIdx.x] + 1u;
    }
    __syncthreads();
    const uint32_t epoch = s_epoch;
    cta_rendezvous(P, rank, world, 0, epoch, timeouts);                 // every rank's gradient buffer is complete
    const AdamConst c = s_c;
    auto adam = [&](float& p, float g, float& mm, float& vv) { adam_update(p, g, mm, vv, c); };
    const int64_t per = (n4 + world - 1) / world;
    const int64_t lo = (int64_t)rank * per, hi = lo + per < n4 ? lo + per : n4;
    const int64_t stride = (int64_t)gridDim.x * SSQ_THREADS;
    for (int64_t base = lo + (int64_t)blockIdx.x * SSQ_THREADS + threadIdx.x; base < hi; base += stride * U) {
        float4 g[U];
#pragma unroll
        for (int u = 0; u < U; ++u) g[u] = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int r0 = 0; r0 < world; r0 += W) {               // rank order 0..R-1 (W ranks' loads in flight at a time)
            float4 pg[U][W];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int64_t i = base + (int64_t)u * stride;
#pragma unroll
                for (int k = 0; k < W; ++k)
                    pg[u][k] = (r0 + k < world && i < hi) ? ld_peer4(P.grad[r0 + k] + i * 4, mode) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
#pragma unroll
                for (int k = 0; k < W; ++k)
                    if (r0 + k < world) { g[u].x += pg[u][k].x; g[u].y += pg[u][k].y; g[u].z += pg[u][k].z; g[u].w += pg[u][k].w; }
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int64_t i = base + (int64_t)u * stride;
            if (i < hi) {
                float4 p = *reinterpret_cast<const float4*>(P.flat[rank] + i * 4);
                float4 mm = *reinterpret_cast<const float4*>(m_shard + (i - lo) * 4);
                float4 vv = *reinterpret_cast<const float4*>(v_shard + (i - lo) * 4);
                if (reduced_out) *reinterpret_cast<float4*>(reduced_out + (i - lo) * 4) = g[u];
                adam(p.x, g[u].x, mm.x, vv.x); adam(p.y, g[u].y, mm.y, vv.y);
                adam(p.z, g[u].z, mm.z, vv.z); adam(p.w, g[u].w, mm.w, vv.w);
                *reinterpret_cast<float4*>(m_shard + (i - lo) * 4) = mm;
                *reinterpret_cast<float4*>(v_shard + (i - lo) * 4) = vv;
                for (int r = 0; r < world; ++r) st_peer4(P.flat[r] + i * 4, p, mode);        // all-gather: every rank's parameter buffer
            }
        }
    }
    cta_rendezvous(P, rank, world, 1, epoch, timeouts);                 // everybody's stores have landed everywhere
    if (threadIdx.x == 0) {
        epochs[blockIdx.x] = epoch;
        __threadfence();
        if (atomicAdd(ticket, 1u) == gridDim.x - 1) { *step_dev = *step_dev + 1; *ticket = 0u; __threadfence(); }
    }
}

}  // namespace ssq

using namespace ssq;

extern "C" size_t ssq_exchange_pad_bytes(void) { return (size_t)XCHG_MAX_CTAS * 2 * XCHG_MAX_WORLD * sizeof(uint32_t); }

extern "C" int64_t ssq_exchange_shard_elems(int64_t n, int world) {
    if (world < 1 || n < 0) return 0;
    const int64_t n4 = (n + 3) / 4;
    return ((n4 + world - 1) / world) * 4;
}

extern "C" int ssq_grad_exchange_adam(float* const* flat_ptrs, const float* const* grad_ptrs, uint32_t* const* pad_ptrs,
                                      int rank, int world, int64_t n,
                                      float* exp_avg_shard, float* exp_avg_sq_shard,
                                      const float* lr_dev, int64_t* step_dev, double beta1, double beta2, double eps,
                                      float* reduced_shard_out, unsigned int* timeouts, uint32_t* epochs,
                                      void* ws, size_t ws_bytes, void* stream) {
    if (!flat_ptrs || !grad_ptrs || !pad_ptrs || !exp_avg_shard || !exp_avg_sq_shard || !lr_dev || !step_dev || !timeouts || !epochs) return SSQ_ERR_NULL;
    if (world < 1 || world > XCHG_MAX_WORLD || rank < 0 || rank >= world || n < 0 || (n & 3)) return SSQ_ERR_SIZE;
    if (!ws || ws_bytes < ssq_ws_bytes(1)) return SSQ_ERR_WORKSPACE;
    PeerTable P;
    for (int r = 0; r < world; ++r) {
        if (!flat_ptrs[r] || !grad_ptrs[r] || !pad_ptrs[r]) return SSQ_ERR_NULL;
        if (!aligned16(flat_ptrs[r]) || !aligned16(grad_ptrs[r])) return SSQ_ERR_ALIGN;
        P.flat[r] = flat_ptrs[r]; P.grad[r] = grad_ptrs[r]; P.pad[r] = pad_ptrs[r];
    }
    for (int r = world; r < XCHG_MAX_WORLD; ++r) { P.flat[r] = nullptr; P.grad[r] = nullptr; P.pad[r] = nullptr; }
    const int64_t n4 = n >> 2;
    const int64_t per = (n4 + world - 1) / world;
    static const int mode = getenv("SSQ_XCHG_MODE") ? atoi(getenv("SSQ_XCHG_MODE")) : 0;
    const int U = world <= 2 ? 8 : (world <= 4 ? 4 : 2);
    int64_t grid = (per + (int64_t)SSQ_THREADS * U - 1) / ((int64_t)SSQ_THREADS * U);   // a function of (n, world) only
    if (grid > XCHG_MAX_CTAS) grid = XCHG_MAX_CTAS;
    if (grid < 1) grid = 1;
    unsigned int* ticket = reinterpret_cast<unsigned int*>(ws) + (SSQ_WS_TICKETS - 1);
#define XCHG_LAUNCH(W_, U_) exchange_adam_kernel<W_, U_><<<(unsigned)grid, SSQ_THREADS, 0, (cudaStream_t)stream>>>( \
        P, rank, world, n4, exp_avg_shard, exp_avg_sq_shard, lr_dev, step_dev, ticket, timeouts, beta1, beta2, eps, reduced_shard_out, mode, epochs)
    if (world <= 2) XCHG_LAUNCH(2, 8);
    else if (world <= 4) XCHG_LAUNCH(4, 4);
    else XCHG_LAUNCH(8, 2);
#undef XCHG_LAUNCH
    return launch_status();
}
