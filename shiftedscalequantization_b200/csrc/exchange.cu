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
// (put = CAS 0->1 on the peer's flag with release.sys, wait = CAS 1->0 on the own flag with acquire.sys; flags reset
// themselves, so the pad needs no epoch). Arrival 1: "my gradient buffer is complete" (it was written by the previous kernel
// on this stream). Arrival 2: "my stores into your parameter buffer are done". A kernel completes only after all its CTAs
// have passed arrival 2, hence after every CTA of every rank has finished storing: the next kernel on any rank reads
// consistent alphas, and nobody overwrites a gradient buffer that is still being read. The grid is the same on every rank
// (it depends on n and R only) and small enough to be co-resident.
#include "ssq_common.cuh"

namespace ssq {

constexpr int XCHG_MAX_WORLD = 16;
constexpr int XCHG_MAX_CTAS = 64;
constexpr int XCHG_U = 2;                  // float4 vectors per thread per trip (x R gradient loads in flight each)
constexpr uint32_t XCHG_SPIN_LIMIT = 1u << 25;   // flag polls before a rendezvous gives up (tens of seconds): never hang the GPU

struct PeerTable {
    float* flat[XCHG_MAX_WORLD];           // parameter buffers (symmetric)
    const float* grad[XCHG_MAX_WORLD];     // gradient buffers (symmetric)
    uint32_t* pad[XCHG_MAX_WORLD];         // flag pads [XCHG_MAX_CTAS * 2][XCHG_MAX_WORLD]
};

__device__ __forceinline__ bool flag_put(uint32_t* addr) {
    uint32_t old, spins = 0;
    do {
        asm volatile("atom.global.release.sys.cas.b32 %0, [%1], 0, 1;" : "=r"(old) : "l"(addr) : "memory");
    } while (old != 0u && ++spins < XCHG_SPIN_LIMIT);
    return old == 0u;
}
__device__ __forceinline__ bool flag_wait(uint32_t* addr) {
    uint32_t old, spins = 0;
    do {
        asm volatile("atom.global.acquire.sys.cas.b32 %0, [%1], 1, 0;" : "=r"(old) : "l"(addr) : "memory");
    } while (old != 1u && ++spins < XCHG_SPIN_LIMIT);
    return old == 1u;
}
// all ranks' CTA `blockIdx.x` meet; `phase` selects one of the two flag rows of this CTA
__device__ __forceinline__ void cta_rendezvous(const PeerTable& P, int rank, int world, int phase, unsigned int* timeouts) {
    __syncthreads();
    const int row = (blockIdx.x * 2 + phase) * XCHG_MAX_WORLD;
    if ((int)threadIdx.x < world && (int)threadIdx.x != rank) {
        __threadfence_system();
        bool ok = flag_put(P.pad[threadIdx.x] + row + rank);       // tell peer t: rank `rank` arrived
        ok = flag_wait(P.pad[rank] + row + threadIdx.x) && ok;     // wait for peer t
        if (!ok) atomicAdd(timeouts, 1u);                           // a peer never showed up: the host raises after the run
    }
    __syncthreads();
}

__device__ __forceinline__ float4 ld_peer4(const float* p) {
    float4 v;
    asm volatile("ld.global.relaxed.sys.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_peer4(float* p, const float4& v) {
    asm volatile("st.global.relaxed.sys.v4.f32 [%0], {%1,%2,%3,%4};" :: "l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// n4 = number of float4 vectors of the flat buffers (padded to a multiple of 4 floats by the engine);
// shard of rank r = vectors [r*per, min((r+1)*per, n4)), per = ceil(n4 / world)
__global__ void __launch_bounds__(SSQ_THREADS)
exchange_adam_kernel(const __grid_constant__ PeerTable P, int rank, int world, int64_t n4,
                     float* __restrict__ m_shard, float* __restrict__ v_shard,
                     const float* __restrict__ lr_dev, int64_t* step_dev, unsigned int* ticket, unsigned int* timeouts,
                     double beta1d, double beta2d, double epsd, float* __restrict__ reduced_out) {
    __shared__ float s_step_size, s_bc2_sqrt;
    if (threadIdx.x == 0) {
        const double t = (double)(*step_dev + 1);
        s_step_size = (float)((double)__ldg(lr_dev) / (1.0 - pow(beta1d, t)));
        s_bc2_sqrt = (float)sqrt(1.0 - pow(beta2d, t));
    }
    const float w1 = (float)(1.0 - beta1d), w2 = (float)(1.0 - beta2d), beta2 = (float)beta2d, eps = (float)epsd;
    cta_rendezvous(P, rank, world, 0, timeouts);                        // every rank's gradient buffer is complete
    const float step_size = s_step_size, bc2_sqrt = s_bc2_sqrt;
    auto adam = [&](float& p, float g, float& mm, float& vv) {
        mm = mm + w1 * (g - mm);
        vv = vv * beta2 + w2 * g * g;
        const float denom = sqrtf(vv) / bc2_sqrt + eps;
        p = p - step_size * (mm / denom);
    };
    const int64_t per = (n4 + world - 1) / world;
    const int64_t lo = (int64_t)rank * per, hi = lo + per < n4 ? lo + per : n4;
    const int64_t stride = (int64_t)gridDim.x * SSQ_THREADS;
    for (int64_t base = lo + (int64_t)blockIdx.x * SSQ_THREADS + threadIdx.x; base < hi; base += stride * XCHG_U) {
        float4 g[XCHG_U], p[XCHG_U], mm[XCHG_U], vv[XCHG_U];
        // issue every load of the trip first: U x R gradient vectors (R-1 of them over NVLink) + the local shard state
#pragma unroll
        for (int u = 0; u < XCHG_U; ++u) {
            const int64_t i = base + (int64_t)u * stride;
            g[u] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (i < hi) {
                p[u] = *reinterpret_cast<const float4*>(P.flat[rank] + i * 4);
                mm[u] = *reinterpret_cast<const float4*>(m_shard + (i - lo) * 4);
                vv[u] = *reinterpret_cast<const float4*>(v_shard + (i - lo) * 4);
            }
        }
        float4 pg[XCHG_U][XCHG_MAX_WORLD > 8 ? 8 : XCHG_MAX_WORLD];
        for (int r0 = 0; r0 < world; r0 += 8) {               // rank order 0..R-1, eight peers' loads in flight at a time
#pragma unroll
            for (int u = 0; u < XCHG_U; ++u) {
                const int64_t i = base + (int64_t)u * stride;
#pragma unroll
                for (int k = 0; k < 8; ++k)
                    if (r0 + k < world && i < hi) pg[u][k] = ld_peer4(P.grad[r0 + k] + i * 4);
            }
#pragma unroll
            for (int u = 0; u < XCHG_U; ++u) {
#pragma unroll
                for (int k = 0; k < 8; ++k)
                    if (r0 + k < world) { g[u].x += pg[u][k].x; g[u].y += pg[u][k].y; g[u].z += pg[u][k].z; g[u].w += pg[u][k].w; }
            }
        }
#pragma unroll
        for (int u = 0; u < XCHG_U; ++u) {
            const int64_t i = base + (int64_t)u * stride;
            if (i < hi) {
                if (reduced_out) *reinterpret_cast<float4*>(reduced_out + (i - lo) * 4) = g[u];
                adam(p[u].x, g[u].x, mm[u].x, vv[u].x); adam(p[u].y, g[u].y, mm[u].y, vv[u].y);
                adam(p[u].z, g[u].z, mm[u].z, vv[u].z); adam(p[u].w, g[u].w, mm[u].w, vv[u].w);
                *reinterpret_cast<float4*>(m_shard + (i - lo) * 4) = mm[u];
                *reinterpret_cast<float4*>(v_shard + (i - lo) * 4) = vv[u];
                for (int r = 0; r < world; ++r) st_peer4(P.flat[r] + i * 4, p[u]);     // all-gather: every rank's parameter buffer
            }
        }
    }
    cta_rendezvous(P, rank, world, 1, timeouts);                        // everybody's stores have landed everywhere
    if (threadIdx.x == 0) {
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
                                      float* reduced_shard_out, unsigned int* timeouts, void* ws, size_t ws_bytes, void* stream) {
    if (!flat_ptrs || !grad_ptrs || !pad_ptrs || !exp_avg_shard || !exp_avg_sq_shard || !lr_dev || !step_dev || !timeouts) return SSQ_ERR_NULL;
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
    int64_t grid = (per + (int64_t)SSQ_THREADS * XCHG_U - 1) / ((int64_t)SSQ_THREADS * XCHG_U);   // a function of (n, world) only
    if (grid > XCHG_MAX_CTAS) grid = XCHG_MAX_CTAS;
    if (grid < 1) grid = 1;
    exchange_adam_kernel<<<(unsigned)grid, SSQ_THREADS, 0, (cudaStream_t)stream>>>(
        P, rank, world, n4, exp_avg_shard, exp_avg_sq_shard, lr_dev, step_dev,
        reinterpret_cast<unsigned int*>(ws) + (SSQ_WS_TICKETS - 1), timeouts, beta1, beta2, eps, reduced_shard_out);
    return launch_status();
}
