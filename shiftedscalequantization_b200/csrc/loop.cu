// Calibration-loop plumbing that lets one reconstruction iteration live inside a CUDA graph:
//  - ssq_adam_step: torch.optim.Adam semantics (quant/block_recon.py:60,103) on one flat buffer per
//    reconstruction unit, step count and learning rate read from device memory;
//  - ssq_loop_advance: device-side iteration state (mini-batch indices = the reference's
//    torch.randperm(N)[:B] stream, precomputed on the host: quant/block_recon.py:90;
//    temperature b: quant/block_recon.py:185-202; lr: CosineAnnealingLR of :73).
#include "ssq_common.cuh"

namespace ssq {

__global__ void __launch_bounds__(SSQ_THREADS)
adam_kernel(float* __restrict__ param, const float* __restrict__ grad, float* __restrict__ m, float* __restrict__ v,
            int64_t n, const float* __restrict__ lr_dev, double beta1d, double beta2d, double epsd,
            int64_t* step_dev, unsigned int* ticket /* nullable: non-null = end the iteration */, int t_offset) {
    const bool vec = aligned16(param) && aligned16(grad) && aligned16(m) && aligned16(v);
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    // address-ordered tiles of SSQ_THREADS*U float4s, one per CTA (ssq_common.cuh): 8 loads in flight per thread
    constexpr int U = 2;
    const int64_t n4 = vec ? (n >> 2) : 0;
    const int64_t i0 = (int64_t)blockIdx.x * (SSQ_THREADS * U) + threadIdx.x;
    float4 p4[U], g4[U], m4[U], v4[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
        const int64_t i = i0 + (int64_t)u * SSQ_THREADS;
        if (i < n4) {
            p4[u] = *reinterpret_cast<float4*>(param + i * 4);
            g4[u] = ld_stream4(grad + i * 4);
            m4[u] = *reinterpret_cast<float4*>(m + i * 4);
            v4[u] = *reinterpret_cast<float4*>(v + i * 4);
        }
    }
    // host-side doubles of torch/optim/adam.py (_single_tensor_adam): bias corrections from the device step count,
    // evaluated by one thread per CTA while the tile's loads are in flight
    __shared__ AdamConst s_c;
    if (threadIdx.x == 0) s_c = adam_const(beta1d, beta2d, epsd, (double)(*step_dev + t_offset), __ldg(lr_dev));
    __syncthreads();
    const AdamConst c = s_c;
    auto one = [&](float& p, float g, float& mm, float& vv) { adam_update(p, g, mm, vv, c); };
#pragma unroll
    for (int u = 0; u < U; ++u) {
        const int64_t i = i0 + (int64_t)u * SSQ_THREADS;
        if (i < n4) {
            one(p4[u].x, g4[u].x, m4[u].x, v4[u].x); one(p4[u].y, g4[u].y, m4[u].y, v4[u].y);
            one(p4[u].z, g4[u].z, m4[u].z, v4[u].z); one(p4[u].w, g4[u].w, m4[u].w, v4[u].w);
            *reinterpret_cast<float4*>(param + i * 4) = p4[u];
            *reinterpret_cast<float4*>(m + i * 4) = m4[u];
            *reinterpret_cast<float4*>(v + i * 4) = v4[u];
        }
    }
    // tail (n % 4 elements, or everything when a pointer is not 16-byte aligned): grid-stride, scalar
    for (int64_t i = n4 * 4 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        one(param[i], grad[i], m[i], v[i]);
    if (ticket) {            // every CTA read *step_dev before the barrier above; the last one to retire ends the iteration
        __syncthreads();
        if (threadIdx.x == 0) {
            __threadfence();
            if (atomicAdd(ticket, 1u) == gridDim.x - 1) { *step_dev = *step_dev + 1; *ticket = 0u; __threadfence(); }
        }
    }
}

__global__ void loop_advance_kernel(int64_t* step_dev, const int64_t* __restrict__ idx_table, int64_t* idx_live, int batch,
                                    const float* __restrict__ b_table, float* b_live,
                                    const float* __restrict__ lr_table, float* lr_live, int64_t n_steps) {
    int64_t s = *step_dev;
    if (s >= n_steps) s = n_steps - 1;   // replays past the schedule keep the last row
    if (idx_table && idx_live)
        for (int j = threadIdx.x; j < batch; j += blockDim.x) idx_live[j] = idx_table[s * batch + j];
    if (threadIdx.x == 0) {
        if (b_table && b_live) *b_live = b_table[s];
        if (lr_table && lr_live) *lr_live = lr_table[s];
    }
    __syncthreads();
    if (threadIdx.x == 0) *step_dev = *step_dev + 1;
}

// Host-resident feature cache, pull variant: the SMs read the mini-batch rows straight out of mapped pinned host
// memory (UVA) with 128-bit loads, 4 in flight per thread, and write them to HBM. The row numbers come from the
// device-side index table at *step_dev + lookahead, so the transfer is a node of the captured iteration graph
// (forked beside the iteration it prefetches for) and costs the host nothing per step. PCIe-bound: a few dozen
// CTAs keep > 200 KB in flight, which is what a Gen5 x16 link needs; the grid is capped so the convolutions
// running beside it keep their SMs.
__global__ void __launch_bounds__(SSQ_THREADS)
pull_rows_kernel(const float* __restrict__ host_src, const int64_t* __restrict__ idx_table,
                 const int64_t* __restrict__ step_dev, int64_t lookahead, int64_t n_steps,
                 float* __restrict__ dst, int64_t batch, int64_t per_sample, bool vec) {
    int64_t s = *step_dev + lookahead;
    if (s >= n_steps) s = n_steps - 1;
    if (s < 0) s = 0;
    const int64_t* __restrict__ rows = idx_table + s * batch;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x, first = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (vec) {
        const int64_t ps4 = per_sample >> 2, total4 = batch * ps4;
        int64_t i = first;
        for (; i + 3 * stride < total4; i += 4 * stride) {
            float4 v[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int64_t j = i + k * stride, r = j / ps4;
                v[k] = ld_stream4(host_src + (__ldg(rows + r) * ps4 + (j - r * ps4)) * 4);
            }
#pragma unroll
            for (int k = 0; k < 4; ++k) st_stream4(dst + (i + k * stride) * 4, v[k]);
        }
        for (; i < total4; i += stride) {
            const int64_t r = i / ps4;
            st_stream4(dst + i * 4, ld_stream4(host_src + (__ldg(rows + r) * ps4 + (i - r * ps4)) * 4));
        }
    } else {
        const int64_t total = batch * per_sample;
        for (int64_t i = first; i < total; i += stride) {
            const int64_t r = i / per_sample;
            dst[i] = host_src[__ldg(rows + r) * per_sample + (i - r * per_sample)];
        }
    }
}


// Host-resident feature cache, zero-packed pull variant. Cached features are post-ReLU tensors: a third to a half of their
// elements are +0.0f, and the host-resident mode is PCIe-bound on the large early units. The host cache therefore keeps only the
// non-zero values, packed in element order; the index that locates them stays on the device (3 % of the dense size): a bit mask
// (1 bit per element: bit pattern != 0) and `chunk_off`, the offset of every 1024-element chunk's values. With the index in HBM a
// chunk costs ONE PCIe round trip. One warp expands one chunk: lane = one 32-bit mask word, the chunk's packed values
// are read with aligned 128-bit loads (all issued before the first use) into a per-warp shared-memory stage, a warp scan of the
// popcounts gives each lane its first value, and the dense row is written to HBM. Lossless by construction: only elements whose
// 32 bits are all zero are dropped (so -0.0f, denormals, NaN payloads survive).
#define SSQ_PACK_CHUNK 1024
struct PackedChunk {                       // one warp's view of one 1024-element chunk: everything that is in flight for it
    float4 v[9];                           // 9 x 32 float4 >= 3 + 1024 + 3 floats
    float* out;
    uint32_t m;
    int nvec, lead;                        // float4s to stage; values of the aligned over-read that precede the chunk's first
};
__global__ void __launch_bounds__(SSQ_THREADS, 2)
pull_rows_packed_kernel(const uint32_t* __restrict__ mask, const float* __restrict__ host_vals,
                        const int64_t* __restrict__ chunk_off, const int64_t* __restrict__ idx_table,
                        const int64_t* __restrict__ step_dev, int64_t lookahead, int64_t n_steps,
                        float* __restrict__ dst, int64_t batch, int64_t per_sample) {
    constexpr int WARPS = SSQ_THREADS / 32, NV = 9;
    __shared__ __align__(16) float stage[WARPS][SSQ_PACK_CHUNK + 16];
    int64_t s = *step_dev + lookahead;
    if (s >= n_steps) s = n_steps - 1;
    if (s < 0) s = 0;
    const int64_t* __restrict__ rows = idx_table + s * batch;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t C = per_sample / SSQ_PACK_CHUNK, W = per_sample / 32;
    const int64_t nwarps = (int64_t)gridDim.x * WARPS, total = batch * C;
    float* __restrict__ st = stage[warp];
    // issue every load of chunk `ch` (mask word from HBM, packed values over PCIe); nothing waits here
    auto load = [&](PackedChunk& k, int64_t ch) {
        const int64_t j = ch / C, c = ch - j * C, r = __ldg(rows + j);
        const int64_t g = r * C + c;
        const int64_t base = __ldg(chunk_off + g), cnt = __ldg(chunk_off + g + 1) - base;
        k.m = __ldg(mask + r * W + c * 32 + lane);
        const int64_t start = base & ~(int64_t)3;                 // aligned over-read of at most 3 + 3 values
        k.nvec = (int)((base + cnt - start + 3) >> 2);
        k.lead = (int)(base - start);
        k.out = dst + j * per_sample + c * SSQ_PACK_CHUNK + lane * 32;
#pragma unroll
        for (int u = 0; u < NV; ++u) {
            const int q = lane + 32 * u;
            k.v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (q < k.nvec) k.v[u] = ld_stream4(host_vals + start + 4 * (int64_t)q);
        }
    };
    // stage the values, scan the popcounts, write the dense 1024 elements
    auto expand = [&](const PackedChunk& k) {
        // every store's address depends on every load (x * 0 -> 0, NaN/Inf -> NaN -> cvt 0: always 0, but only known once all
        // loads have landed): ptxas otherwise stores the first vectors before issuing the last loads, i.e. two PCIe round trips
        float dep = fmaf(__uint_as_float(k.m), 0.0f, 0.0f);
#pragma unroll
        for (int u = 0; u < NV; ++u) dep = fmaf(k.v[u].x, 0.0f, dep);
        float* __restrict__ stz = st + __float2int_rz(dep);
#pragma unroll
        for (int u = 0; u < NV; ++u) {
            const int q = lane + 32 * u;
            if (q < k.nvec) *reinterpret_cast<float4*>(stz + 4 * q) = k.v[u];
        }
        __syncwarp();
        const int pc = __popc(k.m);
        int incl = pc;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
        int pos = k.lead + incl - pc;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            float o[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const bool on = (k.m >> (4 * q + e)) & 1u;
                o[e] = on ? st[pos] : 0.f;
                pos += on ? 1 : 0;
            }
            st_stream4(k.out + 4 * q, make_float4(o[0], o[1], o[2], o[3]));
        }
        __syncwarp();
    };
    // two chunks per warp in flight: the next chunk's loads are issued before the current one is expanded
    PackedChunk A, B;
    int64_t ch = (int64_t)blockIdx.x * WARPS + warp;
    bool has_a = ch < total;
    if (has_a) load(A, ch);
    while (has_a) {
        const int64_t nb = ch + nwarps;
        const bool has_b = nb < total;
        if (has_b) load(B, nb);
        expand(A);
        if (!has_b) break;
        ch = nb + nwarps;
        has_a = ch < total;
        if (has_a) load(A, ch);
        expand(B);
    }
}

}  // namespace ssq

using namespace ssq;

static int adam_launch(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n,
                       const float* lr_dev, double beta1, double beta2, double eps,
                       int64_t* step_dev, unsigned int* ticket, int t_offset, void* stream) {
    if (!param || !grad || !exp_avg || !exp_avg_sq || !lr_dev || !step_dev) return SSQ_ERR_NULL;
    if (n < 0) return SSQ_ERR_SIZE;
    const bool vec = aligned16(param) && aligned16(grad) && aligned16(exp_avg) && aligned16(exp_avg_sq);
    unsigned grid = vec ? tile_grid(((n >> 2) + SSQ_THREADS * 2 - 1) / (SSQ_THREADS * 2), false)
                        : (unsigned)grid_for((n + SSQ_THREADS - 1) / SSQ_THREADS);
    if (grid < 1) grid = 1;
    adam_kernel<<<grid, SSQ_THREADS, 0, (cudaStream_t)stream>>>(param, grad, exp_avg, exp_avg_sq, n, lr_dev, beta1, beta2, eps, step_dev, ticket, t_offset);
    return launch_status();
}

extern "C" int ssq_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n,
                             const float* lr_dev, double beta1, double beta2, double eps,
                             const int64_t* step_dev, void* stream) {
    if (n == 0) return SSQ_OK;
    return adam_launch(param, grad, exp_avg, exp_avg_sq, n, lr_dev, beta1, beta2, eps, const_cast<int64_t*>(step_dev), nullptr, 0, stream);
}

extern "C" int ssq_adam_step_pending(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n,
                                     const float* lr_dev, double beta1, double beta2, double eps,
                                     const int64_t* step_dev, void* stream) {
    if (n == 0) return SSQ_OK;
    return adam_launch(param, grad, exp_avg, exp_avg_sq, n, lr_dev, beta1, beta2, eps, const_cast<int64_t*>(step_dev), nullptr, 1, stream);
}

extern "C" int ssq_adam_step_end_iteration(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n,
                                           const float* lr_dev, double beta1, double beta2, double eps,
                                           int64_t* step_dev, void* ws, size_t ws_bytes, void* stream) {
    if (!ws || ws_bytes < ssq_ws_bytes(1)) return SSQ_ERR_WORKSPACE;
    return adam_launch(param, grad, exp_avg, exp_avg_sq, n, lr_dev, beta1, beta2, eps, step_dev,
                       reinterpret_cast<unsigned int*>(ws) + (SSQ_WS_TICKETS - 1), 1, stream);
}

extern "C" int ssq_loop_advance(int64_t* step_dev, const int64_t* idx_table, int64_t* idx_live, int batch,
                                const float* b_table, float* b_live, const float* lr_table, float* lr_live,
                                int64_t n_steps, void* stream) {
    if (!step_dev) return SSQ_ERR_NULL;
    if (n_steps <= 0 || batch < 0) return SSQ_ERR_SIZE;
    loop_advance_kernel<<<1, 128, 0, (cudaStream_t)stream>>>(step_dev, idx_table, idx_live, batch, b_table, b_live, lr_table, lr_live, n_steps);
    return launch_status();
}

extern "C" int ssq_stage_rows_h2d(const float* host_src, const int64_t* rows, float* dev_dst,
                                  int64_t batch, int64_t per_sample, void* stream) {
    if (batch == 0 || per_sample == 0) return SSQ_OK;
    if (!host_src || !rows || !dev_dst) return SSQ_ERR_NULL;
    if (batch < 0 || per_sample < 0) return SSQ_ERR_SIZE;
    const size_t bytes = (size_t)per_sample * sizeof(float);
    for (int64_t n = 0; n < batch; ++n) {
        cudaError_t e = cudaMemcpyAsync(dev_dst + n * per_sample, host_src + rows[n] * per_sample, bytes,
                                        cudaMemcpyHostToDevice, (cudaStream_t)stream);
        if (e != cudaSuccess) return (int)e;
    }
    return SSQ_OK;
}

extern "C" int ssq_pull_rows_host(const float* host_src_mapped, const int64_t* idx_table, const int64_t* step_dev,
                                  int64_t lookahead, int64_t n_steps, float* dev_dst, int64_t batch, int64_t per_sample,
                                  int max_ctas, void* stream) {
    if (batch == 0 || per_sample == 0) return SSQ_OK;
    if (!host_src_mapped || !idx_table || !step_dev || !dev_dst) return SSQ_ERR_NULL;
    if (batch < 0 || per_sample < 0 || n_steps <= 0) return SSQ_ERR_SIZE;
    const bool vec = (per_sample % 4 == 0) && aligned16(host_src_mapped) && aligned16(dev_dst);
    const int64_t total = batch * per_sample;
    const int64_t per_cta = (int64_t)SSQ_THREADS * (vec ? 16 : 1);
    int64_t grid = (total + per_cta - 1) / per_cta;
    const int64_t cap = max_ctas > 0 ? max_ctas : 32;
    if (grid > cap) grid = cap;
    if (grid < 1) grid = 1;
    pull_rows_kernel<<<(int)grid, SSQ_THREADS, 0, (cudaStream_t)stream>>>(host_src_mapped, idx_table, step_dev, lookahead, n_steps,
                                                                      dev_dst, batch, per_sample, vec);
    return launch_status();
}

extern "C" int ssq_pull_rows_host_packed(const uint32_t* mask, const float* host_vals_mapped, const int64_t* chunk_off,
                                         const int64_t* idx_table, const int64_t* step_dev, int64_t lookahead, int64_t n_steps,
                                         float* dev_dst, int64_t batch, int64_t per_sample, int max_ctas, void* stream) {
    if (batch == 0 || per_sample == 0) return SSQ_OK;
    if (!mask || !host_vals_mapped || !chunk_off || !idx_table || !step_dev || !dev_dst) return SSQ_ERR_NULL;
    if (batch < 0 || per_sample < 0 || n_steps <= 0 || per_sample % SSQ_PACK_CHUNK != 0) return SSQ_ERR_SIZE;
    if (!aligned16(host_vals_mapped) || !aligned16(dev_dst)) return SSQ_ERR_ALIGN;
    const int64_t chunks = batch * (per_sample / SSQ_PACK_CHUNK);
    int64_t grid = (chunks + SSQ_THREADS / 32 - 1) / (SSQ_THREADS / 32);
    const int64_t cap = max_ctas > 0 ? max_ctas : 32;
    if (grid > cap) grid = cap;
    pull_rows_packed_kernel<<<(int)grid, SSQ_THREADS, 0, (cudaStream_t)stream>>>(mask, host_vals_mapped, chunk_off, idx_table,
                                                                             step_dev, lookahead, n_steps, dev_dst, batch, per_sample);
    return launch_status();
}

extern "C" int ssq_abi_version(void) { return SSQ_ABI_VERSION; }

extern "C" const char* ssq_status_string(int status) {
    switch (status) {
        case SSQ_OK: return "ok";
        case SSQ_ERR_NULL: return "required pointer is NULL";
        case SSQ_ERR_SIZE: return "inconsistent or out-of-range size";
        case SSQ_ERR_WORKSPACE: return "workspace missing or too small";
        case SSQ_ERR_MODE: return "unknown mode / unsupported combination";
        case SSQ_ERR_ALIGN: return "pointer not 4-byte aligned";
        default: return status > 0 ? cudaGetErrorString((cudaError_t)status) : "unknown ssq status";
    }
}
