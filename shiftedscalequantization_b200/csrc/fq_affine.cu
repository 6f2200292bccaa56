// K1a — uniform affine fake-quant forward / STE backward, and the per-output-channel activation affine.
// HBM-bound streaming kernels: 128-bit loads/stores, several vectors in flight per thread, address-ordered tiles
// (forward) / per-channel slabs (backward), no integer division in the element loop, reciprocal of delta hoisted
// per vector.
//   reference arithmetic: quant/quant_layer.py:92-97, quant/channelQuantMSE.py:134-143,
//   quant/channelQuant.py:79-94, quant/quant_layer.py:258-259
#include "ssq_common.cuh"

namespace ssq {

// torch.clamp semantics (NaN propagates; bounds inclusive)
__device__ __forceinline__ float clampf(float v, float lo, float hi) {
    return v < lo ? lo : (v > hi ? hi : v);
}

// u = x/delta already divided
__device__ __forceinline__ float fq_post(float u, float d, float z, float qmin, float qmax, float& q) {
    q = clampf(__fadd_rn(rintf(u), z), qmin, qmax);
    return __fmul_rn(__fsub_rn(q, z), d);
}
__device__ __forceinline__ float fq_one(float x, const Recip& R, float z, float qmin, float qmax, float& q) {
    return fq_post(div_exact(x, R), R.d, z, qmin, qmax, q);
}
// ChannelQuantMSE: two successive divisions x/s/delta, dequant ((q-z)*delta)*s
__device__ __forceinline__ float fq_one_inscale(float x, float s, const Recip& R, float z, float qmin, float qmax, float& q) {
    q = clampf(__fadd_rn(rintf(div_exact(div_exact(x, s), R)), z), qmin, qmax);
    return __fmul_rn(__fmul_rn(__fsub_rn(q, z), R.d), s);
}

constexpr int UNROLL = 4;

// Vector path. One CTA = one contiguous run of 256*UNROLL float4s (16 KB); CTAs are dispatched in address order by the
// hardware scheduler, so the set of DRAM pages being streamed stays compact and there is no persistent-grid drift or
// ragged last wave (measured on B200: 6.75-6.8 TB/s against 5.3-6.4 TB/s for the same loop body under a persistent
// 148 x k grid-stride grid, and 6.5 TB/s for torch's copy). The channel of a thread's first vector costs one 32-bit
// division; its other vectors step by 256 with compares.
// CHAN: 0 = per-tensor (nchan == 1), 1 = per-channel with inner % 4 == 0
template <int CHAN, bool INSCALE, bool CODES>
__global__ void __launch_bounds__(SSQ_THREADS)
fq_affine_fwd_vec(const float* __restrict__ x, const float* __restrict__ delta, const float* __restrict__ zp,
                  const float* __restrict__ in_scale, float* __restrict__ y, float* __restrict__ codes,
                  uint32_t n4, uint32_t inner4, uint32_t nchan, float qmin, float qmax) {
    const uint32_t base = blockIdx.x * (SSQ_THREADS * UNROLL) + threadIdx.x;
    float4 v[UNROLL];
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
        const uint32_t i = base + u * SSQ_THREADS;
        if (i < n4) v[u] = ld_stream4(x + (size_t)i * 4);
    }
    Recip R; float z;
    uint32_t c = 0, col = 0;
    if (CHAN == 0) { R = make_recip(__ldg(delta)); z = __ldg(zp); }
    else { c = base / inner4; col = base - c * inner4; c %= nchan; }
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
        const uint32_t i = base + u * SSQ_THREADS;
        uint32_t mycol = 0;
        if (CHAN == 1) {
            while (col >= inner4) { col -= inner4; c = (c + 1 == nchan) ? 0 : c + 1; }
            R = make_recip(__ldg(delta + c)); z = __ldg(zp + c);
            mycol = col;
            col += SSQ_THREADS;
        }
        if (i < n4) {
            float4 q, o;
            if (INSCALE) {
                const float4 s = __ldg(reinterpret_cast<const float4*>(in_scale) + mycol);
                o.x = fq_one_inscale(v[u].x, s.x, R, z, qmin, qmax, q.x);
                o.y = fq_one_inscale(v[u].y, s.y, R, z, qmin, qmax, q.y);
                o.z = fq_one_inscale(v[u].z, s.z, R, z, qmin, qmax, q.z);
                o.w = fq_one_inscale(v[u].w, s.w, R, z, qmin, qmax, q.w);
            } else {
                const float4 t = div4_exact(v[u], R);
                o.x = fq_post(t.x, R.d, z, qmin, qmax, q.x);
                o.y = fq_post(t.y, R.d, z, qmin, qmax, q.y);
                o.z = fq_post(t.z, R.d, z, qmin, qmax, q.z);
                o.w = fq_post(t.w, R.d, z, qmin, qmax, q.w);
            }
            st_stream4(y + (size_t)i * 4, o);
            if (CODES) st_stream4(codes + (size_t)i * 4, q);
        }
    }
}

// scalar fallback: any inner / alignment (depthwise 3x3 rows of 9, the 7x7x3 stem, ragged tails)
template <bool INSCALE>
__global__ void __launch_bounds__(SSQ_THREADS)
fq_affine_fwd_scalar(const float* __restrict__ x, const float* __restrict__ delta, const float* __restrict__ zp,
                     const float* __restrict__ in_scale, float* __restrict__ y, float* __restrict__ codes,
                     int64_t n, int64_t inner, int64_t nchan, float qmin, float qmax) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t first = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    ChanWalk cw;
    cw.init(first, stride, inner, nchan);
    for (int64_t i = first; i < n; i += stride) {
        const Recip R = make_recip(__ldg(delta + cw.c));
        const float z = __ldg(zp + cw.c);
        float q;
        const float o = INSCALE ? fq_one_inscale(x[i], __ldg(in_scale + cw.col), R, z, qmin, qmax, q)
                                : fq_one(x[i], R, z, qmin, qmax, q);
        y[i] = o;
        if (codes) codes[i] = q;
        cw.next();
    }
}

// ------------------------------------------------------------------------------- backward
// One CTA per (split, channel): streams its slice of gy/x (outer slabs x a range of the channel's inner extent),
// writes gx, reduces (gdelta, gzp) partials in double; the last CTA of the channel finishes in fixed order.
template <bool VEC>
__global__ void __launch_bounds__(SSQ_THREADS)
fq_affine_bwd_kernel(const float* __restrict__ gy, const float* __restrict__ x, const float* __restrict__ delta,
                     const float* __restrict__ zp, float* __restrict__ gx, float* __restrict__ gdelta,
                     float* __restrict__ gzp, int64_t outer, int64_t inner, int64_t nchan, int64_t chunk,
                     float qmin, float qmax, WsView ws) {
    __shared__ double smem[2 * 32];
    const int64_t c = blockIdx.y;
    const int split = blockIdx.x, nsplit = gridDim.x;
    const Recip R = make_recip(__ldg(delta + c));
    const float d = R.d, z = __ldg(zp + c);
    // this CTA's [k0,k1) of the inner extent, for every outer slab
    const int64_t k0 = (int64_t)split * chunk;
    const int64_t k1 = k0 + chunk < inner ? k0 + chunk : inner;
    double acc[2] = {0.0, 0.0};
    float sd = 0.f, sz = 0.f;  // fp32 running sums flushed to double after every vector
    auto post = [&](float g, float u, float& gxo) {
        const float r = rintf(u);
        const float xi = __fadd_rn(r, z);
        const bool inside = (xi >= qmin) && (xi <= qmax);
        const float q = clampf(xi, qmin, qmax);
        gxo = inside ? g : 0.f;
        sd += g * (inside ? (r - u) : (q - z));
        sz += inside ? 0.f : -(g * d);
    };
    for (int64_t o = 0; o < outer; ++o) {
        const int64_t base = (o * nchan + c) * inner;
        if (VEC) {
            constexpr int U = 2;
            for (int64_t k = k0 + (int64_t)threadIdx.x * 4; k < k1; k += (int64_t)blockDim.x * 4 * U) {
                float4 g[U], xv[U];
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const int64_t kk = k + (int64_t)u * blockDim.x * 4;
                    if (kk < k1) { g[u] = ld_stream4(gy + base + kk); xv[u] = ld_stream4(x + base + kk); }
                }
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const int64_t kk = k + (int64_t)u * blockDim.x * 4;
                    if (kk < k1) {
                        float4 r4;
                        const float4 t = div4_exact(xv[u], R);
                        post(g[u].x, t.x, r4.x); post(g[u].y, t.y, r4.y); post(g[u].z, t.z, r4.z); post(g[u].w, t.w, r4.w);
                        if (gx) st_stream4(gx + base + kk, r4);
                        acc[0] += (double)sd; acc[1] += (double)sz; sd = 0.f; sz = 0.f;
                    }
                }
            }
        } else {
            for (int64_t k = k0 + threadIdx.x; k < k1; k += blockDim.x) {
                float r;
                post(gy[base + k], div_exact(x[base + k], R), r);
                if (gx) gx[base + k] = r;
                acc[0] += (double)sd; acc[1] += (double)sz; sd = 0.f; sz = 0.f;
            }
        }
    }
    if (gdelta == nullptr && gzp == nullptr) return;
    block_sum<2>(acc, smem);
    if (grid_finish<2>(acc, ws, c, split, nsplit, smem) && threadIdx.x == 0) {
        if (gdelta) gdelta[c] = (float)acc[0];
        if (gzp) gzp[c] = (float)acc[1];
    }
}

// ------------------------------------------------------------------------------- channel affine
__global__ void __launch_bounds__(SSQ_THREADS)
chan_affine_fwd_kernel(const float* __restrict__ x, const float* __restrict__ a, const float* __restrict__ b,
                       float* __restrict__ y, int64_t n, int64_t inner, int64_t nchan) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t first = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    ChanWalk cw;
    cw.init(first, stride, inner, nchan);
    for (int64_t i = first; i < n; i += stride) {
        y[i] = __fadd_rn(__fmul_rn(x[i], __ldg(a + cw.c)), __ldg(b + cw.c));
        cw.next();
    }
}

__global__ void __launch_bounds__(SSQ_THREADS)
chan_affine_bwd_kernel(const float* __restrict__ gy, const float* __restrict__ x, const float* __restrict__ a,
                       float* __restrict__ gx, float* __restrict__ ga, float* __restrict__ gb,
                       int64_t outer, int64_t inner, int64_t nchan, int64_t chunk, WsView ws) {
    __shared__ double smem[2 * 32];
    const int64_t c = blockIdx.y;
    const int split = blockIdx.x, nsplit = gridDim.x;
    const float av = __ldg(a + c);
    const int64_t k0 = (int64_t)split * chunk;
    const int64_t k1 = k0 + chunk < inner ? k0 + chunk : inner;
    double acc[2] = {0.0, 0.0};
    for (int64_t o = 0; o < outer; ++o) {
        const int64_t base = (o * nchan + c) * inner;
        for (int64_t k = k0 + threadIdx.x; k < k1; k += blockDim.x) {
            const float g = gy[base + k];
            if (gx) gx[base + k] = g * av;
            acc[0] += (double)(g * x[base + k]);
            acc[1] += (double)g;
        }
    }
    block_sum<2>(acc, smem);
    if (grid_finish<2>(acc, ws, c, split, nsplit, smem) && threadIdx.x == 0) {
        if (ga) ga[c] = (float)acc[0];
        if (gb) gb[c] = (float)acc[1];
    }
}

// split a channel's inner extent into CTAs so that nchan x nsplit CTAs fill the machine
static inline void chan_split(int64_t inner, int64_t outer, int64_t nchan, int vec, int64_t& chunk, int& nsplit) {
    const int64_t cap = SSQ_MAX_SLOTS;          // many small address-ordered slabs, bounded by the workspace's partial slots
    const int64_t want = (cap + nchan - 1) / nchan;
    const int64_t per_cta = (int64_t)SSQ_THREADS * 4 * (vec ? 2 : 1);
    const int64_t by_work = (inner * outer + per_cta * outer - 1) / (per_cta * outer);
    int64_t s = want < by_work ? want : by_work;
    if (s < 1) s = 1;
    chunk = (inner + s - 1) / s;
    chunk = (chunk + 3) / 4 * 4;  // keep float4 alignment of every split
    nsplit = (int)((inner + chunk - 1) / chunk);
    if (nsplit < 1) nsplit = 1;
}

}  // namespace ssq

using namespace ssq;

extern "C" size_t ssq_ws_bytes(int64_t nchan) {
    if (nchan < 1) nchan = 1;
    int64_t cap = SSQ_MAX_SLOTS;
    return ws_ticket_bytes(nchan) + (size_t)(nchan + cap + 64) * 4 * sizeof(double);
}

extern "C" int ssq_fq_affine_fwd(const float* x, const float* delta, const float* zero_point,
                                 const float* in_scale, float* y, float* codes,
                                 int64_t n, int64_t inner, int64_t nchan,
                                 float qmin, float qmax, void* stream) {
    if (n == 0) return SSQ_OK;
    if (!x || !delta || !zero_point || !y) return SSQ_ERR_NULL;
    if (n < 0 || inner <= 0 || nchan <= 0 || n % inner != 0 || (n / inner) % nchan != 0) return SSQ_ERR_SIZE;
    if (in_scale && n != inner * nchan) return SSQ_ERR_SIZE;
    cudaStream_t st = (cudaStream_t)stream;
    bool ptr_ok = aligned16(x) && aligned16(y) && (!codes || aligned16(codes)) && (!in_scale || aligned16(in_scale));
    bool vec = ptr_ok && (inner % 4 == 0) && (n / 4 < (int64_t)0x7fffffff) && n >= 4;
    if (vec) {
        uint32_t n4 = (uint32_t)(n / 4), inner4 = (uint32_t)(inner / 4);
        int64_t ctas = ((int64_t)n4 + SSQ_THREADS * UNROLL - 1) / (SSQ_THREADS * UNROLL);
        bool per_tensor = (nchan == 1 && !in_scale);
#define LAUNCH(CH, IS, CO) fq_affine_fwd_vec<CH, IS, CO><<<(unsigned)ctas, SSQ_THREADS, 0, st>>>( \
        x, delta, zero_point, in_scale, y, codes, n4, inner4, (uint32_t)nchan, qmin, qmax)
        if (per_tensor) { if (codes) LAUNCH(0, false, true); else LAUNCH(0, false, false); }
        else if (in_scale) { if (codes) LAUNCH(1, true, true); else LAUNCH(1, true, false); }
        else { if (codes) LAUNCH(1, false, true); else LAUNCH(1, false, false); }
#undef LAUNCH
    } else {
        int64_t ctas = (n + SSQ_THREADS - 1) / SSQ_THREADS;
        int grid = grid_for(ctas);
        if (in_scale) fq_affine_fwd_scalar<true><<<grid, SSQ_THREADS, 0, st>>>(x, delta, zero_point, in_scale, y, codes, n, inner, nchan, qmin, qmax);
        else fq_affine_fwd_scalar<false><<<grid, SSQ_THREADS, 0, st>>>(x, delta, zero_point, in_scale, y, codes, n, inner, nchan, qmin, qmax);
    }
    return launch_status();
}

extern "C" int ssq_fq_affine_bwd(const float* gy, const float* x, const float* delta, const float* zero_point,
                                 float* gx, float* gdelta, float* gzp,
                                 int64_t n, int64_t inner, int64_t nchan,
                                 float qmin, float qmax, void* ws, size_t ws_bytes, void* stream) {
    if (n == 0) return SSQ_OK;
    if (!gy || !x || !delta || !zero_point) return SSQ_ERR_NULL;
    if (n < 0 || inner <= 0 || nchan <= 0 || n % inner != 0 || (n / inner) % nchan != 0) return SSQ_ERR_SIZE;
    if (nchan > 65535) return SSQ_ERR_SIZE;
    if ((gdelta || gzp) && (!ws || ws_bytes < ssq_ws_bytes(nchan))) return SSQ_ERR_WORKSPACE;
    cudaStream_t st = (cudaStream_t)stream;
    int64_t outer = n / inner / nchan;
    bool vec = aligned16(gy) && aligned16(x) && (!gx || aligned16(gx)) && (inner % 4 == 0);
    int64_t chunk; int nsplit;
    chan_split(inner, outer, nchan, vec, chunk, nsplit);
    dim3 grid(nsplit, (unsigned)nchan);
    WsView v = ws_view(ws, nchan);
    if (vec) fq_affine_bwd_kernel<true><<<grid, SSQ_THREADS, 0, st>>>(gy, x, delta, zero_point, gx, gdelta, gzp, outer, inner, nchan, chunk, qmin, qmax, v);
    else fq_affine_bwd_kernel<false><<<grid, SSQ_THREADS, 0, st>>>(gy, x, delta, zero_point, gx, gdelta, gzp, outer, inner, nchan, chunk, qmin, qmax, v);
    return launch_status();
}

extern "C" int ssq_chan_affine_fwd(const float* x, const float* a, const float* b, float* y,
                                   int64_t n, int64_t inner, int64_t nchan, void* stream) {
    if (n == 0) return SSQ_OK;
    if (!x || !a || !b || !y) return SSQ_ERR_NULL;
    if (n < 0 || inner <= 0 || nchan <= 0 || n % (inner * nchan) != 0) return SSQ_ERR_SIZE;
    int grid = grid_for((n + SSQ_THREADS * 4 - 1) / (SSQ_THREADS * 4));
    chan_affine_fwd_kernel<<<grid, SSQ_THREADS, 0, (cudaStream_t)stream>>>(x, a, b, y, n, inner, nchan);
    return launch_status();
}

extern "C" int ssq_chan_affine_bwd(const float* gy, const float* x, const float* a, float* gx, float* ga, float* gb,
                                   int64_t n, int64_t inner, int64_t nchan, void* ws, size_t ws_bytes, void* stream) {
    if (n == 0) return SSQ_OK;
    if (!gy || !x || !a) return SSQ_ERR_NULL;
    if (n < 0 || inner <= 0 || nchan <= 0 || n % (inner * nchan) != 0 || nchan > 65535) return SSQ_ERR_SIZE;
    if (!ws || ws_bytes < ssq_ws_bytes(nchan)) return SSQ_ERR_WORKSPACE;
    int64_t outer = n / inner / nchan;
    int64_t chunk; int nsplit;
    chan_split(inner, outer, nchan, 0, chunk, nsplit);
    dim3 grid(nsplit, (unsigned)nchan);
    chan_affine_bwd_kernel<<<grid, SSQ_THREADS, 0, (cudaStream_t)stream>>>(gy, x, a, gx, ga, gb, outer, inner, nchan, chunk, ws_view(ws, nchan));
    return launch_status();
}
