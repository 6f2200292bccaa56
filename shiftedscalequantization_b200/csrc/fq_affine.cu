// K1a — uniform affine fake-quant forward / STE backward, and the per-output-channel
// activation affine. HBM-bound streaming kernels: 128-bit loads/stores, 4 vectors in flight per
// thread, persistent grid = 148 SMs x 8 CTAs.
//   reference arithmetic: quant/quant_layer.py:92-97, quant/channelQuantMSE.py:134-143,
//   quant/channelQuant.py:79-94, quant/quant_layer.py:258-259
#include "ssq_common.cuh"

namespace ssq {

// torch.clamp semantics (NaN propagates; bounds inclusive)
__device__ __forceinline__ float clampf(float v, float lo, float hi) {
    return v < lo ? lo : (v > hi ? hi : v);
}

template <bool INSCALE>
__device__ __forceinline__ float fq_one(float x, float d, float z, float s, float qmin, float qmax, float& q) {
    float u = INSCALE ? div_exact(div_exact(x, s), d) : div_exact(x, d);
    q = clampf(__fadd_rn(rintf(u), z), qmin, qmax);
    float y = __fmul_rn(__fsub_rn(q, z), d);
    return INSCALE ? __fmul_rn(y, s) : y;
}

constexpr int UNROLL = 4;

// CHAN: 0 = per-tensor (nchan==1), 1 = per-channel with inner % 4 == 0
template <int CHAN, bool INSCALE, bool CODES>
__global__ void __launch_bounds__(SSQ_THREADS)
fq_affine_fwd_vec(const float* __restrict__ x, const float* __restrict__ delta, const float* __restrict__ zp,
                  const float* __restrict__ in_scale, float* __restrict__ y, float* __restrict__ codes,
                  uint32_t n4, uint32_t inner4, uint32_t nchan, float qmin, float qmax) {
    float d0 = 0.f, z0 = 0.f;
    if (CHAN == 0) { d0 = __ldg(delta); z0 = __ldg(zp); }
    const uint32_t stride = gridDim.x * blockDim.x;
    for (uint32_t base = blockIdx.x * blockDim.x + threadIdx.x; base < n4; base += stride * UNROLL) {
        float4 v[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            uint32_t i = base + u * stride;
            if (i < n4) v[u] = ld_stream4(x + (size_t)i * 4);
        }
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            uint32_t i = base + u * stride;
            if (i >= n4) break;
            float d = d0, z = z0;
            float4 s = make_float4(1.f, 1.f, 1.f, 1.f);
            if (CHAN == 1) {
                uint32_t row = i / inner4;
                uint32_t c = row % nchan;
                d = __ldg(delta + c); z = __ldg(zp + c);
                if (INSCALE) s = __ldg(reinterpret_cast<const float4*>(in_scale) + (i - row * inner4));
            }
            float4 q, o;
            o.x = fq_one<INSCALE>(v[u].x, d, z, s.x, qmin, qmax, q.x);
            o.y = fq_one<INSCALE>(v[u].y, d, z, s.y, qmin, qmax, q.y);
            o.z = fq_one<INSCALE>(v[u].z, d, z, s.z, qmin, qmax, q.z);
            o.w = fq_one<INSCALE>(v[u].w, d, z, s.w, qmin, qmax, q.w);
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
                     int64_t begin, int64_t n, int64_t inner, int64_t nchan, float qmin, float qmax) {
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = begin + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        int64_t row = i / inner;
        int64_t c = row % nchan;
        float s = INSCALE ? __ldg(in_scale + (i - row * inner)) : 1.f;
        float q;
        float o = fq_one<INSCALE>(x[i], __ldg(delta + c), __ldg(zp + c), s, qmin, qmax, q);
        y[i] = o;
        if (codes) codes[i] = q;
    }
}

// ------------------------------------------------------------------------------- backward
// One CTA per (split, channel): streams its slice of gy/x, writes gx, reduces
// (gdelta, gzp) partials in double; last CTA of the channel finishes in fixed order.
template <bool VEC>
__global__ void __launch_bounds__(SSQ_THREADS)
fq_affine_bwd_kernel(const float* __restrict__ gy, const float* __restrict__ x, const float* __restrict__ delta,
                     const float* __restrict__ zp, float* __restrict__ gx, float* __restrict__ gdelta,
                     float* __restrict__ gzp, int64_t outer, int64_t inner, int64_t nchan, int64_t chunk,
                     float qmin, float qmax, WsView ws) {
    __shared__ double smem[2 * 32];
    const int64_t c = blockIdx.y;
    const int split = blockIdx.x, nsplit = gridDim.x;
    const float d = __ldg(delta + c), z = __ldg(zp + c);
    const int64_t per_chan = outer * inner;
    int64_t j0 = (int64_t)split * chunk;
    int64_t j1 = j0 + chunk < per_chan ? j0 + chunk : per_chan;
    double acc[2] = {0.0, 0.0};
    float sd = 0.f, sz = 0.f;  // fp32 running sums flushed to double every few vectors
    auto one = [&](float g, float xv, float& gxo) {
        float u = div_exact(xv, d);
        float r = rintf(u);
        float xi = __fadd_rn(r, z);
        bool inside = (xi >= qmin) && (xi <= qmax);
        float q = clampf(xi, qmin, qmax);
        gxo = inside ? g : 0.f;
        sd += g * (inside ? (r - u) : (q - z));
        sz += inside ? 0.f : -(g * d);
    };
    if (VEC) {
        // inner % 4 == 0, pointers 16B aligned; j runs in units of 4 within the channel
        for (int64_t j = j0 + (int64_t)threadIdx.x * 4; j < j1; j += (int64_t)blockDim.x * 4) {
            int64_t o = j / inner;
            int64_t off = (o * nchan + c) * inner + (j - o * inner);
            float4 g = ld_stream4(gy + off), xv = ld_stream4(x + off), r4;
            one(g.x, xv.x, r4.x); one(g.y, xv.y, r4.y); one(g.z, xv.z, r4.z); one(g.w, xv.w, r4.w);
            if (gx) st_stream4(gx + off, r4);
            acc[0] += (double)sd; acc[1] += (double)sz; sd = 0.f; sz = 0.f;
        }
    } else {
        for (int64_t j = j0 + threadIdx.x; j < j1; j += blockDim.x) {
            int64_t o = j / inner;
            int64_t off = (o * nchan + c) * inner + (j - o * inner);
            float r;
            one(gy[off], x[off], r);
            if (gx) gx[off] = r;
            acc[0] += (double)sd; acc[1] += (double)sz; sd = 0.f; sz = 0.f;
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
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        int64_t c = (i / inner) % nchan;
        y[i] = __fadd_rn(__fmul_rn(x[i], __ldg(a + c)), __ldg(b + c));
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
    const int64_t per_chan = outer * inner;
    int64_t j0 = (int64_t)split * chunk;
    int64_t j1 = j0 + chunk < per_chan ? j0 + chunk : per_chan;
    double acc[2] = {0.0, 0.0};
    for (int64_t j = j0 + threadIdx.x; j < j1; j += blockDim.x) {
        int64_t o = j / inner;
        int64_t off = (o * nchan + c) * inner + (j - o * inner);
        float g = gy[off];
        if (gx) gx[off] = g * av;
        acc[0] += (double)(g * x[off]);
        acc[1] += (double)g;
    }
    block_sum<2>(acc, smem);
    if (grid_finish<2>(acc, ws, c, split, nsplit, smem) && threadIdx.x == 0) {
        if (ga) ga[c] = (float)acc[0];
        if (gb) gb[c] = (float)acc[1];
    }
}

// split a channel's outer*inner elements into CTAs so the grid fills the machine
static inline void chan_split(int64_t per_chan, int64_t nchan, int vec, int64_t& chunk, int& nsplit) {
    int64_t cap = (int64_t)SSQ_NUM_SMS * SSQ_CTAS_PER_SM;
    int64_t want = (cap + nchan - 1) / nchan;
    int64_t per_cta = (int64_t)SSQ_THREADS * 4 * (vec ? 4 : 1);
    int64_t by_work = (per_chan + per_cta - 1) / per_cta;
    int64_t s = want < by_work ? want : by_work;
    if (s < 1) s = 1;
    chunk = (per_chan + s - 1) / s;
    chunk = (chunk + 3) / 4 * 4;  // keep float4 alignment of every split
    nsplit = (int)((per_chan + chunk - 1) / chunk);
    if (nsplit < 1) nsplit = 1;
}

}  // namespace ssq

using namespace ssq;

extern "C" size_t ssq_ws_bytes(int64_t nchan) {
    if (nchan < 1) nchan = 1;
    int64_t cap = (int64_t)SSQ_NUM_SMS * SSQ_CTAS_PER_SM;
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
        int grid = grid_for(ctas);
        bool per_tensor = (nchan == 1 && !in_scale);
#define LAUNCH(CH, IS, CO) fq_affine_fwd_vec<CH, IS, CO><<<grid, SSQ_THREADS, 0, st>>>( \
        x, delta, zero_point, in_scale, y, codes, n4, inner4, (uint32_t)nchan, qmin, qmax)
        if (per_tensor) { if (codes) LAUNCH(0, false, true); else LAUNCH(0, false, false); }
        else if (in_scale) { if (codes) LAUNCH(1, true, true); else LAUNCH(1, true, false); }
        else { if (codes) LAUNCH(1, false, true); else LAUNCH(1, false, false); }
#undef LAUNCH
    } else {
        int64_t ctas = (n + SSQ_THREADS - 1) / SSQ_THREADS;
        int grid = grid_for(ctas);
        if (in_scale) fq_affine_fwd_scalar<true><<<grid, SSQ_THREADS, 0, st>>>(x, delta, zero_point, in_scale, y, codes, 0, n, inner, nchan, qmin, qmax);
        else fq_affine_fwd_scalar<false><<<grid, SSQ_THREADS, 0, st>>>(x, delta, zero_point, in_scale, y, codes, 0, n, inner, nchan, qmin, qmax);
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
    chan_split(outer * inner, nchan, vec, chunk, nsplit);
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
    chan_split(outer * inner, nchan, 0, chunk, nsplit);
    dim3 grid(nsplit, (unsigned)nchan);
    chan_affine_bwd_kernel<<<grid, SSQ_THREADS, 0, (cudaStream_t)stream>>>(gy, x, a, gx, ga, gb, outer, inner, nchan, chunk, ws_view(ws, nchan));
    return launch_status();
}
