// K1b — AdaRound fake-quant: floor + rectified-sigmoid soft / hard rounding, with the rounding regulariser
// sum(1-|2h-1|^b) and its gradient folded into the same pass over (w, alpha).
//   reference arithmetic: quant/adaptive_rounding.py:49-74, quant/block_recon.py:167-174,
//   quant/channelQuant.py:66-78 (sym-aware bounds via qmin/qmax)
// Single-tensor entry points serve the nn.Module surface; the *_mt entry points run every QuantModule of a
// reconstruction unit in one launch from a descriptor table held in kernel-parameter space.
// Per element: one exact division (hoisted reciprocal), 2 MUFU for the sigmoid, a 15-instruction log2 and one
// MUFU.EX2 for t^b — about 45 instructions, under the ~66/element an HBM-bound 12 B/element kernel can issue.
#include "ssq_common.cuh"

namespace ssq {

__device__ __forceinline__ float clampf_(float v, float lo, float hi) {
    return v < lo ? lo : (v > hi ? hi : v);
}

struct AdaOut { float y, q, reg; };

// u = w/delta already divided
// FINITE: u came from the hoisted-reciprocal fast path (finite numerator, delta in range), so floor + r + z is finite and the
// clamp may use FMNMX (2 instructions) instead of the NaN-preserving compare/select pairs (4)
template <bool SOFT, bool REG, bool FINITE = false>
__device__ __forceinline__ AdaOut ada_fwd_one(float u, float a, const Recip& R, float z, float qmin, float qmax, float b) {
    AdaOut o;
    const float fl = floorf(u);
    float r;
    o.reg = 0.f;
    if (SOFT) {
        r = rect_sigmoid(a);
        if (REG) o.reg = reg_term(r, b);
    } else {
        r = (a >= 0.f) ? 1.f : 0.f;
    }
    const float xi = __fadd_rn(__fadd_rn(fl, r), z);
    o.q = FINITE ? fminf(fmaxf(xi, qmin), qmax) : clampf_(xi, qmin, qmax);
    o.y = __fmul_rn(__fsub_rn(o.q, z), R.d);
    return o;
}

// galpha for one element: reconstruction path + regulariser path
template <bool REC, bool REG>
__device__ __forceinline__ float ada_bwd_one(float g, float u, float a, const Recip& R, float z, float qmin, float qmax,
                                             float b, float lam_g) {
    float h;
    const float dh = rect_sigmoid_grad(a, h);
    float out = 0.f;
    if (REC) {
        const float xi = __fadd_rn(__fadd_rn(floorf(u), h), z);
        const bool inside = (xi >= qmin) && (xi <= qmax);
        out = inside ? (g * R.d) * dh : 0.f;
    }
    if (REG) out += lam_g * reg_term_grad(h, b) * dh;
    return out;
}

// body shared by the single-tensor and multi-tensor forward kernels: elements [e0, e1) of one tensor, visited by
// `nthr` threads of which this is number `tid`; REGON is decided once per launch from *b_dev.
template <bool SOFT, bool REGON>
__device__ __forceinline__ double ada_fwd_span(const float* __restrict__ w, const float* __restrict__ alpha,
                                               const float* __restrict__ delta, const float* __restrict__ zp,
                                               float* __restrict__ wq, float* __restrict__ codes, int64_t e0, int64_t e1,
                                               int64_t inner, int64_t nchan, float qmin, float qmax, float b,
                                               int64_t tid, int64_t nthr, bool vec,
                                               const float* __restrict__ gamma = nullptr /* [nchan] output-channel scale folded into wq */) {
    double acc = 0.0;
    ChanWalk cw;
    if (vec) {
        // local 32-bit vector indices relative to e0 (host guarantees (e1-e0)/4 < 2^31)
        const float4* __restrict__ w4 = reinterpret_cast<const float4*>(w + e0);
        const float4* __restrict__ a4 = reinterpret_cast<const float4*>(alpha + e0);
        float4* __restrict__ y4 = reinterpret_cast<float4*>(wq + e0);
        float4* __restrict__ c4 = codes ? reinterpret_cast<float4*>(codes + e0) : nullptr;
        const uint32_t n4 = (uint32_t)((e1 - e0) >> 2), step = (uint32_t)nthr;
        cw.init((uint64_t)(e0 >> 2) + (uint64_t)tid, nthr, inner >> 2, nchan);
        Recip R; float z = 0.f, gm = 1.f; uint32_t c_have = 0xffffffffu;      // channel constants: recomputed only when the channel changes
        auto body = [&](uint32_t j, const float4& wv, const float4& av) {
            if (cw.c != c_have) { R = make_recip(__ldg(delta + cw.c)); z = __ldg(zp + cw.c); if (gamma) gm = __ldg(gamma + cw.c); c_have = cw.c; }
            float4 y, q;
            AdaOut o;
            float rsum = 0.f;
            if (R.ok && small4(wv)) {
#define ONE(F) o = ada_fwd_one<SOFT, REGON, true>(div_fast(wv.F, R), av.F, R, z, qmin, qmax, b); y.F = o.y; q.F = o.q; rsum += o.reg;
                ONE(x) ONE(y) ONE(z) ONE(w)
#undef ONE
            } else {
#define ONE(F) o = ada_fwd_one<SOFT, REGON, false>(__fdiv_rn(wv.F, R.d), av.F, R, z, qmin, qmax, b); y.F = o.y; q.F = o.q; rsum += o.reg;
                ONE(x) ONE(y) ONE(z) ONE(w)
#undef ONE
            }
            if (gamma) { y.x = __fmul_rn(y.x, gm); y.y = __fmul_rn(y.y, gm); y.z = __fmul_rn(y.z, gm); y.w = __fmul_rn(y.w, gm); }
            st_stream4(reinterpret_cast<float*>(y4 + j), y);
            if (c4) st_stream4(reinterpret_cast<float*>(c4 + j), q);
            if (REGON) acc += (double)rsum;
            cw.next();
        };
        uint32_t j = (uint32_t)tid;
        for (; j + step < n4 && j < n4; j += 2 * step) {          // two vectors in flight, no per-vector bounds test
            const float4 w0 = ld_stream4(reinterpret_cast<const float*>(w4 + j)), a0 = ld_stream4(reinterpret_cast<const float*>(a4 + j));
            const float4 w1 = ld_stream4(reinterpret_cast<const float*>(w4 + j + step)), a1 = ld_stream4(reinterpret_cast<const float*>(a4 + j + step));
            body(j, w0, a0);
            body(j + step, w1, a1);
        }
        if (j < n4) body(j, ld_stream4(reinterpret_cast<const float*>(w4 + j)), ld_stream4(reinterpret_cast<const float*>(a4 + j)));
    } else {
        cw.init(e0 + tid, nthr, inner, nchan);
        for (int64_t i = e0 + tid; i < e1; i += nthr) {
            const Recip R = make_recip(__ldg(delta + cw.c));
            const AdaOut o = ada_fwd_one<SOFT, REGON>(div_exact(w[i], R), alpha[i], R, __ldg(zp + cw.c), qmin, qmax, b);
            wq[i] = gamma ? __fmul_rn(o.y, __ldg(gamma + cw.c)) : o.y;
            if (codes) codes[i] = o.q;
            if (REGON) acc += (double)o.reg;
            cw.next();
        }
    }
    return acc;
}

template <bool REC, bool REGON>
__device__ __forceinline__ void ada_bwd_span(const float* __restrict__ gwq, const float* __restrict__ w,
                                             const float* __restrict__ alpha, const float* __restrict__ delta,
                                             const float* __restrict__ zp, float* __restrict__ galpha, int64_t e0, int64_t e1,
                                             int64_t inner, int64_t nchan, float qmin, float qmax, float b, float lam_g,
                                             int accumulate, int64_t tid, int64_t nthr, bool vec) {
    ChanWalk cw;
    if (vec) {
        const float4* __restrict__ w4 = reinterpret_cast<const float4*>(w + e0);
        const float4* __restrict__ a4 = reinterpret_cast<const float4*>(alpha + e0);
        const float4* __restrict__ g4 = REC ? reinterpret_cast<const float4*>(gwq + e0) : nullptr;
        float4* __restrict__ o4 = reinterpret_cast<float4*>(galpha + e0);
        const uint32_t n4 = (uint32_t)((e1 - e0) >> 2), step = (uint32_t)nthr;
        cw.init((uint64_t)(e0 >> 2) + (uint64_t)tid, nthr, inner >> 2, nchan);
        auto body = [&](uint32_t j, const float4& wv, const float4& av, const float4& gv) {
            const Recip R = make_recip(__ldg(delta + cw.c));
            const float z = __ldg(zp + cw.c);
            float4 o, t = make_float4(0.f, 0.f, 0.f, 0.f);
            if (REC) t = div4_exact(wv, R);
            o.x = ada_bwd_one<REC, REGON>(gv.x, t.x, av.x, R, z, qmin, qmax, b, lam_g);
            o.y = ada_bwd_one<REC, REGON>(gv.y, t.y, av.y, R, z, qmin, qmax, b, lam_g);
            o.z = ada_bwd_one<REC, REGON>(gv.z, t.z, av.z, R, z, qmin, qmax, b, lam_g);
            o.w = ada_bwd_one<REC, REGON>(gv.w, t.w, av.w, R, z, qmin, qmax, b, lam_g);
            if (accumulate) {
                const float4 p = o4[j];
                o.x += p.x; o.y += p.y; o.z += p.z; o.w += p.w;
            }
            st_stream4(reinterpret_cast<float*>(o4 + j), o);
            cw.next();
        };
        const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
        auto ld = [&](const float4* p, uint32_t j) { return ld_stream4(reinterpret_cast<const float*>(p + j)); };
        uint32_t j = (uint32_t)tid;
        for (; j + step < n4 && j < n4; j += 2 * step) {
            const float4 w0 = ld(w4, j), a0 = ld(a4, j), g0 = REC ? ld(g4, j) : zero4;
            const float4 w1 = ld(w4, j + step), a1 = ld(a4, j + step), g1 = REC ? ld(g4, j + step) : zero4;
            body(j, w0, a0, g0);
            body(j + step, w1, a1, g1);
        }
        if (j < n4) body(j, ld(w4, j), ld(a4, j), REC ? ld(g4, j) : zero4);
    } else {
        cw.init(e0 + tid, nthr, inner, nchan);
        for (int64_t i = e0 + tid; i < e1; i += nthr) {
            const Recip R = make_recip(__ldg(delta + cw.c));
            const float o = ada_bwd_one<REC, REGON>(REC ? gwq[i] : 0.f, REC ? div_exact(w[i], R) : 0.f, alpha[i], R, __ldg(zp + cw.c),
                                                    qmin, qmax, b, lam_g);
            galpha[i] = accumulate ? galpha[i] + o : o;
            cw.next();
        }
    }
}

// ---- single tensor ----------------------------------------------------------------------------------------
#define SSQ_ST_TILE 8192      // elements per tile: 2048 float4s, 8 per thread, taken two at a time
#ifndef SSQ_K1B_FWD_CTAS
#define SSQ_K1B_FWD_CTAS 5      // 48 registers: 0.953 of the HBM peak (4 CTAs / 60 registers: 0.913, 6: 0.937, 8: 0.82)
#endif
template <bool SOFT, bool REG>
__global__ void __launch_bounds__(SSQ_THREADS, SSQ_K1B_FWD_CTAS)
ada_fwd_kernel(const float* __restrict__ w, const float* __restrict__ alpha, const float* __restrict__ delta,
               const float* __restrict__ zp, float* __restrict__ wq, float* __restrict__ codes,
               int64_t n, int64_t inner, int64_t nchan, float qmin, float qmax, bool vec,
               const float* __restrict__ b_dev, float lambda, float* __restrict__ reg_out, WsView ws, int tiles_per_cta) {
    __shared__ double smem[32];
    const float b = REG ? __ldg(b_dev) : 0.f;
    const bool reg_on = REG && (b > 0.f);
    double acc[1] = {0.0};
    const int64_t ntiles = (n + SSQ_ST_TILE - 1) / SSQ_ST_TILE;
    const int64_t t0 = (int64_t)blockIdx.x * tiles_per_cta;                    // address-ordered tiles (ssq_common.cuh)
    const int64_t t1 = t0 + tiles_per_cta < ntiles ? t0 + tiles_per_cta : ntiles;
    for (int64_t tile = t0; tile < t1; ++tile) {
        const int64_t e0 = tile * SSQ_ST_TILE, e1 = e0 + SSQ_ST_TILE < n ? e0 + SSQ_ST_TILE : n;
        acc[0] += reg_on ? ada_fwd_span<SOFT, true>(w, alpha, delta, zp, wq, codes, e0, e1, inner, nchan, qmin, qmax, b, threadIdx.x, blockDim.x, vec)
                         : ada_fwd_span<SOFT, false>(w, alpha, delta, zp, wq, codes, e0, e1, inner, nchan, qmin, qmax, b, threadIdx.x, blockDim.x, vec);
    }
    if (!REG) return;
    block_sum<1>(acc, smem);
    if (grid_finish<1>(acc, ws, 0, blockIdx.x, gridDim.x, smem) && threadIdx.x == 0)
        reg_out[0] = reg_on ? (float)((double)lambda * acc[0]) : 0.f;
}

__global__ void __launch_bounds__(SSQ_THREADS)
ada_bwd_kernel(const float* __restrict__ gwq, const float* __restrict__ w, const float* __restrict__ alpha,
               const float* __restrict__ delta, const float* __restrict__ zp, float* __restrict__ galpha,
               int64_t n, int64_t inner, int64_t nchan, float qmin, float qmax, bool vec,
               const float* __restrict__ b_dev, float lambda, const float* __restrict__ greg, int accumulate) {
    const float b = b_dev ? __ldg(b_dev) : 0.f;
    const bool reg_on = b_dev && (b > 0.f);
    const float lam_g = reg_on ? lambda * (greg ? __ldg(greg) : 1.f) : 0.f;
    const int64_t ntiles = (n + SSQ_ST_TILE - 1) / SSQ_ST_TILE;
    const int64_t tid = threadIdx.x, nthr = blockDim.x;
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int64_t e0 = tile * SSQ_ST_TILE, e1 = e0 + SSQ_ST_TILE < n ? e0 + SSQ_ST_TILE : n;
        if (gwq) {
            if (reg_on) ada_bwd_span<true, true>(gwq, w, alpha, delta, zp, galpha, e0, e1, inner, nchan, qmin, qmax, b, lam_g, accumulate, tid, nthr, vec);
            else ada_bwd_span<true, false>(gwq, w, alpha, delta, zp, galpha, e0, e1, inner, nchan, qmin, qmax, b, lam_g, accumulate, tid, nthr, vec);
        } else {
            if (reg_on) ada_bwd_span<false, true>(gwq, w, alpha, delta, zp, galpha, e0, e1, inner, nchan, qmin, qmax, b, lam_g, accumulate, tid, nthr, vec);
            else ada_bwd_span<false, false>(gwq, w, alpha, delta, zp, galpha, e0, e1, inner, nchan, qmin, qmax, b, lam_g, accumulate, tid, nthr, vec);
        }
    }
}

// alpha init: rest = w/delta - floor(w/delta); alpha = -log((zeta-gamma)/(rest-gamma) - 1)   (libm logf: one-off)
__global__ void __launch_bounds__(SSQ_THREADS)
ada_init_alpha_kernel(const float* __restrict__ w, const float* __restrict__ delta, float* __restrict__ alpha,
                      int64_t n, int64_t inner, int64_t nchan) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x, first = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    ChanWalk cw;
    cw.init(first, stride, inner, nchan);
    for (int64_t i = first; i < n; i += stride) {
        const float u = __fdiv_rn(w[i], __ldg(delta + cw.c));
        const float rest = __fsub_rn(u, floorf(u));
        const float t = __fsub_rn(__fdiv_rn(SSQ_STRETCH, __fsub_rn(rest, SSQ_GAMMA)), 1.0f);
        alpha[i] = -logf(t);
        cw.next();
    }
}

// ---- stand-alone regulariser ---------------------------------------------------------------
__global__ void __launch_bounds__(SSQ_THREADS)
round_reg_fwd_kernel(const float* __restrict__ v, int64_t n, const float* __restrict__ b_dev, float lambda,
                     float* __restrict__ reg_out, WsView ws) {
    __shared__ double smem[32];
    const float b = __ldg(b_dev);
    double acc[1] = {0.0};
    if (b > 0.f) {
        const int64_t stride = (int64_t)gridDim.x * blockDim.x;
        for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
            acc[0] += (double)reg_term(rect_sigmoid(v[i]), b);
    }
    block_sum<1>(acc, smem);
    if (grid_finish<1>(acc, ws, 0, blockIdx.x, gridDim.x, smem) && threadIdx.x == 0)
        reg_out[0] = (b > 0.f) ? (float)((double)lambda * acc[0]) : 0.f;
}

__global__ void __launch_bounds__(SSQ_THREADS)
round_reg_bwd_kernel(const float* __restrict__ v, int64_t n, const float* __restrict__ b_dev, float lambda,
                     const float* __restrict__ greg, float* __restrict__ gv, int accumulate) {
    const float b = __ldg(b_dev);
    const float lam_g = (b > 0.f) ? lambda * (greg ? __ldg(greg) : 1.f) : 0.f;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        float o = 0.f;
        if (b > 0.f) {
            float h;
            const float dh = rect_sigmoid_grad(v[i], h);
            o = lam_g * reg_term_grad(h, b) * dh;
        }
        gv[i] = accumulate ? gv[i] + o : o;
    }
}

// ---- multi-tensor ----------------------------------------------------------------------------
struct MtTable { ssq_adaround_desc d[SSQ_MT_MAX]; };   // lives in kernel parameter space
struct AffTable { ssq_affine_desc d[SSQ_MT_MAX]; };    // the layers' output affines (gamma^z, varphi^z), parallel to MtTable

__device__ __forceinline__ int find_desc(const MtTable& table, int count, int64_t tile) {
    int lo = 0;
#pragma unroll 1
    for (int i = 1; i < count; ++i) if (table.d[i].tile_begin <= tile) lo = i;
    return lo;
}
__device__ __forceinline__ bool desc_vec(const ssq_adaround_desc& D, bool bwd) {
    bool ok = (D.inner % 4 == 0) && aligned16(D.w) && aligned16(D.alpha);
    return bwd ? (ok && aligned16(D.gwq) && aligned16(D.galpha)) : (ok && aligned16(D.wq));
}

template <bool SOFT>
__global__ void __launch_bounds__(SSQ_THREADS, 4)
ada_fwd_mt_kernel(const __grid_constant__ MtTable table, int count, int64_t total_tiles,
                  const float* __restrict__ b_dev, float lambda, float* __restrict__ reg_out, WsView ws) {
    __shared__ double smem[32];
    __shared__ int s_which;
    const float b = (SOFT && b_dev) ? __ldg(b_dev) : 0.f;
    const bool reg_on = SOFT && reg_out && (b > 0.f);
    double acc[1] = {0.0};
    for (int64_t tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        __syncthreads();
        if (threadIdx.x == 0) s_which = find_desc(table, count, tile);
        __syncthreads();
        const ssq_adaround_desc& D = table.d[s_which];
        const int64_t e0 = (tile - D.tile_begin) * SSQ_MT_TILE;
        const int64_t e1 = e0 + SSQ_MT_TILE < D.n ? e0 + SSQ_MT_TILE : D.n;
        const bool vec = desc_vec(D, false);
        acc[0] += reg_on ? ada_fwd_span<SOFT, true>(D.w, D.alpha, D.delta, D.zero_point, D.wq, nullptr, e0, e1, D.inner, D.nchan,
                                                    D.qmin, D.qmax, b, threadIdx.x, blockDim.x, vec)
                         : ada_fwd_span<SOFT, false>(D.w, D.alpha, D.delta, D.zero_point, D.wq, nullptr, e0, e1, D.inner, D.nchan,
                                                     D.qmin, D.qmax, b, threadIdx.x, blockDim.x, vec);
    }
    if (!reg_out) return;
    block_sum<1>(acc, smem);
    if (grid_finish<1>(acc, ws, 0, blockIdx.x, gridDim.x, smem) && threadIdx.x == 0)
        reg_out[0] = reg_on ? (float)((double)lambda * acc[0]) : 0.f;
}

__global__ void __launch_bounds__(SSQ_THREADS)
ada_bwd_mt_kernel(const __grid_constant__ MtTable table, int count, int64_t total_tiles,
                  const float* __restrict__ b_dev, float lambda) {
    __shared__ int s_which;
    const float b = b_dev ? __ldg(b_dev) : 0.f;
    const bool reg_on = b_dev && (b > 0.f);
    const float lam_g = reg_on ? lambda : 0.f;
    for (int64_t tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        __syncthreads();
        if (threadIdx.x == 0) s_which = find_desc(table, count, tile);
        __syncthreads();
        const ssq_adaround_desc& D = table.d[s_which];
        const int64_t e0 = (tile - D.tile_begin) * SSQ_MT_TILE;
        const int64_t e1 = e0 + SSQ_MT_TILE < D.n ? e0 + SSQ_MT_TILE : D.n;
        const bool vec = desc_vec(D, true);
        if (reg_on) ada_bwd_span<true, true>(D.gwq, D.w, D.alpha, D.delta, D.zero_point, D.galpha, e0, e1, D.inner, D.nchan,
                                             D.qmin, D.qmax, b, lam_g, 0, threadIdx.x, blockDim.x, vec);
        else ada_bwd_span<true, false>(D.gwq, D.w, D.alpha, D.delta, D.zero_point, D.galpha, e0, e1, D.inner, D.nchan,
                                       D.qmin, D.qmax, b, lam_g, 0, threadIdx.x, blockDim.x, vec);
    }
}

// ---- one reconstruction iteration in three launches -------------------------------------------------------
// (quant/block_recon.py:89-105). Launch 1 = this prologue: the iteration's bookkeeping (mini-batch row of the index table,
// temperature b, learning rate — all read through the device iteration counter), the gather of the mini-batch's cached input
// rows, and the soft forward of every quantised layer of the unit + the regulariser. The counter `*step` = iterations completed
// so far is only READ during an iteration; the fused Adam of launch 3 (or ssq_adam_step_ex) increments it when its last CTA
// retires. CTAs [0, gather_tiles) copy rows, CTAs [gather_tiles, gather_tiles + ada_ctas) walk the AdaRound tiles.
struct IterState {
    const int64_t* step; const int64_t* idx_table; int64_t* idx_live;
    const float* b_table; float* b_live; const float* lr_table; float* lr_live;
    int64_t n_steps; int batch;
};
__device__ __forceinline__ int64_t iter_row(const IterState& S) {
    int64_t s = *S.step;
    if (s >= S.n_steps) s = S.n_steps - 1;          // replays past the schedule keep the last row
    return s < 0 ? 0 : s;
}

__global__ void __launch_bounds__(SSQ_THREADS, 4)
iter_prologue_kernel(const __grid_constant__ IterState S, const float* __restrict__ cache, float* __restrict__ cur,
                     int64_t per_sample, int gather_tiles, bool gvec,
                     const __grid_constant__ MtTable table, const __grid_constant__ AffTable aff, bool has_aff,
                     int count, int64_t total_tiles, int ada_ctas,
                     float lambda, float* __restrict__ reg_out, WsView ws) {
    __shared__ double smem[32];
    __shared__ int s_which;
    const int64_t s = iter_row(S);
    if (blockIdx.x == 0) {                           // publish the live row for the kernels that follow
        if (S.idx_live) for (int j = threadIdx.x; j < S.batch; j += blockDim.x) S.idx_live[j] = S.idx_table[s * S.batch + j];
        if (threadIdx.x == 0) {
            if (S.b_live) *S.b_live = S.b_table ? S.b_table[s] : 0.f;
            if (S.lr_live && S.lr_table) *S.lr_live = S.lr_table[s];
        }
    }
    if ((int)blockIdx.x < gather_tiles) {
        constexpr int U = 4;
        const int64_t* __restrict__ rows = S.idx_table + s * S.batch;
        TileWalk tw;
        const int64_t i0 = (int64_t)blockIdx.x * (SSQ_THREADS * U) + threadIdx.x;
        if (gvec) {
            const int64_t ps4 = per_sample >> 2, total4 = (int64_t)S.batch * ps4;
            tw.init((uint64_t)i0, (uint64_t)ps4, (uint64_t)S.batch);
            float4 v[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int64_t i = i0 + (int64_t)u * SSQ_THREADS;
                if (i < total4) v[u] = ld_stream4(cache + (__ldg(rows + tw.c) * ps4 + tw.col) * 4);
                tw.step(SSQ_THREADS);
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int64_t i = i0 + (int64_t)u * SSQ_THREADS;
                if (i < total4) st_stream4(cur + i * 4, v[u]);
            }
        } else {
            const int64_t total = (int64_t)S.batch * per_sample;
            tw.init((uint64_t)i0, (uint64_t)per_sample, (uint64_t)S.batch);
            for (int u = 0; u < U; ++u) {
                const int64_t i = i0 + (int64_t)u * SSQ_THREADS;
                if (i < total) cur[i] = cache[__ldg(rows + tw.c) * per_sample + tw.col];
                tw.step(SSQ_THREADS);
            }
        }
        return;
    }
    if (count == 0) return;
    const int slot = (int)blockIdx.x - gather_tiles;
    const float b = S.b_table ? S.b_table[s] : 0.f;
    const bool reg_on = reg_out && (b > 0.f);
    double acc[1] = {0.0};
    for (int64_t tile = slot; tile < total_tiles; tile += ada_ctas) {
        __syncthreads();
        if (threadIdx.x == 0) s_which = find_desc(table, count, tile);
        __syncthreads();
        const ssq_adaround_desc& D = table.d[s_which];
        const int64_t e0 = (tile - D.tile_begin) * SSQ_MT_TILE;
        const int64_t e1 = e0 + SSQ_MT_TILE < D.n ? e0 + SSQ_MT_TILE : D.n;
        const bool vec = desc_vec(D, false);
        const float* gamma = has_aff ? aff.d[s_which].gamma : nullptr;
        if (has_aff && tile == D.tile_begin) {          // the layer's first tile also folds the bias: b_eff = gamma*b + phi
            const ssq_affine_desc& F = aff.d[s_which];
            for (int64_t c = threadIdx.x; c < D.nchan; c += blockDim.x) {
                const float bb = F.bias ? __ldg(F.bias + c) : 0.f;
                F.beff[c] = __fadd_rn(__fmul_rn(bb, __ldg(F.gamma + c)), __ldg(F.phi + c));
            }
        }
        acc[0] += reg_on ? ada_fwd_span<true, true>(D.w, D.alpha, D.delta, D.zero_point, D.wq, nullptr, e0, e1, D.inner, D.nchan,
                                                    D.qmin, D.qmax, b, threadIdx.x, blockDim.x, vec, gamma)
                         : ada_fwd_span<true, false>(D.w, D.alpha, D.delta, D.zero_point, D.wq, nullptr, e0, e1, D.inner, D.nchan,
                                                     D.qmin, D.qmax, b, threadIdx.x, blockDim.x, vec, gamma);
    }
    if (!reg_out) return;
    block_sum<1>(acc, smem);
    if (grid_finish<1>(acc, ws, 0, slot, ada_ctas, smem) && threadIdx.x == 0)
        reg_out[0] = reg_on ? (float)((double)lambda * acc[0]) : 0.f;
}

// gradients of the folded output affine (quant_layer.py:258-259 under autograd): for layer l, output channel c
//   d/d gamma = sum_k g_Weff[c,k] * W_q[c,k] + g_beff[c] * bias[c],   d/d phi = g_beff[c]
// (W_eff = gamma W_q, b_eff = gamma bias + phi; g_beff = the bias gradient = sum over N,H,W of d loss / d out).
// One warp per (layer, output channel); W_q is re-evaluated from (w, alpha) exactly as the forward did; fixed summation order.
__global__ void __launch_bounds__(SSQ_THREADS)
affine_grad_mt_kernel(const __grid_constant__ MtTable table, const __grid_constant__ AffTable aff, int count, int64_t total_rows,
                      const float* __restrict__ b_dev) {
    const int lane = threadIdx.x & 31;
    const int64_t row = (int64_t)blockIdx.x * (SSQ_THREADS / 32) + (threadIdx.x >> 5);
    if (row >= total_rows) return;
    int l = 0;
    for (int i = 1; i < count; ++i) if (aff.d[i].row_begin <= row) l = i;
    const ssq_adaround_desc& D = table.d[l];
    const ssq_affine_desc& F = aff.d[l];
    const int64_t c = row - F.row_begin, k = D.inner;       // channel-wise layers: inner = elements per output channel
    const Recip R = make_recip(__ldg(D.delta + c));
    const float z = __ldg(D.zero_point + c);
    const float* __restrict__ w = D.w + c * k;
    const float* __restrict__ a = D.alpha + c * k;
    const float* __restrict__ g = D.gwq + c * k;
    double acc = 0.0;
    for (int64_t j0 = lane; j0 < k; j0 += 32 * 8) {
        float sdot = 0.f;
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const int64_t j = j0 + e * 32;
            if (j < k) sdot = fmaf(g[j], ada_fwd_one<true, false>(div_exact(w[j], R), a[j], R, z, D.qmin, D.qmax, 0.f).y, sdot);
        }
        acc += (double)sdot;
    }
    acc = warp_sum(acc);
    if (lane == 0) {
        const float gb = __ldg(F.gbeff + c);
        F.ggamma[c] = (float)acc + (F.bias ? gb * __ldg(F.bias + c) : 0.f);
        F.gphi[c] = gb;
    }
    (void)b_dev;
}

// Launch 3: gradient of every alpha (reconstruction + regulariser) AND its Adam step in one pass over (g_wq, w, alpha, m, v):
// the gradient never goes to memory (unless galpha is wanted), torch.optim.Adam's arithmetic as in loop.cu. t = *step + 1;
// the last CTA to retire increments *step.
struct AdamArgs {
    float* flat; float* m; float* v;               // alpha lives at flat + off, its moments at m + off, v + off
    const float* lr; int64_t* step; unsigned int* ticket;
    double beta1, beta2, eps;
    int store_grad, apply_adam;                    // apply_adam == 0: gradients only (several GPUs: the exchange kernel steps)
};
__global__ void __launch_bounds__(SSQ_THREADS, 4)
ada_bwd_adam_mt_kernel(const __grid_constant__ MtTable table, const __grid_constant__ AffTable aff, bool has_aff,
                       int count, int64_t total_tiles,
                       const float* __restrict__ b_dev, float lambda, const __grid_constant__ AdamArgs A) {
    __shared__ int s_which;
    __shared__ AdamConst s_c;
    const float b = b_dev ? __ldg(b_dev) : 0.f;
    const bool reg_on = b_dev && (b > 0.f);
    const float lam_g = reg_on ? lambda : 0.f;
    if (threadIdx.x == 0) s_c = adam_const(A.beta1, A.beta2, A.eps, (double)(*A.step + 1), __ldg(A.lr));
    __syncthreads();
    const AdamConst c = s_c;
    auto adam = [&](float& p, float g, float& mm, float& vv) { adam_update(p, g, mm, vv, c); };
    for (int64_t tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        __syncthreads();
        if (threadIdx.x == 0) s_which = find_desc(table, count, tile);
        __syncthreads();
        const ssq_adaround_desc& D = table.d[s_which];
        const int64_t e0 = (tile - D.tile_begin) * SSQ_MT_TILE;
        const int64_t e1 = e0 + SSQ_MT_TILE < D.n ? e0 + SSQ_MT_TILE : D.n;
        const int64_t off = D.alpha - A.flat;
        float* __restrict__ pa = A.flat + off;
        float* __restrict__ pm = A.apply_adam ? A.m + off : pa;          // moments are untouched without the Adam step
        float* __restrict__ pv = A.apply_adam ? A.v + off : pa;
        ChanWalk cw;
        if (desc_vec(D, true) && aligned16(pm) && aligned16(pv)) {
            const uint32_t n4 = (uint32_t)((e1 - e0) >> 2);
            cw.init((uint64_t)(e0 >> 2) + threadIdx.x, blockDim.x, D.inner >> 2, D.nchan);
            // 4 vectors per thread (SSQ_MT_TILE / 4 / SSQ_THREADS), taken two at a time: 10 loads in flight per thread at 64
            // registers / 4 resident CTAs (all 20 at once cost 128 registers, 2 CTAs: 0.67 of the HBM peak)
            constexpr int NV = 2;
#pragma unroll 1
            for (int h = 0; h < SSQ_MT_TILE / 4 / SSQ_THREADS / NV; ++h) {
            float4 wv[NV], av[NV], gv[NV], mv[NV], vv[NV];
#pragma unroll
            for (int u = 0; u < NV; ++u) {
                const uint32_t j = threadIdx.x + (h * NV + u) * SSQ_THREADS;
                if (j < n4) {
                    const int64_t e = e0 + 4 * (int64_t)j;
                    wv[u] = ld_stream4(D.w + e); gv[u] = ld_stream4(D.gwq + e);
                    av[u] = *reinterpret_cast<const float4*>(pa + e);
                    if (A.apply_adam) {
                        mv[u] = *reinterpret_cast<const float4*>(pm + e);
                        vv[u] = *reinterpret_cast<const float4*>(pv + e);
                    }
                }
            }
#pragma unroll
            for (int u = 0; u < NV; ++u) {
                const uint32_t j = threadIdx.x + (h * NV + u) * SSQ_THREADS;
                if (j < n4) {
                    const int64_t e = e0 + 4 * (int64_t)j;
                    const Recip R = make_recip(__ldg(D.delta + cw.c));
                    const float z = __ldg(D.zero_point + cw.c);
                    const float4 t = div4_exact(wv[u], R);
                    if (has_aff) {                      // W_eff = gamma W_q: d loss / d W_q = gamma * g_Weff
                        const float gm = __ldg(aff.d[s_which].gamma + cw.c);
                        gv[u].x *= gm; gv[u].y *= gm; gv[u].z *= gm; gv[u].w *= gm;
                    }
                    float4 g;
                    if (reg_on) {
                        g.x = ada_bwd_one<true, true>(gv[u].x, t.x, av[u].x, R, z, D.qmin, D.qmax, b, lam_g);
                        g.y = ada_bwd_one<true, true>(gv[u].y, t.y, av[u].y, R, z, D.qmin, D.qmax, b, lam_g);
                        g.z = ada_bwd_one<true, true>(gv[u].z, t.z, av[u].z, R, z, D.qmin, D.qmax, b, lam_g);
                        g.w = ada_bwd_one<true, true>(gv[u].w, t.w, av[u].w, R, z, D.qmin, D.qmax, b, lam_g);
                    } else {
                        g.x = ada_bwd_one<true, false>(gv[u].x, t.x, av[u].x, R, z, D.qmin, D.qmax, b, lam_g);
                        g.y = ada_bwd_one<true, false>(gv[u].y, t.y, av[u].y, R, z, D.qmin, D.qmax, b, lam_g);
                        g.z = ada_bwd_one<true, false>(gv[u].z, t.z, av[u].z, R, z, D.qmin, D.qmax, b, lam_g);
                        g.w = ada_bwd_one<true, false>(gv[u].w, t.w, av[u].w, R, z, D.qmin, D.qmax, b, lam_g);
                    }
                    if (A.store_grad) st_stream4(D.galpha + e, g);
                    if (A.apply_adam) {
                        adam(av[u].x, g.x, mv[u].x, vv[u].x); adam(av[u].y, g.y, mv[u].y, vv[u].y);
                        adam(av[u].z, g.z, mv[u].z, vv[u].z); adam(av[u].w, g.w, mv[u].w, vv[u].w);
                        *reinterpret_cast<float4*>(pa + e) = av[u];
                        *reinterpret_cast<float4*>(pm + e) = mv[u];
                        *reinterpret_cast<float4*>(pv + e) = vv[u];
                    }
                }
                cw.next();
            }
            }
        } else {
            cw.init(e0 + threadIdx.x, blockDim.x, D.inner, D.nchan);
            for (int64_t i = e0 + threadIdx.x; i < e1; i += blockDim.x) {
                const Recip R = make_recip(__ldg(D.delta + cw.c));
                const float z = __ldg(D.zero_point + cw.c);
                const float u = div_exact(D.w[i], R);
                const float gin = has_aff ? D.gwq[i] * __ldg(aff.d[s_which].gamma + cw.c) : D.gwq[i];
                const float g = reg_on ? ada_bwd_one<true, true>(gin, u, pa[i], R, z, D.qmin, D.qmax, b, lam_g)
                                       : ada_bwd_one<true, false>(gin, u, pa[i], R, z, D.qmin, D.qmax, b, lam_g);
                if (A.store_grad) D.galpha[i] = g;
                if (A.apply_adam) adam(pa[i], g, pm[i], pv[i]);
                cw.next();
            }
        }
    }
    if (!A.apply_adam) return;
    // every CTA has read *step (thread 0, before the barrier above); the last one to arrive bumps it
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        if (atomicAdd(A.ticket, 1u) == gridDim.x - 1) { *A.step = *A.step + 1; *A.ticket = 0u; __threadfence(); }
    }
}

static inline int check_layout(int64_t n, int64_t inner, int64_t nchan) {
    if (n < 0 || inner <= 0 || nchan <= 0 || n % inner != 0 || (n / inner) % nchan != 0) return SSQ_ERR_SIZE;
    return SSQ_OK;
}

}  // namespace ssq

using namespace ssq;

extern "C" int ssq_fq_adaround_fwd(const float* w, const float* alpha, const float* delta, const float* zero_point,
                                   float* wq, float* codes, int64_t n, int64_t inner, int64_t nchan,
                                   float qmin, float qmax, int soft,
                                   const float* b_dev, float lambda, float* reg_out,
                                   void* ws, size_t ws_bytes, void* stream) {
    if (n == 0) return SSQ_OK;
    if (!w || !alpha || !delta || !zero_point || !wq) return SSQ_ERR_NULL;
    if (int e = check_layout(n, inner, nchan)) return e;
    const bool reg = reg_out != nullptr;
    if (reg && (!b_dev || !soft)) return SSQ_ERR_MODE;
    if (reg && (!ws || ws_bytes < ssq_ws_bytes(1))) return SSQ_ERR_WORKSPACE;
    cudaStream_t st = (cudaStream_t)stream;
    bool vec = aligned16(w) && aligned16(alpha) && aligned16(wq) && (!codes || aligned16(codes)) && (inner % 4 == 0) &&
               (n / 4 < (int64_t)0x7fffffff);
    const int64_t ntiles = (n + SSQ_ST_TILE - 1) / SSQ_ST_TILE;
    WsView v = ws_view(ws, 1);
    int per_cta = 1;
    const unsigned grid = reg ? tile_grid_balanced(ntiles, per_cta) : tile_grid(ntiles, false);
#define LAUNCH(S, R) ada_fwd_kernel<S, R><<<grid, SSQ_THREADS, 0, st>>>( \
        w, alpha, delta, zero_point, wq, codes, n, inner, nchan, qmin, qmax, vec, b_dev, lambda, reg_out, v, per_cta)
    if (soft) { if (reg) LAUNCH(true, true); else LAUNCH(true, false); }
    else LAUNCH(false, false);
#undef LAUNCH
    return launch_status();
}

extern "C" int ssq_fq_adaround_bwd(const float* gwq, const float* w, const float* alpha, const float* delta,
                                   const float* zero_point, float* galpha,
                                   int64_t n, int64_t inner, int64_t nchan, float qmin, float qmax,
                                   const float* b_dev, float lambda, const float* greg,
                                   int accumulate, void* stream) {
    if (n == 0) return SSQ_OK;
    if (!w || !alpha || !delta || !zero_point || !galpha) return SSQ_ERR_NULL;
    if (int e = check_layout(n, inner, nchan)) return e;
    bool vec = (!gwq || aligned16(gwq)) && aligned16(w) && aligned16(alpha) && aligned16(galpha) && (inner % 4 == 0) &&
               (n / 4 < (int64_t)0x7fffffff);
    const unsigned grid = tile_grid((n + SSQ_ST_TILE - 1) / SSQ_ST_TILE, false);
    ada_bwd_kernel<<<grid, SSQ_THREADS, 0, (cudaStream_t)stream>>>(gwq, w, alpha, delta, zero_point, galpha, n, inner, nchan,
                                                                   qmin, qmax, vec, b_dev, lambda, greg, accumulate);
    return launch_status();
}

extern "C" int ssq_adaround_init_alpha(const float* w, const float* delta, float* alpha,
                                       int64_t n, int64_t inner, int64_t nchan, void* stream) {
    if (n == 0) return SSQ_OK;
    if (!w || !delta || !alpha) return SSQ_ERR_NULL;
    if (int e = check_layout(n, inner, nchan)) return e;
    int grid = grid_for((n + SSQ_THREADS - 1) / SSQ_THREADS);
    ada_init_alpha_kernel<<<grid, SSQ_THREADS, 0, (cudaStream_t)stream>>>(w, delta, alpha, n, inner, nchan);
    return launch_status();
}

extern "C" int ssq_round_reg_fwd(const float* v, int64_t n, const float* b_dev, float lambda,
                                 float* reg_out, void* ws, size_t ws_bytes, void* stream) {
    if (!v || !b_dev || !reg_out) return SSQ_ERR_NULL;
    if (n < 0) return SSQ_ERR_SIZE;
    if (!ws || ws_bytes < ssq_ws_bytes(1)) return SSQ_ERR_WORKSPACE;
    int grid = grid_for((n + SSQ_THREADS * 4 - 1) / (SSQ_THREADS * 4));
    round_reg_fwd_kernel<<<grid, SSQ_THREADS, 0, (cudaStream_t)stream>>>(v, n, b_dev, lambda, reg_out, ws_view(ws, 1));
    return launch_status();
}

extern "C" int ssq_round_reg_bwd(const float* v, int64_t n, const float* b_dev, float lambda,
                                 const float* greg, float* gv, int accumulate, void* stream) {
    if (n == 0) return SSQ_OK;
    if (!v || !b_dev || !gv) return SSQ_ERR_NULL;
    if (n < 0) return SSQ_ERR_SIZE;
    int grid = grid_for((n + SSQ_THREADS * 4 - 1) / (SSQ_THREADS * 4));
    round_reg_bwd_kernel<<<grid, SSQ_THREADS, 0, (cudaStream_t)stream>>>(v, n, b_dev, lambda, greg, gv, accumulate);
    return launch_status();
}

extern "C" int ssq_fq_adaround_fwd_mt(const ssq_adaround_desc* table, int count, int64_t total_tiles,
                                      int soft, const float* b_dev, float lambda, float* reg_out,
                                      void* ws, size_t ws_bytes, void* stream) {
    if (count == 0 || total_tiles == 0) return SSQ_OK;
    if (!table) return SSQ_ERR_NULL;
    if (count < 0 || count > SSQ_MT_MAX || total_tiles < 0) return SSQ_ERR_SIZE;
    if (reg_out && (!ws || ws_bytes < ssq_ws_bytes(1))) return SSQ_ERR_WORKSPACE;
    if (reg_out && !b_dev) return SSQ_ERR_MODE;
    WsView v = ws_view(ws, 1);
    cudaStream_t st = (cudaStream_t)stream;
    MtTable t;
    for (int i = 0; i < count; ++i) t.d[i] = table[i];
    if (soft) ada_fwd_mt_kernel<true><<<tile_grid(total_tiles, reg_out != nullptr), SSQ_THREADS, 0, st>>>(t, count, total_tiles, b_dev, lambda, reg_out, v);
    else ada_fwd_mt_kernel<false><<<tile_grid(total_tiles, false), SSQ_THREADS, 0, st>>>(t, count, total_tiles, b_dev, lambda, nullptr, v);
    return launch_status();
}

extern "C" int ssq_fq_adaround_bwd_mt(const ssq_adaround_desc* table, int count, int64_t total_tiles,
                                      const float* b_dev, float lambda, void* stream) {
    if (count == 0 || total_tiles == 0) return SSQ_OK;
    if (!table) return SSQ_ERR_NULL;
    if (count < 0 || count > SSQ_MT_MAX || total_tiles < 0) return SSQ_ERR_SIZE;
    const unsigned grid = tile_grid(total_tiles, false);
    MtTable t;
    for (int i = 0; i < count; ++i) t.d[i] = table[i];
    ada_bwd_mt_kernel<<<grid, SSQ_THREADS, 0, (cudaStream_t)stream>>>(t, count, total_tiles, b_dev, lambda);
    return launch_status();
}


static int fill_aff(AffTable& a, const ssq_affine_desc* aff, const ssq_adaround_desc* table, int count) {
    for (int i = 0; i < count; ++i) {
        if (!aff[i].gamma || !aff[i].phi || !aff[i].beff) return SSQ_ERR_NULL;
        if (table[i].n % table[i].nchan != 0 || table[i].inner * table[i].nchan != table[i].n) return SSQ_ERR_MODE;   // per-output-channel layers only
        a.d[i] = aff[i];
    }
    return SSQ_OK;
}

extern "C" int ssq_iter_prologue(const ssq_iter_state* st, const float* cache, float* cur_inp, int64_t per_sample,
                                 const ssq_adaround_desc* table, const ssq_affine_desc* aff, int count, int64_t total_tiles,
                                 float lambda, float* reg_out, void* ws, size_t ws_bytes, void* stream) {
    if (!st || !st->step || !st->idx_table) return SSQ_ERR_NULL;
    if (st->n_steps <= 0 || st->batch < 0 || count < 0 || count > SSQ_MT_MAX || total_tiles < 0 || per_sample < 0) return SSQ_ERR_SIZE;
    if ((cache != nullptr) != (cur_inp != nullptr)) return SSQ_ERR_NULL;
    if (count > 0 && !table) return SSQ_ERR_NULL;
    if (reg_out && (!ws || ws_bytes < ssq_ws_bytes(1))) return SSQ_ERR_WORKSPACE;
    if (reg_out && !st->b_table) return SSQ_ERR_MODE;
    IterState S{st->step, st->idx_table, st->idx_live, st->b_table, st->b_live, st->lr_table, st->lr_live, st->n_steps, st->batch};
    const bool gvec = cache && (per_sample % 4 == 0) && aligned16(cache) && aligned16(cur_inp);
    int64_t gt = 0;
    if (cache) {
        const int64_t units = gvec ? (int64_t)st->batch * (per_sample >> 2) : (int64_t)st->batch * per_sample;
        gt = (units + SSQ_THREADS * 4 - 1) / (SSQ_THREADS * 4);
    }
    if (gt > 0x3fffffff) return SSQ_ERR_SIZE;
    const int ada_ctas = count > 0 ? (int)tile_grid(total_tiles, reg_out != nullptr) : 0;
    int64_t grid = gt + ada_ctas;
    if (grid < 1) grid = 1;
    MtTable t;
    for (int i = 0; i < count; ++i) t.d[i] = table[i];
    AffTable a;
    if (aff) { if (int e = fill_aff(a, aff, table, count)) return e; }
    iter_prologue_kernel<<<(unsigned)grid, SSQ_THREADS, 0, (cudaStream_t)stream>>>(S, cache, cur_inp, per_sample, (int)gt, gvec, t, a, aff != nullptr,
                                                                                   count, total_tiles, ada_ctas, lambda, reg_out, ws_view(ws, 1));
    return launch_status();
}

extern "C" int ssq_fq_adaround_bwd_adam_mt(const ssq_adaround_desc* table, const ssq_affine_desc* aff, int count, int64_t total_tiles,
                                           const float* b_dev, float lambda,
                                           float* flat, float* exp_avg, float* exp_avg_sq, const float* lr_dev,
                                           int64_t* step_dev, double beta1, double beta2, double eps, int store_grad, int apply_adam,
                                           void* ws, size_t ws_bytes, void* stream) {
    if (count == 0 || total_tiles == 0) return SSQ_OK;
    if (!table || !flat || !lr_dev || !step_dev || (apply_adam && (!exp_avg || !exp_avg_sq))) return SSQ_ERR_NULL;
    if (!apply_adam && !store_grad) return SSQ_ERR_MODE;
    if (count < 0 || count > SSQ_MT_MAX || total_tiles < 0) return SSQ_ERR_SIZE;
    if (!ws || ws_bytes < ssq_ws_bytes(1)) return SSQ_ERR_WORKSPACE;
    for (int i = 0; i < count; ++i) if (!table[i].gwq || (store_grad && !table[i].galpha)) return SSQ_ERR_NULL;
    MtTable t;
    for (int i = 0; i < count; ++i) t.d[i] = table[i];
    // the last ticket of the fixed header is reserved for the iteration counter's hand-over
    AffTable a;
    if (aff) { if (int e = fill_aff(a, aff, table, count)) return e; }
    AdamArgs A{flat, exp_avg, exp_avg_sq, lr_dev, step_dev, reinterpret_cast<unsigned int*>(ws) + (SSQ_WS_TICKETS - 1),
               beta1, beta2, eps, store_grad, apply_adam};
    const unsigned grid = tile_grid(total_tiles, false);
    ada_bwd_adam_mt_kernel<<<grid, SSQ_THREADS, 0, (cudaStream_t)stream>>>(t, a, aff != nullptr, count, total_tiles, b_dev, lambda, A);
    return launch_status();
}

extern "C" int ssq_affine_grad_mt(const ssq_adaround_desc* table, const ssq_affine_desc* aff, int count, void* stream) {
    if (count == 0) return SSQ_OK;
    if (!table || !aff) return SSQ_ERR_NULL;
    if (count < 0 || count > SSQ_MT_MAX) return SSQ_ERR_SIZE;
    MtTable t; AffTable a;
    for (int i = 0; i < count; ++i) t.d[i] = table[i];
    if (int e = fill_aff(a, aff, table, count)) return e;
    int64_t rows = 0;
    for (int i = 0; i < count; ++i) {
        if (!table[i].gwq || !aff[i].gbeff || !aff[i].ggamma || !aff[i].gphi) return SSQ_ERR_NULL;
        if (aff[i].row_begin != rows) return SSQ_ERR_SIZE;
        rows += table[i].nchan;
    }
    const int wpc = SSQ_THREADS / 32;
    affine_grad_mt_kernel<<<(unsigned)((rows + wpc - 1) / wpc), SSQ_THREADS, 0, (cudaStream_t)stream>>>(t, a, count, rows, nullptr);
    return launch_status();
}
