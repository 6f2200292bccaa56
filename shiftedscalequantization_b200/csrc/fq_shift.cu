// K1c — shifted-scale ChannelQuant: per-input-channel (conv) or per-element (FC) soft/hard mixture
// over S shifted scales, optionally fused with AdaRound soft rounding ('adaShift').
//   reference arithmetic: quant/channelQuant.py:49-127 (forward modes, shifted_x_quant, soft targets),
//   :201-213 (init_v: dequantised candidates), :279-294 (init_v_beta: integer-floor candidates);
//   regularisers quant/layer_recon_shiftedScale.py:386-393, quant/layer_recon_fused_shiftedScale.py:277-282.
// The reference reads S cached weight-sized tensors x_q[i]; here the S candidates are recomputed from the
// weight in registers (12 B/elem instead of (S+2)*4 B/elem).
#include "ssq_common.cuh"
#include "ssq_fastdiv.h"
#include "ssq_slab_plan.h"

namespace ssq {

constexpr int MS = SSQ_MAX_SHIFTS;
__device__ __forceinline__ float clampk(float v, float lo, float hi) { return v < lo ? lo : (v > hi ? hi : v); }

// ---- group probabilities -----------------------------------------------------------------------------
template <int REG>   // -1 none, 0 entropy, 1 pow
__device__ __forceinline__ void probs_one(const float* a, int S, float* sm, float* v, float* p) {
    float m = a[0];
    _Pragma("unroll") for (int i = 1; i < MS; ++i) if (i < S) m = fmaxf(m, a[i]);
    float den = 0.f;
    _Pragma("unroll") for (int i = 0; i < MS; ++i) if (i < S) { sm[i] = expf(a[i] - m); den += sm[i]; }
    _Pragma("unroll") for (int i = 0; i < MS; ++i) if (i < S) {
        sm[i] = div_exact(sm[i], den);
        v[i] = __fadd_rn(__fmul_rn(sm[i], SSQ_STRETCH), SSQ_GAMMA);
        p[i] = fminf(fmaxf(v[i], 0.f), 1.f);
    }
}

__global__ void __launch_bounds__(SSQ_THREADS)
shift_probs_fwd_kernel(const float* __restrict__ alpha, float* __restrict__ p_out, int64_t groups, int S,
                       int reg_mode, const float* __restrict__ b_dev, float lambda, float* __restrict__ reg_out, WsView ws) {
    __shared__ double smem[32];
    // b_dev given: b <= 0 switches the regulariser off in BOTH modes (warm-up gate of the shifted losses,
    // layer_recon_shiftedScale.py:379-380); the entropy mode without b_dev is always on
    const float b = b_dev ? __ldg(b_dev) : 0.f;
    const bool reg_on = reg_out && (b_dev ? (b > 0.f) : (reg_mode == 0));
    double acc[1] = {0.0};
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < groups; g += stride) {
        float a[MS], sm[MS], v[MS], p[MS];
        _Pragma("unroll") for (int i = 0; i < MS; ++i) if (i < S) a[i] = alpha[g * S + i];
        probs_one<0>(a, S, sm, v, p);
        float r = 0.f;
        _Pragma("unroll") for (int i = 0; i < MS; ++i) if (i < S) {
            p_out[g * S + i] = p[i];
            if (reg_on) r += (reg_mode == 0) ? -(p[i] * logf(p[i] + 1e-10f)) : reg_term(p[i], b);
        }
        acc[0] += (double)r;
    }
    if (!reg_out) return;
    block_sum<1>(acc, smem);
    if (grid_finish<1>(acc, ws, 0, blockIdx.x, gridDim.x, smem) && threadIdx.x == 0)
        reg_out[0] = reg_on ? (float)((double)lambda * acc[0]) : 0.f;
}

__global__ void __launch_bounds__(SSQ_THREADS)
shift_probs_bwd_kernel(const float* __restrict__ alpha, const float* __restrict__ gp, float* __restrict__ galpha,
                       int64_t groups, int S, int reg_mode, const float* __restrict__ b_dev, float lambda,
                       const float* __restrict__ greg) {
    const float b = b_dev ? __ldg(b_dev) : 0.f;
    const bool reg_on = (reg_mode >= 0) && (b_dev ? (b > 0.f) : (reg_mode == 0));
    const float lam_g = reg_on ? lambda * (greg ? __ldg(greg) : 1.f) : 0.f;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < groups; g += stride) {
        float a[MS], sm[MS], v[MS], p[MS], gs[MS];
        _Pragma("unroll") for (int i = 0; i < MS; ++i) if (i < S) a[i] = alpha[g * S + i];
        probs_one<0>(a, S, sm, v, p);
        float dot = 0.f;
        _Pragma("unroll") for (int i = 0; i < MS; ++i) if (i < S) {
            float gpi = gp ? gp[g * S + i] : 0.f;
            if (reg_on) {
                float dr = (reg_mode == 0) ? -(logf(p[i] + 1e-10f) + div_exact(p[i], p[i] + 1e-10f)) : reg_term_grad(p[i], b);
                gpi += lam_g * dr;
            }
            float gv = (v[i] >= 0.f && v[i] <= 1.f) ? gpi : 0.f;   // clamp backward, bounds inclusive
            gs[i] = gv * SSQ_STRETCH;                                // d(sm*1.2-0.1)/dsm
            dot += gs[i] * sm[i];
        }
        _Pragma("unroll") for (int i = 0; i < MS; ++i) if (i < S) galpha[g * S + i] = sm[i] * (gs[i] - dot);   // softmax backward
    }
}

// ---- per-element candidate evaluation -----------------------------------------------------------------
struct ShiftCtx {
    float ds[MS];    // delta*s_i for this row
    float d, z;      // delta, zero point for this row
};

// returns y; optionally the S mixture terms (dy/dp_i) and the adaShift inside flag
template <int MODE>
__device__ __forceinline__ float shift_one(float w, const ShiftCtx& c, const float* p, float beta, int S,
                                           int hard_targets, int hard_round, float qmin, float qmax,
                                           float* terms, bool& inside, float& dh) {
    float t[MS];
    _Pragma("unroll") for (int i = 0; i < MS; ++i) if (i < S) {
        float u = div_exact(w, c.ds[i]);
        if (MODE == SSQ_SHIFT_DEQUANT) {
            float q = clampk(__fadd_rn(rintf(u), c.z), qmin, qmax);
            t[i] = __fmul_rn(__fsub_rn(q, c.z), c.ds[i]);
        } else {
            t[i] = floorf(u);
        }
    }
    float mix;
    if (hard_targets) {
        float pbest = p[0];                       // torch.argmax: first maximum wins
        mix = t[0];
        _Pragma("unroll") for (int i = 1; i < MS; ++i) if (i < S) if (p[i] > pbest) { pbest = p[i]; mix = t[i]; }
    } else {
        mix = __fmul_rn(t[0], p[0]);
        _Pragma("unroll") for (int i = 1; i < MS; ++i) if (i < S) mix = __fadd_rn(mix, __fmul_rn(t[i], p[i]));
    }
    if (terms) _Pragma("unroll") for (int i = 0; i < MS; ++i) if (i < S) terms[i] = t[i];
    inside = true; dh = 0.f;
    if (MODE == SSQ_SHIFT_DEQUANT) return mix;
    float r;
    if (hard_round) r = (beta >= 0.f) ? 1.f : 0.f;
    else { float h; dh = rect_sigmoid_grad(beta, h); r = h; }
    float xi = __fadd_rn(__fadd_rn(mix, r), c.z);
    inside = (xi >= qmin) && (xi <= qmax);
    float q = clampk(xi, qmin, qmax);
    return __fmul_rn(__fsub_rn(q, c.z), c.d);
}

template <int MODE>
__global__ void __launch_bounds__(SSQ_THREADS)
fq_shift_fwd_kernel(const float* __restrict__ w, const float* __restrict__ shift_delta, const float* __restrict__ delta,
                    const float* __restrict__ zp, const float* __restrict__ p, const float* __restrict__ beta,
                    float* __restrict__ y, int64_t oc, int64_t K, int64_t kk, int S, int per_element,
                    int hard_targets, int hard_round, float qmin, float qmax) {
    const int64_t n = oc * K;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += stride) {
        int64_t r = e / K;
        int64_t k = e - r * K;
        int64_t g = per_element ? e : (k / kk);
        ShiftCtx c;
        _Pragma("unroll") for (int i = 0; i < MS; ++i) if (i < S) c.ds[i] = __ldg(shift_delta + (int64_t)i * oc + r);
        c.d = __ldg(delta + r); c.z = __ldg(zp + r);
        float pv[MS];
        _Pragma("unroll") for (int i = 0; i < MS; ++i) if (i < S) pv[i] = __ldg(p + g * S + i);
        bool inside; float dh;
        float bv = (MODE == SSQ_SHIFT_ADASHIFT) ? beta[e] : 0.f;
        y[e] = shift_one<MODE>(w[e], c, pv, bv, S, hard_targets, hard_round, qmin, qmax, nullptr, inside, dh);
    }
}

// backward: CTA = 256 columns x a slab of rows; each thread owns one column and walks the slab's rows,
// accumulating the S per-column sums in registers -> partial[slab][K][S]. (per_element: direct write.)
template <int MODE>
__global__ void __launch_bounds__(SSQ_THREADS)
fq_shift_bwd_kernel(const float* __restrict__ gy, const float* __restrict__ w, const float* __restrict__ shift_delta,
                    const float* __restrict__ delta, const float* __restrict__ zp, const float* __restrict__ p,
                    const float* __restrict__ beta, float* __restrict__ gp, float* __restrict__ gbeta,
                    float* __restrict__ partial, int64_t oc, int64_t K, int64_t kk, int S, int per_element,
                    int hard_round, float qmin, float qmax, int64_t rows_per_slab) {
    const int64_t col = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (col >= K) return;
    const int64_t r0 = (int64_t)blockIdx.y * rows_per_slab;
    const int64_t r1 = r0 + rows_per_slab < oc ? r0 + rows_per_slab : oc;
    float acc[MS];
    for (int i = 0; i < MS; ++i) acc[i] = 0.f;
    float pv[MS];
    if (!per_element) _Pragma("unroll") for (int i = 0; i < MS; ++i) if (i < S) pv[i] = __ldg(p + (col / kk) * S + i);
    for (int64_t r = r0; r < r1; ++r) {
        const int64_t e = r * K + col;
        ShiftCtx c;
        _Pragma("unroll") for (int i = 0; i < MS; ++i) if (i < S) c.ds[i] = __ldg(shift_delta + (int64_t)i * oc + r);
        c.d = __ldg(delta + r); c.z = __ldg(zp + r);
        if (per_element) _Pragma("unroll") for (int i = 0; i < MS; ++i) if (i < S) pv[i] = __ldg(p + e * S + i);
        float terms[MS]; bool inside; float dh;
        float bv = (MODE == SSQ_SHIFT_ADASHIFT) ? beta[e] : 0.f;
        (void)shift_one<MODE>(w[e], c, pv, bv, S, 0, hard_round, qmin, qmax, terms, inside, dh);
        const float g = gy[e];
        float gm = g;                                   // gradient wrt the mixture
        if (MODE == SSQ_SHIFT_ADASHIFT) {
            gm = inside ? g * c.d : 0.f;
            if (gbeta) gbeta[e] = hard_round ? 0.f : gm * dh;
        }
        if (per_element) { _Pragma("unroll") for (int i = 0; i < MS; ++i) if (i < S) gp[e * S + i] = gm * terms[i]; }
        else { _Pragma("unroll") for (int i = 0; i < MS; ++i) if (i < S) acc[i] += gm * terms[i]; }
    }
    if (!per_element)
        _Pragma("unroll") for (int i = 0; i < MS; ++i) if (i < S) partial[((int64_t)blockIdx.y * K + col) * S + i] = acc[i];
}

// ---- vector kernels (conv layers: K % 4 == 0, 16-byte aligned, one group per input channel) ---------------------------
// Same arithmetic as shift_one, organised so that the per-thread fixed cost is small against 16 elements of work:
//  * row / group indices come from host-computed multiply-shift constants (FastDiv) instead of emulated 32-bit divisions;
//  * a row's S reciprocals (MUFU.RCP + Newton) and clamp bounds are computed once per row change (forward) or once per CTA
//    row chunk into shared memory (backward);
//  * ONE range test per float4 decides between the fast path (hoisted-reciprocal exact quotients, FMNMX clamps — every value
//    is finite there) and shift_*_slow, an out-of-line restatement with div.rn and NaN-preserving clamps;
//  * dequantised candidates use clamp(rint(u) + z, qmin, qmax) - z == clamp(rint(u), qmin - z, qmax - z), exact whenever z, qmin,
//    qmax are integers below 2^22 (checked per row; anything else takes the slow path). fma(k, ds, +0) keeps the reference's +0
//    where rint(u) is -0.
struct ShiftArgs {
    const float *w, *shift_delta, *delta, *zp, *p, *beta, *gy;
    float *y, *gbeta, *partial;
    int64_t oc, rows_per_slab;
    uint32_t K4, kk, total4, last_group;
    FastDiv dK4, dkk;
    int S, hard_targets, hard_round, q_int;        // q_int: qmin, qmax are integers of magnitude <= 2^22
    float qmin, qmax;
};

// out-of-line slow paths: one float4 of the forward / one float4 of one row of the backward, scalar reference arithmetic
template <int MODE>
__device__ __noinline__ void shift_fwd_slow(const ShiftArgs& a, uint32_t i) {
    const int64_t K = (int64_t)a.K4 * 4;
    const int64_t r = i / a.K4;
    ShiftCtx c;
    for (int s = 0; s < MS; ++s) if (s < a.S) c.ds[s] = __ldg(a.shift_delta + (int64_t)s * a.oc + r);
    c.d = __ldg(a.delta + r); c.z = __ldg(a.zp + r);
    for (int e = 0; e < 4; ++e) {
        const int64_t idx = (int64_t)i * 4 + e, k = idx - r * K, g = k / a.kk;
        float pv[MS];
        for (int s = 0; s < MS; ++s) if (s < a.S) pv[s] = __ldg(a.p + g * a.S + s);
        bool inside; float dh;
        const float bv = (MODE == SSQ_SHIFT_ADASHIFT) ? a.beta[idx] : 0.f;
        a.y[idx] = shift_one<MODE>(a.w[idx], c, pv, bv, a.S, a.hard_targets, a.hard_round, a.qmin, a.qmax, nullptr, inside, dh);
    }
}
struct SlowTerms { float gm_t[4][MS]; };            // gm * t_i of the four columns
template <int MODE>
__device__ __noinline__ SlowTerms shift_bwd_slow(const ShiftArgs& a, int64_t r, uint32_t col4) {
    SlowTerms o;
    const int64_t K = (int64_t)a.K4 * 4;
    ShiftCtx c;
    for (int s = 0; s < MS; ++s) if (s < a.S) c.ds[s] = __ldg(a.shift_delta + (int64_t)s * a.oc + r);
    c.d = __ldg(a.delta + r); c.z = __ldg(a.zp + r);
    for (int e = 0; e < 4; ++e) {
        const int64_t k = (int64_t)col4 * 4 + e, idx = r * K + k, g = k / a.kk;
        float pv[MS], terms[MS];
        for (int s = 0; s < MS; ++s) { pv[s] = (s < a.S) ? __ldg(a.p + g * a.S + s) : 0.f; terms[s] = 0.f; }
        bool inside; float dh;
        const float bv = (MODE == SSQ_SHIFT_ADASHIFT) ? a.beta[idx] : 0.f;
        (void)shift_one<MODE>(a.w[idx], c, pv, bv, a.S, 0, a.hard_round, a.qmin, a.qmax, terms, inside, dh);
        float gm = a.gy[idx];
        if (MODE == SSQ_SHIFT_ADASHIFT) {
            gm = inside ? gm * c.d : 0.f;
            if (a.gbeta) a.gbeta[idx] = a.hard_round ? 0.f : gm * dh;
        }
        for (int s = 0; s < MS; ++s) o.gm_t[e][s] = gm * terms[s];
    }
    return o;
}

// per-row constants of the fast path
template <int S> struct RowP { float ds[S], r[S], c0, c1; bool ok; };   // (c0, c1) = (qmin - z, qmax - z) dequant | (d, z) adaShift
template <int MODE, int S>
__device__ __forceinline__ void row_params(RowP<S>& R, const ShiftArgs& a, int64_t row) {
    bool ok = true;
#pragma unroll
    for (int s = 0; s < S; ++s) {
        const float ds = __ldg(a.shift_delta + (int64_t)s * a.oc + row);
        const float r0 = rcp_approx(ds);
        R.ds[s] = ds;
        R.r[s] = fmaf(r0, fmaf(-ds, r0, 1.0f), r0);
        ok = ok && (ds >= 0x1p-60f) && (ds <= 0x1p60f);
    }
    const float d = __ldg(a.delta + row), z = __ldg(a.zp + row);
    if (MODE == SSQ_SHIFT_DEQUANT) {
        R.c0 = __fsub_rn(a.qmin, z); R.c1 = __fsub_rn(a.qmax, z);
        ok = ok && a.q_int && (fabsf(z) <= 0x1p22f) && (rintf(z) == z);
    } else {
        R.c0 = d; R.c1 = z;
        ok = ok && (fabsf(z) <= 0x1p60f) && (fabsf(d) <= 0x1p60f);
    }
    R.ok = ok;
}
__device__ __forceinline__ float quot_fast(float x, float ds, float r) {
    const float q0 = __fmul_rn(x, r);
    return fmaf(r, fmaf(-ds, q0, x), q0);
}
// fast-path candidates of one element: dequantised (q - z) * ds, or the integer floors
template <int MODE, int S>
__device__ __forceinline__ void candidates(float x, const RowP<S>& R, float (&t)[S]) {
#pragma unroll
    for (int s = 0; s < S; ++s) {
        const float u = quot_fast(x, R.ds[s], R.r[s]);
        if (MODE == SSQ_SHIFT_DEQUANT) t[s] = fmaf(fminf(fmaxf(rintf(u), R.c0), R.c1), R.ds[s], 0.0f);
        else t[s] = floorf(u);
    }
}
template <int S>
__device__ __forceinline__ float mixture(const float (&t)[S], const float (&p)[S], int hard_targets) {
    float mix;
    if (hard_targets) {
        float pbest = p[0];                       // torch.argmax: first maximum wins
        mix = t[0];
#pragma unroll
        for (int s = 1; s < S; ++s) if (p[s] > pbest) { pbest = p[s]; mix = t[s]; }
    } else {
        mix = __fmul_rn(t[0], p[0]);
#pragma unroll
        for (int s = 1; s < S; ++s) mix = __fadd_rn(mix, __fmul_rn(t[s], p[s]));
    }
    return mix;
}

// group probabilities of the four elements of a float4 starting at in-row offset k0
//   GK 1: kk >= 4 — at most two groups per float4, both rows loaded up front, selected per element
//   GK 2: kk == 1 — the 4*S probabilities are contiguous: S float4 loads
//   GK 0: anything else — per-element lookup
template <int S, int GK>
__device__ __forceinline__ void group_probs(const ShiftArgs& a, uint32_t k0, float (&pe)[4][S]) {
    if (GK == 1) {
        const uint32_t g = fastdiv(k0, a.dkk), n0 = a.kk - (k0 - g * a.kk);   // n0 elements still belong to group g
        const uint32_t g1 = g < a.last_group ? g + 1 : g;
        float p0[S], p1[S];
#pragma unroll
        for (int s = 0; s < S; ++s) { p0[s] = __ldg(a.p + (size_t)g * S + s); p1[s] = __ldg(a.p + (size_t)g1 * S + s); }
#pragma unroll
        for (int e = 0; e < 4; ++e)
#pragma unroll
            for (int s = 0; s < S; ++s) pe[e][s] = (e == 0 || (uint32_t)e < n0) ? p0[s] : p1[s];
    } else if (GK == 2) {
        float flat[4 * S];
        const float4* src = reinterpret_cast<const float4*>(a.p + (size_t)k0 * S);
#pragma unroll
        for (int j = 0; j < S; ++j) { const float4 v = __ldg(src + j); flat[4 * j] = v.x; flat[4 * j + 1] = v.y; flat[4 * j + 2] = v.z; flat[4 * j + 3] = v.w; }
#pragma unroll
        for (int e = 0; e < 4; ++e)
#pragma unroll
            for (int s = 0; s < S; ++s) pe[e][s] = flat[e * S + s];
    } else {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const uint32_t g = a.kk == 1 ? k0 + e : fastdiv(k0 + e, a.dkk);
#pragma unroll
            for (int s = 0; s < S; ++s) pe[e][s] = __ldg(a.p + (size_t)g * S + s);
        }
    }
}

// forward: address-ordered tiles of 256*U float4s (ssq_common.cuh); row = output channel, group = k / kk.
// SOFT: soft targets and soft round known at compile time (the training loops) — the per-element mode branches vanish.
#ifndef SSQ_K1C_FWD_U
#define SSQ_K1C_FWD_U 4
#endif
#ifndef SSQ_K1C_FWD_CTAS
#define SSQ_K1C_FWD_CTAS 4
#endif

template <int MODE, int S, bool SOFT, int GK>
__global__ void __launch_bounds__(SSQ_THREADS, SSQ_K1C_FWD_CTAS)
fq_shift_fwd_vec(const __grid_constant__ ShiftArgs a) {
    constexpr int U = SSQ_K1C_FWD_U;
    const int hard_targets = SOFT ? 0 : a.hard_targets, hard_round = SOFT ? 0 : a.hard_round;
    const uint32_t i0 = blockIdx.x * (uint32_t)(SSQ_THREADS * U) + threadIdx.x;
    float4 wv[U], bv[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
        const uint32_t i = i0 + u * SSQ_THREADS;
        wv[u] = bv[u] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (i < a.total4) {
            wv[u] = ld_stream4(a.w + (size_t)i * 4);
            if (MODE == SSQ_SHIFT_ADASHIFT) bv[u] = ld_stream4(a.beta + (size_t)i * 4);
        }
    }
    RowP<S> R;
    uint32_t r_have = 0xffffffffu, slow_mask = 0;      // slow vectors are redone after the loop: no call inside it
#pragma unroll
    for (int u = 0; u < U; ++u) {
        const uint32_t i = i0 + u * SSQ_THREADS;
        if (i >= a.total4) break;
        const uint32_t row = fastdiv(i, a.dK4), k0 = (i - row * a.K4) * 4;
        if (row != r_have) { row_params<MODE, S>(R, a, row); r_have = row; }
        if (!(R.ok && small4(wv[u]))) { slow_mask |= 1u << u; continue; }
        float pe[4][S];
        group_probs<S, GK>(a, k0, pe);
        const float xe[4] = {wv[u].x, wv[u].y, wv[u].z, wv[u].w}, be[4] = {bv[u].x, bv[u].y, bv[u].z, bv[u].w};
        float out[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            float t[S];
            candidates<MODE, S>(xe[e], R, t);
            const float mix = mixture<S>(t, pe[e], hard_targets);
            if (MODE == SSQ_SHIFT_DEQUANT) out[e] = mix;
            else {
                const float rr = hard_round ? ((be[e] >= 0.f) ? 1.f : 0.f) : rect_sigmoid(be[e]);
                const float xi = __fadd_rn(__fadd_rn(mix, rr), R.c1);
                out[e] = __fmul_rn(__fsub_rn(fminf(fmaxf(xi, a.qmin), a.qmax), R.c1), R.c0);
            }
        }
        st_stream4(a.y + (size_t)i * 4, make_float4(out[0], out[1], out[2], out[3]));
    }
    if (slow_mask) {
        for (int u = 0; u < U; ++u)
            if ((slow_mask >> u) & 1u) shift_fwd_slow<MODE>(a, i0 + u * SSQ_THREADS);
    }
}

// backward: CTA = 1024 adjacent columns (one float4 per thread) x a slab of rows, NR rows in flight per thread; per-column
// sums of d y / d p[g, i] stay in registers -> partial[slab][K][S] (the finish kernel adds slabs and the kk columns of a group in
// fp64). The slab's row constants are staged in shared memory SSQ_SHIFT_RB rows at a time (one thread per row computes them).
// Measured on B200 (scratch/k1c_probe.py): dequantised mixture 4 CTAs/SM x 3 rows in flight (64 registers), adaShift 2 CTAs/SM
// x 4 rows (121 registers, no spills); the slab count makes the grid exactly one wave at that occupancy (slab_plan). A
// TMA-bulk-copy ring (cp.async.bulk + mbarrier, 3-5 stages) was tried instead of register prefetch: 0.79 / 0.66 of peak
// against 0.92 / 0.84 — the per-row barrier polling costs more issue slots than the register pipeline saves.
#define SSQ_SHIFT_RB 64
#ifndef SSQ_K1C_BWD_CTAS_ADA
#define SSQ_K1C_BWD_CTAS_ADA 2
#endif
#ifndef SSQ_K1C_BWD_CTAS_DEQ
#define SSQ_K1C_BWD_CTAS_DEQ 4
#endif
#ifndef SSQ_K1C_BWD_ROWS_ADA
#define SSQ_K1C_BWD_ROWS_ADA 4
#endif
#ifndef SSQ_K1C_BWD_ROWS_DEQ
#define SSQ_K1C_BWD_ROWS_DEQ 3
#endif
#define SSQ_K1C_BWD_CTAS(MODE) ((MODE) == SSQ_SHIFT_ADASHIFT ? SSQ_K1C_BWD_CTAS_ADA : SSQ_K1C_BWD_CTAS_DEQ)
template <int MODE, int S, bool SOFT, int GK>
__global__ void __launch_bounds__(SSQ_THREADS, SSQ_K1C_BWD_CTAS(MODE))
fq_shift_bwd_vec(const __grid_constant__ ShiftArgs a) {
    constexpr int NQ = (2 * S + 2 + 3) / 4;        // float4s per row: ds[S], r[S], c0, c1
    constexpr int NR = (MODE == SSQ_SHIFT_ADASHIFT) ? SSQ_K1C_BWD_ROWS_ADA : SSQ_K1C_BWD_ROWS_DEQ;
    __shared__ float4 srow[SSQ_SHIFT_RB][NQ];
    __shared__ uint32_t sok[SSQ_SHIFT_RB];
    const uint32_t col4 = blockIdx.x * SSQ_THREADS + threadIdx.x;
    const bool live = col4 < a.K4;
    const int hard_round = SOFT ? 0 : a.hard_round;
    const int64_t K = (int64_t)a.K4 * 4;
    const int64_t r0 = (int64_t)blockIdx.y * a.rows_per_slab;
    const int64_t r1 = r0 + a.rows_per_slab < a.oc ? r0 + a.rows_per_slab : a.oc;
    const int nrows = (int)(r1 - r0);
    float pe[4][S], acc[4][S];
#pragma unroll
    for (int e = 0; e < 4; ++e)
#pragma unroll
        for (int s = 0; s < S; ++s) { pe[e][s] = 0.f; acc[e][s] = 0.f; }
    if (live && MODE == SSQ_SHIFT_ADASHIFT) group_probs<S, GK>(a, col4 * 4, pe);
    uint64_t slow_rows = 0;
    // fast-path test of one row's vectors. It reads one lane of EVERY loaded vector (x * 0 keeps NaN / Inf and cannot be folded
    // away), so all loads of a row pair are consumed before the first branch: without that ptxas sinks the second row's loads
    // below the first row's branch and the two-rows-in-flight pipeline degenerates (0.83 -> 0.74 of peak).
    auto fast_ok = [&](int lr, const float4& g4, const float4& w4, const float4& b4) {
        float sum = (fabsf(w4.x) + fabsf(w4.y)) + (fabsf(w4.z) + fabsf(w4.w));
        sum = fmaf(g4.x, 0.0f, sum);
        if (MODE == SSQ_SHIFT_ADASHIFT) sum = fmaf(b4.x, 0.0f, sum);
        return (sok[lr] != 0u) && (sum < SSQ_DIV_XMAX);
    };
    auto row = [&](int64_t r, int lr, const float4& g4, const float4& w4, const float4& b4, bool fast) {
        if (!fast) { slow_rows |= 1ull << lr; return; }   // redone at the end of the chunk: no call in the row loop
        RowP<S> R;
        {
            float f[4 * NQ];
#pragma unroll
            for (int j = 0; j < NQ; ++j) { const float4 v = srow[lr][j]; f[4 * j] = v.x; f[4 * j + 1] = v.y; f[4 * j + 2] = v.z; f[4 * j + 3] = v.w; }
#pragma unroll
            for (int s = 0; s < S; ++s) { R.ds[s] = f[s]; R.r[s] = f[S + s]; }
            R.c0 = f[2 * S]; R.c1 = f[2 * S + 1];
        }
        const float xe[4] = {w4.x, w4.y, w4.z, w4.w}, ge[4] = {g4.x, g4.y, g4.z, g4.w}, be[4] = {b4.x, b4.y, b4.z, b4.w};
        float gb[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            float t[S];
            candidates<MODE, S>(xe[e], R, t);
            float gm = ge[e];                       // gradient wrt the mixture
            gb[e] = 0.f;
            if (MODE == SSQ_SHIFT_ADASHIFT) {
                const float mix = mixture<S>(t, pe[e], 0);
                float rr, dh = 0.f;
                if (hard_round) rr = (be[e] >= 0.f) ? 1.f : 0.f;
                else dh = rect_sigmoid_grad(be[e], rr);
                const float xi = __fadd_rn(__fadd_rn(mix, rr), R.c1);
                const bool inside = (xi >= a.qmin) && (xi <= a.qmax);
                gm = inside ? ge[e] * R.c0 : 0.f;
                gb[e] = hard_round ? 0.f : gm * dh;
            }
#pragma unroll
            for (int s = 0; s < S; ++s) acc[e][s] += gm * t[s];
        }
        if (MODE == SSQ_SHIFT_ADASHIFT && a.gbeta) st_stream4(a.gbeta + r * K + (int64_t)col4 * 4, make_float4(gb[0], gb[1], gb[2], gb[3]));
    };
    const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int c0 = 0; c0 < nrows; c0 += SSQ_SHIFT_RB) {
        const int nrow = nrows - c0 < SSQ_SHIFT_RB ? nrows - c0 : SSQ_SHIFT_RB;
        __syncthreads();
        if ((int)threadIdx.x < nrow) {
            RowP<S> R;
            row_params<MODE, S>(R, a, r0 + c0 + threadIdx.x);
            float f[4 * NQ];
#pragma unroll
            for (int j = 0; j < 4 * NQ; ++j) f[j] = 0.f;
#pragma unroll
            for (int s = 0; s < S; ++s) { f[s] = R.ds[s]; f[S + s] = R.r[s]; }
            f[2 * S] = R.c0; f[2 * S + 1] = R.c1;
            sok[threadIdx.x] = R.ok ? 1u : 0u;
#pragma unroll
            for (int j = 0; j < NQ; ++j) srow[threadIdx.x][j] = make_float4(f[4 * j], f[4 * j + 1], f[4 * j + 2], f[4 * j + 3]);
        }
        __syncthreads();
        if (!live) continue;
        int lr = 0;
        for (; lr + NR <= nrow; lr += NR) {            // NR rows in flight per thread
            float4 gq[NR], wq[NR], bq[NR];
            bool fq[NR];
#pragma unroll
            for (int d = 0; d < NR; ++d) {
                const int64_t ea = (r0 + c0 + lr + d) * K + (int64_t)col4 * 4;
                gq[d] = ld_stream4(a.gy + ea); wq[d] = ld_stream4(a.w + ea);
                bq[d] = (MODE == SSQ_SHIFT_ADASHIFT) ? ld_stream4(a.beta + ea) : zero4;
            }
#pragma unroll
            for (int d = 0; d < NR; ++d) fq[d] = fast_ok(lr + d, gq[d], wq[d], bq[d]);
#pragma unroll
            for (int d = 0; d < NR; ++d) row(r0 + c0 + lr + d, lr + d, gq[d], wq[d], bq[d], fq[d]);
        }
        for (; lr < nrow; ++lr) {
            const int64_t r = r0 + c0 + lr, ea = r * K + (int64_t)col4 * 4;
            const float4 ga = ld_stream4(a.gy + ea), wa = ld_stream4(a.w + ea), ba = (MODE == SSQ_SHIFT_ADASHIFT) ? ld_stream4(a.beta + ea) : zero4;
            row(r, lr, ga, wa, ba, fast_ok(lr, ga, wa, ba));
        }
        while (slow_rows) {
            const int sl = __ffsll((long long)slow_rows) - 1;
            slow_rows &= slow_rows - 1;
            const SlowTerms o = shift_bwd_slow<MODE>(a, r0 + c0 + sl, col4);
#pragma unroll
            for (int e = 0; e < 4; ++e)
#pragma unroll
                for (int s = 0; s < S; ++s) acc[e][s] += o.gm_t[e][s];
        }
    }
    if (!live) return;
#pragma unroll
    for (int e = 0; e < 4; ++e)
#pragma unroll
        for (int s = 0; s < S; ++s)
            a.partial[((int64_t)blockIdx.y * K + (int64_t)col4 * 4 + e) * S + s] = acc[e][s];
}

// one warp per (group, shift): lanes stride over the nslab * kk partials, fixed shuffle tree in fp64 (deterministic)
__global__ void __launch_bounds__(SSQ_THREADS)
fq_shift_bwd_finish_kernel(const float* __restrict__ partial, float* __restrict__ gp, int64_t ic, int64_t K, int64_t kk,
                           int S, int nslab) {
    const int64_t t = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (t >= ic * S) return;                          // whole warps leave together
    const int lane = threadIdx.x & 31;
    const int64_t g = t / S; const int i = (int)(t - g * S);
    const int64_t items = (int64_t)nslab * kk;
    double s = 0.0;
    for (int64_t j = lane; j < items; j += 32) {
        const int64_t sl = j / kk, k = g * kk + (j - sl * kk);
        s += (double)partial[(sl * K + k) * S + i];
    }
    s = warp_sum(s);
    if (lane == 0) gp[t] = (float)s;
}

static inline unsigned finish_grid(int64_t ic, int nshift) {
    return (unsigned)((ic * nshift * 32 + SSQ_THREADS - 1) / SSQ_THREADS);        // one warp per (group, shift)
}
static inline int group_kind(int64_t kk, const float* p) { return kk >= 4 ? 1 : (kk == 1 && aligned16(p) ? 2 : 0); }
static inline ShiftArgs shift_args(const float* w, const float* shift_delta, const float* delta, const float* zp, const float* p,
                                   const float* beta, const float* gy, float* y, float* gbeta, float* partial,
                                   int64_t oc, int64_t K, int64_t kk, int nshift, int hard_targets, int hard_round,
                                   float qmin, float qmax, int64_t rows_per_slab) {
    ShiftArgs a;
    a.w = w; a.shift_delta = shift_delta; a.delta = delta; a.zp = zp; a.p = p; a.beta = beta; a.gy = gy;
    a.y = y; a.gbeta = gbeta; a.partial = partial;
    a.oc = oc; a.rows_per_slab = rows_per_slab;
    a.K4 = (uint32_t)(K / 4); a.kk = (uint32_t)kk; a.total4 = (uint32_t)(oc * (K / 4)); a.last_group = (uint32_t)(K / kk - 1);
    a.dK4 = make_fastdiv(a.K4); a.dkk = make_fastdiv(a.kk >= 2 ? a.kk : 2);   // kk == 1 never divides (GK 2, or k0 + e itself)
    a.S = nshift; a.hard_targets = hard_targets; a.hard_round = hard_round;
    a.q_int = (qmin == (float)(int64_t)qmin) && (qmax == (float)(int64_t)qmax) && qmin >= -0x1p22f && qmin <= 0x1p22f &&
              qmax >= -0x1p22f && qmax <= 0x1p22f;
    a.qmin = qmin; a.qmax = qmax;
    return a;
}

}  // namespace ssq

using namespace ssq;

extern "C" int ssq_shift_probs_fwd(const float* alpha, float* p, int64_t groups, int nshift,
                                   int reg_mode, const float* b_dev, float lambda, float* reg_out,
                                   void* ws, size_t ws_bytes, void* stream) {
    if (groups == 0) return SSQ_OK;
    if (!alpha || !p) return SSQ_ERR_NULL;
    if (groups < 0 || nshift < 1 || nshift > MS) return SSQ_ERR_SIZE;
    if (reg_out && (reg_mode < 0 || reg_mode > 1)) return SSQ_ERR_MODE;
    if (reg_out && reg_mode == 1 && !b_dev) return SSQ_ERR_NULL;
    if (reg_out && (!ws || ws_bytes < ssq_ws_bytes(1))) return SSQ_ERR_WORKSPACE;
    int grid = grid_for((groups + SSQ_THREADS - 1) / SSQ_THREADS);
    shift_probs_fwd_kernel<<<grid, SSQ_THREADS, 0, (cudaStream_t)stream>>>(alpha, p, groups, nshift, reg_mode, b_dev, lambda, reg_out, ws_view(ws, 1));
    return launch_status();
}

extern "C" int ssq_shift_probs_bwd(const float* alpha, const float* gp, float* galpha, int64_t groups, int nshift,
                                   int reg_mode, const float* b_dev, float lambda, const float* greg, void* stream) {
    if (groups == 0) return SSQ_OK;
    if (!alpha || !galpha) return SSQ_ERR_NULL;
    if (groups < 0 || nshift < 1 || nshift > MS) return SSQ_ERR_SIZE;
    if (reg_mode < -1 || reg_mode > 1) return SSQ_ERR_MODE;
    if (reg_mode == 1 && !b_dev) return SSQ_ERR_NULL;
    int grid = grid_for((groups + SSQ_THREADS - 1) / SSQ_THREADS);
    shift_probs_bwd_kernel<<<grid, SSQ_THREADS, 0, (cudaStream_t)stream>>>(alpha, gp, galpha, groups, nshift, reg_mode, b_dev, lambda, greg);
    return launch_status();
}

extern "C" int ssq_fq_shift_fwd(const float* w, const float* shift_delta, const float* delta, const float* zero_point,
                                const float* p, const float* beta, float* y,
                                int64_t oc, int64_t ic, int64_t kk, int nshift, int per_element,
                                int mode, int hard_targets, int hard_round, float qmin, float qmax, void* stream) {
    if (oc == 0 || ic == 0 || kk == 0) return SSQ_OK;
    if (!w || !shift_delta || !delta || !zero_point || !p || !y) return SSQ_ERR_NULL;
    if (oc < 0 || ic < 0 || kk < 0 || nshift < 1 || nshift > MS) return SSQ_ERR_SIZE;
    if (mode == SSQ_SHIFT_ADASHIFT && !beta) return SSQ_ERR_NULL;
    const int64_t K = ic * kk, n = oc * K;
    int grid = grid_for((n + SSQ_THREADS - 1) / SSQ_THREADS);
    cudaStream_t st = (cudaStream_t)stream;
    if (mode != SSQ_SHIFT_DEQUANT && mode != SSQ_SHIFT_ADASHIFT) return SSQ_ERR_MODE;
    const bool vec = !per_element && (K % 4 == 0) && K >= 8 && n / 4 < 0x7fffffff && aligned16(w) && aligned16(y) && (!beta || aligned16(beta));
    if (vec) {
        ShiftArgs a = shift_args(w, shift_delta, delta, zero_point, p, beta, nullptr, y, nullptr, nullptr, oc, K, kk, nshift,
                                 hard_targets, hard_round, qmin, qmax, 0);
        const unsigned vgrid = tile_grid((n / 4 + SSQ_THREADS * SSQ_K1C_FWD_U - 1) / (SSQ_THREADS * SSQ_K1C_FWD_U), false);
        const bool soft = !hard_targets && !hard_round;
        const int gk = group_kind(kk, p);
#define FWDK(M, SS, SO, GK) fq_shift_fwd_vec<M, SS, SO, GK><<<vgrid, SSQ_THREADS, 0, st>>>(a)
#define FWDG(M, SS, SO) do { if (gk == 1) FWDK(M, SS, SO, 1); else if (gk == 2) FWDK(M, SS, SO, 2); else FWDK(M, SS, SO, 0); } while (0)
#define FWDV(M, SS) do { if (soft) FWDG(M, SS, true); else FWDG(M, SS, false); } while (0)
#define FWDS(M) switch (nshift) { case 1: FWDV(M, 1); break; case 2: FWDV(M, 2); break; case 3: FWDV(M, 3); break; default: FWDV(M, 4); }
        if (mode == SSQ_SHIFT_DEQUANT) { FWDS(SSQ_SHIFT_DEQUANT) } else { FWDS(SSQ_SHIFT_ADASHIFT) }
#undef FWDS
#undef FWDV
#undef FWDG
#undef FWDK
        return launch_status();
    }
    if (mode == SSQ_SHIFT_DEQUANT)
        fq_shift_fwd_kernel<SSQ_SHIFT_DEQUANT><<<grid, SSQ_THREADS, 0, st>>>(w, shift_delta, delta, zero_point, p, beta, y, oc, K, kk, nshift, per_element, hard_targets, hard_round, qmin, qmax);
    else if (mode == SSQ_SHIFT_ADASHIFT)
        fq_shift_fwd_kernel<SSQ_SHIFT_ADASHIFT><<<grid, SSQ_THREADS, 0, st>>>(w, shift_delta, delta, zero_point, p, beta, y, oc, K, kk, nshift, per_element, hard_targets, hard_round, qmin, qmax);
    else return SSQ_ERR_MODE;
    return launch_status();
}

extern "C" size_t ssq_shift_bwd_ws_bytes(int64_t oc, int64_t ic, int64_t kk, int nshift, int per_element) {
    if (per_element || oc <= 0 || ic <= 0 || kk <= 0) return 16;
    int nslab, nslab_v = 0; int64_t rps;
    slab_plan(oc, ic * kk, nslab, rps);
    if ((ic * kk) % 4 == 0)
        for (int mode = 0; mode < 2; ++mode) {           // the mode is not known here: room for either plan
            slab_plan(oc, ic * kk, nslab_v, rps, SSQ_K1C_BWD_CTAS(mode));
            if (nslab_v > nslab) nslab = nslab_v;
        }
    return (size_t)nslab * (size_t)(ic * kk) * (size_t)nshift * sizeof(float) + 16;
}

extern "C" int ssq_fq_shift_bwd(const float* gy, const float* w, const float* shift_delta, const float* delta,
                                const float* zero_point, const float* p, const float* beta,
                                float* gp, float* gbeta,
                                int64_t oc, int64_t ic, int64_t kk, int nshift, int per_element,
                                int mode, int hard_round, float qmin, float qmax,
                                void* ws, size_t ws_bytes, void* stream) {
    if (oc == 0 || ic == 0 || kk == 0) return SSQ_OK;
    if (!gy || !w || !shift_delta || !delta || !zero_point || !p || !gp) return SSQ_ERR_NULL;
    if (oc < 0 || ic < 0 || kk < 0 || nshift < 1 || nshift > MS) return SSQ_ERR_SIZE;
    if (mode == SSQ_SHIFT_ADASHIFT && !beta) return SSQ_ERR_NULL;
    if (!per_element && (!ws || ws_bytes < ssq_shift_bwd_ws_bytes(oc, ic, kk, nshift, per_element))) return SSQ_ERR_WORKSPACE;
    const int64_t K = ic * kk;
    if (mode != SSQ_SHIFT_DEQUANT && mode != SSQ_SHIFT_ADASHIFT) return SSQ_ERR_MODE;
    const bool vec = !per_element && (K % 4 == 0) && K >= 8 && (oc * K) / 4 < 0x7fffffff && aligned16(gy) && aligned16(w) &&
                     (!beta || aligned16(beta)) && (!gbeta || aligned16(gbeta));
    int nslab; int64_t rps;
    slab_plan(oc, K, nslab, rps, vec ? SSQ_K1C_BWD_CTAS(mode) : 0);
    if (nslab > 65535) return SSQ_ERR_SIZE;
    cudaStream_t st = (cudaStream_t)stream;
    float* partial = reinterpret_cast<float*>(ws);
    if (vec) {
        dim3 vgrid((unsigned)((K / 4 + SSQ_THREADS - 1) / SSQ_THREADS), (unsigned)nslab);
        ShiftArgs a = shift_args(w, shift_delta, delta, zero_point, p, beta, gy, nullptr, gbeta, partial, oc, K, kk, nshift,
                                 0, hard_round, qmin, qmax, rps);
        const int gk = group_kind(kk, p);
#define BWDK(M, SS, SO, GK) fq_shift_bwd_vec<M, SS, SO, GK><<<vgrid, SSQ_THREADS, 0, st>>>(a)
#define BWDG(M, SS, SO) do { if (gk == 1) BWDK(M, SS, SO, 1); else if (gk == 2) BWDK(M, SS, SO, 2); else BWDK(M, SS, SO, 0); } while (0)
#define BWDV(M, SS) do { if (!hard_round) BWDG(M, SS, true); else BWDG(M, SS, false); } while (0)
#define BWDS(M) switch (nshift) { case 1: BWDV(M, 1); break; case 2: BWDV(M, 2); break; case 3: BWDV(M, 3); break; default: BWDV(M, 4); }
        if (mode == SSQ_SHIFT_DEQUANT) { BWDS(SSQ_SHIFT_DEQUANT) } else { BWDS(SSQ_SHIFT_ADASHIFT) }
#undef BWDS
#undef BWDV
#undef BWDG
#undef BWDK
        int ev = launch_status();
        if (ev) return ev;
        fq_shift_bwd_finish_kernel<<<finish_grid(ic, nshift), SSQ_THREADS, 0, st>>>(partial, gp, ic, K, kk, nshift, nslab);
        return launch_status();
    }
    dim3 grid((unsigned)((K + SSQ_THREADS - 1) / SSQ_THREADS), (unsigned)nslab);
    if (mode == SSQ_SHIFT_DEQUANT)
        fq_shift_bwd_kernel<SSQ_SHIFT_DEQUANT><<<grid, SSQ_THREADS, 0, st>>>(gy, w, shift_delta, delta, zero_point, p, beta, gp, gbeta, partial, oc, K, kk, nshift, per_element, hard_round, qmin, qmax, rps);
    else if (mode == SSQ_SHIFT_ADASHIFT)
        fq_shift_bwd_kernel<SSQ_SHIFT_ADASHIFT><<<grid, SSQ_THREADS, 0, st>>>(gy, w, shift_delta, delta, zero_point, p, beta, gp, gbeta, partial, oc, K, kk, nshift, per_element, hard_round, qmin, qmax, rps);
    else return SSQ_ERR_MODE;
    int e = launch_status();
    if (e || per_element) return e;
    fq_shift_bwd_finish_kernel<<<finish_grid(ic, nshift), SSQ_THREADS, 0, st>>>(partial, gp, ic, K, kk, nshift, nslab);
    return launch_status();
}
