// K1c — shifted-scale ChannelQuant: per-input-channel (conv) or per-element (FC) soft/hard mixture
// over S shifted scales, optionally fused with AdaRound soft rounding ('adaShift').
//   reference arithmetic: quant/channelQuant.py:49-127 (forward modes, shifted_x_quant, soft targets),
//   :201-213 (init_v: dequantised candidates), :279-294 (init_v_beta: integer-floor candidates);
//   regularisers quant/layer_recon_shiftedScale.py:386-393, quant/layer_recon_fused_shiftedScale.py:277-282.
// The reference reads S cached weight-sized tensors x_q[i]; here the S candidates are recomputed from the
// weight in registers (12 B/elem instead of (S+2)*4 B/elem).
#include "ssq_common.cuh"

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
// Same arithmetic as shift_one on quotients that were divided four at a time against a reciprocal hoisted per (row, shift).
template <int MODE, int S>
__device__ __forceinline__ float shift_eval(const float (&u)[S], const float (&ds)[S], float d, float z, const float (&p)[S],
                                            float beta, int hard_targets, int hard_round, float qmin, float qmax,
                                            float (&t)[S], bool& inside, float& dh) {
#pragma unroll
    for (int i = 0; i < S; ++i) {
        if (MODE == SSQ_SHIFT_DEQUANT) {
            const float q = clampk(__fadd_rn(rintf(u[i]), z), qmin, qmax);
            t[i] = __fmul_rn(__fsub_rn(q, z), ds[i]);
        } else {
            t[i] = floorf(u[i]);
        }
    }
    float mix;
    if (hard_targets) {
        float pbest = p[0];                       // torch.argmax: first maximum wins
        mix = t[0];
#pragma unroll
        for (int i = 1; i < S; ++i) if (p[i] > pbest) { pbest = p[i]; mix = t[i]; }
    } else {
        mix = __fmul_rn(t[0], p[0]);
#pragma unroll
        for (int i = 1; i < S; ++i) mix = __fadd_rn(mix, __fmul_rn(t[i], p[i]));
    }
    inside = true; dh = 0.f;
    if (MODE == SSQ_SHIFT_DEQUANT) return mix;
    float r;
    if (hard_round) r = (beta >= 0.f) ? 1.f : 0.f;
    else { float h; dh = rect_sigmoid_grad(beta, h); r = h; }
    const float xi = __fadd_rn(__fadd_rn(mix, r), z);
    inside = (xi >= qmin) && (xi <= qmax);
    return __fmul_rn(__fsub_rn(clampk(xi, qmin, qmax), z), d);
}

// forward: address-ordered tiles of 256*U float4s (ssq_common.cuh); row = output channel, group = (k / kk).
// SOFT: soft targets and soft round known at compile time (the training loops) — the per-element mode branches vanish.
// WIDE: kk >= 4, so a float4 touches at most two groups: both probability rows are loaded up front and selected per element
// (kk < 4, i.e. 1x1 kernels, reloads at every group change).
template <int MODE, int S, bool SOFT, bool WIDE>
__global__ void __launch_bounds__(SSQ_THREADS, 5)
fq_shift_fwd_vec(const float* __restrict__ w, const float* __restrict__ shift_delta, const float* __restrict__ delta,
                 const float* __restrict__ zp, const float* __restrict__ p, const float* __restrict__ beta,
                 float* __restrict__ y, int64_t oc, uint32_t K4, uint32_t kk, int hard_targets_rt, int hard_round_rt,
                 float qmin, float qmax) {
    constexpr int U = 2;
    const int hard_targets = SOFT ? 0 : hard_targets_rt, hard_round = SOFT ? 0 : hard_round_rt;
    const uint32_t last_group = (K4 * 4) / kk - 1;
    const int64_t total4 = oc * (int64_t)K4;
    const int64_t i0 = (int64_t)blockIdx.x * (SSQ_THREADS * U) + threadIdx.x;
    float4 wv[U], bv[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
        const int64_t i = i0 + (int64_t)u * SSQ_THREADS;
        wv[u] = bv[u] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (i < total4) { wv[u] = ld_stream4(w + i * 4); if (MODE == SSQ_SHIFT_ADASHIFT) bv[u] = ld_stream4(beta + i * 4); }
    }
    TileWalk tw;                                   // c = row (output channel), col = float4 column inside the row
    tw.init((uint64_t)i0, K4, (uint64_t)oc);
    float ds[S], d = 0.f, z = 0.f;
    Recip R[S];
    uint32_t r_have = 0xffffffffu;
#pragma unroll
    for (int u = 0; u < U; ++u) {
        const int64_t i = i0 + (int64_t)u * SSQ_THREADS;
        if (i < total4) {
            const uint32_t r = tw.c, k0 = tw.col * 4;
            if (r != r_have) {                     // row parameters: reused by the thread's next vector when it is in the same row
#pragma unroll
                for (int sft = 0; sft < S; ++sft) { ds[sft] = __ldg(shift_delta + (int64_t)sft * oc + r); R[sft] = make_recip(ds[sft]); }
                d = __ldg(delta + r); z = __ldg(zp + r);
                r_have = r;
            }
            float qa[S][4];
            const bool small = small4(wv[u]);
#pragma unroll
            for (int sft = 0; sft < S; ++sft) {
                const float4 q = div4_exact(wv[u], R[sft], small);
                qa[sft][0] = q.x; qa[sft][1] = q.y; qa[sft][2] = q.z; qa[sft][3] = q.w;
            }
            uint32_t g = k0 / kk, rem = k0 - g * kk;
            float pv[S], p1[S];
#pragma unroll
            for (int sft = 0; sft < S; ++sft) pv[sft] = __ldg(p + g * S + sft);
            const uint32_t n0 = kk - rem;          // elements of this vector that still belong to group g
            if (WIDE) {
                const uint32_t g1 = g < last_group ? g + 1 : g;
#pragma unroll
                for (int sft = 0; sft < S; ++sft) p1[sft] = __ldg(p + g1 * S + sft);
            }
            const float be[4] = {bv[u].x, bv[u].y, bv[u].z, bv[u].w};
            float out[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                if (WIDE) {
                    if (e > 0 && (uint32_t)e == n0) {
#pragma unroll
                        for (int sft = 0; sft < S; ++sft) pv[sft] = p1[sft];
                    }
                } else if (rem == kk) {            // next input channel: its own group probabilities
                    rem = 0; ++g;
#pragma unroll
                    for (int sft = 0; sft < S; ++sft) pv[sft] = __ldg(p + g * S + sft);
                }
                float uu[S], t[S]; bool inside; float dh;
#pragma unroll
                for (int sft = 0; sft < S; ++sft) uu[sft] = qa[sft][e];
                out[e] = shift_eval<MODE, S>(uu, ds, d, z, pv, be[e], hard_targets, hard_round, qmin, qmax, t, inside, dh);
                ++rem;
            }
            st_stream4(y + i * 4, make_float4(out[0], out[1], out[2], out[3]));
        }
        tw.step(SSQ_THREADS);
    }
}

// backward: a thread owns 4 adjacent columns and walks the rows of its slab, two rows in flight; per-column sums of
// d y / d p[g, i] in registers -> partial[slab][K][S] (the finish kernel adds slabs and the kk columns of a group in fp64)
template <int MODE, int S, bool SOFT>
__global__ void __launch_bounds__(SSQ_THREADS, 4)
fq_shift_bwd_vec(const float* __restrict__ gy, const float* __restrict__ w, const float* __restrict__ shift_delta,
                 const float* __restrict__ delta, const float* __restrict__ zp, const float* __restrict__ p,
                 const float* __restrict__ beta, float* __restrict__ gbeta, float* __restrict__ partial,
                 int64_t oc, uint32_t K4, uint32_t kk, int hard_round_rt, float qmin, float qmax, int64_t rows_per_slab) {
    const uint32_t col4 = blockIdx.x * blockDim.x + threadIdx.x;
    if (col4 >= K4) return;
    const int hard_round = SOFT ? 0 : hard_round_rt;
    const int64_t K = (int64_t)K4 * 4;
    const int64_t r0 = (int64_t)blockIdx.y * rows_per_slab;
    const int64_t r1 = r0 + rows_per_slab < oc ? r0 + rows_per_slab : oc;
    float pe[4][S], acc[4][S];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        const uint32_t g = (col4 * 4 + e) / kk;
#pragma unroll
        for (int sft = 0; sft < S; ++sft) { pe[e][sft] = __ldg(p + g * S + sft); acc[e][sft] = 0.f; }
    }
    auto row = [&](int64_t r, const float4& g4, const float4& w4, const float4& b4) {
        float ds[S], qa[S][4];
        const bool small = small4(w4);
#pragma unroll
        for (int sft = 0; sft < S; ++sft) {
            ds[sft] = __ldg(shift_delta + (int64_t)sft * oc + r);
            const float4 q = div4_exact(w4, make_recip(ds[sft]), small);
            qa[sft][0] = q.x; qa[sft][1] = q.y; qa[sft][2] = q.z; qa[sft][3] = q.w;
        }
        const float d = __ldg(delta + r), z = __ldg(zp + r);
        const float ge[4] = {g4.x, g4.y, g4.z, g4.w}, be[4] = {b4.x, b4.y, b4.z, b4.w};
        float gb[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            float uu[S], t[S]; bool inside; float dh;
#pragma unroll
            for (int sft = 0; sft < S; ++sft) uu[sft] = qa[sft][e];
            (void)shift_eval<MODE, S>(uu, ds, d, z, pe[e], be[e], 0, hard_round, qmin, qmax, t, inside, dh);
            float gm = ge[e];                       // gradient wrt the mixture
            gb[e] = 0.f;
            if (MODE == SSQ_SHIFT_ADASHIFT) {
                gm = inside ? ge[e] * d : 0.f;
                gb[e] = hard_round ? 0.f : gm * dh;
            }
#pragma unroll
            for (int sft = 0; sft < S; ++sft) acc[e][sft] += gm * t[sft];
        }
        if (MODE == SSQ_SHIFT_ADASHIFT && gbeta) st_stream4(gbeta + r * K + (int64_t)col4 * 4, make_float4(gb[0], gb[1], gb[2], gb[3]));
    };
    const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
    int64_t r = r0;
    for (; r + 1 < r1; r += 2) {
        const int64_t ea = r * K + (int64_t)col4 * 4, eb = ea + K;
        const float4 ga = ld_stream4(gy + ea), wa = ld_stream4(w + ea), gbv = ld_stream4(gy + eb), wb = ld_stream4(w + eb);
        const float4 ba = (MODE == SSQ_SHIFT_ADASHIFT) ? ld_stream4(beta + ea) : zero4;
        const float4 bb = (MODE == SSQ_SHIFT_ADASHIFT) ? ld_stream4(beta + eb) : zero4;
        row(r, ga, wa, ba);
        row(r + 1, gbv, wb, bb);
    }
    if (r < r1) {
        const int64_t ea = r * K + (int64_t)col4 * 4;
        row(r, ld_stream4(gy + ea), ld_stream4(w + ea), (MODE == SSQ_SHIFT_ADASHIFT) ? ld_stream4(beta + ea) : zero4);
    }
#pragma unroll
    for (int e = 0; e < 4; ++e)
#pragma unroll
        for (int sft = 0; sft < S; ++sft)
            partial[((int64_t)blockIdx.y * K + (int64_t)col4 * 4 + e) * S + sft] = acc[e][sft];
}

__global__ void __launch_bounds__(SSQ_THREADS)
fq_shift_bwd_finish_kernel(const float* __restrict__ partial, float* __restrict__ gp, int64_t ic, int64_t K, int64_t kk,
                           int S, int nslab) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= ic * S) return;
    const int64_t g = t / S; const int i = (int)(t - g * S);
    double s = 0.0;
    for (int sl = 0; sl < nslab; ++sl)
        for (int64_t k = g * kk; k < (g + 1) * kk; ++k) s += (double)partial[((int64_t)sl * K + k) * S + i];
    gp[t] = (float)s;
}

static inline void slab_plan(int64_t oc, int64_t K, int& nslab, int64_t& rows_per_slab, bool vec = false) {
    int64_t colblocks = ((vec ? K / 4 : K) + SSQ_THREADS - 1) / SSQ_THREADS;
    // vec: ~3 resident CTAs per SM (about 80 registers), slabs of at least 4 rows so the partials stay small
    int64_t want = ((int64_t)SSQ_NUM_SMS * (vec ? 4 : SSQ_CTAS_PER_SM) + colblocks - 1) / colblocks;
    if (vec && want > (oc + 3) / 4) want = (oc + 3) / 4;
    if (want > oc) want = oc;
    if (want < 1) want = 1;
    rows_per_slab = (oc + want - 1) / want;
    nslab = (int)((oc + rows_per_slab - 1) / rows_per_slab);
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
    const bool vec = !per_element && (K % 4 == 0) && K / 4 < 0x7fffffff && kk < 0x7fffffff && aligned16(w) && aligned16(y) &&
                     (!beta || aligned16(beta));
    if (vec) {
        const unsigned vgrid = tile_grid((n / 4 + SSQ_THREADS * 2 - 1) / (SSQ_THREADS * 2), false);
        const bool soft = !hard_targets && !hard_round, wide = kk >= 4;
#define FWDK(M, SS, SO, WI) fq_shift_fwd_vec<M, SS, SO, WI><<<vgrid, SSQ_THREADS, 0, st>>>(w, shift_delta, delta, zero_point, p, beta, y, oc, \
        (uint32_t)(K / 4), (uint32_t)kk, hard_targets, hard_round, qmin, qmax)
#define FWDV(M, SS) do { if (soft) { if (wide) FWDK(M, SS, true, true); else FWDK(M, SS, true, false); } \
                         else { if (wide) FWDK(M, SS, false, true); else FWDK(M, SS, false, false); } } while (0)
#define FWDS(M) switch (nshift) { case 1: FWDV(M, 1); break; case 2: FWDV(M, 2); break; case 3: FWDV(M, 3); break; default: FWDV(M, 4); }
        if (mode == SSQ_SHIFT_DEQUANT) { FWDS(SSQ_SHIFT_DEQUANT) } else { FWDS(SSQ_SHIFT_ADASHIFT) }
#undef FWDS
#undef FWDV
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
    if ((ic * kk) % 4 == 0) slab_plan(oc, ic * kk, nslab_v, rps, true);
    if (nslab_v > nslab) nslab = nslab_v;
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
    const bool vec = !per_element && (K % 4 == 0) && K / 4 < 0x7fffffff && kk < 0x7fffffff && aligned16(gy) && aligned16(w) &&
                     (!beta || aligned16(beta)) && (!gbeta || aligned16(gbeta));
    int nslab; int64_t rps;
    slab_plan(oc, K, nslab, rps, vec);
    if (nslab > 65535) return SSQ_ERR_SIZE;
    cudaStream_t st = (cudaStream_t)stream;
    float* partial = reinterpret_cast<float*>(ws);
    if (vec) {
        dim3 vgrid((unsigned)((K / 4 + SSQ_THREADS - 1) / SSQ_THREADS), (unsigned)nslab);
#define BWDK(M, SS, SO) fq_shift_bwd_vec<M, SS, SO><<<vgrid, SSQ_THREADS, 0, st>>>(gy, w, shift_delta, delta, zero_point, p, beta, gbeta, partial, \
        oc, (uint32_t)(K / 4), (uint32_t)kk, hard_round, qmin, qmax, rps)
#define BWDV(M, SS) do { if (!hard_round) BWDK(M, SS, true); else BWDK(M, SS, false); } while (0)
#define BWDS(M) switch (nshift) { case 1: BWDV(M, 1); break; case 2: BWDV(M, 2); break; case 3: BWDV(M, 3); break; default: BWDV(M, 4); }
        if (mode == SSQ_SHIFT_DEQUANT) { BWDS(SSQ_SHIFT_DEQUANT) } else { BWDS(SSQ_SHIFT_ADASHIFT) }
#undef BWDS
#undef BWDV
#undef BWDK
        int ev = launch_status();
        if (ev) return ev;
        fq_shift_bwd_finish_kernel<<<(unsigned)((ic * nshift + SSQ_THREADS - 1) / SSQ_THREADS), SSQ_THREADS, 0, st>>>(partial, gp, ic, K, kk, nshift, nslab);
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
    fq_shift_bwd_finish_kernel<<<(unsigned)((ic * nshift + SSQ_THREADS - 1) / SSQ_THREADS), SSQ_THREADS, 0, st>>>(partial, gp, ic, K, kk, nshift, nslab);
    return launch_status();
}
