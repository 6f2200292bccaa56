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

static inline void slab_plan(int64_t oc, int64_t K, int& nslab, int64_t& rows_per_slab) {
    int64_t colblocks = (K + SSQ_THREADS - 1) / SSQ_THREADS;
    int64_t want = ((int64_t)SSQ_NUM_SMS * SSQ_CTAS_PER_SM + colblocks - 1) / colblocks;
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
    if (mode == SSQ_SHIFT_DEQUANT)
        fq_shift_fwd_kernel<SSQ_SHIFT_DEQUANT><<<grid, SSQ_THREADS, 0, st>>>(w, shift_delta, delta, zero_point, p, beta, y, oc, K, kk, nshift, per_element, hard_targets, hard_round, qmin, qmax);
    else if (mode == SSQ_SHIFT_ADASHIFT)
        fq_shift_fwd_kernel<SSQ_SHIFT_ADASHIFT><<<grid, SSQ_THREADS, 0, st>>>(w, shift_delta, delta, zero_point, p, beta, y, oc, K, kk, nshift, per_element, hard_targets, hard_round, qmin, qmax);
    else return SSQ_ERR_MODE;
    return launch_status();
}

extern "C" size_t ssq_shift_bwd_ws_bytes(int64_t oc, int64_t ic, int64_t kk, int nshift, int per_element) {
    if (per_element || oc <= 0 || ic <= 0 || kk <= 0) return 16;
    int nslab; int64_t rps;
    slab_plan(oc, ic * kk, nslab, rps);
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
    int nslab; int64_t rps;
    slab_plan(oc, K, nslab, rps);
    if (nslab > 65535) return SSQ_ERR_SIZE;
    cudaStream_t st = (cudaStream_t)stream;
    dim3 grid((unsigned)((K + SSQ_THREADS - 1) / SSQ_THREADS), (unsigned)nslab);
    float* partial = reinterpret_cast<float*>(ws);
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
