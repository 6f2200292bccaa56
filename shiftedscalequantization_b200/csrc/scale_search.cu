// K2 — scale searches.
//  K2a  per-channel / per-tensor MSE clip-ratio search: quant/quant_layer.py:145-162, :168-175.
//       80 candidates x |x - x_q|^2.4 per element is instruction-bound, not HBM-bound, and libm powf is ~90 % of it.
//       Two passes with an identical argmin:
//         rank   every candidate scored with |d|^p = ex2(p * lg2 |d|) (two MUFU ops; the quantised value itself is
//                computed exactly, so only the power differs: relative score error < 8e-6, see below);
//         settle the candidates whose ranked score lies within 1e-4 of the ranked minimum (1.0-1.1 of the 80 on
//                weight rows) are re-scored with libm powf in a FIXED order; the first strict minimum among them wins.
//       The winner of the full libm scan s* satisfies rank(s*) <= s*(1+e) <= min_rank (1+e)/(1-e), so it is always
//       inside the window, as is every exact tie; everything outside the window is strictly worse in exact arithmetic.
//       Row variants: one CTA per output channel (row staged ONCE into shared memory with a TMA bulk copy,
//       one warp per candidate), or — rows shorter than 128 elements, e.g. depthwise 3x3 — one WARP per row with the
//       lanes spread over the candidates. Tensor variant (activations, up to ~50 M elements): grid-wide, every element
//       read once and ranked against all 80 candidates from registers; a second, short launch settles the survivors.
//  K2b  ChannelQuantMSE input-scale search: quant/channelQuantMSE.py:70-110. One pass over the weights (4 B/element):
//       the reference predicate is monotone in the candidate, which turns the `level`-candidate sweep into one
//       estimate per element, verified with the exact predicate wherever the estimate is near a boundary.
#include "ssq_common.cuh"
#include <atomic>

namespace ssq {

constexpr int NC = SSQ_N_CANDIDATES;
constexpr float RANK_WINDOW = 1e-4f;   // >= 10x the worst-case relative error of the ranked scores (header comment)

__device__ __forceinline__ float clampq(float v, float lo, float hi) { return v < lo ? lo : (v > hi ? hi : v); }

// candidate i -> (delta_i, zp_i, new_min_i), all in the reference's fp32 op order
struct Cand { float d, z, nmin; };
__device__ __forceinline__ Cand make_cand(int i, float x_min, float x_max, float lm1) {
    // 1.0 - (i * 0.01) is Python double arithmetic; ATen casts the scalar to fp32 before the multiply
    float f = (float)(1.0 - ((double)i * 0.01));
    float nmax = __fmul_rn(x_max, f), nmin = __fmul_rn(x_min, f);
    Cand c;
    c.d = __fdiv_rn(__fsub_rn(nmax, nmin), lm1);
    c.z = rintf(__fdiv_rn(-nmin, c.d));
    c.nmin = nmin;
    return c;
}
// settle pass: the reference expression with libm powf (the reciprocal of delta_i is hoisted per candidate; div_exact
// returns the IEEE quotient either way)
__device__ __forceinline__ float cand_err(float x, const Cand& c, const Recip& R, float lm1, float p) {
    float q = clampq(__fadd_rn(rintf(div_exact(x, R)), c.z), 0.f, lm1);
    float xq = __fmul_rn(__fsub_rn(q, c.z), c.d);
    return pow_scalar_accurate(fabsf(__fsub_rn(x, xq)), p);
}
// rank pass. Valid when the row is finite and < 2^67 in magnitude, delta_i in [2^-60, 2^60] and |z_i| < 2^22:
//   clamp(rint(u) + z, 0, L-1) - z == clamp(rint(u), -z, L-1-z) exactly (integers below 2^24), so |x - x_q| is the
//   settle pass's value bit for bit and only the power is approximate: lg2.approx (abs. error 2^-22.6), the fp32
//   product p*lg2 (<= 2^-18 abs. for |p lg2| < 128) and ex2.approx (2^-22 rel.) give < 4e-6 relative per term, hence
//   per sum of non-negative terms; the fp32 warp-tree / 8-term partial sums add < 1e-6.
struct RankCand { float d, r, lo, hi; };
__device__ __forceinline__ bool rank_cand(const Cand& c, float lm1, RankCand& k) {
    const Recip R = make_recip(c.d);
    k.d = c.d; k.r = R.r; k.lo = -c.z; k.hi = lm1 - c.z;
    return R.ok && fabsf(c.z) < 4194304.0f;       // false for NaN z
}
__device__ __forceinline__ float rank_err(float x, const RankCand& k, float p) {
    const float q0 = __fmul_rn(x, k.r);
    const float u = fmaf(k.r, fmaf(-k.d, q0, x), q0);                 // div_fast
    const float q = fminf(fmaxf(rintf(u), k.lo), k.hi);
    const float e = fabsf(__fsub_rn(x, __fmul_rn(q, k.d)));
    return ex2_approx(p * log2_for_pow(e));
}
__device__ __forceinline__ bool needs_rank(float p) { return !(p == 2.0f || p == 1.0f); }   // those powers are exact and cheap
__device__ __forceinline__ void apply_sym(float& x_min, float& x_max, int symmetric) {
    if (symmetric) {
        float am = fmaxf(fabsf(x_min), x_max);
        x_min = (x_min < 0.f) ? -am : 0.f;
        x_max = am;
    }
}
// the winning candidate's outputs (quant_layer.py:157-162); idx < 0 = no candidate scored below 1e10
__device__ __forceinline__ void write_winner(int idx, float best, float x_min, float x_max, float lm1, int symmetric,
                                             int64_t row, float* delta, float* zp, float* raw, float* best_score,
                                             int32_t* best_index) {
    float d = nanf(""), z = nanf(""), r = nanf("");
    if (idx >= 0) {
        Cand c = make_cand(idx, x_min, x_max, lm1);
        d = c.d;
        z = symmetric ? 0.f : c.z;
        r = symmetric ? 0.f : -c.nmin;
    }
    delta[row] = d; zp[row] = z; raw[row] = r;
    if (best_score) best_score[row] = best;
    if (best_index) best_index[row] = idx;
}

// ---- TMA bulk copy of one row into shared memory ------------------------------------------------
__device__ __forceinline__ void tma_row_to_smem(float* srow, const float* grow, uint32_t bytes, uint64_t* bar) {
    uint32_t bar_s = (uint32_t)__cvta_generic_to_shared(bar);
    uint32_t dst_s = (uint32_t)__cvta_generic_to_shared(srow);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(bar_s));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar_s), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     :: "r"(dst_s), "l"(grow), "r"(bytes), "r"(bar_s) : "memory");
    }
    uint32_t done = 0;
    while (!done) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\t"
                     "selp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(bar_s) : "memory");
    }
}

// exact score of one candidate over a shared-memory row, lanes = elements; THE fixed order every variant settles in:
// 8-term fp32 partials at stride 32, fp64 running sum per lane, shuffle tree, / k, -> fp32
__device__ __forceinline__ float settle_row_score(const float* srow, int64_t k, int lane, const Cand& c, float lm1, float p) {
    const Recip R = make_recip(c.d);
    double acc = 0.0;
    for (int64_t j0 = lane; j0 < k; j0 += 32 * 8) {
        float s = 0.f;
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            int64_t j = j0 + e * 32;
            if (j < k) s += cand_err(srow[j], c, R, lm1, p);
        }
        acc += (double)s;
    }
    acc = warp_sum(acc);
    acc = __shfl_sync(0xffffffffu, acc, 0);
    return (float)(acc / (double)k);
}

// ---- K2a rows, CTA per row ------------------------------------------------------------------------
__global__ void __launch_bounds__(SSQ_THREADS)
mse_search_rows_kernel(const float* __restrict__ x, int64_t k, float lm1, int symmetric, float p,
                       float* __restrict__ delta, float* __restrict__ zp, float* __restrict__ raw,
                       float* __restrict__ best_score, int32_t* __restrict__ best_index) {
    extern __shared__ __align__(128) unsigned char dyn[];
    float* srow = reinterpret_cast<float*>(dyn);
    __shared__ uint64_t bar;
    __shared__ float s_rank[NC];            // ranked scores; NaN = "settle me"
    __shared__ float s_scores[NC];
    __shared__ int s_surv[NC];
    __shared__ int s_nsurv;
    __shared__ float s_red[2 * 32];
    __shared__ int s_bad;
    const int64_t row = blockIdx.x;
    const float* grow = x + row * k;
    const uint32_t bytes = (uint32_t)(k * 4);
    if (threadIdx.x == 0) s_bad = 0;
    if ((bytes & 15u) == 0 && aligned16(grow)) {
        tma_row_to_smem(srow, grow, bytes, &bar);
    } else {
        for (int64_t j = threadIdx.x; j < k; j += blockDim.x) srow[j] = ld_stream1(grow + j);
        __syncthreads();
    }
    // row min / max (exact, order independent); `bad` = NaN / Inf / huge element: no ranking for this row
    float mn = INFINITY, mx = -INFINITY; bool bad = false;
    for (int64_t j = threadIdx.x; j < k; j += blockDim.x) {
        float v = srow[j]; mn = fminf(mn, v); mx = fmaxf(mx, v); bad |= !(fabsf(v) < SSQ_DIV_XMAX);
    }
    mn = warp_min(mn); mx = warp_max(mx);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    if (lane == 0) { s_red[warp] = mn; s_red[32 + warp] = mx; }
    if (bad) s_bad = 1;
    __syncthreads();
    mn = s_red[0]; mx = s_red[32];
    for (int w = 1; w < nwarp; ++w) { mn = fminf(mn, s_red[w]); mx = fmaxf(mx, s_red[32 + w]); }
    apply_sym(mn, mx, symmetric);
    const bool rank = needs_rank(p) && !s_bad;
    // ---- rank: one warp per candidate
    if (rank) {
        for (int i = warp; i < NC; i += nwarp) {
            const Cand c = make_cand(i, mn, mx, lm1);
            RankCand rc;
            float out = nanf("");
            if (rank_cand(c, lm1, rc)) {
                double acc = 0.0;
                for (int64_t j0 = lane; j0 < k; j0 += 32 * 8) {
                    float s = 0.f;
#pragma unroll
                    for (int e = 0; e < 8; ++e) {
                        int64_t j = j0 + e * 32;
                        if (j < k) s += rank_err(srow[j], rc, p);
                    }
                    acc += (double)s;
                }
                acc = warp_sum(acc);
                out = (float)(acc / (double)k);
            }
            if (lane == 0) s_rank[i] = out;
        }
        __syncthreads();
    }
    // ---- survivors, in index order
    if (warp == 0) {
        float lo = INFINITY;
        if (rank) {
            for (int i = lane; i < NC; i += 32) lo = fminf(lo, s_rank[i]);     // fminf skips NaN
            lo = warp_min(lo);
        }
        const float thr = lo + lo * RANK_WINDOW;
        int base = 0;
        for (int i0 = 0; i0 < NC; i0 += 32) {
            const int i = i0 + lane;
            const bool keep = i < NC && (!rank || !(s_rank[i] > thr));          // NaN / unranked candidates stay in
            const unsigned m = __ballot_sync(0xffffffffu, keep);
            if (keep) s_surv[base + __popc(m & ((1u << lane) - 1u))] = i;
            base += __popc(m);
        }
        if (lane == 0) s_nsurv = base;
    }
    __syncthreads();
    // ---- settle
    const int ns = s_nsurv;
    for (int s = warp; s < ns; s += nwarp) {
        const int i = s_surv[s];
        const float sc = settle_row_score(srow, k, lane, make_cand(i, mn, mx, lm1), lm1, p);
        if (lane == 0) s_scores[s] = sc;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        float best = 1e10f; int idx = -1;
        for (int s = 0; s < ns; ++s) if (s_scores[s] < best) { best = s_scores[s]; idx = s_surv[s]; }
        write_winner(idx, best, mn, mx, lm1, symmetric, row, delta, zp, raw, best_score, best_index);
    }
}

// ---- K2a rows, warp per row (k < SHORT_ROW): depthwise 3x3 (k = 9), the stem (k = 27 .. 147) ----------------------
constexpr int SHORT_ROW = 128;
constexpr int SHORT_WARPS = SSQ_THREADS / 32;
__global__ void __launch_bounds__(SSQ_THREADS)
mse_search_short_rows_kernel(const float* __restrict__ x, int64_t rows, int k, float lm1, int symmetric, float p,
                             float* __restrict__ delta, float* __restrict__ zp, float* __restrict__ raw,
                             float* __restrict__ best_score, int32_t* __restrict__ best_index) {
    __shared__ float s_rows[SHORT_WARPS][SHORT_ROW];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t row = (int64_t)blockIdx.x * SHORT_WARPS + warp;
    if (row >= rows) return;
    float* srow = s_rows[warp];
    const float* grow = x + row * k;
    float mn = INFINITY, mx = -INFINITY; bool bad = false;
    for (int j = lane; j < k; j += 32) {
        float v = ld_stream1(grow + j); srow[j] = v;
        mn = fminf(mn, v); mx = fmaxf(mx, v); bad |= !(fabsf(v) < SSQ_DIV_XMAX);
    }
    __syncwarp();
    mn = warp_min(mn); mx = warp_max(mx);
    bad = __any_sync(0xffffffffu, bad);
    apply_sym(mn, mx, symmetric);
    const bool rank = needs_rank(p) && !bad;
    // ---- rank: lanes = candidates (lane, lane+32, lane+64), the row broadcast from shared memory
    float rk[3] = {nanf(""), nanf(""), nanf("")};
    if (rank) {
        RankCand rc[3]; bool ok[3];
#pragma unroll
        for (int t = 0; t < 3; ++t) {
            const int i = lane + 32 * t;
            ok[t] = i < NC && rank_cand(make_cand(i < NC ? i : 0, mn, mx, lm1), lm1, rc[t]);
        }
        float s[3] = {0.f, 0.f, 0.f};
        for (int j = 0; j < k; ++j) {
            const float v = srow[j];
#pragma unroll
            for (int t = 0; t < 3; ++t) s[t] += rank_err(v, rc[t], p);
        }
#pragma unroll
        for (int t = 0; t < 3; ++t) if (ok[t]) rk[t] = s[t] / (float)k;
    }
    float lo = INFINITY;
#pragma unroll
    for (int t = 0; t < 3; ++t) if (lane + 32 * t < NC) lo = fminf(lo, rk[t]);
    lo = warp_min(lo);
    const float thr = rank ? lo + lo * RANK_WINDOW : INFINITY;
    // ---- settle the survivors in index order (lanes = elements now), first strict minimum
    float best = 1e10f; int idx = -1;
#pragma unroll
    for (int t = 0; t < 3; ++t) {
        const bool keep = (lane + 32 * t < NC) && (!rank || !(rk[t] > thr));
        unsigned m = __ballot_sync(0xffffffffu, keep);
        while (m) {
            const int i = 32 * t + (__ffs(m) - 1);
            m &= m - 1;
            const float sc = settle_row_score(srow, k, lane, make_cand(i, mn, mx, lm1), lm1, p);
            if (sc < best) { best = sc; idx = i; }
        }
    }
    if (lane == 0) write_winner(idx, best, mn, mx, lm1, symmetric, row, delta, zp, raw, best_score, best_index);
}

// ---- min / max, grid-wide -------------------------------------------------------------------------
// grid (splits, rows): partial min/max per CTA, last CTA per row finishes (order independent => exact).
// row_bad (nullable): OR-ed with 1 when the row holds a NaN / Inf / |x| >= 2^67 element (the consumer resets it)
__global__ void __launch_bounds__(SSQ_THREADS)
row_minmax_kernel(const float* __restrict__ x, int64_t k, int64_t chunk, float* __restrict__ row_min,
                  float* __restrict__ row_max, unsigned int* tickets, float* partial, int* row_bad) {
    __shared__ float s_red[2 * 32];
    __shared__ bool is_last;
    const int64_t row = blockIdx.y;
    const float* grow = x + row * k;
    int64_t j0 = (int64_t)blockIdx.x * chunk, j1 = j0 + chunk < k ? j0 + chunk : k;
    float mn = INFINITY, mx = -INFINITY; bool bad = false;
    for (int64_t j = j0 + threadIdx.x; j < j1; j += blockDim.x) {
        float v = ld_stream1(grow + j); mn = fminf(mn, v); mx = fmaxf(mx, v); bad |= !(fabsf(v) < SSQ_DIV_XMAX);
    }
    if (row_bad && __any_sync(0xffffffffu, bad) && (threadIdx.x & 31) == 0) atomicOr(row_bad + row, 1);
    mn = warp_min(mn); mx = warp_max(mx);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    if (lane == 0) { s_red[warp] = mn; s_red[32 + warp] = mx; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < nwarp; ++w) { mn = fminf(mn, s_red[w]); mx = fmaxf(mx, s_red[32 + w]); }
        float* mine = partial + ((size_t)row * gridDim.x + blockIdx.x) * 2;
        mine[0] = mn; mine[1] = mx;
        __threadfence();
        is_last = atomicAdd(&tickets[row], 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    mn = INFINITY; mx = -INFINITY;
    for (int s = threadIdx.x; s < (int)gridDim.x; s += blockDim.x) {
        const float* pp = partial + ((size_t)row * gridDim.x + s) * 2;
        mn = fminf(mn, __ldcg(pp)); mx = fmaxf(mx, __ldcg(pp + 1));
    }
    mn = warp_min(mn); mx = warp_max(mx);
    __syncthreads();
    if (lane == 0) { s_red[warp] = mn; s_red[32 + warp] = mx; }
    __syncthreads();
    if (threadIdx.x == 0) {
        mn = s_red[0]; mx = s_red[32];
        for (int w = 1; w < nwarp; ++w) { mn = fminf(mn, s_red[w]); mx = fmaxf(mx, s_red[32 + w]); }
        row_min[row] = mn; row_max[row] = mx;
        tickets[row] = 0u;
    }
}

// ---- K2a tensor ----------------------------------------------------------------------------------------------------
// control block in the workspace, written by the rank launch and read by the settle launch
struct TensorCtl { int nsurv; int surv[NC]; };

constexpr int TE = 8;   // elements per thread per trip
// rank launch: every element read once and ranked against all 80 candidates; the last CTA lists the survivors
__global__ void __launch_bounds__(SSQ_THREADS)
mse_rank_tensor_kernel(const float* __restrict__ x, int64_t k, const float* __restrict__ mm /* min,max */,
                       const int* bad, float lm1, int symmetric, float p,
                       unsigned int* ticket, double* partial /* [grid][NC] */, TensorCtl* ctl) {
    __shared__ RankCand s_rc[NC];
    __shared__ int s_ok[NC];
    __shared__ double s_acc[SSQ_THREADS / 32][NC];
    __shared__ float s_rank[NC];
    __shared__ bool is_last;
    float mn = __ldg(mm), mx = __ldg(mm + 1);
    apply_sym(mn, mx, symmetric);
    const bool rank = needs_rank(p) && !*bad;
    if (!rank) {                                     // nothing to rank: every candidate goes to the settle launch
        if (blockIdx.x == 0 && threadIdx.x < NC) { ctl->surv[threadIdx.x] = threadIdx.x; if (threadIdx.x == 0) ctl->nsurv = NC; }
        return;
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    if (threadIdx.x < NC) {
        RankCand rc;
        s_ok[threadIdx.x] = rank_cand(make_cand(threadIdx.x, mn, mx, lm1), lm1, rc);
        s_rc[threadIdx.x] = rc;
    }
    for (int i = lane; i < NC; i += 32) s_acc[warp][i] = 0.0;
    __syncthreads();
    const int64_t tile = (int64_t)blockDim.x * TE;
    for (int64_t base = (int64_t)blockIdx.x * tile; base < k; base += (int64_t)gridDim.x * tile) {
        float xv[TE];
        int nvalid = 0;
#pragma unroll
        for (int e = 0; e < TE; ++e) {
            int64_t j = base + (int64_t)e * blockDim.x + threadIdx.x;
            xv[e] = j < k ? ld_stream1(x + j) : 0.f;
            nvalid += j < k;
        }
        for (int i = 0; i < NC; ++i) {
            const RankCand rc = s_rc[i];
            float s = 0.f;
            if (nvalid == TE) {
#pragma unroll
                for (int e = 0; e < TE; ++e) s += rank_err(xv[e], rc, p);
            } else {
#pragma unroll
                for (int e = 0; e < TE; ++e) if (e < nvalid) s += rank_err(xv[e], rc, p);
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
            if (lane == 0) s_acc[warp][i] += (double)s;
        }
    }
    __syncthreads();
    if (threadIdx.x < NC) {
        double t = 0.0;
        for (int w = 0; w < nwarp; ++w) t += s_acc[w][threadIdx.x];
        partial[(size_t)blockIdx.x * NC + threadIdx.x] = t;
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) is_last = atomicAdd(ticket, 1u) == gridDim.x - 1;
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    if (threadIdx.x < NC) {
        double t = 0.0;
        for (unsigned b = 0; b < gridDim.x; ++b) t += __ldcg(partial + (size_t)b * NC + threadIdx.x);
        s_rank[threadIdx.x] = s_ok[threadIdx.x] ? (float)(t / (double)k) : nanf("");
    }
    __syncthreads();
    if (warp == 0) {
        float lo = INFINITY;
        for (int i = lane; i < NC; i += 32) lo = fminf(lo, s_rank[i]);
        lo = warp_min(lo);
        const float thr = lo + lo * RANK_WINDOW;
        int base = 0;
        for (int i0 = 0; i0 < NC; i0 += 32) {
            const int i = i0 + lane;
            const bool keep = i < NC && !(s_rank[i] > thr);
            const unsigned m = __ballot_sync(0xffffffffu, keep);
            if (keep) ctl->surv[base + __popc(m & ((1u << lane) - 1u))] = i;
            base += __popc(m);
        }
        if (lane == 0) { ctl->nsurv = base; *ticket = 0u; }
    }
}

// settle launch: the survivors re-scored with the reference expression in a fixed order (per-thread 8-term fp32 partials,
// fp64 warp tree, per-CTA fp64 sums, CTAs summed in index order by the last one)
__global__ void __launch_bounds__(SSQ_THREADS)
mse_settle_tensor_kernel(const float* __restrict__ x, int64_t k, const float* __restrict__ mm,
                         float lm1, int symmetric, float p,
                         float* __restrict__ delta, float* __restrict__ zp, float* __restrict__ raw,
                         float* __restrict__ best_score, int32_t* __restrict__ best_index, int64_t out_row,
                         unsigned int* ticket, double* partial /* [grid][NC] */, const TensorCtl* __restrict__ ctl,
                         int* bad) {
    __shared__ float s_d[NC], s_z[NC];
    __shared__ int s_surv[NC];
    __shared__ double s_acc[SSQ_THREADS / 32][NC];
    __shared__ float s_scores[NC];
    __shared__ bool is_last;
    float mn = __ldg(mm), mx = __ldg(mm + 1);
    apply_sym(mn, mx, symmetric);
    const int ns = ctl->nsurv;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    if (threadIdx.x < ns) {
        const int i = ctl->surv[threadIdx.x];
        Cand c = make_cand(i, mn, mx, lm1);
        s_surv[threadIdx.x] = i; s_d[threadIdx.x] = c.d; s_z[threadIdx.x] = c.z;
    }
    for (int i = lane; i < NC; i += 32) s_acc[warp][i] = 0.0;
    __syncthreads();
    const int64_t tile = (int64_t)blockDim.x * TE;
    for (int64_t base = (int64_t)blockIdx.x * tile; base < k; base += (int64_t)gridDim.x * tile) {
        float xv[TE]; bool ok[TE];
#pragma unroll
        for (int e = 0; e < TE; ++e) {
            int64_t j = base + (int64_t)e * blockDim.x + threadIdx.x;
            ok[e] = j < k;
            xv[e] = ok[e] ? ld_stream1(x + j) : 0.f;
        }
        for (int si = 0; si < ns; ++si) {
            Cand c; c.d = s_d[si]; c.z = s_z[si]; c.nmin = 0.f;
            const Recip R = make_recip(c.d);
            float s = 0.f;
#pragma unroll
            for (int e = 0; e < TE; ++e) if (ok[e]) s += cand_err(xv[e], c, R, lm1, p);
            double ds = warp_sum((double)s);
            if (lane == 0) s_acc[warp][si] += ds;
        }
    }
    __syncthreads();
    if (threadIdx.x < ns) {
        double t = 0.0;
        for (int w = 0; w < nwarp; ++w) t += s_acc[w][threadIdx.x];
        partial[(size_t)blockIdx.x * NC + threadIdx.x] = t;
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) is_last = atomicAdd(ticket, 1u) == gridDim.x - 1;
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    if (threadIdx.x < ns) {
        double t = 0.0;
        for (unsigned b = 0; b < gridDim.x; ++b) t += __ldcg(partial + (size_t)b * NC + threadIdx.x);
        s_scores[threadIdx.x] = (float)(t / (double)k);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        float best = 1e10f; int idx = -1;
        for (int s = 0; s < ns; ++s) if (s_scores[s] < best) { best = s_scores[s]; idx = s_surv[s]; }
        write_winner(idx, best, mn, mx, lm1, symmetric, out_row, delta, zp, raw, best_score, best_index);
        *ticket = 0u;
        *bad = 0;
    }
}

// ---- K2b ---------------------------------------------------------------------------------------------------
// Reference predicate (channelQuantMSE.py:91-102) for column j and candidate c:  every row r has
//     lo < g_r(fl(w[r][j] / c)) < hi,   g_r(v) = fl(fl(fl(v / delta_r) + zero_r) / x_range),  zero_r = rint(raw_zp_r / delta_r),
// and inp_scale[j] = the LAST fitting candidate of the descending list cand[0] > cand[1] > ... .
// Every step of g_r is a correctly rounded, weakly increasing function of v, so { v : lo < g_r(v) < hi } is an interval
// [vlo_r, vhi_r] of floats — found ONCE per row by bisection on the float ordering with the exact ops (row_interval).
// fl(w / c) is weakly monotone in c, so when the interval contains 0 (it does whenever 0 <= zero_r <= x_range, lo < 0 < 1 < hi:
// every quantiser this repo or the reference builds) an element fits exactly the candidates c >= c*(w, r): the set of fitting
// indices is a prefix 0..J, J = floor(jr), jr = level (1 - t), t = |w| / |V|, V = vhi_r for w > 0, vlo_r for w < 0, up to
// rounding. The column answer is the MINIMUM prefix over its rows, and jr is monotone in t, so the sweep over the weights only
// has to find max_r t per column: one select, one multiply and one integer max per element (t >= 0, so the IEEE bit patterns
// order like integers and the canonical NaN 0x7fffffff is the largest of all) — 4 B/element and nothing else on the hot path.
// The finish kernel turns max t into the prefix: decisive when jr is further than `margin` (>= 2.5x its error bound) from an
// integer — then every other element of the column has a prefix >= that one, see DESIGN.md K2b — and otherwise (a 2*margin
// fraction of the columns) the warp re-reads that column and settles the elements within 3*margin of the maximum with the exact
// predicate on the neighbouring candidates. Whenever the preconditions fail (flag set by row_interval) the brute-force kernel
// below evaluates all `level` candidates instead.
__device__ __forceinline__ float g_row(float v, float d, float zero, float x_range) {
    return __fdiv_rn(__fadd_rn(__fdiv_rn(v, d), zero), x_range);
}
// order-preserving map float <-> uint32
__device__ __forceinline__ uint32_t f2ord(float f) { uint32_t u = __float_as_uint(f); return (u & 0x80000000u) ? ~u : (u | 0x80000000u); }
__device__ __forceinline__ float ord2f(uint32_t o) { return __uint_as_float((o & 0x80000000u) ? (o & 0x7fffffffu) : ~o); }

struct RowIv { float vlo, vhi, rlo, rhi; };     // interval of v = fl(w/c) that fits; rlo = 1/|vlo|, rhi = 1/vhi (both > 0)
// One WARP per row: the two interval ends are found by a 33-ary search over the ordered float encoding — every round the 32
// lanes probe 32 points of the bracket at once and a ballot picks the sub-bracket (7 rounds instead of the 31 dependent steps of
// a bisection; g_row is two IEEE divisions and an add, ~150 cycles of latency per evaluation).
__global__ void __launch_bounds__(128)
inp_scale_row_interval_kernel(const float* __restrict__ delta, const float* __restrict__ raw_zp, float x_range,
                              float lo, float hi, int64_t oc, int level, const float* __restrict__ cand,
                              RowIv* __restrict__ iv, int* __restrict__ need_brute, int epoch,
                              int* __restrict__ best, int* __restrict__ last_fit, int64_t k) {
    const int64_t gtid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    if (gtid < k) { best[gtid] = 0; last_fit[gtid] = 0; }           // the per-column slots of this call (the grid covers k threads)
    {   // the estimate assumes the reference's list cand[j] = fp32((level - j) / level) (channelQuantMSE.py:79)
        bool okc = true;
        const int64_t nthr = (int64_t)gridDim.x * blockDim.x;
        for (int64_t j = gtid; j < level; j += nthr) okc &= cand[j] == (float)((double)(level - j) / (double)level);
        if (gtid == 0) okc = okc && (lo < 0.f) && (hi > 1.0f) && (x_range >= 1.0f);
        if (!okc) atomicExch(need_brute, epoch);
    }
    const int64_t r = gtid >> 5;
    if (r >= oc) return;                                               // whole warps leave together
    const float d = delta[r];
    const float zero = rintf(__fdiv_rn(raw_zp[r], d));
    bool ok = mid_exponent(d) && d > 0.f && zero >= 0.f && zero <= x_range;
    const float g0 = g_row(0.f, d, zero, x_range);
    ok = ok && g0 > lo && g0 < hi;
    RowIv out = {0.f, 0.f, 0.f, 0.f};
    if (ok) {                                                          // uniform across the warp
        // probe point of this lane inside the open bracket (x, y), y - x > 1: strictly between, non-decreasing in the lane
        auto probe = [&](uint32_t x, uint32_t y) {
            const uint32_t step = (uint32_t)(((uint64_t)(y - x) * (uint32_t)(lane + 1)) / 33u);
            return x + (step ? step : 1u);
        };
        // largest v >= 0 with g(v) < hi; invariant: g(a) < hi, and g(b) >= hi unless b is the top of the range
        uint32_t a = f2ord(0.f), b = f2ord(3.402823466e+38f);
        if (g_row(ord2f(b), d, zero, x_range) < hi) a = b;
        while (b - a > 1u && a != b) {
            const uint32_t m = probe(a, b);
            const unsigned below = __ballot_sync(0xffffffffu, g_row(ord2f(m), d, zero, x_range) < hi);   // a prefix of the lanes (g is monotone)
            const int t = __popc(below);
            const uint32_t m_lo = __shfl_sync(0xffffffffu, m, t > 0 ? t - 1 : 0), m_hi = __shfl_sync(0xffffffffu, m, t < 32 ? t : 31);
            if (t > 0) a = m_lo;
            if (t < 32) b = m_hi;
        }
        out.vhi = ord2f(a);
        // smallest v <= 0 with g(v) > lo; invariant: g(e0) > lo, and g(c0) <= lo unless c0 is the bottom of the range
        uint32_t c0 = f2ord(-3.402823466e+38f), e0 = f2ord(-0.f);
        if (g_row(ord2f(c0), d, zero, x_range) > lo) e0 = c0;
        while (e0 - c0 > 1u && e0 != c0) {
            const uint32_t m = probe(c0, e0);
            const unsigned above = __ballot_sync(0xffffffffu, g_row(ord2f(m), d, zero, x_range) > lo);   // a suffix of the lanes
            const int t = 32 - __popc(above);                          // lanes 0..t-1 are not above
            const uint32_t m_lo = __shfl_sync(0xffffffffu, m, t > 0 ? t - 1 : 0), m_hi = __shfl_sync(0xffffffffu, m, t < 32 ? t : 31);
            if (t > 0) c0 = m_lo;
            if (t < 32) e0 = m_hi;
        }
        out.vlo = ord2f(e0);
        ok = mid_exponent(out.vhi) && out.vhi > 0.f && mid_exponent(out.vlo) && out.vlo < 0.f;
        if (ok) { out.rhi = make_recip(out.vhi).r; out.rlo = make_recip(-out.vlo).r; }
    }
    if (lane == 0) {
        iv[r] = out;
        if (!ok) atomicExch(need_brute, epoch);
    }
}

// exact predicate for one element and candidate index j
__device__ __forceinline__ bool elem_fits(float w, const RowIv& I, const float* __restrict__ cand, int j) {
    const float v = div_exact(w, __ldg(cand + j));
    return v >= I.vlo && v <= I.vhi;
}
// J+1 for one element whose estimate jr is NOT decisive (within `margin` of an integer, or NaN / Inf / huge): settle with the exact
// predicate; the prefix property makes a local walk sufficient. Rare (a 2*margin fraction of the elements that reach it), out of line.
__device__ __noinline__ int elem_prefix_settle(float w, float jr, const RowIv& I, const float* __restrict__ cand, int level) {
    const float flevel = (float)level;
    int j = jr < 0.f ? 0 : (jr >= flevel ? level - 1 : (int)floorf(jr));
    if (!(jr == jr)) j = level - 1;                                    // NaN jr -> level-1, walks down to "none"
    j = min(j + 1, level - 1);
    while (j >= 0 && !elem_fits(w, I, cand, j)) --j;
    while (j + 1 < level && elem_fits(w, I, cand, j + 1)) ++j;
    return j + 1;
}

constexpr int K2B_COLS = 4;                     // columns per thread (one float4 when aligned)
constexpr int K2B_ROWS = 8;                     // rows of loads in flight per thread
constexpr int K2B_CTAS = 6;                     // resident CTAs per SM the launch bounds allow
__device__ __forceinline__ float k2b_t(float x, float rlo, float rhi) { return fabsf(x) * (x > 0.f ? rhi : rlo); }   // >= +0, or NaN
// sweep: tmax[col] = max over the slab's rows of t, as int bit patterns (tmax zeroed by the reset kernel)
__global__ void __launch_bounds__(SSQ_THREADS, K2B_CTAS)
inp_scale_sweep_kernel(const float* __restrict__ w, const RowIv* __restrict__ iv, int64_t oc, int64_t k, int64_t rows_per_cta,
                       const int* __restrict__ need_brute, int epoch, int force_brute, int* __restrict__ tmax) {
    if (force_brute || __ldg(need_brute) == epoch) return;
    const int64_t col0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * K2B_COLS;
    if (col0 >= k) return;
    const int64_t r0 = (int64_t)blockIdx.y * rows_per_cta;
    const int64_t r1 = r0 + rows_per_cta < oc ? r0 + rows_per_cta : oc;
    const bool vec = (k % K2B_COLS == 0) && aligned16(w) && col0 + K2B_COLS <= k;
    int tm[K2B_COLS];
#pragma unroll
    for (int e = 0; e < K2B_COLS; ++e) tm[e] = 0;
    int64_t r = r0;
    if (vec) {
        for (; r + K2B_ROWS <= r1; r += K2B_ROWS) {
            float4 x[K2B_ROWS];
#pragma unroll
            for (int u = 0; u < K2B_ROWS; ++u) x[u] = ld_stream4(w + (r + u) * k + col0);
#pragma unroll
            for (int u = 0; u < K2B_ROWS; ++u) {
                const float2 s = __ldg(reinterpret_cast<const float2*>(&iv[r + u].rlo));
                tm[0] = max(tm[0], __float_as_int(k2b_t(x[u].x, s.x, s.y)));
                tm[1] = max(tm[1], __float_as_int(k2b_t(x[u].y, s.x, s.y)));
                tm[2] = max(tm[2], __float_as_int(k2b_t(x[u].z, s.x, s.y)));
                tm[3] = max(tm[3], __float_as_int(k2b_t(x[u].w, s.x, s.y)));
            }
        }
    }
    for (; r < r1; ++r) {
        const float2 s = __ldg(reinterpret_cast<const float2*>(&iv[r].rlo));
#pragma unroll
        for (int e = 0; e < K2B_COLS; ++e)
            if (col0 + e < k) tm[e] = max(tm[e], __float_as_int(k2b_t(ld_stream1(w + r * k + col0 + e), s.x, s.y)));
    }
#pragma unroll
    for (int e = 0; e < K2B_COLS; ++e)
        if (col0 + e < k && tm[e] != 0) atomicMax(tmax + col0 + e, tm[e]);
}
// finish: one CTA per 256 columns. best[col] holds max t; the prefix length p picks inp_scale[col] = cand[p - 1]. Columns whose
// estimate is not decisive are settled by the WHOLE CTA, one after the other: the column is re-read (strided, 4 rows of loads in
// flight per thread — a lone warp walking 4096 rows one dependent load at a time took longer than the sweep itself).
__global__ void __launch_bounds__(SSQ_THREADS)
inp_scale_finish_kernel(const float* __restrict__ w, const RowIv* __restrict__ iv, const float* __restrict__ cand, int level,
                        int64_t oc, int64_t k, const int* __restrict__ need_brute, int epoch, int force_brute,
                        const int* __restrict__ best, float* __restrict__ inp_scale) {
    if (force_brute || __ldg(need_brute) == epoch) return;
    __shared__ int s_n;
    __shared__ int s_col[SSQ_THREADS], s_min[SSQ_THREADS];
    const int tid = threadIdx.x, lane = tid & 31;
    const int64_t col_base = (int64_t)blockIdx.x * SSQ_THREADS, col = col_base + tid;
    const bool valid = col < k;
    const float flevel = (float)level;
    const float margin = fmaf(flevel, 1e-6f, 1e-6f);                   // >= 2.5x the error bound of jr
    if (tid == 0) s_n = 0;
    __syncthreads();
    const float tm = valid ? __int_as_float(best[col]) : 0.f;
    const float jr = fmaf(-flevel, tm, flevel);                        // candidates 0..floor(jr) fit the column's tightest element
    const float rn = rintf(jr);
    const bool none = !(tm < 1e30f);                                   // NaN / Inf / huge weight in the column: nothing fits
    int p = none ? 0 : max(min(__float2int_rd(jr) + 1, level), 0);
    // not decisive: jr within `margin` of an integer n in [0, level) — the exact prefix is n or n + 1 (n >= level gives `level`
    // either way: all-zero columns land there; n < 0 gives 0 either way)
    const bool settle = valid && !none && !(fabsf(jr - rn) > margin) && rn >= 0.f && rn < flevel;
    int slot = -1;
    if (settle) { slot = atomicAdd(&s_n, 1); s_col[slot] = tid; s_min[slot] = level; }
    __syncthreads();
    const int n_settle = s_n;
    for (int i = 0; i < n_settle; ++i) {
        const int64_t c = col_base + s_col[i];
        const float near = fmaf(-flevel, __int_as_float(best[c]), flevel) + 3.0f * margin;
        int local = level;
        for (int64_t r0 = tid; r0 < oc; r0 += 4 * SSQ_THREADS) {
            float x[4]; RowIv I[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int64_t r = r0 + (int64_t)u * SSQ_THREADS;
                x[u] = r < oc ? w[r * k + c] : 0.f;
                I[u] = iv[r < oc ? r : 0];
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const float je = fmaf(-flevel, k2b_t(x[u], I[u].rlo, I[u].rhi), flevel);
                if (r0 + (int64_t)u * SSQ_THREADS < oc && !(je > near)) local = min(local, elem_prefix_settle(x[u], je, I[u], cand, level));
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) local = min(local, __shfl_xor_sync(0xffffffffu, local, o));
        if (lane == 0 && local < level) atomicMin(&s_min[i], local);
    }
    __syncthreads();
    if (settle) p = s_min[slot];
    if (valid && p > 0) inp_scale[col] = __ldg(cand + p - 1);          // the LAST fitting candidate; none fits: left as it is
}

// brute force: every (column, candidate) with the reference expression; runs only when need_brute is set (or forced)
constexpr int CH = 8;   // candidates per thread
__global__ void __launch_bounds__(SSQ_THREADS)
inp_scale_fit_kernel(const float* __restrict__ w, const float* __restrict__ delta, const float* __restrict__ raw_zp,
                     const float* __restrict__ cand, int level, float x_range, float lo, float hi,
                     int64_t oc, int64_t k, int* __restrict__ need_brute, int epoch, int force_brute, unsigned int* __restrict__ ticket,
                     int* __restrict__ last_fit /* [k], 0 = none */, float* __restrict__ inp_scale) {
    if (!(force_brute || __ldg(need_brute) == epoch)) return;
    __shared__ bool s_last;
    // grid-stride over (column block, candidate block) pairs: the grid is capped so that the usual, idle launch costs a few
    // hundred empty CTAs instead of one per pair (18 432 of them at level 1024 took 13.6 us to do nothing)
    const int64_t colblocks = (k + SSQ_THREADS - 1) / SSQ_THREADS, candblocks = (level + CH - 1) / CH;
    for (int64_t job = blockIdx.x; job < colblocks * candblocks; job += gridDim.x) {
        const int64_t col = (job % colblocks) * SSQ_THREADS + threadIdx.x;
        const int j0 = (int)(job / colblocks) * CH;
        if (col >= k) continue;
        float c[CH]; bool fit[CH];
#pragma unroll
        for (int e = 0; e < CH; ++e) { c[e] = (j0 + e < level) ? __ldg(cand + j0 + e) : 1.f; fit[e] = (j0 + e < level); }
        for (int64_t r = 0; r < oc; ++r) {
            float d = __ldg(delta + r);
            float zero = rintf(__fdiv_rn(__ldg(raw_zp + r), d));
            float xv = w[r * k + col];
#pragma unroll
            for (int e = 0; e < CH; ++e) {
                float u = __fdiv_rn(__fadd_rn(__fdiv_rn(__fdiv_rn(xv, c[e]), d), zero), x_range);
                fit[e] = fit[e] && (u > lo) && (u < hi);
            }
        }
        int last = 0;
#pragma unroll
        for (int e = 0; e < CH; ++e) if (fit[e]) last = j0 + e + 1;
        if (last > 0) atomicMax(last_fit + col, last);
    }
    // the last CTA to finish turns last_fit into inp_scale and leaves the flag clean for the next call on this workspace
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        s_last = atomicAdd(ticket, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    for (int64_t col = threadIdx.x; col < k; col += blockDim.x) {
        const int b = __ldcg(last_fit + col);
        if (b > 0) inp_scale[col] = __ldg(cand + b - 1);
    }
    if (threadIdx.x == 0) { *ticket = 0u; *need_brute = 0; }
}
constexpr int64_t ROW_SMEM_MAX_ELEMS = 48 * 1024;  // 192 KB of the 227 KB a CTA may use
constexpr int TENSOR_GRID = SSQ_NUM_SMS * 4;

}  // namespace ssq

using namespace ssq;

extern "C" size_t ssq_mse_scale_search_ws_bytes(int64_t rows, int64_t k) {
    (void)rows; (void)k;
    // fixed ticket header + min/max/bad slot + control block + per-CTA min/max partials + per-CTA candidate partials
    return ws_ticket_bytes(1) + 256 + 512 + (size_t)SSQ_NUM_SMS * SSQ_CTAS_PER_SM * 2 * sizeof(float) + (size_t)TENSOR_GRID * NC * sizeof(double);
}

static int minmax_launch(const float* x, int64_t rows, int64_t k, float* row_min, float* row_max,
                         unsigned int* tickets, float* partial, int* row_bad, cudaStream_t st) {
    int64_t cap = (int64_t)SSQ_NUM_SMS * SSQ_CTAS_PER_SM;
    int64_t want = (cap + rows - 1) / rows;
    int64_t by_work = (k + SSQ_THREADS * 8 - 1) / (SSQ_THREADS * 8);
    int64_t s = want < by_work ? want : by_work; if (s < 1) s = 1;
    int64_t chunk = (k + s - 1) / s;
    int nsplit = (int)((k + chunk - 1) / chunk);
    row_minmax_kernel<<<dim3(nsplit, (unsigned)rows), SSQ_THREADS, 0, st>>>(x, k, chunk, row_min, row_max, tickets, partial, row_bad);
    return launch_status();
}

extern "C" int ssq_row_minmax(const float* x, int64_t rows, int64_t k, float* row_min, float* row_max,
                              void* ws, size_t ws_bytes, void* stream) {
    if (rows == 0) return SSQ_OK;
    if (!x || !row_min || !row_max) return SSQ_ERR_NULL;
    if (rows < 0 || k <= 0 || rows > 65535) return SSQ_ERR_SIZE;
    if (!ws || ws_bytes < ssq_ws_bytes(rows)) return SSQ_ERR_WORKSPACE;
    WsView v = ws_view(ws, rows);
    return minmax_launch(x, rows, k, row_min, row_max, v.tickets, reinterpret_cast<float*>(v.partials), nullptr, (cudaStream_t)stream);
}

extern "C" int ssq_mse_scale_search(const float* x, int64_t rows, int64_t k, int n_levels, int symmetric,
                                    float p_norm, float* delta, float* zero_point, float* raw_zero_point,
                                    float* best_score, int32_t* best_index,
                                    void* ws, size_t ws_bytes, void* stream) {
    if (rows == 0) return SSQ_OK;
    if (!x || !delta || !zero_point || !raw_zero_point) return SSQ_ERR_NULL;
    if (rows < 0 || k <= 0 || n_levels < 2 || n_levels > 256) return SSQ_ERR_SIZE;
    cudaStream_t st = (cudaStream_t)stream;
    const float lm1 = (float)(n_levels - 1);
    if (k < SHORT_ROW) {
        const unsigned grid = (unsigned)((rows + SHORT_WARPS - 1) / SHORT_WARPS);
        mse_search_short_rows_kernel<<<grid, SSQ_THREADS, 0, st>>>(x, rows, (int)k, lm1, symmetric, p_norm, delta, zero_point,
                                                                  raw_zero_point, best_score, best_index);
        return launch_status();
    }
    if (k <= ROW_SMEM_MAX_ELEMS) {
        size_t smem = (size_t)((k * 4 + 127) / 128) * 128;
        static bool attr_set = false;
        if (!attr_set) {
            cudaError_t e = cudaFuncSetAttribute(mse_search_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                 (int)(ROW_SMEM_MAX_ELEMS * 4));
            if (e != cudaSuccess) return (int)e;
            attr_set = true;
        }
        mse_search_rows_kernel<<<(unsigned)rows, SSQ_THREADS, smem, st>>>(x, k, lm1, symmetric, p_norm, delta, zero_point,
                                                                          raw_zero_point, best_score, best_index);
        return launch_status();
    }
    if (!ws || ws_bytes < ssq_mse_scale_search_ws_bytes(rows, k)) return SSQ_ERR_WORKSPACE;
    unsigned int* tickets = reinterpret_cast<unsigned int*>(ws);          // [0] minmax, [1] rank, [2] settle (shared header)
    char* base = reinterpret_cast<char*>(ws) + ws_ticket_bytes(1);
    float* mm = reinterpret_cast<float*>(base);
    int* bad = reinterpret_cast<int*>(base + 16);
    TensorCtl* ctl = reinterpret_cast<TensorCtl*>(base + 256);
    float* mm_partial = reinterpret_cast<float*>(base + 256 + 512);
    double* partial = reinterpret_cast<double*>(base + 256 + 512 + (size_t)SSQ_NUM_SMS * SSQ_CTAS_PER_SM * 2 * sizeof(float));
    static_assert(sizeof(TensorCtl) <= 512, "control block");
    for (int64_t r = 0; r < rows; ++r) {
        const float* xr = x + r * k;
        int e = minmax_launch(xr, 1, k, mm, mm + 1, tickets, mm_partial, bad, st);
        if (e) return e;
        int64_t tiles = (k + (int64_t)SSQ_THREADS * TE - 1) / ((int64_t)SSQ_THREADS * TE);
        int grid = (int)(tiles < TENSOR_GRID ? tiles : TENSOR_GRID);
        mse_rank_tensor_kernel<<<grid, SSQ_THREADS, 0, st>>>(xr, k, mm, bad, lm1, symmetric, p_norm, tickets + 1, partial, ctl);
        e = launch_status();
        if (e) return e;
        mse_settle_tensor_kernel<<<grid, SSQ_THREADS, 0, st>>>(xr, k, mm, lm1, symmetric, p_norm, delta, zero_point,
                                                               raw_zero_point, best_score, best_index, r, tickets + 2, partial, ctl, bad);
        e = launch_status();
        if (e) return e;
    }
    return SSQ_OK;
}

// workspace: [header 64 B: brute-force epoch flag, ticket][best (max t) : k ints][last_fit : k ints][RowIv : oc]
extern "C" size_t ssq_inp_scale_search_ws_bytes2(int64_t oc, int64_t k) {
    const size_t kk = (size_t)(k > 0 ? k : 1), rr = (size_t)(oc > 0 ? oc : 1);
    return 64 + 2 * ((kk * sizeof(int) + 15) / 16 * 16) + rr * sizeof(RowIv);
}
extern "C" size_t ssq_inp_scale_search_ws_bytes(int64_t k) { return ssq_inp_scale_search_ws_bytes2(65536, k); }

extern "C" int ssq_inp_scale_search_ex(const float* w, const float* delta, const float* raw_zero_point,
                                       const float* cand, int level, float x_range, float lo, float hi,
                                       float* inp_scale, int64_t oc, int64_t k, int force_brute,
                                       void* ws, size_t ws_bytes, void* stream) {
    if (oc == 0 || k == 0 || level == 0) return SSQ_OK;
    if (!w || !delta || !raw_zero_point || !cand || !inp_scale) return SSQ_ERR_NULL;
    if (oc < 0 || k < 0 || level < 0 || level > 65535 * CH) return SSQ_ERR_SIZE;
    if (!ws || ws_bytes < ssq_inp_scale_search_ws_bytes2(oc, k)) return SSQ_ERR_WORKSPACE;
    cudaStream_t st = (cudaStream_t)stream;
    char* base = reinterpret_cast<char*>(ws);
    int* need_brute = reinterpret_cast<int*>(base);
    const size_t karr = ((size_t)k * sizeof(int) + 15) / 16 * 16;
    int* best = reinterpret_cast<int*>(base + 64);
    int* last_fit = reinterpret_cast<int*>(base + 64 + karr);
    RowIv* iv = reinterpret_cast<RowIv*>(base + 64 + 2 * karr);
    unsigned int* ticket = reinterpret_cast<unsigned int*>(base + 4);
    // "this call needs the brute force" is the flag holding THIS call's epoch: nothing has to be cleared before use, and a stale
    // value left by anything else never matches (0 is never an epoch; the brute-force kernel writes 0 back when it is done)
    static std::atomic<int> s_epoch{0};
    int epoch = ++s_epoch;
    if (epoch <= 0) { s_epoch = 1; epoch = 1; }
    const int force = force_brute ? 1 : 0;
    // launch 1: per-row intervals (one warp per row) + this call's per-column slots zeroed
    int64_t prep = (oc + 3) / 4;
    if (prep < (k + 127) / 128) prep = (k + 127) / 128;
    inp_scale_row_interval_kernel<<<(unsigned)prep, 128, 0, st>>>(delta, raw_zero_point, x_range, lo, hi, oc, level, cand, iv, need_brute, epoch,
                                                                   best, last_fit, k);
    int e = launch_status();
    if (e) return e;
    // launch 2, sweep: column blocks x row slabs = ONE wave of resident CTAs (a second, partial wave costs as much as a full one)
    const int64_t colblocks = (k + (int64_t)SSQ_THREADS * K2B_COLS - 1) / ((int64_t)SSQ_THREADS * K2B_COLS);
    int64_t slabs = ((int64_t)SSQ_NUM_SMS * K2B_CTAS) / colblocks;
    if (slabs > (oc + 2 * K2B_ROWS - 1) / (2 * K2B_ROWS)) slabs = (oc + 2 * K2B_ROWS - 1) / (2 * K2B_ROWS);
    if (slabs > 65535) slabs = 65535;
    if (slabs < 1) slabs = 1;
    const int64_t rows_per_cta = (oc + slabs - 1) / slabs;
    slabs = (oc + rows_per_cta - 1) / rows_per_cta;
    inp_scale_sweep_kernel<<<dim3((unsigned)colblocks, (unsigned)slabs), SSQ_THREADS, 0, st>>>(w, iv, oc, k, rows_per_cta, need_brute, epoch, force, best);
    e = launch_status();
    if (e) return e;
    // launch 3: prefix per column -> inp_scale
    const unsigned kgrid = (unsigned)((k + SSQ_THREADS - 1) / SSQ_THREADS);
    inp_scale_finish_kernel<<<kgrid, SSQ_THREADS, 0, st>>>(w, iv, cand, level, oc, k, need_brute, epoch, force, best, inp_scale);
    e = launch_status();
    if (e) return e;
    // launch 4: the brute force, idle (a few hundred empty CTAs) unless the flag carries this epoch or the caller forces it
    int64_t fit_jobs = (int64_t)kgrid * ((level + CH - 1) / CH);
    if (fit_jobs > (int64_t)SSQ_NUM_SMS * 8) fit_jobs = (int64_t)SSQ_NUM_SMS * 8;
    inp_scale_fit_kernel<<<(unsigned)fit_jobs, SSQ_THREADS, 0, st>>>(w, delta, raw_zero_point, cand, level, x_range, lo, hi, oc, k, need_brute, epoch, force,
                                                                     ticket, last_fit, inp_scale);
    return launch_status();
}

extern "C" int ssq_inp_scale_search(const float* w, const float* delta, const float* raw_zero_point,
                                    const float* cand, int level, float x_range, float lo, float hi,
                                    float* inp_scale, int64_t oc, int64_t k,
                                    void* ws, size_t ws_bytes, void* stream) {
    return ssq_inp_scale_search_ex(w, delta, raw_zero_point, cand, level, x_range, lo, hi, inp_scale, oc, k, 0, ws, ws_bytes, stream);
}
