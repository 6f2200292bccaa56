// K2 — scale searches.
//  K2a  per-channel / per-tensor MSE clip-ratio search: quant/quant_layer.py:145-162, :168-175.
//       Row variant: one CTA per output channel; the weight row is staged ONCE into shared memory
//       with a TMA bulk copy (cp.async.bulk + mbarrier) and all 80 candidates are scored from it,
//       one warp per candidate, shuffle-reduced. Tensor variant (activations, up to ~50 M elements):
//       grid-wide, every element read once from HBM and scored against all 80 candidates.
//       Both are powf-bound (80 x |d|^2.4 per 4 bytes), not HBM-bound — see DESIGN.md.
//  K2b  ChannelQuantMSE input-scale search: quant/channelQuantMSE.py:70-110.
#include "ssq_common.cuh"

namespace ssq {

constexpr int NC = SSQ_N_CANDIDATES;

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
__device__ __forceinline__ float cand_err(float x, const Cand& c, float lm1, float p) {
    float q = clampq(__fadd_rn(rintf(div_exact(x, c.d)), c.z), 0.f, lm1);
    float xq = __fmul_rn(__fsub_rn(q, c.z), c.d);
    return pow_scalar_accurate(fabsf(__fsub_rn(x, xq)), p);
}
__device__ __forceinline__ void apply_sym(float& x_min, float& x_max, int symmetric) {
    if (symmetric) {
        float am = fmaxf(fabsf(x_min), x_max);
        x_min = (x_min < 0.f) ? -am : 0.f;
        x_max = am;
    }
}
// first strict minimum below 1e10, then the winning candidate's outputs (quant_layer.py:157-162)
__device__ __forceinline__ void pick_and_write(const float* scores, float x_min, float x_max, float lm1, int symmetric,
                                               int64_t row, float* delta, float* zp, float* raw, float* best_score,
                                               int32_t* best_index) {
    float best = 1e10f; int idx = -1;
    for (int i = 0; i < NC; ++i) if (scores[i] < best) { best = scores[i]; idx = i; }
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

// ---- K2a rows ------------------------------------------------------------------------------------
__global__ void __launch_bounds__(SSQ_THREADS)
mse_search_rows_kernel(const float* __restrict__ x, int64_t k, float lm1, int symmetric, float p,
                       float* __restrict__ delta, float* __restrict__ zp, float* __restrict__ raw,
                       float* __restrict__ best_score, int32_t* __restrict__ best_index) {
    extern __shared__ __align__(128) unsigned char dyn[];
    float* srow = reinterpret_cast<float*>(dyn);
    __shared__ uint64_t bar;
    __shared__ float s_scores[NC];
    __shared__ float s_red[2 * 32];
    const int64_t row = blockIdx.x;
    const float* grow = x + row * k;
    const uint32_t bytes = (uint32_t)(k * 4);
    if ((bytes & 15u) == 0 && aligned16(grow)) {
        tma_row_to_smem(srow, grow, bytes, &bar);
    } else {
        for (int64_t j = threadIdx.x; j < k; j += blockDim.x) srow[j] = ld_stream1(grow + j);
        __syncthreads();
    }
    // row min / max (exact, order independent)
    float mn = INFINITY, mx = -INFINITY;
    for (int64_t j = threadIdx.x; j < k; j += blockDim.x) { float v = srow[j]; mn = fminf(mn, v); mx = fmaxf(mx, v); }
    mn = warp_min(mn); mx = warp_max(mx);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    if (lane == 0) { s_red[warp] = mn; s_red[32 + warp] = mx; }
    __syncthreads();
    mn = s_red[0]; mx = s_red[32];
    for (int w = 1; w < nwarp; ++w) { mn = fminf(mn, s_red[w]); mx = fmaxf(mx, s_red[32 + w]); }
    apply_sym(mn, mx, symmetric);
    // one warp per candidate
    for (int i = warp; i < NC; i += nwarp) {
        Cand c = make_cand(i, mn, mx, lm1);
        double acc = 0.0;
        for (int64_t j0 = lane; j0 < k; j0 += 32 * 8) {
            float s = 0.f;
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                int64_t j = j0 + e * 32;
                if (j < k) s += cand_err(srow[j], c, lm1, p);
            }
            acc += (double)s;
        }
        acc = warp_sum(acc);
        if (lane == 0) s_scores[i] = (float)(acc / (double)k);
    }
    __syncthreads();
    if (threadIdx.x == 0) pick_and_write(s_scores, mn, mx, lm1, symmetric, row, delta, zp, raw, best_score, best_index);
}

// ---- min / max, grid-wide -------------------------------------------------------------------------
// grid (splits, rows): partial min/max per CTA, last CTA per row finishes (order independent => exact).
__global__ void __launch_bounds__(SSQ_THREADS)
row_minmax_kernel(const float* __restrict__ x, int64_t k, int64_t chunk, float* __restrict__ row_min,
                  float* __restrict__ row_max, unsigned int* tickets, float* partial) {
    __shared__ float s_red[2 * 32];
    __shared__ bool is_last;
    const int64_t row = blockIdx.y;
    const float* grow = x + row * k;
    int64_t j0 = (int64_t)blockIdx.x * chunk, j1 = j0 + chunk < k ? j0 + chunk : k;
    float mn = INFINITY, mx = -INFINITY;
    for (int64_t j = j0 + threadIdx.x; j < j1; j += blockDim.x) { float v = ld_stream1(grow + j); mn = fminf(mn, v); mx = fmaxf(mx, v); }
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

// ---- K2a tensor: every element scored against all 80 candidates, read once ----------------------------
constexpr int TE = 8;   // elements per thread per trip
__global__ void __launch_bounds__(SSQ_THREADS)
mse_search_tensor_kernel(const float* __restrict__ x, int64_t k, const float* __restrict__ mm /* min,max */,
                         float lm1, int symmetric, float p,
                         float* __restrict__ delta, float* __restrict__ zp, float* __restrict__ raw,
                         float* __restrict__ best_score, int32_t* __restrict__ best_index, int64_t out_row,
                         unsigned int* ticket, double* partial /* [grid][NC] */) {
    __shared__ float s_d[NC], s_z[NC];
    __shared__ double s_acc[SSQ_THREADS / 32][NC];
    __shared__ float s_scores[NC];
    __shared__ bool is_last;
    float mn = __ldg(mm), mx = __ldg(mm + 1);
    apply_sym(mn, mx, symmetric);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    if (threadIdx.x < NC) { Cand c = make_cand(threadIdx.x, mn, mx, lm1); s_d[threadIdx.x] = c.d; s_z[threadIdx.x] = c.z; }
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
        for (int i = 0; i < NC; ++i) {
            Cand c; c.d = s_d[i]; c.z = s_z[i]; c.nmin = 0.f;
            float s = 0.f;
#pragma unroll
            for (int e = 0; e < TE; ++e) if (ok[e]) s += cand_err(xv[e], c, lm1, p);
            double ds = warp_sum((double)s);
            if (lane == 0) s_acc[warp][i] += ds;
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
        s_scores[threadIdx.x] = (float)(t / (double)k);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        pick_and_write(s_scores, mn, mx, lm1, symmetric, out_row, delta, zp, raw, best_score, best_index);
        *ticket = 0u;
    }
}

// ---- K2b ---------------------------------------------------------------------------------------------------
constexpr int CH = 8;   // candidates per thread
__global__ void __launch_bounds__(SSQ_THREADS)
inp_scale_fit_kernel(const float* __restrict__ w, const float* __restrict__ delta, const float* __restrict__ raw_zp,
                     const float* __restrict__ cand, int level, float x_range, float lo, float hi,
                     int64_t oc, int64_t k, int* __restrict__ best /* [k], 0 = none */) {
    const int64_t col = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int j0 = blockIdx.y * CH;
    if (col >= k) return;
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
    if (last > 0) atomicMax(best + col, last);
}
__global__ void inp_scale_pick_kernel(const float* __restrict__ cand, int* __restrict__ best, float* __restrict__ inp_scale, int64_t k) {
    int64_t col = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (col >= k) return;
    int b = best[col];
    if (b > 0) inp_scale[col] = __ldg(cand + b - 1);
    best[col] = 0;
}

constexpr int64_t ROW_SMEM_MAX_ELEMS = 48 * 1024;  // 192 KB of the 227 KB a CTA may use
constexpr int TENSOR_GRID = SSQ_NUM_SMS * 4;

}  // namespace ssq

using namespace ssq;

extern "C" size_t ssq_mse_scale_search_ws_bytes(int64_t rows, int64_t k) {
    (void)rows; (void)k;
    // fixed ticket header + min/max slot + per-CTA min/max partials + per-CTA candidate partials
    return ws_ticket_bytes(1) + 256 + (size_t)SSQ_NUM_SMS * SSQ_CTAS_PER_SM * 2 * sizeof(float) + (size_t)TENSOR_GRID * NC * sizeof(double);
}

static int minmax_launch(const float* x, int64_t rows, int64_t k, float* row_min, float* row_max,
                         unsigned int* tickets, float* partial, cudaStream_t st) {
    int64_t cap = (int64_t)SSQ_NUM_SMS * SSQ_CTAS_PER_SM;
    int64_t want = (cap + rows - 1) / rows;
    int64_t by_work = (k + SSQ_THREADS * 8 - 1) / (SSQ_THREADS * 8);
    int64_t s = want < by_work ? want : by_work; if (s < 1) s = 1;
    int64_t chunk = (k + s - 1) / s;
    int nsplit = (int)((k + chunk - 1) / chunk);
    row_minmax_kernel<<<dim3(nsplit, (unsigned)rows), SSQ_THREADS, 0, st>>>(x, k, chunk, row_min, row_max, tickets, partial);
    return launch_status();
}

extern "C" int ssq_row_minmax(const float* x, int64_t rows, int64_t k, float* row_min, float* row_max,
                              void* ws, size_t ws_bytes, void* stream) {
    if (rows == 0) return SSQ_OK;
    if (!x || !row_min || !row_max) return SSQ_ERR_NULL;
    if (rows < 0 || k <= 0 || rows > 65535) return SSQ_ERR_SIZE;
    if (!ws || ws_bytes < ssq_ws_bytes(rows)) return SSQ_ERR_WORKSPACE;
    WsView v = ws_view(ws, rows);
    return minmax_launch(x, rows, k, row_min, row_max, v.tickets, reinterpret_cast<float*>(v.partials), (cudaStream_t)stream);
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
    unsigned int* tickets = reinterpret_cast<unsigned int*>(ws);          // [0] minmax, [1] scores (shared header)
    char* base = reinterpret_cast<char*>(ws) + ws_ticket_bytes(1);
    float* mm = reinterpret_cast<float*>(base);
    float* mm_partial = reinterpret_cast<float*>(base + 256);
    double* partial = reinterpret_cast<double*>(base + 256 + (size_t)SSQ_NUM_SMS * SSQ_CTAS_PER_SM * 2 * sizeof(float));
    for (int64_t r = 0; r < rows; ++r) {
        const float* xr = x + r * k;
        int e = minmax_launch(xr, 1, k, mm, mm + 1, tickets, mm_partial, st);
        if (e) return e;
        int64_t tiles = (k + (int64_t)SSQ_THREADS * TE - 1) / ((int64_t)SSQ_THREADS * TE);
        int grid = (int)(tiles < TENSOR_GRID ? tiles : TENSOR_GRID);
        mse_search_tensor_kernel<<<grid, SSQ_THREADS, 0, st>>>(xr, k, mm, lm1, symmetric, p_norm, delta, zero_point,
                                                               raw_zero_point, best_score, best_index, r, tickets + 1, partial);
        e = launch_status();
        if (e) return e;
    }
    return SSQ_OK;
}

extern "C" size_t ssq_inp_scale_search_ws_bytes(int64_t k) { return (size_t)(k > 0 ? k : 1) * sizeof(int); }

extern "C" int ssq_inp_scale_search(const float* w, const float* delta, const float* raw_zero_point,
                                    const float* cand, int level, float x_range, float lo, float hi,
                                    float* inp_scale, int64_t oc, int64_t k,
                                    void* ws, size_t ws_bytes, void* stream) {
    if (oc == 0 || k == 0 || level == 0) return SSQ_OK;
    if (!w || !delta || !raw_zero_point || !cand || !inp_scale) return SSQ_ERR_NULL;
    if (oc < 0 || k < 0 || level < 0 || level > 65535 * CH) return SSQ_ERR_SIZE;
    if (!ws || ws_bytes < ssq_inp_scale_search_ws_bytes(k)) return SSQ_ERR_WORKSPACE;
    cudaStream_t st = (cudaStream_t)stream;
    dim3 grid((unsigned)((k + SSQ_THREADS - 1) / SSQ_THREADS), (unsigned)((level + CH - 1) / CH));
    inp_scale_fit_kernel<<<grid, SSQ_THREADS, 0, st>>>(w, delta, raw_zero_point, cand, level, x_range, lo, hi, oc, k,
                                                       reinterpret_cast<int*>(ws));
    int e = launch_status();
    if (e) return e;
    inp_scale_pick_kernel<<<(unsigned)((k + SSQ_THREADS - 1) / SSQ_THREADS), SSQ_THREADS, 0, st>>>(cand, reinterpret_cast<int*>(ws), inp_scale, k);
    return launch_status();
}
