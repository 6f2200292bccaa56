// K3 — reconstruction loss and its gradient in one pass over (pred, tgt[, fisher]).
//   reference arithmetic: quant/quant_layer.py:25-32 (lp_loss), quant/block_recon.py:154-162
//   (fisher_diag / fisher_full), autograd of those expressions for dpred.
// The target rows can be gathered through an index (the calibration mini-batch,
// quant/block_recon.py:90-92) so the cached FP outputs are read once, in place.
#include "ssq_common.cuh"

namespace ssq {

__device__ __forceinline__ float sgnf(float d) { return d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f); }

struct LossArgs {
    const float* pred; const float* tgt; const float* fisher; const int64_t* tgt_index;
    float* loss; float* dpred; int64_t batch; int64_t per_sample; double denom;
    float p; const float* gscale;
};

// MODE 0: lp with p == 2 (d*d, exact as ATen's pow(2) shortcut); 1: fisher_diag; 2: lp with general p, where
// |d|^p and |d|^(p-1) share one log2 (2 MUFU.EX2). One persistent grid; fixed-order reduction.
template <int MODE, bool VEC, bool GRAD, bool FWD>
__global__ void __launch_bounds__(SSQ_THREADS)
recon_loss_kernel(LossArgs a, WsView ws, int tiles_per_cta) {
    __shared__ double smem[32];
    const float g0 = a.gscale ? __ldg(a.gscale) : 1.f;
    const float inv = __fdiv_rn(g0, (float)a.denom);   // mean backward: grad / (numel/C)
    const float p = a.p, pm1 = a.p - 1.0f;
    const int64_t total = a.batch * a.per_sample;
    double acc[1] = {0.0};
    auto one = [&](float pr, float tg, float fi, float& go) -> float {
        float d = pr - tg;
        float ad = fabsf(d);
        float term;
        if (MODE == 0) {
            term = d * d;
            if (GRAD) go = (inv * (2.0f * ad)) * sgnf(d);
        } else if (MODE == 2) {
            const float l2 = log2_for_pow(ad);
            term = ex2_approx(p * l2);
            if (GRAD) go = (inv * (p * ex2_approx(pm1 * l2))) * sgnf(d);
        } else {
            float f2 = fi * fi;
            term = (d * d) * f2;
            if (GRAD) go = (inv * f2) * (2.0f * d);
        }
        return term;
    };
    // address-ordered tiles of SSQ_THREADS*U vectors (ssq_common.cuh); TileWalk: c = sample, col = offset inside the
    // sample (only needed to follow tgt_index)
    constexpr int U = 4;
    TileWalk tw;
    if (VEC) {
        const int64_t ps4 = a.per_sample >> 2, total4 = total >> 2, tile_v = (int64_t)SSQ_THREADS * U;
        const int64_t ntiles = (total4 + tile_v - 1) / tile_v;
        const int64_t t0 = (int64_t)blockIdx.x * tiles_per_cta, t1 = t0 + tiles_per_cta < ntiles ? t0 + tiles_per_cta : ntiles;
        for (int64_t tile = t0; tile < t1; ++tile) {
            const int64_t i0 = tile * tile_v + threadIdx.x;
            if (a.tgt_index) tw.init((uint64_t)i0, (uint64_t)ps4, (uint64_t)a.batch);
            float4 pr[U], tg[U], fi[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int64_t i = i0 + (int64_t)u * SSQ_THREADS;
                if (i < total4) {
                    const int64_t toff = a.tgt_index ? __ldg(a.tgt_index + tw.c) * ps4 + tw.col : i;
                    pr[u] = ld_stream4(a.pred + i * 4); tg[u] = ld_stream4(a.tgt + toff * 4);
                    if (MODE == 1) fi[u] = ld_stream4(a.fisher + toff * 4);
                }
                if (a.tgt_index) tw.step(SSQ_THREADS);
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int64_t i = i0 + (int64_t)u * SSQ_THREADS;
                if (i < total4) {
                    float4 go;
                    const float4 f = (MODE == 1) ? fi[u] : make_float4(0.f, 0.f, 0.f, 0.f);
                    const float s = one(pr[u].x, tg[u].x, f.x, go.x) + one(pr[u].y, tg[u].y, f.y, go.y)
                                  + one(pr[u].z, tg[u].z, f.z, go.z) + one(pr[u].w, tg[u].w, f.w, go.w);
                    if (GRAD) st_stream4(a.dpred + i * 4, go);
                    acc[0] += (double)s;
                }
            }
        }
    } else {
        const int64_t tile_e = (int64_t)SSQ_THREADS * U, ntiles = (total + tile_e - 1) / tile_e;
        const int64_t t0 = (int64_t)blockIdx.x * tiles_per_cta, t1 = t0 + tiles_per_cta < ntiles ? t0 + tiles_per_cta : ntiles;
        for (int64_t tile = t0; tile < t1; ++tile) {
            const int64_t i0 = tile * tile_e + threadIdx.x;
            if (a.tgt_index) tw.init((uint64_t)i0, (uint64_t)a.per_sample, (uint64_t)a.batch);
            for (int u = 0; u < U; ++u) {
                const int64_t i = i0 + (int64_t)u * SSQ_THREADS;
                if (i < total) {
                    const int64_t toff = a.tgt_index ? __ldg(a.tgt_index + tw.c) * a.per_sample + tw.col : i;
                    float go = 0.f;
                    const float s = one(a.pred[i], a.tgt[toff], MODE == 1 ? a.fisher[toff] : 0.f, go);
                    if (GRAD) a.dpred[i] = go;
                    acc[0] += (double)s;
                }
                if (a.tgt_index) tw.step(SSQ_THREADS);
            }
        }
    }
    if (!FWD) return;
    block_sum<1>(acc, smem);
    if (grid_finish<1>(acc, ws, 0, blockIdx.x, gridDim.x, smem) && threadIdx.x == 0)
        a.loss[0] = (float)(acc[0] / a.denom);
}

// fisher_full pass 1: per-sample dot_n = sum |d||g|; grid (splits, batch)
__global__ void __launch_bounds__(SSQ_THREADS)
fisher_full_dot_kernel(LossArgs a, int64_t chunk, WsView ws, double* dots) {
    __shared__ double smem[32];
    const int64_t n = blockIdx.y;
    const int64_t trow = a.tgt_index ? __ldg(a.tgt_index + n) : n;
    const float* pr = a.pred + n * a.per_sample;
    const float* tg = a.tgt + trow * a.per_sample;
    const float* fi = a.fisher + trow * a.per_sample;
    int64_t j0 = (int64_t)blockIdx.x * chunk;
    int64_t j1 = j0 + chunk < a.per_sample ? j0 + chunk : a.per_sample;
    double acc[1] = {0.0};
    for (int64_t j = j0 + threadIdx.x; j < j1; j += blockDim.x)
        acc[0] += (double)(fabsf(pr[j] - tg[j]) * fabsf(fi[j]));
    block_sum<1>(acc, smem);
    if (grid_finish<1>(acc, ws, n, blockIdx.x, gridDim.x, smem) && threadIdx.x == 0) dots[n] = acc[0];
}

// fisher_full pass 2: loss = sum_n dot_n^2 / (numel*100); dpred = g * 2 dot_n |f| sgn(d) / (numel*100)
template <bool FWD>
__global__ void __launch_bounds__(SSQ_THREADS)
fisher_full_grad_kernel(LossArgs a, const double* dots) {
    const float g0 = a.gscale ? __ldg(a.gscale) : 1.f;
    const int64_t total = a.batch * a.per_sample;
    const double scale = 1.0 / ((double)total * 100.0);
    if (FWD && blockIdx.x == 0 && threadIdx.x == 0) {
        double s = 0.0;
        for (int64_t n = 0; n < a.batch; ++n) s += dots[n] * dots[n];
        a.loss[0] = (float)(s * scale);
    }
    if (!a.dpred) return;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        int64_t n = i / a.per_sample;
        int64_t trow = a.tgt_index ? __ldg(a.tgt_index + n) : n;
        int64_t toff = trow * a.per_sample + (i - n * a.per_sample);
        float d = a.pred[i] - a.tgt[toff];
        a.dpred[i] = g0 * (float)(2.0 * dots[n] * scale) * fabsf(a.fisher[toff]) * sgnf(d);
    }
}

__global__ void __launch_bounds__(SSQ_THREADS)
gather_rows_kernel(const float* __restrict__ src, const int64_t* __restrict__ index, float* __restrict__ dst,
                   int64_t batch, int64_t per_sample, bool vec) {
    constexpr int U = 4;
    TileWalk tw;                      // c = output row, col = offset inside the row
    if (vec) {
        const int64_t ps4 = per_sample >> 2, total4 = batch * ps4;
        const int64_t i0 = (int64_t)blockIdx.x * (SSQ_THREADS * U) + threadIdx.x;       // one address-ordered tile per CTA
        tw.init((uint64_t)i0, (uint64_t)ps4, (uint64_t)batch);
        float4 v[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int64_t i = i0 + (int64_t)u * SSQ_THREADS;
            if (i < total4) v[u] = ld_stream4(src + (__ldg(index + tw.c) * ps4 + tw.col) * 4);
            tw.step(SSQ_THREADS);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int64_t i = i0 + (int64_t)u * SSQ_THREADS;
            if (i < total4) st_stream4(dst + i * 4, v[u]);
        }
    } else {
        const int64_t total = batch * per_sample;
        const int64_t i0 = (int64_t)blockIdx.x * (SSQ_THREADS * U) + threadIdx.x;
        tw.init((uint64_t)i0, (uint64_t)per_sample, (uint64_t)batch);
        for (int u = 0; u < U; ++u) {
            const int64_t i = i0 + (int64_t)u * SSQ_THREADS;
            if (i < total) dst[i] = src[__ldg(index + tw.c) * per_sample + tw.col];
            tw.step(SSQ_THREADS);
        }
    }
}

static int launch_loss(LossArgs a, int mode, bool fwd, void* ws, size_t ws_bytes, cudaStream_t st) {
    if (a.batch == 0 || a.per_sample == 0) return SSQ_OK;
    if (!a.pred || !a.tgt) return SSQ_ERR_NULL;
    if (fwd && !a.loss) return SSQ_ERR_NULL;
    if (!fwd && !a.dpred) return SSQ_ERR_NULL;
    if (a.batch < 0 || a.per_sample < 0 || !(a.denom > 0.0)) return SSQ_ERR_SIZE;
    if (mode < 0 || mode > 2) return SSQ_ERR_MODE;
    if (mode != 0 && !a.fisher) return SSQ_ERR_NULL;
    const int64_t total = a.batch * a.per_sample;
    if (mode == 2) {
        if (a.batch > 65535) return SSQ_ERR_SIZE;
        if (!ws || ws_bytes < ssq_ws_bytes(a.batch)) return SSQ_ERR_WORKSPACE;
        int64_t cap = SSQ_MAX_SLOTS;
        int64_t want = (cap + a.batch - 1) / a.batch;
        int64_t by_work = (a.per_sample + SSQ_THREADS * 4 - 1) / (SSQ_THREADS * 4);
        int64_t s = want < by_work ? want : by_work; if (s < 1) s = 1;
        int64_t chunk = (a.per_sample + s - 1) / s;
        int nsplit = (int)((a.per_sample + chunk - 1) / chunk);
        WsView v = ws_view(ws, a.batch);
        double* dots = v.partials + (size_t)(a.batch + cap + 64) * 2;
        fisher_full_dot_kernel<<<dim3(nsplit, (unsigned)a.batch), SSQ_THREADS, 0, st>>>(a, chunk, v, dots);
        int grid = grid_for((total + SSQ_THREADS * 4 - 1) / (SSQ_THREADS * 4));
        if (fwd) fisher_full_grad_kernel<true><<<grid, SSQ_THREADS, 0, st>>>(a, dots);
        else fisher_full_grad_kernel<false><<<grid, SSQ_THREADS, 0, st>>>(a, dots);
        return launch_status();
    }
    if (fwd && (!ws || ws_bytes < ssq_ws_bytes(1))) return SSQ_ERR_WORKSPACE;
    bool vec = (a.per_sample % 4 == 0) && aligned16(a.pred) && aligned16(a.tgt) &&
               (!a.dpred || aligned16(a.dpred)) && (!a.fisher || aligned16(a.fisher));
    const int64_t per_tile = (int64_t)SSQ_THREADS * 4 * (vec ? 4 : 1);
    const int64_t ntiles = (total + per_tile - 1) / per_tile;
    WsView v = ws_view(ws, 1);
    const bool grad = a.dpred != nullptr;
    int per_cta = 1;
    const unsigned grid = fwd ? tile_grid_balanced(ntiles, per_cta) : tile_grid(ntiles, false);
#define L(M, V, G, F) recon_loss_kernel<M, V, G, F><<<grid, SSQ_THREADS, 0, st>>>(a, v, per_cta)
#define LM(M) do { if (vec) { if (fwd) { if (grad) L(M, true, true, true); else L(M, true, false, true); } \
                              else L(M, true, true, false); } \
                   else { if (fwd) { if (grad) L(M, false, true, true); else L(M, false, false, true); } \
                          else L(M, false, true, false); } } while (0)
    if (mode == 0) { if (a.p == 2.0f) LM(0); else LM(2); } else LM(1);
#undef LM
#undef L
    return launch_status();
}

}  // namespace ssq

using namespace ssq;

extern "C" int ssq_recon_loss(const float* pred, const float* tgt, const float* fisher, const int64_t* tgt_index,
                              float* loss, float* dpred, int64_t batch, int64_t per_sample, double denom,
                              int mode, float p_norm, const float* gscale_dev,
                              void* ws, size_t ws_bytes, void* stream) {
    LossArgs a{pred, tgt, fisher, tgt_index, loss, dpred, batch, per_sample, denom, p_norm, gscale_dev};
    return launch_loss(a, mode, true, ws, ws_bytes, (cudaStream_t)stream);
}

extern "C" int ssq_recon_loss_bwd(const float* pred, const float* tgt, const float* fisher, const int64_t* tgt_index,
                                  const float* gloss, float* dpred, int64_t batch, int64_t per_sample, double denom,
                                  int mode, float p_norm, void* ws, size_t ws_bytes, void* stream) {
    LossArgs a{pred, tgt, fisher, tgt_index, nullptr, dpred, batch, per_sample, denom, p_norm, gloss};
    return launch_loss(a, mode, false, ws, ws_bytes, (cudaStream_t)stream);
}

extern "C" int ssq_gather_rows(const float* src, const int64_t* index, float* dst,
                               int64_t batch, int64_t per_sample, void* stream) {
    if (batch == 0 || per_sample == 0) return SSQ_OK;
    if (!src || !index || !dst) return SSQ_ERR_NULL;
    if (batch < 0 || per_sample < 0) return SSQ_ERR_SIZE;
    bool vec = (per_sample % 4 == 0) && aligned16(src) && aligned16(dst);
    int64_t total = batch * per_sample;
    const int64_t per_tile = (int64_t)SSQ_THREADS * 4 * (vec ? 4 : 1);
    const unsigned grid = tile_grid((total + per_tile - 1) / per_tile, false);
    gather_rows_kernel<<<grid, SSQ_THREADS, 0, (cudaStream_t)stream>>>(src, index, dst, batch, per_sample, vec);
    return launch_status();
}
