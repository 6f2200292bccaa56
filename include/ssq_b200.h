/*
 * ssq_b200.h — C-ABI of libssq_b200.so: the sm_100a kernels behind the
 * ShiftedScaleQuantization PTQ-calibration hot path.
 *
 * Conventions (every entry point):
 *   - plain pointers and sizes only; all tensor pointers are DEVICE fp32,
 *     contiguous, torch default layout (OIHW weights, NCHW activations);
 *   - `stream` is a cudaStream_t passed as void*; work is enqueued on it and the
 *     call returns immediately: never synchronises, never allocates, never throws;
 *   - reductions are deterministic (fixed-order partials + a ticket, no float
 *     atomics), so two runs on the same inputs are bit-identical;
 *   - workspaces are caller-owned device buffers, zero-filled ONCE when allocated
 *     (tickets self-reset); ask ssq_ws_bytes() for the size;
 *   - return value: 0 = SSQ_OK, <0 = argument error (below), >0 = cudaError_t.
 *
 * "channel layout": a tensor of n elements is viewed as [outer, nchan, inner];
 * the quantisation parameters of element i are delta[c], zero_point[c] with
 * c = (i / inner) % nchan. Weights [OC,IC,kh,kw] use nchan=OC, inner=IC*kh*kw;
 * per-tensor activation quantisers use nchan=1, inner=n.
 *
 * Each entry point cites the reference code (path:line under the upstream repo)
 * whose arithmetic it reproduces. Integer codes (rint/floor/clamp results) are
 * bit-exact with the reference's fp32 CPU path; floats agree to <=1e-5 relative.
 */
#ifndef SSQ_B200_H
#define SSQ_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SSQ_OK 0
#define SSQ_ERR_NULL (-1)      /* a required pointer is NULL */
#define SSQ_ERR_SIZE (-2)      /* n/inner/nchan inconsistent or out of range */
#define SSQ_ERR_WORKSPACE (-3) /* workspace missing or too small */
#define SSQ_ERR_MODE (-4)      /* unknown mode / unsupported combination */
#define SSQ_ERR_ALIGN (-5)     /* pointer not 4-byte aligned */

#define SSQ_ABI_VERSION 1
#define SSQ_MAX_SHIFTS 4      /* max len(shiftTarget) handled by the shift kernels */
#define SSQ_N_CANDIDATES 80   /* clip-ratio grid of quant/quant_layer.py:151 */

int ssq_abi_version(void);
const char* ssq_status_string(int status);
/* Bytes of zero-initialised workspace sufficient for any reduction entry point
 * below on a tensor with `nchan` channels. */
size_t ssq_ws_bytes(int64_t nchan);

/* ---- K1a: uniform affine fake-quant ------------------------------------------------
 * forward: quant/quant_layer.py:92-97 (UniformAffineQuantizer.forward);
 *          with in_scale != NULL: quant/channelQuantMSE.py:134-143 (ChannelQuantMSE.forward,
 *          two successive divisions x/in_scale/delta, dequant (q-zp)*delta*in_scale);
 *          ChannelQuant 'none' mode (quant/channelQuant.py:79-94) passes delta*shiftedScale.
 *   q = clamp(rint(x/delta[c]) + zp[c], qmin, qmax);  y = (q - zp[c]) * delta[c]
 * in_scale (nullable) has `inner` elements and requires outer == 1.
 * codes (nullable) receives q (integer-valued fp32, as the reference holds them). */
int ssq_fq_affine_fwd(const float* x, const float* delta, const float* zero_point,
                      const float* in_scale, float* y, float* codes,
                      int64_t n, int64_t inner, int64_t nchan,
                      float qmin, float qmax, void* stream);

/* backward of the expression above as autograd derives it (STE on rint only,
 * clamp passes the gradient on [qmin,qmax] inclusive):
 *   gx = gy*inside;  gdelta[c] = sum gy*(inside ? rint(u)-u : q-zp);  gzp[c] = sum gy*(inside ? 0 : -delta)
 * gx, gdelta, gzp are each nullable (skipped). ws: ssq_ws_bytes(nchan). */
int ssq_fq_affine_bwd(const float* gy, const float* x, const float* delta, const float* zero_point,
                      float* gx, float* gdelta, float* gzp,
                      int64_t n, int64_t inner, int64_t nchan,
                      float qmin, float qmax, void* ws, size_t ws_bytes, void* stream);

/* ---- K1b: AdaRound fake-quant -------------------------------------------------------
 * forward: quant/adaptive_rounding.py:49-59 ('learned_hard_sigmoid'), h() from :63-64;
 *          ChannelQuant 'adaround' mode quant/channelQuant.py:66-78 (sym-aware bounds).
 *   q = clamp(floor(w/delta[c]) + r + zp[c], qmin, qmax), r = h(alpha) if soft else (alpha >= 0)
 *   wq = (q - zp[c]) * delta[c],  h(a) = clamp(sigmoid(a)*1.2 - 0.1, 0, 1)
 * Rounding regulariser folded into the same pass (quant/block_recon.py:173-174):
 *   reg_out[0] = lambda * sum(1 - |2h-1|^b), b = *b_dev; written only if reg_out != NULL
 *   (needs ws). b <= 0 means "regulariser off" (block_recon.py:167) and writes 0. */
int ssq_fq_adaround_fwd(const float* w, const float* alpha, const float* delta, const float* zero_point,
                        float* wq, float* codes, int64_t n, int64_t inner, int64_t nchan,
                        float qmin, float qmax, int soft,
                        const float* b_dev, float lambda, float* reg_out,
                        void* ws, size_t ws_bytes, void* stream);

/* backward to alpha (soft mode): galpha = gwq*delta*inside*h'(alpha) + greg[0]*lambda*dR/dh*h'(alpha)
 * gwq nullable (regulariser-only gradient); greg nullable (=> treated as 1 when b_dev given,
 * regulariser gradient skipped when b_dev == NULL). accumulate != 0 adds into galpha. */
int ssq_fq_adaround_bwd(const float* gwq, const float* w, const float* alpha, const float* delta,
                        const float* zero_point, float* galpha,
                        int64_t n, int64_t inner, int64_t nchan, float qmin, float qmax,
                        const float* b_dev, float lambda, const float* greg,
                        int accumulate, void* stream);

/* alpha init, quant/adaptive_rounding.py:66-72: alpha = -log(1.2/(frac(w/delta)+0.1) - 1) */
int ssq_adaround_init_alpha(const float* w, const float* delta, float* alpha,
                            int64_t n, int64_t inner, int64_t nchan, void* stream);

/* Stand-alone regulariser (when a caller evaluates the loss separately, as
 * LossFunction does at quant/block_recon.py:170-174): mode 0 = sigmoid soft targets
 * (AdaRound h(alpha)), reg = lambda*sum(1-|2h-1|^b). gv nullable. ws: ssq_ws_bytes(1). */
int ssq_round_reg_fwd(const float* v, int64_t n, const float* b_dev, float lambda,
                      float* reg_out, void* ws, size_t ws_bytes, void* stream);
int ssq_round_reg_bwd(const float* v, int64_t n, const float* b_dev, float lambda,
                      const float* greg, float* gv, int accumulate, void* stream);

/* Multi-tensor K1b: every QuantModule of a reconstruction unit in ONE launch.
 * `table` is a HOST array of `count` <= SSQ_MT_MAX descriptors (device pointers inside); it is
 * copied by value into the kernel's parameter space, so the call is CUDA-graph capturable and
 * the caller may reuse the array immediately. tile_begin is the exclusive prefix of
 * ceil(n/SSQ_MT_TILE); a CTA finds its tensor by scanning it. The regulariser partials of all
 * tensors are reduced into reg_out[0]. */
#define SSQ_MT_MAX 16
#define SSQ_MT_TILE 4096
typedef struct ssq_adaround_desc {
    const float* w;
    const float* alpha;
    const float* delta;
    const float* zero_point;
    float* wq;          /* fwd output */
    const float* gwq;   /* bwd input  */
    float* galpha;      /* bwd output */
    int64_t n;
    int64_t inner;
    int64_t nchan;
    int64_t tile_begin;
    float qmin, qmax;
} ssq_adaround_desc;
int ssq_fq_adaround_fwd_mt(const ssq_adaround_desc* table, int count, int64_t total_tiles,
                           int soft, const float* b_dev, float lambda, float* reg_out,
                           void* ws, size_t ws_bytes, void* stream);
int ssq_fq_adaround_bwd_mt(const ssq_adaround_desc* table, int count, int64_t total_tiles,
                           const float* b_dev, float lambda, void* stream);

/* ---- one reconstruction iteration in three launches ----------------------------------------
 * quant/block_recon.py:89-105 (randperm mini-batch :90-92, block forward :95, LossFunction :97,
 * backward :99, optimizer.step :103) and the schedules it reads (LinearTempDecay :185-202,
 * CosineAnnealingLR :72-73). Device-side iteration state: *step = iterations COMPLETED so far
 * (only read during an iteration; the launch that applies Adam increments it when its last CTA
 * retires), idx_table [n_steps, batch] = the reference's CPU torch.randperm(N)[:B] stream
 * consumed up front, b_table / lr_table [n_steps]. Rows past n_steps-1 repeat the last row.
 * ssq_iter_prologue  = bookkeeping (idx_live/b_live/lr_live <- row *step, for the kernels that
 *   follow) + gather of the mini-batch's cached input rows (cache == NULL: skipped) + soft
 *   forward of every quantised layer of the unit with the regulariser (count == 0: skipped).
 *   Same arithmetic and the same fixed-order regulariser sum as ssq_fq_adaround_fwd_mt.
 * ssq_fq_adaround_bwd_adam_mt = ssq_fq_adaround_bwd_mt + ssq_adam_step on the alphas it just
 *   differentiated (flat/exp_avg/exp_avg_sq are the unit's flat buffers; every table[i].alpha
 *   points into flat), t = *step + 1, then ends the iteration. store_grad != 0 also writes
 *   table[i].galpha. Bit-identical to the two separate launches.
 * ssq_adam_step_end_iteration = ssq_adam_step with t = *step + 1, then ends the iteration
 *   (multi-GPU: after the all-reduce; activation phase). */
typedef struct ssq_iter_state {
    const int64_t* step;
    const int64_t* idx_table;
    int64_t* idx_live;        /* nullable */
    const float* b_table;     /* nullable */
    float* b_live;            /* nullable */
    const float* lr_table;    /* nullable */
    float* lr_live;           /* nullable */
    int64_t n_steps;
    int batch;
} ssq_iter_state;
/* The layers' output affine gamma^z / varphi^z (alpha_out / beta_out, quant/quant_layer.py:231-238,258-259; README
 * --bias_cal) FOLDED into the launches above: `aff` (nullable; parallel to `table`, per-output-channel layers only) makes
 * the prologue emit W_eff = gamma_oc * W_q into table[i].wq and b_eff = gamma*bias + varphi into aff[i].beff, so that
 * conv(x, W_eff) + b_eff == (conv(x, W_q) + bias) * gamma + varphi up to fp32 rounding and no activation-sized pass is
 * needed; the backward scales g_wq by gamma_oc. ssq_affine_grad_mt gives the affine's own gradients from the weight and bias
 * gradients of that folded layer: ggamma[c] = sum_k gwq[c,k] * W_q[c,k] + gbeff[c] * bias[c], gphi[c] = gbeff[c]
 * (one warp per output channel, fixed summation order; W_q re-evaluated from (w, alpha) as the forward did). */
typedef struct ssq_affine_desc {
    const float* gamma;      /* [OC] */
    const float* phi;        /* [OC] */
    const float* bias;       /* [OC] or NULL */
    float* beff;             /* [OC] out (prologue) */
    const float* gbeff;      /* [OC] in  (ssq_affine_grad_mt): gradient of the folded bias */
    float* ggamma;           /* [OC] out */
    float* gphi;             /* [OC] out */
    int64_t row_begin;       /* sum of OC over the layers before this one */
} ssq_affine_desc;
int ssq_iter_prologue(const ssq_iter_state* st, const float* cache, float* cur_inp, int64_t per_sample,
                      const ssq_adaround_desc* table, const ssq_affine_desc* aff, int count, int64_t total_tiles,
                      float lambda, float* reg_out, void* ws, size_t ws_bytes, void* stream);
/* apply_adam == 0: gradients only (store_grad must be set), the iteration is ended by whoever applies Adam */
int ssq_fq_adaround_bwd_adam_mt(const ssq_adaround_desc* table, const ssq_affine_desc* aff, int count, int64_t total_tiles,
                                const float* b_dev, float lambda,
                                float* flat, float* exp_avg, float* exp_avg_sq, const float* lr_dev,
                                int64_t* step_dev, double beta1, double beta2, double eps, int store_grad, int apply_adam,
                                void* ws, size_t ws_bytes, void* stream);
int ssq_affine_grad_mt(const ssq_adaround_desc* table, const ssq_affine_desc* aff, int count, void* stream);

/* ---- K1c: shifted-scale ChannelQuant ------------------------------------------------
 * quant/channelQuant.py:49-127. Conv weights [OC,IC,kh,kw] (kk = kh*kw) carry one
 * group-probability row per INPUT channel: alpha [IC,S]; FC weights [OC,IC] carry one per
 * element: alpha [OC*IC,S] (pass ic_groups = OC*IC, kk = 1, per_element = 1).
 * probs: p = clamp(softmax(alpha,-1)*1.2-0.1,0,1) (channelQuant.py:120-121) -> p [G,S];
 * also entropy regulariser -sum p*log(p+1e-10) (layer_recon_shiftedScale.py:393, mode 0) or
 * sum(1-|2p-1|^b) (layer_recon_fused_shiftedScale.py:281-282, mode 1) into reg_out (nullable).
 * When b_dev is given, *b_dev <= 0 switches the regulariser off in either mode (the losses' warm-up gate). */
int ssq_shift_probs_fwd(const float* alpha, float* p, int64_t groups, int nshift,
                        int reg_mode, const float* b_dev, float lambda, float* reg_out,
                        void* ws, size_t ws_bytes, void* stream);
/* galpha from gp (gradient wrt the clamped probabilities) plus the regulariser's own gradient. */
int ssq_shift_probs_bwd(const float* alpha, const float* gp, float* galpha, int64_t groups, int nshift,
                        int reg_mode, const float* b_dev, float lambda, const float* greg,
                        void* stream);

/* mode SSQ_SHIFT_DEQUANT ('learned_hard_sigmoid', channelQuant.py:81-82,96-118 with the x_q of
 *   init_v :201-213): term_i = (clamp(rint(w/(delta*s_i))+zp,qmin,qmax)-zp)*(delta*s_i);
 *   soft: y = sum_i term_i*p[g,i]; hard: y = term_{argmax_i p[g,i]} (first max).
 * mode SSQ_SHIFT_ADASHIFT ('adaShift', channelQuant.py:50-64 with the integer floors of
 *   init_v_beta :279-294): f = mix_i floor(w/(delta*s_i)); q = clamp(f + r + zp, qmin, qmax),
 *   r = h(beta) (soft round) or (beta>=0); y = (q-zp)*delta.
 * shift_delta: [S, nchan] = delta*s_i precomputed by the caller in fp32 (tensor*python-float). */
#define SSQ_SHIFT_DEQUANT 0
#define SSQ_SHIFT_ADASHIFT 1
int ssq_fq_shift_fwd(const float* w, const float* shift_delta, const float* delta, const float* zero_point,
                     const float* p, const float* beta, float* y,
                     int64_t oc, int64_t ic, int64_t kk, int nshift, int per_element,
                     int mode, int hard_targets, int hard_round, float qmin, float qmax, void* stream);
/* gp [G,S] (gradient wrt p, reduced over OC*kk per input channel; ws-backed, deterministic)
 * and gbeta (nullable; adaShift soft-round only). */
int ssq_fq_shift_bwd(const float* gy, const float* w, const float* shift_delta, const float* delta,
                     const float* zero_point, const float* p, const float* beta,
                     float* gp, float* gbeta,
                     int64_t oc, int64_t ic, int64_t kk, int nshift, int per_element,
                     int mode, int hard_round, float qmin, float qmax,
                     void* ws, size_t ws_bytes, void* stream);
size_t ssq_shift_bwd_ws_bytes(int64_t oc, int64_t ic, int64_t kk, int nshift, int per_element);

/* ---- K2a: MSE clip-ratio scale search -------------------------------------------------
 * quant/quant_layer.py:145-162 + quantize() :168-175. For every row (channel) of x [rows, k]:
 * 80 candidates f_i = fp32(1-0.01*i); score_i = mean|x-x_q|^2.4; first strict minimum wins.
 * Outputs per row: delta, zero_point, raw_zero_point (= -new_min), best score and winning
 * index (-1 if no candidate scored below 1e10, i.e. the reference's delta-stays-None case).
 * rows == 1 with large k is the per-tensor activation case (grid-wide variant; needs ws).
 * p_norm is 2.4 in the reference; n_levels = 2^n_bits.
 * Two passes, identical argmin: all candidates are ranked with a MUFU-based |d|^p, the ones
 * within 1e-4 of the ranked minimum are settled with libm powf in a fixed summation order
 * (csrc/scale_search.cu header). Rows shorter than 128 elements get one warp each. */
int ssq_mse_scale_search(const float* x, int64_t rows, int64_t k, int n_levels, int symmetric,
                         float p_norm, float* delta, float* zero_point, float* raw_zero_point,
                         float* best_score, int32_t* best_index,
                         void* ws, size_t ws_bytes, void* stream);
size_t ssq_mse_scale_search_ws_bytes(int64_t rows, int64_t k);
/* row-wise min/max (feeds the 'max' scale method, quant/quant_layer.py:124-142, whose
 * arithmetic is Python double on the host) */
int ssq_row_minmax(const float* x, int64_t rows, int64_t k, float* row_min, float* row_max,
                   void* ws, size_t ws_bytes, void* stream);

/* ---- K2b: ChannelQuantMSE input-scale search -----------------------------------------
 * quant/channelQuantMSE.py:70-110 ('max' mode). w [oc, k] (k = IC*kh*kw); candidates
 * cand[j], j < level, in the reference's order (descending: level/level ... 1/level, fp32);
 * a column fits candidate c when every row has lo < (w/c/delta + zero)/(L-1) < hi;
 * inp_scale[col] = the LAST fitting candidate, else it keeps its incoming value.
 * zero = rint(raw_zp/delta) per row is computed by the kernel.
 * One pass over w (4 B/element, independent of `level`): the predicate is monotone in the
 * candidate and a column's answer is set by its tightest element, so the pass keeps
 * max_r |w| / |V_r| per column (one select, one multiply, one integer max per element) and a
 * finish kernel turns it into the fitting prefix, settling with the exact predicate the
 * columns whose estimate lies within the error margin of a candidate boundary; inputs outside
 * the proof's preconditions (a foreign candidate list, zero outside [0, L-1], lo >= 0,
 * hi <= 1) run the brute-force sweep of all `level` candidates instead (device-side switch, no
 * host round trip). Four launches: row intervals, sweep, finish, brute force (idle).
 * _ex(force_brute=1) selects the brute force explicitly. ws: zero-filled once by the caller. */
int ssq_inp_scale_search(const float* w, const float* delta, const float* raw_zero_point,
                         const float* cand, int level, float x_range, float lo, float hi,
                         float* inp_scale, int64_t oc, int64_t k,
                         void* ws, size_t ws_bytes, void* stream);
int ssq_inp_scale_search_ex(const float* w, const float* delta, const float* raw_zero_point,
                            const float* cand, int level, float x_range, float lo, float hi,
                            float* inp_scale, int64_t oc, int64_t k, int force_brute,
                            void* ws, size_t ws_bytes, void* stream);
size_t ssq_inp_scale_search_ws_bytes(int64_t k);              /* sized for oc <= 65536 */
size_t ssq_inp_scale_search_ws_bytes2(int64_t oc, int64_t k);

/* ---- K3: reconstruction loss -----------------------------------------------------------
 * quant/quant_layer.py:25-32 (lp_loss 'none': sum|d|^p / (numel/C)), quant/block_recon.py:154-162
 * (fisher_diag, fisher_full). mode: 0 = lp, 1 = fisher_diag, 2 = fisher_full.
 * One pass reads pred,tgt[,fisher] and writes loss[0] and, when dpred != NULL,
 * dpred = gscale * dloss/dpred (gscale = *gscale_dev if given else 1).
 * denom = numel / C (N*H*W, or N for 2-D). tgt_index (nullable): int64 [batch]; row n of tgt
 * (and fisher) is read from tgt + tgt_index[n]*per_sample — the loss gathers the calibration
 * batch itself instead of reading a copy (quant/block_recon.py:90-92).
 * fisher_full needs two passes internally (per-sample dot products first). */
int ssq_recon_loss(const float* pred, const float* tgt, const float* fisher, const int64_t* tgt_index,
                   float* loss, float* dpred, int64_t batch, int64_t per_sample, double denom,
                   int mode, float p_norm, const float* gscale_dev,
                   void* ws, size_t ws_bytes, void* stream);
/* backward only (when the forward ran without dpred) */
int ssq_recon_loss_bwd(const float* pred, const float* tgt, const float* fisher, const int64_t* tgt_index,
                       const float* gloss, float* dpred, int64_t batch, int64_t per_sample, double denom,
                       int mode, float p_norm, void* ws, size_t ws_bytes, void* stream);

/* ---- true-integer export --------------------------------------------------
 * The reference never materialises integer weights; these entry points store the integer intermediate `x_quant` of
 * its hard forwards bit-packed and read it back into the dequantised weight those forwards return:
 *   alpha == NULL : q = clamp(rint(x/delta[c]) + zp[c], qmin, qmax)                 quant/quant_layer.py:92-96
 *   alpha != NULL : q = clamp(floor(x/delta[c]) + (alpha >= 0) + zp[c], qmin, qmax)  quant/adaptive_rounding.py:50-58
 *   in_scale != NULL (k elements): x is first divided by in_scale[j] and the dequantised value multiplied by it
 *                   (quant/channelQuantMSE.py:134-143; pass zp = rint(raw_zp/delta))
 *   import: w_q = ((u + qmin) - zp[c]) * delta[c] [* in_scale[j]]
 * The tensor is [rows, k] with the usual channel layout (inner, nchan) over its rows*k elements. Each row occupies
 * ssq_packed_row_bytes(k, n_bits) bytes; u = q - qmin is stored in sbits = smallest of {1,2,4,8} >= n_bits bits,
 * element j at bit (j % (8/sbits))*sbits of byte j/(8/sbits); padding bits are zero. `packed` is DEVICE memory. */
int64_t ssq_packed_row_bytes(int64_t k, int n_bits);
int ssq_export_codes(const float* w, const float* alpha, const float* in_scale, const float* delta,
                     const float* zero_point, uint8_t* packed, int64_t rows, int64_t k, int64_t inner,
                     int64_t nchan, float qmin, float qmax, int n_bits, void* stream);
int ssq_import_codes(const uint8_t* packed, const float* in_scale, const float* delta, const float* zero_point,
                     float* w_q, int64_t rows, int64_t k, int64_t inner, int64_t nchan, float qmin,
                     int n_bits, void* stream);

/* ---- per-output-channel affine on activations, quant/quant_layer.py:258-259 -----------
 * y = x*a[c] + b[c] with two roundings (mul then add, as the reference) */
int ssq_chan_affine_fwd(const float* x, const float* a, const float* b, float* y,
                        int64_t n, int64_t inner, int64_t nchan, void* stream);
int ssq_chan_affine_bwd(const float* gy, const float* x, const float* a, float* gx, float* ga, float* gb,
                        int64_t n, int64_t inner, int64_t nchan, void* ws, size_t ws_bytes, void* stream);

/* ---- fused multi-tensor Adam (torch.optim.Adam semantics, quant/block_recon.py:60,103) ---
 * one flat parameter/grad/state buffer per reconstruction unit; step count and lr are read
 * from device memory so the launch can live in a CUDA graph. */
int ssq_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n,
                  const float* lr_dev, double beta1, double beta2, double eps,
                  const int64_t* step_dev, void* stream);
/* t = *step_dev + 1 without ending the iteration (a second parameter group stepped before the launch that ends it) */
int ssq_adam_step_pending(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n,
                          const float* lr_dev, double beta1, double beta2, double eps,
                          const int64_t* step_dev, void* stream);
/* the same with t = *step_dev + 1 (step_dev = iterations completed), incrementing *step_dev when the last CTA retires */
int ssq_adam_step_end_iteration(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n,
                                const float* lr_dev, double beta1, double beta2, double eps,
                                int64_t* step_dev, void* ws, size_t ws_bytes, void* stream);

/* ---- the exchange step: gradient SUM over the ranks + Adam, one kernel over peer memory ----
 * reference intent: link.allreduce(p.grad) for every optimised parameter, then optimizer.step
 * (quant/block_recon.py:100-103). flat_ptrs[r] / grad_ptrs[r] / pad_ptrs[r] (HOST arrays of `world`
 * DEVICE pointers) are rank r's parameter buffer, gradient buffer and flag pad in symmetric
 * memory (same layout on every rank; pad = ssq_exchange_pad_bytes() zeroed bytes). Rank `rank`
 * sums elements [rank*S, (rank+1)*S), S = ssq_exchange_shard_elems(n, world), of all gradient
 * buffers in rank order, applies Adam (t = *step_dev + 1) to that shard — exp_avg_shard /
 * exp_avg_sq_shard hold S floats — and stores the new parameters into EVERY rank's parameter
 * buffer; then it ends the iteration (*step_dev += 1). n must be a multiple of 4. Every rank must
 * launch it with the same n and world. reduced_shard_out (nullable): the summed gradient shard.
 * *timeouts is incremented if a peer does not arrive within ~2^25 polls (results are then
 * undefined; the caller raises instead of hanging the device). epochs: ssq_exchange_pad_bytes()/32
 * zeroed uint32 in LOCAL device memory (per-CTA launch counters), owned by the caller for the life
 * of the symmetric buffers. */
size_t ssq_exchange_pad_bytes(void);
int64_t ssq_exchange_shard_elems(int64_t n, int world);
int ssq_grad_exchange_adam(float* const* flat_ptrs, const float* const* grad_ptrs, uint32_t* const* pad_ptrs,
                           int rank, int world, int64_t n,
                           float* exp_avg_shard, float* exp_avg_sq_shard,
                           const float* lr_dev, int64_t* step_dev, double beta1, double beta2, double eps,
                           float* reduced_shard_out, unsigned int* timeouts, uint32_t* epochs,
                           void* ws, size_t ws_bytes, void* stream);

/* ---- calibration loop plumbing -----------------------------------------------------------
 * gather rows of a cached feature tensor: dst[n] = src[index[n]] (quant/block_recon.py:91) */
int ssq_gather_rows(const float* src, const int64_t* index, float* dst,
                    int64_t batch, int64_t per_sample, void* stream);
/* host-resident feature cache (the reference's keep_gpu=False mode, quant/data_utils.py:34-36 +
 * `cached_inps[idx].to(device)` at quant/block_recon.py:91-92): dst[n] (device) = src[rows[n]] (PINNED host memory),
 * one cudaMemcpyAsync per row on `stream`; rows is a HOST array. No host-side gather, no staging copy. */
int ssq_stage_rows_h2d(const float* host_src, const int64_t* rows, float* dev_dst,
                       int64_t batch, int64_t per_sample, void* stream);
/* same mode, pull variant: the SMs read the rows out of MAPPED pinned host memory (cudaHostAlloc/UVA: the host
 * pointer is valid on the device) and write them to dev_dst. The rows are row `min(*step_dev + lookahead,
 * n_steps-1)` of the DEVICE index table (the reference's torch.randperm(N)[:B] stream, quant/block_recon.py:90), so
 * the transfer needs no host work per iteration and can be a node of a captured graph. max_ctas caps the grid
 * (<= 0: 32) so it can run beside the iteration it prefetches for. */
int ssq_pull_rows_host(const float* host_src_mapped, const int64_t* idx_table, const int64_t* step_dev,
                       int64_t lookahead, int64_t n_steps, float* dev_dst, int64_t batch, int64_t per_sample,
                       int max_ctas, void* stream);
/* same transfer from a ZERO-PACKED host cache (cached features are post-ReLU: a third to a half of the elements are +0.0f, and
 * this mode is PCIe-bound). The host keeps only the non-zero values of every row, packed in element order (host_vals: mapped
 * pinned memory, 16-byte aligned, >= 4 floats of slack after the last value); the index stays on the DEVICE: mask[r][per_sample/32]
 * (bit e%32 of word e/32 set iff the 32 bits of element e are not all zero) and chunk_off[N*per_sample/1024 + 1], the exclusive
 * prefix sum of the per-1024-element chunk counts, row-major. per_sample must be a multiple of 1024. The expansion is lossless
 * (-0.0f, denormals, NaNs are values). The reference keeps the cache as dense CPU tensors (quant/data_utils.py:29-36); the packed
 * form is this library's own. */
int ssq_pull_rows_host_packed(const uint32_t* mask, const float* host_vals_mapped, const int64_t* chunk_off,
                              const int64_t* idx_table, const int64_t* step_dev, int64_t lookahead, int64_t n_steps,
                              float* dev_dst, int64_t batch, int64_t per_sample, int max_ctas, void* stream);
/* advance the device-side iteration state used by a graph-captured loop: step += 1, and
 * copy row `step` of idx_table/b_table/lr_table into the live slots. */
int ssq_loop_advance(int64_t* step_dev, const int64_t* idx_table, int64_t* idx_live, int batch,
                     const float* b_table, float* b_live, const float* lr_table, float* lr_live,
                     int64_t n_steps, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SSQ_B200_H */
