/*
 * siss_b200.h — C ABI of libsiss_b200.so: the B200 (sm_100a) kernels behind the SISS
 * data-unlearning hot path.
 *
 * The reference (claserken/SISS) is pure Python/PyTorch: it has no FFI of its own, so each entry
 * point below cites the reference *call site* whose sequence of ATen launches it replaces
 * (paths relative to the reference checkout). The Python host layer (siss_b200/) binds these
 * with ctypes; INTEGRATION.md shows the stub a reference maintainer would add.
 *
 * Conventions
 *   - plain pointers + sizes, no torch types; every pointer is a DEVICE pointer unless the
 *     parameter name starts with `h_`;
 *   - tensors are contiguous, row-major `[B, D]` with D = C*H*W (NCHW flattened per sample);
 *   - `dtype` / `pred_dtype` are SISS_F32 / SISS_BF16 / SISS_F16;
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream);
 *   - all entry points are asynchronous w.r.t. the host and never synchronise the device;
 *   - return value: 0 on success, a positive cudaError_t, or a negative SISS_E* code.
 *     `siss_error_string` renders either.
 *   - reductions are fixed-order: results are bitwise reproducible run to run for a given
 *     shape on a given GPU.
 */
#ifndef SISS_B200_H
#define SISS_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SISS_B200_ABI_VERSION 3

enum siss_dtype { SISS_F32 = 0, SISS_BF16 = 1, SISS_F16 = 2 };

enum siss_status {
    SISS_OK = 0,
    SISS_EINVAL = -1,       /* null pointer / non-positive size / bad enum            */
    SISS_EUNSUPPORTED = -2, /* dtype combination not compiled in                       */
    SISS_EARCH = -3         /* device is not sm_100 (this library has no other target) */
};

enum siss_combine_mode {
    SISS_COMBINE_SCALING_NORM = 0, /* s = scaling_norm / ||g_a||        (delete_celeb.py:746)     */
    SISS_COMBINE_ERASEDIFF = 1,    /* s = -max(eta - <g_x,g_a>/||g_a||^2, 0)  (delete_celeb.py:741-742) */
    SISS_COMBINE_NONE = 2          /* s = 0: single-loss methods, only the clip applies (delete_celeb.py:682-684,767) */
};

typedef void* siss_stream_t;

int siss_abi_version(void);
const char* siss_error_string(int code);
/* Fails with SISS_EARCH unless the current device is compute capability 10.x. */
int siss_check_device(int* sm_count, int* cc_major, int* cc_minor);

/* ------------------------------------------------------------------------------------------
 * K1 — DDPM forward noising  x_t = sqrt(abar_t) * x0 + sqrt(1 - abar_t) * eps
 * Replaces diffusers==0.27.2 `DDPMScheduler.add_noise` as called at delete_celeb.py:602-603,
 * delete_tshirt.py:544-545, delete_sd.py:922-934. `alphas_cumprod` is the fp32 table [T]; it is
 * cast to `dtype` BEFORE the square roots and every product / the sum is rounded to `dtype`,
 * which is the eager rounding sequence. The pair form shares eps and t between the keep batch
 * (x0) and the forget batch (a0) as the reference does (delete_celeb.py:579-603).
 * Algorithmic bytes/element: single 3*s, pair 5*s  (s = sizeof dtype).
 * ---------------------------------------------------------------------------------------- */
int siss_add_noise(const void* x0, const void* noise, const int64_t* timesteps,
                   const float* alphas_cumprod, int T, void* xt,
                   int64_t B, int64_t D, int dtype, siss_stream_t stream);

int siss_add_noise_pair(const void* x0, const void* a0, const void* noise, const int64_t* timesteps,
                        const float* alphas_cumprod, int T, void* xt_x, void* xt_a,
                        int64_t B, int64_t D, int dtype, siss_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Workspace for the per-sample (row) reductions and the dynamic span queue of K2 / K3. Allocate
 * once, ZERO it once (cudaMemset) and reuse on one stream: the kernels leave every counter reset,
 * and the layout does not depend on the per-call B, so calls with different batch sizes can share
 * it. Size depends only on the largest B it will be used with (~150 KB + 4 B per row).
 * ---------------------------------------------------------------------------------------- */
int64_t siss_row_workspace_bytes(int64_t B);

/* ------------------------------------------------------------------------------------------
 * K2 — defensive-mixture sampling + importance weights.
 * Replaces losses/ddpm_deletion_loss.py:12-23 (gamma/sigma gather, Bernoulli row select) and
 * :32-45 (Gaussian exponents d_x, d_a and the two importance weights).
 *   keep_mask[b] != 0  -> row b of x_mix is taken from the keep batch, else from the forget batch
 *                        (the host draws it exactly as the reference: CPU `torch.rand(B) > lambd`).
 *   dist_x[b] = sum_D (x_mix - gamma_t x0)^2 / (2 sigma_t^2),  dist_a likewise with a0
 *   w_x[b] = 1 / ((1-lambd) + lambd * exp(dist_x - dist_a))
 *   w_a[b] = 1 / ((1-lambd) * exp(dist_a - dist_x) + lambd)
 * The exponent difference is accumulated directly (sum of (r_x - r_a)(r_x + r_a)) instead of
 * subtracting two ~D/2-sized fp32 sums, then fed through the reference's literal formula, so the
 * inf/0 saturation of the weights is reproduced ("importance weights").
 * Algorithmic bytes/element: 4*s (selected x_t row, x0, a0 read; x_mix written).
 * ---------------------------------------------------------------------------------------- */
int siss_mixture_weights(const void* xt_x, const void* xt_a, const void* x0, const void* a0,
                         const uint8_t* keep_mask, const int64_t* timesteps,
                         const float* gamma, const float* sigma, int T, double lambd,
                         void* x_mix, float* dist_x, float* dist_a, float* w_x, float* w_a,
                         void* workspace, int64_t B, int64_t D, int dtype, siss_stream_t stream);

/* K1 o K2 fused: x_t is formed only for the selected source and never written for the other.
 * Same outputs as siss_mixture_weights. Algorithmic bytes/element: 4*s (x0, a0, eps read; x_mix
 * written). */
int siss_add_noise_mixture(const void* x0, const void* a0, const void* noise,
                           const uint8_t* keep_mask, const int64_t* timesteps,
                           const float* alphas_cumprod, const float* gamma, const float* sigma, int T,
                           double lambd,
                           void* x_mix, float* dist_x, float* dist_a, float* w_x, float* w_a,
                           void* workspace, int64_t B, int64_t D, int dtype, siss_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * K3 — weighted epsilon-MSE.
 * Replaces losses/ddpm_deletion_loss.py:26-30 (targets eps_x, eps_a and squared errors), :51-53
 * (importance weighting), the scalarisation at delete_celeb.py:686-687 and autograd's backward
 * of all of that into the UNet output (delete_celeb.py:691,702).
 *   eps_x = (x_mix - gamma_t x0) / sigma_t     eps_a = (x_mix - gamma_t a0) / sigma_t   (fp32)
 *   loss_x = (pred - eps_x)^2                  loss_a = (pred - eps_a)^2
 *
 * siss_wmse_fwd_bwd (fast path): one pass that emits BOTH upstream gradients
 *   grad_x = (go_x * w_x[b]) * (2 (pred - eps_x))   grad_a = (go_a * w_a[b]) * (2 (pred - eps_a))
 * in pred's dtype, plus per-sample sums row_loss_{x,a}[b] = sum_D loss_{x,a} (for the stats block
 * delete_celeb.py:626-663). go_x / go_a are d(total)/d(weighted_loss element), i.e.
 * 1/(train_batch_size * grad_accum) in the reference. Nothing [B,D]-sized is materialised except
 * the two gradients. Algorithmic bytes/element: sizeof(pred) + 3*s + 2*sizeof(pred).
 * ---------------------------------------------------------------------------------------- */
int siss_wmse_fwd_bwd(const void* pred, int pred_dtype,
                      const void* x_mix, const void* x0, const void* a0, int dtype,
                      const int64_t* timesteps, const float* gamma, const float* sigma, int T,
                      const float* w_x, const float* w_a, float go_x, float go_a,
                      void* grad_x, void* grad_a, float* row_loss_x, float* row_loss_a,
                      void* workspace, int64_t B, int64_t D, siss_stream_t stream);

/* API-compatible forward: materialises the four fp32 [B,D] tensors of the reference 7-tuple
 * (losses/ddpm_deletion_loss.py:56). Any of the four outputs may be NULL (skipped). */
int siss_wmse_fwd(const void* pred, int pred_dtype,
                  const void* x_mix, const void* x0, const void* a0, int dtype,
                  const int64_t* timesteps, const float* gamma, const float* sigma, int T,
                  const float* w_x, const float* w_a,
                  float* loss_x, float* loss_a, float* wloss_x, float* wloss_a,
                  int64_t B, int64_t D, siss_stream_t stream);

/* API-compatible backward: grad_pred = sum over the present upstream gradients of
 *   go_loss_x * 2u_x + (go_wloss_x * w_x) * 2u_x + go_loss_a * 2u_a + (go_wloss_a * w_a) * 2u_a.
 * Each go_* is NULL (absent), a [B,D] fp32 tensor (stride 1) or one broadcast fp32 scalar in
 * device memory (stride 0: what autograd hands back for `.sum()`), selected by *_stride. */
int siss_wmse_bwd(const void* pred, int pred_dtype,
                  const void* x_mix, const void* x0, const void* a0, int dtype,
                  const int64_t* timesteps, const float* gamma, const float* sigma, int T,
                  const float* w_x, const float* w_a,
                  const float* go_loss_x, int go_loss_x_stride,
                  const float* go_loss_a, int go_loss_a_stride,
                  const float* go_wloss_x, int go_wloss_x_stride,
                  const float* go_wloss_a, int go_wloss_a_stride,
                  void* grad_pred, int64_t B, int64_t D, siss_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Plain squared error against a given target — the No-IS / EraseDiff / NegGrad / naive losses,
 * losses/ddpm_deletion_loss.py:60-96:  loss = (pred - target)^2 ; scaled = alpha * loss.
 * Output dtype is the torch promotion of (pred, target): fp32 unless both are the same 16-bit
 * type, in which case every op rounds to that type like eager does.
 * ---------------------------------------------------------------------------------------- */
int siss_sqerr_fwd(const void* pred, int pred_dtype, const void* target, int target_dtype,
                   void* loss, void* scaled /* nullable */, float alpha,
                   int64_t n, siss_stream_t stream);

/* grad_pred = go_loss * 2u + (go_scaled * alpha) * 2u ; go_* as in siss_wmse_bwd (fp32 or the
 * promoted 16-bit type, given by go_dtype). */
int siss_sqerr_bwd(const void* pred, int pred_dtype, const void* target, int target_dtype,
                   const void* go_loss, int go_loss_stride,
                   const void* go_scaled, int go_scaled_stride, float alpha, int go_dtype,
                   void* grad_pred, int64_t n, siss_stream_t stream);

/* Fast path for double_forward_with_neg_del / erasediff: both squared errors, both gradients
 * (grad = go * 2 (pred - target), in pred's dtype) and per-sample sums in one pass.
 * target_a may equal target_x (No-IS shares eps); it is then read once.
 * Algorithmic bytes/element (fp32, shared target): 3 reads + 2 writes = 20. */
int siss_dual_mse_fwd_bwd(const void* pred_x, const void* pred_a, int pred_dtype,
                          const void* target_x, const void* target_a, int target_dtype,
                          float go_x, float go_a, void* grad_x, void* grad_a,
                          float* row_loss_x, float* row_loss_a,
                          void* workspace, int64_t B, int64_t D, siss_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * K4 — two-term gradient combine on flat fp32 buffers.
 * Replaces the per-parameter Python loops at delete_celeb.py:714-753 (= delete_tshirt.py:656-697,
 * delete_sd.py:1071-1109) and the clip at delete_celeb.py:767.
 *
 * K4a siss_norm3: sums3 = { sum g_x^2, sum g_a^2, sum g_x*g_a } as three doubles (exact fp64
 * products, fp64 accumulation, fixed order). 8 bytes/parameter. `workspace` from siss_norm3_workspace_bytes,
 * zeroed once.
 *
 * K4b siss_combine: reads sums3 FROM DEVICE MEMORY (so an all-reduce of the three scalars can sit
 * between K4a and K4b with no host sync) and writes
 *     out = clip * (g_x - s * g_a)
 *   s    per `mode` (see siss_combine_mode); if inf_guard and s is inf, s = 0 (delete_tshirt.py:688-690)
 *   clip = min(1, max_norm / (||g_x - s g_a|| + 1e-6))   (torch.nn.utils.clip_grad_norm_);
 *          max_norm <= 0 disables it. The norm of the combination is obtained algebraically from
 *          sums3: ||g_x||^2 - 2 s <g_x,g_a> + s^2 ||g_a||^2, evaluated in fp64.
 *   stats5 (nullable) = { ||g_x||, ||g_a||, s, ||g_x - s g_a||, clip } — the three wandb scalars
 *          of delete_celeb.py:748 plus what clip_grad_norm_ returns.
 *   `out` may alias g_x or g_a. 12 bytes/parameter.
 * ---------------------------------------------------------------------------------------- */
int64_t siss_norm3_workspace_bytes(void);

int siss_norm3(const float* g_x, const float* g_a, int64_t n, double* sums3,
               void* workspace, siss_stream_t stream);

int siss_combine(const float* g_x, const float* g_a, float* out, int64_t n,
                 const double* sums3, int mode, float value, float max_norm, int inf_guard,
                 float* stats5, siss_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Multi-tensor K4 — K4a / K4b over a LIST of separately allocated gradient tensors (one per UNet
 * parameter, the layout of the reference loop, delete_celeb.py:717-750), for callers that keep
 * per-parameter tensors instead of GradCombiner's flat buffers. One launch for the whole list.
 * All `d_*` arrays live in DEVICE memory and are built once by the caller:
 *   d_gx[i], d_ga[i], d_out[i]  pointers of tensor i (d_out[i] may equal d_gx[i]);
 *   d_sizes[i]                  elements of tensor i;
 *   d_chunk_prefix[i]           number of siss_mt_chunk_elems()-element chunks in tensors 0..i-1,
 *                               d_chunk_prefix[n_tensors] == total_chunks.
 * Same arithmetic, scalars and stats as siss_norm3 / siss_combine; workspace as siss_norm3.
 * ---------------------------------------------------------------------------------------- */
int siss_mt_chunk_elems(void);

int siss_mt_norm3(const float* const* d_gx, const float* const* d_ga, const int64_t* d_sizes,
                  const int64_t* d_chunk_prefix, int n_tensors, int64_t total_chunks, double* sums3,
                  void* workspace, siss_stream_t stream);

int siss_mt_combine(const float* const* d_gx, const float* const* d_ga, float* const* d_out, const int64_t* d_sizes,
                    const int64_t* d_chunk_prefix, int n_tensors, int64_t total_chunks, const double* sums3,
                    int mode, float value, float max_norm, int inf_guard, float* stats5, siss_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * K4b fused with the optimiser step: g = clip * (g_x - s * g_a) is consumed in registers by a
 * torch.optim.AdamW update (decoupled weight decay; config/delete_celeb.yaml:127-134, stepped at
 * delete_celeb.py:769) of the flat fp32 parameter buffer `param` and its moments; with zero_grads the
 * two gradient buffers are cleared in the same pass (replaces optimizer.zero_grad(), :773).
 *   sums3 == NULL        : no combine scalars — g = g_x (already combined / exchanged), no clip;
 *   mode == SISS_COMBINE_NONE : single-term methods, g = clip * g_x (g_a may be NULL);
 *   grad_out (nullable)  : also materialise g (may alias g_x), e.g. for logging or a later hook;
 *   step                 : 1-based optimiser step count (bias corrections 1 - beta^step);
 *   d_step (nullable)    : if given, the step count is read from this DEVICE int64 instead (and `step` is
 *                          ignored), so a captured CUDA graph stays valid across optimiser steps; advance it
 *                          on the stream with siss_counter_add before the call.
 *   d_sched (nullable)   : DEVICE double[2] = {lr, ema_decay}; if given they replace the scalar arguments, so a
 *                          learning-rate schedule (lr_scheduler.step(), delete_celeb.py:770) and EMA warm-up
 *                          survive CUDA-graph replay;
 *   ema_param (nullable) : flat fp32 shadow parameters, updated in the same pass as diffusers' EMAModel.step does
 *                          right after the optimiser step (delete_celeb.py:776-777; cfg.ema.use_ema):
 *                          shadow -= (1 - ema_decay) * (shadow - param_new); +8 bytes/parameter.
 * 40 bytes/parameter (5 reads + 5 writes) instead of 48 in 4 launches for K4b + AdamW + 2 memsets.
 * ---------------------------------------------------------------------------------------- */
int siss_combine_adamw(float* g_x, float* g_a, int64_t n, const double* sums3, int mode, float value,
                       float max_norm, int inf_guard, float* param, float* exp_avg, float* exp_avg_sq,
                       double lr, double beta1, double beta2, double eps, double weight_decay, int64_t step,
                       const int64_t* d_step, const double* d_sched, float* ema_param, double ema_decay,
                       int zero_grads, float* grad_out, float* stats5, siss_stream_t stream);

/* *d_counter += value, stream-ordered (one thread). For the device-side optimiser step count. */
int siss_counter_add(int64_t* d_counter, int64_t value, siss_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Opt-in device-side draws from a counter-based stream (Philox4x32-10 + Box-Muller; definition in
 * csrc/philox.cuh, CPU restatement oracle/philox.py). They replace torch.randn (delete_celeb.py:581),
 * torch.randint (:593) and the CPU `torch.rand(B) > lambd` with its H2D copy
 * (losses/ddpm_deletion_loss.py:18). This is a NEW seed semantic (hence opt-in): a value depends only on
 * (seed, draw, global element / row index), so data-parallel ranks passing their global offsets draw exactly
 * the slices of the 1-rank tensors. `draw` (< 2^62) names one independent draw, e.g. the micro-step index;
 * d_draw (nullable): read the draw index from this DEVICE uint64 instead (advance it on the stream with
 * siss_counter_add), so a captured CUDA graph draws fresh values on every replay.
 *   siss_randn      : out[i] = N(0,1) sample of global element elem_offset + i, rounded to `dtype`;
 *   siss_draw_rows  : timesteps[r] uniform in [t_lo, t_hi) and keep_mask[r] = (uniform > lambd) for global row
 *                     row_offset + r; either output may be NULL.
 *   siss_add_noise_mixture_rng : siss_add_noise_mixture (K1oK2) with eps generated in registers from the same
 *                     stream instead of being read: bit-identical to siss_randn(noise, ...) followed by
 *                     siss_add_noise_mixture(..., noise, ...), at 3*s instead of 4*s bytes per element and without
 *                     the randn launch. noise_out (nullable) also materialises eps, for the methods whose
 *                     target is eps itself. elem_offset = global index of this rank's first element.
 * ---------------------------------------------------------------------------------------- */
int siss_randn(void* out, int64_t n, int dtype, uint64_t seed, uint64_t draw, const uint64_t* d_draw, uint64_t elem_offset,
               siss_stream_t stream);
int siss_draw_rows(int64_t* timesteps, uint8_t* keep_mask, int64_t B, uint64_t seed, uint64_t draw, const uint64_t* d_draw,
                   uint64_t row_offset, int64_t t_lo, int64_t t_hi, double lambd, siss_stream_t stream);
int siss_add_noise_mixture_rng(const void* x0, const void* a0, const uint8_t* keep_mask, const int64_t* timesteps,
                               const float* alphas_cumprod, const float* gamma, const float* sigma, int T,
                               double lambd, uint64_t seed, uint64_t draw, const uint64_t* d_draw, uint64_t elem_offset,
                               void* x_mix, void* noise_out, float* dist_x, float* dist_a, float* w_x, float* w_a,
                               void* workspace, int64_t B, int64_t D, int dtype, siss_stream_t stream);
/* EraseDiff with the forget target drawn in-kernel (opt-in device RNG): siss_dual_mse_fwd_bwd where target_a is not
 * read but generated — uniform [0, 1) from the aux domain of the same stream, rounded to the prediction dtype, like
 * torch.rand_like(eps_hat_a) (losses/ddpm_deletion_loss.py:75). 16 + s bytes/element instead of 20 + the rand_like
 * launch's 4. target_a_out (nullable, prediction dtype) also materialises the target. */
int siss_dual_mse_rng_fwd_bwd(const void* pred_x, const void* pred_a, int pred_dtype, const void* target_x, int target_dtype,
                              uint64_t seed, uint64_t draw, const uint64_t* d_draw, uint64_t elem_offset,
                              float go_x, float go_a, void* grad_x, void* grad_a, void* target_a_out,
                              float* row_loss_x, float* row_loss_a, void* workspace, int64_t B, int64_t D,
                              siss_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Membership-loss metric (metrics/class_membership.py:66-116): I sampled images x n_noise shared noise
 * draws at one timestep. Expanded row r = i * n_noise + j pairs image i with noise j (:76-86); the
 * expansion is never materialised. Both calls work on the expanded-row slice [row0, row0 + rows) — one
 * eval batch (:101-105).
 *   siss_membership_add_noise : xt_x[q], xt_a[q] = add_noise(x0[i] | a0[i], noise[j], timestep) (:92-93),
 *                               x0, a0 [I, D], noise [n_noise, D], outputs [rows, D], all of `dtype`;
 *   siss_membership_sqerr     : sum_x[q] = sum_d (pred_x[q, d] - noise[j, d])^2, sum_a likewise (:108-109);
 *                               pred_* fp32 [rows, D] (UNet outputs), noise of `dtype`; fp32 accumulation.
 *                               workspace: siss_row_workspace_bytes(rows) zeroed bytes.
 * ---------------------------------------------------------------------------------------- */
int siss_membership_add_noise(const void* x0, const void* a0, const void* noise, const float* alphas_cumprod,
                              int T, int64_t timestep, void* xt_x, void* xt_a, int64_t row0, int64_t rows,
                              int64_t n_noise, int64_t D, int dtype, siss_stream_t stream);
int siss_membership_sqerr(const float* pred_x, const float* pred_a, const void* noise, int dtype, float* sum_x,
                          float* sum_a, void* workspace, int64_t row0, int64_t rows, int64_t n_noise, int64_t D,
                          siss_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Fused statistics epilogue — the per-batch logging scalars of delete_celeb.py:626-656 (mean over
 * all elements; max / min / unbiased std of the per-sample means; mean / max / min / std of the
 * importance weights) in one launch from the O(B) per-sample sums of K2/K3.
 * out16 = [loss_x: mean,max,min,std | loss_a: ... | importance_weight_x: ... | importance_weight_a: ...].
 * Any input may be NULL (its four outputs are NaN). row_loss_* are per-sample SUMS over D elements.
 * ---------------------------------------------------------------------------------------- */
int siss_batch_stats(const float* row_loss_x, const float* row_loss_a, const float* w_x, const float* w_a,
                     int64_t B, int64_t D, float* out16, siss_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Data-parallel exchange fused with K4 over NVLink peer memory (one process per GPU, 2/4/8 GPUs of
 * one NVSwitch box). Replaces what DDP's bucketed all-reduce does implicitly inside
 * `accelerator.backward` on the sync step (delete_celeb.py:691) — and also reduces the NegGrad term,
 * which the reference never does.
 * `h_*` parameters are HOST arrays of `world` DEVICE pointers, one per rank, all mapped into this
 * process (CUDA IPC / symmetric memory). The caller must barrier all ranks on the stream before the
 * first kernel (every rank's G_x, G_a complete), between the two (every rank's scalar slot written)
 * and after the second (every rank's output complete).
 *
 * siss_p2p_reduce_norm3: reduce-scatter(G_x) + reduce-scatter(G_a) + K4a in one kernel. Rank `rank`
 * reads elements [rank*shard_len, (rank+1)*shard_len) of every peer's buffers, sums them in rank
 * order, writes the reduced shard to shard_x / shard_a (local), writes this rank's three partial sums
 * to sums3_local and to doubles [4*rank .. 4*rank+2] of every peer's scalar buffer h_peer_scalars[r]
 * (each at least 4*world doubles). Inbound NVLink bytes per rank: (world-1)/world * 8 per parameter.
 * x_prereduced == 1: shard_x already holds the reduced G_x shard (its reduce-scatter was overlapped with
 * the second backward pass, G_x being final after the first); only G_a crosses NVLink (4 B/parameter)
 * and h_peers_x is ignored. x_prereduced == 2: G_a ONLY (first phase of the pipelined exchange below):
 * h_peers_x / shard_x are ignored, the published sums are {0, sum a^2, 0}.
 *
 * siss_p2p_combine_allgather: K4b + all-gather in one kernel. Sums the `world` scalar slots (local
 * copy, rank order, so every rank derives identical s and clip), computes
 * clip * (shard_x - s * shard_a) and stores it to elements [rank*shard_len, ...) of EVERY peer's
 * h_peers_out[r]. Outbound NVLink bytes per rank: (world-1)/world * 4 per parameter.
 * ---------------------------------------------------------------------------------------- */
int64_t siss_p2p_workspace_bytes(void);

int siss_p2p_reduce_norm3(const float* const* h_peers_x, const float* const* h_peers_a,
                          double* const* h_peer_scalars, int world, int rank, int64_t shard_len,
                          float* shard_x, float* shard_a, double* sums3_local, int x_prereduced,
                          void* workspace, siss_stream_t stream);

/* Region-wise use of the kernels of this section (GradCombiner issues the reduce of each REGION of the flat buffer as
 * soon as autograd has finalised it, under the second backward pass): call any reduce / gather entry point once per
 * region with the peer tables pre-offset to the region start and shard_len = the rank's slice of the region — the
 * kernels address `peer + rank * shard_len + i`, which is then exactly that slice — and the shard pointers offset to the
 * slice's place. Each such reduce call writes its sums to its own `sums3_local`; siss_publish_sums adds the `regions`
 * triples in region order and stores the total to doubles [4*rank ..] of every peer's scalar buffer. */
int siss_publish_sums(const double* region_sums, int regions, double* const* h_peer_scalars, int world, int rank,
                      siss_stream_t stream);

int siss_p2p_combine_allgather(const float* shard_x, const float* shard_a, const double* scalar_slots,
                               float* const* h_peers_out, int world, int rank, int64_t shard_len,
                               int mode, float value, float max_norm, int inf_guard, float* stats5,
                               siss_stream_t stream);

/* Sharded (ZeRO-1) optimiser step fused with the PARAMETER all-gather — the data-parallel form of
 * siss_combine_adamw (delete_celeb.py:714-773): K4b on this rank's gradient shard, the AdamW (+EMA) update of this
 * rank's shard of parameters / moments in registers, new parameters stored to elements [rank*shard_len, ...) of
 * every peer's flat parameter buffer h_peers_param[r] (own included). exp_avg / exp_avg_sq / ema_shard are LOCAL
 * shard-sized buffers. Follows siss_p2p_reduce_norm3 + barrier exactly like siss_p2p_combine_allgather, and needs a
 * barrier after it before any rank reads its parameters. step / d_step / d_sched / ema_* as in siss_combine_adamw. */
int siss_p2p_adamw_allgather(const float* shard_x, const float* shard_a, const double* scalar_slots,
                             float* const* h_peers_param, int world, int rank, int64_t shard_len,
                             int mode, float value, float max_norm, int inf_guard,
                             float* exp_avg, float* exp_avg_sq, double lr, double beta1, double beta2, double eps,
                             double weight_decay, int64_t step, const int64_t* d_step, const double* d_sched,
                             float* ema_shard, double ema_decay, float* stats5, siss_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * The same exchange through the NVSwitch's in-fabric reduction / replication (NVLS). `mc_*` are MULTICAST device
 * addresses bound to the same symmetric allocation on every rank (cuMulticast*, e.g. torch symmetric memory's
 * `multicast_ptr`): `multimem.ld_reduce` returns the sum over all ranks of the addressed 16 bytes (inbound bytes /
 * world), `multimem.st` replicates a store into every rank's buffer (outbound bytes / world). The fp32 additions
 * happen in the switch in a fabric-defined order: every rank still holds bit-identical results (each element is
 * reduced once, by its owner), but they are not bit-equal to the rank-ordered sums of the siss_p2p_* kernels.
 * Barriers exactly as for the siss_p2p_* pair. Replaces the same reference sites (delete_celeb.py:691, 714-767).
 *
 * siss_nvls_reduce_norm3      reduce-scatter + K4a. x_mode 0: G_x and G_a through the switch; 1: shard_x already
 *                             reduced, G_a through the switch; 2: G_a only (published sums {0, sum a^2, 0}).
 * siss_nvls_combine_allgather K4b on the shard + one multimem.st per vector (all-gather in the switch) into mc_out.
 * siss_nvls_adamw_allgather   ZeRO-1 step as siss_p2p_adamw_allgather, new parameters replicated through mc_param;
 *                             param_local = THIS rank's (unicast) flat parameter buffer.
 *
 * PIPELINED exchange for the scaling-norm modes (SISS / SISS No-IS, s = scaling_norm / ||G_a||, delete_celeb.py:746):
 * the reduce of G_x (outbound-heavy) and the gather of the result (inbound-heavy) run concurrently in one kernel.
 *   1. siss_p2p_reduce_norm3(x_prereduced = 2) or siss_nvls_reduce_norm3(x_mode = 2)      | barrier
 *   2. siss_nvls_xcombine_bcast: for the own shard x = multimem.ld_reduce(G_x), y = x - s * shard_a (s from the
 *      phase-1 slots `scalar_slots1`), multimem.st(y) IN PLACE into every rank's G_x; publishes
 *      {sum x^2, sum x a, sum y^2} to doubles [4*rank ..] of every peer's second slot array h_peer_scalars2[r] | barrier
 *   3. siss_scale_finalize (local): n_x, n_a, s, total_norm = sqrt(sum y^2), clip (clip_grad_norm_, :767) -> stats5,
 *      and g *= clip over the whole local buffer (skipped when clip == 1).
 * ---------------------------------------------------------------------------------------- */
int siss_nvls_reduce_norm3(const float* mc_x, const float* mc_a, double* const* h_peer_scalars,
                           int world, int rank, int64_t shard_len, float* shard_x, float* shard_a,
                           double* sums3_local, int x_mode, void* workspace, siss_stream_t stream);

int siss_nvls_combine_allgather(const float* shard_x, const float* shard_a, const double* scalar_slots,
                                float* mc_out, int world, int rank, int64_t shard_len,
                                int mode, float value, float max_norm, int inf_guard, float* stats5,
                                siss_stream_t stream);

int siss_nvls_adamw_allgather(const float* shard_x, const float* shard_a, const double* scalar_slots,
                              float* mc_param, const float* param_local, int world, int rank, int64_t shard_len,
                              int mode, float value, float max_norm, int inf_guard,
                              float* exp_avg, float* exp_avg_sq, double lr, double beta1, double beta2, double eps,
                              double weight_decay, int64_t step, const int64_t* d_step, const double* d_sched,
                              float* ema_shard, double ema_decay, float* stats5, siss_stream_t stream);

int siss_nvls_xcombine_bcast(float* mc_x, const float* shard_a, const double* scalar_slots1,
                             double* const* h_peer_scalars2, int world, int rank, int64_t shard_len,
                             float scaling_norm, int inf_guard, void* workspace, siss_stream_t stream);

int siss_scale_finalize(float* g, int64_t n, const double* scalar_slots1, const double* scalar_slots2, int world,
                        float scaling_norm, float max_norm, int inf_guard, float* stats5, siss_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * The same exchange with the bytes moved by the COPY ENGINES (DMA over NVLink, ~700-730 GB/s per direction measured
 * against 530-620 GB/s for SM-issued peer / multicast traffic) and the SMs working on local memory only. `staging` is a
 * local buffer of 2 * (world-1) * shard_len floats. The shard is
 * processed in `chunks` (1..16) pieces so that the kernels overlap the copies. Results are bit-identical to the
 * siss_p2p_* pair (same rank-ordered sums). Side streams / events are owned by the library; all work is ordered against
 * `stream`. Barriers between ranks exactly as for siss_p2p_*; h_* arrays as there (entry [rank] = the local buffer).
 *
 * siss_ce_reduce_norm3      DMA-pull this rank's shard of every peer's G_a (and G_x, x_mode 0) + rank-ordered sum + K4a.
 * siss_ce_combine_allgather K4b on the shard into the own buffer + DMA-push to every peer.
 * (x_mode 2 of the reduce serves as phase 1 of the pipelined schedule above, followed by siss_nvls_xcombine_bcast.)
 * ---------------------------------------------------------------------------------------- */
int siss_ce_reduce_norm3(const float* const* h_peers_x, const float* const* h_peers_a, double* const* h_peer_scalars,
                         int world, int rank, int64_t shard_len, float* staging, float* shard_x, float* shard_a,
                         double* sums3_local, int x_mode, int chunks, void* workspace, siss_stream_t stream);

int siss_ce_combine_allgather(const float* shard_x, const float* shard_a, const double* scalar_slots,
                              float* const* h_peers_out, int world, int rank, int64_t shard_len, int chunks,
                              int mode, float value, float max_norm, int inf_guard, float* stats5, siss_stream_t stream);

#ifdef __cplusplus
}
#endif

#endif /* SISS_B200_H */
