/*
 * sdt_b200.h -- C ABI of libsdt_b200.so: the B200 (sm_100a) implementation of
 * SCAL-SDT's LoRA training-step hot path.
 *
 * The reference (MooerFoes/scal-sdt) is pure Python and has no FFI of its own;
 * each entry point below replaces the arithmetic that sits under one reference
 * call site (citations relative to /root/reference).  The Python host layer in
 * scal_sdt_b200/ binds these symbols with ctypes and keeps the reference's
 * module-injection / loss / EMA interfaces unchanged.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer owned by the caller (PyTorch), 16-byte
 *     aligned and contiguous, unless the comment says "host";
 *   - every call is asynchronous on `stream` (pass torch's current stream) and
 *     is CUDA-graph capturable; the library never allocates tensor memory and
 *     never synchronises;
 *   - return value 0 = success, negative = error; sdt_last_error() returns the
 *     thread-local message.  Unsupported dtype/shape is an ERROR, never a
 *     fallback (there is no CPU path in this library);
 *   - `void* stream` is a cudaStream_t.
 */
#ifndef SDT_B200_H_
#define SDT_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* element types */
enum { SDT_F32 = 0, SDT_BF16 = 1, SDT_F16 = 2 };   /* SDT_F16 (IEEE half): the LoRA projection entry points and sdt_lora_pack */
/* prediction target, modules/model.py:306-314 */
enum { SDT_TARGET_EPSILON = 0, SDT_TARGET_SAMPLE = 1, SDT_TARGET_V = 2 };
/* error codes */
enum { SDT_OK = 0, SDT_ERR_ARG = -1, SDT_ERR_UNSUPPORTED = -2, SDT_ERR_CUDA = -3, SDT_ERR_NCCL = -4 };

int         sdt_version(void);
const char* sdt_last_error(void);
/* 0 when the current device is compute capability 10.x (B200), else SDT_ERR_UNSUPPORTED */
int         sdt_device_check(void);
/* number of CUDA kernels this library has launched in the calling process (bench.py's gpu_launches) */
long long   sdt_launch_count(void);

/* ---- K1: fused LoRA projection, forward ---------------------------------------------------
 * Replaces loralib.Linear.forward reached through get_lora (modules/lora.py:12-14):
 *     Ts = scaling * (X A^T)            [M,r]   (kept on-chip; also written to t_save)
 *     Y  = X W^T + bias + Ts B^T        [M,N]
 * SDT_BF16: x[M,K], w[N,K], A[r,K], B[N,r], y[M,N], t_save[M,r] bf16, bias[N] f32 or NULL;
 *           r in {16,32,64} (host layer zero-pads smaller ranks); K%8==0, N%8==0.
 *           One tcgen05/TMEM/TMA kernel launch.
 * SDT_F16 : as SDT_BF16 with every 16-bit tensor in IEEE half (the same kernels; the format is an instruction-descriptor
 *           field plus the conversions in the epilogues).
 * SDT_F32 : all f32, any r >= 1; FFMA kernels (parity path for the reference's fp32 CPU config).
 * A == NULL (r == 0) gives the plain frozen projection Y = X W^T + bias.
 */
int sdt_lora_linear_fwd(const void* x, const void* w, const float* bias, const void* A, const void* B,
                        float scaling, void* y, void* t_save,
                        int64_t M, int64_t K, int64_t N, int r, int dtype, void* stream);

/* ---- K1 + residual: y = round(round(x W^T + bias + scaling (x A^T) B^T) + residual) -----------------------------------
 * The output of ff.net.2 / proj_out is added to the residual stream right away (diffusers BasicTransformerBlock:
 * hidden_states = ff(norm3(hidden_states)) + hidden_states; Transformer2DModel: output = proj_out(...) + residual).  The
 * epilogue reads the residual tile and adds it to the rounded output -- the values torch's separate add produces, without
 * that add's pass over both tensors.  residual [M,N], same dtype as y; SDT_BF16 / SDT_F16.
 */
int sdt_lora_linear_fwd_res(const void* x, const void* w, const float* bias, const void* A, const void* B, float scaling,
                            const void* residual, void* y, void* t_save, int64_t M, int64_t K, int64_t N, int r, int dtype,
                            void* stream);

/* ---- K1 (grouped): several projections of ONE shape in one launch ---------------------------
 * The reference calls to_q / to_k / to_v of a self-attention on the same normalised hidden states, and to_k / to_v
 * of a cross-attention on the same text context (diffusers CrossAttention.forward under modules/lora.py:12-14): three
 * (two) launches of 6.7 GFLOP each.  This entry runs up to SDT_MAX_GROUP such projections -- identical (M, K, N, r,
 * scaling), own W / A / B / bias / y / t_save, x may be shared -- as the work items of ONE persistent kernel, so launch,
 * prologue and pipeline-drain costs are paid once and the wave quantisation is that of the combined tile count.
 * `problems` is a HOST array; SDT_BF16 or SDT_F16; every problem has a bias or none has.
 */
#define SDT_MAX_GROUP 4
typedef struct {
  const void* x; const void* w; const float* bias; const void* A; const void* B; void* y; void* t_save;
  const void* residual;       /* NULL, or [M,N] added to y in the epilogue (as sdt_lora_linear_fwd_res) */
} sdt_lora_problem;
int sdt_lora_linear_fwd_group(const sdt_lora_problem* problems /* host */, int n_problems, float scaling,
                              int64_t M, int64_t K, int64_t N, int r, int dtype, void* stream);

/* ---- K1 (one input, many widths): up to SDT_MAX_MULTI projections of ONE (M, K, padded rank) with their own output width ----
 * to_k / to_v of EVERY cross-attention of the UNet read the same frozen text context (SD1.5: 32 projections of the 616 x 768
 * context to 320 / 640 / 1280 columns): one launch whose work items are (row tile, problem, column tile) instead of eight launches
 * of 9-15 us at 5-25 % of the tensor peak.  Ns[q] = output width of problem q (HOST array); x may be shared; every problem has a
 * bias or none has.  Needs M >= 256, K >= 256 (sdt_lora_linear_fwd_multi_supported); SDT_BF16 / SDT_F16.
 */
#define SDT_MAX_MULTI 32
int sdt_lora_linear_fwd_multi_supported(int n_problems, int64_t M, int64_t K, const int64_t* Ns /* host */, int r);
int sdt_lora_linear_fwd_multi(const sdt_lora_problem* problems /* host */, const int64_t* Ns /* host */, int n_problems,
                              float scaling, int64_t M, int64_t K, int r, int dtype, void* stream);

/* ---- K1 + GEGLU epilogue: ff.net.0.proj and the activation that follows it, one launch ------------------
 * diffusers GEGLU.forward: proj = self.proj(x); h, gate = proj.chunk(2, -1); return h * gelu(gate) -- with self.proj a LoRA site
 * (configs/optim_targets/lora.yaml:23-27).  w [2I,K], bias [2I], B [2I,r] in their reference row order ([h rows ; gate rows]).
 *     proj [M,2I] = x W^T + bias + scaling (x A^T) B^T          (written: the backward needs it; reference column order)
 *     act  [M,I]  = proj[:, :I] * gelu(proj[:, I:])             (erf GELU, evaluated on the rounded proj values)
 * A 256 x 128 tile holds 64 h columns and the 64 matching gate columns (the two CTAs of a pair bring the two row blocks), so the
 * activation is formed in the epilogue and the separate pass over proj disappears.  Needs M >= 256, I % 64 == 0, padded rank
 * 16/32/64 (sdt_lora_linear_geglu_supported); otherwise call sdt_lora_linear_fwd and sdt_geglu.  SDT_BF16 / SDT_F16.
 */
int sdt_lora_linear_geglu_supported(int64_t M, int64_t K, int64_t I, int r);
int sdt_lora_linear_geglu_fwd(const void* x, const void* w, const float* bias, const void* A, const void* B, float scaling,
                              void* proj, void* act, void* t_save, int64_t M, int64_t K, int64_t I, int r, int dtype, void* stream);

/* ---- K2: fused LoRA projection, backward ---------------------------------------------------
 * Autograd of the above with W, bias frozen (modules/model.py:137):
 *     G  = scaling * (dY B)             [M,r]   (on-chip; also written to g_ws)
 *     dX = dY W + G A                   [M,K]   (skipped when dx == NULL)
 *     dA (+)= G^T X                     [r_true,K] f32
 *     dB (+)= dY^T Ts                   [N,r_true] f32
 * SDT_BF16: dy[M,N], x[M,K], t_save[M,r], g_ws[M,r] bf16; wt = W^T [K,N] bf16 (cached copy of the
 *           frozen weight); At = A^T [K,r] bf16; Bt = B^T [r,N] bf16 (from sdt_lora_pack).
 *           dA/dB are f32 and are ACCUMULATED into: zero them first unless accumulating across
 *           micro-batches.  r_true <= r is the un-padded rank: only dA[:r_true,:] and
 *           dB[:, :r_true] (row stride r_true) are written.
 *           `ws`: device workspace of sdt_lora_wgrad_workspace_bytes() bytes, 16-byte aligned, ZEROED ONCE by the
 *           caller (the kernels leave it zeroed where it matters), used by one stream at a time.  With it the
 *           token-slice partial sums of dA / dB are combined in a fixed order (two stages inside the one launch):
 *           gradients are bit-identical from run to run, like the reference's torch path.  ws == NULL: the partials
 *           leave through red.global.add (same values up to f32 summation order, not reproducible).
 * SDT_F32 : wt = W [N,K] (no transposed copy needed), At = A [r,K], Bt = B [N,r], r_true == r; ws unused.
 */
size_t sdt_lora_wgrad_workspace_bytes(void);
int sdt_lora_linear_bwd(const void* dy, const void* x, const void* wt, const void* At, const void* Bt,
                        const void* t_save, float scaling, void* dx, void* g_ws, float* dA, float* dB,
                        int64_t M, int64_t K, int64_t N, int r, int r_true, int dtype, void* ws, void* stream);

/* ---- K2 (grouped): backward of projections that read ONE input (to_q / to_k / to_v) ----------------------
 * The reference's autograd runs three backward GEMMs and two adds for dX = sum_q (dY_q W_q + s (dY_q B_q) A_q).  Here the sum
 * is ONE K loop over the concatenated contraction (each projection is a "source" with its own rank-R accumulator in TMEM), so
 * dX is written once; G_q is written to g_ws of each problem and the dA / dB reductions follow, one launch per projection.
 * 2..3 problems of identical (M, K, N, r, scaling); padded rank 16 or 32; M >= 256 (sdt_lora_linear_bwd_group_supported says
 * whether a shape qualifies; otherwise call sdt_lora_linear_bwd per site and add).  `problems` is a HOST array; bf16 only.
 * dx == NULL (the input needs no gradient: to_k / to_v on the text context): up to SDT_MAX_GROUP problems, the rank
 * projections G_q = s dY_q B_q run as the work items of one launch (wt may be NULL).
 */
typedef struct {
  const void* dy; const void* x; const void* wt; const void* At; const void* Bt; const void* t_save;
  void* g_ws; float* dA; float* dB;
} sdt_lora_bwd_problem;
int sdt_lora_linear_bwd_group_supported(int n_problems, int need_dx, int64_t M, int64_t K, int64_t N, int r);
int sdt_lora_linear_bwd_group(const sdt_lora_bwd_problem* problems /* host */, int n_problems, float scaling, void* dx,
                              int64_t M, int64_t K, int64_t N, int r, int r_true, int dtype, void* ws /* as above */,
                              void* stream);

/* ---- K2 (batched reductions): dA / dB of up to sdt_lora_wgrad_max_sites() sites of ANY shapes in ONE launch -----------
 * sdt_lora_linear_bwd(_group) called with dA == NULL and dB == NULL computes only dX and G = s dY B (g_ws) and leaves the two
 * token reductions to the caller, who collects the sites of a transformer block (q / k / v / out of both attentions, the
 * feed-forward projections, proj_in / proj_out: different M, K, N, one padded rank r) and reduces them here:
 *     dA_i (+)= G_i^T X_i        dB_i (+)= dY_i^T Ts_i
 * One launch instead of one per site: prologue, ramp-up and the tail of the deterministic second stage are paid once.  The
 * operands must stay alive and unchanged until the launch has run.  `sites` is a HOST array.  ws as for sdt_lora_linear_bwd.
 */
typedef struct {
  const void* x; const void* g_ws; float* dA; const void* dy; const void* t_save; float* dB;
  int64_t M, K, N;
} sdt_wgrad_site;
int sdt_lora_wgrad_max_sites(void);
int sdt_lora_wgrad_batch(const sdt_wgrad_site* sites /* host */, int n_sites, int r, int r_true, int dtype, void* ws, void* stream);

/* ---- LoRA dropout on the rank path (modules/lora.py:12; loralib 0.1: (dropout(x) A^T) B^T * scaling) --------------------
 * Compatibility path (every shipped optim_target has dropout 0): the projection runs on the concatenated contraction
 * X' = [x | xd], W' = [W | 0], A' = [0 | A] with xd = x * keep / (1 - p), through the same fused kernels.
 *   backward == 0: in = x [M,K] -> out = X' [M,2K]
 *   backward == 1: in = dX' [M,2K] -> out = dx [M,K] = dX'[:, :K] + dX'[:, K:] * keep / (1 - p)
 * keep is regenerated from Philox4x32-10 keyed by *seed_dev (device int64: a captured graph draws fresh masks per replay
 * when the caller refreshes it) and `salt` (distinct per site and call).  p is realised on a 1/65536 grid.  bf16 / fp16.
 */
int sdt_lora_dropout(const void* in, void* out, int64_t M, int64_t K, float p, const int64_t* seed_dev, uint64_t salt,
                     int backward, int dtype, void* stream);

/* ---- LoRA operand packing (multi-tensor, one launch for all sites) ---------------------------
 * For every site i: from the f32 master lora_A[r_true,K], lora_B[N,r_true] write the four bf16
 * operand layouts the tensor-core kernels read, zero-padded to rank r:
 *     A_p[r,K]  At_p[K,r]  B_p[N,r]  Bt_p[r,N]
 * `sites` is a DEVICE array of n_sites sdt_pack_site records (built once by the host layer).  With dtype SDT_F16 the
 * outputs are IEEE half instead of bf16 (the reference's stock `trainer.precision: 16`, configs/lora.yaml:50).
 */
typedef struct {
  const float* A;  const float* B;          /* f32 masters */
  void* A_p; void* At_p; void* B_p; void* Bt_p;   /* bf16 outputs */
  int32_t K, N, r_true, r;
} sdt_pack_site;
int sdt_lora_pack(const sdt_pack_site* sites, int n_sites, int64_t max_site_elems, int dtype /* SDT_BF16 | SDT_F16 */,
                  void* stream);

/* ---- K3: DDPM noising + prediction target --------------------------------------------------
 * Replaces scheduler.add_noise / get_velocity (modules/model.py:302,312):
 *     a = sqrt(abar[t_b]), s = sqrt(1 - abar[t_b])     (abar first cast to the sample dtype)
 *     noisy  = a*x0 + s*eps
 *     target = a*eps - s*x0      (mode V only; for EPSILON / SAMPLE pass target == NULL: the
 *                                 reference aliases eps / x0, no copy is made)
 * x0, eps, noisy, target: [B, chw] of `dtype`; t: int64[B]; alphas_cumprod: f32[T].
 * Bit-exact with the torch op sequence (no FMA contraction).  t outside [0,T) is clamped on the
 * device and reported through *oob_flag (int32, device, may be NULL), the async analogue of the
 * reference's IndexError.
 */
int sdt_noise_target(const void* x0, const void* eps, const int64_t* t, const float* alphas_cumprod,
                     int num_train_timesteps, void* noisy, void* target, int mode,
                     int64_t B, int64_t chw, int dtype, int32_t* oob_flag, void* stream);

/* ---- K4: MSE loss, two-segment mean, fused dPred ---------------------------------------------
 * Replaces F.mse_loss(reduction="none") + chunk + mean (modules/model.py:316,338-342):
 *     L = (pred - target)^2 in f32
 *     split == B : loss = mean(L)
 *     split <  B : loss = mean(L[:split]) + w_prior * mean(L[split:])   (instance | class halves)
 * loss_out: f32[3] = {loss, mean(first segment), mean(second segment or 0)}.
 * dpred (NULL or [B,chw] of pred_dtype) = grad_scale * dloss/dpred.
 * loss_elem (NULL or f32[B,chw]) = L, for callers of the reference's per-element _denoise_loss.
 * nan_flag (NULL or int32 device) is set to 1 if any L is NaN (async analogue of raise_if_nan).
 * workspace: sdt_mse_loss_workspace_bytes() bytes, zero-initialised once by the caller; the kernel
 * leaves it zeroed.  Deterministic two-stage reduction (no float atomics).
 */
size_t sdt_mse_loss_workspace_bytes(void);
int sdt_mse_loss(const void* pred, int pred_dtype, const void* target, int target_dtype,
                 float* loss_out, void* dpred, float* loss_elem, int32_t* nan_flag,
                 int64_t B, int64_t chw, int64_t split, float w_prior, float grad_scale,
                 void* workspace, void* stream);

/* ---- K5: EMA update -----------------------------------------------------------------------------
 * Replaces ExponentialMovingAverage.update's per-tensor loop (modules/ema.py:56-61):
 *     tmp = s - p ; tmp *= (1-d) ; s -= tmp        (same three roundings, bit-exact in f32)
 * flat : one contiguous arena of n elements.
 * multi: n_tensors separate tensors; `chunks` is a device array of n_chunks {tensor, offset}
 *        records covering every tensor in pieces of <= chunk_elems elements.
 * one_minus_decay_dev: optional DEVICE f32 scalar that overrides the immediate (lets a captured
 * CUDA graph follow the warm-up schedule min(d,(1+n)/(10+n)), modules/ema.py:47-54).
 */
typedef struct { int32_t tensor; int32_t pad; int64_t offset; } sdt_chunk;
int sdt_ema_update_flat(void* shadow, const void* param, int64_t n, float one_minus_decay,
                        const float* one_minus_decay_dev, int dtype, void* stream);
int sdt_ema_update_multi(void* const* shadow_ptrs, const void* const* param_ptrs, const int64_t* numels,
                         const sdt_chunk* chunks, int n_chunks, int chunk_elems, float one_minus_decay,
                         const float* one_minus_decay_dev, int dtype, void* stream);

/* ---- f1: fused AdamW over the flat LoRA arena (torch.optim.AdamW semantics) -------------------
 * modules/model.py:33-64 builds one torch AdamW param group per site; over the flat arena this is
 * a single pass.  hyper (host values) = {lr, beta1, beta2, eps, weight_decay, bias_corr1, bias_corr2};
 * hyper_dev (optional device f32[7]) overrides it for graph replay.  grad_scale multiplies g first
 * (1/world after a sum-allreduce, or 1).  If ema_shadow != NULL the EMA lerp of K5 is applied to the
 * freshly updated parameter in the same pass.
 */
int sdt_adamw_flat(float* p, const float* g, float* m, float* v, int64_t n, const float* hyper,
                   const float* hyper_dev, float grad_scale, float* ema_shadow, float ema_one_minus_decay,
                   const float* ema_one_minus_decay_dev, void* stream);

/* ---- f2: GEGLU, the activation after ff.net.0.proj (diffusers GEGLU: proj = [h | gate], out = h * gelu_erf(gate)) ----
 * backward == 0: out_or_dproj [M,I]  = h * gelu(gate)                         (dout ignored)
 * backward == 1: out_or_dproj [M,2I] = [dout * gelu(gate) | dout * h * gelu'(gate)]
 * proj [M,2I]; dtype SDT_BF16 (128-bit vectorised, I % 8 == 0) or SDT_F32.  HBM-bound: 3 I (fwd) / 5 I (bwd) elements per row.
 */
int sdt_geglu(const void* proj, const void* dout, void* out_or_dproj, int64_t M, int64_t I, int backward, int dtype,
              void* stream);

/* ---- f2: GroupNorm (+ SiLU) on channels-last bf16 activations [B, HW, C], frozen affine (gamma, beta f32[C]) ----
 * chan_bias (NULL or bf16 [B, C]): the tensor that is normalised is x + chan_bias[b, c] -- the time-embedding add in front of
 * norm2 of a ResNet block (diffusers ResnetBlock2D: hidden = hidden + temb[:, :, None, None]; hidden = norm2(hidden)) folded
 * into the norm; x itself is not rewritten.
 * forward : stats f32[B,G,2] (written: per-group sum, sum of squares); y = act((x' - mean) * rstd * gamma + beta), x' = x + chan_bias
 * backward: dx (= dx') from (x, chan_bias, dout, stats).  d chan_bias = sum over HW of dx (caller).
 * ws: scratch of ws_floats >= sdt_group_norm_workspace_floats(B, G) f32 (no initialisation needed): the statistics pass leaves one
 *     partial per CTA there and the apply pass sums them in CTA order -- no float atomics, results are bit-reproducible like
 *     torch's native GroupNorm, and there is no zero-fill in front of the statistics launch.
 * Needs C % 8 == 0 and C / G >= 8.
 */
int64_t sdt_group_norm_workspace_floats(int64_t B, int G);
int sdt_group_norm_nhwc(const void* x, const void* chan_bias, const float* gamma, const float* beta, float* stats, void* y,
                        int64_t B, int64_t HW, int C, int G, float eps, int silu, float* ws, int64_t ws_floats, void* stream);
int sdt_group_norm_nhwc_bwd(const void* x, const void* chan_bias, const void* dout, const float* gamma, const float* beta,
                            const float* stats, float* ws, int64_t ws_floats, void* dx, int64_t B, int64_t HW, int C, int G, float eps,
                            int silu, void* stream);

/* ---- f2: residual add with folded per-channel bias on channels-last bf16 [rows, C]: out = a + b + bias[c] (bias f32[C]) ----
 * End of diffusers ResnetBlock2D: out = shortcut(x) + conv2(h); the convolutions run without bias and their frozen biases are
 * added in this pass (torch adds each conv bias as a separate non-vectorised broadcast kernel).  C % 8 == 0.
 */
int sdt_residual_bias_add(const void* a, const void* b, const float* bias, void* out, int64_t rows, int C, void* stream);

/* ---- f2: LayerNorm over C of token-major bf16 activations [M, C], frozen affine (gamma, beta f32[C]), fused residual add ----
 * diffusers BasicTransformerBlock (the UNet loaded at modules/model.py:82-91): x = x + attn(...); h = norm(x).
 * forward : res != NULL: xs_out = bf16(x + res) (the new residual stream) and y = LN(xs_out); res == NULL: y = LN(x).
 *           stats f32[M,2] = (mean, rstd) per row.
 * backward: dx = LN'(dy) (+ dres, the gradient arriving on the residual stream); xs = the tensor that was normalised.
 * One warp per row, row resident in registers, 128-bit accesses; C % 8 == 0, C <= 2048.
 */
int sdt_layer_norm_fwd(const void* x, const void* res, const float* gamma, const float* beta, void* xs_out, void* y,
                       float* stats, int64_t M, int C, float eps, void* stream);
int sdt_layer_norm_bwd(const void* xs, const void* dy, const void* dres, const float* gamma, const float* stats, void* dx,
                       int64_t M, int C, void* stream);

/* ---- K6: data-parallel LoRA-gradient exchange (replaces Lightning DDP, train.py:98-109) -------
 * One NCCL communicator per process (libnccl is resolved at run time with dlopen, so the library
 * loads on machines without NCCL).  sdt_allreduce averages `count` elements in place.
 */
int sdt_comm_unique_id(void* out_128_bytes /* host */);
int sdt_comm_init(const void* unique_id_128_bytes /* host */, int rank, int world);
int sdt_comm_world(void);
int sdt_allreduce(void* buf, int64_t count, int dtype, void* stream);
int sdt_comm_destroy(void);

/* ---- test support: slow SIMT GEMM used by the GPU tests as an on-device cross-check -----------
 * C[M,N] = alpha * sum_k A[m*lda_m + k*lda_k] * B[n*ldb_n + k*ldb_k] (+ beta * C), f32.
 */
int sdt_simt_gemm_f32(const float* A, int64_t lda_m, int64_t lda_k, const float* B, int64_t ldb_n, int64_t ldb_k,
                      float* C, int64_t ldc, const float* bias, float alpha, float beta,
                      int64_t M, int64_t N, int64_t K, void* stream);

/* test / measurement support (tools/): process-wide switches, 0 = default behaviour.
 *   0..9   weight-gradient kernel: override shared-memory / instruction descriptor constants (layout probing)
 *   10     device pointer to 128 int64 slots: clock64 timeline of CTA 0 of the next GEMM launches (tools/gemm_trace.py)
 *   11     1: force the single-CTA GEMM kernel          12  1: never use 224-wide tiles
 *   13     2: weight-stationary schedule of the single-CTA kernel (K <= 320; measured: no gain)
 *   14     K threshold above which the CTA-pair kernel is used (default 256)
 *   15     weight-gradient grid: min 64-token chunks per CTA | CTAs per SM << 8
 *   20     force the number of column tiles per work item in the CTA-pair kernel (tools/gemm_ab.py groups)
 *   22     minimum N for 224-wide (192-wide at rank 64) tiles            24  programmatic dependent launch: 1 off, 2 on
 *   30     1: A operand of the CTA-pair kernel through tensor memory (tcgen05.cp + TS-mode UMMAs; measured slower: tools/gemm_ab.py ts)
 *   31     double tiles (two column tiles of a work item in one joint K loop): minimum K (default 1280); 1 = never (tools/gemm_ab.py dt)
 */
int sdt_debug_set(int key, uint64_t value);

#ifdef __cplusplus
}
#endif
#endif /* SDT_B200_H_ */
