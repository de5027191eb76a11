/* vdt_b200.h — C ABI of the B200-native sampling hot path for tqch/v-diffusion-torch.
 *
 * The reference has no FFI: its boundary for this path is a Python object protocol
 * (SURVEY.md §8b).  Each entry point below names the reference interface it stands in for;
 * the Python shim in v-diffusion-torch_b200/ (UNet, GaussianDiffusion) binds them with ctypes
 * and INTEGRATION.md shows the binding a reference maintainer would add.
 *
 * Conventions: every function returns 0 on success, non-zero on error (vdt_last_error() gives the
 * message, thread-local).  Unless the name ends in _host, data pointers are DEVICE pointers on the
 * current CUDA device and the call is stream-ordered on `stream` (a cudaStream_t, may be NULL).
 * A plan owns its workspace; calls on one plan must not overlap in time.
 */
#ifndef VDT_B200_H
#define VDT_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VDT_MAX_LEVELS 8

/* UNet constructor integers — v_diffusion/models/unet.py:155-171 (UNet.__init__), as they arrive from
 * config["model"] after the defaults merge (generate.py:86-91).  head_dim / num_heads: 0 = None. */
typedef struct vdt_unet_config {
    int32_t in_channels;
    int32_t hid_channels;
    int32_t out_channels;
    int32_t num_levels;
    int32_t ch_multipliers[VDT_MAX_LEVELS];
    int32_t num_res_blocks;
    int32_t apply_attn[VDT_MAX_LEVELS];
    int32_t embedding_dim;        /* 0 = 4 * hid_channels */
    int32_t head_dim;
    int32_t num_heads;
    int32_t num_classes;
    int32_t multitags;            /* 1: labels are fp32 multi-hot [B, num_classes] (CelebA attributes, unet.py:209-210, 290-294) */
    int32_t resolution;           /* H = W of the images this plan serves (DATA_INFO[...]["resolution"]) */
    int32_t max_rows;             /* UNet batch rows processed per pass; larger batches are chunked */
    int32_t operand_dtype;        /* 16-bit tensor-core operand format: 0 = fp16 (default), 1 = bf16; accumulation,
                                     GroupNorm statistics, softmax and the residual stream are fp32 either way.
                                     2 = split-precision validation mode: every operand is an fp16 hi/lo pair and every
                                     product three K segments (hi*hi + lo*hi + hi*lo, ~22-bit mantissa) on the same
                                     tcgen05 kernels; attention runs in fp32 on CUDA cores. ~3x slower. */
} vdt_unet_config;

/* GaussianDiffusion constructor + get_logsnr_schedule arguments — diffusion.py:260-291, 42-112. */
enum { VDT_OUT_X0 = 0, VDT_OUT_EPS = 1, VDT_OUT_BOTH = 2, VDT_OUT_V = 3 };
enum { VDT_VAR_FIXED_SMALL = 0, VDT_VAR_FIXED_LARGE = 1, VDT_VAR_FIXED_MEDIUM = 2 };
enum { VDT_SCHED_COSINE = 0, VDT_SCHED_LINEAR = 1, VDT_SCHED_SIGMOID = 2, VDT_SCHED_LEGACY = 3 };
typedef struct vdt_sampler_config {
    int32_t sample_timesteps;     /* T */
    int32_t model_out_type;       /* VDT_OUT_* */
    int32_t model_var_type;       /* VDT_VAR_* */
    int32_t logsnr_schedule;      /* VDT_SCHED_* */
    int32_t use_ddim;             /* p_sample(..., use_ddim=) */
    int32_t x0eps_coef;           /* GaussianDiffusion(x0eps_coef=): posterior mean as c1*eps + c2*x0 (diffusion.py:137-140, 335-343) */
    int32_t t_fp32;               /* 1: the step tensor is fp32 as in p_sample_progressive (diffusion.py:421): s, t are fp32
                                     quotients and the timestep embedding is evaluated in fp32; 0: fp64 as in p_sample (:399) */
    int32_t reserved;             /* 0 */
    double intp_frac;
    double logsnr_min, logsnr_max;
    double w_guide;
    uint64_t seed;                /* on-device noise stream for ancestral sampling without injected noise */
} vdt_sampler_config;

typedef struct vdt_plan vdt_plan;

const char* vdt_last_error(void);
int vdt_version(void);

/* UNet(...) — unet.py:155-283.  Builds the block list, the weight table and (lazily) the workspace. */
int vdt_plan_create(const vdt_unet_config* cfg, vdt_plan** out);
void vdt_plan_destroy(vdt_plan* plan);

/* state_dict layout — SURVEY §8b.  Enumerate the expected keys / shapes (OIHW fp32 like the reference). */
int vdt_plan_num_weights(const vdt_plan* plan);
const char* vdt_plan_weight_name(const vdt_plan* plan, int index);
int vdt_plan_weight_shape(const vdt_plan* plan, int index, int64_t* shape4, int* ndim);

/* load_state_dict — generate.py:92-98.  `data` is contiguous fp32 with the reference's shape;
 * on_device != 0: device pointer, else host pointer.  Re-loading a key re-packs it. */
int vdt_plan_load_weight(vdt_plan* plan, const char* key, const float* data, int64_t numel, int on_device);
/* After all keys are loaded: pack bf16 GEMM operands.  Fails listing the first missing key. */
int vdt_plan_finalize(vdt_plan* plan);

/* UNet.forward(x, t, y) — unet.py:286-322.  x fp32 NCHW [B, in, R, R]; t fp64 [B]; y NULL, int64 [B] class ids
 * (0 = none), or fp32 multi-hot [B, num_classes] when the plan was created with multitags; out fp32 NCHW [B, out, R, R]. */
int vdt_unet_forward(vdt_plan* plan, const float* x, const double* t, const void* y, float* out, int32_t batch,
                     void* stream);

/* UNet.forward in .train() mode: the same network with nn.Dropout(drop_rate, inplace=True) active between act2 and conv2 of
 * every ResidualBlock (unet.py:135, 146).  The masks come from this library's Philox4x32-10 stream keyed by (seed, block,
 * element); drop_rate = 0 is bit-identical to vdt_unet_forward.  Forward only (the plan keeps no activations): the differentiable path is
 * the kernel tape of v-diffusion-torch_b200/training.py over the vdt_op_* entry points below. */
int vdt_unet_forward_train(vdt_plan* plan, const float* x, const double* t, const void* y, float* out, int32_t batch,
                           float drop_rate, uint64_t seed, void* stream);

/* GaussianDiffusion.p_sample(denoise_fn=UNet, shape, noise, label, use_ddim) — diffusion.py:394-414,
 * with p_sample_step (360-392) and p_mean_var (317-356) fused into one kernel per step and the step
 * captured as a CUDA graph.  noise fp32 [B, C, R, R] (x_T); label NULL, int64 [B], or fp32 multi-hot [B, num_classes] (multitags); step_noise NULL or
 * fp32 [T, B, C, R, R] (the per-step normal draws of diffusion.py:389, indexed by step); out fp32 [B, C, R, R]. */
int vdt_p_sample(vdt_plan* plan, const vdt_sampler_config* sc, const float* noise, const void* label,
                 const float* step_noise, float* out, int32_t batch, void* stream);
/* A slice of the same loop, in place on x (fp32 [B, C, R, R]): runs `num_steps` consecutive steps starting
 * at step index `first_step` (T-1 is the first step of a trajectory) — p_sample_step (diffusion.py:360-392)
 * applied num_steps times.  vdt_p_sample == copy noise, range(T-1, T), copy out.  Used by bench.py to time
 * K denoising steps of the real loop. */
int vdt_p_sample_range(vdt_plan* plan, const vdt_sampler_config* sc, float* x, const void* label,
                       const float* step_noise, int32_t batch, int32_t first_step, int32_t num_steps,
                       float* pred_x0 /* optional [B, C, R, R]: guided x0 prediction of the last step run
                                         (p_sample_step(return_pred=True); used by p_sample_progressive, 416-441) */,
                       void* stream);
/* Same with HOST buffers (pinned or pageable); copies in/out inside the call and synchronises. */
int vdt_p_sample_host(vdt_plan* plan, const vdt_sampler_config* sc, const float* noise, const void* label,
                      const float* step_noise, float* out, int32_t batch);

/* ---- training step: GaussianDiffusion.train_loss around the model call (diffusion.py:492-545) ------- */
enum { VDT_REWEIGHT_CONSTANT = 0, VDT_REWEIGHT_SNR = 1, VDT_REWEIGHT_SNR_TRUNC = 2, VDT_REWEIGHT_SNR_1PLUS = 3 };
/* Per-sample scalars from continuous times t in [0, 1] (HOST fp64 [B], Trainer.loss train_utils.py:137-147): out HOST
 * [B][16] floats, the vdt_step_coefficients slots 0-5, 11-13 (alpha, sigma, rsqrt(sigmoid l), exp(-l/2), sigmoid l,
 * sigmoid -l, l, rsqrt(sigmoid -l), exp(l/2)); only the schedule fields of sc are read. */
int vdt_train_coefficients(const vdt_sampler_config* sc, const double* t_host, int32_t batch, float* out);
/* q_sample (diffusion.py:242-245): x_t = x_0 * alpha_t + eps * sigma_t; device fp32 [B, C*H*W]; coef_dev = the table above on the device. */
int vdt_q_sample(const float* x0, const float* noise, const float* coef_dev, float* x_t, int32_t batch, int32_t chw, void* stream);
/* from_model_out_to_pred + the re-weighted MSE (diffusion.py:466-490, 518-541): loss fp32 [B]; grad_out (optional, shaped
 * like model_out) receives d loss.mean() / d model_out, the tensor autograd would hand to the UNet's backward pass. */
int vdt_train_loss(const float* model_out, const float* x0, const float* noise, const float* x_t, const float* coef_dev,
                   float* loss, float* grad_out, int32_t batch, int32_t c, int32_t hw, int32_t model_out_type,
                   int32_t reweight_type, void* stream);

/* The tail of generate.py's batch loop (generate.py:149): fp32 NCHW samples -> uint8 NHWC pixels,
 * (x * 127.5 + 127.5).clamp(0, 255).to(uint8).permute(0, 2, 3, 1).  x fp32 [B, C, HW], out uint8 [B, HW, C]; device pointers. */
int vdt_images_to_uint8(const float* x_nchw, uint8_t* out_nhwc, int32_t batch, int32_t c, int32_t hw, void* stream);

/* logsnr schedule + posterior coefficients — diffusion.py:42-112, 126-203 (host, fp64 with the reference's
 * fp32 rounding points).  out: [T][16] floats: alpha_t, sigma_t, rsqrt(sigmoid l_t), exp(-l_t/2), sigmoid(l_t),
 * sigmoid(-l_t), c1, c2, std, logvar, logsnr_s, logsnr_t, rsqrt(sigmoid -l_t), exp(l_t/2), x0eps_coef (0/1), 0.
 * With x0eps_coef and DDIM, c1/c2 are the LOGARITHMS 0.5*logsigmoid(-+l_s): the reference only exponentiates
 * them for eta != 0 (diffusion.py:180-182 vs 199) and p_sample always runs eta = 0; reproduced as is. */
int vdt_step_coefficients(const vdt_sampler_config* sc, float* out);

/* fp16 range monitor.  With fp16 operands (operand_dtype 0 / 2) every conversion of an unbounded value -- the raw
 * residual stream copied as the skip conv's operand, conv1 / q / k / v outputs kept in 16 bits -- saturates at +-65504
 * instead of overflowing to inf.  This returns how many (warp, 32x32-chunk) / thread events clamped a value since the
 * plan was finalized (or since the last call with reset != 0).  Non-zero means the checkpoint's activations leave the
 * fp16 range: use operand_dtype = 1 (bf16, fp32 exponent range) for it.  Synchronises the plan's stream. */
int vdt_plan_saturations(vdt_plan* plan, uint64_t* count, int reset);

/* Counters: kernels launched by this library since process start (graph replays count their nodes). */
uint64_t vdt_kernel_launches(void);

/* Algorithmic FLOPs (1 MAC = 2 FLOP) of one UNet.forward batch row, split like SURVEY §6:
 * convolutions (incl. 1x1), attention matmuls, embedding linears. */
int vdt_plan_flops(const vdt_plan* plan, double* conv, double* attn, double* linear);
/* FLOPs per UNet row the conv kernels execute (padded in_conv / out_conv GEMMs, sub-pixel upsampling convs; with
 * cfg_rows != 0 the label-independent prefix -- in_conv and conv1 of block 0 -- counted once per CFG row pair). */
int vdt_plan_conv_flops_executed(const vdt_plan* plan, int32_t cfg_rows, double* out);

/* Per-kernel-family device timing.  While enabled, steps run eagerly (no graph) with a CUDA-event pair
 * around every launch on the launching stream; vdt_profile_read returns accumulated milliseconds and launch
 * counts for VDT_PROF_* families and resets them. */
enum { VDT_PROF_CONV = 0, VDT_PROF_GROUPNORM = 1, VDT_PROF_ATTENTION = 2, VDT_PROF_OTHER = 3, VDT_PROF_FAMILIES = 4 };
int vdt_profile_enable(int on);
int vdt_profile_read(double* ms4, uint64_t* launches4);

/* ---- optimizer step of the reference trainer (train_utils.py:159-166) -------------------------------------------
 * nn.utils.clip_grad_norm_(params, max_norm) -> torch.optim.AdamW.step() (train.py:158) -> EMA.update() (utils.py:144-149),
 * without a host sync: vdt_grad_sq_accumulate adds one gradient tensor's sum of squares to a device double (zero it once
 * per step; fixed summation order), vdt_adamw_ema_step then updates one parameter tensor in place: g *= min(1, max_norm /
 * (sqrt(total) + 1e-6)) (skipped when grad_sq_total is null or max_norm <= 0), p *= 1 - lr * wd, Adam moments, bias-corrected
 * update with `step` = the 1-based step count, and shadow += (1 - ema_decay) * (p - shadow) when ema_shadow is given
 * (ema_decay = min(decay, (1 + num_updates) / (10 + num_updates)) is the caller's, utils.py:146). */
int64_t vdt_grad_sq_scratch_bytes(int64_t n);
int vdt_grad_sq_accumulate(const float* grad, int64_t n, void* scratch, int64_t scratch_bytes, double* sq_accum, void* stream);
int vdt_adamw_ema_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, float* ema_shadow, int64_t n,
                       double lr, double beta1, double beta2, double eps, double weight_decay, int32_t step,
                       const double* grad_sq_total, double max_norm, double ema_decay, void* stream);

/* ---- kernel-level entry points (used by the parity tests; device pointers) ------------------------- */
/* f16: 1 = fp16 operands, 0 = bf16.  conv2d(k x k, pad k/2) on 16-bit NHWC input with fp32 OIHW weights -> fp32 NHWC
 * (+bias, +residual), or 16-bit NHWC when out16 is given.  stats_out (optional): GroupNorm partial statistics,
 * float2 (sum, sum of squares) per (slab of 32 rows, stat_cols = 4 or 2 columns): [rows/32][cout/stat_cols]. */
int vdt_op_conv(const void* x_16_nhwc, int32_t batch, int32_t h, int32_t w, int32_t cin, const float* w_oihw,
                int32_t cout, int32_t ksize, const float* bias, const float* residual, float* out_nhwc, int32_t f16,
                void* out16, void* stats_out, int32_t stat_cols, void* stream);
/* Backward of a conv layer (F.conv2d under autograd, modules.py:141-144).
 * dgrad: dX fp32 NHWC [B, H, W, cin] = conv(dY, W rotated 180 degrees with its channel axes swapped) on the forward kernel;
 *        dy 16-bit NHWC [B, H, W, cout], w fp32 OIHW [cout, cin, k, k].
 * wgrad: dW fp32 OIHW [cout, cin, k, k] = sum over pixels of dY[p][co] * X[p + tap][ci] (tcgen05, both operands MN-major,
 *        K split over CTAs with a fixed-order reduction); optional dbias fp32 [cout] = sum over pixels of dY.  x, dy 16-bit
 *        NHWC.  Needs cout % 128 == 0, cin % 64 == 0 and feature maps that tile into 128-pixel boxes. */
int vdt_op_conv_dgrad(const void* dy_16_nhwc, int32_t batch, int32_t h, int32_t w, int32_t cin, const float* w_oihw, int32_t cout,
                      int32_t ksize, float* dx_nhwc, int32_t f16, void* stream);
int vdt_op_conv_wgrad(const void* x_16_nhwc, const void* dy_16_nhwc, int32_t batch, int32_t h, int32_t w, int32_t cin, int32_t cout,
                      int32_t ksize, float* dw_oihw, float* dbias, int32_t f16, void* stream);
/* Rows of the statistics table vdt_op_conv writes per image: stats_out is [batch * slabs (+ 4 rows of slack: the last
 * M tile always writes its four quarters)][cout / stat_cols] (sum, sumsq)
 * float2, one slab per 32-row quarter of an M tile of the conv kernel's image-aligned tiling; 0 = this feature-map size
 * has no such layout (the GroupNorm then makes its own statistics pass). */
int vdt_stat_slabs_per_image(int32_t h, int32_t w);
/* GroupNorm(32, 1e-6) [+FiLM] [+SiLU] [+resample 0 none / 1 avgpool2 / 2 nearest x2] over concat(src1, src2).
 * stats1/stats2 (optional): partial statistics from vdt_op_conv for src1/src2 -> single-pass kernel;
 * in16: src1 is 16-bit (needs stats1). */
int vdt_op_groupnorm(const void* src1, int32_t c1, const float* src2, int32_t c2, int32_t batch, int32_t h, int32_t w,
                     const float* gamma, const float* beta, const float* film, int32_t film_stride, int32_t film_off,
                     int32_t silu, int32_t resample, void* out_act_16, void* out_raw_16, float* out_res,
                     int32_t f16, const void* stats1, const void* stats2, int32_t stat_cols, int32_t in16, void* stream);
/* GroupNorm + SiLU with training dropout on a plain fp32 NHWC tensor (two-pass statistics): the op norm2 -> act2 -> dropout
 * of a ResidualBlock in .train() mode, exposed for the mask-statistics test. */
int vdt_op_groupnorm_dropout(const void* src1, int32_t c1, int32_t batch, int32_t h, int32_t w, const float* gamma,
                             const float* beta, int32_t silu, void* out_act_16, int32_t f16, float drop_p, uint64_t seed,
                             int32_t layer, void* stream);
/* Backward of norm -> FiLM -> SiLU -> dropout (the activation chain of a ResidualBlock, unet.py:131-135, 143-146): x and
 * grad_out fp32 [batch, h*w, c] (c in 128 / 256 / 512 / 1024), grad_out = gradient at the activation; film = null or
 * [batch][2c] (shift | scale per sample); (drop_p, seed, layer) name the forward's dropout stream (drop_p = 0: none).
 * Writes grad_x [batch, h*w, c], grad_gamma / grad_beta [c] and, when given, grad_film [batch][2c] (d shift | d scale).
 * Replaces autograd through F.group_norm / F.silu / F.dropout (torch/nn/functional.py) on this chain. */
int vdt_op_groupnorm_backward(const float* x, const float* grad_out, int32_t c, int32_t batch, int32_t h, int32_t w,
                              const float* gamma, const float* beta, const float* film, int32_t silu, float drop_p,
                              uint64_t seed, int32_t layer, float* grad_x, float* grad_gamma, float* grad_beta,
                              float* grad_film, void* stream);
/* The same GroupNorm kernel with the FiLM table and the training dropout stream both given: norm2 -> FiLM -> act2 -> dropout
 * of a ResidualBlock in .train() mode (unet.py:143-146) on a plain fp32 NHWC tensor; drop_p = 0: no dropout.  The mask is the
 * one vdt_op_groupnorm_backward regenerates from (seed, layer). */
int vdt_op_groupnorm_train(const void* src1, int32_t c1, int32_t batch, int32_t h, int32_t w, const float* gamma,
                           const float* beta, const float* film, int32_t film_stride, int32_t film_off, int32_t silu,
                           void* out_act_16, int32_t f16, float drop_p, uint64_t seed, int32_t layer, void* stream);
/* F.linear(x, W, b) [+ SiLU] in fp32 (modules.py:77-78; time_embed, class_embed and every ResidualBlock.fc, unet.py:142,
 * 287-295): x [rows, K], W [N, K], b [N] (required) -> out [rows, N]; K <= 1536.  The composed training step also forms the
 * backward products with it (dX = dY W: pass W^T; dW = dY^T X: pass the transposed operands). */
int vdt_op_linear(const float* x, const float* w, const float* b, float* out, int32_t rows, int32_t k, int32_t n,
                  int32_t silu_out, void* stream);
/* get_timestep_embedding(t, dim) (functions.py:10-29) for fp64 t [rows] -> fp32 [rows, dim] (sin | cos halves). */
int vdt_op_timestep_embedding(const double* t, float* out, int32_t rows, int32_t dim, void* stream);
/* attention on qkv 16-bit [B*N, 3*hid] (q | k | v thirds, heads contiguous inside each, the layout proj_in writes)
 * -> 16-bit [B*N, hid]; any N >= 1 (ragged last key / query tiles are masked) */
int vdt_op_attention(const void* qkv_16, void* out_16, int32_t batch, int32_t n, int32_t heads,
                     int32_t d, int32_t f16, void* stream);
/* Backward of the attention core (autograd through the einsum / softmax / einsum of unet.py:55-64), reference-grade: fp32 on
 * CUDA cores, two passes (per query row: softmax statistics, D = dO . O, dQ; per key row: dK, dV), nothing N x N stored.
 * qkv fp32 [B*N, 3*hid] (q | k | v thirds, heads contiguous), grad_out fp32 [B*N, hid] -> grad_qkv fp32 [B*N, 3*hid]. */
int vdt_op_attention_backward(const float* qkv, const float* grad_out, float* grad_qkv, int32_t batch, int32_t n, int32_t heads,
                              int32_t d, void* stream);
/* one sampler update with explicit step index; coef = one row of vdt_step_coefficients (host pointer). */
int vdt_op_sampler_step(const float* model_out, const float* x_t, const float* noise, float* x_s, int32_t batch,
                        int32_t c, int32_t hw, int32_t cfg, int32_t model_out_type, int32_t step, const float* coef_host,
                        float w, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* VDT_B200_H */
