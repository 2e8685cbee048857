/* lr_b200.h — C ABI of liblr_b200.so: the B200-native (sm_100a) implementation of LeftRefill's DDIM/UNet hot path.
 *
 * Plain pointers and sizes only; no torch / C++ types. All pointers are CUDA device pointers unless stated.
 * Every entry point returns 0 on success, non-zero on failure; lr_last_error() then gives the reason (thread local).
 * Calls are asynchronous on the `stream` argument (a cudaStream_t passed as void*); no hidden device synchronisation.
 * Threading: one host thread per engine handle at a time (the reference drives its model from a single Python thread).
 * Per-device lazy initialisation (kernel attributes, SM count) is keyed by the CURRENT CUDA device and atomic; the
 * op-level entry points keep one split-K scratch buffer per device (callers on one device use one stream at a time);
 * the engine (lr_unet_* / lr_vae_*) owns its memory per handle and must be used with its device current.
 * Citations are relative to the reference tree (ewrfcas/LeftRefill @ 893c3220).
 */
#ifndef LR_B200_H_
#define LR_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* 2: lr_vae_*, lr_unet_cfg.use_sep, lr_unet_set_c_input, GroupNorm-fusion entry points (round 2).
 * 3: lr_groupnorm_scratch_bytes (the scratch of lr_groupnorm_f16 now depends on P); GEGLU weight / bias rows in groups of
 *    four (value, value', gate, gate') - lr_repack_linear_weight(geglu = 1) writes that order. */
#define LR_B200_ABI_VERSION 3

/* ---- library ---------------------------------------------------------------------------------------------- */
int lr_abi_version(void);
const char* lr_last_error(void);
/* number of CUDA kernels launched by this library since the last reset (bench.py's gpu_launches) */
long long lr_launch_count(void);
void lr_launch_count_reset(void);
/* Kernels replayed through a captured CUDA graph are launched by the driver, not by this library: the host side adds
 * the number of kernel nodes of every replay here so that lr_launch_count() keeps counting executed kernels. */
void lr_launch_count_add(long long n);
/* Bring-up aid (no reference counterpart): when LR_GEMM_TRACE / LR_ATTN_TRACE is set in the environment before the first
 * op, CTA 0 of the GEMM / attention kernels records clock64() samples per warp role and phase; this copies them to HOST
 * memory `dst` (synchronises the device) and optionally clears the buffer. Fails when tracing is not enabled. */
int lr_debug_read_trace(void* dst, long long bytes, int clear);

/* ---- UNet engine: replaces UNetModel.forward (ldm/modules/diffusionmodules/openaimodel.py:755-787) ------------ */
typedef struct lr_unet lr_unet;

/* Mirrors the constructor kwargs the yaml passes (openaimodel.py:442-472, configs/ref_inpainting.yaml:20-36). */
typedef struct lr_unet_cfg {
  int in_channels;         /* 9 */
  int model_channels;      /* 320 */
  int out_channels;        /* 4 */
  int num_levels;          /* len(channel_mult) */
  int channel_mult[8];     /* 1,2,4,4 */
  int num_res_blocks[8];   /* per level */
  int attention_ds[8];     /* attention_resolutions (downsample rates), n_attention_ds entries */
  int n_attention_ds;
  int num_head_channels;   /* 64 (d_head); the attention kernel requires 64 */
  int transformer_depth;   /* 1 */
  int context_dim;         /* 1024 */
  int use_linear_in_transformer; /* 1: nn.Linear proj_in/out, 0: 1x1 conv (same math) */
  /* multiview variant (ldm/modules/multiview_attention.py:431-468): self-attention runs over view_num views */
  int view_num;            /* 1 = plain UNetModel */
  int concat_target;       /* 0/1 (multiview only) */
  /* NVSUnetModel (inpainting_ldm/NVS_ldm.py:22-31): 1 = a learned separator column `sep_token.<channels>` is inserted
   * between the two halves of the stitched canvas before every non-resampling block and removed after it (:57-97);
   * the weight table then ends with sep_token.{in_channels, ...} in the reference's order. */
  int use_sep;
} lr_unet_cfg;

/* Host-only: builds the layer graph and the weight table. Device memory is allocated lazily. */
int lr_unet_create(const lr_unet_cfg* cfg, lr_unet** out);
void lr_unet_destroy(lr_unet* h);

/* Weight table = the reference state-dict keys relative to UNetModel ("input_blocks.1.0.in_layers.2.weight", ...),
 * i.e. `model.diffusion_model.<name>` in SD2 checkpoints (test_inpainting.py:26-53). */
int lr_unet_num_weights(const lr_unet* h);
const char* lr_unet_weight_name(const lr_unet* h, int index);
/* shape_out receives up to 4 dims in PyTorch order; returns ndim (or -1). */
int lr_unet_weight_shape(const lr_unet* h, int index, int64_t shape_out[4]);
/* Upload one fp32 tensor in its PyTorch layout (OIHW conv, [out,in] linear); the engine repacks it to fp16 GEMM
 * layout on `stream`. `data` must stay valid until the stream reaches this point. */
int lr_unet_set_weight(lr_unet* h, const char* name, const float* data, const int64_t* shape, int ndim, void* stream);
/* Number of weights not yet uploaded (forward refuses to run while > 0). */
int lr_unet_missing_weights(const lr_unet* h);

/* Cross-attention K/V of `context` [n, L, context_dim] fp32 are step-invariant (attention.py:170-171): compute them
 * once for all layers. lr_unet_forward with context == NULL reuses this cache. */
int lr_unet_set_context(lr_unet* h, const float* context, int n, int L, void* stream);

/* NVS input refinement (NVS_ldm.py:49,64-68): c_input [n, model_channels, H, Wc] fp32 NCHW is added to the output of
 * the input conv of every following lr_unet_forward, over the whole feature width (Wc == Wh) or over its right part
 * (Wc == Wh - Wh/2, columns from Wh/2 on), where Wh = W (+1 with use_sep) and W is the UNet input width. NULL clears it.
 * The engine keeps its own fp16 copy: `c_input` may be released once `stream` has passed this call. */
int lr_unet_set_c_input(lr_unet* h, const float* c_input, int n, int C, int H, int Wc, int W, void* stream);

/* x [n, in_channels, H, W] fp32 NCHW, timesteps [n] int64, context [n, L, context_dim] fp32 or NULL (cached),
 * out [n, out_channels, H, W] fp32 NCHW. */
int lr_unet_forward(lr_unet* h, const float* x, const int64_t* timesteps, const float* context, int L, float* out,
                    int n, int H, int W, void* stream);
/* Classifier-free-guidance pair (p_sample_ddim, ldm/models/diffusion/ddim.py:317-343): the UNet batch is
 * [uncond | cond] with IDENTICAL x, timesteps and c_concat in both halves; only the cross-attention context differs.
 * x [n_canvas, in_channels, H, W], timesteps [n_canvas]; the 2*n_canvas contexts (uncond first) must have been cached
 * with lr_unet_set_context; out [2*n_canvas, out_channels, H, W]. Everything before the first cross-attention (input
 * conv, first ResBlock, first self-attention) is computed once and replicated; the result is bit-identical to
 * lr_unet_forward on the doubled batch. */
int lr_unet_forward_cfg_pair(lr_unet* h, const float* x, const int64_t* timesteps, float* out, int n_canvas, int H,
                             int W, void* stream);
/* Per-op CUDA-event timing of forward (bench.py's roofline): when enabled, every forward records an event between
 * plan steps on `stream`. lr_unet_read_profile waits for the last profiled forward and sums by kernel class:
 * 0 = gemm_conv_kernel (all convs + linears), 1 = attention_kernel, 2 = GroupNorm kernels, 3 = LayerNorm, 4 = other. */
int lr_unet_set_profiling(lr_unet* h, int enable);
int lr_unet_read_profile(lr_unet* h, double ms_by_class[5], double flops_by_class[5], int steps_by_class[5]);
/* Per plan step: duration of the last profiled forward (ms, -1 if none), algorithmic FLOPs, class, description. */
int lr_unet_num_steps(const lr_unet* h);
int lr_unet_step_info(lr_unet* h, int index, double* ms, double* flops, int* cls, char* desc, int desc_len);
/* Incremented whenever the engine rebuilds its static plan (new batch / latent shape / context length): device
 * pointers baked into a captured CUDA graph of lr_unet_forward* are valid only while this value is unchanged. */
long long lr_unet_plan_generation(const lr_unet* h);
/* Algorithmic FLOPs (2*M*N*K convs/linears + 4*Tq*Tk*d attention) of the last planned forward. */
double lr_unet_last_flops(const lr_unet* h);
/* Bytes of device memory held by the engine (weights + activation plan). */
long long lr_unet_device_bytes(const lr_unet* h);

/* ---- first-stage decoder: replaces AutoencoderKL.decode (ldm/models/autoencoder.py:87-90: decoder(post_quant_conv(z)))
 * with the Decoder of ldm/modules/diffusionmodules/model.py:547-653, as called by LatentDiffusion.decode_first_stage
 * (ldm/models/diffusion/ddpm.py:835-843, which first multiplies z by 1 / scale_factor = z_scale here). ----------- */
typedef struct lr_vae lr_vae;
/* first_stage_config.params of configs/ref_inpainting.yaml:38-58 (ddconfig + embed_dim); attn_resolutions must be
 * empty (only the mid-block attention of the SD VAE exists), resamp_with_conv true, tanh_out / give_pre_end false. */
typedef struct lr_vae_cfg {
  int ch;              /* 128 */
  int out_ch;          /* 3 */
  int num_levels;      /* len(ch_mult) */
  int ch_mult[8];      /* 1,2,4,4 */
  int num_res_blocks;  /* 2 */
  int z_channels;      /* 4 */
  int embed_dim;       /* 4 */
} lr_vae_cfg;
int lr_vae_create(const lr_vae_cfg* cfg, lr_vae** out);
void lr_vae_destroy(lr_vae* h);
/* weight table = AutoencoderKL state-dict keys "decoder.*" and "post_quant_conv.*" (SD checkpoints:
 * first_stage_model.<name>), fp32 PyTorch layouts, same protocol as lr_unet_set_weight */
int lr_vae_num_weights(const lr_vae* h);
const char* lr_vae_weight_name(const lr_vae* h, int index);
int lr_vae_weight_shape(const lr_vae* h, int index, int64_t shape_out[4]);
int lr_vae_set_weight(lr_vae* h, const char* name, const float* data, const int64_t* shape, int ndim, void* stream);
int lr_vae_missing_weights(const lr_vae* h);
/* z [n, embed_dim, H, W] fp32 NCHW -> out [n, out_ch, 8H, 8W] fp32 NCHW (2^(num_levels-1) = 8 for the SD VAE) */
int lr_vae_decode(lr_vae* h, const float* z, float z_scale, float* out, int n, int H, int W, void* stream);
double lr_vae_last_flops(const lr_vae* h);
long long lr_vae_device_bytes(const lr_vae* h);
int lr_vae_num_steps(const lr_vae* h);

/* ---- fused CFG + DDIM update: replaces p_sample_ddim's tail (ldm/models/diffusion/ddim.py:343,359-381) ----------
 * eps_uncond/eps_cond: the two halves of the CFG-doubled UNet output (eps_cond may be NULL: no guidance).
 * noise may be NULL when sigma == 0. All tensors fp32 with `numel` elements. */
int lr_ddim_update(const float* x, const float* eps_uncond, const float* eps_cond, const float* noise, float cfg_scale,
                   float a_t, float a_prev, float sigma_t, float sqrt_one_minus_at, float temperature, int64_t numel,
                   float* x_prev, float* pred_x0, void* stream);
/* Same update with the per-step scalars in DEVICE memory, coef = {cfg_scale, a_t, a_prev, sigma_t, sqrt(1 - a_t)}
 * (fp32): the launch is identical for every DDIM step, which lets the drop-in DDIMSampler capture ONE CUDA graph of
 * (UNet forward + update) and replay it for all steps of ddim_sampling (ddim.py:253-296). x_prev may alias x. */
int lr_ddim_update_dev(const float* x, const float* eps_uncond, const float* eps_cond, const float* noise,
                       const float* coef, float temperature, int64_t numel, float* x_prev, float* pred_x0,
                       void* stream);

/* ---- op-level entry points (what CrossAttention / ResBlock / SpatialTransformer mirrors call stand-alone) ------
 * Activations are NHWC fp16: row = (n*H + y)*W + x, channels contiguous. */

/* force_block_n (testing hook, 0 = heuristic): 1000*cg + block_n with cg in {0: heuristic, 1: single CTA, 2: CTA pair
 * (tcgen05 cta_group::2)} and block_n a multiple of 32 (0 = heuristic).
 * out[M, n_out] = A[M, K] * W[n_out(*2 if geglu), K]^T (+bias) (+residual) ; geglu: W/bias rows in groups of four
 * (value_2k, value_2k+1, gate_2k, gate_2k+1), the order lr_repack_linear_weight(geglu = 1) writes; n_out even
 * (value_j, gate_j) and out[:, j] = v_j * gelu(g_j) (attention.py:51-58). fp16 in/out, fp32 accumulate. */
int lr_linear_f16(const void* a, int lda, int M, int K, const void* w, int ldw, int n_cols, const float* bias,
                  const void* residual, int ld_res, void* out, int ld_out, int geglu, int force_block_n, void* stream);
/* 3x3 conv, pad 1, stride 1|2, over the channel concat of x0 [n,h,w,c0] and optional x1 [n,h,w,c1];
 * wt [cout, 9*(c0+c1)] fp16 with k = (ky*3+kx)*(c0+c1) + c; bias [cout]; bias_img [n, cout] (time embedding);
 * residual/out [n*ho*wo, cout] fp16. */
int lr_conv3x3_f16(const void* x0, int c0, const void* x1, int c1, int n, int h, int w, int stride, const void* wt,
                   int cout, const float* bias, const float* bias_img, const void* residual, void* out,
                   int force_block_n, void* stream);
/* ---- GroupNorm fused into its producer and its consumer (ResBlock in_layers / out_layers = GroupNorm32 -> SiLU -> conv,
 * openaimodel.py:200-204,224-231,254-274; SpatialTransformer norm -> proj_in, attention.py:399-404) -----------------
 * A GroupNorm is y = x * scale[n, c] + shift[n, c] once its statistics are known. The kernel that WRITES a tensor can
 * leave per-tile (sum, sum of squares) partials per channel (stats_out, lr_conv_stats_rows() x cout float2 entries;
 * *stats_ppi receives the table rows per image, 0 if this geometry cannot produce them); lr_gn_finalize reduces the
 * partials of one or two (channel-concatenated) tensors in a fixed order to scale / shift [n, c0 + c1]; the kernel that
 * READS the tensor applies them (and the SiLU) to its activation tiles in shared memory (gn_scale / gn_shift non-NULL;
 * 3x3: stride 1 and an image of at least 16 rows x 8 columns). The normalised tensor never exists in memory. */
long long lr_conv_stats_rows(int n, int h, int w, int stride, int taps, int rows_per_img);
int lr_gn_conv3x3_f16(const void* x0, int c0, const void* x1, int c1, int n, int h, int w, const float* gn_scale,
                      const float* gn_shift, int silu, const void* wt, int cout, const float* bias, const float* bias_img,
                      const void* residual, void* out, float* stats_out, int* stats_ppi, int force_block_n, void* stream);
/* token matrix a [M, K] whose image index is row / rows_per_img (rows_per_img = H*W) */
int lr_gn_linear_f16(const void* a, int M, int K, int rows_per_img, const float* gn_scale, const float* gn_shift,
                     int silu, const void* w, int n_cols, const float* bias, const void* residual, void* out,
                     float* stats_out, int* stats_ppi, int force_block_n, void* stream);
int lr_gn_finalize(const float* part0, int ppi0, int c0, const float* part1, int ppi1, int c1, int n, int P, int groups,
                   float eps, const float* gamma, const float* beta, float* scale, float* shift, void* stream);

/* Upsample.forward (openaimodel.py:108-116; model.py:62-66) = F.interpolate(x, scale_factor=2, mode="nearest") + 3x3 conv,
 * without materialising the upsampled tensor: four 2x2-tap convs on x [n, h, w, cin], one per output phase, with the 3x3
 * taps that read the same source pixel pre-summed (4/9 of the FLOPs). wt [cout, 9*cin] as for lr_conv3x3_f16;
 * wfold_scratch: 16*cout*cin fp16 elements; out [n, 2h, 2w, cout]. cout % 32 == 0. */
int lr_upsample2x_conv3x3_f16(const void* x, int n, int h, int w, int cin, const void* wt, int cout, const float* bias,
                              void* wfold_scratch, void* out, void* stream);
/* softmax(q k^T * scale) v, d_head = 64. q [batch*tq, ldq] (head h at columns q_col0 + 64h), likewise k, v over tk
 * tokens; out [batch*tq, ld_out]. Replaces attention.py:176-195. */
int lr_attention_f16(const void* q, int ldq, int q_col0, const void* k, int ldk, int k_col0, const void* v, int ldv,
                     int v_col0, void* out, int ld_out, int batch, int heads, int tq, int tk, float scale,
                     void* stream);
/* GroupNorm(groups, eps) [+ SiLU] over concat(x0, x1) -> out [n, P, c0+c1] fp16 (util.py:217-219, attention.py:90-91).
 * scratch: lr_groupnorm_scratch_bytes(n, groups, P) bytes of device memory, 16-byte aligned (per-chunk partial sums and
 * the grid-barrier counters of the single-launch kernel; contents need not be preserved between calls). */
size_t lr_groupnorm_scratch_bytes(int n, int groups, int P);
int lr_groupnorm_f16(const void* x0, int c0, const void* x1, int c1, int n, int P, int groups, float eps,
                     const float* gamma, const float* beta, int silu, void* out, void* scratch, void* stream);
int lr_layernorm_f16(const void* x, int M, int C, const float* gamma, const float* beta, float eps, void* out,
                     void* stream);
/* layout helpers for the NCHW fp32 boundary */
int lr_nchw_f32_to_nhwc_f16(const float* x, int n, int c, int h, int w, void* out, void* stream);
int lr_nhwc_f16_to_nchw_f32(const void* x, int ld, int n, int c, int h, int w, float* out, void* stream);
/* fp32 OIHW / [out,in] -> fp16 GEMM layouts used above */
int lr_repack_conv3x3_weight(const float* w_oihw, int cout, int cin, void* out, void* stream);
int lr_repack_linear_weight(const float* w, int n_out, int n_in, int geglu, void* out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* LR_B200_H_ */
