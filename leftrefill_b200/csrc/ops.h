// Host-side op layer: builds TMA tensor maps + launch geometry once ("op" objects), launches them on a stream.
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>

namespace lr {

// ---- error plumbing (C-ABI returns int status; text via lr_last_error) ----
void set_error(const std::string& msg);
const char* last_error();
#define LR_CUDA(expr)                                                                                  \
  do {                                                                                                 \
    cudaError_t _e = (expr);                                                                           \
    if (_e != cudaSuccess) {                                                                           \
      ::lr::set_error(std::string(#expr) + ": " + cudaGetErrorString(_e));                             \
      return 1;                                                                                        \
    }                                                                                                  \
  } while (0)
#define LR_CHECK(cond, msg)                                                                            \
  do {                                                                                                 \
    if (!(cond)) {                                                                                     \
      ::lr::set_error(std::string(msg));                                                               \
      return 1;                                                                                        \
    }                                                                                                  \
  } while (0)
#define LR_TRY(expr)                                                                                   \
  do {                                                                                                 \
    int _r = (expr);                                                                                   \
    if (_r != 0) return _r;                                                                            \
  } while (0)

// ---- GEMM / implicit conv ----
struct ConvSpec {
  const __half* a0 = nullptr;  // source 0, NHWC [n_img, in_h, in_w, c0] with row stride lda0 (elements)
  int c0 = 0, lda0 = 0;
  const __half* a1 = nullptr;  // optional source 1 (skip connection), same spatial dims
  int c1 = 0, lda1 = 0;
  int n_img = 1, in_h = 1, in_w = 1;
  int stride = 1;  // 1 or 2
  int taps = 1;    // 1 (linear / 1x1 conv), 9 (3x3, pad 1) or 4 (2x2 phase of a folded nearest-x2 upsample + 3x3 conv:
                   // taps at input offsets (up_oy + a, up_ox + b), a, b in {0, 1}; needs out_sx / out_sy / out_sn)
  int up_oy = 0, up_ox = 0;
  // output pixel strides in elements (0 = dense [n, Ho, Wo, ld_out]); `out` is the first pixel written
  size_t out_sx = 0, out_sy = 0, out_sn = 0;
  const __half* w = nullptr;  // [ncols, ldw] fp16, k = tap*(c0+c1) + c
  int ldw = 0;
  int ncols = 0;
  const float* bias = nullptr;
  const float* bias_img = nullptr;  // [n_img, ld_bias_img] per-image additive bias (0 -> ld = ncols)
  int ld_bias_img = 0;
  const __half* residual = nullptr;
  int ld_res = 0;
  __half* out = nullptr;
  int ld_out = 0;
  int geglu = 0;
  float out_scale = 1.0f;  // multiplies the final value (softmax scale of the VAE AttnBlock logits)
  const float* ln_stats = nullptr;  // folded LayerNorm: [M] float2 (mean, rstd) of the input rows, or nullptr
  const float* ln_s = nullptr;      //                   [ncols] column sums of the gamma-folded weight
  // ... or per-row partials left by the producing GEMM's epilogue (rowstats_out of that op): [M][ln_ld] float2, the first
  // ln_slots entries of a row are summed; mean / rstd over c0 channels with ln_eps
  const float* ln_part = nullptr;
  int ln_slots = 0, ln_ld = 0;
  float ln_eps = 1e-5f;
  // producer side: (sum, sum of squares) per output row and (N-tile, warp half) -> rowstats_out [M][rowstats_ld] float2;
  // ConvOp::rowstats_slots entries per row are written (rowstats_ld >= 2 * ceil(ncols / 32) is always enough)
  float* rowstats_out = nullptr;
  int rowstats_ld = 0;
  // GroupNorm fusion (gemm_tc.cuh header): the input is x * xf_scale[n, c] + xf_shift[n, c] (+ SiLU), applied in shared
  // memory by the transform warps. Needs a halo-mode 3x3 conv (stride 1, image >= 16 rows x 8 columns) or taps == 1.
  const float* xf_scale = nullptr;  // [n_img (or M / xf_rows_per_img), c0 + c1]
  const float* xf_shift = nullptr;
  int xf_silu = 0;
  int xf_rows_per_img = 0;  // taps == 1 on a token matrix: rows per image
  // per-tile-half (sum, sum of squares) of the fp16 output per channel -> stats_out [conv_stats_rows()][ncols] float2;
  // written only when the op ends up eligible (ConvOp::stats_ok), i.e. TMA-store epilogue, no split-K, every 128-row tile
  // inside one image (stats_rows_per_img: rows per image of a token-matrix output, 0 = conv geometry)
  float* stats_out = nullptr;
  int stats_rows_per_img = 0;
  float* workspace = nullptr;  // optional split-K scratch (fp32 partial tiles), workspace_bytes >= 3 * M_out * ncols * 4
  size_t workspace_bytes = 0;
  int force_block_n = 0;  // testing hook: 0 = heuristic
  int force_cg = 0;       // testing hook: 0 = heuristic, 1 = single CTA, 2 = CTA pair (cta_group::2)
};
struct ConvOp {
  alignas(64) unsigned char params[2048];  // GemmParams (opaque here so headers stay CUDA-kernel free)
  int grid = 0, smem = 0;
  int out_h = 0, out_w = 0;
  int block_n = 0, stages = 0, tiles = 0, cg = 1;
  // split-K (ksplit > 1): the GEMM kernel writes fp32 partials, launch_conv_op then runs the reduction
  int ksplit = 1, ncols = 0, rows_per_img = 0, ld_bias_img = 0, ld_res = 0, ld_out = 0;
  size_t m_out = 0;
  float* partial = nullptr;
  const float* bias = nullptr;
  const float* bias_img = nullptr;
  const __half* residual = nullptr;
  __half* out = nullptr;
  double flops = 0;
  int xf = 0;         // launches the transform-warp instantiation
  int stats_ok = 0;   // stats_out is written by this op
  int stats_ppi = 0;  // statistics rows per image (2 per 128-row tile)
  int rowstats_slots = 0;  // entries per row written to rowstats_out (0: none)
};
// 3x3 convs with at most this many OUTPUT pixels per image run split-K over the taps when scratch is provided
constexpr int kSplitKMaxPixels = 256;
// rows of the statistics table a conv / linear with this geometry needs (2 per M-tile); 0 if it can never produce them
size_t conv_stats_rows(const ConvSpec& s);
// true if build_conv_op will run this 3x3 conv in halo mode (the only 3x3 mode the transform warps support)
bool conv_is_halo(const ConvSpec& s);
int build_conv_op(ConvOp* op, const ConvSpec& s);
int launch_conv_op(const ConvOp& op, cudaStream_t st);
float* op_level_workspace(size_t bytes);

// ---- attention ----
struct AttnSpec {
  const __half* q = nullptr; int ldq = 0, q_col0 = 0;
  const __half* k = nullptr; int ldk = 0, k_col0 = 0;
  const __half* v = nullptr; int ldv = 0, v_col0 = 0;
  __half* out = nullptr; int ld_out = 0;
  int batch = 1, heads = 1, tq = 0, tk = 0;
  float scale = 0.125f;
};
struct AttnOp {
  alignas(64) unsigned char params[1024];
  dim3 grid;
  int persist = 0;  // attention_persist_kernel (grid = CTAs walking the work items) instead of one CTA per item
  double flops = 0;
};
int build_attn_op(AttnOp* op, const AttnSpec& s);
int launch_attn_op(const AttnOp& op, cudaStream_t st);

// ---- normalisation / elementwise launchers ----
// GroupNorm over concat(x0[c0], x1[c1]) -> out [n, P, c0+c1]; scratch: groupnorm_scratch_bytes(n, groups, P) bytes whose
// first 16 bytes (grid-barrier counters) must be zero on entry: zeroed here unless the caller guarantees
// scratch_is_zero; the kernel leaves them zero.
size_t groupnorm_scratch_bytes(int n_img, int groups, int P);
int launch_groupnorm(const __half* x0, int c0, const __half* x1, int c1, int n_img, int P, int groups, float eps,
                     const float* gamma, const float* beta, int do_silu, void* scratch, int scratch_is_zero,
                     __half* out, cudaStream_t st);
// GroupNorm statistics from producer partials -> per-(image, channel) affine coefficients (elementwise.cuh)
int launch_gn_finalize(const float* part0, int ppi0, int c0, const float* part1, int ppi1, int c1, int n_img, int P,
                       int groups, float eps, const float* gamma, const float* beta, float* scale, float* shift,
                       cudaStream_t st);
// y = [silu](x * scale[n, c] + shift[n, c]) over concat(x0, x1): GroupNorm apply from finalized producer statistics
int launch_gn_apply_coef(const __half* x0, int c0, const __half* x1, int c1, int n_img, int P, const float* scale,
                         const float* shift, int do_silu, __half* out, cudaStream_t st);
int launch_layernorm(const __half* x, int M, int C, const float* gamma, const float* beta, float eps, __half* out,
                     cudaStream_t st);
// (mean, rstd) per row -> stats [M] float2; the normalisation itself is folded into the consuming GEMM
int launch_layernorm_stats(const __half* x, int M, int C, float eps, float* stats, cudaStream_t st);
// the same (mean, rstd) table from the per-row partials a GEMM epilogue left (ConvSpec::rowstats_out)
int launch_ln_rows_finalize(const float* part, int ld, int slots, int M, int C, float eps, float* stats, cudaStream_t st);
int launch_ln_fold(const __half* w, int rows, int K, const float* gamma, const float* beta, const float* bias, __half* wf,
                   float* s_out, float* bf_out, cudaStream_t st);
int launch_im2col_nchw_f32(const float* x, int n_img, int cin, int H, int W, int kpad, __half* out, cudaStream_t st);
int launch_upsample2x(const __half* x, int n_img, int H, int W, int C, __half* out, cudaStream_t st);
int launch_mv_gather(const __half* src, int ld_src, int ncols, int b, int v, int hh, int side, __half* dst,
                     cudaStream_t st);
int launch_mv_scatter(const __half* src, int ncols, int b, int v, int hh, int side, __half* dst, cudaStream_t st);
// NVSUnetModel(use_sep=True) separator column (NVS_ldm.py:57-71) and c_input staging (:64-68)
int launch_sep_insert(const __half* x, const float* sep, int n_img, int H, int W, int C, __half* out, cudaStream_t st);
int launch_sep_remove(const __half* x, int n_img, int H, int W1, int C, __half* out, cudaStream_t st);
int launch_sep_insert_nchw_f32(const float* x, const float* sep, int n_img, int C, int H, int W, float* out,
                               cudaStream_t st);
int launch_cinput_to_nhwc(const float* x, int n_img, int C, int H, int Wc, int x_off, int Wh, __half* out,
                          cudaStream_t st);
// first-stage decoder helpers (elementwise.cuh)
int launch_vae_in(const float* z, int n_img, int e, int zc, int H, int W, float z_scale, const float* pq_w,
                  const float* pq_b, int kpad, __half* out, cudaStream_t st);
int launch_softmax_rows(__half* x, int rows, int T, size_t ld, cudaStream_t st);
int launch_transpose_f16(const __half* in, int T, int C, size_t ld_in, __half* out, cudaStream_t st);
// nearest-x2 upsample folded into the following 3x3 conv: w [O, 9*I] (k = tap*I + i) -> 4 phase matrices [4][O, 4*I]
// (phase = py*2 + px, k = (a*2 + b)*I + i) with the taps that read the same source pixel summed in fp32
int launch_upfold_weights(const __half* w, int O, int I, __half* out, cudaStream_t st);
int launch_cast_f32_f16(const float* x, size_t n, __half* out, cudaStream_t st);
int launch_nhwc_to_nchw_f32(const __half* x, int ld, int n_img, int cout, int H, int W, float* out, cudaStream_t st);
int launch_nchw_f32_to_nhwc(const float* x, int n_img, int C, int H, int W, __half* out, cudaStream_t st);
int launch_small_linear(const float* in, int ld_in, int n_rows, int K, const __half* w, const float* bias, int n_out,
                        int silu_in, int silu_out, float* out, int ld_out, cudaStream_t st);
int launch_timestep_embedding(const long long* t, int t_count, int n, int dim, float* out, cudaStream_t st);
int launch_ddim_update(const float* x, const float* e_u, const float* e_c, const float* noise, float cfg, float a_t,
                       float a_prev, float sigma, float sqrt_one_minus_at, float temperature, size_t n, float* x_prev,
                       float* pred_x0, cudaStream_t st);
int launch_ddim_update_dev(const float* x, const float* e_u, const float* e_c, const float* noise, const float* coef,
                           float temperature, size_t n, float* x_prev, float* pred_x0, cudaStream_t st);
int launch_repack_conv(const float* w, int O, int I, int ldk, __half* out, cudaStream_t st);
int launch_repack_linear(const float* w, int O, int I, int geglu, int dst_row0, __half* out, cudaStream_t st);
int launch_repack_bias(const float* b, int O, int geglu, float* out, cudaStream_t st);

unsigned long long* debug_trace_buffer();
int debug_read_trace(void* dst, size_t bytes, int clear);

long long launches_since_reset();
void reset_launch_counter();
void add_launches(long long n);

}  // namespace lr
