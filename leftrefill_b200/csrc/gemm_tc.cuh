// tcgen05 GEMM / implicit-GEMM 3x3 convolution for sm_100a.
//
//   out[M, N] = epilogue( sum_{tap, c} A[pixel(m) + tap, c] * Wt[n, tap*Ctot + c] )
//
// Activations are NHWC fp16 (pixel-major rows, channels contiguous), so
//   * a Linear layer is the degenerate geometry  W = M, H = 1, images = 1, taps = 1;
//   * a 3x3 convolution (ResBlock in/out layers, Downsample, Upsample conv: reference
//     ldm/modules/diffusionmodules/openaimodel.py:90-118,133-159,200-231) is 9 taps, each a TMA box shifted by
//     (dy, dx) in the (x, y) coordinates of a 4-D tensor map [C, W, H, N]; TMA out-of-bounds zero fill IS the
//     conv padding, and the tensor map's element strides implement stride 2.
//   * the skip concat `th.cat([h, hs.pop()], 1)` (openaimodel.py:781) is never materialised: the K loop walks two
//     tensor maps (source 0 = h, source 1 = skip) back to back.
//   * Upsample (openaimodel.py:108-116, model.py:62-66) = nearest x2 followed by a 3x3 conv is never materialised
//     either: output pixel (2y + py, 2x + px) only sees the 2 x 2 source pixels (y + py - 1 + a, x + px - 1 + b), so
//     each of the four output phases (py, px) is a 4-tap conv on the LOW-resolution input with pre-summed weights
//     (upfold_weights_kernel) whose TMA store walks the output with a pixel stride of 2: 4/9 of the FLOPs and no 4x
//     tensor.
//   * HALO mode (stride-1 3x3 convs on images at least 16 rows high): the output tile is 8 pixels wide x 16 rows, and
//     for every 64-channel chunk ONE TMA box of (8+2) x (16+2) input pixels is loaded and re-used by all nine taps: the
//     A descriptor of tap (dy, dx) simply starts ((dy+1)*10 + (dx+1)) rows further into that box, with a stride of
//     10 rows between its 8-row groups (the 128B swizzle is a function of the shared-memory address bits, so a
//     row-shifted, row-padded operand reads correctly with base_offset 0 - tests/umma_halo_probe.cu). Activation
//     traffic from L2 drops 6.25x (23 KB instead of 9 x 16 KB per chunk); weights stream through their own ring, one
//     tile per (chunk, tap). The main loop of the non-halo path is bound by L2->SM bytes.
//
// Structure (one CTA per SM, persistent over output tiles of 128 x block_n):
//   warp 0 (1 lane)  TMA producer      -> smem ring of {A 128x64, B block_n x 64} fp16 stages, 128B swizzle
//   warp 1 (1 lane)  tcgen05.mma issue -> fp32 accumulator in TMEM, double buffered (2 x 256 columns)
// CG = 2 (large problems): the two CTAs of a cluster form a CTA pair. Each loads its own 128 pixel rows of A and HALF
// of the weight tile; the leader issues tcgen05.mma.cta_group::2 (M = 256), which reads B halves from both SMs, so
// every SM ingests (128 + block_n/2) x 128 B per k-chunk instead of (128 + block_n) x 128 B — the per-SM L2->smem
// ingest rate is what capped the 1-CTA kernel at ~55 % tensor-pipe utilisation.
//   warps 2..9       epilogue          -> tcgen05.ld, bias (smem staged) / per-image bias (time embedding) /
//                                         residual (prefetched one chunk ahead) / GEGLU -> fp16 tile staged in
//                                         swizzled smem -> TMA store (full 128 B lines; per-thread 16 B row-strided
//                                         global stores made the kernel epilogue-bound for K <= ~3000)
//
// GroupNorm fusion (round 2; reference ResBlock in_layers / out_layers = GroupNorm32 -> SiLU -> conv, openaimodel.py:200-204,
// 224-231,254-274, and SpatialTransformer norm -> proj_in, attention.py:399-404): a GroupNorm is an affine map per
// (image, channel), y = x * a[n, c] + b[n, c], once its statistics are known. Neither half needs its own pass:
//   * STATISTICS come from the PRODUCER: the epilogue that writes a tensor also sums, per output tile half (64 rows) and
//     channel, (sum x, sum x^2) of the fp16 values it staged for the TMA store and writes them to a fixed slot of a
//     partials table (GemmParams::stats_out). gn_finalize_kernel (elementwise.cuh) combines the slots of an image in a
//     fixed order (bit-reproducible, batch invariant) into a[n, c], b[n, c].
//   * APPLY (+ SiLU) can happen in the CONSUMER (template parameter XF): four extra "transform" warps apply the affine
//     map [and the SiLU] IN PLACE in shared memory on every activation tile the TMA producer lands, before the MMA warp
//     may read it (barrier xready). In halo mode one (8+2) x (16+2) box serves all nine taps, so every element is
//     transformed once per output tile - not once per tap - and the out-of-image halo is re-zeroed after the affine map
//     (the conv pads the NORMALISED tensor with zeros).
//     Measured on B200 (profiles/r2_ab_gn_fusion.txt): for the Linear consumers (SpatialTransformer norm -> proj_in, no
//     SiLU) this costs +13 us on a 40 us GEMM and replaces a 42 us GroupNorm pass: the engine uses it. For the 3x3
//     convs (GroupNorm + SiLU of 180 x 64 elements per k-chunk) four transform warps run at ~0.15 IPC and take longer
//     than the MMAs of the chunk (120 us vs 83 us at 320->320 @64x128; a variant in which these warps load the tile
//     from global memory themselves was slower still, 184 us), which is no better than the stand-alone pass; the
//     engine therefore applies GroupNorm + SiLU for convs with gn_apply_coef_kernel (ONE read + ONE write pass over
//     the producers' statistics) and keeps the in-kernel conv transform as a tested op-level path (LR_GN_FUSE_CONV=1).
#pragma once
#include "ptx.cuh"

namespace lr {

struct GemmParams {
  CUtensorMap tmA0;  // source 0 activations [C0, W, H, N] (conv) or [K, M, 1, 1] (linear)
  CUtensorMap tmA1;  // source 1 (skip connection) or a copy of tmA0
  CUtensorMap tmB;   // weights [Ktot, Ncols], K contiguous
  CUtensorMap tmC;   // output [n_valid, W, H, N], box (64 cols, bw, bh, bn), 128B swizzle   (TMA-store epilogue)
  CUtensorMap tmC2;  // same tensor, box (32 cols, ...), no swizzle: the 32-column remainder slab of a tile
  int tma_store;     // 1: epilogue stages the fp16 tile in smem and writes it with TMA (needs ld_out % 8 == 0)
  int cstage_off;    // byte offset of the staging buffer(s) inside dynamic smem
  int cstage_bufs;   // 1 or 2 staging buffers (2: short main loops, where the epilogue sets the tile period: it then never
                     // waits for a TMA store to drain its buffer)
  int cstage_bytes;  // bytes of one staging buffer (multiple of 1024)
  int halo;          // 1: halo mode (see the header comment); tile box is bw = 8, bh = 16, bn = 1
  int a_stages;      // halo: depth of the activation (halo tile) ring; `stages` is then the depth of the weight ring
  int a_slot_bytes;  // halo: bytes of one activation slot (multiple of 1024)
  int b_group;       // halo: weight tiles (taps) per weight-ring stage: 1, or 3 = one filter row (dx = -1, 0, 1) per barrier
                     // round trip (the producer <-> MMA handshake costs ~440 cycles, more than the MMAs of one
                     // 160-column tap)
  int ring_bytes;    // bytes of all pipeline rings = offset of the barrier block inside dynamic smem
  int res_tma;       // 1: the residual tile is TMA-loaded into the (double-buffered) staging buffer one tile ahead and
                     // the epilogue adds it in place. A thread-per-row LDG.128 of the residual touches 32 different
                     // 128-byte lines per instruction (rows are ld_res apart): ~3000 L1 wavefront cycles per tile,
                     // which is what made "+ residual" cost 17 us on the 65536 x 320 -> 320 linears.
  CUtensorMap tmR;   // residual [n_valid, W, H, N], box (64 cols, bw, bh, bn), 128B swizzle (same geometry as tmC)
  CUtensorMap tmR2;  // 32-column remainder slab, no swizzle (same geometry as tmC2)
  // LayerNorm folded into this GEMM (attention.py:279-283: every LayerNorm feeds a Linear). The weight carries gamma
  // (W' = W * gamma, fp16), `bias` carries b + W beta, and the epilogue finishes the normalisation per output row r:
  //   out[r, j] = rstd_r * acc[r, j] - rstd_r * mean_r * ln_s[j] + bias[j],   ln_s[j] = sum_k W'[j, k]
  // so the normalised activation tensor is never written or re-read; a stats-only pass produces (mean, rstd) per row.
  const float2* ln_stats;  // [M] (mean, rstd) or nullptr
  const float* ln_s;       // [ncols]
  // ... or the statistics come as per-row PARTIALS left by the epilogue of the GEMM that wrote the rows (rowstats_out
  // there): ln_part [M][ln_ld] float2 (sum, sum of squares), the first ln_slots entries of a row are summed in order,
  // mean = S / ln_c, rstd = rsqrt(Q / ln_c - mean^2 + ln_eps). No separate statistics pass over the activation.
  const float2* ln_part;
  int ln_slots, ln_ld, ln_c;
  float ln_eps;
  // producer side: every epilogue thread owns one output row and adds up (x, x^2) over the columns it converts
  // (fp32 values before the fp16 rounding); slot = 2 * (N-tile index) + (warp half) of row `grow`
  float2* rowstats_out;    // [M][rowstats_ld] or nullptr
  int rowstats_ld;
  int ksplit;        // split-K over the filter taps (1 = off, 3 = taps {0-2}, {3-5}, {6-8} as separate work units): for
                     // convs whose M is too small to fill the GPU. Units then write raw fp32 partial tiles to `partial`
                     // ([ksplit][M][ncols]) and splitk_reduce_kernel applies bias / residual and converts to fp16.
  float* partial;
  int n_img, H, W;   // OUTPUT pixel grid
  int bw, bh, bn;    // tile box: bw*bh*bn == 128 output pixels
  int tiles_x, tiles_y, tiles_b, tiles_n;
  int stride;        // 1 or 2 (input coordinate = stride * output coordinate + tap offset)
  int taps;          // 1, 9 (3x3) or 4 (one 2x2 phase of a nearest-x2 upsample folded into the following 3x3 conv)
  signed char tap_dy[9], tap_dx[9];  // input offset of every tap (plain mode)
  int kc0, kc1;      // 64-channel chunks per tap in source 0 / source 1
  int c0, ctot;      // channels of source 0, total channels (K extent of one tap in B)
  int ncols;         // GEMM N
  int block_n;       // UMMA N (multiple of 32, <= 256)
  int stages;        // smem ring depth
  const float* bias;       // [ncols] or nullptr
  const float* bias_img;   // [n_img, ld_bias_img] or nullptr (ResBlock emb_layers output, openaimodel.py:263-272)
  int ld_bias_img;
  const __half* residual;  // [M, ld_res] or nullptr, added after bias
  int ld_res;
  __half* out;             // [M, ld_out]
  int ld_out;
  int geglu;               // accumulator columns are (value, gate) pairs -> out[:, j] = v * gelu(g) (attention.py:51-58)
  int n_valid;             // valid output columns (after GEGLU halving)
  float out_scale;         // multiplies the final value (1.0 normally)
  int epi_lean;            // 1: the epilogue takes its LEAN path (host-checked: TMA-store staging, no split-K / folded
                           // LayerNorm / per-image bias / row statistics / out_scale, residual absent or TMA-loaded,
                           // ncols % 32 == 0): bias straight from global (L1 broadcast), 32-bit shared addresses, no
                           // per-row bookkeeping, columns split evenly between the two warps of a TMEM lane quarter
  // ---- GroupNorm fusion (see the header comment) ----
  const float* xf_scale;   // XF kernels: a[n_img][xf_ld] (rstd * gamma) of the (concatenated) input channels
  const float* xf_shift;   //             b[n_img][xf_ld] (beta - mean * rstd * gamma)
  int xf_ld;               // channels of the (concatenated) input
  int xf_silu;             // 1: SiLU after the affine map
  int xf_rows_per_img;     // linear geometry: token rows per image (the image index of row m is m / xf_rows_per_img)
  float2* stats_out;       // [tiles_m * 2][stats_ld] (sum, sum of squares) per tile half (64 rows) and output channel of the
                           // fp16 values this launch writes, or nullptr. Requires the TMA-store epilogue.
  int stats_ld;
  int dbg;                 // bring-up experiments only (LR_GEMM_DEBUG): 1 = no A loads, 2 = no B loads, 4 = no MMAs
  unsigned long long* trace;  // LR_GEMM_TRACE: CTA 0 records clock64() per role / tile / phase ([3][16][8]); else null
};

// role: 0 producer, 1 MMA issuer, 2 epilogue (warp 2). One lane of CTA 0 writes; first 16 tiles only.
#define LR_GEMM_TR(role, tcount, slot)                                                                       \
  do {                                                                                                       \
    if (p.trace != nullptr && blockIdx.x == 0 && (tcount) < 16 && lane == 0)                                 \
      p.trace[((role) * 16 + (tcount)) * 8 + (slot)] = static_cast<unsigned long long>(clock64());           \
  } while (0)

// Warp roles: role 0 = TMA producer, role 1 = MMA issuer, roles 2..9 = epilogue. -DLR_HI_WARP_ISSUE=1 moves the two
// single-lane issuing roles to the highest hardware warp ids (8, 9); measured on B200 (round 1, r1h): no difference for
// this kernel, so the natural order stays the default.
#ifndef LR_HI_WARP_ISSUE
#define LR_HI_WARP_ISSUE 0
#endif
// (Round 1 also tried a dedicated STORE warp issuing the TMA stores, profiles/r1_ab_store_warp.txt: no gain for the whole
// forward. Round 2 spreads the issue over the epilogue warps instead: see kStoreIssuers.)
constexpr int kGemmThreads = 320;
constexpr int kXfWarps = 4;                                 // transform warps of the XF instantiation (warps 10..13)
constexpr int kGemmThreadsXf = 320 + kXfWarps * 32;
constexpr int kEpiWarps = 8;
constexpr int kEpiThreads = kEpiWarps * 32;
// Issuing one bulk tensor copy costs the issuing thread 300-450 cycles (in-kernel trace, profiles/r2_trace_linears.txt),
// and an output tile is up to four slabs: lane 0 of epilogue warp i issues (and later confirms) the store of slab i,
// and the residual tile of a later unit is fetched by the TMA producer warp, not by an epilogue thread.
constexpr int kStoreIssuers = 4;
constexpr int kBlockM = 128;
constexpr int kBlockK = 64;
constexpr int kATileBytes = kBlockM * kBlockK * 2;  // 16 KB
constexpr int kMaxStages = 8;
constexpr int kMaxAStages = 4;
constexpr int kHaloBoxRows = (8 + 2) * (16 + 2);  // pixels of one halo box
constexpr int kHaloBW = 8, kHaloBH = 16;  // halo-mode output tile: 8 pixels wide x 16 rows
constexpr int kBarBytes = 512;
constexpr int kGemmAuxBytes = kBarBytes /*barriers*/ + 4 * 256 * 4 /*bias + LayerNorm column-sum staging, double buffered*/;
static_assert((5 * kMaxStages + 4 + 6) * 8 + 4 <= kBarBytes, "barrier block overflows");

// bytes of one pipeline stage in ONE CTA (cg = CTAs cooperating on a tile: each holds block_n / cg weight rows)
__host__ __device__ inline int gemm_stage_bytes(int block_n, int cg = 1) {
  return kATileBytes + (block_n / cg) * kBlockK * 2;
}

// erf with |abs error| <= 1.5e-7 (Abramowitz & Stegun 7.1.26): one MUFU.RCP + one MUFU.EX2 + 8 FMA-pipe ops.
// GELU here is the exact-erf form of F.gelu (attention.py:58); the output is rounded to fp16 (2^-11) afterwards.
__device__ __forceinline__ float fast_erf(float x) {
  const float ax = fabsf(x);
  float t;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.3275911f, ax, 1.0f)));
  float p = fmaf(1.061405429f, t, -1.453152027f);
  p = fmaf(p, t, 1.421413741f);
  p = fmaf(p, t, -0.284496736f);
  p = fmaf(p, t, 0.254829592f);
  p *= t;
  float e;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(-1.4426950408889634f * ax * ax));
  return copysignf(fmaf(-p, e, 1.0f), x);
}
__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + fast_erf(x * 0.70710678118654752f)); }
// The same GELU on a packed pair (FFMA2 / FMUL2: half the issue slots of the FMA-pipe part; the four MUFU operations and
// the sign handling stay scalar). z = x / sqrt(2); erf(|z|) = 1 - (a1 t + ... + a5 t^5) e^{-z^2}, t = 1 / (1 + p |z|).
__device__ __forceinline__ f32x2 gelu_erf2(f32x2 x) {
  float x0, x1;
  upk2(x, x0, x1);
  const f32x2 ax = pk2(fabsf(x0), fabsf(x1));
  float d0, d1;
  upk2(fma2(pk2(0.3275911f * 0.70710678118654752f, 0.3275911f * 0.70710678118654752f), ax, pk2(1.0f, 1.0f)), d0, d1);
  const f32x2 t = pk2(rcp_ftz(d0), rcp_ftz(d1));
  // -(a1 t + a2 t^2 + a3 t^3 + a4 t^4 + a5 t^5): the sign is folded into the coefficients
  f32x2 p = fma2(pk2(-1.061405429f, -1.061405429f), t, pk2(1.453152027f, 1.453152027f));
  p = fma2(p, t, pk2(-1.421413741f, -1.421413741f));
  p = fma2(p, t, pk2(0.284496736f, 0.284496736f));
  p = fma2(p, t, pk2(-0.254829592f, -0.254829592f));
  p = mul2(p, t);
  float a0, a1;  // -z^2 log2(e) = x^2 * (-0.5 log2(e))
  upk2(mul2(mul2(x, x), pk2(-0.72134752044448170f, -0.72134752044448170f)), a0, a1);
  float e0, e1;
  upk2(fma2(p, pk2(ex2_ftz(a0), ex2_ftz(a1)), pk2(1.0f, 1.0f)), e0, e1);  // erf(|z|)
  const f32x2 erf = pk2(copysignf(e0, x0), copysignf(e1, x1));
  const f32x2 h = mul2(x, pk2(0.5f, 0.5f));
  return fma2(h, erf, h);  // 0.5 x (1 + erf(z))
}

template <int V>
struct IntTag {
  static constexpr int value = V;
};

__device__ __forceinline__ float fast_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float fast_rcp(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// y = [silu](x * sc + sh) on 8 packed fp16 channels; the result is ANDed with `keep` (all ones / zero). Branch free on
// purpose: the transform warps have one warp per scheduler, their throughput comes from instruction-level parallelism
// across the 32 independent elements of four unrolled rows, which per-row branches (reconvergence points) destroy.
template <bool kSilu>
__device__ __forceinline__ uint4 xf_apply8(const uint4 v, const float (&sc)[8], const float (&sh)[8], const uint32_t keep) {
  const uint32_t w[4] = {v.x, v.y, v.z, v.w};
  uint32_t o[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 t = unpack_half2(w[i]);
    float y0 = fmaf(t.x, sc[2 * i], sh[2 * i]);
    float y1 = fmaf(t.y, sc[2 * i + 1], sh[2 * i + 1]);
    if (kSilu) {  // x * sigmoid(x) = x * rcp(1 + 2^(-x log2 e)): one MUFU.EX2 + one MUFU.RCP, flush-to-zero forms
      y0 *= fast_rcp(1.0f + fast_ex2(-1.4426950408889634f * y0));
      y1 *= fast_rcp(1.0f + fast_ex2(-1.4426950408889634f * y1));
    }
    o[i] = pack_half2(y0, y1) & keep;
  }
  return make_uint4(o[0], o[1], o[2], o[3]);
}

template <int CG, bool XF>
__global__ void __launch_bounds__(XF ? kGemmThreadsXf : kGemmThreads, 1) gemm_conv_kernel(const __grid_constant__ GemmParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // identical smem offsets in both CTAs of a pair are required (UMMA descriptors / multicast commits use offsets)
  uint8_t* smem = smem_raw;
  const int stage_bytes = gemm_stage_bytes(p.block_n, CG);
  const uint32_t rank = (CG == 2) ? cluster_ctarank() : 0u;
  const bool leader = rank == 0;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + p.ring_bytes);
  uint64_t* full = bars;                   // [kMaxStages]
  uint64_t* empty = bars + kMaxStages;     // [kMaxStages]
  uint64_t* tfull = bars + 2 * kMaxStages; // [2]
  uint64_t* tempty = tfull + 2;            // [2]
  uint64_t* a_full = tempty + 2;           // [kMaxStages]  activation tile landed (halo mode; XF: also the plain mode)
  uint64_t* a_empty = a_full + kMaxStages;   // [kMaxStages]
  uint64_t* xready = a_empty + kMaxStages;   // [kMaxStages]  XF: the transform warps (of both CTAs) are done with the tile
  uint64_t* cfree = xready + kMaxStages + 2;  // [2] the TMA stores of staging buffer b have finished reading it
  uint64_t* rfull = cfree + 2;               // [2] residual tile has landed in staging buffer b (res_tma)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(rfull + 2);
  float* sbias = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + kBarBytes);  // [2][256]
  float* slns = sbias + 512;                                                               // [2][256] ln_s slice

  const int hw_warp = threadIdx.x >> 5;
#if LR_HI_WARP_ISSUE
  const int warp = (hw_warp + 2) % 10;  // role: hardware warps 0..7 -> epilogue (2..9), 8 -> producer (0), 9 -> MMA (1)
#else
  const int warp = hw_warp;
#endif
  const int lane = threadIdx.x & 31;
  pdl_launch_dependents();

  if (threadIdx.x == 0) {
    if ((smem_u32(smem) & 1023u) != 0) {
      printf("lr_b200: gemm smem base not 1024-aligned\n");
      __trap();
    }
    tma_prefetch_desc(&p.tmA0);
    tma_prefetch_desc(&p.tmA1);
    tma_prefetch_desc(&p.tmB);
    for (int i = 0; i < p.stages; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull[i], 1);
      mbar_init(&tempty[i], kEpiWarps * CG);  // the leader's MMA thread waits for the epilogues of BOTH CTAs
    }
    for (int i = 0; i < kMaxStages; ++i) {
      mbar_init(&a_full[i], 1);
      mbar_init(&a_empty[i], 1);
      mbar_init(&xready[i], kXfWarps * CG);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&cfree[i], kEpiWarps);
      mbar_init(&rfull[i], 1);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    if (CG == 2) {
      tmem_alloc_2sm(tmem_slot, 512);
      tmem_relinquish_2sm();
    } else {
      tmem_alloc(tmem_slot, 512);
      tmem_relinquish();
    }
  }
  tc_fence_before();
  if (CG == 2) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();  // predecessors complete: from here on global memory may be read and written

  // Work unit = one 128*CG x block_n output tile per CTA (pair). CTA `rank` of a pair owns M-tile CG*pair + rank; when
  // the number of M-tiles is odd the last pair's second CTA recomputes the last tile and simply does not store it.
  const int tiles_m = p.tiles_x * p.tiles_y * p.tiles_b;
  const int tiles_mp = (tiles_m + CG - 1) / CG;
  const int num_tiles = tiles_mp * p.tiles_n;
  const int num_units = num_tiles * p.ksplit;  // unit u = (tile u / ksplit, tap range u % ksplit)
  const int unit0 = blockIdx.x / CG, unit_step = gridDim.x / CG;
  const int b_rows = p.block_n / CG;

  // res_tma: the TMA producer warp fetches the residual tile of its unit number `tcount` into staging buffer
  // (tcount & 1) right after the operands of that unit, i.e. about one tile period before the epilogue needs it. The
  // buffer was last read by the TMA stores (and the statistics pass) of unit tcount - 2; the epilogue warps confirm the
  // end of those reads at the start of unit tcount - 1 and arrive on cfree[].
  auto fetch_residual = [&](int tcount, int tn, int tx, int ty, int tb) {  // whole warp
    if (tcount >= 2) mbar_wait(&cfree[tcount & 1], ((tcount >> 1) - 1) & 1);
    if (elect_one()) {
      const int buf = tcount & 1;
      const int ocols_tile = p.block_n, full_slabs = p.block_n >> 6;  // (never combined with GEGLU)
      const int oc0 = tn * p.block_n;
      uint8_t* dst = smem + p.cstage_off + buf * p.cstage_bytes;
      int bytes = 0;
      for (int sl = 0; sl < full_slabs; ++sl)
        if (oc0 + sl * 64 < p.n_valid) bytes += kBlockM * 128;
      const bool rem = (ocols_tile & 63) != 0 && oc0 + full_slabs * 64 < p.n_valid;
      if (rem) bytes += kBlockM * 64;
      mbar_arrive_expect_tx(&rfull[buf], bytes);
      for (int sl = 0; sl < full_slabs; ++sl)
        if (oc0 + sl * 64 < p.n_valid)
          tma_load_4d(dst + sl * (kBlockM * 128), &p.tmR, &rfull[buf], oc0 + sl * 64, tx * p.bw, ty * p.bh, tb * p.bn);
      if (rem)
        tma_load_4d(dst + full_slabs * (kBlockM * 128), &p.tmR2, &rfull[buf], oc0 + full_slabs * 64, tx * p.bw, ty * p.bh,
                    tb * p.bn);
    }
    __syncwarp();
  };

  if (warp == 0 && p.halo) {
    // ------------------------------- TMA producer, halo mode ------------------------------------------------
    uint8_t* b_ring = smem + p.a_stages * p.a_slot_bytes;
    const int b_bytes = b_rows * kBlockK * 2;
    const int a_tx = (kHaloBW + 2) * (kHaloBH + 2) * kBlockK * 2;  // TMA credits the full box, zero-filled parts included
    auto produce = [&](auto g_tag) {
    constexpr int kG = decltype(g_tag)::value;  // weight tiles (taps) per weight-ring stage
    int sa = 0, sb = 0;
    uint32_t pha = 0, phb = 0;
    int tcount = 0;
    for (int tile = unit0; tile < num_tiles; tile += unit_step, ++tcount) {
      LR_GEMM_TR(0, tcount, 0);
      const int tn = tile % p.tiles_n;
      int tm = min((tile / p.tiles_n) * CG + static_cast<int>(rank), tiles_m - 1);
      const int tx = tm % p.tiles_x;
      tm /= p.tiles_x;
      const int ty = tm % p.tiles_y;
      const int tb = tm / p.tiles_y;
      const int x0 = tx * kHaloBW, y0 = ty * kHaloBH, n0 = tb;
      const int ncol0 = tn * p.block_n + static_cast<int>(rank) * b_rows;  // this CTA's share of the weight rows
      for (int kc = 0; kc < p.kc0 + p.kc1; ++kc) {
        const bool src0 = kc < p.kc0;
        const int kch = (src0 ? kc : kc - p.kc0) * kBlockK;
        const CUtensorMap* ta = src0 ? &p.tmA0 : &p.tmA1;
        mbar_wait(&a_empty[sa], pha ^ 1);
        if (elect_one()) {
          uint8_t* a_s = smem + sa * p.a_slot_bytes;
          if (CG == 2 && !XF) {
            if (leader) mbar_arrive_expect_tx(&a_full[sa], 2 * a_tx);
            tma_load_4d_2sm(a_s, ta, &a_full[sa], kch, x0 - 1, y0 - 1, n0);
          } else {  // XF: every CTA's transform warps wait for their OWN box, so it is credited to the local barrier
            mbar_arrive_expect_tx(&a_full[sa], a_tx);
            tma_load_4d(a_s, ta, &a_full[sa], kch, x0 - 1, y0 - 1, n0);
          }
        }
        __syncwarp();
        if (++sa == p.a_stages) { sa = 0; pha ^= 1; }
        for (int tap = 0; tap < 9; tap += kG) {
          mbar_wait(&empty[sb], phb ^ 1);
          if (elect_one()) {
            uint8_t* b_s = b_ring + sb * b_bytes * kG;
            if (CG == 2) {
              if (leader) mbar_arrive_expect_tx(&full[sb], 2 * b_bytes * kG);
            } else {
              mbar_arrive_expect_tx(&full[sb], b_bytes * kG);
            }
#pragma unroll
            for (int g = 0; g < kG; ++g) {
              const int kb = (tap + g) * p.ctot + (src0 ? 0 : p.c0) + kch;
              if (CG == 2) tma_load_2d_2sm(b_s + g * b_bytes, &p.tmB, &full[sb], kb, ncol0);
              else tma_load_2d(b_s + g * b_bytes, &p.tmB, &full[sb], kb, ncol0);
            }
          }
          __syncwarp();
          if (++sb == p.stages) { sb = 0; phb ^= 1; }
        }
      }
      if (p.res_tma) fetch_residual(tcount, tn, tx, ty, tb);
      LR_GEMM_TR(0, tcount, 1);
    }
    };
    if (p.b_group == 3) produce(IntTag<3>{}); else produce(IntTag<1>{});
  } else if (warp == 0) {
    // ------------------------------- TMA producer (whole warp loops, one elected lane issues) ---------------
    int s = 0;
    uint32_t ph = 0;
    int tcount = 0;
    for (int u = unit0; u < num_units; u += unit_step, ++tcount) {
      LR_GEMM_TR(0, tcount, 0);
      const int tile = u / p.ksplit, sp = u % p.ksplit;
      const int tap_b = sp * p.taps / p.ksplit, tap_e = (sp + 1) * p.taps / p.ksplit;
      const int tn = tile % p.tiles_n;
      int tm = min((tile / p.tiles_n) * CG + static_cast<int>(rank), tiles_m - 1);
      const int tx = tm % p.tiles_x;
      tm /= p.tiles_x;
      const int ty = tm % p.tiles_y;
      const int tb = tm / p.tiles_y;
      const int x0 = tx * p.bw * p.stride, y0 = ty * p.bh * p.stride, n0 = tb * p.bn;
      const int ncol0 = tn * p.block_n + static_cast<int>(rank) * b_rows;  // this CTA's share of the weight rows
      for (int tap = tap_b; tap < tap_e; ++tap) {
        const int dy = p.tap_dy[tap], dx = p.tap_dx[tap];
        for (int kc = 0; kc < p.kc0 + p.kc1; ++kc) {
          mbar_wait(&empty[s], ph ^ 1);
          if (elect_one()) {
            uint8_t* a_s = smem + s * stage_bytes;
            uint8_t* b_s = a_s + kATileBytes;
            const bool src0 = kc < p.kc0;
            const int kch = (src0 ? kc : kc - p.kc0) * kBlockK;
            const int kb = tap * p.ctot + (src0 ? 0 : p.c0) + kch;
            const CUtensorMap* ta = src0 ? &p.tmA0 : &p.tmA1;
            const int tx_bytes = ((p.dbg & 1) ? 0 : kATileBytes) + ((p.dbg & 2) ? 0 : stage_bytes - kATileBytes);
            if (XF) {
              // activation tile -> this CTA's a_full (its transform warps wait on it); weight tile -> full (leader's)
              mbar_arrive_expect_tx(&a_full[s], kATileBytes);
              tma_load_4d(a_s, ta, &a_full[s], kch, x0 + dx, y0 + dy, n0);
              if (CG == 2) {
                if (leader) mbar_arrive_expect_tx(&full[s], 2 * (stage_bytes - kATileBytes));
                tma_load_2d_2sm(b_s, &p.tmB, &full[s], kb, ncol0);
              } else {
                mbar_arrive_expect_tx(&full[s], stage_bytes - kATileBytes);
                tma_load_2d(b_s, &p.tmB, &full[s], kb, ncol0);
              }
            } else if (CG == 2) {
              // the leader's barrier collects the bytes of both CTAs
              if (leader) mbar_arrive_expect_tx(&full[s], 2 * tx_bytes);
              if (!(p.dbg & 1)) tma_load_4d_2sm(a_s, ta, &full[s], kch, x0 + dx, y0 + dy, n0);
              if (!(p.dbg & 2)) tma_load_2d_2sm(b_s, &p.tmB, &full[s], kb, ncol0);
            } else {
              mbar_arrive_expect_tx(&full[s], tx_bytes);
              if (!(p.dbg & 1)) tma_load_4d(a_s, ta, &full[s], kch, x0 + dx, y0 + dy, n0);
              if (!(p.dbg & 2)) tma_load_2d(b_s, &p.tmB, &full[s], kb, ncol0);
            }
          }
          __syncwarp();
          if (++s == p.stages) { s = 0; ph ^= 1; }
        }
      }
      if (p.res_tma) fetch_residual(tcount, tn, tx, ty, tb);
      LR_GEMM_TR(0, tcount, 1);
    }
  } else if (warp == 1 && p.halo) {
    // ------------------------------- MMA issuer, halo mode -------------------------------------------------
    if (leader) {
      const uint32_t idesc = umma_idesc_f16(kBlockM * CG, p.block_n, 0);
      const uint32_t desc_hi_a = umma_desc_hi_sw128((kHaloBW + 2) * 128);  // 8-row groups = image rows, 10 pixels apart
      const uint32_t desc_hi_b = umma_desc_hi_sw128(1024);
      const uint32_t a_lo0 = umma_desc_lo(smem_u32(smem), 16);
      const uint32_t b_lo0 = umma_desc_lo(smem_u32(smem) + p.a_stages * p.a_slot_bytes, 16);
      const uint32_t a_units = static_cast<uint32_t>(p.a_slot_bytes) >> 4;
      const uint32_t b_units = static_cast<uint32_t>(b_rows * kBlockK * 2) >> 4;
      auto issue = [&](auto g_tag) {
      constexpr int kG = decltype(g_tag)::value;
      int sa = 0, sb = 0;
      uint32_t pha = 0, phb = 0;
      int as = 0;
      uint32_t aph = 0;
      int tcount = 0;
      for (int tile = unit0; tile < num_tiles; tile += unit_step, ++tcount) {
        LR_GEMM_TR(1, tcount, 0);
        mbar_wait(&tempty[as], aph ^ 1);
        tc_fence_after();
        LR_GEMM_TR(1, tcount, 1);
        const uint32_t d_tmem = tmem_base + as * 256;
        uint32_t started = 0;
        for (int kc = 0; kc < p.kc0 + p.kc1; ++kc) {
          if (XF) { if (p.dbg & 128) mbar_wait_cluster(&xready[sa], pha); else mbar_wait(&xready[sa], pha); }
          else mbar_wait(&a_full[sa], pha);
          tc_fence_after();
          if (kc == 0) LR_GEMM_TR(1, tcount, 2);
          for (int tap0 = 0; tap0 < 9; tap0 += kG) {
            mbar_wait(&full[sb], phb);
            tc_fence_after();
            if (elect_one()) {
#pragma unroll
              for (int g = 0; g < kG; ++g) {
                const int tap = tap0 + g;
                // tap (dy, dx) = (tap / 3 - 1, tap % 3 - 1): start (dy + 1) * 10 + (dx + 1) rows (128 B = 8 units) into the box
                const uint32_t a_lo = a_lo0 + sa * a_units + static_cast<uint32_t>((tap / 3) * (kHaloBW + 2) + tap % 3) * 8u;
                const uint32_t b_lo = b_lo0 + (sb * kG + g) * b_units;
#pragma unroll
                for (int k = 0; k < kBlockK / 16; ++k) {
                  const uint64_t ad = umma_desc_make(desc_hi_a, a_lo + 2 * k);
                  const uint64_t bd = umma_desc_make(desc_hi_b, b_lo + 2 * k);
                  if (CG == 2) umma_f16_2sm(d_tmem, ad, bd, idesc, (started | g | k) != 0 ? 1u : 0u);
                  else umma_f16(d_tmem, ad, bd, idesc, (started | g | k) != 0 ? 1u : 0u);
                }
              }
              if (CG == 2) umma_commit_2sm(&empty[sb]); else umma_commit(&empty[sb]);
            }
            started = 1;
            __syncwarp();
            if (++sb == p.stages) { sb = 0; phb ^= 1; }
          }
          if (elect_one()) {
            if (CG == 2) umma_commit_2sm(&a_empty[sa]); else umma_commit(&a_empty[sa]);
          }
          __syncwarp();
          if (++sa == p.a_stages) { sa = 0; pha ^= 1; }
        }
        if (elect_one()) {
          if (CG == 2) umma_commit_2sm(&tfull[as]); else umma_commit(&tfull[as]);
        }
        __syncwarp();
        LR_GEMM_TR(1, tcount, 3);
        if (++as == 2) { as = 0; aph ^= 1; }
      }
      };
      if (p.b_group == 3) issue(IntTag<3>{}); else issue(IntTag<1>{});
    }
  } else if (warp == 1) {
    // ------------------------------- MMA issuer (whole warp loops, one elected lane issues) ----------------
    if (leader) {
      const uint32_t idesc = umma_idesc_f16(kBlockM * CG, p.block_n, 0);
      const uint32_t desc_hi = umma_desc_hi_sw128(1024);
      const uint32_t a_lo0 = umma_desc_lo(smem_u32(smem), 16);
      const uint32_t b_lo0 = umma_desc_lo(smem_u32(smem) + kATileBytes, 16);
      const uint32_t stage_units = static_cast<uint32_t>(stage_bytes) >> 4;
      int s = 0;
      uint32_t ph = 0;
      int as = 0;
      uint32_t aph = 0;
      int tcount = 0;
      for (int u = unit0; u < num_units; u += unit_step, ++tcount) {
        LR_GEMM_TR(1, tcount, 0);
        mbar_wait(&tempty[as], aph ^ 1);
        tc_fence_after();
        LR_GEMM_TR(1, tcount, 1);
        const uint32_t d_tmem = tmem_base + as * 256;
        const int sp = u % p.ksplit;
        const int kiters = ((sp + 1) * p.taps / p.ksplit - sp * p.taps / p.ksplit) * (p.kc0 + p.kc1);
        for (int it = 0; it < kiters; ++it) {
          mbar_wait(&full[s], ph);
          if (XF) { if (p.dbg & 128) mbar_wait_cluster(&xready[s], ph); else mbar_wait(&xready[s], ph); }
          tc_fence_after();
          if (it == 0) LR_GEMM_TR(1, tcount, 2);
          if (elect_one()) {
            const uint32_t a_lo = a_lo0 + s * stage_units;
            const uint32_t b_lo = b_lo0 + s * stage_units;
            if (!(p.dbg & 4)) {
#pragma unroll
              for (int k = 0; k < kBlockK / 16; ++k) {  // +32 bytes (= 2 descriptor units) per 16-element K step
                const uint64_t ad = umma_desc_make(desc_hi, a_lo + 2 * k);
                const uint64_t bd = umma_desc_make(desc_hi, b_lo + 2 * k);
                if (CG == 2) umma_f16_2sm(d_tmem, ad, bd, idesc, (it | k) != 0 ? 1u : 0u);
                else umma_f16(d_tmem, ad, bd, idesc, (it | k) != 0 ? 1u : 0u);
              }
            }
            if (CG == 2) umma_commit_2sm(&empty[s]); else umma_commit(&empty[s]);
          }
          __syncwarp();
          if (++s == p.stages) { s = 0; ph ^= 1; }
        }
        if (elect_one()) {
          if (CG == 2) umma_commit_2sm(&tfull[as]); else umma_commit(&tfull[as]);
        }
        __syncwarp();
        LR_GEMM_TR(1, tcount, 3);
        if (++as == 2) { as = 0; aph ^= 1; }
      }
    }
  } else if (XF && warp >= 2 + kEpiWarps) {
    // ------------------------------- transform warps (GroupNorm apply [+ SiLU] in shared memory) -------------
    // 128 threads; thread -> (physical 16-byte chunk q of a 128-byte row, rows prow0, prow0 + 16, ...). With the 128B
    // swizzle the LOGICAL channel chunk of physical chunk q in row r is q ^ (r & 7); rows advance by 16, so it is the
    // same for all rows of a thread: its 8 scale / shift values are loaded once per 64-channel chunk, before the tile
    // has landed.
    const int tt = static_cast<int>(threadIdx.x) - (2 + kEpiWarps) * 32;
    const int q = tt & 7, prow0 = tt >> 3;
    const int j8 = (q ^ (prow0 & 7)) * 8;
    const bool xf_skip = (p.dbg & 64) != 0;  // timing experiment: barrier hand-offs only, no transform work
    auto load_coef = [&](int img, int kc, float (&sc)[8], float (&sh)[8]) {
      const bool src0 = kc < p.kc0;
      const int kch = (src0 ? kc : kc - p.kc0) * kBlockK + j8;       // channel inside its source
      const int climit = src0 ? p.c0 : p.ctot - p.c0;
      if (kch + 8 <= climit) {  // channels beyond the source are TMA zero fill and meet zero weights: leave them zero
        const float* a = p.xf_scale + static_cast<size_t>(img) * p.xf_ld + (src0 ? 0 : p.c0) + kch;
        const float* b = p.xf_shift + static_cast<size_t>(img) * p.xf_ld + (src0 ? 0 : p.c0) + kch;
        const float4 a0 = __ldg(reinterpret_cast<const float4*>(a)), a1 = __ldg(reinterpret_cast<const float4*>(a) + 1);
        const float4 b0 = __ldg(reinterpret_cast<const float4*>(b)), b1 = __ldg(reinterpret_cast<const float4*>(b) + 1);
        sc[0] = a0.x; sc[1] = a0.y; sc[2] = a0.z; sc[3] = a0.w; sc[4] = a1.x; sc[5] = a1.y; sc[6] = a1.z; sc[7] = a1.w;
        sh[0] = b0.x; sh[1] = b0.y; sh[2] = b0.z; sh[3] = b0.w; sh[4] = b1.x; sh[5] = b1.y; sh[6] = b1.z; sh[7] = b1.w;
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) sc[i] = sh[i] = 0.f;
      }
    };
    auto signal = [&](uint64_t* bar) {
      fence_proxy_async_smem();  // generic-proxy writes -> visible to the tensor core's (async proxy) operand reads
      __syncwarp();
      if (lane == 0) {
        // the peer CTA's warps arrive on the leader's barrier (the leader issues the cta_group::2 MMA for both). Plain
        // arrives, like the accumulator-empty handshake; cluster-scope release / acquire on this barrier measured
        // +15 % on the 3x3 convs (LR_GEMM_DEBUG bit 128 selects them for A/B runs).
        if (CG == 2 && (p.dbg & 128)) {
          if (leader) mbar_arrive_release_cluster(bar); else mbar_arrive_leader_release_cluster(bar);
        } else if (CG == 2 && !leader) {
          mbar_arrive_leader(bar);
        } else {
          mbar_arrive(bar);
        }
      }
    };
    if (p.halo) {
      int sa = 0;
      uint32_t pha = 0;
      for (int tile = unit0; tile < num_tiles; tile += unit_step) {
        int tm = min((tile / p.tiles_n) * CG + static_cast<int>(rank), tiles_m - 1);
        const int tx = tm % p.tiles_x;
        tm /= p.tiles_x;
        const int ty = tm % p.tiles_y;
        const int n0 = tm / p.tiles_y;
        const int gx0 = tx * kHaloBW - 1, gy0 = ty * kHaloBH - 1;
        for (int kc = 0; kc < p.kc0 + p.kc1; ++kc) {
          float sc[8], sh[8];
          load_coef(n0, kc, sc, sh);
          mbar_wait(&a_full[sa], pha);
          uint8_t* slot = smem + sa * p.a_slot_bytes + q * 16;
          // 12 row slots per thread (rows >= 180 do not exist: the 12th slot of threads with prow0 >= 4), in three
          // straight-line groups of four. Pixels outside the image stay / become ZERO: the conv pads the NORMALISED
          // tensor (keep mask), so the out-of-bounds zero fill of the TMA box is re-established after the affine map.
          auto rows = [&](auto silu_tag) {
            constexpr bool kSilu = decltype(silu_tag)::value != 0;
#pragma unroll
            for (int g = 0; g < 3; ++g) {
              uint4 v[4];
              uint32_t keep[4];
              int off[4];
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                const int pr = prow0 + (g * 4 + k) * 16;
                const int prc = min(pr, kHaloBoxRows - 1);
                const int yl = prc / (kHaloBW + 2), xl = prc - yl * (kHaloBW + 2);
                const bool inb = static_cast<unsigned>(gx0 + xl) < static_cast<unsigned>(p.W) &&
                                 static_cast<unsigned>(gy0 + yl) < static_cast<unsigned>(p.H);
                keep[k] = inb ? 0xffffffffu : 0u;
                off[k] = pr < kHaloBoxRows ? prc * 128 : -1;
                v[k] = *reinterpret_cast<const uint4*>(slot + prc * 128);
              }
#pragma unroll
              for (int k = 0; k < 4; ++k) v[k] = xf_apply8<kSilu>(v[k], sc, sh, keep[k]);
#pragma unroll
              for (int k = 0; k < 4; ++k)
                if (off[k] >= 0) *reinterpret_cast<uint4*>(slot + off[k]) = v[k];
            }
          };
          if (!xf_skip) {
            if (p.xf_silu) rows(IntTag<1>{}); else rows(IntTag<0>{});
          }
          signal(&xready[sa]);
          if (++sa == p.a_stages) { sa = 0; pha ^= 1; }
        }
      }
    } else {
      // plain (Linear / 1x1) mode: one 128-row x 64-channel tile per stage. The host guarantees a token-matrix geometry
      // (H == 1, one "image", tiles of 128 consecutive rows): row r of M-tile tm is token row tm * 128 + r, its image
      // is row / xf_rows_per_img. Tiles inside one image (the normal case: rows per image a multiple of 128) take
      // their coefficients once per chunk.
      int s = 0;
      uint32_t ph = 0;
      const int M = p.W;
      for (int u = unit0; u < num_units; u += unit_step) {
        const int tile = u / p.ksplit;
        const int tm = min((tile / p.tiles_n) * CG + static_cast<int>(rank), tiles_m - 1);
        const int m0 = tm * kBlockM;
        const int img_first = m0 / p.xf_rows_per_img;
        const int img_last = min(m0 + kBlockM - 1, M - 1) / p.xf_rows_per_img;
        for (int kc = 0; kc < p.kc0 + p.kc1; ++kc) {  // taps == 1 (checked on the host)
          float sc[8], sh[8];
          load_coef(img_first, kc, sc, sh);
          mbar_wait(&a_full[s], ph);
          uint8_t* slot = smem + s * stage_bytes + q * 16;
          if (xf_skip) {
          } else if (img_first == img_last) {
#pragma unroll
            for (int it = 0; it < kBlockM / 16; ++it) {
              uint4* ptr = reinterpret_cast<uint4*>(slot + (prow0 + it * 16) * 128);
              *ptr = p.xf_silu ? xf_apply8<true>(*ptr, sc, sh, 0xffffffffu)   // rows beyond M are zero fill and are
                               : xf_apply8<false>(*ptr, sc, sh, 0xffffffffu);  // never stored
            }
          } else {
            int cur_img = img_first;
            for (int r = prow0; r < kBlockM; r += 16) {
              const int img = min(m0 + r, M - 1) / p.xf_rows_per_img;
              if (img != cur_img) {
                cur_img = img;
                load_coef(img, kc, sc, sh);
              }
              uint4* ptr = reinterpret_cast<uint4*>(slot + r * 128);
              *ptr = p.xf_silu ? xf_apply8<true>(*ptr, sc, sh, 0xffffffffu) : xf_apply8<false>(*ptr, sc, sh, 0xffffffffu);
            }
          }
          signal(&xready[s]);
          if (++s == p.stages) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp < 2 + kEpiWarps) {
    // ------------------------------- epilogue warps -----------------------------
    // 8 warps: warp (2+w) reads TMEM lane quarter (w & 3); the two warps of a quarter interleave 32-column chunks.
    // The XF instantiation (GroupNorm-consuming convs / proj_in) never runs GEGLU, a folded LayerNorm or split-K: pruning
    // those paths at compile time keeps it inside the 128 registers that 448 threads leave per thread.
    const bool k_geglu = !XF && p.geglu != 0;
    const bool k_ln = !XF && (p.ln_stats != nullptr || p.ln_part != nullptr);
    const bool k_partial = !XF && p.partial != nullptr;
    const int ew = warp - 2;
    const int q = hw_warp & 3;  // TMEM lane quarter this warp may access (hardware: warp id % 4)
    const int half = ew >> 2;
    const int r = q * 32 + lane;
    const int etid = ew * 32 + lane;
    int as = 0;
    uint32_t aph = 0;
    const bool vec_ok = (p.ld_out % 8 == 0) && (p.residual == nullptr || p.ld_res % 8 == 0);
    const int ocols_tile = k_geglu ? p.block_n / 2 : p.block_n;  // output columns of one tile
    const int full_slabs = ocols_tile >> 6;                      // 64-column slabs [128 rows][128 B], swizzled
    bool stores_pending = false;
    // bias: one column per epilogue thread (block_n <= 256 = kEpiThreads), fetched ONE TILE AHEAD into a register
    auto fetch_bias = [&](int tile) -> float {
      if (tile >= num_tiles || p.bias == nullptr || etid >= p.block_n || (p.dbg & 16)) return 0.f;  // tile = u / ksplit
      const int col = (tile % p.tiles_n) * p.block_n + etid;
      return col < p.ncols ? __ldg(p.bias + col) : 0.f;
    };
    auto fetch_lns = [&](int tile) -> float {
      if (!k_ln || tile >= num_tiles || etid >= p.block_n) return 0.f;
      const int col = (tile % p.tiles_n) * p.block_n + etid;
      return col < p.ncols ? __ldg(p.ln_s + col) : 0.f;
    };
    const bool lean = p.epi_lean != 0;
    float bias_next = lean ? 0.f : fetch_bias(unit0 / p.ksplit);
    float lns_next = lean ? 0.f : fetch_lns(unit0 / p.ksplit);
    // pixel of this thread's row inside a tile (the same for every tile)
    const int xi = r % p.bw;
    const int yi = (r / p.bw) % p.bh;
    const int ni = r / (p.bw * p.bh);
    // lean path: accumulator columns [lean_c0, lean_c1) of every tile belong to this warp (granularity: 16 columns, or
    // 32 = 16 outputs with GEGLU), so that block_n = 160 / 224 / 96 split 80 / 112 / 48 per warp instead of 96 + 64 ...
    const int lean_gran = k_geglu ? 32 : 16;
    const int lean_mid = (((p.block_n / lean_gran) + 1) >> 1) * lean_gran;
    const int lean_c0 = half ? lean_mid : 0, lean_c1 = half ? p.block_n : lean_mid;
    const uint32_t cstage_u32 = smem_u32(smem + p.cstage_off);
    const uint32_t r7x = static_cast<uint32_t>(r & 7) << 4;
    const size_t m_total = static_cast<size_t>(p.n_img) * p.H * p.W;
    // lane 0 of the first kStoreIssuers epilogue warps: issues the TMA store of slab `ew` of every tile (see kStoreIssuers)
    const bool issuer = lane == 0 && ew < kStoreIssuers;
    // the stores this thread issued have finished reading their staging buffer; with two buffers that frees the buffer
    // of the previous unit for the producer's next residual fetch
    // Every epilogue warp arrives (after ITS reads of the previous unit's staging buffer: residual, statistics pass), so
    // a completed cfree phase means nobody in this CTA touches that buffer any more.
    auto confirm_stores = [&](int tcount) {
      __syncwarp();
      if (lane == 0 && tcount > 0) {
        if (stores_pending) tma_store_wait_read();
        stores_pending = false;
        mbar_arrive(&cfree[(tcount - 1) & 1]);
      }
    };
    // unit -> (N-tile, M-tile pair) without a division per unit (ksplit == 1; split-K units divide below)
    int tn_i = unit0 % p.tiles_n, tmq_i = unit0 / p.tiles_n;
    const int step_q = unit_step / p.tiles_n, step_r = unit_step % p.tiles_n;
    const bool lin_geom = p.tiles_y == 1 && p.tiles_b == 1;
    int tcount = 0;
    for (int u = unit0; u < num_units; u += unit_step, ++tcount) {
      if (ew == 0) LR_GEMM_TR(2, tcount, 0);
      int sp = 0, tn, tmq;
      if (p.ksplit == 1) {
        tn = tn_i;
        tmq = tmq_i;
        tn_i += step_r;
        tmq_i += step_q;
        if (tn_i >= p.tiles_n) {
          tn_i -= p.tiles_n;
          ++tmq_i;
        }
      } else {
        const int tile = u / p.ksplit;
        sp = u % p.ksplit;
        tn = tile % p.tiles_n;
        tmq = tile / p.tiles_n;
      }
      uint8_t* cstage = smem + p.cstage_off + ((p.cstage_bufs == 2) ? (tcount & 1) * p.cstage_bytes : 0);
      int tm = tmq * CG + static_cast<int>(rank);
      const bool tile_ok = tm < tiles_m;
      tm = min(tm, tiles_m - 1);
      const int tm_idx = tm;  // linear M-tile index: the row pair of this tile in the statistics table
      int tx = tm, ty = 0, tb = 0;
      if (!lin_geom) {
        tx = tm % p.tiles_x;
        tm /= p.tiles_x;
        ty = tm % p.tiles_y;
        tb = tm / p.tiles_y;
      }
      const int ncol0 = tn * p.block_n;
      float rs_s = 0.f, rs_q = 0.f;  // rowstats_out: this thread's row, the columns this warp half converts
      bool row_ok = false;
      size_t grow = 0;

      if (lean) {
        // ---------------- lean path (see GemmParams::epi_lean) ----------------
        if (p.cstage_bufs == 1) {  // single staging buffer: the previous tile's TMA stores must have drained it
          confirm_stores(tcount);
          asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads) : "memory");
        }
        const uint32_t cst = cstage_u32 + ((p.cstage_bufs == 2) ? (tcount & 1) * p.cstage_bytes : 0);
        // Shared address of 16-byte chunk k of the output columns [oc, oc + 8 * nchunks) of this thread's staged row:
        // row_base(oc) ^ (k << 4). Full slabs are [128 rows][128 B] with the 128-byte swizzle (chunk index ^ (row & 7)),
        // the 32-column remainder slab is [128 rows][64 B] unswizzled; a group of 2 (4) chunks starts at an even
        // (multiple-of-4) chunk index, so adding k never carries and the whole address is one three-input XOR.
        auto row_base = [&](int oc) -> uint32_t {
          const int slab = oc >> 6;
          const uint32_t cpx = static_cast<uint32_t>((oc >> 3) & 7) << 4;
          return (slab < full_slabs) ? (cst + slab * (kBlockM * 128) + r * 128) ^ (cpx ^ r7x)
                                     : (cst + full_slabs * (kBlockM * 128) + r * 64) ^ cpx;
        };
        // folded LayerNorm (linear geometry: row index = 128 * M-tile + r): out = rstd * acc + (-mean * rstd) * ln_s + bias.
        // The row statistics do not depend on the MMA: fetched before the accumulator wait.
        float ln_rstd = 1.f, ln_nrm = 0.f;
        const int lrow = tm_idx * kBlockM + r;
        const bool lrow_ok = tile_ok && lrow < p.W;
        if (k_ln && lrow_ok) {
          float2 st;
          if (p.ln_part != nullptr) {
            float su = 0.f, sq = 0.f;
            const float2* pr = p.ln_part + static_cast<size_t>(lrow) * p.ln_ld;
            if (p.ln_slots <= 16 && (p.ln_slots & 1) == 0 && (p.ln_ld & 1) == 0) {
              // all partials of the row in flight at once (two per 16-byte load), then summed in slot order
              float4 v[8];
#pragma unroll
              for (int i = 0; i < 8; ++i)
                v[i] = (2 * i < p.ln_slots) ? __ldg(reinterpret_cast<const float4*>(pr) + i) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                if (2 * i < p.ln_slots) {
                  su += v[i].x;
                  sq += v[i].y;
                  su += v[i].z;
                  sq += v[i].w;
                }
              }
            } else {
              for (int i = 0; i < p.ln_slots; ++i) {  // fixed order: bit-reproducible
                const float2 t = __ldg(pr + i);
                su += t.x;
                sq += t.y;
              }
            }
            const float inv_c = 1.0f / static_cast<float>(p.ln_c);
            const float mean = su * inv_c;
            st = make_float2(mean, rsqrtf(fmaxf(fmaf(-mean, mean, sq * inv_c), 0.f) + p.ln_eps));
          } else {
            st = __ldg(p.ln_stats + lrow);
          }
          ln_rstd = st.y;
          ln_nrm = -st.x * st.y;
        }
        if (ew == 0) LR_GEMM_TR(2, tcount, 1);
        mbar_wait(&tfull[as], aph);
        tc_fence_after();
        if (p.cstage_bufs == 2) confirm_stores(tcount);
        if (p.res_tma) mbar_wait(&rfull[tcount & 1], (tcount >> 1) & 1);  // residual tile is in the staging buffer
        if (ew == 0) LR_GEMM_TR(2, tcount, 2);
        const uint32_t t_row = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + as * 256;
        f32x2 rs2_s = pk2(0.f, 0.f), rs2_q = pk2(0.f, 0.f);  // row statistics, even / odd columns
        auto run = [&](auto bias_tag, auto res_tag, auto ln_tag, auto rs_tag) {
          constexpr bool kBias = decltype(bias_tag)::value != 0;
          constexpr bool kRes = decltype(res_tag)::value != 0;
          constexpr bool kLn = decltype(ln_tag)::value != 0;  // (implies kBias)
          constexpr bool kRs = decltype(rs_tag)::value != 0;  // row statistics of the values written (rowstats_out)
          const float4* bias4 = reinterpret_cast<const float4*>(p.bias + ncol0);
          const float4* lns4 = reinterpret_cast<const float4*>(p.ln_s + ncol0);
          // accumulator value -> pre-activation: + bias, or the folded LayerNorm's per-row affine map
          auto pre = [&](uint32_t acc, float b, float sj) -> float {
            return kLn ? fmaf(ln_rstd, __uint_as_float(acc), fmaf(ln_nrm, sj, b)) : __uint_as_float(acc) + b;
          };
          if (k_geglu) {
            if constexpr (!kRes) {
              for (int c = lean_c0; c < lean_c1; c += 32) {
                if (ncol0 + c >= p.ncols) break;
                uint32_t v[32];
                tmem_ld32(t_row + c, v);
                float4 b4[8], s4[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                  b4[j] = kBias ? __ldg(bias4 + (c >> 2) + j) : make_float4(0.f, 0.f, 0.f, 0.f);
                  s4[j] = kLn ? __ldg(lns4 + (c >> 2) + j) : make_float4(0.f, 0.f, 0.f, 0.f);
                }
                const uint32_t a0 = row_base(c >> 1);
                tmem_ld_wait();
                // accumulator columns 4j .. 4j+3 = (value, value', gate, gate') of output pair j: packed fp32 throughout
                uint32_t o16[8];
                const f32x2 rstd2 = pk2(ln_rstd, ln_rstd), nrm2 = pk2(ln_nrm, ln_nrm);
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                  f32x2 val = pk2(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]));
                  f32x2 gate = pk2(__uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3]));
                  if constexpr (kLn) {
                    val = fma2(rstd2, val, fma2(nrm2, pk2(s4[j].x, s4[j].y), pk2(b4[j].x, b4[j].y)));
                    gate = fma2(rstd2, gate, fma2(nrm2, pk2(s4[j].z, s4[j].w), pk2(b4[j].z, b4[j].w)));
                  } else {
                    val = add2(val, pk2(b4[j].x, b4[j].y));
                    gate = add2(gate, pk2(b4[j].z, b4[j].w));
                  }
                  float r0, r1;
                  upk2(mul2(val, gelu_erf2(gate)), r0, r1);
                  o16[j] = pack_half2(r0, r1);
                }
                sts128(a0, make_uint4(o16[0], o16[1], o16[2], o16[3]));
                sts128(a0 ^ 16u, make_uint4(o16[4], o16[5], o16[6], o16[7]));
              }
            }
            return;
          }
          // kN accumulator columns (16 or 32) -> kN / 8 chunks of the staged row, the residual added in place
          auto chunk = [&](auto n_tag, int c) {
            constexpr int kN = decltype(n_tag)::value;
            uint32_t v[kN];
            if constexpr (kN == 32) tmem_ld32(t_row + c, v); else tmem_ld16(t_row + c, v);
            float4 b4[kN / 4], s4[kN / 4];
#pragma unroll
            for (int j = 0; j < kN / 4; ++j) {
              b4[j] = kBias ? __ldg(bias4 + (c >> 2) + j) : make_float4(0.f, 0.f, 0.f, 0.f);
              s4[j] = kLn ? __ldg(lns4 + (c >> 2) + j) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
            const uint32_t a0 = row_base(c);
            uint4 rr[kN / 8];
            (void)rr;
            if constexpr (kRes) {
#pragma unroll
              for (int k = 0; k < kN / 8; ++k) rr[k] = lds128(a0 ^ (k << 4));
            }
            tmem_ld_wait();
#pragma unroll
            for (int k = 0; k < kN / 8; ++k) {
              float f[8];
              f[0] = pre(v[8 * k + 0], b4[2 * k].x, s4[2 * k].x);
              f[1] = pre(v[8 * k + 1], b4[2 * k].y, s4[2 * k].y);
              f[2] = pre(v[8 * k + 2], b4[2 * k].z, s4[2 * k].z);
              f[3] = pre(v[8 * k + 3], b4[2 * k].w, s4[2 * k].w);
              f[4] = pre(v[8 * k + 4], b4[2 * k + 1].x, s4[2 * k + 1].x);
              f[5] = pre(v[8 * k + 5], b4[2 * k + 1].y, s4[2 * k + 1].y);
              f[6] = pre(v[8 * k + 6], b4[2 * k + 1].z, s4[2 * k + 1].z);
              f[7] = pre(v[8 * k + 7], b4[2 * k + 1].w, s4[2 * k + 1].w);
              if constexpr (kRes) {
                const uint32_t rw[4] = {rr[k].x, rr[k].y, rr[k].z, rr[k].w};
#pragma unroll
                for (int h = 0; h < 4; ++h) {
                  const float2 t = unpack_half2(rw[h]);
                  f[2 * h] += t.x;
                  f[2 * h + 1] += t.y;
                }
              }
              if constexpr (kRs) {
#pragma unroll
                for (int h = 0; h < 4; ++h) {
                  const f32x2 v2 = pk2(f[2 * h], f[2 * h + 1]);
                  rs2_s = add2(rs2_s, v2);
                  rs2_q = fma2(v2, v2, rs2_q);
                }
              }
              sts128(a0 ^ (k << 4),
                     make_uint4(pack_half2(f[0], f[1]), pack_half2(f[2], f[3]), pack_half2(f[4], f[5]), pack_half2(f[6], f[7])));
            }
          };
          int c = lean_c0;
          const int c_end = min(lean_c1, p.ncols - ncol0);  // ncols % 32 == 0: whole 16-column groups
          if ((c & 16) && c < c_end) {
            chunk(IntTag<16>{}, c);
            c += 16;
          }
          for (; c + 32 <= c_end; c += 32) chunk(IntTag<32>{}, c);
          if (c < c_end) chunk(IntTag<16>{}, c);
        };
        using T0 = IntTag<0>;
        using T1 = IntTag<1>;
        if (k_ln) {
          if (p.res_tma) run(T1{}, T1{}, T1{}, T0{}); else run(T1{}, T0{}, T1{}, T0{});
        } else if (p.rowstats_out != nullptr) {  // (host: only with a bias, never with GEGLU or a folded LayerNorm)
          if (p.res_tma) run(T1{}, T1{}, T0{}, T1{}); else run(T1{}, T0{}, T0{}, T1{});
          float s0, s1, q0, q1;
          upk2(rs2_s, s0, s1);
          upk2(rs2_q, q0, q1);
          rs_s = s0 + s1;
          rs_q = q0 + q1;
          row_ok = lrow_ok;
          grow = static_cast<size_t>(lrow);
        } else if (p.bias != nullptr) {
          if (p.res_tma) run(T1{}, T1{}, T0{}, T0{}); else run(T1{}, T0{}, T0{}, T0{});
        } else {
          if (p.res_tma) run(T0{}, T1{}, T0{}, T0{}); else run(T0{}, T0{}, T0{}, T0{});
        }
      } else {
      // ---------------- general path ----------------
      {
        const int x = tx * p.bw + xi, y = ty * p.bh + yi, n = tb * p.bn + ni;
        row_ok = tile_ok && (x < p.W) && (y < p.H) && (n < p.n_img);
        grow = (static_cast<size_t>(n) * p.H + y) * p.W + x;
      }
      const int n = tb * p.bn + ni;  // image of this thread's row (per-image bias)
      // single staging buffer: the previous tile's TMA stores must have drained it before anyone overwrites it
      if (p.cstage_bufs == 1) confirm_stores(tcount);
      // stage this tile's bias slice (double buffered by accumulator stage; the named barrier orders reuse)
      float* sb = sbias + as * 256;
      float* sl = slns + as * 256;
      if (etid < p.block_n) {
        sb[etid] = bias_next;
        sl[etid] = lns_next;
      }
      asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads) : "memory");
      bias_next = fetch_bias((u + unit_step) < num_units ? (u + unit_step) / p.ksplit : num_tiles);
      lns_next = fetch_lns((u + unit_step) < num_units ? (u + unit_step) / p.ksplit : num_tiles);
      // folded LayerNorm: this thread's row statistics (they do not depend on the MMA either)
      float ln_rstd = 1.f, ln_nrm = 0.f;
      if (k_ln && row_ok) {
        float2 st;
        if (p.ln_part != nullptr) {
          float su = 0.f, sq = 0.f;
          const float2* pr = p.ln_part + grow * p.ln_ld;
          for (int i = 0; i < p.ln_slots; ++i) {  // fixed order: bit-reproducible
            const float2 t = __ldg(pr + i);
            su += t.x;
            sq += t.y;
          }
          const float inv_c = 1.0f / static_cast<float>(p.ln_c);
          const float mean = su * inv_c;
          st = make_float2(mean, rsqrtf(fmaxf(fmaf(-mean, mean, sq * inv_c), 0.f) + p.ln_eps));
        } else {
          st = __ldg(p.ln_stats + grow);
        }
        ln_rstd = st.y;
        ln_nrm = -st.x * st.y;
      }

      // residual of the first chunk is fetched before the accumulator is ready (it does not depend on the MMA)
      const bool fast = vec_ok && row_ok && !k_geglu && !p.res_tma;
      const __half* res_row = (p.residual != nullptr) ? p.residual + grow * p.ld_res : nullptr;
      uint4 rnext[4] = {make_uint4(0, 0, 0, 0), make_uint4(0, 0, 0, 0), make_uint4(0, 0, 0, 0), make_uint4(0, 0, 0, 0)};
      int c = half * 32;
      if (fast && res_row != nullptr && c < p.block_n && ncol0 + c + 32 <= p.n_valid) {
        const uint4* rp = reinterpret_cast<const uint4*>(res_row + ncol0 + c);
#pragma unroll
        for (int k = 0; k < 4; ++k) rnext[k] = __ldg(rp + k);
      }

      if (ew == 0) LR_GEMM_TR(2, tcount, 1);
      mbar_wait(&tfull[as], aph);
      tc_fence_after();
      if (p.cstage_bufs == 2) confirm_stores(tcount);
      if (p.res_tma) mbar_wait(&rfull[tcount & 1], (tcount >> 1) & 1);  // residual tile is in the staging buffer
      if (ew == 0) LR_GEMM_TR(2, tcount, 2);
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + as * 256;
      for (; c < p.block_n; c += 64) {
        uint32_t v[32];
        if (p.dbg & 32) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = 0;
        } else {
          tmem_ld32(t_row + c, v);
        }
        uint4 rcur[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) rcur[k] = rnext[k];
        if (p.res_tma) {
          // the residual chunk sits exactly where this thread will write its output chunk
          const int slab = c >> 6, cp0 = (c & 63) >> 3;
          const uint8_t* rowp = (slab < full_slabs) ? cstage + slab * (kBlockM * 128) + r * 128
                                                    : cstage + full_slabs * (kBlockM * 128) + r * 64;
#pragma unroll
          for (int k = 0; k < 4; ++k)
            rcur[k] = (slab < full_slabs) ? *reinterpret_cast<const uint4*>(rowp + (((cp0 + k) ^ (r & 7)) << 4))
                                          : *reinterpret_cast<const uint4*>(rowp + ((cp0 + k) << 4));
        }
        const int cn = c + 64;
        if (fast && res_row != nullptr && cn < p.block_n && ncol0 + cn + 32 <= p.n_valid) {
          const uint4* rp = reinterpret_cast<const uint4*>(res_row + ncol0 + cn);
#pragma unroll
          for (int k = 0; k < 4; ++k) rnext[k] = __ldg(rp + k);
        }
        tmem_ld_wait();
        const int col0 = ncol0 + c;
        if (k_partial) {
          // split-K unit: raw fp32 partial sums; bias / residual / fp16 conversion happen in splitk_reduce_kernel
          if (row_ok && col0 < p.ncols) {
            float* po = p.partial + (static_cast<size_t>(sp) * m_total + grow) * p.ncols + col0;
            if (col0 + 32 <= p.ncols && (p.ncols & 3) == 0) {
#pragma unroll
              for (int j = 0; j < 8; ++j)
                reinterpret_cast<uint4*>(po)[j] = make_uint4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
            } else {
              for (int j = 0; j < 32; ++j)
                if (col0 + j < p.ncols) po[j] = __uint_as_float(v[j]);
            }
          }
          continue;
        }
        if (row_ok && col0 < p.ncols) {
          float f[32];
          if (k_ln) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              const float4 b4 = *reinterpret_cast<const float4*>(sb + c + j);
              const float4 s4 = *reinterpret_cast<const float4*>(sl + c + j);
              f[j] = fmaf(ln_rstd, __uint_as_float(v[j]), fmaf(ln_nrm, s4.x, b4.x));
              f[j + 1] = fmaf(ln_rstd, __uint_as_float(v[j + 1]), fmaf(ln_nrm, s4.y, b4.y));
              f[j + 2] = fmaf(ln_rstd, __uint_as_float(v[j + 2]), fmaf(ln_nrm, s4.z, b4.z));
              f[j + 3] = fmaf(ln_rstd, __uint_as_float(v[j + 3]), fmaf(ln_nrm, s4.w, b4.w));
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              const float4 b4 = *reinterpret_cast<const float4*>(sb + c + j);
              f[j] = __uint_as_float(v[j]) + b4.x;
              f[j + 1] = __uint_as_float(v[j + 1]) + b4.y;
              f[j + 2] = __uint_as_float(v[j + 2]) + b4.z;
              f[j + 3] = __uint_as_float(v[j + 3]) + b4.w;
            }
          }
          if (p.bias_img != nullptr) {
            const float* bi = p.bias_img + static_cast<size_t>(n) * p.ld_bias_img + col0;
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (col0 + j < p.ncols) f[j] += __ldg(bi + j);
          }
          if (k_geglu) {
            const int oc0 = col0 >> 1;
            __half* o = p.out + grow * p.ld_out + oc0;
            float g[16];
#pragma unroll
            for (int j = 0; j < 8; ++j) {  // columns (v, v', g, g') -> outputs v * gelu(g), v' * gelu(g')
              // the SAME packed arithmetic as the lean path: which path a launch takes depends on its geometry (hence on
              // the batch size for tiny token counts), the result must not (tests/test_multigpu.py)
              upk2(mul2(pk2(f[4 * j], f[4 * j + 1]), gelu_erf2(pk2(f[4 * j + 2], f[4 * j + 3]))), g[2 * j], g[2 * j + 1]);
            }  // (out_scale is never used with GEGLU)
            if (p.tma_store) {
              const uint4 w0 = make_uint4(pack_half2(g[0], g[1]), pack_half2(g[2], g[3]), pack_half2(g[4], g[5]),
                                          pack_half2(g[6], g[7]));
              const uint4 w1 = make_uint4(pack_half2(g[8], g[9]), pack_half2(g[10], g[11]), pack_half2(g[12], g[13]),
                                          pack_half2(g[14], g[15]));
              const int oc = c >> 1;  // output column inside the tile (multiple of 16)
              const int slab = oc >> 6, cp0 = (oc & 63) >> 3;
              if (slab < full_slabs) {
                uint8_t* rowp = cstage + slab * (kBlockM * 128) + r * 128;
                *reinterpret_cast<uint4*>(rowp + ((cp0 ^ (r & 7)) << 4)) = w0;
                *reinterpret_cast<uint4*>(rowp + (((cp0 + 1) ^ (r & 7)) << 4)) = w1;
              } else {
                uint8_t* rowp = cstage + full_slabs * (kBlockM * 128) + r * 64 + (cp0 << 4);
                *reinterpret_cast<uint4*>(rowp) = w0;
                *reinterpret_cast<uint4*>(rowp + 16) = w1;
              }
            } else if (vec_ok && oc0 + 16 <= p.n_valid) {
              uint4 w0 = make_uint4(pack_half2(g[0], g[1]), pack_half2(g[2], g[3]), pack_half2(g[4], g[5]),
                                    pack_half2(g[6], g[7]));
              uint4 w1 = make_uint4(pack_half2(g[8], g[9]), pack_half2(g[10], g[11]), pack_half2(g[12], g[13]),
                                    pack_half2(g[14], g[15]));
              reinterpret_cast<uint4*>(o)[0] = w0;
              reinterpret_cast<uint4*>(o)[1] = w1;
            } else {
              for (int j = 0; j < 16; ++j)
                if (oc0 + j < p.n_valid) o[j] = __float2half_rn(g[j]);
            }
          } else {
            __half* o = p.out + grow * p.ld_out + col0;
            if (p.tma_store || (vec_ok && col0 + 32 <= p.n_valid)) {
              if (res_row != nullptr) {
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                  const uint32_t rw[4] = {rcur[k].x, rcur[k].y, rcur[k].z, rcur[k].w};
#pragma unroll
                  for (int h = 0; h < 4; ++h) {
                    const float2 t = unpack_half2(rw[h]);
                    f[k * 8 + h * 2] += t.x;
                    f[k * 8 + h * 2 + 1] += t.y;
                  }
                }
              }
              if (p.rowstats_out != nullptr) {
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                  if (col0 + j < p.n_valid) {
                    rs_s += f[j];
                    rs_q = fmaf(f[j], f[j], rs_q);
                  }
                }
              }
              const int slab = c >> 6, cp0 = (c & 63) >> 3;
              uint8_t* rowp = (slab < full_slabs) ? cstage + slab * (kBlockM * 128) + r * 128
                                                  : cstage + full_slabs * (kBlockM * 128) + r * 64;
              if (p.out_scale != 1.0f) {  // only the VAE attention logits carry a scale: keep the multiply off the common path
#pragma unroll
                for (int j = 0; j < 32; ++j) f[j] *= p.out_scale;
              }
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                uint4 w = make_uint4(pack_half2(f[k * 8 + 0], f[k * 8 + 1]), pack_half2(f[k * 8 + 2], f[k * 8 + 3]),
                                     pack_half2(f[k * 8 + 4], f[k * 8 + 5]), pack_half2(f[k * 8 + 6], f[k * 8 + 7]));
                if (!p.tma_store) reinterpret_cast<uint4*>(o)[k] = w;
                else if (slab < full_slabs) *reinterpret_cast<uint4*>(rowp + (((cp0 + k) ^ (r & 7)) << 4)) = w;
                else *reinterpret_cast<uint4*>(rowp + ((cp0 + k) << 4)) = w;
              }
            } else {
              for (int j = 0; j < 32; ++j) {
                if (col0 + j < p.n_valid) {
                  float t = f[j];
                  if (res_row != nullptr) t += __half2float(res_row[col0 + j]);
                  rs_s += t;
                  rs_q = fmaf(t, t, rs_q);
                  o[j] = __float2half_rn(t * p.out_scale);
                }
              }
            }
          }
        }
      }
      }  // general path
      tc_fence_before();
      __syncwarp();
      if (ew == 0) LR_GEMM_TR(2, tcount, 3);
      if (lane == 0) {
        if (CG == 2) mbar_arrive_leader(&tempty[as]); else mbar_arrive(&tempty[as]);
      }
      if (p.rowstats_out != nullptr && row_ok)
        p.rowstats_out[grow * p.rowstats_ld + tn * 2 + half] = make_float2(rs_s, rs_q);
      if (p.tma_store) {
        // the fp16 tile is complete in smem: lane 0 of epilogue warp i writes slab i out with TMA (rows / columns outside
        // the tensor are clipped by the tensor map, so partial tiles need no masking). The stores that last read this
        // staging buffer were confirmed before anyone wrote to it (confirm_stores).
        fence_proxy_async_smem();
        asm volatile("bar.sync 2, %0;" ::"n"(kEpiThreads) : "memory");
        if (ew == 0) LR_GEMM_TR(2, tcount, 4);
        if (issuer && tile_ok && !(p.dbg & 8)) {
          const int oc_tile0 = k_geglu ? (ncol0 >> 1) : ncol0;
          const int x0 = tx * p.bw, y0 = ty * p.bh, n0 = tb * p.bn;
          const int oc = oc_tile0 + ew * 64;
          if (oc < p.n_valid && ew * 64 < ocols_tile) {
            if (ew < full_slabs) tma_store_4d(&p.tmC, cstage + ew * (kBlockM * 128), oc, x0, y0, n0);
            else tma_store_4d(&p.tmC2, cstage + full_slabs * (kBlockM * 128), oc, x0, y0, n0);
            tma_store_commit();
            stores_pending = true;
          }
        }
        if (ew == 0) LR_GEMM_TR(2, tcount, 6);
        if (p.stats_out != nullptr && tile_ok) {
          // GroupNorm statistics of the tile just staged (the fp16 values the TMA store is writing): thread -> (pair of
          // output columns, half of the rows). 32 consecutive threads read 128 contiguous bytes of one row: no bank
          // conflicts, the swizzle only permutes 16-byte chunks inside a row. The staging buffer is not rewritten before
          // every epilogue thread has passed the bias barrier of the next tile.
          const int cpair = etid & 127, rh = etid >> 7;
          const int col = 2 * cpair;  // output column inside the tile
          const int oc_tile0 = k_geglu ? (ncol0 >> 1) : ncol0;
          if (col < ocols_tile && oc_tile0 + col < p.n_valid) {
            const int slab = col >> 6, cin = col & 63;
            const bool in_slab = slab < full_slabs;
            const uint8_t* base = in_slab ? cstage + slab * (kBlockM * 128) + (cin & 7) * 2
                                          : cstage + full_slabs * (kBlockM * 128) + cin * 2;
            const int cp0 = cin >> 3;
            const int xlim = p.W - tx * p.bw, ylim = p.H - ty * p.bh, nlim = p.n_img - tb * p.bn;
            const bool full_tile = xlim >= p.bw && ylim >= p.bh && nlim >= p.bn;
            float s0 = 0.f, s1 = 0.f, q0 = 0.f, q1 = 0.f;
            auto acc = [&](uint32_t w) {
              const float2 t = unpack_half2(w);
              s0 += t.x;
              q0 = fmaf(t.x, t.x, q0);
              s1 += t.y;
              q1 = fmaf(t.y, t.y, q1);
            };
            const uint8_t* hbase = base + rh * 64 * (in_slab ? 128 : 64);
            if (full_tile) {
              // straight-line: 8 independent loads in flight; (row & 7) is the unrolled index k
#pragma unroll
              for (int i8 = 0; i8 < 8; ++i8) {
                uint32_t w[8];
#pragma unroll
                for (int k = 0; k < 8; ++k)
                  w[k] = in_slab ? *reinterpret_cast<const uint32_t*>(hbase + (i8 * 8 + k) * 128 + ((cp0 ^ k) << 4))
                                 : *reinterpret_cast<const uint32_t*>(hbase + (i8 * 8 + k) * 64);
#pragma unroll
                for (int k = 0; k < 8; ++k) acc(w[k]);
              }
            } else {
              for (int i = 0; i < 64; ++i) {  // tiles that stick out of the image: rows outside hold stale data
                const int rr = rh * 64 + i;
                const int xi2 = rr % p.bw, yi2 = (rr / p.bw) % p.bh, ni2 = rr / (p.bw * p.bh);
                if (xi2 >= xlim || yi2 >= ylim || ni2 >= nlim) continue;
                acc(in_slab ? *reinterpret_cast<const uint32_t*>(hbase + i * 128 + ((cp0 ^ (rr & 7)) << 4))
                            : *reinterpret_cast<const uint32_t*>(hbase + i * 64));
              }
            }
            float4* dst = reinterpret_cast<float4*>(p.stats_out + (static_cast<size_t>(tm_idx) * 2 + rh) * p.stats_ld +
                                                    oc_tile0 + col);
            *dst = make_float4(s0, q0, s1, q1);
          }
        }
        if (ew == 0) LR_GEMM_TR(2, tcount, 5);
      }
      if (++as == 2) { as = 0; aph ^= 1; }
    }
    if (issuer && stores_pending) tma_store_wait_read();
  }

  tc_fence_before();
  if (CG == 2) cluster_sync_all(); else __syncthreads();
  if (warp == 1) {
    if (CG == 2) tmem_dealloc_2sm(tmem_base, 512); else tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace lr
