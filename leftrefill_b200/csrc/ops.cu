// Op layer implementation: TMA descriptor construction, launch geometry heuristics, kernel launches.
#include "ops.h"

#include <atomic>
#include <cstdlib>
#include <cstring>
#include <mutex>

#include "attention_tc.cuh"
#include "elementwise.cuh"
#include "gemm_tc.cuh"

namespace lr {

static thread_local std::string g_err;
void set_error(const std::string& msg) { g_err = msg; }
const char* last_error() { return g_err.c_str(); }

static std::atomic<long long> g_launches{0};
long long launches_since_reset() { return g_launches.load(); }
void reset_launch_counter() { g_launches.store(0); }
void add_launches(long long n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
#define LR_LAUNCHED()                                        \
  do {                                                       \
    g_launches.fetch_add(1, std::memory_order_relaxed);      \
    LR_CUDA(cudaGetLastError());                             \
  } while (0)

static_assert(sizeof(GemmParams) <= sizeof(ConvOp::params), "GemmParams too large");
static_assert(sizeof(AttnParams) <= sizeof(AttnOp::params), "AttnParams too large");

// ------------------------------------------------------------------------------------------------------------
// cuTensorMapEncodeTiled through the runtime's driver entry point (no link-time libcuda dependency)
// ------------------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

// fp16 tensor map, 128B swizzle; dims innermost first; strides (bytes) for dims 1..rank-1.
static int make_tmap(CUtensorMap* m, const void* ptr, int rank, const uint64_t* dims, const uint64_t* strides,
                     const uint32_t* box, const uint32_t* estr, bool swizzle128 = true) {
  EncodeTiledFn fn = get_encode_fn();
  LR_CHECK(fn != nullptr, "cuTensorMapEncodeTiled entry point unavailable (no CUDA driver?)");
  LR_CHECK((reinterpret_cast<uintptr_t>(ptr) & 15) == 0, "TMA: global address must be 16-byte aligned");
  for (int i = 0; i < rank - 1; ++i) LR_CHECK(strides[i] % 16 == 0, "TMA: global strides must be multiples of 16 B");
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, static_cast<cuuint32_t>(rank), const_cast<void*>(ptr),
                  reinterpret_cast<const cuuint64_t*>(dims), reinterpret_cast<const cuuint64_t*>(strides),
                  reinterpret_cast<const cuuint32_t*>(box), reinterpret_cast<const cuuint32_t*>(estr),
                  CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with CUresult " + std::to_string(static_cast<int>(r)) + " (rank " +
              std::to_string(rank) + ", dims " + std::to_string(dims[0]) + "," + std::to_string(dims[1]) + ", box " +
              std::to_string(box[0]) + "," + std::to_string(box[1]) + ")");
    return 1;
  }
  return 0;
}

// All hot kernels go through this launcher: programmatic dependent launch (+ optional 2-CTA cluster).
template <typename... KArgs, typename... Args>
static cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, int cluster,
                              Args... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[3];
  int na = 0;
  static const bool use_pdl = getenv("LR_NO_PDL") == nullptr;
  // cluster < 0: COOPERATIVE launch (kernels with a grid-wide barrier inside): the driver only starts the grid when all of
  // its CTAs can be resident at once, also when other streams / processes share the GPU. Not combined with programmatic
  // dependent launch (the driver accepts both attributes together; measured: no difference, 18.1-18.3 ms either way).
  if (cluster < 0) {
    attr[na].id = cudaLaunchAttributeCooperative;
    attr[na].val.cooperative = 1;
    ++na;
  } else if (use_pdl) {
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  if (cluster > 1) {
    attr[na].id = cudaLaunchAttributeClusterDimension;
    attr[na].val.clusterDim.x = cluster;
    attr[na].val.clusterDim.y = 1;
    attr[na].val.clusterDim.z = 1;
    ++na;
  }
  cfg.attrs = attr;
  cfg.numAttrs = na;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// Debug trace buffer (LR_GEMM_TRACE / LR_ATTN_TRACE set): kernels of CTA 0 drop clock64() samples here; read back with
// lr_debug_read_trace. Null (no instrumentation executed) otherwise.
constexpr size_t kTraceBytes = 16384;
unsigned long long* debug_trace_buffer() {
  static unsigned long long* buf = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    if (getenv("LR_GEMM_TRACE") == nullptr && getenv("LR_ATTN_TRACE") == nullptr) return;
    if (cudaMalloc(&buf, kTraceBytes) != cudaSuccess) { buf = nullptr; return; }
    cudaMemset(buf, 0, kTraceBytes);
  });
  return buf;
}
int debug_read_trace(void* dst, size_t bytes, int clear) {
  unsigned long long* b = debug_trace_buffer();
  LR_CHECK(b != nullptr, "trace buffer not enabled (set LR_GEMM_TRACE or LR_ATTN_TRACE before the first op)");
  if (bytes > kTraceBytes) bytes = kTraceBytes;
  LR_CUDA(cudaDeviceSynchronize());
  LR_CUDA(cudaMemcpy(dst, b, bytes, cudaMemcpyDeviceToHost));
  if (clear) LR_CUDA(cudaMemset(b, 0, kTraceBytes));
  return 0;
}

// Per-device caches (a process may drive several GPUs: cudaFuncSetAttribute and the SM count are per device).
constexpr int kMaxDevices = 64;
static int current_device() {
  int dev = 0;
  cudaGetDevice(&dev);
  return (dev >= 0 && dev < kMaxDevices) ? dev : 0;
}
static int sm_count() {
  static std::atomic<int> n[kMaxDevices];
  const int dev = current_device();
  int v = n[dev].load(std::memory_order_relaxed);
  if (v == 0) {
    cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev);
    if (v <= 0) v = 148;
    n[dev].store(v, std::memory_order_relaxed);
  }
  return v;
}
// true exactly once per (device, slot): the caller then sets its per-device function attributes
static bool first_use_on_device(int slot) {
  static std::atomic<unsigned> done[kMaxDevices];
  const int dev = current_device();
  const unsigned bit = 1u << slot;
  return (done[dev].fetch_or(bit, std::memory_order_acq_rel) & bit) == 0;
}

// ------------------------------------------------------------------------------------------------------------
// GEMM / conv op
// ------------------------------------------------------------------------------------------------------------
static inline int cdiv(int a, int b) { return (a + b - 1) / b; }
static int env_int(const char* name, int dflt) {
  const char* e = getenv(name);
  return e ? atoi(e) : dflt;
}

namespace {
struct Tiling {
  int Ho, Wo, bw, bh, bn, tiles_x, tiles_y, tiles_b;
  bool halo;
};
Tiling conv_tiling(const ConvSpec& s) {
  Tiling t;
  t.Ho = (s.stride == 1) ? s.in_h : (s.in_h - 1) / 2 + 1;
  t.Wo = (s.stride == 1) ? s.in_w : (s.in_w - 1) / 2 + 1;
  // halo mode (gemm_tc.cuh): stride-1 3x3 convs on images at least one halo tile high
  static const bool halo_enabled = getenv("LR_NO_HALO") == nullptr;
  t.halo = halo_enabled && s.taps == 9 && s.stride == 1 && s.in_h >= kHaloBH && s.in_w >= kHaloBW;
  // tile box: power-of-two (bw, bh, bn) with product 128 minimising the number of (partially empty) tiles
  int best_tiles = INT32_MAX;
  t.bw = 128;
  t.bh = 1;
  t.bn = 1;
  if (t.halo) {
    t.bw = kHaloBW;
    t.bh = kHaloBH;
    t.bn = 1;
  }
  for (int w = 128; w >= 1 && !t.halo; w >>= 1) {
    for (int h = 128 / w; h >= 1; h >>= 1) {
      const int n = 128 / (w * h);
      const long long nt = 1LL * cdiv(t.Wo, w) * cdiv(t.Ho, h) * cdiv(s.n_img, n);
      if (nt < best_tiles) {
        best_tiles = static_cast<int>(nt);
        t.bw = w;
        t.bh = h;
        t.bn = n;
      }
    }
  }
  t.tiles_x = cdiv(t.Wo, t.bw);
  t.tiles_y = cdiv(t.Ho, t.bh);
  t.tiles_b = cdiv(s.n_img, t.bn);
  return t;
}
// every 128-row output tile lies inside one image, and the tiles of an image are consecutive
bool tiles_per_image(const ConvSpec& s, const Tiling& t, int* ppi) {
  if (s.stats_rows_per_img > 0) {  // token matrix [M, C]: in_h == 1, n_img == 1, tiles of 128 consecutive rows
    if (s.n_img != 1 || s.in_h != 1 || t.bw != 128 || s.stats_rows_per_img % 128 != 0) return false;
    *ppi = 2 * (s.stats_rows_per_img / 128);
    return true;
  }
  if (t.bn != 1) return false;
  *ppi = 2 * t.tiles_x * t.tiles_y;
  return true;
}
}  // namespace

size_t conv_stats_rows(const ConvSpec& s) {
  const Tiling t = conv_tiling(s);
  int ppi = 0;
  if (!tiles_per_image(s, t, &ppi)) return 0;
  return static_cast<size_t>(2) * t.tiles_x * t.tiles_y * t.tiles_b;
}
bool conv_is_halo(const ConvSpec& s) { return conv_tiling(s).halo; }

int build_conv_op(ConvOp* op, const ConvSpec& s) {
  LR_CHECK(s.a0 && s.w && s.out, "conv: null pointer");
  LR_CHECK(s.taps == 1 || s.taps == 9 || s.taps == 4, "conv: taps must be 1, 9 or 4 (folded upsample phase)");
  LR_CHECK(s.taps != 4 || (s.stride == 1 && s.out_sx > 0 && !s.geglu && s.residual == nullptr && s.xf_scale == nullptr),
           "conv: a folded-upsample phase needs stride 1, explicit output strides and a plain epilogue");
  LR_CHECK(s.stride == 1 || s.stride == 2, "conv: stride must be 1 or 2");
  LR_CHECK(s.c0 % 8 == 0 && s.c1 % 8 == 0 && s.lda0 % 8 == 0 && (s.a1 == nullptr || s.lda1 % 8 == 0),
           "conv: channel counts / leading dims must be multiples of 8 (use the im2col path otherwise)");
  LR_CHECK(s.ldw % 8 == 0, "conv: weight leading dim must be a multiple of 8");
  LR_CHECK(s.a1 != nullptr || s.c1 == 0, "conv: c1 without a1");
  LR_CHECK(s.a1 == nullptr || s.c0 % 64 == 0, "conv: two-source K split requires c0 % 64 == 0");
  GemmParams p;
  memset(&p, 0, sizeof(p));
  const Tiling tl = conv_tiling(s);
  const int Ho = tl.Ho, Wo = tl.Wo;
  op->out_h = Ho;
  op->out_w = Wo;
  const bool halo = tl.halo;
  const int bw = tl.bw, bh = tl.bh, bn = tl.bn;
  const bool xf = s.xf_scale != nullptr;
  LR_CHECK(!xf || (s.xf_shift != nullptr && (halo || s.taps == 1) && !s.geglu && s.ln_stats == nullptr),
           "conv: the fused GroupNorm transform needs a halo-mode 3x3 conv or a Linear, without GEGLU / folded LayerNorm");
  p.n_img = s.n_img;
  p.H = Ho;
  p.W = Wo;
  p.bw = bw;
  p.bh = bh;
  p.bn = bn;
  p.tiles_x = cdiv(Wo, bw);
  p.tiles_y = cdiv(Ho, bh);
  p.tiles_b = cdiv(s.n_img, bn);
  p.stride = s.stride;
  p.taps = s.taps;
  for (int t = 0; t < 9; ++t) {
    p.tap_dy[t] = static_cast<signed char>(s.taps == 9 ? t / 3 - 1 : (s.taps == 4 ? s.up_oy + t / 2 : 0));
    p.tap_dx[t] = static_cast<signed char>(s.taps == 9 ? t % 3 - 1 : (s.taps == 4 ? s.up_ox + t % 2 : 0));
  }
  p.kc0 = cdiv(s.c0, kBlockK);
  p.kc1 = cdiv(s.c1, kBlockK);
  p.c0 = s.c0;
  p.ctot = s.c0 + s.c1;
  p.ncols = s.ncols;
  const int tiles_m = p.tiles_x * p.tiles_y * p.tiles_b;
  // (cg, block_n): minimise waves * per-tile cost. The per-tile cost models the measured behaviour of the kernel:
  // tensor time ~ block_n, but the mainloop is bound by the bytes each SM pulls from L2 per k-chunk,
  // (128 + block_n / cg) rows, when that exceeds what the SM can ingest while the MMAs run.
  int block_n = s.force_block_n, cg = s.force_cg, ksplit = 1;
  {
    // split-K over the nine taps (3 work units per output tile + a reduction pass) for convs whose images are too
    // small for a halo tile (<= 256 output pixels per image: the 8x16 level): needs scratch for the fp32 partial tiles.
    // The decision depends on the per-image geometry ONLY, never on the batch size: split-K changes the summation
    // order, and results must not depend on how a batch is sharded over GPUs (tests/test_multigpu.py).
    const size_t m_out = static_cast<size_t>(s.n_img) * Ho * Wo;
    static const bool splitk_enabled = getenv("LR_NO_SPLITK") == nullptr;
    const bool can_split = splitk_enabled && s.taps == 9 && !halo && !s.geglu && Ho * Wo <= kSplitKMaxPixels &&
                           s.workspace != nullptr && s.workspace_bytes >= 3 * m_out * s.ncols * sizeof(float) &&
                           s.ncols % 8 == 0 && s.ld_out % 8 == 0 && (s.residual == nullptr || s.ld_res % 8 == 0);
    double best = 1e300;
    int best_bn = 0, best_cg = 0, best_ks = 1;
    for (int ks = (can_split ? 3 : 1); ks <= (can_split ? 3 : 1); ks += 2) {
      for (int c = 1; c <= 2; ++c) {
        if (s.force_cg && c != s.force_cg) continue;
        if (c == 2 && tiles_m < 2) continue;
        for (int bnn = 256; bnn >= 32; bnn -= 32) {
          if (s.force_block_n && bnn != s.force_block_n) continue;
          const long long units = 1LL * cdiv(tiles_m, c) * cdiv(s.ncols, bnn) * ks;
          const long long slots = sm_count() / c;
          const long long waves = (units + slots - 1) / slots;
          const double tensor = bnn;  // ~ cycles / 0.5 per k16 step
          // bytes-bound mainloop (calibrated on B200, ncu r1): rows of activations + weights each SM pulls per k-chunk;
          // in halo mode one 180-row box serves nine chunks
          const double ingest = (halo ? 20.0 : 128.0) + bnn / c;
          const double per_tile = (tensor > ingest ? tensor : ingest) / ks + 40.0 + (ks > 1 ? 25.0 : 0.0);
          const double cost = waves * per_tile;
          if (cost < best) {
            best = cost;
            best_bn = bnn;
            best_cg = c;
            best_ks = ks;
          }
        }
      }
    }
    LR_CHECK(best_bn != 0, "conv: no feasible tile configuration");
    block_n = best_bn;
    cg = best_cg;
    ksplit = best_ks;
  }
  LR_CHECK(block_n % 32 == 0 && block_n >= 32 && block_n <= 256, "conv: bad block_n");
  LR_CHECK(!s.geglu || (s.ncols % 4 == 0), "conv: GEGLU needs an even number of outputs (rows in groups of four)");
  p.block_n = block_n;
  p.tiles_n = cdiv(s.ncols, block_n);
  // TMA-store epilogue: needs 16-byte aligned output rows, output-tile widths made of 64-column slabs plus at most
  // one 32-column remainder, and (with a residual) whole 32-column chunks
  const int n_valid = s.geglu ? s.ncols / 2 : s.ncols;
  const int ocols_tile = s.geglu ? block_n / 2 : block_n;
  const bool tma_store = ksplit == 1 && (s.ld_out % 8 == 0) && ((ocols_tile % 64) == 0 || (ocols_tile % 64) == 32) &&
                         (s.residual == nullptr || (n_valid % 32 == 0 && s.ld_res % 8 == 0)) &&
                         getenv("LR_NO_TMA_STORE") == nullptr;
  const int cstage_bytes = tma_store ? kBlockM * ocols_tile * 2 : 0;  // multiple of 8 KB: keeps 1024 B alignment
  // Short main loops (small K): the epilogue sets the tile period, so it gets two staging buffers and never waits for
  // a TMA store to drain; long main loops keep the shared memory for pipeline stages.
  const int kiters = s.taps * (cdiv(s.c0, kBlockK) + cdiv(s.c1, kBlockK));
  static const int two_below = env_int("LR_GEMM_TWO_CSTAGE_KITERS", 24);
  int bufs = (tma_store && kiters <= two_below) ? 2 : 1;
  const int budget = 232448 - 2048 - kGemmAuxBytes;
  int stages, a_stages = 0, ring_bytes;
  const int a_slot = ((kHaloBW + 2) * (kHaloBH + 2) * kBlockK * 2 + 1023) & ~1023;
  if (halo) {
    // two rings: activation halo tiles (one per 64-channel chunk) and weight tiles (one per chunk and tap)
    static const int bgroup_max_bn = env_int("LR_HALO_BGROUP_MAX_BN", 192);  // one filter row per stage up to this block_n
    int b_group = (block_n <= bgroup_max_bn) ? 3 : 1;
    int b_bytes = (block_n / cg) * kBlockK * 2 * b_group;
    a_stages = 3;
    stages = (budget - a_stages * a_slot - bufs * cstage_bytes) / b_bytes;
    if (b_group == 3 && stages < 3) {  // not enough room for three grouped stages: one tap per stage
      b_group = 1;
      b_bytes = (block_n / cg) * kBlockK * 2;
      stages = (budget - a_stages * a_slot - bufs * cstage_bytes) / b_bytes;
    }
    if (stages < 4 && b_group == 1) {
      a_stages = 2;
      stages = (budget - a_stages * a_slot - bufs * cstage_bytes) / b_bytes;
    }
    if (stages > kMaxStages) stages = kMaxStages;
    LR_CHECK(stages >= 2, "conv (halo): not enough shared memory for 2 weight stages");
    p.b_group = b_group;
    ring_bytes = a_stages * a_slot + stages * b_bytes;
  } else {
    stages = (budget - bufs * cstage_bytes) / gemm_stage_bytes(block_n, cg);
    if (bufs == 2 && stages < 3) {
      bufs = 1;
      stages = (budget - cstage_bytes) / gemm_stage_bytes(block_n, cg);
    }
    if (stages > kMaxStages) stages = kMaxStages;
    LR_CHECK(stages >= 2, "conv: not enough shared memory for 2 stages");
    ring_bytes = stages * gemm_stage_bytes(block_n, cg);
  }
  p.stages = stages;
  p.halo = halo ? 1 : 0;
  p.a_stages = a_stages;
  p.a_slot_bytes = a_slot;
  p.ring_bytes = ring_bytes;
  p.tma_store = tma_store ? 1 : 0;
  p.cstage_off = (ring_bytes + kGemmAuxBytes + 1023) & ~1023;  // swizzle needs 1024 B
  p.cstage_bufs = bufs;
  p.cstage_bytes = cstage_bytes;
  LR_CHECK((s.ln_stats == nullptr && s.ln_part == nullptr) ||
               (s.ln_s != nullptr && s.bias != nullptr && s.taps == 1 && ksplit == 1),
           "conv: folded LayerNorm needs ln_s, a (folded) bias and a Linear geometry");
  LR_CHECK(s.ln_part == nullptr || (s.ln_slots > 0 && s.ln_ld >= s.ln_slots && s.ln_stats == nullptr && s.c1 == 0),
           "conv: folded LayerNorm from row partials needs ln_slots <= ln_ld and a single source");
  p.ln_stats = reinterpret_cast<const float2*>(s.ln_stats);
  p.ln_s = s.ln_s;
  p.ln_part = reinterpret_cast<const float2*>(s.ln_part);
  p.ln_slots = s.ln_slots;
  p.ln_ld = s.ln_ld;
  p.ln_c = s.c0;
  p.ln_eps = s.ln_eps;
  p.rowstats_out = nullptr;
  p.rowstats_ld = s.rowstats_ld;
  op->rowstats_slots = 0;
  if (s.rowstats_out != nullptr && ksplit == 1 && !s.geglu && s.taps == 1 && s.out_sx == 0 &&
      s.rowstats_ld >= 2 * cdiv(s.ncols, block_n)) {
    p.rowstats_out = reinterpret_cast<float2*>(s.rowstats_out);
    op->rowstats_slots = 2 * cdiv(s.ncols, block_n);
  }
  p.ksplit = ksplit;
  p.partial = ksplit > 1 ? s.workspace : nullptr;
  p.bias = ksplit > 1 ? nullptr : s.bias;  // split-K: bias / per-image bias / residual are applied by the reduction
  p.bias_img = ksplit > 1 ? nullptr : s.bias_img;
  p.ld_bias_img = s.ld_bias_img > 0 ? s.ld_bias_img : s.ncols;
  p.residual = ksplit > 1 ? nullptr : s.residual;
  p.ld_res = s.ld_res;
  p.out = s.out;
  p.ld_out = s.ld_out;
  p.geglu = s.geglu;
  p.n_valid = s.geglu ? s.ncols / 2 : s.ncols;
  p.out_scale = s.out_scale;
  {
    const char* e = getenv("LR_GEMM_DEBUG");
    p.dbg = e ? atoi(e) : 0;
  }
  p.trace = debug_trace_buffer();
  LR_CHECK(!(s.geglu && s.residual), "conv: GEGLU + residual not supported");
  LR_CHECK(!(s.geglu && s.out_scale != 1.0f), "conv: GEGLU + out_scale not supported");

  // activations: [C, W, H, N]
  {
    uint64_t dims[4] = {static_cast<uint64_t>(s.c0), static_cast<uint64_t>(s.in_w), static_cast<uint64_t>(s.in_h),
                        static_cast<uint64_t>(s.n_img)};
    uint64_t str[3] = {static_cast<uint64_t>(s.lda0) * 2, static_cast<uint64_t>(s.lda0) * 2 * s.in_w,
                       static_cast<uint64_t>(s.lda0) * 2 * s.in_w * s.in_h};
    uint32_t box[4] = {kBlockK, static_cast<uint32_t>(halo ? bw + 2 : bw * s.stride),
                       static_cast<uint32_t>(halo ? bh + 2 : bh * s.stride), static_cast<uint32_t>(bn)};
    uint32_t es[4] = {1, static_cast<uint32_t>(s.stride), static_cast<uint32_t>(s.stride), 1};
    LR_TRY(make_tmap(&p.tmA0, s.a0, 4, dims, str, box, es));
    if (s.a1 != nullptr) {
      dims[0] = static_cast<uint64_t>(s.c1);
      str[0] = static_cast<uint64_t>(s.lda1) * 2;
      str[1] = str[0] * s.in_w;
      str[2] = str[1] * s.in_h;
      LR_TRY(make_tmap(&p.tmA1, s.a1, 4, dims, str, box, es));
    } else {
      p.tmA1 = p.tmA0;
    }
  }
  {
    uint64_t dims[2] = {static_cast<uint64_t>(s.taps) * (s.c0 + s.c1), static_cast<uint64_t>(s.ncols)};
    LR_CHECK(dims[0] <= static_cast<uint64_t>(s.ldw), "conv: weight leading dim smaller than K");
    uint64_t str[1] = {static_cast<uint64_t>(s.ldw) * 2};
    uint32_t box[2] = {kBlockK, static_cast<uint32_t>(block_n / cg)};
    uint32_t es[2] = {1, 1};
    LR_TRY(make_tmap(&p.tmB, s.w, 2, dims, str, box, es));
  }
  p.res_tma = 0;
  static const bool res_tma_enabled = getenv("LR_NO_RES_TMA") == nullptr;
  if (res_tma_enabled && tma_store && s.residual != nullptr && bufs == 2 && ksplit == 1) {
    // residual tile loaded by TMA into the staging buffer, same tiling as the output
    uint64_t dims[4] = {static_cast<uint64_t>(n_valid), static_cast<uint64_t>(Wo), static_cast<uint64_t>(Ho),
                        static_cast<uint64_t>(s.n_img)};
    uint64_t str[3] = {static_cast<uint64_t>(s.ld_res) * 2, static_cast<uint64_t>(s.ld_res) * 2 * Wo,
                       static_cast<uint64_t>(s.ld_res) * 2 * Wo * Ho};
    uint32_t box[4] = {64, static_cast<uint32_t>(bw), static_cast<uint32_t>(bh), static_cast<uint32_t>(bn)};
    uint32_t es[4] = {1, 1, 1, 1};
    LR_TRY(make_tmap(&p.tmR, s.residual, 4, dims, str, box, es, true));
    box[0] = 32;
    LR_TRY(make_tmap(&p.tmR2, s.residual, 4, dims, str, box, es, false));
    p.res_tma = 1;
  }
  {
    // lean epilogue path (GemmParams::epi_lean): everything the UNet's Linears and residual convs need, nothing else
    const bool lean_enabled = getenv("LR_NO_LEAN_EPI") == nullptr;  // (read per op: tests switch it inside one process)
    // (a folded LayerNorm only in its usual form: per-row (mean, rstd) table, token-matrix geometry)
    const bool token_matrix = p.tiles_y == 1 && p.tiles_b == 1 && bw == kBlockM;
    const bool ln_ok = (p.ln_stats == nullptr && p.ln_part == nullptr) ||
                       (token_matrix && (reinterpret_cast<uintptr_t>(s.ln_s) & 15) == 0);
    // (per-row statistics of the output: token matrix, plain epilogue with a bias)
    const bool rs_ok = p.rowstats_out == nullptr ||
                       (token_matrix && s.bias != nullptr && !s.geglu && p.ln_stats == nullptr && p.ln_part == nullptr);
    p.epi_lean = (lean_enabled && tma_store && ksplit == 1 && ln_ok && rs_ok && p.bias_img == nullptr &&
                  s.out_scale == 1.0f && (s.residual == nullptr || p.res_tma) &&
                  s.ncols % 32 == 0 && p.dbg == 0 && (reinterpret_cast<uintptr_t>(s.bias) & 15) == 0)
                     ? 1
                     : 0;
  }
  if (tma_store) {
    // output tensor as the kernel tiles it: [cols, W, H, N] over the OUTPUT pixel grid
    uint64_t dims[4] = {static_cast<uint64_t>(n_valid), static_cast<uint64_t>(Wo), static_cast<uint64_t>(Ho),
                        static_cast<uint64_t>(s.n_img)};
    uint64_t str[3] = {static_cast<uint64_t>(s.ld_out) * 2, static_cast<uint64_t>(s.ld_out) * 2 * Wo,
                       static_cast<uint64_t>(s.ld_out) * 2 * Wo * Ho};
    if (s.out_sx > 0) {  // strided output (folded upsample phase: every other pixel of the high-resolution tensor)
      str[0] = s.out_sx * 2;
      str[1] = s.out_sy * 2;
      str[2] = s.out_sn * 2;
    }
    uint32_t box[4] = {64, static_cast<uint32_t>(bw), static_cast<uint32_t>(bh), static_cast<uint32_t>(bn)};
    uint32_t es[4] = {1, 1, 1, 1};
    LR_TRY(make_tmap(&p.tmC, s.out, 4, dims, str, box, es, true));
    box[0] = 32;
    LR_TRY(make_tmap(&p.tmC2, s.out, 4, dims, str, box, es, false));
  }
  // ---- GroupNorm fusion ----
  LR_CHECK(!xf || ksplit == 1, "conv: the fused GroupNorm transform cannot be combined with split-K");
  LR_CHECK(!xf || halo || (s.n_img == 1 && s.in_h == 1 && bw == 128 && s.xf_rows_per_img > 0),
           "conv: the fused GroupNorm transform of a Linear needs a token matrix (n_img = in_h = 1) and xf_rows_per_img");
  p.xf_scale = s.xf_scale;
  p.xf_shift = s.xf_shift;
  p.xf_ld = s.c0 + s.c1;
  p.xf_silu = s.xf_silu;
  p.xf_rows_per_img = s.xf_rows_per_img;

  LR_CHECK(!xf || (s.c0 + s.c1) % 8 == 0, "conv: fused GroupNorm needs channel counts that are multiples of 8");
  op->xf = xf ? 1 : 0;
  op->stats_ok = 0;
  op->stats_ppi = 0;
  p.stats_out = nullptr;
  p.stats_ld = s.ncols;
  {
    int ppi = 0;
    if (s.stats_out != nullptr && tma_store && ksplit == 1 && !s.geglu && s.ncols % 2 == 0 &&
        tiles_per_image(s, tl, &ppi)) {
      LR_CHECK((reinterpret_cast<uintptr_t>(s.stats_out) & 15) == 0, "conv: statistics table must be 16-byte aligned");
      p.stats_out = reinterpret_cast<float2*>(s.stats_out);
      op->stats_ok = 1;
      op->stats_ppi = ppi;
    }
  }
  LR_CHECK(s.out_sx == 0 || (tma_store && ksplit == 1),
           "conv: a strided output needs the TMA-store epilogue (cout a multiple of 32, 16-byte aligned rows)");
  const int num_units = cdiv(tiles_m, cg) * p.tiles_n * ksplit;
  op->ksplit = ksplit;
  op->ncols = s.ncols;
  op->rows_per_img = Ho * Wo;
  op->ld_bias_img = s.ld_bias_img > 0 ? s.ld_bias_img : s.ncols;
  op->ld_res = s.ld_res;
  op->ld_out = s.ld_out;
  op->m_out = static_cast<size_t>(s.n_img) * Ho * Wo;
  op->partial = s.workspace;
  op->bias = s.bias;
  op->bias_img = s.bias_img;
  op->residual = s.residual;
  op->out = s.out;
  const int slots = sm_count() / cg;
  op->grid = (num_units < slots ? num_units : slots) * cg;
  op->smem = tma_store ? p.cstage_off + bufs * cstage_bytes : 1024 + ring_bytes + kGemmAuxBytes;
  op->block_n = block_n;
  op->stages = stages;
  op->tiles = num_units;
  op->cg = cg;
  op->flops = 2.0 * s.n_img * Ho * Wo * static_cast<double>(s.ncols) * s.taps * (s.c0 + s.c1);
  memcpy(op->params, &p, sizeof(p));
  if (first_use_on_device(0)) {
    LR_CUDA(cudaFuncSetAttribute(gemm_conv_kernel<1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448));
    LR_CUDA(cudaFuncSetAttribute(gemm_conv_kernel<2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448));
    LR_CUDA(cudaFuncSetAttribute(gemm_conv_kernel<1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448));
    LR_CUDA(cudaFuncSetAttribute(gemm_conv_kernel<2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448));
  }
  return 0;
}

int launch_conv_op(const ConvOp& op, cudaStream_t st) {
  const GemmParams* p = reinterpret_cast<const GemmParams*>(op.params);
  if (op.xf) {
    if (op.cg == 2) {
      LR_CUDA(launch_pdl(gemm_conv_kernel<2, true>, dim3(op.grid), dim3(kGemmThreadsXf), op.smem, st, 2, *p));
    } else {
      LR_CUDA(launch_pdl(gemm_conv_kernel<1, true>, dim3(op.grid), dim3(kGemmThreadsXf), op.smem, st, 1, *p));
    }
  } else if (op.cg == 2) {
    LR_CUDA(launch_pdl(gemm_conv_kernel<2, false>, dim3(op.grid), dim3(kGemmThreads), op.smem, st, 2, *p));
  } else {
    LR_CUDA(launch_pdl(gemm_conv_kernel<1, false>, dim3(op.grid), dim3(kGemmThreads), op.smem, st, 1, *p));
  }
  LR_LAUNCHED();
  if (op.ksplit > 1) {
    const size_t total = op.m_out * (op.ncols / 8);
    LR_CUDA(launch_pdl(splitk_reduce_kernel, dim3(static_cast<unsigned>((total + 255) / 256)), dim3(256), 0, st, 1,
                       static_cast<const float*>(op.partial), op.ksplit, op.m_out, op.ncols, op.rows_per_img, op.bias,
                       op.bias_img, op.ld_bias_img, op.residual, op.ld_res, op.out, op.ld_out));
    LR_LAUNCHED();
  }
  return 0;
}

// Scratch for the op-level entry points (lr_conv3x3_f16): grown on demand, never shrunk, one per device (op-level calls
// on one device are expected to be issued from one stream at a time: the engine never uses this). The engine
// passes plan-owned scratch instead.
float* op_level_workspace(size_t bytes) {
  static float* ws_[kMaxDevices];
  static size_t ws_bytes_[kMaxDevices];
  static std::mutex mu;
  std::lock_guard<std::mutex> lock(mu);
  const int dev = current_device();
  float*& ws = ws_[dev];
  size_t& ws_bytes = ws_bytes_[dev];
  if (bytes > (size_t(512) << 20)) return nullptr;
  if (bytes > ws_bytes) {
    if (ws) {
      cudaDeviceSynchronize();  // an earlier op on another stream may still be using the old block
      cudaFree(ws);
    }
    ws = nullptr;
    ws_bytes = 0;
    if (cudaMalloc(&ws, bytes) != cudaSuccess) return nullptr;
    ws_bytes = bytes;
  }
  return ws;
}

// ------------------------------------------------------------------------------------------------------------
// attention op
// ------------------------------------------------------------------------------------------------------------
int build_attn_op(AttnOp* op, const AttnSpec& s) {
  LR_CHECK(s.q && s.k && s.v && s.out, "attention: null pointer");
  LR_CHECK(s.ldq % 8 == 0 && s.ldk % 8 == 0 && s.ldv % 8 == 0 && s.ld_out % 8 == 0,
           "attention: leading dims must be multiples of 8");
  LR_CHECK(s.tq > 0 && s.tk > 0 && s.batch > 0 && s.heads > 0, "attention: empty problem");
  AttnParams p;
  memset(&p, 0, sizeof(p));
  auto mk = [&](CUtensorMap* m, const __half* ptr, int ld, int T) -> int {
    uint64_t dims[3] = {static_cast<uint64_t>(ld), static_cast<uint64_t>(T), static_cast<uint64_t>(s.batch)};
    uint64_t str[2] = {static_cast<uint64_t>(ld) * 2, static_cast<uint64_t>(ld) * 2 * T};
    uint32_t box[3] = {kAttnD, kAttnTile, 1};
    uint32_t es[3] = {1, 1, 1};
    return make_tmap(m, ptr, 3, dims, str, box, es);
  };
  LR_TRY(mk(&p.tmQ, s.q, s.ldq, s.tq));
  LR_TRY(mk(&p.tmK, s.k, s.ldk, s.tk));
  LR_TRY(mk(&p.tmV, s.v, s.ldv, s.tk));
  p.heads = s.heads;
  p.tq = s.tq;
  p.tk = s.tk;
  p.batch = s.batch;
  p.q_col0 = s.q_col0;
  p.k_col0 = s.k_col0;
  p.v_col0 = s.v_col0;
  p.out = s.out;
  p.ld_out = s.ld_out;
  p.scale_log2 = s.scale * 1.4426950408889634f;
  p.trace = getenv("LR_ATTN_TRACE") ? debug_trace_buffer() : nullptr;
  op->grid = dim3(cdiv(s.tq, kAttnQBlock), s.heads, s.batch);
  // persistent scheduling (attention_persist_kernel), LR_ATTN_PERSIST=1: for key sequences of up to
  // LR_ATTN_PERSIST_MAX_TILES KV steps (default 2: the cross-attentions) whenever every CTA gets whole 256-query blocks.
  // Measured on B200 (profiles/r2_ab_attention_persistent.txt): the cross-attention launches get 17 % faster when timed
  // alone, the CUDA-graph replay of the whole forward does not change (18.31 vs 18.32 ms), so it stays an option.
  // (read per op: ops are built at plan time, and tests switch it inside one process)
  const bool persist_enabled = env_int("LR_ATTN_PERSIST", 0) != 0;
  const int persist_max_tiles = env_int("LR_ATTN_PERSIST_MAX_TILES", 2);
  op->persist = 0;
  if (persist_enabled && s.tq % kAttnQBlock == 0 && cdiv(s.tk, kAttnTile) <= persist_max_tiles) {
    const long long items = static_cast<long long>(s.tq / kAttnQBlock) * s.heads * s.batch;
    op->persist = 1;
    op->grid = dim3(static_cast<unsigned>(items < sm_count() ? items : sm_count()));
  }
  op->flops = 4.0 * s.batch * s.heads * static_cast<double>(s.tq) * s.tk * kAttnD;
  memcpy(op->params, &p, sizeof(p));
  if (first_use_on_device(1)) {
    LR_CUDA(cudaFuncSetAttribute(attention_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kAttnSmemBytes));
    LR_CUDA(cudaFuncSetAttribute(attention_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kAttnSmemBytes));
    LR_CUDA(cudaFuncSetAttribute(attention_persist_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 kAttnPersistSmemBytes));
    LR_CUDA(cudaFuncSetAttribute(attention_persist_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 kAttnPersistSmemBytes));
  }
  return 0;
}

int launch_attn_op(const AttnOp& op, cudaStream_t st) {
  const AttnParams* p = reinterpret_cast<const AttnParams*>(op.params);
  if (op.persist) {
    if (p->trace != nullptr) {
      LR_CUDA(launch_pdl(attention_persist_kernel<true>, op.grid, dim3(kAttnThreads), kAttnPersistSmemBytes, st, 1, *p));
    } else {
      LR_CUDA(launch_pdl(attention_persist_kernel<false>, op.grid, dim3(kAttnThreads), kAttnPersistSmemBytes, st, 1, *p));
    }
  } else if (p->trace != nullptr) {
    LR_CUDA(launch_pdl(attention_kernel<true>, op.grid, dim3(kAttnThreads), kAttnSmemBytes, st, 1, *p));
  } else {
    LR_CUDA(launch_pdl(attention_kernel<false>, op.grid, dim3(kAttnThreads), kAttnSmemBytes, st, 1, *p));
  }
  LR_LAUNCHED();
  return 0;
}

// ------------------------------------------------------------------------------------------------------------
// normalisation / elementwise
// ------------------------------------------------------------------------------------------------------------
size_t groupnorm_scratch_bytes(int n_img, int groups, int P) {
  return (gn_scratch_bytes(n_img, groups, P) + 255) & ~size_t(255);
}

int launch_groupnorm(const __half* x0, int c0, const __half* x1, int c1, int n_img, int P, int groups, float eps,
                     const float* gamma, const float* beta, int do_silu, void* scratch, int scratch_is_zero,
                     __half* out, cudaStream_t st) {
  const int C = c0 + c1;
  LR_CHECK(c0 % 8 == 0 && c1 % 8 == 0, "groupnorm: channels must be multiples of 8");
  LR_CHECK(C % groups == 0, "groupnorm: channels not divisible by groups");
  LR_CHECK(C / 8 <= kNormThreads, "groupnorm: too many channels");
  LR_CHECK(groups <= kNormThreads, "groupnorm: too many groups");
  LR_CHECK(x1 != nullptr || c1 == 0, "groupnorm: c1 without x1");
  const int rpi = kNormThreads / (C / 8);
  const size_t smem = static_cast<size_t>(rpi) * 2 * C * sizeof(float);
  // Images of up to 8 x LR_GN_FUSED_KB KB take the single-launch cluster kernel (one 8-CTA cluster per image, second pass
  // hits L2), larger ones the persistent kernel. Measured on B200 at N = 8 (round 1, r1m): the cluster kernel wins up to
  // 1.3 MB per image (16x32x1280, 8x16x1280); 16-CTA (non-portable) clusters were slower everywhere. No choice depends
  // on the batch size (bit-exact batch invariance).
  static const int fused_kb = env_int("LR_GN_FUSED_KB", 160);
  constexpr int kCS = 8;
  const size_t img_bytes = static_cast<size_t>(P) * C * sizeof(__half);
  // the persistent kernel shares the chunk sums of a group between 512 / groups lanes of one warp
  const bool persistent_ok = groups >= 16 && groups <= 64 && (groups & (groups - 1)) == 0;
  if (groups <= 64 && ((fused_kb > 0 && img_bytes <= static_cast<size_t>(kCS) * fused_kb * 1024) || !persistent_ok)) {
    LR_CUDA(launch_pdl(gn_fused_cluster_kernel, dim3(kCS, n_img), dim3(kNormThreads), smem, st, kCS, x0, c0, x1, c1, P,
                       groups, eps, gamma, beta, do_silu, out));
    LR_LAUNCHED();
    return 0;
  }
  LR_CHECK(persistent_ok, "groupnorm: unsupported group count");
  LR_CHECK(scratch != nullptr, "groupnorm: scratch buffer required");
  LR_CHECK((reinterpret_cast<uintptr_t>(scratch) & 15) == 0, "groupnorm: scratch must be 16-byte aligned");
  const long long items = static_cast<long long>(n_img) * gn_chunks(P);
  // all CTAs must be resident (grid-wide barrier inside): two per SM (__launch_bounds__(512, 2): 64 registers, <= 33 KB
  // of shared memory each)
  const long long slots = 2LL * sm_count();
  const int grid = static_cast<int>(items < slots ? items : slots);
  if (!scratch_is_zero) LR_CUDA(cudaMemsetAsync(scratch, 0, 16, st));
  static const bool gn_coop = getenv("LR_GN_NO_COOP") == nullptr;  // cooperative launch of the grid-barrier kernel
  // bring-up (tests/gpu_time_gn_passes.py, gpu_gn_trace.py): 1 = skip pass 1, 2 = skip pass 2, 16 = print phase cycles
  static const int gn_dbg = env_int("LR_GN_DEBUG", 0);
  LR_CUDA(launch_pdl(gn_persistent_kernel, dim3(grid), dim3(kNormThreads), smem, st, gn_coop ? -1 : 1, x0, c0, x1, c1, P, n_img, groups,
                     eps, gamma, beta, do_silu, out, static_cast<unsigned char*>(scratch), gn_dbg));
  LR_LAUNCHED();
  return 0;
}

int launch_gn_finalize(const float* part0, int ppi0, int c0, const float* part1, int ppi1, int c1, int n_img, int P,
                       int groups, float eps, const float* gamma, const float* beta, float* scale, float* shift,
                       cudaStream_t st) {
  LR_CHECK(part0 != nullptr && (part1 != nullptr || c1 == 0), "gn_finalize: null statistics table");
  LR_CHECK((c0 + c1) % groups == 0, "gn_finalize: channels not divisible by groups");
  LR_CUDA(launch_pdl(gn_finalize_kernel, dim3(groups, n_img), dim3(kGnFinalizeThreads), 0, st, 1,
                     reinterpret_cast<const float2*>(part0), ppi0, c0, reinterpret_cast<const float2*>(part1), ppi1, c1,
                     static_cast<double>(P) * ((c0 + c1) / groups), eps, gamma, beta, scale, shift));
  LR_LAUNCHED();
  return 0;
}

int launch_gn_apply_coef(const __half* x0, int c0, const __half* x1, int c1, int n_img, int P, const float* scale,
                         const float* shift, int do_silu, __half* out, cudaStream_t st) {
  const int C = c0 + c1;
  LR_CHECK(c0 % 8 == 0 && c1 % 8 == 0 && C / 8 <= kNormThreads, "gn_apply_coef: bad channel counts");
  static const int chunk_div = env_int("LR_GN_APPLY_CHUNK_DIV", 32);
  int chunk = P / chunk_div;
  if (chunk < 16) chunk = 16;
  if (chunk > 512) chunk = 512;
  LR_CUDA(launch_pdl(gn_apply_coef_kernel, dim3(cdiv(P, chunk), n_img), dim3(kNormThreads), 0, st, 1, x0, c0, x1, c1, P,
                     chunk, scale, shift, do_silu, out));
  LR_LAUNCHED();
  return 0;
}

template <int VPL, bool kStatsOnly>
static int launch_ln_t(const __half* x, int M, int C, const float* gamma, const float* beta, float eps, __half* out,
                       cudaStream_t st) {
  const int tile_bytes = kLnTileRows * C * 2;
  int stages = (96 * 1024) / tile_bytes;
  if (stages > 8) stages = 8;
  if (stages < 2) stages = 2;
  const size_t smem = static_cast<size_t>(stages) * tile_bytes + static_cast<size_t>(2) * C * sizeof(float) +
                      static_cast<size_t>(2) * stages * sizeof(uint64_t);
  // per instantiation and device (slots 2..17)
  if (first_use_on_device(2 + (VPL - 1) * 2 + (kStatsOnly ? 1 : 0))) {
    LR_CUDA(cudaFuncSetAttribute(layernorm_kernel<VPL, kStatsOnly>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 200 * 1024));
  }
  const int ntiles = cdiv(M, kLnTileRows);
  int ctas_per_sm = static_cast<int>((220 * 1024) / (smem + 1024));
  if (ctas_per_sm > 4) ctas_per_sm = 4;
  if (ctas_per_sm < 1) ctas_per_sm = 1;
  int blocks = sm_count() * ctas_per_sm;
  if (blocks > ntiles) blocks = ntiles;
  LR_CUDA(launch_pdl(layernorm_kernel<VPL, kStatsOnly>, dim3(blocks), dim3(kLnThreads), smem, st, 1, x, M, C, gamma, beta,
                     eps, out, stages));
  LR_LAUNCHED();
  return 0;
}

template <bool kStatsOnly>
static int launch_ln_any(const __half* x, int M, int C, const float* gamma, const float* beta, float eps, __half* out,
                         cudaStream_t st) {
  LR_CHECK(C % 8 == 0, "layernorm: C must be a multiple of 8");
  LR_CHECK((reinterpret_cast<uintptr_t>(x) & 15) == 0, "layernorm: input must be 16-byte aligned");
  const int nvec = C / 8;
  const int vpl = cdiv(nvec, 32);
  switch (vpl) {
    case 1: return launch_ln_t<1, kStatsOnly>(x, M, C, gamma, beta, eps, out, st);
    case 2: return launch_ln_t<2, kStatsOnly>(x, M, C, gamma, beta, eps, out, st);
    case 3: return launch_ln_t<3, kStatsOnly>(x, M, C, gamma, beta, eps, out, st);
    case 4: return launch_ln_t<4, kStatsOnly>(x, M, C, gamma, beta, eps, out, st);
    case 5: return launch_ln_t<5, kStatsOnly>(x, M, C, gamma, beta, eps, out, st);
    case 6: case 7: case 8: return launch_ln_t<8, kStatsOnly>(x, M, C, gamma, beta, eps, out, st);
    default: LR_CHECK(false, "layernorm: C > 2048 unsupported");
  }
  return 0;
}

int launch_layernorm(const __half* x, int M, int C, const float* gamma, const float* beta, float eps, __half* out,
                     cudaStream_t st) {
  return launch_ln_any<false>(x, M, C, gamma, beta, eps, out, st);
}

template <int VPL, int RPW>
static int launch_ln_stats_t(const __half* x, int M, int C, float eps, float* stats, cudaStream_t st) {
  int blocks = cdiv(M, 8 * RPW);
  const int cap = sm_count() * 8;
  if (blocks > cap) blocks = cap;
  LR_CUDA(launch_pdl(ln_stats_kernel<VPL, RPW>, dim3(blocks), dim3(256), 0, st, 1, x, M, C, eps,
                     reinterpret_cast<float2*>(stats)));
  LR_LAUNCHED();
  return 0;
}

int launch_ln_rows_finalize(const float* part, int ld, int slots, int M, int C, float eps, float* stats, cudaStream_t st) {
  LR_CHECK(part != nullptr && stats != nullptr && slots > 0 && ld >= slots, "ln_rows_finalize: bad arguments");
  LR_CUDA(launch_pdl(ln_rows_finalize_kernel, dim3(cdiv(M, 256)), dim3(256), 0, st, 1,
                     reinterpret_cast<const float2*>(part), ld, slots, M, C, eps, reinterpret_cast<float2*>(stats)));
  LR_LAUNCHED();
  return 0;
}

int launch_layernorm_stats(const __half* x, int M, int C, float eps, float* stats, cudaStream_t st) {
  static const bool tma_stats = getenv("LR_LN_STATS_TMA") != nullptr;  // A/B: the TMA-pipeline kernel in stats-only mode
  if (tma_stats) return launch_ln_any<true>(x, M, C, nullptr, nullptr, eps, reinterpret_cast<__half*>(stats), st);
  LR_CHECK(C % 8 == 0, "layernorm: C must be a multiple of 8");
  LR_CHECK((reinterpret_cast<uintptr_t>(x) & 15) == 0, "layernorm: input must be 16-byte aligned");
  switch (cdiv(C / 8, 32)) {
    case 1: return launch_ln_stats_t<1, 4>(x, M, C, eps, stats, st);
    case 2: return launch_ln_stats_t<2, 4>(x, M, C, eps, stats, st);
    case 3: return launch_ln_stats_t<3, 2>(x, M, C, eps, stats, st);
    case 4: return launch_ln_stats_t<4, 2>(x, M, C, eps, stats, st);
    case 5: return launch_ln_stats_t<5, 2>(x, M, C, eps, stats, st);
    case 6: case 7: case 8: return launch_ln_stats_t<8, 1>(x, M, C, eps, stats, st);
    default: LR_CHECK(false, "layernorm: C > 2048 unsupported");
  }
  return 0;
}

int launch_ln_fold(const __half* w, int rows, int K, const float* gamma, const float* beta, const float* bias, __half* wf,
                   float* s_out, float* bf_out, cudaStream_t st) {
  ln_fold_kernel<<<cdiv(rows, 8), 256, 0, st>>>(w, rows, K, gamma, beta, bias, wf, s_out, bf_out);
  LR_LAUNCHED();
  return 0;
}

int launch_im2col_nchw_f32(const float* x, int n_img, int cin, int H, int W, int kpad, __half* out, cudaStream_t st) {
  LR_CHECK(kpad % 8 == 0, "im2col: kpad must be a multiple of 8");
  const size_t total = static_cast<size_t>(n_img) * H * W * (kpad / 8);
  LR_CUDA(launch_pdl(im2col_nchw_f32_kernel, dim3(static_cast<unsigned>((total + 255) / 256)), dim3(256), 0, st, 1, x,
                     n_img, cin, H, W, kpad, out));
  LR_LAUNCHED();
  return 0;
}
int launch_upsample2x(const __half* x, int n_img, int H, int W, int C, __half* out, cudaStream_t st) {
  LR_CHECK(C % 8 == 0, "upsample: C must be a multiple of 8");
  const size_t total = static_cast<size_t>(n_img) * 4 * H * W * (C / 8);
  LR_CUDA(launch_pdl(upsample2x_nhwc_kernel, dim3(static_cast<unsigned>((total + 255) / 256)), dim3(256), 0, st, 1, x,
                     n_img, H, W, C, out));
  LR_LAUNCHED();
  return 0;
}
int launch_mv_gather(const __half* src, int ld_src, int ncols, int b, int v, int hh, int side, __half* dst,
                     cudaStream_t st) {
  LR_CHECK(ncols % 8 == 0 && ld_src % 8 == 0, "mv_gather: columns must be multiples of 8");
  const size_t total = static_cast<size_t>(b) * (v + 1) * hh * side * (ncols / 8);
  LR_CUDA(launch_pdl(mv_gather_kernel, dim3(static_cast<unsigned>((total + 255) / 256)), dim3(256), 0, st, 1, src, ld_src,
                     ncols, b, v, hh, side, dst));
  LR_LAUNCHED();
  return 0;
}
int launch_mv_scatter(const __half* src, int ncols, int b, int v, int hh, int side, __half* dst, cudaStream_t st) {
  LR_CHECK(ncols % 8 == 0, "mv_scatter: columns must be multiples of 8");
  const size_t total = static_cast<size_t>(b) * v * hh * 2 * side * (ncols / 8);
  LR_CUDA(launch_pdl(mv_scatter_kernel, dim3(static_cast<unsigned>((total + 255) / 256)), dim3(256), 0, st, 1, src, ncols,
                     b, v, hh, side, dst));
  LR_LAUNCHED();
  return 0;
}
int launch_sep_insert(const __half* x, const float* sep, int n_img, int H, int W, int C, __half* out, cudaStream_t st) {
  LR_CHECK(C % 8 == 0, "sep_insert: C must be a multiple of 8");
  const size_t total = static_cast<size_t>(n_img) * H * (W + 1) * (C / 8);
  LR_CUDA(launch_pdl(sep_insert_nhwc_kernel, dim3(static_cast<unsigned>((total + 255) / 256)), dim3(256), 0, st, 1, x, sep,
                     n_img, H, W, C, out));
  LR_LAUNCHED();
  return 0;
}
int launch_sep_remove(const __half* x, int n_img, int H, int W1, int C, __half* out, cudaStream_t st) {
  LR_CHECK(C % 8 == 0, "sep_remove: C must be a multiple of 8");
  const size_t total = static_cast<size_t>(n_img) * H * (W1 - 1) * (C / 8);
  LR_CUDA(launch_pdl(sep_remove_nhwc_kernel, dim3(static_cast<unsigned>((total + 255) / 256)), dim3(256), 0, st, 1, x,
                     n_img, H, W1, C, out));
  LR_LAUNCHED();
  return 0;
}
int launch_sep_insert_nchw_f32(const float* x, const float* sep, int n_img, int C, int H, int W, float* out,
                               cudaStream_t st) {
  const size_t total = static_cast<size_t>(n_img) * C * H * (W + 1);
  LR_CUDA(launch_pdl(sep_insert_nchw_f32_kernel, dim3(static_cast<unsigned>((total + 255) / 256)), dim3(256), 0, st, 1, x,
                     sep, n_img, C, H, W, out));
  LR_LAUNCHED();
  return 0;
}
int launch_cinput_to_nhwc(const float* x, int n_img, int C, int H, int Wc, int x_off, int Wh, __half* out,
                          cudaStream_t st) {
  const size_t total = static_cast<size_t>(n_img) * H * Wh * C;
  cinput_nchw_to_nhwc_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, st>>>(x, n_img, C, H, Wc, x_off, Wh,
                                                                                          out);
  LR_LAUNCHED();
  return 0;
}
int launch_vae_in(const float* z, int n_img, int e, int zc, int H, int W, float z_scale, const float* pq_w,
                  const float* pq_b, int kpad, __half* out, cudaStream_t st) {
  LR_CHECK(e <= kVaeMaxZ && zc <= kVaeMaxZ && 9 * zc <= kpad, "vae_in: latent channel count out of range");
  const size_t total = static_cast<size_t>(n_img) * H * W * 9;
  LR_CUDA(launch_pdl(vae_in_kernel, dim3(static_cast<unsigned>((total + 255) / 256)), dim3(256), 0, st, 1, z, n_img, e, zc,
                     H, W, z_scale, pq_w, pq_b, kpad, out));
  LR_LAUNCHED();
  return 0;
}
int launch_softmax_rows(__half* x, int rows, int T, size_t ld, cudaStream_t st) {
  LR_CHECK(T % 8 == 0 && ld % 8 == 0 && T <= 256 * 8 * kSoftmaxVecs, "softmax_rows: unsupported row length");
  LR_CUDA(launch_pdl(softmax_rows_kernel, dim3(rows), dim3(256), 0, st, 1, x, T, ld));
  LR_LAUNCHED();
  return 0;
}
int launch_transpose_f16(const __half* in, int T, int C, size_t ld_in, __half* out, cudaStream_t st) {
  LR_CUDA(launch_pdl(transpose_f16_kernel, dim3(cdiv(T, 32), cdiv(C, 32)), dim3(32, 8), 0, st, 1, in, T, C, ld_in, out));
  LR_LAUNCHED();
  return 0;
}
int launch_upfold_weights(const __half* w, int O, int I, __half* out, cudaStream_t st) {
  const size_t total = static_cast<size_t>(4) * O * 4 * I;
  upfold_weights_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, st>>>(w, O, I, out);
  LR_LAUNCHED();
  return 0;
}
int launch_cast_f32_f16(const float* x, size_t n, __half* out, cudaStream_t st) {
  cast_f32_f16_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, st>>>(x, n, out);
  LR_LAUNCHED();
  return 0;
}
int launch_nhwc_to_nchw_f32(const __half* x, int ld, int n_img, int cout, int H, int W, float* out, cudaStream_t st) {
  const size_t total = static_cast<size_t>(n_img) * cout * H * W;
  LR_CUDA(launch_pdl(nhwc_f16_to_nchw_f32_kernel, dim3(static_cast<unsigned>((total + 255) / 256)), dim3(256), 0, st, 1,
                     x, ld, n_img, cout, H, W, out));
  LR_LAUNCHED();
  return 0;
}
int launch_nchw_f32_to_nhwc(const float* x, int n_img, int C, int H, int W, __half* out, cudaStream_t st) {
  const size_t total = static_cast<size_t>(n_img) * C * H * W;
  nchw_f32_to_nhwc_f16_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, st>>>(x, n_img, C, H, W, out);
  LR_LAUNCHED();
  return 0;
}
int launch_small_linear(const float* in, int ld_in, int n_rows, int K, const __half* w, const float* bias, int n_out,
                        int silu_in, int silu_out, float* out, int ld_out, cudaStream_t st) {
  LR_CHECK(K % 8 == 0, "small_linear: K must be a multiple of 8");
  const int rows = n_rows < kMaxSmallBatch ? n_rows : kMaxSmallBatch;
  const size_t smem = static_cast<size_t>(rows) * K * sizeof(float);
  LR_CHECK(smem <= 48 * 1024, "small_linear: K too large for the activation staging buffer");
  LR_CHECK(ld_in % 4 == 0 && (reinterpret_cast<uintptr_t>(in) & 15) == 0, "small_linear: input rows must be 16-byte aligned");
  if (n_out <= 4096) {  // narrow layers: one output per warp keeps >= 40 CTAs x 4 busy instead of 40
    LR_CUDA(launch_pdl(small_linear_kernel<1>, dim3(cdiv(n_out, 8)), dim3(256), smem, st, 1, in, ld_in, n_rows, K, w, bias,
                       n_out, silu_in, silu_out, out, ld_out));
  } else {
    LR_CUDA(launch_pdl(small_linear_kernel<kSmallOutPerWarp>, dim3(cdiv(n_out, 8 * kSmallOutPerWarp)), dim3(256), smem, st, 1,
                       in, ld_in, n_rows, K, w, bias, n_out, silu_in, silu_out, out, ld_out));
  }
  LR_LAUNCHED();
  return 0;
}
int launch_timestep_embedding(const long long* t, int t_count, int n, int dim, float* out, cudaStream_t st) {
  const int total = n * (dim / 2);
  timestep_embedding_kernel<<<cdiv(total, 256), 256, 0, st>>>(t, t_count, n, dim, out);
  LR_LAUNCHED();
  return 0;
}
int launch_ddim_update(const float* x, const float* e_u, const float* e_c, const float* noise, float cfg, float a_t,
                       float a_prev, float sigma, float sqrt_one_minus_at, float temperature, size_t n, float* x_prev,
                       float* pred_x0, cudaStream_t st) {
  LR_CUDA(launch_pdl(ddim_update_kernel, dim3(static_cast<unsigned>((n + 255) / 256)), dim3(256), 0, st, 1, x, e_u, e_c,
                     noise, cfg, a_t, a_prev, sigma, sqrt_one_minus_at, temperature, n, x_prev, pred_x0));
  LR_LAUNCHED();
  return 0;
}
int launch_ddim_update_dev(const float* x, const float* e_u, const float* e_c, const float* noise, const float* coef,
                           float temperature, size_t n, float* x_prev, float* pred_x0, cudaStream_t st) {
  LR_CUDA(launch_pdl(ddim_update_dev_kernel, dim3(static_cast<unsigned>((n + 255) / 256)), dim3(256), 0, st, 1, x, e_u,
                     e_c, noise, coef, temperature, n, x_prev, pred_x0));
  LR_LAUNCHED();
  return 0;
}
int launch_repack_conv(const float* w, int O, int I, int ldk, __half* out, cudaStream_t st) {
  const size_t total = static_cast<size_t>(O) * ldk;
  repack_conv_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, st>>>(w, O, I, ldk, out);
  LR_LAUNCHED();
  return 0;
}
int launch_repack_linear(const float* w, int O, int I, int geglu, int dst_row0, __half* out, cudaStream_t st) {
  const size_t total = static_cast<size_t>(O) * I;
  repack_linear_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, st>>>(w, O, I, geglu, dst_row0, out);
  LR_LAUNCHED();
  return 0;
}
int launch_repack_bias(const float* b, int O, int geglu, float* out, cudaStream_t st) {
  repack_bias_kernel<<<cdiv(O, 256), 256, 0, st>>>(b, O, geglu, out);
  LR_LAUNCHED();
  return 0;
}

}  // namespace lr
