// UNet engine: builds the layer graph of UNetModel (reference ldm/modules/diffusionmodules/openaimodel.py:412-787)
// from the constructor config, owns fp16 repacked weights, and executes UNetModel.forward as a static plan of
// tcgen05 GEMM/conv ops, fused attention ops and HBM-bound norm kernels on NHWC fp16 activations.
#include <algorithm>
#include <cstdio>
#include <cstring>
#include <functional>
#include <map>
#include <memory>
#include <string>
#include <vector>

#include "../../include/lr_b200.h"
#include "ops.h"

namespace lr {

enum WKind { K_CONV3, K_LINEAR, K_LINEAR_GEGLU, K_VEC, K_VEC_GEGLU, K_F32 /* all elements, fp32, as is */ };

struct Weight {
  std::string name;
  int64_t shape[4] = {0, 0, 0, 0};
  int ndim = 0;
  WKind kind = K_VEC;
  size_t off = 0;    // element offset into the fp16 arena (matrices) or fp32 arena (vectors)
  int dst_row0 = 0;  // row offset inside a fused matrix (QKV / KV)
  int ld = 0;        // row length (elements) of the destination matrix
  bool loaded = false;
  int region = 0;    // 0: off is final; 1 / 2: off is relative to the contiguous emb_layers weight / bias block
};

struct ResW {
  int cin = 0, cout = 0;
  size_t gn1_g, gn1_b, conv1_w, conv1_b, emb_w, emb_b, gn2_g, gn2_b, conv2_w, conv2_b, skip_w, skip_b;
  bool has_skip = false;
  bool has_emb = true;  // UNet ResBlock: + Linear(SiLU(emb)) per image (openaimodel.py:263-272); VAE ResnetBlock: none
  float eps = 1e-5f;    // GroupNorm32 (util.py:208) 1e-5; the VAE's Normalize (model.py:45) 1e-6
  int emb_col0 = 0;  // column offset of this block's emb_layers output inside the batched [n, emb_total] buffer
};
struct TBlockW {
  size_t ln1_g, ln1_b, qkv_w, out1_w, out1_b;
  size_t ln2_g, ln2_b, q2_w, kv2_w, out2_w, out2_b;
  size_t ln3_g, ln3_b, ff1_w, ff1_b, ff2_w, ff2_b;
  // LayerNorm folded into its consuming Linear (gemm_tc.cuh, GemmParams::ln_stats): gamma-scaled fp16 copies of the
  // weights, their column sums and the beta-shifted biases, rebuilt by run_folds() whenever a weight changes
  size_t qkv_wf, qkv_s, qkv_bf, q2_wf, q2_s, q2_bf, ff1_wf, ff1_s, ff1_bf;
  int kv_slot = 0;
};
struct STW {
  int C = 0, heads = 0;
  size_t gn_g, gn_b, pin_w, pin_b, pout_w, pout_b;
  std::vector<TBlockW> blocks;
};
struct ConvW {
  int cin = 0, cout = 0;
  size_t w, b;
  // Upsample convs: the four 2x2 phase matrices [4][cout][4 * cin] of the folded nearest-x2 + 3x3 conv (gemm_tc.cuh
  // header), rebuilt by run_folds() whenever a weight changes
  bool is_up = false;
  size_t wfold = 0;
};
enum NodeKind { N_CONV_IN, N_RES, N_ST, N_DOWN, N_UP };
struct Node {
  NodeKind kind;
  int idx;
};

struct Act {
  __half* p = nullptr;
  int C = 0, H = 0, W = 0;
  // GroupNorm statistics left by the producing kernel (gemm_tc.cuh header): [n * ppi][C] float2 partials, or nullptr
  float* stats = nullptr;
  int ppi = 0;
};

struct Pool {
  struct Buf {
    void* p;
    size_t bytes;
    bool used;
  };
  std::vector<Buf> bufs;
  size_t total = 0;
  int acquire(size_t bytes, void** out) {
    bytes = (bytes + 255) & ~size_t(255);
    int best = -1;
    for (size_t i = 0; i < bufs.size(); ++i)
      if (!bufs[i].used && bufs[i].bytes >= bytes && (best < 0 || bufs[i].bytes < bufs[best].bytes))
        best = static_cast<int>(i);
    if (best >= 0 && bufs[best].bytes <= bytes * 2) {
      bufs[best].used = true;
      *out = bufs[best].p;
      return 0;
    }
    void* p = nullptr;
    LR_CUDA(cudaMalloc(&p, bytes));
    bufs.push_back({p, bytes, true});
    total += bytes;
    *out = p;
    return 0;
  }
  void release(void* p) {
    for (auto& b : bufs)
      if (b.p == p) b.used = false;
  }
  void clear() {
    for (auto& b : bufs) cudaFree(b.p);
    bufs.clear();
    total = 0;
  }
};

}  // namespace lr

using namespace lr;

// Shared machinery of the two engines (UNet forward, first-stage decoder): weight table + fp16 / fp32 arenas, the
// activation pool, the static plan (list of launches) and the planners of the blocks both networks are made of.
struct lr_engine {
  lr_unet_cfg cfg;
  std::vector<Weight> weights;
  std::map<std::string, int> windex;
  size_t half_elems = 0, float_elems = 0;
  __half* harena = nullptr;
  float* farena = nullptr;

  std::vector<ResW> res;
  std::vector<STW> sts;
  std::vector<ConvW> convs;  // conv_in, downs, ups, head
  std::vector<std::vector<Node>> input_blocks, output_blocks;
  std::vector<Node> middle;
  size_t te0_w, te0_b, te2_w, te2_b, head_gn_g, head_gn_b;
  int conv_in_idx = 0, head_idx = 0, kpad_in = 0;
  int n_kv_slots = 0;
  int emb_total = 0;            // sum of ResBlock out_channels
  size_t emb_w_base = 0, emb_b_base = 0;
  int n_gn_sites = 0;
  // NVSUnetModel(use_sep=True): one learned separator token per channel count (NVS_ldm.py:24-31)
  std::map<int, size_t> sep_w;
  // NVS input refinement: c_input staged as an NHWC fp16 residual of the input conv (NVS_ldm.py:64-68)
  __half* cin_h = nullptr;
  int cin_n = 0, cin_H = 0, cin_Wh = 0;
  bool cin_on = false;
  int pcin = 0;  // planned with c_input

  // ---- plan state ----
  int pn = 0, ph = 0, pw = 0;  // planned shape
  long long plan_generation = 0;  // bumped whenever the plan (and with it every activation pointer) is rebuilt
  int pshared = 0;             // planned with the CFG-pair shared prefix
  int shared_ns = 0;           // > 0 while planning the shared prefix: number of images actually computed
  Pool pool;
  struct Step {
    std::function<int(cudaStream_t)> fn;
    int cls;       // 0 gemm/conv (gemm_conv_kernel), 1 attention, 2 groupnorm, 3 layernorm, 4 other
    double flops;  // algorithmic FLOPs of this step
    std::string desc;
  };
  std::vector<Step> steps;
  bool profiling = false;
  std::vector<cudaEvent_t> prof_events;  // steps.size() + 1 events of the last profiled forward
  void push(std::function<int(cudaStream_t)> fn, int cls = 4, double fl = 0.0, std::string desc = "") {
    steps.push_back(Step{std::move(fn), cls, fl, std::move(desc)});
  }
  std::vector<std::unique_ptr<ConvOp>> conv_ops;
  std::vector<std::unique_ptr<AttnOp>> attn_ops;
  double flops = 0;
  const float* in_x = nullptr;     // bound per call
  const int64_t* in_t = nullptr;
  float* out_y = nullptr;
  // context cache
  int ctx_n = 0, ctx_L = 0;
  __half* ctx_h = nullptr;               // [n*L, context_dim]
  std::vector<__half*> kv;               // per kv slot [n*L, 2C]
  std::vector<int> kv_C;
  std::vector<std::unique_ptr<ConvOp>> kv_ops;
  bool ctx_valid = false;
  size_t persistent_bytes = 0;

  virtual ~lr_engine() {
    for (auto e : prof_events) cudaEventDestroy(e);
    pool.clear();
    if (harena) cudaFree(harena);
    if (farena) cudaFree(farena);
    if (ctx_h) cudaFree(ctx_h);
    if (cin_h) cudaFree(cin_h);
    for (auto p : kv)
      if (p) cudaFree(p);
  }

  // ------------------------------------------------------------------------------------------------------
  size_t reg(const std::string& name, std::vector<int64_t> shape, WKind kind, size_t off, int dst_row0, int ld) {
    Weight w;
    w.name = name;
    w.ndim = static_cast<int>(shape.size());
    for (int i = 0; i < w.ndim; ++i) w.shape[i] = shape[i];
    w.kind = kind;
    w.off = off;
    w.dst_row0 = dst_row0;
    w.ld = ld;
    windex[name] = static_cast<int>(weights.size());
    weights.push_back(w);
    return off;
  }
  size_t halloc(size_t elems) {
    size_t o = half_elems;
    half_elems += (elems + 127) & ~size_t(127);
    return o;
  }
  size_t falloc(size_t elems) {
    size_t o = float_elems;
    float_elems += (elems + 63) & ~size_t(63);
    return o;
  }
  size_t reg_vec(const std::string& name, int n, bool geglu = false) {
    return reg(name, {n}, geglu ? K_VEC_GEGLU : K_VEC, falloc(n), 0, 0);
  }
  size_t reg_linear(const std::string& name, int O, int I, bool geglu = false) {
    return reg(name, {O, I}, geglu ? K_LINEAR_GEGLU : K_LINEAR, halloc(static_cast<size_t>(O) * I), 0, I);
  }
  size_t reg_conv3(const std::string& name, int O, int I, int ld) {
    return reg(name, {O, I, 3, 3}, K_CONV3, halloc(static_cast<size_t>(O) * ld), 0, ld);
  }

  int add_res(const std::string& pfx, int cin, int cout, int temb) {
    ResW r;
    r.cin = cin;
    r.cout = cout;
    r.gn1_g = reg_vec(pfx + "in_layers.0.weight", cin);
    r.gn1_b = reg_vec(pfx + "in_layers.0.bias", cin);
    r.conv1_w = reg_conv3(pfx + "in_layers.2.weight", cout, cin, 9 * cin);
    r.conv1_b = reg_vec(pfx + "in_layers.2.bias", cout);
    // all emb_layers (openaimodel.py:217-223) live in one contiguous [emb_total, temb] matrix: one launch per forward
    r.emb_col0 = emb_total;
    r.emb_w = reg(pfx + "emb_layers.1.weight", {cout, temb}, K_LINEAR, static_cast<size_t>(emb_total) * temb, 0, temb);
    weights.back().region = 1;
    r.emb_b = reg(pfx + "emb_layers.1.bias", {cout}, K_VEC, static_cast<size_t>(emb_total), 0, 0);
    weights.back().region = 2;
    emb_total += cout;
    r.gn2_g = reg_vec(pfx + "out_layers.0.weight", cout);
    r.gn2_b = reg_vec(pfx + "out_layers.0.bias", cout);
    r.conv2_w = reg_conv3(pfx + "out_layers.3.weight", cout, cout, 9 * cout);
    r.conv2_b = reg_vec(pfx + "out_layers.3.bias", cout);
    r.has_skip = cin != cout;
    if (r.has_skip) {
      // nn.Conv2d(cin, cout, 1): [cout, cin, 1, 1] (openaimodel.py:237-240) — same memory layout as a Linear
      r.skip_w = reg(pfx + "skip_connection.weight", {cout, cin, 1, 1}, K_LINEAR,
                     halloc(static_cast<size_t>(cout) * cin), 0, cin);
      r.skip_b = reg_vec(pfx + "skip_connection.bias", cout);
    }
    res.push_back(r);
    return static_cast<int>(res.size()) - 1;
  }
  int add_st(const std::string& pfx, int C) {
    STW s;
    s.C = C;
    s.heads = C / cfg.num_head_channels;
    s.gn_g = reg_vec(pfx + "norm.weight", C);
    s.gn_b = reg_vec(pfx + "norm.bias", C);
    if (cfg.use_linear_in_transformer) {
      s.pin_w = reg_linear(pfx + "proj_in.weight", C, C);
    } else {
      s.pin_w = reg(pfx + "proj_in.weight", {C, C, 1, 1}, K_LINEAR, halloc(static_cast<size_t>(C) * C), 0, C);
    }
    s.pin_b = reg_vec(pfx + "proj_in.bias", C);
    const int ctx = cfg.context_dim;
    for (int d = 0; d < cfg.transformer_depth; ++d) {
      const std::string b = pfx + "transformer_blocks." + std::to_string(d) + ".";
      TBlockW t;
      // key order follows the reference module registration order (attention.py:259-268): attn1, ff, attn2, norms
      t.qkv_w = halloc(static_cast<size_t>(3) * C * C);
      reg(b + "attn1.to_q.weight", {C, C}, K_LINEAR, t.qkv_w, 0, C);
      reg(b + "attn1.to_k.weight", {C, C}, K_LINEAR, t.qkv_w, C, C);
      reg(b + "attn1.to_v.weight", {C, C}, K_LINEAR, t.qkv_w, 2 * C, C);
      t.out1_w = reg_linear(b + "attn1.to_out.0.weight", C, C);
      t.out1_b = reg_vec(b + "attn1.to_out.0.bias", C);
      t.ff1_w = reg_linear(b + "ff.net.0.proj.weight", 8 * C, C, true);
      t.ff1_b = reg_vec(b + "ff.net.0.proj.bias", 8 * C, true);
      t.ff2_w = reg_linear(b + "ff.net.2.weight", C, 4 * C);
      t.ff2_b = reg_vec(b + "ff.net.2.bias", C);
      t.q2_w = reg_linear(b + "attn2.to_q.weight", C, C);
      t.kv2_w = halloc(static_cast<size_t>(2) * C * ctx);
      reg(b + "attn2.to_k.weight", {C, ctx}, K_LINEAR, t.kv2_w, 0, ctx);
      reg(b + "attn2.to_v.weight", {C, ctx}, K_LINEAR, t.kv2_w, C, ctx);
      t.out2_w = reg_linear(b + "attn2.to_out.0.weight", C, C);
      t.out2_b = reg_vec(b + "attn2.to_out.0.bias", C);
      t.ln1_g = reg_vec(b + "norm1.weight", C);
      t.ln1_b = reg_vec(b + "norm1.bias", C);
      t.ln2_g = reg_vec(b + "norm2.weight", C);
      t.ln2_b = reg_vec(b + "norm2.bias", C);
      t.ln3_g = reg_vec(b + "norm3.weight", C);
      t.ln3_b = reg_vec(b + "norm3.bias", C);
      t.qkv_wf = halloc(static_cast<size_t>(3) * C * C);
      t.qkv_s = falloc(3 * C);
      t.qkv_bf = falloc(3 * C);
      t.q2_wf = halloc(static_cast<size_t>(C) * C);
      t.q2_s = falloc(C);
      t.q2_bf = falloc(C);
      t.ff1_wf = halloc(static_cast<size_t>(8) * C * C);
      t.ff1_s = falloc(8 * C);
      t.ff1_bf = falloc(8 * C);
      t.kv_slot = n_kv_slots++;
      kv_C.push_back(C);
      s.blocks.push_back(t);
    }
    if (cfg.use_linear_in_transformer) {
      s.pout_w = reg_linear(pfx + "proj_out.weight", C, C);
    } else {
      s.pout_w = reg(pfx + "proj_out.weight", {C, C, 1, 1}, K_LINEAR, halloc(static_cast<size_t>(C) * C), 0, C);
    }
    s.pout_b = reg_vec(pfx + "proj_out.bias", C);
    sts.push_back(s);
    return static_cast<int>(sts.size()) - 1;
  }
  int add_conv(const std::string& pfx, int cin, int cout, int ld) {
    ConvW c;
    c.cin = cin;
    c.cout = cout;
    c.w = reg_conv3(pfx + "weight", cout, cin, ld);
    c.b = reg_vec(pfx + "bias", cout);
    convs.push_back(c);
    return static_cast<int>(convs.size()) - 1;
  }

  bool attn_at(int ds) const {
    for (int i = 0; i < cfg.n_attention_ds; ++i)
      if (cfg.attention_ds[i] == ds) return true;
    return false;
  }

  // Graph construction, same traversal as UNetModel.__init__ (openaimodel.py:527-731)
  int build_graph() {
    const int mc = cfg.model_channels, temb = 4 * mc;
    LR_CHECK(cfg.num_levels >= 1 && cfg.num_levels <= 8, "num_levels out of range");
    LR_CHECK(cfg.num_head_channels == 64, "only num_head_channels == 64 is supported by the attention kernel");
    LR_CHECK(mc % 64 == 0, "model_channels must be a multiple of 64");
    LR_CHECK(cfg.context_dim % 8 == 0, "context_dim must be a multiple of 8");
    te0_w = reg_linear("time_embed.0.weight", temb, mc);
    te0_b = reg_vec("time_embed.0.bias", temb);
    te2_w = reg_linear("time_embed.2.weight", temb, temb);
    te2_b = reg_vec("time_embed.2.bias", temb);
    kpad_in = ((9 * cfg.in_channels + 63) / 64) * 64;
    conv_in_idx = add_conv("input_blocks.0.0.", cfg.in_channels, mc, kpad_in);
    input_blocks.push_back({Node{N_CONV_IN, conv_in_idx}});
    std::vector<int> chans{mc};
    int ch = mc, ds = 1, ib = 1;
    for (int level = 0; level < cfg.num_levels; ++level) {
      const int mult = cfg.channel_mult[level];
      for (int nr = 0; nr < cfg.num_res_blocks[level]; ++nr) {
        const std::string p = "input_blocks." + std::to_string(ib) + ".";
        std::vector<Node> blk;
        blk.push_back({N_RES, add_res(p + "0.", ch, mult * mc, temb)});
        ch = mult * mc;
        if (attn_at(ds)) blk.push_back({N_ST, add_st(p + "1.", ch)});
        input_blocks.push_back(blk);
        chans.push_back(ch);
        ++ib;
      }
      if (level != cfg.num_levels - 1) {
        const std::string p = "input_blocks." + std::to_string(ib) + ".0.op.";
        input_blocks.push_back({Node{N_DOWN, add_conv(p, ch, ch, 9 * ch)}});
        chans.push_back(ch);
        ds *= 2;
        ++ib;
      }
    }
    middle.push_back({N_RES, add_res("middle_block.0.", ch, ch, temb)});
    middle.push_back({N_ST, add_st("middle_block.1.", ch)});
    middle.push_back({N_RES, add_res("middle_block.2.", ch, ch, temb)});
    int ob = 0;
    for (int level = cfg.num_levels - 1; level >= 0; --level) {
      const int mult = cfg.channel_mult[level];
      for (int i = 0; i <= cfg.num_res_blocks[level]; ++i) {
        const int ich = chans.back();
        chans.pop_back();
        const std::string p = "output_blocks." + std::to_string(ob) + ".";
        std::vector<Node> blk;
        blk.push_back({N_RES, add_res(p + "0.", ch + ich, mc * mult, temb)});
        ch = mc * mult;
        int sub = 1;
        if (attn_at(ds)) {
          blk.push_back({N_ST, add_st(p + std::to_string(sub) + ".", ch)});
          ++sub;
        }
        if (level > 0 && i == cfg.num_res_blocks[level]) {
          blk.push_back({N_UP, add_conv(p + std::to_string(sub) + ".conv.", ch, ch, 9 * ch)});
          mark_up_conv(blk.back().idx);
          ds /= 2;
        }
        output_blocks.push_back(blk);
        ++ob;
      }
    }
    head_gn_g = reg_vec("out.0.weight", ch);
    head_gn_b = reg_vec("out.0.bias", ch);
    LR_CHECK(ch == mc, "head channel mismatch");
    head_idx = add_conv("out.2.", mc, cfg.out_channels, 9 * mc);
    if (cfg.use_sep) {
      // channel counts that receive a separator, in order of first use: the inputs of every block that does not end
      // in a Downsample / Upsample (for SD2: 9, 320, 640, 1280, 2560, 1920, 960 - the reference's hard-coded list)
      std::vector<int> order;
      auto want = [&](int c) {
        if (std::find(order.begin(), order.end(), c) == order.end()) order.push_back(c);
      };
      auto in_ch = [&](const std::vector<Node>& blk) {
        const Node& f = blk[0];
        if (f.kind == N_CONV_IN) return cfg.in_channels;
        if (f.kind == N_RES) return res[f.idx].cin;
        return convs[f.idx].cin;
      };
      auto sep_block = [&](const std::vector<Node>& blk) { return blk.back().kind != N_DOWN && blk.back().kind != N_UP; };
      for (const auto& blk : input_blocks)
        if (sep_block(blk)) want(in_ch(blk));
      want(in_ch(middle));
      for (const auto& blk : output_blocks)
        if (sep_block(blk)) want(in_ch(blk));
      for (int c : order) sep_w[c] = reg_vec("sep_token." + std::to_string(c), c);
    }
    emb_w_base = halloc(static_cast<size_t>(emb_total) * temb);
    emb_b_base = falloc(emb_total);
    for (Weight& w : weights) {
      if (w.region == 1) w.off += emb_w_base;
      if (w.region == 2) w.off += emb_b_base;
    }
    return 0;
  }

  bool fold_dirty = true;
  bool ln_fold = getenv("LR_NO_LN_FOLD") == nullptr;
  bool up_fold = getenv("LR_NO_UPFOLD") == nullptr;
  // marks conv `idx` as the conv of an Upsample and reserves its folded weights
  void mark_up_conv(int idx) {
    ConvW& c = convs[idx];
    c.is_up = true;
    c.wfold = halloc(static_cast<size_t>(16) * c.cout * c.cin);
  }
  int run_folds(cudaStream_t st) {
    if (!fold_dirty) return 0;
    if (up_fold)
      for (const ConvW& c : convs)
        if (c.is_up) LR_TRY(launch_upfold_weights(H(c.w), c.cout, c.cin, H(c.wfold), st));
    if (!ln_fold) {
      fold_dirty = false;
      return 0;
    }
    for (const STW& s : sts) {
      const int C = s.C;
      for (const TBlockW& b : s.blocks) {
        LR_TRY(launch_ln_fold(H(b.qkv_w), 3 * C, C, F(b.ln1_g), F(b.ln1_b), nullptr, H(b.qkv_wf), F(b.qkv_s), F(b.qkv_bf), st));
        LR_TRY(launch_ln_fold(H(b.q2_w), C, C, F(b.ln2_g), F(b.ln2_b), nullptr, H(b.q2_wf), F(b.q2_s), F(b.q2_bf), st));
        LR_TRY(launch_ln_fold(H(b.ff1_w), 8 * C, C, F(b.ln3_g), F(b.ln3_b), F(b.ff1_b), H(b.ff1_wf), F(b.ff1_s), F(b.ff1_bf), st));
      }
    }
    fold_dirty = false;
    return 0;
  }

  int ensure_arenas() {
    if (harena == nullptr) {
      LR_CUDA(cudaMalloc(&harena, half_elems * sizeof(__half)));
      LR_CUDA(cudaMemset(harena, 0, half_elems * sizeof(__half)));
      LR_CUDA(cudaMalloc(&farena, float_elems * sizeof(float)));
      LR_CUDA(cudaMemset(farena, 0, float_elems * sizeof(float)));
      persistent_bytes += half_elems * sizeof(__half) + float_elems * sizeof(float);
    }
    return 0;
  }
  __half* H(size_t off) { return harena + off; }
  float* F(size_t off) { return farena + off; }

  // ------------------------------------------------------------------------------------------------------
  // plan helpers
  // ------------------------------------------------------------------------------------------------------
  int acquire_h(size_t elems, __half** out) {
    void* p = nullptr;
    LR_TRY(pool.acquire(elems * sizeof(__half), &p));
    *out = static_cast<__half*>(p);
    return 0;
  }
  int acquire_f(size_t elems, float** out) {
    void* p = nullptr;
    LR_TRY(pool.acquire(elems * sizeof(float), &p));
    *out = static_cast<float*>(p);
    return 0;
  }
  // Statistics tables live exactly as long as the activation buffer they describe.
  std::map<const void*, void*> stats_of;
  void release(void* p) {
    auto it = stats_of.find(p);
    if (it != stats_of.end()) {
      pool.release(it->second);
      stats_of.erase(it);
    }
    pool.release(p);
  }
  // GroupNorm sites. Three implementations exist (gemm_tc.cuh header and profiles/r2_ab_gn_fusion.txt have the
  // measurements behind the defaults):
  //   1. consumer-fused: statistics from the producer's epilogue, gn_finalize_kernel, and the consuming GEMM applies the
  //      affine map in shared memory (transform warps). Default for SpatialTransformer norm -> proj_in on images of at
  //      least gn_fuse_linear_min_rows tokens (64x128 latents: +17 us on the GEMM instead of a 42 us pass); 3x3 convs
  //      only with LR_GN_FUSE_CONV=1 (the SiLU of 11.5 K elements per k-chunk outlasts the chunk's MMAs).
  //   2. coefficient apply (LR_GN_COEF_APPLY=1): producer statistics + ONE read / ONE write pass (gn_apply_coef_kernel);
  //      measured 28 + 4 us against 42 us at the top level, nothing below it: off by default.
  //   3. the stand-alone kernels (elementwise.cuh): everything else.
  bool gn_fuse = getenv("LR_NO_GN_FUSE") == nullptr;
  bool gn_fuse_conv = gn_fuse && getenv("LR_GN_FUSE_CONV") != nullptr && atoi(getenv("LR_GN_FUSE_CONV")) != 0;
  bool gn_coef_apply = gn_fuse && getenv("LR_GN_COEF_APPLY") != nullptr && atoi(getenv("LR_GN_COEF_APPLY")) != 0;
  int gn_fuse_linear_min_rows = getenv("LR_GN_FUSE_LINEAR_MIN_ROWS") ? atoi(getenv("LR_GN_FUSE_LINEAR_MIN_ROWS")) : 4096;
  bool stats_all() const { return gn_fuse_conv || gn_coef_apply; }  // every GroupNorm input gets producer statistics
  bool st_fused(int P) const { return gn_fuse && P >= gn_fuse_linear_min_rows; }
  bool next_is_fused_st = false;  // set by plan_block: the ResBlock being planned feeds a consumer-fused norm
  // GroupNorm [+ SiLU] of concat(x0, x1) into `out` by way 2 or 3
  int add_gn_auto(const Act& x0, const Act& x1, int n, float eps, const float* g, const float* b, int silu, __half* out) {
    float *sc = nullptr, *sh = nullptr;
    int err = 0;
    if (gn_coef_apply && plan_gn_coef(x0, x1, n, eps, g, b, &sc, &sh, &err)) {
      const __half* p0 = x0.p;
      const __half* p1 = x1.p;
      const int c0 = x0.C, c1 = x1.C, P = x0.H * x0.W;
      push([=](cudaStream_t st) { return launch_gn_apply_coef(p0, c0, p1, c1, n, P, sc, sh, silu, out, st); }, 2, 0.0,
           "gn-apply n=" + std::to_string(n) + " P=" + std::to_string(P) + " c=" + std::to_string(c0) + "+" +
               std::to_string(c1));
      release(sc);
      release(sh);
      return 0;
    }
    LR_TRY(err);
    return add_gn(x0.p, x0.C, x1.p, x1.C, n, x0.H * x0.W, eps, g, b, silu, out);
  }
  // Requests producer statistics for the output of conv spec `s` ([n_tab images] table so that a CFG-shared prefix can be
  // replicated); call before add_conv_step, then finish_stats with the built op.
  int want_stats(ConvSpec* s, int n_tab_factor = 1, bool needed = false) {
    s->stats_out = nullptr;
    if (!gn_fuse || !(needed || stats_all()) || s->ncols % 8 != 0) return 0;
    const size_t rows = conv_stats_rows(*s) * n_tab_factor;
    if (rows == 0) return 0;
    float* t;
    LR_TRY(acquire_f(rows * s->ncols * 2, &t));
    s->stats_out = t;
    return 0;
  }
  void finish_stats(const ConvSpec& s, const ConvOp* op, Act* out) {
    if (s.stats_out == nullptr) return;
    if (op->stats_ok) {
      out->stats = s.stats_out;
      out->ppi = op->stats_ppi;
      stats_of[out->p] = s.stats_out;
    } else {
      pool.release(s.stats_out);
    }
  }
  // GroupNorm site whose consumer applies the affine map itself: combines the producers' partials into scale / shift
  // [n][c0 + c1]. Returns false (nothing planned) when a source has no statistics.
  bool plan_gn_coef(const Act& x0, const Act& x1, int n, float eps, const float* g, const float* b, float** scale,
                    float** shift, int* err) {
    *err = 0;
    if (!gn_fuse || x0.stats == nullptr || (x1.p != nullptr && x1.stats == nullptr)) return false;
    const int C = x0.C + x1.C;
    float *sc, *sh;
    if ((*err = acquire_f(static_cast<size_t>(n) * C, &sc)) != 0) return false;
    if ((*err = acquire_f(static_cast<size_t>(n) * C, &sh)) != 0) return false;
    const float* p0 = x0.stats;
    const float* p1 = x1.stats;
    const int ppi0 = x0.ppi, ppi1 = x1.ppi, c0 = x0.C, c1 = x1.C, P = x0.H * x0.W;
    push([=](cudaStream_t st) { return launch_gn_finalize(p0, ppi0, c0, p1, ppi1, c1, n, P, 32, eps, g, b, sc, sh, st); },
         2, 0.0, "gn-finalize n=" + std::to_string(n) + " P=" + std::to_string(P) + " c=" + std::to_string(c0) + "+" +
                     std::to_string(c1));
    *scale = sc;
    *shift = sh;
    return true;
  }
  int add_conv_step(const ConvSpec& s_in, ConvOp** out_op = nullptr) {
    ConvSpec s = s_in;
    void* ws = nullptr;
    if (s.taps == 9 && s.workspace == nullptr) {
      // split-K scratch for convs with few output pixels (the 8x16 level): plan-owned, recycled right after this step
      const int Ho = s.stride == 1 ? s.in_h : (s.in_h - 1) / 2 + 1, Wo = s.stride == 1 ? s.in_w : (s.in_w - 1) / 2 + 1;
      const size_t m_out = static_cast<size_t>(s.n_img) * Ho * Wo;
      if (Ho * Wo <= kSplitKMaxPixels) {
        const size_t bytes = 3 * m_out * s.ncols * sizeof(float);
        LR_TRY(pool.acquire(bytes, &ws));
        s.workspace = static_cast<float*>(ws);
        s.workspace_bytes = bytes;
      }
    }
    auto op = std::make_unique<ConvOp>();
    LR_TRY(build_conv_op(op.get(), s));
    if (ws) release(ws);
    flops += op->flops;
    ConvOp* raw = op.get();
    if (out_op) *out_op = raw;
    conv_ops.push_back(std::move(op));
    char d[200];
    snprintf(d, sizeof(d), "%s n=%d %dx%d s%d c=%d+%d->%d%s bn=%d st=%d tiles=%d%s%s%s",
             s.taps == 9 ? "conv3x3" : "linear", s.n_img, s.in_h, s.in_w, s.stride, s.c0, s.c1, s.ncols,
             s.geglu ? " geglu" : (s.residual ? " +res" : ""), raw->block_n, raw->stages, raw->tiles,
             raw->ksplit > 1 ? " splitK3" : "", raw->xf ? (s.xf_silu ? " gn+silu" : " gn") : "",
             raw->stats_ok ? " stats" : "");
    push([raw](cudaStream_t st) { return launch_conv_op(*raw, st); }, 0, raw->flops, d);
    return 0;
  }
  // Row-statistics hand-over between the Linears of a transformer block (gemm_tc.cuh GemmParams::rowstats_out / ln_part):
  // the GEMM that writes the residual stream leaves per-row (sum, sum of squares) partials, the GEMM that consumes its
  // LayerNorm finishes mean / rstd in its epilogue; no statistics pass over the activation in between.
  struct RowStats {
    float* table = nullptr;  // [M][ld] float2
    int ld = 0;
    int slots = 0;           // valid entries per row (0: the current content of the stream has no partials)
  };
  // Measured on B200 (round 2, profiles/r2_ab_ln_rowstats.txt), before and after the lean epilogue: the 48
  // ln_stats_kernel passes disappear, but the producers' extra work plus either the consumers' scattered partial loads
  // or 48 tiny finalize kernels cost as much (18.4 ms either way). OFF by default (LR_LN_ROWSTATS=1 enables;
  // parity-tested).
  bool ln_rowstats = getenv("LR_LN_ROWSTATS") != nullptr && atoi(getenv("LR_LN_ROWSTATS")) != 0;
  int add_linear(const __half* a, int M, int K, const __half* w, int ncols, const float* bias, const __half* residual,
                 int ld_res, __half* out, int ld_out, int geglu, const float* ln_stats = nullptr,
                 const float* ln_s = nullptr, const RowStats* ln_from = nullptr, RowStats* stats_to = nullptr) {
    ConvSpec s;
    s.ln_stats = ln_stats;
    s.ln_s = ln_s;
    if (ln_from != nullptr) {
      // LayerNorm statistics from the producer's per-row partials: a tiny kernel turns them into the (mean, rstd) table
      // the stats pass would have written (the consumer can also sum them itself, ConvSpec::ln_part: measured slower,
      // every N-tile of the consumer repeats the scattered loads - profiles/r2_ab_ln_rowstats.txt)
      LR_CHECK(ln_stats != nullptr, "internal: row statistics need the (mean, rstd) table of the block");
      const float* tb = ln_from->table;
      const int ld = ln_from->ld, slots = ln_from->slots;
      float* dst = const_cast<float*>(ln_stats);
      push([=](cudaStream_t st) { return launch_ln_rows_finalize(tb, ld, slots, M, K, 1e-5f, dst, st); }, 3, 0.0,
           "layernorm-finalize M=" + std::to_string(M) + " C=" + std::to_string(K));
    }
    if (stats_to != nullptr && stats_to->table != nullptr) {
      s.rowstats_out = stats_to->table;
      s.rowstats_ld = stats_to->ld;
    }
    s.a0 = a;
    s.c0 = K;
    s.lda0 = K;
    s.n_img = 1;
    s.in_h = 1;
    s.in_w = M;
    s.taps = 1;
    s.w = w;
    s.ldw = K;
    s.ncols = ncols;
    s.bias = bias;
    s.residual = residual;
    s.ld_res = ld_res;
    s.out = out;
    s.ld_out = ld_out;
    s.geglu = geglu;
    ConvOp* op = nullptr;
    LR_TRY(add_conv_step(s, &op));
    if (stats_to != nullptr) stats_to->slots = op->rowstats_slots;
    return 0;
  }
  // GroupNorm statistics: one (sum, sumsq) slot per call site, all zeroed by a single memset at the start of forward
  unsigned char* gn_stats = nullptr;
  size_t gn_site_bytes = 0;
  int gn_sites_planned = 0, gn_sites_cap = 0;
  int add_gn(const __half* x0, int c0, const __half* x1, int c1, int n, int P, float eps, const float* g,
             const float* b, int silu, __half* out) {
    LR_CHECK(gn_sites_planned < gn_sites_cap, "internal: GroupNorm statistics arena too small");
    unsigned char* st_ = gn_stats + static_cast<size_t>(gn_sites_planned++) * gn_site_bytes;
    push([=](cudaStream_t st) {
      return launch_groupnorm(x0, c0, x1, c1, n, P, 32, eps, g, b, silu, st_, 1, out, st);
    }, 2, 0.0, "groupnorm n=" + std::to_string(n) + " P=" + std::to_string(P) + " c=" + std::to_string(c0) + "+" +
                   std::to_string(c1));
    return 0;
  }
  int add_ln_stats(const __half* x, int M, int C, float* stats) {
    push([=](cudaStream_t st) { return launch_layernorm_stats(x, M, C, 1e-5f, stats, st); }, 3, 0.0,
         "layernorm-stats M=" + std::to_string(M) + " C=" + std::to_string(C));
    return 0;
  }
  int add_ln(const __half* x, int M, int C, const float* g, const float* b, __half* out) {
    push([=](cudaStream_t st) { return launch_layernorm(x, M, C, g, b, 1e-5f, out, st); }, 3, 0.0,
         "layernorm M=" + std::to_string(M) + " C=" + std::to_string(C));
    return 0;
  }
  int add_attn(const AttnSpec& s) {
    auto op = std::make_unique<AttnOp>();
    LR_TRY(build_attn_op(op.get(), s));
    flops += op->flops;
    AttnOp* raw = op.get();
    attn_ops.push_back(std::move(op));
    char d[200];
    snprintf(d, sizeof(d), "attention b=%d h=%d tq=%d tk=%d", s.batch, s.heads, s.tq, s.tk);
    push([raw](cudaStream_t st) { return launch_attn_op(*raw, st); }, 1, raw->flops, d);
    return 0;
  }

  float* emb_all = nullptr;  // [n, emb_total]: every ResBlock's emb_layers output

  // CFG pair: the first half of a [n, ...] buffer was computed once for both halves; replicate it
  void add_dup_half(__half* p, size_t half_elems) {
    push([=](cudaStream_t st) {
      cudaError_t e = cudaMemcpyAsync(p + half_elems, p, half_elems * sizeof(__half), cudaMemcpyDeviceToDevice, st);
      if (e != cudaSuccess) {
        set_error(std::string("cudaMemcpyAsync(dup half): ") + cudaGetErrorString(e));
        return 1;
      }
      return 0;
    });
  }

  // ... and its statistics table: the rows of the first `ns` images are replicated for the second half
  void add_dup_stats(const Act& a, int ns) {
    if (a.stats == nullptr) return;
    float* t = a.stats;
    const size_t half = static_cast<size_t>(ns) * a.ppi * a.C * 2;  // floats
    push([=](cudaStream_t st) {
      cudaError_t e = cudaMemcpyAsync(t + half, t, half * sizeof(float), cudaMemcpyDeviceToDevice, st);
      if (e != cudaSuccess) {
        set_error(std::string("cudaMemcpyAsync(dup stats): ") + cudaGetErrorString(e));
        return 1;
      }
      return 0;
    });
  }

  // ResBlock._forward (openaimodel.py:254-274), x = concat(x0, x1) when x1.p != nullptr
  //
  // GroupNorm + SiLU of both halves are fused into the convs that consume them whenever the producers left statistics
  // and the conv runs in halo mode (gemm_tc.cuh header): no normalised tensor is written or re-read. Otherwise (the
  // 8x16 level, whose convs run split-K on 128-pixel images, and tiny test shapes) the stand-alone kernel runs.
  int plan_res(const ResW& r, Act x0, Act x1, int n_full, Act* out) {
    const int Hh = x0.H, Ww = x0.W, P = Hh * Ww;
    // inside the shared CFG prefix only the first shared_ns images are computed (buffers stay full size)
    const int n = shared_ns > 0 ? shared_ns : n_full;
    const int tab = n_full / n;  // statistics tables are sized for the full batch (the prefix is replicated afterwards)
    const size_t M = static_cast<size_t>(n) * P;
    const size_t Mfull = static_cast<size_t>(n_full) * P;
    LR_CHECK(x0.C + x1.C == r.cin, "resblock: channel mismatch");
    __half *xn = nullptr, *h, *hn = nullptr, *o, *skip = nullptr;
    Act hact;
    {
      ConvSpec s;
      s.n_img = n;
      s.in_h = Hh;
      s.in_w = Ww;
      s.taps = 9;
      s.w = H(r.conv1_w);
      s.ldw = 9 * r.cin;
      s.ncols = r.cout;
      s.bias = F(r.conv1_b);
      s.bias_img = r.has_emb ? emb_all + r.emb_col0 : nullptr;
      s.ld_bias_img = emb_total;
      s.ld_out = r.cout;
      float *sc = nullptr, *sh = nullptr;
      int err = 0;
      const bool fused = gn_fuse_conv && conv_is_halo(s) &&
                         plan_gn_coef(x0, x1, n, r.eps, F(r.gn1_g), F(r.gn1_b), &sc, &sh, &err);
      LR_TRY(err);
      if (fused) {
        s.a0 = x0.p;
        s.c0 = x0.C;
        s.lda0 = x0.C;
        s.a1 = x1.p;
        s.c1 = x1.C;
        s.lda1 = x1.C;
        s.xf_scale = sc;
        s.xf_shift = sh;
        s.xf_silu = 1;
      } else {
        LR_TRY(acquire_h(M * r.cin, &xn));
        LR_TRY(add_gn_auto(x0, x1, n, r.eps, F(r.gn1_g), F(r.gn1_b), 1, xn));
        s.a0 = xn;
        s.c0 = r.cin;
        s.lda0 = r.cin;
      }
      LR_TRY(acquire_h(M * r.cout, &h));
      s.out = h;
      hact = Act{h, r.cout, Hh, Ww};
      LR_TRY(want_stats(&s));
      ConvOp* op;
      LR_TRY(add_conv_step(s, &op));
      finish_stats(s, op, &hact);
      if (xn) release(xn);
      if (sc) { release(sc); release(sh); }
    }
    const __half* resid;
    if (r.has_skip) {
      LR_TRY(acquire_h(M * r.cout, &skip));
      ConvSpec s;
      s.a0 = x0.p;
      s.c0 = x0.C;
      s.lda0 = x0.C;
      s.a1 = x1.p;
      s.c1 = x1.C;
      s.lda1 = x1.C;
      s.n_img = 1;
      s.in_h = 1;
      s.in_w = static_cast<int>(M);
      s.taps = 1;
      s.w = H(r.skip_w);
      s.ldw = r.cin;
      s.ncols = r.cout;
      s.bias = F(r.skip_b);
      s.out = skip;
      s.ld_out = r.cout;
      LR_TRY(add_conv_step(s));
      resid = skip;
    } else {
      LR_CHECK(x1.p == nullptr, "resblock: identity skip with concat input");
      resid = x0.p;
    }
    LR_TRY(acquire_h(Mfull * r.cout, &o));
    Act oact{o, r.cout, Hh, Ww};
    {
      ConvSpec s;
      s.c0 = r.cout;
      s.lda0 = r.cout;
      s.n_img = n;
      s.in_h = Hh;
      s.in_w = Ww;
      s.taps = 9;
      s.w = H(r.conv2_w);
      s.ldw = 9 * r.cout;
      s.ncols = r.cout;
      s.bias = F(r.conv2_b);
      s.residual = resid;
      s.ld_res = r.cout;
      s.out = o;
      s.ld_out = r.cout;
      float *sc = nullptr, *sh = nullptr;
      int err = 0;
      const bool fused = gn_fuse_conv && conv_is_halo(s) &&
                         plan_gn_coef(hact, Act{}, n, r.eps, F(r.gn2_g), F(r.gn2_b), &sc, &sh, &err);
      LR_TRY(err);
      if (fused) {
        s.a0 = h;
        s.xf_scale = sc;
        s.xf_shift = sh;
        s.xf_silu = 1;
      } else {
        LR_TRY(acquire_h(M * r.cout, &hn));
        LR_TRY(add_gn_auto(hact, Act{}, n, r.eps, F(r.gn2_g), F(r.gn2_b), 1, hn));
        s.a0 = hn;
      }
      LR_TRY(want_stats(&s, tab, next_is_fused_st));
      ConvOp* op;
      LR_TRY(add_conv_step(s, &op));
      finish_stats(s, op, &oact);
      if (hn) release(hn);
      if (sc) { release(sc); release(sh); }
    }
    release(h);
    if (skip) release(skip);
    *out = oact;
    return 0;
  }

  // SpatialTransformer.forward + BasicTransformerBlock._forward (attention.py:393-419, 279-283)
  int plan_st(const STW& s, Act x, int n_full, Act* out) {
    const int C = s.C, P = x.H * x.W;
    const int Mfull = n_full * P;
    LR_CHECK(x.C == C, "spatial transformer: channel mismatch");
    // CFG pair: everything up to and including the FIRST self-attention is independent of the context, so inside the
    // shared prefix it is computed for the first shared_ns images only and then replicated (ddim.py:317-326 feeds both
    // halves the same x / t / c_concat; only c_crossattn differs). Buffers are full size throughout.
    int n = shared_ns > 0 ? shared_ns : n_full;
    int M = n * P;
    __half *xn, *h, *t, *qkv, *a, *g, *o;
    LR_TRY(acquire_h(static_cast<size_t>(Mfull) * C, &h));
    // per-row LayerNorm partials of the residual stream h, rewritten by every Linear that writes h (RowStats)
    RowStats rs;
    if (ln_fold && ln_rowstats) {
      rs.ld = 2 * ((C + 31) / 32);
      LR_TRY(acquire_f(static_cast<size_t>(Mfull) * rs.ld * 2, &rs.table));
    }
    auto dup_rowstats = [&](int rows) {  // CFG-pair prefix: replicate the partials of the first `rows` rows
      if (rs.table == nullptr || rs.slots == 0) return;
      float* t_ = rs.table;
      const size_t half = static_cast<size_t>(rows) * rs.ld * 2;
      push([=](cudaStream_t st) {
        cudaError_t e = cudaMemcpyAsync(t_ + half, t_, half * sizeof(float), cudaMemcpyDeviceToDevice, st);
        if (e != cudaSuccess) {
          set_error(std::string("cudaMemcpyAsync(dup row stats): ") + cudaGetErrorString(e));
          return 1;
        }
        return 0;
      });
    };
    {
      // norm (GroupNorm eps 1e-6, no activation) -> proj_in (attention.py:399-404): the affine map is applied to the
      // activation tiles of the proj_in GEMM in shared memory when the producer of x left statistics
      float *sc = nullptr, *sh = nullptr;
      int err = 0;
      const bool fused = st_fused(P) && plan_gn_coef(x, Act{}, n, 1e-6f, F(s.gn_g), F(s.gn_b), &sc, &sh, &err);
      LR_TRY(err);
      if (fused) {
        ConvSpec cs;
        cs.a0 = x.p;
        cs.c0 = C;
        cs.lda0 = C;
        cs.n_img = 1;
        cs.in_h = 1;
        cs.in_w = M;
        cs.taps = 1;
        cs.w = H(s.pin_w);
        cs.ldw = C;
        cs.ncols = C;
        cs.bias = F(s.pin_b);
        cs.out = h;
        cs.ld_out = C;
        cs.xf_scale = sc;
        cs.xf_shift = sh;
        cs.xf_silu = 0;
        cs.xf_rows_per_img = P;
        if (rs.table != nullptr) {
          cs.rowstats_out = rs.table;
          cs.rowstats_ld = rs.ld;
        }
        ConvOp* op;
        LR_TRY(add_conv_step(cs, &op));
        rs.slots = op->rowstats_slots;
        release(sc);
        release(sh);
      } else {
        LR_TRY(acquire_h(static_cast<size_t>(Mfull) * C, &xn));
        LR_TRY(add_gn_auto(x, Act{}, n, 1e-6f, F(s.gn_g), F(s.gn_b), 0, xn));
        LR_TRY(add_linear(xn, M, C, H(s.pin_w), C, F(s.pin_b), nullptr, 0, h, C, 0, nullptr, nullptr, nullptr, &rs));
        release(xn);
      }
    }
    LR_TRY(acquire_h(static_cast<size_t>(Mfull) * C, &t));
    LR_TRY(acquire_h(static_cast<size_t>(Mfull) * C, &a));
    float* lnst;  // (mean, rstd) per token row for the folded LayerNorms
    LR_TRY(acquire_f(static_cast<size_t>(Mfull) * 2, &lnst));
    bool first_block = true;
    for (const TBlockW& b : s.blocks) {
      if (!first_block && n != n_full) {  // depth > 1: only block 0's self-attention is shared
        add_dup_half(h, static_cast<size_t>(M) * C);
        dup_rowstats(M);
        add_dup_half(x.p, static_cast<size_t>(M) * C);
        n = n_full;
        M = Mfull;
      }
      // self-attention: x = attn1(norm1(x)) + x
      LR_TRY(acquire_h(static_cast<size_t>(M) * 3 * C, &qkv));
      if (ln_fold) {  // LayerNorm folded into the QKV projection: stats pass + epilogue correction
        if (rs.slots == 0) LR_TRY(add_ln_stats(h, M, C, lnst));
        LR_TRY(add_linear(h, M, C, H(b.qkv_wf), 3 * C, F(b.qkv_bf), nullptr, 0, qkv, 3 * C, 0, lnst, F(b.qkv_s),
                          rs.slots > 0 ? &rs : nullptr));
      } else {
        LR_TRY(add_ln(h, M, C, F(b.ln1_g), F(b.ln1_b), t));
        LR_TRY(add_linear(t, M, C, H(b.qkv_w), 3 * C, nullptr, nullptr, 0, qkv, 3 * C, 0));
      }
      LR_CHECK(!(shared_ns > 0 && cfg.view_num > 1), "CFG-pair sharing is not implemented for the multiview UNet");
      if (cfg.view_num > 1 && cfg.concat_target) {
        // multiview_attention.py:436-462: rows are stitched [ref_i | target] canvases; attend over
        // [target(row 0), ref_1..ref_v] and write the target block back to every row.
        const int v = cfg.view_num - 1;
        LR_CHECK(n % v == 0, "multiview: UNet batch not divisible by view_num - 1");
        const int side = x.W / 2, hh = x.H;
        LR_CHECK(x.W == 2 * x.H, "multiview concat_target expects stitched canvases with W == 2*H (as the reference)");
        const int bs = n / v;
        const int Tp = (v + 1) * hh * side;
        const int Mp = bs * Tp;
        __half *qkv_r, *a_r, *h_r, *o_r;
        LR_TRY(acquire_h(static_cast<size_t>(Mp) * 3 * C, &qkv_r));
        LR_TRY(acquire_h(static_cast<size_t>(Mp) * C, &a_r));
        LR_TRY(acquire_h(static_cast<size_t>(Mp) * C, &h_r));
        LR_TRY(acquire_h(static_cast<size_t>(Mp) * C, &o_r));
        {
          const __half* src = qkv;
          push([=](cudaStream_t st) { return launch_mv_gather(src, 3 * C, 3 * C, bs, v, hh, side, qkv_r, st); });
          const __half* hsrc = h;
          push([=](cudaStream_t st) { return launch_mv_gather(hsrc, C, C, bs, v, hh, side, h_r, st); });
        }
        AttnSpec as;
        as.q = qkv_r; as.ldq = 3 * C; as.q_col0 = 0;
        as.k = qkv_r; as.ldk = 3 * C; as.k_col0 = C;
        as.v = qkv_r; as.ldv = 3 * C; as.v_col0 = 2 * C;
        as.out = a_r; as.ld_out = C;
        as.batch = bs; as.heads = s.heads; as.tq = Tp; as.tk = Tp;
        as.scale = 0.125f;
        LR_TRY(add_attn(as));
        LR_TRY(add_linear(a_r, Mp, C, H(b.out1_w), C, F(b.out1_b), h_r, C, o_r, C, 0));
        {
          __half* hdst = h;
          push([=](cudaStream_t st) { return launch_mv_scatter(o_r, C, bs, v, hh, side, hdst, st); });
          rs.slots = 0;  // h was rewritten by a copy kernel: the next LayerNorm needs its own statistics pass
        }
        release(qkv_r);
        release(a_r);
        release(h_r);
        release(o_r);
        release(qkv);
      } else {
        // multiview (multiview_attention.py:448,462, concat_target=False): '(b v) hw c -> b (v hw) c' is a pure
        // reshape of the token matrix, so only batch / sequence length change.
        const int v = cfg.view_num > 1 ? cfg.view_num : 1;
        LR_CHECK(n % v == 0, "multiview: batch not divisible by view_num");
        AttnSpec as;
        as.q = qkv; as.ldq = 3 * C; as.q_col0 = 0;
        as.k = qkv; as.ldk = 3 * C; as.k_col0 = C;
        as.v = qkv; as.ldv = 3 * C; as.v_col0 = 2 * C;
        as.out = a; as.ld_out = C;
        as.batch = n / v; as.heads = s.heads; as.tq = P * v; as.tk = P * v;
        as.scale = 0.125f;
        LR_TRY(add_attn(as));
        release(qkv);
        LR_TRY(add_linear(a, M, C, H(b.out1_w), C, F(b.out1_b), h, C, h, C, 0, nullptr, nullptr, nullptr, &rs));
      }
      if (first_block && n != n_full) {  // end of the shared prefix: replicate the residual stream and the ST input
        add_dup_half(h, static_cast<size_t>(M) * C);
        dup_rowstats(M);
        add_dup_half(x.p, static_cast<size_t>(M) * C);
        n = n_full;
        M = Mfull;
      }
      first_block = false;
      // cross-attention against the cached context K/V
      __half* q2;
      LR_TRY(acquire_h(static_cast<size_t>(M) * C, &q2));
      if (ln_fold) {
        if (rs.slots == 0) LR_TRY(add_ln_stats(h, M, C, lnst));
        LR_TRY(add_linear(h, M, C, H(b.q2_wf), C, F(b.q2_bf), nullptr, 0, q2, C, 0, lnst, F(b.q2_s),
                          rs.slots > 0 ? &rs : nullptr));
      } else {
        LR_TRY(add_ln(h, M, C, F(b.ln2_g), F(b.ln2_b), t));
        LR_TRY(add_linear(t, M, C, H(b.q2_w), C, nullptr, nullptr, 0, q2, C, 0));
      }
      {
        AttnSpec as;
        as.q = q2; as.ldq = C; as.q_col0 = 0;
        as.k = kv[b.kv_slot]; as.ldk = 2 * C; as.k_col0 = 0;
        as.v = kv[b.kv_slot]; as.ldv = 2 * C; as.v_col0 = C;
        as.out = a; as.ld_out = C;
        as.batch = n; as.heads = s.heads; as.tq = P; as.tk = ctx_L;
        as.scale = 0.125f;
        LR_TRY(add_attn(as));
      }
      release(q2);
      LR_TRY(add_linear(a, M, C, H(b.out2_w), C, F(b.out2_b), h, C, h, C, 0, nullptr, nullptr, nullptr, &rs));
      // GEGLU feed-forward
      LR_TRY(acquire_h(static_cast<size_t>(M) * 4 * C, &g));
      if (ln_fold) {
        if (rs.slots == 0) LR_TRY(add_ln_stats(h, M, C, lnst));
        LR_TRY(add_linear(h, M, C, H(b.ff1_wf), 8 * C, F(b.ff1_bf), nullptr, 0, g, 4 * C, 1, lnst, F(b.ff1_s),
                          rs.slots > 0 ? &rs : nullptr));
      } else {
        LR_TRY(add_ln(h, M, C, F(b.ln3_g), F(b.ln3_b), t));
        LR_TRY(add_linear(t, M, C, H(b.ff1_w), 8 * C, F(b.ff1_b), nullptr, 0, g, 4 * C, 1));
      }
      LR_TRY(add_linear(g, M, 4 * C, H(b.ff2_w), C, F(b.ff2_b), h, C, h, C, 0, nullptr, nullptr, nullptr, &rs));
      release(g);
    }
    release(t);
    release(a);
    release(lnst);
    if (rs.table) release(rs.table);
    LR_TRY(acquire_h(static_cast<size_t>(M) * C, &o));
    Act oact{o, C, x.H, x.W};
    {
      // proj_out (+ the transformer's input as residual): its output feeds the next ResBlock's GroupNorm
      ConvSpec cs;
      cs.a0 = h;
      cs.c0 = C;
      cs.lda0 = C;
      cs.n_img = 1;
      cs.in_h = 1;
      cs.in_w = M;
      cs.taps = 1;
      cs.w = H(s.pout_w);
      cs.ldw = C;
      cs.ncols = C;
      cs.bias = F(s.pout_b);
      cs.residual = x.p;
      cs.ld_res = C;
      cs.out = o;
      cs.ld_out = C;
      cs.stats_rows_per_img = P;
      LR_TRY(want_stats(&cs));
      ConvOp* op;
      LR_TRY(add_conv_step(cs, &op));
      finish_stats(cs, op, &oact);
    }
    release(h);
    *out = oact;
    return 0;
  }

  int plan_conv3(const ConvW& c, Act x, int n, int stride, Act* out) {
    ConvSpec s;
    s.a0 = x.p;
    s.c0 = x.C;
    s.lda0 = x.C;
    s.n_img = n;
    s.in_h = x.H;
    s.in_w = x.W;
    s.stride = stride;
    s.taps = 9;
    s.w = H(c.w);
    s.ldw = 9 * c.cin;
    s.ncols = c.cout;
    s.bias = F(c.b);
    const int Ho = stride == 1 ? x.H : (x.H - 1) / 2 + 1, Wo = stride == 1 ? x.W : (x.W - 1) / 2 + 1;
    __half* o;
    LR_TRY(acquire_h(static_cast<size_t>(n) * Ho * Wo * c.cout, &o));
    s.out = o;
    s.ld_out = c.cout;
    Act oact{o, c.cout, Ho, Wo};
    LR_TRY(want_stats(&s));
    ConvOp* op;
    LR_TRY(add_conv_step(s, &op));
    finish_stats(s, op, &oact);
    *out = oact;
    return 0;
  }

  // Upsample.forward (openaimodel.py:108-116; model.py:62-66): nearest x2 then conv3x3. Large images run the FOLDED
  // form (four 2x2-tap convs on the low-resolution input, each writing one output phase through a strided TMA store:
  // 4/9 of the FLOPs, no 4x tensor); below kUpFoldMinPixels low-resolution pixels PER IMAGE four quarter-size launches
  // cannot fill the GPU and the upsampled tensor is materialised instead. The choice must not depend on the batch size:
  // the folded weights round differently, and results are batch / world-size invariant by contract.
  static constexpr int kUpFoldMinPixels = 512;
  int plan_upsample(const ConvW& c, Act x, int n, Act* out) {
    const int hh = x.H, ww = x.W, cc = x.C;
    const bool fold = up_fold && c.is_up && hh * ww >= kUpFoldMinPixels && c.cout % 32 == 0;
    if (!fold) {
      __half* up;
      LR_TRY(acquire_h(static_cast<size_t>(n) * 4 * hh * ww * cc, &up));
      const __half* src = x.p;
      push([=](cudaStream_t st) { return launch_upsample2x(src, n, hh, ww, cc, up, st); });
      Act u{up, cc, 2 * hh, 2 * ww};
      LR_TRY(plan_conv3(c, u, n, 1, out));
      release(up);
      return 0;
    }
    __half* o;
    const size_t co = c.cout;
    LR_TRY(acquire_h(static_cast<size_t>(n) * 4 * hh * ww * co, &o));
    for (int phase = 0; phase < 4; ++phase) {
      const int py = phase >> 1, px = phase & 1;
      ConvSpec s;
      s.a0 = x.p;
      s.c0 = cc;
      s.lda0 = cc;
      s.n_img = n;
      s.in_h = hh;
      s.in_w = ww;
      s.taps = 4;
      s.up_oy = py - 1;
      s.up_ox = px - 1;
      s.w = H(c.wfold) + static_cast<size_t>(phase) * co * 4 * cc;
      s.ldw = 4 * cc;
      s.ncols = c.cout;
      s.bias = F(c.b);
      s.out = o + (static_cast<size_t>(py) * 2 * ww + px) * co;
      s.ld_out = c.cout;
      s.out_sx = 2 * co;
      s.out_sy = 2 * (2 * static_cast<size_t>(ww)) * co;
      s.out_sn = static_cast<size_t>(4) * hh * ww * co;
      LR_TRY(add_conv_step(s));
    }
    *out = Act{o, c.cout, 2 * hh, 2 * ww};
    return 0;
  }

  int plan_block(const std::vector<Node>& blk, Act h, Act skip, int n, Act* out, bool release_input) {
    Act cur = h;
    bool cur_owned = false;  // whether `cur` is a temporary we may release
    for (size_t i = 0; i < blk.size(); ++i) {
      Act nxt;
      const Node& nd = blk[i];
      if (nd.kind == N_RES) {
        next_is_fused_st = i + 1 < blk.size() && blk[i + 1].kind == N_ST && st_fused(cur.H * cur.W);
        LR_TRY(plan_res(res[nd.idx], cur, (i == 0) ? skip : Act{}, n, &nxt));
        next_is_fused_st = false;
      } else if (nd.kind == N_ST) {
        LR_TRY(plan_st(sts[nd.idx], cur, n, &nxt));
      } else if (nd.kind == N_DOWN) {
        LR_TRY(plan_conv3(convs[nd.idx], cur, n, 2, &nxt));
      } else if (nd.kind == N_UP) {
        LR_TRY(plan_upsample(convs[nd.idx], cur, n, &nxt));
      } else {
        LR_CHECK(false, "unexpected node");
      }
      if (cur_owned || (i == 0 && release_input)) release(cur.p);
      if (i == 0 && skip.p != nullptr) release(skip.p);
      cur = nxt;
      cur_owned = true;
    }
    *out = cur;
    return 0;
  }

  // ---- NVSUnetModel(use_sep=True): separator column around every non-resampling block (NVS_ldm.py:57-97) ----
  int plan_sep_insert(Act x, const float* sep, int n, Act* out) {
    __half* o;
    LR_TRY(acquire_h(static_cast<size_t>(n) * x.H * (x.W + 1) * x.C, &o));
    const __half* src = x.p;
    const int hh = x.H, ww = x.W, cc = x.C;
    push([=](cudaStream_t st) { return launch_sep_insert(src, sep, n, hh, ww, cc, o, st); });
    *out = Act{o, x.C, x.H, x.W + 1};
    return 0;
  }
  int plan_sep_remove(Act x, int n, Act* out) {
    __half* o;
    LR_TRY(acquire_h(static_cast<size_t>(n) * x.H * (x.W - 1) * x.C, &o));
    const __half* src = x.p;
    const int hh = x.H, w1 = x.W, cc = x.C;
    push([=](cudaStream_t st) { return launch_sep_remove(src, n, hh, w1, cc, o, st); });
    *out = Act{o, x.C, x.H, x.W - 1};
    return 0;
  }
  // plan_block with the separator inserted before and removed after when the block does not end in a resampling op
  int plan_block_sep(const std::vector<Node>& blk, Act h, Act skip, int n, Act* out, bool release_input) {
    const bool sep_blk = cfg.use_sep && blk.back().kind != N_DOWN && blk.back().kind != N_UP;
    if (!sep_blk) return plan_block(blk, h, skip, n, out, release_input);
    const int ctot = h.C + skip.C;
    auto it = sep_w.find(ctot);
    LR_CHECK(it != sep_w.end(), "use_sep: no separator token for " + std::to_string(ctot) + " channels");
    Act hs_, ss_{};
    LR_TRY(plan_sep_insert(h, F(it->second), n, &hs_));
    if (skip.p != nullptr) LR_TRY(plan_sep_insert(skip, F(it->second) + h.C, n, &ss_));
    if (release_input) release(h.p);
    if (skip.p != nullptr) release(skip.p);
    Act t;
    LR_TRY(plan_block(blk, hs_, ss_, n, &t, true));  // releases the two widened copies
    LR_TRY(plan_sep_remove(t, n, out));
    release(t.p);
    return 0;
  }

  int set_c_input(const float* c_input, int n, int C, int Hh, int Wc, int Wh, cudaStream_t st) {
    if (c_input == nullptr) {
      if (cin_on) { cin_on = false; pn = 0; }
      return 0;
    }
    LR_CHECK(C == cfg.model_channels, "c_input must have model_channels channels (it is added to the input conv's output)");
    LR_CHECK(Wc == Wh || Wc == Wh - Wh / 2, "c_input width must equal the feature width or its right part (NVS_ldm.py:65-68)");
    if (cin_h == nullptr || cin_n != n || cin_H != Hh || cin_Wh != Wh) {
      if (cin_h) cudaFree(cin_h);
      cin_h = nullptr;
      cin_n = cin_H = cin_Wh = 0;
      pn = 0;
      ++plan_generation;
      LR_CUDA(cudaMalloc(&cin_h, static_cast<size_t>(n) * Hh * Wh * C * sizeof(__half)));
      cin_n = n; cin_H = Hh; cin_Wh = Wh;
    }
    LR_TRY(launch_cinput_to_nhwc(c_input, n, C, Hh, Wc, Wc == Wh ? 0 : Wh / 2, Wh, cin_h, st));
    if (!cin_on) { cin_on = true; pn = 0; }
    return 0;
  }

  int build_kv_cache(int n, int L) {
    if (ctx_n == n && ctx_L == L && ctx_h != nullptr) return 0;
    // The old buffers go away: every attention op of the current plan and every captured step graph points at them.
    // Invalidate both BEFORE freeing (a failed allocation below must not leave a state that looks valid).
    if (ctx_h) cudaFree(ctx_h);
    ctx_h = nullptr;
    for (auto p : kv)
      if (p) cudaFree(p);
    kv.assign(n_kv_slots, nullptr);
    kv_ops.clear();
    ctx_n = ctx_L = 0;
    ctx_valid = false;
    pn = 0;             // attention ops hold kv pointers: force a re-plan
    ++plan_generation;  // DDIMSampler's cached step graphs compare this counter (ddim.py _StepGraph.valid)
    const size_t rows = static_cast<size_t>(n) * L;
    LR_CUDA(cudaMalloc(&ctx_h, rows * cfg.context_dim * sizeof(__half)));
    for (int i = 0; i < n_kv_slots; ++i) LR_CUDA(cudaMalloc(&kv[i], rows * 2 * kv_C[i] * sizeof(__half)));
    // one GEMM per transformer block: [n*L, ctx] x [2C, ctx]^T
    for (const STW& s : sts) {
      for (const TBlockW& b : s.blocks) {
        ConvSpec cs;
        cs.a0 = ctx_h;
        cs.c0 = cfg.context_dim;
        cs.lda0 = cfg.context_dim;
        cs.n_img = 1;
        cs.in_h = 1;
        cs.in_w = static_cast<int>(rows);
        cs.taps = 1;
        cs.w = H(b.kv2_w);
        cs.ldw = cfg.context_dim;
        cs.ncols = 2 * s.C;
        cs.out = kv[b.kv_slot];
        cs.ld_out = 2 * s.C;
        auto op = std::make_unique<ConvOp>();
        LR_TRY(build_conv_op(op.get(), cs));
        kv_ops.push_back(std::move(op));
      }
    }
    ctx_n = n;  // committed only after every allocation succeeded
    ctx_L = L;
    return 0;
  }

  int set_context(const float* ctx, int n, int L, cudaStream_t st) {
    LR_TRY(ensure_arenas());
    LR_TRY(build_kv_cache(n, L));
    LR_TRY(launch_cast_f32_f16(ctx, static_cast<size_t>(n) * L * cfg.context_dim, ctx_h, st));
    for (auto& op : kv_ops) LR_TRY(launch_conv_op(*op, st));
    ctx_valid = true;
    return 0;
  }

  int build_plan(int n, int Hh, int Ww, int shared = 0) {
    if (pn == n && ph == Hh && pw == Ww && pshared == shared && pcin == (cin_on ? 1 : 0)) return 0;
    LR_CHECK(!shared || n % 2 == 0, "CFG-pair forward needs an even UNet batch");
    LR_CHECK(!(shared && (cfg.use_sep || cin_on)), "CFG-pair sharing is not implemented with use_sep / c_input");
    const int Wh = Ww + (cfg.use_sep ? 1 : 0);  // width seen by the input conv (separator column inserted)
    LR_CHECK(!cin_on || (cin_n == n && cin_H == Hh && cin_Wh == Wh), "c_input was staged for another batch / latent shape");
    const int ns = shared ? n / 2 : n;  // images whose x / t are distinct
    steps.clear();
    conv_ops.clear();
    attn_ops.clear();
    pool.clear();
    stats_of.clear();
    flops = 0;
    pn = 0;
    const int mc = cfg.model_channels, temb = 4 * mc;
    {
      gn_sites_cap = 2 * static_cast<int>(res.size()) + static_cast<int>(sts.size()) + 1;
      gn_sites_planned = 0;
      gn_site_bytes = groupnorm_scratch_bytes(n, 32, Hh * Wh);  // sized for the largest map
      const size_t bytes = gn_site_bytes * gn_sites_cap;
      void* p;
      LR_TRY(pool.acquire(bytes, &p));
      gn_stats = static_cast<unsigned char*>(p);
      unsigned char* gs = gn_stats;
      const size_t pitch = gn_site_bytes;
      const int sites = gn_sites_cap;
      push([=](cudaStream_t st) {
        // only the grid-barrier counters at the head of every site need to be zero (launch_groupnorm)
        cudaError_t e = cudaMemset2DAsync(gs, pitch, 0, 16, sites, st);
        if (e != cudaSuccess) {
          set_error(std::string("cudaMemsetAsync(gn stats): ") + cudaGetErrorString(e));
          return 1;
        }
        return 0;
      });
    }
    // --- timestep path (openaimodel.py:768-769 + every ResBlock's emb_layers, :263): 4 launches per forward ---
    float *tsin, *e1, *emb;
    LR_TRY(acquire_f(static_cast<size_t>(n) * mc, &tsin));
    LR_TRY(acquire_f(static_cast<size_t>(n) * temb, &e1));
    LR_TRY(acquire_f(static_cast<size_t>(n) * temb, &emb));
    LR_TRY(acquire_f(static_cast<size_t>(n) * emb_total, &emb_all));
    push([=](cudaStream_t st) {
      return launch_timestep_embedding(reinterpret_cast<const long long*>(this->in_t), ns, n, mc, tsin, st);
    });
    {
      const __half* w0 = H(te0_w);
      const float* b0 = F(te0_b);
      const __half* w2 = H(te2_w);
      const float* b2 = F(te2_b);
      const __half* wa = H(emb_w_base);
      const float* ba = F(emb_b_base);
      float* ea = emb_all;
      const int et = emb_total;
      push([=](cudaStream_t st) { return launch_small_linear(tsin, mc, n, mc, w0, b0, temb, 0, 1, e1, temb, st); });
      push([=](cudaStream_t st) { return launch_small_linear(e1, temb, n, temb, w2, b2, temb, 0, 0, emb, temb, st); });
      push([=](cudaStream_t st) { return launch_small_linear(emb, temb, n, temb, wa, ba, et, 1, 0, ea, et, st); });
    }
    // --- input conv: im2col of the NCHW fp32 boundary tensor, then a GEMM ---
    const size_t M0 = static_cast<size_t>(n) * Hh * Ww;
    const size_t M0s = static_cast<size_t>(ns) * Hh * Ww;  // rows computed by the input conv (x is [ns, ...])
    const size_t M0h = static_cast<size_t>(n) * Hh * Wh, M0hs = static_cast<size_t>(ns) * Hh * Wh;  // with the separator
    __half* col;
    LR_TRY(acquire_h(M0hs * kpad_in, &col));
    {
      const int cin = cfg.in_channels, kp = kpad_in;
      if (cfg.use_sep) {
        float* xsep;
        LR_TRY(acquire_f(static_cast<size_t>(ns) * cin * Hh * Wh, &xsep));
        auto it = sep_w.find(cin);
        LR_CHECK(it != sep_w.end(), "use_sep: no separator token for the input channels");
        const float* sp = F(it->second);
        push([=](cudaStream_t st) { return launch_sep_insert_nchw_f32(this->in_x, sp, ns, cin, Hh, Ww, xsep, st); });
        push([=](cudaStream_t st) { return launch_im2col_nchw_f32(xsep, ns, cin, Hh, Wh, kp, col, st); });
        release(xsep);
      } else {
        push(
            [=](cudaStream_t st) { return launch_im2col_nchw_f32(this->in_x, ns, cin, Hh, Ww, kp, col, st); });
      }
    }
    Act h;
    {
      __half* o;
      LR_TRY(acquire_h(M0h * mc, &o));
      const ConvW& c = convs[conv_in_idx];
      // c_input (NVS_ldm.py:64-68) is added to the input conv's output BEFORE the separator is removed
      ConvSpec cs;
      cs.a0 = col;
      cs.c0 = kpad_in;
      cs.lda0 = kpad_in;
      cs.n_img = 1;
      cs.in_h = 1;
      cs.in_w = static_cast<int>(M0hs);
      cs.taps = 1;
      cs.w = H(c.w);
      cs.ldw = kpad_in;
      cs.ncols = mc;
      cs.bias = F(c.b);
      cs.residual = cin_on ? cin_h : nullptr;
      cs.ld_res = mc;
      cs.out = o;
      cs.ld_out = mc;
      cs.stats_rows_per_img = Hh * Wh;
      h = Act{o, mc, Hh, Wh};
      if (!cfg.use_sep) LR_TRY(want_stats(&cs, n / ns));  // (with use_sep the separator column is cut out again below)
      ConvOp* op;
      LR_TRY(add_conv_step(cs, &op));
      finish_stats(cs, op, &h);
      if (cfg.use_sep) {
        Act t;
        LR_TRY(plan_sep_remove(h, n, &t));
        release(h.p);
        h = t;
      }
    }
    release(col);
    std::vector<Act> hs{h};
    for (size_t i = 1; i < input_blocks.size(); ++i) {
      Act o;
      if (shared && i == 1) {
        // shared CFG prefix: ResBlock (+ the first self-attention) of input_blocks.1 on the first half only
        const bool has_st = input_blocks[1].size() > 1 && input_blocks[1][1].kind == N_ST;
        shared_ns = ns;
        if (has_st) {
          LR_TRY(plan_block(input_blocks[1], h, Act{}, n, &o, false));  // plan_st leaves the prefix after attn1
          shared_ns = 0;
        } else {
          LR_TRY(plan_block(input_blocks[1], h, Act{}, n, &o, false));
          shared_ns = 0;
          add_dup_half(o.p, M0s * o.C);
          add_dup_stats(o, ns);
        }
        add_dup_half(hs[0].p, M0s * mc);  // conv_in output: also a skip connection for both halves
        add_dup_stats(hs[0], ns);
      } else {
        LR_TRY(plan_block_sep(input_blocks[i], h, Act{}, n, &o, false));
      }
      hs.push_back(o);
      h = o;
    }
    {
      Act o;
      LR_TRY(plan_block_sep(middle, h, Act{}, n, &o, false));
      h = o;
    }
    for (size_t i = 0; i < output_blocks.size(); ++i) {
      Act skip = hs.back();
      hs.pop_back();
      LR_CHECK(skip.H == h.H && skip.W == h.W, "skip connection spatial mismatch (H, W must be divisible by 2^levels)");
      Act o;
      // h is a temporary except right after the middle block when it still aliases nothing in hs
      LR_TRY(plan_block_sep(output_blocks[i], h, skip, n, &o, true));
      h = o;
    }
    // --- head: GroupNorm32 -> SiLU -> conv3x3 (openaimodel.py:727-731,787) ---
    {
      __half *hn = nullptr, *y;
      const ConvW& c = convs[head_idx];
      LR_TRY(acquire_h(M0 * c.cout, &y));
      ConvSpec s;
      s.n_img = n;
      s.in_h = Hh;
      s.in_w = Ww;
      s.taps = 9;
      float *sc = nullptr, *sh = nullptr;
      int err = 0;
      const bool fused = gn_fuse_conv && conv_is_halo(s) &&
                         plan_gn_coef(h, Act{}, n, 1e-5f, F(head_gn_g), F(head_gn_b), &sc, &sh, &err);
      LR_TRY(err);
      if (fused) {
        s.a0 = h.p;
        s.xf_scale = sc;
        s.xf_shift = sh;
        s.xf_silu = 1;
      } else {
        LR_TRY(acquire_h(M0 * mc, &hn));
        LR_TRY(add_gn_auto(h, Act{}, n, 1e-5f, F(head_gn_g), F(head_gn_b), 1, hn));
        s.a0 = hn;
      }
      s.c0 = mc;
      s.lda0 = mc;
      s.n_img = n;
      s.in_h = Hh;
      s.in_w = Ww;
      s.taps = 9;
      s.w = H(c.w);
      s.ldw = 9 * mc;
      s.ncols = c.cout;
      s.bias = F(c.b);
      s.out = y;
      s.ld_out = c.cout;
      LR_TRY(add_conv_step(s));
      const int co = c.cout;
      push(
          [=](cudaStream_t st) { return launch_nhwc_to_nchw_f32(y, co, n, co, Hh, Ww, this->out_y, st); });
    }
    pn = n;
    ph = Hh;
    pw = Ww;
    pshared = shared;
    pcin = cin_on ? 1 : 0;
    ++plan_generation;
    return 0;
  }
};

struct lr_unet : lr_engine {};

// ==============================================================================================================
// C ABI
// ==============================================================================================================
// ==============================================================================================================
// First-stage decoder engine: AutoencoderKL.decode = decoder(post_quant_conv(z)) (ldm/models/autoencoder.py:87-90) with
// the Decoder of ldm/modules/diffusionmodules/model.py:547-653, as a static plan of the same kernels the UNet uses:
// ResnetBlock (model.py:82-150) = plan_res without the time-embedding term (GroupNorm eps 1e-6 + swish + tcgen05
// conv3x3, nin_shortcut as a 1x1 GEMM, the skip added in the conv2 epilogue); Upsample (model.py:51-66) = nearest x2
// + conv3x3; AttnBlock (model.py:153-204) = one fused QKV GEMM, then per image the single-head d = C attention as two
// tcgen05 GEMMs around a row-softmax kernel (logits [T, T] fp16 with the 1/sqrt(C) scale applied to the fp32
// accumulator - under torch.autocast the reference materialises the same fp16 logits with torch.bmm), proj_out with the
// residual in its epilogue. The d_head = 64 flash kernel does not apply (d = 512); this runs once per batch, not per step.
// ==============================================================================================================
struct VaeAttnW {
  int C = 0;
  size_t gn_g, gn_b, qkv_w, qkv_b, out_w, out_b;
};
struct VaeLevel {
  std::vector<int> blocks;  // indices into res
  int up_conv = -1;         // index into convs, -1: no upsample
};

struct lr_vae : lr_engine {
  lr_vae_cfg vcfg;
  int top = 0, last = 0, kpad = 0;
  size_t conv_in_w = 0, conv_in_b = 0, pq_w = 0, pq_b = 0, no_g = 0, no_b = 0;
  int mid1 = 0, mid2 = 0, out_conv = 0;
  VaeAttnW attn;
  std::vector<VaeLevel> levels;  // from the lowest resolution up
  const float* in_z = nullptr;
  float* out_img = nullptr;
  float z_scale = 1.0f;
  float pz_scale = 0.f;  // planned

  int add_vres(const std::string& pfx, int cin, int cout) {
    ResW r;
    r.cin = cin;
    r.cout = cout;
    r.has_emb = false;
    r.eps = 1e-6f;
    r.gn1_g = reg_vec(pfx + "norm1.weight", cin);
    r.gn1_b = reg_vec(pfx + "norm1.bias", cin);
    r.conv1_w = reg_conv3(pfx + "conv1.weight", cout, cin, 9 * cin);
    r.conv1_b = reg_vec(pfx + "conv1.bias", cout);
    r.gn2_g = reg_vec(pfx + "norm2.weight", cout);
    r.gn2_b = reg_vec(pfx + "norm2.bias", cout);
    r.conv2_w = reg_conv3(pfx + "conv2.weight", cout, cout, 9 * cout);
    r.conv2_b = reg_vec(pfx + "conv2.bias", cout);
    r.has_skip = cin != cout;
    if (r.has_skip) {
      r.skip_w = reg(pfx + "nin_shortcut.weight", {cout, cin, 1, 1}, K_LINEAR, halloc(static_cast<size_t>(cout) * cin), 0, cin);
      r.skip_b = reg_vec(pfx + "nin_shortcut.bias", cout);
    }
    res.push_back(r);
    return static_cast<int>(res.size()) - 1;
  }

  // registration order = the reference state dict (decoder.*, then post_quant_conv.*)
  int build_graph_vae() {
    const lr_vae_cfg& c = vcfg;
    LR_CHECK(c.num_levels >= 1 && c.num_levels <= 8, "vae: num_levels out of range");
    LR_CHECK(c.z_channels >= 1 && c.z_channels <= 7 && c.embed_dim >= 1 && c.embed_dim <= 8, "vae: z_channels / embed_dim out of range");
    LR_CHECK(c.ch % 32 == 0, "vae: ch must be a multiple of 32 (GroupNorm groups)");
    top = c.ch * c.ch_mult[c.num_levels - 1];
    kpad = ((9 * c.z_channels + 63) / 64) * 64;
    const std::string d = "decoder.";
    conv_in_w = reg_conv3(d + "conv_in.weight", top, c.z_channels, kpad);
    conv_in_b = reg_vec(d + "conv_in.bias", top);
    mid1 = add_vres(d + "mid.block_1.", top, top);
    {
      const std::string a = d + "mid.attn_1.";
      attn.C = top;
      attn.gn_g = reg_vec(a + "norm.weight", top);
      attn.gn_b = reg_vec(a + "norm.bias", top);
      attn.qkv_w = halloc(static_cast<size_t>(3) * top * top);
      attn.qkv_b = falloc(3 * top);
      const char* nm[3] = {"q", "k", "v"};
      for (int i = 0; i < 3; ++i) {
        reg(a + nm[i] + ".weight", {top, top, 1, 1}, K_LINEAR, attn.qkv_w, i * top, top);
        reg(a + nm[i] + ".bias", {top}, K_VEC, attn.qkv_b + static_cast<size_t>(i) * top, 0, 0);
      }
      attn.out_w = reg(a + "proj_out.weight", {top, top, 1, 1}, K_LINEAR, halloc(static_cast<size_t>(top) * top), 0, top);
      attn.out_b = reg_vec(a + "proj_out.bias", top);
    }
    mid2 = add_vres(d + "mid.block_2.", top, top);
    // the ModuleList order is up.0 .. up.N-1 (model.py:606 prepends) while execution runs from level N-1 down
    std::vector<VaeLevel> by_level(c.num_levels);
    std::vector<int> in_ch(c.num_levels);
    {
      int block_in = top;
      for (int lvl = c.num_levels - 1; lvl >= 0; --lvl) {
        in_ch[lvl] = block_in;
        block_in = c.ch * c.ch_mult[lvl];
      }
      last = block_in;
    }
    for (int lvl = 0; lvl < c.num_levels; ++lvl) {
      int block_in = in_ch[lvl];
      const int block_out = c.ch * c.ch_mult[lvl];
      for (int i = 0; i <= c.num_res_blocks; ++i) {
        by_level[lvl].blocks.push_back(add_vres(d + "up." + std::to_string(lvl) + ".block." + std::to_string(i) + ".",
                                                block_in, block_out));
        block_in = block_out;
      }
      if (lvl != 0) {
        by_level[lvl].up_conv = add_conv(d + "up." + std::to_string(lvl) + ".upsample.conv.", block_out, block_out,
                                         9 * block_out);
        mark_up_conv(by_level[lvl].up_conv);
      }
    }
    for (int lvl = c.num_levels - 1; lvl >= 0; --lvl) levels.push_back(by_level[lvl]);
    no_g = reg_vec(d + "norm_out.weight", last);
    no_b = reg_vec(d + "norm_out.bias", last);
    out_conv = add_conv(d + "conv_out.", last, c.out_ch, 9 * last);
    pq_w = reg("post_quant_conv.weight", {c.z_channels, c.embed_dim, 1, 1}, K_F32,
               falloc(static_cast<size_t>(c.z_channels) * c.embed_dim), 0, 0);
    pq_b = reg_vec("post_quant_conv.bias", c.z_channels);
    return 0;
  }

  // AttnBlock (model.py:153-204) on x [n, P, C]
  int plan_attn(Act x, int n, Act* out) {
    const int C = attn.C, P = x.H * x.W, M = n * P;
    LR_CHECK(P % 8 == 0 && C % 8 == 0, "vae attention: H*W and C must be multiples of 8");
    __half *xn, *qkv, *S, *vt, *att, *o;
    LR_TRY(acquire_h(static_cast<size_t>(M) * C, &xn));
    LR_TRY(add_gn_auto(x, Act{}, n, 1e-6f, F(attn.gn_g), F(attn.gn_b), 0, xn));
    LR_TRY(acquire_h(static_cast<size_t>(M) * 3 * C, &qkv));
    LR_TRY(add_linear(xn, M, C, H(attn.qkv_w), 3 * C, F(attn.qkv_b), nullptr, 0, qkv, 3 * C, 0));
    release(xn);
    LR_TRY(acquire_h(static_cast<size_t>(P) * P, &S));
    LR_TRY(acquire_h(static_cast<size_t>(C) * P, &vt));
    LR_TRY(acquire_h(static_cast<size_t>(M) * C, &att));
    for (int b = 0; b < n; ++b) {
      const __half* qb = qkv + static_cast<size_t>(b) * P * 3 * C;
      {  // S = (Q K^T) / sqrt(C): "weights" = the K rows of this image
        ConvSpec cs;
        cs.a0 = qb;
        cs.c0 = C;
        cs.lda0 = 3 * C;
        cs.n_img = 1;
        cs.in_h = 1;
        cs.in_w = P;
        cs.taps = 1;
        cs.w = qb + C;
        cs.ldw = 3 * C;
        cs.ncols = P;
        cs.out = S;
        cs.ld_out = P;
        cs.out_scale = 1.0f / sqrtf(static_cast<float>(C));
        LR_TRY(add_conv_step(cs));
      }
      {
        __half* Sb = S;
        push([=](cudaStream_t st) { return launch_softmax_rows(Sb, P, P, static_cast<size_t>(P), st); }, 4, 0.0,
             "vae softmax rows=" + std::to_string(P));
        const __half* vb = qb + 2 * C;
        push([=](cudaStream_t st) { return launch_transpose_f16(vb, P, C, static_cast<size_t>(3) * C, vt, st); }, 4, 0.0,
             "vae V^T");
      }
      {  // O = P V: "weights" = V^T [C, T]
        ConvSpec cs;
        cs.a0 = S;
        cs.c0 = P;
        cs.lda0 = P;
        cs.n_img = 1;
        cs.in_h = 1;
        cs.in_w = P;
        cs.taps = 1;
        cs.w = vt;
        cs.ldw = P;
        cs.ncols = C;
        cs.out = att + static_cast<size_t>(b) * P * C;
        cs.ld_out = C;
        LR_TRY(add_conv_step(cs));
      }
    }
    release(S);
    release(vt);
    release(qkv);
    LR_TRY(acquire_h(static_cast<size_t>(M) * C, &o));
    LR_TRY(add_linear(att, M, C, H(attn.out_w), C, F(attn.out_b), x.p, C, o, C, 0));
    release(att);
    *out = Act{o, C, x.H, x.W};
    return 0;
  }

  int build_plan_vae(int n, int Hh, int Ww) {
    if (pn == n && ph == Hh && pw == Ww && pz_scale == z_scale) return 0;
    steps.clear();
    conv_ops.clear();
    attn_ops.clear();
    pool.clear();
    stats_of.clear();
    flops = 0;
    pn = 0;
    {
      gn_sites_cap = 2 * static_cast<int>(res.size()) + 2;
      gn_sites_planned = 0;
      gn_site_bytes = groupnorm_scratch_bytes(n, 32, 64 * Hh * Ww);  // sized for the decoded resolution
      const size_t bytes = gn_site_bytes * gn_sites_cap;
      void* p;
      LR_TRY(pool.acquire(bytes, &p));
      gn_stats = static_cast<unsigned char*>(p);
      unsigned char* gs = gn_stats;
      const size_t pitch = gn_site_bytes;
      const int sites = gn_sites_cap;
      push([=](cudaStream_t st) {
        cudaError_t e = cudaMemset2DAsync(gs, pitch, 0, 16, sites, st);
        if (e != cudaSuccess) {
          set_error(std::string("cudaMemsetAsync(gn stats): ") + cudaGetErrorString(e));
          return 1;
        }
        return 0;
      });
    }
    const lr_vae_cfg& c = vcfg;
    const size_t M0 = static_cast<size_t>(n) * Hh * Ww;
    __half* col;
    LR_TRY(acquire_h(M0 * kpad, &col));
    {
      const float zs = z_scale;
      const int e = c.embed_dim, zc = c.z_channels, kp = kpad;
      const float* pw_ = F(pq_w);
      const float* pb_ = F(pq_b);
      push([=](cudaStream_t st) { return launch_vae_in(this->in_z, n, e, zc, Hh, Ww, zs, pw_, pb_, kp, col, st); });
    }
    Act h;
    {
      __half* o;
      LR_TRY(acquire_h(M0 * top, &o));
      LR_TRY(add_linear(col, static_cast<int>(M0), kpad, H(conv_in_w), top, F(conv_in_b), nullptr, 0, o, top, 0));
      h = Act{o, top, Hh, Ww};
    }
    release(col);
    auto step_res = [&](int idx) -> int {
      Act o;
      LR_TRY(plan_res(res[idx], h, Act{}, n, &o));
      release(h.p);
      h = o;
      return 0;
    };
    LR_TRY(step_res(mid1));
    {
      Act o;
      LR_TRY(plan_attn(h, n, &o));
      release(h.p);
      h = o;
    }
    LR_TRY(step_res(mid2));
    for (const VaeLevel& lv : levels) {
      for (int idx : lv.blocks) LR_TRY(step_res(idx));
      if (lv.up_conv >= 0) {
        // Upsample (model.py:62-66): nearest x2, then conv3x3 (folded: plan_upsample)
        Act o;
        LR_TRY(plan_upsample(convs[lv.up_conv], h, n, &o));
        release(h.p);
        h = o;
      }
    }
    {
      // norm_out -> swish -> conv_out (model.py:645-650); out_ch (3) columns, rows padded to 4 halves
      const ConvW& cw = convs[out_conv];
      const size_t Mo = static_cast<size_t>(n) * h.H * h.W;
      __half *hn, *y;
      LR_TRY(acquire_h(Mo * h.C, &hn));
      LR_TRY(add_gn_auto(h, Act{}, n, 1e-6f, F(no_g), F(no_b), 1, hn));
      release(h.p);
      const int ldy = 4;
      LR_CHECK(cw.cout <= ldy, "vae: out_ch > 4 is not supported");
      LR_TRY(acquire_h(Mo * ldy, &y));
      ConvSpec s;
      s.a0 = hn;
      s.c0 = h.C;
      s.lda0 = h.C;
      s.n_img = n;
      s.in_h = h.H;
      s.in_w = h.W;
      s.taps = 9;
      s.w = H(cw.w);
      s.ldw = 9 * h.C;
      s.ncols = cw.cout;
      s.bias = F(cw.b);
      s.out = y;
      s.ld_out = ldy;
      LR_TRY(add_conv_step(s));
      release(hn);
      const int co = cw.cout, oh = h.H, ow = h.W;
      push([=](cudaStream_t st) { return launch_nhwc_to_nchw_f32(y, ldy, n, co, oh, ow, this->out_img, st); });
    }
    pn = n;
    ph = Hh;
    pw = Ww;
    pz_scale = z_scale;
    ++plan_generation;
    return 0;
  }
};


static int engine_set_weight(lr_engine* h, const char* name, const float* data, const int64_t* shape, int ndim,
                             void* stream) {
  LR_CHECK(h && name && data && shape, "set_weight: null argument");
  auto it = h->windex.find(name);
  LR_CHECK(it != h->windex.end(), std::string("set_weight: unknown weight '") + name + "'");
  Weight& w = h->weights[it->second];
  bool ok = (ndim == w.ndim);
  for (int i = 0; ok && i < ndim; ++i) ok = (shape[i] == w.shape[i]);
  // a Linear registered as [O, I] also accepts the 1x1-conv form [O, I, 1, 1] and vice versa
  if (!ok && w.kind == K_LINEAR) {
    ok = (ndim == 2 || ndim == 4) && shape[0] == w.shape[0] && shape[1] == w.shape[1] &&
         (ndim == 2 || (shape[2] == 1 && shape[3] == 1));
  }
  LR_CHECK(ok, std::string("set_weight: shape mismatch for '") + name + "'");
  LR_TRY(h->ensure_arenas());
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int O = static_cast<int>(w.shape[0]);
  switch (w.kind) {
    case K_CONV3:
      LR_TRY(launch_repack_conv(data, O, static_cast<int>(w.shape[1]), w.ld, h->H(w.off), st));
      break;
    case K_LINEAR:
      LR_TRY(launch_repack_linear(data, O, static_cast<int>(w.shape[1]), 0, w.dst_row0, h->H(w.off), st));
      break;
    case K_LINEAR_GEGLU:
      LR_TRY(launch_repack_linear(data, O, static_cast<int>(w.shape[1]), 1, 0, h->H(w.off), st));
      break;
    case K_VEC:
      LR_TRY(launch_repack_bias(data, O, 0, h->F(w.off), st));
      break;
    case K_VEC_GEGLU:
      LR_TRY(launch_repack_bias(data, O, 1, h->F(w.off), st));
      break;
    case K_F32: {
      size_t numel = 1;
      for (int i = 0; i < w.ndim; ++i) numel *= static_cast<size_t>(w.shape[i]);
      LR_CUDA(cudaMemcpyAsync(h->F(w.off), data, numel * sizeof(float), cudaMemcpyDeviceToDevice, st));
      break;
    }
  }
  w.loaded = true;
  h->fold_dirty = true;  // LayerNorm-folded weight copies depend on norm{1,2,3} and their consuming Linears
  h->ctx_valid = false;  // cached K/V depend on attn2.to_k / to_v
  return 0;
}


extern "C" {

int lr_abi_version(void) { return LR_B200_ABI_VERSION; }
const char* lr_last_error(void) { return lr::last_error(); }
long long lr_launch_count(void) { return lr::launches_since_reset(); }
void lr_launch_count_reset(void) { lr::reset_launch_counter(); }
void lr_launch_count_add(long long n) { lr::add_launches(n); }
int lr_debug_read_trace(void* dst, long long bytes, int clear) {
  LR_CHECK(dst != nullptr && bytes > 0, "lr_debug_read_trace: bad argument");
  return lr::debug_read_trace(dst, static_cast<size_t>(bytes), clear);
}

int lr_unet_create(const lr_unet_cfg* cfg, lr_unet** out) {
  LR_CHECK(cfg != nullptr && out != nullptr, "lr_unet_create: null argument");
  auto h = std::make_unique<lr_unet>();
  h->cfg = *cfg;
  if (h->cfg.view_num < 1) h->cfg.view_num = 1;
  LR_CHECK(h->cfg.transformer_depth >= 1, "transformer_depth must be >= 1");
  LR_TRY(h->build_graph());
  *out = h.release();
  return 0;
}
void lr_unet_destroy(lr_unet* h) { delete h; }

static int engine_num_weights(const lr_engine* h) { return h ? static_cast<int>(h->weights.size()) : 0; }
static const char* engine_weight_name(const lr_engine* h, int i) {
  if (!h || i < 0 || i >= static_cast<int>(h->weights.size())) return nullptr;
  return h->weights[i].name.c_str();
}
static int engine_weight_shape(const lr_engine* h, int i, int64_t shape_out[4]) {
  if (!h || i < 0 || i >= static_cast<int>(h->weights.size())) return -1;
  for (int d = 0; d < 4; ++d) shape_out[d] = h->weights[i].shape[d];
  return h->weights[i].ndim;
}
static int engine_missing_weights(const lr_engine* h) {
  int m = 0;
  for (const auto& w : h->weights) m += w.loaded ? 0 : 1;
  return m;
}
int lr_unet_num_weights(const lr_unet* h) { return engine_num_weights(h); }
const char* lr_unet_weight_name(const lr_unet* h, int i) { return engine_weight_name(h, i); }
int lr_unet_weight_shape(const lr_unet* h, int i, int64_t shape_out[4]) { return engine_weight_shape(h, i, shape_out); }
int lr_unet_missing_weights(const lr_unet* h) { return engine_missing_weights(h); }

int lr_unet_set_weight(lr_unet* h, const char* name, const float* data, const int64_t* shape, int ndim, void* stream) {
  return engine_set_weight(h, name, data, shape, ndim, stream);
}

int lr_unet_set_context(lr_unet* h, const float* context, int n, int L, void* stream) {
  LR_CHECK(h && context, "lr_unet_set_context: null argument");
  LR_CHECK(lr_unet_missing_weights(h) == 0, "lr_unet_set_context: weights missing");
  // A captured step graph replays lr_unet_forward* without passing through it: weight-derived state (the LayerNorm-folded
  // copies) is therefore refreshed here as well, the one call the sampler makes before every loop.
  LR_TRY(h->run_folds(static_cast<cudaStream_t>(stream)));
  return h->set_context(context, n, L, static_cast<cudaStream_t>(stream));
}

int lr_unet_set_c_input(lr_unet* h, const float* c_input, int n, int C, int H, int Wc, int W, void* stream) {
  LR_CHECK(h != nullptr, "lr_unet_set_c_input: null handle");
  LR_CHECK(c_input == nullptr || (n > 0 && C > 0 && H > 0 && Wc > 0 && W > 0), "lr_unet_set_c_input: empty input");
  // feature width at the point of the addition: the input conv runs on the separator-widened canvas when use_sep
  const int Wh = W + (h->cfg.use_sep ? 1 : 0);
  return h->set_c_input(c_input, n, C, H, Wc, Wh, static_cast<cudaStream_t>(stream));
}

int lr_unet_forward(lr_unet* h, const float* x, const int64_t* timesteps, const float* context, int L, float* out,
                    int n, int H, int W, void* stream) {
  LR_CHECK(h && x && timesteps && out, "lr_unet_forward: null argument");
  LR_CHECK(n > 0 && H > 0 && W > 0, "lr_unet_forward: empty input");
  const int missing = lr_unet_missing_weights(h);
  LR_CHECK(missing == 0, "lr_unet_forward: " + std::to_string(missing) + " weights not uploaded");
  const int div = 1 << (h->cfg.num_levels - 1);
  LR_CHECK(H % div == 0 && W % div == 0, "lr_unet_forward: H and W must be divisible by 2^(levels-1)");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (context != nullptr) {
    LR_TRY(h->set_context(context, n, L, st));
  } else {
    LR_CHECK(h->ctx_valid && h->ctx_n == n, "lr_unet_forward: no cached context for this batch size");
  }
  LR_TRY(h->run_folds(st));
  LR_TRY(h->build_plan(n, H, W));
  h->in_x = x;
  h->in_t = timesteps;
  h->out_y = out;
  if (!h->profiling) {
    for (auto& s : h->steps) LR_TRY(s.fn(st));
    return 0;
  }
  while (h->prof_events.size() < h->steps.size() + 1) {
    cudaEvent_t e;
    LR_CUDA(cudaEventCreate(&e));
    h->prof_events.push_back(e);
  }
  LR_CUDA(cudaEventRecord(h->prof_events[0], st));
  for (size_t i = 0; i < h->steps.size(); ++i) {
    LR_TRY(h->steps[i].fn(st));
    LR_CUDA(cudaEventRecord(h->prof_events[i + 1], st));
  }
  return 0;
}

int lr_unet_set_profiling(lr_unet* h, int enable) {
  LR_CHECK(h != nullptr, "lr_unet_set_profiling: null handle");
  h->profiling = enable != 0;
  return 0;
}

int lr_unet_read_profile(lr_unet* h, double ms_by_class[5], double flops_by_class[5], int steps_by_class[5]) {
  LR_CHECK(h != nullptr, "lr_unet_read_profile: null handle");
  LR_CHECK(h->prof_events.size() >= h->steps.size() + 1 && !h->steps.empty(),
           "lr_unet_read_profile: no profiled forward has run");
  LR_CUDA(cudaEventSynchronize(h->prof_events[h->steps.size()]));
  for (int c = 0; c < 5; ++c) {
    ms_by_class[c] = 0.0;
    flops_by_class[c] = 0.0;
    steps_by_class[c] = 0;
  }
  for (size_t i = 0; i < h->steps.size(); ++i) {
    float ms = 0.f;
    LR_CUDA(cudaEventElapsedTime(&ms, h->prof_events[i], h->prof_events[i + 1]));
    const int c = h->steps[i].cls;
    ms_by_class[c] += ms;
    flops_by_class[c] += h->steps[i].flops;
    steps_by_class[c] += 1;
  }
  return 0;
}

int lr_unet_num_steps(const lr_unet* h) { return h ? static_cast<int>(h->steps.size()) : 0; }

int lr_unet_step_info(lr_unet* h, int i, double* ms, double* flops, int* cls, char* desc, int desc_len) {
  LR_CHECK(h != nullptr && i >= 0 && i < static_cast<int>(h->steps.size()), "lr_unet_step_info: bad index");
  *ms = -1.0;
  if (h->prof_events.size() >= h->steps.size() + 1) {
    float t = 0.f;
    if (cudaEventElapsedTime(&t, h->prof_events[i], h->prof_events[i + 1]) == cudaSuccess) *ms = t;
  }
  *flops = h->steps[i].flops;
  *cls = h->steps[i].cls;
  if (desc && desc_len > 0) {
    strncpy(desc, h->steps[i].desc.c_str(), desc_len - 1);
    desc[desc_len - 1] = 0;
  }
  return 0;
}

int lr_unet_forward_cfg_pair(lr_unet* h, const float* x, const int64_t* timesteps, float* out, int n_canvas, int H,
                             int W, void* stream) {
  LR_CHECK(h && x && timesteps && out, "lr_unet_forward_cfg_pair: null argument");
  LR_CHECK(n_canvas > 0 && H > 0 && W > 0, "lr_unet_forward_cfg_pair: empty input");
  const int missing = lr_unet_missing_weights(h);
  LR_CHECK(missing == 0, "lr_unet_forward_cfg_pair: " + std::to_string(missing) + " weights not uploaded");
  const int div = 1 << (h->cfg.num_levels - 1);
  LR_CHECK(H % div == 0 && W % div == 0, "lr_unet_forward_cfg_pair: H and W must be divisible by 2^(levels-1)");
  LR_CHECK(h->ctx_valid && h->ctx_n == 2 * n_canvas,
           "lr_unet_forward_cfg_pair: lr_unet_set_context must have cached 2*n_canvas contexts (uncond first)");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  LR_TRY(h->run_folds(st));
  LR_TRY(h->build_plan(2 * n_canvas, H, W, 1));
  h->in_x = x;
  h->in_t = timesteps;
  h->out_y = out;
  for (auto& s : h->steps) LR_TRY(s.fn(st));
  return 0;
}

long long lr_unet_plan_generation(const lr_unet* h) { return h ? h->plan_generation : -1; }

int lr_ddim_update_dev(const float* x, const float* eps_uncond, const float* eps_cond, const float* noise,
                       const float* coef, float temperature, int64_t numel, float* x_prev, float* pred_x0,
                       void* stream) {
  LR_CHECK(x && eps_uncond && coef && x_prev && pred_x0 && numel >= 0, "lr_ddim_update_dev: bad argument");
  if (numel == 0) return 0;
  return launch_ddim_update_dev(x, eps_uncond, eps_cond, noise, coef, temperature, static_cast<size_t>(numel), x_prev,
                                pred_x0, static_cast<cudaStream_t>(stream));
}

double lr_unet_last_flops(const lr_unet* h) { return h ? h->flops : 0.0; }
long long lr_unet_device_bytes(const lr_unet* h) {
  return h ? static_cast<long long>(h->persistent_bytes + h->pool.total) : 0;
}

int lr_ddim_update(const float* x, const float* eps_uncond, const float* eps_cond, const float* noise, float cfg_scale,
                   float a_t, float a_prev, float sigma_t, float sqrt_one_minus_at, float temperature, int64_t numel,
                   float* x_prev, float* pred_x0, void* stream) {
  LR_CHECK(x && eps_uncond && x_prev && pred_x0 && numel >= 0, "lr_ddim_update: bad argument");
  if (numel == 0) return 0;
  return launch_ddim_update(x, eps_uncond, eps_cond, noise, cfg_scale, a_t, a_prev, sigma_t, sqrt_one_minus_at,
                            temperature, static_cast<size_t>(numel), x_prev, pred_x0,
                            static_cast<cudaStream_t>(stream));
}

// ---- first-stage decoder -------------------------------------------------------------------------------------
int lr_vae_create(const lr_vae_cfg* cfg, lr_vae** out) {
  LR_CHECK(cfg != nullptr && out != nullptr, "lr_vae_create: null argument");
  auto h = std::make_unique<lr_vae>();
  h->vcfg = *cfg;
  memset(&h->cfg, 0, sizeof(h->cfg));
  h->cfg.view_num = 1;
  LR_TRY(h->build_graph_vae());
  *out = h.release();
  return 0;
}
void lr_vae_destroy(lr_vae* h) { delete h; }
int lr_vae_num_weights(const lr_vae* h) { return engine_num_weights(h); }
const char* lr_vae_weight_name(const lr_vae* h, int i) { return engine_weight_name(h, i); }
int lr_vae_weight_shape(const lr_vae* h, int i, int64_t shape_out[4]) { return engine_weight_shape(h, i, shape_out); }
int lr_vae_missing_weights(const lr_vae* h) { return engine_missing_weights(h); }
int lr_vae_set_weight(lr_vae* h, const char* name, const float* data, const int64_t* shape, int ndim, void* stream) {
  return engine_set_weight(h, name, data, shape, ndim, stream);
}
int lr_vae_decode(lr_vae* h, const float* z, float z_scale, float* out, int n, int H, int W, void* stream) {
  LR_CHECK(h && z && out, "lr_vae_decode: null argument");
  LR_CHECK(n > 0 && H > 0 && W > 0, "lr_vae_decode: empty input");
  const int missing = engine_missing_weights(h);
  LR_CHECK(missing == 0, "lr_vae_decode: " + std::to_string(missing) + " weights not uploaded");
  LR_TRY(h->ensure_arenas());
  LR_TRY(h->run_folds(static_cast<cudaStream_t>(stream)));
  h->z_scale = z_scale;
  LR_TRY(h->build_plan_vae(n, H, W));
  h->in_z = z;
  h->out_img = out;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  for (auto& s : h->steps) LR_TRY(s.fn(st));
  return 0;
}
double lr_vae_last_flops(const lr_vae* h) { return h ? h->flops : 0.0; }
long long lr_vae_device_bytes(const lr_vae* h) {
  return h ? static_cast<long long>(h->persistent_bytes + h->pool.total) : 0;
}
int lr_vae_num_steps(const lr_vae* h) { return h ? static_cast<int>(h->steps.size()) : 0; }

// ---- op level ----------------------------------------------------------------------------------------------
int lr_linear_f16(const void* a, int lda, int M, int K, const void* w, int ldw, int n_cols, const float* bias,
                  const void* residual, int ld_res, void* out, int ld_out, int geglu, int force_block_n, void* stream) {
  if (M == 0) return 0;
  ConvSpec s;
  s.a0 = static_cast<const __half*>(a);
  s.c0 = K;
  s.lda0 = lda;
  s.n_img = 1;
  s.in_h = 1;
  s.in_w = M;
  s.taps = 1;
  s.w = static_cast<const __half*>(w);
  s.ldw = ldw;
  s.ncols = geglu ? 2 * n_cols : n_cols;
  s.bias = bias;
  s.residual = static_cast<const __half*>(residual);
  s.ld_res = ld_res;
  s.out = static_cast<__half*>(out);
  s.ld_out = ld_out;
  s.geglu = geglu;
  s.force_block_n = force_block_n % 1000;
  s.force_cg = force_block_n / 1000;
  ConvOp op;
  LR_TRY(build_conv_op(&op, s));
  return launch_conv_op(op, static_cast<cudaStream_t>(stream));
}

int lr_conv3x3_f16(const void* x0, int c0, const void* x1, int c1, int n, int h, int w, int stride, const void* wt,
                   int cout, const float* bias, const float* bias_img, const void* residual, void* out,
                   int force_block_n, void* stream) {
  if (n == 0) return 0;
  ConvSpec s;
  s.a0 = static_cast<const __half*>(x0);
  s.c0 = c0;
  s.lda0 = c0;
  s.a1 = static_cast<const __half*>(x1);
  s.c1 = x1 ? c1 : 0;
  s.lda1 = c1;
  s.n_img = n;
  s.in_h = h;
  s.in_w = w;
  s.stride = stride;
  s.taps = 9;
  s.w = static_cast<const __half*>(wt);
  s.ldw = 9 * (c0 + s.c1);
  s.ncols = cout;
  s.bias = bias;
  s.bias_img = bias_img;
  s.residual = static_cast<const __half*>(residual);
  s.ld_res = cout;
  s.out = static_cast<__half*>(out);
  s.ld_out = cout;
  s.force_block_n = force_block_n % 1000;
  s.force_cg = force_block_n / 1000;
  {
    const int Ho = stride == 1 ? h : (h - 1) / 2 + 1, Wo = stride == 1 ? w : (w - 1) / 2 + 1;
    const size_t m_out = static_cast<size_t>(n) * Ho * Wo;
    if (Ho * Wo <= kSplitKMaxPixels) {
      s.workspace_bytes = 3 * m_out * cout * sizeof(float);
      s.workspace = op_level_workspace(s.workspace_bytes);
      if (s.workspace == nullptr) s.workspace_bytes = 0;
    }
  }
  ConvOp op;
  LR_TRY(build_conv_op(&op, s));
  return launch_conv_op(op, static_cast<cudaStream_t>(stream));
}

// ---- GroupNorm fusion, op level (gemm_tc.cuh header) -------------------------------------------------------------
long long lr_conv_stats_rows(int n, int h, int w, int stride, int taps, int rows_per_img) {
  ConvSpec s;
  s.n_img = n;
  s.in_h = h;
  s.in_w = w;
  s.stride = stride;
  s.taps = taps;
  s.stats_rows_per_img = rows_per_img;
  return static_cast<long long>(conv_stats_rows(s));
}

int lr_gn_conv3x3_f16(const void* x0, int c0, const void* x1, int c1, int n, int h, int w, const float* gn_scale,
                      const float* gn_shift, int silu, const void* wt, int cout, const float* bias, const float* bias_img,
                      const void* residual, void* out, float* stats_out, int* stats_ppi, int force_block_n, void* stream) {
  if (stats_ppi) *stats_ppi = 0;
  if (n == 0) return 0;
  ConvSpec s;
  s.a0 = static_cast<const __half*>(x0);
  s.c0 = c0;
  s.lda0 = c0;
  s.a1 = static_cast<const __half*>(x1);
  s.c1 = x1 ? c1 : 0;
  s.lda1 = c1;
  s.n_img = n;
  s.in_h = h;
  s.in_w = w;
  s.stride = 1;
  s.taps = 9;
  s.w = static_cast<const __half*>(wt);
  s.ldw = 9 * (c0 + s.c1);
  s.ncols = cout;
  s.bias = bias;
  s.bias_img = bias_img;
  s.residual = static_cast<const __half*>(residual);
  s.ld_res = cout;
  s.out = static_cast<__half*>(out);
  s.ld_out = cout;
  s.force_block_n = force_block_n % 1000;
  s.force_cg = force_block_n / 1000;
  s.xf_scale = gn_scale;
  s.xf_shift = gn_shift;
  s.xf_silu = silu;
  s.stats_out = stats_out;
  LR_CHECK(gn_scale == nullptr || conv_is_halo(s),
           "lr_gn_conv3x3_f16: the fused GroupNorm transform needs an image of at least 16 rows x 8 columns (halo mode)");
  ConvOp op;
  LR_TRY(build_conv_op(&op, s));
  if (stats_ppi) *stats_ppi = op.stats_ok ? op.stats_ppi : 0;
  return launch_conv_op(op, static_cast<cudaStream_t>(stream));
}

int lr_gn_linear_f16(const void* a, int M, int K, int rows_per_img, const float* gn_scale, const float* gn_shift,
                     int silu, const void* w, int n_cols, const float* bias, const void* residual, void* out,
                     float* stats_out, int* stats_ppi, int force_block_n, void* stream) {
  if (stats_ppi) *stats_ppi = 0;
  if (M == 0) return 0;
  ConvSpec s;
  s.a0 = static_cast<const __half*>(a);
  s.c0 = K;
  s.lda0 = K;
  s.n_img = 1;
  s.in_h = 1;
  s.in_w = M;
  s.taps = 1;
  s.w = static_cast<const __half*>(w);
  s.ldw = K;
  s.ncols = n_cols;
  s.bias = bias;
  s.residual = static_cast<const __half*>(residual);
  s.ld_res = n_cols;
  s.out = static_cast<__half*>(out);
  s.ld_out = n_cols;
  s.force_block_n = force_block_n % 1000;
  s.force_cg = force_block_n / 1000;
  s.xf_scale = gn_scale;
  s.xf_shift = gn_shift;
  s.xf_silu = silu;
  s.xf_rows_per_img = rows_per_img;
  s.stats_out = stats_out;
  s.stats_rows_per_img = rows_per_img;
  ConvOp op;
  LR_TRY(build_conv_op(&op, s));
  if (stats_ppi) *stats_ppi = op.stats_ok ? op.stats_ppi : 0;
  return launch_conv_op(op, static_cast<cudaStream_t>(stream));
}

int lr_gn_finalize(const float* part0, int ppi0, int c0, const float* part1, int ppi1, int c1, int n, int P, int groups,
                   float eps, const float* gamma, const float* beta, float* scale, float* shift, void* stream) {
  LR_CHECK(part0 && gamma && beta && scale && shift, "lr_gn_finalize: null argument");
  if (n == 0) return 0;
  return launch_gn_finalize(part0, ppi0, c0, part1, part1 ? ppi1 : 0, part1 ? c1 : 0, n, P, groups, eps, gamma, beta, scale,
                            shift, static_cast<cudaStream_t>(stream));
}

int lr_upsample2x_conv3x3_f16(const void* x, int n, int h, int w, int cin, const void* wt, int cout, const float* bias,
                              void* wfold_scratch, void* out, void* stream) {
  LR_CHECK(x && wt && wfold_scratch && out, "lr_upsample2x_conv3x3_f16: null argument");
  LR_CHECK(cout % 32 == 0 && cin % 8 == 0, "lr_upsample2x_conv3x3_f16: cout must be a multiple of 32, cin of 8");
  if (n == 0) return 0;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  __half* wf = static_cast<__half*>(wfold_scratch);
  LR_TRY(launch_upfold_weights(static_cast<const __half*>(wt), cout, cin, wf, st));
  __half* o = static_cast<__half*>(out);
  const size_t co = cout;
  for (int phase = 0; phase < 4; ++phase) {
    const int py = phase >> 1, px = phase & 1;
    ConvSpec s;
    s.a0 = static_cast<const __half*>(x);
    s.c0 = cin;
    s.lda0 = cin;
    s.n_img = n;
    s.in_h = h;
    s.in_w = w;
    s.taps = 4;
    s.up_oy = py - 1;
    s.up_ox = px - 1;
    s.w = wf + static_cast<size_t>(phase) * co * 4 * cin;
    s.ldw = 4 * cin;
    s.ncols = cout;
    s.bias = bias;
    s.out = o + (static_cast<size_t>(py) * 2 * w + px) * co;
    s.ld_out = cout;
    s.out_sx = 2 * co;
    s.out_sy = 2 * (2 * static_cast<size_t>(w)) * co;
    s.out_sn = static_cast<size_t>(4) * h * w * co;
    ConvOp op;
    LR_TRY(build_conv_op(&op, s));
    LR_TRY(launch_conv_op(op, st));
  }
  return 0;
}

int lr_attention_f16(const void* q, int ldq, int q_col0, const void* k, int ldk, int k_col0, const void* v, int ldv,
                     int v_col0, void* out, int ld_out, int batch, int heads, int tq, int tk, float scale,
                     void* stream) {
  if (batch == 0 || tq == 0) return 0;
  AttnSpec s;
  s.q = static_cast<const __half*>(q); s.ldq = ldq; s.q_col0 = q_col0;
  s.k = static_cast<const __half*>(k); s.ldk = ldk; s.k_col0 = k_col0;
  s.v = static_cast<const __half*>(v); s.ldv = ldv; s.v_col0 = v_col0;
  s.out = static_cast<__half*>(out); s.ld_out = ld_out;
  s.batch = batch; s.heads = heads; s.tq = tq; s.tk = tk; s.scale = scale;
  AttnOp op;
  LR_TRY(build_attn_op(&op, s));
  return launch_attn_op(op, static_cast<cudaStream_t>(stream));
}

size_t lr_groupnorm_scratch_bytes(int n, int groups, int P) { return groupnorm_scratch_bytes(n, groups, P); }
int lr_groupnorm_f16(const void* x0, int c0, const void* x1, int c1, int n, int P, int groups, float eps,
                     const float* gamma, const float* beta, int silu, void* out, void* scratch, void* stream) {
  LR_CHECK(x0 && gamma && beta && out && scratch, "lr_groupnorm_f16: null argument");
  if (n == 0 || P == 0) return 0;
  return launch_groupnorm(static_cast<const __half*>(x0), c0, static_cast<const __half*>(x1), x1 ? c1 : 0, n, P, groups,
                          eps, gamma, beta, silu, scratch, 0, static_cast<__half*>(out),
                          static_cast<cudaStream_t>(stream));
}
int lr_layernorm_f16(const void* x, int M, int C, const float* gamma, const float* beta, float eps, void* out,
                     void* stream) {
  LR_CHECK(x && gamma && beta && out, "lr_layernorm_f16: null argument");
  if (M == 0) return 0;
  return launch_layernorm(static_cast<const __half*>(x), M, C, gamma, beta, eps, static_cast<__half*>(out),
                          static_cast<cudaStream_t>(stream));
}
int lr_nchw_f32_to_nhwc_f16(const float* x, int n, int c, int h, int w, void* out, void* stream) {
  if (static_cast<size_t>(n) * c * h * w == 0) return 0;
  return launch_nchw_f32_to_nhwc(x, n, c, h, w, static_cast<__half*>(out), static_cast<cudaStream_t>(stream));
}
int lr_nhwc_f16_to_nchw_f32(const void* x, int ld, int n, int c, int h, int w, float* out, void* stream) {
  if (static_cast<size_t>(n) * c * h * w == 0) return 0;
  return launch_nhwc_to_nchw_f32(static_cast<const __half*>(x), ld, n, c, h, w, out, static_cast<cudaStream_t>(stream));
}
int lr_repack_conv3x3_weight(const float* w_oihw, int cout, int cin, void* out, void* stream) {
  return launch_repack_conv(w_oihw, cout, cin, 9 * cin, static_cast<__half*>(out), static_cast<cudaStream_t>(stream));
}
int lr_repack_linear_weight(const float* w, int n_out, int n_in, int geglu, void* out, void* stream) {
  return launch_repack_linear(w, n_out, n_in, geglu, 0, static_cast<__half*>(out), static_cast<cudaStream_t>(stream));
}

}  // extern "C"
