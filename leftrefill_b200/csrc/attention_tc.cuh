// Fused softmax(Q K^T * scale) V for sm_100a with tcgen05 tensor cores and TMEM accumulators.
// Replaces the einsum/softmax/einsum of CrossAttention.forward (reference ldm/modules/attention.py:165-196), which
// materialises the [b*h, Tq, Tk] fp32 logits (1.34 GB per canvas at T = 8192). Precision islands are kept:
// fp16 operands, fp32 QK^T accumulation (the reference's _ATTN_PRECISION="fp32" path), fp32 softmax statistics.
//
// One CTA = one (batch, head, 256-query block) = two 128-row query tiles that ping-pong on the tensor core. d_head = 64.
//   warp 0 (1 lane)   : TMA — Q0/Q1 once, then a 4-stage ring of {K_j, V_j} 128-token tiles shared by both query tiles
//   warps 1, 2 (1 lane): tcgen05.mma for query tile 0 / 1 — S_t = Q_t K_j^T (128x128 fp32 in TMEM), O_t += P_t V_j
//                       (128x64 fp32 in TMEM) with P_t read from TENSOR MEMORY (fp16 pairs, 64 columns): the softmax
//                       result never touches shared memory. S_t(j+1) is issued as soon as warpgroup t has pulled
//                       S_t(j) into registers, so in steady state a warpgroup never waits for the tensor pipe.
//   warps 4..7, 8..11 : softmax warpgroup for tile 0 / tile 1 — thread <-> query row (tcgen05.ld 32x32b), the whole
//                       128-wide logits row lives in registers, online max with lazy rescaling of the TMEM-resident O
//                       (threshold 2^8), P_t = exp2(.) packed to fp16 pairs and stored with tcgen05.st.
// V is consumed as an MN-major B operand straight from the [token, d] layout the QKV GEMM produces: no transposes.
// TMEM (512 columns): S0 [0,128) S1 [128,256) P0 [256,320) P1 [320,384) O0 [384,448) O1 [448,512).
#pragma once
#include "ptx.cuh"

namespace lr {

struct AttnParams {
  CUtensorMap tmQ, tmK, tmV;  // 3-D [cols, tokens, batch], box (64, 128, 1), 128B swizzle
  int heads, tq, tk, batch;
  int q_col0, k_col0, v_col0;  // column of head 0 inside the Q / K / V tensors
  __half* out;                 // [batch*tq, ld_out], head h -> columns [h*64, h*64+64)
  int ld_out;
  float scale_log2;            // softmax scale * log2(e)
  unsigned long long* trace;   // LR_ATTN_TRACE: block (0,0,0) accumulates clock64() per softmax phase; else null
};

constexpr int kAttnThreads = 384;
constexpr int kAttnTile = 128;                          // rows of one query tile / keys per KV tile
constexpr int kAttnQBlock = 2 * kAttnTile;              // queries per CTA
constexpr int kAttnD = 64;
constexpr int kAttnTileBytes = kAttnTile * kAttnD * 2;  // 16 KB
#ifndef LR_ATTN_STAGES
#define LR_ATTN_STAGES 4
#endif
constexpr int kAttnStages = LR_ATTN_STAGES;
constexpr int kAttnSmemBytes = 2 * kAttnTileBytes /*Q0,Q1*/ + kAttnStages * 2 * kAttnTileBytes /*K,V ring*/ +
                               256 /*barriers*/;
constexpr int kTmemS = 0, kTmemP = 256, kTmemO = 384;  // column bases (S: 128 per tile, P: 64, O: 64)
// Softmax ping-pong (A/B: -DLR_ATTN_PINGPONG=0): the two warpgroups take turns in the exp2 phase (named barriers 2/3),
// so one warpgroup owns the MUFU unit (16 ex2/clk/SM) while the other loads S, finds the row max and stores P. Without
// it both drift into lock step, share the MUFU at half rate each and leave it idle during their common non-exp phases.
#ifndef LR_HI_WARP_ISSUE
#define LR_HI_WARP_ISSUE 0
#endif
#ifndef LR_ATTN_PINGPONG
#define LR_ATTN_PINGPONG 1
#endif
constexpr float kRescaleThreshold = 8.0f;  // log2 domain: P stays <= 256, exact in fp16/fp32 accumulators

template <bool B>
struct BoolTag {
  static constexpr bool value = B;
};

__device__ __forceinline__ float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// exp2 on the FMA pipe (Cody-Waite range reduction + degree-3 minimax polynomial, max relative error 7.6e-5, i.e. below
// half an fp16 ulp of the P value it feeds). -DLR_ATTN_POLY_GROUPS=G computes one exponential in every G groups of four
// here instead of on the MUFU unit (16 ex2/clk/SM, 77 % busy in ncu, FMA pipe 22 %). Measured on B200 with the softmax
// ping-pong in place (round 1, r1l, 8x5 heads x 8192^2): 875 us all-MUFU, 907 us with 25 %, 981 us with 12.5 % on the
// polynomial - the softmax warps are bound by issue slots and latency as much as by MUFU throughput, so the default is 0.
// Other measured dead ends: ex2.approx.f16x2 is split by ptxas into two MUFU ops + PRMT; a ones-column in V to get the
// row sum from the MMA was correct but not faster.
#ifndef LR_ATTN_POLY_GROUPS
#define LR_ATTN_POLY_GROUPS 0
#endif
// Logits (of the 128 per row and KV tile) whose exponential is computed BEFORE a warpgroup takes its ping-pong turn:
// the non-exp phases of a tile are shorter than the other warpgroup's turn, so a strict hand-over leaves the waiting
// warpgroup idle; a head start on its own exponentials fills MUFU bubbles of the owner instead. (Round 1's schedule had
// this by accident - ptxas had hoisted 39 of 128 MUFU ops above the barrier; see pingpong_wait.)
#ifndef LR_ATTN_PRE_EXP
#define LR_ATTN_PRE_EXP 40
#endif
static_assert(LR_ATTN_PRE_EXP % 4 == 0 && LR_ATTN_PRE_EXP >= 0 && LR_ATTN_PRE_EXP <= 128, "LR_ATTN_PRE_EXP");
__device__ __forceinline__ float poly_exp2(float x) {
  x = fmaxf(x, -125.0f);
  const float xr = x + 12582912.0f;        // 1.5 * 2^23: the nearest integer to x lands in the low mantissa bits
  const float f = x - (xr - 12582912.0f);  // [-0.5, 0.5]
  float p = fmaf(0.05517028f, f, 0.2426077f);
  p = fmaf(p, f, 0.6932609f);
  p = fmaf(p, f, 0.9999283f);
  return __int_as_float(__float_as_int(p) + (__float_as_int(xr) << 23));  // p * 2^round(x)
}
// Ping-pong turn taking must bracket the exp2 phase, but named barriers order MEMORY operations only: ptxas moved the
// register-only exp2 work of one instantiation above `bar.sync` (round 2: both warpgroups then shared the MUFU and the
// kernel lost 20 %). The hand-over is therefore tied into the data flow through shared memory: the exponent offset
// passes through a volatile load issued after the barrier, and the row sum through a volatile store issued before the
// arrive.
__device__ __forceinline__ float pingpong_wait(int bar_id, float moff, uint32_t zero_addr) {
  float z;
  asm volatile("bar.sync %1, 256;\n\tld.volatile.shared.f32 %0, [%2];" : "=f"(z) : "r"(bar_id), "r"(zero_addr) : "memory");
  return moff + z;  // z == 0.0f
}
__device__ __forceinline__ void pingpong_pass(int bar_id, float rowsum, uint32_t sink_addr) {
  asm volatile("st.volatile.shared.f32 [%2], %1;\n\tbar.arrive %0, 256;" ::"r"(bar_id), "f"(rowsum), "r"(sink_addr) : "memory");
}
__device__ __forceinline__ float fmax3(float a, float b, float c) {
  float d;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
  return d;
}

// kTrace = true (debug builds of the launch only): one lane per softmax warpgroup of block (0,0,0) accumulates the
// cycles of each phase of tile_body into p.trace[t*8 + phase]:
//   0 wait S ready | 1 tcgen05.ld S | 2 row max / rescale decision | 3 exp2 + pack | 4 wait PV(j-1) (+ O rescale)
//   5 tcgen05.st P + arrive | 6 total | 7 tiles
template <bool kTrace>
__global__ void __launch_bounds__(kAttnThreads, 1) attention_kernel(const __grid_constant__ AttnParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* q_s = smem;                                       // [2] query tiles
  uint8_t* k_s = q_s + 2 * kAttnTileBytes;                   // [stages]
  uint8_t* v_s = k_s + kAttnStages * kAttnTileBytes;         // [stages]
  uint64_t* bars = reinterpret_cast<uint64_t*>(v_s + kAttnStages * kAttnTileBytes);
  uint64_t* q_full = bars;                         // 1
  uint64_t* kv_full = bars + 1;                    // [stages]
  uint64_t* kv_empty = kv_full + kAttnStages;      // [stages]
  uint64_t* s_full = kv_empty + kAttnStages;       // [2] S_t(j) is in TMEM
  uint64_t* s_empty = s_full + 2;                  // [2] warpgroup t holds S_t(j) in registers
  uint64_t* p_full = s_empty + 2;                  // [2] P_t(j) is in TMEM
  uint64_t* pv_done = p_full + 2;                  // [2] O_t += P_t(j) V_j has completed
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(pv_done + 2);

  // roles: 0 = TMA, 1 / 2 = MMA issuer of query tile 0 / 1, 3 = idle, 4..7 / 8..11 = softmax warpgroup 0 / 1.
  // -DLR_HI_WARP_ISSUE=1 moves the issuing roles to the highest hardware warp ids (8..10); measured on B200 (round 1,
  // r1h): 893 us instead of 854 us at 8 x 5 heads x 8192^2, so the issuers stay on the lowest ids.
  const int hw_warp = threadIdx.x >> 5;
#if LR_HI_WARP_ISSUE
  const int warp = (hw_warp + 4) % 12;
#else
  const int warp = hw_warp;
#endif
  const int lane = threadIdx.x & 31;
  const int qb = blockIdx.x, head = blockIdx.y, b = blockIdx.z;
  const int ntiles = (p.tk + kAttnTile - 1) / kAttnTile;
  // the second query tile may lie completely beyond tq (e.g. the 8x16 level has 128 tokens): it is then skipped
  // everywhere (no fully out-of-bounds TMA box, no MMA, no softmax work)
  const int ntq = (qb * kAttnQBlock + kAttnTile < p.tq) ? 2 : 1;
  pdl_launch_dependents();

  if (threadIdx.x == 0) {
    if ((smem_u32(smem) & 1023u) != 0) {
      printf("lr_b200: attention smem base not 1024-aligned\n");
      __trap();
    }
    tma_prefetch_desc(&p.tmQ);
    tma_prefetch_desc(&p.tmK);
    tma_prefetch_desc(&p.tmV);
    mbar_init(q_full, 1);
    for (int i = 0; i < kAttnStages; ++i) {
      mbar_init(&kv_full[i], 1);
      mbar_init(&kv_empty[i], ntq);  // one tcgen05.commit per query tile's issuing warp
    }
    for (int t = 0; t < 2; ++t) {
      mbar_init(&s_full[t], 1);
      mbar_init(&s_empty[t], 128);
      mbar_init(&p_full[t], 128);
      mbar_init(&pv_done[t], 1);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();  // predecessors complete: from here on global memory may be read and written
  if (warp < 4) {
    if (warp == 0) {
      // ------------------------------- TMA producer (whole warp loops, one elected lane issues) ---------------
      if (elect_one()) {
        mbar_arrive_expect_tx(q_full, ntq * kAttnTileBytes);
        tma_load_3d(q_s, &p.tmQ, q_full, p.q_col0 + head * kAttnD, qb * kAttnQBlock, b);
        if (ntq == 2)
          tma_load_3d(q_s + kAttnTileBytes, &p.tmQ, q_full, p.q_col0 + head * kAttnD, qb * kAttnQBlock + kAttnTile, b);
      }
      __syncwarp();
      for (int j = 0; j < ntiles; ++j) {
        const int s = j % kAttnStages;
        const uint32_t ph = (j / kAttnStages) & 1;
        mbar_wait(&kv_empty[s], ph ^ 1);
        if (elect_one()) {
          mbar_arrive_expect_tx(&kv_full[s], 2 * kAttnTileBytes);
          tma_load_3d(k_s + s * kAttnTileBytes, &p.tmK, &kv_full[s], p.k_col0 + head * kAttnD, j * kAttnTile, b);
          tma_load_3d(v_s + s * kAttnTileBytes, &p.tmV, &kv_full[s], p.v_col0 + head * kAttnD, j * kAttnTile, b);
        }
        __syncwarp();
      }
    } else if (warp - 1 < ntq) {
      // ------------------------------- MMA issuers: warp 1 -> query tile 0, warp 2 -> query tile 1 -------------
      // One issuing warp PER query tile, each following its own warpgroup with BLOCKING mbarrier waits in program
      // order (a suspended try_wait wakes ~60 cycles after the arrive). The earlier single-warp event loop polled six
      // barriers with mbarrier.test_wait (~150 cycles each): a warpgroup waited ~740 cycles per KV tile for an S tile
      // whose inputs had been ready for a long time (in-kernel trace, profiles/r1b_trace.log).
      // Whole warp loops, one elected lane issues. Descriptors are a constant high word plus a 32-bit low word
      // (address >> 4) kept in uniform registers: advancing along K is one add.
      const int t = warp - 1;
      const uint32_t idesc_s = umma_idesc_f16(128, kAttnTile, 0);  // S: N = 128 keys, K-major B
      const uint32_t idesc_o = umma_idesc_f16(128, kAttnD, 1);     // O: N = 64 channels, MN-major B (V)
      const uint32_t desc_hi = umma_desc_hi_sw128(1024);
      const uint32_t q_lo = umma_desc_lo(smem_u32(q_s), 16);
      const uint32_t k_lo = umma_desc_lo(smem_u32(k_s), 16);
      const uint32_t v_lo = umma_desc_lo(smem_u32(v_s), 16);
      constexpr uint32_t kTileUnits = kAttnTileBytes >> 4;
      auto issue_s = [&](int stage) {  // S_t = Q_t K^T : 4 k-steps of 16 channels (32 B = 2 units each)
#pragma unroll
        for (int k = 0; k < kAttnD / 16; ++k) {
          umma_f16(tmem_base + kTmemS + t * 128, umma_desc_make(desc_hi, q_lo + t * kTileUnits + 2 * k),
                   umma_desc_make(desc_hi, k_lo + stage * kTileUnits + 2 * k), idesc_s, k != 0 ? 1u : 0u);
        }
        umma_commit(&s_full[t]);
      };
      auto issue_pv = [&](int stage, uint32_t accumulate) {
        // A = P_t from TMEM (8 columns = 16 fp16 per k step); B = 16 token rows of V (16 x 128 B = 128 units), MN-major
#pragma unroll
        for (int k = 0; k < kAttnTile / 16; ++k) {
          umma_f16_ts(tmem_base + kTmemO + t * 64, tmem_base + kTmemP + t * 64 + k * 8,
                      umma_desc_make(desc_hi, v_lo + stage * kTileUnits + k * 128), idesc_o, (accumulate | k) != 0 ? 1u : 0u);
        }
        umma_commit(&pv_done[t]);
        umma_commit(&kv_empty[stage]);  // the stage is free once BOTH tiles' PV MMAs have read it (count = ntq)
      };
      mbar_wait(q_full, 0);
      mbar_wait(&kv_full[0], 0);
      tc_fence_after();
      if (elect_one()) issue_s(0);
      __syncwarp();
      for (int j = 0; j < ntiles; ++j) {
        if (j + 1 < ntiles) {
          // S_t(j+1) runs one tile ahead: it needs K_{j+1} and the warpgroup to have pulled S_t(j) into registers
          const int s1 = (j + 1) % kAttnStages;
          mbar_wait(&kv_full[s1], ((j + 1) / kAttnStages) & 1);
          mbar_wait(&s_empty[t], j & 1);
          tc_fence_after();
          if (elect_one()) issue_s(s1);
          __syncwarp();
        }
        mbar_wait(&p_full[t], j & 1);
        tc_fence_after();
        if (elect_one()) issue_pv(j % kAttnStages, j > 0 ? 1u : 0u);
        __syncwarp();
      }
    }
  } else {
    // ------------------------------- softmax warpgroups ---------------------------
    const int t = (warp - 4) >> 2;  // query tile of this warpgroup
    const int q = hw_warp & 3;      // TMEM lane quarter (hardware warp id % 4)
    const int r = q * 32 + lane;    // row inside the tile
    const uint32_t lane_addr = static_cast<uint32_t>(q * 32) << 16;
    const uint32_t tmem_S = tmem_base + kTmemS + t * 128 + lane_addr;
    const uint32_t tmem_P = tmem_base + kTmemP + t * 64 + lane_addr;
    const uint32_t tmem_O = tmem_base + kTmemO + t * 64 + lane_addr;
    float m_used = -INFINITY;  // max the current O / l are scaled against (raw logit units)
    float l = 0.f;
    const int my_tiles = (t < ntq) ? ntiles : 0;

    unsigned long long tr[6] = {0, 0, 0, 0, 0, 0};
    long long tr_t = 0, tr_begin = 0;
    const bool tracing = kTrace && p.trace != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0;
    if (kTrace) tr_begin = tr_t = clock64();
    auto mark = [&](int phase) {
      if (kTrace) {
        const long long now = clock64();
        tr[phase] += static_cast<unsigned long long>(now - tr_t);
        tr_t = now;
      }
    };
#if LR_ATTN_PINGPONG
    if (ntq == 2 && t == 1) asm volatile("bar.arrive 2, 256;" ::: "memory");  // warpgroup 0 goes first
#endif
    auto tile_body = [&](int j, auto mask_tag) {
      constexpr bool kMask = decltype(mask_tag)::value;
      mbar_wait(&s_full[t], j & 1);
      tc_fence_after();
      mark(0);
      uint32_t s[128];
      {
        uint32_t(&s0)[32] = *reinterpret_cast<uint32_t(*)[32]>(&s[0]);
        uint32_t(&s1)[32] = *reinterpret_cast<uint32_t(*)[32]>(&s[32]);
        uint32_t(&s2)[32] = *reinterpret_cast<uint32_t(*)[32]>(&s[64]);
        uint32_t(&s3)[32] = *reinterpret_cast<uint32_t(*)[32]>(&s[96]);
        tmem_ld32(tmem_S + 0, s0);
        tmem_ld32(tmem_S + 32, s1);
        tmem_ld32(tmem_S + 64, s2);
        tmem_ld32(tmem_S + 96, s3);
        tmem_ld_wait();
      }
      tc_fence_before();
      mbar_arrive(&s_empty[t]);  // S_t may be overwritten by the next QK^T right away
      mark(1);
      if (kMask) {               // only the last KV tile can be partial (keys beyond tk were zero-filled by TMA)
        const int kv_valid = p.tk - j * kAttnTile;
#pragma unroll
        for (int i = 0; i < 128; ++i)
          if (i >= kv_valid) s[i] = 0xff800000u;  // -inf
      }
      float mx[8];
#pragma unroll
      for (int c = 0; c < 8; ++c) mx[c] = fmax3(__uint_as_float(s[c]), __uint_as_float(s[c + 8]), __uint_as_float(s[c + 16]));
#pragma unroll
      for (int i = 24; i < 120; i += 16) {
#pragma unroll
        for (int c = 0; c < 8; ++c) mx[c] = fmax3(mx[c], __uint_as_float(s[i + c]), __uint_as_float(s[i + c + 8]));
      }
#pragma unroll
      for (int c = 0; c < 8; ++c) mx[c] = fmaxf(mx[c], __uint_as_float(s[120 + c]));
      const float mrow = fmaxf(fmax3(mx[0], mx[1], mx[2]), fmaxf(fmax3(mx[3], mx[4], mx[5]), fmaxf(mx[6], mx[7])));
      const float m_new = fmaxf(m_used, mrow);
      const bool need = (m_new - m_used) * p.scale_log2 > kRescaleThreshold;  // true on the first tile (-inf)
      float alpha = 1.0f;
      if (need) {
        alpha = (m_used == -INFINITY) ? 0.f : fast_exp2((m_used - m_new) * p.scale_log2);
        m_used = m_new;
        l *= alpha;
      }
      mark(2);
#if LR_ATTN_PINGPONG
      if (ntq == 2) asm volatile("bar.sync %0, 256;" ::"r"(2 + t) : "memory");  // my turn on the MUFU
#endif
      // P = exp2((s - m_used) * scale_log2) as fp16 pairs, kept in registers until the previous PV has released P_t
      const float moff = m_used * p.scale_log2;
      uint32_t h[64];
      float l0 = 0.f, l1 = 0.f, l2 = 0.f, l3 = 0.f;
#pragma unroll
      for (int i = 0; i < 64; i += 2) {
        const float e0 = fast_exp2(fmaf(__uint_as_float(s[2 * i]), p.scale_log2, -moff));
        const float e1 = fast_exp2(fmaf(__uint_as_float(s[2 * i + 1]), p.scale_log2, -moff));
        const float e2 = fast_exp2(fmaf(__uint_as_float(s[2 * i + 2]), p.scale_log2, -moff));
        const float x3 = fmaf(__uint_as_float(s[2 * i + 3]), p.scale_log2, -moff);
        const float e3 = (LR_ATTN_POLY_GROUPS > 0 && ((i >> 1) % (LR_ATTN_POLY_GROUPS > 0 ? LR_ATTN_POLY_GROUPS : 1)) == 0)
                             ? poly_exp2(x3) : fast_exp2(x3);
        l0 += e0;
        l1 += e1;
        l2 += e2;
        l3 += e3;
        h[i] = pack_half2(e0, e1);
        h[i + 1] = pack_half2(e2, e3);
      }
      l += (l0 + l1) + (l2 + l3);
#if LR_ATTN_PINGPONG
      // hand the MUFU to the other warpgroup (its last turn needs no successor)
      if (ntq == 2 && !(t == 1 && j == my_tiles - 1)) asm volatile("bar.arrive %0, 256;" ::"r"(3 - t) : "memory");
#endif
      mark(3);
      if (j > 0) {
        mbar_wait(&pv_done[t], (j - 1) & 1);  // PV(j-1) finished reading P_t and updating O_t
        tc_fence_after();
        if (__any_sync(0xffffffffu, need)) {
          // rescale the TMEM-resident O row (warp-collective; lanes that do not need it use alpha = 1)
#pragma unroll 1
          for (int c = 0; c < kAttnD; c += 32) {
            uint32_t v[32];
            tmem_ld32(tmem_O + c, v);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) * alpha);
            tmem_st32(tmem_O + c, v);
          }
        }
      }
      mark(4);
      {
        uint32_t(&h0)[32] = *reinterpret_cast<uint32_t(*)[32]>(&h[0]);
        uint32_t(&h1)[32] = *reinterpret_cast<uint32_t(*)[32]>(&h[32]);
        tmem_st32(tmem_P, h0);
        tmem_st32(tmem_P + 32, h1);
      }
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(&p_full[t]);
      mark(5);
    };

    for (int j = 0; j < my_tiles - 1; ++j) tile_body(j, BoolTag<false>{});
    if (my_tiles > 0) tile_body(my_tiles - 1, BoolTag<true>{});
    if (kTrace && tracing && q == 0 && lane == 0) {
      for (int i = 0; i < 6; ++i) p.trace[t * 8 + i] = tr[i];
      p.trace[t * 8 + 6] = static_cast<unsigned long long>(clock64() - tr_begin);
      p.trace[t * 8 + 7] = static_cast<unsigned long long>(my_tiles);
    }

    // epilogue: O / rowsum -> fp16
    if (t < ntq) {
      mbar_wait(&pv_done[t], (ntiles - 1) & 1);
      tc_fence_after();
      const int row = qb * kAttnQBlock + t * kAttnTile + r;
      const float inv_l = 1.0f / l;
      __half* o = p.out + (static_cast<size_t>(b) * p.tq + row) * p.ld_out + head * kAttnD;
#pragma unroll 1
      for (int c = 0; c < kAttnD; c += 32) {
        uint32_t v[32];
        tmem_ld32(tmem_O + c, v);
        tmem_ld_wait();
        if (row < p.tq) {
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            uint4 w;
            w.x = pack_half2(__uint_as_float(v[8 * k + 0]) * inv_l, __uint_as_float(v[8 * k + 1]) * inv_l);
            w.y = pack_half2(__uint_as_float(v[8 * k + 2]) * inv_l, __uint_as_float(v[8 * k + 3]) * inv_l);
            w.z = pack_half2(__uint_as_float(v[8 * k + 4]) * inv_l, __uint_as_float(v[8 * k + 5]) * inv_l);
            w.w = pack_half2(__uint_as_float(v[8 * k + 6]) * inv_l, __uint_as_float(v[8 * k + 7]) * inv_l);
            *reinterpret_cast<uint4*>(o + c + 8 * k) = w;
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

// ------------------------------------------------------------------------------------------------------------------
// Persistent variant (opt-in: LR_ATTN_PERSIST=1, tq a multiple of 256, short key sequences - see build_attn_op): at most
// one CTA per SM, each walking the work items
// idx = blockIdx.x + k * gridDim.x, idx -> (batch, head, 256-query block) with the query block fastest (neighbouring CTAs
// share a head's K / V in L2). The pipeline never drains between items: mbarrier phases follow a GLOBAL KV-step counter,
// the query tiles are double buffered (q_full / q_empty), S of the next item's first KV tile is issued while the
// warpgroups still exponentiate / write out the current item, and TMEM, barriers and descriptors are set up once per
// CTA. What this buys: the cross-attention launches (77 keys = ONE KV step per item) were a chain of
// setup -> load -> MMA -> softmax -> MMA -> store per CTA, 1280 CTAs at 64x128: 60 -> 50 us. For the 8192-token
// self-attention (64 KV steps per item, fill / drain is 3 %) it measured 2 % SLOWER than attention_kernel, whose exp2
// phase ptxas happens to schedule better (profiles/r2_ab_attention_persistent.txt), so long sequences keep that kernel.
// Same roles, same tile_body arithmetic and the same TMEM map as attention_kernel; the ping-pong hand-over is pinned
// through shared memory here (pingpong_wait / pingpong_pass).
// ------------------------------------------------------------------------------------------------------------------
constexpr int kAttnPersistSmemBytes = 4 * kAttnTileBytes /*Q: 2 buffers x 2 tiles*/ +
                                      kAttnStages * 2 * kAttnTileBytes /*K,V ring*/ + 256 /*barriers*/;

template <bool kTrace>
__global__ void __launch_bounds__(kAttnThreads, 1) attention_persist_kernel(const __grid_constant__ AttnParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* q_s = smem;                                       // [2 buffers][2 query tiles]
  uint8_t* k_s = q_s + 4 * kAttnTileBytes;                   // [stages]
  uint8_t* v_s = k_s + kAttnStages * kAttnTileBytes;         // [stages]
  uint64_t* bars = reinterpret_cast<uint64_t*>(v_s + kAttnStages * kAttnTileBytes);
  uint64_t* q_full = bars;                         // [2] both query tiles of buffer b have landed
  uint64_t* q_empty = bars + 2;                    // [2] every S MMA that reads buffer b has completed
  uint64_t* kv_full = q_empty + 2;                 // [stages]
  uint64_t* kv_empty = kv_full + kAttnStages;      // [stages]
  uint64_t* s_full = kv_empty + kAttnStages;       // [2] S_t(step) is in TMEM
  uint64_t* s_empty = s_full + 2;                  // [2] warpgroup t holds S_t(step) in registers
  uint64_t* p_full = s_empty + 2;                  // [2] P_t(step) is in TMEM
  uint64_t* pv_done = p_full + 2;                  // [2] O_t += P_t(step) V has completed
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(pv_done + 2);
  const uint32_t pp_zero = smem_u32(tmem_slot + 1), pp_sink = smem_u32(tmem_slot + 2);  // see pingpong_wait / pingpong_pass

  const int hw_warp = threadIdx.x >> 5;
  const int warp = hw_warp;
  const int lane = threadIdx.x & 31;
  const int ntiles = (p.tk + kAttnTile - 1) / kAttnTile;
  const int nqb = p.tq / kAttnQBlock;
  const int total = nqb * p.heads * p.batch;
  const int first = blockIdx.x, stride = gridDim.x;
  const int n_items = (total - first + stride - 1) / stride;  // >= 1: the grid never exceeds the item count
  const int total_steps = n_items * ntiles;
  auto decode = [&](int k, int& qb, int& head, int& b) {
    const int idx = first + k * stride;
    qb = idx % nqb;
    const int hb = idx / nqb;
    head = hb % p.heads;
    b = hb / p.heads;
  };
  pdl_launch_dependents();

  if (threadIdx.x == 0) {
    if ((smem_u32(smem) & 1023u) != 0) {
      printf("lr_b200: attention smem base not 1024-aligned\n");
      __trap();
    }
    tma_prefetch_desc(&p.tmQ);
    tma_prefetch_desc(&p.tmK);
    tma_prefetch_desc(&p.tmV);
    *reinterpret_cast<float*>(tmem_slot + 1) = 0.0f;
    for (int i = 0; i < 2; ++i) {
      mbar_init(&q_full[i], 1);
      mbar_init(&q_empty[i], 2);  // one tcgen05.commit per issuing warp
    }
    for (int i = 0; i < kAttnStages; ++i) {
      mbar_init(&kv_full[i], 1);
      mbar_init(&kv_empty[i], 2);
    }
    for (int t = 0; t < 2; ++t) {
      mbar_init(&s_full[t], 1);
      mbar_init(&s_empty[t], 128);
      mbar_init(&p_full[t], 128);
      mbar_init(&pv_done[t], 1);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();  // predecessors complete: from here on global memory may be read and written
  if (warp < 4) {
    if (warp == 0) {
      // ------------------------------- TMA producer ------------------------------------------------------------
      int g = 0;
      for (int k = 0; k < n_items; ++k) {
        int qb, head, b;
        decode(k, qb, head, b);
        const int qbuf = k & 1;
        mbar_wait(&q_empty[qbuf], ((k >> 1) & 1) ^ 1);
        if (elect_one()) {
          uint8_t* q_dst = q_s + qbuf * 2 * kAttnTileBytes;
          mbar_arrive_expect_tx(&q_full[qbuf], 2 * kAttnTileBytes);
          tma_load_3d(q_dst, &p.tmQ, &q_full[qbuf], p.q_col0 + head * kAttnD, qb * kAttnQBlock, b);
          tma_load_3d(q_dst + kAttnTileBytes, &p.tmQ, &q_full[qbuf], p.q_col0 + head * kAttnD, qb * kAttnQBlock + kAttnTile, b);
        }
        __syncwarp();
        for (int j = 0; j < ntiles; ++j, ++g) {
          const int s = g % kAttnStages;
          const uint32_t ph = (g / kAttnStages) & 1;
          mbar_wait(&kv_empty[s], ph ^ 1);
          if (elect_one()) {
            mbar_arrive_expect_tx(&kv_full[s], 2 * kAttnTileBytes);
            tma_load_3d(k_s + s * kAttnTileBytes, &p.tmK, &kv_full[s], p.k_col0 + head * kAttnD, j * kAttnTile, b);
            tma_load_3d(v_s + s * kAttnTileBytes, &p.tmV, &kv_full[s], p.v_col0 + head * kAttnD, j * kAttnTile, b);
          }
          __syncwarp();
        }
      }
    } else if (warp < 3) {
      // ------------------------------- MMA issuers: warp 1 -> query tile 0, warp 2 -> query tile 1 -------------
      const int t = warp - 1;
      const uint32_t idesc_s = umma_idesc_f16(128, kAttnTile, 0);  // S: N = 128 keys, K-major B
      const uint32_t idesc_o = umma_idesc_f16(128, kAttnD, 1);     // O: N = 64 channels, MN-major B (V)
      const uint32_t desc_hi = umma_desc_hi_sw128(1024);
      const uint32_t q_lo = umma_desc_lo(smem_u32(q_s), 16);
      const uint32_t k_lo = umma_desc_lo(smem_u32(k_s), 16);
      const uint32_t v_lo = umma_desc_lo(smem_u32(v_s), 16);
      constexpr uint32_t kTileUnits = kAttnTileBytes >> 4;
      auto issue_s = [&](int stage, int qbuf) {  // S_t = Q_t K^T : 4 k-steps of 16 channels (32 B = 2 units each)
#pragma unroll
        for (int k = 0; k < kAttnD / 16; ++k) {
          umma_f16(tmem_base + kTmemS + t * 128, umma_desc_make(desc_hi, q_lo + (qbuf * 2 + t) * kTileUnits + 2 * k),
                   umma_desc_make(desc_hi, k_lo + stage * kTileUnits + 2 * k), idesc_s, k != 0 ? 1u : 0u);
        }
        umma_commit(&s_full[t]);
      };
      auto issue_pv = [&](int stage, uint32_t accumulate) {
#pragma unroll
        for (int k = 0; k < kAttnTile / 16; ++k) {
          umma_f16_ts(tmem_base + kTmemO + t * 64, tmem_base + kTmemP + t * 64 + k * 8,
                      umma_desc_make(desc_hi, v_lo + stage * kTileUnits + k * 128), idesc_o, (accumulate | k) != 0 ? 1u : 0u);
        }
        umma_commit(&pv_done[t]);
        umma_commit(&kv_empty[stage]);  // the stage is free once BOTH tiles' PV MMAs have read it
      };
      mbar_wait(&q_full[0], 0);
      mbar_wait(&kv_full[0], 0);
      tc_fence_after();
      if (elect_one()) issue_s(0, 0);
      __syncwarp();
      int g = 0;
      for (int k = 0; k < n_items; ++k) {
        for (int j = 0; j < ntiles; ++j, ++g) {
          if (g + 1 < total_steps) {
            // S of the next step runs one KV tile ahead: it needs the keys (and, across an item boundary, the next
            // item's queries) and the warpgroup to have pulled S(g) into registers
            int qbuf1 = k & 1;
            if (j + 1 == ntiles) {
              qbuf1 = (k + 1) & 1;
              mbar_wait(&q_full[qbuf1], ((k + 1) >> 1) & 1);
            }
            const int s1 = (g + 1) % kAttnStages;
            mbar_wait(&kv_full[s1], ((g + 1) / kAttnStages) & 1);
            mbar_wait(&s_empty[t], g & 1);
            tc_fence_after();
            if (elect_one()) issue_s(s1, qbuf1);
            __syncwarp();
          }
          mbar_wait(&p_full[t], g & 1);
          tc_fence_after();
          if (elect_one()) {
            issue_pv(g % kAttnStages, j > 0 ? 1u : 0u);
            if (j == ntiles - 1) umma_commit(&q_empty[k & 1]);  // every S MMA of this item has been issued before
          }
          __syncwarp();
        }
      }
    }
  } else {
    // ------------------------------- softmax warpgroups ---------------------------
    const int t = (warp - 4) >> 2;  // query tile of this warpgroup
    const int q = hw_warp & 3;      // TMEM lane quarter (hardware warp id % 4)
    const int r = q * 32 + lane;    // row inside the tile
    const uint32_t lane_addr = static_cast<uint32_t>(q * 32) << 16;
    const uint32_t tmem_S = tmem_base + kTmemS + t * 128 + lane_addr;
    const uint32_t tmem_P = tmem_base + kTmemP + t * 64 + lane_addr;
    const uint32_t tmem_O = tmem_base + kTmemO + t * 64 + lane_addr;
    float m_used = -INFINITY;  // max the current O / l are scaled against (raw logit units)
    float l = 0.f;

    unsigned long long tr[6] = {0, 0, 0, 0, 0, 0};
    long long tr_t = 0, tr_begin = 0;
    const bool tracing = kTrace && p.trace != nullptr && blockIdx.x == 0;
    if (kTrace) tr_begin = tr_t = clock64();
    auto mark = [&](int phase) {
      if (kTrace) {
        const long long now = clock64();
        tr[phase] += static_cast<unsigned long long>(now - tr_t);
        tr_t = now;
      }
    };
#if LR_ATTN_PINGPONG
    if (t == 1) asm volatile("bar.arrive 2, 256;" ::: "memory");  // warpgroup 0 goes first
#endif
    // g: global KV step of this CTA (mbarrier phases), j: KV tile inside the current item
    auto tile_body = [&](int g, int j, auto mask_tag) {
      constexpr bool kMask = decltype(mask_tag)::value;
      mbar_wait(&s_full[t], g & 1);
      tc_fence_after();
      mark(0);
      uint32_t s[128];
      {
        uint32_t(&s0)[32] = *reinterpret_cast<uint32_t(*)[32]>(&s[0]);
        uint32_t(&s1)[32] = *reinterpret_cast<uint32_t(*)[32]>(&s[32]);
        uint32_t(&s2)[32] = *reinterpret_cast<uint32_t(*)[32]>(&s[64]);
        uint32_t(&s3)[32] = *reinterpret_cast<uint32_t(*)[32]>(&s[96]);
        tmem_ld32(tmem_S + 0, s0);
        tmem_ld32(tmem_S + 32, s1);
        tmem_ld32(tmem_S + 64, s2);
        tmem_ld32(tmem_S + 96, s3);
        tmem_ld_wait();
      }
      tc_fence_before();
      mbar_arrive(&s_empty[t]);  // S_t may be overwritten by the next QK^T right away
      mark(1);
      if (kMask) {               // only the last KV tile can be partial (keys beyond tk were zero-filled by TMA)
        const int kv_valid = p.tk - j * kAttnTile;
#pragma unroll
        for (int i = 0; i < 128; ++i)
          if (i >= kv_valid) s[i] = 0xff800000u;  // -inf
      }
      float mx[8];
#pragma unroll
      for (int c = 0; c < 8; ++c) mx[c] = fmax3(__uint_as_float(s[c]), __uint_as_float(s[c + 8]), __uint_as_float(s[c + 16]));
#pragma unroll
      for (int i = 24; i < 120; i += 16) {
#pragma unroll
        for (int c = 0; c < 8; ++c) mx[c] = fmax3(mx[c], __uint_as_float(s[i + c]), __uint_as_float(s[i + c + 8]));
      }
#pragma unroll
      for (int c = 0; c < 8; ++c) mx[c] = fmaxf(mx[c], __uint_as_float(s[120 + c]));
      const float mrow = fmaxf(fmax3(mx[0], mx[1], mx[2]), fmaxf(fmax3(mx[3], mx[4], mx[5]), fmaxf(mx[6], mx[7])));
      const float m_new = fmaxf(m_used, mrow);
      const bool need = (m_new - m_used) * p.scale_log2 > kRescaleThreshold;  // true on the first tile (-inf)
      float alpha = 1.0f;
      if (need) {
        alpha = (m_used == -INFINITY) ? 0.f : fast_exp2((m_used - m_new) * p.scale_log2);
        m_used = m_new;
        l *= alpha;
      }
      mark(2);
      const float moff_pre = m_used * p.scale_log2;
      uint32_t h[64];
      float l0 = 0.f, l1 = 0.f, l2 = 0.f, l3 = 0.f;
      float e_last = 0.f;  // the last exponential issued: the hand-over waits for it, not for the row sum / packing
      auto exp4 = [&](int i, float moff) {  // logits 2i .. 2i+3 -> packed fp16 pairs h[i], h[i+1]
        const float e0 = fast_exp2(fmaf(__uint_as_float(s[2 * i]), p.scale_log2, -moff));
        const float e1 = fast_exp2(fmaf(__uint_as_float(s[2 * i + 1]), p.scale_log2, -moff));
        const float e2 = fast_exp2(fmaf(__uint_as_float(s[2 * i + 2]), p.scale_log2, -moff));
        const float e3 = fast_exp2(fmaf(__uint_as_float(s[2 * i + 3]), p.scale_log2, -moff));
        l0 += e0;
        l1 += e1;
        l2 += e2;
        l3 += e3;
        h[i] = pack_half2(e0, e1);
        h[i + 1] = pack_half2(e2, e3);
        e_last = e3;
      };
#pragma unroll
      for (int i = 0; i < LR_ATTN_PRE_EXP / 2; i += 2) exp4(i, moff_pre);  // head start (see LR_ATTN_PRE_EXP)
      float moff = moff_pre;
#if LR_ATTN_PINGPONG
      moff = pingpong_wait(2 + t, moff, pp_zero);  // my turn on the MUFU
#endif
#pragma unroll
      for (int i = LR_ATTN_PRE_EXP / 2; i < 64; i += 2) exp4(i, moff);
#if LR_ATTN_PINGPONG
      // hand the MUFU to the other warpgroup (its very last turn needs no successor)
      if (!(t == 1 && g == total_steps - 1)) pingpong_pass(3 - t, e_last, pp_sink);
#endif
      l += (l0 + l1) + (l2 + l3);
      mark(3);
      if (j > 0) {
        mbar_wait(&pv_done[t], (g - 1) & 1);  // PV of the previous step finished reading P_t and updating O_t
        tc_fence_after();
        if (__any_sync(0xffffffffu, need)) {
#pragma unroll 1
          for (int c = 0; c < kAttnD; c += 32) {
            uint32_t v[32];
            tmem_ld32(tmem_O + c, v);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) * alpha);
            tmem_st32(tmem_O + c, v);
          }
        }
      }
      mark(4);
      {
        uint32_t(&h0)[32] = *reinterpret_cast<uint32_t(*)[32]>(&h[0]);
        uint32_t(&h1)[32] = *reinterpret_cast<uint32_t(*)[32]>(&h[32]);
        tmem_st32(tmem_P, h0);
        tmem_st32(tmem_P + 32, h1);
      }
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(&p_full[t]);
      mark(5);
    };

    int g = 0;
    for (int k = 0; k < n_items; ++k) {
      m_used = -INFINITY;
      l = 0.f;
      for (int j = 0; j < ntiles - 1; ++j, ++g) tile_body(g, j, BoolTag<false>{});
      tile_body(g, ntiles - 1, BoolTag<true>{});
      ++g;
      // item epilogue: O / rowsum -> fp16. (The next item's first PV overwrites O_t only after this warpgroup's next
      // p_full arrive, which follows these loads in program order; its S is already being computed.)
      mbar_wait(&pv_done[t], (g - 1) & 1);
      tc_fence_after();
      int qb, head, b;
      decode(k, qb, head, b);
      const int row = qb * kAttnQBlock + t * kAttnTile + r;
      const float inv_l = 1.0f / l;
      __half* o = p.out + (static_cast<size_t>(b) * p.tq + row) * p.ld_out + head * kAttnD;
#pragma unroll 1
      for (int c = 0; c < kAttnD; c += 32) {
        uint32_t v[32];
        tmem_ld32(tmem_O + c, v);
        tmem_ld_wait();
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
          uint4 w;
          w.x = pack_half2(__uint_as_float(v[8 * kk + 0]) * inv_l, __uint_as_float(v[8 * kk + 1]) * inv_l);
          w.y = pack_half2(__uint_as_float(v[8 * kk + 2]) * inv_l, __uint_as_float(v[8 * kk + 3]) * inv_l);
          w.z = pack_half2(__uint_as_float(v[8 * kk + 4]) * inv_l, __uint_as_float(v[8 * kk + 5]) * inv_l);
          w.w = pack_half2(__uint_as_float(v[8 * kk + 6]) * inv_l, __uint_as_float(v[8 * kk + 7]) * inv_l);
          *reinterpret_cast<uint4*>(o + c + 8 * kk) = w;
        }
      }
    }
    if (kTrace && tracing && q == 0 && lane == 0) {
      for (int i = 0; i < 6; ++i) p.trace[t * 8 + i] = tr[i];
      p.trace[t * 8 + 6] = static_cast<unsigned long long>(clock64() - tr_begin);
      p.trace[t * 8 + 7] = static_cast<unsigned long long>(total_steps);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}


}  // namespace lr
