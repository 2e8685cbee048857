// Fused softmax(Q K^T * scale) V for sm_100a with tcgen05 tensor cores and TMEM accumulators.
// Replaces the einsum/softmax/einsum of CrossAttention.forward (reference ldm/modules/attention.py:165-196), which
// materialises the [b*h, Tq, Tk] fp32 logits (1.34 GB per canvas at T = 8192). Precision islands are kept:
// fp16 operands, fp32 QK^T accumulation (the reference's _ATTN_PRECISION="fp32" path), fp32 softmax statistics.
//
// One CTA = one (batch, head, 128-query tile). d_head = 64.
//   warp 0 (1 lane): TMA — Q once, then a 2-stage ring of {K_j, V_j} 128-token tiles (OOB rows zero-filled)
//   warp 1 (1 lane): tcgen05.mma — S = Q K_j^T (128x128, TMEM cols 0..127), O += P_j V_j (128x64, cols 128..191)
//   warps 2..5     : softmax — thread <-> query row (tcgen05.ld 32x32b), online max with lazy rescaling of the
//                    TMEM-resident O, P_j written as fp16 into a 128B-swizzled smem tile that feeds the PV MMA.
// V is consumed as an MN-major B operand straight from the [token, d] layout the QKV GEMM produces: no transposes.
// ~113 KB smem and 256 TMEM columns per CTA, so two CTAs share an SM and overlap each other's MMA/softmax phases.
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
};

constexpr int kAttnThreads = 192;
constexpr int kAttnTile = 128;
constexpr int kAttnD = 64;
constexpr int kAttnTileBytes = kAttnTile * kAttnD * 2;  // 16 KB
constexpr int kAttnStages = 2;
constexpr int kAttnSmemBytes = kAttnTileBytes * (1 + 2 * kAttnStages) + 2 * kAttnTileBytes /*P*/ + 128 /*barriers*/;
constexpr float kRescaleThreshold = 8.0f;  // log2 domain: P stays <= 256, exact in fp16/fp32 accumulators

__device__ __forceinline__ float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__global__ void __launch_bounds__(kAttnThreads, 2) attention_kernel(const __grid_constant__ AttnParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* q_s = smem;
  uint8_t* k_s = q_s + kAttnTileBytes;                       // [stages]
  uint8_t* v_s = k_s + kAttnStages * kAttnTileBytes;         // [stages]
  uint8_t* p_s = v_s + kAttnStages * kAttnTileBytes;         // 2 swizzle atoms of [128 x 64]
  uint64_t* bars = reinterpret_cast<uint64_t*>(p_s + 2 * kAttnTileBytes);
  uint64_t* q_full = bars;
  uint64_t* kv_full = bars + 1;    // [2]
  uint64_t* kv_empty = bars + 3;   // [2]
  uint64_t* s_full = bars + 5;
  uint64_t* p_full = bars + 6;
  uint64_t* o_full = bars + 7;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int qt = blockIdx.x, head = blockIdx.y, b = blockIdx.z;
  const int ntiles = (p.tk + kAttnTile - 1) / kAttnTile;

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
      mbar_init(&kv_empty[i], 1);
    }
    mbar_init(s_full, 1);
    mbar_init(p_full, 128);
    mbar_init(o_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 256);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_S = tmem_base;
  const uint32_t tmem_O = tmem_base + 128;

  if (warp == 0) {
    if (lane == 0) {
      mbar_arrive_expect_tx(q_full, kAttnTileBytes);
      tma_load_3d(q_s, &p.tmQ, q_full, p.q_col0 + head * kAttnD, qt * kAttnTile, b);
      for (int j = 0; j < ntiles; ++j) {
        const int s = j % kAttnStages;
        const uint32_t ph = (j / kAttnStages) & 1;
        mbar_wait(&kv_empty[s], ph ^ 1);
        mbar_arrive_expect_tx(&kv_full[s], 2 * kAttnTileBytes);
        tma_load_3d(k_s + s * kAttnTileBytes, &p.tmK, &kv_full[s], p.k_col0 + head * kAttnD, j * kAttnTile, b);
        tma_load_3d(v_s + s * kAttnTileBytes, &p.tmV, &kv_full[s], p.v_col0 + head * kAttnD, j * kAttnTile, b);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc_s = umma_idesc_f16(128, kAttnTile, 0);  // S: N = 128 keys, K-major B
      const uint32_t idesc_o = umma_idesc_f16(128, kAttnD, 1);     // O: N = 64 channels, MN-major B (V)
      const uint32_t q_addr = smem_u32(q_s);
      const uint32_t p_addr = smem_u32(p_s);
      mbar_wait(q_full, 0);
      for (int j = 0; j < ntiles; ++j) {
        const int s = j % kAttnStages;
        const uint32_t ph = (j / kAttnStages) & 1;
        mbar_wait(&kv_full[s], ph);
        tc_fence_after();
        const uint32_t k_addr = smem_u32(k_s + s * kAttnTileBytes);
        const uint32_t v_addr = smem_u32(v_s + s * kAttnTileBytes);
#pragma unroll
        for (int k = 0; k < kAttnD / 16; ++k) {
          umma_f16(tmem_S, umma_smem_desc_sw128(q_addr + k * 32, 1024, 16),
                   umma_smem_desc_sw128(k_addr + k * 32, 1024, 16), idesc_s, k != 0 ? 1u : 0u);
        }
        umma_commit(s_full);
        mbar_wait(p_full, j & 1);
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < kAttnTile / 16; ++k) {
          const uint32_t pa = p_addr + (k >> 2) * kAttnTileBytes + (k & 3) * 32;
          const uint32_t va = v_addr + k * 16 * 128;  // 16 token rows of 128 B
          umma_f16(tmem_O, umma_smem_desc_sw128(pa, 1024, 16), umma_smem_desc_sw128(va, 1024, 16), idesc_o,
                   (j | k) != 0 ? 1u : 0u);
        }
        umma_commit(&kv_empty[s]);
        if (j == ntiles - 1) umma_commit(o_full);
      }
    }
  } else {
    const int q = warp & 3;
    const int r = q * 32 + lane;
    const uint32_t lane_addr = static_cast<uint32_t>(q * 32) << 16;
    float m_used = -INFINITY;  // max the current O / l are scaled against (raw logit units)
    float l = 0.f;
    for (int j = 0; j < ntiles; ++j) {
      mbar_wait(s_full, j & 1);  // also implies PV_{j-1} has completed (commit covers all prior MMAs)
      tc_fence_after();
      const int kv_valid = min(kAttnTile, p.tk - j * kAttnTile);
      // pass 1: row max
      float mx = -INFINITY;
#pragma unroll 1
      for (int c = 0; c < kAttnTile; c += 32) {
        uint32_t v[32];
        tmem_ld32(tmem_S + lane_addr + c, v);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i)
          if (c + i < kv_valid) mx = fmaxf(mx, __uint_as_float(v[i]));
      }
      const float m_new = fmaxf(m_used, mx);
      const bool need = (m_new - m_used) * p.scale_log2 > kRescaleThreshold;  // true on the first tile (-inf)
      float alpha = 1.0f;
      if (need) {
        alpha = (m_used == -INFINITY) ? 0.f : fast_exp2((m_used - m_new) * p.scale_log2);
        m_used = m_new;
        l *= alpha;
      }
      if (j > 0 && __any_sync(0xffffffffu, need)) {
        // rescale the TMEM-resident O row (warp-collective; lanes that do not need it use alpha = 1)
#pragma unroll 1
        for (int c = 0; c < kAttnD; c += 32) {
          uint32_t v[32];
          tmem_ld32(tmem_O + lane_addr + c, v);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) * alpha);
          tmem_st32(tmem_O + lane_addr + c, v);
        }
        tmem_st_wait();
      }
      // pass 2: P = exp2((s - m_used) * scale_log2) -> fp16, swizzled K-major smem tile
      const float moff = m_used * p.scale_log2;
      uint8_t* prow = p_s + r * 128;
#pragma unroll 1
      for (int c = 0; c < kAttnTile; c += 32) {
        uint32_t v[32];
        tmem_ld32(tmem_S + lane_addr + c, v);
        tmem_ld_wait();
        uint32_t h[16];
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          float p0 = (c + i < kv_valid) ? fast_exp2(__uint_as_float(v[i]) * p.scale_log2 - moff) : 0.f;
          float p1 = (c + i + 1 < kv_valid) ? fast_exp2(__uint_as_float(v[i + 1]) * p.scale_log2 - moff) : 0.f;
          l += p0 + p1;
          h[i >> 1] = pack_half2(p0, p1);
        }
        // columns [c, c+32) = 4 chunks of 16 B inside atom (c / 64)
        uint8_t* atom = prow + (c >> 6) * kAttnTileBytes;
        const int chunk0 = (c & 63) >> 3;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int ch = (chunk0 + k) ^ (r & 7);
          *reinterpret_cast<uint4*>(atom + ch * 16) = make_uint4(h[4 * k], h[4 * k + 1], h[4 * k + 2], h[4 * k + 3]);
        }
      }
      fence_proxy_async_smem();
      tc_fence_before();
      mbar_arrive(p_full);
    }
    // epilogue: O / l -> fp16
    mbar_wait(o_full, 0);
    tc_fence_after();
    const int row = qt * kAttnTile + r;
    const float inv_l = 1.0f / l;
    __half* o = p.out + (static_cast<size_t>(b) * p.tq + row) * p.ld_out + head * kAttnD;
#pragma unroll 1
    for (int c = 0; c < kAttnD; c += 32) {
      uint32_t v[32];
      tmem_ld32(tmem_O + lane_addr + c, v);
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

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 256);
}

}  // namespace lr
