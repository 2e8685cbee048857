// HBM-bound support kernels of the UNet forward: GroupNorm (statistics + apply + SiLU, two-source so the skip concat
// is read in place), LayerNorm, first-layer im2col from the NCHW fp32 boundary tensor, nearest x2 upsample,
// timestep-embedding MLPs, NHWC->NCHW output conversion, and the fused CFG + DDIM update.
// All activations are NHWC fp16: row = (n*H + y)*W + x, channels contiguous; 16-byte (8 x fp16) vector accesses.
#pragma once
#include "ptx.cuh"

namespace lr {

constexpr int kNormThreads = 512;
// -DLR_GN_TRACE=1: gn_persistent_kernel keeps clock64 stamps of its phases and prints them with LR_GN_DEBUG=16
// (tests/gpu_gn_trace.py; costs registers, off in the shipped build)
#ifndef LR_GN_TRACE
#define LR_GN_TRACE 0
#endif

__device__ __forceinline__ void load8(const __half* p, float (&f)[8]) {
  const uint4 v = *reinterpret_cast<const uint4*>(p);
  const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 t = unpack_half2(w[i]);
    f[2 * i] = t.x;
    f[2 * i + 1] = t.y;
  }
}
__device__ __forceinline__ void store8(__half* p, const float (&f)[8]) {
  *reinterpret_cast<uint4*>(p) =
      make_uint4(pack_half2(f[0], f[1]), pack_half2(f[2], f[3]), pack_half2(f[4], f[5]), pack_half2(f[6], f[7]));
}
// x * sigmoid(x) with MUFU.EX2 + MUFU.RCP (no IEEE division: the slow path of `/` made GroupNorm+SiLU issue bound;
// flush-to-zero forms: __expf / __fdividef add three range-fixup instructions per element that change nothing here,
// 1 + e rounds to 1 long before e is denormal)
__device__ __forceinline__ float silu_f(float x) { return x * rcp_ftz(1.0f + ex2_ftz(x * -1.4426950408889634f)); }
// [silu](x * sc + sh) on the two fp16 values of one packed word, in packed fp32 pairs (FFMA2 / FMUL2 / FADD2)
__device__ __forceinline__ uint32_t affine_silu_h2(uint32_t w, f32x2 sc, f32x2 sh, int do_silu) {
  const float2 t = unpack_half2(w);
  f32x2 y = fma2(pk2(t.x, t.y), sc, sh);
  if (do_silu) {
    float a0, a1;
    upk2(mul2(y, pk2(-1.4426950408889634f, -1.4426950408889634f)), a0, a1);
    float d0, d1;
    upk2(add2(pk2(ex2_ftz(a0), ex2_ftz(a1)), pk2(1.0f, 1.0f)), d0, d1);
    y = mul2(y, pk2(rcp_ftz(d0), rcp_ftz(d1)));
  }
  float y0, y1;
  upk2(y, y0, y1);
  return pack_half2(y0, y1);
}

// ------------------------------------------------------------------------------------------------------------
// GroupNorm (reference: GroupNorm32, ldm/modules/diffusionmodules/util.py:217-219, fp32 math; Normalize,
// ldm/modules/attention.py:90-91). Sources x0 [n, P, c0] and x1 [n, P, c1] form the channel concat.
//
// Thread mapping (all GroupNorm kernels): thread -> (8-channel vector column `vec`, pixel slice `rsub`), rpi = 512 / nvec
// slices sweep the pixel range rpi rows at a time, kGnBatch independent 16-byte loads in flight per thread (the loads
// of a batch are predicated, never serialised by a tail loop). Every reduction runs in a FIXED order (thread-local
// order, shared-memory slice order, cluster rank order), so statistics are bit-reproducible and independent of the
// batch size.
//
// Two execution schemes, chosen per call site by launch_groupnorm:
//   * fused (gn_fused_cluster_kernel): one cluster of CS CTAs per image. Pass 1 reads the image once and reduces to
//     per-group (sum, sum of squares) in fp64; the CS partials are exchanged through distributed shared memory and
//     every CTA finalises (mean, rstd) itself; pass 2 re-reads its own pixels (L2 / L1 resident: the whole image is
//     a few MB at most) and writes y = [silu]((x - mean) * rstd * gamma + beta). One launch, no atomics, no scratch.
//   * persistent (gn_persistent_kernel) for images too large for one cluster: ONE launch of at most one CTA per SM.
//     Every image is cut into gn_chunks(P) pixel chunks (a function of the image size only); each CTA owns a contiguous
//     span of (image, chunk) items. Pass 1 streams its items from HBM and leaves one fp64 (sum, sum of squares) per
//     (item, group) in the scratch table - plain stores, no atomics; a grid-wide barrier (all CTAs are resident: the
//     grid never exceeds the SM count); pass 2 sums the chunk partials of an image in chunk order (fixed tree), derives
//     (mean, rstd) and re-reads its own items - L2 hits, the items were just streamed through it - to write
//     y = [silu]((x - mean) * rstd * gamma + beta). Bit-reproducible and independent of the batch size and of the
//     number of CTAs. (Round 1 used a statistics kernel with fp64 atomics + an apply kernel: two launches of short-lived
//     CTAs that reached 2.2 TB/s; profiles/r2_ab_gn_persistent.txt.)
// Scratch layout per call site: unsigned counters[4] (arrive, depart: zero on entry, the kernel leaves them zero);
// at byte 256: double2 part[n][chunks][G].
// ------------------------------------------------------------------------------------------------------------
constexpr int kGnBatch = 8;
constexpr int kGnPersistBatch = 4;  // gn_persistent_kernel: two CTAs per SM at <= 64 registers instead of deeper batches

// pixel chunks per image (depends on the image size ONLY): 37 chunks per image - 4 images fill the 148 SMs exactly, 8
// images give every CTA two items (32 chunks left 40 CTAs with one item and 108 with two) - of 16 to 2048 pixels; the
// first-stage decoder's 512 x 1024 maps get 256 chunks, enough for a single image to occupy the GPU
__host__ __device__ inline int gn_chunk_pixels(int P) {
  int chunk = (P + 36) / 37;
  if (chunk < 16) chunk = 16;
  if (chunk > 2048) chunk = 2048;
  return chunk;
}
__host__ __device__ inline int gn_chunks(int P) { const int c = gn_chunk_pixels(P); return (P + c - 1) / c; }
__host__ __device__ inline size_t gn_scratch_bytes(int n, int groups, int P) {
  return 256 + static_cast<size_t>(n) * gn_chunks(P) * groups * 2 * sizeof(double);
}

__device__ __forceinline__ uint4 ldg16(const __half* p) { return __ldg(reinterpret_cast<const uint4*>(p)); }
__device__ __forceinline__ void unpack8(const uint4& v, float (&f)[8]) {
  const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 t = unpack_half2(w[i]);
    f[2 * i] = t.x;
    f[2 * i + 1] = t.y;
  }
}

// Compiler barrier on a packed vector: stops ptxas/nvcc from keeping the UNPACKED fp32 copy of a register-resident row
// live across passes (which doubles the register footprint); re-unpacking costs 4 cheap converts.
__device__ __forceinline__ void keep_packed(uint4& v) { asm volatile("" : "+r"(v.x), "+r"(v.y), "+r"(v.z), "+r"(v.w)); }

// per-thread partial (sum, sum of squares) of 8 channels over pixels p0, p0 + rpi, ... < p_end
template <int kB = kGnBatch>
__device__ __forceinline__ void gn_accumulate(const __half* __restrict__ base, int ld, int p0, int p_end, int rpi,
                                              float (&s)[8], float (&ss)[8]) {
  for (int p = p0; p < p_end; p += kB * rpi) {
    uint4 v[kB];
#pragma unroll
    for (int u = 0; u < kB; ++u) {
      const int pp = p + u * rpi;
      v[u] = (pp < p_end) ? ldg16(base + static_cast<size_t>(pp) * ld) : make_uint4(0u, 0u, 0u, 0u);
    }
#pragma unroll
    for (int u = 0; u < kB; ++u) {
      float f[8];
      unpack8(v[u], f);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        s[j] += f[j];
        ss[j] = fmaf(f[j], f[j], ss[j]);
      }
    }
  }
}

// CTA-wide fixed-order reduction of the per-thread partials to per-group fp64 (sum, sum of squares).
// sm: [rpi][2][C] floats. On return threads g < groups hold their group's sums in (a, b).
__device__ __forceinline__ void gn_cta_group_sums(float* sm, int C, int rpi, int rsub, int ch, int groups,
                                                  const float (&s)[8], const float (&ss)[8], double& a, double& b) {
  if (rsub < rpi) {
    float* dst = sm + static_cast<size_t>(rsub) * 2 * C + ch;
    *reinterpret_cast<float4*>(dst) = make_float4(s[0], s[1], s[2], s[3]);
    *reinterpret_cast<float4*>(dst + 4) = make_float4(s[4], s[5], s[6], s[7]);
    *reinterpret_cast<float4*>(dst + C) = make_float4(ss[0], ss[1], ss[2], ss[3]);
    *reinterpret_cast<float4*>(dst + C + 4) = make_float4(ss[4], ss[5], ss[6], ss[7]);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) {
    float acc = sm[i];
    for (int rr = 1; rr < rpi; ++rr) acc += sm[static_cast<size_t>(rr) * 2 * C + i];
    sm[i] = acc;
  }
  __syncthreads();
  a = 0.0;
  b = 0.0;
  const int cpg = C / groups;
  if (static_cast<int>(threadIdx.x) < groups) {
    const int g = threadIdx.x;
    for (int j = 0; j < cpg; ++j) {
      a += static_cast<double>(sm[g * cpg + j]);
      b += static_cast<double>(sm[C + g * cpg + j]);
    }
  }
}

__device__ __forceinline__ float2 gn_finalize(double sum, double sq, double cnt, float eps) {
  const double mean = sum / cnt;
  double var = sq / cnt - mean * mean;
  if (var < 0.0) var = 0.0;
  return make_float2(static_cast<float>(mean), static_cast<float>(1.0 / sqrt(var + static_cast<double>(eps))));
}

// y = [silu](x * sc + sh) for pixels p0, p0 + rpi, ... < p_end of one 8-channel column
template <int kB = kGnBatch>
__device__ __forceinline__ void gn_apply_rows(const __half* __restrict__ base, int ld, __half* __restrict__ o, int C,
                                              int p0, int p_end, int rpi, const float (&sc)[8], const float (&sh)[8],
                                              int do_silu) {
  // Software pipeline of two half batches: the loads of one are in flight while the other is normalised and stored.
  // With SiLU the math alone runs near the HBM rate (EX2 -> RCP per element at 16 / clk / SM = 16 B / clk / SM at best),
  // so load latency has to hide behind it, not add to it (tests/gpu_time_gn_passes.py: 21 -> 13 us on 8 x 64x128x320).
  constexpr int kH = kB / 2;
  const int step = kH * rpi;
  f32x2 sc2[4], sh2[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    sc2[i] = pk2(sc[2 * i], sc[2 * i + 1]);
    sh2[i] = pk2(sh[2 * i], sh[2 * i + 1]);
  }
  auto load = [&](int p, uint4 (&v)[kH]) {
#pragma unroll
    for (int u = 0; u < kH; ++u) {
      const int pp = p + u * rpi;
      v[u] = (pp < p_end) ? ldg16(base + static_cast<size_t>(pp) * ld) : make_uint4(0u, 0u, 0u, 0u);
    }
  };
  auto proc = [&](int p, const uint4 (&v)[kH]) {
#pragma unroll
    for (int u = 0; u < kH; ++u) {
      const int pp = p + u * rpi;
      if (pp < p_end) {
        uint4 y;
        y.x = affine_silu_h2(v[u].x, sc2[0], sh2[0], do_silu);
        y.y = affine_silu_h2(v[u].y, sc2[1], sh2[1], do_silu);
        y.z = affine_silu_h2(v[u].z, sc2[2], sh2[2], do_silu);
        y.w = affine_silu_h2(v[u].w, sc2[3], sh2[3], do_silu);
        *reinterpret_cast<uint4*>(o + static_cast<size_t>(pp) * C) = y;
      }
    }
  };
  if (p0 >= p_end) return;
  uint4 va[kH], vb[kH];
  int p = p0;
  load(p, va);
  while (true) {
    const int pn = p + step;
    if (pn < p_end) load(pn, vb);
    proc(p, va);
    if (pn >= p_end) break;
    p = pn + step;
    if (p < p_end) load(p, va);
    proc(pn, vb);
    if (p >= p_end) break;
  }
}

// Persistent two-pass GroupNorm (see the header comment). grid = min(2 * #SM, n * chunks) CTAs; block = kNormThreads;
// dynamic smem = rpi * 2 * C floats.
__global__ void __launch_bounds__(kNormThreads, 2) gn_persistent_kernel(const __half* __restrict__ x0, int c0,
                                                                        const __half* __restrict__ x1, int c1, int P,
                                                                        int n_img, int groups, float eps,
                                                                        const float* __restrict__ gamma,
                                                                        const float* __restrict__ beta, int do_silu,
                                                                        __half* __restrict__ out,
                                                                        unsigned char* __restrict__ scratch, int dbg) {
  pdl_launch_dependents();
  pdl_wait();
  extern __shared__ float sm[];  // [rpi][2][C]
  __shared__ float2 s_mr[64];
  const int C = c0 + c1;
  const int nvec = C / 8;
  const int chunk = gn_chunk_pixels(P), chunks = gn_chunks(P);
  unsigned* cnt = reinterpret_cast<unsigned*>(scratch);
  double2* part = reinterpret_cast<double2*>(scratch + 256);
  const int items = n_img * chunks;
  const int G = gridDim.x;
  const int i0 = static_cast<int>(static_cast<long long>(blockIdx.x) * items / G);
  const int i1 = static_cast<int>(static_cast<long long>(blockIdx.x + 1) * items / G);
  const int rpi = blockDim.x / nvec;  // pixel rows handled per sweep
  const int vec = threadIdx.x % nvec;
  const int rsub = threadIdx.x / nvec;
  const int ch = vec * 8;
  const __half* src = (ch < c0) ? x0 + ch : x1 + (ch - c0);
  const int ld = (ch < c0) ? c0 : c1;

  // ---- pass 1: per-(item, group) partial sums ----
#if LR_GN_TRACE
  long long tr[6] = {0, 0, 0, 0, 0, 0};
  tr[0] = clock64();
#endif
  for (int it = i0; it < i1 && !(dbg & 1); ++it) {
    const int n = it / chunks, k = it - n * chunks;
    const int p_begin = k * chunk, p_end = min(P, p_begin + chunk);
    float s[8], ss[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) s[j] = ss[j] = 0.f;
    if (rsub < rpi) gn_accumulate<kGnPersistBatch>(src + static_cast<size_t>(n) * P * ld, ld, p_begin + rsub, p_end, rpi, s, ss);
    double a, b;
    gn_cta_group_sums(sm, C, rpi, rsub, ch, groups, s, ss, a, b);
    if (static_cast<int>(threadIdx.x) < groups) part[static_cast<size_t>(it) * groups + threadIdx.x] = make_double2(a, b);
    __syncthreads();  // sm is rewritten by the next item
  }

  // ---- grid barrier: every CTA's partials are visible ----
#if LR_GN_TRACE
  tr[1] = clock64();
#endif
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    atomicAdd(&cnt[0], 1u);
    const long long t0 = clock64();
    unsigned seen;
    do {
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(cnt) : "memory");
      if (seen < static_cast<unsigned>(G) && clock64() - t0 > 4000000000LL) {
        printf("lr_b200: groupnorm grid barrier timed out (%u of %d CTAs arrived)\n", seen, G);
        __trap();
      }
    } while (seen < static_cast<unsigned>(G));
  }
  __syncthreads();

  // ---- pass 2: finalise per image, normalise own items ----
#if LR_GN_TRACE
  tr[2] = clock64();
#endif
  const int cpg = C / groups;
  const int tpg = blockDim.x / groups;  // threads that share the chunk sum of one group (a power of two <= 32)
  const double cntd = static_cast<double>(P) * cpg;
  int cur_n = -1;
  float sc[8], sh[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) sc[j] = sh[j] = 0.f;
  for (int it = i0; it < i1; ++it) {
    const int n = it / chunks, k = it - n * chunks;
    if (n != cur_n) {
      __syncthreads();  // everybody is done with the previous image's s_mr
      // gamma / beta of this thread's channels (in flight while the chunk partials are summed)
      float4 gam[2], bet[2];
      if (rsub < rpi) {
        gam[0] = __ldg(reinterpret_cast<const float4*>(gamma + ch));
        gam[1] = __ldg(reinterpret_cast<const float4*>(gamma + ch) + 1);
        bet[0] = __ldg(reinterpret_cast<const float4*>(beta + ch));
        bet[1] = __ldg(reinterpret_cast<const float4*>(beta + ch) + 1);
      }
      const int g = threadIdx.x / tpg, sub = threadIdx.x % tpg;
      double a = 0.0, b = 0.0;
      if (g < groups) {
        const double2* pr = part + (static_cast<size_t>(n) * chunks) * groups + g;
        // chunk order within a lane (four loads in flight), then a fixed butterfly over the lanes
        for (int j = sub; j < chunks; j += 4 * tpg) {
          double2 v[4];
#pragma unroll
          for (int u = 0; u < 4; ++u)
            v[u] = (j + u * tpg < chunks) ? __ldcg(pr + static_cast<size_t>(j + u * tpg) * groups) : make_double2(0.0, 0.0);
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            a += v[u].x;
            b += v[u].y;
          }
        }
      }
      for (int o = tpg >> 1; o > 0; o >>= 1) {
        a += __shfl_xor_sync(0xffffffffu, a, o);
        b += __shfl_xor_sync(0xffffffffu, b, o);
      }
      if (g < groups && sub == 0) s_mr[g] = gn_finalize(a, b, cntd, eps);
      __syncthreads();
      cur_n = n;
      if (rsub < rpi) {  // the affine coefficients of this thread's 8 channels change with the image only
        const float gm[8] = {gam[0].x, gam[0].y, gam[0].z, gam[0].w, gam[1].x, gam[1].y, gam[1].z, gam[1].w};
        const float bt[8] = {bet[0].x, bet[0].y, bet[0].z, bet[0].w, bet[1].x, bet[1].y, bet[1].z, bet[1].w};
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float2 m = s_mr[(ch + j) / cpg];
          sc[j] = m.y * gm[j];
          sh[j] = bt[j] - m.x * sc[j];
        }
      }
#if LR_GN_TRACE
      if (it == i0) tr[3] = clock64();
#endif
    }
    if (rsub < rpi && !(dbg & 2)) {
      const int p_begin = k * chunk, p_end = min(P, p_begin + chunk);
      gn_apply_rows<kGnPersistBatch>(src + static_cast<size_t>(n) * P * ld, ld, out + static_cast<size_t>(n) * P * C + ch, C, p_begin + rsub,
                    p_end, rpi, sc, sh, do_silu);
    }
  }

  // ---- leave the counters zero for the next launch on this scratch ----
#if LR_GN_TRACE
  tr[4] = clock64();
  if ((dbg & 16) && threadIdx.x == 0 && (blockIdx.x == 0 || blockIdx.x == 77))
    printf("gn trace cta %d items %d: pass1 %lld barrier %lld stats %lld apply %lld\n", blockIdx.x, i1 - i0, tr[1] - tr[0],
           tr[2] - tr[1], tr[3] - tr[2], tr[4] - tr[3]);
#endif
  if (threadIdx.x == 0) {
    const unsigned left = atomicAdd(&cnt[1], 1u);
    if (left == static_cast<unsigned>(G) - 1) {  // everybody has passed the barrier: nobody reads cnt[0] any more
      cnt[0] = 0;
      cnt[1] = 0;
      __threadfence();
    }
  }
}

// GroupNorm apply [+ SiLU] from the per-(image, channel) coefficients gn_finalize_kernel derived from the PRODUCER's
// statistics: y = [silu](x * scale[n, c] + shift[n, c]). One read + one write pass (the stand-alone scheme reads twice).
// Same thread mapping as the other GroupNorm kernels. grid = (pixel chunks, n); block = kNormThreads.
__global__ void __launch_bounds__(kNormThreads) gn_apply_coef_kernel(const __half* __restrict__ x0, int c0,
                                                                     const __half* __restrict__ x1, int c1, int P,
                                                                     int chunk, const float* __restrict__ scale,
                                                                     const float* __restrict__ shift, int do_silu,
                                                                     __half* __restrict__ out) {
  pdl_launch_dependents();
  pdl_wait();
  const int C = c0 + c1;
  const int nvec = C / 8;
  const int n = blockIdx.y;
  const int p_begin = blockIdx.x * chunk;
  const int p_end = min(P, p_begin + chunk);
  const int rpi = blockDim.x / nvec;
  const int vec = threadIdx.x % nvec;
  const int rsub = threadIdx.x / nvec;
  if (rsub >= rpi) return;
  const int ch = vec * 8;
  float sc[8], sh[8];
  {
    const float4* a = reinterpret_cast<const float4*>(scale + static_cast<size_t>(n) * C + ch);
    const float4* b = reinterpret_cast<const float4*>(shift + static_cast<size_t>(n) * C + ch);
    const float4 a0 = __ldg(a), a1 = __ldg(a + 1), b0 = __ldg(b), b1 = __ldg(b + 1);
    sc[0] = a0.x; sc[1] = a0.y; sc[2] = a0.z; sc[3] = a0.w; sc[4] = a1.x; sc[5] = a1.y; sc[6] = a1.z; sc[7] = a1.w;
    sh[0] = b0.x; sh[1] = b0.y; sh[2] = b0.z; sh[3] = b0.w; sh[4] = b1.x; sh[5] = b1.y; sh[6] = b1.z; sh[7] = b1.w;
  }
  const __half* base = (ch < c0) ? x0 + ch : x1 + (ch - c0);
  const int ld = (ch < c0) ? c0 : c1;
  base += static_cast<size_t>(n) * P * ld;
  __half* o = out + static_cast<size_t>(n) * P * C + ch;
  gn_apply_rows(base, ld, o, C, p_begin + rsub, p_end, rpi, sc, sh, do_silu);
}

// Fused single-launch GroupNorm: grid = (CS, n) with cluster dims (CS, 1, 1); block = kNormThreads;
// dynamic smem = rpi * 2 * C floats.
__device__ __forceinline__ double ld_dsmem_f64(const double* p, uint32_t rank) {
  uint32_t ra;
  double v;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(smem_u32(p)), "r"(rank));
  asm volatile("ld.shared::cluster.f64 %0, [%1];" : "=d"(v) : "r"(ra) : "memory");
  return v;
}
// split cluster barrier (non-.aligned forms: callers may arrive right after thread-divergent code)
__device__ __forceinline__ void cluster_arrive() { asm volatile("barrier.cluster.arrive.release;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait() { asm volatile("barrier.cluster.wait.acquire;" ::: "memory"); }

__global__ void __launch_bounds__(kNormThreads) gn_fused_cluster_kernel(const __half* __restrict__ x0, int c0,
                                                                        const __half* __restrict__ x1, int c1, int P,
                                                                        int groups, float eps,
                                                                        const float* __restrict__ gamma,
                                                                        const float* __restrict__ beta, int do_silu,
                                                                        __half* __restrict__ out) {
  pdl_launch_dependents();
  pdl_wait();
  extern __shared__ float sm[];        // [rpi][2][C]
  __shared__ double s_part[2 * 64];    // this CTA's per-group (sum, sum of squares), read by the whole cluster
  __shared__ float2 s_mr[64];
  const int C = c0 + c1;
  const int nvec = C / 8;
  const int n = blockIdx.y;
  const int cs = gridDim.x;  // == cluster size
  const uint32_t rank = cluster_ctarank();
  const int per = (P + cs - 1) / cs;
  const int p_begin = static_cast<int>(rank) * per;
  const int p_end = min(P, p_begin + per);
  const int rpi = blockDim.x / nvec;
  const int vec = threadIdx.x % nvec;
  const int rsub = threadIdx.x / nvec;
  const int ch = vec * 8;
  const __half* base = (ch < c0) ? x0 + ch : x1 + (ch - c0);
  const int ld = (ch < c0) ? c0 : c1;
  base += static_cast<size_t>(n) * P * ld;
  float s[8], ss[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) s[j] = ss[j] = 0.f;
  if (rsub < rpi) gn_accumulate(base, ld, p_begin + rsub, p_end, rpi, s, ss);
  double a, b;
  gn_cta_group_sums(sm, C, rpi, rsub, ch, groups, s, ss, a, b);
  if (static_cast<int>(threadIdx.x) < groups) {
    s_part[2 * threadIdx.x] = a;
    s_part[2 * threadIdx.x + 1] = b;
  }
  cluster_arrive();
  cluster_wait();
  if (static_cast<int>(threadIdx.x) < groups) {
    double sum = 0.0, sq = 0.0;
    for (int r = 0; r < cs; ++r) {  // rank order: identical in every CTA
      sum += ld_dsmem_f64(&s_part[2 * threadIdx.x], r);
      sq += ld_dsmem_f64(&s_part[2 * threadIdx.x + 1], r);
    }
    s_mr[threadIdx.x] = gn_finalize(sum, sq, static_cast<double>(P) * (C / groups), eps);
  }
  cluster_arrive();  // our remote reads are done; matched by the wait at the end (keeps every CTA's smem alive)
  __syncthreads();
  if (rsub < rpi) {
    const int cpg = C / groups;
    float sc[8], sh[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float2 m = s_mr[(ch + j) / cpg];
      sc[j] = m.y * gamma[ch + j];
      sh[j] = beta[ch + j] - m.x * sc[j];
    }
    __half* o = out + static_cast<size_t>(n) * P * C + ch;
    gn_apply_rows(base, ld, o, C, p_begin + rsub, p_end, rpi, sc, sh, do_silu);
  }
  cluster_wait();
}

// ------------------------------------------------------------------------------------------------------------
// GroupNorm from PRODUCER statistics (gemm_tc.cuh header): the kernel that wrote the tensor left, per 64-row tile half
// and channel, (sum, sum of squares) of the fp16 values in part[rows][C]; the rows of image n are [n*ppi, (n+1)*ppi).
// One CTA per (group, image) sums its group's entries in a FIXED order (thread-local sequence, then a fixed shared-memory
// tree; fp64), and writes the affine coefficients the consumer's transform warps apply:
//   scale[n, c] = rstd * gamma[c],  shift[n, c] = beta[c] - mean * scale[n, c]      (same formulas as the stand-alone GroupNorm kernels)
// Two sources = the channel concat of the skip connection (openaimodel.py:781); their tables may have different ppi.
// ------------------------------------------------------------------------------------------------------------
constexpr int kGnFinalizeThreads = 128;
__global__ void __launch_bounds__(kGnFinalizeThreads) gn_finalize_kernel(const float2* __restrict__ part0, int ppi0, int c0,
                                                                         const float2* __restrict__ part1, int ppi1, int c1,
                                                                         double count, float eps,
                                                                         const float* __restrict__ gamma,
                                                                         const float* __restrict__ beta,
                                                                         float* __restrict__ scale,
                                                                         float* __restrict__ shift) {
  pdl_launch_dependents();
  pdl_wait();
  __shared__ double red[2][kGnFinalizeThreads];
  const int groups = gridDim.x, g = blockIdx.x, n = blockIdx.y;
  const int C = c0 + c1, cpg = C / groups;
  const int ch0 = g * cpg;
  double s = 0.0, q = 0.0;
  // channels of this group that live in source 0 / source 1 (a group may straddle the concat boundary)
  const int a0 = min(max(c0 - ch0, 0), cpg);  // first a0 channels of the group are in source 0
  // entries of one source: eight independent loads in flight per thread, added in index order (as a plain loop would)
  auto sum_source = [&](const float2* __restrict__ part, int ppi, int cs, int chs, int na) {
    const int tot = ppi * na;
    for (int base = threadIdx.x; base < tot; base += 8 * kGnFinalizeThreads) {
      float2 v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int idx = base + u * kGnFinalizeThreads;
        const int i = idx / na, c = idx - i * na;
        v[u] = (idx < tot) ? __ldg(part + (static_cast<size_t>(n) * ppi + i) * cs + chs + c) : make_float2(0.f, 0.f);
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        if (base + u * kGnFinalizeThreads < tot) {
          s += static_cast<double>(v[u].x);
          q += static_cast<double>(v[u].y);
        }
      }
    }
  };
  if (a0 > 0) sum_source(part0, ppi0, c0, ch0, a0);
  const int a1 = cpg - a0;
  const int ch1 = ch0 + a0 - c0;  // first channel of the group inside source 1
  if (a1 > 0) sum_source(part1, ppi1, c1, ch1, a1);
  red[0][threadIdx.x] = s;
  red[1][threadIdx.x] = q;
  __syncthreads();
  for (int o = kGnFinalizeThreads / 2; o > 0; o >>= 1) {
    if (static_cast<int>(threadIdx.x) < o) {
      red[0][threadIdx.x] += red[0][threadIdx.x + o];
      red[1][threadIdx.x] += red[1][threadIdx.x + o];
    }
    __syncthreads();
  }
  const float2 mr = gn_finalize(red[0][0], red[1][0], count, eps);
  for (int c = threadIdx.x; c < cpg; c += kGnFinalizeThreads) {
    const float sc = mr.y * gamma[ch0 + c];
    scale[static_cast<size_t>(n) * C + ch0 + c] = sc;
    shift[static_cast<size_t>(n) * C + ch0 + c] = beta[ch0 + c] - mr.x * sc;
  }
}

// ------------------------------------------------------------------------------------------------------------
// LayerNorm over the channel dim (nn.LayerNorm eps 1e-5, attention.py:266-268).
// HBM-streaming structure: the token matrix is contiguous, so a tile of kLnTileRows rows is ONE 1-D bulk copy
// (cp.async.bulk global -> shared, completion on an mbarrier). A producer warp keeps `stages` tiles in flight per CTA
// (bytes in flight are bounded by shared memory, not by registers); 8 consumer warps each normalise two rows of the
// tile (two-pass statistics on register-resident rows, interleaved for ILP) and store fp16 rows straight to global.
// Persistent: grid-stride over tiles. gamma / beta are staged once per CTA.
// Dynamic smem: stages * kLnTileRows * C * 2 (tiles) + 2 * C * 4 (gamma, beta) + 2 * stages * 8 (barriers).
// ------------------------------------------------------------------------------------------------------------
// (mean, rstd) per token row from the per-row (sum, sum of squares) partials the producing GEMM's epilogue left
// (GemmParams::rowstats_out): part [M][ld] float2, the first `slots` entries of a row summed in slot order.
// Replaces the statistics pass over the activation (ln_stats_kernel) when the producer is one of our GEMMs.
__global__ void __launch_bounds__(256) ln_rows_finalize_kernel(const float2* __restrict__ part, int ld, int slots, int M,
                                                               int C, float eps, float2* __restrict__ stats) {
  pdl_launch_dependents();
  pdl_wait();
  const int row = blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= M) return;
  const float2* pr = part + static_cast<size_t>(row) * ld;
  float su = 0.f, sq = 0.f;
  if ((slots & 1) == 0 && (ld & 1) == 0) {
    for (int i0 = 0; i0 < slots; i0 += 8) {  // four 16-byte loads (8 slots) in flight
      float4 v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u)
        v[u] = (i0 + 2 * u < slots) ? __ldg(reinterpret_cast<const float4*>(pr + i0) + u) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        if (i0 + 2 * u < slots) {
          su += v[u].x;
          sq += v[u].y;
          su += v[u].z;
          sq += v[u].w;
        }
      }
    }
  } else {
    for (int i = 0; i < slots; ++i) {
      const float2 t = __ldg(pr + i);
      su += t.x;
      sq += t.y;
    }
  }
  const float inv_c = 1.0f / static_cast<float>(C);
  const float mean = su * inv_c;
  stats[row] = make_float2(mean, rsqrtf(fmaxf(fmaf(-mean, mean, sq * inv_c), 0.f) + eps));
}

constexpr int kLnConsumerWarps = 8;
constexpr int kLnRowsPerWarp = 2;
constexpr int kLnTileRows = kLnConsumerWarps * kLnRowsPerWarp;  // 16
constexpr int kLnThreads = (kLnConsumerWarps + 1) * 32;

__device__ __forceinline__ void bulk_load_1d(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(src)), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// kStatsOnly: writes (mean, rstd) per row to `out` (reinterpreted as float2*) instead of the normalised rows: the
// LayerNorm itself is then folded into the consuming GEMM (gemm_tc.cuh, GemmParams::ln_stats).
template <int VPL, bool kStatsOnly>  // 8-channel vectors per lane: ceil(C / 256)
__global__ void __launch_bounds__(kLnThreads) layernorm_kernel(const __half* __restrict__ x, int M, int C,
                                                               const float* __restrict__ gamma,
                                                               const float* __restrict__ beta, float eps,
                                                               __half* __restrict__ out, int stages) {
  extern __shared__ __align__(128) uint8_t ln_smem[];
  const int tile_bytes = kLnTileRows * C * 2;
  float* gb = reinterpret_cast<float*>(ln_smem + static_cast<size_t>(stages) * tile_bytes);  // [2][C]
  uint64_t* full = reinterpret_cast<uint64_t*>(gb + 2 * C);
  uint64_t* empty = full + stages;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  pdl_launch_dependents();
  if (threadIdx.x == 0) {
    for (int i = 0; i < stages; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], kLnConsumerWarps);
    }
    fence_barrier_init();
  }
  // weights are never written inside a forward: safe to read before pdl_wait
  if (!kStatsOnly) {
    for (int i = threadIdx.x; i < C; i += blockDim.x) {
      gb[i] = gamma[i];
      gb[C + i] = beta[i];
    }
  }
  __syncthreads();
  pdl_wait();
  const int ntiles = (M + kLnTileRows - 1) / kLnTileRows;
  if (warp == kLnConsumerWarps) {
    // ------------------------------- producer warp ---------------------------------------------------------
    int s = 0;
    uint32_t ph = 0;
    for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
      mbar_wait(&empty[s], ph ^ 1);
      if (lane == 0) {
        const int rows = min(kLnTileRows, M - t * kLnTileRows);
        const uint32_t bytes = static_cast<uint32_t>(rows) * C * 2;
        mbar_arrive_expect_tx(&full[s], bytes);
        bulk_load_1d(ln_smem + static_cast<size_t>(s) * tile_bytes,
                     x + static_cast<size_t>(t) * kLnTileRows * C, bytes, &full[s]);
      }
      __syncwarp();
      if (++s == stages) { s = 0; ph ^= 1; }
    }
    return;
  }
  // ------------------------------- consumer warps ------------------------------------------------------------
  const int nvec = C / 8;
  const float inv_c = 1.0f / static_cast<float>(C);
  int s = 0;
  uint32_t ph = 0;
  for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
    mbar_wait(&full[s], ph);
    const uint8_t* tile = ln_smem + static_cast<size_t>(s) * tile_bytes;
    const int r0 = warp * kLnRowsPerWarp;
    uint4 raw[kLnRowsPerWarp][VPL];
#pragma unroll
    for (int r = 0; r < kLnRowsPerWarp; ++r) {
#pragma unroll
      for (int v = 0; v < VPL; ++v) {
        const int vi = lane + v * 32;
        raw[r][v] = (vi < nvec) ? *reinterpret_cast<const uint4*>(tile + (static_cast<size_t>(r0 + r) * C + vi * 8) * 2)
                                : make_uint4(0u, 0u, 0u, 0u);
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty[s]);  // rows are in registers: the slot can be refilled
    if (++s == stages) { s = 0; ph ^= 1; }
    float mean[kLnRowsPerWarp], rstd[kLnRowsPerWarp];
#pragma unroll
    for (int r = 0; r < kLnRowsPerWarp; ++r) {
      float acc = 0.f;
#pragma unroll
      for (int v = 0; v < VPL; ++v) {
        float f[8];
        unpack8(raw[r][v], f);
        acc += ((f[0] + f[1]) + (f[2] + f[3])) + ((f[4] + f[5]) + (f[6] + f[7]));
      }
      mean[r] = acc;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
      for (int r = 0; r < kLnRowsPerWarp; ++r) mean[r] += __shfl_xor_sync(0xffffffffu, mean[r], o);
    }
#pragma unroll
    for (int r = 0; r < kLnRowsPerWarp; ++r) {
      mean[r] *= inv_c;
      float q = 0.f;
#pragma unroll
      for (int v = 0; v < VPL; ++v) {
        if (lane + v * 32 < nvec) {
          float f[8];
          unpack8(raw[r][v], f);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float d = f[j] - mean[r];
            q = fmaf(d, d, q);
          }
        }
      }
      rstd[r] = q;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
      for (int r = 0; r < kLnRowsPerWarp; ++r) rstd[r] += __shfl_xor_sync(0xffffffffu, rstd[r], o);
    }
#pragma unroll
    for (int r = 0; r < kLnRowsPerWarp; ++r) rstd[r] = rsqrtf(rstd[r] * inv_c + eps);
    const long long row_base = static_cast<long long>(t) * kLnTileRows + r0;
    if (kStatsOnly) {
      if (lane == 0) {
#pragma unroll
        for (int r = 0; r < kLnRowsPerWarp; ++r)
          if (row_base + r < M) reinterpret_cast<float2*>(out)[row_base + r] = make_float2(mean[r], rstd[r]);
      }
    }
#pragma unroll
    for (int v = 0; v < (kStatsOnly ? 0 : VPL); ++v) {
      const int vi = lane + v * 32;
      if (vi < nvec) {
        const float4 g0 = *reinterpret_cast<const float4*>(gb + vi * 8);
        const float4 g1 = *reinterpret_cast<const float4*>(gb + vi * 8 + 4);
        const float4 b0 = *reinterpret_cast<const float4*>(gb + C + vi * 8);
        const float4 b1 = *reinterpret_cast<const float4*>(gb + C + vi * 8 + 4);
        const float gg[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
        const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
        for (int r = 0; r < kLnRowsPerWarp; ++r) {
          if (row_base + r < M) {
            float f[8], y[8];
            unpack8(raw[r][v], f);
#pragma unroll
            for (int j = 0; j < 8; ++j) y[j] = (f[j] - mean[r]) * rstd[r] * gg[j] + bb[j];
            store8(out + static_cast<size_t>(row_base + r) * C + vi * 8, y);
          }
        }
      }
    }
  }
}

// Stats-only LayerNorm pass, read-only: (mean, rstd) per row -> stats[M]. One warp handles RPW rows per iteration with all
// of their 16-byte loads issued up front (no shared memory, no stores besides 8 bytes per row): a plain grid-stride
// kernel beats the TMA pipeline above here because nothing has to be written back.
template <int VPL, int RPW>
__global__ void __launch_bounds__(256) ln_stats_kernel(const __half* __restrict__ x, int M, int C, float eps,
                                                       float2* __restrict__ stats) {
  pdl_launch_dependents();
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const int nvec = C / 8;
  const float inv_c = 1.0f / static_cast<float>(C);
  const long long stride = static_cast<long long>(gridDim.x) * (blockDim.x >> 5) * RPW;
  for (long long row0 = (static_cast<long long>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5)) * RPW; row0 < M;
       row0 += stride) {
    uint4 raw[RPW][VPL];
#pragma unroll
    for (int r = 0; r < RPW; ++r) {
#pragma unroll
      for (int v = 0; v < VPL; ++v) {
        const int vi = lane + v * 32;
        raw[r][v] = (row0 + r < M && vi < nvec) ? ldg16(x + static_cast<size_t>(row0 + r) * C + vi * 8)
                                                : make_uint4(0u, 0u, 0u, 0u);
      }
    }
    float mean[RPW], var[RPW];
#pragma unroll
    for (int r = 0; r < RPW; ++r) {
      float acc = 0.f;
#pragma unroll
      for (int v = 0; v < VPL; ++v) {
        float f[8];
        unpack8(raw[r][v], f);
        acc += ((f[0] + f[1]) + (f[2] + f[3])) + ((f[4] + f[5]) + (f[6] + f[7]));
      }
      mean[r] = acc;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
      for (int r = 0; r < RPW; ++r) mean[r] += __shfl_xor_sync(0xffffffffu, mean[r], o);
    }
#pragma unroll
    for (int r = 0; r < RPW; ++r) {
      mean[r] *= inv_c;
      float q = 0.f;
#pragma unroll
      for (int v = 0; v < VPL; ++v) {
        if (lane + v * 32 < nvec) {
          float f[8];
          unpack8(raw[r][v], f);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float d = f[j] - mean[r];
            q = fmaf(d, d, q);
          }
        }
      }
      var[r] = q;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
      for (int r = 0; r < RPW; ++r) var[r] += __shfl_xor_sync(0xffffffffu, var[r], o);
    }
    if (lane == 0) {
#pragma unroll
      for (int r = 0; r < RPW; ++r)
        if (row0 + r < M) stats[row0 + r] = make_float2(mean[r], rsqrtf(var[r] * inv_c + eps));
    }
  }
}

// LayerNorm folding (run once per weight update): for a Linear W [rows, K] (fp16, GEMM layout) that consumes
// LayerNorm(gamma, beta): Wf[j, k] = fp16(W[j, k] * gamma[k]), s[j] = sum_k Wf[j, k], bf[j] = bias[j] + sum_k W[j, k] beta[k].
// One warp per output row.
__global__ void __launch_bounds__(256) ln_fold_kernel(const __half* __restrict__ w, int rows, int K,
                                                      const float* __restrict__ gamma, const float* __restrict__ beta,
                                                      const float* __restrict__ bias, __half* __restrict__ wf,
                                                      float* __restrict__ s_out, float* __restrict__ bf_out) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  float s = 0.f, bacc = 0.f;
  for (int k = lane; k < K; k += 32) {
    const float wv = __half2float(w[static_cast<size_t>(row) * K + k]);
    const __half h = __float2half_rn(wv * gamma[k]);
    wf[static_cast<size_t>(row) * K + k] = h;
    s += __half2float(h);
    bacc = fmaf(wv, beta[k], bacc);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s += __shfl_xor_sync(0xffffffffu, s, o);
    bacc += __shfl_xor_sync(0xffffffffu, bacc, o);
  }
  if (lane == 0) {
    s_out[row] = s;
    bf_out[row] = (bias != nullptr ? bias[row] : 0.f) + bacc;
  }
}

// ------------------------------------------------------------------------------------------------------------
// Boundary conversions
// ------------------------------------------------------------------------------------------------------------
// First conv (in_channels=9, openaimodel.py:539-545): gather the 3x3 neighbourhood of the NCHW fp32 input into
// an fp16 [M, kpad] matrix, k = tap*cin + c, zero padded, consumed by the GEMM as a Linear. One thread per 8 consecutive
// k (one 16-byte store; kpad % 8 == 0); consecutive threads walk k first, so a warp covers 2-3 pixels and its gathers of
// one (tap, channel) plane hit neighbouring addresses.
__global__ void im2col_nchw_f32_kernel(const float* __restrict__ x, int n_img, int cin, int H, int W, int kpad,
                                       __half* __restrict__ out) {
  pdl_launch_dependents();
  pdl_wait();
  const int kv = kpad >> 3;
  const size_t idx = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const size_t total = static_cast<size_t>(n_img) * H * W * kv;
  if (idx >= total) return;
  const int k0 = static_cast<int>(idx % kv) * 8;
  const size_t m = idx / kv;
  const int xw = static_cast<int>(m % W);
  const int yh = static_cast<int>((m / W) % H);
  const int n = static_cast<int>(m / (static_cast<size_t>(W) * H));
  float v[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int k = k0 + j;
    v[j] = 0.f;
    if (k < 9 * cin) {
      const int tap = k / cin, c = k - tap * cin;
      const int yy = yh + tap / 3 - 1, xx = xw + tap % 3 - 1;
      if (yy >= 0 && yy < H && xx >= 0 && xx < W) v[j] = __ldg(x + ((static_cast<size_t>(n) * cin + c) * H + yy) * W + xx);
    }
  }
  store8(out + m * kpad + k0, v);
}

// Upsample (openaimodel.py:108-116): F.interpolate(scale_factor=2, mode="nearest"), NHWC fp16, 8-channel vectors.
__global__ void upsample2x_nhwc_kernel(const __half* __restrict__ x, int n_img, int H, int W, int C,
                                       __half* __restrict__ out) {
  pdl_launch_dependents();
  pdl_wait();
  const int nvec = C / 8;
  const size_t idx = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const size_t total = static_cast<size_t>(n_img) * (2 * H) * (2 * W) * nvec;
  if (idx >= total) return;
  const int v = static_cast<int>(idx % nvec);
  const size_t m = idx / nvec;
  const int xo = static_cast<int>(m % (2 * W));
  const int yo = static_cast<int>((m / (2 * W)) % (2 * H));
  const int n = static_cast<int>(m / (static_cast<size_t>(4) * W * H));
  const uint4 val =
      *reinterpret_cast<const uint4*>(x + ((static_cast<size_t>(n) * H + yo / 2) * W + xo / 2) * C + v * 8);
  *reinterpret_cast<uint4*>(out + m * C + v * 8) = val;
}

// Upsample folded into its conv (gemm_tc.cuh header): w [O, 9*I] fp16 (k = (ky*3 + kx)*I + i) -> out [4][O][4*I],
// phase = py*2 + px, k = (a*2 + b)*I + i. Output row 2y + py of the upsampled image reads, through tap ky, upsampled row
// 2y + py + ky - 1 = source row y + floor((py + ky - 1) / 2):
//   py = 0: y - 1 (tap ky = 0) and y (taps ky = 1, 2);   py = 1: y (taps ky = 0, 1) and y + 1 (tap ky = 2)
// so source offset index a in {0, 1} (offset py - 1 + a) collects the taps with floor((py + ky - 1) / 2) == py - 1 + a;
// columns alike. The fp16 tap weights are summed in fp32 and rounded to fp16 once.
__global__ void upfold_weights_kernel(const __half* __restrict__ w, int O, int I, __half* __restrict__ out) {
  const size_t idx = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const size_t total = static_cast<size_t>(4) * O * 4 * I;
  if (idx >= total) return;
  const int i = static_cast<int>(idx % I);
  const int ab = static_cast<int>((idx / I) % 4);
  const int o = static_cast<int>((idx / (static_cast<size_t>(4) * I)) % O);
  const int phase = static_cast<int>(idx / (static_cast<size_t>(4) * I * O));
  const int py = phase >> 1, px = phase & 1, a = ab >> 1, b = ab & 1;
  float acc = 0.f;
  for (int ky = 0; ky < 3; ++ky) {
    const int sy = (py + ky - 1 + 2) / 2 - 1;  // floor((py + ky - 1) / 2) for arguments >= -2
    if (sy != py - 1 + a) continue;
    for (int kx = 0; kx < 3; ++kx) {
      const int sx = (px + kx - 1 + 2) / 2 - 1;
      if (sx != px - 1 + b) continue;
      acc += __half2float(w[static_cast<size_t>(o) * 9 * I + (ky * 3 + kx) * I + i]);
    }
  }
  out[idx] = __float2half_rn(acc);
}

// Multiview re-arranged self-attention (ldm/modules/multiview_attention.py:436-462, concat_target=True): every UNet
// batch row is a stitched [ref_i | target] canvas of hh x (2*side) tokens. The attention sequence of sample b is
// [target (taken from row 0), ref_1 .. ref_v], each block hh*side tokens. gather: dst[b, k, y, x] <- src; scatter:
// the attended sequence is written back with the target block broadcast to all v rows (:456-460).
__device__ __forceinline__ size_t mv_src_row(int bi, int k, int y, int x, int v, int hh, int side) {
  const int row_img = (k == 0) ? bi * v : bi * v + (k - 1);
  const int xx = (k == 0) ? side + x : x;
  return (static_cast<size_t>(row_img) * hh + y) * (2 * side) + xx;
}
__global__ void mv_gather_kernel(const __half* __restrict__ src, int ld_src, int ncols, int b, int v, int hh, int side,
                                 __half* __restrict__ dst) {
  pdl_launch_dependents();
  pdl_wait();
  const int nvec = ncols / 8;
  const size_t idx = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const size_t total = static_cast<size_t>(b) * (v + 1) * hh * side * nvec;
  if (idx >= total) return;
  const int vec = static_cast<int>(idx % nvec);
  size_t r = idx / nvec;
  const int x = static_cast<int>(r % side);
  r /= side;
  const int y = static_cast<int>(r % hh);
  r /= hh;
  const int k = static_cast<int>(r % (v + 1));
  const int bi = static_cast<int>(r / (v + 1));
  const uint4 val = *reinterpret_cast<const uint4*>(src + mv_src_row(bi, k, y, x, v, hh, side) * ld_src + vec * 8);
  *reinterpret_cast<uint4*>(dst + (idx / nvec) * ncols + vec * 8) = val;
}
__global__ void mv_scatter_kernel(const __half* __restrict__ src, int ncols, int b, int v, int hh, int side,
                                  __half* __restrict__ dst) {
  pdl_launch_dependents();
  pdl_wait();
  const int nvec = ncols / 8;
  const size_t idx = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const size_t total = static_cast<size_t>(b) * v * hh * 2 * side * nvec;
  if (idx >= total) return;
  const int vec = static_cast<int>(idx % nvec);
  size_t r = idx / nvec;
  const int xx = static_cast<int>(r % (2 * side));
  r /= 2 * side;
  const int y = static_cast<int>(r % hh);
  r /= hh;
  const int i = static_cast<int>(r % v);
  const int bi = static_cast<int>(r / v);
  const int k = (xx < side) ? i + 1 : 0;
  const int x = (xx < side) ? xx : xx - side;
  const size_t srow = ((static_cast<size_t>(bi) * (v + 1) + k) * hh + y) * side + x;
  *reinterpret_cast<uint4*>(dst + (idx / nvec) * ncols + vec * 8) =
      *reinterpret_cast<const uint4*>(src + srow * ncols + vec * 8);
}

// NVSUnetModel(use_sep=True) (inpainting_ldm/NVS_ldm.py:57-71,74-97): before every non-resampling block ONE learned
// separator column sep[c] is inserted between the left (reference) and right (target) halves of the stitched canvas,
// `cat([h[..., :W/2], sep, h[..., W/2:]], -1)`, and removed again after the block, `cat([h[..., :W/2], h[..., -W/2:]])`.
// NHWC: in [n, H, W, C] -> out [n, H, W + 1, C]; sep points at this source's slice of the (concatenated) token.
__global__ void sep_insert_nhwc_kernel(const __half* __restrict__ x, const float* __restrict__ sep, int n_img, int H,
                                       int W, int C, __half* __restrict__ out) {
  pdl_launch_dependents();
  pdl_wait();
  const int nvec = C / 8;
  const size_t idx = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const size_t total = static_cast<size_t>(n_img) * H * (W + 1) * nvec;
  if (idx >= total) return;
  const int v = static_cast<int>(idx % nvec);
  const size_t m = idx / nvec;
  const int xo = static_cast<int>(m % (W + 1));
  const size_t ny = m / (W + 1);
  uint4 val;
  if (xo == W / 2) {
    const float* sp = sep + v * 8;
    val = make_uint4(pack_half2(sp[0], sp[1]), pack_half2(sp[2], sp[3]), pack_half2(sp[4], sp[5]), pack_half2(sp[6], sp[7]));
  } else {
    const int xi = xo < W / 2 ? xo : xo - 1;
    val = *reinterpret_cast<const uint4*>(x + (ny * W + xi) * C + v * 8);
  }
  *reinterpret_cast<uint4*>(out + m * C + v * 8) = val;
}
// in [n, H, W1, C] -> out [n, H, W1 - 1, C]: drops column (W1 - 1) / 2
__global__ void sep_remove_nhwc_kernel(const __half* __restrict__ x, int n_img, int H, int W1, int C,
                                       __half* __restrict__ out) {
  pdl_launch_dependents();
  pdl_wait();
  const int nvec = C / 8;
  const int W = W1 - 1;
  const size_t idx = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const size_t total = static_cast<size_t>(n_img) * H * W * nvec;
  if (idx >= total) return;
  const int v = static_cast<int>(idx % nvec);
  const size_t m = idx / nvec;
  const int xo = static_cast<int>(m % W);
  const size_t ny = m / W;
  const int xi = xo < W / 2 ? xo : xo + 1;
  *reinterpret_cast<uint4*>(out + m * C + v * 8) = *reinterpret_cast<const uint4*>(x + (ny * W1 + xi) * C + v * 8);
}
// the same insertion on the NCHW fp32 UNet input (in_channels = 9 is not a multiple of 8)
__global__ void sep_insert_nchw_f32_kernel(const float* __restrict__ x, const float* __restrict__ sep, int n_img, int C,
                                           int H, int W, float* __restrict__ out) {
  pdl_launch_dependents();
  pdl_wait();
  const size_t idx = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const size_t total = static_cast<size_t>(n_img) * C * H * (W + 1);
  if (idx >= total) return;
  const int xo = static_cast<int>(idx % (W + 1));
  const size_t ncy = idx / (W + 1);
  const int c = static_cast<int>((ncy / H) % C);
  out[idx] = (xo == W / 2) ? sep[c] : x[ncy * W + (xo < W / 2 ? xo : xo - 1)];
}
// NVS input refinement (NVS_ldm.py:64-68): c_input [n, C, H, Wc] fp32 NCHW is added to the output of the input conv,
// either over the whole width or over the columns from x_off on. Staged as an NHWC fp16 [n, H, Wh, C] residual (zeros
// outside [x_off, x_off + Wc)) that the input conv's epilogue adds.
__global__ void cinput_nchw_to_nhwc_kernel(const float* __restrict__ x, int n_img, int C, int H, int Wc, int x_off,
                                           int Wh, __half* __restrict__ out) {
  const size_t idx = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const size_t total = static_cast<size_t>(n_img) * H * Wh * C;
  if (idx >= total) return;
  const int c = static_cast<int>(idx % C);
  const size_t m = idx / C;
  const int xo = static_cast<int>(m % Wh);
  const int y = static_cast<int>((m / Wh) % H);
  const int n = static_cast<int>(m / (static_cast<size_t>(Wh) * H));
  const int xi = xo - x_off;
  float v = 0.f;
  if (xi >= 0 && xi < Wc) v = x[((static_cast<size_t>(n) * C + c) * H + y) * Wc + xi];
  out[idx] = __float2half_rn(v);
}

// ------------------------------------------------------------------------------------------------------------
// First-stage decoder helpers (AutoencoderKL.decode: ldm/models/autoencoder.py:87-90; Decoder / AttnBlock:
// ldm/modules/diffusionmodules/model.py:153-204,547-653)
// ------------------------------------------------------------------------------------------------------------
// z [n, zc, H, W] fp32 NCHW (the sampler's latent) -> z * z_scale (decode_first_stage divides by scale_factor,
// ddpm.py:842) -> post_quant_conv (1x1, zc x e, fp32) -> im2col rows [n*H*W, kpad] fp16 of the 3x3 conv_in
// (k = tap * zc + c, zero outside the image: the conv pads post_quant_conv's OUTPUT with zeros).
constexpr int kVaeMaxZ = 8;
__global__ void vae_in_kernel(const float* __restrict__ z, int n_img, int e, int zc, int H, int W, float z_scale,
                              const float* __restrict__ pq_w, const float* __restrict__ pq_b, int kpad,
                              __half* __restrict__ out) {
  pdl_launch_dependents();
  pdl_wait();
  const size_t idx = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;  // (pixel, tap)
  const size_t total = static_cast<size_t>(n_img) * H * W * 9;
  if (idx >= total) return;
  const int tap = static_cast<int>(idx % 9);
  const size_t m = idx / 9;
  const int xw = static_cast<int>(m % W);
  const int yh = static_cast<int>((m / W) % H);
  const int n = static_cast<int>(m / (static_cast<size_t>(W) * H));
  const int yy = yh + tap / 3 - 1, xx = xw + tap % 3 - 1;
  __half* o = out + m * kpad + tap * zc;
  if (yy < 0 || yy >= H || xx < 0 || xx >= W) {
    for (int c = 0; c < zc; ++c) o[c] = __float2half_rn(0.f);
  } else {
    float zin[kVaeMaxZ];
    for (int i = 0; i < e; ++i) zin[i] = z[((static_cast<size_t>(n) * e + i) * H + yy) * W + xx] * z_scale;
    for (int c = 0; c < zc; ++c) {
      float acc = pq_b[c];
      for (int i = 0; i < e; ++i) acc = fmaf(pq_w[c * e + i], zin[i], acc);
      o[c] = __float2half_rn(acc);
    }
  }
  if (tap == 8)
    for (int k = 9 * zc; k < kpad; ++k) out[m * kpad + k] = __float2half_rn(0.f);
}

// Row softmax of an fp16 matrix in place (AttnBlock: softmax over the key axis of the scaled logits, model.py:185),
// fp32 statistics. One CTA of 256 threads per row; the row is read once into registers (T <= 256 * 8 * kSoftmaxVecs).
constexpr int kSoftmaxVecs = 8;  // rows of up to 16384 elements
__global__ void __launch_bounds__(256) softmax_rows_kernel(__half* __restrict__ x, int T, size_t ld) {
  pdl_launch_dependents();
  pdl_wait();
  __shared__ float red[8];
  __half* row = x + static_cast<size_t>(blockIdx.x) * ld;
  const int nvec = T / 8;
  uint4 raw[kSoftmaxVecs];
  float mx = -INFINITY;
#pragma unroll
  for (int v = 0; v < kSoftmaxVecs; ++v) {
    const int vi = threadIdx.x + v * 256;
    raw[v] = make_uint4(0xfc00fc00u, 0xfc00fc00u, 0xfc00fc00u, 0xfc00fc00u);  // -inf
    if (vi < nvec) {
      raw[v] = *reinterpret_cast<const uint4*>(row + vi * 8);
      float f[8];
      unpack8(raw[v], f);
#pragma unroll
      for (int j = 0; j < 8; ++j) mx = fmaxf(mx, f[j]);
    }
  }
  auto block_reduce = [&](float v, bool is_max) -> float {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float t = __shfl_xor_sync(0xffffffffu, v, o);
      v = is_max ? fmaxf(v, t) : v + t;
    }
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    float r = red[0];
#pragma unroll
    for (int w = 1; w < 8; ++w) r = is_max ? fmaxf(r, red[w]) : r + red[w];  // fixed order
    return r;
  };
  mx = block_reduce(mx, true);
  float sum = 0.f;
#pragma unroll
  for (int v = 0; v < kSoftmaxVecs; ++v) {
    if (threadIdx.x + v * 256 < nvec) {
      float f[8];
      unpack8(raw[v], f);
#pragma unroll
      for (int j = 0; j < 8; ++j) sum += __expf(f[j] - mx);
    }
  }
  sum = block_reduce(sum, false);
  const float inv = 1.0f / sum;
#pragma unroll
  for (int v = 0; v < kSoftmaxVecs; ++v) {
    const int vi = threadIdx.x + v * 256;
    if (vi < nvec) {
      float f[8];
      unpack8(raw[v], f);
#pragma unroll
      for (int j = 0; j < 8; ++j) f[j] = __expf(f[j] - mx) * inv;
      store8(row + vi * 8, f);
    }
  }
}

// out [C, T] = in [T, C]^T (fp16, in has row stride ld_in): the value matrix of the AttnBlock as the K-major "weight"
// operand of the P V GEMM. 32 x 32 tiles through shared memory.
__global__ void transpose_f16_kernel(const __half* __restrict__ in, int T, int C, size_t ld_in, __half* __restrict__ out) {
  pdl_launch_dependents();
  pdl_wait();
  __shared__ __half tile[32][33];
  const int t0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int t = t0 + i, c = c0 + threadIdx.x;
    tile[i][threadIdx.x] = (t < T && c < C) ? in[static_cast<size_t>(t) * ld_in + c] : __float2half_rn(0.f);
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int c = c0 + i, t = t0 + threadIdx.x;
    if (c < C && t < T) out[static_cast<size_t>(c) * T + t] = tile[threadIdx.x][i];
  }
}

// fp32 -> fp16 cast (context tokens)
__global__ void cast_f32_f16_kernel(const float* __restrict__ x, size_t n, __half* __restrict__ out) {
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) out[i] = __float2half_rn(x[i]);
}

// UNet output: NHWC fp16 [M, ld] (first cout columns) -> NCHW fp32 [n, cout, H, W]
__global__ void nhwc_f16_to_nchw_f32_kernel(const __half* __restrict__ x, int ld, int n_img, int cout, int H, int W,
                                            float* __restrict__ out) {
  pdl_launch_dependents();
  pdl_wait();
  const size_t idx = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const size_t total = static_cast<size_t>(n_img) * cout * H * W;
  if (idx >= total) return;
  const int xw = static_cast<int>(idx % W);
  const int yh = static_cast<int>((idx / W) % H);
  const int c = static_cast<int>((idx / (static_cast<size_t>(W) * H)) % cout);
  const int n = static_cast<int>(idx / (static_cast<size_t>(W) * H * cout));
  out[idx] = __half2float(x[((static_cast<size_t>(n) * H + yh) * W + xw) * ld + c]);
}
// generic NCHW fp32 -> NHWC fp16 and back (op-level API convenience)
__global__ void nchw_f32_to_nhwc_f16_kernel(const float* __restrict__ x, int n_img, int C, int H, int W,
                                            __half* __restrict__ out) {
  const size_t idx = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const size_t total = static_cast<size_t>(n_img) * C * H * W;
  if (idx >= total) return;
  const int c = static_cast<int>(idx % C);
  const size_t m = idx / C;
  const int xw = static_cast<int>(m % W);
  const int yh = static_cast<int>((m / W) % H);
  const int n = static_cast<int>(m / (static_cast<size_t>(W) * H));
  out[idx] = __float2half_rn(x[((static_cast<size_t>(n) * C + c) * H + yh) * W + xw]);
}

// ------------------------------------------------------------------------------------------------------------
// Small-M linear layers of the timestep path (time_embed, 22 x emb_layers; openaimodel.py:527-532,217-223).
// out[n, o] = bias[o] + sum_k act(in[n, k]) * W[o, k]; fp16 weights streamed once, fp32 activations/accumulate.
// One warp per kOut output features (4 for the big emb_layers block; 1 for the 1280-wide time_embed layers, which would
// otherwise run on 40 CTAs); handles up to kMaxSmallBatch batch rows per pass.
// ------------------------------------------------------------------------------------------------------------
constexpr int kMaxSmallBatch = 8;
constexpr int kSmallOutPerWarp = 4;
// dynamic smem: min(n_rows, 8) * K floats (the activation block, SiLU already applied when silu_in)
template <int kOut>
__global__ void __launch_bounds__(256) small_linear_kernel(const float* __restrict__ in, int ld_in, int n_rows, int K,
                                                           const __half* __restrict__ w,
                                                           const float* __restrict__ bias, int n_out, int silu_in,
                                                           int silu_out, float* __restrict__ out, int ld_out) {
  pdl_launch_dependents();
  pdl_wait();
  extern __shared__ float act[];  // [rows][K]
  const int lane = threadIdx.x & 31;
  constexpr int kSmallOutPerWarp = kOut;
  const int o0 = (blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * kSmallOutPerWarp;
  for (int r0 = 0; r0 < n_rows; r0 += kMaxSmallBatch) {
    const int rows = min(kMaxSmallBatch, n_rows - r0);
    __syncthreads();
    // stage the activation block: 16-byte loads, four in flight per thread (a scalar load -> SiLU -> store loop left
    // every CTA of the 630-CTA emb_layers launch waiting ~40 dependent L2 round trips before its first weight load)
    const int kq = K >> 2;  // float4 per row (K % 8 == 0)
    for (int i0 = threadIdx.x; i0 < rows * kq; i0 += 4 * blockDim.x) {
      float4 v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int i = i0 + u * blockDim.x;
        v[u] = (i < rows * kq) ? *reinterpret_cast<const float4*>(in + static_cast<size_t>(r0 + i / kq) * ld_in + (i % kq) * 4)
                               : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int i = i0 + u * blockDim.x;
        if (i < rows * kq) {
          if (silu_in) v[u] = make_float4(silu_f(v[u].x), silu_f(v[u].y), silu_f(v[u].z), silu_f(v[u].w));
          *reinterpret_cast<float4*>(act + static_cast<size_t>(i / kq) * K + (i % kq) * 4) = v[u];
        }
      }
    }
    __syncthreads();
    if (o0 < n_out) {
      float acc[kSmallOutPerWarp][kMaxSmallBatch];
#pragma unroll
      for (int u = 0; u < kSmallOutPerWarp; ++u)
#pragma unroll
        for (int r = 0; r < kMaxSmallBatch; ++r) acc[u][r] = 0.f;
#pragma unroll 4
      for (int k = lane * 8; k < K; k += 32 * 8) {
        float wv[kSmallOutPerWarp][8];
#pragma unroll
        for (int u = 0; u < kSmallOutPerWarp; ++u) {
          const int o = min(o0 + u, n_out - 1);
          load8(w + static_cast<size_t>(o) * K + k, wv[u]);
        }
#pragma unroll
        for (int r = 0; r < kMaxSmallBatch; ++r) {
          if (r < rows) {
            const float4 a0 = *reinterpret_cast<const float4*>(act + r * K + k);
            const float4 a1 = *reinterpret_cast<const float4*>(act + r * K + k + 4);
            const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
#pragma unroll
            for (int u = 0; u < kSmallOutPerWarp; ++u)
#pragma unroll
              for (int j = 0; j < 8; ++j) acc[u][r] += av[j] * wv[u][j];
          }
        }
      }
#pragma unroll
      for (int u = 0; u < kSmallOutPerWarp; ++u) {
#pragma unroll
        for (int r = 0; r < kMaxSmallBatch; ++r) {
          float a = acc[u][r];
#pragma unroll
          for (int sft = 16; sft > 0; sft >>= 1) a += __shfl_xor_sync(0xffffffffu, a, sft);
          if (lane == 0 && r < rows && o0 + u < n_out) {
            a += bias ? bias[o0 + u] : 0.f;
            if (silu_out) a = silu_f(a);
            out[static_cast<size_t>(r0 + r) * ld_out + o0 + u] = a;
          }
        }
      }
    }
  }
}

// timestep_embedding (util.py:154-174): [cos(t f_i) | sin(t f_i)], f_i = exp(-ln(1e4) i / half), fp32.
// `t` holds t_count entries; row b uses t[b % t_count] (the CFG-pair forward passes one timestep per canvas).
__global__ void timestep_embedding_kernel(const long long* __restrict__ t, int t_count, int n, int dim,
                                          float* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int half = dim / 2;
  if (i >= n * half) return;
  const int b = i / half, j = i % half;
  const float freq = expf(-logf(10000.0f) * static_cast<float>(j) / static_cast<float>(half));
  const float arg = static_cast<float>(t[b % t_count]) * freq;
  out[static_cast<size_t>(b) * dim + j] = cosf(arg);
  out[static_cast<size_t>(b) * dim + half + j] = sinf(arg);
  if ((dim & 1) && j == 0) out[static_cast<size_t>(b) * dim + dim - 1] = 0.f;
}

// ------------------------------------------------------------------------------------------------------------
// Fused classifier-free guidance + DDIM update (ldm/models/diffusion/ddim.py:343,359-381), fp32 state.
//   e = e_u + s (e_c - e_u);  pred_x0 = (x - sqrt(1-a_t) e) / sqrt(a_t)
//   x_prev = sqrt(a_prev) pred_x0 + sqrt(1 - a_prev - sigma^2) e + sigma * noise * temperature
// eps is the UNet output for the CFG-doubled batch [2B, ...] (uncond first), or [B, ...] when e_c == nullptr.
// ------------------------------------------------------------------------------------------------------------
__global__ void ddim_update_kernel(const float* __restrict__ x, const float* __restrict__ e_u,
                                   const float* __restrict__ e_c, const float* __restrict__ noise, float cfg,
                                   float a_t, float a_prev, float sigma, float sqrt_one_minus_at, float temperature,
                                   size_t n, float* __restrict__ x_prev, float* __restrict__ pred_x0) {
  pdl_launch_dependents();
  pdl_wait();
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float e = e_u[i];
  if (e_c != nullptr) e = e + cfg * (e_c[i] - e);
  const float p0 = (x[i] - sqrt_one_minus_at * e) / sqrtf(a_t);
  const float dir = sqrtf(1.0f - a_prev - sigma * sigma) * e;
  const float nz = noise != nullptr ? sigma * noise[i] * temperature : 0.f;
  pred_x0[i] = p0;
  x_prev[i] = sqrtf(a_prev) * p0 + dir + nz;
}

// Same update with the per-step scalars read from DEVICE memory, coef = {cfg, a_t, a_prev, sigma, sqrt(1 - a_t)}:
// the launch arguments do not change from step to step, so one captured CUDA graph serves all DDIM steps.
__global__ void ddim_update_dev_kernel(const float* x /* may alias x_prev */, const float* __restrict__ e_u,
                                       const float* __restrict__ e_c, const float* __restrict__ noise,
                                       const float* __restrict__ coef, float temperature, size_t n,
                                       float* x_prev, float* __restrict__ pred_x0) {
  pdl_launch_dependents();
  pdl_wait();
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float cfg = coef[0], a_t = coef[1], a_prev = coef[2], sigma = coef[3], sqrt_one_minus_at = coef[4];
  float e = e_u[i];
  if (e_c != nullptr) e = e + cfg * (e_c[i] - e);
  const float p0 = (x[i] - sqrt_one_minus_at * e) / sqrtf(a_t);
  const float dir = sqrtf(1.0f - a_prev - sigma * sigma) * e;
  const float nz = noise != nullptr ? sigma * noise[i] * temperature : 0.f;
  pred_x0[i] = p0;
  x_prev[i] = sqrtf(a_prev) * p0 + dir + nz;
}

// ------------------------------------------------------------------------------------------------------------
// Split-K reduction of gemm_conv_kernel partial tiles: out[m, n] = fp16( sum_s partial[s][m][n] + bias[n]
// + bias_img[img(m)][n] + residual[m][n] ), 8 columns per thread, fixed summation order s = 0, 1, 2 (bit-reproducible).
// ------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) splitk_reduce_kernel(const float* __restrict__ partial, int ksplit, size_t M,
                                                            int ncols, int rows_per_img, const float* __restrict__ bias,
                                                            const float* __restrict__ bias_img, int ld_bias_img,
                                                            const __half* __restrict__ residual, int ld_res,
                                                            __half* __restrict__ out, int ld_out) {
  pdl_launch_dependents();
  pdl_wait();
  const int nvec = ncols / 8;
  const size_t idx = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= M * nvec) return;
  const size_t m = idx / nvec;
  const int col = static_cast<int>(idx % nvec) * 8;
  float acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = 0.f;
  for (int s = 0; s < ksplit; ++s) {
    const float4* pp = reinterpret_cast<const float4*>(partial + (static_cast<size_t>(s) * M + m) * ncols + col);
    const float4 a = __ldg(pp), b = __ldg(pp + 1);
    acc[0] += a.x; acc[1] += a.y; acc[2] += a.z; acc[3] += a.w;
    acc[4] += b.x; acc[5] += b.y; acc[6] += b.z; acc[7] += b.w;
  }
  if (bias != nullptr) {
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] += __ldg(bias + col + j);
  }
  if (bias_img != nullptr) {
    const float* bi = bias_img + (m / rows_per_img) * ld_bias_img + col;
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] += __ldg(bi + j);
  }
  if (residual != nullptr) {
    float r[8];
    load8(residual + m * ld_res + col, r);
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] += r[j];
  }
  store8(out + m * ld_out + col, acc);
}

// ------------------------------------------------------------------------------------------------------------
// Weight repacking (fp32 PyTorch layouts -> fp16 GEMM layouts), run once per weight update.
// ------------------------------------------------------------------------------------------------------------
// conv OIHW fp32 -> [O][tap][I] fp16 (tap = ky*3+kx), row stride ldk (>= 9*I, zero padded)
__global__ void repack_conv_kernel(const float* __restrict__ w, int O, int I, int ldk, __half* __restrict__ out) {
  const size_t idx = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= static_cast<size_t>(O) * ldk) return;
  const int k = static_cast<int>(idx % ldk);
  const int o = static_cast<int>(idx / ldk);
  float v = 0.f;
  if (k < 9 * I) {
    const int tap = k / I, i = k % I;
    v = w[(static_cast<size_t>(o) * I + i) * 9 + tap];
  }
  out[idx] = __float2half_rn(v);
}
// GEGLU row order of the fused projection: groups of four rows (value_2k, value_2k+1, gate_2k, gate_2k+1), so that the
// epilogue finds the two values and the two gates of an output PAIR in adjacent accumulator columns (packed fp32 math,
// one fp16x2 result). Source rows: [0, O/2) values, [O/2, O) gates (attention.py:51-58: chunk(2, dim=-1)).
__host__ __device__ inline int geglu_row(int o, int O) {
  const int half = O / 2;
  const int j = (o < half) ? o : o - half;
  return 4 * (j >> 1) + (j & 1) + ((o < half) ? 0 : 2);
}
// linear [O, I] fp32 -> fp16 rows at dst_row0 (+ the GEGLU row order)
__global__ void repack_linear_kernel(const float* __restrict__ w, int O, int I, int geglu, int dst_row0,
                                     __half* __restrict__ out) {
  const size_t idx = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= static_cast<size_t>(O) * I) return;
  const int i = static_cast<int>(idx % I);
  const int o = static_cast<int>(idx / I);
  const int d = geglu ? geglu_row(o, O) : o;
  out[(static_cast<size_t>(dst_row0) + d) * I + i] = __float2half_rn(w[idx]);
}
__global__ void repack_bias_kernel(const float* __restrict__ b, int O, int geglu, float* __restrict__ out) {
  const int o = blockIdx.x * blockDim.x + threadIdx.x;
  if (o >= O) return;
  out[geglu ? geglu_row(o, O) : o] = b[o];
}

}  // namespace lr
