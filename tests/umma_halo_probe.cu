// Hardware probe (bring-up aid, not product code): can ONE tcgen05.mma read its 128-row A operand out of a TMA-loaded
// HALO tile, i.e. with a start address that is not 1024-byte aligned (shift by whole 128-byte rows) and an 8-row-group
// stride (SBO) that is not a multiple of 1024 bytes?  That is what re-using one (bw+2) x (bh+2) input tile for all nine
// taps of a 3x3 convolution needs.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -I leftrefill_b200/csrc tests/umma_halo_probe.cu -o /tmp/probe
//
// A source: NHWC [H = 18][W = 10][C = 64] fp16 with A[y][x][c] = small integers (exact in fp16), loaded with ONE TMA box
// (64, 10, 18), 128B swizzle -> 180 smem rows of 128 B. B = 64 x 64 identity, so D[r][n] = the A element the MMA
// actually read for row r, column n. For tap (dy, dx) the expected row r = 8*yy + xx (tile 8 wide x 16 high) is source
// pixel (y = yy + dy + 1, x = xx + dx + 1).  Variants: base_offset field 0 / (start >> 7) & 7.
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <vector>

#include "ptx.cuh"

using namespace lr;

constexpr int HW = 10, HH = 18, C = 64;
constexpr int A_BYTES = HW * HH * 128;  // 23040
constexpr int A_SLOT = 24 * 1024;

struct Params {
  CUtensorMap tmA, tmB;
  float* d;        // [128][64]
  int dy, dx;
  int use_base_offset;
  int sbo;         // bytes between 8-row groups
};

__global__ void __launch_bounds__(128, 1) probe_kernel(const __grid_constant__ Params p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* a_s = smem;
  uint8_t* b_s = smem + A_SLOT;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + A_SLOT + 8192);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    fence_barrier_init();
  }
  if (warp == 0) {
    tmem_alloc(tmem_slot, 64);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (threadIdx.x == 0) {
    mbar_arrive_expect_tx(&bars[0], A_BYTES + 64 * 128);
    tma_load_3d(a_s, &p.tmA, &bars[0], 0, 0, 0);
    tma_load_2d(b_s, &p.tmB, &bars[0], 0, 0);
  }
  mbar_wait(&bars[0], 0);
  tc_fence_after();
  if (threadIdx.x == 0) {
    const uint32_t idesc = umma_idesc_f16(128, 64, 0);
    const uint32_t a_start = smem_u32(a_s) + ((p.dy + 1) * HW + (p.dx + 1)) * 128;
    for (int k = 0; k < 4; ++k) {
      uint64_t ad = umma_smem_desc_sw128(a_start + 32 * k, p.sbo, 16);
      if (p.use_base_offset) ad |= static_cast<uint64_t>((a_start >> 7) & 7) << 49;
      const uint64_t bd = umma_smem_desc_sw128(smem_u32(b_s) + 32 * k, 1024, 16);
      umma_f16(tmem_base, ad, bd, idesc, k != 0 ? 1u : 0u);
    }
    umma_commit(&bars[1]);
  }
  mbar_wait(&bars[1], 0);
  tc_fence_after();
  const int r = warp * 32 + lane;
  for (int c = 0; c < 64; c += 32) {
    uint32_t v[32];
    tmem_ld32(tmem_base + (static_cast<uint32_t>(warp * 32) << 16) + c, v);
    tmem_ld_wait();
    for (int j = 0; j < 32; ++j) p.d[r * 64 + c + j] = __uint_as_float(v[j]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 64);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
  void* fnp = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fnp, cudaEnableDefault, &q) != cudaSuccess || !fnp) {
    printf("no cuTensorMapEncodeTiled\n");
    return 1;
  }
  EncodeTiledFn enc = reinterpret_cast<EncodeTiledFn>(fnp);
  std::vector<__half> ha(HH * HW * C), hb(64 * 64);
  auto aval = [](int y, int x, int c) { return static_cast<float>(((y * HW + x) * 7 + c * 3) % 201 - 100); };
  for (int y = 0; y < HH; ++y)
    for (int x = 0; x < HW; ++x)
      for (int c = 0; c < C; ++c) ha[(y * HW + x) * C + c] = __float2half(aval(y, x, c));
  for (int n = 0; n < 64; ++n)
    for (int k = 0; k < 64; ++k) hb[n * 64 + k] = __float2half(n == k ? 1.f : 0.f);
  __half *da, *db;
  float* dd;
  cudaMalloc(&da, ha.size() * 2);
  cudaMalloc(&db, hb.size() * 2);
  cudaMalloc(&dd, 128 * 64 * 4);
  cudaMemcpy(da, ha.data(), ha.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(db, hb.data(), hb.size() * 2, cudaMemcpyHostToDevice);
  Params p;
  {
    cuuint64_t dims[3] = {C, HW, HH};
    cuuint64_t str[2] = {C * 2, static_cast<cuuint64_t>(C) * 2 * HW};
    cuuint32_t box[3] = {64, HW, HH};
    cuuint32_t es[3] = {1, 1, 1};
    CUresult r = enc(&p.tmA, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, da, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("encode A failed %d\n", (int)r); return 1; }
  }
  {
    cuuint64_t dims[2] = {64, 64};
    cuuint64_t str[1] = {64 * 2};
    cuuint32_t box[2] = {64, 64};
    cuuint32_t es[2] = {1, 1};
    CUresult r = enc(&p.tmB, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, db, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("encode B failed %d\n", (int)r); return 1; }
  }
  p.d = dd;
  p.sbo = HW * 128;
  const int smem = A_SLOT + 8192 + 64;
  cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem + 1024);
  std::vector<float> hd(128 * 64);
  int all_ok = 1;
  for (int ubo = 0; ubo < 2; ++ubo) {
    for (int dy = -1; dy <= 1; ++dy) {
      for (int dx = -1; dx <= 1; ++dx) {
        p.dy = dy;
        p.dx = dx;
        p.use_base_offset = ubo;
        cudaMemset(dd, 0, 128 * 64 * 4);
        probe_kernel<<<1, 128, smem>>>(p);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) {
          printf("base_offset=%d tap(%d,%d): CUDA error %s\n", ubo, dy, dx, cudaGetErrorString(e));
          return 2;
        }
        cudaMemcpy(hd.data(), dd, 128 * 64 * 4, cudaMemcpyDeviceToHost);
        int bad = 0, first = -1;
        for (int r = 0; r < 128; ++r)
          for (int n = 0; n < 64; ++n) {
            const float exp = aval(r / 8 + dy + 1, r % 8 + dx + 1, n);
            if (hd[r * 64 + n] != exp) {
              if (first < 0) first = r * 64 + n;
              ++bad;
            }
          }
        printf("base_offset_field=%d tap(dy=%d,dx=%d): %s (%d / 8192 wrong%s)\n", ubo, dy, dx, bad ? "MISMATCH" : "ok", bad,
               bad ? "" : "");
        if (bad && first >= 0)
          printf("   first wrong: row %d col %d got %g expected %g\n", first / 64, first % 64, hd[first],
                 aval((first / 64) / 8 + dy + 1, (first / 64) % 8 + dx + 1, first % 64));
        if (bad && ubo == 0) all_ok = 0;
      }
    }
  }
  printf("HALO PROBE (base_offset field 0): %s\n", all_ok ? "PASS - shifted / padded A descriptors read the halo tile correctly" : "FAIL");
  return 0;
}
