// Probe: does a K-major SWIZZLE_128B UMMA descriptor whose start address is NOT 1024-byte aligned (row offset dx inside
// the swizzle atom) read the right rows, and does it need the descriptor's base_offset field?
// Tile: one TMA box [64 ch, 16 px, 18 rows] (row pitch 16 px = 2048 B).  View for tap (dy,dx): rows (y+dy)*16 + x+dx,
// x = 0..7, y = 0..15 -> start = base + (dy*16+dx)*128, SBO = 2048.  B = identity (N = 64), so D[m][n] = A[m][n].
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../sr_gan_fd_b200/csrc/ptx.cuh"
using namespace b200sr;

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

__global__ void __launch_bounds__(128, 1) probe(const __grid_constant__ CUtensorMap tm, float* out, int dy, int dx, int use_base_offset,
                                                int box_w) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  uint8_t* A = smem;                 // box_w*18 rows * 128 B
  uint8_t* Bt = smem + 40960;        // 64 rows x 128 B identity, swizzled
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 40960 + 8192);
  uint64_t* done = bar + 1;
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // identity B: element (n, k) = (n == k); row n at n*128, 16B chunk j stored at j ^ (n & 7)
  for (int i = threadIdx.x; i < 64 * 64; i += 128) {
    int n = i / 64, k = i % 64;
    int chunk = (k / 8) ^ (n & 7);
    reinterpret_cast<__nv_bfloat16*>(Bt + n * 128 + chunk * 16)[k % 8] = __float2bfloat16(n == k ? 1.f : 0.f);
  }
  fence_proxy_async_smem();
  if (threadIdx.x == 0) { mbar_init(bar, 1); mbar_init(done, 1); fence_barrier_init(); }
  if (warp == 0) tmem_alloc_imm<64>(slot);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem = *slot;
  if (threadIdx.x == 0) {
    mbar_arrive_expect_tx(bar, box_w * 18 * 128);
    tma_load_4d(A, &tm, bar, 0, 4 - 1, 8 - 1, 0);   // tile origin x0 = 4, y0 = 8 -> box starts at (3, 7)
  }
  mbar_wait(bar, 0);
  tcgen05_fence_after();
  if (threadIdx.x == 0) {
    const uint32_t pitch = box_w * 128;
    const uint32_t a_addr = smem_u32(A) + (dy * box_w + dx) * 128;
    const uint32_t idesc = make_idesc_bf16(128, 64, 0, 0);
    for (int ks = 0; ks < 4; ++ks) {
      uint64_t ad = make_smem_desc(a_addr + ks * 32, 16, pitch);
      if (use_base_offset) ad |= static_cast<uint64_t>((a_addr >> 7) & 7) << 49;
      uint64_t bd = make_smem_desc(smem_u32(Bt) + ks * 32, 16, 1024);
      umma_bf16_ss(tmem, ad, bd, idesc, ks > 0);
    }
    umma_commit(done);
  }
  mbar_wait(done, 0);
  tcgen05_fence_after();
  uint32_t r[32];
  for (int c0 = 0; c0 < 64; c0 += 32) {
    tmem_ld_32x32b_x32(tmem + (static_cast<uint32_t>(warp * 32) << 16) + c0, r);
    tmem_ld_wait();
    for (int i = 0; i < 32; ++i) out[(warp * 32 + lane) * 64 + c0 + i] = __uint_as_float(r[i]);
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc_imm<64>(tmem);
}

int main() {
  const int H = 40, W = 32, C = 64;
  std::vector<__nv_bfloat16> h(H * W * C);
  for (int y = 0; y < H; ++y) for (int x = 0; x < W; ++x) for (int c = 0; c < C; ++c)
    h[(y * W + x) * C + c] = __float2bfloat16((float)((y * 37 + x * 11 + c * 3) % 251) - 125.f);
  __nv_bfloat16* d; cudaMalloc(&d, h.size() * 2); cudaMemcpy(d, h.data(), h.size() * 2, cudaMemcpyHostToDevice);
  float* dout; cudaMalloc(&dout, 128 * 64 * 4);
  void* fn = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
  EncodeTiledFn enc = (EncodeTiledFn)fn;
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  for (int box_w : {16, 10}) {
    CUtensorMap tm;
    cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, 1};
    cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
    cuuint32_t box[4] = {64, (cuuint32_t)box_w, 18, 1};
    cuuint32_t es[4] = {1, 1, 1, 1};
    CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return 1; }
    for (int ubo = 0; ubo < 2; ++ubo) {
      int bad_total = 0;
      for (int dy = 0; dy < 3; ++dy) for (int dx = 0; dx < 3; ++dx) {
        cudaMemset(dout, 0, 128 * 64 * 4);
        probe<<<1, 128, 64 * 1024>>>(tm, dout, dy, dx, ubo, box_w);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("box_w %d base_offset %d tap (%d,%d): CUDA error %s\n", box_w, ubo, dy, dx, cudaGetErrorString(e)); return 2; }
        std::vector<float> o(128 * 64);
        cudaMemcpy(o.data(), dout, o.size() * 4, cudaMemcpyDeviceToHost);
        int bad = 0;
        for (int m = 0; m < 128; ++m) for (int c = 0; c < 64; ++c) {
          int y = 8 + (m >> 3) + dy - 1, x = 4 + (m & 7) + dx - 1;
          float ref = __bfloat162float(h[(y * W + x) * C + c]);
          if (o[m * 64 + c] != ref) ++bad;
        }
        bad_total += bad;
        printf("box_w %2d base_offset_field %d tap (%d,%d): %s (%d mismatches)\n", box_w, ubo, dy, dx, bad ? "BAD" : "ok", bad);
      }
      printf("== box_w %d, base_offset field %s: %s\n", box_w, ubo ? "set" : "zero", bad_total ? "FAILS" : "ALL TAPS EXACT");
    }
  }
  return 0;
}
