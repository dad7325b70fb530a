// Probe: L2 -> shared memory bandwidth of TMA tile loads on B200, with the access patterns of the conv kernel:
//  (a) activation boxes [64 ch, 10 px, 34 rows] out of an NHWC bf16 tensor with 192 channels per pixel (128 B rows, 384 B pitch)
//  (b) weight boxes [64, 64 rows] out of a dense [rows][64] bf16 matrix (contiguous 8 KB)
// Every CTA (one per SM, 200 KB smem ring) streams `iters` boxes from a working set that fits L2; no compute.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../sr_gan_fd_b200/csrc/ptx.cuh"
using namespace b200sr;
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

__global__ void __launch_bounds__(64, 1) stream_kernel(const __grid_constant__ CUtensorMap tm, int mode, int iters, int depth, int box_bytes,
                                                       int tiles_x, int tiles_y, int nimg, int rows_total, const void* bulk_src) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  const int slot = (box_bytes + 1023) & ~1023;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + depth * slot);
  if (threadIdx.x == 0) { for (int i = 0; i < depth; ++i) mbar_init(&full[i], 1); fence_barrier_init(); }
  __syncthreads();
  if (threadIdx.x == 0) {
    // keep `depth` boxes in flight: issue, then wait for the oldest before reusing its slot
    uint32_t phase_bits = 0;
    for (int i = 0; i < iters + depth; ++i) {
      const int s = i % depth;
      if (i >= depth) { mbar_wait(&full[s], (phase_bits >> s) & 1u); phase_bits ^= 1u << s; }
      if (i < iters) {
        const unsigned id = blockIdx.x * 7919u + i * 104729u;
        mbar_arrive_expect_tx(&full[s], box_bytes);
        if (mode == 0) {
          const int n = id % nimg, tx = (id / nimg) % tiles_x, ty = (id / (nimg * tiles_x)) % tiles_y, c = ((id >> 20) % 3) * 64;
          tma_load_4d(smem + s * slot, &tm, &full[s], c, tx * 8 - 1, ty * 32 - 1, n);
        } else if (mode == 1) {
          const int wrows = box_bytes / 128;
          tma_load_2d(smem + s * slot, &tm, &full[s], 0, (id % (rows_total / wrows)) * wrows);
        } else {
          const char* src = reinterpret_cast<const char*>(bulk_src) + (size_t)(id % (rows_total * 128 / box_bytes)) * box_bytes;
          asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                       :: "r"(smem_u32(smem + s * slot)), "l"(src), "r"(box_bytes), "r"(smem_u32(&full[s])) : "memory");
        }
      }
    }
  }
}

int main() {
  void* fn = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
  EncodeTiledFn enc = (EncodeTiledFn)fn;
  int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  int clk = 0; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  const int N = 16, H = 64, W = 64, C = 192;  // 25 MB dense buffer (L2 resident)
  __nv_bfloat16* act; cudaMalloc(&act, (size_t)N * H * W * C * 2); cudaMemset(act, 0, (size_t)N * H * W * C * 2);
  const int rows = 65536;  // 8 MB weight matrix
  __nv_bfloat16* wt; cudaMalloc(&wt, (size_t)rows * 128); cudaMemset(wt, 0, (size_t)rows * 128);
  CUtensorMap tmA, tmW;
  { cuuint64_t dims[4] = {C, W, H, N}; cuuint64_t st[3] = {C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
    cuuint32_t box[4] = {64, 10, 34, 1}, es[4] = {1, 1, 1, 1};
    enc(&tmA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, act, dims, st, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
        CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE); }
  { cuuint64_t dims[2] = {64, rows}; cuuint64_t st[1] = {128}; cuuint32_t box[2] = {64, 64}, es[2] = {1, 1};
    enc(&tmW, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, wt, dims, st, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
        CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE); }
  cudaFuncSetAttribute(stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  struct Case { int mode; int bytes; const char* name; };
  Case cases[] = {{0, 10 * 34 * 128, "act 4D [64,10,34]"}, {1, 32 * 128, "w 2D [64,32]"}, {1, 64 * 128, "w 2D [64,64]"}, {1, 96 * 128, "w 2D [64,96]"},
                  {1, 192 * 128, "w 2D [64,192]"}, {1, 256 * 128, "w 2D [64,256]"}, {2, 4096, "bulk 1D 4K"}, {2, 8192, "bulk 1D 8K"},
                  {2, 24576, "bulk 1D 24K"}, {2, 73728, "bulk 1D 72K"}};
  for (const Case& cs : cases) {
    CUtensorMap tm = tmA;
    if (cs.mode == 1) {
      cuuint64_t dims[2] = {64, rows}; cuuint64_t st[1] = {128}; cuuint32_t box[2] = {64, (cuuint32_t)(cs.bytes / 128)}, es[2] = {1, 1};
      enc(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, wt, dims, st, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
          CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    }
    for (int depth : {1, 2}) {
      const int iters = 1000;
      float ms = 0;
      for (int rep = 0; rep < 2; ++rep) {
        cudaEventRecord(e0);
        stream_kernel<<<sms, 64, 220 * 1024>>>(tm, cs.mode, iters, depth, cs.bytes, W / 8, H / 32, N, rows, wt);
        cudaEventRecord(e1);
        cudaError_t err = cudaDeviceSynchronize();
        if (err != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(err)); return 1; }
        cudaEventElapsedTime(&ms, e0, e1);
      }
      const double bytes = (double)sms * iters * cs.bytes;
      printf("%-20s %6d B, %d in flight: %7.1f GB/s per SM, %6.2f TB/s aggregate, %6.0f ns per op\n", cs.name, cs.bytes, depth,
             bytes / sms / (ms * 1e-3) / 1e9, bytes / (ms * 1e-3) / 1e12, ms * 1e6 / iters);
    }
  }
  return 0;
}
