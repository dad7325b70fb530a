// Probe: cycles per tcgen05.mma (kind::f16, M=128, K=16, SS operands, 128B swizzle) as a function of N, on B200.
// One CTA per SM issues `reps` MMAs back to back from one thread (operands resident in smem, no loads), then commits.
#include <cstdio>
#include <vector>
#include "../sr_gan_fd_b200/csrc/ptx.cuh"
using namespace b200sr;

template <int d_rot>
__global__ void __launch_bounds__(128, 1) mma_probe(int n_cols, int reps, int a_off_rows, int sbo_a, long long* out_cycles, int commits_per_16, int d_stride = 128) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 200 * 1024);
  uint64_t* scratch_bar = bar + 2;   // commits in the loop arrive here (count large, never waited on)
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 1);
  for (int i = threadIdx.x; i < 200 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;  // bf16 ~0.0078
  fence_proxy_async_smem();
  if (threadIdx.x == 0) { mbar_init(bar, 1); mbar_init(scratch_bar, 1 << 20); fence_barrier_init(); }
  if (threadIdx.x < 32) tmem_alloc_imm<512>(slot);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem = *slot;
  if (threadIdx.x == 0) {
    const uint32_t idesc = make_idesc_bf16(128, n_cols, 0, 0);
    const uint32_t a0 = smem_u32(smem), b0 = smem_u32(smem + 128 * 1024);
    const uint32_t hiA = smem_desc_hi(sbo_a), hiB = smem_desc_hi(1024);
    const long long t0 = clock64();
    const uint32_t a_lo = smem_desc_lo(a0 + a_off_rows * 128, 16), b_lo = smem_desc_lo(b0, 16);
    for (int r = 0; r < reps; r += 16) {
#pragma unroll
      for (int u = 0; u < 16; ++u)  // 4 taps x 4 k-steps, constant offsets -> one add per operand at most
        umma_bf16_ss_lohi2(tmem + (u % d_rot) * d_stride, a_lo + (u >> 2) * 80 + (u & 3) * 2, hiA, b_lo + (u >> 2) * 512 + (u & 3) * 2, hiB, idesc, 1u);
      for (int c = 0; c < commits_per_16; ++c) umma_commit(scratch_bar);
    }
    umma_commit(bar);
    mbar_wait(bar, 0);
    const long long t1 = clock64();
    if (blockIdx.x == 0) out_cycles[0] = t1 - t0;
  }
  tcgen05_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) { tcgen05_fence_after(); tmem_dealloc_imm<512>(tmem); }
}

int main() {
  long long* d; cudaMalloc(&d, 8);
  cudaFuncSetAttribute(mma_probe<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 210 * 1024);
  cudaFuncSetAttribute(mma_probe<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 210 * 1024);
  cudaFuncSetAttribute(mma_probe<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 210 * 1024);
  int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const int reps = 4096;
  for (int commits : {0, 2}) {
    for (int n : {32, 96, 192}) {
      long long cyc = 0;
      for (int it = 0; it < 2; ++it) {
        mma_probe<1><<<sms, 128, 210 * 1024>>>(n, reps, 0, 1280, d, commits);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("CUDA error %s (n=%d)\n", cudaGetErrorString(e), n); return 1; }
        cudaMemcpy(&cyc, d, 8, cudaMemcpyDeviceToHost);
      }
      const double per = (double)cyc / reps;
      printf("%2d commits per 16 MMAs, N=%3d : %6.1f cycles per MMA (ideal %5.1f)\n", commits, n, per, n / 2.0);
    }
  }
  // independent accumulators: rotate the destination among d_rot TMEM column blocks
  for (int rot : {1, 2, 4}) {
    for (int n : {16, 32, 64, 128}) {
      long long cyc = 0;
      for (int it = 0; it < 2; ++it) {
        if (rot == 1) mma_probe<1><<<sms, 128, 210 * 1024>>>(n, reps, 0, 1280, d, 0, 128);
        else if (rot == 2) mma_probe<2><<<sms, 128, 210 * 1024>>>(n, reps, 0, 1280, d, 0, 128);
        else mma_probe<4><<<sms, 128, 210 * 1024>>>(n, reps, 0, 1280, d, 0, 128);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("CUDA error %s (n=%d)\n", cudaGetErrorString(e), n); return 1; }
        cudaMemcpy(&cyc, d, 8, cudaMemcpyDeviceToHost);
      }
      printf("%d accumulators in rotation, N=%3d : %6.1f cycles per MMA (ideal %5.1f)\n", rot, n, (double)cyc / reps, n / 2.0);
    }
  }
  return 0;
}
