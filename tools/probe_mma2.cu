// Probe: the conv kernel's exact MMA issue pattern (taps x k-steps x two 128-pixel halves sharing each weight tile,
// one commit per weight stage) in isolation, to separate tensor-pipe cost from everything else in the chain kernel.
#include <cstdio>
#include "../sr_gan_fd_b200/csrc/ptx.cuh"
using namespace b200sr;

template <int TAPS, int KS, int HALVES>
__global__ void __launch_bounds__(128, 1) mma_probe(int n_cols, int stages, long long* out_cycles) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = raw + ((1024u - (smem_u32(raw) & 1023u)) & 1023u);
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 200 * 1024);
  uint64_t* scratch_bar = bar + 2;
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 1);
  for (int i = threadIdx.x; i < 200 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  fence_proxy_async_smem();
  if (threadIdx.x == 0) { mbar_init(bar, 1); mbar_init(scratch_bar, 1 << 20); fence_barrier_init(); }
  if (threadIdx.x < 32) tmem_alloc_imm<512>(slot);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem = *slot;
  if (threadIdx.x == 0) {
    const uint32_t idesc = make_idesc_bf16(128, n_cols, 0, 0);
    const uint32_t a0 = smem_u32(smem), b0 = smem_u32(smem + 96 * 1024);
    constexpr uint32_t hiA = smem_desc_hi(1280), hiB = smem_desc_hi(1024);
    const uint32_t b_dy = n_cols * 8;
    const long long t0 = clock64();
    for (int s = 0; s < stages; ++s) {
      const uint32_t a_lo = smem_desc_lo(a0 + (s & 1) * 44032, 16), b_lo = smem_desc_lo(b0 + (s % 3) * 24576, 16);
#pragma unroll
      for (int t = 0; t < TAPS; ++t) {
#pragma unroll
        for (int k = 0; k < KS; ++k) {
#pragma unroll
          for (int h = 0; h < HALVES; ++h)
            umma_bf16_ss_lohi2(tmem + h * 128, a_lo + ((t % 3) * 10 + t / 3) * 8 + k * 2 + h * 1280, hiA, b_lo + t * b_dy + k * 2, hiB, idesc, 1u);
        }
      }
      umma_commit(scratch_bar);
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

template <int TAPS, int KS, int HALVES>
void run(const char* name, int n, int sms, long long* d) {
  cudaFuncSetAttribute(mma_probe<TAPS, KS, HALVES>, cudaFuncAttributeMaxDynamicSharedMemorySize, 210 * 1024);
  const int stages = 256;
  long long cyc = 0;
  for (int it = 0; it < 2; ++it) {
    mma_probe<TAPS, KS, HALVES><<<sms, 128, 210 * 1024>>>(n, stages, d);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("CUDA error %s (%s)\n", cudaGetErrorString(e), name); return; }
    cudaMemcpy(&cyc, d, 8, cudaMemcpyDeviceToHost);
  }
  printf("%-44s N=%3d : %6.1f cycles per MMA\n", name, n, (double)cyc / (stages * TAPS * KS * HALVES));
}

int main() {
  long long* d; cudaMalloc(&d, 8);
  int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  run<9, 4, 2>("9 taps x 4 k-steps x 2 halves / commit", 32, sms, d);
  run<9, 2, 2>("9 taps x 2 k-steps x 2 halves / commit", 32, sms, d);
  run<3, 4, 2>("3 taps x 4 k-steps x 2 halves / commit", 64, sms, d);
  run<3, 4, 2>("3 taps x 4 k-steps x 2 halves / commit", 32, sms, d);
  run<3, 4, 1>("3 taps x 4 k-steps x 1 half   / commit", 64, sms, d);
  run<9, 4, 1>("9 taps x 4 k-steps x 1 half   / commit", 32, sms, d);
  run<3, 4, 2>("3 taps x 4 k-steps x 2 halves / commit", 96, sms, d);
  run<3, 4, 2>("3 taps x 4 k-steps x 2 halves / commit", 128, sms, d);
  return 0;
}
