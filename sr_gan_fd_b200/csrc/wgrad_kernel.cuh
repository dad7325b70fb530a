// Weight-gradient of the 3x3 convs as a tcgen05 GEMM whose reduction dimension is the PIXEL axis:
//
//   dW[tap][ci][co] = sum_{pixel} X[pixel + tap, ci] * dY[pixel, co]
//
// Both operands are NHWC bf16 tiles (rows = pixels, 64 channels = 128 B per row) brought in by TMA with the 128B
// swizzle, i.e. exactly the UMMA "MN-major" canonical layout (the GEMM K index walks smem rows).  One CTA owns one
// horizontal tap dx and a slice of the pixel tiles (split-K over pixels across CTAs); the three vertical taps are
// three accumulators fed from the same haloed X tile at +0/+1024/+2048 bytes.  M = 128 input channels (two 64-ch
// TMA boxes), N = up to 160 output-gradient channels -- for a dense block that is the concatenation
// [dY5|dY4|dY3|dY2|dY1] of all consumers of the input slice, so the X tile is read once for all five convs
// ("re-associated by input slice").  The accumulators (3 x N <= 480 TMEM columns) live in TMEM for the whole
// pixel loop and are flushed once with 16-byte vector reductions (red.global.add.v4.f32) into a [tap][ci][co] fp32
// staging tensor; a small unpack kernel transposes that to the OIHW gradient layout per gradient bucket.
#pragma once
#include "conv_kernel.cuh"

namespace b200sr {

constexpr int kWgBBytes = kTileH * kTileW * 128;  // 16384 B per 64-channel dY tile
constexpr int kWgThreads = 192;
constexpr int kWgMaxSeg = 5;

struct WgradSegment {
  int col_begin, col_end;  // accumulator columns of this conv's output channels (multiples of 16)
  float* out;              // gradient staging tensor, [tap][ci_total][co_pad] fp32 (co fastest -> 16 B vector reductions)
  int ci_total;            // input channels of that conv
  int ci0;                 // conv input channel of accumulator row 0
  int co_pad;              // output channels rounded up to a multiple of 4
};

struct WgradParams {
  int N, H, W;
  int tiles_x, tiles_y, num_tiles;
  int a_c0;      // first X channel (accumulator row 0); two 64-channel boxes are loaded
  int b_c0;      // first dY channel (accumulator column 0)
  int n_cols;    // UMMA N (multiple of 16, <= 160)
  int n_blocks;  // ceil(n_cols / 64) dY boxes per tile
  int num_stages;
  int num_seg;
  WgradSegment seg[kWgMaxSeg];
};

__host__ __device__ inline int wgrad_stage_bytes(int n_blocks) { return 2 * kABytes + n_blocks * kWgBBytes; }
__host__ inline int wgrad_smem_bytes(int n_blocks, int stages) {
  return stages * wgrad_stage_bytes(n_blocks) + 1024 + 256;
}
__host__ inline int wgrad_pick_stages(int n_blocks) {
  int s = (227 * 1024 - 1024 - 256) / wgrad_stage_bytes(n_blocks);
  return s > 4 ? 4 : s;
}

__global__ void __launch_bounds__(kWgThreads, 1)
wgrad3x3_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmDY,
                const WgradParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int S = p.num_stages;
  const int stage_bytes = wgrad_stage_bytes(p.n_blocks);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S * stage_bytes);
  uint64_t* full = bars;
  uint64_t* empty = bars + 4;
  uint64_t* done = bars + 8;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 9);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int dxi = blockIdx.y;  // horizontal tap owned by this CTA
  uint32_t tmem_cols = 32;
  while (tmem_cols < 3u * p.n_cols) tmem_cols <<= 1;

  if (warp == 0 && lane == 0) {
    prefetch_tensormap(&tmX);
    prefetch_tensormap(&tmDY);
    for (int s = 0; s < S; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    mbar_init(done, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, tmem_cols);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int tiles_per_img = p.tiles_x * p.tiles_y;
  const bool has_work = static_cast<int>(blockIdx.x) < p.num_tiles;

  if (warp == 0) {
    int s = 0;
    uint32_t ph = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
      const int n = tile / tiles_per_img;
      const int t2 = tile - n * tiles_per_img;
      const int ty = t2 / p.tiles_x;
      const int x0 = (t2 - ty * p.tiles_x) * kTileW;
      const int y0 = ty * kTileH;
      mbar_wait(&empty[s], ph ^ 1);
      if (elect_one_sync()) {
        uint8_t* st = smem + s * stage_bytes;
        mbar_arrive_expect_tx(&full[s], stage_bytes);
        tma_load_4d(st, &tmX, &full[s], p.a_c0, x0 + dxi - 1, y0 - 1, n);
        tma_load_4d(st + kABytes, &tmX, &full[s], p.a_c0 + 64, x0 + dxi - 1, y0 - 1, n);
        for (int j = 0; j < p.n_blocks; ++j)
          tma_load_4d(st + 2 * kABytes + j * kWgBBytes, &tmDY, &full[s], p.b_c0 + 64 * j, x0, y0, n);
      }
      __syncwarp();
      if (++s == S) { s = 0; ph ^= 1; }
    }
  } else if (warp == 1) {
    const uint32_t idesc = make_idesc_bf16(128, p.n_cols, 1, 1);  // both operands MN-major
    constexpr uint32_t kHi = smem_desc_hi(1024);
    int s = 0;
    uint32_t ph = 0;
    uint32_t acc = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
      mbar_wait(&full[s], ph);
      tcgen05_fence_after();
      if (elect_one_sync()) {
        const uint32_t a0 = smem_u32(smem + s * stage_bytes);
        const uint32_t a_lo = smem_desc_lo(a0, kABytes);                  // LBO = stride between the two 64-ch M blocks
        const uint32_t b_lo = smem_desc_lo(a0 + 2 * kABytes, kWgBBytes);  // LBO = stride between 64-ch N blocks
#pragma unroll
        for (int dyi = 0; dyi < 3; ++dyi) {
#pragma unroll
          for (int ks = 0; ks < 8; ++ks)  // 16 pixels (two 8-pixel patch rows) per UMMA
            umma_bf16_ss_lohi(tmem_base + dyi * p.n_cols, a_lo + dyi * 64 + ks * 128, b_lo + ks * 128, kHi, idesc,
                              (ks == 0) ? acc : 1u);
        }
        umma_commit(&empty[s]);
      }
      __syncwarp();
      acc = 1;
      if (++s == S) { s = 0; ph ^= 1; }
    }
    if (elect_one_sync()) umma_commit(done);
    __syncwarp();
  } else if (has_work) {
    // epilogue: flush the three accumulators with 16-byte fp32 vector reductions into [tap][ci][co] staging tensors
    const int q = warp & 3;
    const int m = q * 32 + lane;  // accumulator row = input channel a_c0 + m
    mbar_wait(done, 0);
    tcgen05_fence_after();
    for (int dyi = 0; dyi < 3; ++dyi) {
      const int tap = dyi * 3 + dxi;
      for (int c0 = 0; c0 < p.n_cols; c0 += 16) {
        uint32_t r[16];
        tmem_ld_32x32b_x16(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + dyi * p.n_cols + c0, r);
        tmem_ld_wait();
        int sgi = 0;
        while (sgi < p.num_seg - 1 && c0 >= p.seg[sgi].col_end) ++sgi;
        const WgradSegment& sg = p.seg[sgi];
        const int co = c0 - sg.col_begin;
        const int ci = sg.ci0 + m;
        if (c0 >= sg.col_begin && c0 < sg.col_end && ci < sg.ci_total) {
          float* dst = sg.out + (static_cast<long long>(tap) * sg.ci_total + ci) * sg.co_pad + co;
#pragma unroll
          for (int j = 0; j < 4; ++j)
            if (co + 4 * j < sg.co_pad)
              red_add_v4_f32(dst + 4 * j, __uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1]), __uint_as_float(r[4 * j + 2]),
                             __uint_as_float(r[4 * j + 3]));
        }
      }
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    tmem_dealloc(tmem_base, tmem_cols);
  }
}

}  // namespace b200sr
