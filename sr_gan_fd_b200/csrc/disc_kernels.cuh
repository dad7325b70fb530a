// Helpers of the U-Net discriminator path (BSRGAN/model.py:91-167 = Real_ESRGAN/model.py:29-105).
// Its ten convs run on conv3x3_chain_kernel / wgrad3x3_kernel (the 4x4 stride-2 convs as 3x3 convs over the pixel-UNSHUFFLED
// input: "U layout", [N, H/2, W/2, 4C] with channel (py*2+px)*C + c); these HBM-bound kernels do the rest: the bilinear x2
// upsampling (align_corners=False) with the U-Net skip addition folded in, its transpose for the backward pass with the
// LeakyReLU derivative folded in, and the two element-wise joins around conv2.
#pragma once
#include "ptx.cuh"

namespace b200sr {

// element offset of (n, y, x, channel 0) of a C-channel tensor over an (H, W) lattice kept in U layout
__device__ __forceinline__ long long u_layout_off(int n, int y, int x, int H, int W, int C) {
  return ((static_cast<long long>(n) * (H >> 1) + (y >> 1)) * (W >> 1) + (x >> 1)) * (4LL * C) + (((y & 1) << 1) | (x & 1)) * C;
}

// eight 16-bit values <-> fp32; F16 = false: bf16 (default), true: fp16 (the reference's autocast format)
template <bool F16>
__device__ __forceinline__ void cvt8_to_f32(const uint4 v, float (&f)[8]) {
  f[0] = lo16_to_f32(v.x, F16); f[1] = hi16_to_f32(v.x, F16); f[2] = lo16_to_f32(v.y, F16); f[3] = hi16_to_f32(v.y, F16);
  f[4] = lo16_to_f32(v.z, F16); f[5] = hi16_to_f32(v.z, F16); f[6] = lo16_to_f32(v.w, F16); f[7] = hi16_to_f32(v.w, F16);
}
template <bool F16>
__device__ __forceinline__ uint4 f32x8_to_16(const float (&f)[8]) {
  return make_uint4(pack_16x2(f[0], f[1], F16), pack_16x2(f[2], f[3], F16), pack_16x2(f[4], f[5], F16), pack_16x2(f[6], f[7], F16));
}

// x: [N, C, H, W] (any strides; fp32 / fp16 / bf16; C = 3, other widths take the generator's ingest kernel) -> [N*H*W, 64] bf16 = [hi(C) | lo(C) | hi(C) | 0...] (conv1's weights are
// packed [w_hi | w_hi | w_lo]).  One thread per pixel: the row is assembled in registers and leaves as eight 16-byte stores (the
// generator's ingest kernel stores element by element, fine for its 64 x 64 inputs, 0.55 ms for sixteen 256 x 256 images).
template <typename T, int C, bool F16>
__global__ void disc_ingest_input_kernel(const T* __restrict__ x, long long sn, long long sc, long long sh, long long sw, int N, int H, int W,
                                         __nv_bfloat16* __restrict__ out) {
  const long long pix = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (pix >= static_cast<long long>(N) * H * W) return;
  const int xw = static_cast<int>(pix % W);
  const int yh = static_cast<int>((pix / W) % H);
  const int n = static_cast<int>(pix / (static_cast<long long>(W) * H));
  unsigned short row[64];
#pragma unroll
  for (int c = 0; c < 64; ++c) row[c] = 0;
#pragma unroll
  for (int c = 0; c < C; ++c) {
    const float v = static_cast<float>(x[n * sn + c * sc + yh * sh + xw * sw]);
    unsigned short h, l;
    if (F16) {
      const __half hh = __float2half_rn(v);
      h = __half_as_ushort(hh); l = __half_as_ushort(__float2half_rn(v - __half2float(hh)));
    } else {
      const __nv_bfloat16 hh = __float2bfloat16_rn(v);
      h = __bfloat16_as_ushort(hh); l = __bfloat16_as_ushort(__float2bfloat16_rn(v - __bfloat162float(hh)));
    }
    row[c] = h; row[C + c] = l; row[2 * C + c] = h;
  }
  uint4* o = reinterpret_cast<uint4*>(out + pix * 64);
  const uint4* r = reinterpret_cast<const uint4*>(row);
#pragma unroll
  for (int k = 0; k < 8; ++k) o[k] = r[k];
}

// F.interpolate(scale_factor=2, mode="bilinear", align_corners=False) of s = in (+ skip), BSRGAN/model.py:150-159.
// Output row 2Y reads rows (Y-1, Y) with weights (0.25, 0.75), row 2Y+1 rows (Y, Y+1) with (0.75, 0.25); indices clamp at the
// border (so the first / last output row copies the first / last input row).  in: [N, h, w, C] bf16; skip_u: nullptr or the
// tensor added to `in` first, in U layout; out: [N, 2h, 2w, C].  One thread per INPUT pixel and 8 channels: it reads the 3 x 3
// source neighbourhood once (18 loads with the skip) and writes the 2 x 2 output pixels that sit on top of its pixel.
template <bool F16>
__global__ void __launch_bounds__(256, 4)
disc_bilinear_up_kernel(const __nv_bfloat16* __restrict__ in, const __nv_bfloat16* __restrict__ skip_u,
                        __nv_bfloat16* __restrict__ out, int N, int h, int w, int C) {
  // grid: (ceil(w * C/8 / 256), h, N) -- 32-bit index math only
  const int c8 = C >> 3;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= w * c8) return;
  const int X = t / c8, cc = t - X * c8;
  const int Y = blockIdx.y, n = blockIdx.z;
  const int xs[3] = {max(X - 1, 0), X, min(X + 1, w - 1)};
  const long long orow = (static_cast<long long>(n) * (2 * h) + 2 * Y) * (2 * w) + 2 * X;  // output pixel (2Y, 2X)
  float top[2][8], bot[2][8];  // output rows 2Y / 2Y+1 (two columns each), accumulated source row by source row
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    const int sy = (a == 0) ? max(Y - 1, 0) : (a == 1 ? Y : min(Y + 1, h - 1));
    float s[3][8];
#pragma unroll
    for (int b = 0; b < 3; ++b) {
      cvt8_to_f32<F16>(__ldg(reinterpret_cast<const uint4*>(in + ((static_cast<long long>(n) * h + sy) * w + xs[b]) * C) + cc), s[b]);
      if (skip_u) {
        float g[8];
        cvt8_to_f32<F16>(__ldg(reinterpret_cast<const uint4*>(skip_u + u_layout_off(n, sy, xs[b], h, w, C)) + cc), g);
#pragma unroll
        for (int k = 0; k < 8; ++k) s[b][k] += g[k];
      }
    }
    // horizontal pass: column 2X = 0.25 s[X-1] + 0.75 s[X], column 2X+1 = 0.75 s[X] + 0.25 s[X+1] (clamped neighbours repeat the edge)
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const float h0 = 0.25f * s[0][k] + 0.75f * s[1][k], h1 = 0.75f * s[1][k] + 0.25f * s[2][k];
      if (a == 0) { top[0][k] = 0.25f * h0; top[1][k] = 0.25f * h1; }
      else if (a == 1) { top[0][k] += 0.75f * h0; top[1][k] += 0.75f * h1; bot[0][k] = 0.75f * h0; bot[1][k] = 0.75f * h1; }
      else { bot[0][k] += 0.25f * h0; bot[1][k] += 0.25f * h1; }
    }
    if (a == 1) {
      reinterpret_cast<uint4*>(out + orow * C)[cc] = f32x8_to_16<F16>(top[0]);
      reinterpret_cast<uint4*>(out + (orow + 1) * C)[cc] = f32x8_to_16<F16>(top[1]);
    }
  }
  reinterpret_cast<uint4*>(out + (orow + 2 * w) * C)[cc] = f32x8_to_16<F16>(bot[0]);
  reinterpret_cast<uint4*>(out + (orow + 2 * w + 1) * C)[cc] = f32x8_to_16<F16>(bot[1]);
}

// Transpose of the above: gin[Y, X] = sum over the (up to) 4 x 4 output positions that read input (Y, X).  Per axis the output
// indices 2Y-1, 2Y, 2Y+1, 2Y+2 contribute with weights 0.25 (Y >= 1), 0.75 (1.0 when Y == 0), 0.75 (1.0 when Y == h-1),
// 0.25 (Y <= h-2).  gout: [N, 2h, 2w, C] bf16.  gs_out (or nullptr): the sum itself = gradient w.r.t. the upsampled tensor
// (in + skip), which is also the skip connection's gradient; ga_out (or nullptr): the sum times the LeakyReLU(0.2) derivative
// taken from the saved activation `act` ([N, h, w, C], same lattice) = gradient w.r.t. the producing conv's pre-activation.
template <bool F16>
__global__ void __launch_bounds__(256, 4)
disc_bilinear_bwd_kernel(const __nv_bfloat16* __restrict__ gout, const __nv_bfloat16* __restrict__ act,
                         __nv_bfloat16* __restrict__ gs_out, __nv_bfloat16* __restrict__ ga_out, int N, int h, int w, int C) {
  // grid: (ceil(w * C/8 / 256), h, N) -- 32-bit index math only
  const int c8 = C >> 3;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= w * c8) return;
  const int X = t / c8, cc = t - X * c8;
  const int Y = blockIdx.y, n = blockIdx.z;
  const long long ip = (static_cast<long long>(n) * h + Y) * w + X;
  const int Ho = 2 * h, Wo = 2 * w;
  float wy[4], wx[4];
  wy[0] = (Y >= 1) ? 0.25f : 0.f; wy[1] = (Y == 0) ? 1.f : 0.75f; wy[2] = (Y == h - 1) ? 1.f : 0.75f; wy[3] = (Y <= h - 2) ? 0.25f : 0.f;
  wx[0] = (X >= 1) ? 0.25f : 0.f; wx[1] = (X == 0) ? 1.f : 0.75f; wx[2] = (X == w - 1) ? 1.f : 0.75f; wx[3] = (X <= w - 2) ? 0.25f : 0.f;
  float acc[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) acc[k] = 0.f;
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    const int oy = 2 * Y - 1 + a;
    if (wy[a] == 0.f) continue;
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      const int ox = 2 * X - 1 + b;
      if (wx[b] == 0.f) continue;
      float f[8];
      cvt8_to_f32<F16>(__ldg(reinterpret_cast<const uint4*>(gout + ((static_cast<long long>(n) * Ho + oy) * Wo + ox) * C) + cc), f);
      const float wgt = wy[a] * wx[b];
#pragma unroll
      for (int k = 0; k < 8; ++k) acc[k] += wgt * f[k];
    }
  }
  if (gs_out) reinterpret_cast<uint4*>(gs_out + ip * C)[cc] = f32x8_to_16<F16>(acc);
  if (ga_out) {
    float m[8];
    cvt8_to_f32<F16>(__ldg(reinterpret_cast<const uint4*>(act + ip * C) + cc), m);
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[k] *= (m[k] > 0.f) ? 1.f : 0.2f;
    reinterpret_cast<uint4*>(ga_out + ip * C)[cc] = f32x8_to_16<F16>(acc);
  }
}

// out[N, H, W, C] = a (plain layout) + b_u (U layout): up3 + out1 in front of conv2 (BSRGAN/model.py:161)
template <bool F16>
__global__ void disc_add_u_kernel(const __nv_bfloat16* __restrict__ a, const __nv_bfloat16* __restrict__ b_u, __nv_bfloat16* __restrict__ out,
                                  int N, int H, int W, int C) {
  const int c8 = C >> 3;
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= static_cast<long long>(N) * H * W * c8) return;
  const int cc = static_cast<int>(i % c8);
  const long long p = i / c8;
  const int x = static_cast<int>(p % W), y = static_cast<int>((p / W) % H), n = static_cast<int>(p / (static_cast<long long>(W) * H));
  float f[8], g[8];
  cvt8_to_f32<F16>(__ldg(reinterpret_cast<const uint4*>(a + p * C) + cc), f);
  cvt8_to_f32<F16>(__ldg(reinterpret_cast<const uint4*>(b_u + u_layout_off(n, y, x, H, W, C)) + cc), g);
#pragma unroll
  for (int k = 0; k < 8; ++k) f[k] += g[k];
  reinterpret_cast<uint4*>(out + p * C)[cc] = f32x8_to_16<F16>(f);
}

// out = g * LeakyReLU'(act): the gradient of (up3 + out1) taken through up3's activation
template <bool F16>
__global__ void disc_lrelu_mask_kernel(const __nv_bfloat16* __restrict__ g, const __nv_bfloat16* __restrict__ act, __nv_bfloat16* __restrict__ out,
                                       long long n8) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n8) return;
  float f[8], m[8];
  cvt8_to_f32<F16>(__ldg(reinterpret_cast<const uint4*>(g) + i), f);
  cvt8_to_f32<F16>(__ldg(reinterpret_cast<const uint4*>(act) + i), m);
#pragma unroll
  for (int k = 0; k < 8; ++k) f[k] *= (m[k] > 0.f) ? 1.f : 0.2f;
  reinterpret_cast<uint4*>(out)[i] = f32x8_to_16<F16>(f);
}

// dy: [N, C, H, W] fp32 contiguous -> [N*H*W, out_stride] bf16, channels [C, 16) zero (upstream gradient of the logit map)
template <bool F16>
__global__ void disc_ingest_grad_kernel(const float* __restrict__ dy, int N, int C, int H, int W, __nv_bfloat16* __restrict__ out, int out_stride) {
  const long long pix = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long plane = static_cast<long long>(H) * W;
  if (pix >= static_cast<long long>(N) * plane) return;
  const long long n = pix / plane, r = pix - n * plane;
  float v[16];
#pragma unroll
  for (int c = 0; c < 16; ++c) v[c] = (c < C) ? dy[(n * C + c) * plane + r] : 0.f;
  uint4* o = reinterpret_cast<uint4*>(out + pix * out_stride);
  o[0] = make_uint4(pack_16x2(v[0], v[1], F16), pack_16x2(v[2], v[3], F16), pack_16x2(v[4], v[5], F16), pack_16x2(v[6], v[7], F16));
  o[1] = make_uint4(pack_16x2(v[8], v[9], F16), pack_16x2(v[10], v[11], F16), pack_16x2(v[12], v[13], F16), pack_16x2(v[14], v[15], F16));
}

// Staged weight gradient of a 4x4 stride-2 conv computed as a 3x3 conv over the U-layout input -> [co][c][4][4].
// src: [tap = dy*3+dx][co_pad/4][4C][4] fp32; kernel position (ky, kx) lives at tap (dy, dx) and phase (py, px) with
// ky = 2 dy + py - 1: ky 0,1,2,3 -> (dy, py) = (0,1), (1,0), (1,1), (2,0).  One thread per destination element.
__global__ void unpack_wgrad_down_kernel(const float* __restrict__ src, float* __restrict__ dst, int co, int C, int co_pad) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= static_cast<long long>(co) * C * 16) return;
  const int kx = static_cast<int>(i & 3), ky = static_cast<int>((i >> 2) & 3);
  const int c = static_cast<int>((i >> 4) % C), o = static_cast<int>((i >> 4) / C);
  const int dy = (ky + 1) >> 1, py = (ky + 1) & 1, dx = (kx + 1) >> 1, px = (kx + 1) & 1;
  const int cu = (py * 2 + px) * C + c;
  dst[i] = src[((static_cast<long long>(dy * 3 + dx) * (co_pad >> 2) + (o >> 2)) * (4LL * C) + cu) * 4 + (o & 3)];
}

}  // namespace b200sr
