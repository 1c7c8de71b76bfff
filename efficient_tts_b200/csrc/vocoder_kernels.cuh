// Element-wise / small kernels of the HiFi-GAN generator path; the convolutions run on the
// tap-GEMM of gemm2_sm100.cuh.  Citations are relative to /root/reference/nntts/vocoders/hifigan_model.py.
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "path_kernels.cuh"

namespace efts {

// mel [B, C, T] fp32 (the layout the reference's Generator takes, :120) -> operand planes [B, T, C].
// 32 x 32 shared-memory tile transpose; grid (ceil(T/32), ceil(C/32), B), block (32, 8).
__global__ void voc_mel_planes_kernel(const float* __restrict__ mel, int C, int T, __half* __restrict__ hi,
                                      __half* __restrict__ lo, int* __restrict__ err_flag) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z;
  const int t0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  for (int r = threadIdx.y; r < 32; r += 8) {
    const int c = c0 + r, t = t0 + threadIdx.x;
    tile[r][threadIdx.x] = (c < C && t < T) ? mel[(static_cast<size_t>(b) * C + c) * T + t] : 0.0f;
  }
  __syncthreads();
  for (int r = threadIdx.y; r < 32; r += 8) {
    const int t = t0 + r, c = c0 + threadIdx.x;
    if (t < T && c < C) {
      const float v = tile[threadIdx.x][r];
      if (!(fabsf(v) <= 65504.0f) && err_flag != nullptr) atomicOr(err_flag, 8);
      const __half h = __float2half_rn(v);
      const size_t o = (static_cast<size_t>(b) * T + t) * C + c;
      hi[o] = h;
      lo[o] = __float2half_rn((v - __half2float(h)) * kSplitScale);
    }
  }
}

// xs = r0 + r1 + ... (left to right, :126-130); x = xs / num_kernels (:131).  Either the operand planes of
// LeakyReLU(0.1)(x) for the next transposed conv (:123), or fp32 LeakyReLU(0.01)(x) for conv_post (:132).
struct VocAvgArgs { const float* r[4]; int n; };
__global__ void voc_average_kernel(VocAvgArgs a, size_t n4, float slope, __half* __restrict__ hi, __half* __restrict__ lo,
                                   float* __restrict__ out_f, int* __restrict__ err_flag) {
  const float div = static_cast<float>(a.n);
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n4;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    float4 v = reinterpret_cast<const float4*>(a.r[0])[i];
    for (int k = 1; k < a.n; ++k) {
      const float4 u = reinterpret_cast<const float4*>(a.r[k])[i];
      v.x = __fadd_rn(v.x, u.x); v.y = __fadd_rn(v.y, u.y); v.z = __fadd_rn(v.z, u.z); v.w = __fadd_rn(v.w, u.w);
    }
    v.x = __fdiv_rn(v.x, div); v.y = __fdiv_rn(v.y, div); v.z = __fdiv_rn(v.z, div); v.w = __fdiv_rn(v.w, div);
    v.x = v.x > 0.0f ? v.x : __fmul_rn(v.x, slope); v.y = v.y > 0.0f ? v.y : __fmul_rn(v.y, slope);
    v.z = v.z > 0.0f ? v.z : __fmul_rn(v.z, slope); v.w = v.w > 0.0f ? v.w : __fmul_rn(v.w, slope);
    if (out_f != nullptr) reinterpret_cast<float4*>(out_f)[i] = v;
    if (hi != nullptr) {
      if (outside_fp16_range(v) && err_flag != nullptr)
        atomicOr(err_flag, 8);
      uint2 h, l;
      split4(v, &h, &l);
      reinterpret_cast<uint2*>(hi)[i] = h;
      reinterpret_cast<uint2*>(lo)[i] = l;
    }
  }
}

// conv_post (:133) + tanh (:134): x [B, L, C] fp32 (already LeakyReLU'd), w [taps][C], one output sample per
// thread: y[b, t] = tanh(bias + sum_j sum_c w[j][c] * x[b, t + j - pad, c]).  C multiple of 4, C * taps <= 1024.
__global__ void voc_post_kernel(const float* __restrict__ x, const float* __restrict__ w, float bias, int L, int C,
                                int taps, float* __restrict__ y) {
  __shared__ float sw[1024];
  for (int i = threadIdx.x; i < taps * C; i += blockDim.x) sw[i] = w[i];
  __syncthreads();
  const int b = blockIdx.y;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= L) return;
  const int pad = (taps - 1) / 2;
  float acc = 0.0f;
  for (int j = 0; j < taps; ++j) {
    const int s = t + j - pad;
    if (s < 0 || s >= L) continue;
    const float4* row = reinterpret_cast<const float4*>(x + (static_cast<size_t>(b) * L + s) * C);
    for (int c4 = 0; c4 < C / 4; ++c4) {
      const float4 v = row[c4];
      const float* ww = sw + j * C + c4 * 4;
      acc = fmaf(v.x, ww[0], acc); acc = fmaf(v.y, ww[1], acc); acc = fmaf(v.z, ww[2], acc); acc = fmaf(v.w, ww[3], acc);
    }
  }
  y[static_cast<size_t>(b) * L + t] = tanhf(acc + bias);
}

// The same arithmetic (same order: taps outer, channels inner) with the rows of a 256-sample block staged in shared
// memory by coalesced loads; C <= 32.  The one-row-per-thread global reads of voc_post_kernel cost 0.73 ms for
// 16 x 800 frames (420 MB, 10 x the HBM time); this form reads every row once.
constexpr int VOC_POST_BLOCK = 256;
constexpr int VOC_POST_CMAX = 32;
__global__ void __launch_bounds__(VOC_POST_BLOCK)
voc_post_tiled_kernel(const float* __restrict__ x, const float* __restrict__ w, float bias, int L, int C, int taps,
                      float* __restrict__ y) {
  __shared__ float sw[1024];
  __shared__ float sx[(VOC_POST_BLOCK + 16) * (VOC_POST_CMAX + 1)];
  for (int i = threadIdx.x; i < taps * C; i += blockDim.x) sw[i] = w[i];
  const int b = blockIdx.y;
  const int t0 = blockIdx.x * VOC_POST_BLOCK;
  const int pad = (taps - 1) / 2;
  const int rows = VOC_POST_BLOCK + taps - 1;           // rows t0 - pad .. t0 + 255 + pad
  const int ld = C + 1;                                 // odd stride: lanes (= consecutive rows) hit distinct banks
  const int c4n = C / 4;
  for (int i = threadIdx.x; i < rows * c4n; i += blockDim.x) {
    const int r = i / c4n, c4 = i - r * c4n;
    const int s = t0 - pad + r;
    float4 v = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    if (s >= 0 && s < L) v = __ldg(reinterpret_cast<const float4*>(x + (static_cast<size_t>(b) * L + s) * C) + c4);
    float* d = sx + r * ld + c4 * 4;
    d[0] = v.x; d[1] = v.y; d[2] = v.z; d[3] = v.w;
  }
  __syncthreads();
  const int t = t0 + threadIdx.x;
  if (t >= L) return;
  float acc = 0.0f;
  for (int j = 0; j < taps; ++j) {
    const int s = t + j - pad;
    if (s < 0 || s >= L) continue;
    const float* row = sx + (threadIdx.x + j) * ld;
    const float* ww = sw + j * C;
    for (int c = 0; c < C; ++c) acc = fmaf(row[c], ww[c], acc);
  }
  y[static_cast<size_t>(b) * L + t] = tanhf(acc + bias);
}

}  // namespace efts
