// Training slice: element-wise / layout kernels around the tensor-core GEMMs of the ResConv1d
// backward.  Forward of one layer (layers/efts_modules.py:48-51): u = lrelu_{0.1}(conv_k(x) + b), y = x + u.
// Backward, with g = dL/dy:   G' = g * (u > 0 ? 1 : 0.1)            (LeakyReLU'; torch takes the slope at pre == 0)
//                             dL/dx = g + conv_k^T(G')               -> tap-GEMM with flipped, transposed weights
//                             dL/dW[o, c, j] = sum_{b,t} G'[b,t,o] x[b, t + j - pad, c]   -> GEMM over positions
//                             dL/db[o] = sum_{b,t} G'[b,t,o]
// The position-reduction GEMM wants both operands K-major with K = positions: G'^T and x^T are laid out
// [C][B * Tp] with Tp = round8(T + 2 * pad) and `pad` zero columns in front of every utterance, so a tap is a
// column shift of x^T that never reads across an utterance boundary.  TMA box coordinates must be 16-byte aligned,
// so the shift cannot be a read offset of one or two fp16 elements: x^T is written once per tap with the shift
// applied on the store side.  (Known cost: k transposes of x per layer.  The better operand is the activation plane
// itself read as an MN-major UMMA operand, where a tap is a row shift of the shared-memory tile exactly like in the
// forward kernel; see DESIGN.md, "what comes next".)
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "path_kernels.cuh"

namespace efts {

// Conv1d weight fp32 [N, K, taps] (torch layout) -> operand planes [taps][N][K]; FLIP: the operator of the data
// gradient, Wf[j][c][o] = W[o][c][taps - 1 - j], as planes [taps][K][N].  One thread per element.
template <bool FLIP>
__global__ void pack_conv_weight_kernel(const float* __restrict__ w, int N, int K, int taps, __half* __restrict__ hi,
                                        __half* __restrict__ lo, int* __restrict__ err_flag) {
  const size_t n = static_cast<size_t>(N) * K * taps;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    // i enumerates the destination: [tap][row][col]
    const int col = static_cast<int>(i % (FLIP ? N : K));
    const int row = static_cast<int>((i / (FLIP ? N : K)) % (FLIP ? K : N));
    const int tap = static_cast<int>(i / (static_cast<size_t>(N) * K));
    const int o = FLIP ? col : row, c = FLIP ? row : col, j = FLIP ? taps - 1 - tap : tap;
    const float x = w[(static_cast<size_t>(o) * K + c) * taps + j];
    if (!(fabsf(x) <= 65504.0f)) atomicOr(err_flag, 8);
    const __half h = __float2half_rn(x);
    hi[i] = h;
    lo[i] = __float2half_rn((x - __half2float(h)) * kSplitScale);
  }
}

// y = x + u (fp32, the residual add of the layer) + operand planes of y.  n4 = elements / 4.
__global__ void residual_add_split_kernel(const float* __restrict__ x, const float* __restrict__ u, size_t n4,
                                          float* __restrict__ y, __half* __restrict__ hi, __half* __restrict__ lo,
                                          int* __restrict__ err_flag) {
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n4;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(x) + i);
    const float4 b = __ldg(reinterpret_cast<const float4*>(u) + i);
    float4 v;
    v.x = __fadd_rn(a.x, b.x); v.y = __fadd_rn(a.y, b.y); v.z = __fadd_rn(a.z, b.z); v.w = __fadd_rn(a.w, b.w);
    reinterpret_cast<float4*>(y)[i] = v;
    if (outside_fp16_range(v)) atomicOr(err_flag, 8);
    uint2 h, l;
    split4(v, &h, &l);
    reinterpret_cast<uint2*>(hi)[i] = h;
    reinterpret_cast<uint2*>(lo)[i] = l;
  }
}

// G' = g * (u > 0 ? 1 : 0.1) as operand planes [B, T, C] (A operand of the data-gradient GEMM).
__global__ void lrelu_grad_split_kernel(const float* __restrict__ g, const float* __restrict__ u, size_t n4,
                                        __half* __restrict__ hi, __half* __restrict__ lo, int* __restrict__ err_flag) {
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n4;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(g) + i);
    const float4 b = __ldg(reinterpret_cast<const float4*>(u) + i);
    float4 v;
    v.x = b.x > 0.0f ? a.x : __fmul_rn(a.x, 0.1f); v.y = b.y > 0.0f ? a.y : __fmul_rn(a.y, 0.1f);
    v.z = b.z > 0.0f ? a.z : __fmul_rn(a.z, 0.1f); v.w = b.w > 0.0f ? a.w : __fmul_rn(a.w, 0.1f);
    if (outside_fp16_range(v)) atomicOr(err_flag, 8);
    uint2 h, l;
    split4(v, &h, &l);
    reinterpret_cast<uint2*>(hi)[i] = h;
    reinterpret_cast<uint2*>(lo)[i] = l;
  }
}

// fp32 [B, T, C] (optionally times the LeakyReLU' mask of u) -> transposed operand planes [C][B * Tp], element
// (b, t, c) at column b * Tp + pad + t - shift of row c (|shift| <= pad).  The margins (pad columns in front of every utterance, the rest of
// Tp behind it) must be zero: the buffer is cleared by the caller once.  32 x 32 tiles through shared memory; grid
// (ceil(T / 32), C / 32, B), block (32, 8).
__global__ void transpose_split_kernel(const float* __restrict__ x, const float* __restrict__ u_mask, int T, int C, int Tp,
                                       int pad, int shift, size_t ktot, __half* __restrict__ hiT, __half* __restrict__ loT) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z;
  const int t0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  for (int r = threadIdx.y; r < 32; r += 8) {
    const int t = t0 + r;
    float v = 0.0f;
    if (t < T) {
      const size_t i = (static_cast<size_t>(b) * T + t) * C + c0 + threadIdx.x;
      v = x[i];
      if (u_mask != nullptr && !(u_mask[i] > 0.0f)) v = __fmul_rn(v, 0.1f);
    }
    tile[r][threadIdx.x] = v;
  }
  __syncthreads();
  for (int r = threadIdx.y; r < 32; r += 8) {
    const int t = t0 + threadIdx.x;
    if (t < T) {
      const float v = tile[threadIdx.x][r];
      const __half h = __float2half_rn(v);
      const size_t o = static_cast<size_t>(c0 + r) * ktot + static_cast<size_t>(b) * Tp + pad + t - shift;
      hiT[o] = h;
      loT[o] = __float2half_rn((v - __half2float(h)) * kSplitScale);
    }
  }
}

// db[c] = sum over rows of G'[row, c] (G' = g * LeakyReLU'(u)), fp32 result, double accumulation, deterministic:
// block j sums rows j, j + gridDim.x, ... for 128 columns (blockIdx.y picks the column group) into part[j][C];
// bias_grad_finish_kernel adds the partials in block order.
__global__ void bias_grad_partial_kernel(const float* __restrict__ g, const float* __restrict__ u, size_t rows, int C,
                                         double* __restrict__ part) {
  const int c = blockIdx.y * 128 + threadIdx.x;
  if (c >= C) return;
  double acc = 0.0;
  for (size_t r = blockIdx.x; r < rows; r += gridDim.x) {
    const float a = g[r * C + c];
    acc += static_cast<double>(u[r * C + c] > 0.0f ? a : __fmul_rn(a, 0.1f));
  }
  part[static_cast<size_t>(blockIdx.x) * C + c] = acc;
}
__global__ void bias_grad_finish_kernel(const double* __restrict__ part, int nparts, int C, float* __restrict__ db) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  double acc = 0.0;
  for (int j = 0; j < nparts; ++j) acc += part[static_cast<size_t>(j) * C + c];
  db[c] = static_cast<float>(acc);
}

// dWt [taps][N][K] fp32 (what the position-reduction GEMMs write) -> torch layout [N][K][taps].
__global__ void weight_grad_permute_kernel(const float* __restrict__ dwt, int N, int K, int taps, float* __restrict__ dw) {
  const size_t n = static_cast<size_t>(N) * K * taps;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int j = static_cast<int>(i % taps);
    const size_t ok = i / taps;                        // o * K + c
    dw[i] = dwt[static_cast<size_t>(j) * N * K + ok];
  }
}

}  // namespace efts
