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

// fp32 [B, T, C] (optionally times the LeakyReLU' mask of u) -> transposed operand planes [C][B * Tp], one copy per
// shift s in [first_shift, first_shift + nshift): column b * Tp + q of row c of copy s holds x[b, q - pad + s, c], zero
// where that time step does not exist -- so the pad columns in front of every utterance and the rest of Tp behind it are
// written here too (no memset).  Why copies: the weight gradient's position-reduction GEMM for tap j needs x^T shifted
// by j - pad along its contiguous dimension, and neither a TMA coordinate nor a UMMA start address can carry a 2-byte
// offset.  One pass reads x once and writes every copy with aligned 16-byte stores:
//   block (256 threads) = 64 output columns x 32 channels of one utterance; it stages time steps
//   [64 j - 4, 64 j + 64) as fp16 hi / lo rows [channel][step] in shared memory (conflict-free strides), then thread
//   (channel, 8-column group) loads the 12-step window its five possible shifts share and funnel-shifts the words.
// grid (ceil(Tp / 64), C / 32, B).
constexpr int TS_COLS = 64, TS_HALO = 4, TS_STRIDE = 70;     // 68 staged steps per channel, row stride 35 words
__device__ __forceinline__ void ts_store_shifted(const uint32_t (&w)[6], int off, __half* dst) {
  uint32_t o[4];
  const int k0 = off >> 1;
#pragma unroll
  for (int k = 0; k < 4; ++k)
    o[k] = (off & 1) ? __funnelshift_r(w[k0 + k], w[k0 + k + 1], 16) : w[k0 + k];
  *reinterpret_cast<uint4*>(dst) = make_uint4(o[0], o[1], o[2], o[3]);
}
__global__ void __launch_bounds__(256)
transpose_shift_split_kernel(const float* __restrict__ x, const float* __restrict__ u_mask, int T, int C, int Tp, int pad,
                             int first_shift, int nshift, size_t ktot, size_t copy_stride, __half* __restrict__ hiT,
                             __half* __restrict__ loT) {
  __shared__ __align__(16) __half s_hi[32 * TS_STRIDE];
  __shared__ __align__(16) __half s_lo[32 * TS_STRIDE];
  const int b = blockIdx.z;
  const int q0 = blockIdx.x * TS_COLS, c0 = blockIdx.y * 32;
  const int lane = threadIdx.x & 31, wrp = threadIdx.x >> 5;
  // stage: step index i <-> time step q0 - TS_HALO + i
  for (int i = wrp; i < TS_COLS + TS_HALO; i += 8) {
    const int t = q0 - TS_HALO + i;
    float v = 0.0f;
    if (t >= 0 && t < T) {
      const size_t g = (static_cast<size_t>(b) * T + t) * C + c0 + lane;
      v = x[g];
      if (u_mask != nullptr && !(u_mask[g] > 0.0f)) v = __fmul_rn(v, 0.1f);
    }
    const __half h = __float2half_rn(v);
    s_hi[lane * TS_STRIDE + i] = h;
    s_lo[lane * TS_STRIDE + i] = __float2half_rn((v - __half2float(h)) * kSplitScale);
  }
  __syncthreads();
  const int ch = threadIdx.x >> 3, grp = threadIdx.x & 7;
  const int q = q0 + 8 * grp;                    // first output column of this thread's group
  if (q >= Tp) return;
  uint32_t wh[6], wl[6];
#pragma unroll
  for (int k = 0; k < 6; ++k) {
    wh[k] = *reinterpret_cast<const uint32_t*>(s_hi + ch * TS_STRIDE + 8 * grp + 2 * k);
    wl[k] = *reinterpret_cast<const uint32_t*>(s_lo + ch * TS_STRIDE + 8 * grp + 2 * k);
  }
  const size_t o = static_cast<size_t>(c0 + ch) * ktot + static_cast<size_t>(b) * Tp + q;
  for (int m = 0; m < nshift; ++m) {
    // column q + e <- time step q + e - pad + s = staged step 8 grp + e + (TS_HALO - pad + s)
    const int off = TS_HALO - pad + first_shift + m;           // in [0, 4] for |s| <= pad <= 2
    ts_store_shifted(wh, off, hiT + m * copy_stride + o);
    ts_store_shifted(wl, off, loT + m * copy_stride + o);
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
