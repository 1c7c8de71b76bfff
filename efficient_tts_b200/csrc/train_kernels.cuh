// Training slice: element-wise / layout kernels around the tensor-core GEMMs of the ResConv1d
// backward.  Forward of one layer (layers/efts_modules.py:48-51): u = lrelu_{0.1}(conv_k(x) + b), y = x + u.
// Backward, with g = dL/dy:   G' = g * (u > 0 ? 1 : 0.1)            (LeakyReLU'; torch takes the slope at pre == 0)
//                             dL/dx = g + conv_k^T(G')               -> tap-GEMM with flipped, transposed weights
//                             dL/dW[o, c, j] = sum_{b,t} G'[b,t,o] x[b, t + j - pad, c]   -> GEMM over positions
//                             dL/db[o] = sum_{b,t} G'[b,t,o]
// The position-reduction GEMM reads BOTH operands from the activations' own operand planes [B, T, C] as MN-major UMMA
// operands (gemm2_kernel's BMN = 2 instantiation): A = G' (M = its channels), B = the layer input (N = its channels), the
// reduction index is the row of the 64-row TMA boxes, utterance by utterance; a conv tap is a row offset of B's box and
// rows outside the utterance are zero-filled by the TMA unit -- exactly the forward kernel's treatment of taps.  No
// transposed copy of either tensor exists.  (Earlier forms of this slice transposed G' and wrote k shifted transposes of
// x per layer: a K-major operand cannot be shifted by one element, TMA coordinates and UMMA start addresses being
// 16-byte aligned.)
// Further down: the duration predictor's LayerNorm / head kernels and the criterion with gradients.
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

// G' = g * (u > 0 ? 1 : slope) as operand planes [B, T, C] (A operand of the data-gradient GEMM); slope = 0.1 for the
// LeakyReLU of the ResConv layers, 0 for the ReLU of the duration predictor.
__global__ void lrelu_grad_split_kernel(const float* __restrict__ g, const float* __restrict__ u, size_t n4, float slope,
                                        __half* __restrict__ hi, __half* __restrict__ lo, int* __restrict__ err_flag) {
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n4;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(g) + i);
    const float4 b = __ldg(reinterpret_cast<const float4*>(u) + i);
    float4 v;
    v.x = b.x > 0.0f ? a.x : __fmul_rn(a.x, slope); v.y = b.y > 0.0f ? a.y : __fmul_rn(a.y, slope);
    v.z = b.z > 0.0f ? a.z : __fmul_rn(a.z, slope); v.w = b.w > 0.0f ? a.w : __fmul_rn(a.w, slope);
    if (outside_fp16_range(v)) atomicOr(err_flag, 8);
    uint2 h, l;
    split4(v, &h, &l);
    reinterpret_cast<uint2*>(hi)[i] = h;
    reinterpret_cast<uint2*>(lo)[i] = l;
  }
}

// db[c] = sum over rows of G'[row, c] (G' = g * LeakyReLU'(u)), fp32 result, double accumulation, deterministic:
// block j sums rows j, j + gridDim.x, ... for 128 columns (blockIdx.y picks the column group) into part[j][C];
// bias_grad_finish_kernel adds the partials in block order.
__global__ void bias_grad_partial_kernel(const float* __restrict__ g, const float* __restrict__ u, float slope, size_t rows,
                                         int C, double* __restrict__ part) {
  const int c = blockIdx.y * 128 + threadIdx.x;
  if (c >= C) return;
  double acc = 0.0;
  for (size_t r = blockIdx.x; r < rows; r += gridDim.x) {
    const float a = g[r * C + c];
    acc += static_cast<double>(u[r * C + c] > 0.0f ? a : __fmul_rn(a, slope));
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

// ------------------------------------------------------------------------------------------------
// Duration predictor (layers/duration_predictor.py:57-77), training.  Layer: u = relu(conv_k(x) + b) (the GEMM's
// epilogue), h = LayerNorm_C(u) * gamma + beta [* keep], head: out = h_L . w + b, masked positions 0 (:85-86).
// `keep` is the train-mode dropout mask already scaled by 1 / (1 - p) (drawn by the caller; nullptr = no dropout).
// One warp per row; the row arithmetic is layernorm_row's (path_kernels.cuh), eps 1e-12 (layers/layer_norm.py:22).
__global__ void ln_train_fwd_kernel(const float* __restrict__ u, size_t rows, const float* __restrict__ gamma,
                                    const float* __restrict__ beta, const float* __restrict__ keep,
                                    float* __restrict__ h, __half* __restrict__ hi, __half* __restrict__ lo,
                                    const float* __restrict__ head_w, const float* __restrict__ head_b,
                                    const uint8_t* __restrict__ mask, float* __restrict__ out, int* __restrict__ err_flag) {
  constexpr int C = 512;
  const size_t row = blockIdx.x * static_cast<size_t>(blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float4* ur = reinterpret_cast<const float4*>(u + row * C);
  float4 v[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) v[j] = ur[j * 32 + lane];
  float s = 0.0f;
#pragma unroll
  for (int j = 0; j < 4; ++j) s += (v[j].x + v[j].y) + (v[j].z + v[j].w);
  const float mean = warp_sum(s) * (1.0f / C);
  float ss = 0.0f;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float a = v[j].x - mean, b2 = v[j].y - mean, c = v[j].z - mean, d = v[j].w - mean;
    ss += (a * a + b2 * b2) + (c * c + d * d);
  }
  const float rstd = __fdiv_rn(1.0f, sqrtf(warp_sum(ss) * (1.0f / C) + 1e-12f));
  float dot = 0.0f;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int c = (j * 32 + lane) * 4;
    const float4 g = __ldg(reinterpret_cast<const float4*>(gamma + c));
    const float4 bb = __ldg(reinterpret_cast<const float4*>(beta + c));
    float4 y;
    y.x = (v[j].x - mean) * rstd * g.x + bb.x; y.y = (v[j].y - mean) * rstd * g.y + bb.y;
    y.z = (v[j].z - mean) * rstd * g.z + bb.z; y.w = (v[j].w - mean) * rstd * g.w + bb.w;
    if (keep != nullptr) {
      const float4 k = __ldg(reinterpret_cast<const float4*>(keep + row * C + c));
      y.x *= k.x; y.y *= k.y; y.z *= k.z; y.w *= k.w;
    }
    *reinterpret_cast<float4*>(h + row * C + c) = y;
    if (hi != nullptr) {
      if (outside_fp16_range(y)) atomicOr(err_flag, 8);
      uint2 hh, ll;
      split4(y, &hh, &ll);
      *reinterpret_cast<uint2*>(hi + row * C + c) = hh;
      *reinterpret_cast<uint2*>(lo + row * C + c) = ll;
    }
    if (head_w != nullptr) {
      const float4 w = __ldg(reinterpret_cast<const float4*>(head_w + c));
      dot += (y.x * w.x + y.y * w.y) + (y.z * w.z + y.w * w.w);
    }
  }
  if (head_w != nullptr) {
    dot = warp_sum(dot);
    if (lane == 0) out[row] = (mask != nullptr && mask[row]) ? 0.0f : dot + head_b[0];
  }
}

// Backward of the same row: gh = dL/dh (from the next conv's data gradient), or, for the last layer, dy * head_w with
// dy = dL/dout (0 at masked positions).  With xh = (u - mean) * rstd, gk = gh * keep, gg = gk * gamma:
//   dL/du = rstd * (gg - mean_c(gg) - xh * mean_c(gg * xh))     (before ReLU'; the conv backward applies the mask of u)
//   dgamma[c] += gk * xh, dbeta[c] += gk, and for the last layer dw_head[c] += dy * h, db_head += dy.
// Column sums: every warp walks rows warp, warp + W, ... and keeps its 16 columns per lane in double; the per-warp
// partials part[4][W][C] (gamma, beta, head_w, head_b in column 0) are added in warp order by ln_train_finish_kernel.
__global__ void ln_train_bwd_kernel(const float* __restrict__ gh, const float* __restrict__ dy,
                                    const uint8_t* __restrict__ mask, const float* __restrict__ head_w,
                                    const float* __restrict__ u, const float* __restrict__ h,
                                    const float* __restrict__ keep, const float* __restrict__ gamma, size_t rows,
                                    float* __restrict__ gu, double* __restrict__ part) {
  constexpr int C = 512;
  const int lane = threadIdx.x & 31;
  const size_t W = static_cast<size_t>(gridDim.x) * (blockDim.x >> 5);
  const size_t wid = blockIdx.x * static_cast<size_t>(blockDim.x >> 5) + (threadIdx.x >> 5);
  double ag[16], ab[16], aw[16], adb = 0.0;
#pragma unroll
  for (int i = 0; i < 16; ++i) { ag[i] = 0.0; ab[i] = 0.0; aw[i] = 0.0; }
  for (size_t row = wid; row < rows; row += W) {
    float4 v[4], g[4];
    const float d = dy != nullptr ? ((mask != nullptr && mask[row]) ? 0.0f : dy[row]) : 0.0f;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int c = (j * 32 + lane) * 4;
      v[j] = *reinterpret_cast<const float4*>(u + row * C + c);
      if (dy != nullptr) {
        const float4 w = __ldg(reinterpret_cast<const float4*>(head_w + c));
        g[j] = make_float4(d * w.x, d * w.y, d * w.z, d * w.w);
        const float4 hh = *reinterpret_cast<const float4*>(h + row * C + c);
        aw[4 * j + 0] += static_cast<double>(d * hh.x); aw[4 * j + 1] += static_cast<double>(d * hh.y);
        aw[4 * j + 2] += static_cast<double>(d * hh.z); aw[4 * j + 3] += static_cast<double>(d * hh.w);
      } else {
        g[j] = *reinterpret_cast<const float4*>(gh + row * C + c);
      }
      if (keep != nullptr) {
        const float4 k = __ldg(reinterpret_cast<const float4*>(keep + row * C + c));
        g[j].x *= k.x; g[j].y *= k.y; g[j].z *= k.z; g[j].w *= k.w;
      }
    }
    if (dy != nullptr && lane == 0) adb += static_cast<double>(d);
    float s = 0.0f;
#pragma unroll
    for (int j = 0; j < 4; ++j) s += (v[j].x + v[j].y) + (v[j].z + v[j].w);
    const float mean = warp_sum(s) * (1.0f / C);
    float ss = 0.0f;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float a = v[j].x - mean, b2 = v[j].y - mean, c = v[j].z - mean, e = v[j].w - mean;
      ss += (a * a + b2 * b2) + (c * c + e * e);
    }
    const float rstd = __fdiv_rn(1.0f, sqrtf(warp_sum(ss) * (1.0f / C) + 1e-12f));
    float m1 = 0.0f, m2 = 0.0f;
    float xh[16], gg[16];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int c = (j * 32 + lane) * 4;
      const float4 gm = __ldg(reinterpret_cast<const float4*>(gamma + c));
      const float vv[4] = {v[j].x, v[j].y, v[j].z, v[j].w};
      const float gk[4] = {g[j].x, g[j].y, g[j].z, g[j].w};
      const float gam[4] = {gm.x, gm.y, gm.z, gm.w};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        xh[4 * j + i] = (vv[i] - mean) * rstd;
        gg[4 * j + i] = gk[i] * gam[i];
        m1 += gg[4 * j + i];
        m2 += gg[4 * j + i] * xh[4 * j + i];
        ag[4 * j + i] += static_cast<double>(gk[i] * xh[4 * j + i]);
        ab[4 * j + i] += static_cast<double>(gk[i]);
      }
    }
    m1 = warp_sum(m1) * (1.0f / C);
    m2 = warp_sum(m2) * (1.0f / C);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int c = (j * 32 + lane) * 4;
      float4 o;
      o.x = rstd * (gg[4 * j + 0] - m1 - xh[4 * j + 0] * m2); o.y = rstd * (gg[4 * j + 1] - m1 - xh[4 * j + 1] * m2);
      o.z = rstd * (gg[4 * j + 2] - m1 - xh[4 * j + 2] * m2); o.w = rstd * (gg[4 * j + 3] - m1 - xh[4 * j + 3] * m2);
      *reinterpret_cast<float4*>(gu + row * C + c) = o;
    }
  }
#pragma unroll
  for (int j = 0; j < 4; ++j)
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const size_t c = static_cast<size_t>((j * 32 + lane) * 4 + i);
      part[(0 * W + wid) * C + c] = ag[4 * j + i];
      part[(1 * W + wid) * C + c] = ab[4 * j + i];
      part[(2 * W + wid) * C + c] = aw[4 * j + i];
    }
  if (lane == 0) part[(3 * W + wid) * C] = adb;
}
// out[q][c] = sum over the W warp partials, in warp order; q = 0 gamma, 1 beta, 2 head_w, 3 head_b (column 0 only)
__global__ void ln_train_finish_kernel(const double* __restrict__ part, int W, float* __restrict__ dgamma,
                                       float* __restrict__ dbeta, float* __restrict__ dhead_w, float* __restrict__ dhead_b) {
  constexpr int C = 512;
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  const int q = blockIdx.y;
  if (c >= C) return;
  float* dst = q == 0 ? dgamma : q == 1 ? dbeta : q == 2 ? dhead_w : dhead_b;
  if (dst == nullptr || (q == 3 && c != 0)) return;
  double acc = 0.0;
  for (int j = 0; j < W; ++j) acc += part[(static_cast<size_t>(q) * W + j) * C + c];
  dst[c] = static_cast<float>(acc);
}

// ------------------------------------------------------------------------------------------------
// Criterion with gradients (losses/fastspeech_loss.py:54-67, after_outs = None, use_weighted_masking = False):
//   mel term  = mean over the selected elements of (before_outs - ys)^2  (use_mse) or |before_outs - ys|
//   dur term  = mean over the selected tokens of |d_outs - ds|
// "selected" = positions inside the lengths when use_masking, everything otherwise.  Block partial sums in double,
// added in block order by fastspeech_loss_finish_kernel (deterministic); the element gradients d(term)/d(input) are
// written alongside (0 outside the selection; sign(0) = 0 like torch's L1 gradient).
constexpr int kLossBlocks = 592;
__global__ void fastspeech_loss_kernel(const float* __restrict__ mel, const float* __restrict__ ys,
                                       const float* __restrict__ dur, const float* __restrict__ ds,
                                       const long long* __restrict__ ilens, const long long* __restrict__ olens, int B,
                                       int T1, int T2, int odim, int use_masking, int use_mse,
                                       float* __restrict__ grad_mel, float* __restrict__ grad_dur,
                                       double* __restrict__ part, int* __restrict__ flags) {
  // selected counts (every block computes them: B is small)
  __shared__ double s_cnt[2];
  __shared__ double s_red[2][8];
  if (threadIdx.x == 0) {
    double nm = 0.0, nd = 0.0;
    for (int b = 0; b < B; ++b) {
      long long ol = olens[b], il = ilens[b];
      if (ol < 0 || ol > T2 || il < 0 || il > T1) { atomicOr(flags, 16); ol = max(0ll, min(ol, (long long)T2)); il = max(0ll, min(il, (long long)T1)); }
      nm += use_masking ? static_cast<double>(ol) * odim : static_cast<double>(T2) * odim;
      nd += use_masking ? static_cast<double>(il) : static_cast<double>(T1);
    }
    s_cnt[0] = nm; s_cnt[1] = nd;
  }
  __syncthreads();
  const float inv_m = s_cnt[0] > 0.0 ? static_cast<float>(1.0 / s_cnt[0]) : 0.0f;
  const float inv_d = s_cnt[1] > 0.0 ? static_cast<float>(1.0 / s_cnt[1]) : 0.0f;
  double am = 0.0, ad = 0.0;
  const size_t nmel = static_cast<size_t>(B) * T2 * odim;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < nmel;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const size_t r = i / odim;
    const int b = static_cast<int>(r / T2), t = static_cast<int>(r % T2);
    float gr = 0.0f;
    if (!use_masking || t < min(max(olens[b], 0ll), (long long)T2)) {
      const float d = __fsub_rn(mel[i], ys[i]);
      if (use_mse) { am += static_cast<double>(__fmul_rn(d, d)); gr = __fmul_rn(__fmul_rn(2.0f, d), inv_m); }
      else { am += static_cast<double>(fabsf(d)); gr = d > 0.0f ? inv_m : (d < 0.0f ? -inv_m : 0.0f); }
    }
    if (grad_mel != nullptr) grad_mel[i] = gr;
  }
  const size_t ntok = static_cast<size_t>(B) * T1;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < ntok;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int b = static_cast<int>(i / T1), t = static_cast<int>(i % T1);
    float gr = 0.0f;
    if (!use_masking || t < min(max(ilens[b], 0ll), (long long)T1)) {
      const float d = __fsub_rn(dur[i], ds[i]);
      ad += static_cast<double>(fabsf(d));
      gr = d > 0.0f ? inv_d : (d < 0.0f ? -inv_d : 0.0f);
    }
    if (grad_dur != nullptr) grad_dur[i] = gr;
  }
  am = warp_sum(am);
  ad = warp_sum(ad);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (lane == 0) { s_red[0][w] = am; s_red[1][w] = ad; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0.0, d = 0.0;
    for (int j = 0; j < static_cast<int>(blockDim.x >> 5); ++j) { a += s_red[0][j]; d += s_red[1][j]; }
    part[2 * blockIdx.x] = a;
    part[2 * blockIdx.x + 1] = d;
    if (blockIdx.x == 0) { part[2 * gridDim.x] = s_cnt[0]; part[2 * gridDim.x + 1] = s_cnt[1]; }
  }
}
__global__ void fastspeech_loss_finish_kernel(const double* __restrict__ part, int nblocks, float* __restrict__ losses) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  double a = 0.0, d = 0.0;
  for (int j = 0; j < nblocks; ++j) { a += part[2 * j]; d += part[2 * j + 1]; }
  const double nm = part[2 * nblocks], nd = part[2 * nblocks + 1];
  losses[0] = static_cast<float>(nm > 0.0 ? a / nm : 0.0 / 0.0);      // mean over nothing is NaN, like torch
  losses[1] = static_cast<float>(nd > 0.0 ? d / nd : 0.0 / 0.0);
}
// out = in * scalar[0] (the chain rule of a loss term: its upstream gradient is a device scalar)
__global__ void scale_by_scalar_kernel(const float* __restrict__ in, const float* __restrict__ scalar, size_t n,
                                       float* __restrict__ out) {
  const float s = scalar[0];
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<size_t>(gridDim.x) * blockDim.x)
    out[i] = __fmul_rn(in[i], s);
}

}  // namespace efts
