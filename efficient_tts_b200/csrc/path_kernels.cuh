// HBM-bound kernels of the EFTS-CNN forward path: embedding gather, operand-plane split, the IMV
// monotonic-alignment chain (softmax-expectation, ReLU-diff prefix scan, aligned positions, Gaussian
// reconstruction), channel LayerNorm + duration head, masked losses and the length regulator.
// All of them are coalesced / vectorised streaming kernels built on warp shuffles; citations are
// relative to /root/reference/nntts.
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>

#include "gemm_params.cuh"

namespace efts {

constexpr float kSplitScale = 2048.0f;   // 2^11, see gemm_params.cuh

// Lengths are loop bounds over rows of T entries (and over shared-memory tiles sized for T): a length outside
// [0, T] is clamped wherever one is loaded (the forward entry points additionally report it, prep_lengths_kernel).
__device__ __forceinline__ int clamp_len(int len, int T) { return min(max(len, 0), T); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// fp32 -> (hi, lo) fp16 operand planes: hi = fp16(x), lo = fp16((x - hi) * 2^11).
__device__ __forceinline__ void split4(const float4 v, uint2* hi, uint2* lo) {
  const float x[4] = {v.x, v.y, v.z, v.w};
  __align__(8) __half h[4];
  __align__(8) __half l[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    h[j] = __float2half_rn(x[j]);
    l[j] = __float2half_rn((x[j] - __half2float(h[j])) * kSplitScale);
  }
  *hi = *reinterpret_cast<const uint2*>(h);
  *lo = *reinterpret_cast<const uint2*>(l);
}

// ------------------------------------------------------------------------------------------------
// lengths int64 -> int32 (+ the checks the reference performs on the host, utils/nets_utils.py:146-156)
// flags[0] |= 1 if max(text_lengths) != T1, |= 2 if max(speech_lengths) != T2, |= 16 if any length lies outside
// [0, padded dim] (the stored int32 lengths are clamped, so no kernel ever walks past a row; the reference fails
// with a shape error on such input and the binding raises when it reads the flags)
__global__ void prep_lengths_kernel(const int64_t* __restrict__ tl, const int64_t* __restrict__ sl, int B,
                                    int T1, int T2, int* __restrict__ tl32, int* __restrict__ sl32,
                                    int* __restrict__ flags) {
  int mt = 0, ms = 0, bad = 0;
  for (int b = threadIdx.x; b < B; b += blockDim.x) {
    const int64_t a = tl[b];
    bad |= (a < 0 || a > T1);
    const int ac = static_cast<int>(min(max(a, static_cast<int64_t>(0)), static_cast<int64_t>(T1)));
    tl32[b] = ac;
    mt = max(mt, ac);
    if (sl != nullptr) {
      const int64_t c = sl[b];
      bad |= (c < 0 || c > T2);
      const int cc = static_cast<int>(min(max(c, static_cast<int64_t>(0)), static_cast<int64_t>(T2)));
      sl32[b] = cc;
      ms = max(ms, cc);
    }
  }
  __shared__ int s_mt, s_ms;
  if (threadIdx.x == 0) { s_mt = 0; s_ms = 0; }
  __syncthreads();
  atomicMax(&s_mt, mt);
  atomicMax(&s_ms, ms);
  if (bad) atomicOr(flags, 16);
  __syncthreads();
  if (threadIdx.x == 0) {
    int f = 0;
    if (s_mt != T1) f |= 1;
    if (sl != nullptr && s_ms != T2) f |= 2;
    if (f) atomicOr(flags, f);
  }
}

// ------------------------------------------------------------------------------------------------
// Embedding gather (models/efficient_tts.py:144,246; no padding_idx: id 0 is a real row).
// One block of C/4 threads per (b, i); writes the fp32 master and the operand planes.
// `lens` (ragged batched synthesis only): rows i >= lens[b] are written as zeros, so that an utterance
// sees the same zero padding it would see when run alone and unpadded.
__global__ void embed_kernel(const int64_t* __restrict__ text, const float* __restrict__ table,
                             int num_symbols, int C, float* __restrict__ out, __half* __restrict__ hi,
                             __half* __restrict__ lo, int* __restrict__ flags, const int* __restrict__ lens,
                             int T) {
  const size_t row = blockIdx.x;
  if (lens != nullptr && static_cast<int>(row % T) >= lens[row / T]) {
    const int c0 = threadIdx.x * 4;
    if (c0 < C) {
      *reinterpret_cast<float4*>(out + row * C + c0) = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
      *reinterpret_cast<uint2*>(hi + row * C + c0) = make_uint2(0u, 0u);
      *reinterpret_cast<uint2*>(lo + row * C + c0) = make_uint2(0u, 0u);
    }
    return;
  }
  long long id = text[row];
  if (id < 0 || id >= num_symbols) {
    if (threadIdx.x == 0) atomicOr(flags, 4);
    id = 0;
  }
  const int c = threadIdx.x * 4;
  if (c >= C) return;
  const float4 v = __ldg(reinterpret_cast<const float4*>(table + static_cast<size_t>(id) * C + c));
  if (outside_fp16_range(v))
    atomicOr(flags, 8);                           // outside the fp16 operand range
  *reinterpret_cast<float4*>(out + row * C + c) = v;
  uint2 h, l;
  split4(v, &h, &l);
  *reinterpret_cast<uint2*>(hi + row * C + c) = h;
  *reinterpret_cast<uint2*>(lo + row * C + c) = l;
}

// fp32 [n] -> operand planes; n multiple of 4.
__global__ void split_planes_kernel(const float* __restrict__ x, size_t n4, __half* __restrict__ hi,
                                    __half* __restrict__ lo, int* __restrict__ err_flag) {
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n4;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(x) + i);
    if (outside_fp16_range(v) && err_flag != nullptr)
      atomicOr(err_flag, 8);                      // outside the fp16 operand range
    uint2 h, l;
    split4(v, &h, &l);
    reinterpret_cast<uint2*>(hi)[i] = h;
    reinterpret_cast<uint2*>(lo)[i] = l;
  }
}

// fp32 [B, T, C] -> transposed planes [B, C, ldt] (t contiguous), zero for t in [T, ldt).
// 32x32 shared-memory tile transpose; grid (ceil(ldt/32), C/32, B), block (32, 8).
__global__ void split_transpose_kernel(const float* __restrict__ x, int T, int C, int ldt,
                                       __half* __restrict__ hiT, __half* __restrict__ loT) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z;
  const int t0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  for (int r = threadIdx.y; r < 32; r += 8) {
    const int t = t0 + r;
    tile[r][threadIdx.x] = t < T ? x[(static_cast<size_t>(b) * T + t) * C + c0 + threadIdx.x] : 0.0f;
  }
  __syncthreads();
  for (int r = threadIdx.y; r < 32; r += 8) {
    const int t = t0 + threadIdx.x;
    if (t < ldt) {
      const float v = tile[threadIdx.x][r];
      const __half h = __float2half_rn(v);
      const size_t o = (static_cast<size_t>(b) * C + c0 + r) * ldt + t;
      hiT[o] = h;
      loT[o] = __float2half_rn((v - __half2float(h)) * kSplitScale);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// B2: imv_generator tail (models/efficient_tts.py:314-323): Delta = relu(diff), Delta[0] = 0;
// prefix sum over frames accumulated in double and rounded to fp32 per prefix (what torch's CPU
// cumsum does for fp32); * mel_mask; / clamp(max, 1e-8); * (T1_b - 1).  One warp per utterance.
// The position expectation imv'[t] comes either from `imv_raw` or, when `part` is given, from the
// per-column-tile softmax partials the energy GEMM epilogue wrote (max, sum exp, sum exp * i).
__global__ void imv_scan_kernel(const float* __restrict__ imv_raw, const float4* __restrict__ part, int n_part,
                                const int* __restrict__ tl, const int* __restrict__ sl, int B, int T2,
                                float* __restrict__ imv) {
  const int lane = threadIdx.x & 31;
  const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (b >= B) return;
  float* y = imv + static_cast<size_t>(b) * T2;
  const int L2 = clamp_len(sl[b], T2);
  double carry = 0.0;
  float vmax = -CUDART_INF_F;
  float prev_last = 0.0f;
  for (int base = 0; base < T2; base += 32) {
    const int t = base + lane;
    float r = 0.0f;
    if (t < T2) {
      if (part == nullptr) {
        r = imv_raw[static_cast<size_t>(b) * T2 + t];
      } else if (t < L2) {                       // pad frames: alpha is zeroed (:168) -> expectation 0
        const float4* pp = part + (static_cast<size_t>(b) * T2 + t) * n_part;
        float mx = -CUDART_INF_F;
        for (int k = 0; k < n_part; ++k) mx = fmaxf(mx, pp[k].x);
        double den = 0.0, num = 0.0;
        for (int k = 0; k < n_part; ++k) {
          const float4 v = pp[k];
          const double sc = static_cast<double>(expf(v.x - mx));
          den = fma(static_cast<double>(v.y), sc, den);
          num = fma(static_cast<double>(v.z), sc, num);
        }
        r = static_cast<float>(num / den);
      }
    }
    float prev = __shfl_up_sync(0xffffffffu, r, 1);
    if (lane == 0) prev = prev_last;
    prev_last = __shfl_sync(0xffffffffu, r, 31);
    float d = 0.0f;
    if (t > 0 && t < T2) d = fmaxf(__fsub_rn(r, prev), 0.0f);
    double s = static_cast<double>(d);
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const double n = __shfl_up_sync(0xffffffffu, s, o);
      if (lane >= o) s += n;
    }
    s += carry;
    carry = __shfl_sync(0xffffffffu, s, 31);
    if (t < T2) {
      const float c = __fmul_rn(static_cast<float>(s), t < L2 ? 1.0f : 0.0f);
      y[t] = c;
      vmax = fmaxf(vmax, c);
    }
  }
  vmax = warp_max(vmax);
  const float last = fmaxf(vmax, 1e-8f);
  const float scale = static_cast<float>(tl[b]) - 1.0f;
  __syncwarp();
  for (int t = lane; t < T2; t += 32) y[t] = __fmul_rn(__fdiv_rn(y[t], last), scale);
}

// B3: get_aligned_positions (models/efficient_tts.py:338-345): for each valid token i,
// e[b,i] = sum_t softmax_t(-(imv[t]-i)^2 * sigma_e over valid frames) * t.  One warp per (b, i).
__global__ void aligned_positions_kernel(const float* __restrict__ imv, const int* __restrict__ tl,
                                         const int* __restrict__ sl, int T1, int T2, float sigma_e,
                                         float* __restrict__ e, const float* __restrict__ pvec = nullptr) {
  const int lane = threadIdx.x & 31;
  const int b = blockIdx.y;
  const int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (i >= T1) return;
  if (i >= tl[b]) {
    if (lane == 0) e[static_cast<size_t>(b) * T1 + i] = 0.0f;
    return;
  }
  const float* x = imv + static_cast<size_t>(b) * T2;
  const int L2 = clamp_len(sl[b], T2);
  const float p = pvec != nullptr ? pvec[static_cast<size_t>(b) * T1 + i] : static_cast<float>(i);
  float m = -CUDART_INF_F;
  for (int t = lane; t < L2; t += 32) {
    const float d = __fsub_rn(x[t], p);
    m = fmaxf(m, __fmul_rn(__fmul_rn(-1.0f, __fmul_rn(d, d)), sigma_e));
  }
  m = warp_max(m);
  float den = 0.0f, num = 0.0f;
  for (int t = lane; t < L2; t += 32) {
    const float d = __fsub_rn(x[t], p);
    const float g = __fmul_rn(__fmul_rn(-1.0f, __fmul_rn(d, d)), sigma_e);
    const float ev = expf(g - m);
    den += ev;
    num = fmaf(ev, static_cast<float>(t), num);
  }
  den = warp_sum(den);
  num = warp_sum(num);
  if (lane == 0) e[static_cast<size_t>(b) * T1 + i] = __fdiv_rn(num, den);
}

// B2 / B3, one block per utterance (the launches forward() uses).  Same arithmetic, element for element and in
// the same order, as imv_scan_kernel / aligned_positions_kernel above -- those run one warp per utterance /
// per token straight from global memory and spend their time waiting on dependent loads (59 + 49 us at C3 for
// a few MB).  Here the whole block merges the softmax partials into shared memory, one warp runs the prefix
// scan out of shared memory, and the block normalises and stores; the aligned-position kernel stages imv[b,:]
// in shared memory once per 32 tokens and takes max_t G from min_t |imv[t] - p| (G is monotone in |d|).
constexpr int IMV_BLOCK_THREADS = 256;
__global__ void __launch_bounds__(IMV_BLOCK_THREADS)
imv_scan_block_kernel(const float* __restrict__ imv_raw, const float4* __restrict__ part, int n_part,
                      const int* __restrict__ tl, const int* __restrict__ sl, int T2, float* __restrict__ imv) {
  extern __shared__ float scan_smem[];                       // [T2]
  __shared__ float s_last;
  const int b = blockIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int L2 = clamp_len(sl[b], T2);
  float* y = imv + static_cast<size_t>(b) * T2;
  for (int t = threadIdx.x; t < T2; t += IMV_BLOCK_THREADS) {
    float r = 0.0f;
    if (part == nullptr) {
      r = imv_raw[static_cast<size_t>(b) * T2 + t];
    } else if (t < L2) {                                     // pad frames: alpha is zeroed (:168) -> expectation 0
      const float4* pp = part + (static_cast<size_t>(b) * T2 + t) * n_part;
      float mx = -CUDART_INF_F;
      for (int k = 0; k < n_part; ++k) mx = fmaxf(mx, pp[k].x);
      double den = 0.0, num = 0.0;
      for (int k = 0; k < n_part; ++k) {
        const float4 v = pp[k];
        const double sc = static_cast<double>(expf(v.x - mx));
        den = fma(static_cast<double>(v.y), sc, den);
        num = fma(static_cast<double>(v.z), sc, num);
      }
      r = static_cast<float>(num / den);
    }
    scan_smem[t] = r;
  }
  __syncthreads();
  if (warp == 0) {
    double carry = 0.0;
    float vmax = -CUDART_INF_F;
    float prev_last = 0.0f;
    for (int base = 0; base < T2; base += 32) {
      const int t = base + lane;
      const float r = t < T2 ? scan_smem[t] : 0.0f;
      float prev = __shfl_up_sync(0xffffffffu, r, 1);
      if (lane == 0) prev = prev_last;
      prev_last = __shfl_sync(0xffffffffu, r, 31);
      float d = 0.0f;
      if (t > 0 && t < T2) d = fmaxf(__fsub_rn(r, prev), 0.0f);
      double s = static_cast<double>(d);
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const double n = __shfl_up_sync(0xffffffffu, s, o);
        if (lane >= o) s += n;
      }
      s += carry;
      carry = __shfl_sync(0xffffffffu, s, 31);
      if (t < T2) {
        const float c = __fmul_rn(static_cast<float>(s), t < L2 ? 1.0f : 0.0f);
        scan_smem[t] = c;
        vmax = fmaxf(vmax, c);
      }
    }
    vmax = warp_max(vmax);
    if (lane == 0) s_last = fmaxf(vmax, 1e-8f);
  }
  __syncthreads();
  const float last = s_last;
  const float scale = static_cast<float>(tl[b]) - 1.0f;
  for (int t = threadIdx.x; t < T2; t += IMV_BLOCK_THREADS) y[t] = __fmul_rn(__fdiv_rn(scan_smem[t], last), scale);
}

constexpr int AP_TOKENS = 32;                                // tokens per block: 8 warps x 4
__global__ void __launch_bounds__(IMV_BLOCK_THREADS)
aligned_positions_block_kernel(const float* __restrict__ imv, const int* __restrict__ tl,
                               const int* __restrict__ sl, int T1, int T2, float sigma_e,
                               float* __restrict__ e, const float* __restrict__ pvec) {
  extern __shared__ float ap_smem[];                         // [L2]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int b = blockIdx.y;
  const int i0 = blockIdx.x * AP_TOKENS;
  const int L1 = clamp_len(tl[b], T1), L2 = clamp_len(sl[b], T2);
  if (i0 < L1) {
    const float* x = imv + static_cast<size_t>(b) * T2;
    for (int t = threadIdx.x; t < L2; t += IMV_BLOCK_THREADS) ap_smem[t] = x[t];
  }
  __syncthreads();
  // each warp owns four consecutive tokens and walks the frames once for all of them (one shared-memory load
  // and one int->float conversion per frame serve four softmax rows); per token the sums run over the frames in
  // the same order as in aligned_positions_kernel
  const int ib = i0 + warp * 4;
  if (ib >= T1) return;
  if (ib >= L1) {
    if (lane < 4 && ib + lane < T1) e[static_cast<size_t>(b) * T1 + ib + lane] = 0.0f;
    return;
  }
  float p[4], dmin[4], m[4], den[4], num[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int i = min(ib + k, T1 - 1);
    p[k] = pvec != nullptr ? pvec[static_cast<size_t>(b) * T1 + i] : static_cast<float>(ib + k);
    dmin[k] = CUDART_INF_F;
    den[k] = 0.0f; num[k] = 0.0f;
  }
  for (int t = lane; t < L2; t += 32) {
    const float x = ap_smem[t];
#pragma unroll
    for (int k = 0; k < 4; ++k) dmin[k] = fminf(dmin[k], fabsf(__fsub_rn(x, p[k])));
  }
#pragma unroll
  for (int k = 0; k < 4; ++k) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) dmin[k] = fminf(dmin[k], __shfl_xor_sync(0xffffffffu, dmin[k], o));
    // max_t G: for L2 == 0 the reference's softmax over an empty row gives NaN; dmin = inf reproduces it
    m[k] = __fmul_rn(__fmul_rn(-1.0f, __fmul_rn(dmin[k], dmin[k])), sigma_e);
  }
#pragma unroll 2
  for (int t = lane; t < L2; t += 32) {
    const float x = ap_smem[t];
    const float tf = static_cast<float>(t);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float d = __fsub_rn(x, p[k]);
      const float g = __fmul_rn(__fmul_rn(-1.0f, __fmul_rn(d, d)), sigma_e);
      const float ev = expf(g - m[k]);
      den[k] += ev;
      num[k] = fmaf(ev, tf, num[k]);
    }
  }
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    den[k] = warp_sum(den[k]);
    num[k] = warp_sum(num[k]);
    const int i = ib + k;
    if (lane == 0 && i < T1) e[static_cast<size_t>(b) * T1 + i] = i < L1 ? __fdiv_rn(num[k], den[k]) : 0.0f;
  }
}

// B4: reconstruct_align_from_aligned_position (models/efficient_tts.py:366-375, :186):
// R[b,i,t] = softmax_i( fl32(-sigma) * (q_t - e_i)^2 over valid tokens ), zero at pad tokens/frames;
// q_t = t on valid frames.  One thread per frame t, e[b,:] staged in shared memory.  Writes the
// returned fp32 matrix [B,T1,T2] (coalesced along t) and the K-major operand planes [B,T2,ldp]
// (zero for i >= T1_b) the expansion GEMM reads.  tl / sl == nullptr: no masks (inference path).
__global__ void reconstruct_alignment_kernel(const float* __restrict__ e, const int* __restrict__ tl,
                                             const int* __restrict__ sl, int T1, int T2, int ldp,
                                             float neg_sigma, float* __restrict__ R,
                                             __half* __restrict__ p_hi, __half* __restrict__ p_lo) {
  extern __shared__ float se[];
  const int b = blockIdx.y;
  const int L1 = tl != nullptr ? clamp_len(tl[b], T1) : T1;
  const int L2 = sl != nullptr ? clamp_len(sl[b], T2) : T2;
  for (int i = threadIdx.x; i < T1; i += blockDim.x) se[i] = e[static_cast<size_t>(b) * T1 + i];
  __syncthreads();
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= T2) return;
  const bool live = t < L2;
  const float q = live ? static_cast<float>(t) : 0.0f;
  float m = -CUDART_INF_F;
  for (int i = 0; i < L1; ++i) {
    const float d = __fsub_rn(q, se[i]);
    m = fmaxf(m, __fmul_rn(neg_sigma, __fmul_rn(d, d)));
  }
  float den = 0.0f;
  for (int i = 0; i < L1; ++i) {
    const float d = __fsub_rn(q, se[i]);
    den += expf(__fmul_rn(neg_sigma, __fmul_rn(d, d)) - m);
  }
  float* Rb = R + static_cast<size_t>(b) * T1 * T2 + t;
  __half* ph = p_hi + (static_cast<size_t>(b) * T2 + t) * ldp;
  __half* pl = p_lo + (static_cast<size_t>(b) * T2 + t) * ldp;
  for (int i0 = 0; i0 < ldp; i0 += 8) {
    __align__(16) __half h[8];
    __align__(16) __half l[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int i = i0 + j;
      float r = 0.0f;
      if (live && i < L1) {
        const float d = __fsub_rn(q, se[i]);
        r = __fdiv_rn(expf(__fmul_rn(neg_sigma, __fmul_rn(d, d)) - m), den);
      }
      if (i < T1) Rb[static_cast<size_t>(i) * T2] = r;
      h[j] = __float2half_rn(r);
      l[j] = __float2half_rn((r - __half2float(h[j])) * kSplitScale);
    }
    if (p_hi != nullptr) {
      *reinterpret_cast<uint4*>(ph + i0) = *reinterpret_cast<const uint4*>(h);
      *reinterpret_cast<uint4*>(pl + i0) = *reinterpret_cast<const uint4*>(l);
    }
  }
}

// B4, frame-per-lane: a warp-per-frame organisation issues ~150 M warp instructions for 404 MB at C3 and is bound
// by instruction issue (ncu: 80 % issue-active, 25 % of DRAM peak), not by HBM.  Here every lane owns two frames
// of a 64-frame tile and walks the tokens eight at a time (the warps of a block interleave over the 8-token
// chunks), so that
//   * the softmax over tokens needs no shuffles: per-frame partial minima of |q - e_i| / partial sums go through
//     4 KB of shared memory once per pass.  max_i h_i is h at min_i |q - e_i| exactly (fl(d^2) and
//     fl(neg_sigma * x) are monotone), so the first pass costs two instructions per element;
//   * the exponentials are computed once: the second pass parks them in shared memory ([frame][token] fp32, 32
//     bytes per lane and chunk), the third pass multiplies by the frame's reciprocal sum;
//   * rows of the returned fp32 matrix [B,T1,T2] are stored straight from registers (lanes = consecutive
//     frames, 128 B per warp store);
//   * the K-major fp16 hi/lo operand rows [B,T2,ldp] overwrite the parked exponentials in place (the 32 bytes
//     of a chunk become 16 B of hi + 16 B of lo) and are copied out with 16-byte vectors, each frame's row
//     being one contiguous run in global memory.
// Pad tokens of the last live chunk carry e = 3e38: d^2 = inf, h = -inf, exp = 0 -- no per-element predicate.
// Same arithmetic per element as the kernel above (expf(h - m) * rcp(sum)); only the order of the partial sums
// of the denominator differs.
constexpr int R3_FRAMES = 64;
constexpr int R3_WARPS = 8;
// shared-memory row of one frame in bytes: 4 * ldp, padded so that (bytes / 16) is odd (conflict-free 16-byte
// accesses with lanes = frames)
__host__ __device__ __forceinline__ int r3_row_bytes(int ldp) { return (((ldp >> 2) | 1) << 4); }
__host__ __device__ __forceinline__ size_t r3_smem_bytes(int ldp) {
  return static_cast<size_t>(ldp) * 4 + 2 * R3_WARPS * R3_FRAMES * 4 + static_cast<size_t>(R3_FRAMES) * r3_row_bytes(ldp);
}

// eight fp32 values -> 16 B of fp16 hi + 16 B of fp16 lo (lo = fp16((x - hi) * 2^11))
__device__ __forceinline__ void r3_split8(const float (&r)[8], uint4& h, uint4& l) {
  uint32_t hh[4], ll[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const __half2 h2 = __floats2half2_rn(r[2 * k], r[2 * k + 1]);
    const float2 f = __half22float2(h2);
    const __half2 l2 = __floats2half2_rn((r[2 * k] - f.x) * kSplitScale, (r[2 * k + 1] - f.y) * kSplitScale);
    hh[k] = *reinterpret_cast<const uint32_t*>(&h2);
    ll[k] = *reinterpret_cast<const uint32_t*>(&l2);
  }
  h = make_uint4(hh[0], hh[1], hh[2], hh[3]);
  l = make_uint4(ll[0], ll[1], ll[2], ll[3]);
}

__global__ void __launch_bounds__(32 * R3_WARPS)
reconstruct_alignment_rows_kernel(const float* __restrict__ e, const int* __restrict__ tl,
                                  const int* __restrict__ sl, int T1, int T2, int ldp, float neg_sigma,
                                  float* __restrict__ R, __half* __restrict__ p_hi, __half* __restrict__ p_lo) {
  extern __shared__ __align__(16) uint8_t r3_smem[];
  float* se = reinterpret_cast<float*>(r3_smem);                      // [ldp]
  float* red_m = se + ldp;                                            // [warps][64] partial min |q - e|
  float* red_d = red_m + R3_WARPS * R3_FRAMES;                        // [warps][64] partial denominators
  uint8_t* tile = reinterpret_cast<uint8_t*>(red_d + R3_WARPS * R3_FRAMES);   // [64][row_bytes]
  const int rowb = r3_row_bytes(ldp);
  const int b = blockIdx.y;
  const int t0 = blockIdx.x * R3_FRAMES;
  const int L1 = tl != nullptr ? clamp_len(tl[b], T1) : T1;
  const int L2 = sl != nullptr ? clamp_len(sl[b], T2) : T2;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float* Rb = R + static_cast<size_t>(b) * T1 * T2;
  const int ta = t0 + lane, tb = ta + 32;                              // this lane's two frames
  const bool in_a = ta < T2, in_b = tb < T2;
  if (t0 >= L2) {                                                      // whole tile is padding
    float* pa = Rb + static_cast<size_t>(warp) * T2 + ta;
    for (int i = warp; i < T1; i += R3_WARPS) {
      if (in_a) pa[0] = 0.0f;
      if (in_b) pa[32] = 0.0f;
      pa += static_cast<size_t>(R3_WARPS) * T2;
    }
    return;
  }
  for (int i = threadIdx.x; i < ldp; i += 32 * R3_WARPS) se[i] = i < L1 ? e[static_cast<size_t>(b) * T1 + i] : 3.0e38f;
  __syncthreads();
  const int nch = (L1 + 7) >> 3;                                       // 8-token chunks that hold valid tokens
  const bool live_a = ta < L2, live_b = tb < L2;
  const float qa = live_a ? static_cast<float>(ta) : 0.0f;            // q = t * mel_mask (:368)
  const float qb = live_b ? static_cast<float>(tb) : 0.0f;
  // pass 1: min_i |q - e_i|
  float da = CUDART_INF_F, db = CUDART_INF_F;
  for (int c = warp; c < nch; c += R3_WARPS) {
    const float4 e0 = *reinterpret_cast<const float4*>(se + 8 * c);
    const float4 e1 = *reinterpret_cast<const float4*>(se + 8 * c + 4);
    const float ev[8] = {e0.x, e0.y, e0.z, e0.w, e1.x, e1.y, e1.z, e1.w};
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      da = fminf(da, fabsf(__fsub_rn(qa, ev[j])));
      db = fminf(db, fabsf(__fsub_rn(qb, ev[j])));
    }
  }
  red_m[warp * R3_FRAMES + lane] = da;
  red_m[warp * R3_FRAMES + lane + 32] = db;
  __syncthreads();
#pragma unroll
  for (int w = 0; w < R3_WARPS; ++w) {
    da = fminf(da, red_m[w * R3_FRAMES + lane]);
    db = fminf(db, red_m[w * R3_FRAMES + lane + 32]);
  }
  const float ma = __fmul_rn(neg_sigma, __fmul_rn(da, da));            // = max_i h_i
  const float mb = __fmul_rn(neg_sigma, __fmul_rn(db, db));
  // pass 2: exponentials (parked in shared memory) and denominators
  uint8_t* rowa = tile + lane * rowb;
  uint8_t* rowbp = rowa + 32 * rowb;
  float sa = 0.0f, sb = 0.0f;
  for (int c = warp; c < nch; c += R3_WARPS) {
    const float4 e0 = *reinterpret_cast<const float4*>(se + 8 * c);
    const float4 e1 = *reinterpret_cast<const float4*>(se + 8 * c + 4);
    const float ev[8] = {e0.x, e0.y, e0.z, e0.w, e1.x, e1.y, e1.z, e1.w};
    float xa[8], xb[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float ua = __fsub_rn(qa, ev[j]), ub = __fsub_rn(qb, ev[j]);
      xa[j] = expf(__fmul_rn(neg_sigma, __fmul_rn(ua, ua)) - ma);
      xb[j] = expf(__fmul_rn(neg_sigma, __fmul_rn(ub, ub)) - mb);
      sa += xa[j];
      sb += xb[j];
    }
    *reinterpret_cast<float4*>(rowa + 32 * c) = make_float4(xa[0], xa[1], xa[2], xa[3]);
    *reinterpret_cast<float4*>(rowa + 32 * c + 16) = make_float4(xa[4], xa[5], xa[6], xa[7]);
    *reinterpret_cast<float4*>(rowbp + 32 * c) = make_float4(xb[0], xb[1], xb[2], xb[3]);
    *reinterpret_cast<float4*>(rowbp + 32 * c + 16) = make_float4(xb[4], xb[5], xb[6], xb[7]);
  }
  red_d[warp * R3_FRAMES + lane] = sa;
  red_d[warp * R3_FRAMES + lane + 32] = sb;
  __syncthreads();
  sa = 0.0f; sb = 0.0f;
#pragma unroll
  for (int w = 0; w < R3_WARPS; ++w) {                                 // same order in every warp: one sum per frame
    sa += red_d[w * R3_FRAMES + lane];
    sb += red_d[w * R3_FRAMES + lane + 32];
  }
  // one correctly rounded reciprocal per frame instead of a division per element
  const float inv_a = live_a ? __frcp_rn(sa) : 0.0f;
  const float inv_b = live_b ? __frcp_rn(sb) : 0.0f;
  // pass 3 (each thread revisits exactly what it parked): values -> fp32 rows from registers, hi/lo groups in place
  const bool planes = p_hi != nullptr;
  for (int c = warp; c < nch; c += R3_WARPS) {
    const float4 a0 = *reinterpret_cast<const float4*>(rowa + 32 * c);
    const float4 a1 = *reinterpret_cast<const float4*>(rowa + 32 * c + 16);
    const float4 b0 = *reinterpret_cast<const float4*>(rowbp + 32 * c);
    const float4 b1 = *reinterpret_cast<const float4*>(rowbp + 32 * c + 16);
    float ra[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
    float rb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      ra[j] = __fmul_rn(ra[j], inv_a);
      rb[j] = __fmul_rn(rb[j], inv_b);
    }
    float* pa = Rb + static_cast<size_t>(8 * c) * T2 + ta;              // frame b sits 32 floats further
    if (8 * c + 8 <= T1) {                                             // warp-uniform
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        if (in_a) pa[0] = ra[j];
        if (in_b) pa[32] = rb[j];
        pa += T2;
      }
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        if (8 * c + j < T1) {
          if (in_a) pa[0] = ra[j];
          if (in_b) pa[32] = rb[j];
        }
        pa += T2;
      }
    }
    if (planes) {
      uint4 h, l;
      r3_split8(ra, h, l);
      *reinterpret_cast<uint4*>(rowa + 32 * c) = h;
      *reinterpret_cast<uint4*>(rowa + 32 * c + 16) = l;
      r3_split8(rb, h, l);
      *reinterpret_cast<uint4*>(rowbp + 32 * c) = h;
      *reinterpret_cast<uint4*>(rowbp + 32 * c + 16) = l;
    }
  }
  {                                                                    // pad tokens past the last live chunk
    float* pa = Rb + static_cast<size_t>(8 * nch + warp) * T2 + ta;
    for (int i = 8 * nch + warp; i < T1; i += R3_WARPS) {
      if (in_a) pa[0] = 0.0f;
      if (in_b) pa[32] = 0.0f;
      pa += static_cast<size_t>(R3_WARPS) * T2;
    }
  }
  if (!planes) return;
  __syncthreads();
  // operand rows of the live frames: ldp / 8 sixteen-byte groups per frame and plane, zeros past the live chunks
  const int nlive = min(min(R3_FRAMES, T2 - t0), L2 - t0);
  const int upr = ldp >> 3;
  for (int r = warp; r < nlive; r += R3_WARPS) {
    const size_t o = (static_cast<size_t>(b) * T2 + t0 + r) * ldp;
    const uint8_t* src = tile + r * rowb;
    for (int u = lane; u < upr; u += 32) {
      uint4 vh = make_uint4(0u, 0u, 0u, 0u), vl = vh;
      if (u < nch) {
        vh = *reinterpret_cast<const uint4*>(src + 32 * u);
        vl = *reinterpret_cast<const uint4*>(src + 32 * u + 16);
      }
      *reinterpret_cast<uint4*>(p_hi + o + 8 * u) = vh;
      *reinterpret_cast<uint4*>(p_lo + o + 8 * u) = vl;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Stand-alone forms of the reference's helper methods (models/efficient_tts.py:287-398); forward() itself
// uses the fused kernels above.

// bool prefix mask [B, T] -> int32 lengths (the reference only builds masks with make_non_pad_mask).
__global__ void mask_lengths_kernel(const uint8_t* __restrict__ mask, int B, int T, int* __restrict__ lens) {
  const int lane = threadIdx.x & 31;
  const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (b >= B) return;
  int n = 0;
  for (int t = lane; t < T; t += 32) n += mask[static_cast<size_t>(b) * T + t] != 0;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) n += __shfl_xor_sync(0xffffffffu, n, o);
  if (lane == 0) lens[b] = n;
}

// generate_index_vector (:287-297): p[b, i] = i on valid tokens, 0 on padding.
__global__ void index_vector_kernel(const int* __restrict__ tl, int B, int T1, float* __restrict__ p) {
  const size_t k = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
  if (k >= static_cast<size_t>(B) * T1) return;
  const int b = static_cast<int>(k / T1), i = static_cast<int>(k % T1);
  p[k] = i < tl[b] ? static_cast<float>(i) : 0.0f;
}

// scaled_dot_product_attention tail (:392-398): softmax over valid tokens of the (already scaled) scores
// S[B*T2, ldS], zero on padding, written transposed as alpha[B, T1, T2].  One warp per (b, t).
__global__ void attention_alpha_kernel(const float* __restrict__ S, int ldS, const int* __restrict__ tl, int T1,
                                       int T2, size_t rows, float* __restrict__ alpha) {
  const int lane = threadIdx.x & 31;
  const size_t row = blockIdx.x * static_cast<size_t>(blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int b = static_cast<int>(row / T2), t = static_cast<int>(row % T2);
  const int L = clamp_len(tl[b], T1);
  const float* s = S + row * ldS;
  float m = -CUDART_INF_F;
  for (int i = lane; i < L; i += 32) m = fmaxf(m, s[i]);
  m = warp_max(m);
  float den = 0.0f;
  for (int i = lane; i < L; i += 32) den += expf(s[i] - m);
  den = warp_sum(den);
  float* a = alpha + static_cast<size_t>(b) * T1 * T2 + t;
  for (int i = lane; i < T1; i += 32) a[static_cast<size_t>(i) * T2] = i < L ? __fdiv_rn(expf(s[i] - m), den) : 0.0f;
}

// imv_generator head (:312): imv'[b, t] = sum_i alpha[b, i, t] * p[b, i].  One thread per (b, t).
__global__ void alpha_expectation_kernel(const float* __restrict__ alpha, const float* __restrict__ p, int T1,
                                         int T2, float* __restrict__ out) {
  const int b = blockIdx.y;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= T2) return;
  const float* a = alpha + static_cast<size_t>(b) * T1 * T2 + t;
  float acc = 0.0f;
  for (int i = 0; i < T1; ++i) acc = fmaf(a[static_cast<size_t>(i) * T2], p[static_cast<size_t>(b) * T1 + i], acc);
  out[static_cast<size_t>(b) * T2 + t] = acc;
}

// ------------------------------------------------------------------------------------------------
// Channel LayerNorm of the duration predictor (layers/layer_norm.py:16,30: eps 1e-12, biased
// variance over the C channels) on channels-last rows; one warp per row, C = 512.
// HEAD == 0: writes the normalised row as operand planes (input of the next conv).
// HEAD == 1: fuses the Linear(C -> 1) head (layers/duration_predictor.py:76) and the output mode:
//   0 log domain, 1 clamp(exp(x) - offset, 0), 2 clamp(round(exp(x) - offset), 0) as int64
//   (layers/duration_predictor.py:79-88); rows t >= lens[b] are written as 0.
// The row arithmetic, shared with the resident layer-stack kernel (stack_sm100.cuh) so both produce the same bits.
// v[j] holds columns (j * 32 + lane) * 4 .. + 3 of the row.  HEAD == 0: writes operand planes at hi / lo (row
// base pointers; zeros when !live).  HEAD == 1: returns the Linear(C -> 1) dot product (valid on every lane).
template <int HEAD>
__device__ __forceinline__ float layernorm_row(const float4 (&v)[4], int lane, const float* __restrict__ gamma,
                                               const float* __restrict__ beta, __half* __restrict__ hi,
                                               __half* __restrict__ lo, const float* __restrict__ head_w, bool live) {
  constexpr int C = 512;
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
  const float var = warp_sum(ss) * (1.0f / C);
  const float rstd = __fdiv_rn(1.0f, sqrtf(var + 1e-12f));
  float dot = 0.0f;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int c = (j * 32 + lane) * 4;
    const float4 g = __ldg(reinterpret_cast<const float4*>(gamma + c));
    const float4 bb = __ldg(reinterpret_cast<const float4*>(beta + c));
    float4 y;
    y.x = (v[j].x - mean) * rstd * g.x + bb.x;
    y.y = (v[j].y - mean) * rstd * g.y + bb.y;
    y.z = (v[j].z - mean) * rstd * g.z + bb.z;
    y.w = (v[j].w - mean) * rstd * g.w + bb.w;
    if (HEAD == 0) {
      if (!live) y = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
      uint2 h, l;
      split4(y, &h, &l);
      *reinterpret_cast<uint2*>(hi + c) = h;
      *reinterpret_cast<uint2*>(lo + c) = l;
    } else {
      const float4 w = __ldg(reinterpret_cast<const float4*>(head_w + c));
      dot += (y.x * w.x + y.y * w.y) + (y.z * w.z + y.w * w.w);
    }
  }
  return HEAD == 1 ? warp_sum(dot) : 0.0f;
}

// Output mode of the duration head (layers/duration_predictor.py:79-88) for one row; call from one lane.
__device__ __forceinline__ void duration_head_store(float dot, const float* __restrict__ head_b, bool live, int mode,
                                                    float offset, void* __restrict__ out, size_t row) {
  float r = dot + head_b[0];
  if (mode == 0) {
    reinterpret_cast<float*>(out)[row] = live ? r : 0.0f;
  } else if (mode == 1) {
    r = fmaxf(__fsub_rn(expf(r), offset), 0.0f);
    reinterpret_cast<float*>(out)[row] = live ? r : 0.0f;
  } else {
    r = fmaxf(rintf(__fsub_rn(expf(r), offset)), 0.0f);
    reinterpret_cast<long long*>(out)[row] = live ? static_cast<long long>(r) : 0ll;
  }
}

template <int HEAD>
__global__ void layernorm_kernel(const float* __restrict__ x, size_t rows, int T,
                                 const float* __restrict__ gamma, const float* __restrict__ beta,
                                 __half* __restrict__ hi, __half* __restrict__ lo,
                                 const float* __restrict__ head_w, const float* __restrict__ head_b,
                                 const int* __restrict__ lens, int mode, float offset,
                                 void* __restrict__ out) {
  constexpr int C = 512;
  const int lane = threadIdx.x & 31;
  const size_t row = blockIdx.x * static_cast<size_t>(blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const float4* xr = reinterpret_cast<const float4*>(x + row * C);
  float4 v[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) v[j] = xr[j * 32 + lane];
  const bool live = lens == nullptr || static_cast<int>(row % T) < lens[row / T];
  const float dot = layernorm_row<HEAD>(v, lane, gamma, beta, HEAD == 0 ? hi + row * C : nullptr,
                                        HEAD == 0 ? lo + row * C : nullptr, head_w, live);
  if (HEAD == 1 && lane == 0) duration_head_store(dot, head_b, live, mode, offset, out, row);
}

// ------------------------------------------------------------------------------------------------
// Masked losses (losses/fastspeech_loss.py:54-67 with use_masking=True; duration target
// models/efficient_tts.py:204,215-216).  acc[0] += sum (mel_pred - speech)^2 over valid frames,
// acc[1] += sum |d - log(delta_e + offset)| over valid tokens; double accumulation.
__global__ void loss_partial_kernel(const float* __restrict__ mel_pred, const float* __restrict__ speech,
                                    const int* __restrict__ sl, int T2, int odim,
                                    const float* __restrict__ dur, const float* __restrict__ e,
                                    const int* __restrict__ tl, int T1, int B, float offset,
                                    int use_masking, double* __restrict__ acc) {
  double sq = 0.0, ab = 0.0;
  const int row4 = odim >> 2;
  const size_t total4 = static_cast<size_t>(B) * T2 * row4;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total4;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const size_t r = i / row4;
    const int b = static_cast<int>(r / T2);
    const int t = static_cast<int>(r % T2);
    if (use_masking && t >= sl[b]) continue;
    const float4 a = reinterpret_cast<const float4*>(mel_pred)[i];
    const float4 c = __ldg(reinterpret_cast<const float4*>(speech) + i);
    const float d0 = a.x - c.x, d1 = a.y - c.y, d2 = a.z - c.z, d3 = a.w - c.w;
    sq += static_cast<double>(d0 * d0) + static_cast<double>(d1 * d1) +
          static_cast<double>(d2 * d2) + static_cast<double>(d3 * d3);
  }
  const size_t ntok = static_cast<size_t>(B) * T1;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < ntok;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int b = static_cast<int>(i / T1);
    const int k = static_cast<int>(i % T1);
    const bool live = k < tl[b];
    if (use_masking && !live) continue;
    const float de = k == 0 ? e[i] : __fsub_rn(e[i], e[i - 1]);
    // pad tokens: both the prediction and the target are masked to 0 before the loss (:216, DP :86)
    const float tgt = live ? logf(__fadd_rn(de, offset)) : 0.0f;
    ab += static_cast<double>(fabsf(__fsub_rn(dur[i], tgt)));
  }
  sq = warp_sum(sq);
  ab = warp_sum(ab);
  __shared__ double s_sq[32], s_ab[32];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (lane == 0) { s_sq[w] = sq; s_ab[w] = ab; }
  __syncthreads();
  if (w == 0) {
    const int nw = blockDim.x >> 5;
    sq = lane < nw ? s_sq[lane] : 0.0;
    ab = lane < nw ? s_ab[lane] : 0.0;
    sq = warp_sum(sq);
    ab = warp_sum(ab);
    if (lane == 0) {
      atomicAdd(acc, sq);
      atomicAdd(acc + 1, ab);
    }
  }
}

// scalars = {loss, mel_loss, duration_loss, sum_sq, n_mel, sum_abs, n_tok, flags}
__global__ void loss_finalize_kernel(const double* __restrict__ acc, const int* __restrict__ tl,
                                     const int* __restrict__ sl, int B, int T1, int T2, int odim,
                                     int use_masking, const int* __restrict__ flags,
                                     const int* __restrict__ err_flag, float* __restrict__ scalars) {
  double nt = 0.0, nm = 0.0;
  for (int b = threadIdx.x; b < B; b += 32) {
    nt += use_masking ? tl[b] : T1;
    nm += static_cast<double>(use_masking ? sl[b] : T2) * odim;
  }
  nt = warp_sum(nt);
  nm = warp_sum(nm);
  if (threadIdx.x == 0) {
    const float mel = static_cast<float>(acc[0] / nm);
    const float dur = static_cast<float>(acc[1] / nt);
    scalars[0] = mel + dur;
    scalars[1] = mel;
    scalars[2] = dur;
    scalars[3] = static_cast<float>(acc[0]);
    scalars[4] = static_cast<float>(nm);
    scalars[5] = static_cast<float>(acc[1]);
    scalars[6] = static_cast<float>(nt);
    scalars[7] = static_cast<float>(flags[0] | err_flag[0]);
  }
}

// ------------------------------------------------------------------------------------------------
// inference(): e = cumsum(durations) (models/efficient_tts.py:260; fp32 cumsum accumulated in
// double like torch CPU) and T2 = round_half_even(e[T1-1]) (:361).  One warp, B = 1.
// Batched form (ragged synthesis): one warp per utterance over its first lens[b] tokens; e beyond is 0.
__global__ void duration_cumsum_batch_kernel(const float* __restrict__ d, const int* __restrict__ lens, int B,
                                             int T1, float* __restrict__ e, int* __restrict__ t2_out) {
  const int lane = threadIdx.x & 31;
  const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (b >= B) return;
  const int L = min(max(lens[b], 0), T1);
  const float* db = d + static_cast<size_t>(b) * T1;
  float* eb = e + static_cast<size_t>(b) * T1;
  double carry = 0.0;
  for (int base = 0; base < T1; base += 32) {
    const int i = base + lane;
    double s = i < L ? static_cast<double>(db[i]) : 0.0;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const double n = __shfl_up_sync(0xffffffffu, s, o);
      if (lane >= o) s += n;
    }
    s += carry;
    carry = __shfl_sync(0xffffffffu, s, 31);
    if (i < T1) eb[i] = i < L ? static_cast<float>(s) : 0.0f;
  }
  if (lane == 0) t2_out[b] = L > 0 ? static_cast<int>(rintf(static_cast<float>(carry))) : 0;
}

// One warp: e = cumsum(d[0..T1)), t2_out[0] = round_half_even(e[T1-1]).
__device__ __forceinline__ void duration_cumsum_warp(const float* __restrict__ d, int T1, float* __restrict__ e,
                                                     int* __restrict__ t2_out, int lane) {
  double carry = 0.0;
  float last = 0.0f;
  for (int base = 0; base < T1; base += 32) {
    const int i = base + lane;
    double s = i < T1 ? static_cast<double>(d[i]) : 0.0;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const double n = __shfl_up_sync(0xffffffffu, s, o);
      if (lane >= o) s += n;
    }
    s += carry;
    carry = __shfl_sync(0xffffffffu, s, 31);
    if (i < T1) {
      e[i] = static_cast<float>(s);
      if (i == T1 - 1) last = static_cast<float>(s);
    }
  }
  last = __shfl_sync(0xffffffffu, last, (T1 - 1) & 31);
  if (lane == 0) t2_out[0] = static_cast<int>(rintf(last));
}

__global__ void duration_cumsum_kernel(const float* __restrict__ d, int T1, float* __restrict__ e,
                                       int* __restrict__ t2_out) {
  duration_cumsum_warp(d, T1, e, t2_out, threadIdx.x);
}

// ------------------------------------------------------------------------------------------------
// Length regulator (layers/length_regulator.py:35-79).
// plan: ds_eff = ds (alpha == 1) or round_half_even(float(ds) * alpha) (:49-50); rows whose valid
// durations sum to 0 become all ones (:76-78; written through to the caller's ds when alpha == 1,
// because the reference works on views); out_lens[b] = sum; plan[0] = max_b out_lens (atomicMax),
// plan[1] |= 1 on a negative duration.  One warp per row.
__global__ void length_regulator_plan_kernel(long long* __restrict__ ds, const long long* __restrict__ ilens,
                                             float alpha, int alpha_is_one, int B, int T1,
                                             long long* __restrict__ ds_eff, long long* __restrict__ out_lens,
                                             long long* __restrict__ plan) {
  const int lane = threadIdx.x & 31;
  const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (b >= B) return;
  const int n = static_cast<int>(min(static_cast<long long>(T1), max(0ll, ilens[b])));
  long long* drow = ds + static_cast<size_t>(b) * T1;
  long long* erow = ds_eff + static_cast<size_t>(b) * T1;
  long long sum = 0;
  int neg = 0;
  for (int i = lane; i < T1; i += 32) {
    long long d = drow[i];
    if (!alpha_is_one) d = static_cast<long long>(rintf(__fmul_rn(static_cast<float>(d), alpha)));
    erow[i] = d;
    if (i < n) {
      sum += d;
      neg |= d < 0;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    sum += __shfl_xor_sync(0xffffffffu, sum, o);
    neg |= __shfl_xor_sync(0xffffffffu, neg, o);
  }
  __syncwarp();
  if (sum == 0 && n > 0) {
    for (int i = lane; i < n; i += 32) {
      erow[i] = 1;
      if (alpha_is_one) drow[i] = 1;
    }
    sum = n;
  }
  if (lane == 0) {
    out_lens[b] = sum;
    atomicMax(reinterpret_cast<unsigned long long*>(plan), static_cast<unsigned long long>(max(sum, 0ll)));
    if (neg) atomicOr(reinterpret_cast<unsigned long long*>(plan + 1), 1ull);
  }
}

// fwd: per row an int64 inclusive scan of the durations into shared memory, then every output frame
// j finds its source token idx = #{i : cumsum_i <= j} by binary search and copies the D-float row
// with 16-byte accesses.  grid (ceil(Tout / FR), B); block 256.
constexpr int LR_FRAMES = 64;
__global__ void length_regulator_fwd_kernel(const float* __restrict__ xs, const long long* __restrict__ ds_eff,
                                            const long long* __restrict__ ilens,
                                            const long long* __restrict__ out_lens, int T1, int D,
                                            long long Tout, float pad_value, float* __restrict__ out,
                                            long long* __restrict__ idx) {
  extern __shared__ long long cs[];   // [T1] inclusive cumsum of valid durations
  __shared__ long long warp_tot[32];
  __shared__ int src[LR_FRAMES];
  const int b = blockIdx.y;
  const int n = static_cast<int>(min(static_cast<long long>(T1), max(0ll, ilens[b])));
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
  // block-wide inclusive scan, chunk of blockDim.x at a time
  long long carry = 0;
  for (int base = 0; base < n; base += blockDim.x) {
    const int i = base + threadIdx.x;
    long long v = i < n ? ds_eff[static_cast<size_t>(b) * T1 + i] : 0;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const long long u = __shfl_up_sync(0xffffffffu, v, o);
      if (lane >= o) v += u;
    }
    if (lane == 31) warp_tot[w] = v;
    __syncthreads();
    long long pre = carry;
    for (int k = 0; k < w; ++k) pre += warp_tot[k];
    if (i < n) cs[i] = v + pre;
    long long tot = 0;
    for (int k = 0; k < nw; ++k) tot += warp_tot[k];
    carry += tot;
    __syncthreads();
  }
  const long long len = out_lens[b];
  const long long j0 = static_cast<long long>(blockIdx.x) * LR_FRAMES;
  for (int f = threadIdx.x; f < LR_FRAMES; f += blockDim.x) {
    const long long j = j0 + f;
    int s = -1;
    if (j < len) {
      int lo = 0, hi = n;                      // first i with cs[i] > j
      while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (cs[mid] <= j) lo = mid + 1; else hi = mid;
      }
      s = lo;
    }
    src[f] = s;
    if (idx != nullptr && j < Tout) idx[static_cast<size_t>(b) * Tout + j] = s;
  }
  __syncthreads();
  if ((D & 3) == 0) {
    const int d4 = D >> 2;
    const float4 padv = make_float4(pad_value, pad_value, pad_value, pad_value);
    // one warp per output frame: the row is one contiguous run on both sides (512 B per warp access), no
    // index arithmetic per element; four independent 16-byte loads are in flight per lane before the stores
    for (int f = w; f < LR_FRAMES; f += nw) {
      const long long j = j0 + f;
      if (j >= Tout) break;
      const int s = src[f];
      const float4* sp = reinterpret_cast<const float4*>(xs + (static_cast<size_t>(b) * T1 + max(s, 0)) * D);
      float4* dp = reinterpret_cast<float4*>(out + (static_cast<size_t>(b) * Tout + j) * D);
      for (int c0 = 0; c0 < d4; c0 += 128) {
        float4 v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int c = c0 + u * 32 + lane;
          v[u] = (s >= 0 && c < d4) ? __ldg(sp + c) : padv;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int c = c0 + u * 32 + lane;
          if (c < d4) dp[c] = v[u];
        }
      }
    }
  } else {   // D not a multiple of 4: scalar copies (never on the EFTS path, D = 512)
    for (int k = threadIdx.x; k < LR_FRAMES * D; k += blockDim.x) {
      const int f = k / D, c = k % D;
      const long long j = j0 + f;
      if (j >= Tout) break;
      const int s = src[f];
      out[(static_cast<size_t>(b) * Tout + j) * D + c] =
          s >= 0 ? xs[(static_cast<size_t>(b) * T1 + s) * D + c] : pad_value;
    }
  }
}

}  // namespace efts
