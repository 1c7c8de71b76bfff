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

namespace efts {

constexpr float kSplitScale = 2048.0f;   // 2^11, see gemm_sm100.cuh

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
// flags[0] |= 1 if max(text_lengths) != T1, |= 2 if max(speech_lengths) != T2
__global__ void prep_lengths_kernel(const int64_t* __restrict__ tl, const int64_t* __restrict__ sl, int B,
                                    int T1, int T2, int* __restrict__ tl32, int* __restrict__ sl32,
                                    int* __restrict__ flags) {
  int mt = 0, ms = 0;
  for (int b = threadIdx.x; b < B; b += blockDim.x) {
    const int a = static_cast<int>(tl[b]);
    tl32[b] = a;
    mt = max(mt, a);
    if (sl != nullptr) {
      const int c = static_cast<int>(sl[b]);
      sl32[b] = c;
      ms = max(ms, c);
    }
  }
  __shared__ int s_mt, s_ms;
  if (threadIdx.x == 0) { s_mt = 0; s_ms = 0; }
  __syncthreads();
  atomicMax(&s_mt, mt);
  atomicMax(&s_ms, ms);
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
  if (!(fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w))) <= 65504.0f))
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
    if (fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w))) > 65504.0f && err_flag != nullptr)
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
// B1: scaled-dot-product softmax over tokens fused with the position expectation
// (models/efficient_tts.py:392-398 softmax with pad keys at -inf, :168 pad frames zeroed, :312
// bmm(alpha^T, p) with p[i] = i on valid tokens).  One warp per (b, t) row of the energy matrix
// S[B*T2, ldS] (already divided by sqrt(D)); alpha itself is never written.
__global__ void energy_softmax_expect_kernel(const float* __restrict__ S, int ldS,
                                             const int* __restrict__ tl, const int* __restrict__ sl,
                                             int T2, size_t rows, float* __restrict__ imv_raw) {
  const int lane = threadIdx.x & 31;
  const size_t row = blockIdx.x * static_cast<size_t>(blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int b = static_cast<int>(row / T2);
  const int t = static_cast<int>(row % T2);
  if (t >= sl[b]) {
    if (lane == 0) imv_raw[row] = 0.0f;
    return;
  }
  const int L = tl[b];
  const float* s = S + row * ldS;
  float m = -CUDART_INF_F;
  for (int i = lane; i < L; i += 32) m = fmaxf(m, s[i]);
  m = warp_max(m);
  float den = 0.0f, num = 0.0f;
  for (int i = lane; i < L; i += 32) {
    const float ev = expf(s[i] - m);
    den += ev;
    num = fmaf(ev, static_cast<float>(i), num);
  }
  den = warp_sum(den);
  num = warp_sum(num);
  if (lane == 0) imv_raw[row] = __fdiv_rn(num, den);
}

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
  const int L2 = sl[b];
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
  const int L2 = sl[b];
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
  const int L1 = tl != nullptr ? tl[b] : T1;
  const int L2 = sl != nullptr ? sl[b] : T2;
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

// B4, tiled: same arithmetic as reconstruct_alignment_kernel, organised so that both output layouts are
// written with full coalescing.  One block per (utterance, 64-frame tile): each warp evaluates the
// token-softmax of 8 frames with the tokens spread over its lanes (one expf per element, the exps stay
// in registers between the sum and the normalisation), the 64 x T1 tile is staged in shared memory
// (row stride 65 words: conflict-free both ways) and then stored as rows of the returned fp32 matrix
// [B,T1,T2] (256 B per row segment) and as K-major fp16 hi/lo operand rows [B,T2,ldp] (half2 per lane).
// Frames t >= L2 get zeros in the fp32 matrix; their operand rows are not written (the expansion GEMM
// masks those rows by select).  Requires T1 <= 32 * RT_KMAX.
constexpr int RT_FRAMES = 64;
constexpr int RT_KMAX = 16;
__global__ void __launch_bounds__(256)
reconstruct_alignment_tiled_kernel(const float* __restrict__ e, const int* __restrict__ tl,
                                   const int* __restrict__ sl, int T1, int T2, int ldp, float neg_sigma,
                                   float* __restrict__ R, __half* __restrict__ p_hi, __half* __restrict__ p_lo) {
  extern __shared__ float rt_smem[];
  float* tile = rt_smem;                                   // [ldp][65]
  float* se = rt_smem + static_cast<size_t>(ldp) * (RT_FRAMES + 1);
  const int b = blockIdx.y;
  const int t0 = blockIdx.x * RT_FRAMES;
  const int L1 = tl != nullptr ? tl[b] : T1;
  const int L2 = sl != nullptr ? sl[b] : T2;
  const int nt = min(RT_FRAMES, T2 - t0);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float* Rb = R + static_cast<size_t>(b) * T1 * T2 + t0;
  if (t0 >= L2) {                                          // whole tile is padding
    for (int i = warp; i < T1; i += 8)
      for (int tt = lane; tt < nt; tt += 32) Rb[static_cast<size_t>(i) * T2 + tt] = 0.0f;
    return;
  }
  for (int i = threadIdx.x; i < T1; i += 256) se[i] = e[static_cast<size_t>(b) * T1 + i];
  __syncthreads();
  const int kcount = (L1 + 31) >> 5;                       // lane-strided token chunks that hold valid tokens
  for (int tt = warp; tt < RT_FRAMES; tt += 8) {
    const int t = t0 + tt;
    const bool live = t < L2;
    const float q = live ? static_cast<float>(t) : 0.0f;
    float h[RT_KMAX];
    float m = -CUDART_INF_F;
#pragma unroll
    for (int k = 0; k < RT_KMAX; ++k) {
      h[k] = 0.0f;
      if (k < kcount) {                                    // warp-uniform
        const int i = lane + 32 * k;
        const float d = __fsub_rn(q, se[min(i, T1 - 1)]);
        h[k] = i < L1 ? __fmul_rn(neg_sigma, __fmul_rn(d, d)) : -CUDART_INF_F;
        m = fmaxf(m, h[k]);
      }
    }
    m = warp_max(m);
    float den = 0.0f;
#pragma unroll
    for (int k = 0; k < RT_KMAX; ++k) {
      if (k < kcount) {
        h[k] = expf(h[k] - m);                             // exp(-inf) = 0 on pad tokens
        den += h[k];
      }
    }
    den = warp_sum(den);
    // one correctly rounded reciprocal per frame instead of a division per element (differs from
    // exp / sum by at most one ulp of a value <= 1)
    const float inv = live ? __frcp_rn(den) : 0.0f;
#pragma unroll
    for (int k = 0; k < RT_KMAX; ++k) {
      const int i = lane + 32 * k;
      if (i < ldp) tile[i * (RT_FRAMES + 1) + tt] = (k < kcount) ? __fmul_rn(h[k], inv) : 0.0f;
    }
  }
  __syncthreads();
  for (int i = warp; i < T1; i += 8)
    for (int tt = lane; tt < nt; tt += 32) Rb[static_cast<size_t>(i) * T2 + tt] = tile[i * (RT_FRAMES + 1) + tt];
  const int nlive = p_hi != nullptr ? min(nt, L2 - t0) : 0;
  for (int tt = warp; tt < nlive; tt += 8) {
    const size_t o = (static_cast<size_t>(b) * T2 + t0 + tt) * ldp;
    for (int i2 = lane; 2 * i2 < ldp; i2 += 32) {
      const float a = tile[(2 * i2) * (RT_FRAMES + 1) + tt];
      const float c = tile[(2 * i2 + 1) * (RT_FRAMES + 1) + tt];
      const __half ah = __float2half_rn(a), ch = __float2half_rn(c);
      const __half al = __float2half_rn((a - __half2float(ah)) * kSplitScale);
      const __half cl = __float2half_rn((c - __half2float(ch)) * kSplitScale);
      *reinterpret_cast<__half2*>(p_hi + o + 2 * i2) = __halves2half2(ah, ch);
      *reinterpret_cast<__half2*>(p_lo + o + 2 * i2) = __halves2half2(al, cl);
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
  const int L = tl[b];
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
  float s = 0.0f;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    v[j] = xr[j * 32 + lane];
    s += (v[j].x + v[j].y) + (v[j].z + v[j].w);
  }
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
      if (lens != nullptr && static_cast<int>(row % T) >= lens[row / T]) y = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
      uint2 h, l;
      split4(y, &h, &l);
      *reinterpret_cast<uint2*>(hi + row * C + c) = h;
      *reinterpret_cast<uint2*>(lo + row * C + c) = l;
    } else {
      const float4 w = __ldg(reinterpret_cast<const float4*>(head_w + c));
      dot += (y.x * w.x + y.y * w.y) + (y.z * w.z + y.w * w.w);
    }
  }
  if (HEAD == 1) {
    dot = warp_sum(dot);
    if (lane == 0) {
      float r = dot + head_b[0];
      const int b = static_cast<int>(row / T);
      const int t = static_cast<int>(row % T);
      const bool live = lens == nullptr || t < lens[b];
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
  }
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

__global__ void duration_cumsum_kernel(const float* __restrict__ d, int T1, float* __restrict__ e,
                                       int* __restrict__ t2_out) {
  const int lane = threadIdx.x;
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
    for (int k = threadIdx.x; k < LR_FRAMES * d4; k += blockDim.x) {
      const int f = k / d4, c = k % d4;
      const long long j = j0 + f;
      if (j >= Tout) break;
      const int s = src[f];
      const float4 v =
          s >= 0 ? __ldg(reinterpret_cast<const float4*>(xs + (static_cast<size_t>(b) * T1 + s) * D) + c)
                 : padv;
      reinterpret_cast<float4*>(out + (static_cast<size_t>(b) * Tout + j) * D)[c] = v;
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
