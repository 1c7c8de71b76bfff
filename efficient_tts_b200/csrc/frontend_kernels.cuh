// Log-mel front-end (datasets/meldataset.py:49-82): the element-wise kernels either side of the two
// tensor-core GEMMs.  The STFT is a tap-GEMM over hop-sized chunks of the reflect-padded waveform -- frame f is rows
// f .. f + n_fft/hop - 1 of the chunk matrix, one tap per row, against the windowed DFT basis -- and the mel
// projection is a plain GEMM whose epilogue applies log(clamp(x, 1e-5)).
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "path_kernels.cuh"

namespace efts {

// Frames of an utterance of `len` samples: reflect padding of (n_fft - hop) / 2 on both sides, center = False (:66-70).
__host__ __device__ __forceinline__ int frontend_frames(long long len, int n_fft, int hop) {
  const long long padded = len + 2 * ((n_fft - hop) / 2);
  return padded < n_fft ? 0 : static_cast<int>(1 + (padded - n_fft) / hop);
}

// audio fp32 [B, Lmax] (+ lengths[B] or nullptr = all Lmax) -> operand planes of the chunk matrix [B, R, hop],
// R = Tmax + n_fft / hop - 1: chunk[b, r, k] = y_pad_b[r * hop + k] with y_pad_b the utterance's OWN reflect padding
// (torch.nn.functional.pad(mode='reflect'), :66); zero beyond the utterance's last frame.  Also mel_lengths[b].
// One thread per four samples.  flags |= 16 when a length lies outside [0, Lmax], |= 32 when an utterance is too
// short for the reflect padding (torch raises: padding must be smaller than the input), |= 8 on a sample outside
// the fp16 operand range.
__global__ void frontend_chunk_planes_kernel(const float* __restrict__ audio, const long long* __restrict__ lengths,
                                             int B, int Lmax, int R, int n_fft, int hop, __half* __restrict__ hi,
                                             __half* __restrict__ lo, long long* __restrict__ mel_lengths,
                                             int* __restrict__ lens32, int* __restrict__ flags) {
  const int b = blockIdx.y;
  const int pad = (n_fft - hop) / 2;
  long long len = lengths != nullptr ? lengths[b] : Lmax;
  if (len < 0 || len > Lmax) {
    if (threadIdx.x == 0 && blockIdx.x == 0) atomicOr(flags, 16);
    len = len < 0 ? 0 : Lmax;
  }
  if (len <= pad) {                                   // reflect padding needs pad < len
    if (threadIdx.x == 0 && blockIdx.x == 0) atomicOr(flags, 32);
    len = 0;
  }
  const int frames = len > 0 ? frontend_frames(len, n_fft, hop) : 0;
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    if (mel_lengths != nullptr) mel_lengths[b] = frames;
    lens32[b] = frames;
  }
  const int L = static_cast<int>(len);
  const int used = frames > 0 ? (frames - 1) * hop + n_fft : 0;        // padded samples the frames read
  const size_t row_elems = static_cast<size_t>(R) * hop;
  const float* y = audio + static_cast<size_t>(b) * Lmax;
  for (size_t i4 = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i4 * 4 < row_elems;
       i4 += static_cast<size_t>(gridDim.x) * blockDim.x) {
    float v[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int i = static_cast<int>(i4 * 4) + j;     // index into the padded signal
      float x = 0.0f;
      if (i < used) {
        int s = i - pad;
        if (s < 0) s = -s;
        if (s >= L) s = 2 * (L - 1) - s;
        x = y[s];
      }
      v[j] = x;
    }
    const float4 f = make_float4(v[0], v[1], v[2], v[3]);
    if (outside_fp16_range(f)) atomicOr(flags, 8);
    uint2 h, l;
    split4(f, &h, &l);
    const size_t o = static_cast<size_t>(b) * row_elems + i4 * 4;
    *reinterpret_cast<uint2*>(hi + o) = h;
    *reinterpret_cast<uint2*>(lo + o) = l;
  }
}

}  // namespace efts
