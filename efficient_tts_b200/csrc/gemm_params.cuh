// Parameters of the split-fp16 tap-GEMM on tcgen05 tensor cores (gemm2_sm100.cuh).
//
// The kernel computes, for every position row m = (b, t) and output column n,
//
//     acc[m, n] = sum_{tap < ntaps} sum_{k < K}  A[b, t + (tap - pad) * dil, k] * W[z(tap, b), n, k]
//
// which covers the reference's Conv1d residual layers (layers/efts_modules.py:32-36,50: ntaps=5,
// pad=2), the duration-predictor convs (layers/duration_predictor.py:58: ntaps=3, pad=1), every
// torch.nn.Linear on the path (ntaps=1) and the two batched matmuls of the alignment block
// (models/efficient_tts.py:390 and :190: z = b).
//
// fp32 parity on fp16 tensor cores: every fp32 operand x is stored as two fp16 planes,
// hi = fp16(x) and lo = fp16((x - hi) * 2^11), and the product is assembled from three products,
//     acc0 += Ahi*Bhi            acc1 += Ahi*Blo + Alo*Bhi          acc = acc0 + 2^-11 * acc1
// with both accumulators in fp32 tensor memory (hard part 1).
// Operands are staged by TMA into 128-byte-swizzled K-major tiles; out-of-range rows
// (t < 0, t >= T: the conv zero padding; k >= K; n >= N) are zero-filled by the TMA unit.
#pragma once
#include <cuda_fp16.h>
#include <stdint.h>

namespace efts {

constexpr float SPLIT_SCALE = 2048.0f;           // 2^11
constexpr float SPLIT_INV_SCALE = 1.0f / 2048.0f;

// True when any of the four values cannot become an fp16 operand: |x| > 65504, +-inf, or NaN.  Compared on the
// bit patterns (|x| as an unsigned integer is monotone in |x|, and every NaN pattern lies above +inf) because a
// float max drops NaNs.
__device__ __forceinline__ bool outside_fp16_range(const float4 v) {
  const uint32_t m = max(max(__float_as_uint(v.x) & 0x7fffffffu, __float_as_uint(v.y) & 0x7fffffffu),
                         max(__float_as_uint(v.z) & 0x7fffffffu, __float_as_uint(v.w) & 0x7fffffffu));
  return m > 0x477fe000u;   // bits of 65504.0f
}

// ACT_LOGCLAMP: log(max(x, 1e-5)), the dynamic-range compression of the log-mel front-end (datasets/meldataset.py:29-30)
enum GemmAct { ACT_NONE = 0, ACT_LRELU = 1, ACT_RELU = 2, ACT_LOGCLAMP = 3 };

struct GemmParams {
  int B, T;              // A is [B, T, K]; one CTA tile never crosses a batch row
  int K;                 // reduction length per tap
  int N;                 // output columns written (multiple of 8)
  int ntaps, pad;        // row shift of tap j is (j - pad)
  int b_batched;         // 0: B-operand z = tap (weights [ntaps, N, K]); 1: z = b ([B, N, K])
  int act;               // GemmAct
  float divisor;         // acc is divided by this before bias when != 1 (energy / sqrt(D))
  const float* bias;     // [N] or nullptr
  const float* resid;    // fp32 [B, T, ld_out] added after the activation, or nullptr
  const int* lens;       // [B] or nullptr: rows with t >= lens[b] are written as zeros
  const int* skip_lens;  // [B] or nullptr: tiles with t0 >= skip_lens[b] + skip_halo are not computed
  int skip_halo;         //   (their rows cannot reach a valid output; SURVEY.md 7, hard part 2)
  float* out;            // fp32 [B, T, ld_out] or nullptr
  int ld_out;
  __half* out_hi;        // fp16 planes [B, T, ld_pl] or nullptr
  __half* out_lo;
  int ld_pl;
  __half* outT_hi;       // transposed fp16 planes [B, N, ld_t] (t contiguous) or nullptr
  __half* outT_lo;
  int ld_t;
  const int2* tile_list; // compacted live row tiles (b, t0), or nullptr = all B * ceil(T/128)
  const int* tile_count; // device count of tile_list entries
  int chunk_kb;          // k-blocks per main-accumulator flush (0 = never flush)
  // softmax-partial epilogue (energy GEMM): instead of storing the scores, every 128-column tile writes
  // (max, sum exp, sum exp * column, 0) over its columns n < col_lens[b] to softmax_part[(b*T + t) * n_tiles + tile]
  float4* softmax_part;
  const int* col_lens;
  int err_code;          // extra bits OR-ed into err_flag with bit 3 (identifies the launch kind in diagnostics)
  int* err_flag;         // |= 8 when an activation leaves the fp16 operand range (|x| > 65504)
  // split reduction (fused-B kernel, small problems): work item w covers accumulation chunk w % splits of tile
  // w / splits and stores its raw fp32 partial at out + (w % splits) * split_stride; splitk_reduce_kernel finishes
  int splits;            // 0 / 1 = off
  size_t split_stride;   // elements between the partial planes
  float* split_scratch;  // host side only: room for the partial planes (kSplitScratchBytes), or nullptr = never split
  // dilated taps / vocoder layers
  int dil;               // row shift of tap j is (j - pad) * dil; 0 is read as 1
  int plane_act;         // 1: the fp16 operand planes hold LeakyReLU(0.1) of the stored fp32 value (the next layer's
                         // input activation, vocoders/hifigan_model.py:58,123), the fp32 store stays pre-activation
  int bmn_per;           // MN-major B operand (weight gradients): k-blocks per utterance of the activation planes; k-block kb
                         // is rows (kb % bmn_per) * 64 + b_koff .. of utterance kb / bmn_per (0 = B is K-major)
  int amn;               // host side: the A operand is MN-major too: OpA describes activation planes [B, T, C] (M = its channels)
  int bmn_batches;       // utterances of those planes (an invalid M tile of an MN-major A reads utterance bmn_batches: zeros)
  int mag_pairs;         // 1: columns (2f, 2f + 1) are (re_f, im_f) of an STFT; the epilogue writes sqrt(re^2 + im^2 + 1e-9)
                         // as operand planes [B, T, N / 2] (out_hi / out_lo, ld_pl) and nothing else (front-end only)
  int long_taps;         // host side only: use the 184-row A box variant (128 + (ntaps - 1) * dil <= 184)
  // weight-gradient GEMMs (reduction over positions)
  int b_koff;            // offset of the B operand along the reduction: with an MN-major B (bmn_per > 0) a ROW offset of the
                         // TMA box, any integer (the conv tap; rows outside the utterance are zero-filled); with a K-major B
                         // an element offset of the K coordinate, which must be a multiple of 8 (16-byte aligned boxes)
  int split_kb;          // > 0 with splits > 1: every work item covers split_kb k-blocks (several accumulation chunks)
};

}  // namespace efts
