// Split-fp16 "3-pass" tap-GEMM on tcgen05 tensor cores (sm_100a).
//
// Computes, for every position row m = (b, t) and output column n,
//
//     acc[m, n] = sum_{tap < ntaps} sum_{k < K}  A[b, t + tap - pad, k] * W[z(tap, b), n, k]
//
// which covers the reference's Conv1d residual layers (layers/efts_modules.py:32-36,50: ntaps=5,
// pad=2), the duration-predictor convs (layers/duration_predictor.py:58: ntaps=3, pad=1), every
// torch.nn.Linear on the path (ntaps=1) and the two batched matmuls of the alignment block
// (models/efficient_tts.py:390 and :190: z = b).
//
// fp32 parity on fp16 tensor cores: every fp32 operand x is stored as two fp16 planes,
// hi = fp16(x) and lo = fp16((x - hi) * 2^11), and the product is assembled from three MMAs,
//     acc0 += Ahi*Bhi            acc1 += Ahi*Blo + Alo*Bhi          acc = acc0 + 2^-11 * acc1
// with both accumulators in fp32 tensor memory (hard part 1).
//
// Roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM owner + MMA issuer,
// warps 2..5 = epilogue (one thread per accumulator row / TMEM lane).
// Operands are staged by TMA into 128-byte-swizzled K-major tiles; out-of-range rows
// (t < 0, t >= T: the conv zero padding; k >= K; n >= N) are zero-filled by the TMA unit.
#pragma once
#include <cuda_fp16.h>

#include "sm100_ptx.cuh"

namespace efts {

constexpr int GEMM_BM = 128;        // rows (positions) per CTA tile == TMEM lanes
constexpr int GEMM_BK = 64;         // fp16 per k-block == one 128-byte swizzle row
constexpr int GEMM_THREADS = 192;
constexpr int GEMM_AROWS_SHIFT = 136;   // A box rows when one tile serves every tap (128 + halo, /8)
constexpr float SPLIT_SCALE = 2048.0f;           // 2^11
constexpr float SPLIT_INV_SCALE = 1.0f / 2048.0f;

enum GemmAct { ACT_NONE = 0, ACT_LRELU = 1, ACT_RELU = 2 };

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
  // second-generation kernel only (gemm2_sm100.cuh)
  const int2* tile_list; // compacted live row tiles (b, t0), or nullptr = all B * ceil(T/128)
  const int* tile_count; // device count of tile_list entries
  int chunk_kb;          // k-blocks per main-accumulator flush (0 = never flush)
  // softmax-partial epilogue (energy GEMM): instead of storing the scores, every 128-column tile writes
  // (max, sum exp, sum exp * column, 0) over its columns n < col_lens[b] to softmax_part[(b*T + t) * n_tiles + tile]
  float4* softmax_part;
  const int* col_lens;
  int err_code;          // extra bits OR-ed into err_flag with bit 3 (identifies the launch kind in diagnostics)
  int* err_flag;         // |= 8 when an activation leaves the fp16 operand range (|x| > 65504)
  int debug_mask;        // timing experiments only (results become wrong): 1 = no fp32 store, 2 = no plane stores
  // split reduction (fused-B kernel, small problems): work item w covers accumulation chunk w % splits of tile
  // w / splits and stores its raw fp32 partial at out + (w % splits) * split_stride; splitk_reduce_kernel finishes
  int splits;            // 0 / 1 = off
  size_t split_stride;   // elements between the partial planes
  float* split_scratch;  // host side only: room for the partial planes (kSplitScratchBytes), or nullptr = never split
  // dilated taps / vocoder layers (second-generation kernel only)
  int dil;               // row shift of tap j is (j - pad) * dil; 0 is read as 1
  int plane_act;         // 1: the fp16 operand planes hold LeakyReLU(0.1) of the stored fp32 value (the next layer's
                         // input activation, vocoders/hifigan_model.py:58,123), the fp32 store stays pre-activation
  int long_taps;         // host side only: use the 184-row A box variant (128 + (ntaps - 1) * dil <= 184)
};

template <int BN, int AMODE>
struct GemmCfg {
  // AMODE 0: one 128-row A box per (k-block, tap).  AMODE 1/2: one 136-row A box per k-block,
  // taps read it through row-shifted descriptors (2 additionally sets the descriptor base offset).
  static constexpr int A_ROWS = AMODE == 0 ? GEMM_BM : GEMM_AROWS_SHIFT;
  static constexpr int A_PLANE = A_ROWS * 128;
  static constexpr int B_PLANE = BN * 128;
  static constexpr int B_STAGES = BN == 256 ? 2 : (BN == 128 ? (AMODE == 0 ? 3 : 4) : (AMODE == 0 ? 4 : 6));
  static constexpr int A_STAGES = AMODE == 0 ? B_STAGES : 2;
  static constexpr int SMEM_TILES = A_STAGES * 2 * A_PLANE + B_STAGES * 2 * B_PLANE;
  static constexpr int SMEM_BYTES = SMEM_TILES + 1024 /*align slack*/ + 256 /*barriers*/;
  static constexpr int TMEM_COLS = 2 * BN < 32 ? 32 : 2 * BN;   // BN in {64,128,256} -> pow2
  static_assert(A_PLANE % 1024 == 0 && B_PLANE % 1024 == 0, "swizzle atoms need 1024B alignment");
  static_assert(SMEM_BYTES <= 232448, "exceeds 227 KB");
};

__device__ __forceinline__ void split_store8(__half* hi, __half* lo, const float* v) {
  __align__(16) __half h[8];
  __align__(16) __half l[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    h[j] = __float2half_rn(v[j]);
    l[j] = __float2half_rn((v[j] - __half2float(h[j])) * SPLIT_SCALE);
  }
  *reinterpret_cast<uint4*>(hi) = *reinterpret_cast<const uint4*>(h);
  *reinterpret_cast<uint4*>(lo) = *reinterpret_cast<const uint4*>(l);
}

template <int BN, int AMODE>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_split_kernel(const __grid_constant__ CUtensorMap tmA_hi, const __grid_constant__ CUtensorMap tmA_lo,
                  const __grid_constant__ CUtensorMap tmB_hi, const __grid_constant__ CUtensorMap tmB_lo,
                  const GemmParams p) {
  using Cfg = GemmCfg<BN, AMODE>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sA = smem_base;
  const uint32_t sB = sA + Cfg::A_STAGES * 2 * Cfg::A_PLANE;
  const uint32_t sBar = sB + Cfg::B_STAGES * 2 * Cfg::B_PLANE;
  // barrier slots (8 B each): fullA[4] emptyA[4] fullB[8] emptyB[8] tmem_full[1]; tmem ptr after
  auto fullA = [&](int s) { return sBar + 8u * s; };
  auto emptyA = [&](int s) { return sBar + 32u + 8u * s; };
  auto fullB = [&](int s) { return sBar + 64u + 8u * s; };
  auto emptyB = [&](int s) { return sBar + 128u + 8u * s; };
  const uint32_t tmem_full = sBar + 192u;
  const uint32_t tmem_slot = sBar + 200u;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  // grid.x = n-tiles (fastest, so the CTAs sharing one A tile run together), grid.y = row tiles
  // of one utterance, grid.z = utterance.
  const int b = blockIdx.z;
  const int t0 = blockIdx.y * GEMM_BM;
  const int n0 = blockIdx.x * BN;
  const int num_kb = (p.K + GEMM_BK - 1) / GEMM_BK;
  if (p.skip_lens != nullptr && t0 >= p.skip_lens[b] + p.skip_halo) return;   // CTA-uniform

  if (threadIdx.x == 0) {
    for (int s = 0; s < Cfg::A_STAGES; ++s) { ptx::mbar_init(fullA(s), 1); ptx::mbar_init(emptyA(s), 1); }
    for (int s = 0; s < Cfg::B_STAGES; ++s) { ptx::mbar_init(fullB(s), 1); ptx::mbar_init(emptyB(s), 1); }
    ptx::mbar_init(tmem_full, 1);
    ptx::fence_barrier_init();
  }
  if (warp == 1) ptx::tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    if (lane == 0) {
      ptx::prefetch_tensormap(&tmA_hi); ptx::prefetch_tensormap(&tmA_lo);
      ptx::prefetch_tensormap(&tmB_hi); ptx::prefetch_tensormap(&tmB_lo);
      int ia = 0, ib = 0;
      for (int kb = 0; kb < num_kb; ++kb) {
        if (AMODE != 0) {
          const int s = ia % Cfg::A_STAGES; const uint32_t ph = (ia / Cfg::A_STAGES) & 1;
          ptx::mbar_wait(emptyA(s), ph ^ 1u);
          ptx::mbar_expect_tx(fullA(s), 2 * Cfg::A_PLANE);
          const uint32_t dst = sA + s * 2 * Cfg::A_PLANE;
          ptx::tma_load_3d(&tmA_hi, fullA(s), dst, kb * GEMM_BK, t0 - p.pad, b);
          ptx::tma_load_3d(&tmA_lo, fullA(s), dst + Cfg::A_PLANE, kb * GEMM_BK, t0 - p.pad, b);
          ++ia;
        }
        for (int tap = 0; tap < p.ntaps; ++tap) {
          if (AMODE == 0) {
            const int s = ia % Cfg::A_STAGES; const uint32_t ph = (ia / Cfg::A_STAGES) & 1;
            ptx::mbar_wait(emptyA(s), ph ^ 1u);
            ptx::mbar_expect_tx(fullA(s), 2 * Cfg::A_PLANE);
            const uint32_t dst = sA + s * 2 * Cfg::A_PLANE;
            ptx::tma_load_3d(&tmA_hi, fullA(s), dst, kb * GEMM_BK, t0 + tap - p.pad, b);
            ptx::tma_load_3d(&tmA_lo, fullA(s), dst + Cfg::A_PLANE, kb * GEMM_BK, t0 + tap - p.pad, b);
            ++ia;
          }
          const int s = ib % Cfg::B_STAGES; const uint32_t ph = (ib / Cfg::B_STAGES) & 1;
          ptx::mbar_wait(emptyB(s), ph ^ 1u);
          ptx::mbar_expect_tx(fullB(s), 2 * Cfg::B_PLANE);
          const uint32_t dst = sB + s * 2 * Cfg::B_PLANE;
          const int z = p.b_batched ? b : tap;
          ptx::tma_load_3d(&tmB_hi, fullB(s), dst, kb * GEMM_BK, n0, z);
          ptx::tma_load_3d(&tmB_lo, fullB(s), dst + Cfg::B_PLANE, kb * GEMM_BK, n0, z);
          ++ib;
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc = ptx::make_idesc_f16(GEMM_BM, BN);
      const uint32_t acc0 = tmem_base;           // hi*hi
      const uint32_t acc1 = tmem_base + BN;      // hi*lo + lo*hi
      int ia = 0, ib = 0;
      uint32_t first = 1;
      for (int kb = 0; kb < num_kb; ++kb) {
        int sa = 0;
        if (AMODE != 0) {
          sa = ia % Cfg::A_STAGES;
          ptx::mbar_wait(fullA(sa), (ia / Cfg::A_STAGES) & 1);
          ++ia;
        }
        for (int tap = 0; tap < p.ntaps; ++tap) {
          if (AMODE == 0) {
            sa = ia % Cfg::A_STAGES;
            ptx::mbar_wait(fullA(sa), (ia / Cfg::A_STAGES) & 1);
            ++ia;
          }
          const int sb = ib % Cfg::B_STAGES;
          ptx::mbar_wait(fullB(sb), (ib / Cfg::B_STAGES) & 1);
          ++ib;
          ptx::tc_fence_after();
          const uint32_t a_addr = sA + sa * 2 * Cfg::A_PLANE + (AMODE == 0 ? 0 : tap * 128);
          const uint32_t b_addr = sB + sb * 2 * Cfg::B_PLANE;
          const uint32_t boff = AMODE == 2 ? ((a_addr >> 7) & 7u) : 0u;
          const uint64_t dAh = ptx::make_desc_sw128(a_addr, boff);
          const uint64_t dAl = ptx::make_desc_sw128(a_addr + Cfg::A_PLANE, boff);
          const uint64_t dBh = ptx::make_desc_sw128(b_addr, 0);
          const uint64_t dBl = ptx::make_desc_sw128(b_addr + Cfg::B_PLANE, 0);
#pragma unroll
          for (int k = 0; k < GEMM_BK / 16; ++k) {
            const uint64_t ko = static_cast<uint64_t>(k * 2);   // 32 bytes >> 4 per K=16 step
            ptx::mma_f16_ss(acc0, dAh + ko, dBh + ko, idesc, first ? 0u : 1u);
            ptx::mma_f16_ss(acc1, dAh + ko, dBl + ko, idesc, first ? 0u : 1u);
            ptx::mma_f16_ss(acc1, dAl + ko, dBh + ko, idesc, 1u);
            first = 0;
          }
          ptx::tc_commit(emptyB(sb));
          if (AMODE == 0) ptx::tc_commit(emptyA(sa));
        }
        if (AMODE != 0) ptx::tc_commit(emptyA(sa));
      }
      ptx::tc_commit(tmem_full);
    }
  } else {
    // ------------------------------------------------------------ epilogue (warps 2..5)
    const int q = warp & 3;                       // TMEM lane quadrant this warp may read
    const int row = q * 32 + lane;
    const int t = t0 + row;
    const bool row_ok = t < p.T;
    const bool row_live = row_ok && (p.lens == nullptr || t < p.lens[b]);
    const size_t m = static_cast<size_t>(b) * p.T + (row_ok ? t : 0);
    ptx::mbar_wait(tmem_full, 0);
    ptx::tc_fence_after();
    const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
#pragma unroll 1
    for (int c0 = 0; c0 < BN; c0 += 32) {
      if (n0 + c0 >= p.N) break;                  // warp-uniform
      __syncwarp();                               // tcgen05.ld is .aligned: reconverge after waits / skips
      uint32_t r0[32], r1[32];
      ptx::tmem_ld_32x32(lane_addr + c0, r0);
      ptx::tmem_ld_32x32(lane_addr + BN + c0, r1);
      ptx::tmem_ld_wait();
      if (!row_ok) continue;
      float v[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        float a = __uint_as_float(r0[j]) + __uint_as_float(r1[j]) * SPLIT_INV_SCALE;
        if (p.divisor != 1.0f) a = __fdiv_rn(a, p.divisor);
        v[j] = a;
      }
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        const int n = n0 + c0 + g * 8;
        if (n >= p.N) break;
        float* vv = v + g * 8;
        if (p.bias != nullptr) {
          const float4 b0 = __ldg(reinterpret_cast<const float4*>(p.bias + n));
          const float4 b1 = __ldg(reinterpret_cast<const float4*>(p.bias + n + 4));
          vv[0] += b0.x; vv[1] += b0.y; vv[2] += b0.z; vv[3] += b0.w;
          vv[4] += b1.x; vv[5] += b1.y; vv[6] += b1.z; vv[7] += b1.w;
        }
        if (p.act == ACT_LRELU) {
#pragma unroll
          for (int j = 0; j < 8; ++j) vv[j] = vv[j] > 0.0f ? vv[j] : vv[j] * 0.1f;
        } else if (p.act == ACT_RELU) {
#pragma unroll
          for (int j = 0; j < 8; ++j) vv[j] = fmaxf(vv[j], 0.0f);
        }
        if (p.resid != nullptr) {
          const float* rp = p.resid + m * p.ld_out + n;
          const float4 x0 = *reinterpret_cast<const float4*>(rp);
          const float4 x1 = *reinterpret_cast<const float4*>(rp + 4);
          vv[0] = x0.x + vv[0]; vv[1] = x0.y + vv[1]; vv[2] = x0.z + vv[2]; vv[3] = x0.w + vv[3];
          vv[4] = x1.x + vv[4]; vv[5] = x1.y + vv[5]; vv[6] = x1.z + vv[6]; vv[7] = x1.w + vv[7];
        }
        if (!row_live) {
#pragma unroll
          for (int j = 0; j < 8; ++j) vv[j] = 0.0f;
        }
        if (p.out != nullptr) {
          float* op = p.out + m * p.ld_out + n;
          *reinterpret_cast<float4*>(op) = make_float4(vv[0], vv[1], vv[2], vv[3]);
          *reinterpret_cast<float4*>(op + 4) = make_float4(vv[4], vv[5], vv[6], vv[7]);
        }
        if (p.out_hi != nullptr) {
          split_store8(p.out_hi + m * p.ld_pl + n, p.out_lo + m * p.ld_pl + n, vv);
        }
        if (p.outT_hi != nullptr) {
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const size_t o = (static_cast<size_t>(b) * p.N + (n + j)) * p.ld_t + t;
            const __half h = __float2half_rn(vv[j]);
            p.outT_hi[o] = h;
            p.outT_lo[o] = __float2half_rn((vv[j] - __half2float(h)) * SPLIT_SCALE);
          }
        }
      }
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) ptx::tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
}

}  // namespace efts
