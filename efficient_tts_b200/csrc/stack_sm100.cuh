// Resident layer stack for B = 1 synthesis (models/efficient_tts.py:230-285): ONE persistent kernel walks a list
// of tensor-core layers -- the five text-encoder convs, the value projection, the duration predictor (two convs with
// LayerNorm and the Linear head) and the duration cumsum in phase 1; the six decoder convs and the mel head in
// phase 2 -- with a grid-wide barrier between layers instead of a kernel boundary.
//
// Why: at B = 1 a conv layer is a handful of 128 x 128 tiles of ~4 us of tensor work each.  As separate launches
// (a split-reduction GEMM + a reduce kernel per layer) the layer costs 20 - 28 us of launch latency, barrier /
// tensor-memory set-up, pipeline fill and teardown, and the host spends ~9 us enqueueing each of the ~32 launches:
// 0.49 ms for 9 GFLOP.  Here barriers, tensor memory and the TMA / MMA pipelines are set up once, activations stay in
// L2 between layers, and the host enqueues two kernels.
//
// Per layer, two phases separated by grid barriers:
//   1. GEMM: work item = (128-row tile, 128-column tile, accumulation chunk of `chunk_kb` k-blocks); the item's raw
//      fp32 partial goes to scratch[chunk][T][N].  Same MMAs in the same order as gemm2_kernel's fused-B variant
//      (Ahi x [Bhi ; Blo] as one N = 256 MMA, then Alo x Bhi into the upper half; here on single CTAs, so the
//      [Bhi ; Blo] operand is one 256-row shared-memory tile).
//   2. reduce + epilogue, one warp per row: partials summed in chunk order (the additions the unsplit kernel's
//      running sums perform, in the same order: results are bitwise those of gemm2_kernel / splitk_reduce_kernel /
//      layernorm_kernel), bias, activation, then residual + fp32 / operand-plane stores, or LayerNorm (+ head).
// The reduce phase writes with ordinary stores what the next layer's TMA loads read: every writer issues
// fence.proxy.async before the barrier and the producer after it.
#pragma once
#include "gemm2_sm100.cuh"
#include "path_kernels.cuh"

namespace efts {

constexpr int ST_THREADS = 256;                 // warp 0 TMA, warp 1 MMA, warps 4-7 accumulator drain; all 8 reduce
constexpr int ST_A_PLANE = G2_A_ROWS * 128;     // 136-row A box (128 rows + conv halo), one fp16 plane
constexpr int ST_A_STAGE = 2 * ST_A_PLANE;      // hi + lo
constexpr int ST_B_PLANE = G2_BN * 128;
constexpr int ST_B_STAGE = 2 * ST_B_PLANE;      // [Bhi ; Blo]: one 256-row K-major tile
constexpr int ST_A_STAGES = 2;
constexpr int ST_B_STAGES = 4;                  // of 256 weight rows; with 64-column tiles the same bytes hold 8 stages
constexpr int ST_B_STAGES_MAX = 8;
constexpr int ST_SMEM_TILES = ST_A_STAGES * ST_A_STAGE + ST_B_STAGES * ST_B_STAGE;
constexpr int ST_LAYER_SMEM = 2048;              // shared-memory copy of the layer table
constexpr int ST_SMEM_BYTES = ST_SMEM_TILES + 1024 + 512 + 4 * G2_STAGE_WARP_BYTES + ST_LAYER_SMEM;
constexpr int ST_MAX_LAYERS = 9;
static_assert(ST_A_PLANE % 1024 == 0 && ST_B_PLANE % 1024 == 0, "swizzle atoms need 1024-byte alignment");
static_assert(ST_SMEM_BYTES <= 232448, "exceeds 227 KB");

enum StackMode { ST_PLAIN = 0, ST_LN_PLANES = 1, ST_LN_HEAD = 2 };

struct alignas(64) StackMaps {
  CUtensorMap a_hi, a_lo;      // input operand planes [T, K], box 64 x 136
  CUtensorMap b_hi, b_lo;      // weights [taps, N, K], box 64 x 128
};

// Everything else a layer needs.  Lives in kernel-parameter space; the kernel copies the array to shared memory once
// (indexed constant-bank loads are slow and the reduce phase reads a dozen fields per row).
struct alignas(16) StackLayer {
  int T, K, N, ntaps, pad;
  int chunk_kb;                // k-blocks per accumulation chunk (0 = one chain over all of K)
  int act;                     // GemmAct after the bias
  int mode;                    // StackMode
  int err_code;                // OR-ed into err_flag with bit 3 on an operand-range violation
  int ld_t;
  float head_offset;
  int head_mode;               // duration head: 0 log, 1 clamp(exp - offset, 0), 2 rounded int64
  const float* bias;           // [N]
  const float* resid;          // fp32 [T, N] or nullptr
  float* out;                  // fp32 [T, N] or nullptr
  __half* out_hi;              // operand planes [T, N] or nullptr
  __half* out_lo;
  __half* outT_hi;             // transposed operand planes [N, ld_t] or nullptr
  __half* outT_lo;
  const float* ln_g;           // LayerNorm over the N = 512 channels (modes 1, 2)
  const float* ln_b;
  const float* head_w;
  const float* head_b;
  void* head_out;
};

static_assert(sizeof(StackLayer) % 16 == 0, "copied to shared memory 16 bytes at a time");

struct StackParams {
  StackMaps maps[ST_MAX_LAYERS];
  StackLayer layer[ST_MAX_LAYERS];
  int n_layers;
  int bn;                      // column tile of this launch: 128, or 64 when that still fits one wave of CTAs (twice the
                               // work items of half the MMA depth each; same per-element arithmetic)
  float* scratch;              // partial planes [chunk][T][N]
  size_t split_stride;         // elements between partial planes (>= max T * N)
  unsigned* sync;              // grid barrier counter (monotonic across launches)
  unsigned sync_base;          // its value when this launch starts
  int* err_flag;
  long long* trace;            // measurement hook: clock64 of CTA 0 at every phase boundary (nullptr = off)
  // prologue: embedding gather (models/efficient_tts.py:246) when text != nullptr
  const int64_t* text;
  const float* emb;
  int num_symbols, T_embed;
  float* x0_f; __half* x0_hi; __half* x0_lo;
  int* flags;                  // |= 4 token id out of range, |= 8 operand range, |= 64 grid barrier timed out
  // epilogue: e = cumsum(durations), T2 = round(e[-1]) (:260, :361) when dur != nullptr
  const float* dur; float* e; int* t2_out; int T_cumsum;
};

// clock64 that cannot be scheduled before the listed values exist (measurement stamps)
__device__ __forceinline__ long long clock_after(float a, float b, float c, float d) {
  long long t;
  asm volatile("mov.u64 %0, %%clock64;" : "=l"(t) : "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
  return t;
}


// All CTAs of the grid are resident (grid <= SM count, one CTA per SM by shared memory).
// PROXY = true: the phase before the barrier wrote, with ordinary stores, operand planes that TMA loads (async proxy)
// read after it: every writer orders its stores against the async proxy before arriving (the TMA producer issues
// the matching fence before its first load of the next layer).
// One releasing reduction and relaxed polling per CTA; the acquire fence is paid once, after the last arrival.
template <bool PROXY>
__device__ __forceinline__ void stack_grid_barrier(unsigned* ctr, unsigned& target, bool& dead, int* err_flag) {
  if (PROXY) asm volatile("fence.proxy.async.global;" ::: "memory");
  __syncthreads();
  target += gridDim.x;
  if (threadIdx.x == 0) {
    asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(ctr) : "memory");
    const long long t0 = clock64();
    while (!dead) {
      unsigned v;
      asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(ctr) : "memory");
      if (static_cast<int>(v - target) >= 0) break;
      // A protocol bug must not hang the GPU: give up after ~2 s, flag it (bit 6), and let this CTA run through the
      // remaining barriers without waiting.  (Not __trap(): an abort edge inside the layer loop makes the compiler
      // treat the loop's state as divergent, and the MMA warp's descriptor arithmetic leaves the uniform datapath.)
      if (clock64() - t0 > 4000000000LL) {
        dead = true;
        if (err_flag != nullptr) atomicOr(err_flag, 64);
      }
    }
    asm volatile("fence.acq_rel.gpu;" ::: "memory");
  }
  __syncthreads();
}

__global__ void __launch_bounds__(ST_THREADS, 1)
stack_kernel(const __grid_constant__ StackParams sp) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sA = smem_base;
  const uint32_t sB = sA + ST_A_STAGES * ST_A_STAGE;
  const uint32_t sBar = sB + ST_B_STAGES * ST_B_STAGE;
  auto fullA = [&](int s) { return sBar + 8u * s; };
  auto emptyA = [&](int s) { return sBar + 32u + 8u * s; };
  auto fullB = [&](int s) { return sBar + 64u + 8u * s; };
  auto emptyB = [&](int s) { return sBar + 128u + 8u * s; };
  auto acc_full = [&](int s) { return sBar + 192u + 8u * s; };
  auto acc_empty = [&](int s) { return sBar + 224u + 8u * s; };
  const uint32_t tmem_slot = sBar + 256u;
  const uint32_t sStage = sBar + 512u;            // per-drain-warp transposition buffers [4][32 rows][144 B]
  const StackLayer* layers = reinterpret_cast<const StackLayer*>(smem_raw + (sStage + 4u * G2_STAGE_WARP_BYTES -
                                                                              ptx::smem_u32(smem_raw)));
  static_assert(sizeof(StackLayer) * ST_MAX_LAYERS <= ST_LAYER_SMEM, "layer table does not fit its shared-memory slot");

  const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0);   // provably warp-uniform role index
  const int lane = threadIdx.x & 31;
  // weight ring: the deeper it is, the more of a layer's weights are in flight while the grid still synchronises
  const uint32_t nb_stages = sp.bn == 64 ? ST_B_STAGES_MAX : ST_B_STAGES;
  const uint32_t nb_shift = sp.bn == 64 ? 3u : 2u;
  const uint32_t b_stage = static_cast<uint32_t>(2 * sp.bn * 128);

  if (threadIdx.x == 0) {
    for (int s = 0; s < ST_A_STAGES; ++s) { ptx::mbar_init(fullA(s), 1); ptx::mbar_init(emptyA(s), 1); }
    for (int s = 0; s < ST_B_STAGES_MAX; ++s) { ptx::mbar_init(fullB(s), 1); ptx::mbar_init(emptyB(s), 1); }
    for (int s = 0; s < 2; ++s) { ptx::mbar_init(acc_full(s), 1); ptx::mbar_init(acc_empty(s), 4); }
    ptx::fence_barrier_init();
  }
  if (warp == 1) ptx::tmem_alloc(tmem_slot, 512);
  {
    const uint4* src = reinterpret_cast<const uint4*>(sp.layer);
    uint4* dst = reinterpret_cast<uint4*>(const_cast<StackLayer*>(layers));
    for (int i = threadIdx.x; i < static_cast<int>(sizeof(StackLayer) * ST_MAX_LAYERS / 16); i += ST_THREADS) dst[i] = src[i];
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
  ptx::pdl_wait();

  unsigned bar_target = sp.sync_base;
  bool bar_dead = false;
  int trace_n = 0;
  auto stamp = [&]() {
    if (sp.trace != nullptr && blockIdx.x == 0 && threadIdx.x == 0) sp.trace[trace_n++] = clock64();
  };
  stamp();
  // role-level stamps of CTA 0 during layer 1 (trace[48..55]): MMA first operands landed / last commit issued,
  // drain accumulator ready / TMEM released / stores issued, reduce loads consumed
  auto stamp_role = [&](int li, int k) {
    if (sp.trace != nullptr && blockIdx.x == 0 && li == 1) sp.trace[48 + k] = clock64();
  };
  const int gwarp = blockIdx.x * (ST_THREADS / 32) + warp;
  const int nwarps = gridDim.x * (ST_THREADS / 32);

  // ---- prologue: embedding rows -> fp32 master + operand planes (embed_kernel's arithmetic)
  if (sp.text != nullptr) {
    constexpr int C = 512;
    for (int row = gwarp; row < sp.T_embed; row += nwarps) {
      long long id = sp.text[row];
      if (id < 0 || id >= sp.num_symbols) {
        if (lane == 0) atomicOr(sp.flags, 4);
        id = 0;
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int c = (j * 32 + lane) * 4;
        const float4 v = __ldg(reinterpret_cast<const float4*>(sp.emb + static_cast<size_t>(id) * C + c));
        if (outside_fp16_range(v)) atomicOr(sp.flags, 8);
        *reinterpret_cast<float4*>(sp.x0_f + static_cast<size_t>(row) * C + c) = v;
        uint2 h, l;
        split4(v, &h, &l);
        *reinterpret_cast<uint2*>(sp.x0_hi + static_cast<size_t>(row) * C + c) = h;
        *reinterpret_cast<uint2*>(sp.x0_lo + static_cast<size_t>(row) * C + c) = l;
      }
    }
    stamp();
    stack_grid_barrier<true>(sp.sync, bar_target, bar_dead, sp.err_flag);
    stamp();
  }

  // ring positions persist across layers (the pipelines are never torn down).  One set per role: the producer's
  // advance under a lane-dependent branch, and a counter shared with it would no longer be warp-uniform for the compiler
  // -- the MMA warp's stage and descriptor arithmetic would leave the uniform datapath (see gemm2_kernel).
  uint32_t p_ia = 0, p_ib = 0;               // TMA producer (lane 0 of warp 0)
  uint32_t m_ia = 0, m_ib = 0, m_it = 0;     // MMA warp
  uint32_t d_it = 0;                         // drain warps
  int pre_b = 0;                 // producer: weight tiles of the coming layer's first item already in flight

  for (int li = 0; li < sp.n_layers; ++li) {
    const StackLayer& L = layers[li];           // shared-memory copy: what the reduce phase reads per row
    const StackLayer& Lc = sp.layer[li];        // parameter space: everything that steers control flow (loads through
                                                // the constant bank are warp-uniform for the compiler; shared-memory
                                                // loads are not, and would push the MMA warp off the uniform datapath)
    const StackMaps& M = sp.maps[li];
    const int BN = sp.bn;
    const int ntaps = Lc.ntaps, pad = Lc.pad;
    const int n_nt = (Lc.N + BN - 1) / BN;
    const int n_rt = (Lc.T + G2_BM - 1) / G2_BM;
    const int num_kb = (Lc.K + G2_BK - 1) / G2_BK;
    const int ckb = Lc.chunk_kb < 1 ? num_kb : Lc.chunk_kb;
    const int nchunks = (num_kb + ckb - 1) / ckb;
    const int total = n_rt * n_nt * nchunks;
    const int w_first = blockIdx.x, w_step = gridDim.x;

    if (warp == 0) {
      // ---------------------------------------------------------- TMA producer
      if (lane == 0) {
        if (li == 0) { ptx::prefetch_tensormap(&M.a_hi); ptx::prefetch_tensormap(&M.a_lo); }
        asm volatile("fence.proxy.async.global;" ::: "memory");   // reader side of the barrier's proxy fence
        auto load_b = [&](const StackMaps& Mx, int kb, int n0, int tap) {
          const int s = p_ib & (nb_stages - 1u);
          ptx::mbar_wait(emptyB(s), ((p_ib >> nb_shift) & 1u) ^ 1u);
          const uint32_t dst = sB + s * b_stage;
          ptx::mbar_expect_tx(fullB(s), 2 * BN * 128);              // [Bhi (BN rows) ; Blo (BN rows)], one 2 BN-row tile
          ptx::tma_load_3d(&Mx.b_hi, fullB(s), dst, kb * G2_BK, n0, tap);
          ptx::tma_load_3d(&Mx.b_lo, fullB(s), dst + BN * 128, kb * G2_BK, n0, tap);
          ++p_ib;
        };
        int skip_b = pre_b;                          // weight tiles of the first item that were issued ahead
        for (int w = w_first; w < total; w += w_step) {
          const int ch = w % nchunks, tile = w / nchunks;
          const int n0 = (tile % n_nt) * BN, t0 = (tile / n_nt) * G2_BM;
          const int kb1 = min(ch * ckb + ckb, num_kb);
          for (int kb = ch * ckb; kb < kb1; ++kb) {
            {
              const int s = p_ia % ST_A_STAGES;
              ptx::mbar_wait(emptyA(s), ((p_ia / ST_A_STAGES) & 1u) ^ 1u);
              const uint32_t dst = sA + s * ST_A_STAGE;
              ptx::mbar_expect_tx(fullA(s), ST_A_STAGE);
              ptx::tma_load_3d(&M.a_hi, fullA(s), dst, kb * G2_BK, t0 - pad, 0);
              ptx::tma_load_3d(&M.a_lo, fullA(s), dst + ST_A_PLANE, kb * G2_BK, t0 - pad, 0);
              ++p_ia;
            }
            for (int tap = 0; tap < ntaps; ++tap) {
              if (skip_b > 0) { --skip_b; continue; }
              load_b(M, kb, n0, tap);
            }
          }
        }
        // The next layer's weights do not depend on this layer's output: its first weight tiles are requested now and
        // land while the grid reduces and synchronises; only the activation loads wait for the barrier.
        pre_b = 0;
        if (li + 1 < sp.n_layers) {
          const StackLayer& Ln = sp.layer[li + 1];
          const StackMaps& Mn = sp.maps[li + 1];
          ptx::prefetch_tensormap(&Mn.b_hi); ptx::prefetch_tensormap(&Mn.b_lo);
          ptx::prefetch_tensormap(&Mn.a_hi); ptx::prefetch_tensormap(&Mn.a_lo);
          const int nn_nt = (Ln.N + BN - 1) / BN, nn_rt = (Ln.T + G2_BM - 1) / G2_BM;
          const int nnum_kb = (Ln.K + G2_BK - 1) / G2_BK;
          const int nckb = Ln.chunk_kb < 1 ? nnum_kb : Ln.chunk_kb;
          const int nnch = (nnum_kb + nckb - 1) / nckb;
          const int w = blockIdx.x;
          if (w < nn_rt * nn_nt * nnch) {
            const int ch = w % nnch, n0 = ((w / nnch) % nn_nt) * BN;
            const int kb0 = ch * nckb, kb1 = min(kb0 + nckb, nnum_kb);
            const int avail = (kb1 - kb0) * Ln.ntaps;
            for (int q = 0; q < min(static_cast<int>(nb_stages), avail); ++q) {
              load_b(Mn, kb0 + q / Ln.ntaps, n0, q % Ln.ntaps);
              ++pre_b;
            }
          }
        }
      }
    } else if (warp == 1) {
      // ---------------------------------------------------------- MMA issuer: the whole warp walks the schedule so that
      // the descriptor arithmetic stays on the uniform datapath (see gemm2_kernel), lane 0 issues
      {
        const bool issuer = lane == 0;
        // working copies the compiler can see are warp-uniform (the persistent ones merge across the role branches)
        uint32_t ia = __shfl_sync(0xffffffffu, m_ia, 0), ib = __shfl_sync(0xffffffffu, m_ib, 0);
        uint32_t it = __shfl_sync(0xffffffffu, m_it, 0);
        const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
        // ring geometry re-derived from the constant bank (the kernel-scope copies live in per-thread registers)
        const uint32_t nbm = sp.bn == 64 ? 7u : 3u, nbs = sp.bn == 64 ? 3u : 2u, bsh = sp.bn == 64 ? 14u : 15u;
        const uint32_t idesc_n256 = ptx::make_idesc_f16(G2_BM, 2 * BN);     // Ahi x [Bhi ; Blo]
        const uint32_t idesc_n128 = ptx::make_idesc_f16(G2_BM, BN);         // Alo x Bhi
        for (int w = w_first; w < total; w += w_step, ++it) {
          const int ch = w % nchunks;
          const int kb1 = min(ch * ckb + ckb, num_kb);
          const uint32_t buf = it & 1u;
          ptx::mbar_wait(acc_empty(buf), ((it >> 1) & 1u) ^ 1u);
          ptx::tc_fence_after();
          const uint32_t acc = tmem_u + buf * (2 * G2_BN);
          uint32_t first = 1;
          for (int kb = ch * ckb; kb < kb1; ++kb) {
            const int sa = ia % ST_A_STAGES;
            ptx::mbar_wait(fullA(sa), (ia / ST_A_STAGES) & 1u);
            ++ia;
            for (int tap = 0; tap < ntaps; ++tap) {
              const uint32_t sb = ib & nbm;
              ptx::mbar_wait(fullB(sb), (ib >> nbs) & 1u);
              ++ib;
              ptx::tc_fence_after();
              if (first && issuer) stamp_role(li, 0);
              if (issuer && sp.trace != nullptr && blockIdx.x == 0 && li == 1 && (kb - ch * ckb) * ntaps + tap < 12)
                sp.trace[36 + (kb - ch * ckb) * ntaps + tap] = clock64();     // operands of step i landed
              const uint32_t a_addr = sA + sa * ST_A_STAGE + tap * 128;       // row shift = tap
              const uint32_t b_addr = sB + (sb << bsh);
              const uint64_t dAh = ptx::make_desc_sw128(a_addr, 0);
              const uint64_t dAl = ptx::make_desc_sw128(a_addr + ST_A_PLANE, 0);
              const uint64_t dB = ptx::make_desc_sw128(b_addr, 0);
              if (issuer) {
#pragma unroll
                for (int k = 0; k < G2_BK / 16; ++k) {
                  const uint64_t ko = static_cast<uint64_t>(k * 2);
                  // columns [0,128): Ahi*Bhi, [128,256): Ahi*Blo + Alo*Bhi
                  ptx::mma_f16_ss(acc, dAh + ko, dB + ko, idesc_n256, (first && k == 0) ? 0u : 1u);
                  ptx::mma_f16_ss(acc + BN, dAl + ko, dB + ko, idesc_n128, 1u);
                }
                ptx::tc_commit(emptyB(sb));
              }
              first = 0;
              __syncwarp();
            }
            if (issuer) ptx::tc_commit(emptyA(sa));
          }
          if (issuer) {
            ptx::tc_commit(acc_full(buf));
            stamp_role(li, 1);
          }
        }
        m_ia = ia; m_ib = ib; m_it = it;
      }
    } else if (warp >= 4) {
      // ---------------------------------------------------------- accumulator drain -> raw partial planes
      // (rolled loops: the whole per-layer code path has to stay resident in the instruction cache, a layer is a few
      // microseconds and an unrolled epilogue is fetch-bound)
      const int q = warp & 3;
      const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
      const uint32_t stg = sStage + static_cast<uint32_t>(q) * G2_STAGE_WARP_BYTES;
      const int sub_row = lane >> 3, sub_col = (lane & 7) * 4;
      for (int w = w_first; w < total; w += w_step, ++d_it) {
        const int ch = w % nchunks, tile = w / nchunks;
        const int n0 = (tile % n_nt) * BN, t0 = (tile / n_nt) * G2_BM;
        const uint32_t buf = d_it & 1u;
        float* part = sp.scratch + static_cast<size_t>(ch) * sp.split_stride;
        ptx::mbar_wait(acc_full(buf), (d_it >> 1) & 1u);
        ptx::tc_fence_after();
        if (q == 0 && lane == 0) stamp_role(li, 2);
#pragma unroll 1
        for (int c32 = 0; c32 < BN / 32; ++c32) {
          __syncwarp();                               // the previous block's transposed reads are done
          uint32_t r0[32], r1[32];
          ptx::tmem_ld_32x32(lane_addr + buf * (2 * G2_BN) + c32 * 32, r0);
          ptx::tmem_ld_32x32(lane_addr + buf * (2 * G2_BN) + BN + c32 * 32, r1);
          ptx::tmem_ld_wait();
          if (c32 == BN / 32 - 1) {                   // everything this warp needs has left tensor memory
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(acc_empty(buf));
            if (q == 0 && lane == 0) stamp_role(li, 3);
          }
#pragma unroll
          for (int k4 = 0; k4 < 8; ++k4) {
            float vv[4];
#pragma unroll
            for (int j = 0; j < 4; ++j)             // gemm2_kernel's chunk fold: one rounding, then 0 + x
              vv[j] = __fadd_rn(0.0f, __fmaf_rn(__uint_as_float(r1[k4 * 4 + j]), SPLIT_INV_SCALE,
                                                __uint_as_float(r0[k4 * 4 + j])));
            asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};"
                         ::"r"(stg + lane * G2_STAGE_ROW_BYTES + k4 * 16), "f"(vv[0]), "f"(vv[1]), "f"(vv[2]), "f"(vv[3]) : "memory");
          }
          __syncwarp();
          // transposed store (see g2_store_block32): eight lanes write one row's 128 contiguous bytes
          const int nn = n0 + c32 * 32 + sub_col;
          if (nn < L.N) {
#pragma unroll 2
            for (int itr = 0; itr < 8; ++itr) {
              const int rr = itr * 4 + sub_row;
              const int tr = t0 + q * 32 + rr;
              float4 v;
              asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                           : "r"(stg + rr * G2_STAGE_ROW_BYTES + sub_col * 4) : "memory");
              if (tr < L.T) *reinterpret_cast<float4*>(part + static_cast<size_t>(tr) * L.N + nn) = v;
            }
          }
        }
        if (q == 0 && lane == 0) stamp_role(li, 4);
      }
    }
    __syncwarp();
    __syncthreads();
    stamp();
    stack_grid_barrier<false>(sp.sync, bar_target, bar_dead, sp.err_flag);
    stamp();

    // ------------------------------------------------------------ reduce + epilogue
    // sum of the partial planes in chunk order, four loads in flight at a time, + bias, activation
    auto bias_act = [&](float4 x, int c) -> float4 {
      if (L.bias != nullptr) {
        const float4 bb = __ldg(reinterpret_cast<const float4*>(L.bias + c));
        x.x = __fadd_rn(x.x, bb.x); x.y = __fadd_rn(x.y, bb.y); x.z = __fadd_rn(x.z, bb.z); x.w = __fadd_rn(x.w, bb.w);
      }
      if (L.act == ACT_LRELU) {
        x.x = x.x > 0.0f ? x.x : __fmul_rn(x.x, 0.1f); x.y = x.y > 0.0f ? x.y : __fmul_rn(x.y, 0.1f);
        x.z = x.z > 0.0f ? x.z : __fmul_rn(x.z, 0.1f); x.w = x.w > 0.0f ? x.w : __fmul_rn(x.w, 0.1f);
      } else if (L.act == ACT_RELU) {
        x.x = fmaxf(x.x, 0.0f); x.y = fmaxf(x.y, 0.0f); x.z = fmaxf(x.z, 0.0f); x.w = fmaxf(x.w, 0.0f);
      }
      return x;
    };
    auto reduced = [&](int row, int c) -> float4 {
      const float* p0 = sp.scratch + static_cast<size_t>(row) * L.N + c;
      float4 x = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
#pragma unroll 1
      for (int s0 = 0; s0 < nchunks; s0 += 4) {
        float4 u[4];
#pragma unroll
        for (int s = 0; s < 4; ++s)
          if (s0 + s < nchunks) u[s] = __ldcg(reinterpret_cast<const float4*>(p0 + (s0 + s) * sp.split_stride));
#pragma unroll
        for (int s = 0; s < 4; ++s)
          if (s0 + s < nchunks) {
            x.x = __fadd_rn(x.x, u[s].x); x.y = __fadd_rn(x.y, u[s].y);
            x.z = __fadd_rn(x.z, u[s].z); x.w = __fadd_rn(x.w, u[s].w);
          }
      }
      return bias_act(x, c);
    };
    // residual, fp32 store, operand planes (+ transposed planes) of one (row, four columns) unit
    auto store_plain = [&](int row, int c, float4 r, float4 x) {
      const size_t o = static_cast<size_t>(row) * L.N + c;
      if (L.resid != nullptr) {
        x.x = __fadd_rn(r.x, x.x); x.y = __fadd_rn(r.y, x.y); x.z = __fadd_rn(r.z, x.z); x.w = __fadd_rn(r.w, x.w);
      }
      if (L.out != nullptr) *reinterpret_cast<float4*>(L.out + o) = x;
      if (L.out_hi != nullptr) {
        if (outside_fp16_range(x) && sp.err_flag != nullptr) atomicOr(sp.err_flag, 8 | L.err_code);
        uint2 h, l;
        split4(x, &h, &l);
        *reinterpret_cast<uint2*>(L.out_hi + o) = h;
        *reinterpret_cast<uint2*>(L.out_lo + o) = l;
        if (L.outT_hi != nullptr) {                 // K-major operand of the expansion matmul: [N, ld_t], t contiguous
          const float xs[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const __half hh = __float2half_rn(xs[k]);
            const size_t ot = static_cast<size_t>(c + k) * L.ld_t + row;
            L.outT_hi[ot] = hh;
            L.outT_lo[ot] = __float2half_rn((xs[k] - __half2float(hh)) * SPLIT_SCALE);
          }
        }
      }
    };
    if (sp.trace != nullptr && blockIdx.x == 0 && threadIdx.x == 0 && li == 1) sp.trace[53] = clock64();
    if (L.mode == ST_PLAIN) {
      // one thread per (row, four columns)
      const int n4 = L.N >> 2;
      const int units = L.T * n4;
#pragma unroll 1
      for (int i = blockIdx.x * ST_THREADS + threadIdx.x; i < units; i += gridDim.x * ST_THREADS) {
        const int row = i / n4, c = (i - row * n4) * 4;
        float4 r = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
        if (L.resid != nullptr) r = __ldcg(reinterpret_cast<const float4*>(L.resid + static_cast<size_t>(row) * L.N + c));
        store_plain(row, c, r, reduced(row, c));
      }
    } else {
      // LayerNorm over the 512 channels: one warp per row
#pragma unroll 1
      for (int row = gwarp; row < L.T; row += nwarps) {
        float4 v[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) v[j] = reduced(row, (j * 32 + lane) * 4);
        if (L.mode == ST_LN_PLANES) {
          layernorm_row<0>(v, lane, L.ln_g, L.ln_b, L.out_hi + static_cast<size_t>(row) * L.N,
                           L.out_lo + static_cast<size_t>(row) * L.N, nullptr, true);
        } else {
          const float dot = layernorm_row<1>(v, lane, L.ln_g, L.ln_b, nullptr, nullptr, L.head_w, true);
          if (lane == 0) duration_head_store(dot, L.head_b, true, L.head_mode, L.head_offset, L.head_out, row);
        }
      }
    }
    __syncthreads();
    stamp();
    stack_grid_barrier<true>(sp.sync, bar_target, bar_dead, sp.err_flag);
    stamp();
  }

  if (bar_dead && sp.flags != nullptr) atomicOr(sp.flags, 64);     // thread 0 only: a grid barrier gave up
  // ---- epilogue: duration cumsum and T2 (one warp)
  if (sp.dur != nullptr && blockIdx.x == 0 && warp == 0) duration_cumsum_warp(sp.dur, sp.T_cumsum, sp.e, sp.t2_out, lane);

  __syncwarp();
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) ptx::tmem_dealloc(tmem_base, 512);
}

}  // namespace efts
