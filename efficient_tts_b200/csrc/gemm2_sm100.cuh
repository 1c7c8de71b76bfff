// The tap-GEMM kernel: persistent, CTA-pair (tcgen05 cta_group::2) capable, with the main
// accumulator flushed to fp32 registers every few k-blocks (contraction and split-fp16 scheme: gemm_params.cuh).
//
//   * one 136-row A box per k-block serves all taps through row-shifted UMMA descriptors;
//   * CG = 2: two CTAs (one TPC) form a 256-row x 128-column tile; each loads its own 128 A rows and
//     half of the W tile, so weight traffic from L2 halves;
//   * persistent CTAs walk a (compacted) list of live row tiles; TMEM holds two 128-column main
//     accumulators and two correction accumulators, so MMAs of the next chunk / tile overlap the
//     drain and the epilogue of the previous one;
//   * tensor-core fp32 accumulation truncates (measured on B200: the error of a 160-step chain is
//     biased towards zero and 16x the noise of an fp32 FMA chain).  The main hi*hi accumulator is
//     therefore restarted every `chunk_kb` k-blocks and the partial sums are added in round-to-nearest
//     fp32 on the CUDA cores; the correction accumulator is 2^-11 down and needs no flushing.
#pragma once
#include <cuda_fp16.h>
#include <math_constants.h>

#include "gemm_params.cuh"
#include "sm100_ptx.cuh"

namespace efts {

constexpr int G2_BM = 128;
constexpr int G2_BN = 128;
constexpr int G2_BK = 64;
constexpr int G2_A_ROWS = 136;                      // A box: 128 rows + the halo of 9 undilated taps
constexpr int G2_A_ROWS_LONG = 184;                 // vocoder layers: up to 15 taps, dilation up to 5 (halo <= 56 rows)
constexpr int G2_A_ROWS_XLONG = 200;                // ResBlock2 generators: k = 7 with dilation 12 (halo 72 rows)
constexpr int G2_EPI_WARPS = 8;                    // two groups of four (one warp per TMEM lane quadrant)
constexpr int G2_THREADS = 128 + 32 * G2_EPI_WARPS;   // warpgroup 0: TMA, MMA, 2 idle warps; warpgroups 1, 2: epilogue groups
constexpr int G2_BIAS_MAX = 1024;                     // columns whose bias is staged in shared memory
constexpr int G2_BIAS_MAX_LONG = 2048;                // long-tap variant (the transposed-conv GEMM has N = stride * C)
constexpr int G2_REGS_CTRL = 72;                     // setmaxnreg budgets: 3 warps per SM sub-partition,
constexpr int G2_REGS_EPI = 216;                     // 32 * (72 + 2 * 216) = 16128 <= 16384 registers (an exact fit, 80 + 2 * 216, fails to launch)

constexpr int G2_STAGE_ROW_BYTES = 144;               // wide epilogue staging: 32 fp32 + pad per row (conflict-free v4 stores)
constexpr int G2_STAGE_WARP_BYTES = 32 * G2_STAGE_ROW_BYTES;

// FUSE = 1 (CTA pairs, long reductions): the products that share the A operand, Ahi*Bhi and Ahi*Blo, are issued as
// ONE N = 256 MMA whose B operand is [Bhi (leader CTA's 128 rows) ; Blo (peer CTA's 128 rows)], the third product
// Alo*Bhi follows as an N = 128 MMA into the upper 128 accumulator columns (same 2^-11 scale as Ahi*Blo).  Two MMAs
// per k-step instead of three, A fetched from shared memory twice instead of three times.  The N = 128 MMA needs
// Bhi split 64 + 64 over the two CTAs at one common offset, so a third 64-row block per stage holds Bhi[0:64]
// (leader, a duplicate) / Bhi[64:128] (peer).  Both accumulator halves restart with every chunk.
// BN = 64 (long-box fused-B kernel only): 64 output columns per tile for the vocoder's 64-channel layers and the
// 32-channel layers grouped two time steps per row -- Ahi x [Bhi|Blo] is then an N = 128 MMA, Alo x Bhi an N = 64 one.
template <int CG, int WIDE = 0, int FUSE = 0, int AR = G2_A_ROWS, int BN = G2_BN>
struct G2Cfg {
  static_assert(!FUSE || (CG == 2 && !WIDE), "the fused-B variant is the CTA-pair, long-reduction kernel");
  static_assert(AR == G2_A_ROWS || FUSE, "the long A box exists for the fused-B kernel only");
  static_assert(BN == G2_BN || (BN == 64 && FUSE && AR != G2_A_ROWS), "narrow tiles exist for the long-box kernel only");
  static constexpr int A_ROWS = AR;
  static constexpr int A_PLANE = AR * 128;
  static constexpr int A_STAGE = 2 * A_PLANE;
  static constexpr int BIAS_MAX = AR == G2_A_ROWS ? G2_BIAS_MAX : G2_BIAS_MAX_LONG;
  static constexpr int B_ROWS = BN / CG;
  static constexpr int B_PLANE = B_ROWS * 128;
  static constexpr int B_STAGE = (FUSE ? 3 : 2) * B_PLANE;
  // the wide (short-reduction) variant trades pipeline depth for 16 per-warp transposition buffers
  static constexpr int A_STAGES = WIDE ? 2 : (CG == 2 ? (FUSE ? 2 : 3) : 2);
  static constexpr int B_STAGES = WIDE ? (CG == 2 ? 4 : 2)
                                       : (CG == 2 ? (FUSE ? (AR == G2_A_ROWS ? 4 : (BN == 64 ? 6 : 3)) : 5) : 3);
  static constexpr int EPI_WARPS = WIDE ? 16 : 8;
  static constexpr int SMEM_TILES = A_STAGES * A_STAGE + B_STAGES * B_STAGE;
  static constexpr int SMEM_BYTES = SMEM_TILES + 1024 + 512 + 4 * BIAS_MAX + EPI_WARPS * G2_STAGE_WARP_BYTES;
  static_assert(SMEM_BYTES <= 232448, "exceeds 227 KB");
};

// Compacts the row tiles that can reach a valid output: utterance b keeps ceil(min(T, L_b + halo) / 128).
__global__ void build_tile_list_kernel(const int* __restrict__ lens, int B, int T, int halo,
                                       int2* __restrict__ list, int* __restrict__ count) {
  __shared__ int warp_tot[32];
  __shared__ int carry_s;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
  if (threadIdx.x == 0) carry_s = 0;
  __syncthreads();
  for (int base = 0; base < B; base += blockDim.x) {
    const int b = base + threadIdx.x;
    int n = 0;
    if (b < B) {
      const int lim = min(T, max(lens[b], 0) + halo);
      n = (lim + G2_BM - 1) / G2_BM;
    }
    int v = n;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int u = __shfl_up_sync(0xffffffffu, v, o);
      if (lane >= o) v += u;
    }
    if (lane == 31) warp_tot[w] = v;
    __syncthreads();
    int pre = carry_s;
    for (int k = 0; k < w; ++k) pre += warp_tot[k];
    const int start = pre + v - n;
    for (int j = 0; j < n; ++j) list[start + j] = make_int2(b, j * G2_BM);
    __syncthreads();
    if (threadIdx.x == 0) {
      int tot = 0;
      for (int k = 0; k < nw; ++k) tot += warp_tot[k];
      carry_s += tot;
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) *count = carry_s;
}

// Transposed store of one warp's 32 rows x 32 columns block staged in shared memory (row stride 144 B):
// lane (sub_row, sub_col) owns four consecutive columns of row 4 * itr + sub_row, so eight lanes cover one row's
// 128 contiguous bytes and every global access is made of whole lines -- scattered one-row-per-thread 16-byte
// accesses cost the LSU two cycles per row and instruction and an L2 request per half sector.  Adds the fp32
// residual (all eight row loads are issued before the first store), zeroes rows past the utterance, writes the
// fp32 result and its fp16 hi/lo operand planes, and range-checks what becomes an operand.
template <bool VOC = true>
__device__ __forceinline__ void g2_store_block32(const GemmParams& p, uint32_t stg, int b, int t_base, int n,
                                                 int lens_b, int check_b, int lane, size_t out_off = 0) {
  const int sub_row = lane >> 3, sub_col = (lane & 7) * 4;
  const int nn = n + sub_col;
  const bool col_ok = nn < p.N;
  float4 res[8];
  if (p.resid != nullptr) {
#pragma unroll
    for (int itr = 0; itr < 8; ++itr) {
      const int tr = t_base + itr * 4 + sub_row;
      res[itr] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
      if (tr < p.T && col_ok)
        res[itr] = *reinterpret_cast<const float4*>(p.resid + (static_cast<size_t>(b) * p.T + tr) * p.ld_out + nn);
    }
  }
#pragma unroll
  for (int itr = 0; itr < 8; ++itr) {
    const int rr = itr * 4 + sub_row;
    const int tr = t_base + rr;
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                 : "r"(stg + rr * G2_STAGE_ROW_BYTES + sub_col * 4) : "memory");
    if (tr < p.T && col_ok) {
      const size_t mr = static_cast<size_t>(b) * p.T + tr;
      if (p.resid != nullptr) {
        v.x = res[itr].x + v.x; v.y = res[itr].y + v.y; v.z = res[itr].z + v.z; v.w = res[itr].w + v.w;
      }
      if (tr >= lens_b) v = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
      if (p.out != nullptr) *reinterpret_cast<float4*>(p.out + out_off + mr * p.ld_out + nn) = v;
      if (p.out_hi != nullptr) {
        if (tr < check_b && outside_fp16_range(v) && p.err_flag != nullptr) atomicOr(p.err_flag, 8 | p.err_code);
        if (VOC && p.plane_act) {                   // operand of the next layer = LeakyReLU(0.1) of the stored value
          v.x = v.x > 0.0f ? v.x : v.x * 0.1f; v.y = v.y > 0.0f ? v.y : v.y * 0.1f;
          v.z = v.z > 0.0f ? v.z : v.z * 0.1f; v.w = v.w > 0.0f ? v.w : v.w * 0.1f;
        }
        const __half2 h01 = __floats2half2_rn(v.x, v.y), h23 = __floats2half2_rn(v.z, v.w);
        const float2 f01 = __half22float2(h01), f23 = __half22float2(h23);
        const __half2 l01 = __floats2half2_rn((v.x - f01.x) * SPLIT_SCALE, (v.y - f01.y) * SPLIT_SCALE);
        const __half2 l23 = __floats2half2_rn((v.z - f23.x) * SPLIT_SCALE, (v.w - f23.y) * SPLIT_SCALE);
        uint2 ph, pl;
        ph.x = *reinterpret_cast<const uint32_t*>(&h01); ph.y = *reinterpret_cast<const uint32_t*>(&h23);
        pl.x = *reinterpret_cast<const uint32_t*>(&l01); pl.y = *reinterpret_cast<const uint32_t*>(&l23);
        *reinterpret_cast<uint2*>(p.out_hi + mr * p.ld_pl + nn) = ph;
        *reinterpret_cast<uint2*>(p.out_lo + mr * p.ld_pl + nn) = pl;
      }
    }
  }
}

// EPI selects how much epilogue is compiled in (the fully unrolled column loop makes every option cost
// instruction-cache footprint in all launches): 0 = bias / activation / residual / fp32 + plane stores (conv and
// Linear layers), 1 = also the divisor and the transposed plane store, 2 = token-softmax partials only, 3 = magnitude
// of (re, im) column pairs as operand planes (the STFT GEMM of the log-mel front-end; unfused epilogue variant only).
enum Gemm2Epi { EPI_STD = 0, EPI_FULL = 1, EPI_SOFTMAX = 2, EPI_MAG = 3 };

// WIDE = 1 is the variant for short reductions (a single accumulation chunk: Linear layers, mel prenet, the
// expansion matmul), whose cost is the epilogue, not the MMAs: sixteen epilogue warps instead of eight (two per
// TMEM lane quadrant and group, 64 columns each) read the finished accumulators straight from tensor memory 16
// columns at a time -- no running sums, 96 registers per thread, twice the warps to hide the store latency.
constexpr int G2_THREADS_WIDE = 128 + 32 * 16;

// SPLIT = 1 compiles the split-reduction bookkeeping in (work item = k-block range of a tile, partial planes).  It is a
// template parameter because the run-time form of it cost the unsplit conv launches 7.5 % more issued instructions and
// 6.7 points of tensor-pipe activity (measured by bisecting the commits on one box: 198 M -> 213 M warp instructions,
// 86.6 % -> 79.9 %).
// BMN = 1: the B operand is read MN-major from activation planes [B, T, C] (TMA boxes of 64 columns x 64 rows): the
// position-reduction GEMMs of the weight gradient, where the reduction index is the row of the activation and a conv tap
// is a row offset of the box -- no transposed copies of the activation (fused-B pair variant only).  BMN = 2: the A
// operand as well (the masked gradient's planes [B, T, C]: M = its channels, 64-channel x 64-row boxes, two per plane).
template <int CG, int EPI, int WIDE, int FUSE = 0, int AR = G2_A_ROWS, int BN = 128, int SPLIT = 0, int BMN = 0>
__global__ void __launch_bounds__(WIDE ? G2_THREADS_WIDE : G2_THREADS, 1)
gemm2_kernel(const __grid_constant__ CUtensorMap tmA_hi, const __grid_constant__ CUtensorMap tmA_lo,
             const __grid_constant__ CUtensorMap tmB_hi, const __grid_constant__ CUtensorMap tmB_lo,
             const GemmParams p) {
  using Cfg = G2Cfg<CG, WIDE, FUSE, AR, BN>;
  constexpr int G2_BN = BN;                          // column tile of this instantiation (shadows the default)
  constexpr int G2_A_PLANE = Cfg::A_PLANE;
  constexpr int G2_A_STAGE = Cfg::A_STAGE;
  // dilated taps and activated operand planes exist for the vocoder's layers only (long A boxes, and its short layers
  // on the wide variant); the acoustic model's conv instantiation compiles them out
  constexpr bool VOC = WIDE || AR != G2_A_ROWS;
  const int dil = (VOC && p.dil > 1) ? p.dil : 1;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sA = smem_base;
  const uint32_t sB = sA + Cfg::A_STAGES * G2_A_STAGE;
  const uint32_t sBar = sB + Cfg::B_STAGES * Cfg::B_STAGE;
  auto fullA = [&](int s) { return sBar + 8u * s; };               // [4]
  auto emptyA = [&](int s) { return sBar + 32u + 8u * s; };        // [4]
  auto fullB = [&](int s) { return sBar + 64u + 8u * s; };         // [8]
  auto emptyB = [&](int s) { return sBar + 128u + 8u * s; };       // [8]
  // acc0_full is per (epilogue group, buffer): a group only ever waits on its own barriers, so it is
  // never more than one phase away from them even though it idles through the other group's tiles.
  auto acc0_full = [&](int grp, int s) { return sBar + 192u + 8u * (grp * 2 + s); };   // [2][2]
  auto acc0_empty = [&](int s) { return sBar + 224u + 8u * s; };   // [2]
  auto acc1_full = [&](int s) { return sBar + 240u + 8u * s; };    // [2] (buffer == owning group)
  auto acc1_empty = [&](int s) { return sBar + 256u + 8u * s; };   // [2]
  const uint32_t tmem_slot = sBar + 272u;
  const uint32_t sBias = sBar + 512u;              // [G2_BIAS_MAX] fp32 copy of the bias (epilogue reads it per row)
  const uint32_t sStage = sBias + 4u * Cfg::BIAS_MAX; // per-warp transposition buffers [warps][32 rows][144 B]

  // through a shuffle: the compiler then knows the role branches below are warp-uniform (and keeps the MMA warp's
  // stage / descriptor arithmetic on the uniform datapath in every instantiation)
  const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  const uint32_t rank = CG == 2 ? ptx::cluster_ctarank() : 0u;
  const bool leader = rank == 0;

  if (threadIdx.x == 0) {
    for (int s = 0; s < Cfg::A_STAGES; ++s) { ptx::mbar_init(fullA(s), 1); ptx::mbar_init(emptyA(s), 1); }
    for (int s = 0; s < Cfg::B_STAGES; ++s) { ptx::mbar_init(fullB(s), 1); ptx::mbar_init(emptyB(s), 1); }
    for (int s = 0; s < 2; ++s) {
      ptx::mbar_init(acc0_full(0, s), 1);
      ptx::mbar_init(acc0_full(1, s), 1);
      ptx::mbar_init(acc1_full(s), 1);
      ptx::mbar_init(acc0_empty(s), (WIDE ? 8 : 4) * CG);   // one arrival per warp of the owning group, per CTA
      ptx::mbar_init(acc1_empty(s), (WIDE ? 8 : 4) * CG);
    }
    ptx::fence_barrier_init();
  }
  if (warp == 1) {
    if (CG == 2) ptx::tmem_alloc_pair(tmem_slot, 512); else ptx::tmem_alloc(tmem_slot, 512);
  }
  if (p.bias != nullptr) {                         // launch_gemm guarantees N <= G2_BIAS_MAX when a bias is given
    for (int i = threadIdx.x; i < p.N; i += blockDim.x)
      asm volatile("st.shared.f32 [%0], %1;" ::"r"(sBias + 4u * i), "f"(__ldg(p.bias + i)) : "memory");
  }
  ptx::tc_fence_before();
  if (CG == 2) ptx::cluster_sync(); else __syncthreads();
  ptx::tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
  // Everything above (barriers, tensor memory, bias = weights) is independent of the previous launch: with
  // programmatic dependent launch it overlaps the predecessor's tail.  From here on activations are touched.
  ptx::pdl_wait();
  ptx::pdl_launch_dependents();

  // ---- work list (identical in every role) ----
  const int n_nt = (p.N + G2_BN - 1) / G2_BN;
  const int tiles_per_b = (p.T + G2_BM - 1) / G2_BM;
  const int n_rt = p.tile_list != nullptr ? *p.tile_count : p.B * tiles_per_b;
  const int n_prt = (n_rt + CG - 1) / CG;
  const int num_kb = (p.K + G2_BK - 1) / G2_BK;
  const int chunk_kb = p.chunk_kb < 1 ? num_kb : p.chunk_kb;
  // split reduction: item w is chunk (w % splits) of tile (w / splits); every item is one accumulation chunk
  static_assert(!SPLIT || FUSE, "split reductions run on the fused-B kernel");
  const int splits = (SPLIT && p.splits > 1) ? p.splits : 1;
  // work items are counted in 32 bits (launch_gemm2_t checks the range): every epilogue thread locates its tile once per
  // item, and 64-bit divisions cost ~100 instructions each
  const unsigned total = static_cast<unsigned>(n_prt) * n_nt * splits;
  const int cid = blockIdx.x / CG, ncl = gridDim.x / CG;
  const int item_kb = (SPLIT && splits > 1) ? (p.split_kb > 0 ? p.split_kb : chunk_kb) : num_kb;   // k-blocks per work item
  auto kb_first = [&](unsigned w) { return (SPLIT && splits > 1) ? static_cast<int>(w % static_cast<unsigned>(splits)) * item_kb : 0; };

  auto locate = [&](unsigned w_in, int& b, int& t0, int& n0, bool& valid) {
    const unsigned w = (SPLIT && splits > 1) ? w_in / static_cast<unsigned>(splits) : w_in;
    const unsigned wq = w / static_cast<unsigned>(n_nt);
    const int nt = static_cast<int>(w - wq * n_nt);
    const int rt = static_cast<int>(wq) * CG + static_cast<int>(rank);
    n0 = nt * G2_BN;
    valid = rt < n_rt;
    if (!valid) { b = p.B; t0 = 0; return; }                 // b == B: every TMA row is out of range -> zeros
    if (p.tile_list != nullptr) {
      const int2 e = p.tile_list[rt];
      b = e.x; t0 = e.y;
    } else {
      b = rt / tiles_per_b; t0 = (rt % tiles_per_b) * G2_BM;
    }
  };

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer (every CTA)
    if (!WIDE) asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(G2_REGS_CTRL));
    if (lane == 0) {
      ptx::prefetch_tensormap(&tmA_hi); ptx::prefetch_tensormap(&tmA_lo);
      ptx::prefetch_tensormap(&tmB_hi); ptx::prefetch_tensormap(&tmB_lo);
      uint32_t ia = 0, ib = 0;
      for (unsigned w = cid; w < total; w += ncl) {
        int b, t0, n0; bool valid;
        locate(w, b, t0, n0, valid);
        const int kbf = kb_first(w);
        for (int kb = kbf; kb < min(kbf + item_kb, num_kb); ++kb) {
          {
            const int s = ia % Cfg::A_STAGES; const uint32_t ph = (ia / Cfg::A_STAGES) & 1;
            ptx::mbar_wait(emptyA(s), ph ^ 1u);
            const uint32_t dst = sA + s * G2_A_STAGE;
            if (CG == 2 && BMN == 2) {
              // MN-major A: channels t0 .. t0 + 127 (this CTA's M rows) of rows tk .. tk + 63 of utterance bb, two
              // 64-column atoms per plane; an invalid tile (b == p.B) reads utterance p.bmn_batches: zeros
              const uint32_t bar = ptx::map_to_cta(fullA(s), 0);
              if (leader) ptx::mbar_expect_tx(fullA(s), 2 * 4 * 64 * 128);
              const int bb = valid ? kb / p.bmn_per : p.bmn_batches, tk = (kb % p.bmn_per) * G2_BK;
              ptx::tma_load_3d_pair(&tmA_hi, bar, dst, t0, tk, bb);
              ptx::tma_load_3d_pair(&tmA_hi, bar, dst + 64 * 128, t0 + 64, tk, bb);
              ptx::tma_load_3d_pair(&tmA_lo, bar, dst + G2_A_PLANE, t0, tk, bb);
              ptx::tma_load_3d_pair(&tmA_lo, bar, dst + G2_A_PLANE + 64 * 128, t0 + 64, tk, bb);
            } else if (CG == 2) {
              const uint32_t bar = ptx::map_to_cta(fullA(s), 0);
              if (leader) ptx::mbar_expect_tx(fullA(s), 2 * G2_A_STAGE);
              ptx::tma_load_3d_pair(&tmA_hi, bar, dst, kb * G2_BK, t0 - p.pad * dil, b);
              ptx::tma_load_3d_pair(&tmA_lo, bar, dst + G2_A_PLANE, kb * G2_BK, t0 - p.pad * dil, b);
            } else {
              ptx::mbar_expect_tx(fullA(s), G2_A_STAGE);
              ptx::tma_load_3d(&tmA_hi, fullA(s), dst, kb * G2_BK, t0 - p.pad * dil, b);
              ptx::tma_load_3d(&tmA_lo, fullA(s), dst + G2_A_PLANE, kb * G2_BK, t0 - p.pad * dil, b);
            }
            ++ia;
          }
          for (int tap = 0; tap < p.ntaps; ++tap) {
            const int s = ib % Cfg::B_STAGES; const uint32_t ph = (ib / Cfg::B_STAGES) & 1;
            ptx::mbar_wait(emptyB(s), ph ^ 1u);
            const uint32_t dst = sB + s * Cfg::B_STAGE;
            const int z = p.b_batched ? b : tap;
            const int nrow = n0 + static_cast<int>(rank) * Cfg::B_ROWS;
            if (FUSE && BMN) {
              // the same three blocks per CTA, each a 64-column x 64-row box of the activation planes (an MN-major atom
              // column): columns n0 .. n0 + 127 of rows tk .. tk + 63 of utterance bb
              const uint32_t bar = ptx::map_to_cta(fullB(s), 0);
              if (leader) ptx::mbar_expect_tx(fullB(s), 2 * Cfg::B_STAGE);
              const CUtensorMap* own = leader ? &tmB_hi : &tmB_lo;
              const int bb = kb / p.bmn_per, tk = (kb - bb * p.bmn_per) * G2_BK + p.b_koff;
              ptx::tma_load_3d_pair(own, bar, dst, n0, tk, bb);
              ptx::tma_load_3d_pair(own, bar, dst + Cfg::B_PLANE, n0 + Cfg::B_ROWS, tk, bb);
              ptx::tma_load_3d_pair(&tmB_hi, bar, dst + 2 * Cfg::B_PLANE, nrow, tk, bb);
            } else if (FUSE) {
              // leader: Bhi[0:64], Bhi[64:128], Bhi[0:64] again; peer: Blo[0:64], Blo[64:128], Bhi[64:128]
              const uint32_t bar = ptx::map_to_cta(fullB(s), 0);
              if (leader) ptx::mbar_expect_tx(fullB(s), 2 * Cfg::B_STAGE);
              const CUtensorMap* own = leader ? &tmB_hi : &tmB_lo;
              ptx::tma_load_3d_pair(own, bar, dst, kb * G2_BK + p.b_koff, n0, z);
              ptx::tma_load_3d_pair(own, bar, dst + Cfg::B_PLANE, kb * G2_BK + p.b_koff, n0 + Cfg::B_ROWS, z);
              ptx::tma_load_3d_pair(&tmB_hi, bar, dst + 2 * Cfg::B_PLANE, kb * G2_BK + p.b_koff, nrow, z);
            } else if (CG == 2) {
              const uint32_t bar = ptx::map_to_cta(fullB(s), 0);
              if (leader) ptx::mbar_expect_tx(fullB(s), 2 * Cfg::B_STAGE);
              ptx::tma_load_3d_pair(&tmB_hi, bar, dst, kb * G2_BK + p.b_koff, nrow, z);
              ptx::tma_load_3d_pair(&tmB_lo, bar, dst + Cfg::B_PLANE, kb * G2_BK + p.b_koff, nrow, z);
            } else {
              ptx::mbar_expect_tx(fullB(s), Cfg::B_STAGE);
              ptx::tma_load_3d(&tmB_hi, fullB(s), dst, kb * G2_BK + p.b_koff, nrow, z);
              ptx::tma_load_3d(&tmB_lo, fullB(s), dst + Cfg::B_PLANE, kb * G2_BK + p.b_koff, nrow, z);
            }
            ++ib;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer (leader CTA)
    // The WHOLE warp walks the schedule and lane 0 issues.  With a lane-dependent branch around the loop the compiler
    // keeps the stage / descriptor arithmetic in per-thread registers and moves every operand of every tcgen05.mma into
    // uniform registers through an elect + R2UR.BROADCAST sequence right before the instruction -- five dependent
    // moves on the issue path of each MMA, 6.7 points of tensor-pipe activity on the conv layers (found by bisecting
    // the round-1 commits on one box).  Warp-uniform control flow keeps that arithmetic on the uniform datapath.
    if (!WIDE) asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(G2_REGS_CTRL));
    if (leader) {
      const bool issuer = lane == 0;
      constexpr uint32_t idesc = ptx::make_idesc_f16(G2_BM * CG, G2_BN, BMN != 0, BMN == 2);
      constexpr uint32_t idesc_wide = ptx::make_idesc_f16(G2_BM * CG, 2 * G2_BN, BMN != 0, BMN == 2);
      auto commit = [&](uint32_t bar) {
        if (issuer) {
          if (CG == 2) ptx::tc_commit_pair(bar, 3); else ptx::tc_commit(bar);
        }
      };
      uint32_t ia = 0, ib = 0, g = 0, it = 0;
      for (unsigned w = cid; w < total; w += ncl) {
        const uint32_t tb = it & 1u;
        if (!FUSE) ptx::mbar_wait(acc1_empty(tb), ((it >> 1) & 1u) ^ 1u);
        const uint32_t acc1 = tmem_base + 256u + tb * G2_BN;
        uint32_t first1 = 1;
        const int kbf = kb_first(w), kbl = min(kbf + item_kb, num_kb);
        for (int kb0 = kbf; kb0 < kbl; kb0 += chunk_kb) {
          const uint32_t buf = g & 1u;
          ptx::mbar_wait(acc0_empty(buf), ((g >> 1) & 1u) ^ 1u);
          ptx::tc_fence_after();
          const uint32_t acc0 = tmem_base + buf * (FUSE ? 2 * G2_BN : G2_BN);
          uint32_t first0 = 1;
          const int kb1 = min(kb0 + chunk_kb, kbl);
          for (int kb = kb0; kb < kb1; ++kb) {
            const int sa = ia % Cfg::A_STAGES;
            ptx::mbar_wait(fullA(sa), (ia / Cfg::A_STAGES) & 1);
            ++ia;
            for (int tap = 0; tap < p.ntaps; ++tap) {
              const int sb = ib % Cfg::B_STAGES;
              ptx::mbar_wait(fullB(sb), (ib / Cfg::B_STAGES) & 1);
              ++ib;
              ptx::tc_fence_after();
              const uint32_t a_addr = sA + sa * G2_A_STAGE + tap * dil * 128;     // row shift = tap * dilation
              const uint32_t b_addr = sB + sb * Cfg::B_STAGE;
              const uint64_t dAh = BMN == 2 ? ptx::make_desc_sw128_mn(a_addr, 64 * 128) : ptx::make_desc_sw128(a_addr, 0);
              const uint64_t dAl = BMN == 2 ? ptx::make_desc_sw128_mn(a_addr + G2_A_PLANE, 64 * 128)
                                            : ptx::make_desc_sw128(a_addr + G2_A_PLANE, 0);
              const uint64_t dBh = BMN ? ptx::make_desc_sw128_mn(b_addr, Cfg::B_PLANE) : ptx::make_desc_sw128(b_addr, 0);
              const uint64_t dBl = ptx::make_desc_sw128(b_addr + Cfg::B_PLANE, 0);
              const uint64_t dB3 = BMN ? ptx::make_desc_sw128_mn(b_addr + 2 * Cfg::B_PLANE, Cfg::B_PLANE)
                                       : ptx::make_desc_sw128(b_addr + 2 * Cfg::B_PLANE, 0);
              if (issuer) {
#pragma unroll
                for (int k = 0; k < G2_BK / 16; ++k) {
                  const uint64_t ko = static_cast<uint64_t>(k * 2);
                  const uint64_t kob = BMN ? static_cast<uint64_t>(k * (16 * 128 / 16)) : ko;   // MN-major: 16 rows on
                  const uint64_t koa = BMN == 2 ? kob : ko;
                  if (FUSE) {
                    // columns [0,128): Ahi*Bhi, [128,256): Ahi*Blo + Alo*Bhi (dBl + one plane = the third block)
                    ptx::mma_f16_ss_pair(acc0, dAh + koa, dBh + kob, idesc_wide, (first0 && k == 0) ? 0u : 1u);
                    ptx::mma_f16_ss_pair(acc0 + G2_BN, dAl + koa, dB3 + kob, idesc, 1u);
                  } else if (CG == 2) {
                    ptx::mma_f16_ss_pair(acc0, dAh + ko, dBh + ko, idesc, (first0 && k == 0) ? 0u : 1u);
                    ptx::mma_f16_ss_pair(acc1, dAh + ko, dBl + ko, idesc, (first1 && k == 0) ? 0u : 1u);
                    ptx::mma_f16_ss_pair(acc1, dAl + ko, dBh + ko, idesc, 1u);
                  } else {
                    ptx::mma_f16_ss(acc0, dAh + ko, dBh + ko, idesc, (first0 && k == 0) ? 0u : 1u);
                    ptx::mma_f16_ss(acc1, dAh + ko, dBl + ko, idesc, (first1 && k == 0) ? 0u : 1u);
                    ptx::mma_f16_ss(acc1, dAl + ko, dBh + ko, idesc, 1u);
                  }
                }
              }
              first0 = 0; first1 = 0;
              __syncwarp();
              commit(emptyB(sb));
            }
            commit(emptyA(sa));
          }
          commit(acc0_full(static_cast<int>(it & 1u), static_cast<int>(buf)));
          ++g;
        }
        if (!FUSE) commit(acc1_full(tb));
        ++it;
      }
    }
  } else if (warp < 4) {
    if (!WIDE) asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(G2_REGS_CTRL));   // idle warps of warpgroup 0
  } else if (WIDE) {
    // ------------------------------------------------------------ direct epilogue (warps 4..19), one chunk per tile
    // Scattered 16-byte stores (one row per thread) cost the LSU two cycles per row and instruction and are
    // what bounds these launches, so each warp transposes its 32 x 32 block through shared memory: eight lanes
    // then write one row's 128 contiguous bytes, four rows per instruction.
    const int q = warp & 3;
    const uint32_t ew = static_cast<uint32_t>(warp - 4);
    const uint32_t grp = ew >> 3;
    const int cbase = static_cast<int>((ew >> 2) & 1u) * (G2_BN / 2);
    const int row = q * 32 + lane;
    const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    const uint32_t stg = sStage + ew * G2_STAGE_WARP_BYTES;
    auto release = [&](uint32_t bar) {
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (CG == 2) ptx::mbar_arrive_cluster(ptx::map_to_cta(bar, 0)); else ptx::mbar_arrive(bar);
      }
    };
    uint32_t g = 0, it = 0, uses0 = 0u, uses1 = 0u;
    for (unsigned w = cid; w < total; w += ncl, ++it, ++g) {       // exactly one chunk per tile
      if ((it & 1u) != grp) continue;
      int b, t0, n0; bool valid;
      locate(w, b, t0, n0, valid);
      const int t = t0 + row;
      const bool tile_live = valid && !(p.skip_lens != nullptr && t0 >= p.skip_lens[b] + p.skip_halo);
      const bool row_ok = tile_live && t < p.T;
      const bool row_live = row_ok && (p.lens == nullptr || t < p.lens[b]);
      const int lens_b = (valid && p.lens != nullptr) ? p.lens[b] : p.T;
      const int check_b = (valid && p.skip_lens != nullptr) ? p.skip_lens[b] + p.skip_halo : p.T;
      const uint32_t buf = g & 1u, tb = it & 1u;
      ptx::mbar_wait(acc0_full(static_cast<int>(grp), static_cast<int>(buf)), (buf ? uses1 : uses0) & 1u);
      if (buf) ++uses1; else ++uses0;
      ptx::mbar_wait(acc1_full(tb), (it >> 1) & 1u);
      ptx::tc_fence_after();
      if (EPI == EPI_SOFTMAX) {
        // token softmax of this warp's 64 columns, 16 at a time with a running maximum (online softmax); the two
        // sums are kept in double.  One partial per (row, column half): softmax_part[row][2 * tile + half].
        const int L = valid ? p.col_lens[b] : 0;
        float mx = -CUDART_INF_F;
        double den = 0.0, num = 0.0;
#pragma unroll
        for (int c16 = 0; c16 < G2_BN / 32; ++c16) {
          uint32_t r0[16], r1[16];
          __syncwarp();
          ptx::tmem_ld_32x16(lane_addr + buf * G2_BN + cbase + c16 * 16, r0);
          ptx::tmem_ld_32x16(lane_addr + 256u + tb * G2_BN + cbase + c16 * 16, r1);
          ptx::tmem_ld_wait();
          if (c16 == G2_BN / 32 - 1) {
            release(acc0_empty(buf));
            release(acc1_empty(tb));
          }
          const int nb = n0 + cbase + c16 * 16;
          float sv[16];
          float cm = -CUDART_INF_F;
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            sv[j] = __fdiv_rn(__fadd_rn(__uint_as_float(r0[j]), __uint_as_float(r1[j]) * SPLIT_INV_SCALE), p.divisor);
            if (nb + j < L) cm = fmaxf(cm, sv[j]);
          }
          if (cm > mx) {
            const double sc = static_cast<double>(expf(mx - cm));      // exp(-inf) = 0 on the first block
            den *= sc; num *= sc;
            mx = cm;
          }
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            if (nb + j < L) {
              const double ev = static_cast<double>(expf(sv[j] - mx));
              den += ev;
              num = fma(ev, static_cast<double>(nb + j), num);
            }
          }
        }
        if (row_ok)
          p.softmax_part[(static_cast<size_t>(b) * p.T + t) * (2 * n_nt) + 2 * (n0 / G2_BN) + (cbase ? 1 : 0)] =
              make_float4(mx, static_cast<float>(den), static_cast<float>(num), 0.0f);
        continue;
      }
#pragma unroll
      for (int c32 = 0; c32 < G2_BN / 64; ++c32) {
        const int n = n0 + cbase + c32 * 32;
        __syncwarp();                               // previous block's transposed reads are done; ld is .aligned
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          uint32_t r0[16], r1[16];
          ptx::tmem_ld_32x16(lane_addr + buf * G2_BN + cbase + c32 * 32 + h * 16, r0);
          ptx::tmem_ld_32x16(lane_addr + 256u + tb * G2_BN + cbase + c32 * 32 + h * 16, r1);
          ptx::tmem_ld_wait();
          if (c32 == G2_BN / 64 - 1 && h == 1) {    // everything this warp needs has left tensor memory
            release(acc0_empty(buf));
            release(acc1_empty(tb));
          }
#pragma unroll
          for (int k4 = 0; k4 < 4; ++k4) {
            const int nn = n + h * 16 + k4 * 4;
            float vv[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              vv[j] = __fadd_rn(__uint_as_float(r0[k4 * 4 + j]), __uint_as_float(r1[k4 * 4 + j]) * SPLIT_INV_SCALE);
              if (EPI == EPI_FULL && p.divisor != 1.0f) vv[j] = __fdiv_rn(vv[j], p.divisor);
            }
            if (p.bias != nullptr && nn < p.N) {
              float4 bb;
              asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(bb.x), "=f"(bb.y), "=f"(bb.z), "=f"(bb.w)
                           : "r"(sBias + 4u * nn));
              vv[0] += bb.x; vv[1] += bb.y; vv[2] += bb.z; vv[3] += bb.w;
            }
            if (p.act == ACT_LRELU) {
#pragma unroll
              for (int j = 0; j < 4; ++j) vv[j] = vv[j] > 0.0f ? vv[j] : vv[j] * 0.1f;
            } else if (p.act == ACT_RELU) {
#pragma unroll
              for (int j = 0; j < 4; ++j) vv[j] = fmaxf(vv[j], 0.0f);
            } else if (p.act == ACT_LOGCLAMP) {
#pragma unroll
              for (int j = 0; j < 4; ++j) vv[j] = logf(fmaxf(vv[j], 1e-5f));
            }
            if (EPI == EPI_FULL && p.outT_hi != nullptr && row_ok && nn < p.N) {   // t-contiguous planes: thread = row
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const float x = row_live ? vv[j] : 0.0f;
                const size_t o = (static_cast<size_t>(b) * p.N + (nn + j)) * p.ld_t + t;
                const __half hh = __float2half_rn(x);
                p.outT_hi[o] = hh;
                p.outT_lo[o] = __float2half_rn((x - __half2float(hh)) * SPLIT_SCALE);
              }
            }
            asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(stg + lane * G2_STAGE_ROW_BYTES + h * 64 + k4 * 16),
                         "f"(vv[0]), "f"(vv[1]), "f"(vv[2]), "f"(vv[3]) : "memory");
          }
        }
        __syncwarp();
        if (!tile_live) continue;                   // warp-uniform
        g2_store_block32<VOC>(p, stg, b, t0 + q * 32, n, lens_b, check_b, lane);
      }
    }
  } else {
    // ------------------------------------------------------------ accumulate + epilogue (warps 4..11)
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(G2_REGS_EPI));
    // Two groups of four warps (one warp per TMEM lane quadrant, warp w may touch lanes (w % 4) * 32 ..)
    // take alternate tiles: while one group runs the store-heavy epilogue of tile i, the other drains
    // the chunks of tile i + 1 as soon as the MMAs commit them, so the tensor pipe never waits for stores.
    const int q = warp & 3;
    const uint32_t grp = static_cast<uint32_t>(warp - 4) >> 2;
    const int row = q * 32 + lane;
    const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    // accumulation chunks of work item w (the last split of a reduction may be shorter than the others)
    const uint32_t nchunks_all = static_cast<uint32_t>((num_kb + chunk_kb - 1) / chunk_kb);
    auto chunks_of = [&](unsigned w) {
      if (!SPLIT) return nchunks_all;
      const int kbf = kb_first(w);
      return static_cast<uint32_t>((min(kbf + item_kb, num_kb) - kbf + chunk_kb - 1) / chunk_kb);
    };
    auto release = [&](uint32_t bar) {            // one arrival per warp, on the leader's barrier
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (CG == 2) ptx::mbar_arrive_cluster(ptx::map_to_cta(bar, 0)); else ptx::mbar_arrive(bar);
      }
    };
    uint32_t g = 0, it = 0;
    uint32_t uses0 = 0u, uses1 = 0u;              // completed waits on this group's acc0_full[buf]
    for (unsigned w = cid; w < total; w += ncl, ++it) {
      const uint32_t nchunks = chunks_of(w);
      if ((it & 1u) != grp) { g += nchunks; continue; }
      int b, t0, n0; bool valid;
      locate(w, b, t0, n0, valid);
      const int t = t0 + row;
      const bool tile_live = valid && !(p.skip_lens != nullptr && t0 >= p.skip_lens[b] + p.skip_halo);
      const bool row_ok = tile_live && t < p.T;
      const bool row_live = row_ok && (p.lens == nullptr || t < p.lens[b]);
      // rows past L_b + halo of a live tile may be fed by never-written workspace: they cannot reach a valid
      // output, so their values are not range-checked
      const bool row_checked = row_ok && (p.skip_lens == nullptr || t < p.skip_lens[b] + p.skip_halo);
      const size_t m = row_ok ? static_cast<size_t>(b) * p.T + t : 0;
      float sum[G2_BN];
#pragma unroll
      for (int j = 0; j < G2_BN; ++j) sum[j] = 0.0f;
      for (uint32_t ch = 0; ch < nchunks; ++ch, ++g) {
        const uint32_t buf = g & 1u;
        ptx::mbar_wait(acc0_full(static_cast<int>(grp), static_cast<int>(buf)), (buf ? uses1 : uses0) & 1u);
        if (buf) ++uses1; else ++uses0;
        ptx::tc_fence_after();
        __syncwarp();
        if (FUSE) {
          // both halves of the chunk: main product and the 2^-11 correction, folded with one rounding
#pragma unroll
          for (int c = 0; c < G2_BN / 16; ++c) {
            uint32_t r0[16], r1[16];
            ptx::tmem_ld_32x16(lane_addr + buf * (2 * G2_BN) + c * 16, r0);
            ptx::tmem_ld_32x16(lane_addr + buf * (2 * G2_BN) + G2_BN + c * 16, r1);
            ptx::tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 16; ++j)
              sum[c * 16 + j] = __fadd_rn(sum[c * 16 + j],
                                          __fmaf_rn(__uint_as_float(r1[j]), SPLIT_INV_SCALE, __uint_as_float(r0[j])));
          }
        } else {
#pragma unroll
          for (int c = 0; c < G2_BN / 32; ++c) {     // two TMEM loads in flight per wait
            uint32_t r0[16], r1[16];
            ptx::tmem_ld_32x16(lane_addr + buf * G2_BN + c * 32, r0);
            ptx::tmem_ld_32x16(lane_addr + buf * G2_BN + c * 32 + 16, r1);
            ptx::tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              sum[c * 32 + j] = __fadd_rn(sum[c * 32 + j], __uint_as_float(r0[j]));
              sum[c * 32 + 16 + j] = __fadd_rn(sum[c * 32 + 16 + j], __uint_as_float(r1[j]));
            }
          }
        }
        release(acc0_empty(buf));
      }
      if (!FUSE) {
        // fold the correction accumulator in and hand its TMEM back before the stores start
        const uint32_t tb = it & 1u;
        ptx::mbar_wait(acc1_full(tb), (it >> 1) & 1u);
        ptx::tc_fence_after();
        __syncwarp();
#pragma unroll
        for (int c = 0; c < G2_BN / 32; ++c) {
          uint32_t r0[16], r1[16];
          ptx::tmem_ld_32x16(lane_addr + 256u + tb * G2_BN + c * 32, r0);
          ptx::tmem_ld_32x16(lane_addr + 256u + tb * G2_BN + c * 32 + 16, r1);
          ptx::tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            sum[c * 32 + j] = __fadd_rn(sum[c * 32 + j], __uint_as_float(r0[j]) * SPLIT_INV_SCALE);
            sum[c * 32 + 16 + j] = __fadd_rn(sum[c * 32 + 16 + j], __uint_as_float(r1[j]) * SPLIT_INV_SCALE);
          }
        }
        release(acc1_empty(tb));
      }
      if (!tile_live) continue;                     // warp-uniform: the stores below are warp-cooperative
      if (EPI == EPI_SOFTMAX) {
        if (!row_ok) continue;
        // scaled-dot-product softmax over tokens, one partial per column tile (models/efficient_tts.py:390-398):
        // the scores never leave the SM; imv_scan_kernel merges the partials into the position expectation.
        const int L = p.col_lens[b];
        float mx = -CUDART_INF_F;
#pragma unroll
        for (int j = 0; j < G2_BN; ++j) {
          sum[j] = __fdiv_rn(sum[j], p.divisor);
          if (n0 + j < L) mx = fmaxf(mx, sum[j]);
        }
        // 128 sequential fp32 additions would cost ~1e-4 of absolute error on the position expectation (values up
        // to T1); the two running sums are kept in double and rounded once
        double den = 0.0, num = 0.0;
#pragma unroll
        for (int j = 0; j < G2_BN; ++j) {
          if (n0 + j < L) {
            const double ev = static_cast<double>(expf(sum[j] - mx));
            den += ev;
            num = fma(ev, static_cast<double>(n0 + j), num);
          }
        }
        p.softmax_part[m * n_nt + n0 / G2_BN] = make_float4(mx, static_cast<float>(den), static_cast<float>(num), 0.0f);
        continue;
      }
      if (EPI == EPI_SOFTMAX) continue;
      if (EPI == EPI_MAG) {
        // STFT columns (2f, 2f + 1) = (re_f, im_f): sqrt(re^2 + im^2 + 1e-9) evaluated like the reference's
        // spec.pow(2).sum(-1) + 1e-9 (datasets/meldataset.py:72), written as operand planes [B, T, N / 2]; frames past an
        // utterance's length are zero.  32 magnitudes at a time through the transposition buffer, like the fp32 stores.
        const uint32_t stgm = sStage + static_cast<uint32_t>(warp - 4) * G2_STAGE_WARP_BYTES;
        const int lens_m = p.lens != nullptr ? p.lens[b] : p.T;
        const int NM = p.N >> 1;
        const int sub_row = lane >> 3, sub_col = (lane & 7) * 4;
#pragma unroll
        for (int c32 = 0; c32 < G2_BN / 64; ++c32) {
          const int nm = (n0 >> 1) + c32 * 32;
          if (nm >= NM) break;                      // warp-uniform
          __syncwarp();
#pragma unroll
          for (int k4 = 0; k4 < 8; ++k4) {
            float vv[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const float re = sum[2 * (c32 * 32 + k4 * 4 + j)], im = sum[2 * (c32 * 32 + k4 * 4 + j) + 1];
              vv[j] = row_live ? sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(re, re), __fmul_rn(im, im)), 1e-9f)) : 0.0f;
            }
            asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(stgm + lane * G2_STAGE_ROW_BYTES + k4 * 16),
                         "f"(vv[0]), "f"(vv[1]), "f"(vv[2]), "f"(vv[3]) : "memory");
          }
          __syncwarp();
          const int nn = nm + sub_col;
          if (nn < NM) {
#pragma unroll
            for (int itr = 0; itr < 8; ++itr) {
              const int rr = itr * 4 + sub_row;
              const int tr = t0 + q * 32 + rr;
              float4 v;
              asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                           : "r"(stgm + rr * G2_STAGE_ROW_BYTES + sub_col * 4) : "memory");
              if (tr < p.T) {
                if (tr < lens_m && outside_fp16_range(v) && p.err_flag != nullptr) atomicOr(p.err_flag, 8 | p.err_code);
                const __half2 h01 = __floats2half2_rn(v.x, v.y), h23 = __floats2half2_rn(v.z, v.w);
                const float2 f01 = __half22float2(h01), f23 = __half22float2(h23);
                const __half2 l01 = __floats2half2_rn((v.x - f01.x) * SPLIT_SCALE, (v.y - f01.y) * SPLIT_SCALE);
                const __half2 l23 = __floats2half2_rn((v.z - f23.x) * SPLIT_SCALE, (v.w - f23.y) * SPLIT_SCALE);
                uint2 ph, pl;
                ph.x = *reinterpret_cast<const uint32_t*>(&h01); ph.y = *reinterpret_cast<const uint32_t*>(&h23);
                pl.x = *reinterpret_cast<const uint32_t*>(&l01); pl.y = *reinterpret_cast<const uint32_t*>(&l23);
                const size_t o = (static_cast<size_t>(b) * p.T + tr) * p.ld_pl + nn;
                *reinterpret_cast<uint2*>(p.out_hi + o) = ph;
                *reinterpret_cast<uint2*>(p.out_lo + o) = pl;
              }
            }
          }
        }
        continue;
      }
      // stores: 32 columns at a time through this warp's transposition buffer (see g2_store_block32)
      const uint32_t stg = sStage + static_cast<uint32_t>(warp - 4) * G2_STAGE_WARP_BYTES;
      const int lens_b = p.lens != nullptr ? p.lens[b] : p.T;
      const int check_b = p.skip_lens != nullptr ? p.skip_lens[b] + p.skip_halo : p.T;
#pragma unroll
      for (int c32 = 0; c32 < G2_BN / 32; ++c32) {
        const int n = n0 + c32 * 32;
        if (n >= p.N) break;                        // warp-uniform
        __syncwarp();                               // the previous block's transposed reads are done
#pragma unroll
        for (int k4 = 0; k4 < 8; ++k4) {
          const int nn = n + k4 * 4;
          float vv[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            vv[j] = sum[c32 * 32 + k4 * 4 + j];
            if (EPI == EPI_FULL && p.divisor != 1.0f) vv[j] = __fdiv_rn(vv[j], p.divisor);
          }
          if (p.bias != nullptr && nn < p.N) {
            float4 bb;
            asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(bb.x), "=f"(bb.y), "=f"(bb.z), "=f"(bb.w)
                         : "r"(sBias + 4u * nn));
            vv[0] += bb.x; vv[1] += bb.y; vv[2] += bb.z; vv[3] += bb.w;
          }
          if (p.act == ACT_LRELU) {
#pragma unroll
            for (int j = 0; j < 4; ++j) vv[j] = vv[j] > 0.0f ? vv[j] : vv[j] * 0.1f;
          } else if (p.act == ACT_RELU) {
#pragma unroll
            for (int j = 0; j < 4; ++j) vv[j] = fmaxf(vv[j], 0.0f);
          }                                           // ACT_LOGCLAMP exists in the wide variant only (launch_gemm checks)
          if (EPI == EPI_FULL && p.outT_hi != nullptr && row_ok && nn < p.N) {   // t-contiguous planes: thread = row
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const float x = row_live ? vv[j] : 0.0f;
              const size_t o = (static_cast<size_t>(b) * p.N + (nn + j)) * p.ld_t + t;
              const __half hh = __float2half_rn(x);
              p.outT_hi[o] = hh;
              p.outT_lo[o] = __float2half_rn((x - __half2float(hh)) * SPLIT_SCALE);
            }
          }
          asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(stg + lane * G2_STAGE_ROW_BYTES + k4 * 16),
                       "f"(vv[0]), "f"(vv[1]), "f"(vv[2]), "f"(vv[3]) : "memory");
        }
        __syncwarp();
        g2_store_block32<VOC>(p, stg, b, t0 + q * 32, n, lens_b, check_b, lane,
                         (SPLIT && splits > 1) ? static_cast<size_t>(w % static_cast<unsigned>(splits)) * p.split_stride : 0);
      }
    }
  }

  // ---- teardown: nobody may leave while the pair can still touch this CTA's smem / TMEM ----
  __syncwarp();
  ptx::tc_fence_before();
  if (CG == 2) ptx::cluster_sync(); else __syncthreads();
  if (warp == 1) {
    if (CG == 2) ptx::tmem_dealloc_pair(tmem_base, 512); else ptx::tmem_dealloc(tmem_base, 512);
  }
}

// Second half of a split reduction: sums the `splits` partial planes in chunk order (the same fp32 additions, in
// the same order, as the running sums of the unsplit kernel: results are bitwise equal to it) and applies the
// epilogue -- bias, activation, residual, row mask, fp32 store, fp16 hi/lo operand planes, operand range check.
// One thread per (row, 4 columns).  Launches without tile lists only (every row of [B*T, N] is computed).
__global__ void splitk_reduce_kernel(const GemmParams p, const float* __restrict__ part, int splits, size_t split_stride) {
  ptx::pdl_wait();
  ptx::pdl_launch_dependents();
  const size_t n4 = static_cast<size_t>(p.N) >> 2;
  const size_t idx = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
  const size_t rows = static_cast<size_t>(p.B) * p.T;
  if (idx >= rows * n4) return;
  const size_t mr = idx / n4;
  const int nn = static_cast<int>(idx % n4) * 4;
  float4 v = *reinterpret_cast<const float4*>(part + mr * p.N + nn);
  v.x = __fadd_rn(0.0f, v.x); v.y = __fadd_rn(0.0f, v.y); v.z = __fadd_rn(0.0f, v.z); v.w = __fadd_rn(0.0f, v.w);
  for (int s = 1; s < splits; ++s) {
    const float4 u = *reinterpret_cast<const float4*>(part + s * split_stride + mr * p.N + nn);
    v.x = __fadd_rn(v.x, u.x); v.y = __fadd_rn(v.y, u.y); v.z = __fadd_rn(v.z, u.z); v.w = __fadd_rn(v.w, u.w);
  }
  if (p.bias != nullptr) {
    const float4 bb = __ldg(reinterpret_cast<const float4*>(p.bias + nn));
    v.x = __fadd_rn(v.x, bb.x); v.y = __fadd_rn(v.y, bb.y); v.z = __fadd_rn(v.z, bb.z); v.w = __fadd_rn(v.w, bb.w);
  }
  if (p.act == ACT_LRELU) {
    v.x = v.x > 0.0f ? v.x : __fmul_rn(v.x, 0.1f); v.y = v.y > 0.0f ? v.y : __fmul_rn(v.y, 0.1f);
    v.z = v.z > 0.0f ? v.z : __fmul_rn(v.z, 0.1f); v.w = v.w > 0.0f ? v.w : __fmul_rn(v.w, 0.1f);
  } else if (p.act == ACT_RELU) {
    v.x = fmaxf(v.x, 0.0f); v.y = fmaxf(v.y, 0.0f); v.z = fmaxf(v.z, 0.0f); v.w = fmaxf(v.w, 0.0f);
  } else if (p.act == ACT_LOGCLAMP) {
    v.x = logf(fmaxf(v.x, 1e-5f)); v.y = logf(fmaxf(v.y, 1e-5f)); v.z = logf(fmaxf(v.z, 1e-5f)); v.w = logf(fmaxf(v.w, 1e-5f));
  }
  if (p.resid != nullptr) {
    const float4 r = *reinterpret_cast<const float4*>(p.resid + mr * p.ld_out + nn);
    v.x = __fadd_rn(r.x, v.x); v.y = __fadd_rn(r.y, v.y); v.z = __fadd_rn(r.z, v.z); v.w = __fadd_rn(r.w, v.w);
  }
  const int b = static_cast<int>(mr / p.T), tr = static_cast<int>(mr % p.T);
  if (p.lens != nullptr && tr >= p.lens[b]) v = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
  if (p.out != nullptr) *reinterpret_cast<float4*>(p.out + mr * p.ld_out + nn) = v;
  if (p.out_hi != nullptr) {
    if (outside_fp16_range(v) && p.err_flag != nullptr) atomicOr(p.err_flag, 8 | p.err_code);
    if (p.plane_act) {
      v.x = v.x > 0.0f ? v.x : __fmul_rn(v.x, 0.1f); v.y = v.y > 0.0f ? v.y : __fmul_rn(v.y, 0.1f);
      v.z = v.z > 0.0f ? v.z : __fmul_rn(v.z, 0.1f); v.w = v.w > 0.0f ? v.w : __fmul_rn(v.w, 0.1f);
    }
    const __half2 h01 = __floats2half2_rn(v.x, v.y), h23 = __floats2half2_rn(v.z, v.w);
    const float2 f01 = __half22float2(h01), f23 = __half22float2(h23);
    const __half2 l01 = __floats2half2_rn((v.x - f01.x) * SPLIT_SCALE, (v.y - f01.y) * SPLIT_SCALE);
    const __half2 l23 = __floats2half2_rn((v.z - f23.x) * SPLIT_SCALE, (v.w - f23.y) * SPLIT_SCALE);
    uint2 ph, pl;
    ph.x = *reinterpret_cast<const uint32_t*>(&h01); ph.y = *reinterpret_cast<const uint32_t*>(&h23);
    pl.x = *reinterpret_cast<const uint32_t*>(&l01); pl.y = *reinterpret_cast<const uint32_t*>(&l23);
    *reinterpret_cast<uint2*>(p.out_hi + mr * p.ld_pl + nn) = ph;
    *reinterpret_cast<uint2*>(p.out_lo + mr * p.ld_pl + nn) = pl;
  }
}

}  // namespace efts
