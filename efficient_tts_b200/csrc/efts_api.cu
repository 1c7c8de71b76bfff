// C ABI of the B200-native EFTS-CNN forward path (include/efts_b200.h): context, weight prepack,
// TMA tensor maps, kernel launches and the forward()/inference() schedules.
// Reference citations are relative to /root/reference/nntts.
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <tuple>
#include <vector>

#include "../../include/efts_b200.h"
#include "gemm2_sm100.cuh"
#include "path_kernels.cuh"
#include "stack_sm100.cuh"
#include "vocoder_kernels.cuh"
#include "frontend_kernels.cuh"
#include "train_kernels.cuh"

namespace {

using namespace efts;

thread_local char g_err[512] = "";

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

#define CUDA_TRY(expr)                                                                         \
  do {                                                                                         \
    cudaError_t _e = (expr);                                                                   \
    if (_e != cudaSuccess)                                                                     \
      return fail(EFTS_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, \
                  __LINE__);                                                                   \
  } while (0)

#define TRY(expr)              \
  do {                         \
    int _r = (expr);           \
    if (_r != EFTS_OK) return _r; \
  } while (0)

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);

// fp16 hi/lo operand planes of one weight tensor, tap-major [Z][N][K].
struct PackedW {
  __half* hi = nullptr;
  __half* lo = nullptr;
  float* bias = nullptr;
  int Z = 0, N = 0, K = 0;
};

// [B, T, ld], K valid columns.  map_rows > T: every utterance owns map_rows rows in memory of which the kernel computes the
// first T (its taps may read on into the others): the front-end's chunk matrix, T frames from T + taps - 1 chunk rows.
struct OpA { const __half* hi; const __half* lo; int B, T, K, ld; int map_rows = 0; };
struct OpB { const __half* hi; const __half* lo; int Z, N, K, ld; };   // [Z, N, ld], K valid columns

inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }
inline int round8(int x) { return (x + 7) / 8 * 8; }

// Bump allocator over the caller's workspace.
struct Arena {
  char* base;
  size_t cap, off = 0;
  bool ok = true;
  Arena(void* p, size_t n) : base(static_cast<char*>(p)), cap(n) {}
  template <typename T>
  T* get(size_t count) {
    off = align_up(off, 256);
    T* r = reinterpret_cast<T*>(base + off);
    off += count * sizeof(T);
    if (off > cap) ok = false;
    return r;
  }
};

}  // namespace

struct efts_ctx {
  efts_config cfg;
  int sm_count = 0;
  int skip_pad_tiles = 1;
  int pair = 1;              // CTA pairs (cta_group::2) for the weight GEMMs
  int cur_tag = 15;          // ProfTag of the launch being issued (diagnostics)
  int wide = 1;              // short-reduction launches use the 16-epilogue-warp variant
  int stack_trace_on = 0;    // measurement hook: the stack kernel records its phase stamps (efts_profile_stack_trace)
  int chunk_kb = 2;          // k-blocks per main-accumulator flush (1 = most accurate, 0 = never)
  int voc_group = 1;             // vocoder: grouped (super-tap) packing of the 32 / 64-channel layers (read at finalize)
  int voc_wide = 1;              // vocoder: short-reach, short-reduction layers use the wide-epilogue variant
  int voc_narrow = 0;            // vocoder: 64-column tiles for layers with N <= 64 (read at finalize and at launch).
                                 // Off: measured 36.4 vs 34.5 ms at 16 x 800 frames -- the narrow layers are bound by
                                 // their tile count (A-box halo, epilogue), which the grouped packing halves, not by MMA columns
  int stack = 1;                 // B = 1 synthesis: resident layer-stack kernel (stack_sm100.cuh) instead of one launch
                                 // per layer (bitwise the same results; the switch keeps the per-layer path tested)
  unsigned* sync_ctr = nullptr;  // grid-barrier counter of the stack kernel (device, monotonic)
  unsigned sync_base = 0;        // its value once every enqueued stack launch has finished
  long long* stack_trace = nullptr;   // device buffer [64] for the stack kernel's phase stamps (option "stack_trace")
  int split_k = 1;               // fused-B kernel: split the reduction of small launches over more SMs
  int pdl = 1;                   // programmatic dependent launch for the GEMM and split-reduce kernels
  int fuse_b = 1;                // conv layers: Ahi*[Bhi|Blo] as one N = 256 MMA (two MMAs per k-step instead of three)
  int imv_version = 2;           // 2: block-per-utterance scan / aligned positions out of shared memory; 1: the
                                 // warp-per-row kernels that serve rows too long for shared memory (same arithmetic
                                 // in the same order, bitwise equal -- the switch exists so that path stays tested)
  int64_t launches = 0;
  bool finalized = false;
  EncodeTiledFn encode = nullptr;
  std::map<std::string, std::vector<float>> raw;          // host staging until finalize
  std::map<std::string, std::vector<int64_t>> raw_shape;
  std::vector<void*> device_allocs;
  PackedW text[16], mel[16], dec[16], dp[4];
  PackedW key, value, prenet, melout;
  float* emb = nullptr;
  float* ln_g[4] = {nullptr, nullptr, nullptr, nullptr};
  float* ln_b[4] = {nullptr, nullptr, nullptr, nullptr};
  float* head_w = nullptr;
  float* head_b = nullptr;
  int* err_flag = nullptr;   // device word: bit 3 = activation outside the fp16 operand range
  // HiFi-GAN generator contexts (efts_vocoder_create) carry their layers here; cfg above is unused for them
  struct Vocoder {
    efts_vocoder_config cfg;
    PackedW conv_pre;
    PackedW ups[8];                    // transposed conv as a 3-tap GEMM with N = rate * C (polyphase columns)
    PackedW c1[32][3], c2[32][3];      // resblocks[n].convs1[m] / convs2[m]
    int g1[32][3], g2[32][3];          // time steps per GEMM row of that layer (1 = plain [B, L, C] view, see pack_grouped)
    int d1[32][3];                     // dilation handed to the kernel (1 when the grouped packing absorbed it)
    float* post_w = nullptr;           // [7][C_last]
    float post_b = 0.0f;
  };
  Vocoder* voc = nullptr;
  // log-mel front-end contexts (efts_frontend_create): windowed DFT basis [taps][2 * half][hop] and mel basis [mels][Kp]
  struct Frontend {
    efts_frontend_config cfg;
    PackedW dft, mel;
    int taps = 0, half = 0, Kp = 0;   // Kp: bins kept (a multiple of 8): up to the last one any mel filter weighs
  };
  Frontend* fe = nullptr;
  // measurement hooks (efts_profile_*): CUDA-event pairs around tagged launches
  struct ProfRec { cudaEvent_t a, b; int tag; };
  std::vector<ProfRec> prof;
  size_t prof_used = 0;
  typedef std::tuple<const void*, int, int, int, int, int> MapKey;   // (pointer, inner, rows, z, ld, box rows)
  std::map<MapKey, CUtensorMap> map_cache;                           // encoded TMA tensor maps (make_map)
  int32_t* pinned_words = nullptr;                                   // host staging of efts_read_words (pinned, 64 words)
  char tag_kernel[16][64] = {};    // instantiation of the last GEMM launched under each tag (efts_profile_kernel_name)
  uint32_t profile_mask = 0;
};

namespace {

// Records an event pair around the launches issued while it is alive (only when the tag is enabled).
struct ProfScope {
  efts_ctx* c; cudaStream_t st; int idx = -1;
  ProfScope(efts_ctx* c_, cudaStream_t st_, int tag) : c(c_), st(st_) {
    c->cur_tag = tag;
    if (!(c->profile_mask & (1u << tag))) return;
    if (c->prof_used == c->prof.size()) {
      efts_ctx::ProfRec r;
      r.tag = tag;
      if (cudaEventCreate(&r.a) != cudaSuccess || cudaEventCreate(&r.b) != cudaSuccess) return;
      c->prof.push_back(r);
    }
    idx = static_cast<int>(c->prof_used++);
    c->prof[idx].tag = tag;
    cudaEventRecord(c->prof[idx].a, st);
  }
  ~ProfScope() { if (idx >= 0) cudaEventRecord(c->prof[idx].b, st); }
};
enum ProfTag { TAG_TEXT_CONV = 0, TAG_MEL_CONV = 1, TAG_DEC_CONV = 2, TAG_LINEAR = 3, TAG_ENERGY = 4,
               TAG_SOFTMAX = 5, TAG_SCAN = 6, TAG_ALIGNED = 7, TAG_RECONSTRUCT = 8, TAG_EXPAND = 9,
               TAG_DURATION = 10, TAG_LOSS = 11, TAG_PREP = 12 };

// ------------------------------------------------------------------------------------------------
// tensor maps
// Encoding a map costs ~1 us of host time and every GEMM launch needs four; the operands of a schedule recur
// (weights always, activations whenever the caller's workspace and sizes repeat), so encoded maps are kept per context.
int make_map(efts_ctx* c, CUtensorMap* m, const __half* ptr, int inner, int rows, int z, int ld,
             int box_rows) {
  const efts_ctx::MapKey key{ptr, inner, rows, z, ld, box_rows};
  auto it = c->map_cache.find(key);
  if (it != c->map_cache.end()) {
    *m = it->second;
    return EFTS_OK;
  }
  cuuint64_t dims[3] = {static_cast<cuuint64_t>(inner), static_cast<cuuint64_t>(rows),
                        static_cast<cuuint64_t>(z)};
  cuuint64_t strides[2] = {static_cast<cuuint64_t>(ld) * 2, static_cast<cuuint64_t>(rows) * ld * 2};
  cuuint32_t box[3] = {static_cast<cuuint32_t>(G2_BK), static_cast<cuuint32_t>(box_rows), 1};
  cuuint32_t estr[3] = {1, 1, 1};
  if ((reinterpret_cast<uintptr_t>(ptr) & 15) != 0 || (strides[0] & 15) != 0)
    return fail(EFTS_ERR_ARG, "TMA operand not 16-byte aligned (ptr %p, ld %d)", (const void*)ptr, ld);
  CUresult r = c->encode(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, const_cast<__half*>(ptr), dims, strides, box,
                         estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                         CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return fail(EFTS_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d) inner=%d rows=%d z=%d ld=%d box=%d",
                (int)r, inner, rows, z, ld, box_rows);
  if (c->map_cache.size() >= 4096) c->map_cache.clear();
  c->map_cache.emplace(key, *m);
  return EFTS_OK;
}

constexpr size_t kReconstructSmemMax = 200 * 1024;
constexpr size_t kSplitScratchBytes = 16u << 20;       // partial planes of a split reduction (workspace, text side)

template <int CG, int EPI, int WIDE, int FUSE = 0, int AR = G2_A_ROWS, int BN = G2_BN, int SPLIT = 0, int BMN = 0>
int launch_gemm2_t(efts_ctx* c, cudaStream_t st, const OpA& a, const OpB& b, const GemmParams& p) {
  using Cfg = G2Cfg<CG, WIDE, FUSE, AR, BN>;
  auto kern = gemm2_kernel<CG, EPI, WIDE, FUSE, AR, BN, SPLIT, BMN>;
  if (!SPLIT && p.splits > 1) return fail(EFTS_ERR_ARG, "split reduction requested from an unsplit kernel instantiation");
  const int dil = p.dil > 1 ? p.dil : 1;
  if (G2_BM + (p.ntaps - 1) * dil > AR)
    return fail(EFTS_ERR_ARG, "%d taps with dilation %d need a %d-row A box (this variant holds %d)", p.ntaps, dil,
                G2_BM + (p.ntaps - 1) * dil, AR);
  if (p.bias != nullptr && p.N > Cfg::BIAS_MAX) return fail(EFTS_ERR_ARG, "bias supports at most %d columns", Cfg::BIAS_MAX);
  alignas(64) CUtensorMap ma_hi, ma_lo, mb_hi, mb_lo;
  const int a_rows = a.map_rows > a.T ? a.map_rows : a.T;
  // MN-major A (BMN == 2): 64-row boxes of the planes [B, T, C] themselves, like the B operand's
  TRY(make_map(c, &ma_hi, a.hi, a.K, a_rows, a.B, a.ld, BMN == 2 ? Cfg::B_ROWS : AR));
  TRY(make_map(c, &ma_lo, a.lo, a.K, a_rows, a.B, a.ld, BMN == 2 ? Cfg::B_ROWS : AR));
  TRY(make_map(c, &mb_hi, b.hi, b.K, b.N, b.Z, b.ld, Cfg::B_ROWS));
  TRY(make_map(c, &mb_lo, b.lo, b.K, b.N, b.Z, b.ld, Cfg::B_ROWS));
  // persistent: one CTA per SM (CTA pairs when CG == 2), never more than there are tiles
  const long long n_rt = static_cast<long long>(p.B) * ((p.T + G2_BM - 1) / G2_BM);
  const long long work = ((n_rt + CG - 1) / CG) * ((p.N + BN - 1) / BN) * (SPLIT && p.splits > 1 ? p.splits : 1);
  if (work >= (1ll << 31)) return fail(EFTS_ERR_ARG, "launch of %lld work items exceeds the kernel's 32-bit item index", work);
  long long ctas = std::min<long long>(c->sm_count / CG, work) * CG;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(static_cast<unsigned>(ctas));
  cfg.blockDim = dim3(WIDE ? G2_THREADS_WIDE : G2_THREADS);
  cfg.dynamicSmemBytes = Cfg::SMEM_BYTES;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CG;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;   // prologue overlaps the predecessor's tail
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = c->pdl ? 2 : 1;
  CUDA_TRY(cudaLaunchKernelEx(&cfg, kern, ma_hi, ma_lo, mb_hi, mb_lo, p));
  c->launches++;
  if (c->profile_mask)
    snprintf(c->tag_kernel[c->cur_tag & 15], sizeof(c->tag_kernel[0]), "gemm2_kernel<%d, %d, %d, %d, %d, %d, %d, %d>", CG, EPI,
             WIDE, FUSE, AR, BN, SPLIT, BMN);
  return EFTS_OK;
}

int launch_gemm(efts_ctx* c, cudaStream_t st, const OpA& a, const OpB& b, GemmParams p) {
  p.B = a.B; p.T = a.T; p.K = a.K;
  static const bool trace = getenv("EFTS_TRACE_GEMM") != nullptr;      // diagnostics: one line per tensor-core launch
  if (trace)
    fprintf(stderr, "[efts gemm] B=%d T=%d K=%d N=%d taps=%d dil=%d act=%d resid=%d out=%d planes=%d plane_act=%d long=%d batched=%d\n",
            a.B, a.T, a.K, p.N, p.ntaps, p.dil, p.act, p.resid != nullptr, p.out != nullptr, p.out_hi != nullptr,
            p.plane_act, p.long_taps, p.b_batched);
  if (a.K != b.K && p.bmn_per == 0) return fail(EFTS_ERR_ARG, "gemm K mismatch %d vs %d", a.K, b.K);
  if (p.N % 8 != 0) return fail(EFTS_ERR_ARG, "gemm N=%d must be a multiple of 8", p.N);
  {
    if (p.ntaps > 15) return fail(EFTS_ERR_ARG, "at most 15 taps");
    p.chunk_kb = c->chunk_kb;
    p.err_flag = c->err_flag;
    p.err_code = 1 << (8 + c->cur_tag);
    if (!c->skip_pad_tiles) { p.tile_list = nullptr; p.tile_count = nullptr; p.skip_lens = nullptr; }
    const int epi = p.softmax_part != nullptr ? EPI_SOFTMAX
                    : (p.divisor != 1.0f || p.outT_hi != nullptr) ? EPI_FULL : EPI_STD;
    const bool pair = c->pair && !p.b_batched;
    // short reductions (<= 40 MMA steps: one accumulation chain is as accurate as a flushed one) are bound by
    // their epilogue: they take the wide variant, which reads the single chunk straight from tensor memory
    const int steps = p.ntaps * ((p.K + G2_BK - 1) / G2_BK) * (G2_BK / 16);
    // vocoder layers with a short reach and a short reduction (k = 3, most grouped layers: <= 40 MMA steps) are bound by
    // their store epilogue like the Linear layers: they take the wide variant (16 epilogue warps) through the normal path
    if (p.long_taps && c->voc_wide && c->wide && epi == EPI_STD && !p.b_batched && (p.N > 64 || c->voc_wide > 1) &&
        G2_BM + (p.ntaps - 1) * (p.dil > 1 ? p.dil : 1) <= G2_A_ROWS && (p.bias == nullptr || p.N <= G2_BIAS_MAX) &&
        p.ntaps * ((p.K + G2_BK - 1) / G2_BK) * (G2_BK / 16) <= 40)
      p.long_taps = 0;
    if (p.long_taps) {        // vocoder layers: up to 11 taps / dilation 5, always the fused-B pair kernel
      if (epi != EPI_STD || p.b_batched || p.chunk_kb < 1)
        return fail(EFTS_ERR_ARG, "long-tap launches are plain weight GEMMs");
      p.splits = 0;
      const int box = G2_BM + (p.ntaps - 1) * (p.dil > 1 ? p.dil : 1);
      const bool xlong = box > G2_A_ROWS_LONG;
      if (p.N <= 64 && c->voc_narrow)     // 64-column tiles: half the MMA columns of a 128-column tile
        return xlong ? launch_gemm2_t<2, EPI_STD, 0, 1, G2_A_ROWS_XLONG, 64>(c, st, a, b, p)
                     : launch_gemm2_t<2, EPI_STD, 0, 1, G2_A_ROWS_LONG, 64>(c, st, a, b, p);
      if (xlong) return launch_gemm2_t<2, EPI_STD, 0, 1, G2_A_ROWS_XLONG>(c, st, a, b, p);
      return launch_gemm2_t<2, EPI_STD, 0, 1, G2_A_ROWS_LONG>(c, st, a, b, p);
    }
    if (p.mag_pairs) {                       // STFT of the log-mel front-end: magnitude planes from (re, im) column pairs
      if (epi != EPI_STD || !pair || !c->fuse_b || p.chunk_kb < 1 || p.N % 16 != 0 || p.out_hi == nullptr || p.out != nullptr ||
          p.bias != nullptr || p.resid != nullptr)
        return fail(EFTS_ERR_ARG, "magnitude epilogue: plain paired GEMM with plane output only");
      return launch_gemm2_t<2, EPI_MAG, 0, 1>(c, st, a, b, p);
    }
    if (p.bmn_per > 0) {
      // Weight gradients: B = activation planes [Z = utterances, N = rows, K = channels] read MN-major by the split
      // fused-B kernel, whatever the size (the reduction index is the activation's row, so no other variant applies).
      // amn: A is given the same way -- planes [B, T, C] of the masked gradient, M = its channels -- instead of as
      // K-major transposed planes: the kernel then sees one "utterance" of C rows and K = B * bmn_per k-blocks.
      if (p.amn) {
        if (a.B != b.Z || a.T != b.N) return fail(EFTS_ERR_ARG, "MN-major operands: planes of different shapes");
        p.bmn_batches = a.B;
        p.K = a.B * p.bmn_per * G2_BK; p.T = a.K; p.B = 1;
      }
      const int num_kb = (p.K + G2_BK - 1) / G2_BK;
      if (p.split_kb <= 0) p.split_kb = (num_kb + p.chunk_kb - 1) / std::max(1, p.chunk_kb) * std::max(1, p.chunk_kb);
      const int splits = (num_kb + p.split_kb - 1) / p.split_kb;
      const size_t plane = static_cast<size_t>(p.B) * p.T * p.N;
      if (epi != EPI_STD || !pair || !c->fuse_b || p.chunk_kb < 1 || p.N != b.K || p.N % (2 * G2Cfg<2, 0, 1>::B_ROWS) != 0 ||
          p.split_scratch == nullptr || p.split_kb % p.chunk_kb != 0 || plane * splits * sizeof(float) > kSplitScratchBytes ||
          p.ld_out % 4 != 0 || p.tile_list != nullptr || p.skip_lens != nullptr || p.ntaps != 1)
        return fail(EFTS_ERR_ARG, "MN-major B operand: %d output columns over planes of %d channels, %d splits", p.N, b.K, splits);
      GemmParams q = p;
      q.bias = nullptr; q.act = ACT_NONE; q.resid = nullptr; q.lens = nullptr;
      q.out = p.split_scratch; q.ld_out = p.N; q.out_hi = nullptr; q.out_lo = nullptr;
      q.splits = splits; q.split_stride = plane;
      if (p.amn) {
        TRY((launch_gemm2_t<2, EPI_STD, 0, 1, G2_A_ROWS, G2_BN, 1, 2>(c, st, a, b, q)));
      } else {
        TRY((launch_gemm2_t<2, EPI_STD, 0, 1, G2_A_ROWS, G2_BN, 1, 1>(c, st, a, b, q)));
      }
      splitk_reduce_kernel<<<static_cast<unsigned>((plane / 4 + 255) / 256), 256, 0, st>>>(p, p.split_scratch, splits, plane);
      CUDA_TRY(cudaGetLastError());
      c->launches++;
      return EFTS_OK;
    }
    const bool wide = c->wide && steps <= 40;
    if (p.act == ACT_LOGCLAMP && !wide)
      return fail(EFTS_ERR_UNSUPPORTED, "the log epilogue is compiled into the short-reduction (wide) variant only");
    if (wide) {
      p.chunk_kb = 0;
      if (epi == EPI_STD) return pair ? launch_gemm2_t<2, EPI_STD, 1>(c, st, a, b, p) : launch_gemm2_t<1, EPI_STD, 1>(c, st, a, b, p);
      if (epi == EPI_SOFTMAX) return pair ? launch_gemm2_t<2, EPI_SOFTMAX, 1>(c, st, a, b, p) : launch_gemm2_t<1, EPI_SOFTMAX, 1>(c, st, a, b, p);
      return pair ? launch_gemm2_t<2, EPI_FULL, 1>(c, st, a, b, p) : launch_gemm2_t<1, EPI_FULL, 1>(c, st, a, b, p);
    }
    if (epi == EPI_STD && pair && c->fuse_b && p.chunk_kb > 0) {
      // small problems (B = 1 synthesis: a handful of 27-us tiles on 148 SMs): one work item per accumulation
      // chunk, partial planes summed in chunk order by splitk_reduce_kernel -- bitwise the unsplit result
      const int num_kb = (p.K + G2_BK - 1) / G2_BK;
      if (p.split_kb > 0 && p.split_scratch != nullptr) {
        // long reductions over few output tiles (weight gradients: K = positions): every work item covers split_kb
        // k-blocks, the partial planes are summed by splitk_reduce_kernel
        const int splits = (num_kb + p.split_kb - 1) / p.split_kb;
        const size_t plane = static_cast<size_t>(p.B) * p.T * p.N;
        if (splits < 2) { p.split_kb = 0; return launch_gemm2_t<2, EPI_STD, 0, 1>(c, st, a, b, p); }
        if (p.split_kb % p.chunk_kb != 0 || plane * splits * sizeof(float) > kSplitScratchBytes || p.N % 4 != 0 ||
            p.ld_out % 4 != 0 || p.tile_list != nullptr || p.skip_lens != nullptr)
          return fail(EFTS_ERR_ARG, "split reduction: %d splits of %d k-blocks do not fit (plane %zu)", splits, p.split_kb, plane);
        GemmParams q = p;
        q.bias = nullptr; q.act = ACT_NONE; q.resid = nullptr; q.lens = nullptr;
        q.out = p.split_scratch; q.ld_out = p.N; q.out_hi = nullptr; q.out_lo = nullptr;
        q.splits = splits; q.split_stride = plane;
        TRY((launch_gemm2_t<2, EPI_STD, 0, 1, G2_A_ROWS, G2_BN, 1>(c, st, a, b, q)));
        const size_t n = plane / 4;
        const float* part = p.split_scratch;
        splitk_reduce_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, st>>>(p, part, splits, plane);
        CUDA_TRY(cudaGetLastError());
        c->launches++;
        return EFTS_OK;
      }
      const int nchunks = (num_kb + p.chunk_kb - 1) / p.chunk_kb;
      const long long n_rt = static_cast<long long>(p.B) * ((p.T + G2_BM - 1) / G2_BM);
      const long long items = ((n_rt + 1) / 2) * ((p.N + G2_BN - 1) / G2_BN);
      const size_t plane = static_cast<size_t>(p.B) * p.T * p.N;
      if (c->split_k && p.split_scratch != nullptr && p.tile_list == nullptr && p.skip_lens == nullptr &&
          nchunks > 1 && nchunks <= 8 && items * 4 <= c->sm_count && p.N % 4 == 0 && p.ld_out % 4 == 0 &&
          p.ld_pl % 4 == 0 && plane * nchunks * sizeof(float) <= kSplitScratchBytes) {
        GemmParams q = p;
        q.bias = nullptr; q.act = ACT_NONE; q.resid = nullptr; q.lens = nullptr;
        q.out = p.split_scratch; q.ld_out = p.N; q.out_hi = nullptr; q.out_lo = nullptr;
        q.splits = nchunks; q.split_stride = plane;
        TRY((launch_gemm2_t<2, EPI_STD, 0, 1, G2_A_ROWS, G2_BN, 1>(c, st, a, b, q)));
        const size_t n = plane / 4;
        cudaLaunchConfig_t rc = {};
        rc.gridDim = dim3(static_cast<unsigned>((n + 255) / 256));
        rc.blockDim = dim3(256);
        rc.stream = st;
        cudaLaunchAttribute ra[1];
        ra[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        ra[0].val.programmaticStreamSerializationAllowed = 1;
        rc.attrs = ra;
        rc.numAttrs = c->pdl ? 1 : 0;
        const float* part = p.split_scratch;
        CUDA_TRY(cudaLaunchKernelEx(&rc, splitk_reduce_kernel, p, part, nchunks, plane));
        c->launches++;
        return EFTS_OK;
      }
      return launch_gemm2_t<2, EPI_STD, 0, 1>(c, st, a, b, p);
    }
    if (epi == EPI_STD) return pair ? launch_gemm2_t<2, EPI_STD, 0>(c, st, a, b, p) : launch_gemm2_t<1, EPI_STD, 0>(c, st, a, b, p);
    if (epi == EPI_FULL) return pair ? launch_gemm2_t<2, EPI_FULL, 0>(c, st, a, b, p) : launch_gemm2_t<1, EPI_FULL, 0>(c, st, a, b, p);
    return pair ? launch_gemm2_t<2, EPI_SOFTMAX, 0>(c, st, a, b, p) : launch_gemm2_t<1, EPI_SOFTMAX, 0>(c, st, a, b, p);
  }
}

// Opt every kernel that needs more than 48 KB of dynamic shared memory in, once per context (the attribute
// is per device, so it is not cached in a process-wide static).
int set_kernel_attributes() {
#define EFTS_OPT_IN_V2(CG_, EPI_, W_) \
  CUDA_TRY(cudaFuncSetAttribute(gemm2_kernel<CG_, EPI_, W_>, cudaFuncAttributeMaxDynamicSharedMemorySize, G2Cfg<CG_, W_>::SMEM_BYTES))
  EFTS_OPT_IN_V2(1, EPI_STD, 0); EFTS_OPT_IN_V2(1, EPI_FULL, 0); EFTS_OPT_IN_V2(1, EPI_SOFTMAX, 0);
  EFTS_OPT_IN_V2(2, EPI_STD, 0); EFTS_OPT_IN_V2(2, EPI_FULL, 0); EFTS_OPT_IN_V2(2, EPI_SOFTMAX, 0);
  EFTS_OPT_IN_V2(1, EPI_STD, 1); EFTS_OPT_IN_V2(1, EPI_FULL, 1); EFTS_OPT_IN_V2(1, EPI_SOFTMAX, 1);
  EFTS_OPT_IN_V2(2, EPI_STD, 1); EFTS_OPT_IN_V2(2, EPI_FULL, 1); EFTS_OPT_IN_V2(2, EPI_SOFTMAX, 1);
#undef EFTS_OPT_IN_V2
  CUDA_TRY(cudaFuncSetAttribute(gemm2_kernel<2, EPI_STD, 0, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                G2Cfg<2, 0, 1>::SMEM_BYTES));
  CUDA_TRY(cudaFuncSetAttribute(gemm2_kernel<2, EPI_STD, 0, 1, G2_A_ROWS, G2_BN, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                G2Cfg<2, 0, 1>::SMEM_BYTES));
  CUDA_TRY(cudaFuncSetAttribute(gemm2_kernel<2, EPI_MAG, 0, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                G2Cfg<2, 0, 1>::SMEM_BYTES));
  CUDA_TRY(cudaFuncSetAttribute(gemm2_kernel<2, EPI_STD, 0, 1, G2_A_ROWS, G2_BN, 1, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                G2Cfg<2, 0, 1>::SMEM_BYTES));
  CUDA_TRY(cudaFuncSetAttribute(gemm2_kernel<2, EPI_STD, 0, 1, G2_A_ROWS, G2_BN, 1, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                G2Cfg<2, 0, 1>::SMEM_BYTES));
  CUDA_TRY(cudaFuncSetAttribute(gemm2_kernel<2, EPI_STD, 0, 1, G2_A_ROWS_LONG>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                G2Cfg<2, 0, 1, G2_A_ROWS_LONG>::SMEM_BYTES));
  CUDA_TRY(cudaFuncSetAttribute(gemm2_kernel<2, EPI_STD, 0, 1, G2_A_ROWS_XLONG>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                G2Cfg<2, 0, 1, G2_A_ROWS_XLONG>::SMEM_BYTES));
  CUDA_TRY(cudaFuncSetAttribute(gemm2_kernel<2, EPI_STD, 0, 1, G2_A_ROWS_LONG, 64>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                G2Cfg<2, 0, 1, G2_A_ROWS_LONG, 64>::SMEM_BYTES));
  CUDA_TRY(cudaFuncSetAttribute(gemm2_kernel<2, EPI_STD, 0, 1, G2_A_ROWS_XLONG, 64>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                G2Cfg<2, 0, 1, G2_A_ROWS_XLONG, 64>::SMEM_BYTES));
  CUDA_TRY(cudaFuncSetAttribute(stack_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ST_SMEM_BYTES));
  CUDA_TRY(cudaFuncSetAttribute(reconstruct_alignment_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                static_cast<int>(kReconstructSmemMax)));
  CUDA_TRY(cudaFuncSetAttribute(imv_scan_block_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                static_cast<int>(kReconstructSmemMax)));
  CUDA_TRY(cudaFuncSetAttribute(aligned_positions_block_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                static_cast<int>(kReconstructSmemMax)));
  return EFTS_OK;
}

// IMV scan / aligned positions: version 2 = one block per utterance out of shared memory, 1 = one warp per
// utterance / token from global memory (A/B baseline, and the fallback when a row does not fit shared memory).
int launch_imv_scan(int version, cudaStream_t st, const float* imv_raw, const float4* part, int n_part, const int* tl,
                    const int* sl, int B, int T2, float* imv) {
  const size_t smem = static_cast<size_t>(T2) * sizeof(float);
  if (version >= 2 && smem <= kReconstructSmemMax)
    imv_scan_block_kernel<<<B, IMV_BLOCK_THREADS, smem, st>>>(imv_raw, part, n_part, tl, sl, T2, imv);
  else
    imv_scan_kernel<<<(B + 3) / 4, 128, 0, st>>>(imv_raw, part, n_part, tl, sl, B, T2, imv);
  CUDA_TRY(cudaGetLastError());
  return EFTS_OK;
}
int launch_aligned_positions(int version, cudaStream_t st, const float* imv, const int* tl, const int* sl, int B,
                             int T1, int T2, float sigma_e, float* e, const float* pvec) {
  const size_t smem = static_cast<size_t>(T2) * sizeof(float);
  if (version >= 2 && smem <= kReconstructSmemMax)
    aligned_positions_block_kernel<<<dim3((T1 + AP_TOKENS - 1) / AP_TOKENS, B), IMV_BLOCK_THREADS, smem, st>>>(
        imv, tl, sl, T1, T2, sigma_e, e, pvec);
  else
    aligned_positions_kernel<<<dim3((T1 + 7) / 8, B), 256, 0, st>>>(imv, tl, sl, T1, T2, sigma_e, e, pvec);
  CUDA_TRY(cudaGetLastError());
  return EFTS_OK;
}

// Gaussian reconstruction launch: the frame-per-lane kernel; token counts that do not fit its shared-memory tile
// take the one-thread-per-frame kernel (same arithmetic, partial sums grouped differently).
int launch_reconstruct(cudaStream_t st, const float* e, const int* tl, const int* sl, int B, int T1, int T2, int T1p,
                       float neg_sigma, float* reconst_alpha, __half* R_hi, __half* R_lo) {
  if (r3_smem_bytes(T1p) <= kReconstructSmemMax) {
    dim3 grid((T2 + R3_FRAMES - 1) / R3_FRAMES, B);
    reconstruct_alignment_rows_kernel<<<grid, 32 * R3_WARPS, r3_smem_bytes(T1p), st>>>(
        e, tl, sl, T1, T2, T1p, neg_sigma, reconst_alpha, R_hi, R_lo);
  } else {
    dim3 grid((T2 + 127) / 128, B);
    reconstruct_alignment_kernel<<<grid, 128, T1 * sizeof(float), st>>>(e, tl, sl, T1, T2, T1p, neg_sigma,
                                                                        reconst_alpha, R_hi, R_lo);
  }
  CUDA_TRY(cudaGetLastError());
  return EFTS_OK;
}

GemmParams gemm_defaults() {
  GemmParams p;
  memset(&p, 0, sizeof(p));
  p.ntaps = 1;
  p.divisor = 1.0f;
  return p;
}

OpB weight_op(const PackedW& w) { return OpB{w.hi, w.lo, w.Z, w.N, w.K, w.K}; }

// ------------------------------------------------------------------------------------------------
// weights
int upload(efts_ctx* c, const void* host, size_t bytes, void** dev) {
  CUDA_TRY(cudaMalloc(dev, bytes));
  c->device_allocs.push_back(*dev);
  CUDA_TRY(cudaMemcpy(*dev, host, bytes, cudaMemcpyHostToDevice));
  return EFTS_OK;
}

int need(efts_ctx* c, const std::string& name, std::initializer_list<int64_t> shape,
         const std::vector<float>** out) {
  auto it = c->raw.find(name);
  if (it == c->raw.end()) return fail(EFTS_ERR_STATE, "weight '%s' was never set", name.c_str());
  const std::vector<int64_t>& got = c->raw_shape[name];
  std::vector<int64_t> want(shape);
  if (got != want) {
    std::string g, w;
    for (auto v : got) g += std::to_string(v) + ",";
    for (auto v : want) w += std::to_string(v) + ",";
    return fail(EFTS_ERR_ARG, "weight '%s' has shape [%s], expected [%s]", name.c_str(), g.c_str(), w.c_str());
  }
  *out = &it->second;
  return EFTS_OK;
}

// [N, K, taps] (torch Conv1d / Linear layout, taps == 1 for Linear) -> planes [taps][N][K]
int pack_weight(efts_ctx* c, const std::string& wname, const std::string& bname, int N, int K, int taps,
                PackedW* out) {
  const std::vector<float>* w;
  const std::vector<float>* b;
  if (taps == 1 && c->raw_shape.count(wname) && c->raw_shape[wname].size() == 2) {
    TRY(need(c, wname, {N, K}, &w));
  } else {
    TRY(need(c, wname, {N, K, taps}, &w));
  }
  TRY(need(c, bname, {N}, &b));
  const size_t n = static_cast<size_t>(taps) * N * K;
  std::vector<__half> hi(n), lo(n);
  for (int z = 0; z < taps; ++z)
    for (int o = 0; o < N; ++o)
      for (int k = 0; k < K; ++k) {
        const float x = (*w)[(static_cast<size_t>(o) * K + k) * taps + z];
        if (!(fabsf(x) <= 65504.0f))
          return fail(EFTS_ERR_UNSUPPORTED, "weight '%s' has a value outside the fp16 operand range", wname.c_str());
        const __half h = __float2half_rn(x);
        const size_t d = (static_cast<size_t>(z) * N + o) * K + k;
        hi[d] = h;
        lo[d] = __float2half_rn((x - __half2float(h)) * SPLIT_SCALE);
      }
  out->Z = taps; out->N = N; out->K = K;
  TRY(upload(c, hi.data(), n * sizeof(__half), reinterpret_cast<void**>(&out->hi)));
  TRY(upload(c, lo.data(), n * sizeof(__half), reinterpret_cast<void**>(&out->lo)));
  TRY(upload(c, b->data(), N * sizeof(float), reinterpret_cast<void**>(&out->bias)));
  return EFTS_OK;
}

int upload_vec(efts_ctx* c, const std::string& name, std::initializer_list<int64_t> shape, float** dev) {
  const std::vector<float>* v;
  TRY(need(c, name, shape, &v));
  return upload(c, v->data(), v->size() * sizeof(float), reinterpret_cast<void**>(dev));
}

// ------------------------------------------------------------------------------------------------
// workspace layouts
struct FwdWs {
  int *tl32, *sl32, *flags;
  double* acc;
  float* xt_f[2]; __half *xt_hi[2], *xt_lo[2];
  __half *key_hi, *key_lo, *val_hi, *val_lo, *valT_hi, *valT_lo;
  float* dp_f; __half *dp_hi, *dp_lo;
  float *dur, *e;
  __half *sp_hi, *sp_lo;
  float* xm_f[2]; __half *xm_hi[2], *xm_lo[2];
  float *S, *imv_raw;
  __half *R_hi, *R_lo;
  float* splitk;              // partial planes of split reductions (kSplitScratchBytes)
  int2 *list_t, *list_m;      // compacted live row tiles (text side / mel side)
  int *cnt_t, *cnt_m;
  int T1p;
};

// Which rows of a padded batch can still reach a valid output: lengths + (optional) compacted tiles.
struct Skip { const int* lens; const int2* list; const int* count; };

void carve_text(Arena& a, FwdWs& w, int B, int T1, int C) {
  const size_t m1 = static_cast<size_t>(B) * T1;
  w.T1p = round8(T1);
  w.tl32 = a.get<int>(B);
  w.sl32 = a.get<int>(B);
  w.flags = a.get<int>(4);
  w.acc = a.get<double>(2);
  for (int i = 0; i < 2; ++i) {
    w.xt_f[i] = a.get<float>(m1 * C);
    w.xt_hi[i] = a.get<__half>(m1 * C);
    w.xt_lo[i] = a.get<__half>(m1 * C);
  }
  w.key_hi = a.get<__half>(m1 * C); w.key_lo = a.get<__half>(m1 * C);
  w.val_hi = a.get<__half>(m1 * C); w.val_lo = a.get<__half>(m1 * C);
  w.valT_hi = a.get<__half>(static_cast<size_t>(B) * C * w.T1p);
  w.valT_lo = a.get<__half>(static_cast<size_t>(B) * C * w.T1p);
  w.dp_f = a.get<float>(m1 * C);
  w.dp_hi = a.get<__half>(m1 * C); w.dp_lo = a.get<__half>(m1 * C);
  w.dur = a.get<float>(m1);
  w.e = a.get<float>(m1);
  w.list_t = a.get<int2>(static_cast<size_t>(B) * ((T1 + G2_BM - 1) / G2_BM));
  w.cnt_t = a.get<int>(4);
  w.splitk = a.get<float>(kSplitScratchBytes / sizeof(float));
}

void carve_mel(Arena& a, FwdWs& w, int B, int T2, int C, int odim, bool teacher_forced) {
  const size_t m2 = static_cast<size_t>(B) * T2;
  for (int i = 0; i < 2; ++i) {
    w.xm_f[i] = a.get<float>(m2 * C);
    w.xm_hi[i] = a.get<__half>(m2 * C);
    w.xm_lo[i] = a.get<__half>(m2 * C);
  }
  w.R_hi = a.get<__half>(m2 * w.T1p);
  w.R_lo = a.get<__half>(m2 * w.T1p);
  w.list_m = a.get<int2>(static_cast<size_t>(B) * ((T2 + G2_BM - 1) / G2_BM));
  w.cnt_m = a.get<int>(4);
  if (teacher_forced) {
    w.sp_hi = a.get<__half>(m2 * odim);
    w.sp_lo = a.get<__half>(m2 * odim);
    w.S = a.get<float>(m2 * w.T1p);
    w.imv_raw = a.get<float>(m2);
  }
}

// ------------------------------------------------------------------------------------------------
// schedules shared by the entry points

// x <- x + lrelu(conv_k(x) + b), n layers, ping-ponging between two (fp32, hi, lo) buffer sets.
// The input lives in set `cur`; returns the index of the set holding the output.  If `final_f`
// is non-null the last layer writes its fp32 result there instead (planes still go to the set).
int run_conv_stack(efts_ctx* c, cudaStream_t st, const PackedW* layers, int n, int B, int T, float* f[2],
                   __half* hi[2], __half* lo[2], const float* first_resid, float* final_f,
                   const Skip* skip, int* cur_io, int tag, bool ragged = false, float* splitk = nullptr) {
  const int C = c->cfg.n_channels;
  int cur = *cur_io;
  for (int l = 0; l < n; ++l) {
    const int nxt = cur ^ 1;
    GemmParams p = gemm_defaults();
    p.N = C;
    p.ntaps = layers[l].Z;
    p.pad = (layers[l].Z - 1) / 2;
    p.act = ACT_LRELU;
    p.bias = layers[l].bias;
    p.resid = (l == 0 && first_resid != nullptr) ? first_resid : f[cur];
    // the fp32 master of a layer is the next layer's residual; after the last layer only the caller can want it
    // (final_f) -- the consumers inside the path (Linear layers, energy GEMM, mel head) read the operand planes
    p.out = l == n - 1 ? final_f : f[nxt];
    p.ld_out = C;
    p.out_hi = hi[nxt]; p.out_lo = lo[nxt]; p.ld_pl = C;
    if (skip != nullptr && c->skip_pad_tiles) {
      p.skip_lens = skip->lens; p.tile_list = skip->list; p.tile_count = skip->count;
      p.skip_halo = ragged ? p.pad : p.pad * (n - 1 - l);
    }
    // ragged synthesis: every layer's rows t >= L_b are written as zeros, so utterance b convolves against
    // the zero padding it would see alone (tiles within `pad` rows of L_b are computed to produce those zeros)
    if (ragged) p.lens = skip->lens;
    p.split_scratch = splitk;
    {
      ProfScope ps(c, st, tag);
      TRY(launch_gemm(c, st, OpA{hi[cur], lo[cur], B, T, C, C}, weight_op(layers[l]), p));
    }
    cur = nxt;
  }
  *cur_io = cur;
  return EFTS_OK;
}

// Duration predictor body: two (conv k3 -> ReLU -> LayerNorm) layers and the linear head.
// Input operand planes in_hi/in_lo [B,T,C]; scratch dp_f, dp_hi/lo.
int run_duration_predictor(efts_ctx* c, cudaStream_t st, const __half* in_hi, const __half* in_lo, int B,
                           int T, float* dp_f, __half* dp_hi, __half* dp_lo, const int* lens, int mode,
                           void* out, bool mask_hidden = false, const Skip* skip = nullptr, float* splitk = nullptr) {
  const int C = c->cfg.n_channels;
  const size_t rows = static_cast<size_t>(B) * T;
  const int nl = c->cfg.n_duration_layer;
  const __half* ahi = in_hi;
  const __half* alo = in_lo;
  ProfScope ps(c, st, TAG_DURATION);
  for (int l = 0; l < nl; ++l) {
    GemmParams p = gemm_defaults();
    p.N = C;
    p.ntaps = c->dp[l].Z;
    p.pad = (c->dp[l].Z - 1) / 2;
    p.act = ACT_RELU;
    p.bias = c->dp[l].bias;
    p.out = dp_f; p.ld_out = C;
    if (skip != nullptr && c->skip_pad_tiles) {   // only rows that can reach a valid token's duration
      p.skip_lens = skip->lens; p.tile_list = skip->list; p.tile_count = skip->count;
      p.skip_halo = p.pad * (nl - 1 - l);
    }
    p.split_scratch = splitk;
    TRY(launch_gemm(c, st, OpA{ahi, alo, B, T, C, C}, weight_op(c->dp[l]), p));
    const int wpb = 8;
    const unsigned grid = static_cast<unsigned>((rows + wpb - 1) / wpb);
    if (l < nl - 1) {
      layernorm_kernel<0><<<grid, wpb * 32, 0, st>>>(dp_f, rows, T, c->ln_g[l], c->ln_b[l], dp_hi, dp_lo,
                                                     nullptr, nullptr, mask_hidden ? lens : nullptr, 0, 0.0f,
                                                     nullptr);
    } else {
      layernorm_kernel<1><<<grid, wpb * 32, 0, st>>>(dp_f, rows, T, c->ln_g[l], c->ln_b[l], nullptr, nullptr,
                                                     c->head_w, c->head_b, lens, mode,
                                                     c->cfg.duration_offset, out);
    }
    CUDA_TRY(cudaGetLastError());
    c->launches++;
    ahi = dp_hi; alo = dp_lo;
  }
  return EFTS_OK;
}

// Gaussian reconstruction + expansion (models/efficient_tts.py:184-194 / :270-280):
// e [B,T1] -> reconst_alpha [B,T1,T2] and expanded value at frame rate (fp32 + planes).
int run_reconstruct_expand(efts_ctx* c, cudaStream_t st, const float* e, const int* tl, const Skip* mel, int B,
                           int T1, int T2, int T1p, __half* R_hi, __half* R_lo, const __half* valT_hi,
                           const __half* valT_lo, float* reconst_alpha, float* out_f, __half* out_hi,
                           __half* out_lo) {
  const int C = c->cfg.n_channels;
  const int* sl = mel != nullptr ? mel->lens : nullptr;
  {
    ProfScope ps(c, st, TAG_RECONSTRUCT);
    const float neg_sigma = -1.0f * c->cfg.sigma;
    TRY(launch_reconstruct(st, e, tl, sl, B, T1, T2, T1p, neg_sigma, reconst_alpha, R_hi, R_lo));
    c->launches++;
  }
  GemmParams p = gemm_defaults();
  p.N = C;
  p.b_batched = 1;
  p.lens = sl;
  p.out = out_f; p.ld_out = C;
  p.out_hi = out_hi; p.out_lo = out_lo; p.ld_pl = C;
  if (mel != nullptr && c->skip_pad_tiles) {
    p.skip_lens = mel->lens; p.tile_list = mel->list; p.tile_count = mel->count;
    p.skip_halo = ((c->cfg.k_size - 1) / 2) * c->cfg.n_decoder_layer;
  }
  ProfScope ps(c, st, TAG_EXPAND);
  return launch_gemm(c, st, OpA{R_hi, R_lo, B, T2, T1p, T1p}, OpB{valT_hi, valT_lo, B, C, T1p, T1p}, p);
}

// Energy -> softmax/expectation -> scan -> aligned positions (models/efficient_tts.py:167-178).
int run_imv(efts_ctx* c, cudaStream_t st, const __half* q_hi, const __half* q_lo, const __half* key_hi,
            const __half* key_lo, const int* tl, const Skip* mel, int B, int T1, int T2, int T1p, float* S,
            float* imv_raw, float* imv, float* e) {
  const int C = c->cfg.n_channels;
  const int* sl = mel->lens;
  GemmParams p = gemm_defaults();
  p.N = T1p;
  p.b_batched = 1;
  p.divisor = static_cast<float>(std::sqrt(static_cast<double>(C)));   // np.sqrt(float(D)), :390
  // the token softmax runs in the GEMM epilogue: only (max, sum, weighted sum) per column tile reach memory (the
  // buffer S holds them); neither the scores nor alpha exist.
  // the wide (16-warp) variant emits one partial per 64-column half tile, the narrow one per 128-column tile
  const bool wide_softmax = c->wide && ((C + G2_BK - 1) / G2_BK) * (G2_BK / 16) <= 40;
  const int n_part = ((T1p + G2_BN - 1) / G2_BN) * (wide_softmax ? 2 : 1);
  p.softmax_part = reinterpret_cast<float4*>(S);
  p.col_lens = tl;
  if (c->skip_pad_tiles) { p.skip_lens = mel->lens; p.tile_list = mel->list; p.tile_count = mel->count; p.skip_halo = 0; }
  {
    ProfScope ps(c, st, TAG_ENERGY);
    TRY(launch_gemm(c, st, OpA{q_hi, q_lo, B, T2, C, C}, OpB{key_hi, key_lo, B, T1, C, C}, p));
  }
  {
    ProfScope ps(c, st, TAG_SCAN);
    TRY(launch_imv_scan(c->imv_version, st, imv_raw, reinterpret_cast<const float4*>(S), n_part, tl, sl,
                        B, T2, imv));
  }
  {
    ProfScope ps(c, st, TAG_ALIGNED);
    TRY(launch_aligned_positions(c->imv_version, st, imv, tl, sl, B, T1, T2, c->cfg.sigma_e, e, nullptr));
  }
  c->launches += 3;
  return EFTS_OK;
}


// ------------------------------------------------------------------------------------------------
// Resident layer stack (stack_sm100.cuh): B = 1 synthesis as two kernels instead of ~32 launches.
struct StackBuilder {
  efts_ctx* c;
  StackParams sp;
  int max_items = 1, max_items64 = 1;
  size_t max_plane = 0;
  int barriers = 0;
  const PackedW* wts[ST_MAX_LAYERS] = {};
  explicit StackBuilder(efts_ctx* c_) : c(c_) { memset(&sp, 0, sizeof(sp)); }

  // one tensor-core layer: A planes [T, K] x weights w -> epilogue described by the caller through the returned slot
  int add(const __half* a_hi, const __half* a_lo, int T, int K, const PackedW& w, int chunk_kb, int tag, StackLayer** out) {
    if (sp.n_layers >= ST_MAX_LAYERS) return fail(EFTS_ERR_UNSUPPORTED, "layer stack holds at most %d layers", ST_MAX_LAYERS);
    if (w.K != K) return fail(EFTS_ERR_ARG, "stack layer K mismatch %d vs %d", w.K, K);
    StackMaps& M = sp.maps[sp.n_layers];
    StackLayer& L = sp.layer[sp.n_layers++];
    TRY(make_map(c, &M.a_hi, a_hi, K, T, 1, K, G2_A_ROWS));
    TRY(make_map(c, &M.a_lo, a_lo, K, T, 1, K, G2_A_ROWS));
    wts[sp.n_layers - 1] = &w;                         // weight maps are encoded at launch (their box depends on the tile)
    L.T = T; L.K = K; L.N = w.N; L.ntaps = w.Z; L.pad = (w.Z - 1) / 2;
    L.chunk_kb = chunk_kb;
    L.bias = w.bias;
    L.err_code = 1 << (8 + tag);
    const int num_kb = (K + G2_BK - 1) / G2_BK;
    const int ckb = chunk_kb < 1 ? num_kb : chunk_kb;
    const int nchunks = (num_kb + ckb - 1) / ckb;
    max_items = std::max(max_items, ((T + G2_BM - 1) / G2_BM) * ((w.N + G2_BN - 1) / G2_BN) * nchunks);
    max_items64 = std::max(max_items64, ((T + G2_BM - 1) / G2_BM) * ((w.N + 63) / 64) * nchunks);
    max_plane = std::max(max_plane, static_cast<size_t>(T) * w.N);
    sp.split_stride = std::max(sp.split_stride, static_cast<size_t>(T) * w.N);
    if (nchunks > 8) return fail(EFTS_ERR_UNSUPPORTED, "layer stack: at most 8 accumulation chunks per layer");
    if (static_cast<size_t>(T) * w.N * nchunks * sizeof(float) > kSplitScratchBytes)
      return fail(EFTS_ERR_WORKSPACE, "layer stack: %d rows do not fit the partial-plane scratch", T);
    *out = &L;
    return EFTS_OK;
  }

  int launch(cudaStream_t st, float* scratch) {
    sp.scratch = scratch;
    sp.sync = c->sync_ctr;
    sp.sync_base = c->sync_base;
    sp.err_flag = c->err_flag;
    sp.trace = c->stack_trace_on ? c->stack_trace : nullptr;
    if (sp.text != nullptr) barriers += 1;
    // 64-column tiles when they still fit one wave: twice the CTAs, half the MMA depth per layer
    sp.bn = max_items64 <= c->sm_count ? 64 : G2_BN;
    if (sp.bn == 64) max_items = max_items64;
    for (int i = 0; i < sp.n_layers; ++i) {
      TRY(make_map(c, &sp.maps[i].b_hi, wts[i]->hi, wts[i]->K, wts[i]->N, wts[i]->Z, wts[i]->K, sp.bn));
      TRY(make_map(c, &sp.maps[i].b_lo, wts[i]->lo, wts[i]->K, wts[i]->N, wts[i]->Z, wts[i]->K, sp.bn));
    }
    // enough CTAs for the widest layer's work items and one (row, four columns) unit per thread of the largest
    // reduce, never more than SMs (every CTA must be resident: the layers synchronise through a grid barrier)
    size_t units = static_cast<size_t>(sp.T_embed) * 32;
    for (int i = 0; i < sp.n_layers; ++i)
      units = std::max(units, static_cast<size_t>(sp.layer[i].T) * (sp.layer[i].mode == ST_PLAIN ? sp.layer[i].N / 4 : 32));
    const int grid = std::max(1, std::min(c->sm_count, std::max(max_items, static_cast<int>((units + ST_THREADS - 1) / ST_THREADS))));
    cudaLaunchConfig_t cfg = {};
    cfg.blockDim = dim3(ST_THREADS);
    cfg.dynamicSmemBytes = ST_SMEM_BYTES;
    cfg.stream = st;
    barriers += 2 * sp.n_layers;
    cfg.gridDim = dim3(static_cast<unsigned>(grid));
    CUDA_TRY(cudaLaunchKernelEx(&cfg, stack_kernel, sp));
    c->sync_base += static_cast<unsigned>(grid) * static_cast<unsigned>(barriers);
    c->launches++;
    return EFTS_OK;
  }
};

// Whether the B = 1 schedules may take the resident stack: default numerics (fused-B pairs, flushed accumulator)
// and sizes the partial-plane scratch holds; anything else runs one launch per layer.
bool stack_usable(const efts_ctx* c, int T, int n_layers) {
  const int nchunks = c->chunk_kb < 1 ? 1 : (c->cfg.n_channels / G2_BK + c->chunk_kb - 1) / c->chunk_kb;
  return c->stack && c->pair && c->fuse_b && c->wide && c->split_k && c->chunk_kb >= 1 &&
         n_layers <= ST_MAX_LAYERS && c->cfg.n_channels == 512 && c->profile_mask == 0 &&
         static_cast<size_t>(T) * c->cfg.n_channels * nchunks * sizeof(float) <= kSplitScratchBytes;
}

int check_ready(const efts_ctx* c) {
  if (c == nullptr) return fail(EFTS_ERR_ARG, "null context");
  if (!c->finalized) return fail(EFTS_ERR_STATE, "weights not finalised (call efts_finalize_weights)");
  return EFTS_OK;
}

int split_planes(efts_ctx* c, cudaStream_t st, const float* x, size_t n, __half* hi, __half* lo) {
  if (n % 4 != 0) return fail(EFTS_ERR_ARG, "split_planes: element count must be a multiple of 4");
  const size_t n4 = n / 4;
  const unsigned grid = static_cast<unsigned>(std::min<size_t>((n4 + 255) / 256, 148 * 16));
  split_planes_kernel<<<grid ? grid : 1, 256, 0, st>>>(x, n4, hi, lo, c->err_flag);
  CUDA_TRY(cudaGetLastError());
  c->launches++;
  return EFTS_OK;
}

}  // namespace

// ================================================================================================
namespace { int create_base(int device, efts_ctx** out); }

extern "C" {

const char* efts_last_error(void) { return g_err; }
#ifndef EFTS_SOURCE_SHA
#define EFTS_SOURCE_SHA "unknown"
#endif
const char* efts_version(void) { return "efts_b200 0.2 (sm_100a, split-fp16 tcgen05) src " EFTS_SOURCE_SHA; }

int efts_create(const efts_config* cfg, efts_ctx** out) {
  if (cfg == nullptr || out == nullptr) return fail(EFTS_ERR_ARG, "null argument");
  if (cfg->n_channels != 512) return fail(EFTS_ERR_UNSUPPORTED, "n_channels=%d (kernels are built for 512)", cfg->n_channels);
  if (cfg->k_size != 5 && cfg->k_size != 3 && cfg->k_size != 1)
    return fail(EFTS_ERR_UNSUPPORTED, "k_size=%d (supported: 1, 3, 5)", cfg->k_size);
  if (cfg->duration_kernel_size != 3 && cfg->duration_kernel_size != 1 && cfg->duration_kernel_size != 5)
    return fail(EFTS_ERR_UNSUPPORTED, "duration_kernel_size=%d", cfg->duration_kernel_size);
  if (cfg->odim % 8 != 0 || cfg->odim <= 0 || cfg->odim > 256) return fail(EFTS_ERR_UNSUPPORTED, "odim=%d must be a multiple of 8, <= 256", cfg->odim);
  if (cfg->n_text_encoder_layer < 1 || cfg->n_text_encoder_layer > 16 || cfg->n_mel_encoder_layer < 1 ||
      cfg->n_mel_encoder_layer > 16 || cfg->n_decoder_layer < 1 || cfg->n_decoder_layer > 16 ||
      cfg->n_duration_layer < 1 || cfg->n_duration_layer > 4)
    return fail(EFTS_ERR_UNSUPPORTED, "layer counts out of range");
  if (cfg->leaky_relu_slope != 0.1f) return fail(EFTS_ERR_UNSUPPORTED, "leaky_relu_slope must be 0.1");
  efts_ctx* c = nullptr;
  TRY(create_base(cfg->device, &c));
  c->cfg = *cfg;
  *out = c;
  return EFTS_OK;
}

}  // extern "C"

namespace {
// Device checks, kernel attributes, the tensor-map encoder and the error word shared by both context kinds.
int create_base(int device, efts_ctx** out) {
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0)
    return fail(EFTS_ERR_CUDA, "no CUDA device: %s (this library has no CPU path)", cudaGetErrorString(e));
  if (device < 0 || device >= ndev) return fail(EFTS_ERR_ARG, "device %d out of range", device);
  cudaDeviceProp prop;
  CUDA_TRY(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10)
    return fail(EFTS_ERR_UNSUPPORTED, "device %d is sm_%d%d; this library is built for sm_100a only", device,
                prop.major, prop.minor);
  CUDA_TRY(cudaSetDevice(device));
  TRY(set_kernel_attributes());
  efts_ctx* c = new efts_ctx();
  c->sm_count = prop.multiProcessorCount;
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
  if (e != cudaSuccess || fn == nullptr || qres != cudaDriverEntryPointSuccess) {
    delete c;
    return fail(EFTS_ERR_CUDA, "cuTensorMapEncodeTiled unavailable: %s", cudaGetErrorString(e));
  }
  c->encode = reinterpret_cast<EncodeTiledFn>(fn);
  if (cudaMalloc(reinterpret_cast<void**>(&c->err_flag), sizeof(int)) != cudaSuccess ||
      cudaMemset(c->err_flag, 0, sizeof(int)) != cudaSuccess) {
    delete c;
    return fail(EFTS_ERR_CUDA, "cudaMalloc failed");
  }
  c->device_allocs.push_back(c->err_flag);
  if (cudaMalloc(reinterpret_cast<void**>(&c->sync_ctr), sizeof(unsigned)) != cudaSuccess ||
      cudaMemset(c->sync_ctr, 0, sizeof(unsigned)) != cudaSuccess) {
    delete c;
    return fail(EFTS_ERR_CUDA, "cudaMalloc failed");
  }
  c->device_allocs.push_back(c->sync_ctr);
  if (cudaMalloc(reinterpret_cast<void**>(&c->stack_trace), 64 * sizeof(long long)) != cudaSuccess) {
    delete c;
    return fail(EFTS_ERR_CUDA, "cudaMalloc failed");
  }
  c->device_allocs.push_back(c->stack_trace);
  *out = c;
  return EFTS_OK;
}
}  // namespace

extern "C" {

void efts_destroy(efts_ctx* c) {
  if (c == nullptr) return;
  for (void* p : c->device_allocs) cudaFree(p);
  for (auto& r : c->prof) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
  if (c->pinned_words != nullptr) cudaFreeHost(c->pinned_words);
  delete c->voc;
  delete c->fe;
  delete c;
}

int efts_set_weight(efts_ctx* c, const char* name, const float* data_host, const int64_t* shape, int32_t ndim) {
  if (c == nullptr || name == nullptr || data_host == nullptr || shape == nullptr || ndim < 1 || ndim > 4)
    return fail(EFTS_ERR_ARG, "bad argument to efts_set_weight");
  if (c->finalized) return fail(EFTS_ERR_STATE, "weights already finalised");
  size_t n = 1;
  std::vector<int64_t> sh(shape, shape + ndim);
  for (auto v : sh) n *= static_cast<size_t>(v);
  c->raw[name].assign(data_host, data_host + n);
  c->raw_shape[name] = sh;
  return EFTS_OK;
}

int efts_finalize_weights(efts_ctx* c) {
  if (c == nullptr) return fail(EFTS_ERR_ARG, "null context");
  if (c->finalized) return fail(EFTS_ERR_STATE, "weights already finalised");
  CUDA_TRY(cudaSetDevice(c->cfg.device));
  const efts_config& g = c->cfg;
  const int C = g.n_channels;
  TRY(upload_vec(c, "text_embedding_table.weight", {g.num_symbols, C}, &c->emb));
  struct { const char* name; int n; PackedW* dst; } stacks[3] = {
      {"text_encoder", g.n_text_encoder_layer, c->text},
      {"mel_encoder", g.n_mel_encoder_layer, c->mel},
      {"decoder", g.n_decoder_layer, c->dec}};
  for (auto& s : stacks)
    for (int i = 0; i < s.n; ++i) {
      const std::string p = std::string(s.name) + ".layers." + std::to_string(i) + ".conv.0.";
      TRY(pack_weight(c, p + "weight", p + "bias", C, C, g.k_size, &s.dst[i]));
    }
  TRY(pack_weight(c, "text_encoder_key.weight", "text_encoder_key.bias", C, C, 1, &c->key));
  TRY(pack_weight(c, "text_encoder_value.weight", "text_encoder_value.bias", C, C, 1, &c->value));
  TRY(pack_weight(c, "mel_prenet.0.weight", "mel_prenet.0.bias", C, g.odim, 1, &c->prenet));
  TRY(pack_weight(c, "mel_output_layer.weight", "mel_output_layer.bias", g.odim, C, 1, &c->melout));
  for (int i = 0; i < g.n_duration_layer; ++i) {
    const std::string p = "duration_predictor.conv." + std::to_string(i);
    TRY(pack_weight(c, p + ".0.weight", p + ".0.bias", C, C, g.duration_kernel_size, &c->dp[i]));
    TRY(upload_vec(c, p + ".2.weight", {C}, &c->ln_g[i]));
    TRY(upload_vec(c, p + ".2.bias", {C}, &c->ln_b[i]));
  }
  TRY(upload_vec(c, "duration_predictor.linear.weight", {1, C}, &c->head_w));
  TRY(upload_vec(c, "duration_predictor.linear.bias", {1}, &c->head_b));
  c->raw.clear();
  c->raw_shape.clear();
  c->finalized = true;
  return EFTS_OK;
}

size_t efts_workspace_bytes(const efts_ctx* c, int32_t B, int32_t T1, int32_t T2) {
  if (c == nullptr || B < 1 || T1 < 1 || T2 < 1) return 0;
  Arena a(nullptr, ~static_cast<size_t>(0));
  FwdWs w;
  carve_text(a, w, B, T1, c->cfg.n_channels);
  carve_mel(a, w, B, T2, c->cfg.n_channels, c->cfg.odim, true);
  // efts_alignment_fwd additionally stages key / value / query planes (3 text-sized, 1 mel-sized)
  const size_t extra = (static_cast<size_t>(B) * T1 * 3 + static_cast<size_t>(B) * T2) * c->cfg.n_channels * 4 + 4096;
  return a.off + extra + 4096;
}

int efts_set_option(efts_ctx* c, const char* name, int32_t value) {
  if (c == nullptr || name == nullptr) return fail(EFTS_ERR_ARG, "null argument");
  if (strcmp(name, "skip_pad_tiles") == 0) { c->skip_pad_tiles = value != 0; return EFTS_OK; }
  if (strcmp(name, "pair") == 0) { c->pair = value != 0; return EFTS_OK; }
  if (strcmp(name, "stack_trace") == 0) { c->stack_trace_on = value != 0; return EFTS_OK; }
  if (strcmp(name, "wide") == 0) { c->wide = value != 0; return EFTS_OK; }
  if (strcmp(name, "fuse_b") == 0) { c->fuse_b = value != 0; return EFTS_OK; }
  if (strcmp(name, "split_k") == 0) { c->split_k = value != 0; return EFTS_OK; }
  if (strcmp(name, "stack") == 0) { c->stack = value != 0; return EFTS_OK; }
  if (strcmp(name, "pdl") == 0) { c->pdl = value != 0; return EFTS_OK; }
  if (strcmp(name, "voc_group") == 0) { c->voc_group = value; return EFTS_OK; }
  if (strcmp(name, "voc_narrow") == 0) { c->voc_narrow = value != 0; return EFTS_OK; }
  if (strcmp(name, "voc_wide") == 0) { c->voc_wide = value; return EFTS_OK; }
  if (strcmp(name, "imv_version") == 0) {
    if (value != 1 && value != 2) return fail(EFTS_ERR_ARG, "imv_version must be 1 or 2");
    c->imv_version = value;
    return EFTS_OK;
  }
  if (strcmp(name, "chunk_kb") == 0) {
    if (value < 0 || value > 64) return fail(EFTS_ERR_ARG, "chunk_kb out of range");
    c->chunk_kb = value;
    return EFTS_OK;
  }
  return fail(EFTS_ERR_ARG, "unknown option '%s'", name);
}

int64_t efts_launch_count(const efts_ctx* c) { return c ? c->launches : 0; }

int efts_error_flags(efts_ctx* c, void* stream, int32_t* flags_host) {
  if (c == nullptr || flags_host == nullptr) return fail(EFTS_ERR_ARG, "null argument");
  return efts_read_words(c, c->err_flag, flags_host, 1, stream);
}

int efts_profile_enable(efts_ctx* c, uint32_t tag_mask) {
  if (c == nullptr) return fail(EFTS_ERR_ARG, "null context");
  c->profile_mask = tag_mask;
  c->prof_used = 0;
  return EFTS_OK;
}

int efts_profile_read(efts_ctx* c, int32_t tag, double* total_ms, int64_t* count) {
  if (c == nullptr || total_ms == nullptr || count == nullptr) return fail(EFTS_ERR_ARG, "null argument");
  double ms = 0.0;
  int64_t n = 0;
  for (size_t i = 0; i < c->prof_used; ++i) {
    if (c->prof[i].tag != tag) continue;
    CUDA_TRY(cudaEventSynchronize(c->prof[i].b));
    float t = 0.0f;
    CUDA_TRY(cudaEventElapsedTime(&t, c->prof[i].a, c->prof[i].b));
    ms += t;
    ++n;
  }
  *total_ms = ms;
  *count = n;
  return EFTS_OK;
}

int efts_read_words(efts_ctx* c, const int32_t* dev, int32_t* host, int32_t n, void* stream) {
  if (c == nullptr || dev == nullptr || host == nullptr || n < 1 || n > 64) return fail(EFTS_ERR_ARG, "bad argument");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (c->pinned_words == nullptr) CUDA_TRY(cudaMallocHost(reinterpret_cast<void**>(&c->pinned_words), 64 * sizeof(int32_t)));
  CUDA_TRY(cudaMemcpyAsync(c->pinned_words, dev, static_cast<size_t>(n) * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaStreamSynchronize(st));
  memcpy(host, c->pinned_words, static_cast<size_t>(n) * sizeof(int32_t));
  return EFTS_OK;
}

int efts_profile_stack_trace(efts_ctx* c, int64_t* out, int32_t n) {
  if (c == nullptr || out == nullptr || n < 1 || n > 64) return fail(EFTS_ERR_ARG, "bad argument");
  CUDA_TRY(cudaMemcpy(out, c->stack_trace, static_cast<size_t>(n) * sizeof(long long), cudaMemcpyDeviceToHost));
  return EFTS_OK;
}

int efts_profile_kernel_name(const efts_ctx* c, int32_t tag, char* buf, size_t n) {
  if (c == nullptr || buf == nullptr || n == 0 || tag < 0 || tag > 15) return fail(EFTS_ERR_ARG, "bad argument");
  snprintf(buf, n, "%s", c->tag_kernel[tag]);
  return EFTS_OK;
}

// ------------------------------------------------------------------------------------------------
int efts_forward(efts_ctx* c, const int64_t* text, const int64_t* text_lengths, const float* speech,
                 const int64_t* speech_lengths, int32_t B, int32_t T1, int32_t T2, float* imv,
                 float* reconst_alpha, float* mel_pred, float* scalars, void* workspace, size_t workspace_bytes,
                 void* stream) {
  TRY(check_ready(c));
  if (!text || !text_lengths || !speech || !speech_lengths || !imv || !reconst_alpha || !mel_pred || !scalars ||
      !workspace)
    return fail(EFTS_ERR_ARG, "efts_forward: null pointer");
  if (B < 1 || T1 < 1 || T2 < 1) return fail(EFTS_ERR_ARG, "efts_forward: bad sizes B=%d T1=%d T2=%d", B, T1, T2);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const efts_config& g = c->cfg;
  const int C = g.n_channels;
  Arena a(workspace, workspace_bytes);
  FwdWs w;
  carve_text(a, w, B, T1, C);
  carve_mel(a, w, B, T2, C, g.odim, true);
  if (!a.ok) return fail(EFTS_ERR_WORKSPACE, "workspace too small: need %zu, have %zu", a.off, workspace_bytes);
  const size_t m1 = static_cast<size_t>(B) * T1, m2 = static_cast<size_t>(B) * T2;

  // 0. lengths, flags, accumulators
  CUDA_TRY(cudaMemsetAsync(w.flags, 0, 4 * sizeof(int), st));
  CUDA_TRY(cudaMemsetAsync(w.acc, 0, 2 * sizeof(double), st));
  CUDA_TRY(cudaMemsetAsync(c->err_flag, 0, sizeof(int), st));
  prep_lengths_kernel<<<1, 256, 0, st>>>(text_lengths, speech_lengths, B, T1, T2, w.tl32, w.sl32, w.flags);
  CUDA_TRY(cudaGetLastError());
  // live row tiles of both sides (largest halo any launch uses; launches re-check their own halo)
  const int pad_k = (g.k_size - 1) / 2;
  build_tile_list_kernel<<<1, 256, 0, st>>>(w.tl32, B, T1, pad_k * g.n_text_encoder_layer, w.list_t, w.cnt_t);
  CUDA_TRY(cudaGetLastError());
  build_tile_list_kernel<<<1, 256, 0, st>>>(w.sl32, B, T2, pad_k * std::max(g.n_mel_encoder_layer, g.n_decoder_layer),
                                            w.list_m, w.cnt_m);
  CUDA_TRY(cudaGetLastError());
  c->launches += 2;
  const Skip skip_t{w.tl32, w.list_t, w.cnt_t};
  const Skip skip_m{w.sl32, w.list_m, w.cnt_m};
  // 1. embedding (:144) -> text encoder (:148)
  {
    ProfScope ps(c, st, TAG_PREP);
    embed_kernel<<<static_cast<unsigned>(m1), C / 4, 0, st>>>(text, c->emb, g.num_symbols, C, w.xt_f[0], w.xt_hi[0],
                                                              w.xt_lo[0], w.flags, nullptr, T1);
    CUDA_TRY(cudaGetLastError());
  }
  c->launches += 2;
  int cur = 0;
  TRY(run_conv_stack(c, st, c->text, g.n_text_encoder_layer, B, T1, w.xt_f, w.xt_hi, w.xt_lo, nullptr, nullptr,
                     &skip_t, &cur, TAG_TEXT_CONV));
  // 2. key / value projections, zero at pad tokens (:149-157)
  {
    GemmParams p = gemm_defaults();
    p.N = C; p.bias = c->key.bias; p.lens = w.tl32;
    p.out_hi = w.key_hi; p.out_lo = w.key_lo; p.ld_pl = C;
    { ProfScope ps(c, st, TAG_LINEAR); TRY(launch_gemm(c, st, OpA{w.xt_hi[cur], w.xt_lo[cur], B, T1, C, C}, weight_op(c->key), p)); }
    p.bias = c->value.bias;
    p.out_hi = w.val_hi; p.out_lo = w.val_lo;
    p.outT_hi = w.valT_hi; p.outT_lo = w.valT_lo; p.ld_t = w.T1p;
    if (w.T1p != T1) {   // K padding of the expansion GEMM must be finite (it multiplies zeros)
      CUDA_TRY(cudaMemsetAsync(w.valT_hi, 0, static_cast<size_t>(B) * C * w.T1p * sizeof(__half), st));
      CUDA_TRY(cudaMemsetAsync(w.valT_lo, 0, static_cast<size_t>(B) * C * w.T1p * sizeof(__half), st));
    }
    { ProfScope ps(c, st, TAG_LINEAR); TRY(launch_gemm(c, st, OpA{w.xt_hi[cur], w.xt_lo[cur], B, T1, C, C}, weight_op(c->value), p)); }
  }
  // 3. mel prenet (:161) -> mel encoder (:162)
  { ProfScope ps(c, st, TAG_PREP); TRY(split_planes(c, st, speech, m2 * g.odim, w.sp_hi, w.sp_lo)); }
  {
    GemmParams p = gemm_defaults();
    p.N = C; p.bias = c->prenet.bias; p.act = ACT_LRELU;
    p.out = w.xm_f[0]; p.ld_out = C;
    p.out_hi = w.xm_hi[0]; p.out_lo = w.xm_lo[0]; p.ld_pl = C;
    if (c->skip_pad_tiles) {
      p.skip_lens = skip_m.lens; p.tile_list = skip_m.list; p.tile_count = skip_m.count;
      p.skip_halo = pad_k * g.n_mel_encoder_layer;
    }
    { ProfScope ps(c, st, TAG_LINEAR); TRY(launch_gemm(c, st, OpA{w.sp_hi, w.sp_lo, B, T2, g.odim, g.odim}, weight_op(c->prenet), p)); }
  }
  int curm = 0;
  TRY(run_conv_stack(c, st, c->mel, g.n_mel_encoder_layer, B, T2, w.xm_f, w.xm_hi, w.xm_lo, nullptr, nullptr,
                     &skip_m, &curm, TAG_MEL_CONV));
  // 4. alignment: energy/softmax/expectation, scan, aligned positions (:167-178)
  TRY(run_imv(c, st, w.xm_hi[curm], w.xm_lo[curm], w.key_hi, w.key_lo, w.tl32, &skip_m, B, T1, T2, w.T1p, w.S,
              w.imv_raw, imv, w.e));
  // 5. Gaussian reconstruction + expansion (:184-194) into mel buffer set 0
  TRY(run_reconstruct_expand(c, st, w.e, w.tl32, &skip_m, B, T1, T2, w.T1p, w.R_hi, w.R_lo, w.valT_hi, w.valT_lo,
                             reconst_alpha, w.xm_f[0], w.xm_hi[0], w.xm_lo[0]));
  // 6. decoder (:197) and mel head (:198-200)
  curm = 0;
  TRY(run_conv_stack(c, st, c->dec, g.n_decoder_layer, B, T2, w.xm_f, w.xm_hi, w.xm_lo, nullptr, nullptr, &skip_m,
                     &curm, TAG_DEC_CONV));
  {
    GemmParams p = gemm_defaults();
    p.N = g.odim; p.bias = c->melout.bias; p.lens = w.sl32;
    p.out = mel_pred; p.ld_out = g.odim;
    { ProfScope ps(c, st, TAG_LINEAR); TRY(launch_gemm(c, st, OpA{w.xm_hi[curm], w.xm_lo[curm], B, T2, C, C}, weight_op(c->melout), p)); }
  }
  // 7. duration predictor on value (:219), log domain, zero at pad tokens
  TRY(run_duration_predictor(c, st, w.val_hi, w.val_lo, B, T1, w.dp_f, w.dp_hi, w.dp_lo, w.tl32, 0, w.dur, false,
                             &skip_t));
  // 8. losses (:220-227)
  ProfScope ps_loss(c, st, TAG_LOSS);
  loss_partial_kernel<<<c->sm_count * 4, 256, 0, st>>>(mel_pred, speech, w.sl32, T2, g.odim, w.dur, w.e, w.tl32,
                                                       T1, B, g.duration_offset, g.use_masking, w.acc);
  CUDA_TRY(cudaGetLastError());
  loss_finalize_kernel<<<1, 32, 0, st>>>(w.acc, w.tl32, w.sl32, B, T1, T2, g.odim, g.use_masking, w.flags, c->err_flag,
                                         scalars);
  CUDA_TRY(cudaGetLastError());
  c->launches += 2;
  return EFTS_OK;
}

// ------------------------------------------------------------------------------------------------
int efts_inference_phase1(efts_ctx* c, const int64_t* text, int32_t T1, int32_t* t2_dev, void* workspace,
                          size_t workspace_bytes, void* stream) {
  TRY(check_ready(c));
  if (!text || !t2_dev || !workspace || T1 < 1) return fail(EFTS_ERR_ARG, "efts_inference_phase1: bad argument");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const efts_config& g = c->cfg;
  const int C = g.n_channels;
  Arena a(workspace, workspace_bytes);
  FwdWs w;
  carve_text(a, w, 1, T1, C);
  if (!a.ok) return fail(EFTS_ERR_WORKSPACE, "workspace too small: need %zu, have %zu", a.off, workspace_bytes);
  CUDA_TRY(cudaMemsetAsync(t2_dev, 0, 2 * sizeof(int), st));
  CUDA_TRY(cudaMemsetAsync(c->err_flag, 0, sizeof(int), st));
  const int nl = g.n_duration_layer;
  if (stack_usable(c, T1, g.n_text_encoder_layer + 1 + nl)) {
    // one resident kernel: embedding, text encoder, value projection, duration predictor, cumsum / T2
    StackBuilder sb(c);
    sb.sp.text = text; sb.sp.emb = c->emb; sb.sp.num_symbols = g.num_symbols; sb.sp.T_embed = T1;
    sb.sp.x0_f = w.xt_f[0]; sb.sp.x0_hi = w.xt_hi[0]; sb.sp.x0_lo = w.xt_lo[0]; sb.sp.flags = t2_dev + 1;
    int cur = 0;
    StackLayer* L;
    for (int l = 0; l < g.n_text_encoder_layer; ++l, cur ^= 1) {
      TRY(sb.add(w.xt_hi[cur], w.xt_lo[cur], T1, C, c->text[l], c->chunk_kb, TAG_TEXT_CONV, &L));
      L->act = ACT_LRELU; L->resid = w.xt_f[cur]; L->out = l == g.n_text_encoder_layer - 1 ? nullptr : w.xt_f[cur ^ 1];
      L->out_hi = w.xt_hi[cur ^ 1]; L->out_lo = w.xt_lo[cur ^ 1];
    }
    // value only: the key projection at :251 is computed by the reference but never used
    TRY(sb.add(w.xt_hi[cur], w.xt_lo[cur], T1, C, c->value, 0, TAG_LINEAR, &L));
    L->out_hi = w.val_hi; L->out_lo = w.val_lo; L->outT_hi = w.valT_hi; L->outT_lo = w.valT_lo; L->ld_t = w.T1p;
    if (w.T1p != T1) {
      CUDA_TRY(cudaMemsetAsync(w.valT_hi, 0, static_cast<size_t>(C) * w.T1p * sizeof(__half), st));
      CUDA_TRY(cudaMemsetAsync(w.valT_lo, 0, static_cast<size_t>(C) * w.T1p * sizeof(__half), st));
    }
    // durations clamp(exp(x) - offset, 0) (:258): [conv k3 -> ReLU -> LayerNorm] x n, Linear head
    const __half* ahi = w.val_hi;
    const __half* alo = w.val_lo;
    for (int l = 0; l < nl; ++l) {
      TRY(sb.add(ahi, alo, T1, C, c->dp[l], c->chunk_kb, TAG_DURATION, &L));
      L->act = ACT_RELU; L->ln_g = c->ln_g[l]; L->ln_b = c->ln_b[l];
      if (l < nl - 1) {
        L->mode = ST_LN_PLANES; L->out_hi = w.dp_hi; L->out_lo = w.dp_lo;
      } else {
        L->mode = ST_LN_HEAD; L->head_w = c->head_w; L->head_b = c->head_b; L->head_mode = 1;
        L->head_offset = g.duration_offset; L->head_out = w.dur;
      }
      ahi = w.dp_hi; alo = w.dp_lo;
    }
    // their cumsum (:260); T2 = round(e[-1]) (:361)
    sb.sp.dur = w.dur; sb.sp.e = w.e; sb.sp.t2_out = t2_dev; sb.sp.T_cumsum = T1;
    return sb.launch(st, w.splitk);
  }
  embed_kernel<<<T1, C / 4, 0, st>>>(text, c->emb, g.num_symbols, C, w.xt_f[0], w.xt_hi[0], w.xt_lo[0], t2_dev + 1,
                                     nullptr, T1);
  CUDA_TRY(cudaGetLastError());
  c->launches++;
  int cur = 0;
  TRY(run_conv_stack(c, st, c->text, g.n_text_encoder_layer, 1, T1, w.xt_f, w.xt_hi, w.xt_lo, nullptr, nullptr,
                     nullptr, &cur, TAG_TEXT_CONV, false, w.splitk));
  {   // value only: the key projection at :251 is computed by the reference but never used
    GemmParams p = gemm_defaults();
    p.N = C; p.bias = c->value.bias;
    p.out_hi = w.val_hi; p.out_lo = w.val_lo; p.ld_pl = C;
    p.outT_hi = w.valT_hi; p.outT_lo = w.valT_lo; p.ld_t = w.T1p;
    if (w.T1p != T1) {
      CUDA_TRY(cudaMemsetAsync(w.valT_hi, 0, static_cast<size_t>(C) * w.T1p * sizeof(__half), st));
      CUDA_TRY(cudaMemsetAsync(w.valT_lo, 0, static_cast<size_t>(C) * w.T1p * sizeof(__half), st));
    }
    TRY(launch_gemm(c, st, OpA{w.xt_hi[cur], w.xt_lo[cur], 1, T1, C, C}, weight_op(c->value), p));
  }
  // durations clamp(exp(x) - offset, 0) (:258) and their cumsum (:260); T2 = round(e[-1]) (:361)
  TRY(run_duration_predictor(c, st, w.val_hi, w.val_lo, 1, T1, w.dp_f, w.dp_hi, w.dp_lo, nullptr, 1, w.dur, false,
                             nullptr, w.splitk));
  duration_cumsum_kernel<<<1, 32, 0, st>>>(w.dur, T1, w.e, t2_dev);
  CUDA_TRY(cudaGetLastError());
  c->launches++;
  return EFTS_OK;
}

int efts_inference_phase2(efts_ctx* c, int32_t T1, int32_t T2, float* mel_pred, float* reconst_alpha,
                          void* workspace, size_t workspace_bytes, void* stream) {
  TRY(check_ready(c));
  if (!mel_pred || !reconst_alpha || !workspace || T1 < 1) return fail(EFTS_ERR_ARG, "efts_inference_phase2: bad argument");
  if (T2 < 1)
    return fail(EFTS_ERR_DATA, "predicted length T2=%d: the reference's decoder conv fails on an empty sequence", T2);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const efts_config& g = c->cfg;
  const int C = g.n_channels;
  Arena a(workspace, workspace_bytes);
  FwdWs w;
  carve_text(a, w, 1, T1, C);
  carve_mel(a, w, 1, T2, C, g.odim, false);
  if (!a.ok) return fail(EFTS_ERR_WORKSPACE, "workspace too small: need %zu, have %zu", a.off, workspace_bytes);
  TRY(run_reconstruct_expand(c, st, w.e, nullptr, nullptr, 1, T1, T2, w.T1p, w.R_hi, w.R_lo, w.valT_hi, w.valT_lo,
                             reconst_alpha, w.xm_f[0], w.xm_hi[0], w.xm_lo[0]));
  if (stack_usable(c, T2, g.n_decoder_layer + 1)) {
    // decoder (:282) and mel head (:283-284) as one resident kernel
    StackBuilder sb(c);
    int cur = 0;
    StackLayer* L;
    for (int l = 0; l < g.n_decoder_layer; ++l, cur ^= 1) {
      TRY(sb.add(w.xm_hi[cur], w.xm_lo[cur], T2, C, c->dec[l], c->chunk_kb, TAG_DEC_CONV, &L));
      L->act = ACT_LRELU; L->resid = w.xm_f[cur]; L->out = l == g.n_decoder_layer - 1 ? nullptr : w.xm_f[cur ^ 1];
      L->out_hi = w.xm_hi[cur ^ 1]; L->out_lo = w.xm_lo[cur ^ 1];
    }
    TRY(sb.add(w.xm_hi[cur], w.xm_lo[cur], T2, C, c->melout, 0, TAG_LINEAR, &L));
    L->out = mel_pred;
    return sb.launch(st, w.splitk);
  }
  int curm = 0;
  TRY(run_conv_stack(c, st, c->dec, g.n_decoder_layer, 1, T2, w.xm_f, w.xm_hi, w.xm_lo, nullptr, nullptr, nullptr,
                     &curm, TAG_DEC_CONV, false, w.splitk));
  GemmParams p = gemm_defaults();
  p.N = g.odim; p.bias = c->melout.bias;
  p.out = mel_pred; p.ld_out = g.odim;
  return launch_gemm(c, st, OpA{w.xm_hi[curm], w.xm_lo[curm], 1, T2, C, C}, weight_op(c->melout), p);
}

// ------------------------------------------------------------------------------------------------
// Ragged batched synthesis.  Every utterance is computed as if it were alone and unpadded:
// rows beyond its length are zero after every layer, exactly the zero padding the B = 1 path convolves against.
int efts_inference_batch_phase1(efts_ctx* c, const int64_t* text, const int64_t* text_lengths, int32_t B, int32_t T1,
                                int32_t* t2_dev, void* workspace, size_t workspace_bytes, void* stream) {
  TRY(check_ready(c));
  if (!text || !text_lengths || !t2_dev || !workspace || B < 1 || T1 < 1 || B > 65535)
    return fail(EFTS_ERR_ARG, "efts_inference_batch_phase1: bad argument");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const efts_config& g = c->cfg;
  const int C = g.n_channels;
  const int pad_k = (g.k_size - 1) / 2;
  Arena a(workspace, workspace_bytes);
  FwdWs w;
  carve_text(a, w, B, T1, C);
  if (!a.ok) return fail(EFTS_ERR_WORKSPACE, "workspace too small: need %zu, have %zu", a.off, workspace_bytes);
  const size_t m1 = static_cast<size_t>(B) * T1;
  CUDA_TRY(cudaMemsetAsync(t2_dev, 0, (B + 1) * sizeof(int), st));
  CUDA_TRY(cudaMemsetAsync(c->err_flag, 0, sizeof(int), st));
  prep_lengths_kernel<<<1, 256, 0, st>>>(text_lengths, nullptr, B, T1, 0, w.tl32, w.sl32, t2_dev + B);
  CUDA_TRY(cudaGetLastError());
  build_tile_list_kernel<<<1, 256, 0, st>>>(w.tl32, B, T1, pad_k, w.list_t, w.cnt_t);
  CUDA_TRY(cudaGetLastError());
  embed_kernel<<<static_cast<unsigned>(m1), C / 4, 0, st>>>(text, c->emb, g.num_symbols, C, w.xt_f[0], w.xt_hi[0],
                                                            w.xt_lo[0], t2_dev + B, w.tl32, T1);
  CUDA_TRY(cudaGetLastError());
  c->launches += 3;
  const Skip skip_t{w.tl32, w.list_t, w.cnt_t};
  int cur = 0;
  TRY(run_conv_stack(c, st, c->text, g.n_text_encoder_layer, B, T1, w.xt_f, w.xt_hi, w.xt_lo, nullptr, nullptr,
                     &skip_t, &cur, TAG_TEXT_CONV, true));
  {
    GemmParams p = gemm_defaults();
    p.N = C; p.bias = c->value.bias; p.lens = w.tl32;
    p.out_hi = w.val_hi; p.out_lo = w.val_lo; p.ld_pl = C;
    p.outT_hi = w.valT_hi; p.outT_lo = w.valT_lo; p.ld_t = w.T1p;
    if (w.T1p != T1) {
      CUDA_TRY(cudaMemsetAsync(w.valT_hi, 0, static_cast<size_t>(B) * C * w.T1p * sizeof(__half), st));
      CUDA_TRY(cudaMemsetAsync(w.valT_lo, 0, static_cast<size_t>(B) * C * w.T1p * sizeof(__half), st));
    }
    ProfScope ps(c, st, TAG_LINEAR);
    TRY(launch_gemm(c, st, OpA{w.xt_hi[cur], w.xt_lo[cur], B, T1, C, C}, weight_op(c->value), p));
  }
  TRY(run_duration_predictor(c, st, w.val_hi, w.val_lo, B, T1, w.dp_f, w.dp_hi, w.dp_lo, w.tl32, 1, w.dur, true));
  duration_cumsum_batch_kernel<<<(B + 3) / 4, 128, 0, st>>>(w.dur, w.tl32, B, T1, w.e, t2_dev);
  CUDA_TRY(cudaGetLastError());
  c->launches++;
  return EFTS_OK;
}

int efts_inference_batch_phase2(efts_ctx* c, int32_t B, int32_t T1, int32_t T2max, const int32_t* t2_dev,
                                float* mel_pred, float* reconst_alpha, void* workspace, size_t workspace_bytes,
                                void* stream) {
  TRY(check_ready(c));
  if (!t2_dev || !mel_pred || !reconst_alpha || !workspace || B < 1 || T1 < 1 || T2max < 1 || B > 65535)
    return fail(EFTS_ERR_ARG, "efts_inference_batch_phase2: bad argument");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const efts_config& g = c->cfg;
  const int C = g.n_channels;
  const int pad_k = (g.k_size - 1) / 2;
  Arena a(workspace, workspace_bytes);
  FwdWs w;
  carve_text(a, w, B, T1, C);
  carve_mel(a, w, B, T2max, C, g.odim, false);
  if (!a.ok) return fail(EFTS_ERR_WORKSPACE, "workspace too small: need %zu, have %zu", a.off, workspace_bytes);
  const int* t2 = t2_dev;                         // per-utterance frame counts are the mel-side lengths
  build_tile_list_kernel<<<1, 256, 0, st>>>(t2, B, T2max, pad_k, w.list_m, w.cnt_m);
  CUDA_TRY(cudaGetLastError());
  c->launches++;
  const Skip skip_m{t2, w.list_m, w.cnt_m};
  TRY(run_reconstruct_expand(c, st, w.e, w.tl32, &skip_m, B, T1, T2max, w.T1p, w.R_hi, w.R_lo, w.valT_hi,
                             w.valT_lo, reconst_alpha, w.xm_f[0], w.xm_hi[0], w.xm_lo[0]));
  int curm = 0;
  TRY(run_conv_stack(c, st, c->dec, g.n_decoder_layer, B, T2max, w.xm_f, w.xm_hi, w.xm_lo, nullptr, nullptr,
                     &skip_m, &curm, TAG_DEC_CONV, true));
  GemmParams p = gemm_defaults();
  p.N = g.odim; p.bias = c->melout.bias; p.lens = t2;
  p.out = mel_pred; p.ld_out = g.odim;
  ProfScope ps(c, st, TAG_LINEAR);
  return launch_gemm(c, st, OpA{w.xm_hi[curm], w.xm_lo[curm], B, T2max, C, C}, weight_op(c->melout), p);
}

// ------------------------------------------------------------------------------------------------
int efts_conv_stack_fwd(efts_ctx* c, int32_t stack, const float* x, float* y, int32_t B, int32_t T, void* workspace,
                        size_t workspace_bytes, void* stream) {
  TRY(check_ready(c));
  if (!x || !y || !workspace || B < 1 || T < 1) return fail(EFTS_ERR_ARG, "efts_conv_stack_fwd: bad argument");
  const PackedW* layers;
  int n;
  if (stack == 0) { layers = c->text; n = c->cfg.n_text_encoder_layer; }
  else if (stack == 1) { layers = c->mel; n = c->cfg.n_mel_encoder_layer; }
  else if (stack == 2) { layers = c->dec; n = c->cfg.n_decoder_layer; }
  else return fail(EFTS_ERR_ARG, "stack must be 0 (text), 1 (mel) or 2 (decoder)");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int C = c->cfg.n_channels;
  const size_t m = static_cast<size_t>(B) * T;
  Arena a(workspace, workspace_bytes);
  float* f[2]; __half* hi[2]; __half* lo[2];
  for (int i = 0; i < 2; ++i) {
    f[i] = a.get<float>(m * C);
    hi[i] = a.get<__half>(m * C);
    lo[i] = a.get<__half>(m * C);
  }
  if (!a.ok) return fail(EFTS_ERR_WORKSPACE, "workspace too small: need %zu, have %zu", a.off, workspace_bytes);
  TRY(split_planes(c, st, x, m * C, hi[0], lo[0]));
  int cur = 0;
  return run_conv_stack(c, st, layers, n, B, T, f, hi, lo, x, y, nullptr, &cur, stack);
}

int efts_duration_predictor_fwd(efts_ctx* c, const float* x, const int32_t* lengths, int32_t B, int32_t T,
                                int32_t mode, void* out, void* workspace, size_t workspace_bytes, void* stream) {
  TRY(check_ready(c));
  if (!x || !out || !workspace || B < 1 || T < 1 || mode < 0 || mode > 2)
    return fail(EFTS_ERR_ARG, "efts_duration_predictor_fwd: bad argument");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int C = c->cfg.n_channels;
  const size_t m = static_cast<size_t>(B) * T;
  Arena a(workspace, workspace_bytes);
  __half* in_hi = a.get<__half>(m * C);
  __half* in_lo = a.get<__half>(m * C);
  float* dp_f = a.get<float>(m * C);
  __half* dp_hi = a.get<__half>(m * C);
  __half* dp_lo = a.get<__half>(m * C);
  if (!a.ok) return fail(EFTS_ERR_WORKSPACE, "workspace too small: need %zu, have %zu", a.off, workspace_bytes);
  TRY(split_planes(c, st, x, m * C, in_hi, in_lo));
  return run_duration_predictor(c, st, in_hi, in_lo, B, T, dp_f, dp_hi, dp_lo, lengths, mode, out);
}

int efts_tap_gemm(efts_ctx* c, const float* x, const float* wgt, float* out, int32_t B, int32_t T, int32_t K,
                  int32_t N, int32_t ntaps, int32_t pad, int32_t batched, void* workspace, size_t workspace_bytes,
                  void* stream) {
  if (c == nullptr) return fail(EFTS_ERR_ARG, "null context");
  if (!x || !wgt || !out || !workspace || B < 1 || T < 1 || K < 8 || N < 8 || K % 8 || N % 8 || ntaps < 1)
    return fail(EFTS_ERR_ARG, "efts_tap_gemm: bad argument");
  if (batched && ntaps != 1) return fail(EFTS_ERR_ARG, "efts_tap_gemm: batched requires ntaps == 1");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int Z = batched ? B : ntaps;
  const size_t nx = static_cast<size_t>(B) * T * K, nw = static_cast<size_t>(Z) * N * K;
  Arena a(workspace, workspace_bytes);
  __half* xh = a.get<__half>(nx); __half* xl = a.get<__half>(nx);
  __half* wh = a.get<__half>(nw); __half* wl = a.get<__half>(nw);
  if (!a.ok) return fail(EFTS_ERR_WORKSPACE, "workspace too small: need %zu, have %zu", a.off, workspace_bytes);
  TRY(split_planes(c, st, x, nx, xh, xl));
  TRY(split_planes(c, st, wgt, nw, wh, wl));
  GemmParams p = gemm_defaults();
  p.N = N; p.ntaps = ntaps; p.pad = pad; p.b_batched = batched;
  p.out = out; p.ld_out = N;
  return launch_gemm(c, st, OpA{xh, xl, B, T, K, K}, OpB{wh, wl, Z, N, K, K}, p);
}

int efts_alignment_fwd(efts_ctx* c, const float* mel_h, const float* key, const float* value,
                       const int32_t* text_lengths, const int32_t* speech_lengths, int32_t B, int32_t T1, int32_t T2,
                       float* imv, float* e, float* reconst_alpha, float* expanded, void* workspace,
                       size_t workspace_bytes, void* stream) {
  if (c == nullptr) return fail(EFTS_ERR_ARG, "null context");
  if (!mel_h || !key || !value || !text_lengths || !speech_lengths || !imv || !e || !reconst_alpha || !expanded ||
      !workspace || B < 1 || T1 < 1 || T2 < 1)
    return fail(EFTS_ERR_ARG, "efts_alignment_fwd: bad argument");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int C = c->cfg.n_channels;
  const int T1p = round8(T1);
  const size_t m1 = static_cast<size_t>(B) * T1, m2 = static_cast<size_t>(B) * T2;
  Arena a(workspace, workspace_bytes);
  __half* q_hi = a.get<__half>(m2 * C); __half* q_lo = a.get<__half>(m2 * C);
  __half* k_hi = a.get<__half>(m1 * C); __half* k_lo = a.get<__half>(m1 * C);
  __half* vT_hi = a.get<__half>(static_cast<size_t>(B) * C * T1p);
  __half* vT_lo = a.get<__half>(static_cast<size_t>(B) * C * T1p);
  float* S = a.get<float>(m2 * T1p);
  float* imv_raw = a.get<float>(m2);
  __half* R_hi = a.get<__half>(m2 * T1p); __half* R_lo = a.get<__half>(m2 * T1p);
  __half* x_hi = a.get<__half>(m2 * C); __half* x_lo = a.get<__half>(m2 * C);
  if (!a.ok) return fail(EFTS_ERR_WORKSPACE, "workspace too small: need %zu, have %zu", a.off, workspace_bytes);
  TRY(split_planes(c, st, mel_h, m2 * C, q_hi, q_lo));
  TRY(split_planes(c, st, key, m1 * C, k_hi, k_lo));
  split_transpose_kernel<<<dim3((T1p + 31) / 32, C / 32, B), dim3(32, 8), 0, st>>>(value, T1, C, T1p, vT_hi, vT_lo);
  CUDA_TRY(cudaGetLastError());
  c->launches++;
  const Skip mel{speech_lengths, nullptr, nullptr};
  TRY(run_imv(c, st, q_hi, q_lo, k_hi, k_lo, text_lengths, &mel, B, T1, T2, T1p, S, imv_raw, imv, e));
  return run_reconstruct_expand(c, st, e, text_lengths, &mel, B, T1, T2, T1p, R_hi, R_lo, vT_hi, vT_lo,
                                reconst_alpha, expanded, x_hi, x_lo);
}

// ------------------------------------------------------------------------------------------------
// Stand-alone helper methods of EfficientTTSCNN (models/efficient_tts.py:287-398).
int efts_mask_lengths(const uint8_t* mask, int32_t B, int32_t T, int32_t* lengths, void* stream) {
  if (!mask || !lengths || B < 1 || T < 1) return fail(EFTS_ERR_ARG, "efts_mask_lengths: bad argument");
  mask_lengths_kernel<<<(B + 3) / 4, 128, 0, static_cast<cudaStream_t>(stream)>>>(mask, B, T, lengths);
  CUDA_TRY(cudaGetLastError());
  return EFTS_OK;
}

int efts_index_vector(const int32_t* text_lengths, int32_t B, int32_t T1, float* p, void* stream) {
  if (!text_lengths || !p || B < 1 || T1 < 1) return fail(EFTS_ERR_ARG, "efts_index_vector: bad argument");
  const size_t n = static_cast<size_t>(B) * T1;
  index_vector_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      text_lengths, B, T1, p);
  CUDA_TRY(cudaGetLastError());
  return EFTS_OK;
}

int efts_attention_alpha(efts_ctx* c, const float* query, const float* key, const int32_t* text_lengths, int32_t B,
                         int32_t T1, int32_t T2, float* alpha, void* workspace, size_t workspace_bytes, void* stream) {
  if (c == nullptr) return fail(EFTS_ERR_ARG, "null context");
  if (!query || !key || !text_lengths || !alpha || !workspace || B < 1 || T1 < 1 || T2 < 1 || B > 65535)
    return fail(EFTS_ERR_ARG, "efts_attention_alpha: bad argument");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int C = c->cfg.n_channels;
  const int T1p = round8(T1);
  const size_t m1 = static_cast<size_t>(B) * T1, m2 = static_cast<size_t>(B) * T2;
  Arena a(workspace, workspace_bytes);
  __half* q_hi = a.get<__half>(m2 * C); __half* q_lo = a.get<__half>(m2 * C);
  __half* k_hi = a.get<__half>(m1 * C); __half* k_lo = a.get<__half>(m1 * C);
  float* S = a.get<float>(m2 * T1p);
  if (!a.ok) return fail(EFTS_ERR_WORKSPACE, "workspace too small: need %zu, have %zu", a.off, workspace_bytes);
  TRY(split_planes(c, st, query, m2 * C, q_hi, q_lo));
  TRY(split_planes(c, st, key, m1 * C, k_hi, k_lo));
  GemmParams p = gemm_defaults();
  p.N = T1p;
  p.b_batched = 1;
  p.divisor = static_cast<float>(std::sqrt(static_cast<double>(C)));
  p.out = S; p.ld_out = T1p;
  TRY(launch_gemm(c, st, OpA{q_hi, q_lo, B, T2, C, C}, OpB{k_hi, k_lo, B, T1, C, C}, p));
  attention_alpha_kernel<<<static_cast<unsigned>((m2 + 7) / 8), 256, 0, st>>>(S, T1p, text_lengths, T1, T2, m2, alpha);
  CUDA_TRY(cudaGetLastError());
  c->launches++;
  return EFTS_OK;
}

int efts_imv_generator(const float* alpha, const float* p, const int32_t* text_lengths, const int32_t* speech_lengths,
                       int32_t B, int32_t T1, int32_t T2, float* imv, void* workspace, size_t workspace_bytes,
                       void* stream) {
  if (!alpha || !p || !text_lengths || !speech_lengths || !imv || !workspace || B < 1 || T1 < 1 || T2 < 1 || B > 65535)
    return fail(EFTS_ERR_ARG, "efts_imv_generator: bad argument");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  Arena a(workspace, workspace_bytes);
  float* raw = a.get<float>(static_cast<size_t>(B) * T2);
  if (!a.ok) return fail(EFTS_ERR_WORKSPACE, "workspace too small: need %zu, have %zu", a.off, workspace_bytes);
  alpha_expectation_kernel<<<dim3((T2 + 127) / 128, B), 128, 0, st>>>(alpha, p, T1, T2, raw);
  CUDA_TRY(cudaGetLastError());
  CUDA_TRY(cudaFuncSetAttribute(imv_scan_block_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                static_cast<int>(kReconstructSmemMax)));
  return launch_imv_scan(2, st, raw, nullptr, 0, text_lengths, speech_lengths, B, T2, imv);
}

int efts_aligned_positions(const float* imv, const float* p, const int32_t* text_lengths, const int32_t* speech_lengths,
                           int32_t B, int32_t T1, int32_t T2, float sigma_e, float* e, void* stream) {
  if (!imv || !text_lengths || !speech_lengths || !e || B < 1 || T1 < 1 || T2 < 1 || B > 65535)
    return fail(EFTS_ERR_ARG, "efts_aligned_positions: bad argument");
  CUDA_TRY(cudaFuncSetAttribute(aligned_positions_block_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                static_cast<int>(kReconstructSmemMax)));
  return launch_aligned_positions(2, static_cast<cudaStream_t>(stream), imv, text_lengths, speech_lengths, B, T1, T2,
                                  sigma_e, e, p);
}

int efts_reconstruct_alignment(const float* e, const int32_t* text_lengths, const int32_t* speech_lengths, int32_t B,
                               int32_t T1, int32_t T2, float delta, float* reconst_alpha, void* stream) {
  if (!e || !reconst_alpha || B < 1 || T1 < 1 || T2 < 1 || B > 65535)
    return fail(EFTS_ERR_ARG, "efts_reconstruct_alignment: bad argument");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int T1p = round8(T1);
  const float neg_sigma = -1.0f * delta;
  CUDA_TRY(cudaFuncSetAttribute(reconstruct_alignment_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                static_cast<int>(kReconstructSmemMax)));
  return launch_reconstruct(st, e, text_lengths, speech_lengths, B, T1, T2, T1p, neg_sigma, reconst_alpha, nullptr,
                            nullptr);
}

// ------------------------------------------------------------------------------------------------
int efts_length_regulator_plan(int64_t* ds, const int64_t* ilens, float alpha, int32_t B, int32_t T1,
                               int64_t* ds_eff, int64_t* out_lens, int64_t* plan, void* stream) {
  if (!ds || !ilens || !ds_eff || !out_lens || !plan || B < 1 || T1 < 1)
    return fail(EFTS_ERR_ARG, "efts_length_regulator_plan: bad argument");
  if (!(alpha > 0.0f)) return fail(EFTS_ERR_ARG, "alpha must be > 0 (layers/length_regulator.py:48)");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  CUDA_TRY(cudaMemsetAsync(plan, 0, 2 * sizeof(int64_t), st));
  length_regulator_plan_kernel<<<(B + 3) / 4, 128, 0, st>>>(
      reinterpret_cast<long long*>(ds), reinterpret_cast<const long long*>(ilens), alpha, alpha == 1.0f ? 1 : 0, B,
      T1, reinterpret_cast<long long*>(ds_eff), reinterpret_cast<long long*>(out_lens),
      reinterpret_cast<long long*>(plan));
  CUDA_TRY(cudaGetLastError());
  return EFTS_OK;
}

int efts_length_regulator_fwd(const float* xs, const int64_t* ds_eff, const int64_t* ilens, const int64_t* out_lens,
                              int32_t B, int32_t T1, int32_t D, int64_t Tout, float pad_value, float* out,
                              int64_t* idx, void* stream) {
  if (!xs || !ds_eff || !ilens || !out_lens || !out || B < 1 || T1 < 1 || D < 1 || Tout < 0)
    return fail(EFTS_ERR_ARG, "efts_length_regulator_fwd: bad argument");
  if (Tout == 0) return EFTS_OK;
  if (B > 65535) return fail(EFTS_ERR_ARG, "B too large");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const size_t smem = static_cast<size_t>(T1) * sizeof(long long);
  if (smem > 200 * 1024) return fail(EFTS_ERR_ARG, "T1=%d too long for the shared-memory scan", T1);
  if (smem > 48 * 1024)
    CUDA_TRY(cudaFuncSetAttribute(length_regulator_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  static_cast<int>(smem)));
  dim3 grid(static_cast<unsigned>((Tout + LR_FRAMES - 1) / LR_FRAMES), B);
  length_regulator_fwd_kernel<<<grid, 256, smem, st>>>(
      xs, reinterpret_cast<const long long*>(ds_eff), reinterpret_cast<const long long*>(ilens),
      reinterpret_cast<const long long*>(out_lens), T1, D, Tout, pad_value, out, reinterpret_cast<long long*>(idx));
  CUDA_TRY(cudaGetLastError());
  return EFTS_OK;
}

}  // extern "C"

// ================================================================================================
// HiFi-GAN V1 generator (vocoders/hifigan_model.py:95-136).  Every convolution is a launch of
// the fused-B tap-GEMM with the long A box (up to 11 taps, dilation up to 5):
//   * Conv1d(k, dilation d): taps shifted by (j - pad) * d rows;
//   * ConvTranspose1d(k = 2u, stride u, padding u/2): a 3-tap GEMM over input rows (q-1, q, q+1) whose N = u * C
//     output columns are the u output phases of row q -- the fp32 result [B, L, u*C] IS [B, L*u, C] in memory.
//     Phase r takes tap j = u * (1 - tau) + r + u/2 of input row q + tau - 1 (two of the three are in range);
//   * the activation in front of a conv (F.leaky_relu(x) then conv, :58,60,123) is applied when its operand planes
//     are written: act = LeakyReLU in the producer's epilogue when only the planes are needed, plane_act when the
//     pre-activation fp32 value is needed too (the residual x of ResBlock1, :62).
namespace {

struct VocWs {
  __half *m_hi, *m_lo;                 // mel planes [B, T, num_mels]
  float *x_f, *ra_f, *rb_f, *fin[4];   // fp32 [B, L, C] (stage-sized)
  __half *x_hi, *x_lo, *t_hi, *t_lo, *ra_hi, *ra_lo, *rb_hi, *rb_lo;
};

size_t voc_max_elems(const efts_vocoder_config& g, int B, int T) {
  size_t L = T, best = static_cast<size_t>(T) * g.upsample_initial_channel;
  int C = g.upsample_initial_channel;
  for (int i = 0; i < g.num_upsamples; ++i) {
    L *= g.upsample_rates[i];
    C /= 2;
    best = std::max(best, L * C);
  }
  return best * B;
}

void carve_voc(Arena& a, VocWs& w, const efts_vocoder_config& g, int B, int T) {
  const size_t n = voc_max_elems(g, B, T);
  w.m_hi = a.get<__half>(static_cast<size_t>(B) * T * g.num_mels);
  w.m_lo = a.get<__half>(static_cast<size_t>(B) * T * g.num_mels);
  w.x_f = a.get<float>(n); w.ra_f = a.get<float>(n); w.rb_f = a.get<float>(n);
  for (int k = 0; k < g.num_kernels; ++k) w.fin[k] = a.get<float>(n);
  w.x_hi = a.get<__half>(n); w.x_lo = a.get<__half>(n);
  w.t_hi = a.get<__half>(n); w.t_lo = a.get<__half>(n);
  w.ra_hi = a.get<__half>(n); w.ra_lo = a.get<__half>(n);
  w.rb_hi = a.get<__half>(n); w.rb_lo = a.get<__half>(n);
}

// One convolution of the generator: planes in -> bias, optional activation / residual -> fp32 and / or planes out.
int voc_conv(efts_ctx* c, cudaStream_t st, const PackedW& w, int dil, const __half* ahi, const __half* alo, int B,
             int L, int act, const float* resid, float* out_f, __half* ohi, __half* olo, int plane_act, int group = 1) {
  L /= group;                          // grouped packing: [B, L, C] read as [B, L / G, G * C] (pack_grouped)
  GemmParams p = gemm_defaults();
  p.N = w.N;
  p.ntaps = w.Z;
  p.pad = (w.Z - 1) / 2;
  p.dil = dil;
  p.act = act;
  p.bias = w.bias;
  p.resid = resid;
  p.out = out_f; p.ld_out = w.N;
  p.out_hi = ohi; p.out_lo = olo; p.ld_pl = w.N;
  p.plane_act = plane_act;
  p.long_taps = 1;
  return launch_gemm(c, st, OpA{ahi, alo, B, L, w.K, w.K}, weight_op(w), p);
}

// ConvTranspose1d weight [Cin, Cout, k] (k = 2u, padding u/2) -> 3-tap GEMM planes [3][u*Cout][Cin], bias tiled.
// fp32 weights [taps][N][K] (tap-major GEMM layout) -> fp16 hi / lo operand planes on the device.
int upload_planes(efts_ctx* c, const std::vector<float>& t, const std::string& what, __half** hi_dev, __half** lo_dev) {
  std::vector<__half> hi(t.size()), lo(t.size());
  for (size_t i = 0; i < t.size(); ++i) {
    const float x = t[i];
    if (!(fabsf(x) <= 65504.0f))
      return fail(EFTS_ERR_UNSUPPORTED, "weight '%s' has a value outside the fp16 operand range", what.c_str());
    const __half h = __float2half_rn(x);
    hi[i] = h;
    lo[i] = __float2half_rn((x - __half2float(h)) * SPLIT_SCALE);
  }
  TRY(upload(c, hi.data(), hi.size() * sizeof(__half), reinterpret_cast<void**>(hi_dev)));
  TRY(upload(c, lo.data(), lo.size() * sizeof(__half), reinterpret_cast<void**>(lo_dev)));
  return EFTS_OK;
}

// ConvTranspose1d weight [Cin, Cout, k] (k = 2u, stride u, padding u/2) -> [3][u*Cout][Cin]: output phase r of input
// row q takes tap j = u (1 - tau) + r + u/2 of input row q + tau - 1.
void map_transposed(const float* w, int Cin, int Cout, int k, int u, float* out) {
  const int N = u * Cout;
  for (int tau = 0; tau < 3; ++tau)
    for (int r = 0; r < u; ++r) {
      const int j = u * (1 - tau) + r + u / 2;
      for (int co = 0; co < Cout; ++co)
        for (int ci = 0; ci < Cin; ++ci)
          out[(static_cast<size_t>(tau) * N + r * Cout + co) * Cin + ci] =
              (j >= 0 && j < k) ? w[(static_cast<size_t>(ci) * Cout + co) * k + j] : 0.0f;
    }
}

int pack_ups(efts_ctx* c, const std::string& wname, const std::string& bname, int Cin, int Cout, int k, int u,
             PackedW* out) {
  const std::vector<float>* w;
  const std::vector<float>* b;
  TRY(need(c, wname, {Cin, Cout, k}, &w));
  TRY(need(c, bname, {Cout}, &b));
  const int N = u * Cout;
  std::vector<float> t(static_cast<size_t>(3) * N * Cin);
  map_transposed(w->data(), Cin, Cout, k, u, t.data());
  std::vector<float> bias(N);
  for (int r = 0; r < u; ++r)
    for (int co = 0; co < Cout; ++co) bias[r * Cout + co] = (*b)[co];
  out->Z = 3; out->N = N; out->K = Cin;
  TRY(upload_planes(c, t, wname, &out->hi, &out->lo));
  TRY(upload(c, bias.data(), N * sizeof(float), reinterpret_cast<void**>(&out->bias)));
  return EFTS_OK;
}

// Narrow layers (C = 32 / 64 channels) waste a 128 x 128 x 64 MMA tile: only C of 128 output columns and C of 64
// reduction columns are real.  The same buffer [B, L, C] read as [B, L / G, G * C] (G = 128 / C consecutive time
// steps per row) turns the conv into an undilated "super-tap" GEMM with K = N = 128:
//   y[G q + r, co] = sum_j sum_ci W[co, ci, j] * x[G q + r + (j - pad) d, ci]
//                  = sum_tau sum_{r', ci} W'[tau][(r, co)][(r', ci)] * X[q + tau][(r', ci)],
//   W'[tau][(r, co)][(r', ci)] = W[co, ci, j]  where  G tau + r' - r = (j - pad) d  (zero if no such j).
// Super-taps: 2 * floor((pad * d + G - 1) / G) + 1.  Every other part of the layer (bias, activation, residual,
// planes) is element-wise, so only the view changes.  k = 11, d = 1 at C = 32: 5 x 128 x 128 MACs per 4 steps
// instead of 4 x 11 x 64 x 128 (4.4 x fewer); the dilation moves into the weights.
int grouped_taps(int k, int d, int G) { return 2 * (((k - 1) / 2 * d + G - 1) / G) + 1; }

// Conv1d weight [C, C, k] with dilation d -> super-tap weights [S][G*C][G*C] (zero where no tap lands).
void map_grouped(const float* w, int C, int k, int d, int G, float* out) {
  const int S = grouped_taps(k, d, G), tmax = (S - 1) / 2, pad = (k - 1) / 2, N = G * C;
  std::fill(out, out + static_cast<size_t>(S) * N * N, 0.0f);
  for (int tau = -tmax; tau <= tmax; ++tau)
    for (int r = 0; r < G; ++r)
      for (int rp = 0; rp < G; ++rp) {
        const int off = G * tau + rp - r;                   // = (j - pad) * d
        if (off % d != 0) continue;
        const int j = off / d + pad;
        if (j < 0 || j >= k) continue;
        for (int co = 0; co < C; ++co)
          for (int ci = 0; ci < C; ++ci)
            out[(static_cast<size_t>(tau + tmax) * N + r * C + co) * N + rp * C + ci] =
                w[(static_cast<size_t>(co) * C + ci) * k + j];
      }
}

int pack_grouped(efts_ctx* c, const std::string& wname, const std::string& bname, int C, int k, int d, int G,
                 PackedW* out) {
  const std::vector<float>* w;
  const std::vector<float>* b;
  TRY(need(c, wname, {C, C, k}, &w));
  TRY(need(c, bname, {C}, &b));
  const int S = grouped_taps(k, d, G), N = G * C;
  std::vector<float> t(static_cast<size_t>(S) * N * N);
  map_grouped(w->data(), C, k, d, G, t.data());
  std::vector<float> bias(N);
  for (int r = 0; r < G; ++r)
    for (int co = 0; co < C; ++co) bias[r * C + co] = (*b)[co];
  out->Z = S; out->N = N; out->K = N;
  TRY(upload_planes(c, t, wname, &out->hi, &out->lo));
  TRY(upload(c, bias.data(), N * sizeof(float), reinterpret_cast<void**>(&out->bias)));
  return EFTS_OK;
}

// Plain or grouped packing of one resblock conv, whichever issues fewer MMA columns (grouping needs <= 15 super-taps
// and a sequence length that is a multiple of G: L = T * prod(rates) always is when the last rates cover G).
int pack_voc_conv(efts_ctx* c, const std::string& prefix, int C, int k, int d, int Lmult, PackedW* out, int* group,
                  int* dil) {
  // MMA columns issued per time step with G steps per row: taps * (K padded to 64) * (column tile: 64 when the
  // row has <= 64 columns and narrow tiles are on, else 128) / G
  auto cost = [&](int G) -> double {
    const int taps = G == 1 ? k : grouped_taps(k, d, G);
    if (taps > 15) return 1e30;
    const int W = G * C, Kp = (W + 63) / 64 * 64;
    const int tile = (W <= 64 && c->voc_narrow) ? 64 : 128;
    return static_cast<double>(taps) * Kp * tile * ((W + tile - 1) / tile) / G;
  };
  int G = 1;
  if (c->voc_group)
    for (int g2 = 2; g2 * C <= 128; g2 *= 2)
      if (Lmult % g2 == 0 && (cost(g2) < cost(G) || (c->voc_group > 1 && cost(g2) < c->voc_group * cost(G)))) G = g2;
  *group = G;
  *dil = G == 1 ? d : 1;
  if (G == 1) return pack_weight(c, prefix + ".weight", prefix + ".bias", C, C, k, out);
  return pack_grouped(c, prefix + ".weight", prefix + ".bias", C, k, d, G, out);
}

}  // namespace

extern "C" {

int efts_vocoder_create(const efts_vocoder_config* g, efts_ctx** out) {
  if (g == nullptr || out == nullptr) return fail(EFTS_ERR_ARG, "null argument");
  if (g->num_upsamples < 1 || g->num_upsamples > 8 || g->num_kernels < 1 || g->num_kernels > 4)
    return fail(EFTS_ERR_UNSUPPORTED, "num_upsamples / num_kernels out of range");
  if (g->num_mels % 8 != 0 || g->num_mels < 8) return fail(EFTS_ERR_UNSUPPORTED, "num_mels=%d must be a multiple of 8", g->num_mels);
  int C = g->upsample_initial_channel;
  if (C % 8 != 0 || C > G2_BIAS_MAX_LONG) return fail(EFTS_ERR_UNSUPPORTED, "upsample_initial_channel=%d", C);
  for (int i = 0; i < g->num_upsamples; ++i) {
    const int u = g->upsample_rates[i], k = g->upsample_kernel_sizes[i];
    if (u < 2 || (u & 1) || k != 2 * u)
      return fail(EFTS_ERR_UNSUPPORTED, "upsample %d: rate %d, kernel %d (supported: even rate, kernel = 2 * rate)", i, u, k);
    if (C % 2) return fail(EFTS_ERR_UNSUPPORTED, "odd channel count");
    C /= 2;
    if (C % 8 != 0) return fail(EFTS_ERR_UNSUPPORTED, "stage %d has %d channels (must be a multiple of 8)", i, C);
    if (u * C > G2_BIAS_MAX_LONG) return fail(EFTS_ERR_UNSUPPORTED, "rate * channels = %d exceeds %d", u * C, G2_BIAS_MAX_LONG);
  }
  if (C * 7 > 1024) return fail(EFTS_ERR_UNSUPPORTED, "conv_post with %d channels", C);
  if (g->num_upsamples * g->num_kernels > 32) return fail(EFTS_ERR_UNSUPPORTED, "too many resblocks");
  if (g->resblock_type != 1 && g->resblock_type != 2) return fail(EFTS_ERR_UNSUPPORTED, "resblock type %d", g->resblock_type);
  if (g->num_dilations < 1 || g->num_dilations > 3) return fail(EFTS_ERR_UNSUPPORTED, "num_dilations=%d", g->num_dilations);
  for (int j = 0; j < g->num_kernels; ++j) {
    const int k = g->resblock_kernel_sizes[j];
    if (k < 1 || k > 11 || !(k & 1)) return fail(EFTS_ERR_UNSUPPORTED, "resblock kernel size %d (odd, <= 11)", k);
    for (int m = 0; m < g->num_dilations; ++m) {
      const int d = g->resblock_dilations[j][m];
      if (d < 1 || G2_BM + (k - 1) * d > G2_A_ROWS_XLONG)
        return fail(EFTS_ERR_UNSUPPORTED, "resblock kernel %d with dilation %d exceeds the %d-row operand box", k, d,
                    G2_A_ROWS_XLONG);
    }
  }
  efts_ctx* c = nullptr;
  TRY(create_base(g->device, &c));
  c->voc = new efts_ctx::Vocoder();
  c->voc->cfg = *g;
  *out = c;
  return EFTS_OK;
}

int efts_vocoder_finalize(efts_ctx* c) {
  if (c == nullptr || c->voc == nullptr) return fail(EFTS_ERR_ARG, "not a vocoder context");
  if (c->finalized) return EFTS_OK;
  efts_ctx::Vocoder& v = *c->voc;
  const efts_vocoder_config& g = v.cfg;
  CUDA_TRY(cudaSetDevice(g.device));
  int C = g.upsample_initial_channel;
  int lmult = 1;                       // L is a multiple of this at the current stage
  TRY(pack_weight(c, "conv_pre.weight", "conv_pre.bias", C, g.num_mels, 7, &v.conv_pre));
  for (int i = 0; i < g.num_upsamples; ++i) {
    const std::string p = "ups." + std::to_string(i);
    TRY(pack_ups(c, p + ".weight", p + ".bias", C, C / 2, g.upsample_kernel_sizes[i], g.upsample_rates[i], &v.ups[i]));
    C /= 2;
    lmult *= g.upsample_rates[i];
    for (int j = 0; j < g.num_kernels; ++j) {
      const int n = i * g.num_kernels + j;
      for (int m = 0; m < g.num_dilations; ++m) {
        const std::string q = "resblocks." + std::to_string(n);
        int unused = 1;
        if (g.resblock_type == 2) {                 // ResBlock2: one conv per dilation, "convs.m"
          TRY(pack_voc_conv(c, q + ".convs." + std::to_string(m), C, g.resblock_kernel_sizes[j], g.resblock_dilations[j][m],
                            lmult, &v.c1[n][m], &v.g1[n][m], &v.d1[n][m]));
          continue;
        }
        TRY(pack_voc_conv(c, q + ".convs1." + std::to_string(m), C, g.resblock_kernel_sizes[j], g.resblock_dilations[j][m],
                          lmult, &v.c1[n][m], &v.g1[n][m], &v.d1[n][m]));
        TRY(pack_voc_conv(c, q + ".convs2." + std::to_string(m), C, g.resblock_kernel_sizes[j], 1, lmult, &v.c2[n][m],
                          &v.g2[n][m], &unused));
      }
    }
  }
  {   // conv_post: weight [1, C, 7] -> [7][C] fp32 for the direct kernel
    const std::vector<float>* w;
    const std::vector<float>* b;
    TRY(need(c, "conv_post.weight", {1, C, 7}, &w));
    TRY(need(c, "conv_post.bias", {1}, &b));
    std::vector<float> t(static_cast<size_t>(7) * C);
    for (int j = 0; j < 7; ++j)
      for (int ch = 0; ch < C; ++ch) t[static_cast<size_t>(j) * C + ch] = (*w)[static_cast<size_t>(ch) * 7 + j];
    TRY(upload(c, t.data(), t.size() * sizeof(float), reinterpret_cast<void**>(&v.post_w)));
    v.post_b = (*b)[0];
  }
  c->raw.clear();
  c->raw_shape.clear();
  c->finalized = true;
  return EFTS_OK;
}

size_t efts_vocoder_workspace_bytes(const efts_ctx* c, int32_t B, int32_t T) {
  if (c == nullptr || c->voc == nullptr || B < 1 || T < 1) return 0;
  Arena a(nullptr, ~static_cast<size_t>(0));
  VocWs w;
  carve_voc(a, w, c->voc->cfg, B, T);
  return a.off + 256;
}

int efts_vocoder_forward(efts_ctx* c, const float* mel, int32_t B, int32_t T, float* audio, void* workspace,
                         size_t workspace_bytes, void* stream) {
  TRY(check_ready(c));
  if (c->voc == nullptr) return fail(EFTS_ERR_ARG, "not a vocoder context");
  if (!mel || !audio || !workspace || B < 1 || T < 1 || B > 65535) return fail(EFTS_ERR_ARG, "efts_vocoder_forward: bad argument");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  efts_ctx::Vocoder& v = *c->voc;
  const efts_vocoder_config& g = v.cfg;
  Arena a(workspace, workspace_bytes);
  VocWs w;
  carve_voc(a, w, g, B, T);
  if (!a.ok) return fail(EFTS_ERR_WORKSPACE, "workspace too small: need %zu, have %zu", a.off, workspace_bytes);
  CUDA_TRY(cudaMemsetAsync(c->err_flag, 0, sizeof(int), st));
  voc_mel_planes_kernel<<<dim3((T + 31) / 32, (g.num_mels + 31) / 32, B), dim3(32, 8), 0, st>>>(mel, g.num_mels, T, w.m_hi,
                                                                                              w.m_lo, c->err_flag);
  CUDA_TRY(cudaGetLastError());
  c->launches++;
  // conv_pre (:121); its only consumer is leaky_relu -> ups[0] (:123-124): planes of the activated value
  TRY(voc_conv(c, st, v.conv_pre, 1, w.m_hi, w.m_lo, B, T, ACT_LRELU, nullptr, nullptr, w.t_hi, w.t_lo, 0));
  size_t L = T;
  int C = g.upsample_initial_channel;
  for (int i = 0; i < g.num_upsamples; ++i) {
    const int u = g.upsample_rates[i];
    // ups[i] (:124): x fp32 [B, L*u, C/2] (the residual input of every resblock) + planes of leaky_relu(x) (:58)
    TRY(voc_conv(c, st, v.ups[i], 1, w.t_hi, w.t_lo, B, static_cast<int>(L), ACT_NONE, nullptr, w.x_f, w.x_hi, w.x_lo, 1));
    L *= u;
    C /= 2;
    if (L > 0x7fffffff / 2) return fail(EFTS_ERR_ARG, "sequence too long");
    const int Li = static_cast<int>(L);
    for (int j = 0; j < g.num_kernels; ++j) {           // ResBlock1.forward (:56-63)
      const int n = i * g.num_kernels + j;
      const float* cur_f = w.x_f;
      const __half *cur_hi = w.x_hi, *cur_lo = w.x_lo;
      const int nd = g.num_dilations;
      for (int m = 0; m < nd && g.resblock_type == 2; ++m) {          // ResBlock2.forward (:83-88): x = c(lrelu(x)) + x
        const bool last = m == nd - 1;
        float* of = last ? w.fin[j] : ((m & 1) ? w.rb_f : w.ra_f);
        __half* oh = last ? nullptr : ((m & 1) ? w.rb_hi : w.ra_hi);
        __half* ol = last ? nullptr : ((m & 1) ? w.rb_lo : w.ra_lo);
        TRY(voc_conv(c, st, v.c1[n][m], v.d1[n][m], cur_hi, cur_lo, B, Li, ACT_NONE, cur_f, of, oh, ol, 1, v.g1[n][m]));
        cur_f = of; cur_hi = oh; cur_lo = ol;
      }
      for (int m = 0; m < nd && g.resblock_type == 1; ++m) {
        // xt = c1(leaky_relu(x)); only leaky_relu(xt) is consumed (:59-60)
        TRY(voc_conv(c, st, v.c1[n][m], v.d1[n][m], cur_hi, cur_lo, B, Li, ACT_LRELU, nullptr, nullptr, w.t_hi, w.t_lo, 0,
                     v.g1[n][m]));
        // x = c2(...) + x (:61-62): fp32 x for the next residual, planes of leaky_relu(x) for the next c1
        const bool last = m == nd - 1;
        float* of = last ? w.fin[j] : ((m & 1) ? w.rb_f : w.ra_f);
        __half* oh = last ? nullptr : ((m & 1) ? w.rb_hi : w.ra_hi);
        __half* ol = last ? nullptr : ((m & 1) ? w.rb_lo : w.ra_lo);
        TRY(voc_conv(c, st, v.c2[n][m], 1, w.t_hi, w.t_lo, B, Li, ACT_NONE, cur_f, of, oh, ol, 1, v.g2[n][m]));
        cur_f = of; cur_hi = oh; cur_lo = ol;
      }
    }
    // x = (r0 + r1 + r2) / num_kernels (:126-131), then leaky_relu for the next ups (:123) or conv_post (:132)
    VocAvgArgs av;
    av.n = g.num_kernels;
    for (int k = 0; k < 4; ++k) av.r[k] = k < g.num_kernels ? w.fin[k] : nullptr;
    const size_t n4 = static_cast<size_t>(B) * L * C / 4;
    const unsigned grid = static_cast<unsigned>(std::min<size_t>((n4 + 255) / 256, static_cast<size_t>(c->sm_count) * 16));
    const bool final_stage = i == g.num_upsamples - 1;
    voc_average_kernel<<<grid, 256, 0, st>>>(av, n4, final_stage ? 0.01f : 0.1f, final_stage ? nullptr : w.t_hi,
                                             final_stage ? nullptr : w.t_lo, final_stage ? w.x_f : nullptr, c->err_flag);
    CUDA_TRY(cudaGetLastError());
    c->launches++;
  }
  // conv_post + tanh (:133-134)
  if (C <= VOC_POST_CMAX && C % 4 == 0)
    voc_post_tiled_kernel<<<dim3(static_cast<unsigned>((L + VOC_POST_BLOCK - 1) / VOC_POST_BLOCK), B), VOC_POST_BLOCK, 0, st>>>(
        w.x_f, v.post_w, v.post_b, static_cast<int>(L), C, 7, audio);
  else
    voc_post_kernel<<<dim3(static_cast<unsigned>((L + 255) / 256), B), 256, 0, st>>>(w.x_f, v.post_w, v.post_b,
                                                                                   static_cast<int>(L), C, 7, audio);
  CUDA_TRY(cudaGetLastError());
  c->launches++;
  return EFTS_OK;
}


// Host-only views of the two weight re-arrangements above (no device needed): what the packers feed the tap-GEMM,
// in fp32, so that the index algebra can be checked against torch's conv1d / conv_transpose1d on the CPU.
int efts_host_map_transposed(const float* w, int32_t Cin, int32_t Cout, int32_t k, int32_t u, float* out) {
  if (!w || !out || Cin < 1 || Cout < 1 || u < 2 || (u & 1) || k != 2 * u) return fail(EFTS_ERR_ARG, "efts_host_map_transposed: bad argument");
  map_transposed(w, Cin, Cout, k, u, out);
  return EFTS_OK;
}
int efts_host_map_grouped(const float* w, int32_t C, int32_t k, int32_t d, int32_t G, float* out, int32_t* taps) {
  if (!w || !taps || C < 1 || k < 1 || !(k & 1) || d < 1 || G < 1) return fail(EFTS_ERR_ARG, "efts_host_map_grouped: bad argument");
  *taps = grouped_taps(k, d, G);
  if (out != nullptr) map_grouped(w, C, k, d, G, out);
  return EFTS_OK;
}

}  // extern "C"

// ------------------------------------------------------------------------------------------------
// Log-mel front-end: datasets/meldataset.py:49-82 on the GPU.
namespace {
struct FeWs {
  int* lens32;
  __half *ch_hi, *ch_lo;       // chunk matrix planes [B, R, hop]
  __half *mg_hi, *mg_lo;       // magnitude planes [B, Tmax, Kp]
};
void carve_frontend(Arena& a, FeWs& w, const efts_ctx::Frontend& f, int B, int Tmax) {
  const size_t R = static_cast<size_t>(Tmax) + f.taps - 1;
  w.lens32 = a.get<int>(B);
  w.ch_hi = a.get<__half>(B * R * f.cfg.hop_size);
  w.ch_lo = a.get<__half>(B * R * f.cfg.hop_size);
  w.mg_hi = a.get<__half>(static_cast<size_t>(B) * Tmax * f.Kp);
  w.mg_lo = a.get<__half>(static_cast<size_t>(B) * Tmax * f.Kp);
}
}  // namespace

extern "C" {

int efts_frontend_create(const efts_frontend_config* g, efts_ctx** out) {
  if (g == nullptr || out == nullptr) return fail(EFTS_ERR_ARG, "null argument");
  if (g->n_fft < 64 || (g->n_fft & (g->n_fft - 1)) != 0) return fail(EFTS_ERR_UNSUPPORTED, "n_fft=%d must be a power of two", g->n_fft);
  if (g->win_size != g->n_fft) return fail(EFTS_ERR_UNSUPPORTED, "win_size=%d must equal n_fft=%d", g->win_size, g->n_fft);
  if (g->hop_size < 64 || g->hop_size % 64 != 0 || g->n_fft % g->hop_size != 0 || g->n_fft / g->hop_size > 9)
    return fail(EFTS_ERR_UNSUPPORTED, "hop_size=%d: must be a multiple of 64 dividing n_fft, at most 9 hops per window", g->hop_size);
  if (g->num_mels % 8 != 0 || g->num_mels < 8 || g->num_mels > 1024) return fail(EFTS_ERR_UNSUPPORTED, "num_mels=%d", g->num_mels);
  efts_ctx* c = nullptr;
  TRY(create_base(g->device, &c));
  c->fe = new efts_ctx::Frontend();
  c->fe->cfg = *g;
  // A DFT bin far below the frame's peak is a small difference of large partial sums, and the tensor core's fp32
  // accumulation truncates: the main accumulator is flushed after every k-block here (measured: log-mel error on
  // bands above 1e-3 of the frame's peak 1.3e-4 -> 6.6e-5 against a float64 evaluation; the fp32 FFT of the reference
  // is at 1.8e-5)
  c->chunk_kb = 1;
  c->fe->taps = g->n_fft / g->hop_size;
  c->fe->half = g->n_fft / 2;
  c->fe->Kp = round8(g->n_fft / 2 + 1);     // until finalize has seen the mel filter bank
  *out = c;
  return EFTS_OK;
}

int efts_frontend_finalize(efts_ctx* c) {
  if (c == nullptr || c->fe == nullptr) return fail(EFTS_ERR_ARG, "not a front-end context");
  if (c->finalized) return EFTS_OK;
  efts_ctx::Frontend& f = *c->fe;
  CUDA_TRY(cudaSetDevice(f.cfg.device));
  // Uploaded: "stft.weight" [2 half, hop, taps] (rows 0 .. half: re_0 .. re_half, rows half + 1 ..: im_1 .. im_{half-1}; im_0
  // and im_half vanish for a real signal) and "mel_basis.weight" [mels, round8(half + 1)].  Packed: only the bins some mel
  // filter weighs (fmax = 8 kHz at 22.05 kHz keeps 372 of 513 -- the others are multiplied by zero in :74), and the two
  // columns of a bin side by side, (re_f, im_f), so that the STFT GEMM's epilogue can write the magnitude itself.
  const int full = round8(f.half + 1);
  const std::vector<float>* dft;
  const std::vector<float>* mel;
  TRY(need(c, "stft.weight", {2 * f.half, f.cfg.hop_size, f.taps}, &dft));
  TRY(need(c, "mel_basis.weight", {f.cfg.num_mels, full}, &mel));
  int nb = 1;
  for (int m = 0; m < f.cfg.num_mels; ++m)
    for (int k = 0; k <= f.half; ++k)
      if ((*mel)[static_cast<size_t>(m) * full + k] != 0.0f) nb = std::max(nb, k + 1);
  f.Kp = round8(nb);
  const size_t per_col = static_cast<size_t>(f.cfg.hop_size) * f.taps;
  std::vector<float> pairs(static_cast<size_t>(2) * f.Kp * per_col, 0.0f), trim(static_cast<size_t>(f.cfg.num_mels) * f.Kp, 0.0f);
  for (int k = 0; k < f.Kp && k <= f.half; ++k) {
    std::copy(dft->begin() + k * per_col, dft->begin() + (k + 1) * per_col, pairs.begin() + (2 * k) * per_col);
    if (k > 0 && k < f.half)
      std::copy(dft->begin() + (f.half + k) * per_col, dft->begin() + (f.half + k + 1) * per_col,
                pairs.begin() + (2 * k + 1) * per_col);
  }
  for (int m = 0; m < f.cfg.num_mels; ++m)
    for (int k = 0; k < f.Kp && k < full; ++k) trim[static_cast<size_t>(m) * f.Kp + k] = (*mel)[static_cast<size_t>(m) * full + k];
  c->raw["stft.pairs.weight"] = pairs;
  c->raw_shape["stft.pairs.weight"] = {2 * f.Kp, f.cfg.hop_size, f.taps};
  c->raw["stft.pairs.bias"] = std::vector<float>(static_cast<size_t>(2) * f.Kp, 0.0f);
  c->raw_shape["stft.pairs.bias"] = {2 * f.Kp};
  c->raw["mel_trim.weight"] = trim;
  c->raw_shape["mel_trim.weight"] = {f.cfg.num_mels, f.Kp};
  TRY(pack_weight(c, "stft.pairs.weight", "stft.pairs.bias", 2 * f.Kp, f.cfg.hop_size, f.taps, &f.dft));
  TRY(pack_weight(c, "mel_trim.weight", "mel_basis.bias", f.cfg.num_mels, f.Kp, 1, &f.mel));
  c->raw.clear();
  c->raw_shape.clear();
  c->finalized = true;
  return EFTS_OK;
}

int32_t efts_frontend_frames(const efts_ctx* c, int64_t length) {
  if (c == nullptr || c->fe == nullptr || length < 0) return -1;
  return frontend_frames(length, c->fe->cfg.n_fft, c->fe->cfg.hop_size);
}

size_t efts_frontend_workspace_bytes(const efts_ctx* c, int32_t B, int32_t Lmax) {
  if (c == nullptr || c->fe == nullptr || B < 1 || Lmax < 1) return 0;
  Arena a(nullptr, ~static_cast<size_t>(0));
  FeWs w;
  carve_frontend(a, w, *c->fe, B, std::max(1, frontend_frames(Lmax, c->fe->cfg.n_fft, c->fe->cfg.hop_size)));
  return a.off + 4096;
}

int efts_frontend_forward(efts_ctx* c, const float* audio, const int64_t* lengths, int32_t B, int32_t Lmax, float* mel,
                          int64_t* mel_lengths, void* workspace, size_t workspace_bytes, void* stream) {
  if (c == nullptr || c->fe == nullptr) return fail(EFTS_ERR_ARG, "not a front-end context");
  if (!c->finalized) return fail(EFTS_ERR_STATE, "front-end weights not finalised");
  if (!audio || !mel || !workspace || B < 1 || Lmax < 1 || B > 65535) return fail(EFTS_ERR_ARG, "efts_frontend_forward: bad argument");
  const efts_ctx::Frontend& f = *c->fe;
  const int Tmax = frontend_frames(Lmax, f.cfg.n_fft, f.cfg.hop_size);
  if (Tmax < 1) return fail(EFTS_ERR_DATA, "%d samples are shorter than one analysis window after padding", Lmax);
  if (Tmax > 65535) return fail(EFTS_ERR_ARG, "utterances of more than 65535 frames are not supported");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int R = Tmax + f.taps - 1;
  Arena a(workspace, workspace_bytes);
  FeWs w;
  carve_frontend(a, w, f, B, Tmax);
  if (!a.ok) return fail(EFTS_ERR_WORKSPACE, "workspace too small: need %zu, have %zu", a.off, workspace_bytes);
  CUDA_TRY(cudaMemsetAsync(c->err_flag, 0, sizeof(int), st));
  // 1. reflect padding (:66) + hop-sized chunk matrix as operand planes
  const size_t n4 = static_cast<size_t>(R) * f.cfg.hop_size / 4;
  const unsigned gx = static_cast<unsigned>(std::min<size_t>((n4 + 255) / 256, 1024));
  frontend_chunk_planes_kernel<<<dim3(gx, B), 256, 0, st>>>(audio, reinterpret_cast<const long long*>(lengths), B, Lmax, R,
                                                            f.cfg.n_fft, f.cfg.hop_size, w.ch_hi, w.ch_lo,
                                                            reinterpret_cast<long long*>(mel_lengths), w.lens32, c->err_flag);
  CUDA_TRY(cudaGetLastError());
  c->launches++;
  // 2. STFT (:69-70): frame t = rows t .. t + taps - 1 of the chunk matrix against the windowed DFT basis, and
  // 3. magnitude sqrt(re^2 + im^2 + 1e-9) (:72) in its epilogue (columns are (re, im) pairs), as operand planes
  {
    GemmParams p = gemm_defaults();
    p.N = 2 * f.Kp; p.ntaps = f.taps; p.pad = 0; p.mag_pairs = 1; p.lens = w.lens32;
    p.out_hi = w.mg_hi; p.out_lo = w.mg_lo; p.ld_pl = f.Kp;
    ProfScope ps(c, st, TAG_LINEAR);
    OpA chunks{w.ch_hi, w.ch_lo, B, Tmax, f.cfg.hop_size, f.cfg.hop_size};
    chunks.map_rows = R;
    TRY(launch_gemm(c, st, chunks, weight_op(f.dft), p));
  }
  // 4. mel projection (:74) with log(clamp(x, 1e-5)) (:75) in the epilogue; frames beyond an utterance's length are zero
  GemmParams p = gemm_defaults();
  p.N = f.cfg.num_mels; p.act = ACT_LOGCLAMP; p.lens = w.lens32;
  p.out = mel; p.ld_out = f.cfg.num_mels;
  ProfScope ps(c, st, TAG_LINEAR);
  return launch_gemm(c, st, OpA{w.mg_hi, w.mg_lo, B, Tmax, f.Kp, f.Kp}, weight_op(f.mel), p);
}

}  // extern "C"

// ------------------------------------------------------------------------------------------------
// Training slice: ResConvBlock forward with saved activations and its backward
// (layers/efts_modules.py:48-51,77-79 under autograd; trainers/efficient_tts_trainer.py:152-154).
namespace {
struct TrainWs {
  __half *w_hi, *w_lo;                 // packed weights of the current layer [k][C][C]
  __half *a_hi[2], *a_lo[2];           // operand planes of a [B, T, C] activation (ping-pong)
  float* dwt;                          // [k][C][C]
  float* splitk;                       // partial planes of the split reductions
  double* db_part;                     // [kBiasParts][C]
  float* g[2];                         // running data gradient [B, T, C] (ping-pong)
  size_t ktot; int Tp;
};
constexpr int kBiasParts = 592;
void carve_train(Arena& a, TrainWs& w, int B, int T, int C, int k, bool backward) {
  const size_t m = static_cast<size_t>(B) * T;
  (void)k;
  w.Tp = (T + G2_BK - 1) / G2_BK * G2_BK;            // positions of an utterance = whole k-blocks of the weight gradient
  w.ktot = static_cast<size_t>(B) * w.Tp;
  w.w_hi = a.get<__half>(static_cast<size_t>(k) * C * C);
  w.w_lo = a.get<__half>(static_cast<size_t>(k) * C * C);
  for (int i = 0; i < 2; ++i) { w.a_hi[i] = a.get<__half>(m * C); w.a_lo[i] = a.get<__half>(m * C); }
  if (backward) {
    w.dwt = a.get<float>(static_cast<size_t>(k) * C * C);
    w.splitk = a.get<float>(kSplitScratchBytes / sizeof(float));
    w.db_part = a.get<double>(static_cast<size_t>(kBiasParts) * C);
    for (int i = 0; i < 2; ++i) w.g[i] = a.get<float>(m * C);
  }
}
unsigned ew_grid(size_t n) { return static_cast<unsigned>(std::max<size_t>(1, std::min<size_t>((n + 255) / 256, 148 * 16))); }

// Backward of one conv layer  u = act(conv_k(x) + b)  given g = dL/du-side gradient BEFORE the activation mask:
// G' = g * (u > 0 ? 1 : slope);  grad_w [C][C][k], grad_b [C];  dx = conv_k^T(G') (+ resid when given).
int conv_layer_bwd(efts_ctx* c, cudaStream_t st, TrainWs& w, const float* g, const float* u, float slope, const float* xl,
                   const float* weights_l, int k, int B, int T, const float* resid, float* dx, float* grad_w_l,
                   float* grad_b_l) {
  const int C = 512;
  const int pad = (k - 1) / 2;
  const size_t rows = static_cast<size_t>(B) * T, n = rows * C, wn = static_cast<size_t>(k) * C * C;
  // split of the position reduction: ~one work item per CTA pair
  const int num_kb = static_cast<int>((w.ktot + G2_BK - 1) / G2_BK);
  const int items = ((C / G2_BM + 1) / 2) * (C / G2_BN);
  int want = std::max(1, std::min(16, (c->sm_count / 2) / items));
  int split_kb = (num_kb + want - 1) / want;
  split_kb = (split_kb + c->chunk_kb - 1) / std::max(1, c->chunk_kb) * std::max(1, c->chunk_kb);
  // G' planes (both gradients' A operand), the planes of the layer input (weight gradient's B operand), bias gradient
  lrelu_grad_split_kernel<<<ew_grid(n / 4), 256, 0, st>>>(g, u, n / 4, slope, w.a_hi[0], w.a_lo[0], c->err_flag);
  CUDA_TRY(cudaGetLastError());
  TRY(split_planes(c, st, xl, n, w.a_hi[1], w.a_lo[1]));
  bias_grad_partial_kernel<<<dim3(kBiasParts, C / 128), 128, 0, st>>>(g, u, slope, rows, C, w.db_part);
  CUDA_TRY(cudaGetLastError());
  bias_grad_finish_kernel<<<(C + 127) / 128, 128, 0, st>>>(w.db_part, kBiasParts, C, grad_b_l);
  CUDA_TRY(cudaGetLastError());
  c->launches += 3;
  // dL/dW[o, c, j] = sum over positions of G'[pos, o] x[pos + j - pad, c]: one position-reduction GEMM per tap.  Both
  // operands are the activations' own planes [B, T, C] read MN-major, utterance by utterance in 64-row boxes (rows
  // outside the utterance are zero-filled by the TMA unit: the conv's zero padding, and the tail of the last box); the
  // tap is a row offset of B's box -- no transposed copy of either tensor.
  for (int j = 0; j < k; ++j) {
    GemmParams p = gemm_defaults();
    p.N = C; p.out = w.dwt + static_cast<size_t>(j) * C * C; p.ld_out = C;
    p.split_kb = split_kb; p.split_scratch = w.splitk;
    p.bmn_per = w.Tp / G2_BK; p.b_koff = j - pad; p.amn = 1;
    ProfScope ps(c, st, TAG_LINEAR);
    TRY(launch_gemm(c, st, OpA{w.a_hi[0], w.a_lo[0], B, T, C, C}, OpB{w.a_hi[1], w.a_lo[1], B, T, C, C}, p));
  }
  weight_grad_permute_kernel<<<ew_grid(wn), 256, 0, st>>>(w.dwt, C, C, k, grad_w_l);
  CUDA_TRY(cudaGetLastError());
  // dL/dx = [resid +] conv^T(G'): the tap-GEMM with flipped, transposed weights
  pack_conv_weight_kernel<true><<<ew_grid(wn), 256, 0, st>>>(weights_l, C, C, k, w.w_hi, w.w_lo, c->err_flag);
  CUDA_TRY(cudaGetLastError());
  c->launches += 2;
  GemmParams p = gemm_defaults();
  p.N = C; p.ntaps = k; p.pad = pad; p.resid = resid; p.out = dx; p.ld_out = C;
  { ProfScope ps(c, st, TAG_DEC_CONV); TRY(launch_gemm(c, st, OpA{w.a_hi[0], w.a_lo[0], B, T, C, C}, OpB{w.w_hi, w.w_lo, k, C, C, C}, p)); }
  return EFTS_OK;
}
}  // namespace

extern "C" {

size_t efts_resconv_train_workspace_bytes(const efts_ctx* c, int32_t B, int32_t T, int32_t k) {
  if (c == nullptr || B < 1 || T < 1 || k < 1) return 0;
  Arena a(nullptr, ~static_cast<size_t>(0));
  TrainWs w;
  carve_train(a, w, B, T, 512, k, true);
  return a.off + 4096;
}

int efts_resconv_train_fwd(efts_ctx* c, const float* x, const float* weights, const float* biases, int32_t n_layers,
                           int32_t k, int32_t B, int32_t T, float* acts, float* us, void* workspace,
                           size_t workspace_bytes, void* stream) {
  if (c == nullptr) return fail(EFTS_ERR_ARG, "null context");
  const int C = 512;
  if (!x || !weights || !biases || !acts || !us || !workspace || n_layers < 1 || B < 1 || T < 1 || (k != 1 && k != 3 && k != 5))
    return fail(EFTS_ERR_ARG, "efts_resconv_train_fwd: bad argument");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  Arena a(workspace, workspace_bytes);
  TrainWs w;
  carve_train(a, w, B, T, C, k, false);
  if (!a.ok) return fail(EFTS_ERR_WORKSPACE, "workspace too small: need %zu, have %zu", a.off, workspace_bytes);
  const size_t n = static_cast<size_t>(B) * T * C;
  CUDA_TRY(cudaMemsetAsync(c->err_flag, 0, sizeof(int), st));
  CUDA_TRY(cudaMemcpyAsync(acts, x, n * sizeof(float), cudaMemcpyDeviceToDevice, st));
  TRY(split_planes(c, st, x, n, w.a_hi[0], w.a_lo[0]));
  int cur = 0;
  for (int l = 0; l < n_layers; ++l, cur ^= 1) {
    const size_t wn = static_cast<size_t>(k) * C * C;
    pack_conv_weight_kernel<false><<<ew_grid(wn), 256, 0, st>>>(weights + l * wn, C, C, k, w.w_hi, w.w_lo, c->err_flag);
    CUDA_TRY(cudaGetLastError());
    float* u = us + l * n;
    GemmParams p = gemm_defaults();
    p.N = C; p.ntaps = k; p.pad = (k - 1) / 2; p.act = ACT_LRELU; p.bias = biases + static_cast<size_t>(l) * C;
    p.out = u; p.ld_out = C;
    { ProfScope ps(c, st, TAG_DEC_CONV); TRY(launch_gemm(c, st, OpA{w.a_hi[cur], w.a_lo[cur], B, T, C, C}, OpB{w.w_hi, w.w_lo, k, C, C, C}, p)); }
    residual_add_split_kernel<<<ew_grid(n / 4), 256, 0, st>>>(acts + l * n, u, n / 4, acts + (l + 1) * n, w.a_hi[cur ^ 1],
                                                              w.a_lo[cur ^ 1], c->err_flag);
    CUDA_TRY(cudaGetLastError());
    c->launches += 2;
  }
  return EFTS_OK;
}

int efts_resconv_train_bwd(efts_ctx* c, const float* grad_out, const float* acts, const float* us, const float* weights,
                           int32_t n_layers, int32_t k, int32_t B, int32_t T, float* grad_x, float* grad_w, float* grad_b,
                           void* workspace, size_t workspace_bytes, void* stream) {
  if (c == nullptr) return fail(EFTS_ERR_ARG, "null context");
  const int C = 512;
  if (!grad_out || !acts || !us || !weights || !grad_x || !grad_w || !grad_b || !workspace || n_layers < 1 || B < 1 || T < 1 ||
      (k != 1 && k != 3 && k != 5) || B > 65535)
    return fail(EFTS_ERR_ARG, "efts_resconv_train_bwd: bad argument");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  Arena a(workspace, workspace_bytes);
  TrainWs w;
  carve_train(a, w, B, T, C, k, true);
  if (!a.ok) return fail(EFTS_ERR_WORKSPACE, "workspace too small: need %zu, have %zu", a.off, workspace_bytes);
  const size_t n = static_cast<size_t>(B) * T * C, wn = static_cast<size_t>(k) * C * C;
  CUDA_TRY(cudaMemsetAsync(c->err_flag, 0, sizeof(int), st));
  const float* g = grad_out;
  for (int l = n_layers - 1; l >= 0; --l) {
    float* dx = l == 0 ? grad_x : w.g[l & 1];
    // y = x + lrelu(conv(x) + b): dL/dx = g + conv^T(G'), the residual slot of the data-gradient GEMM carries g
    TRY(conv_layer_bwd(c, st, w, g, us + l * n, 0.1f, acts + l * n, weights + l * wn, k, B, T, g, dx, grad_w + l * wn,
                       grad_b + static_cast<size_t>(l) * C));
    g = dx;
  }
  return EFTS_OK;
}

// ---- duration predictor, training (layers/duration_predictor.py:57-88) ----
namespace {
constexpr int kLnWarps = 148 * 8;     // warps of ln_train_bwd_kernel = partial rows of its column sums
}
size_t efts_duration_train_workspace_bytes(const efts_ctx* c, int32_t B, int32_t T, int32_t k) {
  if (c == nullptr || B < 1 || T < 1 || k < 1) return 0;
  Arena a(nullptr, ~static_cast<size_t>(0));
  TrainWs w;
  carve_train(a, w, B, T, 512, k, true);
  a.get<double>(static_cast<size_t>(4) * kLnWarps * 512);
  return a.off + 4096;
}

int efts_duration_train_fwd(efts_ctx* c, const float* x, const float* conv_w, const float* conv_b, const float* ln_g,
                            const float* ln_b, const float* head_w, const float* head_b, const uint8_t* mask,
                            const float* keep, int32_t n_layers, int32_t k, int32_t B, int32_t T, float* acts, float* us,
                            float* out, void* workspace, size_t workspace_bytes, void* stream) {
  if (c == nullptr) return fail(EFTS_ERR_ARG, "null context");
  const int C = 512;
  if (!x || !conv_w || !conv_b || !ln_g || !ln_b || !head_w || !head_b || !acts || !us || !out || !workspace ||
      n_layers < 1 || B < 1 || T < 1 || (k != 1 && k != 3 && k != 5))
    return fail(EFTS_ERR_ARG, "efts_duration_train_fwd: bad argument");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  Arena a(workspace, workspace_bytes);
  TrainWs w;
  carve_train(a, w, B, T, C, k, false);
  if (!a.ok) return fail(EFTS_ERR_WORKSPACE, "workspace too small: need %zu, have %zu", a.off, workspace_bytes);
  const size_t rows = static_cast<size_t>(B) * T, n = rows * C, wn = static_cast<size_t>(k) * C * C;
  CUDA_TRY(cudaMemsetAsync(c->err_flag, 0, sizeof(int), st));
  CUDA_TRY(cudaMemcpyAsync(acts, x, n * sizeof(float), cudaMemcpyDeviceToDevice, st));
  TRY(split_planes(c, st, x, n, w.a_hi[0], w.a_lo[0]));
  int cur = 0;
  for (int l = 0; l < n_layers; ++l, cur ^= 1) {
    pack_conv_weight_kernel<false><<<ew_grid(wn), 256, 0, st>>>(conv_w + l * wn, C, C, k, w.w_hi, w.w_lo, c->err_flag);
    CUDA_TRY(cudaGetLastError());
    float* u = us + l * n;
    GemmParams p = gemm_defaults();
    p.N = C; p.ntaps = k; p.pad = (k - 1) / 2; p.act = ACT_RELU; p.bias = conv_b + static_cast<size_t>(l) * C;
    p.out = u; p.ld_out = C;
    { ProfScope ps(c, st, TAG_DURATION); TRY(launch_gemm(c, st, OpA{w.a_hi[cur], w.a_lo[cur], B, T, C, C}, OpB{w.w_hi, w.w_lo, k, C, C, C}, p)); }
    const bool last = l == n_layers - 1;
    ln_train_fwd_kernel<<<static_cast<unsigned>((rows + 7) / 8), 256, 0, st>>>(
        u, rows, ln_g + static_cast<size_t>(l) * C, ln_b + static_cast<size_t>(l) * C, keep ? keep + l * n : nullptr,
        acts + (l + 1) * n, last ? nullptr : w.a_hi[cur ^ 1], last ? nullptr : w.a_lo[cur ^ 1], last ? head_w : nullptr,
        head_b, mask, out, c->err_flag);
    CUDA_TRY(cudaGetLastError());
    c->launches += 2;
  }
  return EFTS_OK;
}

int efts_duration_train_bwd(efts_ctx* c, const float* grad_out, const float* acts, const float* us, const float* conv_w,
                            const float* ln_g, const float* head_w, const uint8_t* mask, const float* keep,
                            int32_t n_layers, int32_t k, int32_t B, int32_t T, float* grad_x, float* grad_conv_w,
                            float* grad_conv_b, float* grad_ln_g, float* grad_ln_b, float* grad_head_w,
                            float* grad_head_b, void* workspace, size_t workspace_bytes, void* stream) {
  if (c == nullptr) return fail(EFTS_ERR_ARG, "null context");
  const int C = 512;
  if (!grad_out || !acts || !us || !conv_w || !ln_g || !head_w || !grad_x || !grad_conv_w || !grad_conv_b || !grad_ln_g ||
      !grad_ln_b || !grad_head_w || !grad_head_b || !workspace || n_layers < 1 || B < 1 || T < 1 ||
      (k != 1 && k != 3 && k != 5) || B > 65535)
    return fail(EFTS_ERR_ARG, "efts_duration_train_bwd: bad argument");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  Arena a(workspace, workspace_bytes);
  TrainWs w;
  carve_train(a, w, B, T, C, k, true);
  double* ln_part = a.get<double>(static_cast<size_t>(4) * kLnWarps * C);
  if (!a.ok) return fail(EFTS_ERR_WORKSPACE, "workspace too small: need %zu, have %zu", a.off, workspace_bytes);
  const size_t rows = static_cast<size_t>(B) * T, n = rows * C, wn = static_cast<size_t>(k) * C * C;
  CUDA_TRY(cudaMemsetAsync(c->err_flag, 0, sizeof(int), st));
  const float* gh = nullptr;                         // dL/dh of the layer below the one being processed
  for (int l = n_layers - 1; l >= 0; --l) {
    const bool last = l == n_layers - 1;
    float* gu = w.g[0];                              // dL/du before ReLU'
    float* dx = l == 0 ? grad_x : w.g[1];
    ln_train_bwd_kernel<<<kLnWarps / 8, 256, 0, st>>>(gh, last ? grad_out : nullptr, mask, head_w, us + l * n,
                                                      acts + (l + 1) * n, keep ? keep + l * n : nullptr,
                                                      ln_g + static_cast<size_t>(l) * C, rows, gu, ln_part);
    CUDA_TRY(cudaGetLastError());
    ln_train_finish_kernel<<<dim3((C + 127) / 128, 4), 128, 0, st>>>(ln_part, kLnWarps, grad_ln_g + static_cast<size_t>(l) * C,
                                                                    grad_ln_b + static_cast<size_t>(l) * C,
                                                                    last ? grad_head_w : nullptr, last ? grad_head_b : nullptr);
    CUDA_TRY(cudaGetLastError());
    c->launches += 2;
    TRY(conv_layer_bwd(c, st, w, gu, us + l * n, 0.0f, acts + l * n, conv_w + l * wn, k, B, T, nullptr, dx,
                       grad_conv_w + l * wn, grad_conv_b + static_cast<size_t>(l) * C));
    gh = dx;
  }
  return EFTS_OK;
}

// ---- criterion with gradients (losses/fastspeech_loss.py:54-67) ----
size_t efts_fastspeech_loss_workspace_bytes(const efts_ctx* c) {
  return c == nullptr ? 0 : (2 * static_cast<size_t>(kLossBlocks) + 2) * sizeof(double) + 256;
}

int efts_fastspeech_loss(efts_ctx* c, const float* before_outs, const float* d_outs, const float* ys, const float* ds,
                         const int64_t* ilens, const int64_t* olens, int32_t B, int32_t T1, int32_t T2, int32_t odim,
                         int32_t use_masking, int32_t use_mse, float* losses, float* grad_before, float* grad_d,
                         void* workspace, size_t workspace_bytes, void* stream) {
  if (c == nullptr) return fail(EFTS_ERR_ARG, "null context");
  if (!before_outs || !d_outs || !ys || !ds || !ilens || !olens || !losses || !workspace || B < 1 || T1 < 1 || T2 < 1 || odim < 1)
    return fail(EFTS_ERR_ARG, "efts_fastspeech_loss: bad argument");
  if (workspace_bytes < efts_fastspeech_loss_workspace_bytes(c))
    return fail(EFTS_ERR_WORKSPACE, "workspace too small: need %zu, have %zu", efts_fastspeech_loss_workspace_bytes(c), workspace_bytes);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  double* part = reinterpret_cast<double*>((reinterpret_cast<uintptr_t>(workspace) + 255) & ~static_cast<uintptr_t>(255));
  CUDA_TRY(cudaMemsetAsync(c->err_flag, 0, sizeof(int), st));
  const size_t n = static_cast<size_t>(B) * T2 * odim;
  const int blocks = static_cast<int>(std::max<size_t>(1, std::min<size_t>((n + 255) / 256, kLossBlocks)));
  fastspeech_loss_kernel<<<blocks, 256, 0, st>>>(before_outs, ys, d_outs, ds, reinterpret_cast<const long long*>(ilens),
                                                 reinterpret_cast<const long long*>(olens), B, T1, T2, odim, use_masking,
                                                 use_mse, grad_before, grad_d, part, c->err_flag);
  CUDA_TRY(cudaGetLastError());
  fastspeech_loss_finish_kernel<<<1, 32, 0, st>>>(part, blocks, losses);
  CUDA_TRY(cudaGetLastError());
  c->launches += 2;
  return EFTS_OK;
}

int efts_scale_by_scalar(efts_ctx* c, const float* in, const float* scalar, size_t n, float* out, void* stream) {
  if (c == nullptr || !in || !scalar || !out) return fail(EFTS_ERR_ARG, "efts_scale_by_scalar: bad argument");
  if (n == 0) return EFTS_OK;
  scale_by_scalar_kernel<<<ew_grid(n), 256, 0, static_cast<cudaStream_t>(stream)>>>(in, scalar, n, out);
  CUDA_TRY(cudaGetLastError());
  c->launches++;
  return EFTS_OK;
}

}  // extern "C"
