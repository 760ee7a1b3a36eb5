// Item-Item cosine kNN on the GPU -- replaces graphs/build_ii_knn.py:56-99, which computes
// sklearn's cosine_similarity in 1000-row batches on the CPU (O(n^2 d), 78-100 s for 63k items in the reference's report)
// and keeps the k best neighbours of every item.
//
// Four passes:
//   1. prepare   : the reference's two normalisations (x / (|x| + 1e-8), then sklearn's row normalisation) in fp32;
//                  writes the normalised fp32 rows and a bf16 image of them, already laid out the way the UMMA wants
//                  its shared-memory operands (128-row blocks, 128-byte swizzle), so that a whole operand tile is ONE
//                  contiguous 32 KB bulk copy (TMA unit).
//   2. candidates: bf16 tcgen05 GEMM of the image against itself.  A CTA keeps a 256-row block of A resident (128 rows
//                  for 384-d) and streams every 128-row block of B, one 128-wide K chunk at a time, through a 3-deep
//                  ring of bulk copies; accumulators live in TMEM (2 row blocks x 2 stages x 128 columns = all 512
//                  columns); the epilogue warps read them back (tcgen05.ld) and produce, per row, the 48 best APPROXIMATE
//                  similarities: below 8,192 items as a register-resident running list, above as a pipeline of three sweeps
//                  (group maxima -> sampled appends -> full-sweep appends against a per-row threshold, see the modes below).
//   3. re-rank   : exact fp32 dot products for the 48 candidates of every row (warp per row), top-k selection in
//                  descending order, min_similarity filter.  The final similarities are plain fp32 like the reference's;
//                  the bf16 pass only decides WHICH 48 columns get the exact treatment, and a per-row guard lists the
//                  rows where the approximate margin cannot prove the selection exact (dense near-duplicates);
//   4. exact     : those rows are redone with exact fp32 dots against ALL columns.
#include <cuda_bf16.h>
#include <stdlib.h>

#include "common.cuh"
#include "tc_common.cuh"
#include "../../include/b200gat.h"

namespace b200gat {
namespace knn {

using namespace tc;

// The embedding width is KC chunks of 128 (KC = 1: the fused 128-d features, KC = 3: the 384-d text embeddings).
constexpr int kChunkD = 128;
constexpr int kBlk = 128;                       // rows per image block
constexpr int kChunkBytes = kBlk * kChunkD * 2; // 32 KB: [kb 2][128 rows][128 B swizzled]; a block is KC consecutive chunks
constexpr int kCand = 48;                  // approximate candidates kept per row (register-resident list)
constexpr int kBStages = 3;
constexpr int kEpiWarps = 8;
constexpr int kThreads = (kEpiWarps + 2) * 32;   // + MMA warp + producer warp
// The append sweep runs TWO epilogue warps per (row block, TMEM lane quarter): one takes chunks 0-1 of a tile, the other chunks 2-3,
// each with its own half of the row's list.  Its epilogue is bound by instruction issue (two warps per scheduler could not keep
// up with the tensor pipe: 90 ms), not by latency or bytes.
#ifndef KNN_EPI16
#define KNN_EPI16 1    // 0: A/B builds with one epilogue warp per (row block, lane quarter) in the append sweep as well
#endif
constexpr int kAppendHalves = KNN_EPI16 ? 2 : 1;
constexpr int kEpiWarpsAppend = 8 * kAppendHalves;
constexpr int kThreadsAppend = (kEpiWarpsAppend + 2) * 32;
constexpr int kStage = 32;                 // staged (similarity, column) pairs per row between two merges into the register list
constexpr int kStageBytesPerWarp = kStage * 32 * 8;   // [entry][lane] uint2: 8 KB
template <int KC> struct Cfg {
  static constexpr int kNA = KC == 1 ? 2 : 1;                       // A row blocks resident per CTA
  static constexpr int kD = KC * kChunkD;
  static constexpr int kBlkBytes = KC * kChunkBytes;
  static constexpr int kSmem = 1024 + kNA * kBlkBytes + kBStages * kChunkBytes + kNA * 4 * kStageBytesPerWarp + 256;
};
// Bound on |bf16 similarity - exact| used by the safety guard: both operands are rounded to nearest (u = 2^-9), so the
// error is at most (2u + u^2) * sum|a_i b_i| <= 3.91e-3 for unit rows of any width; fp32 accumulation adds ~1e-6.
constexpr float kApproxErr = 4e-3f;

// ---- 1. prepare -----------------------------------------------------------------------------------------------------
template <int KC>
__global__ void __launch_bounds__(128) prepare_kernel(const float* __restrict__ emb, int64_t n, int64_t n_pad,
                                                      float* __restrict__ en, uint8_t* __restrict__ image) {
  constexpr int D = KC * kChunkD;
  const int lane = threadIdx.x & 31;
  const int64_t r = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  if (r >= n_pad) return;
  float4 v[KC];
#pragma unroll
  for (int kc = 0; kc < KC; ++kc) v[kc] = make_float4(0.f, 0.f, 0.f, 0.f);
  if (r < n) {
    float ss = 0.f;
#pragma unroll
    for (int kc = 0; kc < KC; ++kc) {
      v[kc] = ldg4(emb + r * D + kc * kChunkD + lane * 4);
      ss += dot4(v[kc], v[kc]);
    }
    const float s1 = 1.f / (sqrtf(warp_sum(ss)) + 1e-8f);          // embeddings / (norms + 1e-8)      (:56-57)
    ss = 0.f;
#pragma unroll
    for (int kc = 0; kc < KC; ++kc) {
      v[kc].x *= s1; v[kc].y *= s1; v[kc].z *= s1; v[kc].w *= s1;
      ss += dot4(v[kc], v[kc]);
    }
    float n2 = sqrtf(warp_sum(ss));                                // sklearn cosine_similarity normalises again (:76)
    if (n2 == 0.f) n2 = 1.f;
    const float s2 = 1.f / n2;
#pragma unroll
    for (int kc = 0; kc < KC; ++kc) { v[kc].x *= s2; v[kc].y *= s2; v[kc].z *= s2; v[kc].w *= s2; }
  }
  // bf16 image: block = r / 128, then chunk kc, k-block = (lane*4) / 64, 16-byte piece = ((lane*4) % 64) / 8
  const int rb = (int)(r % kBlk), kb = lane >> 4, e = (lane & 15) * 4;   // e: element offset inside the 64-wide k-block
#pragma unroll
  for (int kc = 0; kc < KC; ++kc) {
    *reinterpret_cast<float4*>(en + r * D + kc * kChunkD + lane * 4) = v[kc];
    const __nv_bfloat162 lo = __floats2bfloat162_rn(v[kc].x, v[kc].y), hi = __floats2bfloat162_rn(v[kc].z, v[kc].w);
    uint2 pk;
    pk.x = *reinterpret_cast<const uint32_t*>(&lo);
    pk.y = *reinterpret_cast<const uint32_t*>(&hi);
    uint8_t* dst = image + ((r / kBlk) * KC + kc) * (size_t)kChunkBytes + kb * (kBlk * 128) + sw128(rb, e >> 3) + (e & 7) * 2;
    *reinterpret_cast<uint2*>(dst) = pk;
  }
}

// ---- 2. candidates --------------------------------------------------------------------------------------------------
// The running top-48 of a row lives in REGISTERS as a descending list.  Insertion is a fully unrolled bubble pass
// (compare, conditional swap) -- no local memory, no dependent address chain; lanes that have nothing to insert carry
// v = -inf through the pass and leave their list untouched.
//
// A pass costs the whole warp ~250 instructions whichever of its 32 rows it serves, and a row accepts
// ~kCand * ln(n / kCand) columns over a sweep (444 at 498 k items), almost never in the same chunk as its neighbours: with one
// pass per accepted column the epilogue, not the tensor pipe, bounded the kernel (measured with diagnostic builds at
// 498,196 x 128-d: 45.7 ms when the epilogue only scans -- chunk maximum + vote --, 113.6 ms with the insertions).  Accepted
// columns are therefore only APPENDED to a per-row staging area in shared memory (one predicated 8-byte store each), and
// the lists are brought up to date in batches: when some row of the warp has fewer than 8 free staging slots, every row
// merges what it has staged -- pass t serves entry t of all 32 rows at once.  Between two merges a row's admission
// threshold is stale (too low), which only means a few more appends; nothing a fresh threshold would have kept is lost.
template <int N>
struct TopList {
  float s[N];
  int i[N];
  __device__ __forceinline__ void reset() {
#pragma unroll
    for (int k = 0; k < N; ++k) { s[k] = -INFINITY; i[k] = -1; }
  }
  __device__ __forceinline__ void insert(float v, int col) {
#pragma unroll
    for (int k = 0; k < N; ++k) {
      const bool sw = v > s[k];
      const float ts = s[k];
      const int ti = i[k];
      s[k] = sw ? v : ts;
      i[k] = sw ? col : ti;
      v = sw ? ts : v;
      col = sw ? ti : col;
    }
  }
};

// Modes of the candidates kernel.
//   kModeLists   : the running top-48 per row in registers (above).  Complete on its own; used for catalogues below
//                  kAppendMinBlocks column blocks, and as the fallback of the append pipeline.
// With a running threshold alone a row accepts ~kCand * ln(n / kCand) columns over a sweep (444 at 498 k items) and the
// register lists, not the tensor pipe, bound the kernel (113.6 -> 90.4 ms with staged merges against a 45.7 ms scan-only
// floor).  Larger catalogues therefore go through a pipeline in which the full sweep keeps NO list at all:
//   kModeGroupMax: sweep every kStrideA-th column block; a row keeps the maxima of 64 disjoint column groups (column position
//                  inside the 32-wide chunk x chunk parity: 64 statically indexed registers, 32 FMNMX per chunk).  The 48th
//                  largest of the 64 maxima is attained by 48 different columns, hence a lower bound of the row's 48th best
//                  similarity: thr_A (expected rank among all columns ~ 86 * kStrideA).
//   kModeAppend  : sweep every jstep-th block and APPEND every (similarity, column) with similarity >= thr to the row's list
//                  in global memory (one predicated 8-byte store; a per-lane counter).  First with jstep = kStrideB and
//                  thr_A (~170 appends per row); select_kernel takes the 48th largest of them -- again 48 different columns --
//                  as thr_B (expected rank 48 * kStrideB); then the full sweep with thr_B (<= 384 +- 52 appends per row) and
//                  select_kernel keeps the 48 best as the candidate lists of the re-rank.  Inside an append sweep a row's
//                  threshold climbs a ladder: a rung becomes the threshold once 48 different appended columns have reached it.
// Every threshold is a valid bound whatever the layout of the catalogue (the columns that produced it are swept again by
// the next pass, `>=` re-admits them), so each full-sweep list holds at least 48 entries and the bound handed to the
// re-rank guard, the 48th best approximate similarity, is the same as the register lists'.  Only speed depends on the
// estimates: a row that appends more than kCap entries in the full sweep (a clump of > kCap near-duplicates that no sampled block
// showed) is marked unsafe and redone by the exact path; if more than kMaxOverflowRows rows do, the register-list kernel
// (launched behind a device-side gate, it exits at once otherwise) redoes the sweep for all rows, starting from thr_B.
// Measured at 498,196 x 128-d (profiles/r2ah_knn_append.md): 3.1 + 12.7 + 2.2 (select) + 59.0 + 3.3 (select) ms for the three
// sweeps, 84.7 ms for the whole call against 89.6 ms with the register lists (63,001 items: 3.6 against 4.3 ms); the full sweep
// runs the tensor pipe 51 % active and is bound by the epilogue's ALU/issue rate (136 warp instructions per 32 x 32 chunk).
#ifndef KNN_DIAG
#define KNN_DIAG 0     // diagnostic builds only: 1 the epilogue scans (chunk maximum + vote) but never inserts, 2 it does not read TMEM at all
#endif
constexpr int kModeLists = 0, kModeGroupMax = 1, kModeAppend = 2;
constexpr int kStrideA = 16, kStrideB = 8, kAppendMinBlocks = 64;
constexpr int kCap = 1024;                  // appended entries kept per row (8 KB), one half per epilogue warp of the pair; mean 2 x 192,
                                            // sd 2 x 37 in the full sweep
constexpr int kMaxOverflowRows = 64;

// Order-preserving map float -> uint32 (0 is below every real value: the "empty" key)
__device__ __forceinline__ uint32_t fkey(float f) {
  const uint32_t b = __float_as_uint(f);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float fkey_inv(uint32_t k) { return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k); }

// Eight columns of one row against the row's threshold: every similarity >= thr goes to list[min(cnt, cap - 1)] as (similarity,
// column) and advances cnt.  One PTX block, so that the compares stay where they are (hoisted out of the per-group switch they
// were paid by every chunk with a hit anywhere: the epilogue is bound by instruction issue) and the stores are predicated, not
// branched around.  A full list keeps rewriting its last slot; such rows are recognised by their count.
// Measured alternatives at 498,196 x 128-d (full sweep): C++ with branches 90 ms, predicated stores from C++ 102-120 ms, this block
// 60.8 ms, staging the 8 values in shared memory and letting lanes 0-7 test one column each 60.4 ms.
__device__ __forceinline__ void append8(uint2* list, int& cnt, float thr, int cbase, float x0, float x1, float x2, float x3, float x4,
                                        float x5, float x6, float x7) {
#define KNN_A1(X, T)                                            \
  "setp.ge.f32 p, " X ", %2;\n\t"                               \
  "min.s32 idx, %0, %4;\n\t"                                    \
  "mad.wide.s32 a, idx, 8, %1;\n\t"                             \
  "mov.b32 xb, " X ";\n\t"                                      \
  "add.s32 col, %3, " T ";\n\t"                                 \
  "@p st.global.v2.b32 [a], {xb, col};\n\t"                     \
  "@p add.s32 %0, %0, 1;\n\t"
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .s32 idx;\n\t.reg .b32 xb, col;\n\t.reg .s64 a;\n\t"
      KNN_A1("%5", "0") KNN_A1("%6", "1") KNN_A1("%7", "2") KNN_A1("%8", "3") KNN_A1("%9", "4") KNN_A1("%10", "5") KNN_A1("%11", "6")
      KNN_A1("%12", "7") "}"
      : "+r"(cnt)
      : "l"(list), "f"(thr), "r"(cbase), "n"(kCap / kAppendHalves - 1), "f"(x0), "f"(x1), "f"(x2), "f"(x3), "f"(x4), "f"(x5), "f"(x6), "f"(x7)
      : "memory");
#undef KNN_A1
}

// kModeLists   : sweep from thr_in[row] (may be NULL: -inf; strict >), write the candidate lists and bound_out[row] =
//                max(thr_in[row], 48th entry) >= the approximate similarity of every column that is NOT in the list.
//                gate != NULL: the whole launch is a no-op unless *gate > gate_limit.
// kModeGroupMax: thr_out[row] = 48th largest of the 64 group maxima.
// kModeAppend  : lists[row][0 .. min(cnt, kCap)) = the columns with similarity >= thr_in[row], acc_cnt[row] = cnt (may exceed kCap).
template <int KC, int MODE>
__global__ void __launch_bounds__(MODE == 2 ? kThreadsAppend : kThreads, 1) candidates_kernel(const uint8_t* __restrict__ image, int64_t n, int n_blocks, int jstep,
                                                                 float* __restrict__ cand_sim, int32_t* __restrict__ cand_idx,
                                                                 const float* __restrict__ thr_in, float* __restrict__ thr_out,
                                                                 float* __restrict__ bound_out, uint2* __restrict__ lists,
                                                                 int32_t* __restrict__ acc_cnt, float* __restrict__ dlt,
                                                                 const int32_t* __restrict__ gate, int gate_limit) {
  if (MODE == kModeLists && gate && *gate <= gate_limit) return;      // fallback launch that is not needed (uniform: before any barrier)
  constexpr int kList = kCand;
  constexpr int NA = Cfg<KC>::kNA;
  constexpr int kEpi = MODE == kModeAppend ? kEpiWarpsAppend : kEpiWarps;      // epilogue warps; then the MMA warp, then the producer warp
  constexpr int kHalves = MODE == kModeAppend ? kAppendHalves : 1;
  constexpr int kBlkBytes = Cfg<KC>::kBlkBytes;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* sm = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t sA = base, sB = base + NA * kBlkBytes, sStage = sB + kBStages * kChunkBytes, sBar = sStage + NA * 4 * kStageBytesPerWarp;
  const uint32_t bar_afull = sBar, bar_aempty = sBar + 8, bar_bfull = sBar + 16, bar_bempty = sBar + 48, bar_tfull = sBar + 80,
                 bar_tempty = sBar + 96;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(sm + (sBar - base) + 128);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_super = (n_blocks + NA - 1) / NA;      // NA*128-row super blocks

  if (threadIdx.x == 0) {
    mbar_init(bar_afull, 1);
    mbar_init(bar_aempty, 1);
    for (int i = 0; i < kBStages; ++i) { mbar_init(bar_bfull + 8 * i, 1); mbar_init(bar_bempty + 8 * i, 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(bar_tfull + 8 * i, 1); mbar_init(bar_tempty + 8 * i, NA * 128 * kHalves); }
    fence_barrier_init();
  }
  if (warp == kEpi) tmem_alloc(smem_u32(tmem_ptr_smem), 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp == kEpi + 1) {
    // =============================== producer: bulk copies through the TMA unit ===============================
    if (lane == 0) {
      uint32_t bs = 0, bph = 0, aph = 0;
      for (int sb = blockIdx.x; sb < n_super; sb += gridDim.x) {
        mbar_wait(bar_aempty, aph ^ 1);
        mbar_expect_tx(bar_afull, NA * kBlkBytes);
#pragma unroll
        for (int a = 0; a < NA; ++a) {
          const int ab = min(NA * sb + a, n_blocks - 1);                 // clamps on a ragged tail: those rows are ignored
#pragma unroll
          for (int kc = 0; kc < KC; ++kc)
            bulk_g2s(sA + a * kBlkBytes + kc * kChunkBytes, image + ((size_t)ab * KC + kc) * kChunkBytes, kChunkBytes, bar_afull);
        }
        for (int j = 0; j < n_blocks; j += jstep) {
#pragma unroll 1
          for (int kc = 0; kc < KC; ++kc) {
            mbar_wait(bar_bempty + 8 * bs, bph ^ 1);
            mbar_expect_tx(bar_bfull + 8 * bs, kChunkBytes);
            bulk_g2s(sB + bs * kChunkBytes, image + ((size_t)j * KC + kc) * kChunkBytes, kChunkBytes, bar_bfull + 8 * bs);
            if (++bs == kBStages) { bs = 0; bph ^= 1; }
          }
        }
        aph ^= 1;
      }
    }
  } else if (warp == kEpi) {
    // =============================== MMA issuer ===============================================================
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(128, 128);
      uint32_t bs = 0, bph = 0, ts = 0, tph = 0, aph = 0;
      for (int sb = blockIdx.x; sb < n_super; sb += gridDim.x) {
        mbar_wait(bar_afull, aph);
        for (int j = 0; j < n_blocks; j += jstep) {
          mbar_wait(bar_tempty + 8 * ts, tph ^ 1);
#pragma unroll 1
          for (int kc = 0; kc < KC; ++kc) {
            mbar_wait(bar_bfull + 8 * bs, bph);
            tc_fence_after();
            const uint32_t b0 = sB + bs * kChunkBytes;
#pragma unroll
            for (int a = 0; a < NA; ++a) {
              const uint32_t d = tmem_base + (ts * 2 + a) * 128, a0 = sA + a * kBlkBytes + kc * kChunkBytes;
#pragma unroll
              for (int kb = 0; kb < 2; ++kb)
#pragma unroll
                for (int k = 0; k < 4; ++k)
                  umma_bf16(d, make_desc(a0 + kb * (kBlk * 128) + k * 32, 16, 1024),
                            make_desc(b0 + kb * (kBlk * 128) + k * 32, 16, 1024), idesc, (kc | kb | k) != 0);
            }
            umma_commit(bar_bempty + 8 * bs);
            if (++bs == kBStages) { bs = 0; bph ^= 1; }
          }
          umma_commit(bar_tfull + 8 * ts);
          if (++ts == 2) { ts = 0; tph ^= 1; }
        }
        umma_commit(bar_aempty);
        aph ^= 1;
      }
    }
    __syncwarp();
  } else if ((warp >> 2) < NA * kHalves) {
    // =============================== epilogue =====================================================================
    const int half = warp / (4 * NA);             // append sweep: which two chunks of a tile (and which half of the list) this warp owns
    const int a = (warp >> 2) % NA, q = warp & 3; // A block of the super block, TMEM lane quarter
    const int ni = (int)n;                        // n < 2^31 (checked by the entry point)
    uint32_t ts = 0, tph = 0;
    // one 32-column chunk of this row's similarities; the diagonal and the columns past the last item become -inf (only the row
    // block's own column block and the last block can contain them: warp-uniform test)
    auto load_chunk = [&](float(&v)[32], int j, int c, int row, bool edge) {
      tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (ts * 2 + a) * 128 + c * 32, v);
      if (edge) {
        const int col0 = j * kBlk + c * 32;
#pragma unroll
        for (int t = 0; t < 32; ++t)
          if (col0 + t == row || col0 + t >= ni) v[t] = -INFINITY;
      }
    };
    // maxima of the four 8-column groups of a chunk
    auto group_max = [](const float(&v)[32], float(&g)[4]) {
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        g[u] = fmaxf(fmaxf(fmaxf(v[8 * u], v[8 * u + 1]), fmaxf(v[8 * u + 2], v[8 * u + 3])),
                     fmaxf(fmaxf(v[8 * u + 4], v[8 * u + 5]), fmaxf(v[8 * u + 6], v[8 * u + 7])));
      }
    };

    if constexpr (MODE == kModeGroupMax) {
      // ---------- maxima of 64 disjoint column groups per row; thr_out = the 48th largest of them
      for (int sb = blockIdx.x; sb < n_super; sb += gridDim.x) {
        const int row = (NA * sb + a) * kBlk + q * 32 + lane;
        const bool live = (NA * sb + a) < n_blocks && row < ni;
        float gA[32], gB[32];
#pragma unroll
        for (int t = 0; t < 32; ++t) { gA[t] = -INFINITY; gB[t] = -INFINITY; }
        for (int j = 0; j < n_blocks; j += jstep) {
          mbar_wait(bar_tfull + 8 * ts, tph);
          tc_fence_after();
          const bool edge = j == NA * sb + a || (j + 1) * kBlk > ni;
#pragma unroll 1
          for (int c2 = 0; c2 < 2; ++c2) {
            float v[32];
            load_chunk(v, j, 2 * c2, row, edge);
#pragma unroll
            for (int t = 0; t < 32; ++t) gA[t] = fmaxf(gA[t], v[t]);
            load_chunk(v, j, 2 * c2 + 1, row, edge);
#pragma unroll
            for (int t = 0; t < 32; ++t) gB[t] = fmaxf(gB[t], v[t]);
          }
          tc_fence_before();
          mbar_arrive(bar_tempty + 8 * ts);
          if (++ts == 2) { ts = 0; tph ^= 1; }
        }
        TopList<kCand> top;
        top.reset();
#pragma unroll 1
        for (int half = 0; half < 2; ++half) {
#pragma unroll
          for (int t = 0; t < 32; ++t) top.insert(half ? gB[t] : gA[t], t);
        }
        if (live) {
          thr_out[row] = top.s[kCand - 1];
          dlt[row] = fmaxf(top.s[kCand / 2 - 1] - top.s[kCand - 1], 1e-6f);   // ladder step of the next sweep: ~ a factor 2 in rank
        }
      }
    } else if constexpr (MODE == kModeAppend) {
      // ---------- no list: every column with similarity >= thr is appended to the row's list in global memory
      for (int sb = blockIdx.x; sb < n_super; sb += gridDim.x) {
        const int row = (NA * sb + a) * kBlk + q * 32 + lane;
        const bool live = (NA * sb + a) < n_blocks && row < ni;
        float thr = live ? thr_in[row] : INFINITY;                     // padding rows never append
        // Ladder: thr2 < thr3 are the next two rungs; cnt2 / cnt3 count the (chunk, 8-column group) cells of this row seen so far
        // whose maximum reached them -- that many DIFFERENT columns, all of them appended.  48 of them make the rung a valid
        // threshold for the rest of the sweep (~48 (1 + ln 8) appends per row instead of 384).
        const float dl = live ? dlt[row] : 0.f;
        float thr2 = thr + dl, thr3 = thr2 + dl;
        int cnt2 = 0, cnt3 = 0;
        uint2* lp = lists + (size_t)(live ? row : 0) * kCap + half * (kCap / kAppendHalves);
        int cnt = 0;
        for (int j = 0; j < n_blocks; j += jstep) {
          mbar_wait(bar_tfull + 8 * ts, tph);
          tc_fence_after();
          const bool edge = j == NA * sb + a || (j + 1) * kBlk > ni;
#pragma unroll 1
          for (int c = (4 / kHalves) * half; c < (4 / kHalves) * (half + 1); ++c) {
            float v[32];
            load_chunk(v, j, c, row, edge);
            float g[4];
            group_max(v, g);
            const float vmax = fmaxf(fmaxf(g[0], g[1]), fmaxf(g[2], g[3]));
            if (!__any_sync(kFull, vmax >= thr)) continue;          // nothing in this chunk reaches any row's threshold
            unsigned gmask = __reduce_or_sync(kFull, (g[0] >= thr ? 1u : 0u) | (g[1] >= thr ? 2u : 0u) | (g[2] >= thr ? 4u : 0u) | (g[3] >= thr ? 8u : 0u));
            const int col0 = j * kBlk + c * 32;
            while (gmask) {
              const int u = __ffs(gmask) - 1;
              gmask &= gmask - 1;
              const int cbase = col0 + 8 * u;
#define KNN_APPEND8(U) append8(lp, cnt, thr, cbase, v[8 * (U)], v[8 * (U) + 1], v[8 * (U) + 2], v[8 * (U) + 3], v[8 * (U) + 4], \
                               v[8 * (U) + 5], v[8 * (U) + 6], v[8 * (U) + 7]);
              switch (u) {
                case 0: KNN_APPEND8(0) break;
                case 1: KNN_APPEND8(1) break;
                case 2: KNN_APPEND8(2) break;
                default: KNN_APPEND8(3) break;
              }
#undef KNN_APPEND8
            }
            cnt2 += (g[0] >= thr2 ? 1 : 0) + (g[1] >= thr2 ? 1 : 0) + (g[2] >= thr2 ? 1 : 0) + (g[3] >= thr2 ? 1 : 0);
            cnt3 += (g[0] >= thr3 ? 1 : 0) + (g[1] >= thr3 ? 1 : 0) + (g[2] >= thr3 ? 1 : 0) + (g[3] >= thr3 ? 1 : 0);
            if (cnt2 >= kCand) { thr = thr2; thr2 = thr3; cnt2 = cnt3; thr3 += dl; cnt3 = 0; }
          }
          tc_fence_before();
          mbar_arrive(bar_tempty + 8 * ts);
          if (++ts == 2) { ts = 0; tph ^= 1; }
        }
        if (live) {
          acc_cnt[2 * row + half] = cnt;
          if (kHalves == 1) acc_cnt[2 * row + 1] = 0;
        }
      }
    } else {
      // ---------- running top-48 per row in registers
      TopList<kList> top;
      const uint32_t stage0 = sStage + warp * kStageBytesPerWarp + lane * 8;     // entry k of this row: stage0 + k * 256
      for (int sb = blockIdx.x; sb < n_super; sb += gridDim.x) {
        const int row = (NA * sb + a) * kBlk + q * 32 + lane;
        const bool live = (NA * sb + a) < n_blocks && row < ni;
        top.reset();
        const float thr0 = live ? (thr_in ? thr_in[row] : -INFINITY) : INFINITY;      // padding rows never append
        float thr = thr0;
        int cnt = 0;                                  // staged entries of this row
        // merge the staged entries of all 32 rows into their lists: pass k serves entry k of every row that has one
        auto flush = [&]() {
          const int most = __reduce_max_sync(kFull, cnt);
          for (int k = 0; k < most; ++k) {
            float sv = -INFINITY;
            int sc = -1;
            if (k < cnt) {
              const uint2 e = ld_shared_u2(stage0 + k * 256);
              sv = __uint_as_float(e.x);
              sc = (int)e.y;
            }
            if (!__any_sync(kFull, sv > top.s[kList - 1])) continue;     // stale threshold: nothing of this pass still qualifies
            top.insert(sv, sc);
          }
          cnt = 0;
          if (live) thr = fmaxf(thr0, top.s[kList - 1]);
        };
        for (int j = 0; j < n_blocks; j += jstep) {
          mbar_wait(bar_tfull + 8 * ts, tph);
          tc_fence_after();
          const bool edge = j == NA * sb + a || (j + 1) * kBlk > ni;
#pragma unroll 1
          for (int c = 0; c < 4; ++c) {
            if (KNN_DIAG & 2) continue;
            float v[32];
            load_chunk(v, j, c, row, edge);
            const int col0 = j * kBlk + c * 32;
            float g[4];
            group_max(v, g);
            const float vmax = fmaxf(fmaxf(g[0], g[1]), fmaxf(g[2], g[3]));
            if (!__any_sync(kFull, vmax > thr)) continue;          // nothing in this chunk beats any row's threshold
            if (KNN_DIAG & 1) { if (vmax == 12345.f) thr = 0.f; continue; }
            // groups in which some row has a hit (warp-uniform mask); a group appends at most 8 entries per row
            unsigned gmask = __reduce_or_sync(kFull, (g[0] > thr ? 1u : 0u) | (g[1] > thr ? 2u : 0u) | (g[2] > thr ? 4u : 0u) | (g[3] > thr ? 8u : 0u));
            while (gmask) {
              const int u = __ffs(gmask) - 1;
              gmask &= gmask - 1;
              if (__any_sync(kFull, cnt > kStage - 8)) flush();     // single call site: the merge is ~1.5 KB of code
              const int cbase = col0 + 8 * u;
#define KNN_APPEND8(U)                                                                                     \
  _Pragma("unroll") for (int t = 0; t < 8; ++t) {                                                          \
    const float x = v[8 * (U) + t];                                                                        \
    const int hit = x > thr ? 1 : 0;                                                                       \
    st_shared_u2_if(hit, stage0 + cnt * 256, __float_as_uint(x), (uint32_t)(cbase + t));                  \
    cnt += hit;                                                                                            \
  }
              switch (u) {
                case 0: KNN_APPEND8(0) break;
                case 1: KNN_APPEND8(1) break;
                case 2: KNN_APPEND8(2) break;
                default: KNN_APPEND8(3) break;
              }
#undef KNN_APPEND8
            }
          }
          tc_fence_before();
          mbar_arrive(bar_tempty + 8 * ts);
          if (++ts == 2) { ts = 0; tph ^= 1; }
        }
        flush();
        if (live) {
#pragma unroll
          for (int k = 0; k < kList; ++k) {
            cand_sim[(int64_t)row * kCand + k] = top.s[k];
            cand_idx[(int64_t)row * kCand + k] = top.i[k];
          }
          bound_out[row] = fmaxf(thr0, top.s[kList - 1]);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == kEpi) tmem_dealloc(tmem_base, 512);
}

// ---- 2b. selection over the appended lists ----------------------------------------------------------------------------
// Warp per row.  The (at most kCap) appended entries sit in registers, 32 per lane, as order-preserving integer keys; the
// 48th largest key is found bit by bit (the largest T with |{key >= T}| >= 48: 32 counting rounds).
//   FINAL = false: thr[row] = that value (the admission threshold of the next sweep).  Fewer than 48 entries: thr stays.
//   FINAL = true : the 48 best entries become the row's candidate list (unordered: the re-rank does not care), bound[row] =
//                  the 48th best (no column outside the list has a larger approximate similarity), or thr[row] if the row
//                  appended fewer than 48, or +inf if it appended more than kCap (entries were dropped: the guard then
//                  sends the row to the exact path); n_over counts those rows.
template <bool FINAL>
__global__ void __launch_bounds__(128) select_kernel(const uint2* __restrict__ lists, const int32_t* __restrict__ acc_cnt, int64_t n,
                                                     float* __restrict__ thr, float* __restrict__ dlt, float* __restrict__ cand_sim,
                                                     int32_t* __restrict__ cand_idx, float* __restrict__ bound,
                                                     int32_t* __restrict__ n_over) {
  constexpr int kSlots = kCap / 32;
  static_assert(kCap % 32 == 0 && kCand > 32 && kCand <= 64, "select_kernel lane mapping");
  const int lane = threadIdx.x & 31;
  const int64_t r = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  if (r >= n) return;
  // the two halves of the list (one per epilogue warp of the pair) sit in slots [0, kSlots/2) and [kSlots/2, kSlots)
  constexpr int kHalfCap = kCap / kAppendHalves, kHalfSlots = kSlots / kAppendHalves;
  const int c_raw0 = acc_cnt[2 * r], c_raw1 = acc_cnt[2 * r + 1];
  const bool over = c_raw0 > kHalfCap || c_raw1 > kHalfCap;
  const int c0 = min(c_raw0, kHalfCap), c1 = min(c_raw1, kHalfCap);
  uint32_t key[kSlots];
  int col[kSlots];
#pragma unroll
  for (int s = 0; s < kSlots; ++s) {
    key[s] = 0u;
    col[s] = -1;
    const int e = (s % kHalfSlots) * 32 + lane;
    if (e < (s < kHalfSlots ? c0 : c1)) {
      const uint2 x = lists[r * kCap + (s < kHalfSlots ? 0 : kHalfCap) + e];
      key[s] = fkey(__uint_as_float(x.x));
      col[s] = (int)x.y;
    }
  }
  const int c = c0 + c1;
  // warp-uniform: does slot s hold any entry
  auto slot_live = [&](int s) { return (s % kHalfSlots) * 32 < (s < kHalfSlots ? c0 : c1); };
  uint32_t T = 0u;         // fewer than 48 entries (cannot happen after a sweep that re-admits the 48 columns behind its threshold): keep all
  if (c >= kCand) {
#pragma unroll 1
    for (int bit = 31; bit >= 0; --bit) {
      const uint32_t tr = T | (1u << bit);
      int local = 0;
#pragma unroll
      for (int s = 0; s < kSlots; ++s)
        if (slot_live(s)) local += key[s] >= tr ? 1 : 0;
      if (__reduce_add_sync(kFull, local) >= kCand) T = tr;
    }
  }
  if (!FINAL) {
    if (lane == 0 && c >= kCand) {
      const float t_new = fkey_inv(T), t_old = thr[r];
      thr[r] = t_new;
      dlt[r] = fmaxf(0.5f * (t_new - t_old), 1e-6f);      // the two thresholds are a factor ~3.6 apart in rank: ~1.9 per rung
    }
    return;
  }
  // keys above T (fewer than 48 of them), then keys equal to T until the list is full, then padding
  const unsigned lt = (1u << lane) - 1u;
  int base = 0;
#pragma unroll
  for (int s = 0; s < kSlots; ++s) {
    if (slot_live(s)) {
      const bool p = key[s] > T;
      const unsigned m = __ballot_sync(kFull, p);
      const int pos = base + __popc(m & lt);
      if (p) { cand_sim[r * kCand + pos] = fkey_inv(key[s]); cand_idx[r * kCand + pos] = col[s]; }
      base += __popc(m);
    }
  }
  if (T != 0u) {
#pragma unroll
    for (int s = 0; s < kSlots; ++s) {
      if (slot_live(s) && base < kCand) {
        const bool p = key[s] == T;
        const unsigned m = __ballot_sync(kFull, p);
        const int pos = base + __popc(m & lt);
        if (p && pos < kCand) { cand_sim[r * kCand + pos] = fkey_inv(key[s]); cand_idx[r * kCand + pos] = col[s]; }
        base += __popc(m);
      }
    }
  }
  for (int pos = min(base, kCand) + lane; pos < kCand; pos += 32) { cand_sim[r * kCand + pos] = -INFINITY; cand_idx[r * kCand + pos] = -1; }
  if (lane == 0) {
    bound[r] = over ? INFINITY : (c >= kCand ? fkey_inv(T) : thr[r]);
    if (over) atomicAdd(n_over, 1);
  }
}

// ---- 3. exact re-rank ------------------------------------------------------------------------------------------------
template <int KC>
__global__ void __launch_bounds__(128) rerank_kernel(const float* __restrict__ en, int64_t n, const float* __restrict__ cand_sim,
                                                     const int32_t* __restrict__ cand_idx, int k, float min_similarity,
                                                     int32_t* __restrict__ nbr_idx, float* __restrict__ nbr_sim,
                                                     int32_t* __restrict__ counts, int32_t* __restrict__ n_unsafe,
                                                     int32_t* __restrict__ unsafe_rows, const float* __restrict__ bound) {
  constexpr int D = KC * kChunkD;
  const int lane = threadIdx.x & 31;
  const int64_t r = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  if (r >= n) return;
  float4 u[KC];
#pragma unroll
  for (int kc = 0; kc < KC; ++kc) u[kc] = ldg4(en + r * D + kc * kChunkD + lane * 4);
  auto dot_row = [&](int64_t c) {
    float acc = 0.f;
#pragma unroll
    for (int kc = 0; kc < KC; ++kc) acc += dot4(u[kc], ldg4(en + c * D + kc * kChunkD + lane * 4));
    return acc;
  };
  // lane l owns candidates l and l + 32 (kCand <= 64)
  static_assert(kCand > 32 && kCand <= 64, "rerank lane mapping");
  int id[2] = {cand_idx[r * kCand + lane], (32 + lane < kCand) ? cand_idx[r * kCand + 32 + lane] : -1};
  float ex[2] = {-INFINITY, -INFINITY};
#pragma unroll
  for (int half = 0; half < 2; ++half) {
    for (int c0 = 0; c0 < 32; c0 += 4) {
      float d[4];
      int cid[4];
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        cid[t] = __shfl_sync(kFull, id[half], c0 + t);
        d[t] = cid[t] >= 0 ? dot_row(cid[t]) : 0.f;
      }
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const float s = warp_sum(d[t]);
        if (lane == c0 + t && cid[t] >= 0) ex[half] = s;
      }
    }
  }
  const float approx_last = bound[r];   // no column outside the list has a larger approximate similarity (see candidates_kernel)
  // k rounds of warp arg-max over the 64 exact similarities (ties: lower candidate slot first)
  float kth = -INFINITY;
  int valid = 0;
  for (int round = 0; round < k; ++round) {
    float best = ex[0];
    int slot = lane;
    if (ex[1] > best) { best = ex[1]; slot = lane + 32; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ob = __shfl_xor_sync(kFull, best, o);
      const int os = __shfl_xor_sync(kFull, slot, o);
      if (ob > best || (ob == best && os < slot)) { best = ob; slot = os; }
    }
    const int wid = __shfl_sync(kFull, id[slot >> 5], slot & 31);
    if (lane == (slot & 31)) ex[slot >> 5] = -INFINITY;           // consumed
    const bool ok = best > -INFINITY && best >= min_similarity;  // the list is descending: once false, always false
    if (lane == 0) {
      nbr_sim[r * k + round] = best > -INFINITY ? best : 0.f;
      nbr_idx[r * k + round] = best > -INFINITY ? wid : -1;
    }
    valid += ok ? 1 : 0;
    kth = best;
  }
  if (lane == 0) {
    counts[r] = valid;
    // guard: a column outside the 48 candidates has approximate similarity <= approx_last, hence exact similarity
    // <= approx_last + err; the selection is provably the exact top-k iff the exact k-th beats that bound
    // (kth == -inf: fewer than k candidates cleared the sampled threshold -- columns below it may still belong to the top-k)
    if (n - 1 > kCand && !(kth > approx_last + kApproxErr)) unsafe_rows[atomicAdd(n_unsafe, 1)] = (int32_t)r;
  }
}


// ---- 4. exact path for the rows the guard could not clear -------------------------------------------------------------
// One block per such row (grid-stride over the list): every warp scans a stripe of ALL columns with exact fp32 dots and
// keeps its own top-k (lane l holds entry l, replacement of the current minimum); the 8 per-warp lists are merged by
// warp 0.  255 MB of reads per row at 498k items, so this is only meant for the few rows with dense near-duplicates.
template <int KC>
__global__ void __launch_bounds__(256) exact_rows_kernel(const float* __restrict__ en, int64_t n, const int32_t* __restrict__ unsafe_rows,
                                                         const int32_t* __restrict__ n_unsafe, int k, float min_similarity,
                                                         int32_t* __restrict__ nbr_idx, float* __restrict__ nbr_sim,
                                                         int32_t* __restrict__ counts) {
  __shared__ float s_sim[8][32];
  __shared__ int s_idx[8][32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int total = *n_unsafe;
  for (int u = blockIdx.x; u < total; u += gridDim.x) {
    constexpr int D = KC * kChunkD;
    const int64_t r = unsafe_rows[u];
    float4 q[KC];
#pragma unroll
    for (int kc = 0; kc < KC; ++kc) q[kc] = ldg4(en + r * D + kc * kChunkD + lane * 4);
    float my = -INFINITY;       // lane l < k holds one entry of this warp's top-k
    int my_i = -1;
    float wmin = -INFINITY;     // smallest entry of the warp's list (valid once the list is full)
    int filled = 0;
    for (int64_t c0 = warp; c0 < n; c0 += 8 * 4) {
      float d[4];
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const int64_t c = c0 + 8 * t;
        float acc = 0.f;
        if (c < n && c != r) {
#pragma unroll
          for (int kc = 0; kc < KC; ++kc) acc += dot4(q[kc], ldg4(en + c * D + kc * kChunkD + lane * 4));
        }
        d[t] = acc;
      }
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const int64_t c = c0 + 8 * t;
        const float sc = warp_sum(d[t]);
        if (c >= n || c == r) continue;                       // warp-uniform
        if (filled < k) {
          if (lane == filled) { my = sc; my_i = (int)c; }
          ++filled;
          if (filled == k) {                                  // list full: find its minimum
            float m = lane < k ? my : INFINITY;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) m = fminf(m, __shfl_xor_sync(kFull, m, o));
            wmin = m;
          }
        } else if (sc > wmin) {
          // replace (one of) the minimum entries, then recompute the minimum
          const unsigned holders = __ballot_sync(kFull, lane < k && my == wmin);
          if (lane == __ffs(holders) - 1) { my = sc; my_i = (int)c; }
          float m = lane < k ? my : INFINITY;
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) m = fminf(m, __shfl_xor_sync(kFull, m, o));
          wmin = m;
        }
      }
    }
    s_sim[warp][lane] = lane < k ? my : -INFINITY;
    s_idx[warp][lane] = lane < k ? my_i : -1;
    __syncthreads();
    if (warp == 0) {
      // 8 lists x 32 slots = 256 candidates, 8 per lane; k rounds of arg-max (ties: lower column index first)
      float cs[8];
      int ci[8];
#pragma unroll
      for (int w = 0; w < 8; ++w) { cs[w] = s_sim[w][lane]; ci[w] = s_idx[w][lane]; }
      int valid = 0;
      for (int round = 0; round < k; ++round) {
        float best = -INFINITY;
        int bi = 0x7fffffff, bw = 0;
#pragma unroll
        for (int w = 0; w < 8; ++w)
          if (cs[w] > best || (cs[w] == best && ci[w] >= 0 && ci[w] < bi)) { best = cs[w]; bi = ci[w] >= 0 ? ci[w] : 0x7fffffff; bw = w; }
        float gb = best;
        int gi = bi, gl = lane;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          const float ob = __shfl_xor_sync(kFull, gb, o);
          const int oi = __shfl_xor_sync(kFull, gi, o), ol = __shfl_xor_sync(kFull, gl, o);
          if (ob > gb || (ob == gb && oi < gi)) { gb = ob; gi = oi; gl = ol; }
        }
        if (lane == gl) {
#pragma unroll
          for (int w = 0; w < 8; ++w)
            if (w == bw) cs[w] = -INFINITY;                   // consumed
        }
        if (lane == 0) {
          nbr_sim[r * k + round] = gb > -INFINITY ? gb : 0.f;
          nbr_idx[r * k + round] = gb > -INFINITY ? gi : -1;
        }
        valid += (gb > -INFINITY && gb >= min_similarity) ? 1 : 0;
      }
      if (lane == 0) counts[r] = valid;
    }
    __syncthreads();
  }
}

}  // namespace knn
}  // namespace b200gat

using namespace b200gat;

template <int KC>
static int knn_run(const float* emb, int64_t n_items, int k, float min_similarity, int32_t* nbr_idx, float* nbr_sim, int32_t* counts,
                   int32_t* n_unsafe, void* workspace, cudaStream_t st) {
  using C = knn::Cfg<KC>;
  static DeviceOnce once;        // one per KC instantiation
  if (once.pending()) {
    B200GAT_CUDA(cudaFuncSetAttribute(knn::candidates_kernel<KC, knn::kModeLists>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::kSmem));
    B200GAT_CUDA(cudaFuncSetAttribute(knn::candidates_kernel<KC, knn::kModeGroupMax>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::kSmem));
    B200GAT_CUDA(cudaFuncSetAttribute(knn::candidates_kernel<KC, knn::kModeAppend>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::kSmem));
    once.done();
  }
  const int64_t n_pad = (n_items + knn::kBlk - 1) / knn::kBlk * knn::kBlk;
  const int n_blocks = (int)(n_pad / knn::kBlk);
  char* p = (char*)workspace;
  float* en = (float*)p;                                p += (size_t)n_pad * C::kD * 4;
  uint8_t* image = (uint8_t*)p;                         p += (size_t)n_pad * C::kD * 2;
  float* cand_sim = (float*)p;                          p += (size_t)n_pad * knn::kCand * 4;
  int32_t* cand_idx = (int32_t*)p;                      p += (size_t)n_pad * knn::kCand * 4;
  int32_t* unsafe_rows = (int32_t*)p;                   p += (size_t)n_pad * 4;
  float* thr = (float*)p;                               p += (size_t)n_pad * 4;
  float* bound = (float*)p;                             p += (size_t)n_pad * 4;
  int32_t* acc_cnt = (int32_t*)p;                       p += (size_t)n_pad * 8;      // one counter per half list
  int32_t* n_over = (int32_t*)p;                        p += 256;
  float* dlt = (float*)p;                               p += (size_t)n_pad * 4;
  uint2* lists = (uint2*)p;                             // [n_pad][kCap], only laid out for >= kAppendMinBlocks column blocks
  // B200GAT_KNN_MODE=lists (read per call) forces the register-list kernel at every size (A/B measurements, tests)
  const char* env = getenv("B200GAT_KNN_MODE");
  const bool append = n_blocks >= knn::kAppendMinBlocks && !(env && env[0] == 'l');
  // B200GAT_KNN_MAX_OVERFLOW (tests): how many overflowed rows the exact path may take before the register-list kernel redoes the sweep
  const char* env_o = getenv("B200GAT_KNN_MAX_OVERFLOW");
  const int max_over = env_o ? atoi(env_o) : knn::kMaxOverflowRows;
  count_launch(), knn::prepare_kernel<KC><<<ceil_div(n_pad * 32, 128), 128, 0, st>>>(emb, n_items, n_pad, en, image);
  const int n_super = (n_blocks + C::kNA - 1) / C::kNA;
  const int grid = n_super < kNumSMs ? n_super : kNumSMs;
  const int sel_grid = ceil_div(n_items * 32, 128);
  if (append) {
    B200GAT_CUDA(cudaMemsetAsync(n_over, 0, sizeof(int32_t), st));
    // thr_A: 48th largest of 64 group maxima over every 16th column block
    count_launch(), knn::candidates_kernel<KC, knn::kModeGroupMax><<<grid, knn::kThreads, C::kSmem, st>>>(
        image, n_items, n_blocks, knn::kStrideA, nullptr, nullptr, nullptr, thr, nullptr, nullptr, nullptr, dlt, nullptr, 0);
    // thr_B: 48th largest of the columns >= thr_A among every 8th block
    count_launch(), knn::candidates_kernel<KC, knn::kModeAppend><<<grid, knn::kThreadsAppend, C::kSmem, st>>>(
        image, n_items, n_blocks, knn::kStrideB, nullptr, nullptr, thr, nullptr, nullptr, lists, acc_cnt, dlt, nullptr, 0);
    count_launch(), knn::select_kernel<false><<<sel_grid, 128, 0, st>>>(lists, acc_cnt, n_items, thr, dlt, nullptr, nullptr, nullptr, nullptr);
    // the full sweep: every column >= thr_B, then the 48 best of them
    count_launch(), knn::candidates_kernel<KC, knn::kModeAppend><<<grid, knn::kThreadsAppend, C::kSmem, st>>>(
        image, n_items, n_blocks, 1, nullptr, nullptr, thr, nullptr, nullptr, lists, acc_cnt, dlt, nullptr, 0);
    count_launch(), knn::select_kernel<true><<<sel_grid, 128, 0, st>>>(lists, acc_cnt, n_items, thr, dlt, cand_sim, cand_idx, bound, n_over);
    // fallback behind a device-side gate: too many rows overflowed their lists -> register lists for all rows, from thr_B
    count_launch(), knn::candidates_kernel<KC, knn::kModeLists><<<grid, knn::kThreads, C::kSmem, st>>>(
        image, n_items, n_blocks, 1, cand_sim, cand_idx, thr, nullptr, bound, nullptr, nullptr, nullptr, n_over, max_over);
  } else {
    count_launch(), knn::candidates_kernel<KC, knn::kModeLists><<<grid, knn::kThreads, C::kSmem, st>>>(
        image, n_items, n_blocks, 1, cand_sim, cand_idx, nullptr, nullptr, bound, nullptr, nullptr, nullptr, nullptr, 0);
  }
#if KNN_DIAG == 0
  count_launch(), knn::rerank_kernel<KC><<<ceil_div(n_items * 32, 128), 128, 0, st>>>(en, n_items, cand_sim, cand_idx, k, min_similarity,
                                                                                      nbr_idx, nbr_sim, counts, n_unsafe, unsafe_rows,
                                                                                      bound);
  // rows whose bf16 margin was too thin: exact scan of all columns (no-op when the list is empty)
  count_launch(), knn::exact_rows_kernel<KC><<<kNumSMs * 2, 256, 0, st>>>(en, n_items, unsafe_rows, n_unsafe, k, min_similarity, nbr_idx,
                                                                         nbr_sim, counts);
#endif
  B200GAT_LAUNCH_CHECK();
  return kOk;
}

extern "C" int b200gat_knn_workspace_bytes(int64_t n_items, int dim, size_t* bytes) {
  B200GAT_CHECK_ARG(bytes && n_items >= 0, "bad args");
  if (dim != 128 && dim != 384) {
    set_error("cosine kNN: embedding width %d is not supported on the tensor-core path (128 or 384)", dim);
    return kErrUnsupported;
  }
  const int64_t n_pad = (n_items + knn::kBlk - 1) / knn::kBlk * knn::kBlk;
  *bytes = (size_t)n_pad * dim * 4 /*en*/ + (size_t)n_pad * dim * 2 /*image*/ + (size_t)n_pad * knn::kCand * 8 /*cands*/ +
           (size_t)n_pad * 24 /*unsafe row list, thresholds, ladder steps, bounds, append counters (2)*/ + 256 /*overflow counter*/ + 1024;
  if (n_pad / knn::kBlk >= knn::kAppendMinBlocks) *bytes += (size_t)n_pad * knn::kCap * 8;   // appended (similarity, column) lists
  return kOk;
}

// nbr_idx / nbr_sim [n, k]: the k most similar other items of every item, descending (unused slots: -1 / 0);
// counts[n]: how many of them pass `>= min_similarity` (they are a prefix).  n_unsafe: device int32, number of rows
// whose bf16 candidate margin was too thin to prove exactness; those rows were recomputed by the exact path.
extern "C" int b200gat_knn_cosine_f32(const float* emb, int64_t n_items, int dim, int k, float min_similarity,
                                      int32_t* nbr_idx, float* nbr_sim, int32_t* counts, int32_t* n_unsafe, void* workspace,
                                      size_t workspace_bytes, void* stream) {
  B200GAT_CHECK_ARG(emb && nbr_idx && nbr_sim && counts && n_unsafe && workspace, "null pointer");
  B200GAT_CHECK_ARG(k >= 1 && k <= 32, "k=%d must be in [1, 32]", k);
  size_t need;
  int rc = b200gat_knn_workspace_bytes(n_items, dim, &need);
  if (rc) return rc;
  B200GAT_CHECK_ARG(workspace_bytes >= need, "workspace too small: %zu < %zu", workspace_bytes, need);
  B200GAT_CHECK_ARG(n_items < 2147483647LL, "too many items");
  cudaStream_t st = (cudaStream_t)stream;
  B200GAT_CUDA(cudaMemsetAsync(n_unsafe, 0, sizeof(int32_t), st));
  if (n_items == 0) return kOk;
  if (dim == 128) return knn_run<1>(emb, n_items, k, min_similarity, nbr_idx, nbr_sim, counts, n_unsafe, workspace, st);
  return knn_run<3>(emb, n_items, k, min_similarity, nbr_idx, nbr_sim, counts, n_unsafe, workspace, st);
}
