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
//                  columns); 8 epilogue warps read them back (tcgen05.ld) and keep, per row, the 48 best APPROXIMATE
//                  similarities seen so far in a register-resident list.
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

// Sampled admission threshold (opt-in, see knn_run).  With a running threshold alone a row accepts ~kCand * ln(n / kCand) columns over a sweep
// (444 at 498 k items), and every acceptance of any of a warp's 32 rows costs the whole warp one insertion pass: the
// epilogue, not the tensor pipe, bounded the kernel (tensor pipe 24 % active).  A first sweep over every kSampleStride-th
// column block keeps only the kSampleKeep best per row; the kSampleKeep-th best of that SUBSET is a valid lower bound of the
// row's kSampleKeep-th best overall, and in expectation its rank among all columns is kSampleKeep * kSampleStride = 192.
// The full sweep starts from that threshold: ~48 * (1 + ln 4) = 115 acceptances instead of 444.  Correctness does not rest
// on the estimate: columns outside the final list have approximate similarity <= max(threshold, 48th entry), which is the
// bound the re-rank guard uses; a row whose exact k-th neighbour does not clear it is redone exactly (that needs the
// threshold to rank above ~30 among all columns: P < 1e-7 for a random sample, certain only for adversarial layouts).
#ifndef KNN_DIAG
#define KNN_DIAG 0     // diagnostic builds only: 1 the epilogue scans (chunk maximum + vote) but never inserts, 2 it does not read TMEM at all
#endif
constexpr int kSampleStride = 16, kSampleKeep = 12, kSampleMinBlocks = 64;

// SAMPLE: sweep the column blocks j0, j0 + kSampleStride, ... and write thr_out[row] = the kSampleKeep-th best similarity seen.
// otherwise: sweep all column blocks starting from thr_in[row] (may be NULL: -inf), write the candidate lists and
// bound_out[row] = max(thr_in[row], 48th entry) >= the approximate similarity of every column that is NOT in the list.
template <int KC, bool SAMPLE>
__global__ void __launch_bounds__(kThreads, 1) candidates_kernel(const uint8_t* __restrict__ image, int64_t n, int n_blocks,
                                                                 float* __restrict__ cand_sim, int32_t* __restrict__ cand_idx,
                                                                 const float* __restrict__ thr_in, float* __restrict__ thr_out,
                                                                 float* __restrict__ bound_out) {
  constexpr int kJStep = SAMPLE ? kSampleStride : 1;
  constexpr int kList = SAMPLE ? kSampleKeep : kCand;
  constexpr int NA = Cfg<KC>::kNA;
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
    for (int i = 0; i < 2; ++i) { mbar_init(bar_tfull + 8 * i, 1); mbar_init(bar_tempty + 8 * i, NA * 128); }
    fence_barrier_init();
  }
  if (warp == kEpiWarps) tmem_alloc(smem_u32(tmem_ptr_smem), 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp == kEpiWarps + 1) {
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
        for (int j = 0; j < n_blocks; j += kJStep) {
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
  } else if (warp == kEpiWarps) {
    // =============================== MMA issuer ===============================================================
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(128, 128);
      uint32_t bs = 0, bph = 0, ts = 0, tph = 0, aph = 0;
      for (int sb = blockIdx.x; sb < n_super; sb += gridDim.x) {
        mbar_wait(bar_afull, aph);
        for (int j = 0; j < n_blocks; j += kJStep) {
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
  } else if ((warp >> 2) < NA) {
    // =============================== epilogue: running top-48 per row ==========================================
    const int a = warp >> 2, q = warp & 3;        // A block of the super block, TMEM lane quarter
    TopList<kList> top;
    const uint32_t stage0 = sStage + warp * kStageBytesPerWarp + lane * 8;     // entry k of this row: stage0 + k * 256
    uint32_t ts = 0, tph = 0;
    for (int sb = blockIdx.x; sb < n_super; sb += gridDim.x) {
      const int row = (NA * sb + a) * kBlk + q * 32 + lane;          // n < 2^31 (checked by the entry point)
      const int ni = (int)n;
      const bool live = (NA * sb + a) < n_blocks && row < ni;
      top.reset();
      const float thr0 = live ? ((!SAMPLE && thr_in) ? thr_in[row] : -INFINITY) : INFINITY;      // padding rows never append
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
      for (int j = 0; j < n_blocks; j += kJStep) {
        mbar_wait(bar_tfull + 8 * ts, tph);
        tc_fence_after();
        // the diagonal and the columns past the last item must never be appended: only the row block's own column block and
        // the last block can contain them (warp-uniform test)
        const bool edge = j == NA * sb + a || (j + 1) * kBlk > ni;
#pragma unroll 1
        for (int c = 0; c < 4; ++c) {
          if (KNN_DIAG & 2) continue;
          float v[32];
          tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (ts * 2 + a) * 128 + c * 32, v);
          const int col0 = j * kBlk + c * 32;
          if (edge) {
#pragma unroll
            for (int t = 0; t < 32; ++t)
              if (col0 + t == row || col0 + t >= ni) v[t] = -INFINITY;
          }
          // maxima of the four 8-column groups, then of the chunk
          float g[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            g[u] = fmaxf(fmaxf(fmaxf(v[8 * u], v[8 * u + 1]), fmaxf(v[8 * u + 2], v[8 * u + 3])),
                         fmaxf(fmaxf(v[8 * u + 4], v[8 * u + 5]), fmaxf(v[8 * u + 6], v[8 * u + 7])));
          }
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
    if (x > thr) {                                                                                         \
      st_shared_u2(stage0 + cnt * 256, __float_as_uint(x), (uint32_t)(cbase + t));                        \
      ++cnt;                                                                                               \
    }                                                                                                      \
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
        if (SAMPLE) {
          thr_out[row] = top.s[kList - 1];
        } else {
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
  if (warp == kEpiWarps) tmem_dealloc(tmem_base, 512);
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
    B200GAT_CUDA(cudaFuncSetAttribute(knn::candidates_kernel<KC, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::kSmem));
    B200GAT_CUDA(cudaFuncSetAttribute(knn::candidates_kernel<KC, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::kSmem));
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
  float* bound = (float*)p;
  // Opt-in (B200GAT_KNN_SAMPLE=1, read per call).  Measured at 498,196 x 128-d items: 113.7 ms without, 100.6 ms with.  Off by
  // default because a catalogue stored cluster by cluster defeats the estimate: rows whose own cluster falls into a sampled
  // block get a threshold above their k-th neighbour and are redone by the exact path (255 MB of reads per row at that size).
  const char* env = getenv("B200GAT_KNN_SAMPLE");
  const bool sample = n_blocks >= knn::kSampleMinBlocks && env && atoi(env) != 0;
  count_launch(), knn::prepare_kernel<KC><<<ceil_div(n_pad * 32, 128), 128, 0, st>>>(emb, n_items, n_pad, en, image);
  const int n_super = (n_blocks + C::kNA - 1) / C::kNA;
  const int grid = n_super < kNumSMs ? n_super : kNumSMs;
  if (sample)
    count_launch(), knn::candidates_kernel<KC, true><<<grid, knn::kThreads, C::kSmem, st>>>(image, n_items, n_blocks, nullptr, nullptr,
                                                                                           nullptr, thr, nullptr);
  count_launch(), knn::candidates_kernel<KC, false><<<grid, knn::kThreads, C::kSmem, st>>>(image, n_items, n_blocks, cand_sim, cand_idx,
                                                                                          sample ? thr : nullptr, nullptr, bound);
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
           (size_t)n_pad * 12 /*unsafe row list, sampled thresholds, bounds*/ + 1024;
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
