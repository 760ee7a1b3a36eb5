// Tensor-core (tcgen05 / TMEM) dense kernels for the projection h = x W^T and its backward, sm_100a.
//
// The reference computes self.lin(x) in fp32 (scripts/train_gat_custom.py:77) and the north star asks
// for rtol 1e-5, which plain TF32 (10-bit mantissa) cannot give.  Every fp32 operand is therefore
// split into hi = tf32(x) and lo = x - hi while it is staged into shared memory, and each K step
// issues three UMMAs, hi*hi + lo*hi + hi*lo (the dropped lo*lo term is 2^-22 relative).
// Accumulation is fp32 in TMEM.
//
// All three kernels share one structure (persistent CTAs, 1 per SM):
//   producer warps : coalesced 128-bit global loads -> split -> st.shared into the UMMA canonical
//                    128B-swizzled layout -> fence.proxy.async -> mbarrier arrive
//   MMA warp       : one elected lane issues tcgen05.mma (cta_group::1, kind::tf32, M=128, N=128, K=8),
//                    tcgen05.commit releases the smem stage / publishes the accumulator
//   epilogue warps : tcgen05.ld (32x32b.x32) TMEM -> registers -> smem transpose -> coalesced stores
//
//   proj_fwd : h[n,128] = x[n,128] . W^T  (A K-major streamed, B = W resident, double-buffered TMEM),
//              logits s_src/s_dst = rows of h dotted with a_src/a_dst in the epilogue (train_gat_custom.py:79)
//   proj_dx  : same kernel with A = dh + ds_src*a_src + ds_dst*a_dst formed on the fly, B = W^T
//   proj_dw  : dW[128,128] = dh_full^T . x, reduction over the node dimension: both operands are
//              MN-major in memory, staged untransposed (a_major = b_major = MN); per-CTA partial
//              tiles are reduced in a fixed order (deterministic).  v = [ds_src|ds_dst]^T x is
//              accumulated by the producers on the side.
//
// Shapes: in_features = 128 and channels = 128 per head (heads handled as grid.y for proj_fwd);
// anything else is served by dense_simt.cu.
#include <cuda_bf16.h>

#include <stdlib.h>

#include "common.cuh"
#include "tc_common.cuh"
#include "../../include/b200gat.h"

namespace b200gat {
namespace tc {

// ---- B-operand image: W split into hi/lo and laid out exactly as it sits in shared memory -----------
// image[term][kb][row n][32 floats swizzled], term 0 = hi, 1 = lo; B(n,k) = w[n*ldn + k*ldk]
__global__ void build_b_image_kernel(const float* __restrict__ w, int64_t ldn, int64_t ldk, float* __restrict__ image) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;  // one float4 (4 consecutive k) each
  if (idx >= kTileN * kK / 4) return;
  const int n = idx / (kK / 4), k4 = idx % (kK / 4);
  const int kb = k4 / 8, chunk = k4 % 8;
  float4 v;
  v.x = w[n * ldn + (k4 * 4 + 0) * ldk];
  v.y = w[n * ldn + (k4 * 4 + 1) * ldk];
  v.z = w[n * ldn + (k4 * 4 + 2) * ldk];
  v.w = w[n * ldn + (k4 * 4 + 3) * ldk];
  float4 hi, lo;
  split4(v, hi, lo);
  const uint32_t off = kb * (kTileN * 128) + sw128(n, chunk);
  *reinterpret_cast<float4*>(reinterpret_cast<char*>(image) + off) = hi;
  *reinterpret_cast<float4*>(reinterpret_cast<char*>(image) + (kK / kKB) * kTileN * 128 + off) = lo;
}

// ------------------------------------------------------------------------------------------------
// proj_fwd / proj_dx
// ------------------------------------------------------------------------------------------------
#ifndef TC_PRODUCER_WARPS
#define TC_PRODUCER_WARPS 8      // A/B knob (multiple of 4, divides 32): warps that load, split and stage the A operand.
                                 // Measured at config-2 size: 8 warps fwd 0.181 ms, 16 warps 0.206 ms (bwd unchanged)
#endif
constexpr int kFwdProducerWarps = TC_PRODUCER_WARPS, kFwdEpiWarps = 4;
constexpr int kProdStride = kFwdProducerWarps * 4;            // tile rows covered by one pass of the producer threads
constexpr int kProdRows = kTileM / kProdStride;               // rows per producer thread and k-block
constexpr int kFwdThreads = (kFwdProducerWarps + kFwdEpiWarps + 1) * 32;  // + MMA warp
constexpr int kAStages = 2;
constexpr int kAStageBytes = 2 * kTileM * 128;                 // hi + lo of one 32-wide k block (32 KB)
constexpr int kBImageBytes = 2 * (kK / kKB) * kTileN * 128;    // 128 KB
constexpr int kEpiStageBytes = kFwdEpiWarps * 32 * 128;        // 16 KB
constexpr int kFwdSmem = 1024 + kBImageBytes + kAStages * kAStageBytes + kEpiStageBytes + 2 * kTileN * 4 + 256;

constexpr int kMaxPeers = 7;
struct FwdParams {
  const float* a;         // [n_rows, lda]  (x, or dh for the dx flavour)
  int64_t lda;
  const float* b_images;  // [grid.y][kBImageBytes]
  float* out;             // [n_rows, ldo]
  int64_t ldo;
  int64_t n_rows;
  // flavour LOGITS: s[n, 2*heads] from the rows of h;  flavour DX: a is corrected on the fly
  const float* att_src;   // [heads, 128]
  const float* att_dst;
  float* s;               // [n_rows, 2*heads]
  const float* ds;        // DX flavour: [n_rows, ds_ld] logit gradients; this launch uses columns ds_src_col / ds_dst_col
  int heads;
  int ds_ld, ds_src_col, ds_dst_col;
  int accumulate;         // DX flavour, heads > 1: out += result (one launch per head, the contraction runs over all heads)
  // fused exchange (row-sharded path): every output tile -- and the logits -- is also stored into the same place of each
  // peer's exchange buffer (NVLink stores on peer-mapped pointers), so the transfer runs under the rest of the GEMM
  float* peer_out[kMaxPeers];   // where `out` lives in peer q's buffer
  float* peer_s[kMaxPeers];     // where `s` lives there (LOGITS flavour)
  int n_peers;
};

// FLAVOR 0: logits epilogue (projection forward); 1: A corrected on the fly (dx); 2: bias added in the epilogue (plain Linear)
template <int FLAVOR>
__global__ void __launch_bounds__(kFwdThreads, 1) proj_kernel(FwdParams p) {
  constexpr bool DX = FLAVOR == 1;
  constexpr bool LOGITS = FLAVOR == 0;
  constexpr bool BIAS = FLAVOR == 2;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* sm = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t sB = base;
  const uint32_t sA = sB + kBImageBytes;
  const uint32_t sEpi = sA + kAStages * kAStageBytes;
  float* att = reinterpret_cast<float*>(sm + kBImageBytes + kAStages * kAStageBytes + kEpiStageBytes);  // [2][128]
  const uint32_t sBar = sEpi + kEpiStageBytes + 2 * kTileN * 4;
  // barriers: full[2], empty[2], tmem_full[2], tmem_empty[2], b_ready, tmem_ptr
  const uint32_t bar_full = sBar, bar_empty = sBar + 16, bar_tfull = sBar + 32, bar_tempty = sBar + 48, bar_b = sBar + 64;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(sm + (sBar - base) + 80);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int head = blockIdx.y;
  const int64_t n_tiles = (p.n_rows + kTileM - 1) / kTileM;

  if (threadIdx.x == 0) {
    for (int i = 0; i < kAStages; ++i) {
      mbar_init(bar_full + 8 * i, kFwdProducerWarps * 32);
      mbar_init(bar_empty + 8 * i, 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(bar_tfull + 8 * i, 1);
      mbar_init(bar_tempty + 8 * i, kFwdEpiWarps * 32);
    }
    mbar_init(bar_b, 1);
    fence_barrier_init();
  }
  if (warp == kFwdProducerWarps + kFwdEpiWarps) tmem_alloc(smem_u32(tmem_ptr_smem), 256);
  if (LOGITS) {
    for (int i = threadIdx.x; i < 2 * kTileN; i += kFwdThreads)
      att[i] = i < kTileN ? p.att_src[head * kTileN + i] : p.att_dst[head * kTileN + i - kTileN];
  } else if (DX) {
    for (int i = threadIdx.x; i < 2 * kTileN; i += kFwdThreads)
      att[i] = i < kTileN ? p.att_src[i] : p.att_dst[i - kTileN];
  } else {
    for (int i = threadIdx.x; i < kTileN; i += kFwdThreads) att[i] = p.att_src ? p.att_src[head * kTileN + i] : 0.f;  // bias
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp < kFwdProducerWarps) {
    // =========================== producers: A tiles =============================================
    // Software pipelined: the global loads of k-block it+1 are in flight while k-block it is split and stored.
    const int t = threadIdx.x;          // 0..255
    const int chunk = t & 7, r0 = t >> 3;   // rows r0, r0 + kProdStride, ... of the tile
    constexpr int KB = kK / kKB;
    constexpr int R = kProdRows;
    const int64_t my_tiles = blockIdx.x < n_tiles ? (n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    const int64_t n_it = my_tiles * KB;
    struct Ld { float4 v[R]; float ds[R][2]; };
    auto load = [&](int64_t it, Ld& L) {
      const int64_t row0 = (blockIdx.x + (it / KB) * gridDim.x) * kTileM;
      const int kb = (int)(it % KB);
#pragma unroll
      for (int i = 0; i < R; ++i) {
        const int64_t row = row0 + r0 + kProdStride * i;
        if (row < p.n_rows) {
          L.v[i] = ld_stream4(p.a + row * p.lda + kb * kKB + chunk * 4);
          if (DX) { L.ds[i][0] = __ldg(p.ds + row * p.ds_ld + p.ds_src_col); L.ds[i][1] = __ldg(p.ds + row * p.ds_ld + p.ds_dst_col); }
        } else {
          L.v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (DX) L.ds[i][0] = L.ds[i][1] = 0.f;
        }
      }
    };
    auto store = [&](int64_t it, Ld& L) {
      const int kb = (int)(it % KB);
      const uint32_t stage = (uint32_t)(it % kAStages), phase = (uint32_t)((it / kAStages) & 1);
      if (DX) {  // dh_full = dh + ds_src * a_src + ds_dst * a_dst
        const float4 as = *reinterpret_cast<const float4*>(att + kb * kKB + chunk * 4);
        const float4 ad = *reinterpret_cast<const float4*>(att + kTileN + kb * kKB + chunk * 4);
#pragma unroll
        for (int i = 0; i < R; ++i) {
          L.v[i].x += L.ds[i][0] * as.x + L.ds[i][1] * ad.x;
          L.v[i].y += L.ds[i][0] * as.y + L.ds[i][1] * ad.y;
          L.v[i].z += L.ds[i][0] * as.z + L.ds[i][1] * ad.z;
          L.v[i].w += L.ds[i][0] * as.w + L.ds[i][1] * ad.w;
        }
      }
      mbar_wait(bar_empty + 8 * stage, phase ^ 1);
      uint8_t* dst = sm + kBImageBytes + stage * kAStageBytes;
#pragma unroll
      for (int i = 0; i < R; ++i) {
        float4 hi, lo;
        split4(L.v[i], hi, lo);
        const uint32_t off = sw128(r0 + kProdStride * i, chunk);
        *reinterpret_cast<float4*>(dst + off) = hi;
        *reinterpret_cast<float4*>(dst + kTileM * 128 + off) = lo;
      }
      fence_proxy_async();
      mbar_arrive(bar_full + 8 * stage);
    };
    // four k-blocks of loads in flight per thread (64 KB per SM)
    Ld l0, l1, l2, l3;
    if (0 < n_it) load(0, l0);
    if (1 < n_it) load(1, l1);
    if (2 < n_it) load(2, l2);
    for (int64_t it = 0; it < n_it; it += 4) {
      if (it + 3 < n_it) load(it + 3, l3);
      store(it, l0);
      if (it + 1 < n_it) { if (it + 4 < n_it) load(it + 4, l0); store(it + 1, l1); }
      if (it + 2 < n_it) { if (it + 5 < n_it) load(it + 5, l1); store(it + 2, l2); }
      if (it + 3 < n_it) { if (it + 6 < n_it) load(it + 6, l2); store(it + 3, l3); }
    }
  } else if (warp == kFwdProducerWarps + kFwdEpiWarps) {
    // =========================== MMA issuer ======================================================
    if (lane == 0) {
      // resident B image through the TMA unit (1D bulk copies)
      mbar_expect_tx(bar_b, kBImageBytes);
      const char* src = reinterpret_cast<const char*>(p.b_images) + (size_t)head * kBImageBytes;
      for (int c = 0; c < kBImageBytes; c += 32768) bulk_g2s(sB + c, src + c, 32768, bar_b);
      mbar_wait(bar_b, 0);
      constexpr uint32_t idesc = make_idesc(kTileM, kTileN, 0, 0);
      uint32_t stage = 0, phase = 0, acc = 0, acc_phase = 0;
      for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        mbar_wait(bar_tempty + 8 * acc, acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d = tmem_base + acc * kTileN;
        for (int kb = 0; kb < kK / kKB; ++kb) {
          mbar_wait(bar_full + 8 * stage, phase);
          tc_fence_after();
          const uint32_t a_hi = sA + stage * kAStageBytes, a_lo = a_hi + kTileM * 128;
          const uint32_t b_hi = sB + kb * kTileN * 128, b_lo = b_hi + (kK / kKB) * kTileN * 128;
#pragma unroll
          for (int k = 0; k < kKB / kUmmaK; ++k) {
            const uint32_t ko = k * kUmmaK * 4;
            const uint64_t dah = make_desc(a_hi + ko, 16, 1024), dal = make_desc(a_lo + ko, 16, 1024);
            const uint64_t dbh = make_desc(b_hi + ko, 16, 1024), dbl = make_desc(b_lo + ko, 16, 1024);
            umma_tf32(d, dal, dbl, idesc, (kb | k) != 0);
            umma_tf32(d, dal, dbh, idesc, 1);
            umma_tf32(d, dah, dbl, idesc, 1);
            umma_tf32(d, dah, dbh, idesc, 1);
          }
          umma_commit(bar_empty + 8 * stage);
          if (++stage == kAStages) { stage = 0; phase ^= 1; }
        }
        umma_commit(bar_tfull + 8 * acc);
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
    __syncwarp();
  } else {
    // =========================== epilogue =========================================================
    const int q = warp & 3;               // TMEM lane quarter this warp may read
    uint8_t* stg = sm + kBImageBytes + kAStages * kAStageBytes + (warp - kFwdProducerWarps) * 32 * 128;
    uint32_t acc = 0, acc_phase = 0;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      const int64_t row0 = tile * kTileM + q * 32;
      mbar_wait(bar_tfull + 8 * acc, acc_phase);
      tc_fence_after();
      float ps = 0.f, pd = 0.f;
      for (int c = 0; c < kTileN / 32; ++c) {
        float v[32];
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + acc * kTileN + c * 32, v);
        if (LOGITS) {
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            ps = fmaf(v[j], att[c * 32 + j], ps);
            pd = fmaf(v[j], att[kTileN + c * 32 + j], pd);
          }
        }
        if (BIAS) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] += att[c * 32 + j];
        }
        // transpose through smem: lane = row writes its 32 columns, then 8 lanes store one 128 B row segment
#pragma unroll
        for (int j = 0; j < 8; ++j)
          *reinterpret_cast<float4*>(stg + lane * 128 + (((j ^ lane) & 7) << 4)) = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
        __syncwarp();
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int r = (lane >> 3) + 4 * i, ch = lane & 7;
          float4 o = *reinterpret_cast<const float4*>(stg + r * 128 + (((ch ^ r) & 7) << 4));
          const int64_t row = row0 + r;
          if (row < p.n_rows) {
            float* dst = p.out + row * p.ldo + head * kTileN + c * 32 + ch * 4;
            if (DX && p.accumulate) {
              const float4 prev = ld_stream4(dst);
              o.x += prev.x; o.y += prev.y; o.z += prev.z; o.w += prev.w;
            }
            st_stream4(dst, o);
            const int64_t off = dst - p.out;
            for (int q = 0; q < p.n_peers; ++q) {     // rotate the first peer with the tile so that the links fill evenly
              int qq = q + (int)(tile % (p.n_peers > 0 ? p.n_peers : 1));
              if (qq >= p.n_peers) qq -= p.n_peers;
              st_stream4(p.peer_out[qq] + off, o);
            }
          }
        }
        __syncwarp();
      }
      tc_fence_before();
      mbar_arrive(bar_tempty + 8 * acc);
      if (LOGITS) {
        const int64_t row = row0 + lane;
        if (row < p.n_rows) {
          const int64_t i0 = row * (2 * p.heads) + head, i1 = i0 + p.heads;
          p.s[i0] = ps;
          p.s[i1] = pd;
          for (int q = 0; q < p.n_peers; ++q) { p.peer_s[q][i0] = ps; p.peer_s[q][i1] = pd; }
        }
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == kFwdProducerWarps + kFwdEpiWarps) tmem_dealloc(tmem_base, 256);
}

// ------------------------------------------------------------------------------------------------
// proj_dw : dW[c, f] = sum_n dh_full[n, c] * x[n, f]   (+ v[q, f] = sum_n ds[n, q] x[n, f])
// ------------------------------------------------------------------------------------------------
constexpr int kDwProducerWarps = 8, kDwEpiWarps = 4;
constexpr int kDwThreads = (kDwProducerWarps + kDwEpiWarps + 1) * 32;
constexpr int kDwRows = 32;                                 // node rows per stage (4 UMMA K steps)
constexpr int kDwOperandBytes = 2 * kDwRows * 512;          // hi + lo of one operand (32 KB)
constexpr int kDwStageBytes = 2 * kDwOperandBytes;          // A + B (64 KB)
constexpr int kDwStages = 3;
// The tensor core truncates when it adds into the fp32 accumulator, so a long accumulation chain drifts
// (measured 1.4e-5 relative over 2000 rows).  Every kDwGroup stages (256 node rows, 96 UMMAs) the running tile is
// therefore "promoted": the epilogue warps add it into a second TMEM tile with round-to-nearest FADDs while the
// MMA warp continues into the other running tile.  TMEM columns: [0,128) running A, [128,256) running B, [256,384) sum.
#ifndef TC_DW_GROUP
#define TC_DW_GROUP 2          // stages (x 32 node rows x 16 UMMAs) accumulated by the tensor core before a promotion
#endif
constexpr int kDwGroup = TC_DW_GROUP;
// Only the hi*hi product is large enough for that truncation to matter: the three cross terms (2^-11 and 2^-22 of it) go
// into a tile of their own, [384,512), which runs over all of the CTA's rows and is added once at the end.  The running
// tiles then see one truncating accumulation per K step instead of four.
#ifndef TC_DW_XTILE
#define TC_DW_XTILE 1
#endif
constexpr int kDwSmem = 1024 + kDwStages * kDwStageBytes + kDwEpiWarps * 32 * 128 + 2 * kTileN * 4 + 256;

struct DwParams {
  const float* dh;   // [n_rows, ld_dh] aggregation part of dh (this head's 128 columns)
  int64_t ld_dh;
  const float* ds;   // [n_rows, ds_ld], columns ds_src_col / ds_dst_col
  int ds_ld, ds_src_col, ds_dst_col;
  const void* x;     // [n_rows, ld_x], this launch's 128 columns; fp32, or bf16 when x_bf16 (the sharded bf16 tier keeps the
  int64_t ld_x;      // layer inputs it exchanges as bf16)
  int x_bf16;
  const float* att_src;
  const float* att_dst;
  int64_t n_rows;
  int64_t rows_per_cta;   // multiple of kDwRows
  float* part_dw;    // [grid][128*128]
  float* part_v;     // [grid][2*128]
};

// MN-major tf32 staging (SWIZZLE_128B_BASE32B): atoms of 4 K rows x 128 B (32 floats along MN); inside an atom
// the 32-byte chunk index is XORed with the row index (Swizzle<2,5,2>).  Atoms are laid out K-group major:
// offset = ((r/4)*4 + mn_block) * 512, so LBO (next MN block) = 512 B and SBO (next 4 K rows) = 2048 B.
// Returns the byte offset of 16-byte chunk c4 (0..31) of node row r (0..31).
__device__ __forceinline__ uint32_t mn_off(int r, int c4) {
  const int j = c4 & 7;
  return (uint32_t)(((r >> 2) * 4 + (c4 >> 3)) * 512 + (r & 3) * 128 + ((((j >> 1) ^ r) & 3) << 5) + ((j & 1) << 4));
}

__global__ void __launch_bounds__(kDwThreads, 1) proj_dw_kernel(DwParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* sm = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t sStage = base;
  const uint32_t sEpi = sStage + kDwStages * kDwStageBytes;
  float* att = reinterpret_cast<float*>(sm + kDwStages * kDwStageBytes + kDwEpiWarps * 32 * 128);
  const uint32_t sBar = sEpi + kDwEpiWarps * 32 * 128 + 2 * kTileN * 4;
  const uint32_t bar_full = sBar, bar_empty = sBar + 32, bar_tfull = sBar + 64, bar_tempty = sBar + 80, bar_done = sBar + 96;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(sm + (sBar - base) + 112);
  float* vred = reinterpret_cast<float*>(sm);   // reused after the main loop: [8 warps][2][128]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t r_begin = (int64_t)blockIdx.x * p.rows_per_cta;
  const int64_t r_end = min(p.n_rows, r_begin + p.rows_per_cta);
  const int n_stages_total = r_begin < r_end ? (int)((r_end - r_begin + kDwRows - 1) / kDwRows) : 0;
  const int n_groups = (n_stages_total + kDwGroup - 1) / kDwGroup;

  if (threadIdx.x == 0) {
    for (int i = 0; i < kDwStages; ++i) {
      mbar_init(bar_full + 8 * i, kDwProducerWarps * 32);
      mbar_init(bar_empty + 8 * i, 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(bar_tfull + 8 * i, 1);
      mbar_init(bar_tempty + 8 * i, kDwEpiWarps * 32);
    }
    mbar_init(bar_done, 1);
    fence_barrier_init();
  }
  if (warp == kDwProducerWarps + kDwEpiWarps) tmem_alloc(smem_u32(tmem_ptr_smem), 512);
  for (int i = threadIdx.x; i < 2 * kTileN; i += kDwThreads)
    att[i] = p.ds ? (i < kTileN ? p.att_src[i] : p.att_dst[i - kTileN]) : 0.f;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp < kDwProducerWarps) {
    const int c4 = lane;                       // this lane's 16-byte column chunk, fixed for the whole kernel
    const float4 as = *reinterpret_cast<const float4*>(att + c4 * 4);
    const float4 ad = *reinterpret_cast<const float4*>(att + kTileN + c4 * 4);
    float4 vs = make_float4(0.f, 0.f, 0.f, 0.f), vd = make_float4(0.f, 0.f, 0.f, 0.f);
    struct Ld { float4 g[4], xv[4]; float d0[4], d1[4]; };
    auto load = [&](int it, Ld& L) {
      const int64_t row0 = r_begin + (int64_t)it * kDwRows;
#pragma unroll
      for (int i = 0; i < 4; ++i) {             // warp w handles rows w, w+8, w+16, w+24 of the stage
        const int64_t row = row0 + warp + 8 * i;
        if (row < r_end) {
          L.g[i] = ld_stream4(p.dh + row * p.ld_dh + c4 * 4);
          if (p.x_bf16) {
            const uint2 raw = __ldg(reinterpret_cast<const uint2*>(reinterpret_cast<const __nv_bfloat16*>(p.x) + row * p.ld_x + c4 * 4));
            L.xv[i] = make_float4(__uint_as_float(raw.x << 16), __uint_as_float(raw.x & 0xffff0000u), __uint_as_float(raw.y << 16),
                                  __uint_as_float(raw.y & 0xffff0000u));
          } else {
            L.xv[i] = ld_stream4(reinterpret_cast<const float*>(p.x) + row * p.ld_x + c4 * 4);
          }
          L.d0[i] = p.ds ? __ldg(p.ds + row * p.ds_ld + p.ds_src_col) : 0.f;
          L.d1[i] = p.ds ? __ldg(p.ds + row * p.ds_ld + p.ds_dst_col) : 0.f;
        } else {
          L.g[i] = L.xv[i] = make_float4(0.f, 0.f, 0.f, 0.f);
          L.d0[i] = L.d1[i] = 0.f;
        }
      }
    };
    auto store = [&](int it, Ld& L) {
      const uint32_t stage = (uint32_t)(it % kDwStages), phase = (uint32_t)((it / kDwStages) & 1);
      mbar_wait(bar_empty + 8 * stage, phase ^ 1);
      uint8_t* dA = sm + stage * kDwStageBytes;
      uint8_t* dB = dA + kDwOperandBytes;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        float4 a = L.g[i];
        a.x += L.d0[i] * as.x + L.d1[i] * ad.x;
        a.y += L.d0[i] * as.y + L.d1[i] * ad.y;
        a.z += L.d0[i] * as.z + L.d1[i] * ad.z;
        a.w += L.d0[i] * as.w + L.d1[i] * ad.w;
        vs = fma4(L.d0[i], L.xv[i], vs);
        vd = fma4(L.d1[i], L.xv[i], vd);
        float4 hi, lo;
        const uint32_t off = mn_off(warp + 8 * i, c4);
        split4(a, hi, lo);
        *reinterpret_cast<float4*>(dA + off) = hi;
        *reinterpret_cast<float4*>(dA + kDwRows * 512 + off) = lo;
        split4(L.xv[i], hi, lo);
        *reinterpret_cast<float4*>(dB + off) = hi;
        *reinterpret_cast<float4*>(dB + kDwRows * 512 + off) = lo;
      }
      fence_proxy_async();
      mbar_arrive(bar_full + 8 * stage);
    };
    Ld l0, l1;   // two stages (64 KB per SM) of loads in flight
    if (0 < n_stages_total) load(0, l0);
    for (int it = 0; it < n_stages_total; it += 2) {
      if (it + 1 < n_stages_total) load(it + 1, l1);
      store(it, l0);
      if (it + 1 < n_stages_total) { if (it + 2 < n_stages_total) load(it + 2, l0); store(it + 1, l1); }
    }
    // park the side sums; reduced after the block barrier below
    // (the MMA pipeline may still be reading the stages: wait until the accumulator is published)
    mbar_wait(bar_done, 0);
    *reinterpret_cast<float4*>(vred + (warp * 2 + 0) * kTileN + c4 * 4) = vs;
    *reinterpret_cast<float4*>(vred + (warp * 2 + 1) * kTileN + c4 * 4) = vd;
  } else if (warp == kDwProducerWarps + kDwEpiWarps) {
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc(kTileM, kTileN, 1, 1);
      uint32_t stage = 0, phase = 0;
      for (int it = 0; it < n_stages_total; ++it) {
        const int grp = it / kDwGroup, buf = grp & 1, in_grp = it - grp * kDwGroup;
        if (in_grp == 0) {  // the promotion of this buffer's previous group must have drained it
          mbar_wait(bar_tempty + 8 * buf, ((grp >> 1) & 1) ^ 1);
          tc_fence_after();
        }
        const uint32_t d = tmem_base + buf * kTileN;
        mbar_wait(bar_full + 8 * stage, phase);
        tc_fence_after();
        const uint32_t a_hi = sStage + stage * kDwStageBytes, a_lo = a_hi + kDwRows * 512;
        const uint32_t b_hi = a_hi + kDwOperandBytes, b_lo = b_hi + kDwRows * 512;
#pragma unroll
        for (int kg = 0; kg < kDwRows / kUmmaK; ++kg) {
          const uint32_t ko = kg * 4096;
          const uint64_t dah = make_desc(a_hi + ko, 512, 2048, 1), dal = make_desc(a_lo + ko, 512, 2048, 1);
          const uint64_t dbh = make_desc(b_hi + ko, 512, 2048, 1), dbl = make_desc(b_lo + ko, 512, 2048, 1);
#if TC_DW_XTILE
          const uint32_t dx_ = tmem_base + 3 * kTileN;
          umma_tf32(dx_, dal, dbl, idesc, (it | kg) != 0);
          umma_tf32(dx_, dal, dbh, idesc, 1);
          umma_tf32(dx_, dah, dbl, idesc, 1);
          umma_tf32(d, dah, dbh, idesc, (in_grp | kg) != 0);
#else
          umma_tf32(d, dal, dbl, idesc, (in_grp | kg) != 0);
          umma_tf32(d, dal, dbh, idesc, 1);
          umma_tf32(d, dah, dbl, idesc, 1);
          umma_tf32(d, dah, dbh, idesc, 1);
#endif
        }
        umma_commit(bar_empty + 8 * stage);
        if (in_grp == kDwGroup - 1 || it == n_stages_total - 1) umma_commit(bar_tfull + 8 * buf);
        if (++stage == kDwStages) { stage = 0; phase ^= 1; }
      }
      umma_commit(bar_done);   // every UMMA has finished reading shared memory
    }
    __syncwarp();
  } else {
    const int q = warp & 3;
    uint8_t* stg = sm + kDwStages * kDwStageBytes + (warp - kDwProducerWarps) * 32 * 128;
    float* out = p.part_dw + (size_t)blockIdx.x * kTileM * kTileN;
    const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16);
    // promote all groups but the last into the sum tile
    for (int grp = 0; grp + 1 < n_groups; ++grp) {
      const int buf = grp & 1;
      mbar_wait(bar_tfull + 8 * buf, (grp >> 1) & 1);
      tc_fence_after();
      for (int c = 0; c < kTileN / 32; ++c) {
        float v[32], acc[32];
        tmem_ld32(lane_base + buf * kTileN + c * 32, v);
        if (grp > 0) {
          tmem_ld32(lane_base + 2 * kTileN + c * 32, acc);
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] += acc[j];
        }
        tmem_st32(lane_base + 2 * kTileN + c * 32, v);
      }
      tc_fence_before();
      mbar_arrive(bar_tempty + 8 * buf);
    }
    if (n_groups > 0) {
      const int grp = n_groups - 1;
      mbar_wait(bar_tfull + 8 * (grp & 1), (grp >> 1) & 1);
      tc_fence_after();
    }
    for (int c = 0; c < kTileN / 32; ++c) {
      float v[32];
      if (n_groups > 0) {
        tmem_ld32(lane_base + ((n_groups - 1) & 1) * kTileN + c * 32, v);
        if (n_groups > 1) {
          float acc[32];
          tmem_ld32(lane_base + 2 * kTileN + c * 32, acc);
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] += acc[j];
        }
#if TC_DW_XTILE
        {
          float xt[32];   // the cross terms of all rows (complete: the last group's commit covers every earlier UMMA)
          tmem_ld32(lane_base + 3 * kTileN + c * 32, xt);
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] += xt[j];
        }
#endif
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = 0.f;
      }
#pragma unroll
      for (int j = 0; j < 8; ++j)
        *reinterpret_cast<float4*>(stg + lane * 128 + (((j ^ lane) & 7) << 4)) = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
      __syncwarp();
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int r = (lane >> 3) + 4 * i, ch = lane & 7;
        const float4 o = *reinterpret_cast<const float4*>(stg + r * 128 + (((ch ^ r) & 7) << 4));
        *reinterpret_cast<float4*>(out + (q * 32 + r) * kTileN + c * 32 + ch * 4) = o;
      }
      __syncwarp();
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  // side sums v[q, f]: fixed-order reduction over the 8 producer warps
  if (p.part_v && threadIdx.x < 2 * kTileN) {
    float a = 0.f;
#pragma unroll
    for (int w = 0; w < kDwProducerWarps; ++w) a += vred[(w * 2 + threadIdx.x / kTileN) * kTileN + threadIdx.x % kTileN];
    p.part_v[(size_t)blockIdx.x * 2 * kTileN + threadIdx.x] = a;
  }
  if (warp == kDwProducerWarps + kDwEpiWarps) tmem_dealloc(tmem_base, 512);
}

// out[(i / width) * ld_out + i % width] = sum_z part[z*stride + i]   (a [n / width, width] tile of a wider matrix)
__global__ void reduce_parts_kernel(const float* __restrict__ part, int n_parts, int64_t stride, int64_t n, float* __restrict__ out,
                                    int width, int64_t ld_out) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  float a = 0.f;
  for (int z = 0; z < n_parts; ++z) a += part[z * stride + i];
  out[(i / width) * ld_out + i % width] = a;
}

// da_src[c] = W[c,:] . v[0,:],  da_dst[c] = W[c,:] . v[1,:]     (heads == 1, F == 128)
__global__ void att_grad_tc_kernel(const float* __restrict__ W, const float* __restrict__ v, float* __restrict__ da_src,
                                   float* __restrict__ da_dst) {
  const int lane = threadIdx.x & 31;
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (row >= kTileN) return;
  const float4 w = ldg4(W + row * 128 + lane * 4);
  float ps = warp_sum(dot4(w, ldg4(v + lane * 4)));
  float pd = warp_sum(dot4(w, ldg4(v + 128 + lane * 4)));
  if (lane == 0) { da_src[row] = ps; da_dst[row] = pd; }
}

}  // namespace tc

static int g_gemm_mode = B200GAT_GEMM_TF32X3;

static int ensure_attrs() {
  static DeviceOnce once;
  if (!once.pending()) return kOk;
  B200GAT_CUDA(cudaFuncSetAttribute(tc::proj_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::kFwdSmem));
  B200GAT_CUDA(cudaFuncSetAttribute(tc::proj_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::kFwdSmem));
  B200GAT_CUDA(cudaFuncSetAttribute(tc::proj_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::kFwdSmem));
  B200GAT_CUDA(cudaFuncSetAttribute(tc::proj_dw_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::kDwSmem));
  once.done();
  return kOk;
}

bool tc_supported(int in_features, int heads, int channels) {
  (void)heads;
  return g_gemm_mode == B200GAT_GEMM_TF32X3 && in_features == 128 && channels == 128;
}

// Diagnostic: B200GAT_TC_PARTS (bit mask, default all) limits the tensor-core path to some of the GEMMs so that a parity
// difference can be attributed: 1 projection forward, 2 dx, 4 dW + att gradients, 8 linear forward, 16 linear dW.
int tc_parts() {
  static int parts = -1;
  if (parts < 0) {
    const char* e = getenv("B200GAT_TC_PARTS");
    parts = e ? atoi(e) & 31 : 31;
  }
  return parts;
}

size_t tc_workspace_bytes(int heads) {
  // B images (one per head, or W^T), dW / v partials
  return (size_t)(heads > 1 ? heads : 1) * tc::kBImageBytes + (size_t)kNumSMs * (128 * 128 + 2 * 128) * sizeof(float) +
         2 * 128 * sizeof(float) + 1024;
}

// h[n, heads*128] = x W^T, s = row dots.  workspace >= tc_workspace_bytes(heads)
int tc_project_fwd(const float* x, const float* W, const float* a_src, const float* a_dst, int64_t n_rows, int heads,
                   float* h, float* s, void* workspace, cudaStream_t st, float* const* peer_h, float* const* peer_s, int n_peers) {
  int rc = ensure_attrs();
  if (rc) return rc;
  float* images = (float*)workspace;
  for (int hh = 0; hh < heads; ++hh) {
    count_launch(), tc::build_b_image_kernel<<<ceil_div(128 * 128 / 4, 256), 256, 0, st>>>(
        W + (size_t)hh * 128 * 128, 128, 1, images + (size_t)hh * tc::kBImageBytes / 4);
  }
  tc::FwdParams p{};
  p.a = x; p.lda = 128; p.b_images = images; p.out = h; p.ldo = (int64_t)heads * 128; p.n_rows = n_rows;
  p.att_src = a_src; p.att_dst = a_dst; p.s = s; p.ds = nullptr; p.heads = heads;
  p.n_peers = n_peers;
  for (int q = 0; q < n_peers; ++q) { p.peer_out[q] = peer_h[q]; p.peer_s[q] = peer_s[q]; }
  const int64_t n_tiles = (n_rows + 127) / 128;
  dim3 grid((unsigned)(n_tiles < kNumSMs ? n_tiles : kNumSMs), heads);
  count_launch(), tc::proj_kernel<0><<<grid, tc::kFwdThreads, tc::kFwdSmem, st>>>(p);
  B200GAT_LAUNCH_CHECK();
  return kOk;
}

// dx = dh_full W;  dW = dh_full^T x;  da_src/da_dst  (per head: 128-column slices of dh, rows of W).  dh is NOT modified
// (the logit-gradient correction is applied on the fly).  heads > 1: dx accumulates over one launch per head.
int tc_project_bwd(const float* x, const float* W, const float* a_src, const float* a_dst, const float* dh, const float* ds,
                   int64_t n_rows, int heads, float* dx, float* dW, float* da_src, float* da_dst, void* workspace, cudaStream_t st,
                   float* const* peer_dx, int n_peers) {
  int rc = ensure_attrs();
  if (rc) return rc;
  float* image = (float*)workspace;
  float* part_dw = image + tc::kBImageBytes / 4;
  float* part_v = part_dw + (size_t)kNumSMs * 128 * 128;
  float* v = part_v + (size_t)kNumSMs * 2 * 128;
  const int64_t n_tiles = (n_rows + 127) / 128;
  int64_t per = (n_rows + kNumSMs - 1) / kNumSMs;
  per = (per + tc::kDwRows - 1) / tc::kDwRows * tc::kDwRows;
  const int grid = (int)((n_rows + per - 1) / per);
  for (int hh = 0; hh < heads; ++hh) {
    const float* Wh = W + (size_t)hh * 128 * 128;
    if (dx) {
      // B(n = f, k = c) = W_h[c, f]  ->  ldn = 1, ldk = 128
      count_launch(), tc::build_b_image_kernel<<<ceil_div(128 * 128 / 4, 256), 256, 0, st>>>(Wh, 1, 128, image);
      tc::FwdParams p{};
      p.a = dh + hh * 128; p.lda = (int64_t)heads * 128; p.b_images = image; p.out = dx; p.ldo = 128; p.n_rows = n_rows;
      p.att_src = a_src + hh * 128; p.att_dst = a_dst + hh * 128; p.s = nullptr; p.ds = ds; p.heads = 1;
      p.ds_ld = 2 * heads; p.ds_src_col = hh; p.ds_dst_col = heads + hh; p.accumulate = hh > 0;
      p.n_peers = heads == 1 ? n_peers : 0;      // the fused exchange needs the finished rows: single launch only
      for (int q = 0; q < p.n_peers; ++q) p.peer_out[q] = peer_dx[q];
      count_launch(), tc::proj_kernel<1><<<dim3((unsigned)(n_tiles < kNumSMs ? n_tiles : kNumSMs), 1), tc::kFwdThreads, tc::kFwdSmem, st>>>(p);
    }
    if (!dW) continue;       // dx only (diagnostic split, see tc_parts)
    tc::DwParams q{};
    q.dh = dh + hh * 128; q.ld_dh = (int64_t)heads * 128; q.ds = ds; q.ds_ld = 2 * heads; q.ds_src_col = hh; q.ds_dst_col = heads + hh;
    q.x = x; q.ld_x = 128; q.att_src = a_src + hh * 128; q.att_dst = a_dst + hh * 128; q.n_rows = n_rows;
    q.rows_per_cta = per;
    q.part_dw = part_dw; q.part_v = part_v;
    count_launch(), tc::proj_dw_kernel<<<grid, tc::kDwThreads, tc::kDwSmem, st>>>(q);
    count_launch(), tc::reduce_parts_kernel<<<ceil_div(128 * 128, 256), 256, 0, st>>>(part_dw, grid, 128 * 128, 128 * 128,
                                                                                      dW + (size_t)hh * 128 * 128, 128, 128);
    count_launch(), tc::reduce_parts_kernel<<<1, 256, 0, st>>>(part_v, grid, 256, 256, v, 256, 256);
    count_launch(), tc::att_grad_tc_kernel<<<ceil_div(128 * 32, 128), 128, 0, st>>>(Wh, v, da_src + hh * 128, da_dst + hh * 128);
  }
  B200GAT_LAUNCH_CHECK();
  return kOk;
}

// out[n, ldo] (first 128 columns) = x[n,128] W[128,128]^T + bias
int tc_linear_fwd(const float* x, const float* W, const float* bias, int64_t n_rows, float* out, int64_t ldo, void* workspace,
                  cudaStream_t st) {
  int rc = ensure_attrs();
  if (rc) return rc;
  float* image = (float*)workspace;
  count_launch(), tc::build_b_image_kernel<<<ceil_div(128 * 128 / 4, 256), 256, 0, st>>>(W, 128, 1, image);
  tc::FwdParams p{};
  p.a = x; p.lda = 128; p.b_images = image; p.out = out; p.ldo = ldo; p.n_rows = n_rows;
  p.att_src = bias; p.att_dst = nullptr; p.s = nullptr; p.ds = nullptr; p.heads = 1;
  const int64_t n_tiles = (n_rows + 127) / 128;
  count_launch(), tc::proj_kernel<2><<<dim3((unsigned)(n_tiles < kNumSMs ? n_tiles : kNumSMs), 1), tc::kFwdThreads, tc::kFwdSmem, st>>>(p);
  B200GAT_LAUNCH_CHECK();
  return kOk;
}

// dW[128,128] = dy^T x
int tc_linear_dw(const float* x, const float* dy, int64_t n_rows, float* dW, void* workspace, cudaStream_t st) {
  int rc = ensure_attrs();
  if (rc) return rc;
  float* part_dw = (float*)workspace + tc::kBImageBytes / 4;
  tc::DwParams q{};
  q.dh = dy; q.ld_dh = 128; q.ds = nullptr; q.ds_ld = 0; q.ds_src_col = 0; q.ds_dst_col = 0;
  q.x = x; q.ld_x = 128; q.att_src = nullptr; q.att_dst = nullptr; q.n_rows = n_rows;
  int64_t per = (n_rows + kNumSMs - 1) / kNumSMs;
  per = (per + tc::kDwRows - 1) / tc::kDwRows * tc::kDwRows;
  q.rows_per_cta = per;
  const int grid = (int)((n_rows + per - 1) / per);
  q.part_dw = part_dw; q.part_v = nullptr;
  count_launch(), tc::proj_dw_kernel<<<grid, tc::kDwThreads, tc::kDwSmem, st>>>(q);
  count_launch(), tc::reduce_parts_kernel<<<ceil_div(128 * 128, 256), 256, 0, st>>>(part_dw, grid, 128 * 128, 128 * 128, dW, 128, 128);
  B200GAT_LAUNCH_CHECK();
  return kOk;
}

// dW[H*C, F] = dh_full^T x and v[2H, F] = [ds_src | ds_dst]^T x for any H*C, F that are multiples of 128: one proj_dw launch per
// (128 rows of dW) x (128 columns of x) tile pair, fp32-accurate (TF32 split).  workspace >= tc_dw_tiles_workspace_bytes().
size_t tc_dw_tiles_workspace_bytes(int heads, int in_features) {
  return (size_t)kNumSMs * (128 * 128 + 2 * 128) * sizeof(float) + (size_t)2 * heads * in_features * sizeof(float) + 1024;
}
int tc_project_dw_tiles(const void* x, int x_is_bf16, const float* a_src, const float* a_dst, const float* dh, const float* ds,
                        int64_t n_rows, int F, int H, int C, float* dW, float* v /*[2H, F]*/, void* workspace, cudaStream_t st) {
  int rc = ensure_attrs();
  if (rc) return rc;
  float* part_dw = (float*)workspace;
  float* part_v = part_dw + (size_t)kNumSMs * 128 * 128;
  int64_t per = (n_rows + kNumSMs - 1) / kNumSMs;
  per = (per + tc::kDwRows - 1) / tc::kDwRows * tc::kDwRows;
  const int grid = (int)((n_rows + per - 1) / per);
  const int HC = H * C;
  for (int mt = 0; mt < HC / 128; ++mt) {
    const int hh = (mt * 128) / C;
    const bool first_of_head = (mt * 128) % C == 0;        // the side sums v do not depend on the dW row tile: take them once
    for (int ft = 0; ft < F / 128; ++ft) {
      tc::DwParams q{};
      q.dh = dh + mt * 128; q.ld_dh = HC; q.ds = ds; q.ds_ld = 2 * H; q.ds_src_col = hh; q.ds_dst_col = H + hh;
      q.x = x_is_bf16 ? (const void*)((const __nv_bfloat16*)x + ft * 128) : (const void*)((const float*)x + ft * 128);
      q.x_bf16 = x_is_bf16; q.ld_x = F; q.att_src = a_src + mt * 128; q.att_dst = a_dst + mt * 128; q.n_rows = n_rows;
      q.rows_per_cta = per;
      q.part_dw = part_dw; q.part_v = first_of_head ? part_v : nullptr;
      count_launch(), tc::proj_dw_kernel<<<grid, tc::kDwThreads, tc::kDwSmem, st>>>(q);
      count_launch(), tc::reduce_parts_kernel<<<ceil_div(128 * 128, 256), 256, 0, st>>>(part_dw, grid, 128 * 128, 128 * 128,
                                                                                        dW + (size_t)mt * 128 * F + ft * 128, 128, F);
      if (first_of_head) {   // part_v [grid][2][128]: row 0 -> v[hh, ft*128..], row 1 -> v[H + hh, ft*128..]
        count_launch(), tc::reduce_parts_kernel<<<1, 128, 0, st>>>(part_v, grid, 256, 128, v + (size_t)hh * F + ft * 128, 128, F);
        count_launch(), tc::reduce_parts_kernel<<<1, 128, 0, st>>>(part_v + 128, grid, 256, 128, v + (size_t)(H + hh) * F + ft * 128, 128, F);
      }
    }
  }
  B200GAT_LAUNCH_CHECK();
  return kOk;
}

}  // namespace b200gat

namespace b200gat {  // gemm_bf16.cu, dense_simt.cu
bool bf16_gemm_supported(int in_features, int heads, int channels);
size_t bf16_gemm_workspace_bytes(int in_features, int heads, int channels);
int bf16_project_fwd(const void* x, int x_is_bf16, const float* W, const float* a_src, const float* a_dst, int64_t n_rows, int F, int H,
                     int C, void* h_bf16, float* s, void* workspace, cudaStream_t st);
int bf16_project_dx(const float* dh, const float* ds, const float* W, const float* a_src, const float* a_dst, int64_t n_rows, int F,
                    int H, int C, float* dx, int accumulate, void* workspace, cudaStream_t st);
int att_grad_launch(const float* W, const float* v, int H, int C, int F, float* da_src, float* da_dst, cudaStream_t st);
}  // namespace b200gat

// h_bf16 [n, heads*C] = bf16(x) bf16(W)^T (fp32 accumulate), s = fp32 row dots of the accumulator with a_src / a_dst.
// in_features and channels in {128, 256}, heads * channels <= 1024.  x_is_bf16: x is already stored as bf16.
extern "C" int b200gat_project_bf16_ex(const void* x, int x_is_bf16, const float* W, const float* a_src, const float* a_dst,
                                       int64_t n_rows, int in_features, int heads, int channels, void* h_bf16, float* s,
                                       void* workspace, size_t workspace_bytes, void* stream) {
  using namespace b200gat;
  B200GAT_CHECK_ARG(x && W && a_src && a_dst && h_bf16 && s && workspace, "null pointer");
  if (!bf16_gemm_supported(in_features, heads, channels)) {
    set_error("bf16 projection needs in_features and channels in {128, 256} and heads * channels <= 1024 (got %d, %d x %d)",
              in_features, heads, channels);
    return kErrUnsupported;
  }
  B200GAT_CHECK_ARG(workspace_bytes >= bf16_gemm_workspace_bytes(in_features, heads, channels), "workspace too small");
  if (n_rows == 0) return kOk;
  return bf16_project_fwd(x, x_is_bf16, W, a_src, a_dst, n_rows, in_features, heads, channels, h_bf16, s, workspace, (cudaStream_t)stream);
}
extern "C" int b200gat_project_bf16(const float* x, const float* W, const float* a_src, const float* a_dst, int64_t n_rows,
                                    int in_features, int heads, int channels, void* h_bf16, float* s, void* workspace,
                                    size_t workspace_bytes, void* stream) {
  return b200gat_project_bf16_ex(x, 0, W, a_src, a_dst, n_rows, in_features, heads, channels, h_bf16, s, workspace, workspace_bytes, stream);
}

// Backward of the bf16 projection: dx = dh_full W (one bf16 tensor-core launch over all heads, K = heads * channels),
// dW = dh_full^T x (fp32-accurate TF32-split tiles), da_src / da_dst; dh_full = dh + ds_src (x) a_src + ds_dst (x) a_dst is formed
// on the fly (dh is not modified).  Same shape limits as b200gat_project_bf16.  _ex: x may be stored as bf16, and dx may
// accumulate (per-head streaming: one call per head into the same dx).
extern "C" int b200gat_project_bwd_bf16_ex(const void* x, int x_is_bf16, const float* W, const float* a_src, const float* a_dst,
                                           const float* dh, const float* ds, int64_t n_rows, int in_features, int heads,
                                           int channels, float* dx /*nullable*/, int accumulate_dx, float* dW, float* da_src,
                                           float* da_dst, void* workspace, size_t workspace_bytes, void* stream) {
  using namespace b200gat;
  B200GAT_CHECK_ARG(x && W && a_src && a_dst && dh && ds && dW && da_src && da_dst && workspace, "null pointer");
  if (!bf16_gemm_supported(in_features, heads, channels)) {
    set_error("bf16 projection backward needs in_features and channels in {128, 256} and heads * channels <= 1024 (got %d, %d x %d)",
              in_features, heads, channels);
    return kErrUnsupported;
  }
  const size_t img = (bf16_gemm_workspace_bytes(in_features, heads, channels) + 255) / 256 * 256;
  B200GAT_CHECK_ARG(workspace_bytes >= img + tc_dw_tiles_workspace_bytes(heads, in_features), "workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  const int HC = heads * channels;
  if (n_rows == 0) {
    B200GAT_CUDA(cudaMemsetAsync(dW, 0, sizeof(float) * HC * in_features, st));
    B200GAT_CUDA(cudaMemsetAsync(da_src, 0, sizeof(float) * HC, st));
    B200GAT_CUDA(cudaMemsetAsync(da_dst, 0, sizeof(float) * HC, st));
    return kOk;
  }
  int rc;
  if (dx) {
    rc = bf16_project_dx(dh, ds, W, a_src, a_dst, n_rows, in_features, heads, channels, dx, accumulate_dx, workspace, st);
    if (rc) return rc;
  }
  char* ws2 = (char*)workspace + img;
  float* v = (float*)(ws2 + (size_t)kNumSMs * (128 * 128 + 2 * 128) * sizeof(float));
  rc = tc_project_dw_tiles(x, x_is_bf16, a_src, a_dst, dh, ds, n_rows, in_features, heads, channels, dW, v, ws2, st);
  if (rc) return rc;
  return att_grad_launch(W, v, heads, channels, in_features, da_src, da_dst, st);
}
extern "C" int b200gat_project_bwd_bf16(const float* x, const float* W, const float* a_src, const float* a_dst, const float* dh,
                                        const float* ds, int64_t n_rows, int in_features, int heads, int channels,
                                        float* dx /*nullable*/, float* dW, float* da_src, float* da_dst, void* workspace,
                                        size_t workspace_bytes, void* stream) {
  return b200gat_project_bwd_bf16_ex(x, 0, W, a_src, a_dst, dh, ds, n_rows, in_features, heads, channels, dx, 0, dW, da_src, da_dst,
                                     workspace, workspace_bytes, stream);
}

extern "C" int b200gat_set_gemm_mode(int mode) {
  B200GAT_CHECK_ARG(mode == B200GAT_GEMM_FP32 || mode == B200GAT_GEMM_TF32X3, "unsupported gemm mode %d", mode);
  b200gat::g_gemm_mode = mode;
  return b200gat::kOk;
}
extern "C" int b200gat_get_gemm_mode(void) { return b200gat::g_gemm_mode; }
