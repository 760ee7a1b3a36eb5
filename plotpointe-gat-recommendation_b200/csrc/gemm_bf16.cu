// bf16 tensor-core (tcgen05 / TMEM) dense kernel for the "bf16 projection" tier (BASELINE configs 3 and 5), any
// contraction length K that is a multiple of 64 and output tiles of 128 or 256 columns:
//
//   FWD : h_bf16[n, H*C] = bf16(x[n, K]) . bf16(W[H*C, K])^T   + logits s_src / s_dst from the fp32 accumulator
//         (scripts/train_gat_pyg.py:77,87 -> GATConv's lin + (h * att).sum(-1));   K = F_in in {128, 256}, C in {128, 256}
//   DX  : dx[n, F] = bf16(dh + ds_src (x) a_src + ds_dst (x) a_dst)[n, H*C] . bf16(W[H*C, F])
//         -- the contraction runs over ALL heads in one launch (K = H*C up to 1024): dh is read once, dx is written once
//         (the per-head launches of the fp32 path re-read x and read-modify-write dx once per head).
//
// Structure (persistent CTAs, 1 per SM): 8 producer warps load the fp32 A rows, round to bf16 and store them in the UMMA
// canonical 128B-swizzled layout; the B operand is a pre-swizzled bf16 image in global memory (built once per call from W,
// L2-resident) and each 64-wide K block of it is fetched into the same pipeline stage by ONE bulk copy (TMA unit,
// cp.async.bulk + mbarrier complete_tx); one elected lane issues tcgen05.mma kind::f16 (M = 128, N = NT, K = 16) into a
// double-buffered TMEM accumulator; 4 epilogue warps read it back with tcgen05.ld.
#include <cuda_bf16.h>

#include "common.cuh"
#include "tc_common.cuh"
#include "../../include/b200gat.h"

namespace b200gat {
namespace tcb {
using namespace tc;

constexpr int kProdWarps = 8, kEpiWarps = 4;
constexpr int kThreads = (kProdWarps + kEpiWarps + 1) * 32;
constexpr int kStages = 4;
constexpr int kKB = 64;                         // bf16 elements per 128-byte swizzle row
constexpr int kAStage = kTileM * 128;           // 16 KB
constexpr int kMaxAtt = 2 * 1024;               // a_src | a_dst of up to H*C = 1024 columns

template <int NT>
constexpr int smem_bytes() { return 1024 + kStages * (kAStage + NT * 128) + kMaxAtt * 4 + 256; }

// image[tile][kb][row n of the tile][64 bf16, swizzled];  B(n, k) = w[n * ldn + k * ldk]
__global__ void build_image_kernel(const float* __restrict__ w, int64_t ldn, int64_t ldk, int n_total, int K, int NT,
                                   uint8_t* __restrict__ image) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;   // one 16-byte chunk (8 consecutive k) each
  const int chunks_per_row = K / 8;
  if (idx >= n_total * chunks_per_row) return;
  const int n = idx / chunks_per_row, k8 = idx % chunks_per_row;
  const int kb = k8 / 8, chunk = k8 % 8;
  float v[8];
#pragma unroll
  for (int q = 0; q < 8; ++q) v[q] = w[n * ldn + (int64_t)(k8 * 8 + q) * ldk];
  const int tile = n / NT, nl = n % NT;
  const size_t off = ((size_t)tile * (K / kKB) + kb) * ((size_t)NT * 128) + sw128(nl, chunk);
  *reinterpret_cast<uint4*>(image + off) = pack8_bf16(make_float4(v[0], v[1], v[2], v[3]), make_float4(v[4], v[5], v[6], v[7]));
}

struct Params {
  const void* a;           // [n_rows, lda]: fp32 (x, or dh for DX), or bf16 x when the kernel is instantiated with A16
  int64_t lda;
  const uint8_t* image;    // B image, see build_image_kernel
  int K;                   // contraction length (multiple of 64)
  int64_t n_rows;
  // FWD
  __nv_bfloat16* out_bf16; // [n_rows, ldo]
  float* s;                // [n_rows, 2 * heads]
  int heads;
  // DX
  float* out_f32;          // [n_rows, ldo]
  const float* ds;         // [n_rows, 2 * heads]
  int channels;            // C: column k of A belongs to head k / C
  int64_t ldo;
  const float* att_src;    // FWD: [heads, NT]   DX: [heads * C]
  const float* att_dst;
  int accumulate;          // DX: out += result (per-head streaming: one launch per head into the same dx)
};

// FLAVOR 0: FWD (bf16 out + logits, blockIdx.y = head / column tile), 1: DX (fp32 out, A corrected on the fly).
// A16: the A operand is already bf16 in memory (the exchanged layer input of the sharded bf16 tier): staged without conversion.
template <int NT, int FLAVOR, bool A16 = false>
__global__ void __launch_bounds__(kThreads, 1) gemm_bf16_kernel(Params p) {
  constexpr bool DX = FLAVOR == 1;
  constexpr int kBStage = NT * 128;
  constexpr int kStageBytes = kAStage + kBStage;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* sm = smem_raw + (base - smem_u32(smem_raw));
  float* att = reinterpret_cast<float*>(sm + kStages * kStageBytes);            // FWD: [2][NT]  DX: [2][H*C]
  const uint32_t sBar = base + kStages * kStageBytes + kMaxAtt * 4;
  const uint32_t bar_full = sBar, bar_empty = sBar + 32, bar_tfull = sBar + 64, bar_tempty = sBar + 80;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(sm + (sBar - base) + 112);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tile_n = blockIdx.y;
  const int64_t n_tiles = (p.n_rows + kTileM - 1) / kTileM;
  const int n_kb = p.K / kKB;
  const int n_att = DX ? p.heads * p.channels : NT;

  if (threadIdx.x == 0) {
    for (int i = 0; i < kStages; ++i) {
      mbar_init(bar_full + 8 * i, kProdWarps * 32 + 1);     // + the expect_tx arrival that announces the B bytes
      mbar_init(bar_empty + 8 * i, 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(bar_tfull + 8 * i, 1);
      mbar_init(bar_tempty + 8 * i, kEpiWarps * 32);
    }
    fence_barrier_init();
  }
  if (warp == kProdWarps + kEpiWarps) tmem_alloc(smem_u32(tmem_ptr_smem), 2 * NT);
  for (int i = threadIdx.x; i < n_att; i += kThreads) {
    att[i] = p.att_src[(DX ? 0 : tile_n * NT) + i];
    att[n_att + i] = p.att_dst[(DX ? 0 : tile_n * NT) + i];
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp < kProdWarps) {
    // =========================== producers: A tile (registers -> bf16 -> swizzled smem) + the B bulk copy ==============
    const int t = threadIdx.x;
    const int chunk = t & 7, r0 = t >> 3;             // rows r0, r0 + 32, r0 + 64, r0 + 96 of the tile
    constexpr int R = kTileM / 32;
    const int64_t my_tiles = blockIdx.x < n_tiles ? (n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    const int64_t n_it = my_tiles * n_kb;
    const uint8_t* img = p.image + (size_t)tile_n * n_kb * kBStage;
    struct Ld { float4 v[R][2]; float d[R][2]; };     // A16: v[i][0] holds the 8 bf16 values as raw bits
    auto load = [&](int64_t it, Ld& L) {
      const int64_t row0 = (blockIdx.x + (it / n_kb) * gridDim.x) * kTileM;
      const int kb = (int)(it % n_kb);
#pragma unroll
      for (int i = 0; i < R; ++i) {
        const int64_t row = row0 + r0 + 32 * i;
        if (row < p.n_rows) {
          if (A16) {
            L.v[i][0] = ld_stream4(reinterpret_cast<const float*>(reinterpret_cast<const __nv_bfloat16*>(p.a) + row * p.lda + kb * kKB + chunk * 8));
          } else {
            const float* src = reinterpret_cast<const float*>(p.a) + row * p.lda + kb * kKB + chunk * 8;
            L.v[i][0] = ld_stream4(src);
            L.v[i][1] = ld_stream4(src + 4);
          }
          if (DX) {
            const int head = (kb * kKB) / p.channels;
            L.d[i][0] = __ldg(p.ds + row * (2 * p.heads) + head);
            L.d[i][1] = __ldg(p.ds + row * (2 * p.heads) + p.heads + head);
          }
        } else {
          L.v[i][0] = L.v[i][1] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (DX) L.d[i][0] = L.d[i][1] = 0.f;
        }
      }
    };
    auto store = [&](int64_t it, Ld& L) {
      const int kb = (int)(it % n_kb);
      const uint32_t stage = (uint32_t)(it % kStages), phase = (uint32_t)((it / kStages) & 1);
      mbar_wait(bar_empty + 8 * stage, phase ^ 1);
      if (t == 0) {   // this K block of the B image: one bulk copy, its bytes complete the same "full" barrier
        mbar_expect_tx(bar_full + 8 * stage, kBStage);
        bulk_g2s(base + stage * kStageBytes + kAStage, img + (size_t)kb * kBStage, kBStage, bar_full + 8 * stage);
      }
      uint8_t* dst = sm + stage * kStageBytes;
      if (DX) {       // dh_full = dh + ds_src * a_src + ds_dst * a_dst   (gradient of the two logit row dots)
        const float* as = att + kb * kKB + chunk * 8;
        const float* ad = att + n_att + kb * kKB + chunk * 8;
        const float4 as0 = *reinterpret_cast<const float4*>(as), as1 = *reinterpret_cast<const float4*>(as + 4);
        const float4 ad0 = *reinterpret_cast<const float4*>(ad), ad1 = *reinterpret_cast<const float4*>(ad + 4);
#pragma unroll
        for (int i = 0; i < R; ++i) {
          const float a = L.d[i][0], b = L.d[i][1];
          L.v[i][0].x += a * as0.x + b * ad0.x; L.v[i][0].y += a * as0.y + b * ad0.y;
          L.v[i][0].z += a * as0.z + b * ad0.z; L.v[i][0].w += a * as0.w + b * ad0.w;
          L.v[i][1].x += a * as1.x + b * ad1.x; L.v[i][1].y += a * as1.y + b * ad1.y;
          L.v[i][1].z += a * as1.z + b * ad1.z; L.v[i][1].w += a * as1.w + b * ad1.w;
        }
      }
#pragma unroll
      for (int i = 0; i < R; ++i) {
        uint4 pk;
        if (A16) pk = make_uint4(__float_as_uint(L.v[i][0].x), __float_as_uint(L.v[i][0].y), __float_as_uint(L.v[i][0].z), __float_as_uint(L.v[i][0].w));
        else pk = pack8_bf16(L.v[i][0], L.v[i][1]);
        *reinterpret_cast<uint4*>(dst + sw128(r0 + 32 * i, chunk)) = pk;
      }
      fence_proxy_async();
      mbar_arrive(bar_full + 8 * stage);
    };
    Ld l0, l1;   // two K blocks of loads in flight per thread
    if (0 < n_it) load(0, l0);
    for (int64_t it = 0; it < n_it; it += 2) {
      if (it + 1 < n_it) load(it + 1, l1);
      store(it, l0);
      if (it + 1 < n_it) { if (it + 2 < n_it) load(it + 2, l0); store(it + 1, l1); }
    }
  } else if (warp == kProdWarps + kEpiWarps) {
    // =========================== MMA issuer =========================================================================
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(kTileM, NT);
      uint32_t stage = 0, phase = 0, acc = 0, acc_phase = 0;
      for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        mbar_wait(bar_tempty + 8 * acc, acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d = tmem_base + acc * NT;
        for (int kb = 0; kb < n_kb; ++kb) {
          mbar_wait(bar_full + 8 * stage, phase);
          tc_fence_after();
          const uint32_t a0 = base + stage * kStageBytes, b0 = a0 + kAStage;
#pragma unroll
          for (int k = 0; k < kKB / 16; ++k)
            umma_bf16(d, make_desc(a0 + k * 32, 16, 1024), make_desc(b0 + k * 32, 16, 1024), idesc, (kb | k) != 0);
          umma_commit(bar_empty + 8 * stage);
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
        umma_commit(bar_tfull + 8 * acc);
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
    __syncwarp();
  } else {
    // =========================== epilogue ===========================================================================
    const int q = warp & 3;               // TMEM lane quarter this warp may read
    uint32_t acc = 0, acc_phase = 0;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      const int64_t row = tile * kTileM + q * 32 + lane;
      mbar_wait(bar_tfull + 8 * acc, acc_phase);
      tc_fence_after();
      float ps = 0.f, pd = 0.f;
      for (int c = 0; c < NT / 32; ++c) {
        float v[32];
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + acc * NT + c * 32, v);
        if (!DX) {
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            ps = fmaf(v[j], att[c * 32 + j], ps);
            pd = fmaf(v[j], att[NT + c * 32 + j], pd);
          }
          if (row < p.n_rows) {   // 64 contiguous bytes of this thread's row
            uint4* o = reinterpret_cast<uint4*>(p.out_bf16 + row * p.ldo + (int64_t)tile_n * NT + c * 32);
#pragma unroll
            for (int j = 0; j < 4; ++j)
              o[j] = pack8_bf16(make_float4(v[8 * j], v[8 * j + 1], v[8 * j + 2], v[8 * j + 3]),
                                make_float4(v[8 * j + 4], v[8 * j + 5], v[8 * j + 6], v[8 * j + 7]));
          }
        } else if (row < p.n_rows) {   // 128 contiguous bytes of this thread's row
          float* o = p.out_f32 + row * p.ldo + (int64_t)tile_n * NT + c * 32;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            float4 r = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
            if (p.accumulate) {
              const float4 prev = ld_stream4(o + 4 * j);
              r.x += prev.x; r.y += prev.y; r.z += prev.z; r.w += prev.w;
            }
            st_stream4(o + 4 * j, r);
          }
        }
      }
      tc_fence_before();
      mbar_arrive(bar_tempty + 8 * acc);
      if (!DX && row < p.n_rows) {
        p.s[row * (2 * p.heads) + tile_n] = ps;
        p.s[row * (2 * p.heads) + p.heads + tile_n] = pd;
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == kProdWarps + kEpiWarps) tmem_dealloc(tmem_base, 2 * NT);
}

static int ensure_attrs() {
  static DeviceOnce once;
  if (!once.pending()) return kOk;
  B200GAT_CUDA(cudaFuncSetAttribute(gemm_bf16_kernel<128, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes<128>()));
  B200GAT_CUDA(cudaFuncSetAttribute(gemm_bf16_kernel<128, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes<128>()));
  B200GAT_CUDA(cudaFuncSetAttribute(gemm_bf16_kernel<256, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes<256>()));
  B200GAT_CUDA(cudaFuncSetAttribute(gemm_bf16_kernel<256, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes<256>()));
  B200GAT_CUDA(cudaFuncSetAttribute((gemm_bf16_kernel<128, 0, true>), cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes<128>()));
  B200GAT_CUDA(cudaFuncSetAttribute((gemm_bf16_kernel<256, 0, true>), cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes<256>()));
  once.done();
  return kOk;
}

}  // namespace tcb

bool bf16_gemm_supported(int in_features, int heads, int channels) {
  return (in_features == 128 || in_features == 256) && (channels == 128 || channels == 256) && heads >= 1 && heads * channels <= 1024;
}
size_t bf16_gemm_workspace_bytes(int in_features, int heads, int channels) {
  return (size_t)heads * channels * in_features * 2 + 1024;       // the B image (forward and dx have the same size)
}

// h_bf16 [n, H*C] = bf16(x) bf16(W)^T, s = fp32 row dots of the accumulator with a_src / a_dst
int bf16_project_fwd(const void* x, int x_is_bf16, const float* W, const float* a_src, const float* a_dst, int64_t n_rows, int F, int H,
                     int C, void* h_bf16, float* s, void* workspace, cudaStream_t st) {
  int rc = tcb::ensure_attrs();
  if (rc) return rc;
  uint8_t* image = (uint8_t*)workspace;
  const int HC = H * C;
  count_launch(), tcb::build_image_kernel<<<ceil_div((int64_t)HC * F / 8, 256), 256, 0, st>>>(W, F, 1, HC, F, C, image);
  tcb::Params p{};
  p.a = x; p.lda = F; p.image = image; p.K = F; p.n_rows = n_rows; p.out_bf16 = (__nv_bfloat16*)h_bf16; p.s = s; p.heads = H;
  p.ldo = HC; p.att_src = a_src; p.att_dst = a_dst; p.channels = C;
  const int64_t n_tiles = (n_rows + 127) / 128;
  const int per = kNumSMs / H > 0 ? kNumSMs / H : 1;
  dim3 grid((unsigned)(n_tiles < per ? n_tiles : per), H);
  if (C == 128 && !x_is_bf16) count_launch(), tcb::gemm_bf16_kernel<128, 0><<<grid, tcb::kThreads, tcb::smem_bytes<128>(), st>>>(p);
  else if (C == 128) count_launch(), tcb::gemm_bf16_kernel<128, 0, true><<<grid, tcb::kThreads, tcb::smem_bytes<128>(), st>>>(p);
  else if (!x_is_bf16) count_launch(), tcb::gemm_bf16_kernel<256, 0><<<grid, tcb::kThreads, tcb::smem_bytes<256>(), st>>>(p);
  else count_launch(), tcb::gemm_bf16_kernel<256, 0, true><<<grid, tcb::kThreads, tcb::smem_bytes<256>(), st>>>(p);
  B200GAT_LAUNCH_CHECK();
  return kOk;
}

// dx[n, F] = (dh + ds_src (x) a_src + ds_dst (x) a_dst)[n, H*C] . W[H*C, F], operands rounded to bf16, fp32 accumulation
int bf16_project_dx(const float* dh, const float* ds, const float* W, const float* a_src, const float* a_dst, int64_t n_rows, int F,
                    int H, int C, float* dx, int accumulate, void* workspace, cudaStream_t st) {
  int rc = tcb::ensure_attrs();
  if (rc) return rc;
  uint8_t* image = (uint8_t*)workspace;
  const int HC = H * C;
  // B(n = f, k = hc) = W[hc, f]  ->  ldn = 1, ldk = F
  count_launch(), tcb::build_image_kernel<<<ceil_div((int64_t)F * HC / 8, 256), 256, 0, st>>>(W, 1, F, F, HC, F, image);
  tcb::Params p{};
  p.a = dh; p.lda = HC; p.image = image; p.K = HC; p.n_rows = n_rows; p.out_f32 = dx; p.ds = ds; p.heads = H; p.channels = C;
  p.ldo = F; p.att_src = a_src; p.att_dst = a_dst; p.accumulate = accumulate;
  const int64_t n_tiles = (n_rows + 127) / 128;
  dim3 grid((unsigned)(n_tiles < kNumSMs ? n_tiles : kNumSMs), 1);
  if (F == 128) count_launch(), tcb::gemm_bf16_kernel<128, 1><<<grid, tcb::kThreads, tcb::smem_bytes<128>(), st>>>(p);
  else count_launch(), tcb::gemm_bf16_kernel<256, 1><<<grid, tcb::kThreads, tcb::smem_bytes<256>(), st>>>(p);
  B200GAT_LAUNCH_CHECK();
  return kOk;
}

}  // namespace b200gat
