// fp32 CUDA-core dense kernels around the projection h = x W^T (scripts/train_gat_custom.py:77):
// a strided tiled SGEMM (NT / NN / TN-with-split reduction), the attention-logit row dots
// ((h*a).sum(-1), train_gat_custom.py:79), and the small parameter-gradient finalisers.
// This is the strict-fp32 path (true FFMA accumulation); the tensor-core path lives in
// gemm_tcgen05.cu and is validated against this one.
#include <stdint.h>
#include <stdlib.h>

#include "common.cuh"
#include "../../include/b200gat.h"

namespace b200gat {

// C[m,n] (+)= sum_k A(m,k) * B(k,n);  A(m,k) = A[m*sam + k*sak], B(k,n) = B[k*sbk + n*sbn].
// blockIdx.z selects a K-slab [z*k_slab, min(K,(z+1)*k_slab)) written to C + z*c_slab_stride.
constexpr int kBM = 64, kBN = 64, kBK = 16;

template <bool A_KMAJOR /*sak==1*/, bool B_NMAJOR /*sbn==1*/>
__global__ void __launch_bounds__(256) sgemm_kernel(const float* __restrict__ A, int64_t sam, int64_t sak,
                                                    const float* __restrict__ B, int64_t sbk, int64_t sbn,
                                                    float* __restrict__ Cm, int64_t ldc, int M, int N, int64_t K,
                                                    int64_t k_slab, int64_t c_slab_stride,
                                                    const float* __restrict__ bias /*[N] or null*/) {
  __shared__ float As[kBK][kBM + 4];
  __shared__ float Bs[kBK][kBN + 4];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;  // 16 x 16 threads, 4x4 micro-tile
  const int m0 = blockIdx.x * kBM, n0 = blockIdx.y * kBN;   // M tiles on x: the row count may need more than 65535 of them
  const int64_t kbeg = blockIdx.z * k_slab;
  const int64_t kend = min(K, kbeg + k_slab);
  float acc[4][4] = {};
  for (int64_t k0 = kbeg; k0 < kend; k0 += kBK) {
    // A tile: 64 x 16
#pragma unroll
    for (int it = 0; it < 4; ++it) {
      const int idx = tid + it * 256;
      int mm, kk;
      if (A_KMAJOR) { kk = idx & 15; mm = idx >> 4; } else { mm = idx & 63; kk = idx >> 6; }
      const int gm = m0 + mm;
      const int64_t gk = k0 + kk;
      As[kk][mm] = (gm < M && gk < kend) ? __ldg(A + gm * sam + gk * sak) : 0.f;
    }
#pragma unroll
    for (int it = 0; it < 4; ++it) {
      const int idx = tid + it * 256;
      int nn, kk;
      if (B_NMAJOR) { nn = idx & 63; kk = idx >> 6; } else { kk = idx & 15; nn = idx >> 4; }
      const int gn = n0 + nn;
      const int64_t gk = k0 + kk;
      Bs[kk][nn] = (gn < N && gk < kend) ? __ldg(B + gk * sbk + gn * sbn) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < kBK; ++kk) {
      const float4 a = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
      const float4 b = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
  float* Cz = Cm + blockIdx.z * c_slab_stride;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int gm = m0 + ty * 4 + i;
    if (gm >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int gn = n0 + tx * 4 + j;
      if (gn < N) Cz[(int64_t)gm * ldc + gn] = acc[i][j] + (bias ? bias[gn] : 0.f);
    }
  }
}

// out[i] = sum_z part[z*stride + i], fixed order (deterministic split-K reduction)
__global__ void reduce_slabs_kernel(const float* __restrict__ part, int n_slabs, int64_t stride, int64_t n,
                                    float* __restrict__ out) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  float a = 0.f;
  for (int z = 0; z < n_slabs; ++z) a += part[z * stride + i];
  out[i] = a;
}

static int launch_sgemm(const float* A, int64_t sam, int64_t sak, const float* B, int64_t sbk, int64_t sbn, float* C,
                        int64_t ldc, int M, int N, int64_t K, int n_slabs, int64_t k_slab, int64_t c_slab_stride,
                        const float* bias, cudaStream_t st) {
  dim3 grid(ceil_div(M, kBM), ceil_div(N, kBN), n_slabs);
  const bool ak = sak == 1, bn = sbn == 1;
  if (ak && bn) count_launch(), sgemm_kernel<true, true><<<grid, 256, 0, st>>>(A, sam, sak, B, sbk, sbn, C, ldc, M, N, K, k_slab, c_slab_stride, bias);
  else if (ak && !bn) count_launch(), sgemm_kernel<true, false><<<grid, 256, 0, st>>>(A, sam, sak, B, sbk, sbn, C, ldc, M, N, K, k_slab, c_slab_stride, bias);
  else if (!ak && bn) count_launch(), sgemm_kernel<false, true><<<grid, 256, 0, st>>>(A, sam, sak, B, sbk, sbn, C, ldc, M, N, K, k_slab, c_slab_stride, bias);
  else count_launch(), sgemm_kernel<false, false><<<grid, 256, 0, st>>>(A, sam, sak, B, sbk, sbn, C, ldc, M, N, K, k_slab, c_slab_stride, bias);
  B200GAT_LAUNCH_CHECK();
  return kOk;
}

// s[n, 0:H] = (h[n,hh,:] . a_src[hh,:]),  s[n, H:2H] = (h[n,hh,:] . a_dst[hh,:])   (warp per row)
__global__ void __launch_bounds__(128) logits_kernel(const float* __restrict__ h, const float* __restrict__ a_src,
                                                     const float* __restrict__ a_dst, int64_t n_rows, int H, int C,
                                                     float* __restrict__ s) {
  const int lane = threadIdx.x & 31;
  const int64_t r = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  if (r >= n_rows) return;
  for (int hh = 0; hh < H; ++hh) {
    float ps = 0.f, pd = 0.f;
    for (int c = lane * 4; c < C; c += 128) {
      const float4 v = ldg4(h + (r * H + hh) * C + c);
      ps += dot4(v, ldg4(a_src + hh * C + c));
      pd += dot4(v, ldg4(a_dst + hh * C + c));
    }
    ps = warp_sum(ps);
    pd = warp_sum(pd);
    if (lane == 0) {
      s[r * (2 * H) + hh] = ps;
      s[r * (2 * H) + H + hh] = pd;
    }
  }
}

// dh[n,hh,c] += ds_src[n,hh]*a_src[hh,c] + ds_dst[n,hh]*a_dst[hh,c]   (gradient of the two row dots)
__global__ void add_logit_grad_kernel(float* __restrict__ dh, const float* __restrict__ ds, const float* __restrict__ a_src,
                                      const float* __restrict__ a_dst, int64_t n_rows, int H, int C) {
  const int64_t HC4 = (int64_t)H * C / 4;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < n_rows * HC4;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = idx / HC4;
    const int off = (int)(idx - r * HC4) * 4;
    const int hh = off / C;
    const float gs = ds[r * (2 * H) + hh], gd = ds[r * (2 * H) + H + hh];
    float4 v = *reinterpret_cast<float4*>(dh + r * H * C + off);
    const float4 as = ldg4(a_src + off), ad = ldg4(a_dst + off);
    v.x += gs * as.x + gd * ad.x;
    v.y += gs * as.y + gd * ad.y;
    v.z += gs * as.z + gd * ad.z;
    v.w += gs * as.w + gd * ad.w;
    *reinterpret_cast<float4*>(dh + r * H * C + off) = v;
  }
}

// da_src[hh,c] = W[hh*C+c,:] . v[hh,:],  da_dst[hh,c] = W[hh*C+c,:] . v[H+hh,:]
//   where v = [ds_src | ds_dst]^T x   (== h^T ds since h = x W^T)
__global__ void att_grad_kernel(const float* __restrict__ W, const float* __restrict__ v, int H, int C, int F,
                                float* __restrict__ da_src, float* __restrict__ da_dst) {
  const int lane = threadIdx.x & 31;
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;  // hh*C + c
  if (row >= H * C) return;
  const int hh = row / C;
  float ps = 0.f, pd = 0.f;
  for (int k = lane; k < F; k += 32) {
    const float w = W[(int64_t)row * F + k];
    ps += w * v[(int64_t)hh * F + k];
    pd += w * v[(int64_t)(H + hh) * F + k];
  }
  ps = warp_sum(ps);
  pd = warp_sum(pd);
  if (lane == 0) { da_src[row] = ps; da_dst[row] = pd; }
}

// column sums of a [n, C] matrix with a fixed reduction tree: slabs of rows -> partial[slab, C] -> out[C]
__global__ void __launch_bounds__(256) colsum_partial_kernel(const float* __restrict__ a, int64_t n, int C, int64_t rows_per_slab,
                                                             float* __restrict__ part) {
  // 8 warps stride over the slab's rows, lanes over 128-bit column chunks; fixed row -> warp map, fixed combine order
  __shared__ float4 red[8][32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t r0 = blockIdx.x * rows_per_slab, r1 = min(n, r0 + rows_per_slab);
  for (int c = lane * 4; c < C; c += 128) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int64_t r = r0 + warp; r < r1; r += 8) {
      const float4 v = ld_stream4(a + r * C + c);
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    red[warp][lane] = acc;
    __syncthreads();
    if (warp == 0) {
      float4 t = red[0][lane];
#pragma unroll
      for (int w = 1; w < 8; ++w) { t.x += red[w][lane].x; t.y += red[w][lane].y; t.z += red[w][lane].z; t.w += red[w][lane].w; }
      *reinterpret_cast<float4*>(part + (int64_t)blockIdx.x * C + c) = t;
    }
    __syncthreads();
  }
}

}  // namespace b200gat

using namespace b200gat;

static const int kSlabs = 2 * kNumSMs;  // split count for reductions over the node dimension

namespace b200gat {  // gemm_tc.cu
bool tc_supported(int in_features, int heads, int channels);
int tc_parts();
size_t tc_workspace_bytes(int heads);
int tc_project_fwd(const float* x, const float* W, const float* a_src, const float* a_dst, int64_t n_rows, int heads,
                   float* h, float* s, void* workspace, cudaStream_t st, float* const* peer_h = nullptr, float* const* peer_s = nullptr,
                   int n_peers = 0);
int tc_project_bwd(const float* x, const float* W, const float* a_src, const float* a_dst, const float* dh, const float* ds,
                   int64_t n_rows, int heads, float* dx, float* dW, float* da_src, float* da_dst, void* workspace, cudaStream_t st,
                   float* const* peer_dx = nullptr, int n_peers = 0);
int tc_linear_fwd(const float* x, const float* W, const float* bias, int64_t n_rows, float* out, int64_t ldo, void* workspace,
                  cudaStream_t st);
int tc_linear_dw(const float* x, const float* dy, int64_t n_rows, float* dW, void* workspace, cudaStream_t st);
size_t tc_dw_tiles_workspace_bytes(int heads, int in_features);
size_t bf16_gemm_workspace_bytes(int in_features, int heads, int channels);
int att_grad_launch(const float* W, const float* v, int H, int C, int F, float* da_src, float* da_dst, cudaStream_t st) {
  count_launch(), att_grad_kernel<<<ceil_div((int64_t)H * C * 32, 128), 128, 0, st>>>(W, v, H, C, F, da_src, da_dst);
  B200GAT_LAUNCH_CHECK();
  return kOk;
}
}  // namespace b200gat

static size_t simt_workspace_bytes(int heads, int channels, int in_features) {
  const size_t hc = (size_t)heads * channels;
  return ((size_t)kSlabs * (hc + 2 * heads) * in_features + (size_t)kSlabs * channels + 2 * heads * in_features) * sizeof(float) + 1024;
}

extern "C" int b200gat_dense_workspace_bytes(int heads, int channels, int in_features, size_t* bytes) {
  B200GAT_CHECK_ARG(bytes, "null");
  const size_t a = simt_workspace_bytes(heads, channels, in_features), b = tc_workspace_bytes(heads);
  const size_t c = (bf16_gemm_workspace_bytes(in_features, heads, channels) + 255) / 256 * 256 + tc_dw_tiles_workspace_bytes(heads, in_features);
  *bytes = a > b ? (a > c ? a : c) : (b > c ? b : c);
  return kOk;
}

// h = x W^T ; s = [h.a_src | h.a_dst]
extern "C" int b200gat_project_f32(const float* x, const float* W, const float* a_src, const float* a_dst, int64_t n_rows,
                                   int in_features, int heads, int channels, float* h, float* s, void* workspace,
                                   size_t workspace_bytes, void* stream) {
  B200GAT_CHECK_ARG(x && W && a_src && a_dst && h && s, "null pointer");
  B200GAT_CHECK_ARG(channels % 4 == 0 && in_features > 0, "bad dims");
  if (n_rows == 0) return kOk;
  cudaStream_t st = (cudaStream_t)stream;
  if (tc_supported(in_features, heads, channels) && (tc_parts() & 1)) {
    B200GAT_CHECK_ARG(workspace && workspace_bytes >= tc_workspace_bytes(heads), "workspace too small for the tensor-core path");
    return tc_project_fwd(x, W, a_src, a_dst, n_rows, heads, h, s, workspace, st);
  }
  const int HC = heads * channels;
  B200GAT_CHECK_ARG(n_rows < 2147483647LL, "n_rows too large");
  int rc = launch_sgemm(x, in_features, 1, W, 1, in_features, h, HC, (int)n_rows, HC, in_features, 1, in_features, 0,
                        nullptr, st);
  if (rc) return rc;
  count_launch(), logits_kernel<<<ceil_div(n_rows * 32, 128), 128, 0, st>>>(h, a_src, a_dst, n_rows, heads, channels, s);
  B200GAT_LAUNCH_CHECK();
  return kOk;
}

// Fused projection + exchange for the row-sharded path (heads == 1, in_features == channels == 128, tensor-core mode): the
// GEMM epilogue stores every h tile and the logits into this rank's block of each peer's exchange buffer as well as into the
// local one, so the NVLink transfer runs under the GEMM instead of after it.  peer_h[q] / peer_s[q]: where h / s (row 0 of
// this call) live in peer q's mapped buffer.  Returns kErrUnsupported for other shapes (the caller then projects and pushes).
extern "C" int b200gat_project_push_f32(const float* x, const float* W, const float* a_src, const float* a_dst, int64_t n_rows,
                                        int in_features, int heads, int channels, float* h, float* s, void* const* peer_h,
                                        void* const* peer_s, int n_peers, void* workspace, size_t workspace_bytes, void* stream) {
  B200GAT_CHECK_ARG(x && W && a_src && a_dst && h && s, "null pointer");
  B200GAT_CHECK_ARG(n_peers >= 0 && n_peers <= 7 && (n_peers == 0 || (peer_h && peer_s)), "0..7 peers");
  if (!(heads == 1 && tc_supported(in_features, heads, channels) && (tc_parts() & 1))) {
    set_error("fused projection + exchange needs heads == 1, in_features == channels == 128 and the tensor-core GEMM mode");
    return kErrUnsupported;
  }
  B200GAT_CHECK_ARG(workspace && workspace_bytes >= tc_workspace_bytes(heads), "workspace too small for the tensor-core path");
  if (n_rows == 0) return kOk;
  float* ph[7];
  float* ps[7];
  for (int q = 0; q < n_peers; ++q) {
    B200GAT_CHECK_ARG(peer_h[q] && peer_s[q], "null peer pointer");
    ph[q] = (float*)peer_h[q];
    ps[q] = (float*)peer_s[q];
  }
  return tc_project_fwd(x, W, a_src, a_dst, n_rows, heads, h, s, workspace, (cudaStream_t)stream, ph, ps, n_peers);
}

// The same for the backward: dx (the next layer's dout) is stored into the peers' buffers from the dx GEMM's epilogue.
extern "C" int b200gat_project_bwd_push_f32(const float* x, const float* W, const float* a_src, const float* a_dst, const float* dh,
                                            const float* ds, int64_t n_rows, int in_features, int heads, int channels, float* dx,
                                            void* const* peer_dx, int n_peers, float* dW, float* da_src, float* da_dst,
                                            void* workspace, size_t workspace_bytes, void* stream) {
  B200GAT_CHECK_ARG(x && W && a_src && a_dst && dh && ds && dx && dW && da_src && da_dst && workspace, "null pointer");
  B200GAT_CHECK_ARG(n_peers >= 0 && n_peers <= 7 && (n_peers == 0 || peer_dx), "0..7 peers");
  if (!(heads == 1 && tc_supported(in_features, heads, channels) && (tc_parts() & 6) == 6)) {
    set_error("fused projection backward + exchange needs heads == 1, in_features == channels == 128 and the tensor-core GEMM mode");
    return kErrUnsupported;
  }
  size_t need;
  b200gat_dense_workspace_bytes(heads, channels, in_features, &need);
  B200GAT_CHECK_ARG(workspace_bytes >= need, "workspace too small: %zu < %zu", workspace_bytes, need);
  if (n_rows == 0) return b200gat_project_bwd_f32(x, W, a_src, a_dst, const_cast<float*>(dh), ds, n_rows, in_features, heads, channels, dx,
                                                  dW, da_src, da_dst, workspace, workspace_bytes, stream);
  float* pd[7];
  for (int q = 0; q < n_peers; ++q) {
    B200GAT_CHECK_ARG(peer_dx[q], "null peer pointer");
    pd[q] = (float*)peer_dx[q];
  }
  return tc_project_bwd(x, W, a_src, a_dst, dh, ds, n_rows, heads, dx, dW, da_src, da_dst, workspace, (cudaStream_t)stream, pd, n_peers);
}

// Given dh (aggregation part, overwritten with the full dh) and ds = [ds_src|ds_dst]:
//   dx = dh W ; dW = dh^T x ; da_src, da_dst
extern "C" int b200gat_project_bwd_f32(const float* x, const float* W, const float* a_src, const float* a_dst,
                                       float* dh, const float* ds, int64_t n_rows, int in_features, int heads,
                                       int channels, float* dx /*nullable*/, float* dW, float* da_src, float* da_dst,
                                       void* workspace, size_t workspace_bytes, void* stream) {
  B200GAT_CHECK_ARG(x && W && a_src && a_dst && dh && ds && dW && da_src && da_dst && workspace, "null pointer");
  size_t need;
  b200gat_dense_workspace_bytes(heads, channels, in_features, &need);
  B200GAT_CHECK_ARG(workspace_bytes >= need, "workspace too small: %zu < %zu", workspace_bytes, need);
  cudaStream_t st = (cudaStream_t)stream;
  const int HC = heads * channels, F = in_features, H2 = 2 * heads;
  const int parts = tc_supported(in_features, heads, channels) ? (tc_parts() & 6) : 0;
  if (n_rows > 0 && parts == 6)
    return tc_project_bwd(x, W, a_src, a_dst, dh, ds, n_rows, heads, dx, dW, da_src, da_dst, workspace, st);
  if (n_rows > 0 && parts) {   // diagnostic split: one of dx / dW on the tensor cores (they read dh before it is rewritten below)
    int rc0 = tc_project_bwd(x, W, a_src, a_dst, dh, ds, n_rows, heads, (parts & 2) ? dx : nullptr, (parts & 4) ? dW : nullptr, da_src,
                             da_dst, workspace, st);
    if (rc0) return rc0;
  }
  float* part = (float*)workspace;                       // [kSlabs, HC + 2H, F]
  float* v = part + (size_t)kSlabs * (HC + H2) * F + (size_t)kSlabs * channels;  // [2H, F]
  if (n_rows == 0) {
    B200GAT_CUDA(cudaMemsetAsync(dW, 0, sizeof(float) * HC * F, st));
    B200GAT_CUDA(cudaMemsetAsync(da_src, 0, sizeof(float) * HC, st));
    B200GAT_CUDA(cudaMemsetAsync(da_dst, 0, sizeof(float) * HC, st));
    return kOk;
  }
  count_launch(), add_logit_grad_kernel<<<kNumSMs * 8, 256, 0, st>>>(dh, ds, a_src, a_dst, n_rows, heads, channels);
  B200GAT_LAUNCH_CHECK();
  int rc;
  if (dx && !(parts & 2)) {  // dx[n,f] = sum_m dh[n,m] W[m,f]
    rc = launch_sgemm(dh, HC, 1, W, F, 1, dx, F, (int)n_rows, F, HC, 1, HC, 0, nullptr, st);
    if (rc) return rc;
  }
  if (parts & 4) return kOk;
  const int64_t k_slab = (n_rows + kSlabs - 1) / kSlabs;
  const int n_slabs = (int)((n_rows + k_slab - 1) / k_slab);
  // dW[m,f] = sum_n dh[n,m] x[n,f]
  rc = launch_sgemm(dh, 1, HC, x, F, 1, part, F, HC, F, n_rows, n_slabs, k_slab, (int64_t)HC * F, nullptr, st);
  if (rc) return rc;
  count_launch(), reduce_slabs_kernel<<<ceil_div((int64_t)HC * F, 256), 256, 0, st>>>(part, n_slabs, (int64_t)HC * F, (int64_t)HC * F, dW);
  // v[q,f] = sum_n ds[n,q] x[n,f]
  float* part_v = part + (size_t)kSlabs * HC * F;
  rc = launch_sgemm(ds, 1, H2, x, F, 1, part_v, F, H2, F, n_rows, n_slabs, k_slab, (int64_t)H2 * F, nullptr, st);
  if (rc) return rc;
  count_launch(), reduce_slabs_kernel<<<ceil_div((int64_t)H2 * F, 256), 256, 0, st>>>(part_v, n_slabs, (int64_t)H2 * F, (int64_t)H2 * F, v);
  count_launch(), att_grad_kernel<<<ceil_div((int64_t)HC * 32, 128), 128, 0, st>>>(W, v, heads, channels, F, da_src, da_dst);
  B200GAT_LAUNCH_CHECK();
  return kOk;
}

// ------------------------------------------------------------------------------------------------
// linear128_ffma : y[n, 0:128] = x[n, 0:128] W[128,128]^T + bias with true fp32 FFMA accumulation (round to nearest).
// The item-feature projection is the one GEMM of the path whose output is mixed with rows that carry no GEMM error at all
// (the user embedding rows of the same node-feature matrix): the tensor core's truncating accumulation shrinks every item
// row by ~2e-6, and config 1's cancellation-heavy gradients amplify exactly that imbalance (measured: every parameter
// gradient within 6.5e-6 of the fp64 oracle with this kernel, up to 1.7e-3 with the tensor-core flavour).
// Persistent CTAs; W^T resident in shared memory; x tiles of 128 rows double-buffered with cp.async; 8 x 8 outputs per thread.
// ------------------------------------------------------------------------------------------------
constexpr int kLinRows = 128, kLinThreads = 256, kLinPitch = 132;   // pitch: 16-byte aligned, rows 4 banks apart
constexpr int kLinSmem = (128 * 128 + 2 * kLinRows * kLinPitch) * 4;

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc, bool valid) {
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
  const int n = valid ? 16 : 0;     // src-size 0: the 16 destination bytes are zero-filled
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(gsrc), "r"(n) : "memory");
}

__global__ void __launch_bounds__(kLinThreads, 1) linear128_ffma_kernel(const float* __restrict__ x, const float* __restrict__ W,
                                                                        const float* __restrict__ bias, int64_t n_rows,
                                                                        float* __restrict__ y, int64_t ldy) {
  extern __shared__ float4 lin_smem4[];
  float* Ws = reinterpret_cast<float*>(lin_smem4);          // [k][n] = W[n][k]
  float* Xs = Ws + 128 * 128;                               // [2][kLinRows][kLinPitch]
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int64_t n_tiles = (n_rows + kLinRows - 1) / kLinRows;
  auto stage = [&](int64_t tile, int buf) {
    float* dst = Xs + (size_t)buf * kLinRows * kLinPitch;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const int idx = tid + i * kLinThreads;                // 0 .. 4095 : 128 rows x 32 chunks of 16 B
      const int r = idx >> 5, c = idx & 31;
      const int64_t row = tile * kLinRows + r;
      cp_async16(dst + r * kLinPitch + c * 4, x + (row < n_rows ? row : 0) * 128 + c * 4, row < n_rows);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  int64_t tile = blockIdx.x;
  if (tile < n_tiles) stage(tile, 0);
  for (int i = tid; i < 128 * 128; i += kLinThreads) {      // transpose W once per CTA (coalesced reads, 32-way spread writes)
    const int n = i >> 7, k = i & 127;
    Ws[k * 128 + n] = __ldg(W + i);
  }
  float4 b0 = make_float4(0.f, 0.f, 0.f, 0.f), b1 = b0;
  if (bias) { b0 = ldg4(bias + tx * 4); b1 = ldg4(bias + 64 + tx * 4); }
  int buf = 0;
  for (; tile < n_tiles; tile += gridDim.x, buf ^= 1) {
    const int64_t next = tile + gridDim.x;
    if (next < n_tiles) {
      stage(next, buf ^ 1);
      asm volatile("cp.async.wait_group 1;" ::: "memory");
    } else {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    __syncthreads();
    const float* X = Xs + (size_t)buf * kLinRows * kLinPitch + (ty * 8) * kLinPitch;
    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
#pragma unroll 2
    for (int k4 = 0; k4 < 128; k4 += 4) {
      float4 a[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) a[i] = *reinterpret_cast<const float4*>(X + i * kLinPitch + k4);
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        const float4 w0 = *reinterpret_cast<const float4*>(Ws + (k4 + kk) * 128 + tx * 4);
        const float4 w1 = *reinterpret_cast<const float4*>(Ws + (k4 + kk) * 128 + 64 + tx * 4);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float av = kk == 0 ? a[i].x : kk == 1 ? a[i].y : kk == 2 ? a[i].z : a[i].w;
          acc[i][0] = fmaf(av, w0.x, acc[i][0]); acc[i][1] = fmaf(av, w0.y, acc[i][1]);
          acc[i][2] = fmaf(av, w0.z, acc[i][2]); acc[i][3] = fmaf(av, w0.w, acc[i][3]);
          acc[i][4] = fmaf(av, w1.x, acc[i][4]); acc[i][5] = fmaf(av, w1.y, acc[i][5]);
          acc[i][6] = fmaf(av, w1.z, acc[i][6]); acc[i][7] = fmaf(av, w1.w, acc[i][7]);
        }
      }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int64_t row = tile * kLinRows + ty * 8 + i;
      if (row < n_rows) {
        float* dst = y + row * ldy;
        *reinterpret_cast<float4*>(dst + tx * 4) = make_float4(acc[i][0] + b0.x, acc[i][1] + b0.y, acc[i][2] + b0.z, acc[i][3] + b0.w);
        *reinterpret_cast<float4*>(dst + 64 + tx * 4) = make_float4(acc[i][4] + b1.x, acc[i][5] + b1.y, acc[i][6] + b1.z, acc[i][7] + b1.w);
      }
    }
    __syncthreads();      // everybody is done with this buffer before the next iteration's cp.async overwrites it
  }
}

static int launch_linear128_ffma(const float* x, const float* W, const float* bias, int64_t n_rows, float* y, int64_t ldy,
                                 cudaStream_t st) {
  static DeviceOnce once;
  if (once.pending()) {
    B200GAT_CUDA(cudaFuncSetAttribute(linear128_ffma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kLinSmem));
    once.done();
  }
  const int64_t n_tiles = (n_rows + kLinRows - 1) / kLinRows;
  const int grid = (int)(n_tiles < kNumSMs ? n_tiles : kNumSMs);
  count_launch(), linear128_ffma_kernel<<<grid, kLinThreads, kLinSmem, st>>>(x, W, bias, n_rows, y, ldy);
  B200GAT_LAUNCH_CHECK();
  return kOk;
}

// y[n, ldy] = x W^T + bias   (CustomGAT.item_proj, scripts/train_gat_custom.py:100,107; writes straight into the tail of the
// [N, C] node-feature buffer, which removes the torch.cat copy of :109)
extern "C" int b200gat_linear_f32(const float* x, const float* W, const float* bias, int64_t n_rows, int in_features,
                                  int out_features, float* y, int64_t ldy, void* workspace, size_t workspace_bytes,
                                  void* stream) {
  B200GAT_CHECK_ARG(x && W && y && ldy >= out_features, "null pointer / bad ld");
  if (n_rows == 0) return kOk;
  cudaStream_t st = (cudaStream_t)stream;
  // default: fp32 FFMA (see linear128_ffma_kernel); the tensor-core flavour only on request (B200GAT_TC_PARTS bit 8 AND
  // B200GAT_LINEAR_TC=1), kept for A/B measurements
  static const bool linear_tc = getenv("B200GAT_LINEAR_TC") && atoi(getenv("B200GAT_LINEAR_TC")) != 0;
  if (linear_tc && tc_supported(in_features, 1, out_features) && ldy % 4 == 0 && (tc_parts() & 8)) {
    B200GAT_CHECK_ARG(workspace && workspace_bytes >= tc_workspace_bytes(1), "workspace too small for the tensor-core path");
    return tc_linear_fwd(x, W, bias, n_rows, y, ldy, workspace, st);
  }
  if (in_features == 128 && out_features == 128 && ldy % 4 == 0 && ((uintptr_t)x % 16 == 0) && ((uintptr_t)y % 16 == 0))
    return launch_linear128_ffma(x, W, bias, n_rows, y, ldy, st);
  return launch_sgemm(x, in_features, 1, W, 1, in_features, y, ldy, (int)n_rows, out_features, in_features, 1, in_features, 0,
                      bias, st);
}

// The same projection on the tensor cores (TF32 hi/lo split, ~1e-7 per element but with the tensor core's truncating
// accumulation): for the bf16 tier, whose next step rounds x to bf16 anyway.  Falls back to b200gat_linear_f32 for other shapes.
extern "C" int b200gat_linear_tc_f32(const float* x, const float* W, const float* bias, int64_t n_rows, int in_features,
                                     int out_features, float* y, int64_t ldy, void* workspace, size_t workspace_bytes,
                                     void* stream) {
  B200GAT_CHECK_ARG(x && W && y && ldy >= out_features, "null pointer / bad ld");
  if (n_rows == 0) return kOk;
  if (tc_supported(in_features, 1, out_features) && ldy % 4 == 0 && (tc_parts() & 8)) {
    B200GAT_CHECK_ARG(workspace && workspace_bytes >= tc_workspace_bytes(1), "workspace too small for the tensor-core path");
    return tc_linear_fwd(x, W, bias, n_rows, y, ldy, workspace, (cudaStream_t)stream);
  }
  return b200gat_linear_f32(x, W, bias, n_rows, in_features, out_features, y, ldy, workspace, workspace_bytes, stream);
}

// dW = dy^T x, dbias = column sums of dy   (backward of the above; x needs no gradient: item features are inputs)
extern "C" int b200gat_linear_bwd_f32(const float* x, const float* dy, int64_t ldy, int64_t n_rows, int in_features,
                                      int out_features, float* dW, float* dbias, void* workspace, size_t workspace_bytes,
                                      void* stream) {
  B200GAT_CHECK_ARG(x && dy && dW && workspace, "null pointer");
  size_t need;
  b200gat_dense_workspace_bytes(1, out_features, in_features, &need);
  B200GAT_CHECK_ARG(workspace_bytes >= need, "workspace too small: %zu < %zu", workspace_bytes, need);
  cudaStream_t st = (cudaStream_t)stream;
  if (n_rows == 0) {
    B200GAT_CUDA(cudaMemsetAsync(dW, 0, sizeof(float) * out_features * in_features, st));
    if (dbias) B200GAT_CUDA(cudaMemsetAsync(dbias, 0, sizeof(float) * out_features, st));
    return kOk;
  }
  int rc;
  if (tc_supported(in_features, 1, out_features) && ldy == out_features && (tc_parts() & 16)) {
    rc = tc_linear_dw(x, dy, n_rows, dW, workspace, st);
  } else {
    const int64_t k_slab = (n_rows + kSlabs - 1) / kSlabs;
    const int n_slabs = (int)((n_rows + k_slab - 1) / k_slab);
    float* part = (float*)workspace;
    rc = launch_sgemm(dy, 1, ldy, x, in_features, 1, part, in_features, out_features, in_features, n_rows, n_slabs, k_slab,
                      (int64_t)out_features * in_features, nullptr, st);
    if (rc) return rc;
    count_launch(), reduce_slabs_kernel<<<ceil_div((int64_t)out_features * in_features, 256), 256, 0, st>>>(
        part, n_slabs, (int64_t)out_features * in_features, (int64_t)out_features * in_features, dW);
  }
  if (rc) return rc;
  if (dbias) {
    B200GAT_CHECK_ARG(ldy == out_features, "dbias needs contiguous dy");
    return b200gat_colsum_f32(dy, n_rows, out_features, dbias, workspace, workspace_bytes, stream);
  }
  B200GAT_LAUNCH_CHECK();
  return kOk;
}

// dbias[c] = sum_n dout[n,c]
extern "C" int b200gat_colsum_f32(const float* a, int64_t n_rows, int channels, float* out, void* workspace,
                                  size_t workspace_bytes, void* stream) {
  B200GAT_CHECK_ARG(a && out && workspace, "null pointer");
  B200GAT_CHECK_ARG(workspace_bytes >= (size_t)kSlabs * channels * sizeof(float), "workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  if (n_rows == 0) { B200GAT_CUDA(cudaMemsetAsync(out, 0, sizeof(float) * channels, st)); return kOk; }
  const int64_t rps = (n_rows + kSlabs - 1) / kSlabs;
  const int n_slabs = (int)((n_rows + rps - 1) / rps);
  count_launch(), colsum_partial_kernel<<<n_slabs, 256, 0, st>>>(a, n_rows, channels, rps, (float*)workspace);
  count_launch(), reduce_slabs_kernel<<<ceil_div(channels, 256), 256, 0, st>>>((const float*)workspace, n_slabs, channels, channels, out);
  B200GAT_LAUNCH_CHECK();
  return kOk;
}
