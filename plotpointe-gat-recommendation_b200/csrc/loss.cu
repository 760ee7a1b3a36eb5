// Fused ranking loss over sampled (u, i, j) triples: positive/negative dot products, BPR or BCE,
// mean -- replaces scripts/train_gat_custom.py:350-359 (three S x C gathers, two S x C products and
// the reductions) and its autograd backward (an index_put with duplicate rows).
//
// forward : 8 lanes per triple; per-triple gradient coefficients are kept (2S floats); block partial
//           sums are reduced in a fixed order -> deterministic loss.
// backward: the 3S (node, triple) incidences are stably sorted by node once per triple set
//           (radix sort from graph.cu); 8 lanes per node then accumulate its incidences in that
//           order -> dZ[N, C] without atomics, every row written exactly once.
// Both kernels are chains of dependent loads per unit of work, so they use narrow lane groups to keep many in flight.
#include <cuda_bf16.h>

#include "common.cuh"
#include "../../include/b200gat.h"

namespace b200gat {

constexpr int kBpr = B200GAT_LOSS_BPR;
constexpr int kBce = B200GAT_LOSS_BCE;

__device__ __forceinline__ float softplusf(float a) { return fmaxf(a, 0.f) + log1pf(expf(-fabsf(a))); }
__device__ __forceinline__ float sigmoidf_(float a) { return 1.f / (1.f + expf(-a)); }

// 8 lanes per triple (32 triples per block of 256 threads): a triple is a chain triple ids -> node_map -> three rows, so
// the kernel's speed is the number of triples in flight; a lane owns 4-channel pieces sl, sl + 8, ... of each row.
constexpr int kTripleW = 8;
constexpr int kTriplesPerBlock = 256 / kTripleW;

// Where the rows of Z live.  Single GPU: one [N, C] matrix.  Row-sharded: rank b owns rows [b * n_max, (b+1) * n_max) of the
// gathered row space and keeps them in its own (peer-mapped) buffer; a triple's rows are then read straight out of the
// owners' memory over NVLink -- S x 3 rows instead of an all-gather of all N rows.
constexpr int kMaxBlocks = 8;
struct RowSource {
  const float* base[kMaxBlocks];
  int64_t n_max;
  int n_blocks;
  __device__ __forceinline__ const float* row(int64_t g, int C) const {
    if (n_blocks <= 1) return base[0] + g * C;
    const int64_t b = g / n_max;
    return base[b] + (g - b * n_max) * C;
  }
};

template <int LOSS>
__global__ void __launch_bounds__(256) loss_fwd_kernel(const RowSource z, int C, int64_t n_users, int64_t n_items,
                                                       const int64_t* __restrict__ u, const int64_t* __restrict__ i,
                                                       const int64_t* __restrict__ j, int64_t S, int64_t t_begin, int64_t t_end,
                                                       const int32_t* __restrict__ node_map,
                                                       float* __restrict__ coef /*[2S]: d/dpos, d/dneg*/,
                                                       double* __restrict__ partial, int32_t* __restrict__ n_bad) {
  constexpr int W = kTripleW;
  __shared__ float wsum[kTriplesPerBlock];
  const int sl = threadIdx.x & (W - 1), g = threadIdx.x / W;
  const int64_t t = t_begin + blockIdx.x * (int64_t)kTriplesPerBlock + g;   // this launch covers triples [t_begin, t_end)
  float term = 0.f;
  float pos = 0.f, neg = 0.f;
  if (t < t_end) {
    int64_t uu = u[t], ii = i[t], jj = j[t];
    const bool ok = uu >= 0 && uu < n_users && ii >= 0 && ii < n_items && jj >= 0 && jj < n_items;
    if (!ok) { uu = 0; ii = 0; jj = 0; if (sl == 0) atomicAdd(n_bad, 1); }
    int64_t nu_ = uu, ni_ = n_users + ii, nj_ = n_users + jj;
    if (node_map) { nu_ = node_map[nu_]; ni_ = node_map[ni_]; nj_ = node_map[nj_]; }
    const float* zu = z.row(nu_, C);
    const float* zi = z.row(ni_, C);
    const float* zj = z.row(nj_, C);
    for (int c = sl * 4; c < C; c += W * 4) {
      const float4 a = ldg4(zu + c);
      pos += dot4(a, ldg4(zi + c));
      neg += dot4(a, ldg4(zj + c));
    }
  }
  // all 32 lanes take part (the groups of a warp are independent; xor offsets < 8 stay inside a group)
#pragma unroll
  for (int o = W / 2; o > 0; o >>= 1) {
    pos += __shfl_xor_sync(kFull, pos, o);
    neg += __shfl_xor_sync(kFull, neg, o);
  }
  if (t < t_end) {
    float gp, gn;
    if (LOSS == kBpr) {  // -log(sigmoid(pos-neg) + 1e-8), train_gat_custom.py:355
      const float sg = sigmoidf_(pos - neg);
      term = -logf(sg + 1e-8f);
      gp = -(sg * (1.f - sg)) / (sg + 1e-8f);
      gn = -gp;
    } else {  // BCE-with-logits on [pos -> 1, neg -> 0], train_gat_custom.py:357-359
      term = softplusf(-pos) + softplusf(neg);
      gp = -sigmoidf_(-pos);
      gn = sigmoidf_(neg);
    }
    if (sl == 0) { coef[t] = gp; coef[S + t] = gn; }
  }
  if (sl == 0) wsum[g] = term;
  __syncthreads();
  if (threadIdx.x == 0) {   // fixed order -> deterministic loss
    double a = 0.0;
#pragma unroll
    for (int k = 0; k < kTriplesPerBlock; ++k) a += (double)wsum[k];
    partial[blockIdx.x] = a;
  }
}

__global__ void __launch_bounds__(256) loss_finalize_kernel(const double* __restrict__ partial, int n, double scale,
                                                            const int32_t* __restrict__ n_bad, float* __restrict__ loss) {
  __shared__ double sm[256];
  double a = 0.0;
  for (int k = threadIdx.x; k < n; k += 256) a += partial[k];
  sm[threadIdx.x] = a;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) sm[threadIdx.x] += sm[threadIdx.x + o];
    __syncthreads();
  }
  // an out-of-range triple index poisons the loss instead of silently reading another row
  if (threadIdx.x == 0) loss[0] = *n_bad ? __int_as_float(0x7fc00000) : (float)(sm[0] * scale);
}

__global__ void incidence_keys_kernel(const int64_t* __restrict__ u, const int64_t* __restrict__ i,
                                      const int64_t* __restrict__ j, int64_t S, int64_t n_users, int64_t n_items,
                                      int32_t* __restrict__ keys) {
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < S; t += (int64_t)gridDim.x * blockDim.x) {
    int64_t uu = u[t], ii = i[t], jj = j[t];
    if (uu < 0 || uu >= n_users) uu = 0;
    if (ii < 0 || ii >= n_items) ii = 0;
    if (jj < 0 || jj >= n_items) jj = 0;
    keys[t] = (int32_t)uu;
    keys[S + t] = (int32_t)(n_users + ii);
    keys[2 * S + t] = (int32_t)(n_users + jj);
  }
}

// kLossW lanes per node (4 nodes per warp): most nodes carry 0-2 incidences, so the kernel is a chain of dependent loads per
// node (ptr -> ids -> triple -> node_map -> partner row) and its speed is the number of chains in flight, not bytes.
// A lane owns 4-channel pieces sl, sl + W, sl + 2W, sl + 3W of every block of 16 W channels.  No shuffles: the groups of a
// warp may run different trip counts.
constexpr int kLossW = 8;

__global__ void __launch_bounds__(128) loss_bwd_kernel(const RowSource z, int C, int64_t n_nodes, int64_t n_users,
                                                       const int64_t* __restrict__ u, const int64_t* __restrict__ i,
                                                       const int64_t* __restrict__ j, int64_t S,
                                                       const float* __restrict__ coef, const int32_t* __restrict__ ptr,
                                                       const int32_t* __restrict__ ids, const float* __restrict__ grad_out,
                                                       float scale, const int32_t* __restrict__ node_list, int64_t node_begin,
                                                       int64_t node_count, const int32_t* __restrict__ node_map,
                                                       float* __restrict__ dz /*[node_count, C]*/,
                                                       __nv_bfloat16* __restrict__ dz_bf16 /*optional bf16 copy*/) {
  constexpr int W = kLossW;
  const int sl = threadIdx.x & (W - 1);
  const int64_t local = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) / W;
  if (local >= node_count) return;
  const int64_t n = node_list ? (int64_t)node_list[local] : node_begin + local;
  const bool real = n >= 0;                      // node_list entries < 0 are padding rows: written as zeros
  const int beg = real ? ptr[n] : 0, end = real ? ptr[n + 1] : 0;
  const float g = grad_out[0] * scale;
  const int64_t n_items = n_nodes - n_users;
  auto safe = [](int64_t v, int64_t lim) { return (v < 0 || v >= lim) ? (int64_t)0 : v; };
  auto zrow = [&](int64_t node) { return node_map ? (int64_t)node_map[node] : node; };
  for (int cb = 0; cb < C; cb += 16 * W) {
    float4 acc[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) acc[k] = make_float4(0.f, 0.f, 0.f, 0.f);
    auto add_row = [&](float cf, int64_t node) {
      const float* r = z.row(zrow(node), C) + cb + sl * 4;
#pragma unroll
      for (int k = 0; k < 4; ++k)
        if (cb + (k * W + sl) * 4 < C) acc[k] = fma4(cf, ldg4(r + k * W * 4), acc[k]);
    };
    for (int q = beg; q < end; ++q) {            // node-sorted, stable: a fixed accumulation order
      const int id = ids[q];
      const int role = id / (int)S;
      const int64_t t = id - (int64_t)role * S;
      if (role == 0) {
        add_row(coef[t], n_users + safe(i[t], n_items));
        add_row(coef[S + t], n_users + safe(j[t], n_items));
      } else {
        add_row(coef[(role - 1) * S + t], safe(u[t], n_users));
      }
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int c = cb + (k * W + sl) * 4;
      if (c >= C) continue;
      float4 a = acc[k];
      a.x *= g; a.y *= g; a.z *= g; a.w *= g;
      *reinterpret_cast<float4*>(dz + local * C + c) = a;
      if (dz_bf16) {
        const __nv_bfloat162 lo = __floats2bfloat162_rn(a.x, a.y), hi = __floats2bfloat162_rn(a.z, a.w);
        uint2 pk;
        pk.x = *reinterpret_cast<const uint32_t*>(&lo);
        pk.y = *reinterpret_cast<const uint32_t*>(&hi);
        *reinterpret_cast<uint2*>(dz_bf16 + local * C + c) = pk;
      }
    }
  }
}

static inline size_t al(size_t x) { return (x + 255) / 256 * 256; }

}  // namespace b200gat

using namespace b200gat;

extern "C" int b200gat_loss_workspace_bytes(int64_t n_nodes, int64_t n_triples, size_t* bytes) {
  B200GAT_CHECK_ARG(bytes && n_triples >= 0 && n_nodes >= 0, "bad args");
  B200GAT_CHECK_ARG(3 * n_triples < 2147483647LL && n_nodes < 2147483647LL, "sizes must be < 2^31");
  const int64_t S = n_triples;
  *bytes = al(2 * S * 4 + 4) /*coef*/ + al((size_t)(S / 8 + 2) * 8) /*partials*/ + 3 * al(3 * S * 4 + 4) /*keys, sorted, ids*/ +
           al((n_nodes + 1) * 4) /*ptr*/ + al(sort_workspace_bytes(3 * S)) + 256 /*n_bad*/;
  return kOk;
}

namespace {
struct LossWs {
  float* coef; double* partial; int32_t *keys, *sorted, *ids, *ptr; void* sort_ws; int32_t* n_bad;
};
LossWs carve(void* ws, int64_t n_nodes, int64_t S) {
  char* p = (char*)ws;
  auto take = [&](size_t b) { char* q = p; p += al(b); return (void*)q; };
  LossWs w;
  w.coef = (float*)take(2 * S * 4 + 4);
  w.partial = (double*)take((size_t)(S / 8 + 2) * 8);
  w.keys = (int32_t*)take(3 * S * 4 + 4);
  w.sorted = (int32_t*)take(3 * S * 4 + 4);
  w.ids = (int32_t*)take(3 * S * 4 + 4);
  w.ptr = (int32_t*)take((n_nodes + 1) * 4);
  w.sort_ws = take(sort_workspace_bytes(3 * S));
  w.n_bad = (int32_t*)take(4);
  return w;
}
}  // namespace

namespace {
int make_source(const float* z, const void* const* blocks, int n_blocks, int64_t n_max, RowSource* rs) {
  if (blocks) {
    B200GAT_CHECK_ARG(n_blocks >= 1 && n_blocks <= kMaxBlocks && n_max > 0, "bad row-block table (%d blocks)", n_blocks);
    for (int b = 0; b < kMaxBlocks; ++b) rs->base[b] = b < n_blocks ? (const float*)blocks[b] : nullptr;
    rs->n_blocks = n_blocks;
    rs->n_max = n_max;
  } else {
    B200GAT_CHECK_ARG(z, "null pointer");
    rs->base[0] = z;
    rs->n_blocks = 1;
    rs->n_max = 0;
  }
  return kOk;
}

// forward over triples [t_begin, t_begin + t_count): coef halves written in place, `loss` = (sum of the terms) * scale
int loss_fwd_impl(const RowSource& rs, int64_t n_users, int64_t n_items, int channels, const int64_t* u, const int64_t* i,
                  const int64_t* j, int64_t S, int64_t t_begin, int64_t t_count, const int32_t* node_map, int loss_kind,
                  int need_backward, float* coef, float* loss, LossWs& w, cudaStream_t st) {
  const int64_t N = n_users + n_items;
  B200GAT_CUDA(cudaMemsetAsync(w.n_bad, 0, 4, st));
  const int blocks = ceil_div(t_count > 0 ? t_count : 1, kTriplesPerBlock);
  const double scale = (loss_kind == kBpr ? 1.0 : 0.5) / (double)S;
  if (loss_kind == kBpr)
    count_launch(), loss_fwd_kernel<kBpr><<<blocks, 256, 0, st>>>(rs, channels, n_users, n_items, u, i, j, S, t_begin, t_begin + t_count, node_map, coef, w.partial, w.n_bad);
  else
    count_launch(), loss_fwd_kernel<kBce><<<blocks, 256, 0, st>>>(rs, channels, n_users, n_items, u, i, j, S, t_begin, t_begin + t_count, node_map, coef, w.partial, w.n_bad);
  count_launch(), loss_finalize_kernel<<<1, 256, 0, st>>>(w.partial, blocks, scale, w.n_bad, loss);
  B200GAT_LAUNCH_CHECK();
  if (need_backward) {
    count_launch(), incidence_keys_kernel<<<min(ceil_div(S, 256), kNumSMs * 8), 256, 0, st>>>(u, i, j, S, n_users, n_items, w.keys);
    B200GAT_LAUNCH_CHECK();
    int rc = sort_pairs_stable(w.keys, 3 * S, N, w.sorted, w.ids, w.sort_ws, st);
    if (rc) return rc;
    rc = node_ptr_from_sorted(w.sorted, 3 * S, N, w.ptr, st);
    if (rc) return rc;
  }
  return kOk;
}
}  // namespace

// loss[0] = mean over triples.  Leaves coef + the node-sorted incidence lists in `workspace` for the backward.
extern "C" int b200gat_rank_loss_fwd_f32(const float* z, int64_t n_users, int64_t n_items, int channels, const int64_t* u,
                                         const int64_t* i, const int64_t* j, int64_t n_triples, const int32_t* node_map,
                                         int loss_kind, int need_backward, float* loss, void* workspace,
                                         size_t workspace_bytes, void* stream) {
  B200GAT_CHECK_ARG(z && loss && workspace, "null pointer");
  B200GAT_CHECK_ARG(loss_kind == kBpr || loss_kind == kBce, "bad loss kind %d", loss_kind);
  B200GAT_CHECK_ARG(channels % 4 == 0, "channels must be a multiple of 4");
  B200GAT_CHECK_ARG(n_triples > 0 && u && i && j, "empty triple set");
  size_t need;
  int rc = b200gat_loss_workspace_bytes(n_users + n_items, n_triples, &need);
  if (rc) return rc;
  B200GAT_CHECK_ARG(workspace_bytes >= need, "workspace too small: %zu < %zu", workspace_bytes, need);
  LossWs w = carve(workspace, n_users + n_items, n_triples);
  RowSource rs;
  rc = make_source(z, nullptr, 0, 0, &rs);
  if (rc) return rc;
  return loss_fwd_impl(rs, n_users, n_items, channels, u, i, j, n_triples, 0, n_triples, node_map, loss_kind, need_backward,
                       w.coef, loss, w, (cudaStream_t)stream);
}

// Row-sharded forward: this rank evaluates triples [t_begin, t_begin + t_count) only; their Z rows are read out of the owners'
// row blocks (z_blocks[b] = rank b's [n_max, C] block, peer-mapped).  coef [2 * n_triples]: this rank's slices of the two
// halves are written; loss_partial[0] = this rank's share of the mean.  need_backward also sorts ALL incidences by node.
extern "C" int b200gat_rank_loss_fwd_peer_f32(const void* const* z_blocks, int n_blocks, int64_t n_max, int64_t n_users,
                                              int64_t n_items, int channels, const int64_t* u, const int64_t* i, const int64_t* j,
                                              int64_t n_triples, int64_t t_begin, int64_t t_count, const int32_t* node_map,
                                              int loss_kind, int need_backward, float* coef, float* loss_partial, void* workspace,
                                              size_t workspace_bytes, void* stream) {
  B200GAT_CHECK_ARG(z_blocks && coef && loss_partial && workspace && node_map, "null pointer");
  B200GAT_CHECK_ARG(loss_kind == kBpr || loss_kind == kBce, "bad loss kind %d", loss_kind);
  B200GAT_CHECK_ARG(channels % 4 == 0, "channels must be a multiple of 4");
  B200GAT_CHECK_ARG(n_triples > 0 && u && i && j && t_begin >= 0 && t_count >= 0 && t_begin + t_count <= n_triples, "bad triple range");
  size_t need;
  int rc = b200gat_loss_workspace_bytes(n_users + n_items, n_triples, &need);
  if (rc) return rc;
  B200GAT_CHECK_ARG(workspace_bytes >= need, "workspace too small: %zu < %zu", workspace_bytes, need);
  LossWs w = carve(workspace, n_users + n_items, n_triples);
  RowSource rs;
  rc = make_source(nullptr, z_blocks, n_blocks, n_max, &rs);
  if (rc) return rc;
  return loss_fwd_impl(rs, n_users, n_items, channels, u, i, j, n_triples, t_begin, t_count, node_map, loss_kind, need_backward,
                       coef, loss_partial, w, (cudaStream_t)stream);
}

namespace {
int loss_bwd_impl(const RowSource& rs, int64_t n_users, int64_t n_items, int channels, const int64_t* u, const int64_t* i,
                  const int64_t* j, int64_t S, const int32_t* node_map, int loss_kind, const float* coef, const float* grad_out,
                  const int32_t* node_list, int64_t node_begin, int64_t node_count, float* dz, void* dz_bf16, LossWs& w,
                  cudaStream_t st) {
  const int64_t N = n_users + n_items;
  B200GAT_CHECK_ARG(node_count >= 0 && (node_list || (node_begin >= 0 && node_begin + node_count <= N)), "bad node range");
  if (node_count == 0) return kOk;
  const float scale = loss_kind == kBpr ? 1.f / (float)S : 0.5f / (float)S;
  count_launch(), loss_bwd_kernel<<<ceil_div(node_count * kLossW, 128), 128, 0, st>>>(rs, channels, N, n_users, u, i, j, S, coef, w.ptr, w.ids,
                                                                 grad_out, scale, node_list, node_begin, node_count, node_map, dz,
                                                                 (__nv_bfloat16*)dz_bf16);
  B200GAT_LAUNCH_CHECK();
  return kOk;
}
}  // namespace

extern "C" int b200gat_rank_loss_bwd_f32(const float* z, int64_t n_users, int64_t n_items, int channels, const int64_t* u,
                                         const int64_t* i, const int64_t* j, int64_t n_triples, const int32_t* node_map,
                                         int loss_kind, const float* grad_out, const int32_t* node_list, int64_t node_begin,
                                         int64_t node_count, float* dz, void* dz_bf16, void* workspace,
                                         size_t workspace_bytes, void* stream) {
  B200GAT_CHECK_ARG(z && grad_out && dz && workspace && u && i && j, "null pointer");
  B200GAT_CHECK_ARG(loss_kind == kBpr || loss_kind == kBce, "bad loss kind %d", loss_kind);
  size_t need;
  int rc = b200gat_loss_workspace_bytes(n_users + n_items, n_triples, &need);
  if (rc) return rc;
  B200GAT_CHECK_ARG(workspace_bytes >= need, "workspace too small");
  LossWs w = carve(workspace, n_users + n_items, n_triples);
  RowSource rs;
  rc = make_source(z, nullptr, 0, 0, &rs);
  if (rc) return rc;
  return loss_bwd_impl(rs, n_users, n_items, channels, u, i, j, n_triples, node_map, loss_kind, w.coef, grad_out, node_list,
                       node_begin, node_count, dz, dz_bf16, w, (cudaStream_t)stream);
}

// Row-sharded backward: dZ rows of this rank's nodes (node_list), partner rows read out of the owners' blocks; coef [2 * n_triples]
// must hold every rank's slices (gathered by the caller).  `workspace` is the one the peer forward of the same triples used.
extern "C" int b200gat_rank_loss_bwd_peer_f32(const void* const* z_blocks, int n_blocks, int64_t n_max, int64_t n_users,
                                              int64_t n_items, int channels, const int64_t* u, const int64_t* i, const int64_t* j,
                                              int64_t n_triples, const int32_t* node_map, int loss_kind, const float* coef,
                                              const float* grad_out, const int32_t* node_list, int64_t node_count, float* dz,
                                              void* dz_bf16, void* workspace, size_t workspace_bytes, void* stream) {
  B200GAT_CHECK_ARG(z_blocks && coef && grad_out && dz && workspace && u && i && j && node_map && node_list, "null pointer");
  B200GAT_CHECK_ARG(loss_kind == kBpr || loss_kind == kBce, "bad loss kind %d", loss_kind);
  size_t need;
  int rc = b200gat_loss_workspace_bytes(n_users + n_items, n_triples, &need);
  if (rc) return rc;
  B200GAT_CHECK_ARG(workspace_bytes >= need, "workspace too small");
  LossWs w = carve(workspace, n_users + n_items, n_triples);
  RowSource rs;
  rc = make_source(nullptr, z_blocks, n_blocks, n_max, &rs);
  if (rc) return rc;
  return loss_bwd_impl(rs, n_users, n_items, channels, u, i, j, n_triples, node_map, loss_kind, coef, grad_out, node_list, 0,
                       node_count, dz, dz_bf16, w, (cudaStream_t)stream);
}
