// The two callers either side of the hot path:
//   adam_step  : torch.optim.Adam(lr, weight_decay) as the reference constructs it (scripts/train_gat_custom.py:335,362),
//                one fused elementwise kernel per parameter tensor (L2-in-gradient weight decay, bias correction).
//   eval_ranks : the inner loop of eval_sampled (scripts/train_gat_custom.py:200-206): for each evaluated user, the
//                scores of 1 positive + K sampled negatives, i_emb @ u_emb, and rank = (scores > scores[0]).sum() + 1.
#include "common.cuh"
#include "../../include/b200gat.h"

namespace b200gat {

__device__ __forceinline__ void adam_update(float& p, float g, float& m, float& v, float lr_c, float b1, float b2, float eps,
                                            float wd, float inv_sqrt_bc2) {
  const float gi = g + wd * p;                         // weight decay enters the gradient (Adam, not AdamW)
  m = m + (gi - m) * (1.f - b1);                       // exp_avg.lerp_(grad, 1 - beta1)
  v = v * b2 + (1.f - b2) * gi * gi;                   // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, 1 - beta2)
  const float denom = sqrtf(v) * inv_sqrt_bc2 + eps;
  p = p - lr_c * (m / denom);
}

// 128-bit accesses, two independent float4 groups per thread and iteration (7 streams: p, g, m, v in; p, m, v out)
__global__ void __launch_bounds__(256) adam_step_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                                        float* __restrict__ v, int64_t n, float lr, float b1, float b2, float eps,
                                                        float wd, float inv_bc1, float inv_sqrt_bc2, int vec_ok) {
  const float lr_c = lr * inv_bc1;
  const int64_t tid = blockIdx.x * (int64_t)blockDim.x + threadIdx.x, nth = (int64_t)gridDim.x * blockDim.x;
  const int64_t n4 = vec_ok ? n / 4 : 0;
  float4* p4 = reinterpret_cast<float4*>(p);
  const float4* g4 = reinterpret_cast<const float4*>(g);
  float4* m4 = reinterpret_cast<float4*>(m);
  float4* v4 = reinterpret_cast<float4*>(v);
  for (int64_t i = tid; i < n4; i += 2 * nth) {
    const int64_t i2 = i + nth;
    const bool two = i2 < n4;
    float4 pa = p4[i], ga = ld_stream4(reinterpret_cast<const float*>(g4 + i)), ma = m4[i], va = v4[i];
    float4 pb = pa, gb = ga, mb = ma, vb = va;
    if (two) { pb = p4[i2]; gb = ld_stream4(reinterpret_cast<const float*>(g4 + i2)); mb = m4[i2]; vb = v4[i2]; }
    adam_update(pa.x, ga.x, ma.x, va.x, lr_c, b1, b2, eps, wd, inv_sqrt_bc2);
    adam_update(pa.y, ga.y, ma.y, va.y, lr_c, b1, b2, eps, wd, inv_sqrt_bc2);
    adam_update(pa.z, ga.z, ma.z, va.z, lr_c, b1, b2, eps, wd, inv_sqrt_bc2);
    adam_update(pa.w, ga.w, ma.w, va.w, lr_c, b1, b2, eps, wd, inv_sqrt_bc2);
    p4[i] = pa; m4[i] = ma; v4[i] = va;
    if (two) {
      adam_update(pb.x, gb.x, mb.x, vb.x, lr_c, b1, b2, eps, wd, inv_sqrt_bc2);
      adam_update(pb.y, gb.y, mb.y, vb.y, lr_c, b1, b2, eps, wd, inv_sqrt_bc2);
      adam_update(pb.z, gb.z, mb.z, vb.z, lr_c, b1, b2, eps, wd, inv_sqrt_bc2);
      adam_update(pb.w, gb.w, mb.w, vb.w, lr_c, b1, b2, eps, wd, inv_sqrt_bc2);
      p4[i2] = pb; m4[i2] = mb; v4[i2] = vb;
    }
  }
  for (int64_t i = n4 * 4 + tid; i < n; i += nth) {     // tail, or everything when a pointer is not 16-byte aligned
    float pi = p[i], mi = m[i], vi = v[i];
    adam_update(pi, g[i], mi, vi, lr_c, b1, b2, eps, wd, inv_sqrt_bc2);
    p[i] = pi; m[i] = mi; v[i] = vi;
  }
}

constexpr int kEvalThreads = 128;
constexpr int kEvalU = 4;   // candidate rows in flight per warp

__global__ void __launch_bounds__(kEvalThreads) eval_ranks_kernel(const float* __restrict__ z, int64_t n_users, int64_t n_items, int C,
                                                                  const int64_t* __restrict__ users,
                                                                  const int64_t* __restrict__ cand, int64_t n_eval, int K1,
                                                                  int32_t* __restrict__ rank_out, int32_t* __restrict__ n_bad) {
  __shared__ float s_pos;
  __shared__ int s_cnt[kEvalThreads / 32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t q = blockIdx.x;
  if (q >= n_eval) return;
  int64_t u = users[q];
  if (u < 0 || u >= n_users) { u = 0; if (threadIdx.x == 0) atomicAdd(n_bad, 1); }
  const int64_t* cq = cand + q * K1;
  // this lane's slice of the user row (C <= 512: up to 4 chunks of 128 channels)
  float4 ur[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) ur[k] = (k * 128 + lane * 4 < C) ? ldg4(z + u * C + k * 128 + lane * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
  auto score = [&](int64_t item) {
    if (item < 0 || item >= n_items) { item = 0; if (lane == 0) atomicAdd(n_bad, 1); }
    const float* r = z + (n_users + item) * C;
    float d = 0.f;
#pragma unroll
    for (int k = 0; k < 4; ++k)
      if (k * 128 + lane * 4 < C) d += dot4(ur[k], ldg4(r + k * 128 + lane * 4));
    return d;
  };
  if (warp == 0) {
    const float d = warp_sum(score(cq[0]));
    if (lane == 0) s_pos = d;
  }
  __syncthreads();
  const float pos = s_pos;
  int cnt = 0;
  for (int c0 = 1 + warp * kEvalU; c0 < K1; c0 += (kEvalThreads / 32) * kEvalU) {
    float d[kEvalU];
#pragma unroll
    for (int k = 0; k < kEvalU; ++k) d[k] = (c0 + k < K1) ? score(cq[c0 + k]) : 0.f;
#pragma unroll
    for (int k = 0; k < kEvalU; ++k) {
      const float t = warp_sum(d[k]);
      cnt += (c0 + k < K1 && t > pos) ? 1 : 0;
    }
  }
  if (lane == 0) s_cnt[warp] = cnt;
  __syncthreads();
  if (threadIdx.x == 0) {
    int total = 0;
    for (int w = 0; w < kEvalThreads / 32; ++w) total += s_cnt[w];
    rank_out[q] = total + 1;     // (scores > scores[0]).sum() + 1
  }
}

}  // namespace b200gat

using namespace b200gat;

extern "C" int b200gat_adam_step_f32(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n, float lr,
                                     float beta1, float beta2, float eps, float weight_decay, int64_t step, void* stream) {
  B200GAT_CHECK_ARG(n == 0 || (param && grad && exp_avg && exp_avg_sq), "null pointer");
  B200GAT_CHECK_ARG(step >= 1, "step counts from 1");
  if (n == 0) return kOk;
  const double bc1 = 1.0 - pow((double)beta1, (double)step), bc2 = 1.0 - pow((double)beta2, (double)step);
  const int vec_ok = (((uintptr_t)param | (uintptr_t)grad | (uintptr_t)exp_avg | (uintptr_t)exp_avg_sq) & 15) == 0;
  const int64_t want = (n / 8 + 255) / 256 + 1, cap = (int64_t)kNumSMs * 8;
  const int grid = (int)(want < cap ? want : cap);
  count_launch(), adam_step_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(param, grad, exp_avg, exp_avg_sq, n, lr, beta1, beta2, eps,
                                                                          weight_decay, (float)(1.0 / bc1), (float)(1.0 / sqrt(bc2)),
                                                                          vec_ok);
  B200GAT_LAUNCH_CHECK();
  return kOk;
}

extern "C" int b200gat_eval_ranks_f32(const float* z, int64_t n_users, int64_t n_items, int channels, const int64_t* users,
                                      const int64_t* candidates, int64_t n_eval, int n_candidates, int32_t* ranks,
                                      int32_t* n_bad, void* stream) {
  B200GAT_CHECK_ARG(z && ranks && n_bad && (n_eval == 0 || (users && candidates)), "null pointer");
  B200GAT_CHECK_ARG(channels % 4 == 0 && channels > 0 && channels <= 512, "channels must be a multiple of 4, at most 512");
  B200GAT_CHECK_ARG(n_candidates >= 1, "need at least the positive candidate");
  B200GAT_CHECK_ARG(n_eval < 2147483647LL, "too many evaluated users");
  cudaStream_t st = (cudaStream_t)stream;
  B200GAT_CUDA(cudaMemsetAsync(n_bad, 0, sizeof(int32_t), st));
  if (n_eval == 0) return kOk;
  count_launch(), eval_ranks_kernel<<<(unsigned)n_eval, kEvalThreads, 0, st>>>(z, n_users, n_items, channels, users, candidates, n_eval,
                                                                              n_candidates, ranks, n_bad);
  B200GAT_LAUNCH_CHECK();
  return kOk;
}

// ---- f4: BPR triple sampler (scripts/train_gat_custom.py:213-224) ---------------------------------------------------
// The reference draws, per sample, a uniform user among those with training positives, a uniform positive of that user
// and a uniform item that is not one of the user's positives (rejection).  Python's `random` stream cannot be replayed on
// a GPU, so parity is distributional; the draws come from Philox-4x32-10 keyed on (seed, sample, attempt).  The user's
// positives are read from the graph's CSC (out-edges of a user node are exactly its items).
namespace b200gat {
__device__ __forceinline__ void philox4(uint64_t seed, uint32_t c0, uint32_t c1, uint32_t (&out)[4]) {
  uint32_t c2 = 0x243F6A88u, c3 = 0x85A308D3u;
  uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    c0 = hi1 ^ c1 ^ k0; c1 = lo1; c2 = hi0 ^ c3 ^ k1; c3 = lo0;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
// unbiased integer in [0, n): 64-bit multiply-shift on 32 random bits is biased by < n / 2^32 (< 2e-4 here), as is
// Python's own randrange for non powers of two before its rejection step; good to distribution-level parity
__device__ __forceinline__ int64_t bounded(uint32_t r, int64_t n) { return (int64_t)(((uint64_t)r * (uint64_t)n) >> 32); }

__global__ void sample_bpr_kernel(const int32_t* __restrict__ colptr, const int32_t* __restrict__ row, int64_t n_users,
                                  int64_t n_items, int64_t n_samples, uint64_t seed, const int32_t* __restrict__ active_users,
                                  int64_t n_active, int64_t* __restrict__ u_out, int64_t* __restrict__ i_out,
                                  int64_t* __restrict__ j_out, int32_t* __restrict__ n_fail) {
  const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (t >= n_samples) return;
  uint32_t r[4];
  int64_t u = 0;
  int beg = 0, deg = 0;
  uint32_t attempt = 0;
  if (active_users) {                            // the caller's list of users with positives: one draw, no rejection
    philox4(seed, (uint32_t)t, ((uint32_t)(t >> 32) << 8), r);
    u = active_users[bounded(r[0], n_active)];
    beg = colptr[u];
    deg = colptr[u + 1] - beg;
  } else {
    for (; attempt < 64; ++attempt) {            // users without positives are not in the reference's dict: redraw
      philox4(seed, (uint32_t)t, ((uint32_t)(t >> 32) << 8) | attempt, r);
      u = bounded(r[0], n_users);
      beg = colptr[u];
      deg = colptr[u + 1] - beg;
      if (deg > 0) break;
    }
  }
  if (deg <= 0) { atomicAdd(n_fail, 1); u_out[t] = 0; i_out[t] = 0; j_out[t] = 0; return; }
  const int64_t i = (int64_t)row[beg + (int)bounded(r[1], deg)] - n_users;
  int64_t j = -1;
  uint32_t rr[4] = {r[2], r[3], 0, 0};
  int next = 0, avail = 2;                          // two words are left over from the first draw
  for (uint32_t a2 = 0; a2 < 1024 && j < 0; ++a2) {
    if (next == avail) { philox4(seed ^ 0x9E3779B97F4A7C15ull, (uint32_t)t, ((uint32_t)(t >> 32) << 12) | (a2 & 0xfff), rr); next = 0; avail = 4; }
    const uint32_t word = next == 0 ? rr[0] : (next == 1 ? rr[1] : (next == 2 ? rr[2] : rr[3]));
    const int64_t cand = bounded(word, n_items);
    ++next;
    bool in = false;
    for (int q = beg; q < beg + deg; ++q) in |= ((int64_t)row[q] - n_users == cand);
    if (!in) j = cand;
  }
  if (j < 0) {
    // 1024 rejections in a row: this user has interacted with nearly every item (the reference's loop would spin as long).
    // Walk forward from a random item to the first one that is not a positive: always terminates unless EVERY item is one.
    philox4(seed ^ 0xD1B54A32D192ED03ull, (uint32_t)t, (uint32_t)(t >> 32), rr);
    int64_t cand = bounded(rr[0], n_items);
    for (int64_t step = 0; step < n_items && j < 0; ++step, cand = cand + 1 == n_items ? 0 : cand + 1) {
      bool in = false;
      for (int q = beg; q < beg + deg; ++q) in |= ((int64_t)row[q] - n_users == cand);
      if (!in) j = cand;
    }
  }
  if (j < 0) { atomicAdd(n_fail, 1); j = 0; }      // the user's positives cover the whole catalogue: no negative exists
  u_out[t] = u;
  i_out[t] = i;
  j_out[t] = j;
}
}  // namespace b200gat

extern "C" int b200gat_sample_bpr(const int32_t* colptr, const int32_t* row, int64_t n_users, int64_t n_items, int64_t n_samples,
                                  uint64_t seed, int64_t* u, int64_t* i, int64_t* j, int32_t* n_fail, void* stream) {
  return b200gat_sample_bpr_ex(colptr, row, n_users, n_items, n_samples, seed, nullptr, 0, u, i, j, n_fail, stream);
}

extern "C" int b200gat_sample_bpr_ex(const int32_t* colptr, const int32_t* row, int64_t n_users, int64_t n_items, int64_t n_samples,
                                     uint64_t seed, const int32_t* active_users, int64_t n_active, int64_t* u, int64_t* i, int64_t* j,
                                     int32_t* n_fail, void* stream) {
  B200GAT_CHECK_ARG(colptr && row && n_fail && (n_samples == 0 || (u && i && j)), "null pointer");
  B200GAT_CHECK_ARG(n_users > 0 && n_items > 0 && n_samples >= 0, "bad sizes");
  B200GAT_CHECK_ARG(!active_users || n_active > 0, "empty list of users with positives");
  cudaStream_t st = (cudaStream_t)stream;
  B200GAT_CUDA(cudaMemsetAsync(n_fail, 0, sizeof(int32_t), st));
  if (n_samples == 0) return kOk;
  count_launch(), sample_bpr_kernel<<<ceil_div(n_samples, 256), 256, 0, st>>>(colptr, row, n_users, n_items, n_samples, seed,
                                                                             active_users, n_active, u, i, j, n_fail);
  B200GAT_LAUNCH_CHECK();
  return kOk;
}
