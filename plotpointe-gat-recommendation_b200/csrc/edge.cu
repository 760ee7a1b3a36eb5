// Fused GAT edge kernels: one forward launch and one backward (CSC) launch per layer.
//
// Forward  (replaces scripts/train_gat_custom.py:79-92 and the PyG GATConv message passing called
//          at scripts/train_gat_pyg.py:87): one lane group per destination row over the CSR built in
//          graph.cu; logits from the per-node scalars s_src/s_dst; LeakyReLU; segment softmax with
//          shuffles (custom dialect: clamp[-10,10], no max, +1e-9; PyG dialect: online max,
//          +1e-16); 128-bit gathers of h[src]; weighted aggregation; head mean + bias.  No E-sized
//          float tensor is ever written.
// Backward (replaces autograd's replay of the same lines, train_gat_custom.py:361): one lane group per
//          SOURCE row over the CSC; gathers dout[dst] once per edge, recomputes alpha from per-node
//          scalars, accumulates dh[src] in registers (no atomics), emits the logit gradient
//          de per edge (E*H floats, the only E-sized scratch) and ds_src; a scalar CSR pass then
//          reduces de per destination into ds_dst.
//
// Two kernel families, both templated on the gathered dtype T (fp32 or bf16 rows; accumulation is always fp32):
//   * H*C == 128 (the headline shape): edge_fwd16_kernel / edge_bwd16_kernel, HALF a warp per row, a lane owns
//     channels [4 sl, 4 sl + 4) and [64 + 4 sl, 64 + 4 sl + 4) -- see the comment above them;
//   * any H in {1,2,4} x C in {128,256}: edge_fwd_kernel / edge_bwd_kernel, a warp per row, lane l owns channels
//     [4l, 4l+4) of every 128-wide chunk and head.
// Rows come from a descending-degree schedule; rows longer than 128 edges arrive as segments whose partial states a
// small combine kernel merges in order.
#include <cuda_bf16.h>

#include "common.cuh"
#include "../../include/b200gat.h"

namespace b200gat {

constexpr int kCustom = B200GAT_POLICY_CUSTOM;
constexpr int kPyG = B200GAT_POLICY_PYG;

// ---- counter-based RNG for attention dropout (Philox-4x32-10) -----------------------------------
// keyed on (seed, original edge id, head) so the forward (CSR order) and backward (CSC order)
// passes regenerate the same mask without storing it.
__device__ __forceinline__ uint32_t philox_uniform_bits(uint64_t seed, uint32_t edge_id, uint32_t head) {
  uint32_t c0 = edge_id, c1 = head, c2 = 0x9E3779B9u, c3 = 0xBB67AE85u;
  uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    c0 = hi1 ^ c1 ^ k0;
    c1 = lo1;
    c2 = hi0 ^ c3 ^ k1;
    c3 = lo0;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  return c0;
}
// keep-scale: 1/(1-p) with probability 1-p, else 0
__device__ __forceinline__ float dropout_scale(uint64_t seed, uint32_t edge_id, uint32_t head, float p, float inv_keep) {
  float u = (philox_uniform_bits(seed, edge_id, head) >> 8) * (1.0f / 16777216.0f);  // [0,1)
  return u >= p ? inv_keep : 0.0f;
}

template <int POLICY>
__device__ __forceinline__ float activate(float z0, float neg_slope) {
  float z = z0 > 0.f ? z0 : z0 * neg_slope;
  if (POLICY == kCustom) z = fminf(fmaxf(z, -10.f), 10.f);  // train_gat_custom.py:82
  return z;
}

// One lane's 4-channel piece of a feature row, as stored (fp32: 16 B, bf16: 8 B) and as computed on (float4).
template <typename T> struct Raw4;
template <> struct Raw4<float> { using type = float4; };
template <> struct Raw4<__nv_bfloat16> { using type = uint2; };
__device__ __forceinline__ float4 ld_raw4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ uint2 ld_raw4(const __nv_bfloat16* p) { return __ldg(reinterpret_cast<const uint2*>(p)); }
#ifndef EDGE_L2_HINTS
#define EDGE_L2_HINTS 1      // A/B knob: the half-warp backward marks what it touches once per launch (edge ids, the h row a
#endif                       // half-warp owns, the dh row it writes) evict-first in L2.  Measured at config 2, two runs each:
                             // backward 2.651 -> 2.628 ms per step (bf16 tier 1.937 -> 1.901); the same hints in the forward
                             // kernel cost 0.01 ms (its output is the next kernel's input) and are not applied
// once-per-launch variants of the loads above
__device__ __forceinline__ float4 ld_once4(const float* p) { return EDGE_L2_HINTS ? ld_cs4(p) : ld_raw4(p); }
__device__ __forceinline__ uint2 ld_once4(const __nv_bfloat16* p) { return EDGE_L2_HINTS ? ld_cs2u(p) : ld_raw4(p); }
__device__ __forceinline__ int ld_once(const int32_t* p) { return EDGE_L2_HINTS ? ld_cs_i32(p) : __ldg(p); }
__device__ __forceinline__ void st_once4(float* p, float4 v) {
  if (EDGE_L2_HINTS) st_cs4(p, v);
  else *reinterpret_cast<float4*>(p) = v;
}
__device__ __forceinline__ void st_once(float* p, float v) {
  if (EDGE_L2_HINTS) st_cs1(p, v);
  else *p = v;
}
__device__ __forceinline__ float4 to_f4(float4 v) { return v; }
__device__ __forceinline__ float4 to_f4(uint2 v) {   // bf16 -> fp32 is a 16-bit shift
  return make_float4(__uint_as_float(v.x << 16), __uint_as_float(v.x & 0xffff0000u), __uint_as_float(v.y << 16),
                     __uint_as_float(v.y & 0xffff0000u));
}


// --------------------------------------------------------------------------------------------
// row schedule: rows in descending-degree order, one int4 (row, beg, end, 0) each.  Persistent warps walk
// the schedule round-robin, so every warp gets the same mix of long and short rows (hubs first) and the next
// row's descriptor is a plain strided load that can be prefetched.
// --------------------------------------------------------------------------------------------
__global__ void degree_keys_kernel(const int32_t* __restrict__ ptr, int64_t n_rows, int32_t* __restrict__ keys) {
  for (int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; r < n_rows; r += (int64_t)gridDim.x * blockDim.x)
    keys[r] = ptr[r + 1] - ptr[r];
}
__global__ void build_schedule_kernel(const int32_t* __restrict__ ptr, const int32_t* __restrict__ asc_rows, int64_t n_rows,
                                      int4* __restrict__ sched) {
  for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < n_rows; k += (int64_t)gridDim.x * blockDim.x) {
    const int r = asc_rows ? asc_rows[n_rows - 1 - k] : (int)k;   // descending degree (or natural order)
    sched[k] = make_int4(r, ptr[r], ptr[r + 1], 0);
  }
}

// --------------------------------------------------------------------------------------------
// forward
// --------------------------------------------------------------------------------------------
constexpr int kEdgeThreads = 128;
// tuning knobs (A/B builds): rows per load group for the H*C == 128 instantiation, minimum resident blocks per SM
#ifndef EDGE_U_FWD
#define EDGE_U_FWD 4
#endif
#ifndef EDGE_U_BWD
#define EDGE_U_BWD 4
#endif
#ifndef EDGE_MINB
#define EDGE_MINB 8
#endif
// resident blocks per SM for 2 <= H*C/128 <= 4.  Config 3 (heads 4, bf16 rows), ms for both layers:
//   forward 4 blocks 7.28 | 5 blocks 6.29 | 6 blocks 6.23      backward 4 blocks 6.25 | 5 blocks 6.41 (spills) | 6 blocks 6.55
#ifndef EDGE_MINB_H4_FWD
#define EDGE_MINB_H4_FWD 5
#endif
#ifndef EDGE_MINB_H4_BWD
#define EDGE_MINB_H4_BWD 4
#endif

// > 0: the warp-per-row kernels (heads > 1 or 256 channels) stage their gathered rows in per-thread cp.async rings of that many load
// groups as well (see EDGE_FWD_RING below); 0 = two register-buffered groups.  Config 3 (heads 4, bf16 rows), ms for both layers
// (profiles/r2am_cfg3_ring_ab.txt): forward 6.29 -> 5.14 (ring 3, 5 blocks) | 4.99 (ring 3, 6 blocks) | 5.04 (ring 2, 6 blocks);
// backward 6.25 -> 4.59 (ring 4, 4 blocks) | 4.93 (ring 4, 5 blocks: spills) | 4.66 (ring 2, 5 blocks); step 19.0 -> 16.2 ms.
#ifndef EDGE_GEN_FWD_RING
#define EDGE_GEN_FWD_RING 3
#endif
#ifndef EDGE_GEN_BWD_RING
#define EDGE_GEN_BWD_RING 4
#endif

template <int BYTES>
__device__ __forceinline__ void cp_async_row(uint32_t dst_smem, const void* src) {
  if (BYTES == 16) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst_smem), "l"(src) : "memory");
  else asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst_smem), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

template <typename T, int H, int CV>
struct RowBuf {
  typename Raw4<T>::type v[H][CV];
};

// How a launch's result lands in `out`: out = (accumulate ? out : 0) + scale * (head mean) + bias.  The plain layer call is
// {1, 0}; per-head streaming (one launch per head of a heads > 1 layer, config 5) uses {1 / heads, head > 0} and passes the
// bias with the first head only.
struct OutMode {
  float scale;
  int accumulate;
};

template <int POLICY, int H, int CV>
__device__ __forceinline__ void finalize_row(int r, bool has_edges, const float (&m)[H], const float (&lt)[H], float4 (&acc)[H][CV],
                                             const float* __restrict__ bias, float* __restrict__ out,
                                             float* __restrict__ out_heads, float2* __restrict__ rowstat, int lane, OutMode om) {
  constexpr int C = CV * 128;
  constexpr int HC = H * C;
  float inv[H];
#pragma unroll
  for (int hh = 0; hh < H; ++hh) {
    inv[hh] = 1.f / (lt[hh] + (POLICY == kCustom ? 1e-9f : 1e-16f));
    if (lane == 0 && rowstat) rowstat[(size_t)r * H + hh] = make_float2(has_edges ? m[hh] : 0.f, inv[hh]);
  }
#pragma unroll
  for (int cv = 0; cv < CV; ++cv) {
    float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int hh = 0; hh < H; ++hh) {
      float4 a = acc[hh][cv];
      a.x *= inv[hh]; a.y *= inv[hh]; a.z *= inv[hh]; a.w *= inv[hh];
      if (out_heads) st_stream4(out_heads + (size_t)r * HC + hh * C + cv * 128 + lane * 4, a);
      o.x += a.x; o.y += a.y; o.z += a.z; o.w += a.w;
    }
    if (H > 1) { o.x *= 1.f / H; o.y *= 1.f / H; o.z *= 1.f / H; o.w *= 1.f / H; }
    if (om.scale != 1.f) { o.x *= om.scale; o.y *= om.scale; o.z *= om.scale; o.w *= om.scale; }
    if (bias) {
      const float4 b = ldg4(bias + cv * 128 + lane * 4);
      o.x += b.x; o.y += b.y; o.z += b.z; o.w += b.w;
    }
    float4* op = reinterpret_cast<float4*>(out + (size_t)r * C + cv * 128 + lane * 4);
    if (om.accumulate) {
      const float4 prev = *op;
      o.x += prev.x; o.y += prev.y; o.z += prev.z; o.w += prev.w;
    }
    *op = o;
  }
}

template <typename T, int POLICY, int H, int CV, bool DROPOUT>
__global__ void __launch_bounds__(kEdgeThreads, (H * CV == 1) ? EDGE_MINB : (H * CV <= 4 ? EDGE_MINB_H4_FWD : 2)) edge_fwd_kernel(const T* __restrict__ h, const float* __restrict__ s,
                                                                const int4* __restrict__ sched,
                                                                const int32_t* __restrict__ col,
                                                                const int32_t* __restrict__ perm, int n_rows, int row_offset,
                                                                float neg_slope, const float* __restrict__ bias,
                                                                float* __restrict__ out, float* __restrict__ out_heads,
                                                                float2* __restrict__ rowstat, float* __restrict__ partial,
                                                                float p_drop, uint64_t seed, OutMode om) {
  constexpr int C = CV * 128;
  constexpr int HC = H * C;
  // rows per load group (EDGE_GEN_FWD_RING groups in flight through the cp.async ring, or two in registers).  bf16 rows are half
  // the bytes, so twice the rows make a group
  constexpr int U = ((H * CV >= 4) ? 1 : (H * CV >= 2 ? 2 : EDGE_U_FWD)) * (sizeof(T) == 2 ? 2 : 1);
  const int lane = threadIdx.x & 31;
  const int idx = blockIdx.x * (kEdgeThreads / 32) + (threadIdx.x >> 5);
  if (idx >= n_rows) return;
  const float inv_keep = DROPOUT ? 1.f / (1.f - p_drop) : 1.f;
  const int4 d = __ldg(sched + idx);

  auto load_group = [&](RowBuf<T, H, CV>(&buf)[U], int c, int k) {
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int ck = __shfl_sync(kFull, c, (k + u) & 31);
      const T* hp = h + (size_t)ck * HC + lane * 4;
#pragma unroll
      for (int hh = 0; hh < H; ++hh)
#pragma unroll
        for (int cv = 0; cv < CV; ++cv) buf[u].v[hh][cv] = ld_raw4(hp + hh * C + cv * 128);
    }
  };

  {
    const int r = d.x, beg = d.y, end = d.z;

    float sd[H], m[H], l[H];
    float4 acc[H][CV];
#pragma unroll
    for (int hh = 0; hh < H; ++hh) {
      sd[hh] = __ldg(s + (size_t)(row_offset + r) * (2 * H) + H + hh);
      m[hh] = POLICY == kPyG ? -INFINITY : 0.f;
      l[hh] = 0.f;
#pragma unroll
      for (int cv = 0; cv < CV; ++cv) acc[hh][cv] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    for (int base = beg; base < end; base += 32) {
      const int e = base + lane;
      const bool valid = e < end;
      int c = valid ? __ldg(col + e) : 0;
      float sv[H];
#pragma unroll
      for (int hh = 0; hh < H; ++hh) sv[hh] = valid ? __ldg(s + (size_t)c * (2 * H) + hh) : 0.f;
      // zero-weight tail lanes re-read lane 0's row so they add no new cache lines
      const int c0 = __shfl_sync(kFull, c, 0);
      if (!valid) c = c0;
      const int cnt = min(32, end - base);
#if EDGE_GEN_FWD_RING > 0
      constexpr int NB = EDGE_GEN_FWD_RING;
      using Raw = typename Raw4<T>::type;
      __shared__ Raw ring[NB][U][H][CV][kEdgeThreads];     // private per thread; one commit per call whether or not anything is copied
      auto ring_load = [&](int b, int k, bool on) {
        if (on) {
#pragma unroll
          for (int u = 0; u < U; ++u) {
            const int ck = __shfl_sync(kFull, c, (k + u) & 31);
            const T* hp = h + (size_t)ck * HC + lane * 4;
#pragma unroll
            for (int hh = 0; hh < H; ++hh)
#pragma unroll
              for (int cv = 0; cv < CV; ++cv)
                cp_async_row<sizeof(Raw)>((uint32_t)__cvta_generic_to_shared(&ring[b][u][hh][cv][threadIdx.x]), hp + hh * C + cv * 128);
          }
        }
        cp_async_commit();
      };
#pragma unroll
      for (int b = 0; b < NB; ++b) ring_load(b, b * U, b * U < cnt);
#else
      RowBuf<T, H, CV> bufA[U], bufB[U];
      load_group(bufA, c, 0);                              // gathers start before the softmax math
      if (U < cnt) load_group(bufB, c, U);
#endif

      float p[H];
#pragma unroll
      for (int hh = 0; hh < H; ++hh) {
        float z = -INFINITY;
        if (valid) z = activate<POLICY>(sv[hh] + sd[hh], neg_slope);
        if (POLICY == kPyG) {
          const float nm = fmaxf(m[hh], warp_max(z));
          if (nm != m[hh]) {  // warp-uniform
            const float scale = expf(m[hh] - nm);
            l[hh] *= scale;
#pragma unroll
            for (int cv = 0; cv < CV; ++cv) {
              acc[hh][cv].x *= scale; acc[hh][cv].y *= scale; acc[hh][cv].z *= scale; acc[hh][cv].w *= scale;
            }
            m[hh] = nm;
          }
          p[hh] = valid ? expf(z - nm) : 0.f;
        } else {
          p[hh] = valid ? expf(z) : 0.f;
        }
        l[hh] += p[hh];  // the denominator sees every edge; dropout acts on alpha afterwards (:88-89)
        if (DROPOUT && valid) p[hh] *= dropout_scale(seed, (uint32_t)__ldg(perm + e), hh, p_drop, inv_keep);
      }
#if EDGE_GEN_FWD_RING > 0
      for (int k = 0; k < cnt; k += NB * U) {
#pragma unroll
        for (int b = 0; b < NB; ++b) {
          if (k + b * U < cnt) {                           // warp-uniform
            cp_async_wait<NB - 1>();
#pragma unroll
            for (int u = 0; u < U; ++u)
#pragma unroll
              for (int hh = 0; hh < H; ++hh) {
                const float pk = __shfl_sync(kFull, p[hh], (k + b * U + u) & 31);
#pragma unroll
                for (int cv = 0; cv < CV; ++cv) acc[hh][cv] = fma4(pk, to_f4(ring[b][u][hh][cv][threadIdx.x]), acc[hh][cv]);
              }
            ring_load(b, k + (b + NB) * U, k + (b + NB) * U < cnt);
          }
        }
      }
      cp_async_wait<0>();
#else
      auto consume = [&](RowBuf<T, H, CV>(&buf)[U], int k) {
#pragma unroll
        for (int u = 0; u < U; ++u)
#pragma unroll
          for (int hh = 0; hh < H; ++hh) {
            const float pk = __shfl_sync(kFull, p[hh], (k + u) & 31);
#pragma unroll
            for (int cv = 0; cv < CV; ++cv) acc[hh][cv] = fma4(pk, to_f4(buf[u].v[hh][cv]), acc[hh][cv]);
          }
      };
      for (int k = 0; k < cnt; k += 2 * U) {
        consume(bufA, k);
        if (k + 2 * U < cnt) load_group(bufA, c, k + 2 * U);
        if (k + U < cnt) {
          consume(bufB, k + U);
          if (k + 3 * U < cnt) load_group(bufB, c, k + 3 * U);
        }
      }
#endif
    }
    float lt[H];
#pragma unroll
    for (int hh = 0; hh < H; ++hh) lt[hh] = warp_sum(l[hh]);
    if (d.w > 0) {  // segment of a split (long) row: park the un-normalised state, fwd_combine_kernel finishes the row
      float* ps = partial + (size_t)(d.w - 1) * (H * (C + 4));
#pragma unroll
      for (int hh = 0; hh < H; ++hh) {
        if (lane == 0) { ps[hh * 4] = m[hh]; ps[hh * 4 + 1] = lt[hh]; }
#pragma unroll
        for (int cv = 0; cv < CV; ++cv)
          *reinterpret_cast<float4*>(ps + 4 * H + hh * C + cv * 128 + lane * 4) = acc[hh][cv];
      }
      return;
    }
    finalize_row<POLICY, H, CV>(r, beg < end, m, lt, acc, bias, out, out_heads, rowstat, lane, om);
  }
}

// merges the segments of the split rows in segment order (deterministic) and finishes them
template <int POLICY, int H, int CV>
__global__ void __launch_bounds__(kEdgeThreads) fwd_combine_kernel(const float* __restrict__ partial, const int4* __restrict__ table,
                                                                   int n_long, const float* __restrict__ bias,
                                                                   float* __restrict__ out, float* __restrict__ out_heads,
                                                                   float2* __restrict__ rowstat, OutMode om) {
  constexpr int C = CV * 128;
  const int lane = threadIdx.x & 31;
  const int idx = blockIdx.x * (kEdgeThreads / 32) + (threadIdx.x >> 5);
  if (idx >= n_long) return;
  const int4 t = __ldg(table + idx);   // (row, first slot, n segments, degree)
  float m[H], lt[H];
  float4 acc[H][CV];
#pragma unroll
  for (int hh = 0; hh < H; ++hh) {
    m[hh] = POLICY == kPyG ? -INFINITY : 0.f;
    lt[hh] = 0.f;
#pragma unroll
    for (int cv = 0; cv < CV; ++cv) acc[hh][cv] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  for (int k = 0; k < t.z; ++k) {
    const float* ps = partial + (size_t)(t.y + k) * (H * (C + 4));
#pragma unroll
    for (int hh = 0; hh < H; ++hh) {
      const float mk = ps[hh * 4], lk = ps[hh * 4 + 1];
      float so = 1.f, sn = 1.f;
      if (POLICY == kPyG) {
        const float nm = fmaxf(m[hh], mk);
        so = expf(m[hh] - nm);
        sn = expf(mk - nm);
        m[hh] = nm;
      }
      lt[hh] = lt[hh] * so + lk * sn;
#pragma unroll
      for (int cv = 0; cv < CV; ++cv) {
        const float4 a = *reinterpret_cast<const float4*>(ps + 4 * H + hh * C + cv * 128 + lane * 4);
        acc[hh][cv].x = acc[hh][cv].x * so + a.x * sn;
        acc[hh][cv].y = acc[hh][cv].y * so + a.y * sn;
        acc[hh][cv].z = acc[hh][cv].z * so + a.z * sn;
        acc[hh][cv].w = acc[hh][cv].w * so + a.w * sn;
      }
    }
  }
  finalize_row<POLICY, H, CV>(t.x, true, m, lt, acc, bias, out, out_heads, rowstat, lane, om);
}

// --------------------------------------------------------------------------------------------
// backward, step 0: per-destination scalars  nodestat[i,h] = (s_dst, m, 1/D, t)
//   t[i,h] = (1/H) * dout[i,:] . out_h[i,h,:]   (= sum_k alpha_ik dalpha_ik)
// --------------------------------------------------------------------------------------------
template <int H, int CV>
__global__ void __launch_bounds__(256) node_prep_kernel(const float* __restrict__ dout,       // [n_rows, C]
                                                        const float* __restrict__ out_heads,  // [n_rows, H, C]
                                                        const float* __restrict__ bias,       // subtracted (H==1, PyG)
                                                        const float* __restrict__ s, const float2* __restrict__ rowstat,
                                                        int n_rows, int row_offset, float4* __restrict__ nodestat,
                                                        float* __restrict__ colsum_part /*[gridDim.x, C] or null*/,
                                                        __nv_bfloat16* __restrict__ dout_bf16 /*[n_rows, C] or null*/) {
  constexpr int C = CV * 128;
  __shared__ float4 red[8][CV][32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float4 csum[CV];
  float4 b[CV];
#pragma unroll
  for (int cv = 0; cv < CV; ++cv) {
    csum[cv] = make_float4(0.f, 0.f, 0.f, 0.f);
    b[cv] = bias ? ldg4(bias + cv * 128 + lane * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  for (int r = blockIdx.x * 8 + warp; r < n_rows; r += gridDim.x * 8) {   // fixed row -> warp map: deterministic sums
    float t[H];
#pragma unroll
    for (int hh = 0; hh < H; ++hh) t[hh] = 0.f;
#pragma unroll
    for (int cv = 0; cv < CV; ++cv) {
      const float4 g = ld_stream4(dout + (size_t)r * C + cv * 128 + lane * 4);
      if (dout_bf16) {
        const __nv_bfloat162 lo = __floats2bfloat162_rn(g.x, g.y), hi = __floats2bfloat162_rn(g.z, g.w);
        uint2 pk;
        pk.x = *reinterpret_cast<const uint32_t*>(&lo);
        pk.y = *reinterpret_cast<const uint32_t*>(&hi);
        *reinterpret_cast<uint2*>(dout_bf16 + (size_t)r * C + cv * 128 + lane * 4) = pk;
      }
      csum[cv].x += g.x; csum[cv].y += g.y; csum[cv].z += g.z; csum[cv].w += g.w;
#pragma unroll
      for (int hh = 0; hh < H; ++hh) {
        float4 o = ld_stream4(out_heads + ((size_t)r * H + hh) * C + cv * 128 + lane * 4);
        o.x -= b[cv].x; o.y -= b[cv].y; o.z -= b[cv].z; o.w -= b[cv].w;
        t[hh] += dot4(g, o);
      }
    }
#pragma unroll
    for (int hh = 0; hh < H; ++hh) {
      const float tt = warp_sum(t[hh]) * (1.f / H);
      if (lane == 0) {
        const float2 rs = rowstat[(size_t)r * H + hh];
        nodestat[(size_t)r * H + hh] = make_float4(s[(size_t)(row_offset + r) * (2 * H) + H + hh], rs.x, rs.y, tt);
      }
    }
  }
  if (colsum_part) {
#pragma unroll
    for (int cv = 0; cv < CV; ++cv) red[warp][cv][lane] = csum[cv];
    __syncthreads();
    if (warp == 0) {
#pragma unroll
      for (int cv = 0; cv < CV; ++cv) {
        float4 a = red[0][cv][lane];
#pragma unroll
        for (int w = 1; w < 8; ++w) {
          const float4 v = red[w][cv][lane];
          a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
        }
        *reinterpret_cast<float4*>(colsum_part + (size_t)blockIdx.x * C + cv * 128 + lane * 4) = a;
      }
    }
  }
}

// out[c] = sum over parts; 8 interleaved stripes per column, combined in a fixed order
// out[c] = sum_k part[k, c] in a fixed order: 8 stripes of partial rows, then the stripes in order.  One block per 32 columns
// (a single 1024-thread block took 34 us for 1184 partial rows).
__global__ void __launch_bounds__(256) colsum_finish_kernel(const float* __restrict__ part, int n_parts, int C, float* __restrict__ out) {
  __shared__ float red[8][32];
  const int lc = threadIdx.x & 31, stripe = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + lc;
  float a = 0.f;
  if (c < C)
    for (int k = stripe; k < n_parts; k += 8) a += part[(size_t)k * C + c];
  red[stripe][lc] = a;
  __syncthreads();
  if (stripe == 0 && c < C) {
    float t = red[0][lc];
#pragma unroll
    for (int w = 1; w < 8; ++w) t += red[w][lc];
    out[c] = t;
  }
}

// --------------------------------------------------------------------------------------------
// backward, step 1: CSC pass (persistent warps over the source-row schedule)
// --------------------------------------------------------------------------------------------
// TP ("two-phase", per-head streaming without saved per-head outputs): t[i] = sum_k alpha_ik dalpha_ik is not known yet, so
// this pass emits w_ij = alpha_ij * dalpha_ij into `de` (nodestat.w is ignored, ds_src is not meaningful); the caller reduces w
// per destination into t, and edge_de_kernel turns w into de = slope * (w - alpha * t) in a second, scalar pass.
template <typename T, int POLICY, int H, int CV, bool DROPOUT, bool TP = false>
__global__ void __launch_bounds__(kEdgeThreads, (H * CV == 1) ? EDGE_MINB : (H * CV <= 4 ? EDGE_MINB_H4_BWD : 2)) edge_bwd_kernel(const T* __restrict__ h, const float* __restrict__ s,
                                                                const T* __restrict__ dout,
                                                                const float4* __restrict__ nodestat,
                                                                const int4* __restrict__ sched,
                                                                const int32_t* __restrict__ row,
                                                                const int32_t* __restrict__ perm_csc, int n_rows, int row_offset,
                                                                float neg_slope, float* __restrict__ dh,
                                                                float* __restrict__ de, float* __restrict__ ds_src,
                                                                int ld_ds, float* __restrict__ partial, float p_drop,
                                                                uint64_t seed) {
  constexpr int C = CV * 128;
  constexpr int HC = H * C;
  constexpr int U = (CV >= 2 ? 2 : EDGE_U_BWD) * (sizeof(T) == 2 ? 2 : 1);   // dout rows per load group; EDGE_GEN_BWD_RING groups in flight (ring) or two (registers)
  const int lane = threadIdx.x & 31;
  const int idx = blockIdx.x * (kEdgeThreads / 32) + (threadIdx.x >> 5);
  if (idx >= n_rows) return;
  const float inv_keep = DROPOUT ? 1.f / (1.f - p_drop) : 1.f;
  constexpr float invH = 1.f / H;
  const int4 d = __ldg(sched + idx);

  struct GBuf { typename Raw4<T>::type g[CV]; };
  auto load_group = [&](GBuf(&buf)[U], int i, int k) {
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int ik = __shfl_sync(kFull, i, (k + u) & 31);
#pragma unroll
      for (int cv = 0; cv < CV; ++cv) buf[u].g[cv] = ld_raw4(dout + (size_t)ik * C + cv * 128 + lane * 4);
    }
  };

  {
    const int r = d.x, beg = d.y, end = d.z;
    const size_t j = (size_t)row_offset + r;

    float4 hj[H][CV], acc[H][CV];
    float ssj[H], dss[H];
#pragma unroll
    for (int hh = 0; hh < H; ++hh) {
      ssj[hh] = __ldg(s + j * (2 * H) + hh);
      dss[hh] = 0.f;
#pragma unroll
      for (int cv = 0; cv < CV; ++cv) {
        hj[hh][cv] = (beg < end) ? to_f4(ld_raw4(h + j * HC + hh * C + cv * 128 + lane * 4)) : make_float4(0.f, 0.f, 0.f, 0.f);
        acc[hh][cv] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
    for (int base = beg; base < end; base += 32) {
      const int q = base + lane;
      const bool valid = q < end;
      int i = valid ? __ldg(row + q) : 0;
      const int i0 = __shfl_sync(kFull, i, 0);
      if (!valid) i = i0;
      const int cnt = min(32, end - base);
#if EDGE_GEN_BWD_RING > 0
      constexpr int NB = EDGE_GEN_BWD_RING;
      using Raw = typename Raw4<T>::type;
      __shared__ Raw ring[NB][U][CV][kEdgeThreads];
      auto ring_load = [&](int b, int k, bool on) {
        if (on) {
#pragma unroll
          for (int u = 0; u < U; ++u) {
            const int ik = __shfl_sync(kFull, i, (k + u) & 31);
#pragma unroll
            for (int cv = 0; cv < CV; ++cv)
              cp_async_row<sizeof(Raw)>((uint32_t)__cvta_generic_to_shared(&ring[b][u][cv][threadIdx.x]), dout + (size_t)ik * C + cv * 128 + lane * 4);
          }
        }
        cp_async_commit();
      };
#pragma unroll
      for (int b = 0; b < NB; ++b) ring_load(b, b * U, b * U < cnt);
#else
      GBuf bufA[U], bufB[U];
      load_group(bufA, i, 0);                      // dout gathers start before the per-edge scalar math
      if (U < cnt) load_group(bufB, i, U);
#endif

      // per lane (= per edge of this chunk): agg = alpha' / H (weight of dout_i in dh_j), and the two coefficients of
      // de = alpha * (dalpha * keep - t) * slope  written as  de = gA * <dout_i, h_j> - cB
      float agg[H], gA[H], cB[H], my_de[H];
#pragma unroll
      for (int hh = 0; hh < H; ++hh) {
        agg[hh] = gA[hh] = cB[hh] = my_de[hh] = 0.f;
        if (valid) {
          const float4 st = __ldg(nodestat + (size_t)i * H + hh);  // (s_dst, m, 1/D, t)
          const float z0 = ssj[hh] + st.x;
          const float zl = z0 > 0.f ? z0 : z0 * neg_slope;
          float zc = zl, pass = 1.f;
          if (POLICY == kCustom) {
            zc = fminf(fmaxf(zl, -10.f), 10.f);
            pass = (zl >= -10.f && zl <= 10.f) ? 1.f : 0.f;  // clamp passes gradient only inside [-10,10]
          }
          const float alpha = expf(zc - st.y) * st.z;
          const float gsc = (z0 > 0.f ? 1.f : neg_slope) * pass;
          float ks = 1.f;
          if (DROPOUT) ks = dropout_scale(seed, (uint32_t)__ldg(perm_csc + q), hh, p_drop, inv_keep);
          agg[hh] = alpha * ks * invH;
          gA[hh] = agg[hh] * gsc;
          cB[hh] = alpha * st.w * gsc;
        }
      }
#if EDGE_GEN_BWD_RING > 0
      auto consume = [&](int b, int k) {
        // dalpha of the U edges of this group: per-lane partial dots first, then ONE multi-value warp reduction per head
        // (7 shuffles for 4 edges instead of 20), then each edge's owner lane fetches its total.
        float dsum[H][U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int kk = (k + u) & 31;
#pragma unroll
          for (int hh = 0; hh < H; ++hh) {
            const float a = __shfl_sync(kFull, agg[hh], kk);
            float dp = 0.f;
#pragma unroll
            for (int cv = 0; cv < CV; ++cv) {
              const float4 gv = to_f4(ring[b][u][cv][threadIdx.x]);
              acc[hh][cv] = fma4(a, gv, acc[hh][cv]);
              dp += dot4(hj[hh][cv], gv);
            }
            dsum[hh][u] = dp;
          }
        }
        const int mine = (lane - k) & 31;                   // which edge of the group this lane owns (if < U)
        const bool owner = mine < U && k + mine < cnt;
#pragma unroll
        for (int hh = 0; hh < H; ++hh) {
          const float tot = warp_multi_sum<U>(dsum[hh], lane);
          const float dot = __shfl_sync(kFull, tot, (mine & (U - 1)) << (5 - Log2<U>::value));   // <dout_i, h_j> of my edge
          if (owner) my_de[hh] = TP ? agg[hh] * dot : gA[hh] * dot - cB[hh];
        }
      };
      for (int k = 0; k < cnt; k += NB * U) {
#pragma unroll
        for (int b = 0; b < NB; ++b) {
          if (k + b * U < cnt) {                       // warp-uniform
            cp_async_wait<NB - 1>();
            consume(b, k + b * U);
            ring_load(b, k + (b + NB) * U, k + (b + NB) * U < cnt);
          }
        }
      }
      cp_async_wait<0>();
#else
      auto consume = [&](GBuf(&buf)[U], int k) {
        // dalpha of the U edges of this group: per-lane partial dots first, then ONE multi-value warp reduction per head
        // (7 shuffles for 4 edges instead of 20), then each edge's owner lane fetches its total.
        float dsum[H][U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int kk = (k + u) & 31;
#pragma unroll
          for (int hh = 0; hh < H; ++hh) {
            const float a = __shfl_sync(kFull, agg[hh], kk);
            float dp = 0.f;
#pragma unroll
            for (int cv = 0; cv < CV; ++cv) {
              const float4 gv = to_f4(buf[u].g[cv]);
              acc[hh][cv] = fma4(a, gv, acc[hh][cv]);
              dp += dot4(hj[hh][cv], gv);
            }
            dsum[hh][u] = dp;
          }
        }
        const int mine = (lane - k) & 31;                   // which edge of the group this lane owns (if < U)
        const bool owner = mine < U && k + mine < cnt;
#pragma unroll
        for (int hh = 0; hh < H; ++hh) {
          const float tot = warp_multi_sum<U>(dsum[hh], lane);
          const float dot = __shfl_sync(kFull, tot, (mine & (U - 1)) << (5 - Log2<U>::value));   // <dout_i, h_j> of my edge
          if (owner) my_de[hh] = TP ? agg[hh] * dot : gA[hh] * dot - cB[hh];
        }
      };
      for (int k = 0; k < cnt; k += 2 * U) {
        consume(bufA, k);
        if (k + 2 * U < cnt) load_group(bufA, i, k + 2 * U);
        if (k + U < cnt) {
          consume(bufB, k + U);
          if (k + 3 * U < cnt) load_group(bufB, i, k + 3 * U);
        }
      }
#endif
      if (valid) {
#pragma unroll
        for (int hh = 0; hh < H; ++hh) {
          de[(size_t)q * H + hh] = my_de[hh];
          dss[hh] += my_de[hh];
        }
      }
    }
    if (d.w > 0) {  // segment of a split (long) row: bwd_combine_kernel sums the segments in order
      float* ps = partial + (size_t)(d.w - 1) * (HC + 4);
#pragma unroll
      for (int hh = 0; hh < H; ++hh) {
        const float t = warp_sum(dss[hh]);
        if (lane == 0) ps[hh] = t;
#pragma unroll
        for (int cv = 0; cv < CV; ++cv) *reinterpret_cast<float4*>(ps + 4 + hh * C + cv * 128 + lane * 4) = acc[hh][cv];
      }
      return;
    }
#pragma unroll
    for (int hh = 0; hh < H; ++hh) {
      const float t = warp_sum(dss[hh]);
      if (lane == 0) ds_src[(size_t)r * ld_ds + hh] = t;
#pragma unroll
      for (int cv = 0; cv < CV; ++cv)
        *reinterpret_cast<float4*>(dh + (size_t)r * HC + hh * C + cv * 128 + lane * 4) = acc[hh][cv];
    }
  }
}

template <int H, int CV>
__global__ void __launch_bounds__(kEdgeThreads) bwd_combine_kernel(const float* __restrict__ partial, const int4* __restrict__ table,
                                                                   int n_long, float* __restrict__ dh, float* __restrict__ ds_src,
                                                                   int ld_ds) {
  constexpr int C = CV * 128;
  constexpr int HC = H * C;
  const int lane = threadIdx.x & 31;
  const int idx = blockIdx.x * (kEdgeThreads / 32) + (threadIdx.x >> 5);
  if (idx >= n_long) return;
  const int4 t = __ldg(table + idx);
  float dss[H];
  float4 acc[H][CV];
#pragma unroll
  for (int hh = 0; hh < H; ++hh) {
    dss[hh] = 0.f;
#pragma unroll
    for (int cv = 0; cv < CV; ++cv) acc[hh][cv] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  for (int k = 0; k < t.z; ++k) {
    const float* ps = partial + (size_t)(t.y + k) * (HC + 4);
#pragma unroll
    for (int hh = 0; hh < H; ++hh) {
      dss[hh] += ps[hh];
#pragma unroll
      for (int cv = 0; cv < CV; ++cv) {
        const float4 a = *reinterpret_cast<const float4*>(ps + 4 + hh * C + cv * 128 + lane * 4);
        acc[hh][cv].x += a.x; acc[hh][cv].y += a.y; acc[hh][cv].z += a.z; acc[hh][cv].w += a.w;
      }
    }
  }
#pragma unroll
  for (int hh = 0; hh < H; ++hh) {
    if (lane == 0) ds_src[(size_t)t.x * ld_ds + hh] = dss[hh];
#pragma unroll
    for (int cv = 0; cv < CV; ++cv)
      *reinterpret_cast<float4*>(dh + (size_t)t.x * HC + hh * C + cv * 128 + lane * 4) = acc[hh][cv];
  }
}

// --------------------------------------------------------------------------------------------
// backward, step 2: ds_dst[i,h] = sum over in-edges of de (CSR order, via the CSR->CSC map)
// --------------------------------------------------------------------------------------------
// W lanes per row (32 / W rows per warp): a row is a chain rowptr -> csr2csc -> de, so rows in flight set the speed
template <int H, int W>
__global__ void __launch_bounds__(128) ds_dst_kernel(const float* __restrict__ de, const int32_t* __restrict__ rowptr,
                                                     const int32_t* __restrict__ csr2csc, int n_rows,
                                                     float* __restrict__ ds_dst, int ld_ds) {
  const int sl = threadIdx.x & (W - 1);
  const int64_t r = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) / W;
  const bool live = r < n_rows;                      // no early return: the shuffles below are warp-wide
  const int beg = live ? rowptr[r] : 0, end = live ? rowptr[r + 1] : 0;
  float a[H];
#pragma unroll
  for (int hh = 0; hh < H; ++hh) a[hh] = 0.f;
  for (int e = beg + sl; e < end; e += W) {
    const size_t q = (size_t)csr2csc[e];
#pragma unroll
    for (int hh = 0; hh < H; ++hh) a[hh] += de[q * H + hh];
  }
#pragma unroll
  for (int hh = 0; hh < H; ++hh) {
    float t = a[hh];
#pragma unroll
    for (int o = W / 2; o > 0; o >>= 1) t += __shfl_xor_sync(kFull, t, o);
    if (live && sl == 0) ds_dst[(size_t)r * ld_ds + hh] = t;
  }
}

// thread-per-row variant for short rows (average degree < 8, e.g. one rank's slice of the backward graph)
template <int H>
__global__ void __launch_bounds__(256) ds_dst_thread_kernel(const float* __restrict__ de, const int32_t* __restrict__ rowptr,
                                                            const int32_t* __restrict__ csr2csc, int n_rows,
                                                            float* __restrict__ ds_dst, int ld_ds) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n_rows) return;
  const int beg = rowptr[r], end = rowptr[r + 1];
  float a[H];
#pragma unroll
  for (int hh = 0; hh < H; ++hh) a[hh] = 0.f;
  for (int e = beg; e < end; ++e) {
    const size_t q = (size_t)csr2csc[e];
#pragma unroll
    for (int hh = 0; hh < H; ++hh) a[hh] += de[q * H + hh];
  }
#pragma unroll
  for (int hh = 0; hh < H; ++hh) ds_dst[(size_t)r * ld_ds + hh] = a[hh];
}

// second phase of the two-phase backward: de_ij = slope_ij * (w_ij - alpha_ij * t_i) in place over w (CSC order), and
// ds_src[j] = sum over the out-edges of j.  nodestat[i] = (s_dst, m, 1/D, t) now carries t.  W lanes per source row.
template <int POLICY, int H, int W>
__global__ void __launch_bounds__(128) edge_de_kernel(float* __restrict__ de, const int32_t* __restrict__ colptr,
                                                      const int32_t* __restrict__ row, const float* __restrict__ s,
                                                      const float4* __restrict__ nodestat, int n_rows, int row_offset,
                                                      float neg_slope, float* __restrict__ ds_src, int ld_ds) {
  const int sl = threadIdx.x & (W - 1);
  const int64_t r = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) / W;
  const bool live = r < n_rows;
  const int beg = live ? colptr[r] : 0, end = live ? colptr[r + 1] : 0;
  float ssj[H], a[H];
#pragma unroll
  for (int hh = 0; hh < H; ++hh) {
    ssj[hh] = live ? __ldg(s + (size_t)(row_offset + r) * (2 * H) + hh) : 0.f;
    a[hh] = 0.f;
  }
  for (int q = beg + sl; q < end; q += W) {
    const int i = __ldg(row + q);
#pragma unroll
    for (int hh = 0; hh < H; ++hh) {
      const float4 st = __ldg(nodestat + (size_t)i * H + hh);
      const float z0 = ssj[hh] + st.x;
      const float zl = z0 > 0.f ? z0 : z0 * neg_slope;
      float zc = zl, pass = 1.f;
      if (POLICY == kCustom) {
        zc = fminf(fmaxf(zl, -10.f), 10.f);
        pass = (zl >= -10.f && zl <= 10.f) ? 1.f : 0.f;
      }
      const float alpha = expf(zc - st.y) * st.z;
      const float gsc = (z0 > 0.f ? 1.f : neg_slope) * pass;
      const float d = gsc * (de[(size_t)q * H + hh] - alpha * st.w);
      de[(size_t)q * H + hh] = d;
      a[hh] += d;
    }
  }
#pragma unroll
  for (int hh = 0; hh < H; ++hh) {
    float t = a[hh];
#pragma unroll
    for (int o = W / 2; o > 0; o >>= 1) t += __shfl_xor_sync(kFull, t, o);
    if (live && sl == 0) ds_src[(size_t)r * ld_ds + hh] = t;
  }
}

// nodestat[r, h].w = t[r, h]  (the per-destination sums of the first phase, written into the slot the second phase reads)
__global__ void set_t_kernel(float4* __restrict__ nodestat, const float* __restrict__ t, int64_t n) {
  const int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (k < n) nodestat[k].w = t[k];
}

// nodestat[r, h] = (s_dst, m, 1/D, 0) from the saved per-node scalars (no per-head outputs needed), and an optional scaled bf16
// copy of dout for the gathers
__global__ void node_stat_kernel(const float* __restrict__ s, const float2* __restrict__ rowstat, int64_t n_rows, int64_t row_offset,
                                 int H, float4* __restrict__ nodestat) {
  const int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (k >= n_rows * H) return;
  const int64_t r = k / H;
  const int hh = (int)(k - r * H);
  const float2 rs = rowstat[k];
  nodestat[k] = make_float4(s[(size_t)(row_offset + r) * (2 * H) + H + hh], rs.x, rs.y, 0.f);
}
__global__ void cast_bf16_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, int64_t n4, float scale) {
  for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < n4; k += (int64_t)gridDim.x * blockDim.x) {
    const float4 v = ld_stream4(src + 4 * k);
    const __nv_bfloat162 lo = __floats2bfloat162_rn(v.x * scale, v.y * scale), hi = __floats2bfloat162_rn(v.z * scale, v.w * scale);
    uint2 pk;
    pk.x = *reinterpret_cast<const uint32_t*>(&lo);
    pk.y = *reinterpret_cast<const uint32_t*>(&hi);
    *reinterpret_cast<uint2*>(dst + 4 * k) = pk;
  }
}

// --------------------------------------------------------------------------------------------
// H * C == 128 (the headline shape): HALF a warp per scheduled row.
// Rows are short (mean in-degree 19, a third of them below 10), so what bounds the warp-per-row kernels is the
// number of rows being worked on at once (every row is a chain of dependent loads: schedule entry -> edge ids ->
// feature rows), not the bytes each warp keeps in flight.  Here a warp carries two neighbouring schedule entries
// (neighbours in the degree-sorted schedule have almost the same length), 16 lanes each; a lane owns channels
// [4 sl, 4 sl + 4) and [64 + 4 sl, 64 + 4 sl + 4), so both 128-bit loads of a row stay fully coalesced.  Bytes in
// flight per warp and registers per lane are unchanged; the number of concurrent row chains doubles.
// Every shuffle is executed by the full warp (width 16 keeps the data inside each half), so trip counts are the
// maximum over the two halves and the shorter half is predicated off.
// --------------------------------------------------------------------------------------------
#ifndef EDGE_W
#define EDGE_W 16      // 16: half-warp kernels for H*C == 128; 32: the generic warp-per-row kernels everywhere
#endif
// Tuning knobs, set from A/B builds on config 2 (ms for both layers; generic warp-per-row kernels: fp32 fwd 2.63, bwd 3.32).
// U = rows per load group and half (x2 for bf16 rows), NB = load groups in flight, MINB = resident blocks per SM.
//   fp32 fwd: U1 NB2 MINB9 1.93 | U1 NB4 MINB8 1.86 | U1 NB3 MINB9 2.02 | U2 NB2 MINB8 2.42 | U1 NB2 MINB10 2.17 | U1 NB8 MINB6 2.99
//   fp32 bwd: U1 NB2 MINB9 2.72 | U1 NB1 MINB10 2.65 | U1 NB1 MINB9 2.70 | U1 NB1 MINB12 2.74 | U2 NB2 MINB8 3.33 | U1 NB1 MINB16 3.39
//   bf16 fwd: U2 NB2 MINB9 1.67 | U2 NB4 MINB8 1.84          bf16 bwd: U2 NB2 MINB9 2.21 | U2 NB1 MINB10 2.03
// Fewer bytes in flight per row chain and more chains win: the rows are short, and spills or a lower block count cost more
// than the extra loads in flight buy.
#ifndef EDGE_U16_FWD
#define EDGE_U16_FWD 1
#endif
#ifndef EDGE_U16_BWD
#define EDGE_U16_BWD 1
#endif
#ifndef EDGE_NB16_FWD
#define EDGE_NB16_FWD 4
#endif
#ifndef EDGE_NB16_FWD_BF16
#define EDGE_NB16_FWD_BF16 2
#endif
#ifndef EDGE_NB16_BWD
#define EDGE_NB16_BWD 1
#endif
#ifndef EDGE_MINB16_FWD
#define EDGE_MINB16_FWD 12     // 8 with register buffers (EDGE_FWD_RING 0)
#endif
#ifndef EDGE_MINB16_FWD_BF16
#define EDGE_MINB16_FWD_BF16 12   // 9 with register buffers
#endif
#ifndef EDGE_MINB16_BWD
#define EDGE_MINB16_BWD 10
#endif
// > 0 = the gathered rows (h in the forward, dout in the backward) go through a per-thread ring of that many load groups in
// shared memory, filled by cp.async (LDGSTS): every lane reads back exactly the bytes it copied, so no barrier is involved,
// and the gathers in flight no longer cost registers -- which is what capped the register-buffered versions (backward: one
// group, NB 1 beat NB 2 only through occupancy; forward: 64 registers, 8 blocks per SM).  0 = the register buffers above.
// Measured on config 2, ms for both layers (profiles/r2ak_edge_ring_ab.md):
//   fp32 backward  ring 0: 2.61 | 2: 2.05 | 4: 1.94 | 6: 2.12       bf16 backward  0: 1.90 | 2 (x2 rows): 1.56 | 4: 1.62 | 6: 2.29
//   fp32 forward   ring 0 / 8 blocks: 1.80 | 4 / 8: 1.73 | 4 / 10: 1.71 | 6 / 9: 1.74 | 3 / 12: 1.70
//   bf16 forward   ring 0 / 9 blocks: 1.68 | 2 / 9: 1.15 | 2 / 10: 1.16 | 3 / 9: 1.16 | 2 / 12: 1.10
#ifndef EDGE_BWD_RING
#define EDGE_BWD_RING 4
#endif
#ifndef EDGE_BWD_RING_BF16
#define EDGE_BWD_RING_BF16 2
#endif
#ifndef EDGE_FWD_RING
#define EDGE_FWD_RING 3
#endif
#ifndef EDGE_FWD_RING_BF16
#define EDGE_FWD_RING_BF16 2
#endif


__device__ __forceinline__ float half_sum(float v) {
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
  return v;
}
__device__ __forceinline__ float half_max(float v) {
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(kFull, v, o));
  return v;
}

template <typename T>
struct HalfRow {
  typename Raw4<T>::type a, b;
};

template <typename T, int POLICY, bool DROPOUT>
__global__ void __launch_bounds__(kEdgeThreads, sizeof(T) == 2 ? EDGE_MINB16_FWD_BF16 : EDGE_MINB16_FWD) edge_fwd16_kernel(const T* __restrict__ h, const float* __restrict__ s,
                                                                            const int4* __restrict__ sched,
                                                                            const int32_t* __restrict__ col,
                                                                            const int32_t* __restrict__ perm, int n_rows,
                                                                            int row_offset, float neg_slope,
                                                                            const float* __restrict__ bias, float* __restrict__ out,
                                                                            float* __restrict__ out_heads,
                                                                            float2* __restrict__ rowstat, float* __restrict__ partial,
                                                                            float p_drop, uint64_t seed, OutMode om) {
  constexpr int C = 128;
  constexpr int U = EDGE_U16_FWD * (sizeof(T) == 2 ? 2 : 1);   // rows per load group and half; EDGE_FWD_RING groups in flight (ring) or EDGE_NB16_FWD (registers)
  const int lane = threadIdx.x & 31, sl = lane & 15;
  const int idx = (blockIdx.x * (kEdgeThreads / 32) + (threadIdx.x >> 5)) * 2 + (lane >> 4);
  if ((idx & ~1) >= n_rows) return;     // both halves past the end
  const bool live = idx < n_rows;
  const float inv_keep = DROPOUT ? 1.f / (1.f - p_drop) : 1.f;
  const int4 d = live ? __ldg(sched + idx) : make_int4(0, 0, 0, 0);
  const int r = d.x, beg = d.y, end = d.z;
  const int ch0 = sl * 4, ch1 = 64 + sl * 4;

  auto load_group = [&](HalfRow<T>(&buf)[U], int c, int k, int cnt) {
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int ck = __shfl_sync(kFull, c, (k + u) & 15, 16);
      if (k + u < cnt) {
        const T* hp = h + (size_t)ck * C;
        buf[u].a = ld_raw4(hp + ch0);
        buf[u].b = ld_raw4(hp + ch1);
      }
    }
  };

  const float sd = live ? __ldg(s + (size_t)(row_offset + r) * 2 + 1) : 0.f;
  float m = POLICY == kPyG ? -INFINITY : 0.f, l = 0.f;
  float4 acc0 = make_float4(0.f, 0.f, 0.f, 0.f), acc1 = acc0;
  const int nch_mine = (end - beg + 15) >> 4;
  const int nch = max(nch_mine, __shfl_xor_sync(kFull, nch_mine, 16));
  int c_next = (beg + sl < end) ? __ldg(col + beg + sl) : 0;
  for (int ch = 0; ch < nch; ++ch) {
    const int base = beg + ch * 16;
    const int e = base + sl;
    const bool valid = e < end;
    const int c = c_next;
    if (ch + 1 < nch) c_next = (e + 16 < end) ? __ldg(col + e + 16) : 0;   // next chunk's ids: one load latency off the chain
    const float sv = valid ? __ldg(s + (size_t)c * 2) : 0.f;
    const int cnt = min(16, end - base);                                  // <= 0 once this half is done
    const int cmax = max(cnt, __shfl_xor_sync(kFull, cnt, 16));           // warp-uniform trip count
#if EDGE_FWD_RING > 0
    constexpr int NB = sizeof(T) == 2 ? EDGE_FWD_RING_BF16 : EDGE_FWD_RING;
    using Raw = typename Raw4<T>::type;
    __shared__ Raw ring[NB][U][2][kEdgeThreads];          // see edge_bwd16_kernel
    auto ring_load = [&](int b, int k, bool on) {
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int ck = __shfl_sync(kFull, c, (k + u) & 15, 16);
        if (on && k + u < cnt) {
          const T* hp = h + (size_t)ck * C;
          cp_async_row<sizeof(Raw)>((uint32_t)__cvta_generic_to_shared(&ring[b][u][0][threadIdx.x]), hp + ch0);
          cp_async_row<sizeof(Raw)>((uint32_t)__cvta_generic_to_shared(&ring[b][u][1][threadIdx.x]), hp + ch1);
        }
      }
      cp_async_commit();
    };
#pragma unroll
    for (int b = 0; b < NB; ++b) ring_load(b, b * U, b * U < cmax);
#else
    constexpr int NB = sizeof(T) == 2 ? EDGE_NB16_FWD_BF16 : EDGE_NB16_FWD;
    HalfRow<T> buf[NB][U];
    load_group(buf[0], c, 0, cnt);                                        // gathers start before the softmax math
#pragma unroll
    for (int b = 1; b < NB; ++b)
      if (b * U < cmax) load_group(buf[b], c, b * U, cnt);

#endif
    float z = -INFINITY, p;
    if (valid) z = activate<POLICY>(sv + sd, neg_slope);
    if (POLICY == kPyG) {
      const float nm = fmaxf(m, half_max(z));
      if (nm != m) {   // uniform inside the half
        const float scale = expf(m - nm);
        l *= scale;
        acc0.x *= scale; acc0.y *= scale; acc0.z *= scale; acc0.w *= scale;
        acc1.x *= scale; acc1.y *= scale; acc1.z *= scale; acc1.w *= scale;
        m = nm;
      }
      p = valid ? expf(z - nm) : 0.f;
    } else {
      p = valid ? expf(z) : 0.f;
    }
    l += p;   // the denominator sees every edge; dropout acts on alpha afterwards (:88-89)
    if (DROPOUT && valid) p *= dropout_scale(seed, (uint32_t)__ldg(perm + e), 0, p_drop, inv_keep);

#if EDGE_FWD_RING > 0
    for (int k = 0; k < cmax; k += NB * U) {
#pragma unroll
      for (int b = 0; b < NB; ++b) {
        if (k + b * U < cmax) {                                           // warp-uniform
          cp_async_wait<NB - 1>();
#pragma unroll
          for (int u = 0; u < U; ++u) {
            const float pk = __shfl_sync(kFull, p, (k + b * U + u) & 15, 16);
            if (k + b * U + u < cnt) {
              acc0 = fma4(pk, to_f4(ring[b][u][0][threadIdx.x]), acc0);
              acc1 = fma4(pk, to_f4(ring[b][u][1][threadIdx.x]), acc1);
            }
          }
          ring_load(b, k + (b + NB) * U, k + (b + NB) * U < cmax);
        }
      }
    }
    cp_async_wait<0>();
#else
    auto consume = [&](HalfRow<T>(&buf)[U], int k) {
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const float pk = __shfl_sync(kFull, p, (k + u) & 15, 16);
        if (k + u < cnt) {
          acc0 = fma4(pk, to_f4(buf[u].a), acc0);
          acc1 = fma4(pk, to_f4(buf[u].b), acc1);
        }
      }
    };
    for (int k = 0; k < cmax; k += NB * U) {
      consume(buf[0], k);
      if (k + NB * U < cmax) load_group(buf[0], c, k + NB * U, cnt);
#pragma unroll
      for (int b = 1; b < NB; ++b) {
        if (k + b * U < cmax) {                                           // warp-uniform
          consume(buf[b], k + b * U);
          if (k + (b + NB) * U < cmax) load_group(buf[b], c, k + (b + NB) * U, cnt);
        }
      }
    }
#endif
  }
  const float lt = half_sum(l);
  if (!live) return;
  if (d.w > 0) {  // segment of a split (long) row: park the un-normalised state, fwd_combine_kernel finishes the row
    float* ps = partial + (size_t)(d.w - 1) * (C + 4);
    if (sl == 0) { ps[0] = m; ps[1] = lt; }
    *reinterpret_cast<float4*>(ps + 4 + ch0) = acc0;
    *reinterpret_cast<float4*>(ps + 4 + ch1) = acc1;
    return;
  }
  const float inv = 1.f / (lt + (POLICY == kCustom ? 1e-9f : 1e-16f));
  if (sl == 0 && rowstat) rowstat[r] = make_float2(beg < end ? m : 0.f, inv);
  acc0.x *= inv; acc0.y *= inv; acc0.z *= inv; acc0.w *= inv;
  acc1.x *= inv; acc1.y *= inv; acc1.z *= inv; acc1.w *= inv;
  if (out_heads) {
    st_stream4(out_heads + (size_t)r * C + ch0, acc0);
    st_stream4(out_heads + (size_t)r * C + ch1, acc1);
  }
  if (om.scale != 1.f) {
    acc0.x *= om.scale; acc0.y *= om.scale; acc0.z *= om.scale; acc0.w *= om.scale;
    acc1.x *= om.scale; acc1.y *= om.scale; acc1.z *= om.scale; acc1.w *= om.scale;
  }
  if (bias) {
    const float4 b0 = ldg4(bias + ch0), b1 = ldg4(bias + ch1);
    acc0.x += b0.x; acc0.y += b0.y; acc0.z += b0.z; acc0.w += b0.w;
    acc1.x += b1.x; acc1.y += b1.y; acc1.z += b1.z; acc1.w += b1.w;
  }
  float4* o0 = reinterpret_cast<float4*>(out + (size_t)r * C + ch0);
  float4* o1 = reinterpret_cast<float4*>(out + (size_t)r * C + ch1);
  if (om.accumulate) {
    const float4 p0 = *o0, p1 = *o1;
    acc0.x += p0.x; acc0.y += p0.y; acc0.z += p0.z; acc0.w += p0.w;
    acc1.x += p1.x; acc1.y += p1.y; acc1.z += p1.z; acc1.w += p1.w;
  }
  *o0 = acc0;
  *o1 = acc1;
}

template <typename T, int POLICY, bool DROPOUT>
__global__ void __launch_bounds__(kEdgeThreads, EDGE_MINB16_BWD) edge_bwd16_kernel(const T* __restrict__ h, const float* __restrict__ s,
                                                                            const T* __restrict__ dout,
                                                                            const float4* __restrict__ nodestat,
                                                                            const int4* __restrict__ sched,
                                                                            const int32_t* __restrict__ row,
                                                                            const int32_t* __restrict__ perm_csc, int n_rows,
                                                                            int row_offset, float neg_slope, float* __restrict__ dh,
                                                                            float* __restrict__ de, float* __restrict__ ds_src,
                                                                            int ld_ds, float* __restrict__ partial, float p_drop,
                                                                            uint64_t seed) {
  constexpr int C = 128;
  constexpr int U = EDGE_U16_BWD * (sizeof(T) == 2 ? 2 : 1);
  const int lane = threadIdx.x & 31, sl = lane & 15;
  const int idx = (blockIdx.x * (kEdgeThreads / 32) + (threadIdx.x >> 5)) * 2 + (lane >> 4);
  if ((idx & ~1) >= n_rows) return;
  const bool live = idx < n_rows;
  const float inv_keep = DROPOUT ? 1.f / (1.f - p_drop) : 1.f;
  const int4 d = live ? __ldg(sched + idx) : make_int4(0, 0, 0, 0);
  const int r = d.x, beg = d.y, end = d.z;
  const int ch0 = sl * 4, ch1 = 64 + sl * 4;
  const size_t j = (size_t)row_offset + r;

  auto load_group = [&](HalfRow<T>(&buf)[U], int i, int k, int cnt) {
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int ik = __shfl_sync(kFull, i, (k + u) & 15, 16);
      if (k + u < cnt) {
        const T* gp = dout + (size_t)ik * C;
        buf[u].a = ld_raw4(gp + ch0);
        buf[u].b = ld_raw4(gp + ch1);
      }
    }
  };

  float4 hj0 = make_float4(0.f, 0.f, 0.f, 0.f), hj1 = hj0, acc0 = hj0, acc1 = hj0;
  float ssj = 0.f, dss = 0.f;
  if (beg < end) {   // false for the dead half too (beg = end = 0)
    ssj = __ldg(s + j * 2);
    hj0 = to_f4(ld_once4(h + j * C + ch0));
    hj1 = to_f4(ld_once4(h + j * C + ch1));
  }
  const int nch_mine = (end - beg + 15) >> 4;
  const int nch = max(nch_mine, __shfl_xor_sync(kFull, nch_mine, 16));
  int i_next = (beg + sl < end) ? ld_once(row + beg + sl) : 0;
  for (int ch = 0; ch < nch; ++ch) {
    const int base = beg + ch * 16;
    const int q = base + sl;
    const bool valid = q < end;
    const int i = i_next;
    if (ch + 1 < nch) i_next = (q + 16 < end) ? ld_once(row + q + 16) : 0;   // next chunk's ids: one load latency off the chain
    const int cnt = min(16, end - base);
    const int cmax = max(cnt, __shfl_xor_sync(kFull, cnt, 16));
#if EDGE_BWD_RING > 0
    constexpr int NB = sizeof(T) == 2 ? EDGE_BWD_RING_BF16 : EDGE_BWD_RING;
    using Raw = typename Raw4<T>::type;
    __shared__ Raw ring[NB][U][2][kEdgeThreads];          // [group slot][row of the group][piece][thread]: conflict-free, private per thread
    // group b <- rows k .. k + U - 1 of this chunk (or nothing); ALWAYS one commit, so that "all but the newest NB - 1 groups have
    // landed" means the same thing at every point of the chunk
    auto ring_load = [&](int b, int k, bool on) {
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int ik = __shfl_sync(kFull, i, (k + u) & 15, 16);
        if (on && k + u < cnt) {
          const T* gp = dout + (size_t)ik * C;
          cp_async_row<sizeof(Raw)>((uint32_t)__cvta_generic_to_shared(&ring[b][u][0][threadIdx.x]), gp + ch0);
          cp_async_row<sizeof(Raw)>((uint32_t)__cvta_generic_to_shared(&ring[b][u][1][threadIdx.x]), gp + ch1);
        }
      }
      cp_async_commit();
    };
#pragma unroll
    for (int b = 0; b < NB; ++b) ring_load(b, b * U, b * U < cmax);
#else
    constexpr int NB = EDGE_NB16_BWD;
    HalfRow<T> buf[NB][U];
    load_group(buf[0], i, 0, cnt);                    // dout gathers start before the per-edge scalar math
#pragma unroll
    for (int b = 1; b < NB; ++b)
      if (b * U < cmax) load_group(buf[b], i, b * U, cnt);

#endif
    // per lane (= per edge of this chunk): agg = alpha' (weight of dout_i in dh_j), and the two coefficients of
    // de = alpha * (dalpha * keep - t) * slope  written as  de = gA * <dout_i, h_j> - cB
    float agg = 0.f, gA = 0.f, cB = 0.f, my_de = 0.f;
    if (valid) {
      const float4 st = __ldg(nodestat + i);  // (s_dst, m, 1/D, t)
      const float z0 = ssj + st.x;
      const float zl = z0 > 0.f ? z0 : z0 * neg_slope;
      float zc = zl, pass = 1.f;
      if (POLICY == kCustom) {
        zc = fminf(fmaxf(zl, -10.f), 10.f);
        pass = (zl >= -10.f && zl <= 10.f) ? 1.f : 0.f;  // clamp passes gradient only inside [-10,10]
      }
      const float alpha = expf(zc - st.y) * st.z;
      const float gsc = (z0 > 0.f ? 1.f : neg_slope) * pass;
      float ks = 1.f;
      if (DROPOUT) ks = dropout_scale(seed, (uint32_t)ld_once(perm_csc + q), 0, p_drop, inv_keep);
      agg = alpha * ks;
      gA = agg * gsc;
      cB = alpha * st.w * gsc;
    }
#if EDGE_BWD_RING > 0
    auto consume_ring = [&](int b, int k) {
      float dsum[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const float a = __shfl_sync(kFull, agg, (k + u) & 15, 16);
        dsum[u] = 0.f;
        if (k + u < cnt) {
          const float4 g0 = to_f4(ring[b][u][0][threadIdx.x]), g1 = to_f4(ring[b][u][1][threadIdx.x]);
          acc0 = fma4(a, g0, acc0);
          acc1 = fma4(a, g1, acc1);
          dsum[u] = dot4(hj0, g0) + dot4(hj1, g1);
        }
      }
      const int mine = (sl - k) & 15;                    // which edge of the group this lane owns (if < U)
      const bool owner = mine < U && k + mine < cnt;
      const float tot = MultiReduce<U, 8>::run(dsum, lane);
      const float dot = __shfl_sync(kFull, tot, (mine & (U - 1)) << (4 - Log2<U>::value), 16);
      if (owner) my_de = gA * dot - cB;
    };
    for (int k = 0; k < cmax; k += NB * U) {
#pragma unroll
      for (int b = 0; b < NB; ++b) {
        if (k + b * U < cmax) {                         // warp-uniform
          cp_async_wait<NB - 1>();
          consume_ring(b, k + b * U);
          ring_load(b, k + (b + NB) * U, k + (b + NB) * U < cmax);
        }
      }
    }
    cp_async_wait<0>();                                 // nothing of this chunk is left in flight when the next one reuses the slots
#else
    auto consume = [&](HalfRow<T>(&buf)[U], int k) {
      float dsum[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const float a = __shfl_sync(kFull, agg, (k + u) & 15, 16);
        dsum[u] = 0.f;
        if (k + u < cnt) {
          const float4 g0 = to_f4(buf[u].a), g1 = to_f4(buf[u].b);
          acc0 = fma4(a, g0, acc0);
          acc1 = fma4(a, g1, acc1);
          dsum[u] = dot4(hj0, g0) + dot4(hj1, g1);
        }
      }
      const int mine = (sl - k) & 15;                    // which edge of the group this lane owns (if < U)
      const bool owner = mine < U && k + mine < cnt;
      const float tot = MultiReduce<U, 8>::run(dsum, lane);               // total of edge u sits in lanes sl >> (4 - log2 U) == u
      const float dot = __shfl_sync(kFull, tot, (mine & (U - 1)) << (4 - Log2<U>::value), 16);
      if (owner) my_de = gA * dot - cB;
    };
    for (int k = 0; k < cmax; k += NB * U) {
      consume(buf[0], k);
      if (k + NB * U < cmax) load_group(buf[0], i, k + NB * U, cnt);
#pragma unroll
      for (int b = 1; b < NB; ++b) {
        if (k + b * U < cmax) {                       // warp-uniform
          consume(buf[b], k + b * U);
          if (k + (b + NB) * U < cmax) load_group(buf[b], i, k + (b + NB) * U, cnt);
        }
      }
    }
#endif
    if (valid) {
      de[q] = my_de;          // read again by ds_dst right after this kernel: left to the default policy
      dss += my_de;
    }
  }
  const float t = half_sum(dss);
  if (!live) return;
  if (d.w > 0) {  // segment of a split (long) row: bwd_combine_kernel sums the segments in order
    float* ps = partial + (size_t)(d.w - 1) * (C + 4);
    if (sl == 0) ps[0] = t;
    *reinterpret_cast<float4*>(ps + 4 + ch0) = acc0;
    *reinterpret_cast<float4*>(ps + 4 + ch1) = acc1;
    return;
  }
  if (sl == 0) ds_src[(size_t)r * ld_ds] = t;
  st_once4(dh + (size_t)r * C + ch0, acc0);
  st_once4(dh + (size_t)r * C + ch1, acc1);
}

// ---- dispatch helpers ---------------------------------------------------------------------------
#define B200GAT_DISPATCH_HC(H_, CV_, ...)                                        \
  if (H_ == 1 && CV_ == 1) { constexpr int kH = 1, kCV = 1; __VA_ARGS__; }        \
  else if (H_ == 2 && CV_ == 1) { constexpr int kH = 2, kCV = 1; __VA_ARGS__; }   \
  else if (H_ == 4 && CV_ == 1) { constexpr int kH = 4, kCV = 1; __VA_ARGS__; }   \
  else if (H_ == 1 && CV_ == 2) { constexpr int kH = 1, kCV = 2; __VA_ARGS__; }   \
  else if (H_ == 4 && CV_ == 2) { constexpr int kH = 4, kCV = 2; __VA_ARGS__; }   \
  else { set_error("unsupported heads=%d channels=%d (heads in {1,2,4} x C=128, or heads in {1,4} x C=256)", H_, CV_ * 128); return kErrUnsupported; }

static int check_shape(int heads, int channels) {
  if (channels % 128 != 0 || channels <= 0) {
    set_error("out_channels=%d must be a positive multiple of 128 (128-bit lane gathers)", channels);
    return kErrUnsupported;
  }
  (void)heads;
  return kOk;
}

}  // namespace b200gat

using namespace b200gat;

// one warp per scheduled row; the hardware block scheduler hands out blocks in schedule order (longest rows first)
template <typename K>
static int persistent_grid(K, int threads, int64_t n_rows, int* grid) {
  *grid = (int)((n_rows + threads / 32 - 1) / (threads / 32));
  return kOk;
}

extern "C" int b200gat_schedule_workspace_bytes(int64_t n_rows, size_t* bytes) {
  B200GAT_CHECK_ARG(bytes && n_rows >= 0, "bad args");
  const size_t e = ((size_t)(n_rows > 0 ? n_rows : 1) * 4 + 255) / 256 * 256;
  *bytes = 3 * e + sort_workspace_bytes(n_rows) + 256;
  return kOk;
}

// sched[k] = (row, ptr[row], ptr[row+1], 0) for the k-th row in descending-degree order (ties: descending row id).
// degree_bound: any value > the largest row degree (e.g. number of edges + 1); it only sets the radix pass count.
extern "C" int b200gat_build_schedule(const int32_t* ptr, int64_t n_rows, int64_t degree_bound, int32_t* sched,
                                      void* workspace, size_t workspace_bytes, void* stream) {
  B200GAT_CHECK_ARG(ptr && (sched || n_rows == 0) && workspace, "bad arguments");
  size_t need;
  b200gat_schedule_workspace_bytes(n_rows, &need);
  B200GAT_CHECK_ARG(workspace_bytes >= need, "workspace too small: %zu < %zu", workspace_bytes, need);
  if (n_rows == 0) return kOk;
  cudaStream_t st = (cudaStream_t)stream;
  char* p = (char*)workspace;
  const size_t e = ((size_t)n_rows * 4 + 255) / 256 * 256;
  int32_t* keys = (int32_t*)p;
  int32_t* sorted = (int32_t*)(p + e);
  int32_t* asc_rows = (int32_t*)(p + 2 * e);
  void* sort_ws = p + 3 * e;
  const int grid = min(ceil_div(n_rows, 256), kNumSMs * 8);
  if (degree_bound <= 0) {  // natural row order
    count_launch(), build_schedule_kernel<<<grid, 256, 0, st>>>(ptr, nullptr, n_rows, (int4*)sched);
    B200GAT_LAUNCH_CHECK();
    return kOk;
  }
  count_launch(), degree_keys_kernel<<<grid, 256, 0, st>>>(ptr, n_rows, keys);
  int rc = sort_pairs_stable(keys, n_rows, degree_bound, sorted, asc_rows, sort_ws, st);
  if (rc) return rc;
  count_launch(), build_schedule_kernel<<<grid, 256, 0, st>>>(ptr, asc_rows, n_rows, (int4*)sched);
  B200GAT_LAUNCH_CHECK();
  return kOk;
}

template <typename T>
static int edge_fwd_impl(const T* h, const float* s, const int32_t* sched, int64_t n_rows, const int32_t* long_table,
                         int64_t n_long, float* partial, const int32_t* col, const int32_t* perm, int64_t row_offset, int heads,
                         int channels, int policy, float negative_slope, const float* bias, float* out, float* out_heads,
                         float* rowstat, float p_drop, uint64_t seed, void* stream, OutMode om = OutMode{1.f, 0}) {
  B200GAT_CHECK_ARG(h && s && sched && out, "null pointer");
  B200GAT_CHECK_ARG(n_long == 0 || (long_table && partial), "split rows need long_table and partial");
  B200GAT_CHECK_ARG(policy == kCustom || policy == kPyG, "bad policy %d", policy);
  B200GAT_CHECK_ARG(p_drop >= 0.f && p_drop < 1.f, "dropout p=%f outside [0,1)", p_drop);
  B200GAT_CHECK_ARG(p_drop == 0.f || perm, "dropout needs perm");
  int rc = check_shape(heads, channels);
  if (rc) return rc;
  if (n_rows == 0) return kOk;
  const int cv = channels / 128;
  cudaStream_t st = (cudaStream_t)stream;
  const bool drop = p_drop > 0.f;
  int grid = 0;
#define LAUNCH_FWD(P, D)                                                                                                \
  do {                                                                                                                  \
    if (kH * kCV == 1 && EDGE_W == 16) {                                                                                \
      grid = (int)ceil_div(n_rows, 2 * (kEdgeThreads / 32));                                                            \
      count_launch(), edge_fwd16_kernel<T, P, D><<<grid, kEdgeThreads, 0, st>>>(                                        \
          h, s, (const int4*)sched, col, perm, (int)n_rows, (int)row_offset, negative_slope, bias, out, out_heads,      \
          (float2*)rowstat, partial, p_drop, seed, om);                                                                 \
    } else {                                                                                                            \
      rc = persistent_grid(edge_fwd_kernel<T, P, kH, kCV, D>, kEdgeThreads, n_rows, &grid);                             \
      if (rc) return rc;                                                                                                \
      count_launch(), edge_fwd_kernel<T, P, kH, kCV, D><<<grid, kEdgeThreads, 0, st>>>(                                 \
          h, s, (const int4*)sched, col, perm, (int)n_rows, (int)row_offset, negative_slope, bias, out, out_heads,      \
          (float2*)rowstat, partial, p_drop, seed, om);                                                                 \
    }                                                                                                                   \
    if (n_long > 0)                                                                                                     \
      count_launch(), fwd_combine_kernel<P, kH, kCV><<<ceil_div(n_long * 32, kEdgeThreads), kEdgeThreads, 0, st>>>(     \
          partial, (const int4*)long_table, (int)n_long, bias, out, out_heads, (float2*)rowstat, om);                   \
  } while (0)
  B200GAT_DISPATCH_HC(heads, cv, {
    if (policy == kCustom) { if (drop) LAUNCH_FWD(kCustom, true); else LAUNCH_FWD(kCustom, false); }
    else { if (drop) LAUNCH_FWD(kPyG, true); else LAUNCH_FWD(kPyG, false); }
  })
#undef LAUNCH_FWD
  B200GAT_LAUNCH_CHECK();
  return kOk;
}

extern "C" int b200gat_edge_fwd_f32(const float* h, const float* s, const int32_t* sched, int64_t n_rows,
                                    const int32_t* long_table, int64_t n_long, float* partial, const int32_t* col,
                                    const int32_t* perm, int64_t row_offset, int heads, int channels,
                                    int policy, float negative_slope, const float* bias, float* out, float* out_heads,
                                    float* rowstat, float p_drop, uint64_t seed, void* stream) {
  return edge_fwd_impl<float>(h, s, sched, n_rows, long_table, n_long, partial, col, perm, row_offset, heads, channels, policy,
                              negative_slope, bias, out, out_heads, rowstat, p_drop, seed, stream);
}
// same with the gathered matrix h stored as bf16 (half the gather bytes; fp32 accumulation and outputs)
extern "C" int b200gat_edge_fwd_bf16(const void* h_bf16, const float* s, const int32_t* sched, int64_t n_rows,
                                     const int32_t* long_table, int64_t n_long, float* partial, const int32_t* col,
                                     const int32_t* perm, int64_t row_offset, int heads, int channels,
                                     int policy, float negative_slope, const float* bias, float* out, float* out_heads,
                                     float* rowstat, float p_drop, uint64_t seed, void* stream) {
  return edge_fwd_impl<__nv_bfloat16>((const __nv_bfloat16*)h_bf16, s, sched, n_rows, long_table, n_long, partial, col, perm,
                                      row_offset, heads, channels, policy, negative_slope, bias, out, out_heads, rowstat, p_drop,
                                      seed, stream);
}

// dbias (nullable): column sums of dout in a fixed order; workspace >= 8*148*channels floats when dbias != NULL
extern "C" int b200gat_node_prep_f32(const float* dout, const float* out_heads, const float* bias, const float* s,
                                     const float* rowstat, int64_t n_rows, int64_t row_offset, int heads, int channels,
                                     float* nodestat, float* dbias, void* dout_bf16, void* workspace, size_t workspace_bytes,
                                     void* stream) {
  B200GAT_CHECK_ARG(dout && out_heads && s && rowstat && nodestat, "null pointer");
  int rc = check_shape(heads, channels);
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  if (n_rows == 0) {
    if (dbias) B200GAT_CUDA(cudaMemsetAsync(dbias, 0, sizeof(float) * channels, st));
    return kOk;
  }
  const int cv = channels / 128;
  const int64_t want_blocks = (n_rows + 7) / 8;
  const int grid = (int)(want_blocks < (int64_t)kNumSMs * 8 ? want_blocks : (int64_t)kNumSMs * 8);
  B200GAT_CHECK_ARG(!dbias || (workspace && workspace_bytes >= (size_t)grid * channels * sizeof(float)),
                    "node_prep workspace too small");
  B200GAT_DISPATCH_HC(heads, cv, {
    count_launch(), node_prep_kernel<kH, kCV><<<grid, 256, 0, st>>>(dout, out_heads, bias, s, (const float2*)rowstat, (int)n_rows,
                                                                   (int)row_offset, (float4*)nodestat,
                                                                   dbias ? (float*)workspace : nullptr, (__nv_bfloat16*)dout_bf16);
  })
  if (dbias) count_launch(), colsum_finish_kernel<<<ceil_div(channels, 32), 256, 0, st>>>((const float*)workspace, grid, channels, dbias);
  B200GAT_LAUNCH_CHECK();
  return kOk;
}

template <typename T>
static int edge_bwd_impl(const T* h, const float* s, const T* dout, const float* nodestat, const int32_t* sched, int64_t n_rows,
                         const int32_t* long_table, int64_t n_long, float* partial, const int32_t* row, const int32_t* perm_csc,
                         int64_t row_offset, int heads, int channels, int policy, float negative_slope, float* dh, float* de,
                         float* ds_src, int ld_ds, float p_drop, uint64_t seed, void* stream, bool two_phase = false) {
  B200GAT_CHECK_ARG(h && s && dout && nodestat && sched && dh && ds_src && ld_ds >= heads, "null pointer / bad ld");
  B200GAT_CHECK_ARG(n_long == 0 || (long_table && partial), "split rows need long_table and partial");
  B200GAT_CHECK_ARG(policy == kCustom || policy == kPyG, "bad policy %d", policy);
  B200GAT_CHECK_ARG(p_drop >= 0.f && p_drop < 1.f, "dropout p=%f outside [0,1)", p_drop);
  B200GAT_CHECK_ARG(p_drop == 0.f || perm_csc, "dropout needs perm_csc");
  int rc = check_shape(heads, channels);
  if (rc) return rc;
  if (n_rows == 0) return kOk;
  const int cv = channels / 128;
  cudaStream_t st = (cudaStream_t)stream;
  const bool drop = p_drop > 0.f;
  int grid = 0;
#define LAUNCH_BWD(P, D)                                                                                               \
  do {                                                                                                                 \
    if (two_phase) {                                                                                                   \
      if constexpr (kH == 1) {   /* the two-phase form exists for per-head streaming: one head per launch */            \
        rc = persistent_grid(edge_bwd_kernel<T, P, kH, kCV, D, true>, kEdgeThreads, n_rows, &grid);                    \
        if (rc) return rc;                                                                                             \
        count_launch(), edge_bwd_kernel<T, P, kH, kCV, D, true><<<grid, kEdgeThreads, 0, st>>>(                        \
            h, s, dout, (const float4*)nodestat, (const int4*)sched, row, perm_csc, (int)n_rows, (int)row_offset,      \
            negative_slope, dh, de, ds_src, ld_ds, partial, p_drop, seed);                                             \
      } else {                                                                                                         \
        set_error("the two-phase backward runs one head per launch (heads=%d)", heads);                                \
        return kErrUnsupported;                                                                                        \
      }                                                                                                                \
    } else if (kH * kCV == 1 && EDGE_W == 16) {                                                                        \
      grid = (int)ceil_div(n_rows, 2 * (kEdgeThreads / 32));                                                           \
      count_launch(), edge_bwd16_kernel<T, P, D><<<grid, kEdgeThreads, 0, st>>>(                                       \
          h, s, dout, (const float4*)nodestat, (const int4*)sched, row, perm_csc, (int)n_rows, (int)row_offset,        \
          negative_slope, dh, de, ds_src, ld_ds, partial, p_drop, seed);                                               \
    } else {                                                                                                           \
      rc = persistent_grid(edge_bwd_kernel<T, P, kH, kCV, D>, kEdgeThreads, n_rows, &grid);                            \
      if (rc) return rc;                                                                                               \
      count_launch(), edge_bwd_kernel<T, P, kH, kCV, D><<<grid, kEdgeThreads, 0, st>>>(                                \
          h, s, dout, (const float4*)nodestat, (const int4*)sched, row, perm_csc, (int)n_rows, (int)row_offset,        \
          negative_slope, dh, de, ds_src, ld_ds, partial, p_drop, seed);                                               \
    }                                                                                                                  \
    if (n_long > 0)                                                                                                    \
      count_launch(), bwd_combine_kernel<kH, kCV><<<ceil_div(n_long * 32, kEdgeThreads), kEdgeThreads, 0, st>>>(       \
          partial, (const int4*)long_table, (int)n_long, dh, ds_src, ld_ds);                                           \
  } while (0)
  B200GAT_DISPATCH_HC(heads, cv, {
    if (policy == kCustom) { if (drop) LAUNCH_BWD(kCustom, true); else LAUNCH_BWD(kCustom, false); }
    else { if (drop) LAUNCH_BWD(kPyG, true); else LAUNCH_BWD(kPyG, false); }
  })
#undef LAUNCH_BWD
  B200GAT_LAUNCH_CHECK();
  return kOk;
}

extern "C" int b200gat_edge_bwd_f32(const float* h, const float* s, const float* dout, const float* nodestat,
                                    const int32_t* sched, int64_t n_rows, const int32_t* long_table, int64_t n_long,
                                    float* partial, const int32_t* row, const int32_t* perm_csc,
                                    int64_t row_offset, int heads, int channels, int policy, float negative_slope,
                                    float* dh, float* de, float* ds_src, int ld_ds, float p_drop, uint64_t seed,
                                    void* stream) {
  return edge_bwd_impl<float>(h, s, dout, nodestat, sched, n_rows, long_table, n_long, partial, row, perm_csc, row_offset, heads,
                              channels, policy, negative_slope, dh, de, ds_src, ld_ds, p_drop, seed, stream);
}
// same with h and the gathered dout stored as bf16 (dout_bf16 is written by b200gat_node_prep_f32)
extern "C" int b200gat_edge_bwd_bf16(const void* h_bf16, const float* s, const void* dout_bf16, const float* nodestat,
                                     const int32_t* sched, int64_t n_rows, const int32_t* long_table, int64_t n_long,
                                     float* partial, const int32_t* row, const int32_t* perm_csc,
                                     int64_t row_offset, int heads, int channels, int policy, float negative_slope,
                                     float* dh, float* de, float* ds_src, int ld_ds, float p_drop, uint64_t seed,
                                     void* stream) {
  return edge_bwd_impl<__nv_bfloat16>((const __nv_bfloat16*)h_bf16, s, (const __nv_bfloat16*)dout_bf16, nodestat, sched, n_rows,
                                      long_table, n_long, partial, row, perm_csc, row_offset, heads, channels, policy,
                                      negative_slope, dh, de, ds_src, ld_ds, p_drop, seed, stream);
}

extern "C" int b200gat_ds_dst_f32(const float* de, const int32_t* rowptr, const int32_t* csr2csc, int64_t n_rows,
                                  int64_t n_edges, int heads, float* ds_dst, int ld_ds, void* stream) {
  B200GAT_CHECK_ARG(rowptr && ds_dst && ld_ds >= heads, "null pointer / bad ld");
  if (n_rows == 0) return kOk;
  if (n_edges < 8 * n_rows) {   // short rows: one thread per row (the per-row sums stay in edge order: same result)
    cudaStream_t st2 = (cudaStream_t)stream;
    const int g2 = ceil_div(n_rows, 256);
    if (heads == 1) count_launch(), ds_dst_thread_kernel<1><<<g2, 256, 0, st2>>>(de, rowptr, csr2csc, (int)n_rows, ds_dst, ld_ds);
    else if (heads == 2) count_launch(), ds_dst_thread_kernel<2><<<g2, 256, 0, st2>>>(de, rowptr, csr2csc, (int)n_rows, ds_dst, ld_ds);
    else if (heads == 4) count_launch(), ds_dst_thread_kernel<4><<<g2, 256, 0, st2>>>(de, rowptr, csr2csc, (int)n_rows, ds_dst, ld_ds);
    else { set_error("unsupported heads=%d", heads); return kErrUnsupported; }
    B200GAT_LAUNCH_CHECK();
    return kOk;
  }
  constexpr int kW = 8;
  const int grid = ceil_div(n_rows * kW, 128);
  cudaStream_t st = (cudaStream_t)stream;
  if (heads == 1) count_launch(), ds_dst_kernel<1, kW><<<grid, 128, 0, st>>>(de, rowptr, csr2csc, (int)n_rows, ds_dst, ld_ds);
  else if (heads == 2) count_launch(), ds_dst_kernel<2, kW><<<grid, 128, 0, st>>>(de, rowptr, csr2csc, (int)n_rows, ds_dst, ld_ds);
  else if (heads == 4) count_launch(), ds_dst_kernel<4, kW><<<grid, 128, 0, st>>>(de, rowptr, csr2csc, (int)n_rows, ds_dst, ld_ds);
  else { set_error("unsupported heads=%d", heads); return kErrUnsupported; }
  B200GAT_LAUNCH_CHECK();
  return kOk;
}

// ---- per-head streaming (BASELINE config 5: heads x channels = 4 x 256 over 30 M nodes does not fit as one [N, H*C] matrix) ----
// One launch per head with heads = 1 arguments (h_head [N, C], s_head [N, 2]); `out` collects the head mean:
// out = (accumulate ? out : 0) + out_scale * result (+ bias).
extern "C" int b200gat_edge_fwd_stream_f32(const float* h, const float* s, const int32_t* sched, int64_t n_rows,
                                           const int32_t* long_table, int64_t n_long, float* partial, const int32_t* col,
                                           const int32_t* perm, int64_t row_offset, int channels, int policy,
                                           float negative_slope, const float* bias, float* out, float* rowstat, float p_drop,
                                           uint64_t seed, float out_scale, int accumulate, void* stream) {
  return edge_fwd_impl<float>(h, s, sched, n_rows, long_table, n_long, partial, col, perm, row_offset, 1, channels, policy,
                              negative_slope, bias, out, nullptr, rowstat, p_drop, seed, stream, OutMode{out_scale, accumulate});
}
extern "C" int b200gat_edge_fwd_stream_bf16(const void* h_bf16, const float* s, const int32_t* sched, int64_t n_rows,
                                            const int32_t* long_table, int64_t n_long, float* partial, const int32_t* col,
                                            const int32_t* perm, int64_t row_offset, int channels, int policy,
                                            float negative_slope, const float* bias, float* out, float* rowstat, float p_drop,
                                            uint64_t seed, float out_scale, int accumulate, void* stream) {
  return edge_fwd_impl<__nv_bfloat16>((const __nv_bfloat16*)h_bf16, s, sched, n_rows, long_table, n_long, partial, col, perm,
                                      row_offset, 1, channels, policy, negative_slope, bias, out, nullptr, rowstat, p_drop, seed,
                                      stream, OutMode{out_scale, accumulate});
}

// Two-phase backward of ONE head without its saved output (see edge_bwd_kernel's TP note).  phase 1: same arguments as
// b200gat_edge_bwd_* with heads = 1; nodestat[i] = (s_dst, m, 1/D, ignored); writes dh and w = alpha * dalpha into `de`.
// The caller sums w per destination (b200gat_ds_dst_f32 on `de`), stores the sums into nodestat[.].w (set_t), and runs phase 2,
// which rewrites `de` in place with the logit gradients and produces ds_src; ds_dst is then b200gat_ds_dst_f32 on `de` again.
extern "C" int b200gat_edge_bwd_phase1_f32(const float* h, const float* s, const float* dout, const float* nodestat,
                                           const int32_t* sched, int64_t n_rows, const int32_t* long_table, int64_t n_long,
                                           float* partial, const int32_t* row, const int32_t* perm_csc, int64_t row_offset,
                                           int channels, int policy, float negative_slope, float* dh, float* de, float* ds_src,
                                           int ld_ds, float p_drop, uint64_t seed, void* stream) {
  return edge_bwd_impl<float>(h, s, dout, nodestat, sched, n_rows, long_table, n_long, partial, row, perm_csc, row_offset, 1,
                              channels, policy, negative_slope, dh, de, ds_src, ld_ds, p_drop, seed, stream, true);
}
extern "C" int b200gat_edge_bwd_phase1_bf16(const void* h_bf16, const float* s, const void* dout_bf16, const float* nodestat,
                                            const int32_t* sched, int64_t n_rows, const int32_t* long_table, int64_t n_long,
                                            float* partial, const int32_t* row, const int32_t* perm_csc, int64_t row_offset,
                                            int channels, int policy, float negative_slope, float* dh, float* de, float* ds_src,
                                            int ld_ds, float p_drop, uint64_t seed, void* stream) {
  return edge_bwd_impl<__nv_bfloat16>((const __nv_bfloat16*)h_bf16, s, (const __nv_bfloat16*)dout_bf16, nodestat, sched, n_rows,
                                      long_table, n_long, partial, row, perm_csc, row_offset, 1, channels, policy, negative_slope,
                                      dh, de, ds_src, ld_ds, p_drop, seed, stream, true);
}
extern "C" int b200gat_edge_bwd_phase2_f32(float* de, const int32_t* colptr, const int32_t* row, const float* s,
                                           const float* nodestat, int64_t n_rows, int64_t row_offset, int policy,
                                           float negative_slope, float* ds_src, int ld_ds, void* stream) {
  B200GAT_CHECK_ARG(de && colptr && row && s && nodestat && ds_src && ld_ds >= 1, "null pointer / bad ld");
  B200GAT_CHECK_ARG(policy == kCustom || policy == kPyG, "bad policy %d", policy);
  if (n_rows == 0) return kOk;
  constexpr int kW = 8;
  const int grid = ceil_div(n_rows * kW, 128);
  cudaStream_t st = (cudaStream_t)stream;
  if (policy == kCustom)
    count_launch(), edge_de_kernel<kCustom, 1, kW><<<grid, 128, 0, st>>>(de, colptr + row_offset, row, s, (const float4*)nodestat, (int)n_rows,
                                                                         (int)row_offset, negative_slope, ds_src, ld_ds);
  else
    count_launch(), edge_de_kernel<kPyG, 1, kW><<<grid, 128, 0, st>>>(de, colptr + row_offset, row, s, (const float4*)nodestat, (int)n_rows,
                                                                      (int)row_offset, negative_slope, ds_src, ld_ds);
  B200GAT_LAUNCH_CHECK();
  return kOk;
}
// nodestat[r, h] = (s_dst, m, 1/D, 0) for rows [0, n_rows) of a block whose first row is row_offset in s
extern "C" int b200gat_node_stat_f32(const float* s, const float* rowstat, int64_t n_rows, int64_t row_offset, int heads,
                                     float* nodestat, void* stream) {
  B200GAT_CHECK_ARG(s && rowstat && nodestat && heads >= 1, "null pointer");
  if (n_rows == 0) return kOk;
  count_launch(), node_stat_kernel<<<ceil_div(n_rows * heads, 256), 256, 0, (cudaStream_t)stream>>>(s, (const float2*)rowstat, n_rows,
                                                                                                   row_offset, heads, (float4*)nodestat);
  B200GAT_LAUNCH_CHECK();
  return kOk;
}
extern "C" int b200gat_node_stat_set_t_f32(float* nodestat, const float* t, int64_t n, void* stream) {
  B200GAT_CHECK_ARG(nodestat && t, "null pointer");
  if (n == 0) return kOk;
  count_launch(), set_t_kernel<<<ceil_div(n, 256), 256, 0, (cudaStream_t)stream>>>((float4*)nodestat, t, n);
  B200GAT_LAUNCH_CHECK();
  return kOk;
}
// dst_bf16[k] = bf16(scale * src[k]), n a multiple of 4
extern "C" int b200gat_cast_bf16(const float* src, void* dst_bf16, int64_t n, float scale, void* stream) {
  B200GAT_CHECK_ARG(src && dst_bf16 && n % 4 == 0, "null pointer / n must be a multiple of 4");
  if (n == 0) return kOk;
  const int64_t n4 = n / 4;
  const int grid = (int)(ceil_div(n4, 256) < kNumSMs * 16 ? ceil_div(n4, 256) : kNumSMs * 16);
  count_launch(), cast_bf16_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(src, (__nv_bfloat16*)dst_bf16, n4, scale);
  B200GAT_LAUNCH_CHECK();
  return kOk;
}
