// Shared helpers for the b200gat CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <atomic>

namespace b200gat {

// ---- error plumbing -------------------------------------------------------------------------
// Every extern "C" entry point returns 0 on success or a negative code; the message of the last
// failure on this host thread is kept for b200gat_last_error().
enum : int { kOk = 0, kErrArg = -1, kErrCuda = -2, kErrUnsupported = -3, kErrWorkspace = -4 };

void set_error(const char* fmt, ...);
void count_launch();  // bumps the process-wide kernel launch counter (b200gat_launch_count)

// cudaFuncSetAttribute applies to the CURRENT device only, so "done once" has to be remembered per device.
// Usage: static DeviceOnce once; if (once.pending()) { ...set attributes...; once.done(); }   (setting twice is harmless)
struct DeviceOnce {
  std::atomic<uint64_t> mask{0};
  static uint64_t bit() {
    int dev = 0;
    cudaGetDevice(&dev);
    return 1ull << (dev & 63);
  }
  bool pending() const { return !(mask.load(std::memory_order_acquire) & bit()); }
  void done() { mask.fetch_or(bit(), std::memory_order_release); }
};

#define B200GAT_CHECK_ARG(cond, ...)                    \
  do {                                                  \
    if (!(cond)) {                                      \
      ::b200gat::set_error(__VA_ARGS__);                \
      return ::b200gat::kErrArg;                        \
    }                                                   \
  } while (0)

#define B200GAT_CUDA(call)                                                                   \
  do {                                                                                       \
    cudaError_t e__ = (call);                                                                \
    if (e__ != cudaSuccess) {                                                                \
      ::b200gat::set_error("%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); \
      return ::b200gat::kErrCuda;                                                            \
    }                                                                                        \
  } while (0)

#define B200GAT_LAUNCH_CHECK() B200GAT_CUDA(cudaGetLastError())

// ---- device helpers -------------------------------------------------------------------------
constexpr int kWarp = 32;
constexpr unsigned kFull = 0xffffffffu;
constexpr int kNumSMs = 148;  // B200: 2 dies x 74 SMs

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
  return v;
}
// Sums U (power of two <= 32) per-lane values across the warp at once.  A plain butterfly per value costs 5 shuffles each;
// here every level first halves the number of live values (a lane keeps one half and ships the other half to its
// partner), so the cost is (U - 1) + (5 - log2 U) shuffles in total.  On return, the total of value u is held by the
// lanes with (lane >> (5 - log2 U)) == u; v[] is clobbered.
template <int N, int OFF>
struct MultiReduce {
  template <int U>
  static __device__ __forceinline__ float run(float (&v)[U], int lane) {
    const bool hi = (lane & OFF) != 0;
#pragma unroll
    for (int t = 0; t < N / 2; ++t) {
      const float keep = hi ? v[N / 2 + t] : v[t];
      const float send = hi ? v[t] : v[N / 2 + t];
      v[t] = keep + __shfl_xor_sync(0xffffffffu, send, OFF);
    }
    return MultiReduce<N / 2, OFF / 2>::run(v, lane);
  }
};
template <int OFF>
struct MultiReduce<1, OFF> {
  template <int U>
  static __device__ __forceinline__ float run(float (&v)[U], int) {
    float x = v[0];
#pragma unroll
    for (int o = OFF; o >= 1; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
    return x;
  }
};
template <>
struct MultiReduce<1, 0> {
  template <int U>
  static __device__ __forceinline__ float run(float (&v)[U], int) { return v[0]; }
};
template <int U>
__device__ __forceinline__ float warp_multi_sum(float (&v)[U], int lane) {
  static_assert(U >= 1 && U <= 32 && (U & (U - 1)) == 0, "U must be a power of two <= 32");
  return MultiReduce<U, 16>::run(v, lane);
}
template <int U>
struct Log2 { static constexpr int value = 1 + Log2<U / 2>::value; };
template <>
struct Log2<1> { static constexpr int value = 0; };

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(kFull, v, o));
  return v;
}
__device__ __forceinline__ int warp_sum_i(int v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
  return v;
}

// 128-bit read-only gather load (feature rows are read many times: keep them in L1/L2).
__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
// streaming 128-bit load / store for data touched once (do not pollute L1)
__device__ __forceinline__ float4 ld_stream4(const float* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p));
  return r;
}
__device__ __forceinline__ void st_stream4(float* p, float4 v) {
  asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z),
               "f"(v.w)
               : "memory");
}

// "cache streaming" accesses: the line is allocated evict-first in L2, so data that is touched once per launch (index lists,
// the rows a kernel writes) does not push the randomly gathered feature rows out of L2
__device__ __forceinline__ void st_cs4(float* p, float4 v) {
  asm volatile("st.global.cs.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ void st_cs1(float* p, float v) { asm volatile("st.global.cs.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory"); }
__device__ __forceinline__ int ld_cs_i32(const int32_t* p) {
  int r;
  asm volatile("ld.global.cs.s32 %0, [%1];" : "=r"(r) : "l"(p));
  return r;
}
__device__ __forceinline__ float4 ld_cs4(const float* p) {
  float4 r;
  asm volatile("ld.global.cs.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ uint2 ld_cs2u(const void* p) {
  uint2 r;
  asm volatile("ld.global.cs.v2.u32 {%0,%1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(p));
  return r;
}

__device__ __forceinline__ float4 fma4(float a, float4 b, float4 c) {
  c.x = fmaf(a, b.x, c.x);
  c.y = fmaf(a, b.y, c.y);
  c.z = fmaf(a, b.z, c.z);
  c.w = fmaf(a, b.w, c.w);
  return c;
}
__device__ __forceinline__ float dot4(float4 a, float4 b) {
  return fmaf(a.x, b.x, fmaf(a.y, b.y, fmaf(a.z, b.z, a.w * b.w)));
}

inline int ceil_div(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

// ---- shared host-side building blocks (graph.cu) -------------------------------------------------
// Stable LSD radix sort of (key, iota) pairs by int32 key in [0, key_range); vals_out = the stable
// argsort.  workspace: sort_workspace_bytes(n).
size_t sort_workspace_bytes(int64_t n);
int sort_pairs_stable(const int32_t* keys, int64_t n, int64_t key_range, int32_t* keys_out, int32_t* vals_out,
                      void* workspace, cudaStream_t st);
// ptr[v] = first position in `sorted` with key >= v, for v in [0, n_nodes]
int node_ptr_from_sorted(const int32_t* sorted, int64_t n, int64_t n_nodes, int32_t* ptr, cudaStream_t st);

}  // namespace b200gat
