// Peer-memory exchange for the row-sharded path ("layer outputs are exchanged by an all-gather over
// NVLink").  Every rank keeps its block of rows in a buffer that its peers map into their own address space (CUDA IPC);
// an all-gather is then each rank PULLING the peers' blocks with the copy engines (cudaMemcpyAsync on peer pointers, one
// stream per peer), which keeps the SMs free and -- unlike the NCCL ring -- lets all seven NVLink-attached peers stream
// into a GPU at once.  Ordering is the caller's: a stream-ordered barrier (a one-element all-reduce) before the pulls
// makes sure every producer kernel has finished.
#include <string.h>

#include "common.cuh"
#include "../../include/b200gat.h"

using namespace b200gat;

extern "C" int b200gat_peer_alloc(size_t bytes, void** ptr) {
  B200GAT_CHECK_ARG(ptr && bytes > 0, "bad arguments");
  B200GAT_CUDA(cudaMalloc(ptr, bytes));   // a base allocation of its own: that is what an IPC handle can name
  return kOk;
}

extern "C" int b200gat_peer_free(void* ptr) {
  if (ptr) B200GAT_CUDA(cudaFree(ptr));
  return kOk;
}

extern "C" int b200gat_peer_export(const void* ptr, void* handle, size_t handle_bytes) {
  B200GAT_CHECK_ARG(ptr && handle, "null pointer");
  B200GAT_CHECK_ARG(handle_bytes >= sizeof(cudaIpcMemHandle_t), "handle buffer too small: %zu < %zu", handle_bytes,
                    sizeof(cudaIpcMemHandle_t));
  cudaIpcMemHandle_t h;
  B200GAT_CUDA(cudaIpcGetMemHandle(&h, const_cast<void*>(ptr)));
  memcpy(handle, &h, sizeof(h));
  return kOk;
}

extern "C" int b200gat_peer_open(const void* handle, void** ptr) {
  B200GAT_CHECK_ARG(handle && ptr, "null pointer");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle, sizeof(h));
  B200GAT_CUDA(cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess));
  return kOk;
}

extern "C" int b200gat_peer_close(void* ptr) {
  if (ptr) B200GAT_CUDA(cudaIpcCloseMemHandle(ptr));
  return kOk;
}

extern "C" int b200gat_peer_pull(void* dst, const void* src, size_t bytes, void* stream) {
  B200GAT_CHECK_ARG((dst && src) || bytes == 0, "null pointer");
  if (bytes == 0) return kOk;
  B200GAT_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
  return kOk;
}

// ================================================================================================================
// Device-side exchange over the peer-mapped buffers: no host round trip, no library collective in the step.
//
// Every rank owns ONE exported buffer with the same layout on all ranks ("fabric"): a block of flags followed by
// regions.  A gathered region holds `world` row blocks; rank r produces block r in place (its projection / node_prep /
// dx kernel writes straight into it) and pulls the other blocks out of its peers' copies of the same region.
//
//   signal(channel)      : after the producer kernels (stream order), one thread per peer stores this step's epoch into
//                          flags[channel][rank] of THAT PEER's buffer (release, system scope).
//   allgather / reduce   : the consuming kernel first waits until flags[channel][p] >= epoch for every p (its own flags,
//                          local memory), then reads the peers' blocks over NVLink with all SMs: consecutive 32 KB chunks
//                          go to different peers, so every link is busy from the first wave on.
//
// Why the data is there when the flag is: the producer kernel has completed (its writes are in the producer's L2, the
// coherence point of its memory) before the signal kernel starts; peers read that memory through NVLink into the same L2.
// Epochs only grow (one per step), so flags are never reset; a region is rewritten one step later, after at least one
// later exchange of the step has proven that every peer is past its pulls (sharded.py lists the argument per region).
// Waits are bounded by wall time and trap instead of hanging the GPU.
// ================================================================================================================
namespace b200gat {

constexpr int kMaxWorld = 8;
struct PeerTable { uint64_t base[kMaxWorld]; };
struct GatherPart { uint64_t off; uint64_t block_bytes; };

__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint64_t global_ns() {
  uint64_t t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ uint4 ld_peer16(const void* p) {   // read once, keep it out of L1
  uint4 v;
  asm volatile("ld.global.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  return v;
}

constexpr uint64_t kWaitLimitNs = 20ull * 1000 * 1000 * 1000;   // 20 s: a peer that never signals is a crashed peer

// block-wide: returns once every peer's flag of `channel` has reached `epoch`
__device__ __forceinline__ void wait_channel(const uint32_t* flags, int channel, int world, uint32_t epoch) {
  if ((int)threadIdx.x < world) {
    const uint32_t* f = flags + (size_t)channel * kMaxWorld + threadIdx.x;
    const uint64_t t0 = global_ns();
    while ((int32_t)(ld_acquire_sys(f) - epoch) < 0) {
      __nanosleep(200);
      if (global_ns() - t0 > kWaitLimitNs) __trap();
    }
  }
  __syncthreads();
}

__global__ void peer_signal_kernel(PeerTable t, int channel, int rank, int world, uint32_t epoch) {
  if ((int)threadIdx.x < world) {
    __threadfence_system();
    st_release_sys(reinterpret_cast<uint32_t*>(t.base[threadIdx.x]) + (size_t)channel * kMaxWorld + rank, epoch);
  }
}

__global__ void peer_wait_kernel(const uint32_t* flags, int channel, int world, uint32_t epoch) {
  wait_channel(flags, channel, world, epoch);
}

constexpr int kPullThreads = 512, kPullUnroll = 4;
constexpr uint64_t kPullChunk = (uint64_t)kPullThreads * kPullUnroll * 16;   // 32 KB

__global__ void __launch_bounds__(kPullThreads) peer_allgather_kernel(PeerTable t, GatherPart p0, GatherPart p1, int n_parts, int rank,
                                                                      int world, int channel, uint32_t epoch) {
  wait_channel(reinterpret_cast<const uint32_t*>(t.base[rank]), channel, world, epoch);
  const int others = world - 1;
  for (int part = 0; part < n_parts; ++part) {
    const GatherPart gp = part == 0 ? p0 : p1;
    const uint64_t chunks_per_block = (gp.block_bytes + kPullChunk - 1) / kPullChunk;
    const uint64_t n_chunks = chunks_per_block * others;
    for (uint64_t c = blockIdx.x; c < n_chunks; c += gridDim.x) {
      const int peer = (rank + 1 + (int)(c % others)) % world;          // neighbouring chunks -> different peers
      const uint64_t pos = (c / others) * kPullChunk;
      const uint64_t block_off = gp.off + (uint64_t)peer * gp.block_bytes;
      const char* src = reinterpret_cast<const char*>(t.base[peer]) + block_off;
      char* dst = reinterpret_cast<char*>(t.base[rank]) + block_off;
      uint4 v[kPullUnroll];
#pragma unroll
      for (int k = 0; k < kPullUnroll; ++k) {
        const uint64_t b = pos + ((uint64_t)k * kPullThreads + threadIdx.x) * 16;
        if (b < gp.block_bytes) v[k] = ld_peer16(src + b);
      }
#pragma unroll
      for (int k = 0; k < kPullUnroll; ++k) {
        const uint64_t b = pos + ((uint64_t)k * kPullThreads + threadIdx.x) * 16;
        if (b < gp.block_bytes) *reinterpret_cast<uint4*>(dst + b) = v[k];
      }
    }
  }
}

// Push variant of the gather: this rank reads ITS block once (local HBM) and stores it into the same place of every peer's
// buffer (posted NVLink writes).  Nothing to wait for before it starts -- the peers are done with that region, by the same
// argument that lets a producer overwrite its own block -- and the "my pushes have landed" signal is raised afterwards (the
// kernel boundary orders the stores before the signal kernel's fence + release store); the consumer waits on that channel.
__global__ void __launch_bounds__(kPullThreads) peer_push_kernel(PeerTable t, GatherPart p0, GatherPart p1, int n_parts, int rank, int world) {
  const int others = world - 1;
  for (int part = 0; part < n_parts; ++part) {
    const GatherPart gp = part == 0 ? p0 : p1;
    const uint64_t n_chunks = (gp.block_bytes + kPullChunk - 1) / kPullChunk;
    const uint64_t block_off = gp.off + (uint64_t)rank * gp.block_bytes;
    const char* src = reinterpret_cast<const char*>(t.base[rank]) + block_off;
    for (uint64_t c = blockIdx.x; c < n_chunks; c += gridDim.x) {
      const uint64_t pos = c * kPullChunk;
      uint4 v[kPullUnroll];
#pragma unroll
      for (int k = 0; k < kPullUnroll; ++k) {
        const uint64_t b = pos + ((uint64_t)k * kPullThreads + threadIdx.x) * 16;
        if (b < gp.block_bytes) v[k] = *reinterpret_cast<const uint4*>(src + b);
      }
      for (int q = 0; q < others; ++q) {
        const int peer = (rank + 1 + (int)((c + q) % others)) % world;      // neighbouring blocks start with different peers
        char* dst = reinterpret_cast<char*>(t.base[peer]) + block_off;
#pragma unroll
        for (int k = 0; k < kPullUnroll; ++k) {
          const uint64_t b = pos + ((uint64_t)k * kPullThreads + threadIdx.x) * 16;
          if (b < gp.block_bytes) *reinterpret_cast<uint4*>(dst + b) = v[k];
        }
      }
    }
  }
}

// out[i] = sum over ranks p = 0..world-1 (fixed order: every rank gets the same bits) of peer_p.region[first + i]
__global__ void __launch_bounds__(256) peer_reduce_kernel(PeerTable t, uint64_t off, int64_t first, int64_t n, float* __restrict__ out,
                                                          int rank, int world, int channel, uint32_t epoch) {
  wait_channel(reinterpret_cast<const uint32_t*>(t.base[rank]), channel, world, epoch);
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float a = 0.f;
    for (int p = 0; p < world; ++p) {
      const float* src = reinterpret_cast<const float*>(t.base[p] + off) + first + i;
      float v;
      asm volatile("ld.global.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(src));
      a += v;
    }
    out[i] = a;
  }
}

static int make_table(const void* const* bases, int world, PeerTable* t) {
  B200GAT_CHECK_ARG(bases && world >= 1 && world <= kMaxWorld, "world size %d outside [1, %d]", world, kMaxWorld);
  for (int r = 0; r < kMaxWorld; ++r) t->base[r] = r < world ? (uint64_t)bases[r] : 0;
  return kOk;
}

}  // namespace b200gat

extern "C" int b200gat_peer_flag_bytes(int n_channels, size_t* bytes) {
  B200GAT_CHECK_ARG(bytes && n_channels > 0, "bad arguments");
  *bytes = ((size_t)n_channels * kMaxWorld * 4 + 255) / 256 * 256;
  return kOk;
}

extern "C" int b200gat_peer_signal(const void* const* bases, int world, int rank, int channel, uint32_t epoch, void* stream) {
  PeerTable t;
  int rc = make_table(bases, world, &t);
  if (rc) return rc;
  count_launch(), peer_signal_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(t, channel, rank, world, epoch);
  B200GAT_LAUNCH_CHECK();
  return kOk;
}

extern "C" int b200gat_peer_wait(const void* const* bases, int world, int rank, int channel, uint32_t epoch, void* stream) {
  PeerTable t;
  int rc = make_table(bases, world, &t);
  if (rc) return rc;
  count_launch(), peer_wait_kernel<<<1, 32, 0, (cudaStream_t)stream>>>((const uint32_t*)t.base[rank], channel, world, epoch);
  B200GAT_LAUNCH_CHECK();
  return kOk;
}

extern "C" int b200gat_peer_allgather(const void* const* bases, int world, int rank, int channel, uint32_t epoch, int n_parts,
                                      const uint64_t* offsets, const uint64_t* block_bytes, void* stream) {
  PeerTable t;
  int rc = make_table(bases, world, &t);
  if (rc) return rc;
  B200GAT_CHECK_ARG(n_parts >= 1 && n_parts <= 2 && offsets && block_bytes, "1 or 2 parts per gather");
  GatherPart p[2] = {{0, 0}, {0, 0}};
  uint64_t total = 0;
  for (int k = 0; k < n_parts; ++k) {
    B200GAT_CHECK_ARG(offsets[k] % 16 == 0 && block_bytes[k] % 16 == 0, "gather parts must be 16-byte aligned");
    p[k].off = offsets[k];
    p[k].block_bytes = block_bytes[k];
    total += block_bytes[k];
  }
  if (world == 1 || total == 0) return b200gat_peer_wait(bases, world, rank, channel, epoch, stream);
  const uint64_t chunks = (total / kPullChunk + 2) * (uint64_t)(world - 1);
  const int grid = (int)(chunks < (uint64_t)kNumSMs * 4 ? chunks : (uint64_t)kNumSMs * 4);
  count_launch(), peer_allgather_kernel<<<grid, kPullThreads, 0, (cudaStream_t)stream>>>(t, p[0], p[1], n_parts, rank, world, channel, epoch);
  B200GAT_LAUNCH_CHECK();
  return kOk;
}

extern "C" int b200gat_peer_push(const void* const* bases, int world, int rank, int n_parts, const uint64_t* offsets,
                                 const uint64_t* block_bytes, void* stream) {
  PeerTable t;
  int rc = make_table(bases, world, &t);
  if (rc) return rc;
  B200GAT_CHECK_ARG(n_parts >= 1 && n_parts <= 2 && offsets && block_bytes, "1 or 2 parts per gather");
  GatherPart p[2] = {{0, 0}, {0, 0}};
  uint64_t total = 0;
  for (int k = 0; k < n_parts; ++k) {
    B200GAT_CHECK_ARG(offsets[k] % 16 == 0 && block_bytes[k] % 16 == 0, "gather parts must be 16-byte aligned");
    p[k].off = offsets[k];
    p[k].block_bytes = block_bytes[k];
    total += block_bytes[k];
  }
  if (world == 1 || total == 0) return kOk;
  const uint64_t chunks = total / kPullChunk + 2;
  const int grid = (int)(chunks < (uint64_t)kNumSMs * 4 ? chunks : (uint64_t)kNumSMs * 4);
  count_launch(), peer_push_kernel<<<grid, kPullThreads, 0, (cudaStream_t)stream>>>(t, p[0], p[1], n_parts, rank, world);
  B200GAT_LAUNCH_CHECK();
  return kOk;
}

extern "C" int b200gat_peer_reduce_f32(const void* const* bases, int world, int rank, int channel, uint32_t epoch, uint64_t offset,
                                       int64_t first, int64_t n, float* out, void* stream) {
  PeerTable t;
  int rc = make_table(bases, world, &t);
  if (rc) return rc;
  B200GAT_CHECK_ARG(out && n >= 0 && first >= 0 && offset % 4 == 0, "bad arguments");
  if (n == 0) return b200gat_peer_wait(bases, world, rank, channel, epoch, stream);
  const int64_t blocks = (n + 255) / 256;
  const int grid = (int)(blocks < (int64_t)kNumSMs * 8 ? blocks : (int64_t)kNumSMs * 8);
  count_launch(), peer_reduce_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(t, offset, first, n, out, rank, world, channel, epoch);
  B200GAT_LAUNCH_CHECK();
  return kOk;
}
