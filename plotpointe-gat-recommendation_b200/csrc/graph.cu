// Device-side COO -> CSR (by destination) / CSC (by source) builder.
//
// Replaces the edge ordering implied by the reference's per-edge scatter ops
// (scripts/train_gat_custom.py:85-92 on the edge list built at :166-175): the CSR order produced
// here is exactly ``dst.argsort(stable=True)`` -- duplicates kept, original intra-row order kept.
//
// Pipeline (all integer work, bit-exact and deterministic; integer atomics only on counters
// whose final value is order-independent):
//   1. narrow + validate   : int64 [2,E] -> int32 src/dst, count out-of-range ids
//   2. LSD radix sort       : stable, 8-bit digits, key = node id, value = edge id
//        per pass: tile histogram -> device-wide exclusive scan (digit-major) -> stable scatter
//   3. rowptr / colptr      : lower_bound of every node id in the sorted keys (no atomics)
//   4. col / row gathers, CSR<->CSC position maps
#include "common.cuh"
#include "../../include/b200gat.h"

namespace b200gat {

// --------------------------------------------------------------------------------------------
// 1. narrow + validate
// --------------------------------------------------------------------------------------------
__global__ void narrow_validate_kernel(const int64_t* __restrict__ ei, int64_t n_edges, int64_t n_nodes,
                                       int32_t* __restrict__ src32, int32_t* __restrict__ dst32,
                                       int32_t* __restrict__ n_bad) {
  int bad = 0;
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < n_edges; e += (int64_t)gridDim.x * blockDim.x) {
    int64_t s = ei[e], d = ei[n_edges + e];
    bool ok = (s >= 0) & (s < n_nodes) & (d >= 0) & (d < n_nodes);
    bad += !ok;
    src32[e] = ok ? (int32_t)s : 0;
    dst32[e] = ok ? (int32_t)d : 0;
  }
  bad = warp_sum_i(bad);
  if ((threadIdx.x & 31) == 0 && bad) atomicAdd(n_bad, bad);
}

// --------------------------------------------------------------------------------------------
// device-wide exclusive scan of int32 (three launches, arbitrary length)
// --------------------------------------------------------------------------------------------
constexpr int kScanThreads = 256;
constexpr int kScanItems = 16;
constexpr int kScanTile = kScanThreads * kScanItems;  // 4096

__device__ __forceinline__ int block_exclusive_scan(int v, int* smem /*[kScanThreads/32]*/, int& total) {
  // inclusive warp scan
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  int x = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int y = __shfl_up_sync(kFull, x, o);
    if (lane >= o) x += y;
  }
  if (lane == 31) smem[w] = x;
  __syncthreads();
  if (w == 0) {
    int s = lane < kScanThreads / 32 ? smem[lane] : 0;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int y = __shfl_up_sync(kFull, s, o);
      if (lane >= o) s += y;
    }
    if (lane < kScanThreads / 32) smem[lane] = s;  // inclusive over warps
  }
  __syncthreads();
  int warp_prefix = w ? smem[w - 1] : 0;
  total = smem[kScanThreads / 32 - 1];
  __syncthreads();
  return warp_prefix + x - v;
}

__global__ void __launch_bounds__(kScanThreads) scan_tile_sums_kernel(const int32_t* __restrict__ in, int64_t n,
                                                                      int32_t* __restrict__ tile_sums) {
  __shared__ int sm[kScanThreads / 32];
  int64_t base = (int64_t)blockIdx.x * kScanTile + (int64_t)threadIdx.x * kScanItems;
  int s = 0;
#pragma unroll
  for (int k = 0; k < kScanItems; ++k) s += (base + k < n) ? in[base + k] : 0;
  int total;
  block_exclusive_scan(s, sm, total);
  if (threadIdx.x == 0) tile_sums[blockIdx.x] = total;
}

// single block: in-place exclusive scan of the tile sums (sequential over chunks of 256)
__global__ void __launch_bounds__(kScanThreads) scan_sums_kernel(int32_t* __restrict__ sums, int n) {
  __shared__ int sm[kScanThreads / 32];
  int carry = 0;
  for (int base = 0; base < n; base += kScanThreads) {
    int i = base + threadIdx.x;
    int v = i < n ? sums[i] : 0;
    int total;
    int ex = block_exclusive_scan(v, sm, total);
    if (i < n) sums[i] = carry + ex;
    carry += total;
  }
}

__global__ void __launch_bounds__(kScanThreads) scan_apply_kernel(const int32_t* in, int64_t n,  // in/out may alias
                                                                  const int32_t* __restrict__ tile_prefix,
                                                                  int32_t* out) {
  __shared__ int sm[kScanThreads / 32];
  int64_t base = (int64_t)blockIdx.x * kScanTile + (int64_t)threadIdx.x * kScanItems;
  int v[kScanItems];
  int s = 0;
#pragma unroll
  for (int k = 0; k < kScanItems; ++k) {
    v[k] = (base + k < n) ? in[base + k] : 0;
    s += v[k];
  }
  int total;
  int ex = block_exclusive_scan(s, sm, total) + tile_prefix[blockIdx.x];
#pragma unroll
  for (int k = 0; k < kScanItems; ++k) {
    if (base + k < n) out[base + k] = ex;
    ex += v[k];
  }
}

static size_t scan_ws_ints(int64_t n) { return (size_t)ceil_div(n, kScanTile) + 1; }

// in/out may alias. ws: scan_ws_ints(n) int32.
static int exclusive_scan_i32(const int32_t* in, int32_t* out, int64_t n, int32_t* ws, cudaStream_t st) {
  if (n == 0) return kOk;
  int tiles = ceil_div(n, kScanTile);
  count_launch(), scan_tile_sums_kernel<<<tiles, kScanThreads, 0, st>>>(in, n, ws);
  count_launch(), scan_sums_kernel<<<1, kScanThreads, 0, st>>>(ws, tiles);
  count_launch(), scan_apply_kernel<<<tiles, kScanThreads, 0, st>>>(in, n, ws, out);
  B200GAT_LAUNCH_CHECK();
  return kOk;
}

// --------------------------------------------------------------------------------------------
// 2. stable LSD radix sort of (key = node id, value = edge id)
// --------------------------------------------------------------------------------------------
constexpr int kSortThreads = 256;
constexpr int kSortWarps = kSortThreads / 32;
constexpr int kSortIters = 8;                            // keys per lane
constexpr int kSortTile = kSortThreads * kSortIters;     // 2048 keys per block
constexpr int kRadix = 256;

// Tile element order (defines stability): warp w owns keys [w*256, (w+1)*256) of the tile and
// visits them 32 at a time, lane-major.
__device__ __forceinline__ int64_t sort_index(int64_t tile_base, int w, int it, int lane) {
  return tile_base + w * (kSortIters * 32) + it * 32 + lane;
}

__global__ void __launch_bounds__(kSortThreads) radix_hist_kernel(const int32_t* __restrict__ keys, int64_t n, int shift,
                                                                  int n_tiles, int32_t* __restrict__ table) {
  __shared__ int hist[kRadix];
  hist[threadIdx.x] = 0;
  __syncthreads();
  int64_t tile_base = (int64_t)blockIdx.x * kSortTile;
  int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int it = 0; it < kSortIters; ++it) {
    int64_t idx = sort_index(tile_base, w, it, lane);
    if (idx < n) atomicAdd(&hist[(keys[idx] >> shift) & (kRadix - 1)], 1);
  }
  __syncthreads();
  table[(int64_t)threadIdx.x * n_tiles + blockIdx.x] = hist[threadIdx.x];  // digit-major
}

__global__ void __launch_bounds__(kSortThreads) radix_scatter_kernel(const int32_t* __restrict__ keys_in,
                                                                     const int32_t* __restrict__ vals_in,  // null: iota
                                                                     int64_t n, int shift, int n_tiles,
                                                                     const int32_t* __restrict__ table_scanned,
                                                                     int32_t* __restrict__ keys_out,
                                                                     int32_t* __restrict__ vals_out) {
  __shared__ int cnt[kSortWarps][kRadix + 1];  // +1: bin for out-of-range lanes
  for (int i = threadIdx.x; i < kSortWarps * (kRadix + 1); i += kSortThreads) (&cnt[0][0])[i] = 0;
  __syncthreads();
  int64_t tile_base = (int64_t)blockIdx.x * kSortTile;
  int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  unsigned lt_mask = (1u << lane) - 1u;
  int key[kSortIters], val[kSortIters], rank[kSortIters];
#pragma unroll
  for (int it = 0; it < kSortIters; ++it) {
    int64_t idx = sort_index(tile_base, w, it, lane);
    bool ok = idx < n;
    key[it] = ok ? keys_in[idx] : 0;
    val[it] = ok ? (vals_in ? vals_in[idx] : (int32_t)idx) : 0;
    int d = ok ? ((key[it] >> shift) & (kRadix - 1)) : kRadix;
    unsigned peers = __match_any_sync(kFull, d);
    int leader = __ffs(peers) - 1;
    int before = 0;
    if (lane == leader) {
      before = cnt[w][d];
      cnt[w][d] = before + __popc(peers);
    }
    before = __shfl_sync(kFull, before, leader);
    rank[it] = before + __popc(peers & lt_mask);  // stable: earlier lane / earlier iteration first
    __syncwarp();
  }
  __syncthreads();
  {  // one thread per digit: turn per-warp counts into global start offsets
    int d = threadIdx.x;
    int run = table_scanned[(int64_t)d * n_tiles + blockIdx.x];
#pragma unroll
    for (int ww = 0; ww < kSortWarps; ++ww) {
      int c = cnt[ww][d];
      cnt[ww][d] = run;
      run += c;
    }
  }
  __syncthreads();
#pragma unroll
  for (int it = 0; it < kSortIters; ++it) {
    int64_t idx = sort_index(tile_base, w, it, lane);
    if (idx < n) {
      int d = (key[it] >> shift) & (kRadix - 1);
      int pos = cnt[w][d] + rank[it];
      keys_out[pos] = key[it];
      vals_out[pos] = val[it];
    }
  }
}

static int key_bits(int64_t n_nodes) {
  int bits = 1;
  while (((int64_t)1 << bits) < n_nodes) ++bits;
  return bits;
}

struct SortWs {
  int32_t *k0, *v0, *k1, *v1, *table, *scan;
};
static inline size_t align_up(size_t x, size_t a = 256) { return (x + a - 1) / a * a; }

// Sort (keys, iota) by key, stable. Result in keys_out / vals_out. tmp buffers each n int32.
static int radix_sort_by_key(const int32_t* keys, int64_t n, int64_t n_nodes, int32_t* keys_out, int32_t* vals_out,
                             const SortWs& ws, cudaStream_t st) {
  if (n == 0) return kOk;
  int passes = (key_bits(n_nodes) + 7) / 8;
  int n_tiles = ceil_div(n, kSortTile);
  const int32_t* kin = keys;
  const int32_t* vin = nullptr;
  for (int p = 0; p < passes; ++p) {
    bool last = p == passes - 1;
    int32_t* ko = last ? keys_out : ((p & 1) ? ws.k1 : ws.k0);
    int32_t* vo = last ? vals_out : ((p & 1) ? ws.v1 : ws.v0);
    count_launch(), radix_hist_kernel<<<n_tiles, kSortThreads, 0, st>>>(kin, n, 8 * p, n_tiles, ws.table);
    int rc = exclusive_scan_i32(ws.table, ws.table, (int64_t)kRadix * n_tiles, ws.scan, st);
    if (rc) return rc;
    count_launch(), radix_scatter_kernel<<<n_tiles, kSortThreads, 0, st>>>(kin, vin, n, 8 * p, n_tiles, ws.table, ko, vo);
    B200GAT_LAUNCH_CHECK();
    kin = ko;
    vin = vo;
  }
  return kOk;
}

// --------------------------------------------------------------------------------------------
// 3./4. pointers, gathers, maps
// --------------------------------------------------------------------------------------------
__global__ void lower_bound_ptr_kernel(const int32_t* __restrict__ sorted_keys, int64_t n, int64_t n_nodes,
                                       int32_t* __restrict__ ptr /*[n_nodes+1]*/) {
  int64_t v = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (v > n_nodes) return;
  int64_t lo = 0, hi = n;  // first position with key >= v
  while (lo < hi) {
    int64_t mid = (lo + hi) >> 1;
    if (sorted_keys[mid] < v) lo = mid + 1; else hi = mid;
  }
  ptr[v] = (int32_t)lo;
}

__global__ void gather_i32_kernel(const int32_t* __restrict__ table, const int32_t* __restrict__ idx, int64_t n,
                                  int32_t* __restrict__ out) {
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x)
    out[e] = table[idx[e]];
}

__global__ void invert_perm_kernel(const int32_t* __restrict__ perm, int64_t n, int32_t* __restrict__ inv) {
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x)
    inv[perm[e]] = (int32_t)e;
}

// ---- entry points shared with loss.cu (declared in common.cuh) -----------------------------------
size_t sort_workspace_bytes(int64_t n) {
  size_t e = align_up((size_t)(n > 0 ? n : 1) * 4);
  int n_tiles = ceil_div(n, kSortTile) + 1;
  return 4 * e + align_up((size_t)kRadix * n_tiles * 4) + align_up(scan_ws_ints((int64_t)kRadix * n_tiles) * 4);
}

int sort_pairs_stable(const int32_t* keys, int64_t n, int64_t key_range, int32_t* keys_out, int32_t* vals_out,
                      void* workspace, cudaStream_t st) {
  char* p = (char*)workspace;
  size_t e = align_up((size_t)(n > 0 ? n : 1) * 4);
  int n_tiles = ceil_div(n, kSortTile) + 1;
  auto take = [&](size_t b) { char* q = p; p += b; return (int32_t*)q; };
  SortWs ws;
  ws.k0 = take(e); ws.v0 = take(e); ws.k1 = take(e); ws.v1 = take(e);
  ws.table = take(align_up((size_t)kRadix * n_tiles * 4));
  ws.scan = take(align_up(scan_ws_ints((int64_t)kRadix * n_tiles) * 4));
  return radix_sort_by_key(keys, n, key_range, keys_out, vals_out, ws, st);
}

int node_ptr_from_sorted(const int32_t* sorted, int64_t n, int64_t n_nodes, int32_t* ptr, cudaStream_t st) {
  count_launch(), lower_bound_ptr_kernel<<<ceil_div(n_nodes + 1, 256), 256, 0, st>>>(sorted, n, n_nodes, ptr);
  B200GAT_LAUNCH_CHECK();
  return kOk;
}

}  // namespace b200gat

using namespace b200gat;

extern "C" int b200gat_graph_workspace_bytes(int64_t n_nodes, int64_t n_edges, size_t* bytes) {
  B200GAT_CHECK_ARG(bytes != nullptr, "bytes is null");
  B200GAT_CHECK_ARG(n_nodes >= 0 && n_edges >= 0, "negative sizes");
  B200GAT_CHECK_ARG(n_nodes < 2147483647LL && n_edges < 2147483647LL,
                    "int32 index build: n_nodes=%lld n_edges=%lld must be < 2^31", (long long)n_nodes,
                    (long long)n_edges);
  size_t e = align_up((size_t)(n_edges > 0 ? n_edges : 1) * 4);
  int n_tiles = ceil_div(n_edges, kSortTile) + 1;
  size_t table = align_up((size_t)kRadix * n_tiles * 4);
  size_t scan = align_up(scan_ws_ints((int64_t)kRadix * n_tiles) * 4);
  // src32, dst32, k0, v0, k1, v1, sorted_keys, inv  + table + scan + status
  *bytes = 8 * e + table + scan + 256;
  return kOk;
}

extern "C" int b200gat_build_graph(const int64_t* edge_index, int64_t n_edges, int64_t n_nodes, int32_t* rowptr,
                                   int32_t* col, int32_t* perm, int32_t* colptr, int32_t* row, int32_t* perm_csc,
                                   int32_t* csr2csc, int32_t* n_bad_out, void* workspace, size_t workspace_bytes,
                                   void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  size_t need = 0;
  int rc = b200gat_graph_workspace_bytes(n_nodes, n_edges, &need);
  if (rc) return rc;
  B200GAT_CHECK_ARG(workspace_bytes >= need, "workspace too small: %zu < %zu", workspace_bytes, need);
  B200GAT_CHECK_ARG(rowptr && colptr && n_bad_out && workspace, "null output pointer");
  B200GAT_CHECK_ARG(n_edges == 0 || (edge_index && col && perm && row && perm_csc && csr2csc), "null edge array");

  char* p = (char*)workspace;
  size_t e = align_up((size_t)(n_edges > 0 ? n_edges : 1) * 4);
  int n_tiles = ceil_div(n_edges, kSortTile) + 1;
  auto take = [&](size_t b) { char* q = p; p += b; return (int32_t*)q; };
  int32_t* src32 = take(e);
  int32_t* dst32 = take(e);
  SortWs ws;
  ws.k0 = take(e); ws.v0 = take(e); ws.k1 = take(e); ws.v1 = take(e);
  int32_t* sorted = take(e);
  int32_t* inv = take(e);
  ws.table = take(align_up((size_t)kRadix * n_tiles * 4));
  ws.scan = take(align_up(scan_ws_ints((int64_t)kRadix * n_tiles) * 4));

  B200GAT_CUDA(cudaMemsetAsync(n_bad_out, 0, sizeof(int32_t), st));
  const int T = 256;
  int grid = n_edges ? min(ceil_div(n_edges, T), kNumSMs * 8) : 1;
  if (n_edges) count_launch(), narrow_validate_kernel<<<grid, T, 0, st>>>(edge_index, n_edges, n_nodes, src32, dst32, n_bad_out);
  B200GAT_LAUNCH_CHECK();

  // CSR: stable by destination
  rc = radix_sort_by_key(dst32, n_edges, n_nodes, sorted, perm, ws, st);
  if (rc) return rc;
  count_launch(), lower_bound_ptr_kernel<<<ceil_div(n_nodes + 1, T), T, 0, st>>>(sorted, n_edges, n_nodes, rowptr);
  if (n_edges) count_launch(), gather_i32_kernel<<<grid, T, 0, st>>>(src32, perm, n_edges, col);
  B200GAT_LAUNCH_CHECK();

  // CSC: stable by source
  rc = radix_sort_by_key(src32, n_edges, n_nodes, sorted, perm_csc, ws, st);
  if (rc) return rc;
  count_launch(), lower_bound_ptr_kernel<<<ceil_div(n_nodes + 1, T), T, 0, st>>>(sorted, n_edges, n_nodes, colptr);
  if (n_edges) {
    count_launch(), gather_i32_kernel<<<grid, T, 0, st>>>(dst32, perm_csc, n_edges, row);
    count_launch(), invert_perm_kernel<<<grid, T, 0, st>>>(perm_csc, n_edges, inv);      // inv[edge id] = CSC position
    count_launch(), gather_i32_kernel<<<grid, T, 0, st>>>(inv, perm, n_edges, csr2csc);   // CSR position -> CSC position
  }
  B200GAT_LAUNCH_CHECK();
  return kOk;
}
