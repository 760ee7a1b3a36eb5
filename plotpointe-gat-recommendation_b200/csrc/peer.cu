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
